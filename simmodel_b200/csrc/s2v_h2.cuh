// fp32-accurate tensor-core arithmetic on TWO fp16 pieces per operand with exact power-of-two scales (r02d; DESIGN.md 4f,
// profiles/r02d_fp16x2.md): the kernels the tensor path runs by default.
//   k_embed_round_bwd_h2   backward round      du^{t-1} = (g theta2) [mu^{t-1} > 0],  dtheta2 += g^T mu^{t-1}
//   k_embed_round_h2       forward round       mu^t = relu(base + (sum_nbr mu^{t-1}) theta2^T)
//   k_embed_final_bwd_h2   last backward stage D = sum_t du^t, dtheta1b, dh, dtheta1a, column sums
//   k_head_h2              Q / logit head      q_i = w5 . relu([theta6 pool_g ; theta7 mu_i]) + b5
// Included by s2v_embed_tc.cu inside its anonymous namespace (uses its gather helpers, tile constants, tc:: wrappers).
#pragma once

constexpr uint32_t TILE16X2 = 2 * TILE16;

__device__ __forceinline__ float pow2_scale_for(float amax) {      // 2^k with amax 2^k in [2^14, 2^15); 2^127 for (near-)zero tiles
    const int e = (__float_as_int(amax) >> 23) & 255;
    int k = 141 - e;
    k = k > 127 ? 127 : (k < -126 ? -126 : k);
    return __int_as_float((k + 127) << 23);
}
__device__ __forceinline__ void split2h(const float4& v, float s, uint2& p1, uint2& p2) {
    const float x = v.x * s, y = v.y * s, z = v.z * s, w = v.w * s;
    const __half2 a = __floats2half2_rn(x, y), b = __floats2half2_rn(z, w);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    const __half2 c = __floats2half2_rn(x - fa.x, y - fa.y), d = __floats2half2_rn(z - fb.x, w - fb.y);
    p1.x = *reinterpret_cast<const uint32_t*>(&a); p1.y = *reinterpret_cast<const uint32_t*>(&b);
    p2.x = *reinterpret_cast<const uint32_t*>(&c); p2.y = *reinterpret_cast<const uint32_t*>(&d);
}
__device__ __forceinline__ void store_split2h(uint8_t* t1, int r, int c4, const float4& v, float s) {
    uint2 p1, p2;
    split2h(v, s, p1, p2);
    const uint32_t off = tc::tile_off16(r, c4);
    *reinterpret_cast<uint2*>(t1 + off) = p1;
    *reinterpret_cast<uint2*>(t1 + TILE16 + off) = p2;
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the three products with the smallest first: (h2, h1'), (h1, h2'), (h1, h1')
__device__ __forceinline__ void issue_transform_h2(uint32_t d_tmem, uint32_t a1, uint32_t b1) {
    constexpr uint32_t idesc = idesc_f16(64, 64, 0, 0);
    const int ii[3] = {1, 0, 0}, jj[3] = {0, 1, 0};
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_bf16(d_tmem, tc::desc_k16(a1 + ii[p] * TILE16, ks), tc::desc_k16(b1 + jj[p] * TILE16, ks), idesc, p > 0 || ks > 0);
}
__device__ __forceinline__ void issue_reduce_h2(uint32_t d_tmem, uint32_t a1, uint32_t b1) {
    constexpr uint32_t idesc = idesc_f16(64, 64, 1, 1);
    const int ii[3] = {1, 0, 0}, jj[3] = {0, 1, 0};
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_bf16(d_tmem, tc::desc_mn16(a1 + ii[p] * TILE16, ks, TILE16), tc::desc_mn16(b1 + jj[p] * TILE16, ks, TILE16), idesc,
                         p > 0 || ks > 0);
}
// swizzled fp32 staging tile [64][64]: 16-byte chunk c of row r lives at chunk c ^ (r & 15)
__device__ __forceinline__ uint32_t stg_off(int r, int chunk) { return (uint32_t)(r * 64 + ((chunk ^ (r & 15)) << 2)); }


// ------------------------------------------------------------------------------------
// The backward round on TWO fp16 pieces per operand with per-tile power-of-two scaling (r02d, k_embed_round_bwd_h2).
// bf16x3 costs three 8 KB pieces per operand tile and six products per contraction; fp16 carries 11 significand bits,
// so x s = h1 + h2 (h1 = RN16(x s), h2 = RN16(x s - h1)) is exact to 2^-22 |x| and three products (h2 h1', h1 h2',
// h1 h1') give the same ~2^-22 -- IF nothing underflows: fp16 has 5 exponent bits, and gradient tiles live at 1e-8.
// Every operand tile is therefore scaled by a power of two (exact) that puts its largest magnitude in [2^14, 2^15):
// elements down to 2^-18 of the tile's largest keep all 22 bits (h2 stays normal), below that the error is the fp16
// subnormal spacing, 2^-40 of the largest -- the absolute error of a tile is <= 2^-22 of its largest entry, the
// quantity the parity bound is stated in.  The scales leave through the epilogue (D1 2^-(eg + ew)) and the read-out of
// D2 (per tile, 2^-(eg + em): D2 is never chained across tiles, so the tensor core's truncating accumulator sees 12 MMAs).
// What it buys on chip: 16 KB instead of 24 KB per operand tile, 24 MMAs instead of 48 per tile, a third fewer split
// instructions and shared-memory stores: 1.24 -> 1.0x ms in the otherwise identical 4-warp kernel.  (Running the MMAs
// of tile k behind the gather of tile k+1 on a second g stage -- 64 KB, still three CTAs -- measured 1.33 - 1.37 ms.)
// Staging is a 16 KB swizzled [64][64] fp32 tile in the dead g tile.
// ------------------------------------------------------------------------------------
struct H2Shared {
    uint64_t bar;
    uint32_t tmem_base;
    uint32_t pad;
    float wmax[8];           // per warp: max |g|, max |m| of the tile being produced
    float wscale;
};

template <int CTAS>
__global__ void __launch_bounds__(TC_THREADS, CTAS)
k_embed_round_bwd_h2(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ du_t,
                     const float* __restrict__ mu_prev, float* __restrict__ du_prev, float* __restrict__ partial,
                     int partial_stride, int accumulate, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta2^T 2^ew: tile(row k, col n) = theta2[n][k], two fp16 pieces
    uint8_t* a1 = w1 + TILE16X2;                 // gathered du^t tile 2^eg; reused as the staging tile
    uint8_t* m1 = a1 + TILE16X2;                 // own mu^{t-1} tile 2^em
    H2Shared* sh = reinterpret_cast<H2Shared*>(m1 + TILE16X2);
    float* stg = reinterpret_cast<float*>(a1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const unsigned FULL = 0xffffffffu;

    if (warp == 0) tc::tmem_alloc<128>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 2); tc::fence_barrier_init(); }
    {   // weights: largest magnitude -> scale -> two scaled fp16 pieces (transposed)
        float wm = 0.f;
        for (int e = t; e < 64 * 64; e += TC_THREADS) wm = fmaxf(wm, fabsf(__ldg(theta2_w + e)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(FULL, wm, o));
        if (lane == 0) sh->wmax[warp] = wm;
        __syncthreads();
        if (t == 0) sh->wscale = pow2_scale_for(fmaxf(fmaxf(sh->wmax[0], sh->wmax[1]), fmaxf(sh->wmax[2], sh->wmax[3])));
        __syncthreads();
        const float sw = sh->wscale;
        for (int e = t; e < 64 * 16; e += TC_THREADS) {
            const int row = e >> 4, cc = e & 15;
            const float4 v = make_float4(__ldg(theta2_w + (4 * cc) * 64 + row), __ldg(theta2_w + (4 * cc + 1) * 64 + row),
                                         __ldg(theta2_w + (4 * cc + 2) * 64 + row), __ldg(theta2_w + (4 * cc + 3) * 64 + row));
            store_split2h(w1, row, cc, v, sw);
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const float inv_sw = 1.0f / sh->wscale;
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1), s_m = tc::smem_u32(m1);
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    uint32_t phase = 0;
    const int c4 = lane & 15, rl0 = warp * 16 + 8 * (lane >> 4);
    float racc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) racc[i] = 0.f;
    Gather8Meta meta;
    if (blockIdx.x < num_tiles)
        gather8_prefetch(meta, g, g.rowptr_t, g.col_t, blockIdx.x * TCR, min(TCR, g.num_nodes - blockIdx.x * TCR), rl0);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        uint32_t mbits = 0;                      // [mu^{t-1} > 0] of this lane's 8 rows x 4 columns
        float inv_g, inv_m;
        {
            float4 mv[8], gv[8];                 // own rows first: they are in flight together with the gather
#pragma unroll
            for (int i = 0; i < 8; ++i)
                mv[i] = rl0 + i < rows ? ldg4(mu_prev + (size_t)(row0 + rl0 + i) * 64 + 4 * c4) : f4zero();
            gather8_consume(meta, g.col_t, du_t, rl0, [&](int rl, const float4& acc) { gv[rl - rl0] = acc; });
            float gm = 0.f, mm = 0.f;            // the tile's largest magnitudes: per lane, per warp, per CTA
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                gm = fmaxf(gm, fmaxf(fmaxf(fabsf(gv[i].x), fabsf(gv[i].y)), fmaxf(fabsf(gv[i].z), fabsf(gv[i].w))));
                mm = fmaxf(mm, fmaxf(fmaxf(fabsf(mv[i].x), fabsf(mv[i].y)), fmaxf(fabsf(mv[i].z), fabsf(mv[i].w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { gm = fmaxf(gm, __shfl_xor_sync(FULL, gm, o)); mm = fmaxf(mm, __shfl_xor_sync(FULL, mm, o)); }
            if (lane == 0) { sh->wmax[warp] = gm; sh->wmax[4 + warp] = mm; }
            __syncthreads();
            const float sg = pow2_scale_for(fmaxf(fmaxf(sh->wmax[0], sh->wmax[1]), fmaxf(sh->wmax[2], sh->wmax[3])));
            const float sm = pow2_scale_for(fmaxf(fmaxf(sh->wmax[4], sh->wmax[5]), fmaxf(sh->wmax[6], sh->wmax[7])));
            inv_g = 1.0f / sg; inv_m = 1.0f / sm;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                store_split2h(a1, rl0 + i, c4, gv[i], sg);
                store_split2h(m1, rl0 + i, c4, mv[i], sm);
                mbits |= (mv[i].x > 0.f ? 1u : 0u) << (4 * i) | (mv[i].y > 0.f ? 2u : 0u) << (4 * i) |
                         (mv[i].z > 0.f ? 4u : 0u) << (4 * i) | (mv[i].w > 0.f ? 8u : 0u) << (4 * i);
            }
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {                         // two issuing warps, one per accumulator (the order inside each stays fixed)
            tc::tc_fence_after_sync();
            issue_transform_h2(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        } else if (warp == 2) {
            tc::tc_fence_after_sync();
            issue_reduce_h2(d2, s_a, s_m);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        {
            const int nt = tile + gridDim.x;     // CSR metadata of the next tile: overlaps MMA + epilogue
            if (nt < num_tiles) gather8_prefetch(meta, g, g.rowptr_t, g.col_t, nt * TCR, min(TCR, g.num_nodes - nt * TCR), rl0);
        }
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        {
            const float c1 = inv_g * inv_sw;    // the g pieces are dead once the MMAs have completed: D1 -> staging
            float v[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(d1 + lane_sel + 32 * half, v);
                if (lane < 16) {
                    const int r = 16 * warp + lane;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        st4(stg + stg_off(r, 8 * half + i), make_float4(v[4 * i] * c1, v[4 * i + 1] * c1, v[4 * i + 2] * c1, v[4 * i + 3] * c1));
                }
            }
            // D2 of this tile alone (its scales are its own): lanes 0-15 own the TMEM lanes M = 64 uses and keep columns 0-31,
            // lanes 16-31 take columns 32-63 by shuffle; the long sum is fp32 adds in tile order
            const float c2 = inv_g * inv_m;
            tc::tmem_ld32(d2 + lane_sel, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (lane < 16) racc[i] = fmaf(v[i], c2, racc[i]);
            tc::tmem_ld32(d2 + lane_sel + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float o = __shfl_sync(FULL, v[i], lane & 15);
                if (lane >= 16) racc[i] = fmaf(o, c2, racc[i]);
            }
        }
        tc::tc_fence_before_sync();
        __syncwarp();                            // the warp reads back the staging rows it wrote itself (rows 16 w .. 16 w + 15)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = rl0 + i;
            if (rl < rows) {
                float4 v = ld4(stg + stg_off(rl, c4));
                const uint32_t b = mbits >> (4 * i);
                st4(du_prev + (size_t)(row0 + rl) * 64 + 4 * c4,
                    make_float4(b & 1u ? v.x : 0.f, b & 2u ? v.y : 0.f, b & 4u ? v.z : 0.f, b & 8u ? v.w : 0.f));
            }
        }
        // no barrier here: the next tile's pieces are stored behind ITS scale-exchange barrier, which every thread reaches
        // only after this epilogue
    }
    {   // thread (warp, lane) holds dtheta2[n = 16 warp + lane % 16][32 (lane / 16) .. + 32)
        float* my = partial + (size_t)blockIdx.x * partial_stride + (16 * warp + (lane & 15)) * 64 + 32 * (lane >> 4);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 x4 = make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]);
            if (accumulate) x4 = f4add(ld4(my + i), x4);
            st4(my + i, x4);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(sh->tmem_base);
}

// ------------------------------------------------------------------------------------
// Forward round on two fp16 pieces per operand (same arithmetic scheme as k_embed_round_bwd_h2): 32 KB of tiles and 12
// MMAs per tile instead of 48 KB and 24; the gathered tile is scaled by the power of two that puts its largest entry in
// [2^14, 2^15), theta2 once per CTA, both leave through the epilogue.
// ------------------------------------------------------------------------------------
template <int CTAS>
__global__ void __launch_bounds__(TC_THREADS, CTAS)
k_embed_round_h2(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ base,
                 const float* __restrict__ mu_prev, float* __restrict__ mu_next, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta2 [n][k] 2^ew, two fp16 pieces
    uint8_t* a1 = w1 + TILE16X2;                 // gathered tile 2^ea, two fp16 pieces; reused as the staging tile
    H2Shared* sh = reinterpret_cast<H2Shared*>(a1 + TILE16X2);
    float* stg = reinterpret_cast<float*>(a1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const unsigned FULL = 0xffffffffu;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    {
        float wm = 0.f;
        for (int e = t; e < 64 * 64; e += TC_THREADS) wm = fmaxf(wm, fabsf(__ldg(theta2_w + e)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(FULL, wm, o));
        if (lane == 0) sh->wmax[warp] = wm;
        __syncthreads();
        if (t == 0) sh->wscale = pow2_scale_for(fmaxf(fmaxf(sh->wmax[0], sh->wmax[1]), fmaxf(sh->wmax[2], sh->wmax[3])));
        __syncthreads();
        const float sw = sh->wscale;
        for (int e = t; e < 64 * 16; e += TC_THREADS) store_split2h(w1, e >> 4, e & 15, ldg4(theta2_w + (e >> 4) * 64 + 4 * (e & 15)), sw);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const float inv_sw = 1.0f / sh->wscale;
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1);
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    uint32_t phase = 0;
    const int c4 = lane & 15, rl0 = warp * 16 + 8 * (lane >> 4);
    Gather8Meta meta;
    if (blockIdx.x < num_tiles)
        gather8_prefetch(meta, g, g.rowptr, g.col, blockIdx.x * TCR, min(TCR, g.num_nodes - blockIdx.x * TCR), rl0);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        float c1;
        {
            float4 gv[8];
            gather8_consume(meta, g.col, mu_prev, rl0, [&](int rl, const float4& acc) { gv[rl - rl0] = acc; });
            float gm = 0.f;                      // scale per WARP = per 16 rows: the rows of a row transform are independent, the warp
#pragma unroll                                   // that splits rows 16 w .. 16 w + 15 is the warp that reads them back from TMEM -- no exchange
            for (int i = 0; i < 8; ++i)
                gm = fmaxf(gm, fmaxf(fmaxf(fabsf(gv[i].x), fabsf(gv[i].y)), fmaxf(fabsf(gv[i].z), fabsf(gv[i].w))));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(FULL, gm, o));
            const float sa = pow2_scale_for(gm);
            c1 = (1.0f / sa) * inv_sw;
#pragma unroll
            for (int i = 0; i < 8; ++i) store_split2h(a1, rl0 + i, c4, gv[i], sa);
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            issue_transform_h2(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        float4 bv[8];                            // base rows of the epilogue while the tensor core works (requested before the
#pragma unroll                                   // metadata prefetch, which stalls on its own dependent loads)
        for (int i = 0; i < 8; ++i)
            bv[i] = rl0 + i < rows ? ldg4(base + (size_t)(row0 + rl0 + i) * 64 + 4 * c4) : f4zero();
        {
            const int nt = tile + gridDim.x;
            if (nt < num_tiles) gather8_prefetch(meta, g, g.rowptr, g.col, nt * TCR, min(TCR, g.num_nodes - nt * TCR), rl0);
        }
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        {
            float v[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(d1 + lane_sel + 32 * half, v);
                if (lane < 16) {
                    const int r = 16 * warp + lane;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        st4(stg + stg_off(r, 8 * half + i), make_float4(v[4 * i] * c1, v[4 * i + 1] * c1, v[4 * i + 2] * c1, v[4 * i + 3] * c1));
                }
            }
        }
        tc::tc_fence_before_sync();
        __syncwarp();                            // the warp reads back the staging rows it wrote itself (rows 16 w .. 16 w + 15)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = rl0 + i;
            if (rl < rows) {
                float4 v = ld4(stg + stg_off(rl, c4));
                st4(mu_next + (size_t)(row0 + rl) * 64 + 4 * c4,
                    make_float4(relu(v.x + bv[i].x), relu(v.y + bv[i].y), relu(v.z + bv[i].z), relu(v.w + bv[i].w)));
            }
        }
        __syncthreads();                        // the next tile's pieces overwrite the staging rows of OTHER warps
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(d1);
}

// ------------------------------------------------------------------------------------
// Last backward stage on two fp16 pieces per operand, in the shape of the forward round (r02d): 4-warp CTAs, three per SM.
//   D = sum_t du^t (fixed order t = 1..T) ;  dtheta1b += D^T h ;  dh = (D theta1b) * [h > 0] ;  dtheta1a += dh^T x ;
//   column sums of D, c D, (Npad - c) D, dh.        h = relu(theta1a x + b1a) is recomputed (K = F <= 8, CUDA cores).
// A lane owns rows rl0 .. rl0 + 7, columns 4 c4 .. 4 c4 + 3 of the tile in every phase (plane sums, h, masks, epilogue,
// column sums, dh x^T), so nothing but the two MMA operands and the staged D theta1b crosses threads; all long sums are
// per-lane fp32 registers over the CTA's tiles, reduced across the eight lanes that share a column group in fixed order
// at the end.  The D and h tiles are scaled per tile (k_embed_round_bwd_h2), D2 is read out per tile.
// Partial layout: EmbedPartial<64> rows TH1B .. TH1A (s2v_embed.cu).
// ------------------------------------------------------------------------------------
constexpr int FH2_FS = 8;                        // features padded to 8 (F <= 8)
__global__ void __launch_bounds__(TC_THREADS, 3)
k_embed_final_bwd_h2(s2v_graph_t g, s2v_embed_params_t prm, const float* __restrict__ x, const float* __restrict__ dus,
                     const float* __restrict__ du_last, int T, float* __restrict__ partial, int partial_stride, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta1b^T 2^ew: tile(row k, col n) = theta1b[n][k]
    uint8_t* a1 = w1 + TILE16X2;                 // D tile 2^ed; reused as the staging tile
    uint8_t* h1 = a1 + TILE16X2;                 // h tile 2^eh
    float* W1a = reinterpret_cast<float*>(h1 + TILE16X2);       // [FS][64]
    float* xs = W1a + FH2_FS * 64;                              // [64][FS]
    float* b1a = xs + TCR * FH2_FS;                             // [64]
    float* cw = b1a + 64;                                       // [64] in-degree c_r, [64] Npad - c_r
    H2Shared* sh = reinterpret_cast<H2Shared*>(cw + 128);
    float* stg = reinterpret_cast<float*>(a1);
    const int F = prm.node_dim;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const unsigned FULL = 0xffffffffu;
    const size_t plane = (size_t)g.num_nodes * 64;

    if (warp == 0) tc::tmem_alloc<128>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 2); tc::fence_barrier_init(); }
    {
        float wm = 0.f;
        for (int e = t; e < 64 * 64; e += TC_THREADS) wm = fmaxf(wm, fabsf(__ldg(prm.theta1b_w + e)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(FULL, wm, o));
        if (lane == 0) sh->wmax[warp] = wm;
        for (int e = t; e < 64 * FH2_FS; e += TC_THREADS) { int f = e >> 6, n = e & 63; W1a[e] = f < F ? __ldg(prm.theta1a_w + n * F + f) : 0.f; }
        if (t < 64) b1a[t] = __ldg(prm.theta1a_b + t);
        __syncthreads();
        if (t == 0) sh->wscale = pow2_scale_for(fmaxf(fmaxf(sh->wmax[0], sh->wmax[1]), fmaxf(sh->wmax[2], sh->wmax[3])));
        __syncthreads();
        const float sw = sh->wscale;
        for (int e = t; e < 64 * 16; e += TC_THREADS) {
            const int row = e >> 4, cc = e & 15;
            const float* W = prm.theta1b_w;
            store_split2h(w1, row, cc, make_float4(__ldg(W + (4 * cc) * 64 + row), __ldg(W + (4 * cc + 1) * 64 + row),
                                                   __ldg(W + (4 * cc + 2) * 64 + row), __ldg(W + (4 * cc + 3) * 64 + row)), sw);
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const float inv_sw = 1.0f / sh->wscale;
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1), s_h = tc::smem_u32(h1);
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    uint32_t phase = 0;
    const int c4 = lane & 15, rl0 = warp * 16 + 8 * (lane >> 4);
    float racc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) racc[i] = 0.f;
    float4 pD = f4zero(), pA1 = f4zero(), pA0 = f4zero(), pB1a = f4zero();      // columns 4 c4 .. + 3, this lane's rows
    float4 pth1a[FH2_FS];                                                      // [feature] x columns 4 c4 .. + 3
#pragma unroll
    for (int f = 0; f < FH2_FS; ++f) pth1a[f] = f4zero();

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        // ---- requests: x elements, per-row degree data, then the T planes of the lane's rows ----
        float xr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = t + TC_THREADS * j;
            xr[j] = e < rows * F ? __ldg(x + (size_t)row0 * F + e) : 0.f;
        }
        float cn = 0.f, dn = 0.f;
        if (t < rows) {
            int li, off, gi;
            row_info(g, row0 + t, li, off, gi);
            cn = (float)__ldg(g.indeg + li);
            dn = (float)graph_npad(g, gi) - cn;
        }
        float4 dv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dv[i] = f4zero();
        {
            const size_t o = (size_t)(row0 + rl0) * 64 + 4 * c4;
            int tt = 0;
            for (; tt + 1 < T; tt += 2) {        // two planes (16 independent 128-bit loads) at a time, summed in the order t = 1..T
                const float* pa = dus + (size_t)tt * plane + o;          // (four rows x five planes per batch: more spills, no faster)
                const float* pb = (tt + 1 < T - 1 ? dus + (size_t)(tt + 1) * plane : du_last) + o;
                float4 a[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = rl0 + i < rows ? ldg4(pa + (size_t)i * 64) : f4zero();
#pragma unroll
                for (int i = 0; i < 8; ++i) b[i] = rl0 + i < rows ? ldg4(pb + (size_t)i * 64) : f4zero();
#pragma unroll
                for (int i = 0; i < 8; ++i) dv[i] = f4add(f4add(dv[i], a[i]), b[i]);
            }
            if (tt < T) {
                const float* pa = (tt < T - 1 ? dus + (size_t)tt * plane : du_last) + o;
#pragma unroll
                for (int i = 0; i < 8; ++i) if (rl0 + i < rows) dv[i] = f4add(dv[i], ldg4(pa + (size_t)i * 64));
            }
        }
        // ---- x tile and degree weights to shared memory (rows past the end: x = 0, c = 0) ----
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = t + TC_THREADS * j;
            if (e < TCR * F) xs[(e / F) * FH2_FS + (e % F)] = xr[j];
        }
        if (t < TCR) { cw[t] = cn; cw[64 + t] = dn; }
        __syncthreads();
        // ---- h, masks, column sums of D, tile maxima ----
        float4 hv[8];
        uint32_t hbits = 0;
        {
            const float4 bb = ld4(b1a + 4 * c4);
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[i] = bb;
#pragma unroll
            for (int f0 = 0; f0 < FH2_FS; f0 += 4) {
                const float4 w0 = ld4(W1a + f0 * 64 + 4 * c4), w1v = ld4(W1a + (f0 + 1) * 64 + 4 * c4),
                             w2v = ld4(W1a + (f0 + 2) * 64 + 4 * c4), w3v = ld4(W1a + (f0 + 3) * 64 + 4 * c4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 xv = ld4(xs + (rl0 + i) * FH2_FS + f0);
                    hv[i].x = fmaf(xv.x, w0.x, hv[i].x); hv[i].y = fmaf(xv.x, w0.y, hv[i].y);
                    hv[i].z = fmaf(xv.x, w0.z, hv[i].z); hv[i].w = fmaf(xv.x, w0.w, hv[i].w);
                    hv[i].x = fmaf(xv.y, w1v.x, hv[i].x); hv[i].y = fmaf(xv.y, w1v.y, hv[i].y);
                    hv[i].z = fmaf(xv.y, w1v.z, hv[i].z); hv[i].w = fmaf(xv.y, w1v.w, hv[i].w);
                    hv[i].x = fmaf(xv.z, w2v.x, hv[i].x); hv[i].y = fmaf(xv.z, w2v.y, hv[i].y);
                    hv[i].z = fmaf(xv.z, w2v.z, hv[i].z); hv[i].w = fmaf(xv.z, w2v.w, hv[i].w);
                    hv[i].x = fmaf(xv.w, w3v.x, hv[i].x); hv[i].y = fmaf(xv.w, w3v.y, hv[i].y);
                    hv[i].z = fmaf(xv.w, w3v.z, hv[i].z); hv[i].w = fmaf(xv.w, w3v.w, hv[i].w);
                }
            }
        }
        float gm = 0.f, hm = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            hv[i] = make_float4(relu(hv[i].x), relu(hv[i].y), relu(hv[i].z), relu(hv[i].w));
            hbits |= (hv[i].x > 0.f ? 1u : 0u) << (4 * i) | (hv[i].y > 0.f ? 2u : 0u) << (4 * i) |
                     (hv[i].z > 0.f ? 4u : 0u) << (4 * i) | (hv[i].w > 0.f ? 8u : 0u) << (4 * i);
            hm = fmaxf(hm, fmaxf(fmaxf(hv[i].x, hv[i].y), fmaxf(hv[i].z, hv[i].w)));
            gm = fmaxf(gm, fmaxf(fmaxf(fabsf(dv[i].x), fabsf(dv[i].y)), fmaxf(fabsf(dv[i].z), fabsf(dv[i].w))));
            const float c = cw[rl0 + i], d = cw[64 + rl0 + i];
            pD = f4add(pD, dv[i]);
            pA1.x = fmaf(c, dv[i].x, pA1.x); pA1.y = fmaf(c, dv[i].y, pA1.y); pA1.z = fmaf(c, dv[i].z, pA1.z); pA1.w = fmaf(c, dv[i].w, pA1.w);
            pA0.x = fmaf(d, dv[i].x, pA0.x); pA0.y = fmaf(d, dv[i].y, pA0.y); pA0.z = fmaf(d, dv[i].z, pA0.z); pA0.w = fmaf(d, dv[i].w, pA0.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { gm = fmaxf(gm, __shfl_xor_sync(FULL, gm, o)); hm = fmaxf(hm, __shfl_xor_sync(FULL, hm, o)); }
        if (lane == 0) { sh->wmax[warp] = gm; sh->wmax[4 + warp] = hm; }
        __syncthreads();
        const float sd = pow2_scale_for(fmaxf(fmaxf(sh->wmax[0], sh->wmax[1]), fmaxf(sh->wmax[2], sh->wmax[3])));
        const float shh = pow2_scale_for(fmaxf(fmaxf(sh->wmax[4], sh->wmax[5]), fmaxf(sh->wmax[6], sh->wmax[7])));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            store_split2h(a1, rl0 + i, c4, dv[i], sd);
            store_split2h(h1, rl0 + i, c4, hv[i], shh);
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            issue_transform_h2(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        } else if (warp == 2) {
            tc::tc_fence_after_sync();
            issue_reduce_h2(d2, s_a, s_h);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        {
            const float c1 = (1.0f / sd) * inv_sw;
            float v[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(d1 + lane_sel + 32 * half, v);
                if (lane < 16) {
                    const int r = 16 * warp + lane;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        st4(stg + stg_off(r, 8 * half + i), make_float4(v[4 * i] * c1, v[4 * i + 1] * c1, v[4 * i + 2] * c1, v[4 * i + 3] * c1));
                }
            }
            const float c2 = (1.0f / sd) * (1.0f / shh);
            tc::tmem_ld32(d2 + lane_sel, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (lane < 16) racc[i] = fmaf(v[i], c2, racc[i]);
            tc::tmem_ld32(d2 + lane_sel + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float o = __shfl_sync(FULL, v[i], lane & 15);
                if (lane >= 16) racc[i] = fmaf(o, c2, racc[i]);
            }
        }
        tc::tc_fence_before_sync();
        __syncwarp();                            // the warp reads back the staging rows it wrote itself
        // ---- dh = (D theta1b) * [h > 0];  db1a += dh;  dtheta1a += dh x^T  (rows past the end: D = 0 -> dh = 0) ----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = rl0 + i;
            float4 v = ld4(stg + stg_off(rl, c4));
            const uint32_t b = hbits >> (4 * i);
            const float4 dh = make_float4(b & 1u ? v.x : 0.f, b & 2u ? v.y : 0.f, b & 4u ? v.z : 0.f, b & 8u ? v.w : 0.f);
            pB1a = f4add(pB1a, dh);
            const float4 x0 = ld4(xs + rl * FH2_FS), x1 = ld4(xs + rl * FH2_FS + 4);
            const float xf[FH2_FS] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int f = 0; f < FH2_FS; ++f) {
                pth1a[f].x = fmaf(dh.x, xf[f], pth1a[f].x); pth1a[f].y = fmaf(dh.y, xf[f], pth1a[f].y);
                pth1a[f].z = fmaf(dh.z, xf[f], pth1a[f].z); pth1a[f].w = fmaf(dh.w, xf[f], pth1a[f].w);
            }
        }
        __syncthreads();                        // staging, xs, cw are overwritten by the next tile
    }
    float* my = partial + (size_t)blockIdx.x * partial_stride;
    {   // dtheta1b[n = 16 warp + lane % 16][32 (lane / 16) .. + 32)
        float* dst = my + EP_TH1B + (16 * warp + (lane & 15)) * 64 + 32 * (lane >> 4);
#pragma unroll
        for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]));
    }
    // per-lane column sums -> per-CTA sums: the eight lanes (warp, half) that share a column group, in fixed order
    float* red = reinterpret_cast<float*>(w1);   // [8][(4 + FS) * 64] floats = 24 KB: weight, D and h tiles are dead
    {
        const int slot = 2 * warp + (lane >> 4);
        float* r = red + slot * ((4 + FH2_FS) * 64) + 4 * c4;
        st4(r, pD); st4(r + 64, pA1); st4(r + 128, pA0); st4(r + 192, pB1a);
#pragma unroll
        for (int f = 0; f < FH2_FS; ++f) st4(r + (4 + f) * 64, pth1a[f]);
    }
    __syncthreads();
    for (int e = t; e < (4 + FH2_FS) * 64; e += TC_THREADS) {
        float sum = 0.f;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) sum += red[sl * ((4 + FH2_FS) * 64) + e];
        const int q = e >> 6, col = e & 63;
        if (q == 0) my[EP_COLD + col] = sum;
        else if (q == 1) my[EP_A1 + col] = sum;
        else if (q == 2) my[EP_A0 + col] = sum;
        else if (q == 3) my[EP_B1A + col] = sum;
        else if (q - 4 < F) my[EP_TH1A + col * F + (q - 4)] = sum;
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(sh->tmem_base);
}

// ------------------------------------------------------------------------------------
// Q / logit head on two fp16 pieces per operand (k_head_tc's pipeline, k_embed_round_h2's arithmetic): 32 KB of tiles, 12
// MMAs per tile; the mu tile is scaled per tile, theta7 once per CTA, both leave in the row owner's epilogue.
// ------------------------------------------------------------------------------------
template <int CTAS>
__global__ void __launch_bounds__(TC_THREADS, CTAS)
k_head_h2(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ mu, const float* __restrict__ gstate,
          float* __restrict__ q, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta7 [n][k] 2^ew, two fp16 pieces
    uint8_t* a1 = w1 + TILE16X2;                 // mu tile 2^ea, two fp16 pieces
    float* vec = reinterpret_cast<float*>(a1 + TILE16X2);      // b7, w5a, w5b
    H2Shared* sh = reinterpret_cast<H2Shared*>(vec + 3 * 64);
    float* b7 = vec, *w5a = vec + 64, *w5b = vec + 128;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const unsigned FULL = 0xffffffffu;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    {
        float wm = 0.f;
        for (int e = t; e < 64 * 64; e += TC_THREADS) wm = fmaxf(wm, fabsf(__ldg(prm.theta7_w + e)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(FULL, wm, o));
        if (lane == 0) sh->wmax[warp] = wm;
        if (t < 64) {
            b7[t] = __ldg(prm.theta7_b + t);
            w5a[t] = __ldg(prm.theta5_w + t);
            w5b[t] = __ldg(prm.theta5_w + 64 + t);
        }
        __syncthreads();
        if (t == 0) sh->wscale = pow2_scale_for(fmaxf(fmaxf(sh->wmax[0], sh->wmax[1]), fmaxf(sh->wmax[2], sh->wmax[3])));
        __syncthreads();
        const float sw = sh->wscale;
        for (int e = t; e < 64 * 16; e += TC_THREADS) store_split2h(w1, e >> 4, e & 15, ldg4(prm.theta7_w + (e >> 4) * 64 + 4 * (e & 15)), sw);
    }
    const float b5 = __ldg(prm.theta5_b);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const float inv_sw = 1.0f / sh->wscale;
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1);
    uint32_t phase = 0;
    float4 m[8];
    auto load_rows = [&](int tile) {
        const int row0 = tile * TCR, rows = min(TCR, g.num_nodes - row0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {                  // warp w holds rows 16 w .. 16 w + 15 (8 per half-warp): the rows its lanes own in TMEM
            const int rl = 16 * warp + 8 * (lane >> 4) + i, c4 = lane & 15;
            m[i] = rl < rows ? ldg4(mu + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero();
        }
    };
    if (blockIdx.x < num_tiles) load_rows(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        float am = 0.f;                          // scale per warp = per 16 rows (see k_embed_round_h2): no exchange
#pragma unroll
        for (int i = 0; i < 8; ++i) am = fmaxf(am, fmaxf(fmaxf(fabsf(m[i].x), fabsf(m[i].y)), fmaxf(fabsf(m[i].z), fabsf(m[i].w))));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(FULL, am, o));
        const float sa = pow2_scale_for(am);
        const float c1 = (1.0f / sa) * inv_sw;
#pragma unroll
        for (int i = 0; i < 8; ++i) store_split2h(a1, 16 * warp + 8 * (lane >> 4) + i, lane & 15, m[i], sa);
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            issue_transform_h2(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        if (tile + (int)gridDim.x < num_tiles) load_rows(tile + gridDim.x);     // in flight during the MMAs and the epilogue
        const int rl = 16 * warp + (lane & 15);  // graph part of the row owner's q while the tensor core works
        float sg = 0.f;
        if (lane < 16 && rl < rows) {
            int li, off, gi;
            row_info(g, row0 + rl, li, off, gi);
            const float* G = gstate + (size_t)gi * 64;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 gv = ldg4(G + 4 * c), wa = ld4(w5a + 4 * c);
                sg = fmaf(wa.x, relu(gv.x), sg); sg = fmaf(wa.y, relu(gv.y), sg);
                sg = fmaf(wa.z, relu(gv.z), sg); sg = fmaf(wa.w, relu(gv.w), sg);
            }
        }
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        float sl = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tc::tmem_ld32(d1 + ((uint32_t)(32 * warp) << 16) + 32 * half, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 bb = ld4(b7 + 32 * half + 4 * c), wb = ld4(w5b + 32 * half + 4 * c);
                sl = fmaf(wb.x, relu(fmaf(v[4 * c], c1, bb.x)), sl); sl = fmaf(wb.y, relu(fmaf(v[4 * c + 1], c1, bb.y)), sl);
                sl = fmaf(wb.z, relu(fmaf(v[4 * c + 2], c1, bb.z)), sl); sl = fmaf(wb.w, relu(fmaf(v[4 * c + 3], c1, bb.w)), sl);
            }
        }
        if (lane < 16 && rl < rows) q[row0 + rl] = (sg + sl) + b5;
        tc::tc_fence_before_sync();            // no barrier here: the next tile's pieces are stored behind ITS scale-exchange barrier
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(d1);
}

