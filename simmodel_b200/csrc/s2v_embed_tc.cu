// Tensor-core (tcgen05 / TMEM) variants of the propagation round, p = 64 only.
//
// Same math as k_embed_round / k_embed_round_bwd (s2v_embed.cu) with the p x p contractions
// moved to the 5th-generation tensor cores as 3xTF32 split-precision UMMAs (s2v_tc.cuh):
// CUDA cores only gather neighbour rows, split them into hi/lo TF32 tiles in shared memory
// (canonical SWIZZLE_128B layout, written directly by the gather lanes) and run the epilogue
// from TMEM.  A CTA of 128 threads owns 64 consecutive node rows per tile (UMMA M = 64),
// several CTAs per SM interleave their gather / MMA / epilogue phases.
//
//   forward :  D1[64x64]  = agg * theta2^T                        (A K-major, B K-major)
//   backward:  D1[64x64]  = g * theta2                            (A K-major, B = theta2^T K-major)
//              D2[128x64] += [g_hi ; g_lo]^T * mu_hi + [g_hi ; g_lo]^T * mu_lo
//                            (both MN-major views of the same tiles; K = the 64 rows of the
//                             tile; rows 0-63 of D2 + rows 64-127 of D2 = dtheta2 partial,
//                             kept in TMEM across all tiles of the CTA)
#include "s2v_common.cuh"
#include "s2v_tc.cuh"

#ifdef S2V_TC_TIMELINE
__device__ long long g_timeline[16 * 8];
#define TL(slot) do { if (blockIdx.x == 0 && threadIdx.x == 32 && tl_n < 8) g_timeline[tl_n * 16 + (slot)] = clock64(); } while (0)
#define TL3(slot) do { if (blockIdx.x == 0 && threadIdx.x == 32 && tl_n >= 0 && tl_n < 8) g_timeline[tl_n * 16 + (slot)] = clock64(); } while (0)
#define TL2(slot) do { if (blockIdx.x == 0 && threadIdx.x == 32 && tl_n >= 0 && tl_n < 8) g_timeline[tl_n * 16 + (slot)] = clock64(); } while (0)
#else
#define TL(slot) do {} while (0)
#define TL2(slot) do {} while (0)
#define TL3(slot) do {} while (0)
#endif

namespace {

constexpr int TCR = 64;                         // rows per tile
constexpr int TC_THREADS = 128;
constexpr uint32_t TILE_BYTES = TCR * 256;      // one hi or lo tile: 2 column blocks x 64 rows x 128 B
constexpr int STG_LD = 68;                      // staging leading dimension (floats)

// split a float4 chunk of row r into the hi / lo tiles
__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, int r, int c4, const float4& v) {
    float4 hi, lo;
    tc::split4(v, hi, lo);
    uint32_t off = tc::tile_off<TCR>(r, c4);
    *reinterpret_cast<float4*>(hi_tile + off) = hi;
    *reinterpret_cast<float4*>(lo_tile + off) = lo;
}

// same for an MN-major (SWIZZLE_128B_BASE32B) tile
__device__ __forceinline__ void store_split_mn(uint8_t* hi_tile, uint8_t* lo_tile, int r, int c4, const float4& v) {
    float4 hi, lo;
    tc::split4(v, hi, lo);
    uint32_t off = tc::tile_off_mn<TCR>(r, c4);
    *reinterpret_cast<float4*>(hi_tile + off) = hi;
    *reinterpret_cast<float4*>(lo_tile + off) = lo;
}

// Row owners (lanes 0-15 of each warp own row 16*warp + lane, the UMMA M = 64 lane mapping
// lane = 32*(row/16) + row%16) copy their row of an MN-major tile into TMEM as the A operand
// of the row transform: 64 columns of their lane.  (The 4-row swizzle of the tile spreads
// consecutive rows over four bank groups: 2-way conflicts at most.)
__device__ __forceinline__ void row_to_tmem(const uint8_t* tile, uint32_t a_tmem) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = 16 * warp + (lane & 15);
    float v[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 x = *reinterpret_cast<const float4*>(tile + tc::tile_off_mn<TCR>(r, 8 * half + j));
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
        }
        tc::tmem_st32(a_tmem + ((uint32_t)(32 * warp) << 16) + 32 * half, v);
    }
}

// D[64x64] = A * B^T with A in TMEM (hi at a_tmem, lo at a_tmem + 64 columns), B K-major in smem
__device__ __forceinline__ void issue_transform_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::idesc_tf32(64, 64, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32_ts(d_tmem, a_tmem + 64 + 8 * ks, tc::desc_kmajor<TCR>(b_hi, ks), idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32_ts(d_tmem, a_tmem + 8 * ks, tc::desc_kmajor<TCR>(b_lo, ks), idesc, true);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32_ts(d_tmem, a_tmem + 8 * ks, tc::desc_kmajor<TCR>(b_hi, ks), idesc, true);
}

// W [64][64] row-major -> hi/lo tiles with tile(row, col) = transpose ? W[col][row] : W[row][col]
__device__ __forceinline__ void load_weight_tiles(uint8_t* hi_tile, uint8_t* lo_tile, const float* __restrict__ W, bool transpose) {
    for (int e = threadIdx.x; e < 64 * 64; e += TC_THREADS) {
        int a = e >> 6, b = e & 63;
        float v = __ldg(W + e);
        int row = transpose ? b : a, col = transpose ? a : b;
        float hi = tc::tf32_rn(v), lo = tc::tf32_rn(v - hi);
        uint32_t off = tc::tile_off<TCR>(row, col >> 2) + (col & 3) * 4;
        *reinterpret_cast<float*>(hi_tile + off) = hi;
        *reinterpret_cast<float*>(lo_tile + off) = lo;
    }
}

// D[64x64] = A * B^T with A (rows r, K-major) and B (rows n, K-major), 3xTF32; small terms first
__device__ __forceinline__ void issue_transform(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::idesc_tf32(64, 64, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32(d_tmem, tc::desc_kmajor<TCR>(a_lo, ks), tc::desc_kmajor<TCR>(b_hi, ks), idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32(d_tmem, tc::desc_kmajor<TCR>(a_hi, ks), tc::desc_kmajor<TCR>(b_lo, ks), idesc, true);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32(d_tmem, tc::desc_kmajor<TCR>(a_hi, ks), tc::desc_kmajor<TCR>(b_hi, ks), idesc, true);
}

// D2[128x64] = [A_hi ; A_lo]^T * B_lo + [A_hi ; A_lo]^T * B_hi, K = the 64 rows of the tiles;
// all tiles MN-major (SWIZZLE_128B_BASE32B).  a_hi must be followed directly by a_lo (M = 128
// spans the four 32-column blocks at LBO = 8 KB).  One tile per call: the tensor core adds
// into its fp32 accumulator with truncation, so long chains are summed on the CUDA cores
// instead (flush_d2).
__device__ __forceinline__ void issue_reduce(uint32_t d2_tmem, uint32_t a_hi, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 1, 1);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32(d2_tmem, tc::desc_mnmajor<TCR>(a_hi, ks), tc::desc_mnmajor<TCR>(b_lo, ks), idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_tf32(d2_tmem, tc::desc_mnmajor<TCR>(a_hi, ks), tc::desc_mnmajor<TCR>(b_hi, ks), idesc, true);
}

// racc[0..32) += D2[n][32h..32h+32) + D2[64+n][32h..32h+32) with n = 32*(warp&1) + lane, h = warp>>1.
// Warp w can only read TMEM lanes 32w..32w+31 (rows 32w+lane of D2): each thread keeps the
// column half it accumulates and hands the other half to warp w^2 through xbuf ([4][32][36] floats).
__device__ __forceinline__ void flush_d2(uint32_t d2_tmem, float* xbuf, float (&racc)[32]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = warp >> 1;
    float keep[32], give[32];
    tc::tmem_ld32(d2_tmem + ((uint32_t)(32 * warp) << 16) + 32 * h, keep);
    tc::tmem_ld32(d2_tmem + ((uint32_t)(32 * warp) << 16) + 32 * (1 - h), give);
    float* mine = xbuf + (warp * 32 + lane) * 36;
#pragma unroll
    for (int i = 0; i < 32; i += 4) st4(mine + i, make_float4(give[i], give[i + 1], give[i + 2], give[i + 3]));
    tc::tc_fence_before_sync();
    __syncthreads();
    const float* theirs = xbuf + ((warp ^ 2) * 32 + lane) * 36;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float4 o = ld4(theirs + i);
        racc[i] += keep[i] + o.x; racc[i + 1] += keep[i + 1] + o.y;
        racc[i + 2] += keep[i + 2] + o.z; racc[i + 3] += keep[i + 3] + o.w;
    }
}

// D1 (UMMA M = 64: row i lives in TMEM lane 32*(i/16) + i%16) -> staging[row][STG_LD] in shared memory
__device__ __forceinline__ void d1_to_staging(uint32_t d1_tmem, float* stg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tc::tmem_ld32(d1_tmem + ((uint32_t)(32 * warp) << 16) + 32 * half, v);
        if (lane < 16) {
            float* dst = stg + (16 * warp + lane) * STG_LD + 32 * half;
#pragma unroll
            for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        }
    }
}

struct TcShared {
    uint64_t bar;
    uint32_t tmem_base;
    uint32_t pad;
};

// ------------------------------------------------------------------------------------
// forward round
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
k_embed_round_tc(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ base,
                 const float* __restrict__ mu_prev, float* __restrict__ mu_next, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;                        // theta2 [n][k]
    uint8_t* b_lo = b_hi + TILE_BYTES;
    uint8_t* a_hi = b_lo + TILE_BYTES;           // gathered tile; a_lo follows; both reused as staging
    uint8_t* a_lo = a_hi + TILE_BYTES;
    TcShared* sh = reinterpret_cast<TcShared*>(a_lo + TILE_BYTES);
    float* stg = reinterpret_cast<float*>(a_hi);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles(b_hi, b_lo, theta2_w, false);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a_hi = tc::smem_u32(a_hi), s_a_lo = tc::smem_u32(a_lo), s_b_hi = tc::smem_u32(b_hi), s_b_lo = tc::smem_u32(b_lo);
    uint32_t phase = 0;
    const int rl0 = warp * 16 + 8 * (lane >> 4);
    Gather8Meta meta;
    if (blockIdx.x < num_tiles)
        gather8_prefetch(meta, g, g.rowptr, g.col, blockIdx.x * TCR, min(TCR, g.num_nodes - blockIdx.x * TCR), rl0);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
#ifdef S2V_TC_TIMELINE
        const int tl_n = (tile - blockIdx.x) / gridDim.x - 2;
#endif
        TL(0);
        // gather (8 rows per half-warp, fixed neighbour order; CSR metadata was prefetched) and split into hi/lo tiles
        gather8_consume(meta, g.col, mu_prev, rl0,
                        [&](int rl, const float4& acc) { store_split(a_hi, a_lo, rl, lane & 15, acc); });
        TL(1);
        tc::fence_proxy_async_smem();
        __syncthreads();
        TL(2);
        if (t == 0) {
            tc::tc_fence_after_sync();
            issue_transform(d1, s_a_hi, s_a_lo, s_b_hi, s_b_lo);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();                           // lanes 1-31 of warp 0 must not sleep in try_wait before lane 0 has issued
        TL(3);
        {
            const int nt = tile + gridDim.x;     // CSR metadata of the next tile: overlaps MMA + epilogue
            if (nt < num_tiles) gather8_prefetch(meta, g, g.rowptr, g.col, nt * TCR, min(TCR, g.num_nodes - nt * TCR), rl0);
        }
        TL(4);
        // prefetch the base rows of the epilogue while the tensor core works
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            bv[i] = rl < rows ? ldg4(base + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero();
        }
        TL(5);
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        TL(6);
        d1_to_staging(d1, stg);                 // the A tiles are dead once the MMAs have completed
        tc::tc_fence_before_sync();
        __syncthreads();
        TL(7);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            if (rl < rows) {
                float4 v = ld4(stg + rl * STG_LD + 4 * c4);
                st4(mu_next + (size_t)(row0 + rl) * 64 + 4 * c4,
                    make_float4(relu(v.x + bv[i].x), relu(v.y + bv[i].y), relu(v.z + bv[i].z), relu(v.w + bv[i].w)));
            }
        }
        TL(8);
        __syncthreads();                        // staging is overwritten by the next gather
        TL(9);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(d1);
}

// ------------------------------------------------------------------------------------
// backward round, 8 warps per CTA.
// The tile working set (96 KB of operand tiles, 256 TMEM columns) allows only two CTAs per SM, so
// the CTA is made wider instead: every half-warp gathers 4 rows (16 neighbour loads + 4 own-row
// loads in flight per lane, 16 warps per SM), and the TMEM copies, the epilogue and the dtheta2
// flush are split over twice as many threads.  Warp w works on TMEM lane quarter q = w & 3 and on
// the column half hh = w >> 2.
// ------------------------------------------------------------------------------------
constexpr int TC_THREADS2 = 256;

__device__ __forceinline__ void load_weight_tiles256(uint8_t* hi_tile, uint8_t* lo_tile, const float* __restrict__ W, bool transpose) {
    for (int e = threadIdx.x; e < 64 * 64; e += TC_THREADS2) {
        int a = e >> 6, b = e & 63;
        float v = __ldg(W + e);
        int row = transpose ? b : a, col = transpose ? a : b;
        float hi = tc::tf32_rn(v), lo = tc::tf32_rn(v - hi);
        uint32_t off = tc::tile_off<TCR>(row, col >> 2) + (col & 3) * 4;
        *reinterpret_cast<float*>(hi_tile + off) = hi;
        *reinterpret_cast<float*>(lo_tile + off) = lo;
    }
}

__global__ void __launch_bounds__(TC_THREADS2, 2)
k_embed_round_bwd_tc256(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ du_t,
                        const float* __restrict__ mu_prev, float* __restrict__ du_prev, float* __restrict__ partial,
                        int partial_stride, int accumulate, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;                        // theta2^T: tile(row k, col n) = theta2[n][k]
    uint8_t* b_lo = b_hi + TILE_BYTES;
    uint8_t* g_hi = b_lo + TILE_BYTES;           // gathered du tile (g_lo must follow g_hi)
    uint8_t* g_lo = g_hi + TILE_BYTES;
    uint8_t* m_hi = g_lo + TILE_BYTES;           // mu^{t-1} tile
    uint8_t* m_lo = m_hi + TILE_BYTES;
    TcShared* sh = reinterpret_cast<TcShared*>(m_lo + TILE_BYTES);
    float* stg = reinterpret_cast<float*>(g_hi);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, hh = warp >> 2;

    if (warp == 0) tc::tmem_alloc<256>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles256(b_hi, b_lo, theta2_w, true);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    // TMEM columns: D1 [0,64)  D2 [64,128)  A_hi [128,192)  A_lo [192,256)
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64, a_tm = sh->tmem_base + 128;
    const uint32_t s_g_hi = tc::smem_u32(g_hi), s_b_hi = tc::smem_u32(b_hi), s_b_lo = tc::smem_u32(b_lo),
                   s_m_hi = tc::smem_u32(m_hi), s_m_lo = tc::smem_u32(m_lo);
    const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
    uint32_t phase = 0;
    float racc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) racc[i] = 0.f;
    const int c4 = lane & 15, rl0 = 4 * (2 * warp + (lane >> 4));
    GatherMeta<4> meta;
    if (blockIdx.x < num_tiles)
        gatherN_prefetch<4>(meta, g, g.rowptr_t, g.col_t, blockIdx.x * TCR, min(TCR, g.num_nodes - blockIdx.x * TCR), rl0);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
#ifdef S2V_TC_TIMELINE
        const int tl_n = (tile - blockIdx.x) / gridDim.x - 2;
#endif
        TL3(0);
        {
            float4 m[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                m[i] = rl0 + i < rows ? ldg4(mu_prev + (size_t)(row0 + rl0 + i) * 64 + 4 * c4) : f4zero();
            gatherN_consume<4, 4>(meta, g.col_t, du_t, rl0,
                                  [&](int rl, const float4& acc) { store_split_mn(g_hi, g_lo, rl, c4, acc); });
#pragma unroll
            for (int i = 0; i < 4; ++i) store_split_mn(m_hi, m_lo, rl0 + i, c4, m[i]);
            const int nt = tile + gridDim.x;
            if (nt < num_tiles)
                gatherN_prefetch<4>(meta, g, g.rowptr_t, g.col_t, nt * TCR, min(TCR, g.num_nodes - nt * TCR), rl0);
        }
        TL3(1);
        tc::fence_proxy_async_smem();
        __syncthreads();
        TL3(2);
        if (t == 0) {                            // the reduction only needs the shared-memory tiles: start it now
            tc::tc_fence_after_sync();
            issue_reduce(d2, s_g_hi, s_m_hi, s_m_lo);
        }
        __syncwarp();
        {   // A operand of the row transform -> TMEM: row owner = lane < 16 of quarter q, column half hh
            const int r = 16 * q + (lane & 15);
            float v[32];
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const uint8_t* tile_p = part ? g_lo : g_hi;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 x4 = *reinterpret_cast<const float4*>(tile_p + tc::tile_off_mn<TCR>(r, 8 * hh + j));
                    v[4 * j] = x4.x; v[4 * j + 1] = x4.y; v[4 * j + 2] = x4.z; v[4 * j + 3] = x4.w;
                }
                tc::tmem_st32(a_tm + 64 * part + lane_sel + 32 * hh, v);
            }
            tc::tmem_st_wait();
        }
        tc::tc_fence_before_sync();
        __syncthreads();
        TL3(3);
        if (t == 0) {
            tc::tc_fence_after_sync();
            issue_transform_ts(d1, a_tm, s_b_hi, s_b_lo);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        TL3(4);
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        TL3(5);
        {   // D1 -> staging (overwrites the g tiles: every MMA reading them has completed)
            float v[32];
            tc::tmem_ld32(d1 + lane_sel + 32 * hh, v);
            if (lane < 16) {
                float* dst = stg + (16 * q + lane) * STG_LD + 32 * hh;
#pragma unroll
                for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
            }
        }
        tc::tc_fence_before_sync();
        __syncthreads();
        TL3(6);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = t + TC_THREADS2 * i, rl = e >> 4, cc = e & 15;
            if (rl < rows) {
                float4 v = ld4(stg + rl * STG_LD + 4 * cc);
                float4 mm = *reinterpret_cast<const float4*>(m_hi + tc::tile_off_mn<TCR>(rl, cc));   // sign(mu) == sign(mu_hi)
                st4(du_prev + (size_t)(row0 + rl) * 64 + 4 * cc,
                    make_float4(mm.x > 0.f ? v.x : 0.f, mm.y > 0.f ? v.y : 0.f, mm.z > 0.f ? v.z : 0.f, mm.w > 0.f ? v.w : 0.f));
            }
        }
        TL3(7);
        __syncthreads();                        // staging is re-used as the exchange buffer
        {   // dtheta2 flush: thread holds D2[32q+lane][32hh..32hh+32); rows 0-63 = hi part, 64-127 = lo part
            // of n = 32(q&1)+lane.  The q<2 thread accumulates columns [32hh, 32hh+16), its partner (q^2)
            // the columns [32hh+16, 32hh+32); each hands the other 16 values through the exchange buffer.
            float v[32];
            tc::tmem_ld32(d2 + lane_sel + 32 * hh, v);
            const bool upper = (q >> 1) != 0;      // warp-uniform: keep columns [16,32) and give [0,16), or the reverse
            float keep[16], give[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { keep[i] = upper ? v[16 + i] : v[i]; give[i] = upper ? v[i] : v[16 + i]; }
            float* xb = stg + (size_t)(warp * 32 + lane) * 20;
#pragma unroll
            for (int i = 0; i < 16; i += 4) st4(xb + i, make_float4(give[i], give[i + 1], give[i + 2], give[i + 3]));
            tc::tc_fence_before_sync();
            __syncthreads();
            const float* theirs = stg + (size_t)((warp ^ 2) * 32 + lane) * 20;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float4 o = ld4(theirs + i);
                racc[i] += keep[i] + o.x; racc[i + 1] += keep[i + 1] + o.y;
                racc[i + 2] += keep[i + 2] + o.z; racc[i + 3] += keep[i + 3] + o.w;
            }
        }
        __syncthreads();                        // ... and overwritten by the next gather
        TL3(8);
    }
    {
        float* my = partial + (size_t)blockIdx.x * partial_stride + (32 * (q & 1) + lane) * 64 + 32 * hh + 16 * (q >> 1);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float4 x4 = make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]);
            if (accumulate) x4 = f4add(ld4(my + i), x4);
            st4(my + i, x4);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(sh->tmem_base);
}

// ------------------------------------------------------------------------------------
// node MLP + degree term + round 1 (same math as k_embed_first, s2v_embed.cu):
//   h = relu(theta1a x + b1a) on the CUDA cores (K = F is tiny), s1 = h theta1b^T on the tensor cores,
//   base = s1 + b1b + b2 + b3 + c u1 + (Npad - c) u0,  mu^1 = relu(base)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
k_embed_first_tc(s2v_graph_t g, s2v_embed_params_t prm, const float* __restrict__ x, float* __restrict__ base,
                 float* __restrict__ mu1, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;                        // theta1b [n][k]
    uint8_t* b_lo = b_hi + TILE_BYTES;
    uint8_t* a_hi = b_lo + TILE_BYTES;           // h tile; a_lo follows; both reused as staging
    uint8_t* a_lo = a_hi + TILE_BYTES;
    float* W1a = reinterpret_cast<float*>(a_lo + TILE_BYTES);   // [F][64]
    float* xs = W1a + S2V_MAX_F * 64;                           // [64][F]
    float* vec = xs + TCR * S2V_MAX_F;                          // b1a, cst, u1, u0, r1, r0 (6 x 64)
    float* cw = vec + 6 * 64;                                   // [64] c_r, [64] Npad - c_r
    TcShared* sh = reinterpret_cast<TcShared*>(cw + 128);
    float* stg = reinterpret_cast<float*>(a_hi);
    float* b1a = vec, *cst = vec + 64, *u1 = vec + 128, *u0 = vec + 192, *r1 = vec + 256, *r0 = vec + 320;
    const int F = prm.node_dim;
    const int t = threadIdx.x, warp = t >> 5;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles(b_hi, b_lo, prm.theta1b_w, false);
    for (int e = t; e < 64 * F; e += TC_THREADS) { int n = e / F, f = e - n * F; W1a[f * 64 + n] = __ldg(prm.theta1a_w + e); }
    if (t < 64) {
        b1a[t] = __ldg(prm.theta1a_b + t);
        float w4 = __ldg(prm.theta4_w + t), b4 = __ldg(prm.theta4_b + t);
        r1[t] = relu(w4 + b4);
        r0[t] = relu(b4);
    }
    __syncthreads();
    if (t < 64) {
        float a1 = 0.f, a0 = 0.f;
        for (int k = 0; k < 64; ++k) {
            float w = __ldg(prm.theta3_w + t * 64 + k);
            a1 = fmaf(w, r1[k], a1);
            a0 = fmaf(w, r0[k], a0);
        }
        u1[t] = a1;
        u0[t] = a0;
        cst[t] = __ldg(prm.theta1b_b + t) + __ldg(prm.theta2_b + t) + __ldg(prm.theta3_b + t);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a_hi = tc::smem_u32(a_hi), s_a_lo = tc::smem_u32(a_lo), s_b_hi = tc::smem_u32(b_hi), s_b_lo = tc::smem_u32(b_lo);
    uint32_t phase = 0;
    const int c4 = t & 15, rbase = t >> 4;       // this thread's chunks: rows rbase + 8*i, columns 4*c4..4*c4+3

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        for (int e = t; e < TCR * F; e += TC_THREADS) xs[e] = e < rows * F ? __ldg(x + (size_t)row0 * F + e) : 0.f;
        if (t < TCR) {
            float c = 0.f, d = 0.f;
            if (t < rows) {
                int li, off, gi;
                row_info(g, row0 + t, li, off, gi);
                c = (float)__ldg(g.indeg + li);
                d = (float)graph_npad(g, gi) - c;
            }
            cw[t] = c;
            cw[64 + t] = d;
        }
        __syncthreads();
        {
            float4 h[8];
            const float4 bb = ld4(b1a + 4 * c4);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = bb;
            for (int f = 0; f < F; ++f) {
                const float4 w = ld4(W1a + f * 64 + 4 * c4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float xv = xs[(rbase + 8 * i) * F + f];
                    h[i].x = fmaf(xv, w.x, h[i].x); h[i].y = fmaf(xv, w.y, h[i].y);
                    h[i].z = fmaf(xv, w.z, h[i].z); h[i].w = fmaf(xv, w.w, h[i].w);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                store_split(a_hi, a_lo, rbase + 8 * i, c4, make_float4(relu(h[i].x), relu(h[i].y), relu(h[i].z), relu(h[i].w)));
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (t == 0) {
            tc::tc_fence_after_sync();
            issue_transform(d1, s_a_hi, s_a_lo, s_b_hi, s_b_lo);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        d1_to_staging(d1, stg);
        tc::tc_fence_before_sync();
        __syncthreads();
        {
            const float4 c0 = ld4(cst + 4 * c4), v1 = ld4(u1 + 4 * c4), v0 = ld4(u0 + 4 * c4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int rl = rbase + 8 * i;
                if (rl < rows) {
                    float4 v = ld4(stg + rl * STG_LD + 4 * c4);
                    float c = cw[rl], d = cw[64 + rl];
                    float4 b;
                    b.x = v.x + (c0.x + (c * v1.x + d * v0.x));
                    b.y = v.y + (c0.y + (c * v1.y + d * v0.y));
                    b.z = v.z + (c0.z + (c * v1.z + d * v0.z));
                    b.w = v.w + (c0.w + (c * v1.w + d * v0.w));
                    size_t o = (size_t)(row0 + rl) * 64 + 4 * c4;
                    st4(base + o, b);
                    st4(mu1 + o, make_float4(relu(b.x), relu(b.y), relu(b.z), relu(b.w)));
                }
            }
        }
        __syncthreads();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(d1);
}

// ------------------------------------------------------------------------------------
// last backward stage (same math as k_embed_final_bwd, s2v_embed.cu):
//   D = sum_t du^t ;  dtheta1b += D^T h ;  dh = (D theta1b) * [h > 0] ;  dtheta1a += dh^T x ;
//   column sums of D, c*D, (Npad-c)*D, dh
// D and h tiles are MN-major in shared memory (operands of the reduction D2 = [D_hi;D_lo]^T h),
// the row transform D * theta1b takes D from TMEM.  Partial layout: EmbedPartial<64> (s2v_embed.cu).
// ------------------------------------------------------------------------------------
constexpr int EP_TH1B = 64 * 64, EP_COLD = 2 * 64 * 64, EP_A1 = EP_COLD + 64, EP_A0 = EP_A1 + 64,
              EP_B1A = EP_A0 + 64, EP_TH1A = EP_B1A + 64;

__global__ void __launch_bounds__(TC_THREADS)
k_embed_final_bwd_tc(s2v_graph_t g, s2v_embed_params_t prm, const float* __restrict__ x, const float* __restrict__ dus,
                     const float* __restrict__ du_last, int T, float* __restrict__ partial, int partial_stride,
                     int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;                        // theta1b^T: tile(row k, col n) = theta1b[n][k]
    uint8_t* b_lo = b_hi + TILE_BYTES;
    uint8_t* d_hi = b_lo + TILE_BYTES;           // D tile (d_lo must follow d_hi); later staging / dh
    uint8_t* d_lo = d_hi + TILE_BYTES;
    uint8_t* h_hi = d_lo + TILE_BYTES;           // h tile; later the D2 exchange buffer
    uint8_t* h_lo = h_hi + TILE_BYTES;
    float* W1a = reinterpret_cast<float*>(h_lo + TILE_BYTES);   // [F][64]
    float* xs = W1a + S2V_MAX_F * 64;                           // [64][F]
    float* b1a = xs + TCR * S2V_MAX_F;                          // [64]
    float* cw = b1a + 64;                                       // [64] in-degree c_r, [64] Npad - c_r
    TcShared* sh = reinterpret_cast<TcShared*>(cw + 128);
    float* stg = reinterpret_cast<float*>(d_hi);
    float* xbuf = reinterpret_cast<float*>(h_hi);
    const int F = prm.node_dim;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const size_t plane = (size_t)g.num_nodes * 64;

    if (warp == 0) tc::tmem_alloc<256>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles(b_hi, b_lo, prm.theta1b_w, true);
    for (int e = t; e < 64 * F; e += TC_THREADS) { int n = e / F, f = e - n * F; W1a[f * 64 + n] = __ldg(prm.theta1a_w + e); }
    if (t < 64) b1a[t] = __ldg(prm.theta1a_b + t);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64, a_tm = sh->tmem_base + 128;
    const uint32_t s_d_hi = tc::smem_u32(d_hi), s_b_hi = tc::smem_u32(b_hi), s_b_lo = tc::smem_u32(b_lo),
                   s_h_hi = tc::smem_u32(h_hi), s_h_lo = tc::smem_u32(h_lo);
    uint32_t phase = 0;
    float racc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) racc[i] = 0.f;
    float pD = 0.f, pA1 = 0.f, pA0 = 0.f, pB1a = 0.f;        // threads 0..63: column t
    constexpr int MAXJ = S2V_MAX_F / 2;                       // thread -> column t % 64, features t / 64 + 2 j
    const int kcol = t & 63, fgrp = t >> 6;
    float pth1a[MAXJ];
#pragma unroll
    for (int i = 0; i < MAXJ; ++i) pth1a[i] = 0.f;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
#ifdef S2V_TC_TIMELINE
        const int tl_n = (tile - blockIdx.x) / gridDim.x - 2;
#endif
        TL2(0);
        // x tile and per-row degree weights
        for (int e = t; e < TCR * F; e += TC_THREADS) xs[e] = e < rows * F ? __ldg(x + (size_t)row0 * F + e) : 0.f;
        if (t < TCR) {
            float c = 0.f, d = 0.f;
            if (t < rows) {
                int li, off, gi;
                row_info(g, row0 + t, li, off, gi);
                c = (float)__ldg(g.indeg + li);
                d = (float)graph_npad(g, gi) - c;
            }
            cw[t] = c;
            cw[64 + t] = d;
        }
        // D tile: fixed summation order t = 1..T
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            float4 s4 = f4zero();
            if (rl < rows) {
                size_t o = (size_t)(row0 + rl) * 64 + 4 * c4;
                for (int tt = 0; tt < T - 1; ++tt) s4 = f4add(s4, ldg4(dus + tt * plane + o));
                s4 = f4add(s4, ldg4(du_last + o));
            }
            store_split_mn(d_hi, d_lo, rl, c4, s4);
        }
        TL2(1);
        __syncthreads();                        // xs complete
        // h tile (rows past the end give relu(b1a) but their D is zero)
        {
            const int c4 = t & 15, rbase = t >> 4;
            float4 h[8];
            const float4 bb = ld4(b1a + 4 * c4);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = bb;
            for (int f = 0; f < F; ++f) {
                const float4 w = ld4(W1a + f * 64 + 4 * c4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float xv = xs[(rbase + 8 * i) * F + f];
                    h[i].x = fmaf(xv, w.x, h[i].x); h[i].y = fmaf(xv, w.y, h[i].y);
                    h[i].z = fmaf(xv, w.z, h[i].z); h[i].w = fmaf(xv, w.w, h[i].w);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                store_split_mn(h_hi, h_lo, rbase + 8 * i, c4, make_float4(relu(h[i].x), relu(h[i].y), relu(h[i].z), relu(h[i].w)));
        }
        tc::fence_proxy_async_smem();
        TL2(2);
        __syncthreads();
        row_to_tmem(d_hi, a_tm);
        row_to_tmem(d_lo, a_tm + 64);
        tc::tmem_st_wait();
        tc::tc_fence_before_sync();
        __syncthreads();
        TL2(3);
        if (t == 0) {
            tc::tc_fence_after_sync();
            issue_transform_ts(d1, a_tm, s_b_hi, s_b_lo);
            issue_reduce(d2, s_d_hi, s_h_hi, s_h_lo);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        TL2(4);
        // column sums of D on the CUDA cores while the tensor core works (D = hi + lo)
        if (t < 64) {
            const int c4 = t >> 2, w = t & 3;
            for (int r = 0; r < rows; ++r) {
                uint32_t off = tc::tile_off_mn<TCR>(r, c4) + 4 * w;
                float v = *reinterpret_cast<const float*>(d_hi + off) + *reinterpret_cast<const float*>(d_lo + off);
                pD += v;
                pA1 = fmaf(cw[r], v, pA1);
                pA0 = fmaf(cw[64 + r], v, pA0);
            }
        }
        TL2(5);
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        TL2(6);
        __syncthreads();                        // column sums done before the D tiles become staging
        d1_to_staging(d1, stg);
        tc::tc_fence_before_sync();
        __syncthreads();
        // dh = (D theta1b) * [h > 0], kept in staging for dtheta1a / db1a
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            float4 v = ld4(stg + rl * STG_LD + 4 * c4);
            float4 h = *reinterpret_cast<const float4*>(h_hi + tc::tile_off_mn<TCR>(rl, c4));
            float4 dh = make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f, h.w > 0.f ? v.w : 0.f);
            if (rl >= rows) dh = f4zero();
            st4(stg + rl * STG_LD + 4 * c4, dh);
        }
        __syncthreads();
        TL2(7);
        if (t < 64)
            for (int r = 0; r < rows; ++r) pB1a += stg[r * STG_LD + t];
        for (int r = 0; r < rows; ++r) {
            const float a = stg[r * STG_LD + kcol];
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) {
                int f = fgrp + 2 * j;
                if (f < F) pth1a[j] = fmaf(a, xs[r * F + f], pth1a[j]);
            }
        }
        TL2(8);
        flush_d2(d2, xbuf, racc);               // h tiles are dead: exchange buffer
        __syncthreads();
        TL2(9);
    }
    float* my = partial + (size_t)blockIdx.x * partial_stride;
    {
        float* dst = my + EP_TH1B + (32 * (warp & 1) + lane) * 64 + 32 * (warp >> 1);
#pragma unroll
        for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]));
    }
    if (t < 64) { my[EP_COLD + t] = pD; my[EP_A1 + t] = pA1; my[EP_A0 + t] = pA0; my[EP_B1A + t] = pB1a; }
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        int f = fgrp + 2 * j;
        if (f < F) my[EP_TH1A + kcol * F + f] = pth1a[j];
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(sh->tmem_base);
}

// ------------------------------------------------------------------------------------
// debug / self-test kernel: out = A * W^T (mode 0), out = A * W (mode 1), out[64x64] = A^T * B (mode 2),
// out = A * W with the A operand staged through TMEM (mode 3, the backward round's transform)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
k_tc_selftest(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int rows) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;
    uint8_t* b_lo = b_hi + TILE_BYTES;
    uint8_t* a_hi = b_lo + TILE_BYTES;
    uint8_t* a_lo = a_hi + TILE_BYTES;
    uint8_t* m_hi = a_lo + TILE_BYTES;
    uint8_t* m_lo = m_hi + TILE_BYTES;
    TcShared* sh = reinterpret_cast<TcShared*>(m_lo + TILE_BYTES);
    float* stg = reinterpret_cast<float*>(a_hi);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (warp == 0) tc::tmem_alloc<256>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    if (mode != 2) load_weight_tiles(b_hi, b_lo, B, mode != 0);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64, a_tm = sh->tmem_base + 128;
    uint32_t phase = 0;
    float racc[32];
    for (int i = 0; i < 32; ++i) racc[i] = 0.f;
    const int num_tiles = (rows + TCR - 1) / TCR;
    for (int tile = 0; tile < num_tiles; ++tile) {
        const int row0 = tile * TCR, nr = min(TCR, rows - row0);
        for (int e = t; e < TCR * 16; e += TC_THREADS) {
            int rl = e >> 4, c4 = e & 15;
            float4 a = rl < nr ? ldg4(A + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero();
            if (mode < 2) store_split(a_hi, a_lo, rl, c4, a);
            else store_split_mn(a_hi, a_lo, rl, c4, a);
            if (mode == 2) {
                float4 b = rl < nr ? ldg4(B + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero();
                store_split_mn(m_hi, m_lo, rl, c4, b);
            }
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (mode == 3) {
            row_to_tmem(a_hi, a_tm);
            row_to_tmem(a_lo, a_tm + 64);
            tc::tmem_st_wait();
            tc::tc_fence_before_sync();
            __syncthreads();
        }
        if (t == 0) {
            tc::tc_fence_after_sync();
            if (mode < 2) issue_transform(d1, tc::smem_u32(a_hi), tc::smem_u32(a_lo), tc::smem_u32(b_hi), tc::smem_u32(b_lo));
            else if (mode == 2) issue_reduce(d2, tc::smem_u32(a_hi), tc::smem_u32(m_hi), tc::smem_u32(m_lo));
            else issue_transform_ts(d1, a_tm, tc::smem_u32(b_hi), tc::smem_u32(b_lo));
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();                           // lanes 1-31 of warp 0 must not sleep in try_wait before lane 0 has issued
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        if (mode != 2) {
            d1_to_staging(d1, stg);
            tc::tc_fence_before_sync();
            __syncthreads();
            for (int e = t; e < TCR * 16; e += TC_THREADS) {
                int rl = e >> 4, c4 = e & 15;
                if (rl < nr) st4(out + (size_t)(row0 + rl) * 64 + 4 * c4, ld4(stg + rl * STG_LD + 4 * c4));
            }
        } else {
            flush_d2(d2, stg, racc);
        }
        __syncthreads();
    }
    if (mode == 2) {
        float* my = out + (32 * (warp & 1) + lane) * 64 + 32 * (warp >> 1);
        for (int i = 0; i < 32; ++i) my[i] = racc[i];
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(sh->tmem_base);
}

constexpr int FWD_SMEM = 4 * TILE_BYTES + 64 + 1024;
constexpr int FIRST_SMEM = 4 * TILE_BYTES + (S2V_MAX_F * 64 + TCR * S2V_MAX_F + 6 * 64 + 128) * 4 + 64 + 1024;
constexpr int FIN_SMEM = 6 * TILE_BYTES + (S2V_MAX_F * 64 + TCR * S2V_MAX_F + 64 + 128) * 4 + 64 + 1024;
constexpr int BWD_SMEM = 6 * TILE_BYTES + 64 + 1024;

// Resident CTAs per SM from the kernel's own resource use (registers, shared memory, TMEM columns).
int tc_grid(const void* kernel, int smem, int tmem_cols, int tiles) {
    int dev = 0, sms = 0, smem_sm = 0, regs_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    cudaFuncAttributes fa;
    int per_sm = 1;
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) {
        int regs_cta = ((fa.numRegs + 7) / 8 * 8) * TC_THREADS;
        int by_regs = regs_cta > 0 ? regs_sm / regs_cta : 1;
        int by_smem = smem_sm / (smem + (int)fa.sharedSizeBytes + 1024);      // 1 KB reserved per CTA
        int by_tmem = 512 / tmem_cols;
        per_sm = by_regs < by_smem ? by_regs : by_smem;
        if (by_tmem < per_sm) per_sm = by_tmem;
        if (per_sm < 1) per_sm = 1;
    }
    int grid = sms * per_sm;
    if (grid > S2V_MAX_GRID) grid = S2V_MAX_GRID;
    if (grid > tiles) grid = tiles;
    return grid < 1 ? 1 : grid;
}

}  // namespace

// Launchers used by s2v_embed.cu (p = 64 only).  grid_out: number of CTAs (= dtheta2 partials).
int s2v_tc_round_forward(const s2v_graph_t& g, const float* theta2_w, const float* base, const float* mu_prev,
                         float* mu_next, cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    int grid = tc_grid((const void*)k_embed_round_tc, FWD_SMEM, 64, tiles);
    k_embed_round_tc<<<grid, TC_THREADS, FWD_SMEM, st>>>(g, theta2_w, base, mu_prev, mu_next, tiles);
    return (int)cudaGetLastError();
}

int s2v_tc_round_backward_grid(int num_nodes) {
    const int tiles = (num_nodes + TCR - 1) / TCR;
    cudaFuncSetAttribute(k_embed_round_bwd_tc256, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = 2 * sms;                          // two 8-warp CTAs per SM (96 KB of tiles, 256 TMEM columns each)
    if (grid > S2V_MAX_GRID) grid = S2V_MAX_GRID;
    return grid < tiles ? grid : (tiles < 1 ? 1 : tiles);
}

int s2v_tc_round_backward(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                          float* du_prev, float* partial, int partial_stride, int accumulate, int grid, cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_bwd_tc256, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    k_embed_round_bwd_tc256<<<grid, TC_THREADS2, BWD_SMEM, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride,
                                                                accumulate, tiles);
    return (int)cudaGetLastError();
}

#ifdef S2V_TC_TIMELINE
extern "C" int s2v_debug_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 16 * 8);
}
#endif

int s2v_tc_embed_first(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, float* base, float* mu1,
                       cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_first_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FIRST_SMEM));
    int grid = tc_grid((const void*)k_embed_first_tc, FIRST_SMEM, 64, tiles);
    k_embed_first_tc<<<grid, TC_THREADS, FIRST_SMEM, st>>>(g, prm, x, base, mu1, tiles);
    return (int)cudaGetLastError();
}

int s2v_tc_final_backward_grid(int num_nodes) {
    const int tiles = (num_nodes + TCR - 1) / TCR;
    cudaFuncSetAttribute(k_embed_final_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FIN_SMEM);
    return tc_grid((const void*)k_embed_final_bwd_tc, FIN_SMEM, 256, tiles);
}

int s2v_tc_final_backward(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, const float* dus,
                          const float* du_last, float* partial, int partial_stride, int grid, cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_final_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FIN_SMEM));
    k_embed_final_bwd_tc<<<grid, TC_THREADS, FIN_SMEM, st>>>(g, prm, x, dus, du_last, prm.rounds, partial, partial_stride, tiles);
    return (int)cudaGetLastError();
}

extern "C" int s2v_debug_tc_gemm(int32_t mode, const float* A, const float* B, float* out, int64_t rows, void* stream) {
    if (!A || !B || !out || rows < 1 || mode < 0 || mode > 3) return S2V_ERR_BAD_ARG;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    k_tc_selftest<<<1, TC_THREADS, BWD_SMEM, (cudaStream_t)stream>>>(mode, A, B, out, (int)rows);
    return (int)cudaGetLastError();
}
