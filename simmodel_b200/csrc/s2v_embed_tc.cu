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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_fp16.h>
#include "s2v_common.cuh"
#include "s2v_tc.cuh"

#ifndef S2V_HEAD_H2_DEFAULT
#define S2V_HEAD_H2_DEFAULT 5          /* Q head: 0 = bf16x3 (k_head_tc), 4..6 = fp16x2 with that many CTAs per SM */
#endif
#ifndef S2V_FWD_H2_DEFAULT
#define S2V_FWD_H2_DEFAULT 5           /* forward round: 0 = bf16x3 (k_embed_round_tc16), 4..6 = fp16x2 with that many CTAs per SM */
#endif

#ifdef S2V_TC_TIMELINE
__device__ long long g_timeline[16 * 8];
#define TL(slot) do { if (blockIdx.x == 0 && threadIdx.x == 32 && tl_n < 8) g_timeline[tl_n * 16 + (slot)] = clock64(); } while (0)
#define TL3(slot) do { if (blockIdx.x == 0 && threadIdx.x == 32 && tl_n >= 0 && tl_n < 8) g_timeline[tl_n * 16 + (slot)] = clock64(); } while (0)
#define TLW(cond, k, slot) do { if (blockIdx.x == 0 && (cond) && (k) >= 4 && (k) < 12) g_timeline[((k) - 4) * 16 + (slot)] = clock64(); } while (0)
#define TL2(slot) do { if (blockIdx.x == 0 && threadIdx.x == 32 && tl_n >= 0 && tl_n < 8) g_timeline[tl_n * 16 + (slot)] = clock64(); } while (0)
#define TLS(cond, k, slot) do { if (blockIdx.x == 0 && (cond) && (k) % 50 == 0 && (k) / 50 < 8) g_timeline[((k) / 50) * 16 + (slot)] = clock64(); } while (0)
#else
#define TLS(cond, k, slot) do {} while (0)
#define TL(slot) do {} while (0)
#define TL2(slot) do {} while (0)
#define TLW(cond, k, slot) do {} while (0)
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
// bf16x3 building blocks (backward kernels): 64-row x 64-column bf16 tiles of 8 KB, three per fp32 tile.
// ------------------------------------------------------------------------------------
constexpr uint32_t TILE16 = TCR * 128;          // one bf16 tile
constexpr uint32_t TILE16X3 = 3 * TILE16;

// split a float4 chunk of row r into the three bf16 tiles at t1, t1 + TILE16, t1 + 2 TILE16
__device__ __forceinline__ void store_split3(uint8_t* t1, int r, int c4, const float4& v) {
    uint2 p1, p2, p3;
    tc::split3(v, p1, p2, p3);
    const uint32_t off = tc::tile_off16(r, c4);
    *reinterpret_cast<uint2*>(t1 + off) = p1;
    *reinterpret_cast<uint2*>(t1 + TILE16 + off) = p2;
    *reinterpret_cast<uint2*>(t1 + 2 * TILE16 + off) = p3;
}
// the six (i, j) products with i + j <= 4, smallest first
__device__ __forceinline__ void pair6(int p, int& i, int& j) {
    const int ii[6] = {2, 0, 1, 1, 0, 0}, jj[6] = {0, 2, 1, 0, 1, 0};
    i = ii[p]; j = jj[p];
}
// D[64x64] (+)= A * B^T, A rows = tile rows (K-major), B rows = output columns (K-major); 24 MMAs
__device__ __forceinline__ void issue_transform16(uint32_t d_tmem, uint32_t a1, uint32_t b1) {
    constexpr uint32_t idesc = tc::idesc_bf16(64, 64, 0, 0);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        int i, j;
        pair6(p, i, j);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_bf16(d_tmem, tc::desc_k16(a1 + i * TILE16, ks), tc::desc_k16(b1 + j * TILE16, ks), idesc, p > 0 || ks > 0);
    }
}
// D[64x64] (+)= A^T * B over the 64 rows of the tiles (both MN-major views); 24 MMAs
__device__ __forceinline__ void issue_reduce16(uint32_t d_tmem, uint32_t a1, uint32_t b1, bool accumulate) {
    constexpr uint32_t idesc = tc::idesc_bf16(64, 64, 1, 1);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        int i, j;
        pair6(p, i, j);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_bf16(d_tmem, tc::desc_mn16(a1 + i * TILE16, ks, TILE16), tc::desc_mn16(b1 + j * TILE16, ks, TILE16), idesc,
                         accumulate || p > 0 || ks > 0);
    }
}
// W [64][64] row-major -> three bf16 tiles with tile(row, col) = transpose ? W[col][row] : W[row][col]
__device__ __forceinline__ void load_weight_tiles16(uint8_t* t1, const float* __restrict__ W, bool transpose, int nthreads) {
    for (int e = threadIdx.x; e < 64 * 16; e += nthreads) {
        const int row = e >> 4, c4 = e & 15;
        float4 v;
        if (transpose) v = make_float4(__ldg(W + (4 * c4) * 64 + row), __ldg(W + (4 * c4 + 1) * 64 + row),
                                       __ldg(W + (4 * c4 + 2) * 64 + row), __ldg(W + (4 * c4 + 3) * 64 + row));
        else v = ldg4(W + row * 64 + 4 * c4);
        store_split3(t1, row, c4, v);
    }
}

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
        if (warp == 0) {
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
        if (warp == 0) {                            // the reduction only needs the shared-memory tiles: start it now
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
        if (warp == 0) {
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
__global__ void __launch_bounds__(TC_THREADS, 4)
k_embed_first_tc(s2v_graph_t g, s2v_embed_params_t prm, const float* __restrict__ x, float* __restrict__ base,
                 float* __restrict__ mu1, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta1b [n][k], three bf16 tiles (bf16x3: 24 mantissa bits)
    uint8_t* a1 = w1 + TILE16X3;                 // h tile, three bf16 tiles; reused as staging
    const int F = prm.node_dim;
    const int FS = (F + 3) & ~3;                 // x rows and theta1a are padded to a multiple of four features (zeros)
    float* W1a = reinterpret_cast<float*>(a1 + TILE16X3);       // [FS][64]   (sized by the actual F: F = 7 leaves room
    float* xs = W1a + FS * 64;                                  // [64][FS]    for a third CTA per SM, first_smem_bytes)
    float* vec = xs + TCR * FS;                                 // b1a, cst, u1, u0, r1, r0 (6 x 64)
    float* cw = vec + 6 * 64;                                   // [64] c_r, [64] Npad - c_r
    TcShared* sh = reinterpret_cast<TcShared*>(cw + 128);
    float* stg = reinterpret_cast<float*>(a1);
    float* b1a = vec, *cst = vec + 64, *u1 = vec + 128, *u0 = vec + 192, *r1 = vec + 256, *r0 = vec + 320;
    const int t = threadIdx.x, warp = t >> 5;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles16(w1, prm.theta1b_w, false, TC_THREADS);
    for (int e = t; e < 64 * FS; e += TC_THREADS) { int f = e >> 6, n = e & 63; W1a[e] = f < F ? __ldg(prm.theta1a_w + n * F + f) : 0.f; }
    if (t < 64) {
        b1a[t] = __ldg(prm.theta1a_b + t);
        float w4 = __ldg(prm.theta4_w + t), b4 = __ldg(prm.theta4_b + t);
        r1[t] = relu(w4 + b4);
        r0[t] = relu(b4);
    }
    __syncthreads();
    if (t < 64) {
        u1[t] = row_dot<64>(prm.theta3_w + t * 64, r1, 0.f);
        u0[t] = row_dot<64>(prm.theta3_w + t * 64, r0, 0.f);
        cst[t] = __ldg(prm.theta1b_b + t) + __ldg(prm.theta2_b + t) + __ldg(prm.theta3_b + t);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1);
    uint32_t phase = 0;
    const int c4 = t & 15, rbase = t >> 4;       // this thread's chunks: rows rbase + 8*i, columns 4*c4..4*c4+3

    // Row t of the CTA's next tile, requested one tile ahead: features (F <= 8), in-degree c and padded size Npad as RAW
    // integers -- converting them here would make the 64 staging threads (and the barrier behind them) wait for the index
    // loads every tile (1.3 k cycles on the timeline).  The (graph, local row) of a row advances by a fixed stride per tile:
    // the only division is done once.
    float xr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int ci = 0, npi = 0, nli = 0, ngi = 0, step_q = 0, step_r = 0;
    if (g.period > 0) {
        const int stride = (int)gridDim.x * TCR, r = (int)blockIdx.x * TCR + (t < TCR ? t : 0);
        step_q = stride / g.period; step_r = stride - step_q * g.period;
        ngi = r / g.period; nli = r - ngi * g.period;
    }
    auto prefetch = [&](int tl) {                // tiles are requested in the order the CTA visits them
        const int r = tl * TCR + t;
        ci = 0; npi = g.npad_scalar;
#pragma unroll
        for (int f = 0; f < 8; ++f) xr[f] = 0.f;
        if (r < g.num_nodes) {
            if (g.period > 0) {
                ci = __ldg(g.indeg + nli);
                if (g.npad) npi = __ldg(g.npad + ngi);
            } else {
                ci = __ldg(g.indeg + r);
                if (g.npad) npi = __ldg(g.npad + __ldg(g.gid + r));
            }
            if (F <= 8) {
#pragma unroll
                for (int f = 0; f < 8; ++f) if (f < F) xr[f] = __ldg(x + (size_t)r * F + f);
            }
        }
        nli += step_r; ngi += step_q;
        if (g.period > 0 && nli >= g.period) { nli -= g.period; ++ngi; }
    };
    if (t < TCR && blockIdx.x < num_tiles) prefetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
#ifdef S2V_TC_TIMELINE
        const int tl_n = (tile - blockIdx.x) / gridDim.x - 2;
#endif
        TL2(0);
        // thread t < 64 stages row t: x (padded to FS floats), in-degree c and Npad - c; for F <= 8 the values were
        // loaded one tile ahead
        if (t < TCR) {
            if (F > 8) {
                for (int f = 0; f < FS; ++f) xs[t * FS + f] = (t < rows && f < F) ? __ldg(x + (size_t)(row0 + t) * F + f) : 0.f;
            } else {
                st4(xs + t * FS, make_float4(xr[0], xr[1], xr[2], xr[3]));
                if (FS > 4) st4(xs + t * FS + 4, make_float4(xr[4], xr[5], xr[6], xr[7]));
            }
            const float cn = (float)ci;
            cw[t] = cn;
            cw[64 + t] = (float)npi - cn;                // rows past the end: x = 0 and c = 0 (their outputs are not stored)
            prefetch(tile + gridDim.x);                  // zeros past the last row
        }
        __syncthreads();
        TL2(1);
        {
            float4 h[8];
            const float4 bb = ld4(b1a + 4 * c4);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = bb;
            for (int f0 = 0; f0 < FS; f0 += 4) {     // padded features are zero in x and in theta1a
                const float4 w0 = ld4(W1a + f0 * 64 + 4 * c4), w1 = ld4(W1a + (f0 + 1) * 64 + 4 * c4),
                             w2 = ld4(W1a + (f0 + 2) * 64 + 4 * c4), w3 = ld4(W1a + (f0 + 3) * 64 + 4 * c4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 xv = ld4(xs + (rbase + 8 * i) * FS + f0);
                    h[i].x = fmaf(xv.x, w0.x, h[i].x); h[i].y = fmaf(xv.x, w0.y, h[i].y);
                    h[i].z = fmaf(xv.x, w0.z, h[i].z); h[i].w = fmaf(xv.x, w0.w, h[i].w);
                    h[i].x = fmaf(xv.y, w1.x, h[i].x); h[i].y = fmaf(xv.y, w1.y, h[i].y);
                    h[i].z = fmaf(xv.y, w1.z, h[i].z); h[i].w = fmaf(xv.y, w1.w, h[i].w);
                    h[i].x = fmaf(xv.z, w2.x, h[i].x); h[i].y = fmaf(xv.z, w2.y, h[i].y);
                    h[i].z = fmaf(xv.z, w2.z, h[i].z); h[i].w = fmaf(xv.z, w2.w, h[i].w);
                    h[i].x = fmaf(xv.w, w3.x, h[i].x); h[i].y = fmaf(xv.w, w3.y, h[i].y);
                    h[i].z = fmaf(xv.w, w3.z, h[i].z); h[i].w = fmaf(xv.w, w3.w, h[i].w);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                store_split3(a1, rbase + 8 * i, c4, make_float4(relu(h[i].x), relu(h[i].y), relu(h[i].z), relu(h[i].w)));
        }
        TL2(2);
        tc::fence_proxy_async_smem();
        __syncthreads();
        TL2(3);
        if (warp == 0) {
            tc::tc_fence_after_sync();
            issue_transform16(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        TL2(4);
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        TL2(5);
        d1_to_staging(d1, stg);
        tc::tc_fence_before_sync();
        __syncthreads();
        TL2(6);
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
        TL2(7);
        __syncthreads();
        TL2(8);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(d1);
}

// ------------------------------------------------------------------------------------
// Q / logit head (same math as k_head, s2v_head.cu): L = mu theta7^T + b7 on the tensor cores (3xTF32),
//   q_i = w5[:p] . relu(G_g(i)) + w5[p:] . relu(L_i) + b5
// The row owner (lane < 16 of each warp, UMMA M = 64 layout) reads its accumulator row from TMEM and does both
// dot products; G comes from k_pool (per graph, L2-resident).  mu rows are loaded one tile ahead.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 4)
k_head_tc(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ mu, const float* __restrict__ gstate,
          float* __restrict__ q, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta7 [n][k], three bf16 tiles (bf16x3: 24 mantissa bits)
    uint8_t* a1 = w1 + TILE16X3;                 // mu tile, three bf16 tiles
    float* vec = reinterpret_cast<float*>(a1 + TILE16X3);      // b7, w5a, w5b
    TcShared* sh = reinterpret_cast<TcShared*>(vec + 3 * 64);
    float* b7 = vec, *w5a = vec + 64, *w5b = vec + 128;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles16(w1, prm.theta7_w, false, TC_THREADS);
    if (t < 64) {
        b7[t] = __ldg(prm.theta7_b + t);
        w5a[t] = __ldg(prm.theta5_w + t);
        w5b[t] = __ldg(prm.theta5_w + 64 + t);
    }
    const float b5 = __ldg(prm.theta5_b);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1);
    uint32_t phase = 0;
    float4 m[8];
    auto load_rows = [&](int tile) {
        const int row0 = tile * TCR, rows = min(TCR, g.num_nodes - row0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            m[i] = rl < rows ? ldg4(mu + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero();
        }
    };
    if (blockIdx.x < num_tiles) load_rows(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i;
            store_split3(a1, e >> 4, e & 15, m[i]);
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            issue_transform16(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        if (tile + (int)gridDim.x < num_tiles) load_rows(tile + gridDim.x);     // in flight during the MMAs and the epilogue
        // graph part of the row owner's q while the tensor core works
        const int rl = 16 * warp + (lane & 15);
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
                sl = fmaf(wb.x, relu(v[4 * c] + bb.x), sl); sl = fmaf(wb.y, relu(v[4 * c + 1] + bb.y), sl);
                sl = fmaf(wb.z, relu(v[4 * c + 2] + bb.z), sl); sl = fmaf(wb.w, relu(v[4 * c + 3] + bb.w), sl);
            }
        }
        if (lane < 16 && rl < rows) q[row0 + rl] = (sg + sl) + b5;
        tc::tc_fence_before_sync();
        __syncthreads();                        // the accumulator and the A tiles are free again
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
        if (warp == 0) {
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
        if (warp == 0) {
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

// ------------------------------------------------------------------------------------
// forward round on bf16x3 (default): same pipeline as k_embed_round_tc, operands split into three bf16 tiles
// (24 mantissa bits instead of the 21 of 3xTF32) -- 48 KB of tiles instead of 64 KB, so four CTAs fit on an SM.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 4)
k_embed_round_tc16(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ base,
                   const float* __restrict__ mu_prev, float* __restrict__ mu_next, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta2 [n][k], three bf16 tiles
    uint8_t* a1 = w1 + TILE16X3;                 // gathered tile, three bf16 tiles; reused as staging
    TcShared* sh = reinterpret_cast<TcShared*>(a1 + TILE16X3);
    float* stg = reinterpret_cast<float*>(a1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles16(w1, theta2_w, false, TC_THREADS);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1);
    uint32_t phase = 0;
    const int rl0 = warp * 16 + 8 * (lane >> 4);
    Gather8Meta meta;
    if (blockIdx.x < num_tiles)
        gather8_prefetch(meta, g, g.rowptr, g.col, blockIdx.x * TCR, min(TCR, g.num_nodes - blockIdx.x * TCR), rl0);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        gather8_consume(meta, g.col, mu_prev, rl0,
                        [&](int rl, const float4& acc) { store_split3(a1, rl, lane & 15, acc); });
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            issue_transform16(d1, s_a, s_w);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        float4 bv[8];                            // base rows of the epilogue while the tensor core works: requested FIRST -- the
#pragma unroll                                   // metadata prefetch below stalls on its own dependent loads (rowptr -> col)
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            bv[i] = rl < rows ? ldg4(base + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero();
        }
        {
            const int nt = tile + gridDim.x;     // CSR metadata of the next tile: overlaps MMA + epilogue.  (Extents a tile earlier
            // than the ids -- no dependent stall at all -- measured 0.936 ms against 0.895 ms: the extra registers spill.)
            if (nt < num_tiles) gather8_prefetch(meta, g, g.rowptr, g.col, nt * TCR, min(TCR, g.num_nodes - nt * TCR), rl0);
        }
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        d1_to_staging(d1, stg);                 // the A tiles are dead once the MMAs have completed
        tc::tc_fence_before_sync();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
            if (rl < rows) {
                float4 v = ld4(stg + rl * STG_LD + 4 * c4);
                st4(mu_next + (size_t)(row0 + rl) * 64 + 4 * c4,
                    make_float4(relu(v.x + bv[i].x), relu(v.y + bv[i].y), relu(v.z + bv[i].z), relu(v.w + bv[i].w)));
            }
        }
        __syncthreads();                        // staging is overwritten by the next gather
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(d1);
}

// ------------------------------------------------------------------------------------
// backward round in the shape of the forward round (r02d): 4-warp CTAs, three per SM, no roles.
//   du^{t-1} = (g theta2) * [mu^{t-1} > 0],  dtheta2 += g^T mu^{t-1},  g = gather of du^t over the transposed CSR.
// The forward round reaches 0.87 of the copy peak with four independent 4-warp CTAs per SM whose gather / MMA /
// epilogue phases interleave; the warp-specialised backward kernel (one 16-warp CTA per SM, below) sits at 0.64 because
// only its eight producer warps have loads in flight.  Here every warp gathers (16 neighbour loads + 8 own-row loads per
// lane) and the SM holds three such CTAs: 72 KB of operand tiles (theta2^T, g, mu^{t-1}: three bf16 tiles each, one
// SWIZZLE_128B tile serves the K-major and the MN-major view) and 128 TMEM columns per CTA (D1 = g theta2, D2 = g^T mu,
// M = N = 64, six products each).  D2 is read out every `d2_group` tiles (the tensor core's accumulator truncates:
// chains stay short) into 32 registers per thread: lanes 0-15 of a warp own the TMEM lanes M = 64 uses and keep
// columns 0-31, lanes 16-31 take columns 32-63 by shuffle.  Partials per CTA, summed in fixed order by
// k_reduce_partials: bit-stable run to run.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 3)
k_embed_round_bwd_tc16(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ du_t,
                       const float* __restrict__ mu_prev, float* __restrict__ du_prev, float* __restrict__ partial,
                       int partial_stride, int accumulate, int num_tiles, int d2_group) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta2^T: tile(row k, col n) = theta2[n][k]
    uint8_t* a1 = w1 + TILE16X3;                 // gathered du^t tile; reused as staging
    uint8_t* m1 = a1 + TILE16X3;                 // own mu^{t-1} tile
    TcShared* sh = reinterpret_cast<TcShared*>(m1 + TILE16X3);
    float* stg = reinterpret_cast<float*>(a1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const unsigned FULL = 0xffffffffu;

    if (warp == 0) tc::tmem_alloc<128>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    load_weight_tiles16(w1, theta2_w, true, TC_THREADS);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1), s_m = tc::smem_u32(m1);
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    uint32_t phase = 0;
    const int c4 = lane & 15, rl0 = warp * 16 + 8 * (lane >> 4);
    float racc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) racc[i] = 0.f;
    int chained = 0;                             // tiles accumulated in D2 since it was last read out
#ifdef S2V_BWD16_EXPERIMENTS
    const int exp_bits = d2_group >> 8;          // 1: no reduce MMAs, 2: no MMAs at all, 4: no m-tile stores, 8: no D2 read-out, 16: no du stores
    d2_group &= 255;
#else
    constexpr int exp_bits = 0;
#endif
    Gather8Meta meta;
    if (blockIdx.x < num_tiles)
        gather8_prefetch(meta, g, g.rowptr_t, g.col_t, blockIdx.x * TCR, min(TCR, g.num_nodes - blockIdx.x * TCR), rl0);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        uint32_t mbits = 0;                      // [mu^{t-1} > 0] of this lane's 8 rows x 4 columns
        {
            float4 mv[8];                        // own rows first: they are in flight together with the gather
#pragma unroll
            for (int i = 0; i < 8; ++i)
                mv[i] = rl0 + i < rows ? ldg4(mu_prev + (size_t)(row0 + rl0 + i) * 64 + 4 * c4) : f4zero();
            gather8_consume(meta, g.col_t, du_t, rl0,
                            [&](int rl, const float4& acc) { store_split3(a1, rl, c4, acc); });
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (!(exp_bits & 4)) store_split3(m1, rl0 + i, c4, mv[i]);
                mbits |= (mv[i].x > 0.f ? 1u : 0u) << (4 * i) | (mv[i].y > 0.f ? 2u : 0u) << (4 * i) |
                         (mv[i].z > 0.f ? 4u : 0u) << (4 * i) | (mv[i].w > 0.f ? 8u : 0u) << (4 * i);
            }
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            if (!(exp_bits & 2)) issue_transform16(d1, s_a, s_w);
            if (!(exp_bits & 3)) issue_reduce16(d2, s_a, s_m, chained != 0);
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
        d1_to_staging(d1, stg);                 // the g tiles are dead once the MMAs have completed
        if ((++chained >= d2_group || tile + (int)gridDim.x >= num_tiles) && !(exp_bits & 8)) {
            chained = 0;
            float v[32];
            tc::tmem_ld32(d2 + lane_sel, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (lane < 16) racc[i] += v[i];
            tc::tmem_ld32(d2 + lane_sel + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float o = __shfl_sync(FULL, v[i], lane & 15);
                if (lane >= 16) racc[i] += o;
            }
        }
        tc::tc_fence_before_sync();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = rl0 + i;
            if (rl < rows && !(exp_bits & 16)) {
                float4 v = ld4(stg + rl * STG_LD + 4 * c4);
                const uint32_t b = mbits >> (4 * i);
                st4(du_prev + (size_t)(row0 + rl) * 64 + 4 * c4,
                    make_float4(b & 1u ? v.x : 0.f, b & 2u ? v.y : 0.f, b & 4u ? v.z : 0.f, b & 8u ? v.w : 0.f));
            }
        }
        __syncthreads();                        // staging is overwritten by the next gather
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
// The same round with the MMAs off the critical path (r02d): 8-warp CTAs, two per SM, TWO g stages.
// In k_embed_round_bwd_tc16 the 48 MMAs of a tile (~1.7 k cycles + commit / wait latency) sit in every CTA's serial
// chain (ablation: 0.27 ms of 1.24 ms).  Here the MMAs of tile k run while the CTA gathers tile k+1 into the other g
// stage; the own mu^{t-1} rows of tile k+1 wait in 16 registers per thread and are split into the (single) m tile once
// the MMAs of tile k have completed.  Per tile: gather(k+1) | wait MMA(k) | D1(k) -> staging (the dead g stage) |
// [D2 read-out] | sync | du^{t-1}(k) stores, m(k+1) stores | sync | issue MMA(k+1).  All 16 warps of the SM gather
// (4 rows per half-warp: 16 neighbour loads + 4 own-row loads per lane).  96 KB of shared memory and 128 TMEM columns
// per CTA.  Warp w reads TMEM lane quarter w & 3, column half w >> 2.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS2, 2)
k_embed_round_bwd_tc16p(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ du_t,
                        const float* __restrict__ mu_prev, float* __restrict__ du_prev, float* __restrict__ partial,
                        int partial_stride, int accumulate, int num_tiles, int d2_group) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // theta2^T: tile(row k, col n) = theta2[n][k]
    uint8_t* ga = w1 + TILE16X3;                 // gathered du^t tile, stage 0 / 1; a dead stage is the staging buffer
    uint8_t* m1 = ga + 2 * TILE16X3;             // own mu^{t-1} tile
    TcShared* sh = reinterpret_cast<TcShared*>(m1 + TILE16X3);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, hh = warp >> 2;
    const unsigned FULL = 0xffffffffu;

    if (warp == 0) tc::tmem_alloc<128>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 2); tc::fence_barrier_init(); }
    load_weight_tiles16(w1, theta2_w, true, TC_THREADS2);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d1 = sh->tmem_base, d2 = sh->tmem_base + 64;
    const uint32_t s_g = tc::smem_u32(ga), s_w = tc::smem_u32(w1), s_m = tc::smem_u32(m1);
    const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
    uint32_t phase = 0;
    const int c4 = lane & 15, rl0 = 4 * (2 * warp + (lane >> 4));
    float racc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) racc[i] = 0.f;
    int chained = 0;                             // tiles accumulated in D2 since it was last read out
    uint32_t mbits = 0;                          // [mu^{t-1} > 0] of this lane's 4 rows x 4 columns, tile k
    GatherMeta<4> meta;
#ifdef S2V_BWD16_EXPERIMENTS
    const int exp_bits = d2_group >> 8;          // 1: no reduce MMAs, 2: no MMAs at all, 4: no m-tile stores, 8: no D2 read-out, 16: no du stores,
    d2_group &= 255;                             // 32: no g-tile stores, 64: no staging
#else
    constexpr int exp_bits = 0;
#endif

    auto own_rows = [&](int tile, float4 (&mv)[4]) {
        const int row0 = tile * TCR, rows = min(TCR, g.num_nodes - row0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            mv[i] = rl0 + i < rows ? ldg4(mu_prev + (size_t)(row0 + rl0 + i) * 64 + 4 * c4) : f4zero();
    };
    auto store_m = [&](const float4 (&mv)[4]) {
        mbits = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!(exp_bits & 4)) store_split3(m1, rl0 + i, c4, mv[i]);
            mbits |= (mv[i].x > 0.f ? 1u : 0u) << (4 * i) | (mv[i].y > 0.f ? 2u : 0u) << (4 * i) |
                     (mv[i].z > 0.f ? 4u : 0u) << (4 * i) | (mv[i].w > 0.f ? 8u : 0u) << (4 * i);
        }
    };
    auto prefetch = [&](int tile) {
        if (tile < num_tiles)
            gatherN_prefetch<4>(meta, g, g.rowptr_t, g.col_t, tile * TCR, min(TCR, g.num_nodes - tile * TCR), rl0);
    };
    auto issue = [&](int stage) {
        if (warp == 0) {                         // two issuing warps (one per accumulator: the order inside each stays fixed)
            tc::tc_fence_after_sync();
            if (!(exp_bits & 2)) issue_transform16(d1, s_g + stage * TILE16X3, s_w);
            tc::mma_commit(&sh->bar);
        } else if (warp == 4) {
            tc::tc_fence_after_sync();
            if (!(exp_bits & 3)) issue_reduce16(d2, s_g + stage * TILE16X3, s_m, chained != 0);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
    };

    if (blockIdx.x < num_tiles) {                // tile 0 of this CTA
        float4 mv[4];
        prefetch(blockIdx.x);
        own_rows(blockIdx.x, mv);
        gatherN_consume<4, 4>(meta, g.col_t, du_t, rl0, [&](int rl, const float4& acc) { if (!(exp_bits & 32)) store_split3(ga, rl, c4, acc); else if (acc.x == 123.f) mbits = 7; });
        store_m(mv);
        prefetch(blockIdx.x + gridDim.x);
        tc::fence_proxy_async_smem();
        __syncthreads();
        issue(0);
    }
    int s = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, s ^= 1) {
        const int row0 = tile * TCR;
        const int rows = min(TCR, g.num_nodes - row0);
        const int nt = tile + gridDim.x;
        const bool more = nt < num_tiles;        // CTA-uniform
        float4 mv[4];
        if (more) {                              // gather tile k+1 into the other stage while the tensor core works on tile k
            uint8_t* gn = ga + (s ^ 1) * TILE16X3;
            own_rows(nt, mv);
            gatherN_consume<4, 4>(meta, g.col_t, du_t, rl0, [&](int rl, const float4& acc) { if (!(exp_bits & 32)) store_split3(gn, rl, c4, acc); else if (acc.x == 123.f) mbits = 7; });
            prefetch(nt + gridDim.x);            // (extents a tile ahead of the ids, no dependent stall: 1.204 -> 1.247 ms, not kept)
        }
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        float* stg = reinterpret_cast<float*>(ga + s * TILE16X3);      // the g pieces of tile k are dead
        if (!(exp_bits & 64)) {
            float v[32];
            tc::tmem_ld32(d1 + lane_sel + 32 * hh, v);
            if (lane < 16) {
                float* dst = stg + (16 * q + lane) * STG_LD + 32 * hh;
#pragma unroll
                for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
            }
        }
        if ((++chained >= d2_group || !more) && !(exp_bits & 8)) {    // D2 -> registers: lanes 0-15 own the TMEM lanes M = 64 uses and keep 16 columns,
            chained = 0;                         // lanes 16-31 take the other 16 by shuffle
            float v[32];
            tc::tmem_ld32(d2 + lane_sel + 32 * hh, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float o = __shfl_sync(FULL, v[16 + i], lane & 15);
                racc[i] += lane < 16 ? v[i] : o;
            }
        }
        tc::tc_fence_before_sync();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = rl0 + i;
            if (rl < rows && !(exp_bits & 16)) {
                float4 v = ld4(stg + rl * STG_LD + 4 * c4);
                const uint32_t b = mbits >> (4 * i);
                st4(du_prev + (size_t)(row0 + rl) * 64 + 4 * c4,
                    make_float4(b & 1u ? v.x : 0.f, b & 2u ? v.y : 0.f, b & 4u ? v.z : 0.f, b & 8u ? v.w : 0.f));
            }
        }
        if (more) {
            store_m(mv);                         // the MMAs reading the m tile of tile k have completed
            tc::fence_proxy_async_smem();
        }
        __syncthreads();                         // staging read out, m tile complete
        if (more) issue(s ^ 1);
    }
    {   // thread holds dtheta2[n = 16 q + lane % 16][32 hh + 16 (lane / 16) .. + 16)
        float* my = partial + (size_t)blockIdx.x * partial_stride + (16 * q + (lane & 15)) * 64 + 32 * hh + 16 * (lane >> 4);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float4 x4 = make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]);
            if (accumulate) x4 = f4add(ld4(my + i), x4);
            st4(my + i, x4);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(sh->tmem_base);
}

#include "s2v_h2.cuh"
#include "s2v_ring.cuh"

// self-test: mode 6: out[rows x 64] = A * W;  mode 7: out[64 x 64] = A^T * B over the rows (per-tile accumulators
// summed on the CUDA cores)
__global__ void __launch_bounds__(TC_THREADS)
k_tc_selftest16(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int rows) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                          // 3 tiles each
    uint8_t* a1 = w1 + TILE16X3;
    uint8_t* m1 = a1 + TILE16X3;
    TcShared* sh = reinterpret_cast<TcShared*>(m1 + TILE16X3);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (warp == 0) tc::tmem_alloc<64>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    if (mode == 6) load_weight_tiles16(w1, B, true, TC_THREADS);      // tile(row n, col k) = W[k][n]
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d = sh->tmem_base;
    uint32_t phase = 0;
    float racc[64];
    for (int i = 0; i < 64; ++i) racc[i] = 0.f;
    const int num_tiles = (rows + TCR - 1) / TCR;
    for (int tile = 0; tile < num_tiles; ++tile) {
        const int row0 = tile * TCR, nr = min(TCR, rows - row0);
        for (int e = t; e < TCR * 16; e += TC_THREADS) {
            int rl = e >> 4, c4 = e & 15;
            store_split3(a1, rl, c4, rl < nr ? ldg4(A + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero());
            if (mode == 7) store_split3(m1, rl, c4, rl < nr ? ldg4(B + (size_t)(row0 + rl) * 64 + 4 * c4) : f4zero());
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
            if (mode == 6) issue_transform16(d, tc::smem_u32(a1), tc::smem_u32(w1));
            else issue_reduce16(d, tc::smem_u32(a1), tc::smem_u32(m1), false);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        float v[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            tc::tmem_ld32(d + ((uint32_t)(32 * warp) << 16) + 32 * half, v);
            if (mode == 6) {
                if (lane < 16 && 16 * warp + lane < nr)
                    for (int i = 0; i < 32; ++i) out[(size_t)(row0 + 16 * warp + lane) * 64 + 32 * half + i] = v[i];
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) racc[32 * half + i] += v[i];
            }
        }
        tc::tc_fence_before_sync();
        __syncthreads();
    }
    if (mode == 7 && lane < 16)
        for (int i = 0; i < 64; ++i) out[(16 * warp + lane) * 64 + i] = racc[i];
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(sh->tmem_base);
}

#include "s2v_ws.cuh"

// Tensor-pipe timing probe (one CTA): cycles per MMA sequence of the backward round, issued back to back by one
// thread and waited for with one commit: out[0] = reduce (16 SS MMAs, MN-major, M = 128), out[1] = transform with A
// from TMEM (24 MMAs, M = 64), out[2] = transform SS K-major (24 MMAs, M = 64), out[3] = both backward sequences;
// out[4..7] = the bf16 sequences (see the loop).
__global__ void __launch_bounds__(TC_THREADS)
k_tc_rate(float* __restrict__ out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    TcShared* sh = reinterpret_cast<TcShared*>(smem + 6 * TILE_BYTES);
    const int t = threadIdx.x, warp = t >> 5;
    for (int e = t; e < 6 * (int)TILE_BYTES / 4; e += TC_THREADS) reinterpret_cast<float*>(smem)[e] = 1.0f;
    if (warp == 0) tc::tmem_alloc<256>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (warp == 0) {                             // warp-collective issue (s2v_tc.cuh)
        const uint32_t d1 = sh->tmem_base, d2 = d1 + 64, a_tm = d1 + 128;
        const uint32_t b_hi = tc::smem_u32(smem), b_lo = b_hi + TILE_BYTES, g_hi = b_lo + TILE_BYTES, g_lo = g_hi + TILE_BYTES,
                       m_hi = g_lo + TILE_BYTES, m_lo = m_hi + TILE_BYTES;
        uint32_t phase = 0;
        for (int mode = 0; mode < 8; ++mode) {
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (mode == 0 || mode == 3) issue_reduce(d2, g_hi, m_hi, m_lo);
                if (mode == 1 || mode == 3) issue_transform_ts(d1, a_tm, b_hi, b_lo);
                if (mode == 2) issue_transform(d1, g_hi, g_lo, b_hi, b_lo);
                if (mode == 4 || mode == 6) issue_transform16(d1, g_hi, b_hi);         // 24 bf16 MMAs, K-major, M = 64
                if (mode == 5 || mode == 6) issue_reduce16(d2, g_hi, m_hi, false);     // 24 bf16 MMAs, MN-major, M = 64
                if (mode == 7) {                                                       // 8 bf16 MMAs, MN-major, M = 128
                    constexpr uint32_t idesc = tc::idesc_bf16(128, 64, 1, 1);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tc::mma_bf16(d2, tc::desc_mn16(g_hi, ks & 3, TILE16), tc::desc_mn16(m_hi, ks & 3, TILE16), idesc, ks > 0);
                }
            }
            tc::mma_commit(&sh->bar);
            tc::mbar_wait(&sh->bar, phase);
            phase ^= 1;
            if (t == 0) out[mode] = (float)(clock64() - t0) / reps;
        }
        // cycles per single bf16 MMA (K = 16) by shape: out[8 + 6*major + 3*(M==128) + {0,1,2 for N = 64,128,256}]
        for (int major = 0; major < 2; ++major)
            for (int mi = 0; mi < 2; ++mi)
                for (int ni = 0; ni < 3; ++ni) {
                    const int M = mi ? 128 : 64, N = 64 << ni;
                    const uint32_t idesc = tc::idesc_bf16(M, N, major, major);
                    long long t0 = clock64();
                    for (int r = 0; r < reps; ++r)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            uint64_t da = major ? tc::desc_mn16(g_hi, ks, TILE16) : tc::desc_k16(g_hi, ks);
                            uint64_t db = major ? tc::desc_mn16(b_hi, ks, TILE16) : tc::desc_k16(b_hi, ks);
                            tc::mma_bf16(d1, da, db, idesc, true);
                        }
                    tc::mma_commit(&sh->bar);
                    tc::mbar_wait(&sh->bar, phase);
                    phase ^= 1;
                    if (t == 0) out[8 + 6 * major + 3 * mi + ni] = (float)(clock64() - t0) / (4 * reps);
                }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(sh->tmem_base);
}

constexpr int FWD_SMEM = 4 * TILE_BYTES + 64 + 1024;
constexpr int WS_SMEM = TILE16X3 + WS_STAGES * WS_STAGE_BYTES + 2 * WS_STG_FLOATS * 4 + (3 * TCR * WS_MAX_F + 16 * 64) * 4 + 256 + 256 + 1024;
constexpr int WSF_SMEM = TILE16X3 + WSF_STAGES * WS_STAGE_BYTES + 2 * WS_STG_FLOATS * 4 + 3 * TCR * WS_MAX_F * 4 + 256 + 128 +
                         WSF_SLOTS * TCR * 256 + 1024;
static_assert(WSF_SMEM <= 232448 && 2 * WSF_SLOTS * 8 <= 128, "last-stage shared memory");
static_assert(sizeof(WsShared) <= 256, "WsShared");
static inline int first_smem_bytes(int F) {
    const int FS = (F + 3) & ~3;
    return 2 * (int)TILE16X3 + (FS * 64 + TCR * FS + 6 * 64 + 128) * 4 + 64 + 1024;
}
constexpr int FIN_SMEM = 6 * TILE_BYTES + (S2V_MAX_F * 64 + TCR * S2V_MAX_F + 64 + 128) * 4 + 64 + 1024;
constexpr int BWD_SMEM = 6 * TILE_BYTES + 64 + 1024;

#include "s2v_rows.cuh"

// out[(o_row0 + r) * ldo + o_col0 + c] = sum over CTAs of partial[b][r * 64 + c], r < rows, c < cols: fixed order
__global__ void k_sum_cta_partials_win(const float* __restrict__ partial, int nblocks, int rows, int cols, float* __restrict__ out,
                                       int ldo, int o_row0, int o_col0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * 64) return;
    const int r = e >> 6, c = e & 63;
    if (c >= cols) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * (rows * 64) + e];
    out[(size_t)(o_row0 + r) * ldo + o_col0 + c] = s;
}

// out[c] = sum over CTAs of partial[b][c], fixed order (one thread per element, coalesced across threads)
__global__ void k_sum_cta_partials(const float* __restrict__ partial, int nblocks, int width, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= width) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * width + c];
    out[c] = s;
}

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

template <int NS>
int ring_forward_launch_ns(int slots, int cap, const s2v_graph_t& g, const float* theta2_w, const float* base,
                           const float* mu_prev, float* mu_next, int tiles, cudaStream_t st) {
    int dev = 0, sms = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    S2V_RETURN_IF(cudaFuncGetAttributes(&fa, k_embed_round_ring<NS>));
    const int fixed = 1024 + (1 + NS) * (int)TILE16X3 + TCR * STG_LD * 4 + 2 * TCR * 256;
    const int max_slots = (smem_max - (int)fa.sharedSizeBytes - fixed) / 256;
    if (slots <= 0 || slots > max_slots) slots = max_slots;
    if (cap < 32) cap = 32;
    if (cap > RG_CAP_MAX) cap = RG_CAP_MAX;
    if (cap > slots) return S2V_ERR_UNSUPPORTED;
    const int smem = fixed + slots * 256;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_ring<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = sms < tiles ? sms : tiles;
    k_embed_round_ring<NS><<<grid, RG_THREADS, smem, st>>>(g, theta2_w, base, mu_prev, mu_next, tiles, slots, cap);
    return (int)cudaGetLastError();
}
int ring_forward_launch(int ns, int slots, int cap, const s2v_graph_t& g, const float* theta2_w, const float* base,
                        const float* mu_prev, float* mu_next, int tiles, cudaStream_t st) {
    return ns == 2 ? ring_forward_launch_ns<2>(slots, cap, g, theta2_w, base, mu_prev, mu_next, tiles, st)
                   : ring_forward_launch_ns<1>(slots, cap, g, theta2_w, base, mu_prev, mu_next, tiles, st);
}

}  // namespace

// Launchers used by s2v_embed.cu (p = 64 only).  grid_out: number of CTAs (= dtheta2 partials).
int s2v_tc_round_forward(const s2v_graph_t& g, const float* theta2_w, const float* base, const float* mu_prev,
                         float* mu_next, cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    static const bool tf32 = getenv("S2V_B200_FWD_TF32") != nullptr;      // the 3xTF32 forward round, kept for A/B runs
    if (tf32) {
        S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
        int grid = tc_grid((const void*)k_embed_round_tc, FWD_SMEM, 64, tiles);
        k_embed_round_tc<<<grid, TC_THREADS, FWD_SMEM, st>>>(g, theta2_w, base, mu_prev, mu_next, tiles);
        return (int)cudaGetLastError();
    }
    // asynchronous-gather variant (s2v_ring.cuh): S2V_B200_FWD_RING=<stages 1|2>[,<ring slots>[,<edges per chunk>]]
    const char* ring_env = getenv("S2V_B200_FWD_RING");
    if (ring_env && (g.period == 0 || g.period >= TCR) && tiles > 0) {
        int ns = 1, slots = 0, cap = 128;
        sscanf(ring_env, "%d,%d,%d", &ns, &slots, &cap);
        return ring_forward_launch(ns == 2 ? 2 : 1, slots, cap, g, theta2_w, base, mu_prev, mu_next, tiles, st);
    }
    const int h2_ctas = [] {                   // S2V_B200_FWD_ROUND=tc16 | h2[,<CTAs per SM: 4 5 6>]  (read per call: tests switch it)
        const char* e = getenv("S2V_B200_FWD_ROUND");
        if (!e) return S2V_FWD_H2_DEFAULT;
        if (strncmp(e, "h2", 2) != 0) return 0;
        int c = S2V_FWD_H2_DEFAULT ? S2V_FWD_H2_DEFAULT : 5;
        if (e[2] == ',') c = atoi(e + 3);
        return c < 4 ? 4 : (c > 6 ? 6 : c);
    }();
    if (h2_ctas) {
        constexpr int FWDH2_SMEM = 2 * (int)TILE16X2 + 128 + 1024;
        auto launch = [&](auto kern) -> int {
            S2V_RETURN_IF(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FWDH2_SMEM));
            int grid = tc_grid((const void*)kern, FWDH2_SMEM, 64, tiles);
            kern<<<grid, TC_THREADS, FWDH2_SMEM, st>>>(g, theta2_w, base, mu_prev, mu_next, tiles);
            return (int)cudaGetLastError();
        };
        if (h2_ctas == 4) return launch(k_embed_round_h2<4>);
        if (h2_ctas == 5) return launch(k_embed_round_h2<5>);
        return launch(k_embed_round_h2<6>);
    }
    constexpr int FWD16_SMEM = 2 * (int)TILE16X3 + 64 + 1024;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_tc16, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD16_SMEM));
    int grid = tc_grid((const void*)k_embed_round_tc16, FWD16_SMEM, 64, tiles);
    k_embed_round_tc16<<<grid, TC_THREADS, FWD16_SMEM, st>>>(g, theta2_w, base, mu_prev, mu_next, tiles);
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

int s2v_tc_round_backward_ws_grid(int num_nodes) {
    const int tiles = (num_nodes + TCR - 1) / TCR;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms < tiles ? sms : (tiles < 1 ? 1 : tiles);
}

int s2v_tc_round_backward_ws(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                             float* du_prev, float* partial, int partial_stride, int accumulate, int grid, cudaStream_t st) {
    WsArgs a{};
    a.w = theta2_w; a.src = du_t; a.own = mu_prev; a.dst = du_prev;
    a.partial = partial; a.partial_stride = partial_stride; a.accumulate = accumulate;
    a.num_tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_bwd_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
#ifdef S2V_WS_EXPERIMENTS
    if (const char* e = getenv("S2V_WS_EXP")) a.accumulate |= atoi(e) << 8;
#endif
    k_embed_bwd_ws<false><<<grid, WS_THREADS, WS_SMEM, st>>>(g, a);
    return (int)cudaGetLastError();
}

#ifndef S2V_BWD16_DEFAULT_GROUP
#define S2V_BWD16_DEFAULT_GROUP 1      /* tiles chained in D2 between two read-outs */
#endif
#ifndef S2V_BWD16_H2
#define S2V_BWD16_H2 1
#endif
#ifndef S2V_BWD16_PIPELINED
#define S2V_BWD16_PIPELINED 1
#endif
#ifndef S2V_BWD16_DEFAULT
#define S2V_BWD16_DEFAULT 1            /* auto: 0 = warp-specialised backward round, 1 = the S2V_B200_BWD16_KERNEL family (default: k_embed_round_bwd_h2) */
#endif
// Which backward-round kernel the tensor path runs: flags S2V_PATH_TENSOR_WS -> warp-specialised, S2V_PATH_TENSOR_16 ->
// k_embed_round_bwd_tc16, auto -> S2V_B200_BWD_ROUND=ws|tc16 (default: S2V_BWD16_DEFAULT).  Returns 0 for the
// warp-specialised kernel, else the number of tiles k_embed_round_bwd_tc16 chains in D2 (S2V_B200_BWD16_GROUP, 1..8).
static int bwd16_group(int flags) {
    static const int grp = [] {
        const char* e = getenv("S2V_B200_BWD16_GROUP");
        int v = e ? atoi(e) : S2V_BWD16_DEFAULT_GROUP;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    static const bool dflt16 = [] {
        const char* e = getenv("S2V_B200_BWD_ROUND");
        return e ? strncmp(e, "tc16", 4) == 0 : (S2V_BWD16_DEFAULT != 0);
    }();
    if (flags == S2V_PATH_TENSOR_WS) return 0;
    if (flags == S2V_PATH_TENSOR_16) return grp;
    return dflt16 ? grp : 0;
}
constexpr int BWD16_SMEM = 3 * (int)TILE16X3 + 64 + 1024;
constexpr int BWD16P_SMEM = 4 * (int)TILE16X3 + 64 + 1024;
constexpr int BWDH2_SMEM = 3 * (int)TILE16X2 + 128 + 1024;
static int bwdh2_ctas() { static const int c = [] { const char* e = getenv("S2V_B200_H2_CTAS"); return e && atoi(e) == 4 ? 4 : 3; }(); return c; }
static bool bwd16_h2() {                     // S2V_B200_BWD16_KERNEL=h: two fp16 pieces per operand, per-tile scales (k_embed_round_bwd_h2)
    static const bool h = [] { const char* e = getenv("S2V_B200_BWD16_KERNEL"); return e ? e[0] == 'h' : (S2V_BWD16_H2 != 0); }();
    return h;
}
static bool bwd16_pipelined() {              // S2V_B200_BWD16_KERNEL=s: 4-warp CTAs, MMAs in the serial chain; p: 8-warp CTAs, two g stages
    static const bool p = [] { const char* e = getenv("S2V_B200_BWD16_KERNEL"); return e ? e[0] == 'p' : (S2V_BWD16_PIPELINED != 0); }();
    return p;
}

int s2v_tc_round_backward_sel_grid(int num_nodes, int flags) {
    if (bwd16_group(flags) == 0) return s2v_tc_round_backward_ws_grid(num_nodes);
    const int tiles = (num_nodes + TCR - 1) / TCR;
    if (bwd16_h2()) {
        if (bwdh2_ctas() == 4) {
            cudaFuncSetAttribute(k_embed_round_bwd_h2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWDH2_SMEM);
            return tc_grid((const void*)k_embed_round_bwd_h2<4>, BWDH2_SMEM, 128, tiles);
        }
        cudaFuncSetAttribute(k_embed_round_bwd_h2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWDH2_SMEM);
        return tc_grid((const void*)k_embed_round_bwd_h2<3>, BWDH2_SMEM, 128, tiles);
    }
    if (bwd16_pipelined()) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = 2 * sms;                // two 8-warp CTAs per SM (97 KB, 128 registers, 128 TMEM columns each)
        return grid < tiles ? grid : (tiles < 1 ? 1 : tiles);
    }
    cudaFuncSetAttribute(k_embed_round_bwd_tc16, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD16_SMEM);
    return tc_grid((const void*)k_embed_round_bwd_tc16, BWD16_SMEM, 128, tiles);
}

int s2v_tc_round_backward_sel(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                              float* du_prev, float* partial, int partial_stride, int accumulate, int grid, int flags,
                              cudaStream_t st) {
    const int grp = bwd16_group(flags);
    if (grp == 0) return s2v_tc_round_backward_ws(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride, accumulate, grid, st);
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    if (bwd16_h2()) {
        if (bwdh2_ctas() == 4) {
            S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_bwd_h2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWDH2_SMEM));
            k_embed_round_bwd_h2<4><<<grid, TC_THREADS, BWDH2_SMEM, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride,
                                                                        accumulate, tiles);
            return (int)cudaGetLastError();
        }
        S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_bwd_h2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWDH2_SMEM));
        k_embed_round_bwd_h2<3><<<grid, TC_THREADS, BWDH2_SMEM, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride,
                                                                    accumulate, tiles);
        return (int)cudaGetLastError();
    }
    if (bwd16_pipelined()) {
        S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_bwd_tc16p, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD16P_SMEM));
        int grp_p = grp;
#ifdef S2V_BWD16_EXPERIMENTS
        if (const char* e = getenv("S2V_B200_BWD16_EXP")) grp_p |= atoi(e) << 8;
#endif
        k_embed_round_bwd_tc16p<<<grid, TC_THREADS2, BWD16P_SMEM, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride,
                                                                       accumulate, tiles, grp_p);
        return (int)cudaGetLastError();
    }
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_bwd_tc16, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD16_SMEM));
    int grp_arg = grp;
#ifdef S2V_BWD16_EXPERIMENTS
    if (const char* e = getenv("S2V_B200_BWD16_EXP")) grp_arg |= atoi(e) << 8;
#endif
    k_embed_round_bwd_tc16<<<grid, TC_THREADS, BWD16_SMEM, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride,
                                                                accumulate, tiles, grp_arg);
    return (int)cudaGetLastError();
}

// Last backward stage, fp16x2 forward-shaped kernel (node_dim <= 8): chosen with the k_embed_round_bwd_h2 family
constexpr int FINH2_SMEM = 3 * (int)TILE16X2 + (FH2_FS * 64 + TCR * FH2_FS + 64 + 128) * 4 + 128 + 1024;
bool s2v_tc_final_backward_h2_selected(int flags, int node_dim) {
    static const bool off = getenv("S2V_B200_FINAL_WS") != nullptr;      // A/B: keep the warp-specialised last stage
    return !off && node_dim <= FH2_FS && bwd16_group(flags) != 0 && bwd16_h2();
}
int s2v_tc_final_backward_h2_grid(int num_nodes) {
    const int tiles = (num_nodes + TCR - 1) / TCR;
    cudaFuncSetAttribute(k_embed_final_bwd_h2, cudaFuncAttributeMaxDynamicSharedMemorySize, FINH2_SMEM);
    return tc_grid((const void*)k_embed_final_bwd_h2, FINH2_SMEM, 128, tiles);
}
int s2v_tc_final_backward_h2(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, const float* dus,
                             const float* du_last, float* partial, int partial_stride, int grid, cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_final_bwd_h2, cudaFuncAttributeMaxDynamicSharedMemorySize, FINH2_SMEM));
    k_embed_final_bwd_h2<<<grid, TC_THREADS, FINH2_SMEM, st>>>(g, prm, x, dus, du_last, prm.rounds, partial, partial_stride, tiles);
    return (int)cudaGetLastError();
}

// Last backward stage on the same pipeline (node_dim <= WS_MAX_F); partial rows [0, grid) receive TH1B .. TH1A.
bool s2v_tc_final_backward_ws_ok(int node_dim) { return node_dim <= WS_MAX_F; }
int s2v_tc_final_backward_ws(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, const float* dus,
                             const float* du_last, float* partial, int partial_stride, int grid, cudaStream_t st) {
    WsArgs a{};
    a.w = prm.theta1b_w; a.src = dus; a.own = du_last; a.dst = nullptr;
    a.x = x; a.th1a_w = prm.theta1a_w; a.th1a_b = prm.theta1a_b; a.F = prm.node_dim; a.T = prm.rounds;
    a.partial = partial; a.partial_stride = partial_stride; a.accumulate = 0;
    a.num_tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_bwd_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WSF_SMEM));
#ifdef S2V_WS_EXPERIMENTS
    if (const char* e = getenv("S2V_WS_EXP_FINAL")) a.accumulate |= atoi(e) << 8;
#endif
    k_embed_bwd_ws<true><<<grid, WS_THREADS, WSF_SMEM, st>>>(g, a);
    return (int)cudaGetLastError();
}

int s2v_tc_round_backward(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                          float* du_prev, float* partial, int partial_stride, int accumulate, int grid, cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_round_bwd_tc256, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    k_embed_round_bwd_tc256<<<grid, TC_THREADS2, BWD_SMEM, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial, partial_stride,
                                                                accumulate, tiles);
    return (int)cudaGetLastError();
}

#ifdef S2V_WSF_PROF
extern "C" int s2v_debug_wsf_prof(long long* host_out) { return (int)cudaMemcpyFromSymbol(host_out, g_wsf_prof, sizeof(long long) * 16); }
#endif
#ifdef S2V_RING_PROF
extern "C" int s2v_debug_ring_prof(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_ring_prof, sizeof(long long) * 32);
}
#endif
#ifdef S2V_TC_TIMELINE
extern "C" int s2v_debug_timeline(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 16 * 8);
}
#endif

int s2v_tc_embed_first(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, float* base, float* mu1,
                       cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    // (The same stage on two fp16 pieces -- k_embed_first_h2, r02d -- measured 0.599 ms at five CTAs per SM against 0.592 ms:
    // this stage is bound by its 2 GB of stores and its per-tile barrier chain, not by the MMAs or the tile footprint.)
    const int FIRST_SMEM = first_smem_bytes(prm.node_dim);
    S2V_RETURN_IF(cudaFuncSetAttribute(k_embed_first_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FIRST_SMEM));
    int grid = tc_grid((const void*)k_embed_first_tc, FIRST_SMEM, 64, tiles);
    k_embed_first_tc<<<grid, TC_THREADS, FIRST_SMEM, st>>>(g, prm, x, base, mu1, tiles);
    return (int)cudaGetLastError();
}

int s2v_tc_head_forward(const s2v_graph_t& g, const s2v_head_params_t& prm, const float* mu, const float* gstate, float* q,
                        cudaStream_t st) {
    const int tiles = (g.num_nodes + TCR - 1) / TCR;
    const int h2_ctas = [] {                   // S2V_B200_HEAD=tc16 | h2[,<CTAs per SM: 4 5 6>]
        const char* e = getenv("S2V_B200_HEAD");
        if (!e) return S2V_HEAD_H2_DEFAULT;
        if (strncmp(e, "h2", 2) != 0) return 0;
        int c = S2V_HEAD_H2_DEFAULT ? S2V_HEAD_H2_DEFAULT : 5;
        if (e[2] == ',') c = atoi(e + 3);
        return c < 4 ? 4 : (c > 6 ? 6 : c);
    }();
    if (h2_ctas) {
        constexpr int HEADH2_SMEM = 2 * (int)TILE16X2 + 3 * 64 * 4 + 128 + 1024;
        auto launch = [&](auto kern) -> int {
            S2V_RETURN_IF(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADH2_SMEM));
            int grid = tc_grid((const void*)kern, HEADH2_SMEM, 64, tiles);
            kern<<<grid, TC_THREADS, HEADH2_SMEM, st>>>(g, prm, mu, gstate, q, tiles);
            return (int)cudaGetLastError();
        };
        if (h2_ctas == 4) return launch(k_head_h2<4>);
        if (h2_ctas == 5) return launch(k_head_h2<5>);
        return launch(k_head_h2<6>);
    }
    constexpr int HEAD_SMEM = 2 * (int)TILE16X3 + 3 * 64 * 4 + 64 + 1024;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_head_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM));
    int grid = tc_grid((const void*)k_head_tc, HEAD_SMEM, 64, tiles);
    k_head_tc<<<grid, TC_THREADS, HEAD_SMEM, st>>>(g, prm, mu, gstate, q, tiles);
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
    if (mode == 5) {                             // timing probe: out[0..3] = cycles per MMA sequence, rows = repetitions
        if (!out || rows < 1) return S2V_ERR_BAD_ARG;
        S2V_RETURN_IF(cudaFuncSetAttribute(k_tc_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
        k_tc_rate<<<1, TC_THREADS, BWD_SMEM, (cudaStream_t)stream>>>(out, (int)rows);
        return (int)cudaGetLastError();
    }
    if (mode == 6 || mode == 7) {                // bf16x3 building blocks
        if (!A || !B || !out || rows < 1) return S2V_ERR_BAD_ARG;
        S2V_RETURN_IF(cudaFuncSetAttribute(k_tc_selftest16, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
        k_tc_selftest16<<<1, TC_THREADS, BWD_SMEM, (cudaStream_t)stream>>>(mode, A, B, out, (int)rows);
        return (int)cudaGetLastError();
    }
    if (!A || !B || !out || rows < 1 || mode < 0 || mode > 3) return S2V_ERR_BAD_ARG;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    k_tc_selftest<<<1, TC_THREADS, BWD_SMEM, (cudaStream_t)stream>>>(mode, A, B, out, (int)rows);
    return (int)cudaGetLastError();
}


// ------------------------------------------------------------------------------------
// C ABI: node-wise linear layers of the GATv2 branch on the tensor cores (include/s2v_b200.h)
// ------------------------------------------------------------------------------------
template <int KB, int NB, bool MASK>
static int launch_rows_gemm(const RowsGemmArgs& a, cudaStream_t st) {
    constexpr int SMEM = (NB * KB + KB) * (int)TILE16X3 + 64 + 1024;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_rows_gemm16<KB, NB, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int grid = tc_grid((const void*)k_rows_gemm16<KB, NB, MASK>, SMEM, 64 * NB, a.num_tiles);
    k_rows_gemm16<KB, NB, MASK><<<grid, TC_THREADS, SMEM, st>>>(a);
    return (int)cudaGetLastError();
}

// out[n, m] = A[n, k] W^T (W [m, k], transpose_w = 0) or A[n, k] W (W [k, m], transpose_w = 1), optionally * (mask > 0).
// k <= 256 (any value: ragged K is zero padded), m a multiple of 64, <= 256.  Planned as passes of at most 128 output
// columns x 128 contraction columns (144 KB of bf16x3 tiles per CTA); later K passes accumulate into out.
extern "C" int s2v_rows_linear(const float* A, int32_t lda, int32_t k, const float* W, int32_t ldw, int32_t m,
                               int32_t transpose_w, const float* mask, int32_t ldm, float* out, int32_t ldo, int64_t n,
                               void* stream) {
    if (!A || !W || !out || n < 0 || k < 1 || k > 256 || m < 64 || m > 256 || (m & 63) || lda < k || ldo < m || (ldo & 3))
        return S2V_ERR_BAD_ARG;
    if (mask && (ldm < m || (ldm & 3))) return S2V_ERR_BAD_ARG;
    if (((uintptr_t)out & 15) || (mask && ((uintptr_t)mask & 15))) return S2V_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (n > 0x7fffffff - 64) return S2V_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    RowsGemmArgs a{};
    a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.transpose_w = transpose_w; a.M = mask; a.ldm = ldm; a.out = out; a.ldo = ldo;
    a.n = (int)n; a.num_tiles = (int)((n + TCR - 1) / TCR);
    for (int oc = 0; oc < m; oc += 128) {
        const int nb = (m - oc) >= 128 ? 2 : 1;
        for (int kc = 0; kc < k; kc += 128) {
            const int kv = k - kc < 128 ? k - kc : 128, kb = kv > 64 ? 2 : 1;
            const bool last_k = kc + 128 >= k;
            a.a_col0 = kc; a.k_valid = kv; a.o_col0 = oc; a.m_col0 = oc; a.accumulate = kc > 0; a.w_n_valid = 64 * nb;
            if (transpose_w) { a.w_row0 = kc; a.w_col0 = oc; } else { a.w_row0 = oc; a.w_col0 = kc; }
            const bool msk = mask != nullptr && last_k;
            int rc;
            if (kb == 1 && nb == 1) rc = msk ? launch_rows_gemm<1, 1, true>(a, st) : launch_rows_gemm<1, 1, false>(a, st);
            else if (kb == 1 && nb == 2) rc = msk ? launch_rows_gemm<1, 2, true>(a, st) : launch_rows_gemm<1, 2, false>(a, st);
            else if (kb == 2 && nb == 1) rc = msk ? launch_rows_gemm<2, 1, true>(a, st) : launch_rows_gemm<2, 1, false>(a, st);
            else rc = msk ? launch_rows_gemm<2, 2, true>(a, st) : launch_rows_gemm<2, 2, false>(a, st);
            S2V_RETURN_IF(rc);
        }
    }
    return 0;
}

extern "C" int s2v_rows_gemm64(const float* A, int32_t lda, int32_t k_blocks, const float* W, int32_t n_blocks,
                               int32_t transpose_w, const float* mask, int32_t ldm, float* out, int32_t ldo, int64_t n,
                               void* stream) {
    if (k_blocks < 1 || n_blocks < 1) return S2V_ERR_BAD_ARG;
    if (k_blocks > 4 || n_blocks > 4) return S2V_ERR_UNSUPPORTED;
    return s2v_rows_linear(A, lda, 64 * k_blocks, W, transpose_w ? 64 * n_blocks : 64 * k_blocks, 64 * n_blocks, transpose_w, mask,
                           ldm, out, ldo, n, stream);
}

extern "C" int64_t s2v_rows_reduce64_workspace_bytes(int32_t a_blocks) {
    (void)a_blocks;
    return (int64_t)S2V_MAX_GRID * 2 * 64 * 64 * (int64_t)sizeof(float);
}

template <int AB>
static int launch_rows_reduce(const RowsReduceArgs& a, float* out, int ldo, int o_row0, int o_col0, cudaStream_t st) {
    constexpr int SMEM = (AB + 1) * (int)TILE16X3 + 64 + 1024;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_rows_reduce16<AB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int grid = tc_grid((const void*)k_rows_reduce16<AB>, SMEM, 64 * AB, a.num_tiles);
    k_rows_reduce16<AB><<<grid, TC_THREADS, SMEM, st>>>(a);
    S2V_RETURN_IF(cudaGetLastError());
    const int elems = AB * 64 * 64;
    k_sum_cta_partials_win<<<(elems + 255) / 256, 256, 0, st>>>(a.partial, grid, AB * 64, a.b_valid, out, ldo, o_row0, o_col0);
    return (int)cudaGetLastError();
}

// out[m, k] = A[n, m]^T B[n, k] (weight gradient dW = dY^T X): m a multiple of 64 <= 256, k <= 128 (ragged: F = 7).
// Per-CTA partials in workspace (s2v_rows_reduce64_workspace_bytes), summed in fixed order.
extern "C" int s2v_rows_outer(const float* A, int32_t lda, int32_t m, const float* B, int32_t ldb, int32_t k, float* out,
                              int32_t ldo, void* workspace, int64_t workspace_bytes, int64_t n, void* stream) {
    if (!A || !B || !out || !workspace || n < 0 || m < 64 || m > 256 || (m & 63) || k < 1 || k > 128 || lda < m || ldb < k ||
        ldo < k || (lda & 3) || ((uintptr_t)A & 15))
        return S2V_ERR_BAD_ARG;
    if (workspace_bytes < s2v_rows_reduce64_workspace_bytes(2)) return S2V_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        for (int r = 0; r < m; ++r) S2V_RETURN_IF(cudaMemsetAsync(out + (size_t)r * ldo, 0, sizeof(float) * k, st));
        return 0;
    }
    if (n > 0x7fffffff - 64) return S2V_ERR_BAD_ARG;
    RowsReduceArgs a{};
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.partial = (float*)workspace; a.n = (int)n; a.num_tiles = (int)((n + TCR - 1) / TCR);
    for (int ac = 0; ac < m; ac += 128) {
        const int ab = (m - ac) >= 128 ? 2 : 1;
        for (int bc = 0; bc < k; bc += 64) {
            a.a_col0 = ac; a.b_col0 = bc; a.b_valid = k - bc < 64 ? k - bc : 64;
            S2V_RETURN_IF(ab == 2 ? launch_rows_reduce<2>(a, out, ldo, ac, bc, st) : launch_rows_reduce<1>(a, out, ldo, ac, bc, st));
        }
    }
    return 0;
}

extern "C" int s2v_rows_reduce64(const float* A, int32_t lda, int32_t a_blocks, const float* B, int32_t ldb, float* out,
                                 void* workspace, int64_t workspace_bytes, int64_t n, void* stream) {
    if (a_blocks < 1 || a_blocks > 4) return S2V_ERR_UNSUPPORTED;
    return s2v_rows_outer(A, lda, 64 * a_blocks, B, ldb, 64, out, 64, workspace, workspace_bytes, n, stream);
}
