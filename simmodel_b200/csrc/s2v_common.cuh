// Shared device helpers for the s2v kernels (sm_100a).
//
// Tile model used by every node-wise kernel: a CTA of S2V_THREADS threads owns a
// tile of S2V_TM consecutive node rows.  Row tiles live in shared memory row-major
// with a leading dimension of P+4 floats (16-byte aligned rows, bank-conflict-free
// for the access patterns below); weight matrices live in shared memory as
// [k][n] so that one float4 read yields four output columns.
//
// Thread mapping (same for the row transform C = A * W and for the reduction
// C = A^T * B over the rows of a tile):  tx = t % TXP selects four consecutive
// output columns, ty = t / TXP selects the row group; thread rows are ty + RG*i.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "s2v_b200.h"

#define S2V_THREADS 128
#define S2V_TM 64
#define S2V_MAX_GRID 1024
#define S2V_MAX_F 32

// TM = tile height.  64 everywhere except the small-problem forward stages (16: four times the CTAs, a quarter of the
// work per stage -- the stage latency, not the throughput, is what a 975-row problem pays).
template <int P, int TM = S2V_TM>
struct Cfg {
    static_assert(P % 4 == 0 && P >= 8 && P <= 64, "emb_dim must be a multiple of 4 in [8,64]");
    static constexpr int TXP = P / 4;                            // float4 column groups per row
    static constexpr int RG = S2V_THREADS / TXP;                 // row groups
    static constexpr int ACTIVE = TXP * RG;                      // threads that own outputs
    static constexpr int NR = (TM + RG - 1) / RG;                // tile rows per thread (transform)
    static constexpr int NRR = (P + RG - 1) / RG;                // output rows per thread (reduction, P x P)
    static constexpr int LDA = P + 4;                            // leading dimension of row tiles
    static constexpr int TILE_FLOATS = TM * LDA;
    static constexpr int W_FLOATS = P * P;
    // gather: LPR lanes (float4 each) cover one row, RPP rows per warp pass
    static constexpr int LPR = P / 4;
    static constexpr int RPP = 32 / LPR;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float relu(float v) { return v > 0.f ? v : 0.f; }

// Row loads of the gathers.  LD = 0: ld.global.nc (the source is read-only for the whole kernel: one launch per round);
// LD = 1: ld.global.cg (L2): the single-launch small-graph kernel reads rows other CTAs wrote earlier in the SAME launch.
template <int LD>
__device__ __forceinline__ float4 ld_row4(const float* p) {
    if constexpr (LD == 1) return __ldcg(reinterpret_cast<const float4*>(p));
    else return __ldg(reinterpret_cast<const float4*>(p));
}


// Row r of the batch -> (row in the CSR, offset to add to CSR column ids, graph id).
__device__ __forceinline__ void row_info(const s2v_graph_t& g, int r, int& li, int& off, int& gi) {
    if (g.period > 0) {
        gi = r / g.period;
        off = gi * g.period;
        li = r - off;
    } else {
        li = r;
        off = 0;
        gi = __ldg(g.gid + r);
    }
}
__device__ __forceinline__ int graph_npad(const s2v_graph_t& g, int gi) {
    return g.npad ? __ldg(g.npad + gi) : g.npad_scalar;
}
__device__ __forceinline__ void graph_range(const s2v_graph_t& g, int gi, int& lo, int& hi) {
    if (g.period > 0) { lo = gi * g.period; hi = lo + g.period; }
    else { lo = __ldg(g.ptr + gi); hi = __ldg(g.ptr + gi + 1); }
}

// a + sum_k w_row[k] * v[k], k ascending (the order of the plain loop), with the P/4 row loads issued together: a thread
// that walks a weight row with scalar loads pays one L2 round trip per element -- 64 in a row are ~10 us, the whole
// budget of a small-graph launch.
template <int P>
__device__ __forceinline__ float row_dot(const float* __restrict__ w_row, const float* v, float a) {
    float4 w[P / 4];
#pragma unroll
    for (int k4 = 0; k4 < P / 4; ++k4) w[k4] = ldg4(w_row + 4 * k4);
#pragma unroll
    for (int k4 = 0; k4 < P / 4; ++k4) {
        a = fmaf(w[k4].x, v[4 * k4], a); a = fmaf(w[k4].y, v[4 * k4 + 1], a);
        a = fmaf(w[k4].z, v[4 * k4 + 2], a); a = fmaf(w[k4].w, v[4 * k4 + 3], a);
    }
    return a;
}

// smem Wt[k*P + n] = W[n*P + k]   (y = x * W^T form)
template <int P>
__device__ __forceinline__ void load_w_transposed(float* Ws, const float* __restrict__ W) {
    // consecutive threads take consecutive rows n: the transposed shared-memory stores are conflict-free and the
    // P*P/4/128 vector loads of a thread are independent (one L2 round trip for the whole matrix)
    constexpr int V = P / 4;
#pragma unroll
    for (int e0 = 0; e0 < P * V; e0 += S2V_THREADS) {
        const int e = e0 + threadIdx.x;
        if (P * V % S2V_THREADS != 0 && e >= P * V) break;
        const int c = e / P, n = e - c * P;
        const float4 w = ldg4(W + n * P + 4 * c);
        Ws[(4 * c) * P + n] = w.x; Ws[(4 * c + 1) * P + n] = w.y;
        Ws[(4 * c + 2) * P + n] = w.z; Ws[(4 * c + 3) * P + n] = w.w;
    }
}
// smem Ws[n*P + k] = W[n*P + k]   (y = g * W form)
template <int P>
__device__ __forceinline__ void load_w_plain(float* Ws, const float* __restrict__ W) {
    for (int e = threadIdx.x * 4; e < P * P; e += S2V_THREADS * 4) st4(Ws + e, ldg4(W + e));
}

// Copy `rows` consecutive rows of a [*,P] global matrix into a row tile; rows beyond are zeroed.
template <int P, int LD = 0, int TM = S2V_TM>
__device__ __forceinline__ void load_tile(float* As, const float* src, int rows) {
    constexpr int LDA = Cfg<P>::LDA;
    constexpr int V = P / 4;
    for (int e = threadIdx.x; e < TM * V; e += S2V_THREADS) {
        int r = e / V, c = e - r * V;
        float4 v = r < rows ? ld_row4<LD>(src + (size_t)r * P + 4 * c) : f4zero();
        st4(As + r * LDA + 4 * c, v);
    }
}

// acc[i][0..3] += sum_k As[row_i][k] * Ws[k][4tx..4tx+3],  row_i = ty + RG*i
template <int P, int TM = S2V_TM>
__device__ __forceinline__ void tile_gemm(const float* As, const float* Ws, float (&acc)[Cfg<P, TM>::NR][4], int tx, int ty) {
    constexpr int LDA = Cfg<P>::LDA, RG = Cfg<P>::RG, NR = Cfg<P, TM>::NR;
#pragma unroll 2
    for (int k0 = 0; k0 < P; k0 += 4) {
        float4 w0 = ld4(Ws + (k0 + 0) * P + 4 * tx);
        float4 w1 = ld4(Ws + (k0 + 1) * P + 4 * tx);
        float4 w2 = ld4(Ws + (k0 + 2) * P + 4 * tx);
        float4 w3 = ld4(Ws + (k0 + 3) * P + 4 * tx);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            int r = ty + RG * i;
            if (NR * RG > TM && r >= TM) continue;
            float4 a = ld4(As + r * LDA + k0);
            acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
        }
    }
}

// racc[i][0..3] += sum_{r<rows} As[r][NRR*ty + i] * Bs[r][4tx..4tx+3]     (P x P output)
// The thread's output rows are consecutive so that its A values are 128-bit shared-memory loads.
template <int P>
__device__ __forceinline__ void tile_reduce_gemm(const float* As, const float* Bs, int rows,
                                                 float (&racc)[Cfg<P>::NRR][4], int tx, int ty) {
    constexpr int LDA = Cfg<P>::LDA, NRR = Cfg<P>::NRR;
    const int n0 = NRR * ty;
    if (n0 >= P) return;
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
        float4 b = ld4(Bs + r * LDA + 4 * tx);
        if constexpr (NRR % 4 == 0) {
#pragma unroll
            for (int i = 0; i < NRR; i += 4) {
                float4 a = ld4(As + r * LDA + n0 + i);
                racc[i][0] = fmaf(a.x, b.x, racc[i][0]); racc[i][1] = fmaf(a.x, b.y, racc[i][1]);
                racc[i][2] = fmaf(a.x, b.z, racc[i][2]); racc[i][3] = fmaf(a.x, b.w, racc[i][3]);
                racc[i + 1][0] = fmaf(a.y, b.x, racc[i + 1][0]); racc[i + 1][1] = fmaf(a.y, b.y, racc[i + 1][1]);
                racc[i + 1][2] = fmaf(a.y, b.z, racc[i + 1][2]); racc[i + 1][3] = fmaf(a.y, b.w, racc[i + 1][3]);
                racc[i + 2][0] = fmaf(a.z, b.x, racc[i + 2][0]); racc[i + 2][1] = fmaf(a.z, b.y, racc[i + 2][1]);
                racc[i + 2][2] = fmaf(a.z, b.z, racc[i + 2][2]); racc[i + 2][3] = fmaf(a.z, b.w, racc[i + 2][3]);
                racc[i + 3][0] = fmaf(a.w, b.x, racc[i + 3][0]); racc[i + 3][1] = fmaf(a.w, b.y, racc[i + 3][1]);
                racc[i + 3][2] = fmaf(a.w, b.z, racc[i + 3][2]); racc[i + 3][3] = fmaf(a.w, b.w, racc[i + 3][3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NRR; ++i) {
                int n = n0 + i;
                if (n >= P) continue;
                float a = As[r * LDA + n];
                racc[i][0] = fmaf(a, b.x, racc[i][0]); racc[i][1] = fmaf(a, b.y, racc[i][1]);
                racc[i][2] = fmaf(a, b.z, racc[i][2]); racc[i][3] = fmaf(a, b.w, racc[i][3]);
            }
        }
    }
}

// Store the P x P reduction accumulators of this CTA to dst[n*P + k] (assign or accumulate).
template <int P>
__device__ __forceinline__ void store_racc(float* dst, const float (&racc)[Cfg<P>::NRR][4], int tx, int ty,
                                           bool active, bool accumulate) {
    constexpr int NRR = Cfg<P>::NRR;
    if (!active) return;
#pragma unroll
    for (int i = 0; i < NRR; ++i) {
        int n = NRR * ty + i;
        if (n >= P) continue;
        float4 v = make_float4(racc[i][0], racc[i][1], racc[i][2], racc[i][3]);
        float* d = dst + n * P + 4 * tx;
        if (accumulate) v = f4add(ld4(d), v);
        st4(d, v);
    }
}

// Column sums held per thread (four columns, partial over the thread's rows) ->
// dst[0..P) summed over row groups in fixed order.  scratch: RG*P floats of smem.
template <int P>
__device__ __forceinline__ void reduce_columns(float* scratch, const float (&v)[4], float* dst, int tx, int ty,
                                               bool active, bool accumulate) {
    constexpr int RG = Cfg<P>::RG;
    __syncthreads();
    if (active) st4(scratch + ty * P + 4 * tx, make_float4(v[0], v[1], v[2], v[3]));
    __syncthreads();
    if (threadIdx.x < P) {
        float s = 0.f;
        for (int g = 0; g < RG; ++g) s += scratch[g * P + threadIdx.x];
        dst[threadIdx.x] = accumulate ? dst[threadIdx.x] + s : s;
    }
    __syncthreads();
}

// Sum of the neighbour rows of CSR row `li` (float4 slice c4 of each row), fixed order.
template <int LD = 0>
__device__ __forceinline__ float4 gather_row(const float* __restrict__ src, const int32_t* __restrict__ rowptr,
                                             const int32_t* __restrict__ col, int li, int off, int P, int c4) {
    int beg = __ldg(rowptr + li), end = __ldg(rowptr + li + 1);
    float4 acc = f4zero();
    int k = beg;
    for (; k + 4 <= end; k += 4) {
        int j0 = __ldg(col + k), j1 = __ldg(col + k + 1), j2 = __ldg(col + k + 2), j3 = __ldg(col + k + 3);
        float4 v0 = ld_row4<LD>(src + (size_t)(off + j0) * P + 4 * c4);
        float4 v1 = ld_row4<LD>(src + (size_t)(off + j1) * P + 4 * c4);
        float4 v2 = ld_row4<LD>(src + (size_t)(off + j2) * P + 4 * c4);
        float4 v3 = ld_row4<LD>(src + (size_t)(off + j3) * P + 4 * c4);
        acc = f4add(f4add(f4add(f4add(acc, v0), v1), v2), v3);
    }
    if (k < end) {
        int n = end - k;
        int j0 = __ldg(col + k);
        int j1 = n > 1 ? __ldg(col + k + 1) : j0;
        int j2 = n > 2 ? __ldg(col + k + 2) : j0;
        float4 v0 = ld_row4<LD>(src + (size_t)(off + j0) * P + 4 * c4);
        float4 v1 = n > 1 ? ld_row4<LD>(src + (size_t)(off + j1) * P + 4 * c4) : f4zero();
        float4 v2 = n > 2 ? ld_row4<LD>(src + (size_t)(off + j2) * P + 4 * c4) : f4zero();
        acc = f4add(acc, v0);
        if (n > 1) acc = f4add(acc, v1);
        if (n > 2) acc = f4add(acc, v2);
    }
    return acc;
}

// p = 64 gather with memory-level parallelism.  A half-warp (16 lanes x float4 = one 256-byte
// row) produces 8 consecutive tile rows [rl0, rl0+8): (1) lanes 0-7 fetch the CSR extents of the 8
// rows, (2) all 8 neighbour-id lists (up to 16 ids each, one per lane) are fetched in one round
// trip, (3) the neighbour rows are loaded in straight-line batches of 2 rows x 4 neighbours (8
// independent 128-bit loads in flight per lane); rows with more than 4 (16) neighbours take extra
// passes.  The per-row summation order is the CSR order, identical to gather_row, so results are
// bit-identical to the simple form.  sink(rl, acc) is called by all 16 lanes for each row.
// Step (1)+(2) for one tile can be issued while the previous tile is still in its GEMM / epilogue
// (gather8_prefetch), leaving a single memory round trip in front of the data (gather8_consume).
struct Gather8Meta {
    int b, d, off;       // lanes 0-7 (and their mirrors 8-15): CSR begin, degree, graph offset of row rl0 + (lane&7)
    int idx[8];          // lane c: c-th neighbour id of row rl0 + i
};

__device__ __forceinline__ void gather8_prefetch(Gather8Meta& m, const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                                 const int32_t* __restrict__ col, int row0, int rows, int rl0) {
    const unsigned FULL = 0xffffffffu;
    const int c = threadIdx.x & 15;
    m.b = 0; m.d = 0; m.off = 0;
    int rl = rl0 + (c & 7);
    if (rl < rows) {
        int r = row0 + rl, li = r;
        if (g.period > 0) { int gi = r / g.period; m.off = gi * g.period; li = r - m.off; }
        m.b = __ldg(rowptr + li);
        m.d = __ldg(rowptr + li + 1) - m.b;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int bi = __shfl_sync(FULL, m.b, i, 16), di = __shfl_sync(FULL, m.d, i, 16);
        m.idx[i] = c < di ? __ldg(col + bi + c) : 0;
    }
}

template <int LD = 0, class Sink>
__device__ __forceinline__ void gather8_consume(const Gather8Meta& m, const int32_t* __restrict__ col,
                                                const float* __restrict__ src, int rl0, Sink&& sink) {
    const unsigned FULL = 0xffffffffu;
    const int c = threadIdx.x & 15;
    // all warp shuffles first: the loads below must not be separated by convergent operations,
    // otherwise the compiler serialises them row by row (one memory round trip per row)
    int dd[8], oo[8], jj[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dd[i] = __shfl_sync(FULL, m.d, i, 16);
        oo[i] = __shfl_sync(FULL, m.off, i, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) jj[i][k] = __shfl_sync(FULL, m.idx[i], k, 16);
    }
    float4 acc[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {             // two batches of 4 rows x 4 neighbours = 16 independent 128-bit loads
        float4 v[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                v[i][k] = dd[4 * h + i] > k ? ld_row4<LD>(src + (size_t)(oo[4 * h + i] + jj[4 * h + i][k]) * 64 + 4 * c) : f4zero();
#pragma unroll
        for (int i = 0; i < 4; ++i)
            acc[4 * h + i] = f4add(f4add(f4add(f4add(f4zero(), v[i][0]), v[i][1]), v[i][2]), v[i][3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int dmax = max(dd[i], __shfl_xor_sync(FULL, dd[i], 16));    // uniform over the warp
        if (dmax > 4) {
            int bi = __shfl_sync(FULL, m.b, i, 16);
            for (int k = 4; k < dmax; ++k) {
                int j = k < 16 ? __shfl_sync(FULL, m.idx[i], k, 16) : (k < dd[i] ? __ldg(col + bi + k) : 0);
                if (k < dd[i]) acc[i] = f4add(acc[i], ld_row4<LD>(src + (size_t)(oo[i] + j) * 64 + 4 * c));
            }
        }
        sink(rl0 + i, acc[i]);
    }
}

// Same scheme for NR (<= 8) rows per half-warp, loads issued in batches of BATCH rows x 4 neighbours.
template <int NR>
struct GatherMeta {
    int b, d, off;
    int idx[NR];
};

template <int NR>
__device__ __forceinline__ void gatherN_prefetch(GatherMeta<NR>& m, const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                                 const int32_t* __restrict__ col, int row0, int rows, int rl0) {
    const unsigned FULL = 0xffffffffu;
    const int c = threadIdx.x & 15;
    m.b = 0; m.d = 0; m.off = 0;
    int rl = rl0 + (c % NR);
    if (rl < rows) {
        int r = row0 + rl, li = r;
        if (g.period > 0) { int gi = r / g.period; m.off = gi * g.period; li = r - m.off; }
        m.b = __ldg(rowptr + li);
        m.d = __ldg(rowptr + li + 1) - m.b;
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        int bi = __shfl_sync(FULL, m.b, i, 16), di = __shfl_sync(FULL, m.d, i, 16);
        m.idx[i] = c < di ? __ldg(col + bi + c) : 0;
    }
}

// The same prefetch split in its two dependent round trips, so that a software pipeline can keep three tiles in
// flight per lane: CSR extents of tile k+2, neighbour ids of tile k+1, neighbour rows of tile k.
struct GatherExt { int b, d, off; };

template <int NR>
__device__ __forceinline__ void gatherN_extents(GatherExt& e, const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                                int row0, int rows, int rl0) {
    const int c = threadIdx.x & 15;
    e.b = 0; e.d = 0; e.off = 0;
    int rl = rl0 + (c % NR);
    if (rl < rows) {
        int r = row0 + rl, li = r;
        if (g.period > 0) { int gi = r / g.period; e.off = gi * g.period; li = r - e.off; }
        e.b = __ldg(rowptr + li);
        e.d = __ldg(rowptr + li + 1) - e.b;
    }
}

template <int NR>
__device__ __forceinline__ void gatherN_ids(GatherMeta<NR>& m, const GatherExt& e, const int32_t* __restrict__ col) {
    const unsigned FULL = 0xffffffffu;
    const int c = threadIdx.x & 15;
    m.b = e.b; m.d = e.d; m.off = e.off;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        int bi = __shfl_sync(FULL, e.b, i, 16), di = __shfl_sync(FULL, e.d, i, 16);
        m.idx[i] = c < di ? __ldg(col + bi + c) : 0;
    }
}

template <int NR, int BATCH, int LD, class Sink>
__device__ __forceinline__ void gatherN_consume_ld(const GatherMeta<NR>& m, const int32_t* __restrict__ col,
                                                   const float* src, int rl0, Sink&& sink) {
    static_assert(NR % BATCH == 0, "NR must be a multiple of BATCH");
    const unsigned FULL = 0xffffffffu;
    const int c = threadIdx.x & 15;
    int dd[NR], oo[NR], jj[NR][4];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        dd[i] = __shfl_sync(FULL, m.d, i, 16);
        oo[i] = __shfl_sync(FULL, m.off, i, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) jj[i][k] = __shfl_sync(FULL, m.idx[i], k, 16);
    }
    float4 acc[NR];
#pragma unroll
    for (int h = 0; h < NR / BATCH; ++h) {
        float4 v[BATCH][4];
#pragma unroll
        for (int i = 0; i < BATCH; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                v[i][k] = dd[BATCH * h + i] > k
                              ? ld_row4<LD>(src + (size_t)(oo[BATCH * h + i] + jj[BATCH * h + i][k]) * 64 + 4 * c) : f4zero();
#pragma unroll
        for (int i = 0; i < BATCH; ++i)
            acc[BATCH * h + i] = f4add(f4add(f4add(f4add(f4zero(), v[i][0]), v[i][1]), v[i][2]), v[i][3]);
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        int dmax = max(dd[i], __shfl_xor_sync(FULL, dd[i], 16));
        if (dmax > 4) {
            int bi = __shfl_sync(FULL, m.b, i, 16);
            for (int k = 4; k < dmax; ++k) {
                int j = k < 16 ? __shfl_sync(FULL, m.idx[i], k, 16) : (k < dd[i] ? __ldg(col + bi + k) : 0);
                if (k < dd[i]) acc[i] = f4add(acc[i], ld_row4<LD>(src + (size_t)(oo[i] + j) * 64 + 4 * c));
            }
        }
        sink(rl0 + i, acc[i]);
    }
}

template <int NR, int BATCH, class Sink>
__device__ __forceinline__ void gatherN_consume(const GatherMeta<NR>& m, const int32_t* __restrict__ col,
                                                const float* __restrict__ src, int rl0, Sink&& sink) {
    gatherN_consume_ld<NR, BATCH, 0>(m, col, src, rl0, sink);
}

// Software-pipelined form for a persistent producer warp (p = 64, 4 rows per half-warp): one call issues, in this
// order, every warp shuffle it needs (on registers loaded by the PREVIOUS call) and then every load -- the neighbour
// rows of tile k, the neighbour ids of tile k+1 and the CSR extents of tile k+2 -- so that no shuffle sits between
// two loads and a single memory round trip separates two calls.  Rows of degree 5 or 6 (1 % of the NWB rows, but
// one in two 64-row tiles contains such a row) would cost the whole tile a second, serialised round trip in
// gatherN_consume; here the first such row of the half-warp gets its two extra neighbour loads issued with the
// main batch.  Anything beyond (a second long row in the same 4 rows, degree > 6) takes the slow loop.
// Summation order per row = CSR order, bit-identical to gather_row.
struct GatherPipe4 {
    GatherExt ext1;          // CSR extents of the rows of tile k+1 (lane c: row rl0 + c % 4)
    GatherMeta<4> cur;       // extents + neighbour ids of tile k
};

// rows_k2(row0, rows): extent of tile k+2 (rows = 0 past the end); more_loads(): the caller's own loads of tile k,
// issued with the batch.
template <class NextRows, class MoreLoads>
__device__ __forceinline__ void gather4_pipe_step(GatherPipe4& p, const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                                  const int32_t* __restrict__ col, const float* __restrict__ src,
                                                  int rl0, NextRows&& rows_k2, MoreLoads&& more_loads, float4 (&acc)[4]) {
    const unsigned FULL = 0xffffffffu;
    const int c = threadIdx.x & 15;
    // ---- shuffles (registers loaded one call ago) ----
    int dd[4], oo[4], jj[4][4], b1[4], d1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        dd[i] = __shfl_sync(FULL, p.cur.d, i, 16);
        oo[i] = __shfl_sync(FULL, p.cur.off, i, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) jj[i][k] = __shfl_sync(FULL, p.cur.idx[i], k, 16);
        b1[i] = __shfl_sync(FULL, p.ext1.b, i, 16);
        d1[i] = __shfl_sync(FULL, p.ext1.d, i, 16);
    }
    int istar = -1, dstar = 0, ostar = 0, xstar = 0;      // first row of the half-warp with more than 4 neighbours
#pragma unroll
    for (int i = 3; i >= 0; --i)
        if (dd[i] > 4) { istar = i; dstar = dd[i]; ostar = oo[i]; xstar = p.cur.idx[i]; }
    const int j4 = __shfl_sync(FULL, xstar, 4, 16), j5 = __shfl_sync(FULL, xstar, 5, 16);
    // ---- loads ----
    float4 v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[i][k] = dd[i] > k ? ldg4(src + (size_t)(oo[i] + jj[i][k]) * 64 + 4 * c) : f4zero();
    const float4 e4 = dstar > 4 ? ldg4(src + (size_t)(ostar + j4) * 64 + 4 * c) : f4zero();
    const float4 e5 = dstar > 5 ? ldg4(src + (size_t)(ostar + j5) * 64 + 4 * c) : f4zero();
    GatherMeta<4> nxt;
    nxt.b = p.ext1.b; nxt.d = p.ext1.d; nxt.off = p.ext1.off;
#pragma unroll
    for (int i = 0; i < 4; ++i) nxt.idx[i] = c < d1[i] ? __ldg(col + b1[i] + c) : 0;
    GatherExt ext2;
    ext2.b = 0; ext2.d = 0; ext2.off = 0;
    {
        int row0, rows;
        rows_k2(row0, rows);
        int rl = rl0 + (c & 3);
        if (rl < rows) {
            int r = row0 + rl, li = r;
            if (g.period > 0) { int gi = r / g.period; ext2.off = gi * g.period; li = r - ext2.off; }
            ext2.b = __ldg(rowptr + li);
            ext2.d = __ldg(rowptr + li + 1);               // end; turned into the degree below, after the data adds
        }
    }
    more_loads();
    // ---- sums (CSR order) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[i] = f4add(f4add(f4add(f4add(f4zero(), v[i][0]), v[i][1]), v[i][2]), v[i][3]);
        if (i == istar) {                                  // acc is never -0.0, so adding the zero of an absent neighbour is exact
            if (dstar > 4) acc[i] = f4add(acc[i], e4);
            if (dstar > 5) acc[i] = f4add(acc[i], e5);
        }
    }
    // ---- slow path: further long rows (warp-uniform control flow: the shuffles below are full-warp) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k0 = i == istar ? 6 : 4;                 // first neighbour not covered above (per half-warp)
        const int need = dd[i] > k0 ? dd[i] : 0;
        const int dmax = max(need, __shfl_xor_sync(FULL, need, 16));
        if (dmax > 4) {
            const int bi = __shfl_sync(FULL, p.cur.b, i, 16);
            for (int k = 4; k < dmax; ++k) {
                int j = k < 16 ? __shfl_sync(FULL, p.cur.idx[i], k, 16) : (k < need ? __ldg(col + bi + k) : 0);
                if (k >= k0 && k < need) acc[i] = f4add(acc[i], ldg4(src + (size_t)(oo[i] + j) * 64 + 4 * c));
            }
        }
    }
    ext2.d -= ext2.b;
    p.cur = nxt;
    p.ext1 = ext2;
}

template <int LD = 0, class Sink>
__device__ __forceinline__ void gather8_p64(const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                            const int32_t* __restrict__ col, const float* __restrict__ src,
                                            int row0, int rows, int rl0, Sink&& sink) {
    Gather8Meta m;
    gather8_prefetch(m, g, rowptr, col, row0, rows, rl0);
    gather8_consume<LD>(m, col, src, rl0, sink);
}

// Gather a tile: As[rl][:] = sum of neighbour rows of batch row row0+rl (zeros past num_nodes).
template <int P, int LD = 0, int TM = S2V_TM>
__device__ __forceinline__ void gather_tile(float* As, const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                            const int32_t* __restrict__ col, const float* __restrict__ src,
                                            int row0, int rows) {
    constexpr int LDA = Cfg<P>::LDA, LPR = Cfg<P>::LPR, RPP = Cfg<P>::RPP;
    constexpr int ROWS_PER_WARP = TM / (S2V_THREADS / 32);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if constexpr (P == 64 && ROWS_PER_WARP < 16) {          // short tiles: ROWS_PER_WARP / 2 rows per half-warp, same scheme
        constexpr int NRH = ROWS_PER_WARP / 2;
        const int rl0 = warp * ROWS_PER_WARP + NRH * (lane >> 4);
        GatherMeta<NRH> m;
        gatherN_prefetch<NRH>(m, g, rowptr, col, row0, rows, rl0);
        gatherN_consume_ld<NRH, NRH, LD>(m, col, src, rl0,
                                         [&](int rl, const float4& acc) { st4(As + rl * LDA + 4 * (lane & 15), acc); });
        return;
    }
    if constexpr (P == 64) {
        gather8_p64<LD>(g, rowptr, col, src, row0, rows, warp * ROWS_PER_WARP + 8 * (lane >> 4),
                    [&](int rl, const float4& acc) { st4(As + rl * LDA + 4 * (lane & 15), acc); });
        return;
    }
    int sub = lane / LPR, c4 = lane - sub * LPR;
    if (sub >= RPP) return;
    for (int rr = sub; rr < ROWS_PER_WARP; rr += RPP) {
        int rl = warp * ROWS_PER_WARP + rr;
        float4 acc = f4zero();
        if (rl < rows) {
            int li, off, gi;
            int r = row0 + rl;
            if (g.period > 0) { gi = r / g.period; off = gi * g.period; li = r - off; }
            else { li = r; off = 0; }
            acc = gather_row<LD>(src, rowptr, col, li, off, P, c4);
        }
        st4(As + rl * LDA + 4 * c4, acc);
    }
}

#define S2V_RETURN_IF(e) do { int _c = (int)(e); if (_c != 0) return _c; } while (0)

// Persistent launch: one wave of resident CTAs (SM count x occupancy), never more than the tile count.
static inline int s2v_persistent_grid(const void* kernel, int smem_bytes, int num_tiles) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, S2V_THREADS, smem_bytes);
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > S2V_MAX_GRID) grid = S2V_MAX_GRID;
    if (grid > num_tiles) grid = num_tiles;
    return grid < 1 ? 1 : grid;
}

// dst[off0 + e] = sum_{c < n_part} partial[c*stride + off0 + e], e < len: fixed summation order
// (four contiguous c-segments per element, combined in order) => deterministic.
static __global__ void __launch_bounds__(256)
k_reduce_partials(const float* __restrict__ partial, int stride, int n_part, int off0, int len, float* __restrict__ dst) {
    __shared__ float sh[4][64];
    const int x = threadIdx.x & 63, y = threadIdx.x >> 6;
    const int e = blockIdx.x * 64 + x;
    float s = 0.f;
    if (e < len) {
        int seg = (n_part + 3) / 4;
        int c0 = y * seg, c1 = min(n_part, c0 + seg);
        const float* p = partial + (size_t)c0 * stride + off0 + e;
        for (int c = c0; c < c1; ++c, p += stride) s += *p;
    }
    sh[y][x] = s;
    __syncthreads();
    if (y == 0 && e < len) dst[off0 + e] = ((sh[0][x] + sh[1][x]) + sh[2][x]) + sh[3][x];
}
static inline void s2v_reduce_partials(const float* partial, int stride, int n_part, int off0, int len, float* dst,
                                       cudaStream_t st) {
    if (len <= 0) return;
    k_reduce_partials<<<(len + 63) / 64, 256, 0, st>>>(partial, stride, n_part, off0, len, dst);
}

#define S2V_PATH_AUTO 0
#define S2V_PATH_FFMA 1
#define S2V_PATH_TENSOR 2
#define S2V_PATH_TENSOR_WS 3    // tensor path with the warp-specialised (producer / MMA / consumer) backward round
#define S2V_PATH_TENSOR_16 4    // tensor path with the forward-shaped bf16x3 backward round (k_embed_round_bwd_tc16)
#define S2V_TENSOR_MIN_NODES 4096
#ifndef S2V_TENSOR_FINAL
#define S2V_TENSOR_FINAL 0      // tensor-core variant of the last backward stage (slower than FFMA so far)
#endif

#define S2V_DISPATCH_P(P_, CALL)                         \
    switch (P_) {                                        \
        case 8: { constexpr int PP = 8; CALL; } break;   \
        case 16: { constexpr int PP = 16; CALL; } break; \
        case 24: { constexpr int PP = 24; CALL; } break; \
        case 32: { constexpr int PP = 32; CALL; } break; \
        case 48: { constexpr int PP = 48; CALL; } break; \
        case 64: { constexpr int PP = 64; CALL; } break; \
        default: return S2V_ERR_UNSUPPORTED;             \
    }
