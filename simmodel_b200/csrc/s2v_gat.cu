// GATv2 edge-attention layer (--qnet gat2), forward and backward, sm_100a.
//
// One GATv2Conv layer (PyG ~2.0.4, concat heads, add_self_loops, share_weights=False; call sites
// comb_opt.py:48-62,163-171 and models_basic_lstm.py:326-338,414-426) splits into
//   (1) two node-wise linear layers  x_l = lin_l(x), x_r = lin_r(x)              -- plain GEMMs, done by the host
//   (2) the per-edge attention + per-target segment softmax + weighted neighbour sum  -- the kernels below.
// For a target i with incoming sources j (the graph arrives WITHOUT self loops; every kernel appends the
// self edge i -> i last, which is the order remove_self_loops + add_self_loops produces):
//   e_ij[h]  = sum_c att[h,c] * lrelu(x_r[i,h,c] + x_l[j,h,c]),  lrelu slope 0.2
//   a_ij[h]  = exp(e_ij - max_j e_ij) / (sum_j exp(e_ij - max_j) + 1e-16)
//   out[i]   = sum_j a_ij[h] * x_l[j,h,:] + bias        (optional ReLU: BasicGNN applies it between layers)
//
// Mapping: one warp per target (forward, backward A) or per source (backward B); lane l owns VEC consecutive
// channels [l*VEC, l*VEC+VEC) of the H*C-wide row, so a row is one coalesced 128/64/32-bit load per lane and
// a head is a group of C/VEC consecutive lanes reduced with shuffles.  Rows with at most GAT_FAST neighbours
// (incl. the self edge; every reference graph: max degree 6) keep all neighbour rows in registers: one memory
// round trip per row, all shuffles hoisted in front of the loads.  Longer rows take a looped path that re-reads
// the neighbour rows from L1/L2.  Nothing per-edge is stored: backward recomputes e and a from the saved
// per-node softmax statistics (max, denominator).  No atomics; att/bias gradients are per-CTA partial sums
// reduced in fixed order.
#include "s2v_common.cuh"

#define GAT_THREADS 256
#define GAT_WARPS (GAT_THREADS / 32)
#define GAT_FAST 8
#define GAT_SLOPE 0.2f

namespace {

template <int VEC> struct Vec { float v[VEC]; };

template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv(const float* p) {
    Vec<VEC> r;
    if (VEC == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else if (VEC == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); r.v[0] = t.x; r.v[1] = t.y; }
    else { r.v[0] = __ldg(p); }
    return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const Vec<VEC>& r) {
    if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else if (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
    else *p = r.v[0];
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> vzero() { Vec<VEC> r;
#pragma unroll
    for (int v = 0; v < VEC; ++v) r.v[v] = 0.f;
    return r; }

template <int VEC>
__device__ __forceinline__ Vec<VEC> vadd(Vec<VEC> a, const Vec<VEC>& b) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) a.v[v] += b.v[v];
    return a;
}
__device__ __forceinline__ float lrelu(float s) { return fmaxf(s, s * GAT_SLOPE); }   // slope < 1
// ex / den from the stored reciprocal with one Newton step (== the correctly rounded quotient up to the last bit)
__device__ __forceinline__ float qdiv(float ex, float inv, float den) {
    const float q = ex * inv;
    return fmaf(fmaf(-q, den, ex), inv, q);
}

// sum over the lanes of one head (lph consecutive lanes); every lane of the warp takes part
__device__ __forceinline__ float head_sum(float v, int lph, int lane) {
    if ((lph & (lph - 1)) == 0) {
        for (int o = lph >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    int base = lane - lane % lph;
    float s = 0.f;
    for (int l = 0; l < lph; ++l) s += __shfl_sync(0xffffffffu, v, (base + l) & 31);
    return s;
}

template <int VEC>
__device__ __forceinline__ float score_partial(const Vec<VEC>& a, const Vec<VEC>& xr, const Vec<VEC>& xl) {
    float e = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) e = fmaf(a.v[v], lrelu(xr.v[v] + xl.v[v]), e);
    return e;
}
template <int VEC>
__device__ __forceinline__ float dot_partial(const Vec<VEC>& a, const Vec<VEC>& b) {
    float e = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) e = fmaf(a.v[v], b.v[v], e);
    return e;
}

// ------------------------------------------------------------------------------------------------------------
// forward: out[i] = sum_j softmax_j(e_ij) x_l[j] + bias ; stats[i] = (max, denominator) per head
// ------------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_fwd(const s2v_graph_t g, int H, int C, int ld, const float* __restrict__ xl, const float* __restrict__ xr,
          const float* __restrict__ att, const float* __restrict__ bias, const float* __restrict__ lin_bias,
          int apply_relu, float* __restrict__ out, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int HC = H * C, lanes = HC / VEC, lph = C / VEC;
    const bool active = lane < lanes;
    const int ch = active ? lane * VEC : 0;
    const int head = ch / C;
    const Vec<VEC> a = ldv<VEC>(att + ch);
    const Vec<VEC> b = ldv<VEC>(bias + ch);
    // x_l / x_r may arrive without their Linear biases (see the header): every row gets its bias as it is loaded, the
    // same fp32 add a GEMM epilogue would have done
    const Vec<VEC> bl = lin_bias ? ldv<VEC>(lin_bias + ch) : vzero<VEC>(), br = lin_bias ? ldv<VEC>(lin_bias + HC + ch) : vzero<VEC>();
    const int warp = blockIdx.x * GAT_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GAT_WARPS;
    for (int r = warp; r < g.num_nodes; r += nwarps) {
        int li, off, gi;
        row_info(g, r, li, off, gi);
        const int beg = __ldg(g.rowptr_t + li), deg = __ldg(g.rowptr_t + li + 1) - beg, cnt = deg + 1;
        const Vec<VEC> xri = vadd<VEC>(ldv<VEC>(xr + (size_t)r * ld + ch), br);
        Vec<VEC> acc = vzero<VEC>();
        float m = -INFINITY, inv, den;
        if (cnt <= GAT_FAST) {
            const int jl = lane < deg ? __ldg(g.col_t + beg + lane) + off : r;
            int js[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) js[k] = __shfl_sync(0xffffffffu, jl, k);
            Vec<VEC> rows[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) rows[k] = k < cnt ? vadd<VEC>(ldv<VEC>(xl + (size_t)js[k] * ld + ch), bl) : vzero<VEC>();
            float sc[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) sc[k] = head_sum(active ? score_partial<VEC>(a, xri, rows[k]) : 0.f, lph, lane);
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) m = fmaxf(m, sc[k]);
            float S = 0.f;
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) { sc[k] = expf(sc[k] - m); S += sc[k]; }
            den = S + 1e-16f; inv = 1.f / den;
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) {
                const float al = qdiv(sc[k], inv, den);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc.v[v] = fmaf(rows[k].v[v], al, acc.v[v]);
            }
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = vadd<VEC>(ldv<VEC>(xl + (size_t)j * ld + ch), bl);
                m = fmaxf(m, head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane));
            }
            float S = 0.f;
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = vadd<VEC>(ldv<VEC>(xl + (size_t)j * ld + ch), bl);
                S += expf(head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane) - m);
            }
            den = S + 1e-16f; inv = 1.f / den;
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = vadd<VEC>(ldv<VEC>(xl + (size_t)j * ld + ch), bl);
                const float al = qdiv(expf(head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane) - m), inv, den);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc.v[v] = fmaf(row.v[v], al, acc.v[v]);
            }
        }
        if (active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { float o = acc.v[v] + b.v[v]; acc.v[v] = apply_relu ? relu(o) : o; }
            stv<VEC>(out + (size_t)r * HC + ch, acc);
            if (lane % lph == 0) *reinterpret_cast<float4*>(stats + ((size_t)r * H + head) * 4) = make_float4(m, inv, 0.f, den);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward A (warp per target i): D_i = sum_j a_ij da_ij, dx_r[i], partial sums of datt and dbias
// ------------------------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void bwd_a_edge(const Vec<VEC>& a, const Vec<VEC>& xri, const Vec<VEC>& row, float al, float dal, float D,
                                           Vec<VEC>& dxr, Vec<VEC>& datt) {
    const float de = al * (dal - D);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float s = xri.v[v] + row.v[v];
        dxr.v[v] = fmaf(de * a.v[v], s > 0.f ? 1.f : GAT_SLOPE, dxr.v[v]);
        datt.v[v] = fmaf(de, lrelu(s), datt.v[v]);
    }
}

template <int VEC>
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_bwd_target(const s2v_graph_t g, int H, int C, int ld, const float* __restrict__ xl, const float* __restrict__ xr,
                 const float* __restrict__ att, const float* __restrict__ lin_bias, float* __restrict__ stats,
                 const float* __restrict__ dout, float* __restrict__ dxr, float* __restrict__ partials) {
    __shared__ float red[GAT_WARPS][3 * 128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int HC = H * C, lanes = HC / VEC, lph = C / VEC;
    const bool active = lane < lanes;
    const int ch = active ? lane * VEC : 0;
    const int head = ch / C;
    const Vec<VEC> a = ldv<VEC>(att + ch);
    const Vec<VEC> bl = lin_bias ? ldv<VEC>(lin_bias + ch) : vzero<VEC>(), br = lin_bias ? ldv<VEC>(lin_bias + HC + ch) : vzero<VEC>();
    Vec<VEC> datt = vzero<VEC>(), dbias = vzero<VEC>(), dbr = vzero<VEC>();
    const int warp = blockIdx.x * GAT_WARPS + wid, nwarps = gridDim.x * GAT_WARPS;
    for (int r = warp; r < g.num_nodes; r += nwarps) {
        int li, off, gi;
        row_info(g, r, li, off, gi);
        const int beg = __ldg(g.rowptr_t + li), deg = __ldg(g.rowptr_t + li + 1) - beg, cnt = deg + 1;
        const Vec<VEC> xri = vadd<VEC>(ldv<VEC>(xr + (size_t)r * ld + ch), br);
        const Vec<VEC> doi = ldv<VEC>(dout + (size_t)r * HC + ch);
        const float m = stats[((size_t)r * H + head) * 4], inv = stats[((size_t)r * H + head) * 4 + 1], den = stats[((size_t)r * H + head) * 4 + 3];
        Vec<VEC> dx = vzero<VEC>();
        float D = 0.f;
        if (cnt <= GAT_FAST) {
            const int jl = lane < deg ? __ldg(g.col_t + beg + lane) + off : r;
            int js[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) js[k] = __shfl_sync(0xffffffffu, jl, k);
            Vec<VEC> rows[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) rows[k] = k < cnt ? vadd<VEC>(ldv<VEC>(xl + (size_t)js[k] * ld + ch), bl) : vzero<VEC>();
            float al[GAT_FAST], dal[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) {
                const float e = head_sum(active ? score_partial<VEC>(a, xri, rows[k]) : 0.f, lph, lane);
                dal[k] = head_sum(active ? dot_partial<VEC>(doi, rows[k]) : 0.f, lph, lane);
                al[k] = k < cnt ? qdiv(expf(e - m), inv, den) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) D = fmaf(al[k], dal[k], D);
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) bwd_a_edge<VEC>(a, xri, rows[k], al[k], dal[k], D, dx, datt);
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = vadd<VEC>(ldv<VEC>(xl + (size_t)j * ld + ch), bl);
                const float e = head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(doi, row) : 0.f, lph, lane);
                D = fmaf(qdiv(expf(e - m), inv, den), dl, D);
            }
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = vadd<VEC>(ldv<VEC>(xl + (size_t)j * ld + ch), bl);
                const float e = head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(doi, row) : 0.f, lph, lane);
                bwd_a_edge<VEC>(a, xri, row, qdiv(expf(e - m), inv, den), dl, D, dx, datt);
            }
        }
        if (active) {
            stv<VEC>(dxr + (size_t)r * ld + ch, dx);
            if (lane % lph == 0) stats[((size_t)r * H + head) * 4 + 2] = D;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { dbias.v[v] += doi.v[v]; dbr.v[v] += dx.v[v]; }
        }
    }
    // per-CTA partial sums, fixed order over warps
    if (active) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { red[wid][ch + v] = datt.v[v]; red[wid][128 + ch + v] = dbias.v[v]; red[wid][256 + ch + v] = dbr.v[v]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * HC; c += GAT_THREADS) {
        const int idx = 128 * (c / HC) + c % HC;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < GAT_WARPS; ++w) s += red[w][idx];
        partials[(size_t)blockIdx.x * 4 * HC + c] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward B (warp per source j): dx_l[j] = sum_{i: j -> i} a_ij dout[i] + de_ij att lrelu'(x_r[i] + x_l[j])
// ------------------------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void bwd_b_edge(const Vec<VEC>& a, const Vec<VEC>& xlj, const Vec<VEC>& xri, const Vec<VEC>& doi,
                                           float al, float dal, float D, Vec<VEC>& dx) {
    const float de = al * (dal - D);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float s = xri.v[v] + xlj.v[v];
        dx.v[v] = fmaf(al, doi.v[v], dx.v[v]);
        dx.v[v] = fmaf(de * a.v[v], s > 0.f ? 1.f : GAT_SLOPE, dx.v[v]);
    }
}

template <int VEC>
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_bwd_source(const s2v_graph_t g, int H, int C, int ld, const float* __restrict__ xl, const float* __restrict__ xr,
                 const float* __restrict__ att, const float* __restrict__ lin_bias, const float* __restrict__ stats,
                 const float* __restrict__ dout, float* __restrict__ dxl, float* __restrict__ partials) {
    __shared__ float red[GAT_WARPS][128];
    Vec<VEC> dbl = vzero<VEC>();
    const int lane = threadIdx.x & 31;
    const int HC = H * C, lanes = HC / VEC, lph = C / VEC;
    const bool active = lane < lanes;
    const int ch = active ? lane * VEC : 0;
    const int head = ch / C;
    const Vec<VEC> a = ldv<VEC>(att + ch);
    const Vec<VEC> bl = lin_bias ? ldv<VEC>(lin_bias + ch) : vzero<VEC>(), br = lin_bias ? ldv<VEC>(lin_bias + HC + ch) : vzero<VEC>();
    const int warp = blockIdx.x * GAT_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GAT_WARPS;
    for (int r = warp; r < g.num_nodes; r += nwarps) {
        int li, off, gi;
        row_info(g, r, li, off, gi);
        const int beg = __ldg(g.rowptr + li), deg = __ldg(g.rowptr + li + 1) - beg, cnt = deg + 1;
        const Vec<VEC> xlj = vadd<VEC>(ldv<VEC>(xl + (size_t)r * ld + ch), bl);
        Vec<VEC> dx = vzero<VEC>();
        if (cnt <= GAT_FAST) {
            const int il = lane < deg ? __ldg(g.col + beg + lane) + off : r;
            int is[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) is[k] = __shfl_sync(0xffffffffu, il, k);
            Vec<VEC> xrs[GAT_FAST], dos[GAT_FAST];
            float4 st[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) {
                const bool on = k < cnt;
                xrs[k] = on ? vadd<VEC>(ldv<VEC>(xr + (size_t)is[k] * ld + ch), br) : vzero<VEC>();
                dos[k] = on ? ldv<VEC>(dout + (size_t)is[k] * HC + ch) : vzero<VEC>();
                st[k] = on ? ldg4(stats + ((size_t)is[k] * H + head) * 4) : f4zero();
            }
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) {
                const float e = head_sum(active ? score_partial<VEC>(a, xrs[k], xlj) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(dos[k], xlj) : 0.f, lph, lane);
                if (k < cnt) bwd_b_edge<VEC>(a, xlj, xrs[k], dos[k], qdiv(expf(e - st[k].x), st[k].y, st[k].w), dl, st[k].z, dx);
            }
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int i = k < deg ? __ldg(g.col + beg + k) + off : r;
                const Vec<VEC> xri = vadd<VEC>(ldv<VEC>(xr + (size_t)i * ld + ch), br);
                const Vec<VEC> doi = ldv<VEC>(dout + (size_t)i * HC + ch);
                const float4 st = ldg4(stats + ((size_t)i * H + head) * 4);
                const float e = head_sum(active ? score_partial<VEC>(a, xri, xlj) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(doi, xlj) : 0.f, lph, lane);
                bwd_b_edge<VEC>(a, xlj, xri, doi, qdiv(expf(e - st.x), st.y, st.w), dl, st.z, dx);
            }
        }
        if (active) {
            stv<VEC>(dxl + (size_t)r * ld + ch, dx);
#pragma unroll
            for (int v = 0; v < VEC; ++v) dbl.v[v] += dx.v[v];
        }
    }
    if (active) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) red[threadIdx.x >> 5][ch + v] = dbl.v[v];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < HC; c += GAT_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < GAT_WARPS; ++w) s += red[w][c];
        partials[(size_t)blockIdx.x * 4 * HC + 3 * HC + c] = s;
    }
}

// ============================================================================================================
// Specialised kernels for the reference's shapes: 128-bit lanes, LPR lanes per row (H*C = 4*LPR), LPH lanes per
// head (C = 4*LPH), 32/LPR rows per warp; CSR extents and neighbour ids of the next rows are prefetched two / one
// iterations ahead (one memory round trip per row on the critical path instead of three).
//   <32,8>: H*C = 128, C = 32 (PPO hidden layers)   <16,8>: H*C = 64, C = 32 (DQN QNet_GATv2)
//   <16,4>: H*C = 64,  C = 16 (PPO output layer)
// ============================================================================================================
template <int LPH>
__device__ __forceinline__ float hsum(float v) {
#pragma unroll
    for (int o = LPH >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float score4(const float4& a, const float4& xr, const float4& xl) {
    float e = a.x * lrelu(xr.x + xl.x);
    e = fmaf(a.y, lrelu(xr.y + xl.y), e);
    e = fmaf(a.z, lrelu(xr.z + xl.z), e);
    return fmaf(a.w, lrelu(xr.w + xl.w), e);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ float dlr(float s) { return s > 0.f ? 1.f : GAT_SLOPE; }

// Pipelined row metadata: (beg, deg, off, graph copy) of a row from the CSR given by (rp, ci); ids of its first LPR
// neighbours.
struct RowMeta { int r, beg, deg, off, gi; };
__device__ __forceinline__ RowMeta row_meta(const s2v_graph_t& g, const int* __restrict__ rp, int r) {
    RowMeta m; m.r = r; m.beg = 0; m.deg = -1; m.off = 0; m.gi = 0;          // deg -1 => cnt 0: row out of range
    if (r < g.num_nodes) {
        int li = r;
        if (g.period > 0) { m.gi = r / g.period; m.off = m.gi * g.period; li = r - m.off; }
        m.beg = __ldg(rp + li);
        m.deg = __ldg(rp + li + 1) - m.beg;
    }
    return m;
}
__device__ __forceinline__ int row_ids(const RowMeta& m, const int* __restrict__ ci, int sl) {
    return sl < m.deg ? __ldg(ci + m.beg + sl) + m.off : m.r;
}
__device__ __forceinline__ int row_map(const RowMeta& m, const int* __restrict__ map, int sl) {
    return sl < m.deg ? __ldg(map + m.beg + sl) : -1;
}

#define GAT_ROW_SETUP()                                                                                         \
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;                                       \
    const int ch = sl * 4, head = sl / LPH;                                                                     \
    const int step = gridDim.x * GAT_WARPS * RPW;                                                               \
    int r0 = (blockIdx.x * GAT_WARPS + (threadIdx.x >> 5)) * RPW;

#define GAT_ROW_LOOP(RP, CI)                                                                                    \
    RowMeta cur = row_meta(g, RP, r0 + sub);                                                                    \
    int jl = row_ids(cur, CI, sl);                                                                              \
    RowMeta nxt = row_meta(g, RP, r0 + step + sub);                                                             \
    for (; r0 < g.num_nodes; r0 += step)

#define GAT_ROW_ADVANCE(RP, CI)                                                                                 \
    const int jl_n = row_ids(nxt, CI, sl);                                                                      \
    const RowMeta nn = row_meta(g, RP, nxt.r + step);

#define GAT_ROW_ROTATE() cur = nxt; jl = jl_n; nxt = nn;

template <int LPR>
__device__ __forceinline__ int warp_max_cnt(int cnt) {
    if (LPR < 32) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 16));
    return cnt;
}

// Per-edge attention weights: alpha / dedge are [E_total + n][H] with E_total = copies * E_csr -- the incoming edges of
// target r at ((gi * E_csr + beg_t + k) * H + head) in column-form (CSC) order, its self edge at ((E_total + r) * H + head).
struct EdgeIdx {
    size_t e0, s0;          // index of edge 0 / of the self edge for this lane's head
    __device__ __forceinline__ size_t at(int k, int deg, int H) const { return k < deg ? e0 + (size_t)k * H : s0; }
};
template <int H>
__device__ __forceinline__ EdgeIdx edge_idx(const s2v_graph_t& g, const RowMeta& m, int head) {
    const int e_csr = g.num_edges;
    const size_t copies = g.period > 0 ? (size_t)(g.num_nodes / g.period) : 1;
    EdgeIdx x;
    x.e0 = ((size_t)m.gi * e_csr + m.beg) * H + head;
    x.s0 = (copies * e_csr + (size_t)m.r) * H + head;
    return x;
}

// The NS per-edge values of a head are identical in all LPH lanes of the head: lane q of the head stores the value of
// edge q (and q + LPH when NS > LPH), so one store instruction covers every edge of the rows of the warp.
template <int LPH, int NS, int H>
__device__ __forceinline__ void store_edge_values(float* __restrict__ dst, const float (&v)[NS], int q, int cnt, const EdgeIdx& ex) {
#pragma unroll
    for (int base = 0; base < NS; base += LPH) {
        const int k = base + q;
        float x = v[base];
#pragma unroll
        for (int i = 1; i < LPH; ++i) if (base + i < NS) x = (q == i) ? v[base + i] : x;
        if (k < cnt && k < NS) dst[ex.at(k, cnt - 1, H)] = x;
    }
}

// Register-resident paths for rows with at most NS neighbours (incl. the self edge); NS = 4 or 8 is chosen per warp.
template <int LPR, int LPH, int NS>
__device__ __forceinline__ void fwd_fast(const float* __restrict__ xl, int ld, int ch, int jl, int cnt, const float4& a,
                                         const float4& bl, const float4& xri, float4& acc, float* __restrict__ alpha,
                                         const EdgeIdx& ex, int q) {
    constexpr int H = LPR / LPH;
    int js[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) js[k] = __shfl_sync(0xffffffffu, jl, k, LPR);
    float4 rows[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) rows[k] = k < cnt ? f4add(ldg4(xl + (size_t)js[k] * ld + ch), bl) : f4zero();
    float sc[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) sc[k] = hsum<LPH>(score4(a, xri, rows[k]));
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NS; ++k) if (k < cnt) m = fmaxf(m, sc[k]);
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < NS; ++k) if (k < cnt) { sc[k] = expf(sc[k] - m); S += sc[k]; }
    const float den = S + 1e-16f, inv = 1.f / den;
#pragma unroll
    for (int k = 0; k < NS; ++k) if (k < cnt) {
        sc[k] = qdiv(sc[k], inv, den);
        acc.x = fmaf(rows[k].x, sc[k], acc.x); acc.y = fmaf(rows[k].y, sc[k], acc.y);
        acc.z = fmaf(rows[k].z, sc[k], acc.z); acc.w = fmaf(rows[k].w, sc[k], acc.w);
    }
    store_edge_values<LPH, NS, H>(alpha, sc, q, cnt, ex);
}

template <int LPR, int LPH>
__global__ void __launch_bounds__(GAT_THREADS, 4)   // latency-bound: occupancy beats the few spilled registers (A/B measured)
k_gat_fwd4(const s2v_graph_t g, int ld, const float* __restrict__ xl, const float* __restrict__ xr,
           const float* __restrict__ att, const float* __restrict__ bias, const float* __restrict__ lin_bias,
           int apply_relu, float* __restrict__ out, float* __restrict__ alpha) {
    constexpr int RPW = 32 / LPR, HC = 4 * LPR, H = LPR / LPH;
    GAT_ROW_SETUP()
    const float4 a = ldg4(att + ch), b = ldg4(bias + ch);
    // x_l / x_r may arrive without their Linear biases: every row gets its bias as it is loaded
    const float4 bl = lin_bias ? ldg4(lin_bias + ch) : f4zero(), br = lin_bias ? ldg4(lin_bias + HC + ch) : f4zero();
    const int q = sl % LPH;
    const bool writer = q == 0;
    GAT_ROW_LOOP(g.rowptr_t, g.col_t) {
        GAT_ROW_ADVANCE(g.rowptr_t, g.col_t)
        const int r = cur.r, cnt = cur.deg + 1;
        const bool valid = cnt > 0;
        const float4 xri = valid ? f4add(ldg4(xr + (size_t)r * ld + ch), br) : f4zero();
        const EdgeIdx ex = edge_idx<H>(g, cur, head);
        float4 acc = f4zero();
        const int mc = warp_max_cnt<LPR>(cnt);
        if (mc <= 4) fwd_fast<LPR, LPH, 4>(xl, ld, ch, jl, cnt, a, bl, xri, acc, alpha, ex, q);
        else if (mc <= GAT_FAST) fwd_fast<LPR, LPH, GAT_FAST>(xl, ld, ch, jl, cnt, a, bl, xri, acc, alpha, ex, q);
        else {
            float m = -INFINITY, S = 0.f, inv = 0.f, den = 1.f;
            for (int pass = 0; pass < 3; ++pass) {
                for (int k = 0; k < mc; ++k) {
                    const bool on = k < cnt;
                    const int j = on && k < cur.deg ? __ldg(g.col_t + cur.beg + k) + cur.off : r;
                    const float4 row = on ? f4add(ldg4(xl + (size_t)j * ld + ch), bl) : f4zero();
                    const float e = hsum<LPH>(score4(a, xri, row));
                    if (!on) continue;
                    if (pass == 0) m = fmaxf(m, e);
                    else if (pass == 1) S += expf(e - m);
                    else {
                        const float al = qdiv(expf(e - m), inv, den);
                        if (writer) alpha[ex.at(k, cur.deg, H)] = al;
                        acc.x = fmaf(row.x, al, acc.x); acc.y = fmaf(row.y, al, acc.y);
                        acc.z = fmaf(row.z, al, acc.z); acc.w = fmaf(row.w, al, acc.w);
                    }
                }
                if (pass == 1) { den = S + 1e-16f; inv = 1.f / den; }
            }
        }
        if (valid) {
            float4 o = make_float4(acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w);
            if (apply_relu) o = make_float4(relu(o.x), relu(o.y), relu(o.z), relu(o.w));
            st4(out + (size_t)r * HC + ch, o);
        }
        GAT_ROW_ROTATE()
    }
}

__device__ __forceinline__ void bwd_a_edge4(const float4& a, const float4& xri, const float4& row, float de, float4& dx, float4& datt) {
    float s;
    s = xri.x + row.x; dx.x = fmaf(de * a.x, dlr(s), dx.x); datt.x = fmaf(de, lrelu(s), datt.x);
    s = xri.y + row.y; dx.y = fmaf(de * a.y, dlr(s), dx.y); datt.y = fmaf(de, lrelu(s), datt.y);
    s = xri.z + row.z; dx.z = fmaf(de * a.z, dlr(s), dx.z); datt.z = fmaf(de, lrelu(s), datt.z);
    s = xri.w + row.w; dx.w = fmaf(de * a.w, dlr(s), dx.w); datt.w = fmaf(de, lrelu(s), datt.w);
}

template <int LPR, int LPH, int NS>
__device__ __forceinline__ void bwd_target_fast(const float* __restrict__ xl, int ld, int ch, int jl, int cnt, const float4& a,
                                                const float4& bl, const float4& xri, const float4& doi,
                                                const float* __restrict__ alpha, float* __restrict__ dedge, const EdgeIdx& ex,
                                                int q, float4& dx, float4& datt) {
    constexpr int H = LPR / LPH;
    int js[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) js[k] = __shfl_sync(0xffffffffu, jl, k, LPR);
    float4 rows[NS];
    float al[NS], dal[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        rows[k] = k < cnt ? f4add(ldg4(xl + (size_t)js[k] * ld + ch), bl) : f4zero();
        al[k] = k < cnt ? __ldg(alpha + ex.at(k, cnt - 1, H)) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) dal[k] = hsum<LPH>(dot4(doi, rows[k]));
    float D = 0.f;
#pragma unroll
    for (int k = 0; k < NS; ++k) if (k < cnt) D = fmaf(al[k], dal[k], D);
#pragma unroll
    for (int k = 0; k < NS; ++k) if (k < cnt) {
        dal[k] = al[k] * (dal[k] - D);
        bwd_a_edge4(a, xri, rows[k], dal[k], dx, datt);
    }
    store_edge_values<LPH, NS, H>(dedge, dal, q, cnt, ex);
}

template <int LPR, int LPH>
__global__ void __launch_bounds__(GAT_THREADS, 3)
k_gat_bwd_target4(const s2v_graph_t g, int ld, const float* __restrict__ xl, const float* __restrict__ xr,
                  const float* __restrict__ att, const float* __restrict__ lin_bias, const float* __restrict__ alpha,
                  float* __restrict__ dedge, const float* __restrict__ dout, float* __restrict__ dxr,
                  float* __restrict__ partials) {
    constexpr int RPW = 32 / LPR, HC = 4 * LPR, H = LPR / LPH;
    __shared__ float red[GAT_WARPS * RPW][3 * HC];
    float4 datt = f4zero(), dbias = f4zero(), dbr = f4zero();
    GAT_ROW_SETUP()
    const float4 a = ldg4(att + ch);
    const float4 bl = lin_bias ? ldg4(lin_bias + ch) : f4zero(), br = lin_bias ? ldg4(lin_bias + HC + ch) : f4zero();
    const int q = sl % LPH;
    const bool writer = q == 0;
    GAT_ROW_LOOP(g.rowptr_t, g.col_t) {
        GAT_ROW_ADVANCE(g.rowptr_t, g.col_t)
        const int r = cur.r, cnt = cur.deg + 1;
        const bool valid = cnt > 0;
        const float4 xri = valid ? f4add(ldg4(xr + (size_t)r * ld + ch), br) : f4zero();
        const float4 doi = valid ? ldg4(dout + (size_t)r * HC + ch) : f4zero();
        const EdgeIdx ex = edge_idx<H>(g, cur, head);
        float4 dx = f4zero();
        const int mc = warp_max_cnt<LPR>(cnt);
        if (mc <= 4) bwd_target_fast<LPR, LPH, 4>(xl, ld, ch, jl, cnt, a, bl, xri, doi, alpha, dedge, ex, q, dx, datt);
        else if (mc <= GAT_FAST) bwd_target_fast<LPR, LPH, GAT_FAST>(xl, ld, ch, jl, cnt, a, bl, xri, doi, alpha, dedge, ex, q, dx, datt);
        else {
            float D = 0.f;
            for (int pass = 0; pass < 2; ++pass)
                for (int k = 0; k < mc; ++k) {
                    const bool on = k < cnt;
                    const int j = on && k < cur.deg ? __ldg(g.col_t + cur.beg + k) + cur.off : r;
                    const float4 row = on ? f4add(ldg4(xl + (size_t)j * ld + ch), bl) : f4zero();
                    const float dl = hsum<LPH>(dot4(doi, row));
                    if (!on) continue;
                    const float alk = __ldg(alpha + ex.at(k, cur.deg, H));
                    if (pass == 0) D = fmaf(alk, dl, D);
                    else {
                        const float de = alk * (dl - D);
                        if (writer) dedge[ex.at(k, cur.deg, H)] = de;
                        bwd_a_edge4(a, xri, row, de, dx, datt);
                    }
                }
        }
        if (valid) {
            st4(dxr + (size_t)r * ld + ch, dx);
            dbias = f4add(dbias, doi);
            dbr = f4add(dbr, dx);
        }
        GAT_ROW_ROTATE()
    }
    {
        const int lane_ = threadIdx.x & 31, slot = (threadIdx.x >> 5) * RPW + lane_ / LPR, c4 = (lane_ % LPR) * 4;
        st4(&red[slot][c4], datt);
        st4(&red[slot][HC + c4], dbias);
        st4(&red[slot][2 * HC + c4], dbr);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * HC; c += GAT_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < GAT_WARPS * RPW; ++w) s += red[w][c];
        partials[(size_t)blockIdx.x * 4 * HC + c] = s;
    }
}

__device__ __forceinline__ void bwd_b_edge4(const float4& a, const float4& xlj, const float4& xri, const float4& doi,
                                            float al, float de, float4& dx) {
    dx.x = fmaf(al, doi.x, dx.x); dx.x = fmaf(de * a.x, dlr(xri.x + xlj.x), dx.x);
    dx.y = fmaf(al, doi.y, dx.y); dx.y = fmaf(de * a.y, dlr(xri.y + xlj.y), dx.y);
    dx.z = fmaf(al, doi.z, dx.z); dx.z = fmaf(de * a.z, dlr(xri.z + xlj.z), dx.z);
    dx.w = fmaf(al, doi.w, dx.w); dx.w = fmaf(de * a.w, dlr(xri.w + xlj.w), dx.w);
}

// Batch of four out-edges of a source: no score, no softmax -- alpha and de = alpha (dalpha - D) of every edge were
// stored by the target passes (column-form order) and are fetched through the row-form -> column-form edge map.
template <int LPR, int LPH>
__device__ __forceinline__ void bwd_source_fast(const float* __restrict__ xr, const float* __restrict__ dout,
                                                const float* __restrict__ alpha, const float* __restrict__ dedge, int ld, int ch,
                                                int il, int ml, int cnt, int k0, size_t ebase, size_t sidx, int head,
                                                const float4& a, const float4& br, const float4& xlj, float4& dx) {
    constexpr int NS = 4, HC = 4 * LPR, H = LPR / LPH;
    int is[NS], ms[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) { is[k] = __shfl_sync(0xffffffffu, il, k0 + k, LPR); ms[k] = __shfl_sync(0xffffffffu, ml, k0 + k, LPR); }
    cnt -= k0;
    float4 xrs[NS], dos[NS];
    float al[NS], de[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const bool on = k < cnt;
        const size_t idx = ms[k] >= 0 ? (ebase + (size_t)ms[k]) * H + head : sidx;
        xrs[k] = on ? f4add(ldg4(xr + (size_t)is[k] * ld + ch), br) : f4zero();
        dos[k] = on ? ldg4(dout + (size_t)is[k] * HC + ch) : f4zero();
        al[k] = on ? __ldg(alpha + idx) : 0.f;
        de[k] = on ? __ldg(dedge + idx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) if (k < cnt) bwd_b_edge4(a, xlj, xrs[k], dos[k], al[k], de[k], dx);
}

template <int LPR, int LPH>
__global__ void __launch_bounds__(GAT_THREADS, 3)
k_gat_bwd_source4(const s2v_graph_t g, int ld, const float* __restrict__ xl, const float* __restrict__ xr,
                  const float* __restrict__ att, const float* __restrict__ lin_bias, const float* __restrict__ alpha,
                  const float* __restrict__ dedge, const int* __restrict__ row_to_col, const float* __restrict__ dout,
                  float* __restrict__ dxl, float* __restrict__ partials) {
    constexpr int RPW = 32 / LPR, HC = 4 * LPR, H = LPR / LPH;
    __shared__ float red[GAT_WARPS * RPW][HC];
    float4 dbl = f4zero();
    GAT_ROW_SETUP()
    const float4 a = ldg4(att + ch);
    const float4 bl = lin_bias ? ldg4(lin_bias + ch) : f4zero(), br = lin_bias ? ldg4(lin_bias + HC + ch) : f4zero();
    const size_t copies = g.period > 0 ? (size_t)(g.num_nodes / g.period) : 1;
    GAT_ROW_LOOP(g.rowptr, g.col) {
        GAT_ROW_ADVANCE(g.rowptr, g.col)
        const int ml = row_map(cur, row_to_col, sl);          // -1 beyond the out-edges: the self edge
        const int r = cur.r, cnt = cur.deg + 1;
        const bool valid = cnt > 0;
        const float4 xlj = valid ? f4add(ldg4(xl + (size_t)r * ld + ch), bl) : f4zero();
        const size_t ebase = (size_t)cur.gi * g.num_edges, sidx = (copies * g.num_edges + (size_t)r) * H + head;
        float4 dx = f4zero();
        const int mc = warp_max_cnt<LPR>(cnt);
        if (mc <= GAT_FAST) {             // the edges of a source are independent: batches of four keep the registers low
            bwd_source_fast<LPR, LPH>(xr, dout, alpha, dedge, ld, ch, jl, ml, cnt, 0, ebase, sidx, head, a, br, xlj, dx);
            if (mc > 4) bwd_source_fast<LPR, LPH>(xr, dout, alpha, dedge, ld, ch, jl, ml, cnt, 4, ebase, sidx, head, a, br, xlj, dx);
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int i = k < cur.deg ? __ldg(g.col + cur.beg + k) + cur.off : r;
                const size_t idx = k < cur.deg ? (ebase + (size_t)__ldg(row_to_col + cur.beg + k)) * H + head : sidx;
                const float4 xri = f4add(ldg4(xr + (size_t)i * ld + ch), br);
                const float4 doi = ldg4(dout + (size_t)i * HC + ch);
                bwd_b_edge4(a, xlj, xri, doi, __ldg(alpha + idx), __ldg(dedge + idx), dx);
            }
        }
        if (valid) { st4(dxl + (size_t)r * ld + ch, dx); dbl = f4add(dbl, dx); }
        GAT_ROW_ROTATE()
    }
    {
        const int lane_ = threadIdx.x & 31;
        st4(&red[(threadIdx.x >> 5) * RPW + lane_ / LPR][(lane_ % LPR) * 4], dbl);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < HC; c += GAT_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < GAT_WARPS * RPW; ++w) s += red[w][c];
        partials[(size_t)blockIdx.x * 4 * HC + 3 * HC + c] = s;
    }
}

// Column sums of the per-CTA partials [grid][datt | dbias | sum dx_r | sum dx_l]; one CTA per column, fixed per-thread
// strides and a fixed shared-memory tree => bit-stable.  grad_lin_bias = [sum dx_l | sum dx_r] (lin_l.bias | lin_r.bias).
__global__ void __launch_bounds__(128)
k_gat_reduce(const float* __restrict__ partials, int nblocks, int HC, float* __restrict__ grad_att,
             float* __restrict__ grad_bias, float* __restrict__ grad_lin_bias) {
    __shared__ float red[128];
    const int c = blockIdx.x;
    float s = 0.f;
    for (int b = threadIdx.x; b < nblocks; b += 128) s += partials[(size_t)b * 4 * HC + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (c < HC) grad_att[c] = red[0];
        else if (c < 2 * HC) grad_bias[c - HC] = red[0];
        else if (c < 3 * HC) grad_lin_bias[HC + (c - 2 * HC)] = red[0];
        else grad_lin_bias[c - 3 * HC] = red[0];
    }
}

int gat_vec(int H, int C) {
    const int HC = H * C;
    if (H <= 0 || C <= 0 || HC > 128) return 0;
    for (int v = 1; v <= 4; v *= 2)
        if (HC % v == 0 && C % v == 0 && HC / v <= 32) return v;
    return 0;
}

// 0 = generic kernels, else the specialised <LPR,LPH> variant
int gat_special(int H, int C, int ld) {
    if (ld % 4 != 0) return 0;
    if (H * C == 128 && C == 32) return 1;
    if (H * C == 64 && C == 32) return 2;
    if (H * C == 64 && C == 16) return 3;
    return 0;
}

int gat_check(const s2v_graph_t* g, int H, int C, int ld) {
    if (!g || g->num_nodes < 0 || g->period < 0) return S2V_ERR_BAD_ARG;
    if (g->period == 0 && g->num_nodes > 0 && !g->gid) return S2V_ERR_BAD_ARG;
    if (!g->rowptr || !g->rowptr_t) return S2V_ERR_BAD_ARG;
    const int vec = gat_vec(H, C);
    if (!vec) return S2V_ERR_UNSUPPORTED;
    if (ld < H * C || ld % vec != 0) return S2V_ERR_BAD_ARG;
    return 0;
}

int gat_grid(int num_nodes, int rows_per_cta) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (num_nodes + rows_per_cta - 1) / rows_per_cta;
    int cap = sms * 8;                             // 8 CTAs of 256 threads per SM; a multiple of the SM count
    if (cap > S2V_GAT_MAX_GRID) cap = S2V_GAT_MAX_GRID;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace

#define GAT_DISPATCH(VEC_, CALL)          \
    switch (VEC_) {                       \
        case 4: { constexpr int V = 4; CALL; break; } \
        case 2: { constexpr int V = 2; CALL; break; } \
        default: { constexpr int V = 1; CALL; break; } \
    }
#define GAT_SPECIAL(SP_, CALL)                                            \
    switch (SP_) {                                                        \
        case 1: { constexpr int LPR = 32, LPH = 8; CALL; break; }         \
        case 2: { constexpr int LPR = 16, LPH = 8; CALL; break; }         \
        default: { constexpr int LPR = 16, LPH = 4; CALL; break; }        \
    }

static int64_t gat_edge_floats(const s2v_graph_t* g, int heads) {
    const int64_t copies = g->period > 0 ? g->num_nodes / g->period : 1;
    return (copies * (int64_t)g->num_edges + g->num_nodes) * heads;
}

extern "C" int64_t s2v_gat_edge_floats(const s2v_graph_t* g, int32_t heads) {
    return g && heads > 0 ? gat_edge_floats(g, heads) : 0;
}

extern "C" int s2v_gat_attn_forward(const s2v_graph_t* g, int32_t heads, int32_t channels, int32_t ld,
                                    const float* xl, const float* xr, const float* att, const float* bias,
                                    const float* lin_bias, int32_t apply_relu, float* out, float* stats, float* alpha,
                                    void* stream) {
    S2V_RETURN_IF(gat_check(g, heads, channels, ld));
    if (!xl || !xr || !att || !bias || !out || (!stats && !alpha)) return S2V_ERR_BAD_ARG;
    if (g->num_nodes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int sp = alpha ? gat_special(heads, channels, ld) : 0;
    if (sp) {
        const int grid = gat_grid(g->num_nodes, GAT_WARPS * (sp == 1 ? 1 : 2));
        GAT_SPECIAL(sp, (k_gat_fwd4<LPR, LPH><<<grid, GAT_THREADS, 0, st>>>(*g, ld, xl, xr, att, bias, lin_bias, apply_relu, out, alpha)));
    } else {
        if (!stats) return S2V_ERR_BAD_ARG;
        const int grid = gat_grid(g->num_nodes, GAT_WARPS);
        GAT_DISPATCH(gat_vec(heads, channels),
                     (k_gat_fwd<V><<<grid, GAT_THREADS, 0, st>>>(*g, heads, channels, ld, xl, xr, att, bias, lin_bias, apply_relu, out, stats)));
    }
    return (int)cudaGetLastError();
}

// 1 when the (heads, channels, ld) combination runs the specialised kernels, which keep per-edge attention weights
// (alpha, [s2v_gat_edge_floats]) instead of per-node softmax statistics and need the row_to_col edge map in backward
extern "C" int s2v_gat_uses_edge_alpha(int32_t heads, int32_t channels, int32_t ld) {
    return gat_vec(heads, channels) && gat_special(heads, channels, ld) ? 1 : 0;
}

extern "C" int64_t s2v_gat_attn_backward_workspace_bytes(int32_t heads, int32_t channels, int64_t edge_floats) {
    // per-CTA partials (datt, dbias, sum dx_r, sum dx_l) + de per edge and head (specialised kernels)
    return (int64_t)sizeof(float) * ((int64_t)S2V_GAT_MAX_GRID * 4 * heads * channels + (edge_floats > 0 ? edge_floats : 0));
}

extern "C" int s2v_gat_attn_backward(const s2v_graph_t* g, int32_t heads, int32_t channels, int32_t ld,
                                     const float* xl, const float* xr, const float* att, const float* lin_bias,
                                     float* stats, const float* alpha, const int32_t* row_to_col, const float* dout,
                                     float* dxl, float* dxr, float* grad_att, float* grad_bias,
                                     float* grad_lin_bias, void* workspace, int64_t workspace_bytes, void* stream) {
    S2V_RETURN_IF(gat_check(g, heads, channels, ld));
    if (!xl || !xr || !att || (!stats && !alpha) || !dout || !dxl || !dxr || !grad_att || !grad_bias || !grad_lin_bias || !workspace) return S2V_ERR_BAD_ARG;
    const int sp = alpha ? gat_special(heads, channels, ld) : 0;
    if (sp && g->num_edges > 0 && !row_to_col) return S2V_ERR_BAD_ARG;
    if (!sp && !stats) return S2V_ERR_BAD_ARG;
    if (workspace_bytes < s2v_gat_attn_backward_workspace_bytes(heads, channels, sp ? gat_edge_floats(g, heads) : 0)) return S2V_ERR_WORKSPACE;
    const int HC = heads * channels;
    cudaStream_t st = (cudaStream_t)stream;
    if (g->num_nodes == 0) {
        cudaMemsetAsync(grad_att, 0, sizeof(float) * HC, st);
        cudaMemsetAsync(grad_bias, 0, sizeof(float) * HC, st);
        cudaMemsetAsync(grad_lin_bias, 0, sizeof(float) * 2 * HC, st);
        return (int)cudaGetLastError();
    }
    float* partials = (float*)workspace;
    float* dedge = partials + (size_t)S2V_GAT_MAX_GRID * 4 * HC;
    int grid;
    if (sp) {
        grid = gat_grid(g->num_nodes, GAT_WARPS * (sp == 1 ? 1 : 2));
        GAT_SPECIAL(sp, (k_gat_bwd_target4<LPR, LPH><<<grid, GAT_THREADS, 0, st>>>(*g, ld, xl, xr, att, lin_bias, alpha, dedge, dout, dxr, partials)));
        S2V_RETURN_IF(cudaGetLastError());
        GAT_SPECIAL(sp, (k_gat_bwd_source4<LPR, LPH><<<grid, GAT_THREADS, 0, st>>>(*g, ld, xl, xr, att, lin_bias, alpha, dedge, row_to_col, dout, dxl, partials)));
    } else {
        grid = gat_grid(g->num_nodes, GAT_WARPS);
        GAT_DISPATCH(gat_vec(heads, channels),
                     (k_gat_bwd_target<V><<<grid, GAT_THREADS, 0, st>>>(*g, heads, channels, ld, xl, xr, att, lin_bias, stats, dout, dxr, partials)));
        S2V_RETURN_IF(cudaGetLastError());
        GAT_DISPATCH(gat_vec(heads, channels),
                     (k_gat_bwd_source<V><<<grid, GAT_THREADS, 0, st>>>(*g, heads, channels, ld, xl, xr, att, lin_bias, stats, dout, dxl, partials)));
    }
    S2V_RETURN_IF(cudaGetLastError());
    k_gat_reduce<<<4 * HC, 128, 0, st>>>(partials, grid, HC, grad_att, grad_bias, grad_lin_bias);
    return (int)cudaGetLastError();
}
