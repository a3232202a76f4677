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

__device__ __forceinline__ float lrelu(float s) { return s > 0.f ? s : s * GAT_SLOPE; }

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
          const float* __restrict__ att, const float* __restrict__ bias, int apply_relu,
          float* __restrict__ out, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int HC = H * C, lanes = HC / VEC, lph = C / VEC;
    const bool active = lane < lanes;
    const int ch = active ? lane * VEC : 0;
    const int head = ch / C;
    const Vec<VEC> a = ldv<VEC>(att + ch), b = ldv<VEC>(bias + ch);
    const int warp = blockIdx.x * GAT_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GAT_WARPS;
    for (int r = warp; r < g.num_nodes; r += nwarps) {
        int li, off, gi;
        row_info(g, r, li, off, gi);
        const int beg = __ldg(g.rowptr_t + li), deg = __ldg(g.rowptr_t + li + 1) - beg, cnt = deg + 1;
        const Vec<VEC> xri = ldv<VEC>(xr + (size_t)r * ld + ch);
        Vec<VEC> acc = vzero<VEC>();
        float m = -INFINITY, den;
        if (cnt <= GAT_FAST) {
            const int jl = lane < deg ? __ldg(g.col_t + beg + lane) + off : r;
            int js[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) js[k] = __shfl_sync(0xffffffffu, jl, k);
            Vec<VEC> rows[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) rows[k] = k < cnt ? ldv<VEC>(xl + (size_t)js[k] * ld + ch) : vzero<VEC>();
            float sc[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) sc[k] = head_sum(active ? score_partial<VEC>(a, xri, rows[k]) : 0.f, lph, lane);
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) m = fmaxf(m, sc[k]);
            float S = 0.f;
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) { sc[k] = expf(sc[k] - m); S += sc[k]; }
            den = S + 1e-16f;
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) {
                const float al = sc[k] / den;
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc.v[v] = fmaf(rows[k].v[v], al, acc.v[v]);
            }
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = ldv<VEC>(xl + (size_t)j * ld + ch);
                m = fmaxf(m, head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane));
            }
            float S = 0.f;
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = ldv<VEC>(xl + (size_t)j * ld + ch);
                S += expf(head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane) - m);
            }
            den = S + 1e-16f;
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = ldv<VEC>(xl + (size_t)j * ld + ch);
                const float al = expf(head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane) - m) / den;
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc.v[v] = fmaf(row.v[v], al, acc.v[v]);
            }
        }
        if (active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { float o = acc.v[v] + b.v[v]; acc.v[v] = apply_relu ? relu(o) : o; }
            stv<VEC>(out + (size_t)r * HC + ch, acc);
            if (lane % lph == 0) { stats[(size_t)r * 2 * H + 2 * head] = m; stats[(size_t)r * 2 * H + 2 * head + 1] = den; }
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
                 const float* __restrict__ att, const float* __restrict__ stats, const float* __restrict__ dout,
                 float* __restrict__ dxr, float* __restrict__ dsum, float* __restrict__ partials) {
    __shared__ float red[GAT_WARPS][2 * 128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int HC = H * C, lanes = HC / VEC, lph = C / VEC;
    const bool active = lane < lanes;
    const int ch = active ? lane * VEC : 0;
    const int head = ch / C;
    const Vec<VEC> a = ldv<VEC>(att + ch);
    Vec<VEC> datt = vzero<VEC>(), dbias = vzero<VEC>();
    const int warp = blockIdx.x * GAT_WARPS + wid, nwarps = gridDim.x * GAT_WARPS;
    for (int r = warp; r < g.num_nodes; r += nwarps) {
        int li, off, gi;
        row_info(g, r, li, off, gi);
        const int beg = __ldg(g.rowptr_t + li), deg = __ldg(g.rowptr_t + li + 1) - beg, cnt = deg + 1;
        const Vec<VEC> xri = ldv<VEC>(xr + (size_t)r * ld + ch);
        const Vec<VEC> doi = ldv<VEC>(dout + (size_t)r * HC + ch);
        const float m = __ldg(stats + (size_t)r * 2 * H + 2 * head), den = __ldg(stats + (size_t)r * 2 * H + 2 * head + 1);
        Vec<VEC> dx = vzero<VEC>();
        float D = 0.f;
        if (cnt <= GAT_FAST) {
            const int jl = lane < deg ? __ldg(g.col_t + beg + lane) + off : r;
            int js[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) js[k] = __shfl_sync(0xffffffffu, jl, k);
            Vec<VEC> rows[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) rows[k] = k < cnt ? ldv<VEC>(xl + (size_t)js[k] * ld + ch) : vzero<VEC>();
            float al[GAT_FAST], dal[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) {
                const float e = head_sum(active ? score_partial<VEC>(a, xri, rows[k]) : 0.f, lph, lane);
                dal[k] = head_sum(active ? dot_partial<VEC>(doi, rows[k]) : 0.f, lph, lane);
                al[k] = k < cnt ? expf(e - m) / den : 0.f;
            }
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) D = fmaf(al[k], dal[k], D);
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) if (k < cnt) bwd_a_edge<VEC>(a, xri, rows[k], al[k], dal[k], D, dx, datt);
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = ldv<VEC>(xl + (size_t)j * ld + ch);
                const float e = head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(doi, row) : 0.f, lph, lane);
                D = fmaf(expf(e - m) / den, dl, D);
            }
            for (int k = 0; k < cnt; ++k) {
                const int j = k < deg ? __ldg(g.col_t + beg + k) + off : r;
                const Vec<VEC> row = ldv<VEC>(xl + (size_t)j * ld + ch);
                const float e = head_sum(active ? score_partial<VEC>(a, xri, row) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(doi, row) : 0.f, lph, lane);
                bwd_a_edge<VEC>(a, xri, row, expf(e - m) / den, dl, D, dx, datt);
            }
        }
        if (active) {
            stv<VEC>(dxr + (size_t)r * ld + ch, dx);
            if (lane % lph == 0) dsum[(size_t)r * H + head] = D;
#pragma unroll
            for (int v = 0; v < VEC; ++v) dbias.v[v] += doi.v[v];
        }
    }
    // per-CTA partial sums, fixed order over warps
    if (active) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { red[wid][ch + v] = datt.v[v]; red[wid][128 + ch + v] = dbias.v[v]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * HC; c += GAT_THREADS) {
        const int idx = c < HC ? c : 128 + (c - HC);
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < GAT_WARPS; ++w) s += red[w][idx];
        partials[(size_t)blockIdx.x * 2 * HC + c] = s;
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
                 const float* __restrict__ att, const float* __restrict__ stats, const float* __restrict__ dout,
                 const float* __restrict__ dsum, float* __restrict__ dxl) {
    const int lane = threadIdx.x & 31;
    const int HC = H * C, lanes = HC / VEC, lph = C / VEC;
    const bool active = lane < lanes;
    const int ch = active ? lane * VEC : 0;
    const int head = ch / C;
    const Vec<VEC> a = ldv<VEC>(att + ch);
    const int warp = blockIdx.x * GAT_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GAT_WARPS;
    for (int r = warp; r < g.num_nodes; r += nwarps) {
        int li, off, gi;
        row_info(g, r, li, off, gi);
        const int beg = __ldg(g.rowptr + li), deg = __ldg(g.rowptr + li + 1) - beg, cnt = deg + 1;
        const Vec<VEC> xlj = ldv<VEC>(xl + (size_t)r * ld + ch);
        Vec<VEC> dx = vzero<VEC>();
        if (cnt <= GAT_FAST) {
            const int il = lane < deg ? __ldg(g.col + beg + lane) + off : r;
            int is[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) is[k] = __shfl_sync(0xffffffffu, il, k);
            Vec<VEC> xrs[GAT_FAST], dos[GAT_FAST];
            float ms[GAT_FAST], dens[GAT_FAST], Ds[GAT_FAST];
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) {
                const bool on = k < cnt;
                xrs[k] = on ? ldv<VEC>(xr + (size_t)is[k] * ld + ch) : vzero<VEC>();
                dos[k] = on ? ldv<VEC>(dout + (size_t)is[k] * HC + ch) : vzero<VEC>();
                ms[k] = on ? __ldg(stats + (size_t)is[k] * 2 * H + 2 * head) : 0.f;
                dens[k] = on ? __ldg(stats + (size_t)is[k] * 2 * H + 2 * head + 1) : 1.f;
                Ds[k] = on ? __ldg(dsum + (size_t)is[k] * H + head) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < GAT_FAST; ++k) {
                const float e = head_sum(active ? score_partial<VEC>(a, xrs[k], xlj) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(dos[k], xlj) : 0.f, lph, lane);
                if (k < cnt) bwd_b_edge<VEC>(a, xlj, xrs[k], dos[k], expf(e - ms[k]) / dens[k], dl, Ds[k], dx);
            }
        } else {
            for (int k = 0; k < cnt; ++k) {
                const int i = k < deg ? __ldg(g.col + beg + k) + off : r;
                const Vec<VEC> xri = ldv<VEC>(xr + (size_t)i * ld + ch);
                const Vec<VEC> doi = ldv<VEC>(dout + (size_t)i * HC + ch);
                const float m = __ldg(stats + (size_t)i * 2 * H + 2 * head), den = __ldg(stats + (size_t)i * 2 * H + 2 * head + 1);
                const float D = __ldg(dsum + (size_t)i * H + head);
                const float e = head_sum(active ? score_partial<VEC>(a, xri, xlj) : 0.f, lph, lane);
                const float dl = head_sum(active ? dot_partial<VEC>(doi, xlj) : 0.f, lph, lane);
                bwd_b_edge<VEC>(a, xlj, xri, doi, expf(e - m) / den, dl, D, dx);
            }
        }
        if (active) stv<VEC>(dxl + (size_t)r * ld + ch, dx);
    }
}

// grad_att[c] / grad_bias[c] = sum over CTAs of partials[b][c] / partials[b][HC + c], fixed order
__global__ void k_gat_reduce(const float* __restrict__ partials, int nblocks, int HC, float* __restrict__ grad_att,
                             float* __restrict__ grad_bias) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * HC) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * 2 * HC + c];
    if (c < HC) grad_att[c] = s; else grad_bias[c - HC] = s;
}

int gat_vec(int H, int C) {
    const int HC = H * C;
    if (H <= 0 || C <= 0 || HC > 128) return 0;
    for (int v = 1; v <= 4; v *= 2)
        if (HC % v == 0 && C % v == 0 && HC / v <= 32) return v;
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

int gat_grid(int num_nodes) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (num_nodes + GAT_WARPS - 1) / GAT_WARPS;
    const int cap = sms * 8;                       // 8 CTAs of 256 threads per SM = full occupancy; multiple of the SM count
    return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace

#define GAT_DISPATCH(VEC_, CALL)          \
    switch (VEC_) {                       \
        case 4: { constexpr int V = 4; CALL; break; } \
        case 2: { constexpr int V = 2; CALL; break; } \
        default: { constexpr int V = 1; CALL; break; } \
    }

extern "C" int s2v_gat_attn_forward(const s2v_graph_t* g, int32_t heads, int32_t channels, int32_t ld,
                                    const float* xl, const float* xr, const float* att, const float* bias,
                                    int32_t apply_relu, float* out, float* stats, void* stream) {
    S2V_RETURN_IF(gat_check(g, heads, channels, ld));
    if (!xl || !xr || !att || !bias || !out || !stats) return S2V_ERR_BAD_ARG;
    if (g->num_nodes == 0) return 0;
    const int grid = gat_grid(g->num_nodes);
    GAT_DISPATCH(gat_vec(heads, channels),
                 (k_gat_fwd<V><<<grid, GAT_THREADS, 0, (cudaStream_t)stream>>>(*g, heads, channels, ld, xl, xr, att, bias,
                                                                                apply_relu, out, stats)));
    return (int)cudaGetLastError();
}

extern "C" int64_t s2v_gat_attn_backward_workspace_bytes(int32_t heads, int32_t channels, int64_t num_nodes) {
    // per-CTA partials of (datt, dbias) for the largest grid, + D_i per node and head
    return (int64_t)sizeof(float) * ((int64_t)S2V_GAT_MAX_GRID * 2 * heads * channels + num_nodes * heads);
}

extern "C" int s2v_gat_attn_backward(const s2v_graph_t* g, int32_t heads, int32_t channels, int32_t ld,
                                     const float* xl, const float* xr, const float* att, const float* stats,
                                     const float* dout, float* dxl, float* dxr, float* grad_att, float* grad_bias,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
    S2V_RETURN_IF(gat_check(g, heads, channels, ld));
    if (!xl || !xr || !att || !stats || !dout || !dxl || !dxr || !grad_att || !grad_bias || !workspace) return S2V_ERR_BAD_ARG;
    if (workspace_bytes < s2v_gat_attn_backward_workspace_bytes(heads, channels, g->num_nodes)) return S2V_ERR_WORKSPACE;
    const int HC = heads * channels;
    cudaStream_t st = (cudaStream_t)stream;
    if (g->num_nodes == 0) {
        cudaMemsetAsync(grad_att, 0, sizeof(float) * HC, st);
        cudaMemsetAsync(grad_bias, 0, sizeof(float) * HC, st);
        return (int)cudaGetLastError();
    }
    int grid = gat_grid(g->num_nodes);
    if (grid > S2V_GAT_MAX_GRID) grid = S2V_GAT_MAX_GRID;
    float* partials = (float*)workspace;
    float* dsum = partials + (size_t)S2V_GAT_MAX_GRID * 2 * HC;
    GAT_DISPATCH(gat_vec(heads, channels),
                 (k_gat_bwd_target<V><<<grid, GAT_THREADS, 0, st>>>(*g, heads, channels, ld, xl, xr, att, stats, dout, dxr, dsum, partials)));
    S2V_RETURN_IF(cudaGetLastError());
    GAT_DISPATCH(gat_vec(heads, channels),
                 (k_gat_bwd_source<V><<<grid, GAT_THREADS, 0, st>>>(*g, heads, channels, ld, xl, xr, att, stats, dout, dsum, dxl)));
    S2V_RETURN_IF(cudaGetLastError());
    // partials are [grid][datt HC | dbias HC]
    k_gat_reduce<<<(2 * HC + 127) / 128, 128, 0, st>>>(partials, grid, HC, grad_att, grad_bias);
    return (int)cudaGetLastError();
}
