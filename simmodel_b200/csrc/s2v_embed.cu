// Embedding stage of struct2vec: node MLP, degree term, T propagation rounds, and
// the matching backward.  Math: SURVEY.md Appendix A (derived from
// modules/gnn/comb_opt.py:230-242 and modules/ppo/models_basic_lstm.py:491-503).
//
//   h_i    = relu(theta1a x_i + b1a)            s1_i = theta1b h_i + b1b
//   r1     = relu(w4 + b4), r0 = relu(b4)       z_j  = c_j r1 + (Npad - c_j) r0
//   s3_j   = theta3 z_j + b3 = c_j u1 + (Npad - c_j) u0 + b3,   u1 = theta3 r1, u0 = theta3 r0
//   base_i = s1_i + s3_i + b2
//   mu^1   = relu(base)                          (mu^0 = 0, so the first gather is zero)
//   mu^t_i = relu(base_i + theta2 sum_{j in out(i)} mu^{t-1}_j)
//
// FP32 CUDA-core path (FFMA); see s2v_common.cuh for the tile model.
#include "s2v_common.cuh"
#include "s2v_bodies.cuh"

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int P, int TM = S2V_TM>
__global__ void __launch_bounds__(S2V_THREADS)
k_embed_first(s2v_graph_t g, s2v_embed_params_t prm, const float* __restrict__ x,
              float* __restrict__ base, float* __restrict__ mu1, int num_tiles) {
    extern __shared__ __align__(16) float smem[];
    embed_first_body<P, TM>(smem, g, prm, x, base, mu1, num_tiles);
}

template <int P, int TM = S2V_TM>
__global__ void __launch_bounds__(S2V_THREADS)
k_embed_round(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ base,
              const float* __restrict__ mu_prev, float* __restrict__ mu_next, int num_tiles) {
    extern __shared__ __align__(16) float smem[];
    embed_round_body<P, 0, TM>(smem, g, theta2_w, base, mu_prev, mu_next, num_tiles);
}

// ------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------
// Per-CTA partial sums (floats), reduced in fixed order by k_embed_finalize.
template <int P>
struct EmbedPartial {
    static constexpr int TH2 = 0;                       // P*P   sum_j g_j mu_j^T          (all rounds)
    static constexpr int TH1B = P * P;                  // P*P   sum_i D_i h_i^T
    static constexpr int COLD = 2 * P * P;              // P     sum_i D_i
    static constexpr int A1 = COLD + P;                 // P     sum_i c_i D_i
    static constexpr int A0 = A1 + P;                   // P     sum_i (Npad - c_i) D_i
    static constexpr int B1A = A0 + P;                  // P     sum_i dh_i
    static constexpr int TH1A = B1A + P;                // P*F   sum_i dh_i x_i^T
    static __host__ __device__ int size(int F) { return TH1A + P * F; }
};

// One backward round: rows j of the tile.
//   g_j = sum_{i : j in out(i)} du^t_i           (gather on the transposed CSR)
//   dtheta2 += g_j mu^{t-1}_j^T ;  du^{t-1}_j = (theta2^T g_j) * [mu^{t-1}_j > 0]
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_embed_round_bwd(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ du_t,
                  const float* __restrict__ mu_prev, float* __restrict__ du_prev, float* __restrict__ partial,
                  int partial_stride, int accumulate, int num_tiles) {
    using C = Cfg<P>;
    extern __shared__ __align__(16) float smem[];
    float* W2 = smem;                         // theta2 as stored: [n][k]
    float* As = W2 + C::W_FLOATS;             // g tile
    float* Bs = As + C::TILE_FLOATS;          // mu^{t-1} tile
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    load_w_plain<P>(W2, theta2_w);
    float racc[C::NRR][4];
#pragma unroll
    for (int i = 0; i < C::NRR; ++i) racc[i][0] = racc[i][1] = racc[i][2] = racc[i][3] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * S2V_TM;
        const int rows = min(S2V_TM, g.num_nodes - row0);
        gather_tile<P>(As, g, g.rowptr_t, g.col_t, du_t, row0, rows);
        load_tile<P>(Bs, mu_prev + (size_t)row0 * P, rows);
        __syncthreads();
        if (active) {
            tile_reduce_gemm<P>(As, Bs, rows, racc, tx, ty);
            float acc[C::NR][4];
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P>(As, W2, acc, tx, ty);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= rows) continue;
                float4 m = ld4(Bs + rl * C::LDA + 4 * tx);
                st4(du_prev + (size_t)(row0 + rl) * P + 4 * tx,
                    make_float4(m.x > 0.f ? acc[i][0] : 0.f, m.y > 0.f ? acc[i][1] : 0.f,
                                m.z > 0.f ? acc[i][2] : 0.f, m.w > 0.f ? acc[i][3] : 0.f));
            }
        }
        __syncthreads();
    }
    store_racc<P>(partial + (size_t)blockIdx.x * partial_stride + EmbedPartial<P>::TH2, racc, tx, ty, active, accumulate != 0);
}

// Last backward stage: D_i = sum_t du^t_i and everything that depends on it.
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_embed_final_bwd(s2v_graph_t g, s2v_embed_params_t prm, const float* __restrict__ x,
                  const float* __restrict__ dus, const float* __restrict__ du_last, int T,
                  float* __restrict__ partial, int partial_stride, int num_tiles, float* __restrict__ dx) {
    using C = Cfg<P>;
    using EP = EmbedPartial<P>;
    extern __shared__ __align__(16) float smem[];
    const int F = prm.node_dim;
    float* W1b = smem;                        // theta1b as stored [n][k]
    float* As = W1b + C::W_FLOATS;            // D tile, later dh tile
    float* Bs = As + C::TILE_FLOATS;          // h tile
    float* W1a = Bs + C::TILE_FLOATS;         // [F][P]
    float* xs = W1a + S2V_MAX_F * P;          // [TM][F]
    float* b1a = xs + S2V_TM * S2V_MAX_F;     // [P]
    float* scratch = b1a + P;                 // [RG][P]
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    const size_t plane = (size_t)g.num_nodes * P;

    load_w_plain<P>(W1b, prm.theta1b_w);
    for (int e = t; e < P * F; e += S2V_THREADS) { int n = e / F, f = e - n * F; W1a[f * P + n] = __ldg(prm.theta1a_w + e); }
    if (t < P) b1a[t] = __ldg(prm.theta1a_b + t);

    float racc[C::NRR][4];
#pragma unroll
    for (int i = 0; i < C::NRR; ++i) racc[i][0] = racc[i][1] = racc[i][2] = racc[i][3] = 0.f;
    float pD[4] = {0.f, 0.f, 0.f, 0.f}, pA1[4] = {0.f, 0.f, 0.f, 0.f}, pA0[4] = {0.f, 0.f, 0.f, 0.f}, pB1a[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int MAXE = (P * S2V_MAX_F + S2V_THREADS - 1) / S2V_THREADS;
    float pth1a[MAXE];
#pragma unroll
    for (int i = 0; i < MAXE; ++i) pth1a[i] = 0.f;
    __syncthreads();

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * S2V_TM;
        const int rows = min(S2V_TM, g.num_nodes - row0);
        // D tile: fixed summation order t = 1..T
        constexpr int V = P / 4;
        for (int e = t; e < S2V_TM * V; e += S2V_THREADS) {
            int r = e / V, c = e - r * V;
            float4 s = f4zero();
            if (r < rows) {
                size_t o = (size_t)(row0 + r) * P + 4 * c;
                for (int tt = 0; tt < T - 1; ++tt) s = f4add(s, ldg4(dus + tt * plane + o));
                s = f4add(s, ldg4(du_last + o));
            }
            st4(As + r * C::LDA + 4 * c, s);
        }
        for (int e = t; e < S2V_TM * F; e += S2V_THREADS)
            xs[e] = e < rows * F ? __ldg(x + (size_t)row0 * F + e) : 0.f;
        __syncthreads();
        if (active) {
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int r = ty + C::RG * i;
                if (r >= S2V_TM) continue;
                float4 h = ld4(b1a + 4 * tx);
                for (int f = 0; f < F; ++f) {
                    float xv = xs[r * F + f];
                    float4 w = ld4(W1a + f * P + 4 * tx);
                    h.x = fmaf(xv, w.x, h.x); h.y = fmaf(xv, w.y, h.y); h.z = fmaf(xv, w.z, h.z); h.w = fmaf(xv, w.w, h.w);
                }
                // rows past the end must not contribute to dtheta1b: their D is zero, h value is irrelevant
                st4(Bs + r * C::LDA + 4 * tx, make_float4(relu(h.x), relu(h.y), relu(h.z), relu(h.w)));
                if (r < rows) {
                    int rr = row0 + r, li, off, gi;
                    row_info(g, rr, li, off, gi);
                    float c = (float)__ldg(g.indeg + li);
                    float d = (float)graph_npad(g, gi) - c;
                    float4 D = ld4(As + r * C::LDA + 4 * tx);
                    pD[0] += D.x; pD[1] += D.y; pD[2] += D.z; pD[3] += D.w;
                    pA1[0] = fmaf(c, D.x, pA1[0]); pA1[1] = fmaf(c, D.y, pA1[1]); pA1[2] = fmaf(c, D.z, pA1[2]); pA1[3] = fmaf(c, D.w, pA1[3]);
                    pA0[0] = fmaf(d, D.x, pA0[0]); pA0[1] = fmaf(d, D.y, pA0[1]); pA0[2] = fmaf(d, D.z, pA0[2]); pA0[3] = fmaf(d, D.w, pA0[3]);
                }
            }
        }
        __syncthreads();
        float acc[C::NR][4];
        if (active) {
            tile_reduce_gemm<P>(As, Bs, rows, racc, tx, ty);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P>(As, W1b, acc, tx, ty);
        }
        __syncthreads();                       // everyone is done reading D from As
        if (active) {
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= S2V_TM) continue;
                float4 h = ld4(Bs + rl * C::LDA + 4 * tx);
                float4 dh = make_float4(h.x > 0.f ? acc[i][0] : 0.f, h.y > 0.f ? acc[i][1] : 0.f,
                                        h.z > 0.f ? acc[i][2] : 0.f, h.w > 0.f ? acc[i][3] : 0.f);
                if (rl >= rows) dh = f4zero();
                pB1a[0] += dh.x; pB1a[1] += dh.y; pB1a[2] += dh.z; pB1a[3] += dh.w;
                st4(As + rl * C::LDA + 4 * tx, dh);
            }
        }
        __syncthreads();
        if (dx != nullptr) {                   // gradient w.r.t. the node features (lstm_type 'FE': an LSTM feeds x): dx = dh theta1a
            for (int e = t; e < rows * F; e += S2V_THREADS) {
                const int r = e / F, f = e - r * F;
                float s = 0.f;
                for (int k = 0; k < P; ++k) s = fmaf(As[r * C::LDA + k], W1a[f * P + k], s);
                dx[(size_t)row0 * F + e] = s;
            }
        }
#pragma unroll
        for (int i = 0; i < MAXE; ++i) {
            int e = t + i * S2V_THREADS;
            if (e < P * F) {
                int k = e / F, f = e - k * F;
                float s = pth1a[i];
                for (int r = 0; r < rows; ++r) s = fmaf(As[r * C::LDA + k], xs[r * F + f], s);
                pth1a[i] = s;
            }
        }
        __syncthreads();
    }
    float* my = partial + (size_t)blockIdx.x * partial_stride;
    store_racc<P>(my + EP::TH1B, racc, tx, ty, active, false);
    reduce_columns<P>(scratch, pD, my + EP::COLD, tx, ty, active, false);
    reduce_columns<P>(scratch, pA1, my + EP::A1, tx, ty, active, false);
    reduce_columns<P>(scratch, pA0, my + EP::A0, tx, ty, active, false);
    reduce_columns<P>(scratch, pB1a, my + EP::B1A, tx, ty, active, false);
#pragma unroll
    for (int i = 0; i < MAXE; ++i) {
        int e = t + i * S2V_THREADS;
        if (e < P * F) my[EP::TH1A + e] = pth1a[i];
    }
}

// Closed-form theta3/theta4 gradients and copies, from the fixed-order sums of the per-CTA
// partials (k_reduce_partials, s2v_common.cuh).  sums has the EmbedPartial layout.
// grad layout: theta1a.w [P,F], theta1a.b, theta1b.w [P,P], theta1b.b, theta2.w, theta2.b,
//              theta3.w, theta3.b, theta4.w [P], theta4.b [P]
template <int P>
__global__ void __launch_bounds__(256)
k_embed_finalize(s2v_embed_params_t prm, const float* __restrict__ sums, int have_th2, float* __restrict__ grad) {
    using EP = EmbedPartial<P>;
    const int F = prm.node_dim;
    __shared__ float A1[P], A0[P], r1[P], r0[P];
    const int t = threadIdx.x;
    float* g1a_w = grad;
    float* g1a_b = g1a_w + P * F;
    float* g1b_w = g1a_b + P;
    float* g1b_b = g1b_w + P * P;
    float* g2_w = g1b_b + P;
    float* g2_b = g2_w + P * P;
    float* g3_w = g2_b + P;
    float* g3_b = g3_w + P * P;
    float* g4_w = g3_b + P;
    float* g4_b = g4_w + P;
    for (int e = t; e < P; e += blockDim.x) {
        A1[e] = sums[EP::A1 + e];
        A0[e] = sums[EP::A0 + e];
        float w4 = prm.theta4_w[e], b4 = prm.theta4_b[e];
        r1[e] = relu(w4 + b4);
        r0[e] = relu(b4);
        float cd = sums[EP::COLD + e];
        g1a_b[e] = sums[EP::B1A + e];
        g1b_b[e] = cd;
        g2_b[e] = cd;
        g3_b[e] = cd;
    }
    __syncthreads();
    for (int e = t; e < P; e += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int n = 0; n < P; ++n) {
            float w = prm.theta3_w[n * P + e];
            a = fmaf(w, A1[n], a);
            b = fmaf(w, A0[n], b);
        }
        float w4 = prm.theta4_w[e], b4 = prm.theta4_b[e];
        float m1 = (w4 + b4) > 0.f ? a : 0.f;
        g4_w[e] = m1;
        g4_b[e] = m1 + (b4 > 0.f ? b : 0.f);
    }
    for (int e = t; e < P * P; e += blockDim.x) {
        int n = e / P, k = e - n * P;
        g3_w[e] = A1[n] * r1[k] + A0[n] * r0[k];
        g2_w[e] = have_th2 ? sums[EP::TH2 + e] : 0.f;
        g1b_w[e] = sums[EP::TH1B + e];
    }
    for (int e = t; e < P * F; e += blockDim.x) g1a_w[e] = sums[EP::TH1A + e];
}

__global__ void k_relu_mask(const float4* __restrict__ dmu, const float4* __restrict__ mu, float4* __restrict__ du, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 d = dmu[i], m = mu[i];
        du[i] = make_float4(m.x > 0.f ? d.x : 0.f, m.y > 0.f ? d.y : 0.f, m.z > 0.f ? d.z : 0.f, m.w > 0.f ? d.w : 0.f);
    }
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------

template <typename K>
static int set_smem(K kernel, int bytes) {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// tensor-core round kernels (s2v_embed_tc.cu), p = 64 only
int s2v_tc_round_forward(const s2v_graph_t& g, const float* theta2_w, const float* base, const float* mu_prev,
                         float* mu_next, cudaStream_t st);
int s2v_tc_round_backward_grid(int num_nodes);
int s2v_tc_round_backward_ws_grid(int num_nodes);
int s2v_tc_round_backward_ws(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                             float* du_prev, float* partial, int partial_stride, int accumulate, int grid, cudaStream_t st);
int s2v_tc_round_backward(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                          float* du_prev, float* partial, int partial_stride, int accumulate, int grid, cudaStream_t st);
// the tensor path's default backward round: warp-specialised or forward-shaped (S2V_B200_BWD_ROUND, s2v_embed_tc.cu)
int s2v_tc_round_backward_sel_grid(int num_nodes, int flags);
int s2v_tc_round_backward_sel(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                              float* du_prev, float* partial, int partial_stride, int accumulate, int grid, int flags,
                              cudaStream_t st);
bool s2v_tc_final_backward_ws_ok(int node_dim);
int s2v_tc_final_backward_ws(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, const float* dus,
                             const float* du_last, float* partial, int partial_stride, int grid, cudaStream_t st);

bool s2v_tc_final_backward_h2_selected(int flags, int node_dim);
int s2v_tc_final_backward_h2_grid(int num_nodes);
int s2v_tc_final_backward_h2(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, const float* dus,
                             const float* du_last, float* partial, int partial_stride, int grid, cudaStream_t st);

int s2v_tc_embed_first(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, float* base, float* mu1,
                       cudaStream_t st);
int s2v_tc_final_backward_grid(int num_nodes);
int s2v_tc_final_backward(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x, const float* dus,
                          const float* du_last, float* partial, int partial_stride, int grid, cudaStream_t st);

// flags (s2v_embed_params_t.flags): 0 = auto, 1 = CUDA-core FP32 path, 2 = tensor-core path with the 8-warp
// backward round, 3 = tensor-core path with the warp-specialised backward round (what auto picks for large batches)
static bool use_tensor_path(int P, int flags, int num_nodes) {
    if (P != 64 || flags == S2V_PATH_FFMA) return false;
    if (flags >= S2V_PATH_TENSOR && flags <= S2V_PATH_TENSOR_16) return true;
    return num_nodes >= S2V_TENSOR_MIN_NODES;
}
// Backward round.  Three implementations of the same tile math (profiles/ holds the timings, AMS, 4.0 M rows):
//   CUDA-core FP32 (k_embed_round_bwd)                      1.73 ms   (issue / shared-memory bound)
//   tcgen05 3xTF32, 8-warp CTAs (k_embed_round_bwd_tc256)    1.76 ms   (latency-starved: 2 CTAs per SM)
//   tcgen05 bf16x3, warp-specialised (k_embed_round_bwd_ws)  1.39 ms   <- auto, whenever the tensor path is on
// S2V_PATH_TENSOR forces the 8-warp kernel, S2V_PATH_TENSOR_WS the warp-specialised one.
static bool use_tensor_path_bwd_round(int P, int flags, int num_nodes) {
    return use_tensor_path(P, flags, num_nodes);
}
static bool use_ws_bwd_round(int flags) { return flags != S2V_PATH_TENSOR; }

template <int P>
static int embed_forward_impl(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x,
                              float* base, float* mus, cudaStream_t st) {
    using C = Cfg<P>;
    const int tiles = (g.num_nodes + S2V_TM - 1) / S2V_TM;
    const size_t plane = (size_t)g.num_nodes * P;
    const bool tensor = use_tensor_path(P, prm.flags, g.num_nodes);
    // small problems (emb_dim 64, at most 4096 rows): 16-row tiles -- stage latency, not throughput, is the cost there
    constexpr int TS = 16;
    if constexpr (P == 64) if (!tensor && g.num_nodes <= S2V_TENSOR_MIN_NODES) {
        using CS = Cfg<P, TS>;
        const int tiles_s = (g.num_nodes + TS - 1) / TS;
        int smem = (CS::W_FLOATS + CS::TILE_FLOATS + S2V_MAX_F * P + TS * S2V_MAX_F + 6 * P) * 4;
        S2V_RETURN_IF(set_smem(k_embed_first<P, TS>, smem));
        int grid = s2v_persistent_grid((const void*)k_embed_first<P, TS>, smem, tiles_s);
        k_embed_first<P, TS><<<grid, S2V_THREADS, smem, st>>>(g, prm, x, base, mus, tiles_s);
        smem = (CS::W_FLOATS + CS::TILE_FLOATS) * 4;
        S2V_RETURN_IF(set_smem(k_embed_round<P, TS>, smem));
        grid = s2v_persistent_grid((const void*)k_embed_round<P, TS>, smem, tiles_s);
        for (int t = 1; t < prm.rounds; ++t)
            k_embed_round<P, TS><<<grid, S2V_THREADS, smem, st>>>(g, prm.theta2_w, base, mus + (t - 1) * plane, mus + t * plane, tiles_s);
        return (int)cudaGetLastError();
    }
    if (tensor) {
        S2V_RETURN_IF(s2v_tc_embed_first(g, prm, x, base, mus, st));
    } else {
        int smem = (C::W_FLOATS + C::TILE_FLOATS + S2V_MAX_F * P + S2V_TM * S2V_MAX_F + 6 * P) * 4;
        S2V_RETURN_IF(set_smem(k_embed_first<P>, smem));
        int grid = s2v_persistent_grid((const void*)k_embed_first<P>, smem, tiles);
        k_embed_first<P><<<grid, S2V_THREADS, smem, st>>>(g, prm, x, base, mus, tiles);
    }
    if (tensor) {
        for (int t = 1; t < prm.rounds; ++t)
            S2V_RETURN_IF(s2v_tc_round_forward(g, prm.theta2_w, base, mus + (t - 1) * plane, mus + t * plane, st));
    } else {
        int smem = (C::W_FLOATS + C::TILE_FLOATS) * 4;
        S2V_RETURN_IF(set_smem(k_embed_round<P>, smem));
        int grid = s2v_persistent_grid((const void*)k_embed_round<P>, smem, tiles);
        for (int t = 1; t < prm.rounds; ++t)
            k_embed_round<P><<<grid, S2V_THREADS, smem, st>>>(g, prm.theta2_w, base, mus + (t - 1) * plane,
                                                              mus + t * plane, tiles);
    }
    return (int)cudaGetLastError();
}

template <int P>
static int embed_backward_impl(const s2v_graph_t& g, const s2v_embed_params_t& prm, const float* x,
                               const float* base, const float* mus, const float* dmu, int dmu_masked,
                               float* dus, float* partial, float* grad, float* dx, cudaStream_t st) {
    using C = Cfg<P>;
    using EP = EmbedPartial<P>;
    const int T = prm.rounds;
    const int tiles = (g.num_nodes + S2V_TM - 1) / S2V_TM;
    const size_t plane = (size_t)g.num_nodes * P;
    const int stride = EP::size(prm.node_dim);
    float* sums = partial + (size_t)S2V_MAX_GRID * stride;
    // du^T
    const float* du_last = dmu;
    if (!dmu_masked) {
        float* dst = dus + (size_t)(T - 1) * plane;
        int64_t n4 = (int64_t)plane / 4;
        int blocks = (int)((n4 + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        k_relu_mask<<<blocks, 256, 0, st>>>((const float4*)dmu, (const float4*)(mus + (size_t)(T - 1) * plane),
                                            (float4*)dst, n4);
        du_last = dst;
    }
    int smem_f = (C::W_FLOATS + 2 * C::TILE_FLOATS + S2V_MAX_F * P + S2V_TM * S2V_MAX_F + P + C::RG * P) * 4;
    S2V_RETURN_IF(set_smem(k_embed_final_bwd<P>, smem_f));
    int grid_f = s2v_persistent_grid((const void*)k_embed_final_bwd<P>, smem_f, tiles);
    int grid_r = 0;
    // rounds T..2 : du^t -> du^{t-1}
    if (use_tensor_path_bwd_round(P, prm.flags, g.num_nodes)) {
        const bool ws = use_ws_bwd_round(prm.flags);
        grid_r = ws ? s2v_tc_round_backward_sel_grid(g.num_nodes, prm.flags) : s2v_tc_round_backward_grid(g.num_nodes);
        for (int t = T; t >= 2; --t) {
            const float* du_t = (t == T) ? du_last : dus + (size_t)(t - 1) * plane;
            if (ws)
                S2V_RETURN_IF(s2v_tc_round_backward_sel(g, prm.theta2_w, du_t, mus + (size_t)(t - 2) * plane,
                                                        dus + (size_t)(t - 2) * plane, partial, stride, t != T, grid_r, prm.flags, st));
            else
                S2V_RETURN_IF(s2v_tc_round_backward(g, prm.theta2_w, du_t, mus + (size_t)(t - 2) * plane,
                                                    dus + (size_t)(t - 2) * plane, partial, stride, t != T, grid_r, st));
        }
    } else {
        int smem_r = (C::W_FLOATS + 2 * C::TILE_FLOATS) * 4;
        S2V_RETURN_IF(set_smem(k_embed_round_bwd<P>, smem_r));
        grid_r = s2v_persistent_grid((const void*)k_embed_round_bwd<P>, smem_r, tiles);
        for (int t = T; t >= 2; --t) {
            const float* du_t = (t == T) ? du_last : dus + (size_t)(t - 1) * plane;
            k_embed_round_bwd<P><<<grid_r, S2V_THREADS, smem_r, st>>>(g, prm.theta2_w, du_t, mus + (size_t)(t - 2) * plane,
                                                                      dus + (size_t)(t - 2) * plane, partial, stride,
                                                                      t != T, tiles);
        }
    }
    if (dx == nullptr && use_tensor_path(P, prm.flags, g.num_nodes) && use_ws_bwd_round(prm.flags) &&
        s2v_tc_final_backward_h2_selected(prm.flags, prm.node_dim)) {
        grid_f = s2v_tc_final_backward_h2_grid(g.num_nodes);
        S2V_RETURN_IF(s2v_tc_final_backward_h2(g, prm, x, dus, du_last, partial, stride, grid_f, st));
    } else if (dx == nullptr && use_tensor_path(P, prm.flags, g.num_nodes) && use_ws_bwd_round(prm.flags) && s2v_tc_final_backward_ws_ok(prm.node_dim)) {
        grid_f = s2v_tc_round_backward_ws_grid(g.num_nodes);
        S2V_RETURN_IF(s2v_tc_final_backward_ws(g, prm, x, dus, du_last, partial, stride, grid_f, st));
    } else if (dx == nullptr && S2V_TENSOR_FINAL && use_tensor_path(P, prm.flags, g.num_nodes)) {
        grid_f = s2v_tc_final_backward_grid(g.num_nodes);
        S2V_RETURN_IF(s2v_tc_final_backward(g, prm, x, dus, du_last, partial, stride, grid_f, st));
    } else {
        k_embed_final_bwd<P><<<grid_f, S2V_THREADS, smem_f, st>>>(g, prm, x, dus, du_last, T, partial, stride, tiles, dx);   // dx != NULL: CUDA-core last stage
    }
    if (T >= 2) s2v_reduce_partials(partial, stride, grid_r, EP::TH2, P * P, sums, st);
    s2v_reduce_partials(partial, stride, grid_f, EP::TH1B, stride - EP::TH1B, sums, st);
    k_embed_finalize<P><<<1, 256, 0, st>>>(prm, sums, T >= 2, grad);
    (void)base;
    return (int)cudaGetLastError();
}

static int check_graph(const s2v_graph_t* g) {
    if (!g || g->num_nodes < 0 || g->num_graphs < 0) return S2V_ERR_BAD_ARG;
    if (!g->rowptr || !g->rowptr_t || !g->indeg) return S2V_ERR_BAD_ARG;
    if (g->num_edges > 0 && (!g->col || !g->col_t)) return S2V_ERR_BAD_ARG;
    if (g->period == 0 && (!g->ptr || !g->gid) && g->num_nodes > 0) return S2V_ERR_BAD_ARG;
    if (g->period < 0) return S2V_ERR_BAD_ARG;
    if (g->period > 0 && (int64_t)g->period * g->num_graphs != g->num_nodes) return S2V_ERR_BAD_ARG;
    return 0;
}

extern "C" int64_t s2v_embed_grad_floats(int32_t F, int32_t P) {
    return (int64_t)P * F + P + 3 * ((int64_t)P * P + P) + 2 * P;
}

extern "C" int64_t s2v_embed_backward_workspace_bytes(int32_t F, int32_t P) {
    return (int64_t)(S2V_MAX_GRID + 1) * (2 * P * P + 4 * P + P * F) * 4;
}

extern "C" int s2v_embed_forward(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                                 float* base, float* mus, void* stream) {
    S2V_RETURN_IF(check_graph(g));
    if (!prm || !x || !base || !mus || prm->rounds < 1) return S2V_ERR_BAD_ARG;
    if (prm->node_dim < 1 || prm->node_dim > S2V_MAX_F) return S2V_ERR_UNSUPPORTED;
    if (g->num_nodes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    S2V_DISPATCH_P(prm->emb_dim, return embed_forward_impl<PP>(*g, *prm, x, base, mus, st));
    return 0;
}

extern "C" int s2v_embed_backward_dx(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                                     const float* base, const float* mus, const float* dmu, int32_t dmu_masked,
                                     float* dus, void* workspace, int64_t workspace_bytes, float* grad, float* dx, void* stream);

extern "C" int s2v_embed_backward(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                                  const float* base, const float* mus, const float* dmu, int32_t dmu_masked,
                                  float* dus, void* workspace, int64_t workspace_bytes, float* grad, void* stream) {
    return s2v_embed_backward_dx(g, prm, x, base, mus, dmu, dmu_masked, dus, workspace, workspace_bytes, grad, nullptr, stream);
}

extern "C" int s2v_embed_backward_dx(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                                     const float* base, const float* mus, const float* dmu, int32_t dmu_masked,
                                     float* dus, void* workspace, int64_t workspace_bytes, float* grad, float* dx, void* stream) {
    S2V_RETURN_IF(check_graph(g));
    if (!prm || !x || !mus || !dmu || !dus || !workspace || !grad || prm->rounds < 1) return S2V_ERR_BAD_ARG;
    if (prm->node_dim < 1 || prm->node_dim > S2V_MAX_F) return S2V_ERR_UNSUPPORTED;
    if (workspace_bytes < s2v_embed_backward_workspace_bytes(prm->node_dim, prm->emb_dim)) return S2V_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (g->num_nodes == 0) {
        return (int)cudaMemsetAsync(grad, 0, s2v_embed_grad_floats(prm->node_dim, prm->emb_dim) * 4, st);
    }
    S2V_DISPATCH_P(prm->emb_dim, return embed_backward_impl<PP>(*g, *prm, x, base, mus, dmu, dmu_masked, dus,
                                                                 (float*)workspace, grad, dx, st));
    return 0;
}

extern "C" int s2v_relu_mask(const float* dmu, const float* mu, float* du, int64_t n, void* stream) {
    if (!dmu || !mu || !du || n < 0 || (n & 3)) return S2V_ERR_BAD_ARG;
    if (n == 0) return 0;
    int64_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_relu_mask<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)dmu, (const float4*)mu, (float4*)du, n4);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// single-round entry points (unit tests, roofline measurement of the dominant kernels)
// ------------------------------------------------------------------------------------
template <int P>
static int round_forward_impl(const s2v_graph_t& g, const float* theta2_w, const float* base, const float* mu_prev,
                              float* mu_next, cudaStream_t st) {
    using C = Cfg<P>;
    const int tiles = (g.num_nodes + S2V_TM - 1) / S2V_TM;
    int smem = (C::W_FLOATS + C::TILE_FLOATS) * 4;
    S2V_RETURN_IF(set_smem(k_embed_round<P>, smem));
    int grid = s2v_persistent_grid((const void*)k_embed_round<P>, smem, tiles);
    k_embed_round<P><<<grid, S2V_THREADS, smem, st>>>(g, theta2_w, base, mu_prev, mu_next, tiles);
    return (int)cudaGetLastError();
}

template <int P>
static int round_backward_impl(const s2v_graph_t& g, const float* theta2_w, const float* du_t, const float* mu_prev,
                               float* du_prev, float* partial, int F, cudaStream_t st) {
    using C = Cfg<P>;
    const int tiles = (g.num_nodes + S2V_TM - 1) / S2V_TM;
    int smem = (C::W_FLOATS + 2 * C::TILE_FLOATS) * 4;
    S2V_RETURN_IF(set_smem(k_embed_round_bwd<P>, smem));
    int grid = s2v_persistent_grid((const void*)k_embed_round_bwd<P>, smem, tiles);
    k_embed_round_bwd<P><<<grid, S2V_THREADS, smem, st>>>(g, theta2_w, du_t, mu_prev, du_prev, partial,
                                                          EmbedPartial<P>::size(F), 0, tiles);
    return (int)cudaGetLastError();
}

extern "C" int s2v_embed_round_forward(const s2v_graph_t* g, int32_t emb_dim, int32_t flags, const float* theta2_w,
                                       const float* base, const float* mu_prev, float* mu_next, void* stream) {
    S2V_RETURN_IF(check_graph(g));
    if (!theta2_w || !base || !mu_prev || !mu_next) return S2V_ERR_BAD_ARG;
    if (g->num_nodes == 0) return 0;
    if (use_tensor_path(emb_dim, flags, g->num_nodes))
        return s2v_tc_round_forward(*g, theta2_w, base, mu_prev, mu_next, (cudaStream_t)stream);
    S2V_DISPATCH_P(emb_dim, return round_forward_impl<PP>(*g, theta2_w, base, mu_prev, mu_next, (cudaStream_t)stream));
    return 0;
}

extern "C" int s2v_embed_round_backward(const s2v_graph_t* g, int32_t node_dim, int32_t emb_dim, int32_t flags,
                                        const float* theta2_w, const float* du_t, const float* mu_prev, float* du_prev,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    S2V_RETURN_IF(check_graph(g));
    if (!theta2_w || !du_t || !mu_prev || !du_prev || !workspace) return S2V_ERR_BAD_ARG;
    if (workspace_bytes < s2v_embed_backward_workspace_bytes(node_dim, emb_dim)) return S2V_ERR_WORKSPACE;
    if (g->num_nodes == 0) return 0;
    if (use_tensor_path_bwd_round(emb_dim, flags, g->num_nodes)) {
        const int pstride = 2 * emb_dim * emb_dim + 4 * emb_dim + emb_dim * node_dim;
        if (use_ws_bwd_round(flags))
            return s2v_tc_round_backward_sel(*g, theta2_w, du_t, mu_prev, du_prev, (float*)workspace, pstride, 0,
                                             s2v_tc_round_backward_sel_grid(g->num_nodes, flags), flags, (cudaStream_t)stream);
        return s2v_tc_round_backward(*g, theta2_w, du_t, mu_prev, du_prev, (float*)workspace, pstride, 0,
                                     s2v_tc_round_backward_grid(g->num_nodes), (cudaStream_t)stream);
    }
    S2V_DISPATCH_P(emb_dim, return round_backward_impl<PP>(*g, theta2_w, du_t, mu_prev, du_prev, (float*)workspace,
                                                           node_dim, (cudaStream_t)stream));
    return 0;
}
