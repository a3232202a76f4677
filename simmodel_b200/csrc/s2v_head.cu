// Pooling + Q / logit head of struct2vec and its backward.
// Math (modules/gnn/comb_opt.py:248-273,
// modules/ppo/models_basic_lstm.py:530-538,548-562):
//   pool_g = (1/n_g) sum_{i in g} mu_i         (norm_agg; plain sum otherwise; count clamped >= 1)
//   G_g    = theta6 pool_g + b6 ;  L_i = theta7 mu_i + b7
//   q_i    = w5[:p] . relu(G_g(i)) + w5[p:] . relu(L_i) + b5
#include "s2v_common.cuh"
#include "s2v_reduce.cuh"
#include "s2v_bodies.cuh"

// ------------------------------------------------------------------------------------
// pooling: one CTA per graph, fixed summation order
// ------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_pool(s2v_graph_t g, int norm_agg, const float* __restrict__ mu, float* __restrict__ pool,
       const float* __restrict__ theta6_w, const float* __restrict__ theta6_b, float* __restrict__ gstate) {
    using C = Cfg<P>;
    __shared__ __align__(16) float pool_smem[C::RG * P + P];
    pool_stage_body<P>(pool_smem, g, norm_agg, mu, pool, theta6_w, theta6_b, gstate);
}

template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_pool_bwd(s2v_graph_t g, int norm_agg, const float* __restrict__ dpool, float* __restrict__ dmu) {
    constexpr int V = P / 4;
    const int64_t total = (int64_t)g.num_nodes * V;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(e / V), c = (int)(e - (int64_t)r * V);
        int li, off, gi;
        row_info(g, r, li, off, gi);
        int lo, hi;
        graph_range(g, gi, lo, hi);
        int n = hi - lo;
        float4 d = ldg4(dpool + (size_t)gi * P + 4 * c);
        if (norm_agg) { float dn = (float)(n > 1 ? n : 1); d.x /= dn; d.y /= dn; d.z /= dn; d.w /= dn; }
        st4(dmu + (size_t)r * P + 4 * c, d);
    }
}

// ------------------------------------------------------------------------------------
// head forward: tile kernel
// ------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_head(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ mu, const float* __restrict__ gstate,
       float* __restrict__ q, int num_tiles) {
    extern __shared__ __align__(16) float smem[];
    head_rows_body<P>(smem, g, prm, mu, gstate, q, num_tiles);
}

// ------------------------------------------------------------------------------------
// fused head: pool -> theta6 -> tile head -> invalid-action mask + per-graph reduction, ONE launch.
// One CTA per graph (persistent over graphs).  The three phases are the bodies of k_pool, k_head and the mask kernels
// (s2v_mask.cu) run back to back by the same 128 threads with the same thread mapping and summation orders, so every
// output is bit-identical to the three-launch sequence pool / head / mask on the CUDA-core path; q stays in L1/L2
// between the phases.  Used where launches are the cost (acting, small batches, the per-rollout PPO call pattern):
//   mode 0  q only                                  QNet.forward                        comb_opt.py:248-273
//   mode 1  prob = softmax(q | reachable)           PPO pi                              models_basic_lstm.py:530-540
//   mode 2  (max, argmax) of q over reachable       PPO v (critic q), greedy evaluation models_basic_lstm.py:548-564
//   mode 3  argmax over a neighbour list            DQN greedy action                   comb_opt.py:319-326
// ------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_head_fused(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ mu, int mode,
             const uint8_t* __restrict__ mask, const int32_t* __restrict__ reach_ptr, const int32_t* __restrict__ reach_idx,
             float* __restrict__ pool, float* __restrict__ gstate, float* q, float* __restrict__ prob,
             int64_t* __restrict__ arg_out, int64_t* __restrict__ node_out, float* __restrict__ val_out) {
    extern __shared__ __align__(16) float smem[];
    head_fused_body<P, 0>(smem, g, prm, mu, mode, mask, reach_ptr, reach_idx, pool, gstate, q, prob, arg_out, node_out, val_out);
}

// ------------------------------------------------------------------------------------
// head backward
// ------------------------------------------------------------------------------------
// per graph workspace row: dG [P], dpool_scaled [P], s_g, pad
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_head_bwd_graph(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ gstate,
                 const float* __restrict__ dq, float* __restrict__ per_graph, unsigned int* __restrict__ nnz_total) {
    __shared__ float red[S2V_THREADS];
    __shared__ float dG[P];
    const int gi = blockIdx.x, t = threadIdx.x;
    int lo, hi;
    graph_range(g, gi, lo, hi);
    float s = 0.f;
    int nz = 0;
    for (int r = lo + t; r < hi; r += S2V_THREADS) { const float d = __ldg(dq + r); s += d; nz += d != 0.f; }
    red[t] = s;
    nz = __reduce_add_sync(0xffffffffu, nz);            // integer count: the atomic sum is order-independent
    if ((t & 31) == 0 && nz) atomicAdd(nnz_total, (unsigned int)nz);
    __syncthreads();
    for (int w = S2V_THREADS / 2; w > 0; w >>= 1) {
        if (t < w) red[t] += red[t + w];
        __syncthreads();
    }
    const float sg = red[0];
    float* out = per_graph + (size_t)gi * (2 * P + 4);
    if (t < P) {
        float G = __ldg(gstate + (size_t)gi * P + t);
        float d = G > 0.f ? sg * __ldg(prm.theta5_w + t) : 0.f;
        dG[t] = d;
        out[t] = d;
    }
    __syncthreads();
    if (t < P) {
        float a = 0.f;
        for (int n = 0; n < P; ++n) a = fmaf(__ldg(prm.theta6_w + n * P + t), dG[n], a);
        int cnt = hi - lo;
        if (prm.norm_agg) a /= (float)(cnt > 1 ? cnt : 1);
        out[P + t] = a;
    }
    if (t == 0) { out[2 * P] = sg; out[2 * P + 1] = 0.f; }
}

template <int P>
struct HeadPartial {
    static constexpr int TH7 = 0;              // P*P  sum_i dL_i mu_i^T
    static constexpr int B7 = P * P;           // P    sum_i dL_i
    static constexpr int W5B = B7 + P;         // P    sum_i dq_i relu(L_i)
    static constexpr int SIZE = W5B + P;
};

// dq with few non-zeros (the DQN loss touches one node per graph): dmu_i = dpool_g(i) [* (mu_i > 0)] for every row, as
// one streaming pass with all loads of a thread in flight; k_head_bwd_rows then revisits only the rows with a non-zero dq
// (k_head_bwd, the tile kernel, handles dense dq -- the PPO losses -- and small batches).
#define HEAD_SPARSE(nnz, n) ((unsigned long long)(nnz) * 8ull < (unsigned long long)(n))
template <int P>
__global__ void __launch_bounds__(256)
k_head_bwd_stream(s2v_graph_t g, const float* __restrict__ mu, const float* __restrict__ per_graph, int mask_relu,
                  float* __restrict__ dmu, const unsigned int* __restrict__ nnz_total) {
    if (!HEAD_SPARSE(*nnz_total, g.num_nodes)) return;
    constexpr int V = P / 4, U = 8;
    const long long total = (long long)g.num_nodes * V, stride = (long long)gridDim.x * 256;
    for (long long e0 = (long long)blockIdx.x * 256 + threadIdx.x; e0 < total; e0 += stride * U) {
        float4 o[U], m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = e0 + u * stride;
            o[u] = f4zero(); m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (e < total) {
                const int r = (int)(e / V), c = (int)(e - (long long)r * V);
                const int gi = g.period > 0 ? r / g.period : __ldg(g.gid + r);
                o[u] = ldg4(per_graph + (size_t)gi * (2 * P + 4) + P + 4 * c);
                if (mask_relu) m[u] = ldg4(mu + (size_t)r * P + 4 * c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = e0 + u * stride;
            if (e < total)
                st4(dmu + e * 4, make_float4(m[u].x > 0.f ? o[u].x : 0.f, m[u].y > 0.f ? o[u].y : 0.f,
                                             m[u].z > 0.f ? o[u].z : 0.f, m[u].w > 0.f ? o[u].w : 0.f));
        }
    }
}

// Sparse dq, second pass: only the ROWS with dq != 0 carry a local term.  Per such row (one warp):
//   L = theta7 mu_i + b7 ; dL = dq_i w5b * [L > 0] ; dtheta7 += dL mu_i^T ; db7 += dL ; dw5b += dq_i relu(L) ;
//   dmu_i = theta7^T dL + dpool_g(i)  [* (mu_i > 0)]   (overwrites what k_head_bwd_stream wrote for the row).
// CTA c owns a contiguous slab of rows, warp w a quarter of it, scanned 32 rows at a time in row order; dtheta7 is
// accumulated per warp in shared memory (lane = column k: conflict-free) and the four warps are summed in fixed order.
template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_head_bwd_rows(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ mu, const float* __restrict__ dq,
                const float* __restrict__ per_graph, int mask_relu, float* __restrict__ dmu,
                float* __restrict__ partial, int rows_per_cta, const unsigned int* __restrict__ nnz_total) {
    using HP = HeadPartial<P>;
    if (!HEAD_SPARSE(*nnz_total, g.num_nodes)) return;
    constexpr int LDW = P + 1, NW = S2V_THREADS / 32, KPL = (P + 31) / 32;     // KPL = outputs / columns per lane
    extern __shared__ __align__(16) float smem[];
    float* W7 = smem;                              // theta7[n][k], row stride P+1
    float* acc = W7 + P * LDW;                     // [NW][P*P] per-warp dtheta7
    float* bmu = acc + NW * P * P;                 // [NW][P]
    float* bdl = bmu + NW * P;                     // [NW][P]
    float* vec = bdl + NW * P;                     // b7 [P], w5b [P]
    float* red = vec + 2 * P;                      // [NW][2P]
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int e = t; e < P * P; e += S2V_THREADS) W7[(e / P) * LDW + e % P] = __ldg(prm.theta7_w + e);
    for (int e = t; e < NW * P * P; e += S2V_THREADS) acc[e] = 0.f;
    if (t < P) { vec[t] = __ldg(prm.theta7_b + t); vec[P + t] = __ldg(prm.theta5_w + P + t); }
    __syncthreads();
    float pB7[KPL], pW5[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) pB7[q] = pW5[q] = 0.f;
    float* my_acc = acc + w * P * P, *my_mu = bmu + w * P, *my_dl = bdl + w * P;
    const int sub = rows_per_cta / NW;
    const long long lo = (long long)blockIdx.x * rows_per_cta + (long long)w * sub;
    const long long hi = lo + sub < (long long)g.num_nodes ? lo + sub : (long long)g.num_nodes;
    for (long long base = lo; base < hi; base += 32) {
        const long long r = base + lane;
        const float d = r < hi ? __ldg(dq + r) : 0.f;
        unsigned live = __ballot_sync(0xffffffffu, d != 0.f);
        while (live) {
            const int b = __ffs(live) - 1;
            live &= live - 1;
            const int rr = (int)(base + b);
            const float dd = __shfl_sync(0xffffffffu, d, b);
            float muv[KPL];
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                const int k = lane + 32 * q;
                muv[q] = k < P ? __ldg(mu + (size_t)rr * P + k) : 0.f;
                if (k < P) my_mu[k] = muv[q];
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < KPL; ++q) {                                     // lane = output n
                const int n = lane + 32 * q;
                if (n < P) {
                    float L = vec[n];
                    for (int k = 0; k < P; ++k) L = fmaf(W7[n * LDW + k], my_mu[k], L);
                    const float dl = L > 0.f ? dd * vec[P + n] : 0.f;
                    pW5[q] = fmaf(dd, relu(L), pW5[q]);
                    pB7[q] += dl;
                    my_dl[n] = dl;
                }
            }
            __syncwarp();
            int li, off, gi;
            row_info(g, rr, li, off, gi);
#pragma unroll
            for (int q = 0; q < KPL; ++q) {                                     // lane = column k
                const int k = lane + 32 * q;
                if (k < P) {
                    float o = __ldg(per_graph + (size_t)gi * (2 * P + 4) + P + k);
                    for (int n = 0; n < P; ++n) {
                        const float dl = my_dl[n];                              // warp-uniform
                        if (dl != 0.f) {
                            my_acc[n * P + k] = fmaf(dl, muv[q], my_acc[n * P + k]);
                            o = fmaf(W7[n * LDW + k], dl, o);
                        }
                    }
                    dmu[(size_t)rr * P + k] = (!mask_relu || muv[q] > 0.f) ? o : 0.f;
                }
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
        const int n = lane + 32 * q;
        if (n < P) { red[w * 2 * P + n] = pB7[q]; red[w * 2 * P + P + n] = pW5[q]; }
    }
    __syncthreads();
    float* my = partial + (size_t)blockIdx.x * HP::SIZE;
    for (int e = t; e < P * P; e += S2V_THREADS) {
        float s2 = 0.f;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) s2 += acc[ww * P * P + e];
        my[HP::TH7 + e] = s2;
    }
    if (t < 2 * P) {
        float s2 = 0.f;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) s2 += red[ww * 2 * P + t];
        my[(t < P ? HP::B7 : HP::W5B - P) + t] = s2;
    }
}

template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_head_bwd(s2v_graph_t g, s2v_head_params_t prm, const float* __restrict__ mu, const float* __restrict__ dq,
           const float* __restrict__ per_graph, int mask_relu, float* __restrict__ dmu,
           float* __restrict__ partial, int num_tiles, const unsigned int* __restrict__ nnz_total) {
    if (nnz_total != nullptr && HEAD_SPARSE(*nnz_total, g.num_nodes)) return;   // sparse dq: stream + rows kernels
    using C = Cfg<P>;
    using HP = HeadPartial<P>;
    extern __shared__ __align__(16) float smem[];
    float* Wt7 = smem;                        // Wt[k][n] = theta7[n][k]   (L = mu theta7^T)
    float* W7 = Wt7 + C::W_FLOATS;            // theta7 as stored [n][k]     (dmu = dL theta7)
    float* As = W7 + C::W_FLOATS;             // mu tile
    float* Bs = As + C::TILE_FLOATS;          // dL tile
    float* vec = Bs + C::TILE_FLOATS;         // b7, w5b
    float* b7 = vec, *w5b = vec + P;
    float* scratch = vec + 2 * P;             // [RG][P]
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    load_w_transposed<P>(Wt7, prm.theta7_w);
    load_w_plain<P>(W7, prm.theta7_w);
    if (t < P) { b7[t] = __ldg(prm.theta7_b + t); w5b[t] = __ldg(prm.theta5_w + P + t); }
    float racc[C::NRR][4];
#pragma unroll
    for (int i = 0; i < C::NRR; ++i) racc[i][0] = racc[i][1] = racc[i][2] = racc[i][3] = 0.f;
    float pB7[4] = {0.f, 0.f, 0.f, 0.f}, pW5[4] = {0.f, 0.f, 0.f, 0.f};
    __syncthreads();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * S2V_TM;
        const int rows = min(S2V_TM, g.num_nodes - row0);
        // A tile whose dq entries are all zero has dL = 0: nothing to add to dtheta7/db7/dw5b and
        // dmu reduces to the pooled term.  The DQN loss touches ONE node per graph
        // (qvals[actions + ptr[:-1]], comb_opt.py:355-356), so almost every tile takes this exit.
        const bool live = __syncthreads_or(t < rows && __ldg(dq + row0 + t) != 0.f);
        if (!live) {
            // streaming exit: all loads of the thread's chunks are issued before the first store; the graph of a row
            // comes from one division per tile (replicated layout) instead of one per chunk
            constexpr int V = P / 4;
            constexpr int NI = (S2V_TM * V + S2V_THREADS - 1) / S2V_THREADS;
            int gi0 = 0, nb = S2V_TM;                           // rows [0, nb) of the tile belong to graph gi0
            if (g.period > 0) { gi0 = row0 / g.period; nb = (gi0 + 1) * g.period - row0; }
            float4 o[NI], m[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int e = t + i * S2V_THREADS, rl = e / V, c = e - rl * V;
                o[i] = f4zero(); m[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (e < rows * V) {
                    int gi;
                    if (g.period > 0) gi = rl < nb ? gi0 : gi0 + 1 + (rl - nb) / g.period;
                    else gi = __ldg(g.gid + row0 + rl);
                    o[i] = ldg4(per_graph + (size_t)gi * (2 * P + 4) + P + 4 * c);
                    if (mask_relu) m[i] = ldg4(mu + (size_t)(row0 + rl) * P + 4 * c);
                }
            }
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int e = t + i * S2V_THREADS, rl = e / V, c = e - rl * V;
                if (e < rows * V)
                    st4(dmu + (size_t)(row0 + rl) * P + 4 * c,
                        make_float4(m[i].x > 0.f ? o[i].x : 0.f, m[i].y > 0.f ? o[i].y : 0.f,
                                    m[i].z > 0.f ? o[i].z : 0.f, m[i].w > 0.f ? o[i].w : 0.f));
            }
            continue;
        }
        load_tile<P>(As, mu + (size_t)row0 * P, rows);
        __syncthreads();
        if (active) {
            float acc[C::NR][4];
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P>(As, Wt7, acc, tx, ty);
            float4 bb = ld4(b7 + 4 * tx), wb = ld4(w5b + 4 * tx);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= S2V_TM) continue;
                float4 dL = f4zero();
                if (rl < rows) {
                    float d = __ldg(dq + row0 + rl);
                    float l0 = acc[i][0] + bb.x, l1 = acc[i][1] + bb.y, l2 = acc[i][2] + bb.z, l3 = acc[i][3] + bb.w;
                    dL = make_float4(l0 > 0.f ? d * wb.x : 0.f, l1 > 0.f ? d * wb.y : 0.f,
                                     l2 > 0.f ? d * wb.z : 0.f, l3 > 0.f ? d * wb.w : 0.f);
                    pW5[0] = fmaf(d, relu(l0), pW5[0]); pW5[1] = fmaf(d, relu(l1), pW5[1]);
                    pW5[2] = fmaf(d, relu(l2), pW5[2]); pW5[3] = fmaf(d, relu(l3), pW5[3]);
                    pB7[0] += dL.x; pB7[1] += dL.y; pB7[2] += dL.z; pB7[3] += dL.w;
                }
                st4(Bs + rl * C::LDA + 4 * tx, dL);
            }
        }
        __syncthreads();
        if (active) {
            tile_reduce_gemm<P>(Bs, As, rows, racc, tx, ty);
            float acc[C::NR][4];
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P>(Bs, W7, acc, tx, ty);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= rows) continue;
                int li, off, gi;
                row_info(g, row0 + rl, li, off, gi);
                float4 dp = ldg4(per_graph + (size_t)gi * (2 * P + 4) + P + 4 * tx);
                float4 o = make_float4(acc[i][0] + dp.x, acc[i][1] + dp.y, acc[i][2] + dp.z, acc[i][3] + dp.w);
                if (mask_relu) {
                    float4 m = ld4(As + rl * C::LDA + 4 * tx);
                    o = make_float4(m.x > 0.f ? o.x : 0.f, m.y > 0.f ? o.y : 0.f, m.z > 0.f ? o.z : 0.f, m.w > 0.f ? o.w : 0.f);
                }
                st4(dmu + (size_t)(row0 + rl) * P + 4 * tx, o);
            }
        }
        __syncthreads();
    }
    float* my = partial + (size_t)blockIdx.x * HP::SIZE;
    store_racc<P>(my + HP::TH7, racc, tx, ty, active, false);
    reduce_columns<P>(scratch, pB7, my + HP::B7, tx, ty, active, false);
    reduce_columns<P>(scratch, pW5, my + HP::W5B, tx, ty, active, false);
}

// Per-graph quantities reduced over the graphs of the batch, 64 graphs per tile:
//   dtheta6 += dG_g pool_g^T,  db6 += dG_g,  dw5[:P] += s_g relu(G_g),  db5 += s_g
template <int P>
struct HeadGraphPartial {
    static constexpr int TH6 = 0;              // P*P
    static constexpr int B6 = P * P;           // P
    static constexpr int W5A = B6 + P;         // P
    static constexpr int B5 = W5A + P;         // 4 (one used)
    static constexpr int SIZE = B5 + 4;
};

template <int P>
__global__ void __launch_bounds__(S2V_THREADS)
k_head_graph_partials(int num_graphs, const float* __restrict__ per_graph, const float* __restrict__ pool,
                      const float* __restrict__ gstate, float* __restrict__ partial, int num_tiles) {
    using C = Cfg<P>;
    using GP = HeadGraphPartial<P>;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                         // dG tile   [64 graphs][P]
    float* Bs = As + C::TILE_FLOATS;          // pool tile
    float* Gs = Bs + C::TILE_FLOATS;          // s_g relu(G_g) tile
    float* sg = Gs + C::TILE_FLOATS;          // [64]
    float* scratch = sg + S2V_TM;             // [RG][P]
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    constexpr int PG = 2 * P + 4, V = P / 4;
    float racc[C::NRR][4];
#pragma unroll
    for (int i = 0; i < C::NRR; ++i) racc[i][0] = racc[i][1] = racc[i][2] = racc[i][3] = 0.f;
    float pB6[4] = {0.f, 0.f, 0.f, 0.f}, pW5[4] = {0.f, 0.f, 0.f, 0.f};
    float pB5 = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g0 = tile * S2V_TM;
        const int rows = min(S2V_TM, num_graphs - g0);
        if (t < S2V_TM) sg[t] = t < rows ? __ldg(per_graph + (size_t)(g0 + t) * PG + 2 * P) : 0.f;
        __syncthreads();
        for (int e = t; e < S2V_TM * V; e += S2V_THREADS) {
            int r = e / V, c = e - r * V;
            float4 a = f4zero(), b = f4zero(), gq = f4zero();
            if (r < rows) {
                a = ldg4(per_graph + (size_t)(g0 + r) * PG + 4 * c);
                b = ldg4(pool + (size_t)(g0 + r) * P + 4 * c);
                float4 G = ldg4(gstate + (size_t)(g0 + r) * P + 4 * c);
                float s = sg[r];
                gq = make_float4(s * relu(G.x), s * relu(G.y), s * relu(G.z), s * relu(G.w));
            }
            st4(As + r * C::LDA + 4 * c, a);
            st4(Bs + r * C::LDA + 4 * c, b);
            st4(Gs + r * C::LDA + 4 * c, gq);
        }
        __syncthreads();
        if (active) {
            tile_reduce_gemm<P>(As, Bs, rows, racc, tx, ty);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int r = ty + C::RG * i;
                if (r >= rows) continue;
                float4 a = ld4(As + r * C::LDA + 4 * tx), gq = ld4(Gs + r * C::LDA + 4 * tx);
                pB6[0] += a.x; pB6[1] += a.y; pB6[2] += a.z; pB6[3] += a.w;
                pW5[0] += gq.x; pW5[1] += gq.y; pW5[2] += gq.z; pW5[3] += gq.w;
            }
        }
        if (t == 0) for (int r = 0; r < rows; ++r) pB5 += sg[r];
        __syncthreads();
    }
    float* my = partial + (size_t)blockIdx.x * GP::SIZE;
    store_racc<P>(my + GP::TH6, racc, tx, ty, active, false);
    reduce_columns<P>(scratch, pB6, my + GP::B6, tx, ty, active, false);
    reduce_columns<P>(scratch, pW5, my + GP::W5A, tx, ty, active, false);
    if (t == 0) { my[GP::B5] = pB5; my[GP::B5 + 1] = 0.f; my[GP::B5 + 2] = 0.f; my[GP::B5 + 3] = 0.f; }
}

// grad layout: theta5.w [2P] (global part, local part), theta5.b [1], theta6.w [P,P], theta6.b [P],
//              theta7.w [P,P], theta7.b [P];  sums: HeadPartial | HeadGraphPartial
template <int P>
__global__ void __launch_bounds__(256)
k_head_finalize(const float* __restrict__ sums, float* __restrict__ grad) {
    using HP = HeadPartial<P>;
    using GP = HeadGraphPartial<P>;
    const float* sg = sums + HP::SIZE;
    const int t = threadIdx.x;
    float* g5_w = grad;
    float* g5_b = g5_w + 2 * P;
    float* g6_w = g5_b + 1;
    float* g6_b = g6_w + P * P;
    float* g7_w = g6_b + P;
    float* g7_b = g7_w + P * P;
    for (int e = t; e < P; e += blockDim.x) {
        g5_w[e] = sg[GP::W5A + e];
        g5_w[P + e] = sums[HP::W5B + e];
        g6_b[e] = sg[GP::B6 + e];
        g7_b[e] = sums[HP::B7 + e];
    }
    if (t == 0) g5_b[0] = sg[GP::B5];
    for (int e = t; e < P * P; e += blockDim.x) {
        g6_w[e] = sg[GP::TH6 + e];
        g7_w[e] = sums[HP::TH7 + e];
    }
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------

int s2v_tc_head_forward(const s2v_graph_t& g, const s2v_head_params_t& prm, const float* mu, const float* gstate, float* q,
                        cudaStream_t st);          // s2v_embed_tc.cu (p = 64)
// same rule as the embedding (s2v_embed.cu:use_tensor_path)
static bool head_tensor_path(int P, int flags, int num_nodes) {
    if (P != 64 || flags == S2V_PATH_FFMA) return false;
    if (flags >= S2V_PATH_TENSOR && flags <= S2V_PATH_TENSOR_16) return true;
    return num_nodes >= S2V_TENSOR_MIN_NODES;
}

template <int P>
static int head_forward_impl(const s2v_graph_t& g, const s2v_head_params_t& prm, const float* mu,
                             float* pool, float* gstate, float* q, cudaStream_t st) {
    using C = Cfg<P>;
    k_pool<P><<<g.num_graphs, S2V_THREADS, 0, st>>>(g, prm.norm_agg, mu, pool, prm.theta6_w, prm.theta6_b, gstate);
    const int tiles = (g.num_nodes + S2V_TM - 1) / S2V_TM;
    if (tiles > 0 && head_tensor_path(P, prm.flags, g.num_nodes)) {
        S2V_RETURN_IF(s2v_tc_head_forward(g, prm, mu, gstate, q, st));
    } else if (tiles > 0) {
        int smem = (C::W_FLOATS + C::TILE_FLOATS + 2 * S2V_TM * C::TXP + 3 * P) * 4;
        S2V_RETURN_IF(cudaFuncSetAttribute(k_head<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int grid = s2v_persistent_grid((const void*)k_head<P>, smem, tiles);
        k_head<P><<<grid, S2V_THREADS, smem, st>>>(g, prm, mu, gstate, q, tiles);
    }
    return (int)cudaGetLastError();
}

static inline int64_t head_ws_floats(int P, int num_graphs) {
    return (int64_t)num_graphs * (2 * P + 4) + (int64_t)(S2V_MAX_GRID + 1) * (2 * P * P + 4 * P + 4);
}

template <int P>
static int head_backward_impl(const s2v_graph_t& g, const s2v_head_params_t& prm, const float* mu,
                              const float* pool, const float* gstate, const float* dq, int mask_relu,
                              float* dmu, float* ws, float* grad, cudaStream_t st) {
    using C = Cfg<P>;
    using HP = HeadPartial<P>;
    float* per_graph = ws;
    float* partial = ws + (size_t)g.num_graphs * (2 * P + 4);
    unsigned int* nnz_total = reinterpret_cast<unsigned int*>(ws + head_ws_floats(P, g.num_graphs));
    S2V_RETURN_IF(cudaMemsetAsync(nnz_total, 0, sizeof(unsigned int), st));
    k_head_bwd_graph<P><<<g.num_graphs, S2V_THREADS, 0, st>>>(g, prm, gstate, dq, per_graph, nnz_total);
    const int tiles = (g.num_nodes + S2V_TM - 1) / S2V_TM;
    const bool split = g.num_nodes >= S2V_TENSOR_MIN_NODES;      // small batches: one launch less matters more
    if (split) {
        const long long chunks = (long long)g.num_nodes * (P / 4);
        long long want = (chunks + 256 * 8 - 1) / (256 * 8);
        int sgrid = (int)(want < 148 * 8 ? want : 148 * 8);
        k_head_bwd_stream<P><<<sgrid, 256, 0, st>>>(g, mu, per_graph, mask_relu, dmu, nnz_total);
    }
    int smem = (2 * C::W_FLOATS + 2 * C::TILE_FLOATS + 2 * P + C::RG * P) * 4;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_head_bwd<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = s2v_persistent_grid((const void*)k_head_bwd<P>, smem, tiles);
    k_head_bwd<P><<<grid, S2V_THREADS, smem, st>>>(g, prm, mu, dq, per_graph, mask_relu, dmu, partial, tiles,
                                                   split ? nnz_total : nullptr);
    if (split) {                                   // same grid => same partial rows, whichever of the two kernels is active
        int rpc = (g.num_nodes + grid - 1) / grid;
        rpc = (rpc + S2V_THREADS - 1) / S2V_THREADS * S2V_THREADS;
        int smem_r = (P * (P + 1) + 4 * P * P + 4 * P + 4 * P + 2 * P + 4 * 2 * P) * 4;
        S2V_RETURN_IF(cudaFuncSetAttribute(k_head_bwd_rows<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
        k_head_bwd_rows<P><<<grid, S2V_THREADS, smem_r, st>>>(g, prm, mu, dq, per_graph, mask_relu, dmu, partial, rpc, nnz_total);
    }
    // per-graph terms, then fixed-order sums of both partial sets
    using GP = HeadGraphPartial<P>;
    float* partial_g = partial + (size_t)S2V_MAX_GRID * HP::SIZE;
    float* sums = partial_g + (size_t)S2V_MAX_GRID * GP::SIZE;
    const int gtiles = (g.num_graphs + S2V_TM - 1) / S2V_TM;
    int smem_g = (3 * C::TILE_FLOATS + S2V_TM + C::RG * P) * 4;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_head_graph_partials<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_g));
    int grid_g = gtiles < 128 ? gtiles : 128;
    k_head_graph_partials<P><<<grid_g, S2V_THREADS, smem_g, st>>>(g.num_graphs, per_graph, pool, gstate, partial_g, gtiles);
    s2v_reduce_partials(partial, HP::SIZE, grid, 0, HP::SIZE, sums, st);
    s2v_reduce_partials(partial_g, GP::SIZE, grid_g, 0, GP::SIZE, sums + HP::SIZE, st);
    k_head_finalize<P><<<1, 256, 0, st>>>(sums, grad);
    return (int)cudaGetLastError();
}


static int check_graph_h(const s2v_graph_t* g) {
    if (!g || g->num_nodes < 0 || g->num_graphs < 0 || g->period < 0) return S2V_ERR_BAD_ARG;
    if (g->period == 0 && (!g->ptr || !g->gid) && g->num_nodes > 0) return S2V_ERR_BAD_ARG;
    if (g->period > 0 && (int64_t)g->period * g->num_graphs != g->num_nodes) return S2V_ERR_BAD_ARG;
    return 0;
}

extern "C" int64_t s2v_head_grad_floats(int32_t P) { return 2 * (int64_t)P + 1 + 2 * ((int64_t)P * P + P); }

extern "C" int64_t s2v_head_backward_workspace_bytes(int32_t P, int32_t num_graphs) {
    return head_ws_floats(P, num_graphs) * 4 + 64;      // + the non-zero counter of dq
}

extern "C" int s2v_head_forward(const s2v_graph_t* g, const s2v_head_params_t* prm, const float* mu,
                                float* pool, float* gstate, float* q, void* stream) {
    S2V_RETURN_IF(check_graph_h(g));
    if (!prm || !mu || !pool || !gstate || !q) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    S2V_DISPATCH_P(prm->emb_dim, return head_forward_impl<PP>(*g, *prm, mu, pool, gstate, q, st));
    return 0;
}

extern "C" int s2v_head_backward(const s2v_graph_t* g, const s2v_head_params_t* prm, const float* mu,
                                 const float* pool, const float* gstate, const float* dq, int32_t mask_relu,
                                 float* dmu, void* workspace, int64_t workspace_bytes, float* grad, void* stream) {
    S2V_RETURN_IF(check_graph_h(g));
    if (!prm || !mu || !pool || !gstate || !dq || !dmu || !workspace || !grad) return S2V_ERR_BAD_ARG;
    if (workspace_bytes < s2v_head_backward_workspace_bytes(prm->emb_dim, g->num_graphs)) return S2V_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (g->num_nodes == 0 || g->num_graphs == 0)
        return (int)cudaMemsetAsync(grad, 0, s2v_head_grad_floats(prm->emb_dim) * 4, st);
    S2V_DISPATCH_P(prm->emb_dim, return head_backward_impl<PP>(*g, *prm, mu, pool, gstate, dq, mask_relu, dmu,
                                                                (float*)workspace, grad, st));
    return 0;
}

extern "C" int s2v_head_fused_forward(const s2v_graph_t* g, const s2v_head_params_t* prm, const float* mu, int32_t mode,
                                      const uint8_t* mask, const int32_t* reach_ptr, const int32_t* reach_idx,
                                      float* pool, float* gstate, float* q, float* prob, int64_t* arg_out, int64_t* node_out,
                                      float* val_out, void* stream) {
    S2V_RETURN_IF(check_graph_h(g));
    if (!prm || !mu || !pool || !gstate || !q || mode < 0 || mode > 3) return S2V_ERR_BAD_ARG;
    if (mode == 1 && (!mask || !prob)) return S2V_ERR_BAD_ARG;
    if (mode == 2 && (!mask || !arg_out || !val_out)) return S2V_ERR_BAD_ARG;
    if (mode == 3 && (!reach_ptr || !arg_out || !node_out || !val_out)) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define S2V_HEAD_FUSED_LAUNCH(P_)                                                                                          \
    {                                                                                                                      \
        using C = Cfg<P_>;                                                                                                 \
        const int smem = (C::W_FLOATS + C::TILE_FLOATS + S2V_TM * C::TXP + 3 * P_ + C::RG * P_ + 2 * P_) * 4;             \
        S2V_RETURN_IF(cudaFuncSetAttribute(k_head_fused<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));          \
        const int grid = s2v_persistent_grid((const void*)k_head_fused<P_>, smem, g->num_graphs);                          \
        k_head_fused<P_><<<grid, S2V_THREADS, smem, st>>>(*g, *prm, mu, mode, mask, reach_ptr, reach_idx, pool, gstate, q,   \
                                                         prob, arg_out, node_out, val_out);                                \
        return (int)cudaGetLastError();                                                                                    \
    }
    S2V_DISPATCH_P(prm->emb_dim, S2V_HEAD_FUSED_LAUNCH(PP));
#undef S2V_HEAD_FUSED_LAUNCH
    return 0;
}

/* 1 when the fused launch is the better choice: the problem runs the CUDA-core head (so the fused launch is
 * bit-identical to pool / head / mask) and its graphs are small (at most 256 nodes on average). */
extern "C" int s2v_head_fused_eligible(int32_t emb_dim, int32_t flags, int64_t num_nodes, int64_t num_graphs) {
    if (head_tensor_path(emb_dim, flags, (int)(num_nodes > 0x7fffffff ? 0x7fffffff : num_nodes))) return 0;
    // one CTA walks a whole graph: right for many small graphs (Manhattan5, MemTask, rollout steps of small worlds); a
    // few large graphs (B = 1 acting on a 975-node road network) keep the tile-parallel head + separate reduction
    return num_nodes <= 256 * num_graphs ? 1 : 0;
}

extern "C" int s2v_pool_forward(const s2v_graph_t* g, int32_t P, int32_t norm_agg, const float* mu, float* pool, void* stream) {
    S2V_RETURN_IF(check_graph_h(g));
    if (!mu || !pool) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    S2V_DISPATCH_P(P, (k_pool<PP><<<g->num_graphs, S2V_THREADS, 0, st>>>(*g, norm_agg, mu, pool, nullptr, nullptr, nullptr)));
    return (int)cudaGetLastError();
}

extern "C" int s2v_pool_backward(const s2v_graph_t* g, int32_t P, int32_t norm_agg, const float* dpool, float* dmu, void* stream) {
    S2V_RETURN_IF(check_graph_h(g));
    if (!dpool || !dmu) return S2V_ERR_BAD_ARG;
    if (g->num_nodes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = (int64_t)g->num_nodes * (P / 4);
    int blocks = (int)((total + S2V_THREADS - 1) / S2V_THREADS);
    if (blocks > 148 * 16) blocks = 148 * 16;
    S2V_DISPATCH_P(P, (k_pool_bwd<PP><<<blocks, S2V_THREADS, 0, st>>>(*g, norm_agg, dpool, dmu)));
    return (int)cudaGetLastError();
}
