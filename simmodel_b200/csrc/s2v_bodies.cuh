// Stage bodies of the CUDA-core (FP32 FFMA) forward path, shared by the one-launch-per-stage kernels
// (s2v_embed.cu, s2v_head.cu) and by the single-launch small-graph kernel (s2v_small.cu): the same code, the same
// thread mapping, the same summation orders -> bit-identical results whichever way a problem is launched.
#pragma once
#include "s2v_common.cuh"
#include "s2v_reduce.cuh"

// node MLP + degree term + round 1 (comb_opt.py:230-238; mu^0 = 0)
template <int P, int TM = S2V_TM>
__device__ __forceinline__ void embed_first_body(float* smem, const s2v_graph_t& g, const s2v_embed_params_t& prm,
                                                 const float* __restrict__ x, float* __restrict__ base,
                                                 float* __restrict__ mu1, int num_tiles) {
    using C = Cfg<P, TM>;
    const int F = prm.node_dim;
    float* Wt1b = smem;                       // [P][P]  Wt[k][n] = theta1b[n][k]
    float* As = Wt1b + C::W_FLOATS;           // h tile
    float* W1a = As + C::TILE_FLOATS;         // [F][P]  W1a[f][n] = theta1a[n][f]
    float* xs = W1a + S2V_MAX_F * P;          // [TM][F]
    float* vec = xs + TM * S2V_MAX_F;     // b1a, cst, u1, u0, r1, r0   (6 x P)
    float* b1a = vec, *cst = vec + P, *u1 = vec + 2 * P, *u0 = vec + 3 * P, *r1 = vec + 4 * P, *r0 = vec + 5 * P;
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;

    load_w_transposed<P>(Wt1b, prm.theta1b_w);
    for (int e = t; e < P * F; e += S2V_THREADS) { int n = e / F, f = e - n * F; W1a[f * P + n] = __ldg(prm.theta1a_w + e); }
    if (t < P) {
        b1a[t] = __ldg(prm.theta1a_b + t);
        float w4 = __ldg(prm.theta4_w + t), b4 = __ldg(prm.theta4_b + t);
        r1[t] = relu(w4 + b4);
        r0[t] = relu(b4);
    }
    __syncthreads();
    if (t < P) {
        u1[t] = row_dot<P>(prm.theta3_w + t * P, r1, 0.f);
        u0[t] = row_dot<P>(prm.theta3_w + t * P, r0, 0.f);
        cst[t] = __ldg(prm.theta1b_b + t) + __ldg(prm.theta2_b + t) + __ldg(prm.theta3_b + t);
    }
    __syncthreads();

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TM;
        const int rows = min(TM, g.num_nodes - row0);
        // in-degree and padded size of this thread's output rows: requested here, used in the epilogue (as raw integers,
        // so that nothing waits for them before the tile product is done)
        int ci[C::NR], npi[C::NR];
#pragma unroll
        for (int i = 0; i < C::NR; ++i) {
            const int rl = ty + C::RG * i;
            ci[i] = 0; npi[i] = 0;
            if (active && rl < rows) {
                int li, off, gi;
                row_info(g, row0 + rl, li, off, gi);
                ci[i] = __ldg(g.indeg + li);
                npi[i] = graph_npad(g, gi);
            }
        }
        for (int e = t; e < TM * F; e += S2V_THREADS)
            xs[e] = e < rows * F ? __ldg(x + (size_t)row0 * F + e) : 0.f;
        __syncthreads();
        if (active) {
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int r = ty + C::RG * i;
                if (r >= TM) continue;
                float4 h = ld4(b1a + 4 * tx);
                for (int f = 0; f < F; ++f) {
                    float xv = xs[r * F + f];
                    float4 w = ld4(W1a + f * P + 4 * tx);
                    h.x = fmaf(xv, w.x, h.x); h.y = fmaf(xv, w.y, h.y); h.z = fmaf(xv, w.z, h.z); h.w = fmaf(xv, w.w, h.w);
                }
                st4(As + r * C::LDA + 4 * tx, make_float4(relu(h.x), relu(h.y), relu(h.z), relu(h.w)));
            }
        }
        __syncthreads();
        if (active) {
            float acc[C::NR][4];
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P, TM>(As, Wt1b, acc, tx, ty);
            float4 c0 = ld4(cst + 4 * tx), v1 = ld4(u1 + 4 * tx), v0 = ld4(u0 + 4 * tx);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= rows) continue;
                const int r = row0 + rl;
                const float c = (float)ci[i];
                const float d = (float)npi[i] - c;
                float4 b;
                b.x = acc[i][0] + (c0.x + (c * v1.x + d * v0.x));
                b.y = acc[i][1] + (c0.y + (c * v1.y + d * v0.y));
                b.z = acc[i][2] + (c0.z + (c * v1.z + d * v0.z));
                b.w = acc[i][3] + (c0.w + (c * v1.w + d * v0.w));
                st4(base + (size_t)r * P + 4 * tx, b);
                st4(mu1 + (size_t)r * P + 4 * tx, make_float4(relu(b.x), relu(b.y), relu(b.z), relu(b.w)));
            }
        }
        __syncthreads();
    }
}

// one propagation round (comb_opt.py:241-242)
template <int P, int LD = 0, int TM = S2V_TM>
__device__ __forceinline__ void embed_round_body(float* smem, const s2v_graph_t& g, const float* __restrict__ theta2_w,
                                                 const float* __restrict__ base, const float* mu_prev,
                                                 float* __restrict__ mu_next, int num_tiles, bool load_weights = true) {
    using C = Cfg<P, TM>;
    float* Wt2 = smem;                        // Wt[k][n] = theta2[n][k]
    float* As = Wt2 + C::W_FLOATS;            // gathered tile
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    if (load_weights) { load_w_transposed<P>(Wt2, theta2_w); __syncthreads(); }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TM;
        const int rows = min(TM, g.num_nodes - row0);
        gather_tile<P, LD, TM>(As, g, g.rowptr, g.col, mu_prev, row0, rows);
        __syncthreads();
        if (active) {
            float acc[C::NR][4];
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P, TM>(As, Wt2, acc, tx, ty);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= rows) continue;
                size_t o = (size_t)(row0 + rl) * P + 4 * tx;
                float4 b = ld_row4<LD>(base + o);
                st4(mu_next + o, make_float4(relu(acc[i][0] + b.x), relu(acc[i][1] + b.y),
                                             relu(acc[i][2] + b.z), relu(acc[i][3] + b.w)));
            }
        }
        __syncthreads();
    }
}

// Pooling of one graph, fixed summation order.  Row r (relative to the graph start) belongs to thread group
// ty = r % RG; a thread first sums its rows INSIDE every 64-row chunk of the graph (ascending), then the chunk sums
// (ascending), and the RG group totals are added in order.  The chunked association is what lets several CTAs pool one
// large graph (k_qnet_small: one partial per chunk and group, summed here in the same order) with bit-identical results.
#define S2V_POOL_CHUNK 64
template <int P, int LD = 0>
__device__ __forceinline__ float4 pool_chunk_partial(const float* mu, int lo, int c0, int c1, int ty, int tx) {
    using C = Cfg<P>;
    float4 c = f4zero();                                   // rows c0 <= r < c1 (relative) with r % RG == ty
    const int r0 = c0 + ((ty - c0 % C::RG) + C::RG) % C::RG;
    constexpr int NPT = S2V_POOL_CHUNK / C::RG + 1;        // rows of a chunk a thread can own
    float4 v[NPT];                                         // all loads first: one memory round trip per chunk, not one per row
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
        const int r = r0 + u * C::RG;
        v[u] = r < c1 ? ld_row4<LD>(mu + (size_t)(lo + r) * P + 4 * tx) : f4zero();
    }
#pragma unroll
    for (int u = 0; u < NPT; ++u)
        if (r0 + u * C::RG < c1) c = f4add(c, v[u]);       // same order as the plain loop; absent rows add nothing
    return c;
}
template <int P>
__device__ __forceinline__ void pool_finish(float4 s, float* scratch, float* pool_s, int tx, int ty, bool active) {
    using C = Cfg<P>;
    if (active) st4(scratch + ty * P + 4 * tx, s);
    __syncthreads();
    if (threadIdx.x < P) {
        float a = 0.f;
        for (int k = 0; k < C::RG; ++k) a += scratch[k * P + threadIdx.x];
        pool_s[threadIdx.x] = a;
    }
    __syncthreads();
}
template <int P, int LD = 0>
__device__ __forceinline__ void pool_graph(const float* mu, int lo, int hi, float* scratch, float* pool_s, float scale) {
    using C = Cfg<P>;
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    float4 s = f4zero();
    if (t < C::ACTIVE) {
        const int n = hi - lo;
        int c0 = 0;
        for (; c0 + 4 * S2V_POOL_CHUNK <= n; c0 += 4 * S2V_POOL_CHUNK) {     // four chunk sums in flight, added in chunk order
            float4 c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                c[u] = pool_chunk_partial<P, LD>(mu, lo, c0 + u * S2V_POOL_CHUNK, c0 + (u + 1) * S2V_POOL_CHUNK, ty, tx);
#pragma unroll
            for (int u = 0; u < 4; ++u) s = f4add(s, c[u]);
        }
        for (; c0 < n; c0 += S2V_POOL_CHUNK)
            s = f4add(s, pool_chunk_partial<P, LD>(mu, lo, c0, min(n, c0 + S2V_POOL_CHUNK), ty, tx));
    }
    pool_finish<P>(s, scratch, pool_s, tx, ty, t < C::ACTIVE);
    (void)scale;
}

// Graph term of q: sg = sum over the column slices tx (ascending) of the fma chain wa[c] * relu(G[c]) over the slice's four
// columns.  q_i = (b5 + sum_tx sL_i[tx]) + sg, with sL the same chain over wb[c] * relu(L_i[c] + b7[c]): the graph term is
// a per-graph constant, kept apart from the row term so that rows and graphs can be finished by different CTAs.
template <int P, int LD = 0>
__device__ __forceinline__ float head_graph_slice(const float* w5a, const float* G, int tx) {
    const float4 wa = ld4(w5a + 4 * tx), g4 = LD ? ld_row4<LD>(G + 4 * tx) : ld4(G + 4 * tx);
    float s = 0.f;
    s = fmaf(wa.x, relu(g4.x), s); s = fmaf(wa.y, relu(g4.y), s);
    s = fmaf(wa.z, relu(g4.z), s); s = fmaf(wa.w, relu(g4.w), s);
    return s;
}

// Invalid-action mask + per-graph reduction over q[lo, hi) (the code of s2v_mask.cu's kernels, one CTA of MK_THREADS).
//   mode 1 masked softmax -> prob; 2 masked max / arg-max; 3 arg-max over a neighbour list; anything else: nothing.
__device__ __forceinline__ void masked_reduce_graph(int mode, int gi, int lo, int hi, const float* q,
                                                    const uint8_t* __restrict__ mask, const int32_t* __restrict__ reach_ptr,
                                                    const int32_t* __restrict__ reach_idx, float* __restrict__ prob,
                                                    int64_t* __restrict__ arg_out, int64_t* __restrict__ node_out,
                                                    float* __restrict__ val_out, ArgPair* sh_arg, float* sh_f) {
    const int t = threadIdx.x;
    if (mode == 1) {
        float m = -CUDART_INF_F;
        for (int r = lo + t; r < hi; r += MK_THREADS)
            if (mask[r]) m = fmaxf(m, q[r]);
        m = block_max(m, sh_f);
        float ssum = 0.f;
        for (int r = lo + t; r < hi; r += MK_THREADS)
            if (mask[r]) ssum += expf(q[r] - m);
        ssum = block_sum(ssum, sh_f);
        for (int r = lo + t; r < hi; r += MK_THREADS) {
            float v = mask[r] ? expf(q[r] - m) / ssum : 0.f;
            if (m == -CUDART_INF_F) v = CUDART_NAN_F;
            prob[r] = v;
        }
    } else if (mode == 2) {
        ArgPair p; p.v = -CUDART_INF_F; p.i = -1;
        for (int r = lo + t; r < hi; r += MK_THREADS)
            if (mask[r]) { ArgPair c; c.v = q[r]; c.i = r - lo; p = better(p, c); }
        p = block_argmax(p, sh_arg);
        if (t == 0) {
            arg_out[gi] = p.i;
            val_out[gi] = p.i >= 0 ? p.v : -CUDART_INF_F;
        }
    } else if (mode == 3) {
        const int beg = reach_ptr[gi], end = reach_ptr[gi + 1];
        ArgPair p; p.v = -CUDART_INF_F; p.i = -1;
        for (int k = beg + t; k < end; k += MK_THREADS) {
            ArgPair c; c.v = q[lo + reach_idx[k]]; c.i = k - beg;
            p = better(p, c);
        }
        p = block_argmax(p, sh_arg);
        if (t == 0) {
            arg_out[gi] = p.i;
            node_out[gi] = p.i >= 0 ? (int64_t)reach_idx[beg + p.i] : -1;
            val_out[gi] = p.i >= 0 ? p.v : -CUDART_INF_F;
        }
    }
}

// pool + theta6 + head + mask + per-graph reduction (see k_head_fused, s2v_head.cu)
template <int P, int LD = 0>
__device__ __forceinline__ void head_fused_body(float* smem, const s2v_graph_t& g, const s2v_head_params_t& prm, const float* mu,
                                                int mode, const uint8_t* __restrict__ mask, const int32_t* __restrict__ reach_ptr,
                                                const int32_t* __restrict__ reach_idx, float* __restrict__ pool,
                                                float* __restrict__ gstate, float* q, float* __restrict__ prob,
                                                int64_t* __restrict__ arg_out, int64_t* __restrict__ node_out,
                                                float* __restrict__ val_out) {
    using C = Cfg<P>;
    static_assert(S2V_THREADS == MK_THREADS, "the fused kernel runs the mask reductions with their own thread count");
    float* Wt7 = smem;                        // Wt[k][n] = theta7[n][k]
    float* As = Wt7 + C::W_FLOATS;            // mu tile
    float* red = As + C::TILE_FLOATS;         // [TM][TXP]
    float* vec = red + S2V_TM * C::TXP;       // b7, w5a, w5b
    float* scratch = vec + 3 * P;             // [RG][P] pooling partials
    float* pool_s = scratch + C::RG * P;      // [P]
    float* g_s = pool_s + P;                  // [P] theta6 pool + b6
    __shared__ ArgPair sh_arg[MK_THREADS / 32];
    __shared__ float sh_f[MK_THREADS / 32];
    float* b7 = vec, *w5a = vec + P, *w5b = vec + 2 * P;
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    load_w_transposed<P>(Wt7, prm.theta7_w);
    if (t < P) {
        b7[t] = __ldg(prm.theta7_b + t);
        w5a[t] = __ldg(prm.theta5_w + t);
        w5b[t] = __ldg(prm.theta5_w + P + t);
    }
    const float b5 = __ldg(prm.theta5_b);
    __syncthreads();
    for (int gi = blockIdx.x; gi < g.num_graphs; gi += gridDim.x) {
        int lo, hi;
        graph_range(g, gi, lo, hi);
        const int n = hi - lo;
        // ---- phase 1: pool + theta6 (k_pool) ----
        pool_graph<P, LD>(mu, lo, hi, scratch, pool_s, 1.f);
        if (t < P) {
            float pv = prm.norm_agg ? pool_s[t] / (float)(n > 1 ? n : 1) : pool_s[t];
            pool_s[t] = pv;
            pool[(size_t)gi * P + t] = pv;
        }
        __syncthreads();
        if (t < P) {
            float a = row_dot<P>(prm.theta6_w + t * P, pool_s, __ldg(prm.theta6_b + t));
            g_s[t] = a;
            gstate[(size_t)gi * P + t] = a;
        }
        __syncthreads();
        if (t < C::TXP) scratch[t] = head_graph_slice<P>(w5a, g_s, t);       // graph term of q (scratch is free again)
        __syncthreads();
        float sg = 0.f;
        for (int k = 0; k < C::TXP; ++k) sg += scratch[k];
        __syncthreads();
        // ---- phase 2: q of the graph's rows, 64 at a time (k_head) ----
        for (int row0 = lo; row0 < hi; row0 += S2V_TM) {
            const int rows = min(S2V_TM, hi - row0);
            load_tile<P, LD>(As, mu + (size_t)row0 * P, rows);
            __syncthreads();
            if (active) {
                float acc[C::NR][4];
#pragma unroll
                for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                tile_gemm<P>(As, Wt7, acc, tx, ty);
                const float4 bb = ld4(b7 + 4 * tx), wb = ld4(w5b + 4 * tx);
#pragma unroll
                for (int i = 0; i < C::NR; ++i) {
                    int rl = ty + C::RG * i;
                    if (rl >= S2V_TM) continue;
                    float sacc = 0.f;
                    if (rl < rows) {
                        sacc = fmaf(wb.x, relu(acc[i][0] + bb.x), sacc); sacc = fmaf(wb.y, relu(acc[i][1] + bb.y), sacc);
                        sacc = fmaf(wb.z, relu(acc[i][2] + bb.z), sacc); sacc = fmaf(wb.w, relu(acc[i][3] + bb.w), sacc);
                    }
                    red[rl * C::TXP + tx] = sacc;
                }
            }
            __syncthreads();
            if (t < rows) {
                float sacc = b5;
                for (int k = 0; k < C::TXP; ++k) sacc += red[t * C::TXP + k];
                q[row0 + t] = sacc + sg;
            }
            __syncthreads();                   // also orders the q stores of this CTA before the reads of phase 3
        }
        // ---- phase 3: invalid-action mask + per-graph reduction (s2v_mask.cu) ----
        masked_reduce_graph(mode, gi, lo, hi, q, mask, reach_ptr, reach_idx, prob, arg_out, node_out, val_out, sh_arg, sh_f);
        __syncthreads();
    }
}

// ---- the head as three tile-parallel stages (k_pool | k_head | mask kernels): the per-stage kernels and the cooperative
// ---- kernel of s2v_small.cu (graphs too large for one CTA each) run these bodies
// pool + theta6 of the graphs blockIdx.x, + gridDim.x, ...      smem: [RG][P] scratch + [P]
template <int P, int LD = 0>
__device__ __forceinline__ void pool_stage_body(float* smem, const s2v_graph_t& g, int norm_agg, const float* mu,
                                                float* __restrict__ pool, const float* __restrict__ theta6_w,
                                                const float* __restrict__ theta6_b, float* __restrict__ gstate) {
    using C = Cfg<P>;
    float* scratch = smem;
    float* pool_s = scratch + C::RG * P;
    const int t = threadIdx.x;
    for (int gi = blockIdx.x; gi < g.num_graphs; gi += gridDim.x) {
        int lo, hi;
        graph_range(g, gi, lo, hi);
        const int n = hi - lo;
        pool_graph<P, LD>(mu, lo, hi, scratch, pool_s, 1.f);
        if (t < P) {
            // torch: sum / count (a true division), not sum * (1/count)
            float pv = norm_agg ? pool_s[t] / (float)(n > 1 ? n : 1) : pool_s[t];
            pool_s[t] = pv;
            pool[(size_t)gi * P + t] = pv;
        }
        __syncthreads();
        if (gstate != nullptr && t < P) {
            float a = row_dot<P>(theta6_w + t * P, pool_s, __ldg(theta6_b + t));
            gstate[(size_t)gi * P + t] = a;
        }
        __syncthreads();
    }
}

// The same pooling by MANY CTAs (one large graph: acting on a road network).  Item = (graph, 64-row chunk of the graph);
// items blockIdx.x, + gridDim.x, ... : thread (ty, tx) stores its chunk partial to ws[(item * RG + ty) * P + 4 tx].
// pool_reduce_body then sums a graph's chunk partials per thread in chunk order and finishes as pool_graph does: the same
// association, so the result has the same bits as one CTA walking the whole graph.
template <int P, int LD = 0>
__device__ __forceinline__ void pool_partials_body(const s2v_graph_t& g, const float* mu, float* __restrict__ ws) {
    using C = Cfg<P>;
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    if (t >= C::ACTIVE) return;
    int item0 = 0;
    for (int gi = 0; gi < g.num_graphs; ++gi) {
        int lo, hi;
        graph_range(g, gi, lo, hi);
        const int n = hi - lo, nch = (n + S2V_POOL_CHUNK - 1) / S2V_POOL_CHUNK;
        const int first = item0 + (((int)blockIdx.x - item0 % (int)gridDim.x) + (int)gridDim.x) % (int)gridDim.x;
        for (int it = first; it < item0 + nch; it += gridDim.x) {
            const int c0 = (it - item0) * S2V_POOL_CHUNK;
            const float4 c = pool_chunk_partial<P, LD>(mu, lo, c0, min(n, c0 + S2V_POOL_CHUNK), ty, tx);
            st4(ws + ((size_t)it * C::RG + ty) * P + 4 * tx, c);
        }
        item0 += nch;
    }
}

template <int P>
__device__ __forceinline__ void pool_reduce_body(float* smem, const s2v_graph_t& g, int norm_agg, const float* ws,
                                                 float* __restrict__ pool, const float* __restrict__ theta6_w,
                                                 const float* __restrict__ theta6_b, float* __restrict__ gstate) {
    using C = Cfg<P>;
    float* scratch = smem;
    float* pool_s = scratch + C::RG * P;
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    int item0 = 0;
    for (int gi = 0; gi < g.num_graphs; ++gi) {
        int lo, hi;
        graph_range(g, gi, lo, hi);
        const int n = hi - lo, nch = (n + S2V_POOL_CHUNK - 1) / S2V_POOL_CHUNK;
        if (gi % (int)gridDim.x == (int)blockIdx.x) {
            float4 s = f4zero();
            if (t < C::ACTIVE) {
                const float* base = ws + ((size_t)item0 * C::RG + ty) * P + 4 * tx;
                int c = 0;
                for (; c + 8 <= nch; c += 8) {               // eight chunk partials in flight, added in chunk order
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = ld_row4<1>(base + (size_t)(c + u) * C::RG * P);
#pragma unroll
                    for (int u = 0; u < 8; ++u) s = f4add(s, v[u]);
                }
                for (; c < nch; ++c) s = f4add(s, ld_row4<1>(base + (size_t)c * C::RG * P));
            }
            pool_finish<P>(s, scratch, pool_s, tx, ty, t < C::ACTIVE);
            if (t < P) {
                float pv = norm_agg ? pool_s[t] / (float)(n > 1 ? n : 1) : pool_s[t];
                pool_s[t] = pv;
                pool[(size_t)gi * P + t] = pv;
            }
            __syncthreads();
            if (t < P) {
                float a = row_dot<P>(theta6_w + t * P, pool_s, __ldg(theta6_b + t));
                gstate[(size_t)gi * P + t] = a;
            }
            __syncthreads();
        }
        item0 += nch;
    }
}

// q of the row tiles blockIdx.x, + gridDim.x, ... (TM rows each); gstate = theta6 pool + b6 per graph
template <int P, int LD = 0, int TM = S2V_TM>
__device__ __forceinline__ void head_rows_body(float* smem, const s2v_graph_t& g, const s2v_head_params_t& prm, const float* mu,
                                               const float* gstate, float* __restrict__ q, int num_tiles) {
    using C = Cfg<P, TM>;
    float* Wt7 = smem;                        // Wt[k][n] = theta7[n][k]
    float* As = Wt7 + C::W_FLOATS;            // mu tile
    float* red = As + C::TILE_FLOATS;         // [TM][TXP] row term slices
    float* redg = red + TM * C::TXP;          // [TM][TXP] graph term slices of the row's graph
    float* vec = redg + TM * C::TXP;          // b7, w5a, w5b
    float* b7 = vec, *w5a = vec + P, *w5b = vec + 2 * P;
    const int t = threadIdx.x;
    const int tx = t % C::TXP, ty = t / C::TXP;
    const bool active = t < C::ACTIVE;
    if ((int)blockIdx.x >= num_tiles) return;
    load_w_transposed<P>(Wt7, prm.theta7_w);
    if (t < P) {
        b7[t] = __ldg(prm.theta7_b + t);
        w5a[t] = __ldg(prm.theta5_w + t);
        w5b[t] = __ldg(prm.theta5_w + P + t);
    }
    const float b5 = __ldg(prm.theta5_b);
    __syncthreads();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TM;
        const int rows = min(TM, g.num_nodes - row0);
        load_tile<P, LD, TM>(As, mu + (size_t)row0 * P, rows);
        __syncthreads();
        if (active) {
            float acc[C::NR][4];
#pragma unroll
            for (int i = 0; i < C::NR; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            tile_gemm<P, TM>(As, Wt7, acc, tx, ty);
            float4 bb = ld4(b7 + 4 * tx), wb = ld4(w5b + 4 * tx);
#pragma unroll
            for (int i = 0; i < C::NR; ++i) {
                int rl = ty + C::RG * i;
                if (rl >= TM) continue;
                float s = 0.f, sgs = 0.f;
                if (rl < rows) {
                    int li, off, gi;
                    row_info(g, row0 + rl, li, off, gi);
                    sgs = head_graph_slice<P, LD>(w5a, gstate + (size_t)gi * P, tx);   // graph term (same value for every row of a graph)
                    s = fmaf(wb.x, relu(acc[i][0] + bb.x), s); s = fmaf(wb.y, relu(acc[i][1] + bb.y), s);
                    s = fmaf(wb.z, relu(acc[i][2] + bb.z), s); s = fmaf(wb.w, relu(acc[i][3] + bb.w), s);
                }
                red[rl * C::TXP + tx] = s;
                redg[rl * C::TXP + tx] = sgs;
            }
        }
        __syncthreads();
        if (t < rows) {
            float s = b5, sg = 0.f;
            for (int k = 0; k < C::TXP; ++k) s += red[t * C::TXP + k];
            for (int k = 0; k < C::TXP; ++k) sg += redg[t * C::TXP + k];
            q[row0 + t] = s + sg;
        }
        __syncthreads();
    }
}

// invalid-action mask + per-graph reduction of the graphs blockIdx.x, + gridDim.x, ...
__device__ __forceinline__ void mask_stage_body(int mode, const s2v_graph_t& g, const float* q, const uint8_t* __restrict__ mask,
                                                const int32_t* __restrict__ reach_ptr, const int32_t* __restrict__ reach_idx,
                                                float* __restrict__ prob, int64_t* __restrict__ arg_out,
                                                int64_t* __restrict__ node_out, float* __restrict__ val_out) {
    __shared__ ArgPair sh_arg2[MK_THREADS / 32];
    __shared__ float sh_f2[MK_THREADS / 32];
    for (int gi = blockIdx.x; gi < g.num_graphs; gi += gridDim.x) {
        int lo, hi;
        graph_range(g, gi, lo, hi);
        masked_reduce_graph(mode, gi, lo, hi, q, mask, reach_ptr, reach_idx, prob, arg_out, node_out, val_out, sh_arg2, sh_f2);
        __syncthreads();
    }
}
