// Small-graph fast path: the WHOLE forward pass of the s2v Q-network / policy head in ONE cooperative launch.
//
//   stage 0   node MLP + degree term + round 1          (embed_first_body)
//   stage t   propagation round t = 2..T                (embed_round_body)       grid-wide barrier in between
//   stage T+1 pool + theta6 + head + mask + reduction   (head_fused_body, one CTA per graph; graphs of more than 256 nodes:
//             pool_stage_body | head_rows_body | mask_stage_body, tile-parallel with two more grid barriers)
//
// For acting (B = 1: 975 rows = 16 tiles), the reference's Manhattan5 / MemTask batches and the per-rollout PPO calls
// the work of a stage is a few microseconds, so the launch-per-stage path (T + 1 launches, each with its own drain and
// ramp) is latency-bound; here the stages are separated by cooperative-groups grid barriers, the embeddings stay in
// L2 between them, and the stage bodies are the ones the multi-launch CUDA-core path runs (s2v_bodies.cuh): results
// are bit-identical.  Rows written by other CTAs earlier in the launch are read with ld.global.cg (LD = 1).
// Replaces comb_opt.py:297-331 (predict / get_best_action), models_basic_lstm.py:526-566 (pi / v per rollout).
#include <cooperative_groups.h>
#include "s2v_bodies.cuh"

namespace cg = cooperative_groups;

#ifdef S2V_SMALL_PROF
__device__ unsigned long long g_small_prof[32];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define SP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_small_prof[i] = gtime(); } while (0)
extern "C" int s2v_debug_small_prof(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_small_prof, sizeof(unsigned long long) * 32); }
#else
#define SP(i) do {} while (0)
#endif

template <int P, int TM>
__global__ void __launch_bounds__(S2V_THREADS)
k_qnet_small(s2v_graph_t g, s2v_embed_params_t eprm, s2v_head_params_t hprm, const float* __restrict__ x,
             float* base, float* mus, int mode, const uint8_t* __restrict__ mask, const int32_t* __restrict__ reach_ptr,
             const int32_t* __restrict__ reach_idx, float* __restrict__ pool, float* __restrict__ gstate, float* q,
             float* __restrict__ prob, int64_t* __restrict__ arg_out, int64_t* __restrict__ node_out,
             float* __restrict__ val_out, float* pool_ws) {
    extern __shared__ __align__(16) float smem[];
    cg::grid_group grid = cg::this_grid();
    const int tiles = (g.num_nodes + TM - 1) / TM;          // embedding stages: TM-row tiles (the head walks graphs)
    const size_t plane = (size_t)g.num_nodes * P;
    SP(0);
    embed_first_body<P, TM>(smem, g, eprm, x, base, mus, tiles);
    SP(1);
    for (int t = 1; t < eprm.rounds; ++t) {
        __threadfence();
        grid.sync();
        SP(2 * t);
        embed_round_body<P, 1, TM>(smem, g, eprm.theta2_w, base, mus + (t - 1) * plane, mus + t * plane, tiles, t == 1);
        SP(2 * t + 1);
    }
    if (mode < 0) return;                         // embedding only
    __threadfence();
    grid.sync();
    SP(20);
    const float* muT = mus + (size_t)(eprm.rounds - 1) * plane;
    if ((int64_t)g.num_nodes <= 256 * (int64_t)g.num_graphs) {          // small graphs: one CTA walks a whole graph
        head_fused_body<P, 1>(smem, g, hprm, muT, mode, mask, reach_ptr, reach_idx, pool, gstate, q, prob, arg_out, node_out,
                              val_out);
        SP(21);
        return;
    }
    // large graphs (acting on a road network: B = 1, 975 rows): the head as three tile-parallel stages -- pool + theta6
    // per graph | q per row tile | mask + reduction per graph -- with grid barriers in between (k_pool, k_head and the
    // mask kernels run the same bodies: bit-identical)
    if (pool_ws != nullptr) {                     // the rows of a graph pooled by many CTAs: chunk partials | per-graph finish
        pool_partials_body<P, 1>(g, muT, pool_ws);
        __threadfence();
        grid.sync();
        pool_reduce_body<P>(smem, g, hprm.norm_agg, pool_ws, pool, hprm.theta6_w, hprm.theta6_b, gstate);
    } else {
        pool_stage_body<P, 1>(smem, g, hprm.norm_agg, muT, pool, hprm.theta6_w, hprm.theta6_b, gstate);
    }
    __threadfence();
    grid.sync();
    SP(21);
    head_rows_body<P, 1, TM>(smem, g, hprm, muT, gstate, q, tiles);
    SP(22);
    if (mode >= 1) {
        __threadfence();
        grid.sync();
        mask_stage_body(mode, g, q, mask, reach_ptr, reach_idx, prob, arg_out, node_out, val_out);
    }
    SP(23);
}

// Short tiles (16 rows) for emb_dim 64: a stage of one 64-row tile per CTA costs 5.6 us on 128 threads (gather round trips
// + a 64 x 64 x 64 FFMA tile); with 16-row tiles four times as many CTAs share the rows and a stage is ~2 us.
template <int P>
struct SmallTile { static constexpr int TM = P == 64 ? 16 : S2V_TM; };

template <int P>
static int small_forward_impl(const s2v_graph_t& g, const s2v_embed_params_t& eprm, const s2v_head_params_t& hprm,
                              const float* x, float* base, float* mus, int mode, const uint8_t* mask,
                              const int32_t* reach_ptr, const int32_t* reach_idx, float* pool, float* gstate, float* q,
                              float* prob, int64_t* arg_out, int64_t* node_out, float* val_out, float* pool_ws,
                              cudaStream_t st) {
    using C = Cfg<P>;
    constexpr int TM = SmallTile<P>::TM;
    using CT = Cfg<P, TM>;
    const int smem_first = (CT::W_FLOATS + CT::TILE_FLOATS + S2V_MAX_F * P + TM * S2V_MAX_F + 6 * P) * 4;
    const int smem_head = (C::W_FLOATS + C::TILE_FLOATS + S2V_TM * C::TXP + 3 * P + C::RG * P + 2 * P) * 4;
    const int smem_rows = (CT::W_FLOATS + CT::TILE_FLOATS + 2 * TM * CT::TXP + 3 * P) * 4;      // head_rows_body
    int smem = smem_first > smem_head ? smem_first : smem_head;
    if (smem_rows > smem) smem = smem_rows;
    auto kernel = k_qnet_small<P, TM>;
    S2V_RETURN_IF(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    S2V_RETURN_IF(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, S2V_THREADS, smem));
    if (per_sm < 1) return S2V_ERR_UNSUPPORTED;
    const int tiles = (g.num_nodes + TM - 1) / TM;
    int want = tiles > g.num_graphs ? tiles : g.num_graphs;
    int grid = sms * per_sm;                      // every CTA must be resident: cooperative launch
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    s2v_graph_t gg = g;
    s2v_embed_params_t ep = eprm;
    s2v_head_params_t hp = hprm;
    void* args[] = {&gg, &ep, &hp, (void*)&x, &base, &mus, &mode, (void*)&mask, (void*)&reach_ptr, (void*)&reach_idx,
                    &pool, &gstate, &q, &prob, &arg_out, &node_out, &val_out, &pool_ws};
    return (int)cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(S2V_THREADS), args, smem, st);
}

// mode -1: embedding only (base, mus);  0..3: as s2v_head_fused_forward.  Returns S2V_ERR_UNSUPPORTED when the problem
// belongs to the tensor-core path (s2v_qnet_small_eligible() == 0): the caller then uses the per-stage launches.
extern "C" int s2v_qnet_small_eligible(int32_t emb_dim, int32_t flags, int64_t num_nodes, int64_t num_graphs) {
    if (emb_dim == 64 && flags >= S2V_PATH_TENSOR && flags <= S2V_PATH_TENSOR_16) return 0;
    if (emb_dim == 64 && flags != S2V_PATH_FFMA && num_nodes >= S2V_TENSOR_MIN_NODES) return 0;
    if (num_nodes > 16 * S2V_TENSOR_MIN_NODES) return 0;        // beyond that a stage fills the machine: per-stage launches
    (void)num_graphs;                                           // graphs of any size: fused or tile-parallel head stage
    return 1;
}

// Workspace of the many-CTA pool (floats): one [RG][P] partial per 64-row chunk of every graph; 0 when the graphs are small
// enough for one CTA each (the fused head stage).
extern "C" int64_t s2v_qnet_small_workspace_floats(int32_t emb_dim, int64_t num_nodes, int64_t num_graphs) {
    if (emb_dim < 8 || emb_dim > 64 || emb_dim % 4 || num_nodes <= 256 * num_graphs) return 0;
    const int64_t rg = S2V_THREADS / (emb_dim / 4);
    return (num_nodes / S2V_POOL_CHUNK + num_graphs) * rg * emb_dim;
}

extern "C" int s2v_qnet_small_forward(const s2v_graph_t* g, const s2v_embed_params_t* eprm, const s2v_head_params_t* hprm,
                                      const float* x, float* base, float* mus, int32_t mode, const uint8_t* mask,
                                      const int32_t* reach_ptr, const int32_t* reach_idx, float* pool, float* gstate,
                                      float* q, float* prob, int64_t* arg_out, int64_t* node_out, float* val_out,
                                      void* stream) {
    return s2v_qnet_small_forward_ws(g, eprm, hprm, x, base, mus, mode, mask, reach_ptr, reach_idx, pool, gstate, q, prob,
                                     arg_out, node_out, val_out, nullptr, 0, stream);
}

extern "C" int s2v_qnet_small_forward_ws(const s2v_graph_t* g, const s2v_embed_params_t* eprm, const s2v_head_params_t* hprm,
                                         const float* x, float* base, float* mus, int32_t mode, const uint8_t* mask,
                                         const int32_t* reach_ptr, const int32_t* reach_idx, float* pool, float* gstate,
                                         float* q, float* prob, int64_t* arg_out, int64_t* node_out, float* val_out,
                                         float* workspace, int64_t workspace_floats, void* stream) {
    if (!g || !eprm || !x || !base || !mus || eprm->rounds < 1 || mode < -1 || mode > 3) return S2V_ERR_BAD_ARG;
    if (g->num_nodes < 0 || g->num_graphs < 0 || g->period < 0 || !g->rowptr || !g->indeg) return S2V_ERR_BAD_ARG;
    if (g->num_edges > 0 && !g->col) return S2V_ERR_BAD_ARG;
    if (g->period == 0 && (!g->ptr || !g->gid) && g->num_nodes > 0) return S2V_ERR_BAD_ARG;
    if (g->period > 0 && (int64_t)g->period * g->num_graphs != g->num_nodes) return S2V_ERR_BAD_ARG;
    if (mode >= 0 && (!hprm || !pool || !gstate || !q || hprm->emb_dim != eprm->emb_dim)) return S2V_ERR_BAD_ARG;
    if (mode == 1 && (!mask || !prob)) return S2V_ERR_BAD_ARG;
    if (mode == 2 && (!mask || !arg_out || !val_out)) return S2V_ERR_BAD_ARG;
    if (mode == 3 && (!reach_ptr || !arg_out || !node_out || !val_out)) return S2V_ERR_BAD_ARG;
    if (eprm->node_dim < 1 || eprm->node_dim > S2V_MAX_F) return S2V_ERR_UNSUPPORTED;
    if (!s2v_qnet_small_eligible(eprm->emb_dim, eprm->flags, g->num_nodes, g->num_graphs)) return S2V_ERR_UNSUPPORTED;
    if (g->num_nodes == 0) return 0;
    s2v_head_params_t hp{};
    if (hprm) hp = *hprm;
    cudaStream_t st = (cudaStream_t)stream;
    // the many-CTA pool needs its workspace; without one (or with too small a one) one CTA pools each graph
    float* pool_ws = workspace;
    const int64_t need = s2v_qnet_small_workspace_floats(eprm->emb_dim, g->num_nodes, g->num_graphs);
    if (mode < 0 || need == 0 || workspace_floats < need) pool_ws = nullptr;
    S2V_DISPATCH_P(eprm->emb_dim, return small_forward_impl<PP>(*g, *eprm, hp, x, base, mus, mode, mask, reach_ptr, reach_idx,
                                                                 pool, gstate, q, prob, arg_out, node_out, val_out, pool_ws, st));
    return 0;
}
