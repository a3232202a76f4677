/*
 * s2v_b200.h -- C ABI of the B200-native struct2vec (s2v) hot path.
 *
 * Drop-in boundary for the s2v / GNN-embedding path of rvdweerd/simmodel.  The
 * reference is pure Python; what its hot path "binds" are the torch / torch_scatter /
 * torch_geometric library ops it calls.  Each entry point below names the reference
 * lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocates all
 *     buffers, workspaces and outputs); nothing is allocated or freed here;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and never
 *     synchronise; calls are capturable into CUDA graphs;
 *   - return value: 0 on success, otherwise a cudaError_t value (> 0) or one of the
 *     S2V_ERR_* codes (< 0); no exceptions cross the ABI; s2v_error_string() decodes;
 *   - no global mutable state: re-entrant, callable from any host thread;
 *   - all floating point is fp32, node/edge indices are int32 on the device
 *     (edge_index arrives as the int64 torch/PyG tensor and is narrowed by
 *     s2v_csr_build);
 *   - reductions are atomics-free with a fixed summation order: results are
 *     run-to-run bit-stable for a given device and problem size.
 *
 * Graph layout (s2v_graph_t)
 *   Row form  : row i gathers the columns j with W[i,j] != 0, i.e. the edges with
 *               edge_index[0] == i (conn_matrices.matmul(mu), comb_opt.py:241).
 *   Transpose : column j lists the rows i that gather it (used by backward and for
 *               the in-degree c_j of the theta4 term, comb_opt.py:236-237).
 *   period==0 : "batched" layout, the CSR covers all num_nodes rows of the batch
 *               (PyG Batch.from_data_list, comb_opt.py:299,344); ptr/gid give the
 *               graph of each node.
 *   period>0  : "replicated" layout, the CSR describes ONE graph of `period` nodes
 *               and the batch is num_graphs copies of it with different features
 *               (the PPO path: S time steps sharing one ei,
 *               models_basic_lstm.py:507-518).  Row r = s*period + i.
 */
#ifndef S2V_B200_H
#define S2V_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2V_ABI_VERSION 5

#define S2V_ERR_BAD_ARG        (-1)  /* null pointer / negative size / inconsistent sizes     */
#define S2V_ERR_UNSUPPORTED    (-2)  /* emb_dim not in {8,16,24,32,48,64}, node_dim > 32; GAT: heads*channels > 128 */
#define S2V_ERR_WORKSPACE      (-3)  /* workspace smaller than s2v_*_workspace_bytes()        */

typedef struct s2v_graph {
    int32_t num_nodes;        /* sum of real node counts over the batch (rows)            */
    int32_t num_graphs;       /* B                                                         */
    int32_t period;           /* 0 = batched CSR over num_nodes rows, >0 = replicated      */
    int32_t num_edges;        /* entries in col / col_t                                    */
    const int32_t* rowptr;    /* [n_csr+1], n_csr = period ? period : num_nodes            */
    const int32_t* col;       /* [num_edges]                                               */
    const int32_t* rowptr_t;  /* [n_csr+1]                                                 */
    const int32_t* col_t;     /* [num_edges]                                               */
    const int32_t* indeg;     /* [n_csr] column counts c_j                                 */
    const int32_t* ptr;       /* [B+1] node offset of each graph (NULL iff period>0)       */
    const int32_t* gid;       /* [num_nodes] graph id of each node (NULL iff period>0)     */
    const int32_t* npad;      /* [B] padded size per graph, or NULL -> npad_scalar         */
    int32_t npad_scalar;      /* xv.shape[1] of the reference call (comb_opt.py:221)       */
    int32_t reserved;
} s2v_graph_t;

/* theta1a..theta4 of QNet (comb_opt.py:204-209) / PPO_GNN_Single_LSTM (models_basic_lstm.py:432-436),
 * row-major nn.Linear layout: weight [out,in], bias [out]. */
typedef struct s2v_embed_params {
    int32_t node_dim;         /* F  */
    int32_t emb_dim;          /* p  */
    int32_t rounds;           /* T  */
    int32_t flags;            /* kernel path: 0 auto (tensor cores when p == 64 and num_nodes >= 4096: the fp16x2
                                 kernels of flag 4), 1 CUDA-core FP32 (FFMA), 2 tcgen05 with the 8-warp 3xTF32 backward
                                 round, 3 tcgen05 with the warp-specialised bf16x3 backward round and last stage,
                                 4 tcgen05 with the forward-shaped backward round and last stage (two fp16 pieces per
                                 operand, power-of-two tile scales) at every size; 2, 3 and 4 need p == 64       */
    const float* theta1a_w;   /* [p,F] */
    const float* theta1a_b;   /* [p]   */
    const float* theta1b_w;   /* [p,p] */
    const float* theta1b_b;
    const float* theta2_w;    /* [p,p] */
    const float* theta2_b;
    const float* theta3_w;    /* [p,p] */
    const float* theta3_b;
    const float* theta4_w;    /* [p,1] */
    const float* theta4_b;
} s2v_embed_params_t;

/* theta5..theta7 (comb_opt.py:210-212; theta{5,6,7}_{pi,v}, models_basic_lstm.py:449-460). */
typedef struct s2v_head_params {
    int32_t emb_dim;
    int32_t norm_agg;         /* 1 = mean pool (scatter_mean), 0 = sum pool (scatter_add)  */
    const float* theta5_w;    /* [1,2p] */
    const float* theta5_b;    /* [1]    */
    const float* theta6_w;    /* [p,p]  */
    const float* theta6_b;
    const float* theta7_w;    /* [p,p]  */
    const float* theta7_b;
    int32_t flags;            /* kernel path, as s2v_embed_params_t.flags (ABI 2)            */
    int32_t reserved;
} s2v_head_params_t;

int         s2v_abi_version(void);
const char* s2v_error_string(int code);

/* Number of floats in the flat gradient buffers (parameter order of the reference
 * constructors, so the concatenation [embed | head] is QNet.parameters() order). */
int64_t s2v_embed_grad_floats(int32_t node_dim, int32_t emb_dim);   /* theta1a.w,b 1b.w,b 2.w,b 3.w,b 4.w,b */
int64_t s2v_head_grad_floats(int32_t emb_dim);                      /* theta5.w,b 6.w,b 7.w,b               */

/* ---------------------------------------------------------------------------------
 * K1  deterministic CSR / CSC builder over a (batched) edge_index.
 * Replaces: dense W = nx.to_numpy_matrix / F.pad / torch.stack(Ws).to(device)
 * (simdata_utils.py:513, comb_opt.py:538,342), dense_to_sparse (simdata_utils.py:519),
 * to_dense_adj (models_basic_lstm.py:515).
 * edge_index: int64 [2,E] (row 0 = gathering node i, row 1 = column j), ids in [0,n).
 * Output order inside a row / column is the input edge order (stable), so the result
 * is bit-identical to a stable argsort (oracle/s2v_ref.py:csr_from_edge_index).
 * workspace: s2v_csr_workspace_bytes(n, E) bytes.
 * --------------------------------------------------------------------------------- */
int64_t s2v_csr_workspace_bytes(int64_t n, int64_t num_edges);
int s2v_csr_build(const int64_t* edge_index, int64_t num_edges, int64_t n,
                  int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t, int32_t* indeg,
                  void* workspace, int64_t workspace_bytes, void* stream);
/* Same builder over the edge list as a trainer holds it on the host: int32 or int64 ids (index_bytes 4 | 8) that are
 * either global (eptr == nptr == NULL) or LOCAL to their graph -- then eptr[B+1] is the first edge and nptr[B+1] the
 * first node of every graph, and the builder adds nptr[g] to both endpoints of the edges of graph g on the device.
 * Replaces the offset pass of Batch.from_data_list (comb_opt.py:299,344) and halves the bytes a batch ships
 * (int32 local ids instead of int64 global ids).  Output identical to s2v_csr_build on the offset list. */
int s2v_csr_build_local(const void* edge_index, int32_t index_bytes, int64_t num_edges, int64_t n,
                        const int32_t* eptr, const int32_t* nptr, int32_t num_graphs,
                        int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t, int32_t* indeg,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* Optional: row_to_col[k] = position in col_t of the edge stored at position k of col (the GATv2 backward walks both
 * forms and exchanges per-edge values).  build_workspace: the workspace of the s2v_csr_build call that produced the
 * CSR, untouched since; scratch: [num_edges] int32. */
int s2v_csr_row_to_col(const void* build_workspace, int64_t n, int64_t num_edges, int32_t* scratch, int32_t* row_to_col,
                       void* stream);

/* ---------------------------------------------------------------------------------
 * K2+K3  embedding forward: node MLP, theta4/theta3 degree term, T propagation rounds.
 * Replaces comb_opt.py:230-242 and models_basic_lstm.py:491-503.
 *   x    [num_nodes,F]      node features of the REAL nodes (padded rows never enter)
 *   base [num_nodes,p] out  s1 + s3 + b2 (the round-invariant part of the pre-activation)
 *   mus  [T,num_nodes,p] out  mu^1..mu^T (mu^T = mus[T-1] is the embedding; all rounds
 *                            are kept for backward)
 * --------------------------------------------------------------------------------- */
int s2v_embed_forward(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                      float* base, float* mus, void* stream);

/* Embedding backward (autograd of the lines above, comb_opt.py:362 / models_basic_lstm.py:650).
 *   dmu        [num_nodes,p]   gradient w.r.t. mu^T
 *   dmu_masked 1 if dmu is already multiplied by [mu^T > 0] (fused head backward)
 *   dus        [T,num_nodes,p] workspace for the pre-activation gradients du^t
 *   workspace  s2v_embed_backward_workspace_bytes() bytes (per-CTA partial sums)
 *   grad       [s2v_embed_grad_floats] out, overwritten                                  */
int64_t s2v_embed_backward_workspace_bytes(int32_t node_dim, int32_t emb_dim);
int s2v_embed_backward(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                       const float* base, const float* mus, const float* dmu, int32_t dmu_masked,
                       float* dus, void* workspace, int64_t workspace_bytes, float* grad, void* stream);

/* The same with the gradient w.r.t. the node features: dx [num_nodes, F] out (NULL = not wanted).  Needed when a
 * trainable module feeds x: lstm_type 'FE' runs an LSTM over the node features before the GNN
 * (models_basic_lstm.py:442-444,473).  dx = dh theta1a with dh = (sum_t du^t theta1b) * [h > 0]; the last backward
 * stage then runs on the CUDA cores (the tcgen05 variant does not keep dh). */
int s2v_embed_backward_dx(const s2v_graph_t* g, const s2v_embed_params_t* prm, const float* x,
                          const float* base, const float* mus, const float* dmu, int32_t dmu_masked,
                          float* dus, void* workspace, int64_t workspace_bytes, float* grad, float* dx, void* stream);

/* One propagation round alone (comb_opt.py:241-242): mu_next = relu(base + theta2 (W mu_prev)),
 * and one backward round: g = W^T du_t ; du_prev = (theta2^T g) * [mu_prev > 0] (the per-CTA
 * partial sums of dtheta2 land in workspace).  Used by unit tests and to time the dominant
 * kernels in isolation. */
int s2v_embed_round_forward(const s2v_graph_t* g, int32_t emb_dim, int32_t flags, const float* theta2_w,
                            const float* base, const float* mu_prev, float* mu_next, void* stream);
int s2v_embed_round_backward(const s2v_graph_t* g, int32_t node_dim, int32_t emb_dim, int32_t flags,
                             const float* theta2_w, const float* du_t, const float* mu_prev, float* du_prev,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* Self-test of the tcgen05 building blocks on one CTA (rows x 64 operands, fp32):
 * mode 0: out[rows,64] = A W^T ; mode 1: out[rows,64] = A W ; mode 2: out[64,64] = A^T B (B [rows,64]);
 * mode 3: out[rows,64] = A W with the A operand in tensor memory. */
int s2v_debug_tc_gemm(int32_t mode, const float* A, const float* B, float* out, int64_t rows, void* stream);

/* ---------------------------------------------------------------------------------
 * K5  pooling + Q / logit head.  Replaces comb_opt.py:248-273 (select list, scatter_mean /
 * scatter_add, repeat_interleave, theta6, theta7, cat, relu, theta5) and
 * models_basic_lstm.py:530-538, 548-562.
 *   mu   [num_nodes,p]
 *   pool [B,p] out, gstate [B,p] out (theta6 pool + b6), q [num_nodes] out
 * --------------------------------------------------------------------------------- */
int s2v_head_forward(const s2v_graph_t* g, const s2v_head_params_t* prm, const float* mu,
                     float* pool, float* gstate, float* q, void* stream);

/* The same head with the invalid-action mask and the per-graph reduction in ONE launch (one CTA per graph: pool,
 * theta6, the graph's q rows, then the masked reduction over them).  Outputs pool / gstate / q as s2v_head_forward, plus
 *   mode 0: nothing else (QNet.forward, comb_opt.py:248-273)
 *   mode 1: prob [num_nodes] = softmax of q over mask != 0, masked entries exactly 0 (models_basic_lstm.py:530-540)
 *   mode 2: arg_out [B] (local node id, -1 if the mask is empty), val_out [B] = masked max (models_basic_lstm.py:548-564)
 *   mode 3: arg_out [B] = position in the neighbour list, node_out [B], val_out [B] (comb_opt.py:319-326)
 * Bit-identical to s2v_head_forward followed by the s2v_masked_* / s2v_list_argmax call whenever
 * s2v_head_fused_eligible() is 1: the problem runs the CUDA-core head (emb_dim != 64, flags == 1, or fewer than 4096
 * rows) and its graphs are small (<= 256 nodes on average: one CTA walks a whole graph).  Large batches keep the tcgen05
 * head + separate reduction (launches do not matter there), a few large graphs the tile-parallel CUDA-core head. */
int s2v_head_fused_forward(const s2v_graph_t* g, const s2v_head_params_t* prm, const float* mu, int32_t mode,
                           const uint8_t* mask, const int32_t* reach_ptr, const int32_t* reach_idx,
                           float* pool, float* gstate, float* q, float* prob, int64_t* arg_out, int64_t* node_out,
                           float* val_out, void* stream);
int s2v_head_fused_eligible(int32_t emb_dim, int32_t flags, int64_t num_nodes, int64_t num_graphs);

/* Small-graph fast path: the whole forward pass -- node MLP, T rounds, pool, head, mask, reduction -- in ONE cooperative
 * launch with grid-wide barriers between the stages (csrc/s2v_small.cu).  For acting (B = 1), the reference's
 * Manhattan5 / MemTask batches and the per-rollout PPO calls, where a stage is microseconds and T + 3 launches are the
 * cost.  Same stage code as the per-stage CUDA-core kernels: results bit-identical to s2v_embed_forward followed by
 * s2v_head_fused_forward (graphs of at most 256 nodes on average: one CTA walks a graph) or by s2v_head_forward and the
 * mask kernels (larger graphs, e.g. acting on a 975-node road network: the head stage is tile-parallel -- pool + theta6
 * per graph | q per row tile | mask + reduction per graph -- with two more grid barriers).  mode -1: embedding only
 * (hprm and the head outputs may be NULL); 0..3 as s2v_head_fused_forward.  s2v_qnet_small_eligible(): 1 when the
 * problem runs the CUDA-core path and has at most 65 536 rows; otherwise s2v_qnet_small_forward returns
 * S2V_ERR_UNSUPPORTED.
 * Replaces comb_opt.py:297-331 (predict / get_best_action) and models_basic_lstm.py:526-566 (pi / v per rollout). */
int s2v_qnet_small_eligible(int32_t emb_dim, int32_t flags, int64_t num_nodes, int64_t num_graphs);
/* ABI 5: the same launch with a caller-owned workspace of s2v_qnet_small_workspace_floats() floats, which lets MANY CTAs
 * pool the rows of a large graph (one [row group][p] partial per 64-row chunk, summed per thread in chunk order: the same
 * association, hence the same bits, as one CTA walking the graph) -- the pool stage of B = 1 on a 975-node network drops
 * from 11 us to ~4 us.  workspace NULL or too small, or graphs of at most 256 nodes: exactly s2v_qnet_small_forward. */
int64_t s2v_qnet_small_workspace_floats(int32_t emb_dim, int64_t num_nodes, int64_t num_graphs);
int s2v_qnet_small_forward_ws(const s2v_graph_t* g, const s2v_embed_params_t* eprm, const s2v_head_params_t* hprm,
                              const float* x, float* base, float* mus, int32_t mode, const uint8_t* mask,
                              const int32_t* reach_ptr, const int32_t* reach_idx, float* pool, float* gstate, float* q,
                              float* prob, int64_t* arg_out, int64_t* node_out, float* val_out, float* workspace,
                              int64_t workspace_floats, void* stream);
int s2v_qnet_small_forward(const s2v_graph_t* g, const s2v_embed_params_t* eprm, const s2v_head_params_t* hprm,
                           const float* x, float* base, float* mus, int32_t mode, const uint8_t* mask,
                           const int32_t* reach_ptr, const int32_t* reach_idx, float* pool, float* gstate, float* q,
                           float* prob, int64_t* arg_out, int64_t* node_out, float* val_out, void* stream);

/* Head backward.  dq [num_nodes]; dmu [num_nodes,p] out (multiplied by [mu>0] when
 * mask_relu != 0, i.e. it is du^T); per_graph [B, 2p+4] workspace;
 * grad [s2v_head_grad_floats] out, overwritten. */
int64_t s2v_head_backward_workspace_bytes(int32_t emb_dim, int32_t num_graphs);
int s2v_head_backward(const s2v_graph_t* g, const s2v_head_params_t* prm, const float* mu,
                      const float* pool, const float* gstate, const float* dq, int32_t mask_relu,
                      float* dmu, void* workspace, int64_t workspace_bytes, float* grad, void* stream);

/* Per-graph mean / sum pool alone (critic 'v': theta5_v(mu.mean(dim=1)), models_basic_lstm.py:548-552)
 * and its backward (dmu[i] = dpool[g(i)] / n_g). */
int s2v_pool_forward(const s2v_graph_t* g, int32_t emb_dim, int32_t norm_agg, const float* mu, float* pool, void* stream);
int s2v_pool_backward(const s2v_graph_t* g, int32_t emb_dim, int32_t norm_agg, const float* dpool, float* dmu, void* stream);

/* du = dmu * [mu > 0]  (ReLU backward of the last round when the head is not fused). */
int s2v_relu_mask(const float* dmu, const float* mu, float* du, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------
 * K6/K7  invalid-action mask + per-graph reductions (bit-exact integer results).
 *   s2v_list_argmax : DQN greedy action, comb_opt.py:319-326: np.argmax over
 *       q[reachable_nodes]; reach_idx holds LOCAL node ids per graph (ascending neighbour
 *       list, simdata_utils.py:641), reach_ptr [B+1].  Outputs per graph: position in the
 *       list (first maximum), local node id, value.  Empty list -> (-1,-1,-inf).
 *   s2v_masked_argmax : first maximum of q over mask != 0 (qvals[~reachable] = -inf;
 *       max(dim=1), models_basic_lstm.py:563-564; probs.argmax(), rl_policy.py:536).
 *       Empty mask -> arg -1, val -inf.  arg is the LOCAL node id.
 *   s2v_masked_softmax_* : prob_logits[~reachable] = -inf; softmax(dim=1)
 *       (models_basic_lstm.py:539-540); masked entries are exactly 0.
 *   s2v_scatter_rows : dq = 0; dq[ptr[g] + arg[g]] = dval[g]  (backward of the masked max).
 * --------------------------------------------------------------------------------- */
int s2v_list_argmax(const s2v_graph_t* g, const float* q, const int32_t* reach_ptr, const int32_t* reach_idx,
                    int64_t* pos_out, int64_t* node_out, float* val_out, void* stream);
int s2v_masked_argmax(const s2v_graph_t* g, const float* q, const uint8_t* mask,
                      int64_t* arg_out, float* val_out, void* stream);
int s2v_masked_softmax_forward(const s2v_graph_t* g, const float* logits, const uint8_t* mask,
                               float* prob, void* stream);
int s2v_masked_softmax_backward(const s2v_graph_t* g, const float* prob, const float* dprob,
                                float* dlogits, void* stream);
int s2v_scatter_rows(const s2v_graph_t* g, const int64_t* arg, const float* dval, float* dq, void* stream);

/* ---------------------------------------------------------------------------------
 * GATv2 edge-attention layer (ABI 3; SURVEY.md 8f rank 1, --qnet gat2).
 * Replaces GATv2Conv.propagate/message + torch_geometric.utils.softmax + the scatter-add
 * aggregation of the third-party PyG layer behind self.gat(pyg_data.x, pyg_data.edge_index)
 * (comb_opt.py:48-62,94,163-171; models_basic_lstm.py:326-338,414-426,509-513).
 * The two node-wise linear layers x_l = lin_l(x), x_r = lin_r(x) are plain GEMMs and stay
 * with the caller (one cuBLAS GEMM over the concatenated weights).
 *   g        layout built from the edge_index WITHOUT self loops (remove_self_loops);
 *            the kernels append the self edge of every node last (add_self_loops order).
 *            Message direction: source edge_index[0] -> target edge_index[1], i.e. target j
 *            aggregates the rows listed in the transposed CSR (rowptr_t / col_t).
 *   heads H, channels C (per head); H*C <= 128
 *   xl, xr   [num_nodes, H*C] with row stride ld floats (two halves of one [n, 2HC] GEMM output)
 *   att [H*C], bias [H*C]
 *   lin_bias [2*H*C] = lin_l.bias | lin_r.bias, or NULL.  When given, xl / xr are the BIAS-FREE products x W_l^T and
 *            x W_r^T (the caller's GEMM then needs no bias epilogue pass over [n, 2HC]) and the kernels add b_l / b_r
 *            to every x_l / x_r row as it is loaded -- the same fp32 add the epilogue would have done, so results are
 *            identical to passing biased rows.  NULL: xl / xr already contain their biases.
 *   out      [num_nodes, H*C] = sum_j softmax_j(att . lrelu(xr[i] + xl[j], 0.2)) xl[j] + bias,
 *            ReLU applied when apply_relu != 0 (BasicGNN between layers)
 *   What the forward keeps for the backward depends on the kernel variant (s2v_gat_uses_edge_alpha):
 *   alpha    [s2v_gat_edge_floats] out (specialised shapes 4x32, 2x32, 4x16): the attention weight of every edge and
 *            head -- the incoming edges of target r in col_t order, copy-major for replicated layouts, then the self
 *            edges [num_nodes, H].  Backward then recomputes neither scores nor softmax: the target pass stores
 *            de = alpha (dalpha - D) per edge next to alpha and the source pass fetches both through row_to_col
 *            (s2v_csr_row_to_col).  Pass NULL to force the generic kernels.
 *   stats    [num_nodes, H, 4] out (generic kernels; may be NULL when alpha is given and the shape is specialised):
 *            per target and head (max score, 1 / den, D, den = softmax denominator + 1e-16); the backward writes
 *            D = sum_j a_ij da_ij into the third slot and recomputes alpha from it
 * Backward: dout [num_nodes, H*C] (already multiplied by the ReLU mask when the layer had one)
 *   -> dxl, dxr (row stride ld), grad_att [H*C], grad_bias [H*C], grad_lin_bias [2*H*C] = column sums of
 *      dxl | dxr, i.e. the gradients of lin_l.bias | lin_r.bias (all overwritten; fixed-order sums).
 * --------------------------------------------------------------------------------- */
#define S2V_GAT_MAX_GRID 2048
int s2v_gat_attn_forward(const s2v_graph_t* g, int32_t heads, int32_t channels, int32_t ld,
                         const float* xl, const float* xr, const float* att, const float* bias,
                         const float* lin_bias, int32_t apply_relu, float* out, float* stats, float* alpha, void* stream);
int64_t s2v_gat_edge_floats(const s2v_graph_t* g, int32_t heads);
int s2v_gat_uses_edge_alpha(int32_t heads, int32_t channels, int32_t ld);
int64_t s2v_gat_attn_backward_workspace_bytes(int32_t heads, int32_t channels, int64_t edge_floats);
int s2v_gat_attn_backward(const s2v_graph_t* g, int32_t heads, int32_t channels, int32_t ld,
                          const float* xl, const float* xr, const float* att, const float* lin_bias,
                          float* stats, const float* alpha, const int32_t* row_to_col, const float* dout, float* dxl, float* dxr, float* grad_att, float* grad_bias,
                          float* grad_lin_bias, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------
 * Node-wise linear layers of the GATv2 branch on the tensor cores (tcgen05, bf16x3 split precision, fp32
 * accumulate; ~3e-7 of the output scale).  Replace the lin_l / lin_r matmuls inside GATv2Conv.forward and their
 * autograd (torch_geometric GATv2Conv: x_l = self.lin_l(x), x_r = self.lin_r(x); comb_opt.py:94) for feature
 * widths that are multiples of 64.  Rows = nodes; all matrices row-major fp32, leading dimensions in floats.
 *   s2v_rows_gemm64 :  out[n, 64 nb] = A[n, 64 kb] * Wt, optionally * (mask > 0)  (mask [n, 64 nb])
 *        transpose_w = 0: W is [64 nb, 64 kb], out = A W^T      (forward: kb = 1, nb = 2)
 *        transpose_w = 1: W is [64 kb, 64 nb], out = A W        (backward dx: kb = 2, nb = 1; mask = the layer
 *                                                                input, i.e. the ReLU mask of the previous layer)
 *        supported (kb, nb, transpose_w): (1, 2, 0), (2, 1, 1); anything else returns S2V_ERR_UNSUPPORTED
 *   s2v_rows_reduce64 : out[64 ab, 64] = A[n, 64 ab]^T B[n, 64]  (weight gradient; ab = 2), per-CTA partials in
 *        workspace summed in fixed order.
 * --------------------------------------------------------------------------------- */
/* General form (ABI 4): every lin_l / lin_r of the configurations the reference ships -- QNet_GATv2 (in 7 | 64, 2 x 32) and
 * PPO_GNN_Single_LSTM(qnet='gat2') (in 7 | 128, hidden 128 = 4 x 32, out 64 = 4 x 16; models_basic_lstm.py:418-422,
 * ppo_custom.py:256-257) -- runs on the tcgen05 kernels, no library GEMM left:
 *   s2v_rows_linear : out[n, m] = A[n, k] W^T (W [m, k] row stride ldw, transpose_w = 0) or A[n, k] W (W [k, m],
 *        transpose_w = 1), optionally * (mask > 0); k <= 256 of any value (the F = 7 input layer is zero padded), m a
 *        multiple of 64 <= 256.  Planned as passes of <= 128 output x <= 128 contraction columns; K passes accumulate.
 *   s2v_rows_outer  : out[m, k] = A[n, m]^T B[n, k] (dW = dY^T X), m a multiple of 64 <= 256, k <= 128 (ragged allowed);
 *        workspace as s2v_rows_reduce64. */
int s2v_rows_linear(const float* A, int32_t lda, int32_t k, const float* W, int32_t ldw, int32_t m, int32_t transpose_w,
                    const float* mask, int32_t ldm, float* out, int32_t ldo, int64_t n, void* stream);
int s2v_rows_outer(const float* A, int32_t lda, int32_t m, const float* B, int32_t ldb, int32_t k, float* out, int32_t ldo,
                   void* workspace, int64_t workspace_bytes, int64_t n, void* stream);
int s2v_rows_gemm64(const float* A, int32_t lda, int32_t k_blocks, const float* W, int32_t n_blocks, int32_t transpose_w,
                    const float* mask, int32_t ldm, float* out, int32_t ldo, int64_t n, void* stream);
int64_t s2v_rows_reduce64_workspace_bytes(int32_t a_blocks);
int s2v_rows_reduce64(const float* A, int32_t lda, int32_t a_blocks, const float* B, int32_t ldb, float* out,
                      void* workspace, int64_t workspace_bytes, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2V_B200_H */
