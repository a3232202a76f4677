// K6/K7: invalid-action masking with per-graph reductions.  One CTA per graph;
// warp-shuffle reductions with a fixed combine order (deterministic), first-index
// tie break (np.argmax / torch.argmax semantics).
//   DQN greedy action      comb_opt.py:319-326
//   PPO masked softmax     models_basic_lstm.py:539-540
//   PPO masked max         models_basic_lstm.py:563-564 ; greedy eval rl_policy.py:536
#include "s2v_common.cuh"
#include <math_constants.h>

#include "s2v_reduce.cuh"

__global__ void __launch_bounds__(MK_THREADS)
k_list_argmax(s2v_graph_t g, const float* __restrict__ q, const int32_t* __restrict__ reach_ptr,
              const int32_t* __restrict__ reach_idx, int64_t* __restrict__ pos_out, int64_t* __restrict__ node_out,
              float* __restrict__ val_out) {
    __shared__ ArgPair sh[MK_THREADS / 32];
    const int gi = blockIdx.x;
    int lo, hi;
    graph_range(g, gi, lo, hi);
    int beg = reach_ptr[gi], end = reach_ptr[gi + 1];
    ArgPair p; p.v = -CUDART_INF_F; p.i = -1;
    for (int k = beg + threadIdx.x; k < end; k += MK_THREADS) {
        ArgPair c; c.v = q[lo + reach_idx[k]]; c.i = k - beg;
        p = better(p, c);
    }
    p = block_argmax(p, sh);
    if (threadIdx.x == 0) {
        pos_out[gi] = p.i;
        node_out[gi] = p.i >= 0 ? (int64_t)reach_idx[beg + p.i] : -1;
        val_out[gi] = p.i >= 0 ? p.v : -CUDART_INF_F;
    }
}

__global__ void __launch_bounds__(MK_THREADS)
k_masked_argmax(s2v_graph_t g, const float* __restrict__ q, const uint8_t* __restrict__ mask,
                int64_t* __restrict__ arg_out, float* __restrict__ val_out) {
    __shared__ ArgPair sh[MK_THREADS / 32];
    const int gi = blockIdx.x;
    int lo, hi;
    graph_range(g, gi, lo, hi);
    ArgPair p; p.v = -CUDART_INF_F; p.i = -1;
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS) {
        if (mask[r]) { ArgPair c; c.v = q[r]; c.i = r - lo; p = better(p, c); }
    }
    p = block_argmax(p, sh);
    if (threadIdx.x == 0) {
        arg_out[gi] = p.i;
        val_out[gi] = p.i >= 0 ? p.v : -CUDART_INF_F;
    }
}

__global__ void __launch_bounds__(MK_THREADS)
k_masked_softmax(s2v_graph_t g, const float* __restrict__ logits, const uint8_t* __restrict__ mask,
                 float* __restrict__ prob) {
    __shared__ float sh[MK_THREADS / 32];
    const int gi = blockIdx.x;
    int lo, hi;
    graph_range(g, gi, lo, hi);
    float m = -CUDART_INF_F;
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS)
        if (mask[r]) m = fmaxf(m, logits[r]);
    m = block_max(m, sh);
    float s = 0.f;
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS)
        if (mask[r]) s += expf(logits[r] - m);
    s = block_sum(s, sh);
    // all-masked graph: softmax over all -inf is NaN in the reference as well
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS) {
        float v = mask[r] ? expf(logits[r] - m) / s : 0.f;
        if (m == -CUDART_INF_F) v = CUDART_NAN_F;
        prob[r] = v;
    }
}

__global__ void __launch_bounds__(MK_THREADS)
k_masked_softmax_bwd(s2v_graph_t g, const float* __restrict__ prob, const float* __restrict__ dprob,
                     float* __restrict__ dlogits) {
    __shared__ float sh[MK_THREADS / 32];
    const int gi = blockIdx.x;
    int lo, hi;
    graph_range(g, gi, lo, hi);
    float s = 0.f;
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS) s = fmaf(prob[r], dprob[r], s);
    s = block_sum(s, sh);
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS) dlogits[r] = prob[r] * (dprob[r] - s);
}

__global__ void __launch_bounds__(MK_THREADS)
k_scatter_rows(s2v_graph_t g, const int64_t* __restrict__ arg, const float* __restrict__ dval, float* __restrict__ dq) {
    const int gi = blockIdx.x;
    int lo, hi;
    graph_range(g, gi, lo, hi);
    int a = (int)arg[gi];
    float d = dval[gi];
    for (int r = lo + threadIdx.x; r < hi; r += MK_THREADS) dq[r] = (r - lo == a) ? d : 0.f;
}

static int check_graph_m(const s2v_graph_t* g) {
    if (!g || g->num_nodes < 0 || g->num_graphs < 0 || g->period < 0) return S2V_ERR_BAD_ARG;
    if (g->period == 0 && !g->ptr && g->num_graphs > 0) return S2V_ERR_BAD_ARG;
    return 0;
}

extern "C" int s2v_list_argmax(const s2v_graph_t* g, const float* q, const int32_t* reach_ptr, const int32_t* reach_idx,
                               int64_t* pos_out, int64_t* node_out, float* val_out, void* stream) {
    S2V_RETURN_IF(check_graph_m(g));
    if (!q || !reach_ptr || !pos_out || !node_out || !val_out) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    k_list_argmax<<<g->num_graphs, MK_THREADS, 0, (cudaStream_t)stream>>>(*g, q, reach_ptr, reach_idx, pos_out, node_out, val_out);
    return (int)cudaGetLastError();
}

extern "C" int s2v_masked_argmax(const s2v_graph_t* g, const float* q, const uint8_t* mask,
                                 int64_t* arg_out, float* val_out, void* stream) {
    S2V_RETURN_IF(check_graph_m(g));
    if (!q || !mask || !arg_out || !val_out) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    k_masked_argmax<<<g->num_graphs, MK_THREADS, 0, (cudaStream_t)stream>>>(*g, q, mask, arg_out, val_out);
    return (int)cudaGetLastError();
}

extern "C" int s2v_masked_softmax_forward(const s2v_graph_t* g, const float* logits, const uint8_t* mask,
                                          float* prob, void* stream) {
    S2V_RETURN_IF(check_graph_m(g));
    if (!logits || !mask || !prob) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    k_masked_softmax<<<g->num_graphs, MK_THREADS, 0, (cudaStream_t)stream>>>(*g, logits, mask, prob);
    return (int)cudaGetLastError();
}

extern "C" int s2v_masked_softmax_backward(const s2v_graph_t* g, const float* prob, const float* dprob,
                                           float* dlogits, void* stream) {
    S2V_RETURN_IF(check_graph_m(g));
    if (!prob || !dprob || !dlogits) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    k_masked_softmax_bwd<<<g->num_graphs, MK_THREADS, 0, (cudaStream_t)stream>>>(*g, prob, dprob, dlogits);
    return (int)cudaGetLastError();
}

extern "C" int s2v_scatter_rows(const s2v_graph_t* g, const int64_t* arg, const float* dval, float* dq, void* stream) {
    S2V_RETURN_IF(check_graph_m(g));
    if (!arg || !dval || !dq) return S2V_ERR_BAD_ARG;
    if (g->num_graphs == 0) return 0;
    k_scatter_rows<<<g->num_graphs, MK_THREADS, 0, (cudaStream_t)stream>>>(*g, arg, dval, dq);
    return (int)cudaGetLastError();
}

extern "C" int s2v_abi_version(void) { return S2V_ABI_VERSION; }

extern "C" const char* s2v_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case S2V_ERR_BAD_ARG: return "s2v: bad argument (null pointer, negative or inconsistent size)";
        case S2V_ERR_UNSUPPORTED: return "s2v: unsupported emb_dim (need 8,16,24,32,48,64) or node_dim > 32";
        case S2V_ERR_WORKSPACE: return "s2v: workspace too small";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "s2v: unknown error";
    }
}
