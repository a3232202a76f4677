// K1: deterministic CSR / CSC builder over a (batched) int64 edge_index.
// Integer-only.  Result is bit-identical to a stable sort of the edges by row
// (resp. column): inside a row the entries keep the input edge order
// (oracle/s2v_ref.py:csr_from_edge_index).  Replaces the dense adjacency the
// reference ships to the device (comb_opt.py:342,538; models_basic_lstm.py:515).
//
// histogram (integer atomics: order-independent counts) -> exclusive scan ->
// bucket fill with edge ids (order inside a bucket arbitrary) -> per-row sort of the
// edge ids (restores input order => deterministic) -> column ids.
#include "s2v_common.cuh"

#define SCAN_BLOCK 1024

// Edge list as the caller holds it: int64 (torch / PyG) or int32 ids, global (batch-wide) or LOCAL to their graph.
// Local ids come with eptr[B+1] (first edge of every graph) and nptr[B+1] (first node): the node offset of edge e
// is nptr[g] for the g with eptr[g] <= e < eptr[g+1] (binary search; both arrays are a few KB and stay in L1 / L2).
// Shipping int32 local ids halves the host-to-device bytes of a batch and needs no offset pass on the host.
struct EdgeList {
    const void* ei;          // [2, E]
    int64_t E;
    const int32_t* eptr;     // NULL: ids are global
    const int32_t* nptr;
    int32_t B;
    int32_t is32;
};
__device__ __forceinline__ int64_t edge_offset(const EdgeList& L, int64_t e) {
    if (!L.eptr) return 0;
    int lo = 0, hi = L.B;                       // invariant: eptr[lo] <= e < eptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(L.eptr + mid) <= e) lo = mid; else hi = mid;
    }
    return __ldg(L.nptr + lo);
}
__device__ __forceinline__ int64_t edge_raw(const EdgeList& L, int which, int64_t e) {
    return L.is32 ? (int64_t)__ldg((const int32_t*)L.ei + which * L.E + e) : __ldg((const int64_t*)L.ei + which * L.E + e);
}
// local ids must also stay inside their own graph: [0, nptr[g+1] - nptr[g]) -- checked through the global range of the
// batch only (an id past its graph lands in a later graph of the batch; the caller validates per-graph sizes)

__global__ void k_csr_count(EdgeList L, int64_t n, int32_t* __restrict__ cnt_row,
                            int32_t* __restrict__ cnt_col, int32_t* __restrict__ err) {
    const int64_t E = L.E;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = edge_offset(L, e);
        int64_t s = edge_raw(L, 0, e) + off, d = edge_raw(L, 1, e) + off;
        if (s < 0 || s >= n || d < 0 || d >= n) { atomicOr(err, 1); continue; }
        atomicAdd(cnt_row + s, 1);
        atomicAdd(cnt_col + d, 1);
    }
}

// out[i] = exclusive prefix inside the block, block_sums[b] = block total
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_local(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t* __restrict__ block_sums, int64_t n) {
    __shared__ int32_t warp_tot[SCAN_BLOCK / 32];
    int64_t i = blockIdx.x * (int64_t)SCAN_BLOCK + threadIdx.x;
    int32_t v = i < n ? in[i] : 0;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    int32_t incl = x + (warp > 0 ? warp_tot[warp - 1] : 0);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}

// single block: exclusive scan of block_sums in place; writes the grand total to *total
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_blocks(int32_t* __restrict__ block_sums, int nb, int32_t* __restrict__ total) {
    __shared__ int32_t warp_tot[SCAN_BLOCK / 32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += SCAN_BLOCK) {
        int i = base + threadIdx.x;
        int32_t v = i < nb ? block_sums[i] : 0;
        int32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        int32_t incl = x + (warp > 0 ? warp_tot[warp - 1] : 0) + carry_s;
        if (i < nb) block_sums[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_add(int32_t* __restrict__ out, const int32_t* __restrict__ block_sums, const int32_t* __restrict__ total, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
    if (i == 0) out[n] = *total;
}

__global__ void k_csr_fill(EdgeList L, int64_t n, const int32_t* __restrict__ rowptr,
                           const int32_t* __restrict__ rowptr_t, int32_t* __restrict__ cur_row,
                           int32_t* __restrict__ cur_col, int32_t* __restrict__ eid_row, int32_t* __restrict__ eid_col) {
    const int64_t E = L.E;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = edge_offset(L, e);
        int64_t s = edge_raw(L, 0, e) + off, d = edge_raw(L, 1, e) + off;
        if (s < 0 || s >= n || d < 0 || d >= n) continue;
        eid_row[rowptr[s] + atomicAdd(cur_row + s, 1)] = (int32_t)e;
        eid_col[rowptr_t[d] + atomicAdd(cur_col + d, 1)] = (int32_t)e;
    }
}

// One thread per row: sort the edge ids of the bucket ascending (restores the input order), then emit the other
// endpoint.  which = 1: emit ei[1][e] (row form), 0: ei[0][e].  Rows with more than CSR_HEAVY entries (hubs, star
// graphs) are left to k_csr_sort_emit_heavy: an insertion sort by one thread is quadratic in the degree.
// Duplicate edges (the same (i, j) pair twice) are flagged in err bit 1: the reference builds W with to_dense_adj,
// which SUMS duplicates into a weight of 2 (theta4 term relu(2 w4 + b4)), while a CSR would count the edge twice
// (2 relu(w4 + b4)) -- the layout builders reject such input instead of diverging silently.
#define CSR_HEAVY 32

__global__ void k_csr_sort_emit(EdgeList L, int64_t n, const int32_t* __restrict__ ptr,
                                int32_t* __restrict__ eid, int32_t* __restrict__ out, int which, int32_t* __restrict__ err,
                                int32_t* __restrict__ heavy_count, int32_t* __restrict__ heavy_list) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int beg = ptr[i], end = ptr[i + 1];
        if (end - beg > CSR_HEAVY) {            // left to k_csr_sort_emit_heavy (list order is irrelevant: one CTA per row)
            heavy_list[atomicAdd(heavy_count, 1)] = (int32_t)i;
            continue;
        }
        for (int a = beg + 1; a < end; ++a) {
            int32_t key = eid[a];
            int b = a - 1;
            while (b >= beg && eid[b] > key) { eid[b + 1] = eid[b]; --b; }
            eid[b + 1] = key;
        }
        for (int a = beg; a < end; ++a) {
            const int64_t e = eid[a];
            out[a] = (int32_t)(edge_raw(L, which, e) + edge_offset(L, e));
        }
        if (which) {
            bool dup = false;
            for (int a = beg + 1; a < end; ++a)
                for (int b = beg; b < a; ++b) dup |= out[a] == out[b];
            if (dup) atomicOr(err, 2);
        }
    }
}

// Heavy rows: one CTA per row, bitonic sort of the edge ids in a power-of-two scratch segment (padded with INT_MAX),
// O(d log^2 d / threads).  scratch: 2 * E ints; row i uses [2 * ptr[i], 2 * ptr[i] + n2).  Duplicate detection sorts
// the emitted endpoints the same way and compares neighbours.
__device__ void cta_bitonic_sort(int32_t* a, int n2) {
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const int32_t x = a[i], y = a[l];
                    if ((x > y) == ((i & k) == 0)) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(256)
k_csr_sort_emit_heavy(EdgeList L, const int32_t* __restrict__ ptr, int32_t* __restrict__ eid,
                      int32_t* __restrict__ out, int which, int32_t* __restrict__ scratch, int32_t* __restrict__ err,
                      const int32_t* __restrict__ heavy_count, const int32_t* __restrict__ heavy_list) {
    __shared__ int dup_s;
    const int count = *heavy_count;             // rows listed by k_csr_sort_emit (usually none: the kernel exits at once)
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const int64_t i = heavy_list[idx];
        const int beg = ptr[i], end = ptr[i + 1], d = end - beg;
        int n2 = 1;
        while (n2 < d) n2 <<= 1;
        int32_t* sc = scratch + 2 * (int64_t)beg;
        for (int k = threadIdx.x; k < n2; k += blockDim.x) sc[k] = k < d ? eid[beg + k] : INT32_MAX;
        if (threadIdx.x == 0) dup_s = 0;
        __syncthreads();
        cta_bitonic_sort(sc, n2);
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
            const int64_t e = sc[k];
            eid[beg + k] = (int32_t)e;
            out[beg + k] = (int32_t)(edge_raw(L, which, e) + edge_offset(L, e));
        }
        __syncthreads();
        if (which) {
            for (int k = threadIdx.x; k < n2; k += blockDim.x) sc[k] = k < d ? out[beg + k] : INT32_MAX;
            __syncthreads();
            cta_bitonic_sort(sc, n2);
            int dup = 0;
            for (int k = threadIdx.x + 1; k < d; k += blockDim.x) dup |= sc[k] == sc[k - 1];
            if (dup) dup_s = 1;
            __syncthreads();
            if (threadIdx.x == 0 && dup_s) atomicOr(err, 2);
        }
        __syncthreads();
    }
}

// position of every edge in the column form, then row-form position -> column-form position
__global__ void k_csr_inv(const int32_t* __restrict__ eid_col, int64_t E, int32_t* __restrict__ inv) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) inv[eid_col[p]] = (int32_t)p;
}
__global__ void k_csr_map(const int32_t* __restrict__ eid_row, const int32_t* __restrict__ inv, int64_t E, int32_t* __restrict__ row_to_col) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) row_to_col[p] = inv[eid_row[p]];
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// workspace layout (int32): err[64] | cnt_row[n] | cnt_col[n] | eid_row[E] | eid_col[E] | block_sums[nb+1] x2 | totals[64]
//                           | heavy-row sort scratch [2E] | heavy-row lists [E/32 + 64] x2
// err[0]: bit 0 = node id out of range, bit 1 = duplicate edge; err[1], err[2]: number of heavy rows / columns
extern "C" int64_t s2v_csr_workspace_bytes(int64_t n, int64_t E) {
    int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1;
    return (64 + align_up(n, 64) * 2 + align_up(E, 64) * 2 + align_up(nb, 64) * 2 + 64 + 2 * align_up(E, 64) +
            2 * align_up(E / 32 + 64, 64)) * 4;
}

// Entry k of row i (edge i -> col[k]) is entry row_to_col[k] of column col[k] in the transposed form.  Needs the
// workspace of the s2v_csr_build call that produced the CSR (same n, E, same stream order) plus scratch [E] int32.
extern "C" int s2v_csr_row_to_col(const void* build_workspace, int64_t n, int64_t E, int32_t* scratch, int32_t* row_to_col,
                                  void* stream) {
    if (E < 0 || n < 0 || n >= INT32_MAX || E >= INT32_MAX) return S2V_ERR_BAD_ARG;
    if (E == 0) return 0;
    if (!build_workspace || !scratch || !row_to_col) return S2V_ERR_BAD_ARG;
    const int32_t* w = (const int32_t*)build_workspace;
    const int32_t* eid_row = w + 64 + align_up(n, 64) * 2;
    const int32_t* eid_col = eid_row + align_up(E, 64);
    int eb = (int)((E + 255) / 256);
    if (eb > 148 * 32) eb = 148 * 32;
    cudaStream_t st = (cudaStream_t)stream;
    k_csr_inv<<<eb, 256, 0, st>>>(eid_col, E, scratch);
    k_csr_map<<<eb, 256, 0, st>>>(eid_row, scratch, E, row_to_col);
    return (int)cudaGetLastError();
}

static int scan_exclusive(const int32_t* in, int32_t* out, int32_t* block_sums, int32_t* total, int64_t n, cudaStream_t st) {
    int nb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (nb < 1) nb = 1;
    k_scan_local<<<nb, SCAN_BLOCK, 0, st>>>(in, out, block_sums, n);
    k_scan_blocks<<<1, SCAN_BLOCK, 0, st>>>(block_sums, nb, total);
    k_scan_add<<<nb, SCAN_BLOCK, 0, st>>>(out, block_sums, total, n);
    return (int)cudaGetLastError();
}

static int csr_build_impl(const EdgeList& L, int64_t n,
                          int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t, int32_t* indeg,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    const int64_t E = L.E;
    if (E < 0 || n < 0 || n >= INT32_MAX || E >= INT32_MAX) return S2V_ERR_BAD_ARG;
    if (!rowptr || !rowptr_t || !workspace || (n > 0 && !indeg)) return S2V_ERR_BAD_ARG;
    if (E > 0 && (!L.ei || !col || !col_t)) return S2V_ERR_BAD_ARG;
    if (workspace_bytes < s2v_csr_workspace_bytes(n, E)) return S2V_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* w = (int32_t*)workspace;
    int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1;
    int32_t* err = w;
    int32_t* cnt_row = err + 64;
    int32_t* cnt_col = cnt_row + align_up(n, 64);
    int32_t* eid_row = cnt_col + align_up(n, 64);
    int32_t* eid_col = eid_row + align_up(E, 64);
    int32_t* bs_row = eid_col + align_up(E, 64);
    int32_t* bs_col = bs_row + align_up(nb, 64);
    int32_t* totals = bs_col + align_up(nb, 64);
    int32_t* heavy_scratch = totals + 64;
    int32_t* heavy_rows = heavy_scratch + 2 * align_up(E, 64);
    int32_t* heavy_cols = heavy_rows + align_up(E / 32 + 64, 64);
    S2V_RETURN_IF(cudaMemsetAsync(w, 0, (64 + align_up(n, 64) * 2) * 4, st));
    int eb = (int)((E + 255) / 256);
    if (eb > 148 * 32) eb = 148 * 32;
    if (eb < 1) eb = 1;
    int rb = (int)((n + 127) / 128);
    if (rb > 148 * 32) rb = 148 * 32;
    if (rb < 1) rb = 1;
    if (E > 0) k_csr_count<<<eb, 256, 0, st>>>(L, n, cnt_row, cnt_col, err);
    if (n > 0) S2V_RETURN_IF(cudaMemcpyAsync(indeg, cnt_col, n * 4, cudaMemcpyDeviceToDevice, st));
    S2V_RETURN_IF(scan_exclusive(cnt_row, rowptr, bs_row, totals, n, st));
    S2V_RETURN_IF(scan_exclusive(cnt_col, rowptr_t, bs_col, totals + 1, n, st));
    if (E > 0) {
        S2V_RETURN_IF(cudaMemsetAsync(cnt_row, 0, align_up(n, 64) * 2 * 4, st));
        k_csr_fill<<<eb, 256, 0, st>>>(L, n, rowptr, rowptr_t, cnt_row, cnt_col, eid_row, eid_col);
        k_csr_sort_emit<<<rb, 128, 0, st>>>(L, n, rowptr, eid_row, col, 1, err, err + 1, heavy_rows);
        k_csr_sort_emit<<<rb, 128, 0, st>>>(L, n, rowptr_t, eid_col, col_t, 0, err, err + 2, heavy_cols);
        // hubs: rows / columns with more than CSR_HEAVY entries, one CTA each (the kernels exit at once when there are none)
        k_csr_sort_emit_heavy<<<148, 256, 0, st>>>(L, rowptr, eid_row, col, 1, heavy_scratch, err, err + 1, heavy_rows);
        k_csr_sort_emit_heavy<<<148, 256, 0, st>>>(L, rowptr_t, eid_col, col_t, 0, heavy_scratch, err, err + 2, heavy_cols);
    }
    return (int)cudaGetLastError();
}

extern "C" int s2v_csr_build(const int64_t* edge_index, int64_t E, int64_t n,
                             int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t, int32_t* indeg,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    EdgeList L{edge_index, E, nullptr, nullptr, 0, 0};
    return csr_build_impl(L, n, rowptr, col, rowptr_t, col_t, indeg, workspace, workspace_bytes, stream);
}

extern "C" int s2v_csr_build_local(const void* edge_index, int32_t index_bytes, int64_t E, int64_t n,
                                   const int32_t* eptr, const int32_t* nptr, int32_t num_graphs,
                                   int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t, int32_t* indeg,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    if (index_bytes != 4 && index_bytes != 8) return S2V_ERR_BAD_ARG;
    if ((eptr == nullptr) != (nptr == nullptr) || (eptr && num_graphs < 1)) return S2V_ERR_BAD_ARG;
    EdgeList L{edge_index, E, eptr, nptr, num_graphs, index_bytes == 4};
    return csr_build_impl(L, n, rowptr, col, rowptr_t, col_t, indeg, workspace, workspace_bytes, stream);
}
