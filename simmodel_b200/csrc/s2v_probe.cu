// Gather-pipeline probe (debug entry point, not part of the product path): how fast can one persistent CTA per SM
// pull the neighbour rows of 64-row tiles into a shared-memory ring with ASYNCHRONOUS copies, when nothing else
// competes?  The answer sizes the loader of the warp-specialised round kernels (s2v_embed_tc.cu), which use the
// same three roles:
//   M (1 warp)  : metadata.  rowptr of the next tile by cp.async into shared memory, chunk descriptors (first ring
//                 slot of every row), ring flow control, neighbour ids of the chunk by cp.async into an id ring.
//   L (4 warps) : copy issue.  Edge-parallel: a lane owns an edge id, every LDGSTS moves two 256-byte neighbour rows
//                 (16 lanes x 16 B each); completion arrives on the chunk's mbarrier (cp.async.mbarrier.arrive).
//                 Mode bit 1: the tile's own rows ride along as one cp.async.bulk (UBLKCP) of 16 KB.
//   C (8 warps) : wait for the chunk, sum the landed rows per destination row in CSR order, store the row.
// No role ever waits on a global load in registers: every dependent address comes from shared memory.
//
// out[r] = sum_{j in N(r)} src[j]  (+ own[r] with mode bit 1): checked against torch on the host side.
#include "s2v_common.cuh"
#include "s2v_tc.cuh"

namespace {

constexpr int PR_ROWS = 64;
constexpr int PR_NC = 8;                 // chunk descriptors in flight
constexpr int PR_L = 4, PR_C = 8;        // loader / consumer warps
constexpr int PR_THREADS = 32 * (1 + PR_L + PR_C);
constexpr int PR_CAP_MAX = 384;          // edges per chunk (id ring)

struct PrDesc {
    int slot0, n, n_run0, off0, off1, last, final_chunk, tile, own_slot, rows, need, pad;
    int beg[PR_ROWS + 1];
    int pad2[3];
};
struct PrShared {
    uint64_t meta_full[PR_NC], full[PR_NC], empty[PR_NC];
    PrDesc desc[PR_NC];
    int ids[PR_NC][PR_CAP_MAX];
    int rp[2][PR_ROWS + 4];              // rowptr of the tile's rows: [0..64] + [65] = end of the first run
};

__device__ long long g_probe_prof[16];

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_if(uint32_t dst, const void* src, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(dst), "l"(src), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(PR_THREADS, 1)
k_gather_probe(s2v_graph_t g, const float* __restrict__ src, const float* __restrict__ own, float* __restrict__ dst,
               int num_tiles, int R, int cap, int mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                            // R slots of 256 B
    __shared__ PrShared sh_static;
    PrShared* sh = &sh_static;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const bool with_own = (mode & 2) != 0, bulk_rows = (mode & 1) != 0;
    if (t == 0) {
        for (int i = 0; i < PR_NC; ++i) {
            tc::mbar_init(&sh->meta_full[i], 33);
            tc::mbar_init(&sh->full[i], bulk_rows ? 1 : PR_L * 32 + 1);
            tc::mbar_init(&sh->empty[i], PR_C);
        }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    if (my_tiles == 0) return;
    const uint32_t ring_s = tc::smem_u32(ring);
    const int32_t* __restrict__ rowptr = g.rowptr_t;
    const int32_t* __restrict__ col = g.col_t;
    auto tile_row0 = [&](int k) { return (blockIdx.x + k * (int)gridDim.x) * PR_ROWS; };

    if (warp == 0) {
        // ============================ M: metadata ============================
        // rowptr entries of tile k -> sh->rp[k & 1]: entry r (0..rows) of the run the row belongs to, [65] = end of run 0
        auto fetch_rp = [&](int k) {
            const int row0 = tile_row0(k), rows = min(PR_ROWS, g.num_nodes - row0);
            int* dstp = sh->rp[k & 1];
            int la = row0, first = rows;
            if (g.period > 0) { const int gi = row0 / g.period; la = row0 - gi * g.period; first = min(g.period - la, rows); }
            for (int r = lane; r <= rows; r += 32)
                cp_async4(dstp + r, rowptr + ((r < first || g.period == 0) ? la + r : r - first));
            if (lane == 0) cp_async4(dstp + 65, rowptr + la + first);
        };
        int head = 0, used = 0, tail = 0, cidx = 0;
        long long t_wait = 0, t_all = clock64();
        fetch_rp(0);
        cp_async_commit();
        for (int k = 0; k < my_tiles; ++k) {
            if (k + 1 < my_tiles) fetch_rp(k + 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const int row0 = tile_row0(k), rows = min(PR_ROWS, g.num_nodes - row0);
            const int* rp = sh->rp[k & 1];
            int off0 = 0, first = rows;
            if (g.period > 0) { const int gi = row0 / g.period; off0 = gi * g.period; first = min(g.period - (row0 - off0), rows); }
            const int e0 = rp[0], n0 = rp[65] - e0, n1 = first < rows ? rp[rows] : 0, ne = n0 + n1;
            auto row_beg = [&](int r) -> int { return r >= rows ? ne : (r < first ? rp[r] - e0 : n0 + rp[r]); };
            const int pb0 = row_beg(lane), pb1 = row_beg(lane + 32);
            __syncwarp();                                              // rp[k & 1] is rewritten by fetch_rp(k + 2)
            const int nchunks = ne > 0 ? (ne + cap - 1) / cap : 1;
            for (int ch = 0; ch < nchunks; ++ch) {
                const int lo = ch * cap, hi = min(ne, lo + cap), n = hi - lo;
                const bool last = ch == nchunks - 1;
                const int need = n + ((last && with_own) ? PR_ROWS : 0);
                long long c0 = clock64();
                while (used + need > R || cidx - tail >= PR_NC) {
                    tc::mbar_wait(&sh->empty[tail % PR_NC], (tail / PR_NC) & 1);
                    used -= sh->desc[tail % PR_NC].need;
                    ++tail;
                }
                t_wait += clock64() - c0;
                PrDesc* d = &sh->desc[cidx % PR_NC];
                d->beg[lane] = min(max(pb0 - lo, 0), n);
                d->beg[lane + 32] = min(max(pb1 - lo, 0), n);
                if (lane == 0) {
                    d->beg[PR_ROWS] = n; d->slot0 = head; d->n = n; d->last = last; d->tile = k; d->rows = rows;
                    d->n_run0 = min(max(n0 - lo, 0), n); d->off0 = off0; d->off1 = off0 + g.period;
                    d->final_chunk = last && k == my_tiles - 1;
                    int os = head + n; if (os >= R) os -= R;
                    d->own_slot = os; d->need = need;
                }
                int* idp = sh->ids[cidx % PR_NC];
                for (int te = lo + lane; te < hi; te += 32)
                    cp_async4(idp + te - lo, col + (te < n0 ? e0 + te : te - n0));
                cp_async_arrive_noinc(&sh->meta_full[cidx % PR_NC]);
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->meta_full[cidx % PR_NC]);
                head += need; if (head >= R) head -= R;
                used += need;
                ++cidx;
            }
        }
        if (blockIdx.x == 0 && lane == 0) { g_probe_prof[0] = t_wait; g_probe_prof[3] = clock64() - t_all; }
    } else if (warp <= PR_L) {
        // ============================ L: copy issue ============================
        const unsigned FULL = 0xffffffffu;
        const int lw = warp - 1, c = lane & 15, half = lane >> 4;
        long long t_wait = 0, t_issue = 0;
        for (int cidx = 0;; ++cidx) {
            long long c0 = clock64();
            tc::mbar_wait(&sh->meta_full[cidx % PR_NC], (cidx / PR_NC) & 1);
            long long c1 = clock64();
            t_wait += c1 - c0;
            const PrDesc* d = &sh->desc[cidx % PR_NC];
            const int slot0 = d->slot0, n = d->n, nr0 = d->n_run0, off0 = d->off0, off1 = d->off1;
            const int fin = d->final_chunk;
            uint64_t* full = &sh->full[cidx % PR_NC];
            const int* idp = sh->ids[cidx % PR_NC];
            if (bulk_rows) {
                for (int te = lw * 32 + lane; te < n; te += PR_L * 32) {
                    int slot = slot0 + te; if (slot >= R) slot -= R;
                    const int off = te < nr0 ? off0 : off1;
                    bulk_g2s(ring_s + slot * 256, src + (size_t)(off + idp[te]) * 64, 256, full);
                }
            } else
            for (int base = lw * 32; base < n; base += PR_L * 32) {
                const int idv = base + lane < n ? idp[base + lane] : 0;
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                    const int el = 2 * it + half, te = base + el;
                    const int j = __shfl_sync(FULL, idv, el);
                    int slot = slot0 + te; if (slot >= R) slot -= R;
                    const int off = te < nr0 ? off0 : off1;
                    cp_async16_if(ring_s + slot * 256 + 16 * c, src + (size_t)(off + j) * 64 + 4 * c, te < n);
                }
            }
            if (!bulk_rows) cp_async_arrive_noinc(full);
            if (lw == 0 && lane == 0) {
                const int rows = d->rows;
                if (bulk_rows) {
                    uint32_t tx = n * 256;
                    if (d->last && with_own && rows > 0) {
                        const int os = d->own_slot, row0 = tile_row0(d->tile);
                        const int first = min(rows, R - os);
                        bulk_g2s(ring_s + os * 256, own + (size_t)row0 * 64, first * 256, full);
                        if (first < rows) bulk_g2s(ring_s, own + (size_t)(row0 + first) * 64, (rows - first) * 256, full);
                        tx += rows * 256;
                    }
                    if (tx) mbar_arrive_expect_tx(full, tx); else tc::mbar_arrive(full);
                } else if (d->last && with_own && rows > 0) {       // the tile's own rows: one (or two, at the ring end) bulk copies
                    const int os = d->own_slot, row0 = tile_row0(d->tile);
                    const int first = min(rows, R - os);
                    bulk_g2s(ring_s + os * 256, own + (size_t)row0 * 64, first * 256, full);
                    if (first < rows) bulk_g2s(ring_s, own + (size_t)(row0 + first) * 64, (rows - first) * 256, full);
                    mbar_arrive_expect_tx(full, rows * 256);
                } else {
                    tc::mbar_arrive(full);
                }
            }
            t_issue += clock64() - c1;
            if (fin) break;
        }
        if (blockIdx.x == 0 && lane == 0 && lw == 0) { g_probe_prof[4] = t_wait; g_probe_prof[6] = t_issue; }
    } else {
        // ============================ C: consumer ============================
        const int cw = warp - 1 - PR_L, c = lane & 15, hw = 2 * cw + (lane >> 4);
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = f4zero();
        long long t_wait = 0, t_sum = 0, t_store = 0, t_desc = 0, t_all = clock64();
        for (int cidx = 0;; ++cidx) {
            long long c0 = clock64();
            tc::mbar_wait(&sh->full[cidx % PR_NC], (cidx / PR_NC) & 1);
            long long c1 = clock64();
            t_wait += c1 - c0;
            const PrDesc* d = &sh->desc[cidx % PR_NC];
            const int slot0 = d->slot0, last = d->last, fin = d->final_chunk;
            int sb[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) sb[i] = d->beg[4 * hw + i];
            long long c1b = clock64() + (sb[0] & 0) + (slot0 & 0);
            t_desc += c1b - c1;
            auto slot_ptr = [&](int s) -> const float* {
                int slot = slot0 + s; if (slot >= R) slot -= R;
                return reinterpret_cast<const float*>(ring + slot * 256) + 4 * c;
            };
            {   // first four neighbours of the four rows: 16 independent shared-memory loads, summed in CSR order
                float4 v[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[i][q] = ld4(slot_ptr(sb[i] + q < sb[i + 1] ? sb[i] + q : 0));   // unconditional: batched
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (sb[i] + q < sb[i + 1]) acc[i] = f4add(acc[i], v[i][q]);
                    for (int s = sb[i] + 4; s < sb[i + 1]; ++s) acc[i] = f4add(acc[i], ld4(slot_ptr(s)));
                }
            }
            long long c2 = clock64();
            t_sum += c2 - c1;
            if (last) {
                const int row0 = tile_row0(d->tile), rows = d->rows, os = d->own_slot;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 4 * hw + i;
                    int slot = os + r; if (slot >= R) slot -= R;
                    if (with_own && r < rows) acc[i] = f4add(acc[i], ld4(reinterpret_cast<const float*>(ring + slot * 256) + 4 * c));
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->empty[cidx % PR_NC]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 4 * hw + i;
                    if (r < rows) st4(dst + (size_t)(row0 + r) * 64 + 4 * c, acc[i]);
                    acc[i] = f4zero();
                }
                t_store += clock64() - c2;
            } else {
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->empty[cidx % PR_NC]);
            }
            if (fin) break;
        }
        if (blockIdx.x == 0 && lane == 0 && cw == 0) {
            g_probe_prof[8] = t_wait; g_probe_prof[9] = clock64() - t_all; g_probe_prof[10] = my_tiles;
            g_probe_prof[12] = t_sum; g_probe_prof[13] = t_store; g_probe_prof[14] = t_desc;
        }
    }
}

}  // namespace

// mode bit 1: own rows through the ring (one bulk copy per tile).  ring_slots: 256-byte slots of the ring (<= 800),
// cap: edges per chunk (<= 384).
extern "C" int s2v_debug_gather_probe(const s2v_graph_t* g, const float* src, const float* own, float* dst, int32_t mode,
                                      int32_t ring_slots, int32_t cap, void* stream) {
    if (!g || !src || !own || !dst || ring_slots < 128 || ring_slots > 800 || cap < 32 || cap > PR_CAP_MAX ||
        cap + PR_ROWS > ring_slots || (g->period > 0 && g->period < PR_ROWS))
        return S2V_ERR_BAD_ARG;
    const int tiles = (g->num_nodes + PR_ROWS - 1) / PR_ROWS;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms < tiles ? sms : (tiles < 1 ? 1 : tiles);
    const int smem = ring_slots * 256 + 1024;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_gather_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_gather_probe<<<grid, PR_THREADS, smem, (cudaStream_t)stream>>>(*g, src, own, dst, tiles, ring_slots, cap, mode);
    return (int)cudaGetLastError();
}

extern "C" int s2v_debug_gather_probe_prof(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_probe_prof, sizeof(long long) * 16);
}

// ---- cp.async.bulk issue / service rate (debug): `warps` warps of one CTA per SM, lane 0 of each issues `n` bulk copies of
// `bytes` bytes back to back into its own ring of 4 slots (no reuse protocol: timing only) and waits for all of them once.
// out[0] = cycles lane 0 of warp 0 spent issuing, out[1] = cycles until everything had landed.
namespace {
__global__ void __launch_bounds__(512, 1)
k_bulk_rate(const float* __restrict__ src, long long* __restrict__ out, int n, int bytes, int warps, int planes, long long plane_floats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) tc::mbar_init(&bars[i], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    if (warp < warps && lane == 0) {
        uint8_t* ring = smem_raw + warp * 4 * bytes;
        // planes > 1: copy i reads plane i % planes (planes are plane_floats apart), like the T planes of the last backward stage
        const float* p = src + ((size_t)blockIdx.x * warps + warp) * (size_t)(n / planes + 1) * (bytes / 4);
        const long long t0 = clock64();
        for (int i = 0; i < n; ++i)
            bulk_g2s(tc::smem_u32(ring + (i & 3) * bytes), p + (size_t)(i % planes) * plane_floats + (size_t)(i / planes) * (bytes / 4), bytes, &bars[warp]);
        const long long t1 = clock64();
        mbar_arrive_expect_tx(&bars[warp], (uint32_t)(n * bytes));
        tc::mbar_wait(&bars[warp], 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
}
}  // namespace

extern "C" int s2v_debug_bulk_rate(const float* src, long long* out_dev, int32_t n, int32_t bytes, int32_t warps, int32_t planes,
                                   int64_t plane_floats, void* stream) {
    if (planes < 1) return S2V_ERR_BAD_ARG;
    if (warps < 1 || warps > 16 || bytes % 16 || bytes * 4 * warps > 200 * 1024 || n * bytes >= (1 << 20)) return S2V_ERR_BAD_ARG;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = bytes * 4 * warps + 1024;
    S2V_RETURN_IF(cudaFuncSetAttribute(k_bulk_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_bulk_rate<<<sms, 512, smem, (cudaStream_t)stream>>>(src, out_dev, n, bytes, warps, planes, plane_floats);
    return (int)cudaGetLastError();
}
