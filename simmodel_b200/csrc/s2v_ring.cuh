// Forward propagation round with an ASYNCHRONOUS neighbour gather (included by s2v_embed_tc.cu, inside its namespace).
//
// k_embed_round_tc16 keeps the gathered rows in registers between issue and use: 16 warps x 16 x 128-bit loads per
// lane is what four 128-thread CTAs can hold, and the round runs at bytes-in-flight / latency.  Here the neighbour
// rows land in a shared-memory ring by cp.async (LDGSTS, 16 B per lane, L1 bypassed) -- no register is tied up
// while a row is in flight, and the ring (~120-145 KB) holds 2-3 tiles' worth of rows.  Measured alone
// (s2v_probe.cu, profiles/r02_gather_probe.md) this gather moves 5.8 TB/s with 128 KB of ring and 7.2 TB/s with 196 KB.
// One persistent CTA per SM, 18 warps:
//   M  (warp 0)     metadata: rowptr of the next tile (cp.async), chunk descriptors (<= cap edges, first ring slot of
//                   every row), ring flow control, neighbour ids of the chunk (cp.async into an id ring)
//   L  (warps 1-4)  copy issue, edge-parallel: a lane owns an edge id, every LDGSTS moves two 256-byte neighbour rows;
//                   completion arrives on the chunk's mbarrier (cp.async.mbarrier.arrive.noinc)
//   C  (warps 5-12) wait for the chunk, sum the landed rows per destination row in CSR order (bit-identical to
//                   gather_row), release the chunk; at the end of a tile split3 -> bf16x3 operand stage -> a_full
//   X  (warp 13)    tcgen05.mma issue: D1[k & 1] = agg theta2^T (24 MMAs, issue_transform16), commits -> a_free, d_full
//   E  (warps 14-17) epilogue: D1 -> staging -> + base, ReLU -> mu^t rows (coalesced 128-bit stores); the base rows of
//                   the next tile are requested before the accumulator is awaited
// Same arithmetic in the same order as k_embed_round_tc16: results are bit-identical.

#ifdef S2V_RING_PROF
__device__ long long g_ring_prof[32];
#define RGP_DECL(n) long long rgp_##n = 0
#define RGP_T(v) long long v = clock64()
#define RGP_ADD(n, a, b) rgp_##n += (b) - (a)
#define RGP_OUT(slot, n, cond) do { if (blockIdx.x == 0 && (cond)) g_ring_prof[slot] = rgp_##n; } while (0)
#else
#define RGP_DECL(n) do {} while (0)
#define RGP_T(v) do {} while (0)
#define RGP_ADD(n, a, b) do {} while (0)
#define RGP_OUT(slot, n, cond) do {} while (0)
#endif

constexpr int RG_NC = 8;                  // chunk descriptors in flight
constexpr int RG_L = 4, RG_C = 8;         // loader / consumer warps
constexpr int RG_X_WARP = 1 + RG_L + RG_C, RG_E_WARP0 = RG_X_WARP + 1;
constexpr int RG_THREADS = 32 * (RG_E_WARP0 + 4);
constexpr int RG_CAP_MAX = 256;           // edges per chunk (id ring)

struct RgDesc {
    int slot0, n, n_run0, off0, off1, last, final_chunk, tile, rows, pad0, pad1, pad2;
    int beg[TCR + 1];
    int pad3[3];
};
struct RgShared {
    uint64_t meta_full[RG_NC], full[RG_NC], empty[RG_NC];
    uint64_t a_full[2], a_free[2], d_full[2], d_free[2], base_full[2];
    uint32_t tmem_base, pad;
    RgDesc desc[RG_NC];
    int ids[RG_NC][RG_CAP_MAX];
    int rp[2][TCR + 4];                   // rowptr of the tile's rows: [0..64] + [65] = end of the first run
};

// ---- M: metadata + ring flow control (one warp) ----
__device__ __forceinline__ void ring_meta_role(RgShared* sh, const s2v_graph_t& g, const int32_t* __restrict__ rowptr,
                                               const int32_t* __restrict__ col, int my_tiles, int R, int cap) {
    const int lane = threadIdx.x & 31;
    auto tile_row0 = [&](int k) { return (blockIdx.x + k * (int)gridDim.x) * TCR; };
    // rowptr entries of tile k -> sh->rp[k & 1]: entry r (0..rows) of the run the row belongs to, [65] = end of run 0
    auto fetch_rp = [&](int k) {
        const int row0 = tile_row0(k), rows = min(TCR, g.num_nodes - row0);
        int* dstp = sh->rp[k & 1];
        int la = row0, first = rows;
        if (g.period > 0) { const int gi = row0 / g.period; la = row0 - gi * g.period; first = min(g.period - la, rows); }
        for (int r = lane; r <= rows; r += 32)
            tc::cp_async4(dstp + r, rowptr + ((r < first || g.period == 0) ? la + r : r - first));
        if (lane == 0) tc::cp_async4(dstp + 65, rowptr + la + first);
    };
    int head = 0, used = 0, tail = 0, cidx = 0;
    RGP_DECL(mwait); RGP_DECL(mall); RGP_T(m_t0);
    fetch_rp(0);
    tc::cp_async_commit();
    for (int k = 0; k < my_tiles; ++k) {
        if (k + 1 < my_tiles) fetch_rp(k + 1);
        tc::cp_async_commit();
        tc::cp_async_wait<1>();
        __syncwarp();
        const int row0 = tile_row0(k), rows = min(TCR, g.num_nodes - row0);
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
            RGP_T(m_a);
            while (used + n > R || cidx - tail >= RG_NC) {
                tc::mbar_wait(&sh->empty[tail % RG_NC], (tail / RG_NC) & 1);
                used -= sh->desc[tail % RG_NC].n;
                ++tail;
            }
            RGP_T(m_b); RGP_ADD(mwait, m_a, m_b);
            RgDesc* d = &sh->desc[cidx % RG_NC];
            d->beg[lane] = min(max(pb0 - lo, 0), n);
            d->beg[lane + 32] = min(max(pb1 - lo, 0), n);
            if (lane == 0) {
                d->beg[TCR] = n; d->slot0 = head; d->n = n; d->last = last; d->tile = k; d->rows = rows;
                d->n_run0 = min(max(n0 - lo, 0), n); d->off0 = off0; d->off1 = off0 + g.period;
                d->final_chunk = last && k == my_tiles - 1;
            }
            int* idp = sh->ids[cidx % RG_NC];
            for (int te = lo + lane; te < hi; te += 32)
                tc::cp_async4(idp + te - lo, col + (te < n0 ? e0 + te : te - n0));
            tc::cp_async_arrive_noinc(&sh->meta_full[cidx % RG_NC]);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->meta_full[cidx % RG_NC]);
            head += n; if (head >= R) head -= R;
            used += n;
            ++cidx;
        }
    }
    RGP_T(m_t1); RGP_ADD(mall, m_t0, m_t1);
    RGP_OUT(0, mwait, lane == 0); RGP_OUT(1, mall, lane == 0);
}

// ---- L: copy issue (RG_L warps; lw = index of this warp among them) ----
__device__ __forceinline__ void ring_load_role(RgShared* sh, const float* __restrict__ src, uint32_t ring_s, int R, int lw) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, c = lane & 15, half = lane >> 4;
    RGP_DECL(lwait); RGP_DECL(lissue);
    for (int cidx = 0;; ++cidx) {
        RGP_T(l_a);
        tc::mbar_wait(&sh->meta_full[cidx % RG_NC], (cidx / RG_NC) & 1);
        RGP_T(l_b); RGP_ADD(lwait, l_a, l_b);
        const RgDesc* d = &sh->desc[cidx % RG_NC];
        const int slot0 = d->slot0, n = d->n, nr0 = d->n_run0, off0 = d->off0, off1 = d->off1;
        const int fin = d->final_chunk;
        uint64_t* full = &sh->full[cidx % RG_NC];
        const int* idp = sh->ids[cidx % RG_NC];
        for (int base = lw * 32; base < n; base += RG_L * 32) {
            const int idv = base + lane < n ? idp[base + lane] : 0;
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int el = 2 * it + half, te = base + el;
                const int j = __shfl_sync(FULL, idv, el);
                int slot = slot0 + te; if (slot >= R) slot -= R;
                const int off = te < nr0 ? off0 : off1;
                tc::cp_async16_if(ring_s + slot * 256 + 16 * c, src + (size_t)(off + j) * 64 + 4 * c, te < n);
            }
        }
        tc::cp_async_arrive_noinc(full);
        if (lw == 0 && lane == 0) tc::mbar_arrive(full);
        RGP_T(l_c); RGP_ADD(lissue, l_b, l_c);
        if (fin) break;
    }
    RGP_OUT(2, lwait, lw == 0 && lane == 0); RGP_OUT(3, lissue, lw == 0 && lane == 0);
}

// ---- C: sum the rows of one landed chunk into acc (rows 4 hw .. 4 hw + 3 of the tile), CSR order ----
struct RgChunk { int last, fin, tile, rows; };
__device__ __forceinline__ RgChunk ring_consume_chunk(RgShared* sh, const uint8_t* ring, int R, int cidx, int hw, int c,
                                                      float4 (&acc)[4]) {
    const RgDesc* d = &sh->desc[cidx % RG_NC];
    const int slot0 = d->slot0;
    RgChunk r;
    r.last = d->last; r.fin = d->final_chunk; r.tile = d->tile; r.rows = d->rows;
    int sb[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) sb[i] = d->beg[4 * hw + i];
    auto slot_ptr = [&](int s) -> const float* {
        int slot = slot0 + s; if (slot >= R) slot -= R;
        return reinterpret_cast<const float*>(ring + slot * 256) + 4 * c;
    };
    float4 v[4][4];                           // first four neighbours of the four rows: 16 independent shared-memory loads
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[i][q] = ld4(slot_ptr(sb[i] + q < sb[i + 1] ? sb[i] + q : 0));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (sb[i] + q < sb[i + 1]) acc[i] = f4add(acc[i], v[i][q]);
        for (int s = sb[i] + 4; s < sb[i + 1]; ++s) acc[i] = f4add(acc[i], ld4(slot_ptr(s)));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(&sh->empty[cidx % RG_NC]);     // slots and descriptor are free again
    return r;
}

// D1 (UMMA M = 64) -> staging[row][STG_LD]; q = TMEM lane quarter of this warp (warp id % 4)
__device__ __forceinline__ void d1_to_staging_q(uint32_t d1_tmem, float* stg, int q, int lane) {
    float v[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tc::tmem_ld32(d1_tmem + ((uint32_t)(32 * q) << 16) + 32 * half, v);
        if (lane < 16) {
            float* dst = stg + (16 * q + lane) * STG_LD + 32 * half;
#pragma unroll
            for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        }
    }
}

template <int NS>
__global__ void __launch_bounds__(RG_THREADS, 1)
k_embed_round_ring(s2v_graph_t g, const float* __restrict__ theta2_w, const float* __restrict__ base,
                   const float* __restrict__ mu_prev, float* __restrict__ mu_next, int num_tiles, int R, int cap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                                   // theta2 [n][k], three bf16 tiles
    uint8_t* a0 = w1 + TILE16X3;                          // NS operand stages (gathered tile, three bf16 tiles each)
    float* stg = reinterpret_cast<float*>(a0 + NS * TILE16X3);
    float* bbuf = stg + TCR * STG_LD;                     // base rows of tiles k, k + 1: two 16 KB buffers filled by cp.async.bulk
    uint8_t* ring = reinterpret_cast<uint8_t*>(bbuf + 2 * TCR * 64);     // R slots of 256 B
    __shared__ RgShared sh_static;
    RgShared* sh = &sh_static;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (warp == RG_X_WARP) tc::tmem_alloc<512>(&sh->tmem_base);
    if (t == 0) {
        for (int i = 0; i < RG_NC; ++i) {
            tc::mbar_init(&sh->meta_full[i], 33);
            tc::mbar_init(&sh->full[i], RG_L * 32 + 1);
            tc::mbar_init(&sh->empty[i], RG_C);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&sh->a_full[i], RG_C);
            tc::mbar_init(&sh->a_free[i], 1);
            tc::mbar_init(&sh->d_full[i], 1);
            tc::mbar_init(&sh->d_free[i], 4);
            tc::mbar_init(&sh->base_full[i], 1);
        }
        tc::fence_barrier_init();
    }
    load_weight_tiles16(w1, theta2_w, false, RG_THREADS);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (sh->tmem_base != 0) __trap();                     // the CTA owns all 512 columns: D1[b] at column 64 b
    const int my_tiles = (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;      // grid <= num_tiles
    auto tile_row0 = [&](int k) { return (blockIdx.x + k * (int)gridDim.x) * TCR; };

    if (warp == 0) {
        ring_meta_role(sh, g, g.rowptr, g.col, my_tiles, R, cap);
    } else if (warp <= RG_L) {
        ring_load_role(sh, mu_prev, tc::smem_u32(ring), R, warp - 1);
    } else if (warp < RG_X_WARP) {
        // ================================ C ================================
        const int cw = warp - 1 - RG_L, c = lane & 15, hw = 2 * cw + (lane >> 4);
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = f4zero();
        RGP_DECL(cwait); RGP_DECL(csum); RGP_DECL(cfree); RGP_DECL(csplit); RGP_DECL(call); RGP_T(c_t0);
        for (int cidx = 0;; ++cidx) {
            RGP_T(c_a);
            tc::mbar_wait(&sh->full[cidx % RG_NC], (cidx / RG_NC) & 1);
            RGP_T(c_b); RGP_ADD(cwait, c_a, c_b);
            const RgChunk ch = ring_consume_chunk(sh, ring, R, cidx, hw, c, acc);
            RGP_T(c_c); RGP_ADD(csum, c_b, c_c);
            if (ch.last) {
                const int k = ch.tile, s = k % NS;
                tc::mbar_wait(&sh->a_free[s], ((k / NS) & 1) ^ 1);       // the MMAs of tile k - NS have read stage s
                RGP_T(c_d); RGP_ADD(cfree, c_c, c_d);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    store_split3(a0 + s * TILE16X3, 4 * hw + i, c, acc[i]);
                    acc[i] = f4zero();
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->a_full[s]);
                RGP_T(c_e); RGP_ADD(csplit, c_d, c_e);
            }
            if (ch.fin) break;
        }
        RGP_T(c_t1); RGP_ADD(call, c_t0, c_t1);
        RGP_OUT(4, cwait, cw == 0 && lane == 0); RGP_OUT(5, csum, cw == 0 && lane == 0); RGP_OUT(6, cfree, cw == 0 && lane == 0);
        RGP_OUT(7, csplit, cw == 0 && lane == 0); RGP_OUT(8, call, cw == 0 && lane == 0);
    } else if (warp == RG_X_WARP) {
        // ================================ X ================================
        const uint32_t s_w = tc::smem_u32(w1), s_a = tc::smem_u32(a0);
        RGP_DECL(xwait); RGP_DECL(xissue);
        for (int k = 0; k < my_tiles; ++k) {
            const int s = k % NS, b = k & 1;
            RGP_T(x_a);
            tc::mbar_wait2(&sh->a_full[s], (k / NS) & 1, &sh->d_free[b], ((k >> 1) & 1) ^ 1);
            RGP_T(x_b); RGP_ADD(xwait, x_a, x_b);
            tc::tc_fence_after_sync();
            issue_transform16(64 * b, s_a + s * TILE16X3, s_w);
            tc::mma_commit(&sh->a_free[s]);
            tc::mma_commit(&sh->d_full[b]);
            RGP_T(x_c); RGP_ADD(xissue, x_b, x_c);
        }
        RGP_OUT(9, xwait, lane == 0); RGP_OUT(10, xissue, lane == 0);
    } else {
        // ================================ E ================================
        const int et = t - 32 * RG_E_WARP0, q = warp & 3;
        // the tile's base rows are one contiguous 16 KB block: a single cp.async.bulk (TMA) per tile, two tiles ahead,
        // issued by one E thread -- no register is tied up while they are in flight
        auto request_base = [&](int k) {
            const int row0 = tile_row0(k), rows = min(TCR, g.num_nodes - row0);
            tc::fence_proxy_async_smem();                 // the buffer was read through the generic proxy (tile k - 2)
            tc::bulk_g2s(tc::smem_u32(bbuf + (k & 1) * TCR * 64), base + (size_t)row0 * 64, rows * 256, &sh->base_full[k & 1]);
            tc::mbar_arrive_expect_tx(&sh->base_full[k & 1], rows * 256);
        };
        if (et == 0) { request_base(0); if (my_tiles > 1) request_base(1); }
        RGP_DECL(ewait); RGP_DECL(ework); RGP_DECL(e1); RGP_DECL(e2); RGP_DECL(e3);
        for (int k = 0; k < my_tiles; ++k) {
            const int b = k & 1;
            const int row0 = tile_row0(k), rows = min(TCR, g.num_nodes - row0);
            const float* bb = bbuf + b * TCR * 64;
            RGP_T(e_a);
            tc::mbar_wait(&sh->d_full[b], (k >> 1) & 1);
            RGP_T(e_b); RGP_ADD(ewait, e_a, e_b);
            tc::tc_fence_after_sync();
            d1_to_staging_q(64 * b, stg, q, lane);
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->d_free[b]);
            RGP_T(e_p); RGP_ADD(e1, e_b, e_p);
            tc::mbar_wait(&sh->base_full[b], (k >> 1) & 1);
            tc::named_bar_sync(1, 128);
            RGP_T(e_q); RGP_ADD(e2, e_p, e_q);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = et + 128 * i, rl = e >> 4, c4 = e & 15;
                if (rl < rows) {
                    const float4 v = ld4(stg + rl * STG_LD + 4 * c4), bv = ld4(bb + rl * 64 + 4 * c4);
                    st4(mu_next + (size_t)(row0 + rl) * 64 + 4 * c4,
                        make_float4(relu(v.x + bv.x), relu(v.y + bv.y), relu(v.z + bv.z), relu(v.w + bv.w)));
                }
            }
            RGP_T(e_r); RGP_ADD(e3, e_q, e_r);
            tc::named_bar_sync(1, 128);                   // staging and the base buffer are free again
            if (et == 0 && k + 2 < my_tiles) request_base(k + 2);
            RGP_T(e_c); RGP_ADD(ework, e_b, e_c);
        }
        RGP_OUT(11, ewait, et == 0); RGP_OUT(12, ework, et == 0);
        RGP_OUT(14, e1, et == 0); RGP_OUT(15, e2, et == 0); RGP_OUT(16, e3, et == 0);
        if (blockIdx.x == 0 && et == 0) {
#ifdef S2V_RING_PROF
            g_ring_prof[13] = my_tiles;
#endif
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == RG_X_WARP) tc::tmem_dealloc<512>(0);
}
