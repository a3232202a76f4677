// Warp-specialised bf16x3 backward pipeline (r01 / r02): the backward round and the last backward stage as ONE persistent
// 16-warp CTA per SM with producer / MMA-issue / epilogue roles ("The warp-specialised backward pipeline").
// The default of the tensor path until r02d (now path flag 3, S2V_B200_BWD_ROUND=ws, S2V_B200_FINAL_WS=1).
// Included by s2v_embed_tc.cu inside its anonymous namespace.
#pragma once

// ------------------------------------------------------------------------------------
// backward round, warp-specialised, bf16x3: one persistent CTA of 16 warps per SM.
//   du^{t-1} = (g theta2) * [mu^{t-1} > 0],  dtheta2 += g^T mu^{t-1},  g = gather of du^t over the transposed CSR.
// The gathered tile is needed by two contractions -- over its columns (row transform, K-major view) and over its
// rows (theta-gradient, MN-major view).  For 16-bit operands both views are the same SWIZZLE_128B bytes
// (s2v_tc.cuh), so the tile is split into three bf16 tiles once and read by both; fp32 accuracy comes from the six
// bf16 products with i + j <= 4.  A bf16 UMMA costs 65 cycles for any M <= 128, N <= 128 (scripts/tc_rate.py), so
// the products are stacked along M and N (tile orders [x3][x1][x2] for g, [x2][x1][x3] for mu and theta2 make
// every stack contiguous) and the stacked blocks are summed when the accumulators are read: 24 MMAs per tile.
//   warps 0-7   P : gather into REGISTERS (software pipeline over three tiles per lane: CSR extents of k+2,
//                   neighbour ids of k+1, rows of k), split3, store the g tiles of stage s
//   warps 8-11  F : lane 0 of warp 8 issues the MMAs of tile k:
//                       D1w[k&1][64 x 128]  = g1 t3 | g3 t1  +  g2 [t2|t1]  +  g1 [t2|t1]   (K-major, M = 64, N = 64/128/128)
//                       D2x[128 x 128]    (+)= [g1;g2]^T [m2|m1]                            (MN-major, all four blocks wanted)
//                       D2y[128 x 128]     += [g3;g1]^T [m1|m3]                             (blocks g3 m1 and g1 m3 wanted; 2^-16 of
//                                                                                            the result: lives in TMEM for the whole kernel)
//                   all four warps then move D1w of tile k-1 to the staging buffer (summing its two column halves)
//                   and, after every D2_GROUP tiles, add D2x into registers (the tensor core accumulates with
//                   truncation: chains are kept short, the long sum runs on the CUDA cores)
//   warps 12-15 E : stream the own mu^{t-1} rows (coalesced, two tiles ahead), split3 and store the m tiles, keep the
//                   sign bits; staging -> ReLU mask -> du^{t-1} with coalesced 128-bit stores
// Variants measured and dropped (AMS, 4.0 M rows; this kernel: 1.31 ms): producers that keep the rows of tile k+1
// in flight while storing tile k (setmaxnreg 152/128/80): 1.40 ms; queueing both MMA groups of a tile before the
// previous tile is drained: 1.44 ms; compile-time stage indices (descriptor = base + immediate): 1.47 ms.  With
// the MMAs, the shared-memory stores and the global stores switched off one at a time (-DS2V_WS_EXPERIMENTS,
// scripts/ws_experiments.py) the gather alone runs in 0.81 ms; MMAs add 0.34 ms, tile stores 0.27 ms.
// Three shared-memory stages of 48 KB, two D1w buffers.  mbarriers (s = k % 3, n = k / 3):
//   full_g[s]    8  P -> issuer                      full_m[s]   4  E -> issuer
//   mma_done[s]  1  tcgen05.commit -> P, E (stage s reusable: tile k-3), F (D1w[k&1] complete)
//   d1_free[k&1] 4  F -> issuer                      d2_free     4  F -> issuer (D2x read out after a group)
//   stg_full[k&1] 4  F -> E                          stg_empty[k&1] 4  E -> F
// ------------------------------------------------------------------------------------
constexpr int WS_P_WARPS = 8, WS_F_WARP0 = 8, WS_E_WARP0 = 12;
constexpr int WS_THREADS = 512;
#ifndef S2V_WS_STAGES
#define S2V_WS_STAGES 2      /* a third stage bought nothing: backward round 1.192 ms with two, 1.234 ms with three */
#endif
constexpr int WS_STAGES = S2V_WS_STAGES;
#ifndef S2V_WSF_STAGES
#define S2V_WSF_STAGES 1
#define S2V_WSF_SLOTS 7
#define S2V_WSF_LAG 2
#endif
constexpr int WSF_STAGES = S2V_WSF_STAGES;                            // last stage: ONE operand stage (its producers hold the next tile in registers
                                                         // while the MMAs read it) -- the other two are the TMA landing ring
constexpr int WSF_SLOTS = S2V_WSF_SLOTS;                             // landing slots of one tile plane (64 rows x 256 B = 16 KB): 112 KB in flight
constexpr int WSF_LAG = S2V_WSF_LAG;                               // a slot is re-armed WSF_LAG reads after it was read
constexpr int WS_D2_GROUP = 4;                           // tiles accumulated in TMEM between two read-outs of D2x
constexpr uint32_t WS_STAGE_BYTES = 2 * TILE16X3;        // [g3][g1][g2][m2][m1][m3]
constexpr int WS_STG_FLOATS = TCR * STG_LD;              // one staging buffer (two: tile parity)
constexpr uint32_t WS_G1 = TILE16, WS_G2 = 2 * TILE16, WS_G3 = 0;           // piece offsets inside the g tiles
constexpr uint32_t WS_X1 = TILE16, WS_X2 = 0, WS_X3 = 2 * TILE16;           // ... inside the m / theta tiles

struct WsShared {
    uint64_t full_g[WS_STAGES], full_m[WS_STAGES], mma_done[WS_STAGES];
    uint64_t d1_free[2], stg_full[2], stg_empty[2], d2_free;
    uint32_t tmem_base;
    uint32_t pad;
};

// The tensor core's fp32 accumulator rounds in ONE direction, so every element of D = g theta2 carries a bias of the same
// sign (~2.5e-7 of its magnitude); column sums over 10^4..10^6 rows with mixed signs amplify that to 1e-4 of the
// (cancelled) sum.  Odd tiles therefore run with BOTH operand tiles negated: the reductions (-g)^T (-m) are unchanged,
// the row transform (-g) theta2 is negated back in the epilogue, and its bias flips sign from tile to tile.
#ifndef S2V_WS_ALTSIGN
#define S2V_WS_ALTSIGN 0
#endif
__device__ __forceinline__ float4 ws_alt(const float4& v, int tile) {
    return (S2V_WS_ALTSIGN && (tile & 1)) ? make_float4(-v.x, -v.y, -v.z, -v.w) : v;
}
__device__ __forceinline__ void store_split3_at(uint8_t* base, uint32_t o1, uint32_t o2, uint32_t o3, int r, int c4, const float4& v) {
    uint2 p1, p2, p3;
    tc::split3(v, p1, p2, p3);
    const uint32_t off = tc::tile_off16(r, c4);
    *reinterpret_cast<uint2*>(base + o1 + off) = p1;
    *reinterpret_cast<uint2*>(base + o2 + off) = p2;
    *reinterpret_cast<uint2*>(base + o3 + off) = p3;
}

// D1w[64 x 128] = g1 t3 | g3 t1, += g2 [t2|t1], += g1 [t2|t1]   (g = stage tiles, w = theta2^T tiles [t2][t1][t3]).
// (Compile-time stage indices with base + immediate descriptors were tried: fewer instructions per MMA, but the
// six-way dispatch measured 12 % slower end to end.)
__device__ __forceinline__ void issue_ws_transform(uint32_t d1w, uint32_t g, uint32_t w) {
    constexpr uint32_t i64 = tc::idesc_bf16(64, 64, 0, 0), i128 = tc::idesc_bf16(64, 128, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tc::mma_bf16(d1w, tc::desc_k16(g + WS_G1, ks), tc::desc_k16(w + WS_X3, ks), i64, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tc::mma_bf16(d1w + 64, tc::desc_k16(g + WS_G3, ks), tc::desc_k16(w + WS_X1, ks), i64, ks > 0);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tc::mma_bf16(d1w, tc::desc_k16(g + WS_G2, ks), tc::desc_k16(w, ks), i128, true);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tc::mma_bf16(d1w, tc::desc_k16(g + WS_G1, ks), tc::desc_k16(w, ks), i128, true);
}
// D2y += [g3;g1]^T [m1|m3],  D2x (+)= [g1;g2]^T [m2|m1]      (m = the stage's m tiles [m2][m1][m3])
__device__ __forceinline__ void issue_ws_reduce(uint32_t d2x, uint32_t d2y, uint32_t g, uint32_t m, bool acc_x, bool acc_y) {
    constexpr uint32_t idesc = tc::idesc_bf16(128, 128, 1, 1);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        tc::mma_bf16(d2y, tc::desc_mn16(g + WS_G3, ks, TILE16), tc::desc_mn16(m + WS_X1, ks, TILE16), idesc, acc_y || ks > 0);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        tc::mma_bf16(d2x, tc::desc_mn16(g + WS_G1, ks, TILE16), tc::desc_mn16(m + WS_X2, ks, TILE16), idesc, acc_x || ks > 0);
}

// The same pipeline runs the LAST backward stage (FINAL = true; math of k_embed_final_bwd, s2v_embed.cu):
//   D = sum_t du^t ;  dtheta1b += D^T h ;  dh = (D theta1b) * [h > 0] ;  dtheta1a += dh^T x ;  column sums of D, c D,
//   (Npad - c) D, dh.
// D takes the place of g (P streams and sums the T planes instead of gathering), h = relu(theta1a x + b1a) the place
// of mu (E recomputes it from the x tile), theta1b the place of theta2; instead of storing dh, E reduces it against x.
#ifdef S2V_WSF_PROF
__device__ long long g_wsf_prof[16];
#define WSFP(i, a, b) do { if (FINAL && blockIdx.x == 0 && t == 0) wsfp[i] += (b) - (a); } while (0)
#define WSFT(v) const long long v = clock64()
#if S2V_WSF_PROF == 3
#define WSFP3(i, a, b) WSFP(i, a, b)
#define WSFT3(v) WSFT(v)
#else
#define WSFP3(i, a, b) do {} while (0)
#define WSFT3(v) do {} while (0)
#endif
#if S2V_WSF_PROF >= 2
#define WSFP2(i, a, b) do {} while (0)
#define WSFT2(v) do {} while (0)
#else
#define WSFP2(i, a, b) WSFP(i, a, b)
#define WSFT2(v) WSFT(v)
#endif
#else
#define WSFP(i, a, b) do {} while (0)
#define WSFT(v) do {} while (0)
#define WSFP2(i, a, b) do {} while (0)
#define WSFT2(v) do {} while (0)
#define WSFP3(i, a, b) do {} while (0)
#define WSFT3(v) do {} while (0)
#endif
struct WsArgs {
    const float* w;          // theta2 | theta1b  [64][64]
    const float* src;        // du^t (gathered) | du^1..du^{T-1} planes
    const float* own;        // mu^{t-1}        | du^T (last plane)
    float* dst;              // du^{t-1}        | unused
    const float* x;          // FINAL: node features [n][F]
    const float* th1a_w;     // FINAL: theta1a [64][F]
    const float* th1a_b;     // FINAL: [64]
    int F, T;                // FINAL: node_dim (<= WS_MAX_F), rounds
    float* partial;          // per-CTA partial sums (EmbedPartial<64> layout, s2v_embed.cu)
    int partial_stride, accumulate, num_tiles;
};
constexpr int WS_MAX_F = 8;
constexpr int WS_EP_TH1B = 64 * 64, WS_EP_COLD = 2 * 64 * 64, WS_EP_A1 = WS_EP_COLD + 64, WS_EP_A0 = WS_EP_A1 + 64,
              WS_EP_B1A = WS_EP_A0 + 64, WS_EP_TH1A = WS_EP_B1A + 64;

template <bool FINAL>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_embed_bwd_ws(s2v_graph_t g, WsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w_tiles = smem;                                 // W^T: tile(row n, col k) = W[k][n], pieces [t2][t1][t3]
    // The last stage runs on TWO operand stages: the third one's 48 KB are the landing ring of its TMA plane streams.
    constexpr int NST = FINAL ? WSF_STAGES : WS_STAGES;
    uint8_t* stage0 = w_tiles + TILE16X3;
    float* stg0 = reinterpret_cast<float*>(stage0 + NST * WS_STAGE_BYTES);          // 2 staging buffers
    float* xs0 = stg0 + 2 * WS_STG_FLOATS;                   // FINAL: NST x tiles [64][F]
    WsShared* sh = reinterpret_cast<WsShared*>(xs0 + 3 * TCR * WS_MAX_F);
    // FINAL: per-producer-warp landing rings (WSF_SLOTS slots of 8 rows x 256 B) + their mbarriers; the first 4 KB double as
    // the producers' reduction scratch [16][64] once every copy has landed and been read
    uint64_t* lbar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sh) + 256);     // full[WSF_SLOTS], empty[WSF_SLOTS]
    float* land = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(lbar) + 128);
    float* pscr = FINAL ? land : reinterpret_cast<float*>(lbar);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const float* __restrict__ theta2_w = a.w;
    const float* __restrict__ du_t = a.src;
    const float* __restrict__ mu_prev = a.own;
    float* __restrict__ du_prev = a.dst;
    float* __restrict__ partial = a.partial;
    const int partial_stride = a.partial_stride, num_tiles = a.num_tiles;
    int accumulate = a.accumulate;
#ifdef S2V_WS_EXPERIMENTS
    const int xf = accumulate >> 8;             // timing experiments: bit i switches one piece of work off (results are garbage)
    accumulate &= 1;
#define WS_X(bit) (xf & (1 << (bit)))
#else
#define WS_X(bit) 0
#endif

    if (warp == WS_F_WARP0) tc::tmem_alloc<512>(&sh->tmem_base);
    if (t == 0) {
        for (int i = 0; i < WS_STAGES; ++i) {
            tc::mbar_init(&sh->full_g[i], WS_P_WARPS);
            tc::mbar_init(&sh->full_m[i], 4);
            tc::mbar_init(&sh->mma_done[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&sh->d1_free[i], 4);
            tc::mbar_init(&sh->stg_full[i], 4);
            tc::mbar_init(&sh->stg_empty[i], 4);
        }
        tc::mbar_init(&sh->d2_free, 4);
        if (FINAL)
            for (int i = 0; i < WSF_SLOTS; ++i) { tc::mbar_init(&lbar[i], 1); tc::mbar_init(&lbar[WSF_SLOTS + i], WS_P_WARPS); }
        tc::fence_barrier_init();
    }
    for (int e = t; e < 64 * 16; e += WS_THREADS) {          // row n of the tile = column n of theta2
        const int n = e >> 4, c4 = e & 15;
        float4 v = make_float4(__ldg(theta2_w + (4 * c4) * 64 + n), __ldg(theta2_w + (4 * c4 + 1) * 64 + n),
                               __ldg(theta2_w + (4 * c4 + 2) * 64 + n), __ldg(theta2_w + (4 * c4 + 3) * 64 + n));
        store_split3_at(w_tiles, WS_X1, WS_X2, WS_X3, n, c4, v);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    // The CTA owns all 512 TMEM columns, so the allocation starts at lane 0, column 0: the constant keeps the MMA
    // operands in uniform registers.
    if (sh->tmem_base != 0) __trap();
    constexpr uint32_t tmem = 0;
    // TMEM columns: D1w[b] at 128 b (b < 2), D2x at 256, D2y at 384
    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    auto tile_rows = [&](int k, int& row0, int& rows) {
        row0 = (blockIdx.x + k * (int)gridDim.x) * TCR;
        rows = min(TCR, g.num_nodes - row0);
    };

    if (warp < WS_P_WARPS) {
        // ================================ P: gather (FINAL: sum of the T planes) ================================
        const int c4 = lane & 15, rl0 = 4 * (2 * warp + (lane >> 4));
        GatherPipe4 pipe;
        pipe.ext1.b = pipe.ext1.d = pipe.ext1.off = 0;
        pipe.cur.b = pipe.cur.d = pipe.cur.off = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) pipe.cur.idx[i] = 0;
        if (!FINAL && my_tiles > 0) {
            int row0, rows;
            tile_rows(0, row0, rows);
            gatherN_prefetch<4>(pipe.cur, g, g.rowptr_t, g.col_t, row0, rows, rl0);
            if (my_tiles > 1) { tile_rows(1, row0, rows); gatherN_extents<4>(pipe.ext1, g, g.rowptr_t, row0, rows, rl0); }
        }
        float4 pD = f4zero(), pA1 = f4zero(), pA0 = f4zero();  // FINAL: column sums of D, c D, (Npad - c) D over this lane's rows
#ifdef S2V_WSF_PROF
        long long wsfp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long wsf_t0 = clock64();
#endif
        // FINAL: the T planes of a tile arrive by TMA.  Request r = (tile r / T, plane r % T): ONE cp.async.bulk of the
        // tile's 64 rows (16 KB, contiguous) into slot r % WSF_SLOTS of the landing ring, completion as transaction bytes
        // on full[slot].  All eight producer warps read their rows of the slot and arrive on empty[slot]; request
        // r + WSF_SLOTS re-arms the slot.  A cp.async.bulk costs its issuing thread 200 - 400 cycles (scripts/bulk_rate.py),
        // so the issues rotate over the warps -- warp r % 8 issues request r, WSF_LAG reads after the slot's previous
        // content was read, when every warp is past it -- instead of stalling one warp five times per tile.
        uint64_t* lfull = lbar;
        uint64_t* lempty = lbar + WSF_SLOTS;
        const int fT = FINAL ? a.T : 1;
        const int ftotal = FINAL ? my_tiles * fT : 0;
        int freq = 0, fslot = 0, fphase = 0;                   // next request to read: slot freq % WSF_SLOTS, parity (freq / WSF_SLOTS) & 1
        auto final_request = [&](int r) {                      // one lane: issue request r
            const int kk = r / fT, j = r - kk * fT, slot = r % WSF_SLOTS;
            int row0, rows;
            tile_rows(kk, row0, rows);
            const float* src_plane = j < fT - 1 ? a.src + (size_t)j * ((size_t)g.num_nodes * 64) : a.own;
            tc::mbar_arrive_expect_tx(&lfull[slot], rows * 256);
            tc::bulk_g2s(tc::smem_u32(land + slot * (TCR * 64)), src_plane + (size_t)row0 * 64, rows * 256, &lfull[slot]);
        };
        if (FINAL && lane == 0)
            for (int r = warp; r < WSF_SLOTS && r < ftotal; r += WS_P_WARPS) final_request(r);
        // FINAL: in-degree c and padded size Npad of this lane's rows, requested one tile ahead.  The (graph, local row) of a
        // row advance from tile to tile by a fixed stride, so the only division is done once, here (four divisions, two
        // dependent index loads and the branches around them cost the producers ~1.9 k cycles per tile inside the loop).
        int cdi[4] = {0, 0, 0, 0}, npi[4] = {0, 0, 0, 0};
        int nli[4] = {0, 0, 0, 0}, ngi[4] = {0, 0, 0, 0};       // local row / graph of row i of the next tile to request (period > 0)
        int step_q = 0, step_r = 0;
        if (FINAL && g.period > 0) {
            const int stride = (int)gridDim.x * TCR;
            step_q = stride / g.period; step_r = stride - step_q * g.period;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = (int)blockIdx.x * TCR + rl0 + i;
                ngi[i] = r / g.period; nli[i] = r - ngi[i] * g.period;
            }
        }
        auto load_cd = [&](int kk) {                           // tiles are requested in order kk = 0, 1, 2, ...
            int row0, rows;
            tile_rows(kk, row0, rows);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cdi[i] = 0; npi[i] = g.npad_scalar;
                if (rl0 + i < rows) {
                    // volatile: the compiler may not re-issue these loads at their use a tile later
                    const int r = row0 + rl0 + i;
                    if (g.period > 0) {
                        asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(cdi[i]) : "l"(g.indeg + nli[i]));
                        if (g.npad) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(npi[i]) : "l"(g.npad + ngi[i]));
                    } else {
                        asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(cdi[i]) : "l"(g.indeg + r));
                        if (g.npad) {
                            int gi;
                            asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(gi) : "l"(g.gid + r));
                            asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(npi[i]) : "l"(g.npad + gi));
                        }
                    }
                }
                nli[i] += step_r; ngi[i] += step_q;
                if (nli[i] >= g.period && g.period > 0) { nli[i] -= g.period; ++ngi[i]; }
            }
        };
        if (FINAL && my_tiles > 0) load_cd(0);
        int s = 0, par = 0;                                   // s = k % NST, par = (k / NST) & 1
        for (int k = 0; k < my_tiles; ++k) {
            uint8_t* gt = stage0 + s * WS_STAGE_BYTES;
            float4 acc[4];
            TLW(t == 0, k, 0);
            TLS(t == 0, k, 11);
            if constexpr (!FINAL) {
                gather4_pipe_step(pipe, g, g.rowptr_t, g.col_t, du_t, rl0,
                                  [&](int& r2, int& n2) { r2 = 0; n2 = 0; if (k + 2 < my_tiles) tile_rows(k + 2, r2, n2); },
                                  [&]() {}, acc);
            } else {
                // D = sum_t du^t, planes in the order t = 1..T, read from the landing ring
                WSFT3(q_a);
                int row0, rows;
                tile_rows(k, row0, rows);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = f4zero();
                WSFT3(q_b); WSFP3(1, q_a, q_b);
                for (int j = 0; j < fT; ++j, ++freq) {
                    WSFT(p_a);
                    tc::mbar_wait(&lfull[fslot], fphase);
                    WSFT(p_b); WSFP(0, p_a, p_b);
                    const float* lp = land + fslot * (TCR * 64) + rl0 * 64 + 4 * c4;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (rl0 + i < rows) acc[i] = f4add(acc[i], ld4(lp + i * 64));
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&lempty[fslot]);
                    WSFT2(p_c); WSFP2(1, p_b, p_c);
                    // request rq re-arms the slot that was read WSF_LAG reads ago; its issuer is warp rq % 8
                    const int rq = freq - WSF_LAG + WSF_SLOTS;
                    if (rq >= WSF_SLOTS && rq < ftotal && (rq % WS_P_WARPS) == warp && lane == 0) {
                        tc::mbar_wait(&lempty[rq % WSF_SLOTS], ((rq / WSF_SLOTS) - 1) & 1);     // all eight warps have read request rq - WSF_SLOTS
                        final_request(rq);
                    }
                    WSFT2(p_d); WSFP2(2, p_c, p_d);
                    if (++fslot == WSF_SLOTS) { fslot = 0; fphase ^= 1; }
                }
                WSFT3(q_c); WSFP3(2, q_b, q_c);
                float cf[4], df[4];                            // in-degree c and Npad - c of the rows (raw values requested a tile ago)
#pragma unroll
                for (int i = 0; i < 4; ++i) { cf[i] = (float)cdi[i]; df[i] = (float)npi[i] - cf[i]; }
                if (k + 1 < my_tiles) load_cd(k + 1);           // dependent index loads: a tile period to land, nothing waits for them
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    pD = f4add(pD, acc[i]);
                    pA1.x = fmaf(cf[i], acc[i].x, pA1.x); pA1.y = fmaf(cf[i], acc[i].y, pA1.y);
                    pA1.z = fmaf(cf[i], acc[i].z, pA1.z); pA1.w = fmaf(cf[i], acc[i].w, pA1.w);
                    pA0.x = fmaf(df[i], acc[i].x, pA0.x); pA0.y = fmaf(df[i], acc[i].y, pA0.y);
                    pA0.z = fmaf(df[i], acc[i].z, pA0.z); pA0.w = fmaf(df[i], acc[i].w, pA0.w);
                }
                WSFT3(q_d); WSFP3(3, q_c, q_d);
            }
            TLW(t == 0, k, 1);
            WSFT2(p_e);
            tc::mbar_wait(&sh->mma_done[s], par ^ 1);            // the MMAs of tile k-3 have finished reading stage s
            WSFT2(p_f); WSFP2(3, p_e, p_f);
            TLW(t == 0, k, 2);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (!WS_X(3)) store_split3_at(gt, WS_G1, WS_G2, WS_G3, rl0 + i, c4, ws_alt(acc[i], k));
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->full_g[s]);
            WSFT2(p_g); WSFP2(4, p_f, p_g);
            TLW(t == 0, k, 3);
            if (++s == NST) { s = 0; par ^= 1; }
        }
#ifdef S2V_WSF_PROF
        if (FINAL && blockIdx.x == 0 && t == 0) {
            for (int i = 0; i < 5; ++i) g_wsf_prof[i] = wsfp[i];
            g_wsf_prof[5] = clock64() - wsf_t0; g_wsf_prof[6] = my_tiles;
        }
#endif
        if constexpr (FINAL) {                                 // column sums over the 16 half-warps, fixed order
            tc::named_bar_sync(3, 32 * WS_P_WARPS);            // pscr aliases the landing ring: every warp has read its last slot
            float* my = partial + (size_t)blockIdx.x * partial_stride;
            const int hw = t >> 4;                             // 0..15
            const float4 vals[3] = {pD, pA1, pA0};
            const int offs[3] = {WS_EP_COLD, WS_EP_A1, WS_EP_A0};
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3) {
                st4(pscr + hw * 64 + 4 * c4, vals[q3]);
                tc::named_bar_sync(3, 32 * WS_P_WARPS);
                if (t < 64) {
                    float sum = 0.f;
                    for (int h2 = 0; h2 < 16; ++h2) sum += pscr[h2 * 64 + t];
                    my[offs[q3] + t] = sum;
                }
                tc::named_bar_sync(3, 32 * WS_P_WARPS);
            }
        }
    } else if (warp < WS_E_WARP0) {
        // ================================ F: MMA issue, D1 -> staging, dtheta2 sums ================================
        const int q = warp & 3;                               // TMEM lane quarter this warp can access
        const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
        const bool issuer = warp == WS_F_WARP0;
        const uint32_t s_w = tc::smem_u32(w_tiles);
        const uint32_t d2x = tmem + 256, d2y = tmem + 384;
        float racc[64];                                       // sum over the column halves of D2x row 32q + lane
#pragma unroll
        for (int i = 0; i < 64; ++i) racc[i] = 0.f;

        auto drain = [&](int j) {                              // tile j: D1w -> staging; group end: D2x -> racc
            const int sj = j % NST, pj = (j / NST) & 1, bj = j & 1;
            float* stg = stg0 + bj * WS_STG_FLOATS;
            tc::mbar_wait2(&sh->mma_done[sj], pj, &sh->stg_empty[bj], ((j >> 1) & 1) ^ 1);
            TLW(t == 32 * WS_F_WARP0, j, 13);
            tc::tc_fence_after_sync();
            TLW(t == 32 * WS_F_WARP0, j, 6);
#pragma unroll
            for (int c = 0; c < 4; ++c) {                      // M = 64 layout: lanes < 16 of quarter q own row 16q + lane
                float va[16], vb[16];
                tc::tmem_ld16x2(tmem + 128 * bj + lane_sel + 16 * c, tmem + 128 * bj + lane_sel + 64 + 16 * c, va, vb);
                if (lane < 16 && !WS_X(5)) {
                    float* dst = stg + (16 * q + lane) * STG_LD + 16 * c;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        st4(dst + i, make_float4(va[i] + vb[i], va[i + 1] + vb[i + 1], va[i + 2] + vb[i + 2], va[i + 3] + vb[i + 3]));
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&sh->d1_free[bj]); tc::mbar_arrive(&sh->stg_full[bj]); }
            if (j % WS_D2_GROUP == WS_D2_GROUP - 1 || j == my_tiles - 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float va[16], vb[16];
                    tc::tmem_ld16x2(d2x + lane_sel + 16 * c, d2x + lane_sel + 64 + 16 * c, va, vb);
#pragma unroll
                    for (int i = 0; i < 16; ++i) racc[16 * c + i] += va[i] + vb[i];
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh->d2_free);
            }
            TLW(t == 32 * WS_F_WARP0, j, 7);
        };

        int s = 0, par = 0;
        for (int k = 0; k < my_tiles; ++k) {
            const uint32_t s_g = tc::smem_u32(stage0 + s * WS_STAGE_BYTES), s_m = s_g + TILE16X3;
            if (issuer) {                                      // whole warp (uniform operands), one elected lane issues;
                tc::mbar_wait2(&sh->full_g[s], par, &sh->d1_free[k & 1], ((k >> 1) & 1) ^ 1);   // row transform first: it does not touch D2x
                tc::tc_fence_after_sync();
                TLW(lane == 0, k, 4);
                if (!WS_X(1)) issue_ws_transform(128 * (k & 1), s_g, s_w);
                TLW(lane == 0, k, 12);
            }
            if (k > 0) drain(k - 1);                           // reads D2x out at a group end -> d2_free
            if (issuer) {
                const int grp = k / WS_D2_GROUP;
                if (k % WS_D2_GROUP == 0) tc::mbar_wait2(&sh->full_m[s], par, &sh->d2_free, (grp & 1) ^ 1);
                else tc::mbar_wait(&sh->full_m[s], par);
                tc::tc_fence_after_sync();
                TLW(lane == 0, k, 14);
                if (!WS_X(0)) issue_ws_reduce(256, 384, s_g, s_m, k % WS_D2_GROUP != 0, k > 0);
                tc::mma_commit(&sh->mma_done[s]);
                TLW(lane == 0, k, 5);
            }
            if (++s == NST) { s = 0; par ^= 1; }
        }
        if (my_tiles > 0) {
            drain(my_tiles - 1);
            // D2y: rows 0-63 (g3) x columns 0-63 (m1) and rows 64-127 (g1) x columns 64-127 (m3)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float va[16];
                tc::tmem_ld16(d2y + lane_sel + (q >= 2 ? 64 : 0) + 16 * c, va);
#pragma unroll
                for (int i = 0; i < 16; ++i) racc[16 * c + i] += va[i];
            }
        }
        // dtheta2 partial of this CTA, row n = 32 (q & 1) + lane: warps q < 2 hold the g1 (+ g3) part, q >= 2 the g2
        // (+ g1 m3) part.  Every MMA has completed, so stage 0 serves as the exchange buffer [64][65].
        float* xch = reinterpret_cast<float*>(stage0) + ((q & 1) * 32 + lane) * 65;
        if (q >= 2) {
#pragma unroll
            for (int i = 0; i < 64; ++i) xch[i] = racc[i];
        }
        tc::tc_fence_before_sync();
        tc::named_bar_sync(2, 128);
        if (q < 2) {
            float* my = partial + (size_t)blockIdx.x * partial_stride + (FINAL ? WS_EP_TH1B : 0) + (32 * q + lane) * 64;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                float4 x4 = make_float4(racc[i] + xch[i], racc[i + 1] + xch[i + 1], racc[i + 2] + xch[i + 2], racc[i + 3] + xch[i + 3]);
                if (accumulate) x4 = f4add(ld4(my + i), x4);
                st4(my + i, x4);
            }
        }
    } else {
        // ================================ E: own rows (FINAL: h tile), epilogue ================================
        const int et = t - 32 * WS_E_WARP0;                   // 0..127: chunk e = et + 128 i -> row e >> 4, columns 4 (e & 15)
        uint32_t sgn_a = 0u, sgn_b = 0u, sgn_c = 0u;          // [mu > 0] bits of this thread's 8 chunks: tiles k-1, k, k+1
        float4 m[8];
        // FINAL: theta1a columns 4cc..4cc+3 of this thread, its partial sums of dh and dh x^T, the x values it stages
        const int F = FINAL ? a.F : 0;
        float4 w1a[WS_MAX_F], b1a4 = f4zero(), pB1a = f4zero(), pth1a[WS_MAX_F];
        float xr[WS_MAX_F];                                   // FINAL: thread et < 64 stages row et of the x tile
#pragma unroll
        for (int f = 0; f < WS_MAX_F; ++f) xr[f] = 0.f;
        if constexpr (FINAL) {
            const int cc = et & 15;
            for (int e = et; e < 3 * TCR * WS_MAX_F; e += 128) xs0[e] = 0.f;   // rows are padded to WS_MAX_F floats; the pads stay zero
            tc::named_bar_sync(1, 128);
#pragma unroll
            for (int f = 0; f < WS_MAX_F; ++f) {
                pth1a[f] = f4zero();
                w1a[f] = f < F ? make_float4(__ldg(a.th1a_w + (4 * cc) * F + f), __ldg(a.th1a_w + (4 * cc + 1) * F + f),
                                             __ldg(a.th1a_w + (4 * cc + 2) * F + f), __ldg(a.th1a_w + (4 * cc + 3) * F + f))
                               : f4zero();
            }
            b1a4 = ldg4(a.th1a_b + 4 * cc);
        }
        auto load_m = [&](int k) {
            int row0, rows;
            tile_rows(k, row0, rows);
            if constexpr (FINAL) {
#pragma unroll
                for (int f = 0; f < WS_MAX_F; ++f)
                    xr[f] = (et < rows && f < F) ? __ldg(a.x + (size_t)(row0 + et) * F + f) : 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int e = et + 128 * i, rl = e >> 4, cc = e & 15;
                    m[i] = rl < rows ? ldg4(mu_prev + (size_t)(row0 + rl) * 64 + 4 * cc) : f4zero();
                }
            }
        };
        auto store_m = [&](int k) -> uint32_t {                // registers -> m tiles of stage k % 3; returns the sign bits
            const int sk = k % NST, pk = (k / NST) & 1;
            uint8_t* mt = stage0 + sk * WS_STAGE_BYTES + TILE16X3;
            const float* xs = xs0 + (k % 3) * TCR * WS_MAX_F;    // three x tiles whatever the number of operand stages: tile k+1
            if constexpr (FINAL) {                             // is staged while the dh of tile k-1 is still reduced against its x
                float* xw = xs0 + (k % 3) * TCR * WS_MAX_F;
                tc::named_bar_sync(1, 128);                    // every E thread is done with the x tile of k-3
                if (et < TCR) {
                    st4(xw + et * WS_MAX_F, make_float4(xr[0], xr[1], xr[2], xr[3]));
                    st4(xw + et * WS_MAX_F + 4, make_float4(xr[4], xr[5], xr[6], xr[7]));
                }
                tc::named_bar_sync(1, 128);
            }
            tc::mbar_wait(&sh->mma_done[sk], pk ^ 1);            // stage free (tile k-3)
            uint32_t bits = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = et + 128 * i;
                float4 val;
                if constexpr (FINAL) {                         // h = relu(theta1a x + b1a), chunk by chunk
                    const int rl = e >> 4;
                    const float4 xa = ld4(xs + rl * WS_MAX_F), xb = ld4(xs + rl * WS_MAX_F + 4);
                    const float xv[WS_MAX_F] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                    float4 h = b1a4;
#pragma unroll
                    for (int f = 0; f < WS_MAX_F; ++f) {       // theta1a columns past F are zero
                        h.x = fmaf(xv[f], w1a[f].x, h.x); h.y = fmaf(xv[f], w1a[f].y, h.y);
                        h.z = fmaf(xv[f], w1a[f].z, h.z); h.w = fmaf(xv[f], w1a[f].w, h.w);
                    }
                    val = make_float4(relu(h.x), relu(h.y), relu(h.z), relu(h.w));
                } else {
                    val = m[i];
                }
                if (!WS_X(4)) store_split3_at(mt, WS_X1, WS_X2, WS_X3, e >> 4, e & 15, ws_alt(val, k));
                bits |= (val.x > 0.f ? 1u : 0u) << (4 * i) | (val.y > 0.f ? 1u : 0u) << (4 * i + 1) |
                        (val.z > 0.f ? 1u : 0u) << (4 * i + 2) | (val.w > 0.f ? 1u : 0u) << (4 * i + 3);
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->full_m[sk]);
            return bits;
        };
        auto store_du = [&](int j, uint32_t bits) {            // staging -> ReLU mask -> du^{t-1} rows of tile j (FINAL: reduce dh)
            const int bj = j & 1;
            int row0, rows;
            tile_rows(j, row0, rows);
            const float* stg = stg0 + bj * WS_STG_FLOATS;
            const float* xs = xs0 + (j % 3) * TCR * WS_MAX_F;
            tc::mbar_wait(&sh->stg_full[bj], (j >> 1) & 1);
            TLW(et == 0, j, 9);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int e = et + 128 * i, rl = e >> 4, cc = e & 15;
                if (rl < rows && !WS_X(2)) {
                    float4 v = ws_alt(ld4(stg + rl * STG_LD + 4 * cc), j);
                    v = make_float4((bits >> (4 * i)) & 1u ? v.x : 0.f, (bits >> (4 * i + 1)) & 1u ? v.y : 0.f,
                                    (bits >> (4 * i + 2)) & 1u ? v.z : 0.f, (bits >> (4 * i + 3)) & 1u ? v.w : 0.f);
                    if constexpr (FINAL) {
                        pB1a = f4add(pB1a, v);
                        const float4 xa = ld4(xs + rl * WS_MAX_F), xb = ld4(xs + rl * WS_MAX_F + 4);
                        const float xv[WS_MAX_F] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                        for (int f = 0; f < WS_MAX_F; ++f) {   // x columns past F are zero
                            pth1a[f].x = fmaf(v.x, xv[f], pth1a[f].x); pth1a[f].y = fmaf(v.y, xv[f], pth1a[f].y);
                            pth1a[f].z = fmaf(v.z, xv[f], pth1a[f].z); pth1a[f].w = fmaf(v.w, xv[f], pth1a[f].w);
                        }
                    } else {
                        st4(du_prev + (size_t)(row0 + rl) * 64 + 4 * cc, v);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->stg_empty[bj]);
            TLW(et == 0, j, 10);
        };

        if (my_tiles > 0) {
            load_m(0);
            sgn_b = store_m(0);
            if (my_tiles > 1) load_m(1);
        }
        for (int k = 0; k < my_tiles; ++k) {
            if (k + 1 < my_tiles) {
                sgn_c = store_m(k + 1);
                TLW(et == 0, k, 8);
                if (k + 2 < my_tiles) load_m(k + 2);            // in flight during the store below
            }
            if (k > 0) store_du(k - 1, sgn_a);
            sgn_a = sgn_b;
            sgn_b = sgn_c;
        }
        if (my_tiles > 0) store_du(my_tiles - 1, sgn_a);
        if constexpr (FINAL) {                                 // sum_i dh_i and sum_i dh_i x_i^T over the 8 row groups, fixed order
            float* my = partial + (size_t)blockIdx.x * partial_stride;
            float* scr = xs0;                                  // [8][64]
            const int cc = et & 15, rg = et >> 4;
            for (int f = -1; f < F; ++f) {
                float4 val = pB1a;
#pragma unroll
                for (int ff = 0; ff < WS_MAX_F; ++ff) if (ff == f) val = pth1a[ff];
                tc::named_bar_sync(1, 128);
                st4(scr + rg * 64 + 4 * cc, val);
                tc::named_bar_sync(1, 128);
                if (et < 64) {
                    float sum = 0.f;
                    for (int r8 = 0; r8 < 8; ++r8) sum += scr[r8 * 64 + et];
                    if (f < 0) my[WS_EP_B1A + et] = sum;
                    else my[WS_EP_TH1A + et * F + f] = sum;
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == WS_F_WARP0) tc::tmem_dealloc<512>(tmem);
}

