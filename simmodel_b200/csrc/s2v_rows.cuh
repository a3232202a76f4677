// Node-wise linear layers of the GATv2 branch on the tensor cores (k_rows_gemm16 / k_rows_reduce16, bf16x3): the lin_l / lin_r
// GEMMs of every shipped GATv2 layer shape.  Included by s2v_embed_tc.cu inside its anonymous namespace.
#pragma once

// ------------------------------------------------------------------------------------
// Node-wise linear layers of the GATv2 branch on the tensor cores (bf16x3, fp32 accuracy ~3e-7 of scale):
//   k_rows_gemm16<KB,NB,MASK>:  out[n, 64 NB] = A[n, 64 KB] * Wt   (+ optional (mask > 0) on the output)
//       forward  x_l | x_r = x [Wl;Wr]^T :  KB = 1, NB = 2, B tile(nb)(n, k) = W[64 nb + n][k]
//       backward dx = [dx_l | dx_r] [Wl;Wr] * [x > 0] :  KB = 2, NB = 1, B tile(kb)(k, n) = W[64 kb + n][k]
//   k_rows_reduce16<AB>:  partial[cta][64 AB, 64] = sum over the CTA's tiles of A[:, 64 ab ..]^T B   (dW = dxlr^T x)
// Tile = 64 rows; same single-tile pipeline as k_head_tc: rows of the next tile are in flight (registers) during the
// MMAs and the epilogue, several CTAs per SM interleave their phases.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void issue_transform16_acc(uint32_t d_tmem, uint32_t a1, uint32_t b1, bool accumulate) {
    constexpr uint32_t idesc = tc::idesc_bf16(64, 64, 0, 0);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        int i, j;
        pair6(p, i, j);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_bf16(d_tmem, tc::desc_k16(a1 + i * TILE16, ks), tc::desc_k16(b1 + j * TILE16, ks), idesc,
                         accumulate || p > 0 || ks > 0);
    }
}

// Generalised operands (r02): A / mask / out are column windows of wider row-major matrices, the weight block is a
// window of W (either orientation), K may be ragged (the F = 7 input layer: rows read with scalar loads, zero padded to
// 64), the output may accumulate (a 256-long contraction runs as two passes of 128).
struct RowsGemmArgs {
    const float* A; int lda, a_col0, k_valid;          // A[:, a_col0 .. a_col0 + k_valid), k_valid <= 64 KB
    const float* W; int ldw, w_row0, w_col0, transpose_w, w_n_valid;
    //   transpose_w = 0: tile(nb, kb)(n, k) = W[w_row0 + 64 nb + n][w_col0 + 64 kb + k]      (out = A W^T)
    //   transpose_w = 1: tile(nb, kb)(n, k) = W[w_row0 + 64 kb + k][w_col0 + 64 nb + n]      (out = A W)
    const float* M; int ldm, m_col0;                   // optional mask window (same columns as out)
    float* out; int ldo, o_col0;                       // out[:, o_col0 .. o_col0 + 64 NB)
    int accumulate, n, num_tiles;
};

__device__ __forceinline__ void load_weight_tiles16g(uint8_t* t1, const RowsGemmArgs& a, int nb, int kb, int nthreads) {
    for (int e = threadIdx.x; e < 64 * 16; e += nthreads) {
        const int nrow = e >> 4, c4 = e & 15;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 64 * kb + 4 * c4 + j, nn = 64 * nb + nrow;
            const bool ok = k < a.k_valid && nn < a.w_n_valid;
            v[j] = !ok ? 0.f : (a.transpose_w ? __ldg(a.W + (size_t)(a.w_row0 + k) * a.ldw + a.w_col0 + nn)
                                              : __ldg(a.W + (size_t)(a.w_row0 + nn) * a.ldw + a.w_col0 + k));
        }
        store_split3(t1, nrow, c4, make_float4(v[0], v[1], v[2], v[3]));
    }
}

template <int KB, int NB, bool MASK>
__global__ void __launch_bounds__(TC_THREADS)
k_rows_gemm16(RowsGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w1 = smem;                                   // NB x KB weight blocks, three bf16 tiles each
    uint8_t* a1 = w1 + NB * KB * TILE16X3;                // KB row tiles; reused as the fp32 staging of the epilogue
    TcShared* sh = reinterpret_cast<TcShared*>(a1 + KB * TILE16X3);
    float* stg = reinterpret_cast<float*>(a1);
    const int t = threadIdx.x, warp = t >> 5;
    const int n = a.n, num_tiles = a.num_tiles;
    if (warp == 0) tc::tmem_alloc<64 * NB>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    for (int nb = 0; nb < NB; ++nb)
        for (int kb = 0; kb < KB; ++kb) load_weight_tiles16g(w1 + (nb * KB + kb) * TILE16X3, a, nb, kb, TC_THREADS);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_w = tc::smem_u32(w1);
    uint32_t phase = 0;
    // vector loads when the window is 16-byte aligned and complete, scalar + zero padding otherwise (F = 7)
    const bool a_vec = a.k_valid == 64 * KB && (a.lda & 3) == 0 && (a.a_col0 & 3) == 0 && ((uintptr_t)a.A & 15) == 0;
    float4 m[KB][8];
    auto load_rows = [&](int tile) {
        const int row0 = tile * TCR, rows = min(TCR, n - row0);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
                float4 v = f4zero();
                if (rl < rows) {
                    const float* src = a.A + (size_t)(row0 + rl) * a.lda + a.a_col0 + 64 * kb + 4 * c4;
                    if (a_vec) v = ldg4(src);
                    else {
                        const int k0 = 64 * kb + 4 * c4;
                        if (k0 < a.k_valid) v.x = __ldg(src);
                        if (k0 + 1 < a.k_valid) v.y = __ldg(src + 1);
                        if (k0 + 2 < a.k_valid) v.z = __ldg(src + 2);
                        if (k0 + 3 < a.k_valid) v.w = __ldg(src + 3);
                    }
                }
                m[kb][i] = v;
            }
    };
    if (blockIdx.x < num_tiles) load_rows(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = tile * TCR, rows = min(TCR, n - row0);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = t + TC_THREADS * i;
                store_split3(a1 + kb * TILE16X3, e >> 4, e & 15, m[kb][i]);
            }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
                    issue_transform16_acc(d + 64 * nb, s_a + kb * TILE16X3, s_w + (nb * KB + kb) * TILE16X3, kb > 0);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        float4 mk[MASK ? NB : 1][8];                       // the mask rows of this tile: in flight during the MMAs
        if (MASK) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
                    mk[nb][i] = rl < rows ? ldg4(a.M + (size_t)(row0 + rl) * a.ldm + a.m_col0 + 64 * nb + 4 * c4) : f4zero();
                }
        }
        if (tile + (int)gridDim.x < num_tiles) load_rows(tile + gridDim.x);      // in flight during the MMAs and the epilogue
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            d1_to_staging(d + 64 * nb, stg);               // the A tiles are dead once the MMAs have completed
            tc::tc_fence_before_sync();
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
                if (rl < rows) {
                    float4 v = ld4(stg + rl * STG_LD + 4 * c4);
                    float* dst = a.out + (size_t)(row0 + rl) * a.ldo + a.o_col0 + 64 * nb + 4 * c4;
                    if (a.accumulate) v = f4add(ld4(dst), v);
                    if (MASK) {
                        const float4 k4 = mk[nb][i];
                        v = make_float4(k4.x > 0.f ? v.x : 0.f, k4.y > 0.f ? v.y : 0.f, k4.z > 0.f ? v.z : 0.f, k4.w > 0.f ? v.w : 0.f);
                    }
                    st4(dst, v);
                }
            }
            __syncthreads();                               // staging (= A tiles) free again
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64 * NB>(d);
}

struct RowsReduceArgs {
    const float* A; int lda, a_col0;                   // A[:, a_col0 .. a_col0 + 64 AB)
    const float* B; int ldb, b_col0, b_valid;          // B[:, b_col0 .. b_col0 + b_valid), b_valid <= 64 (zero padded)
    float* partial; int n, num_tiles;
};

template <int AB>
__global__ void __launch_bounds__(TC_THREADS)
k_rows_reduce16(RowsReduceArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a1 = smem;                                   // AB column blocks of A, three bf16 tiles each
    uint8_t* b1 = a1 + AB * TILE16X3;
    TcShared* sh = reinterpret_cast<TcShared*>(b1 + TILE16X3);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n = a.n, num_tiles = a.num_tiles;
    if (warp == 0) tc::tmem_alloc<64 * AB>(&sh->tmem_base);
    if (t == 0) { tc::mbar_init(&sh->bar, 1); tc::fence_barrier_init(); }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t d = sh->tmem_base;
    const uint32_t s_a = tc::smem_u32(a1), s_b = tc::smem_u32(b1);
    uint32_t phase = 0;
    float racc[AB][64];                                   // row 16 warp + lane (lane < 16) of each 64 x 64 block
#pragma unroll
    for (int ab = 0; ab < AB; ++ab)
#pragma unroll
        for (int i = 0; i < 64; ++i) racc[ab][i] = 0.f;
    const bool b_vec = a.b_valid == 64 && (a.ldb & 3) == 0 && (a.b_col0 & 3) == 0 && ((uintptr_t)a.B & 15) == 0;
    float4 ma[AB][8], mb[8];
    auto load_rows = [&](int tile) {
        const int row0 = tile * TCR, rows = min(TCR, n - row0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = t + TC_THREADS * i, rl = e >> 4, c4 = e & 15;
#pragma unroll
            for (int ab = 0; ab < AB; ++ab)
                ma[ab][i] = rl < rows ? ldg4(a.A + (size_t)(row0 + rl) * a.lda + a.a_col0 + 64 * ab + 4 * c4) : f4zero();
            float4 v = f4zero();
            if (rl < rows) {
                const float* src = a.B + (size_t)(row0 + rl) * a.ldb + a.b_col0 + 4 * c4;
                if (b_vec) v = ldg4(src);
                else {
                    if (4 * c4 < a.b_valid) v.x = __ldg(src);
                    if (4 * c4 + 1 < a.b_valid) v.y = __ldg(src + 1);
                    if (4 * c4 + 2 < a.b_valid) v.z = __ldg(src + 2);
                    if (4 * c4 + 3 < a.b_valid) v.w = __ldg(src + 3);
                }
            }
            mb[i] = v;
        }
    };
    if (blockIdx.x < num_tiles) load_rows(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = t + TC_THREADS * i;
#pragma unroll
            for (int ab = 0; ab < AB; ++ab) store_split3(a1 + ab * TILE16X3, e >> 4, e & 15, ma[ab][i]);
            store_split3(b1, e >> 4, e & 15, mb[i]);
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        // the tensor core adds into its accumulator with truncation: at most four tiles are chained in TMEM, then the
        // group is summed on the CUDA cores (same rule as the s2v backward pipeline)
        const bool first = (it & 3) == 0, last = (it & 3) == 3 || tile + (int)gridDim.x >= num_tiles;
        if (warp == 0) {
            tc::tc_fence_after_sync();
#pragma unroll
            for (int ab = 0; ab < AB; ++ab) issue_reduce16(d + 64 * ab, s_a + ab * TILE16X3, s_b, !first);
            tc::mma_commit(&sh->bar);
        }
        __syncwarp();
        if (tile + (int)gridDim.x < num_tiles) load_rows(tile + gridDim.x);
        tc::mbar_wait(&sh->bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        if (last) {
#pragma unroll
            for (int ab = 0; ab < AB; ++ab)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[32];
                    tc::tmem_ld32(d + 64 * ab + ((uint32_t)(32 * warp) << 16) + 32 * half, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) racc[ab][32 * half + i] += v[i];
                }
        }
        tc::tc_fence_before_sync();
        __syncthreads();                                   // accumulators and tiles free again
    }
    if (lane < 16) {
        float* my = a.partial + (size_t)blockIdx.x * (AB * 64 * 64);
#pragma unroll
        for (int ab = 0; ab < AB; ++ab)
#pragma unroll
            for (int i = 0; i < 64; i += 4)
                st4(my + (size_t)(ab * 64 + 16 * warp + lane) * 64 + i, make_float4(racc[ab][i], racc[ab][i + 1], racc[ab][i + 2], racc[ab][i + 3]));
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64 * AB>(d);
}

