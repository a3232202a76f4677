// tcgen05 / TMEM building blocks for the p = 64 node transforms (sm_100a only).
//
// Every p x p contraction of the path is run as a 3xTF32 split-precision UMMA:
//   x = hi + lo,  hi = tf32_rn(x),  lo = tf32_rn(x - hi)
//   A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        (fp32 accumulate in TMEM)
// which keeps ~21 mantissa bits per product (the north star asks for 1e-5 of tensor
// scale; plain TF32 would give 5e-4).
//
// Shared-memory operand layout (both operands, both majors): a tile of R rows x 64 fp32
// columns is stored as two "column blocks" of 32 columns; inside a block a row is 128
// bytes, rows are packed at 128 B, and the eight 16-byte chunks of a row are XOR-swizzled
// with (row & 7) -- the canonical UMMA SWIZZLE_128B atom (8 rows x 128 B = 1024 B).
//   element (r, c):  block(c/32) * R*128  +  r*128  +  (((c%32)/4) ^ (r&7))*16  +  (c%4)*4
// This is a K-major operand (MN = rows, K = columns; SBO = 1024 B, k-step of 8 columns =
// +32 B inside a block).  Tiles must start at a 1024-byte aligned shared address.
//
// Operands that are contracted over their ROWS (the theta-gradient reductions A^T B) are
// MN-major; for 32-bit data the only legal MN-major layout is SWIZZLE_128B_BASE32B
// (cute Layout_MN_SW128_32B_Atom): same 128-byte rows and 32-column blocks, but the swizzle
// works on 32-byte granules with a 4-row period -- byte-address bits [5,7) ^= bits [7,9)
// (mapping probed on the hardware with one-hot operands, scripts/tc_debug.py):
//   element (r, c):  block(c/32) * R*128  +  r*128  +  (((c%32)/4) ^ ((r&3)<<1))*16  +  (c%4)*4
// K atom = 4 rows (SBO = 512 B), LBO = R*128 B between column blocks, k-step of 8 rows = +1024 B.
// A tile cannot be both; the backward round keeps its tiles MN-major in shared memory and feeds
// the row transform from TMEM (A operand in tensor memory, written by tcgen05.st).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- split precision -------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
    hi.x = tf32_rn(v.x); hi.y = tf32_rn(v.y); hi.z = tf32_rn(v.z); hi.w = tf32_rn(v.w);
    lo.x = tf32_rn(v.x - hi.x); lo.y = tf32_rn(v.y - hi.y); lo.z = tf32_rn(v.z - hi.z); lo.w = tf32_rn(v.w - hi.w);
}

// byte offset of the 16-byte chunk holding columns 4*c4..4*c4+3 of row r in a tile of R rows
template <int R>
__device__ __forceinline__ uint32_t tile_off(int r, int c4) {
    return (uint32_t)((c4 >> 3) * (R * 128) + r * 128 + (((c4 & 7) ^ (r & 7)) << 4));
}

// MN-major (SWIZZLE_128B_BASE32B) tile: byte offset of the 16-byte chunk c4 of row r
template <int R>
__device__ __forceinline__ uint32_t tile_off_mn(int r, int c4) {
    return (uint32_t)((c4 >> 3) * (R * 128) + r * 128 + (((c4 & 7) ^ ((r & 3) << 1)) << 4));
}

// ---- descriptors -----------------------------------------------------------------------
// SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand: rows = MN, k-step ks covers columns 8*ks..8*ks+7
template <int R>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, int ks) {
    return smem_desc(tile_saddr + (ks >> 2) * (R * 128) + (ks & 3) * 32, 16, 1024);
}
// MN-major operand (SWIZZLE_128B_BASE32B, layout type 1): MN = columns (blocks of 32 at
// LBO = R*128), K atoms of 4 rows at SBO = 512 B, k-step ks covers rows 8*ks..8*ks+7
template <int R>
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, int ks) {
    uint64_t d = smem_desc(tile_saddr + ks * 1024, R * 128, 512);
    return (d & ~(7ull << 61)) | (1ull << 61);
}
// 64-column bf16 tiles (128-byte rows, SWIZZLE_128B; see the bf16x3 notes below).  K-major view: rows = MN, k-step ks
// covers columns 16*ks..16*ks+15.  MN-major view: columns = MN (64 per tile, further tiles at lbo_bytes), k-step ks
// covers rows 16*ks..16*ks+15.
__device__ __forceinline__ uint64_t desc_k16(uint32_t tile_saddr, int ks) { return smem_desc(tile_saddr + ks * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn16(uint32_t tile_saddr, int ks, uint32_t lbo_bytes) {
    return smem_desc(tile_saddr + ks * 2048, lbo_bytes, 1024);
}
// Instruction descriptor, kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- bf16x3 split precision (backward kernels) ---------------------------------------------
// x = x1 + x2 + x3 with x_i bf16 (8 mantissa bits each, 24 in total); A*B ~= sum of the six products with
// i + j <= 4 (the dropped ones are below 2^-24 relative).  For 16-bit operands the K-major and the MN-major
// SWIZZLE_128B atoms are the SAME bytes: a row of 64 bf16 is 128 B, its eight 16-byte chunks XOR (row & 7):
//   element (r, c):  r*128 + (((c/8) ^ (r&7)) * 16) + (c%8)*2
// read as K-major (MN = rows, K = columns, k-step of 16 columns = +32 B, SBO = 1024 B) by a row transform and
// as MN-major (MN = columns, K = rows, K atom = 8 rows at SBO = 1024 B, k-step of 16 rows = +2048 B, LBO = the
// stride between stacked 64-column tiles) by a reduction over the rows -- one tile serves both contractions,
// which 32-bit operands cannot do (their MN-major layout swizzles 32-byte granules with a 4-row period).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
// four consecutive columns -> three packed groups (8 bytes each)
__device__ __forceinline__ void split3(const float4& v, uint2& p1, uint2& p2, uint2& p3) {
    p1.x = pack_bf16x2(v.x, v.y); p1.y = pack_bf16x2(v.z, v.w);
    float rx = v.x - bf16_lo_f32(p1.x), ry = v.y - bf16_hi_f32(p1.x), rz = v.z - bf16_lo_f32(p1.y), rw = v.w - bf16_hi_f32(p1.y);
    p2.x = pack_bf16x2(rx, ry); p2.y = pack_bf16x2(rz, rw);
    rx -= bf16_lo_f32(p2.x); ry -= bf16_hi_f32(p2.x); rz -= bf16_lo_f32(p2.y); rw -= bf16_hi_f32(p2.y);
    p3.x = pack_bf16x2(rx, ry); p3.y = pack_bf16x2(rz, rw);
}
// byte offset of the 8-byte group holding columns 4*c4..4*c4+3 of row r (64-column bf16 tile, 128-byte rows)
__device__ __forceinline__ uint32_t tile_off16(int r, int c4) {
    return (uint32_t)(r * 128 + ((((c4 >> 1) ^ (r & 7))) << 4) + ((c4 & 1) << 3));
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- mbarrier --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrival traps (sticky error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 28); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// Two barriers polled together: a try_wait costs ~100 cycles even when its phase has completed, and two waits in
// sequence pay that twice on the critical path of a role that has nothing else to do.
__device__ __forceinline__ void mbar_wait2(uint64_t* bar_a, uint32_t par_a, uint64_t* bar_b, uint32_t par_b) {
    bool a = false, b = false;
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        const bool ta = a || mbar_try_wait(bar_a, par_a), tb = b || mbar_try_wait(bar_b, par_b);
        a = ta; b = tb;
        if (a && b) return;
    }
    __trap();
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- tcgen05 ----------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; writes the TMEM base address to *slot (shared memory)
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}

// The issue helpers below are WARP-COLLECTIVE: all 32 lanes of a converged warp call them with identical operands
// and one elected lane issues.  Called under `if (lane == 0)` the operands live in per-thread registers and every
// tcgen05.mma costs a waterfall loop with five R2UR moves (65 cycles measured, 140-190 next to busy warps);
// called by the whole warp with operands derived from uniform values (kernel parameters, the shared-memory
// window, loop counters) the descriptors stay in uniform registers.
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// bf16 operands (kind::f16), fp32 accumulate
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// same with the A operand in tensor memory (lane = row, one column per tf32 element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs have completed (warp-collective like the MMAs; elect.sync
// picks the same lane every time, so the commit follows that lane's MMAs)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive columns: thread i of the warp gets lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    // the registers are in/out operands of the wait so that no use can be scheduled before it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr_a, uint32_t taddr_b, float (&va)[16], float (&vb)[16]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr_a));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr_b));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { va[i] = __uint_as_float(r[i]); vb[i] = __uint_as_float(r[16 + i]); }
}

// registers -> TMEM: thread i of the warp writes lane (lane_base + i), 32 consecutive columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
          "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
          "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
          "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
          "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- asynchronous copies (global -> shared without registers) ----------------------------------
// cp.async (LDGSTS): 16 bytes per lane; .cg bypasses L1.  Completion is reported to an mbarrier with
// cp.async.mbarrier.arrive.noinc (one pending-count unit per executing thread, triggered when all of the thread's
// earlier cp.async have landed).
__device__ __forceinline__ void cp_async16_if(uint32_t dst, const void* src, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(dst), "l"(src), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cp.async.bulk (TMA, one thread): `bytes` contiguous bytes (multiple of 16) global -> shared, completion as
// transaction bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

}  // namespace tc
