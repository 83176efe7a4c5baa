// tc_common.cuh - sm_100a primitives for the tensor-core path: mbarrier, bulk async copy (TMA unit,
// SASS UBLKCP), tcgen05 (alloc / mma / commit / ld / fences), UMMA descriptors, and the split-f16
// operand format.
//
// Arithmetic ("f16x3"): every fp32 value a is carried as a_hi = f16(a), a_lo = f16(a - a_hi)
// (22 significand bits).  A product a*b is accumulated in fp32 on the tensor cores as
//   a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (the dropped a_lo*b_lo term is ~2^-22 relative)
// i.e. three tcgen05.mma.kind::f16 passes per K step into one TMEM accumulator.
//
// Shared-memory operand image (used for BOTH the A and the B operand, K-major, no swizzle -
// UMMA "interleave" canonical layout  ((8,m),(8,k)) : ((16 B, SBO), (2 B, LBO))):
//   byte(r, k) = (k / 8) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2
// with rows r in [0,128), LBO = 2048 (one 8-column group of all 128 rows), SBO = 128.  A warp
// that writes rows r..r+31 of one 8-column group writes 512 contiguous bytes.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tib {
namespace tc {

constexpr int kRows = 128;                    // rows of every operand image / MMA M and N
constexpr uint32_t kLBO = 2048;               // bytes between 8-column groups
constexpr uint32_t kSBO = 128;                // bytes between 8-row groups
constexpr uint32_t kKStepBytes = 2 * kLBO;    // one MMA K step = 16 columns = two column groups
constexpr int kChunkK = 32;                   // columns per streamed weight chunk
constexpr uint32_t kChunkHalfBytes = kRows * kChunkK * 2;   // 8 KB (hi or lo image)
constexpr uint32_t kChunkBytes = 2 * kChunkHalfBytes;       // 16 KB: hi image then lo image
constexpr uint32_t kOperandHalfBytes = kRows * 128 * 2;     // 32 KB: [128 x 128] f16 image
constexpr uint32_t kOperandBytes = 2 * kOperandHalfBytes;   // 64 KB: hi image then lo image

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
// Bounded wait: a protocol bug must not hang the GPU.  On time-out the error word is set and
// every later wait returns at once, so the kernel drains (with garbage results) and the host
// reports the failure.
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, volatile int* err) {
  if (*err) return;
  for (uint32_t spins = 0; spins < (1u << 22); ++spins) {
    if (mbar_try_wait(bar, parity)) return;       // try_wait itself suspends the thread for a bounded time
    if ((spins & 1023u) == 1023u && *err) return;
  }
  *err = 1;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile int* err) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, err);
}

// wait + add the stalled cycles to *acc (pipeline diagnostics; acc may be a dummy)
// (mbarrier.try_wait itself may block for a hardware time limit, so the clock brackets the first try too)
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, volatile int* err, long long& acc,
                                                bool diag = true) {
  if (!diag) { mbar_wait(bar, parity, err); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity, err);
  acc += clock64() - t0;
}

// ---- bulk async copy global -> shared (TMA unit; completes on an mbarrier) -------------------------
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on `bar` when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; both operands K-major, f16 in, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo);
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand [128 x 16] f16 lives in TMEM (lane = row, 8 columns of packed
// f16 pairs: column c holds k = 2c in its low half and k = 2c + 1 in its high half)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 consecutive K values (one MMA K step) of this thread's row -> hi and lo A-operand columns in TMEM
__device__ __forceinline__ void tmem_store_kstep(uint32_t t_hi, uint32_t t_lo, const float (&v)[16]) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  tmem_st8(t_hi, h);
  tmem_st8(t_lo, l);
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// two 32-column loads under one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t ta, uint32_t tb, float (&a)[32], float (&b)[32]) {
  uint32_t r[32], q[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(ta)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
        "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
        "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
      : "r"(tb)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(q[i]); }
}

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// two 16-column loads under one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(tb)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(q[i]); }
}

// two 8-column loads (e.g. the phi and w accumulators of the same rows) under one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, int col, float (&a)[8], float (&b)[8]) {
  uint32_t r[8], q[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(ta + col)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
               : "r"(tb + col)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(q[i]); }
}

// sin and cos of a moderate fp32 argument (|a| < ~1e4): 3-term Cody-Waite reduction by pi/2 and the
// minimax polynomials of the Cephes sinf / cosf kernels on [-pi/4, pi/4] (~1 ulp); no slow path, no stack.
__device__ __forceinline__ void sincos_cw(float a, float& s, float& c) {
  const float q = rintf(a * 0.636619772367581343f);
  float r = fmaf(q, -1.5707962512969971f, a);
  r = fmaf(q, -7.5497894158615964e-08f, r);
  r = fmaf(q, -5.3903029534742384e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  const float sr = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  const float cr = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const int n = (int)q;
  const float ss = (n & 1) ? cr : sr, cc = (n & 1) ? sr : cr;
  s = (n & 2) ? -ss : ss;
  c = ((n + 1) & 2) ? -cc : cc;
}

// ---- UMMA descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 (no swizzle) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format=F32 [4,6), a/b_format=F16 (0),
// a/b K-major (0), N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t kIdesc128x128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// One K range of `ksteps` 16-column steps: D (+)= A * B^T in f16x3.  a_hi / b_hi are the shared
// addresses of the hi images at the first step; the lo images sit a_lo_off / b_lo_off bytes above.
__device__ __forceinline__ void mma_f16x3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo_off, uint32_t b_hi,
                                          uint32_t b_lo_off, int ksteps, bool accumulate_first, int passes) {
  uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t ah = make_desc(a_hi + ks * kKStepBytes), al = make_desc(a_hi + a_lo_off + ks * kKStepBytes);
    const uint64_t bh = make_desc(b_hi + ks * kKStepBytes), bl = make_desc(b_hi + b_lo_off + ks * kKStepBytes);
    tc_mma_f16(d_tmem, ah, bh, kIdesc128x128, acc);
    if (passes == 3) {
      tc_mma_f16(d_tmem, ah, bl, kIdesc128x128, 1u);
      tc_mma_f16(d_tmem, al, bh, kIdesc128x128, 1u);
    }
    acc = 1u;
  }
}

// ---- split-f16 packing ----------------------------------------------------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);            // one packed convert
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Raw state (s, v, e) enters the GEMMs scaled by 2^-4 and the accumulators are scaled back in the epilogue (both
// exact): a split-f16 operand overflows at |x| >= 65504, which the node features of a random-weight network reach
// (|s| ~ 1e5 after q^2-scaled updates); with the scale the bound is 1.05e6 and values of O(1) still keep ~21 bits
// (the f16 subnormal quantum 2^-24 becomes 2^-20 in true units).  Hidden activations (after LayerNorm + SiLU) and
// positional encodings are O(1) and enter unscaled.
constexpr float kStateScale = 0.0625f;
constexpr float kStateUnscale = 16.0f;

// z * sigmoid(z) with the two MUFU approximations (ex2, rcp): ~2 ulp
__device__ __forceinline__ float silu_fast(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return z * r;
}
// MUFU square root (~1 ulp), for values that are rounded to split-f16 operands right afterwards
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 8 consecutive columns (one column group) of row r -> hi and lo images of an operand buffer
__device__ __forceinline__ void store_group(unsigned char* op_hi, uint32_t lo_off, int r, int kgroup, const float (&v)[8]) {
  uint4 h, l;
  split2(v[0], v[1], h.x, l.x);
  split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z);
  split2(v[6], v[7], h.w, l.w);
  unsigned char* p = op_hi + (size_t)kgroup * kLBO + (size_t)r * 16;
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + lo_off) = l;
}

}  // namespace tc
}  // namespace tib
