// train_gemm.cuh - general fp32-faithful GEMM on tcgen05 for the training step (forward, data gradients, weight gradients):
//
//     C[m][n] (=, +=, atomic +=)  alpha * sum_k A(m, k) * B(n, k)  (+ bias[n])
//
// Both operands are read straight from row-major fp32 tensors in global memory - there is no pre-packed weight image,
// because the weights change every optimiser step - and are turned into split-f16 operand images (tc_common.cuh) in
// shared memory by the CTA's own threads.  Each operand is described by (ptr, ld, idx, trans):
//     trans = 0:  Op(r, k) = ptr[row(r) * ld + k]      r indexes source rows (optionally gathered), k is contiguous
//     trans = 1:  Op(r, k) = ptr[row(k) * ld + r]      the contraction runs over source ROWS (optionally gathered)
// which covers the three GEMM forms of a Linear layer y = x W^T  (W [out][in] as the reference stores it):
//     forward          y [R][out]  : A = x (trans 0),   B = W (trans 0)            K = in
//     data gradient    dx [R][in]  : A = dy (trans 0),  B = W (trans 1)            K = out
//     weight gradient  dW [out][in]: A = dy (trans 1),  B = x (trans 1)            K = R   (split over CTAs, atomic)
// A trans-0 tile is built row-per-lane-octet (32 B per lane, 128 B per row and chunk); a trans-1 tile is built with
// lanes along r, so that every load instruction reads 128 contiguous bytes and every store writes 512.
//
// Range: f16 overflows at 65504 and loses relative precision below 2^-14, so every operand carries a power-of-two
// scale: a constant (states 2^-4, hidden activations / weights 1) or, for gradient tensors, one derived on the device
// from the tensor's |max| (tracked by the producing kernel with atomicMax), mapping it into [2^12, 2^13).  The scale is
// undone on the accumulators (exact).
//
// Pipeline: 256 threads build K chunks of 32 into a 2-stage ring (32 KB per stage); thread 0 issues the three MMA
// passes of a chunk and commits them to the stage's mbarrier; the next chunk is built while they run.  64 KB of shared
// memory and 128 TMEM columns per CTA, so three CTAs share an SM and cover each other's build / issue gaps.
#pragma once
#include "tc_common.cuh"

namespace tib {
namespace train {

struct GemmOperand {
  const float* ptr;
  long long ld;
  const int* idx;       // optional gather of source rows
  int trans;
  const float* amax;    // DEVICE |max| of the tensor (dynamic scale) or nullptr (use `scale`)
  float scale;
};

enum { GEMM_STORE = 0, GEMM_ATOMIC = 1, GEMM_ACCUM = 2 };

struct GemmP {
  int M, N, K;
  GemmOperand A, B;
  float* C;
  long long ldc;
  const int* c_idx;     // optional: result row m is added to row c_idx[m] of C (GEMM_ATOMIC only)
  const float* bias;    // [N] or nullptr
  int mode;             // GEMM_*
  int k_splits;         // == gridDim.z; > 1 requires GEMM_ATOMIC
  float alpha;
  int passes;           // 3 = split-f16 x3, 1 = single f16 pass
  int* err;
};

constexpr int kGemmThreads = 256;
constexpr int kGemmKC = 32;                                   // K per chunk
constexpr int kGemmStages = 2;
constexpr uint32_t kGemmHalf = tc::kChunkHalfBytes;           // 8 KB: one [128 x 32] f16 image
constexpr uint32_t kGemmStageBytes = 4 * kGemmHalf;           // A hi, A lo, B hi, B lo
constexpr uint32_t kGemmSmem = kGemmStages * kGemmStageBytes + 64;

// power-of-two scale that maps amax into [2^12, 2^13); 1 for an all-zero (or non-finite) tensor
__device__ __forceinline__ float pow2_scale_for(float amax) {
  if (!(amax > 0.0f) || !(amax < 3.0e38f)) return 1.0f;
  int ex;
  frexpf(amax, &ex);                       // amax = f * 2^ex, f in [0.5, 1)
  int p = 13 - ex;
  p = p > 100 ? 100 : (p < -100 ? -100 : p);
  return ldexpf(1.0f, p);
}
__device__ __forceinline__ float operand_scale(const GemmOperand& o) { return o.amax ? pow2_scale_for(__ldg(o.amax)) : o.scale; }

// One [128 x 32] operand tile (rows r0.., columns k0..) -> hi / lo images at `img` / `img + kGemmHalf`.
__device__ __forceinline__ void build_tile(unsigned char* img, const GemmOperand& o, int r0, int r_end, int k0, int k_end,
                                           float scale, int warp, int lane) {
  if (o.trans == 0) {
    // lane = (row & 7, column group): a warp covers 8 rows x 32 columns per step; two steps per warp cover its 16 rows
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int rl = 16 * warp + 8 * half + (lane & 7), g = lane >> 3;
      const int r = r0 + rl, k = k0 + 8 * g;
      float v[8];
      if (r < r_end && k < k_end) {
        const long long row = o.idx ? (long long)__ldg(o.idx + r) : (long long)r;
        const float4* src = reinterpret_cast<const float4*>(o.ptr + row * o.ld + k);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
        v[4] = b.x * scale; v[5] = b.y * scale; v[6] = b.z * scale; v[7] = b.w * scale;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.0f;
      }
      tc::store_group(img, kGemmHalf, rl, g, v);
    }
  } else {
    // lanes along r: warp w covers rows 32 (w & 3) .. +31 and the column groups 2 (w >> 2), 2 (w >> 2) + 1
    const int rl = 32 * (warp & 3) + lane, r = r0 + rl;
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      const int g = 2 * (warp >> 2) + gg;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + 8 * g + i;
        float x = 0.0f;
        if (r < r_end && k < k_end) {
          const long long row = o.idx ? (long long)__ldg(o.idx + k) : (long long)k;
          x = __ldg(o.ptr + row * o.ld + r) * scale;
        }
        v[i] = x;
      }
      tc::store_group(img, kGemmHalf, rl, g, v);
    }
  }
}

__global__ void __launch_bounds__(kGemmThreads, 2) k_gemm_tc(const GemmP p) {
  using namespace tc;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);   // [0..1] stage free, [2] done
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;

  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128;
  const int chunks_total = (p.K + kGemmKC - 1) / kGemmKC;
  const int cps = (chunks_total + p.k_splits - 1) / p.k_splits;
  const int c_begin = blockIdx.z * cps, c_end = min(chunks_total, c_begin + cps);
  if (c_begin >= c_end) return;                                   // uniform over the CTA

  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float sa = operand_scale(p.A), sb = operand_scale(p.B);

  for (int c = c_begin; c < c_end; ++c) {
    const int it = c - c_begin, st = it & 1;
    if (it >= kGemmStages) mbar_wait(&bars[st], ((it >> 1) - 1) & 1, err);      // the MMAs that read this stage are done
    unsigned char* const stage = smem + st * kGemmStageBytes;
    const int k0 = c * kGemmKC;
    build_tile(stage, p.A, m0, p.M, k0, p.K, sa, warp, lane);
    build_tile(stage + 2 * kGemmHalf, p.B, n0, p.N, k0, p.K, sb, warp, lane);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a = smem_u32(stage), b = a + 2 * kGemmHalf;
      mma_f16x3(tmem, a, kGemmHalf, b, kGemmHalf, kGemmKC / 16, it > 0, p.passes);
      tc_commit(&bars[st]);
      if (c + 1 == c_end) tc_commit(&bars[2]);
    }
  }
  mbar_wait(&bars[2], 0, err);
  tc_fence_after();

  // ---- epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31 (rows) and columns 64 (w >> 2) .. +63
  const int m = m0 + 32 * (warp & 3) + lane;
  const float unscale = p.alpha / (sa * sb);
  const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
  const long long crow = (m < p.M) ? (p.c_idx ? (long long)__ldg(p.c_idx + m) : (long long)m) : 0;
#pragma unroll 1
  for (int blk = 0; blk < 2; ++blk) {
    const int cn = 64 * (warp >> 2) + 32 * blk;
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cn, v);
    if (m < p.M) {
      float* dst = p.C + crow * p.ldc + n0 + cn;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = n0 + cn + i;
        if (n < p.N) {
          float r = v[i] * unscale;
          if (add_bias) r += __ldg(p.bias + n);
          if (p.mode == GEMM_STORE) dst[i] = r;
          else if (p.mode == GEMM_ACCUM) dst[i] += r;
          else atomicAdd(dst + i, r);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace train
}  // namespace tib
