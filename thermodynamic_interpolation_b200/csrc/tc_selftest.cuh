// tc_selftest.cuh - single-tile self test of the tensor-core plumbing shared by the drift kernels.
#pragma once
#include "tc_common.cuh"

namespace tib {
namespace tc {

constexpr int kSelfThreads = 192;
constexpr int kSelfStages = 4;

// rows [32*warp, +32) x 16 column groups of an operand image from row-major fp32 global rows:
// lane = (row & 7, group quad) so each 128 B line of a row is read by 4 lanes and every store
// instruction writes 4 x 128 contiguous bytes.
template <typename RowPtr>
__device__ __forceinline__ void selftest_build(unsigned char* op, int warp, int lane, int rows, RowPtr row_ptr) {
#pragma unroll 1
  for (int oct = 0; oct < 4; ++oct) {
    const int r = 32 * warp + 8 * oct + (lane & 7);
    const float* src = r < rows ? row_ptr(r) : nullptr;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const int g = 4 * kq + (lane >> 3);
      float v[8];
      if (src) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + g * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + g * 8 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.0f;
      }
      store_group(op, kOperandHalfBytes, r, g, v);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Self test of the tensor-core plumbing (descriptors, operand image, ring, TMEM addressing):
//   transposed = 0:  out[r][c] = sum_k A[r][k] * W[c][k]      (lane = row of A)
//   transposed = 1:  out[r][c] = sum_k W[r][k] * A[c][k]      (lane = row of W)
//   transposed = 2:  as 0, but the A operand is written to TMEM (tcgen05.st) and read from there
// A is fp32 [128][128] row-major, wchunks = 4 packed chunks of W [128][128], out fp32 [128][128].
__global__ void __launch_bounds__(kSelfThreads, 1) k_tc_selftest(const float* __restrict__ A, const unsigned char* __restrict__ wchunks,
                                                             float* __restrict__ out, int transposed, int* err_ptr) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const X = smem;
  unsigned char* const RING = smem + kOperandBytes;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kOperandBytes + kSelfStages * kChunkBytes);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = err_ptr;
  if (tid == 0) {
    for (int i = 0; i < kSelfStages; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[4], 128);   // operand full
    mbar_init(&bars[5], 1);     // accumulator full
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 4) {
    if (lane == 0)
      for (int c = 0; c < 4; ++c) {
        mbar_arrive_expect_tx(&bars[c], kChunkBytes);
        bulk_g2s(RING + c * kChunkBytes, wchunks + (size_t)c * kChunkBytes, kChunkBytes, &bars[c]);
      }
  } else if (warp == 5) {
    if (lane == 0) {
      mbar_wait(&bars[4], 0, err);
      tc_fence_after();
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&bars[kb], 0, err);
        tc_fence_after();
        const uint32_t wst = smem_u32(RING) + kb * kChunkBytes, opk = smem_u32(X) + kb * (2 * kKStepBytes);
        if (transposed == 0) mma_f16x3(tmem, opk, kOperandHalfBytes, wst, kChunkHalfBytes, 2, kb > 0, 3);
        else if (transposed == 1) mma_f16x3(tmem, wst, kChunkHalfBytes, opk, kOperandHalfBytes, 2, kb > 0, 3);
        else {
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a_hi = tmem + 128 + 8 * (2 * kb + ks), a_lo = a_hi + 64;
            const uint64_t bh = make_desc(wst + ks * kKStepBytes), bl = make_desc(wst + kChunkHalfBytes + ks * kKStepBytes);
            tc_mma_f16_ts(tmem, a_hi, bh, kIdesc128x128, (kb | ks) ? 1u : 0u);
            tc_mma_f16_ts(tmem, a_hi, bl, kIdesc128x128, 1u);
            tc_mma_f16_ts(tmem, a_lo, bh, kIdesc128x128, 1u);
          }
        }
      }
      tc_commit(&bars[5]);
    }
  } else {
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    if (transposed == 2) {
      for (int ks = 0; ks < 8; ++ks) {
        float v[16];
        for (int i = 0; i < 16; ++i) v[i] = A[(size_t)tid * 128 + 16 * ks + i];
        tmem_store_kstep(lane_base + 128 + 8 * ks, lane_base + 192 + 8 * ks, v);
      }
      tmem_wait_st();
      tc_fence_before();
    } else {
      selftest_build(X, warp, lane, 128, [&](int r) { return A + (size_t)r * 128; });
      fence_proxy_async();
    }
    mbar_arrive(&bars[4]);
    mbar_wait(&bars[5], 0, err);
    tc_fence_after();
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) {
      float t[32];
      tmem_ld32(lane_taddr + 32 * c, t);
      for (int i = 0; i < 32; ++i) out[(size_t)tid * 128 + 32 * c + i] = t[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}


}  // namespace tc
}  // namespace tib
