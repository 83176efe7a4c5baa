// common.cuh - shared device helpers for libtib (sm_100a).
// Compiled with -fmad=false: every fused multiply-add in this library is an explicit fmaf(), and
// every place where the reference rounds a product and a sum separately (torch eager ops) stays
// un-fused.  No --use_fast_math: sincosf/expf/sqrtf/div are the IEEE-accurate versions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TIB_THREADS 256
#define TIB_WARPS 8
#define TIB_MAX_ATOMS 64

namespace tib {

constexpr float kPiF = 3.14159265358979323846f;  // np.pi rounded to fp32 (embedding.py:156-157)

// Programmatic dependent launch: a kernel first lets the NEXT kernel of the stream be scheduled (it runs its prologue and
// parks at its own wait), then - before it first touches data the previous kernels produce - waits until they have completed
// and their writes are visible.  Only launches that carry cudaLaunchAttributeProgrammaticStreamSerialization start early;
// for every other launch both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// PositionalEncoder argument ((x / len) * rank) * pi, each op rounded to fp32
// (mdqm9/thermo/ambient/models/embedding.py:156-157).
__device__ __forceinline__ float pe_arg(float x, float len, int rank) {
  return __fmul_rn(__fmul_rn(__fdiv_rn(x, len), (float)rank), kPiF);
}

__device__ __forceinline__ float silu(float y) {
  // torch.nn.SiLU: y * sigmoid(y)
  return __fdiv_rn(y, 1.0f + expf(-y));
}

template <int CPL>
__device__ __forceinline__ void ldg_vec(float (&w)[CPL], const float* __restrict__ p) {
  if constexpr (CPL == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else if constexpr (CPL == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    w[0] = t.x; w[1] = t.y;
  } else {
    w[0] = __ldg(p);
  }
}

template <int CPL>
__device__ __forceinline__ void st_vec(float* p, const float (&w)[CPL]) {
  if constexpr (CPL == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]);
  } else if constexpr (CPL == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(w[0], w[1]);
  } else {
    p[0] = w[0];
  }
}

// Column tiling of a feature row of width F over one warp: chunks of CW columns, CPL per lane.
template <int F>
struct Cols {
  static_assert(F == 32 || F == 64 || F == 128 || F == 256, "n_features must be 32, 64, 128 or 256");
  static constexpr int CW = F < 128 ? F : 128;
  static constexpr int CPL = CW / 32;
  static constexpr int NCH = F / CW;
};

// One MLP (embedding.py:26-34) with transposed weights: W1t [k_in][F], W2t [F][F], W3t [F][n_out].
struct MlpW {
  const float *W1t, *b1, *g1, *be1;
  const float *W2t, *b2, *g2, *be2;
  const float *W3t, *b3;
  int k_in, n_out;
};

}  // namespace tib
