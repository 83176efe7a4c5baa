// simt_mlp.cuh - fp32 CUDA-core building blocks for the reference MLP
// (Linear -> LayerNorm -> SiLU -> Linear -> LayerNorm -> SiLU -> Linear; embedding.py:26-34).
//
// Work decomposition: a CTA of 8 warps owns a tile of TR = 8*RPT rows (edges or nodes) whose
// activations live in shared memory.  Row r = q*8 + warp is owned by one warp for the whole MLP
// chain, so consecutive layers need only __syncwarp(), never a CTA barrier.  Inside a warp, lane l
// owns columns [ch*CW + l*CPL, +CPL) of every chunk ch: the weight row Wt[k][...] is one coalesced
// 128-bit load per lane (weights are stored transposed, [K][N]; all 8 warps hit the same lines so
// L1 serves 7 of 8), and the activation A[r][k] is a shared-memory broadcast.  Per 4 values of k a
// thread issues RPT LDS.128 + 4 LDG.128 for 16*RPT FMAs.
#pragma once
#include "common.cuh"

namespace tib {

// acc[q][c] = sum_k A[(q*8+warp)*lda + k] * Wt[k*ldw + col + c]      (k ascending, fmaf chain)
// KU weight rows are in flight per lane (double buffered); KU = 16 is for launches with a handful of
// CTAs where the loop is bound by the latency of the weight loads, not by their throughput.
template <int RPT, int CPL, int KU = 4>
__device__ __forceinline__ void gemm_rows(float (&acc)[RPT][CPL], const float* A, int lda, int K,
                                          const float* __restrict__ Wt, int ldw, int col, int warp) {
  static_assert(KU % 4 == 0, "KU must be a multiple of 4");
#pragma unroll
  for (int q = 0; q < RPT; ++q)
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[q][c] = 0.0f;
  const float* wp = Wt + col;
  const float* ap = A + warp * lda;
  float w[KU][CPL], wn[KU][CPL];
#pragma unroll
  for (int kk = 0; kk < KU; ++kk) ldg_vec<CPL>(w[kk], wp + (size_t)kk * ldw);
  for (int k0 = 0; k0 < K; k0 += KU) {
    if (k0 + KU < K) {
#pragma unroll
      for (int kk = 0; kk < KU; ++kk) ldg_vec<CPL>(wn[kk], wp + (size_t)(k0 + KU + kk) * ldw);
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
#pragma unroll
      for (int k4 = 0; k4 < KU; k4 += 4) {
        const float4 a = *reinterpret_cast<const float4*>(ap + q * 8 * lda + k0 + k4);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float t = acc[q][c];
          t = fmaf(a.x, w[k4 + 0][c], t);
          t = fmaf(a.y, w[k4 + 1][c], t);
          t = fmaf(a.z, w[k4 + 2][c], t);
          t = fmaf(a.w, w[k4 + 3][c], t);
          acc[q][c] = t;
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < KU; ++kk)
#pragma unroll
      for (int c = 0; c < CPL; ++c) w[kk][c] = wn[kk][c];
  }
}

// Hidden layer: Xout[r][:] = SiLU(LayerNorm(A[r][:] @ Wt + b) * g + be) for the warp's rows.
// LayerNorm (eps 1e-5, biased variance, two-pass) is done in registers across the warp.
template <int F, int RPT, int KU = 4>
__device__ __forceinline__ void layer_ln_silu(const float* A, int lda, int K, const float* __restrict__ Wt,
                                              const float* __restrict__ b, const float* __restrict__ g,
                                              const float* __restrict__ be, float* Xout, int ldo,
                                              int warp, int lane) {
  using C = Cols<F>;
  float acc[C::NCH][RPT][C::CPL];
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch)
    gemm_rows<RPT, C::CPL, KU>(acc[ch], A, lda, K, Wt, F, ch * C::CW + lane * C::CPL, warp);
  float bb[C::NCH][C::CPL], gg[C::NCH][C::CPL], ee[C::NCH][C::CPL];
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch) {
    ldg_vec<C::CPL>(bb[ch], b + ch * C::CW + lane * C::CPL);
    ldg_vec<C::CPL>(gg[ch], g + ch * C::CW + lane * C::CPL);
    ldg_vec<C::CPL>(ee[ch], be + ch * C::CW + lane * C::CPL);
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    float sum = 0.0f;
#pragma unroll
    for (int ch = 0; ch < C::NCH; ++ch)
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) {
        acc[ch][q][c] += bb[ch][c];
        sum += acc[ch][q][c];
      }
    const float mean = warp_sum(sum) * (1.0f / F);
    float ss = 0.0f;
#pragma unroll
    for (int ch = 0; ch < C::NCH; ++ch)
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) {
        const float d = acc[ch][q][c] - mean;
        acc[ch][q][c] = d;
        ss = fmaf(d, d, ss);
      }
    const float var = warp_sum(ss) * (1.0f / F);
    const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(var + 1e-5f));
    float* orow = Xout + (size_t)(q * 8 + warp) * ldo;
#pragma unroll
    for (int ch = 0; ch < C::NCH; ++ch) {
      float y[C::CPL];
#pragma unroll
      for (int c = 0; c < C::CPL; ++c)
        y[c] = silu(__fadd_rn(__fmul_rn(__fmul_rn(acc[ch][q][c], rstd), gg[ch][c]), ee[ch][c]));
      st_vec<C::CPL>(orow + ch * C::CW + lane * C::CPL, y);
    }
  }
  __syncwarp();
}

// Output layer, one chunk of CW columns starting at column c0 of W3t [F][n_out]:
// acc[q][c] = H[r][:] @ W3t[:, c0 + lane*CPL + c] + b3[...]
template <int F, int RPT, int KU = 4>
__device__ __forceinline__ void out_chunk(float (&acc)[RPT][Cols<F>::CPL], const float* H, int ldh,
                                          const float* __restrict__ W3t, int n_out,
                                          const float* __restrict__ b3, int c0, int warp, int lane) {
  using C = Cols<F>;
  gemm_rows<RPT, C::CPL, KU>(acc, H, ldh, F, W3t, n_out, c0 + lane * C::CPL, warp);
  float bb[C::CPL];
  ldg_vec<C::CPL>(bb, b3 + c0 + lane * C::CPL);
#pragma unroll
  for (int q = 0; q < RPT; ++q)
#pragma unroll
    for (int c = 0; c < C::CPL; ++c) acc[q][c] += bb[c];
}

}  // namespace tib
