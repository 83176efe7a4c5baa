// adw.cuh - asymmetric double well: FCNetMultiBeta drift b(x, t, beta0, beta1) and its exact 1-D
// divergence d b / d x in fp64 (adw/thermo/models/simple.py:38-41; ode_wrapper.py:55-67).
// The reference differentiates b with autograd; for a scalar state the same number is the forward
// tangent pushed through the net (the beta embedding does not depend on x), so one pass yields both.
// One CTA = kAdwRows samples, thread t owns hidden unit t (H = 256); activations and tangents are
// ping-ponged through shared memory, hidden weights are stored transposed ([in][out]) so the
// per-k weight read is coalesced and the activation read is a broadcast.
#pragma once
#include "common.cuh"

namespace tib {

constexpr int kAdwRows = 16;
constexpr int kAdwMaxHidden = 8;

struct AdwW {
  const double *e_W1, *e_b1, *e_W2t, *e_b2, *e_W3, *e_b3;   // beta_embed: 3->H, H->H, H->1
  const double *n_W1, *n_b1;                                // net: 3->H
  const double* n_Wt[kAdwMaxHidden];                        // hidden H->H (transposed)
  const double* n_b[kAdwMaxHidden];
  const double *n_Wo, *n_bo;                                // H->1
  int n_hidden;
};

template <int H> constexpr size_t adw_smem() { return sizeof(double) * (4 * (size_t)kAdwRows * H + 2 * kAdwRows); }

__device__ __forceinline__ double sigmoid_d(double z) { return 1.0 / (1.0 + exp(-z)); }

template <int H>
__global__ void __launch_bounds__(H, 1) k_adw(AdwW w, const double* __restrict__ x, const double* __restrict__ beta0,
                                              const double* __restrict__ beta1, float t, double* __restrict__ out_b,
                                              double* __restrict__ out_div, size_t n) {
  constexpr int R = kAdwRows;
  extern __shared__ __align__(16) double sm[];
  double* A0 = sm;              // [R][H]
  double* A1 = A0 + R * H;
  double* D0 = A1 + R * H;      // tangents
  double* D1 = D0 + R * H;
  double* EMB = D1 + R * H;     // [R]
  double* XR = EMB + R;         // [R]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t row0 = (size_t)blockIdx.x * R;
  const double td = (double)t;

  // beta_embed layer 1: cat[beta0, beta1, t] (simple.py:39)
  {
    const double w0 = w.e_W1[tid * 3 + 0], w1 = w.e_W1[tid * 3 + 1], w2 = w.e_W1[tid * 3 + 2], b = w.e_b1[tid];
    for (int r = 0; r < R; ++r) {
      const size_t i = row0 + r;
      double z = 0.0;
      if (i < n) z = w0 * beta0[i] + w1 * beta1[i] + w2 * td + b;
      A0[r * H + tid] = z * sigmoid_d(z);
    }
    if (tid < R) XR[tid] = (row0 + tid < n) ? x[row0 + tid] : 0.0;
  }
  __syncthreads();
  // beta_embed layer 2
  {
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    for (int k = 0; k < H; ++k) {
      const double wv = w.e_W2t[(size_t)k * H + tid];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fma(A0[r * H + k], wv, acc[r]);
    }
    const double b = w.e_b2[tid];
#pragma unroll
    for (int r = 0; r < R; ++r) { const double z = acc[r] + b; A1[r * H + tid] = z * sigmoid_d(z); }
  }
  __syncthreads();
  // beta_embed output (scalar per sample)
  for (int r = warp; r < R; r += H / 32) {
    double a = 0.0;
    for (int k = lane; k < H; k += 32) a = fma(A1[r * H + k], w.e_W3[k], a);
    a = warp_sum_d(a);
    if (lane == 0) EMB[r] = a + w.e_b3[0];
  }
  __syncthreads();
  // net layer 1: cat[x, t, emb] (simple.py:40); tangent wrt x is the first weight column
  {
    const double w0 = w.n_W1[tid * 3 + 0], w1 = w.n_W1[tid * 3 + 1], w2 = w.n_W1[tid * 3 + 2], b = w.n_b1[tid];
    for (int r = 0; r < R; ++r) {
      const double z = w0 * XR[r] + w1 * td + w2 * EMB[r] + b;
      const double sg = sigmoid_d(z);
      A0[r * H + tid] = z * sg;
      D0[r * H + tid] = sg * (1.0 + z * (1.0 - sg)) * w0;
    }
  }
  __syncthreads();
  double* Ain = A0; double* Din = D0; double* Aout = A1; double* Dout = D1;
  for (int l = 0; l < w.n_hidden; ++l) {
    double acc[R], dac[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r] = 0.0; dac[r] = 0.0; }
    const double* Wt = w.n_Wt[l];
    for (int k = 0; k < H; ++k) {
      const double wv = Wt[(size_t)k * H + tid];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r] = fma(Ain[r * H + k], wv, acc[r]);
        dac[r] = fma(Din[r * H + k], wv, dac[r]);
      }
    }
    const double b = w.n_b[l][tid];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double z = acc[r] + b;
      const double sg = sigmoid_d(z);
      Aout[r * H + tid] = z * sg;
      Dout[r * H + tid] = sg * (1.0 + z * (1.0 - sg)) * dac[r];
    }
    __syncthreads();
    double* tmp = Ain; Ain = Aout; Aout = tmp;
    tmp = Din; Din = Dout; Dout = tmp;
  }
  for (int r = warp; r < R; r += H / 32) {
    double a = 0.0, d = 0.0;
    for (int k = lane; k < H; k += 32) {
      const double wo = w.n_Wo[k];
      a = fma(Ain[r * H + k], wo, a);
      d = fma(Din[r * H + k], wo, d);
    }
    a = warp_sum_d(a); d = warp_sum_d(d);
    const size_t i = row0 + r;
    if (lane == 0 && i < n) {
      out_b[i] = a + w.n_bo[0];
      if (out_div) out_div[i] = d;
    }
  }
}

}  // namespace tib
