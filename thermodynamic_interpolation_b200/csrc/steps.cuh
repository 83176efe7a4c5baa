// steps.cuh - K1: fused integrator state updates (HBM-bound elementwise / reduction kernels).
// Every kernel is a grid-stride loop over a flat fp32 state with 128-bit accesses where the
// length allows; rounding follows the torch eager expressions of torchdiffeq 0.2.5 op by op
// (products and sums rounded separately unless the reference itself uses a dot product).
#pragma once
#include "common.cuh"

namespace tib {

// x_out = x + dt*b [+ (dt*eps)*score] [+ sqrt(2*eps*dt)*noise];  frame = x_out (optional)
// Euler: FixedGridODESolver.integrate `y1 = y0 + dt*f0` (solvers.py), product then sum.
__global__ void k_step_euler(const float* __restrict__ x, const float* __restrict__ b,
                             const float* __restrict__ score, const float* __restrict__ noise,
                             float dt, float dt_eps, float sig, float* __restrict__ x_out,
                             float* __restrict__ frame, size_t n, int vec_ok) {
  // 128-bit path only when every pointer is 16-byte aligned (frames of a [T,N,3] trajectory with
  // 3N % 4 != 0 are not)
  const size_t n4 = vec_ok ? n / 4 : 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 bv = reinterpret_cast<const float4*>(b)[i];
    float4 o;
    o.x = __fadd_rn(xv.x, __fmul_rn(dt, bv.x));
    o.y = __fadd_rn(xv.y, __fmul_rn(dt, bv.y));
    o.z = __fadd_rn(xv.z, __fmul_rn(dt, bv.z));
    o.w = __fadd_rn(xv.w, __fmul_rn(dt, bv.w));
    if (score) {
      const float4 sv = reinterpret_cast<const float4*>(score)[i];
      o.x = __fadd_rn(o.x, __fmul_rn(dt_eps, sv.x)); o.y = __fadd_rn(o.y, __fmul_rn(dt_eps, sv.y));
      o.z = __fadd_rn(o.z, __fmul_rn(dt_eps, sv.z)); o.w = __fadd_rn(o.w, __fmul_rn(dt_eps, sv.w));
    }
    if (noise) {
      const float4 zv = reinterpret_cast<const float4*>(noise)[i];
      o.x = __fadd_rn(o.x, __fmul_rn(sig, zv.x)); o.y = __fadd_rn(o.y, __fmul_rn(sig, zv.y));
      o.z = __fadd_rn(o.z, __fmul_rn(sig, zv.z)); o.w = __fadd_rn(o.w, __fmul_rn(sig, zv.w));
    }
    reinterpret_cast<float4*>(x_out)[i] = o;
    if (frame) reinterpret_cast<float4*>(frame)[i] = o;
  }
  // tail (n % 4 elements)
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float o = __fadd_rn(x[i], __fmul_rn(dt, b[i]));
    if (score) o = __fadd_rn(o, __fmul_rn(dt_eps, score[i]));
    if (noise) o = __fadd_rn(o, __fmul_rn(sig, noise[i]));
    x_out[i] = o;
    if (frame) frame[i] = o;
  }
}

// torchdiffeq rk4 = 3/8 rule (rk_common.rk4_alt_step_func); `third` = fp32(1/3).
//   mode 0: out = y + (dt*k1)*third
//   mode 1: out = y + dt*(k2 - k1*third)
//   mode 2: out = y + dt*((k1 - k2) + k3)
//   mode 3: out = y + (((k1 + 3*(k2+k3)) + k4)*dt)*0.125     (+ frame)
__global__ void k_rk4_stage(int mode, const float* __restrict__ y, const float* __restrict__ k1,
                            const float* __restrict__ k2, const float* __restrict__ k3,
                            const float* __restrict__ k4, float dt, float* __restrict__ out,
                            float* __restrict__ frame, size_t n) {
  const float third = (float)(1.0 / 3.0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float inc;
    if (mode == 0) inc = __fmul_rn(__fmul_rn(dt, k1[i]), third);
    else if (mode == 1) inc = __fmul_rn(dt, __fsub_rn(k2[i], __fmul_rn(k1[i], third)));
    else if (mode == 2) inc = __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[i], k2[i]), k3[i]));
    else inc = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1[i], __fmul_rn(3.0f, __fadd_rn(k2[i], k3[i]))), k4[i]), dt), 0.125f);
    const float o = __fadd_rn(y[i], inc);
    out[i] = o;
    if (frame) frame[i] = o;
  }
}

// ---- dopri5 (torchdiffeq rk_common._runge_kutta_step) -----------------------------------------
// right-hand side of the (x, dlogp) state: out[0, n3) *= mb (the drift, negated for reverse_ode), out[n3, n3 + n_mol) *= md
// (the divergence times -scale, or +scale)                                          (ode_wrapper.py:39-49,91)
__global__ void k_dlogp_rhs_finish(float* __restrict__ out, size_t n3, size_t n_mol, float mb, float md) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n3 + n_mol; i += stride) {
    if (i < n3) { if (mb != 1.0f) out[i] = __fmul_rn(out[i], mb); }
    else out[i] = __fmul_rn(out[i], md);
  }
}

struct StageCoef { float c[7]; int n; };

// out = y + sum_{s<n} k_s * c_s   with c_s = fp32(beta_s * dt) (k[..., :i+1].matmul(beta_i * dt))
__global__ void k_dopri_stage(const float* __restrict__ y, const float* __restrict__ k, size_t kstride,
                              StageCoef coef, float* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float dot = 0.0f;
#pragma unroll
    for (int s = 0; s < 7; ++s)
      if (s < coef.n) dot = fmaf(k[(size_t)s * kstride + i], coef.c[s], dot);
    out[i] = __fadd_rn(y[i], dot);
  }
}

// partial[blockIdx] = sum_i ( (sum_s k_s*c_s) / (atol + rtol*max(|y0|,|y1|)) )^2 in fp64
// (misc._compute_error_ratio; y1_error = k.matmul(dt * c_error)).
__global__ void k_dopri_error(const float* __restrict__ y0, const float* __restrict__ y1,
                              const float* __restrict__ k, size_t kstride, StageCoef cerr,
                              double rtol, double atol, double* __restrict__ partial, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float err = 0.0f;
#pragma unroll
    for (int s = 0; s < 7; ++s) err = fmaf(k[(size_t)s * kstride + i], cerr.c[s], err);
    const double tol = atol + rtol * (double)fmaxf(fabsf(y0[i]), fabsf(y1[i]));
    const double r = (double)err / tol;
    acc += r * r;
  }
  __shared__ double red[32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

// partial[blockIdx] = sum_i ( (a_i - b_i) / (atol + |y_i|*rtol) )^2 in fp64; b may be NULL.
// Used by misc._select_initial_step for d0, d1, d2.
__global__ void k_scaled_sq(const float* __restrict__ a, const float* __restrict__ b,
                            const float* __restrict__ y, double rtol, double atol,
                            double* __restrict__ partial, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float num = b ? __fsub_rn(a[i], b[i]) : a[i];
    const double scale = atol + (double)fabsf(y[i]) * rtol;
    const double r = (double)num / scale;
    acc += r * r;
  }
  __shared__ double red[32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

// out[0] = sum(partial[0..m)) in a fixed order (deterministic).
__global__ void k_reduce_partials(const double* __restrict__ partial, int m, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) acc += partial[i];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) out[0] = v;
  }
}

// Dense output of an accepted step (interp._interp_fit / _interp_evaluate, rk_common._interp_fit):
// y_mid = y0 + sum_s k_s * fp32(dt*c_mid_s); quartic through (y0, y_mid, y1, f0=k_0, f1=k_6);
// evaluated at up to 8 normalised times xs[] -> frames[j].
struct DenseArgs { float xs[8]; float* frames[8]; int n; };

__global__ void k_dopri_dense(const float* __restrict__ y0, const float* __restrict__ y1,
                              const float* __restrict__ k, size_t kstride, StageCoef cmid, float dt,
                              DenseArgs d, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float dot = 0.0f;
#pragma unroll
    for (int s = 0; s < 7; ++s) dot = fmaf(k[(size_t)s * kstride + i], cmid.c[s], dot);
    const float a0 = y0[i], a1 = y1[i];
    const float ym = __fadd_rn(a0, dot);
    const float f0 = k[i], f1 = k[(size_t)6 * kstride + i];
    // a = 2*dt*(f1-f0) - 8*(y1+y0) + 16*y_mid
    const float ca = __fadd_rn(__fsub_rn(__fmul_rn(__fmul_rn(2.0f, dt), __fsub_rn(f1, f0)), __fmul_rn(8.0f, __fadd_rn(a1, a0))), __fmul_rn(16.0f, ym));
    // b = dt*(5*f0 - 3*f1) + 18*y0 + 14*y1 - 32*y_mid
    const float cb = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(dt, __fsub_rn(__fmul_rn(5.0f, f0), __fmul_rn(3.0f, f1))), __fmul_rn(18.0f, a0)), __fmul_rn(14.0f, a1)), __fmul_rn(32.0f, ym));
    // c = dt*(f1 - 4*f0) - 11*y0 - 5*y1 + 16*y_mid
    const float cc = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(dt, __fsub_rn(f1, __fmul_rn(4.0f, f0))), __fmul_rn(11.0f, a0)), __fmul_rn(5.0f, a1)), __fmul_rn(16.0f, ym));
    const float cd = __fmul_rn(dt, f0);
    for (int j = 0; j < d.n; ++j) {
      const float xx = d.xs[j];
      float total = __fadd_rn(a0, __fmul_rn(xx, cd));
      float xp = __fmul_rn(xx, xx);
      total = __fadd_rn(total, __fmul_rn(xp, cc));
      xp = __fmul_rn(xp, xx);
      total = __fadd_rn(total, __fmul_rn(xp, cb));
      xp = __fmul_rn(xp, xx);
      total = __fadd_rn(total, __fmul_rn(xp, ca));
      d.frames[j][i] = total;
    }
  }
}

// ---- reweighting statistics (ess.py:8-10,32-35; free_energy.py:41-46) -------------------------
// partial[block*5 + {0..4}] = sum w, sum w^2, sum exp(-phi)*weight, sum weight, count
__global__ void k_reweight_partials(const double* __restrict__ E0, const double* __restrict__ E1,
                                    const double* __restrict__ nd, const double* __restrict__ wt,
                                    size_t n, double* __restrict__ partial) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double a[5] = {0, 0, 0, 0, 0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double phi = E1[i] - E0[i] + (nd ? nd[i] : 0.0);
    const double w = exp(-phi);
    const double g = wt ? wt[i] : 1.0;
    a[0] += w; a[1] += w * w; a[2] += w * g; a[3] += g; a[4] += 1.0;
  }
  __shared__ double red[5][32];
  for (int c = 0; c < 5; ++c) {
    const double v = warp_sum_d(a[c]);
    if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int c = 0; c < 5; ++c) {
      double v = threadIdx.x < (blockDim.x >> 5) ? red[c][threadIdx.x] : 0.0;
      v = warp_sum_d(v);
      if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 5 + c] = v;
    }
  }
}

__global__ void k_reweight_final(const double* __restrict__ partial, int m, double* __restrict__ out) {
  const int c = threadIdx.x;
  if (c < 5) {
    double acc = 0.0;
    for (int i = 0; i < m; ++i) acc += partial[(size_t)i * 5 + c];
    out[c] = acc;
  }
}

}  // namespace tib
