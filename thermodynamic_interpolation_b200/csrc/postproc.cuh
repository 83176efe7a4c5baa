// postproc.cuh - internal coordinates of sampled conformers on the device.
//
// construct_z_matrix_batch (mdqm9/analysis/utils/z_matrix.py:56-102 with compute_distance / compute_angle /
// compute_torsion of mol_geometry.py:25-81): for placed atom a >= 1 with reference triplet (r3, r2, r1)
//   z[c][a-1][0] = | X[order[a]] - X[r3] |
//   z[c][a-1][1] = angle at X[r3] between X[order[a]] and X[r2]                       (a >= 2, else 0)
//   z[c][a-1][2] = atan2 torsion of (X[r1], X[r2], X[r3], X[order[a]]) in (-pi, pi]    (a >= 3, else 0)
// The torsions are the features of the reference's TICA / histogram analysis (results_00031.py:140-141,217-225).
// HBM bound: 12 n bytes read and 12 (n - 1) bytes written per conformer; one thread per (conformer, atom).
#pragma once
#include "common.cuh"

namespace tib {

__global__ void k_zmatrix(const float* __restrict__ x, long long n_conf, int n_atoms, const int* __restrict__ order,
                          const int* __restrict__ ref, float* __restrict__ z) {
  const long long total = n_conf * (long long)(n_atoms - 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / (n_atoms - 1);
    const int a = 1 + (int)(idx % (n_atoms - 1));
    const float* X = x + c * (long long)n_atoms * 3;
    const int i4 = __ldg(order + a), i3 = __ldg(ref + 3 * a), i2 = __ldg(ref + 3 * a + 1), i1 = __ldg(ref + 3 * a + 2);
    const float p4[3] = {X[3 * i4], X[3 * i4 + 1], X[3 * i4 + 2]};
    const float p3[3] = {X[3 * i3], X[3 * i3 + 1], X[3 * i3 + 2]};
    float out[3] = {0.0f, 0.0f, 0.0f};
    // compute_distance(x4, x3) = norm(x3 - x4)
    const float d34[3] = {p3[0] - p4[0], p3[1] - p4[1], p3[2] - p4[2]};
    out[0] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d34[0], d34[0]), __fmul_rn(d34[1], d34[1])), __fmul_rn(d34[2], d34[2])));
    if (a >= 2) {
      const float p2[3] = {X[3 * i2], X[3 * i2 + 1], X[3 * i2 + 2]};
      // compute_angle(x1 = p4, x2 = p3, x3 = p2): acos(<x1-x2, x3-x2> / (|x1-x2| |x3-x2|))
      const float u[3] = {p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]};
      const float w[3] = {p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]};
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(u[0], w[0]), __fmul_rn(u[1], w[1])), __fmul_rn(u[2], w[2]));
      const float nu = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(u[0], u[0]), __fmul_rn(u[1], u[1])), __fmul_rn(u[2], u[2])));
      const float nw = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2])));
      out[1] = acosf(__fdiv_rn(dot, __fmul_rn(nu, nw)));
      if (a >= 3) {
        const float p1[3] = {X[3 * i1], X[3 * i1 + 1], X[3 * i1 + 2]};
        // compute_torsion(x1 = p1, x2 = p2, x3 = p3, x4 = p4)
        const float x12[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        const float x23[3] = {p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2]};
        const float x34[3] = {p4[0] - p3[0], p4[1] - p3[1], p4[2] - p3[2]};
        const float c2334[3] = {__fmul_rn(x23[1], x34[2]) - __fmul_rn(x23[2], x34[1]),
                                __fmul_rn(x23[2], x34[0]) - __fmul_rn(x23[0], x34[2]),
                                __fmul_rn(x23[0], x34[1]) - __fmul_rn(x23[1], x34[0])};
        const float c1223[3] = {__fmul_rn(x12[1], x23[2]) - __fmul_rn(x12[2], x23[1]),
                                __fmul_rn(x12[2], x23[0]) - __fmul_rn(x12[0], x23[2]),
                                __fmul_rn(x12[0], x23[1]) - __fmul_rn(x12[1], x23[0])};
        const float n23 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x23[0], x23[0]), __fmul_rn(x23[1], x23[1])), __fmul_rn(x23[2], x23[2])));
        const float yy = __fmul_rn(n23, __fadd_rn(__fadd_rn(__fmul_rn(x12[0], c2334[0]), __fmul_rn(x12[1], c2334[1])), __fmul_rn(x12[2], c2334[2])));
        const float xx = __fadd_rn(__fadd_rn(__fmul_rn(c1223[0], c2334[0]), __fmul_rn(c1223[1], c2334[1])), __fmul_rn(c1223[2], c2334[2]));
        out[2] = atan2f(yy, xx);
      }
    }
    float* zo = z + idx * 3;
    zo[0] = out[0]; zo[1] = out[1]; zo[2] = out[2];
  }
}

// Torsion features -> TICA projection -> histograms (plots/10506_main.ipynb cells 3-4): per conformer
//   feat[2 j] = cos(torsion_j), feat[2 j + 1] = sin(torsion_j)              (enc(): np.stack((cos, sin), -1).reshape)
//   proj[k]   = sum_i (feat[i] - mean[i]) R[i][k],  k < dim <= 4             (deeptime TICA.transform)
//   hist[k][bin(proj[k])] += weight                                          (plt.hist(bins=linspace(lo, hi, n_bins + 1)))
// torsions: [n_conf][stride] floats with the torsion of feature j at column col0 + j * col_step (a z-matrix row block, or a
// plain [n_conf][n_tors] array).  HBM bound: 4 n_tors bytes read + 4 dim written per conformer; one thread per conformer,
// histogram counts accumulated in shared memory per block, then added to global memory with one atomic per bin.
__global__ void k_tica_project(const float* __restrict__ tors, long long n_conf, int n_tors, long long stride, int col0, int col_step,
                               const double* __restrict__ mean, const double* __restrict__ R, int dim, const double* __restrict__ weight,
                               float* __restrict__ proj, double* __restrict__ hist, int n_bins, double lo, double hi) {
  extern __shared__ double sh_hist[];           // [dim][n_bins]
  for (int i = threadIdx.x; i < dim * n_bins; i += blockDim.x) sh_hist[i] = 0.0;
  __syncthreads();
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_conf; c += (long long)gridDim.x * blockDim.x) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const float* t = tors + c * stride + col0;
    for (int j = 0; j < n_tors; ++j) {
      float sn, cs;
      sincosf(t[(long long)j * col_step], &sn, &cs);
      const double fc = (double)cs - mean[2 * j], fs = (double)sn - mean[2 * j + 1];
      for (int k = 0; k < dim; ++k) acc[k] += fc * R[(2 * j) * dim + k] + fs * R[(2 * j + 1) * dim + k];
    }
    const double w = weight ? weight[c] : 1.0;
    for (int k = 0; k < dim; ++k) {
      if (proj) proj[c * dim + k] = (float)acc[k];
      if (hist && acc[k] >= lo && acc[k] <= hi) {
        int b = (int)((acc[k] - lo) / (hi - lo) * n_bins);
        b = b >= n_bins ? n_bins - 1 : b;      // the right edge belongs to the last bin (numpy.histogram)
        atomicAdd(&sh_hist[k * n_bins + b], w);
      }
    }
  }
  __syncthreads();
  if (hist)
    for (int i = threadIdx.x; i < dim * n_bins; i += blockDim.x)
      if (sh_hist[i] != 0.0) atomicAdd(&hist[i], sh_hist[i]);
}

}  // namespace tib
