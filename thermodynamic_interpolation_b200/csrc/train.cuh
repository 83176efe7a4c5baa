// train.cuh - element-wise / gather / scatter kernels of the training step: everything of
//   loss = StandardVelocityLoss(LinearInterpolant)(batch0, batch1, cPaiNN)        mdqm9/thermo/ambient/losses.py:30-85,126-133
//   loss.backward(); clip_grad_norm_(params, 1); Adam.step()                       mdqm9/train_ambient.py:144-148
// that is not a dense contraction (those run on tcgen05, train_gemm.cuh).  The two antithetic evaluations of the drift
// (x_t^+, x_t^-) are one batch of 2 B molecules: N2 = 2 N nodes, E2 = 2 E edge rows.  Layouts: s [N2][F], v [N2][3][F]
// (xyz-major planes), e [E2][F] with edge rows ordered by (dst, src) so that the incoming edges of a node are contiguous;
// the w MLP depends on an edge only through its length, so its rows are the P2 = E2 / 2 undirected pairs.
// Every kernel that produces a tensor a later GEMM reads as an operand records the tensor's |max| (atomicMax on the
// float bits) - the GEMM derives its power-of-two operand scale from it.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace tib {
namespace train {

constexpr int kEW = 128;   // threads of the element-wise kernels

// Programmatic dependent launch (common.cuh; train_api.cu::launch_pdl sets the attribute on the small launches): the
// element-wise kernels trigger and wait at entry.
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }

// sigmoid / SiLU with the MUFU approximations (ex2, rcp; ~2 ulp) - the IEEE expf + divide chain made the LayerNorm kernels
// instruction-bound (60 instructions per element); the sampling path's tensor-core epilogues use the same approximation
__device__ __forceinline__ float sigmoid_fast(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float silu_mufu(float z) { return z * sigmoid_fast(z); }

__device__ __forceinline__ void atomic_amax(float* slot, float v) {       // v >= 0
  atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(v));
}
// One atomicMax per BLOCK (every thread of the block must call it): atomics on one address are serialised in L2 at about
// 0.7 ns each, so one per warp made the |max| bookkeeping the longest part of the small element-wise kernels
// (18 k warps -> 13 us) - and of a 46 k-block experiment, 255 us.
__device__ __forceinline__ void warp_amax(float* slot, float v) {
  __shared__ float s_amax[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) s_amax[warp] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float m = threadIdx.x < nwarps ? s_amax[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0 && slot) atomic_amax(slot, m);
  }
  __syncthreads();          // the staging array may be reused by a second call
}

// ---- interpolant (interpolants.py:16-33, 53-108) and loss targets (losses.py:126-133) -----------------------------------------
enum { GAMMA_BROWNIAN = 0, GAMMA_SIN2 = 1 };

__device__ __forceinline__ void gamma_of(int kind, float a, float t, float& g, float& gd) {
  if (kind == GAMMA_SIN2) {
    float sn, cs;
    sincosf(kPiF * t, &sn, &cs);
    g = sn * sn;
    gd = 2.0f * kPiF * sn * cs;
  } else {
    const float r = sqrtf(a * t * (1.0f - t));
    g = r;
    gd = (1.0f / (2.0f * r)) * a * (1.0f - 2.0f * t);
  }
}

// xt[p][n] = (1 - t) x0 + t x1 +- gamma(t) z ;  tgt[p][n] = (x1 - x0) +- gamma_dot(t) z ; colsum[p][c] += xt  (p = 0: +, 1: -)
__global__ void k_tr_interp(int N, const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ t,
                            const float* __restrict__ z, int gamma_kind, float a, float* __restrict__ xt, float* __restrict__ tgt,
                            float* __restrict__ colsum) {
  pdl_entry();
  __shared__ float red[6];
  if (threadIdx.x < 6) red[threadIdx.x] = 0.0f;
  __syncthreads();
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float tt = t[n];
    float g, gd;
    gamma_of(gamma_kind, a, tt, g, gd);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a0 = x0[3 * n + c], a1 = x1[3 * n + c], zz = z[3 * n + c];
      const float it = (1.0f - tt) * a0 + tt * a1, gz = g * zz, dti = -1.0f * a0 + 1.0f * a1, gdz = gd * zz;
      const float xp = it + gz, xm = it - gz;
      xt[3 * n + c] = xp;
      xt[3 * (N + n) + c] = xm;
      tgt[3 * n + c] = dti + gdz;
      tgt[3 * (N + n) + c] = dti - gdz;
      acc[c] += xp;
      acc[3 + c] += xm;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float w = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], w);
  }
  __syncthreads();
  if (threadIdx.x < 6) atomicAdd(&colsum[threadIdx.x], red[threadIdx.x]);
}

// losses.py:56-57: xt -= mean over ALL atoms of the batch (per pass)
__global__ void k_tr_center(int N, float* __restrict__ xt, const float* __restrict__ colsum) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * N) return;
  const int p = i / (3 * N), c = i % 3;
  xt[i] = xt[i] - __fdiv_rn(colsum[3 * p + c], (float)N);
}

// ---- graph tables: one block per (molecule, pass) ----------------------------------------------------------------------------------
struct GraphP {
  int B, N;
  long long E;
  const int* mol_ptr;            // [B + 1]
  const long long* edge_ptr;     // [B + 1]
  const unsigned char* edge_type;// [E] in (src, dst) order (the batch contract)
  const float* xt;               // [2 N][3]
  int *src, *dst, *pair, *etype, *in_ptr;      // [E2], [E2], [E2], [E2], [N2 + 1]
  float4* dir;                   // [E2] (ex, ey, ez, dist): graph.py:27-29
  float* pair_dist;              // [P2]
};

__global__ void k_tr_graph(const GraphP g) {
  pdl_entry();
  const int m = blockIdx.x, pass = blockIdx.y;
  const int nb0 = g.mol_ptr[m], n = g.mol_ptr[m + 1] - nb0;
  const long long eb0 = g.edge_ptr[m];
  const int nbase = pass * g.N + nb0;
  const long long ebase = (long long)pass * g.E + eb0;
  const int ne = n * (n - 1);
  for (int k = threadIdx.x; k < ne; k += blockDim.x) {
    const int jl = k / (n - 1), kk = k - jl * (n - 1), il = kk + (kk >= jl ? 1 : 0);
    const long long row = ebase + k;
    const int s = nbase + il, d = nbase + jl;
    g.src[row] = s;
    g.dst[row] = d;
    g.etype[row] = g.edge_type[eb0 + (long long)il * (n - 1) + jl - (jl > il ? 1 : 0)];
    const int a = min(il, jl), b = max(il, jl);
    const long long pr = ebase / 2 + (long long)a * (n - 1) - (long long)a * (a - 1) / 2 + (b - a - 1);
    g.pair[row] = (int)pr;
    const float rx = g.xt[3 * s] - g.xt[3 * d], ry = g.xt[3 * s + 1] - g.xt[3 * d + 1], rz = g.xt[3 * s + 2] - g.xt[3 * d + 2];
    const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    const float den = 1.0f + dist;
    g.dir[row] = make_float4(__fdiv_rn(rx, den), __fdiv_rn(ry, den), __fdiv_rn(rz, den), dist);
    if (il < jl) g.pair_dist[pr] = dist;
  }
  for (int jl = threadIdx.x; jl < n; jl += blockDim.x) g.in_ptr[nbase + jl] = (int)(ebase + (long long)jl * (n - 1));
  if (m == g.B - 1 && pass == 1 && threadIdx.x == 0) g.in_ptr[2 * g.N] = (int)(2 * g.E);
}

// ---- embedding inputs (embedding.py:68-86, 127-160, 200-212) --------------------------------------------------------------------------
// X0[n] = cat[atom_emb[atoms[n]], PE((T0 - mean) / range; temp_length), PE(T1 ...), PE(t[n]; time_length)]   [N][(2 + n_temp) F]
__global__ void k_tr_embed_in(int N, int F, int n_temp, const int* __restrict__ atoms, const float* __restrict__ T0,
                              const float* __restrict__ T1, const float* __restrict__ t, const float* __restrict__ atom_emb,
                              float temp_mean, float temp_range, float temp_length, float time_length, float* __restrict__ X0) {
  pdl_entry();
  const int n = blockIdx.x;
  const int width = (2 + n_temp) * F;
  float* row = X0 + (long long)n * width;
  const float* emb = atom_emb + (long long)atoms[n] * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) row[f] = emb[f];
  for (int seg = 0; seg <= n_temp; ++seg) {
    float val, len;
    if (seg < n_temp) {
      const float T = seg == 0 ? T0[n] : T1[n];
      val = __fdiv_rn(__fsub_rn(T, __fmul_rn(temp_mean, 1.0f)), temp_range);
      len = temp_length;
    } else {
      val = t[n];
      len = time_length;
    }
    for (int r = threadIdx.x; r < F / 2; r += blockDim.x) {
      float sn, cs;
      sincosf(pe_arg(val, len, r + 1), &sn, &cs);
      row[(1 + seg) * F + 2 * r] = cs;
      row[(1 + seg) * F + 2 * r + 1] = sn;
    }
  }
}

// PE of the pair distances [P2][F] (embedding.py:137-160, cpainn.py:282) and e0 = Emb4(edge_type) [E2][F] (cpainn.py:70)
__global__ void k_tr_pair_pe(long long P2, int F, const float* __restrict__ pair_dist, float length, float* __restrict__ out) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = F / 2;
  if (i >= P2 * half) return;
  const long long p = i / half;
  const int r = (int)(i - p * half);
  float sn, cs;
  sincosf(pe_arg(pair_dist[p], length, r + 1), &sn, &cs);
  out[p * F + 2 * r] = cs;
  out[p * F + 2 * r + 1] = sn;
}
__global__ void k_tr_gather_rows(long long rows, int F, const int* __restrict__ idx, int idx_mod, const float* __restrict__ table,
                                 float* __restrict__ out) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const long long r = i / F;
  const int f = (int)(i - r * F);
  const long long sr = idx ? (long long)idx[r] : (idx_mod ? r % idx_mod : r);
  out[i] = table[sr * F + f];
}
// grad_table[idx[r]] += d[r]  for a table with few rows (edge / atom embeddings): block-local sums in shared memory first
// (shared-memory atomics: several rows of the block may hit the same table row), then one global atomic per entry and block
__global__ void k_tr_scatter_rows(long long rows, int F, int table_rows, const int* __restrict__ idx, const float* __restrict__ d,
                                  float* __restrict__ grad_table) {
  pdl_entry();
  extern __shared__ float acc[];                      // [table_rows][F]
  for (int i = threadIdx.x; i < table_rows * F; i += blockDim.x) acc[i] = 0.0f;
  __syncthreads();
  const int lanes_per_row = F < (int)blockDim.x ? F : (int)blockDim.x, rows_per_it = blockDim.x / lanes_per_row;
  const int f0 = threadIdx.x % lanes_per_row, rsub = threadIdx.x / lanes_per_row;
  const long long per = (rows + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  if (rsub < rows_per_it)
    for (long long r = r0 + rsub; r < r1; r += rows_per_it) {
      const int tr = idx[r];
      for (int f = f0; f < F; f += lanes_per_row) atomicAdd(&acc[tr * F + f], d[r * F + f]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < table_rows * F; i += blockDim.x)
    if (acc[i] != 0.0f) atomicAdd(&grad_table[i], acc[i]);
}
// out[n] = a[n] + a[N + n]  (the two passes share the x-independent embedding)
__global__ void k_tr_fold_passes(long long n, const float* __restrict__ a, float* __restrict__ out, float* amax) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.0f;
  if (i < n) { v = a[i] + a[n + i]; out[i] = v; }
  warp_amax(amax, fabsf(v));
}

// ---- Linear -> LayerNorm -> SiLU (embedding.py:26-34), one warp per row ---------------------------------------------------------------
// z [R][F] (pre-LayerNorm, bias included) -> n = (z - mean) / sqrt(var + eps) in place, rstd [R], h = SiLU(gamma n + beta)
__global__ void k_tr_ln_silu_fwd(long long R, int F, float* __restrict__ zn, float* __restrict__ rstd, float* __restrict__ h,
                                 const float* __restrict__ gamma, const float* __restrict__ beta) {
  pdl_entry();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * wpb) {
    float* z = zn + r * F;
    float sum = 0.0f;
    for (int f = lane; f < F; f += 32) sum += z[f];
    const float mean = warp_sum(sum) / (float)F;
    float var = 0.0f;
    for (int f = lane; f < F; f += 32) { const float d = z[f] - mean; var += d * d; }
    const float rs = 1.0f / sqrtf(warp_sum(var) / (float)F + 1e-5f);
    if (lane == 0) rstd[r] = rs;
    for (int f = lane; f < F; f += 32) {
      const float nn = (z[f] - mean) * rs;
      z[f] = nn;
      h[r * F + f] = silu_mufu(fmaf(nn, gamma[f], beta[f]));
    }
  }
}

// dh [R][F] (gradient wrt h) -> dz in place (gradient wrt the Linear output); accumulates d gamma, d beta, d bias
// (= column sums of dz) with one atomic per column and block; records |max| of dz.
template <int KM>                                       // columns per lane: F <= 32 KM
__global__ void __launch_bounds__(256, 4) k_tr_ln_silu_bwd(long long R, int F, float* __restrict__ dh, const float* __restrict__ nrm,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ g_gamma,
                                                           float* __restrict__ g_beta, float* __restrict__ g_bias, float* amax) {
  pdl_entry();
  extern __shared__ float sacc[];                       // [3][F]
  for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) sacc[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float a_g[KM], a_b[KM], a_z[KM], gam[KM], bet[KM];
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    const int f = lane + 32 * k;
    a_g[k] = a_b[k] = a_z[k] = 0.0f;
    gam[k] = f < F ? gamma[f] : 0.0f;
    bet[k] = f < F ? beta[f] : 0.0f;
  }
  float mx = 0.0f;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * wpb) {
    float* d = dh + r * F;
    const float* nn = nrm + r * F;
    const float rs = rstd[r];
    float dn[KM], nv[KM];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const int f = lane + 32 * k;
      dn[k] = nv[k] = 0.0f;
      if (f < F) {
        const float n_ = nn[f], u = fmaf(n_, gam[k], bet[k]);
        const float sg = sigmoid_fast(u);
        const float dpre = d[f] * (sg * (1.0f + u * (1.0f - sg)));       // SiLU'(u)
        a_g[k] += dpre * n_;
        a_b[k] += dpre;
        nv[k] = n_;
        dn[k] = dpre * gam[k];
        s1 += dn[k];
        s2 += dn[k] * n_;
      }
    }
    s1 = warp_sum(s1) / (float)F;
    s2 = warp_sum(s2) / (float)F;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const int f = lane + 32 * k;
      if (f < F) {
        const float dz = (dn[k] - s1 - nv[k] * s2) * rs;
        d[f] = dz;
        a_z[k] += dz;
        mx = fmaxf(mx, fabsf(dz));
      }
    }
  }
  warp_amax(amax, mx);
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    const int f = lane + 32 * k;
    if (f < F) { atomicAdd(&sacc[f], a_g[k]); atomicAdd(&sacc[F + f], a_b[k]); atomicAdd(&sacc[2 * F + f], a_z[k]); }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    atomicAdd(&g_gamma[f], sacc[f]);
    atomicAdd(&g_beta[f], sacc[F + f]);
    atomicAdd(&g_bias[f], sacc[2 * F + f]);
  }
}

// LayerNorm parameter gradients from the by-products of the fused GEMM epilogue (train_gemm.cuh, EPI_LN_BWD):
//   g_gamma[c] += sum_r dpre[r][c] n[r][c],  g_beta[c] += sum_r dpre[r][c],  g_bias[c] += sum_r dz[r][c]
// grid = (column blocks, row slabs); runs on the side stream
__global__ void k_tr_ln_param_grads(long long R, int C, const float* __restrict__ dpre, const float* __restrict__ nrm,
                                    const float* __restrict__ dz, float* __restrict__ g_gamma, float* __restrict__ g_beta,
                                    float* __restrict__ g_bias) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long per = (R + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(R, r0 + per);
  float ag[2] = {0.0f, 0.0f}, ab[2] = {0.0f, 0.0f}, az[2] = {0.0f, 0.0f};
  long long r = r0;
  for (; r + 2 <= r1; r += 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float dp = __ldg(dpre + (r + u) * C + c);
      ag[u] += dp * __ldg(nrm + (r + u) * C + c);
      ab[u] += dp;
      az[u] += __ldg(dz + (r + u) * C + c);
    }
  }
  for (; r < r1; ++r) {
    const float dp = __ldg(dpre + r * C + c);
    ag[0] += dp * __ldg(nrm + r * C + c);
    ab[0] += dp;
    az[0] += __ldg(dz + r * C + c);
  }
  atomicAdd(&g_gamma[c], ag[0] + ag[1]);
  atomicAdd(&g_beta[c], ab[0] + ab[1]);
  atomicAdd(&g_bias[c], az[0] + az[1]);
}

// out[c] += sum_r d[r][c]   (bias gradient of an output Linear); also records |max| of d.  grid = (column blocks, row slabs)
__global__ void k_tr_colsum(long long R, int C, const float* __restrict__ d, float* __restrict__ out, float* amax) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (R + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(R, r0 + per);
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, mx = 0.0f;
  if (c < C) {
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {                        // four independent loads in flight per thread
#pragma unroll
      for (int u = 0; u < 4; ++u) { const float v = __ldg(d + (r + u) * C + c); acc[u] += v; mx = fmaxf(mx, fabsf(v)); }
    }
    for (; r < r1; ++r) { const float v = __ldg(d + r * C + c); acc[0] += v; mx = fmaxf(mx, fabsf(v)); }
    const float tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    if (tot != 0.0f) atomicAdd(&out[c], tot);
  }
  warp_amax(amax, mx);
}

// ---- SE3Message (cpainn.py:263-310): m = phi3 * w3, gated scatter over the incoming edges; one block per destination node ------
struct CombineP {
  int N2, F;
  const int *in_ptr, *src, *pair;
  const float4* dir;
  const float *phi3, *w3;                    // [E2][5F], [P2][5F]
  const float *s_in, *v_in, *e_in;           // layer inputs
  float *s_out, *v_out, *e_out;
};

__global__ void k_tr_combine_fwd(const CombineP p) {
  pdl_entry();
  const int F = p.F;
  for (int j = blockIdx.x; j < p.N2; j += gridDim.x) {
    const int r0 = p.in_ptr[j], r1 = p.in_ptr[j + 1];
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const float vjx = p.v_in[((long long)j * 3 + 0) * F + f], vjy = p.v_in[((long long)j * 3 + 1) * F + f],
                  vjz = p.v_in[((long long)j * 3 + 2) * F + f];
      float ds = 0.0f, dvx = 0.0f, dvy = 0.0f, dvz = 0.0f;
      for (int r = r0; r < r1; ++r) {
        const float* ph = p.phi3 + (long long)r * 5 * F + f;
        const float* ww = p.w3 + (long long)p.pair[r] * 5 * F + f;
        const float mg = ph[0] * ww[0], msd = ph[F] * ww[F], mds = ph[2 * F] * ww[2 * F], mde = ph[3 * F] * ww[3 * F],
                    mcg = ph[4 * F] * ww[4 * F];
        const float4 d = p.dir[r];
        const long long i = p.src[r];
        const float vix = p.v_in[(i * 3 + 0) * F + f], viy = p.v_in[(i * 3 + 1) * F + f], viz = p.v_in[(i * 3 + 2) * F + f];
        // cross = dir x v[dst]  (cpainn.py:296-298)
        const float cx = d.y * vjz - d.z * vjy, cy = d.z * vjx - d.x * vjz, cz = d.x * vjy - d.y * vjx;
        dvx += (msd * d.x + mg * vix) + mcg * cx;
        dvy += (msd * d.y + mg * viy) + mcg * cy;
        dvz += (msd * d.z + mg * viz) + mcg * cz;
        ds += mds;
        p.e_out[(long long)r * F + f] = p.e_in[(long long)r * F + f] + mde;
      }
      p.s_out[(long long)j * F + f] = p.s_in[(long long)j * F + f] + ds;
      p.v_out[((long long)j * 3 + 0) * F + f] = vjx + dvx;
      p.v_out[((long long)j * 3 + 1) * F + f] = vjy + dvy;
      p.v_out[((long long)j * 3 + 2) * F + f] = vjz + dvz;
    }
  }
}

// Adjoint of the above.  In: ds, dv = gradients wrt (s_out, v_out) [running buffers, updated in place to the gradients
// wrt (s_in, v_in) minus the parts that arrive later through the phi MLP], de = gradient wrt e_out (left as is: e_in's
// identity part; the phi MLP's input gradient is accumulated onto it afterwards).  Out: d_phi3 [E2][5F], d_w3 [P2][5F]
// (atomic: two directed edges per pair; zeroed by the caller), dv_src [N2][3][F] (atomic scatter to the source nodes; zeroed by
// the caller and added to dv by k_tr_add afterwards).
struct CombineBwdP {
  CombineP c;
  float *ds, *dv;
  const float* de;
  float *d_phi3, *d_w3, *dv_src;
  float *amax_phi3, *amax_w3;
};

__global__ void k_tr_combine_bwd(const CombineBwdP q) {
  pdl_entry();
  const CombineP& p = q.c;
  const int F = p.F;
  float mx_p = 0.0f, mx_w = 0.0f;
  for (int j = blockIdx.x; j < p.N2; j += gridDim.x) {
    const int r0 = p.in_ptr[j], r1 = p.in_ptr[j + 1];
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const long long jb = (long long)j * 3 * F + f;
      const float vjx = p.v_in[jb], vjy = p.v_in[jb + F], vjz = p.v_in[jb + 2 * F];
      const float gx = q.dv[jb], gy = q.dv[jb + F], gz = q.dv[jb + 2 * F];     // d loss / d v_out[j]
      const float gs = q.ds[(long long)j * F + f];
      float ax = 0.0f, ay = 0.0f, az = 0.0f;                                    // extra gradient wrt v_in[j] through the cross term
      for (int r = r0; r < r1; ++r) {
        const long long pr = p.pair[r];
        const float* ph = p.phi3 + (long long)r * 5 * F + f;
        const float* ww = p.w3 + pr * 5 * F + f;
        const float4 d = p.dir[r];
        const long long i = p.src[r];
        const long long ib = i * 3 * F + f;
        const float vix = p.v_in[ib], viy = p.v_in[ib + F], viz = p.v_in[ib + 2 * F];
        const float cx = d.y * vjz - d.z * vjy, cy = d.z * vjx - d.x * vjz, cz = d.x * vjy - d.y * vjx;
        float dm[5];
        dm[0] = gx * vix + gy * viy + gz * viz;            // gates
        dm[1] = gx * d.x + gy * d.y + gz * d.z;            // scale_edge_dir
        dm[2] = gs;                                        // ds
        dm[3] = q.de[(long long)r * F + f];                // de
        dm[4] = gx * cx + gy * cy + gz * cz;               // cross_gates
        const float mg = ph[0] * ww[0], mcg = ph[4 * F] * ww[4 * F];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float dp = dm[k] * ww[k * F], dw = dm[k] * ph[k * F];
          q.d_phi3[(long long)r * 5 * F + k * F + f] = dp;
          atomicAdd(&q.d_w3[pr * 5 * F + k * F + f], dw);
          mx_p = fmaxf(mx_p, fabsf(dp));
          mx_w = fmaxf(mx_w, fabsf(dw));
        }
        atomicAdd(&q.dv_src[ib], mg * gx);
        atomicAdd(&q.dv_src[ib + F], mg * gy);
        atomicAdd(&q.dv_src[ib + 2 * F], mg * gz);
        // (dir x v) . g = v . (g x dir)
        ax += mcg * (gy * d.z - gz * d.y);
        ay += mcg * (gz * d.x - gx * d.z);
        az += mcg * (gx * d.y - gy * d.x);
      }
      q.dv[jb] = gx + ax;
      q.dv[jb + F] = gy + ay;
      q.dv[jb + 2 * F] = gz + az;
    }
  }
  warp_amax(q.amax_phi3, mx_p);
  warp_amax(q.amax_w3, 2.0f * mx_w);       // two addends per pair: an upper bound of |d_w3|
}

__global__ void k_tr_add(long long n, float* __restrict__ a, const float* __restrict__ b) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// ---- Update (cpainn.py:345-376) ---------------------------------------------------------------------------------------------------
// uvvv [3 N2][2F] = v [U; V]^T : columns [0, F) = U v, [F, 2F) = V v.   q = |V v| over xyz.
__global__ void k_tr_upd_q(long long NF, int F, const float* __restrict__ uvvv, float* __restrict__ qv) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NF) return;
  const long long n = i / F;
  const int f = (int)(i - n * F);
  const float* b = uvvv + n * 3 * 2 * F + F + f;
  const float x = b[0], y = b[2 * F], z = b[4 * F];
  qv[i] = sqrtf(x * x + y * y + z * z);
}
// v_out = v + (U v) g ; s_out = s + (q^2 a + c)        gac [N2][3F] = (gates, scale_sq, add_s)
__global__ void k_tr_upd_apply(long long NF, int F, const float* __restrict__ uvvv, const float* __restrict__ qv,
                               const float* __restrict__ gac, const float* __restrict__ s, const float* __restrict__ v,
                               float* __restrict__ s_out, float* __restrict__ v_out) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NF) return;
  const long long n = i / F;
  const int f = (int)(i - n * F);
  const float g = gac[n * 3 * F + f], a = gac[n * 3 * F + F + f], c = gac[n * 3 * F + 2 * F + f], qq = qv[i];
  s_out[i] = s[i] + ((qq * qq) * a + c);
#pragma unroll
  for (int k = 0; k < 3; ++k) v_out[(n * 3 + k) * F + f] = v[(n * 3 + k) * F + f] + uvvv[(n * 3 + k) * 2 * F + f] * g;
}
// gradients wrt (s_out, v_out) in ds, dv (left in place: identity parts) -> d_gac, d_uvvv[:, 0:F] (U v part), dq (direct part)
__global__ void k_tr_upd_bwd1(long long NF, int F, const float* __restrict__ uvvv, const float* __restrict__ qv,
                              const float* __restrict__ gac, const float* __restrict__ ds, const float* __restrict__ dv,
                              float* __restrict__ d_gac, float* __restrict__ d_uvvv, float* __restrict__ dq, float* amax_gac) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float mx = 0.0f;
  if (i < NF) {
    const long long n = i / F;
    const int f = (int)(i - n * F);
    const float g = gac[n * 3 * F + f], a = gac[n * 3 * F + F + f], qq = qv[i], gs = ds[i];
    float dg = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float gv = dv[(n * 3 + k) * F + f];
      dg += gv * uvvv[(n * 3 + k) * 2 * F + f];
      d_uvvv[(n * 3 + k) * 2 * F + f] = gv * g;
    }
    const float da = gs * (qq * qq);
    d_gac[n * 3 * F + f] = dg;
    d_gac[n * 3 * F + F + f] = da;
    d_gac[n * 3 * F + 2 * F + f] = gs;
    dq[i] = gs * (2.0f * qq * a);
    mx = fmaxf(fabsf(dg), fmaxf(fabsf(da), fabsf(gs)));
  }
  warp_amax(amax_gac, mx);
}
// d (V v)[c] = dq (V v)[c] / q   (0 where q = 0, as torch's norm backward)   -> d_uvvv[:, F:2F]; |max| of all of d_uvvv
__global__ void k_tr_upd_bwd2(long long NF, int F, const float* __restrict__ uvvv, const float* __restrict__ qv,
                              const float* __restrict__ dq, float* __restrict__ d_uvvv, float* amax) {
  pdl_entry();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float mx = 0.0f;
  if (i < NF) {
    const long long n = i / F;
    const int f = (int)(i - n * F);
    const float qq = qv[i], w = qq > 0.0f ? dq[i] / qq : 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float dvv = w * uvvv[(n * 3 + k) * 2 * F + F + f];
      d_uvvv[(n * 3 + k) * 2 * F + F + f] = dvv;
      mx = fmaxf(mx, fmaxf(fabsf(dvv), fabsf(d_uvvv[(n * 3 + k) * 2 * F + f])));
    }
  }
  warp_amax(amax, mx);
}

// ---- LayerReadout with one output feature (cpainn.py:425-437) + loss (losses.py:126-133); one warp per node --------------------
// y = W3 h2 + b3 (2 values: s_out, gate), ev[c] = Vout . v[n][c], out[n][c] = ev[c] * gate
struct ReadoutP {
  int N2, F, N;                  // N = atoms per pass (the loss is a mean over N)
  const float *h2, *v, *W3, *b3, *Vout;      // W3 [2][F]
  const float* tgt;              // [N2][3]
  float* out;                    // [N2][3] the drift b
  float* gate;                   // [N2]
  double* loss;                  // += (0.5 |b|^2 - tgt . b) / N
  // backward outputs
  float *dh2, *dv;               // [N2][F], [N2][3][F]
  float *g_W3, *g_b3, *g_Vout;
  float* amax_dh2;
};

__global__ void k_tr_readout(const ReadoutP p) {
  pdl_entry();
  extern __shared__ float sacc[];                    // [2][F]: g_W3 row 1, g_Vout
  const int F = p.F, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) sacc[i] = 0.0f;
  __syncthreads();
  double lsum = 0.0;
  float gb = 0.0f, mx = 0.0f;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < p.N2; n += gridDim.x * wpb) {
    float y1 = 0.0f, ev[3] = {0.0f, 0.0f, 0.0f};
    for (int f = lane; f < F; f += 32) {
      y1 = fmaf(p.W3[F + f], p.h2[(long long)n * F + f], y1);
      const float vo = p.Vout[f];
#pragma unroll
      for (int c = 0; c < 3; ++c) ev[c] = fmaf(vo, p.v[((long long)n * 3 + c) * F + f], ev[c]);
    }
    y1 = warp_sum(y1) + p.b3[1];
    float dgate = 0.0f, dout[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ev[c] = warp_sum(ev[c]);
      const float b = ev[c] * y1, tg = p.tgt[3 * n + c];
      if (lane == 0) {
        p.out[3 * n + c] = b;
        lsum += (0.5 * (double)b * b - (double)tg * b);
      }
      dout[c] = (b - tg) / (float)p.N;               // d loss / d b
      dgate += dout[c] * ev[c];
    }
    if (lane == 0) { p.gate[n] = y1; gb += dgate; }
    for (int f = lane; f < F; f += 32) {
      const float d = dgate * p.W3[F + f];
      p.dh2[(long long)n * F + f] = d;
      mx = fmaxf(mx, fabsf(d));
      atomicAdd(&sacc[f], dgate * p.h2[(long long)n * F + f]);
      const float vo = p.Vout[f];
      float gv = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float w = dout[c] * y1;
        p.dv[((long long)n * 3 + c) * F + f] = w * vo;
        gv += w * p.v[((long long)n * 3 + c) * F + f];
      }
      atomicAdd(&sacc[F + f], gv);
    }
  }
  warp_amax(p.amax_dh2, mx);
  if (lane == 0) {
    if (lsum != 0.0) atomicAdd(p.loss, lsum / (double)p.N);
    if (gb != 0.0f) atomicAdd(&p.g_b3[1], gb);
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    atomicAdd(&p.g_W3[F + f], sacc[f]);
    atomicAdd(&p.g_Vout[f], sacc[F + f]);
  }
}

// ---- optimiser: clip_grad_norm_(params, max_norm) + torch.optim.Adam (train_ambient.py:96,146-148) ------------------------------
__global__ void k_tr_sqnorm(long long n, const float* __restrict__ g, double* __restrict__ out) {
  pdl_entry();
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = g[i];
    acc += v * v;
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}
// sqnorm = sum g^2 BEFORE clipping (device); clip coefficient min(1, max_norm / (norm + 1e-6)) (max_norm <= 0: no clipping)
__global__ void k_tr_adam(long long n, float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                          const double* __restrict__ sqnorm, float max_norm, float lr, float beta1, float beta2, float eps,
                          float weight_decay, float bc1, float bc2_sqrt) {
  pdl_entry();
  float coef = 1.0f;
  if (max_norm > 0.0f) {
    const float total = (float)sqrt(*sqnorm);
    coef = fminf(1.0f, max_norm / (total + 1e-6f));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * coef;
    if (weight_decay != 0.0f) gi = gi + weight_decay * w[i];
    const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);            // lerp_
    const float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    w[i] = w[i] + (-lr / bc1) * (mi / denom);
  }
}

}  // namespace train
}  // namespace tib
