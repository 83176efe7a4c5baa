// simt_tangent.cuh - forward-mode tangents of the cPaiNN drift in fp32 on CUDA cores: the exact
// divergence div_x b = sum_{atom a, coordinate c} d b[a][c] / d x[a][c] per molecule that the reference
// obtains from 3n reverse passes (ODEWrapper.compute_divergence, ambient ode_wrapper.py:59-91, latent
// :57-86).
//
// One pass propagates D tangent directions next to the primal values; direction q = dir0 + dd seeds
// x_dot = unit vector on coordinate q % 3 of atom q / 3 of EVERY molecule (molecules are independent, so
// one pass serves all of them; a molecule with fewer atoms simply carries a zero tangent).  The embedded
// node and edge features do not depend on x, so all tangents start at zero and enter through the edge
// geometry of each message layer.  Every kernel below is the dual-number version of its namesake in
// simt_drift.cuh: the primal arithmetic is op-for-op the same (b from this path equals TIB_MATH_FP32_SIMT
// up to the grouping of the per-node sums, because the edge tiles are smaller), the tangent arithmetic is
// plain fp32.
//
// Tangent state (HBM, caller's workspace), D copies each at a fixed stride:
//   ts [D][N][F], tv [D][N][3][F], te [D][E][F], tout [D][N][3]
#pragma once
#include "simt_drift.cuh"

namespace tib {

// y * sigmoid(y) exactly as silu(), and its derivative sigmoid(y) * (1 + y * (1 - sigmoid(y)))
__device__ __forceinline__ void silu_dual(float y, float& out, float& dout) {
  const float den = 1.0f + expf(-y);
  out = __fdiv_rn(y, den);
  const float sg = __fdiv_rn(1.0f, den);
  dout = sg * (1.0f + y * (1.0f - sg));
}

// acc[i][q][c] = sum_k A_i[(q*8+warp)*lda + k] * Wt[k*ldw + col + c] for the primal input A_0 = A and the tangent
// inputs A_i = At + (i-1)*at_stride: one pass over the weight rows serves all NI inputs (the same fmaf chain
// per accumulator as gemm_rows, so acc[0] is bit-identical to it).
template <int RPT, int CPL, int NI>
__device__ __forceinline__ void gemm_rows_multi(float (&acc)[NI][RPT][CPL], const float* A, const float* At, int at_stride,
                                                int lda, int K, const float* __restrict__ Wt, int ldw, int col, int warp) {
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[i][q][c] = 0.0f;
  const float* wp = Wt + col;
  const float* ap[NI];
  ap[0] = A + warp * lda;
#pragma unroll
  for (int i = 1; i < NI; ++i) ap[i] = At + (size_t)(i - 1) * at_stride + warp * lda;
  float w[4][CPL], wn[4][CPL];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldg_vec<CPL>(w[kk], wp + (size_t)kk * ldw);
  for (int k0 = 0; k0 < K; k0 += 4) {
    if (k0 + 4 < K) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ldg_vec<CPL>(wn[kk], wp + (size_t)(k0 + 4 + kk) * ldw);
    }
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(ap[i] + q * 8 * lda + k0);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float t = acc[i][q][c];
          t = fmaf(a.x, w[0][c], t);
          t = fmaf(a.y, w[1][c], t);
          t = fmaf(a.z, w[2][c], t);
          t = fmaf(a.w, w[3][c], t);
          acc[i][q][c] = t;
        }
      }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int c = 0; c < CPL; ++c) w[kk][c] = wn[kk][c];
  }
}

// Hidden layer with tangents: Xout = SiLU(LN(A @ Wt + b)), Xt[dd] = J_{SiLU o LN}(A @ Wt + b) (At[dd] @ Wt).
// With n = (z - mean) * rstd:  d n = rstd * (dz - mean(dz) - n * mean(n * (dz - mean(dz)))).
template <int F, int RPT, int D>
__device__ __forceinline__ void layer_ln_silu_jvp(const float* A, const float* At, int at_stride, int lda, int K,
                                                  const float* __restrict__ Wt, const float* __restrict__ b,
                                                  const float* __restrict__ g, const float* __restrict__ be,
                                                  float* Xout, float* Xt, int xt_stride, int ldo, int warp, int lane) {
  using C = Cols<F>;
  float acc[C::NCH][1 + D][RPT][C::CPL];   // [.][0] primal (becomes the normalised value n), [.][1+dd] tangents
  float coef[C::NCH][RPT][C::CPL];
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch)
    gemm_rows_multi<RPT, C::CPL, 1 + D>(acc[ch], A, At, at_stride, lda, K, Wt, F, ch * C::CW + lane * C::CPL, warp);
  {
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
          acc[ch][0][q][c] += bb[ch][c];
          sum += acc[ch][0][q][c];
        }
      const float mean = warp_sum(sum) * (1.0f / F);
      float ss = 0.0f;
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch)
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) {
          const float d = acc[ch][0][q][c] - mean;
          acc[ch][0][q][c] = d;
          ss = fmaf(d, d, ss);
        }
      const float var = warp_sum(ss) * (1.0f / F);
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(var + 1e-5f));
      float* orow = Xout + (size_t)(q * 8 + warp) * ldo;
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) {
        float y[C::CPL];
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) {
          const float n = __fmul_rn(acc[ch][0][q][c], rstd);
          float ds;
          silu_dual(__fadd_rn(__fmul_rn(n, gg[ch][c]), ee[ch][c]), y[c], ds);
          acc[ch][0][q][c] = n;
          coef[ch][q][c] = ds * gg[ch][c] * rstd;
        }
        st_vec<C::CPL>(orow + ch * C::CW + lane * C::CPL, y);
      }
    }
  }
#pragma unroll
  for (int dd = 0; dd < D; ++dd) {
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      float sum = 0.0f;
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch)
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) sum += acc[ch][1 + dd][q][c];
      const float mean = warp_sum(sum) * (1.0f / F);
      float cn = 0.0f;
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch)
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) {
          acc[ch][1 + dd][q][c] -= mean;
          cn = fmaf(acc[ch][0][q][c], acc[ch][1 + dd][q][c], cn);
        }
      cn = warp_sum(cn) * (1.0f / F);
      float* orow = Xt + (size_t)dd * xt_stride + (size_t)(q * 8 + warp) * ldo;
#pragma unroll
      for (int ch = 0; ch < C::NCH; ++ch) {
        float y[C::CPL];
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) y[c] = coef[ch][q][c] * (acc[ch][1 + dd][q][c] - acc[ch][0][q][c] * cn);
        st_vec<C::CPL>(orow + ch * C::CW + lane * C::CPL, y);
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// message with tangents: SE3Message.forward and its directional derivative   (cpainn.py:263-310)
// ---------------------------------------------------------------------------------------------
struct MessageJvpP {
  MessageP m;              // primal arguments, exactly as k_message
  const float* ts_old;     // [D][N][F]
  const float* tv_old;     // [D][N][3][F]
  float* ts_new;
  float* tv_new;
  float* te;               // [D][E][F] updated in place
  size_t st_s, st_v, st_e; // strides between directions (floats)
  int dir0;                // first direction of this pass
};

template <int F, int RPT, int D>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_message_jvp(MessageJvpP pp) {
  using C = Cols<F>;
  constexpr int TR = 8 * RPT;
  constexpr int CW = C::CW;
  const MessageP& p = pp.m;
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;                       // [TR][2F]
  float* XA = X0 + TR * 2 * F;            // [TR][F]
  float* HP = XA + TR * F;                // [TR][F]
  float* HW = HP + TR * F;                // [TR][F]
  float* GEO = HW + TR * F;               // [TR][4]   d, dir.xyz
  float* X0t = GEO + TR * 4;              // [D][TR][2F]
  float* XAt = X0t + D * TR * 2 * F;      // [D][TR][F]
  float* HPt = XAt + D * TR * F;          // [D][TR][F]
  float* HWt = HPt + D * TR * F;          // [D][TR][F]
  float* GEOt = HWt + D * TR * F;         // [D][TR][4]  d_dot, dir_dot.xyz
  float* XS = GEOt + D * TR * 4;          // [TIB_MAX_ATOMS][3]
  constexpr int SX0 = TR * 2 * F, SXF = TR * F, SG = TR * 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mol = blockIdx.x;
  const int n0 = __ldg(p.b.mol_ptr + mol);
  const int n = __ldg(p.b.mol_ptr + mol + 1) - n0;
  const int ne = n * (n - 1);
  const long long e0 = __ldg(p.b.edge_ptr + mol);
  const int nm1 = n - 1;
  const bool first = p.first_layer != 0;

  for (int idx = tid; idx < n * F; idx += TIB_THREADS) {
    p.s_new[(size_t)n0 * F + idx] = p.s_old[(size_t)n0 * F + idx];
#pragma unroll
    for (int dd = 0; dd < D; ++dd)
      pp.ts_new[dd * pp.st_s + (size_t)n0 * F + idx] = first ? 0.0f : pp.ts_old[dd * pp.st_s + (size_t)n0 * F + idx];
  }
  for (int idx = tid; idx < n * 3 * F; idx += TIB_THREADS) {
    p.v_new[(size_t)n0 * 3 * F + idx] = first ? 0.0f : p.v_old[(size_t)n0 * 3 * F + idx];
#pragma unroll
    for (int dd = 0; dd < D; ++dd)
      pp.tv_new[dd * pp.st_v + (size_t)n0 * 3 * F + idx] = first ? 0.0f : pp.tv_old[dd * pp.st_v + (size_t)n0 * 3 * F + idx];
  }
  for (int idx = tid; idx < n * 3; idx += TIB_THREADS) XS[idx] = p.x[(size_t)n0 * 3 + idx];
  __syncthreads();

  for (int r0 = 0; r0 < ne; r0 += TR) {
    const int rows = min(TR, ne - r0);
    // ---- geometry and its tangents: r = x[src]-x[dst], d = |r|, dir = r/(1+d)          (graph.py:27-29)
    if (tid < TR) {
      float d = 0.f, dir[3] = {0.f, 0.f, 0.f}, r[3] = {0.f, 0.f, 0.f};
      int i = -1, j = -1;
      if (tid < rows) {
        const int rr = r0 + tid;
        i = rr / nm1;
        const int k = rr % nm1;
        j = k + (k >= i);
        r[0] = XS[i * 3 + 0] - XS[j * 3 + 0];
        r[1] = XS[i * 3 + 1] - XS[j * 3 + 1];
        r[2] = XS[i * 3 + 2] - XS[j * 3 + 2];
        d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2])));
        const float den = 1.0f + d;
        dir[0] = __fdiv_rn(r[0], den); dir[1] = __fdiv_rn(r[1], den); dir[2] = __fdiv_rn(r[2], den);
      }
      GEO[tid * 4 + 0] = d; GEO[tid * 4 + 1] = dir[0]; GEO[tid * 4 + 2] = dir[1]; GEO[tid * 4 + 3] = dir[2];
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        const int q = pp.dir0 + dd, a = q / 3, c = q % 3;
        const float sgn = (tid < rows) ? (float)((i == a) - (j == a)) : 0.0f;   // d r[c] / d x[a][c]
        float dt = 0.f, dirt[3] = {0.f, 0.f, 0.f};
        if (sgn != 0.0f) {
          const float den = 1.0f + d;
          dt = d > 0.0f ? sgn * r[c] / d : 0.0f;
          const float k2 = dt / (den * den);
#pragma unroll
          for (int k = 0; k < 3; ++k) dirt[k] = (k == c ? sgn / den : 0.0f) - r[k] * k2;
        }
        float* gt = GEOt + dd * SG + tid * 4;
        gt[0] = dt; gt[1] = dirt[0]; gt[2] = dirt[1]; gt[3] = dirt[2];
      }
    }
    // ---- phi input cat[s[src], e] and its tangents                                 (cpainn.py:275-281)
    for (int idx = tid; idx < TR * (F / 4); idx += TIB_THREADS) {
      const int row = idx / (F / 4), f4 = idx % (F / 4);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 sv = zero, ev = zero;
      const bool live = row < rows;
      const int i = live ? (r0 + row) / nm1 : 0;
      if (live) {
        sv = reinterpret_cast<const float4*>(p.s_old + (size_t)(n0 + i) * F)[f4];
        ev = reinterpret_cast<const float4*>(p.e + (size_t)(e0 + r0 + row) * F)[f4];
      }
      reinterpret_cast<float4*>(X0 + row * 2 * F)[f4] = sv;
      reinterpret_cast<float4*>(X0 + row * 2 * F + F)[f4] = ev;
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        float4 ts = zero, te = zero;
        if (live && !first) {
          ts = reinterpret_cast<const float4*>(pp.ts_old + dd * pp.st_s + (size_t)(n0 + i) * F)[f4];
          te = reinterpret_cast<const float4*>(pp.te + dd * pp.st_e + (size_t)(e0 + r0 + row) * F)[f4];
        }
        reinterpret_cast<float4*>(X0t + dd * SX0 + row * 2 * F)[f4] = ts;
        reinterpret_cast<float4*>(X0t + dd * SX0 + row * 2 * F + F)[f4] = te;
      }
    }
    __syncthreads();

    layer_ln_silu_jvp<F, RPT, D>(X0, X0t, SX0, 2 * F, 2 * F, p.phi.W1t, p.phi.b1, p.phi.g1, p.phi.be1, XA, XAt, SXF, F, warp, lane);
    layer_ln_silu_jvp<F, RPT, D>(XA, XAt, SXF, F, F, p.phi.W2t, p.phi.b2, p.phi.g2, p.phi.be2, HP, HPt, SXF, F, warp, lane);
    // ---- w input PositionalEncoder(d) over X0[:, 0:F]; tangent = d/dd (cos, sin)(arg) * d_dot   (cpainn.py:283)
    for (int idx = lane; idx < RPT * (F / 2); idx += 32) {
      const int q = idx / (F / 2), rank = 1 + idx % (F / 2);
      const int row = q * 8 + warp;
      float sn, cs;
      sincosf(pe_arg(GEO[row * 4], p.length_scale, rank), &sn, &cs);
      X0[row * 2 * F + 2 * (rank - 1)] = cs;
      X0[row * 2 * F + 2 * (rank - 1) + 1] = sn;
      const float k = (float)rank * kPiF / p.length_scale;
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        const float kd = k * GEOt[dd * SG + row * 4];
        X0t[dd * SX0 + row * 2 * F + 2 * (rank - 1)] = -sn * kd;
        X0t[dd * SX0 + row * 2 * F + 2 * (rank - 1) + 1] = cs * kd;
      }
    }
    __syncwarp();
    layer_ln_silu_jvp<F, RPT, D>(X0, X0t, SX0, 2 * F, F, p.w.W1t, p.w.b1, p.w.g1, p.w.be1, XA, XAt, SXF, F, warp, lane);
    layer_ln_silu_jvp<F, RPT, D>(XA, XAt, SXF, F, F, p.w.W2t, p.w.b2, p.w.g2, p.w.be2, HW, HWt, SXF, F, warp, lane);
    __syncthreads();   // X0 / X0t become the CTA-wide product buffers

    // ---- output layer chunk by chunk: m = phi * w, m_dot = phi_dot * w + phi * w_dot         (cpainn.py:285-290)
    float* MB = X0;    // [TR][CW]
    const int i_lo = r0 / nm1, i_hi = (r0 + rows - 1) / nm1;
    for (int sp = 0; sp < 5; ++sp) {
      if (first && (sp == 0 || sp == 4)) continue;   // multiply v == 0 and v_dot == 0
      for (int ch = 0; ch < C::NCH; ++ch) {
        const int c0 = sp * F + ch * CW;
        {
          float acc[1 + D][RPT][C::CPL], accw[1 + D][RPT][C::CPL];
          gemm_rows_multi<RPT, C::CPL, 1 + D>(acc, HP, HPt, SXF, F, F, p.phi.W3t, 5 * F, c0 + lane * C::CPL, warp);
          gemm_rows_multi<RPT, C::CPL, 1 + D>(accw, HW, HWt, SXF, F, F, p.w.W3t, 5 * F, c0 + lane * C::CPL, warp);
          float bp[C::CPL], bw[C::CPL];
          ldg_vec<C::CPL>(bp, p.phi.b3 + c0 + lane * C::CPL);
          ldg_vec<C::CPL>(bw, p.w.b3 + c0 + lane * C::CPL);
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            float m[C::CPL];
#pragma unroll
            for (int c = 0; c < C::CPL; ++c) {
              acc[0][q][c] += bp[c];
              accw[0][q][c] += bw[c];
              m[c] = __fmul_rn(acc[0][q][c], accw[0][q][c]);
            }
            st_vec<C::CPL>(MB + (q * 8 + warp) * CW + lane * C::CPL, m);
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
#pragma unroll
              for (int c = 0; c < C::CPL; ++c) m[c] = acc[1 + dd][q][c] * accw[0][q][c] + acc[0][q][c] * accw[1 + dd][q][c];
              st_vec<C::CPL>(X0t + dd * SX0 + (q * 8 + warp) * CW + lane * C::CPL, m);
            }
          }
        }
        __syncthreads();
        const int fbase = ch * CW;
        if (sp == 3) {
          // e += de, e_dot += de_dot                                                 (cpainn.py:308)
          for (int idx = tid; idx < rows * (CW / 4); idx += TIB_THREADS) {
            const int row = idx / (CW / 4), f4 = idx % (CW / 4);
            float4* ep = reinterpret_cast<float4*>(p.e + (size_t)(e0 + r0 + row) * F + fbase) + f4;
            float4 ev = *ep;
            const float4 mv = reinterpret_cast<const float4*>(MB + row * CW)[f4];
            ev.x += mv.x; ev.y += mv.y; ev.z += mv.z; ev.w += mv.w;
            *ep = ev;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
              float4* tp = reinterpret_cast<float4*>(pp.te + dd * pp.st_e + (size_t)(e0 + r0 + row) * F + fbase) + f4;
              float4 tv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *tp;
              const float4 tm = reinterpret_cast<const float4*>(X0t + dd * SX0 + row * CW)[f4];
              tv.x += tm.x; tv.y += tm.y; tv.z += tm.z; tv.w += tm.w;
              *tp = tv;
            }
          }
        } else {
          for (int idx = tid; idx < n * CW; idx += TIB_THREADS) {
            const int j = idx / CW, f = idx % CW, fg = fbase + f;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            float t0[D], t1[D], t2[D];
#pragma unroll
            for (int dd = 0; dd < D; ++dd) { t0[dd] = 0.f; t1[dd] = 0.f; t2[dd] = 0.f; }
            float vj0 = 0.f, vj1 = 0.f, vj2 = 0.f;
            float wj0[D], wj1[D], wj2[D];
            if (sp == 4) {
              const float* vj = p.v_old + (size_t)(n0 + j) * 3 * F + fg;
              vj0 = vj[0]; vj1 = vj[F]; vj2 = vj[2 * F];
#pragma unroll
              for (int dd = 0; dd < D; ++dd) {
                const float* wj = pp.tv_old + dd * pp.st_v + (size_t)(n0 + j) * 3 * F + fg;
                wj0[dd] = wj[0]; wj1[dd] = wj[F]; wj2[dd] = wj[2 * F];
              }
            }
            for (int i = i_lo; i <= i_hi; ++i) {
              if (i == j) continue;
              const int rl = i * nm1 + j - (j > i) - r0;
              if (rl < 0 || rl >= rows) continue;
              const float m = MB[rl * CW + f];
              if (sp == 0) {          // gates * v[src]
                const float* vi = p.v_old + (size_t)(n0 + i) * 3 * F + fg;
                const float v0 = vi[0], v1 = vi[F], v2 = vi[2 * F];
                a0 = fmaf(m, v0, a0); a1 = fmaf(m, v1, a1); a2 = fmaf(m, v2, a2);
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                  const float tm = X0t[dd * SX0 + rl * CW + f];
                  const float* wi = pp.tv_old + dd * pp.st_v + (size_t)(n0 + i) * 3 * F + fg;
                  t0[dd] += tm * v0 + m * wi[0]; t1[dd] += tm * v1 + m * wi[F]; t2[dd] += tm * v2 + m * wi[2 * F];
                }
              } else if (sp == 1) {   // scale_edge_dir * dir
                const float dx = GEO[rl * 4 + 1], dy = GEO[rl * 4 + 2], dz = GEO[rl * 4 + 3];
                a0 = fmaf(m, dx, a0); a1 = fmaf(m, dy, a1); a2 = fmaf(m, dz, a2);
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                  const float tm = X0t[dd * SX0 + rl * CW + f];
                  const float* gt = GEOt + dd * SG + rl * 4;
                  t0[dd] += tm * dx + m * gt[1]; t1[dd] += tm * dy + m * gt[2]; t2[dd] += tm * dz + m * gt[3];
                }
              } else if (sp == 2) {   // ds
                a0 += m;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) t0[dd] += X0t[dd * SX0 + rl * CW + f];
              } else {                // cross_product_gates * (dir x v[dst])           (cpainn.py:296-300)
                const float dx = GEO[rl * 4 + 1], dy = GEO[rl * 4 + 2], dz = GEO[rl * 4 + 3];
                const float c0x = __fmul_rn(dy, vj2) - __fmul_rn(dz, vj1);
                const float c1x = __fmul_rn(dz, vj0) - __fmul_rn(dx, vj2);
                const float c2x = __fmul_rn(dx, vj1) - __fmul_rn(dy, vj0);
                a0 = fmaf(m, c0x, a0); a1 = fmaf(m, c1x, a1); a2 = fmaf(m, c2x, a2);
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                  const float tm = X0t[dd * SX0 + rl * CW + f];
                  const float* gt = GEOt + dd * SG + rl * 4;
                  const float ex = gt[1], ey = gt[2], ez = gt[3];
                  const float d0 = (ey * vj2 - ez * vj1) + (dy * wj2[dd] - dz * wj1[dd]);
                  const float d1 = (ez * vj0 - ex * vj2) + (dz * wj0[dd] - dx * wj2[dd]);
                  const float d2 = (ex * vj1 - ey * vj0) + (dx * wj1[dd] - dy * wj0[dd]);
                  t0[dd] += tm * c0x + m * d0; t1[dd] += tm * c1x + m * d1; t2[dd] += tm * c2x + m * d2;
                }
              }
            }
            if (sp == 2) {
              p.s_new[(size_t)(n0 + j) * F + fg] += a0;
#pragma unroll
              for (int dd = 0; dd < D; ++dd) pp.ts_new[dd * pp.st_s + (size_t)(n0 + j) * F + fg] += t0[dd];
            } else {
              float* vn = p.v_new + (size_t)(n0 + j) * 3 * F + fg;
              vn[0] += a0; vn[F] += a1; vn[2 * F] += a2;
#pragma unroll
              for (int dd = 0; dd < D; ++dd) {
                float* wn = pp.tv_new + dd * pp.st_v + (size_t)(n0 + j) * 3 * F + fg;
                wn[0] += t0[dd]; wn[F] += t1[dd]; wn[2 * F] += t2[dd];
              }
            }
          }
        }
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// update with tangents: Update.forward                                       (cpainn.py:345-376)
// ---------------------------------------------------------------------------------------------
struct UpdateJvpP {
  UpdateP u;
  float* ts;               // [D][N][F]    in place
  float* tv;               // [D][N][3][F] in place
  size_t st_s, st_v;
};

template <int F, int RPT, int D>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_update_jvp(UpdateJvpP pp) {
  using C = Cols<F>;
  constexpr int TR = 8 * RPT;
  const UpdateP& p = pp.u;
  extern __shared__ __align__(16) float smem[];
  constexpr int SET = TR * 10 * F;        // one set: VIN [3][TR][F], UV [3][TR][F], X0 [TR][2F], XA [TR][F], XB [TR][F]
  float* VIN = smem;
  float* UV = VIN + 3 * TR * F;
  float* X0 = UV + 3 * TR * F;
  float* XA = X0 + TR * 2 * F;
  float* XB = XA + TR * F;
  float* T = smem + SET;                  // D tangent sets with the same layout
  constexpr int O_UV = 3 * TR * F, O_X0 = 6 * TR * F, O_XA = 8 * TR * F, O_XB = 9 * TR * F;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node0 = blockIdx.x * TR;

  for (int idx = tid; idx < TR * 3 * (F / 4); idx += TIB_THREADS) {
    const int row = idx / (3 * (F / 4)), rem = idx % (3 * (F / 4)), xyz = rem / (F / 4), f4 = rem % (F / 4);
    const int node = node0 + row;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 val = zero;
    if (node < p.n_nodes) val = reinterpret_cast<const float4*>(p.v + ((size_t)node * 3 + xyz) * F)[f4];
    reinterpret_cast<float4*>(VIN + ((size_t)xyz * TR + row) * F)[f4] = val;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      float4 tv = zero;
      if (node < p.n_nodes) tv = reinterpret_cast<const float4*>(pp.tv + dd * pp.st_v + ((size_t)node * 3 + xyz) * F)[f4];
      reinterpret_cast<float4*>(T + dd * SET + ((size_t)xyz * TR + row) * F)[f4] = tv;
    }
  }
  for (int idx = tid; idx < TR * (F / 4); idx += TIB_THREADS) {
    const int row = idx / (F / 4), f4 = idx % (F / 4), node = node0 + row;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 val = zero;
    if (node < p.n_nodes) val = reinterpret_cast<const float4*>(p.s + (size_t)node * F)[f4];
    reinterpret_cast<float4*>(X0 + row * 2 * F + F)[f4] = val;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      float4 ts = zero;
      if (node < p.n_nodes) ts = reinterpret_cast<const float4*>(pp.ts + dd * pp.st_s + (size_t)node * F)[f4];
      reinterpret_cast<float4*>(T + dd * SET + O_X0 + row * 2 * F + F)[f4] = ts;
    }
  }
  __syncthreads();

  // q = |V v|, q_dot = (V v . V v_dot) / q ; U v and U v_dot                  (cpainn.py:355-361,392-403)
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch) {
    const int col = ch * C::CW + lane * C::CPL;
    float sq[RPT][C::CPL], dot[D][RPT][C::CPL];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) {
        sq[q][c] = 0.0f;
#pragma unroll
        for (int dd = 0; dd < D; ++dd) dot[dd][q][c] = 0.0f;
      }
#pragma unroll 1
    for (int xyz = 0; xyz < 3; ++xyz) {
      float acc[1 + D][RPT][C::CPL];
      gemm_rows_multi<RPT, C::CPL, 1 + D>(acc, VIN + (size_t)xyz * TR * F, T + (size_t)xyz * TR * F, SET, F, F, p.Vt, F, col, warp);
#pragma unroll
      for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) {
          sq[q][c] = __fadd_rn(sq[q][c], __fmul_rn(acc[0][q][c], acc[0][q][c]));
#pragma unroll
          for (int dd = 0; dd < D; ++dd) dot[dd][q][c] += acc[0][q][c] * acc[1 + dd][q][c];
        }
      gemm_rows_multi<RPT, C::CPL, 1 + D>(acc, VIN + (size_t)xyz * TR * F, T + (size_t)xyz * TR * F, SET, F, F, p.Ut, F, col, warp);
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        st_vec<C::CPL>(UV + ((size_t)xyz * TR + q * 8 + warp) * F + col, acc[0][q]);
#pragma unroll
        for (int dd = 0; dd < D; ++dd)
          st_vec<C::CPL>(T + dd * SET + O_UV + ((size_t)xyz * TR + q * 8 + warp) * F + col, acc[1 + dd][q]);
      }
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      float nrm[C::CPL];
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) nrm[c] = __fsqrt_rn(sq[q][c]);   // vv.norm(dim=-1), cpainn.py:361
      st_vec<C::CPL>(X0 + (q * 8 + warp) * 2 * F + col, nrm);
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        float qd[C::CPL];
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) qd[c] = nrm[c] > 0.0f ? dot[dd][q][c] / nrm[c] : 0.0f;
        st_vec<C::CPL>(T + dd * SET + O_X0 + (q * 8 + warp) * 2 * F + col, qd);
      }
    }
  }
  __syncwarp();

  layer_ln_silu_jvp<F, RPT, D>(X0, T + O_X0, SET, 2 * F, 2 * F, p.mlp.W1t, p.mlp.b1, p.mlp.g1, p.mlp.be1, XA, T + O_XA, SET, F, warp, lane);
  layer_ln_silu_jvp<F, RPT, D>(XA, T + O_XA, SET, F, F, p.mlp.W2t, p.mlp.b2, p.mlp.g2, p.mlp.be2, XB, T + O_XB, SET, F, warp, lane);

  // split order: gates, scale_squared_norm, add_invariant_features               (cpainn.py:366-368)
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch) {
    const int col = ch * C::CW + lane * C::CPL;
    {
      float g[1 + D][RPT][C::CPL], bg[C::CPL];
      gemm_rows_multi<RPT, C::CPL, 1 + D>(g, XB, T + O_XB, SET, F, F, p.mlp.W3t, 3 * F, col, warp);
      ldg_vec<C::CPL>(bg, p.mlp.b3 + col);
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int row = q * 8 + warp, node = node0 + row;
        if (node >= p.n_nodes) continue;
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) g[0][q][c] += bg[c];
#pragma unroll
        for (int xyz = 0; xyz < 3; ++xyz) {
          const size_t at = ((size_t)xyz * TR + row) * F + col;
          float o[C::CPL];
#pragma unroll
          for (int c = 0; c < C::CPL; ++c)   // v + uv * gates                        (cpainn.py:370,374)
            o[c] = __fadd_rn(VIN[at + c], __fmul_rn(UV[at + c], g[0][q][c]));
          st_vec<C::CPL>(p.v + ((size_t)node * 3 + xyz) * F + col, o);
#pragma unroll
          for (int dd = 0; dd < D; ++dd) {
            const float* Ts = T + dd * SET;
#pragma unroll
            for (int c = 0; c < C::CPL; ++c) o[c] = Ts[at + c] + Ts[O_UV + at + c] * g[0][q][c] + UV[at + c] * g[1 + dd][q][c];
            st_vec<C::CPL>(pp.tv + dd * pp.st_v + ((size_t)node * 3 + xyz) * F + col, o);
          }
        }
      }
    }
    float a[1 + D][RPT][C::CPL], cc[1 + D][RPT][C::CPL], ba[C::CPL], bc[C::CPL];
    gemm_rows_multi<RPT, C::CPL, 1 + D>(a, XB, T + O_XB, SET, F, F, p.mlp.W3t, 3 * F, F + col, warp);
    gemm_rows_multi<RPT, C::CPL, 1 + D>(cc, XB, T + O_XB, SET, F, F, p.mlp.W3t, 3 * F, 2 * F + col, warp);
    ldg_vec<C::CPL>(ba, p.mlp.b3 + F + col);
    ldg_vec<C::CPL>(bc, p.mlp.b3 + 2 * F + col);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = q * 8 + warp, node = node0 + row;
      if (node >= p.n_nodes) continue;
      float o[C::CPL], nq[C::CPL];
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) {
        a[0][q][c] += ba[c];
        cc[0][q][c] += bc[c];
        nq[c] = X0[row * 2 * F + col + c];
        const float ds = __fadd_rn(__fmul_rn(__fmul_rn(nq[c], nq[c]), a[0][q][c]), cc[0][q][c]);   // cpainn.py:371
        o[c] = __fadd_rn(X0[row * 2 * F + F + col + c], ds);                                         // cpainn.py:373
      }
      st_vec<C::CPL>(p.s + (size_t)node * F + col, o);
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        const float* Ts = T + dd * SET;
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) {
          const float nqd = Ts[O_X0 + row * 2 * F + col + c];
          o[c] = Ts[O_X0 + row * 2 * F + F + col + c] + 2.0f * nq[c] * nqd * a[0][q][c] + nq[c] * nq[c] * a[1 + dd][q][c] +
                 cc[1 + dd][q][c];
        }
        st_vec<C::CPL>(pp.ts + dd * pp.st_s + (size_t)node * F + col, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// readout with tangents: LayerReadout.forward, n_features_out = 1             (cpainn.py:425-437)
// ---------------------------------------------------------------------------------------------
struct ReadoutJvpP {
  ReadoutP r;              // r.out may be null (primal drift not wanted in this pass)
  const float* ts;         // [D][N][F]
  const float* tv;         // [D][N][3][F]
  float* tout;             // [D][N][3]
  size_t st_s, st_v, st_o;
};

template <int F, int RPT, int D>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_readout_jvp(ReadoutJvpP pp) {
  constexpr int TR = 8 * RPT;
  const ReadoutP& p = pp.r;
  extern __shared__ __align__(16) float smem[];
  constexpr int SET = TR * 3 * F;         // X0, XA, XB [TR][F]
  float* X0 = smem;
  float* XA = X0 + TR * F;
  float* XB = XA + TR * F;
  float* T = smem + SET;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node0 = blockIdx.x * TR;
  for (int idx = tid; idx < TR * (F / 4); idx += TIB_THREADS) {
    const int row = idx / (F / 4), f4 = idx % (F / 4), node = node0 + row;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 val = zero;
    if (node < p.n_nodes) val = reinterpret_cast<const float4*>(p.s + (size_t)node * F)[f4];
    reinterpret_cast<float4*>(X0 + row * F)[f4] = val;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      float4 ts = zero;
      if (node < p.n_nodes) ts = reinterpret_cast<const float4*>(pp.ts + dd * pp.st_s + (size_t)node * F)[f4];
      reinterpret_cast<float4*>(T + dd * SET + row * F)[f4] = ts;
    }
  }
  __syncthreads();
  layer_ln_silu_jvp<F, RPT, D>(X0, T, SET, F, F, p.mlp.W1t, p.mlp.b1, p.mlp.g1, p.mlp.be1, XA, T + TR * F, SET, F, warp, lane);
  layer_ln_silu_jvp<F, RPT, D>(XA, T + TR * F, SET, F, F, p.mlp.W2t, p.mlp.b2, p.mlp.g2, p.mlp.be2, XB, T + 2 * TR * F, SET, F, warp, lane);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int row = q * 8 + warp, node = node0 + row;
    if (node >= p.n_nodes) continue;   // warp-uniform
    float g = 0.0f;
    for (int k = lane; k < F; k += 32) g = fmaf(XB[row * F + k], __ldg(p.mlp.W3t + F + k), g);
    g = warp_sum(g) + __ldg(p.mlp.b3 + 1);
    float vo[3], o[3];
#pragma unroll
    for (int xyz = 0; xyz < 3; ++xyz) {
      float a = 0.0f;
      for (int k = lane; k < F; k += 32) a = fmaf(__ldg(p.Vout + k), p.v[((size_t)node * 3 + xyz) * F + k], a);
      vo[xyz] = warp_sum(a);
      o[xyz] = __fmul_rn(vo[xyz], g);
    }
    if (p.out && lane < 3) p.out[(size_t)node * 3 + lane] = o[lane];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      float tg = 0.0f;
      for (int k = lane; k < F; k += 32) tg = fmaf(T[dd * SET + 2 * TR * F + row * F + k], __ldg(p.mlp.W3t + F + k), tg);
      tg = warp_sum(tg);
      float to[3];
#pragma unroll
      for (int xyz = 0; xyz < 3; ++xyz) {
        float a = 0.0f;
        for (int k = lane; k < F; k += 32)
          a = fmaf(__ldg(p.Vout + k), pp.tv[dd * pp.st_v + ((size_t)node * 3 + xyz) * F + k], a);
        to[xyz] = warp_sum(a) * g + vo[xyz] * tg;
      }
      if (lane < 3) pp.tout[dd * pp.st_o + (size_t)node * 3 + lane] = to[lane];
    }
  }
}

// div[mol] (+)= sum_dd tout[dd][mol_ptr[mol] + a_dd][c_dd]; the pass with dir0 == 0 starts the sum.
// skip_last_of = n_max > 0: the directions of atom index n_max - 1 are NOT propagated.  The drift depends on x through
// differences only (graph.py:27), so sum_a d b / d x_{a,c} = 0 as a field and, for a molecule with n = n_max atoms,
//   d b_{n-1,c} / d x_{n-1,c} = - sum_{a < n-1} d b_{n-1,c} / d x_{a,c}:
// the missing diagonal entries come out of the other directions' tangents at the last atom.
__global__ void k_div_pick(const int* __restrict__ mol_ptr, int n_mol, const float* __restrict__ tout, size_t st_o,
                           int dir0, int D, float* __restrict__ div, int skip_last_of = 0) {
  const int mol = blockIdx.x * blockDim.x + threadIdx.x;
  if (mol >= n_mol) return;
  const int n0 = __ldg(mol_ptr + mol), n = __ldg(mol_ptr + mol + 1) - n0;
  float acc = dir0 == 0 ? 0.0f : div[mol];
  for (int dd = 0; dd < D; ++dd) {
    const int q = dir0 + dd, a = q / 3, c = q % 3;
    if (a < n) acc += tout[dd * st_o + (size_t)(n0 + a) * 3 + c];
    if (skip_last_of > 0 && n == skip_last_of && a < n - 1) acc -= tout[dd * st_o + (size_t)(n0 + n - 1) * 3 + c];
  }
  div[mol] = acc;
}

template <int F, int RPT, int D> constexpr size_t smem_message_jvp() {
  return sizeof(float) * ((size_t)(1 + D) * (8 * RPT) * (5 * F + 4) + TIB_MAX_ATOMS * 3);
}
template <int F, int RPT, int D> constexpr size_t smem_update_jvp() { return sizeof(float) * (size_t)(1 + D) * (8 * RPT) * (10 * F); }
template <int F, int RPT, int D> constexpr size_t smem_readout_jvp() { return sizeof(float) * (size_t)(1 + D) * (8 * RPT) * (3 * F); }

}  // namespace tib
