// simt_drift.cuh - the cPaiNN drift b(t, x, T) in fp32 on CUDA cores (TIB_MATH_FP32_SIMT).
//
// State between kernels (HBM, owned by the caller's workspace):
//   s [N][F]      invariant node features
//   v [N][3][F]   equivariant node features, xyz-major planes so every feature row is contiguous
//                 (the reference keeps [N,F,3]; the transposition never leaves the library)
//   e [E][F]      invariant edge features, rows in the reference's (src,dst)-lexicographic order
// Kernels per drift evaluation: embed, edge_init, L x (message, update), readout.
#pragma once
#include "simt_mlp.cuh"

namespace tib {

struct DriftBatch {
  int n_mol, n_nodes;
  long long n_edges;
  const int* mol_ptr;
  const long long* edge_ptr;
  const int* atom_id;
  const unsigned char* edge_type;
  const float* temp0;
  const float* temp1;
};

struct EmbedP {
  DriftBatch b;
  MlpW mlp;
  const float* atom_emb;  // [n_types][F]
  int n_temp;             // number of temperature encoders (0,1,2)
  float t, temp_mean, temp_range, temp_length, time_length;
  float* s_out;           // [N][F]
};

// ---------------------------------------------------------------------------------------------
// embed: s0 = MLP(cat[Emb(atom), PE(T0'), PE(T1'), PE(t)])          (embedding.py:68-86,249-261)
// ---------------------------------------------------------------------------------------------
template <int F, int RPT, int KU = 4>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_embed(EmbedP p) {
  using C = Cols<F>;
  constexpr int TR = 8 * RPT;
  extern __shared__ __align__(16) float smem[];
  const int kin = (2 + p.n_temp) * F;
  float* X0 = smem;                 // [TR][kin]
  float* XA = X0 + TR * 4 * F;      // [TR][F]   (X0 sized for the ambient worst case 4F)
  float* XB = XA + TR * F;          // [TR][F]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node0 = blockIdx.x * TR;
  const int nseg = 2 + p.n_temp;

  // segment 0: atom embedding rows
  for (int idx = tid; idx < TR * F; idx += TIB_THREADS) {
    const int row = idx / F, f = idx % F, node = node0 + row;
    X0[row * kin + f] = node < p.b.n_nodes ? __ldg(p.atom_emb + (size_t)__ldg(p.b.atom_id + node) * F + f) : 0.0f;
  }
  // segments 1..: positional encodings, one (cos,sin) pair per thread-iteration
  const int npair = F / 2;
  for (int idx = tid; idx < TR * (nseg - 1) * npair; idx += TIB_THREADS) {
    const int row = idx / ((nseg - 1) * npair);
    const int rem = idx % ((nseg - 1) * npair);
    const int seg = 1 + rem / npair, rank = 1 + rem % npair;
    const int node = node0 + row;
    float cs = 0.0f, sn = 0.0f;
    if (node < p.b.n_nodes) {
      float val, len;
      if (seg <= p.n_temp) {
        const float T = __ldg((seg == 1 ? p.b.temp0 : p.b.temp1) + node);
        val = __fdiv_rn(T - p.temp_mean, p.temp_range);   // TemperatureEncoder, embedding.py:208-209
        len = p.temp_length;
      } else {
        val = p.t;                                        // batch.t, ode_wrapper.py:112
        len = p.time_length;
      }
      sincosf(pe_arg(val, len, rank), &sn, &cs);
    }
    float* o = X0 + row * kin + seg * F + 2 * (rank - 1);
    o[0] = cs;                                            // torch.stack((cos, sin)), embedding.py:160
    o[1] = sn;
  }
  __syncthreads();

  layer_ln_silu<F, RPT, KU>(X0, kin, kin, p.mlp.W1t, p.mlp.b1, p.mlp.g1, p.mlp.be1, XA, F, warp, lane);
  layer_ln_silu<F, RPT, KU>(XA, F, F, p.mlp.W2t, p.mlp.b2, p.mlp.g2, p.mlp.be2, XB, F, warp, lane);
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch) {
    float acc[RPT][C::CPL];
    out_chunk<F, RPT, KU>(acc, XB, F, p.mlp.W3t, F, p.mlp.b3, ch * C::CW, warp, lane);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int node = node0 + q * 8 + warp;
      if (node < p.b.n_nodes) st_vec<C::CPL>(p.s_out + (size_t)node * F + ch * C::CW + lane * C::CPL, acc[q]);
    }
  }
}

// First message layer, phi's hidden layers per distinct input: s0 takes n_rows distinct values (the de-duplicated
// embedding rows) and e0 = Emb(edge type) n_et values, so H[u * n_et + t] = SiLU(LN(W2 SiLU(LN(W1 cat[s0[u], e0[t]] + b1)) + b2))
// is evaluated once per (u, t) instead of once per edge (cpainn.py:275-281 with embedding.py:26-34).
struct PhiTabP {
  MlpW phi;
  const float* s0;        // [n_rows][F]
  const float* edge_emb;  // [n_et][F]
  int n_rows, n_et;
  float* out;             // [n_rows * n_et][F]
};

template <int F, int KU>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_phi_table(PhiTabP p) {
  constexpr int TR = 8;
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;                 // [TR][2F]
  float* XA = X0 + TR * 2 * F;      // [TR][F]
  float* XB = XA + TR * F;          // [TR][F]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * TR, total = p.n_rows * p.n_et;
  for (int idx = tid; idx < TR * 2 * F; idx += TIB_THREADS) {
    const int row = idx / (2 * F), f = idx % (2 * F), r = r0 + row;
    float val = 0.0f;
    if (r < total) val = f < F ? __ldg(p.s0 + (size_t)(r / p.n_et) * F + f) : __ldg(p.edge_emb + (size_t)(r % p.n_et) * F + f - F);
    X0[idx] = val;
  }
  __syncthreads();
  layer_ln_silu<F, 1, KU>(X0, 2 * F, 2 * F, p.phi.W1t, p.phi.b1, p.phi.g1, p.phi.be1, XA, F, warp, lane);
  layer_ln_silu<F, 1, KU>(XA, F, F, p.phi.W2t, p.phi.b2, p.phi.g2, p.phi.be2, XB, F, warp, lane);
  const int r = r0 + warp;
  if (r < total)
    for (int f = lane; f < F; f += 32) p.out[(size_t)r * F + f] = XB[warp * F + f];
}

// ---------------------------------------------------------------------------------------------
// One CTA per table row, K split over the 8 warps: for the few de-duplicated rows of a single-species batch the
// row-per-warp kernels above are one long chain of dependent weight loads (42 + 30 us for 9 + 36 rows); here every
// layer is 1/8 of that chain plus a shared-memory reduction.  The partial sums are added in a fixed order, which
// is NOT the k-ascending fmaf chain of the fp32 path - used by the tensor-core modes only (F = 128).
// ---------------------------------------------------------------------------------------------
template <int F>
__device__ __forceinline__ void splitk_linear(const float* x, int K, const float* __restrict__ Wt, int ldw, float* part,
                                              int warp, int lane) {
  // part[warp][c] = sum over this warp's slice of k of x[k] * Wt[k][c],  c = 4 * lane .. 4 * lane + 3
  static_assert(F == 128, "split-K helpers are built for F = 128");
  const int ks = K / TIB_WARPS, k0 = warp * ks;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int kb = 0; kb < ks; kb += 16) {
    float4 w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + kb + i) * ldw) + lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = x[k0 + kb + i];
      acc[0] = fmaf(a, w[i].x, acc[0]); acc[1] = fmaf(a, w[i].y, acc[1]);
      acc[2] = fmaf(a, w[i].z, acc[2]); acc[3] = fmaf(a, w[i].w, acc[3]);
    }
  }
  *reinterpret_cast<float4*>(part + warp * F + 4 * lane) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// out[c] = SiLU(LN(x @ Wt + b)) (ln = true) or x @ Wt + b (ln = false); one row, all 256 threads
template <int F>
__device__ __forceinline__ void splitk_layer(const float* x, int K, const float* __restrict__ Wt, const float* __restrict__ b,
                                             const float* __restrict__ g, const float* __restrict__ be, bool ln, float* out,
                                             float* part, int warp, int lane) {
  splitk_linear<F>(x, K, Wt, F, part, warp, lane);
  __syncthreads();
  if (warp == 0) {
    float z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < TIB_WARPS; ++w) t += part[w * F + 4 * lane + i];
      z[i] = t + __ldg(b + 4 * lane + i);
    }
    if (ln) {
      const float mean = warp_sum((z[0] + z[1]) + (z[2] + z[3])) * (1.0f / F);
      float ss = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { z[i] -= mean; ss = fmaf(z[i], z[i], ss); }
      const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(warp_sum(ss) * (1.0f / F) + 1e-5f));
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = silu(z[i] * rstd * __ldg(g + 4 * lane + i) + __ldg(be + 4 * lane + i));
    }
    *reinterpret_cast<float4*>(out + 4 * lane) = make_float4(z[0], z[1], z[2], z[3]);
  }
  __syncthreads();
}

struct EmbedTabP {
  EmbedP e;               // e.b.n_nodes = number of table rows U; e.s_out = s0 table [U][F]
  MlpW phi;               // first message layer's phi (hidden layers only)
  const float* edge_emb;  // [n_et][F]
  int n_et;
  float* phitab;          // [U * n_et][F] or NULL (then the grid is U CTAs)
};

// CTA (u, t): s0[u] = InvariantFeatures MLP of table row u (written by t == 0), then the phi table row (u, t)
template <int F>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_embed_tab(EmbedTabP pp) {
  const EmbedP& p = pp.e;
  __shared__ __align__(16) float X0[4 * F], XA[2 * F], XB[F], PART[TIB_WARPS * F];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_et = pp.phitab ? pp.n_et : 1;
  const int u = blockIdx.x / n_et, et = blockIdx.x % n_et;
  const int nseg = 2 + p.n_temp, kin = nseg * F, npair = F / 2;
  for (int f = tid; f < F; f += TIB_THREADS) X0[f] = __ldg(p.atom_emb + (size_t)__ldg(p.b.atom_id + u) * F + f);
  for (int idx = tid; idx < (nseg - 1) * npair; idx += TIB_THREADS) {
    const int seg = 1 + idx / npair, rank = 1 + idx % npair;
    float val, len;
    if (seg <= p.n_temp) {
      const float T = __ldg((seg == 1 ? p.b.temp0 : p.b.temp1) + u);
      val = __fdiv_rn(T - p.temp_mean, p.temp_range);
      len = p.temp_length;
    } else {
      val = p.t;
      len = p.time_length;
    }
    float sn, cs;
    sincosf(pe_arg(val, len, rank), &sn, &cs);
    X0[seg * F + 2 * (rank - 1)] = cs;
    X0[seg * F + 2 * (rank - 1) + 1] = sn;
  }
  __syncthreads();
  splitk_layer<F>(X0, kin, p.mlp.W1t, p.mlp.b1, p.mlp.g1, p.mlp.be1, true, XA, PART, warp, lane);
  splitk_layer<F>(XA, F, p.mlp.W2t, p.mlp.b2, p.mlp.g2, p.mlp.be2, true, XB, PART, warp, lane);
  splitk_layer<F>(XB, F, p.mlp.W3t, p.mlp.b3, nullptr, nullptr, false, XA, PART, warp, lane);   // s0[u] in XA[0:F]
  if (et == 0)
    for (int f = tid; f < F; f += TIB_THREADS) p.s_out[(size_t)u * F + f] = XA[f];
  if (!pp.phitab) return;
  for (int f = tid; f < F; f += TIB_THREADS) XA[F + f] = __ldg(pp.edge_emb + (size_t)et * F + f);
  __syncthreads();
  splitk_layer<F>(XA, 2 * F, pp.phi.W1t, pp.phi.b1, pp.phi.g1, pp.phi.be1, true, XB, PART, warp, lane);
  splitk_layer<F>(XB, F, pp.phi.W2t, pp.phi.b2, pp.phi.g2, pp.phi.be2, true, X0, PART, warp, lane);
  for (int f = tid; f < F; f += TIB_THREADS) pp.phitab[(size_t)blockIdx.x * F + f] = X0[f];
}

// out[r][:] = table[index[r]][:]   (de-duplicated node embedding -> per-node rows)
__global__ void k_gather_rows(const float* __restrict__ table, const int* __restrict__ index, float* __restrict__ out,
                              int n_rows, int F) {
  const long long total = (long long)n_rows * (F / 4);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / (F / 4);
    const int f4 = (int)(idx % (F / 4));
    reinterpret_cast<float4*>(out + (size_t)row * F)[f4] =
        __ldg(reinterpret_cast<const float4*>(table + (size_t)__ldg(index + row) * F) + f4);
  }
}

// e0 = Emb4(edge_type)                                              (embedding.py:89-103, cpainn.py:70)
__global__ void k_edge_init(const unsigned char* __restrict__ edge_type, const float* __restrict__ edge_emb,
                            float* __restrict__ e, long long n_edges, int F) {
  const long long total = n_edges * (F / 4);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / (F / 4);
    const int f4 = (int)(idx % (F / 4));
    const float4 val = __ldg(reinterpret_cast<const float4*>(edge_emb + (size_t)edge_type[row] * F) + f4);
    reinterpret_cast<float4*>(e + (size_t)row * F)[f4] = val;
  }
}

// ---------------------------------------------------------------------------------------------
// message: SE3Message.forward                                        (cpainn.py:263-310)
//   one CTA per molecule; edges tiled TR rows at a time in (src,dst) order.
// ---------------------------------------------------------------------------------------------
struct MessageP {
  DriftBatch b;
  MlpW phi, w;
  const float* x;         // [N][3]
  const float* s_old;     // [N][F]
  const float* v_old;     // [N][3][F]
  float* s_new;
  float* v_new;
  float* e;               // [E][F] updated in place
  float length_scale;
  int first_layer;        // v_old == 0: skip reading it (gates/cross terms vanish)
};

template <int F, int RPT>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_message(MessageP p) {
  using C = Cols<F>;
  constexpr int TR = 8 * RPT;
  constexpr int CW = C::CW;
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;                 // [TR][2F]  phi input, then PE(d), then the phi*w product chunk
  float* XA = X0 + TR * 2 * F;      // [TR][F]
  float* HP = XA + TR * F;          // [TR][F]   phi hidden 2
  float* HW = HP + TR * F;          // [TR][F]   w hidden 2
  float* GEO = HW + TR * F;         // [TR][4]   d, dir.xyz
  float* XS = GEO + TR * 4;         // [TIB_MAX_ATOMS][3]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mol = blockIdx.x;
  const int n0 = __ldg(p.b.mol_ptr + mol);
  const int n = __ldg(p.b.mol_ptr + mol + 1) - n0;
  const int ne = n * (n - 1);
  const long long e0 = __ldg(p.b.edge_ptr + mol);
  const int nm1 = n - 1;

  // v_new = v_old, s_new = s_old for this molecule; coordinates to shared memory
  for (int idx = tid; idx < n * F; idx += TIB_THREADS) p.s_new[(size_t)n0 * F + idx] = p.s_old[(size_t)n0 * F + idx];
  for (int idx = tid; idx < n * 3 * F; idx += TIB_THREADS)
    p.v_new[(size_t)n0 * 3 * F + idx] = p.first_layer ? 0.0f : p.v_old[(size_t)n0 * 3 * F + idx];
  for (int idx = tid; idx < n * 3; idx += TIB_THREADS) XS[idx] = p.x[(size_t)n0 * 3 + idx];
  __syncthreads();

  for (int r0 = 0; r0 < ne; r0 += TR) {
    const int rows = min(TR, ne - r0);
    // ---- geometry of the tile's edges: r = x[src]-x[dst], d = |r|, dir = r/(1+d)   (graph.py:27-29)
    if (tid < TR) {
      float d = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
      if (tid < rows) {
        const int r = r0 + tid, i = r / nm1, k = r % nm1, j = k + (k >= i);
        const float rx = XS[i * 3 + 0] - XS[j * 3 + 0];
        const float ry = XS[i * 3 + 1] - XS[j * 3 + 1];
        const float rz = XS[i * 3 + 2] - XS[j * 3 + 2];
        d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
        const float den = 1.0f + d;
        dx = __fdiv_rn(rx, den); dy = __fdiv_rn(ry, den); dz = __fdiv_rn(rz, den);
      }
      GEO[tid * 4 + 0] = d; GEO[tid * 4 + 1] = dx; GEO[tid * 4 + 2] = dy; GEO[tid * 4 + 3] = dz;
    }
    // ---- phi input: cat[s[src], e]                                              (cpainn.py:275-281)
    for (int idx = tid; idx < TR * (F / 4); idx += TIB_THREADS) {
      const int row = idx / (F / 4), f4 = idx % (F / 4);
      float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), ev = sv;
      if (row < rows) {
        const int i = (r0 + row) / nm1;
        sv = reinterpret_cast<const float4*>(p.s_old + (size_t)(n0 + i) * F)[f4];
        ev = reinterpret_cast<const float4*>(p.e + (size_t)(e0 + r0 + row) * F)[f4];
      }
      reinterpret_cast<float4*>(X0 + row * 2 * F)[f4] = sv;
      reinterpret_cast<float4*>(X0 + row * 2 * F + F)[f4] = ev;
    }
    __syncthreads();

    // ---- phi hidden layers (warp-local rows)
    layer_ln_silu<F, RPT>(X0, 2 * F, 2 * F, p.phi.W1t, p.phi.b1, p.phi.g1, p.phi.be1, XA, F, warp, lane);
    layer_ln_silu<F, RPT>(XA, F, F, p.phi.W2t, p.phi.b2, p.phi.g2, p.phi.be2, HP, F, warp, lane);
    // ---- w input: PositionalEncoder(edge_dist) written over X0[:, 0:F]           (cpainn.py:283)
    for (int idx = lane; idx < RPT * (F / 2); idx += 32) {
      const int q = idx / (F / 2), rank = 1 + idx % (F / 2);
      const int row = q * 8 + warp;
      float sn, cs;
      sincosf(pe_arg(GEO[row * 4], p.length_scale, rank), &sn, &cs);
      X0[row * 2 * F + 2 * (rank - 1)] = cs;
      X0[row * 2 * F + 2 * (rank - 1) + 1] = sn;
    }
    __syncwarp();
    layer_ln_silu<F, RPT>(X0, 2 * F, F, p.w.W1t, p.w.b1, p.w.g1, p.w.be1, XA, F, warp, lane);
    layer_ln_silu<F, RPT>(XA, F, F, p.w.W2t, p.w.b2, p.w.g2, p.w.be2, HW, F, warp, lane);
    __syncthreads();   // X0 is re-used as the CTA-wide product buffer from here on

    // ---- output layer, one (split, chunk) at a time: m = phi(...) * w(...)        (cpainn.py:285-290)
    // split order: gates, scale_edge_dir, ds, de, cross_product_gates
    float* MB = X0;    // [TR][CW]
    const int i_lo = r0 / nm1, i_hi = (r0 + rows - 1) / nm1;
    for (int sp = 0; sp < 5; ++sp) {
      if (p.first_layer && (sp == 0 || sp == 4)) continue;   // multiply v == 0
      for (int ch = 0; ch < C::NCH; ++ch) {
        const int c0 = sp * F + ch * CW;
        {
          float acc[RPT][C::CPL], accw[RPT][C::CPL];
          out_chunk<F, RPT>(acc, HP, F, p.phi.W3t, 5 * F, p.phi.b3, c0, warp, lane);
          out_chunk<F, RPT>(accw, HW, F, p.w.W3t, 5 * F, p.w.b3, c0, warp, lane);
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            float m[C::CPL];
#pragma unroll
            for (int c = 0; c < C::CPL; ++c) m[c] = __fmul_rn(acc[q][c], accw[q][c]);
            st_vec<C::CPL>(MB + (q * 8 + warp) * CW + lane * C::CPL, m);
          }
        }
        __syncthreads();
        const int fbase = ch * CW;
        if (sp == 3) {
          // e += de                                                                (cpainn.py:308)
          for (int idx = tid; idx < rows * (CW / 4); idx += TIB_THREADS) {
            const int row = idx / (CW / 4), f4 = idx % (CW / 4);
            float4* ep = reinterpret_cast<float4*>(p.e + (size_t)(e0 + r0 + row) * F + fbase) + f4;
            float4 ev = *ep;
            const float4 mv = reinterpret_cast<const float4*>(MB + row * CW)[f4];
            ev.x += mv.x; ev.y += mv.y; ev.z += mv.z; ev.w += mv.w;
            *ep = ev;
          }
        } else {
          // segmented sum over incoming edges of node j, sources in ascending order (the order
          // index_add_ visits them in the reference's scatter, cpainn.py:303-304)
          for (int idx = tid; idx < n * CW; idx += TIB_THREADS) {
            const int j = idx / CW, f = idx % CW, fg = fbase + f;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            float vj0 = 0.f, vj1 = 0.f, vj2 = 0.f;
            if (sp == 4) {
              const float* vj = p.v_old + (size_t)(n0 + j) * 3 * F + fg;
              vj0 = vj[0]; vj1 = vj[F]; vj2 = vj[2 * F];
            }
            for (int i = i_lo; i <= i_hi; ++i) {
              if (i == j) continue;
              const int rl = i * nm1 + j - (j > i) - r0;
              if (rl < 0 || rl >= rows) continue;
              const float m = MB[rl * CW + f];
              if (sp == 0) {          // gates * v[src]
                const float* vi = p.v_old + (size_t)(n0 + i) * 3 * F + fg;
                a0 = fmaf(m, vi[0], a0); a1 = fmaf(m, vi[F], a1); a2 = fmaf(m, vi[2 * F], a2);
              } else if (sp == 1) {   // scale_edge_dir * dir
                a0 = fmaf(m, GEO[rl * 4 + 1], a0); a1 = fmaf(m, GEO[rl * 4 + 2], a1); a2 = fmaf(m, GEO[rl * 4 + 3], a2);
              } else if (sp == 2) {   // ds
                a0 += m;
              } else {                // cross_product_gates * (dir x v[dst])       (cpainn.py:296-300)
                const float dx = GEO[rl * 4 + 1], dy = GEO[rl * 4 + 2], dz = GEO[rl * 4 + 3];
                const float c0x = __fmul_rn(dy, vj2) - __fmul_rn(dz, vj1);
                const float c1x = __fmul_rn(dz, vj0) - __fmul_rn(dx, vj2);
                const float c2x = __fmul_rn(dx, vj1) - __fmul_rn(dy, vj0);
                a0 = fmaf(m, c0x, a0); a1 = fmaf(m, c1x, a1); a2 = fmaf(m, c2x, a2);
              }
            }
            if (sp == 2) {
              p.s_new[(size_t)(n0 + j) * F + fg] += a0;
            } else {
              float* vn = p.v_new + (size_t)(n0 + j) * 3 * F + fg;
              vn[0] += a0; vn[F] += a1; vn[2 * F] += a2;
            }
          }
        }
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// update: Update.forward                                             (cpainn.py:345-376)
// ---------------------------------------------------------------------------------------------
struct UpdateP {
  int n_nodes;
  MlpW mlp;               // 2F -> F -> F -> 3F
  const float* Ut;        // [F][F] transposed
  const float* Vt;        // [F][F] transposed
  float* s;               // [N][F]    in place
  float* v;               // [N][3][F] in place
};

template <int F, int RPT>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_update(UpdateP p) {
  using C = Cols<F>;
  constexpr int TR = 8 * RPT;
  extern __shared__ __align__(16) float smem[];
  float* VIN = smem;                  // [3][TR][F]
  float* UV = VIN + 3 * TR * F;       // [3][TR][F]
  float* X0 = UV + 3 * TR * F;        // [TR][2F]  cat[|Vv|, s]
  float* XA = X0 + TR * 2 * F;        // [TR][F]
  float* XB = XA + TR * F;            // [TR][F]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node0 = blockIdx.x * TR;

  for (int idx = tid; idx < TR * 3 * (F / 4); idx += TIB_THREADS) {
    const int row = idx / (3 * (F / 4)), rem = idx % (3 * (F / 4)), xyz = rem / (F / 4), f4 = rem % (F / 4);
    const int node = node0 + row;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (node < p.n_nodes) val = reinterpret_cast<const float4*>(p.v + ((size_t)node * 3 + xyz) * F)[f4];
    reinterpret_cast<float4*>(VIN + ((size_t)xyz * TR + row) * F)[f4] = val;
  }
  for (int idx = tid; idx < TR * (F / 4); idx += TIB_THREADS) {
    const int row = idx / (F / 4), f4 = idx % (F / 4), node = node0 + row;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (node < p.n_nodes) val = reinterpret_cast<const float4*>(p.s + (size_t)node * F)[f4];
    reinterpret_cast<float4*>(X0 + row * 2 * F + F)[f4] = val;
  }
  __syncthreads();

  // |V v| over xyz and U v  (EquivariantLinear, cpainn.py:392-403)
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch) {
    float sq[RPT][C::CPL];
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) sq[q][c] = 0.0f;
    for (int xyz = 0; xyz < 3; ++xyz) {
      float acc[RPT][C::CPL];
      gemm_rows<RPT, C::CPL>(acc, VIN + (size_t)xyz * TR * F, F, F, p.Vt, F, ch * C::CW + lane * C::CPL, warp);
#pragma unroll
      for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int c = 0; c < C::CPL; ++c) sq[q][c] = __fadd_rn(sq[q][c], __fmul_rn(acc[q][c], acc[q][c]));
      gemm_rows<RPT, C::CPL>(acc, VIN + (size_t)xyz * TR * F, F, F, p.Ut, F, ch * C::CW + lane * C::CPL, warp);
#pragma unroll
      for (int q = 0; q < RPT; ++q)
        st_vec<C::CPL>(UV + ((size_t)xyz * TR + q * 8 + warp) * F + ch * C::CW + lane * C::CPL, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      float nrm[C::CPL];
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) nrm[c] = __fsqrt_rn(sq[q][c]);   // vv.norm(dim=-1), cpainn.py:361
      st_vec<C::CPL>(X0 + (q * 8 + warp) * 2 * F + ch * C::CW + lane * C::CPL, nrm);
    }
  }
  __syncwarp();

  layer_ln_silu<F, RPT>(X0, 2 * F, 2 * F, p.mlp.W1t, p.mlp.b1, p.mlp.g1, p.mlp.be1, XA, F, warp, lane);
  layer_ln_silu<F, RPT>(XA, F, F, p.mlp.W2t, p.mlp.b2, p.mlp.g2, p.mlp.be2, XB, F, warp, lane);

  // split order: gates, scale_squared_norm, add_invariant_features               (cpainn.py:366-368)
#pragma unroll
  for (int ch = 0; ch < C::NCH; ++ch) {
    const int col = ch * C::CW + lane * C::CPL;
    float acc[RPT][C::CPL];
    out_chunk<F, RPT>(acc, XB, F, p.mlp.W3t, 3 * F, p.mlp.b3, ch * C::CW, warp, lane);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = q * 8 + warp, node = node0 + row;
      if (node >= p.n_nodes) continue;
#pragma unroll
      for (int xyz = 0; xyz < 3; ++xyz) {
        float o[C::CPL];
#pragma unroll
        for (int c = 0; c < C::CPL; ++c)   // v + uv * gates                        (cpainn.py:370,374)
          o[c] = __fadd_rn(VIN[((size_t)xyz * TR + row) * F + col + c],
                           __fmul_rn(UV[((size_t)xyz * TR + row) * F + col + c], acc[q][c]));
        st_vec<C::CPL>(p.v + ((size_t)node * 3 + xyz) * F + col, o);
      }
    }
    float acc_c[RPT][C::CPL];
    out_chunk<F, RPT>(acc, XB, F, p.mlp.W3t, 3 * F, p.mlp.b3, F + ch * C::CW, warp, lane);
    out_chunk<F, RPT>(acc_c, XB, F, p.mlp.W3t, 3 * F, p.mlp.b3, 2 * F + ch * C::CW, warp, lane);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = q * 8 + warp, node = node0 + row;
      if (node >= p.n_nodes) continue;
      float o[C::CPL];
#pragma unroll
      for (int c = 0; c < C::CPL; ++c) {
        const float nq = X0[row * 2 * F + col + c];
        const float ds = __fadd_rn(__fmul_rn(__fmul_rn(nq, nq), acc[q][c]), acc_c[q][c]);   // cpainn.py:371
        o[c] = __fadd_rn(X0[row * 2 * F + F + col + c], ds);                                 // cpainn.py:373
      }
      st_vec<C::CPL>(p.s + (size_t)node * F + col, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// readout: LayerReadout.forward with n_features_out = 1               (cpainn.py:425-437)
// ---------------------------------------------------------------------------------------------
struct ReadoutP {
  int n_nodes;
  MlpW mlp;               // F -> F -> F -> 2 ; W3 kept [2][F] (row-major, NOT transposed)
  const float* Vout;      // [F]
  const float* s;         // [N][F]
  const float* v;         // [N][3][F]
  float* out;             // [N][3]
};

template <int F, int RPT>
__global__ void __launch_bounds__(TIB_THREADS, 1) k_readout(ReadoutP p) {
  constexpr int TR = 8 * RPT;
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;             // [TR][F]
  float* XA = X0 + TR * F;
  float* XB = XA + TR * F;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node0 = blockIdx.x * TR;
  for (int idx = tid; idx < TR * (F / 4); idx += TIB_THREADS) {
    const int row = idx / (F / 4), f4 = idx % (F / 4), node = node0 + row;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (node < p.n_nodes) val = reinterpret_cast<const float4*>(p.s + (size_t)node * F)[f4];
    reinterpret_cast<float4*>(X0 + row * F)[f4] = val;
  }
  __syncthreads();
  layer_ln_silu<F, RPT>(X0, F, F, p.mlp.W1t, p.mlp.b1, p.mlp.g1, p.mlp.be1, XA, F, warp, lane);
  layer_ln_silu<F, RPT>(XA, F, F, p.mlp.W2t, p.mlp.b2, p.mlp.g2, p.mlp.be2, XB, F, warp, lane);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int row = q * 8 + warp, node = node0 + row;
    if (node >= p.n_nodes) continue;   // warp-uniform
    // gate = second output of the MLP (first is the discarded invariant output)
    float g = 0.0f;
    for (int k = lane; k < F; k += 32) g = fmaf(XB[row * F + k], __ldg(p.mlp.W3t + F + k), g);
    g = warp_sum(g) + __ldg(p.mlp.b3 + 1);
    float o[3];
#pragma unroll
    for (int xyz = 0; xyz < 3; ++xyz) {
      float a = 0.0f;
      for (int k = lane; k < F; k += 32) a = fmaf(__ldg(p.Vout + k), p.v[((size_t)node * 3 + xyz) * F + k], a);
      o[xyz] = __fmul_rn(warp_sum(a), g);
    }
    if (lane < 3) p.out[(size_t)node * 3 + lane] = o[lane];
  }
}

// shared-memory footprints (bytes) of the kernels above
template <int F, int RPT> constexpr size_t smem_embed() { return sizeof(float) * (size_t)(8 * RPT) * (4 * F + 2 * F); }
template <int F, int RPT> constexpr size_t smem_message() {
  return sizeof(float) * ((size_t)(8 * RPT) * (5 * F + 4) + TIB_MAX_ATOMS * 3);
}
template <int F, int RPT> constexpr size_t smem_update() { return sizeof(float) * (size_t)(8 * RPT) * (10 * F); }
template <int F, int RPT> constexpr size_t smem_readout() { return sizeof(float) * (size_t)(8 * RPT) * (3 * F); }

}  // namespace tib
