// layered.cuh - the element-wise / gather-scatter kernels of the "layered" path (tc_chain.cuh does the GEMMs):
// the F = 256 drift and the forward-mode tangents of the exact divergence (ode_wrapper.py:59-91).
//
// Layout: rows of every edge array are in (dst, src) order (k_edge_tables: node_in_ptr, RowA {src, dst, flags | type << 16,
// dist}, RowB {dir}); phi3 / w3 / w3d are [E][5F] with the reference's split order gates | scale_edge_dir | ds | de |
// cross_gates (cpainn.py:285-290); tangent arrays carry a leading direction axis with explicit strides.
// Direction q = 3 a + c seeds x_dot = unit vector on coordinate c of atom a of EVERY molecule (molecules do not interact).
// All kernels: 128 threads, feature f = threadIdx.x + 128 j; HBM-bound (coalesced along the feature axis).
#pragma once
#include "common.cuh"

namespace tib {
namespace lay {

// row_et[r] = edge type of (dst,src)-ordered row r; node_local[i] = index of node i inside its molecule;
// row_pair[r] = index of the UNDIRECTED pair {src, dst} of row r (pairs (lo, hi), lo < hi, lexicographic per molecule, molecule
// m's pairs start at edge_ptr[m] / 2) and pair_dist[pair] = |x_src - x_dst| - the w MLP depends on an edge through its distance
// only (cpainn.py:283), and d(i,j) == d(j,i) bit for bit, so it is evaluated once per pair.
__global__ void k_row_aux(const uint4* __restrict__ rowa, long long n_edges, int* __restrict__ row_et, const int* __restrict__ mol_ptr,
                          const long long* __restrict__ edge_ptr, int n_mol, int* __restrict__ node_local,
                          int* __restrict__ row_pair, float* __restrict__ pair_dist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_mol) {
    const int n0 = mol_ptr[i], n1 = mol_ptr[i + 1];
    for (int j = n0; j < n1; ++j) node_local[j] = j - n0;
  }
  if (i < n_edges) {
    const uint4 ra = rowa[i];
    row_et[i] = (int)((ra.z >> 16) & 0xFFu);
    // molecule of this row: binary search over edge_ptr
    int lo = 0, hi = n_mol - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (edge_ptr[mid] <= i) lo = mid; else hi = mid - 1; }
    const int n0 = mol_ptr[lo], n = mol_ptr[lo + 1] - n0;
    const int il = (int)ra.x - n0, jl = (int)ra.y - n0;
    const int a = min(il, jl), b = max(il, jl);
    const int pr = (int)(edge_ptr[lo] / 2) + a * n - a * (a + 1) / 2 + (b - a - 1);
    row_pair[i] = pr;
    if (il > jl) pair_dist[pr] = __uint_as_float(ra.w);
  }
}

struct CombineP {
  int n_nodes, F;
  const int* node_in_ptr;
  const uint4* rowa;
  const float4* rowb;
  const float* phi3;        // [E][5F]
  const float* w3;          // [E/2][5F], one row per undirected pair
  const int* row_pair;      // [E]
  const float* s_old;       // [N][F]
  const float* v_old;       // [N][3][F]
  float* s_new;
  float* v_new;
  float* e;                 // [E][F] in place
  const float* edge_emb;    // first layer: e0 = edge_emb[type]
  int first_layer;
};

// SE3Message.forward after the two MLPs (cpainn.py:285-310): m = phi3 * w3, gated scatter over the incoming edges of
// one destination node per block, e += de.
__global__ void __launch_bounds__(128) k_combine(CombineP p) {
  const int F = p.F;
  for (int j = blockIdx.x; j < p.n_nodes; j += gridDim.x) {
    const int r0 = __ldg(p.node_in_ptr + j), r1 = __ldg(p.node_in_ptr + j + 1);
    for (int f = threadIdx.x; f < F; f += 128) {
      float as = 0.0f, av0 = 0.0f, av1 = 0.0f, av2 = 0.0f, d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
      for (int r = r0; r < r1; ++r) {
        const uint4 ra = __ldg(p.rowa + r);
        const float4 rb = __ldg(p.rowb + r);
        const float* ph = p.phi3 + (size_t)r * 5 * F + f;
        const float* wh = p.w3 + (size_t)__ldg(p.row_pair + r) * 5 * F + f;
        const float m0 = __fmul_rn(ph[0], wh[0]), m1 = __fmul_rn(ph[F], wh[F]), m2 = __fmul_rn(ph[2 * F], wh[2 * F]);
        const float m3 = __fmul_rn(ph[3 * F], wh[3 * F]), m4 = __fmul_rn(ph[4 * F], wh[4 * F]);
        if (!p.first_layer) {
          const float* vi = p.v_old + (size_t)ra.x * 3 * F + f;
          av0 = fmaf(m0, __ldg(vi), av0); av1 = fmaf(m0, __ldg(vi + F), av1); av2 = fmaf(m0, __ldg(vi + 2 * F), av2);
          d0 = fmaf(m4, rb.x, d0); d1 = fmaf(m4, rb.y, d1); d2 = fmaf(m4, rb.z, d2);
        }
        av0 = fmaf(m1, rb.x, av0); av1 = fmaf(m1, rb.y, av1); av2 = fmaf(m1, rb.z, av2);
        as += m2;
        float* ep = p.e + (size_t)r * F + f;
        const float e0 = p.first_layer ? __ldg(p.edge_emb + (size_t)((ra.z >> 16) & 0xFFu) * F + f) : *ep;
        *ep = e0 + m3;
      }
      const size_t o = (size_t)j * 3 * F + f;
      float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
      if (!p.first_layer) {
        v0 = __ldg(p.v_old + o); v1 = __ldg(p.v_old + o + F); v2 = __ldg(p.v_old + o + 2 * F);
        av0 += __fmul_rn(d1, v2) - __fmul_rn(d2, v1);          // (sum g dir) x v[dst]   (cpainn.py:296-300)
        av1 += __fmul_rn(d2, v0) - __fmul_rn(d0, v2);
        av2 += __fmul_rn(d0, v1) - __fmul_rn(d1, v0);
      }
      p.s_new[(size_t)j * F + f] = __ldg(p.s_old + (size_t)j * F + f) + as;
      p.v_new[o] = v0 + av0; p.v_new[o + F] = v1 + av1; p.v_new[o + 2 * F] = v2 + av2;
    }
  }
}

struct CombineJvpP {
  CombineP c;               // primal arrays (s_new / v_new / e are NOT written here; s_old / v_old are the layer inputs)
  const float* x;           // [N][3]
  const int* node_local;    // [N]
  const float* w3d;         // [E/2][5F]: d w3 / d dist per undirected pair
  const float* phi3d;       // [D][E][5F] tangents of phi3 (null in the first layer: s0, e0 do not depend on x)
  long long st_phi;         // floats between directions of phi3d
  const float* ts_old; const float* tv_old; float* ts_new; float* tv_new; float* te;
  long long st_s, st_v, st_e;
  int dir0, n_dirs;         // directions dir0 .. dir0 + n_dirs - 1; tangent arrays are indexed by the ABSOLUTE direction,
                            // phi3d by (direction - dir0)
};

// Tangent of k_combine.  One block iteration = one destination node and the THREE coordinate directions of one atom a
// (dir0 and n_dirs are multiples of 3): the primal products m = phi3 * w3 and the gather of v[src] are shared by the
// three directions, and the geometry tangents exist only on the edges that touch atom a (sign = +1 source, -1
// destination):  d_dot_c = sign r_c / d,  dir_dot_c = sign e_c / (1 + d) - r d_dot_c / (1 + d)^2   (graph.py:27-29).
__global__ void __launch_bounds__(128, 4) k_combine_jvp(CombineJvpP pp) {
  const CombineP& p = pp.c;
  const int F = p.F;
  // every array but te / ts_new / tv_new is read-only here: non-coherent loads (__ldg) may be hoisted above the te stores,
  // which is what keeps enough loads in flight (the plain-pointer version ran at 2 TB/s)
  const float* __restrict__ phi3 = p.phi3; const float* __restrict__ w3 = p.w3; const float* __restrict__ w3d = pp.w3d;
  const float* __restrict__ phi3d = pp.phi3d; const float* __restrict__ v_old = p.v_old; const float* __restrict__ tv_old = pp.tv_old;
  float* __restrict__ te_base = pp.te;
  const long long n_items = (long long)p.n_nodes * (pp.n_dirs / 3);
  for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
    // the atoms of one destination node run next to each other (the primal rows they share are then L2 hits)
    const int na = pp.n_dirs / 3, j = (int)(it / na), ac = (int)(it - (long long)j * na);
    const int q0 = pp.dir0 + 3 * ac, a = q0 / 3;
    const int r0 = __ldg(p.node_in_ptr + j), r1 = __ldg(p.node_in_ptr + j + 1);
    const int jl = __ldg(pp.node_local + j);
    const float xj0 = __ldg(pp.x + 3 * (size_t)j), xj1 = __ldg(pp.x + 3 * (size_t)j + 1), xj2 = __ldg(pp.x + 3 * (size_t)j + 2);
    for (int f = threadIdx.x; f < F; f += 128) {
      float as[3] = {0.f, 0.f, 0.f}, av[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
      float ec[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};     // per direction: sum (m4_dot dir + m4 dir_dot)
      float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;                                    // sum m4 dir
      // per-direction base pointers once per item: the 64-bit address arithmetic of ~40 loads per edge was most of the
      // instruction stream
      const float* pd_c[3]; const float* tv_c[3]; float* te_c[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        pd_c[c] = phi3d ? phi3d + (size_t)(3 * ac + c) * pp.st_phi + f : nullptr;
        tv_c[c] = tv_old + (size_t)(q0 + c) * pp.st_v + f;
        te_c[c] = te_base + (size_t)(q0 + c) * pp.st_e + f;
      }
#pragma unroll 2
      for (int r = r0; r < r1; ++r) {
        const uint4 ra = __ldg(p.rowa + r);
        const float4 rb = __ldg(p.rowb + r);
        const int i = (int)ra.x;
        const float dir[3] = {rb.x, rb.y, rb.z};
        const int sg = (__ldg(pp.node_local + i) == a ? 1 : 0) - (jl == a ? 1 : 0);
        const size_t pr = (size_t)__ldg(p.row_pair + r);
        const float* ph = phi3 + (size_t)r * 5 * F + f;
        const float* wh = w3 + pr * 5 * F + f;
        // ---- all loads of this edge first
        float phv[5], w[5], pdv[3][5], vi[3] = {0.f, 0.f, 0.f}, tvi[3][3], teo[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 5; ++k) { phv[k] = __ldg(ph + (size_t)k * F); w[k] = __ldg(wh + (size_t)k * F); }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int k = 0; k < 5; ++k)
            pdv[c][k] = phi3d ? __ldg(pd_c[c] + (size_t)r * 5 * F + (size_t)k * F) : 0.0f;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            tvi[c][k] = p.first_layer ? 0.0f : __ldg(tv_c[c] + (size_t)i * 3 * F + k * F);
          if (!p.first_layer) teo[c] = te_c[c][(size_t)r * F];
        }
        if (!p.first_layer) {
          const float* vp = v_old + (size_t)i * 3 * F + f;
          vi[0] = __ldg(vp); vi[1] = __ldg(vp + F); vi[2] = __ldg(vp + 2 * F);
        }
        float m[5], mw[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 5; ++k) m[k] = phv[k] * w[k];
        float dd[3] = {0.f, 0.f, 0.f}, gd[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        if (sg != 0) {                                          // block-uniform
          const float dist = __uint_as_float(ra.w), fs = (float)sg;
          const float rv[3] = {__ldg(pp.x + 3 * (size_t)i) - xj0, __ldg(pp.x + 3 * (size_t)i + 1) - xj1, __ldg(pp.x + 3 * (size_t)i + 2) - xj2};
          const float inv = 1.0f / (1.0f + dist);
          const float* wd = w3d + pr * 5 * F + f;
#pragma unroll
          for (int k = 0; k < 5; ++k) mw[k] = phv[k] * __ldg(wd + (size_t)k * F);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            dd[c] = dist > 0.0f ? fs * rv[c] / dist : 0.0f;
            const float kk = dd[c] * inv * inv;
#pragma unroll
            for (int k = 0; k < 3; ++k) gd[c][k] = -rv[k] * kk + (k == c ? fs * inv : 0.0f);
          }
        }
        if (!p.first_layer) { d0 = fmaf(m[4], dir[0], d0); d1 = fmaf(m[4], dir[1], d1); d2 = fmaf(m[4], dir[2], d2); }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float md[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) md[k] = fmaf(pdv[c][k], w[k], mw[k] * dd[c]);
          if (!p.first_layer) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              av[c][k] += md[0] * vi[k] + m[0] * tvi[c][k];
              ec[c][k] += md[4] * dir[k] + m[4] * gd[c][k];
            }
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) av[c][k] += md[1] * dir[k] + m[1] * gd[c][k];
          as[c] += md[2];
          te_c[c][(size_t)r * F] = teo[c] + md[3];
        }
      }
      const size_t o = (size_t)j * 3 * F + f;
      float v[3] = {0.f, 0.f, 0.f};
      if (!p.first_layer) { v[0] = __ldg(v_old + o); v[1] = __ldg(v_old + o + F); v[2] = __ldg(v_old + o + 2 * F); }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int q = q0 + c;
        float t[3] = {0.f, 0.f, 0.f}, ts = 0.0f;
        if (!p.first_layer) {
          const float* tv_o = tv_old + (size_t)q * pp.st_v + o;
          t[0] = __ldg(tv_o); t[1] = __ldg(tv_o + F); t[2] = __ldg(tv_o + 2 * F);
          ts = __ldg(pp.ts_old + (size_t)q * pp.st_s + (size_t)j * F + f);
          av[c][0] += (ec[c][1] * v[2] - ec[c][2] * v[1]) + (d1 * t[2] - d2 * t[1]);
          av[c][1] += (ec[c][2] * v[0] - ec[c][0] * v[2]) + (d2 * t[0] - d0 * t[2]);
          av[c][2] += (ec[c][0] * v[1] - ec[c][1] * v[0]) + (d0 * t[1] - d1 * t[0]);
        }
        pp.ts_new[(size_t)q * pp.st_s + (size_t)j * F + f] = ts + as[c];
        float* tv_n = pp.tv_new + (size_t)q * pp.st_v + o;
        tv_n[0] = t[0] + av[c][0]; tv_n[F] = t[1] + av[c][1]; tv_n[2 * F] = t[2] + av[c][2];
      }
    }
  }
}

// ---- Update (cpainn.py:345-376), element-wise parts; the GEMMs run in k_chain_tc --------------------------------------------
// vvuv [3N][2F]: columns [0,F) = V v, [F,2F) = U v per (node, xyz) row.  q = |V v| over xyz  -> qs[N][F]
struct UpdQP {
  int n_nodes, F;
  const float* vvuv;        // primal [3N][2F]
  float* q;                 // [N][F]
  // tangents (n_dirs = 0: primal only): tvvuv [D][3N][2F], tq [D][N][F]: q_dot = (V v . V v_dot) / q
  int n_dirs;
  const float* tvvuv; long long st_tvvuv;
  float* tq; long long st_tq;
};
__global__ void __launch_bounds__(128) k_upd_q(UpdQP p) {
  const int F = p.F;
  for (int j = blockIdx.x; j < p.n_nodes; j += gridDim.x)
    for (int f = threadIdx.x; f < F; f += 128) {
      const float* vv = p.vvuv + (size_t)j * 3 * 2 * F + f;
      const float a0 = vv[0], a1 = vv[2 * F], a2 = vv[4 * F];
      const float q = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2)));   // cpainn.py:361
      p.q[(size_t)j * F + f] = q;
      for (int d = 0; d < p.n_dirs; ++d) {
        const float* tv = p.tvvuv + (size_t)d * p.st_tvvuv + (size_t)j * 3 * 2 * F + f;
        const float dot = a0 * tv[0] + a1 * tv[2 * F] + a2 * tv[4 * F];
        p.tq[(size_t)d * p.st_tq + (size_t)j * F + f] = q > 0.0f ? dot / q : 0.0f;
      }
    }
}

// v += (U v) g ; s += q^2 a + c   with (g, a, c) = split(gac [N][3F])     (cpainn.py:366-374); tangents first (they need
// the primal values of this layer's inputs), then the primal update in place.
struct UpdApplyP {
  int n_nodes, F;
  const float* vvuv; const float* q; const float* gac;
  float* s; float* v;
  int n_dirs;
  const float* tvvuv; long long st_tvvuv;
  const float* tq; long long st_tq;
  const float* tgac; long long st_tgac;
  float* ts; long long st_s;
  float* tv; long long st_v;
};
__global__ void __launch_bounds__(128) k_upd_apply(UpdApplyP p) {
  const int F = p.F;
  for (int j = blockIdx.x; j < p.n_nodes; j += gridDim.x)
    for (int f = threadIdx.x; f < F; f += 128) {
      const float* gac = p.gac + (size_t)j * 3 * F + f;
      const float g = gac[0], a = gac[F], c = gac[2 * F];
      const float q = p.q[(size_t)j * F + f];
      const float* uv = p.vvuv + (size_t)j * 3 * 2 * F + F + f;
      const float u0 = uv[0], u1 = uv[2 * F], u2 = uv[4 * F];
      for (int d = 0; d < p.n_dirs; ++d) {
        const float* tg = p.tgac + (size_t)d * p.st_tgac + (size_t)j * 3 * F + f;
        const float gd = tg[0], ad = tg[F], cd = tg[2 * F];
        const float* tu = p.tvvuv + (size_t)d * p.st_tvvuv + (size_t)j * 3 * 2 * F + F + f;
        float* tv = p.tv + (size_t)d * p.st_v + (size_t)j * 3 * F + f;
        tv[0] += tu[0] * g + u0 * gd; tv[F] += tu[2 * F] * g + u1 * gd; tv[2 * F] += tu[4 * F] * g + u2 * gd;
        const float qd = p.tq[(size_t)d * p.st_tq + (size_t)j * F + f];
        p.ts[(size_t)d * p.st_s + (size_t)j * F + f] += 2.0f * q * qd * a + q * q * ad + cd;
      }
      float* v = p.v + (size_t)j * 3 * F + f;
      v[0] = __fadd_rn(v[0], __fmul_rn(u0, g)); v[F] = __fadd_rn(v[F], __fmul_rn(u1, g)); v[2 * F] = __fadd_rn(v[2 * F], __fmul_rn(u2, g));
      float* s = p.s + (size_t)j * F + f;
      *s = __fadd_rn(*s, __fadd_rn(__fmul_rn(__fmul_rn(q, q), a), c));                                // cpainn.py:371,373
    }
}

}  // namespace lay
}  // namespace tib
