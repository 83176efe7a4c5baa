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

// row_et[r] = edge type of (dst,src)-ordered row r; node_local[i] = index of node i inside its molecule
__global__ void k_row_aux(const uint4* __restrict__ rowa, long long n_edges, int* __restrict__ row_et, const int* __restrict__ mol_ptr,
                          int n_mol, int* __restrict__ node_local) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_edges) row_et[i] = (int)((rowa[i].z >> 16) & 0xFFu);
  if (i < n_mol) {
    const int n0 = mol_ptr[i], n1 = mol_ptr[i + 1];
    for (int j = n0; j < n1; ++j) node_local[j] = j - n0;
  }
}

struct CombineP {
  int n_nodes, F;
  const int* node_in_ptr;
  const uint4* rowa;
  const float4* rowb;
  const float* phi3;        // [E][5F]
  const float* w3;          // [E][5F]
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
        const float* wh = p.w3 + (size_t)r * 5 * F + f;
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
  const float* w3d;         // [E][5F]: d w3 / d dist
  const float* phi3d;       // [D][E][5F] tangents of phi3 (null in the first layer: s0, e0 do not depend on x)
  long long st_phi;         // floats between directions of phi3d
  const float* ts_old; const float* tv_old; float* ts_new; float* tv_new; float* te;
  long long st_s, st_v, st_e;
  int dir0, n_dirs;         // directions dir0 .. dir0 + n_dirs - 1; tangent arrays are indexed by the ABSOLUTE direction,
                            // phi3d by (direction - dir0)
};

// Tangent of k_combine for one (destination node, direction) per block iteration.
__global__ void __launch_bounds__(128) k_combine_jvp(CombineJvpP pp) {
  const CombineP& p = pp.c;
  const int F = p.F;
  const long long n_items = (long long)p.n_nodes * pp.n_dirs;
  for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int dl = (int)(it / p.n_nodes), j = (int)(it - (long long)dl * p.n_nodes);
    const int q = pp.dir0 + dl, a = q / 3, c = q % 3;
    const int r0 = __ldg(p.node_in_ptr + j), r1 = __ldg(p.node_in_ptr + j + 1);
    const int jl = __ldg(pp.node_local + j);
    const float xj0 = __ldg(pp.x + 3 * (size_t)j), xj1 = __ldg(pp.x + 3 * (size_t)j + 1), xj2 = __ldg(pp.x + 3 * (size_t)j + 2);
    const float* ts_o = pp.ts_old + (size_t)q * pp.st_s;
    const float* tv_o = pp.tv_old + (size_t)q * pp.st_v;
    float* te = pp.te + (size_t)q * pp.st_e;
    for (int f = threadIdx.x; f < F; f += 128) {
      float as = 0.0f, av0 = 0.0f, av1 = 0.0f, av2 = 0.0f;
      float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, e0 = 0.0f, e1 = 0.0f, e2 = 0.0f;   // sum m4 dir, sum (m4d dir + m4 dird)
      for (int r = r0; r < r1; ++r) {
        const uint4 ra = __ldg(p.rowa + r);
        const float4 rb = __ldg(p.rowb + r);
        const int i = (int)ra.x;
        const float dist = __uint_as_float(ra.w);
        // geometry tangents (graph.py:27-29): r = x_i - x_j, d = |r|, dir = r / (1 + d)
        const int sg = (__ldg(pp.node_local + i) == a ? 1 : 0) - (jl == a ? 1 : 0);
        float dd = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
        if (sg != 0) {
          const float rx = __ldg(pp.x + 3 * (size_t)i) - xj0, ry = __ldg(pp.x + 3 * (size_t)i + 1) - xj1, rz = __ldg(pp.x + 3 * (size_t)i + 2) - xj2;
          const float rc = c == 0 ? rx : (c == 1 ? ry : rz);
          const float inv = 1.0f / (1.0f + dist);
          dd = dist > 0.0f ? (float)sg * rc / dist : 0.0f;
          const float k = dd * inv * inv;
          g0 = -rx * k; g1 = -ry * k; g2 = -rz * k;
          if (c == 0) g0 += (float)sg * inv; else if (c == 1) g1 += (float)sg * inv; else g2 += (float)sg * inv;
        }
        const float* ph = p.phi3 + (size_t)r * 5 * F + f;
        const float* wh = p.w3 + (size_t)r * 5 * F + f;
        const float* wd = pp.w3d + (size_t)r * 5 * F + f;
        const float* pd = pp.phi3d ? pp.phi3d + (size_t)dl * pp.st_phi + (size_t)r * 5 * F + f : nullptr;
        float m[5], md[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float phk = ph[(size_t)k * F], whk = wh[(size_t)k * F];
          m[k] = phk * whk;
          md[k] = phk * (wd[(size_t)k * F] * dd);
          if (pd) md[k] = fmaf(pd[(size_t)k * F], whk, md[k]);
        }
        if (!p.first_layer) {
          const float* vi = p.v_old + (size_t)i * 3 * F + f;
          const float* tvi = tv_o + (size_t)i * 3 * F + f;
          av0 += md[0] * __ldg(vi) + m[0] * __ldg(tvi);
          av1 += md[0] * __ldg(vi + F) + m[0] * __ldg(tvi + F);
          av2 += md[0] * __ldg(vi + 2 * F) + m[0] * __ldg(tvi + 2 * F);
          d0 = fmaf(m[4], rb.x, d0); d1 = fmaf(m[4], rb.y, d1); d2 = fmaf(m[4], rb.z, d2);
          e0 += md[4] * rb.x + m[4] * g0; e1 += md[4] * rb.y + m[4] * g1; e2 += md[4] * rb.z + m[4] * g2;
        }
        av0 += md[1] * rb.x + m[1] * g0; av1 += md[1] * rb.y + m[1] * g1; av2 += md[1] * rb.z + m[1] * g2;
        as += md[2];
        float* ep = te + (size_t)r * F + f;
        *ep = (p.first_layer ? 0.0f : *ep) + md[3];
      }
      const size_t o = (size_t)j * 3 * F + f;
      float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
      if (!p.first_layer) {
        const float v0 = __ldg(p.v_old + o), v1 = __ldg(p.v_old + o + F), v2 = __ldg(p.v_old + o + 2 * F);
        t0 = __ldg(tv_o + o); t1 = __ldg(tv_o + o + F); t2 = __ldg(tv_o + o + 2 * F);
        av0 += (e1 * v2 - e2 * v1) + (d1 * t2 - d2 * t1);
        av1 += (e2 * v0 - e0 * v2) + (d2 * t0 - d0 * t2);
        av2 += (e0 * v1 - e1 * v0) + (d0 * t1 - d1 * t0);
      }
      float* ts_n = pp.ts_new + (size_t)q * pp.st_s;
      float* tv_n = pp.tv_new + (size_t)q * pp.st_v;
      ts_n[(size_t)j * F + f] = (p.first_layer ? 0.0f : __ldg(ts_o + (size_t)j * F + f)) + as;
      tv_n[o] = t0 + av0; tv_n[o + F] = t1 + av1; tv_n[o + 2 * F] = t2 + av2;
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
