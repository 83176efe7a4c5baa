// tc_message.cuh - SE3Message.forward (cpainn.py:263-310) fused on the sm_100a tensor cores.
//
// Work unit: a TILE = up to kTileNodes consecutive destination nodes and all their incoming edges
// (<= 128 rows).  Edge features e[E][F] are kept in (dst, src)-lexicographic order inside the
// library, so the rows of a tile are contiguous in HBM and every destination node is owned by
// exactly one tile: the scatter-sum over incoming edges needs neither atomics nor a second pass.
//
// Per tile (F = 128):
//   hidden layers   D[edge][feat] = A[edge][k] * W[feat][k]^T      (edge = TMEM lane: LayerNorm is thread-local)
//       w   : PE(d) -> LN/SiLU -> LN/SiLU                          (2 GEMMs)
//       phi : cat[s[src], e] -> LN/SiLU -> LN/SiLU                 (3 GEMMs: the 2F input in two K halves)
//   output layer    D^T[feat][edge] = W3[feat][k] * H2[edge][k]^T  (feat = TMEM lane: the gated
//       scatter over the edges of a tile is a thread-local loop over TMEM columns), one 128-feature
//       split (gates, scale_edge_dir, ds, de, cross_gates) at a time, phi and w side by side,
//       double-buffered in TMEM so the MMAs of split i+1 overlap the scatter of split i.
// Weights stream from L2 through a 4-stage ring of 16 KB chunks with cp.async.bulk (TMA unit).
//
// Warp roles (192 threads): warps 0-3 operand builders / epilogue (thread = TMEM lane),
// warp 4 weight producer (+ TMEM allocation), warp 5 MMA issuer.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace tib {
namespace tc {

constexpr int kF = 128;
constexpr int kThreads = 192;
constexpr int kStages = 4;
constexpr int kTileNodes = 16;                 // destination nodes per tile (delta-v window in smem)
constexpr int kChunksPerLayer = 60;

// streamed chunk order of one message layer (every entry is 4 chunks = one [128 x 128] matrix):
//   0 w.W1 | 4 phi.W1[:, :F] | 8 w.W2 | 12 phi.W1[:, F:] | 16 phi.W2 | 20+8sp phi.W3[sp] | 24+8sp w.W3[sp]
struct MsgParams {
  const float *phi_b1, *phi_g1, *phi_be1, *phi_b2, *phi_g2, *phi_be2, *phi_b3;
  const float *w_b1, *w_g1, *w_be1, *w_b2, *w_g2, *w_be2, *w_b3;
};

struct TcMsgP {
  int n_nodes, n_tiles, nodes_per_tile;
  const int* node_in_ptr;   // [N+1] first (dst-major) edge row of each node
  const int* node_mol;      // [N]
  const int* mol_ptr;       // [B+1]
  const float* x;           // [N][3]
  const float* s_old;       // [N][F]
  const float* v_old;       // [N][3][F]
  float* s_new;
  float* v_new;
  float* e;                 // [E][F] (dst,src) order, updated in place
  const unsigned char* wblob;   // kChunksPerLayer chunks of this layer
  MsgParams prm;
  float length_scale;
  int first_layer;
  int passes;               // 3 = split-f16 (fp32-faithful), 1 = single f16 pass
  int* err;
};

struct RowInfo { int src; int slot; float dx, dy, dz; int last; int dst; int pad; };   // 32 B

struct MsgSmem {
  // offsets (bytes) into dynamic shared memory
  static constexpr uint32_t X = 0;
  static constexpr uint32_t Y = X + kOperandBytes;
  static constexpr uint32_t RING = Y + kOperandBytes;
  static constexpr uint32_t DV = RING + kStages * kChunkBytes;                 // [kTileNodes][3][F] fp32
  static constexpr uint32_t ROWS = DV + kTileNodes * 3 * kF * 4;               // RowInfo[128]
  static constexpr uint32_t BARS = ROWS + 128 * sizeof(RowInfo);
  static constexpr uint32_t TOTAL = BARS + 256;
};
// barrier indices
enum { B_FULL = 0, B_EMPTY = B_FULL + kStages, B_XFULL = B_EMPTY + kStages, B_YFULL, B_YFREE, B_ACC0, B_ACC1,
       B_TFULL0, B_TFULL1, B_TEMPTY0, B_TEMPTY1, B_COUNT };

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// rows [32*warp, +32) x 16 column groups of an operand image from row-major fp32 global rows:
// lane = (row & 7, group quad) so each 128 B line of a row is read by 4 lanes and every store
// instruction writes 4 x 128 contiguous bytes.
template <typename RowPtr>
__device__ __forceinline__ void build_from_global(unsigned char* op, int warp, int lane, int rows, RowPtr row_ptr) {
#pragma unroll 1
  for (int oct = 0; oct < 4; ++oct) {
    const int r = 32 * warp + 8 * oct + (lane & 7);
    const float* src = r < rows ? row_ptr(r) : nullptr;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const int g = 4 * kq + (lane >> 3);
      float v[8];
      if (src) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + g * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + g * 8 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.0f;
      }
      store_group(op, kOperandHalfBytes, r, g, v);
    }
  }
}

// accumulator row (this thread's TMEM lane, 128 columns) -> + bias -> LayerNorm -> SiLU -> operand image
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr, const float* __restrict__ b, const float* __restrict__ g,
                                                const float* __restrict__ be, unsigned char* op, int row) {
  float v[128];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float t[32];
    tmem_ld32(taddr + 32 * c, t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[32 * c + i] = t[i];
  }
  float sum = 0.0f;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + c);
    v[4 * c + 0] += bb.x; v[4 * c + 1] += bb.y; v[4 * c + 2] += bb.z; v[4 * c + 3] += bb.w;
    sum += (v[4 * c + 0] + v[4 * c + 1]) + (v[4 * c + 2] + v[4 * c + 3]);
  }
  const float mean = sum * (1.0f / 128.0f);
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    v[i] -= mean;
    ss = fmaf(v[i], v[i], ss);
  }
  const float rstd = rsqrtf(ss * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
  for (int kg = 0; kg < 16; ++kg) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + 2 * kg), g1 = __ldg(reinterpret_cast<const float4*>(g) + 2 * kg + 1);
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(be) + 2 * kg), e1 = __ldg(reinterpret_cast<const float4*>(be) + 2 * kg + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float z = fmaf(v[8 * kg + i] * rstd, gg[i], ee[i]);
      y[i] = __fdividef(z, 1.0f + __expf(-z));
    }
    store_group(op, kOperandHalfBytes, row, kg, y);
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_message_tc(TcMsgP p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const X = smem + MsgSmem::X;
  unsigned char* const Y = smem + MsgSmem::Y;
  unsigned char* const RING = smem + MsgSmem::RING;
  float* const DV = reinterpret_cast<float*>(smem + MsgSmem::DV);
  RowInfo* const ROWS = reinterpret_cast<RowInfo*>(smem + MsgSmem::ROWS);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + MsgSmem::BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + MsgSmem::BARS + 8 * B_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
    mbar_init(&bars[B_XFULL], 128); mbar_init(&bars[B_YFULL], 128);
    mbar_init(&bars[B_YFREE], 1); mbar_init(&bars[B_ACC0], 1); mbar_init(&bars[B_ACC1], 1);
    mbar_init(&bars[B_TFULL0], 1); mbar_init(&bars[B_TFULL1], 1);
    mbar_init(&bars[B_TEMPTY0], 128); mbar_init(&bars[B_TEMPTY1], 128);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_splits = p.first_layer ? 3 : 5;

  if (warp == 4) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < kChunksPerLayer; ++c) {
          if (p.first_layer && ((c >= 20 && c < 28) || c >= 52)) continue;   // splits 0 and 4 multiply v = 0
          mbar_wait(&bars[B_EMPTY + stage], ph ^ 1, err);
          mbar_arrive_expect_tx(&bars[B_FULL + stage], kChunkBytes);
          bulk_g2s(RING + stage * kChunkBytes, p.wblob + (size_t)c * kChunkBytes, kChunkBytes, &bars[B_FULL + stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      uint32_t px = 0, py = 0, pte[2] = {0, 0};
      const uint32_t xa = smem_u32(X), ya = smem_u32(Y), ring = smem_u32(RING);
      // one [128 x 128] matrix = 4 chunks.  transposed = weights are the A operand.
      auto gemm = [&](uint32_t d, uint32_t op, bool transposed, bool accumulate) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&bars[B_FULL + stage], ph, err);
          tc_fence_after();
          const uint32_t wst = ring + stage * kChunkBytes, opk = op + kb * (2 * kKStepBytes);
          if (!transposed) mma_f16x3(d, opk, kOperandHalfBytes, wst, kChunkHalfBytes, 2, accumulate || kb > 0, p.passes);
          else             mma_f16x3(d, wst, kChunkHalfBytes, opk, kOperandHalfBytes, 2, accumulate || kb > 0, p.passes);
          tc_commit(&bars[B_EMPTY + stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
      };
      const uint32_t acc0 = tmem, acc1 = tmem + 128;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(&bars[B_XFULL], px, err); px ^= 1; tc_fence_after();
        gemm(acc0, xa, false, false);                       // w layer 1   : PE(d)
        tc_commit(&bars[B_ACC0]);
        mbar_wait(&bars[B_YFULL], py, err); py ^= 1; tc_fence_after();
        gemm(acc1, ya, false, false);                       // phi layer 1 : s[src] half
        tc_commit(&bars[B_YFREE]);
        mbar_wait(&bars[B_XFULL], px, err); px ^= 1; tc_fence_after();
        gemm(acc0, xa, false, false);                       // w layer 2
        tc_commit(&bars[B_ACC0]);
        mbar_wait(&bars[B_YFULL], py, err); py ^= 1; tc_fence_after();
        gemm(acc1, ya, false, true);                        // phi layer 1 : e half (accumulates)
        tc_commit(&bars[B_ACC1]);
        mbar_wait(&bars[B_YFULL], py, err); py ^= 1; tc_fence_after();
        gemm(acc1, ya, false, false);                       // phi layer 2
        tc_commit(&bars[B_ACC1]);
        mbar_wait(&bars[B_YFULL], py, err); py ^= 1;
        mbar_wait(&bars[B_XFULL], px, err); px ^= 1; tc_fence_after();
        for (int it = 0; it < n_splits; ++it) {
          const int pb = it & 1;
          mbar_wait(&bars[B_TEMPTY0 + pb], pte[pb] ^ 1, err); pte[pb] ^= 1; tc_fence_after();
          gemm(tmem + 256 * pb, ya, true, false);           // phi layer 3, one split (transposed)
          gemm(tmem + 256 * pb + 128, xa, true, false);     // w layer 3, same split
          tc_commit(&bars[B_TFULL0 + pb]);
        }
      }
    }
  } else {
    // =========================== builders / epilogue (128 threads) ===========================
    uint32_t pa0 = 0, pa1 = 0, pyf = 0, ptf[2] = {0, 0};
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const int f = tid;                                      // feature owned in the transposed epilogue
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int node_lo = tile * p.nodes_per_tile;
      const int node_hi = min(node_lo + p.nodes_per_tile, p.n_nodes);
      const int row0 = __ldg(p.node_in_ptr + node_lo);
      const int rows = __ldg(p.node_in_ptr + node_hi) - row0;
      epi_bar_sync();                                       // previous tile's readers of ROWS are done
      float dist = 0.0f;
      {
        RowInfo ri; ri.src = 0; ri.slot = 0; ri.dx = ri.dy = ri.dz = 0.0f; ri.last = 0; ri.dst = node_lo; ri.pad = 0;
        if (tid < rows) {
          const int er = row0 + tid;
          int j = node_lo;
          while (j + 1 < node_hi && __ldg(p.node_in_ptr + j + 1) <= er) ++j;
          const int mol = __ldg(p.node_mol + j);
          const int n0 = __ldg(p.mol_ptr + mol), n = __ldg(p.mol_ptr + mol + 1) - n0;
          const int jl = j - n0, ip = er - __ldg(p.node_in_ptr + j);
          const int il = ip + (ip >= jl);
          const int src = n0 + il;
          // r = x[src] - x[dst], d = |r|, dir = r / (1 + d)                       (graph.py:27-29)
          const float rx = __ldg(p.x + 3 * src + 0) - __ldg(p.x + 3 * j + 0);
          const float ry = __ldg(p.x + 3 * src + 1) - __ldg(p.x + 3 * j + 1);
          const float rz = __ldg(p.x + 3 * src + 2) - __ldg(p.x + 3 * j + 2);
          dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
          const float den = 1.0f + dist;
          ri.src = src; ri.slot = j - node_lo; ri.dst = j; ri.last = (ip == n - 2);
          ri.dx = __fdiv_rn(rx, den); ri.dy = __fdiv_rn(ry, den); ri.dz = __fdiv_rn(rz, den);
        }
        ROWS[tid] = ri;
      }
#pragma unroll
      for (int i = 0; i < kTileNodes * 3; ++i) DV[i * kF + f] = 0.0f;
      epi_bar_sync();

      // ---- E1: PositionalEncoder(edge_dist) -> X                                 (cpainn.py:283)
#pragma unroll 1
      for (int kg = 0; kg < 16; ++kg) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float sn = 0.0f, cs = 0.0f;
          if (tid < rows) sincosf(pe_arg(dist, p.length_scale, 4 * kg + q + 1), &sn, &cs);
          v[2 * q] = cs; v[2 * q + 1] = sn;
        }
        store_group(X, kOperandHalfBytes, tid, kg, v);
      }
      fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
      // ---- E2: s[src] -> Y                                                       (cpainn.py:275-281)
      build_from_global(Y, warp, lane, rows, [&](int r) { return p.s_old + (size_t)ROWS[r].src * kF; });
      fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
      // ---- E3: w hidden 1 -> X
      mbar_wait(&bars[B_ACC0], pa0, err); pa0 ^= 1; tc_fence_after();
      hidden_epilogue(lane_taddr, p.prm.w_b1, p.prm.w_g1, p.prm.w_be1, X, tid);
      tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
      // ---- E4: e rows -> Y (after the s[src] half has been consumed)
      mbar_wait(&bars[B_YFREE], pyf, err); pyf ^= 1;
      build_from_global(Y, warp, lane, rows, [&](int r) { return p.e + (size_t)(row0 + r) * kF; });
      fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
      // ---- E5: w hidden 2 -> X (final: B operand of the output layer)
      mbar_wait(&bars[B_ACC0], pa0, err); pa0 ^= 1; tc_fence_after();
      hidden_epilogue(lane_taddr, p.prm.w_b2, p.prm.w_g2, p.prm.w_be2, X, tid);
      tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
      // ---- E6: phi hidden 1 -> Y
      mbar_wait(&bars[B_ACC1], pa1, err); pa1 ^= 1; tc_fence_after();
      hidden_epilogue(lane_taddr + 128, p.prm.phi_b1, p.prm.phi_g1, p.prm.phi_be1, Y, tid);
      tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
      // ---- E7: phi hidden 2 -> Y (final)
      mbar_wait(&bars[B_ACC1], pa1, err); pa1 ^= 1; tc_fence_after();
      hidden_epilogue(lane_taddr + 128, p.prm.phi_b2, p.prm.phi_g2, p.prm.phi_be2, Y, tid);
      tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);

      // ---- E8..: output layer, transposed: this thread owns feature f, TMEM columns are edges.
      // m = phi3 * w3, split order gates | scale_edge_dir | ds | de | cross_gates   (cpainn.py:285-290)
      for (int it = 0; it < n_splits; ++it) {
        const int sp = p.first_layer ? it + 1 : it;
        const int pb = it & 1;
        const float bphi = __ldg(p.prm.phi_b3 + sp * kF + f), bw = __ldg(p.prm.w_b3 + sp * kF + f);
        mbar_wait(&bars[B_TFULL0 + pb], ptf[pb], err); ptf[pb] ^= 1; tc_fence_after();
        const uint32_t tphi = lane_taddr + 256 * pb, tw = tphi + 128;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        float vj0 = 0.0f, vj1 = 0.0f, vj2 = 0.0f;
        bool have_vj = false;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (32 * c >= rows) break;
          float P[32], Q[32];
          tmem_ld32(tphi + 32 * c, P);
          tmem_ld32(tw + 32 * c, Q);
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int r = 32 * c + q;
            if (r < rows) {
              const float m = __fmul_rn(P[q] + bphi, Q[q] + bw);
              const RowInfo ri = ROWS[r];
              if (sp == 0) {            // gates * v[src]
                const float* vi = p.v_old + (size_t)ri.src * 3 * kF + f;
                a0 = fmaf(m, __ldg(vi), a0); a1 = fmaf(m, __ldg(vi + kF), a1); a2 = fmaf(m, __ldg(vi + 2 * kF), a2);
              } else if (sp == 1) {     // scale_edge_dir * dir
                a0 = fmaf(m, ri.dx, a0); a1 = fmaf(m, ri.dy, a1); a2 = fmaf(m, ri.dz, a2);
              } else if (sp == 2) {     // ds
                a0 += m;
              } else if (sp == 3) {     // e += de                                   (cpainn.py:308)
                float* ep = p.e + (size_t)(row0 + r) * kF + f;
                *ep = *ep + m;
              } else {                  // cross_gates * (dir x v[dst])              (cpainn.py:296-300)
                if (!have_vj) {
                  const float* vj = p.v_old + (size_t)ri.dst * 3 * kF + f;
                  vj0 = __ldg(vj); vj1 = __ldg(vj + kF); vj2 = __ldg(vj + 2 * kF);
                  have_vj = true;
                }
                const float c0 = __fmul_rn(ri.dy, vj2) - __fmul_rn(ri.dz, vj1);
                const float c1 = __fmul_rn(ri.dz, vj0) - __fmul_rn(ri.dx, vj2);
                const float c2 = __fmul_rn(ri.dx, vj1) - __fmul_rn(ri.dy, vj0);
                a0 = fmaf(m, c0, a0); a1 = fmaf(m, c1, a1); a2 = fmaf(m, c2, a2);
              }
              if (ri.last && sp != 3) {   // all incoming edges of this destination node seen
                if (sp == 2) {
                  const size_t o = (size_t)ri.dst * kF + f;
                  p.s_new[o] = __ldg(p.s_old + o) + a0;                               // cpainn.py:306
                } else {
                  float* dv = DV + ri.slot * 3 * kF + f;
                  dv[0] += a0; dv[kF] += a1; dv[2 * kF] += a2;
                }
                a0 = a1 = a2 = 0.0f;
                have_vj = false;
              }
            }
          }
        }
        tc_fence_before(); mbar_arrive(&bars[B_TEMPTY0 + pb]);
      }
      // ---- v_new = v_old + sum of the three equivariant contributions              (cpainn.py:305)
      for (int j = node_lo; j < node_hi; ++j) {
        const float* dv = DV + (j - node_lo) * 3 * kF + f;
        const size_t o = (size_t)j * 3 * kF + f;
#pragma unroll
        for (int xyz = 0; xyz < 3; ++xyz)
          p.v_new[o + xyz * kF] = (p.first_layer ? 0.0f : __ldg(p.v_old + o + xyz * kF)) + dv[xyz * kF];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// -------------------------------------------------------------------------------------------------
// Self test of the tensor-core plumbing (descriptors, operand image, ring, TMEM addressing):
//   transposed = 0:  out[r][c] = sum_k A[r][k] * W[c][k]      (lane = row of A)
//   transposed = 1:  out[r][c] = sum_k W[r][k] * A[c][k]      (lane = row of W)
// A is fp32 [128][128] row-major, wchunks = 4 packed chunks of W [128][128], out fp32 [128][128].
__global__ void __launch_bounds__(kThreads, 1) k_tc_selftest(const float* __restrict__ A, const unsigned char* __restrict__ wchunks,
                                                             float* __restrict__ out, int transposed, int* err_ptr) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const X = smem;
  unsigned char* const RING = smem + kOperandBytes;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kOperandBytes + kStages * kChunkBytes);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = err_ptr;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[4], 128);   // operand full
    mbar_init(&bars[5], 1);     // accumulator full
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 4) {
    if (lane == 0)
      for (int c = 0; c < 4; ++c) {
        mbar_arrive_expect_tx(&bars[c], kChunkBytes);
        bulk_g2s(RING + c * kChunkBytes, wchunks + (size_t)c * kChunkBytes, kChunkBytes, &bars[c]);
      }
  } else if (warp == 5) {
    if (lane == 0) {
      mbar_wait(&bars[4], 0, err);
      tc_fence_after();
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&bars[kb], 0, err);
        tc_fence_after();
        const uint32_t wst = smem_u32(RING) + kb * kChunkBytes, opk = smem_u32(X) + kb * (2 * kKStepBytes);
        if (!transposed) mma_f16x3(tmem, opk, kOperandHalfBytes, wst, kChunkHalfBytes, 2, kb > 0, 3);
        else             mma_f16x3(tmem, wst, kChunkHalfBytes, opk, kOperandHalfBytes, 2, kb > 0, 3);
      }
      tc_commit(&bars[5]);
    }
  } else {
    build_from_global(X, warp, lane, 128, [&](int r) { return A + (size_t)r * 128; });
    fence_proxy_async();
    mbar_arrive(&bars[4]);
    mbar_wait(&bars[5], 0, err);
    tc_fence_after();
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) {
      float t[32];
      tmem_ld32(lane_taddr + 32 * c, t);
      for (int i = 0; i < 32; ++i) out[(size_t)tid * 128 + 32 * c + i] = t[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

// ---- small helpers of the tensor-core drift ---------------------------------------------------------
// node_mol[j], node_in_ptr[j] = first (dst,src)-ordered edge row of node j; node_in_ptr[N] = E
__global__ void k_node_tables(const int* __restrict__ mol_ptr, const long long* __restrict__ edge_ptr, int n_mol,
                              int* __restrict__ node_mol, int* __restrict__ node_in_ptr) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mol) return;
  const int n0 = mol_ptr[m], n = mol_ptr[m + 1] - n0;
  const int e0 = (int)edge_ptr[m];
  for (int j = 0; j < n; ++j) { node_mol[n0 + j] = m; node_in_ptr[n0 + j] = e0 + j * (n - 1); }
  if (m == n_mol - 1) node_in_ptr[n0 + n] = (int)edge_ptr[n_mol];
}

// e0 = Emb4(edge_type) written in (dst,src) order; edge_type is given in the reference's (src,dst) order
__global__ void k_edge_init_dst(const unsigned char* __restrict__ edge_type, const float* __restrict__ edge_emb,
                                const int* __restrict__ mol_ptr, const long long* __restrict__ edge_ptr,
                                float* __restrict__ e, int F) {
  const int m = blockIdx.x;
  const int n = mol_ptr[m + 1] - mol_ptr[m];
  const long long e0 = edge_ptr[m];
  const int ne = n * (n - 1), f4n = F / 4;
  for (int idx = threadIdx.x; idx < ne * f4n; idx += blockDim.x) {
    const int row = idx / f4n, f4 = idx % f4n;
    const int jl = row / (n - 1), ip = row % (n - 1), il = ip + (ip >= jl);
    const int src_major = il * (n - 1) + jl - (jl > il);
    const float4 val = __ldg(reinterpret_cast<const float4*>(edge_emb + (size_t)edge_type[e0 + src_major] * F) + f4);
    reinterpret_cast<float4*>(e + (size_t)(e0 + row) * F)[f4] = val;
  }
}

}  // namespace tc
}  // namespace tib
