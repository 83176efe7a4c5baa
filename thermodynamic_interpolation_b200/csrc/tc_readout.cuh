// tc_readout.cuh - LayerReadout.forward (cpainn.py:425-437, n_features_out = 1) on the sm_100a tensor cores (F = 128).
//
//   (_, gate) = split(MLP_{F->F->F->2}(s));   out[node][xyz] = (Vout . v[node][xyz]) * gate
//
// Work unit: a tile of 128 consecutive nodes, D[node][feat] = A[node][k] * W[feat][k]^T as in tc_update.cuh
// (same warp roles, weight ring and barrier pair).  Only the two hidden layers are GEMMs; the last Linear has two
// output columns of which the reference keeps one, so it is a dot product in the layer-2 epilogue.  The
// 3 x 128-float dot products with Vout are done warp-per-row (coalesced 512-byte reads) while the first GEMM runs.
#pragma once
#include "tc_update.cuh"

namespace tib {
namespace tc {

constexpr int kRoChunks = 8;       // W1 | W2, 4 chunks each

struct TcRoP {
  int n_nodes, n_tiles;
  int tile_nodes;               // nodes per tile (<= 128), see tc_tile_nodes
  const float* s;               // [N][F]
  const float* v;               // [N][3][F]
  float* out;                   // [N][3]
  const unsigned char* wblob;   // 2 matrices x 4 chunks
  const float *b1, *g1, *be1, *b2, *g2, *be2;
  const float* w3;              // row 1 of W3 [2][F] (the gate)
  const float* b3;              // &b3[1]
  const float* vout;            // [F]
  int passes;
  int* err;                     // [0] pipeline time-out, [1] non-finite readout
};

struct RoSmem {
  static constexpr uint32_t X = 0;
  static constexpr uint32_t RING = X + kOperandBytes;
  static constexpr uint32_t PRM = RING + kStages * kChunkBytes;    // b1 g1 be1 b2 g2 be2 | w3 | vout
  static constexpr uint32_t STAT = PRM + 8 * 128 * 4;              // float2 [4 groups][128 rows]
  static constexpr uint32_t VO = STAT + 4 * 128 * 8;               // float [128 rows][4]
  static constexpr uint32_t BARS = VO + 128 * 16;
  static constexpr uint32_t TOTAL = BARS + 256;
};

// layer-2 accumulator row, columns [32*grp, +32): + bias -> LayerNorm -> SiLU -> partial dot with w3; returns the
// full dot product (4-way exchange through `stat`)
__device__ __noinline__ float ro_gate(uint32_t taddr, int grp, int row, const float* b, const float* g, const float* be,
                                      const float* w3, float2* stat) {
  const uint32_t t0 = taddr + 32 * grp;
  float t[32];
  tmem_ld32(t0, t);
  float sum = 0.0f, ss = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    t[i] += b[32 * grp + i];
    sum += t[i];
    ss = fmaf(t[i], t[i], ss);
  }
  stat[grp * 128 + row] = make_float2(sum, ss);
  named_bar_sync(NB_QUARTER + (row >> 5), kQuarterThreads);
  const float2 s0 = stat[row], s1 = stat[128 + row], s2 = stat[256 + row], s3 = stat[384 + row];
  const float mean = ((s0.x + s1.x) + (s2.x + s3.x)) * (1.0f / 128.0f);
  const float var = fmaxf(((s0.y + s1.y) + (s2.y + s3.y)) * (1.0f / 128.0f) - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
  float dot = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    dot = fmaf(silu_fast(fmaf(fmaf(t[i], rstd, nmr), g[32 * grp + i], be[32 * grp + i])), w3[32 * grp + i], dot);
  named_bar_sync(NB_QUARTER + (row >> 5), kQuarterThreads);     // the row's four threads have read the statistics
  stat[grp * 128 + row].x = dot;
  named_bar_sync(NB_QUARTER + (row >> 5), kQuarterThreads);
  return (stat[row].x + stat[128 + row].x) + (stat[256 + row].x + stat[384 + row].x);
}

__global__ void __launch_bounds__(kThreads, 1) k_readout_tc(TcRoP p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const X = smem + RoSmem::X;
  unsigned char* const RING = smem + RoSmem::RING;
  float* const PRM = reinterpret_cast<float*>(smem + RoSmem::PRM);
  float2* const STAT = reinterpret_cast<float2*>(smem + RoSmem::STAT);
  float* const VO = reinterpret_cast<float*>(smem + RoSmem::VO);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + RoSmem::BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + RoSmem::BARS + 8 * U_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;
  pdl_trigger();        // programmatic dependent launch: the next kernel of the stream may run its prologue now

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars[U_FULL + i], 1); mbar_init(&bars[U_EMPTY + i], 1); }
    mbar_init(&bars[U_OPS], kEpiThreads);
    mbar_init(&bars[U_ACC], 1);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 128);
  {
    const float* src[8] = {p.b1, p.g1, p.be1, p.b2, p.g2, p.be2, p.w3, p.vout};
    for (int i = tid; i < 8 * kF; i += kThreads) PRM[i] = __ldg(src[i >> 7] + (i & 127));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != 16) pdl_wait();      // barriers, TMEM and parameters were set up under the previous kernel; the weight producer
                                   // (warp 16) reads only weights, which no kernel writes, and starts streaming at once
  const uint32_t tmem = *tmem_slot;

  if (warp == 16) {
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
        for (int c = 0; c < kRoChunks; ++c) {
          mbar_wait(&bars[U_EMPTY + stage], ph ^ 1, err);
          mbar_arrive_expect_tx(&bars[U_FULL + stage], kChunkBytes);
          bulk_g2s(RING + stage * kChunkBytes, p.wblob + (size_t)c * kChunkBytes, kChunkBytes, &bars[U_FULL + stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
    }
  } else if (warp == 17) {
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0, pops = 0;
      const uint32_t xa = smem_u32(X), ring = smem_u32(RING);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(&bars[U_OPS], pops, err); pops ^= 1; tc_fence_after();
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&bars[U_FULL + stage], ph, err);
            tc_fence_after();
            mma_f16x3(tmem, xa + kb * (2 * kKStepBytes), kOperandHalfBytes, ring + stage * kChunkBytes, kChunkHalfBytes, 2,
                      kb > 0, p.passes);
            tc_commit(&bars[U_EMPTY + stage]);
            if (++stage == kStages) { stage = 0; ph ^= 1; }
          }
          tc_commit(&bars[U_ACC]);
        }
    }
  } else {
    const int grp = warp >> 2, wq = warp & 3;
    const int row = 32 * wq + lane;
    const uint32_t T0 = tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t pacc = 0;
    auto ops_done = [&]() { fence_proxy_async(); tc_fence_before(); mbar_arrive(&bars[U_OPS]); };
    auto acc_ready = [&]() { mbar_wait(&bars[U_ACC], pacc, err); pacc ^= 1; tc_fence_after(); };
    const float4 vo4 = *reinterpret_cast<const float4*>(PRM + 7 * kF + 4 * lane);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int node0 = tile * p.tile_nodes;
      const int rows = min(p.tile_nodes, p.n_nodes - node0);
      upd_build(X, wq, grp, lane, rows, p.s + (size_t)node0 * kF, kF);
      ops_done();
      // Vout . v[node][xyz], one node per warp iteration, under the first GEMM          (cpainn.py:434-436)
#pragma unroll 1
      for (int r = warp; r < rows; r += 16) {
        const float* vr = p.v + (size_t)(node0 + r) * 3 * kF + 4 * lane;
        float a[3];
#pragma unroll
        for (int xyz = 0; xyz < 3; ++xyz) {
          const float4 t = *reinterpret_cast<const float4*>(vr + xyz * kF);
          a[xyz] = fmaf(t.x, vo4.x, fmaf(t.y, vo4.y, fmaf(t.z, vo4.z, t.w * vo4.w)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
          a[1] += __shfl_xor_sync(0xffffffffu, a[1], o);
          a[2] += __shfl_xor_sync(0xffffffffu, a[2], o);
        }
        if (lane < 3) VO[r * 4 + lane] = a[lane];
      }
      acc_ready();
      upd_hidden(T0, grp, row, PRM, PRM + kF, PRM + 2 * kF, X, STAT, kStateUnscale);
      ops_done();
      acc_ready();
      const float gate = ro_gate(T0, grp, row, PRM + 3 * kF, PRM + 4 * kF, PRM + 5 * kF, PRM + 6 * kF, STAT) + __ldg(p.b3);
      if (grp == 0 && row < rows) {
        if (!(fabsf(gate) <= 3.0e38f)) p.err[1] = 1;     // NaN / inf: beyond the split-f16 range (tib_model_status reports it)
        float* o = p.out + (size_t)(node0 + row) * 3;
        o[0] = VO[row * 4 + 0] * gate; o[1] = VO[row * 4 + 1] * gate; o[2] = VO[row * 4 + 2] * gate;
      }
      tc_fence_before();
      named_bar_sync(NB_ALL, kEpiThreads);     // TMEM, STAT and VO are reused by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 128);
}

}  // namespace tc
}  // namespace tib
