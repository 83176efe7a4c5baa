// tc_update.cuh - Update.forward (cpainn.py:345-376) on the sm_100a tensor cores (F = 128).
//
//   vv = V v, uv = U v (per xyz plane);  q = |vv| over xyz;  (g, a, c) = split(MLP_{2F->F->F->3F}(cat[q, s]));
//   v += uv * g;   s += q^2 * a + c
//
// Work unit: a tile of 128 consecutive nodes.  Every GEMM is D[node][feat] = A[node][k] * W[feat][k]^T
// (node = TMEM lane), so LayerNorm, the norm over xyz and the gated residuals are all row-local.  Each plane
// of v is turned into an operand image ONCE and multiplied by V and by U while it is resident:
//   steps 1-3 (xyz = 0,1,2): plane -> X / Y / X;  vv = P V^T -> T0,  uv_xyz = P U^T -> T1+xyz;
//                            epilogue: q2 += vv^2 (registers)
//   step 4: q = sqrt(q2) -> X,  s -> Y;   layer 1 = X*W1[:, :F]^T + Y*W1[:, F:]^T -> T0
//   step 5: LN/SiLU(T0) -> X;             layer 2 -> T0
//   step 6: LN/SiLU(T0) -> X;             g = X*W3g^T -> T0
//   step 7: v_new[xyz] = v + uv_xyz * g  (T1..T3, T0);      a = X*W3a^T -> T1 once plane 0 is updated, c -> T2 after plane 1
//   step 8: s_new = s + q2 * a + c
// The epilogue (512 threads) and the MMA issuer alternate through one operand-ready / accumulator-ready
// barrier pair; weights stream through the same bulk-copy ring as the message kernel.
// Warp roles: warps 0-15 epilogue (group g = warp / 4 owns feature columns [32g, 32g+32) of every row),
// warp 16 weight producer (+ TMEM allocation), warp 17 MMA issuer.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_message.cuh"

namespace tib {
namespace tc {

constexpr int kUpdChunks = 48;     // streamed per tile: (V | U) x 3 planes | W1q | W1s | W2 | W3g | W3a | W3c
// weight blob of one update layer (matrix index -> 4 chunks): 0 V, 1 W1[:, :F], 2 W1[:, F:], 3 W2,
// 4 W3 rows [F,2F) (a), 5 W3 rows [2F,3F) (c), 6 W3 rows [0,F) (g), 7 U
__constant__ int kUpdOrder[12] = {0, 7, 0, 7, 0, 7, 1, 2, 3, 6, 4, 5};

struct TcUpdP {
  int n_nodes, n_tiles;
  int tile_nodes;      // nodes per tile (<= 128 TMEM lanes): 128 for full grids, fewer when the batch is small (tc_tile_nodes)
  float* s;                 // [N][F]    in place
  float* v;                 // [N][3][F] in place
  const unsigned char* wblob;   // 8 matrices x 4 chunks
  const float *b1, *g1, *be1, *b2, *g2, *be2, *b3;   // MLP parameters (fp32)
  int passes;
  int* err;
  long long* dbg;           // optional phase counters (diagnostics): [gridDim.x][16] cycles of thread 0 per step
};

struct UpdSmem {
  static constexpr uint32_t X = 0;
  static constexpr uint32_t Y = X + kOperandBytes;
  static constexpr uint32_t RING = Y + kOperandBytes;
  static constexpr uint32_t PRM = RING + kStages * kChunkBytes;    // b1 g1 be1 b2 g2 be2 (6F) | b3 (3F)
  static constexpr uint32_t STAT = PRM + 9 * 128 * 4;              // float2 [4 groups][128 rows]
  static constexpr uint32_t BARS = STAT + 4 * 128 * 8;
  static constexpr uint32_t TOTAL = BARS + 256;
};
enum { U_FULL = 0, U_EMPTY = U_FULL + kStages, U_OPS = U_EMPTY + kStages, U_ACC, U_COUNT };

// rows [32*wq, +32) x column groups [4*grp, +4) of an operand image from fp32 global rows base + r*stride
// (raw state: scaled by kStateScale, see tc_common.cuh)
__device__ __noinline__ void upd_build(unsigned char* op, int wq, int grp, int lane, int rows, const float* base, size_t stride) {
  const int g = 4 * grp + (lane >> 3);
  float4 a[4], b[4];
  // all eight 16-byte loads are issued before the first conversion (the build is bound by their latency)
#pragma unroll
  for (int oct = 0; oct < 4; ++oct) {
    const int r = 32 * wq + 8 * oct + (lane & 7);
    a[oct] = b[oct] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const float* src = base + (size_t)r * stride + g * 8;
      a[oct] = *reinterpret_cast<const float4*>(src);
      b[oct] = *reinterpret_cast<const float4*>(src + 4);
    }
  }
#pragma unroll
  for (int oct = 0; oct < 4; ++oct) {
    const int r = 32 * wq + 8 * oct + (lane & 7);
    const float v[8] = {a[oct].x * kStateScale, a[oct].y * kStateScale, a[oct].z * kStateScale, a[oct].w * kStateScale,
                        b[oct].x * kStateScale, b[oct].y * kStateScale, b[oct].z * kStateScale, b[oct].w * kStateScale};
    store_group(op, kOperandHalfBytes, r, g, v);
  }
}

// accumulator row, columns [32*grp, +32): + bias -> LayerNorm over all 128 columns (4-way statistics
// exchange) -> SiLU -> operand image
__device__ __noinline__ void upd_hidden(uint32_t taddr, int grp, int row, const float* b, const float* g, const float* be,
                                        unsigned char* op, float2* stat, float ascale) {
  const uint32_t t0 = taddr + 32 * grp;
  const float* bq = b + 32 * grp;
  float sum = 0.0f, ss = 0.0f;
#pragma unroll 1
  for (int kg = 0; kg < 4; kg += 2) {
    float t[8], u[8];
    tmem_ld8x2(t0 + 8 * kg, t0 + 8 * kg + 8, 0, t, u);
    const float4* bp = reinterpret_cast<const float4*>(bq) + 2 * kg;
    const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
    const float bb[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x0 = fmaf(t[i], ascale, bb[i]), x1 = fmaf(u[i], ascale, bb[8 + i]);
      sum += x0 + x1;
      ss = fmaf(x0, x0, fmaf(x1, x1, ss));
    }
  }
  stat[grp * 128 + row] = make_float2(sum, ss);
  named_bar_sync(NB_QUARTER + (row >> 5), kQuarterThreads);
  const float2 s0 = stat[row], s1 = stat[128 + row], s2 = stat[256 + row], s3 = stat[384 + row];
  const float mean = ((s0.x + s1.x) + (s2.x + s3.x)) * (1.0f / 128.0f);
  const float var = fmaxf(((s0.y + s1.y) + (s2.y + s3.y)) * (1.0f / 128.0f) - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
#pragma unroll 1
  for (int kg = 0; kg < 4; ++kg) {
    float t[8];
    tmem_ld8(t0 + 8 * kg, t);
    const float4* bp = reinterpret_cast<const float4*>(bq) + 2 * kg;
    const float4* gp = reinterpret_cast<const float4*>(g + 32 * grp) + 2 * kg;
    const float4* ep = reinterpret_cast<const float4*>(be + 32 * grp) + 2 * kg;
    const float4 b0 = bp[0], b1 = bp[1], g0 = gp[0], g1 = gp[1], e0 = ep[0], e1 = ep[1];
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = silu_fast(fmaf(fmaf(fmaf(t[i], ascale, bb[i]), rstd, nmr), gg[i], ee[i]));
    store_group(op, kOperandHalfBytes, row, 4 * grp + kg, y);
  }
  named_bar_sync(NB_QUARTER + (row >> 5), kQuarterThreads);     // `stat` rows of this quarter may be rewritten by the next call
}

// DIAG = the phase counters of tib_debug_counters (sixteen 64-bit counters): a template flag, as in k_message_tc
template <bool DIAG>
__global__ void __launch_bounds__(kThreads, 1) k_update_tc(TcUpdP p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const X = smem + UpdSmem::X;
  unsigned char* const Y = smem + UpdSmem::Y;
  unsigned char* const RING = smem + UpdSmem::RING;
  float* const PRM = reinterpret_cast<float*>(smem + UpdSmem::PRM);
  float2* const STAT = reinterpret_cast<float2*>(smem + UpdSmem::STAT);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + UpdSmem::BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + UpdSmem::BARS + 8 * U_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;
  pdl_trigger();        // programmatic dependent launch: the next kernel of the stream may run its prologue now

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars[U_FULL + i], 1); mbar_init(&bars[U_EMPTY + i], 1); }
    mbar_init(&bars[U_OPS], kEpiThreads);
    mbar_init(&bars[U_ACC], 1);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 9 * kF; i += kThreads) {
    const int r = i >> 7;
    const float* src = r == 0 ? p.b1 : r == 1 ? p.g1 : r == 2 ? p.be1 : r == 3 ? p.b2 : r == 4 ? p.g2 : r == 5 ? p.be2 : p.b3 + (r - 6) * kF;
    PRM[i] = __ldg(src + (i & 127));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != 16) pdl_wait();      // barriers, TMEM and parameters were set up under the previous kernel; the weight producer
                                   // (warp 16) reads only weights, which no kernel writes, and starts streaming at once
  const uint32_t tmem = *tmem_slot;

  if (warp == 16) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x)
        for (int c = 0; c < kUpdChunks; ++c) {
          mbar_wait(&bars[U_EMPTY + stage], ph ^ 1, err);
          mbar_arrive_expect_tx(&bars[U_FULL + stage], kChunkBytes);
          bulk_g2s(RING + stage * kChunkBytes, p.wblob + (size_t)(kUpdOrder[c >> 2] * 4 + (c & 3)) * kChunkBytes, kChunkBytes,
                   &bars[U_FULL + stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
    }
  } else if (warp == 17) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0, pops = 0;
      const uint32_t xa = smem_u32(X), ya = smem_u32(Y), ring = smem_u32(RING);
      // one streamed [128 x 128] weight matrix against one or two operand images
      auto gemm = [&](uint32_t d0, uint32_t op0, uint32_t d1, uint32_t op1, bool accumulate) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&bars[U_FULL + stage], ph, err);
          tc_fence_after();
          const uint32_t wst = ring + stage * kChunkBytes, off = kb * (2 * kKStepBytes);
          mma_f16x3(d0, op0 + off, kOperandHalfBytes, wst, kChunkHalfBytes, 2, accumulate || kb > 0, p.passes);
          if (op1) mma_f16x3(d1, op1 + off, kOperandHalfBytes, wst, kChunkHalfBytes, 2, kb > 0, p.passes);
          tc_commit(&bars[U_EMPTY + stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
      };
      auto ops_ready = [&]() { mbar_wait(&bars[U_OPS], pops, err); pops ^= 1; tc_fence_after(); };
      const uint32_t T0 = tmem, T1 = tmem + 128, T2 = tmem + 256, T3 = tmem + 384;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        ops_ready(); gemm(T0, xa, 0, 0, false); gemm(T1, xa, 0, 0, false); tc_commit(&bars[U_ACC]);   // 1: vv, uv0 (plane 0 in X)
        ops_ready(); gemm(T0, ya, 0, 0, false); gemm(T2, ya, 0, 0, false); tc_commit(&bars[U_ACC]);   // 2: vv, uv1 (plane 1 in Y)
        ops_ready(); gemm(T0, xa, 0, 0, false); gemm(T3, xa, 0, 0, false); tc_commit(&bars[U_ACC]);   // 3: vv, uv2 (plane 2 in X)
        ops_ready(); gemm(T0, xa, 0, 0, false); gemm(T0, ya, 0, 0, true); tc_commit(&bars[U_ACC]);    // 4: layer 1 (q in X, s in Y)
        ops_ready(); gemm(T0, xa, 0, 0, false); tc_commit(&bars[U_ACC]);                               // 5: layer 2
        ops_ready(); gemm(T0, xa, 0, 0, false); tc_commit(&bars[U_ACC]);                               // 6: g
        ops_ready(); gemm(T1, xa, 0, 0, false);                                                        // 7a: a -> T1 (uv0 consumed)
        ops_ready(); gemm(T2, xa, 0, 0, false); tc_commit(&bars[U_ACC]);                               // 7b: c -> T2 (uv1 consumed)
      }
    }
  } else {
    // =========================== builders / epilogue (512 threads) ===========================
    const int grp = warp >> 2, wq = warp & 3;
    const int row = 32 * wq + lane;
    const uint32_t lt = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t T0 = lt, T1 = lt + 128;
    uint32_t pacc = 0;
    const bool diag = DIAG && tid == 0;
    long long phc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = diag ? clock64() : 0;
#define TIB_UPH(i) do { if (diag) { const long long _t = clock64(); phc[i] += _t - tlast; tlast = _t; } } while (0)
    auto ops_done = [&]() { fence_proxy_async(); tc_fence_before(); mbar_arrive(&bars[U_OPS]); };
    auto acc_ready = [&]() { mbar_wait(&bars[U_ACC], pacc, err); pacc ^= 1; tc_fence_after(); };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int node0 = tile * p.tile_nodes;
      const int rows = min(p.tile_nodes, p.n_nodes - node0);
      const float* vbase = p.v + (size_t)node0 * 3 * kF;
      const bool live = row < rows;
      const size_t node = (size_t)(node0 + row);
      // The next tile's v and s rows (two contiguous blocks, 2048 lines) start towards L2 now: its builds then pay
      // the L2 latency, not the HBM latency.  (The first tile of a CTA is not prefetched.)
      {
        const int nt = tile + gridDim.x;
        if (nt < p.n_tiles) {
          const int nrows = min(p.tile_nodes, p.n_nodes - nt * p.tile_nodes);
          const char* vb = reinterpret_cast<const char*>(p.v + (size_t)nt * p.tile_nodes * 3 * kF);
          const char* sb = reinterpret_cast<const char*>(p.s + (size_t)nt * p.tile_nodes * kF);
          for (int i = tid; i < nrows * 12; i += kEpiThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (size_t)i * 128));
          if (tid < nrows * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(sb + (size_t)tid * 128));
        }
      }
      // steps 1-3: planes of v; q2 = sum over xyz of (V v)^2                            (cpainn.py:358-361)
      float q2[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) q2[i] = 0.0f;
      upd_build(X, wq, grp, lane, rows, vbase, 3 * kF);
      ops_done();
      upd_build(Y, wq, grp, lane, rows, vbase + kF, 3 * kF);          // under the MMAs of plane 0
      TIB_UPH(0);
#pragma unroll 1
      for (int xyz = 0; xyz < 3; ++xyz) {
        acc_ready();                                                  // vv, uv_xyz of plane xyz
        TIB_UPH(1);
        {
          float t[32];
          tmem_ld32(T0 + 32 * grp, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) q2[i] = fmaf(t[i], t[i], q2[i]);
        }
        if (xyz < 2) ops_done();                                      // T0 drained: the next plane's MMAs (operand built earlier) may start
        // operands that are needed two steps from now, under the MMAs just released
        if (xyz == 0) upd_build(X, wq, grp, lane, rows, vbase + 2 * kF, 3 * kF);   // plane 2 (X is free: plane 0 is done)
        if (xyz == 1) upd_build(Y, wq, grp, lane, rows, p.s + (size_t)node0 * kF, kF);   // s (Y is free: plane 1 is done)
        TIB_UPH(2);
      }
      // step 4: q -> X (X is free: the MMAs of plane 2 have completed).  q2 holds (kStateScale |V v|)^2, so its
      // root is q already scaled like every other raw-state operand.
#pragma unroll
      for (int kg = 0; kg < 4; ++kg) {
        float q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = sqrt_fast(q2[8 * kg + i]);
        store_group(X, kOperandHalfBytes, row, 4 * grp + kg, q);
      }
      ops_done();
      TIB_UPH(4);
      // step 5
      acc_ready();
      TIB_UPH(5);
      upd_hidden(T0, grp, row, PRM, PRM + kF, PRM + 2 * kF, X, STAT, kStateUnscale);
      ops_done();
      TIB_UPH(6);
      // step 6
      acc_ready();
      TIB_UPH(7);
      upd_hidden(T0, grp, row, PRM + 3 * kF, PRM + 4 * kF, PRM + 5 * kF, X, STAT, 1.0f);
      ops_done();
      TIB_UPH(8);
      // step 7: v += (U v) * g                                                        (cpainn.py:370,374)
      acc_ready();
      TIB_UPH(9);
      {
        // Row-per-thread global accesses would cost 32 wavefronts per instruction (rows are 1.5 KB apart), so
        // each plane goes through the (free) Y buffer: coalesced load -> swizzled fp32 tile -> row-local
        // update -> coalesced store.  16-byte chunk c of row r lives at chunk (c ^ (r & 31)).
        // two halves of 16 columns: g and (U v) of 16 columns are live next to q2[32], not 2 x 32 (no spills at 96 registers)
        float4* const stage = reinterpret_cast<float4*>(Y);
#pragma unroll 1
        for (int xyz = 0; xyz < 3; ++xyz) {
          float* const vplane = p.v + ((size_t)node0 * 3 + xyz) * kF;
          // coalesced load: warp w, iteration it -> row 8*it' ... : 512 threads x 8 chunks = 128 rows x 32 chunks
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = (tid >> 5) + 16 * it, c = tid & 31;         // one row (512 B) per warp instruction
            if (r < rows) stage[r * 32 + (c ^ (r & 31))] = *reinterpret_cast<const float4*>(vplane + (size_t)r * 3 * kF + 4 * c);
          }
          named_bar_sync(NB_ALL, kEpiThreads);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float g[16], u[16];
            tmem_ld16x2(T0 + 32 * grp + 16 * h, lt + 128 * (1 + xyz) + 32 * grp + 16 * h, g, u);
#pragma unroll
            for (int i = 0; i < 16; ++i) g[i] = (g[i] + PRM[6 * kF + 32 * grp + 16 * h + i]) * kStateUnscale;   // T1..T3 hold kStateScale * U v
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4* sp = stage + row * 32 + ((8 * grp + 4 * h + j) ^ (row & 31));
              float4 t = *sp;
              t.x = fmaf(u[4 * j + 0], g[4 * j + 0], t.x); t.y = fmaf(u[4 * j + 1], g[4 * j + 1], t.y);
              t.z = fmaf(u[4 * j + 2], g[4 * j + 2], t.z); t.w = fmaf(u[4 * j + 3], g[4 * j + 3], t.w);
              *sp = t;
            }
          }
          named_bar_sync(NB_ALL, kEpiThreads);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = (tid >> 5) + 16 * it, c = tid & 31;
            if (r < rows) *reinterpret_cast<float4*>(vplane + (size_t)r * 3 * kF + 4 * c) = stage[r * 32 + (c ^ (r & 31))];
          }
          named_bar_sync(NB_ALL, kEpiThreads);                        // the stage is reused by the next plane
          if (xyz < 2) ops_done();      // uv_xyz (T1 / T2) has been read by everyone: a / c may be accumulated there now
        }
      }
      TIB_UPH(10);
      // step 8: s += q^2 * a + c                                                      (cpainn.py:371,373)
      {
        float4* const stage = reinterpret_cast<float4*>(Y);
        float* const srows = p.s + (size_t)node0 * kF;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = (tid >> 5) + 16 * it, c = tid & 31;
          if (r < rows) stage[r * 32 + (c ^ (r & 31))] = *reinterpret_cast<const float4*>(srows + (size_t)r * kF + 4 * c);
        }
        acc_ready();
        TIB_UPH(11);
        named_bar_sync(NB_ALL, kEpiThreads);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float a[16], c[16];
          tmem_ld16x2(T1 + 32 * grp + 16 * h, lt + 256 + 32 * grp + 16 * h, a, c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4* sp = stage + row * 32 + ((8 * grp + 4 * h + j) ^ (row & 31));
            float4 t = *sp;
            const float* pa = PRM + 7 * kF + 32 * grp + 16 * h + 4 * j;
            const float* pc = PRM + 8 * kF + 32 * grp + 16 * h + 4 * j;
            constexpr float kU2 = kStateUnscale * kStateUnscale;   // q2 is the square of the scaled norm
            t.x += fmaf(q2[16 * h + 4 * j + 0] * kU2, a[4 * j + 0] + pa[0], c[4 * j + 0] + pc[0]);
            t.y += fmaf(q2[16 * h + 4 * j + 1] * kU2, a[4 * j + 1] + pa[1], c[4 * j + 1] + pc[1]);
            t.z += fmaf(q2[16 * h + 4 * j + 2] * kU2, a[4 * j + 2] + pa[2], c[4 * j + 2] + pc[2]);
            t.w += fmaf(q2[16 * h + 4 * j + 3] * kU2, a[4 * j + 3] + pa[3], c[4 * j + 3] + pc[3]);
            *sp = t;
          }
        }
        tc_fence_before();
        named_bar_sync(NB_ALL, kEpiThreads);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = (tid >> 5) + 16 * it, c2 = tid & 31;
          if (r < rows) *reinterpret_cast<float4*>(srows + (size_t)r * kF + 4 * c2) = stage[r * 32 + (c2 ^ (r & 31))];
        }
      }
      tc_fence_before();
      named_bar_sync(NB_ALL, kEpiThreads);                 // the tile's TMEM reads finish before the next tile's MMAs
      TIB_UPH(14);
    }
    if (diag) for (int i = 0; i < 16; ++i) p.dbg[(size_t)blockIdx.x * 16 + i] = phc[i];
#undef TIB_UPH
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace tib
