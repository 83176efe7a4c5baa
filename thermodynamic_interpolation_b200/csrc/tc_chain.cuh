// tc_chain.cuh - one MLP of the reference (embedding.py:8-49: Linear -> LayerNorm -> SiLU -> Linear -> LayerNorm ->
// SiLU -> Linear) over a tall matrix of rows, on the sm_100a tensor cores, for F = 128 and F = 256, in two flavours:
//   PRIMAL : out = MLP(A)            (optionally saving the normalised pre-activations n1, n2 and 1/std per row)
//   JVP    : out = J_MLP(A) . Adot   (forward-mode tangent rows through the same weights, using the saved n1, n2, 1/std)
// and as a plain GEMM (n_hidden = 0): out = A W^T.  This is the "layered" engine: the F = 256 drift (the fused
// k_message_tc holds two 64 KB operand images per tile, which F = 256 doubles past the 227 KB of shared memory) and
// the exact divergence (ode_wrapper.py:59-91), whose tangent GEMMs are the same weight matrices applied to D tangent
// rows per primal row.
//
// Work item = 128 consecutive rows of one direction.  Hidden layers  D[row][feat] = A[row][k] W[feat][k]^T  (row = TMEM
// lane: LayerNorm and its Jacobian are row-local); the output layer is TRANSPOSED, D^T[feat][row] = W3[feat][k] H2[row][k]^T
// (feat = TMEM lane), so that a warp stores 32 consecutive features of one row: coalesced 128-byte stores without a
// shared-memory transpose.  Operand images, the streamed weight ring and the warp roles are those of tc_message.cuh.
//
// Split-f16 range: primal operands use the fixed scales of tc_common.cuh.  Tangent operands have no natural scale,
// so every tangent operand image carries a power-of-two scale - per tile for the input rows (two-pass build: max, then
// convert), per row for the hidden tangents (from a bound that needs no second exchange) - undone on the accumulators.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_message.cuh"

namespace tib {
namespace tc {

enum { SRC_NONE = 0, SRC_ROWS = 1, SRC_PE = 2, SRC_PE_D = 3 };
enum { EPI_PRIMAL = 0, EPI_JVP = 1 };

struct ChainSrc {
  int kind;              // SRC_*
  const float* base;     // SRC_ROWS: fp32 rows [*, ld]; SRC_PE / SRC_PE_D: one scalar per row at base[row * ld]
  long long ld;          // floats between rows
  const int* idx;        // optional gather: source row = idx[row * idx_stride]
  int idx_stride;
  long long dir_stride;  // floats added to base per direction
  float scale;           // SRC_ROWS (primal): multiplier of the raw values; SRC_PE*: the encoder's max_length
};

struct ChainP {
  int n_rows, n_tiles, n_dirs;
  int n_halves;                      // K_in / F: 1 or 2 (half h comes from src[h])
  ChainSrc src[2];
  int n_hidden;                      // 2 = MLP, 0 = plain GEMM (n_halves = 1)
  int epi;                           // EPI_*
  const unsigned char* wblob;        // 16 KB chunks in consumption order (pack_chain_* in tib_api.cu)
  const float *b1, *g1, *be1, *b2, *g2, *be2, *b3;
  float ascale1;                     // primal: multiplies the layer-1 accumulator (undoes src scale)
  float gmax1, gmax2;                // JVP: max |LayerNorm gain| of the two hidden layers (bound of the hidden tangents)
  float* save_n[2];                  // [n_tiles][F][128] normalised pre-activations (primal writes if non-null, JVP reads)
  float* save_r[2];                  // [n_rows] 1/std
  int n_out;                         // multiple of 128
  float* out;                        // [n_dirs][n_rows][ld_out]
  long long ld_out, out_dir_stride;
  float out_scale;                   // plain GEMM, primal: multiplies the accumulator (undoes src scale)
  int passes;
  int* err;
};

// NG = epilogue warp groups of 4 warps (one warp per TMEM lane quarter in each group).  NG = 4: 16 epilogue warps, one CTA
// per SM (F = 256 needs 128 KB of operand image).  NG = 2 (F = 128): 8 epilogue warps, one 64 KB operand buffer, a two-stage
// weight ring, 256 TMEM columns - TWO CTAs per SM, so that one CTA's MMAs run under the other's epilogue (a single chain is
// strictly serial: build -> MMA -> LayerNorm -> MMA -> LayerNorm -> MMA).
template <int F, int NG>
struct ChainSmem {
  static constexpr uint32_t OP_HALF = 128u * F * 2u;            // hi (or lo) image of a [128 x F] operand
  static constexpr uint32_t OP_BYTES = 2u * OP_HALF;
  static constexpr int NBUF = (F == 128 && NG == 4) ? 2 : 1;
  static constexpr int STAGES = NG == 4 ? 4 : 2;
  static constexpr int NSLOT = NG == 4 ? 2 : 1;                 // output accumulator slots of 128 columns
  static constexpr int TMEM_COLS = NG == 4 ? 512 : 256;
  static constexpr int EPI = NG * 128;                          // epilogue threads
  static constexpr int THREADS = EPI + 64;
  static constexpr int CTAS = NG == 4 ? 1 : 2;
  static constexpr uint32_t BUF = 0;
  static constexpr uint32_t RING = BUF + NBUF * OP_BYTES;
  static constexpr uint32_t PRM = RING + STAGES * kChunkBytes;   // b1 g1 be1 b2 g2 be2
  static constexpr uint32_t STAT = PRM + 6 * F * 4;              // 2 x float4 [NG groups][128 rows]
  static constexpr uint32_t INVS = STAT + 2 * NG * 128 * 16;     // float [128]: un-scale factor of each output column
  static constexpr uint32_t RED = INVS + 512;                    // float [16]: per-warp maxima of the tile-scale pass
  static constexpr uint32_t BARS = RED + 128;
  static constexpr uint32_t TOTAL = BARS + 256;
};
// C_OPS1: the second K half of layer 1 is announced on its own barrier - a thread arrives for both halves without a wait in
// between, and two arrivals of one thread must not complete a phase that another thread has not arrived on yet
enum { C_FULL = 0, C_EMPTY = C_FULL + 4, C_OPS = C_EMPTY + 4, C_OPS1, C_ACC, C_FREE0, C_FREE1, C_TFULL0, C_TFULL1,
       C_TEMPTY0, C_TEMPTY1, C_COUNT };

// 2^k with k clamped so that the result is a normal float
__device__ __forceinline__ float pow2i(int k) { return __uint_as_float((uint32_t)(min(max(k, -126), 127) + 127) << 23); }
__device__ __forceinline__ int floor_log2f(float x) { return (int)((__float_as_uint(x) >> 23) & 255u) - 127; }

// number of streamed chunks of one work item
template <int F>
__host__ __device__ constexpr int chain_chunks(int n_hidden, int n_halves, int n_out) {
  return (n_hidden ? (n_halves + 1) * (F / 128) * (F / 32) : 0) + (n_out / 128) * (F / 32);
}

// ---- input build: rows [32*wq, +32) x this thread's column groups of one K half -> operand image -------------------------
// thread (wq, grp, lane): rows 32*wq + 8*oct + (lane & 7), column groups 4*grp + (lane >> 3) + 4*NG*j
template <int F>
__device__ __forceinline__ const float* chain_row_ptr(const ChainSrc s, long long dir, long long row) {
  const long long r = s.idx ? (long long)__ldg(s.idx + row * s.idx_stride) : row;
  return s.base + dir * s.dir_stride + r * s.ld;
}

template <int F, int NG>
__device__ __noinline__ float chain_absmax(const ChainSrc s, long long dir, long long row0, int rows, int wq, int grp, int lane) {
  float m = 0.0f;
  if (s.kind != SRC_ROWS) return m;
#pragma unroll 1
  for (int j = 0; j < F / (32 * NG); ++j) {
    const int g = 4 * grp + (lane >> 3) + 4 * NG * j;
    float4 a[4], b[4];
#pragma unroll
    for (int oct = 0; oct < 4; ++oct) {
      const int r = 32 * wq + 8 * oct + (lane & 7);
      a[oct] = b[oct] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        const float* p = chain_row_ptr<F>(s, dir, row0 + r) + g * 8;
        a[oct] = __ldg(reinterpret_cast<const float4*>(p));
        b[oct] = __ldg(reinterpret_cast<const float4*>(p + 4));
      }
    }
#pragma unroll
    for (int oct = 0; oct < 4; ++oct) {
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a[oct].x), fabsf(a[oct].y)), fmaxf(fabsf(a[oct].z), fabsf(a[oct].w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(b[oct].x), fabsf(b[oct].y)), fmaxf(fabsf(b[oct].z), fabsf(b[oct].w))));
    }
  }
  return m;
}

template <int F, int NG>
__device__ __noinline__ void chain_build(unsigned char* op, const ChainSrc s, long long dir, long long row0, int rows, int wq,
                                         int grp, int lane, float scale) {
  constexpr uint32_t LO = 128u * F * 2u;
#pragma unroll 1
  for (int j = 0; j < F / (32 * NG); ++j) {
    const int g = 4 * grp + (lane >> 3) + 4 * NG * j;
    if (s.kind == SRC_ROWS) {
      float4 a[4], b[4];
#pragma unroll
      for (int oct = 0; oct < 4; ++oct) {
        const int r = 32 * wq + 8 * oct + (lane & 7);
        a[oct] = b[oct] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
          const float* p = chain_row_ptr<F>(s, dir, row0 + r) + g * 8;
          a[oct] = __ldg(reinterpret_cast<const float4*>(p));
          b[oct] = __ldg(reinterpret_cast<const float4*>(p + 4));
        }
      }
#pragma unroll
      for (int oct = 0; oct < 4; ++oct) {
        const int r = 32 * wq + 8 * oct + (lane & 7);
        const float v[8] = {a[oct].x * scale, a[oct].y * scale, a[oct].z * scale, a[oct].w * scale,
                            b[oct].x * scale, b[oct].y * scale, b[oct].z * scale, b[oct].w * scale};
        store_group(op, LO, r, g, v);
      }
    } else {
      // PositionalEncoder(d) (embedding.py:137-160) or its derivative with respect to d; columns 8g .. 8g+7 are the
      // (cos, sin) pairs of ranks 4g+1 .. 4g+4
#pragma unroll 1
      for (int oct = 0; oct < 4; ++oct) {
        const int r = 32 * wq + 8 * oct + (lane & 7);
        float v[8];
        const bool live = r < rows;
        const float d = live ? __ldg(s.base + dir * s.dir_stride + (row0 + r) * s.ld) : 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rank = 4 * g + q + 1;
          float sn = 0.0f, cs = 0.0f;
          if (live) sincos_cw(pe_arg(d, s.scale, rank), sn, cs);
          if (s.kind == SRC_PE) { v[2 * q] = cs; v[2 * q + 1] = sn; }
          else {
            const float w = __fmul_rn(__fdiv_rn((float)rank, s.scale), kPiF) * scale;   // d arg / d d
            v[2 * q] = -sn * w; v[2 * q + 1] = cs * w;
          }
        }
        store_group(op, LO, r, g, v);
      }
    }
  }
}

// ---- hidden epilogue -------------------------------------------------------------------------------------------------------
// accumulator row `row` (TMEM lane), columns [grp * F/4, +F/4).  Two passes over TMEM (statistics, then the values): the
// row quarter is not held in registers, so F = 256 fits the same register budget.
//   PRIMAL: z = acc * ascale + b ; n = (z - mean) / std ; h = SiLU(n g + be)                       -> image (unscaled)
//   JVP   : zd = acc * ainv ; nd = rstd (zd - mean zd - n mean(n zd)) ; hd = SiLU'(n g + be) g nd   -> image * 2^k(row)
// returns the row's image scale (1 for PRIMAL)
template <int F, int NG, int EPI>
__device__ __noinline__ float chain_hidden(uint32_t taddr, int grp, int row, bool live, const float* b, const float* g,
                                           const float* be, unsigned char* op, float4* stat, float ascale, float gmax,
                                           float* save_n, float* save_r) {
  constexpr int CPT = F / NG;                // columns per thread
  constexpr uint32_t LO = 128u * F * 2u;
  const int col0 = grp * CPT;
  float sum = 0.0f, ss = 0.0f, zmax = 0.0f;
#pragma unroll 1
  for (int pc = 0; pc < CPT / 32; ++pc) {
    float v[32];
    tmem_ld32(taddr + col0 + 32 * pc, v);
    if (EPI == EPI_PRIMAL) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float z = fmaf(v[i], ascale, b[col0 + 32 * pc + i]);
        sum += z; ss = fmaf(z, z, ss);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float z = v[i] * ascale;
        const float nn = live ? __ldg(save_n + (size_t)(col0 + 32 * pc + i) * 128 + row) : 0.0f;
        sum += z; ss = fmaf(nn, z, ss); zmax = fmaxf(zmax, fabsf(z));
      }
    }
  }
  stat[grp * 128 + row] = make_float4(sum, ss, zmax, 0.0f);
  named_bar_sync(NB_QUARTER + (row >> 5), 32 * NG);
  float4 s0 = stat[row], s1 = stat[128 + row], s2 = make_float4(0.f, 0.f, 0.f, 0.f), s3 = s2;
  if (NG == 4) { s2 = stat[256 + row]; s3 = stat[384 + row]; }
  const float m1 = ((s0.x + s1.x) + (s2.x + s3.x)) * (1.0f / F);
  const float m2 = ((s0.y + s1.y) + (s2.y + s3.y)) * (1.0f / F);
  float rstd, c0, c1, rowscale = 1.0f;
  if (EPI == EPI_PRIMAL) {
    const float var = fmaxf(m2 - m1 * m1, 0.0f);
    rstd = rsqrtf(var + 1e-5f);
    c0 = -m1 * rstd; c1 = 0.0f;
    if (save_r && grp == 0 && live) save_r[0] = rstd;
  } else {
    rstd = live ? __ldg(save_r) : 0.0f;
    c0 = m1; c1 = m2;            // mean(zd), mean(n zd)
    const float zm = fmaxf(fmaxf(s0.z, s1.z), fmaxf(s2.z, s3.z));
    const float bound = 1.1f * gmax * (2.0f + (F == 128 ? 11.4f : 16.0f)) * rstd * zm;
    if (bound > 0.0f && bound < 3.0e38f) rowscale = pow2i(13 - floor_log2f(bound));
  }
#pragma unroll 1
  for (int pc = 0; pc < CPT / 32; ++pc) {
    float v[32];
    tmem_ld32(taddr + col0 + 32 * pc, v);
#pragma unroll
    for (int kg = 0; kg < 4; ++kg) {
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = col0 + 32 * pc + 8 * kg + i;
        if (EPI == EPI_PRIMAL) {
          const float n = fmaf(fmaf(v[8 * kg + i], ascale, b[c]), rstd, c0);
          if (save_n && live) save_n[(size_t)c * 128 + row] = n;
          y[i] = silu_fast(fmaf(n, g[c], be[c]));
        } else {
          const float nn = live ? __ldg(save_n + (size_t)c * 128 + row) : 0.0f;
          const float nd = rstd * ((v[8 * kg + i] * ascale - c0) - nn * c1);
          const float zt = fmaf(nn, g[c], be[c]);
          float e, sg;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zt * -1.4426950408889634f));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + e));
          y[i] = (sg * fmaf(zt, 1.0f - sg, 1.0f)) * (g[c] * nd) * rowscale;
        }
      }
      store_group(op, LO, row, (col0 + 32 * pc) / 8 + kg, y);
    }
  }
  return rowscale;
}

template <int F, int NG>
__global__ void __launch_bounds__(ChainSmem<F, NG>::THREADS, ChainSmem<F, NG>::CTAS) k_chain_tc(ChainP p) {
  using S = ChainSmem<F, NG>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const BUF = smem + S::BUF;
  unsigned char* const RING = smem + S::RING;
  float* const PRM = reinterpret_cast<float*>(smem + S::PRM);
  float4* const STAT = reinterpret_cast<float4*>(smem + S::STAT);
  float* const INVS = reinterpret_cast<float*>(smem + S::INVS);
  float* const RED = reinterpret_cast<float*>(smem + S::RED);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + S::BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BARS + 8 * C_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;
  constexpr int KC = F / 32;          // chunks per K = F
  constexpr int NBH = F / 128;        // 128-column blocks of a hidden layer
  constexpr int EW = 4 * NG;          // epilogue warps; warp EW = weight producer, warp EW + 1 = MMA issuer
  constexpr int OCOLS = 128 / NG;     // output columns (= rows of the tile) per epilogue thread
  const int n_work = p.n_tiles * p.n_dirs;
  const int n_ob = p.n_out / 128;

  if (tid == 0) {
    for (int i = 0; i < S::STAGES; ++i) { mbar_init(&bars[C_FULL + i], 1); mbar_init(&bars[C_EMPTY + i], 1); }
    mbar_init(&bars[C_OPS], S::EPI); mbar_init(&bars[C_OPS1], S::EPI);
    mbar_init(&bars[C_ACC], 1);
    mbar_init(&bars[C_FREE0], 1); mbar_init(&bars[C_FREE1], 1);
    mbar_init(&bars[C_TFULL0], 1); mbar_init(&bars[C_TFULL1], 1);
    mbar_init(&bars[C_TEMPTY0], S::EPI); mbar_init(&bars[C_TEMPTY1], S::EPI);
    fence_mbar_init();
  }
  if (warp == EW) tmem_alloc(tmem_slot, S::TMEM_COLS);
  if (p.n_hidden) {
    for (int i = tid; i < 6 * F; i += S::THREADS) {
      const int r = i / F;
      const float* src = r == 0 ? p.b1 : r == 1 ? p.g1 : r == 2 ? p.be1 : r == 3 ? p.b2 : r == 4 ? p.g2 : p.be2;
      PRM[i] = __ldg(src + (i - r * F));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_hid = tmem, t_out = tmem + F;       // hidden accumulator | NSLOT output slots of 128 columns

  if (warp == EW) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      const int per_item = chain_chunks<F>(p.n_hidden, p.n_halves, p.n_out);
      int stage = 0; uint32_t ph = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x)
        for (int c = 0; c < per_item; ++c) {
          mbar_wait(&bars[C_EMPTY + stage], ph ^ 1, err);
          mbar_arrive_expect_tx(&bars[C_FULL + stage], kChunkBytes);
          bulk_g2s(RING + stage * kChunkBytes, p.wblob + (size_t)c * kChunkBytes, kChunkBytes, &bars[C_FULL + stage]);
          if (++stage == S::STAGES) { stage = 0; ph ^= 1; }
        }
    }
  } else if (warp == EW + 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0, pops[2] = {0, 0}, pte[2] = {0, 0};
      const uint32_t buf = smem_u32(BUF), ring = smem_u32(RING);
      // one streamed chunk (128 weight rows x 32 k) against the operand image at `op`
      auto chunk = [&](uint32_t d, uint32_t op, int kc, bool transposed, bool acc) {
        mbar_wait(&bars[C_FULL + stage], ph, err);
        tc_fence_after();
        const uint32_t wst = ring + stage * kChunkBytes, opk = op + kc * (2 * kKStepBytes);
        if (!transposed) mma_f16x3(d, opk, S::OP_HALF, wst, kChunkHalfBytes, 2, acc, p.passes);
        else             mma_f16x3(d, wst, kChunkHalfBytes, opk, S::OP_HALF, 2, acc, p.passes);
        tc_commit(&bars[C_EMPTY + stage]);
        if (++stage == S::STAGES) { stage = 0; ph ^= 1; }
      };
      auto ops_ready = [&](int h = 0) { mbar_wait(&bars[C_OPS + h], pops[h], err); pops[h] ^= 1; tc_fence_after(); };
      int oc = 0;                                                  // running output-block counter (slot = oc % NSLOT)
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        if (p.n_hidden) {
          for (int h = 0; h < p.n_halves; ++h) {                   // layer 1, one K half at a time
            ops_ready(h);
            const uint32_t op = buf + (h % S::NBUF) * S::OP_BYTES;
#pragma unroll 1
            for (int nb = 0; nb < NBH; ++nb)
#pragma unroll 1
              for (int kc = 0; kc < KC; ++kc) chunk(t_hid + 128 * nb, op, kc, false, h > 0 || kc > 0);
            if (h + S::NBUF < p.n_halves) tc_commit(&bars[C_FREE0 + (h % S::NBUF)]);
          }
          tc_commit(&bars[C_ACC]);
          ops_ready();                                             // layer 2 (H1 in buffer 0)
#pragma unroll 1
          for (int nb = 0; nb < NBH; ++nb)
#pragma unroll 1
            for (int kc = 0; kc < KC; ++kc) chunk(t_hid + 128 * nb, buf, kc, false, kc > 0);
          tc_commit(&bars[C_ACC]);
        }
        ops_ready();                                               // output layer (H2, or the plain input, in buffer 0)
#pragma unroll 1
        for (int ob = 0; ob < n_ob; ++ob, ++oc) {
          const int sl = oc % S::NSLOT;
          mbar_wait(&bars[C_TEMPTY0 + sl], pte[sl] ^ 1, err); pte[sl] ^= 1; tc_fence_after();
#pragma unroll 1
          for (int kc = 0; kc < KC; ++kc) chunk(t_out + 128 * sl, buf, kc, true, kc > 0);
          tc_commit(&bars[C_TFULL0 + sl]);
        }
      }
    }
  } else {
    // =========================== builders / epilogue (128 NG threads) ===========================
    const int grp = warp >> 2, wq = warp & 3;
    const int row = 32 * wq + lane;
    const uint32_t lt = tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t pacc = 0, pfree[2] = {0, 0}, ptf[2] = {0, 0};
    int oc = 0, statsel = 0;
    auto ops_done = [&](int h = 0) { fence_proxy_async(); tc_fence_before(); mbar_arrive(&bars[C_OPS + h]); };
    auto acc_ready = [&]() { mbar_wait(&bars[C_ACC], pacc, err); pacc ^= 1; tc_fence_after(); };
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int dir = w / p.n_tiles, tile = w - dir * p.n_tiles;
      const long long row0 = (long long)tile * 128;
      const int rows = (int)min((long long)128, (long long)p.n_rows - row0);
      const bool live = row < rows;
      // ---- input scale: fixed (primal), or 2^k from the tile's largest tangent magnitude
      float in_scale = 1.0f, in_inv = 1.0f;
      if (p.epi == EPI_JVP) {
        float m = 0.0f;
        m = chain_absmax<F, NG>(p.src[0], dir, row0, rows, wq, grp, lane);
        if (p.n_halves > 1) m = fmaxf(m, chain_absmax<F, NG>(p.src[1], dir, row0, rows, wq, grp, lane));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        named_bar_sync(NB_ALL, S::EPI);                             // RED of the previous item has been read
        if (lane == 0) RED[warp] = m;
        named_bar_sync(NB_ALL, S::EPI);
#pragma unroll
        for (int i = 0; i < EW; ++i) m = fmaxf(m, RED[i]);
        if (p.src[0].kind == SRC_PE_D) m = fmaxf(m, 64.0f);          // derivative of the encoder: at most F/2 * pi / length
        if (m > 0.0f && m < 3.0e38f) { const int e = floor_log2f(m); in_scale = pow2i(9 - e); in_inv = pow2i(e - 9); }
      }
      float rowscale = 1.0f;
      if (p.n_hidden) {
        for (int h = 0; h < p.n_halves; ++h) {
          const int bi = h % S::NBUF;
          if (h >= S::NBUF) { mbar_wait(&bars[C_FREE0 + bi], pfree[bi], err); pfree[bi] ^= 1; }
          // constant indices only: indexing the kernel parameter with h would make ptxas copy it to local memory
          if (h == 0) chain_build<F, NG>(BUF + bi * S::OP_BYTES, p.src[0], dir, row0, rows, wq, grp, lane, p.epi == EPI_JVP ? in_scale : p.src[0].scale);
          else        chain_build<F, NG>(BUF + bi * S::OP_BYTES, p.src[1], dir, row0, rows, wq, grp, lane, p.epi == EPI_JVP ? in_scale : p.src[1].scale);
          ops_done(h);
        }
        const size_t so = (size_t)tile * F * 128;
        float4* const st0 = STAT + statsel * (NG * 128);
        float4* const st1 = STAT + (statsel ^ 1) * (NG * 128);
        acc_ready();
        if (p.epi == EPI_PRIMAL)
          chain_hidden<F, NG, EPI_PRIMAL>(lt, grp, row, live, PRM, PRM + F, PRM + 2 * F, BUF, st0, p.ascale1, 0.0f,
                                          p.save_n[0] ? p.save_n[0] + so : nullptr, p.save_r[0] ? p.save_r[0] + row0 + row : nullptr);
        else
          rowscale = chain_hidden<F, NG, EPI_JVP>(lt, grp, row, live, PRM, PRM + F, PRM + 2 * F, BUF, st0, in_inv, p.gmax1,
                                                  p.save_n[0] + so, p.save_r[0] + row0 + row);
        ops_done();
        acc_ready();
        if (p.epi == EPI_PRIMAL)
          chain_hidden<F, NG, EPI_PRIMAL>(lt, grp, row, live, PRM + 3 * F, PRM + 4 * F, PRM + 5 * F, BUF, st1, 1.0f, 0.0f,
                                          p.save_n[1] ? p.save_n[1] + so : nullptr, p.save_r[1] ? p.save_r[1] + row0 + row : nullptr);
        else
          rowscale = chain_hidden<F, NG, EPI_JVP>(lt, grp, row, live, PRM + 3 * F, PRM + 4 * F, PRM + 5 * F, BUF, st1, 1.0f / rowscale,
                                                  p.gmax2, p.save_n[1] + so, p.save_r[1] + row0 + row);
        if (grp == 0) INVS[row] = 1.0f / rowscale;
      } else {
        chain_build<F, NG>(BUF, p.src[0], dir, row0, rows, wq, grp, lane, p.epi == EPI_JVP ? in_scale : p.src[0].scale);
        if (grp == 0) INVS[row] = p.epi == EPI_JVP ? in_inv : p.out_scale;
      }
      ops_done();
      named_bar_sync(NB_ALL, S::EPI);                               // INVS is complete
      // ---- output layer, transposed: this thread owns feature f = TMEM lane of each 128-feature block; columns are rows
      float* const obase = p.out + (long long)dir * p.out_dir_stride + (row0 + OCOLS * grp) * p.ld_out + row;
      const int qn = min(OCOLS, rows - OCOLS * grp);
#pragma unroll 1
      for (int ob = 0; ob < n_ob; ++ob, ++oc) {
        const int sl = oc % S::NSLOT;
        const float bias = (p.epi == EPI_PRIMAL && p.b3) ? __ldg(p.b3 + 128 * ob + row) : 0.0f;
        mbar_wait(&bars[C_TFULL0 + sl], ptf[sl], err); ptf[sl] ^= 1; tc_fence_after();
        float* o = obase + 128 * ob;
#pragma unroll 1
        for (int pc = 0; pc < OCOLS / 32; ++pc) {
          float v[32];
          tmem_ld32(lt + F + 128 * sl + OCOLS * grp + 32 * pc, v);
          if (pc == OCOLS / 32 - 1) { tc_fence_before(); mbar_arrive(&bars[C_TEMPTY0 + sl]); }   // the slot may be refilled while we store
          const float4* inv4 = reinterpret_cast<const float4*>(INVS + OCOLS * grp + 32 * pc);
          const int qv = qn - 32 * pc;       // valid columns of this piece
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 iv = inv4[q4];
            const float is[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (4 * q4 + i < qv) *o = fmaf(v[4 * q4 + i], is[i], bias);
              o += p.ld_out;
            }
          }
        }
      }
      named_bar_sync(NB_ALL, S::EPI);                               // INVS / STAT may be rewritten by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EW) tmem_dealloc(tmem, S::TMEM_COLS);
}

}  // namespace tc
}  // namespace tib
