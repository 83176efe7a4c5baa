// train_gemm.cuh - general fp32-faithful GEMM on tcgen05 for the training step (forward, data gradients, weight gradients):
//
//     C[m][n] (=, +=, atomic +=)  alpha * sum_k A(m, k) * B(n, k)  (+ bias[n])
//
// Both operands are read straight from row-major fp32 tensors in global memory - there is no pre-packed weight image,
// because the weights change every optimiser step - and are turned into split-f16 operand images (tc_common.cuh) in
// shared memory by the CTA's own threads.  Each operand is described by (ptr, ld, idx, trans):
//     trans = 0:  Op(r, k) = ptr[row(r) * ld + k]      r indexes source rows (optionally gathered), k is contiguous
//     trans = 1:  Op(r, k) = ptr[row(k) * ld + r]      the contraction runs over source ROWS (optionally gathered)
// which covers the three GEMM forms of a Linear layer y = x W^T  (W [out][in] as the reference stores it):
//     forward          y [R][out]  : A = x (trans 0),   B = W (trans 0)            K = in
//     data gradient    dx [R][in]  : A = dy (trans 0),  B = W (trans 1)            K = out
//     weight gradient  dW [out][in]: A = dy (trans 1),  B = x (trans 1)            K = R   (split over CTAs, atomic)
// A trans-0 tile is built row-per-lane-octet (32 B per lane, 128 B per row and chunk); a trans-1 tile is built with
// lanes along r, so that every load instruction reads 128 contiguous bytes and every store writes 512.
//
// Range: f16 overflows at 65504 and loses relative precision below 2^-14, so every operand carries a power-of-two
// scale: a constant (states 2^-4, hidden activations / weights 1) or, for gradient tensors, one derived on the device
// from the tensor's |max| (tracked by the producing kernel with atomicMax), mapping it into [2^12, 2^13).  The scale is
// undone on the accumulators (exact).
//
// Pipeline: 256 threads build K chunks of 32 into a 2-stage ring (32 KB per stage): the global loads of chunk c + 1 are
// issued into registers before chunk c's barrier, thread 0 issues the three MMA passes of a chunk and commits them to
// the stage's mbarrier, and the next chunk is converted and stored while they run.  The accumulator tile leaves through
// warp-private pieces of the idle ring (coalesced 256-byte row segments; float4 atomics for the split-K weight
// gradients).  64 KB of shared memory and 128 TMEM columns per CTA, so three CTAs share an SM.  N, ldc, the bias and C
// must be multiples of 4 floats / 16-byte aligned (every Linear of the network is).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "train.cuh"

namespace tib {
namespace train {

struct GemmOperand {
  const float* ptr;
  long long ld;
  const int* idx;       // optional gather of source rows
  int trans;
  const float* amax;    // DEVICE |max| of the tensor (dynamic scale) or nullptr (use `scale`)
  float scale;
};

enum { GEMM_STORE = 0, GEMM_ATOMIC = 1, GEMM_ACCUM = 2 };

struct GemmP {
  int M, N, K;
  GemmOperand A, B;
  float* C;
  long long ldc;
  const int* c_idx;     // optional: result row m is added to row c_idx[m] of C (GEMM_ATOMIC only)
  const float* bias;    // [N] or nullptr
  int mode;             // GEMM_*
  int k_splits;         // == gridDim.z; > 1 requires GEMM_ATOMIC
  float alpha;
  int passes;           // 3 = split-f16 x3, 1 = single f16 pass
  int* err;
  long long* dbg;       // optional: clock64() stamps of CTA (0,0,0) thread 0 (pipeline diagnostics), else nullptr
  // optional pre-packed B operand (k_pack_operand): the split-f16 images of B in (n tile, K chunk) order, 16 KB each.  The
  // chunk then arrives by one bulk copy of the TMA unit instead of 256 threads loading, converting and storing it - used
  // for the weights, whose images are made once per optimiser step and shared by every m tile of every GEMM that reads them.
  const unsigned char* b_img;
  int b_chunk0, b_chunks;     // first K chunk of this GEMM inside the image, K chunks per n tile of the image
  // optional second A segment: K chunks >= a_split read A2 (its columns restart at 0) - cat[x1, x2] W^T as one GEMM
  GemmOperand A2;
  int a_split;                // in chunks; >= the chunk count when there is no second segment
  // optional fused epilogue over whole rows (N <= 128: one n tile, no K split, GEMM_STORE):
  //   EPI_LN_FWD: z = acc + bias -> C = n = (z - mean) / sqrt(var + 1e-5), ln_rstd[m], ln_h = SiLU(gamma n + beta)
  // (the LayerNorm ADJOINT was fused the same way - dz and dpre from the data-gradient GEMM's epilogue, the parameter sums on
  // the side stream - and measured slower than the separate kernel, 5.21 against 4.96 ms per step: its row-per-thread reads of
  // the saved n and the exp / divide chain sit badly in a 3-CTA-per-SM GEMM; removed)
  int epi;
  const float *ln_gamma, *ln_beta;
  float *ln_rstd, *ln_h;
};
enum { EPI_NONE = 0, EPI_LN_FWD = 1 };

constexpr int kGemmThreads = 256;
constexpr int kGemmKC = 32;                                   // K per chunk
constexpr int kGemmStages = 2;
constexpr uint32_t kGemmHalf = tc::kChunkHalfBytes;           // 8 KB: one [128 x 32] f16 image
constexpr uint32_t kGemmStageBytes = 4 * kGemmHalf;           // A hi, A lo, B hi, B lo
constexpr uint32_t kGemmEpiFloats = 3 * 128 + 4 * 128;     // bias / gamma / beta rows + two pairs of per-row partial sums
constexpr uint32_t kGemmSmem = kGemmStages * kGemmStageBytes + 64 + kGemmEpiFloats * 4;

// power-of-two scale that maps amax into [2^12, 2^13); 1 for an all-zero (or non-finite) tensor
__device__ __forceinline__ float pow2_scale_for(float amax) {
  if (!(amax > 0.0f) || !(amax < 3.0e38f)) return 1.0f;
  int ex;
  frexpf(amax, &ex);                       // amax = f * 2^ex, f in [0.5, 1)
  int p = 13 - ex;
  p = p > 100 ? 100 : (p < -100 ? -100 : p);
  return ldexpf(1.0f, p);
}
__device__ __forceinline__ float operand_scale(const GemmOperand& o) { return o.amax ? pow2_scale_for(__ldg(o.amax)) : o.scale; }

// One [128 x 32] operand tile (rows r0.., columns k0..) in two steps, so that the global loads of the NEXT chunk are in
// flight while the current one is converted, stored and multiplied:
//   load_tile : 16 floats per thread into registers (trans 0: 2 x 32 B of two rows; trans 1: 2 column groups x 8 source rows)
//   store_tile: scale, split into hi / lo f16 and write the images at `img` / `img + kGemmHalf`
struct TileRows { long long off[2]; bool ok[2]; };          // trans 0: the thread's two source rows (fixed over the K loop)

__device__ __forceinline__ TileRows tile_rows(const GemmOperand& o, int r0, int r_end, int warp, int lane) {
  TileRows t;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + 16 * warp + 8 * half + (lane & 7);
    t.ok[half] = o.trans == 0 && r < r_end;
    t.off[half] = t.ok[half] ? (o.idx ? (long long)__ldg(o.idx + r) : (long long)r) * o.ld : 0;
  }
  return t;
}

__device__ __forceinline__ void load_tile(float (&v)[16], const GemmOperand& o, const TileRows& tr, int r0, int r_end, int k0,
                                          int k_end, int warp, int lane) {
  if (o.trans == 0) {
    const int k = k0 + 8 * (lane >> 3);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (tr.ok[half] && k < k_end) {
        const float4* src = reinterpret_cast<const float4*>(o.ptr + tr.off[half] + k);
        a = __ldg(src);
        b = __ldg(src + 1);
      }
      v[8 * half + 0] = a.x; v[8 * half + 1] = a.y; v[8 * half + 2] = a.z; v[8 * half + 3] = a.w;
      v[8 * half + 4] = b.x; v[8 * half + 5] = b.y; v[8 * half + 6] = b.z; v[8 * half + 7] = b.w;
    }
  } else {
    // lanes along r; rows beyond the tile are clamped (their values are zeroed by store_tile's caller via `rok`), full
    // chunks take a branch-free path with one IMAD.WIDE per address
    const int r = min(r0 + 32 * (warp & 3) + lane, r_end - 1);
    const int kb = k0 + 16 * (warp >> 2);
    const char* const base = reinterpret_cast<const char*>(o.ptr + r);
    const long long ldb = o.ld * 4;
    if (k0 + kGemmKC <= k_end) {
      if (o.idx) {
        int rows[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) rows[i] = __ldg(o.idx + kb + i);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldg(reinterpret_cast<const float*>(base + rows[i] * ldb));
      } else {
        const char* q = base + kb * ldb;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldg(reinterpret_cast<const float*>(q + i * ldb));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = kb + i;
        v[i] = 0.0f;
        if (k < k_end) v[i] = __ldg(reinterpret_cast<const float*>(base + (o.idx ? (long long)__ldg(o.idx + k) : (long long)k) * ldb));
      }
    }
  }
}

__device__ __forceinline__ void store_tile(unsigned char* img, int trans, float scale, int warp, int lane, const float (&v)[16]) {
  // scale == 0 marks a thread whose trans-1 row lies beyond the tile (its clamped loads are discarded)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = v[8 * h + i] * scale;
    if (trans == 0) tc::store_group(img, kGemmHalf, 16 * warp + 8 * h + (lane & 7), lane >> 3, w);
    else tc::store_group(img, kGemmHalf, 32 * (warp & 3) + lane, 2 * (warp >> 2) + h, w);
  }
}

__global__ void __launch_bounds__(kGemmThreads, 3) k_gemm_tc(const GemmP p) {
  using namespace tc;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);   // [0..1] stage free, [2] done,
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);                          // [3..4] packed B chunk arrived
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;
  const bool packed_b = p.b_img != nullptr;
  pdl_trigger();                             // the next kernel of the stream may be scheduled (it waits at its own griddepcontrol.wait)

  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128;
  const int chunks_total = (p.K + kGemmKC - 1) / kGemmKC;
  const int cps = (chunks_total + p.k_splits - 1) / p.k_splits;
  const int c_begin = blockIdx.z * cps, c_end = min(chunks_total, c_begin + cps);
  if (c_begin >= c_end) return;                                   // uniform over the CTA

  const bool dbg = p.dbg != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int di = 0;
  if (dbg) p.dbg[di++] = clock64();
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1); mbar_init(&bars[4], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  pdl_wait();                                // barriers and TMEM are set up under the previous kernel's tail; its results are needed from here
  // the first chunk's operands are requested before the allocation barrier
  const TileRows ra = tile_rows(p.A, m0, p.M, warp, lane);
  const bool two_seg = p.a_split < chunks_total;
  TileRows ra2{};
  if (two_seg) ra2 = tile_rows(p.A2, m0, p.M, warp, lane);
  const int k_split = p.a_split * kGemmKC;                      // columns of the first segment
  auto load_a = [&](float (&v)[16], int c) {
    if (c < p.a_split) load_tile(v, p.A, ra, m0, p.M, c * kGemmKC, two_seg ? k_split : p.K, warp, lane);
    else load_tile(v, p.A2, ra2, m0, p.M, (c - p.a_split) * kGemmKC, p.K - k_split, warp, lane);
  };
  TileRows rb{};
  float va[16], vb[16];
  load_a(va, c_begin);
  if (!packed_b) {
    rb = tile_rows(p.B, n0, p.N, warp, lane);
    load_tile(vb, p.B, rb, n0, p.N, c_begin * kGemmKC, p.K, warp, lane);
  }
  const float sa = operand_scale(p.A), sb = operand_scale(p.B);
  // a trans-1 thread whose row lies beyond the tile loads a clamped row and stores zeros
  const float sa_t = (p.A.trans && m0 + 32 * (warp & 3) + lane >= p.M) ? 0.0f : sa;
  const float sb_t = (p.B.trans && n0 + 32 * (warp & 3) + lane >= p.N) ? 0.0f : sb;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (dbg) p.dbg[di++] = clock64();            // [1] after TMEM allocation

  for (int c = c_begin; c < c_end; ++c) {
    const int it = c - c_begin, st = it & 1;
    if (it >= kGemmStages) mbar_wait(&bars[st], ((it >> 1) - 1) & 1, err);      // the MMAs that read this stage are done
    unsigned char* const stage = smem + st * kGemmStageBytes;
    if (packed_b && tid == 0) {                // the B chunk of this stage: one 16 KB bulk copy, under the A conversion
      mbar_arrive_expect_tx(&bars[3 + st], 2 * kGemmHalf);
      bulk_g2s(stage + 2 * kGemmHalf, p.b_img + ((size_t)blockIdx.x * p.b_chunks + p.b_chunk0 + c) * (2 * kGemmHalf), 2 * kGemmHalf, &bars[3 + st]);
    }
    store_tile(stage, c < p.a_split ? p.A.trans : p.A2.trans, sa_t, warp, lane, va);
    if (!packed_b) store_tile(stage + 2 * kGemmHalf, p.B.trans, sb_t, warp, lane, vb);
    if (c + 1 < c_end) {                       // next chunk's loads fly under the barrier, the MMAs and the next wait
      load_a(va, c + 1);
      if (!packed_b) load_tile(vb, p.B, rb, n0, p.N, (c + 1) * kGemmKC, p.K, warp, lane);
    }
    if (dbg && di < 40) p.dbg[di++] = clock64();          // after the build of this chunk (thread 0's part)
    fence_proxy_async();
    __syncthreads();
    if (dbg && di < 40) p.dbg[di++] = clock64();          // after the barrier
    if (tid == 0) {
      if (packed_b) mbar_wait(&bars[3 + st], (it >> 1) & 1, err);
      tc_fence_after();
      const uint32_t a = smem_u32(stage), b = a + 2 * kGemmHalf;
      mma_f16x3(tmem, a, kGemmHalf, b, kGemmHalf, kGemmKC / 16, it > 0, p.passes);
      tc_commit(&bars[st]);
      if (c + 1 == c_end) tc_commit(&bars[2]);
    }
  }
  if (dbg) p.dbg[di++] = clock64();            // all chunks issued
  mbar_wait(&bars[2], 0, err);
  tc_fence_after();
  if (dbg) p.dbg[di++] = clock64();            // accumulator complete

  if (p.epi != EPI_NONE) {
    // ---- fused row epilogue (LayerNorm + SiLU forward): thread = (tile row, column half); the two halves of a row
    // exchange their partial sums through shared memory; the accumulators are simply re-read from TMEM for every pass
    float* const xs = reinterpret_cast<float*>(smem + kGemmStages * kGemmStageBytes + 64);
    float* const bias_s = xs, *const gam_s = xs + 128, *const bet_s = xs + 256, *const part = xs + 384;     // part[4][128]
    if (tid < 128) {
      const bool ok = tid < p.N;
      bias_s[tid] = (ok && p.bias) ? __ldg(p.bias + tid) : 0.0f;
      gam_s[tid] = ok ? __ldg(p.ln_gamma + tid) : 0.0f;
      bet_s[tid] = ok ? __ldg(p.ln_beta + tid) : 0.0f;
    }
    __syncthreads();
    const int hcol = warp >> 2, row = 32 * (warp & 3) + lane, m = m0 + row;
    const uint32_t tbase = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 64 * hcol;
    const float unscale = p.alpha / (sa * sb), inv_n = 1.0f / (float)p.N;
    float4* const tile = reinterpret_cast<float4*>(smem + warp * 8192);
    const int j = lane & 15;
    // coalesced copy of the warp's staged [32 x 64] block to rows of `dst` (leading dimension ldc)
    auto flush = [&](float* dst) {
      __syncwarp();
      const int n = 64 * hcol + 4 * j;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int r = 2 * i + (lane >> 4), mm = m0 + 32 * (warp & 3) + r;
        if (mm < p.M && n < p.N) *reinterpret_cast<float4*>(dst + (long long)mm * p.ldc + n) = tile[r * 16 + (j ^ (r & 15))];
      }
      __syncwarp();
    };
    {
      float sum = 0.0f;
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        float v[32];
        tmem_ld32(tbase + 32 * blk, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) { const int col = 64 * hcol + 32 * blk + i; if (col < p.N) sum += v[i] * unscale + bias_s[col]; }
      }
      part[hcol * 128 + row] = sum;
      __syncthreads();
      const float mean = (part[row] + part[128 + row]) * inv_n;
      float sq = 0.0f;
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        float v[32];
        tmem_ld32(tbase + 32 * blk, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = 64 * hcol + 32 * blk + i;
          if (col < p.N) { const float d = (v[i] * unscale + bias_s[col]) - mean; sq += d * d; }
        }
      }
      part[256 + hcol * 128 + row] = sq;
      __syncthreads();
      const float rs = 1.0f / sqrtf((part[256 + row] + part[384 + row]) * inv_n + 1e-5f);
      if (hcol == 0 && m < p.M) p.ln_rstd[m] = rs;
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        float v[32];
        tmem_ld32(tbase + 32 * blk, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = ((v[4 * q + e] * unscale + bias_s[64 * hcol + 32 * blk + 4 * q + e]) - mean) * rs;
          tile[lane * 16 + ((8 * blk + q) ^ (lane & 15))] = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      flush(p.C);                                              // n
#pragma unroll 4
      for (int sl = 0; sl < 16; ++sl) {                        // h = SiLU(gamma n + beta), own row, in place
        float4 x = tile[lane * 16 + (sl ^ (lane & 15))];
        const int col = 64 * hcol + 4 * sl;
        x.x = tc::silu_fast(fmaf(x.x, gam_s[col], bet_s[col]));         x.y = tc::silu_fast(fmaf(x.y, gam_s[col + 1], bet_s[col + 1]));
        x.z = tc::silu_fast(fmaf(x.z, gam_s[col + 2], bet_s[col + 2])); x.w = tc::silu_fast(fmaf(x.w, gam_s[col + 3], bet_s[col + 3]));
        tile[lane * 16 + (sl ^ (lane & 15))] = x;
      }
      flush(p.ln_h);
    }
  } else {
  // ---- epilogue.  Warp w owns TMEM lanes 32 (w & 3) .. +31 (rows) and columns 64 (w >> 2) .. +63; its [32 x 64] block goes
  // through a warp-private 8 KB piece of the (now idle) operand ring so that global rows are written 256 contiguous bytes at
  // a time: element (row r, float4 slot j) sits at r * 256 + ((j ^ (r & 15)) * 16) - conflict-free both ways.
  {
    float4* const tile = reinterpret_cast<float4*>(smem + warp * 8192);
    const float unscale = p.alpha / (sa * sb);
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 64 * (warp >> 2) + 32 * blk, v);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        tile[lane * 16 + ((8 * blk + q) ^ (lane & 15))] =
            make_float4(v[4 * q] * unscale, v[4 * q + 1] * unscale, v[4 * q + 2] * unscale, v[4 * q + 3] * unscale);
    }
    __syncwarp();
    const int j = lane & 15;
    const int n = n0 + 64 * (warp >> 2) + 4 * j;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr && blockIdx.z == 0 && n < p.N) bias = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int r = 2 * i + (lane >> 4);
      const int m = m0 + 32 * (warp & 3) + r;
      if (m < p.M && n < p.N) {
        float4 x = tile[r * 16 + (j ^ (r & 15))];
        x.x += bias.x; x.y += bias.y; x.z += bias.z; x.w += bias.w;
        const long long crow = p.c_idx ? (long long)__ldg(p.c_idx + m) : (long long)m;
        float4* dst = reinterpret_cast<float4*>(p.C + crow * p.ldc + n);
        if (p.mode == GEMM_STORE) *dst = x;
        else if (p.mode == GEMM_ACCUM) { const float4 o = *dst; *dst = make_float4(o.x + x.x, o.y + x.y, o.z + x.z, o.w + x.w); }
        else atomicAdd(dst, x);
      }
    }
  }
  }
  if (dbg) p.dbg[di++] = clock64();            // epilogue done
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
  if (dbg) { p.dbg[di++] = clock64(); p.dbg[63] = di; }
}

// ---- operand images in global memory --------------------------------------------------------------------------------------------
// One entry = one matrix W [rows_o][cols_i] (leading dimension ld, a block of the flat weight vector) packed twice:
//   fwd : B(n, k) = W[n][k]   (forward GEMMs;       n tiles over rows_o, K chunks over cols_i)   at img + fwd_off
//   tr  : B(n, k) = W[k][n]   (data-gradient GEMMs; n tiles over cols_i, K chunks over rows_o)   at img + tr_off
// Chunk (n tile, K chunk) is 16 KB (hi image then lo image) at (n tile * n_chunks + chunk) * 16 KB; rows / columns beyond
// the matrix are zero.
struct PackEntry { long long src; int ld, rows_o, cols_i; long long fwd_off, tr_off; int first_block; int pad; };
constexpr int kPackMaxEntries = 96;
struct PackTable { PackEntry e[kPackMaxEntries]; int n; int total_blocks; };

__global__ void __launch_bounds__(kGemmThreads) k_pack_operand(const PackTable tab, const float* __restrict__ W,
                                                                unsigned char* __restrict__ img) {
  pdl_entry();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.x;
  int ei = 0;
  while (ei + 1 < tab.n && tab.e[ei + 1].first_block <= blk) ++ei;          // tab.n <= 96: a short uniform scan
  const PackEntry& e = tab.e[ei];
  int local = blk - e.first_block;
  const int fwd_chunks = (e.cols_i + kGemmKC - 1) / kGemmKC, fwd_tiles = (e.rows_o + 127) / 128;
  const int tr_chunks = (e.rows_o + kGemmKC - 1) / kGemmKC;
  GemmOperand o{};
  o.ptr = W + e.src; o.ld = e.ld; o.idx = nullptr; o.scale = 1.0f; o.amax = nullptr;
  unsigned char* dst;
  int r0, r_end, k0, k_end;
  if (local < fwd_tiles * fwd_chunks) {
    const int nt = local / fwd_chunks, c = local % fwd_chunks;
    o.trans = 0; r0 = nt * 128; r_end = e.rows_o; k0 = c * kGemmKC; k_end = e.cols_i;
    dst = img + e.fwd_off + (size_t)local * (2 * kGemmHalf);
  } else {
    local -= fwd_tiles * fwd_chunks;
    const int nt = local / tr_chunks, c = local % tr_chunks;
    o.trans = 1; r0 = nt * 128; r_end = e.cols_i; k0 = c * kGemmKC; k_end = e.rows_o;
    dst = img + e.tr_off + (size_t)local * (2 * kGemmHalf);
  }
  const TileRows tr = tile_rows(o, r0, r_end, warp, lane);
  float v[16];
  load_tile(v, o, tr, r0, r_end, k0, k_end, warp, lane);
  const float sc = (o.trans && r0 + 32 * (warp & 3) + lane >= r_end) ? 0.0f : 1.0f;
  store_tile(dst, o.trans, sc, warp, lane, v);
}

}  // namespace train
}  // namespace tib
