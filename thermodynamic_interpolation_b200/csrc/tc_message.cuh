// tc_message.cuh - SE3Message.forward (cpainn.py:263-310) fused on the sm_100a tensor cores.
//
// Work unit: a TILE = up to kTileNodes consecutive destination nodes and all their incoming edges
// (<= 128 rows).  Edge features e[E][F] are kept in (dst, src)-lexicographic order inside the
// library, so the rows of a tile are contiguous in HBM and every destination node is owned by
// exactly one tile: the scatter-sum over incoming edges needs neither global atomics nor a second pass.
//
// Per tile (F = 128):
//   hidden layers   D[edge][feat] = A[edge][k] * W[feat][k]^T      (edge = TMEM lane: LayerNorm is row-local)
//       w   : PE(d) -> LN/SiLU -> LN/SiLU                          (2 GEMMs)
//       phi : cat[s[src], e] -> LN/SiLU -> LN/SiLU                 (3 GEMMs: the 2F input in two K halves)
//   output layer    D^T[feat][edge] = W3[feat][k] * H2[edge][k]^T  (feat = TMEM lane: the gated scatter
//       over the edges of a tile is a thread-local loop over TMEM columns), one 128-feature split
//       (gates, scale_edge_dir, ds, de, cross_gates) at a time, phi and w side by side, double-buffered
//       in TMEM so the MMAs of split i+1 overlap the scatter of split i.
// Weights stream from L2 through a ring of 16 KB chunks with cp.async.bulk (TMA unit).
// Not recomputed per layer / per edge: the positional-encoding operand image of a tile (stored by the first layer,
// bulk-copied into X by the others; the issuer then starts the next tile's first hidden GEMM while the epilogue still
// scatters the previous tile's last split) and, in the first layer, phi's hidden activations (a table per
// (embedding row, edge type), gathered per edge).
//
// Warp roles (576 threads): warps 0-15 = four epilogue groups (group g = warp / 4; a warp reads the
// TMEM lane quarter warp % 4), warp 16 weight producer (+ TMEM allocation), warp 17 MMA issuer.
//   hidden phase : every epilogue runs on all 16 warps (group g = feature columns [32g, 32g+32) of
//                  every row; LayerNorm statistics exchanged through shared memory); the w and phi
//                  chains are interleaved so that the MMAs of one run under the epilogue of the other.
//   output phase : group g owns whole destination nodes and scatters their TMEM columns (= edges).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace tib {
namespace tc {

constexpr int kF = 128;
constexpr int kEpiThreads = 512;
constexpr int kThreads = kEpiThreads + 64;
constexpr int kStages = 4;
// fp32 layer parameters staged in shared memory once per CTA:
//   [0,6F): w  {b1,g1,be1,b2,g2,be2}   [6F,12F): phi {b1,g1,be1,b2,g2,be2}   [12F,17F): phi b3   [17F,22F): w b3
constexpr int kPrmW = 0, kPrmPhi = 6 * 128, kPrmB3 = 12 * 128, kPrmFloats = 22 * 128;
constexpr int kTileNodes = 16;                 // destination nodes per tile (4 per epilogue group: delta-s / delta-v live in registers)
constexpr int kChunksPerLayer = 60;

// streamed chunk order of one message layer (every entry is 4 chunks = one [128 x 128] matrix):
//   0 w.W1 | 4 phi.W1[:, :F] | 8 w.W2 | 12 phi.W1[:, F:] | 16 phi.W2 | 20+8sp: phi.W3[sp] and w.W3[sp] interleaved
//   chunk by chunk (phi k0, w k0, phi k1, w k1, ...) so the two independent accumulators alternate in the MMA queue
struct MsgParams {
  const float* vec[12];   // w {b1,g1,be1,b2,g2,be2} then phi {b1,g1,be1,b2,g2,be2}: rows [0,12) of PRM
  const float *phi_b3, *w_b3;
};

struct TcMsgP {
  int n_nodes, n_tiles, nodes_per_tile;
  const int* tile_node_ptr; // [n_tiles+1] or NULL (uniform tiles of nodes_per_tile nodes)
  const int* node_in_ptr;   // [N+1] first (dst-major) edge row of each node
  const uint4* rowa;        // [E] RowA per (dst,src)-ordered edge row (k_edge_tables; slot filled per tile)
  const uint4* rowb;        // [E] RowB
  const float* s_old;       // [N][F]
  const float* v_old;       // [N][3][F]
  float* s_new;
  float* v_new;
  float* e;                 // [E][F] (dst,src) order, updated in place
  unsigned char* pe_img;    // optional [n_tiles][64 KB]: operand images of PositionalEncoder(edge_dist), identical in every layer
  int pe_mode;              //   0 compute per tile, 1 compute and store the image (first layer), 2 load the stored image
  const float* phi_tab;     // first layer, optional: phi's second hidden activation per (embedding row, edge type), [U * n_et][F]
  const int* embed_index;   //   row of each node in the de-duplicated embedding table (with phi_tab)
  int n_et;                 //   number of edge types (with phi_tab)
  const float* edge_emb;    // first layer only: e0 = edge_emb[edge type] is formed on the fly, never read from e (embedding.py:89-103)
  const unsigned char* wblob;   // kChunksPerLayer chunks of this layer
  MsgParams prm;
  float length_scale;
  int first_layer;
  int passes;               // 3 = split-f16 (fp32-faithful), 1 = single f16 pass
  int* err;
  long long* dbg;           // optional [gridDim.x][8] stall-cycle counters (diagnostics), or NULL
};

struct RowA { int src; int dst; int slot_last; float dist; };   // slot | last << 8 | first << 9 | edge type << 16
struct RowB { float dx, dy, dz, pad; };

struct MsgSmem {
  // offsets (bytes) into dynamic shared memory
  static constexpr uint32_t X = 0;
  static constexpr uint32_t Y = X + kOperandBytes;
  static constexpr uint32_t RING = Y + kOperandBytes;
  static constexpr uint32_t PRM = RING + kStages * kChunkBytes;                // layer parameters (fp32), see kPrm*
  // tile tables, two sets: the next tile's set is filled while this tile waits for its last hidden-layer MMAs
  static constexpr uint32_t ROWA = PRM + kPrmFloats * 4;                       // 2 x RowA[128]
  static constexpr uint32_t ROWB = ROWA + 2 * 128 * 16;                        // 2 x RowB[128]
  static constexpr uint32_t STAT = ROWB + 2 * 128 * 16;                        // 2 x float2 [4 groups][128 rows] (alternating)
  static constexpr uint32_t SLOTROW = STAT + 2 * 4 * 128 * 8;              // 2 x int[32]: first row of each slot (kTileNodes + 1 used)
  static constexpr uint32_t BARS = SLOTROW + 2 * 128;
  static constexpr uint32_t TOTAL = BARS + 256;
};
// barrier indices
enum { B_FULL = 0, B_EMPTY = B_FULL + kStages /* one per PAIR of stages */, B_XFULL = B_EMPTY + kStages / 2, B_YFULL, B_YFREE, B_ACC0, B_ACC1,
       B_TFULL0, B_TFULL1, B_TEMPTY0, B_TEMPTY1, B_XFREE, B_PEFULL, B_COUNT };
// named barriers: all 512 epilogue threads, or the 4 warps (one per column group) that share TMEM lane quarter wq
enum { NB_ALL = 1, NB_QUARTER = 2 /* + wq */ };
constexpr int kQuarterThreads = 128;

// The hot loops below are deliberately ROLLED (small bodies, TMEM re-read per pass): the straight-line
// version of this kernel was ~220 KB of SASS and ran instruction-fetch bound (16 warps streaming
// through code far larger than the 32 KB instruction cache).
__device__ __noinline__ void hidden_epilogue(uint32_t taddr, int grp, int row, const float* b, const float* g, const float* be,
                                             unsigned char* op, float* stat_f, float ascale) {
  // Accumulator row `row` (TMEM lane), columns [32*grp, +32): * ascale + bias -> LayerNorm over all 128 columns
  // (sum and sum of squares exchanged between the four threads of a row) -> SiLU -> operand image.
  float2* stat = reinterpret_cast<float2*>(stat_f);
  float v[32];
  tmem_ld32(taddr + 32 * grp, v);                       // one TMEM read; the row quarter stays in registers
  const float4* bp = reinterpret_cast<const float4*>(b + 32 * grp);
  float sum = 0.0f, ss = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 bb = bp[c];
    v[4 * c + 0] = fmaf(v[4 * c + 0], ascale, bb.x); v[4 * c + 1] = fmaf(v[4 * c + 1], ascale, bb.y);
    v[4 * c + 2] = fmaf(v[4 * c + 2], ascale, bb.z); v[4 * c + 3] = fmaf(v[4 * c + 3], ascale, bb.w);
    sum += (v[4 * c + 0] + v[4 * c + 1]) + (v[4 * c + 2] + v[4 * c + 3]);
    ss = fmaf(v[4 * c + 0], v[4 * c + 0], fmaf(v[4 * c + 1], v[4 * c + 1], fmaf(v[4 * c + 2], v[4 * c + 2], fmaf(v[4 * c + 3], v[4 * c + 3], ss))));
  }
  stat[grp * 128 + row] = make_float2(sum, ss);
  named_bar_sync(NB_QUARTER + (row >> 5), kQuarterThreads);   // only the four threads of a row exchange statistics
  const float2 s0 = stat[row], s1 = stat[128 + row], s2 = stat[256 + row], s3 = stat[384 + row];
  const float mean = ((s0.x + s1.x) + (s2.x + s3.x)) * (1.0f / 128.0f);
  const float var = fmaxf(((s0.y + s1.y) + (s2.y + s3.y)) * (1.0f / 128.0f) - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
  const float4* gp = reinterpret_cast<const float4*>(g + 32 * grp);
  const float4* ep = reinterpret_cast<const float4*>(be + 32 * grp);
#pragma unroll
  for (int kg = 0; kg < 4; ++kg) {
    const float4 g0 = gp[2 * kg], g1 = gp[2 * kg + 1], e0 = ep[2 * kg], e1 = ep[2 * kg + 1];
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = silu_fast(fmaf(fmaf(v[8 * kg + i], rstd, nmr), gg[i], ee[i]));
    store_group(op, kOperandHalfBytes, row, 4 * grp + kg, y);
  }
  // Consecutive calls alternate between two `stat` buffers; a buffer is reused only after an MMA that
  // needed every thread's operand-ready arrival has completed, i.e. after all of these reads.
}

// rows [32*wq, +32) x column groups [4*grp, +4) of an operand image from row-major fp32 global rows
// `base + row_index(r) * 128`, where row_index(r) = ROWA[r].src (gather 1), the edge type (gather 2: rows of
// the edge-type embedding, first layer), embed_index[src] * n_et + edge type (gather 3: rows of the first layer's
// phi table, scale 1) or row0 + r (gather 0).  Raw state rows are scaled by kStateScale.
__device__ __noinline__ void build_rows(unsigned char* op, int wq, int grp, int lane, int rows, const float* base,
                                        const RowA* rowa, int gather, int row0, float scale = kStateScale,
                                        const int* embed_index = nullptr, int n_et = 0) {
  const int g = 4 * grp + (lane >> 3);
  float4 a[4], b[4];
  // all eight 16-byte loads are issued before the first conversion (the build is bound by their latency)
#pragma unroll
  for (int oct = 0; oct < 4; ++oct) {
    const int r = 32 * wq + 8 * oct + (lane & 7);
    a[oct] = b[oct] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const size_t et = (size_t)((rowa[r].slot_last >> 16) & 0xFF);
      const size_t ri = gather == 1 ? (size_t)rowa[r].src : gather == 2 ? et
                        : gather == 3 ? (size_t)__ldg(embed_index + rowa[r].src) * n_et + et : (size_t)(row0 + r);
      const float* src = base + ri * kF + g * 8;
      a[oct] = __ldg(reinterpret_cast<const float4*>(src));
      b[oct] = __ldg(reinterpret_cast<const float4*>(src + 4));
    }
  }
#pragma unroll
  for (int oct = 0; oct < 4; ++oct) {
    const int r = 32 * wq + 8 * oct + (lane & 7);
    const float v[8] = {a[oct].x * scale, a[oct].y * scale, a[oct].z * scale, a[oct].w * scale,
                        b[oct].x * scale, b[oct].y * scale, b[oct].z * scale, b[oct].w * scale};
    store_group(op, kOperandHalfBytes, r, g, v);
  }
}

// One [128 x 128] matrix = 4 streamed chunks; transposed = the weights are the A operand.
struct MmaRing { int stage; uint32_t ph; long long w_weights; long long t_issue; long long t_commit; };
__device__ __forceinline__ void gemm_job(uint64_t* bars, uint32_t ring, MmaRing& st, uint32_t d, uint32_t op, int transposed,
                                      int accumulate, int passes, volatile int* err, bool diag) {
#pragma unroll 1
  for (int kb = 0; kb < 4; ++kb) {
    mbar_wait_timed(&bars[B_FULL + st.stage], st.ph, err, st.w_weights, diag);
    tc_fence_after();
    const uint32_t wst = ring + st.stage * kChunkBytes, opk = op + kb * (2 * kKStepBytes);
    const long long ta = diag ? clock64() : 0;
    if (!transposed) mma_f16x3(d, opk, kOperandHalfBytes, wst, kChunkHalfBytes, 2, accumulate || kb > 0, passes);
    else             mma_f16x3(d, wst, kChunkHalfBytes, opk, kOperandHalfBytes, 2, accumulate || kb > 0, passes);
    const long long tb = diag ? clock64() : 0;
    if (st.stage & 1) tc_commit(&bars[B_EMPTY + (st.stage >> 1)]);     // ring slots are released in pairs
    if (diag) { st.t_issue += tb - ta; st.t_commit += clock64() - tb; }
    if (++st.stage == kStages) { st.stage = 0; st.ph ^= 1; }
  }
}

// Output layer of one split: D_phi^T = W3phi * Hphi^T and D_w^T = W3w * Hw^T, chunk pairs (phi k, w k) in
// consecutive ring stages; the MMAs of the two independent accumulators are interleaved.
__device__ __forceinline__ void gemm_pair_job(uint64_t* bars, uint32_t ring, MmaRing& st, uint32_t d_phi, uint32_t op_phi,
                                           uint32_t d_w, uint32_t op_w, int passes, volatile int* err, bool diag) {
#pragma unroll 1
  for (int kb = 0; kb < 4; ++kb) {
    mbar_wait_timed(&bars[B_FULL + st.stage], st.ph, err, st.w_weights, diag);
    mbar_wait_timed(&bars[B_FULL + st.stage + 1], st.ph, err, st.w_weights, diag);
    tc_fence_after();
    const uint32_t wp = ring + st.stage * kChunkBytes, ww = wp + kChunkBytes;
    const uint32_t hp = op_phi + kb * (2 * kKStepBytes), hw = op_w + kb * (2 * kKStepBytes);
    const long long ta = diag ? clock64() : 0;
#pragma unroll 1
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t o = ks * kKStepBytes;
      const uint32_t acc = (kb | ks) ? 1u : 0u;
      const uint64_t wph = make_desc(wp + o), wpl = make_desc(wp + kChunkHalfBytes + o);
      const uint64_t wwh = make_desc(ww + o), wwl = make_desc(ww + kChunkHalfBytes + o);
      const uint64_t hph = make_desc(hp + o), hpl = make_desc(hp + kOperandHalfBytes + o);
      const uint64_t hwh = make_desc(hw + o), hwl = make_desc(hw + kOperandHalfBytes + o);
      tc_mma_f16(d_phi, wph, hph, kIdesc128x128, acc);
      tc_mma_f16(d_w, wwh, hwh, kIdesc128x128, acc);
      if (passes == 3) {
        tc_mma_f16(d_phi, wph, hpl, kIdesc128x128, 1u);
        tc_mma_f16(d_w, wwh, hwl, kIdesc128x128, 1u);
        tc_mma_f16(d_phi, wpl, hph, kIdesc128x128, 1u);
        tc_mma_f16(d_w, wwl, hwh, kIdesc128x128, 1u);
      }
    }
    const long long tb = diag ? clock64() : 0;
    tc_commit(&bars[B_EMPTY + (st.stage >> 1)]);
    if (diag) { st.t_issue += tb - ta; st.t_commit += clock64() - tb; }
    st.stage += 2;
    if (st.stage == kStages) { st.stage = 0; st.ph ^= 1; }
  }
}

// DIAG = the stall-cycle counters of tib_debug_counters; the production instantiation carries none of their registers.
template <bool DIAG>
__global__ void __launch_bounds__(kThreads, 1) k_message_tc(TcMsgP p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const X = smem + MsgSmem::X;
  unsigned char* const Y = smem + MsgSmem::Y;
  unsigned char* const RING = smem + MsgSmem::RING;
  float* const PRM = reinterpret_cast<float*>(smem + MsgSmem::PRM);
  int* const SLOTROW_b = reinterpret_cast<int*>(smem + MsgSmem::SLOTROW);
  RowA* const ROWA_b = reinterpret_cast<RowA*>(smem + MsgSmem::ROWA);
  RowB* const ROWB_b = reinterpret_cast<RowB*>(smem + MsgSmem::ROWB);
  float* const STAT = reinterpret_cast<float*>(smem + MsgSmem::STAT);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + MsgSmem::BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + MsgSmem::BARS + 8 * B_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* err = p.err;
  pdl_trigger();        // programmatic dependent launch: the next kernel of the stream may run its prologue now
  constexpr bool diag = DIAG;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&bars[B_FULL + i], 1);
    for (int i = 0; i < kStages / 2; ++i) mbar_init(&bars[B_EMPTY + i], 1);
    mbar_init(&bars[B_XFULL], kEpiThreads); mbar_init(&bars[B_YFULL], kEpiThreads);
    mbar_init(&bars[B_YFREE], 1); mbar_init(&bars[B_ACC0], 1); mbar_init(&bars[B_ACC1], 1);
    mbar_init(&bars[B_TFULL0], 1); mbar_init(&bars[B_TFULL1], 1);
    mbar_init(&bars[B_TEMPTY0], kEpiThreads); mbar_init(&bars[B_TEMPTY1], kEpiThreads);
    mbar_init(&bars[B_XFREE], 1); mbar_init(&bars[B_PEFULL], 1);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  // PRM rows [0,12) are the 12 vectors of MsgParams in kPrmRow order, rows [12,17) phi b3, rows [17,22) w b3; the pointer
  // table is read from the kernel-parameter bank (a local array of pointers would live on the stack)
  for (int i = tid; i < kPrmFloats; i += kThreads) {
    const int r = i >> 7;
    const float* src = r < 12 ? p.prm.vec[r] : (r < 17 ? p.prm.phi_b3 + (r - 12) * kF : p.prm.w_b3 + (r - 17) * kF);
    PRM[i] = __ldg(src + (i & 127));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();           // barriers, TMEM and parameters were set up under the previous kernel; its results are needed from here
  const uint32_t tmem = *tmem_slot;
  const int n_splits = p.first_layer ? 3 : 5;
  // First layer with a phi table: s0 and e0 take a handful of distinct values, so phi's hidden layers were evaluated
  // once per (embedding row, edge type) (k_phi_table) and the phi chain of every tile is a gather.
  const bool phi_tab = p.first_layer && p.phi_tab != nullptr;
  // The positional encoding of the edge distances is the same in all layers of one drift evaluation: the first layer
  // stores its operand image per tile, the later layers have the bulk-copy unit drop it into X as soon as the previous
  // tile's last output MMAs have released X - that is during the previous tile's last scatter, off the critical path.
  const bool pe_load = p.pe_img != nullptr && p.pe_mode == 2;
  const bool pe_store = p.pe_img != nullptr && p.pe_mode == 1;
  const int first_split_chunk = 20 + 8 * (p.first_layer ? 1 : 0);   // first chunk of the first output split in the blob

  if (warp == 16) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      long long w_empty = 0;
      uint32_t pxf = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        if (pe_load) {
          if (tile != (int)blockIdx.x) { mbar_wait(&bars[B_XFREE], pxf, err); pxf ^= 1; }   // previous tile's output MMAs are done with X
          mbar_arrive_expect_tx(&bars[B_PEFULL], kOperandBytes);
          bulk_g2s(X, p.pe_img + (size_t)tile * kOperandBytes, kOperandBytes, &bars[B_PEFULL]);
        }
        for (int c = 0; c < kChunksPerLayer; ++c) {
          if (p.first_layer && ((c >= 20 && c < 28) || c >= 52)) continue;   // splits 0 and 4 multiply v = 0
          if (phi_tab && ((c >= 4 && c < 8) || (c >= 12 && c < 20))) continue;  // phi W1a, W1b, W2
          // the first output split is issued as w k0..k3 then phi k0..k3 (its w half starts before phi's operand is final);
          // the blob holds every split as (phi k, w k) pairs
          int ci = c;
          if (c >= first_split_chunk && c < first_split_chunk + 8) {
            const int j = c - first_split_chunk;
            ci = first_split_chunk + (j < 4 ? 2 * j + 1 : 2 * (j - 4));
          }
          if (!(stage & 1)) mbar_wait_timed(&bars[B_EMPTY + (stage >> 1)], ph ^ 1, err, w_empty, diag);
          mbar_arrive_expect_tx(&bars[B_FULL + stage], kChunkBytes);
          bulk_g2s(RING + stage * kChunkBytes, p.wblob + (size_t)ci * kChunkBytes, kChunkBytes, &bars[B_FULL + stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
      }
      if (diag) p.dbg[blockIdx.x * 8 + 2] = w_empty;
    }
  } else if (warp == 17) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t px = 0, py = 0, ppe = 0, pte[2] = {0, 0};
      long long w_operands = 0, w_tempty = 0;
      const long long t_start = clock64();
      const uint32_t xa = smem_u32(X), ya = smem_u32(Y), ring = smem_u32(RING);
      MmaRing rs{0, 0u, 0, 0, 0};
      auto gemm = [&](uint32_t d, uint32_t op, bool transposed, bool accumulate) {
        gemm_job(bars, ring, rs, d, op, transposed, accumulate, p.passes, err, diag);
      };
      const uint32_t acc0 = tmem, acc1 = tmem + 128;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        if (pe_load) {
          // X arrives by bulk copy, so this wait says nothing about the epilogue of the previous tile: the hidden
          // accumulators alias output slot 0, whose last readers (split n_splits - 2) must have drained it
          mbar_wait_timed(&bars[B_PEFULL], ppe, err, w_operands, diag); ppe ^= 1;
          mbar_wait_timed(&bars[B_TEMPTY0], pte[0] ^ 1, err, w_tempty, diag);
        } else {
          mbar_wait_timed(&bars[B_XFULL], px, err, w_operands, diag); px ^= 1;
        }
        tc_fence_after();
        gemm(acc0, xa, false, false);                       // w layer 1   : PE(d)
        tc_commit(&bars[B_ACC0]);
        if (!phi_tab) {
          mbar_wait_timed(&bars[B_YFULL], py, err, w_operands, diag); py ^= 1; tc_fence_after();
          gemm(acc1, ya, false, false);                     // phi layer 1 : s[src] half
          tc_commit(&bars[B_YFREE]);
        }
        mbar_wait_timed(&bars[B_XFULL], px, err, w_operands, diag); px ^= 1; tc_fence_after();
        gemm(acc0, xa, false, false);                       // w layer 2
        tc_commit(&bars[B_ACC0]);
        if (!phi_tab) {
          mbar_wait_timed(&bars[B_YFULL], py, err, w_operands, diag); py ^= 1; tc_fence_after();
          gemm(acc1, ya, false, true);                      // phi layer 1 : e half (accumulates)
          tc_commit(&bars[B_ACC1]);
          mbar_wait_timed(&bars[B_YFULL], py, err, w_operands, diag); py ^= 1; tc_fence_after();
          gemm(acc1, ya, false, false);                     // phi layer 2
          tc_commit(&bars[B_ACC1]);
        }
        // first output split: its w half needs only X (final since E5) and runs under the last phi epilogues
        mbar_wait_timed(&bars[B_XFULL], px, err, w_operands, diag); px ^= 1; tc_fence_after();
        {
          const int pb = 1;
          mbar_wait_timed(&bars[B_TEMPTY0 + pb], pte[pb] ^ 1, err, w_tempty, diag); pte[pb] ^= 1; tc_fence_after();
          gemm(tmem + 256 * pb + 128, xa, true, false);     // w layer 3, transposed
          mbar_wait_timed(&bars[B_YFULL], py, err, w_operands, diag); py ^= 1; tc_fence_after();   // final phi operand
          gemm(tmem + 256 * pb, ya, true, false);           // phi layer 3, transposed
          tc_commit(&bars[B_TFULL0 + pb]);
        }
        for (int it = 1; it < n_splits; ++it) {
          const int pb = (it + 1) & 1;   // the LAST split uses slot 1: slot 0 (= the hidden accumulators) drains one split early
          mbar_wait_timed(&bars[B_TEMPTY0 + pb], pte[pb] ^ 1, err, w_tempty, diag); pte[pb] ^= 1; tc_fence_after();
          gemm_pair_job(bars, ring, rs, tmem + 256 * pb, ya, tmem + 256 * pb + 128, xa, p.passes, err, diag);
          tc_commit(&bars[B_TFULL0 + pb]);                  // phi and w layer 3 of this split (transposed)
        }
        tc_commit(&bars[B_XFREE]);                          // X (and Y) may be overwritten for the next tile
      }
      if (diag) {
        p.dbg[blockIdx.x * 8 + 0] = rs.w_weights; p.dbg[blockIdx.x * 8 + 1] = w_operands;
        p.dbg[blockIdx.x * 8 + 3] = w_tempty; p.dbg[blockIdx.x * 8 + 4] = clock64() - t_start;
        p.dbg[blockIdx.x * 8 + 5] = rs.t_issue; p.dbg[blockIdx.x * 8 + 6] = rs.t_commit;
      }
    }
  } else {
    // =========================== builders / epilogue (512 threads) ===========================
    const int grp = warp >> 2, wq = warp & 3;
    const int row = 32 * wq + lane;                         // TMEM lane: edge row (hidden) / feature (output)
    uint32_t pacc0 = 0, pacc1 = 0, pyf = 0, ptf[2] = {0, 0};
    long long w_acc = 0, w_tfull = 0;
    long long phc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#define TIB_PHASE(i) do { if (diag) { const long long _t = clock64(); phc[i] += _t - tlast; tlast = _t; } } while (0)
    const uint32_t lane_taddr = tmem + ((uint32_t)(wq * 32) << 16);
    // bounds of a tile and its tables (set `par`): RowA / RowB of its edge rows and the first row of every node slot
    int nx_node_lo = 0, nx_node_hi = 0, nx_row0 = 0, nx_rows = 0;
    auto load_tables = [&](int tile, int par) {
      nx_node_lo = p.tile_node_ptr ? __ldg(p.tile_node_ptr + tile) : tile * p.nodes_per_tile;
      nx_node_hi = p.tile_node_ptr ? __ldg(p.tile_node_ptr + tile + 1) : min(nx_node_lo + p.nodes_per_tile, p.n_nodes);
      nx_row0 = __ldg(p.node_in_ptr + nx_node_lo);
      nx_rows = __ldg(p.node_in_ptr + nx_node_hi) - nx_row0;
      if (tid < 128) {
        uint4 ra = make_uint4(0u, (uint32_t)nx_node_lo, 0u, 0u), rb = make_uint4(0u, 0u, 0u, 0u);
        if (tid < nx_rows) {
          ra = __ldg(p.rowa + nx_row0 + tid);
          rb = __ldg(p.rowb + nx_row0 + tid);
          ra.z |= (uint32_t)((int)ra.y - nx_node_lo);        // slot of the destination node inside this tile
        }
        reinterpret_cast<uint4*>(ROWA_b + par * 128)[tid] = ra;
        reinterpret_cast<uint4*>(ROWB_b + par * 128)[tid] = rb;
      }
      if (tid >= 128 && tid <= 128 + kTileNodes && nx_node_lo + (tid - 128) <= nx_node_hi)
        SLOTROW_b[par * 32 + tid - 128] = __ldg(p.node_in_ptr + nx_node_lo + (tid - 128)) - nx_row0;
    };
    if ((int)blockIdx.x < p.n_tiles) load_tables(blockIdx.x, 0);
    named_bar_sync(NB_ALL, kEpiThreads);
    int par = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, par ^= 1) {
      // this tile's tables were filled during the previous tile (or above); the end-of-tile barrier ordered them
      const int node_lo = nx_node_lo, node_hi = nx_node_hi, row0 = nx_row0, rows = nx_rows;
      const RowA* const ROWA = ROWA_b + par * 128;
      const RowB* const ROWB = ROWB_b + par * 128;
      const int* const SLOTROW = SLOTROW_b + par * 32;
      // The e rows of this tile (one contiguous block, 4 lines per row) are read ~20 k cycles from now (E4) and again in
      // the e += de split: start them towards L2 now, so that those loads do not pay the HBM latency.
      if (!p.first_layer && tid < 4 * rows)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.e + (size_t)row0 * kF + (size_t)tid * 32));
      TIB_PHASE(0);   // tile tables

      // ---- hidden phase: one sequence, every epilogue on all 16 warps (group g = feature columns
      // [32g, 32g+32) of every row); the MMAs of one chain run under the epilogue of the other.
      {
        // E1: PositionalEncoder(edge_dist) -> X                                     (cpainn.py:283)
        if (!pe_load) {
          const float dist = ROWA[row].dist;
          unsigned char* const img = pe_store ? p.pe_img + (size_t)tile * kOperandBytes : nullptr;
#pragma unroll 2
          for (int kg = 4 * grp; kg < 4 * grp + 4; ++kg) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float sn = 0.0f, cs = 0.0f;
              if (row < rows) sincos_cw(pe_arg(dist, p.length_scale, 4 * kg + q + 1), sn, cs);
              v[2 * q] = cs; v[2 * q + 1] = sn;
            }
            store_group(X, kOperandHalfBytes, row, kg, v);
            if (img) store_group(img, kOperandHalfBytes, row, kg, v);
          }
          fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
        }
        if (phi_tab) {
          // E2': phi hidden 2 of every edge is a row of the table -> Y (final)
          build_rows(Y, wq, grp, lane, rows, p.phi_tab, ROWA, 3, 0, 1.0f, p.embed_index, p.n_et);
          fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
          TIB_PHASE(1);
          // E3: w hidden 1 -> X
          mbar_wait_timed(&bars[B_ACC0], pacc0, err, w_acc, diag); pacc0 ^= 1; tc_fence_after();
          hidden_epilogue(lane_taddr, grp, row, PRM + kPrmW, PRM + kPrmW + kF, PRM + kPrmW + 2 * kF, X, STAT, 1.0f);
          tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
          if (tile + (int)gridDim.x < p.n_tiles) load_tables(tile + gridDim.x, par ^ 1);   // under the MMAs of w layer 2
          TIB_PHASE(2);
          // E5: w hidden 2 -> X (final)
          mbar_wait_timed(&bars[B_ACC0], pacc0, err, w_acc, diag); pacc0 ^= 1; tc_fence_after();
          hidden_epilogue(lane_taddr, grp, row, PRM + kPrmW + 3 * kF, PRM + kPrmW + 4 * kF, PRM + kPrmW + 5 * kF, X, STAT + 1024, 1.0f);
          tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
          TIB_PHASE(3);
        } else {
        // E2: s[src] -> Y                                                            (cpainn.py:275-281)
        build_rows(Y, wq, grp, lane, rows, p.s_old, ROWA, 1, 0);
        fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
        TIB_PHASE(1);   // E1 + E2
        // E3: w hidden 1 -> X
        mbar_wait_timed(&bars[B_ACC0], pacc0, err, w_acc, diag); pacc0 ^= 1; tc_fence_after();
        hidden_epilogue(lane_taddr, grp, row, PRM + kPrmW, PRM + kPrmW + kF, PRM + kPrmW + 2 * kF, X, STAT, 1.0f);
        tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
        // E4: e rows -> Y (after the s[src] half has been consumed)
        mbar_wait_timed(&bars[B_YFREE], pyf, err, w_acc, diag); pyf ^= 1;
        if (p.first_layer) build_rows(Y, wq, grp, lane, rows, p.edge_emb, ROWA, 2, 0);
        else build_rows(Y, wq, grp, lane, rows, p.e, ROWA, 0, row0);
        fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
        TIB_PHASE(2);   // E3 + E4
        // E5: w hidden 2 -> X (final: B operand of the output layer)
        mbar_wait_timed(&bars[B_ACC0], pacc0, err, w_acc, diag); pacc0 ^= 1; tc_fence_after();
        hidden_epilogue(lane_taddr, grp, row, PRM + kPrmW + 3 * kF, PRM + kPrmW + 4 * kF, PRM + kPrmW + 5 * kF, X, STAT + 1024, 1.0f);
        tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_XFULL]);
        // E6: phi hidden 1 -> Y
        mbar_wait_timed(&bars[B_ACC1], pacc1, err, w_acc, diag); pacc1 ^= 1; tc_fence_after();
        hidden_epilogue(lane_taddr + 128, grp, row, PRM + kPrmPhi, PRM + kPrmPhi + kF, PRM + kPrmPhi + 2 * kF, Y, STAT, kStateUnscale);
        tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
        if (tile + (int)gridDim.x < p.n_tiles) load_tables(tile + gridDim.x, par ^ 1);   // under the MMAs of phi layer 2
        TIB_PHASE(3);   // E5 + E6 (+ the next tile's tables)
        // E7: phi hidden 2 -> Y (final)
        mbar_wait_timed(&bars[B_ACC1], pacc1, err, w_acc, diag); pacc1 ^= 1; tc_fence_after();
        hidden_epilogue(lane_taddr + 128, grp, row, PRM + kPrmPhi + 3 * kF, PRM + kPrmPhi + 4 * kF, PRM + kPrmPhi + 5 * kF, Y, STAT + 1024, 1.0f);
        tc_fence_before(); fence_proxy_async(); mbar_arrive(&bars[B_YFULL]);
        }
        TIB_PHASE(4);   // E7
      }

      // ---- output layer, transposed: this thread owns feature f = TMEM lane; TMEM columns are edges.
      // m = phi3 * w3, split order gates | scale_edge_dir | ds | de | cross_gates     (cpainn.py:285-290)
      // Epilogue group g owns whole destination nodes (slots [g*per, g*per+per)), so the sums over
      // incoming edges live in registers across all splits: no shared-memory scatter, no atomics.
      const int f = row;
      const int nslots = node_hi - node_lo;
      const int per = (nslots + 3) >> 2;                    // <= kTileNodes / 4 = 4 destination nodes per group
      // regular tile: this group's four nodes own the 32 aligned columns [32*grp, +32), 8 each (always for n = 9)
      bool regular = nslots == kTileNodes;
#pragma unroll
      for (int k = 0; k <= 4; ++k) regular = regular && (SLOTROW[min(4 * grp + k, kTileNodes)] == 32 * grp + 8 * k);
      // accumulators of the (up to) 4 owned nodes: the node loop is rolled and always works on set 0,
      // rotating the sets after every node (4 rotations restore the order)
      float acc_s[4], acc_v[4][3];
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc_s[k] = 0.0f; acc_v[k][0] = acc_v[k][1] = acc_v[k][2] = 0.0f; }
      for (int it = 0; it < n_splits; ++it) {
        const int sp = p.first_layer ? it + 1 : it;
        const int pb = (it + 1) & 1;   // the LAST split uses slot 1: slot 0 (= the hidden accumulators) drains one split early
        const float bphi = PRM[kPrmB3 + sp * kF + f], bw = PRM[kPrmB3 + 5 * kF + sp * kF + f];
        mbar_wait_timed(&bars[B_TFULL0 + pb], ptf[pb], err, w_tfull, diag); ptf[pb] ^= 1; tc_fence_after();
        const uint32_t tphi = lane_taddr + 256 * pb, tw = tphi + 128;
        if (regular) {
          // ---- fast path: the group's 4 nodes x 8 edges in two TMEM round trips of 2 nodes each (phi and w accumulators
          // of 16 edges = 32 registers live at the product instead of 64: no spills at 96 registers)
          const int c0 = 32 * grp;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = c0 + 16 * h;
            float P[16], Q[16];
            tmem_ld16x2(tphi + ch, tw + ch, P, Q);
#pragma unroll
            for (int q = 0; q < 16; ++q) P[q] = __fmul_rn(P[q] + bphi, Q[q] + bw);
            if (sp == 0) {              // gates * v[src]
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const float* vi = p.v_old + (size_t)ROWA[ch + q].src * 3 * kF + f;
                acc_v[2 * h + (q >> 3)][0] = fmaf(P[q], __ldg(vi), acc_v[2 * h + (q >> 3)][0]);
                acc_v[2 * h + (q >> 3)][1] = fmaf(P[q], __ldg(vi + kF), acc_v[2 * h + (q >> 3)][1]);
                acc_v[2 * h + (q >> 3)][2] = fmaf(P[q], __ldg(vi + 2 * kF), acc_v[2 * h + (q >> 3)][2]);
              }
            } else if (sp == 1) {       // scale_edge_dir * dir
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const RowB rb = ROWB[ch + q];
                acc_v[2 * h + (q >> 3)][0] = fmaf(P[q], rb.dx, acc_v[2 * h + (q >> 3)][0]);
                acc_v[2 * h + (q >> 3)][1] = fmaf(P[q], rb.dy, acc_v[2 * h + (q >> 3)][1]);
                acc_v[2 * h + (q >> 3)][2] = fmaf(P[q], rb.dz, acc_v[2 * h + (q >> 3)][2]);
              }
            } else if (sp == 2) {       // ds
#pragma unroll
              for (int q = 0; q < 16; ++q) acc_s[2 * h + (q >> 3)] += P[q];
            } else if (sp == 3) {       // e += de                                       (cpainn.py:308)
              float* ep = p.e + (size_t)(row0 + ch) * kF + f;
              if (p.first_layer) {
#pragma unroll
                for (int q = 0; q < 16; ++q) Q[q] = __ldg(p.edge_emb + ((ROWA[ch + q].slot_last >> 16) & 0xFF) * kF + f);
              } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) Q[q] = ep[(size_t)q * kF];
              }
#pragma unroll
              for (int q = 0; q < 16; ++q) ep[(size_t)q * kF] = Q[q] + P[q];
            } else {                    // cross_gates * (dir x v[dst]), linear in dir     (cpainn.py:296-300)
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int k = 2 * h + k2;
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const RowB rb = ROWB[ch + 8 * k2 + q];
                  d0 = fmaf(P[8 * k2 + q], rb.dx, d0); d1 = fmaf(P[8 * k2 + q], rb.dy, d1); d2 = fmaf(P[8 * k2 + q], rb.dz, d2);
                }
                const float* vj = p.v_old + (size_t)(node_lo + 4 * grp + k) * 3 * kF + f;
                const float vj0 = __ldg(vj), vj1 = __ldg(vj + kF), vj2 = __ldg(vj + 2 * kF);
                acc_v[k][0] += __fmul_rn(d1, vj2) - __fmul_rn(d2, vj1);
                acc_v[k][1] += __fmul_rn(d2, vj0) - __fmul_rn(d0, vj2);
                acc_v[k][2] += __fmul_rn(d0, vj1) - __fmul_rn(d1, vj0);
              }
            }
          }
          tc_fence_before(); mbar_arrive(&bars[B_TEMPTY0 + pb]);
          continue;
        }
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int slot = grp * per + k;
          if (k < per && slot < nslots) {
            const int rbeg = SLOTROW[slot], rend = SLOTROW[slot + 1];
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;   // sp 4: sum of cross_gates * dir over the node's edges
#pragma unroll 1
            for (int r0 = rbeg & ~7; r0 < rend; r0 += 8) {      // 8-column aligned TMEM pieces
              float P[8], Q[8];
              tmem_ld8x2(tphi, tw, r0, P, Q);
              const int qlo = rbeg - r0, qhi = rend - r0;       // rows [qlo, qhi) of the piece belong to this node
              if (qlo <= 0 && qhi >= 8) {                       // whole piece belongs to the node (always, for n = 9)
#pragma unroll
                for (int q = 0; q < 8; ++q) P[q] = __fmul_rn(P[q] + bphi, Q[q] + bw);
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) P[q] = (q >= qlo && q < qhi) ? __fmul_rn(P[q] + bphi, Q[q] + bw) : 0.0f;
              }
              if (sp == 0) {            // gates * v[src]
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float* vi = p.v_old + (size_t)ROWA[r0 + q].src * 3 * kF + f;
                  acc_v[0][0] = fmaf(P[q], __ldg(vi), acc_v[0][0]);
                  acc_v[0][1] = fmaf(P[q], __ldg(vi + kF), acc_v[0][1]);
                  acc_v[0][2] = fmaf(P[q], __ldg(vi + 2 * kF), acc_v[0][2]);
                }
              } else if (sp == 1) {     // scale_edge_dir * dir
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const RowB rb = ROWB[r0 + q];
                  acc_v[0][0] = fmaf(P[q], rb.dx, acc_v[0][0]);
                  acc_v[0][1] = fmaf(P[q], rb.dy, acc_v[0][1]);
                  acc_v[0][2] = fmaf(P[q], rb.dz, acc_v[0][2]);
                }
              } else if (sp == 2) {     // ds
#pragma unroll
                for (int q = 0; q < 8; ++q) acc_s[0] += P[q];
              } else if (sp == 3) {     // e += de                                     (cpainn.py:308)
                float* ep = p.e + (size_t)(row0 + r0) * kF + f;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  Q[q] = !(q >= qlo && q < qhi) ? 0.0f
                         : p.first_layer ? __ldg(p.edge_emb + ((ROWA[r0 + q].slot_last >> 16) & 0xFF) * kF + f)
                                         : ep[(size_t)q * kF];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (q >= qlo && q < qhi) ep[(size_t)q * kF] = Q[q] + P[q];
              } else {                  // cross_gates * (dir x v[dst])                (cpainn.py:296-300)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const RowB rb = ROWB[r0 + q];
                  d0 = fmaf(P[q], rb.dx, d0); d1 = fmaf(P[q], rb.dy, d1); d2 = fmaf(P[q], rb.dz, d2);
                }
              }
            }
            if (sp == 4) {
              // sum_i g_i (dir_i x v[dst]) = (sum_i g_i dir_i) x v[dst]: the cross product is linear in dir
              const float* vj = p.v_old + (size_t)(node_lo + slot) * 3 * kF + f;
              const float vj0 = __ldg(vj), vj1 = __ldg(vj + kF), vj2 = __ldg(vj + 2 * kF);
              acc_v[0][0] += __fmul_rn(d1, vj2) - __fmul_rn(d2, vj1);
              acc_v[0][1] += __fmul_rn(d2, vj0) - __fmul_rn(d0, vj2);
              acc_v[0][2] += __fmul_rn(d0, vj1) - __fmul_rn(d1, vj0);
            }
          }
          {   // rotate the accumulator sets
            const float ts = acc_s[0], t0 = acc_v[0][0], t1 = acc_v[0][1], t2 = acc_v[0][2];
#pragma unroll
            for (int j = 0; j < 3; ++j) { acc_s[j] = acc_s[j + 1]; acc_v[j][0] = acc_v[j + 1][0]; acc_v[j][1] = acc_v[j + 1][1]; acc_v[j][2] = acc_v[j + 1][2]; }
            acc_s[3] = ts; acc_v[3][0] = t0; acc_v[3][1] = t1; acc_v[3][2] = t2;
          }
        }
        tc_fence_before(); mbar_arrive(&bars[B_TEMPTY0 + pb]);
      }
      TIB_PHASE(5);     // output layer (all splits, including waits)
      // ---- s_new = s_old + sum ds (cpainn.py:306), v_new = v_old + sum dv (cpainn.py:305)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int slot = grp * per + k;
        if (k < per && slot < nslots) {
          const size_t j = (size_t)(node_lo + slot);
          p.s_new[j * kF + f] = __ldg(p.s_old + j * kF + f) + acc_s[k];
#pragma unroll
          for (int xyz = 0; xyz < 3; ++xyz) {
            const size_t o = (j * 3 + xyz) * kF + f;
            p.v_new[o] = (p.first_layer ? 0.0f : __ldg(p.v_old + o)) + acc_v[k][xyz];
          }
        }
      }
      named_bar_sync(NB_ALL, kEpiThreads);                  // every reader of this tile's tables is done
      TIB_PHASE(6);     // write-back + end-of-tile barrier
    }
    if (diag && (tid == 0 || tid == 256)) {
      if (tid == 0) p.dbg[blockIdx.x * 8 + 7] = w_tfull + w_acc;
      for (int i = 0; i < 8; ++i) p.dbg[(size_t)(gridDim.x + blockIdx.x * 2 + (tid ? 1 : 0)) * 8 + i] = phc[i];
    }
#undef TIB_PHASE
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// ---- small helpers of the tensor-core drift ---------------------------------------------------------
// Per drift evaluation (geometry depends on x): node_in_ptr[j] = first (dst,src)-ordered edge row of node j
// (node_in_ptr[N] = E) and, per edge row, RowA {src, dst, last << 8 | first << 9 | edge type << 16, dist} and RowB {dir, 0}:
// r = x[src] - x[dst], d = |r|, dir = r / (1 + d)                                       (graph.py:27-29)
__global__ void k_edge_tables(const int* __restrict__ mol_ptr, const long long* __restrict__ edge_ptr, int n_mol,
                              const float* __restrict__ x, const unsigned char* __restrict__ edge_type,
                              int* __restrict__ node_in_ptr, uint4* __restrict__ rowa, uint4* __restrict__ rowb) {
  const int m = blockIdx.x;
  const int n0 = mol_ptr[m], n = mol_ptr[m + 1] - n0;
  const int e0 = (int)edge_ptr[m];
  for (int j = threadIdx.x; j < n; j += blockDim.x) node_in_ptr[n0 + j] = e0 + j * (n - 1);
  if (m == n_mol - 1 && threadIdx.x == 0) node_in_ptr[n0 + n] = (int)edge_ptr[n_mol];
  for (int row = threadIdx.x; row < n * (n - 1); row += blockDim.x) {
    const int jl = row / (n - 1), ip = row % (n - 1), il = ip + (ip >= jl);
    const int src = n0 + il, dst = n0 + jl;
    const float rx = x[3 * src + 0] - x[3 * dst + 0];
    const float ry = x[3 * src + 1] - x[3 * dst + 1];
    const float rz = x[3 * src + 2] - x[3 * dst + 2];
    const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    const float den = 1.0f + dist;
    // edge_type is given in the reference's (src,dst) order
    const uint32_t et = edge_type[(size_t)e0 + il * (n - 1) + jl - (jl > il)];
    rowa[e0 + row] = make_uint4((uint32_t)src, (uint32_t)dst, (uint32_t)(((ip == n - 2) << 8) | ((ip == 0) << 9)) | (et << 16),
                                __float_as_uint(dist));
    rowb[e0 + row] = make_uint4(__float_as_uint(__fdiv_rn(rx, den)), __float_as_uint(__fdiv_rn(ry, den)),
                                __float_as_uint(__fdiv_rn(rz, den)), 0u);
  }
}

}  // namespace tc
}  // namespace tib
