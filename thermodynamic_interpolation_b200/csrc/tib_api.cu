// tib_api.cu - extern "C" entry points of libtib.so (see include/tib.h).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tib.h"
#include "simt_drift.cuh"
#include "simt_tangent.cuh"
#include "steps.cuh"
#include "adw.cuh"
#include "postproc.cuh"
#include "tc_message.cuh"
#include "tc_selftest.cuh"
#include "tc_update.cuh"
#include "tc_readout.cuh"
#include "tc_chain.cuh"
#include "layered.cuh"

namespace {

thread_local std::string g_err;
thread_local uint64_t g_launches = 0;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}

// ---- optional per-kernel-class timing (tib_profile_begin/end): an event pair around each launch
struct ProfState {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[TIB_K_COUNT];
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
};
thread_local ProfState g_prof;

struct ProfScope {
  int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(int k, cudaStream_t s) : kind(k), st(s) {
    if (!g_prof.on) return;
    if (!g_prof.pool.empty()) { a = g_prof.pool.back().first; b = g_prof.pool.back().second; g_prof.pool.pop_back(); }
    else { cudaEventCreate(&a); cudaEventCreate(&b); }
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    g_prof.ev[kind].push_back({a, b});
  }
};

struct DeviceGuard {
  int prev = -1; bool ok = false;
  explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; ok = cudaSetDevice(dev) == cudaSuccess; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_CHECK()                                                                        \
  do {                                                                                        \
    ++g_launches;                                                                             \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---- programmatic dependent launch of the tensor-core kernels (common.cuh: pdl_trigger / pdl_wait) -------------------------
// When a launch leaves SMs idle (fewer tiles than SMs: the reference's own batch sizes of 12 and 64 conformers are 7 - 36
// tiles), the next kernel of the drift is scheduled on them right away, sets up its barriers, allocates its tensor memory,
// loads its parameters and - update / readout - starts streaming its weights, and parks at griddepcontrol.wait until the
// kernel before it has finished.  Full grids get no attribute: there is no idle SM to start on.
template <typename P>
void launch_tc(void (*kernel)(P), int grid, int block, size_t smem, cudaStream_t st, int n_sms, const P& prm) {
  static const bool off = getenv("TIB_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (!off && grid < n_sms) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, prm);          // errors surface in LAUNCH_CHECK (cudaGetLastError)
}

// Nodes per tile of the node-row tensor-core kernels (update, readout).  A tile always occupies the 128 TMEM lanes of one
// SM, but its memory phases (operand builds from v and s, the residual updates) scale with the rows it holds: when the
// batch cannot fill the GPU with 128-node tiles, smaller tiles on more SMs cut the latency of the launch - 12 conformers
// (108 nodes) are one 128-row tile on one SM, or seven 16-row tiles on seven.  Full grids keep 128.  The arithmetic of a
// node does not depend on its tile (row-local GEMM rows and LayerNorms), so results are bit-identical.
int tc_tile_nodes(int n_nodes, int n_sms) {
  static const bool off = getenv("TIB_FIXED_NODE_TILES") != nullptr;
  if (off) return 128;
  const int per_sm = (n_nodes + n_sms - 1) / n_sms;
  return std::min(128, std::max(16, (per_sm + 15) / 16 * 16));
}

// ---- packed-weight walking ---------------------------------------------------------------------
struct MlpShape { int k_in, h, n_out; };
size_t mlp_floats(const MlpShape& s) {
  return (size_t)s.h * s.k_in + 3 * (size_t)s.h + (size_t)s.h * s.h + 3 * (size_t)s.h + (size_t)s.n_out * s.h + s.n_out;
}
int n_temp_of(int variant) {
  return variant == TIB_VARIANT_AMBIENT ? 2 : (variant == TIB_VARIANT_LATENT_MULTI_T ? 1 : 0);
}

struct HostMlp { tib::MlpW w; };

}  // namespace

namespace tib_internal {      // shared with train_api.cu (the library's second translation unit)
int set_error(const char* msg) { g_err = msg; return -1; }
void count_launches(uint64_t n) { g_launches += n; }
// event pair around a launch of the other translation unit, when tib_profile_begin is active
void* prof_open(int kind, void* stream) { return g_prof.on ? new ProfScope(kind, (cudaStream_t)stream) : nullptr; }
void prof_close(void* h) { delete static_cast<ProfScope*>(h); }
}  // namespace tib_internal

struct tib_model {
  tib_model_desc d;
  int device = 0;
  int math = TIB_MATH_FP32_SIMT;   // tib_model_create selects F16X3_TC when n_features == 128
  int n_temp = 0;
  float* dev = nullptr;   // one allocation holding every repacked tensor
  size_t dev_floats = 0;
  const float* edge_emb = nullptr;
  const float* atom_emb = nullptr;
  tib::MlpW combine{};
  struct Layer {
    tib::MlpW phi, w, upd; const float *Ut, *Vt; const unsigned char* tc_msg = nullptr; const unsigned char* tc_upd = nullptr;
    // layered path (tc_chain.cuh): weight streams of the phi / w / update MLPs and of the [V; U] GEMM, max |LayerNorm gain|
    const unsigned char *ch_phi = nullptr, *ch_w = nullptr, *ch_uv = nullptr, *ch_upd = nullptr;
    float gmax_phi[2], gmax_w[2], gmax_upd[2];
  };
  std::vector<Layer> layers;
  tib::MlpW readout{};
  const unsigned char* tc_ro = nullptr;   // readout W1 | W2 as tensor-core chunks (F = 128)
  const float* Vout = nullptr;
  bool attrs_set = false;
  // tensor-core path (F = 128): per-layer streamed weight chunks in split-f16 operand images
  unsigned char* tc_blob = nullptr;
  unsigned char* ch_blob = nullptr;   // layered path (F = 128, 256)
  bool tc_attrs_set = false;
  bool ch_attrs_set = false;
  bool pe_cache = true;       // TIB_NO_PE_CACHE (read once, at model creation) turns the positional-encoding image cache off
  bool jvp_attrs_set = false;
  int* dev_err = nullptr;     // device error words: [0] bounded mbarrier waits, [1] non-finite tensor-core readout
  int n_sms = 148;
  long long* dev_dbg = nullptr;   // optional stall counters of the last tensor-core message launch
};

namespace {

// Copies one reference MLP into `dst` (host staging) in device layout and records device pointers.
//   in : W1[h,k] b1 g1 be1 W2[h,h] b2 g2 be2 W3[o,h] b3     (row-major [out,in])
//   out: W1t[k][h] b1 g1 be1 W2t[h][h] b2 g2 be2 W3t[h][o] (or W3 untransposed) b3
const float* repack_mlp(const float* src, const MlpShape& s, std::vector<float>& stage, float* dev_base,
                        tib::MlpW* out, bool transpose_w3) {
  auto push_T = [&](const float* W, int rows_out, int cols_in) {  // W[rows_out][cols_in] -> [cols_in][rows_out]
    size_t off = stage.size();
    stage.resize(off + (size_t)rows_out * cols_in);
    for (int o = 0; o < rows_out; ++o)
      for (int k = 0; k < cols_in; ++k) stage[off + (size_t)k * rows_out + o] = W[(size_t)o * cols_in + k];
    return dev_base + off;
  };
  auto push = [&](const float* v, size_t n) {
    size_t off = stage.size();
    // keep every tensor 16-byte aligned for 128-bit loads
    stage.insert(stage.end(), v, v + n);
    while (stage.size() % 4) stage.push_back(0.0f);
    return dev_base + off;
  };
  const int h = s.h;
  out->k_in = s.k_in;
  out->n_out = s.n_out;
  out->W1t = push_T(src, h, s.k_in); src += (size_t)h * s.k_in;
  out->b1 = push(src, h); src += h;
  out->g1 = push(src, h); src += h;
  out->be1 = push(src, h); src += h;
  out->W2t = push_T(src, h, h); src += (size_t)h * h;
  out->b2 = push(src, h); src += h;
  out->g2 = push(src, h); src += h;
  out->be2 = push(src, h); src += h;
  if (transpose_w3) out->W3t = push_T(src, s.n_out, h);
  else out->W3t = push(src, (size_t)s.n_out * h);
  src += (size_t)s.n_out * h;
  out->b3 = push(src, s.n_out); src += s.n_out;
  return src;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return fail("cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e));
  return 0;
}

// Per-F tile heights (rows per thread): message / node kernels.  See simt_drift.cuh for the budgets.
template <int F> struct Tiles;
template <> struct Tiles<32>  { static constexpr int MSG = 9, NODE = 4; };
template <> struct Tiles<64>  { static constexpr int MSG = 9, NODE = 4; };
template <> struct Tiles<128> { static constexpr int MSG = 9, NODE = 4; };
template <> struct Tiles<256> { static constexpr int MSG = 5, NODE = 2; };

// One streamed weight chunk: rows [row0, row0+128) x columns [col0, col0+32) of the row-major matrix
// W (leading dimension ld) as a split-f16 operand image, hi then lo (layout: csrc/tc_common.cuh).
void pack_tc_chunk(uint16_t* out, const float* W, int ld, int row0, int col0) {
  const size_t half = tib::tc::kChunkHalfBytes / 2;
  for (int k = 0; k < tib::tc::kChunkK; ++k)
    for (int r = 0; r < tib::tc::kRows; ++r) {
      const float w = W[(size_t)(row0 + r) * ld + col0 + k];
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      const size_t idx = (size_t)(k / 8) * (tib::tc::kLBO / 2) + (size_t)r * 8 + (k % 8);
      out[idx] = __half_as_ushort(hi);
      out[half + idx] = __half_as_ushort(lo);
    }
}

struct Workspace {
  float *s[2], *v[2], *e, *drift, *score;
  float *k;          // [7][3N] dopri stages / rk4 stages
  float *ytmp, *ycur, *ynew;
  double *partial, *scalar;
  int *node_in_ptr;
  uint4 *rowa, *rowb;
  float *phitab;     // [kPhiTabRows][F] first-layer phi table (tc_message.cuh)
  unsigned char *peimg; size_t peimg_bytes;
  char* end;         // first byte after the carved region (the layered path carves its arrays from here)   // positional-encoding operand images, 64 KB per message tile (tc_message.cuh)
  static size_t pe_bytes(int n_nodes, long long n_edges) {   // tiles close at 16 nodes or at > 64 rows
    return (size_t)65536 * ((size_t)n_nodes / 16 + (size_t)n_edges / 64 + 2);
  }
  static constexpr int kPhiTabRows = 4096;
  static constexpr int kPartials = 1024;
  static size_t align(size_t x) { return (x + 255) & ~(size_t)255; }
  static size_t bytes(int F, int n_nodes, long long n_edges) {
    size_t b = 0;
    b += 2 * align(sizeof(float) * (size_t)n_nodes * F);
    b += 2 * align(sizeof(float) * (size_t)n_nodes * 3 * F);
    b += align(sizeof(float) * (size_t)n_edges * F);
    b += 12 * align(sizeof(float) * (size_t)n_nodes * 3);   // drift, score, k[7], ytmp, ycur, ynew
    b += align(sizeof(double) * kPartials * 5) + align(sizeof(double) * 8);
    b += align(sizeof(int) * ((size_t)n_nodes + 1)) + 2 * align(sizeof(uint4) * (size_t)n_edges);
    b += align(sizeof(float) * (size_t)kPhiTabRows * F);
    if (F == 128) b += align(pe_bytes(n_nodes, n_edges));
    return b;
  }
  void carve(void* base, int F, int n_nodes, long long n_edges) {
    char* p = (char*)base;
    auto take = [&](size_t nbytes) { char* r = p; p += align(nbytes); return r; };
    s[0] = (float*)take(sizeof(float) * (size_t)n_nodes * F);
    s[1] = (float*)take(sizeof(float) * (size_t)n_nodes * F);
    v[0] = (float*)take(sizeof(float) * (size_t)n_nodes * 3 * F);
    v[1] = (float*)take(sizeof(float) * (size_t)n_nodes * 3 * F);
    e = (float*)take(sizeof(float) * (size_t)n_edges * F);
    const size_t st = sizeof(float) * (size_t)n_nodes * 3;
    drift = (float*)take(st);
    score = (float*)take(st);
    k = (float*)take(st);
    for (int i = 1; i < 7; ++i) take(st);
    ytmp = (float*)take(st);
    ycur = (float*)take(st);
    ynew = (float*)take(st);
    partial = (double*)take(sizeof(double) * kPartials * 5);
    scalar = (double*)take(sizeof(double) * 8);
    node_in_ptr = (int*)take(sizeof(int) * ((size_t)n_nodes + 1));
    rowa = (uint4*)take(sizeof(uint4) * (size_t)n_edges);
    rowb = (uint4*)take(sizeof(uint4) * (size_t)n_edges);
    phitab = (float*)take(sizeof(float) * (size_t)kPhiTabRows * F);
    peimg_bytes = F == 128 ? pe_bytes(n_nodes, n_edges) : 0;
    peimg = F == 128 ? (unsigned char*)take(peimg_bytes) : nullptr;
    end = p;
  }
  static size_t kstride(int n_nodes) { return align(sizeof(float) * (size_t)n_nodes * 3) / sizeof(float); }
};

// s0 = InvariantFeatures(atoms, T0, T1, t) into ws.s[0] (de-duplicated rows + gather when the batch has the tables)
template <int F>
int launch_embed(tib_model* m, const tib_batch* b, const tib::DriftBatch& db, float t, Workspace& ws, cudaStream_t st) {
  using namespace tib;
  constexpr int RN = Tiles<F>::NODE;
  const int node_tiles = (b->n_nodes + 8 * RN - 1) / (8 * RN);
  if (b->embed_index && b->n_embed_rows > 0 && b->n_embed_rows <= b->n_nodes) {
    // de-duplicated embedding: U distinct rows into ws.s[1] (free until the first message layer), then a gather
    if (!b->embed_atom_id || (m->n_temp >= 1 && !b->embed_temp0) || (m->n_temp >= 2 && !b->embed_temp1))
      return fail("embed_index given without the de-duplicated atom/temperature tables");
    DriftBatch du = db;
    du.n_nodes = b->n_embed_rows; du.atom_id = b->embed_atom_id; du.temp0 = b->embed_temp0; du.temp1 = b->embed_temp1;
    EmbedP ep{du, m->combine, m->atom_emb, m->n_temp, t, m->d.temp_mean, m->d.temp_range, m->d.temp_length,
              m->d.time_length, ws.s[1]};
    ProfScope ps(TIB_K_EMBED, st);
    if (b->n_embed_rows <= 8 * 2 * 148)   // few rows: 8 per CTA and 16 weight rows in flight per lane (latency bound)
      k_embed<F, 1, 16><<<(b->n_embed_rows + 7) / 8, TIB_THREADS, smem_embed<F, 1>(), st>>>(ep);
    else
      k_embed<F, RN><<<(b->n_embed_rows + 8 * RN - 1) / (8 * RN), TIB_THREADS, smem_embed<F, RN>(), st>>>(ep);
    LAUNCH_CHECK();
    const long long total = (long long)b->n_nodes * (F / 4);
    k_gather_rows<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(ws.s[1], b->embed_index, ws.s[0],
                                                                                      b->n_nodes, F);
    LAUNCH_CHECK();
  } else {
    EmbedP ep{db, m->combine, m->atom_emb, m->n_temp, t, m->d.temp_mean, m->d.temp_range, m->d.temp_length,
              m->d.time_length, ws.s[0]};
    ProfScope ps(TIB_K_EMBED, st);
    k_embed<F, RN><<<node_tiles, TIB_THREADS, smem_embed<F, RN>(), st>>>(ep);
    LAUNCH_CHECK();
  }
  return 0;
}

template <int F>
int drift_simt(tib_model* m, const tib_batch* b, const float* x, float t, float* out, Workspace& ws, cudaStream_t st) {
  using namespace tib;
  constexpr int RM = Tiles<F>::MSG, RN = Tiles<F>::NODE;
  if (!m->attrs_set) {
    if (set_smem(k_embed<F, RN>, smem_embed<F, RN>())) return -1;
    if (set_smem(k_embed<F, 1, 16>, smem_embed<F, 1>())) return -1;
    if (set_smem(k_message<F, RM>, smem_message<F, RM>())) return -1;
    if (set_smem(k_update<F, RN>, smem_update<F, RN>())) return -1;
    if (set_smem(k_readout<F, RN>, smem_readout<F, RN>())) return -1;
    m->attrs_set = true;
  }
  DriftBatch db{b->n_mol, b->n_nodes, (long long)b->n_edges, b->mol_ptr, (const long long*)b->edge_ptr,
                b->atom_id, b->edge_type, b->temp0, b->temp1};
  const int node_tiles = (b->n_nodes + 8 * RN - 1) / (8 * RN);

  const bool use_tc = (F == 128) && m->math != TIB_MATH_FP32_SIMT;
  int nodes_per_tile = 0, n_tiles = 0;
  bool use_phi_tab = false, fused_tab = false;
  if constexpr (F == 128) {
    // few de-duplicated rows (single-species batches): one CTA per (embedding row, edge type) computes s0 and the first
    // layer's phi table with K split over its warps; the summation order differs from the fp32 path, so tensor-core modes only
    const bool dedupe = b->embed_index && b->n_embed_rows > 0 && b->n_embed_rows <= b->n_nodes;
    if (use_tc && dedupe && (long long)b->n_embed_rows * m->d.n_edge_types <= 2 * m->n_sms) {
      if (!b->embed_atom_id || (m->n_temp >= 1 && !b->embed_temp0) || (m->n_temp >= 2 && !b->embed_temp1))
        return fail("embed_index given without the de-duplicated atom/temperature tables");
      DriftBatch du = db;
      du.n_nodes = b->n_embed_rows; du.atom_id = b->embed_atom_id; du.temp0 = b->embed_temp0; du.temp1 = b->embed_temp1;
      EmbedTabP ep{{du, m->combine, m->atom_emb, m->n_temp, t, m->d.temp_mean, m->d.temp_range, m->d.temp_length,
                    m->d.time_length, ws.s[1]},
                   m->layers[0].phi, m->edge_emb, m->d.n_edge_types, ws.phitab};
      ProfScope ps(TIB_K_EMBED, st);
      k_embed_tab<128><<<b->n_embed_rows * m->d.n_edge_types, TIB_THREADS, 0, st>>>(ep);
      LAUNCH_CHECK();
      const long long total = (long long)b->n_nodes * (F / 4);
      k_gather_rows<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(ws.s[1], b->embed_index, ws.s[0],
                                                                                        b->n_nodes, F);
      LAUNCH_CHECK();
      fused_tab = use_phi_tab = true;
    }
  }
  if (!fused_tab && launch_embed<F>(m, b, db, t, ws, st)) return -1;
  if (use_tc) {
    if (b->n_edges >= (1ll << 31)) return fail("tensor-core path: n_edges=%lld exceeds int32 row indices", (long long)b->n_edges);
    if (!m->tc_attrs_set) {
      if (set_smem(tc::k_message_tc<false>, tc::MsgSmem::TOTAL)) return -1;
      if (set_smem(tc::k_message_tc<true>, tc::MsgSmem::TOTAL)) return -1;
      if (set_smem(tc::k_update_tc<false>, tc::UpdSmem::TOTAL)) return -1;
      if (set_smem(tc::k_update_tc<true>, tc::UpdSmem::TOTAL)) return -1;
      if (set_smem(tc::k_readout_tc, tc::RoSmem::TOTAL)) return -1;
      if (set_smem(k_phi_table<F, 16>, sizeof(float) * 8 * 4 * F)) return -1;
      m->tc_attrs_set = true;
    }
    nodes_per_tile = std::min(tc::kTileNodes, 128 / (b->max_atoms - 1));
    n_tiles = (b->tile_node_ptr && b->n_tiles > 0) ? b->n_tiles : (b->n_nodes + nodes_per_tile - 1) / nodes_per_tile;
    // first layer: phi's hidden layers once per (de-duplicated embedding row, edge type) instead of once per edge
    if (!fused_tab)
      use_phi_tab = b->embed_index && b->n_embed_rows > 0 && b->n_embed_rows <= b->n_nodes &&
                    (long long)b->n_embed_rows * m->d.n_edge_types <= Workspace::kPhiTabRows;
    if (use_phi_tab && !fused_tab) {
      PhiTabP pp{m->layers[0].phi, ws.s[1], m->edge_emb, b->n_embed_rows, m->d.n_edge_types, ws.phitab};
      const int total = b->n_embed_rows * m->d.n_edge_types;
      ProfScope ps(TIB_K_EMBED, st);
      k_phi_table<F, 16><<<(total + 7) / 8, TIB_THREADS, sizeof(float) * 8 * 4 * F, st>>>(pp);
      LAUNCH_CHECK();
    }
    // geometry + edge types per (dst,src)-ordered row; e0 = Emb(edge_type) is formed inside the first message layer
    ProfScope ps(TIB_K_EDGE_INIT, st);
    tc::k_edge_tables<<<b->n_mol, 128, 0, st>>>(b->mol_ptr, (const long long*)b->edge_ptr, b->n_mol, x, b->edge_type,
                                                ws.node_in_ptr, ws.rowa, ws.rowb);
    LAUNCH_CHECK();
  } else {
    const long long total = (long long)b->n_edges * (F / 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    ProfScope ps(TIB_K_EDGE_INIT, st);
    k_edge_init<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(b->edge_type, m->edge_emb, ws.e, (long long)b->n_edges, F);
    LAUNCH_CHECK();
  }
  int cur = 0;
  for (int l = 0; l < m->d.n_layers; ++l) {
    const tib_model::Layer& L = m->layers[l];
    if (use_tc) {
      tc::TcMsgP tp{};
      tp.n_nodes = b->n_nodes; tp.n_tiles = n_tiles; tp.nodes_per_tile = nodes_per_tile;
      tp.node_in_ptr = ws.node_in_ptr; tp.rowa = ws.rowa; tp.rowb = ws.rowb;
      tp.tile_node_ptr = (b->tile_node_ptr && b->n_tiles > 0) ? b->tile_node_ptr : nullptr;
      tp.s_old = ws.s[cur]; tp.v_old = ws.v[cur]; tp.s_new = ws.s[cur ^ 1]; tp.v_new = ws.v[cur ^ 1]; tp.e = ws.e;
      tp.wblob = L.tc_msg; tp.edge_emb = m->edge_emb;
      if (l == 0 && use_phi_tab) { tp.phi_tab = ws.phitab; tp.embed_index = b->embed_index; tp.n_et = m->d.n_edge_types; }
      if (ws.peimg && (size_t)n_tiles * tc::kOperandBytes <= ws.peimg_bytes && m->d.n_layers > 1 && m->pe_cache) {
        tp.pe_img = ws.peimg; tp.pe_mode = l == 0 ? 1 : 2;
      }
      tp.prm = tc::MsgParams{{L.w.b1, L.w.g1, L.w.be1, L.w.b2, L.w.g2, L.w.be2,
                              L.phi.b1, L.phi.g1, L.phi.be1, L.phi.b2, L.phi.g2, L.phi.be2}, L.phi.b3, L.w.b3};
      tp.length_scale = m->d.length_scale; tp.first_layer = (l == 0); tp.passes = (m->math == TIB_MATH_F16_TC) ? 1 : 3;
      tp.err = m->dev_err; tp.dbg = m->dev_dbg;
      ProfScope ps(TIB_K_MESSAGE, st);
      if (tp.dbg) launch_tc(tc::k_message_tc<true>, std::min(n_tiles, m->n_sms), tc::kThreads, tc::MsgSmem::TOTAL, st, m->n_sms, tp);
      else launch_tc(tc::k_message_tc<false>, std::min(n_tiles, m->n_sms), tc::kThreads, tc::MsgSmem::TOTAL, st, m->n_sms, tp);
      LAUNCH_CHECK();
    } else {
      MessageP mp{db, L.phi, L.w, x, ws.s[cur], ws.v[cur], ws.s[cur ^ 1], ws.v[cur ^ 1], ws.e, m->d.length_scale, l == 0};
      ProfScope ps(TIB_K_MESSAGE, st);
      k_message<F, RM><<<b->n_mol, TIB_THREADS, smem_message<F, RM>(), st>>>(mp);
      LAUNCH_CHECK();
    }
    cur ^= 1;
    if (use_tc) {
      tc::TcUpdP up{};
      up.n_nodes = b->n_nodes; up.tile_nodes = tc_tile_nodes(b->n_nodes, m->n_sms);
      up.n_tiles = (b->n_nodes + up.tile_nodes - 1) / up.tile_nodes;
      up.s = ws.s[cur]; up.v = ws.v[cur]; up.wblob = L.tc_upd;
      up.b1 = L.upd.b1; up.g1 = L.upd.g1; up.be1 = L.upd.be1; up.b2 = L.upd.b2; up.g2 = L.upd.g2; up.be2 = L.upd.be2; up.b3 = L.upd.b3;
      up.passes = (m->math == TIB_MATH_F16_TC) ? 1 : 3; up.err = m->dev_err;
      up.dbg = m->dev_dbg ? m->dev_dbg + 8 * 1024 : nullptr;   // second half of the diagnostics buffer
      ProfScope ps(TIB_K_UPDATE, st);
      if (up.dbg) launch_tc(tc::k_update_tc<true>, std::min(up.n_tiles, m->n_sms), tc::kThreads, tc::UpdSmem::TOTAL, st, m->n_sms, up);
      else launch_tc(tc::k_update_tc<false>, std::min(up.n_tiles, m->n_sms), tc::kThreads, tc::UpdSmem::TOTAL, st, m->n_sms, up);
      LAUNCH_CHECK();
    } else {
      UpdateP up{b->n_nodes, L.upd, L.Ut, L.Vt, ws.s[cur], ws.v[cur]};
      ProfScope ps(TIB_K_UPDATE, st);
      k_update<F, RN><<<node_tiles, TIB_THREADS, smem_update<F, RN>(), st>>>(up);
      LAUNCH_CHECK();
    }
  }
  if (use_tc) {
    tc::TcRoP rp{};
    rp.n_nodes = b->n_nodes; rp.tile_nodes = tc_tile_nodes(b->n_nodes, m->n_sms);
    rp.n_tiles = (b->n_nodes + rp.tile_nodes - 1) / rp.tile_nodes;
    rp.s = ws.s[cur]; rp.v = ws.v[cur]; rp.out = out; rp.wblob = m->tc_ro;
    rp.b1 = m->readout.b1; rp.g1 = m->readout.g1; rp.be1 = m->readout.be1;
    rp.b2 = m->readout.b2; rp.g2 = m->readout.g2; rp.be2 = m->readout.be2;
    rp.w3 = m->readout.W3t + F; rp.b3 = m->readout.b3 + 1; rp.vout = m->Vout;
    rp.passes = (m->math == TIB_MATH_F16_TC) ? 1 : 3; rp.err = m->dev_err;
    ProfScope ps(TIB_K_READOUT, st);
    launch_tc(tc::k_readout_tc, std::min(rp.n_tiles, m->n_sms), tc::kThreads, tc::RoSmem::TOTAL, st, m->n_sms, rp);
  } else {
    ReadoutP rp{b->n_nodes, m->readout, m->Vout, ws.s[cur], ws.v[cur], out};
    ProfScope ps(TIB_K_READOUT, st);
    k_readout<F, RN><<<node_tiles, TIB_THREADS, smem_readout<F, RN>(), st>>>(rp);
  }
  LAUNCH_CHECK();
  return 0;
}


// ---- exact divergence by forward-mode tangents (csrc/simt_tangent.cuh) ----------------------------
// D tangent directions per pass and tile heights (rows per thread) of the dual-number kernels; the
// budgets are the 227 KB of shared memory: (1 + D) copies of every activation buffer.
template <int F> struct JvpTiles { static constexpr int D = 3, MSG = 2, NODE = 1; };
template <> struct JvpTiles<256> { static constexpr int D = 1, MSG = 2, NODE = 1; };

struct TangentWs {
  float *ts[2], *tv[2], *te, *tout;
  size_t st_s, st_v, st_e, st_o;   // floats between directions
  static size_t align(size_t x) { return Workspace::align(x); }
  static size_t bytes(int F, int D, int n_nodes, long long n_edges) {
    return (size_t)D * (2 * align(sizeof(float) * (size_t)n_nodes * F) + 2 * align(sizeof(float) * (size_t)n_nodes * 3 * F) +
                        align(sizeof(float) * (size_t)n_edges * F) + align(sizeof(float) * (size_t)n_nodes * 3));
  }
  void carve(void* base, int F, int D, int n_nodes, long long n_edges) {
    char* p = (char*)base;
    st_s = align(sizeof(float) * (size_t)n_nodes * F) / sizeof(float);
    st_v = align(sizeof(float) * (size_t)n_nodes * 3 * F) / sizeof(float);
    st_e = align(sizeof(float) * (size_t)n_edges * F) / sizeof(float);
    st_o = align(sizeof(float) * (size_t)n_nodes * 3) / sizeof(float);
    auto take = [&](size_t floats) { float* r = (float*)p; p += sizeof(float) * floats * D; return r; };
    ts[0] = take(st_s); ts[1] = take(st_s);
    tv[0] = take(st_v); tv[1] = take(st_v);
    te = take(st_e);
    tout = take(st_o);
  }
};

int jvp_dirs(int F) { return F == 256 ? JvpTiles<256>::D : JvpTiles<128>::D; }

template <int F>
int drift_div_simt(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b, float* out_div,
                   Workspace& ws, TangentWs& tw, cudaStream_t st) {
  using namespace tib;
  constexpr int D = JvpTiles<F>::D, RM = JvpTiles<F>::MSG, RN = JvpTiles<F>::NODE;
  if (!m->jvp_attrs_set) {
    if (set_smem(k_embed<F, Tiles<F>::NODE>, smem_embed<F, Tiles<F>::NODE>())) return -1;
    if (set_smem(k_embed<F, 1, 16>, smem_embed<F, 1>())) return -1;
    if (set_smem(k_message_jvp<F, RM, D>, smem_message_jvp<F, RM, D>())) return -1;
    if (set_smem(k_update_jvp<F, RN, D>, smem_update_jvp<F, RN, D>())) return -1;
    if (set_smem(k_readout_jvp<F, RN, D>, smem_readout_jvp<F, RN, D>())) return -1;
    m->jvp_attrs_set = true;
  }
  DriftBatch db{b->n_mol, b->n_nodes, (long long)b->n_edges, b->mol_ptr, (const long long*)b->edge_ptr,
                b->atom_id, b->edge_type, b->temp0, b->temp1};
  const int node_tiles = (b->n_nodes + 8 * RN - 1) / (8 * RN);
  const int n_dirs = 3 * b->max_atoms;
  for (int dir0 = 0; dir0 < n_dirs; dir0 += D) {
    // the primal state is rebuilt in every pass: e is updated in place by the message layers
    if (launch_embed<F>(m, b, db, t, ws, st)) return -1;
    {
      const long long total = (long long)b->n_edges * (F / 4);
      const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
      ProfScope ps(TIB_K_EDGE_INIT, st);
      k_edge_init<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(b->edge_type, m->edge_emb, ws.e, (long long)b->n_edges, F);
      LAUNCH_CHECK();
    }
    int cur = 0;
    for (int l = 0; l < m->d.n_layers; ++l) {
      const tib_model::Layer& L = m->layers[l];
      MessageJvpP mp{{db, L.phi, L.w, x, ws.s[cur], ws.v[cur], ws.s[cur ^ 1], ws.v[cur ^ 1], ws.e, m->d.length_scale, l == 0},
                     tw.ts[cur], tw.tv[cur], tw.ts[cur ^ 1], tw.tv[cur ^ 1], tw.te, tw.st_s, tw.st_v, tw.st_e, dir0};
      {
        ProfScope ps(TIB_K_MESSAGE, st);
        k_message_jvp<F, RM, D><<<b->n_mol, TIB_THREADS, smem_message_jvp<F, RM, D>(), st>>>(mp);
        LAUNCH_CHECK();
      }
      cur ^= 1;
      UpdateJvpP up{{b->n_nodes, L.upd, L.Ut, L.Vt, ws.s[cur], ws.v[cur]}, tw.ts[cur], tw.tv[cur], tw.st_s, tw.st_v};
      ProfScope ps(TIB_K_UPDATE, st);
      k_update_jvp<F, RN, D><<<node_tiles, TIB_THREADS, smem_update_jvp<F, RN, D>(), st>>>(up);
      LAUNCH_CHECK();
    }
    ReadoutJvpP rp{{b->n_nodes, m->readout, m->Vout, ws.s[cur], ws.v[cur], dir0 == 0 ? out_b : nullptr},
                   tw.ts[cur], tw.tv[cur], tw.tout, tw.st_s, tw.st_v, tw.st_o};
    {
      ProfScope ps(TIB_K_READOUT, st);
      k_readout_jvp<F, RN, D><<<node_tiles, TIB_THREADS, smem_readout_jvp<F, RN, D>(), st>>>(rp);
      LAUNCH_CHECK();
    }
    k_div_pick<<<(b->n_mol + 255) / 256, 256, 0, st>>>(b->mol_ptr, b->n_mol, tw.tout, tw.st_o, dir0, D, out_div);
    LAUNCH_CHECK();
  }
  return 0;
}

#include "layered_api.inl"

int check_batch(const tib_model* m, const tib_batch* b) {
  if (!m || !b) return fail("null model or batch");
  if (b->n_mol <= 0 || b->n_nodes <= 0) return fail("empty batch (n_mol=%d, n_nodes=%d)", b->n_mol, b->n_nodes);
  if (b->max_atoms > TIB_MAX_ATOMS) return fail("max_atoms=%d exceeds the supported %d", b->max_atoms, TIB_MAX_ATOMS);
  if (b->max_atoms < 2) return fail("molecules need at least 2 atoms (max_atoms=%d)", b->max_atoms);
  if (!b->mol_ptr || !b->edge_ptr || !b->atom_id || !b->edge_type) return fail("batch pointers must not be null");
  if (m->n_temp >= 1 && !b->temp0) return fail("temp0 is required for this model variant");
  if (m->n_temp >= 2 && !b->temp1) return fail("temp1 is required for the ambient variant");
  return 0;
}

bool uses_layered(const tib_model* m) {
  return (m->d.n_features == 256 && m->math != TIB_MATH_FP32_SIMT) || (m->d.n_features == 128 && m->math == TIB_MATH_F16X3_LAYERED);
}

int drift_dispatch(tib_model* m, const tib_batch* b, const float* x, float t, float* out, Workspace& ws, cudaStream_t st) {
  const int F = m->d.n_features;
  if (m->math != TIB_MATH_FP32_SIMT && F != 128 && F != 256)
    return fail("the tensor-core math modes are built for n_features = 128 and 256 (got %d); use TIB_MATH_FP32_SIMT", F);
  if (uses_layered(m)) {
    LayWs lw{};
    lw.layout((char*)ws.end, F, b->n_nodes, (long long)b->n_edges, b->max_atoms, false);
    return F == 256 ? drift_layered<256>(m, b, x, t, out, ws, lw, st) : drift_layered<128>(m, b, x, t, out, ws, lw, st);
  }
  switch (F) {
    case 32: return drift_simt<32>(m, b, x, t, out, ws, st);
    case 64: return drift_simt<64>(m, b, x, t, out, ws, st);
    case 128: return drift_simt<128>(m, b, x, t, out, ws, st);
    case 256: return drift_simt<256>(m, b, x, t, out, ws, st);
  }
  return fail("unsupported n_features=%d", F);
}

int grid_for(size_t n, int threads = 256) {
  size_t blocks = (n + threads - 1) / threads;
  const size_t cap = 148 * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

size_t drift_ws_bytes(const tib_model* m, int n_nodes, long long n_edges, int max_atoms) {
  size_t need = Workspace::bytes(m->d.n_features, n_nodes, n_edges);
  if (uses_layered(m)) need += LayWs::bytes(m->d.n_features, n_nodes, n_edges, max_atoms, false);
  return need;
}

int prep_ws(const tib_model* m, const tib_batch* b, void* workspace, size_t workspace_bytes, Workspace& ws) {
  const size_t need = drift_ws_bytes(m, b->n_nodes, (long long)b->n_edges, b->max_atoms);
  if (!workspace || workspace_bytes < need) return fail("workspace too small: %zu < %zu bytes", workspace_bytes, need);
  if (((uintptr_t)workspace & 255) != 0) return fail("workspace must be 256-byte aligned");
  ws.carve(workspace, m->d.n_features, b->n_nodes, (long long)b->n_edges);
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* tib_last_error(void) { return g_err.c_str(); }
int tib_abi_version(void) { return TIB_ABI_VERSION; }
uint64_t tib_launch_count(int reset) {
  uint64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int tib_profile_begin(void) {
  for (auto& v : g_prof.ev) { for (auto& p : v) g_prof.pool.push_back(p); v.clear(); }
  g_prof.on = true;
  return 0;
}

int tib_profile_end(double* ms_sum, uint64_t* launches) {
  g_prof.on = false;
  for (int k = 0; k < TIB_K_COUNT; ++k) {
    double tot = 0.0;
    for (auto& p : g_prof.ev[k]) {
      CUDA_TRY(cudaEventSynchronize(p.second));
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, p.first, p.second));
      tot += ms;
      g_prof.pool.push_back(p);
    }
    if (ms_sum) ms_sum[k] = tot;
    if (launches) launches[k] = g_prof.ev[k].size();
    g_prof.ev[k].clear();
  }
  return 0;
}

size_t tib_packed_weight_count(const tib_model_desc* d) {
  if (!d) return 0;
  const int F = d->n_features, nt = n_temp_of(d->variant);
  size_t n = (size_t)d->n_edge_types * F + (size_t)d->n_types * F;
  n += mlp_floats({(2 + nt) * F, F, F});
  for (int l = 0; l < d->n_layers; ++l) {
    n += mlp_floats({2 * F, F, 5 * F}) + mlp_floats({F, F, 5 * F}) + 2 * (size_t)F * F + mlp_floats({2 * F, F, 3 * F});
  }
  n += mlp_floats({F, F, 2}) + F;
  return n;
}

int tib_model_create(tib_model** out, const tib_model_desc* d, const float* w, size_t n_floats, int device) {
  if (!out || !d || !w) return fail("tib_model_create: null argument");
  if (d->abi_version != TIB_ABI_VERSION) return fail("ABI mismatch: header %d, library %d", d->abi_version, TIB_ABI_VERSION);
  const int F = d->n_features;
  if (!(F == 32 || F == 64 || F == 128 || F == 256)) return fail("n_features must be 32, 64, 128 or 256 (got %d)", F);
  if (d->variant < 0 || d->variant > 2) return fail("bad variant %d", d->variant);
  if (d->n_layers < 1 || d->n_types < 1 || d->n_edge_types < 1) return fail("bad layer/type counts");
  if (n_floats != tib_packed_weight_count(d))
    return fail("packed weight count mismatch: got %zu, descriptor needs %zu", n_floats, tib_packed_weight_count(d));
  DeviceGuard guard(device);          // the caller's current device is restored on every return path
  if (!guard.ok) return fail("cudaSetDevice(%d) failed", device);
  tib_model* m = new tib_model();
  m->d = *d;
  m->device = device;
  m->n_temp = n_temp_of(d->variant);
  m->pe_cache = getenv("TIB_NO_PE_CACHE") == nullptr;
  m->math = (F == 128 || F == 256) ? TIB_MATH_F16X3_TC : TIB_MATH_FP32_SIMT;
  const int nt = m->n_temp;

  // upper bound of the staged size: every tensor padded to 4 floats
  std::vector<float> stage;
  stage.reserve(n_floats + 64 * (size_t)(d->n_layers + 2) * 10);
  // device allocation first (pointers are computed relative to it); size known after staging, so
  // stage against a null base and rebase afterwards.
  float* base = nullptr;
  auto push = [&](const float* v, size_t n) {
    size_t off = stage.size();
    stage.insert(stage.end(), v, v + n);
    while (stage.size() % 4) stage.push_back(0.0f);
    return base + off;
  };
  auto push_T = [&](const float* W, int rows_out, int cols_in) {
    size_t off = stage.size();
    stage.resize(off + (size_t)rows_out * cols_in);
    for (int o = 0; o < rows_out; ++o)
      for (int k = 0; k < cols_in; ++k) stage[off + (size_t)k * rows_out + o] = W[(size_t)o * cols_in + k];
    while (stage.size() % 4) stage.push_back(0.0f);
    return base + off;
  };
  const float* src = w;
  m->edge_emb = push(src, (size_t)d->n_edge_types * F); src += (size_t)d->n_edge_types * F;
  m->atom_emb = push(src, (size_t)d->n_types * F); src += (size_t)d->n_types * F;
  src = repack_mlp(src, {(2 + nt) * F, F, F}, stage, base, &m->combine, true);
  m->layers.resize(d->n_layers);
  std::vector<const float*> phi_src(d->n_layers), w_src(d->n_layers), uv_src(d->n_layers);
  for (int l = 0; l < d->n_layers; ++l) {
    auto& L = m->layers[l];
    phi_src[l] = src;
    src = repack_mlp(src, {2 * F, F, 5 * F}, stage, base, &L.phi, true);
    w_src[l] = src;
    src = repack_mlp(src, {F, F, 5 * F}, stage, base, &L.w, true);
    uv_src[l] = src;
    L.Ut = push_T(src, F, F); src += (size_t)F * F;
    L.Vt = push_T(src, F, F); src += (size_t)F * F;
    src = repack_mlp(src, {2 * F, F, 3 * F}, stage, base, &L.upd, true);
  }
  const float* ro_src = src;
  src = repack_mlp(src, {F, F, 2}, stage, base, &m->readout, false);
  m->Vout = push(src, F); src += F;
  if ((size_t)(src - w) != n_floats) { delete m; return fail("internal: weight walk consumed %zu of %zu", (size_t)(src - w), n_floats); }

  m->dev_floats = stage.size();
  cudaError_t e = cudaMalloc(&m->dev, sizeof(float) * m->dev_floats);
  if (e != cudaSuccess) { delete m; return fail("cudaMalloc(%zu floats): %s", stage.size(), cudaGetErrorString(e)); }
  e = cudaMemcpy(m->dev, stage.data(), sizeof(float) * m->dev_floats, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(m->dev); delete m; return fail("cudaMemcpy weights: %s", cudaGetErrorString(e)); }
  // rebase the recorded (null-relative) pointers onto the device allocation
  auto rb = [&](const float*& p) { p = m->dev + (p - (const float*)nullptr); };
  auto rb_mlp = [&](tib::MlpW& q) { rb(q.W1t); rb(q.b1); rb(q.g1); rb(q.be1); rb(q.W2t); rb(q.b2); rb(q.g2); rb(q.be2); rb(q.W3t); rb(q.b3); };
  rb(m->edge_emb); rb(m->atom_emb); rb_mlp(m->combine);
  for (auto& L : m->layers) { rb_mlp(L.phi); rb_mlp(L.w); rb_mlp(L.upd); rb(L.Ut); rb(L.Vt); }
  rb_mlp(m->readout); rb(m->Vout);
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->n_sms = prop.multiProcessorCount;
    e = cudaMalloc(&m->dev_err, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(m->dev_err, 0, 2 * sizeof(int));
    if (e != cudaSuccess) { tib_model_destroy(m); return fail("cudaMalloc(error word): %s", cudaGetErrorString(e)); }
  }
  if (F == 128) {
    // tensor-core weight stream: per layer kChunksPerLayer chunks in consumption order (tc_message.cuh)
    const size_t per_msg = (size_t)tib::tc::kChunksPerLayer * tib::tc::kChunkBytes;
    const size_t per_upd = (size_t)32 * tib::tc::kChunkBytes;      // 8 matrices (tc_update.cuh)
    const size_t per_layer = per_msg + per_upd;
    const size_t per_ro = (size_t)tib::tc::kRoChunks * tib::tc::kChunkBytes;   // readout W1 | W2 (tc_readout.cuh)
    std::vector<uint16_t> blob(per_layer / 2 * d->n_layers + per_ro / 2);
    for (int l = 0; l < d->n_layers; ++l) {
      uint16_t* out16 = blob.data() + per_layer / 2 * l;
      const float* pW1 = phi_src[l];                                   // [F][2F]
      const float* pW2 = pW1 + (size_t)F * 2 * F + 3 * F;              // [F][F]
      const float* pW3 = pW2 + (size_t)F * F + 3 * F;                  // [5F][F]
      const float* wW1 = w_src[l];                                     // [F][F]
      const float* wW2 = wW1 + (size_t)F * F + 3 * F;
      const float* wW3 = wW2 + (size_t)F * F + 3 * F;                  // [5F][F]
      auto mat = [&](const float* W, int ld, int row0, int col0) {     // one [128 x 128] block = 4 chunks
        for (int kb = 0; kb < 4; ++kb) { pack_tc_chunk(out16, W, ld, row0, col0 + 32 * kb); out16 += tib::tc::kChunkBytes / 2; }
      };
      mat(wW1, F, 0, 0);
      mat(pW1, 2 * F, 0, 0);
      mat(wW2, F, 0, 0);
      mat(pW1, 2 * F, 0, F);
      mat(pW2, F, 0, 0);
      for (int sp = 0; sp < 5; ++sp)            // output layer: phi and w chunks interleaved (tc_message.cuh)
        for (int kb = 0; kb < 4; ++kb) {
          pack_tc_chunk(out16, pW3, F, sp * F, 32 * kb); out16 += tib::tc::kChunkBytes / 2;
          pack_tc_chunk(out16, wW3, F, sp * F, 32 * kb); out16 += tib::tc::kChunkBytes / 2;
        }
      // update layer: U [F][F], V [F][F], MLP(2F -> F -> F -> 3F)
      const float* U = uv_src[l];
      const float* V = U + (size_t)F * F;
      const float* uW1 = V + (size_t)F * F;                           // [F][2F]
      const float* uW2 = uW1 + (size_t)F * 2 * F + 3 * F;
      const float* uW3 = uW2 + (size_t)F * F + 3 * F;                 // [3F][F]: rows g | a | c
      mat(V, F, 0, 0);
      mat(uW1, 2 * F, 0, 0);
      mat(uW1, 2 * F, 0, F);
      mat(uW2, F, 0, 0);
      mat(uW3, F, F, 0);
      mat(uW3, F, 2 * F, 0);
      mat(uW3, F, 0, 0);
      mat(U, F, 0, 0);
    }
    {
      uint16_t* out16 = blob.data() + per_layer / 2 * d->n_layers;
      const float* rW1 = ro_src;                                       // [F][F]
      const float* rW2 = rW1 + (size_t)F * F + 3 * F;
      for (const float* W : {rW1, rW2})
        for (int kb = 0; kb < 4; ++kb) { pack_tc_chunk(out16, W, F, 0, 32 * kb); out16 += tib::tc::kChunkBytes / 2; }
    }
    e = cudaMalloc(&m->tc_blob, blob.size() * 2);
    if (e == cudaSuccess) e = cudaMemcpy(m->tc_blob, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { tib_model_destroy(m); return fail("tensor-core weight upload: %s", cudaGetErrorString(e)); }
    for (int l = 0; l < d->n_layers; ++l) { m->layers[l].tc_msg = m->tc_blob + per_layer * l; m->layers[l].tc_upd = m->tc_blob + per_layer * l + per_msg; }
    m->tc_ro = m->tc_blob + per_layer * d->n_layers;
  }
  if (F == 128 || F == 256) {
    // layered path (tc_chain.cuh): per layer the streams of phi, w, [V; U] and the update MLP
    std::vector<uint16_t> blob;
    std::vector<size_t> offs;
    std::vector<float> vu((size_t)2 * F * F);
    for (int l = 0; l < d->n_layers; ++l) {
      auto& L = m->layers[l];
      offs.push_back(blob.size()); pack_chain_mlp(blob, phi_src[l], F, 2 * F, 5 * F, L.gmax_phi);
      offs.push_back(blob.size()); pack_chain_mlp(blob, w_src[l], F, F, 5 * F, L.gmax_w);
      const float* U = uv_src[l];
      const float* V = U + (size_t)F * F;
      std::memcpy(vu.data(), V, sizeof(float) * F * F);
      std::memcpy(vu.data() + (size_t)F * F, U, sizeof(float) * F * F);
      offs.push_back(blob.size()); pack_chain_matrix(blob, vu.data(), F, 2 * F, 0, F);
      offs.push_back(blob.size()); pack_chain_mlp(blob, V + (size_t)F * F, F, 2 * F, 3 * F, L.gmax_upd);
    }
    e = cudaMalloc(&m->ch_blob, blob.size() * 2);
    if (e == cudaSuccess) e = cudaMemcpy(m->ch_blob, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { tib_model_destroy(m); return fail("layered weight upload: %s", cudaGetErrorString(e)); }
    for (int l = 0; l < d->n_layers; ++l) {
      auto& L = m->layers[l];
      L.ch_phi = m->ch_blob + offs[4 * l] * 2; L.ch_w = m->ch_blob + offs[4 * l + 1] * 2;
      L.ch_uv = m->ch_blob + offs[4 * l + 2] * 2; L.ch_upd = m->ch_blob + offs[4 * l + 3] * 2;
    }
  }
  *out = m;
  return 0;
}

void tib_model_destroy(tib_model* m) {
  if (!m) return;
  if (m->dev) cudaFree(m->dev);
  if (m->tc_blob) cudaFree(m->tc_blob);
  if (m->ch_blob) cudaFree(m->ch_blob);
  if (m->dev_err) cudaFree(m->dev_err);
  if (m->dev_dbg) cudaFree(m->dev_dbg);
  delete m;
}

int tib_model_status(tib_model* m, void* stream) {
  if (!m) return fail("null model");
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  int h[2] = {0, 0};
  CUDA_TRY(cudaMemcpy(h, m->dev_err, 2 * sizeof(int), cudaMemcpyDeviceToHost));
  if (h[0] != 0 || h[1] != 0) {
    CUDA_TRY(cudaMemset(m->dev_err, 0, 2 * sizeof(int)));
    if (h[0] != 0)
      return fail("device pipeline error %d: an mbarrier wait timed out inside a tensor-core kernel (results are invalid)", h[0]);
    return fail("non-finite drift from the tensor-core path: node features beyond the split-f16 range (|x| >= 1.0e6) or "
                "non-finite inputs; TIB_MATH_FP32_SIMT has the full fp32 range");
  }
  return 0;
}

int tib_debug_counters(tib_model* m, int enable, long long* out, int max_ctas) {
  if (!m) return fail("null model");
  if (enable && !m->dev_dbg) {
    CUDA_TRY(cudaMalloc(&m->dev_dbg, sizeof(long long) * 8 * 2048));
    CUDA_TRY(cudaMemset(m->dev_dbg, 0, sizeof(long long) * 8 * 2048));
  }
  if (out && m->dev_dbg) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out, m->dev_dbg, sizeof(long long) * 8 * (size_t)std::min(max_ctas, 2048), cudaMemcpyDeviceToHost));
  }
  if (!enable && m->dev_dbg) { cudaFree(m->dev_dbg); m->dev_dbg = nullptr; }
  return 0;
}

int tib_model_set_math(tib_model* m, int math_mode) {
  if (!m) return fail("null model");
  if (math_mode < TIB_MATH_FP32_SIMT || math_mode > TIB_MATH_F16X3_LAYERED) return fail("unknown math mode %d", math_mode);
  const int F = m->d.n_features;
  if (math_mode != TIB_MATH_FP32_SIMT && F != 128 && F != 256)
    return fail("the tensor-core math modes are built for n_features = 128 and 256 (got %d)", F);
  m->math = math_mode;
  return 0;
}

size_t tib_workspace_bytes(const tib_model* m, int32_t n_mol, int32_t n_nodes, int64_t n_edges) {
  (void)n_mol;
  if (!m) return 0;
  return drift_ws_bytes(m, n_nodes, (long long)n_edges, TIB_MAX_ATOMS);
}

int tib_drift(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b, void* workspace,
              size_t workspace_bytes, void* stream) {
  if (check_batch(m, b)) return -1;
  if (!x || !out_b) return fail("tib_drift: null x/out");
  Workspace ws;
  if (prep_ws(m, b, workspace, workspace_bytes, ws)) return -1;
  return drift_dispatch(m, b, x, t, out_b, ws, (cudaStream_t)stream);
}

size_t tib_div_workspace_bytes(const tib_model* m, int32_t n_mol, int32_t n_nodes, int64_t n_edges, int32_t max_atoms) {
  (void)n_mol;
  if (!m) return 0;
  const int F = m->d.n_features;
  const size_t simt = TangentWs::bytes(F, jvp_dirs(F), n_nodes, (long long)n_edges);
  const bool tc = m->math != TIB_MATH_FP32_SIMT && (F == 128 || F == 256);
  const size_t lay = tc ? LayWs::bytes(F, n_nodes, (long long)n_edges, max_atoms, true) : 0;
  return Workspace::bytes(F, n_nodes, (long long)n_edges) + std::max(simt, lay);
}

}  // extern "C"
namespace {
// drift + exact divergence with the workspace already validated (shared by tib_drift_div and the (x, dlogp) rollouts)
int drift_div_dispatch(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b, float* out_div, void* workspace,
                       cudaStream_t st) {
  const int F = m->d.n_features;
  Workspace ws;
  ws.carve(workspace, F, b->n_nodes, (long long)b->n_edges);
  if (m->math != TIB_MATH_FP32_SIMT && (F == 128 || F == 256)) {
    LayWs lw{};
    lw.layout(ws.end, F, b->n_nodes, (long long)b->n_edges, b->max_atoms, true);
    return F == 128 ? drift_div_layered<128>(m, b, x, t, out_b, out_div, ws, lw, st)
                    : drift_div_layered<256>(m, b, x, t, out_b, out_div, ws, lw, st);
  }
  TangentWs tw;
  tw.carve(ws.end, F, jvp_dirs(F), b->n_nodes, (long long)b->n_edges);
  switch (F) {
    case 32: return drift_div_simt<32>(m, b, x, t, out_b, out_div, ws, tw, st);
    case 64: return drift_div_simt<64>(m, b, x, t, out_b, out_div, ws, tw, st);
    case 128: return drift_div_simt<128>(m, b, x, t, out_b, out_div, ws, tw, st);
    case 256: return drift_div_simt<256>(m, b, x, t, out_b, out_div, ws, tw, st);
  }
  return fail("unsupported n_features=%d", F);
}
}  // namespace
extern "C" {

int tib_drift_div(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b, float* out_div, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (check_batch(m, b)) return -1;
  if (!x || !out_b || !out_div) return fail("tib_drift_div: null x/out");
  const size_t need = tib_div_workspace_bytes(m, b->n_mol, b->n_nodes, b->n_edges, b->max_atoms);
  if (!workspace || workspace_bytes < need) return fail("divergence workspace too small: %zu < %zu bytes", workspace_bytes, need);
  if (((uintptr_t)workspace & 255) != 0) return fail("workspace must be 256-byte aligned");
  return drift_div_dispatch(m, b, x, t, out_b, out_div, workspace, (cudaStream_t)stream);
}

int tib_zmatrix(const float* x, int64_t n_conf, int32_t n_atoms, const int32_t* order, const int32_t* ref, float* z, void* stream) {
  if (!x || !order || !ref || !z) return fail("tib_zmatrix: null argument");
  if (n_conf < 0 || n_atoms < 2) return fail("tib_zmatrix: need n_conf >= 0 and n_atoms >= 2 (got %lld, %d)", (long long)n_conf, n_atoms);
  if (n_conf == 0) return 0;
  const long long total = (long long)n_conf * (n_atoms - 1);
  tib::k_zmatrix<<<grid_for((size_t)total), 256, 0, (cudaStream_t)stream>>>(x, (long long)n_conf, n_atoms, order, ref, z);
  LAUNCH_CHECK();
  return 0;
}

int tib_tica_project(const float* torsions, int64_t n_conf, int32_t n_tors, int64_t stride, int32_t col0, int32_t col_step,
                     const double* mean, const double* R, int32_t dim, const double* weight, float* proj, double* hist,
                     int32_t n_bins, double lo, double hi, void* stream) {
  if (!torsions || !mean || !R) return fail("tib_tica_project: null argument");
  if (dim < 1 || dim > 4) return fail("tib_tica_project: dim must be in [1,4] (got %d)", dim);
  if (n_conf < 0 || n_tors < 1) return fail("tib_tica_project: need n_conf >= 0 and n_tors >= 1");
  if (hist && (n_bins < 1 || n_bins > 1024 || !(hi > lo))) return fail("tib_tica_project: bad histogram range / bins");
  if (n_conf == 0) return 0;
  const size_t smem = sizeof(double) * (size_t)dim * (size_t)(hist ? n_bins : 1);
  tib::k_tica_project<<<grid_for((size_t)n_conf), 256, smem, (cudaStream_t)stream>>>(torsions, (long long)n_conf, n_tors, (long long)stride,
                                                                                  col0, col_step, mean, R, dim, weight, proj, hist,
                                                                                  hist ? n_bins : 1, lo, hi);
  LAUNCH_CHECK();
  return 0;
}

int tib_step_euler(const float* x, const float* b, const float* score, const float* noise, float dt, float eps,
                   float* x_out, float* frame, size_t n, void* stream) {
  if (!x || !b || !x_out) return fail("tib_step_euler: null pointer");
  if (n == 0) return 0;
  const float dt_eps = dt * eps;
  const float sig = sqrtf(2.0f * eps * dt);
  const uintptr_t bits = (uintptr_t)x | (uintptr_t)b | (uintptr_t)score | (uintptr_t)noise | (uintptr_t)x_out | (uintptr_t)frame;
  const int vec_ok = (bits & 15) == 0;
  { ProfScope ps(TIB_K_STEP, (cudaStream_t)stream);
    tib::k_step_euler<<<grid_for(vec_ok ? (n + 3) / 4 : n), 256, 0, (cudaStream_t)stream>>>(x, b, score, noise, dt, dt_eps, sig, x_out, frame, n, vec_ok); }
  LAUNCH_CHECK();
  return 0;
}

int tib_rollout_fixed(tib_model* m, const tib_batch* b, const float* x0, const tib_fixed_opts* o, float* out_xts,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (check_batch(m, b)) return -1;
  if (!x0 || !o || !out_xts || !o->t_grid) return fail("tib_rollout_fixed: null argument");
  if (o->n_times < 1) return fail("n_times must be >= 1");
  if (o->method != TIB_METHOD_EULER && (o->eps != 0.0f || o->noise || o->score_model))
    return fail("Euler-Maruyama terms are only defined for TIB_METHOD_EULER");
  Workspace ws;
  if (prep_ws(m, b, workspace, workspace_bytes, ws)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)b->n_nodes * 3;
  const size_t ks = Workspace::kstride(b->n_nodes);
  float* y = ws.ycur;
  CUDA_TRY(cudaMemcpyAsync(y, x0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (o->save_frames) CUDA_TRY(cudaMemcpyAsync(out_xts, x0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  for (int j = 1; j < o->n_times; ++j) {
    const float t0 = o->t_grid[j - 1], t1 = o->t_grid[j];
    const float dt = t1 - t0;   // fp32, as `dt = t1 - t0` on the fp32 grid tensor
    float* frame = o->save_frames ? out_xts + (size_t)j * n : nullptr;
    if (o->method == TIB_METHOD_EULER) {
      if (drift_dispatch(m, b, y, t0, ws.drift, ws, st)) return -1;
      const float* sc = nullptr;
      if (o->score_model && o->eps != 0.0f) {
        if (o->score_model->d.n_features != m->d.n_features) return fail("score model must share n_features");
        if (drift_dispatch(o->score_model, b, y, t0, ws.score, ws, st)) return -1;
        sc = ws.score;
      }
      const float* nz = (o->noise && o->eps != 0.0f) ? o->noise + (size_t)(j - 1) * n : nullptr;
      if (tib_step_euler(y, ws.drift, sc, nz, dt, o->eps, y, frame, n, stream)) return -1;
    } else if (o->method == TIB_METHOD_MIDPOINT) {
      const float half_dt = 0.5f * dt;
      if (drift_dispatch(m, b, y, t0, ws.drift, ws, st)) return -1;
      if (tib_step_euler(y, ws.drift, nullptr, nullptr, half_dt, 0.f, ws.ytmp, nullptr, n, stream)) return -1;
      if (drift_dispatch(m, b, ws.ytmp, t0 + half_dt, ws.drift, ws, st)) return -1;
      if (tib_step_euler(y, ws.drift, nullptr, nullptr, dt, 0.f, y, frame, n, stream)) return -1;
    } else if (o->method == TIB_METHOD_RK4) {
      float* k1 = ws.k; float* k2 = ws.k + ks; float* k3 = ws.k + 2 * ks; float* k4 = ws.k + 3 * ks;
      const float third = (float)(1.0 / 3.0), two_thirds = (float)(2.0 / 3.0);
      const int g = grid_for(n);
      if (drift_dispatch(m, b, y, t0, k1, ws, st)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(0, y, k1, k2, k3, k4, dt, ws.ytmp, nullptr, n); LAUNCH_CHECK();
      if (drift_dispatch(m, b, ws.ytmp, t0 + dt * third, k2, ws, st)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(1, y, k1, k2, k3, k4, dt, ws.ytmp, nullptr, n); LAUNCH_CHECK();
      if (drift_dispatch(m, b, ws.ytmp, t0 + dt * two_thirds, k3, ws, st)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(2, y, k1, k2, k3, k4, dt, ws.ytmp, nullptr, n); LAUNCH_CHECK();
      if (drift_dispatch(m, b, ws.ytmp, t1, k4, ws, st)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(3, y, k1, k2, k3, k4, dt, y, frame, n); LAUNCH_CHECK();
    } else {
      return fail("unknown fixed-grid method %d", o->method);
    }
  }
  if (!o->save_frames) CUDA_TRY(cudaMemcpyAsync(out_xts, y, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
// ---- dopri5 (torchdiffeq 0.2.5 RKAdaptiveStepsizeODESolver; restated in oracle/ode_oracle.py) ----
namespace {
const double DP_ALPHA[6] = {1 / 5., 3 / 10., 4 / 5., 8 / 9., 1., 1.};
const double DP_BETA[6][6] = {
    {1 / 5.},
    {3 / 40., 9 / 40.},
    {44 / 45., -56 / 15., 32 / 9.},
    {19372 / 6561., -25360 / 2187., 64448 / 6561., -212 / 729.},
    {9017 / 3168., -355 / 33., 46732 / 5247., 49 / 176., -5103 / 18656.},
    {35 / 384., 0, 500 / 1113., 125 / 192., -2187 / 6784., 11 / 84.}};
const double DP_C_ERROR[7] = {35 / 384. - 1951 / 21600., 0, 500 / 1113. - 22642 / 50085., 125 / 192. - 451 / 720.,
                              -2187 / 6784. - -12231 / 42400., 11 / 84. - 649 / 6300., -1. / 60.};
const double DP_C_MID[7] = {6025192743. / 30085553152. / 2, 0, 51252292925. / 65400821598. / 2,
                            -2691868925. / 45128329728. / 2, 187940372067. / 1594534317056. / 2,
                            -1776094331. / 19743644256. / 2, 11237099. / 235043384. / 2};

// state buffers of a solver over a flat fp32 state of n floats (x alone, or [x | dlogp])
struct StateBufs { float *ycur, *ynew, *ytmp, *k; size_t ks; double *partial, *scalar; };

// torchdiffeq's norm of the flattened state: RMS over everything (tensor state), or the MAX of the per-component RMS norms
// when the state is the tuple (x, dlogp) flattened with `split` = x.numel() (rk_common / misc._mixed_norm)
struct Reducer {
  StateBufs* sb; cudaStream_t st; size_t n, split; const tib_dopri5_opts* o;
  template <typename Launch>   // launch(offset, count) fills sb->partial for the elements [offset, offset + count)
  int norm(Launch&& launch, double* out) {
    const int parts = split ? 2 : 1;
    for (int p = 0; p < parts; ++p) {
      const size_t off = p == 0 ? 0 : split, cnt = split ? (p == 0 ? split : n - split) : n;
      if (launch(off, cnt)) return -1;
      tib::k_reduce_partials<<<1, 256, 0, st>>>(sb->partial, Workspace::kPartials, sb->scalar + p);
      LAUNCH_CHECK();
    }
    double h[2] = {0.0, 0.0};
    CUDA_TRY(cudaMemcpyAsync(h, sb->scalar, sizeof(double) * parts, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    double best = 0.0;
    for (int p = 0; p < parts; ++p) {
      double v[2] = {h[p], (double)(split ? (p == 0 ? split : n - split) : n)};
      if (o->norm_allreduce) o->norm_allreduce(v, o->norm_user);
      best = std::max(best, std::sqrt(v[0] / v[1]));
      if (!(v[0] == v[0])) best = v[0];      // NaN propagates
    }
    *out = best;
    return 0;
  }
};

// torchdiffeq 0.2.5 RKAdaptiveStepsizeODESolver (dopri5) over a flat state; rhs(y, t, out) evaluates the right-hand side
template <typename Rhs>
int dopri5_impl(Rhs&& rhs, size_t n, size_t split, const float* y0, const tib_dopri5_opts* o, float* out, StateBufs& sb,
                tib_dopri5_stats* stats, cudaStream_t st, void* stream) {
  const size_t ks = sb.ks;
  const int g = grid_for(n);
  const int gp = Workspace::kPartials;   // reduction kernels use exactly kPartials blocks
  const double rtol = o->rtol, atol = o->atol;
  const int max_attempts = o->max_attempts > 0 ? o->max_attempts : 100000;
  Reducer red{&sb, st, n, split, o};
  int nfe = 0, attempts = 0, accepted = 0;
  float* y = sb.ycur;
  float* ynew = sb.ynew;
  float* k = sb.k;   // k_s at k + s*ks
  CUDA_TRY(cudaMemcpyAsync(y, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (o->save_frames) CUDA_TRY(cudaMemcpyAsync(out, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));

  // --- _before_integrate: f0 and Hairer's initial step (misc._select_initial_step, order = 4)
  const double tstart = (double)o->t_grid[0];
  if (rhs(y, (float)tstart, k)) return -1;
  ++nfe;
  double d0, d1, d2, h0, h1, dt;
  if (red.norm([&](size_t off, size_t cnt) { tib::k_scaled_sq<<<gp, 256, 0, st>>>(y + off, nullptr, y + off, rtol, atol, sb.partial, cnt); LAUNCH_CHECK(); return 0; }, &d0)) return -1;
  if (red.norm([&](size_t off, size_t cnt) { tib::k_scaled_sq<<<gp, 256, 0, st>>>(k + off, nullptr, y + off, rtol, atol, sb.partial, cnt); LAUNCH_CHECK(); return 0; }, &d1)) return -1;
  if (d0 < 1e-5 || d1 < 1e-5) h0 = (double)1e-6f; else h0 = 0.01 * d0 / d1;
  h0 = std::fabs(h0);
  // y1 = y0 + h0*f0 (h0 is a 0-dim fp64 tensor times an fp32 tensor -> fp32 arithmetic)
  if (tib_step_euler(y, k, nullptr, nullptr, (float)h0, 0.f, sb.ytmp, nullptr, n, stream)) return -1;
  if (rhs(sb.ytmp, (float)(tstart + h0), k + ks)) return -1;
  ++nfe;
  if (red.norm([&](size_t off, size_t cnt) { tib::k_scaled_sq<<<gp, 256, 0, st>>>(k + ks + off, k + off, y + off, rtol, atol, sb.partial, cnt); LAUNCH_CHECK(); return 0; }, &d2)) return -1;
  d2 = std::fabs(d2 / h0);
  if (d1 <= 1e-15 && d2 <= 1e-15) h1 = std::max((double)1e-6f, h0 * 1e-3);
  else h1 = std::pow(0.01 / std::max(d1, d2), 1.0 / 5.0);
  dt = std::min(100 * h0, std::fabs(h1));

  double t0 = tstart, t1 = tstart;   // rk_state.t0, rk_state.t1
  float dt_step = 0.f;               // fp32 dt of the last accepted step (for the dense output)
  tib::StageCoef cmid{}; cmid.n = 7;
  for (int i = 1; i < o->n_times; ++i) {
    const double tn = (double)o->t_grid[i];
    // collect the frames that fall inside the current accepted step before stepping further
    while (tn > t1) {
      if (attempts >= max_attempts) return fail("dopri5: exceeded %d attempted steps", max_attempts);
      const double t_new = t1 + dt;
      if (!(t1 + dt > t1)) return fail("dopri5: underflow in dt %g", dt);
      const float t0_s = (float)t1, dt_s = (float)dt, t1_s = (float)t_new;
      for (int s = 0; s < 6; ++s) {
        tib::StageCoef c{}; c.n = s + 1;
        for (int q = 0; q <= s; ++q) c.c[q] = (float)DP_BETA[s][q] * dt_s;   // beta_i (fp32) * dt (fp32)
        float* yi = (s == 5) ? ynew : sb.ytmp;
        tib::k_dopri_stage<<<g, 256, 0, st>>>(y, k, ks, c, yi, n); LAUNCH_CHECK();
        float ti;
        if (DP_ALPHA[s] == 1.0) ti = std::nextafterf(t1_s, t1_s - 1.0f);   // Perturb.PREV
        else ti = t0_s + (float)DP_ALPHA[s] * dt_s;
        if (rhs(yi, ti, k + (size_t)(s + 1) * ks)) return -1;
        ++nfe;
      }
      tib::StageCoef ce{}; ce.n = 7;
      for (int q = 0; q < 7; ++q) ce.c[q] = dt_s * (float)DP_C_ERROR[q];
      double ratio;
      if (red.norm([&](size_t off, size_t cnt) { tib::k_dopri_error<<<gp, 256, 0, st>>>(y + off, ynew + off, k + off, ks, ce, rtol, atol, sb.partial, cnt); LAUNCH_CHECK(); return 0; }, &ratio)) return -1;
      ++attempts;
      if (!(ratio == ratio)) return fail("dopri5: non-finite error ratio (state diverged)");
      if (ratio <= 1.0) {
        // accepted: the dense-output kernel needs (y, ynew, k) of THIS step; keep them by flushing the
        // frames that fall inside [t1, t_new] right away.
        ++accepted;
        for (int q = 0; q < 7; ++q) cmid.c[q] = dt_s * (float)DP_C_MID[q];
        dt_step = dt_s;
        t0 = t1; t1 = t_new;
        // frames i.. with t_grid <= t1
        int j = i;
        while (j < o->n_times && (double)o->t_grid[j] <= t1) {
          tib::DenseArgs da{}; da.n = 0;
          while (j < o->n_times && (double)o->t_grid[j] <= t1 && da.n < 8) {
            const double xx = ((double)o->t_grid[j] - t0) / (t1 - t0);
            da.xs[da.n] = (float)xx;
            da.frames[da.n] = o->save_frames ? out + (size_t)j * n : ((j == o->n_times - 1) ? out : nullptr);
            if (da.frames[da.n]) ++da.n;
            ++j;
          }
          if (da.n > 0) { tib::k_dopri_dense<<<g, 256, 0, st>>>(y, ynew, k, ks, cmid, dt_step, da, n); LAUNCH_CHECK(); }
        }
        // FSAL: f0 <- f1 (k_6 -> k_0), y <- ynew
        CUDA_TRY(cudaMemcpyAsync(k, k + 6 * ks, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
        std::swap(y, ynew);
      }
      // _optimal_step_size(dt, ratio, safety=0.9, ifactor=10, dfactor=0.2, order=5)
      if (ratio == 0.0) dt = dt * 10.0;
      else {
        const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
        const double factor = std::min(10.0, std::max(0.9 / std::pow(ratio, 0.2), dfactor));
        dt = dt * factor;
      }
    }
  }
  // n_times == 1: no step is taken and the dense-output kernel never runs; the final state is the initial state
  if (!o->save_frames && o->n_times == 1) CUDA_TRY(cudaMemcpyAsync(out, y, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (stats) { stats->nfe = nfe; stats->attempts = attempts; stats->accepted = accepted; stats->last_dt = dt; }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// torchdiffeq FixedGridODESolver (euler / midpoint / rk4 = 3/8 rule) over a flat state, the output grid is the step grid
template <typename Rhs>
int fixed_impl(Rhs&& rhs, size_t n, const float* y0, const tib_fixed_opts* o, float* out, StateBufs& sb, cudaStream_t st, void* stream) {
  const size_t ks = sb.ks;
  float* y = sb.ycur;
  float* f = sb.k + 6 * ks;     // right-hand side scratch
  CUDA_TRY(cudaMemcpyAsync(y, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (o->save_frames) CUDA_TRY(cudaMemcpyAsync(out, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  for (int j = 1; j < o->n_times; ++j) {
    const float t0 = o->t_grid[j - 1], t1 = o->t_grid[j];
    const float dt = t1 - t0;   // fp32, as `dt = t1 - t0` on the fp32 grid tensor (negative on a decreasing grid)
    float* frame = o->save_frames ? out + (size_t)j * n : nullptr;
    if (o->method == TIB_METHOD_EULER) {
      if (rhs(y, t0, f)) return -1;
      if (tib_step_euler(y, f, nullptr, nullptr, dt, 0.f, y, frame, n, stream)) return -1;
    } else if (o->method == TIB_METHOD_MIDPOINT) {
      const float half_dt = 0.5f * dt;
      if (rhs(y, t0, f)) return -1;
      if (tib_step_euler(y, f, nullptr, nullptr, half_dt, 0.f, sb.ytmp, nullptr, n, stream)) return -1;
      if (rhs(sb.ytmp, t0 + half_dt, f)) return -1;
      if (tib_step_euler(y, f, nullptr, nullptr, dt, 0.f, y, frame, n, stream)) return -1;
    } else if (o->method == TIB_METHOD_RK4) {
      float* k1 = sb.k; float* k2 = sb.k + ks; float* k3 = sb.k + 2 * ks; float* k4 = sb.k + 3 * ks;
      const float third = (float)(1.0 / 3.0), two_thirds = (float)(2.0 / 3.0);
      const int g = grid_for(n);
      if (rhs(y, t0, k1)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(0, y, k1, k2, k3, k4, dt, sb.ytmp, nullptr, n); LAUNCH_CHECK();
      if (rhs(sb.ytmp, t0 + dt * third, k2)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(1, y, k1, k2, k3, k4, dt, sb.ytmp, nullptr, n); LAUNCH_CHECK();
      if (rhs(sb.ytmp, t0 + dt * two_thirds, k3)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(2, y, k1, k2, k3, k4, dt, sb.ytmp, nullptr, n); LAUNCH_CHECK();
      if (rhs(sb.ytmp, t1, k4)) return -1;
      tib::k_rk4_stage<<<g, 256, 0, st>>>(3, y, k1, k2, k3, k4, dt, y, frame, n); LAUNCH_CHECK();
    } else {
      return fail("unknown fixed-grid method %d", o->method);
    }
  }
  if (!o->save_frames) CUDA_TRY(cudaMemcpyAsync(out, y, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// state buffers of the (x, dlogp) rollouts: carved after the divergence workspace
size_t tuple_state_stride(size_t n) { return Workspace::align(sizeof(float) * n) / sizeof(float); }
size_t tuple_state_bytes(size_t n) {
  return 10 * tuple_state_stride(n) * sizeof(float) + Workspace::align(sizeof(double) * Workspace::kPartials * 5) + Workspace::align(sizeof(double) * 8);
}
void tuple_state_carve(char* base, size_t n, StateBufs& sb) {
  sb.ks = tuple_state_stride(n);
  float* f = (float*)base;
  sb.ycur = f; sb.ynew = f + sb.ks; sb.ytmp = f + 2 * sb.ks; sb.k = f + 3 * sb.ks;      // k: 7 slots
  char* p = base + 10 * sb.ks * sizeof(float);
  sb.partial = (double*)p; p += Workspace::align(sizeof(double) * Workspace::kPartials * 5);
  sb.scalar = (double*)p;
}
}  // namespace
extern "C" {

int tib_rollout_dopri5(tib_model* m, const tib_batch* b, const float* x0, const tib_dopri5_opts* o, float* out_xts,
                       tib_dopri5_stats* stats, void* workspace, size_t workspace_bytes, void* stream) {
  if (check_batch(m, b)) return -1;
  if (!x0 || !o || !out_xts || !o->t_grid) return fail("tib_rollout_dopri5: null argument");
  if (o->n_times < 1) return fail("n_times must be >= 1");
  for (int i = 1; i < o->n_times; ++i)
    if (!(o->t_grid[i] > o->t_grid[i - 1])) return fail("t_grid must be strictly increasing");
  Workspace ws;
  if (prep_ws(m, b, workspace, workspace_bytes, ws)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)b->n_nodes * 3;
  StateBufs sb{ws.ycur, ws.ynew, ws.ytmp, ws.k, Workspace::kstride(b->n_nodes), ws.partial, ws.scalar};
  auto rhs = [&](const float* y, float t, float* out) { return drift_dispatch(m, b, y, t, out, ws, st); };
  if (dopri5_impl(rhs, n, 0, x0, o, out_xts, sb, stats, st, stream)) return -1;
  // the stream is idle here anyway: report a device-side pipeline / range fault instead of returning bad frames
  return tib_model_status(m, stream);
}

/* ---- (x, dlogp) rollouts: the state of MoleculeIntegrator(return_dlogp=True) -------------------------------------------- */
namespace {
// rhs over y = [x | dlogp]: out = [b | -div * scale]  (reverse: [-b | +div * scale]; ode_wrapper.py:39-49, latent :38-46)
// mult_b / mult_d carry the wrapper's signs and scale; time_sign = -1 solves a decreasing grid as torchdiffeq does
// (odeint on (-t, -f), _check_inputs): the solver sees the negated, increasing grid
struct DlogpRhs {
  tib_model* m; const tib_batch* b; void* workspace; cudaStream_t st; size_t n3; float mult_b, mult_d, time_sign;
  int operator()(const float* y, float t, float* out) const {
    if (drift_div_dispatch(m, b, y, time_sign * t, out, out + n3, workspace, st)) return -1;
    tib::k_dlogp_rhs_finish<<<grid_for(n3 + (size_t)b->n_mol), 256, 0, st>>>(out, n3, (size_t)b->n_mol, mult_b * time_sign, mult_d * time_sign);
    LAUNCH_CHECK();
    return 0;
  }
};
int check_dlogp_args(tib_model* m, const tib_batch* b, const float* y0, const void* o, const float* out, void* workspace, size_t workspace_bytes) {
  if (check_batch(m, b)) return -1;
  if (!y0 || !o || !out) return fail("dlogp rollout: null argument");
  const size_t need = tib_div_rollout_workspace_bytes(m, b->n_mol, b->n_nodes, b->n_edges, b->max_atoms);
  if (!workspace || workspace_bytes < need) return fail("dlogp rollout workspace too small: %zu < %zu bytes", workspace_bytes, need);
  if (((uintptr_t)workspace & 255) != 0) return fail("workspace must be 256-byte aligned");
  return 0;
}
}  // namespace

size_t tib_div_rollout_workspace_bytes(const tib_model* m, int32_t n_mol, int32_t n_nodes, int64_t n_edges, int32_t max_atoms) {
  if (!m) return 0;
  return Workspace::align(tib_div_workspace_bytes(m, n_mol, n_nodes, n_edges, max_atoms)) + tuple_state_bytes((size_t)n_nodes * 3 + (size_t)n_mol);
}

int tib_rollout_fixed_dlogp(tib_model* m, const tib_batch* b, const float* y0, const tib_fixed_opts* o, float mult_b, float mult_d,
                            float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (check_dlogp_args(m, b, y0, o, out, workspace, workspace_bytes)) return -1;
  if (!o->t_grid || o->n_times < 1) return fail("tib_rollout_fixed_dlogp: bad time grid");
  if (o->eps != 0.0f || o->noise || o->score_model) return fail("Euler-Maruyama terms and the dlogp state are mutually exclusive");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n3 = (size_t)b->n_nodes * 3, n = n3 + (size_t)b->n_mol;
  StateBufs sb;
  tuple_state_carve((char*)workspace + Workspace::align(tib_div_workspace_bytes(m, b->n_mol, b->n_nodes, b->n_edges, b->max_atoms)), n, sb);
  DlogpRhs rhs{m, b, workspace, st, n3, mult_b, mult_d, 1.0f};
  return fixed_impl(rhs, n, y0, o, out, sb, st, stream);
}

int tib_rollout_dopri5_dlogp(tib_model* m, const tib_batch* b, const float* y0, const tib_dopri5_opts* o, float mult_b, float mult_d,
                             float time_sign, float* out, tib_dopri5_stats* stats, void* workspace, size_t workspace_bytes, void* stream) {
  if (check_dlogp_args(m, b, y0, o, out, workspace, workspace_bytes)) return -1;
  if (!o->t_grid || o->n_times < 1) return fail("tib_rollout_dopri5_dlogp: bad time grid");
  for (int i = 1; i < o->n_times; ++i)
    if (!(o->t_grid[i] > o->t_grid[i - 1])) return fail("t_grid must be strictly increasing (a decreasing grid is passed negated, with time_sign = -1)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n3 = (size_t)b->n_nodes * 3, n = n3 + (size_t)b->n_mol;
  StateBufs sb;
  tuple_state_carve((char*)workspace + Workspace::align(tib_div_workspace_bytes(m, b->n_mol, b->n_nodes, b->n_edges, b->max_atoms)), n, sb);
  DlogpRhs rhs{m, b, workspace, st, n3, mult_b, mult_d, time_sign < 0.0f ? -1.0f : 1.0f};
  if (dopri5_impl(rhs, n, n3, y0, o, out, sb, stats, st, stream)) return -1;
  return tib_model_status(m, stream);
}

int tib_reweight_stats(const double* E0, const double* E1, const double* nd, const double* wt, size_t n, double* out,
                       void* stream) {
  if (!E0 || !E1 || !out) return fail("tib_reweight_stats: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = grid_for(n);
  double* partial = nullptr;
  CUDA_TRY(cudaMallocAsync(&partial, sizeof(double) * 5 * blocks, st));
  tib::k_reweight_partials<<<blocks, 256, 0, st>>>(E0, E1, nd, wt, n, partial); LAUNCH_CHECK();
  tib::k_reweight_final<<<1, 32, 0, st>>>(partial, blocks, out); LAUNCH_CHECK();
  CUDA_TRY(cudaFreeAsync(partial, st));
  return 0;
}

// ---- ADW -----------------------------------------------------------------------------------------
struct tib_adw_model {
  int hidden, num_layers, device;
  bool attr_set = false;
  double* dev;
  tib::AdwW w;
};

int tib_adw_create(tib_adw_model** out, int32_t hidden, int32_t num_layers, const double* w, size_t n_doubles, int device) {
  if (!out || !w) return fail("tib_adw_create: null argument");
  if (hidden != 256) return fail("ADW kernel is built for hidden_size 256 (adw/config/settings.json:6), got %d", hidden);
  if (num_layers < 1 || num_layers > 8) return fail("num_layers must be in [1,8]");
  const size_t H = hidden;
  const size_t need = (3 * H + H) + (H * H + H) + (H + 1) + (3 * H + H) + (size_t)(num_layers - 1) * (H * H + H) + (H + 1);
  if (n_doubles != need) return fail("ADW packed weight count mismatch: got %zu, need %zu", n_doubles, need);
  DeviceGuard guard(device);
  if (!guard.ok) return fail("cudaSetDevice(%d) failed", device);
  // device layout: hidden->hidden matrices transposed to [in][out]; everything else as given
  std::vector<double> stage;
  stage.reserve(n_doubles);
  tib_adw_model* m = new tib_adw_model();
  m->hidden = hidden; m->num_layers = num_layers; m->device = device; m->dev = nullptr;
  std::vector<size_t> offs;
  auto push = [&](const double* v, size_t n) { offs.push_back(stage.size()); stage.insert(stage.end(), v, v + n); };
  auto push_T = [&](const double* W) {
    offs.push_back(stage.size());
    size_t off = stage.size(); stage.resize(off + H * H);
    for (size_t o = 0; o < H; ++o) for (size_t k = 0; k < H; ++k) stage[off + k * H + o] = W[o * H + k];
  };
  const double* src = w;
  // beta_embed: Linear(3,H), Linear(H,H), Linear(H,1)
  push(src, 3 * H); src += 3 * H; push(src, H); src += H;
  push_T(src); src += H * H; push(src, H); src += H;
  push(src, H); src += H; push(src, 1); src += 1;
  // net: Linear(3,H), (num_layers-1) x Linear(H,H), Linear(H,1)
  push(src, 3 * H); src += 3 * H; push(src, H); src += H;
  for (int l = 0; l < num_layers - 1; ++l) { push_T(src); src += H * H; push(src, H); src += H; }
  push(src, H); src += H; push(src, 1); src += 1;
  cudaError_t e = cudaMalloc(&m->dev, sizeof(double) * stage.size());
  if (e != cudaSuccess) { delete m; return fail("cudaMalloc: %s", cudaGetErrorString(e)); }
  e = cudaMemcpy(m->dev, stage.data(), sizeof(double) * stage.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(m->dev); delete m; return fail("cudaMemcpy: %s", cudaGetErrorString(e)); }
  size_t i = 0;
  auto next = [&]() { return (const double*)(m->dev + offs[i++]); };
  m->w.e_W1 = next(); m->w.e_b1 = next(); m->w.e_W2t = next(); m->w.e_b2 = next(); m->w.e_W3 = next(); m->w.e_b3 = next();
  m->w.n_W1 = next(); m->w.n_b1 = next();
  m->w.n_hidden = num_layers - 1;
  for (int l = 0; l < num_layers - 1; ++l) { m->w.n_Wt[l] = next(); m->w.n_b[l] = next(); }
  m->w.n_Wo = next(); m->w.n_bo = next();
  *out = m;
  return 0;
}

void tib_adw_destroy(tib_adw_model* m) {
  if (!m) return;
  if (m->dev) cudaFree(m->dev);
  delete m;
}

int tib_adw_drift_div(tib_adw_model* m, const double* x, const double* beta0, const double* beta1, float t,
                      double* out_b, double* out_div, size_t n, void* stream) {
  if (!m || !x || !beta0 || !beta1 || !out_b) return fail("tib_adw_drift_div: null pointer");
  if (n == 0) return 0;
  if (!m->attr_set) {          // per model (= per device): the opt-in shared-memory size is a per-device function attribute
    if (set_smem(tib::k_adw<256>, tib::adw_smem<256>())) return -1;
    m->attr_set = true;
  }
  const int blocks = (int)((n + tib::kAdwRows - 1) / tib::kAdwRows);
  tib::k_adw<256><<<blocks, 256, tib::adw_smem<256>(), (cudaStream_t)stream>>>(m->w, x, beta0, beta1, t, out_b, out_div, n);
  LAUNCH_CHECK();
  return 0;
}

// Tensor-core plumbing self test (see k_tc_selftest): A, out are DEVICE fp32 [128][128]; W is HOST fp32 [128][128].
int tib_selftest_gemm(const float* A, const float* W_host, float* out, int transposed, void* stream) {
  if (!A || !W_host || !out) return fail("tib_selftest_gemm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<uint16_t> chunks(4 * tib::tc::kChunkBytes / 2);
  for (int kb = 0; kb < 4; ++kb) pack_tc_chunk(chunks.data() + (size_t)kb * tib::tc::kChunkBytes / 2, W_host, 128, 0, 32 * kb);
  struct Tmp { unsigned char* dW = nullptr; int* derr = nullptr; ~Tmp() { cudaFree(dW); cudaFree(derr); } } tmp;   // freed on every return path
  CUDA_TRY(cudaMalloc(&tmp.dW, chunks.size() * 2));
  CUDA_TRY(cudaMalloc(&tmp.derr, sizeof(int)));
  CUDA_TRY(cudaMemset(tmp.derr, 0, sizeof(int)));
  CUDA_TRY(cudaMemcpy(tmp.dW, chunks.data(), chunks.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = tib::tc::kOperandBytes + tib::tc::kSelfStages * tib::tc::kChunkBytes + 256;
  if (set_smem(tib::tc::k_tc_selftest, smem)) return -1;
  tib::tc::k_tc_selftest<<<1, tib::tc::kSelfThreads, smem, st>>>(A, tmp.dW, out, transposed, tmp.derr);
  LAUNCH_CHECK();
  CUDA_TRY(cudaStreamSynchronize(st));
  int h = 0;
  CUDA_TRY(cudaMemcpy(&h, tmp.derr, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) return fail("tib_selftest_gemm: an mbarrier wait timed out (pipeline protocol error)");
  return 0;
}

}  // extern "C"
