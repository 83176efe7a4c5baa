// train_api.cu - second translation unit of libtib.so: the training step of the ambient drift network
// (mdqm9/train_ambient.py:124-148, mdqm9/thermo/ambient/losses.py:30-85,126-133):
//   tib_train_loss_grad : interpolant + both antithetic drift evaluations + loss + hand-written adjoints of every stage
//   tib_adam_step       : clip_grad_norm_ + torch.optim.Adam on the flat weight vector
// Dense contractions run on tcgen05 (train_gemm.cuh, split-f16 x3), everything else in train.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <cstdint>
#include <vector>

#include "../../include/tib.h"
#include "train.cuh"
#include "train_gemm.cuh"

namespace tib_internal {      // defined in tib_api.cu (thread-local error string and launch counter of the library)
int set_error(const char* msg);
void count_launches(uint64_t n);
void* prof_open(int kind, void* stream);
void prof_close(void* h);
}  // namespace tib_internal

namespace {

using namespace tib::train;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return tib_internal::set_error(buf);
}

thread_local long long g_trace_rows = 0;      // row count of the MLP being processed (trace labels only)
struct Prof {      // tib_profile_begin / tib_profile_end: device time per kernel class; TIB_TRAIN_TRACE: per-launch times on stderr
  void* h;
  const char* name; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr;
  static bool tracing() { static const bool t = getenv("TIB_TRAIN_TRACE") != nullptr; return t; }
  Prof(int kind, cudaStream_t s, const char* nm = nullptr) : h(tib_internal::prof_open(kind, s)), name(nm), st(s) {
    if (name && tracing()) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaStreamSynchronize(st); cudaEventRecord(e0, st); }
  }
  ~Prof() {
    if (h) tib_internal::prof_close(h);
    if (e0) {
      float ms = 0.0f;
      cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      fprintf(stderr, "[kern] %.*s_rows%lld  %.1f us\n", (int)strcspn(name, "<"), name, g_trace_rows, ms * 1e3);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
  }
};
thread_local double g_gemm_flops = 0.0;
thread_local long long* g_gemm_dbg = nullptr;      // device buffer [64] when tib_gemm_debug is on

#define LAUNCH_CHECK()                                                                        \
  do {                                                                                        \
    tib_internal::count_launches(1);                                                          \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
// Launches that cannot fill the GPU (at most one CTA per SM) carry programmatic stream serialisation: the kernel may be
// scheduled - and run its prologue up to griddepcontrol.wait - while the kernel before it on the stream drains (every kernel
// triggers its dependents at entry).  Inside the captured graph this takes dependent-launch latency off the small nodes of
// the main chain (12 molecules: 2.17 -> 2.05 ms per step).  Large grids do not get it: their early CTAs would sit on shared
// memory and TMEM the running kernel's remaining CTAs need (256 molecules with every launch early: 4.75 -> 4.94 ms).
bool pdl_enabled() { static const bool on = getenv("TIB_TRAIN_NO_PDL") == nullptr; return on; }
int pdl_max_ctas() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl_enabled() && ctas <= pdl_max_ctas()) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);      // errors surface in LAUNCH_CHECK (cudaGetLastError)
}

// an element-wise / scatter kernel of the training step: profile scope (kernel class "train_other"; with TIB_TRAIN_TRACE the
// launch is timed synchronously and printed under the kernel's name) + launch check
#define TRAIN_LAUNCH(stream, ...)                                                             \
  do {                                                                                        \
    Prof pf(TIB_K_TRAIN_OTHER, stream, #__VA_ARGS__);                                         \
    __VA_ARGS__;                                                                              \
    LAUNCH_CHECK();                                                                           \
  } while (0)
#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define TRY(expr) do { if ((expr) != 0) return -1; } while (0)

constexpr float kStateScale = tib::tc::kStateScale;

// ---- offsets into the flat weight / gradient vector (order of tib_packed_weight_count, include/tib.h) ------------------------
// images of a weight matrix as a pre-packed B operand (train_gemm.cuh::k_pack_operand): byte offsets into the image blob
struct ImgOff { size_t fwd, tr; int fwd_chunks, tr_chunks; bool ok; };
struct MlpOff { size_t W1, b1, g1, be1, W2, b2, g2, be2, W3, b3; int k_in, n_out; ImgOff i1, i2, i3; };
struct LayerOff { MlpOff phi, w, upd; size_t UV; ImgOff iuv; };          // U [F][F] then V [F][F]: one [2F][F] matrix
struct Offsets {
  size_t edge_emb, atom_emb; MlpOff combine; std::vector<LayerOff> layers; MlpOff readout; size_t Vout; size_t total;
  PackTable pack; size_t img_bytes;
};

ImgOff add_image(Offsets& o, size_t src, int ld, int rows_o, int cols_i) {
  ImgOff im{};
  if (o.pack.n >= kPackMaxEntries) return im;               // very deep networks: the remaining matrices are built on the fly
  const int fwd_tiles = (rows_o + 127) / 128, tr_tiles = (cols_i + 127) / 128;
  im.fwd_chunks = (cols_i + kGemmKC - 1) / kGemmKC;
  im.tr_chunks = (rows_o + kGemmKC - 1) / kGemmKC;
  im.fwd = o.img_bytes;
  im.tr = im.fwd + (size_t)fwd_tiles * im.fwd_chunks * 2 * kGemmHalf;
  o.img_bytes = im.tr + (size_t)tr_tiles * im.tr_chunks * 2 * kGemmHalf;
  PackEntry& e = o.pack.e[o.pack.n++];
  e.src = (long long)src; e.ld = ld; e.rows_o = rows_o; e.cols_i = cols_i; e.fwd_off = (long long)im.fwd; e.tr_off = (long long)im.tr;
  e.first_block = o.pack.total_blocks; e.pad = 0;
  o.pack.total_blocks += fwd_tiles * im.fwd_chunks + tr_tiles * im.tr_chunks;
  im.ok = true;
  return im;
}
void add_mlp_images(Offsets& o, MlpOff& m, int F, bool with_out) {
  m.i1 = add_image(o, m.W1, m.k_in, F, m.k_in);
  m.i2 = add_image(o, m.W2, F, F, F);
  if (with_out) m.i3 = add_image(o, m.W3, F, m.n_out, F);
}

MlpOff take_mlp(size_t& off, int k_in, int F, int n_out) {
  MlpOff m{};
  m.k_in = k_in; m.n_out = n_out;
  m.W1 = off; off += (size_t)F * k_in; m.b1 = off; off += F; m.g1 = off; off += F; m.be1 = off; off += F;
  m.W2 = off; off += (size_t)F * F; m.b2 = off; off += F; m.g2 = off; off += F; m.be2 = off; off += F;
  m.W3 = off; off += (size_t)n_out * F; m.b3 = off; off += n_out;
  return m;
}
Offsets make_offsets(const tib_model_desc& d, int n_temp) {
  Offsets o{};
  const int F = d.n_features;
  size_t off = 0;
  o.edge_emb = off; off += (size_t)d.n_edge_types * F;
  o.atom_emb = off; off += (size_t)d.n_types * F;
  o.combine = take_mlp(off, (2 + n_temp) * F, F, F);
  for (int l = 0; l < d.n_layers; ++l) {
    LayerOff L{};
    L.phi = take_mlp(off, 2 * F, F, 5 * F);
    L.w = take_mlp(off, F, F, 5 * F);
    L.UV = off; off += 2 * (size_t)F * F;
    L.upd = take_mlp(off, 2 * F, F, 3 * F);
    o.layers.push_back(L);
  }
  o.readout = take_mlp(off, F, F, 2);
  o.Vout = off; off += F;
  o.total = off;
  // the weight images every forward / data-gradient GEMM reads as its B operand (the readout's 2-row output layer and the
  // 1-row equivariant readout are not GEMMs)
  add_mlp_images(o, o.combine, F, true);
  for (LayerOff& L : o.layers) {
    add_mlp_images(o, L.phi, F, true);
    add_mlp_images(o, L.w, F, true);
    L.iuv = add_image(o, L.UV, F, 2 * F, F);
    add_mlp_images(o, L.upd, F, true);
  }
  add_mlp_images(o, o.readout, F, false);
  return o;
}

// ---- workspace ------------------------------------------------------------------------------------------------------------------
struct MlpAct { float *n1, *r1, *h1, *n2, *r2, *h2, *out; long long rows; };
struct LayerAct {
  MlpAct phi, w, upd;
  float *s_in, *v_in, *e_in, *s_mid, *v_mid, *uvvv, *q;
};
struct Ws {
  // graph
  int *src, *dst, *pair, *etype, *in_ptr;
  float4* dir;
  float *pair_dist, *pe;
  // inputs / loss
  float *xt, *tgt, *colsum, *gate;
  float* amax; int n_amax;
  // embedding
  float *X0, *s0; MlpAct emb;
  std::vector<LayerAct> layers;
  float *s_last, *v_last, *e_spare;
  MlpAct ro;
  // backward
  unsigned char* img;      // pre-packed weight images (B operands)
  float *ds, *dv, *de, *dv_src, *d_phi3, *d_w3, *d_gac, *d_uvvv, *dq, *dA, *dB, *dAw, *dBw, *ds0, *dX0;
  size_t bytes;
};

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t layout(Ws& w, char* base, const tib_model_desc& d, int n_temp, long long N, long long E, size_t img_bytes) {
  const int F = d.n_features, L = d.n_layers;
  const long long N2 = 2 * N, E2 = 2 * E, P2 = E;
  size_t off = 0;
  auto take = [&](size_t nbytes) { char* r = base ? base + off : nullptr; off += al256(nbytes); return r; };
  auto tf = [&](long long n) { return (float*)take(sizeof(float) * (size_t)n); };
  auto ti = [&](long long n) { return (int*)take(sizeof(int) * (size_t)n); };
  auto mlp = [&](MlpAct& a, long long rows, int n_out) {
    a.rows = rows;
    a.n1 = tf(rows * F); a.r1 = tf(rows); a.h1 = tf(rows * F);
    a.n2 = tf(rows * F); a.r2 = tf(rows); a.h2 = tf(rows * F);
    a.out = n_out > 0 ? tf(rows * n_out) : nullptr;
  };
  w.img = (unsigned char*)take(img_bytes);
  w.src = ti(E2); w.dst = ti(E2); w.pair = ti(E2); w.etype = ti(E2); w.in_ptr = ti(N2 + 1);
  w.dir = (float4*)take(sizeof(float4) * (size_t)E2);
  w.pair_dist = tf(P2); w.pe = tf(P2 * F);
  w.xt = tf(N2 * 3); w.tgt = tf(N2 * 3); w.colsum = tf(8); w.gate = tf(N2);
  w.n_amax = 16 * (L + 3); w.amax = tf(w.n_amax);
  w.X0 = tf(N * (2 + n_temp) * F); mlp(w.emb, N, F); w.s0 = w.emb.out;
  w.layers.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerAct& a = w.layers[l];
    a.s_in = tf(N2 * F); a.v_in = tf(N2 * 3 * F); a.e_in = tf(E2 * F);
    mlp(a.w, P2, 5 * F); mlp(a.phi, E2, 5 * F);
    a.s_mid = tf(N2 * F); a.v_mid = tf(N2 * 3 * F);
    a.uvvv = tf(N2 * 3 * 2 * F); a.q = tf(N2 * F);
    mlp(a.upd, N2, 3 * F);
  }
  w.s_last = tf(N2 * F); w.v_last = tf(N2 * 3 * F); w.e_spare = tf(E2 * F);
  mlp(w.ro, N2, 0);
  w.ds = tf(N2 * F); w.dv = tf(N2 * 3 * F); w.de = tf(E2 * F); w.dv_src = tf(N2 * 3 * F);
  w.d_phi3 = tf(E2 * 5 * F); w.d_w3 = tf(P2 * 5 * F); w.d_gac = tf(N2 * 3 * F); w.d_uvvv = tf(N2 * 3 * 2 * F); w.dq = tf(N2 * F);
  w.dA = tf(E2 * F); w.dB = tf(E2 * F); w.dAw = tf(P2 * F); w.dBw = tf(P2 * F); w.ds0 = tf(N * F); w.dX0 = tf(N * F);
  w.bytes = off;
  return off;
}

// ---- launch helpers ---------------------------------------------------------------------------------------------------------------
bool g_gemm_attr[64] = {};     // opt-in shared-memory size set, per device

thread_local std::vector<cudaEvent_t> g_events;     // reused by every call of this thread (no timing)
thread_local cudaStream_t g_side[64] = {};          // the side stream of each device

// Launch context.  `st` is the stream the helpers launch on: the caller's stream, or - between to_side() and to_main() -
// the library's side stream, where everything that is off the critical path of the backward pass runs (weight-gradient
// GEMMs, bias column sums, the whole w MLP): at 256 molecules most kernels of the main chain fill a fraction of the GPU.
// Buffers the side stream still reads are tracked, and the main stream waits before it overwrites one.
struct Ctx {
  cudaStream_t st, main, side;
  int depth;         // nesting of to_side()
  int n_sms;
  int* err;          // device error word (bounded mbarrier waits)
  float* amax; int n_amax, next_amax;
  const unsigned char* img;      // pre-packed weight images, or nullptr (every B operand built on the fly)
  size_t next_ev;
  struct Pending { const void* ptr; cudaEvent_t ev; } pend[16];
  int n_pend;
  float* new_amax() { return next_amax < n_amax ? amax + next_amax++ : nullptr; }
  bool two_streams() const { return side != main; }
  cudaEvent_t new_event() {
    if (next_ev == g_events.size()) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      g_events.push_back(e);
    }
    return g_events[next_ev++];
  }
  // order `to` after everything enqueued on `from` so far
  int order(cudaStream_t from, cudaStream_t to) {
    cudaEvent_t e = new_event();
    if (!e) return fail("cudaEventCreate failed");
    CUDA_TRY(cudaEventRecord(e, from));
    CUDA_TRY(cudaStreamWaitEvent(to, e, 0));
    return 0;
  }
  int to_side() {
    if (!two_streams()) return 0;
    if (depth++ == 0) { TRY(order(main, side)); st = side; }
    return 0;
  }
  void to_main() {
    if (!two_streams()) return;
    if (--depth == 0) st = main;
  }
  // the side stream has enqueued its last read of `ptr`
  int side_done(const void* ptr) {
    if (!two_streams()) return 0;
    cudaEvent_t e = new_event();
    if (!e) return fail("cudaEventCreate failed");
    CUDA_TRY(cudaEventRecord(e, side));
    for (int i = 0; i < n_pend; ++i)
      if (pend[i].ptr == ptr) { pend[i].ev = e; return 0; }
    if (n_pend == 16) return fail("internal: too many buffers pending on the side stream");
    pend[n_pend++] = {ptr, e};
    return 0;
  }
  // the main stream is about to overwrite `ptr`
  int before_write(const void* ptr) {
    if (!two_streams() || st != main) return 0;
    for (int i = 0; i < n_pend; ++i)
      if (pend[i].ptr == ptr) {
        CUDA_TRY(cudaStreamWaitEvent(main, pend[i].ev, 0));
        pend[i] = pend[--n_pend];
        return 0;
      }
    return 0;
  }
  int join() { return two_streams() ? order(side, main) : 0; }
};

int make_ctx(Ctx& c, cudaStream_t st, int* err, float* amax, int n_amax) {
  c = Ctx{};
  c.st = c.main = c.side = st; c.err = err; c.amax = amax; c.n_amax = n_amax;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&c.n_sms, cudaDevAttrMultiProcessorCount, dev));
  static const bool one_stream = getenv("TIB_TRAIN_NO_SIDE_STREAM") != nullptr || getenv("TIB_TRAIN_TRACE") != nullptr;
  if (!one_stream && dev >= 0 && dev < 64) {
    if (!g_side[dev]) CUDA_TRY(cudaStreamCreateWithFlags(&g_side[dev], cudaStreamNonBlocking));
    c.side = g_side[dev];
  }
  return 0;
}

GemmOperand op(const float* ptr, long long ld, int trans, float scale, const float* amax = nullptr, const int* idx = nullptr) {
  GemmOperand o{};
  o.ptr = ptr; o.ld = ld; o.idx = idx; o.trans = trans; o.amax = amax; o.scale = scale;
  return o;
}

// a pre-packed B operand: image base, first K chunk of this GEMM, K chunks per n tile of the image
struct BImg { const unsigned char* p; int chunk0, chunks; };

// second A segment and fused LayerNorm epilogue of a GEMM (train_gemm.cuh)
struct GemmExtra {
  const GemmOperand* A2 = nullptr; int k1 = 0;          // cat[A (k1 columns), A2] as the A operand
  int epi = EPI_NONE;
  const float *gamma = nullptr, *beta = nullptr;
  float *rstd = nullptr, *h = nullptr;
};

int gemm(Ctx& c, int M, int N, int K, const GemmOperand& A, const GemmOperand& B, float* C, long long ldc, int mode,
         const float* bias = nullptr, const int* c_idx = nullptr, bool split_k = false, BImg bimg = BImg{nullptr, 0, 0},
         const GemmExtra* ex = nullptr) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GemmP p{};
  p.b_img = bimg.p; p.b_chunk0 = bimg.chunk0; p.b_chunks = bimg.chunks;
  p.a_split = 1 << 30;
  if (ex) {
    if (ex->A2) {
      if (ex->k1 % kGemmKC || A.trans || ex->A2->trans || A.scale != ex->A2->scale || A.amax || ex->A2->amax)
        return fail("internal: a two-segment A operand needs row-major segments with one constant scale");
      p.A2 = *ex->A2; p.a_split = ex->k1 / kGemmKC;
    }
    if (ex->epi != EPI_NONE && (N > 128 || split_k || mode != GEMM_STORE || c_idx))
      return fail("internal: the fused LayerNorm epilogue needs whole rows in one tile (N <= 128), no K split, GEMM_STORE");
    p.epi = ex->epi; p.ln_gamma = ex->gamma; p.ln_beta = ex->beta; p.ln_rstd = ex->rstd; p.ln_h = ex->h;
  }
  p.M = M; p.N = N; p.K = K; p.A = A; p.B = B; p.C = C; p.ldc = ldc; p.c_idx = c_idx; p.bias = bias; p.mode = mode;
  p.alpha = 1.0f; p.passes = 3; p.err = c.err; p.dbg = g_gemm_dbg;
  const int mt = (M + 127) / 128, nt = (N + 127) / 128, chunks = (K + kGemmKC - 1) / kGemmKC;
  int splits = 1;
  if (split_k) {
    splits = std::max(1, std::min(chunks / 4 + 1, (2 * c.n_sms + mt * nt - 1) / (mt * nt)));
    if (mode != GEMM_ATOMIC) return fail("internal: split-K GEMM must accumulate atomically");
  }
  p.k_splits = splits;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !g_gemm_attr[dev]) {
    CUDA_TRY(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    CUDA_TRY(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    g_gemm_attr[dev] = true;
  }
  g_gemm_flops += 2.0 * M * (double)N * K;
  static const bool trace = getenv("TIB_TRAIN_TRACE") != nullptr;      // diagnostics: synchronous per-launch timing on stderr
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (trace) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaStreamSynchronize(c.st); cudaEventRecord(e0, c.st); }
  {
    Prof pf(TIB_K_TRAIN_GEMM, c.st);
    launch_pdl(k_gemm_tc, dim3(nt, mt, splits), kGemmThreads, kGemmSmem, c.st, p);
    LAUNCH_CHECK();
  }
  if (trace) {
    float ms = 0.0f;
    cudaEventRecord(e1, c.st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[gemm] M=%d N=%d K=%d tA=%d tB=%d idxA=%d idxB=%d mode=%d splits=%d ctas=%d  %.1f us  %.1f TFLOP/s\n", M, N, K, A.trans,
            B.trans, A.idx != nullptr, B.idx != nullptr, mode, splits, nt * mt * splits, ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return 0;
}

// forward GEMM over the input columns [k0, k0 + width) of W: chunks k0 / 32 .. of the forward image
BImg img_fwd(const Ctx& c, const ImgOff& im, int k0) {
  if (!c.img || !im.ok) return BImg{nullptr, 0, 0};
  return BImg{c.img + im.fwd, k0 / kGemmKC, im.fwd_chunks};
}
// data-gradient GEMM producing the input columns [k0, k0 + width): n tiles k0 / 128 .. of the transposed image
BImg img_tr(const Ctx& c, const ImgOff& im, int k0) {
  if (!c.img || !im.ok || k0 % 128 != 0) return BImg{nullptr, 0, 0};
  return BImg{c.img + im.tr + (size_t)(k0 / 128) * im.tr_chunks * 2 * kGemmHalf, 0, im.tr_chunks};
}

int blocks_for(long long n, int per_block) { return (int)std::min<long long>((n + per_block - 1) / per_block, 1 << 20); }

// one input segment of an MLP's first Linear: X[:, seg] = rows of `ptr` (optionally gathered), F_seg columns
struct Seg { const float* ptr; long long ld; const int* idx; int width; float scale; };

// Linear -> LN -> SiLU -> Linear -> LN -> SiLU [-> Linear]   (embedding.py:26-34)
int mlp_forward(Ctx& c, const float* W, const MlpOff& m, int F, const Seg* segs, int n_seg, MlpAct& a, bool with_out) {
  const int R = (int)a.rows;
  g_trace_rows = R;
  static const bool no_fuse = getenv("TIB_TRAIN_NO_LN_FUSION") != nullptr;      // diagnostics: separate LayerNorm kernels
  if (F <= 128 && n_seg <= 2 && !no_fuse) {
    // whole rows fit one GEMM tile: LayerNorm + SiLU run in the GEMM epilogue, the input segments are one A operand
    GemmExtra ex{};
    GemmOperand a2{};
    int K = segs[0].width;
    if (n_seg == 2) {
      a2 = op(segs[1].ptr, segs[1].ld, 0, segs[1].scale, nullptr, segs[1].idx);
      ex.A2 = &a2; ex.k1 = segs[0].width; K += segs[1].width;
    }
    ex.epi = EPI_LN_FWD; ex.gamma = W + m.g1; ex.beta = W + m.be1; ex.rstd = a.r1; ex.h = a.h1;
    TRY(gemm(c, R, F, K, op(segs[0].ptr, segs[0].ld, 0, segs[0].scale, nullptr, segs[0].idx), op(W + m.W1, m.k_in, 0, 1.0f), a.n1, F,
             GEMM_STORE, W + m.b1, nullptr, false, img_fwd(c, m.i1, 0), &ex));
    GemmExtra ex2{};
    ex2.epi = EPI_LN_FWD; ex2.gamma = W + m.g2; ex2.beta = W + m.be2; ex2.rstd = a.r2; ex2.h = a.h2;
    TRY(gemm(c, R, F, F, op(a.h1, F, 0, 1.0f), op(W + m.W2, F, 0, 1.0f), a.n2, F, GEMM_STORE, W + m.b2, nullptr, false, img_fwd(c, m.i2, 0),
             &ex2));
  } else {
    int k0 = 0;
    for (int s = 0; s < n_seg; ++s) {
      TRY(gemm(c, R, F, segs[s].width, op(segs[s].ptr, segs[s].ld, 0, segs[s].scale, nullptr, segs[s].idx),
               op(W + m.W1 + k0, m.k_in, 0, 1.0f), a.n1, F, s == 0 ? GEMM_STORE : GEMM_ACCUM, s == 0 ? W + m.b1 : nullptr, nullptr, false,
               img_fwd(c, m.i1, k0)));
      k0 += segs[s].width;
    }
    const int ln_blocks = std::min(blocks_for(R, 4), c.n_sms * 16);
    TRAIN_LAUNCH(c.st, launch_pdl(k_tr_ln_silu_fwd, ln_blocks, kEW, 0, c.st, R, F, a.n1, a.r1, a.h1, W + m.g1, W + m.be1));
    TRY(gemm(c, R, F, F, op(a.h1, F, 0, 1.0f), op(W + m.W2, F, 0, 1.0f), a.n2, F, GEMM_STORE, W + m.b2, nullptr, false, img_fwd(c, m.i2, 0)));
    TRAIN_LAUNCH(c.st, launch_pdl(k_tr_ln_silu_fwd, ln_blocks, kEW, 0, c.st, R, F, a.n2, a.r2, a.h2, W + m.g2, W + m.be2));
  }
  if (with_out)
    TRY(gemm(c, R, m.n_out, F, op(a.h2, F, 0, 1.0f), op(W + m.W3, F, 0, 1.0f), a.out, m.n_out, GEMM_STORE, W + m.b3, nullptr, false,
             img_fwd(c, m.i3, 0)));
  return 0;
}

// where the gradient of one input segment goes
struct SegGrad { float* ptr; long long ld; int mode; const int* c_idx; };   // ptr == nullptr: not needed

// Adjoint of mlp_forward.  dY [R][n_out] with its |max| slot (dY == nullptr: start from dh2 already in `dA`, the readout case).
int mlp_backward(Ctx& c, const float* W, float* G, const MlpOff& m, int F, const Seg* segs, const SegGrad* sg, int n_seg,
                 const MlpAct& a, const float* dY, const float* amax_dY, float* dA, float* dB) {
  const int R = (int)a.rows;
  g_trace_rows = R;
  constexpr int kLnBwdThreads = 256;
  const int ln_blocks = std::min(blocks_for(R, kLnBwdThreads / 32), 4 * c.n_sms);
  auto ln_bwd = F <= 128 ? k_tr_ln_silu_bwd<4> : k_tr_ln_silu_bwd<8>;      // <= 64 registers: 32 warps per SM
  const size_t ln_smem = 3 * (size_t)F * sizeof(float);
  if (dY) {
    TRY(c.to_side());
    TRY(gemm(c, m.n_out, F, R, op(dY, m.n_out, 1, 1.0f, amax_dY), op(a.h2, F, 1, 1.0f), G + m.W3, F, GEMM_ATOMIC, nullptr, nullptr, true));
    TRAIN_LAUNCH(c.st, launch_pdl(k_tr_colsum, dim3((m.n_out + kEW - 1) / kEW, std::min(blocks_for(R, 32), 512)), kEW, 0, c.st, R, m.n_out, dY, G + m.b3, nullptr));
    TRY(c.side_done(dY));
    c.to_main();
    TRY(c.before_write(dA));
    TRY(gemm(c, R, F, m.n_out, op(dY, m.n_out, 0, 1.0f, amax_dY), op(W + m.W3, F, 1, 1.0f), dA, F, GEMM_STORE, nullptr, nullptr, false,
             img_tr(c, m.i3, 0)));
  }
  float* am2 = c.new_amax();
  TRAIN_LAUNCH(c.st, launch_pdl(ln_bwd, ln_blocks, kLnBwdThreads, ln_smem, c.st, R, F, dA, a.n2, a.r2, W + m.g2, W + m.be2, G + m.g2, G + m.be2, G + m.b2, am2));
  TRY(c.to_side());
  TRY(gemm(c, F, F, R, op(dA, F, 1, 1.0f, am2), op(a.h1, F, 1, 1.0f), G + m.W2, F, GEMM_ATOMIC, nullptr, nullptr, true));
  TRY(c.side_done(dA));
  c.to_main();
  TRY(c.before_write(dB));
  TRY(gemm(c, R, F, F, op(dA, F, 0, 1.0f, am2), op(W + m.W2, F, 1, 1.0f), dB, F, GEMM_STORE, nullptr, nullptr, false, img_tr(c, m.i2, 0)));
  float* am1 = c.new_amax();
  TRAIN_LAUNCH(c.st, launch_pdl(ln_bwd, ln_blocks, kLnBwdThreads, ln_smem, c.st, R, F, dB, a.n1, a.r1, W + m.g1, W + m.be1, G + m.g1, G + m.be1, G + m.b1, am1));
  TRY(c.to_side());
  int k0 = 0;
  for (int s = 0; s < n_seg; ++s) {
    TRY(gemm(c, F, segs[s].width, R, op(dB, F, 1, 1.0f, am1), op(segs[s].ptr, segs[s].ld, 1, segs[s].scale, nullptr, segs[s].idx),
             G + m.W1 + k0, m.k_in, GEMM_ATOMIC, nullptr, nullptr, true));
    k0 += segs[s].width;
  }
  TRY(c.side_done(dB));
  c.to_main();
  k0 = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (sg && sg[s].ptr)
      TRY(gemm(c, R, segs[s].width, F, op(dB, F, 0, 1.0f, am1), op(W + m.W1 + k0, m.k_in, 1, 1.0f), sg[s].ptr, sg[s].ld, sg[s].mode,
               nullptr, sg[s].c_idx, false, img_tr(c, m.i1, k0)));
    k0 += segs[s].width;
  }
  return 0;
}

thread_local int* g_dev_err = nullptr;     // one device error word per host thread (never freed: 4 bytes)
thread_local int g_dev_err_device = -1;

}  // namespace

extern "C" {

size_t tib_train_workspace_bytes(const tib_model_desc* desc, int32_t n_mol, int32_t n_nodes, int64_t n_edges) {
  (void)n_mol;
  if (!desc) return 0;
  Ws w{};
  const int n_temp = desc->variant == TIB_VARIANT_AMBIENT ? 2 : (desc->variant == TIB_VARIANT_LATENT_MULTI_T ? 1 : 0);
  return layout(w, nullptr, *desc, n_temp, n_nodes, n_edges, make_offsets(*desc, n_temp).img_bytes);
}

int tib_train_loss_grad(const tib_model_desc* desc, const float* weights, const tib_train_batch* b, const tib_interpolant* ip,
                        double* loss, float* grad, float* out_b, void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc || !weights || !b || !ip || !loss || !grad || !workspace) return fail("tib_train_loss_grad: null argument");
  if (desc->abi_version != TIB_ABI_VERSION) return fail("tib_train_loss_grad: ABI version %d, expected %d", desc->abi_version, TIB_ABI_VERSION);
  if (desc->variant != TIB_VARIANT_AMBIENT) return fail("tib_train_loss_grad: only the ambient variant trains (train_ambient.py)");
  const int F = desc->n_features, L = desc->n_layers;
  if (F % 32 != 0 || F < 32 || F > 256) return fail("tib_train_loss_grad: n_features must be a multiple of 32 in [32, 256]");
  if (b->n_mol < 1 || b->n_nodes < 2 || b->n_edges < 2) return fail("tib_train_loss_grad: empty batch");
  if (2 * b->n_edges >= (1ll << 31) || 6ll * b->n_nodes >= (1ll << 31)) return fail("tib_train_loss_grad: batch exceeds int32 row indices");
  if (ip->gamma_kind != TIB_GAMMA_BROWNIAN && ip->gamma_kind != TIB_GAMMA_SIN2) return fail("tib_train_loss_grad: unknown gamma");
  const int n_temp = 2;
  const long long N = b->n_nodes, E = b->n_edges, N2 = 2 * N, E2 = 2 * E, P2 = E;
  Ws w{};
  const Offsets o = make_offsets(*desc, n_temp);
  if (layout(w, (char*)workspace, *desc, n_temp, N, E, o.img_bytes) > workspace_bytes)
    return fail("tib_train_loss_grad: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!g_dev_err || g_dev_err_device != dev) {
    CUDA_TRY(cudaMalloc(&g_dev_err, sizeof(int)));
    CUDA_TRY(cudaMemset(g_dev_err, 0, sizeof(int)));
    g_dev_err_device = dev;
  }
  Ctx c{};
  TRY(make_ctx(c, st, g_dev_err, w.amax, w.n_amax));
  const float* W = weights;
  float* G = grad;

  CUDA_TRY(cudaMemsetAsync(G, 0, o.total * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(w.amax, 0, w.n_amax * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(w.colsum, 0, 8 * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(double), st));

  // ---- weight images: every Linear's W as a ready-made B operand for the forward and the data-gradient GEMMs (once per call:
  // the weights change with every optimiser step)
  static const bool no_pack = getenv("TIB_TRAIN_NO_PACK") != nullptr;       // diagnostics: build every B operand on the fly
  if (!no_pack && o.pack.n > 0) {
    TRAIN_LAUNCH(st, launch_pdl(k_pack_operand, o.pack.total_blocks, kGemmThreads, 0, st, o.pack, W, w.img));
    c.img = w.img;
  }

  // ---- interpolant, targets, graph (interpolants.py:16-33; losses.py:52-57; graph.py:27-29) -----------------------------------------
  TRAIN_LAUNCH(st, launch_pdl(k_tr_interp, std::min(blocks_for(N, kEW), c.n_sms * 4), kEW, 0, st, (int)N, b->x0, b->x1, b->t, b->z, ip->gamma_kind, ip->a, w.xt, w.tgt, w.colsum));
  TRAIN_LAUNCH(st, launch_pdl(k_tr_center, blocks_for(6 * N, kEW), kEW, 0, st, (int)N, w.xt, w.colsum));
  GraphP gp{b->n_mol, (int)N, E, b->mol_ptr, (const long long*)b->edge_ptr, b->edge_type, w.xt, w.src, w.dst, w.pair, w.etype, w.in_ptr,
            w.dir, w.pair_dist};
  TRAIN_LAUNCH(st, launch_pdl(k_tr_graph, dim3(b->n_mol, 2), kEW, 0, st, gp));

  // ---- embeddings (embedding.py:68-86,249-261; cpainn.py:70-71): x-independent, shared by both passes ------------------------------
  TRAIN_LAUNCH(st, launch_pdl(k_tr_embed_in, (int)N, kEW, 0, st, (int)N, F, n_temp, b->atom_id, b->temp0, b->temp1, b->t, W + o.atom_emb, desc->temp_mean,
                                        desc->temp_range, desc->temp_length, desc->time_length, w.X0));
  const Seg seg_emb[1] = {{w.X0, (2 + n_temp) * F, nullptr, (2 + n_temp) * F, 1.0f}};
  TRY(mlp_forward(c, W, o.combine, F, seg_emb, 1, w.emb, true));
  LayerAct& A0 = w.layers[0];
  TRAIN_LAUNCH(st, launch_pdl(k_tr_gather_rows, blocks_for(N2 * F, kEW), kEW, 0, st, N2, F, nullptr, (int)N, w.s0, A0.s_in));
  TRAIN_LAUNCH(st, launch_pdl(k_tr_gather_rows, blocks_for(E2 * F, kEW), kEW, 0, st, E2, F, w.etype, 0, W + o.edge_emb, A0.e_in));
  CUDA_TRY(cudaMemsetAsync(A0.v_in, 0, sizeof(float) * N2 * 3 * F, st));
  TRAIN_LAUNCH(st, launch_pdl(k_tr_pair_pe, blocks_for(P2 * (F / 2), kEW), kEW, 0, st, P2, F, w.pair_dist, desc->length_scale, w.pe));

  // ---- forward through the layers (cpainn.py:138-150) ----------------------------------------------------------------------------------
  const int node_blocks = (int)std::min<long long>(N2, c.n_sms * 16);
  for (int l = 0; l < L; ++l) {
    LayerAct& a = w.layers[l];
    const LayerOff& lo = o.layers[l];
    float* s_next = l + 1 < L ? w.layers[l + 1].s_in : w.s_last;
    float* v_next = l + 1 < L ? w.layers[l + 1].v_in : w.v_last;
    float* e_next = l + 1 < L ? w.layers[l + 1].e_in : w.e_spare;
    // the w MLP needs only the pair distances: it runs beside the phi MLP
    const Seg seg_w[1] = {{w.pe, F, nullptr, F, 1.0f}};
    TRY(c.to_side());
    TRY(mlp_forward(c, W, lo.w, F, seg_w, 1, a.w, true));
    c.to_main();
    const Seg seg_phi[2] = {{a.s_in, F, w.src, F, kStateScale}, {a.e_in, F, nullptr, F, kStateScale}};
    TRY(mlp_forward(c, W, lo.phi, F, seg_phi, 2, a.phi, true));
    TRY(c.join());
    CombineP cp{(int)N2, F, w.in_ptr, w.src, w.pair, w.dir, a.phi.out, a.w.out, a.s_in, a.v_in, a.e_in, a.s_mid, a.v_mid, e_next};
    TRAIN_LAUNCH(st, launch_pdl(k_tr_combine_fwd, node_blocks, kEW, 0, st, cp));
    TRY(gemm(c, (int)(3 * N2), 2 * F, F, op(a.v_mid, F, 0, kStateScale), op(W + lo.UV, F, 0, 1.0f), a.uvvv, 2 * F, GEMM_STORE, nullptr, nullptr,
             false, img_fwd(c, lo.iuv, 0)));
    TRAIN_LAUNCH(st, launch_pdl(k_tr_upd_q, blocks_for(N2 * F, kEW), kEW, 0, st, N2 * F, F, a.uvvv, a.q));
    const Seg seg_upd[2] = {{a.q, F, nullptr, F, kStateScale}, {a.s_mid, F, nullptr, F, kStateScale}};
    TRY(mlp_forward(c, W, lo.upd, F, seg_upd, 2, a.upd, true));
    TRAIN_LAUNCH(st, launch_pdl(k_tr_upd_apply, blocks_for(N2 * F, kEW), kEW, 0, st, N2 * F, F, a.uvvv, a.q, a.upd.out, a.s_mid, a.v_mid, s_next, v_next));
  }

  // ---- readout, loss and d loss / d b (cpainn.py:425-437; losses.py:126-133) -----------------------------------------------------------
  const Seg seg_ro[1] = {{w.s_last, F, nullptr, F, kStateScale}};
  TRY(mlp_forward(c, W, o.readout, F, seg_ro, 1, w.ro, false));
  float* am_ro = c.new_amax();
  float* bout = out_b ? out_b : w.dX0;      // dX0 is free until the very end ([N][F] >= [2N][3] for F >= 32)
  ReadoutP rp{(int)N2, F, (int)N, w.ro.h2, w.v_last, W + o.readout.W3, W + o.readout.b3, W + o.Vout, w.tgt, bout, w.gate, loss,
              w.dA, w.dv, G + o.readout.W3, G + o.readout.b3, G + o.Vout, am_ro};
  TRAIN_LAUNCH(st, launch_pdl(k_tr_readout, std::min(blocks_for(N2, 4), c.n_sms * 4), kEW, 2 * F * sizeof(float), st, rp));

  // ---- backward ------------------------------------------------------------------------------------------------------------------------------
  {
    const SegGrad sg[1] = {{w.ds, F, GEMM_STORE, nullptr}};
    // dh2 of the readout MLP is in dA; the W3 / b3 / Vout gradients were accumulated by k_tr_readout
    MlpAct ro = w.ro;
    TRY(mlp_backward(c, W, G, o.readout, F, seg_ro, sg, 1, ro, nullptr, am_ro, w.dA, w.dB));
  }
  CUDA_TRY(cudaMemsetAsync(w.de, 0, sizeof(float) * E2 * F, st));      // nothing reads the last layer's e_out
  for (int l = L - 1; l >= 0; --l) {
    LayerAct& a = w.layers[l];
    const LayerOff& lo = o.layers[l];
    // Update (cpainn.py:345-376)
    float* am_gac = c.new_amax();
    TRY(c.before_write(w.d_gac));
    TRY(c.before_write(w.d_uvvv));
    TRAIN_LAUNCH(st, launch_pdl(k_tr_upd_bwd1, blocks_for(N2 * F, kEW), kEW, 0, st, N2 * F, F, a.uvvv, a.q, a.upd.out, w.ds, w.dv, w.d_gac, w.d_uvvv, w.dq, am_gac));
    const Seg seg_upd[2] = {{a.q, F, nullptr, F, kStateScale}, {a.s_mid, F, nullptr, F, kStateScale}};
    const SegGrad sg_upd[2] = {{w.dq, F, GEMM_ACCUM, nullptr}, {w.ds, F, GEMM_ACCUM, nullptr}};
    TRY(mlp_backward(c, W, G, lo.upd, F, seg_upd, sg_upd, 2, a.upd, w.d_gac, am_gac, w.dA, w.dB));
    float* am_uv = c.new_amax();
    TRAIN_LAUNCH(st, launch_pdl(k_tr_upd_bwd2, blocks_for(N2 * F, kEW), kEW, 0, st, N2 * F, F, a.uvvv, a.q, w.dq, w.d_uvvv, am_uv));
    TRY(c.to_side());
    TRY(gemm(c, 2 * F, F, (int)(3 * N2), op(w.d_uvvv, 2 * F, 1, 1.0f, am_uv), op(a.v_mid, F, 1, kStateScale), G + lo.UV, F, GEMM_ATOMIC,
             nullptr, nullptr, true));
    TRY(c.side_done(w.d_uvvv));
    c.to_main();
    TRY(gemm(c, (int)(3 * N2), F, 2 * F, op(w.d_uvvv, 2 * F, 0, 1.0f, am_uv), op(W + lo.UV, F, 1, 1.0f), w.dv, F, GEMM_ACCUM, nullptr, nullptr,
             false, img_tr(c, lo.iuv, 0)));
    // SE3Message (cpainn.py:263-310)
    TRY(c.before_write(w.d_w3));
    TRY(c.before_write(w.d_phi3));
    CUDA_TRY(cudaMemsetAsync(w.d_w3, 0, sizeof(float) * P2 * 5 * F, st));
    CUDA_TRY(cudaMemsetAsync(w.dv_src, 0, sizeof(float) * N2 * 3 * F, st));
    float* am_phi = c.new_amax();
    float* am_w = c.new_amax();
    float* e_next = l + 1 < L ? w.layers[l + 1].e_in : w.e_spare;
    CombineBwdP cb{{(int)N2, F, w.in_ptr, w.src, w.pair, w.dir, a.phi.out, a.w.out, a.s_in, a.v_in, a.e_in, a.s_mid, a.v_mid, e_next},
                   w.ds, w.dv, w.de, w.d_phi3, w.d_w3, w.dv_src, am_phi, am_w};
    TRAIN_LAUNCH(st, launch_pdl(k_tr_combine_bwd, node_blocks, kEW, 0, st, cb));
    TRAIN_LAUNCH(st, launch_pdl(k_tr_add, blocks_for(N2 * 3 * F, kEW), kEW, 0, st, N2 * 3 * F, w.dv, w.dv_src));
    // the w MLP's adjoint produces weight gradients only: all of it runs on the side stream, with its own scratch
    const Seg seg_w[1] = {{w.pe, F, nullptr, F, 1.0f}};
    TRY(c.to_side());
    TRY(mlp_backward(c, W, G, lo.w, F, seg_w, nullptr, 1, a.w, w.d_w3, am_w, w.dAw, w.dBw));
    TRY(c.side_done(w.d_w3));
    c.to_main();
    const Seg seg_phi[2] = {{a.s_in, F, w.src, F, kStateScale}, {a.e_in, F, nullptr, F, kStateScale}};
    const SegGrad sg_phi[2] = {{w.ds, F, GEMM_ATOMIC, w.src}, {w.de, F, GEMM_ACCUM, nullptr}};
    TRY(mlp_backward(c, W, G, lo.phi, F, seg_phi, sg_phi, 2, a.phi, w.d_phi3, am_phi, w.dA, w.dB));
  }
  // embeddings: e0 = Emb4(edge_type), s0 = combine MLP (both passes share it), atom embedding
  {
    const int nb = std::min(blocks_for(E2, 32), c.n_sms * 8);
    TRAIN_LAUNCH(st, launch_pdl(k_tr_scatter_rows, nb, kEW, sizeof(float) * desc->n_edge_types * F, st, E2, F, desc->n_edge_types, w.etype, w.de, G + o.edge_emb));
    float* am_s0 = c.new_amax();
    TRAIN_LAUNCH(st, launch_pdl(k_tr_fold_passes, blocks_for(N * F, kEW), kEW, 0, st, N * F, w.ds, w.ds0, am_s0));
    // weight gradients over all input columns; the input gradient only for the first F (the atom embedding) - the
    // positional-encoding columns carry no parameters
    const int kin = (2 + n_temp) * F;
    const Seg seg2[2] = {{w.X0, kin, nullptr, F, 1.0f}, {w.X0 + F, kin, nullptr, kin - F, 1.0f}};
    const SegGrad sg2[2] = {{w.dX0, F, GEMM_STORE, nullptr}, {nullptr, 0, 0, nullptr}};
    TRY(mlp_backward(c, W, G, o.combine, F, seg2, sg2, 2, w.emb, w.ds0, am_s0, w.dA, w.dB));
    TRY(c.before_write(w.dX0));
    const int nb2 = std::min(blocks_for(N, 64), c.n_sms);
    TRAIN_LAUNCH(st, launch_pdl(k_tr_scatter_rows, nb2, 1024, sizeof(float) * desc->n_types * F, st, N, F, desc->n_types, b->atom_id, w.dX0, G + o.atom_emb));
  }
  TRY(c.join());          // the caller's stream owns the complete gradient again
  return 0;
}

/* Pipeline diagnostics of k_gemm_tc: enable = 1 makes every later launch of this thread record clock64() stamps of its CTA
 * (0,0,0) (start, TMEM allocated, then per K chunk: built, barrier passed; all issued, accumulator complete, epilogue done,
 * end); out (HOST [64], may be NULL) receives the stamps of the last launch, out[63] = their count.  Synchronises the device. */
int tib_gemm_debug(int enable, long long* out) {
  CUDA_TRY(cudaDeviceSynchronize());
  if (out && g_gemm_dbg) CUDA_TRY(cudaMemcpy(out, g_gemm_dbg, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  if (enable && !g_gemm_dbg) {
    CUDA_TRY(cudaMalloc(&g_gemm_dbg, 64 * sizeof(long long)));
    CUDA_TRY(cudaMemset(g_gemm_dbg, 0, 64 * sizeof(long long)));
  }
  if (!enable && g_gemm_dbg) { cudaFree(g_gemm_dbg); g_gemm_dbg = nullptr; }
  return 0;
}

double tib_train_gemm_flops(int reset) {
  const double v = g_gemm_flops;
  if (reset) g_gemm_flops = 0.0;
  return v;
}

int tib_train_status(void* stream) {
  if (!g_dev_err) return 0;
  int h = 0;
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  CUDA_TRY(cudaMemcpy(&h, g_dev_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) {
    CUDA_TRY(cudaMemset(g_dev_err, 0, sizeof(int)));
    return fail("tib_train: a tensor-core pipeline wait timed out (device error word %d)", h);
  }
  return 0;
}

int tib_adam_step(float* weights, const float* grad, float* m, float* v, size_t n, int32_t step, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float max_grad_norm, double* scratch, void* stream) {
  if (!weights || !grad || !m || !v || !scratch) return fail("tib_adam_step: null argument");
  if (step < 1) return fail("tib_adam_step: step counts from 1");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  int dev = 0, n_sms = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  const int nb = (int)std::min<size_t>((n + 255) / 256, (size_t)n_sms * 8);
  if (max_grad_norm > 0.0f) {
    TRAIN_LAUNCH(st, launch_pdl(k_tr_sqnorm, nb, 256, 0, st, (long long)n, grad, scratch));
  }
  const double bc1 = 1.0 - std::pow((double)beta1, step), bc2 = 1.0 - std::pow((double)beta2, step);
  TRAIN_LAUNCH(st, launch_pdl(k_tr_adam, nb, 256, 0, st, (long long)n, weights, grad, m, v, scratch, max_grad_norm, lr, beta1, beta2, eps, weight_decay,
                                (float)bc1, (float)std::sqrt(bc2)));
  return 0;
}

/* General split-f16 GEMM on tcgen05 (train_gemm.cuh) exposed for tests and benchmarks:
 *   C[M][N] (mode 0: =, 1: atomic +=, 2: +=)  sum_k A(m,k) B(n,k),   operand (ptr, ld, trans, idx) as in train_gemm.cuh. */
int tib_gemm_f16x3(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, int32_t trans_a, const int32_t* idx_a, float scale_a,
                   const float* amax_a, const float* B, int64_t ldb, int32_t trans_b, const int32_t* idx_b, float scale_b, float* C,
                   int64_t ldc, const int32_t* c_idx, const float* bias, int32_t mode, int32_t split_k, void* stream) {
  if (!A || !B || !C) return fail("tib_gemm_f16x3: null argument");
  if (!trans_a && (K % 8 || lda % 4)) return fail("tib_gemm_f16x3: a row-major A operand needs K %% 8 == 0 and lda %% 4 == 0");
  if (!trans_b && (K % 8 || ldb % 4)) return fail("tib_gemm_f16x3: a row-major B operand needs K %% 8 == 0 and ldb %% 4 == 0");
  if (N % 4 || ldc % 4 || ((uintptr_t)C & 15) || ((uintptr_t)bias & 15)) return fail("tib_gemm_f16x3: N, ldc must be multiples of 4 and C, bias 16-byte aligned");
  if ((!trans_a && ((uintptr_t)A & 15)) || (!trans_b && ((uintptr_t)B & 15))) return fail("tib_gemm_f16x3: row-major operands must be 16-byte aligned");
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!g_dev_err || g_dev_err_device != dev) {
    CUDA_TRY(cudaMalloc(&g_dev_err, sizeof(int)));
    CUDA_TRY(cudaMemset(g_dev_err, 0, sizeof(int)));
    g_dev_err_device = dev;
  }
  Ctx c{};
  TRY(make_ctx(c, (cudaStream_t)stream, g_dev_err, nullptr, 0));
  c.side = c.main;
  return gemm(c, M, N, K, op(A, lda, trans_a, scale_a, amax_a, idx_a), op(B, ldb, trans_b, scale_b, nullptr, idx_b), C, ldc, mode, bias,
              c_idx, split_k != 0);
}

}  // extern "C"
