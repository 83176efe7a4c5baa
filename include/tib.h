/*
 * tib.h - C ABI of libtib.so, the B200-native (sm_100a) sampling hot path (and training step) of
 * olsson-group/thermodynamic-interpolation.
 *
 * The reference has no FFI of its own: its seams are three Python call signatures
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it sits under
 * (paths relative to the reference repo).  All pointers are plain device or host pointers; no
 * torch types cross this boundary.  Conventions:
 *   - the caller owns every buffer; the library keeps nothing past a call except the
 *     tib_model handle (repacked weights) created by tib_model_create;
 *   - every call is asynchronous on the given CUDA stream (`stream` is a cudaStream_t passed
 *     as void*), except where stated (dopri5 reads one scalar back per attempted step,
 *     exactly like torchdiffeq's accept/reject);
 *   - return value 0 = ok, negative = error; the message is in tib_last_error() (thread local);
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef TIB_H_
#define TIB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIB_ABI_VERSION 6

/* ---- model ------------------------------------------------------------------------------- */

enum { TIB_VARIANT_AMBIENT = 0,       /* T0 and T1 encoders   (mdqm9/thermo/ambient/models/cpainn.py:67-90) */
       TIB_VARIANT_LATENT_MULTI_T = 1,/* one encoder on T     (mdqm9/thermo/latent/models/cpainn.py:43-58)  */
       TIB_VARIANT_LATENT_SINGLE_T = 2/* no temperature input (mdqm9/thermo/latent/models/cpainn.py:59-72)  */ };

/* GEMM arithmetic of the MLP / equivariant-linear layers. */
enum { TIB_MATH_FP32_SIMT = 0,        /* fp32 FMA on CUDA cores: rounds like the reference op by op          */
       TIB_MATH_F16X3_TC = 1,         /* tcgen05, operands split hi+lo f16 (22 bits), 3 MMAs, fp32 accumulate: */
                                      /*   fp32-faithful (~2^-21 per product); n_features = 128 (fused message */
                                      /*   / update / readout kernels) and 256 (layered MLP-chain kernels)     */
       TIB_MATH_F16_TC = 2,           /* tcgen05, single f16 pass (TF32-class, ~2^-11): opt-in only           */
       TIB_MATH_F16X3_LAYERED = 3     /* F16X3 arithmetic on the layered kernels also for n_features = 128     */
                                      /*   (cross-check of the path the divergence and F = 256 use)            */ };

typedef struct tib_model tib_model;

/* Constructor arguments of cPaiNN (ambient cpainn.py:23-32, latent cpainn.py:22-31) plus the
 * constants its sub-modules hard-code (PaiNNBase length_scale=10, cpainn.py:127; edge-type
 * vocabulary 4, cpainn.py:70; TemperatureEncoder mean/range, embedding.py:208-209). */
typedef struct {
  int32_t abi_version;     /* = TIB_ABI_VERSION */
  int32_t variant;         /* TIB_VARIANT_* */
  int32_t n_features;      /* F: 32, 64, 128 or 256 */
  int32_t n_layers;        /* score_layers L */
  int32_t n_types;         /* atom-id vocabulary (25) */
  int32_t n_edge_types;    /* 4 */
  float   temp_length;     /* PositionalEncoder max_length for temperatures */
  float   time_length;     /* ... for t */
  float   length_scale;    /* ... for edge distances */
  float   temp_mean;       /* mean(temperatures)          = 650 */
  float   temp_range;      /* max - min of temperatures   = 700 */
} tib_model_desc;

/* Number of floats tib_model_create expects in `packed_weights` for this descriptor.
 * Packing order ([out,in] row-major, exactly the reference's state_dict tensors):
 *   edge_emb[n_edge_types,F], atom_emb[n_types,F], MLP(combine: (2+n_temp)F -> F -> F),
 *   per layer: MLP(phi: 2F->F->5F), MLP(w: F->F->5F), U[F,F], V[F,F], MLP(update: 2F->F->3F),
 *   MLP(readout: F->F->2), Vout[1,F]
 * where MLP(in->h->out) = W1[h,in] b1[h] ln1_w[h] ln1_b[h] W2[h,h] b2[h] ln2_w[h] ln2_b[h] W3[out,h] b3[out]
 * (embedding.py:26-34). */
size_t tib_packed_weight_count(const tib_model_desc* desc);

/* Replaces cPaiNN(...).load_state_dict(sd).to(device)  (mdqm9/sample_ambient.py:125-131,72).
 * `packed_weights` is a HOST pointer; the weights are copied to `device` and repacked once. */
int  tib_model_create(tib_model** out, const tib_model_desc* desc, const float* packed_weights,
                      size_t n_floats, int device);
void tib_model_destroy(tib_model* m);
int  tib_model_set_math(tib_model* m, int math_mode);     /* TIB_MATH_*; default: F16X3_TC if n_features == 128, else FP32_SIMT */
/* Synchronises `stream` and reports (once) a device-side pipeline error recorded by the tensor-core
 * kernels' bounded barrier waits.  0 = healthy. */
int  tib_model_status(tib_model* m, void* stream);

/* ---- batch ------------------------------------------------------------------------------- */

/* The batch contract of MDQM9SamplerDataset.process + PyG collate
 * (mdqm9/data/mdqm9_ambient.py:160-170; latent mdqm9/data/mdqm9_latent.py:188-205), reduced to
 * what the drift reads.  The graph must be the complete digraph per molecule with edges in
 * (src,dst)-lexicographic order (what `coalesce` produces, mdqm9/thermo/utils.py:74-78); edge row
 * of (i -> j) in molecule m is  edge_ptr[m] + i*(n_m-1) + j - (j>i).  All pointers are DEVICE pointers. */
typedef struct {
  int32_t        n_mol;
  int32_t        n_nodes;      /* N = sum n_m */
  int64_t        n_edges;      /* E = sum n_m (n_m - 1) */
  int32_t        max_atoms;    /* max n_m (<= 64) */
  const int32_t* mol_ptr;      /* [n_mol+1] node offsets  (batch.ptr) */
  const int64_t* edge_ptr;     /* [n_mol+1] edge-row offsets */
  const int32_t* atom_id;      /* [N]  batch.atoms / batch.atom_number */
  const uint8_t* edge_type;    /* [E]  batch.edge_type in {0..3} */
  const float*   temp0;        /* [N]  batch.T0 (ambient) or batch.T (latent multi-T); may be NULL for single-T */
  const float*   temp1;        /* [N]  batch.T1 (ambient); NULL otherwise */
  /* Optional de-duplication of the x-independent node embedding s0 = MLP(cat[Emb(atom), PE(T0), PE(T1), PE(t)])
   * (embedding.py:68-86,249-261): nodes with the same (atom_id, temp0, temp1) share one row.  embed_index == NULL
   * (or n_embed_rows == 0) = evaluate every node. */
  int32_t        n_embed_rows; /* U distinct (atom_id, temp0, temp1) triples */
  const int32_t* embed_index;  /* [N]  row of each node in the U-row table */
  const int32_t* embed_atom_id;/* [U] */
  const float*   embed_temp0;  /* [U] or NULL */
  const float*   embed_temp1;  /* [U] or NULL */
  /* Optional tiling of the destination nodes for the tensor-core message kernel: tile t owns nodes
   * [tile_node_ptr[t], tile_node_ptr[t+1]) - at most 16 nodes whose incoming edges total at most 128 rows.
   * NULL (or n_tiles == 0) = uniform tiles of min(16, 128 / (max_atoms - 1)) nodes, which wastes rows when
   * molecule sizes are mixed. */
  int32_t        n_tiles;
  const int32_t* tile_node_ptr;/* [n_tiles + 1] */
} tib_batch;

/* Scratch the caller must provide to tib_drift / tib_rollout_* for this batch shape. */
size_t tib_workspace_bytes(const tib_model* m, int32_t n_mol, int32_t n_nodes, int64_t n_edges);

/* ---- the drift: cPaiNN.forward via ODEWrapper.forward --------------------------------------
 * Replaces  ODEWrapper(b).forward(t, x, batch)  with return_dlogp=False
 * (mdqm9/thermo/ambient/models/ode_wrapper.py:51-57 -> cpainn.py:93-115).
 *   x [N,3] fp32 (device), t scalar, out_b [N,3] fp32 (device). */
int tib_drift(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b,
              void* workspace, size_t workspace_bytes, void* stream);

/* ---- the drift with its exact divergence: ODEWrapper.forward with return_dlogp=True ---------
 * Replaces  b(batch).output  plus  ODEWrapper.compute_divergence(b, batch)
 * (mdqm9/thermo/ambient/models/ode_wrapper.py:39-49,59-91; latent/models/ode_wrapper.py:38-46,57-86):
 *   out_div[mol] = sum over atoms a and coordinates c of d b[a][c] / d x[a][c], UNSCALED (the ambient
 *   wrapper's x 1e-2 and the sign are applied by the caller), fp32 [n_mol] on the device.
 * The reference runs 3n reverse passes; this runs 3 * max_atoms forward-mode tangent directions and - unlike the
 * reference (ode_wrapper.py:75,79) - accepts molecules of different sizes.  With a tensor-core math mode
 * (n_features 128 / 256) the primal runs layer by layer on the MLP-chain kernels (csrc/tc_chain.cuh) keeping its
 * LayerNorm intermediates, and the tangents go through the same weights on tcgen05: one tangent for the w MLP (it
 * depends on x through the scalar distance), none for the first layer's phi MLP.  With TIB_MATH_FP32_SIMT (and for
 * n_features 32 / 64) dual-number fp32 kernels carry D directions per pass.  The workspace is larger than
 * tib_workspace_bytes: use tib_div_workspace_bytes (max_atoms = the batch's largest molecule). */
size_t tib_div_workspace_bytes(const tib_model* m, int32_t n_mol, int32_t n_nodes, int64_t n_edges, int32_t max_atoms);
int tib_drift_div(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b, float* out_div,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- post-processing: internal coordinates of sampled conformers --------------------------------
 * Replaces  z_matrix.construct_z_matrix_batch(X_batch, ref_atoms, placing_order)
 * (mdqm9/analysis/utils/z_matrix.py:56-102; distances / angles / atan2 torsions of mol_geometry.py:25-81), the
 * step that turns samples into the torsion features of the reference's analysis (results_00031.py:15-18,140-149).
 *   x [n_conf][n_atoms][3] fp32 (device), order [n_atoms] int32 = placing_order, ref [n_atoms][3] int32 = the
 *   reference triplets (distance, angle, torsion partner; rows 0..2 are only partly used, as in the reference),
 *   z [n_conf][n_atoms-1][3] fp32 = (distance, angle, torsion), zero where the reference leaves zeros. */
int tib_zmatrix(const float* x, int64_t n_conf, int32_t n_atoms, const int32_t* order, const int32_t* ref, float* z,
                void* stream);

/* Torsion features -> TICA projection -> histograms: replaces, per conformer, enc() + TICA.transform + plt.hist of the
 * reference's figure script (mdqm9/plots/10506_main.ipynb cells 3-4; deeptime.decomposition.TICA is an un-vendored pin):
 *   feat = (cos t_0, sin t_0, cos t_1, ...), proj = (feat - mean) R [:, :dim], hist[k] = weighted histogram of proj[:, k]
 *   with n_bins equal bins on [lo, hi] (the right edge in the last bin, as numpy.histogram).
 * torsions: DEVICE fp32, the torsion of feature j of conformer c at torsions[c * stride + col0 + j * col_step] (so the
 * z-matrix of tib_zmatrix can be read in place: stride = 3 (n_atoms - 1), col0 = 3 * 2 + 2, col_step = 3, n_tors = n_atoms - 3);
 * mean [2 n_tors], R [2 n_tors][dim] DEVICE fp64 (row-major), dim <= 4; weight [n_conf] DEVICE fp64 or NULL;
 * proj [n_conf][dim] DEVICE fp32 or NULL; hist [dim][n_bins] DEVICE fp64 (ACCUMULATED into; zero it first) or NULL. */
int tib_tica_project(const float* torsions, int64_t n_conf, int32_t n_tors, int64_t stride, int32_t col0, int32_t col_step,
                     const double* mean, const double* R, int32_t dim, const double* weight, float* proj, double* hist,
                     int32_t n_bins, double lo, double hi, void* stream);

/* ---- K1: fused integrator state updates ---------------------------------------------------
 * One explicit Euler / Euler-Maruyama update over a flat state of n floats:
 *     x_out = x + dt*b                       [+ dt*eps*score] [+ sqrt(2*eps*dt)*noise]
 * with the product and the sum rounded separately (torchdiffeq fixed-grid `y0 + dt*f0`,
 * solvers.py FixedGridODESolver.integrate / Euler._step_func).  `score`, `noise` may be NULL
 * (then the result is bit-identical to the Euler path); `frame` (may be NULL) receives a copy of
 * x_out - the saved trajectory frame.  x_out may alias x. */
int tib_step_euler(const float* x, const float* b, const float* score, const float* noise,
                   float dt, float eps, float* x_out, float* frame, size_t n, void* stream);

/* ---- rollouts: MoleculeIntegrator.rollout ------------------------------------------------- */

enum { TIB_METHOD_EULER = 0, TIB_METHOD_MIDPOINT = 1, TIB_METHOD_RK4 = 2 };

typedef struct {
  int32_t      method;        /* TIB_METHOD_* (torchdiffeq fixed-grid solvers) */
  int32_t      n_times;       /* T = len(t_grid) = n_step */
  const float* t_grid;        /* HOST [T], torch.linspace(start,end,n_step) in fp32 */
  int32_t      save_frames;   /* 1: out_xts is [T,N,3]; 0: out_xts is [N,3] (final state only) */
  float        eps;           /* Euler-Maruyama diffusion (0 = ODE); EULER only */
  const float* noise;         /* DEVICE [T-1,N,3] pre-drawn N(0,1), or NULL */
  tib_model*   score_model;   /* second network standing in for the score, or NULL (extension; no reference oracle) */
} tib_fixed_opts;

/* Replaces MoleculeIntegrator(b, method in {'euler','midpoint','rk4'}, n_step).rollout(batch)
 * with return_dlogp=False (mdqm9/thermo/ambient/integrators.py:28-33,55-68).  No host sync. */
int tib_rollout_fixed(tib_model* m, const tib_batch* b, const float* x0, const tib_fixed_opts* o,
                      float* out_xts, void* workspace, size_t workspace_bytes, void* stream);

typedef struct {
  double       rtol, atol;
  int32_t      n_times;       /* T */
  const float* t_grid;        /* HOST [T] fp32, increasing */
  int32_t      save_frames;   /* 1: out_xts [T,N,3]; 0: [N,3] final state */
  int32_t      max_attempts;  /* safety bound on attempted steps (0 = 100000) */
  /* Optional hook run after the local error sums are on the host: all-reduce {sum_sq, count} so
   * that shards share one step sequence (torchdiffeq's RMS norm is over the whole batch).
   * NULL = local norm. */
  void       (*norm_allreduce)(double* sum_sq_and_count /*[2]*/, void* user);
  void*        norm_user;
} tib_dopri5_opts;

typedef struct { int32_t nfe, attempts, accepted; double last_dt; } tib_dopri5_stats;

/* Replaces MoleculeIntegrator(b, 'dopri5', n_step, atol, rtol).rollout(batch), return_dlogp=False
 * (mdqm9/thermo/ambient/integrators.py:55-68 -> torchdiffeq 0.2.5 dopri5).  Synchronises the stream
 * once per attempted step to read the error ratio (as torchdiffeq does). */
int tib_rollout_dopri5(tib_model* m, const tib_batch* b, const float* x0, const tib_dopri5_opts* o,
                       float* out_xts, tib_dopri5_stats* stats,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- rollouts of the tuple state (x, dlogp): MoleculeIntegrator(return_dlogp=True) -----------------------------------
 * Replaces rollout() with return_dlogp=True (mdqm9/thermo/ambient/integrators.py:36-53; latent :57-74): the state is the
 * tuple (x [N,3], dlogp [B]) that torchdiffeq flattens to y = [x | dlogp] (n = 3N + B fp32), the right-hand side is
 * (mult_b * b, mult_d * div) with div the exact divergence (tib_drift_div): mult_b = 1, mult_d = -1e-2 for the ambient
 * wrapper (ode_wrapper.py:47,91), mult_d = -1 for the latent one, signs flipped under reverse_ode.
 * y0 / out are flat: out [T][n] with save_frames, else [n].  The fixed-grid solver takes the grid as given (decreasing
 * grids integrate backwards); dopri5 wants an increasing grid: a decreasing one is passed NEGATED with time_sign = -1
 * (torchdiffeq solves (-t, -f)); its error ratio is the MAX of the RMS norms of the two components.
 * Workspace: tib_div_rollout_workspace_bytes. */
size_t tib_div_rollout_workspace_bytes(const tib_model* m, int32_t n_mol, int32_t n_nodes, int64_t n_edges, int32_t max_atoms);
int tib_rollout_fixed_dlogp(tib_model* m, const tib_batch* b, const float* y0, const tib_fixed_opts* o, float mult_b, float mult_d,
                            float* out, void* workspace, size_t workspace_bytes, void* stream);
int tib_rollout_dopri5_dlogp(tib_model* m, const tib_batch* b, const float* y0, const tib_dopri5_opts* o, float mult_b, float mult_d,
                             float time_sign, float* out, tib_dopri5_stats* stats, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- reweighting statistics ---------------------------------------------------------------
 * Partial sums behind calc_ti_weights / calc_ESS / calc_tfep_dF
 * (mdqm9/analysis/utils/ess.py:8-10,32-35; free_energy.py:41-46):
 *   phi_i = E1_i - E0_i + neg_dlogp_i ; w_i = exp(-phi_i)
 *   out[0] = sum w, out[1] = sum w^2, out[2] = sum exp(-phi)*weight_i (weight = 1 if NULL),
 *   out[3] = sum weight_i, out[4] = n.    All device pointers; out is DEVICE double[5].
 * These are the quantities all-reduced across ranks. */
int tib_reweight_stats(const double* E0, const double* E1, const double* neg_dlogp, const double* weight,
                       size_t n, double* out, void* stream);

/* ---- ADW: FCNetMultiBeta drift + exact 1-D divergence ---------------------------------------
 * Replaces ODEWrapper(b, return_dlogp=True).forward for the 1-D double well
 * (adw/thermo/models/ode_wrapper.py:38-47,55-67; adw/thermo/models/simple.py:38-41).
 * fp64 weights, packed host array: beta_embed (3->H->H->1) then net (3->H x num_layers ->1), each layer W[out,in], b[out]. */
typedef struct tib_adw_model tib_adw_model;
int  tib_adw_create(tib_adw_model** out, int32_t hidden, int32_t num_layers, const double* packed_weights,
                    size_t n_doubles, int device);
void tib_adw_destroy(tib_adw_model* m);
/* x [B] fp64 (the fp32 state widened; torchdiffeq's fixed-grid solvers let the state drift to fp64 with an
 * fp64 model), beta0/beta1 [B] fp64, t scalar (fp32 value) -> b [B] fp64 and div [B] fp64 (= d b/d x,
 * unscaled; may be NULL). */
int  tib_adw_drift_div(tib_adw_model* m, const double* x, const double* beta0, const double* beta1, float t,
                       double* out_b, double* out_div, size_t n, void* stream);

/* ---- training step: loss, gradients, optimiser ------------------------------------------------------------
 * Replaces, for the ambient drift network,
 *     loss = StandardVelocityLoss(LinearInterpolant(a, gamma))(batch0, batch1, model); loss.backward()
 * (mdqm9/thermo/ambient/losses.py:30-85,126-133; interpolants.py:16-33,53-108; mdqm9/train_ambient.py:130,144) and
 *     torch.nn.utils.clip_grad_norm_(model.parameters(), 1); optim.step()      (train_ambient.py:96,146-148).
 * The weights are a DEVICE vector in the packing order of tib_packed_weight_count (they change every optimiser step,
 * so there is no handle and nothing is repacked); the gradient comes back in the same order.  The random draws stay
 * with the caller (the reference draws them from torch's CPU generator: one uniform t per molecule repeated over
 * its atoms, losses.py:46-47, then z = randn(N, 3), interpolants.py:29), so the step is reproducible against the
 * reference on identical draws.  Both antithetic drift evaluations (x_t^+, x_t^-) run as one batch of 2 n_mol
 * molecules; every dense contraction (forward, data gradient, weight gradient) is a split-f16 x3 GEMM on tcgen05. */
enum { TIB_GAMMA_BROWNIAN = 0,        /* gamma = sqrt(a t (1 - t))        (interpolants.py:71-76) */
       TIB_GAMMA_SIN2 = 1             /* gamma = sin^2(pi t)              (interpolants.py:78-82) */ };
typedef struct { int32_t gamma_kind; float a; } tib_interpolant;

/* batch0 / batch1 of the training loop in the contract of MDQM9MultiTempDataset.process
 * (mdqm9/data/mdqm9_ambient.py:87-107) + the draws.  All pointers are DEVICE pointers. */
typedef struct {
  int32_t        n_mol;
  int32_t        n_nodes;      /* N */
  int64_t        n_edges;      /* E = sum n_m (n_m - 1): the complete digraph in (src,dst) order, as tib_batch */
  const int32_t* mol_ptr;      /* [n_mol+1] */
  const int64_t* edge_ptr;     /* [n_mol+1] */
  const int32_t* atom_id;      /* [N] batch0.atoms */
  const uint8_t* edge_type;    /* [E] batch0.edge_type */
  const float*   temp0;        /* [N] batch0.T */
  const float*   temp1;        /* [N] batch1.T */
  const float*   x0;           /* [N,3] batch0.x */
  const float*   x1;           /* [N,3] batch1.x */
  const float*   t;            /* [N]   the time of each atom's molecule */
  const float*   z;            /* [N,3] the interpolant's Gaussian noise */
} tib_train_batch;

size_t tib_train_workspace_bytes(const tib_model_desc* desc, int32_t n_mol, int32_t n_nodes, int64_t n_edges);
/* loss: DEVICE double (overwritten); grad: DEVICE [tib_packed_weight_count] (overwritten); out_b: DEVICE [2N,3] or NULL,
 * the drift at x_t^+ (rows 0..N) and x_t^- (rows N..2N).  Asynchronous on `stream`. */
int tib_train_loss_grad(const tib_model_desc* desc, const float* weights, const tib_train_batch* batch, const tib_interpolant* ip,
                        double* loss, float* grad, float* out_b, void* workspace, size_t workspace_bytes, void* stream);
/* Synchronises `stream` and reports (once) a bounded-wait time-out of the training GEMMs.  0 = healthy. */
int tib_train_status(void* stream);
/* g *= min(1, max_grad_norm / (|g|_2 + 1e-6))  (skipped when max_grad_norm <= 0), then torch.optim.Adam's update
 * (L2 weight decay added to the gradient, bias corrections of step `step` >= 1) on flat DEVICE vectors of n floats.
 * `grad` itself is left un-clipped; scratch: DEVICE double[1] (receives |g|_2^2 before clipping).  For data-parallel
 * training all-reduce `grad` between tib_train_loss_grad and this call. */
int tib_adam_step(float* weights, const float* grad, float* m, float* v, size_t n, int32_t step, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float max_grad_norm, double* scratch, void* stream);
/* The general GEMM behind the training step (csrc/train_gemm.cuh), exposed for tests and benchmarks:
 *   C[M][N] (mode 0: =, 1: atomic +=, 2: +=)  sum_k A(m,k) B(n,k) (+ bias[n]),  fp32 in / out, split-f16 x3 on tcgen05;
 *   trans = 0: Op(r,k) = ptr[row(r) * ld + k], trans = 1: Op(r,k) = ptr[row(k) * ld + r]; idx (may be NULL) gathers source rows;
 *   scale = power-of-two operand scale, or amax_a (DEVICE float, may be NULL) = |max| of A for a dynamic one;
 *   c_idx (may be NULL, mode 1 only) scatters result rows; split_k != 0 cuts K over CTAs (mode 1 only). */
int tib_gemm_f16x3(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, int32_t trans_a, const int32_t* idx_a, float scale_a,
                   const float* amax_a, const float* B, int64_t ldb, int32_t trans_b, const int32_t* idx_b, float scale_b, float* C,
                   int64_t ldc, const int32_t* c_idx, const float* bias, int32_t mode, int32_t split_k, void* stream);

/* Pipeline diagnostics of the training GEMM: enable = 1 makes every later k_gemm_tc launch of this thread record clock64()
 * stamps of its CTA (0,0,0): start, TMEM allocated, per K chunk (operands built, barrier passed), all chunks issued,
 * accumulator complete, epilogue done, end.  out (HOST int64[64], may be NULL) receives the stamps of the last launch,
 * out[63] = their count.  Synchronises the device. */
int tib_gemm_debug(int enable, long long* out);
/* Algorithmic FLOPs (2 M N K) of the k_gemm_tc launches issued by this thread since the last reset. */
double tib_train_gemm_flops(int reset);

/* ---- misc -------------------------------------------------------------------------------- */
const char* tib_last_error(void);
int         tib_abi_version(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset. */
uint64_t    tib_launch_count(int reset);


/* Per-kernel-class device time (CUDA event pairs around each launch on the launching stream) between
 * begin and end, for bench.py's roofline line.  ms_sum / launches are HOST arrays [TIB_K_COUNT]. */
enum { TIB_K_EMBED = 0, TIB_K_EDGE_INIT = 1, TIB_K_MESSAGE = 2, TIB_K_UPDATE = 3, TIB_K_READOUT = 4,
       TIB_K_STEP = 5, TIB_K_TRAIN_GEMM = 6 /* k_gemm_tc launches of the training step */,
       TIB_K_TRAIN_OTHER = 7 /* its element-wise / scatter kernels */, TIB_K_COUNT = 8 };
/* Tensor-core plumbing self test: out = A * W^T (transposed = 0) or W * A^T (transposed = 1) through the
 * same operand images, weight ring and TMEM addressing the drift kernels use.  A, out: DEVICE fp32
 * [128][128]; W: HOST fp32 [128][128].  Synchronous. */
int tib_selftest_gemm(const float* A, const float* W_host, float* out, int transposed, void* stream);

/* Pipeline diagnostics of the tensor-core message kernel: enable = 1 allocates per-CTA stall counters that
 * every later launch overwrites; `out` (HOST, [max_ctas][8] int64, may be NULL) receives, per CTA, cycles the
 * MMA thread waited for {0: weights, 1: operands, 3: accumulator drain}, {2: the producer waited for a free
 * ring slot}, {4: MMA thread total}, {5,6: an epilogue thread of the w / phi chain waited for accumulators},
 * {7: ... for output-layer accumulators}.  Synchronises the device. */
int tib_debug_counters(tib_model* m, int enable, long long* out, int max_ctas);

int tib_profile_begin(void);
int tib_profile_end(double* ms_sum, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* TIB_H_ */
