// layered_api.inl - host orchestration of the layered path (tc_chain.cuh GEMM chains + layered.cuh element-wise kernels):
// the F = 256 drift on tensor cores and the exact divergence by forward-mode tangents on tensor cores.
// Included by tib_api.cu inside its anonymous namespace (uses tib_model, Workspace, fail, LAUNCH_CHECK, ProfScope).

// ---- weight streams -------------------------------------------------------------------------------------------------------
// chunks of W[n][k0 + K) (row-major, leading dimension ld) in (128-row block, 32-column chunk) order
void pack_chain_matrix(std::vector<uint16_t>& blob, const float* W, int ld, int n, int k0, int K) {
  for (int nb = 0; nb < n / 128; ++nb)
    for (int kc = 0; kc < K / 32; ++kc) {
      const size_t off = blob.size();
      blob.resize(off + tib::tc::kChunkBytes / 2);
      pack_tc_chunk(blob.data() + off, W, ld, nb * 128, k0 + kc * 32);
    }
}
// one reference MLP (k_in -> F -> F -> n_out) in the consumption order of k_chain_tc; returns max |gain| of both LayerNorms
void pack_chain_mlp(std::vector<uint16_t>& blob, const float* src, int F, int k_in, int n_out, float gmax[2]) {
  const float* W1 = src;
  const float* g1 = W1 + (size_t)F * k_in + F;
  const float* W2 = g1 + 2 * F;
  const float* g2 = W2 + (size_t)F * F + F;
  const float* W3 = g2 + 2 * F;
  for (int h = 0; h < k_in / F; ++h) pack_chain_matrix(blob, W1, k_in, F, h * F, F);
  pack_chain_matrix(blob, W2, F, F, 0, F);
  pack_chain_matrix(blob, W3, F, n_out, 0, F);
  gmax[0] = gmax[1] = 0.0f;
  for (int i = 0; i < F; ++i) { gmax[0] = std::max(gmax[0], std::fabs(g1[i])); gmax[1] = std::max(gmax[1], std::fabs(g2[i])); }
}

// ---- workspace --------------------------------------------------------------------------------------------------------------
struct LayWs {
  int *row_et, *node_local, *row_pair;
  float* pair_dist;
  float *phi3, *w3, *w3d;
  float *sv_phi_n[2], *sv_phi_r[2], *sv_w_n[2], *sv_w_r[2], *sv_upd_n[2], *sv_upd_r[2];
  float *vvuv, *q, *gac;
  // tangents: D = all directions, Dc = directions per phi-chain pass, Dr = directions per readout pass
  float *ts[2], *tv[2], *te, *tout, *phi3d, *tvvuv, *tq, *tgac;
  size_t st_s, st_v, st_e, st_o, st_phi, st_tvvuv, st_tq, st_tgac;
  int D, Dc, Dr;
  static size_t al(size_t x) { return Workspace::align(x); }
  static int dirs_readout(int F) { return F == 256 ? 1 : 3; }
  static int dirs_chunk(int D) { return D < 9 ? D : 9; }      // a multiple of 3 (k_combine_jvp takes whole atoms)
  // with_div = false: what the F = 256 drift needs; true: everything
  static size_t bytes(int F, int n_nodes, long long n_edges, int max_atoms, bool with_div) {
    LayWs w{}; return w.layout(nullptr, F, n_nodes, n_edges, max_atoms, with_div);
  }
  size_t layout(char* base, int F, int n_nodes, long long n_edges, int max_atoms, bool with_div) {
    size_t off = 0;
    auto take = [&](size_t nbytes) { char* r = base ? base + off : nullptr; off += al(nbytes); return r; };
    const size_t N = (size_t)n_nodes, E = (size_t)n_edges, f4 = sizeof(float);
    const size_t te_rows = ((E + 127) / 128) * 128, tn_rows = ((N + 127) / 128) * 128;
    const size_t P = E / 2, tp_rows = ((P + 127) / 128) * 128;
    row_et = (int*)take(sizeof(int) * E);
    node_local = (int*)take(sizeof(int) * N);
    row_pair = (int*)take(sizeof(int) * E);
    pair_dist = (float*)take(f4 * P);
    phi3 = (float*)take(f4 * E * 5 * F);
    w3 = (float*)take(f4 * P * 5 * F);
    vvuv = (float*)take(f4 * 3 * N * 2 * F);
    q = (float*)take(f4 * N * F);
    gac = (float*)take(f4 * N * 3 * F);
    if (!with_div) return off;
    w3d = (float*)take(f4 * P * 5 * F);
    for (int i = 0; i < 2; ++i) {
      sv_phi_n[i] = (float*)take(f4 * te_rows * F); sv_phi_r[i] = (float*)take(f4 * E);
      sv_w_n[i] = (float*)take(f4 * tp_rows * F);   sv_w_r[i] = (float*)take(f4 * P);
      sv_upd_n[i] = (float*)take(f4 * tn_rows * F); sv_upd_r[i] = (float*)take(f4 * N);
    }
    // translation invariance: the three directions of the last atom index are reconstructed (k_div_pick), not propagated
    static const bool no_skip = getenv("TIB_DIV_ALL_DIRECTIONS") != nullptr;      // diagnostics: propagate all 3 n directions
    D = (max_atoms > 2 && !no_skip) ? 3 * (max_atoms - 1) : 3 * max_atoms; Dc = dirs_chunk(D); Dr = dirs_readout(F);
    st_s = al(f4 * N * F) / f4; st_v = al(f4 * N * 3 * F) / f4; st_e = al(f4 * E * F) / f4; st_o = al(f4 * N * 3) / f4;
    st_phi = al(f4 * E * 5 * F) / f4; st_tvvuv = al(f4 * 3 * N * 2 * F) / f4; st_tq = st_s; st_tgac = st_v;
    for (int i = 0; i < 2; ++i) { ts[i] = (float*)take(f4 * st_s * D); tv[i] = (float*)take(f4 * st_v * D); }
    te = (float*)take(f4 * st_e * D);
    tout = (float*)take(f4 * st_o * Dr);
    phi3d = (float*)take(f4 * st_phi * Dc);
    tvvuv = (float*)take(f4 * st_tvvuv * D);
    tq = (float*)take(f4 * st_tq * D);
    tgac = (float*)take(f4 * st_tgac * D);
    return off;
  }
};

int chain_passes(const tib_model* m) { return m->math == TIB_MATH_F16_TC ? 1 : 3; }

template <int F>
int launch_chain(tib_model* m, tib::tc::ChainP& cp, int kind, cudaStream_t st) {
  using namespace tib::tc;
  cp.passes = chain_passes(m); cp.err = m->dev_err;
  cp.n_tiles = (cp.n_rows + 127) / 128;
  const long long work = (long long)cp.n_tiles * cp.n_dirs;
  if (work <= 0) return 0;
  constexpr int NG = F == 128 ? 2 : 4;
  using S = ChainSmem<F, NG>;
  ProfScope ps(kind, st);
  k_chain_tc<F, NG><<<(int)std::min<long long>(work, (long long)m->n_sms * S::CTAS), S::THREADS, S::TOTAL, st>>>(cp);
  LAUNCH_CHECK();
  return 0;
}

tib::tc::ChainSrc src_rows(const float* base, long long ld, const int* idx, int idx_stride, long long dir_stride, float scale) {
  tib::tc::ChainSrc s{}; s.kind = tib::tc::SRC_ROWS; s.base = base; s.ld = ld; s.idx = idx; s.idx_stride = idx_stride;
  s.dir_stride = dir_stride; s.scale = scale; return s;
}
tib::tc::ChainSrc src_pe(int kind, const float* dist, float length) {
  tib::tc::ChainSrc s{}; s.kind = kind; s.base = dist; s.ld = 1; s.scale = length; return s;
}
void chain_mlp_params(tib::tc::ChainP& cp, const tib::MlpW& w, const unsigned char* blob, int n_out) {
  cp.n_hidden = 2; cp.wblob = blob; cp.b1 = w.b1; cp.g1 = w.g1; cp.be1 = w.be1; cp.b2 = w.b2; cp.g2 = w.g2; cp.be2 = w.be2;
  cp.b3 = w.b3; cp.n_out = n_out; cp.ld_out = n_out; cp.out_scale = 1.0f;
}

template <int F>
int lay_setup(tib_model* m, const tib_batch* b, const float* x, Workspace& ws, LayWs& lw, cudaStream_t st) {
  using namespace tib;
  if (!m->ch_blob) return fail("the layered tensor-core path needs n_features 128 or 256");
  if (b->n_edges >= (1ll << 31)) return fail("layered path: n_edges=%lld exceeds int32 row indices", (long long)b->n_edges);
  if (!m->ch_attrs_set) {
    if (set_smem(tc::k_chain_tc<F, (F == 128 ? 2 : 4)>, tc::ChainSmem<F, (F == 128 ? 2 : 4)>::TOTAL)) return -1;
    // two CTAs of the F = 128 variant share an SM only if the whole 228 KB are configured as shared memory
    cudaFuncSetAttribute(tc::k_chain_tc<F, (F == 128 ? 2 : 4)>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    m->ch_attrs_set = true;
  }
  ProfScope ps(TIB_K_EDGE_INIT, st);
  tc::k_edge_tables<<<b->n_mol, 128, 0, st>>>(b->mol_ptr, (const long long*)b->edge_ptr, b->n_mol, x, b->edge_type,
                                              ws.node_in_ptr, ws.rowa, ws.rowb);
  LAUNCH_CHECK();
  const long long n = std::max<long long>(b->n_edges, b->n_mol);
  lay::k_row_aux<<<(int)((n + 255) / 256), 256, 0, st>>>(ws.rowa, (long long)b->n_edges, lw.row_et, b->mol_ptr,
                                                         (const long long*)b->edge_ptr, b->n_mol, lw.node_local, lw.row_pair, lw.pair_dist);
  LAUNCH_CHECK();
  return 0;
}

// SE3Message (cpainn.py:263-310), primal: w chain, phi chain, combine.  `save` keeps the LayerNorm intermediates for the tangents.
template <int F>
int lay_message(tib_model* m, const tib_batch* b, const tib_model::Layer& L, Workspace& ws, LayWs& lw, int cur, bool first,
                bool save, bool do_combine, cudaStream_t st) {
  using namespace tib;
  const int E = (int)b->n_edges;
  {
    tc::ChainP cp{};
    cp.n_rows = E / 2; cp.n_dirs = 1; cp.n_halves = 1; cp.epi = tc::EPI_PRIMAL;     // one row per undirected pair
    cp.src[0] = src_pe(tc::SRC_PE, lw.pair_dist, m->d.length_scale);
    chain_mlp_params(cp, L.w, L.ch_w, 5 * F);
    cp.ascale1 = 1.0f; cp.out = lw.w3;
    if (save) { cp.save_n[0] = lw.sv_w_n[0]; cp.save_n[1] = lw.sv_w_n[1]; cp.save_r[0] = lw.sv_w_r[0]; cp.save_r[1] = lw.sv_w_r[1]; }
    if (launch_chain<F>(m, cp, TIB_K_MESSAGE, st)) return -1;
  }
  {
    tc::ChainP cp{};
    cp.n_rows = E; cp.n_dirs = 1; cp.n_halves = 2; cp.epi = tc::EPI_PRIMAL;
    cp.src[0] = src_rows(ws.s[cur], F, reinterpret_cast<const int*>(ws.rowa), 4, 0, tc::kStateScale);
    cp.src[1] = first ? src_rows(m->edge_emb, F, lw.row_et, 1, 0, tc::kStateScale) : src_rows(ws.e, F, nullptr, 0, 0, tc::kStateScale);
    chain_mlp_params(cp, L.phi, L.ch_phi, 5 * F);
    cp.ascale1 = tc::kStateUnscale; cp.out = lw.phi3;
    if (save) { cp.save_n[0] = lw.sv_phi_n[0]; cp.save_n[1] = lw.sv_phi_n[1]; cp.save_r[0] = lw.sv_phi_r[0]; cp.save_r[1] = lw.sv_phi_r[1]; }
    if (launch_chain<F>(m, cp, TIB_K_MESSAGE, st)) return -1;
  }
  if (do_combine) {
    lay::CombineP c{b->n_nodes, F, ws.node_in_ptr, ws.rowa, reinterpret_cast<const float4*>(ws.rowb), lw.phi3, lw.w3, lw.row_pair,
                    ws.s[cur], ws.v[cur], ws.s[cur ^ 1], ws.v[cur ^ 1], ws.e, m->edge_emb, first ? 1 : 0};
    ProfScope ps(TIB_K_MESSAGE, st);
    lay::k_combine<<<std::min(b->n_nodes, m->n_sms * 16), 128, 0, st>>>(c);
    LAUNCH_CHECK();
  }
  return 0;
}

// Update (cpainn.py:345-376): [V; U] v as one plain GEMM over the (node, xyz) rows, q = |V v|, MLP(cat[q, s]) -> (g, a, c),
// then v += (U v) g, s += q^2 a + c; with D > 0 the same for every tangent direction (LayerNorm intermediates saved).
template <int F>
int lay_update(tib_model* m, const tib_batch* b, const tib_model::Layer& L, Workspace& ws, LayWs& lw, int cur, int D, cudaStream_t st) {
  using namespace tib;
  const int N = b->n_nodes, nblk = m->n_sms * 16;
  {
    tc::ChainP cp{};
    cp.n_rows = 3 * N; cp.n_dirs = 1; cp.n_halves = 1; cp.epi = tc::EPI_PRIMAL; cp.n_hidden = 0;
    cp.src[0] = src_rows(ws.v[cur], F, nullptr, 0, 0, tc::kStateScale);
    cp.wblob = L.ch_uv; cp.n_out = 2 * F; cp.ld_out = 2 * F; cp.out = lw.vvuv; cp.out_scale = tc::kStateUnscale;
    if (launch_chain<F>(m, cp, TIB_K_UPDATE, st)) return -1;
    if (D > 0) {
      tc::ChainP ct = cp;
      ct.epi = tc::EPI_JVP; ct.n_dirs = D;
      ct.src[0] = src_rows(lw.tv[cur], F, nullptr, 0, (long long)lw.st_v, 1.0f);
      ct.out = lw.tvvuv; ct.out_dir_stride = (long long)lw.st_tvvuv;
      if (launch_chain<F>(m, ct, TIB_K_UPDATE, st)) return -1;
    }
  }
  {
    lay::UpdQP qp{N, F, lw.vvuv, lw.q, D, lw.tvvuv, (long long)lw.st_tvvuv, lw.tq, (long long)lw.st_tq};
    ProfScope ps(TIB_K_UPDATE, st);
    lay::k_upd_q<<<std::min(N, nblk), 128, 0, st>>>(qp);
    LAUNCH_CHECK();
  }
  {
    tc::ChainP cp{};
    cp.n_rows = N; cp.n_dirs = 1; cp.n_halves = 2; cp.epi = tc::EPI_PRIMAL;
    cp.src[0] = src_rows(lw.q, F, nullptr, 0, 0, tc::kStateScale);
    cp.src[1] = src_rows(ws.s[cur], F, nullptr, 0, 0, tc::kStateScale);
    chain_mlp_params(cp, L.upd, L.ch_upd, 3 * F);
    cp.ascale1 = tc::kStateUnscale; cp.out = lw.gac;
    if (D > 0) { cp.save_n[0] = lw.sv_upd_n[0]; cp.save_n[1] = lw.sv_upd_n[1]; cp.save_r[0] = lw.sv_upd_r[0]; cp.save_r[1] = lw.sv_upd_r[1]; }
    if (launch_chain<F>(m, cp, TIB_K_UPDATE, st)) return -1;
    if (D > 0) {
      tc::ChainP ct = cp;
      ct.epi = tc::EPI_JVP; ct.n_dirs = D;
      ct.src[0] = src_rows(lw.tq, F, nullptr, 0, (long long)lw.st_tq, 1.0f);
      ct.src[1] = src_rows(lw.ts[cur], F, nullptr, 0, (long long)lw.st_s, 1.0f);
      ct.gmax1 = L.gmax_upd[0]; ct.gmax2 = L.gmax_upd[1];
      ct.out = lw.tgac; ct.out_dir_stride = (long long)lw.st_tgac;
      if (launch_chain<F>(m, ct, TIB_K_UPDATE, st)) return -1;
    }
  }
  {
    lay::UpdApplyP ap{N, F, lw.vvuv, lw.q, lw.gac, ws.s[cur], ws.v[cur], D, lw.tvvuv, (long long)lw.st_tvvuv,
                      lw.tq, (long long)lw.st_tq, lw.tgac, (long long)lw.st_tgac, lw.ts[cur], (long long)lw.st_s,
                      lw.tv[cur], (long long)lw.st_v};
    ProfScope ps(TIB_K_UPDATE, st);
    lay::k_upd_apply<<<std::min(N, nblk), 128, 0, st>>>(ap);
    LAUNCH_CHECK();
  }
  return 0;
}

// drift on the layered path (F = 256 default; also F = 128 for cross-checks): tensor-core message and update MLPs, fp32 scatter / readout
template <int F>
int drift_layered(tib_model* m, const tib_batch* b, const float* x, float t, float* out, Workspace& ws, LayWs& lw, cudaStream_t st) {
  using namespace tib;
  constexpr int RN = Tiles<F>::NODE;
  if (!m->attrs_set) {
    if (set_smem(k_embed<F, RN>, smem_embed<F, RN>())) return -1;
    if (set_smem(k_embed<F, 1, 16>, smem_embed<F, 1>())) return -1;
    if (set_smem(k_message<F, Tiles<F>::MSG>, smem_message<F, Tiles<F>::MSG>())) return -1;
    if (set_smem(k_update<F, RN>, smem_update<F, RN>())) return -1;
    if (set_smem(k_readout<F, RN>, smem_readout<F, RN>())) return -1;
    m->attrs_set = true;
  }
  DriftBatch db{b->n_mol, b->n_nodes, (long long)b->n_edges, b->mol_ptr, (const long long*)b->edge_ptr,
                b->atom_id, b->edge_type, b->temp0, b->temp1};
  const int node_tiles = (b->n_nodes + 8 * RN - 1) / (8 * RN);
  if (launch_embed<F>(m, b, db, t, ws, st)) return -1;
  if (lay_setup<F>(m, b, x, ws, lw, st)) return -1;
  int cur = 0;
  for (int l = 0; l < m->d.n_layers; ++l) {
    const tib_model::Layer& L = m->layers[l];
    if (lay_message<F>(m, b, L, ws, lw, cur, l == 0, false, true, st)) return -1;
    cur ^= 1;
    if (lay_update<F>(m, b, L, ws, lw, cur, 0, st)) return -1;
  }
  ReadoutP rp{b->n_nodes, m->readout, m->Vout, ws.s[cur], ws.v[cur], out};
  ProfScope ps(TIB_K_READOUT, st);
  k_readout<F, RN><<<node_tiles, TIB_THREADS, smem_readout<F, RN>(), st>>>(rp);
  LAUNCH_CHECK();
  return 0;
}

// drift + exact divergence (ode_wrapper.py:39-49,59-91): primal layer by layer with saved LayerNorm intermediates, then the
// tangents of all 3 * max_atoms directions through the same weights.  Work per evaluation relative to one drift: the w MLP
// has ONE tangent (it depends on x through the scalar distance only), the first layer's phi MLP has none (s0, e0 do not
// depend on x), and the directions of the last atom follow from translation invariance, so the tangent GEMMs cost about
// (3 (n - 1) (L - 1) 16 + 14 L) / (30 L) drifts instead of 3 n.
template <int F>
int drift_div_layered(tib_model* m, const tib_batch* b, const float* x, float t, float* out_b, float* out_div, Workspace& ws,
                      LayWs& lw, cudaStream_t st) {
  using namespace tib;
  constexpr int RN = Tiles<F>::NODE;
  constexpr int DR = JvpTiles<F>::D, JRN = JvpTiles<F>::NODE;
  if (!m->jvp_attrs_set) {
    if (set_smem(k_embed<F, RN>, smem_embed<F, RN>())) return -1;
    if (set_smem(k_embed<F, 1, 16>, smem_embed<F, 1>())) return -1;
    if (set_smem(k_message_jvp<F, JvpTiles<F>::MSG, DR>, smem_message_jvp<F, JvpTiles<F>::MSG, DR>())) return -1;
    if (set_smem(k_update_jvp<F, JRN, DR>, smem_update_jvp<F, JRN, DR>())) return -1;
    if (set_smem(k_readout_jvp<F, JRN, DR>, smem_readout_jvp<F, JRN, DR>())) return -1;
    m->jvp_attrs_set = true;
  }
  DriftBatch db{b->n_mol, b->n_nodes, (long long)b->n_edges, b->mol_ptr, (const long long*)b->edge_ptr,
                b->atom_id, b->edge_type, b->temp0, b->temp1};
  const int N = b->n_nodes, E = (int)b->n_edges, D = lw.D;
  if (launch_embed<F>(m, b, db, t, ws, st)) return -1;
  if (lay_setup<F>(m, b, x, ws, lw, st)) return -1;
  const int nblk = m->n_sms * 16;
  int cur = 0;
  for (int l = 0; l < m->d.n_layers; ++l) {
    const tib_model::Layer& L = m->layers[l];
    const bool first = l == 0;
    // ---- message: primal MLPs (saved), d w3 / d dist, then per direction chunk the phi tangents and the tangent scatter
    if (lay_message<F>(m, b, L, ws, lw, cur, first, true, false, st)) return -1;
    {
      tc::ChainP cp{};
      cp.n_rows = E / 2; cp.n_dirs = 1; cp.n_halves = 1; cp.epi = tc::EPI_JVP;
      cp.src[0] = src_pe(tc::SRC_PE_D, lw.pair_dist, m->d.length_scale);
      chain_mlp_params(cp, L.w, L.ch_w, 5 * F);
      cp.gmax1 = L.gmax_w[0]; cp.gmax2 = L.gmax_w[1];
      cp.save_n[0] = lw.sv_w_n[0]; cp.save_n[1] = lw.sv_w_n[1]; cp.save_r[0] = lw.sv_w_r[0]; cp.save_r[1] = lw.sv_w_r[1];
      cp.out = lw.w3d;
      if (launch_chain<F>(m, cp, TIB_K_MESSAGE, st)) return -1;
    }
    lay::CombineP c{N, F, ws.node_in_ptr, ws.rowa, reinterpret_cast<const float4*>(ws.rowb), lw.phi3, lw.w3, lw.row_pair,
                    ws.s[cur], ws.v[cur], ws.s[cur ^ 1], ws.v[cur ^ 1], ws.e, m->edge_emb, first ? 1 : 0};
    for (int d0 = 0; d0 < D; d0 += lw.Dc) {
      const int nd = std::min(lw.Dc, D - d0);
      if (!first) {
        tc::ChainP cp{};
        cp.n_rows = E; cp.n_dirs = nd; cp.n_halves = 2; cp.epi = tc::EPI_JVP;
        cp.src[0] = src_rows(lw.ts[cur] + (size_t)d0 * lw.st_s, F, reinterpret_cast<const int*>(ws.rowa), 4, (long long)lw.st_s, 1.0f);
        cp.src[1] = src_rows(lw.te + (size_t)d0 * lw.st_e, F, nullptr, 0, (long long)lw.st_e, 1.0f);
        chain_mlp_params(cp, L.phi, L.ch_phi, 5 * F);
        cp.gmax1 = L.gmax_phi[0]; cp.gmax2 = L.gmax_phi[1];
        cp.save_n[0] = lw.sv_phi_n[0]; cp.save_n[1] = lw.sv_phi_n[1]; cp.save_r[0] = lw.sv_phi_r[0]; cp.save_r[1] = lw.sv_phi_r[1];
        cp.out = lw.phi3d; cp.out_dir_stride = (long long)lw.st_phi;
        if (launch_chain<F>(m, cp, TIB_K_MESSAGE, st)) return -1;
      }
      lay::CombineJvpP cj{c, x, lw.node_local, lw.w3d, first ? nullptr : lw.phi3d, (long long)lw.st_phi,
                          lw.ts[cur], lw.tv[cur], lw.ts[cur ^ 1], lw.tv[cur ^ 1], lw.te,
                          (long long)lw.st_s, (long long)lw.st_v, (long long)lw.st_e, d0, nd};
      ProfScope ps(TIB_K_MESSAGE, st);
      lay::k_combine_jvp<<<(int)std::min<long long>((long long)N * (nd / 3), nblk), 128, 0, st>>>(cj);
      LAUNCH_CHECK();
    }
    {
      ProfScope ps(TIB_K_MESSAGE, st);
      lay::k_combine<<<std::min(N, nblk), 128, 0, st>>>(c);
      LAUNCH_CHECK();
    }
    cur ^= 1;
    if (lay_update<F>(m, b, L, ws, lw, cur, D, st)) return -1;
  }
  // ---- readout with tangents (fp32, DR directions per pass) and the trace
  const int jtiles = (N + 8 * JRN - 1) / (8 * JRN);
  for (int d0 = 0; d0 < D; d0 += DR) {
    ReadoutJvpP rp{{N, m->readout, m->Vout, ws.s[cur], ws.v[cur], d0 == 0 ? out_b : nullptr},
                   lw.ts[cur] + (size_t)d0 * lw.st_s, lw.tv[cur] + (size_t)d0 * lw.st_v, lw.tout, lw.st_s, lw.st_v, lw.st_o};
    {
      ProfScope ps(TIB_K_READOUT, st);
      k_readout_jvp<F, JRN, DR><<<jtiles, TIB_THREADS, smem_readout_jvp<F, JRN, DR>(), st>>>(rp);
      LAUNCH_CHECK();
    }
    k_div_pick<<<(b->n_mol + 255) / 256, 256, 0, st>>>(b->mol_ptr, b->n_mol, lw.tout, lw.st_o, d0, DR, out_div,
                                                         D < 3 * b->max_atoms ? b->max_atoms : 0);
    LAUNCH_CHECK();
  }
  return 0;
}
