"""Host-side driver of libtib.so: packs a reference `state_dict` into the C ABI's weight order,
prepares batches (validation + int32/uint8 views) and exposes drift / rollout calls on the current
CUDA stream.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .batch import complete_digraph

DEFAULT_TEMPERATURES = [300, 400, 500, 600, 700, 800, 900, 1000]


@dataclass
class Hyper:
    """Constructor arguments of the reference cPaiNN (ambient cpainn.py:23-32; latent cpainn.py:22-31)."""
    n_features: int = 128
    score_layers: int = 5
    temp_length: float = 100
    time_length: float = 10
    n_types: int = 25
    temperatures: Sequence[float] = tuple(DEFAULT_TEMPERATURES)
    variant: str = "ambient"          # "ambient" | "latent"
    length_scale: float = 10          # PaiNNBase default (cpainn.py:127)
    n_edge_types: int = 4             # cpainn.py:70

    @property
    def n_temp_encoders(self) -> int:
        if self.variant == "ambient":
            return 2
        return 1 if len(self.temperatures) > 1 else 0

    @property
    def c_variant(self) -> int:
        if self.variant == "ambient":
            return _lib.VARIANT_AMBIENT
        return _lib.VARIANT_LATENT_MULTI_T if self.n_temp_encoders == 1 else _lib.VARIANT_LATENT_SINGLE_T

    def key_layout(self) -> Dict[str, str]:
        """Index of each parameter-bearing module inside `cPaiNN.net`
        (ambient cpainn.py:67-90; latent cpainn.py:43-72)."""
        if self.variant == "ambient":
            return dict(edge_emb="net.2", atom_emb="net.3", combine="net.7", base="net.8")
        if self.n_temp_encoders == 1:
            return dict(edge_emb="net.2", atom_emb="net.3", combine="net.6", base="net.7")
        return dict(edge_emb="net.2", atom_emb="net.3", combine="net.5", base="net.6")


def packed_keys(hp: Hyper) -> List[tuple]:
    """(state_dict key, shape) of every tensor in the order documented at tib_packed_weight_count (include/tib.h)."""
    k = hp.key_layout()
    out: List[tuple] = []
    F, L = hp.n_features, hp.score_layers

    def take_mlp(prefix, f_in, f_out):
        out.extend([(f"{prefix}.0.weight", (F, f_in)), (f"{prefix}.0.bias", (F,)),
                    (f"{prefix}.1.weight", (F,)), (f"{prefix}.1.bias", (F,)),
                    (f"{prefix}.3.weight", (F, F)), (f"{prefix}.3.bias", (F,)),
                    (f"{prefix}.4.weight", (F,)), (f"{prefix}.4.bias", (F,)),
                    (f"{prefix}.6.weight", (f_out, F)), (f"{prefix}.6.bias", (f_out,))])

    out.append((f"{k['edge_emb']}.embedding.weight", (hp.n_edge_types, F)))
    out.append((f"{k['atom_emb']}.embedding.weight", (hp.n_types, F)))
    take_mlp(f"{k['combine']}.mlp.mlp", (2 + hp.n_temp_encoders) * F, F)
    for l in range(L):
        msg, upd = f"{k['base']}.layers.{2 * l}", f"{k['base']}.layers.{2 * l + 1}"
        take_mlp(f"{msg}.phi.mlp", 2 * F, 5 * F)
        take_mlp(f"{msg}.w.mlp", F, 5 * F)
        out.append((f"{upd}.u.linear.weight", (F, F)))
        out.append((f"{upd}.v.linear.weight", (F, F)))
        take_mlp(f"{upd}.mlp.mlp", 2 * F, 3 * F)
    ro = f"{k['base']}.layers.{2 * L}"
    take_mlp(f"{ro}.mlp.mlp", F, 2)
    out.append((f"{ro}.V.linear.weight", (1, F)))
    return out


def pack_state_dict(sd: Dict[str, torch.Tensor], hp: Hyper) -> np.ndarray:
    """Flattens the reference `state_dict` into the order documented at tib_packed_weight_count
    (include/tib.h).  Tensors stay in their [out,in] layout; the library transposes."""
    parts: List[torch.Tensor] = []
    for name, shape in packed_keys(hp):
        t = sd[name].detach().to("cpu", torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        parts.append(t.reshape(-1))
    return np.ascontiguousarray(torch.cat(parts).numpy())


def model_desc(hp: Hyper) -> "_lib.ModelDesc":
    """tib_model_desc of a reference constructor call (include/tib.h)."""
    temps = torch.tensor(list(hp.temperatures), dtype=torch.float32)
    return _lib.ModelDesc(
        abi_version=_lib.ABI_VERSION, variant=hp.c_variant, n_features=hp.n_features,
        n_layers=hp.score_layers, n_types=hp.n_types, n_edge_types=hp.n_edge_types,
        temp_length=float(hp.temp_length), time_length=float(hp.time_length),
        length_scale=float(hp.length_scale), temp_mean=float(temps.mean()),
        temp_range=float(temps.max() - temps.min()))


def pack_node_tiles(n_atoms: np.ndarray, max_nodes: int = 16, max_rows: int = 128) -> np.ndarray:
    """Greedy packing of consecutive destination nodes into tiles of at most `max_nodes` nodes and `max_rows`
    incoming-edge rows (node j of a molecule with n atoms has n - 1 rows).  Returns tile_node_ptr (int32)."""
    rows = np.repeat(n_atoms - 1, n_atoms).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(rows)])
    n = rows.shape[0]
    ptr = [0]
    j = 0
    while j < n:
        hi = int(np.searchsorted(cum, cum[j] + max_rows, side="right")) - 1      # last node index bound by rows
        nxt = max(j + 1, min(hi, j + max_nodes, n))
        ptr.append(nxt)
        j = nxt
    return np.asarray(ptr, dtype=np.int32)


class PreparedBatch:
    """Device-side view of a batch in the layout `tib_batch` wants; keeps the tensors alive."""

    def __init__(self, batch, hp: Hyper, device: torch.device, validate: bool = True, sampling_tables: bool = True):
        atoms_name = "atoms" if hp.variant == "ambient" else "atom_number"
        atoms = getattr(batch, atoms_name).to(device)
        N = atoms.shape[0]
        bvec = batch.batch.to(device)
        if bvec.shape[0] != N:
            raise ValueError("batch.batch and the atom-id vector disagree on the node count")
        if N == 0:
            raise ValueError("empty batch")
        if validate and bool((bvec[1:] < bvec[:-1]).any()):
            raise ValueError("batch.batch must be sorted (molecules contiguous)")
        n_mol = int(bvec[-1].item()) + 1
        counts = torch.bincount(bvec, minlength=n_mol)
        if validate and bool((counts < 2).any()):
            raise ValueError("every molecule needs at least 2 atoms")
        edge_index, ptr, eptr = complete_digraph(counts)
        ei = batch.edge_index.to(device)
        if validate:
            if tuple(ei.shape) != tuple(edge_index.shape) or not torch.equal(ei, edge_index):
                raise ValueError(
                    "edge_index must be the complete digraph of every molecule in coalesced "
                    "(src,dst) order (cutoff >= molecule diameter; mdqm9/thermo/utils.py:74-78)")
        et = batch.edge_type.to(device)
        if et.shape[0] != edge_index.shape[1]:
            raise ValueError("edge_type length does not match the number of edges")
        if validate and (int(et.min().item()) < 0 or int(et.max().item()) >= hp.n_edge_types):
            raise ValueError(f"edge_type must lie in [0,{hp.n_edge_types})")
        if validate and (int(atoms.min().item()) < 0 or int(atoms.max().item()) >= hp.n_types):
            raise ValueError(f"atom ids must lie in [0,{hp.n_types})")
        self.n_mol, self.n_nodes, self.n_edges = n_mol, N, int(edge_index.shape[1])
        self.max_atoms = int(counts.max().item())
        self.mol_ptr = ptr.to(torch.int32).contiguous()
        self.edge_ptr = eptr.to(torch.int64).contiguous()
        self.atom_id = atoms.to(torch.int32).contiguous()
        self.edge_type = et.to(torch.uint8).contiguous()
        self.batch_vec = bvec
        self.temp0 = self.temp1 = None
        if hp.variant == "ambient":
            self.temp0 = batch.T0.to(device, torch.float32).contiguous()
            self.temp1 = batch.T1.to(device, torch.float32).contiguous()
        elif hp.n_temp_encoders == 1:
            self.temp0 = batch.T.to(device, torch.float32).contiguous()
        if not sampling_tables:      # the training step needs neither the embedding de-duplication nor the node tiling
            self.c = None
            return
        # x-independent node embedding: one row per distinct (atom id, T0, T1) triple
        keys = [self.atom_id.to(torch.int64)]
        for tt in (self.temp0, self.temp1):
            if tt is not None:
                keys.append(tt.view(torch.int32).to(torch.int64))
        uniq, inverse = torch.unique(torch.stack(keys, dim=1), dim=0, return_inverse=True)
        self.n_embed_rows = int(uniq.shape[0])
        self.embed_index = inverse.to(torch.int32).contiguous()
        self.embed_atom_id = uniq[:, 0].to(torch.int32).contiguous()
        self.embed_temp0 = uniq[:, 1].to(torch.int32).view(torch.float32).contiguous() if self.temp0 is not None else None
        self.embed_temp1 = uniq[:, 2].to(torch.int32).view(torch.float32).contiguous() if self.temp1 is not None else None
        dedupe = self.n_embed_rows * 2 <= N
        # destination-node tiles of the tensor-core message kernel (<= 16 nodes, <= 128 incoming-edge rows)
        self.tile_node_ptr = None
        cmin, cmax = int(counts.min().item()), self.max_atoms
        if cmin != cmax:          # uniform batches tile perfectly with the library's default
            self.tile_node_ptr = torch.from_numpy(pack_node_tiles(counts.cpu().numpy())).to(device)
        self.c = _lib.Batch(
            n_mol=n_mol, n_nodes=N, n_edges=self.n_edges, max_atoms=self.max_atoms,
            mol_ptr=self.mol_ptr.data_ptr(), edge_ptr=self.edge_ptr.data_ptr(),
            atom_id=self.atom_id.data_ptr(), edge_type=self.edge_type.data_ptr(),
            temp0=self.temp0.data_ptr() if self.temp0 is not None else None,
            temp1=self.temp1.data_ptr() if self.temp1 is not None else None,
            n_embed_rows=self.n_embed_rows if dedupe else 0,
            embed_index=self.embed_index.data_ptr() if dedupe else None,
            embed_atom_id=self.embed_atom_id.data_ptr() if dedupe else None,
            embed_temp0=self.embed_temp0.data_ptr() if (dedupe and self.embed_temp0 is not None) else None,
            embed_temp1=self.embed_temp1.data_ptr() if (dedupe and self.embed_temp1 is not None) else None,
            n_tiles=int(self.tile_node_ptr.numel() - 1) if self.tile_node_ptr is not None else 0,
            tile_node_ptr=self.tile_node_ptr.data_ptr() if self.tile_node_ptr is not None else None)


class DriftEngine:
    """Owns one `tib_model` handle (weights repacked on one device)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], hp: Hyper, device):
        self.lib = _lib.load()
        self.hp = hp
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("thermodynamic_interpolation_b200 runs on CUDA devices only (no CPU fallback); "
                               f"got device {self.device}")
        desc = model_desc(hp)
        packed = pack_state_dict(state_dict, hp)
        need = self.lib.tib_packed_weight_count(C.byref(desc))
        if packed.size != need:
            raise ValueError(f"state_dict packs to {packed.size} floats, descriptor needs {need}")
        handle = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.tib_model_create(C.byref(handle), C.byref(desc), packed.ctypes.data_as(C.c_void_p),
                                             packed.size, idx), "tib_model_create")
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None
        self._ws_div: Optional[torch.Tensor] = None
        self._graphs: Dict[tuple, tuple] = {}     # captured fixed-grid rollouts (rollout_fixed(graph=True))

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.tib_model_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ------------------------------------------------------------------------------------------
    def set_math(self, mode: int):
        _lib.check(self.lib.tib_model_set_math(self.handle, mode), "tib_model_set_math")

    def status(self):
        """Synchronises the current stream and raises if a tensor-core kernel recorded a pipeline error."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_model_status(self.handle, self._stream()), "tib_model_status")

    def prepare(self, batch, validate: bool = True) -> PreparedBatch:
        return PreparedBatch(batch, self.hp, self.device, validate)

    def _workspace(self, pb: PreparedBatch) -> torch.Tensor:
        need = self.lib.tib_workspace_bytes(self.handle, pb.n_mol, pb.n_nodes, pb.n_edges)
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _aligned(ws: torch.Tensor):
        p = ws.data_ptr()
        off = (-p) % 256
        return p + off, ws.numel() - off

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def drift(self, pb: PreparedBatch, x: torch.Tensor, t: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._state(x, pb)
        if out is None:
            out = torch.empty_like(x)
        wp, wn = self._aligned(self._workspace(pb))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_drift(self.handle, C.byref(pb.c), x.data_ptr(), float(t), out.data_ptr(),
                                          wp, wn, self._stream()), "tib_drift")
        return out

    def drift_div(self, pb: PreparedBatch, x: torch.Tensor, t: float):
        """(b [N,3], div [n_mol]): the drift and its exact divergence sum_ac d b[a][c] / d x[a][c] per molecule,
        unscaled (reference ODEWrapper.compute_divergence, ode_wrapper.py:59-91, before its x 1e-2)."""
        x = self._state(x, pb)
        out = torch.empty_like(x)
        div = torch.empty(pb.n_mol, dtype=torch.float32, device=self.device)
        need = self.lib.tib_div_workspace_bytes(self.handle, pb.n_mol, pb.n_nodes, pb.n_edges, pb.max_atoms)
        wp, wn = self._alloc_div_ws(need, pb)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_drift_div(self.handle, C.byref(pb.c), x.data_ptr(), float(t), out.data_ptr(),
                                              div.data_ptr(), wp, wn, self._stream()), "tib_drift_div")
        return out, div

    def _alloc_div_ws(self, need: int, pb: PreparedBatch):
        if self._ws_div is None or self._ws_div.numel() < need + 256:
            self._ws_div = None
            free, _total = torch.cuda.mem_get_info(self.device)
            if need + 256 > free:
                raise RuntimeError(
                    f"the exact-divergence workspace for {pb.n_mol} molecules ({pb.n_nodes} atoms, up to {pb.max_atoms} per molecule) "
                    f"is {need / 2**30:.1f} GiB (tangent state of {3 * max(pb.max_atoms - 1, 1)} directions) but only {free / 2**30:.1f} GiB "
                    f"are free on {self.device}: evaluate the batch in molecule chunks (dist.shard_batch)")
            self._ws_div = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        return self._aligned(self._ws_div)

    def _div_rollout_ws(self, pb: PreparedBatch):
        need = self.lib.tib_div_rollout_workspace_bytes(self.handle, pb.n_mol, pb.n_nodes, pb.n_edges, pb.max_atoms)
        return self._alloc_div_ws(need, pb)

    def rollout_dlogp(self, pb: PreparedBatch, x0: torch.Tensor, t_grid: torch.Tensor, method: str, *, mult_b: float,
                      mult_d: float, rtol: float = 1e-4, atol: float = 1e-4, save_frames: bool = True, norm_allreduce=None):
        """The tuple state (x, dlogp) of `return_dlogp=True` integrated inside libtib.so (tib_rollout_fixed_dlogp /
        tib_rollout_dopri5_dlogp): right-hand side (mult_b * b, mult_d * div).  Returns (xts [T,N,3], dlogp [T,B], stats)
        - or ([N,3], [B], stats) without frames."""
        x0 = self._state(x0, pb)
        n3, n_mol = x0.numel(), pb.n_mol
        y0 = torch.cat([x0.reshape(-1), torch.zeros(n_mol, dtype=torch.float32, device=self.device)])
        tg = t_grid.detach().to("cpu", torch.float32)
        T = int(tg.shape[0])
        out = torch.empty((T, n3 + n_mol) if save_frames else (n3 + n_mol,), dtype=torch.float32, device=self.device)
        wp, wn = self._div_rollout_ws(pb)
        stats = {}
        with torch.cuda.device(self.device):
            if method in _lib.METHODS:
                tgn = np.ascontiguousarray(tg.numpy())
                opts = _lib.FixedOpts(method=_lib.METHODS[method], n_times=T, t_grid=tgn.ctypes.data_as(C.POINTER(C.c_float)),
                                      save_frames=int(save_frames), eps=0.0, noise=None, score_model=None)
                _lib.check(self.lib.tib_rollout_fixed_dlogp(self.handle, C.byref(pb.c), y0.data_ptr(), C.byref(opts), float(mult_b),
                                                            float(mult_d), out.data_ptr(), wp, wn, self._stream()),
                           "tib_rollout_fixed_dlogp")
            elif method == "dopri5":
                decreasing = T > 1 and float(tg[-1]) < float(tg[0])
                tgn = np.ascontiguousarray((-tg if decreasing else tg).numpy())
                cb = _lib.NORM_ALLREDUCE(norm_allreduce) if norm_allreduce is not None else _lib.NORM_ALLREDUCE()
                opts = _lib.Dopri5Opts(rtol=float(rtol), atol=float(atol), n_times=T, t_grid=tgn.ctypes.data_as(C.POINTER(C.c_float)),
                                       save_frames=int(save_frames), max_attempts=0, norm_allreduce=cb, norm_user=None)
                st = _lib.Dopri5Stats()
                _lib.check(self.lib.tib_rollout_dopri5_dlogp(self.handle, C.byref(pb.c), y0.data_ptr(), C.byref(opts), float(mult_b),
                                                             float(mult_d), -1.0 if decreasing else 1.0, out.data_ptr(), C.byref(st),
                                                             wp, wn, self._stream()), "tib_rollout_dopri5_dlogp")
                stats = dict(nfe=st.nfe, attempts=st.attempts, accepted=st.accepted, last_dt=st.last_dt)
            else:
                raise ValueError(f"unsupported method {method!r}")
        if save_frames:
            return out[:, :n3].reshape(T, -1, 3), out[:, n3:], stats
        return out[:n3].reshape(-1, 3), out[n3:], stats

    def _state(self, x, pb):
        if x.device != self.device or x.dtype != torch.float32 or tuple(x.shape) != (pb.n_nodes, 3):
            raise ValueError(f"state must be a float32 [{pb.n_nodes},3] tensor on {self.device}")
        return x.contiguous()

    def rollout_fixed(self, pb: PreparedBatch, x0: torch.Tensor, t_grid: torch.Tensor, method: str = "euler",
                      save_frames: bool = True, eps: float = 0.0, noise: Optional[torch.Tensor] = None,
                      score_engine: Optional["DriftEngine"] = None, out: Optional[torch.Tensor] = None,
                      graph: bool = False):
        """`graph=True`: the whole rollout (every drift evaluation and state update of every step - tib_rollout_fixed has
        no host synchronisation) is captured once into a CUDA graph per (prepared batch, grid, method, buffers) and replayed:
        one graph launch instead of ~15 kernel launches per step, which is what bounds the reference's real batch sizes
        (12 / 64 conformers: config/ambient/00031_settings_no_300.json:18, 10506_settings_no_900.json:18)."""
        if graph:
            return self._rollout_fixed_graph(pb, x0, t_grid, method, save_frames, eps, noise, score_engine, out)
        x0 = self._state(x0, pb)
        tg = np.ascontiguousarray(t_grid.detach().to("cpu", torch.float32).numpy())
        T = int(tg.shape[0])
        if out is None:
            out = torch.empty((T, pb.n_nodes, 3) if save_frames else (pb.n_nodes, 3), dtype=torch.float32,
                              device=self.device)
        if noise is not None:
            if tuple(noise.shape) != (T - 1, pb.n_nodes, 3) or noise.dtype != torch.float32 or noise.device != self.device:
                raise ValueError("noise must be float32 [T-1,N,3] on the model device")
            noise = noise.contiguous()
        opts = _lib.FixedOpts(method=_lib.METHODS[method], n_times=T,
                              t_grid=tg.ctypes.data_as(C.POINTER(C.c_float)), save_frames=int(save_frames),
                              eps=float(eps), noise=noise.data_ptr() if noise is not None else None,
                              score_model=score_engine.handle if score_engine is not None else None)
        wp, wn = self._aligned(self._workspace(pb))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_rollout_fixed(self.handle, C.byref(pb.c), x0.data_ptr(), C.byref(opts),
                                                  out.data_ptr(), wp, wn, self._stream()), "tib_rollout_fixed")
        return out

    def _rollout_fixed_graph(self, pb, x0, t_grid, method, save_frames, eps, noise, score_engine, out):
        x0 = self._state(x0, pb)
        tg = tuple(float(v) for v in t_grid.detach().to("cpu", torch.float32).tolist())
        key = (id(pb), tg, method, bool(save_frames), float(eps), None if noise is None else noise.data_ptr(),
               None if score_engine is None else id(score_engine), None if out is None else out.data_ptr())
        hit = self._graphs.get(key)
        if hit is None:
            if len(self._graphs) >= 8:               # a handful of shapes per run; do not grow without bound
                self._graphs.clear()
            static_x0 = torch.empty_like(x0)
            static_x0.copy_(x0)
            kw = dict(method=method, save_frames=save_frames, eps=eps, noise=noise, score_engine=score_engine, out=out)
            # warm-up on a side stream (function attributes, workspace allocation), then capture
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                static_out = self.rollout_fixed(pb, static_x0, t_grid, **kw)
            torch.cuda.current_stream(self.device).wait_stream(side)
            kw["out"] = static_out
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.rollout_fixed(pb, static_x0, t_grid, **kw)
            hit = (g, static_x0, static_out, pb, self._ws)      # kept alive: their device addresses are baked into the graph
            self._graphs[key] = hit
        g, static_x0, static_out = hit[:3]
        static_x0.copy_(x0)
        g.replay()
        return static_out

    def rollout_dopri5(self, pb: PreparedBatch, x0: torch.Tensor, t_grid: torch.Tensor, rtol: float, atol: float,
                       save_frames: bool = True, norm_allreduce=None, max_attempts: int = 0,
                       out: Optional[torch.Tensor] = None):
        x0 = self._state(x0, pb)
        tg = np.ascontiguousarray(t_grid.detach().to("cpu", torch.float32).numpy())
        T = int(tg.shape[0])
        if out is None:
            out = torch.empty((T, pb.n_nodes, 3) if save_frames else (pb.n_nodes, 3), dtype=torch.float32,
                              device=self.device)
        cb = _lib.NORM_ALLREDUCE(norm_allreduce) if norm_allreduce is not None else _lib.NORM_ALLREDUCE()
        opts = _lib.Dopri5Opts(rtol=float(rtol), atol=float(atol), n_times=T,
                               t_grid=tg.ctypes.data_as(C.POINTER(C.c_float)), save_frames=int(save_frames),
                               max_attempts=int(max_attempts), norm_allreduce=cb, norm_user=None)
        stats = _lib.Dopri5Stats()
        wp, wn = self._aligned(self._workspace(pb))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_rollout_dopri5(self.handle, C.byref(pb.c), x0.data_ptr(), C.byref(opts),
                                                   out.data_ptr(), C.byref(stats), wp, wn, self._stream()),
                       "tib_rollout_dopri5")
        return out, dict(nfe=stats.nfe, attempts=stats.attempts, accepted=stats.accepted, last_dt=stats.last_dt)

    def step_euler(self, x, b, dt, *, score=None, noise=None, eps=0.0, out=None, frame=None):
        """K1 alone: out = x + dt*b (+ dt*eps*score + sqrt(2 eps dt)*noise)."""
        if out is None:
            out = torch.empty_like(x)
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tib_step_euler(x.data_ptr(), b.data_ptr(), ptr(score), ptr(noise), float(dt),
                                               float(eps), out.data_ptr(), ptr(frame), x.numel(), self._stream()),
                       "tib_step_euler")
        return out


def selftest_gemm(A: torch.Tensor, W: torch.Tensor, transposed: bool) -> torch.Tensor:
    """Tensor-core plumbing self test: A (cuda fp32 [128,128]), W (fp32 [128,128]) -> A @ W.T or W @ A.T."""
    lib = _lib.load()
    A = A.contiguous()
    Wh = np.ascontiguousarray(W.detach().to("cpu", torch.float32).numpy())
    out = torch.empty(128, 128, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        _lib.check(lib.tib_selftest_gemm(A.data_ptr(), Wh.ctypes.data_as(C.c_void_p), out.data_ptr(), int(transposed),
                                         C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)), "tib_selftest_gemm")
    return out


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().tib_launch_count(1 if reset else 0))
