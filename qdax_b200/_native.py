"""Torch-tensor-level wrappers over the C ABI (include/qdx.h).  torch supplies device memory and streams only;
every computation below is a hand-written kernel in libqdx.so.  No CPU fallback: tensors must live on CUDA."""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from qdax_b200 import _lib
from qdax_b200._lib import CvtIndexDesc, GridDesc, call

import os as _os

DEBUG_SYNC = _os.environ.get("QDX_DEBUG_SYNC", "0") not in ("", "0")     # API boundaries synchronise and check the device error flag
PEER_TIMEOUT_MS = int(_os.environ.get("QDX_PEER_TIMEOUT_MS", "30000"))    # p2p exchange: how long a rank waits for its peers' keys

TASK_IDS = {None: -1, "none": -1, "arm": 0, "rastrigin": 1, "sphere": 2}
KEYMODE_KEEP, KEYMODE_UPDATE, KEYMODE_SCAN, KEYMODE_DIST_UPDATE, KEYMODE_EMIT = 0, 1, 2, 3, 4


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: qdax_b200 runs on CUDA only (no CPU fallback); got device {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def key_words(key) -> Tuple[int, int]:
    k = np.asarray(key.cpu() if isinstance(key, torch.Tensor) else key, dtype=np.uint32).reshape(-1)
    if k.size != 2:
        raise ValueError("an RNG key is two uint32 words")
    return int(k[0]), int(k[1])


# ------------------------------------------------------------------------------------------ workspace
class _ErrSlots:
    """Pool of int32 slots in pinned host memory (one cudaHostAlloc per 1024 slots, not one per workspace)."""

    chunks: list = []
    free: list = []

    @classmethod
    def take(cls):
        if not cls.free:
            t = torch.zeros(1024, dtype=torch.int32).pin_memory()
            a = t.numpy()
            cls.chunks.append(t)
            cls.free.extend((t, a[i:i + 1], t.data_ptr() + 4 * i) for i in range(1024))
        return cls.free.pop()

    @classmethod
    def give(cls, slot) -> None:
        cls.free.append(slot)


class Workspace:
    """Per-repertoire device workspace: selection segments, key chain, 64-bit insertion key table."""

    def __init__(self, K: int, device: torch.device):
        n = C.c_int64(0)
        call("qdx_workspace_bytes", C.c_int64(K), C.byref(n))
        self.K = K
        self.buf = torch.empty(n.value, dtype=torch.uint8, device=device)
        call("qdx_workspace_init", _ptr(self.buf), C.c_int64(K), _stream())
        # True once the occupied-cell list / selection segments in the workspace describe the repertoire this
        # workspace belongs to: set by qdx_select_prepare and by every qdx_commit (whose last CTA rescans the
        # fitness array), so a steady-state generation needs no prepare launch.
        self.sel_valid = False
        self.xchg: Optional["PeerExchange"] = None
        # host mirror of the sticky device error flag: one int32 in pinned host memory that the kernels write when they
        # raise the flag, polled (never waited for) at the API boundaries -- raise_if_error()
        self._err_slot = _ErrSlots.take()
        self._err_np = self._err_slot[1]
        self._err_np[0] = 0
        call("qdx_workspace_set_error_mirror", _ptr(self.buf), C.c_void_p(self._err_slot[2]), _stream())

    def __del__(self):
        slot = getattr(self, "_err_slot", None)
        if slot is not None and _ErrSlots is not None:      # (None while the interpreter shuts down)
            _ErrSlots.give(slot)        # the device buffer dies with this object; stream order keeps late writers off a reused slot

    @property
    def ptr(self) -> C.c_void_p:
        return _ptr(self.buf)

    def keytab(self, with_key_slots: bool = False) -> torch.Tensor:
        """int64 view of the K packed insertion keys (all < 2^63, so signed max == unsigned max); with_key_slots also
        covers the 8 * 64 tail slots that carry every rank's generation keys through the same all-reduce."""
        off = C.c_int64(0)
        call("qdx_workspace_keytab_offset", C.c_int64(self.K), C.byref(off))
        n = self.K + (8 * 64 if with_key_slots else 0)
        return self.buf[off.value: off.value + n * 8].view(torch.int64)

    def set_carry_key(self, key) -> None:
        k0, k1 = key_words(key)
        call("qdx_workspace_set_carry_key", self.ptr, C.c_uint32(k0), C.c_uint32(k1), _stream())

    def read(self):
        """Blocking: (carry key (2,) uint32, metrics (4,) float32, error flag)."""
        ck = (C.c_uint32 * 2)()
        m = (C.c_float * 4)()
        err = C.c_int32(0)
        call("qdx_workspace_read", self.ptr, ck, m, C.byref(err), _stream())
        return np.array(list(ck), dtype=np.uint32), np.array(list(m), dtype=np.float32), int(err.value)

    def check(self) -> None:
        """Blocking: synchronise with the stream and raise if the device error flag is set."""
        _, _, err = self.read()
        if err != 0:
            raise _lib.QdxError("device", err)

    def raise_if_error(self) -> None:
        """Non-blocking: raise if a kernel that has already run reported an error (host mirror).  Called at every API
        boundary of the repertoire / drivers, so a device-side error surfaces at the latest one call after it happened
        (immediately in debug mode, QDX_DEBUG_SYNC=1, where the boundaries call check())."""
        if DEBUG_SYNC:
            self.check()
        err = int(self._err_np[0])
        if err != 0:
            raise _lib.QdxError("device", err)


# ------------------------------------------------------------------------------------------ grid detection
@dataclass
class Grid:
    """Separable tessellation detected from a centroid array (compute_euclidean_centroids layout or any
    product grid).  Keeps the axis tensor alive; `desc` is the ctypes struct handed to the C ABI."""

    n: Tuple[int, ...]
    stride: Tuple[int, ...]
    axes: torch.Tensor
    desc: GridDesc


def detect_grid(centroids: torch.Tensor) -> Optional[Grid]:
    """Host-side, one-off: is `centroids` (K, Dd) a product grid  c[sum_d i_d*stride_d] = (ax_0[i_0], ...)?
    Returns None when it is not, or when the spacing is too fine for the fast path's exactness guard
     -- callers then use the brute-force kernel."""
    c = centroids.detach().cpu().numpy().astype(np.float32)
    K, Dd = c.shape
    if Dd < 1 or Dd > 4 or K < 1 or not np.isfinite(c).all():
        return None
    axes, idx = [], []
    for d in range(Dd):
        ax = np.unique(c[:, d])
        axes.append(ax)
        idx.append(np.searchsorted(ax, c[:, d]))
    n = [len(a) for a in axes]
    if int(np.prod(n)) != K or sum(n) > 4096 or max(n) > 1024:
        return None
    stride = []
    for d in range(Dd):
        if n[d] == 1:
            stride.append(0)
            continue
        mask = idx[d] == 1
        for e in range(Dd):
            if e != d:
                mask &= idx[e] == 0
        ks = np.nonzero(mask)[0]
        if len(ks) != 1:
            return None
        stride.append(int(ks[0]))
    flat = sum(idx[d].astype(np.int64) * stride[d] for d in range(Dd))
    if not np.array_equal(flat, np.arange(K)):
        return None
    # exactness guard: candidates two steps away must never tie with the nearest after rounding of the sum
    h_min = min(float(np.min(np.diff(a))) if len(a) > 1 else np.inf for a in axes)
    span = [float(a[-1] - a[0]) + (float(np.min(np.diff(a))) if len(a) > 1 else 1.0) for a in axes]
    lo = [float(axes[d][0]) - span[d] for d in range(Dd)]
    hi = [float(axes[d][-1]) + span[d] for d in range(Dd)]
    worst = sum((2.0 * s + max(abs(l), abs(u))) ** 2 for s, l, u in zip(span, lo, hi))
    if np.isfinite(h_min) and 2.0 * h_min * h_min <= 16.0 * worst * 2.0**-23:
        return None
    ax_t = torch.from_numpy(np.concatenate(axes).astype(np.float32)).to(centroids.device)
    gd = GridDesc()
    gd.dd = Dd
    for d in range(Dd):
        gd.n[d], gd.stride[d], gd.lo[d], gd.hi[d] = n[d], stride[d], lo[d], hi[d]
    gd.axes = ax_t.data_ptr()
    return Grid(tuple(n), tuple(stride), ax_t, gd)


def grid_of(centroids: torch.Tensor) -> Optional[Grid]:
    """Cached detect_grid (cache lives on the tensor object and is keyed on its storage version)."""
    cache = getattr(centroids, "_qdx_grid_cache", None)
    ver = (centroids.data_ptr(), centroids._version, tuple(centroids.shape))
    if cache is None or cache[0] != ver:
        cache = (ver, detect_grid(centroids))
        try:
            centroids._qdx_grid_cache = cache
        except AttributeError:
            pass
    return cache[1]


# ------------------------------------------------------------------------------------------ CVT bucket index
@dataclass
class CvtIndex:
    """Uniform bucket index over low-dimensional (Dd <= 3) non-grid centroids; built once per tessellation on the host
    by libqdx.so (qdx_cvt_index_plan / qdx_cvt_index_build) and uploaded.  Keeps the device arrays alive."""

    desc: CvtIndexDesc
    start: torch.Tensor
    ids: torch.Tensor
    pts: torch.Tensor


def build_cvt_index(centroids: torch.Tensor) -> Optional[CvtIndex]:
    c = np.ascontiguousarray(centroids.detach().cpu().numpy(), dtype=np.float32)
    K, Dd = c.shape
    plan = CvtIndexDesc()
    nb = C.c_int64(0)
    rc = _lib.lib().qdx_cvt_index_plan(C.c_void_p(c.ctypes.data), C.c_int64(K), C.c_int32(Dd), C.byref(plan), C.byref(nb))
    if rc == -2:                                    # QDX_ERR_UNSUPPORTED: no index applies
        return None
    if rc != 0:
        raise _lib.QdxError("qdx_cvt_index_plan", rc)
    start = np.empty(nb.value + 1, dtype=np.int32)
    ids = np.empty(K, dtype=np.int32)
    pts = np.empty((K, Dd), dtype=np.float32)
    call("qdx_cvt_index_build", C.c_void_p(c.ctypes.data), C.c_int64(K), C.byref(plan), C.c_void_p(start.ctypes.data),
         C.c_void_p(ids.ctypes.data), C.c_void_p(pts.ctypes.data))
    dev = centroids.device
    ts, ti, tp = torch.from_numpy(start).to(dev), torch.from_numpy(ids).to(dev), torch.from_numpy(pts).to(dev)
    plan.start, plan.ids, plan.pts = ts.data_ptr(), ti.data_ptr(), tp.data_ptr()
    return CvtIndex(plan, ts, ti, tp)


INDEX_MAX_DIM, INDEX_MIN_CENTROIDS = 3, 256


def cvt_index_of(centroids: torch.Tensor) -> Optional[CvtIndex]:
    """Cached build_cvt_index (cache on the tensor object, keyed on its storage version); None when the tessellation
    is a grid (fast path), too small, or more than 3-dimensional."""
    K, Dd = centroids.shape
    if Dd > INDEX_MAX_DIM or K < INDEX_MIN_CENTROIDS:
        return None
    cache = getattr(centroids, "_qdx_index_cache", None)
    ver = (centroids.data_ptr(), centroids._version, tuple(centroids.shape))
    if cache is None or cache[0] != ver:
        cache = (ver, build_cvt_index(centroids))
        try:
            centroids._qdx_index_cache = cache
        except AttributeError:
            pass
    return cache[1]


def _index_ptr(index: Optional[CvtIndex]):
    return C.byref(index.desc) if index is not None else C.POINTER(CvtIndexDesc)()


def _grid_ptr(grid: Optional[Grid]):
    return C.byref(grid.desc) if grid is not None else C.POINTER(GridDesc)()


# ------------------------------------------------------------------------------------------ kernels
def select_prepare(rep_f: torch.Tensor, ws: Workspace, key_mode: int = KEYMODE_KEEP, key=None, rank_slot: int = -1) -> None:
    k0, k1 = key_words(key) if key is not None else (0, 0)
    call("qdx_select_prepare", _ptr(rep_f), C.c_int64(rep_f.numel()), ws.ptr, C.c_int32(key_mode), C.c_uint32(k0),
         C.c_uint32(k1), C.c_int32(rank_slot), _stream())
    ws.sel_valid = True


def ensure_selection(rep_f: torch.Tensor, ws: Workspace) -> None:
    """Occupancy scan + selection segments, only if the last commit has not already left them in the workspace."""
    if not ws.sel_valid:
        select_prepare(rep_f, ws, KEYMODE_KEEP)


def host_generation_keys(key_mode: int, key=None, carry: Optional[np.ndarray] = None):
    """The jax.random.split chain of one generation, on the host (a few Threefry blocks): returns the 8 key words
    {sel1, sel2, line, leaf} as a ctypes array for qdx_generate / qdx_xchg_push.  key_mode KEYMODE_SCAN advances
    `carry` (uint32[2] NumPy array) in place."""
    out = (C.c_uint32 * 8)()
    k0, k1 = key_words(key) if key is not None else (0, 0)
    cio = (C.c_uint32 * 2)(*(int(x) for x in carry)) if carry is not None else None
    call("qdx_host_generation_keys", C.c_int32(key_mode), C.c_uint32(k0), C.c_uint32(k1), cio, out)
    if carry is not None:
        carry[0], carry[1] = cio[0], cio[1]
    return out


class PeerExchangeUnavailable(RuntimeError):
    """cudaIpc / peer access is not available between the ranks (raised on every rank alike)."""


class PeerExchange:
    """Peer-memory exchange buffers of DistributedMAPElites(exchange="p2p"): this rank's buffer (cudaMalloc, exported
    with cudaIpc) and the mappings of every peer's buffer.  torch.distributed is used only to hand the 64-byte IPC
    handles around (all_gather_object) and for the set-up barrier.  `attach(ws)` records the mappings in a workspace;
    the generation counter lives in the buffer itself, so workspaces may come and go (cloned repertoires)."""

    def __init__(self, K: int, group=None, B_dev: int = 0, D: int = 0, desc_dim: int = 0):
        """B_dev > 0: the buffer also holds this rank's two offspring blocks (rows B_dev x D, fitness, descriptors), which
        the peers read the winners from (qdx_commit mode 3)."""
        import torch.distributed as dist

        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.K = K
        self.shape = (int(B_dev), int(D), int(desc_dim))
        self.local = C.c_void_p(0)
        self.peers = (C.c_void_p * self.size)()
        self._opened = []
        handle = (C.c_char * 64)()
        err: Optional[Exception] = None
        try:
            call("qdx_xchg_create", C.c_int64(K), C.c_int64(B_dev), C.c_int64(D), C.c_int32(desc_dim), C.byref(self.local), C.cast(handle, C.c_void_p))
        except _lib.QdxError as e:
            err = e
        handles = [None] * self.size
        dist.all_gather_object(handles, None if err is not None else bytes(handle.raw), group=group)
        if err is None and all(h is not None for h in handles):
            try:
                for q in range(self.size):
                    if q == self.rank:
                        self.peers[q] = self.local.value
                    else:
                        pp = C.c_void_p(0)
                        hb = C.create_string_buffer(handles[q], 64)
                        call("qdx_xchg_open", C.cast(hb, C.c_void_p), C.byref(pp))
                        self.peers[q] = pp.value
                        self._opened.append(pp.value)
            except _lib.QdxError as e:
                err = e
        elif err is None:
            err = RuntimeError("a peer could not create its exchange buffer")
        oks = [None] * self.size                    # every rank must take the same decision
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            self.close()
            raise PeerExchangeUnavailable(str(err) if err is not None else "a peer could not map the exchange buffers")
        torch.cuda.synchronize()
        dist.barrier(group=group)          # every rank's buffer is zeroed and mapped before anybody pushes

    def attach(self, ws: "Workspace") -> None:
        if ws.xchg is not self:
            if ws.K != self.K:
                raise ValueError("exchange buffers were created for a different number of cells")
            call("qdx_xchg_attach", ws.ptr, C.c_int32(self.rank), C.c_int32(self.size), self.peers, C.c_int64(self.shape[0]),
                 C.c_int64(self.shape[1]), C.c_int32(self.shape[2]), C.c_int32(PEER_TIMEOUT_MS), _stream())
            ws.xchg = self

    def close(self) -> None:
        torch.cuda.synchronize()
        for p in self._opened:
            call("qdx_xchg_close", C.c_void_p(p))
        self._opened = []
        if self.local.value:
            call("qdx_xchg_destroy", self.local)
            self.local = C.c_void_p(0)


def xchg_push(ws: Workspace, gen_keys) -> None:
    call("qdx_xchg_push", ws.ptr, C.c_int64(ws.K), gen_keys, _stream())


def elect_winners(ws: Workspace, rep_g: torch.Tensor, task: str, desc_dim: int, B_dev: int, nranks: int, iso_sigma: float,
                  line_sigma: float, minval, maxval, first_wins: bool, stage_g, stage_f, stage_d, wait_peers: bool = False) -> None:
    """wait_peers: acquire-spin on the peers' arrival flags for at most PEER_TIMEOUT_MS (then QDX_ERR_PEER_TIMEOUT)."""
    K, D = rep_g.shape
    call("qdx_elect_winners", ws.ptr, C.c_int64(K), C.c_int64(D), C.c_int32(TASK_IDS[task]), C.c_int32(desc_dim), C.c_int64(B_dev),
         C.c_int32(nranks), _ptr(rep_g), C.c_float(iso_sigma), C.c_float(line_sigma), C.c_int32(minval is not None),
         C.c_float(minval or 0.0), C.c_int32(maxval is not None), C.c_float(maxval or 0.0), C.c_int32(bool(first_wins)),
         _ptr(stage_g), _ptr(stage_f), _ptr(stage_d), C.c_int32(PEER_TIMEOUT_MS if wait_peers else 0), _stream())


def regenerate_winners(ws: Workspace, rep_g: torch.Tensor, B_dev: int, nranks: int, iso_sigma: float, line_sigma: float, minval,
                       maxval, first_wins: bool, stage_g: torch.Tensor) -> None:
    K, D = rep_g.shape
    call("qdx_regenerate_winners", ws.ptr, C.c_int64(K), C.c_int64(D), C.c_int64(B_dev), C.c_int32(nranks), _ptr(rep_g),
         C.c_float(iso_sigma), C.c_float(line_sigma), C.c_int32(minval is not None), C.c_float(minval or 0.0),
         C.c_int32(maxval is not None), C.c_float(maxval or 0.0), C.c_int32(bool(first_wins)), _ptr(stage_g), _stream())


def generate(rep_g, rep_f, centroids, ws: Workspace, B: int, iso_sigma: float, line_sigma: float, minval, maxval,
             task: Optional[str], desc_dim: int, grid: Optional[Grid], offer: bool, idx_base: int, first_wins: bool,
             out_g, out_f, out_d, out_cells=None, out_p1=None, out_p2=None, gen_keys=None, index: Optional[CvtIndex] = None,
             fired_rows_only: bool = False, out_xchg: bool = False) -> None:
    """fired_rows_only / out_xchg: QDX_GEN_ROWS_FIRED_ONLY / QDX_GEN_OUT_XCHG of include/qdx.h."""
    K, D = rep_g.shape
    call("qdx_generate", _ptr(rep_g), _ptr(rep_f), _ptr(centroids), ws.ptr, C.c_int64(K), C.c_int64(D), C.c_int64(B),
         C.c_float(iso_sigma), C.c_float(line_sigma), C.c_int32(minval is not None), C.c_float(minval or 0.0),
         C.c_int32(maxval is not None), C.c_float(maxval or 0.0), C.c_int32(TASK_IDS[task]), C.c_int32(desc_dim),
         _grid_ptr(grid), C.c_int32(bool(offer)), C.c_uint32(idx_base), C.c_int32(bool(first_wins)), _ptr(out_g), _ptr(out_f),
         _ptr(out_d), _ptr(out_cells), _ptr(out_p1), _ptr(out_p2), gen_keys, _index_ptr(index),
         C.c_int32((1 if fired_rows_only else 0) | (2 if out_xchg else 0)), _stream())


class GenerationStep:
    """A filled qdx_step_desc + everything it points at (kept alive here): one generation = ONE C-ABI call
    (qdx_map_elites_step).  Built once per (repertoire buffers, offspring buffers, configuration) and reused while the
    repertoire is updated in place."""

    def __init__(self, rep_g, rep_f, rep_d, centroids, ws: Workspace, B: int, cfg: dict, grid: Optional[Grid], index: Optional[CvtIndex],
                 first_wins: bool, buf: dict, rank: int = 0, nranks: int = 1):
        K, D = rep_g.shape
        Dd = cfg["desc_dim"]
        d = _lib.StepDesc()
        self.keep = [rep_g, rep_f, rep_d, centroids, ws, buf, grid, index]
        d.rep_genotypes, d.rep_fitness, d.rep_desc, d.centroids, d.ws = rep_g.data_ptr(), rep_f.data_ptr(), rep_d.data_ptr(), centroids.data_ptr(), ws.buf.data_ptr()
        d.K, d.D, d.B, d.desc_dim, d.task = K, D, B, Dd, TASK_IDS[cfg["task"]]
        d.iso_sigma, d.line_sigma = cfg["iso_sigma"], cfg["line_sigma"]
        d.has_min, d.minval = int(cfg["minval"] is not None), cfg["minval"] or 0.0
        d.has_max, d.maxval = int(cfg["maxval"] is not None), cfg["maxval"] or 0.0
        self.launches = 2
        if grid is not None:
            d.grid = C.pointer(grid.desc)
        elif index is not None:
            d.cvt = C.pointer(index.desc)
        else:
            self.launches += 1
            if TC_MIN_DIM <= Dd <= TC_MAX_DIM and K >= TC_MIN_CENTROIDS:
                prep = tc_prep_of(centroids)
                scratch = torch.empty(B + 64, dtype=torch.int32, device=rep_g.device)
                self.keep += [prep, scratch]
                d.tc_prep, d.tc_scratch = prep.data_ptr(), scratch.data_ptr()
                self.launches += 1
            if nranks > 1:
                self.launches += 1                      # qdx_xchg_push
        d.first_wins, d.qd_offset = int(bool(first_wins)), cfg["qd_offset"]
        d.off_genotypes, d.off_fitness, d.off_desc, d.off_cells = buf["g"].data_ptr(), buf["f"].data_ptr(), buf["d"].data_ptr(), buf["c"].data_ptr()
        d.rank, d.nranks, d.exchange = rank, nranks, (1 if nranks > 1 else 0)
        self.desc = d
        self.ws = ws
        self.ref = C.byref(d)
        self.fn = _lib.lib().qdx_map_elites_step

    def run(self, key_mode: int, key, carry: Optional[np.ndarray], metrics_out: torch.Tensor) -> None:
        k0, k1 = key_words(key) if key is not None else (0, 0)
        cio = None
        if carry is not None:
            cio = (C.c_uint32 * 2)(int(carry[0]), int(carry[1]))
        rc = self.fn(self.ref, key_mode, k0, k1, cio, metrics_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.launch_count += self.launches
        if rc != 0:
            raise _lib.QdxError("qdx_map_elites_step", rc)
        if carry is not None:
            carry[0], carry[1] = cio[0], cio[1]
        self.ws.sel_valid = True


def score(task: str, g: torch.Tensor, desc_dim: int = 2, out_f: Optional[torch.Tensor] = None,
          out_d: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    g = require_cuda(g, "genotypes")
    B, D = g.shape
    f = torch.empty(B, dtype=torch.float32, device=g.device) if out_f is None else out_f
    d = torch.empty(B, desc_dim, dtype=torch.float32, device=g.device) if out_d is None else out_d
    call("qdx_score", C.c_int32(TASK_IDS[task]), _ptr(g), C.c_int64(B), C.c_int64(D), C.c_int32(desc_dim), _ptr(f), _ptr(d), _stream())
    return f, d


def score_noisy_arm(g: torch.Tensor, key, fit_variance: float, desc_variance: float, params_variance: float) -> Tuple[torch.Tensor, torch.Tensor]:
    g = require_cuda(g, "genotypes")
    B, D = g.shape
    k0, k1 = key_words(key)
    f = torch.empty(B, dtype=torch.float32, device=g.device)
    d = torch.empty(B, 2, dtype=torch.float32, device=g.device)
    call("qdx_score_noisy_arm", _ptr(g), C.c_int64(B), C.c_int64(D), C.c_uint32(k0), C.c_uint32(k1), C.c_float(fit_variance),
         C.c_float(desc_variance), C.c_float(params_variance), _ptr(f), _ptr(d), _stream())
    return f, d


TC_MIN_DIM, TC_MAX_DIM, TC_MIN_CENTROIDS = 8, 32, 1024     # where the tensor-core pass pays off


def tc_prep_of(centroids: torch.Tensor) -> torch.Tensor:
    """Per-tessellation buffers of the tensor-core cell assignment (swizzled centroid copy, norms), cached on the
    centroid tensor like the grid descriptor."""
    cache = getattr(centroids, "_qdx_tc_cache", None)
    ver = (centroids.data_ptr(), centroids._version, tuple(centroids.shape))
    if cache is None or cache[0] != ver:
        K, Dd = centroids.shape
        nf, ni = C.c_int64(0), C.c_int64(0)
        call("qdx_cells_tc_workspace", C.c_int64(K), C.c_int64(0), C.byref(nf), C.byref(ni))
        prep = torch.empty(nf.value, dtype=torch.float32, device=centroids.device)
        call("qdx_cells_tc_prepare", _ptr(centroids), C.c_int64(K), C.c_int32(Dd), _ptr(prep), _stream())
        cache = (ver, prep)
        try:
            centroids._qdx_tc_cache = cache
        except AttributeError:
            pass
    return cache[1]


def cells_tc(desc: torch.Tensor, centroids: torch.Tensor, ws: Optional["Workspace"] = None, rep_f=None, fitness=None,
             offer: bool = False, idx_base: int = 0, first_wins: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """get_cells_indices through the tcgen05 pass + exact re-rank (desc_dim <= 32)."""
    B, Dd = desc.shape
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=desc.device)
    prep = tc_prep_of(centroids)
    scratch = torch.empty(B + 64, dtype=torch.int32, device=desc.device)
    call("qdx_cells_tc", _ptr(desc), C.c_int64(B), C.c_int32(Dd), _ptr(centroids), C.c_int64(centroids.shape[0]), _ptr(prep),
         _ptr(scratch), _ptr(out), C.c_void_p(0) if ws is None else ws.ptr, _ptr(rep_f), _ptr(fitness), C.c_int32(bool(offer)),
         C.c_uint32(idx_base), C.c_int32(bool(first_wins)), _stream())
    return out


def cells(desc: torch.Tensor, centroids: torch.Tensor, grid: Optional[Grid] = None, ws: Optional[Workspace] = None,
          rep_f: Optional[torch.Tensor] = None, fitness: Optional[torch.Tensor] = None, offer: bool = False,
          idx_base: int = 0, first_wins: bool = True, out: Optional[torch.Tensor] = None, allow_tc: bool = True,
          allow_index: bool = True) -> torch.Tensor:
    B, Dd = desc.shape
    if grid is None and allow_tc and TC_MIN_DIM <= Dd <= TC_MAX_DIM and centroids.shape[0] >= TC_MIN_CENTROIDS and B > 0:
        return cells_tc(desc, centroids, ws, rep_f, fitness, offer, idx_base, first_wins, out)
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=desc.device)
    index = cvt_index_of(centroids) if (grid is None and allow_index and B > 0) else None
    if index is not None:
        call("qdx_cells_indexed", _ptr(desc), C.c_int64(B), _index_ptr(index), C.c_int64(centroids.shape[0]), _ptr(out),
             C.c_void_p(0) if ws is None else ws.ptr, _ptr(rep_f), _ptr(fitness), C.c_int32(bool(offer)), C.c_uint32(idx_base),
             C.c_int32(bool(first_wins)), _stream())
        return out
    call("qdx_cells", _ptr(desc), C.c_int64(B), C.c_int32(Dd), _ptr(centroids), C.c_int64(centroids.shape[0]), _grid_ptr(grid),
         _ptr(out), C.c_void_p(0) if ws is None else ws.ptr, _ptr(rep_f), _ptr(fitness), C.c_int32(bool(offer)),
         C.c_uint32(idx_base), C.c_int32(bool(first_wins)), _stream())
    return out


def offer_cells(cell_idx, fitness, ws: Workspace, rep_f, idx_base: int = 0, first_wins: bool = True) -> None:
    call("qdx_offer_cells", _ptr(cell_idx), _ptr(fitness), C.c_int64(cell_idx.numel()), C.c_int64(ws.K), ws.ptr, _ptr(rep_f),
         C.c_uint32(idx_base), C.c_int32(bool(first_wins)), _stream())


def commit(ws: Workspace, off_g, off_f, off_d, rep_g, rep_f, rep_d, idx_base: int = 0, first_wins: bool = True,
           qd_offset: float = 0.0, metrics_out: Optional[torch.Tensor] = None, added_cells: Optional[torch.Tensor] = None,
           mode: int = 0) -> None:
    K, D = rep_g.shape
    call("qdx_commit", ws.ptr, C.c_int64(K), C.c_int64(D), C.c_int32(rep_d.shape[1]), _ptr(off_g), _ptr(off_f), _ptr(off_d),
         C.c_uint32(idx_base), C.c_int64(0 if off_f is None else off_f.numel()), C.c_int32(bool(first_wins)), _ptr(rep_g), _ptr(rep_f), _ptr(rep_d),
         C.c_float(qd_offset), _ptr(metrics_out), _ptr(added_cells), C.c_int32(mode), _stream())
    if mode != 1:
        ws.sel_valid = True     # the last CTA of the commit kernel rescanned rep_f


def select_indices(ws: Workspace, key, num: int, device) -> torch.Tensor:
    k0, k1 = key_words(key)
    out = torch.empty(num, dtype=torch.int32, device=device)
    call("qdx_select_indices", ws.ptr, C.c_uint32(k0), C.c_uint32(k1), C.c_int64(num), _ptr(out), _stream())
    return out


def select_indices_without_replacement(rep_f: torch.Tensor, ws: Workspace, key, num: int) -> torch.Tensor:
    k0, k1 = key_words(key)
    K = rep_f.numel()
    if num > K:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    out = torch.empty(num, dtype=torch.int32, device=rep_f.device)
    scratch = torch.empty(K, dtype=torch.float32, device=rep_f.device)
    call("qdx_select_indices_without_replacement", _ptr(rep_f), C.c_int64(K), ws.ptr, C.c_uint32(k0), C.c_uint32(k1), C.c_int64(num), _ptr(scratch),
         _ptr(out), _stream())
    return out


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    src2 = src.reshape(src.shape[0], -1)
    out = torch.empty((idx.numel(), src2.shape[1]), dtype=torch.float32, device=src.device)
    call("qdx_gather_rows", _ptr(src2), _ptr(idx), C.c_int64(idx.numel()), C.c_int64(src2.shape[1]), _ptr(out), _stream())
    return out.reshape((idx.numel(),) + tuple(src.shape[1:]))


def isoline_variation(x1, x2, key, iso_sigma, line_sigma, minval=None, maxval=None) -> torch.Tensor:
    x1 = require_cuda(x1, "x1")
    x2 = require_cuda(x2, "x2")
    B = x1.shape[0]
    D = x1.numel() // max(B, 1)
    k0, k1 = key_words(key)
    out = torch.empty_like(x1)
    call("qdx_isoline_variation", _ptr(x1), _ptr(x2), C.c_int64(B), C.c_int64(D), C.c_uint32(k0), C.c_uint32(k1),
         C.c_float(iso_sigma), C.c_float(line_sigma), C.c_int32(minval is not None), C.c_float(minval or 0.0),
         C.c_int32(maxval is not None), C.c_float(maxval or 0.0), _ptr(out), _stream())
    return out


def generate_leaves(rep_g, rep_f, ws: "Workspace", B: int, iso_sigma: float, line_sigma: float, minval, maxval, out_g, gen_keys,
                    leaves, out_p1=None, out_p2=None) -> None:
    """MixingEmitter.emit (variation only) for a pytree genotype stored as packed rows."""
    K, D = rep_g.shape
    call("qdx_generate_leaves", _ptr(rep_g), _ptr(rep_f), ws.ptr, C.c_int64(K), C.c_int64(D), C.c_int64(B), C.c_float(iso_sigma),
         C.c_float(line_sigma), C.c_int32(minval is not None), C.c_float(minval or 0.0), C.c_int32(maxval is not None),
         C.c_float(maxval or 0.0), _ptr(out_g), _ptr(out_p1), _ptr(out_p2), gen_keys, C.byref(leaves), _stream())


def isoline_variation_leaves(x1, x2, line_key, leaves, iso_sigma, line_sigma, minval=None, maxval=None) -> torch.Tensor:
    """isoline_variation on packed (B, D_total) parents of a pytree genotype; line_key = split(key)[1]."""
    x1 = require_cuda(x1, "x1")
    x2 = require_cuda(x2, "x2")
    B, D = x1.shape
    k0, k1 = key_words(line_key)
    out = torch.empty_like(x1)
    call("qdx_isoline_variation_leaves", _ptr(x1), _ptr(x2), C.c_int64(B), C.c_int64(D), C.c_uint32(k0), C.c_uint32(k1), C.byref(leaves),
         C.c_float(iso_sigma), C.c_float(line_sigma), C.c_int32(minval is not None), C.c_float(minval or 0.0),
         C.c_int32(maxval is not None), C.c_float(maxval or 0.0), _ptr(out), _stream())
    return out


def polynomial_mutation(x, key, proportion_to_mutate: float, eta: float, minval: float, maxval: float) -> torch.Tensor:
    x = require_cuda(x, "x")
    B = x.shape[0]
    D = x.numel() // max(B, 1)
    k0, k1 = key_words(key)
    out = torch.empty_like(x)
    call("qdx_polynomial_mutation", _ptr(x), C.c_int64(B), C.c_int64(D), C.c_uint32(k0), C.c_uint32(k1),
         C.c_int32(int(proportion_to_mutate * D)), C.c_float(1.0 + eta), C.c_float(1.0 / (1.0 + eta)), C.c_float(minval),
         C.c_float(maxval), _ptr(out), _stream())
    return out


def polynomial_crossover(x1, x2, key, proportion_var_to_change: float) -> torch.Tensor:
    x1 = require_cuda(x1, "x1")
    x2 = require_cuda(x2, "x2")
    B = x1.shape[0]
    D = x1.numel() // max(B, 1)
    k0, k1 = key_words(key)
    out = torch.empty_like(x1)
    call("qdx_polynomial_crossover", _ptr(x1), _ptr(x2), C.c_int64(B), C.c_int64(D), C.c_uint32(k0), C.c_uint32(k1),
         C.c_int32(int(proportion_var_to_change * D)), _ptr(out), _stream())
    return out


def random_stream(key, n: int, kind: int, device, minval: float = 0.0, maxval: float = 1.0) -> torch.Tensor:
    k0, k1 = key_words(key)
    out = torch.empty(n, dtype=torch.float32 if kind else torch.int32, device=device)
    call("qdx_random", C.c_uint32(k0), C.c_uint32(k1), C.c_int64(n), C.c_int32(kind), C.c_float(minval), C.c_float(maxval),
         _ptr(out), _stream())
    return out


def metrics(rep_f: torch.Tensor, qd_offset: float = 0.0) -> torch.Tensor:
    out = torch.empty(3, dtype=torch.float32, device=rep_f.device)
    call("qdx_metrics", _ptr(rep_f), C.c_int64(rep_f.numel()), C.c_float(qd_offset), _ptr(out), _stream())
    return out


def dns_add(pop_g, pop_f, pop_d, g, f, d, k: int):
    """Returns (new genotypes, new fitnesses (P,), new descriptors, meta (N,), survivors (P,))."""
    P, D = pop_g.shape[0], pop_g.numel() // pop_g.shape[0]
    B = g.shape[0]
    dev = pop_g.device
    out_g = torch.empty_like(pop_g)
    out_f = torch.empty(P, dtype=torch.float32, device=dev)
    out_d = torch.empty_like(pop_d)
    meta = torch.empty(P + B, dtype=torch.float32, device=dev)
    surv = torch.empty(P, dtype=torch.int32, device=dev)
    call("qdx_dns_add", _ptr(pop_g), _ptr(pop_f), _ptr(pop_d), C.c_int64(P), _ptr(g), _ptr(f), _ptr(d), C.c_int64(B),
         C.c_int64(D), C.c_int32(pop_d.shape[1]), C.c_int32(k), _ptr(out_g), _ptr(out_f), _ptr(out_d), _ptr(meta), _ptr(surv),
         _stream())
    return out_g, out_f, out_d, meta, surv


def host_select_table(M: int) -> Tuple[np.ndarray, int]:
    out = np.zeros(M, dtype=np.float32)
    nseg = C.c_int32(0)
    call("qdx_host_select_table", C.c_int32(M), out.ctypes.data_as(C.c_void_p), C.byref(nseg))
    return out, int(nseg.value)


def host_select_rank(M: int, r: Sequence[float]) -> np.ndarray:
    r = np.ascontiguousarray(r, dtype=np.float32)
    out = np.zeros(r.size, dtype=np.int32)
    call("qdx_host_select_rank", C.c_int32(M), r.ctypes.data_as(C.c_void_p), C.c_int64(r.size), out.ctypes.data_as(C.c_void_p))
    return out
