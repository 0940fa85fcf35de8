"""One catalogue over several B200s: redshift slabs with window-wide halos.

Every per-centre computation of the reference reads only galaxies inside the centre's velocity window
(fof.py:67-69,158-160; helper.py:207-209), so a process that holds a redshift slab plus a halo of one window
evaluates its own centres exactly.  What crosses slabs is small and is exchanged with collectives:

  * the candidate centres (global priority = N(0.5) descending, ties by descending global row),
  * the tracker state (alive / decided byte per centre), merged after every priority round,
  * the candidate clusters' member ids for the ordered overlap contests (replicated on every process),
  * the final lists.

The result is identical to the single-GPU run on the same table (tests/test_slabs_gpu.py).  The communicator is
pluggable: ``TorchComm`` (torch.distributed, NCCL over NVLink or gloo) for one process per GPU, ``ThreadComm`` to
emulate several ranks inside one process (used by the tests on a single GPU).
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from . import params as P
from .fof import _as_matrix
from .pipeline import Catalogue
from .units import as_float

C_KMS = 299792.458


# ------------------------------------------------------------------------------------------------------
# slab planning (host logic, no GPU)
# ------------------------------------------------------------------------------------------------------
def slab_edges(z, world):
    """Equal-count redshift boundaries: world + 1 edges, the outer ones infinite."""
    z = np.asarray(z, dtype=np.float64)
    inner = np.quantile(z, np.arange(1, world) / world) if world > 1 else np.zeros(0)
    return np.concatenate([[-np.inf], inner, [np.inf]])


def _widen(lo, hi, beta):
    """Redshift interval that contains the velocity window of every centre in [lo, hi]."""
    m = 1.0 + 1e-6
    wlo = lo - beta * (1.0 + lo) * m - 1e-9 if np.isfinite(lo) else lo
    whi = hi + beta * (1.0 + hi) * m + 1e-9 if np.isfinite(hi) else hi
    return wlo, whi


def plan_slab(edges, rank, max_velocity):
    """(own, valid, local) redshift intervals of one rank.

    own   [lo, hi): centres this rank evaluates;
    valid [lo, hi]: rows whose own velocity window is complete on this rank (their N(0.5) is exact; FoF() rows);
    local [lo, hi]: rows the rank must hold (valid rows plus one more window: neighbours of the halo rows).
    """
    beta = float(max_velocity) / C_KMS
    own = (float(edges[rank]), float(edges[rank + 1]))
    valid = _widen(own[0], own[1], beta)
    local = _widen(valid[0], valid[1], beta)
    return own, valid, local


def global_priority(N, global_row, device=None):
    """Order of the gathered candidate centres: N(0.5) descending, ties by descending global row (oracle pin D1,
    what ``argsort(kind="stable")[::-1]`` gives on the whole table).  With ``device`` the sort runs on that GPU."""
    N = np.asarray(N, dtype=np.float64)
    global_row = np.asarray(global_row, dtype=np.int64)
    if device is not None and len(N) > 100000 and np.all(N == np.floor(N)) and N.max(initial=0) < 2 ** 20 and \
            global_row.max(initial=0) < 2 ** 40:
        import torch
        key = torch.from_numpy(N.astype(np.int64)).to(device) * (1 << 40) + torch.from_numpy(global_row).to(device)
        return torch.sort(key, descending=True).indices.cpu().numpy()
    return np.lexsort((-global_row, -N))


# ------------------------------------------------------------------------------------------------------
# communicators
# ------------------------------------------------------------------------------------------------------
class SoloComm:
    rank, world = 0, 1

    def allgatherv(self, x):
        return [x]

    def allreduce(self, x, op):
        return x

    def allreduce_bytes(self, alive, decided):
        pass

    def barrier(self):
        pass


class ThreadComm:
    """Ranks emulated by threads of one process (collectives through shared lists and a barrier)."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.slots = [None] * world
            self.barrier = threading.Barrier(world)

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def group(cls, world):
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def _exchange(self, x):
        self.s.slots[self.rank] = x
        self.s.barrier.wait()
        out = list(self.s.slots)
        self.s.barrier.wait()
        return out

    def allgatherv(self, x):
        return [np.array(a, copy=True) for a in self._exchange(x)]

    def allreduce(self, x, op):
        vals = self._exchange(np.asarray(x))
        return {"sum": np.sum, "min": np.min, "max": np.max}[op](np.stack(vals), axis=0)

    def allreduce_bytes(self, alive, decided):
        """alive := elementwise MIN, decided := elementwise MAX over ranks (torch CUDA uint8 tensors)."""
        import torch
        a = self._exchange(alive.clone())
        d = self._exchange(decided.clone())
        alive.copy_(torch.stack(a).min(dim=0).values)
        decided.copy_(torch.stack(d).max(dim=0).values)
        self.s.barrier.wait()

    def barrier(self):
        self.s.barrier.wait()


class TorchComm:
    """torch.distributed (NCCL between GPUs; gloo on CPU for the host-logic tests)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.cuda = dist.get_backend() == "nccl"

    def _dev(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if self.cuda else torch.device("cpu")

    def allgatherv(self, x):
        import torch
        x = np.ascontiguousarray(x)
        dev = self._dev()
        n = torch.tensor([x.shape[0]], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        self.dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
        mx = max(max(sizes), 1)
        pad = torch.zeros((mx,) + x.shape[1:], dtype=torch.from_numpy(x[:0]).dtype, device=dev)
        if x.shape[0]:
            pad[: x.shape[0]] = torch.from_numpy(x).to(dev)
        outs = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(outs, pad)
        return [o[:s].cpu().numpy() for o, s in zip(outs, sizes)]

    def allreduce(self, x, op):
        import torch
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)).to(self._dev())
        self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN,
                                    "max": self.dist.ReduceOp.MAX}[op])
        return t.cpu().numpy()

    def allreduce_bytes(self, alive, decided):
        self.dist.all_reduce(alive, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(decided, op=self.dist.ReduceOp.MAX)

    def barrier(self):
        self.dist.barrier()


# ------------------------------------------------------------------------------------------------------
# the driver
# ------------------------------------------------------------------------------------------------------
def _lib_params(max_velocity, linking_length_factor, virial_radius, richness, overdensity):
    return _lib.default_params(
        H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0,
        max_velocity=as_float(P.max_velocity if max_velocity is None else max_velocity, "km/s"),
        linking_length_factor=float(P.linking_length_factor if linking_length_factor is None else linking_length_factor),
        virial_radius=as_float(P.max_radius if virial_radius is None else virial_radius),
        richness=int(P.richness if richness is None else richness),
        overdensity=float(P.D if overdensity is None else overdensity))


def select_local_rows(z, edges, rank, max_velocity):
    """Global row indices (ascending) this rank holds, and its (own, valid) intervals."""
    own, valid, local = plan_slab(edges, rank, max_velocity)
    rows = np.nonzero((z >= local[0]) & (z <= local[1]))[0]
    return rows, own, valid


def find_clusters_slabs(local_rows, global_row, own, valid, comm, handle, params=None, zcut=None, timings=None,
                        mt_state=None):
    """Run the whole path on this rank's rows and assemble the global catalogue.

    local_rows : (n_local, 7) float64, the rows of the global table this rank holds, in global row order;
    global_row : their indices in the global table;  own / valid: intervals from ``plan_slab``.
    mt_state   : (key[624], pos) of numpy's legacy MT19937 generator to draw the overdensity points from; None uses
                 (and advances) this process's ``np.random`` -- every process must then hold the same state.
    Returns (offsets, member_global_rows, centre_global_row, N, D, props) of the GLOBAL final catalogue, identical
    on every rank, in the order the single-GPU run produces (plus the advanced (key, pos) when mt_state is given).
    """
    import time
    import torch
    h = handle
    L = h.lib
    _t = [time.perf_counter()]
    phases = {}

    def lap(name):
        now = time.perf_counter()
        phases[name] = phases.get(name, 0.0) + 1e3 * (now - _t[0])
        _t[0] = now

    p = params or _lib_params(None, None, None, None, None)
    zcut = zcut or (P.min_redshift, P.max_redshift + 0.02, P.max_redshift)
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    grow = np.asarray(global_row, dtype=np.int64)
    rank = comm.rank

    # stage 0 on the local rows; local candidate centres and FoF() rows
    m, keep = _lib.matrix_of(local_rows)
    ng, mc = C.c_int64(), C.c_int64()
    z3 = (C.c_double * 3)(*zcut)
    v2 = (C.c_double * 2)(*valid)
    o2 = (C.c_double * 2)(*own)
    h.check(L.fofb_dist_begin(h.h, C.byref(p), C.byref(m), z3, v2, o2, C.byref(ng), C.byref(mc)))
    n_g, m_loc = ng.value, mc.value
    cen_rows = np.empty((m_loc, 8))
    cen_src = np.empty(m_loc, dtype=np.int32)
    g_src = np.empty(n_g, dtype=np.int32)
    h.check(L.fofb_dist_local_inputs(h.h, cen_rows.ctypes.data, cen_src.ctypes.data, g_src.ctypes.data, None))

    lap("stage0")
    # global priority order of the centres
    parts = comm.allgatherv(np.column_stack([cen_rows, grow[cen_src].astype(np.float64)]) if m_loc else np.zeros((0, 9)))
    owner = np.concatenate([np.full(len(x), r, dtype=np.int32) for r, x in enumerate(parts)])
    allc = np.concatenate(parts, axis=0)
    order = global_priority(allc[:, 7], allc[:, 8].astype(np.int64), device=torch.device("cuda", h.device))
    centres = np.ascontiguousarray(allc[order, :8])
    centre_grow = allc[order, 8].astype(np.int64)
    owned = np.ascontiguousarray((owner[order] == rank).astype(np.uint8))
    m_glob = len(centres)

    lap("gather_centres")
    bbox = np.empty(4)
    h.check(L.fofb_dist_prepare(h.h, centres.ctypes.data, m_glob, owned.ctypes.data, bbox.ctypes.data))
    lo = comm.allreduce(np.array([bbox[0], bbox[2]]), "min")
    hi = comm.allreduce(np.array([bbox[1], bbox[3]]), "max")
    gbbox = np.array([lo[0], hi[0], lo[1], hi[1]])

    lap("prepare")
    # priority rounds, the tracker state merged after each
    dev = torch.device("cuda", h.device)
    alive = torch.empty(max(m_glob, 1), dtype=torch.uint8, device=dev)
    decided = torch.empty(max(m_glob, 1), dtype=torch.uint8, device=dev)
    rounds = 0
    while m_glob > 0:
        h.check(L.fofb_dist_round_evaluate(h.h))
        h.check(L.fofb_dist_state(h.h, alive.data_ptr(), decided.data_ptr(), 0))
        comm.allreduce_bytes(alive, decided)
        torch.cuda.synchronize(dev)
        h.check(L.fofb_dist_state(h.h, alive.data_ptr(), decided.data_ptr(), 1))
        na = C.c_int64()
        h.check(L.fofb_dist_round_advance(h.h, C.byref(na)))
        rounds += 1
        if int(comm.allreduce(np.array([float(na.value)]), "sum")[0]) == 0:
            break
    kc, kt = C.c_int64(), C.c_int64()
    h.check(L.fofb_dist_finalize(h.h, C.byref(kc), C.byref(kt)))
    K_loc, tot_loc = kc.value, kt.value

    lap("rounds")
    # stage 2 on the all-process cluster list
    lc = np.empty(K_loc, dtype=np.int32)
    loff = np.zeros(K_loc + 1, dtype=np.int64)
    lids = np.empty(tot_loc)
    h.check(L.fofb_dist_local_clusters(h.h, lc.ctypes.data, loff.ctypes.data, lids.ctypes.data))
    g_centre = comm.allgatherv(lc)
    g_sizes = comm.allgatherv(np.diff(loff))
    g_ids = comm.allgatherv(lids)
    c_all = np.concatenate(g_centre)
    s_all = np.concatenate(g_sizes)
    cl_owner = np.concatenate([np.full(len(x), r, dtype=np.int32) for r, x in enumerate(g_centre)])
    starts = np.concatenate([[0], np.cumsum(s_all)])[:-1]
    ids_all = np.concatenate(g_ids) if len(g_ids) else np.zeros(0)
    corder = np.argsort(c_all, kind="stable")  # ascending centre index = the order fof.py:232 appends clusters
    K_glob = len(c_all)
    off_sorted = np.zeros(K_glob + 1, dtype=np.int64)
    np.cumsum(s_all[corder], out=off_sorted[1:])
    if K_glob:
        gather = np.concatenate([np.arange(starts[k], starts[k] + s_all[k]) for k in corder]) if off_sorted[-1] else np.zeros(0, dtype=np.int64)
        ids_sorted = np.ascontiguousarray(ids_all[gather.astype(np.int64)])
    else:
        ids_sorted = np.zeros(0)
    centre_sorted = np.ascontiguousarray(c_all[corder].astype(np.int32))
    lap("gather_clusters")
    keep_glob = np.ones(K_glob, dtype=np.int32)
    h.check(L.fofb_dist_overlap(h.h, K_glob, centre_sorted.ctypes.data, off_sorted.ctypes.data, ids_sorted.ctypes.data,
                                keep_glob.ctypes.data))
    keep_glob = (keep_glob != 0)
    gidx_of_sorted = np.cumsum(keep_glob) - 1
    mine_sorted = cl_owner[corder] == rank          # my clusters, ascending centre index = my local order
    keep_local = np.ascontiguousarray(keep_glob[mine_sorted].astype(np.int32))
    gidx_local = np.ascontiguousarray(gidx_of_sorted[mine_sorted & keep_glob].astype(np.int32))
    K_merged = int(keep_glob.sum())
    h.check(L.fofb_dist_set_merged(h.h, keep_local.ctypes.data, gidx_local.ctypes.data, K_merged, gbbox.ctypes.data))

    lap("overlap")
    # stage 4 (every rank consumes the same np.random stream and uses its own clusters' blocks), stages 5-6
    new_state = None
    if mt_state is None:
        Kf, total = h.pipeline_finish_np_random()
    else:
        Kf, total, nkey, npos = h.pipeline_finish_mt(mt_state[0], mt_state[1])
        new_state = (nkey, npos)
    off, rows, cidx, N, D, props = h.pipeline_fetch(Kf, total)
    lap("finish")
    sizes = np.diff(off)
    # assemble the global catalogue in ascending centre priority
    f_c = comm.allgatherv(cidx)
    f_s = comm.allgatherv(sizes)
    f_rows = comm.allgatherv(grow[rows] if total else np.zeros(0, dtype=np.int64))
    f_N = comm.allgatherv(N)
    f_D = comm.allgatherv(D)
    f_p = comm.allgatherv(props.reshape(-1, 8))
    fc = np.concatenate(f_c)
    fs = np.concatenate(f_s)
    fo = np.argsort(fc, kind="stable")
    fstarts = np.concatenate([[0], np.cumsum(fs)])[:-1]
    allrows = np.concatenate(f_rows)
    out_off = np.zeros(len(fc) + 1, dtype=np.int64)
    np.cumsum(fs[fo], out=out_off[1:])
    out_rows = (np.concatenate([allrows[fstarts[k]:fstarts[k] + fs[k]] for k in fo]) if len(fo) and out_off[-1]
                else np.zeros(0, dtype=np.int64))
    lap("gather_final")
    if timings is not None:
        timings.update({"ms_" + k: round(v, 2) for k, v in phases.items()})
        timings.update(rounds=rounds, centres=m_glob, stage1_clusters=K_glob, merged=K_merged, local_rows=len(local_rows),
                       fof_rows=n_g)
    del keep
    res = (out_off, out_rows.astype(np.int64), centre_grow[fc[fo]], np.concatenate(f_N)[fo], np.concatenate(f_D)[fo],
           np.concatenate(f_p, axis=0)[fo])
    return res + (new_state,) if mt_state is not None else res


def find_clusters_distributed(galaxy_data, comm=None, handle=None, **kw):
    """Convenience wrapper: every rank holds the whole table on the host and keeps its slab.

    Returns a ``pipeline.Catalogue`` over the global table (identical on every rank)."""
    g = _as_matrix(galaxy_data)
    comm = comm or SoloComm()
    handle = handle or _lib.get_handle(0)
    p = _lib_params(kw.get("max_velocity"), kw.get("linking_length_factor"), kw.get("virial_radius"), kw.get("richness"),
                    kw.get("overdensity"))
    edges = slab_edges(g[:, 2], comm.world)
    rows, own, valid = select_local_rows(g[:, 2], edges, comm.rank, p.max_velocity)
    res = find_clusters_slabs(g[rows], rows, own, valid, comm, handle, params=p, timings=kw.get("timings"),
                              mt_state=kw.get("mt_state"))
    cat = Catalogue(g, res[0], res[1], res[2], res[3], res[4], res[5])
    if kw.get("mt_state") is not None:
        cat.mt_state = res[6]
    return cat
