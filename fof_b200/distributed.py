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


def slab_edges_balanced(z, world, max_velocity, iterations=6, pair_weight=1.0):
    """Redshift boundaries that equalise the WORK of the rows every rank holds (its own interval widened twice, see
    ``plan_slab``) rather than the number of rows it owns.  Windows grow with (1 + z), so equal-count slabs leave the
    interior ranks with the largest halos; and a row's cost has a part proportional to the number of galaxies in its
    velocity window, i.e. to the catalogue's density in redshift there, dN/dz * (1 + z) (the link tests of stage 0).
    A row at redshift z weighs 1 + pair_weight * rho(z) (1 + z) / <rho (1 + z)>.  ``z`` may be a sorted sample of the
    table's redshifts (every k-th row is enough)."""
    z = np.sort(np.asarray(z, dtype=np.float64))
    if world <= 1 or len(z) < 4 * world:
        return slab_edges(z, world) if len(z) else np.array([-np.inf] + [np.inf] * world)
    beta = float(max_velocity) / C_KMS
    k = max(8, len(z) // 2000)
    idx = np.arange(len(z))
    dz = z[np.minimum(idx + k, len(z) - 1)] - z[np.maximum(idx - k, 0)]
    rho = (np.minimum(idx + k, len(z) - 1) - np.maximum(idx - k, 0)) / np.maximum(dz, 1e-12)
    pair = rho * (1.0 + z)
    w = 1.0 + pair_weight * pair / pair.mean()
    W = np.concatenate([[0.0], np.cumsum(w)])

    def held(lo, hi):
        a, b = _widen(*_widen(lo, hi, beta), beta)
        return W[np.searchsorted(z, b, side="right")] - W[np.searchsorted(z, a, side="left")]

    edges = slab_edges(z, world)
    for _ in range(iterations):
        target = sum(held(edges[r], edges[r + 1]) for r in range(world)) / world
        new = [-np.inf]
        for r in range(world - 1):
            lo_z, hi_z = (z[0] if r == 0 else new[r]), z[-1]
            for _ in range(50):  # held(new[r], x) grows with x
                mid = 0.5 * (lo_z + hi_z)
                if held(new[r], mid) < target:
                    lo_z = mid
                else:
                    hi_z = mid
            new.append(0.5 * (lo_z + hi_z))
        edges = np.array(new + [np.inf])
    return edges


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

    def allgatherv_t(self, t):
        return [t]

    def allreduce(self, x, op):
        return x

    def allreduce_bytes(self, alive, decided):
        pass

    def allreduce_max_bytes(self, buf):
        pass

    def allgather_shares(self, full, share):
        full.copy_(share)
        return None

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
        """Every rank's x, as private copies.  CUDA tensors are produced and consumed on per-rank streams: the producer
        finishes its work before publishing, the consumers finish copying before anybody moves on."""
        cuda = hasattr(x, "is_cuda") and x.is_cuda
        if cuda:
            import torch
            torch.cuda.current_stream(x.device).synchronize()
        self.s.slots[self.rank] = x
        self.s.barrier.wait()
        if cuda:
            out = [a.clone() for a in self.s.slots]
            torch.cuda.current_stream(x.device).synchronize()
        else:
            out = [np.array(a, copy=True) if isinstance(a, np.ndarray) else a for a in self.s.slots]
        self.s.barrier.wait()
        return out

    def allgatherv(self, x):
        return self._exchange(np.asarray(x))

    def allgatherv_t(self, t):
        return self._exchange(t)

    def allreduce(self, x, op):
        vals = self._exchange(np.asarray(x))
        return {"sum": np.sum, "min": np.min, "max": np.max}[op](np.stack(vals), axis=0)

    def allreduce_bytes(self, alive, decided):
        """alive := elementwise MIN, decided := elementwise MAX over ranks (torch CUDA uint8 tensors)."""
        import torch
        a = self._exchange(alive)
        d = self._exchange(decided)
        alive.copy_(torch.stack(a).min(dim=0).values)
        decided.copy_(torch.stack(d).max(dim=0).values)

    def allreduce_max_bytes(self, buf):
        """buf := elementwise MAX over ranks (torch CUDA uint8 tensor), in place."""
        import torch
        parts = self._exchange(buf)
        buf.copy_(torch.stack(parts).max(dim=0).values)

    def allgather_shares(self, full, share):
        """full := the ranks' equal-sized shares one after the other; returns a handle to wait on (or None)."""
        import torch
        full.copy_(torch.cat(self._exchange(share)))
        return None

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

    def _sizes(self, n, dev):
        """Every rank's first dimension: one small all-gather and ONE host read-back."""
        import torch
        mine = torch.tensor([n], dtype=torch.int64, device=dev)
        out = torch.empty(self.world, dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(out, mine)
        return out.tolist()

    def allgatherv(self, x):
        import torch
        x = np.ascontiguousarray(x)
        dev = self._dev()
        return [o.cpu().numpy() for o in self.allgatherv_t(torch.from_numpy(x).to(dev))]

    def allgatherv_t(self, t):
        """All-gather of tensors whose first dimension differs between ranks (padded all_gather on the device)."""
        import torch
        sizes = self._sizes(t.shape[0], t.device)
        mx = max(max(sizes), 1)
        row = tuple(t.shape[1:])
        pad = torch.empty((mx,) + row, dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = torch.empty((self.world * mx,) + row, dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, pad)
        return [out[r * mx: r * mx + s] for r, s in enumerate(sizes)]

    def allreduce(self, x, op):
        import torch
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)).to(self._dev())
        self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN,
                                    "max": self.dist.ReduceOp.MAX}[op])
        return t.cpu().numpy()

    def allreduce_bytes(self, alive, decided):
        self.dist.all_reduce(alive, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(decided, op=self.dist.ReduceOp.MAX)

    def allreduce_max_bytes(self, buf):
        self.dist.all_reduce(buf, op=self.dist.ReduceOp.MAX)

    def allgather_shares(self, full, share):
        return self.dist.all_gather_into_tensor(full, share, async_op=True)

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


class _DevicePointer:
    """A block of device memory owned by the library, described to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def _device_view(ptr, n_words, dev):
    """int32 torch tensor aliasing n_words 32-bit words at device address ptr (no copy)."""
    import torch
    return torch.as_tensor(_DevicePointer(ptr, n_words), device=dev)


def find_clusters_slabs(local_rows, global_row, own, valid, comm, handle, params=None, zcut=None, timings=None,
                        mt_state=None, gather=True):
    """Run the whole path on this rank's rows and assemble the global catalogue.

    local_rows : (n_local, 7) float64 (numpy, or a torch CUDA tensor already on the handle's device), the rows of
                 the global table this rank holds, in global row order;
    global_row : their indices in the global table;  own / valid: intervals from ``plan_slab``;
    mt_state   : (key[624], pos) of numpy's legacy MT19937 generator to draw the overdensity points from; None uses
                 (and advances) this process's ``np.random`` -- every process must then hold the same state.
    gather     : True: every rank returns the GLOBAL catalogue (one all-gather of the final lists); False: every rank
                 returns only the clusters it owns (same layout, global priority order within the rank).
    Every rank works only with the candidate centres of its own redshift neighbourhood (its own plus the neighbours'
    centres inside its valid interval); centres are identified across ranks by their priority key.  The library's
    kernels, the torch glue and the NCCL collectives are all enqueued on ONE stream per rank, so the only host
    synchronisations are the scalar read-backs that size buffers.
    Returns (offsets, member_global_rows, centre_global_row, N, D, props), in the order the single-GPU run produces
    (plus the advanced (key, pos) when mt_state is given).
    """
    import torch
    h = handle
    dev = torch.device("cuda", h.device)
    stream = h.__dict__.get("_slab_stream")
    if stream is None:
        stream = h.__dict__["_slab_stream"] = torch.cuda.Stream(dev)
    stream.wait_stream(torch.cuda.current_stream(dev))   # inputs prepared by the caller on its own stream
    h.use_stream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            return _slabs_on_stream(local_rows, global_row, own, valid, comm, h, dev, params, zcut, timings, mt_state, gather)
    finally:
        stream.synchronize()
        h.use_stream(None)


def _slabs_on_stream(local_rows, global_row, own, valid, comm, h, dev, params, zcut, timings, mt_state, gather):
    import time
    import torch
    L = h.lib
    _t = [time.perf_counter()]
    phases = {}
    launches = [0]

    def lap(name):
        if timings is None:
            return
        torch.cuda.current_stream(dev).synchronize()
        now = time.perf_counter()
        phases[name] = phases.get(name, 0.0) + 1e3 * (now - _t[0])
        _t[0] = now

    def count():
        launches[0] += h.last_timing()[1]

    def cat0(parts):
        return torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]

    p = params or _lib_params(None, None, None, None, None)
    zcut = zcut or (P.min_redshift, P.max_redshift + 0.02, P.max_redshift)
    rank, world = comm.rank, comm.world
    i64, f64 = torch.int64, torch.float64

    # ---- stage 0 on the local rows; local candidate centres and FoF() rows --------------------------------
    if not hasattr(local_rows, "data_ptr"):
        local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    m, keep = _lib.matrix_of(local_rows)
    ng, mc = C.c_int64(), C.c_int64()
    z3 = (C.c_double * 3)(*zcut)
    v2 = (C.c_double * 2)(*valid)
    o2 = (C.c_double * 2)(*own)
    h.check(L.fofb_dist_begin(h.h, C.byref(p), C.byref(m), z3, v2, o2, C.byref(ng), C.byref(mc)))
    count()
    n_g, m_loc = ng.value, mc.value
    # the MT19937 stream of the overdensity points is generated on a side stream while the rounds run
    if mt_state is None:
        name, mt_key, mt_pos, _, _ = np.random.get_state()
    else:
        mt_key, mt_pos = mt_state
    mt_key = np.ascontiguousarray(mt_key, dtype=np.uint32)
    # Once the number of clusters is known from an earlier call (the same on every rank), the stream is generated in
    # shares: every rank produces 1 / world of it and the shares are all-gathered after the rounds.  The first call
    # has only a per-rank guess and generates the whole stream everywhere.
    last = getattr(h, "_slab_last_merged", 0)
    mt_words_t = mt_share = None
    # (worth the extra collective only for long streams: 1e8 words ~ 8e4 clusters; below that everybody generates all)
    if last and world > 1 and last * 4 * p.n_random >= getattr(h, "_mt_share_min_words", 100_000_000):
        wp, boff, share = C.c_uint64(), C.c_int64(), C.c_int64()
        h.check(L.fofb_dist_prefetch_mt_part(h.h, mt_key.ctypes.data, int(mt_pos), int(last * 1.25) + 256, rank, world,
                                             C.byref(wp), C.byref(boff), C.byref(share)))
        if share.value:
            mt_words_t = _device_view(wp.value, boff.value + share.value * world, dev)[boff.value:]
            mt_share = share.value
    else:
        k_spec = int(last * 1.25) + 256 if last else int(m.rows) * world // 150 + 1024
        h.check(L.fofb_dist_prefetch_mt(h.h, mt_key.ctypes.data, int(mt_pos), k_spec))
    cen = torch.empty((max(m_loc, 1), 8), dtype=f64, device=dev)[:m_loc]
    csrc = torch.empty(max(m_loc, 1), dtype=torch.int32, device=dev)[:m_loc]
    h.check(L.fofb_dist_local_inputs(h.h, cen.data_ptr(), csrc.data_ptr(), None, None))
    # global row ids of the local rows (a torch tensor on the device is used as it is)
    grow_t = global_row.to(dev) if hasattr(global_row, "data_ptr") else torch.from_numpy(np.asarray(global_row, dtype=np.int64)).to(dev)
    c_grow = grow_t[csrc.long()]
    lap("stage0")

    # ---- exchange the centres that lie inside another rank's valid interval ---------------------------------
    iv = np.concatenate(comm.allgatherv(np.array([[valid[0], valid[1]]], dtype=np.float64)), axis=0)
    zc = cen[:, 2]
    needed = torch.zeros(m_loc, dtype=torch.bool, device=dev)
    for s in range(world):
        if s != rank:
            needed |= (zc >= iv[s, 0]) & (zc <= iv[s, 1])
    nz_needed = torch.nonzero(needed).flatten()
    parts_rows = comm.allgatherv_t(torch.cat([cen[nz_needed], c_grow[nz_needed].to(f64).unsqueeze(1)], dim=1))
    counts = [len(x) for x in parts_rows]
    offB = np.concatenate([[0], np.cumsum(counts)])
    B = cat0(parts_rows)
    nB = int(offB[-1])
    B_owner = torch.cat([torch.full((c,), r, dtype=torch.int32, device=dev) for r, c in enumerate(counts)]) if nB else \
        torch.zeros(0, dtype=torch.int32, device=dev)
    imp_idx = torch.nonzero((B_owner != rank) & (B[:, 2] >= valid[0]) & (B[:, 2] <= valid[1])).flatten() if nB else \
        torch.zeros(0, dtype=i64, device=dev)
    loc_rows = torch.cat([cen, B[imp_idx, :8]], dim=0)
    loc_grow = torch.cat([c_grow, B[imp_idx, 8].to(i64)])
    loc_owned = torch.cat([torch.ones(m_loc, dtype=torch.uint8, device=dev), torch.zeros(len(imp_idx), dtype=torch.uint8, device=dev)])
    loc_bidx = torch.full((len(loc_grow),), -1, dtype=i64, device=dev)
    loc_bidx[nz_needed] = int(offB[rank]) + torch.arange(len(nz_needed), device=dev)
    loc_bidx[m_loc:] = imp_idx
    key = loc_rows[:, 7].to(i64) * (1 << 40) + loc_grow
    order = torch.argsort(key, descending=True)
    loc_rows = loc_rows[order].contiguous()
    loc_grow, loc_owned, loc_bidx, key = loc_grow[order], loc_owned[order].contiguous(), loc_bidx[order], key[order]
    m_l = len(loc_grow)
    lpos = torch.nonzero(loc_bidx >= 0).flatten().contiguous()
    bpos = loc_bidx[lpos].contiguous()
    n_shared = len(lpos)
    lap("exchange_centres")

    bbox = np.empty(4)
    h.check(L.fofb_dist_prepare(h.h, loc_rows.data_ptr(), m_l, loc_owned.data_ptr(), bbox.ctypes.data))
    ext = comm.allreduce(np.array([-bbox[0], bbox[1], -bbox[2], bbox[3]]), "max")   # one call: min as -max
    gbbox = np.array([-ext[0], ext[1], -ext[2], ext[3]])
    lap("prepare")

    # ---- priority rounds; ONE collective per round: the states of the shared (boundary) centres, merged by MAX
    #      (0 alive and undecided < 1 switched off < 2 evaluated), plus a byte telling whether anybody still had work
    n_buf = nB + 1
    buf = torch.empty(n_buf, dtype=torch.uint8, device=dev)
    rounds = 0
    while True:
        if m_l > 0:
            h.check(L.fofb_dist_round_evaluate(h.h))
        h.check(L.fofb_dist_state_pack(h.h, n_shared, lpos.data_ptr(), bpos.data_ptr(), buf.data_ptr(), n_buf))
        comm.allreduce_max_bytes(buf)
        h.check(L.fofb_dist_state_unpack(h.h, n_shared, lpos.data_ptr(), bpos.data_ptr(), buf.data_ptr()))
        na, any_active = C.c_int64(0), C.c_int32(0)
        h.check(L.fofb_dist_round_advance2(h.h, buf.data_ptr(), n_buf, C.byref(na), C.byref(any_active)))
        if not any_active.value:
            break
        rounds += 1
    kc, kt = C.c_int64(), C.c_int64()
    h.check(L.fofb_dist_finalize(h.h, C.byref(kc), C.byref(kt)))
    K_loc, tot_loc = kc.value, kt.value
    mt_work = None
    if mt_words_t is not None:  # the shares of the MT19937 stream travel while stage 2 is being sorted out
        h.check(L.fofb_dist_mt_wait(h.h))
        mine = mt_words_t[rank * mt_share:(rank + 1) * mt_share].clone()
        mt_work = comm.allgather_shares(mt_words_t, mine)
    lap("rounds")

    # ---- stage 2: every rank lists the contests of ITS clusters (it needs the neighbours' boundary clusters for that:
    #      a contest joins clusters within one velocity window), the contests are all-gathered and resolved alike
    #      everywhere.  Member ids travel only for the boundary clusters; the cluster centres (10 doubles each) for all.
    lc = torch.empty(max(K_loc, 1), dtype=torch.int32, device=dev)[:K_loc]
    ls = torch.empty(max(K_loc, 1), dtype=i64, device=dev)[:K_loc]
    lids = torch.empty(max(tot_loc, 1), dtype=f64, device=dev)[:tot_loc]
    if K_loc:
        h.check(L.fofb_dist_local_clusters_dev(h.h, lc.data_ptr(), ls.data_ptr(), lids.data_ptr()))
    lcl = lc.long()
    zcl = loc_rows[lcl, 2]
    need_cl = torch.zeros(K_loc, dtype=torch.bool, device=dev)
    for s in range(world):
        if s != rank:
            need_cl |= (zcl >= iv[s, 0]) & (zcl <= iv[s, 1])
    # the int64 priority key travels as its bit pattern (N * 2^40 exceeds 2^53 for N >= 8192)
    meta = torch.cat([loc_rows[lcl], key[lcl].view(f64).unsqueeze(1), ls.to(f64).unsqueeze(1), need_cl.to(f64).unsqueeze(1)], dim=1)
    g_meta = comm.allgatherv_t(meta)                                   # K x 11 per rank
    member_cl = torch.repeat_interleave(torch.arange(K_loc, device=dev), ls, output_size=tot_loc)
    g_bids = comm.allgatherv_t(lids[need_cl[member_cl]] if tot_loc else lids)   # member ids of the boundary clusters only
    kcounts = [len(x) for x in g_meta]
    M_all = cat0(g_meta)
    K_glob = len(M_all)
    cl_owner = torch.cat([torch.full((c,), r, dtype=torch.int32, device=dev) for r, c in enumerate(kcounts)]) if K_glob else \
        torch.zeros(0, dtype=torch.int32, device=dev)
    keep_glob = torch.ones(max(K_glob, 1), dtype=torch.int32, device=dev)[:K_glob]
    n_events_glob = 0
    if K_glob:
        keys_all = M_all[:, 8].contiguous().view(i64)
        corder = torch.argsort(keys_all, descending=True)   # priority order = the order fof.py:232 appends clusters
        neg_sorted = -keys_all[corder]                      # ascending, for searchsorted
        # the list this rank works on: its own clusters plus the other ranks' boundary clusters inside its valid interval
        bmask = M_all[:, 10] != 0
        b_sizes = M_all[bmask, 9].to(i64)
        b_starts = torch.cumsum(b_sizes, 0) - b_sizes
        b_rows = M_all[bmask]
        b_owner = cl_owner[bmask]
        imp = torch.nonzero((b_owner != rank) & (b_rows[:, 2] >= valid[0]) & (b_rows[:, 2] <= valid[1])).flatten()
        bids_all = cat0(g_bids)
        own_starts = torch.cumsum(ls, 0) - ls
        L_rows = torch.cat([loc_rows[lcl], b_rows[imp, :8]], dim=0)
        L_keys = torch.cat([key[lcl], b_rows[imp, 8].contiguous().view(i64)])
        L_sizes = torch.cat([ls, b_sizes[imp]])
        L_starts = torch.cat([own_starts, tot_loc + b_starts[imp]])   # into the pool [own ids | all boundary ids]
        L_owned = torch.cat([torch.ones(K_loc, dtype=torch.bool, device=dev), torch.zeros(len(imp), dtype=torch.bool, device=dev)])
        pool = torch.cat([lids, bids_all])
        lo = torch.argsort(L_keys, descending=True)
        L_rows, L_keys, L_sizes, L_starts, L_owned = L_rows[lo].contiguous(), L_keys[lo], L_sizes[lo], L_starts[lo], L_owned[lo]
        K_L = len(L_keys)
        L_off = torch.zeros(K_L + 1, dtype=i64, device=dev)
        L_off[1:] = torch.cumsum(L_sizes, 0)
        tot_L = int(L_off[-1].item())
        seg = torch.repeat_interleave(torch.arange(K_L, device=dev), L_sizes, output_size=tot_L)
        L_ids = pool[L_starts[seg] + (torch.arange(tot_L, device=dev) - L_off[:-1][seg])].contiguous()
        ne = C.c_int64(0)
        h.check(L.fofb_dist_overlap_events(h.h, K_L, L_rows.data_ptr(), L_off.data_ptr(), tot_L, L_ids.data_ptr(), C.byref(ne)))
        E_L = ne.value
        ev = torch.empty((3, max(E_L, 1)), dtype=torch.int32, device=dev)
        if E_L:
            h.check(L.fofb_dist_overlap_events_fetch(h.h, E_L, ev[0].data_ptr(), ev[1].data_ptr(), ev[2].data_ptr()))
        ev = ev[:, :E_L].long()
        gpos = torch.searchsorted(neg_sorted, -L_keys)       # position of every listed cluster in the all-process order
        mine = L_owned[ev[0]] if E_L else torch.zeros(0, dtype=torch.bool, device=dev)
        e_a, e_c, e_l = ev[0][mine], ev[1][mine], ev[2][mine]
        send = torch.stack([gpos[e_a], gpos[e_c.clamp(min=0)], torch.where(e_l >= 0, gpos[e_l.clamp(min=0)], torch.full_like(e_l, -1))],
                           dim=1).to(torch.int32) if E_L else torch.zeros((0, 3), dtype=torch.int32, device=dev)
        send[:, 1] = torch.where(e_c >= 0, send[:, 1], torch.full_like(send[:, 1], -1)) if E_L else send[:, 1]
        ev_all = cat0(comm.allgatherv_t(send))
        n_events_glob = len(ev_all)
        if n_events_glob:
            eo = torch.argsort(ev_all[:, 0], stable=True)    # contests of one cluster come from one rank, already in order
            ev_all = ev_all[eo]
            ga, gc, gl = ev_all[:, 0].contiguous(), ev_all[:, 1].contiguous(), ev_all[:, 2].contiguous()
        else:
            ga = gc = gl = torch.zeros(1, dtype=torch.int32, device=dev)
        h.check(L.fofb_dist_overlap_resolve(h.h, K_glob, n_events_glob, ga.data_ptr(), gc.data_ptr(), gl.data_ptr(), keep_glob.data_ptr()))
        keepb = keep_glob != 0
        gidx_sorted = torch.cumsum(keepb.to(i64), 0) - 1
        mine_sorted = cl_owner[corder] == rank
        keep_local = keepb[mine_sorted].to(torch.int32).contiguous()
        gidx_local = gidx_sorted[mine_sorted & keepb].to(torch.int32).contiguous()
        K_merged = int(keepb.sum().item())
    else:
        keep_local = torch.zeros(1, dtype=torch.int32, device=dev)
        gidx_local = torch.zeros(1, dtype=torch.int32, device=dev)
        K_merged = 0
    h.check(L.fofb_dist_set_merged(h.h, keep_local.data_ptr(), gidx_local.data_ptr(), K_merged, gbbox.ctypes.data))
    h._slab_last_merged = K_merged
    if mt_work is not None:
        mt_work.wait()
    lap("overlap")

    # ---- stage 4 (every rank consumes the same MT19937 stream and uses its own clusters' blocks), stages 5-6 ----
    new_state = None
    if mt_state is None:
        Kf, total = h.pipeline_finish_np_random()
    else:
        Kf, total, nkey, npos = h.pipeline_finish_mt(mt_state[0], mt_state[1])
        new_state = (nkey, npos)
    count()
    off_t, rows_t, cidx32, N_t, D_t, props_t = h.pipeline_fetch_torch(Kf, total, dev)
    lap("finish")

    # ---- the catalogue in priority order: global (one all-gather of the packed lists) or this rank's share -------
    cidx_t = cidx32.long()
    fmeta = torch.cat([key[cidx_t].view(f64).unsqueeze(1), loc_grow[cidx_t].to(f64).unsqueeze(1),
                       (off_t[1:] - off_t[:-1]).to(f64).unsqueeze(1), N_t.unsqueeze(1), D_t.unsqueeze(1), props_t],
                      dim=1) if Kf else torch.zeros((0, 13), dtype=f64, device=dev)
    frows = grow_t[rows_t.long()] if total else torch.zeros(0, dtype=i64, device=dev)
    if gather and world > 1:
        # one exchange: [K, total] per rank, then one padded buffer holding meta (13 doubles per cluster) and the rows
        sizes = np.concatenate(comm.allgatherv(np.array([[Kf, total]], dtype=np.int64)), axis=0)
        packed = torch.cat([fmeta.reshape(-1), frows.view(f64)])
        parts = comm.allgatherv_t(packed)
        f_meta = cat0([pt[: 13 * int(k)].reshape(-1, 13) for pt, (k, t) in zip(parts, sizes)])
        f_rows = cat0([pt[13 * int(k): 13 * int(k) + int(t)].view(i64) for pt, (k, t) in zip(parts, sizes)])
    else:
        f_meta, f_rows = fmeta, frows
    Kt = len(f_meta)
    if Kt:
        fo = torch.argsort(f_meta[:, 0].contiguous().view(i64), descending=True)
        fs_all = f_meta[:, 2].to(i64)
        fstarts = torch.cumsum(fs_all, 0) - fs_all
        fs = fs_all[fo]
        out_off = torch.zeros(Kt + 1, dtype=i64, device=dev)
        out_off[1:] = torch.cumsum(fs, 0)
        tot = int(f_rows.shape[0])
        seg = torch.repeat_interleave(torch.arange(Kt, device=dev), fs, output_size=tot)
        out_rows = f_rows[fstarts[fo][seg] + (torch.arange(tot, device=dev) - out_off[:-1][seg])]
        fm = f_meta[fo]

        def to_host(name, t):  # pinned, reused host buffers: a fresh pageable allocation per call costs more than the copy
            cache = h.__dict__.setdefault("_slab_pinned", {})
            buf_h = cache.get(name)
            if buf_h is None or buf_h.numel() < t.numel() or buf_h.dtype != t.dtype:
                buf_h = cache[name] = torch.empty(int(t.numel() * 1.25) + 16, dtype=t.dtype, pin_memory=True)
            view = buf_h[: t.numel()].view(t.shape)
            view.copy_(t, non_blocking=True)
            return view

        host = [to_host("off", out_off), to_host("rows", out_rows), to_host("centre", fm[:, 1].to(i64)),
                to_host("N", fm[:, 3].contiguous()), to_host("D", fm[:, 4].contiguous()), to_host("props", fm[:, 5:13].contiguous())]
        torch.cuda.current_stream(dev).synchronize()
        res = tuple(x.numpy() for x in host)  # views of the pinned buffers: valid until the next call on this handle
    else:
        res = (np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0), np.zeros(0),
               np.zeros((0, 8)))
    lap("gather_final")
    if timings is not None:
        timings.update({"ms_" + k: round(v, 2) for k, v in phases.items()})
        timings.update(rounds=rounds, local_centres=m_l, shared_centres=nB, stage1_clusters=K_glob, merged=K_merged,
                       merged_local=int(len(gidx_local)) if K_glob else 0, local_rows=int(m.rows), fof_rows=n_g,
                       contests=n_events_glob, launches=launches[0])
    del keep
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


# ------------------------------------------------------------------------------------------------------
# transitive mode (union-find) over slabs: per-GPU forests merged through the communicator
# ------------------------------------------------------------------------------------------------------
def friends_of_friends_slabs(local_rows, global_row, own, comm, handle, linking_length=0.5, max_velocity=None):
    """Connected components of ONE catalogue held as redshift slabs (fofb_transitive_labels per rank + a merge).

    Every rank holds its own redshift interval ``own`` widened by one velocity window (``plan_slab``'s *valid* interval
    is enough; more does no harm): a link needs each galaxy inside the other's window, so every link that touches an
    owned row has both ends on that rank.  The rank labels its rows with the lock-free union-find of K14; a local label
    is the smallest held row of the local component.  Rows held by two ranks tie their two local components together:
    those (row, label) pairs are all-gathered, the label pairs of one row become edges, and a replicated min-label
    propagation with pointer jumping (the same hook-and-jump scheme, on the few boundary labels) runs to its fixpoint.
    Returns (labels int64[n_own] = smallest GLOBAL row of each owned row's component, owned global rows int64[n_own]);
    together over all ranks they equal the single-GPU ``fofb_transitive_labels`` of the whole table."""
    import torch
    h = handle
    dev = torch.device("cuda", h.device)
    i64 = torch.int64
    p = _lib.default_params(H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0,
                            max_velocity=as_float(P.max_velocity if max_velocity is None else max_velocity, "km/s"))
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64) if not hasattr(local_rows, "data_ptr") else local_rows
    m, keep = _lib.matrix_of(local_rows)
    n_loc = int(m.rows)
    lab32 = torch.empty(max(n_loc, 1), dtype=torch.int32, device=dev)[:n_loc]
    g, l = C.c_int64(), C.c_int64(-1)
    h.check(h.lib.fofb_transitive_labels(h.h, C.byref(p), C.byref(m), float(as_float(linking_length)), lab32.data_ptr(), 1,
                                         C.byref(g), None))
    del keep
    grow = global_row.to(dev) if hasattr(global_row, "data_ptr") else torch.from_numpy(np.asarray(global_row, dtype=np.int64)).to(dev)
    z = local_rows[:, 2] if hasattr(local_rows, "data_ptr") else torch.from_numpy(local_rows[:, 2].copy()).to(dev)
    glab = grow[lab32.long()]                       # local root -> its global row (rows are held in global row order)
    owned = (z >= own[0]) & (z < own[1])
    if comm.world == 1:
        return glab[owned].cpu().numpy(), grow[owned].cpu().numpy()
    # rows another rank may hold as well: everything outside the own interval, and the own rows within two windows of
    # the interval's ends (a neighbour holds at most its own interval widened twice, see plan_slab)
    beta = float(p.max_velocity) / C_KMS
    lo2 = _widen(*_widen(own[0], own[0], beta), beta)[0]
    hi2 = _widen(*_widen(own[1], own[1], beta), beta)[1]
    lo_edge = own[0] + (own[0] - lo2) * 1.05 if np.isfinite(own[0]) else own[0]
    hi_edge = own[1] - (hi2 - own[1]) * 1.05 if np.isfinite(own[1]) else own[1]
    shared = (~owned) | (z <= lo_edge) | (z >= hi_edge)
    pairs = torch.stack([grow[shared], glab[shared]], dim=1)
    allp = torch.cat(comm.allgatherv_t(pairs), dim=0)
    # edges between the labels different ranks gave to one row
    order = torch.argsort(allp[:, 0], stable=True)
    rows_s, labs_s = allp[order, 0], allp[order, 1]
    same = rows_s[1:] == rows_s[:-1]
    ea, eb = labs_s[:-1][same], labs_s[1:][same]
    # dense ids for the labels that occur, then min-label hooking + pointer jumping to the fixpoint
    uniq, inv = torch.unique(torch.cat([ea, eb, glab]), return_inverse=True)
    ia, ib, il = inv[: len(ea)], inv[len(ea): 2 * len(ea)], inv[2 * len(ea):]
    parent = torch.arange(len(uniq), device=dev, dtype=i64)   # uniq is sorted: a smaller id is a smaller global row
    while len(ea):
        pa, pb = parent[ia], parent[ib]
        lo_p = torch.minimum(pa, pb)
        before = parent.clone()
        parent.scatter_reduce_(0, pa, lo_p, reduce="amin")
        parent.scatter_reduce_(0, pb, lo_p, reduce="amin")
        parent = parent[parent]                                  # pointer jumping
        parent = parent[parent]
        if torch.equal(parent, before):
            break
    final = uniq[parent[il]]
    return final[owned].cpu().numpy(), grow[owned].cpu().numpy()
