#!/usr/bin/env python
"""Benchmark of the FoF cluster-finding hot path (BASELINE.json: galaxies/s at 10M galaxies per B200).

One step = one pass of the whole path over one synthetic catalogue:
candidate_center_search -> z cuts -> FoF (stages 1-4) -> interloper_removal -> find_masses("virial")
(pipeline.py:62-64,80-99,113,173), i.e. fofb_pipeline_begin + the overdensity sample points drawn from
numpy's legacy generator + fofb_pipeline_finish.

  value : catalogue resident in HBM before the timed region (device pointer passed to the C ABI);
  e2e   : the same call with the catalogue in pinned HOST memory (H2D inside the timed region) plus the
          device->host read of the final cluster lists and properties;
  roofline : the kernel with the largest share of device time in the timed region, timed with CUDA events on
          the library's stream; algorithmic bytes per unit are the DESIGN.md figures;
  cpu_baseline : the oracle in literal mode (the reference's own O(n^2) control flow), 1 core, on a bounded sample.

N > 1 (torchrun): ONE catalogue of --galaxies rows (the N = 1 workload: strong scaling) is split into redshift slabs
with window-wide halos; boundary centres, the tracker state of every priority round and the cluster lists travel over
NCCL (fof_b200/distributed.py); the time is the max over ranks.  `--decomposition fields` runs N independent replicas
instead.  Every line also carries "wide_field": the 100M-galaxy configuration (BASELINE.json configs[3]) on the same
GPUs, and at N = 1 "cpu_baseline_same_config": GPU and literal CPU path on the same 100k table (configs[0]).
`--impl reference` times the CPU implementation (oracle port) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fof_galaxies_per_sec"
UNIT = "galaxies/s"
ZCUT = (0.5, 2.52, 2.5)  # pipeline.py:80-86 with params.py:19-20

# Algorithmic bytes per unit for each instrumented kernel (DESIGN.md section "Kernels"); unit counts are
# filled in from the run's statistics.  Only HBM-shaped kernels have a meaningful figure.
ALGO_BYTES = {
    "k_count": ("galaxy", 16 + 16 + 24 + 4 + 4),  # packed tile entry once, query record, fp64 ra/dec/cos, row R + N W
    "k_count_prep": ("galaxy", 8 + 16 + 16 + 8),  # z, (ra,dec) R + record and theta W
    "k_cell_gather": ("galaxy", 4 + 4 + 16 + 32 + 16),
    "k_bbox_blocks": ("galaxy", 16 + 64),
    "k_extract": ("galaxy", 24 + 28),
    "k_build_rows": ("galaxy", 64 + 64 + 8),
    "k_kill_targets": ("galaxy", 8 + 4),
    "k_ready": ("candidate_centre", 36),
    "k_ready_r1": ("candidate_centre", 36),
    "k_ready_warp": ("candidate_centre", 36),
    "k_hop2_front": ("stage1_member", 1 + 4),
    "k_centre_prep": ("candidate_centre", 64 + 96),
    "k_virial_count": ("stage1_member", 28),
    "k_virial_fill": ("stage1_member", 28 + 8),
    "k_hop1": ("stage1_member", 8 + 24 + 8),
    "k_hop2": ("stage1_member", 8 + 8 + 64 + 4),
    "k_od_count": ("overdensity_point", 4 + 16 + 16),  # sorted key + 16-byte record R, two 8-byte accumulator updates
    "k_member_vr": ("final_member", 24 + 16),
    "k_clip_bins": ("final_member", 8 + 4 + 4),
    "k_gather_members": ("final_member", 4 + 128),
    "k_virial_mass": ("final_member", 56 + 4),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smmax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_literal_sample(n_sample, seed, workers=1):
    """The oracle in literal mode (reference control flow) on an n_sample-galaxy mock; returns seconds."""
    from fof_b200.mock import make_catalog
    from oracle import fof_oracle as O
    arr = make_catalog(n_sample, seed=seed, as_frame=False)
    p = O.Params()
    t0 = time.perf_counter()
    if workers > 1:
        import multiprocessing as mp
        chunks = np.array_split(np.arange(len(arr)), workers * 4)
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.starmap(_count_rows, [(arr, c) for c in chunks])
        N = np.zeros(len(arr))
        for c, v in zip(chunks, parts):
            N[c] = v
        main_arr = np.column_stack([arr, N])
        cand = main_arr[main_arr[:, -1] >= p.min_count]
        cand = cand[cand[:, -1].argsort(kind="stable")[::-1]]
    else:
        main_arr, cand = O.candidate_center_search(arr, p, mode="literal")
    g = main_arr[(main_arr[:, 2] >= ZCUT[0]) & (main_arr[:, 2] <= ZCUT[1])]
    c = cand[(cand[:, 2] >= ZCUT[0]) & (cand[:, 2] <= ZCUT[2])]
    np.random.seed(1234)
    clusters = O.FoF(g, c, p, mode="literal")
    cleaned, _ = O.interloper_removal(clusters, p)
    O.find_masses(cleaned)
    return time.perf_counter() - t0, len(cleaned)


def cpu_literal_table(arr):
    """The oracle in literal mode, one core, on the given table; returns (seconds, final clusters)."""
    from oracle import fof_oracle as O
    p = O.Params()
    t0 = time.perf_counter()
    main_arr, cand = O.candidate_center_search(arr, p, mode="literal")
    g = main_arr[(main_arr[:, 2] >= ZCUT[0]) & (main_arr[:, 2] <= ZCUT[1])]
    c = cand[(cand[:, 2] >= ZCUT[0]) & (cand[:, 2] <= ZCUT[2])]
    np.random.seed(1234)
    clusters = O.FoF(g, c, p, mode="literal")
    cleaned, _ = O.interloper_removal(clusters, p)
    O.find_masses(cleaned)
    return time.perf_counter() - t0, len(cleaned)


def _count_rows(arr, rows):
    from oracle import fof_oracle as O
    return O.number_counts_literal(arr, O.Params(), rows=rows)[rows]


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the path (oracle port; the reference itself needs astropy and
    GriSPy, absent from this image) on the host cores.  Rank 0 alone runs it."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_literal_sample(n_sample, 777 + i, workers=cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_sample / (ms / 1e3)
    sample = (f"oracle literal mode (per-centre full-catalogue masks and grid rebuilds, the reference's control flow) on a "
              f"{n_sample}-galaxy COSMOS-like mock per step; stage 0 spread over {cores} processes, later stages serial")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.galaxies}-galaxy COSMOS-like mock per GPU (this arm: bounded {n_sample}-galaxy sample, "
                                   "the literal path is O(n^2))", "params": "params.py defaults"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _roofline_from_profile(prof, steps, units, peak, peak_kind):
    """Roofline object for the kernel with the largest share of the event-timed device time."""
    if not prof:
        return None
    tot_ms = sum(v[0] for v in prof.values())
    share = {k: round(v[0] / tot_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]}
    name, (kms, cnt) = max(prof.items(), key=lambda kv: kv[1][0])
    unit_name, bytes_per_unit = ALGO_BYTES.get(name, ("galaxy", 0))
    launches_per_step = cnt / max(steps, 1)
    bytes_per_launch = bytes_per_unit * units[unit_name] / max(launches_per_step, 1)
    avg_ms = kms / cnt
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name)
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_kind": f"of {peak_kind}", "avg_launch_ms": avg_ms, "launches_per_step": launches_per_step,
            "algorithmic_bytes_per_launch": bytes_per_launch, "bytes_per_unit": bytes_per_unit, "unit_of_work": unit_name,
            "share_of_kernel_time": share, "kernel_ms_per_step": {k: round(v[0] / max(steps, 1), 4) for k, v in
                                                                  sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]}}


def _path_bytes(n, m, links, members):
    """algorithmic bytes of the whole path = 256 n + 36 m + 4 L + 60 M."""
    return 256.0 * n + 36.0 * m + 4.0 * links + 60.0 * members


def slab_inputs(arr_or_dev, world, rank, max_velocity):
    """This rank's rows of ONE table that every rank holds (numpy on the host, or a torch CUDA tensor sorted by z)."""
    import torch
    from fof_b200 import distributed as D
    if hasattr(arr_or_dev, "data_ptr"):
        t = arr_or_dev
        n = t.shape[0]
        z = t[:, 2]
        step = max(1, n // 2_000_000)                       # the table is sorted by redshift: a strided sample of z
        edges = D.slab_edges_balanced(z[::step].cpu().numpy(), world, max_velocity)
        own, valid, local = D.plan_slab(edges, rank, max_velocity)
        rows = torch.nonzero((z >= local[0]) & (z <= local[1])).flatten()
        return t[rows].contiguous(), rows, own, valid
    edges = D.slab_edges_balanced(arr_or_dev[::max(1, len(arr_or_dev) // 2_000_000), 2], world, max_velocity)
    rows, own, valid = D.select_local_rows(arr_or_dev[:, 2], edges, rank, max_velocity)
    return np.ascontiguousarray(arr_or_dev[rows]), rows, own, valid


def time_slabs(dist, comm, h, p, local, rows, own, valid, steps, warmup, tim, gather=True):
    """max-over-ranks seconds per step of find_clusters_slabs on (local, rows); returns (seconds, last result)."""
    import torch
    from fof_b200 import distributed as D

    def step():
        np.random.seed(4242)  # every rank must hold the same generator state
        return D.find_clusters_slabs(local, rows, own, valid, comm, h, params=p, zcut=ZCUT, timings=tim, gather=gather)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    res = None
    for _ in range(warmup):
        res = step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = step()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / max(steps, 1), res


def run_slabs(args, rank, world, local_rank, h, p, dist):
    """N > 1: ONE catalogue of args.galaxies rows (the N = 1 workload: strong scaling) split into redshift slabs with
    window-wide halos; candidate centres, the per-round tracker state and the cluster lists travel over NCCL
    (fof_b200/distributed.py).  With --wide-galaxies the 100M wide-field configuration (BASELINE.json configs[3]) is
    timed afterwards on the same ranks and reported under "wide_field"."""
    import torch
    from fof_b200 import distributed as D
    from fof_b200.mock import make_catalog, make_catalog_torch
    n_total = args.galaxies
    arr = make_catalog(n_total, seed=20260103, as_frame=False)  # every rank builds the same table, keeps its slab
    local, rows, own, valid = slab_inputs(arr, world, rank, p.max_velocity)
    del arr
    host = torch.empty(local.shape, dtype=torch.float64, pin_memory=True)
    host.numpy()[:] = local
    local_dev = host.cuda()   # value: the rank's rows resident in HBM
    rows_dev = torch.from_numpy(rows).cuda()
    comm = D.TorchComm()
    tim, tim_e = {}, {}
    clocks = ClockSampler(local_rank)
    # warm-up outside the clock sampling, then the timed region
    time_slabs(dist, comm, h, p, local_dev, rows_dev, own, valid, 0, args.warmup, tim, gather=False)
    clocks.start()
    h.profile(2)
    dt, res = time_slabs(dist, comm, h, p, local_dev, rows_dev, own, valid, args.steps, 0, tim, gather=False)
    h.profile(0)
    prof = h.profile_read()
    clock_info = clocks.stop()
    launches = int(tim.get("launches", 0))
    # end to end: the rank's rows start in pinned HOST memory, the global catalogue ends in host arrays on every rank
    dt_e, res_e = time_slabs(dist, comm, h, p, host.numpy(), rows_dev, own, valid, args.steps, 1, tim_e)
    h2d = torch.tensor([float(host.numpy().nbytes)], dtype=torch.float64, device="cuda")
    dist.all_reduce(h2d)
    d2h = int(sum(np.asarray(x).nbytes for x in res_e[:6]))
    st = h.pipeline_stats()
    red = torch.tensor([float(st["candidate_centres"]), float(st["stage1_members"]), float(tim.get("local_rows", 0))],
                       dtype=torch.float64, device="cuda")
    dist.all_reduce(red)
    by_rank = [None] * world
    dist.all_gather_object(by_rank, {k: v for k, v in tim.items() if k.startswith("ms_")})
    wide = None
    if args.wide_galaxies > 0:
        del local_dev, host
        torch.cuda.empty_cache()
        wide = run_wide(args, rank, world, h, p, dist, comm)
    if rank == 0:
        peak, peak_kind = _peaks()
        members_final = int(res_e[0][-1])
        units = {"galaxy": tim.get("local_rows", 0), "candidate_centre": st["candidate_centres"], "stage1_member": st["stage1_members"],
                 "overdensity_point": tim.get("merged_local", 0) * p.n_random, "final_member": members_final / world}
        roofline = _roofline_from_profile(prof, args.steps, units, peak, peak_kind)
        if roofline:
            roofline["scope"] = "rank 0's kernels on its slab (local rows incl. halos)"
        path_gbs = _path_bytes(n_total, red[0].item(), red[1].item(), members_final) / dt / 1e9
        value = n_total / dt
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"ONE {n_total}-galaxy COSMOS-like mock (the N=1 workload) in {world} redshift slabs, whole path: "
                                       "candidate search, FoF, interloper removal, virial masses",
                           "params": f"params.py defaults, linking_length_factor={args.linking_length_factor}",
                           "parallelism": f"{world} redshift slabs with window-wide halos; NCCL all-gather of boundary centres, cluster "
                                          "lists and final catalogue, one all-reduce of the tracker state per priority round",
                           "l2": "per-rank inputs larger than the 126 MB L2; no explicit flush",
                           "phase_ms_by_rank": {k: [r[k] for r in by_rank] for k in by_rank[0]},
                           "rows_held_by_all_ranks": int(red[2].item()), "halo_overhead": round(red[2].item() / n_total - 1.0, 4),
                           "clusters_final": int(len(res_e[0]) - 1), "members_final": members_final,
                           "value_output": "the final catalogue stays sharded (every rank holds the clusters it owns, in host arrays); "
                                           "e2e gathers the global catalogue on every rank",
                           **{k: (int(v) if float(v).is_integer() else v) for k, v in tim.items()}},
                "clocks": clock_info,
                "e2e": {"value": n_total / dt_e, "unit": UNIT, "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": 1e3 * dt_e,
                        "note": "each rank's rows start in pinned host memory; the assembled global catalogue is read back to host "
                                "arrays on every rank"},
                "gpu_launches": launches * args.steps * world, "roofline": roofline,
                "path_hbm": {"algorithmic_GBps": path_gbs, "frac_of_peak_per_gpu": path_gbs / (peak * world),
                             "note": "SURVEY 8(d): (256 n + 36 m + 4 L + 60 M) bytes / step time, against N x the measured copy peak"},
                "cpu_baseline": None, "wide_field": wide}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_wide(args, rank, world, h, p, dist, comm):
    """BASELINE.json configs[3]: the 100M-galaxy wide-field mock (RA 100-150, Dec -20..20) on `world` GPUs.  The table is
    drawn on the GPU by every rank from one seed (identical everywhere), each rank keeps its slab."""
    import torch
    from fof_b200.mock import make_catalog_torch
    n = args.wide_galaxies
    steps, warmup = max(1, min(args.steps, 3)), 2
    table = make_catalog_torch(n, seed=20260104, kind="wide", device=f"cuda:{torch.cuda.current_device()}")
    tim = {}
    if world > 1:
        local, rows, own, valid = slab_inputs(table, world, rank, p.max_velocity)
        del table
        torch.cuda.empty_cache()
        dt, res = time_slabs(dist, comm, h, p, local, rows, own, valid, steps, warmup, tim, gather=False)
        km = torch.tensor([float(len(res[0]) - 1), float(res[0][-1])], dtype=torch.float64, device="cuda")
        dist.all_reduce(km)   # the catalogue stays sharded: every rank holds the clusters it owns
        K, M = int(km[0].item()), int(km[1].item())
    else:
        def step():
            np.random.seed(4242)
            return h.pipeline_run_np_random(table, p, *ZCUT)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            K, M = step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    st = h.pipeline_stats()
    red = torch.tensor([float(st["candidate_centres"]), float(st["stage1_members"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red)
    by_rank = None
    if world > 1:
        by_rank = [None] * world
        dist.all_gather_object(by_rank, {k: v for k, v in tim.items() if k.startswith("ms_")})
        tim = dict(tim, phase_ms_by_rank={k: [r[k] for r in by_rank] for k in by_rank[0]})
    peak, _ = _peaks()
    gbs = _path_bytes(n, red[0].item(), red[1].item(), M) / dt / 1e9
    return {"workload": f"ONE {n}-galaxy wide-field mock (RA 100-150, Dec -20..20, 0.5<z<2, 20% in Plummer clusters; drawn on the "
                        f"GPU with torch, seed 20260104), whole path, {world} GPU(s)" + (", redshift slabs over NCCL" if world > 1 else ""),
            "value": n / dt, "unit": UNIT, "ms_per_step": 1e3 * dt, "steps": steps, "warmup": warmup, "resident": True,
            "clusters_final": K, "members_final": M, "candidate_centres": int(red[0].item()), "stage1_members": int(red[1].item()),
            "path_hbm_GBps": gbs, "path_hbm_frac_of_peak_per_gpu": gbs / (peak * world),
            **{k: (v if isinstance(v, dict) else (int(v) if float(v).is_integer() else v)) for k, v in tim.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--galaxies", type=int, default=10_000_000, help="galaxies per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--linking-length-factor", type=float, default=0.2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--same-config-galaxies", type=int, default=100_000,
                    help="N = 1: size of the table on which the GPU path (end to end) and the literal CPU path are both timed")
    ap.add_argument("--decomposition", default="slabs", choices=["fields", "slabs"],
                    help="N > 1: ONE catalogue of --galaxies rows in redshift slabs with halos (default: the metric's strong-scaling "
                         "line) or an independent field of --galaxies rows per GPU (replicas, no data-path collective)")
    ap.add_argument("--wide-galaxies", type=int, default=100_000_000,
                    help="also time the wide-field configuration (BASELINE.json configs[3]) of this many rows; 0 skips it")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    assert args.warmup >= 3 or args.steps == 0, "timing rules: at least 3 warm-up steps"

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from fof_b200 import _lib
    from fof_b200.mock import make_catalog
    if not os.path.exists(_lib.LIB_PATH):
        if rank == 0:
            import __graft_entry__ as ge
            ge.build()
        if world > 1:
            dist.barrier()
    h = _lib.Handle(local_rank)
    p = _lib.default_params(linking_length_factor=args.linking_length_factor)

    n = args.galaxies
    if args.decomposition == "slabs" and world > 1:
        return run_slabs(args, rank, world, local_rank, h, p, dist)
    arr = make_catalog(n, seed=20260103 + rank, as_frame=False)
    # host table laid out as the reference's `df.values` is: a consolidated pandas float64 block, every column contiguous
    # (shape (n, 7), strides (8, 8 n)).  The library uploads the three columns stage 0 reads first and streams the other
    # four behind the stage-0 kernels.
    host = torch.empty((7, n), dtype=torch.float64, pin_memory=True)
    host.numpy()[:] = arr.T
    dev = torch.from_numpy(arr).cuda()
    del arr
    host_np = host.numpy().T

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stats = {}

    def step(resident):
        # one C-ABI call; the overdensity points come from np.random's MT19937 state, advanced on the device
        Kf, total = h.pipeline_run_np_random(dev if resident else host_np, p, *ZCUT)
        K = h.pipeline_stats()["stage4_clusters"]
        ms, launches = h.last_timing()
        out = None
        if not resident:
            out = h.pipeline_fetch(Kf, total)
        stats.update(K=K, Kf=Kf, total=total, launches=launches, device_ms=ms, u_bytes=2 * 624 * 4)
        return out

    def timed(resident, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step(resident)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    np.random.seed(99 + rank)
    for _ in range(args.warmup):
        step(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    h.profile(2)
    dt = timed(True, args.steps)
    h.profile(0)
    prof = h.profile_read()
    clock_info = clocks.stop()
    ms_per_step = 1e3 * dt / max(args.steps, 1)
    value = n * world / (dt / max(args.steps, 1))
    launches = stats["launches"] * args.steps
    pstats = h.pipeline_stats()

    # end to end: pinned host input, results read back
    step(False)
    dt_e = timed(False, args.steps)
    e2e_value = n * world / (dt_e / max(args.steps, 1))
    d2h = 8 * (stats["Kf"] + 1) + 4 * stats["total"] + 4 * stats["Kf"] + 8 * stats["Kf"] * 10
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_np.nbytes + stats["u_bytes"]),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt_e / max(args.steps, 1)}

    # roofline of the dominant kernel (largest share of the event-timed device time in the timed region)
    peak, peak_kind = _peaks()
    units = {"galaxy": n, "candidate_centre": pstats["candidate_centres"], "stage1_member": pstats["stage1_members"],
             "overdensity_point": stats["K"] * p.n_random, "final_member": stats["total"]}
    roofline = _roofline_from_profile(prof, args.steps, units, peak, peak_kind)
    # the HBM-shaped hot kernel of the path (the N(0.5) link test) is reported beside the dominant kernel
    if prof and "k_count" in prof and roofline is not None:
        kms, cnt = prof["k_count"]
        b = ALGO_BYTES["k_count"][1] * n / max(cnt / max(args.steps, 1), 1)
        tr = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get("k_count")
        roofline["link_test_kernel"] = {"kernel": "k_count", "bound": "hbm", "achieved": b / (kms / cnt * 1e-3) / 1e9, "peak": peak,
                                        "unit": "GB/s", "frac": b / (kms / cnt * 1e-3) / 1e9 / peak, "traffic": tr,
                                        "avg_launch_ms": kms / cnt, "algorithmic_bytes_per_launch": b}
    path_gbs = _path_bytes(n, pstats["candidate_centres"], pstats["stage1_members"], stats["total"]) / (ms_per_step * 1e-3) / 1e9
    path_hbm = {"algorithmic_GBps": path_gbs, "frac_of_peak_per_gpu": path_gbs / peak,
                "note": "SURVEY 8(d): (256 n + 36 m + 4 L + 60 M) bytes / step time, against the measured copy peak"}

    # BASELINE.json configs[0] (100k galaxies) on both sides: the GPU path end to end from host rows, and -- below -- the
    # literal CPU path on the very same table, so that one ratio in the record compares like with like
    same = None
    if rank == 0 and world == 1 and args.same_config_galaxies > 0 and not args.no_cpu_baseline:
        small = make_catalog(args.same_config_galaxies, seed=20260101, as_frame=False)
        for _ in range(3):
            np.random.seed(1234)
            Kf_s, tot_s = h.pipeline_run_np_random(small, p, *ZCUT)
            h.pipeline_fetch(Kf_s, tot_s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            np.random.seed(1234)
            Kf_s, tot_s = h.pipeline_run_np_random(small, p, *ZCUT)
            h.pipeline_fetch(Kf_s, tot_s)
        torch.cuda.synchronize()
        t_gpu = (time.perf_counter() - t0) / 5
        t_cpu, k_cpu = cpu_literal_table(small)
        same = {"workload": f"BASELINE.json configs[0]: {args.same_config_galaxies}-galaxy COSMOS-like mock (seed 20260101), default "
                            "params.py, whole path, host rows in / host catalogue out on both sides",
                "gpu_e2e": {"value": args.same_config_galaxies / t_gpu, "unit": UNIT, "ms": 1e3 * t_gpu, "clusters": int(Kf_s)},
                "cpu": {"value": args.same_config_galaxies / t_cpu, "unit": UNIT, "seconds": t_cpu, "cores": 1, "kind": "port",
                        "clusters": int(k_cpu), "what": "oracle literal mode = the reference's control flow, one Python process"},
                "ratio": t_cpu / t_gpu}

    wide = None
    if world == 1 and args.wide_galaxies > 0:
        del dev, host, host_np
        torch.cuda.empty_cache()
        wide = run_wide(args, rank, world, h, p, dist, None)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tcpu, _ = cpu_literal_sample(args.cpu_sample, 777, workers=1)
        cpu_baseline = {"value": args.cpu_sample / tcpu, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"oracle literal mode (the reference's per-centre O(n) masks and grid rebuilds), all stages, on a "
                                  f"{args.cpu_sample}-galaxy COSMOS-like mock from the same generator; {tcpu:.1f} s; the literal "
                                  "path is O(n^2), so its galaxies/s falls further at the full 10M size"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if world > 1 else "strong", "vs_baseline": None,
                "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"{n}-galaxy COSMOS-like mock per GPU (0.5<z<2, 80% uniform + 20% Plummer clusters, "
                                       "50k gal/deg^2), whole path: candidate search, FoF, interloper removal, virial masses",
                           "params": f"params.py defaults, linking_length_factor={args.linking_length_factor}",
                           "parallelism": "one process per GPU, independent fields (no data-path collective)" if world > 1 else "single GPU",
                           "l2": "inputs (560 MB per step) larger than the 126 MB L2; no explicit flush",
                           "e2e_input": "pinned host table, column-major like a pandas block (df.values); value: row-major device table",
                           "clusters_final": stats["Kf"], "members_final": stats["total"], **pstats},
                "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "device_ms_per_step": stats["device_ms"],
                "roofline": roofline, "path_hbm": path_hbm, "cpu_baseline": cpu_baseline, "cpu_baseline_same_config": same,
                "wide_field": wide}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
