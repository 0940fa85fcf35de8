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

N > 1 (torchrun): every rank processes its own independent field of the same size (weak scaling, no data-path
collective); the time is the max over ranks, the value the sum of galaxies over ranks.
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
    "k_count": ("galaxy", 28 + 48),          # ra,dec,z-rank 24 R + N 4 W + query record 48 R
    "k_count_prep": ("galaxy", 8 + 16 + 56),   # z, (ra,dec) R + record and theta W
    "k_cell_gather": ("galaxy", 4 + 4 + 16 + 32),
    "k_bbox_blocks": ("galaxy", 16 + 64),
    "k_extract": ("galaxy", 24 + 28),
    "k_build_rows": ("galaxy", 64 + 64 + 8),
    "k_kill_targets": ("galaxy", 8 + 4),
    "k_ready": ("candidate_centre", 36),
    "k_centre_prep": ("candidate_centre", 64 + 96),
    "k_virial_count": ("stage1_member", 28),
    "k_virial_fill": ("stage1_member", 28 + 8),
    "k_hop1": ("stage1_member", 8 + 24 + 8),
    "k_hop2": ("stage1_member", 8 + 8 + 64 + 4),
    "k_od_count": ("overdensity_point", 16 + 4),
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


def run_slabs(args, rank, world, local_rank, h, p, dist):
    """One catalogue of world x galaxies rows, split into redshift slabs with window-wide halos; the per-round tracker
    state, the candidate centres and the cluster lists travel over NCCL (fof_b200/distributed.py)."""
    import torch
    from fof_b200 import distributed as D
    from fof_b200.mock import make_catalog
    n_total = args.galaxies * world
    arr = make_catalog(n_total, seed=20260103, as_frame=False)  # every rank builds the same table, keeps its slab
    edges = D.slab_edges(arr[:, 2], world)
    rows, own, valid = D.select_local_rows(arr[:, 2], edges, rank, p.max_velocity)
    local = np.ascontiguousarray(arr[rows])
    del arr
    local_dev = torch.from_numpy(local).cuda()  # value: the rank's rows resident in HBM
    rows_dev = torch.from_numpy(rows).cuda()
    comm = D.TorchComm()
    tim = {}

    def step():
        np.random.seed(4242)  # every rank must hold the same generator state
        return D.find_clusters_slabs(local_dev, rows_dev, own, valid, comm, h, params=p, zcut=ZCUT, timings=tim)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    clock_info = clocks.stop()
    if rank == 0:
        value = n_total / (dt / max(args.steps, 1))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"ONE {n_total}-galaxy COSMOS-like mock ({args.galaxies} per GPU), whole path",
                           "params": f"params.py defaults, linking_length_factor={args.linking_length_factor}",
                           "parallelism": f"{world} redshift slabs with window-wide halos; NCCL all-gather of centres and cluster lists, "
                                          "all-reduce of the tracker state every priority round",
                           "clusters_final": int(len(res[0]) - 1), "members_final": int(res[0][-1]), **{k: (int(v) if float(v).is_integer() else v) for k, v in tim.items()}},
                "clocks": clock_info,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(res[1].nbytes),
                        "note": "rows resident in HBM; the assembled global catalogue is returned as host arrays on every rank"},
                "gpu_launches": int(h.last_timing()[1]) * args.steps, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


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
    ap.add_argument("--decomposition", default="fields", choices=["fields", "slabs"],
                    help="N > 1: independent fields per GPU (default) or ONE N x galaxies catalogue in redshift slabs with halos")
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
    host = torch.empty((n, 7), dtype=torch.float64, pin_memory=True)
    host.numpy()[:] = arr
    dev = host.cuda(non_blocking=False)
    del arr
    host_np = host.numpy()

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
    roofline = None
    share = {}
    if prof:
        tot_ms = sum(v[0] for v in prof.values())
        share = {k: round(v[0] / tot_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]}
        top = max(prof.items(), key=lambda kv: kv[1][0])
        name, (kms, cnt) = top
        unit_name, bytes_per_unit = ALGO_BYTES.get(name, ("galaxy", 0))
        launches_per_step = cnt / max(args.steps, 1)
        bytes_per_launch = bytes_per_unit * units[unit_name] / max(launches_per_step, 1)
        avg_ms = kms / cnt
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(name)
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_kind": f"of {peak_kind}",
                    "avg_launch_ms": avg_ms, "launches_per_step": launches_per_step,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "bytes_per_unit": bytes_per_unit, "unit_of_work": unit_name,
                    "share_of_kernel_time": share}

    # the one HBM-shaped hot kernel of the path (the N(0.5) link test) is reported beside the dominant kernel
    stream = None
    if prof and "k_count" in prof and roofline is not None:
        kms, cnt = prof["k_count"]
        b = ALGO_BYTES["k_count"][1] * n / max(cnt / max(args.steps, 1), 1)
        tr = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get("k_count")
        stream = {"kernel": "k_count", "bound": "hbm", "achieved": b / (kms / cnt * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                  "frac": b / (kms / cnt * 1e-3) / 1e9 / peak, "traffic": tr, "avg_launch_ms": kms / cnt,
                  "algorithmic_bytes_per_launch": b}
        roofline["link_test_kernel"] = stream

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
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"{n}-galaxy COSMOS-like mock per GPU (0.5<z<2, 80% uniform + 20% Plummer clusters, "
                                       "50k gal/deg^2), whole path: candidate search, FoF, interloper removal, virial masses",
                           "params": f"params.py defaults, linking_length_factor={args.linking_length_factor}",
                           "parallelism": "one process per GPU, independent fields (no data-path collective)" if world > 1 else "single GPU",
                           "l2": "inputs (560 MB per step) larger than the 126 MB L2; no explicit flush",
                           "clusters_final": stats["Kf"], "members_final": stats["total"], **pstats},
                "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "device_ms_per_step": stats["device_ms"],
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
