#!/usr/bin/env python
"""bench.py -- census + SGM + WTA throughput on B200 (BASELINE.json metric: stereo pairs/s and
SGM Mpix*disp/s, % of HBM roofline).

  python bench.py --gpus 1 --steps 20 --warmup 3                      # this repo's CUDA engine
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 3     # CPU arm: OpenMP port of the reference kernels

A step = one pass of the whole hot path (census x2 -> Hamming cost -> 8 SGM sweeps -> WTA) over one
batch of synthetic stereo pairs per GPU.  Default workload = BASELINE.json configs[1]: 1280x720, 128
disparities, 8-path SGM + WTA.  One process per GPU; pairs are independent, so ranks share nothing on
the data path (weak scaling, no collective) -- torch.distributed is used only for the timing barrier
and the max-over-ranks reduction.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (w, h, D, paths, subpix, lrcheck, default pairs per step per GPU, synthetic config id)
    "c1_640x480x64_4path": (640, 480, 64, 4, 0, 0, 16, 1),
    "c2_1280x720x128_8path_wta": (1280, 720, 128, 8, 0, 0, 16, 2),
    "c3_kitti_1242x375x128_4path": (1242, 375, 128, 4, 0, 0, 16, 3),
    "c4_1920x1080x256_8path_subpix_lr": (1920, 1080, 256, 8, 1, 1, 4, 4),
    "c5_3840x2160x256_8path_subpix_lr_single_gpu": (3840, 2160, 256, 8, 1, 1, 1, 5),
}
DEFAULT_WORKLOAD = "c2_1280x720x128_8path_wta"
P1, P2 = 0.01, 0.02  # applications/stereo2/main.cpp:246-247
L2_BYTES = 126e6
_WINDOW = 0  # census window of this run (0 = 9x7, 1 = 11x11, 2 = 16x16), set from --window


def csrc_hash() -> str:
    """sha256 over the sources of the aggregation kernels: ties a committed ncu traffic figure to the code it was measured on"""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "kangaroo_b200", "csrc")
    for f in ("common.cuh", "sgm_step.cuh", "sgm.cu", "sgm_hsweep.cu", "sgm_fused.cu", "census.cu"):
        hsh.update(f.encode())
        hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(wl, window, B, world, materialised=False, generic_hsweep=False):
    """The `config` object of the JSON line -- the same for this repo's arm and the CPU reference arm."""
    w, h, D, paths, subpix, lrcheck, _, _ = WORKLOADS[wl]
    return {"workload": wl, "w": w, "h": h, "disparities": D, "paths": paths, "window": window,
            "popcount": "popc32-compat", "subpix": subpix, "lrcheck": lrcheck, "pairs_per_step_per_gpu": B,
            "sharding": f"pair-batch x{world}, no collective",
            "l2": f"per-step working set {B * w * h * D * 4 / 1e9:.2f} GB (fp32 aggregate) vs "
                  f"{L2_BYTES / 1e6:.0f} MB L2: inputs larger than L2, no flush"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_pairs(w, h, D, cfg, n):
    from kangaroo_b200.synth import stereo_pair
    ls, rs = [], []
    for i in range(n):
        L, R, _ = stereo_pair(w, h, D, config=cfg, index=i % 4)  # 4 distinct pairs, cycled (generation is CPU-bound)
        ls.append(np.roll(L, i // 4, axis=0))
        rs.append(np.roll(R, i // 4, axis=0))
    return np.stack(ls), np.stack(rs)


def cpu_port_throughput(w, h, D, paths, subpix, lrcheck, cfg, rows=None, reps=1):
    """Times the OpenMP port of the reference kernels (oracle/) on `reps` pairs cropped to `rows` rows."""
    import oracle as ko
    from kangaroo_b200.synth import stereo_pair
    ko.use_all_cores()
    L, R, _ = stereo_pair(w, h, D, config=cfg)
    rows = rows or h
    L, R = np.ascontiguousarray(L[:rows]), np.ascontiguousarray(R[:rows])
    t0 = time.perf_counter()
    for _ in range(reps):
        ko.pipeline_u8(L, R, D, window=_WINDOW, dodiag=(paths == 8), subpix=bool(subpix), lrcheck=bool(lrcheck), p1=P1, p2=P2)
    dt = (time.perf_counter() - t0) / reps
    frac = rows / h
    return frac / dt, dt, ko.num_threads(), rows


def run_reference(args, wl):
    """--impl reference: the reference has no CPU implementation (its kernels are CUDA-only), so this arm is
    the OpenMP scalar port of its kernel bodies (oracle/kangaroo_oracle.c) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h, D, paths, subpix, lrcheck, _, cfg = WORKLOADS[wl]
    # bounded sample: full pairs if the run is short, else a row crop so that K+W steps stay within minutes
    total = args.steps + args.warmup
    rows = h if total <= 8 else max(32, int(h * 8 / total))
    for _ in range(args.warmup):
        cpu_port_throughput(w, h, D, paths, subpix, lrcheck, cfg, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, cores, rows = cpu_port_throughput(w, h, D, paths, subpix, lrcheck, cfg, rows)
    dt = time.perf_counter() - t0
    pairs = args.steps * rows / h
    value = pairs / dt
    sample = f"{rows} of {h} rows of one {w}x{h}x{D} pair per step ({paths}-path), {args.steps} steps"
    out = {"impl": "reference", "metric": "stereo_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "mpix_disp_per_s": value * w * h * D / 1e6,
           "config": workload_config(wl, args.window, args.batch or WORKLOADS[wl][6], max(1, int(os.environ.get("WORLD_SIZE", "1")))),
           "note": "reference kernels are CUDA-only; CPU arm = OpenMP scalar port of the kernel bodies (oracle/), one pair at a time",
           "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(_REAL_STDOUT, json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="stereo pairs per step per GPU (0 = workload default)")
    ap.add_argument("--window", default="9x7", choices=["9x7", "11x11", "16x16"],
                    help="census descriptor: 9x7 -> u64 (north_star headline), 16x16 = 8w x 16h -> ulong4 (what the apps run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--materialised-cost", action="store_true",
                    help="A/B: every pass reads the u8 cost volume (no in-sweep cost from census words)")
    ap.add_argument("--generic-hsweep", action="store_true",
                    help="A/B: run the horizontal paths through the generic sweep kernel instead of sgm_hsweep.cu")
    args = ap.parse_args()
    global _WINDOW
    _WINDOW = {"9x7": 0, "11x11": 1, "16x16": 2}[args.window]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = args.workload
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    from kangaroo_b200 import capi, roo
    from kangaroo_b200.sharding import aggregate_throughput, reduce_max   # the code tests/test_sharding_gloo.py covers

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kangaroo_b200 has no CPU path")
    if args.generic_hsweep:
        roo.set_tuning(capi.TUNE_HSWEEP, 0)
    if args.materialised_cost:
        roo.set_tuning(capi.TUNE_INSWEEP_COST, 0)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w, h, D, paths, subpix, lrcheck, dbatch, cfg = WORKLOADS[wl]
    B = args.batch or dbatch
    K, W = args.steps, args.warmup
    Lh, Rh = make_pairs(w, h, D, cfg + 10 * rank, B)
    left = torch.from_numpy(Lh).cuda()
    right = torch.from_numpy(Rh).cuda()
    disp = torch.empty((B, h, w), dtype=torch.float32, device="cuda")
    win = {"9x7": roo.WIN_9x7, "11x11": roo.WIN_11x11, "16x16": roo.WIN_16x16}[args.window]
    eng = roo.StereoEngine(w, h, D, window=win, P1=P1, P2=P2, dohoriz=True, dovert=True, doreverse=True,
                           dodiag=(paths == 8), subpix=bool(subpix), lrcheck=bool(lrcheck), max_batch=B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value) + live per-kernel timing for the roofline ----
    for _ in range(W):
        eng.run_device(left, right, disp)
    barrier()
    eng.set_profiling(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        eng.run_device(left, right, disp)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = capi.launch_count() - launches0
    clk = clocks.stop()
    prof = eng.get_profile()
    eng.set_profiling(False)
    ms_max = reduce_max(ms, device="cuda")                               # slowest rank
    value = aggregate_throughput(B * K, ms * 1e-3, device="cuda")         # all pairs of all ranks / slowest rank's time

    # ---- end to end through the public API with HOST buffers (H2D + compute + D2H every step) ----
    e2e = None
    if not args.no_e2e:
        # two sets of pinned host buffers: a camera thread fills one while the other is in flight
        lp = [torch.from_numpy(Lh).pin_memory() for _ in range(2)]
        rp = [torch.from_numpy(Rh).pin_memory() for _ in range(2)]
        dp = [torch.empty((B, h, w), dtype=torch.float32).pin_memory() for _ in range(2)]
        # one submit = one batch: upload (H2D), the whole path, download (D2H); two batches in flight, so the copies of
        # step k+1 / k-1 overlap the kernels of step k.  Every step's result is read on the host after its wait.
        g = B
        eng_h = roo.StereoEngine(w, h, D, window=win, P1=P1, P2=P2, dodiag=(paths == 8), subpix=bool(subpix),
                                 lrcheck=bool(lrcheck), max_batch=g)

        def run_steps(n):
            prev, acc = None, 0.0
            for k in range(n):
                t = eng_h.submit_host(lp[k & 1], rp[k & 1], dp[k & 1])
                if prev is not None:
                    eng_h.wait(prev)
                    acc += float(dp[(k - 1) & 1][0, h // 2, w // 2])   # host read of the finished step's result
                prev = t
            eng_h.wait(prev)
            acc += float(dp[(n - 1) & 1][0, h // 2, w // 2])
            return acc

        run_steps(W)
        barrier()
        t0 = time.perf_counter()
        run_steps(K)
        dt = time.perf_counter() - t0
        e2e = {"value": aggregate_throughput(B * K, dt, device="cuda"), "unit": "pairs/s", "h2d_bytes_per_step": int(2 * B * w * h),
               "d2h_bytes_per_step": int(B * w * h * 4),
               "api": "roo_engine_submit_host / roo_engine_wait (pinned host buffers, two steps in flight)",
               "pairs_in_flight": 2 * g}
        eng_h.close()

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        unit = float(w) * h * D * B                       # pixel*disparity units one launch processes
        step_kernel_ms = sum(v[0] for k, v in prof.items() if not k.startswith("pass"))
        # ---- the aggregation passes of one step, in plan order (mirrors engine.cu: fused vertical groups when the batch
        # fills the GPU, bulk-copy kernel for the horizontal paths, in-sweep cost where the descriptor allows it)
        S = sum(1 for i in range(8) if prof[f"pass{i}"][1] > 0)
        fused = prof["vgroup"][1] > 0
        if paths == 8:
            names = (["sgm_vgroup_kernel (down + 2 diagonals)", "sgm_vgroup_kernel (up + 2 diagonals)"] if fused else
                     ["sgm_sweep_kernel"] * 6) + ["sgm_hsweep_kernel (right)", "sgm_hsweep_kernel (left, WTA epilogue)"]
        else:
            names = ["sgm_sweep_kernel (down)", "sgm_sweep_kernel (up)", "sgm_hsweep_kernel (right)",
                     "sgm_hsweep_kernel (left, WTA epilogue)"]
        cen_ok = args.window == "9x7" and not args.materialised_cost
        passes = []
        for i in range(S):
            ms_i = prof[f"pass{i}"][0] / max(prof[f"pass{i}"][1], 1)
            nm = names[i] if i < len(names) else "sgm_sweep_kernel"
            # fused groups and the bulk-copy horizontal kernel recompute the cost from census words when the descriptor
            # allows it; the generic single-path sweep only up to 64 disparities (engine.cu)
            in_sweep = cen_ok and ("vgroup" in nm or ("hsweep" in nm and not args.generic_hsweep) or D <= 64)
            alg = (4.0 if i == 0 else 8.0) * unit                      # first pass writes, later passes read + write
            moved = ((0.0 if i == 0 else 4.0) + (0.0 if i == S - 1 else 4.0) + (0.0 if in_sweep else 1.0)) * unit
            passes.append({"pass": i, "kernel": nm, "ms": ms_i, "algorithmic_bytes": alg, "achieved_gbs": alg / ms_i / 1e6,
                           "frac": alg / ms_i / 1e6 / peak, "bytes_this_design_moves": moved,
                           "moved_gbs": moved / ms_i / 1e6, "moved_frac": moved / ms_i / 1e6 / peak,
                           "cost": "in-sweep from census words" if in_sweep else "u8 volume read",
                           "share_of_step": ms_i * K / step_kernel_ms if step_kernel_ms else None})
        dom = max(passes, key=lambda q: q["ms"]) if passes else None
        # measured DRAM traffic of the dominant launch from the committed ncu capture -- only while the kernels are the
        # ones that were profiled (hash of kangaroo_b200/csrc at capture time), else null
        traffic, traffic_note = None, "no ncu capture for this workload"
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if tr.get("csrc_sha256") != csrc_hash():
                traffic_note = "kernel sources changed since the ncu capture (profiles/r2_traffic.json): traffic not reported"
            elif args.materialised_cost or args.generic_hsweep:
                traffic_note = "no ncu capture for this development configuration"
            elif dom and (wl + ("@16x16" if args.window == "16x16" else "")) in tr.get("workloads", {}):
                key = wl + ("@16x16" if args.window == "16x16" else "")      # the capture of this census window
                ent = tr["workloads"][key]["passes"]
                if dom["pass"] < len(ent):
                    traffic = ent[dom["pass"]]["dram_bytes"] * B / tr["workloads"][key]["pairs_per_launch"]
                    traffic_note = f"ncu dram__bytes_read+write, {tr['workloads'][key]['source']}"
        except Exception:
            pass
        agg_ms = sum(q["ms"] for q in passes)
        agg_alg = sum(q["algorithmic_bytes"] for q in passes)
        agg_moved = sum(q["bytes_this_design_moves"] for q in passes)
        out = {
            "metric": "stereo_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mpix_disp_per_s": value * w * h * D / 1e6,
            "config": workload_config(wl, args.window, B, world),
            "matching_cost": ("in-sweep from census words (no cost volume)" if passes and all(q["cost"].startswith("in-sweep") for q in passes)
                              else ("in-sweep in some passes, u8 cost volume in the others" if any(q["cost"].startswith("in-sweep") for q in passes)
                                    else "u8 cost volume")),
            "roofline": ({"bound": "hbm", "kernel": dom["kernel"], "pass": dom["pass"], "achieved": dom["achieved_gbs"],
                          "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic, "traffic_note": traffic_note,
                          "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
                          "avg_launch_ms": dom["ms"], "launches_timed": K, "share_of_step": dom["share_of_step"],
                          "note": "the slowest aggregation launch of the step; every launch is listed in roofline_passes"}
                         if dom else None),
            "roofline_passes": passes,
            "aggregation": {"passes": S, "ms_per_step": agg_ms,
                            "algorithmic_bytes_per_step": agg_alg, "achieved_gbs": agg_alg / agg_ms / 1e6 if agg_ms else None,
                            "frac_of_peak": agg_alg / agg_ms / 1e6 / peak if agg_ms else None,
                            "bytes_this_design_moves_per_step": agg_moved,
                            "moved_gbs": agg_moved / agg_ms / 1e6 if agg_ms else None,
                            "moved_frac_of_peak": agg_moved / agg_ms / 1e6 / peak if agg_ms else None,
                            "note": "algorithmic = SURVEY 8d, 4 B*(2S-1) per pixel*disparity; moved = what these launches "
                                    "read and write (the last sweep writes no aggregate; + 1 B per pass that reads the u8 cost)"},
            "kernel_ms_per_step": {k: v[0] / K for k, v in prof.items() if not k.startswith("pass")},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        # context next to the CPU arm: the reference's OWN CUDA kernels, recompiled unmodified for sm_100a, as measured by
        # tests/test_gpu_reference_speed.py on a B200 (they only launch up to 1024 x 1024 pixels, so not on this workload)
        try:
            rg = json.load(open(os.path.join(ROOT, "profiles", "r2_reference_gpu_speed.json")))
            out["reference_gpu_kernels"] = {
                "source": "profiles/r2_reference_gpu_speed.json (committed record, not measured in this run)",
                **{k: {"reference_ms_per_pair": v["reference_kernels"]["ms_per_pair"],
                       "this_engine_ms_per_pair_batch8": 1e3 / v["engine_batch8"]["pairs_per_s"]}
                   for k, v in rg.items() if isinstance(v, dict) and "reference_kernels" in v}}
        except Exception:
            pass
        if e2e:
            out["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:   # the CPU port is timed at N=1 only (rank 0)
            # bounded sample: whole pairs, about 10-20 s of wall time on all host cores
            _, dt1, _, _ = cpu_port_throughput(w, h, D, paths, subpix, lrcheck, cfg, h)
            reps = max(1, min(8, int(10.0 / dt1)))
            v, dt, cores, rows = cpu_port_throughput(w, h, D, paths, subpix, lrcheck, cfg, h, reps)
            out["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                   "sample": f"{reps} whole {w}x{h}x{D} pair(s), {paths}-path, {dt:.2f} s each "
                                             f"on {cores} threads (after 1 warm-up pair)"}
        _emit(_REAL_STDOUT, json.dumps(out))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _json_only_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when the
    first communicator is created), so everything that goes to fd 1 during the run is sent to stderr and the JSON line is
    written to the real stdout at the end."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_fd: int, line: str) -> None:
    sys.stdout.flush()
    os.write(real_fd, (line + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    _REAL_STDOUT = _json_only_stdout()
    main()
