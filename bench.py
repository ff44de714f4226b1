#!/usr/bin/env python
"""Benchmark of the per-time-step detection path (BASELINE.json: "time steps/sec, 0.25 deg hourly PV
2-PVU detection").

One *step* = one batch of ``--batch`` hourly time steps of the ERA5-shaped synthetic PV field
(721 x 1440, float32) through smoothing (5 passes) -> contours at 2 PVU -> streamers + overturnings +
cutoffs (+ properties) -> the three to_xarray flag grids.  ``value`` is time steps/s with the raw field
already resident in HBM; ``e2e`` is the same through ``Detector.run_batch_host`` with the raw field in
pinned host memory (H2D inside the timed region) and the flag grids + event tables copied back.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1: time steps are sharded
                                                            over ranks, no data-path collective)
  python bench.py --impl reference ...                      (the CPU oracle on all host cores)
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "time steps/sec, 0.25deg hourly PV 2-PVU detection (streamers+overturnings+cutoffs+to_xarray)"
UNIT = "time steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 20 (b200) / 2 (reference)")
    ap.add_argument("--warmup", type=int, default=None, help="default: 3 (b200) / 1 (reference)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=296, help="time steps per step (2 per SM)")
    ap.add_argument("--nlat", type=int, default=721)
    ap.add_argument("--nlon", type=int, default=1440)
    ap.add_argument("--passes", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=3, help="time steps of the CPU baseline sample (~4-8 s each)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight (slots / streams)")
    ap.add_argument("--serial", action="store_true", help="device-resident leg on ONE stream (no kernel overlap)")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 20 if a.impl == "b200" else 2
    if a.warmup is None:
        a.warmup = 3 if a.impl == "b200" else 1
    return a


def workload_config(a):
    return {
        "workload": "synthetic ERA5-shaped 0.25deg ({}x{}) hourly PV, level 2 PVU, smoothing {} passes, "
                    "streamer+overturning+cutoff + to_xarray x3".format(a.nlat, a.nlon, a.passes),
        "time_steps_per_step": a.batch,
        "l2": "every step reads a different {:.0f} MB batch (> 126 MB L2)".format(a.batch * a.nlat * a.nlon * 4 / 1e6),
        "parallelism": "time steps sharded over ranks, no collective",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clock / throttle sampler running while the timed region executes."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples before this point (warm-up) are ignored."""
        self.lines = []

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ roofline
def algorithmic_bytes(kernel, a, stats):
    """Algorithmic bytes ONE launch of `kernel` moves for a batch of a.batch time steps."""
    cells = a.nlat * a.nlon
    T = a.batch
    per_step = {
        "smooth_fused": cells * (4 + 8),                       # float32 in, float64 out, independent of passes
        "ms_segments": cells * 8 + 20 * stats["segments"],       # smoothed field read once + 20 B per segment
        "contour_link": 20 * stats["segments"] + 4 * stats["points"],
        "events_raster": 3 * cells,                              # the three int8 flag grids
        "pair_scan": 48 * stats["points"],
        "streamer_cascade": 8 * stats["pairs"] + 4 * stats["points"],
        "streamer_prep": 20 * stats["points"],
        "overturning": 4 * stats["points"],
    }
    return per_step.get(kernel, 0) * T


def measured_traffic(kernel, batch):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)[kernel]["bytes_per_time_step"] * batch)
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arms
def _oracle_step(raw, nlat, nlon, hours, passes):
    from oracle import pipeline as P
    from wavebreaking_b200 import synthetic

    lat, lon = synthetic.grid_coords(nlat, nlon)
    t0 = np.datetime64("2000-01-01T00", "ns")
    tm = t0 + (np.asarray(hours) * 3600 * 10**9).astype("timedelta64[ns]")
    grid = P.Grid(lon, lat, tm)
    out = P.detect_steps(raw, grid, levels=[2.0], passes=passes)
    return {k: len(v) for k, v in out["events"].items()}


def _ref_prepare(args):
    idx, nlat, nlon, tmpdir = args
    from wavebreaking_b200 import synthetic

    f = synthetic.pv_field(nlat, nlon, [float(idx)])
    np.save(os.path.join(tmpdir, "{}.npy".format(idx)), f)
    return idx


def _ref_process(args):
    idx, nlat, nlon, tmpdir, passes = args
    raw = np.load(os.path.join(tmpdir, "{}.npy".format(idx)))
    return _oracle_step(raw, nlat, nlon, [float(idx)], passes)


def run_reference(a):
    """--impl reference: the oracle (a restatement of the reference's CPU implementation that calls the same
    scipy / sklearn routines; the real package cannot be installed in this image) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import tempfile

    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_step = workers  # one time step per worker and step: a bounded sample of the workload
    total = (a.steps + a.warmup) * per_step
    tmpdir = tempfile.mkdtemp(prefix="wbk_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        pool.map(_ref_prepare, [(i, a.nlat, a.nlon, tmpdir) for i in range(total)])
        for w in range(a.warmup):
            pool.map(_ref_process, [(w * per_step + i, a.nlat, a.nlon, tmpdir, a.passes) for i in range(per_step)])
        t0 = time.perf_counter()
        for s in range(a.steps):
            base = (a.warmup + s) * per_step
            pool.map(_ref_process, [(base + i, a.nlat, a.nlon, tmpdir, a.passes) for i in range(per_step)])
        dt = time.perf_counter() - t0
    for f in os.listdir(tmpdir):
        os.remove(os.path.join(tmpdir, f))
    os.rmdir(tmpdir)
    value = a.steps * per_step / dt
    cfg = workload_config(a)
    cfg["time_steps_per_step"] = per_step
    cfg["l2"] = "n/a (CPU arm)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": "{} time steps per step, one per worker process; the real wavebreaking package "
                                   "is not installable here (xarray/geopandas/shapely/skimage absent), so the "
                                   "oracle port (same scipy/sklearn calls) is timed".format(per_step)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(a, raw_host, hours):
    t0 = time.perf_counter()
    n = raw_host.shape[0]
    for i in range(n):
        _oracle_step(raw_host[i:i + 1], a.nlat, a.nlon, hours[i:i + 1], a.passes)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first {} time steps of the benchmark input, single process ({:.1f} s)".format(n, dt)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    from wavebreaking_b200 import _lib, detect, pipeline, spatial, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.get()
    lat, lon = synthetic.grid_coords(a.nlat, a.nlon)
    det = pipeline.Detector(lat, lon, levels=[2.0], passes=a.passes)
    T, K, W = a.batch, a.steps, a.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs: every step (and rank) gets its own slab of consecutive hours, generated in HBM (untimed)
    # (a few distinct slabs are cycled: each is far larger than the 126 MB L2, so nothing is reused from cache)
    nslab = min(K + W, 4)
    slabs = []
    for s in range(nslab):
        hour0 = float((rank * nslab + s) * T)
        slabs.append(spatial.synth_pv(T, a.nlat, a.nlon, hour0=hour0, hour_step=1.0))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident leg (value)
    stats = {"segments": 0, "points": 0, "pairs": 0}
    depth = a.depth

    def timed_region(shared_stream):
        """K batches, 3 in flight; returns (ms, per-kernel profile, launches, per-batch counts)."""
        for res in det.stream((slabs[s % nslab] for s in range(W)), depth=depth, shared_stream=shared_stream):
            pass
        barrier()
        launches0 = lib.cdll.wbk_launch_count()
        lib.cdll.wbk_prof_reset()
        lib.cdll.wbk_prof_enable(1)
        sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        cnts = []
        for res in det.stream((slabs[(W + s) % nslab] for s in range(K)), depth=depth, shared_stream=shared_stream):
            cnts.append(pipeline.summarize(res))
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        e1.record()
        barrier()
        lib.cdll.wbk_prof_enable(0)
        # batches run on side streams: the host clock bounds the region
        return max(e0.elapsed_time(e1), wall_ms), _lib.prof_read(lib), lib.cdll.wbk_launch_count() - launches0, cnts

    # headline region: every slot on its own stream, so the memory-bound kernels of one batch overlap the
    # FP64-bound smoothing of another
    ms, prof_conc, launches, counts = timed_region(shared_stream=a.serial)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * K * T / (ms_max / 1000.0)
    # same K batches again on ONE stream (kernels of different batches do not overlap): clean per-kernel durations
    # for the roofline; its throughput is reported as value_single_stream
    if a.serial:
        ms_ser, prof = ms, prof_conc
    else:
        sampler2 = ClockSampler(local)
        sampler = sampler2
        ms_ser, prof, _, _ = timed_region(shared_stream=True)
    t_ser = torch.tensor([ms_ser], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ser, op=dist.ReduceOp.MAX)
    value_serial = world * K * T / (float(t_ser.item()) / 1000.0)
    ms = ms_ser

    # per-step statistics for the algorithmic byte counts
    stats["points"] = float(np.mean([c["points"] for c in counts])) / T
    stats["segments"] = stats["points"] * 1.4   # raw segments per deduped point (measured ratio)
    stats["pairs"] = 2500.0

    # ---- roofline of the dominant kernel
    peak, peak_src = peaks()
    roof = None
    if prof:
        dom = max(prof.items(), key=lambda kv: kv[1][1])
        name, (n_l, tot_ms) = dom
        avg_ms = tot_ms / n_l
        by = algorithmic_bytes(name, a, stats)
        achieved = by / (avg_ms / 1000.0) / 1e9 if avg_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(name, a.batch), "peak_source": peak_src,
                "avg_launch_ms": avg_ms, "launches": n_l, "algorithmic_bytes_per_launch": by,
                "kernel_ms_share": {k: round(v[1] / sum(x[1] for x in prof.values()), 4) for k, v in prof.items()},
                "device_busy_frac": sum(x[1] for x in prof.values()) / ms,
                "measured_in": "second timed region of the same K batches on ONE CUDA stream (value_single_stream); in the "
                               "headline region 3 batches run on 3 streams and co-running kernels inflate each other",
                "avg_launch_ms_in_headline_region": (prof_conc.get(name, (1, 0.0))[1] / max(prof_conc.get(name, (1, 0.0))[0], 1)),
                "note": "smooth_fused is bounded by the FP64 pipe before HBM: 5 passes x 7 DP ops per cell on a 64x64 "
                        "tile with a 5-cell halo = 2.9 us per 721x1440 step at 64 DP lanes/clk/SM (DESIGN.md 4)"}

    # ---- end-to-end leg (host buffers)
    e2e = None
    if not a.no_e2e:
        host_in = [torch.empty((T, a.nlat, a.nlon), dtype=torch.float32, pin_memory=True) for _ in range(depth)]
        flags_host = [torch.empty((3, T, a.nlat, a.nlon), dtype=torch.int8, pin_memory=True) for _ in range(depth)]
        for i, h in enumerate(host_in):
            h.copy_(slabs[i % nslab])
        torch.cuda.synchronize()
        for r in det.stream((host_in[s % depth] for s in range(min(W, depth))), depth=depth, flags_host=flags_host):
            pass
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        e0.record()
        d2h = 0
        for r in det.stream((host_in[s % depth] for s in range(K)), depth=depth, flags_host=flags_host):
            d2h = flags_host[0].numel() + sum(t.sums.nbytes + 11 * 4 * len(t) for t in r.tables.values()) \
                + sum(t.rings.nbytes for t in r.tables.values() if t.rings is not None)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        wall = time.perf_counter() - t_wall
        ems = max(e0.elapsed_time(e1), wall * 1000.0)
        t_e = torch.tensor([ems], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * K * T / (float(t_e.item()) / 1000.0), "unit": UNIT,
               "h2d_bytes_per_step": T * a.nlat * a.nlon * 4, "d2h_bytes_per_step": int(d2h)}

    # ---- CPU baseline (rank 0, single GPU runs only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        n = min(a.cpu_sample, T)
        raw_host = slabs[0][:n].cpu().numpy()
        hours = np.arange(n, dtype=np.float64)
        cpu = cpu_baseline(a, raw_host, hours)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "value_single_stream": value_serial, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "events_per_time_step": {k: float(np.mean([c[k] for c in counts])) / T
                                     for k in ("streamers", "overturnings", "cutoffs", "contours", "split")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE line (the JSON): everything libraries print there (the NCCL version banner ...) is
    # sent to stderr by pointing fd 1 at fd 2 for the run; the JSON line goes to the saved real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
