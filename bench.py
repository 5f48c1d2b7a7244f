#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Workload (config.workload): configs[1] = rcontrib daylight coefficients,
Reinhart MF:1 (145 sky bins), -I+ -ab 3 -ad 4096 -lw 2.44e-4, 100 000 sensors,
seeded synthetic 100k-polygon office (pyradiance_b200/scenegen.py, octree built
by our own builder).  One STEP = the daylight-coefficient matrix of 100k sensors
per GPU, delivered to the gathering rank.

  N = 1   the whole 100k-sensor matrix on one GPU.
  N > 1   one process per GPU (torchrun, NCCL).  Headline = WEAK scaling: every
          rank owns its own 100k-sensor block of an N x 100k-record job (RNG keyed
          by the global record index) and its rows are stored -- by the kernel
          that finishes a batch, over NVLink peer memory (dist.RowWindow) -- into
          the ONE [N x 100k, 145, 3] matrix in rank 0's HBM: the row gather is
          inside the timed region.  `strong` (same JSON line) = the FIXED
          100k-sensor job of N = 1 sharded over the N ranks and gathered on rank 0
          three ways (window / exact-size NCCL send-recv from HBM / per-GPU D2H
          into one shared pinned host matrix), with rank 0's own single-GPU time
          for the same job measured in the same run.

  value   traced rays/s, inputs resident in HBM, matrix gathered in rank 0's HBM
  e2e     same metric from pinned HOST ray arrays to the whole matrix in rank 0's
          HOST memory (H2D + D2H inside the timed region)
  roofline  the dominant kernel (k_trace: octree walk + intersection) against
          the measured HBM copy bandwidth, on ALGORITHMIC bytes
  cpu_baseline  the unmodified reference rcontrib -n <cores> (oracle/_ref) on a
          bounded, stratified sample of the same workload

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1]

--config c1 times BASELINE configs[0] (rtrace -I -ab 0, 10k-sensor grid over the
reference's own tests/Resources/trace.oct) through the public Python call, next to
the reference rtrace on the same input; it prints its own JSON line.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NSENS = int(os.environ.get("RB_BENCH_SENSORS", 100_000))
NPOLY = int(os.environ.get("RB_BENCH_POLYS", 100_000))
AB, AD = 3, 4096
LW = 1.0 / AD
OPTS = ["-ab", str(AB), "-ad", str(AD), "-lw", f"{LW:.4e}"]
RB_P = "MF=1,rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"
RB_ARGS = ["-f", "reinhartb.cal", "-p", RB_P, "-bn", "Nrbins", "-b", "rbin", "-m", "skyglow"]
WORKLOAD = (f"rcontrib -I+ -ab {AB} -ad {AD} -lw {LW:.3e} Reinhart MF:1 (145 bins), {NSENS} sensors, "
            f"synthetic {NPOLY}-polygon office (BASELINE configs[1])")
METRIC, UNIT = "traced_rays_per_sec", "rays/s"
RAYS_SAMPLE = 48          # sensors the oracle port counts rays / node visits on (stratified over the grid)


def config_block(world):
    """The workload description; identical in both arms (the measured figures live outside `config`)."""
    return {"workload": WORKLOAD, "sensors_per_gpu": NSENS, "polygons": NPOLY, "bins": 145,
            "sharding": f"records x{world}, scene replicated" + (", rows gathered on rank 0" if world > 1 else ""),
            "l2": "256 MiB flush buffer written before every step; ray queues (GBs per step) exceed L2; "
                  "scene tables are L2-resident by design",
            "rays_per_sensor": f"counted by the CPU oracle on {RAYS_SAMPLE} stratified sensors (both arms); "
                               "the GPU arm's `value` uses its own device counter"}


def scene_paths(tag=""):
    tmp = Path(os.environ.get("RB_TMP", "/tmp/rb200_bench"))
    tmp.mkdir(parents=True, exist_ok=True)
    return tmp / f"office{NPOLY}{tag}.rad", tmp / f"office{NPOLY}{tag}.oct"


def ensure_scene():
    """GPU arm: scene text by the seeded generator, octree by the library's own builder."""
    from pyradiance_b200 import scenegen
    rad, octf = scene_paths()
    if not octf.exists():
        tmp = octf.with_suffix(f".{os.getpid()}.tmp")
        scenegen.write_office(rad.with_suffix(f".{os.getpid()}.rad"), npolys=NPOLY, seed=1234)
        scenegen.build_octree(rad.with_suffix(f".{os.getpid()}.rad"), tmp)
        os.replace(tmp, octf)
    return octf


def ensure_scene_reference():
    """Reference arm: same seeded scene text, octree by the reference's own `oconv -f` -- nothing
    of librb200.so is loaded (the two builders give byte-identical octrees, tests/test_host.py)."""
    from oracle import refrun
    from pyradiance_b200 import scenegen            # pure-Python generator; does not load the CUDA library
    rad, octf = scene_paths("_ref")
    if not octf.exists():
        scenegen.write_office(rad, npolys=NPOLY, seed=1234)
        tmp = octf.with_suffix(f".{os.getpid()}.tmp")
        refrun.oconv([rad], tmp)
        os.replace(tmp, octf)
    return octf


def stratified(n, k):
    """k of n record indices, evenly spread over the (row-major, jittered) sensor grid."""
    k = int(min(n, max(1, k)))
    return np.unique(((np.arange(k) + 0.5) * n / k).astype(np.int64))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_ = index, [], threading.Event()

    def run(self):
        while not self.stop_.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop_.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def reference_rcontrib_rate(octf, sens, cores, target_s=12.0):
    """Time the unmodified reference rcontrib -n cores on a bounded STRATIFIED sample of the sensors.
    Returns (sensors/s, sample size, seconds)."""
    from oracle import refrun
    n = 256
    while True:
        sub = sens[stratified(len(sens), n)]
        t = time.perf_counter()
        refrun.rcontrib(octf, sub, ["-I+"] + OPTS + ["-y", str(len(sub))] + RB_ARGS, nproc=cores, outform="f")
        dt = time.perf_counter() - t
        if dt > 0.4 * target_s or n >= len(sens) or n >= 65536:
            return len(sub) / dt, len(sub), dt
        n = int(min(len(sens), max(n * 2, n * target_s / max(dt, 1e-3) * 0.8)))


def oracle_counts(octf, sens):
    """Rays per sensor and V / E / P per ray of this workload, counted by the CPU restatement
    (oracle/rb_oracle.c) on RAYS_SAMPLE stratified sensors.  SURVEY 8(d): bytes(ray) =
    32 + 32 V + 4 E + 64 P + 16, plus 24 per contribution."""
    from oracle import port
    s = port.Scene(octf, rcontrib=True, ambounce=AB, ambdiv=AD, minweight=LW, seed=5)
    s.add_modifier("skyglow", port.BIN_REINHARTB, 1, (0, 0, -1), (0, 1, 0), 1.0, 145)
    idx = stratified(len(sens), RAYS_SAMPLE)
    s.rcontrib(sens[idx], irrad=2)
    c = s.counters()
    nr = max(1, c["nrays"])
    V, E, P, K = c["nodes"] / nr, c["leafents"] / nr, c["prims"] / nr, c["contribs"] / nr
    bpr = 32 + 32 * V + 4 * E + 64 * P + 16 + 24 * K
    return {"rays_per_sensor": c["nrays"] / len(idx), "bytes_per_ray": bpr,
            "counts": {"V": round(V, 2), "E": round(E, 2), "P": round(P, 2), "contribs_per_ray": round(K, 4)}}


def file_sha16(path):
    try:
        return hashlib.sha256(Path(path).read_bytes()).hexdigest()[:16]
    except OSError:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refrun
    from pyradiance_b200 import scenegen
    if not refrun.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"})
        return 0
    octf = ensure_scene_reference()
    sens = scenegen.office_sensors(NSENS)
    cores = os.cpu_count() or 1
    oc = oracle_counts(octf, sens)
    rays_per_sensor = oc["rays_per_sensor"]
    rate, n0, dt0 = reference_rcontrib_rate(octf, sens, cores, target_s=6.0)
    n = int(max(64, min(NSENS, rate * 6.0)))          # ~6 s per step
    sub = sens[stratified(NSENS, n)]
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        refrun.rcontrib(octf, sub, ["-I+"] + OPTS + ["-y", str(len(sub))] + RB_ARGS, nproc=cores, outform="f")
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    sps = len(sub) * args.steps / tot
    val = sps * rays_per_sensor
    sample = (f"{len(sub)} of {NSENS} sensors per step (stratified over the grid), reference rcontrib -n {cores} "
              f"(oracle/_ref, unmodified, -O3 -ffast-math; octree by the reference oconv), process start-up and "
              f"scene load included; rays/sensor = {rays_per_sensor:.0f} counted by the oracle port on {RAYS_SAMPLE} sensors")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_block(args.gpus),
        "sensors_per_s": sps, "rays_per_sensor": rays_per_sensor,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    return 0


def run_c1(args):
    """BASELINE configs[0]: rtrace -I -ab 0 over a 100 x 100 sensor grid, tests/Resources/trace.oct,
    through the public call (pyradiance_b200.rtrace: bytes in, bytes out, a fresh context and octree load
    per call, like one reference process per call) next to the reference rtrace -n <cores> / -n 1."""
    import pyradiance_b200 as pr
    from oracle import refrun
    octf = ROOT / "tests" / "golden" / "trace.oct"
    gx, gy = np.meshgrid(np.linspace(1, 39, 100), np.linspace(2, 45, 100))
    grid = np.stack([gx.ravel(), gy.ravel(), np.full(10000, 2.5), np.zeros(10000), np.zeros(10000), np.ones(10000)], 1)
    params = ["-I", "-ab", "0", "-dt", "0", "-dj", "0", "-dc", "1"]
    raw = grid.tobytes()

    def ours():
        return pr.rtrace(raw, str(octf), header=False, inform="d", outform="d", params=params)
    for _ in range(max(3, args.warmup)):
        out = ours()
    ts = []
    for _ in range(max(5, args.steps)):
        t = time.perf_counter(); out = ours(); ts.append(time.perf_counter() - t)
    v = np.frombuffer(out, dtype=np.float64).reshape(-1, 3)
    line = {"metric": "c1_wall_ms", "unit": "ms", "higher_is_better": False, "config": {
        "workload": "rtrace -I -ab 0 -dt 0 -dj 0 -dc 1, 100x100 sensor grid, tests/Resources/trace.oct (BASELINE configs[0])"},
        "value": 1e3 * float(np.median(ts)), "best_ms": 1e3 * min(ts), "rays": 10000, "nonzero_rows": int((v[:, 0] > 0).sum()),
        "what": "pyradiance_b200.rtrace(bytes) -> bytes: new context, octree load + upload, H2D, trace, D2H, per call"}
    if refrun.available():
        cores = os.cpu_count() or 1
        for tag, nproc in (("reference_n1_ms", 1), ("reference_ncores_ms", cores)):
            rs = []
            for _ in range(5):
                t = time.perf_counter()
                ref = refrun.run("rtrace", ["-h", "-fdd", "-n", str(nproc)] + params + [str(octf)], raw)
                rs.append(time.perf_counter() - t)
            line[tag] = 1e3 * float(np.median(rs))
        line["cores"] = cores
        w = np.frombuffer(ref, dtype=np.float64).reshape(-1, 3)
        line["max_rel_err_vs_reference"] = float(np.max(np.abs(v - w) / (np.abs(w) + 1e-9)))
    emit(line)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyradiance_b200 import _lib, scenegen
    from pyradiance_b200 import dist as rbd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyradiance_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        octf = ensure_scene()
    if world > 1:
        dist.barrier()
    octf = ensure_scene()

    ctx = _lib.Context(local, _lib.RB_PROGRAM_RCONTRIB)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    t_load = time.perf_counter()
    ctx.load_octree(octf)                  # read + flatten + cell table + upload: once per scene, outside the timed steps
    t_load = time.perf_counter() - t_load
    ctx.set_options(OPTS)
    ctx.cal_load("reinhartb.cal")
    ctx.cal_set(RB_P)
    ctx.add_modifier("skyglow", RB_P, "rbin", int(ctx.cal_eval("Nrbins") + .5))
    ncols = ctx.num_columns()
    flags = _lib.RB_IRRAD_RCONTRIB
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > L2 (126 MB)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA events on the launching stream around `steps` calls, barrier + synchronize on both
        sides, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---------------- headline: weak scaling, rows gathered in rank 0's HBM ----------------
    # rank r owns the records [r*NSENS, (r+1)*NSENS) of an N*NSENS-sensor job
    sens = scenegen.office_sensors(NSENS, seed=42 + rank)
    nrec_all = NSENS * world
    d_rays = torch.from_numpy(sens).to("cuda")
    h_rays = np.ascontiguousarray(sens)
    ctx.pin(h_rays)
    if world > 1:
        win = rbd.RowWindow(ctx, nrec_all, ncols)                  # the one matrix, in rank 0's HBM
        out_ptr, out_floats = win.ptr(rank * NSENS), NSENS * ncols * 3
        host = rbd.SharedHostMatrix(ctx, nrec_all, ncols)          # the one matrix, in pinned host memory
        h_mine = host.array[rank * NSENS:(rank + 1) * NSENS]
    else:
        d_out = torch.empty((NSENS, ncols, 3), dtype=torch.float32, device="cuda")
        out_ptr, out_floats = d_out.data_ptr(), d_out.numel()
        h_mine = np.empty((NSENS, ncols, 3), dtype=np.float32)
        ctx.pin(h_mine)
    row_base = rank * NSENS

    def step_device():
        flush.fill_(1)                                                     # evict L2 between steps
        ctx.rcontrib_device(d_rays.data_ptr(), NSENS, 1, flags, row_base, out_ptr, out_floats)
        if world > 1:
            dist.barrier()                 # every rank's rows are in rank 0's HBM: the step (= the matrix) is complete

    def step_host():
        flush.fill_(1)
        ctx.rcontrib(h_rays, accum=1, flags=flags, row_base=row_base, out=h_mine)
        if world > 1:
            dist.barrier()                 # every rank's rows are in the shared pinned host matrix rank 0 reads
        return float(h_mine[0, 0, 0])

    for _ in range(args.warmup):
        step_device()
    ctx.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_device, args.steps)
    st = ctx.stats()
    for _ in range(args.warmup):           # the host path has first-call costs of its own (staging buffers for rays and rows)
        step_host()
    ctx.reset_stats()
    ms_e2e = timed(step_host, args.steps)
    st_e2e = ctx.stats()
    sampler.stop_.set()
    sampler.join(timeout=2)
    if world > 1:
        if rank == 0:
            full = np.empty((nrec_all, ncols, 3), dtype=np.float32)
            win.download(full)
            assert np.isfinite(full).all() and all(full[r * NSENS:(r + 1) * NSENS].sum() > 0 for r in range(world))
            np.testing.assert_allclose(full.sum(dtype=np.float64), host.array.sum(dtype=np.float64), rtol=1e-3)
            del full
    else:
        assert np.isfinite(h_mine).all() and h_mine.sum() > 0
        np.testing.assert_allclose(h_mine.sum(), float(d_out.sum().item()), rtol=1e-3)   # same job, same seeds

    rays_local = st["nrays"]
    tot = torch.tensor([float(rays_local)], device="cuda", dtype=torch.float64)
    per_rank = torch.tensor([st["kernel_ms"] / args.steps, st["wave_ms"] / args.steps, float(st["launches"]) / args.steps],
                            device="cuda", dtype=torch.float64)
    per_rank_all = [per_rank.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_gather(per_rank_all, per_rank)
    per_rank_all = [[float(x) for x in t.tolist()] for t in per_rank_all]
    rays_all = float(tot.item())
    value = rays_all / (ms / 1e3)
    e2e_value = rays_all / (ms_e2e / 1e3)

    # ---------------- strong scaling: the fixed NSENS-sensor job of N = 1 over N ranks ----------------
    strong = None
    if world > 1:
        sens0 = scenegen.office_sensors(NSENS, seed=42)            # the N = 1 job
        mine, r0, r1 = rbd.local_rays(sens0, 1, rank, world)
        d_mine = torch.from_numpy(np.ascontiguousarray(mine)).to("cuda")
        h_in = mine.copy()                 # (its own buffer: pinning a slice of sens0 would leave sens0 half page-locked)
        ctx.pin(h_in)
        swin = rbd.RowWindow(ctx, NSENS, ncols)
        shost = rbd.SharedHostMatrix(ctx, NSENS, ncols)
        d_rows = torch.empty((r1 - r0, ncols, 3), dtype=torch.float32, device="cuda")
        d_full = torch.empty((NSENS, ncols, 3), dtype=torch.float32, device="cuda") if rank == 0 else None
        h_full = np.empty((NSENS, ncols, 3), dtype=np.float32) if rank == 0 else None
        if rank == 0:
            ctx.pin(h_full)
        reps = 3

        def s_window():                    # device rays -> matrix in rank 0's HBM, rows stored remotely by k_finish
            flush.fill_(1)
            swin.rcontrib(d_mine.data_ptr(), r1 - r0, 1, flags, rays_on_device=True)
            dist.barrier()

        def s_trace_local():
            flush.fill_(1)
            ctx.rcontrib_device(d_mine.data_ptr(), r1 - r0, 1, flags, r0, d_rows.data_ptr(), d_rows.numel())

        def s_gather():
            rbd.gather_rows(d_rows, NSENS, out=d_full)

        def s_nccl():
            s_trace_local(); s_gather()

        def s_host():                      # host rays -> shared pinned host matrix (per-GPU D2H, N links)
            flush.fill_(1)
            ctx.rcontrib(h_in, accum=1, flags=flags, row_base=r0, out=shost.array[r0:r1])
            dist.barrier()

        def s_window_host():               # host rays -> window -> one D2H by rank 0
            flush.fill_(1)
            swin.rcontrib(sens0, accum=1, flags=flags)
            dist.barrier()
            if rank == 0:
                swin.download(h_full)
        res = {}
        for name, fn in (("window", s_window), ("trace_local", s_trace_local), ("nccl", s_nccl), ("gather", s_gather),
                         ("host_shared", s_host), ("host_window_d2h", s_window_host)):
            fn()
            res[name] = timed(fn, reps) / reps
        a0, b0 = rbd.shard_range(NSENS, 0, world)
        gbytes = (NSENS - (b0 - a0)) * ncols * 12          # rows that travel to rank 0
        # rank 0 alone on the same job, same run (the strong-scaling denominator)
        single_ms = None
        if rank == 0:
            d_all = torch.from_numpy(np.ascontiguousarray(sens0)).to("cuda")

            def s_single():
                flush.fill_(1)
                ctx.rcontrib_device(d_all.data_ptr(), NSENS, 1, flags, 0, d_full.data_ptr(), d_full.numel())
            s_single()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(stream)
            for _ in range(reps):
                s_single()
            e1.record(stream); torch.cuda.synchronize()
            single_ms = e0.elapsed_time(e1) / reps
            ref_sum = float(d_full.sum(dtype=torch.float64).item())
            swin.download(h_full)
            np.testing.assert_allclose(h_full.sum(dtype=np.float64), ref_sum, rtol=1e-3)
            np.testing.assert_allclose(shost.array.sum(dtype=np.float64), ref_sum, rtol=1e-3)
        barrier()
        if rank == 0:
            strong = {
                "job": f"{NSENS} sensors (the N = 1 job) sharded by dist.shard_range over {world} ranks, matrix "
                       f"[{NSENS}, {ncols}, 3] float32 delivered to rank 0",
                "single_gpu_ms": single_ms,
                "window_ms": res["window"], "window_speedup": single_ms / res["window"],
                "window_efficiency": single_ms / res["window"] / world,
                "nccl_ms": res["nccl"], "trace_local_ms": res["trace_local"], "gather_ms": res["gather"],
                "gather_bytes": int(gbytes), "gather_GBps": gbytes / (res["gather"] / 1e3) / 1e9,
                "host_shared_ms": res["host_shared"], "host_window_d2h_ms": res["host_window_d2h"],
                "host_speedup": None,
                "limit": "max over ranks of the per-rank trace time (tail waves that do not fill 148 SMs, one host "
                         "round trip per wave) -- compare trace_local_ms x N with single_gpu_ms; the gather itself is "
                         "gather_ms (NCCL) or hidden behind the tracing (window)",
                "timing": f"CUDA events on the launching stream, barrier + synchronize both sides, max over ranks, mean of {reps}",
            }
        swin.close(); shost.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        oc = oracle_counts(octf, sens)
        bpr, counts = oc["bytes_per_ray"], oc["counts"]
        launch_s = (st["wave_ms"] / 1e3) / max(1, st["wave_launches"])
        rays_per_launch = rays_local / max(1, st["wave_launches"])
        achieved = bpr * rays_per_launch / launch_s / 1e9
        # figures that only an ncu --set full capture gives (DRAM bytes, lanes per instruction, L2 hit rate) come from
        # the committed capture named in `ncu_capture`; they describe the build whose hash that file records
        traffic, extra = None, {}
        try:
            prof = json.load(open(ROOT / "profiles" / "k_trace_ncu_current.json"))
            traffic = prof["dram_bytes_per_launch"] / prof["rays_in_launch"] * rays_per_launch
            extra["warp_execution_efficiency_ncu"] = prof.get("thread_inst_per_inst", 0) / 32.0
            extra["l2_hit_rate_ncu"] = prof.get("l2_hit_rate")
            extra["ncu_capture"] = {"file": "profiles/k_trace_ncu_current.json", "capture": prof.get("capture"),
                                    "librb200_sha16_at_capture": prof.get("librb200_sha16")}
        except Exception:
            pass
        try:
            nsm = torch.cuda.get_device_properties(local).multi_processor_count
            extra["k_trace_rays_per_s_per_sm"] = rays_per_launch / launch_s / nsm
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(world),
            "sensors_per_s": NSENS * world * args.steps / (ms / 1e3),
            "rays_per_sensor": rays_local / args.steps / NSENS, "rays_per_sensor_oracle": oc["rays_per_sensor"],
            "rays_per_step_per_gpu": rays_local / args.steps, "dc_matrix_wall_ms": ms / args.steps,
            "per_rank": {"kernel_ms_per_step": [r[0] for r in per_rank_all], "k_trace_ms_per_step": [r[1] for r in per_rank_all],
                         "launches_per_step": [r[2] for r in per_rank_all],
                         "outside_kernels_ms_per_step": ms / args.steps - max(r[0] for r in per_rank_all)},
            "librb200_sha16": file_sha16(_lib.LIB_PATH),
            "scene_load_ms": 1e3 * t_load,      # octree read, flattening, level-K cell table, upload (not in any timed step)
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "kernel_ms_per_step": st_e2e["kernel_ms"] / args.steps, "batches_per_step": st_e2e["batches"] / args.steps,
                    "h2d_bytes_per_step": int(h_rays.nbytes) * world, "d2h_bytes_per_step": int(NSENS * ncols * 12) * world,
                    "path": "pinned host rays -> rb_rcontrib -> rows D2H into " +
                            ("one shared pinned host matrix (dist.SharedHostMatrix), barrier" if world > 1 else "a pinned host matrix")},
            "gpu_launches": int(st["launches"]),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_trace", "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"
                         if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_ray": bpr, "counts_per_ray": counts,
                         "avg_launch_ms": launch_s * 1e3, "launches": int(st["wave_launches"]),
                         "k_trace_share_of_kernel_time": st["wave_ms"] / max(1e-9, st["kernel_ms"]), **extra},
        }
        if world > 1:
            line["gather"] = {"inside_timed_region": True, "how": "peer-memory window (CUDA IPC over NVLink): k_finish of every "
                              "batch stores its rows into the one matrix in rank 0's HBM; a barrier ends the step",
                              "bytes_per_step": int((world - 1) * NSENS * ncols * 12)}
            line["strong"] = strong
        try:
            from oracle import refrun
            if refrun.available():
                cores = os.cpu_count() or 1
                rate, n, dt = reference_rcontrib_rate(octf, sens, cores)
                rps = oc["rays_per_sensor"]
                line["cpu_baseline"] = {"value": rate * rps, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sensors_per_s": rate,
                                        "sample": f"{n} of {NSENS} sensors (stratified) in {dt:.1f} s, unmodified reference "
                                                  f"rcontrib -n {cores} (oracle/_ref), start-up included; rays/sensor = "
                                                  f"{rps:.0f} counted by the oracle port on {RAYS_SAMPLE} sensors"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not present on this box"}
        except Exception as e:      # never lose the GPU numbers to a baseline hiccup
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        emit(line)
    if world > 1:
        win.close(); host.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


_OUT = None


def emit(obj):
    """The one JSON line of the contract, on the process's real stdout."""
    f = _OUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def main():
    # Libraries print to stdout too (NCCL's version banner under NCCL_DEBUG=VERSION/INFO): keep the real
    # stdout for the JSON line and send everything else written to fd 1 to stderr.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c1"])
    args = ap.parse_args()
    if args.config == "c1":
        return run_c1(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
