"""The north-star target run (BASELINE.json configs[2]): the full Reinhart MF:4 (2305-bin)
daylight-coefficient matrix for 1 M sensors over the synthetic 1 M-polygon 10-floor building at
`-ab 5 -ad 10000 -lw 1e-4`, sharded by sensor row over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29531 tools/target_c3.py [--sensors 1000000] [--out profiles/r2_target_c3.json]

The WHOLE matrix (27.7 GB of float32) is gathered in rank 0's HBM while it is computed: every rank's finishing
kernel stores its rows over NVLink into the one peer-mapped window (pyradiance_b200/dist.py RowWindow), so the
timed region ends -- after a barrier -- with the matrix in one caller's hands, as the reference's rcontrib -n N
delivers it.  Timing is CUDA events + barrier, max over ranks (like bench.py).  Checks: every row sum <= pi (a sensor cannot collect more than the
whole sky), no negative / non-finite entries, the per-floor mean daylight falls off with depth the
same way on every floor, and -- when oracle/_ref travelled -- a sample of rows against the unmodified
reference rcontrib (row sums; 6 sigma of the Monte-Carlo noise + 2 %).  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

AB, AD, LW, MF = 5, 10000, 1e-4, 4
NPOLY, FLOORS = 1_000_000, 10
P = f"MF={MF},rNx=0,rNy=0,rNz=-1,Ux=0,Uy=1,Uz=0,RHS=+1"
OPTS = ["-ab", str(AB), "-ad", str(AD), "-lw", f"{LW:g}"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sensors", type=int, default=1_000_000)
    ap.add_argument("--out", default="")
    ap.add_argument("--ref-rows", type=int, default=24)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pyradiance_b200 import _lib, scenegen
    from pyradiance_b200 import dist as rbd
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = Path(os.environ.get("RB_TMP", "/tmp/rbt")); tmp.mkdir(parents=True, exist_ok=True)
    rad, octf = tmp / "bld1m.rad", tmp / "bld1m.oct"
    t_scene = 0.0
    if rank == 0 and not octf.exists():
        t = time.time()
        scenegen.write_office(rad, npolys=NPOLY, floors=FLOORS, seed=77)
        scenegen.build_octree(rad, octf.with_suffix(".tmp"))
        os.replace(octf.with_suffix(".tmp"), octf)
        t_scene = time.time() - t
    if world > 1:
        dist.barrier()
    ctx = _lib.Context(local, _lib.RB_PROGRAM_RCONTRIB)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    t = time.time(); ctx.load_octree(octf); t_load = time.time() - t
    ctx.set_options(OPTS)
    ctx.cal_load("reinhartb.cal"); ctx.cal_set(P)
    ctx.add_modifier("skyglow", P, "rbin", int(ctx.cal_eval("Nrbins") + .5))
    ncols = ctx.num_columns()
    sens = scenegen.office_sensors(args.sensors, floors=FLOORS, seed=3)
    lo, hi = rank * args.sensors // world, (rank + 1) * args.sensors // world
    mine = np.ascontiguousarray(sens[lo:hi])
    d_rays = torch.from_numpy(mine).to("cuda")
    if world > 1:
        win = rbd.RowWindow(ctx, args.sensors, ncols)           # the one matrix, in rank 0's HBM
        out_ptr = win.ptr(lo)
    else:
        d_local = torch.empty((hi - lo, ncols, 3), dtype=torch.float32, device="cuda")
        out_ptr = d_local.data_ptr()
    out_floats = (hi - lo) * ncols * 3
    # warm-up on a sliver (allocations, first launches), then the timed full job
    for _ in range(3):
        ctx.rcontrib_device(d_rays.data_ptr(), min(4000, hi - lo), 1, _lib.RB_IRRAD_RCONTRIB, lo, out_ptr, out_floats)
    ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall0 = time.time()
    e0.record(stream)
    ctx.rcontrib_device(d_rays.data_ptr(), hi - lo, 1, _lib.RB_IRRAD_RCONTRIB, lo, out_ptr, out_floats)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                      # every rank's rows are in rank 0's HBM
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.time() - wall0
    ms = e0.elapsed_time(e1)
    st = ctx.stats()
    agg = torch.tensor([ms, float(st["nrays"]), float(st["wave_ms"]), float(st["kernel_ms"])], device="cuda", dtype=torch.float64)
    mx = agg.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    # ---- checks: rank 0, on the WHOLE gathered matrix (chunks of rows, the matrix is 27.7 GB) ----
    fsum = torch.zeros(FLOORS, device="cuda", dtype=torch.float64)
    fcnt = torch.zeros(FLOORS, device="cuda", dtype=torch.float64)
    chk = torch.zeros(3, device="cuda", dtype=torch.float64)
    rows = None
    if rank == 0:
        full = win.as_tensor() if world > 1 else d_local
        all_sens = sens if world > 1 else mine
        rows = torch.empty(full.shape[0], device="cuda", dtype=torch.float64)
        bad = 0
        step = 20000
        for a in range(0, full.shape[0], step):
            blk = full[a:a + step]
            rows[a:a + step] = blk[:, :, 0].sum(1, dtype=torch.float64)
            bad += int((~torch.isfinite(blk)).sum().item()) + int((blk < 0).sum().item())
        over_pi = int((rows > np.pi * (1 + 1e-4)).sum().item())
        floor_of = torch.from_numpy(np.floor(all_sens[:, 2] / 3.3).astype(np.int64)).to("cuda")
        fsum.index_add_(0, floor_of, rows)
        fcnt.index_add_(0, floor_of, torch.ones_like(rows))
        chk = torch.tensor([float(bad), float(over_pi), float(rows.sum().item())], device="cuda", dtype=torch.float64)
    ref_note = "oracle/_ref not on this box"
    ref_ok = None
    if rank == 0:
        from oracle import refrun
        if refrun.available() and args.ref_rows > 0:
            idx = np.linspace(0, args.sensors - 1, args.ref_rows).astype(int)     # rows of every rank's block
            t = time.time()
            ref = refrun.rcontrib(octf, sens[idx], ["-I+"] + OPTS + ["-f", "reinhartb.cal", "-p", P, "-bn", "Nrbins", "-b", "rbin",
                                                                     "-m", "skyglow"], nproc=os.cpu_count()).reshape(len(idx), -1, 3)
            tref = time.time() - t
            g = rows[torch.from_numpy(idx).to("cuda")].cpu().numpy()
            r = ref[:, :, 0].sum(1)
            sig = np.sqrt(np.maximum(g, r) * np.pi / AD * 2) * 1.5          # first-level Poisson noise, both runs, widened
            ref_ok = bool(np.all(np.abs(g - r) <= 6 * sig + 0.02 * r + 1e-4))
            ref_note = (f"{len(idx)} rows vs reference rcontrib -n {os.cpu_count()} in {tref:.1f} s: max |diff| "
                        f"{np.abs(g - r).max():.4f}, totals {g.sum():.4f} vs {r.sum():.4f}")
        floors = (fsum / torch.clamp(fcnt, min=1)).cpu().numpy()
        nrays = float(agg[1].item())
        line = {
            "what": "north-star target: MF:4 daylight-coefficient matrix, -ab 5 -ad 10000 -lw 1e-4, synthetic 1M-polygon building",
            "n_gpus": world, "sensors": args.sensors, "columns": ncols,
            "matrix_bytes_fp32": args.sensors * ncols * 12,
            "gather": "peer-memory window (CUDA IPC over NVLink): whole matrix in rank 0's HBM at the end of the timed region"
                      if world > 1 else "single GPU",
            "bytes_over_nvlink": (args.sensors - (hi - lo)) * ncols * 12 if world > 1 else 0,
            "device_ms_max_over_ranks": float(mx[0].item()), "wall_s": wall,
            "rays_total": nrays, "rays_per_sec": nrays / (float(mx[0].item()) / 1e3),
            "k_trace_share": float(agg[2].item()) / max(float(agg[3].item()), 1e-9),
            "scene_build_s_rank0": t_scene, "octree_load_s": t_load,
            "rank0_stats": {k: (round(v, 2) if isinstance(v, float) else int(v)) for k, v in st.items()},
            "checks": {"nonfinite_or_negative": int(chk[0].item()), "rows_over_pi": int(chk[1].item()),
                       "mean_row_sum": float(chk[2].item()) / args.sensors,
                       "mean_row_sum_per_floor": [round(float(x), 5) for x in floors],
                       "reference_rows": ref_note, "reference_rows_ok": ref_ok},
        }
        print(json.dumps(line))
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(json.dumps(line, indent=1) + "\n")
    if world > 1:
        if rank == 0:
            del full
        win.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
