#!/usr/bin/env python
"""bench.py — DUNK registration hot path on B200 (contract: task brief ④).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through libdunk_b200.so)
  python bench.py --impl reference ...                     the reference's own CPU path (OpenCV)

A "step" = one batch of query frames (1024 x 1024 u8) taken through the whole hot path —
AKAZE extract -> brute-force Hamming 2-NN + Lowe ratio against the HBM-resident reference
descriptor database -> RANSAC homography — BASELINE.json's metric "query frames/sec
(extract+match+RANSAC)".  Prints ONE JSON line on rank 0.
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

METRIC = "query_frames_per_s"
UNIT = "frames/s"
FRAME = 1024


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        s = sorted(sm)
        return {"sm_mhz": float(np.median(s[len(s) // 2:])), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ data
def build_scene(size, seed=11):
    import synthdata as synth
    return synth.synth_scene(size, seed=seed)


def scene_tiles(scene, tile=FRAME, lods=4):
    """config-4 tiling (preprocessor/src/main.rs:197-246): per LoD the scene is decimated by 2^lod and
    cut into tile x tile images; remainder rows/cols dropped.  Returns (tiles, x_off, y_off, scale)."""
    tiles, xo, yo, sc = [], [], [], []
    cur = scene
    for lod in range(lods):
        if lod > 0:
            h, w = (cur.shape[0] // 2) * 2, (cur.shape[1] // 2) * 2
            c = cur[:h, :w].astype(np.uint16)
            cur = ((c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2] + 2) >> 2).astype(np.uint8)
        rows, cols = cur.shape[0] // tile, cur.shape[1] // tile
        for r in range(rows):
            for c_ in range(cols):
                tiles.append(cur[r * tile:(r + 1) * tile, c_ * tile:(c_ + 1) * tile])
                xo.append(c_ * tile * (1 << lod)); yo.append(r * tile * (1 << lod)); sc.append(float(1 << lod))
        if rows == 0 or cols == 0:
            break
    return (np.ascontiguousarray(np.stack(tiles)), np.array(xo, np.float32), np.array(yo, np.float32),
            np.array(sc, np.float32))


def make_frames(scene, n, seed0=1000):
    """n query frames: known-homography warps of random 1024^2 windows of the scene (config 5)."""
    import synthdata as synth
    try:
        import cv2
    except Exception:
        cv2 = None
    rng = np.random.default_rng(seed0)
    frames, Hs = [], []
    S = scene.shape[0]
    for i in range(n):
        x0, y0 = rng.uniform(64, S - FRAME - 64, 2)
        H = synth.window_homography(x0, y0, seed0 + i)           # scene -> frame
        if cv2 is not None:
            f = cv2.warpPerspective(scene, H, (FRAME, FRAME), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                                    borderValue=1)
        else:
            f = synth.warp_perspective(scene, H, FRAME, FRAME)
        frames.append(f); Hs.append(H)
    return np.ascontiguousarray(np.stack(frames)), np.stack(Hs)


def homography_errors(res, Hs):
    """max-abs error of the recovered frame->scene homography relative to ||H||inf"""
    errs = []
    for r, H in zip(res, Hs):
        if not r["found"]:
            errs.append(np.inf); continue
        Hi = np.linalg.inv(H); Hi /= Hi[2, 2]
        errs.append(float(np.abs(r["H"].reshape(3, 3) - Hi).max() / np.abs(Hi).max()))
    return np.array(errs)


# ------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200._lib import REGISTRATION_DTYPE, check, load

    rank, local_rank, world = env_rank()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load()
    ctx = dunk.Context(local_rank, 4)
    slot = ctx.reserve_slot()
    stream = torch.cuda.ExternalStream(ctx.stream(slot), device=dev)
    B = args.frames

    # ---- reference DB (config 4 style) and query frames (config 5 style); not timed
    scene = build_scene(args.scene)
    tiles, xo, yo, sc = scene_tiles(scene)
    db = dunk.feature_database.DescriptorDatabase(ctx, capacity=tiles.shape[0] * 12000)
    t0 = time.perf_counter()
    counts = db.append_tiles(tiles, xo, yo, sc, np.arange(len(tiles), dtype=np.int32))
    db_build_s = time.perf_counter() - t0
    # every rank registers its own frame batch against the (replicated) DB: frames partition with
    # no collective; the sharded-DB matcher is benchmarked by --workload match
    frames, Hs = make_frames(scene, B, seed0=1000 + 7919 * rank)
    nbytes = frames.nbytes
    f_pin = torch.from_numpy(frames).pin_memory()
    f_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ws_bytes = int(lib.dunk_register_workspace_bytes(db.handle, B, FRAME, FRAME))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    res_dev = torch.zeros(B * REGISTRATION_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    res_pin = torch.zeros(B * REGISTRATION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    f_dev.copy_(f_pin.view(-1))
    torch.cuda.synchronize(dev)

    def device_step():
        check(lib.dunk_register_frames_dev(db.handle, slot, f_dev.data_ptr(), B, FRAME, FRAME, 1, FRAME, FRAME * FRAME,
                                           args.ratio, 3.0, 0, ws.data_ptr(), ws_bytes, res_dev.data_ptr()))

    def e2e_step():
        with torch.cuda.stream(stream):
            f_dev.copy_(f_pin.view(-1), non_blocking=True)
        device_step()
        with torch.cuda.stream(stream):
            res_pin.copy_(res_dev, non_blocking=True)
        ctx.sync(slot)
        return res_pin.numpy().view(REGISTRATION_DTYPE)

    def barrier():
        ctx.sync(slot)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(stream)
    for _ in range(args.steps):
        device_step()
    t1e.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    total_ms = t0e.elapsed_time(t1e)
    clocks = sampler.stop() if sampler else None

    for _ in range(2):
        res = e2e_step()
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        res = e2e_step().copy()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)

    # per-stage device times (CUDA events inside the library) for the roofline of the top kernel
    stage = stage_times(ctx, lib, db, slot, f_dev, B, ws, ws_bytes, res_dev, args) if rank == 0 else None

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    total_ms, e2e_ms = max_over_ranks(total_ms), max_over_ranks(e2e_ms)
    if rank == 0:
        peaks, peak_src = measured_peaks()
        ms_per_step = total_ms / args.steps
        errs = homography_errors(res, Hs)
        kp_mean = float(res["keypoints"].mean())
        out = {
            "metric": METRIC, "value": world * B * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 stencils / u32 popc / f64+f32 RANSAC",
            "data": "synthetic",
            "config": {"workload": f"config5 per-GPU shard: batch of {B} query frames {FRAME}x{FRAME} u8 (known-homography "
                                   f"warps of windows of a {args.scene}^2 synthetic scene) -> AKAZE extract -> Hamming 2-NN + "
                                   f"ratio {args.ratio} vs the HBM-resident reference DB ({len(db)} descriptors from "
                                   f"{len(tiles)} tiles, 4 LoDs) -> RANSAC homography (thr 3.0, 2000 it, conf 0.995)",
                       "frames_per_step_per_gpu": B, "db_rows": len(db), "db_tiles": int(len(tiles)),
                       "keypoints_per_frame_mean": kp_mean, "parallelism": f"frame-batch dp{world}, DB replicated",
                       "l2": "inputs larger than L2 (frame batch %.0f MB + %.1f GB scale-space workspace per step)"
                             % (nbytes / 1e6, ws_bytes / 1e9)},
            "e2e": {"value": world * B * 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
                    "d2h_bytes_per_step": int(B * REGISTRATION_DTYPE.itemsize)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "quality": {"registered": int((res["found"] == 1).sum()), "frames": B,
                        "H_err_median": float(np.median(errs[np.isfinite(errs)])) if np.isfinite(errs).any() else None,
                        "H_err_max": float(errs[np.isfinite(errs)].max()) if np.isfinite(errs).any() else None,
                        "inliers_mean": float(res["inliers"].mean()), "matches_mean": float(res["matches"].mean())},
            "db_build": {"tiles": int(len(tiles)), "rows": len(db), "seconds": db_build_s,
                         "tiles_per_s": len(tiles) / db_build_s},
        }
        if stage:
            out["stages_ms_per_step"] = stage["stages"]
            out["roofline"] = stage["roofline"](peaks, peak_src)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(scene, tiles, xo, yo, sc, frames, args)
        print(json.dumps(out), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)   # torch frees tensors at exit against our external stream; skip the teardown race


def run_ours_sharded(args):
    """N > 1: the reference DB is sharded over the ranks by contiguous row range (each rank extracts its
    share of the scene tiles), every rank extracts + RANSACs its own frame batch, and the matcher is
    the one exchange step: all-gather of the ranks' query descriptors, local top-2 of
    every query against the local shard, all-to-all of the 16-byte top-2 records back to the frame
    owners, lexicographic (distance, index) merge.  Per-GPU work is constant in N -> weak scaling."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200._lib import REGISTRATION_DTYPE, PipelineView, check, load
    from cubesat_apds_b200 import sharding

    rank, local_rank, world = env_rank()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    lib = load()
    ctx = dunk.Context(local_rank, 4)
    slot = ctx.reserve_slot()
    stream = torch.cuda.ExternalStream(ctx.stream(slot), device=dev)
    B = args.frames

    scene = build_scene(args.scene)
    tiles, xo, yo, sc = scene_tiles(scene)
    T = len(tiles)
    t_lo, t_hi = sharding.shard_ranges(T, world)[rank]
    # every rank extracts its share of the tiles, then the rows are re-cut into equal contiguous
    # row ranges (tiles of coarser LoDs carry more keypoints; equal ROW counts balance the matcher)
    tmp = dunk.feature_database.DescriptorDatabase(ctx, capacity=max(1, (t_hi - t_lo)) * 12000)
    t0 = time.perf_counter()
    if t_hi > t_lo:
        tmp.append_tiles(tiles[t_lo:t_hi], xo[t_lo:t_hi], yo[t_lo:t_hi], sc[t_lo:t_hi], np.arange(t_lo, t_hi, dtype=np.int32))
    db_build_s = time.perf_counter() - t0
    n_local = torch.tensor([len(tmp)], dtype=torch.int64, device=dev)
    got = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(got, n_local)
    ext_sizes = [int(x[0]) for x in got]
    ext_bases = np.concatenate([[0], np.cumsum(ext_sizes)]).astype(np.int64)
    total_rows, max_rows = int(ext_bases[-1]), max(ext_sizes)

    def gather_column(ptr, width):
        loc = torch.zeros(max_rows * width, dtype=torch.uint8, device=dev)
        check(lib.dunk_memcpy_dev(ctx.handle, slot, loc.data_ptr(), ptr, len(tmp) * width))
        ctx.sync(slot)
        pad = torch.empty(world * max_rows * width, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(pad, loc)
        full = torch.empty(total_rows * width, dtype=torch.uint8, device=dev)
        for r in range(world):
            full[int(ext_bases[r]) * width:int(ext_bases[r + 1]) * width] = pad[r * max_rows * width:(r * max_rows + ext_sizes[r]) * width]
        return full
    desc_all = gather_column(lib.dunk_db_descriptors_dev(tmp.handle), 64)
    kps_all = gather_column(lib.dunk_db_keypoints_dev(tmp.handle), 28)       # replicated: global row -> keypoint
    torch.cuda.synchronize(dev)
    tmp.close()
    ranges = sharding.shard_ranges(total_rows, world)
    sizes = [b - a for a, b in ranges]
    bases = np.array([a for a, _ in ranges] + [total_rows], dtype=np.int64)
    r_lo, r_hi = ranges[rank]
    db = dunk.feature_database.DescriptorDatabase(ctx, capacity=max(1, r_hi - r_lo))
    check(lib.dunk_db_append_dev(db.handle, slot, desc_all.data_ptr() + r_lo * 64, kps_all.data_ptr() + r_lo * 28, None, r_hi - r_lo))
    del desc_all
    torch.cuda.synchronize(dev)

    frames, Hs = make_frames(scene, B, seed0=1000 + 7919 * rank)
    nbytes = frames.nbytes
    f_pin = torch.from_numpy(frames).pin_memory()
    f_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    f_dev.copy_(f_pin.view(-1))
    ws_bytes = int(lib.dunk_pipeline_workspace_bytes(ctx.handle, B, FRAME, FRAME))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    res_dev = torch.zeros(B * REGISTRATION_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    res_pin = torch.zeros(B * REGISTRATION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    QCAP = B * 4096                      # query rows exchanged per rank (pad); checked every step
    q_pad = torch.zeros(QCAP * 64, dtype=torch.uint8, device=dev)
    q_all = torch.empty(world * QCAP * 64, dtype=torch.uint8, device=dev)
    nq_t = torch.zeros(1, dtype=torch.int32, device=dev)
    nq_all = torch.zeros(world, dtype=torch.int32, device=dev)
    nq_pin = torch.zeros(world, dtype=torch.int32).pin_memory()
    top2_out = torch.empty(world * QCAP * 16, dtype=torch.uint8, device=dev)     # [source rank][query]
    parts = torch.empty(world * QCAP * 16, dtype=torch.uint8, device=dev)        # [shard][my query]
    torch.cuda.synchronize(dev)
    view = PipelineView()

    def device_step():
        check(lib.dunk_pipeline_extract_dev(ctx.handle, slot, f_dev.data_ptr(), B, FRAME, FRAME, 1, FRAME, FRAME * FRAME, 0,
                                            ws.data_ptr(), ws_bytes, C.byref(view)))
        nq = view.total_queries
        assert nq <= QCAP, f"{nq} queries exceed the exchange capacity {QCAP}"
        check(lib.dunk_memcpy_dev(ctx.handle, slot, q_pad.data_ptr(), view.query64_dev, nq * 64))
        with torch.cuda.stream(stream):
            nq_t.fill_(nq)
            dist.all_gather_into_tensor(nq_all, nq_t)
            dist.all_gather_into_tensor(q_all, q_pad)
            nq_pin.copy_(nq_all, non_blocking=True)          # D2H on the pipeline stream, after the all-gather
        ctx.sync(slot)                                      # host sync: the matcher grids depend on the counts
        counts = nq_pin.tolist()
        for r in range(world):
            check(lib.dunk_db_knn2_dev(db.handle, slot, q_all.data_ptr() + r * QCAP * 64, counts[r], int(bases[rank]),
                                       top2_out.data_ptr() + r * QCAP * 16))
        with torch.cuda.stream(stream):
            dist.all_to_all_single(parts, top2_out)
        check(lib.dunk_pipeline_finish_dev(ctx.handle, slot, B, FRAME, FRAME, parts.data_ptr(), world, QCAP, nq,
                                           kps_all.data_ptr(), 0, args.ratio, 3.0, ws.data_ptr(), ws_bytes, res_dev.data_ptr()))

    def e2e_step():
        with torch.cuda.stream(stream):
            f_dev.copy_(f_pin.view(-1), non_blocking=True)
        device_step()
        with torch.cuda.stream(stream):
            res_pin.copy_(res_dev, non_blocking=True)
        ctx.sync(slot)
        return res_pin.numpy().view(REGISTRATION_DTYPE)

    def barrier():
        ctx.sync(slot)
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(stream)
    for _ in range(args.steps):
        device_step()
    t1e.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    total_ms = t0e.elapsed_time(t1e)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        res = e2e_step()
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        res = e2e_step().copy()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    errs = homography_errors(res, Hs)
    ok = torch.tensor([int((res["found"] == 1).sum()), int(np.isfinite(errs).sum() and (errs[np.isfinite(errs)] < 5e-3).sum())],
                      dtype=torch.int64, device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        ms_per_step = total_ms / args.steps
        out = {
            "metric": METRIC, "value": world * B * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 stencils / u32 popc / f64+f32 RANSAC", "data": "synthetic",
            "config": {"workload": f"config5: {world} x {B} query frames {FRAME}x{FRAME} u8 per step (known-homography warps of windows "
                                   f"of a {args.scene}^2 synthetic scene) -> AKAZE extract (frame-partitioned) -> Hamming 2-NN + ratio "
                                   f"{args.ratio} vs the reference DB SHARDED over {world} GPUs by row range ({int(bases[-1])} descriptors, "
                                   f"{T} tiles, 4 LoDs; query all-gather, local top-2, top-2 all-to-all, (dist,idx) merge) -> RANSAC "
                                   f"homography (frame-partitioned)",
                       "frames_per_step_per_gpu": B, "db_rows": int(bases[-1]), "db_rows_per_shard": sizes, "db_tiles": int(T),
                       "parallelism": f"frames dp{world} + DB row-shard{world}",
                       "l2": "inputs larger than L2 (frame batch %.0f MB + %.1f GB scale-space workspace per step per GPU)"
                             % (nbytes / 1e6, ws_bytes / 1e9)},
            "e2e": {"value": world * B * 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(nbytes) * world,
                    "d2h_bytes_per_step": int(B * REGISTRATION_DTYPE.itemsize) * world},
            "gpu_launches": int(launches), "clocks": clocks,
            "quality": {"registered_all_ranks": int(ok[0]), "H_err_below_5e-3_all_ranks": int(ok[1]), "frames_all_ranks": world * B,
                        "inliers_mean_rank0": float(res["inliers"].mean()), "matches_mean_rank0": float(res["matches"].mean())},
            "collectives_per_step": {"all_gather_query_bytes_per_rank": int(QCAP * 64), "all_to_all_top2_bytes_per_rank": int(world * QCAP * 16)},
            "db_build": {"tiles_this_rank": int(t_hi - t_lo), "rows_this_rank": len(db), "seconds": db_build_s},
        }
        print(json.dumps(out), flush=True)
    barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def run_match(args):
    """--workload match — BASELINE config 3: one query frame's descriptors against a DB of args.db_rows
    uniform-random 61-byte rows (seeded, identical whatever N is) sharded over the N ranks by contiguous
    row range; per step: local top-2 on every shard, one NCCL all-gather of the 16-byte records,
    lexicographic (distance, index) merge, ratio test.  Total work is fixed -> strong scaling.  The
    query rows are planted after the random rows of the last shard, so correctness is checkable."""
    import torch
    import torch.distributed as dist
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200._lib import check, load
    from cubesat_apds_b200 import sharding

    rank, local_rank, world = env_rank()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load()
    ctx = dunk.Context(local_rank, 4)
    slot = ctx.reserve_slot()
    stream = torch.cuda.ExternalStream(ctx.stream(slot), device=dev)
    nq, nt = args.queries, args.db_rows
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (nq, 61), dtype=np.uint8)
    q[:, 60] &= 0x3F
    lo, hi = sharding.shard_ranges(nt, world)[rank]
    planted = nq if rank == world - 1 else 0
    db = dunk.feature_database.DescriptorDatabase(ctx, capacity=max(1, hi - lo + planted))
    db.append_random(hi - lo, 7, global_row_offset=lo)
    if planted:
        db.append(q)                                      # global rows nt .. nt + nq - 1
    q64 = np.zeros((nq, 64), np.uint8)
    q64[:, :61] = q
    q_pin = torch.from_numpy(q64).pin_memory()
    q_dev = torch.empty(nq * 64, dtype=torch.uint8, device=dev)
    q_dev.copy_(q_pin.view(-1))
    local = torch.empty(nq * 16, dtype=torch.uint8, device=dev)
    parts = torch.empty(world * nq * 16, dtype=torch.uint8, device=dev)
    merged = torch.empty(nq * 16, dtype=torch.uint8, device=dev)
    matches = torch.empty(nq * 16, dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    m_pin = torch.zeros(nq * 16, dtype=torch.uint8).pin_memory()
    c_pin = torch.zeros(1, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize(dev)

    def device_step():
        check(lib.dunk_db_knn2_dev(db.handle, slot, q_dev.data_ptr(), nq, lo, local.data_ptr()))
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(parts, local)
            check(lib.dunk_top2_merge_dev(ctx.handle, slot, parts.data_ptr(), world, nq, merged.data_ptr()))
            src = merged
        else:
            src = local
        check(lib.dunk_top2_ratio_dev(ctx.handle, slot, src.data_ptr(), nq, args.ratio, matches.data_ptr(), count.data_ptr()))
        return src

    def e2e_step():
        with torch.cuda.stream(stream):
            q_dev.copy_(q_pin.view(-1), non_blocking=True)
        device_step()
        with torch.cuda.stream(stream):
            m_pin.copy_(matches, non_blocking=True)
            c_pin.copy_(count, non_blocking=True)
        ctx.sync(slot)

    def barrier():
        ctx.sync(slot)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        src = device_step()
    barrier()
    top2 = src.cpu().numpy().view(dunk.TOP2_DTYPE)
    planted_ok = bool((top2["d1"] == 0).all() and (top2["i1"] == nt + np.arange(nq)).all())
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(stream)
    for _ in range(args.steps):
        device_step()
    t1e.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    total_ms = t0e.elapsed_time(t1e)
    clocks = sampler.stop() if sampler else None
    e2e_step()
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        ms = total_ms / args.steps
        popc = ctx.microbench_popc()
        pairs = float(nq) * float(nt + nq)
        ach, peak = pairs / (ms * 1e-3) / 1e9, popc * 1e3 / 16.0 * world
        out = {"metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 xor+popc",
               "data": "synthetic",
               "config": {"workload": f"config3-match: 1 query frame ({nq} x 61-B MLDB descriptors) vs {nt} reference descriptors, "
                                      f"brute-force Hamming 2-NN + ratio {args.ratio}, DB sharded over {world} GPU(s) by row range, "
                                      f"all-gather of top-2 records + (distance, index) merge",
                          "stages": "match only", "db_rows": nt, "queries_per_frame": nq, "parallelism": f"db-shard{world}",
                          "l2": "inputs larger than L2 (DB shard %.2f GB)" % ((hi - lo) * 64 / 1e9)},
               "matcher_gpairs_per_s": ach, "planted_rows_found": planted_ok,
               "e2e": {"value": 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(nq * 64),
                       "d2h_bytes_per_step": int(nq * 16 + 4), "matches": int(c_pin[0])},
               "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"bound": "int", "kernel": "hamming_top2_kernel", "achieved": ach, "peak": peak,
                            "unit": "Gpairs/s (16 POPC per pair, all GPUs)", "frac": ach / peak,
                            "peak_source": "POPC-pipe microbenchmark measured in this run (%.2f Tpopc/s per GPU)" % popc,
                            "hbm_achieved_gbs": (nt * 64.0) / (ms * 1e-3) / 1e9, "traffic": None}}
        print(json.dumps(out), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def cv2_akaze():
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    return cv2, cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, cv2.KAZE_DIFF_PM_G2, (1 << 18) - 1)


def config2_frames(n, distinct=16):
    """frames synth(1024, seed = 100 + i); `distinct` images cycled to bound the host time"""
    import synthdata
    base = np.stack([synthdata.synth_image(FRAME, FRAME, 100 + i) for i in range(min(n, distinct))])
    return np.ascontiguousarray(base[np.arange(n) % base.shape[0]])


def extract_cpu(frames, n):
    cv2, ak = cv2_akaze()
    ak.detectAndCompute(frames[0], None)
    t0 = time.perf_counter()
    for i in range(n):
        ak.detectAndCompute(frames[i % len(frames)], None)
    dt = (time.perf_counter() - t0) / n
    return dt, cv2.__version__


def run_extract(args):
    """--workload extract — BASELINE config 2: a batch of 1024 x 1024 u8 frames through AKAZE detect + MLDB
    describe, nothing else.  Frames partition over the ranks with no collective (weak scaling)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200._lib import KEYPOINT_DTYPE, PipelineView, check, load

    rank, local_rank, world = env_rank()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load()
    ctx = dunk.Context(local_rank, 4)
    slot = ctx.reserve_slot()
    stream = torch.cuda.ExternalStream(ctx.stream(slot), device=dev)
    B, sub = args.extract_frames, 64
    frames = config2_frames(B)
    f_pin = torch.from_numpy(frames).pin_memory()
    f_dev = torch.empty(frames.nbytes, dtype=torch.uint8, device=dev)
    f_dev.copy_(f_pin.view(-1))
    ws_bytes = int(lib.dunk_pipeline_workspace_bytes(ctx.handle, sub, FRAME, FRAME))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    view = PipelineView()
    cap = 8192                                              # output rows per frame of the host-buffer call
    kps_host = np.zeros((B, cap), dtype=KEYPOINT_DTYPE)
    desc_host = np.zeros((B, cap, 61), dtype=np.uint8)
    counts = np.zeros(B, dtype=np.int32)
    torch.cuda.synchronize(dev)

    def device_step():
        total = 0
        for f0 in range(0, B, sub):
            nf = min(sub, B - f0)
            check(lib.dunk_pipeline_extract_dev(ctx.handle, slot, f_dev.data_ptr() + f0 * FRAME * FRAME, nf, FRAME, FRAME, 1,
                                                FRAME, FRAME * FRAME, 0, ws.data_ptr(), ws_bytes, C.byref(view)))
            total += view.total_queries
        return total

    def e2e_step():
        # the reference-facing call: host frames in, host keypoints + descriptors out (H2D and D2H inside)
        check(lib.dunk_akaze_extract_batch(ctx.handle, frames.ctypes.data, B, FRAME, FRAME, 1, FRAME, FRAME * FRAME, 0,
                                           kps_host.ctypes.data, desc_host.ctypes.data, cap, counts.ctypes.data))
        return int(counts.sum())

    def barrier():
        ctx.sync(slot)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        n_kp = device_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record(stream)
    for _ in range(args.steps):
        device_step()
    t1e.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    total_ms = t0e.elapsed_time(t1e)
    clocks = sampler.stop() if sampler else None
    n_e2e = e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        peaks, peak_src = measured_peaks()
        stages = profile_stages(ctx, lib, slot, device_step, max(2, args.steps))
        ms = total_ms / args.steps
        out = {"metric": "extract_frames_per_s", "value": world * B * 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32 stencils", "data": "synthetic",
               "config": {"workload": f"config2-extract: batch of {B} frames {FRAME}x{FRAME} u8 gray per GPU (synth(1024, 100+i)) -> AKAZE "
                                      f"detect + MLDB-486 describe only, sub-batches of {sub}",
                          "frames_per_step_per_gpu": B, "keypoints_per_frame_mean": n_kp / B,
                          "parallelism": f"frame-batch dp{world}",
                          "l2": "inputs larger than L2 (%.0f MB of frames + %.1f GB scale-space workspace per sub-batch)"
                                % (frames.nbytes / 1e6, ws_bytes / 1e9)},
               "e2e": {"value": world * B * 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
                       "d2h_bytes_per_step": int(n_e2e * (KEYPOINT_DTYPE.itemsize + 61) + 4 * B),
                       "call": "dunk_akaze_extract_batch (pageable host buffers)"},
               "gpu_launches": int(launches), "clocks": clocks,
               "stages_ms_per_step": stages, "roofline": hbm_roofline(stages, peaks, peak_src)}
        if not args.no_cpu_baseline and world == 1:
            dt, ver = extract_cpu(frames, 16)
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                                   "sample": f"16 of the {B} frames through cv2 {ver} AKAZE.detectAndCompute, {dt * 1e3:.0f} ms/frame"}
        print(json.dumps(out), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def config4_bands(size, seed=11):
    """config 4 scene as the three f32 bands the preprocessor reads (geotiff_extractor); the synthetic scene is
    one u8 plane, so the bands are that plane with per-band gains (band_merger's min-max undoes them)"""
    scene = build_scene(size, seed).astype(np.float32)
    return scene * 40.0, scene * 36.0 + 100.0, scene * 30.0 + 50.0, np.array([0, 255 * 40.0, 100, 100 + 255 * 36.0, 50, 50 + 255 * 30.0], np.float64)


def build_cpu(bands, mm, lods, n_tiles_sample):
    """the preprocessor's per-tile work on the CPU (main.rs:258-301): window -> INTER_AREA down-sample -> band_merger
    (numpy) -> cv2 AKAZE; a bounded sample of LoD-0 tiles plus the one top-LoD tile"""
    from oracle import geo_oracle
    cv2, ak = cv2_akaze()
    H, W = bands[0].shape
    tw, th = W >> (lods - 1), H >> (lods - 1)
    todo = [(0, t % (W // tw), t // (W // tw)) for t in range(n_tiles_sample - 1)] + [(lods - 1, 0, 0)]
    t0 = time.perf_counter()
    for lod, cx, cy in todo:
        s = 1 << lod
        win = [b[cy * th * s:(cy + 1) * th * s, cx * tw * s:(cx + 1) * tw * s] for b in bands]
        if s > 1:
            win = [cv2.resize(w_, (tw, th), interpolation=cv2.INTER_AREA) for w_ in win]
        rgba = geo_oracle.band_merger(win[0].ravel(), win[1].ravel(), win[2].ravel(), mm).reshape(th, tw, 4)
        ak.detectAndCompute(np.ascontiguousarray(rgba[..., [2, 1, 0, 3]]), None)
    return (time.perf_counter() - t0) / len(todo), cv2.__version__, (tw, th)


def run_build(args):
    """--workload build — BASELINE config 4: the reference-DB build.  One scene (three f32 bands, args.build_scene^2)
    -> LoD windows (tile = scene >> 3, 4 LoDs: 64 + 16 + 4 + 1 = 85 tiles) resampled on the device -> band_merger ->
    AKAZE -> rows appended to the HBM DB with x * 2^lod + offset coordinates.  N > 1: every rank builds its own
    scene (independent replicas, no collective)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200._lib import check, load

    rank, local_rank, world = env_rank()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load()
    ctx = dunk.Context(local_rank, 4)
    slot = ctx.reserve_slot()
    S, lods = args.build_scene, 4
    r, g, b, mm = config4_bands(S)
    bands_dev = [torch.from_numpy(x).to(dev) for x in (r, g, b)]
    db = dunk.feature_database.DescriptorDatabase(ctx, capacity=4_000_000)
    n, tw, th = C.c_int(0), C.c_int(0), C.c_int(0)
    torch.cuda.synchronize(dev)

    def device_step():
        db.clear()
        check(lib.dunk_db_build_from_bands_dev(db.handle, bands_dev[0].data_ptr(), bands_dev[1].data_ptr(), bands_dev[2].data_ptr(),
                                               S, S, mm.ctypes.data, lods, 0, 0, C.byref(n), C.byref(tw), C.byref(th)))

    def e2e_step():
        db.clear()
        return db.build_from_bands(r, g, b, mm, lods)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    # the call synchronises internally (it returns row counts), so the device time is its wall time between two syncs
    w0 = time.perf_counter()
    for _ in range(args.steps):
        device_step()
    barrier()
    total_ms = (time.perf_counter() - w0) * 1e3
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    rows, tiles = len(db), n.value
    e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        peaks, peak_src = measured_peaks()
        stages = profile_stages(ctx, lib, slot, device_step, 2)
        ms = total_ms / args.steps
        mpix = sum((S >> lod << lod) ** 2 for lod in range(lods)) / 1e6          # source pixels read per build
        out = {"metric": "db_build_tiles_per_s", "value": world * tiles * 1e3 / ms, "unit": "tiles/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32 stencils", "data": "synthetic",
               "config": {"workload": f"config4-build: {S}x{S} scene (3 f32 bands) -> {tiles} tiles of {tw.value}x{th.value} over {lods} LoDs "
                                      f"(box-mean decimation) -> band_merger -> AKAZE -> {rows} DB rows in HBM",
                          "tiles": tiles, "db_rows": rows, "source_mpix_per_s": world * mpix * 1e3 / ms,
                          "parallelism": f"replicas{world}" if world > 1 else "1 GPU",
                          "l2": "inputs larger than L2 (%.2f GB of bands)" % (3 * r.nbytes / 1e9)},
               "e2e": {"value": world * tiles * 1e3 / (e2e_ms / args.steps), "unit": "tiles/s", "h2d_bytes_per_step": int(3 * r.nbytes),
                       "d2h_bytes_per_step": int(tiles * 8), "call": "dunk_db_build_from_bands (pageable host bands)"},
               "gpu_launches": int(launches), "clocks": clocks,
               "stages_ms_per_step": stages, "roofline": hbm_roofline(stages, peaks, peak_src)}
        if not args.no_cpu_baseline and world == 1:
            dt, ver, _ = build_cpu((r, g, b), mm, lods, 6)
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "tiles/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                   "sample": f"5 LoD-0 tiles + the top-LoD tile: numpy band_merger + cv2 {ver} INTER_AREA + AKAZE, "
                                             f"{dt * 1e3:.0f} ms/tile"}
        print(json.dumps(out), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def measured_traffic(kernel, shape_key):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r1_traffic.json), only when the
    capture was taken at this run's shape; otherwise None (the contract allows null)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        e = t.get(kernel, {})
        return e.get("dram_bytes_per_launch") if e.get("shape") == shape_key else None
    except Exception:
        return None


def profile_stages(ctx, lib, slot, fn, n):
    """run fn() n times between dunk_profile_begin / _end; per-kernel-class device times per call"""
    from cubesat_apds_b200._lib import check
    import ctypes as C
    check(lib.dunk_profile_begin(ctx.handle))
    for _ in range(n):
        fn()
    ctx.sync(slot)
    names = (C.c_char * 4096)()
    ms = (C.c_double * 64)()
    cnt = (C.c_int * 64)()
    alg = (C.c_double * 64)()
    k = lib.dunk_profile_end(ctx.handle, names, 4096, ms, cnt, alg, 64)
    labels = names.value.decode().split(";")[:k]
    return {lab: {"ms": ms[i] / n, "launches": cnt[i] // n, "alg_bytes_or_ops": alg[i] / n} for i, lab in enumerate(labels)}


def hbm_roofline(stages, peaks, peak_src):
    """roofline object for the stage with the largest device time (all extraction stages are HBM-side)"""
    top = max(stages, key=lambda s: stages[s]["ms"])
    t = stages[top]
    ach = t["alg_bytes_or_ops"] / (t["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src, "traffic": None,
            "launches_per_step": t["launches"], "ms_per_step": t["ms"],
            "share_of_step": t["ms"] / sum(x["ms"] for x in stages.values())}


def stage_times(ctx, lib, db, slot, f_dev, B, ws, ws_bytes, res_dev, args):
    """Device time per stage, measured with the library's CUDA-event profiler over extra steps."""
    from cubesat_apds_b200._lib import check
    import ctypes as C
    if not hasattr(lib, "dunk_profile_begin"):
        return None
    n = max(2, args.steps)
    check(lib.dunk_profile_begin(ctx.handle))
    for _ in range(n):
        check(lib.dunk_register_frames_dev(db.handle, slot, f_dev.data_ptr(), B, FRAME, FRAME, 1, FRAME, FRAME * FRAME,
                                           args.ratio, 3.0, 0, ws.data_ptr(), ws_bytes, res_dev.data_ptr()))
    ctx.sync(slot)
    names = (C.c_char * 4096)()
    ms = (C.c_double * 64)()
    cnt = (C.c_int * 64)()
    alg = (C.c_double * 64)()
    k = lib.dunk_profile_end(ctx.handle, names, 4096, ms, cnt, alg, 64)
    labels = names.value.decode().split(";")[:k]
    stages = {lab: {"ms": ms[i] / n, "launches": cnt[i] // n, "alg_bytes_or_ops": alg[i] / n} for i, lab in enumerate(labels)}
    top = max(stages, key=lambda s: stages[s]["ms"])

    def roofline(peaks, peak_src):
        t = stages[top]
        if top == "match.hamming_top2":
            popc = ctx.microbench_popc()
            ach = t["alg_bytes_or_ops"] / (t["ms"] * 1e-3) / 1e9            # Gpairs/s
            peak = popc * 1e3 / 16.0
            return {"bound": "int", "kernel": top, "achieved": ach, "peak": peak, "unit": "Gpairs/s (16 POPC per pair)",
                    "frac": ach / peak, "peak_source": "POPC-pipe microbenchmark measured in this run (%.2f Tpopc/s)" % popc,
                    "traffic": measured_traffic("hamming_top2_kernel", f"pipeline frames={B} scene={args.scene}"),
                    "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read + write, profiles/r1_traffic.json)",
                    "share_of_step": t["ms"] / sum(s["ms"] for s in stages.values())}
        ach = t["alg_bytes_or_ops"] / (t["ms"] * 1e-3) / 1e9                # GB/s
        return {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src, "traffic": None,
                "share_of_step": t["ms"] / sum(s["ms"] for s in stages.values())}
    return {"stages": stages, "roofline": roofline}


# ------------------------------------------------------------------------------------------ reference
def cv2_pipeline(cv2, ak, db_desc, db_pts, frame, ratio):
    """the reference's CPU path: lib.rs:61-92 -> lib.rs:94-114 -> mod.rs:231-259 (OpenCV)."""
    kps, desc = ak.detectAndCompute(frame, None)
    if desc is None or len(kps) < 4:
        return None
    chunk = (1 << 18) - 1                                   # OpenCV asserts train rows < 2^18
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    best = None
    for a in range(0, db_desc.shape[0], chunk):
        m = bf.knnMatch(desc, db_desc[a:a + chunk], 2)
        idx = np.array([[x.trainIdx for x in r] for r in m], dtype=np.int64) + a
        dist = np.array([[x.distance for x in r] for r in m], dtype=np.int32)
        if best is None:
            best = (idx, dist)
        else:
            from oracle import match_oracle as mo
            best = mo.merge_top2([best, (idx, dist)])
    idx, dist = best
    keep = dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(ratio)
    if keep.sum() < 4:
        return None
    src = np.array([kps[i].pt for i in np.nonzero(keep)[0]], np.float32)
    dst = db_pts[idx[keep, 0]]
    H, mask = cv2.findHomography(src, dst, cv2.RANSAC, 3.0)
    return H


def cv2_reference_setup(tiles, xo, yo, sc):
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    ak = cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, cv2.KAZE_DIFF_PM_G2, (1 << 18) - 1)
    return cv2, ak


def cpu_baseline(scene, tiles, xo, yo, sc, frames, args, n_frames=4, db_tiles=8):
    """Reference CPU path (OpenCV via cv2) on a bounded sample: DB of `db_tiles` tiles, `n_frames` frames;
    matcher time scaled linearly to the full DB row count."""
    try:
        cv2, ak = cv2_reference_setup(tiles, xo, yo, sc)
    except Exception as e:     # the oracle port is far too slow for a frames/s figure; report unavailable
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"cv2 unavailable: {e}"}
    cores = os.cpu_count() or 1
    descs, pts = [], []
    t0 = time.perf_counter()
    for t in range(min(db_tiles, len(tiles))):
        k, d = ak.detectAndCompute(tiles[t], None)
        if d is not None:
            descs.append(d)
            pts.append(np.array([p.pt for p in k], np.float32) * sc[t] + np.array([xo[t], yo[t]], np.float32))
    t_extract = (time.perf_counter() - t0) / max(1, min(db_tiles, len(tiles)))
    db_desc, db_pts = np.concatenate(descs), np.concatenate(pts)
    t0 = time.perf_counter()
    for i in range(n_frames):
        cv2_pipeline(cv2, ak, db_desc, db_pts, frames[i], args.ratio)
    dt = (time.perf_counter() - t0) / n_frames
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{n_frames} frames through cv2 {cv2.__version__} AKAZE + BFMatcher(k=2) + findHomography(RANSAC) against "
                      f"a {db_desc.shape[0]}-row DB ({min(db_tiles, len(tiles))} tiles; the GPU arm's DB is larger, so this "
                      f"flatters the CPU); {dt * 1e3:.0f} ms/frame, tile extraction {t_extract * 1e3:.0f} ms/tile"}


def run_reference_workload(args):
    """reference arm of the secondary workloads: the same OpenCV calls on the host cores, bounded samples"""
    cores = os.cpu_count() or 1
    base = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "vs_baseline": None, "data": "synthetic"}
    if args.workload == "extract":
        B = args.extract_frames
        frames = config2_frames(min(B, 16))
        dt, ver = extract_cpu(frames, max(8, 4 * args.steps))
        val, unit, metric, ms = 1.0 / dt, UNIT, "extract_frames_per_s", dt * 1e3 * B
        cfg = {"workload": f"config2-extract on the host CPU: cv2 {ver} AKAZE.detectAndCompute on {FRAME}x{FRAME} u8 frames",
               "frames_per_step_per_gpu": B}
        sample = f"{max(8, 4 * args.steps)} frames, {dt * 1e3:.0f} ms/frame, scaled linearly to {B} frames per step"
        dtype, scaling = "f32 (OpenCV)", "weak"
    elif args.workload == "build":
        r, g, b, mm = config4_bands(args.build_scene)
        dt, ver, (tw, th) = build_cpu((r, g, b), mm, 4, 6)
        val, unit, metric, ms = 1.0 / dt, "tiles/s", "db_build_tiles_per_s", dt * 1e3 * 85
        cfg = {"workload": f"config4-build on the host CPU: {args.build_scene}^2 scene, tiles of {tw}x{th}: numpy band_merger + "
                           f"cv2 {ver} INTER_AREA + AKAZE per tile", "tiles": 85}
        sample = f"5 LoD-0 tiles + the top-LoD tile, {dt * 1e3:.0f} ms/tile, scaled linearly to 85 tiles per step"
        dtype, scaling = "f32 (OpenCV)", "weak"
    else:
        import cv2
        cv2.setNumThreads(cores)
        nq, rows = args.queries, 1_000_000
        rng = np.random.default_rng(0)
        q = rng.integers(0, 256, (nq, 61), dtype=np.uint8)
        t = np.random.default_rng(7).integers(0, 256, (rows, 61), dtype=np.uint8)
        bf, chunk = cv2.BFMatcher(cv2.NORM_HAMMING, False), (1 << 18) - 1
        t0 = time.perf_counter()
        for a in range(0, rows, chunk):
            bf.knnMatch(q, t[a:a + chunk], 2)
        dt = time.perf_counter() - t0
        gp = nq * rows / dt / 1e9
        ms = dt * 1e3 * args.db_rows / rows
        val, unit, metric = 1e3 / ms, UNIT, METRIC
        cfg = {"workload": f"config3-match on the host CPU: cv2 {cv2.__version__} BFMatcher(HAMMING).knnMatch k=2 in <= 262 143-row chunks",
               "db_rows": args.db_rows, "queries_per_frame": nq, "matcher_gpairs_per_s": gp}
        sample = f"{nq} queries x {rows} rows ({gp:.2f} Gpairs/s), scaled linearly to {args.db_rows} rows"
        dtype, scaling = "u8 popcnt (OpenCV)", "strong"
    out = dict(base, metric=metric, value=val, unit=unit, ms_per_step=ms, scaling=scaling, dtype=dtype, config=cfg,
               cpu_baseline={"value": val, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
               e2e={"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(out), flush=True)


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    if args.workload != "pipeline":
        try:
            return run_reference_workload(args)
        except ImportError as e:
            print(json.dumps({"impl": "reference", "unavailable": f"cv2 (OpenCV) not importable: {e}"}))
            return
    scene = build_scene(args.scene)
    tiles, xo, yo, sc = scene_tiles(scene)
    B = args.frames
    try:
        cv2, ak = cv2_reference_setup(tiles, xo, yo, sc)
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 (OpenCV) not importable: {e}"}))
        return
    cores = os.cpu_count() or 1
    # full reference DB through the reference's own extraction (untimed set-up, like our arm)
    descs, pts = [], []
    n_db_tiles = len(tiles) if args.ref_full_db else min(len(tiles), args.ref_db_tiles)
    for t in range(n_db_tiles):
        k, d = ak.detectAndCompute(tiles[t], None)
        if d is not None:
            descs.append(d)
            pts.append(np.array([p.pt for p in k], np.float32) * sc[t] + np.array([xo[t], yo[t]], np.float32))
    db_desc, db_pts = np.concatenate(descs), np.concatenate(pts)
    sample = min(B, args.ref_frames)
    frames, Hs = make_frames(scene, sample, seed0=1000)
    for _ in range(min(1, args.warmup)):
        cv2_pipeline(cv2, ak, db_desc, db_pts, frames[0], args.ratio)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(sample):
            cv2_pipeline(cv2, ak, db_desc, db_pts, frames[i], args.ratio)
    dt_frame = (time.perf_counter() - t0) / (args.steps * sample)
    ms_step = dt_frame * 1e3 * B
    val = 1.0 / dt_frame
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 / u8 popcnt / f64 (OpenCV)", "data": "synthetic",
        "config": {"workload": f"config5 per-GPU shard on the host CPU: batch of {B} query frames {FRAME}x{FRAME} u8 -> "
                               f"cv2.AKAZE -> BFMatcher(HAMMING) knnMatch k=2 + ratio {args.ratio} vs {db_desc.shape[0]} reference "
                               f"descriptors -> findHomography(RANSAC, 3.0)",
                   "frames_per_step_per_gpu": B, "db_rows": int(db_desc.shape[0]), "db_tiles": int(n_db_tiles)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"each step timed on {sample} of {B} frames ({dt_frame * 1e3:.0f} ms/frame, OpenCV {cv2.__version__}, "
                                   f"{cores} threads) and scaled linearly in frames; DB {db_desc.shape[0]} rows from {n_db_tiles} of "
                                   f"{len(tiles)} tiles"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="query frames per step per GPU")
    ap.add_argument("--scene", type=int, default=8192, help="synthetic scene edge (pixels)")
    ap.add_argument("--ratio", type=float, default=0.8)
    ap.add_argument("--ref-frames", type=int, default=4)
    ap.add_argument("--ref-db-tiles", type=int, default=85)
    ap.add_argument("--ref-full-db", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicated-db", action="store_true", help="N>1: replicate the DB instead of sharding it")
    ap.add_argument("--workload", default="pipeline", choices=["pipeline", "match", "extract", "build"],
                    help="pipeline = config 5 (default, the headline metric); match = config 3 (sharded matcher only); "
                         "extract = config 2 (extraction only); build = config 4 (reference-DB build from a scene)")
    ap.add_argument("--extract-frames", type=int, default=256, help="--workload extract: frames per step per GPU")
    ap.add_argument("--build-scene", type=int, default=10980, help="--workload build: scene edge (pixels)")
    ap.add_argument("--db-rows", type=int, default=50_000_000, help="--workload match: reference descriptors")
    ap.add_argument("--queries", type=int, default=3163, help="--workload match: query descriptors per frame")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.workload == "match":
            run_match(args)
        elif args.workload == "extract":
            run_extract(args)
        elif args.workload == "build":
            run_build(args)
        elif env_rank()[2] > 1 and not args.replicated_db:
            run_ours_sharded(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
