#!/usr/bin/env python
"""bench.py — DUNK registration hot path on B200 (contract: task brief ④).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through libdunk_b200.so)
  python bench.py --impl reference ...                     the reference's own CPU path (OpenCV)

Default workload = BASELINE.json config 5, the metric "query frames/sec (extract+match+RANSAC)": a "step" is one
batch of 64 query frames per GPU (1024 x 1024 u8, camera views of windows of the 10980^2 config-4 scene) taken
through the whole hot path — AKAZE extract -> brute-force Hamming 2-NN + Lowe ratio against the HBM-resident
config-4 reference database (sharded by row range over the GPUs, NCCL inside the library) -> RANSAC homography
-> world coordinates + PnP-RANSAC -> attitude.  Prints ONE JSON line on rank 0.

PyTorch appears here only as plumbing around the product: `torch.distributed` for the rendezvous (the 128-byte
NCCL id of the library's own communicator), the barrier and the max-over-ranks of the timings.  Device / pinned
memory, copies, streams, events and every collective on the data path are the library's.
"""
import argparse
import ctypes as C
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
RANSAC_THR = 3.0
PNP = {"method": "SOLVEPNP_EPNP", "iter_count": 1000, "reproj_thres": 3.0, "confidence": 0.99}
DTYPE = "f32 stencils / u32 xor+popc / f64+f32 RANSAC, PnP"


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
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(0.25)          # nvidia-smi needs ~0.2 s before its first sample: short timed regions had none
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
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


# ------------------------------------------------------------------------------------------ harness
class Harness:
    """process set-up / tear-down shared by the workloads.  Nothing of torch ever touches the library's streams, so
    the process exits normally (the driver's hook records the loaded .so at interpreter exit)."""

    def __init__(self, args):
        import torch
        import cubesat_apds_b200 as dunk
        from cubesat_apds_b200 import _lib
        self.torch, self.dunk, self._lib = torch, dunk, _lib
        self.rank, self.local_rank, self.world = env_rank()
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.lib = _lib.load()
        self.ctx = dunk.Context(self.local_rank, 4)
        self.slot = self.ctx.reserve_slot()
        self.group = None
        self._buffers = []

    def shard_group(self):
        """the library's own NCCL communicator; rank 0's 128-byte id travels through torch.distributed"""
        fd = self.dunk.feature_database
        if self.group is None:
            uid = None
            if self.world > 1:
                t = self.torch.zeros(self._lib.SHARD_ID_BYTES, dtype=self.torch.uint8, device=self.dev)
                if self.rank == 0:
                    t.copy_(self.torch.frombuffer(bytearray(fd.ShardGroup.unique_id()), dtype=self.torch.uint8))
                self.dist.broadcast(t, 0)
                uid = bytes(t.cpu().numpy().tobytes())
            self.group = fd.ShardGroup(self.ctx, self.rank, self.world, uid)
        return self.group

    def dev_buffer(self, nbytes):
        b = self._lib.DeviceBuffer(self.ctx, nbytes)
        self._buffers.append(b)
        return b

    def pinned(self, shape, dtype=np.uint8):
        b = self._lib.PinnedBuffer(self.ctx, shape, dtype)
        self._buffers.append(b)
        return b

    def barrier(self):
        self.ctx.sync(self.slot)
        self.torch.cuda.synchronize(self.dev)
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        if not self.dist:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def sum_over_ranks(self, values):
        if not self.dist:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]

    def timed(self, fn, steps):
        """`steps` calls of fn between a barrier + sync on both sides, CUDA events on the library's stream"""
        self.barrier()
        w0 = time.perf_counter()
        self.ctx.timer_begin(self.slot)
        for i in range(steps):
            fn(i)
        ms = self.ctx.timer_end(self.slot)          # records the end event on the slot's stream and waits for it
        wall = (time.perf_counter() - w0) * 1e3
        self.barrier()
        return ms, wall

    def profile(self, fn, n):
        """per-kernel-class device times (CUDA events recorded by the library around every launch class)"""
        self._lib.check(self.lib.dunk_profile_begin(self.ctx.handle))
        for i in range(n):
            fn(i)
        self.ctx.sync(self.slot)
        names = (C.c_char * 8192)()
        ms = (C.c_double * 96)()
        cnt = (C.c_int * 96)()
        alg = (C.c_double * 96)()
        k = self.lib.dunk_profile_end(self.ctx.handle, names, 8192, ms, cnt, alg, 96)
        labels = names.value.decode().split(";")[:k]
        return {lab: {"ms": ms[i] / n, "launches": cnt[i] // n, "alg_bytes_or_ops": alg[i] / n} for i, lab in enumerate(labels)}

    def close(self, *handles):
        """orderly tear-down: drain the stream, free what the bench allocated, destroy group / context / process group"""
        self.barrier()
        for h in handles:
            if h is not None:
                h.close()
        for b in self._buffers:
            b.free()
        if self.group is not None:
            self.group.close()
        self.ctx.release_slot(self.slot)
        self.ctx.close()
        if self.dist:
            self.dist.destroy_process_group()
        sys.stdout.flush()


# ------------------------------------------------------------------------------------------ config 4 / 5 data
def build_scene(size, seed=11):
    import synthdata
    return synthdata.synth_scene(size, seed=seed)


def config4_bands(scene):
    """config 4 scene as the three f32 bands the preprocessor reads (geotiff_extractor); the synthetic scene is
    one u8 plane, so the bands are that plane with per-band gains / offsets (band_merger's min-max undoes them)"""
    s = scene.astype(np.float32)
    return (s * 40.0, s * 36.0 + 100.0, s * 30.0 + 50.0,
            np.array([0, 255 * 40.0, 100, 100 + 255 * 36.0, 50, 50 + 255 * 30.0], np.float64))


def composite_gray(scene, band_merger_fn):
    """What a camera sees of the scene: the preprocessor's RGBA composite (band_merger: min-max, gamma 1/2.2) in
    gray.  The three bands are functions of one u8 plane, so the composite is a 256-entry table; the gray value is
    OpenCV's cvtColor fixed-point formula (R*4899 + G*9617 + B*1868 + 8192) >> 14."""
    v = np.arange(256, dtype=np.uint8)
    r, g, b, mm = config4_bands(v)
    rgba = np.asarray(band_merger_fn(r, g, b, mm)).reshape(256, 4).astype(np.int64)
    lut = ((rgba[:, 0] * 4899 + rgba[:, 1] * 9617 + rgba[:, 2] * 1868 + 8192) >> 14).astype(np.uint8)
    return lut[scene]


def pipeline_config(args, world):
    """identical for both arms (the driver compares the two lines' `config`)"""
    return {
        "workload": f"config5: {args.frames} query frames {FRAME}x{FRAME} u8 per GPU per step ({args.distinct} distinct frames per GPU, "
                    f"cycled; views of a 500 km nadir-ish pinhole camera onto windows of the {args.scene}^2 config-4 scene) -> AKAZE "
                    f"extract -> Hamming 2-NN + ratio {args.ratio} vs the config-4 reference DB (3 f32 bands -> 85 tiles of "
                    f"{args.scene >> 3}^2 over 4 LoDs -> band_merger -> AKAZE rows in HBM) -> findHomography(RANSAC, {RANSAC_THR}, 2000 it, "
                    f"0.995) -> get_world_coordinates (geotransform + 100 m DEM -> ECEF) -> solvePnPRansac(EPNP, {PNP['iter_count']} it, "
                    f"thr {PNP['reproj_thres']}, conf {PNP['confidence']}) -> attitude",
        "frames_per_step_per_gpu": args.frames, "distinct_frames_per_gpu": args.distinct, "scene": args.scene, "db_lods": 4,
        "db_tiles": 85, "ratio": args.ratio, "ransac_thr": RANSAC_THR, "pnp": PNP, "stages": "extract + match + RANSAC homography + PnP",
        "parallelism": f"frames dp{world}" + (f" + DB row-shard{world} (query all-gather, one local top-2 launch, top-2 all-to-all, "
                                              f"(distance, index) merge; NCCL inside libdunk_b200.so)" if world > 1 else ""),
        "l2": "inputs larger than L2 (frame batch %.0f MB + %.0f GB scale-space workspace per step per GPU; a different batch every step)"
              % (args.frames * FRAME * FRAME / 1e6, args.frames * 0.11),
    }


# hamming_top2_kernel, SASS per (query, row) pair: 15 XOR + 14 LOP3 (7 carry-save adders) + 3 IADD3 + 1 ISETP on the ALU pipe,
# 8 POPC on the XU pipe (match_hamming.cu: hamming480; the 16th word is added only on rows that pass the vote)
POPC_PER_PAIR, ALU_PER_PAIR = 8.0, 33.0


def fill_descriptor_bytes(stages, n_keypoints_per_step):
    """k_orientation / k_mldb launch before the keypoint count is known on the host, so the library records no byte
    count for them; the USEFUL bytes per keypoint are fixed by the sampling patterns: orientation = 109 samples x
    (Lx, Ly) f32, MLDB = the shared 21 x 21 lattice (441 samples) x (Lt, Lx, Ly) f32 + the 64-byte row written.  Every
    4-byte sample is a gather that moves a 32-byte sector, so sector traffic is ~8x these figures (L2-gather bound)."""
    per_kp = {"describe.orientation": 109 * 2 * 4, "describe.mldb": 441 * 3 * 4 + 64}
    for k, b in per_kp.items():
        if k in stages and not stages[k]["alg_bytes_or_ops"]:
            stages[k]["alg_bytes_or_ops"] = float(n_keypoints_per_step) * b
            stages[k]["sector_bytes"] = float(n_keypoints_per_step) * (b - (64 if "mldb" in k else 0)) * 8
            stages[k]["bound"] = "L2 gather (one 32-byte sector per 4-byte sample)"


def summarize_quality(res, poses, Hs, Rs, ts):
    import synthdata
    errs = []
    for r, H in zip(res, Hs):
        if not r["found"]:
            errs.append(np.inf); continue
        Hi = np.linalg.inv(H); Hi /= Hi[2, 2]
        errs.append(float(np.abs(r["H"].reshape(3, 3) - Hi).max() / np.abs(Hi).max()))
    errs = np.array(errs)
    ok = np.isfinite(errs)
    q = {"frames": int(len(res)), "registered": int((res["found"] == 1).sum()),
         "H_err_median": float(np.median(errs[ok])) if ok.any() else None, "H_err_max": float(errs[ok].max()) if ok.any() else None,
         "H_err_below_5e-3": int((errs[ok] < 5e-3).sum()), "inliers_mean": float(res["inliers"].mean()),
         "matches_mean": float(res["matches"].mean()), "keypoints_mean": float(res["keypoints"].mean())}
    if poses is not None:
        rot, pos = synthdata.pose_errors(poses["rvec"], poses["tvec"], poses["found"], Rs, ts)
        okp = np.isfinite(rot)
        q.update({"poses_found": int(okp.sum()), "pnp_inliers_mean": float(poses["inliers"].mean()),
                  "rot_err_deg_median": float(np.median(rot[okp])) if okp.any() else None,
                  "rot_err_deg_p90": float(np.percentile(rot[okp], 90)) if okp.any() else None,
                  "cam_pos_err_m_median": float(np.median(pos[okp])) if okp.any() else None,
                  "note": "500 km altitude, 1.2 deg field of view: rotation and lateral position are coupled (1 deg ~ 8.7 km), the "
                          "pair is recovered to the accuracy cv2 reaches on the same correspondences"})
    return q


# ------------------------------------------------------------------------------------------ ours: config 5
def run_pipeline(args):
    import synthdata
    h = Harness(args)
    dunk, lib, ctx, slot, _lib = h.dunk, h.lib, h.ctx, h.slot, h._lib
    check = _lib.check
    fd = dunk.feature_database
    rank, world = h.rank, h.world
    B, S, ND = args.frames, args.scene, args.distinct
    assert ND % B == 0, "--distinct must be a multiple of --frames"
    group = h.shard_group()

    # ---- reference DB = the config-4 build (dunk_db_build_from_bands) partitioned over the ranks by tile, then re-cut
    # into equal row ranges (dunk_shard_group_balance).  Untimed set-up, like the reference arm's.
    scene = build_scene(S)
    r, g, b, mm = config4_bands(scene)
    bands = []
    for x in (r, g, b):
        d = h.dev_buffer(x.nbytes)
        d.upload(slot, x)
        ctx.sync(slot)
        bands.append(d)
    del r, g, b
    tmp = fd.DescriptorDatabase(ctx, capacity=max(400_000, 4_000_000 // world))
    n_t, tw, th = C.c_int(0), C.c_int(0), C.c_int(0)
    t0 = time.perf_counter()
    check(lib.dunk_db_build_from_bands_part_dev(tmp.handle, C.c_void_p(bands[0].ptr), C.c_void_p(bands[1].ptr), C.c_void_p(bands[2].ptr),
                                                S, S, mm.ctypes.data, 4, 0, 0, rank, world, C.byref(n_t), C.byref(tw), C.byref(th)))
    db_build_s = time.perf_counter() - t0
    shard = group.balance(tmp)
    tmp.close()
    for d in bands:
        d.free()
    db_rows = group.total_rows

    # ---- query frames: camera views of the gray composite, warped on the device (bit-exact with cv2.warpPerspective)
    gray = composite_gray(scene, lambda r_, g_, b_, mm_: dunk.image_extractor.band_merger([r_, g_, b_], mm_, ctx=ctx))
    Hs, Rs, ts, fit_resid = synthdata.config5_views(ND, S, 1000 + 7919 * rank)
    scene_dev = h.dev_buffer(gray.nbytes)
    scene_dev.upload(slot, gray)
    ctx.sync(slot)
    frames_dev = h.dev_buffer(ND * FRAME * FRAME)
    for f0 in range(0, ND, 64):
        M = np.ascontiguousarray(Hs[f0:f0 + 64].reshape(-1, 9))
        check(lib.dunk_warp_perspective_batch_dev(ctx.handle, slot, C.c_void_p(scene_dev.ptr), S, S, 1, S, M.ctypes.data, len(M), FRAME,
                                                  FRAME, None, C.c_void_p(frames_dev.ptr + f0 * FRAME * FRAME)))
        ctx.sync(slot)
    frames_pin = h.pinned((ND, FRAME, FRAME))
    frames_dev.download(slot, frames_pin.array)
    ctx.sync(slot)
    scene_dev.free()

    # ---- pose stage: geotransform + DEM in HBM, camera matrix, PnP parameters
    gt_e, heights = synthdata.scene_dem(S)
    geo = fd.Geotransform(synthdata.scene_geotransform(), gt_e, heights, ctx)
    pose = fd.PoseStage(geo, synthdata.CAMERA_K, synthdata.scene_origin(S), PNP["iter_count"], PNP["reproj_thres"], PNP["confidence"], 1)

    RS, PS = _lib.REGISTRATION_DTYPE.itemsize, _lib.POSE_DTYPE.itemsize
    ws_bytes = int(lib.dunk_register_sharded_workspace_bytes(group.handle, B, FRAME, FRAME))
    ws = h.dev_buffer(ws_bytes)
    # e2e: the frames arrive from pinned host memory, double-buffered: while step i computes on `slot`, the frames of
    # step i + 1 are copied on a second slot (stream) -- what a caller streaming frames through the public API does
    stage_dev = [h.dev_buffer(B * FRAME * FRAME), h.dev_buffer(B * FRAME * FRAME)]
    slot_up = ctx.reserve_slot()
    res_dev, pose_dev = h.dev_buffer(B * RS), h.dev_buffer(B * PS)
    res_pin, pose_pin = h.pinned((B,), _lib.REGISTRATION_DTYPE), h.pinned((B,), _lib.POSE_DTYPE)
    n_batches = ND // B

    def step_on(images_ptr):
        check(lib.dunk_register_frames_sharded_dev(group.handle, shard.handle, slot, C.c_void_p(images_ptr), B, FRAME, FRAME, 1, FRAME,
                                                   FRAME * FRAME, args.ratio, RANSAC_THR, 0, C.byref(pose.config), C.c_void_p(ws.ptr),
                                                   ws_bytes, C.c_void_p(res_dev.ptr), C.c_void_p(pose_dev.ptr)))

    def device_step(i):                                       # frames already resident in HBM; a different batch every step
        step_on(frames_dev.ptr + (i % n_batches) * B * FRAME * FRAME)

    seq = {"next": 0, "resident": [None, None]}               # running step counter; batch index held by each staging half

    def upload(j):
        k = j % n_batches
        stage_dev[j % 2].upload(slot_up, frames_pin.array[k * B:(k + 1) * B])
        seq["resident"][j % 2] = k

    def e2e_step(_):
        # host buffers: every step issues one pinned H2D of B frames (the NEXT step's, on the second slot, overlapping this
        # step's kernels) and reads this step's records back (D2H) before it returns
        j = seq["next"]
        seq["next"] = j + 1
        if seq["resident"][j % 2] != j % n_batches:           # very first step: nothing was prefetched
            upload(j)
        ctx.sync(slot_up)                                     # this step's frames have landed
        upload(j + 1)
        step_on(stage_dev[j % 2].ptr)
        res_dev.download(slot, res_pin.array)
        pose_dev.download(slot, pose_pin.array)
        ctx.sync(slot)
        return j % n_batches

    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(h.local_rank) if rank == 0 else None
    h.barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    total_ms, _ = h.timed(device_step, args.steps)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    for i in range(2):
        e2e_step(i)
    e2e_ms, e2e_wall = h.timed(e2e_step, args.steps)
    e2e_ms = max(e2e_ms, e2e_wall)

    # ---- quality over every distinct frame of this rank (untimed), stage times (profiled extra steps on every rank)
    all_res = np.zeros(ND, _lib.REGISTRATION_DTYPE)
    all_pose = np.zeros(ND, _lib.POSE_DTYPE)
    for _ in range(n_batches):
        k = e2e_step(0)
        all_res[k * B:(k + 1) * B] = res_pin.array
        all_pose[k * B:(k + 1) * B] = pose_pin.array
    quality = summarize_quality(all_res, all_pose, Hs, Rs, ts)
    stages = h.profile(device_step, max(2, min(args.steps, 8)))
    fill_descriptor_bytes(stages, quality["keypoints_mean"] * B)
    # rates of the two latency-bound tail kernels (one CTA per frame): hypotheses and point evaluations per second of
    # kernel time, next to the FP32 FMA peak they would be measured against if they were throughput-bound
    fp32_peak = 148 * 128 * 2 * 1.965e9
    tail = {}
    for key, recs, n_pts, flop in (("ransac.find_homography", all_res, all_res["matches"], 30), ("ransac.pnp", all_pose, all_res["inliers"], 40)):
        if key in stages and stages[key]["ms"] > 0:
            hyp = float(recs["hypotheses"].sum()) * B / ND
            evals = float((recs["hypotheses"].astype(np.float64) * n_pts).sum()) * B / ND
            sec = stages[key]["ms"] * 1e-3
            tail[key] = {"hypotheses_per_step": hyp, "hypotheses_per_s": hyp / sec, "point_evals_per_s": evals / sec,
                         "flop_per_eval": flop, "frac_fp32_peak": evals * flop / sec / fp32_peak, "ctas": B, "threads_per_cta": 128,
                         "bound": "latency (serial f64 solves per CTA: DLT refit + LM / EPnP 12x12 Jacobi), see profiles/r2_tail_phase_times*.txt"}
    total_ms, e2e_ms = h.max_over_ranks([total_ms, e2e_ms])
    agg = h.sum_over_ranks([quality["registered"], quality["poses_found"], quality["H_err_below_5e-3"], quality["frames"]])

    if rank == 0:
        peaks, peak_src = measured_peaks()
        ms_per_step = total_ms / args.steps
        out = {
            "metric": METRIC, "value": world * B * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": pipeline_config(args, world),
            "e2e": {"value": world * B * 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(B * FRAME * FRAME) * world,
                    "d2h_bytes_per_step": int(B * (RS + PS)) * world,
                    "call": "dunk_register_frames_sharded_dev between dunk_memcpy_h2d / _d2h on pinned host buffers; the H2D of step "
                            "i + 1 is issued on a second slot while step i computes (one H2D of B frames and one D2H of the "
                            "records inside every timed step)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "measured": {"db_rows": int(db_rows), "db_rows_this_shard": len(shard), "db_build_s_this_rank": db_build_s,
                         "db_tiles_this_rank": int(n_t.value), "tile": [tw.value, th.value], "homography_fit_residual_px": fit_resid,
                         "nccl_version": int(lib.dunk_nccl_version()) if world > 1 else None},
            "quality": dict(quality, all_ranks={"registered": int(agg[0]), "poses_found": int(agg[1]), "H_err_below_5e-3": int(agg[2]),
                                                "frames": int(agg[3])}),
            "stages_ms_per_step": stages, "tail_kernels": tail,
            "roofline": pipeline_roofline(stages, ctx, clocks, peaks, peak_src, f"pipeline frames={B} scene={S}"),
        }
        if not args.no_cpu_baseline and world == 1:
            d, k, _ = shard.read_rows(0, len(shard))
            out["cpu_baseline"] = cpu_pipeline_baseline(args, d, np.stack([k["x"], k["y"]], 1), frames_pin.array, args.cpu_frames)
        print(json.dumps(out), flush=True)
    geo_close = geo
    h.close(shard, geo_close)


def pipeline_roofline(stages, ctx, clocks, peaks, peak_src, shape_key):
    top = max(stages, key=lambda s: stages[s]["ms"])
    t = stages[top]
    share = t["ms"] / sum(s["ms"] for s in stages.values())
    if top == "match.hamming_top2":
        popc = ctx.microbench_popc()                                   # Tpopc/s, measured in this run
        ach = t["alg_bytes_or_ops"] / (t["ms"] * 1e-3) / 1e9            # Gpairs/s
        peak = popc * 1e3 / 16.0
        return {"bound": "int", "kernel": top, "achieved": ach, "peak": peak, "unit": "Gpairs/s (SURVEY 8d: 16 POPC per pair)",
                "frac": ach / peak,
                # the kernel executes POPC_PER_PAIR POPC + ALU_PER_PAIR ALU-class + 3 IMAD per pair (15 XOR words compressed by
                # 7 carry-save adders; the 16th word only on the rare rows that pass the vote): fractions of the physical
                # pipes on EXECUTED instructions
                "frac_executed": ach * POPC_PER_PAIR / (popc * 1e3),
                "frac_executed_unit": "POPC pipe (%d POPC per pair executed)" % POPC_PER_PAIR,
                "frac_executed_alu": ach * ALU_PER_PAIR / (popc * 4.0 * 1e3),
                "peak_popc_tpopc_per_s": popc, "peak_clock_mhz": (clocks or {}).get("sm_mhz"),
                "peak_source": "POPC-pipe microbenchmark run by this process (148 SMs x 16 lanes/clk x SM clock); not in MEASURED_PEAKS.json",
                "traffic": measured_traffic("hamming_top2_kernel", shape_key),
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read + write, profiles/)",
                "launches_per_step": t["launches"], "ms_per_launch": t["ms"] / max(1, t["launches"]), "share_of_step": share}
    ach = t["alg_bytes_or_ops"] / (t["ms"] * 1e-3) / 1e9                # GB/s
    return {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
            "peak_source": peak_src, "traffic": None, "launches_per_step": t["launches"], "share_of_step": share}


def measured_traffic(kernel, shape_key):
    """DRAM bytes per launch of `kernel` from a committed ncu capture (profiles/r*_traffic.json), only when the capture
    was taken at this run's shape; otherwise None (the contract allows null)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            e = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel, {})
            if e.get("shape") == shape_key:
                return e.get("dram_bytes_per_launch")
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------ reference: config 5 on the CPU
def cv2_akaze():
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    return cv2, cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, cv2.KAZE_DIFF_PM_G2, (1 << 18) - 1)


def cv2_pipeline(cv2, ak, db_desc, db_pts, frame, ratio, geo):
    """the reference's CPU path: lib.rs:61-92 -> lib.rs:94-114 -> mod.rs:231-259 -> elevationdb.rs:64-104 -> mod.rs:320-369"""
    from oracle import match_oracle as mo
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
        best = (idx, dist) if best is None else mo.merge_top2([best, (idx, dist)])
    idx, dist = best
    keep = dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(ratio)
    if keep.sum() < 4:
        return None
    src = np.array([kps[i].pt for i in np.nonzero(keep)[0]], np.float32)
    dst = db_pts[idx[keep, 0]]
    H, mask = cv2.findHomography(src, dst, cv2.RANSAC, RANSAC_THR)
    if H is None or geo is None:
        return H, None
    inl = mask.ravel() > 0
    if inl.sum() < 4:
        return H, None
    obj = geo["world"](dst[inl, 0].astype(np.float64), dst[inl, 1].astype(np.float64)) - geo["origin"]
    ok, rv, tv, _ = cv2.solvePnPRansac(obj, src[inl].astype(np.float64), geo["K"], np.zeros((4, 1)), None, None, False, PNP["iter_count"],
                                       PNP["reproj_thres"], PNP["confidence"], None, cv2.SOLVEPNP_EPNP)
    return H, ((rv.ravel(), tv.ravel()) if ok else None)


def cpu_geo(size):
    """world-coordinate chain of the CPU arm: the oracle's restatement of get_world_coordinates (numpy)"""
    import synthdata
    from oracle import geo_oracle as go
    gt, (gt_e, heights) = synthdata.scene_geotransform(), synthdata.scene_dem(size)
    return {"world": lambda x, y: go.world_coordinates(x, y, gt, gt_e, heights, heights.shape[1], heights.shape[0])[0],
            "origin": synthdata.scene_origin(size), "K": synthdata.CAMERA_K}


def cpu_pipeline_baseline(args, db_desc, db_pts, frames, n_frames):
    """cv2 (the reference's OpenCV calls) on a bounded sample of THIS run's frames against the SAME full DB (the rows the
    GPU arm built, read back from HBM: identical to cv2's own by the parity tests)"""
    try:
        cv2, ak = cv2_akaze()
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"cv2 unavailable: {e}"}
    geo = cpu_geo(args.scene)
    db_pts = np.ascontiguousarray(db_pts, np.float32)
    cv2_pipeline(cv2, ak, db_desc, db_pts, frames[0], args.ratio, geo)
    t0 = time.perf_counter()
    for i in range(n_frames):
        cv2_pipeline(cv2, ak, db_desc, db_pts, frames[i % len(frames)], args.ratio, geo)
    dt = (time.perf_counter() - t0) / n_frames
    return {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
            "sample": f"{n_frames} of this run's frames through cv2 {cv2.__version__} AKAZE + BFMatcher(k=2) + findHomography(RANSAC) + "
                      f"solvePnPRansac(EPNP) against the full {db_desc.shape[0]}-row DB; {dt * 1e3:.0f} ms/frame"}


def reference_db(args):
    """the config-4 DB through the reference's own CPU path: LoD windows (box mean) -> band_merger -> cv2 AKAZE per tile"""
    from oracle import geo_oracle as go
    from oracle import lod_oracle as lo
    cv2, ak = cv2_akaze()
    scene = build_scene(args.scene)
    r, g, b, mm = config4_bands(scene)
    descs, pts = [], []
    for lod, col, row, x0, y0, s, tile in lo.lod_tiles(r, g, b, mm, 4, "area"):
        k, d = ak.detectAndCompute(tile, None)
        if d is not None and len(k):
            descs.append(d)
            pts.append(np.array([p.pt for p in k], np.float32) * np.float32(s) + np.array([x0, y0], np.float32))   # main.rs:300-301
    gray = composite_gray(scene, lambda r_, g_, b_, mm_: go.band_merger(r_, g_, b_, mm_))
    return cv2, ak, np.concatenate(descs), np.concatenate(pts), gray


def run_reference_pipeline(args):
    import synthdata
    world = args.gpus
    try:
        cv2, ak, db_desc, db_pts, gray = reference_db(args)
    except ImportError as e:
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 (OpenCV) not importable: {e}"}))
        return
    cores = os.cpu_count() or 1
    sample = max(1, min(args.frames, args.ref_frames))
    n_distinct = min(args.distinct, sample * (args.steps + 1))
    Hs, Rs, ts, _ = synthdata.config5_views(args.distinct, args.scene, 1000)
    frames = [cv2.warpPerspective(gray, Hs[i], (FRAME, FRAME), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=1)
              for i in range(n_distinct)]
    geo = cpu_geo(args.scene)
    for i in range(min(1, args.warmup)):
        cv2_pipeline(cv2, ak, db_desc, db_pts, frames[0], args.ratio, geo)
    found = poses = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(sample):
            out = cv2_pipeline(cv2, ak, db_desc, db_pts, frames[(s * sample + j) % n_distinct], args.ratio, geo)
            found += out is not None and out[0] is not None
            poses += out is not None and out[1] is not None
    total = time.perf_counter() - t0
    ms_step = total * 1e3 / args.steps                       # one reference "step" = `sample` frames, timed in full
    val = args.steps * sample / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 / u8 popcnt / f64 (OpenCV)",
        "data": "synthetic", "config": pipeline_config(args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"each timed step = {sample} of the workload's {args.frames} frames per step, run in full (no extrapolation): "
                                   f"cv2 {cv2.__version__} AKAZE + BFMatcher(k=2) + findHomography(RANSAC) + solvePnPRansac(EPNP), {cores} "
                                   f"threads, full {db_desc.shape[0]}-row DB built by cv2 from the 85 config-4 tiles; "
                                   f"{1e3 / val:.0f} ms/frame; registered {found}, poses {poses} of {args.steps * sample}"},
        "measured": {"db_rows": int(db_desc.shape[0]), "frames_per_reference_step": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ ours: config 3 (matcher only)
def run_match(args):
    """--workload match — BASELINE config 3: one query frame's descriptors against a DB of args.db_rows
    uniform-random 61-byte rows (seeded, identical whatever N is) sharded over the N ranks by contiguous
    row range; per step ONE library call: local top-2 on every shard, one ncclAllGather of the 16-byte records,
    lexicographic (distance, index) merge, ratio test.  Total work is fixed -> strong scaling.  The
    query rows are planted after the random rows of the last shard, so correctness is checkable."""
    from cubesat_apds_b200 import sharding
    h = Harness(args)
    dunk, lib, ctx, slot, _lib = h.dunk, h.lib, h.ctx, h.slot, h._lib
    check = _lib.check
    rank, world = h.rank, h.world
    group = h.shard_group()
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
    q_pin = h.pinned((nq, 64))
    q_pin.array[:] = 0
    q_pin.array[:, :61] = q
    q_dev, merged, matches, count = h.dev_buffer(nq * 64), h.dev_buffer(nq * 16), h.dev_buffer(nq * 16), h.dev_buffer(16)
    m_pin, c_pin = h.pinned((nq,), _lib.DMATCH_DTYPE), h.pinned((4,), np.int32)
    q_dev.upload(slot, q_pin.array)
    ctx.sync(slot)

    def device_step(i):
        check(lib.dunk_db_match_sharded_dev(group.handle, db.handle, slot, C.c_void_p(q_dev.ptr), nq, lo, args.ratio, C.c_void_p(merged.ptr),
                                            C.c_void_p(matches.ptr), C.c_void_p(count.ptr)))

    def e2e_step(i):
        q_dev.upload(slot, q_pin.array)
        device_step(i)
        matches.download(slot, m_pin.array)
        count.download(slot, c_pin.array[:1])
        ctx.sync(slot)

    for i in range(args.warmup):
        device_step(i)
    h.barrier()
    top2 = np.zeros(nq, dunk.TOP2_DTYPE)
    merged.download(slot, top2)
    ctx.sync(slot)
    planted_ok = bool((top2["d1"] == 0).all() and (top2["i1"] == nt + np.arange(nq)).all())
    sampler = ClockSampler(h.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    total_ms, _ = h.timed(device_step, args.steps)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    e2e_step(0)
    e2e_ms, e2e_wall = h.timed(e2e_step, args.steps)
    total_ms, e2e_ms = h.max_over_ranks([total_ms, max(e2e_ms, e2e_wall)])
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
                                      f"ncclAllGather of top-2 records + (distance, index) merge inside dunk_db_match_sharded_dev",
                          "stages": "match only", "db_rows": nt, "queries_per_frame": nq, "parallelism": f"db-shard{world}",
                          "l2": "inputs larger than L2 (DB shard %.2f GB)" % ((hi - lo) * 64 / 1e9)},
               "matcher_gpairs_per_s": ach, "planted_rows_found": planted_ok,
               "e2e": {"value": 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(nq * 64),
                       "d2h_bytes_per_step": int(nq * 16 + 4), "matches": int(c_pin.array[0])},
               "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"bound": "int", "kernel": "hamming_top2_kernel", "achieved": ach, "peak": peak,
                            "unit": "Gpairs/s (16 POPC per pair, all GPUs)", "frac": ach / peak,
                            "frac_executed": ach * POPC_PER_PAIR / (popc * 1e3 * world),
                            "frac_executed_unit": "POPC pipe (%d POPC per pair executed)" % POPC_PER_PAIR,
                            "peak_popc_tpopc_per_s": popc, "peak_clock_mhz": (clocks or {}).get("sm_mhz"),
                            "peak_source": "POPC-pipe microbenchmark run by this process; not in MEASURED_PEAKS.json",
                            "hbm_achieved_gbs": (nt * 64.0) / (ms * 1e-3) / 1e9, "traffic": None}}
        print(json.dumps(out), flush=True)
    h.close(db)


# ------------------------------------------------------------------------------------------ ours: config 2 (extraction only)
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
    h = Harness(args)
    dunk, lib, ctx, slot, _lib = h.dunk, h.lib, h.ctx, h.slot, h._lib
    check = _lib.check
    rank, world = h.rank, h.world
    B, sub = args.extract_frames, args.extract_sub
    frames = config2_frames(B)
    f_dev = h.dev_buffer(frames.nbytes)
    f_dev.upload(slot, frames)
    ctx.sync(slot)
    ws_bytes = int(lib.dunk_pipeline_workspace_bytes(ctx.handle, sub, FRAME, FRAME))
    ws = h.dev_buffer(ws_bytes)
    view = _lib.PipelineView()
    cap = 8192                                              # output rows per frame of the host-buffer call
    kps_host = np.zeros((B, cap), dtype=_lib.KEYPOINT_DTYPE)
    desc_host = np.zeros((B, cap, 61), dtype=np.uint8)
    counts = np.zeros(B, dtype=np.int32)
    n_kp = [0]

    def device_step(i):
        total = 0
        for f0 in range(0, B, sub):
            nf = min(sub, B - f0)
            check(lib.dunk_pipeline_extract_dev(ctx.handle, slot, C.c_void_p(f_dev.ptr + f0 * FRAME * FRAME), nf, FRAME, FRAME, 1,
                                                FRAME, FRAME * FRAME, 0, C.c_void_p(ws.ptr), ws_bytes, C.byref(view)))
            total += view.total_queries
        n_kp[0] = total

    def e2e_step(i):
        # the reference-facing call: host frames in, host keypoints + descriptors out (H2D and D2H inside)
        check(lib.dunk_akaze_extract_batch(ctx.handle, frames.ctypes.data, B, FRAME, FRAME, 1, FRAME, FRAME * FRAME, 0,
                                           kps_host.ctypes.data, desc_host.ctypes.data, cap, counts.ctypes.data))

    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(h.local_rank) if rank == 0 else None
    h.barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    total_ms, _ = h.timed(device_step, args.steps)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    e2e_step(0)
    _, e2e_ms = h.timed(e2e_step, args.steps)
    n_e2e = int(counts.sum())
    total_ms, e2e_ms = h.max_over_ranks([total_ms, e2e_ms])
    stages = h.profile(device_step, max(2, args.steps))
    fill_descriptor_bytes(stages, n_kp[0])
    if rank == 0:
        peaks, peak_src = measured_peaks()
        ms = total_ms / args.steps
        out = {"metric": "extract_frames_per_s", "value": world * B * 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32 stencils", "data": "synthetic",
               "config": {"workload": f"config2-extract: batch of {B} frames {FRAME}x{FRAME} u8 gray per GPU (synth(1024, 100+i)) -> AKAZE "
                                      f"detect + MLDB-486 describe only, sub-batches of {sub}",
                          "frames_per_step_per_gpu": B, "keypoints_per_frame_mean": n_kp[0] / B,
                          "parallelism": f"frame-batch dp{world}",
                          "l2": "inputs larger than L2 (%.0f MB of frames + %.1f GB scale-space workspace per sub-batch)"
                                % (frames.nbytes / 1e6, ws_bytes / 1e9)},
               "e2e": {"value": world * B * 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
                       "d2h_bytes_per_step": int(n_e2e * (_lib.KEYPOINT_DTYPE.itemsize + 61) + 4 * B),
                       "call": "dunk_akaze_extract_batch (pageable host buffers, staged through the library's pinned halves)"},
               "gpu_launches": int(launches), "clocks": clocks,
               "stages_ms_per_step": stages, "roofline": hbm_roofline(stages, peaks, peak_src)}
        if not args.no_cpu_baseline and world == 1:
            dt, ver = extract_cpu(frames, 16)
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                                   "sample": f"16 of the {B} frames through cv2 {ver} AKAZE.detectAndCompute, {dt * 1e3:.0f} ms/frame"}
        print(json.dumps(out), flush=True)
    h.close()


# ------------------------------------------------------------------------------------------ ours: config 4 (DB build)
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
    h = Harness(args)
    dunk, lib, ctx, slot, _lib = h.dunk, h.lib, h.ctx, h.slot, h._lib
    check = _lib.check
    rank, world = h.rank, h.world
    S, lods = args.build_scene, 4
    r, g, b, mm = config4_bands(build_scene(S))
    bands_dev = []
    for x in (r, g, b):
        d = h.dev_buffer(x.nbytes)
        d.upload(slot, x)
        ctx.sync(slot)
        bands_dev.append(d)
    db = dunk.feature_database.DescriptorDatabase(ctx, capacity=4_000_000)
    n, tw, th = C.c_int(0), C.c_int(0), C.c_int(0)

    def device_step(i):
        db.clear()
        check(lib.dunk_db_build_from_bands_dev(db.handle, C.c_void_p(bands_dev[0].ptr), C.c_void_p(bands_dev[1].ptr),
                                               C.c_void_p(bands_dev[2].ptr), S, S, mm.ctypes.data, lods, 0, 0, C.byref(n), C.byref(tw),
                                               C.byref(th)))

    def e2e_step(i):
        db.clear()
        db.build_from_bands(r, g, b, mm, lods)

    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(h.local_rank) if rank == 0 else None
    h.barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    # the call synchronises internally (it returns row counts), so the device time is its wall time between two syncs
    _, total_ms = h.timed(device_step, max(args.steps, 10))
    total_ms *= args.steps / max(args.steps, 10)
    launches = (ctx.launch_count - launches0) * args.steps // max(args.steps, 10)
    clocks = sampler.stop() if sampler else None
    rows, tiles = len(db), n.value
    e2e_step(0)
    _, e2e_ms = h.timed(e2e_step, args.steps)
    total_ms, e2e_ms = h.max_over_ranks([total_ms, e2e_ms])
    stages = h.profile(device_step, 2)
    if rank == 0:
        peaks, peak_src = measured_peaks()
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
                       "d2h_bytes_per_step": int(tiles * 8),
                       "call": "dunk_db_build_from_bands (pageable host bands, staged through the library's pinned ring)"},
               "gpu_launches": int(launches), "clocks": clocks,
               "stages_ms_per_step": stages, "roofline": hbm_roofline(stages, peaks, peak_src)}
        if not args.no_cpu_baseline and world == 1:
            dt, ver, _ = build_cpu((r, g, b), mm, lods, 6)
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "tiles/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                   "sample": f"5 LoD-0 tiles + the top-LoD tile: numpy band_merger + cv2 {ver} INTER_AREA + AKAZE, "
                                             f"{dt * 1e3:.0f} ms/tile"}
        print(json.dumps(out), flush=True)
    h.close(db)


def hbm_roofline(stages, peaks, peak_src):
    """roofline object for the stage with the largest device time (all extraction stages are HBM-side)"""
    top = max(stages, key=lambda s: stages[s]["ms"])
    t = stages[top]
    ach = t["alg_bytes_or_ops"] / (t["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src, "traffic": None,
            "launches_per_step": t["launches"], "ms_per_step": t["ms"],
            "share_of_step": t["ms"] / sum(x["ms"] for x in stages.values())}


# ------------------------------------------------------------------------------------------ reference: secondary workloads
def run_reference_workload(args):
    """reference arm of the secondary workloads: the same OpenCV calls on the host cores, bounded samples, every
    step timed in full"""
    cores = os.cpu_count() or 1
    base = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "vs_baseline": None, "data": "synthetic"}
    if args.workload == "extract":
        B = args.extract_frames
        frames = config2_frames(min(B, 16))
        per_step = 4
        t0 = time.perf_counter()
        dt, ver = extract_cpu(frames, per_step * args.steps)
        val, unit, metric, ms = 1.0 / dt, UNIT, "extract_frames_per_s", dt * 1e3 * per_step
        cfg = {"workload": f"config2-extract on the host CPU: cv2 {ver} AKAZE.detectAndCompute on {FRAME}x{FRAME} u8 frames",
               "frames_per_step_per_gpu": B}
        sample = f"each step = {per_step} frames timed in full, {dt * 1e3:.0f} ms/frame"
        dtype, scaling = "f32 (OpenCV)", "weak"
    elif args.workload == "build":
        r, g, b, mm = config4_bands(build_scene(args.build_scene))
        dt, ver, (tw, th) = build_cpu((r, g, b), mm, 4, 6)
        val, unit, metric, ms = 1.0 / dt, "tiles/s", "db_build_tiles_per_s", dt * 1e3 * 6 / max(1, args.steps)
        cfg = {"workload": f"config4-build on the host CPU: {args.build_scene}^2 scene, tiles of {tw}x{th}: numpy band_merger + "
                           f"cv2 {ver} INTER_AREA + AKAZE per tile", "tiles": 85}
        sample = f"5 LoD-0 tiles + the top-LoD tile in total, {dt * 1e3:.0f} ms/tile"
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
    try:
        if args.workload != "pipeline":
            return run_reference_workload(args)
        return run_reference_pipeline(args)
    except ImportError as e:
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 (OpenCV) not importable: {e}"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="query frames per step per GPU")
    ap.add_argument("--distinct", type=int, default=512, help="distinct query frames per GPU, cycled over the steps")
    ap.add_argument("--scene", type=int, default=10980, help="config-4 scene edge (pixels)")
    ap.add_argument("--ratio", type=float, default=0.8)
    ap.add_argument("--ref-frames", type=int, default=2, help="reference arm: frames per timed step (each step is timed in full)")
    ap.add_argument("--cpu-frames", type=int, default=16, help="our arm's cpu_baseline leg: frames through cv2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="pipeline", choices=["pipeline", "match", "extract", "build"],
                    help="pipeline = config 5 (default, the headline metric); match = config 3 (sharded matcher only); "
                         "extract = config 2 (extraction only); build = config 4 (reference-DB build from a scene)")
    ap.add_argument("--extract-frames", type=int, default=256, help="--workload extract: frames per step per GPU")
    ap.add_argument("--extract-sub", type=int, default=256, help="--workload extract: frames per library call (sub-batch)")
    ap.add_argument("--build-scene", type=int, default=10980, help="--workload build: scene edge (pixels)")
    ap.add_argument("--db-rows", type=int, default=50_000_000, help="--workload match: reference descriptors")
    ap.add_argument("--queries", type=int, default=3163, help="--workload match: query descriptors per frame")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    {"pipeline": run_pipeline, "match": run_match, "extract": run_extract, "build": run_build}[args.workload](args)


if __name__ == "__main__":
    main()
