#!/usr/bin/env python3
"""bench.py -- headless bench harness of the pwnfps per-pixel portal ray cast on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): primary Mrays/s (= pixels/s / 1e6; every pixel also traces two
bounce segments) and frames/s.  A "step" is one pass of the hot path over one batch of
synthetic input: the 600-frame scripted camera fly-through of the default level at 1920x1080
(BASELINE configs[1]), blur pass on as the reference ships it (defs.h:9).

  value      whole-job Mrays/s with frames left resident in HBM (device-timed, CUDA events
             on the library's stream, max over ranks)
  e2e        the same batch through the C ABI with HOST buffers: pwn_trace_poses() into pinned
             host memory, device->host copy of every frame inside the timed region
  roofline   the dominant kernel K1 (trace+shade+pack) against the FP32 issue roofline, plus
             its framebuffer stores against the measured HBM peak
  cpu_baseline  the reference's own OpenMP+SSE2 path (oracle/_ref, built from /root/reference
             by oracle/make_ref.py) timed on this box's host cores on a bounded sample

N > 1: one process per GPU (torchrun).  The path shards by camera pose with no data-path
collective: rank r traces its own 600-frame slice of a 600*N-frame fly-through
("scaling": "weak").  The row-band sharded single-frame mode with the bands gathered into
GPU 0's frame is measured beside it and reported under "band_gather".
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

FRAMES = 600
SIZE = (1920, 1080)
SM_COUNT = 148
FP32_LANES = 128


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed (-lms 100) during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.3)          # let the first sample land before the timed region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                self.rows.append(f)

    def mark(self):
        self.first = len(self.rows)

    def summary(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        rows = self.rows[getattr(self, "first", 0):] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons, "samples": len(rows)}


def cpu_reference_time(lv, poses, idx, w, h, variant="libref.so", nthreads=0, repeat=1):
    """The reference's own trace_screen_centred (oracle/_ref) on the host cores: seconds per frame list."""
    import tempfile
    from oracle import refbind
    ref = refbind.Ref(variant)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "level.txt")
        lv.save_txt(p)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)          # level_load prints its portal table (level.h:220,225)
        try:
            ref.level_load(p)
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
    nt = nthreads or ref.max_threads()
    ref.time_frames(poses["cam"][idx[0]], float(poses["sec_current"][idx[0]]), w, h, nt, 1)   # warm-up frame discarded
    secs = []
    for i in idx:
        secs.append(ref.time_frames(poses["cam"][i], float(poses["sec_current"][i]), w, h, nt, repeat) / repeat)
    return secs, nt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's cores."""
    if rank != 0:
        return
    from pwnfps_b200 import api, levels
    from oracle import refbind
    w, h = SIZE
    lv = levels.default_level()
    poses = api.flythrough(lv, FRAMES)
    idx = list(range(0, FRAMES, 100))          # bounded sample: 6 of the 600 poses per step
    kind = "reference" if refbind.available("libref.so") else "port"
    if kind != "reference":
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so not built (needs /root/reference at build time)"}))
        return
    for _ in range(max(args.warmup, 1)):
        cpu_reference_time(lv, poses, idx[:1], w, h)
    t0 = time.time()
    per = []
    nt = 1
    for _ in range(args.steps):
        secs, nt = cpu_reference_time(lv, poses, idx, w, h)
        per.append(sum(secs))
    total = sum(per)
    mrays = args.steps * len(idx) * w * h / total / 1e6
    line = {
        "impl": "reference", "metric": "primary_mrays_per_s", "value": mrays, "unit": "Mrays/s",
        "frames_per_s": args.steps * len(idx) / total,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "default level, 1920x1080, scripted 600-frame fly-through, blur on (BASELINE configs[1])",
                   "sample": "%d of the 600 poses per step (every 100th)" % len(idx), "resolution": [w, h]},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": nt, "kind": kind,
                         "sample": "%d frames/step x %d steps of the fly-through at %dx%d, OpenMP threads = all host cores" % (len(idx), args.steps, w, h)},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pwnfps_b200 import api, levels

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the trace has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w, h = SIZE
    peaks = load_peaks()
    lv = levels.default_level()
    tr = api.Tracer(local_rank)
    tr.level_upload(lv)
    tr.level_prepare_render()
    tr.frame_config(w, h, want_ids=True)
    all_poses = api.flythrough(lv, FRAMES * world)
    poses = all_poses[rank * FRAMES:(rank + 1) * FRAMES].copy()
    flags = 0 if args.no_blur else api.PWN_F_BLUR

    stream = torch.cuda.ExternalStream(tr.stream_ptr(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def l2_flush():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    def step_device():
        tr.trace_poses(poses, w, h, flags)

    chunk = 150   # 1.24 GB of pinned host memory; fewer pipeline drains per step
    pinned = api.PinnedBuffer((chunk, h, w), np.uint32)

    def step_e2e():
        for c in range(0, FRAMES, chunk):
            tr.trace_poses(poses[c:c + chunk], w, h, flags, out=pinned.ptr)

    def timed(fn, steps):
        """per-step CUDA events on the launching stream; L2 flushed between steps outside them"""
        ms = []
        for _ in range(steps):
            l2_flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return ms

    # ---- value: frames stay in HBM -----------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    tr.sync()
    barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    launches0 = tr.launch_count()
    wall0 = time.time()
    ms = timed(step_device, args.steps)
    barrier()
    torch.cuda.synchronize()
    wall = time.time() - wall0
    launches = tr.launch_count() - launches0
    clocks = sampler.summary()
    dev_ms = allmax(sum(ms))
    wall = allmax(wall)
    pixels_job = float(args.steps) * FRAMES * w * h * world
    value = pixels_job / (dev_ms * 1e-3) / 1e6

    # ---- e2e: host buffers, copies inside the timed region ---------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.time()
    ms_e2e = timed(step_e2e, e2e_steps)
    barrier()
    torch.cuda.synchronize()
    e2e_ms = allmax(sum(ms_e2e))
    e2e_value = float(e2e_steps) * FRAMES * w * h * world / (e2e_ms * 1e-3) / 1e6

    # ---- the drop-in call itself, frame by frame (trace_screen_centred -> pwn_trace_screen): no pipelining
    # across frames, the frame is complete in the host's buffer on return.  Into the host's own malloc'd
    # buffer as it is (pageable) and after pwn_host_register page-locked it in place, as the shim does.
    single = None
    if rank == 0 and world == 1:
        single = {}
        nfr = 150
        for kind in ("pageable", "registered"):
            hostbuf = np.zeros((h, w), np.uint32)
            if kind == "registered":
                api.host_register(hostbuf)
            for i in range(5):
                tr.trace_screen_centred(lv, 0, 0, w, h, poses["cam"][i], float(poses["sec_current"][i]), blur=bool(flags), sbuf=hostbuf)
            t0 = time.perf_counter()
            for i in range(0, FRAMES, FRAMES // nfr):
                tr.trace_screen_centred(lv, 0, 0, w, h, poses["cam"][i], float(poses["sec_current"][i]), blur=bool(flags), sbuf=hostbuf)
            dt = time.perf_counter() - t0
            if kind == "registered":
                api.host_unregister(hostbuf)
            single[kind] = {"ms_per_frame": dt / nfr * 1e3, "mrays_per_s": nfr * w * h / dt / 1e6}
        single["how"] = "pwn_trace_screen per frame (K1, K2, D2H of sbuf, sync), wall clock over %d of the %d poses" % (nfr, FRAMES)

    # ---- roofline of K1, live: per-launch events over the same 600 poses ---------------------
    # (single-frame path: K1 then K2 as two launches, each bracketed by events on the stream)
    k1 = k2 = 0.0
    for i in range(FRAMES):
        tr.trace_cam(poses["cam"][i], w, h, float(poses["sec_current"][i]), flags)
        a, b = tr.last_kernel_ms()
        k1 += a
        k2 += b
    ops = 0
    nsample = 0
    for i in range(0, FRAMES, 10):          # the work model is counted, not timed: every 10th pose
        tr.trace_cam(poses["cam"][i], w, h, float(poses["sec_current"][i]), api.PWN_F_IDS)
        ops += tr.last_opcount()
        nsample += 1
    ops_per_frame = ops / nsample
    k1_ms, k2_ms = k1 / FRAMES, k2 / FRAMES
    step_ms_frame = sum(ms) / args.steps / FRAMES
    sm_mhz = clocks["sm_mhz"] or peaks["sm_max_mhz"]
    peak_issue = SM_COUNT * FP32_LANES * peaks["sm_max_mhz"] * 1e6 / 1e12          # T lane-op/s at max boost
    try:
        peak_measured = tr.fp32_peak()      # non-fused FMUL/FADD chains, measured now on this GPU
    except Exception:
        peak_measured = None
    achieved = ops_per_frame / (k1_ms * 1e-3) / 1e12
    store_bytes = 8.0 * w * h                                                        # 4 B sbuf + 4 B zbuf per pixel
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k_trace_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {
        "bound": "fp32_issue", "kernel": "k_trace (K1: ray gen + walk + shade + bounces + pack)",
        "achieved": achieved, "peak": peak_issue, "unit": "Tlaneop/s", "frac": achieved / peak_issue,
        "peak_how": "148 SMs x 128 FP32 lanes x %.0f MHz max boost (1 lane-op per slot; FMA-counted peak is 2x)" % peaks["sm_max_mhz"],
        "frac_at_measured_clock": achieved / (SM_COUNT * FP32_LANES * sm_mhz * 1e6 / 1e12),
        "peak_fma": 2.0 * peak_issue, "frac_of_fma_peak": achieved / (2.0 * peak_issue),   # both peaks
        "peak_measured": peak_measured, "frac_of_measured_peak": (achieved / peak_measured) if peak_measured else None,
        "peak_measured_how": "k_fp32_peak: 8 independent non-fused FMUL+FADD chains per thread, 148x8 CTAs x 256 threads, best of 4",
        "ops_per_pixel": ops_per_frame / (w * h), "k1_ms_per_launch": k1_ms, "k2_ms_per_launch": k2_ms,
        "k1_share_of_step": k1_ms / step_ms_frame if step_ms_frame > 0 else None,
        "step_ms_per_frame": step_ms_frame,
        "note": "in the timed steps the blur of frame i-1 rides inside the K1 launch of frame i (pose pipeline); "
                "k1/k2_ms_per_launch are the two kernels launched separately on the same poses",
        "traffic": traffic,
        # the kernel of the timed region itself (trace of frame i + blur of frame i-1 in one launch), timed by the
        # step events: the same algorithmic trace work over the whole launch, the blur's own work not counted
        "timed_step": {"kernel": "k_trace<..., BLURP> (K1 + the blur of the previous frame)", "ms_per_launch": step_ms_frame,
                       "achieved": (ops_per_frame / (step_ms_frame * 1e-3) / 1e12) if step_ms_frame > 0 else None,
                       "frac": (ops_per_frame / (step_ms_frame * 1e-3) / 1e12 / peak_issue) if step_ms_frame > 0 else None},
        "hbm": {"bound": "hbm", "achieved": store_bytes / (k1_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": store_bytes / (k1_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["source"],
                "algorithmic_bytes_per_pixel": 8},
    }

    # ---- CPU baseline: the reference's own code on this box's cores (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import refbind
            if refbind.available("libref.so"):
                idx = [0, 150, 300, 450]
                secs, nt = cpu_reference_time(lv, poses, idx, w, h, "libref.so" if flags else "libref_noblur.so")
                secs1, _ = cpu_reference_time(lv, poses, idx[:2], w, h, "libref.so" if flags else "libref_noblur.so", nthreads=1)
                cpu = {"value": len(idx) * w * h / sum(secs) / 1e6, "unit": "Mrays/s", "cores": nt, "kind": "reference",
                       "sample": "4 of the 600 fly-through poses at %dx%d, warm-up frame discarded" % (w, h),
                       "one_thread_value": 2 * w * h / sum(secs1) / 1e6}
        except Exception as e:     # the baseline is a reported number, never a gate
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}

    # ---- the other BASELINE configs, device-timed on rank 0 (reported, not the headline) --------
    others = None
    if rank == 0 and world == 1 and not args.no_others:
      try:
        others = {}
        for name, lvx, (ow, oh), nposes in (("default_level_14_spheres_1080p (game.lua:2-30)", levels.game_spheres(levels.default_level(), 0.3), (1920, 1080), 8),
                                            ("mirror_hall_1080p (configs[2])", levels.mirror_hall(), (1920, 1080), 8),
                                            ("maze_10k_4k (configs[3])", levels.maze(dim=128, seed=1), (3840, 2160), 16),
                                            ("maze_10k_8k (configs[4], 1 GPU)", levels.maze(dim=128, seed=1), (7680, 4320), 16)):
            tr.level_upload(lvx)
            tr.level_prepare_render()
            tr.frame_config(ow, oh)
            if "spheres" in name:
                px = api.flythrough(lvx, 8 * 20)[::20].copy()      # poses around the spawn, where the spheres are
            elif "hall" in name:
                cam0 = api.cam_spawn(lvx)
                px = np.zeros(nposes, api.POSE_DTYPE)
                for k in range(nposes):
                    px["cam"][k] = api.mat4_roty(cam0, 1.5707964 + 0.4 * k)
                    px["sec_current"][k] = 0.25 * k
            else:
                px = levels.poses_random(lvx, nposes, seed=1)
            tr.trace_poses(px, ow, oh, flags)                         # warm-up
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(3):
                tr.trace_poses(px, ow, oh, flags)
            b.record(stream)
            b.synchronize()
            msf = a.elapsed_time(b) / (3 * nposes)
            others[name] = {"resolution": [ow, oh], "poses": nposes, "ms_per_frame": msf,
                            "mrays_per_s": ow * oh / msf / 1e3, "frames_per_s": 1e3 / msf, "blur": bool(flags)}
        # K3 = screen_upscale (screen.h:126-149) of the last 8K frame to a 16K surface and of a 4K frame
        # to 8K: the one HBM-bound kernel of the path (4 B read + 4*rscale^2 B written per frame pixel)
        for (ow, oh, rs) in ((7680, 4320, 2), (3840, 2160, 2), (1920, 1080, 4)):
            tr.frame_config(ow, oh)
            tr.trace_poses(px[:1], ow, oh, flags)
            tr.screen_upscale(ow, oh, rs, read=False)
            tot = 0.0
            for _ in range(10):
                l2_flush()
                tr.screen_upscale(ow, oh, rs, read=False)
                tot += tr.last_upscale_ms()
            msu = tot / 10
            nbytes = ow * oh * 4 * (1 + rs * rs)
            others["screen_upscale_%dx%d_x%d (screen.h:126-149)" % (ow, oh, rs)] = {
                "ms": msu, "algorithmic_bytes": nbytes,
                "roofline": {"bound": "hbm", "achieved": nbytes / (msu * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": nbytes / (msu * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["source"]}}
        tr.level_upload(lv)
        tr.level_prepare_render()
        tr.frame_config(w, h, want_ids=True)
      except Exception as e:
        others = {"error": "%s: %s" % (type(e).__name__, e)}

    band = None
    if world > 1:
        from pwnfps_b200 import shard
        try:        # a side measurement must never take the headline line down with it
            band = shard.bench_band_gather(tr, rank, world, local_rank, stream)
        except Exception as e:
            band = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        line = {
            "metric": "primary_mrays_per_s", "value": value, "unit": "Mrays/s",
            "frames_per_s": value * 1e6 / (w * h),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "default level, 1920x1080, scripted 600-frame fly-through%s (BASELINE configs[1])"
                       % (", blur on" if flags else ", blur off"),
                       "resolution": [w, h], "frames_per_step": FRAMES, "sharding": "pose" if world > 1 else "none",
                       "l2": "256 MiB flush between steps, outside the per-step events; frames are written once and never re-read"},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": FRAMES * api.POSE_DTYPE.itemsize,
                    "d2h_bytes_per_step": FRAMES * w * h * 4, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "how": "pwn_trace_poses into pinned host memory, %d-frame chunks; trace of frame i+1 overlaps the copy of frame i" % chunk,
                    "single_frame_call": single},
            "gpu_launches": int(launches), "wall_s": wall,
        }
        if others is not None:
            line["other_configs"] = others
        if band is not None:
            line["band_gather"] = band
        print(json.dumps(line))
    pinned.free()
    tr.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-blur", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs block")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
