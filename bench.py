#!/usr/bin/env python3
"""bench.py -- headless bench harness of the pwnfps per-pixel portal ray cast on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): primary Mrays/s (= pixels/s / 1e6; every pixel also traces two bounce
segments) and frames/s.  A "step" is one pass of the hot path over one batch of synthetic input:
the 64-pose batch of 7680x4320 frames of the synthetic 10k-sector maze (BASELINE configs[4], the
configuration north_star's target is quoted on), blur pass on as the reference ships it (defs.h:7).

  value      whole-job Mrays/s with the frames left resident in HBM (device-timed, CUDA events on
             the library's stream, max over ranks)
  e2e        the same batch through the C ABI with HOST buffers: every frame copied device->host
             inside the timed region; `e2e.single_frame_call` is the drop-in call itself
             (trace_screen_centred -> pwn_trace_screen), one frame per call, nothing pipelined
             across frames
  roofline   the dominant kernel K1 (trace + shade + pack) against the FP32 issue roofline on THIS
             workload (algorithmic lane-ops counted by the ids build on the same poses), plus its
             framebuffer stores against the measured HBM peak
  cpu_baseline  the reference's own OpenMP+SSE2 path (oracle/_ref, built from /root/reference by
             oracle/make_ref.py) timed on this box's host cores on a bounded sample
  other_configs BASELINE configs[1..3] + the game's 14 spheres, each with its own ops/pixel and
             roofline fraction

N > 1 (torchrun, one process per GPU): STRONG scaling of the same step.  Every frame of the batch
is sharded over the N GPUs in contiguous blocks of 32-row bands and assembled, blurred, in GPU 0's frame
inside the timed events (pwn_trace_shard_blur, PWN_SHARD_BLOCKS: the halo rows go to the neighbouring GPU
and the blurred rows to GPU 0 as peer stores of the trace and blur kernels themselves, the blur running
under the trace, gated by per-band counters on the device).  The
pose-sharded batch (no exchange at all) and the unblurred fused band gather are reported beside it.
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

SIZE = (7680, 4320)
NPOSES = 64
SM_COUNT = 148
FP32_LANES = 128
METRIC = "primary_mrays_per_s"
# algorithmic FP32 lane-ops of the blur pass per blurred pixel, counted from the source like the trace's (SURVEY 8d rules:
# mul, add, cmp, cvt = 1 each): 4 samples x [z = dist - 1 (1) + 2 coordinates x (randfs: cvt, mul, sub (3); x fstr, x z (2);
# int base -> float, add (2); float -> int (1))] = 4 x 17  (screen.h:95-102; the clamps and pavgb are integer work)
W_BLUR = 68


def headline_config(world, blur=True):
    """the `config` object of the JSON line: identical for our arm and the reference arm"""
    return {"workload": "synthetic 10k-sector maze (128x128 cells, 1000 portal pairs, seed 1), 7680x4320, 64-pose batched frame set "
                        "(seed 9)%s (BASELINE configs[4])" % (", blur on" if blur else ", blur off"),
            "resolution": list(SIZE), "frames_per_step": NPOSES,
            "sharding": "none" if world == 1 else "every frame in %d contiguous blocks of 32-row bands over %d GPUs, assembled on GPU 0" % (world, world),
            "l2": "every 8K frame (132.7 MB sbuf + 132.7 MB zbuf + 132.7 MB pre-blur) is larger than the 126 MB L2 and is written "
                  "once; a 256 MiB flush runs between steps, outside the per-step events"}


def headline_workload():
    from pwnfps_b200 import levels
    lv = levels.maze(dim=128, seed=1)
    return lv, levels.poses_random(lv, NPOSES, seed=9)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


# ---- the reference's CPU path (oracle/_ref) -------------------------------------------------------------
def ref_load(lv, blur=True):
    """the reference's own trace_screen_centred with `lv` loaded: the unmodified build for levels inside its
    64x64 / 26-letter limits, the mechanically widened build (pinned to the unmodified one, tests/test_oracle_ref.py)
    beyond them"""
    from oracle import refbind
    wide = lv.w > 64 or lv.h > 64 or len(lv.portals) > 26
    if wide:
        variant = "libref_wide_blur.so" if blur else "libref_wide.so"
    else:
        variant = "libref.so" if blur else "libref_noblur.so"
    ref = refbind.Ref(variant)
    if wide:
        cells = np.full((ref.dim, ref.dim), ord("."), np.int32)
        cells[:lv.h, :lv.w] = lv.cells
        ref.level_set(cells, lv.portals_i32(), lv.sx, lv.sz)
    else:
        import tempfile
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
    return ref, variant


def host_threads():
    """all the host's cores, whatever the launcher put into OMP_NUM_THREADS (torchrun sets it to 1)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_time(ref, poses, idx, w, h, nthreads):
    """seconds per pose of `idx`, each a single call of the reference's trace_screen_centred"""
    return [ref.time_frames(poses["cam"][i], float(poses["sec_current"][i]), w, h, nthreads, 1) for i in idx]


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's cores, same config."""
    if rank != 0:
        return
    from oracle import refbind
    w, h = SIZE
    if not refbind.available("libref_wide_blur.so"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_wide_blur.so not built (needs /root/reference at build time)"}))
        return
    lv, poses = headline_workload()
    ref, variant = ref_load(lv, blur=True)
    nt = host_threads()
    nsample = 4                                # bounded sample: 4 of the 64 poses per step, a different four each step
    for _ in range(max(args.warmup, 1)):
        cpu_reference_time(ref, poses, [0], w, h, nt)
    t0 = time.time()
    total = 0.0
    nframes = 0
    for s in range(args.steps):
        idx = [(s * nsample + k * (NPOSES // nsample)) % NPOSES for k in range(nsample)]
        total += sum(cpu_reference_time(ref, poses, idx, w, h, nt))
        nframes += len(idx)
    mrays = nframes * w * h / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s",
        "frames_per_s": nframes / total,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(world),
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": nt, "kind": "reference", "cpu": cpu_model(),
                         "sample": "%d of the 64 poses per step (a different four each step) x %d steps at %dx%d through the reference's own "
                                   "trace_screen_centred (%s: the reference headers widened to 128x128 cells at build time), "
                                   "OpenMP threads = all %d host cores" % (nsample, args.steps, w, h, variant, nt)},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


# ---- our arm ------------------------------------------------------------------------------------------------
class Rig:
    """what every leg of the bench needs: the tracer, its stream, timing helpers, the collective helpers"""

    def __init__(self, rank, world, local_rank):
        import torch
        from pwnfps_b200 import api
        self.torch = torch
        self.api = api
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.tr = api.Tracer(local_rank)
        self.stream = torch.cuda.ExternalStream(self.tr.stream_ptr(), device=torch.device("cuda", local_rank))
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        self.peaks = load_peaks()
        self.peak_issue = SM_COUNT * FP32_LANES * self.peaks["sm_max_mhz"] * 1e6 / 1e12          # T lane-op/s at max boost

    def l2_flush(self):
        with self.torch.cuda.stream(self.stream):
            self.flush_buf.fill_(1)

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()

    def allmax(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """per-step CUDA events on the launching stream; L2 flushed between steps outside them"""
        ms = []
        for _ in range(steps):
            self.l2_flush()
            self.barrier()
            a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            fn()
            b.record(self.stream)
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return ms

    def k1_roofline(self, poses, w, h, flags, ops_every=1):
        """K1 (and K2) launched separately on `poses`, each bracketed by events on the stream; the algorithmic
        lane-ops of the same poses counted by the ids build"""
        tr, api = self.tr, self.api
        k1 = k2 = 0.0
        for p in poses:
            tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), flags)
            a, b = tr.last_kernel_ms()
            k1 += a
            k2 += b
        ops = 0
        sample = poses[::ops_every]
        for p in sample:          # the work model is counted, not timed
            tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), api.PWN_F_IDS)
            ops += tr.last_opcount()
        n = len(poses)
        ops_per_frame = ops / len(sample)
        k1_ms, k2_ms = k1 / n, k2 / n
        achieved = ops_per_frame / (k1_ms * 1e-3) / 1e12
        return {"ops_per_pixel": ops_per_frame / (w * h), "ops_per_launch": ops_per_frame, "k1_ms_per_launch": k1_ms, "k2_ms_per_launch": k2_ms,
                "achieved": achieved, "frac": achieved / self.peak_issue,
                "ops_counted_on": "%d of the %d poses" % (len(sample), n)}


def other_configs(rig, flags):
    """BASELINE configs[1..3] and the game's 14 spheres, device-timed on rank 0: throughput through the pose
    pipeline and K1's roofline fraction on each workload's own ops/pixel.  Not the headline; no exception is
    swallowed -- a broken path fails the bench."""
    from pwnfps_b200 import levels
    api, tr, torch = rig.api, rig.tr, rig.torch
    others = {}
    dl = levels.default_level()
    fly = api.flythrough(dl, 600)
    hall = levels.mirror_hall()
    hp = np.zeros(8, api.POSE_DTYPE)
    for k in range(8):
        hp["cam"][k] = api.mat4_roty(api.cam_spawn(hall), 1.5707964 + 0.4 * k)
        hp["sec_current"][k] = 0.25 * k
    mz = levels.maze(dim=128, seed=1)
    sph = levels.game_spheres(levels.default_level(), 0.3)
    for name, lvx, (ow, oh), px, every in (
            ("default_level_flythrough_1080p (configs[1])", dl, (1920, 1080), fly, 10),
            ("mirror_hall_1080p (configs[2])", hall, (1920, 1080), hp, 1),
            ("maze_10k_4k (configs[3])", mz, (3840, 2160), levels.poses_random(mz, 16, seed=1), 2),
            ("default_level_14_spheres_1080p (game.lua:2-30)", sph, (1920, 1080), api.flythrough(sph, 8 * 20)[::20].copy(), 1)):
        tr.level_upload(lvx)
        tr.level_prepare_render(lvx)
        tr.frame_config(ow, oh, want_ids=True)
        tr.trace_poses(px, ow, oh, flags)                         # warm-up
        reps = 3 if len(px) < 100 else 1
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(rig.stream)
        for _ in range(reps):
            tr.trace_poses(px, ow, oh, flags)
        b.record(rig.stream)
        b.synchronize()
        msf = a.elapsed_time(b) / (reps * len(px))
        r = rig.k1_roofline(px[::max(1, len(px) // 60)], ow, oh, flags, ops_every=1 if len(px) < 100 else every // 10 or 1)
        others[name] = {"resolution": [ow, oh], "poses": len(px), "ms_per_frame": msf,
                        "mrays_per_s": ow * oh / msf / 1e3, "frames_per_s": 1e3 / msf, "blur": bool(flags),
                        "roofline": {"bound": "fp32_issue", "ops_per_pixel": r["ops_per_pixel"], "k1_ms_per_launch": r["k1_ms_per_launch"],
                                     "k2_ms_per_launch": r["k2_ms_per_launch"], "achieved": r["achieved"], "peak": rig.peak_issue,
                                     "unit": "Tlaneop/s", "frac": r["frac"],
                                     "timed_step_frac": r["ops_per_launch"] / (msf * 1e-3) / 1e12 / rig.peak_issue}}
    # K3 = screen_upscale (screen.h:126-149) of an 8K frame to a 16K surface and of a 4K frame to 8K: the one
    # HBM-bound kernel of the path (4 B read + 4*rscale^2 B written per frame pixel)
    tr.level_upload(dl)
    tr.level_prepare_render(dl)
    for (ow, oh, rs) in ((7680, 4320, 2), (3840, 2160, 2), (1920, 1080, 4)):
        tr.frame_config(ow, oh)
        tr.trace_poses(fly[:1], ow, oh, flags)
        tr.screen_upscale(ow, oh, rs, read=False)
        tot = 0.0
        for _ in range(10):
            rig.l2_flush()
            tr.screen_upscale(ow, oh, rs, read=False)
            tot += tr.last_upscale_ms()
        msu = tot / 10
        nbytes = ow * oh * 4 * (1 + rs * rs)
        others["screen_upscale_%dx%d_x%d (screen.h:126-149)" % (ow, oh, rs)] = {
            "ms": msu, "algorithmic_bytes": nbytes,
            "roofline": {"bound": "hbm", "achieved": nbytes / (msu * 1e-3) / 1e9, "peak": rig.peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": nbytes / (msu * 1e-3) / 1e9 / rig.peaks["hbm_gbs"], "peak_source": rig.peaks["source"]}}
    return others


def single_frame_calls(rig, lv, poses, w, h, blur, nfr):
    """the drop-in call itself, frame by frame (trace_screen_centred -> pwn_trace_screen): the frame is complete in
    the host's buffers on return, nothing is pipelined across frames.  Into the host's own buffer as malloc'd
    (pageable) and after pwn_host_register page-locked it in place, as the shim does; sbuf alone
    (PWN_SCREEN_NO_ZBUF) and sbuf + zbuf (what the reference's function fills)."""
    api, tr = rig.api, rig.tr
    out = {}
    idx = [int(i) for i in np.linspace(0, len(poses) - 1, nfr)]
    for kind in ("pageable", "registered", "registered_with_zbuf"):
        sb = np.zeros((h, w), np.uint32)
        zb = np.zeros((h, w), np.float32) if kind.endswith("zbuf") else None
        if kind != "pageable":
            api.host_register(sb)
            if zb is not None:
                api.host_register(zb)
        for i in idx[:3]:
            tr.trace_screen_centred(lv, 0, 0, w, h, poses["cam"][i], float(poses["sec_current"][i]), blur=blur, sbuf=sb, zbuf=zb)
        t0 = time.perf_counter()
        for i in idx:
            tr.trace_screen_centred(lv, 0, 0, w, h, poses["cam"][i], float(poses["sec_current"][i]), blur=blur, sbuf=sb, zbuf=zb)
        dt = time.perf_counter() - t0
        if kind != "pageable":
            api.host_unregister(sb)
            if zb is not None:
                api.host_unregister(zb)
        out[kind] = {"ms_per_frame": dt / len(idx) * 1e3, "mrays_per_s": len(idx) * w * h / dt / 1e6,
                     "d2h_bytes_per_frame": w * h * 4 * (2 if zb is not None else 1)}
    out["how"] = ("pwn_trace_screen per frame: K1 and K2 in 4 row chunks, the copy of chunk k to the host under the kernels of chunk k+1, "
                  "sync; wall clock over %d poses at %dx%d" % (len(idx), w, h))
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pwnfps_b200 import api, levels

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the trace has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rig = Rig(rank, world, local_rank)
    tr, peaks = rig.tr, rig.peaks
    w, h = SIZE
    lv, poses = headline_workload()
    tr.level_upload(lv)
    tr.level_prepare_render()
    tr.frame_config(w, h)
    blur = not args.no_blur
    flags = api.PWN_F_BLUR if blur else 0
    warm = max(args.warmup, 3)

    bases = None
    bounds = None
    shard_check = None
    if world == 1:
        def step_device():
            tr.trace_poses(poses, w, h, flags)
    else:
        from pwnfps_b200 import shard
        # the frames a single GPU produces, by checksum: what every sharded frame is compared with
        want = tr.trace_poses(poses[:4], w, h, flags, want_sums=True)
        tr.shard_config(world, rank)
        tr.shard_layout(blocks=True)          # contiguous band blocks: halo exchange instead of an all-gather
        # rows below the horizon cost more (floor bounce): the block boundaries are moved until the shards' traces take
        # the same time on a sample of the batch (outside the timed region; the boundaries are printed with the line)
        bounds = shard.balance_blocks(tr, rank, world, dist, poses[::8], w, h)
        rig.barrier()
        if blur:
            bases = shard.open_all_peers(tr, rank, world, dist)

            def one_frame(p):
                tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]))
        else:
            bases = shard.open_peer_frame(tr, rank, dist, with_z=False)

            def one_frame(p):
                tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
        got = []
        for p in poses[:4]:
            if rank == 0:
                tr.frame_clear()
            tr.sync()
            rig.barrier()
            one_frame(p)
            tr.sync()
            rig.barrier()
            if rank == 0:
                s0, _, _ = tr.read_frame(want_z=False)
                got.append(int(api.checksum_ref(s0, w, h)))
        if blur and tr.peers_timed_out():
            raise SystemExit("bench.py: a device-side wait of the sharded frames timed out on rank %d" % rank)
        if rank == 0:
            shard_check = bool(got == [int(v) for v in want])
            if not shard_check:
                raise SystemExit("bench.py: the frames assembled from %d GPUs differ from the one-GPU frames" % world)

        def step_device():
            for p in poses:
                one_frame(p)
            tr.join()          # (two frames are in flight on two streams: the step's closing event comes after both)

    # ---- value: frames stay in HBM -----------------------------------------------------
    for _ in range(warm):
        step_device()
    tr.sync()
    rig.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    launches0 = tr.launch_count()
    wall0 = time.time()
    ms = rig.timed(step_device, args.steps)
    tr.sync()
    rig.barrier()
    torch.cuda.synchronize()
    wall = rig.allmax(time.time() - wall0)
    launches = tr.launch_count() - launches0
    clocks = sampler.summary()
    dev_ms = rig.allmax(sum(ms))
    value = float(args.steps) * NPOSES * w * h / (dev_ms * 1e-3) / 1e6

    # ---- e2e: host buffers, copies inside the timed region ---------------------------------
    e2e_steps = max(2, min(args.steps, 3))
    if world == 1:
        chunk = 8          # 1.06 GB of pinned host memory, reused
        pinned = api.PinnedBuffer((chunk, h, w), np.uint32)

        def step_e2e():
            for c in range(0, NPOSES, chunk):
                tr.trace_poses(poses[c:c + chunk], w, h, flags, out=pinned.ptr)
        step_e2e()
        ms_e2e = rig.timed(step_e2e, e2e_steps)
        e2e_ms = sum(ms_e2e)
        e2e_how = ("pwn_trace_poses into pinned host memory, %d-frame chunks; the trace of frame i+1 and the blur of frame i run under "
                   "the copy of frame i-1" % chunk)
        pinned.free()
    else:
        # every rank copies its own (blurred) bands into one host frame in shared memory that all ranks have
        # page-locked: all the PCIe links at once, nothing gathered on GPU 0.  Wall clock between barriers.
        from pwnfps_b200 import shard
        host, shm = shard._shared_host_frame(rank, dist, h, w)
        api.host_register(host)

        def frame_e2e(p):
            if blur:
                tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]), api.PWN_F_LOCAL_BANDS)
            else:
                tr.frame_redirect(0, 0, 0)
                tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
            tr.read_bands(w, h, host)
        frame_e2e(poses[0])
        rig.barrier()
        if rank == 0 and int(api.checksum_ref(host, w, h)) != int(want[0]):
            raise SystemExit("bench.py: the frame assembled in host memory differs from the one-GPU frame")
        e2e_ms = 0.0
        for _ in range(e2e_steps):
            rig.barrier()
            t0 = time.perf_counter()
            for p in poses:
                frame_e2e(p)
            rig.barrier()
            e2e_ms += (time.perf_counter() - t0) * 1e3
        e2e_ms = rig.allmax(e2e_ms)
        e2e_how = ("per frame: pwn_trace_shard_blur with PWN_F_LOCAL_BANDS, then every rank copies its own bands into one page-locked "
                   "host frame in shared memory (pwn_read_bands); wall clock between barriers, max over ranks")
        tr.sync()
        api.host_unregister(host)
        del host
        shm.close()
        rig.barrier()
        if rank == 0:
            shm.unlink()
    e2e_value = float(e2e_steps) * NPOSES * w * h / (e2e_ms * 1e-3) / 1e6

    # ---- N > 1: the alternatives beside the headline, then back to one shard ----------------------------
    beside = None
    if world > 1:
        from pwnfps_b200 import shard
        if blur:
            timed_out = tr.peers_timed_out()
            shard.close_all_peers(tr, bases)
            if timed_out:
                raise SystemExit("bench.py: a device-side barrier of the sharded frames timed out")
        else:
            shard.close_peer_frame(tr, bases)
        tr.shard_blocks(None)
        tr.shard_layout(blocks=False)
        tr.shard_config(1, 0)
        beside = shard.bench_band_gather(tr, rank, world, local_rank, rig.stream)
        tr.shard_config(1, 0)
        tr.level_upload(lv)
        tr.level_prepare_render()
        tr.frame_config(w, h)

    single = roofline = cpu = others = None
    if rank == 0:
        # ---- the drop-in call, one frame per call ---------------------------------------------------
        if world == 1:
            single = {"8k_maze": single_frame_calls(rig, lv, poses, w, h, blur, 16)}
            dl = levels.default_level()
            tr.level_upload(dl)
            tr.level_prepare_render(dl)
            tr.frame_config(1920, 1080)
            single["1080p_default_level"] = single_frame_calls(rig, dl, api.flythrough(dl, 600), 1920, 1080, blur, 100)
            tr.level_upload(lv)
            tr.level_prepare_render(lv)

        # ---- roofline of K1 on this workload, live: per-launch events over the same 64 poses ------------------
        tr.frame_config(w, h, want_ids=True)
        r = rig.k1_roofline(poses, w, h, flags, ops_every=4)
        step_ms_frame = sum(ms) / args.steps / NPOSES
        sm_mhz = clocks["sm_mhz"] or peaks["sm_max_mhz"]
        peak_measured = tr.fp32_peak()      # non-fused FMUL/FADD chains, measured now on this GPU
        store_bytes = 8.0 * w * h           # 4 B sbuf + 4 B zbuf per pixel
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k_trace_traffic.json")
        if os.path.isfile(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {
            "bound": "fp32_issue", "kernel": "k_trace (K1: ray gen + walk + shade + bounces + pack)",
            "achieved": r["achieved"], "peak": rig.peak_issue, "unit": "Tlaneop/s", "frac": r["frac"],
            "peak_how": "148 SMs x 128 FP32 lanes x %.0f MHz max boost (1 lane-op per slot; FMA-counted peak is 2x)" % peaks["sm_max_mhz"],
            "frac_at_measured_clock": r["achieved"] / (SM_COUNT * FP32_LANES * sm_mhz * 1e6 / 1e12),
            "peak_fma": 2.0 * rig.peak_issue, "frac_of_fma_peak": r["achieved"] / (2.0 * rig.peak_issue),   # both peaks
            "peak_measured": peak_measured, "frac_of_measured_peak": r["achieved"] / peak_measured,
            "peak_measured_how": "k_fp32_peak: 8 independent non-fused FMUL+FADD chains per thread, 148x8 CTAs x 256 threads, best of 4",
            "ops_per_pixel": r["ops_per_pixel"], "ops_counted_on": r["ops_counted_on"],
            "k1_ms_per_launch": r["k1_ms_per_launch"], "k2_ms_per_launch": r["k2_ms_per_launch"],
            "k1_share_of_step": (r["k1_ms_per_launch"] / step_ms_frame) if world == 1 else None,
            "step_ms_per_frame": step_ms_frame,
            "note": "in the timed steps at N = 1 the blur of frame i-1 rides inside the K1 launch of frame i (pose pipeline); "
                    "k1/k2_ms_per_launch are the two kernels launched separately on the same 64 poses on one GPU",
            "traffic": traffic,
            # the kernel of the timed region itself (trace of frame i + blur of frame i-1 in one launch), timed by the
            # step events: the same algorithmic trace work over the whole launch, the blur's own work not counted
            "timed_step": ({"kernel": "k_trace<..., BLURP> (K1 + the blur of the previous frame)", "ms_per_launch": step_ms_frame,
                            "achieved": r["ops_per_launch"] / (step_ms_frame * 1e-3) / 1e12,
                            "frac": r["ops_per_launch"] / (step_ms_frame * 1e-3) / 1e12 / rig.peak_issue,
                            # the same launch with the blur's own lane-ops counted too (the launch does both)
                            "blur_ops_per_pixel": W_BLUR if blur else 0,
                            "frac_with_blur_ops": (r["ops_per_launch"] + (W_BLUR if blur else 0) * w * h) / (step_ms_frame * 1e-3) / 1e12 / rig.peak_issue}
                           if world == 1 else None),
            "hbm": {"bound": "hbm", "achieved": store_bytes / (r["k1_ms_per_launch"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": store_bytes / (r["k1_ms_per_launch"] * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["source"],
                    "algorithmic_bytes_per_pixel": 8},
        }
        tr.frame_config(64, 64)          # drop the 1.6 GB ids buffer

        # ---- CPU baseline: the reference's own code on this box's cores (rank 0, N == 1 only) ----
        if world == 1 and not args.no_cpu:
            from oracle import refbind
            variant = "libref_wide_blur.so" if blur else "libref_wide.so"
            if refbind.available(variant):
                ref, variant = ref_load(lv, blur=blur)
                nt = host_threads()
                cpu_reference_time(ref, poses, [0], w, h, nt)      # warm-up frame discarded
                idx = [5, 21, 37, 53]
                secs = cpu_reference_time(ref, poses, idx, w, h, nt)
                secs1 = cpu_reference_time(ref, poses, idx[:1], 1920, 1080, 1)
                cpu = {"value": len(idx) * w * h / sum(secs) / 1e6, "unit": "Mrays/s", "cores": nt, "kind": "reference", "cpu": cpu_model(),
                       "sample": "poses 5, 21, 37, 53 of the 64-pose batch at %dx%d through the reference's own trace_screen_centred "
                                 "(%s: the reference headers widened to 128x128 cells at build time), warm-up frame discarded" % (w, h, variant),
                       "one_thread_value": 1920 * 1080 / sum(secs1) / 1e6, "one_thread_sample": "pose 5 at 1920x1080, 1 OpenMP thread"}
            else:
                cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/%s not built" % variant}

        # ---- the other BASELINE configs (reported, not the headline) --------
        if world == 1 and not args.no_others:
            others = other_configs(rig, flags)

        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s",
            "frames_per_s": value * 1e6 / (w * h),
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": headline_config(world, blur),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": NPOSES * api.POSE_DTYPE.itemsize,
                    "d2h_bytes_per_step": NPOSES * w * h * 4, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "how": e2e_how, "single_frame_call": single},
            "gpu_launches": int(launches), "wall_s": wall,
        }
        if shard_check is not None:
            line["sharded_frames_equal_one_gpu"] = shard_check
            line["block_boundaries_bands"] = [int(b) for b in bounds]      # (beside `config`, which both arms print identically)
        if others is not None:
            line["other_configs"] = others
        if beside is not None:
            line["band_gather"] = beside
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
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
