"""pwnfps_b200.shard -- multi-GPU sharding of one frame by interleaved 32-row bands.

The reference parallelises a frame over 32-row bands (`#pragma omp parallel for` over
`cy += 32`, screen.h:63-67); pixels are independent, so band b goes to GPU b % N and each GPU
holds the whole (tiny) level.  Two ways to assemble the frame on GPU 0:

  fused   K1 on every rank stores its bands straight into GPU 0's frame, opened through CUDA
          IPC and reached over NVLink (pwn_frame_ipc_* + pwn_frame_redirect): compute and
          gather are one kernel, the transfer overlaps the trace row by row.
  gather  each rank traces into its own frame; the owned rows are packed and gathered with one
          torch.distributed gather (NCCL on GPUs, gloo in the CPU tests).

The blur pass (screen.h:69-123) reads the whole pre-blur frame, so it runs on GPU 0 after
either assembly.  Pose batches shard by pose with no exchange at all (bench.py).
"""
import time

import numpy as np

BAND = 32   # screen.h:64


def owned_bands(dimy, nshards, shard):
    return [b for b in range((dimy + BAND - 1) // BAND) if b % nshards == shard]


def owned_rows(dimy, nshards, shard):
    rows = []
    for b in owned_bands(dimy, nshards, shard):
        rows.extend(range(b * BAND, min(dimy, (b + 1) * BAND)))
    return np.asarray(rows, dtype=np.int64)


def max_owned_rows(dimy, nshards):
    return max(len(owned_rows(dimy, nshards, r)) for r in range(nshards))


def gather_frame(local_frame, dimy, rank, world, dist, device=None):
    """local_frame: (H, W) torch tensor in which this rank's owned rows are valid.
    Returns the stitched (H, W) tensor on rank 0, None elsewhere.  One gather of padded band packs;
    whole bands move as strided views (band b of rank r is row block b*world + r)."""
    import torch
    W = local_frame.shape[1]
    nfull = dimy // BAND                      # whole bands; a ragged last band is handled below
    tail_owner = nfull % world if dimy % BAND else -1
    bands = local_frame[:nfull * BAND].view(nfull, BAND, W)
    nmax = (nfull + world - 1) // world + (1 if tail_owner >= 0 else 0)
    pack = torch.zeros((nmax, BAND, W), dtype=local_frame.dtype, device=local_frame.device)
    mine = bands[rank::world]
    pack[:mine.shape[0]] = mine
    if rank == tail_owner:
        pack[mine.shape[0], :dimy - nfull * BAND] = local_frame[nfull * BAND:dimy]
    if rank == 0:
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist.gather(pack, parts, dst=0)
        out = torch.zeros_like(local_frame)
        ob = out[:nfull * BAND].view(nfull, BAND, W)
        for r in range(world):
            n = ob[r::world].shape[0]
            ob[r::world] = parts[r][:n]
            if r == tail_owner:
                out[nfull * BAND:dimy] = parts[r][n, :dimy - nfull * BAND]
        return out
    dist.gather(pack, None, dst=0)
    return None


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory, for torch.as_tensor."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def frame_tensors(tr):
    """torch views (no copy) of the tracer's device sbuf/zbuf."""
    import torch
    sb, zb, _ = tr.device_ptrs()
    dev = torch.device("cuda", tr.device)
    s = torch.as_tensor(_DevArray(sb, (tr.rheight, tr.rwidth), "<i4"), device=dev)
    z = torch.as_tensor(_DevArray(zb, (tr.rheight, tr.rwidth), "<f4"), device=dev)
    return s, z


def open_peer_frame(tr, rank, dist, with_z=True):
    """Fused path set-up: every rank > 0 redirects its K1 output into rank 0's frame.  zbuf is only
    needed on GPU 0 when the blur pass follows; without it half the NVLink bytes stay home."""
    objs = [None]
    if rank == 0:
        objs = [tr.ipc_export()]
    dist.broadcast_object_list(objs, src=0)
    if rank == 0:
        return None
    handle, so, zo, to = objs[0]
    base = tr.ipc_import(handle)
    tr.frame_redirect(base + so, (base + zo) if with_z else 0, base + to)
    return base


def close_peer_frame(tr, base):
    tr.frame_redirect(0, 0, 0)
    if base is not None:
        tr.ipc_close(base)


def open_all_peers(tr, rank, world, dist):
    """Sharded frame with the blur pass: every rank opens every other rank's frame allocation
    (CUDA IPC) and hands the list to pwn_frame_peers.  Returns the imported bases to close later."""
    handle = tr.ipc_export()[0]
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle))
    bases = [0] * world
    for r in range(world):
        if r != rank:
            bases[r] = tr.ipc_import(handles[r])
    tr.frame_peers(bases)
    dist.barrier()      # pwn_frame_peers restarts the frame count of the device barriers: nobody starts a frame before all have
    return bases


def close_all_peers(tr, bases):
    tr.frame_peers(None)
    for b in bases:
        if b:
            tr.ipc_close(b)


def rebalance(bounds, times, nb):
    """One round of the block balancing: `times[s]` is what shard s took for the bands [bounds[s], bounds[s+1]).  Each
    block's cost is spread evenly over its bands, and the new boundaries are put where the cumulative cost is split
    evenly (nearest band; every shard keeps at least one band).  Pure: every rank computes the same from the same times."""
    world = len(bounds) - 1
    cost = np.zeros(nb)
    for s in range(world):
        n = bounds[s + 1] - bounds[s]
        if n > 0:
            cost[bounds[s]:bounds[s + 1]] = float(times[s]) / n
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    new = [0]
    for s in range(1, world):
        target = cum[-1] * s / world
        b = int(np.searchsorted(cum, target))
        if b > 0 and abs(cum[b - 1] - target) < abs(cum[min(b, nb)] - target):
            b -= 1
        new.append(min(max(b, new[-1] + 1), nb - (world - s)))
    new.append(nb)
    return new


def balance_blocks(tr, rank, world, dist, poses, w, h, rounds=3):
    """Block layout: move the block boundaries until the shards' traces take the same time.  Rows differ in cost (below
    the horizon the floor carries the wavy-normal bounce and the longer secondary rays), so an even split leaves the
    lower shards ~30% more work.  Each round every rank times K1 on its own block for the sample poses, the times are
    all-gathered, each block's cost is spread evenly over its bands, and the boundaries are put where the cumulative
    cost is split evenly.  Returns the boundaries (world + 1 ints), set on `tr`; every rank computes the same."""
    import torch
    nb = (h + BAND - 1) // BAND
    bounds = [s * nb // world for s in range(world + 1)]
    tr.shard_layout(blocks=True)
    tr.shard_config(world, rank)
    for _ in range(rounds):
        tr.shard_blocks(bounds)
        t = 0.0
        for p in poses:
            tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
            t += tr.last_kernel_ms()[0]
        times = torch.zeros(world, dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        times[rank] = t
        dist.all_reduce(times)
        bounds = rebalance(bounds, times.cpu().numpy(), nb)
    tr.shard_blocks(bounds)
    return bounds


def _shared_host_frame(rank, dist, h, w):
    """one (h, w) uint32 frame in POSIX shared memory, mapped by every rank of the box"""
    from multiprocessing import shared_memory
    name = [None]
    shm = None
    if rank == 0:
        shm = shared_memory.SharedMemory(create=True, size=h * w * 4)
        name = [shm.name]
    dist.broadcast_object_list(name, src=0)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=name[0])
        try:        # the creator unlinks; keep this process's resource tracker out of it
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    return np.ndarray((h, w), np.uint32, buffer=shm.buf), shm


def bench_band_gather(tr, rank, world, local_rank, stream, size=(7680, 4320), frames=16):
    """BASELINE configs[4]: one 8K frame of the 10k-sector maze, row bands over `world` GPUs,
    assembled on GPU 0; device-timed, max over ranks.  Returns a dict on every rank."""
    import torch
    import torch.distributed as dist
    from . import api, levels

    w, h = size
    lv = levels.maze(dim=128, seed=1)
    tr.level_upload(lv)
    tr.level_prepare_render()
    tr.frame_config(w, h)
    poses = levels.poses_random(lv, frames, seed=5)
    out = {"workload": "10k-sector maze, %dx%d, %d seeded poses, 32-row bands b %% N == rank (BASELINE configs[4])" % (w, h, frames)}

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # single-GPU time of the same frames, for the speed-up; rank 0 keeps their checksums
    tr.shard_config(1, 0)
    one = 0.0
    for p in poses:
        tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
        one += tr.last_kernel_ms()[0]
    one = allmax(one) / frames
    want = tr.trace_poses(poses, w, h, 0, want_sums=True) if rank == 0 else None

    # fused: K1 stores into GPU 0's frame over NVLink.  Twice: sbuf + zbuf (what a blur on GPU 0
    # needs), and sbuf alone (the frame the host reads; zbuf stays on the tracing GPU)
    tr.shard_config(world, rank)
    fused = fused_s = None
    if True:
        for with_z in (True, False):
            base = open_peer_frame(tr, rank, dist, with_z=with_z)
            got = []
            for it in range(2):
                tot = 0.0
                for k, p in enumerate(poses):
                    check = it == 1 and k % 4 == 0      # every 4th frame is read back and compared
                    if rank == 0 and check:
                        tr.frame_clear()
                        tr.sync()
                    dist.barrier()
                    tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
                    tot += tr.last_kernel_ms()[0]
                    if check:
                        dist.barrier()      # every rank's stores have landed in GPU 0's frame
                        if rank == 0:
                            s0, _, _ = tr.read_frame(want_z=False)
                            got.append(int(api.checksum_ref(s0, w, h)))
                if with_z:
                    fused = allmax(tot) / frames
                else:
                    fused_s = allmax(tot) / frames
            dist.barrier()
            close_peer_frame(tr, base)
            if rank == 0:
                key = "fused_frames_equal_one_gpu" if with_z else "fused_sbuf_only_frames_equal_one_gpu"
                out[key] = bool(got == [int(v) for v in want[::4]])

    # the same frames WITH the blur pass: all-gather fused into K1, device-side barriers, every
    # shard blurs its own bands into GPU 0's frame (pwn_trace_shard_blur); against it, the blur of
    # the gathered frame on GPU 0 alone
    if True:
        want_b = None
        tr.shard_config(1, 0)
        if rank == 0:
            want_b = tr.trace_poses(poses, w, h, api.PWN_F_BLUR, want_sums=True)
            tr.trace_cam(poses[0]["cam"], w, h, float(poses[0]["sec_current"]), api.PWN_F_BLUR)
            out["blur_on_gpu0_ms"] = tr.last_kernel_ms()[1]
        tr.shard_config(world, rank)
        # first the interleaved layout (every K1 stores its bands into the pre-blur buffer of every shard), then the
        # block layout (halo bands only, the blur under the trace): the second is what bench.py's headline times
        for layout in ("interleaved", "blocks"):
            tr.shard_layout(blocks=(layout == "blocks"))
            if layout == "blocks":
                out["block_boundaries_bands"] = [int(b) for b in balance_blocks(tr, rank, world, dist, poses[::4], w, h)]
            dist.barrier()
            bases = open_all_peers(tr, rank, world, dist)
            got = []
            for it in range(2):
                t1 = t2 = 0.0
                for k, p in enumerate(poses):
                    dist.barrier()
                    tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]))
                    a, b = tr.last_kernel_ms()
                    t1 += a
                    t2 += b
                    if it == 1 and k % 4 == 0 and rank == 0:
                        s0, _, _ = tr.read_frame(want_z=False)
                        got.append(int(api.checksum_ref(s0, w, h)))
            key = "sharded_blur" if layout == "interleaved" else "sharded_blur_blocks"
            out[key + "_trace_ms"] = allmax(t1) / frames
            out[key + "_rest_ms"] = allmax(t2) / frames
            out[key + "_ms"] = allmax(t1 + t2) / frames
            out[key + "_mrays"] = w * h / out[key + "_ms"] / 1e3
            if rank == 0:
                out[key + "_frames_equal_one_gpu"] = bool(got == [int(v) for v in want_b[::4]])
            if layout == "interleaved":
                dist.barrier()
                close_all_peers(tr, bases)
        timed_out = tr.peers_timed_out()
        # end to end into HOST memory: the blurred bands stay on their GPU and every rank copies its own
        # into one frame in shared memory that all ranks have page-locked -- all the PCIe links at once,
        # nothing gathered on GPU 0.  Wall clock between barriers, max over ranks.
        host_ok = None
        try:
            host, shm = _shared_host_frame(rank, dist, h, w)
            api.host_register(host)
            try:
                tr.trace_shard_blur(poses[0]["cam"], w, h, float(poses[0]["sec_current"]), api.PWN_F_LOCAL_BANDS)
                tr.read_bands(w, h, host)
                tot = 0.0
                good = True
                for k, p in enumerate(poses):
                    dist.barrier()
                    t0 = time.perf_counter()
                    tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]), api.PWN_F_LOCAL_BANDS)
                    tr.read_bands(w, h, host)
                    dist.barrier()
                    tot += time.perf_counter() - t0
                    if rank == 0 and k % 4 == 0:
                        good = good and int(api.checksum_ref(host, w, h)) == int(want_b[k])
                    dist.barrier()
                out["host_frame_ms"] = allmax(tot) / frames * 1e3
                out["host_frame_mrays"] = w * h / out["host_frame_ms"] / 1e3
                host_ok = good
            finally:
                tr.sync()
                api.host_unregister(host)
                del host
                shm.close()
                dist.barrier()
                if rank == 0:
                    shm.unlink()
        except OSError as e:      # no POSIX shared memory on this box (an environment limit, reported; nothing else is caught)
            out["host_frame_error"] = "%s: %s" % (type(e).__name__, e)
        dist.barrier()
        close_all_peers(tr, bases)
        tr.shard_blocks(None)
        tr.shard_layout(blocks=False)
        if rank == 0:
            out["sharded_blur_timed_out"] = bool(timed_out)
            if host_ok is not None:
                out["host_frame_equal_one_gpu"] = bool(host_ok)

    # gather: local trace, then one NCCL gather of the packed rows (sbuf)
    s, _ = frame_tensors(tr)
    gather_frame(s, h, rank, world, dist)      # untimed: the first collective builds the NCCL communicator
    torch.cuda.synchronize()
    tot_k = tot_g = 0.0
    for p in poses:
        dist.barrier()
        tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
        tot_k += tr.last_kernel_ms()[0]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gather_frame(s, h, rank, world, dist)
        b.record()
        b.synchronize()
        tot_g += a.elapsed_time(b)
    gk, gg = allmax(tot_k) / frames, allmax(tot_g) / frames
    tr.shard_config(1, 0)

    # the 64-pose batch of configs[4]: sharded by pose (pose p -> rank p % N), no exchange at all
    batch = levels.poses_random(lv, 64, seed=9)
    mine = batch[rank::world].copy()
    tr.trace_poses(mine[:2], w, h, api.PWN_F_BLUR)
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    tr.trace_poses(mine, w, h, api.PWN_F_BLUR)
    b.record(stream)
    b.synchronize()
    pb = allmax(a.elapsed_time(b))
    out["pose_batch_64_ms"] = pb
    out["pose_batch_64_mrays"] = 64 * w * h / pb / 1e3
    px = w * h / 1e3
    out.update({"one_gpu_ms": one, "one_gpu_mrays": px / one,
                "fused_ms": fused, "fused_mrays": (px / fused) if fused else None,
                "fused_sbuf_only_ms": fused_s, "fused_sbuf_only_mrays": (px / fused_s) if fused_s else None,
                "gather_trace_ms": gk, "gather_nccl_ms": gg, "gather_mrays": px / (gk + gg),
                "sbuf_bytes": w * h * 4})
    return out
