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
    Returns the stitched (H, W) tensor on rank 0, None elsewhere.  One gather of padded row packs."""
    import torch
    rows = torch.as_tensor(owned_rows(dimy, world, rank), device=local_frame.device)
    nmax = max_owned_rows(dimy, world)
    pack = torch.zeros((nmax, local_frame.shape[1]), dtype=local_frame.dtype, device=local_frame.device)
    pack[:len(rows)] = local_frame[rows]
    if rank == 0:
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist.gather(pack, parts, dst=0)
        out = torch.zeros_like(local_frame)
        for r in range(world):
            rr = torch.as_tensor(owned_rows(dimy, world, r), device=local_frame.device)
            out[rr] = parts[r][:len(rr)]
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


def open_peer_frame(tr, rank, dist):
    """Fused path set-up: every rank > 0 redirects its K1 output into rank 0's frame."""
    objs = [None]
    if rank == 0:
        objs = [tr.ipc_export()]
    dist.broadcast_object_list(objs, src=0)
    if rank == 0:
        return None
    handle, so, zo, to = objs[0]
    base = tr.ipc_import(handle)
    tr.frame_redirect(base + so, base + zo, base + to)
    return base


def close_peer_frame(tr, base):
    tr.frame_redirect(0, 0, 0)
    if base is not None:
        tr.ipc_close(base)


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

    # fused: K1 stores into GPU 0's frame over NVLink
    tr.shard_config(world, rank)
    fused = None
    try:
        base = open_peer_frame(tr, rank, dist)
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
            fused = allmax(tot) / frames
        dist.barrier()
        close_peer_frame(tr, base)
        if rank == 0:
            out["fused_frames_equal_one_gpu"] = bool(got == [int(v) for v in want[::4]])
    except api.PwnError as e:
        out["fused_error"] = str(e)
        tr.frame_redirect(0, 0, 0)

    # gather: local trace, then one NCCL gather of the packed rows (sbuf)
    s, _ = frame_tensors(tr)
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
                "gather_trace_ms": gk, "gather_nccl_ms": gg, "gather_mrays": px / (gk + gg),
                "sbuf_bytes": w * h * 4})
    return out
