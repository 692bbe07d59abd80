import sys, time, numpy as np
sys.path.insert(0, '.')
from pwnfps_b200 import api, levels
lv = levels.mirror_hall()
n = 2
order = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "01")]
ts = [api.Tracer(0) for _ in range(n)]
w, h = 640, 360
for t in ts:
    t.level_upload(lv); t.level_prepare_render(lv); t.frame_config(w, h); t.frame_clear()
bases = [t.device_ptrs()[0] for t in ts]
for r, t in enumerate(ts):
    t.shard_config(n, r); t.shard_layout(blocks=(len(sys.argv) < 3), halo_bands=0); t.frame_peers(bases)
cam = api.mat4_roty(api.cam_spawn(lv), 1.57)
for t in ts:
    t.debug_flags(96)       # creates the debug stream before anything spins
t0 = time.time()
for r in order:
    a = time.time()
    ts[r].trace_shard_blur(cam, w, h, 0.3)
    print("submit shard %d took %.3f s" % (r, time.time() - a))
for k in range(6):
    for r, t in enumerate(ts):
        f = t.debug_flags(96)
        print("%.2fs shard" % (time.time() - t0), r, "A", f[0:2], "B", f[8:10], "err", f[16], "defer", f[32], "bands", f[64:64+12])
    time.sleep(0.2)
for t in ts:
    try:
        t.sync()
    except Exception as e:
        print("sync:", str(e)[:60])
for r, t in enumerate(ts):
    f = t.debug_flags(96)
    print("after: shard", r, "A", f[0:2], "B", f[8:10], "err", f[16], "defer", f[32], "bands", f[64:64+12])
