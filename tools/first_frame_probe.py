"""Scene loads of the 1 M-triangle soup in a process whose CUDA context is warm: host builder against GPU builder
(set YART_B200_BUILD_TRACE=1 for the device build's stage times)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yart_b200 as Y
import bench

path = bench.scene_path(1_000_000, "soup")
sc = Y.Scene(path, bvh_kind=Y.BVH_SAH_HOST)
ctx = Y.Context(max_depth=1)
ctx.upload_scene(sc)
for rep in range(2):
    t0 = time.time()
    c2 = Y.Context(max_depth=1)
    t1 = time.time()
    c2.upload_scene(sc)
    t2 = time.time()
    c2.set_camera(Y.make_camera(1920, 1080, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"]))
    c2.begin_frame(1920, 1080, 4, 64, (0, 0, 0), Y.TONEMAP_AGX)
    t3 = time.time()
    c2.render_wave(0, 4, 0)
    t4 = time.time()
    c2.close()
    print(f"context {1e3 * (t1 - t0):.0f} ms, upload + wide collapse {1e3 * (t2 - t1):.0f} ms, begin_frame {1e3 * (t3 - t2):.0f} ms, first wave {1e3 * (t4 - t3):.0f} ms", flush=True)
for kind, name in ((Y.BVH_SAH_HOST, "host"), (Y.BVH_SAH, "gpu"), (Y.BVH_SAH, "gpu"), (Y.BVH_SAH_HOST, "host"), (Y.BVH_SAH, "gpu")):
    t0 = time.time()
    s = Y.Scene(path, bvh_kind=kind)
    print(f"{name}: load {1e3 * (time.time() - t0):.0f} ms, build_ms {s.build_ms:.0f}, on gpu {s.device_builds}", flush=True)
    s.close()
