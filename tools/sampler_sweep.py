"""How the sampler configuration (log2 of the job's total spp) changes primary-ray coherence: traversal time
of the same 4-spp wave of 1080p primary rays under different totalSamples."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yart_b200 as Y
import bench
sc = Y.Scene(bench.scene_path(1_000_000))
ctx = Y.Context(max_depth=1)
ctx.upload_scene(sc)
W, H = 1920, 1080
ctx.set_camera(Y.make_camera(W, H, 35.0, 0.0, (0, 0, 40), (0, 0, 0)))
n = W * H * 4
rays, hits = ctx.device_alloc(n * 32), ctx.device_alloc(n * 20)
for total in (4, 16, 32, 52, 64, 128, 256, 512, 1024, 2048, 4096):
    for off in (0, 12):
        if off + 4 > total:
            continue
        ctx.begin_frame(W, H, total, 64, (0, 0, 0), Y.TONEMAP_NONE)
        ctx.generate_primary_rays(off, 4, rays)
        ms = ctx.trace_device(rays, n, hits, Y.TRACE_CLOSEST, repeat=3)
        print(f"totalSamples {total:5d} sampleOffset {off:3d}: {ms:.3f} ms")
