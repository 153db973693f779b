"""How long cudaMalloc / cudaFree of the GPU BVH builder's 230 MB take in a process with and without a live render context."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yart_b200 as Y
import bench

rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]


def probe(tag, mb=230):
    out = []
    for _ in range(4):
        p = C.c_void_p()
        t0 = time.time()
        rc = rt.cudaMalloc(C.byref(p), mb << 20)
        t1 = time.time()
        rt.cudaDeviceSynchronize()
        rt.cudaFree(p)
        t2 = time.time()
        out.append(f"malloc {1e3 * (t1 - t0):.1f} free {1e3 * (t2 - t1):.1f} (rc {rc})")
    print(tag, " | ".join(out), flush=True)


rt.cudaSetDevice(0)
rt.cudaFree(None)
probe("fresh process:")
sc = Y.Scene(bench.scene_path(100_000, "soup"))
ctx = Y.Context(max_depth=1)
ctx.upload_scene(sc)
ctx.set_camera(Y.make_camera(1920, 1080, 35.0, 0.0, (0, 0, 40), (0, 0, 0)))
ctx.begin_frame(1920, 1080, 4, 64, (0, 0, 0), Y.TONEMAP_AGX)
ctx.render_wave(0, 4, 0)
probe("with a live render context:")
ctx.close()
probe("after closing it:")
