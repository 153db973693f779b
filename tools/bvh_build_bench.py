"""Times yc_build_bvh_sah (the SAH BVH on the GPU) on the 1 M-triangle soup: cold call, then warm calls."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import yart_b200 as Y
from yart_b200 import capi, scenes
lib = capi.load()
s = scenes.soup(1_000_000)
m = s.meshes[0]
pos = np.ascontiguousarray(m.positions, np.float32); faces = np.ascontiguousarray(m.faces, np.uint32)
n = len(faces)
pool = np.zeros((2*n+2, 10), np.uint32); idx = np.zeros(n, np.uint32)
nn = C.c_uint32(); lv = C.c_uint32()
f = lib.yc_build_bvh_sah
for rep in range(4):
    t0 = time.time()
    rc = f(0, pos.ctypes.data, len(pos), faces.ctypes.data, n, pool.ctypes.data, C.byref(nn), idx.ctypes.data, C.byref(lv))
    print(f"rep {rep}: rc {rc} {1e3*(time.time()-t0):.1f} ms nodes {nn.value} levels {lv.value}", flush=True)
