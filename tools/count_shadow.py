import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import yart_b200 as Y
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "soup"
tris = bench.DEFAULT_TRIS[w]; bench.select_workload(w, tris)
sc = Y.Scene(bench.scene_path(tris, w))
cam = Y.make_camera(bench.W, bench.H, bench.CAM["focal"], bench.CAM["fnum"], bench.CAM["pos"], bench.CAM["target"], (0, 0, 0), bench.CAM["exposure"])
ctx = Y.Context(max_depth=bench.MAX_DEPTH, tail_threshold=-1)
ctx.upload_scene(sc); ctx.set_camera(cam)
Y.lib().yc_set_profiling(ctx._h, 2)
ctx.begin_frame(bench.W, bench.H, 4, 64, (0, 0, 0), Y.TONEMAP_AGX)
ctx.render_wave(0, 4, 0)
st = ctx.stats()
print(w, "extend rays", st.raysExtend, "shadow rays", st.raysShadow, "box", st.boxTests, "tri", st.triTests, "ref rays", st.raysReference)
print("box per traced ray", st.boxTests / (st.raysExtend + st.raysShadow), "tri per ray", st.triTests / (st.raysExtend + st.raysShadow))
