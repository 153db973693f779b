#!/usr/bin/env python
"""bench.py — Mrays/s of the render hot path on BASELINE.json configs[1]:
synthetic 1M-triangle random-soup scene, 1920x1080, primary + shadow rays only (maxDepth = 1).

A step = one pass of the hot path over one batch: one wave of SPP samples of every pixel of the
1080p frame (raygen → extend → shade → shadow → accumulate → finalize).

  value      whole-job Mrays/s (reference ray definition, mis-integrator.cpp:22,126) with the scene,
             BVH and camera resident in HBM; timed with CUDA events on the launching stream
  e2e        same metric through the public Renderer API (yr_render_sync + yr_read): per step the
             frame set-up H2D copies and the HDR+LDR frame D2H copies are inside the timed region
  roofline   extend (closest-hit) kernel: algorithmic bytes per launch / its CUDA-event duration
  cpu_baseline  the reference's own tile-threaded CPU renderer (oracle/_ref/oracle_ref_perf) on the
             box's host cores, bounded to 1 spp of the same frame

N > 1 (torchrun, one rank per GPU): scene replicated, rank r renders sample wave r of the frame
(weak scaling: SPP samples per GPU), the per-GPU HDR frames are combined with an NCCL all-reduce
over NVLink and re-tonemapped; value = rays of all ranks / max-over-ranks time.

--impl reference times the reference CPU renderer alone (all host threads), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
_REAL_STDOUT = sys.stdout
METRIC = "Mrays/s at 1080p (primary + shadow rays, reference ray count)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# workload → (scene generator, kwargs, maxDepth, description).  "soup" is BASELINE.json configs[1] (the bench
# line of record); "sponza" / "mclaren" are configs[2] / configs[3] shapes, full MIS+NEE paths (maxDepth 30).
WORKLOADS = {
    "soup": ("soup", {}, 1, "soup {tris} tris, {W}x{H}, primary+shadow (maxDepth 1)"),
    "sponza": ("sponza", dict(tex_res=512, env_res=1024), 30, "Sponza-shaped {tris} tris, textured PBR + normal maps, HDR env only, {W}x{H}, full paths"),
    "mclaren": ("mclaren", dict(env_res=1024), 30, "McLaren-shaped {tris} tris, clearcoat/chrome/glass+volume, lamps + HDR env, {W}x{H}, full paths"),
}
DEFAULT_TRIS = {"soup": 1_000_000, "sponza": 260_000, "mclaren": 2_000_000}
CAM = dict(pos=(0.0, 0.0, 40.0), target=(0.0, 0.0, 0.0), focal=35.0, fnum=0.0, exposure=0.0)
MAX_DEPTH = 1
WORKLOAD_TEXT = ""


def select_workload(name: str, n_tris: int):
    """Sets the module-level camera / depth / description for the chosen workload."""
    global CAM, MAX_DEPTH, WORKLOAD_TEXT
    from yart_b200 import scenes
    gen, kw, depth, text = WORKLOADS[name]
    small = dict(kw, n_tris=100)
    if "tex_res" in small:
        small["tex_res"] = 4
    if "env_res" in small:
        small["env_res"] = 4
    cam = getattr(scenes, gen)(**small).camera
    CAM = dict(pos=tuple(cam["pos"]), target=tuple(cam["target"]), focal=cam["focal"], fnum=cam["fnum"], exposure=cam["exposure"])
    MAX_DEPTH = depth
    WORKLOAD_TEXT = text.format(tris=n_tris, W=W, H=H)


def scene_path(n_tris: int, workload: str = "soup") -> str:
    import tempfile
    from yart_b200 import scenes
    gen, kw, _, _ = WORKLOADS[workload]
    d = os.environ.get("YART_BENCH_CACHE", os.path.join(tempfile.gettempdir(), "yart_b200_bench"))
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, f"{gen}_{n_tris}.ysc")
    if not os.path.exists(p):
        t0 = time.time()
        tmp = p + f".tmp{os.getpid()}"
        getattr(scenes, gen)(n_tris=n_tris, **kw).write(tmp)
        os.replace(tmp, p)
        log(f"[bench] generated {p} in {time.time() - t0:.1f}s")
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        if index < 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for t, l in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l.split(", ") for _, l in self.lines]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference(path: str, steps: int, warmup: int, spp: int = 1) -> dict:
    """The reference's TileRenderer on all host threads: `steps` timed renderSync() calls of
    `spp` samples of the 1080p frame, after `warmup` untimed ones (one process, one BVH build)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "oracle_ref_perf")
    if not os.path.exists(exe):
        exe = os.path.join(ROOT, "oracle", "_ref", "oracle_ref")
    out = "/tmp/yart_bench_ref.bin"
    cmd = [exe, "render", path, out, f"w={W}", f"h={H}", f"spp={spp}", f"maxdepth={MAX_DEPTH}", "tonemap=agx",
           "pos=%g,%g,%g" % CAM["pos"], "target=%g,%g,%g" % CAM["target"], f"focal={CAM['focal']}", f"fnum={CAM['fnum']}",
           f"exposure={CAM['exposure']}", f"repeat={steps + warmup}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference CPU renderer failed: {r.stderr[-2000:]}")
    j = json.loads(r.stdout.strip().splitlines()[-1])
    rays, prev = [], 0
    for tot, ms in j["steps"]:  # m_totalRays accumulates over renderSync calls
        rays.append((tot - prev, ms))
        prev = tot
    timed = rays[warmup:]
    tot_rays, tot_ms = sum(r for r, _ in timed), sum(m for _, m in timed)
    return dict(mrays=tot_rays / tot_ms / 1e3, ms_per_step=tot_ms / len(timed), rays_per_step=tot_rays / len(timed),
                threads=j["threads"], build_ms=j["build_ms"], spp=spp, exe=os.path.basename(exe))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    path = scene_path(args.tris, args.workload)
    res = cpu_reference(path, args.steps, max(args.warmup, 1), spp=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT, "step": "1 spp of the frame",
                   "bvh_build_ms_excluded": res["build_ms"]},
        "cpu_baseline": {"value": res["mrays"], "unit": "Mrays/s", "cores": res["threads"], "kind": "reference",
                         "sample": f"{args.steps} x 1 spp of the 1080p frame, {res['exe']} (unmodified reference, -O3 -march=x86-64-v3)"},
        "e2e": {"value": res["mrays"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_s": W * H * 1 / (res["ms_per_step"] / 1e3),
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def run_ours(args):
    import numpy as np
    import yart_b200 as Y
    from yart_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        path = scene_path(args.tris, args.workload)
    if dist:
        dist.barrier()
    path = scene_path(args.tris, args.workload)

    Y.use_library(capi.load())  # raises if libyart_b200.so is missing: no fallback
    t0 = time.time()
    scene = Y.Scene(path)
    log(f"[bench r{rank}] scene {scene.n_tris} tris, host SAH build {scene.build_ms:.0f} ms (load {time.time() - t0:.1f}s)")
    cam = Y.make_camera(W, H, CAM["focal"], CAM["fnum"], CAM["pos"], CAM["target"], (0, 0, 0), CAM["exposure"])
    spp = args.spp

    # ---- device-resident arm ---------------------------------------------------------------
    trav = {"auto": Y.TRAVERSAL_AUTO, "reference": Y.TRAVERSAL_REFERENCE_ORDER, "wide": Y.TRAVERSAL_WIDE}[args.traversal]
    ctx = Y.Context(device=local, max_depth=MAX_DEPTH, traversal=trav)
    ctx.upload_scene(scene)
    ctx.set_camera(cam)
    ctx.set_profiling(os.environ.get('YART_BENCH_NO_PROFILE') is None)
    n_warm = max(args.warmup, 3)
    waves_per_rank = n_warm + args.steps
    total_spp = spp * world * waves_per_rank  # the job: every rank renders `waves_per_rank` waves of `spp` samples
    ctx.begin_frame(W, H, total_spp, 64, (0, 0, 0), Y.TONEMAP_AGX)

    frame_t = None
    buckets = world > 1 and args.sharding == "buckets"
    bucket_t, ar_events = None, []
    if buckets:
        import torch
        b_ptr, b_bytes, _, plane_pix = ctx.bucket_device_ptrs()
        m_wave = ctx.wave_buckets(spp * world)

        class _BAlias:  # zero-copy int32 view of the planes a wave of spp * world samples uses
            __cuda_array_interface__ = {"shape": (m_wave * plane_pix * 4,), "typestr": "<i4", "data": (b_ptr, False), "version": 3}
        bucket_t = torch.as_tensor(_BAlias(), device=f"cuda:{local}")
    if dist:
        import torch
        hdr_ptr, _, nbytes = ctx.frame_device_ptrs()

        class _Alias:  # zero-copy view of the context's HDR frame for NCCL
            __cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (hdr_ptr, False), "version": 3}
        frame_t = torch.as_tensor(_Alias(), device=f"cuda:{local}")

    def sync_all():
        if dist:
            import torch
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def step(k):
        """One pass: wave (k * world + rank) of the job — `spp` samples of every pixel — blended into this
        rank's HDR frame with finishTile's sample-count weights."""
        k = min(k, waves_per_rank - 1)
        if buckets:
            # sample sharding inside a wave of spp * world samples by (estimator bucket, pixel class) units: every
            # rank accumulates its units, the GMoN accumulation buffers are summed with NCCL (int32: disjoint
            # slots, bitwise exact), every rank finalizes — bit-identical to one GPU rendering the wave
            import torch
            S = spp * world
            ctx.accumulate_wave(k * S, S, bucket_shard=rank, bucket_shard_count=world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(bucket_t)
            e1.record()
            torch.cuda.synchronize()
            ar_events.append((e0, e1))
            ctx.finalize_wave(S, k * S)
            return
        ctx.render_wave((k * world + rank) * spp, spp, k * spp)

    def combine(scratch=False):
        """N > 1: the per-GPU HDR frames (equal sample counts) are averaged with one NCCL all-reduce over
        NVLink and re-tonemapped — once per job, inside the timed region.  scratch=True (warm-up) runs the
        same collective on a copy so the frames being accumulated are left alone."""
        if not dist or buckets:
            return 0.0
        import torch
        t = frame_t.clone() if scratch else frame_t
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t.mul_(1.0 / world)
        dist.all_reduce(t)
        e1.record()
        torch.cuda.synchronize()
        if not scratch:
            ctx.retonemap()
        return e0.elapsed_time(e1)

    # started before the warm-up: nvidia-smi needs ~0.3 s to print its first sample.  Rank 0 samples its own
    # GPU only: eight concurrent nvidia-smi pollers contend for the driver lock and perturb the step time.
    clocks = ClockSampler(local if rank == 0 else -1)
    for k in range(n_warm):
        step(k)
    if dist:
        combine(scratch=True)  # warms NCCL up (parity of this path: tests/test_multi_gpu_cpu.py)
    sync_all()
    s0 = ctx.stats()
    ar_events.clear()
    w0 = time.time()
    for k in range(args.steps):
        step(n_warm + k)
    ar_total = combine() + sum(a.elapsed_time(b) for a, b in ar_events)
    sync_all()
    w1 = time.time()
    s1 = ctx.stats()
    # keep the GPU under the same load until a few clock samples exist (not timed); the number of
    # extra steps is agreed across ranks so nothing can hang
    n_extra = int((0.5 - (w1 - w0)) / max((w1 - w0) / args.steps, 1e-4)) + 1 if w1 - w0 < 0.5 else 0
    if dist:
        import torch
        t = torch.tensor([n_extra], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_extra = int(t.item())
    for _ in range(min(n_extra, 200)):
        step(waves_per_rank - 1)
    sync_all()
    clk = clocks.stop(w0, time.time())
    dev_ms = (s1.gpuMs - s0.gpuMs) + ar_total
    rays = s1.raysReference - s0.raysReference
    traced = (s1.raysExtend - s0.raysExtend) + (s1.raysShadow - s0.raysShadow)
    launches = s1.kernelLaunches - s0.kernelLaunches

    # ---- extend-kernel roofline: algorithmic bytes of the reference traversal on these rays ------
    rspp = min(spp, max(1, (1 << 23) // (W * H)))  # roofline launch: at most 8 Mi primary rays (4 spp at 1080p)
    n_rays = W * H * rspp
    rays_dev, hits_dev = ctx.device_alloc(n_rays * 32), ctx.device_alloc(n_rays * 20)
    ctx.begin_frame(W, H, total_spp, 64, (0, 0, 0), Y.TONEMAP_AGX)
    ctx.generate_primary_rays((n_warm * world + rank) * spp, rspp, rays_dev)  # the first timed wave's rays
    one = np.array([[0, 0, 40, 0.001, 0.01, 0.02, -0.99975, np.inf]], np.float32)  # any ordinary ray
    ctx.trace(one, Y.TRACE_CLOSEST | Y.TRACE_COUNT)  # zeroes the work counters
    c0 = ctx.stats()
    lib, h = Y.lib(), ctx._h
    import ctypes as C
    ms_c = C.c_float()
    rc = lib.yc_trace_device(h, rays_dev, n_rays, Y.TRACE_CLOSEST | Y.TRACE_COUNT, hits_dev, 1, C.byref(ms_c))
    assert rc == 0
    c1 = ctx.stats()
    box, tri = c1.boxTests - c0.boxTests, c1.triTests - c0.triTests
    trace_ms = ctx.trace_device(rays_dev, n_rays, hits_dev, Y.TRACE_CLOSEST, repeat=3)
    ctx.device_free(rays_dev)
    ctx.device_free(hits_dev)
    algo_bytes = n_rays * (32 + 20) + 32 * box + 52 * tri
    peak, peak_src = peaks()
    # The dominant kernel is the persistent closest-hit traversal (extendKernel / traceKernel: the same
    # tracePersistent body).  Roofline launch = ONE launch over the step's W*H*spp primary rays, timed alone with
    # CUDA events right here (trace_ms, mean of 3 after the step warm-up), against its algorithmic bytes.
    # Inside a step the same rays are traced by `in_step_launches` launches (one per chunk; two chunks of a wave
    # are in flight on two streams), whose event times overlap each other and the other lane's kernels: their
    # aggregate is reported next to it (sum of bytes / (sum of launch times / lanes in flight)).
    n_ext = max(1, s1.extendLaunches - s0.extendLaunches)
    ext_rays = (s1.raysExtend - s0.raysExtend)
    ext_ms_sum = s1.extendMs - s0.extendMs
    lanes = 2
    in_step = None
    if MAX_DEPTH == 1 and ext_ms_sum > 0:
        # maxDepth 1: every extend ray of the timed steps is a primary ray with the bytes counted above per ray
        step_bytes = algo_bytes / n_rays * ext_rays
        in_step = {"launches": int(n_ext), "rays_per_launch": ext_rays / n_ext, "avg_launch_ms": ext_ms_sum / n_ext,
                   "lanes_in_flight": lanes, "aggregate_gbs": step_bytes / (ext_ms_sum / lanes / 1e3) / 1e9}
    roof_ms, roof_kernel = trace_ms, "tracePersistent<closest> (extendKernel / traceKernel body), one launch over the step's primary rays, timed alone"
    achieved = algo_bytes / (roof_ms / 1e3) / 1e9 if roof_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "extend_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        # DRAM bytes of one ncu-captured launch, scaled to this launch's ray count
        traffic = tj.get("dram_bytes_per_launch") * n_rays / tj.get("rays_per_launch", n_rays)

    # ---- end-to-end arm: public Renderer API, host buffers ------------------------------------
    r = Y.Renderer(W, H, cam, scene, samples=spp, first_wave_samples=spp, max_wave_samples=spp, max_depth=MAX_DEPTH,
                   tonemap=Y.TONEMAP_AGX, device=local, traversal=trav)
    e2e_rays = 0

    def e2e_step():
        nonlocal e2e_rays
        d = r.render_sync()
        hdr, ldr, _ = r.read(pinned=True)
        e2e_rays = d["total_rays"]
        return hdr

    for _ in range(2):
        e2e_step()
    sync_all()
    e0 = time.time()
    n_e2e = max(2, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_step()
    sync_all()
    e2e_s = (time.time() - e0)
    r.close()

    # ---- max over ranks, aggregate ------------------------------------------------------------
    if dist:
        import torch
        t = torch.tensor([dev_ms, (w1 - w0) * 1e3, e2e_s], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, e2e_s = t.tolist()
        c = torch.tensor([rays, traced, e2e_rays * n_e2e, launches], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(c)
        rays, traced, e2e_total, launches = c.tolist()
    else:
        wall_ms, e2e_total = (w1 - w0) * 1e3, e2e_rays * n_e2e

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                cr = cpu_reference(path, 2, 1, spp=1)
                cpu = {"value": cr["mrays"], "unit": "Mrays/s", "cores": cr["threads"], "kind": "reference",
                       "sample": f"2 x 1 spp of the same 1080p frame through the unmodified reference's TileRenderer "
                                 f"({cr['exe']}); BVH build {cr['build_ms']:.0f} ms excluded"}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        line = {
            "metric": METRIC, "value": rays / dev_ms / 1e3, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT,
                       "step": f"one wave of {spp} spp per GPU ({W * H * spp} camera paths)",
                       "parallelism": (f"bucket sharding x{world}: every wave of {spp * world} samples is split into (GMoN bucket b, pixel "
                                       "class c) units, rank r takes (b + c) % N == r; scene replicated; NCCL all-reduce(int32 sum) "
                                       "of the accumulation buffers per wave, inside the timed region") if buckets else
                                      (f"sample-wave sharding x{world}: rank r renders waves r, r+N, ... of the job; scene replicated; "
                                       "one NCCL all-reduce of the HDR frames per job (inside the timed region)"),
                       "l2": "inputs larger than L2: BVH + 0.8 GB of path state streamed per step exceed the 126 MB L2; no flush"},
            "wall_ms_per_step": wall_ms / args.steps,
            "traced_mrays_per_s": traced / dev_ms / 1e3,
            "samples_per_s": W * H * spp * world / (dev_ms / args.steps / 1e3),
            "gpu_launches": int(launches),
            "clocks": clk,
            "e2e": {"value": e2e_total / e2e_s / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 256, "d2h_bytes_per_step": 2 * W * H * 16,
                    "call": "Renderer.render_sync() + Renderer.read(pinned=True) (yr_render_sync + yr_read): camera/frame description in, HDR + LDR frames out to page-locked host memory"},
            "metric_note": "value counts rays as the reference does (path segments + unoccluded NEE rays); traced_mrays_per_s counts every ray traced",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if args.workload == "soup" else None, "kernel": roof_kernel, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "rays_per_launch": n_rays, "avg_launch_ms": roof_ms,
                         "box_tests_per_ray": box / n_rays, "tri_tests_per_ray": tri / n_rays,
                         "in_step_extend_launches": in_step,
                         "note": "frac > 1 is possible: the algorithmic bytes (SURVEY 8d: the reference BVH2's node and triangle "
                                 "bytes per test) are served by the L2, where the BVH is resident; DRAM moves only `traffic`. ncu: "
                                 "issue slots 65 %, L1/TEX 77 %, 18.7 of 32 lanes active, top stall L2 latency (profiles/README.md)"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # Only the JSON line may reach stdout: libraries (NCCL prints its version there) write to fd 1 too,
    # so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="soup", choices=sorted(WORKLOADS))
    ap.add_argument("--tris", type=int, default=0, help="triangle count (default: the workload's BASELINE.json size)")
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--sharding", default="waves", choices=["waves", "buckets"],
                    help="N > 1: whole waves per rank (default) or the samples of one wave split by GMoN bucket")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--traversal", default="auto", choices=["auto", "reference", "wide"],
                    help="YcOptions::traversal: auto = the 4-wide BVH for scenes without alpha-tested materials (default), "
                         "reference = the reference-order BVH2 walk (bit-identical frames)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080, help="3840 x 2160 for the BASELINE.json configs[4] shape")
    args = ap.parse_args()
    global W, H, METRIC
    W, H = args.width, args.height
    if (W, H) != (1920, 1080):
        METRIC = METRIC.replace("1080p", f"{W}x{H}")
    if not args.tris:
        args.tris = DEFAULT_TRIS[args.workload]
    select_workload(args.workload, args.tris)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
