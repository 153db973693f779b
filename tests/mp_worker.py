"""Worker for the world_size-2 CPU tests of the multi-GPU host logic (gloo backend).
Each rank plays one GPU: it renders its share with the hostsim build and the frames are combined
with torch.distributed exactly as bench.py / a multi-GPU caller would with NCCL."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(rank: int, world: int, port: int, mode: str, out_dir: str):
    import torch
    import torch.distributed as dist
    import harness as H
    import yart_b200 as Y
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if mode.startswith("nccl_"):
        return run_nccl(rank, world, mode, out_dir)
    Y.use_library(H.hostsim())
    cam = H.scene_camera("cornell")
    sc = Y.Scene(H.scene_file("cornell"))
    w = h = 48
    c = Y.make_camera(w, h, cam["focal"], cam["fnum"], cam["pos"], cam["target"])
    ctx = Y.Context(traversal=Y.TRAVERSAL_REFERENCE_ORDER)  # the parent compares bitwise with a reference-order render
    ctx.upload_scene(sc)
    ctx.set_camera(c)
    spp = 8
    if mode in ("yr_tiles", "yr_buckets"):
        # the whole data plane behind yr_* (yr_create_dist_custom): the library drives the waves and calls back into
        # this process only for its sum collective, which gloo provides here (NCCL on the GPU box)
        import ctypes as C
        dt = {0: (np.float32, torch.float32), 1: (np.int32, torch.int32), 2: (np.int64, torch.int64)}

        def collective(buf, count, dtype, root):
            npd, _ = dt[dtype]
            arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(np.ctypeslib.as_ctypes_type(npd))), shape=(count,))
            t = torch.from_numpy(arr)
            if root < 0:
                dist.all_reduce(t)
            else:
                dist.reduce(t, dst=root)
            return 0

        waves = dict(samples=32, first_wave_samples=16, max_wave_samples=16)  # two waves of 16: m = 3 buckets
        r = Y.Renderer(w, h, c, sc, tile_size=16, tonemap=Y.TONEMAP_AGX, traversal=Y.TRAVERSAL_REFERENCE_ORDER,
                       sharding=Y.SHARD_BUCKETS if mode == "yr_buckets" else Y.SHARD_TILES, dist=(rank, world, collective), **waves)
        seen = []
        r.on_wave_complete(lambda rd, wd: seen.append((wd["wave"], wd["wave_samples"], wd["rays"], rd["total_rays"])))
        data = r.render_sync()
        hdr, ldr, _ = r.read()
        if rank == 0:
            np.save(os.path.join(out_dir, f"{mode}_hdr.npy"), hdr)
            np.save(os.path.join(out_dir, f"{mode}_ldr.npy"), ldr)
            np.save(os.path.join(out_dir, f"{mode}_rays.npy"), np.array([data["total_rays"], len(seen), sum(s[2] for s in seen)], np.int64))
        elif mode == "yr_buckets":
            np.save(os.path.join(out_dir, f"{mode}_hdr_rank{rank}.npy"), hdr)  # bucket sharding: every rank ends with the frame
        r.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    if mode == "tiles":
        # interleaved tile sharding (SURVEY §8e primary): disjoint pixels, sum = full frame, bit-exact
        ctx.begin_frame(w, h, spp, 16, (0, 0, 0), Y.TONEMAP_AGX, shard_index=rank, shard_count=world)
        ctx.render_wave(0, spp, 0)
        hdr, ldr, st = ctx.resolve()
        t_hdr, t_ldr = torch.from_numpy(hdr), torch.from_numpy(ldr)
        dist.all_reduce(t_hdr)
        dist.all_reduce(t_ldr)
    elif mode == "buckets":
        # sample sharding inside a wave by (estimator bucket, pixel class) units (SURVEY §8e alternative B): every
        # rank accumulates its units, the GMoN accumulation buffers are summed across ranks as 32-bit
        # integers (disjoint slots: bitwise exact), every rank finalizes the wave
        waves = [16, 16]
        ctx.begin_frame(w, h, sum(waves), 16, (0, 0, 0), Y.TONEMAP_AGX)
        ptr, nbytes, _, _ = ctx.bucket_device_ptrs()
        taken = 0
        for wv in waves:
            ctx.accumulate_wave(taken, wv, bucket_shard=rank, bucket_shard_count=world)
            acc = np.empty(nbytes // 4, np.int32)
            ctx.d2h(acc, ptr)
            t = torch.from_numpy(acc)
            dist.all_reduce(t)
            ctx.h2d(ptr, acc)
            ctx.finalize_wave(wv, taken)
            taken += wv
        hdr, ldr, st = ctx.resolve()
        t_hdr, t_ldr = torch.from_numpy(hdr), torch.from_numpy(ldr)  # identical on every rank
    else:
        # sample-wave sharding (bench.py's N > 1 path): rank r renders wave r of spp samples of the
        # whole frame; equal wave sizes → finishTile's weights collapse to 1 / world
        ctx.begin_frame(w, h, spp * world, 16, (0, 0, 0), Y.TONEMAP_AGX)
        ctx.render_wave(rank * spp, spp, 0)
        hdr, _, st = ctx.resolve()
        t_hdr = torch.from_numpy(hdr)
        t_hdr.mul_(1.0 / world)
        dist.all_reduce(t_hdr)
        t_ldr = None
    rays = torch.tensor([st.raysReference], dtype=torch.int64)
    dist.all_reduce(rays)
    if rank == 0:
        np.save(os.path.join(out_dir, f"{mode}_hdr.npy"), t_hdr.numpy())
        if t_ldr is not None:
            np.save(os.path.join(out_dir, f"{mode}_ldr.npy"), t_ldr.numpy())
        np.save(os.path.join(out_dir, f"{mode}_rays.npy"), rays.numpy())
    dist.barrier()
    dist.destroy_process_group()


def run_nccl(rank: int, world: int, mode: str, out_dir: str):
    """`-m gpu`, one process per GPU on the real library: yr_create_dist with NCCL inside libyart_b200.so; gloo only
    carries the communicator id.  nccl_tiles: the finalize kernels store into rank 0's frame through the mapped
    allocation (direct); nccl_tiles_reduce: the same with YART_B200_FRAMES_REDUCE=1 (ncclReduce of the frames)."""
    import torch.distributed as dist
    import harness as H
    import yart_b200 as Y
    if mode == "nccl_tiles_reduce":
        os.environ["YART_B200_FRAMES_REDUCE"] = "1"
    box = [Y.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cam = H.scene_camera("material_zoo")
    sc = Y.Scene(H.scene_file("material_zoo"))
    w, h = 160, 90
    c = Y.make_camera(w, h, cam["focal"], cam["fnum"], cam["pos"], cam["target"], (0, 0, 0), cam["exposure"])
    r = Y.Renderer(w, h, c, sc, tile_size=16, tonemap=Y.TONEMAP_AGX, samples=56, first_wave_samples=8, max_wave_samples=16,
                   max_depth=6, traversal=Y.TRAVERSAL_REFERENCE_ORDER, device=rank, dist=(rank, world, box[0]))
    per_wave = []
    target = np.zeros((h, w, 4), np.float32)
    r.set_frame_target(target)
    r.on_wave_complete(lambda rd, wd: per_wave.append(target.copy()))
    for k in range(2):  # the second render reuses the mapped frames
        data = r.render_sync()
        hdr, ldr, _ = r.read()
        if rank == 0:
            np.save(os.path.join(out_dir, f"{mode}_hdr{k}.npy"), hdr)
            np.save(os.path.join(out_dir, f"{mode}_ldr{k}.npy"), ldr)
    if rank == 0:
        np.save(os.path.join(out_dir, f"{mode}_waves.npy"), np.stack(per_wave))
        np.save(os.path.join(out_dir, f"{mode}_info.npy"), np.array([data["total_rays"], int(r.frames_direct())], np.int64))
    r.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
