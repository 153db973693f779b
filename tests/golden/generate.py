"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/oracle_ref,
built by oracle/Makefile from /root/reference) on seeded inputs.

    python tests/golden/generate.py

Each file stores the exact input (KAT blob / rays / render settings) next to the reference's
output, so tests can replay the input through the CUDA path (or the hostsim build) and compare.
The scenes are produced by yart_b200/scenes.py (deterministic, seeded).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness as H  # noqa: E402

OUT = H.GOLDEN

KATS = [  # (file tag, kind, n, kwargs, scene)
    ("sampler16", "sampler", 512, dict(spp=16), None),
    ("sampler128", "sampler", 512, dict(spp=128), None),
    ("sampler1024", "sampler", 512, dict(spp=1024), None),
    ("sampler4096", "sampler", 512, dict(spp=4096), None),
    ("sampler3", "sampler", 256, dict(spp=3), None),      # log2Int rounds in log space: 3 → 2
    ("sampler12", "sampler", 256, dict(spp=12), None),    # 12 → 4
    ("sampler100", "sampler", 256, dict(spp=100), None),  # 100 → 7
    ("ggx", "ggx", 1024, {}, None),
    ("gmon16", "gmon", 256, dict(samples=16), None),
    ("gmon64", "gmon", 256, dict(samples=64), None),
    ("gmon128", "gmon", 128, dict(samples=128), None),
    ("gmon3", "gmon", 64, dict(samples=3), None),
    ("gmonb16", "gmonb", 256, dict(samples=16), None),
    ("gmonb128", "gmonb", 256, dict(samples=128), None),
    ("agx0", "agx", 1024, dict(look=0), None),
    ("agx1", "agx", 512, dict(look=1), None),
    ("agx2", "agx", 512, dict(look=2), None),
    ("camera", "camera", 512, dict(), None),
    ("camera_poly", "camera", 512, dict(sides=6, fnum=1.4), None),
    ("camera_pinhole", "camera", 512, dict(fnum=0.0, w=1920, h=1080, pos=(0, 0, 40), target=(0, 0, 0)), None),
    ("lut", "lut", 2048, {}, "material_zoo"),
    ("texture", "texture", 2048, dict(n_textures=8), "material_zoo"),
    ("bsdf", "bsdf", 4096, dict(n_materials=17), "material_zoo"),
    ("light", "light", 2048, dict(n_lights=6), "material_zoo"),
]

SPONZA_S = dict(n_tris=6000, tex_res=32, env_res=32, n_materials=9)
MCLAREN_S = dict(n_tris=6000, env_res=32)
TRACE_SCENES = [("cornell", {}), ("material_zoo", {}), ("soup", dict(n_tris=20000)), ("two_quads", {}),
                ("sponza", SPONZA_S), ("mclaren", MCLAREN_S)]

RENDERS = [  # (tag, scene, scene kwargs, w, h, spp, first, max, maxdepth, tonemap)
    ("two_quads", "two_quads", {}, 64, 64, 16, 16, 16, 30, "agx"),
    ("cornell", "cornell", {}, 64, 64, 16, 16, 16, 30, "agx"),
    ("cornell_waves", "cornell", {}, 48, 40, 8, 1, 4, 30, "agx"),
    ("zoo", "material_zoo", {}, 96, 64, 16, 16, 16, 30, "agx"),
    ("zoo_waves", "material_zoo", {}, 48, 32, 64, 4, 32, 30, "golden"),
    ("soup_d1", "soup", dict(n_tris=20000), 96, 54, 4, 4, 4, 1, "agx"),
    ("sponza_small", "sponza", SPONZA_S, 96, 54, 16, 16, 16, 30, "agx"),   # C3 shape: textured PBR + env only
    ("mclaren_small", "mclaren", MCLAREN_S, 96, 54, 16, 4, 8, 30, "punchy"),  # C4 shape: coat / glass / volume
]


def trace_rays(cam, n, seed=3, spread=12.0):
    rng = np.random.default_rng(seed)
    o = np.tile(np.asarray(cam["pos"], np.float32), (n, 1)) + rng.normal(0, 0.5, (n, 3)).astype(np.float32)
    tgt = np.asarray(cam["target"], np.float32) + rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((n, 8), np.float32)
    r[:, 0:3], r[:, 3], r[:, 4:7], r[:, 7] = o, 0.001, d, rng.uniform(5, 40, n)
    k = n // 3  # rays born inside the scene, any direction
    r[:k, 0:3] = rng.uniform(-4, 4, (k, 3))
    r[:k, 1] = np.abs(r[:k, 1]) + 0.1
    dd = rng.normal(size=(k, 3))
    r[:k, 4:7] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    r[-8:, 4:7] = np.eye(3, dtype=np.float32)[np.arange(8) % 3]  # axis-aligned: zero components → inf idir
    return r


# TileRenderer<SobolSampler<FastOwenScrambler>, NaiveIntegrator>: tag, scene, kwargs, w, h, spp, first, max, depth
NAIVE_RENDERS = [
    ("cornell", "cornell", {}, 64, 64, 16, 16, 16, 5),
    ("zoo", "material_zoo", {}, 96, 54, 16, 8, 8, 8),
    ("two_quads", "two_quads", {}, 32, 32, 16, 16, 16, 30),
]


def variant_kats():
    """KATs of code that only the variants build carries (libyart_b200_samplers.so / hostsim): UniformLightSampler."""
    blob = H.kat_input("lightuniform", 2048, **dict(n_lights=6))
    out = H.oracle_kat("lightuniform", blob, H.scene_file("material_zoo"))
    np.savez_compressed(os.path.join(OUT, "variantkat_lightuniform.npz"), kind="lightuniform", scene="material_zoo",
                        blob=np.frombuffer(blob, np.uint8), out=out)
    print("variant kat lightuniform", out.size)


def median_bvhs():
    """MedianSplitBVH (bvh.hpp:237-264) over the same meshes."""
    for name, kw in [("cornell", {}), ("material_zoo", {}), ("soup", dict(n_tris=3000))]:
        meshes = H.oracle_bvh(H.scene_file(name, **kw), kind="median")
        np.savez_compressed(os.path.join(OUT, f"medianbvh_{name}.npz"), scene=name, kwargs=repr(kw),
                            **{f"nodes{i}": m[0] for i, m in enumerate(meshes)},
                            **{f"idx{i}": m[1] for i, m in enumerate(meshes)})
        print("median bvh", name, [len(m[0]) for m in meshes])


def naive_renders():
    for tag, name, kw, w, h, spp, first, mx, depth in NAIVE_RENDERS:
        sp, cam = H.scene_file(name, **kw), H.scene_camera(name, **kw)
        r = H.oracle_render(sp, w, h, spp, cam, first=first, max=mx, maxdepth=depth, tonemap="agx", threads=4,
                            integrator="naive")
        np.savez_compressed(os.path.join(OUT, f"naive_{tag}.npz"), scene=name, kwargs=repr(kw),
                            settings=np.array([w, h, spp, first, mx, depth]), tonemap="agx", hdr=r["hdr"], ldr=r["ldr"],
                            rays=r["rays"], integrator="naive")
        print("naive render", tag, r["rays"])


# TileRenderer<SobolSampler<OwenScrambler | BinaryPermuteScrambler>, MISIntegrator>
SCRAMBLER_RENDERS = [
    ("owen_cornell", "owen", "cornell", {}, 64, 64, 16, 16, 16, 6),
    ("owen_zoo", "owen", "material_zoo", {}, 96, 54, 12, 4, 8, 8),
    ("binary_cornell", "binary", "cornell", {}, 64, 64, 16, 16, 16, 6),
    ("binary_zoo", "binary", "material_zoo", {}, 96, 54, 12, 4, 8, 8),
]


# TileRenderer<NaiveSampler | StratifiedSampler, MISIntegrator>
SAMPLER_RENDERS = [
    ("naive_cornell", "naive", "cornell", {}, 64, 64, 16, 16, 16, 6),
    ("naive_zoo", "naive", "material_zoo", {}, 96, 54, 12, 4, 8, 8),
    ("stratified_cornell", "stratified", "cornell", {}, 64, 64, 16, 16, 16, 6),
    ("stratified_zoo", "stratified", "material_zoo", {}, 96, 54, 12, 4, 8, 8),
]


def sampler_renders():
    for tag, smp, name, kw, w, h, spp, first, mx, depth in SAMPLER_RENDERS:
        sp, cam = H.scene_file(name, **kw), H.scene_camera(name, **kw)
        r = H.oracle_render(sp, w, h, spp, cam, first=first, max=mx, maxdepth=depth, tonemap="agx", threads=4, sampler=smp)
        np.savez_compressed(os.path.join(OUT, f"sampler_{tag}.npz"), scene=name, kwargs=repr(kw),
                            settings=np.array([w, h, spp, first, mx, depth]), tonemap="agx", hdr=r["hdr"], ldr=r["ldr"],
                            rays=r["rays"], sampler=smp)
        print("sampler render", tag, r["rays"])


def scrambler_renders():
    for tag, scr, name, kw, w, h, spp, first, mx, depth in SCRAMBLER_RENDERS:
        sp, cam = H.scene_file(name, **kw), H.scene_camera(name, **kw)
        r = H.oracle_render(sp, w, h, spp, cam, first=first, max=mx, maxdepth=depth, tonemap="agx", threads=4, scrambler=scr)
        np.savez_compressed(os.path.join(OUT, f"scrambler_{tag}.npz"), scene=name, kwargs=repr(kw),
                            settings=np.array([w, h, spp, first, mx, depth]), tonemap="agx", hdr=r["hdr"], ldr=r["ldr"],
                            rays=r["rays"], scrambler=scr)
        print("scrambler render", tag, r["rays"])


def main():
    assert H.have_oracle(), "build oracle/_ref first: make -C oracle ref"
    if "--only-naive" in sys.argv:
        return naive_renders()
    if "--only-scramblers" in sys.argv:
        return scrambler_renders()
    if "--only-samplers" in sys.argv:
        return sampler_renders()
    if "--only-median-bvh" in sys.argv:
        return median_bvhs()
    if "--only-variant-kats" in sys.argv:
        return variant_kats()
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only-kat=")]
    if only:  # regenerate single KAT fixtures: --only-kat=gmonb16
        for tag, kind, n, kw, scene in KATS:
            if tag in only:
                blob = H.kat_input(kind, n, **kw)
                out = H.oracle_kat(kind, blob, H.scene_file(scene) if scene else None)
                np.savez_compressed(os.path.join(OUT, f"kat_{tag}.npz"), kind=kind, scene=scene or "",
                                    blob=np.frombuffer(blob, np.uint8), out=out)
                print("kat", tag, out.size)
        return
    for tag, kind, n, kw, scene in KATS:
        blob = H.kat_input(kind, n, **kw)
        sp = H.scene_file(scene) if scene else None
        out = H.oracle_kat(kind, blob, sp)
        np.savez_compressed(os.path.join(OUT, f"kat_{tag}.npz"), kind=kind, scene=scene or "",
                            blob=np.frombuffer(blob, np.uint8), out=out)
        print("kat", tag, out.size)
    for name, kw in TRACE_SCENES:
        sp, cam = H.scene_file(name, **kw), H.scene_camera(name, **kw)
        rays = trace_rays(cam, 3000)
        closest = H.oracle_trace(sp, rays, "closest")
        anyhit = H.oracle_trace(sp, rays, "any")
        np.savez_compressed(os.path.join(OUT, f"trace_{name}.npz"), scene=name, kwargs=repr(kw), rays=rays,
                            closest=closest, anyhit=anyhit)
        print("trace", name, int(closest["didHit"].sum()), int(anyhit["didHit"].sum()))
    for tag, name, kw, w, h, spp, first, mx, depth, tm in RENDERS:
        sp, cam = H.scene_file(name, **kw), H.scene_camera(name, **kw)
        r = H.oracle_render(sp, w, h, spp, cam, first=first, max=mx, maxdepth=depth, tonemap=tm, threads=4)
        np.savez_compressed(os.path.join(OUT, f"render_{tag}.npz"), scene=name, kwargs=repr(kw),
                            settings=np.array([w, h, spp, first, mx, depth]), tonemap=tm, hdr=r["hdr"], ldr=r["ldr"],
                            rays=r["rays"])
        print("render", tag, r["rays"])
    for name, kw in [("cornell", {}), ("material_zoo", {}), ("soup", dict(n_tris=3000))]:
        sp = H.scene_file(name, **kw)
        meshes = H.oracle_bvh(sp)
        np.savez_compressed(os.path.join(OUT, f"bvh_{name}.npz"), scene=name, kwargs=repr(kw),
                            **{f"nodes{i}": m[0] for i, m in enumerate(meshes)},
                            **{f"idx{i}": m[1] for i, m in enumerate(meshes)})
        print("bvh", name, [len(m[0]) for m in meshes])
    median_bvhs()
    variant_kats()
    naive_renders()
    scrambler_renders()
    sampler_renders()


if __name__ == "__main__":
    main()
