"""The SAH BVH built by the device layer (csrc/bvh_build.cuh: level-synchronous, atomics for the folds, the
reference's in-place partition as a scan + two scatters) against the host builder, which reproduces the reference's
tree (tests/test_host_logic.py, goldens bvh_*.npz): node arrays and index arrays identical, byte for byte.
CPU: the same sources compiled for the CPU (stages as loops).  `-m gpu`: the CUDA kernels, up to 1 M triangles."""
import time

import numpy as np
import pytest

import harness as H
import yart_b200 as Y

CASES = [("cornell", {}), ("two_quads", {}), ("material_zoo", {}), ("soup", dict(n_tris=20000)), ("soup", dict(n_tris=49, with_light=False)),
         ("sponza", dict(n_tris=30000, tex_res=8, env_res=8)), ("mclaren", dict(n_tris=60000, env_res=8)),
         ("degenerate_soup", dict(seed=1)), ("degenerate_soup", dict(seed=2, n_tris=20000)), ("degenerate_soup", dict(seed=3, n_tris=700)),
         ("random_scene", dict(seed=5)), ("random_scene", dict(seed=11))]


def compare(name, kw, kind_b=None):
    path = H.scene_file(name, **kw)
    a = Y.Scene(path, bvh_kind=Y.BVH_SAH_HOST)
    t0 = time.time()
    b = Y.Scene(path, bvh_kind=Y.BVH_SAH_DEVICE if kind_b is None else kind_b)
    dt = time.time() - t0
    assert a.device_builds == 0
    m = 0
    while True:
        try:
            na, ia = a.bvh(m)
        except Y.YartError:
            break
        nb, ib = b.bvh(m)
        assert np.array_equal(ia, ib), f"{name} mesh {m}: index order differs"
        assert na.tobytes() == nb.tobytes(), f"{name} mesh {m}: nodes differ"
        m += 1
    assert m >= 1
    return a, b, dt


@pytest.mark.parametrize("name,kw", CASES, ids=[f"{n}-{i}" for i, (n, _) in enumerate(CASES)])
def test_device_layer_builder_equals_host_builder_cpu_build(name, kw, hostsim_lib):
    _, b, _ = compare(name, kw)
    assert b.device_builds >= 1


def test_automatic_choice_stays_on_the_host_in_the_cpu_build(hostsim_lib):
    sc = Y.Scene(H.scene_file("soup", n_tris=40000))
    assert sc.device_builds == 0


GPU_CASES = CASES + [("soup", dict(n_tris=400000))]


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", GPU_CASES, ids=[f"{n}-{i}" for i, (n, _) in enumerate(GPU_CASES)])
def test_gpu_builder_equals_host_builder(name, kw, cuda_lib):
    _, b, _ = compare(name, kw)
    assert b.device_builds >= 1


@pytest.mark.gpu
def test_gpu_builder_one_million_triangles_and_the_automatic_choice(cuda_lib, monkeypatch):
    """C2's scene: the default kind builds the 1 M-triangle mesh on the GPU (the two-triangle light on the host), the tree
    is the host builder's; YART_B200_BVH_DEVICE=0 or ys_set_build_device(-1) keep everything on the host."""
    a, b, dt = compare("soup", dict(n_tris=1_000_000), kind_b=Y.BVH_SAH)
    assert b.device_builds == 1
    print(f"1 M triangles: scene build on the host {a.build_ms:.0f} ms, with the GPU builder {b.build_ms:.0f} ms")
    path = H.scene_file("soup", n_tris=1_000_000)
    monkeypatch.setenv("YART_B200_BVH_DEVICE", "0")
    assert Y.Scene(path).device_builds == 0
    monkeypatch.delenv("YART_B200_BVH_DEVICE")
    Y.set_build_device(-1)
    try:
        assert Y.Scene(path).device_builds == 0
    finally:
        Y.set_build_device(0)


def test_build_entry_point_rejects_bad_input(hostsim_lib):
    import ctypes as C
    from yart_b200 import lib
    pos = np.zeros((3, 3), np.float32)
    pos[1, 0] = pos[2, 1] = 1.0
    faces = np.array([[0, 1, 2, 0]], np.uint32)
    nodes = np.zeros((4, 8), np.uint32)
    idx = np.zeros(1, np.uint32)
    n, lv = C.c_uint32(), C.c_uint32()
    f = lib().yc_build_bvh_sah
    assert f(0, pos.ctypes.data, 3, faces.ctypes.data, 1, nodes.ctypes.data, C.byref(n), idx.ctypes.data, C.byref(lv)) == 0
    assert n.value == 1 and nodes[0, 7] == 1 and nodes[0, 6] == 0 and idx[0] == 0  # one leaf over the one triangle
    assert f(0, pos.ctypes.data, 3, faces.ctypes.data, 0, nodes.ctypes.data, C.byref(n), idx.ctypes.data, C.byref(lv)) == -1
    bad = np.array([[0, 1, 3, 0]], np.uint32)
    assert f(0, pos.ctypes.data, 3, bad.ctypes.data, 1, nodes.ctypes.data, C.byref(n), idx.ctypes.data, C.byref(lv)) == -1
    assert b"out of range" in lib().yc_build_last_error()
    assert f(0, None, 3, faces.ctypes.data, 1, nodes.ctypes.data, C.byref(n), idx.ctypes.data, C.byref(lv)) == -1
