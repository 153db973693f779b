"""The reference-side binding, compiled: integration/wavefront-renderer.hpp (yart::cuda::WavefrontRenderer :
yart::Renderer + SceneFlattener) built against the UNMODIFIED reference by oracle/Makefile (`make -C oracle adapter`).

oracle/_ref/adapter_* builds a yart::Scene through yart's own API, renders it with the reference's
cpu::TileRenderer<SobolSampler<FastOwenScrambler>, MISIntegrator> and with the adapter IN THE SAME PROCESS, and reports
how many words of the LDR frame (Renderer::m_buffer), of the HDR accumulation, and of an asynchronous re-render differ,
plus ray counts and callback counts.  adapter_hostsim links the CPU build of the product sources (runs in this suite);
adapter_cuda links libyart_b200.so (`-m gpu`)."""
import json
import os
import subprocess

import pytest

import harness as H

ADAPTER_CPU = os.path.join(H.ROOT, "oracle", "_ref", "adapter_hostsim")
ADAPTER_GPU = os.path.join(H.ROOT, "oracle", "_ref", "adapter_cuda")

CASES = {
    # scene, extra arguments
    "cornell_one_wave": ("cornell", ["spp=8", "maxdepth=5"]),
    "cornell_progressive_tiles": ("cornell", ["spp=12", "first=1", "max=4", "w=40", "h=40", "tile=16"]),
    "zoo_alpha_transforms_textures": ("material_zoo", ["spp=4", "maxdepth=6", "w=48", "h=32", "tile=8"]),
    "sponza_small_env_light": ("sponza", ["spp=2", "maxdepth=4", "w=48", "h=27"]),
    "mclaren_small_golden_look": ("mclaren", ["spp=2", "maxdepth=8", "w=48", "h=27", "tonemap=golden"]),
    "two_quads_no_tonemapper": ("two_quads", ["spp=4", "tonemap=none"]),
}
SCENE_KW = {"sponza": dict(n_tris=3000, tex_res=32, env_res=32), "mclaren": dict(n_tris=6000, env_res=32)}


def run_adapter(binary, scene, extra, timeout=600):
    H.hostsim()
    kw = SCENE_KW.get(scene, {})
    sp, cam = H.scene_file(scene, **kw), H.scene_camera(scene)
    out = os.path.join(H.CACHE, f"adapter_{os.getpid()}.bin")
    args = [binary, sp, out, "pos=%.9g,%.9g,%.9g" % tuple(cam["pos"]), "target=%.9g,%.9g,%.9g" % tuple(cam["target"]),
            f"focal={cam['focal']}", f"fnum={cam['fnum']}", f"exposure={cam['exposure']}", "threads=1"] + extra
    r = subprocess.run(args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def check(j, exact=True):
    if exact:
        assert j["ldr_words_differing"] == 0 and j["hdr_words_differing"] == 0, j
        assert j["rays_adapter"] == j["rays_reference"], j
    assert j["async_words_differing"] == 0, j                    # render() + wait() reproduces renderSync()
    assert j["wave_ray_sum"] >= j["rays_adapter"], j             # WaveData::rays add up (two renders by then)
    assert j["waves_adapter"] == j["waves_reference"] and j["tiles_adapter"] == j["tiles_reference"], j
    assert j["done_adapter"] == j["done_reference"] == 1 and j["one_completion_per_render"] is True, j


needs_cpu_adapter = pytest.mark.skipif(not os.path.exists(ADAPTER_CPU), reason="oracle/_ref/adapter_hostsim not built")
needs_gpu_adapter = pytest.mark.skipif(not os.path.exists(ADAPTER_GPU), reason="oracle/_ref/adapter_cuda not built")


@needs_cpu_adapter
@pytest.mark.parametrize("case", sorted(CASES))
def test_adapter_equals_tile_renderer_in_process_cpu_build(case):
    scene, extra = CASES[case]
    check(run_adapter(ADAPTER_CPU, scene, extra))


@needs_cpu_adapter
def test_adapter_over_three_shards_cpu_build():
    """WavefrontRenderer::devices with three entries → yr_create_multi_flat: same frame, same callbacks."""
    j = run_adapter(ADAPTER_CPU, "cornell", ["spp=12", "first=1", "max=4", "w=40", "h=40", "tile=16", "devices=0,1,2"])
    check(j)
    assert j["devices"] == 3


@pytest.mark.gpu
@needs_gpu_adapter
@pytest.mark.parametrize("case", sorted(CASES))
def test_adapter_equals_tile_renderer_in_process_gpu(case):
    scene, extra = CASES[case]
    check(run_adapter(ADAPTER_GPU, scene, extra))


@pytest.mark.gpu
@needs_gpu_adapter
def test_adapter_wide_walk_and_two_contexts_gpu():
    """The default traversal (wide BVH for scenes without alpha-tested materials) through the adapter, and two
    contexts (tile shards) on the GPUs present (the same GPU twice on a one-GPU box: in-process group transport)."""
    import torch
    j = run_adapter(ADAPTER_GPU, "cornell", ["spp=8", "maxdepth=5", "traversal=0", "w=96", "h=96"])
    check(j, exact=False)
    assert abs(j["rays_adapter"] - j["rays_reference"]) <= 1e-3 * j["rays_reference"]
    assert j["ldr_words_differing"] <= 0.005 * 96 * 96 * 4
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    j = run_adapter(ADAPTER_GPU, "cornell", ["spp=12", "first=1", "max=4", "w=40", "h=40", "tile=16", f"devices={devs}"])
    check(j)
