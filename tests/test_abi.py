"""The C-ABI boundary: libyart_b200.so loads, exports every entry point include/yart_cuda.h declares,
and refuses to work without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import harness as H
from yart_b200 import capi

HEADER = os.path.join(H.ROOT, "include", "yart_cuda.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ \*]*?\b(y[csr]_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ("yc_create", "yc_upload_scene", "yc_set_camera", "yc_begin_frame", "yc_render_wave", "yc_resolve",
                 "yc_trace", "yc_trace_device", "yc_kat", "ys_scene_load", "ys_camera_make", "yr_create", "yr_render",
                 "yr_abort", "yr_wait", "yr_render_sync", "yr_read"):
        assert must in fns
    assert len(fns) >= 35


def test_product_library_exports_every_declared_symbol():
    assert os.path.exists(capi.PRODUCT_LIB), "libyart_b200.so not built: run __graft_entry__.build()"
    lib = C.CDLL(capi.PRODUCT_LIB)
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, f"declared in include/yart_cuda.h but not exported: {missing}"


def test_samplers_build_exports_the_same_abi():
    assert os.path.exists(capi.SAMPLERS_LIB), "libyart_b200_samplers.so not built: run __graft_entry__.build()"
    lib = C.CDLL(capi.SAMPLERS_LIB)
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, f"declared in include/yart_cuda.h but not exported by the samplers build: {missing}"


def test_python_prototypes_cover_the_header():
    assert sorted(capi.PROTOTYPES) == declared_functions()
    capi.load()  # binds restype/argtypes of all of them


def test_struct_sizes_match_the_c_definitions():
    # sizes the C compiler produces for the PODs crossing the boundary (checked against sizeof in C)
    import subprocess
    import tempfile
    prog = r'''
#include <stdio.h>
#include "yart_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(YcCamera), sizeof(YcOptions), sizeof(YcRect),
         sizeof(YcFrameDesc), sizeof(YcStats), sizeof(YcRay), sizeof(YcHit), sizeof(YcMesh), sizeof(YcScene),
         sizeof(YrSettings), sizeof(YrRenderData), sizeof(YrWaveData), sizeof(YrTileData));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(H.ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    py = [C.sizeof(t) for t in (capi.YcCamera, capi.YcOptions, capi.YcRect, capi.YcFrameDesc, capi.YcStats, capi.YcRay,
                                capi.YcHit, capi.YcMesh, capi.YcScene, capi.YrSettings, capi.YrRenderData, capi.YrWaveData,
                                capi.YrTileData)]
    assert py == sizes


def test_no_device_means_an_error_not_a_fallback():
    """In a container without a GPU the product refuses to create a context (YC_ERR_NO_DEVICE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.load()
    h = C.c_void_p()
    opts = capi.YcOptions()
    rc = lib.yc_create(0, C.byref(opts), C.byref(h))
    assert rc == capi.YC_ERR_NO_DEVICE and not h.value
    import yart_b200 as Y
    Y.use_library(lib)
    with pytest.raises(Y.YartError):
        Y.Context()
    cam = Y.make_camera(8, 8)  # host-only call still works
    with pytest.raises(Y.YartError):
        Y.Renderer(8, 8, cam)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load(str(tmp_path / "libyart_b200.so"))
