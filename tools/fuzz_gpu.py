"""Extra randomised-scene parity seeds on the GPU (same case generator as tests/test_fuzz_parity.py): the CUDA
library against the reference run on the spot, whole frames bit for bit.  python tools/fuzz_gpu.py [first] [count]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import yart_b200 as Y  # noqa: E402
from yart_b200 import capi  # noqa: E402
import test_fuzz_parity as F  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = []
Y.use_library(capi.load())
for seed in range(first, first + count):
    for kw in ({}, {"integrator": "naive"} if seed % 4 == 0 else None):
        if kw is None:
            continue
        try:
            F.run_case(seed, **kw)
        except AssertionError as e:
            bad.append((seed, kw, str(e)[:200]))
Y.use_library(capi.load(capi.SAMPLERS_LIB))
for seed in range(first, first + count, 5):
    for kw in ({"scrambler": "owen"}, {"sampler": "stratified"}, {"sampler": "naive"}):
        try:
            F.run_case(seed, **kw)
        except AssertionError as e:
            bad.append((seed, kw, str(e)[:200]))
print(f"fuzz seeds {first}..{first + count - 1}: {len(bad)} failures")
for b in bad:
    print("  ", b)
sys.exit(1 if bad else 0)
