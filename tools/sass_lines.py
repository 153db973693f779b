"""Attribute an ncu capture's per-instruction counters to source lines.

  python tools/sass_lines.py report.ncu-rep yart_b200/libyart_b200.so [--top 40] [--by file|line|func]

ncu's CSV export of the source page carries per-SASS-instruction metrics but no source mapping; nvdisasm -g on the
cubin of the same build carries the (innermost inlined) file:line of every instruction.  Both list the kernel's
instructions in address order, so they join on the address.  Prints executed warp instructions, stall samples and the
dominant stall reasons per source line (or per file).
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, index):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True, check=True).stdout.splitlines()
    starts = [k for k, ln in enumerate(out) if ln.startswith('"Kernel Name"')] + [len(out)]
    out = out[starts[index]:starts[index + 1]]
    kernel = next(csv.reader([out[0]]))[1]
    rows = [r for r in csv.reader(out[1:]) if r]
    return kernel, rows[0], rows[1:]


def line_map(lib, mangled_part):
    tmp = tempfile.mkdtemp(prefix="sasslines")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cub = [f for f in os.listdir(tmp) if "sm_100" in f][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    m = {}
    cur = None
    inside = False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = mangled_part in ln
            cur = None
            continue
        if not inside:
            continue
        f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            continue
        a = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if a:
            m[int(a.group(1), 16)] = cur
    return m


def mangled_hint(kernel):
    # "void yb::rt::forKernel<ShadeSurfaceK<(bool)1>, (int)8>(...)" -> pieces that appear in the mangled name
    names = re.findall(r"[A-Za-z_][A-Za-z0-9_]*", kernel)
    for n in names:
        if n.endswith("K") or n.endswith("Kernel"):
            if n != "forKernel":
                return n
    return names[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("lib")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by", default="line", choices=["line", "file"])
    ap.add_argument("--index", type=int, default=0, help="which captured launch of the report")
    ap.add_argument("--match", default=None, help="substring of the mangled kernel name (default: derived)")
    a = ap.parse_args()
    kernel, head, rows = sass_rows(a.report, a.index)
    hint = a.match or mangled_hint(kernel)
    # template arguments disambiguate ShadeSurfaceK<0> / <1>
    lm = {}
    for cand in ([hint + "ILb1", hint + "ILb0"] if "(bool)" in kernel else [hint]):
        if "(bool)1" in kernel and cand.endswith("ILb0"):
            continue
        if "(bool)0" in kernel and cand.endswith("ILb1"):
            continue
        lm = line_map(a.lib, cand)
        if lm:
            break
    if not lm:
        lm = line_map(a.lib, hint)
    col = {n: i for i, n in enumerate(head)}
    stalls = [n for n in head if n.startswith("stall_") and "Not Issued" not in n]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    base = None
    for r in rows:
        addr = int(r[col["Address"]], 16)
        if base is None:
            base = addr
        key = lm.get(addr - base)
        if key is None:
            key = ("?", 0)
        if a.by == "file":
            key = (key[0], 0)
        c = agg[key]
        vals = {"inst": int(r[col["Instructions Executed"]] or 0), "samples": int(r[col["# Samples"]] or 0),
                "thr": int(r[col["Thread Instructions Executed"]] or 0),
                "local": int(r[col["L2 Theoretical Sectors Local"]] or 0),
                "glob": int(r[col["L2 Theoretical Sectors Global"]] or 0), "n": 1}
        for s in stalls:
            vals[s] = int(r[col[s]] or 0)
        for k, v in vals.items():
            c[k] += v
            tot[k] += v
    print(f"{kernel}\n{tot['n']} SASS instructions, {tot['inst']} warp instructions executed, {tot['samples']} samples, "
          f"lanes/inst {tot['thr'] / max(1, tot['inst']):.1f}, L2 sectors global {tot['glob']} local {tot['local']}")
    print("stalls: " + ", ".join(f"{s[6:]} {100 * tot[s] / max(1, tot['samples']):.1f}%" for s in
                                 sorted(stalls, key=lambda s: -tot[s])[:8]))
    print(f"{'where':32s} {'sass':>6s} {'inst%':>6s} {'smp%':>6s} {'lanes':>5s} {'glob%':>6s} {'loc%':>6s}  top stalls")
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:a.top]:
        top = ", ".join(f"{s[6:]} {100 * c[s] / max(1, c['samples']):.0f}%" for s in sorted(stalls, key=lambda s: -c[s])[:3])
        where = key[0] if a.by == "file" else f"{key[0]}:{key[1]}"
        print(f"{where:32s} {c['n']:6d} {100 * c['inst'] / max(1, tot['inst']):6.1f} {100 * c['samples'] / max(1, tot['samples']):6.1f} "
              f"{c['thr'] / max(1, c['inst']):5.1f} {100 * c['glob'] / max(1, tot['glob']):6.1f} {100 * c['local'] / max(1, tot['local']):6.1f}  {top}")


if __name__ == "__main__":
    sys.exit(main())
