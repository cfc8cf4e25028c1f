#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals of one profiled kernel, without a GPU:
    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel mangled-name substring> [--top 40] [--lib pgvector_b200/libvecb200.so]
ncu's SASS page (addresses, instructions executed, stall samples) is joined with `nvdisasm -g` of the cubin inside the
shared library (offset -> file:line, innermost inlined frame), because the CUDA source page of `ncu -i` carries no
metrics when the report is read on another machine."""
import argparse
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "header": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["header"] is None:
            cur["header"] = row
        elif cur is not None:
            cur["rows"].append(row)
    return kernels


def line_map(lib, want):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or "sm_100" not in f:
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        fn, m, loc = None, {}, None
        found = {}
        for ln in dis.split("\n"):
            mm = re.match(r"^\.text\.(\S+):", ln)
            if mm:
                fn, loc = mm.group(1), None
                found[fn] = {}
                continue
            mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if mm:
                loc = (os.path.basename(mm.group(1)), int(mm.group(2)))
                continue
            mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if mm and fn:
                found[fn][int(mm.group(1), 16)] = (loc, mm.group(2).strip())
        for k, v in found.items():
            if want in k and v:
                return k, v
    return None, {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--lib", default="pgvector_b200/libvecb200.so")
    ap.add_argument("--launch", type=int, default=0, help="which profiled launch of the report")
    a = ap.parse_args()
    ks = sass_page(a.rep)
    if not ks:
        sys.exit("no SASS page in the report (profile with --import-source on / --set full)")
    k = ks[min(a.launch, len(ks) - 1)]
    h = {n: i for i, n in enumerate(k["header"])}
    want = a.kernel
    if "ILi" not in want:
        # the report names the instantiation ("kernel<(int)2, (int)4, (int)8>"): match it, not the first one of the library
        mm = re.search(r"<([^>]*)>", k["name"])
        if mm:
            args = re.findall(r"\(int\)(-?\d+)", mm.group(1))
            if args:
                want = a.kernel + "I" + "".join(f"Li{v}E" if not v.startswith("-") else f"Lin{v[1:]}E" for v in args) + "E"
    fn, lm = line_map(a.lib, want)
    if not lm:
        sys.exit(f"no function matching {a.kernel} in {a.lib}")
    base = int(k["rows"][0][h["Address"]], 16)
    per = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_i = tot_s = 0
    for r in k["rows"]:
        off = int(r[h["Address"]], 16) - base
        loc, txt = lm.get(off, (None, r[h["Source"]].strip()))
        ins = int(r[h["Instructions Executed"]] or 0)
        smp = int(r[h["Warp Stall Sampling (All Samples)"]] or 0)
        nis = int(r[h["Warp Stall Sampling (Not-issued Samples)"]] or 0)
        p = per[loc]
        p[0] += ins
        p[1] += smp
        p[2] += nis
        p[3][txt.split()[0] if not txt.startswith("@") else txt.split()[1]] += ins
        tot_i += ins
        tot_s += smp
    print(f"kernel: {k['name'][:120]}\nfunction: {fn}\nwarp instructions executed: {tot_i}, stall samples: {tot_s}\n")
    print("| file:line | instructions | share | stall samples | share | top opcodes |")
    print("|---|---|---|---|---|---|")
    for loc, p in sorted(per.items(), key=lambda kv: -kv[1][1])[:a.top]:
        ops = ", ".join(f"{o} {c * 100 // max(p[0], 1)}%" for o, c in sorted(p[3].items(), key=lambda kv: -kv[1])[:3])
        name = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"| {name} | {p[0]} | {100.0 * p[0] / max(tot_i, 1):.1f} % | {p[1]} | {100.0 * p[1] / max(tot_s, 1):.1f} % | {ops} |")


if __name__ == "__main__":
    main()
