#!/usr/bin/env python
"""Which SOURCE LINES of a kernel make its memory requests: joins ncu's per-instruction source page (`ncu -i X.ncu-rep --page source
--csv | gzip`) with the line table of the shipped library (`nvdisasm -g` of the cubin inside bling_b200/libblingcu.so: same build,
same addresses) and prints, per source line, warp instructions executed, L1 tag requests, L2 sectors and warp stall samples.

usage: tools/ncu_shade_lines.py gpurun_out/z6_shade_cfg5.source.csv.gz [kernel-substring] [top]"""
import collections
import csv
import gzip
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def I(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def kernels(path):
    op = gzip.open if path.endswith(".gz") else open
    cur = None
    for r in csv.reader(io.TextIOWrapper(op(path, "rb"))):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            yield cur
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["rows"].append(r)


def line_table(mangled_sub):
    """address -> (file, line) of the first function whose mangled name contains `mangled_sub`"""
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "bling_b200" / "libblingcu.so")], cwd=tmp, capture_output=True)
    cub = max(tmp.glob("*.cubin"), key=lambda p: p.stat().st_size)
    out = subprocess.run(["nvdisasm", "-g", str(cub)], capture_output=True, text=True).stdout.split("\n")
    tab, cur, inside = {}, None, False
    for l in out:
        if l.lstrip().startswith(".section"):
            inside = (".text." in l) and (mangled_sub in l)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/", l)
        if m and cur:
            tab[int(m.group(1), 16)] = cur
    return tab


def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else "ShadeHitBodyILi0E"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    tab = line_table(sub)
    for k in list(kernels(path)):
        if not k["rows"]:
            continue
        ix = {n: i for i, n in enumerate(k["hdr"])}
        base = None
        agg = collections.defaultdict(lambda: [0, 0, 0, 0])
        tot = [0, 0, 0, 0]
        for r in k["rows"]:
            a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else I(r[ix["Address"]])
            if base is None:
                base = a
            v = [I(r[ix["Instructions Executed"]]), I(r[ix["L1 Tag Requests Global"]]), I(r[ix["L2 Theoretical Sectors Global"]]), I(r[ix["# Samples"]])]
            key = tab.get(a - base, ("?", 0))
            for j in range(4):
                agg[key][j] += v[j]; tot[j] += v[j]
        print(f"### `{k['name'][:110]}`\n")
        print(f"warp instructions {tot[0]:,}, L1 tag requests (global) {tot[1]:,}, L2 sectors (global) {tot[2]:,}, stall samples {tot[3]:,}\n")
        print("| source line | warp inst % | L1 requests % | L2 sectors % | stall samples % | text |")
        print("|---|---|---|---|---|---|")
        for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            text = ""
            f = ROOT / "bling_b200" / "csrc" / key[0]
            if f.exists() and key[1] > 0:
                text = f.read_text().split("\n")[key[1] - 1].strip()[:90].replace("|", "\\|")
            print(f"| {key[0]}:{key[1]} | {100 * v[0] / max(1, tot[0]):.1f} | {100 * v[1] / max(1, tot[1]):.1f} | {100 * v[2] / max(1, tot[2]):.1f} | {100 * v[3] / max(1, tot[3]):.1f} | `{text}` |")
        print()
        break


if __name__ == "__main__":
    main()
