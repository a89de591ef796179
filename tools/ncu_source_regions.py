#!/usr/bin/env python
"""Per-instruction view of a traversal kernel from ncu's source page (`ncu -i X.ncu-rep --page source --csv | gzip`):
where the warp instructions and the warp stall samples go. Prints markdown: the stall mix, runs of SASS lines that execute
equally often (= the blocks of the persistent loop: loop head, refill, node step, enqueue, pop, leaf pass, retire) with their
share of all warp instructions and of all stall samples, and the instructions that collect the most samples.

usage: tools/ncu_source_regions.py gpurun_out/r02_near_n.source.csv.gz [min_samples_pct]"""
import collections
import csv
import gzip
import io
import sys


def I(x):
    try:
        return int(x)
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


def main():
    path = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
    seen = set()
    for k in list(kernels(path)):
        if k["name"] in seen or not k["rows"]:
            continue
        seen.add(k["name"])
        ix = {n: i for i, n in enumerate(k["hdr"])}
        R = k["rows"]
        ex = [I(r[ix["Instructions Executed"]]) for r in R]
        sm = [I(r[ix["# Samples"]]) for r in R]
        thr = [r[ix["Avg. Threads Executed"]] for r in R]
        src = [" ".join(r[ix["Source"]].split()) for r in R]
        tot, st = sum(ex), sum(sm)
        print(f"### `{k['name'].split('(const')[0].replace('void bl::', '')}`\n")
        print(f"{len(R)} SASS lines, {tot / 1e9:.2f} G warp instructions, {st} stall samples; "
              f"shared-memory wavefronts {sum(I(r[ix['L1 Wavefronts Shared']]) for r in R) / 1e6:.0f} M "
              f"({sum(I(r[ix['L1 Wavefronts Shared Excessive']]) for r in R) / 1e6:.0f} M of them bank conflicts)\n")
        mix = collections.Counter()
        for r in R:
            for n in k["hdr"]:
                if n.startswith("stall_") and "Not Issued" not in n:
                    mix[n] += I(r[ix[n]])
        print("stall mix: " + ", ".join(f"{n[6:]} {100 * v / st:.1f} %" for n, v in mix.most_common(7)) + "\n")
        print("| SASS lines | instructions | executed (M warps) | threads | share of warp instructions | share of stall samples |\n|---|---|---|---|---|---|")
        seg = []
        for i in range(len(R)):
            if seg and abs(ex[i] - seg[-1][2]) <= 0.02 * max(ex[i], 1):
                seg[-1][1] = i; seg[-1][3] += ex[i]; seg[-1][4] += sm[i]; seg[-1][5] += 1
            else:
                seg.append([i, i, ex[i], ex[i], sm[i], 1, thr[i]])
        for s in seg:
            if s[3] > 0.004 * tot or s[4] > 0.01 * st:
                print(f"| {s[0]}-{s[1]} | {s[5]} | {s[2] / 1e6:.1f} | {s[6]} | {100 * s[3] / tot:.1f} % | {100 * s[4] / st:.1f} % |")
        print("\n| line | instruction | executed (M) | threads | stall samples |\n|---|---|---|---|---|")
        for i in range(len(R)):
            if sm[i] >= min_pct / 100 * st:
                print(f"| {i} | `{src[i][:64]}` | {ex[i] / 1e6:.1f} | {thr[i]} | {100 * sm[i] / st:.1f} % |")
        mem = [(i, I(R[i][ix["L1 Wavefronts Shared"]])) for i in range(len(R)) if I(R[i][ix["L1 Wavefronts Shared"]]) > 0.02 * sum(I(r[ix["L1 Wavefronts Shared"]]) for r in R)]
        print("\n| line | shared-memory instruction | executed (M) | wavefronts (M) | per instruction |\n|---|---|---|---|---|")
        for i, w in mem:
            print(f"| {i} | `{src[i][:64]}` | {ex[i] / 1e6:.1f} | {w / 1e6:.1f} | {w / max(1, ex[i]):.2f} |")
        print()


if __name__ == "__main__":
    main()
