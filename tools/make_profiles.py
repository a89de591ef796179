#!/usr/bin/env python
"""Rebuilds profiles/r01_{launches,trace_final,shade}.md and profiles/traffic.json from the captures in gpurun_out/
(launches_r01.csv, prof_trace_r01.ncu-rep, prof_shade_r01.ncu-rep, bench_default.json). Run here, after a gpurun call."""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "gpurun_out"
P = ROOT / "profiles"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def summary(rep):
    return subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout


bench = json.loads((G / "bench_default.json").read_text().strip().splitlines()[-1])
r = bench["roofline"]

rows = [x for x in csv.reader(open(G / f"launches_{TAG}.csv")) if len(x) > 5]
hdr = rows[0]; i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for x in rows[1:]:
    n = x[i_name].split("(")[0].replace("void ", "")
    if "kTracePersistent" in x[i_name]:
        n = "kTracePersistent<any>" if "(bool)1" in x[i_name] else "kTracePersistent<nearest>"
    v = float(x[i_val].replace(",", "")); u = x[i_unit]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = [f"# {TAG} launch list (final kernels of the round)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes`",
       "(cfg 5: 10 M-triangle soup, 3840x2160, 8 sample indices per step = 66.4 M camera samples per step; the run holds a warm-up step, the timed step and the instrumented counting step, so `kTraceNearestCount` appears here but never inside a timed region).",
       "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live `kernel_ms_by_class`, not absolutes.", "",
       "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {n} | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% |")
tl = sum(r["kernel_ms_by_class"].values())
out += ["", f"Live shares from the same workload NOT under ncu (bench.py default run: {bench['value']:.1f} Msamples/s, {bench['mrays_per_s']:.0f} Mrays/s; CUDA events around every launch, {bench['steps']} timed steps):", "",
        "| class | ms | share |", "|---|---|---|"]
for k, v in sorted(r["kernel_ms_by_class"].items(), key=lambda kv: -kv[1]):
    out.append(f"| {k} | {v:.1f} | {100 * v / tl:.1f}% |")
(P / f"{TAG}_launches.md").write_text("\n".join(out) + "\n")
(P / f"{TAG}_launches.csv").write_text((G / f"launches_{TAG}.csv").read_text())

raw = subprocess.run(["ncu", "-i", str(G / f"prof_trace_{TAG}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u, dat = rr[0], rr[1], rr[2:]
col = h.index
tob = lambda v, unit: float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
near = [x for x in dat if "kTracePersistent<0>" in x[col("Kernel Name")] or "(bool)0" in x[col("Kernel Name")]]
byt = [tob(x[col("dram__bytes_read.sum")], u[col("dram__bytes_read.sum")]) + tob(x[col("dram__bytes_write.sum")], u[col("dram__bytes_write.sum")]) for x in near]
tms = [float(x[col("gpu__time_duration.sum")]) for x in near]
json.dump({"kernel": "kTracePersistent<nearest>",
           "source": f"profiles/{TAG}_trace_final.md (ncu --set full of `bench.py --steps 1 --warmup 1`; the nearest-hit launches of the capture are early bounces, i.e. more rays per launch than the step average)",
           "dram_bytes_per_launch": sum(byt) / len(byt), "launches_sampled": len(byt), "ncu_ms_per_launch": sum(tms) / len(tms)},
          open(P / "traffic.json", "w"), indent=1)

(P / f"{TAG}_trace_final.md").write_text(
    f"# {TAG} final -- kTracePersistent (quantised 64-byte BVH4 nodes, majority-vote stepping, PTX-predicated shared-memory stack)\n\n"
    "Command: `ncu --set full --clock-control none --import-source on -k regex:kTracePersistent -s 3 -c 4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes`; <0> = nearest-hit, <1> = any-hit.\n"
    f"bench.py (not under ncu): {bench['value']:.1f} Msamples/s, {bench['mrays_per_s']:.0f} Mrays/s whole pipeline; trace_nearest {r['mrays_per_s_in_kernel']:.0f} Mrays/s in kernel, "
    f"{r['nodes_per_ray']:.1f} node visits + {r['prims_per_ray']:.1f} triangle tests per ray = {r['bytes_per_ray']:.0f} algorithmic bytes per ray, achieved {r['achieved']:.0f} GB/s = {r['frac']:.2f} of the measured HBM copy bandwidth.\n\n"
    "Reading: the kernel is NOT HBM-bound (DRAM 9-17 %, L2 hit 62-75 %): the busiest units are the ALU pipe (min/max, PRMT, SEL, compares: half rate) and the issue slots, with ~22 of 32 lanes active per instruction, then the LSU data pipe (one L1 wavefront per lane per 16 bytes).\n\n"
    + summary(G / f"prof_trace_{TAG}.ncu-rep"))
(P / f"{TAG}_shade.md").write_text(
    f"# {TAG} -- kRunQueueHeavy<ShadeHitBody<MATTE>> (material-sorted shade kernel, matte queue of cfg 5)\n\n"
    "Same command with `-k regex:kRunQueueHeavy -s 2 -c 2`. 168 registers/thread (16-band spectra: NEE, MIS and extension parts need ~120-128 each, 168 together; the geometry/BSDF set-up alone 72) cap occupancy at 18 %; the kernel is latency/issue-bound, not bandwidth-bound. "
    "Forcing fewer registers spills and is slower (tools/pipeline_ab.py: 56 -> 63/71/78 ms per pass at 4/5/6 CTAs per SM). What helped so far: one instantiation per material kind (glass/mirror/blackbody 127/120/90 registers), per-scene sampler constants with power-of-two fast paths (-11 %).\n\n"
    + summary(G / f"prof_shade_{TAG}.ncu-rep"))
print("profiles written:", bench["value"], bench["mrays_per_s"])
