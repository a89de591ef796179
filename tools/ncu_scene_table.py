#!/usr/bin/env python
"""profiles/r01_named_scenes.md from gpurun_out/scene_metrics_<scene>.csv (tools/gpu_ncu_scenes.sh): SURVEY §8(d) asks, for the
configs whose scene lives in L1/L2 (HBM roofline n/a), for SM issue-slot utilisation, L2 hit rate and warp execution efficiency
instead. One `ncu --metrics ...` pass over `tools/scene_breakdown.py <scene>` per scene; per kernel class the launches are
averaged weighted by their duration."""
import collections
import csv
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "gpurun_out"
SCENES = ["cornell-box", "glass-torus", "specular", "ducky", "sun-sky", "environment"]
COLS = [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per instruction (of 32)"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]


def klass(name):
    if "kTraceWarpQ<(bool)0" in name or "kTraceWarpQ<0" in name: return "trace nearest"
    if "kTraceWarpQ<(bool)1" in name or "kTraceWarpQ<1" in name: return "trace any"
    if "kTracePersistent<(bool)0>" in name or "kTracePersistent<0>" in name: return "trace nearest"
    if "kTracePersistent<(bool)1>" in name or "kTracePersistent<1>" in name: return "trace any"
    if "ClassifyBody" in name: return "classify"
    if "ShadeMissBody" in name: return "shade (miss)"
    if "RaygenBody" in name: return "raygen"
    if "ShadeHitBody" in name: return "shade (hit)"
    if "Resolve" in name: return "resolve"
    if "kFilmTile" in name or "FilmBody" in name or "FinalizeBody" in name: return "film"
    return "other"


def main():
    TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
    summary = {}
    out = [f"# {TAG} -- named scenes (BASELINE.json configs[0..3]) at config size: where the time goes and how the SMs are used", "",
           "SURVEY §8(d): these scenes hold 4 - 14 129 primitives and live in L1/L2, so the HBM roofline does not apply; reported instead:",
           "issue-slot utilisation, warp execution efficiency (threads per executed instruction), cache hit rates.", "",
           "`ncu --metrics gpu__time_duration.sum,smsp__issue_active…,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,"
           "gpu__dram_throughput…,l1tex__t_sector_hit_rate.pct,sm__warps_active… --clock-control none -c 600 python tools/scene_breakdown.py <scene>`",
           "(one warm-up slice + one timed slice of up to 48 M camera samples; launches averaged weighted by duration; times under ncu are",
           "serialised, so read the SHARES).", ""]
    for scene in SCENES:
        f = G / f"scene_metrics_{scene}.csv"
        if not f.exists(): continue
        rows = [r for r in csv.reader(open(f)) if len(r) > 10]
        hdr = rows[0]; ik, im, iv, iu, iid = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
        launches = collections.defaultdict(dict)
        for r in rows[1:]:
            v = float(r[iv].replace(",", "")) if r[iv] not in ("", "n/a") else float("nan")
            if r[im] == "gpu__time_duration.sum":
                v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)   # -> ms
            launches[r[iid]][r[im]] = v; launches[r[iid]]["name"] = r[ik]
        agg = collections.OrderedDict()
        for l in launches.values():
            t = l.get("gpu__time_duration.sum", 0.0)
            a = agg.setdefault(klass(l["name"]), {"t": 0.0, "n": 0, **{m: 0.0 for m, _ in COLS}})
            a["t"] += t; a["n"] += 1
            for m, _ in COLS:
                v = l.get(m, float("nan"))
                if v == v: a[m] += v * t
        tot = sum(a["t"] for a in agg.values())
        out += [f"## {scene}", "", "| kernel class | launches | share of GPU time | " + " | ".join(n for _, n in COLS) + " |", "|---|---|---|" + "---|" * len(COLS)]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
            if a["t"] <= 0: continue
            out.append(f"| {k} | {a['n']} | {100 * a['t'] / tot:.1f} % | " + " | ".join(f"{a[m] / a['t']:.1f}" for m, _ in COLS) + " |")
        out.append("")
        whole = {m: sum(a[m] for a in agg.values()) / tot for m, _ in COLS}
        summary[scene] = {"source": f"ncu --metrics pass of tools/scene_breakdown.py {scene} (profiles/{TAG}_named_scenes.md); duration-weighted over every launch of one slice",
                          "issue_slots_busy_pct": whole[COLS[0][0]], "threads_per_instruction": whole[COLS[1][0]], "warp_execution_efficiency_pct": 100 * whole[COLS[1][0]] / 32,
                          "l2_hit_pct": whole[COLS[2][0]], "l1_hit_pct": whole[COLS[3][0]], "dram_pct_of_peak": whole[COLS[4][0]], "occupancy_pct": whole[COLS[5][0]],
                          "share_of_gpu_time": {k: a["t"] / tot for k, a in agg.items() if a["t"] > 0}}
    import json
    (ROOT / "profiles" / f"{TAG}_named_scenes.json").write_text(json.dumps(summary, indent=1) + "\n")
    (ROOT / "profiles" / f"{TAG}_named_scenes.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
