#!/usr/bin/env python
"""Turns the ncu CSVs a GPU call brought back (gpurun_out/) into the committed summaries under profiles/:

  r02_trace_counters.json   one launch per traversal kernel class from `ncu --set full` (raw page CSV): issue-slot utilisation,
                            threads per instruction, L1 / L2 hit rates, DRAM and pipe utilisations -- bench.py prints them as `ncu`
  r02_traffic.json          DRAM bytes per ray and kernel class over EVERY traversal launch of one step of bench.py
                            (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:kTrace`
                            of `bench.py --steps 1 --warmup 0`, rays of that same step from the run's own JSON line)
  r02_launches.md           the launch list of one step: share of every kernel

usage: tools/make_r02_profiles.py counters gpurun_out/X.raw.csv | traffic gpurun_out/T.csv gpurun_out/T.json | launches gpurun_out/L.csv"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
P = ROOT / "profiles"

COUNTERS = {
    "ms": "gpu__time_duration.sum",
    "registers": "launch__registers_per_thread",
    "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "threads_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1_data_pipe_busy_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "warp_instructions": "smsp__inst_executed.sum",
    "global_load_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
}


def kclass(name):
    if "kTrace" not in name:
        return None
    if "kTraceWarpQ" in name:
        args = name.split("kTraceWarpQ<")[1].split(">")[0].replace("(bool)", "").replace(" ", "").split(",")
        if len(args) > 2 and args[2] in ("1", "true"):
            return None          # the counting instantiation (option traversal_stats): never inside a timed region
        return "trace_any" if args[0] in ("1", "true") else "trace_nearest"
    if "kTracePersistent" in name:
        return "trace_any" if "(bool)1" in name.split("kTracePersistent")[1][:12] else "trace_nearest"
    return None


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def counters(raw_csv):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"source": f"ncu --set full --clock-control none of `bench.py --steps 1 --warmup 1` ({Path(raw_csv).name}); per launch, replayed, cold-ish caches"}
    for r in data:
        c = kclass(r[col["Kernel Name"]])
        if c is None:
            continue
        rec = {k: fnum(r[col[m]]) for k, m in COUNTERS.items() if m in col}
        if units[col[COUNTERS["ms"]]] in ("us", "usecond"): rec["ms"] /= 1e3
        if units[col[COUNTERS["ms"]]] in ("ns", "nsecond"): rec["ms"] /= 1e6
        rec["kernel"] = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        if c not in out or rec["ms"] > out[c]["ms"]:      # keep the longest launch of each class
            out[c] = rec
    (P / "r02_trace_counters.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


def long_rows(path):
    rows = [x for x in csv.reader(open(path)) if len(x) > 5]
    hdr = rows[0]
    i_name, i_met, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    i_id = hdr.index("ID")
    for x in rows[1:]:
        yield x[i_id], x[i_name], x[i_met], fnum(x[i_val]), x[i_unit]


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def to_ms(v, u):
    return v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}.get(u, 1)


def traffic(csv_path, json_path):
    line = json.loads(Path(json_path).read_text().strip().splitlines()[-1])
    steps = line["steps"]
    rays = {"trace_nearest": line["rays"]["nearest_hit_queries"], "trace_any": line["rays"]["any_hit_queries"]}
    agg = collections.defaultdict(lambda: [0.0, 0.0, set()])
    for lid, name, met, v, u in long_rows(csv_path):
        c = kclass(name)
        if c is None or v is None:
            continue
        if met.startswith("dram__bytes"): agg[c][0] += to_bytes(v, u)
        if met.startswith("gpu__time_duration"): agg[c][1] += to_ms(v, u)
        agg[c][2].add(lid)
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kTrace of "
                     f"`bench.py --steps {steps} --warmup 0` ({Path(csv_path).name}): every traversal launch of the timed step(s), rays from that run's own line"}
    for c, (b, ms, ids) in agg.items():
        out[c] = {"dram_bytes": b, "launches": len(ids), "rays": rays[c], "dram_bytes_per_ray": b / max(1.0, rays[c]), "ms_under_ncu": ms}
    (P / "r02_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


def launches(csv_path):
    agg = collections.OrderedDict()
    for lid, name, met, v, u in long_rows(csv_path):
        if not met.startswith("gpu__time_duration") or v is None:
            continue
        n = name.split("(")[0].replace("void ", "")
        c = kclass(name)
        if "kTrace" in name:
            n = n.split("<")[0] + ("<counting>" if c is None else ("<any>" if c == "trace_any" else "<nearest>"))
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += to_ms(v, u)
    tot = sum(a[1] for a in agg.values())
    out = ["# r02 launch list", "",
           f"`ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes` ({Path(csv_path).name}).",
           "cfg 5: 10 M-triangle soup, 3840x2160, 8 sample indices per step. The run holds a warm-up step, the timed step and the counting step",
           "(`<counting>`: the instrumented instantiations of option traversal_stats, never inside a timed region).",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live `kernel_ms_by_class`, not absolutes.", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f} % |")
    (P / "r02_launches.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    {"counters": counters, "traffic": traffic, "launches": launches}[sys.argv[1]](*sys.argv[2:])
