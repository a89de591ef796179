#!/usr/bin/env python
"""Writes profiles/r02_trace_source_view.md: the per-instruction view of the traversal kernels (ncu source pages kept under
profiles/*.source.csv.gz, tables by tools/ncu_source_regions.py), the ncu comparison of the two any-hit child orders
(profiles/r02_any_slot.raw.csv.gz / r02_any_longest.raw.csv.gz) and the A/B log of the second half of round 2.
usage: tools/make_r02_source_view.py"""
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
P = ROOT / "profiles"


def regions(name):
    return subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_source_regions.py"), str(P / name), "0.8"], check=True, capture_output=True, text=True).stdout


ORDER_METRICS = [
    ("kernel time (ms)", "gpu__time_duration.sum", 1.0, "{:.2f}"),
    ("warp instructions (G)", "smsp__inst_executed.sum", 1e-9, "{:.2f}"),
    ("threads per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio", 1.0, "{:.2f}"),
    ("global-load sectors through L1 (G)", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 1e-9, "{:.2f}"),
    ("L1 data pipe busy %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct", 1.0, "{:.1f}"),
    ("DRAM read (GB)", "dram__bytes_read.sum", 1.0, "{:.1f}"),
    ("warps stalled on the long scoreboard per issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0, "{:.2f}"),
]


def order_table():
    def load(f):
        import gzip, io
        rows = list(csv.reader(io.TextIOWrapper(gzip.open(P / (f + ".gz"), "rb"))))
        return [dict(zip(rows[0], r)) for r in rows[2:]]
    a, b = load("r02_any_slot.raw.csv"), load("r02_any_longest.raw.csv")
    out = ["| metric | launch 1: slot order | launch 1: longest first | launch 2: slot order | launch 2: longest first |", "|---|---|---|---|---|"]
    for label, m, sc, fmt in ORDER_METRICS:
        cells = []
        for i in range(2):
            for t in (a, b):
                cells.append(fmt.format(float(t[i][m].replace(",", "")) * sc))
        out.append(f"| {label} (`{m}`) | " + " | ".join(cells) + " |")
    return "\n".join(out)


TEXT = """# r02 — the traversal kernels instruction by instruction (ncu source page), and what the second half of round 2 tried

`ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:kTraceWarpQ<\\(bool\\)0' -s 2 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scenes`
(cfg 5; the nearest-hit launch of the extension rays of bounce 1, ~35 M rays; `(bool)1` for the any-hit launches), exported on
the box with `ncu -i X.ncu-rep --page source --csv | gzip` and kept here as `r02_*.source.csv.gz`; the tables are
`tools/ncu_source_regions.py <file> 0.8`. `tools/gpu_r02_k.sh`, `_n.sh`, `gpu_r02_final.sh`.

The question this view was opened for: the kernels report 0.57-0.63 of the L1 data-pipe roofline with that pipe 73-80 % busy
(`r02_trace_warpq.md`) — so why did 12 % fewer node visits (the any-hit child order below) buy 1-2 % of time?

## 1. Where a trip of the persistent loop goes (nearest-hit kernel, BEFORE the changes of this section)

One trip = every lane that stands at an inner node does one node step; lanes at a leaf queue their items; everybody pops;
when 32 pairs are queued all lanes test one. The runs of SASS lines that execute equally often are the blocks of that loop:
lines 35-47 loop head + refill test (every trip, 62 M warp-trips per launch), 48-169 the refill (10 M), 177-367 the node step
(62 M, 29 of 32 threads), 369-411 the queue append, 413-447 the pop and the leaf-pass test, 448-1308 the leaf pass (17 M),
1309-1329 retire.

@BEFORE@

Reading: the node step is 48 % of all warp instructions (156 per trip), queue append + pop 22 % (71), the leaf pass 12 %, the
refill 4 %; half of all warp stall samples are long-scoreboard waits, and they sit on FIVE instructions: the first use of the
node (line 192, 18 %), of the leaf item (486 + 491, 10.5 %), the refill chain atomic -> q[k] -> ray (77, 100, 107: 8 %) — and
three consumers of **local-memory reloads**: `IMAD R0, R7, 0x200, R42` (329, 6.1 %) and `LEA R0, R47, R2, 0x9` (435, 3.7 %) wait for
`LDL [R1+0x2c8]`, the stack base that ptxas had spilled and re-read twice per trip, and `IMAD.WIDE.U32 R2, R5, 0x10, R2` (1320,
4.6 %) waits for the spilled slot of the ray at retire. With an L1 hit rate of 3-10 % under the scattered node traffic those
reloads are L2 round trips: **14 % of all warp samples waited on 12 bytes of spill**.

## 2. The stack in the warp's own shared-memory region (built: the product)

`trace_warpq.cuh`: the traversal stack moved from `[level][thread]` behind the four warps' queue heads into each warp's region
(`[level][lane]`, 128 B per level), and the stack pointer became the shared-window ADDRESS of the next free entry. Every push
and pop is one `STS` / `LDS` with an immediate offset (`STS [R47+-0x80]`), empty / overflow tests are one subtraction against
the warp's base, and neither a stack base nor a level count lives in a register; the slot of the ray went to a per-lane
shared-memory word (written at refill, read at retire). ptxas: nearest-hit 64 registers, 12 -> 0 bytes of spill; any-hit 56
registers, 12 -> 4 bytes (the ray count, read in the refill only). Results bit-identical (`test_traversal_variants_agree_bit_for_bit`).

@AFTER@

Live in `bench.py` (cfg 5, ms per 5 steps): trace_nearest 635 -> **589** (-7.3 %), trace_any 633 -> 626; 215.2 -> 222.3 Msamples/s
(`tools/gpu_r02_m.sh`). The table above is the final capture (refill threshold 6, section 4): the long-scoreboard share fell from 51
to 40 % of the samples, issue slots went from 61 to 67 % busy (`r02_trace_counters.json`); what is left is the node fetch (22.4 %),
the item fetch (11.8 %) and the refill chain (6.5 %).

The any-hit kernel after the same change (final capture; the node fetch is 29 % of its samples, the item fetch 14 %: its rays
have the worse L2 hit rate, section 3):

@ANY@

## 3. Any-hit child order: enter the child the ray stays in longest (built: the product)

`tools/travsim.cpp` (section "Any-hit child order, second look" of `r02_travsim.md`): distance order, slot order and static
orders all visit the same number of nodes; entering the child with the largest `tfar - tnear` first visits 15 % fewer. Built as
three compares on `node4Near<OVERLAP>`'s keys (+8 instructions per node step, 56 registers kept). The kernel's own counters over
a cfg-5 step: **40.39 -> 35.53 node visits, 13.54 -> 12.93 primitive tests per any-hit ray**; algorithmic bytes per ray 3485 -> 3134.
ncu on the same two launches with both builds (`r02_any_slot.raw.csv.gz`, `r02_any_longest.raw.csv.gz`, `tools/gpu_r02_k.sh`):

@ORDER@

4.6 % fewer warp instructions, 10 % fewer sectors through L1, the L1 data pipe 78 -> 71 % busy — and 0.4-1.8 % less time under
ncu, 2.2 % live (trace_any 649 -> 635 ms per 5 steps). The reason is in the L2 rows: the nodes are laid out depth first, so a walk
in slot order runs forward through memory and neighbouring rays share lines; a walk that picks its child by the ray does not:
L2 hit rate 57-59 -> 44-47 %, DRAM reads +33-40 %, more warps waiting on the long scoreboard. So the kernel is NOT bound by the
L1 pipe alone: it sits on a balance of issue slots (61-63 % busy), the L1 pipe (71-80 %) and memory latency that 32-36 resident
warps per SM do not hide, and taking load off one of them moves the time by a fraction of that load. The `roofline.frac` that
`bench.py` prints went DOWN with this change (0.62 -> 0.57: fewer algorithmic bytes in almost the same time), which is what an
honest algorithmic-bytes roofline does when work is removed without a matching gain.

## 4. What did not pay (all measured on the B200, cfg 5, `bench.py --steps 5 --warmup 3`, ms per 5 steps nearest / any)

| experiment | how | result |
|---|---|---|
| L2 prefetch of leaf items when they are queued | `TQ_PREFETCH=1`, `tools/gpu_r02_l.sh` | 635 / 633 -> 636 / 645 |
| L2 prefetch of inner children when they are pushed (every pushed child IS visited: no cull on pop) | `TQ_PREFETCH=2` | -> 650 / 652 |
| both | `TQ_PREFETCH=3` | -> 653 / 683; cornell-box 437 -> 321 Msamples/s |
| L1 prefetch of the next trip's node right after the pop (~100 instructions ahead of its use) | `TQ_PREFETCH=4`, `tools/gpu_r02_p.sh` | 588 / 626 -> 636 / 668 |
| the same, both sectors | `TQ_PREFETCH=12` | -> 875 / 917 |
| rays reserved 32 at a time, one chunk ahead, `q[base + lane]` loaded once per chunk (refill = one shuffle + the ray load) | not kept; `tools/gpu_r02_m.sh` second run | 589 / 626 -> 602 / 680 (the extra warp-uniform state brought spill reloads back into the node step of the 56-register kernel) |
| L2 eviction policies folded into the load descriptors: leaf items `evict_first` (640 MB, hardly reused), nodes `evict_last`, both | `BL_L2_POLICY=1/2/3`, `tools/gpu_r02_t.sh` | 584 / 621 -> 582 / 626, 585 / 620, 584 / 626: the L2 already keeps what is reused |
| refill once 6 / 8 / 12 lanes are idle instead of 4 | `TR_REFILL`, `tools/gpu_r02_q.sh` | 589 / 628 -> **585 / 622** / 589 / 624 / 610 / 631: 6 is the product |
| binning rays by origin cell and direction octant (upper bound: 4 M uniformly random rays, sorted on the host by Morton code) | `tools/trace_bench.py --sorted`, `tools/gpu_r02_r.sh` | nearest 781 -> 847 Mrays/s (+8 %), any 926 -> 964 (+4 %) for rays that start with NO order at all; the pipeline's queues are already pixel-ordered, and a device sort of 20-35 M keys per launch costs more than that |

`CCTL.E.PF1/PF2` (what `prefetch.global.L1/L2` compiles to) is expensive on this part: one per lane and trip costs 7-8 % of the
kernel, two 49 %. Software prefetch is not a tool here.

## 5. What bounds the kernels now

Per trip the nearest-hit kernel issues ~330 warp instructions (node step 156, append + pop 70, leaf pass 50 amortised, loop
head, refill test and retire 50) and moves 2 x 29 node sectors + 0.28 x 2 x 31 item sectors + ~15 shared-memory wavefronts
through the L1 data pipe; with 8 warps per scheduler both are ~2/3 busy and the remaining third is dependent-load latency
(node fetch, item fetch, refill chain) that more resident warps would hide — and 64 registers x 128 threads x 8 CTAs is the
register file. Fewer instructions per node step (it is 24 `PRMT` + 24 `FFMA` + 22 `FMNMX` + sort + pushes: little slack), fewer
visits WITHOUT losing the memory order (a better tree, not a per-ray order), or a node format whose fetch is one sector are the
three ways left; none is a tuning step.
"""


def main():
    t = TEXT.replace("@BEFORE@", regions("r02_near_before.source.csv.gz")).replace("@AFTER@", regions("r02_near_after.source.csv.gz"))
    t = t.replace("@ANY@", regions("r02_any_after.source.csv.gz")).replace("@ORDER@", order_table())
    (P / "r02_trace_source_view.md").write_text(t)
    print(t[:600])


if __name__ == "__main__":
    main()
