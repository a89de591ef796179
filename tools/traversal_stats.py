#!/usr/bin/env python
"""TraversalStats of the reference's accelerator vs this library's (KdTree.hs:252-281: nodesTraversed, intersections).
The oracle restates the SAH kd-tree of KdTree.hs:107-203 and counts like dbgTraverse; the library's quantised 4-wide BVH
is walked by the kernel-body emulator (same traversal code and order as the CUDA kernels). CPU only.
usage: python tools/traversal_stats.py > profiles/r01_traversal_stats.md"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200 import ir as IR  # noqa: E402
from bling_b200.host.soup import make_soup  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from tests.conftest import camera_rays, compare_hits, load_scene, random_rays  # noqa: E402
from tests.emu.emu_py import EmuContext  # noqa: E402

N = 20000
print("# r01 — traversal statistics: the reference's SAH kd-tree vs the library's quantised BVH4\n")
print("Per-ray means over %d rays per batch (`random`: origins in/around the scene bounds, random directions, 30 %% finite ranges; "
      "`camera`: from the camera into the view cone). kd-tree = `oracle/oracle_kdtree.cpp` (restatement of `KdTree.hs:107-246`, "
      "counters as `dbgTraverse` :260-281); BVH = `bvh_build.cpp` + `bvh.h::traceNearest<true>` (what `blingcu_trace_stats` "
      "returns). Node record: kd-tree node (boxed Haskell value) vs 64-byte quantised 4-wide node; one BVH node visit tests four boxes.\n" % N)
print("| scene | prims | batch | kd nodes | kd prim tests | BVH4 nodes | BVH4 prim tests | same hits |")
print("|---|---|---|---|---|---|---|---|")
for name in ("cornell-box", "ducky", "soup-1M"):
    sc = make_soup(1_000_000, 64, 36, 1, 1) if name == "soup-1M" else load_scene(name)
    o = Oracle(sc, kdtree=True); e = EmuContext(); e.upload_scene(sc)
    for bname, rays in (("random", random_rays(sc, N, 3)), ("camera", camera_rays(None, sc, N, 4))):
        ref = o.trace_nearest(rays, "kd"); kn, kp = o.last_traversal
        got, nodes, prims = e.trace_stats(rays)
        ties, bad = compare_hits(got, ref)
        print(f"| {name} | {sc.n_prims} | {bname} | {kn / N:.1f} | {kp / N:.1f} | {nodes.mean():.1f} | {prims.mean():.1f} | {N - ties - bad}/{N} ({ties} t-ties) |")
