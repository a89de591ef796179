#!/usr/bin/env python
"""Kernel-only A/B bench of the traversal kernels: uploads the cfg-5 soup once, traces a fixed batch of incoherent
(bounce-like) and primary-like rays through the C ABI and reports the per-launch kernel time measured with CUDA
events on the launching stream (option profile_kernels). usage: tools/trace_bench.py [--lib libA.so libB.so ...]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200 import api, ir as IR  # noqa: E402
from bling_b200.host.soup import make_soup  # noqa: E402


def bounce_rays(n, seed, extent=100.0):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, IR.RAY_DTYPE)
    rays["o"] = (rng.random((n, 3)) * 2 - 1).astype(np.float32) * extent
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"] = d.astype(np.float32); rays["tmin"] = 1e-3; rays["tmax"] = np.inf
    return rays


def morton_sorted(rays, extent=100.0, bits=6, octant_major=False):
    """the same rays ordered by the Morton code of their origin's cell (and direction octant): how much the traversal kernels
    would gain from binning the ray queues (profiles/r02_trace_warpq.md)"""
    c = np.clip(((rays["o"] / extent + 1) * 0.5 * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(rays), np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((c[:, a] >> b) & 1) << (3 * b + a)
    octant = ((rays["d"][:, 0] < 0) * 1 + (rays["d"][:, 1] < 0) * 2 + (rays["d"][:, 2] < 0) * 4).astype(np.int64)
    key = (octant << (3 * bits)) | key if octant_major else (key << 3) | octant
    return np.ascontiguousarray(rays[np.argsort(key, kind="stable")])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", nargs="*", default=[str(api.LIB_PATH)])
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--rays", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--option", action="append", default=[])
    ap.add_argument("--sorted", action="store_true", help="also time the bounce batch binned by origin cell / direction octant")
    ap.add_argument("--variants", type=int, nargs="*", default=None, help="trace_variant values to time in ONE context (scene uploaded once)")
    a = ap.parse_args()
    scene = make_soup(a.tris, 64, 36, 1, 1)
    from tests.conftest import camera_rays
    batches = {"bounce": bounce_rays(a.rays, 1), "primary": camera_rays(None, scene, a.rays, 2)}
    if a.sorted:
        batches["bnc-cell"] = morton_sorted(batches["bounce"]); batches["bnc-oct"] = morton_sorted(batches["bounce"], octant_major=True)
        batches["bnc-cell4"] = morton_sorted(batches["bounce"], bits=4)
    for lib in a.lib:
        class Ctx(api.Context):
            _lib_path = Path(lib)
        c = Ctx(0)
        for kv in a.option:
            k, v = kv.split("="); c.set_option(k, float(v))
        c.upload_scene(scene)
        c.set_option("profile_kernels", 1)
        for variant in (a.variants or [None]):
          if variant is not None: c.set_option("trace_variant", variant)
          for name, rays in batches.items():
            c.trace_nearest(rays); c.trace_occluded(rays)          # warm-up
            c.reset_stats()
            for _ in range(a.reps):
                c.trace_nearest(rays)
            kt = c.kernel_times(); ms_n = kt["trace_nearest"][0] / a.reps
            c.reset_stats()
            for _ in range(a.reps):
                c.trace_occluded(rays)
            kt = c.kernel_times(); ms_a = kt["trace_any"][0] / a.reps
            print(f"{Path(lib).name:28s} v{variant} {name:8s} nearest {ms_n:7.3f} ms = {a.rays / ms_n / 1e3:7.1f} Mrays/s   any {ms_a:7.3f} ms = {a.rays / ms_a / 1e3:7.1f} Mrays/s", flush=True)
        c.close()


if __name__ == "__main__":
    main()
