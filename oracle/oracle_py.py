"""ctypes binding of the CPU oracle (liboracle.so). TEST INFRASTRUCTURE ONLY: may be imported from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from bling_b200/."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from bling_b200 import ir as IR

_DIR = Path(__file__).resolve().parent
_LIB = None


def build(force=False):
    so = _DIR / "liboracle.so"
    srcs = [p for p in _DIR.iterdir() if p.suffix in (".cpp", ".h")] + [_DIR.parent / "include" / "blingcu.h"]
    if force or not so.exists() or any(p.stat().st_mtime > so.stat().st_mtime for p in srcs):
        subprocess.check_call(["make", "-s", "-C", str(_DIR)])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        P = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(IR.SceneC), C.c_int, C.POINTER(P)]
        L.oracle_destroy.argtypes = [P]; L.oracle_destroy.restype = None
        L.oracle_trace_nearest.argtypes = [P, P, C.c_size_t, P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_trace_occluded.argtypes = [P, P, C.c_size_t, P, C.c_int]
        L.oracle_debug_eval.argtypes = [C.c_int, P, P]
        L.oracle_light_trace.argtypes = [P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, P, C.c_size_t, C.POINTER(C.c_size_t)]
        L.oracle_read_splat.argtypes = [P, P]
        L.oracle_export_kdtree.argtypes = [P, P, C.POINTER(C.c_uint32), P, C.POINTER(C.c_size_t), C.POINTER(C.c_int32), P]
        L.oracle_trace_kd_stats.argtypes = [P, P, C.c_size_t, P, P, P]
        L.oracle_sample_extent.argtypes = [P] + [C.POINTER(C.c_int32)] * 4
        L.oracle_render_samples.argtypes = [P, C.c_uint32, C.c_uint64, P, P, P, C.c_size_t, P, P]
        L.oracle_render_slice.argtypes = [P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.oracle_read_film.argtypes = [P, P]
        L.oracle_clear_film.argtypes = [P]
        L.oracle_get_stats.argtypes = [P, C.POINTER(IR.Stats)]
        L.oracle_reset_stats.argtypes = [P]
        L.oracle_eval_texture.argtypes = [P, C.c_int32, P, P, C.c_size_t, P]
        L.oracle_add_sample_tile.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, P, P] + [C.POINTER(C.c_int)] * 4
        _LIB = L
    return _LIB


def debug_eval(what: int, args, n_out: int) -> np.ndarray:
    """oracle_debug_eval: one function of oracle_shade.h at explicit arguments (tests/test_third_statement.py)"""
    a = np.ascontiguousarray(np.concatenate([np.atleast_1d(np.asarray(x, np.float32)).ravel() for x in args]), np.float32)
    out = np.zeros(n_out, np.float32)
    rc = lib().oracle_debug_eval(what, a.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_debug_eval({what}) failed with code {rc}")
    return out


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed with code {rc}")


class Oracle:
    """CPU restatement of bling's path integrator on a flat scene IR."""

    def __init__(self, scene: IR.SceneIR, kdtree=True):
        self.scene = scene
        sc, keep = scene.to_c()
        self._h = C.c_void_p()
        _chk(lib().oracle_create(C.byref(sc), int(kdtree), C.byref(self._h)), "create")
        self.kdtree = kdtree

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h); self._h = None

    def __del__(self):
        try: self.close()
        except Exception: pass

    def trace_nearest(self, rays: np.ndarray, mode="brute"):
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE); out = np.zeros(len(rays), IR.HIT_DTYPE)
        nt, ni = C.c_uint64(), C.c_uint64()
        _chk(lib().oracle_trace_nearest(self._h, rays.ctypes.data, len(rays), out.ctypes.data, 1 if mode == "kd" else 0,
                                        C.byref(nt), C.byref(ni)), "trace_nearest")
        self.last_traversal = (nt.value, ni.value)
        return out

    def light_trace(self, pass_index, seed, first_photon, n_photons, records=False):
        """Renderer/LightTracer.hs restated: photons [first, first + n) into the splat buffer; records=True also returns the rows
        {photon, depth, px, py, X, Y, Z} in (photon, depth) order"""
        cap = 64 * n_photons + 16 if records else 0
        out = np.zeros((cap, 7), np.float32); n = C.c_size_t()
        _chk(lib().oracle_light_trace(self._h, pass_index, seed, first_photon, n_photons, out.ctypes.data if records else None, cap, C.byref(n)), "light_trace")
        return out[:min(cap, n.value)] if records else None

    def read_splat(self):
        out = np.zeros((self.scene.height, self.scene.width, 3), np.float32)
        _chk(lib().oracle_read_splat(self._h, out.ctypes.data), "read_splat")
        return out

    def kdtree_flat(self):
        """(nodes, leaf_prims, root, bounds): the SAH kd-tree of KdTree.hs as the flat arrays blingcu_upload_kdtree takes"""
        nn, nl, root = C.c_uint32(), C.c_size_t(), C.c_int32(); b = (C.c_float * 6)()
        _chk(lib().oracle_export_kdtree(self._h, None, C.byref(nn), None, C.byref(nl), C.byref(root), b), "export_kdtree")
        nodes = np.zeros(nn.value, IR.KDNODE_DTYPE); leaf = np.zeros(max(1, nl.value), np.uint32)
        _chk(lib().oracle_export_kdtree(self._h, nodes.ctypes.data, C.byref(nn), leaf.ctypes.data, C.byref(nl), C.byref(root), b), "export_kdtree")
        return nodes, leaf[:nl.value], root.value, np.array(list(b), np.float32)

    def trace_kd_stats(self, rays: np.ndarray):
        """kd-tree traversal with the per-ray counters of dbgTraverse (KdTree.hs:260-281)"""
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE); out = np.zeros(len(rays), IR.HIT_DTYPE)
        nt = np.zeros(len(rays), np.uint32); ni = np.zeros(len(rays), np.uint32)
        _chk(lib().oracle_trace_kd_stats(self._h, rays.ctypes.data, len(rays), out.ctypes.data, nt.ctypes.data, ni.ctypes.data), "trace_kd_stats")
        return out, nt, ni

    def trace_occluded(self, rays: np.ndarray, mode="brute"):
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE); out = np.zeros(len(rays), np.uint8)
        _chk(lib().oracle_trace_occluded(self._h, rays.ctypes.data, len(rays), out.ctypes.data, 1 if mode == "kd" else 0), "trace_occluded")
        return out

    def sample_extent(self):
        v = [C.c_int32() for _ in range(4)]
        lib().oracle_sample_extent(self._h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def render_samples(self, pass_index, seed, px, py, sample):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32)
        sample = np.ascontiguousarray(sample, np.uint32); n = len(px)
        L = np.zeros((n, 16), np.float32); xy = np.zeros((n, 2), np.float32)
        _chk(lib().oracle_render_samples(self._h, pass_index, seed, px.ctypes.data, py.ctypes.data, sample.ctypes.data, n,
                                         L.ctypes.data, xy.ctypes.data), "render_samples")
        return L, xy

    def eval_texture(self, texture, p, uv):
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3); uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((len(p), 16), np.float32)
        _chk(lib().oracle_eval_texture(self._h, texture, p.ctypes.data, uv.ctypes.data, len(p), out.ctypes.data), "eval_texture")
        return out

    def render_slice(self, pass_index, seed, s_begin, s_end, threads=1):
        _chk(lib().oracle_render_slice(self._h, pass_index, seed, s_begin, s_end, threads), "render_slice")

    def render_pass(self, pass_index, seed, threads=1):
        self.render_slice(pass_index, seed, 0, self.scene.spp, threads)

    def read_film(self):
        f = np.zeros((self.scene.height, self.scene.width, 4), np.float32)
        lib().oracle_read_film(self._h, f.ctypes.data)
        return f

    def clear_film(self): lib().oracle_clear_film(self._h)

    def stats(self):
        s = IR.Stats(); lib().oracle_get_stats(self._h, C.byref(s)); return s.as_dict()

    def reset_stats(self): lib().oracle_reset_stats(self._h)

    def add_sample_tile(self, wnd, sx, sy, L16):
        L16 = np.ascontiguousarray(L16, np.float32)
        v = [C.c_int() for _ in range(4)]
        lib().oracle_add_sample_tile(self._h, *wnd, sx, sy, L16.ctypes.data, None, *[C.byref(x) for x in v])
        ox, oy, w, h = (x.value for x in v)
        out = np.zeros((max(h, 0), max(w, 0), 4), np.float32)
        lib().oracle_add_sample_tile(self._h, *wnd, sx, sy, L16.ctypes.data, out.ctypes.data, *[C.byref(x) for x in v])
        return out, (ox, oy)
