"""ctypes binding of the C ABI in include/blingcu.h (libblingcu.so).

There is NO CPU fallback: importing is cheap, but `Context()` raises if the CUDA extension is missing or no
GPU is visible (the library returns BLINGCU_ENOGPU).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import ir as IR

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libblingcu.so"

# every symbol include/blingcu.h declares
SYMBOLS = ["create", "destroy", "last_error", "upload_scene", "trace_nearest", "trace_occluded", "trace_stats",
           "render_pass", "render_slice", "render_samples", "eval_texture", "read_film", "clear_film", "film_add_host", "film_device",
           "synchronize", "set_stream", "get_stats", "reset_stats", "set_option", "sample_extent", "kernel_times",
           "comm_unique_id", "comm_init", "comm_init_all", "comm_destroy", "reduce_film", "reduce_film_group", "comm_wait",
           "read_film_sum", "film_sum_device", "host_alloc", "host_free", "upload_kdtree", "trace_kdtree",
           "light_trace", "light_trace_records", "read_splat"]
COMM_ID_BYTES = 128


class BlingCuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"blingcu error {code}: {msg}")
        self.code = code


def load_library(path=LIB_PATH, prefix="blingcu"):
    path = Path(path)
    if not path.exists():
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(there is no CPU fallback)")
    L = C.CDLL(str(path))
    P = C.c_void_p
    f = lambda n: getattr(L, f"{prefix}_{n}")
    f("create").argtypes = [C.c_int, C.POINTER(P)]
    f("destroy").argtypes = [P]; f("destroy").restype = None
    f("last_error").argtypes = [P]; f("last_error").restype = C.c_char_p
    f("upload_scene").argtypes = [P, C.POINTER(IR.SceneC)]
    f("trace_nearest").argtypes = [P, P, C.c_size_t, P]
    f("trace_occluded").argtypes = [P, P, C.c_size_t, P]
    f("trace_stats").argtypes = [P, P, C.c_size_t, P, P, P]
    f("render_pass").argtypes = [P, C.c_uint32, C.c_uint64]
    f("render_slice").argtypes = [P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
    f("render_samples").argtypes = [P, C.c_uint32, C.c_uint64, P, P, P, C.c_size_t, P, P]
    f("eval_texture").argtypes = [P, C.c_int32, P, P, C.c_size_t, P]
    f("read_film").argtypes = [P, P]
    f("clear_film").argtypes = [P]
    f("film_add_host").argtypes = [P, P]
    f("film_device").argtypes = [P, C.POINTER(P), C.POINTER(C.c_size_t)]
    f("synchronize").argtypes = [P]
    f("set_stream").argtypes = [P, P]
    f("get_stats").argtypes = [P, C.POINTER(IR.Stats)]
    f("reset_stats").argtypes = [P]
    f("set_option").argtypes = [P, C.c_char_p, C.c_double]
    f("sample_extent").argtypes = [P] + [C.POINTER(C.c_int32)] * 4
    f("kernel_times").argtypes = [P, P, P, C.c_int]
    f("comm_unique_id").argtypes = [P]
    f("comm_init").argtypes = [P, C.c_int, C.c_int, P]
    f("comm_init_all").argtypes = [C.POINTER(P), C.c_int]
    f("comm_destroy").argtypes = [P]
    f("reduce_film").argtypes = [P, C.c_int]
    f("reduce_film_group").argtypes = [C.POINTER(P), C.c_int, C.c_int]
    f("comm_wait").argtypes = [P]
    f("read_film_sum").argtypes = [P, P]
    f("film_sum_device").argtypes = [P, C.POINTER(P), C.POINTER(C.c_size_t)]
    f("host_alloc").argtypes = [P, C.c_size_t, C.POINTER(P)]
    f("host_free").argtypes = [P, P]
    f("upload_kdtree").argtypes = [P, P, C.c_uint32, C.c_int32, P, C.c_size_t, P]
    f("trace_kdtree").argtypes = [P, P, C.c_size_t, P, P, P]
    f("light_trace").argtypes = [P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32]
    f("light_trace_records").argtypes = [P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, P, C.c_size_t, C.POINTER(C.c_size_t)]
    f("read_splat").argtypes = [P, P]
    return L


class Context:
    """One GPU context (blingcu_ctx). Mirrors the C ABI one-to-one."""

    _lib_path = LIB_PATH
    _prefix = "blingcu"

    def __init__(self, device: int = 0):
        self._L = load_library(self._lib_path, self._prefix)
        self._h = C.c_void_p()
        rc = self._f("create")(device, C.byref(self._h))
        if rc != 0:
            msg = self._f("last_error")(None)
            self._h = None
            raise BlingCuError(rc, (msg or b"").decode())
        self.scene = None

    def _f(self, n):
        return getattr(self._L, f"{self._prefix}_{n}")

    def _chk(self, rc):
        if rc != 0:
            raise BlingCuError(rc, (self._f("last_error")(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_pinned", []):
                self._f("host_free")(self._h, C.c_void_p(p))
            self._pinned = []
            self._f("destroy")(self._h); self._h = None

    def __del__(self):
        try: self.close()
        except Exception: pass

    def __enter__(self): return self
    def __exit__(self, *a): self.close()

    # ---- scene
    def upload_scene(self, scene: IR.SceneIR):
        sc, keep = scene.to_c()
        self._chk(self._f("upload_scene")(self._h, C.byref(sc)))
        self.scene = scene

    def set_option(self, key: str, value: float):
        self._chk(self._f("set_option")(self._h, key.encode(), float(value)))

    def sample_extent(self):
        v = [C.c_int32() for _ in range(4)]
        self._chk(self._f("sample_extent")(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # ---- explicit ray batches
    def host_array(self, n: int, dtype) -> np.ndarray:
        """an n-element array in page-locked host memory (blingcu_host_alloc): ray / hit buffers in it are transferred in
        place by trace_nearest / trace_occluded. Freed with the context (or host_free)."""
        dt = np.dtype(dtype); p = C.c_void_p()
        self._chk(self._f("host_alloc")(self._h, max(1, n * dt.itemsize), C.byref(p)))
        buf = (C.c_uint8 * max(1, n * dt.itemsize)).from_address(p.value)
        self._pinned = getattr(self, "_pinned", []); self._pinned.append(p.value)
        return np.frombuffer(buf, dtype=dt, count=n)

    def trace_nearest(self, rays: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE)
        if out is None: out = np.zeros(len(rays), IR.HIT_DTYPE)
        self._chk(self._f("trace_nearest")(self._h, rays.ctypes.data, len(rays), out.ctypes.data))
        return out

    def trace_occluded(self, rays: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE)
        if out is None: out = np.zeros(len(rays), np.uint8)
        self._chk(self._f("trace_occluded")(self._h, rays.ctypes.data, len(rays), out.ctypes.data))
        return out

    def upload_kdtree(self, nodes: np.ndarray, leaf_prims: np.ndarray, root: int, bounds):
        """the HOST's own kd-tree (KdTree.hs:29-33, flattened; IR.KDNODE_DTYPE) as an alternative accelerator input"""
        nodes = np.ascontiguousarray(nodes, IR.KDNODE_DTYPE); leaf = np.ascontiguousarray(leaf_prims, np.uint32)
        b = np.ascontiguousarray(bounds, np.float32)
        self._chk(self._f("upload_kdtree")(self._h, nodes.ctypes.data, len(nodes), int(root), leaf.ctypes.data if len(leaf) else None, len(leaf), b.ctypes.data))

    def trace_kdtree(self, rays: np.ndarray):
        """`traverse` of KdTree.hs:223-242 over the uploaded kd-tree: hits + per-ray (nodesTraversed, intersections) of dbgTraverse"""
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE); out = np.zeros(len(rays), IR.HIT_DTYPE)
        nodes = np.zeros(len(rays), np.uint32); prims = np.zeros(len(rays), np.uint32)
        self._chk(self._f("trace_kdtree")(self._h, rays.ctypes.data, len(rays), out.ctypes.data, nodes.ctypes.data, prims.ctypes.data))
        return out, nodes, prims

    def trace_stats(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, IR.RAY_DTYPE); out = np.zeros(len(rays), IR.HIT_DTYPE)
        nodes = np.zeros(len(rays), np.uint32); prims = np.zeros(len(rays), np.uint32)
        self._chk(self._f("trace_stats")(self._h, rays.ctypes.data, len(rays), out.ctypes.data, nodes.ctypes.data, prims.ctypes.data))
        return out, nodes, prims

    # ---- rendering
    def render_pass(self, pass_index: int, seed: int):
        self._chk(self._f("render_pass")(self._h, pass_index, seed))

    def render_slice(self, pass_index: int, seed: int, s_begin: int, s_end: int):
        self._chk(self._f("render_slice")(self._h, pass_index, seed, s_begin, s_end))

    def render_samples(self, pass_index, seed, px, py, sample):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32)
        sample = np.ascontiguousarray(sample, np.uint32); n = len(px)
        L = np.zeros((n, 16), np.float32); xy = np.zeros((n, 2), np.float32)
        self._chk(self._f("render_samples")(self._h, pass_index, seed, px.ctypes.data, py.ctypes.data, sample.ctypes.data, n,
                                            L.ctypes.data, xy.ctypes.data))
        return L, xy

    # ---- light tracer (Renderer/LightTracer.hs)
    def light_trace(self, pass_index: int, seed: int, first_photon: int, n_photons: int):
        """photons [first, first + n) of a pass into the splat buffer (asynchronous like render_slice)"""
        self._chk(self._f("light_trace")(self._h, pass_index, seed, first_photon, n_photons))

    def light_trace_records(self, pass_index: int, seed: int, first_photon: int, n_photons: int, max_records: int = None) -> np.ndarray:
        """the same, returning the splats as rows {photon, depth, px, py, X, Y, Z} in (photon, depth) order (parity)"""
        cap = max_records if max_records is not None else 64 * n_photons + 16
        out = np.zeros((cap, 7), np.float32); n = C.c_size_t()
        self._chk(self._f("light_trace_records")(self._h, pass_index, seed, first_photon, n_photons, out.ctypes.data, cap, C.byref(n)))
        return out[:min(cap, n.value)]

    def read_splat(self, out: np.ndarray = None) -> np.ndarray:
        """the splat buffer _imgS: [H][W]{X, Y, Z}; a pixel of the final image is splat_weight * splat + film / weight (Image.hs:303-315)"""
        if out is None:
            out = np.zeros((self.scene.height, self.scene.width, 3), np.float32)
        self._chk(self._f("read_splat")(self._h, out.ctypes.data))
        return out

    def eval_texture(self, texture: int, p, uv) -> np.ndarray:
        """`Texture a` at explicit (dgP, (dgU, dgV)) points: (n, 16) spectra; scalar textures answer in column 0."""
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3); uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((len(p), 16), np.float32)
        self._chk(self._f("eval_texture")(self._h, texture, p.ctypes.data, uv.ctypes.data, len(p), out.ctypes.data))
        return out

    def read_film(self, out: np.ndarray = None) -> np.ndarray:
        if out is None:
            out = np.zeros((self.scene.height, self.scene.width, 4), np.float32)
        self._chk(self._f("read_film")(self._h, out.ctypes.data))
        return out

    def clear_film(self): self._chk(self._f("clear_film")(self._h))

    def film_add_host(self, film: np.ndarray):
        film = np.ascontiguousarray(film, np.float32)
        self._chk(self._f("film_add_host")(self._h, film.ctypes.data))

    def film_device(self):
        p = C.c_void_p(); n = C.c_size_t()
        self._chk(self._f("film_device")(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- multi-GPU film sum (the library owns NCCL; include/blingcu.h "multi-GPU")
    @classmethod
    def comm_unique_id(cls) -> bytes:
        """128 opaque bytes from rank 0, to be handed to every other rank (ncclGetUniqueId)."""
        L = load_library(cls._lib_path, cls._prefix)
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        rc = getattr(L, f"{cls._prefix}_comm_unique_id")(buf)
        if rc != 0:
            raise BlingCuError(rc, (getattr(L, f"{cls._prefix}_last_error")(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, rank: int, nranks: int, comm_id: bytes = None):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id) if comm_id is not None else None
        self._chk(self._f("comm_init")(self._h, rank, nranks, buf))

    @staticmethod
    def _handles(ctxs):
        return (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])

    @classmethod
    def comm_init_all(cls, ctxs):
        """one process driving several contexts (one per device): rank i = ctxs[i]."""
        rc = ctxs[0]._f("comm_init_all")(cls._handles(ctxs), len(ctxs))
        for c in ctxs:
            if rc != 0: c._chk(rc)

    @classmethod
    def reduce_film_group(cls, ctxs, root: int = -1):
        rc = ctxs[0]._f("reduce_film_group")(cls._handles(ctxs), len(ctxs), root)
        if rc != 0: ctxs[0]._chk(rc)

    def comm_destroy(self): self._chk(self._f("comm_destroy")(self._h))

    def reduce_film(self, root: int = -1):
        """film_sum = sum over ranks of film (asynchronous; overlaps the next render call)."""
        self._chk(self._f("reduce_film")(self._h, root))

    def comm_wait(self): self._chk(self._f("comm_wait")(self._h))

    def read_film_sum(self, out: np.ndarray = None) -> np.ndarray:
        if out is None:
            out = np.zeros((self.scene.height, self.scene.width, 4), np.float32)
        self._chk(self._f("read_film_sum")(self._h, out.ctypes.data))
        return out

    def film_sum_device(self):
        p = C.c_void_p(); n = C.c_size_t()
        self._chk(self._f("film_sum_device")(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def synchronize(self): self._chk(self._f("synchronize")(self._h))

    def set_stream(self, cuda_stream: int | None):
        """run on a caller-owned CUDA stream (raw cudaStream_t as int, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._chk(self._f("set_stream")(self._h, C.c_void_p(cuda_stream or 0)))

    def stats(self) -> dict:
        s = IR.Stats(); self._chk(self._f("get_stats")(self._h, C.byref(s))); return s.as_dict()

    def reset_stats(self): self._chk(self._f("reset_stats")(self._h))

    KERNEL_CLASSES = ["raygen", "trace_nearest", "trace_any", "classify", "shade", "resolve", "film", "other"]

    def kernel_times(self) -> dict:
        """device ms and launch count per kernel class since reset_stats (needs option profile_kernels=1)."""
        ms = np.zeros(8, np.float64); n = np.zeros(8, np.uint64)
        self._chk(self._f("kernel_times")(self._h, ms.ctypes.data, n.ctypes.data, 8))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}
