"""Flat scene IR: ctypes mirror of include/blingcu.h plus a numpy container with npz (de)serialisation.

The IR is what the Haskell host emits where the scene is built; in this repo the
stand-in loader (`bling_b200.loader`) produces it. Nothing here computes anything.
"""
from __future__ import annotations

import ctypes as C
import io
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

BANDS = 16

SHAPE_BOX, SHAPE_CYLINDER, SHAPE_DISK, SHAPE_QUAD, SHAPE_SPHERE = range(5)
MAT_MATTE, MAT_GLASS, MAT_MIRROR, MAT_PLASTIC, MAT_METAL, MAT_BLACKBODY, MAT_SHINYMETAL, MAT_TRANSMATTE, MAT_SUBSTRATE = range(9)
TEX_CONSTANT, TEX_GRAPHPAPER, TEX_CHECKER, TEX_BLEND, TEX_GRADIENT, TEX_IMAGE = range(6)
STEX_CONSTANT, STEX_SCALE, STEX_PERLIN, STEX_FBM, STEX_CELLNOISE, STEX_CRYSTAL, STEX_IMAGE = range(16, 23)
LIGHT_INFINITE, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_AREA = range(4)
ENV_CONSTANT, ENV_RGBTABLE, ENV_SUNSKY = range(3)
CAM_PERSPECTIVE, CAM_ENVIRONMENT = range(2)
SAMPLER_STRATIFIED, SAMPLER_RANDOM = range(2)
INTEGRATOR_PATH, INTEGRATOR_DIRECT, INTEGRATOR_NORMALS, INTEGRATOR_BIDIR = range(4)

f32 = C.c_float
i32 = C.c_int32
u32 = C.c_uint32
u64 = C.c_uint64
PF = C.POINTER(C.c_float)
PI = C.POINTER(C.c_int32)


class Spectrum(C.Structure):
    _fields_ = [("v", f32 * BANDS)]


class Shape(C.Structure):
    _fields_ = [("kind", i32), ("material", i32), ("light", i32), ("prim_id", i32),
                ("p", f32 * 8), ("o2w", f32 * 16), ("w2o", f32 * 16)]


class Texture(C.Structure):
    _fields_ = [("kind", i32), ("child", i32 * 2), ("aux", i32), ("f", f32 * 8), ("s", Spectrum)]


class Material(C.Structure):
    # ftex / bump: 0 = none, else 1 + index of a scalar texture (include/blingcu.h)
    _fields_ = [("kind", i32), ("tex", i32 * 3), ("f", f32 * 3), ("tex3", i32), ("ftex", i32 * 3), ("bump", i32)]


class MaterialV1(C.Structure):
    """Material record of fixtures saved before the scalar-texture fields existed (npz without an `abi` key)."""
    _fields_ = [("kind", i32), ("tex", i32 * 3), ("f", f32 * 3), ("tex3", i32)]


class ImageC(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("channels", i32), ("_pad", i32), ("data", PF)]


class Light(C.Structure):
    _fields_ = [("kind", i32), ("shape", i32), ("env", i32), ("_pad", i32), ("v", f32 * 4), ("s", Spectrum)]


class SunSky(C.Structure):
    _fields_ = [("sun_dir", f32 * 3), ("sun_theta", f32), ("sun_disc_dir", f32 * 3), ("_pad", f32),
                ("perez_x", f32 * 5), ("perez_y", f32 * 5), ("perez_Y", f32 * 5),
                ("zenith_x", f32), ("zenith_y", f32), ("zenith_Y", f32),
                ("s0xyz", f32 * 3), ("s1xyz", f32 * 3), ("s2xyz", f32 * 3), ("sun_radiance", Spectrum)]


class EnvMap(C.Structure):
    _fields_ = [("kind", i32), ("nu", i32), ("nv", i32), ("_pad", i32), ("w2l", f32 * 16), ("l2w", f32 * 16),
                ("s", Spectrum), ("rgb", PF), ("sky", SunSky),
                ("cond_func", PF), ("cond_cdf", PF), ("cond_int", PF), ("marg_func", PF), ("marg_cdf", PF),
                ("marg_int", f32), ("_pad2", f32)]


class Camera(C.Structure):
    _fields_ = [("kind", i32), ("raster2cam", f32 * 16), ("cam2world", f32 * 16),
                ("lens_radius", f32), ("focal_distance", f32), ("env_sx", f32), ("env_sy", f32),
                ("world2raster", f32 * 16), ("pixel_area", f32)]


class SceneC(C.Structure):
    _fields_ = [("n_triangles", u64), ("tri_verts", PF), ("tri_uvs", PF), ("tri_normals", PF),
                ("tri_material", PI), ("tri_prim_id", PI), ("tri_prim_id_base", i32),
                ("n_shapes", u32), ("shapes", C.POINTER(Shape)),
                ("n_materials", u32), ("materials", C.POINTER(Material)),
                ("n_textures", u32), ("textures", C.POINTER(Texture)),
                ("n_lights", u32), ("lights", C.POINTER(Light)),
                ("n_envs", u32), ("envs", C.POINTER(EnvMap)),
                ("camera", Camera),
                ("width", i32), ("height", i32), ("filter_w", f32), ("filter_h", f32), ("filter_table", f32 * 256),
                ("sampler_kind", i32), ("nu", i32), ("nv", i32), ("max_depth", i32), ("sample_depth", i32),
                ("cie_x", Spectrum), ("cie_y", Spectrum), ("cie_z", Spectrum), ("cie_y_sum", f32),
                ("illum_basis", Spectrum * 7), ("integrator_kind", i32),
                ("n_images", u32), ("images", C.POINTER(ImageC)), ("refl_basis", Spectrum * 7)]


class Ray(C.Structure):
    _fields_ = [("o", f32 * 3), ("tmin", f32), ("d", f32 * 3), ("tmax", f32)]


class Hit(C.Structure):
    _fields_ = [("t", f32), ("prim", i32), ("b1", f32), ("b2", f32)]


class Stats(C.Structure):
    _fields_ = [("samples", u64), ("rays_camera", u64), ("rays_extension", u64), ("rays_mis", u64),
                ("rays_shadow", u64), ("dropped_samples", u64), ("nodes_traversed", u64), ("intersections", u64), ("rays_counted", u64),
                ("kernel_launches", u64), ("bvh_nodes", u64), ("bvh_leaf_items", u64), ("last_pass_ms", C.c_double), ("bvh_max_stack", u64), ("rays_mis_culled", u64), ("rays_ext_culled", u64), ("rays_mis_any", u64),
                ("photons", u64), ("rays_light", u64), ("rays_connect", u64), ("splats", u64),
                ("any_nodes_traversed", u64), ("any_intersections", u64), ("any_rays_counted", u64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class KdNode(C.Structure):   # blingcu_kdnode
    _fields_ = [("left", C.c_int32), ("right", C.c_int32), ("split", C.c_float), ("axis", C.c_int32), ("first", C.c_uint32), ("count", C.c_uint32)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("prim", np.int32), ("b1", np.float32), ("b2", np.float32)])
KDNODE_DTYPE = np.dtype([("left", np.int32), ("right", np.int32), ("split", np.float32), ("axis", np.int32), ("first", np.uint32), ("count", np.uint32)])   # blingcu_kdnode


def _fp(a: Optional[np.ndarray]):
    if a is None:
        return PF()
    return a.ctypes.data_as(PF)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return PI()
    return a.ctypes.data_as(PI)


def set_arr(dst, src):
    src = np.asarray(src, dtype=np.float32).ravel()
    for i in range(len(src)):
        dst[i] = float(src[i])


def spectrum(values) -> Spectrum:
    s = Spectrum()
    set_arr(s.v, values)
    return s


@dataclass
class EnvArrays:
    """numpy payload of one blingcu_envmap (the struct holds the POD part)."""
    rgb: Optional[np.ndarray] = None
    cond_func: Optional[np.ndarray] = None
    cond_cdf: Optional[np.ndarray] = None
    cond_int: Optional[np.ndarray] = None
    marg_func: Optional[np.ndarray] = None
    marg_cdf: Optional[np.ndarray] = None


@dataclass
class SceneIR:
    tri_verts: np.ndarray = field(default_factory=lambda: np.zeros((0, 9), np.float32))
    tri_uvs: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    tri_normals: Optional[np.ndarray] = None
    tri_material: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    tri_prim_id: Optional[np.ndarray] = None
    tri_prim_id_base: int = 0
    shapes: List[Shape] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    textures: List[Texture] = field(default_factory=list)
    lights: List[Light] = field(default_factory=list)
    envs: List[EnvMap] = field(default_factory=list)
    env_arrays: List[EnvArrays] = field(default_factory=list)
    camera: Camera = field(default_factory=Camera)
    width: int = 0
    height: int = 0
    filter_w: float = 0.5
    filter_h: float = 0.5
    filter_table: np.ndarray = field(default_factory=lambda: np.ones(256, np.float32))
    sampler_kind: int = SAMPLER_STRATIFIED
    nu: int = 2
    nv: int = 2
    max_depth: int = 7
    sample_depth: int = 3
    integrator_kind: int = INTEGRATOR_PATH
    cie_x: np.ndarray = field(default_factory=lambda: np.zeros(16, np.float32))
    cie_y: np.ndarray = field(default_factory=lambda: np.zeros(16, np.float32))
    cie_z: np.ndarray = field(default_factory=lambda: np.zeros(16, np.float32))
    cie_y_sum: float = 1.0
    illum_basis: np.ndarray = field(default_factory=lambda: np.zeros((7, 16), np.float32))
    images: List[np.ndarray] = field(default_factory=list)    # (h, w, 3) or (h, w, 1) float32, see blingcu_image
    refl_basis: np.ndarray = field(default_factory=lambda: np.zeros((7, 16), np.float32))
    name: str = ""
    cam_fov: float = 0.0   # python-side only: lets host.loader.resized() rebuild raster2cam

    @property
    def n_prims(self) -> int:
        return len(self.tri_verts) + len(self.shapes)

    @property
    def spp(self) -> int:
        return self.nu * self.nv

    def sample_extent(self):
        """Image.hs:162-168 (float32 arithmetic)."""
        fw, fh = np.float32(self.filter_w), np.float32(self.filter_h)
        h = np.float32(0.5)
        x0 = int(np.floor(h - fw)); x1 = int(np.floor(h + np.float32(self.width) + fw))
        y0 = int(np.floor(h - fh)); y1 = int(np.floor(h + np.float32(self.height) + fh))
        return x0, x1, y0, y1

    # ------------------------------------------------------------------ ctypes view
    def to_c(self):
        """Returns (SceneC, keepalive). All pointers reference numpy/ctypes buffers kept in `keepalive`."""
        keep = []
        sc = SceneC()
        tv = np.ascontiguousarray(self.tri_verts, np.float32).reshape(-1, 9)
        tu = np.ascontiguousarray(self.tri_uvs, np.float32).reshape(-1, 6)
        tm = np.ascontiguousarray(self.tri_material, np.int32)
        keep += [tv, tu, tm]
        sc.n_triangles = len(tv)
        sc.tri_verts, sc.tri_uvs, sc.tri_material = _fp(tv), _fp(tu), _ip(tm)
        if self.tri_normals is not None:
            tn = np.ascontiguousarray(self.tri_normals, np.float32).reshape(-1, 9)
            keep.append(tn)
            sc.tri_normals = _fp(tn)
        if self.tri_prim_id is not None:
            tp = np.ascontiguousarray(self.tri_prim_id, np.int32)
            keep.append(tp)
            sc.tri_prim_id = _ip(tp)
        sc.tri_prim_id_base = self.tri_prim_id_base

        def arr(items, ty):
            a = (ty * max(1, len(items)))(*items)
            keep.append(a)
            return a

        sc.n_shapes = len(self.shapes); sc.shapes = arr(self.shapes, Shape)
        sc.n_materials = len(self.materials); sc.materials = arr(self.materials, Material)
        sc.n_textures = len(self.textures); sc.textures = arr(self.textures, Texture)
        sc.n_lights = len(self.lights); sc.lights = arr(self.lights, Light)
        envs = []
        for e, a in zip(self.envs, self.env_arrays):
            e2 = EnvMap.from_buffer_copy(e)
            for name in ("rgb", "cond_func", "cond_cdf", "cond_int", "marg_func", "marg_cdf"):
                v = getattr(a, name)
                if v is not None:
                    v = np.ascontiguousarray(v, np.float32)
                    keep.append(v)
                setattr(e2, name, _fp(v))
            envs.append(e2)
        sc.n_envs = len(envs); sc.envs = arr(envs, EnvMap)
        sc.camera = self.camera
        sc.width, sc.height = self.width, self.height
        sc.filter_w, sc.filter_h = self.filter_w, self.filter_h
        set_arr(sc.filter_table, self.filter_table)
        sc.sampler_kind, sc.nu, sc.nv = self.sampler_kind, self.nu, self.nv
        sc.max_depth, sc.sample_depth = self.max_depth, self.sample_depth
        sc.integrator_kind = self.integrator_kind
        set_arr(sc.cie_x.v, self.cie_x); set_arr(sc.cie_y.v, self.cie_y); set_arr(sc.cie_z.v, self.cie_z)
        sc.cie_y_sum = self.cie_y_sum
        for i in range(7):
            set_arr(sc.illum_basis[i].v, self.illum_basis[i])
            set_arr(sc.refl_basis[i].v, self.refl_basis[i])
        imgs = []
        for a in self.images:
            a = np.ascontiguousarray(a, np.float32); keep.append(a)
            im = ImageC(); im.height, im.width, im.channels = a.shape; im.data = _fp(a)
            imgs.append(im)
        sc.n_images = len(imgs); sc.images = arr(imgs, ImageC)
        return sc, keep

    # ------------------------------------------------------------------ npz
    def save(self, path):
        d = {}

        def raw(items, ty):
            a = (ty * max(1, len(items)))(*items)
            return np.frombuffer(bytes(a), np.uint8)[: C.sizeof(ty) * len(items)].copy()

        d["tri_verts"] = np.asarray(self.tri_verts, np.float32)
        d["tri_uvs"] = np.asarray(self.tri_uvs, np.float32)
        d["tri_material"] = np.asarray(self.tri_material, np.int32)
        if self.tri_normals is not None:
            d["tri_normals"] = np.asarray(self.tri_normals, np.float32)
        if self.tri_prim_id is not None:
            d["tri_prim_id"] = np.asarray(self.tri_prim_id, np.int32)
        d["shapes"] = raw(self.shapes, Shape)
        d["materials"] = raw(self.materials, Material)
        d["textures"] = raw(self.textures, Texture)
        d["lights"] = raw(self.lights, Light)
        envs = []
        for e in self.envs:
            e2 = EnvMap.from_buffer_copy(e)
            for name in ("rgb", "cond_func", "cond_cdf", "cond_int", "marg_func", "marg_cdf"):
                setattr(e2, name, PF())
            envs.append(e2)
        d["envs"] = raw(envs, EnvMap)
        for i, a in enumerate(self.env_arrays):
            for name in ("rgb", "cond_func", "cond_cdf", "cond_int", "marg_func", "marg_cdf"):
                v = getattr(a, name)
                if v is not None:
                    d[f"env{i}_{name}"] = np.asarray(v, np.float32)
        d["camera"] = np.frombuffer(bytes(self.camera), np.uint8).copy()
        d["filter_table"] = np.asarray(self.filter_table, np.float32)
        d["scalars_i"] = np.array([self.tri_prim_id_base, self.width, self.height, self.sampler_kind, self.nu,
                                   self.nv, self.max_depth, self.sample_depth], np.int64)
        d["scalars_f"] = np.array([self.filter_w, self.filter_h, self.cie_y_sum, self.cam_fov], np.float32)
        d["cie"] = np.stack([self.cie_x, self.cie_y, self.cie_z]).astype(np.float32)
        d["illum_basis"] = np.asarray(self.illum_basis, np.float32)
        d["name"] = np.frombuffer(self.name.encode(), np.uint8)
        d["abi"] = np.array([2], np.int32)
        d["integrator_kind"] = np.array([self.integrator_kind], np.int32)
        d["refl_basis"] = np.asarray(self.refl_basis, np.float32)
        for i, a in enumerate(self.images): d[f"image{i}"] = np.asarray(a, np.float32)
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path) -> "SceneIR":
        z = np.load(path)

        def unraw(key, ty):
            b = z[key].tobytes()
            n = len(b) // C.sizeof(ty)
            return [ty.from_buffer_copy(b, i * C.sizeof(ty)) for i in range(n)]

        ir = SceneIR()
        ir.tri_verts = z["tri_verts"]; ir.tri_uvs = z["tri_uvs"]; ir.tri_material = z["tri_material"]
        ir.tri_normals = z["tri_normals"] if "tri_normals" in z else None
        ir.tri_prim_id = z["tri_prim_id"] if "tri_prim_id" in z else None
        ir.shapes = unraw("shapes", Shape)
        if "abi" in z:
            ir.materials = unraw("materials", Material)
        else:
            ir.materials = []
            for o in unraw("materials", MaterialV1):
                m = Material(); m.kind = o.kind; m.tex3 = o.tex3
                for i in range(3): m.tex[i] = o.tex[i]; m.f[i] = o.f[i]
                ir.materials.append(m)
        ir.textures = unraw("textures", Texture); ir.lights = unraw("lights", Light)
        ir.envs = unraw("envs", EnvMap)
        ir.env_arrays = []
        for i in range(len(ir.envs)):
            a = EnvArrays()
            for name in ("rgb", "cond_func", "cond_cdf", "cond_int", "marg_func", "marg_cdf"):
                k = f"env{i}_{name}"
                if k in z:
                    setattr(a, name, z[k])
            ir.env_arrays.append(a)
        cb = z["camera"].tobytes()   # fixtures written before the light-tracer fields existed are shorter: those fields stay zero
        ir.camera = Camera.from_buffer_copy(cb + bytes(max(0, C.sizeof(Camera) - len(cb))))
        ir.filter_table = z["filter_table"]
        si = z["scalars_i"]; sf = z["scalars_f"]
        (ir.tri_prim_id_base, ir.width, ir.height, ir.sampler_kind, ir.nu, ir.nv, ir.max_depth,
         ir.sample_depth) = [int(x) for x in si]
        ir.filter_w, ir.filter_h, ir.cie_y_sum, ir.cam_fov = [float(x) for x in sf]
        ir.cie_x, ir.cie_y, ir.cie_z = z["cie"][0], z["cie"][1], z["cie"][2]
        ir.illum_basis = z["illum_basis"]
        ir.name = z["name"].tobytes().decode()
        if "integrator_kind" in z: ir.integrator_kind = int(z["integrator_kind"][0])
        if "refl_basis" in z: ir.refl_basis = z["refl_basis"]
        i = 0
        while f"image{i}" in z: ir.images.append(z[f"image{i}"]); i += 1
        return ir
