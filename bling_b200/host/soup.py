"""Synthetic triangle soup of BASELINE.json configs[4] / SURVEY.md §8(d) cfg 5 (generator spec: splitmix64
seed 0xB11D6; centroid ~U([-100,100]^3); two edge vectors ~U([-0.6,0.6]^3); 8 matte palette materials by
i mod 8; one constant white infinite light; camera lookAt (0,0,-320)->0, fov 40; box filter; path 5/3)."""
from __future__ import annotations

import numpy as np

from .. import ir as IR
from . import spectra as S
from . import transform as T
from .loader import Filter, Loader

F = np.float32
_PALETTE = [(0.75, 0.75, 0.75), (0.8, 0.3, 0.3), (0.3, 0.8, 0.3), (0.3, 0.3, 0.8),
            (0.8, 0.8, 0.3), (0.8, 0.3, 0.8), (0.3, 0.8, 0.8), (0.5, 0.5, 0.5)]


def splitmix64_uniforms(seed: int, n: int) -> np.ndarray:
    """n uniforms in [0,1) (24-bit) from the splitmix64 stream started at `seed`."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)).astype(F)


def make_soup(n_tris=10_000_000, width=3840, height=2160, nu=32, nv=32, seed=0xB11D6, max_depth=5,
              sample_depth=3) -> IR.SceneIR:
    u = splitmix64_uniforms(seed, 9 * n_tris).reshape(n_tris, 9)
    c = (u[:, 0:3] * F(200) - F(100)).astype(F)
    e1 = (u[:, 3:6] * F(1.2) - F(0.6)).astype(F)
    e2 = (u[:, 6:9] * F(1.2) - F(0.6)).astype(F)
    v0 = (c - ((e1 + e2).astype(F) / F(3)).astype(F)).astype(F)
    verts = np.concatenate([v0, (v0 + e1).astype(F), (v0 + e2).astype(F)], 1)
    ld = Loader(base=None)
    ld.ir.materials.clear(); ld.ir.textures.clear()
    for rgb in _PALETTE:
        ld.add_material(IR.MAT_MATTE, [ld.const_tex(S.rgb_refl(rgb))], [0.0])
    st = ld.st
    st.res = (width, height)
    st.filter = Filter("box")
    st.transform = T.look_at([0, 0, -320], [0, 0, 0], [0, 1, 0])
    st.camera = dict(kind="perspective", fov=F(40), lr=F(0), fd=F(10), c2w=st.transform, res=st.res)
    st.lights.insert(0, ("infinite", T.identity(), ("constant", S.rgb_illum((1, 1, 1)))))
    st.renderer = dict(kind="sampler", sampler=("stratified", nu, nv), integrator=("path", max_depth, sample_depth))
    from .loader import PrimRec
    uvs = np.tile(np.array([0, 0, 1, 0, 1, 1], F), (n_tris, 1))
    st.prims.insert(0, [PrimRec("tris", verts=verts, uvs=uvs, mats=(np.arange(n_tris) % 8).astype(np.int32))])
    ir = ld.finish(f"soup{n_tris}")
    ir.tri_prim_id = None; ir.tri_prim_id_base = 0      # prim id = triangle index
    return ir
