"""Stand-in for bling's `.bling` scene parser (Graphics/Bling/IO/*.hs), producing the flat scene IR.

In a real deployment the Haskell parser emits the IR; this module restates the subset of the
grammar the benchmark configs use, INCLUDING the parser quirks that change results:
  * the last `renderer {}` wins (RendererParser.hs:53-54); prim blocks and lights are PREPENDED
    (RenderJob.hs:49-52, LightParser.hs:26-27)
  * `transform` composes onto the current transform, `newTransform` resets first (RenderJob.hs:55-58)
  * OBJ: third UV/normal index read from vertex i+1 (WaveFront.hs:58-60,71-73); material intervals mix
    face-entry and triangle units so the last `usemtl` group is cut (WaveFront.hs:93-116)
Unsupported constructs raise NotImplementedError (they are outside SURVEY.md §8).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np

from .. import ir as IR
from . import spectra as S
from . import transform as T

F = np.float32


# ------------------------------------------------------------------------------------------- tokens
_TOK = re.compile(r'"[^"]*"|[{}]|,|[^\s{},"]+')


def tokenize(text: str) -> List[str]:
    text = re.sub(r"#[^\n]*", "", text)      # ParserCore.hs:107-108 comment
    return _TOK.findall(text)


class Tokens:
    def __init__(self, toks): self.t = toks; self.i = 0
    def peek(self): return self.t[self.i] if self.i < len(self.t) else None
    def next(self):
        if self.i >= len(self.t): raise ValueError("unexpected end of input")
        x = self.t[self.i]; self.i += 1; return x
    def expect(self, s):
        x = self.next()
        if x != s: raise ValueError(f"expected {s!r}, got {x!r} at token {self.i}")
    def flt(self):                           # ParserCore.hs:119-128: sign digits [. digits], no exponent
        x = self.next()
        if not re.fullmatch(r"[+-]?\d+(\.\d*)?", x): raise ValueError(f"expected float, got {x!r}")
        return F(x)
    def integ(self):
        x = self.next()
        if not re.fullmatch(r"\d+", x): raise ValueError(f"expected integer, got {x!r}")
        return int(x)
    def vec(self): return np.array([self.flt(), self.flt(), self.flt()], F)
    def named_float(self, n): self.expect(n); return self.flt()
    def named_int(self, n): self.expect(n); return self.integ()
    def named_vec(self, n): self.expect(n); return self.vec()
    def qstring(self):
        x = self.next()
        if not x.startswith('"'): raise ValueError(f"expected quoted string, got {x!r}")
        return x[1:-1]


# ------------------------------------------------------------------------------------------- filters
@dataclass
class Filter:
    kind: str = "box"
    p: Tuple[float, ...] = ()

    def size(self):                          # Filter.hs:61-66
        return (F(0.5), F(0.5)) if self.kind == "box" else (F(self.p[0]), F(self.p[1]))

    def eval(self, x, y):                    # Filter.hs:68-96 (float32)
        x, y = F(x), F(y); k = self.kind
        if k == "box": return F(1) if abs(x) < 0.5 and abs(y) < 0.5 else F(0)
        if k == "gauss":
            w, h, a = (F(v) for v in self.p)
            ex, ey = F(np.exp(F(F(-a * w) * w))), F(np.exp(F(F(-a * h) * h)))
            g = lambda d, e: max(F(0), F(F(np.exp(F(F(-a * d) * d))) - e))
            return F(g(x, ex) * g(y, ey))
        if k == "mitchell":
            w, h, b, c = (F(v) for v in self.p)
            iw, ih = F(F(1) / w), F(F(1) / h)

            def m1d(xp):
                q = F(abs(F(F(2) * xp)))
                if q > 1:
                    return F(F(F(F(F(F(-b) - F(6 * c)) * q * q * q) + F(F(F(6 * b) + F(30 * c)) * q * q)) +
                             F(F(F(-12 * b) - F(48 * c)) * q)) + F(F(8 * b) + F(24 * c))) * F(F(1) / F(6))
                return F(F(F(F(F(12) - F(9 * b) - F(6 * c)) * q * q * q) + F(F(F(-18) + F(12 * b) + F(6 * c)) * q * q)) +
                         F(F(6) - F(2 * b))) * F(F(1) / F(6))
            return F(m1d(F(x * iw)) * m1d(F(y * ih)))
        if k == "sinc":
            w, h, tau = (F(v) for v in self.p)

            def s1(v):
                if abs(v) > 1: return F(0)
                if abs(v) < 1e-5: return F(1)
                xp = F(abs(v) * F(np.pi))
                return F(F(np.sin(xp) / xp) * F(np.sin(F(xp * tau)) / F(xp * tau)))
            return F(s1(F(x * F(F(1) / w))) * s1(F(y * F(F(1) / h))))
        if k == "triangle":
            w, h = (F(v) for v in self.p)
            return F(F(max(F(0), F(w - abs(F(x * 2)))) * max(F(0), F(h - abs(F(y * 2))))) / F(w * h))
        raise ValueError(k)

    def table(self) -> np.ndarray:           # Image.hs:46-61 mkTableFilter
        fw, fh = self.size(); t = np.zeros(256, F)
        for y in range(16):
            fy = F(F(F(y) + F(0.5)) * fh / F(16))
            for x in range(16):
                fx = F(F(F(x) + F(0.5)) * fw / F(16))
                t[y * 16 + x] = self.eval(fx, fy)
        return t


# ------------------------------------------------------------------------------------------- parse state
@dataclass
class PrimRec:
    kind: str                    # "tris" | "shape"
    verts: Optional[np.ndarray] = None      # (n,9)
    uvs: Optional[np.ndarray] = None        # (n,6)
    normals: Optional[np.ndarray] = None    # (n,9) or None
    mats: Optional[np.ndarray] = None       # (n,) material ids
    shape: Optional[IR.Shape] = None
    emit: Optional[np.ndarray] = None


@dataclass
class PState:                    # ParserCore.hs:45-58
    res: Tuple[int, int] = (640, 480)
    renderer: Optional[dict] = None
    filter: Filter = field(default_factory=Filter)
    camera: Optional[dict] = None
    transform: T.Transform = field(default_factory=T.identity)
    material: int = 0
    emit: Optional[np.ndarray] = None
    lights: list = field(default_factory=list)
    prims: List[List[PrimRec]] = field(default_factory=list)   # list of blocks, already in PREPENDED order
    base: Path = Path(".")


class Loader:
    def __init__(self, base: Path, env_files: Optional[dict] = None, image_files: Optional[dict] = None):
        self.st = PState(base=base)
        self.ir = IR.SceneIR()
        self.env_files = env_files or {}
        self.image_files = image_files or {}      # name -> uint8 array (h, w[, c]) standing in for a file on disk
        self.image_ids = {}
        # IO/MaterialParser.hs:21-22 defaultMaterial
        self.st.material = self.add_material(IR.MAT_MATTE, [self.const_tex(S.rgb_refl((0.9, 0.9, 0.9)))], [0.0])
        self.st.renderer = dict(kind="sampler", sampler=("stratified", 2, 2), integrator=("path", 7, 3))  # RendererParser.hs:18-24
        self.st.camera = dict(kind="perspective", c2w=T.translate([0, 0, -5]), lr=F(0), fd=F(1), fov=F(90), res=(640, 480))

    # ---- IR pools
    def const_tex(self, spec) -> int:
        t = IR.Texture(); t.kind = IR.TEX_CONSTANT; IR.set_arr(t.s.v, spec)
        self.ir.textures.append(t); return len(self.ir.textures) - 1

    def add_material(self, kind, texs, fs) -> int:
        """fs: scalar parameters, each a float (constant) or ("tex", id) for a computed ScalarTexture."""
        m = IR.Material(); m.kind = kind
        for i in range(3): m.tex[i] = texs[i] if i < len(texs) else -1
        m.tex3 = texs[3] if len(texs) > 3 else -1
        for i, v in enumerate(fs):
            if isinstance(v, tuple): m.ftex[i] = v[1] + 1
            else: m.f[i] = float(v)
        self.ir.materials.append(m); return len(self.ir.materials) - 1

    def add_texture(self, kind, *, child=(-1, -1), aux=0, f=(), s=()) -> int:
        t = IR.Texture(); t.kind = kind; t.child[0], t.child[1] = child; t.aux = aux
        IR.set_arr(t.f, list(f)); IR.set_arr(t.s.v, list(s))
        self.ir.textures.append(t); return len(self.ir.textures) - 1

    # ---- spectra / textures (ParserCore.hs:142-176, MaterialParser.hs)
    def p_spectrum(self, tk: Tokens):
        t = tk.next()
        if t in ("rgbR", "rgbI"):
            if tk.peek().startswith("%"):
                h = tk.next()[1:]
                rgb = [F(int(h[i:i + 2], 16)) / F(255) for i in (0, 2, 4)]
            else:
                rgb = [tk.flt(), tk.flt(), tk.flt()]
            return S.rgb_refl(rgb) if t == "rgbR" else S.rgb_illum(rgb)
        if t == "spd":
            tk.expect("{"); pairs = []
            while True:
                pairs.append((tk.flt(), tk.flt()))
                if tk.peek() == ",": tk.next(); continue
                break
            tk.expect("}")
            return S.from_spd(S.IrregularSpd(pairs))
        if t == "temp": return S.black_body(tk.flt())
        raise ValueError(f"unknown spectrum type {t}")

    def p_mapping2d(self, tk: Tokens, name):   # pTextureMapping2d (MaterialParser.hs:159-178) -> the 9 floats of blingcu.h
        tk.expect(name); tk.expect("{"); mk = tk.next()
        if mk == "uv": m = [0.0, tk.flt(), tk.flt(), tk.flt(), tk.flt()]
        elif mk == "planar":
            vu, vv = tk.vec(), tk.vec()
            m = [1.0] + list(vu) + list(vv) + [tk.flt(), tk.flt()]
        else: raise ValueError(f"unknown 2d mapping {mk}")
        tk.expect("}")
        return m

    def p_mapping3d(self, tk: Tokens, name):   # pTextureMapping3d (:180-187): identityMapping3d w2t = transPoint w2t . dgP
        tk.expect(name); tk.expect("{"); mk = tk.next()
        if mk != "identity": raise ValueError(f"unknown 3d mapping {mk}")
        t = self.p_transform(tk); tk.expect("}")
        return list(t.m.ravel())

    def p_scalar_texture(self, tk: Tokens, name):
        """pScalarTexture (MaterialParser.hs:113-154). Returns a float for `constant`, else ("tex", id)."""
        tk.expect(name); tk.expect("{")
        tp = tk.next()
        if tp == "constant":
            v = tk.flt(); tk.expect("}")
            return float(v)
        if tp == "cellNoise":
            dn = tk.next()
            dist = {"euclidian": 0, "euclidian2": 1, "manhattan": 2, "chebyshev": 3}[dn]
            tid = self.add_texture(IR.STEX_CELLNOISE, aux=dist, s=self.p_mapping3d(tk, "map"))
        elif tp == "fbm":
            octaves, omega = tk.named_int("octaves"), tk.named_float("omega")
            tid = self.add_texture(IR.STEX_FBM, aux=octaves, f=[omega], s=self.p_mapping3d(tk, "map"))
        elif tp == "perlin":
            tid = self.add_texture(IR.STEX_PERLIN, s=self.p_mapping3d(tk, "map"))
        elif tp == "crystal":
            octaves = tk.named_int("octaves")
            tid = self.add_texture(IR.STEX_CRYSTAL, aux=octaves, s=self.p_mapping2d(tk, "map"))
        elif tp == "image":                  # pImageScalar (:104-111): readImageScalarMap accepts 8-bit greyscale only
            tk.expect("{"); tk.expect("file"); img = self.load_image(tk.qstring(), 1)
            tid = self.add_texture(IR.STEX_IMAGE, aux=img, s=self.p_mapping2d(tk, "map")); tk.expect("}")
        elif tp == "scale":
            a, sc = tk.flt(), tk.flt()
            inner = self.scalar_tex_id(self.p_scalar_texture(tk, "tex"))
            tid = self.add_texture(IR.STEX_SCALE, child=(inner, -1), f=[a, sc])
        else:
            raise NotImplementedError(f"scalar texture {tp} (image textures: SURVEY §8(f)2, not built)")
        tk.expect("}")
        return ("tex", tid)

    def load_image(self, name: str, channels: int) -> int:
        """decodeImage stays on the host (JuicyPixels in bling, PIL in this stand-in). 3 channels: RGB8 (RGBA8 drops alpha,
        YCbCr8 is converted) -> (c / 255) ** 2.2, i.e. `unGamma . f` of pixelSpectrum (Texture.hs:87-89); 1 channel: Y8 -> c / 255."""
        key = (name, channels)
        if key in self.image_ids: return self.image_ids[key]
        if name in self.image_files: raw = np.asarray(self.image_files[name])
        else:
            from PIL import Image
            im = Image.open(self.st.base / name)
            if channels == 1:
                if im.mode != "L": raise ValueError(f"unsupported image type {im.mode} for a scalar map (readImageScalarMap: Y8 only)")
            else: im = im.convert("RGB")
            raw = np.asarray(im)
        raw = raw.reshape(raw.shape[0], raw.shape[1], -1)[..., :channels]
        if raw.shape[2] != channels: raise ValueError("image channel count")
        a = raw.astype(F) / F(255)
        if channels == 3: a = np.power(a, F(2.2), dtype=F)
        self.ir.images.append(np.ascontiguousarray(a, F))
        self.image_ids[key] = len(self.ir.images) - 1
        return self.image_ids[key]

    def scalar_tex_id(self, v) -> int:
        """A scalar-texture table entry for either form p_scalar_texture returns."""
        return v[1] if isinstance(v, tuple) else self.add_texture(IR.STEX_CONSTANT, f=[v])

    def p_spectrum_texture(self, tk: Tokens, name) -> int:
        tk.expect(name); tk.expect("{")
        tp = tk.next()
        if tp == "constant":
            tid = self.const_tex(self.p_spectrum(tk))
        elif tp == "graphPaper":             # MaterialParser.hs:229-233
            lw = tk.flt()
            m = self.p_mapping2d(tk, "map")
            c0 = self.p_spectrum_texture(tk, "tex1"); c1 = self.p_spectrum_texture(tk, "tex2")
            if m[0] == 0.0: tid = self.add_texture(IR.TEX_GRAPHPAPER, child=(c0, c1), f=[lw] + m[1:5])
            else: tid = self.add_texture(IR.TEX_GRAPHPAPER, child=(c0, c1), aux=1, f=[lw], s=m)
        elif tp == "checker":                # MaterialParser.hs:209 -> checkerBoard (Texture.hs:209-221)
            sc = tk.vec()
            c0 = self.p_spectrum_texture(tk, "tex1"); c1 = self.p_spectrum_texture(tk, "tex2")
            t = IR.Texture(); t.kind = IR.TEX_CHECKER; t.child[0] = c0; t.child[1] = c1
            IR.set_arr(t.f, list(sc))
            self.ir.textures.append(t); tid = len(self.ir.textures) - 1
        elif tp == "image":                  # pImageTexture (:189-196) -> imageTexture over readImageTextureMap
            tk.expect("{"); tk.expect("file"); img = self.load_image(tk.qstring(), 3)
            tid = self.add_texture(IR.TEX_IMAGE, aux=img, s=self.p_mapping2d(tk, "map")); tk.expect("}")
        elif tp == "blend":                  # spectrumBlend (Texture.hs:129-141)
            c0 = self.p_spectrum_texture(tk, "tex1"); c1 = self.p_spectrum_texture(tk, "tex2")
            f = self.scalar_tex_id(self.p_scalar_texture(tk, "f"))
            tid = self.add_texture(IR.TEX_BLEND, child=(c0, c1), aux=f)
        elif tp == "gradient":               # MaterialParser.hs:212-219 -> gradient (mkGradient steps) f
            f = self.scalar_tex_id(self.p_scalar_texture(tk, "f"))
            tk.expect("steps"); tk.expect("{"); steps = []
            while True:
                pos = tk.flt(); steps.append((pos, self.p_spectrum(tk)))
                if tk.peek() == ",": tk.next(); continue
                break
            tk.expect("}")
            steps.sort(key=lambda e: e[0])    # mkGradient: sortBy (compare `on` fst), stable
            first = len(self.ir.textures)
            for pos, col in steps: self.add_texture(IR.TEX_CONSTANT, f=[pos], s=col)
            tid = self.add_texture(IR.TEX_GRADIENT, child=(first, len(steps)), aux=f)
        else:
            raise NotImplementedError(f"spectrum texture {tp} (image textures: SURVEY §8(f)2, not built)")
        tk.expect("}")
        return tid

    def computes(self, tid: int) -> bool:
        t = self.ir.textures[tid]
        if t.kind in (IR.TEX_BLEND, IR.TEX_GRADIENT, IR.TEX_IMAGE): return True
        if t.kind in (IR.TEX_GRAPHPAPER, IR.TEX_CHECKER): return self.computes(t.child[0]) or self.computes(t.child[1])
        return False

    def map_texture(self, tid: int, fn) -> int:
        """A copy of spectrum-texture tree `tid` with `fn` applied to every constant leaf. Textures only SELECT among
        constant leaves (graphPaper, checker), so f(tex dg) == tex' dg: this is how the host hands the device the
        per-band eta / k spectra of shinyMetal (frApproxEta / frApproxK, Fresnel.hs:72-78) without per-hit arithmetic."""
        t = self.ir.textures[tid]
        if self.computes(tid):
            raise NotImplementedError("shinyMetal over a blend / gradient texture: frApproxEta / frApproxK do not commute with the blend")
        if t.kind == IR.TEX_CONSTANT:
            return self.const_tex(fn(np.array(list(t.s.v), F)))
        n = IR.Texture.from_buffer_copy(t)
        n.child[0] = self.map_texture(t.child[0], fn); n.child[1] = self.map_texture(t.child[1], fn)
        self.ir.textures.append(n); return len(self.ir.textures) - 1

    def p_material_body(self, tk: Tokens) -> int:   # MaterialParser.hs:30-42
        t = tk.next()
        if t == "bumpMap":                   # pBumpMap (:44-48) -> bumpMapped d m (Reflection.hs:344-345)
            d = self.scalar_tex_id(self.p_scalar_texture(tk, "bump"))
            mid = self.p_material_body(tk)
            if self.ir.materials[mid].bump: raise NotImplementedError("nested bumpMap")
            self.ir.materials[mid].bump = d + 1
            return mid
        if t == "matte":
            kd = self.p_spectrum_texture(tk, "kd"); sig = self.p_scalar_texture(tk, "sigma")
            return self.add_material(IR.MAT_MATTE, [kd], [sig])
        if t == "glass":
            ior = self.p_scalar_texture(tk, "ior"); kr = self.p_spectrum_texture(tk, "kr"); kt = self.p_spectrum_texture(tk, "kt")
            return self.add_material(IR.MAT_GLASS, [kr, kt], [ior])
        if t == "mirror":
            return self.add_material(IR.MAT_MIRROR, [self.p_spectrum_texture(tk, "kr")], [])
        if t == "shinyMetal":                # pShinyMetal -> mkShinyMetal (Material.hs:98-108)
            kr = self.p_spectrum_texture(tk, "kr"); ks = self.p_spectrum_texture(tk, "ks"); r = self.p_scalar_texture(tk, "rough")
            return self.add_material(IR.MAT_SHINYMETAL, [self.map_texture(ks, S.fr_approx_eta), self.map_texture(ks, S.fr_approx_k),
                                                         self.map_texture(kr, S.fr_approx_eta), self.map_texture(kr, S.fr_approx_k)], [r])
        if t == "substrate":                 # pSubstrateMaterial -> mkSubstrate (Material.hs:110-127)
            kd = self.p_spectrum_texture(tk, "kd"); ks = self.p_spectrum_texture(tk, "ks"); ka = self.p_spectrum_texture(tk, "ka")
            ur = self.p_scalar_texture(tk, "urough"); vr = self.p_scalar_texture(tk, "vrough"); dp = self.p_scalar_texture(tk, "depth")
            return self.add_material(IR.MAT_SUBSTRATE, [kd, ks, ka], [ur, vr, dp])
        if t == "transMatte":                # pMatteTranslucent -> translucentMatte (Material.hs:43-53)
            kr = self.p_spectrum_texture(tk, "kr"); kt = self.p_spectrum_texture(tk, "kt"); sg = self.p_scalar_texture(tk, "ks")
            return self.add_material(IR.MAT_TRANSMATTE, [kr, kt], [sg])
        if t == "plastic":
            kd = self.p_spectrum_texture(tk, "kd"); ks = self.p_spectrum_texture(tk, "ks"); r = self.p_scalar_texture(tk, "rough")
            return self.add_material(IR.MAT_PLASTIC, [kd, ks], [r])
        if t == "metal":
            eta = self.p_spectrum_texture(tk, "eta"); k = self.p_spectrum_texture(tk, "k"); r = self.p_scalar_texture(tk, "rough")
            return self.add_material(IR.MAT_METAL, [eta, k], [r])
        if t == "blackbody":
            return self.add_material(IR.MAT_BLACKBODY, [], [])
        raise NotImplementedError(f"material {t} (outside SURVEY §8)")

    # ---- transforms (TransformParser.hs)
    def p_transform(self, tk: Tokens) -> T.Transform:
        tk.expect("{"); t = T.identity()
        while tk.peek() != "}":
            n = tk.next()
            if n == "rotateX": x = T.rotate(0, tk.flt())
            elif n == "rotateY": x = T.rotate(1, tk.flt())
            elif n == "rotateZ": x = T.rotate(2, tk.flt())
            elif n == "scale": x = T.scale(tk.vec())
            elif n == "translate": x = T.translate(tk.vec())
            elif n == "lookAt":
                tk.expect("{"); x = T.look_at(tk.named_vec("pos"), tk.named_vec("look"), tk.named_vec("up")); tk.expect("}")
            elif n == "matrix":
                tk.expect("{"); rows = []
                for _ in range(4):
                    tk.expect("m"); rows.append([tk.flt() for _ in range(4)])
                tk.expect("}"); x = T.from_matrix(np.array(rows, F))
            else: raise ValueError(f"unknown transform {n}")
            t = t * x                          # mconcat: apply in order
        tk.expect("}")
        return t

    # ---- primitives (PrimitiveParser.hs)
    def p_shape(self, tk: Tokens) -> IR.Shape:
        tk.expect("{"); t = tk.next(); s = IR.Shape()
        if t == "box":
            a, b = tk.named_vec("pmin"), tk.named_vec("pmax")
            s.kind = IR.SHAPE_BOX; IR.set_arr(s.p, list(np.minimum(a, b)) + list(np.maximum(a, b)))   # Shape.hs:40-44
        elif t == "cylinder":
            r, z0, z1, pm = tk.named_float("radius"), tk.named_float("zmin"), tk.named_float("zmax"), tk.named_float("phiMax")
            s.kind = IR.SHAPE_CYLINDER; IR.set_arr(s.p, [r, min(z0, z1), max(z0, z1), T._radians(np.clip(pm, 0, 360))])
        elif t == "disk":
            h, r0, r1, pm = tk.named_float("height"), tk.named_float("radius"), tk.named_float("innerRadius"), tk.named_float("phiMax")
            s.kind = IR.SHAPE_DISK; IR.set_arr(s.p, [h, max(r0, r1), min(r0, r1), T._radians(np.clip(pm, 0, 360))])
        elif t == "quad":
            s.kind = IR.SHAPE_QUAD; IR.set_arr(s.p, [tk.flt(), tk.flt()])
        elif t == "sphere":
            s.kind = IR.SHAPE_SPHERE; IR.set_arr(s.p, [tk.named_float("radius")])
        else: raise ValueError(f"unknown shape {t}")
        tk.expect("}")
        return s

    @staticmethod
    def triangulate(face):                   # TriangleMesh.hs:23-29: fan (f0, fi, fi+1)
        return [(face[0], face[i], face[i + 1]) for i in range(1, len(face) - 1)]

    def p_primitive(self, tk: Tokens) -> List[PrimRec]:
        st = self.st
        tk.expect("{"); t = tk.next()
        if t == "mesh":                      # pMesh, PrimitiveParser.hs:129-138
            vc, fc = tk.named_int("vertexCount"), tk.named_int("faceCount")
            vs = []
            for _ in range(vc): tk.expect("v"); vs.append(tk.vec())
            tris = []
            for _ in range(fc):
                tk.expect("f"); face = []
                while re.fullmatch(r"\d+", tk.peek() or ""): face.append(tk.integ())
                tris += self.triangulate(face)
            p = T.trans_points(st.transform, np.array(vs, F).reshape(-1, 3))
            idx = np.array(tris, np.int64).reshape(-1, 3)
            verts = p[idx].reshape(-1, 9)
            uvs = np.tile(np.array([0, 0, 1, 0, 1, 1], F), (len(verts), 1))   # TriangleMesh.hs:119-120
            out = [PrimRec("tris", verts=verts, uvs=uvs, mats=np.full(len(verts), st.material, np.int32))]
        elif t == "shape":
            s = self.p_shape(tk)
            IR.set_arr(s.o2w, st.transform.m); IR.set_arr(s.w2o, st.transform.i)
            s.material = st.material; s.light = -1
            out = [PrimRec("shape", shape=s, emit=None if st.emit is None else st.emit.copy())]
        elif t == "waveFront":
            fname = tk.qstring()
            tk.expect("materials"); tk.expect("{")
            mmap = {}
            while tk.peek() != "}":
                n = tk.qstring(); tk.expect("{"); mmap[n] = self.p_material_body(tk); tk.expect("}")
            tk.expect("}")
            out = [self.wavefront(st.base / fname, mmap)]
        elif t == "heightMap":               # PrimitiveParser.hs:39-45 -> heightMap elev (ns, nt) mat tr (Primitive/Heightmap.hs:22-49)
            ns, nt = tk.integ(), tk.integ()
            elev = self.p_scalar_map2d(tk)
            tr = self.p_transform(tk)         # its OWN transform block, not the current transform
            out = [self.height_map(elev, ns, nt, tr)]
        elif t == "bezier":                  # PrimitiveParser.hs:32-37 -> tesselateBezier (Primitive/Bezier.hs:80-105)
            subdivs = tk.named_int("subdivs"); patches = []
            while tk.peek() == "p":
                tk.next(); tk.expect("{"); xs = []
                while tk.peek() != "}":
                    xs.append(tk.flt())
                    if tk.peek() == ",": tk.next()
                tk.expect("}")
                if len(xs) != 48: raise ValueError("error parsing bezier patch: must give 48 values per patch")
                patches.append(np.array(xs, F))
            out = [self.tesselate_bezier(subdivs, patches)]
        else:
            raise NotImplementedError(f"primitive {t} (outside SURVEY §8)")
        tk.expect("}")
        return out

    def p_scalar_map2d(self, tk: Tokens):      # pScalarMap2d (MaterialParser.hs:239-251): a function (x, y) -> Float
        from . import noise
        tk.expect("{"); tp = tk.next()
        if tp == "fbm":                        # texMap3dTo2d (fbm octaves omega) z
            z = tk.flt(); octaves = tk.named_int("octaves"); omega = tk.named_float("omega")
            fn = lambda x, y: noise.fbm(octaves, omega, (F(x), F(y), F(z)))
        elif tp == "scale":
            f = tk.flt(); inner = self.p_scalar_map2d(tk)
            fn = lambda x, y: F(F(f) * inner(x, y))
        else: raise ValueError(f"unknown scalar map type{tp}")
        tk.expect("}")
        return fn

    def height_map(self, elev, ns: int, nt: int, tr: T.Transform) -> PrimRec:
        """Primitive/Heightmap.hs:22-49 in float32: an ns x nt grid over [0,1]^2, two triangles per cell, shading normals from
        central differences of the elevation, then mkTriangleMesh with the primitive's own transform."""
        fns, fnt = F(ns), F(nt)
        coords = [(F(F(x) / F(fns - F(1))), F(F(z) / F(fnt - F(1)))) for z in range(nt) for x in range(ns)]
        ex, ez = F(F(1) / fns), F(F(1) / fnt)
        P = np.zeros((ns * nt, 3), F); N = np.zeros_like(P); UV = np.zeros((ns * nt, 2), F)
        for k, (x, z) in enumerate(coords):
            P[k] = T.trans_point(tr, np.array([x, elev(x, z), z], F))
            dx = F(elev(F(x - ex), z) - elev(F(x + ex), z)); dz = F(elev(x, F(z - ez)) - elev(x, F(z + ez)))
            v = np.array([dx, F(ex + ez), dz], F)
            l2 = F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))
            n = -(v * F(F(1) / F(np.sqrt(l2)))) if l2 != 0 else -np.array([0, 1, 0], F)     # normalize (Math.hs:358-362)
            N[k] = T.trans_normal(tr, n.astype(F))
            UV[k] = (F(x / F(fns - F(1))), F(z / F(fns - F(1))))                             # sic: both by fns - 1
        vert = lambda x, y: x + y * ns
        verts, uvs, nrms = [], [], []
        for y in range(nt - 1):
            for x in range(ns - 1):
                for tri in ((vert(x, y), vert(x + 1, y), vert(x + 1, y + 1)), (vert(x, y), vert(x + 1, y + 1), vert(x, y + 1))):
                    verts.append(P[list(tri)].ravel()); nrms.append(N[list(tri)].ravel()); uvs.append(UV[list(tri)].ravel())
        n = len(verts)
        return PrimRec("tris", verts=np.array(verts, F).reshape(n, 9), uvs=np.array(uvs, F).reshape(n, 6),
                       normals=np.array(nrms, F).reshape(n, 9), mats=np.full(n, self.st.material, np.int32))

    def tesselate_bezier(self, subdivs: int, patches) -> PrimRec:
        """Primitive/Bezier.hs:26-105 in float32: (subdivs+1)^2 vertices per patch with p, the normal dpdu x dpdv (not normalised)
        and uv = (i, j) * step; two triangles (v00 v10 v01) (v10 v11 v01) per cell; then mkTriangleMesh (TriangleMesh.hs:39-60):
        points through o2w, normals through transNormal o2w."""
        st = self.st
        step = F(1) / F(subdivs)
        us = [F(F(i) * step) for i in range(subdivs + 1)]

        def bern(u):
            i = F(1) - u
            return [F(F(F(1) * i) * i) * i, F(F(F(3) * u) * i) * i, F(F(F(3) * u) * u) * i, F(F(F(1) * u) * u) * u]

        def dbern(u):
            i = F(1) - u
            return [F(F(3) * -F(i * i)), F(F(3) * F(F(i * i) - F(F(F(2) * u) * i))), F(F(3) * F(F(F(F(2) * u) * i) - F(u * u))), F(F(3) * F(u * u))]

        B, D = [bern(u) for u in us], [dbern(u) for u in us]
        vstride = subdivs + 1
        verts, uvs, nrms = [], [], []
        for ctrl in patches:
            def ev(bj, bi):                  # s o = sum [c (i*12 + j*3 + o) * bj j * bi i | i <- [0..3], j <- [0..3]]
                out = np.zeros(3, F)
                for o in range(3):
                    acc = F(0)
                    for i in range(4):
                        for j in range(4): acc = F(acc + F(F(ctrl[i * 12 + j * 3 + o] * bj[j]) * bi[i]))
                    out[o] = acc
                return out
            P = np.zeros((vstride * vstride, 3), F); N = np.zeros_like(P); UV = np.zeros((vstride * vstride, 2), F)
            for i in range(vstride):         # evalv (bernstein (i * step)) ...: the OUTER index feeds `bu`
                for j in range(vstride):
                    k = i * vstride + j
                    pt, dpdu, dpdv = ev(B[i], B[j]), ev(D[i], B[j]), ev(B[i], D[j])
                    P[k] = T.trans_point(st.transform, pt)
                    n = np.array([F(dpdu[1] * dpdv[2]) - F(dpdu[2] * dpdv[1]), -(F(dpdu[0] * dpdv[2]) - F(dpdu[2] * dpdv[0])),
                                  F(dpdu[0] * dpdv[1]) - F(dpdu[1] * dpdv[0])], F)   # cross (Math.hs:345-348)
                    N[k] = T.trans_normal(st.transform, n)
                    UV[k] = (us[i], us[j])
            for i in range(subdivs):
                for j in range(subdivs):
                    v00, v10 = i * vstride + j, (i + 1) * vstride + j
                    v01, v11 = i * vstride + j + 1, (i + 1) * vstride + j + 1
                    for tri in ((v00, v10, v01), (v10, v11, v01)):
                        verts.append(P[list(tri)].ravel()); nrms.append(N[list(tri)].ravel()); uvs.append(UV[list(tri)].ravel())
        n = len(verts)
        return PrimRec("tris", verts=np.array(verts, F).reshape(n, 9), uvs=np.array(uvs, F).reshape(n, 6),
                       normals=np.array(nrms, F).reshape(n, 9), mats=np.full(n, st.material, np.int32))

    def wavefront(self, path: Path, mmap: dict) -> PrimRec:   # IO/WaveFront.hs
        pts, nrm, uvs, faces, mtls = [], [], [], [], []
        for line in path.read_text().split("\n"):
            if line.startswith("vn "):
                v = np.array([F(x) for x in line.split()[1:4]], F); nrm.append(T._normalize(v))
            elif line.startswith("vt "):
                f = line.split()[1:]; uvs.append((F(f[0]), F(f[1]) if len(f) > 1 else F(1)))
            elif line.startswith("v "):
                pts.append([F(x) for x in line.split()[1:4]])
            elif line.startswith("f"):
                ent = []
                for w in line[1:].split():
                    parts = w.split("/")
                    vi = int(parts[0]); ti = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    ent.append((vi - 1, ti - 1, ni - 1))
                for tri in self.triangulate(ent): faces += list(tri)
            elif line.startswith("usemtl"):
                mtls.append((line[7:], len(faces)))       # first face ENTRY index (WaveFront.hs:133-138)
        pst = T.trans_points(self.st.transform, np.array(pts, F).reshape(-1, 3))
        cnt = len(faces) // 3
        starts = [("default", 0)] + mtls                   # matIntervals, WaveFront.hs:93-97
        ends = [s for _, s in mtls] + [cnt]
        verts, tuv, tnr, mats = [], [], [], []
        has_n = False
        for (name, s), e in zip(starts, ends):
            l = e - s
            if l <= 0: continue
            mid = mmap.get(name, self.st.material)         # mkMaterialMap default (Material.hs:24-29)
            for i in range(s, s + l, 3):
                if i + 2 >= len(faces): raise IndexError("wavefront triangle out of bounds")
                f0, f1, f2 = faces[i], faces[i + 1], faces[i + 2]
                verts.append(np.concatenate([pst[f0[0]], pst[f1[0]], pst[f2[0]]]))
                i1, i2, i3 = f0[1], f1[1], f1[1]           # Q8: third UV index from vertex i+1
                if i1 >= 0 and i2 >= 0 and i3 >= 0: tuv.append([*uvs[i1], *uvs[i2], *uvs[i3]])
                else: tuv.append([0, 0, 1, 0, 1, 1])
                n1, n2, n3 = f0[2], f1[2], f1[2]           # Q8 for normals; NOT transformed (WaveFront.hs:110-111)
                if n1 < 0 and n2 < 0 and n3 < 0: tnr.append(None)
                else: has_n = True; tnr.append(np.concatenate([nrm[n1], nrm[n2], nrm[n3]]))
                mats.append(mid)
        normals = None
        if has_n:
            if any(n is None for n in tnr): raise NotImplementedError("mixed shaded/flat OBJ triangles")
            normals = np.array(tnr, F)
        return PrimRec("tris", verts=np.array(verts, F).reshape(-1, 9), uvs=np.array(tuv, F).reshape(-1, 6),
                       normals=normals, mats=np.array(mats, np.int32))

    # ---- lights (LightParser.hs, MaterialParser.hs:244-272)
    def p_light(self, tk: Tokens):
        tk.expect("{"); t = tk.next()
        if t == "infinite":
            w2l = self.p_transform(tk)
            tk.expect("l"); tk.expect("{"); mt = tk.next()
            if mt == "constant": env = ("constant", self.p_spectrum(tk))
            elif mt == "file": env = ("file", tk.qstring())
            elif mt == "sunSky": env = ("sunsky", tk.named_vec("east"), tk.named_vec("sunDir"), tk.named_float("turbidity"))
            else: raise ValueError(f"unknown map type {mt}")
            tk.expect("}")
            light = ("infinite", w2l, env)
        elif t == "point": light = ("point", self.named_spectrum(tk, "intensity"), tk.named_vec("position"))
        elif t == "directional": light = ("directional", self.named_spectrum(tk, "intensity"), tk.named_vec("normal"))
        else: raise ValueError(f"unknown light type {t}")
        tk.expect("}")
        self.st.lights.insert(0, light)                    # ls : lights s

    def named_spectrum(self, tk, n): tk.expect(n); return self.p_spectrum(tk)

    # ---- top level (RenderJob.hs:44-63)
    def parse(self, text: str):
        tk = Tokens(tokenize(text)); st = self.st
        while tk.peek() is not None:
            o = tk.next()
            if o == "filter":
                t = tk.next()
                n = {"box": 0, "gauss": 3, "sinc": 3, "triangle": 2, "mitchell": 4}[t]
                st.filter = Filter(t, tuple(tk.flt() for _ in range(n)))
            elif o == "prim": st.prims.insert(0, self.p_primitive(tk))      # p ++ prims s
            elif o == "imageSize": st.res = (tk.integ(), tk.integ())
            elif o == "renderer": self.p_renderer(tk)
            elif o == "transform": st.transform = self.p_transform(tk) * st.transform   # t <> transform s
            elif o == "newTransform": st.transform = T.identity(); st.transform = self.p_transform(tk) * st.transform
            elif o == "camera": self.p_camera(tk)
            elif o == "light": self.p_light(tk)
            elif o == "material": tk.expect("{"); st.material = self.p_material_body(tk); tk.expect("}")
            elif o == "emission":
                tk.expect("{")
                if tk.peek() == "none": tk.next(); st.emit = None
                else: st.emit = self.p_spectrum(tk)
                tk.expect("}")
            else: raise ValueError(f"unknown object type {o}")

    def p_renderer(self, tk: Tokens):          # RendererParser.hs:26-74
        tk.expect("{"); t = tk.next()
        if t == "sampler":
            tk.expect("sampled"); tk.expect("{")
            tk.expect("sampler"); tk.expect("{"); sk = tk.next()
            smp = ("stratified", tk.integ(), tk.integ()) if sk == "stratified" else ("random", tk.integ(), 1)
            tk.expect("}")
            tk.expect("integrator"); tk.expect("{"); ik = tk.next()
            if ik == "path": integ = ("path", tk.named_int("maxDepth"), tk.named_int("sampleDepth"))
            elif ik == "directLighting": integ = ("directLighting", tk.named_int("maxDepth"), 0)   # IntegratorParser: mkDirectLightingIntegrator md
            elif ik == "bidir": integ = ("bidir", tk.named_int("maxDepth"), tk.named_int("sampleDepth"))   # IntegratorParser: mkBidirPathIntegrator md sd
            elif ik == "debug":               # IntegratorParser.hs:27-35: kdtree and reference are `undefined` in Debug.hs
                dt = tk.next()
                if dt != "normals": raise NotImplementedError(f"debug integrator {dt} (undefined in the reference, Integrator/Debug.hs:19-21,35-36)")
                integ = ("normals", 0, 0)
            elif ik == "bidirnod": raise NotImplementedError("integrator bidirnod: `contrib True` is `undefined` in the reference (Integrator/BidirPath.hs:63-65)")
            else: raise NotImplementedError(f"integrator {ik} (outside SURVEY §8)")
            tk.expect("}"); tk.expect("}"); tk.expect("}")
            self.st.renderer = dict(kind="sampler", sampler=smp, integrator=integ)
        elif t == "light":                     # RendererParser.hs:28-30: mkLightTracer ppp
            n = tk.named_int("passPhotons"); tk.expect("}")
            # the sampler / integrator fields of the IR are unused by blingcu_light_trace; they keep their defaults
            self.st.renderer = dict(kind="light", pass_photons=n, sampler=("stratified", 2, 2), integrator=("path", 7, 3))
        else:                                  # other renderers are recorded so "last one wins" stays visible
            depth = 1
            while depth: x = tk.next(); depth += (x == "{") - (x == "}")
            self.st.renderer = dict(kind=t)

    def p_camera(self, tk: Tokens):            # CameraParser.hs:18-45
        tk.expect("{"); t = tk.next(); st = self.st
        if t == "perspective":
            st.camera = dict(kind="perspective", fov=tk.named_float("fov"), lr=tk.named_float("lensRadius"),
                             fd=tk.named_float("focalDistance"), c2w=st.transform, res=st.res)
        elif t == "environment": st.camera = dict(kind="environment", c2w=st.transform, res=st.res)
        else: raise ValueError(f"unknown camera type {t}")
        tk.expect("}")

    # ---- finish: mkScene (Scene.hs:37-43) + mkJob
    def finish(self, name="") -> IR.SceneIR:
        st, ir = self.st, self.ir
        if st.renderer.get("kind") not in ("sampler", "light"):
            raise ValueError(f"active renderer is {st.renderer['kind']!r}, not the sampler or the light-tracer renderer (SURVEY F10)")
        prims = [p for block in st.prims for p in block]
        # explicit lights first, then geometric lights in prim order
        lights = []
        for l in st.lights: lights.append(self.mk_light(l))
        verts, uvs, nrm, mats, tpid = [], [], [], [], []
        any_n = any(p.kind == "tris" and p.normals is not None for p in prims)
        pid = 0
        for p in prims:
            if p.kind == "tris":
                n = len(p.verts)
                verts.append(p.verts); uvs.append(p.uvs); mats.append(p.mats)
                if any_n:
                    # a mesh without normals beside smooth ones: nine zeros per triangle = flat shaded (blingcu.h)
                    nrm.append(p.normals if p.normals is not None else np.zeros((n, 9), np.float32))
                tpid.append(np.arange(pid, pid + n, dtype=np.int32)); pid += n
            else:
                s = p.shape; s.prim_id = pid; pid += 1
                if p.emit is not None:       # Geometry.hs:27-29 mkAreaLight
                    l = IR.Light(); l.kind = IR.LIGHT_AREA; l.shape = len(ir.shapes); l.env = -1
                    IR.set_arr(l.s.v, p.emit)
                    s.light = len(lights); lights.append(l)
                ir.shapes.append(s)
        if verts:
            ir.tri_verts = np.concatenate(verts); ir.tri_uvs = np.concatenate(uvs); ir.tri_material = np.concatenate(mats)
            ir.tri_prim_id = np.concatenate(tpid)
            if any_n: ir.tri_normals = np.concatenate(nrm)
        ir.lights = lights
        # camera (Camera.hs:105-147)
        cam = st.camera; c = IR.Camera(); sx, sy = (F(v) for v in cam["res"])
        IR.set_arr(c.cam2world, cam["c2w"].m)
        if cam["kind"] == "perspective":
            c.kind = IR.CAM_PERSPECTIVE
            IR.set_arr(c.raster2cam, perspective_raster2cam(cam["fov"], sx, sy))
            ir.cam_fov = float(cam["fov"])
            c.lens_radius, c.focal_distance = float(cam["lr"]), float(cam["fd"])
        else:
            c.kind = IR.CAM_ENVIRONMENT; c.env_sx, c.env_sy = float(sx), float(sy)
            IR.set_arr(c.raster2cam, np.eye(4))
        ir.camera = c
        ir.width, ir.height = st.res
        fw, fh = st.filter.size(); ir.filter_w, ir.filter_h = float(fw), float(fh); ir.filter_table = st.filter.table()
        smp = st.renderer["sampler"]; integ = st.renderer["integrator"]
        ir.sampler_kind = IR.SAMPLER_STRATIFIED if smp[0] == "stratified" else IR.SAMPLER_RANDOM
        ir.nu, ir.nv = smp[1], smp[2]; ir.max_depth, ir.sample_depth = integ[1], integ[2]
        ir.integrator_kind = {"directLighting": IR.INTEGRATOR_DIRECT, "normals": IR.INTEGRATOR_NORMALS, "bidir": IR.INTEGRATOR_BIDIR}.get(integ[0], IR.INTEGRATOR_PATH)
        ir.cie_x, ir.cie_y, ir.cie_z, ir.cie_y_sum = S.CIE_X, S.CIE_Y, S.CIE_Z, float(S.CIE_Y_SUM)
        ir.illum_basis = np.stack(S.ILLUM).astype(F)
        ir.refl_basis = np.stack(S.REFL).astype(F)
        ir.name = name
        ir.pass_photons = 0
        if st.renderer["kind"] == "light":     # the camera connection needs world2raster / pixel_area (Camera.hs:78-103)
            ir = with_light_tracer_camera(ir); ir.pass_photons = int(st.renderer["pass_photons"])
        return ir

    def mk_light(self, l) -> IR.Light:
        ir = self.ir; out = IR.Light(); out.shape = -1; out.env = -1
        if l[0] == "point": out.kind = IR.LIGHT_POINT; IR.set_arr(out.s.v, l[1]); IR.set_arr(out.v, l[2])
        elif l[0] == "directional":
            out.kind = IR.LIGHT_DIRECTIONAL; IR.set_arr(out.s.v, l[1]); IR.set_arr(out.v, T._normalize(l[2]))   # Light.hs:53-54
        else:
            from .envmap import make_envmap
            out.kind = IR.LIGHT_INFINITE
            e, arrays = make_envmap(l[2], l[1], self.st.base, self.env_files)
            out.env = len(ir.envs); ir.envs.append(e); ir.env_arrays.append(arrays)
        return out


def perspective_raster2cam(fov, sx, sy) -> np.ndarray:
    """mkProjective / mkPerspectiveCamera (Camera.hs:105-147): the raster-to-camera matrix."""
    sx, sy = F(sx), F(sy)
    proj = T.perspective(fov, 1e-2, 1000)
    aspect = F(sx / sy)
    if aspect > 1: s0, s1, s2, s3 = -aspect, aspect, F(-1), F(1)
    else: s0, s1, s2, s3 = F(-1), F(1), F(F(-1) / aspect), F(F(1) / aspect)
    st1 = T.scale([sx, sy, 1]); st2 = T.scale([F(F(1) / F(s1 - s0)), F(F(1) / F(s2 - s3)), 1])
    tr = T.translate([-s0, -s3, 0])
    s2r = tr * st2 * st1
    return (s2r.inverse() * proj.inverse()).m


def with_light_tracer_camera(ir: IR.SceneIR) -> IR.SceneIR:
    """Fills the fields sampleCam needs (Camera.hs:78-103): world-to-raster `w2r = w2s <> s2r` with `w2s = inverse c2w <> p`, and the
    pixel area `ap = pw * ph` of mkProjective (Camera.hs:105-133). The inverse of cam2world is taken numerically here (the parser
    has it exactly, Transform.hs:120-124); the kernels and the oracle read the same matrix, so parity does not depend on it."""
    import copy
    if ir.camera.kind != IR.CAM_PERSPECTIVE:
        raise ValueError("sampleCam exists for projective cameras only (Camera.hs:101-103)")
    out = copy.copy(ir)
    out.camera = IR.Camera.from_buffer_copy(ir.camera)
    sx, sy = F(ir.width), F(ir.height)
    proj = T.perspective(ir.cam_fov, 1e-2, 1000)
    aspect = F(sx / sy)
    if aspect > 1: s0, s1, s2, s3 = -aspect, aspect, F(-1), F(1)
    else: s0, s1, s2, s3 = F(-1), F(1), F(F(-1) / aspect), F(F(1) / aspect)
    st1 = T.scale([sx, sy, 1]); st2 = T.scale([F(F(1) / F(s1 - s0)), F(F(1) / F(s2 - s3)), 1])
    s2r = T.translate([-s0, -s3, 0]) * st2 * st1
    c2w_m = np.array(list(ir.camera.cam2world), np.float64).reshape(4, 4)
    c2w = T.Transform(c2w_m.astype(F), np.linalg.inv(c2w_m).astype(F))
    w2r = c2w.inverse() * proj * s2r
    IR.set_arr(out.camera.world2raster, w2r.m)
    fl = F(np.tan(F(F(np.radians(F(ir.cam_fov))) / F(2))))
    pw, ph = F(F(fl * F(s1 - s0)) / sx), F(F(fl * F(s3 - s2)) / sy)
    out.camera.pixel_area = float(F(pw * ph))
    return out


def resized(ir: IR.SceneIR, width: int, height: int, nu: int = None, nv: int = None) -> IR.SceneIR:
    """Same scene at another film size / sampler (the camera captures the image size, CameraParser.hs:30-38)."""
    import copy
    out = copy.copy(ir)
    out.camera = IR.Camera.from_buffer_copy(ir.camera)
    out.width, out.height = width, height
    if out.camera.kind == IR.CAM_PERSPECTIVE:
        IR.set_arr(out.camera.raster2cam, perspective_raster2cam(ir.cam_fov, width, height))
    else:
        out.camera.env_sx, out.camera.env_sy = float(width), float(height)
    if nu is not None: out.nu = nu
    if nv is not None: out.nv = nv
    return out


def load_scene(path, *, fixups=(), image_size=None, sampler=None, integrator=None, filter=None, env_files=None,
               drop_lines=(), name=None, image_files=None) -> IR.SceneIR:
    """Parses a `.bling` file into the IR.
    fixups: (regex, replacement) pairs applied to the text first (documented per config in tools/make_scenes.py).
    drop_lines: 1-based line numbers removed before parsing (e.g. a trailing `renderer { sppm ... }`)."""
    path = Path(path)
    lines = path.read_text().split("\n")
    for n in drop_lines: lines[n - 1] = ""
    text = "\n".join(lines)
    for pat, rep in fixups: text = re.sub(pat, rep, text)
    if image_size is not None:               # must precede `camera {}` which captures the size (CameraParser.hs:30-38)
        text = re.sub(r"^\s*imageSize\s+\d+\s+\d+", f"imageSize {image_size[0]} {image_size[1]}", text, flags=re.M)
    ld = Loader(path.parent, env_files, image_files)
    ld.parse(text)
    if sampler is not None: ld.st.renderer["sampler"] = sampler
    if integrator is not None: ld.st.renderer["integrator"] = integrator
    if filter is not None: ld.st.filter = filter
    return ld.finish(name or path.stem)
