#!/usr/bin/env python
"""Flattens the reference's example scenes (BASELINE.json configs 1-4) into IR fixtures under
tests/golden/scenes/. Needs /root/reference (build container only); the fixtures travel to the GPU box.
Per-config fix-ups follow SURVEY.md §8(d) / BASELINE.md and are listed inline.

    python tools/make_scenes.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bling_b200.host.envmap import synthetic_hdr  # noqa: E402
from bling_b200.host.loader import load_scene  # noqa: E402

EX = Path("/root/reference/examples")
OUT = ROOT / "tests" / "golden" / "scenes"

RGB_FIX = [(r"emission\s*\{\s*rgb\b", "emission { rgbI"), (r"\brgb\b", "rgbR")]

CONFIGS = {
    # cfg 1: the last renderer line selects SPPM (:24) -> dropped; 512x512; stratified 8 8
    "cornell-box": dict(file="cornell-box.bling", drop_lines=(24,), image_size=(512, 512), sampler=("stratified", 8, 8)),
    # cfg 2a: parses as shipped (it holds a glass BOX); 1024x1024; 64 spp
    "glass-torus": dict(file="glass-torus.bling", image_size=(1024, 1024), sampler=("stratified", 8, 8)),
    # cfg 2b: stale syntax: `stratified xSamples 2 ySamples 2`, sppm without maxDepth, graphPaper without map{}, rgb spectra
    "specular": dict(file="specular.bling", image_size=(1024, 1024), sampler=("stratified", 8, 8),
                     fixups=[(r"stratified xSamples 2 ySamples 2", "stratified 8 8"),
                             (r"renderer \{ sppm photonCount 5000 radius 0.5 \}", ""),
                             (r"graphPaper 0.1", "graphPaper 0.1 map { uv 1 1 0 0 }")] + RGB_FIX),
    # cfg 3: 1920x1080, stratified 16 16
    "ducky": dict(file="ducky.bling", image_size=(1920, 1080), sampler=("stratified", 16, 16)),
    # cfg 4a: graphPaper lacks map{}; last renderer is SPPM (:20)
    "sun-sky": dict(file="sun-sky.bling", image_size=(1920, 1080), sampler=("stratified", 8, 8), drop_lines=(20,),
                    fixups=[(r"graphPaper 0.05", "graphPaper 0.05 map { uv 10 10 0 0 }")]),
    # cfg 4b: the HDR is a missing blob -> synthetic 1024x512 map; the SPPM renderer precedes the sampler one, fine
    "environment": dict(file="environment.bling", image_size=(1920, 1080), env_files={"*": synthetic_hdr()}),
    # the one example that selects the direct-lighting integrator; parses as shipped (15 blackbody emitters)
    "blackbody-emission": dict(file="blackbody-emission.bling"),
    # SURVEY §8(f)2 "mesh shading normals (bezier ...)": 102 Bezier patches, subdivs 16 -> 26 112 smooth-shaded triangles, thin lens.
    # Stale syntax: `rgbeFile` is not a map type of LightParser.hs (-> `file`); the HDR is a missing blob -> synthetic map
    # more reference examples that parse as shipped and exercise the widened rows: bump mapping over fbm; cellNoise with its four
    # distance functions under bump + gradient; substrate with an fbm coating depth; Oren-Nayar and plastic test scenes
    "bumpmap": dict(file="bumpmap.bling"), "cellnoise": dict(file="cellnoise.bling"), "substrate": dict(file="substrate.bling"),
    "matte-test": dict(file="matte-test.bling"), "plastic-test": dict(file="plastic-test.bling"),
    # examples whose LAST renderer line selects another renderer (SPPM / Metropolis): that line dropped, the sampler renderer above
    # it becomes active, as for cfg 1. trans-matte: translucentMatte; cornell-box-specular: glass + mirror spheres; race: metal + plastic
    "trans-matte": dict(file="trans-matte.bling", drop_lines=(24,)),
    "cornell-box-specular": dict(file="cornell-box-specular.bling", drop_lines=(18,)),
    "race": dict(file="race.bling", drop_lines=(25,)),
    # blend of two constants by a quasi-crystal pattern; `rgbeFile` is stale syntax and the HDR a missing blob (as gumbo)
    "crystal": dict(file="crystal.bling", fixups=[(r"rgbeFile", "file")], env_files={"*": synthetic_hdr()}),
    # stale syntax of older parser versions, rewritten like specular / sun-sky above (metal-test: its last renderer line selects the
    # light tracer -> dropped): `stratified xSamples n ySamples m`,
    # `transform { identity ...}` (= newTransform), graphPaper without map{}, `rgbeFile`
    "shapes": dict(file="shapes.bling", fixups=[(r"stratified xSamples (\d+) ySamples (\d+)", r"stratified \1 \2"),
                                                 (r"transform \{ identity", "newTransform {")] + RGB_FIX),
    "geometric-light": dict(file="geometric-light.bling", fixups=[(r"stratified xSamples (\d+) ySamples (\d+)", r"stratified \1 \2"),
                                                                   (r"graphPaper ([\d.]+)\s+tex1", r"graphPaper \1 map { uv 1 1 0 0 } tex1")] + RGB_FIX),
    "metal-test": dict(file="metal-test.bling", drop_lines=(20,), fixups=[(r"graphPaper ([\d.]+)\s+tex1", r"graphPaper \1 map { uv 1 1 0 0 } tex1"), (r"rgbeFile", "file")],
                       env_files={"*": synthetic_hdr()}),
    # a height-map mesh (fbm elevation, central-difference shading normals) under `integrator { debug normals }`, `random 4`
    # sampler; parses as shipped
    "heightmap": dict(file="heightmap.bling"),
    # bump-mapped (fbm) glass water over a pool of BOX shapes, cylinders textured by a crystal blend, a blackbody
    # emitter; the only example with maxDepth 10, i.e. the only one that reaches the Russian-roulette branch (depth > 7,
    # Path.hs:68-72). Its last renderer line selects Metropolis (:17) -> dropped; stale `stratified xSamples n ySamples m` and the
    # one-parameter `scale s tex` of an older parser (-> `scale 0 s tex`)
    "pool": dict(file="pool.bling", drop_lines=(17,), fixups=[(r"stratified xSamples (\d+) ySamples (\d+)", r"stratified \1 \2"),
                                                              (r"scale 0.2 tex", "scale 0 0.2 tex")]),
    # the Cornell box under a height-map water surface (smooth normals) next to its flat-shaded meshes (nine-zero normals, blingcu.h),
    # glass, maxDepth 15, sinc filter; parses as shipped
    "cornell-box-underwater": dict(file="cornell-box-underwater.bling"),
    "gumbo": dict(file="gumbo.bling", fixups=[(r"rgbeFile", "file")], env_files={"*": synthetic_hdr()}),
}

# this repository's own coverage scenes (tests/golden/scenes_src/*.bling), flattened by the same loader
OWN = ["zoo", "envcam", "smooth", "extras", "textures", "direct"]


def main(only=()):
    OUT.mkdir(parents=True, exist_ok=True)
    for name, cfg in CONFIGS.items():
        if only and name not in only: continue
        cfg = dict(cfg)
        ir = load_scene(EX / cfg.pop("file"), name=name, **cfg)
        ir.save(OUT / f"{name}.npz")
        print(f"{name}: {len(ir.tri_verts)} tris, {len(ir.shapes)} shapes, {len(ir.lights)} lights, "
              f"{len(ir.materials)} materials, {ir.width}x{ir.height}, {ir.nu}x{ir.nv} spp, "
              f"depth {ir.max_depth}/{ir.sample_depth}, extent {ir.sample_extent()}")


def own():
    src = ROOT / "tests" / "golden" / "scenes_src"
    for name in OWN:
        ir = load_scene(src / f"{name}.bling", name=name)
        ir.save(OUT / f"{name}.npz")
        print(f"{name}: {ir.n_prims} prims, {len(ir.lights)} lights, {len(ir.materials)} materials, {len(ir.textures)} textures")


if __name__ == "__main__":
    if sys.argv[1:] == ["own"]: own()
    else: main(sys.argv[1:])          # no arguments: every fixture; else the named ones
