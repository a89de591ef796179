import sys, json
sys.path.insert(0,'/root/repo')
from pathlib import Path
from bling_b200 import api
from bling_b200.host.soup import make_soup
scene = make_soup(10_000_000, 3840, 2160, 32, 32)
for lib in sys.argv[1:]:
    class Ctx(api.Context):
        _lib_path = Path(lib)
    c = Ctx(0); c.upload_scene(scene)
    c.render_slice(1, 1, 0, 4); c.synchronize(); c.reset_stats(); c.set_option("profile_kernels", 1)
    c.render_slice(1, 1, 4, 12); c.synchronize()
    kt = c.kernel_times(); st = c.stats()
    print(Path(lib).name, "pass ms", round(st["last_pass_ms"],1), {k: round(v[0],1) for k,v in kt.items()}, flush=True)
    c.close()
