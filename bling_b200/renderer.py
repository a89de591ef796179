"""Host-side mirror of bling's renderer plug-in seam for the path-integrator hot path.

Reference interface (Graphics/Bling/Rendering.hs):
    class Renderer a where render :: a -> RenderJob -> ProgressReporter -> IO ()      (:77-78)
    data RenderJob = MkJob { jobScene, jobPixelFilter, jobImageSize }                 (:37-41)
    data Progress  = Started | ... | PassDone { progPassNum, finalImg, splatWeight }  (:60-73)
    type ProgressReporter = Progress -> IO Bool     -- False stops the progressive loop (:75, :137-138)

`CudaRenderer.render(job, report)` is `prender` (:111-140) with the tile loop replaced by the CUDA core: upload
the flat scene once, then per pass { render this rank's sample shard; sum films over ranks; report PassDone }.
Multi-GPU: every GPU holds a full scene replica and renders the sample indices [rank*spp/world, (rank+1)*spp/world) of
every pixel; the only exchange is one sum of the [H][W][4] f32 films per report, the analogue of the
reference's sequential addTile merge. The LIBRARY does it (blingcu_reduce_film: ncclAllReduce on its own stream,
overlapping the next pass); the host only hands the communicator id around:
  * `CudaRenderer`       one process per GPU (torchrun): torch.distributed is the launcher plumbing that carries the 128-byte
                         id from rank 0 to the others (an MPI broadcast or a file would do the same);
  * `MultiDeviceRenderer` one process driving n contexts (what a single Haskell process would do): blingcu_comm_init_all +
                         blingcu_reduce_film_group.
With a gloo process group (the CPU tests of the host logic, emulator contexts) the films are summed on the host instead.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import api
from . import ir as IR


@dataclass
class RenderJob:
    """mkJob scene filter size: the flat IR already carries the pixel filter table and the image size."""
    scene: IR.SceneIR

    @property
    def image_size(self):
        return (self.scene.width, self.scene.height)


@dataclass
class Started:
    pass


@dataclass
class PassDone:
    pass_num: int
    final_img: np.ndarray      # [H][W]{weight, X*w, Y*w, Z*w}: Img._imgP (Image.hs:123-129)
    splat_weight: float = 1.0


ProgressReporter = Callable[[object], bool]


def shard_range(spp: int, rank: int, world: int):
    """sample indices of one pass owned by `rank` (contiguous, balanced, covers [0,spp) exactly once)."""
    return (spp * rank) // world, (spp * (rank + 1)) // world


class CudaRenderer:
    """The `Renderer` instance backed by libblingcu.so (keyword `renderer { cuda sampled {...} }` on the bling side)."""

    def __init__(self, device: Optional[int] = None, seed: int = 0x5EED, context_cls=api.Context, process_group=None):
        self.seed = seed
        self.rank, self.world = 0, 1
        self._dist = None
        self._pg = process_group
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        except ImportError:
            pass
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0")) if self._dist is not None else 0
        self.ctx = context_cls(device) if context_cls is api.Context else context_cls()
        self._job = None

    def pretty_print(self) -> str:
        return "cuda sampler renderer"

    def _library_comm(self) -> bool:
        """the library's own communicator sums the films (real GPUs); host-side sum otherwise (gloo CPU tests)."""
        return self._dist is not None and self.world > 1 and self._dist.get_backend(self._pg) == "nccl"

    def upload(self, job: RenderJob):
        if self._library_comm() and not getattr(self, "_comm_ready", False):
            box = [type(self.ctx).comm_unique_id() if self.rank == 0 else None]
            self._dist.broadcast_object_list(box, src=0, group=self._pg)
            self.ctx.comm_init(self.rank, self.world, box[0])
            self._comm_ready = True
        self.ctx.upload_scene(job.scene)
        self._job = job

    def render_pass_shard(self, pass_num: int):
        """this rank's share of pass `pass_num` (device-side accumulate, asynchronous)."""
        s0, s1 = shard_range(self._job.scene.spp, self.rank, self.world)
        if s1 > s0:
            self.ctx.render_slice(pass_num, self.seed, s0, s1)

    def gather_film(self) -> np.ndarray:
        """sum of the per-rank films (all-reduce), returned on the host in Img._imgP layout."""
        sc = self._job.scene
        if self._dist is None or self.world == 1:
            return self.ctx.read_film()
        if self._library_comm():
            self.ctx.reduce_film()                 # asynchronous ncclAllReduce into film_sum on the library's stream
            return self.ctx.read_film_sum()        # waits for it
        import torch
        total = torch.from_numpy(self.ctx.read_film())     # gloo (CPU tests of the host logic)
        self._dist.all_reduce(total, op=self._dist.ReduceOp.SUM, group=self._pg)
        return total.numpy()

    def render(self, job: RenderJob, report: ProgressReporter, first_pass: int = 1):
        """prender (Rendering.hs:111-140): progressive passes until the reporter returns False."""
        self.upload(job)
        report(Started())
        p = first_pass
        while True:
            self.render_pass_shard(p)
            img = self.gather_film()
            cont = report(PassDone(p, img, 1.0))
            if self._dist is not None and self.world > 1:      # every rank must take the same decision
                import torch
                flag = torch.tensor([1 if cont else 0], dtype=torch.int32)
                if self._dist.get_backend(self._pg) == "nccl":
                    flag = flag.cuda()
                self._dist.broadcast(flag, src=0, group=self._pg)
                cont = bool(flag.item())
            if not cont:
                break
            p += 1

    def close(self):
        self.ctx.close()


class LightTracerRenderer:
    """`renderer { light passPhotons n }` (Renderer/LightTracer.hs:1-51, IO/RendererParser.hs:28-30) on the CUDA core: per pass
    `ppp` light paths splatted into the splat buffer, reported as PassDone n img (1 / (n * ppp)) where `img` carries the (empty)
    filtered film and the splats -- here the pair (film [H][W][4], splat [H][W][3]) -- and the reporter's False stops the loop.
    The scene's camera must carry world2raster / pixel_area (bling_b200.host.loader.with_light_tracer_camera)."""

    def __init__(self, pass_photons: int, device: int = 0, seed: int = 0x5EED, context_cls=api.Context):
        self.ppp, self.seed = int(pass_photons), seed
        self.ctx = context_cls(device) if context_cls is api.Context else context_cls()

    def pretty_print(self) -> str:
        return f"light tracer\n{self.ppp} photons per pass"

    def render(self, job: RenderJob, report: ProgressReporter, first_pass: int = 1):
        self.ctx.upload_scene(job.scene)
        report(Started())
        n = first_pass
        while True:
            self.ctx.light_trace(n, self.seed, 0, self.ppp)
            img = (self.ctx.read_film(), self.ctx.read_splat())
            if not report(PassDone(n, img, 1.0 / (n * self.ppp))):
                break
            n += 1

    def close(self):
        self.ctx.close()


def renderer_for(scene: IR.SceneIR, device: int = 0, seed: int = 0x5EED, context_cls=api.Context):
    """The renderer the job's active `renderer {}` block names (RendererParser.hs:26-54, last one wins): the light tracer for
    `renderer { light passPhotons n }` (the loader leaves n in `scene.pass_photons`), else the sampler renderer."""
    if getattr(scene, "pass_photons", 0) > 0:
        return LightTracerRenderer(scene.pass_photons, device, seed, context_cls)
    return CudaRenderer(device, seed, context_cls=context_cls)


class MultiDeviceRenderer:
    """One process, one context per device (SURVEY.md §8b `blingcu_create(devices, ndev)` in spirit): what a single Haskell
    process binding the C ABI does to use every GPU of the box. Render calls are asynchronous, so one host thread keeps all
    devices busy; blingcu_reduce_film_group sums the films (ncclGroupStart/End around one all-reduce per context)."""

    def __init__(self, devices, seed: int = 0x5EED, context_cls=api.Context):
        self.seed = seed
        self.ctxs = [context_cls(d) if context_cls is api.Context else context_cls() for d in devices]
        self._cls = context_cls
        self._cls.comm_init_all(self.ctxs)
        self._job = None

    def upload(self, job: RenderJob):
        for c in self.ctxs:
            c.upload_scene(job.scene)
        self._job = job

    def render_pass(self, pass_num: int) -> np.ndarray:
        spp, n = self._job.scene.spp, len(self.ctxs)
        for r, c in enumerate(self.ctxs):
            s0, s1 = shard_range(spp, r, n)
            if s1 > s0:
                c.render_slice(pass_num, self.seed, s0, s1)
        self._cls.reduce_film_group(self.ctxs, root=0)
        return self.ctxs[0].read_film_sum()

    def render(self, job: RenderJob, report: ProgressReporter, first_pass: int = 1):
        self.upload(job)
        report(Started())
        p = first_pass
        while report(PassDone(p, self.render_pass(p), 1.0)):
            p += 1

    def close(self):
        for c in self.ctxs:
            c.close()
