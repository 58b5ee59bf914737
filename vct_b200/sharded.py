"""One frame of the GI hot path on 1..N GPUs, one process per GPU.

world_size == 1   the frame is `vct_gi_passes` on the library's own stream.
world_size  > 1   z-slab sharding: every rank clears / voxelises / transfers / injects / mip-filters the slab
                  z in [rank*D/N, (rank+1)*D/N) of every level (BOX2 mips are aligned 2x2x2 reductions, so no
                  halo is needed while a slab is at least one texel thick at the coarsest level); then ONE exchange
                  step — an in-place NCCL all-gather per pyramid level, coalesced into a single group launch over
                  NVLink — gives every rank the whole radiance pyramid; `vct_exchange` publishes it to the 3D
                  texture; the cone trace is sharded by screen band and the bands are all-gathered into the image.
                  The library enqueues on torch's current stream, so the collectives need no host synchronisation.

The partition maths (`slab_range`, `level_chunks`, `image_bands`) is pure Python and is unit-tested on CPU with gloo.
"""
import ctypes as C
import os

import numpy as np

from . import params as P


# ------------------------------------------------------------------------------------- partition maths
def slab_range(dim, world, rank):
    """z-slab [lo, hi) owned by `rank` (vct_create uses the same rule; dim % world == 0 is required)."""
    if dim % world:
        raise ValueError("dim must be divisible by world_size")
    return dim * rank // world, dim * (rank + 1) // world


def level_chunks(dim, levels, world):
    """Per level: (offset_in_voxels_of_rank0_chunk_relative_to_level_start, voxels_per_rank).  Level l is a linear
    [z][y][x] array of (dim>>l)^3 voxels; a rank's slab of it is one contiguous chunk of d^3/world voxels."""
    out = []
    for l in range(levels):
        d = max(1, dim >> l)
        if d % world:
            raise ValueError(f"level {l} ({d}^3) is thinner than one slab per rank: use levels <= log2(dim/world)+1")
        out.append(d * d * d // world)
    return out


def image_bands(height, world):
    """Equal bands of whole 8-row tiles; returns (band_rows, [(y_lo, y_hi) per rank]) — mirrors vctk_image_rows."""
    rows8 = (height + 7) // 8
    band = (rows8 + world - 1) // world * 8
    return band, [(min(height, r * band), min(height, (r + 1) * band)) for r in range(world)]


def all_gather_levels(dist, level_tensors, chunks, rank):
    """In-place all-gather of every level (rank r's chunk lives at [r*chunk, (r+1)*chunk) of the level tensor).
    Works for any backend (gloo on CPU in the tests, NCCL over NVLink in production)."""
    for t, n in zip(level_tensors, chunks):
        dist.all_gather_into_tensor(t, t[rank * n:(rank + 1) * n].clone() if t.device.type == "cpu" else t[rank * n:(rank + 1) * n])


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, nbytes):
    import torch
    return torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")


# ------------------------------------------------------------------------------------------- the frame
class ShardedFrame:
    def __init__(self, pipeline, params, world=1, rank=0, workload=None):
        import torch
        self.g, self.p, self.world, self.rank = pipeline, params, world, rank
        # animated workloads (config 4): per-frame actor transforms, and the whole frame graph (producers included) per step
        self.workload, self.frame_index = workload, 0
        self.whole_frame = bool(workload is not None and workload.whole_frame)
        self.torch = torch
        g = pipeline
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            # A dedicated (non-default) stream shared by the library and the NCCL collectives: torch's default stream has
            # handle 0, which vct_set_stream reads as "use a private stream" — the collectives would then not be ordered
            # after the kernels that produce their input.
            self.stream = torch.cuda.Stream()
            self.stream.wait_stream(torch.cuda.current_stream())
            g.set_stream(self.stream.cuda_stream)
            which = P.VOL_RADIANCE if params.draw_radiance else P.VOL_COLOR
            self.levels = [device_tensor(g.device_ptr(which, l), g.level_bytes(which, l)) for l in range(g.L)]
            self.chunks = level_chunks(g.D, g.L, world)
            self.band, self.bands = image_bands(g.H, world)
            self.image = device_tensor(g.device_ptr(P.BUF_IMAGE, 0), g.level_bytes(P.BUF_IMAGE, 0))
            self.band_px = self.band * g.W
            # sparse exchange of level 0 over peer memory (exchange.cu): every rank maps every other rank's staging buffer
            self.sparse_exchange = False
            if os.environ.get("VCT_SPARSE_EXCHANGE", "1") != "0" and dist.get_backend() == "nccl":
                handles = [None] * world
                dist.all_gather_object(handles, g.exchange_setup())
                for r, h in enumerate(handles):
                    g.exchange_import(r, h)
                self.sparse_exchange = True
        else:
            self.stream = torch.cuda.ExternalStream(g.lib.vct_stream(g.h))

    def describe(self):
        if self.world == 1:
            return "1 GPU"
        ex = ("level 0 pushed as flagged x-row segments into every peer's memory over NVLink (cudaIpc), levels >= 1 by coalesced NCCL all-gather"
              if self.sparse_exchange else "coalesced NCCL all-gather of the radiance pyramid")
        return (f"{self.world} GPUs: z-slab sharding of clear/voxelise/transfer/inject/mip, {ex}, "
                f"cone trace sharded by {self.band}-row screen band, NCCL all-gather of the image bands")

    # producers of the reference frame graph that the GI step consumes (replicated on every rank)
    def producers(self, gbuffer=True):
        g, p = self.g, self.p
        g.shadowmap(p)
        if p.warp_texture:
            g.occupancy(p); g.warpmap(p)
        if gbuffer:
            g.gbuffer(p)

    def _gather_levels(self, first):
        dist, r = self.dist, self.rank
        levels, chunks = self.levels[first:], self.chunks[first:]
        try:
            with dist._coalescing_manager(device=self.levels[0].device):
                for t, n in zip(levels, chunks):
                    dist.all_gather_into_tensor(t, t[r * n:(r + 1) * n])
        except (AttributeError, TypeError, RuntimeError):
            for t, n in zip(levels, chunks):
                dist.all_gather_into_tensor(t, t[r * n:(r + 1) * n])

    def _exchange(self):
        """After vct_gi_passes: give every rank the whole traced pyramid, in its 3D texture.  Sparse frame: level 0 goes
        as flagged segments through peer memory; the all-gather of the small levels in between is the barrier that
        orders every rank's pushes before every rank's unpack.  Dense frame (the first, or after an invalidation): all
        levels by all-gather, dense publish.  All ranks take the same branch (same call history)."""
        g = self.g
        if self.sparse_exchange and g.frame_was_sparse():
            g.exchange_push()
            self._gather_levels(1)
            g.exchange_unpack()
            return {}
        self._gather_levels(0)
        g.exchange()
        return {}

    # ---- CUDA graph: the step is ~25 short launches (+ 2 collectives); replaying a captured pair of steps removes the
    # host from the loop.  TWO steps per graph because the segment masks swap roles every frame; the graph is only
    # replayed at the mask parity it was captured at, and only while the frame parameters stay what they were.
    def enable_graph(self):
        torch = self.torch
        self.graph, self.graph_error = None, None
        try:
            self.g.set_profiling(0)
            torch.cuda.synchronize()
            parity = self.g.mask_parity()
            n0 = self.g.launch_count()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=self.stream):
                self.step(); self.step()
            self.graph_launches_per_step = (self.g.launch_count() - n0) // 2
            self.graph, self.graph_parity = gr, parity
            torch.cuda.synchronize()
        except Exception as e:                                   # capture unsupported here: stay eager
            self.graph, self.graph_error = None, f"{type(e).__name__}: {e}"
            torch.cuda.synchronize()
        return self.graph is not None

    def run_steps(self, k):
        """k steps; pairs go through the captured graph when there is one.  Returns how many steps were replayed."""
        replayed = 0
        while k > 0:
            if getattr(self, "graph", None) is not None and k >= 2 and self.g.mask_parity() == self.graph_parity:
                with self.torch.cuda.stream(self.stream):
                    self.graph.replay()
                k -= 2; replayed += 2
            else:
                self.step(); k -= 1
        return replayed

    def _animate(self):
        if self.workload is not None:
            for actor, model in self.workload.models(self.frame_index):
                self.g.set_actor_transform(actor, model)
        self.frame_index += 1

    def step(self):
        g, p = self.g, self.p
        self._animate()
        if self.world == 1:
            (g.frame if self.whole_frame else g.gi_passes)(p)
            return
        with self.torch.cuda.stream(self.stream):
            if self.whole_frame:
                self.producers(gbuffer=False)
            g.gi_passes(p)
            self._exchange()
            if self.whole_frame:
                g.gbuffer(p)
            g.cone_trace(p)
            r, n = self.rank, self.band_px
            self.dist.all_gather_into_tensor(self.image, self.image[r * n:(r + 1) * n])

    def step_e2e(self, host_img, pipelined=False):
        """host_img: pinned int32 tensor of W*H pixels.  Frame parameters go host->device inside vct_gi_passes.
        pipelined (one GPU): the read-back is enqueued on the library's copy stream (vct_read_image_async) and overlaps the next
        step's voxel passes; the caller alternates two host buffers and ends the loop with finish_e2e()."""
        self.step()
        g = self.g
        if self.world == 1:
            if pipelined:
                g.read_image_async(host_img.data_ptr())
            else:
                g._ck(g.lib.vct_read_image(g.h, host_img.data_ptr()))
        elif self.rank == 0:
            with self.torch.cuda.stream(self.stream):
                host_img.copy_(self.image[: g.W * g.H], non_blocking=True)
            self.stream.synchronize()

    def finish_e2e(self):
        """Order the library stream behind the last pipelined read-back (so that an event recorded next covers it)."""
        if self.world == 1:
            self.g.read_image_wait(block_host=False)

    def profiled_step(self):
        """Per-kernel times {name: (ns, launches)} of one step (library profiling level 2)."""
        g, p = self.g, self.p
        self._animate()
        if self.world == 1:
            (g.frame if self.whole_frame else g.gi_passes)(p)
            return g.kernel_times()
        with self.torch.cuda.stream(self.stream):
            g.gi_passes(p)
            kt = g.kernel_times()
            self._exchange(); kt2 = g.kernel_times()          # (kernels of the last library call of the exchange)
            g.cone_trace(p); kt3 = g.kernel_times()
            for extra in (kt2, kt3):
                for k, (ns, n) in extra.items():
                    a = kt.get(k, (0.0, 0)); kt[k] = (a[0] + ns, a[1] + n)
            r, n = self.rank, self.band_px
            self.dist.all_gather_into_tensor(self.image, self.image[r * n:(r + 1) * n])
        return kt

    def close(self):
        """Drop the captured graph before the process group goes away (a graph that outlives NCCL's communicator hangs teardown)."""
        self.graph = None

    def pass_times(self):
        """Per-pass ms of one whole reference frame graph (GLTimer semantics; profiling level 1), incl. producers."""
        g, p = self.g, self.p
        if self.world > 1:
            return {}
        g.frame(p)
        return {k.replace("_ns", ""): round(v / 1e6, 4) for k, v in g.timings().items() if v > 0}
