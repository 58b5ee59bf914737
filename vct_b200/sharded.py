"""One frame of the GI hot path on 1..N GPUs, one process per GPU — a THIN caller: the library runs the
whole sharded frame itself.

world_size == 1   the frame is `vct_gi_passes` (or `vct_frame`) on the library's own stream.
world_size  > 1   once, at start-up, the ranks swap their cudaIpc handle blobs (`vct_exchange_export` / `_import`; the only use of
                  torch.distributed here, `all_gather_object`).  After that `vct_gi_passes` / `vct_frame` does everything on the
                  device: voxel passes on the own z layers (stripes of `slab_stripe` layers dealt round-robin to the ranks; default:
                  one contiguous slab [rank*D/N, (rank+1)*D/N)), the slab exchange over NVLink peer memory
                  with device-side flags (csrc/exchange.cu), the cone trace of the own 64x64 screen tiles, pixels stored into
                  rank 0's image.  No collective, no host synchronisation, and the step is capturable in a CUDA graph.
                  VCT_SPARSE_EXCHANGE=0 (or a backend other than NCCL) selects the round-1 protocol instead: per-level
                  all-gather by the caller, `vct_exchange`, full-image trace on every rank.

The partition maths (`slab_range`, `stripes`, `top_sharded_level`, `level_chunks`, `tile_owner`, `image_bands`) is pure Python, mirrors the library's, and is
unit-tested on CPU with gloo.
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


def stripe_owner(z, stripe, world):
    """Rank that owns voxel layer z when the layers are dealt out in stripes of `stripe` layers (common.cuh `owns_z`)."""
    return (z // stripe) % world


def stripes(dim, stripe, world, rank):
    """[(z_lo, z_hi)] of the stripes `rank` owns; stripe = dim // world is the contiguous slab of `slab_range`."""
    if stripe < 1 or dim % (stripe * world):
        raise ValueError("stripe * world_size must divide dim")
    return [((k * world + rank) * stripe, (k * world + rank + 1) * stripe) for k in range(dim // (stripe * world))]


def top_sharded_level(stripe, levels):
    """Last mip level a rank can filter from its own stripes (texel layers do not straddle a stripe border); the levels above it
    are filtered on every rank from the exchanged level (volume_passes.cu vctk_mip_top_sharded_level / vctk_mip_tail)."""
    l = 0
    while l + 1 < levels and stripe % (1 << (l + 1)) == 0:
        l += 1
    return l


def image_bands(height, world):
    """Equal bands of whole 8-row tiles; returns (band_rows, [(y_lo, y_hi) per rank]) — mirrors vctk_image_rows."""
    rows8 = (height + 7) // 8
    band = (rows8 + world - 1) // world * 8
    return band, [(min(height, r * band), min(height, (r + 1) * band)) for r in range(world)]


SCREEN_TILE = 64


def tile_owner(tx, ty, world):
    """Rank that traces screen tile (tx, ty) of 64x64 pixels — diagonal stripes (cone_trace.cu ensure_trace_tiles)."""
    return (tx + ty) % world


def screen_tiles(width, height, world, rank):
    """[(x0, y0)] of the tiles `rank` traces, in the order the library walks them."""
    ntx, nty = (width + SCREEN_TILE - 1) // SCREEN_TILE, (height + SCREEN_TILE - 1) // SCREEN_TILE
    return [(tx * SCREEN_TILE, ty * SCREEN_TILE) for ty in range(nty) for tx in range(ntx) if tile_owner(tx, ty, world) == rank]


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
        self.graph = None
        g = pipeline
        self.peer_exchange = False
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            # a dedicated (non-default) stream: torch's default stream has handle 0, which vct_set_stream reads as "private stream"
            self.stream = torch.cuda.Stream()
            self.stream.wait_stream(torch.cuda.current_stream())
            g.set_stream(self.stream.cuda_stream)
            if os.environ.get("VCT_SPARSE_EXCHANGE", "1") != "0" and dist.get_backend() == "nccl":
                handles = [None] * world
                dist.all_gather_object(handles, g.exchange_setup())
                for r, h in enumerate(handles):
                    g.exchange_import(r, h)
                dist.barrier()                                   # every rank has mapped every peer before the first frame stores into them
                self.peer_exchange = True
            else:                                                # round-1 protocol: the caller moves the levels
                if g.slab_stripe() != g.D // world:
                    raise ValueError("the caller-side all-gather protocol needs contiguous slabs (Pipeline(..., slab_stripe=-1))")
                which = P.VOL_RADIANCE if params.draw_radiance else P.VOL_COLOR
                self.levels = [device_tensor(g.device_ptr(which, l), g.level_bytes(which, l)) for l in range(g.L)]
                self.chunks = level_chunks(g.D, g.L, world)
        else:
            self.stream = torch.cuda.ExternalStream(g.lib.vct_stream(g.h))

    def describe(self):
        if self.world == 1:
            return "1 GPU"
        if self.peer_exchange:
            T = self.g.slab_stripe()
            return (f"{self.world} GPUs, one process each: z sharding of clear/voxelise/transfer/inject/mip in stripes of {T} layers; slab exchange inside the library over NVLink peer memory "
                    "(level 0 as flagged x-row segments into staging regions, levels >= 1 stored into the peers' pyramids, device-side flags, no collective); "
                    "cone trace sharded by interleaved 64x64 screen tiles, pixels stored into rank 0's image")
        return f"{self.world} GPUs: z-slab sharding, per-level NCCL all-gather of the traced pyramid by the caller, full-image cone trace on every rank"

    # producers of the reference frame graph that the GI step consumes (replicated on every rank)
    def producers(self, gbuffer=True):
        g, p = self.g, self.p
        g.shadowmap(p)
        if p.warp_texture:
            g.occupancy(p); g.warpmap(p)
        if gbuffer:
            g.gbuffer(p)

    # ---- CUDA graph: the step is ~15 short launches; replaying a captured pair of steps removes the host from the loop.  TWO steps
    # per graph because the segment masks swap roles every frame; the graph is only replayed at the mask parity it was captured at,
    # and only while the frame parameters stay what they were.  N > 1: the sharded step holds kernels only (flags, no collectives).
    def enable_graph(self):
        torch = self.torch
        self.graph, self.graph_error = None, None
        if self.world > 1 and not self.peer_exchange:
            self.graph_error = "caller-side all-gather protocol"
            return False
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
            if self.graph is not None and k >= 2 and self.g.mask_parity() == self.graph_parity:
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
        if self.world == 1 or self.peer_exchange:
            (g.frame if self.whole_frame else g.gi_passes)(p)    # N > 1: the whole sharded frame, exchange and image hand-over included
            return
        with self.torch.cuda.stream(self.stream):                # round-1 protocol
            if self.whole_frame:
                self.producers(gbuffer=False)
            g.gi_passes(p)
            all_gather_levels(self.dist, self.levels, self.chunks, self.rank)
            g.exchange()
            if self.whole_frame:
                g.gbuffer(p)
            g.cone_trace(p)

    def step_e2e(self, host_img, pipelined=False):
        """host_img: pinned int32 tensor of W*H pixels.  Frame parameters go host->device inside vct_gi_passes.
        pipelined (one GPU): the read-back is enqueued on the library's copy stream (vct_read_image_async) and overlaps the next
        step's voxel passes; the caller alternates two host buffers and ends the loop with finish_e2e().
        N > 1: rank 0 owns the finished image and reads it back on the frame's stream."""
        self.step()
        g = self.g
        if self.world == 1:
            if pipelined:
                g.read_image_async(host_img.data_ptr())
            else:
                g._ck(g.lib.vct_read_image(g.h, host_img.data_ptr()))
        elif self.rank == 0:
            if pipelined and self.peer_exchange:                 # same pipelining on rank 0: the next frame's exchange and trace wait for the copy
                g.read_image_async(host_img.data_ptr())
            else:
                g._ck(g.lib.vct_read_image(g.h, host_img.data_ptr()))

    def finish_e2e(self):
        """Order the library stream behind the last pipelined read-back (so that an event recorded next covers it)."""
        if self.world == 1 or (self.rank == 0 and self.peer_exchange):
            self.g.read_image_wait(block_host=False)

    def profiled_step(self):
        """Per-kernel times {name: (ns, launches)} of one step (library profiling level 2)."""
        g, p = self.g, self.p
        self.step()
        return g.kernel_times()

    def close(self):
        """Drop the captured graph before the context and the process group go away."""
        self.graph = None

    def pass_times(self):
        """Per-pass ms of one step (GLTimer semantics; profiling level 1).  One GPU: of one whole reference frame graph incl. producers."""
        g, p = self.g, self.p
        if self.world > 1:
            self.step()
        else:
            g.frame(p)
        return {k.replace("_ns", ""): round(v / 1e6, 4) for k, v in g.timings().items() if v > 0}
