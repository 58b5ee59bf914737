#!/usr/bin/env python3
"""Multi-GPU parity (run under torchrun, one rank per GPU): the z-slab sharded frame (exchange inside the library, csrc/exchange.cu)
must give exactly the words of the single-GPU frame — every level of the traced pyramid on rank 0 after the exchange, the
voxelise counters summed over ranks, and the final image assembled in rank 0's buffer from every rank's screen tiles.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           tools/sharded_parity.py [room|room_msaa|sponza|sponza512|animated]     (animated: config 4 at 256^3 / 960x540, whole frames, moving actors)
Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vct_b200 import params as P  # noqa: E402
from vct_b200 import scene as S  # noqa: E402
from vct_b200.pipeline import Pipeline  # noqa: E402
from vct_b200.sharded import ShardedFrame  # noqa: E402


def workload(name):
    """-> (scene, params, D, L, S, W, H, Workload or None)"""
    if name in ("room", "room_msaa"):
        p = S.room_params(320, 240)
        if name == "room_msaa":                  # multisample voxelisation at twice the viewport: extrapolated fragments, no triangle cull per rank
            p.conservative_raster = P.RASTER_MSAA; p.voxelize_multiplier = 2.0
        return S.room_scene(), p, 64, 5, 512, 320, 240, None
    from vct_b200.workloads import Workload
    w = Workload(3) if name == "sponza" else Workload(3, 3840, 2160, 512) if name == "sponza512" else Workload(4, 960, 540, 256)
    return w.scene, w.params, w.D, w.L, w.S, w.W, w.H, (w if name == "animated" else None)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "room"
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc, p, D, L, SS, W, H, wl = workload(name)
    frames = 4 if wl else 3
    g = Pipeline(sc, D, L, SS, W, H, device=local, rank=rank, world_size=world, slab_stripe=-1 if os.environ.get("VCT_SPARSE_EXCHANGE", "1") == "0" else 0)
    fr = ShardedFrame(g, p, world, rank, workload=wl)
    fr.producers()
    for _ in range(frames):                  # later frames: steady state (sparse exchange, publish masks, temporal history) also matches
        fr.step()
    torch.cuda.synchronize()
    info = g.counters()
    cnt = torch.tensor([info.total_fragments, info.unique_voxels], device="cuda", dtype=torch.int64)
    mx = torch.tensor([info.max_fragments_per_voxel], device="cuda", dtype=torch.int64)
    dist.all_reduce(cnt); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    ok, report = True, {"workload": name, "world": world, "exchange": fr.describe()[:100]}
    if rank == 0:
        which = P.VOL_RADIANCE if p.draw_radiance else P.VOL_COLOR
        sharded_levels = [g.read_volume(which, l) for l in range(g.L)]
        sharded_img = g.read_image()
        one = Pipeline(sc, D, L, SS, W, H, device=local)
        try:
            one.shadowmap(p)
            if p.warp_texture:
                one.occupancy(p); one.warpmap(p)
            one.gbuffer(p)
            for k in range(frames):
                if wl:
                    for actor, model in wl.models(k):
                        one.set_actor_transform(actor, model)
                    one.frame(p)
                else:
                    one.gi_passes(p)
            ref_info = one.counters()
            for l in range(one.L):
                same = np.array_equal(sharded_levels[l], one.read_volume(which, l))
                report[f"level{l}_equal"] = bool(same); ok &= same
            img = one.read_image()
            diff = int((img != sharded_img).sum())
            report["image_pixels_differing"] = diff; ok &= diff == 0
            report["counters_sharded"] = [int(cnt[0]), int(cnt[1]), int(mx[0])]
            report["counters_single"] = [ref_info.total_fragments, ref_info.unique_voxels, ref_info.max_fragments_per_voxel]
            ok &= report["counters_sharded"] == report["counters_single"]
        finally:
            one.close()
        report["ok"] = bool(ok)
        print(json.dumps(report), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    g.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
