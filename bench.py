#!/usr/bin/env python3
"""bench.py — GI frame time of the B200 pipeline on the BASELINE.json configurations (default: config 3 = Sponza, 256^3 voxels,
1920x1080, the configuration the metric is quoted on; `--config 1..5` selects the others, vct_b200/workloads.py).

One "step" = one frame of the GI hot path (BASELINE metric): clear + voxelise + transferVoxels + injectRadiance +
mip chain(s) + per-pixel cone trace, everything recomputed every frame.  The shadow map and the visibility buffer
(producer passes of the reference's render(), SURVEY.md §8 a0/a0') are inputs: they are generated on the device
during warm-up and their cost is reported separately in `passes_ms`.

  value      device-timed ms/frame (CUDA events on the library's stream), inputs resident in HBM
  e2e        the same frame through the C ABI with HOST buffers: frame parameters + lights + actor transforms go
             host->device and the final RGBA8 image comes back into pinned host memory, every step
  roofline   the dominant kernel (cone trace) against ITS bound, the L1/texture data pipe; `roofline_hbm` = the voxel passes against
             the measured HBM copy peak (dense algorithmic bytes AND the DRAM bytes ncu measured); `roofline_passes` per kernel
  cpu_baseline / --impl reference   the reference's own GLSL shaders compiled as C++ for the host (oracle/_ref/libvct_glsl_ref.so,
             built from /root/reference/shaders by oracle/Makefile; kind "reference") on the box's host cores, with this
             repository's canonical OpenGL fixed function around them (no OpenGL stack exists on the image); if that
             library did not travel with the repository, the CPU oracle port (kind "port")

N > 1 (torchrun, one rank per GPU): the volume is sharded by z-slab for clear/voxelise/transfer/inject/mip, the
radiance pyramid is all-gathered over NVLink (NCCL), the cone trace is sharded by screen band, and rank 0 gathers
the image bands.  Same frame, so scaling is "strong".
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GI frame ms (voxelize+inject+mip+cone trace) Sponza 256^3 @1080p"      # BASELINE.json; other configs name theirs below
LEVELS, SHADOW = 6, 4096


def metric_name(w):
    if w.config == 3 and (w.D, w.W, w.H) == (256, 1920, 1080):
        return METRIC
    what = "whole frame ms (shadow map+visibility+voxelize+inject+mip+cone trace)" if w.whole_frame else "GI frame ms (voxelize+inject+mip+cone trace)"
    return f"{what} config {w.config} {w.D}^3 @{w.W}x{w.H}"


def peaks():
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(m["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(width=None, height=None, dim=None):
    """(scene, params, D, W, H, data) of config 3 — the tuple the tools and tests of round 1 use."""
    w = make_workload(3, width, height, dim)
    return w.scene, w.params, w.D, w.W, w.H, w.data


def make_workload(config=3, width=None, height=None, dim=None, triangles=None):
    from vct_b200.workloads import Workload
    return Workload(config, width, height, dim, triangles)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(D, L, S, W, H, T, F, U, chains):
    """SURVEY.md §8(d) / BASELINE.md §4 per-pass byte counts, attributed to this build's kernels."""
    d3 = D ** 3
    mip = sum(4 * (D >> l) ** 3 * (1 + 1 / 8) for l in range(L - 1))
    pyr = sum(4 * (D >> l) ** 3 for l in range(L))
    return {
        "k_clear": 8 * d3,                                   # voxelColor + voxelNormal level 0
        "voxelize": 96 * T + 16 * F,                         # all voxeliser kernels together
        "k_transfer": 4 * d3 + 4 * d3 + 8 * U,               # colour read + (fused) radiance clear + 2 stores per occupied voxel
        "k_inject": 4 * S * S,                               # + 8 N_in scattered (not counted: lower bound)
        "k_mip_chain": mip * chains,                         # (the fused kernel also feeds the texture array: not counted)
        "k_cone_trace": pyr + 8 * W * H + 4 * W * H,         # compulsory HBM: pyramid once + visibility + image
    }


VOXELIZE_KERNELS = ("k_transform_vertices", "k_voxel_reset", "k_voxel_bin", "k_voxel_tiles", "k_voxel_resolve", "k_voxel_resolve_medium", "k_voxel_huge_compact", "k_voxel_huge_pick", "k_voxel_huge_gather", "k_voxel_huge_select",
                    "k_voxel_bin_cas", "k_voxel_tiles_cas", "k_voxel_bin_max", "k_voxel_tiles_max")


def measured_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel from the committed
    `ncu --set full` capture of one frame of this same workload (profiles/traffic.json, written by tools/ncu_traffic.py)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_frame(o, r, w, frame_index, rows_stride=1):
    """One step of workload `w` on the host cores; returns per-pass seconds.  r = GlslReference (the reference's own shaders
    compiled as C++) or None (the oracle port).  The fixed-function producers (shadow map, occupancy + warp map, visibility) are
    always the oracle's — only timed when the workload's step contains them (config 4).  The cone trace may run on every
    rows_stride-th image row only; its time is scaled back to the full image."""
    p = w.params
    for actor, model in w.models(frame_index):
        o.set_actor_transform(actor, model)
    x = r if r is not None else o
    passes = []
    if w.whole_frame:
        passes.append(("shadowmap", lambda: o.shadowmap(p)))
        if p.warp_texture:
            passes.append(("warpmap", lambda: (o.occupancy(p), o.warpmap_pass(p))))
    passes += [("voxelize", lambda: x.voxelize(p)), ("transfer", lambda: x.transfer(p)), ("inject", lambda: x.inject(p)),
               ("mip", lambda: (x.mip("radiance"), x.mip("color") if p.mip_color_chain else None))]
    if w.whole_frame:
        passes.append(("gbuffer", lambda: o.visibility(p)))
    passes.append(("cone_trace", lambda: x.shade(p, 0, None, rows_stride)))
    t = {}
    for name, fn in passes:
        t0 = time.perf_counter(); fn(); t[name] = time.perf_counter() - t0
    t["cone_trace"] *= rows_stride
    return t


def oracle_gi_frame(o, p, rows_stride=1):
    """The five GI passes of the metric on the CPU oracle (kept for tests/test_glsl_ref.py)."""
    t = {}
    for name, fn in (("voxelize", lambda: o.voxelize(p)), ("transfer", lambda: o.transfer(p)), ("inject", lambda: o.inject(p)),
                     ("mip", lambda: (o.mip("radiance"), o.mip("color") if p.mip_color_chain else None)),
                     ("cone_trace", lambda: o.shade(p, 0, None, rows_stride))):
        t0 = time.perf_counter(); fn(); t[name] = time.perf_counter() - t0
    t["cone_trace"] *= rows_stride
    return t


def reference_gi_frame(r, p, rows_stride=1):
    """The same five passes with the reference's own GLSL compiled as C++ (tests/oracle_lib.GlslReference)."""
    t = {}
    for name, fn in (("voxelize", lambda: r.voxelize(p)), ("transfer", lambda: r.transfer(p)), ("inject", lambda: r.inject(p)),
                     ("mip", lambda: (r.mip("radiance"), r.mip("color") if p.mip_color_chain else None)),
                     ("cone_trace", lambda: r.shade(p, 0, None, rows_stride))):
        t0 = time.perf_counter(); fn(); t[name] = time.perf_counter() - t0
    t["cone_trace"] *= rows_stride
    return t


REF_NOTE = ("the reference's own GLSL (voxelize.frag, transferVoxels/injectRadiance/filterRadiance.comp, phong.frag) compiled as C++ from "
            "/root/reference/shaders (oracle/_ref/libvct_glsl_ref.so), OpenMP over invocations (voxelize.frag on one thread, canonical order); "
            "rasterisation / interpolation / texture filtering = this repository's canonical OpenGL semantics (no OpenGL/EGL/Mesa on this image)")


def cpu_arm(o):
    """(GlslReference or None, kind, note): the reference's shaders compiled for the host when oracle/_ref/libvct_glsl_ref.so
    travelled with the repository (kind "reference"), else the oracle port."""
    try:
        from tests.oracle_lib import GlslReference
        return GlslReference(o), "reference", REF_NOTE
    except Exception as e:                                          # library not built (no /root/reference at build time)
        return None, "port", f"CPU oracle (C++/OpenMP restatement of the reference GLSL); compiled reference shaders unavailable: {type(e).__name__}"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on ALL host threads.  Rank 0 only; the process never loads
    libvct_b200.so (scenes come from the numpy texture cache of tools/bake_assets.py)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)                  # torchrun hands its workers OMP_NUM_THREADS=1; set before libgomp loads
    from tests.oracle_lib import Oracle, lib
    w = make_workload(args.config, args.width, args.height, args.dim, args.triangles)
    p = w.params
    o = Oracle(w.scene, w.D, w.L, w.S, w.W, w.H)
    cores = lib().orc_num_threads()
    o.shadowmap(p)
    if p.warp_texture:
        o.occupancy(p); o.warpmap_pass(p)
    o.visibility(p)                                              # producers: inputs of the step (timed inside it for whole-frame workloads)
    r, kind, note = cpu_arm(o)
    first = cpu_frame(o, r, w, 0, 16)                            # untimed probe (cone trace on every 16th row) = warm-up 0
    budget = 150.0 / max(1, args.steps + args.warmup)
    stride = 1
    other = sum(v for k, v in first.items() if k != "cone_trace")
    while other + first["cone_trace"] / stride > budget and stride < 64:
        stride *= 2
    for i in range(max(0, args.warmup - 1)):
        cpu_frame(o, r, w, 1 + i, stride)
    per, tot = [], 0.0
    for i in range(args.steps):
        t = cpu_frame(o, r, w, args.warmup + i, stride); per.append(t); tot += sum(t.values())
    ms = 1e3 * tot / args.steps
    sample = (f"every pass but the cone trace at full size; cone trace on every {stride}th image row, time x{stride}"
              if stride > 1 else "full step, every pass at full size")
    passes = {k: round(1e3 * statistics.mean(t[k] for t in per), 3) for k in per[0]}
    port = None
    if kind == "reference":                                      # for transparency: this repository's hand-written port of the same passes, one step
        tp = cpu_frame(o, None, w, args.warmup + args.steps, stride)
        port = {"value": round(1e3 * sum(tp.values()), 3), "unit": "ms", "passes_ms": {k: round(1e3 * v, 3) for k, v in tp.items()},
                "note": "CPU oracle (C++/OpenMP restatement, bit-identical results), same sample, single run"}
    line = {"impl": "reference", "metric": metric_name(w), "value": round(ms, 3), "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 3), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": w.data,
            "config": w.config_dict(),
            "cpu_baseline": {"value": round(ms, 3), "unit": "ms", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(ms, 3), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "passes_ms": passes, "cpu_port": port, "gpu_launches": 0,
            "native_so": "oracle only: libvct_b200.so is not loaded by this arm",
            "note": note}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from vct_b200 import params as P
    from vct_b200.pipeline import Pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = make_workload(args.config, args.width, args.height, args.dim, args.triangles)
    sc, p, D, W, H, L, S = w.scene, w.params, w.D, w.W, w.H, w.L, w.S
    chains = w.chains
    max_frag = max(8 << 20, 3 * sc.n_tris) if w.config == 5 else 0
    g = Pipeline(sc, D, L, S, W, H, device=local, rank=rank, world_size=world, max_fragments=max_frag,
                 slab_stripe=-1 if os.environ.get("VCT_SPARSE_EXCHANGE", "1") == "0" else 0)   # the caller-side all-gather protocol needs contiguous slabs
    from vct_b200.sharded import ShardedFrame
    fr = ShardedFrame(g, p, world, rank, workload=w)              # world == 1: plain vct_gi_passes / vct_frame on the library stream
    stream = fr.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # producers + warm-up (whole reference frame graph, so every buffer the step reads exists)
    fr.producers()
    for _ in range(max(3, args.warmup)):
        fr.step()
    barrier()

    # ---- timed region: exactly K steps, device events on the launching stream, max over ranks
    g.set_profiling(0)
    g.launch_count(reset=True)
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # The timed steps replay a CUDA graph of two captured steps (two: the segment masks swap roles every frame).  Animated workloads
    # change kernel parameters every frame (actor matrices travel by value) and stay host-launched.
    use_graph = (not args.no_graph) and not w.animated and fr.enable_graph()
    g.launch_count(reset=True)
    barrier()
    e0.record(stream)
    replayed = fr.run_steps(args.steps)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = g.launch_count() + (replayed * fr.graph_launches_per_step if replayed else 0)
    clk = clocks.stop() if clocks else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64); dist.all_reduce(lt); launches = int(lt.item())
    ms = ms_total / args.steps

    # ---- e2e: host buffers in, image out, every step
    # One GPU: the read-back of frame i is pipelined behind frame i+1's voxel passes (vct_read_image_async, two pinned host
    # buffers in turn; every step still copies its parameters in and its whole image out, and the timed region ends only
    # after the last image has landed).  --e2e-blocking reads every image back synchronously instead.
    pipelined = (world == 1 or fr.peer_exchange) and not args.e2e_blocking
    host_imgs = [torch.empty(W * H, dtype=torch.int32).pin_memory() for _ in range(2)]
    h2d = C.sizeof(P.FrameParams) + 80 * len(sc.lights) + (64 + 36) * len(sc.meshes)
    d2h = W * H * 4
    for i in range(3):                                           # warm the copy stream / pinned pages
        fr.step_e2e(host_imgs[i & 1], pipelined)
    fr.finish_e2e()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        fr.step_e2e(host_imgs[i & 1], pipelined)
    fr.finish_e2e()
    e1.record(stream)
    barrier()
    e2e_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_total], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_total = float(t.item())
    e2e_ms = e2e_total / args.steps

    # ---- per-kernel breakdown: profiled frames (one event per kernel) right after the timed region, on every rank (rank 0 reports its own)
    kern, passes = {}, {}
    nprof = args.profile_frames
    g.set_profiling(2)
    for _ in range(nprof):
        for name, (ns, n) in fr.profiled_step().items():
            a = kern.setdefault(name, [0.0, 0]); a[0] += ns / 1e6 / nprof; a[1] += n / nprof
    g.set_profiling(1)
    passes = fr.pass_times()
    info = g.counters()
    steps_cone = g.cone_steps()
    g.set_profiling(0)
    if world > 1:                                                # whole-job counters: fragments / voxels / cone steps are per rank (own slab, own band)
        tot = torch.tensor([info.total_fragments, info.unique_voxels, steps_cone], device="cuda", dtype=torch.int64); dist.all_reduce(tot)
        mx = torch.tensor([info.max_fragments_per_voxel], device="cuda", dtype=torch.int64); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        job_frag, job_vox, job_steps, job_max = int(tot[0]), int(tot[1]), int(tot[2]), int(mx[0])
    else:
        job_frag, job_vox, job_steps, job_max = info.total_fragments, info.unique_voxels, steps_cone, info.max_fragments_per_voxel

    if rank == 0:
        peak, peak_src = peaks()
        T = sc.n_tris
        # per-rank algorithmic bytes: a rank's kernels process its own z-slab / screen band, i.e. 1/N of the dense semantics
        ab = {k: v / world for k, v in algorithmic_bytes(D, L, S, W, H, T, job_frag, job_vox, chains).items()}
        kt = {k: v[0] for k, v in kern.items()}
        kt["voxelize"] = sum(kt.get(k, 0.0) for k in VOXELIZE_KERNELS)
        kt["k_inject"] = kt.get("k_inject", 0.0) + kt.get("k_inject_cull", 0.0)
        if "k_transfer" not in kt and "k_voxel_resolve" in kt:
            kt["k_transfer"] = 0.0                                   # sparse frame: transferVoxels runs inside k_voxel_resolve (counted under "voxelize")
        mt = measured_traffic() if (w.config == 3 and world == 1 and (D, W, H) == (256, 1920, 1080)) else {}
        mt["voxelize"] = sum(mt.get(k, 0) for k in ("k_transform_vertices", "k_voxel_bin", "k_voxel_expand", "k_voxel_tiles", "k_voxel_resolve")) or None
        roofs = []
        for name, nbytes in ab.items():
            t_ms = kt.get(name, 0.0)
            if name == "voxelize" and kt.get("k_transfer") == 0.0:
                nbytes += ab["k_transfer"]                           # the fused kernel does both jobs
            if t_ms <= 0:
                continue
            ach = nbytes / (t_ms * 1e-3) / 1e9
            r = {"kernel": name, "ms": round(t_ms, 4), "dense_algorithmic_bytes": int(nbytes),
                 "dense_equivalent_gbs": round(ach, 1), "dense_equivalent_frac": round(ach / peak, 4), "peak": peak, "unit": "GB/s",
                 "note": "dense-equivalent: SURVEY 8(d) bytes of the reference's DENSE sweep over this kernel's time; the sparse kernels move far fewer bytes, so this can exceed 1 and is not a bandwidth claim"}
            if mt.get(name):
                r["measured_dram_bytes"] = int(mt[name]); r["measured_gbs"] = round(mt[name] / (t_ms * 1e-3) / 1e9, 1); r["measured_frac"] = round(mt[name] / (t_ms * 1e-3) / 1e9 / peak, 4)
            roofs.append(r)
        step_kernel_ms = sum(v[0] for k, v in kern.items() if not k.startswith(("memset", "h2d", "<")))
        voxel_names = ("k_clear", "voxelize", "k_transfer", "k_inject", "k_mip_chain")
        voxel_bytes = sum(ab[k] for k in voxel_names)
        voxel_ms = sum(kt.get(k, 0.0) for k in voxel_names)
        voxel_meas = sum(mt.get(k) or 0 for k in voxel_names)
        roofline_hbm = {"what": "voxel passes (clear + voxelise + transfer + inject + mip chains), per rank", "bound": "hbm", "ms": round(voxel_ms, 4),
                        "achieved": round(voxel_bytes / max(voxel_ms, 1e-9) / 1e6, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(voxel_bytes / max(voxel_ms, 1e-9) / 1e6 / peak, 4), "algorithmic_bytes": int(voxel_bytes),
                        "traffic": int(voxel_meas) if voxel_meas else None,
                        "measured_frac": round(voxel_meas / max(voxel_ms, 1e-9) / 1e6 / peak, 4) if voxel_meas else None, "peak_source": peak_src,
                        "note": "achieved = dense algorithmic bytes (SURVEY 8d, / n_gpus) over the summed kernel time; traffic = DRAM bytes of the same kernels in the committed ncu capture"}
        roofline = None
        t_trace = kt.get("k_cone_trace", 0.0)
        if t_trace > 0:
            mhz = float((clk or {}).get("sm_mhz") or 1965.0)
            full = measured_traffic()
            roofline = {"kernel": "k_cone_trace", "bound": "l1tex", "ms": round(t_trace, 4), "share_of_step": round(t_trace / max(step_kernel_ms, 1e-9), 3),
                        "timing": f"per-kernel CUDA events on the library stream, mean of {nprof} profiled frames run right after the timed region",
                        "cone_steps": steps_cone, "traffic": mt.get("k_cone_trace"), "peak_source": "1 TEX data-pipe wavefront / clock / SM x 148 SMs at the SM clock sampled during the timed region"}
            try:                                                   # TEX data-pipe wavefronts per cone step come from the committed ncu capture of this fetch mix
                wf = full["k_cone_trace__tex_wavefronts"] / full["k_cone_trace__cone_steps"] * steps_cone
                peak_wf = 148 * mhz * 1e6 / 1e9
                roofline.update({"achieved": round(wf / (t_trace * 1e-3) / 1e9, 2), "peak": round(peak_wf, 2), "unit": "Gwavefront/s",
                                 "frac": round(wf / (t_trace * 1e-3) / 1e9 / peak_wf, 4), "wavefronts": int(wf), "sm_mhz": mhz,
                                 "floor_ms": round(wf / (148 * mhz * 1e6) * 1e3, 4),
                                 "source": "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum per cone step (profiles/traffic.json) x this run's cone steps"})
            except Exception:
                roofline.update({"achieved": None, "peak": None, "unit": "Gwavefront/s", "frac": None})
            roofline["hbm_compulsory"] = {"bytes": int(ab["k_cone_trace"]), "gbs": round(ab["k_cone_trace"] / (t_trace * 1e-3) / 1e9, 1),
                                          "frac_of_hbm_peak": round(ab["k_cone_trace"] / (t_trace * 1e-3) / 1e9 / peak, 4)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from tests.oracle_lib import Oracle, lib
            o = Oracle(sc, D, L, S, W, H)
            o.shadowmap(p)
            if p.warp_texture:
                o.occupancy(p); o.warpmap_pass(p)
            o.visibility(p)
            r, kind, note = cpu_arm(o)
            probe = cpu_frame(o, r, w, 0, 16)                          # warm the CPU caches / page in the volumes (cone trace on every 16th row)
            stride = 1
            while sum(v for k, v in probe.items() if k != "cone_trace") + probe["cone_trace"] / stride > 25.0 and stride < 64:
                stride *= 2
            t = cpu_frame(o, r, w, 1, stride)
            cpu = {"value": round(1e3 * sum(t.values()), 1), "unit": "ms", "cores": lib().orc_num_threads(), "kind": kind,
                   "sample": ("one full step at full size, single run after one short warm-up" if stride == 1 else
                              f"one step: every pass but the cone trace at full size, cone trace on every {stride}th image row, time x{stride}"),
                   "passes_ms": {k: round(1e3 * v, 1) for k, v in t.items()}, "note": note}
            if kind == "reference":
                tp = cpu_frame(o, None, w, 2, stride)
                cpu["port_ms"] = round(1e3 * sum(tp.values()), 1)       # the hand-written C++/OpenMP port of the same passes (bit-identical results)
        cfg = w.config_dict()                                       # identical in both arms (the driver compares them)
        execution = {"parallelism": fr.describe(),
                     "sparse_frames": os.environ.get("VCT_SPARSE", "1") != "0",
                     "cuda_graph": (f"{replayed} of {args.steps} timed steps replayed from a captured pair of steps" if use_graph else
                                    f"off ({getattr(fr, 'graph_error', None) or ('animated workload: actor matrices change every frame' if w.animated else 'disabled')})")}
        line = {"metric": metric_name(w), "value": round(ms, 4), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": round(ms, 4), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": w.data,
                "config": cfg, "execution": execution,
                "e2e": {"value": round(e2e_ms, 4), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "readback": "rank 0, pipelined on a copy stream behind the next step's voxel passes (vct_read_image_async), 2 pinned host buffers" if pipelined
                                    else "synchronous after every step"},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_passes": roofs,
                "kernels_ms": {k: round(v[0], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])},
                "kernels_note": "rank 0's kernels, mean of profiled host-launched frames after the timed region: shares of the step, not absolutes (their sum exceeds ms_per_step)",
                "passes_ms": passes, "cone_steps": job_steps, "cone_steps_per_s": round(steps_cone / max(t_trace, 1e-9) * 1e3, 0),
                "counters": {"total_fragments": job_frag, "unique_voxels": job_vox, "max_fragments_per_voxel": job_max},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    fr.close()
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (SURVEY 8d numbering); 3 = the metric's")
    ap.add_argument("--triangles", type=int, default=None, help="config 5: triangle count of the synthetic soup (default 1 Mi)")
    ap.add_argument("--profile-frames", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the timed steps from the host instead of replaying a CUDA graph")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--e2e-blocking", action="store_true", help="e2e: read every image back synchronously (no copy/compute overlap)")
    ap.add_argument("--dim", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 20:
            args.steps, args.warmup = 10, 3
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
