// cone_trace_debug.cu — the debug views of phong.frag (vct_frame_params::debug_view != 0; phong.frag:346-447, 489-505) as DBG
// instantiations of the kernel in cone_trace.cuh, compiled WITHOUT fast math and with -fmad=false like the other "exact" units:
// every float operation is one IEEE rounding in source order, the same sequence as the oracle's.  The voxel view at lod <= 0.5 is a
// NEAREST fetch (one voxel index per pixel): with the frustum-aligned grid of voxelizeTesselationWarp, image rows whose centre maps
// exactly onto a voxel boundary ((py + 0.5) * D / H integral — rows 7, 22, 37, ... at D = 64, H = 240) decide their voxel by the last
// ulp, and the fast-math build of the shaded frame differed from the oracle there (29.4 dB; profiles/r02a_diag_voxel_view.txt).
#include "cone_trace.cuh"

int vctk_cone_trace_debug(vct_ctx* c, const TraceArgs& a, int wm, dim3 grid) {
    if (wm == WARP_VOXELS) k_cone_trace<WARP_VOXELS, true><<<grid, kThreads, 0, c->stream>>>(a);
    else if (wm == WARP_TEXTURE) k_cone_trace<WARP_TEXTURE, true><<<grid, kThreads, 0, c->stream>>>(a);
    else if (wm == WARP_TESS) k_cone_trace<WARP_TESS, true><<<grid, kThreads, 0, c->stream>>>(a);
    else k_cone_trace<WARP_NONE, true><<<grid, kThreads, 0, c->stream>>>(a);
    VCT_LAUNCH_CHECK(c, "k_cone_trace_debug");
    return 0;
}
