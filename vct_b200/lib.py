"""ctypes binding of libvct_b200.so (include/vct_b200.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in the product path."""
import ctypes as C
import os

from . import params as P

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "lib", "libvct_b200.so")

# every symbol include/vct_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "vct_create", "vct_destroy", "vct_remake", "vct_last_error", "vct_upload_mesh", "vct_upload_texture",
    "vct_set_material", "vct_set_actor_transform", "vct_set_lights", "vct_shadowmap", "vct_occupancy", "vct_warpmap",
    "vct_voxelize", "vct_transfer", "vct_inject", "vct_fill_holes", "vct_mip", "vct_mip_kernel", "vct_exchange", "vct_gbuffer",
    "vct_cone_trace", "vct_debug_voxels", "vct_frame", "vct_gi_passes", "vct_set_voxel_opacity", "vct_temporal_radiance_filter",
    "vct_filter3d", "vct_normalize_voxels_f16", "vct_read_image", "vct_read_volume", "vct_write_volume",
    "vct_read_shadowmap", "vct_write_shadowmap", "vct_read_visibility", "vct_get_counters", "vct_get_timings",
    "vct_get_cone_steps", "vct_sync", "vct_device_ptr", "vct_level_bytes", "vct_stream", "vct_launch_count",
    "vct_set_stream", "vct_set_profiling", "vct_get_kernel_times",
    "vct_exchange_setup", "vct_exchange_export", "vct_exchange_import", "vct_exchange_local", "vct_exchange_attach",
    "vct_frame_was_sparse", "vct_mask_parity", "vct_slab_stripe",
    "vct_read_image_async", "vct_read_image_wait",
    "vct_ingest_obj", "vct_ingest_image", "vct_ingest_free", "vct_ingest_log", "vct_ingest_get_mesh",
    "vct_ingest_get_material", "vct_ingest_get_texture", "vct_ingest_upload",
]

_lib = None


class VctError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(SO):
        raise VctError(f"{SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(make -C vct_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(SO)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    pp = C.POINTER(P.FrameParams)
    sig = {
        "vct_create": (ci, [C.POINTER(P.Config), C.POINTER(vp)]), "vct_destroy": (ci, [vp]), "vct_remake": (ci, [vp, ci, ci]),
        "vct_last_error": (C.c_char_p, [vp]),
        "vct_upload_mesh": (ci, [vp, ci, vp, C.c_size_t, C.c_size_t, vp, C.c_size_t, vp]),
        "vct_upload_texture": (ci, [vp, ci, ci, ci, ci, ci, vp]),
        "vct_set_material": (ci, [vp, ci, C.POINTER(P.Material)]), "vct_set_actor_transform": (ci, [vp, ci, C.POINTER(cf)]),
        "vct_set_lights": (ci, [vp, vp, ci]),
        "vct_mip": (ci, [vp, ci]), "vct_mip_kernel": (ci, [vp, ci, ci]), "vct_exchange": (ci, [vp]),
        "vct_set_voxel_opacity": (ci, [vp, cf]), "vct_temporal_radiance_filter": (ci, [vp, cf]), "vct_filter3d": (ci, [vp, ci, ci]),
        "vct_normalize_voxels_f16": (ci, [vp, vp, vp, cf]),
        "vct_read_image": (ci, [vp, vp]), "vct_read_image_async": (ci, [vp, vp]), "vct_read_image_wait": (ci, [vp, ci]), "vct_read_volume": (ci, [vp, ci, ci, vp]), "vct_write_volume": (ci, [vp, ci, ci, vp]),
        "vct_read_shadowmap": (ci, [vp, vp]), "vct_write_shadowmap": (ci, [vp, vp]), "vct_read_visibility": (ci, [vp, vp]),
        "vct_get_counters": (ci, [vp, C.POINTER(P.VoxelizeInfo)]), "vct_get_timings": (ci, [vp, C.POINTER(P.Timings)]),
        "vct_get_cone_steps": (ci, [vp, C.POINTER(C.c_ulonglong)]), "vct_sync": (ci, [vp]),
        "vct_device_ptr": (vp, [vp, ci, ci]), "vct_level_bytes": (C.c_size_t, [vp, ci, ci]), "vct_stream": (vp, [vp]),
        "vct_launch_count": (C.c_ulonglong, [vp, ci]),
        "vct_set_stream": (ci, [vp, vp]), "vct_set_profiling": (ci, [vp, ci]),
        "vct_get_kernel_times": (ci, [vp, C.POINTER(P.KernelTime), ci]),
        "vct_exchange_setup": (ci, [vp]), "vct_exchange_export": (ci, [vp, vp]), "vct_exchange_import": (ci, [vp, ci, vp]),
        "vct_ingest_obj": (ci, [C.c_char_p, C.c_char_p, ci, C.POINTER(vp)]), "vct_ingest_image": (ci, [C.c_char_p, ci, C.POINTER(vp)]),
        "vct_ingest_free": (None, [vp]), "vct_ingest_log": (C.c_char_p, [vp]),
        "vct_ingest_get_mesh": (ci, [vp, C.POINTER(P.IngestMesh)]),
        "vct_ingest_get_material": (ci, [vp, ci, C.POINTER(P.Material), C.POINTER(C.c_char_p)]),
        "vct_ingest_get_texture": (ci, [vp, ci, C.POINTER(P.IngestTexture)]),
        "vct_ingest_upload": (ci, [vp, vp, ci, ci, ci, C.POINTER(cf)]),
        "vct_exchange_local": (ci, [vp, C.POINTER(P.Peer)]), "vct_exchange_attach": (ci, [vp, ci, C.POINTER(P.Peer)]),
        "vct_frame_was_sparse": (ci, [vp]), "vct_mask_parity": (ci, [vp]), "vct_slab_stripe": (ci, [vp]),
    }
    for name in ("vct_shadowmap", "vct_occupancy", "vct_warpmap", "vct_voxelize", "vct_transfer", "vct_inject", "vct_fill_holes",
                 "vct_gbuffer", "vct_cone_trace", "vct_debug_voxels", "vct_frame", "vct_gi_passes"):
        sig[name] = (ci, [vp, pp])
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
