"""Host-side mirror of the reference's `Application` + `VCT` for the GI hot path (Python flavour of
vct_b200/host/vct_host.hpp): owns a vct_ctx, uploads a Scene the way Mesh/Scene do, and exposes one method per
GL pass block of Application::render (reference src/Application.cpp:196-1085), same names, same order.
Every method goes through the C ABI; a non-zero status raises VctError with vct_last_error()."""
import ctypes as C

import numpy as np

from . import params as P
from .lib import VctError, load


class Pipeline:
    def __init__(self, scene, dim=256, levels=6, shadow_size=4096, width=1280, height=720, device=0, rank=0,
                 world_size=1, max_fragments=0, devices=None, slab_stripe=0):
        """devices: a list of CUDA ordinals -> ONE handle driving all of them from this process (vct_config.n_devices)."""
        self.lib = load()
        self._devices = (C.c_int * len(devices))(*devices) if devices else None
        self.cfg = P.Config(dim, levels, shadow_size, width, height, device, rank, world_size, max_fragments,
                            len(devices) if devices else 0, self._devices, slab_stripe)
        h = C.c_void_p()
        if self.lib.vct_create(C.byref(self.cfg), C.byref(h)):
            raise VctError(self.lib.vct_last_error(None).decode())
        self.h = h
        self.D, self.S, self.W, self.H = dim, shadow_size, width, height
        self.L = max(1, min(levels, int(np.log2(dim)) + 1))
        if scene is not None:
            self.upload(scene)

    def _ck(self, status):
        if status:
            raise VctError(self.lib.vct_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.vct_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- scene (Mesh::loadMesh buffers, textures, lights)
    def upload(self, scene):
        for i, t in enumerate(scene.textures):
            px = t.packed()
            self._ck(self.lib.vct_upload_texture(self.h, i, t.width, t.height, t.channels, min(16, len(t.levels)), px.ctypes.data))
        for i, m in enumerate(scene.materials):
            self._ck(self.lib.vct_set_material(self.h, i, C.byref(m)))
        for a, (mesh, model) in enumerate(zip(scene.meshes, scene.models)):
            self._ck(self.lib.vct_upload_mesh(self.h, a, mesh.vertices.ctypes.data, len(mesh.vertices), 56, mesh.indices.ctypes.data,
                                              mesh.indices.size, mesh.tri_material.ctypes.data))
            self.set_actor_transform(a, model)
        self.set_lights(scene.lights)

    def set_actor_transform(self, actor, model):
        m = np.ascontiguousarray(np.asarray(model, np.float32).reshape(16))
        self._ck(self.lib.vct_set_actor_transform(self.h, actor, m.ctypes.data_as(C.POINTER(C.c_float))))

    def set_lights(self, lights):
        arr = (P.Light * max(1, len(lights)))(*lights)
        self._ck(self.lib.vct_set_lights(self.h, C.cast(arr, C.c_void_p), len(lights)))

    def remake(self, dim, levels):
        self._ck(self.lib.vct_remake(self.h, dim, levels))
        self.D, self.L = dim, max(1, min(levels, int(np.log2(dim)) + 1))

    # ---- passes
    def shadowmap(self, p): self._ck(self.lib.vct_shadowmap(self.h, C.byref(p)))
    def occupancy(self, p): self._ck(self.lib.vct_occupancy(self.h, C.byref(p)))
    def warpmap(self, p): self._ck(self.lib.vct_warpmap(self.h, C.byref(p)))
    def voxelize(self, p): self._ck(self.lib.vct_voxelize(self.h, C.byref(p)))
    def transfer(self, p): self._ck(self.lib.vct_transfer(self.h, C.byref(p)))
    def inject(self, p): self._ck(self.lib.vct_inject(self.h, C.byref(p)))
    def fill_holes(self, p): self._ck(self.lib.vct_fill_holes(self.h, C.byref(p)))
    def mip(self, which=P.VOL_RADIANCE): self._ck(self.lib.vct_mip(self.h, which))
    def mip_kernel(self, which, mode): self._ck(self.lib.vct_mip_kernel(self.h, which, mode))
    def exchange(self): self._ck(self.lib.vct_exchange(self.h))
    def frame_was_sparse(self): return bool(self.lib.vct_frame_was_sparse(self.h))
    def mask_parity(self): return int(self.lib.vct_mask_parity(self.h))

    def slab_stripe(self): return int(self.lib.vct_slab_stripe(self.h))

    def exchange_setup(self):
        """Allocate the slab-exchange staging buffer and return this rank's cudaIpc handle blob (one process per GPU)."""
        self._ck(self.lib.vct_exchange_setup(self.h))
        buf = C.create_string_buffer(P.EXCHANGE_HANDLE_BYTES)
        self._ck(self.lib.vct_exchange_export(self.h, buf))
        return bytes(buf.raw)

    def exchange_import(self, rank, handle):
        self._ck(self.lib.vct_exchange_import(self.h, rank, C.create_string_buffer(handle, P.EXCHANGE_HANDLE_BYTES)))
    def gbuffer(self, p): self._ck(self.lib.vct_gbuffer(self.h, C.byref(p)))
    def cone_trace(self, p): self._ck(self.lib.vct_cone_trace(self.h, C.byref(p)))
    def debug_voxels(self, p): self._ck(self.lib.vct_debug_voxels(self.h, C.byref(p)))
    def frame(self, p): self._ck(self.lib.vct_frame(self.h, C.byref(p)))
    def gi_passes(self, p): self._ck(self.lib.vct_gi_passes(self.h, C.byref(p)))
    def sync(self): self._ck(self.lib.vct_sync(self.h))

    # ---- outputs
    def read_volume(self, which, level=0):
        nbytes = self.lib.vct_level_bytes(self.h, which, level)
        if not nbytes:
            raise VctError(self.lib.vct_last_error(self.h).decode())
        dt = np.uint16 if which in (P.VOL_WARPMAP, P.VOL_WARP_WEIGHTS_LOW, P.VOL_WARP_WEIGHTS_HIGH) else np.uint32
        out = np.empty(nbytes // np.dtype(dt).itemsize, dt)
        self._ck(self.lib.vct_read_volume(self.h, which, level, out.ctypes.data))
        return out

    def write_volume(self, which, level, data):
        data = np.ascontiguousarray(data)
        assert data.nbytes == self.lib.vct_level_bytes(self.h, which, level)
        self._ck(self.lib.vct_write_volume(self.h, which, level, data.ctypes.data))

    def read_shadowmap(self):
        out = np.empty(self.S * self.S, np.float32); self._ck(self.lib.vct_read_shadowmap(self.h, out.ctypes.data)); return out

    def write_shadowmap(self, d):
        d = np.ascontiguousarray(d, np.float32); self._ck(self.lib.vct_write_shadowmap(self.h, d.ctypes.data))

    def read_visibility(self):
        out = np.empty(self.W * self.H, np.uint64); self._ck(self.lib.vct_read_visibility(self.h, out.ctypes.data)); return out

    def read_image(self, out=None):
        if out is None:
            out = np.empty(self.W * self.H, np.uint32)
        self._ck(self.lib.vct_read_image(self.h, out.ctypes.data)); return out

    def read_image_async(self, pinned_ptr):
        """Enqueue the read-back of the current image into pinned host memory (int address); returns at once."""
        self._ck(self.lib.vct_read_image_async(self.h, pinned_ptr))

    def read_image_wait(self, block_host=True):
        self._ck(self.lib.vct_read_image_wait(self.h, 1 if block_host else 0))

    def image_rgba(self):
        return self.read_image().view(np.uint8).reshape(self.H, self.W, 4)[::-1]

    def counters(self):
        info = P.VoxelizeInfo(); self._ck(self.lib.vct_get_counters(self.h, C.byref(info))); return info

    def cone_steps(self):
        n = C.c_ulonglong(0); self._ck(self.lib.vct_get_cone_steps(self.h, C.byref(n))); return n.value

    def timings(self):
        t = P.Timings(); self._ck(self.lib.vct_get_timings(self.h, C.byref(t))); return {k: getattr(t, k) for k, _ in P.Timings._fields_}

    def set_profiling(self, level):
        self._ck(self.lib.vct_set_profiling(self.h, level))

    def set_stream(self, stream_ptr):
        self._ck(self.lib.vct_set_stream(self.h, stream_ptr))

    def kernel_times(self):
        """{kernel name: (ns, launches)} of the last pass call (profiling level 2)."""
        buf = (P.KernelTime * 64)()
        n = self.lib.vct_get_kernel_times(self.h, buf, 64)
        if n < 0:
            raise VctError("vct_get_kernel_times failed")
        return {buf[i].name.decode(): (buf[i].ns, buf[i].launches) for i in range(n)}

    def launch_count(self, reset=False):
        return int(self.lib.vct_launch_count(self.h, int(reset)))

    def device_ptr(self, which, level=0):
        return self.lib.vct_device_ptr(self.h, which, level)

    def level_bytes(self, which, level=0):
        return self.lib.vct_level_bytes(self.h, which, level)
