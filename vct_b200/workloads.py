"""The five BASELINE.json configurations as runnable workloads (SURVEY.md §8(d) "Config 1..5"): scene, frame parameters,
grid / frame sizes and — for config 4 — the per-frame actor animation.  Used by bench.py (`--config N`), the GPU parity tests
(tests/test_gpu_configs.py) and the tools; host-side harness code, nothing here touches the device."""
import numpy as np

from . import params as P
from . import scene as S

LEVELS, SHADOW = 6, 4096          # VCT::voxelLevels (Application.h:133), SHADOWMAP_WIDTH/HEIGHT (Application.cpp:30-31)

NAMES = {1: "cube.obj (scale 0.5), volume +-1.5", 2: "bunny.obj (no UVs: NaN tangents), diffuse cones only",
         3: "PBR Sponza", 4: "PBR Sponza + 2 animated nanosuits, warp map (warpTexture), temporal radiance filter",
         5: "synthetic triangle soup (PCG seed 0x5EED, sigma 1.5 voxels)"}


class Workload:
    def __init__(self, config=3, width=None, height=None, dim=None, triangles=None, shadow=None, levels=None):
        if config not in NAMES:
            raise ValueError(f"config must be 1..5, got {config}")
        self.config = config
        if config == 5 and triangles:
            sc = S.Scene()
            sc.add_actor(S.soup_mesh(int(triangles), material=sc.add_material(diffuse=sc.add_texture(np.full((1, 1, 3), 255, np.uint8)))))
            sc.lights = P.reference_lights()
            cam, vol, D, size, extra = P.Camera(position=(0, 0, 30), front=(0, 0, -1)), (-20.0, 20.0, (0, 0, 0)), 512, (3840, 2160), {}
        else:
            sc, cam, vol, D, size, extra = S.config_scene(config)
        self.scene, self.camera = sc, cam
        self.D = dim or D
        self.W, self.H = width or size[0], height or size[1]
        self.L, self.S = levels or LEVELS, shadow or SHADOW
        self.data = extra.pop("data", "reference assets (baked)" if config in (1, 2) else "synthetic (seeded)")
        p = P.default_params(self.W, self.H, cam, sc.lights[0], voxel_min=vol[0], voxel_max=vol[1], voxel_center=vol[2])
        for k, v in extra.items():
            setattr(p, k, v)
        self.params = p
        self.animated = config == 4 and len(sc.meshes) >= 3
        # animated actors move the shadow map and the visibility buffer too: the step is the whole frame graph (producers included)
        self.whole_frame = self.animated

    def models(self, frame):
        """Actor transforms of frame `frame` at 60 Hz (config 4: Application.cpp:101-116): [(actor, 4x4 model)]."""
        if not self.animated:
            return []
        a, b = S.nanosuit_models(frame / 60.0)
        return [(1, a), (2, b)]

    @property
    def chains(self):
        return 2 if self.params.mip_color_chain else 1

    def describe(self):
        p = self.params
        flags = []
        if p.warp_texture: flags.append("warp map")
        if p.temporal_filter_radiance: flags.append(f"temporal radiance filter (decay {p.temporal_decay:g})")
        flags.append("diffuse+specular cones" if p.enable_reflections else "diffuse cones only")
        flags.append("actors animated at 60 Hz, whole frame graph per step (shadow map + visibility + GI passes)" if self.whole_frame
                     else "full per-frame revoxelisation; shadow map and visibility buffer are inputs")
        return (f"config {self.config}: {NAMES[self.config]}, {self.D}^3 voxels, {self.L} levels, {self.W}x{self.H}, {self.S}^2 shadow map, "
                f"{self.scene.n_tris} triangles, " + ", ".join(flags))

    def config_dict(self):
        """The `config` object of the bench line — identical for the GPU arm and the reference arm."""
        p = self.params
        return {"workload": self.describe(), "config_index": self.config, "dim": self.D, "levels": self.L, "width": self.W, "height": self.H,
                "shadow": self.S, "triangles": self.scene.n_tris, "mip_chains": self.chains,
                "voxelize_mode": "deterministic running average (canonical draw order)" if p.deterministic else "free-running CAS",
                "l2": "GPU arm: no explicit flush — the inputs of one step (4096^2 shadow map 64 MiB, fragment records, visibility buffer, scene geometry, "
                      "texture pyramid, material textures) exceed the 126 MB L2, so every pass starts L2-cold for its own inputs"}
