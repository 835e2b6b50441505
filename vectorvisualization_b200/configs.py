"""The five BASELINE.json configurations as plain data, plus `apply_scene` which hands a
scene to a `Renderer` in the order VV/3DLIC.cpp:677-799 (`init`) does."""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import LICParams, Renderer, VOLIC_RAYCAST, TF_B, TF_LENGTH, GATE_ALWAYS
from . import fields as F
from . import (OPT_TF_MODE, OPT_GATE_MODE, OPT_NOISE_GATE, OPT_QUIRK_SCALEVOLINV, OPT_QUIRK_LUMINANCE_ALPHA,
               OPT_LICVOL_FP16, OPT_LICVOL_SIZE)


@dataclass
class Scene:
    name: str
    field: np.ndarray                      # [z][y][x][3] float32
    noise: np.ndarray                      # [z][y][x] uint8
    scalar: Optional[np.ndarray]           # [z][y][x] uint8
    filter_row: Optional[np.ndarray]       # uint8 row; None = box filter
    tf: np.ndarray                         # [256][5] uint8
    width: int
    height: int
    with_gradients: bool = False           # -g
    defines: str = ""                      # "#define ILLUM_GRADIENT" ...
    params: dict = field(default_factory=dict)        # LICParams overrides
    camera: dict = field(default_factory=lambda: dict(F.CAMERA_DEFAULT))
    light: dict = field(default_factory=lambda: dict(quat=(0.0, 0.0, 0.0, 1.0), dist=1.0))
    tf_mode: int = TF_B
    gate_mode: int = GATE_ALWAYS
    noise_gate: int = 1
    technique: int = VOLIC_RAYCAST
    lowres: int = 0
    quirk_scalevolinv: int = 1
    quirk_luminance_alpha: int = 0
    licvol_fp16: int = 1
    licvol_size: int = 0
    next_field: Optional[np.ndarray] = None
    interp: tuple = (0, 10)
    slice_dist: tuple = (1.0, 1.0, 1.0)
    mc_offsets: Optional[np.ndarray] = None   # [height][width] float32 in [0,1]; needs "#define USE_MC_OFFSET" in defines
    clip_planes: tuple = ()                # up to 3 active planes (nx, ny, nz, d): n.q + d >= 0 kept, q relative to the centre
    window: Optional[tuple] = None         # (w, h) of the window when the frame is not the window (low-res preset: frame = half the window)
    fbo: int = 0                           # Renderer::enableFBO (key 'F'; off at start-up, VV/renderer.cpp:33): float RGBA16F targets;
                                           # for VOLIC_SLICING it selects the FBO ping-pong program over the blending one

    def lic_params(self):
        return LICParams(**self.params)


def cfg1(n=64, size=512, camera=None):
    """64^3 ABC flow, 64^3 sparse noise seed 1, box filter, h = 0.01, step 1/64, 512^2, illumLIC"""
    return Scene("cfg1", F.abc_flow(n), F.white_noise(n, 1, F.SPARSE_P), F.constant_scalar(), None, F.default_tf(), size, size,
                 params=dict(stepSizeVol=1.0 / 64.0, stepSizeLIC=0.01, freqScale=1.0), camera=dict(camera or F.CAMERA_DEFAULT))


def cfg2(n=128, size=1024, camera=None):
    """128^3 Rankine vortex, dense noise freq 4.0, Gaussian filter, step 1/128, 1024^2"""
    return Scene("cfg2", F.rankine_vortex(n), F.white_noise(n, 2, F.DENSE_P), F.constant_scalar(), F.filter_kernel("gaussian"),
                 F.default_tf(), size, size, params=dict(stepSizeVol=1.0 / 128.0, freqScale=4.0),
                 camera=dict(camera or F.CAMERA_DEFAULT))


def cfg3(n=256, size=1024, camera=None, noise_n=None):
    """256^3 Crawfis tornado, cos^2 filter, length TF + gradient illumination, h = 0.005, step 1/128, 1024^2 (headline)"""
    nn = noise_n or n
    return Scene("cfg3", F.tornado(n), F.white_noise(nn, 3, F.SPARSE_P), F.constant_scalar(), F.filter_kernel("cos2"),
                 F.tf_preset("tf-length"), size, size, with_gradients=True, defines="#define ILLUM_GRADIENT",
                 params=dict(stepSizeVol=1.0 / 128.0, stepSizeLIC=0.005), tf_mode=TF_LENGTH,
                 camera=dict(camera or F.CAMERA_DEFAULT))


def cfg3o(n=256, size=1024, camera=None, noise_n=None):
    """cfg3 with an opaque transfer function -- the reference's default ramp (alpha = opacity = max(0, i - 20), VV/transferEdit.cpp:76-82)
    twice as steep, saturating at 255: samples reach src.a > 0.95, rays terminate early (lic3d_fragment.glsl:91) and the frame is
    computed in depth windows.  (With the default ramp itself and alphaCorrection = 1 no sample of this scene can exceed 0.85.)"""
    s = cfg3(n, size, camera, noise_n)
    s.tf = F.default_tf().copy()
    ramp = np.clip(2 * (np.arange(256) - 20), 0, 255).astype(np.uint8)
    s.tf[:, 3] = ramp
    s.tf[:, 4] = ramp
    s.name = "cfg3o"
    return s


def cfg1t(n=64, size=512, camera=None):
    """cfg1 with a semi-transparent transfer function (no early ray termination): the single-window reference point of cfg1"""
    s = cfg1(n, size, camera)
    s.tf = s.tf.copy()
    s.tf[:, 3] = (s.tf[:, 3].astype(np.int32) * 3 // 10).astype(np.uint8)
    s.name = "cfg1t"
    return s


def cfg4(n=512, size=2048, camera=None, noise_n=256):
    """512^3 curl-noise turbulence, 256^3 sparse noise (REPEAT-tiled), triangle filter, step 1/256, 2048^2"""
    return Scene("cfg4", F.curl_noise(n, 4), F.white_noise(noise_n, 4, F.SPARSE_P), F.constant_scalar(), F.filter_kernel("triangle"),
                 F.default_tf(), size, size, params=dict(stepSizeVol=1.0 / 256.0), camera=dict(camera or F.CAMERA_DEFAULT))


def cfg5(n=1024, size=4096, camera=None, noise_n=256):
    """precomputed LIC volume (n^3) + plain ray-cast, step 1/256, size^2"""
    from . import VOLIC_LICVOLUME
    return Scene("cfg5", F.curl_noise(n, 5), F.white_noise(noise_n, 5, F.SPARSE_P), F.constant_scalar(), F.filter_kernel("triangle"),
                 F.default_tf(), size, size, params=dict(stepSizeVol=1.0 / 256.0), technique=VOLIC_LICVOLUME,
                 camera=dict(camera or F.CAMERA_DEFAULT))


def apply_scene(r: Renderer, s: Scene):
    """the sequence of VV/3DLIC.cpp:677-799 (init) against the C ABI"""
    r.init(s.defines or None)
    r.setVectorField(s.field, s.next_field, s.slice_dist)
    if s.next_field is not None or s.interp[0] != 0:
        r.setTimeInterp(*s.interp)
    r.setNoise(s.noise, s.with_gradients)
    if s.scalar is not None:
        r.setScalar(s.scalar)
    r.setLICFilter(s.filter_row)
    r.setTF(s.tf)
    r.setTechnique(s.technique)          # TF-index / gate options are kept per shader program (ray-cast vs slicing)
    r.setOption(OPT_TF_MODE, s.tf_mode)
    r.setOption(OPT_GATE_MODE, s.gate_mode)
    r.setOption(OPT_NOISE_GATE, s.noise_gate)
    r.setOption(OPT_QUIRK_SCALEVOLINV, s.quirk_scalevolinv)
    r.setOption(OPT_QUIRK_LUMINANCE_ALPHA, s.quirk_luminance_alpha)
    r.setOption(OPT_LICVOL_FP16, s.licvol_fp16)
    r.setOption(OPT_LICVOL_SIZE, s.licvol_size)
    r.enableLowRes(s.lowres)
    r.setLICParams(s.lic_params())
    r.setCamera(**s.camera)
    r.setLight(**s.light)
    r.updateLightPos()
    r.setTechnique(s.technique)
    r.enableFBO(s.fbo)
    r.resize(s.width, s.height)
    r.setWindow(*(s.window or (0, 0)))
    if s.mc_offsets is not None:
        r.setMCOffsets(s.mc_offsets)
    for i in range(3):
        if i < len(s.clip_planes):
            r.setClipPlane(i, s.clip_planes[i], True)
        else:
            r.setClipPlane(i, (0.0, 0.0, -1.0, 0.0), False)
    return r
