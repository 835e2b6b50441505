"""CPU-only checks of the product's boundary: the C-ABI library loads, exports exactly what include/vv_c_api.h declares,
fails loudly without a GPU (no CPU fallback), and its host-only loaders behave like the reference's."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "vv_c_api.h")).read()
    return sorted(set(re.findall(r"VV_API[^;(]*?\b(vv_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(vv):
    lib = vv.load_library()
    names = _declared()
    assert len(names) > 50
    for n in names:
        assert hasattr(lib, n), "libvv_b200.so does not export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", vv.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vv_[a-z0-9_]+)", out)))
    assert exported == names, "header and library disagree: %s" % sorted(set(exported) ^ set(names))


def test_library_is_sm100a_and_has_no_oracle_dependency(vv):
    out = subprocess.run(["cuobjdump", "-lelf", vv.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    deps = subprocess.run(["ldd", vv.LIB_PATH], capture_output=True, text=True).stdout
    assert "vv_oracle" not in deps and "vv_ref" not in deps
    # the product sources never include anything from oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vectorvisualization_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in txt.replace("oracle/_ref", "").replace("the oracle", "") or f == "build.py", f
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_no_cpu_fallback(vv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vv.VVError) as e:
        vv.Renderer(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_default_lic_params(vv):
    p = vv.LICParams()
    q = vv.LICParams(stepSizeVol=0, gradientScale=0)
    vv.load_library().vv_default_lic_params(ctypes.byref(q))
    for f, _ in vv.LICParams._fields_:
        assert getattr(p, f) == getattr(q, f), f
    assert (p.stepSizeVol, p.gradientScale, p.illumScale, p.freqScale) == (1 / 128, 30.0, 1.0, 1.0)      # VV/types.h:93-98
    assert (p.numIterations, p.stepsForward, p.stepsBackward) == (255, 32, 32) and p.stepSizeLIC == np.float32(0.01)


def test_parse_args_surface(vv):
    a = vv.parse_args(["volic", "data/out.64.dat", "-g", "-n", "noise_256_80", "-f", "kernel/gauss.png", "-t", "ct/tf.png"])
    assert a.vol_file == b"data/out.64.dat" and a.use_gradients == 1 and a.noise_file == b"noise_256_80"
    assert a.filter_file == b"kernel/gauss.png" and a.tf_file == b"ct/tf.png"
    a = vv.parse_args(["volic", "-h"])
    assert a.show_help == 1                      # the reference prints the usage and exit(0)s; the library reports it
    assert b"--filter=<png>" in vv.load_library().vv_usage()
    for bad in (["volic"], ["volic", "a.dat", "b.dat"], ["volic", "a.dat", "-q"], ["volic", "a.dat", "-n"], ["volic", "a.dat", "-t", "-g"]):
        with pytest.raises(vv.VVError):
            vv.parse_args(bad)


def test_dat_and_raw_reader(vv, tmp_path):
    from vectorvisualization_b200 import fields as F
    f = F.abc_flow(6)
    dat = F.write_dat(str(tmp_path / "v.dat"), f, slice_thickness=(1, 1, 2))
    info = vv.parse_dat(dat)
    assert tuple(info.resolution) == (6, 6, 6) and tuple(info.slice_thickness) == (1.0, 1.0, 2.0)
    assert info.data_type == vv.FLOAT and info.data_dim == 3 and (info.time_begin, info.time_end) == (0, 0)
    buf = np.zeros_like(f)
    rc = vv.load_library().vv_read_raw(ctypes.byref(info), 0, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    assert rc == 0 and np.array_equal(buf, f)
    assert vv.load_library().vv_read_raw(ctypes.byref(info), 1, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes) != 0      # out of range
    assert vv.load_library().vv_read_raw(ctypes.byref(info), 0, buf.ctypes.data_as(ctypes.c_void_p), 8) != 0               # buffer too small
    # time-dependent set with a printf pattern, comment lines, raw file next to the .dat
    dat2 = F.write_dat(str(tmp_path / "sub" / "t.dat") if os.makedirs(tmp_path / "sub", exist_ok=True) is None else "", None,
                       time_steps=[f, f * 2, f * 3])
    info2 = vv.parse_dat(dat2)
    assert (info2.time_begin, info2.time_end) == (0, 2) and b"%d" in info2.raw_file
    vv.load_library().vv_read_raw(ctypes.byref(info2), 2, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    assert np.array_equal(buf, f * 3)
    with pytest.raises(vv.VVError):
        vv.parse_dat(str(tmp_path / "missing.dat"))
    p = tmp_path / "fmt.dat"
    p.write_text("ObjectFileName: v.raw\nResolution: 6 6 6\nFormat: DOUBLE\n")
    with pytest.raises(vv.VVError):
        vv.parse_dat(str(p))


def test_png_codec_roundtrip(vv, tmp_path):
    from vectorvisualization_b200 import fields as F
    rng = np.random.RandomState(0)
    for ch in (1, 2, 3, 4):
        img = rng.randint(0, 256, size=(7, 13, ch)).astype(np.uint8)
        # written by the python writer (all filter type 0) and by the library, read by the library
        p1 = F.write_png(str(tmp_path / ("a%d.png" % ch)), img)
        assert np.array_equal(vv.png_read(p1), img)
        p2 = str(tmp_path / ("b%d.png" % ch))
        assert vv.load_library().vv_png_write(p2.encode(), img.ctypes.data_as(ctypes.c_void_p), 13, 7, ch) == 0
        assert np.array_equal(vv.png_read(p2), img)
    # PIL-written PNGs use adaptive filters (Sub/Up/Average/Paeth)
    try:
        from PIL import Image
    except Exception:
        return
    img = (np.add.outer(np.arange(64), np.arange(48)) % 251).astype(np.uint8)
    img = np.stack([img, img[::-1], img.T[:64, :48] if False else img, 255 - img], axis=-1)
    path = str(tmp_path / "pil.png")
    Image.fromarray(img, "RGBA").save(path, optimize=True)
    assert np.array_equal(vv.png_read(path), img)
    with pytest.raises(vv.VVError):
        vv.png_read(str(tmp_path / "nope.png"))


def test_synthetic_generators_are_deterministic():
    from vectorvisualization_b200 import fields as F
    a, b = F.curl_noise(16, 4), F.curl_noise(16, 4)
    assert np.array_equal(a, b) and a.shape == (16, 16, 16, 3) and np.isfinite(a).all()
    assert not np.array_equal(a, F.curl_noise(16, 5))
    t = F.tornado(16)
    assert np.isfinite(t).all() and np.abs(t).max() > 0
    r = F.rankine_vortex(16)
    assert np.allclose(r[..., 2], 0.2)
    for name in ("box", "triangle", "gaussian", "cos2"):
        k = F.filter_kernel(name)
        assert k.shape == (256,) and k.max() >= 254 and np.array_equal(k, k[::-1])
    tf = F.tf_preset("tf-length")
    assert tf.shape == (256, 5) and tf[:, 3].max() <= 26          # semi-transparent: no early ray termination


def test_hot_kernel_resources_and_instruction_mix(vv):
    """static guard on the shipped lic_sample_kernel<xy-quad field, gradient build, bf16 noise, guard band + shared cell> (what DESIGN.md
    section 5 measures): 72 registers (7 CTAs x 128 threads per SM) and the sm_100a instructions the design relies on -- FHADD
    (f32 = f16 + f32) for the fp16 field texels, packed FFMA2 / FADD2 lerps, PRMT widening of the bf16 noise, 256-bit field loads, no
    index clamps (FMNMX) in the walk, no accumulator spills; two copies of the walk loop: taps with the field's cell (hot) and with
    the REPEAT arithmetic (ray samples near the faces)"""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from sass_loop_stats import stats
    o = stats(vv.LIB_PATH, "lic_sample_kernelILi2ELi1ELb0ELb0ELi2ELi3E")
    assert "REG:72" in o["usage"].replace(" ", ""), o["usage"]
    assert o["total"] > 2000
    loops = o["loops"][:2]                            # the walk loops: one backward + one forward Heun step and their two noise taps
    assert any(L["ops"].get("FRND", 0) == 0 for L in loops)       # the hot copy carries no REPEAT floor
    for loop in loops:
        ops = loop["ops"]
        assert 420 < loop["n"] < 560, loop["n"]
        assert ops.get("FHADD", 0) == 96              # 24 per field cell; 2 cells per Heun step in the code (the second only when the corrector leaves the predictor's cell)
        assert ops.get("PRMT", 0) >= 64               # 2 noise taps x 4 rows x 4 words x 2 halves
        assert ops.get("FFMA2", 0) >= 90 and ops.get("FADD2", 0) >= 36
        assert ops.get("FMNMX", 0) == 0               # guard band: no coordinate clamp in the walk
        assert ops.get("LDG", 0) == 16                # (2 faces of the field cell as 256-bit loads x 2 for the rare reload + 4 noise rows) x 2 directions
        assert ops.get("STL", 0) <= 4 and ops.get("LDL", 0) <= 4   # at most the two walker positions parked across the rare reload path
    # the scalar builds keep 4 CTAs x 256 threads (64 registers)
    o = stats(vv.LIB_PATH, "lic_sample_kernelILi2ELi0ELb1ELb0ELi2ELi1E")
    assert "REG:64" in o["usage"].replace(" ", ""), o["usage"]
