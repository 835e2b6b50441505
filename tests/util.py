"""shared helpers for the parity tests"""
import numpy as np

# north-star tolerances (BASELINE.json): <= 2/255 max per-channel RGBA difference and PSNR >= 45 dB;
# LIC-volume mode <= 1e-4 relative error
MAX_DIFF_8BIT = 2
MIN_PSNR_DB = 45.0
LICVOL_REL = 1e-4


def psnr8(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def compare_images(vvo, cuda_rgba32f, oracle_rgba32f):
    """both premultiplied float RGBA [h][w][4]; compared after the reference's RGBA8 store (Q18)"""
    a = vvo.quantize_rgba8(cuda_rgba32f)
    b = vvo.quantize_rgba8(oracle_rgba32f)
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), psnr8(a, b), float(np.abs(cuda_rgba32f - oracle_rgba32f).max())


def assert_image_parity(vvo, cuda_rgba32f, oracle_rgba32f, what=""):
    md, ps, mf = compare_images(vvo, cuda_rgba32f, oracle_rgba32f)
    assert md <= MAX_DIFF_8BIT, "%s: max 8-bit diff %d > %d (float max diff %.3g)" % (what, md, MAX_DIFF_8BIT, mf)
    assert ps >= MIN_PSNR_DB, "%s: PSNR %.2f dB < %.1f" % (what, ps, MIN_PSNR_DB)
    return md, ps, mf


def render_cuda(vv, scene, sample_map=True, layout=None):
    from vectorvisualization_b200.configs import apply_scene
    r = vv.Renderer(0)
    if layout is not None:
        r.setOption(vv.OPT_FIELD_LAYOUT, layout)
    apply_scene(r, scene)
    if sample_map:
        r.setOption(vv.OPT_SAMPLE_MAP, 1)
    r.render(True)
    img = r.readRGBA32F()
    img8 = r.readRGBA8()
    cnt = r.readSampleMap() if sample_map else None
    tot = r.lastRaySamples()
    return r, img, img8, cnt, tot
