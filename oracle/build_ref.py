#!/usr/bin/env python
"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.

Builds oracle/_ref/libvv_ref.so from the REFERENCE'S OWN SOURCES, read where they lie under /root/reference
(nothing is copied into the repository; oracle/_ref/ is git-ignored and holds only generated/built files):

  1. the GLSL shaders of the hot path (VV/shader/inc_header.glsl, inc_lic.glsl, inc_illum.glsl, lic3d_fragment.glsl,
     lic3d_volume_fragment.glsl, raycast_lic3d_fragment.glsl, lic3d_slicing_fragment.glsl), concatenated per program
     exactly as Renderer::loadGLSLShader does (VV/renderer.cpp:807-922, VV/GLSLShader.cpp:117-160: "#version 120" +
     defines + files), passed through a purely LEXICAL rewrite and compiled as C++ against oracle/glsl_shim.h;
  2. the C++ loaders / pre-processing (VV/reader.cpp, parseArg.cpp, gradient.cpp, mmath.cpp, dataset.cpp,
     transferEdit.cpp) compiled unmodified against a capturing GL stub (oracle/ref_shim/), see ref_host_driver.cpp.

The lexical rewrite (all of it):
  - strip comments and `#extension` lines;
  - `uniform T x;`            -> `T x;`                      (namespace-scope variable the driver assigns)
  - `in out T x` / `out T x`  -> `T& x`,  `in T x` -> `T x`  (GLSL parameter qualifiers; Q9)
  - float literals get an `f` suffix                        (GLSL literals are single precision)
  - `void main(void)`         -> `void shader_main(void)`
  - `float logEyeDist;`       -> `float logEyeDist = 0.0f;`  (Q3: read before ever being written; 0 on NVIDIA/llvmpipe)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/VectorVisualization"
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
LIB = os.path.join(OUT, "libvv_ref.so")

# program name -> (fragment shader file list as in Renderer::loadGLSLShader, defines, extra C++ defines)
RAYCAST = ["inc_header.glsl", "inc_lic.glsl", "inc_illum.glsl", "lic3d_fragment.glsl"]
LICVOL = ["inc_header.glsl", "inc_lic.glsl", "lic3d_volume_fragment.glsl"]
VOLRAY = ["inc_header.glsl", "inc_illum.glsl", "raycast_lic3d_fragment.glsl"]
SLICING = ["inc_header.glsl", "inc_lic.glsl", "inc_illum.glsl", "lic3d_slicing_fragment.glsl"]
SLICINGBLEND = ["inc_header.glsl", "inc_lic.glsl", "inc_illum.glsl", "lic3d_slicingblend_fragment.glsl"]
PROGRAMS = {
    "raycast_none": (RAYCAST, [], []),
    "raycast_gradient": (RAYCAST, ["ILLUM_GRADIENT"], []),
    "raycast_mallo": (RAYCAST, ["ILLUM_MALLO"], []),
    "raycast_zoeckler": (RAYCAST, ["ILLUM_ZOECKLER"], []),
    "raycast_sof": (RAYCAST, ["SPEED_OF_FLOW"], []),
    "licvol_none": (LICVOL, [], []),
    "licvol_gradient": (LICVOL, ["ILLUM_GRADIENT"], []),
    "licvol_sof": (LICVOL, ["SPEED_OF_FLOW"], []),
    "volraycast": (VOLRAY, [], ["REF_HAS_LICVOL"]),
    "slicing_none": (SLICING, [], ["REF_SLICING"]),
    "slicing_gradient": (SLICING, ["ILLUM_GRADIENT"], ["REF_SLICING"]),
    "slicing_mallo": (SLICING, ["ILLUM_MALLO"], ["REF_SLICING"]),
    "slicing_zoeckler": (SLICING, ["ILLUM_ZOECKLER"], ["REF_SLICING"]),
    # "// TODO: MC offset" builds (inc_header.glsl:14, lic3d_fragment.glsl:31-33, lic3d_slicing_fragment.glsl:31-33)
    "raycast_none_mc": (RAYCAST, ["USE_MC_OFFSET"], []),
    "raycast_gradient_mc": (RAYCAST, ["ILLUM_GRADIENT", "USE_MC_OFFSET"], []),
    "slicing_none_mc": (SLICING, ["USE_MC_OFFSET"], ["REF_SLICING"]),
    # the variant sliceVolume runs without the FBO (the start-up state): the shader returns the premultiplied sample, the GL blends
    "slicingblend_none": (SLICINGBLEND, [], []),
    "slicingblend_gradient": (SLICINGBLEND, ["ILLUM_GRADIENT"], []),
    "slicingblend_mallo": (SLICINGBLEND, ["ILLUM_MALLO"], []),
}
# Source-edit variants: the reference switches the TF index and the LIC gate by (un)commenting lines of
# lic3d_fragment.glsl:53-61.  Each variant swaps the live expression for one of the alternatives the file itself lists.
TF_EXPR = {"a": "vectorData.a", "r": "vectorData.x", "length": "length(vectorData)", "scalar": "scalarData.r"}
EDITS = {}
for _base in ("raycast_none", "raycast_gradient", "raycast_sof", "raycast_mallo", "raycast_zoeckler"):
    for _k, _e in TF_EXPR.items():
        PROGRAMS["%s_tf%s" % (_base, _k)] = PROGRAMS[_base]
        EDITS["%s_tf%s" % (_base, _k)] = [("texture1D(transferRGBASampler, vectorData.b)", "texture1D(transferRGBASampler, %s)" % _e)]
PROGRAMS["raycast_none_gatetf"] = PROGRAMS["raycast_none"]
EDITS["raycast_none_gatetf"] = [("if (scalarData.g > -0.0001)", "if (tfData.a > 0.05)")]
PROGRAMS["raycast_none_gatetf_tfa"] = PROGRAMS["raycast_none"]
EDITS["raycast_none_gatetf_tfa"] = EDITS["raycast_none_gatetf"] + EDITS["raycast_none_tfa"]


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def rewrite(src):
    src = strip_comments(src)
    src = re.sub(r"^[ \t]*#extension[^\n]*$", "", src, flags=re.M)
    src = re.sub(r"\buniform\s+", "", src)
    src = re.sub(r"\bin\s+out\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bout\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*void\s*\)", "void shader_main(void)", src)
    src, n = re.subn(r"\bfloat\s+logEyeDist\s*;", "float logEyeDist = 0.0f;", src)
    return src


def generate():
    os.makedirs(GEN, exist_ok=True)
    files = []
    for name, (shaders, defines, cxxdefs) in PROGRAMS.items():
        body = "".join(open(os.path.join(REF, "shader", f)).read() for f in shaders)   # loadSource concatenation
        body = strip_comments(body)
        for old, new in EDITS.get(name, []):
            assert body.count(old) == 1, (name, old, body.count(old))
            body = body.replace(old, new)
        text = ["// GENERATED by oracle/build_ref.py from %s -- do not commit" % ", ".join("VV/shader/" + f for f in shaders),
                '#include "glsl_shim.h"', '#include "ref_api.h"']
        text += ["#define %s" % d for d in defines + cxxdefs]
        text += ["namespace glsl { namespace %s {" % name, rewrite(body), '#include "ref_frag_driver.inc"', "} }",
                 'extern "C" void vvref_run_%s(const RefUniforms *u, const float *tc, int n, float *out, uint32_t *cnt)' % name,
                 "{ glsl::%s::run(*u, tc, n, out, cnt); }" % name]
        if "REF_SLICING" in cxxdefs:
            text += ['extern "C" void vvref_slice_%s(const RefUniforms *u, const float *fr, const int *st, int np, float *out, uint32_t *cnt)' % name,
                     "{ glsl::%s::run_slicing(*u, fr, st, np, out, cnt); }" % name]
        path = os.path.join(GEN, "prog_%s.cpp" % name)
        with open(path, "w") as f:
            f.write("\n".join(text) + "\n")
        files.append(path)
    # the display pass: background_fragment.glsl is a program of its own (VV/renderer.cpp:811, 851-857)
    body = rewrite(open(os.path.join(REF, "shader", "background_fragment.glsl")).read())
    text = ["// GENERATED by oracle/build_ref.py from VV/shader/background_fragment.glsl -- do not commit",
            '#include "glsl_shim.h"', '#include "ref_api.h"', "namespace glsl { namespace background {", body,
            '#include "ref_bg_driver.inc"', "} }",
            'extern "C" void vvref_run_background(const float *img, int rw, int rh, int ww, int wh, float *out)',
            "{ glsl::background::run_bg(img, rw, rh, ww, wh, out); }"]
    path = os.path.join(GEN, "prog_background.cpp")
    with open(path, "w") as f:
        f.write("\n".join(text) + "\n")
    files.append(path)
    return files


def build(verbose=False):
    if not os.path.isdir(REF):
        raise SystemExit("reference sources not present at %s (oracle/_ref can only be built where they are)" % REF)
    srcs = generate()
    srcs.append(os.path.join(HERE, "ref_common.cpp"))
    host = os.path.join(HERE, "ref_host_driver.cpp")
    extra = []
    if os.path.exists(host):
        extra = build_host_objects(verbose)
    objs = []
    for s in srcs:
        o = os.path.join(GEN, os.path.basename(s) + ".o")
        cmd = ["g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-w", "-I", HERE, "-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    # -Bsymbolic: VV/3DLIC.cpp defines unmangled globals (w, h, aspect, light, noise, ...); bind the library's own references to
    # them whatever else the process has loaded with RTLD_GLOBAL
    cmd = ["g++", "-shared", "-fopenmp", "-Wl,-Bsymbolic", "-o", LIB] + objs + extra
    subprocess.check_call(cmd)
    return LIB


def build_host_objects(verbose=False):
    """reference C++ translation units, unmodified, against the capturing GL stub in oracle/ref_shim/"""
    shim = os.path.join(HERE, "ref_shim")
    objs = []
    units = ["mmath.cpp", "reader.cpp", "parseArg.cpp", "gradient.cpp", "dataset.cpp", "transferEdit.cpp", "texture.cpp", "illumination.cpp", "slicing.cpp",
             "transform.cpp", "trackball.cpp", "camera.cpp", "VolumeBuffer.cpp", "renderer.cpp", "GLSLShader.cpp",
             "fpsCounter.cpp", "timer.cpp", "3DLIC.cpp"]
    flags = ["-O1", "-std=c++14", "-fPIC", "-w", "-fpermissive", "-ffp-contract=off", "-DGLEW_NO_GLU", "-D_USE_MATH_DEFINES",
             "-include", "climits", "-include", "cstring", "-include", "cstdlib", "-include", os.path.join(shim, "ref_prelude.h"),
             "-I", shim, "-I", REF]
    for u in units:
        src = os.path.join(REF, u)
        if not os.path.exists(src):
            continue
        o = os.path.join(GEN, "host_" + u + ".o")
        cmd = ["g++"] + flags + ["-c", src, "-o", o]
        if u == "GLSLShader.cpp":      # `return false;` from a char* function (VV/GLSLShader.cpp): valid only before C++11
            cmd = ["g++"] + [("-std=gnu++98" if f == "-std=c++14" else f) for f in flags] + ["-c", src, "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    for drv in ("ref_host_driver", "ref_app_driver"):
        o = os.path.join(GEN, drv + ".o")
        subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HERE, drv + ".cpp"), "-o", o])
        objs.append(o)
    return objs


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
