"""Build libvv_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m vectorvisualization_b200.build [--force] [--verbose]

The shared library lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvv_b200.so")
VOLIC = os.path.join(HERE, "volic")

CU_SOURCES = ["vv_kernels.cu", "vv_preprocess.cu", "vv_renderer.cu"]
CPP_SOURCES = ["vv_io.cpp", "vv_keys.cpp"]
OPTIONAL_CPP = ["vv_illum.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "vv_c_api.h"))
    cpp = list(CPP_SOURCES) + [f for f in OPTIONAL_CPP if os.path.exists(os.path.join(CSRC, f))]
    defs = ["-DVV_HAVE_ILLUM_TABLES"] if "vv_illum.cpp" in cpp else []
    objs, jobs = [], []
    for src in CU_SOURCES + cpp:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers + [os.path.abspath(__file__)]):
            cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if src.endswith(".cpp"):
                cmd = [nvcc] + NVCC_FLAGS + defs + ["-x", "cu", "-c", s, "-o", o]
            jobs.append(cmd)
    if jobs:
        # the translation units are independent: compile them side by side (vv_kernels.cu alone takes ~3 minutes)
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lz", "-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    main = os.path.join(CSRC, "volic_main.cpp")
    if os.path.exists(main) and (force or _stale(VOLIC, [main, LIB])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), main, "-o", VOLIC,
                               "-L", HERE, "-lvv_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


def build_variant(tag, defines=(), extra_flags=(), only=("vv_kernels.cu",)):
    """A/B experiments: build libvv_b200_<tag>.so next to the shipped library from the same sources with extra -D defines /
    nvcc flags (objects under build/<tag>/).  Only the sources in `only` are recompiled with the extra flags (the kernel
    variants live in vv_kernels.cu); the other objects are those of the shipped build.  Select the library at run time with
    VV_B200_LIB=<path> (scripts/ab.py, scripts/profile_frame.py, tests):

        python -m vectorvisualization_b200.build --variant regs80 -DLIC_MIN_CTAS=3
        python scripts/ab.py cfg=cfg3 vectorvisualization_b200/libvv_b200.so vectorvisualization_b200/libvv_b200_regs80.so
    """
    build()
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    cpp = list(CPP_SOURCES) + [f for f in OPTIONAL_CPP if os.path.exists(os.path.join(CSRC, f))]
    defs = (["-DVV_HAVE_ILLUM_TABLES"] if "vv_illum.cpp" in cpp else []) + list(defines) + list(extra_flags)
    objs = []
    for src in CU_SOURCES + cpp:
        if src not in only:
            objs.append(os.path.join(HERE, "build", src + ".o"))
            continue
        o = os.path.join(objdir, src + ".o")
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + defs + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", os.path.join(CSRC, src), "-o", o]
        subprocess.check_call(cmd)
    lib = os.path.join(HERE, "libvv_b200_%s.so" % tag)
    subprocess.check_call([nvcc, "-shared", "-o", lib] + objs + ["-lz", "-gencode", "arch=compute_100a,code=sm_100a"])
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")],
                            [a for a in sys.argv[i + 2:] if not a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
