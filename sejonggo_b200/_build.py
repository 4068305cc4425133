"""Builds the in-tree CUDA library (sm_100a only) with nvcc.

    python -m sejonggo_b200._build [--force]

The .so lives in sejonggo_b200/lib/ (git-ignored, travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libsejonggo_b200.so")
SOURCES = ["rules.cu", "tree.cu", "tower.cu", "driver.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false",          # PUCT / value sums must not be contracted; tower kernels use explicit fma/mma
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.isfile(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sejonggo_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


ABLATE_SO = os.path.join(LIBDIR, "libsejonggo_b200_ablate.so")
WIDETILES_SO = os.path.join(LIBDIR, "libsejonggo_b200_widetiles.so")      # A/B: tower layers on conv_wide.cuh's 512-row super-tiles (--wide-tiles)


def build(force=False, verbose=False, ablate=False, wide_tiles=False):
    """ablate=True builds a SECOND library with -DSGO_CONV_ABLATE (the conv kernel's timing-ablation switches read
    from SGO_CONV_DEBUG; tools/conv_variants.py loads it through SGO_LIBRARY); wide_tiles=True one whose tower layers run
    on conv_wide.cuh (a measured alternative, profiles/r02_conv_wide_tiles_ab.json).  The product library has neither."""
    out = ABLATE_SO if ablate else (WIDETILES_SO if wide_tiles else SO)
    extra = os.environ.get("SGO_NVCC_EXTRA", "").split()          # ad-hoc A/B variants: SGO_NVCC_EXTRA="-DPR_COALESCED_STORE=1" (or -DSGO_STEM_SEPARATE) SGO_BUILD_OUT=<path>
    if extra:
        out = os.environ.get("SGO_BUILD_OUT") or os.path.join(LIBDIR, "libsejonggo_b200_variant.so")
    if not force and not ablate and not wide_tiles and not extra and not _stale():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    cmd = [nvcc] + NVCC_FLAGS + (["-DSGO_CONV_ABLATE"] if ablate else []) + (["-DSGO_CONV_WIDE_TILES"] if wide_tiles else []) + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ablate="--ablate" in sys.argv, wide_tiles="--wide-tiles" in sys.argv))
