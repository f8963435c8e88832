"""In-tree build of the C-ABI CUDA library (explicit nvcc, sm_100a only).

    python -m sparse_rcnn_b200.build [--force]

Output: sparse_rcnn_b200/csrc/libscn_b200.so (git-ignored, travels to the GPU box with the
snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["grid.cu", "sort.cu", "ops.cu", "conv_fp32.cu", "conv_tc.cu", "conv_ts.cu", "conv_wgrad_tc.cu", "conv_wgrad_ts.cu", "fused.cu", "unet_exec.cu", "dense.cu", "nms.cu", "voxelize.cu"]
OUT = os.path.join(CSRC, "libscn_b200.so")
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "tile_book.cuh"),
                                                      os.path.join(os.path.dirname(HERE), "include", "scn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = OUT + ".tmp%d" % os.getpid()      # link beside the target, then rename: a reader never sees a partial library
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, OUT)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
