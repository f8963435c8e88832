"""Build a kernel-experiment variant of the library: build_variant.py NAME [-DMACRO ...]
-> sparse_rcnn_b200/csrc/variants/libscn_NAME.so; run with SCN_B200_LIB=<that path>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparse_rcnn_b200 import build as B
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.CSRC, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, "libscn_%s.so" % name)
cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + flags + ["-o", out] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
subprocess.run(cmd, check=True)
print(out)
