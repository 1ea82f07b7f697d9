"""Build a VARIANT of libsstg.so for A/B runs (scripts/ab_libs.py): the kernels of another git revision
and/or extra -D flags, without touching the in-tree library.

    python scripts/build_variant.py --out _scratch/libsstg_head.so [--ref HEAD] [-DSSTG_FOO ...]

--ref takes social_stgcnn_b200/csrc and include/ from that revision (default: the working tree).
"""
import argparse
import concurrent.futures
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from social_stgcnn_b200.build import NVCC_FLAGS, SOURCES, _nvcc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--ref", default=None)
a, extra = ap.parse_known_args()
nv = _nvcc()
with tempfile.TemporaryDirectory() as tmp:
    if a.ref:
        subprocess.check_call(f"git -C {ROOT} archive {a.ref} social_stgcnn_b200/csrc include | tar -x -C {tmp}", shell=True)
        src_root = tmp
    else:
        src_root = ROOT
    csrc = os.path.join(src_root, "social_stgcnn_b200", "csrc")
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(csrc, s))]

    def one(src):
        o = os.path.join(tmp, src.replace(".cu", ".o"))
        r = subprocess.run([nv] + NVCC_FLAGS + extra + ["-c", os.path.join(csrc, src), "-o", o], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)
        return o

    with concurrent.futures.ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(one, srcs))
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    subprocess.check_call([nv, "-shared", "-o", a.out] + objs + ["-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"],
                          stderr=subprocess.DEVNULL)
print(a.out)
