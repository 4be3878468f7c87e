"""Builds libudal.so (hand-written CUDA for sm_100a + the C ABI of include/udal.h) in-tree.

    python uncertainty-detection-autolabeling_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The shared library lands next to this file so that it
travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libudal.so")
SOURCES = ["api.cu", "decode_moments.cu", "topk.cu", "nms.cu", "nms_cta.cu", "post.cu", "heads_fp32.cu",
           "heads_tc.cu", "heads_ig.cu", "heads_dw.cu", "heads_l1.cu", "heads_fused.cu", "heads_wide.cu", "run.cu", "nms_np.cu", "autolabel.cu", "decode_sample.cu", "bifpn.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "udal.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        src_path = os.path.join(CSRC, src)
        headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
        headers.append(os.path.join(ROOT, "include", "udal.h"))
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src_path)
                and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in headers)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src_path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    # --no-undefined: a symbol declared with the wrong linkage must fail here, not when ctypes loads the library
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs +
                          ["-lcudart_static", "-ldl", "-lrt", "-lpthread", "-Xlinker", "--no-undefined"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
