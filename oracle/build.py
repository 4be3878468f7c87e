"""Build recipe for the oracle's C pieces (TEST INFRASTRUCTURE ONLY).

``python -m oracle.build`` compiles ``oracle/nms_v5.c`` (the restatement of TF's
NonMaxSuppressionV5 CPU kernel) into ``oracle/_build/liboracle.so`` with plain gcc.

The reference itself is pure Python on TensorFlow 2.10 - it contains no C/C++ sources to
compile, so there is no ``oracle/_ref`` binary for this repository (see DESIGN.md "Oracle").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SOURCES = [os.path.join(HERE, "nms_v5.c")]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB):
        newest = max(os.path.getmtime(s) for s in SOURCES)
        if os.path.getmtime(LIB) >= newest:
            return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", LIB]
    cmd += SOURCES + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
