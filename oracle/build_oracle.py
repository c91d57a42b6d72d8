"""Build the oracle's C restatements with gcc (test infrastructure; never linked into the product)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")


def build():
    os.makedirs(OUT, exist_ok=True)
    libs = {}
    for real, name in (("double", "libctc_ref_f64.so"), ("float", "libctc_ref_f32.so")):
        src = os.path.join(HERE, "ctc_ref.c")
        out = os.path.join(OUT, name)
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-DREAL=" + real, "-o", out, src, "-lm"])
        libs[real] = out
    return libs


if __name__ == "__main__":
    print(build())
