"""Build the oracle's C restatements with gcc (test infrastructure; never linked into the product) and, where the
reference tree is present (the authoring container), stage the UNMODIFIED reference modules of the hot path under
oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot like a built .so) so that bench.py's CPU arm can
run the reference's own CnnOcrModel / ArgmaxDecoder there.  Nothing under oracle/_ref is ever committed or imported
by the product."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
REF_SRC = "/root/reference/src"
REF_OUT = os.path.join(HERE, "_ref")
REF_FILES = ("models/cnnlstm.py", "decoder.py", "alphabet.py")  # hot-path modules only


def stage_reference():
    """Copies the three reference modules byte for byte (a build artefact, like compiling a C reference from its
    sources where they lie).  Returns the staging directory or None when the reference tree is absent."""
    if not os.path.isdir(REF_SRC):
        return REF_OUT if os.path.exists(os.path.join(REF_OUT, "decoder.py")) else None
    for rel in REF_FILES:
        dst = os.path.join(REF_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    return REF_OUT


def build():
    os.makedirs(OUT, exist_ok=True)
    libs = {}
    for real, name in (("double", "libctc_ref_f64.so"), ("float", "libctc_ref_f32.so")):
        src = os.path.join(HERE, "ctc_ref.c")
        out = os.path.join(OUT, name)
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-DREAL=" + real, "-o", out, src, "-lm"])
        libs[real] = out
    stage_reference()
    return libs


if __name__ == "__main__":
    print(build())
