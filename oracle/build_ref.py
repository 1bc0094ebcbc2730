"""ORACLE (test infrastructure).  Byte-compiles the reference's own pure-Python sources, where they lie under
/root/reference, into oracle/_ref/cellregmap/*.bin (CPython byte code, i.e. the content of a .pyc file under an extension the snapshot tools do not
filter out) -- compiled outputs only, no source is copied.  oracle/_ref is
git-ignored (not gpurun-ignored), so the compiled reference travels to the GPU box, where `oracle.ref_shims.load_reference()`
imports it (sourceless) over the dependency stand-ins: the checker of the `-m gpu` tests and the CPU arm of bench.py then run
the reference's unmodified logic.

    python -m oracle.build_ref          # no-op (returns False) when /root/reference is absent
"""
import os
import py_compile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/cellregmap"
OUT = os.path.join(HERE, "_ref", "cellregmap")
FILES = ("__init__.py", "_cellregmap.py", "_math.py", "_types.py", "_simulate.py")


def build(force=False):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(OUT, exist_ok=True)
    for name in FILES:
        src = os.path.join(SRC, name)
        dst = os.path.join(OUT, name[:-3] + ".bin")
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            # unchecked-hash pycs: valid wherever the tree is copied (no source file to compare time stamps with)
            py_compile.compile(src, cfile=dst, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return True


if __name__ == "__main__":
    print("built" if build(force=True) else "reference sources not present; nothing built", OUT)
