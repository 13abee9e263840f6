"""oracle/build_ref_nms.py -- compile the REFERENCE's own Cython NMS (lib/utils/cython_nms.pyx) into
oracle/_ref/ so that the NMS restatement (oracle/nms_oracle.py) can be pinned against it.

The .pyx is read where it lies under /root/reference; nothing is copied into the repository (oracle/_ref/ is
git-ignored).  The only change is the spelling of two numpy type names that numpy >= 1.24 removed
(np.int_t -> np.intp_t, dtype=np.int -> dtype=np.intp): the arithmetic is untouched.
Run by `make -C oracle` when the reference tree is present."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CIM_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "lib", "utils", "cython_nms.pyx")
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.exists(SRC):
        print("reference tree not present: keeping prebuilt oracle/_ref/ref_cython_nms (if any)")
        return 0
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    pyx = os.path.join(OUT, "ref_cython_nms.pyx")
    text = open(SRC).read().replace("np.int_t", "np.intp_t").replace("dtype=np.int)", "dtype=np.intp)")
    open(pyx, "w").write(text)
    c_file = os.path.join(OUT, "ref_cython_nms.c")
    subprocess.run([sys.executable, "-m", "cython", "-2", pyx, "-o", c_file], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    so = os.path.join(OUT, "ref_cython_nms" + sysconfig.get_config_var("EXT_SUFFIX"))
    cc = os.environ.get("CC", "gcc")
    subprocess.run([cc, "-O2", "-fPIC", "-shared", "-w", "-fwrapv", "-I" + sysconfig.get_paths()["include"],
                    "-I" + np.get_include(), c_file, "-o", so], check=True)
    os.remove(pyx)
    os.remove(c_file)
    print("built", os.path.relpath(so, os.path.dirname(HERE)), "from", SRC)
    return 0


if __name__ == "__main__":
    sys.exit(main())
