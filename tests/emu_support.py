"""TEST INFRASTRUCTURE — host-side emulation build of the CUDA sources (tests/cuda_emu).

`emu_package()` returns the package with its ctypes handle pointing at tests/cuda_emu/_build/libtopopt_emu.so: the very same
.cu sources compiled for fibers on the host (see tests/cuda_emu/cuda_emu.h).  It checks kernel *logic* where no GPU
exists — indexing, initialisation (allocations come back 0xFF-filled), barriers, the mbarrier pipeline protocol, CG
recurrences, partition / interface maps — at toy sizes.  It is never a product path: the package itself only ever loads
libtopopt_b200.so and fails loudly without a B200; nothing here is timed or shipped."""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cuda_emu")
# TOE_EMU_ASAN=1: AddressSanitizer build in its own object directory; run the tests with
#   LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 TOE_EMU_ASAN=1 python -m pytest tests/test_emu_*.py
ASAN = os.environ.get("TOE_EMU_ASAN") == "1"
UBSAN = os.environ.get("TOE_EMU_UBSAN") == "1"      # -fsanitize=undefined (no preload needed): TOE_EMU_UBSAN=1 python -m pytest tests/test_emu_*.py
OBJDIR = "_asan" if ASAN else ("_ubsan" if UBSAN else "_build")
EMU_LIB = os.path.join(EMU_DIR, OBJDIR, "libtopopt_emu.so")


def build_emu(extra: str = "", opt: str = "-O2") -> str:
    if ASAN:
        extra = (extra + " -fsanitize=address -fno-omit-frame-pointer").strip()
        opt = "-O1"
    elif UBSAN:
        extra = (extra + " -fsanitize=undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer").strip()
        opt = "-O1"
    cxx = ["CXX=/usr/bin/g++"] if (ASAN or UBSAN) and os.path.exists("/usr/bin/g++") else []      # the distribution's g++ ships libasan
    env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}                    # the sanitizer runtime is for the tests, not for make / g++
    res = subprocess.run(["make", "-C", EMU_DIR, "-j8", "OPT=" + opt, "OBJDIR=" + OBJDIR] + cxx + (["EXTRA=" + extra] if extra else []),
                         capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("building libtopopt_emu.so failed:\n" + res.stdout[-3000:] + res.stderr[-6000:])
    return EMU_LIB


def load_emu():
    # threads of a block are resumed in a pseudo-random order (fixed seed): stricter than 0,1,2,… and still reproducible
    os.environ.setdefault("EMU_SHUFFLE", "20261018")
    build_emu()
    import __graft_entry__ as graft
    pkg = graft.load_package()
    lib = C.CDLL(EMU_LIB, mode=C.RTLD_LOCAL)
    for name, (res, args) in pkg._lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.emu_peak_bytes.restype = C.c_size_t
    lib.emu_live_allocations.restype = C.c_size_t
    lib.emu_check_all_guards.restype = C.c_int
    return pkg, lib


@contextlib.contextmanager
def emulated(pkg, lib):
    """Inside the block `pkg.Context` / the api functions drive the emulated library (tests only)."""
    saved = pkg._lib._lib
    pkg._lib._lib = lib
    try:
        yield pkg
    finally:
        pkg._lib._lib = saved
