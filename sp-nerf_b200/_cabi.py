"""ctypes binding of libspnerf_sm100a.so (the C ABI declared in include/spnerf_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspnerf_sm100a.so")

SELFTEST_MAX_KSTEPS = 64


class SpnerfError(RuntimeError):
    pass


class UmmaSelftest(ctypes.Structure):
    _fields_ = [
        ("a_img", ctypes.c_void_p),
        ("b_img", ctypes.c_void_p),
        ("d_out", ctypes.c_void_p),
        ("a_bytes", ctypes.c_uint32),
        ("b_bytes", ctypes.c_uint32),
        ("n", ctypes.c_uint32),
        ("ksteps", ctypes.c_uint32),
        ("idesc", ctypes.c_uint32),
        ("_pad", ctypes.c_uint32),
        ("a_desc_template", ctypes.c_uint64),
        ("b_desc_template", ctypes.c_uint64),
        ("a_off", ctypes.c_uint32 * SELFTEST_MAX_KSTEPS),
        ("b_off", ctypes.c_uint32 * SELFTEST_MAX_KSTEPS),
    ]


_lib = None


def lib():
    """Load the shared library once.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpnerfError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.spnerf_abi_version.restype = ctypes.c_int
    L.spnerf_abi_version.argtypes = []
    L.spnerf_watchdog_code.restype = ctypes.c_uint
    L.spnerf_watchdog_code.argtypes = []
    L.spnerf_selftest_umma.restype = ctypes.c_int
    L.spnerf_selftest_umma.argtypes = [ctypes.POINTER(UmmaSelftest), ctypes.c_void_p]


def check(rc, what):
    if rc == 0:
        return
    if rc > 0:
        names = {1: "bad argument", 2: "unsupported configuration", 3: "workspace too small"}
        raise SpnerfError(f"{what}: {names.get(rc, rc)}")
    raise SpnerfError(f"{what}: CUDA error {-rc}")
