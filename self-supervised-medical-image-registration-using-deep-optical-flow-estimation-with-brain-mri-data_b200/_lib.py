"""Loader (and builder) of the C-ABI library ``libofr_b200.so`` declared in ``include/ofr_b200.h``.

The product path calls into the library through ctypes with raw device pointers and the current
CUDA stream handle.  There is no fallback: if the library is missing or a call returns non-zero a
``RuntimeError`` is raised.
"""
import ctypes
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# OFR_B200_LIB: development switch of the PYTHON loader (A/B runs of kernel variants built by tools/build_variant.py);
# the library itself reads no environment variable
LIB_PATH = os.environ.get("OFR_B200_LIB") or os.path.join(PKG_DIR, "libofr_b200.so")
SOURCES = ["abi.cu", "channelnorm.cu", "resample2d.cu", "correlation.cu", "correlation_tma.cu", "correlation_umma.cu", "loss.cu", "flowhead.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xptxas=-v", "-Xcompiler", "-fPIC", "-shared"]

OFR_OK = 0
WARP_FLOWNET2, WARP_STN, WARP_PWC, WARP_FLOWNET2_NEAREST = 0, 1, 2, 3
LOSS_MAX_SCALES = 8
LOSS_STATE_DOUBLES = 16

_lock = threading.Lock()
_lib = None


class LossScale(ctypes.Structure):
    """Mirror of ``ofr_loss_scale_t``."""
    _fields_ = [("warped", ctypes.c_void_p), ("flow", ctypes.c_void_p),
                ("g_warped", ctypes.c_void_p), ("g_flow", ctypes.c_void_p),
                ("h", ctypes.c_int), ("w", ctypes.c_int)]


class StnScale(ctypes.Structure):
    """Mirror of ``ofr_stn_scale_t``."""
    _fields_ = [("frame", ctypes.c_void_p), ("flow", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("g_out", ctypes.c_void_p), ("g_flow", ctypes.c_void_p), ("h", ctypes.c_int), ("w", ctypes.c_int)]


STN_MAX_SCALES = 8


class RegLossScale(ctypes.Structure):
    """Mirror of ``ofr_regloss_scale_t``."""
    _fields_ = [("moving", ctypes.c_void_p), ("flow", ctypes.c_void_p), ("warped", ctypes.c_void_p),
                ("g_flow", ctypes.c_void_p), ("h", ctypes.c_int), ("w", ctypes.c_int)]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libofr_b200.so")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(os.path.dirname(PKG_DIR), "include", "ofr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, jobs=None, defines=(), out=None):
    """Compile every kernel for sm_100a into one shared library, in-tree (one nvcc per source file, in parallel, then
    one link).  ``defines`` / ``out``: build a variant (extra -D flags) into another file (tools/build_variant.py)."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build") if out is None else os.path.join(PKG_DIR, "build", os.path.basename(out) + ".d")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-D" + d for d in defines]
    target = out or LIB_PATH

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + compile_flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        return src, obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=jobs or min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, cmd, res in results:
        if verbose or res.returncode != 0:
            print(" ".join(cmd))
            print(res.stdout)
            print(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed compiling {src}:\n" + res.stderr[-4000:])
    tmp = target + ".tmp"
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + \
          [obj for _, obj, _, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libofr_b200.so:\n" + res.stderr[-4000:])
    os.replace(tmp, target)
    return target


def _declare(lib):
    c_int, c_size_t, vp, dbl = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_double
    ip = ctypes.POINTER(ctypes.c_int)
    sig = {
        "ofr_abi_version": (c_int, []),
        "ofr_error_string": (ctypes.c_char_p, [c_int]),
        "ofr_launch_count": (ctypes.c_uint64, []),
        "ofr_corr_out_shape": (c_int, [c_int] * 7 + [ip, ip, ip]),
        "ofr_corr_workspace": (c_size_t, [c_int] * 10),
        "ofr_corr_fwd": (c_int, [vp, vp, vp] + [c_int] * 9 + [vp, c_size_t, vp]),
        "ofr_corr_fwd_act": (c_int, [vp, vp, vp] + [c_int] * 9 + [ctypes.c_float, vp, c_size_t, vp]),
        "ofr_corr_bwd": (c_int, [vp, vp, vp, vp, vp] + [c_int] * 9 + [vp, c_size_t, vp]),
        "ofr_resample_fwd": (c_int, [vp, vp, vp] + [c_int] * 7 + [vp]),
        "ofr_resample_fwd_scaled": (c_int, [vp, vp, vp] + [c_int] * 7 + [ctypes.c_float, vp]),
        "ofr_resample_bwd_scaled": (c_int, [vp, vp, vp, vp, vp] + [c_int] * 7 + [ctypes.c_float, vp, c_size_t, vp]),
        "ofr_resample_bwd_workspace": (c_size_t, [c_int] * 7),
        "ofr_resample_bwd": (c_int, [vp, vp, vp, vp, vp] + [c_int] * 7 + [vp, c_size_t, vp]),
        "ofr_stn_multi_fwd": (c_int, [ctypes.POINTER(StnScale), c_int, c_int, c_int, vp]),
        "ofr_stn_multi_bwd": (c_int, [ctypes.POINTER(StnScale), c_int, c_int, c_int, vp]),
        "ofr_chnorm_fwd": (c_int, [vp, vp] + [c_int] * 4 + [vp]),
        "ofr_chnorm_bwd": (c_int, [vp, vp, vp, vp] + [c_int] * 4 + [vp]),
        "ofr_chnorm_diff_fwd": (c_int, [vp, vp, vp] + [c_int] * 4 + [vp]),
        "ofr_chnorm_diff_bwd": (c_int, [vp, vp, vp, vp, vp, vp] + [c_int] * 4 + [vp]),
        "ofr_loss_workspace": (c_size_t, [c_int]),
        "ofr_loss_fwd": (c_int, [vp, c_int, c_int, c_int, ctypes.POINTER(LossScale), c_int, dbl, dbl, dbl,
                                 vp, vp, vp, c_size_t, vp]),
        "ofr_loss_bwd": (c_int, [vp, c_int, c_int, c_int, ctypes.POINTER(LossScale), c_int, dbl, dbl, dbl,
                                 vp, vp, vp]),
        "ofr_regloss_fwd": (c_int, [vp, c_int, c_int, c_int, ctypes.POINTER(RegLossScale), c_int, dbl, dbl, dbl,
                                    vp, vp, vp, c_size_t, vp]),
        "ofr_regloss_bwd": (c_int, [vp, c_int, c_int, c_int, ctypes.POINTER(RegLossScale), c_int, dbl, dbl, dbl,
                                    vp, vp, vp]),
        "ofr_flowconv3_workspace": (c_size_t, [c_int] * 5),
        "ofr_flowconv3_fwd": (c_int, [vp, vp, vp, vp] + [c_int] * 4 + [vp, c_size_t, vp]),
        "ofr_flowconv3_bwd": (c_int, [vp, vp, vp, vp, vp, vp] + [c_int] * 4 + [vp, c_size_t, vp]),
        "ofr_flowdeconv4_workspace": (c_size_t, [c_int] * 5),
        "ofr_flowdeconv4_fwd": (c_int, [vp, vp, vp, vp] + [c_int] * 4 + [vp, c_size_t, vp]),
        "ofr_flowdeconv4_bwd": (c_int, [vp, vp, vp, vp, vp, vp] + [c_int] * 4 + [vp, c_size_t, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return sorted(sig)


EXPORTS = ["ofr_abi_version", "ofr_error_string", "ofr_launch_count", "ofr_corr_out_shape", "ofr_corr_workspace",
           "ofr_corr_fwd", "ofr_corr_fwd_act",
           "ofr_corr_bwd", "ofr_resample_fwd", "ofr_resample_fwd_scaled", "ofr_resample_bwd_workspace", "ofr_resample_bwd",
           "ofr_resample_bwd_scaled", "ofr_stn_multi_fwd", "ofr_stn_multi_bwd",
           "ofr_chnorm_fwd", "ofr_chnorm_bwd", "ofr_chnorm_diff_fwd", "ofr_chnorm_diff_bwd", "ofr_loss_workspace", "ofr_loss_fwd", "ofr_loss_bwd",
           "ofr_regloss_fwd", "ofr_regloss_bwd", "ofr_flowconv3_workspace", "ofr_flowconv3_fwd", "ofr_flowconv3_bwd",
           "ofr_flowdeconv4_workspace", "ofr_flowdeconv4_fwd", "ofr_flowdeconv4_bwd"]


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                        "There is no CPU or PyTorch fallback for the ofr_b200 layers.")
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def check(code, what):
    if code != OFR_OK:
        msg = lib().ofr_error_string(code).decode()
        raise RuntimeError(f"{what} failed with code {code}: {msg}")


def launch_count():
    return int(lib().ofr_launch_count())
