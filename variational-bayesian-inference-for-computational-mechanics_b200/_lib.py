"""ctypes binding of libvbfem.so (include/vbfem.h).

The library is built in-tree (csrc/libvbfem.so) with nvcc for sm_100a and is the
ONLY compute path: there is no CPU fallback.  ``load()`` raises if the shared
object is missing, and every entry point raises ``VbfemError`` with the
library's own message when a call fails (e.g. no CUDA device).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
REPO = os.path.dirname(_HERE)
INCLUDE = os.path.join(REPO, "include")
LIB_PATH = os.environ.get("VBFEM_LIB", os.path.join(CSRC, "libvbfem.so"))  # VBFEM_LIB: profiling builds
SOURCES = ["vbfem.cu"]
HEADERS = ["vbfem_math.cuh", "vbfem_front.cuh", "vbfem_front_kernel.cuh", "vbfem_panel.cuh", "vbfem_panel2.cuh",
           "vbfem_warp.cuh", "vbfem_warp2.cuh", "vbfem_peer.cuh",
           os.path.join(INCLUDE, "vbfem.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550",
]

INFO_COUNT = 16
INFO_NAMES = ["nfree", "half_bw", "ndof", "nele", "ncolors", "band_in_smem", "smem_bytes", "ctas_per_sm",
              "num_sms", "block_threads", "kernel_variant", "twist_row", "panel_blocks", "panel_ring"]

# every symbol include/vbfem.h declares
SYMBOLS = [
    "vbfem_create", "vbfem_create_ex", "vbfem_fields_elementwise", "vbfem_plan", "vbfem_destroy", "vbfem_last_error", "vbfem_info", "vbfem_reserve", "vbfem_forward",
    "vbfem_backward", "vbfem_keep_ticket", "vbfem_backward_ticket", "vbfem_forward_jac", "vbfem_jac_vjp",
    "vbfem_debug_panel_tables", "vbfem_forward_backward", "vbfem_fields", "vbfem_elbo_step1", "vbfem_elbo_step2", "vbfem_status", "vbfem_forward_host",
    "vbfem_forward_backward_host", "vbfem_measure_peaks",
    "vbfem_peer_open", "vbfem_peer_connect", "vbfem_peer_allreduce", "vbfem_peer_status",
    "vbfem_elbo_step1_allreduce", "vbfem_elbo_step2_allreduce", "vbfem_elbo_step1_loss",
]


class VbfemError(RuntimeError):
    pass


class VbfemMesh(ctypes.Structure):
    """struct vbfem_mesh (include/vbfem.h)."""
    _fields_ = [
        ("nnodes", ctypes.c_int32),
        ("nele", ctypes.c_int32),
        ("coord", ctypes.POINTER(ctypes.c_double)),
        ("ien", ctypes.POINTER(ctypes.c_int32)),
        ("nfree", ctypes.c_int32),
        ("free_dof", ctypes.POINTER(ctypes.c_int32)),
        ("pf", ctypes.POINTER(ctypes.c_double)),
        ("thk", ctypes.c_double),
        ("obs_node", ctypes.c_int32),
        ("obs_ele", ctypes.c_int32),
        ("obs_gp", ctypes.c_int32 * 2),
        ("theta_mean", ctypes.c_double * 2),
        ("theta_std", ctypes.c_double * 2),
    ]


class VbfemOptions(ctypes.Structure):
    """struct vbfem_options (include/vbfem.h)."""
    _fields_ = [("stype", ctypes.c_int32), ("reserved", ctypes.c_int32 * 7)]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libvbfem.so for sm_100a (nvcc cross-compiles
    without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise VbfemError("nvcc not found: cannot build libvbfem.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise VbfemError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def load():
    """dlopen csrc/libvbfem.so and declare the C signatures.  Fails loudly when
    the library has not been built (``python -c 'import __graft_entry__ as g;
    g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VbfemError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                         "this package has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    c_dp = ctypes.c_void_p  # device or host pointers travel as integers
    i64 = ctypes.c_int64
    lib.vbfem_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(VbfemMesh), ctypes.c_int]
    lib.vbfem_create.restype = ctypes.c_int
    lib.vbfem_create_ex.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(VbfemMesh),
                                    ctypes.POINTER(VbfemOptions), ctypes.c_int]
    lib.vbfem_create_ex.restype = ctypes.c_int
    lib.vbfem_fields_elementwise.argtypes = [ctypes.c_void_p, i64] + [c_dp] * 8
    lib.vbfem_fields_elementwise.restype = ctypes.c_int
    lib.vbfem_plan.argtypes = [ctypes.POINTER(VbfemMesh), i64, ctypes.POINTER(i64)]
    lib.vbfem_plan.restype = ctypes.c_int
    lib.vbfem_destroy.argtypes = [ctypes.c_void_p]
    lib.vbfem_destroy.restype = None
    lib.vbfem_last_error.argtypes = []
    lib.vbfem_last_error.restype = ctypes.c_char_p
    lib.vbfem_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(i64)]
    lib.vbfem_info.restype = ctypes.c_int
    lib.vbfem_forward.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, ctypes.c_int, c_dp]
    lib.vbfem_forward.restype = ctypes.c_int
    lib.vbfem_backward.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_backward.restype = ctypes.c_int
    lib.vbfem_reserve.argtypes = [ctypes.c_void_p, i64]
    lib.vbfem_reserve.restype = ctypes.c_int
    lib.vbfem_keep_ticket.argtypes = [ctypes.c_void_p]
    lib.vbfem_keep_ticket.restype = i64
    lib.vbfem_backward_ticket.argtypes = [ctypes.c_void_p, i64, i64, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_backward_ticket.restype = ctypes.c_int
    lib.vbfem_forward_jac.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_forward_jac.restype = ctypes.c_int
    lib.vbfem_jac_vjp.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_jac_vjp.restype = ctypes.c_int
    lib.vbfem_debug_panel_tables.argtypes = [ctypes.POINTER(VbfemMesh), ctypes.c_int, c_dp, i64]
    lib.vbfem_debug_panel_tables.restype = i64
    lib.vbfem_forward_backward.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_forward_backward.restype = ctypes.c_int
    lib.vbfem_fields.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_fields.restype = ctypes.c_int
    lib.vbfem_elbo_step1.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, i64, i64, c_dp, c_dp, c_dp,
                                     c_dp, ctypes.c_double, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_elbo_step1.restype = ctypes.c_int
    lib.vbfem_elbo_step2.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, i64, i64, c_dp, c_dp, c_dp,
                                     c_dp, c_dp, c_dp]
    lib.vbfem_elbo_step2.restype = ctypes.c_int
    lib.vbfem_elbo_step1_allreduce.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, i64, i64, c_dp, c_dp,
                                               c_dp, c_dp, ctypes.c_double, c_dp, c_dp, c_dp]
    lib.vbfem_elbo_step1_allreduce.restype = ctypes.c_int
    lib.vbfem_elbo_step1_loss.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, i64, i64, c_dp, c_dp, c_dp,
                                          c_dp, c_dp, ctypes.c_double, ctypes.c_int32, c_dp, c_dp]
    lib.vbfem_elbo_step1_loss.restype = ctypes.c_int
    lib.vbfem_elbo_step2_allreduce.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, i64, i64, c_dp, c_dp,
                                               c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_elbo_step2_allreduce.restype = ctypes.c_int
    lib.vbfem_peer_open.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_dp,
                                    ctypes.POINTER(ctypes.c_void_p)]
    lib.vbfem_peer_open.restype = ctypes.c_int
    lib.vbfem_peer_connect.argtypes = [ctypes.c_void_p, c_dp, ctypes.POINTER(ctypes.c_void_p)]
    lib.vbfem_peer_connect.restype = ctypes.c_int
    lib.vbfem_peer_allreduce.argtypes = [ctypes.c_void_p, c_dp, ctypes.c_int32, c_dp]
    lib.vbfem_peer_allreduce.restype = ctypes.c_int
    lib.vbfem_peer_status.argtypes = [ctypes.c_void_p]
    lib.vbfem_peer_status.restype = i64
    lib.vbfem_status.argtypes = [ctypes.c_void_p, c_dp, i64]
    lib.vbfem_status.restype = i64
    lib.vbfem_forward_host.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp]
    lib.vbfem_forward_host.restype = ctypes.c_int
    lib.vbfem_forward_backward_host.argtypes = [ctypes.c_void_p, i64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    lib.vbfem_forward_backward_host.restype = ctypes.c_int
    lib.vbfem_measure_peaks.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(ctypes.c_double)]
    lib.vbfem_measure_peaks.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "libvbfem call"):
    if rc < 0:
        msg = load().vbfem_last_error()
        raise VbfemError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")
    return rc
