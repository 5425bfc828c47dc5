"""ctypes binding of librsoccer_b200.so (the C ABI in include/rsoccer_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is
no CPU fallback and nothing here touches ``oracle/``: if the shared object is missing,
importing the engine fails loudly.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# RS_LIB: load another build of the same sources (kernel tuning experiments only)
LIB_PATH = os.environ.get("RS_LIB") or os.path.join(_HERE, "librsoccer_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("rs_capi.cu", "rs_device.cuh", "rs_tasks.cuh", "rs_lanes.cuh")] + [
    os.path.join(_ROOT, "include", f) for f in ("rs_spec.h", "rsoccer_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # flush-to-zero and 2-ulp division / square root: ~10 % fewer issued instructions in the
    # step kernels; the parity tolerance (1e-4 abs, tests/parity.py) is five orders above it
    "-ftz=true", "-prec-div=false", "-prec-sqrt=false",
    "-shared", "-Xcompiler", "-fPIC",
]

RS_OK = 0
ARR_BODY, ARR_ANG, ARR_OU, ARR_PREV, ARR_STEPS, ARR_INFO, ARR_COUNT = 0, 1, 2, 3, 4, 5, 6
TASK_VSS_V0, TASK_SSL_STATIC_DEFENDERS_V0, TASK_SSL_CONTESTED_POSSESSION_V0 = 0, 1, 2
TASK_SSL_DRIBBLING_V0, TASK_SSL_PASS_ENDURANCE_V0 = 3, 4

# every symbol include/rsoccer_b200.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "rs_version", "rs_last_error", "rs_create", "rs_destroy", "rs_state_bytes", "rs_bind_state",
    "rs_layout", "rs_field_params", "rs_reset", "rs_step", "rs_get_state", "rs_set_raw",
    "rs_get_raw", "rs_get_t", "rs_set_t", "rs_sync_t", "rs_task_obs_dim", "rs_task_act_dim", "rs_task_reset", "rs_vss_env_step",
    "rs_ssl_env_step", "rs_vss_env_step_host", "rs_ssl_env_step_host", "rs_launch_count", "rs_kernel_flags",
    "rs_set_option", "rs_get_option", "rs_debug_empty_step",
    "rs_vss_env_step_host_begin", "rs_ssl_env_step_host_begin", "rs_host_step_wait",
)
OPT_STEP_OVERLAP, OPT_PDL, OPT_OVERLAP_ERRORS, OPT_HOST_COPY_ACTIONS = 1, 2, 3, 4


class RsError(RuntimeError):
    pass


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    m = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > m for s in SOURCES)


def build(force=False, verbose=False):
    """nvcc cross-compiles the library for sm_100a (no GPU needed)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RsError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsError(
            "librsoccer_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64, i64 = C.c_void_p, C.c_int, C.c_uint64, C.c_int64
    L.rs_version.restype = i32
    L.rs_last_error.restype = C.c_char_p
    L.rs_create.argtypes = [i32, i32, i32, i32, i32, i32, i32, u64, i64, C.POINTER(vp)]
    L.rs_destroy.argtypes = [vp]
    L.rs_state_bytes.restype = C.c_size_t
    L.rs_state_bytes.argtypes = [vp]
    L.rs_bind_state.argtypes = [vp, vp, vp]
    L.rs_layout.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.rs_field_params.argtypes = [vp, C.POINTER(C.c_double)]
    L.rs_reset.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rs_step.argtypes = [vp, vp, vp]
    L.rs_get_state.argtypes = [vp, vp, vp]
    L.rs_set_raw.argtypes = [vp, vp, vp]
    L.rs_get_raw.argtypes = [vp, vp, vp]
    L.rs_get_t.restype = u64
    L.rs_get_t.argtypes = [vp]
    L.rs_set_t.argtypes = [vp, u64]
    L.rs_sync_t.argtypes = [vp, vp]
    L.rs_task_obs_dim.argtypes = [vp, i32]
    L.rs_task_reset.argtypes = [vp, i32, vp, vp, vp]
    L.rs_vss_env_step.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.rs_ssl_env_step.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.rs_vss_env_step_host.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.rs_ssl_env_step_host.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
    L.rs_vss_env_step_host_begin.argtypes = L.rs_vss_env_step_host.argtypes
    L.rs_ssl_env_step_host_begin.argtypes = L.rs_ssl_env_step_host.argtypes
    L.rs_host_step_wait.argtypes = [vp]
    L.rs_task_act_dim.restype = i32
    L.rs_task_act_dim.argtypes = [i32]
    L.rs_launch_count.restype = u64
    L.rs_launch_count.argtypes = [vp]
    L.rs_kernel_flags.restype = i32
    L.rs_kernel_flags.argtypes = [vp]
    L.rs_debug_empty_step.argtypes = [vp, i32, vp]
    L.rs_set_option.argtypes = [vp, i32, i64]
    L.rs_get_option.argtypes = [vp, i32, C.POINTER(i64), vp]
    _lib = L
    return L


def check(rc, what=""):
    if rc != RS_OK:
        msg = lib().rs_last_error()
        raise RsError("%s failed (%d): %s" % (what or "rsoccer_b200", rc, msg.decode() if msg else "?"))
