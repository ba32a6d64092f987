"""ctypes binding of include/tqsim.h.  Loads the in-tree libtqsim.so and fails loudly when it is missing."""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# (TQ_LIB_PATH: A/B of alternative in-tree builds during kernel work; the default is the library build() makes)
LIB_PATH = os.environ.get("TQ_LIB_PATH") or os.path.join(_HERE, "libtqsim.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["tq_api.cu", "tq_kernels.cu", "tq_stream.cu", "tq_plan.cpp", "tq_cobyla.cpp"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]

_lib = None

c_int_p = ctypes.POINTER(ctypes.c_int32)
c_dbl_p = ctypes.POINTER(ctypes.c_double)
c_u64_p = ctypes.POINTER(ctypes.c_uint64)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_i64_p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes): every symbol include/tqsim.h declares
SIGNATURES = {
    "tq_version": (ctypes.c_int, []),
    "tq_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "tq_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "tq_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "tq_set_pauli_hamiltonian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_u64_p, c_u64_p, c_dbl_p, c_dbl_p]),
    "tq_set_dense_hamiltonian": (ctypes.c_int, [ctypes.c_void_p, c_dbl_p]),
    "tq_set_init_state": (ctypes.c_int, [ctypes.c_void_p, c_dbl_p]),
    "tq_set_circuit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p,
                                      ctypes.c_int]),
    "tq_energy_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "tq_energy_batch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_dbl_p, ctypes.c_int, c_dbl_p]),
    "tq_energy_multi_host": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(c_dbl_p),
                                            ctypes.POINTER(c_u8_p), c_dbl_p]),
    "tq_energy_traj_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "tq_energy_traj_batch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_dbl_p, ctypes.c_int, c_u8_p,
                                                 ctypes.c_int, c_dbl_p]),
    "tq_energy_dm_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "tq_energy_dm_batch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_dbl_p, ctypes.c_int, c_dbl_p]),
    "tq_state_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "tq_evolve_states": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "tq_evolve_states_exchange": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_int, ctypes.c_int, c_u64_p, ctypes.c_void_p]),
    "tq_device_alloc": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]),
    "tq_device_free": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "tq_ipc_export": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, c_u8_p]),
    "tq_ipc_open": (ctypes.c_int, [ctypes.c_int, c_u8_p, ctypes.POINTER(ctypes.c_void_p)]),
    "tq_ipc_close": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "tq_state_batch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_dbl_p, ctypes.c_int, c_dbl_p]),
    "tq_dm_batch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_dbl_p, ctypes.c_int, c_dbl_p]),
    "tq_cobyla_create": (ctypes.c_int, [ctypes.c_int, c_dbl_p, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_void_p)]),
    "tq_cobyla_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "tq_cobyla_ask": (ctypes.c_int, [ctypes.c_void_p, c_dbl_p]),
    "tq_cobyla_tell": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    "tq_cobyla_result": (ctypes.c_int, [ctypes.c_void_p, c_dbl_p, c_dbl_p, c_int_p, c_int_p]),
    "tq_plan_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64_p]),
    "tq_plan_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64_p]),
    "tq_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "tq_plan_cache_stats": (ctypes.c_int, [ctypes.c_void_p, c_i64_p]),
    "tq_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "tq_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int_p, ctypes.POINTER(ctypes.c_float), c_dbl_p,
                                       c_dbl_p, c_int_p]),
    "tq_profile_read_flops": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_dbl_p, c_int_p]),
    "tq_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_dbl_p]),
    "tq_plan_dump": (ctypes.c_void_p, [ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_u64_p]),
    "tq_free": (None, [ctypes.c_void_p]),
}


def build(verbose=False):
    """Compile libtqsim.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    global _lib
    _lib = None
    return LIB_PATH


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "tqsim.h")]
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def lib():
    """The loaded library.  Raises (no fallback) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(tensorrl_qas_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
