"""ctypes binding of libpdlp_b200.so (C ABI in include/pdlp_b200.h).

There is no CPU fallback: if the shared library has not been built the import of
anything that needs it raises, loudly, with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("PDLP_B200_LIB") or os.path.join(HERE, "libpdlp_b200.so")   # env: developer sweeps only
SOURCES = [os.path.join(HERE, "csrc", "pdlp_b200.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("stream_spmv.cuh", "run_spmv.cuh", "panel_spmv.cuh", "pdlp_ops.cuh", "pdlp_block.cuh")] + \
          [os.path.join(ROOT, "include", "pdlp_b200.h")]

NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-O3", "-std=c++17"]


class Residuals(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("pr", "dr", "gap", "prim_obj", "adj_dual", "kkt", "pr_unscaled", "dr_unscaled")]


class State(C.Structure):
    _fields_ = [("eta", C.c_double), ("omega", C.c_double), ("theta", C.c_double), ("eta_sum", C.c_double),
                ("last_weight", C.c_double), ("dx2", C.c_double), ("dy2", C.c_double), ("inter", C.c_double),
                ("kkt_first", C.c_double), ("q_norm", C.c_double), ("c_norm", C.c_double),
                ("k", C.c_int64), ("t", C.c_int32), ("n_restarts", C.c_int32), ("infeas", C.c_int32),
                ("infeas_k", C.c_int32), ("solved", C.c_int32), ("reserved", C.c_int32)]


class Check(C.Structure):
    _fields_ = [("kind", C.c_int32), ("use_average", C.c_int32), ("t", C.c_int32), ("solved", C.c_int32),
                ("k", C.c_int64), ("omega", C.c_double), ("kkt_current", C.c_double), ("kkt_average", C.c_double),
                ("kkt_previous", C.c_double), ("kkt_first_new", C.c_double), ("point", Residuals)]


class SolveOptions(C.Structure):
    _fields_ = [("max_kkt", C.c_int64), ("restart_period", C.c_int32), ("use_graph", C.c_int32),
                ("time_limit", C.c_double), ("time_used", C.c_double), ("max_trace", C.c_int32),
                ("reserved", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("k", C.c_int64), ("j", C.c_int64), ("n", C.c_int32), ("status", C.c_int32),
                ("prim_obj", C.c_double), ("seconds", C.c_double), ("device_ms", C.c_double),
                ("n_trace", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/pdlp_b200.h declares: name -> (restype, argtypes)
class PlanInfo(C.Structure):
    _fields_ = [("rows", C.c_int64), ("nnz", C.c_int64), ("lane_rows", C.c_int64), ("warp_rows", C.c_int64),
                ("cta_rows", C.c_int64), ("slots", C.c_int64), ("packed_bytes", C.c_int64), ("sigma", C.c_int32),
                ("grid", C.c_int32), ("warps", C.c_int32), ("pieces", C.c_int32), ("max_piece_bytes", C.c_int32),
                ("stages", C.c_int32), ("max_warp_bytes", C.c_int64), ("mean_warp_bytes", C.c_int64),
                ("covered_once", C.c_int32), ("ranges_ok", C.c_int32)]


class RunInfo(C.Structure):
    _fields_ = [("stream_entries", C.c_int64), ("rows_finished", C.c_int64), ("grid", C.c_int32), ("warps", C.c_int32),
                ("chunks", C.c_int32), ("long_rows", C.c_int32), ("panels", C.c_int32), ("panel_cap", C.c_int32),
                ("chunk_bytes", C.c_int32), ("max_warp_chunks", C.c_int32)]


_P, _I, _L, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = {
    "pdlp_last_error": (C.c_char_p, []),
    "pdlp_bytes_per_iteration": (_L, [_P]),
    "pdlp_create": (_I, [C.POINTER(_P), _I, _L, _L, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pdlp_create_host": (_I, [C.POINTER(_P), _I, _L, _L, _L, _L, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "pdlp_destroy": (_I, [_P]),
    "pdlp_set_stream": (_I, [_P, _P]),
    "pdlp_set_iterate": (_I, [_P, _P, _P]),
    "pdlp_get_iterate": (_I, [_P, _P, _P]),
    "pdlp_get_average": (_I, [_P, _P, _P]),
    "pdlp_set_step": (_I, [_P, _D, _D, _D, _I]),
    "pdlp_set_options": (_I, [_P, _D, _I, _I, _D]),
    "pdlp_get_state": (_I, [_P, C.POINTER(State)]),
    "pdlp_set_counters": (_I, [_P, _L, C.c_int32]),
    "pdlp_iterate": (_I, [_P, _I, _I, _I]),
    "pdlp_time_kernels": (_I, [_P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pdlp_time_kernels_in_loop": (_I, [_P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pdlp_restart_check": (_I, [_P, _I, C.POINTER(Check)]),
    "pdlp_kkt": (_I, [_P, _P, _P, _D, C.POINTER(Residuals)]),
    "pdlp_primal_objective": (_I, [_P, C.POINTER(_D)]),
    "pdlp_power_iteration": (_I, [_P, _P, _I, C.POINTER(_D)]),
    "pdlp_spmv": (_I, [_P, _I, _P, _P]),
    "pdlp_ruiz": (_I, [_P, _I, _D, C.POINTER(_I)]),
    "pdlp_pock_chambolle": (_I, [_P]),
    "pdlp_get_scaling": (_I, [_P, _P, _P]),
    "pdlp_set_scaling": (_I, [_P, _P, _P]),
    "pdlp_get_problem": (_I, [_P, _P, _P, _P, _P]),
    "pdlp_get_matrix_values": (_I, [_P, _P]),
    "pdlp_get_matrix_structure": (_I, [_P, _P, _P]),
    "pdlp_batch_spmv": (_I, [_P, _P, _P]),
    "pdlp_batch_step": (_I, [_P, _P, _P, _P, _D, _D]),
    "pdlp_batch_gap": (_I, [_P, _P, _P, C.POINTER(_D)]),
    "pdlp_plan_inspect": (_I, [_P, C.c_int64, C.c_int32, C.c_int32, C.POINTER(PlanInfo), _P]),
    "pdlp_run_emulate": (_I, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P,
                              C.POINTER(RunInfo)]),
    "pdlp_shard_enable": (_I, [_P, _D, _D]),
    "pdlp_shard_kty": (_I, [_P, _P]),
    "pdlp_shard_step": (_I, [_P, _P, _I]),
    "pdlp_shard_rule_pack": (_I, [_P, _P]),
    "pdlp_shard_rule_apply": (_I, [_P, _P]),
    "pdlp_shard_check_begin": (_I, [_P, _P, _I]),
    "pdlp_shard_check_mid": (_I, [_P, _P, _P]),
    "pdlp_shard_check_decide": (_I, [_P, _P, _I, _P]),
    "pdlp_shard_check_end": (_I, [_P, _P, C.POINTER(Check)]),
    "pdlp_shard_exchange_create": (_I, [_P, _I, _I, _P]),
    "pdlp_shard_exchange_connect": (_I, [_P, _P]),
    "pdlp_shard_exchange_disconnect": (_I, [_P]),
    "pdlp_shard_iterate_p2p": (_I, [_P, _I, _I]),
    "pdlp_shard_allreduce_p2p": (_I, [_P, _P, C.c_int64]),
    "pdlp_solve": (_I, [_P, C.POINTER(SolveOptions), C.POINTER(SolveResult), C.POINTER(Check)]),
}

_lib = None


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library with argtypes set.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: torchpdlp_b200 has no CPU fallback. Build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)      # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class PdlpError(RuntimeError):
    code = 0


ERR_NOT_CANONICAL = 4


def check(rc: int) -> None:
    if rc != 0:
        err = PdlpError(f"pdlp_b200 error {rc}: {lib().pdlp_last_error().decode()}")
        err.code = rc
        raise err
