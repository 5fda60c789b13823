"""ctypes binding of libggdmc_b200.so (the C ABI declared in include/ggdmc_b200.h).

The library is built in-tree by ``ggdmc_b200/csrc/Makefile`` (nvcc, sm_100a).  If it is missing
or cannot be loaded this module raises -- there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libggdmc_b200.so")

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)

OK, ERR_ARG, ERR_CUDA, ERR_COMM, ERR_CHAINS = 0, 1, 2, 3, 4
SCHEDULE_REFERENCE, SCHEDULE_PARALLEL, SCHEDULE_SIMULTANEOUS = 0, 1, 2
MODEL_LBA, MODEL_DDM = 0, 1  # enum ggdmc_model_type
MODEL_TYPES = {"lba": MODEL_LBA, "fastdm": MODEL_DDM}  # model@type strings (@hdr/likelihood.h:279)


class GgdmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class ModelT(C.Structure):
    _fields_ = [("n_acc", C.c_int32), ("n_cell", C.c_int32), ("npar", C.c_int32), ("n_const", C.c_int32),
                ("param_src", c_i32p), ("const_val", c_dp), ("posdrift", c_u8p), ("type", C.c_int32)]


class TrialsT(C.Structure):
    _fields_ = [("n_subject", C.c_int32), ("subject_offset", c_i64p), ("rt", c_dp), ("cell", c_u16p)]


class PriorT(C.Structure):
    _fields_ = [("npar", C.c_int32), ("p0", c_dp), ("p1", c_dp), ("lower", c_dp), ("upper", c_dp), ("dist", c_i32p),
                ("log_p", c_u8p)]


class ConfigT(C.Structure):
    _fields_ = [("nmc", C.c_int32), ("nchain", C.c_int32), ("thin", C.c_int32), ("report_length", C.c_int32),
                ("pop_migration_prob", C.c_double), ("sub_migration_prob", C.c_double), ("gamma_precursor", C.c_double),
                ("rp", C.c_double), ("is_hblocked", C.c_int32), ("is_pblocked", C.c_int32), ("nparameter", C.c_int32),
                ("schedule", C.c_int32), ("n_replicate", C.c_int32), ("device", C.c_int32), ("seed", c_u64p),
                ("subject_begin", C.c_int32), ("n_subject_total", C.c_int32)]


class SamplesT(C.Structure):
    _fields_ = [("npar", C.c_int32), ("nchain", C.c_int32), ("nmc", C.c_int32), ("theta", c_dp), ("lp", c_dp), ("ll", c_dp)]


class StartT(C.Structure):
    _fields_ = [("theta", c_dp), ("lp", c_dp), ("ll", c_dp)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_int32, C.c_void_p)

EXPORTS = [
    "ggdmc_b200_run_subject", "ggdmc_b200_run_hyper", "ggdmc_b200_run", "ggdmc_b200_trial_logdens", "ggdmc_b200_trial_logdens_hot",
    "ggdmc_b200_sumloglike", "ggdmc_b200_sumloglike_init", "ggdmc_b200_sumlogprior", "ggdmc_b200_select_chains", "ggdmc_b200_engine_create",
    "ggdmc_b200_engine_iterate", "ggdmc_b200_engine_iterate_flushed", "ggdmc_b200_engine_time_likelihood", "ggdmc_b200_engine_state",
    "ggdmc_b200_engine_launch_count", "ggdmc_b200_engine_is_persistent", "ggdmc_b200_engine_destroy", "ggdmc_b200_engine_profile", "ggdmc_b200_engine_counters", "ggdmc_b200_comm_unique_id", "ggdmc_b200_comm_init",
    "ggdmc_b200_comm_finalize", "ggdmc_b200_abi_version", "ggdmc_b200_device_count", "ggdmc_b200_measure_fp64_tflops",
    "ggdmc_b200_philox",
]


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    src_dir = os.path.join(HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "ggdmc_b200.h"))
    def stale():
        return not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale():
        import fcntl
        with open(os.path.join(src_dir, ".build.lock"), "w") as lock:  # the ranks of one launch must not build side by side
            fcntl.flock(lock, fcntl.LOCK_EX)
            if force or stale():
                r = subprocess.run(["make", "-C", src_dir], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError("building libggdmc_b200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA extension; fail loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C ggdmc_b200/csrc` "
                              "(ggdmc_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ggdmc_b200_measure_fp64_tflops.restype = C.c_double
        L.ggdmc_b200_engine_launch_count.restype = C.c_int64
        L.ggdmc_b200_engine_launch_count.argtypes = [C.c_void_p]
        L.ggdmc_b200_engine_is_persistent.argtypes = [C.c_void_p]
        L.ggdmc_b200_engine_destroy.argtypes = [C.c_void_p]
        L.ggdmc_b200_engine_destroy.restype = None
        L.ggdmc_b200_comm_finalize.restype = None
        L.ggdmc_b200_philox.restype = None
        _lib = L
    return _lib


def check(rc: int, err) -> None:
    if rc != 0:
        raise GgdmcError(rc, err.value.decode("utf-8", "replace"))


def errbuf():
    return C.create_string_buffer(256)


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a, t=c_dp):
    return a.ctypes.data_as(t) if a is not None else None
