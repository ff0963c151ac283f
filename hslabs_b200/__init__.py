"""hslabs_b200 -- B200-native gait evaluation (IK + inverse dynamics + contact solve) behind a C ABI.

The compute path is libhsl_b200.so (hand-written sm_100a CUDA, see csrc/); this package is the thin
ctypes host layer that mirrors the reference's evaluation interface.  There is no CPU fallback.
"""
from . import api
from .api import (HSL_FLAG_IGNORE_REACH, HSL_NPARAM, HSL_ST_BAD_PARAMS, HSL_ST_FEW_CONTACTS, HSL_ST_ILLCOND, HSL_ST_SOLVER,
                  HSL_ST_UNREACHABLE, HslError, Model, lib_path, load_preset, make_params, model_path,
                  measure_cot, measure_cot_sweep, pinned_empty)

__all__ = ["Model", "HslError", "make_params", "load_preset", "model_path", "measure_cot", "measure_cot_sweep", "lib_path", "pinned_empty",
           "HSL_NPARAM", "HSL_FLAG_IGNORE_REACH", "HSL_ST_BAD_PARAMS", "HSL_ST_UNREACHABLE", "HSL_ST_SOLVER",
           "HSL_ST_FEW_CONTACTS", "HSL_ST_ILLCOND", "api"]
