"""ctypes loader for libqsim_b200.so (the C ABI declared in include/qsim_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call needs a GPU and none
is present, the error propagates.  Nothing in this package imports the CPU oracle under oracle/.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqsim_b200.so")

#: numpy mirror of qsim_gate_t (== GateOp of the reference, include/Circuit.hpp:64-84)
GATE_DTYPE = np.dtype([("type", "<i4"), ("q0", "<i4"), ("q1", "<i4"), ("q2", "<i4"), ("param", "<f8")], align=True)
assert GATE_DTYPE.itemsize == 24


class QsimError(RuntimeError):
    """std::runtime_error of the C++ API (CUDA failures, unknown gate, zero-probability outcome)."""


class InvalidArgument(ValueError):
    """std::invalid_argument of the C++ API."""


class OutOfRange(IndexError):
    """std::out_of_range of the C++ API."""


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QsimError(
                f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
        _declare(_lib)
        # specialised kernels are compiled on background threads: the interpreter must not start tearing the process down
        # (NVRTC's own exit handlers) while one of them is still inside the compiler
        import atexit
        atexit.register(_lib.qsim_jit_shutdown)
    return _lib


def check(status: int) -> None:
    if status == 0:
        return
    msg = lib().qsim_last_error().decode("utf-8", "replace")
    if status == 1:
        raise InvalidArgument(msg)
    if status == 2:
        raise OutOfRange(msg)
    raise QsimError(msg)


def _declare(L: ctypes.CDLL) -> None:
    P = c_void_p
    PP = POINTER(c_void_p)
    sig = {
        "qsim_last_error": (c_char_p, []),
        "qsim_version": (c_char_p, []),
        "qsim_max_qubits": (c_int, []),
        "qsim_circuit_validate": (c_int, [c_int, P, c_int64]),
        "qsim_circuit_random": (c_int, [c_int, c_int, c_uint, P]),
        "qsim_circuit_ghz": (c_int, [c_int, P]),
        "qsim_circuit_depth": (c_int, [c_int, P, c_int64, POINTER(c_int64)]),
        "qsim_program_compile": (c_int, [c_int, c_int, P, c_int64, PP]),
        "qsim_program_compile_ex": (c_int, [c_int, c_int, P, c_int64, c_uint64, PP]),
        "qsim_program_compile_ex2": (c_int, [c_int, c_int, P, c_int64, c_uint64, c_int, PP]),
        "qsim_program_destroy": (None, [P]),
        "qsim_program_info": (c_int, [P, POINTER(c_int64)]),
        "qsim_program_describe": (c_size_t, [P, c_char_p, c_size_t]),
        "qsim_jit_set_mode": (c_int, [c_int, c_int]),
        "qsim_jit_set_dual": (c_int, [c_int, c_int]),
        "qsim_jit_shutdown": (c_int, []),
        "qsim_jit_stats": (c_int, [POINTER(c_int64)]),
        "qsim_jit_wait": (c_int, []),
        "qsim_program_jit_request": (c_int, [P, c_int, POINTER(c_int)]),
        "qsim_program_set_specialised": (c_int, [P, c_int]),
        "qsim_program_jit_source": (c_size_t, [P, c_int, c_int, c_char_p, c_size_t]),
        "qsim_program_jit_compile": (c_int, [P, c_int, POINTER(c_int64), P, c_size_t]),
        "qsim_sim_create": (c_int, [c_int, PP]),
        "qsim_sim_create_external": (c_int, [c_int, P, PP]),
        "qsim_sim_destroy": (None, [P]),
        "qsim_sim_set_stream": (c_int, [P, P]),
        "qsim_sim_reset": (c_int, [P]),
        "qsim_sim_init_basis": (c_int, [P, c_uint64]),
        "qsim_sim_set_state": (c_int, [P, P]),
        "qsim_sim_run": (c_int, [P, c_int, P, c_int64]),
        "qsim_sim_apply_gate": (c_int, [P, P]),
        "qsim_sim_execute": (c_int, [P, P]),
        "qsim_sim_synchronize": (c_int, [P]),
        "qsim_sim_get_state": (c_int, [P, P]),
        "qsim_sim_get_probabilities": (c_int, [P, P]),
        "qsim_sim_get_probability_range": (c_int, [P, c_uint64, c_uint64, P]),
        "qsim_sim_total_probability": (c_int, [P, POINTER(c_double)]),
        "qsim_sim_marginal": (c_int, [P, P, c_int, P]),
        "qsim_sim_sample_uniforms": (c_int, [P, P, c_int64, P]),
        "qsim_sim_sample_seeded": (c_int, [P, c_uint, c_int64, P]),
        "qsim_sim_measure": (c_int, [P, c_int, c_double, POINTER(c_int)]),
        "qsim_sim_measure_bit": (c_int, [P, c_int, c_double, POINTER(c_int), POINTER(c_double)]),
        "qsim_sim_num_qubits": (c_int, [P]),
        "qsim_sim_device_ptr": (c_void_p, [P]),
        "qsim_sim_launch_count": (c_int64, [P]),
        "qsim_sim_set_timing": (c_int, [P, c_int]),
        "qsim_sim_pass_time_ms": (c_int, [P, POINTER(c_double), POINTER(c_int64)]),
        "qsim_sim_pass_timeline": (c_int, [P, P, c_int64, P]),
        "qsim_sim_pass_times": (c_int, [P, P, c_int64, POINTER(c_int64)]),
        "qsim_shard_create": (c_int, [c_int, c_int, c_int, P, PP]),
        "qsim_shard_swap_p2p": (c_int, [P, P, c_int, c_int]),
        "qsim_shard_cdf_prepare": (c_int, [P, P]),
        "qsim_shard_cdf_classify": (c_int, [P, c_double]),
        "qsim_shard_execute_exchange": (c_int, [P, P, P, P, c_int, c_int]),
        "qsim_program_last_tile_mask": (c_int, [P, P]),
        "qsim_shard_execute_exchange_inplace": (c_int, [P, P, P, c_int, c_int, P, P, c_uint64, c_uint64, P]),
        "qsim_shard_inplace_exchange_possible": (c_int, [P, P, c_int, P]),
        "qsim_shard_execute_exchange_half": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, c_uint64, c_uint64, P]),
        "qsim_shard_split_exchange_possible": (c_int, [P, P, P, c_int, P]),
        "qsim_shard_pack_half": (c_int, [P, c_int, c_int, c_int64, c_int64, P]),
        "qsim_shard_unpack_half": (c_int, [P, c_int, c_int, c_int64, c_int64, P]),
        "qsim_ipc_get_handle": (c_int, [P, P, POINTER(c_uint64)]),
        "qsim_ipc_open_handle": (c_int, [P, PP]),
        "qsim_ipc_close_handle": (c_int, [P]),
        "qsim_shard_partial_probability": (c_int, [P, c_int, POINTER(c_double)]),
        "qsim_shard_collapse": (c_int, [P, c_int, c_int, c_double]),
        "qsim_sharded_unique_id": (c_int, [P]),
        "qsim_sharded_create": (c_int, [c_int, c_int, c_int, P, c_int, PP]),
        "qsim_sharded_destroy": (None, [P]),
        "qsim_sharded_reset": (c_int, [P]),
        "qsim_sharded_run": (c_int, [P, c_int, P, c_int64]),
        "qsim_sharded_compile": (c_int, [P, c_int, P, c_int64, c_int, PP]),
        "qsim_sharded_plan_destroy": (None, [P]),
        "qsim_sharded_plan_info": (c_int, [P, POINTER(c_int64)]),
        "qsim_sharded_execute": (c_int, [P, P]),
        "qsim_sharded_sample": (c_int, [P, P, c_int64, P]),
        "qsim_sharded_measure": (c_int, [P, c_int, c_double, POINTER(c_int)]),
        "qsim_sharded_measure_bit": (c_int, [P, c_int, c_double, POINTER(c_int), POINTER(c_double)]),
        "qsim_sharded_marginal": (c_int, [P, P, c_int, P]),
        "qsim_sharded_total_probability": (c_int, [P, POINTER(c_double)]),
        "qsim_sharded_get_local_state": (c_int, [P, P]),
        "qsim_sharded_set_local_state": (c_int, [P, P]),
        "qsim_sharded_restore_identity_layout": (c_int, [P]),
        "qsim_sharded_layout": (c_int, [P, P, POINTER(c_uint64)]),
        "qsim_sharded_set_identity_layout_only": (c_int, [P, c_int]),
        "qsim_sharded_swap": (c_int, [P, c_int, c_int]),
        "qsim_sharded_relabel_identity": (c_int, [P]),
        "qsim_sharded_info": (c_int, [P, POINTER(c_int64)]),
        "qsim_sharded_inplace_exchanges": (c_int, [P, P]),
        "qsim_sharded_split_exchanges": (c_int, [P, P]),
        "qsim_sharded_local": (c_void_p, [P]),
        "qsim_sharded_set_stream": (c_int, [P, P]),
        "qsim_sharded_synchronize": (c_int, [P]),
        "qsim_sharded_barrier": (c_int, [P]),
        "qsim_sharded_plan_circuit": (c_int, [c_int, c_int, P, c_int64, P, c_int, P, c_int64, P, P, P, POINTER(c_int64)]),
        "qsim_noisy_create": (c_int, [c_int, P, c_int, PP]),
        "qsim_noisy_destroy": (None, [P]),
        "qsim_noisy_set_noise": (c_int, [P, P, c_int]),
        "qsim_noisy_set_seed": (c_int, [P, c_uint]),
        "qsim_noisy_reset": (c_int, [P]),
        "qsim_noisy_run": (c_int, [P, c_int, P, c_int64]),
        "qsim_noisy_apply_gate": (c_int, [P, P]),
        "qsim_noisy_apply_noise": (c_int, [P, P]),
        "qsim_noisy_get_state": (c_int, [P, P]),
        "qsim_noisy_get_probabilities": (c_int, [P, P]),
        "qsim_noisy_sample": (c_int, [P, c_int, P]),
        "qsim_noisy_measure": (c_int, [P, c_int, POINTER(c_int)]),
        "qsim_batched_create": (c_int, [c_int, c_int, P, c_int, PP]),
        "qsim_batched_destroy": (None, [P]),
        "qsim_batched_set_noise": (c_int, [P, P, c_int]),
        "qsim_batched_set_seed": (c_int, [P, c_uint]),
        "qsim_batched_reset": (c_int, [P]),
        "qsim_batched_run": (c_int, [P, c_int, P, c_int64]),
        "qsim_batched_average_probabilities": (c_int, [P, P]),
        "qsim_batched_get_probabilities": (c_int, [P, c_int, P]),
        "qsim_batched_get_state": (c_int, [P, c_int, P]),
        "qsim_batched_sample": (c_int, [P, c_int, P]),
        "qsim_batched_histogram": (c_int, [P, c_int, P]),
        "qsim_batched_total_memory_bytes": (c_size_t, [P]),
        "qsim_dm_create": (c_int, [c_int, P, c_int, PP]),
        "qsim_dm_destroy": (None, [P]),
        "qsim_dm_reset": (c_int, [P]),
        "qsim_dm_run": (c_int, [P, c_int, P, c_int64]),
        "qsim_dm_apply_gate": (c_int, [P, P]),
        "qsim_dm_apply_channel": (c_int, [P, c_int, c_int, c_double]),
        "qsim_dm_init_pure": (c_int, [P, P]),
        "qsim_dm_init_maximally_mixed": (c_int, [P]),
        "qsim_dm_get_probabilities": (c_int, [P, P]),
        "qsim_dm_get_matrix": (c_int, [P, P]),
        "qsim_dm_purity": (c_int, [P, POINTER(c_double)]),
        "qsim_dm_trace": (c_int, [P, POINTER(c_double)]),
        "qsim_dm_is_valid": (c_int, [P, c_double, POINTER(c_int)]),
        "qsim_dm_measure": (c_int, [P, c_int, c_double, POINTER(c_int)]),
        "qsim_shard_sample": (c_int, [P, c_double, c_int, P, c_int64, P, POINTER(c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError here == the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    L._qsim_declared = tuple(sig)


def gates_ptr(gates: np.ndarray) -> c_void_p:
    assert gates.dtype == GATE_DTYPE and gates.flags["C_CONTIGUOUS"]
    return gates.ctypes.data_as(c_void_p)
