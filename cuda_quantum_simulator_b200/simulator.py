"""Simulator: Python mirror of qsim::Simulator (reference include/Simulator.hpp:53-85) over the C ABI."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_int64, c_void_p
from typing import Optional

import numpy as np

from . import _lib
from .circuit import Circuit, gate_record


class CompiledCircuit:
    """A circuit lowered to fused passes and resident on the device (qsim_program_t)."""

    def __init__(self, circuit: Circuit, n_global: int = 0, specialise: Optional[bool] = None):
        """specialise=True: every pass gets a run-time specialised kernel whatever the state size (worth it for a circuit
        that runs many times); None: the library's policy (large states only)."""
        self._h = c_void_p()
        g = circuit.gates
        self.num_qubits = circuit.get_num_qubits()
        self.n_gates = len(g)
        _lib.check(_lib.lib().qsim_program_compile(self.num_qubits, int(n_global), _lib.gates_ptr(g) if len(g) else None,
                                                   len(g), byref(self._h)))
        info = (c_int64 * 8)()
        _lib.check(_lib.lib().qsim_program_info(self._h, info))
        self.n_passes, self.n_ops, _, self.n_sweeps = info[0], info[1], info[2], info[3]
        if specialise:
            _lib.check(_lib.lib().qsim_program_set_specialised(self._h, 1))

    def describe(self) -> str:
        need = _lib.lib().qsim_program_describe(self._h, None, 0)
        buf = ctypes.create_string_buffer(need)
        _lib.lib().qsim_program_describe(self._h, buf, need)
        return buf.value.decode()

    def jit_source(self, pass_index: int = 0, whole_unit: bool = False) -> str:
        """CUDA C++ generated for one pass (the run-time specialised kernel); needs no GPU."""
        need = _lib.lib().qsim_program_jit_source(self._h, int(pass_index), int(whole_unit), None, 0)
        buf = ctypes.create_string_buffer(need)
        _lib.lib().qsim_program_jit_source(self._h, int(pass_index), int(whole_unit), buf, need)
        return buf.value.decode()

    def jit_request(self, pass_index: int = 0) -> str:
        """Queues the background compile of one pass: "ready" | "compiling" | "unavailable" (no GPU needed)."""
        st = ctypes.c_int()
        _lib.check(_lib.lib().qsim_program_jit_request(self._h, int(pass_index), byref(st)))
        return ("ready", "compiling", "unavailable")[st.value]

    def jit_compile(self, pass_index: int = 0, want_cubin: bool = False):
        """NVRTC-compiles one pass's kernel to an sm_100a cubin without loading it (no GPU needed).  Returns the cubin
        size, or the cubin bytes with want_cubin."""
        nb = c_int64()
        _lib.check(_lib.lib().qsim_program_jit_compile(self._h, int(pass_index), byref(nb), None, 0))
        if not want_cubin:
            return nb.value
        buf = ctypes.create_string_buffer(nb.value)
        _lib.check(_lib.lib().qsim_program_jit_compile(self._h, int(pass_index), byref(nb), buf, nb.value))
        return buf.raw

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qsim_program_destroy(self._h)
            self._h = c_void_p()


def jit_set_dual(mode, min_fp64: int = -1):
    """Two-warp-group build of the specialised kernels: mode "off" | "auto" | "always" (or 0 / 1 / 2); see qsim_jit_set_dual in
    include/qsim_b200.h."""
    m = {"off": 0, "auto": 1, "always": 2}.get(mode, mode)
    _lib.check(_lib.lib().qsim_jit_set_dual(int(m), int(min_fp64)))


def jit_set_mode(mode, min_qubits: int = 0):
    """mode: "off" | "auto" | "always" (or 0 / 1 / 2); see qsim_jit_set_mode in include/qsim_b200.h."""
    m = {"off": 0, "auto": 1, "always": 2}.get(mode, mode)
    _lib.check(_lib.lib().qsim_jit_set_mode(int(m), int(min_qubits)))


def jit_wait():
    """Blocks until every specialised kernel queued for background compilation is ready (see qsim_jit_wait)."""
    _lib.check(_lib.lib().qsim_jit_wait())


def jit_stats() -> dict:
    out = (c_int64 * 8)()
    _lib.check(_lib.lib().qsim_jit_stats(out))
    return {"compiles": out[0], "cache_hits": out[1], "launches": out[2], "failures": out[3], "compile_seconds": out[4] / 1e6,
            "disk_hits": out[5], "mode": ("off", "auto", "always")[out[6]], "min_qubits": out[7]}


class Simulator:
    def __init__(self, num_qubits: int, device_ptr: Optional[int] = None, _handle: Optional[c_void_p] = None):
        self._h = c_void_p()
        if _handle is not None:
            self._h = _handle
        elif device_ptr is None:
            _lib.check(_lib.lib().qsim_sim_create(int(num_qubits), byref(self._h)))
        else:
            _lib.check(_lib.lib().qsim_sim_create_external(int(num_qubits), c_void_p(device_ptr), byref(self._h)))
        self._n = int(num_qubits)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qsim_sim_destroy(self._h)
            self._h = c_void_p()

    close = __del__

    # -- execution ----------------------------------------------------------------------------
    def reset(self): _lib.check(_lib.lib().qsim_sim_reset(self._h))

    def run(self, circuit: Circuit):
        g = circuit.gates
        _lib.check(_lib.lib().qsim_sim_run(self._h, circuit.get_num_qubits(), _lib.gates_ptr(g) if len(g) else None, len(g)))

    def apply_gate(self, gtype, q0, q1=-1, q2=-1, param=0.0):
        rec = gate_record(gtype, q0, q1, q2, param)
        _lib.check(_lib.lib().qsim_sim_apply_gate(self._h, _lib.gates_ptr(rec)))

    def execute(self, program: CompiledCircuit): _lib.check(_lib.lib().qsim_sim_execute(self._h, program._h))
    def synchronize(self): _lib.check(_lib.lib().qsim_sim_synchronize(self._h))
    def set_stream(self, cuda_stream: int): _lib.check(_lib.lib().qsim_sim_set_stream(self._h, c_void_p(cuda_stream)))
    def init_basis(self, index: int): _lib.check(_lib.lib().qsim_sim_init_basis(self._h, int(index)))

    def set_state(self, amplitudes: np.ndarray):
        a = np.ascontiguousarray(amplitudes, dtype=np.complex128)
        assert a.size == self.get_state_size()
        _lib.check(_lib.lib().qsim_sim_set_state(self._h, a.ctypes.data_as(c_void_p)))

    # -- inspection ---------------------------------------------------------------------------
    def get_state_vector(self) -> np.ndarray:
        out = np.empty(self.get_state_size(), np.complex128)
        _lib.check(_lib.lib().qsim_sim_get_state(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def get_probabilities(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.get_state_size() - first if count is None else count
        out = np.empty(count, np.float64)
        _lib.check(_lib.lib().qsim_sim_get_probability_range(self._h, int(first), int(count), out.ctypes.data_as(c_void_p)))
        return out

    def marginal(self, qubits) -> np.ndarray:
        """Marginal distribution of up to 12 qubits (qubits[i] -> bit i of the outcome), computed on the device."""
        qs = np.ascontiguousarray(qubits, dtype=np.int32)
        out = np.empty(1 << len(qs), np.float64)
        _lib.check(_lib.lib().qsim_sim_marginal(self._h, qs.ctypes.data_as(c_void_p), len(qs), out.ctypes.data_as(c_void_p)))
        return out

    def get_total_probability(self) -> float:
        v = c_double()
        _lib.check(_lib.lib().qsim_sim_total_probability(self._h, byref(v)))
        return v.value

    # -- measurement --------------------------------------------------------------------------
    def sample(self, n_shots: int, seed: Optional[int] = None, uniforms: Optional[np.ndarray] = None) -> np.ndarray:
        """Sample without collapse.  `uniforms` (one per shot, in [0,1)) or `seed` (mt19937 draws, as the
        reference's NoisySimulator) make the outcome reproducible; indices are int64."""
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64)
            out = np.empty(len(u), np.int64)
            _lib.check(_lib.lib().qsim_sim_sample_uniforms(self._h, u.ctypes.data_as(c_void_p), len(u), out.ctypes.data_as(c_void_p)))
            return out
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        out = np.empty(max(int(n_shots), 0), np.int64)
        _lib.check(_lib.lib().qsim_sim_sample_seeded(self._h, int(seed) & 0xFFFFFFFF, int(n_shots), out.ctypes.data_as(c_void_p)))
        return out

    def measure_qubit(self, qubit: int, uniform: Optional[float] = None) -> int:
        """Simulator::measureQubit: measures index bit n-1-qubit (reference src/StateVector.cu:87-89)."""
        r = float(np.random.random()) if uniform is None else float(uniform)
        res = c_int()
        _lib.check(_lib.lib().qsim_sim_measure(self._h, int(qubit), r, byref(res)))
        return res.value

    def measure_bit(self, bit: int, uniform: float):
        res, p0 = c_int(), c_double()
        _lib.check(_lib.lib().qsim_sim_measure_bit(self._h, int(bit), float(uniform), byref(res), byref(p0)))
        return res.value, p0.value

    # -- info ---------------------------------------------------------------------------------
    def get_num_qubits(self) -> int: return self._n
    def get_state_size(self) -> int: return 1 << self._n
    def device_ptr(self) -> int: return int(_lib.lib().qsim_sim_device_ptr(self._h) or 0)
    def launch_count(self) -> int: return int(_lib.lib().qsim_sim_launch_count(self._h))
    def set_timing(self, on: bool): _lib.check(_lib.lib().qsim_sim_set_timing(self._h, int(bool(on))))

    def pass_time_ms(self):
        t, n = c_double(), c_int64()
        _lib.check(_lib.lib().qsim_sim_pass_time_ms(self._h, byref(t), byref(n)))
        return t.value, n.value

    def pass_times_ms(self):
        out = np.zeros(4096)
        n = c_int64()
        _lib.check(_lib.lib().qsim_sim_pass_times(self._h, out.ctypes.data_as(c_void_p), len(out), byref(n)))
        return out[:min(n.value, len(out))]

    getStateVector, getProbabilities, measureQubit = get_state_vector, get_probabilities, measure_qubit
    getNumQubits, getStateSize, applyGate = get_num_qubits, get_state_size, apply_gate
