"""NoiseModel, NoisySimulator, BatchedSimulator, DensityMatrixSimulator: Python mirrors of the qsim classes
(reference include/NoiseModel.cuh:46-297, include/DensityMatrix.cuh:63-224) over the C ABI."""
from __future__ import annotations

import ctypes
import enum
from ctypes import byref, c_double, c_int, c_int32, c_void_p
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .circuit import Circuit, gate_record


class NoiseType(enum.IntEnum):
    """Same order as `enum class NoiseType` (reference include/NoiseModel.cuh:46-53)."""
    Depolarizing = 0
    AmplitudeDamping = 1
    PhaseDamping = 2
    BitFlip = 3
    PhaseFlip = 4
    BitPhaseFlip = 5


class _CChannel(ctypes.Structure):      # qsim_noise_channel_t
    _fields_ = [("type", c_int32), ("n_qubits", c_int32), ("qubits", ctypes.POINTER(c_int32)), ("probability", c_double)]


class NoiseChannel:
    def __init__(self, ntype: NoiseType, qubits: Sequence[int], probability: float):
        self.type, self.qubits, self.probability = NoiseType(ntype), list(qubits), float(probability)


class NoiseModel:
    """Channel list; a channel added without qubits applies to every qubit."""

    def __init__(self):
        self._channels: List[NoiseChannel] = []

    def _add(self, t, qubits, p):
        if qubits is None:
            self._channels.append(NoiseChannel(t, [], p))
        else:
            for q in qubits:                       # one single-qubit channel per listed qubit, as the reference
                self._channels.append(NoiseChannel(t, [int(q)], p))
        return self

    def add_depolarizing(self, probability, qubits=None): return self._add(NoiseType.Depolarizing, qubits, probability)
    def add_amplitude_damping(self, gamma, qubits=None): return self._add(NoiseType.AmplitudeDamping, qubits, gamma)
    def add_phase_damping(self, gamma, qubits=None): return self._add(NoiseType.PhaseDamping, qubits, gamma)
    def add_bit_flip(self, probability, qubits=None): return self._add(NoiseType.BitFlip, qubits, probability)
    def add_phase_flip(self, probability, qubits=None): return self._add(NoiseType.PhaseFlip, qubits, probability)
    def add_bit_phase_flip(self, probability, qubits=None): return self._add(NoiseType.BitPhaseFlip, qubits, probability)
    def add_depolarizing_all(self, n, p): return self.add_depolarizing(p, range(n))
    def add_amplitude_damping_all(self, n, g): return self.add_amplitude_damping(g, range(n))
    def add_phase_damping_all(self, n, g): return self.add_phase_damping(g, range(n))

    def get_channels(self): return list(self._channels)
    def has_noise(self): return bool(self._channels)
    def clear(self): self._channels = []
    def channel_applies_to_qubit(self, ch: NoiseChannel, q: int): return not ch.qubits or q in ch.qubits

    def _c_array(self):
        """(array, n, keepalive) for the C ABI."""
        n = len(self._channels)
        arr = (_CChannel * max(n, 1))()
        keep = []
        for i, ch in enumerate(self._channels):
            qs = (c_int32 * max(len(ch.qubits), 1))(*ch.qubits)
            keep.append(qs)
            arr[i] = _CChannel(int(ch.type), len(ch.qubits), ctypes.cast(qs, ctypes.POINTER(c_int32)), ch.probability)
        return arr, n, keep


def _gates(circuit: Circuit):
    g = circuit.gates
    return (_lib.gates_ptr(g) if len(g) else None), len(g)


class NoisySimulator:
    def __init__(self, num_qubits: int, noise_model: Optional[NoiseModel] = None):
        self._n = int(num_qubits)
        self._h = c_void_p()
        arr, n, keep = (noise_model or NoiseModel())._c_array()
        _lib.check(_lib.lib().qsim_noisy_create(self._n, arr, n, byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qsim_noisy_destroy(self._h)
            self._h = c_void_p()

    def set_noise_model(self, m: NoiseModel):
        arr, n, keep = m._c_array()
        _lib.check(_lib.lib().qsim_noisy_set_noise(self._h, arr, n))

    def set_seed(self, seed: int): _lib.check(_lib.lib().qsim_noisy_set_seed(self._h, int(seed) & 0xFFFFFFFF))
    def reset(self): _lib.check(_lib.lib().qsim_noisy_reset(self._h))

    def run(self, circuit: Circuit):
        p, k = _gates(circuit)
        _lib.check(_lib.lib().qsim_noisy_run(self._h, circuit.get_num_qubits(), p, k))

    def apply_gate(self, gtype, q0, q1=-1, q2=-1, param=0.0):
        rec = gate_record(gtype, q0, q1, q2, param)
        _lib.check(_lib.lib().qsim_noisy_apply_gate(self._h, _lib.gates_ptr(rec)))

    def apply_noise_to_qubit(self, ntype, qubit, probability):
        m = NoiseModel()._add(NoiseType(ntype), [qubit], probability)
        arr, n, keep = m._c_array()
        _lib.check(_lib.lib().qsim_noisy_apply_noise(self._h, arr))

    def get_state_vector(self):
        out = np.empty(1 << self._n, np.complex128)
        _lib.check(_lib.lib().qsim_noisy_get_state(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def get_probabilities(self):
        out = np.empty(1 << self._n, np.float64)
        _lib.check(_lib.lib().qsim_noisy_get_probabilities(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def sample(self, n_shots: int):
        out = np.empty(max(int(n_shots), 0), np.int32)
        _lib.check(_lib.lib().qsim_noisy_sample(self._h, int(n_shots), out.ctypes.data_as(c_void_p)))
        return out

    def measure_qubit(self, qubit: int) -> int:
        r = c_int()
        _lib.check(_lib.lib().qsim_noisy_measure(self._h, int(qubit), byref(r)))
        return r.value

    def get_num_qubits(self): return self._n
    def get_state_size(self): return 1 << self._n


class BatchedSimulator:
    def __init__(self, num_qubits: int, batch_size: int, noise_model: Optional[NoiseModel] = None):
        self._n, self._b = int(num_qubits), int(batch_size)
        self._h = c_void_p()
        arr, n, keep = (noise_model or NoiseModel())._c_array()
        _lib.check(_lib.lib().qsim_batched_create(self._n, self._b, arr, n, byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qsim_batched_destroy(self._h)
            self._h = c_void_p()

    def set_noise_model(self, m: NoiseModel):
        arr, n, keep = m._c_array()
        _lib.check(_lib.lib().qsim_batched_set_noise(self._h, arr, n))

    def set_seed(self, seed: int): _lib.check(_lib.lib().qsim_batched_set_seed(self._h, int(seed) & 0xFFFFFFFF))
    def reset(self): _lib.check(_lib.lib().qsim_batched_reset(self._h))

    def run(self, circuit: Circuit):
        p, k = _gates(circuit)
        _lib.check(_lib.lib().qsim_batched_run(self._h, circuit.get_num_qubits(), p, k))

    def get_average_probabilities(self):
        out = np.empty(1 << self._n, np.float64)
        _lib.check(_lib.lib().qsim_batched_average_probabilities(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def get_probabilities(self, trajectory: int):
        out = np.empty(1 << self._n, np.float64)
        _lib.check(_lib.lib().qsim_batched_get_probabilities(self._h, int(trajectory), out.ctypes.data_as(c_void_p)))
        return out

    def get_trajectory_state(self, trajectory: int):
        out = np.empty(1 << self._n, np.complex128)
        _lib.check(_lib.lib().qsim_batched_get_state(self._h, int(trajectory), out.ctypes.data_as(c_void_p)))
        return out

    def sample(self, n_shots: int):
        """result[shot][trajectory], as the reference."""
        out = np.empty((max(int(n_shots), 0), self._b), np.int32)
        if n_shots > 0:
            _lib.check(_lib.lib().qsim_batched_sample(self._h, int(n_shots), out.ctypes.data_as(c_void_p)))
        return out

    def get_histogram(self, n_shots: int):
        out = np.zeros(1 << self._n, np.int32)
        _lib.check(_lib.lib().qsim_batched_histogram(self._h, int(n_shots), out.ctypes.data_as(c_void_p)))
        return out

    def get_num_qubits(self): return self._n
    def get_batch_size(self): return self._b
    def get_total_memory_bytes(self): return int(_lib.lib().qsim_batched_total_memory_bytes(self._h))


class DensityMatrixSimulator:
    def __init__(self, num_qubits: int, noise_model: Optional[NoiseModel] = None):
        self._n = int(num_qubits)
        self._h = c_void_p()
        arr, n, keep = (noise_model or NoiseModel())._c_array()
        _lib.check(_lib.lib().qsim_dm_create(self._n, arr, n, byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qsim_dm_destroy(self._h)
            self._h = c_void_p()

    def reset(self): _lib.check(_lib.lib().qsim_dm_reset(self._h))

    def run(self, circuit: Circuit):
        p, k = _gates(circuit)
        _lib.check(_lib.lib().qsim_dm_run(self._h, circuit.get_num_qubits(), p, k))

    def apply_gate(self, gtype, q0, q1=-1, q2=-1, param=0.0):
        rec = gate_record(gtype, q0, q1, q2, param)
        _lib.check(_lib.lib().qsim_dm_apply_gate(self._h, _lib.gates_ptr(rec)))

    def apply_channel(self, ntype, qubit, probability):
        _lib.check(_lib.lib().qsim_dm_apply_channel(self._h, int(ntype), int(qubit), float(probability)))

    def init_from_pure_state(self, state):
        a = np.ascontiguousarray(state, np.complex128)
        assert a.size == 1 << self._n
        _lib.check(_lib.lib().qsim_dm_init_pure(self._h, a.ctypes.data_as(c_void_p)))

    def init_maximally_mixed(self): _lib.check(_lib.lib().qsim_dm_init_maximally_mixed(self._h))

    def get_probabilities(self):
        out = np.empty(1 << self._n, np.float64)
        _lib.check(_lib.lib().qsim_dm_get_probabilities(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def get_density_matrix(self):
        d = 1 << self._n
        out = np.empty((d, d), np.complex128)
        _lib.check(_lib.lib().qsim_dm_get_matrix(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def get_purity(self):
        v = c_double()
        _lib.check(_lib.lib().qsim_dm_purity(self._h, byref(v)))
        return v.value

    def get_trace(self):
        v = c_double()
        _lib.check(_lib.lib().qsim_dm_trace(self._h, byref(v)))
        return v.value

    def is_valid(self, tolerance=1e-10):
        v = c_int()
        _lib.check(_lib.lib().qsim_dm_is_valid(self._h, float(tolerance), byref(v)))
        return bool(v.value)

    def measure_qubit(self, qubit: int, uniform: Optional[float] = None):
        u = float(np.random.random()) if uniform is None else float(uniform)
        r = c_int()
        _lib.check(_lib.lib().qsim_dm_measure(self._h, int(qubit), u, byref(r)))
        return r.value

    def get_num_qubits(self): return self._n
