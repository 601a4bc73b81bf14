"""Circuit IR: Python mirror of the reference's fluent builder (include/Circuit.hpp:89-144).

Gates are stored as a numpy record array with the C ABI's qsim_gate_t layout, so a circuit can be
handed to the engine without conversion.  Validation is done by the library's own builder
(qsim_circuit_validate), so the error behaviour is the C++ API's: OutOfRange for a bad qubit index,
InvalidArgument for duplicate qubits, non-finite angles or a bad qubit count.
"""
from __future__ import annotations

import enum
from ctypes import byref, c_int64
from typing import List

import numpy as np

from . import _lib
from ._lib import GATE_DTYPE


class GateType(enum.IntEnum):
    """Same order as `enum class GateType` (reference include/Circuit.hpp:42-59)."""
    X = 0
    Y = 1
    Z = 2
    H = 3
    S = 4
    T = 5
    Sdag = 6
    Tdag = 7
    Rx = 8
    Ry = 9
    Rz = 10
    CNOT = 11
    CZ = 12
    CRY = 13
    CRZ = 14
    SWAP = 15
    Toffoli = 16


_PARAMETRIC = {GateType.Rx, GateType.Ry, GateType.Rz, GateType.CRY, GateType.CRZ}


def gate_record(gtype: int, q0: int, q1: int = -1, q2: int = -1, param: float = 0.0) -> np.ndarray:
    rec = np.zeros(1, GATE_DTYPE)
    rec[0] = (int(gtype), q0, q1, q2, param)
    return rec


class Circuit:
    def __init__(self, num_qubits: int):
        self._n = int(num_qubits)
        _lib.check(_lib.lib().qsim_circuit_validate(self._n, None, 0))
        self._chunks: List[np.ndarray] = []
        self._count = 0

    # -- builder ------------------------------------------------------------------------------
    def _add(self, gtype, q0, q1=-1, q2=-1, param=0.0) -> "Circuit":
        rec = gate_record(gtype, q0, q1, q2, float(param))
        _lib.check(_lib.lib().qsim_circuit_validate(self._n, _lib.gates_ptr(rec), 1))
        self._chunks.append(rec)
        self._count += 1
        return self

    def x(self, q): return self._add(GateType.X, q)
    def y(self, q): return self._add(GateType.Y, q)
    def z(self, q): return self._add(GateType.Z, q)
    def h(self, q): return self._add(GateType.H, q)
    def s(self, q): return self._add(GateType.S, q)
    def t(self, q): return self._add(GateType.T, q)
    def sdag(self, q): return self._add(GateType.Sdag, q)
    def tdag(self, q): return self._add(GateType.Tdag, q)
    def rx(self, q, theta): return self._add(GateType.Rx, q, param=theta)
    def ry(self, q, theta): return self._add(GateType.Ry, q, param=theta)
    def rz(self, q, theta): return self._add(GateType.Rz, q, param=theta)
    def cnot(self, control, target): return self._add(GateType.CNOT, control, target)
    cx = cnot
    def cz(self, control, target): return self._add(GateType.CZ, control, target)
    def cry(self, control, target, theta): return self._add(GateType.CRY, control, target, param=theta)
    def crz(self, control, target, theta): return self._add(GateType.CRZ, control, target, param=theta)
    def swap(self, q1, q2): return self._add(GateType.SWAP, q1, q2)
    def toffoli(self, c1, c2, target): return self._add(GateType.Toffoli, c1, c2, target)
    ccx = toffoli

    def extend(self, gates: np.ndarray) -> "Circuit":
        """Append pre-built gate records (validated in one call)."""
        gates = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
        _lib.check(_lib.lib().qsim_circuit_validate(self._n, _lib.gates_ptr(gates), len(gates)))
        self._chunks.append(gates)
        self._count += len(gates)
        return self

    # -- access -------------------------------------------------------------------------------
    @property
    def gates(self) -> np.ndarray:
        if len(self._chunks) != 1:
            merged = np.concatenate(self._chunks) if self._chunks else np.zeros(0, GATE_DTYPE)
            self._chunks = [np.ascontiguousarray(merged, dtype=GATE_DTYPE)]
        return self._chunks[0]

    def get_num_qubits(self) -> int: return self._n
    def get_gate_count(self) -> int: return self._count
    def clear(self) -> None:
        self._chunks, self._count = [], 0

    def get_depth(self) -> int:
        d = c_int64(0)
        g = self.gates
        _lib.check(_lib.lib().qsim_circuit_depth(self._n, _lib.gates_ptr(g) if len(g) else None, len(g), byref(d)))
        return d.value

    def to_string(self) -> str:
        lines = [f"Circuit({self._n} qubits, {self._count} gates):"]
        for i, g in enumerate(self.gates):
            t = GateType(int(g["type"]))
            qs = [int(g[k]) for k in ("q0", "q1", "q2") if int(g[k]) >= 0]
            args = ", ".join(str(q) for q in qs)
            if t in _PARAMETRIC:
                args += f", {float(g['param']):g}"
            lines.append(f"  {i}: {t.name}({args})")
        return "\n".join(lines) + "\n"

    __str__ = to_string
    getNumQubits, getGateCount, getDepth, toString = get_num_qubits, get_gate_count, get_depth, to_string


def create_bell_circuit() -> Circuit:
    return Circuit(2).h(0).cnot(0, 1)


def create_ghz_circuit(num_qubits: int) -> Circuit:
    out = np.zeros(max(int(num_qubits), 1), GATE_DTYPE)
    _lib.check(_lib.lib().qsim_circuit_ghz(int(num_qubits), _lib.gates_ptr(out)))
    return Circuit(num_qubits).extend(out)


def create_random_circuit(num_qubits: int, depth: int, seed: int = 42) -> Circuit:
    """createRandomCircuit(n, depth, seed) of the reference (src/Circuit.cpp:252-282): `depth` gates
    drawn from {H, X, CNOT, Rz} with libstdc++'s mt19937 and distributions."""
    c = Circuit(num_qubits)
    out = np.zeros(int(depth), GATE_DTYPE)
    if depth > 0:
        _lib.check(_lib.lib().qsim_circuit_random(int(num_qubits), int(depth), int(seed) & 0xFFFFFFFF, _lib.gates_ptr(out)))
        c.extend(out)
    return c
