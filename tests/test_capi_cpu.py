"""Host-side checks that need no GPU: the C-ABI library loads, exports everything the header
declares, mirrors the reference's error contract, and refuses (loudly) to compute without a device."""
import os
import re

import numpy as np
import pytest

import cuda_quantum_simulator_b200 as q
from cuda_quantum_simulator_b200 import _lib
import helpers as H


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(H.ROOT, "include", "qsim_b200.h")).read()
    declared = set(re.findall(r"QSIM_API[^;(]*?\b(qsim_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/qsim_b200.h but not exported"
    assert declared == set(L._qsim_declared), declared ^ set(L._qsim_declared)
    assert b"sm_100a" in L.qsim_version()


def test_circuit_error_contract():
    """reference src/Circuit.cpp:16-56 and tests/test_boundary.cu:110-152."""
    for bad in (0, -1, 40):
        with pytest.raises(q.InvalidArgument):
            q.Circuit(bad)
    q.Circuit(31)  # deliberate deviation: MAX_QUBITS raised from 30 to 36
    c = q.Circuit(3)
    with pytest.raises(q.OutOfRange):
        c.h(3)
    with pytest.raises(q.OutOfRange):
        c.cnot(0, -1)
    with pytest.raises(q.InvalidArgument):
        c.cnot(1, 1)
    with pytest.raises(q.InvalidArgument):
        c.toffoli(0, 1, 1)
    with pytest.raises(q.InvalidArgument):
        c.rx(0, float("nan"))
    with pytest.raises(q.InvalidArgument):
        c.crz(0, 1, float("inf"))
    assert c.get_gate_count() == 0
    with pytest.raises(q.InvalidArgument):
        q.create_ghz_circuit(1)


def test_circuit_builder_and_depth():
    c = q.Circuit(4).h(0).cnot(0, 1).cx(1, 2).ccx(0, 1, 3).rz(2, 0.5)
    assert c.get_gate_count() == 5 and c.get_num_qubits() == 4
    assert c.get_depth() == 4
    assert "Toffoli(0, 1, 3)" in c.to_string() and "Rz(2, 0.5)" in c.to_string()
    assert q.Circuit(2).get_depth() == 0
    assert q.create_bell_circuit().get_gate_count() == 2
    assert q.create_ghz_circuit(5).get_gate_count() == 5


@pytest.mark.skipif(H.reference() is None, reason="oracle/_ref not built")
def test_random_circuit_generator_matches_reference():
    """createRandomCircuit must draw the same circuit as the reference (libstdc++ distributions)."""
    for n, d, seed in [(30, 20, 42), (5, 29, 7), (1, 10, 3), (11, 99, 9), (24, 300, 12345)]:
        mine = q.create_random_circuit(n, d, seed).gates
        ref = H.ref_random_circuit(n, d, seed)
        assert np.array_equal(mine, ref)
    for n, d, seed in [(7, 40, 1), (12, 64, 5)]:
        g = q.create_random_circuit(n, d, seed)
        ref_depth = H.reference().ref_circuit_depth(n, g.gates.ctypes.data_as(H.P), H.c_int64(d))
        assert g.get_depth() == ref_depth


def test_c2_circuit_is_the_surveyed_one():
    names = [H.NAMES[t] for t in q.create_random_circuit(30, 20, 42).gates["type"]]
    assert names == "X Rz CNOT X H X X H CNOT Rz H H H X CNOT H CNOT H X X".split()
    g36 = q.create_random_circuit(36, 20, 42).gates
    assert (g36["q0"][11], H.NAMES[g36["type"][11]]) == (35, "H")   # the global-qubit gate of config C4


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(q.QsimError, match="no CPU fallback"):
        q.Simulator(3)
    # compiling a circuit is host logic (passes, sweeps, generated kernel source) and needs no device; there is simply
    # nothing to execute it on: every object that owns amplitudes refuses to exist
    prog = q.CompiledCircuit(q.Circuit(3).h(0))
    assert prog.n_passes == 1 and "jit_compute_tile" in prog.jit_source(0)
    for make in (lambda: q.NoisySimulator(3), lambda: q.BatchedSimulator(3, 4), lambda: q.DensityMatrixSimulator(2)):
        with pytest.raises(q.QsimError):
            make()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(H.ROOT, "cuda_quantum_simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libqsim_oracle" not in text and "oracle/_" not in text and "/root/reference" not in text, f
