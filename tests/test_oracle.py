"""Pins the CPU oracle (oracle/qsim_oracle.cpp): known-answer vectors of the reference's tests,
fixtures produced by the unmodified reference CPUSimulator, and — when oracle/_ref is present —
the live reference itself, bit for bit."""
import ctypes
from ctypes import c_int64, c_uint, c_void_p

import numpy as np
import pytest

import helpers as H

P = c_void_p


def _check_expect(state, exp, tol):
    probs = np.abs(state) ** 2
    if "state" in exp:
        want = np.array([complex(r, i) for r, i in exp["state"]])
        assert np.max(np.abs(state - want)) <= tol
    for k, v in exp.get("probs", {}).items():
        assert abs(probs[int(k)] - v) <= tol
    for k, v in exp.get("abs", {}).items():
        assert abs(abs(state[int(k)]) - v) <= tol
    if "probs_sum" in exp:
        idx, v = exp["probs_sum"]
        assert abs(sum(probs[i] for i in idx) - v) <= tol
    for k, v in exp.get("probs_gt", {}).items():
        assert probs[int(k)] > v


@pytest.mark.parametrize("case", H.load_known_answers()["cases"], ids=lambda c: c["name"])
def test_known_answers(case):
    g = H.gates([tuple(x) for x in case["gates"]])
    _check_expect(H.oracle_run(case["n"], g), case["expect"], case["tol"])


def _golden():
    z = np.load(H.GOLDEN + "/ref_cpu_states.npz")
    names = sorted(k[:-3] for k in z.files if k.endswith("__n"))
    return z, names


def test_golden_fixtures_bit_exact():
    """The oracle reproduces the reference CPUSimulator's states exactly (same arithmetic forms)."""
    z, names = _golden()
    assert len(names) >= 140
    for name in names:
        n, g, want = int(z[name + "__n"]), z[name + "__gates"], z[name + "__state"]
        got = H.oracle_run(n, np.ascontiguousarray(g, H.GATE_DTYPE))
        assert np.array_equal(got, want), name


@pytest.mark.skipif(H.reference() is None, reason="oracle/_ref not built")
def test_live_reference_bit_exact():
    rng = np.random.default_rng(7)
    for seed in range(12):
        n, d = int(rng.integers(1, 13)), int(rng.integers(1, 150))
        g = H.ref_random_circuit(n, d, seed)
        assert np.array_equal(H.oracle_run(n, g), H.ref_cpu_run(n, g))
    # all gate kinds the reference CPU path implements (1q set + CNOT/CZ/SWAP)
    kinds = list(range(11)) + [11, 12, 15]
    for trial in range(12):
        n, d = int(rng.integers(2, 11)), 60
        g = H.random_gates(n, d, rng, kinds)
        assert np.array_equal(H.oracle_run(n, g), H.ref_cpu_run(n, g))


@pytest.mark.skipif(H.reference() is None, reason="oracle/_ref not built")
def test_reference_cpu_skips_cry_crz_toffoli():
    """Documents defect D1: the reference CPU path ignores these gates, the oracle applies them."""
    g = H.gates([("X", 0), ("CRY", 0, 1, float(np.pi))])
    assert abs(H.ref_cpu_run(2, g)[1]) == 1.0          # unchanged: still |01>
    assert abs(abs(H.oracle_run(2, g)[3]) - 1.0) < 1e-12


def test_unitarity_and_inverses():
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 9):
        g = H.random_gates(n, 80, rng)
        st = H.oracle_run(n, g, H.random_state(n, rng))
        assert abs(np.linalg.norm(st) - 1.0) < 1e-12


def test_sampling_semantics():
    """lower_bound on the sequential CDF: first index with cum >= r; r == 0 -> 0; past the end -> N."""
    p = np.array([0.0, 0.25, 0.25, 0.5])
    got = H.oracle_sample(p, [0.0, 0.1, 0.25, 0.2500001, 0.5, 0.99, 1.0, 1.5])
    assert got.tolist() == [0, 1, 1, 2, 2, 3, 3, 4]
    # the uniforms are libstdc++'s mt19937 + uniform_real_distribution<double>
    u = H.mt19937_uniforms(42, 4)
    assert np.all((u >= 0) & (u < 1))
    # independent restatement: libstdc++ generate_canonical<double, 53> = (w0 + w1 * 2^32) / 2^64 over
    # the raw MT19937 words (numpy's legacy RandomState(seed) is the same init_genrand seeding)
    w = np.random.RandomState(42).randint(0, 2 ** 32, size=8, dtype=np.uint64)
    want = [(int(w[2 * i]) + int(w[2 * i + 1]) * 2 ** 32) / 2 ** 64 for i in range(4)]
    assert np.array_equal(u, np.array(want))


def test_measure_semantics():
    o = H.oracle()
    st = H.oracle_run(2, H.gates([("H", 0), ("CNOT", 0, 1)]))
    a = st.copy()
    p0 = ctypes.c_double()
    r = o.orc_measure(a.ctypes.data_as(P), 2, 1, ctypes.c_double(0.3), ctypes.byref(p0))
    assert r == 0 and abs(p0.value - 0.5) < 1e-15 and abs(abs(a[0]) - 1) < 1e-12
    b = st.copy()
    r = o.orc_measure(b.ctypes.data_as(P), 2, 0, ctypes.c_double(0.7), None)
    assert r == 1 and abs(abs(b[3]) - 1) < 1e-12
    z = H.zero_state(1)
    assert o.orc_measure(z.ctypes.data_as(P), 1, 0, ctypes.c_double(0.0), None) == 0


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    o = H.oracle()
    def run(ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); out = (ctypes.c_uint32 * 4)()
        o.orc_philox4x32_10(c, k, out)
        return [x for x in out]
    assert run([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_density_matrix_channels():
    o = H.oracle()
    n = 3
    rng = np.random.default_rng(5)
    psi = H.oracle_run(n, H.random_gates(n, 30, rng))
    rho = np.zeros((1 << n, 1 << n), np.complex128)
    o.orc_dm_from_pure(rho.ctypes.data_as(P), n, psi.ctypes.data_as(P))
    assert abs(o.orc_dm_trace(rho.ctypes.data_as(P), n) - 1) < 1e-12
    assert abs(o.orc_dm_purity(rho.ctypes.data_as(P), n) - 1) < 1e-12
    # unitary evolution of rho matches the state-vector path
    g = H.random_gates(n, 25, rng)
    psi2 = H.oracle_run(n, g, psi)
    for i in range(len(g)):
        assert o.orc_dm_apply_gate(rho.ctypes.data_as(P), n, g[i:i + 1].ctypes.data_as(P)) == 0
    assert np.max(np.abs(rho - np.outer(psi2, psi2.conj()))) < 1e-12
    # every channel is trace preserving and (p > 0) reduces purity; p = 0 is the identity
    for ch in range(6):
        r = rho.copy()
        assert o.orc_dm_channel(r.ctypes.data_as(P), n, ch, 1, ctypes.c_double(0.0)) == 0
        assert np.max(np.abs(r - rho)) < 1e-14
        assert o.orc_dm_channel(r.ctypes.data_as(P), n, ch, 1, ctypes.c_double(0.2)) == 0
        assert abs(o.orc_dm_trace(r.ctypes.data_as(P), n) - 1) < 1e-12
        assert o.orc_dm_purity(r.ctypes.data_as(P), n) < 1 - 1e-3
        assert np.max(np.abs(r - r.conj().T)) < 1e-14
    # bit flip with p = 1 is exactly X (reference tests/test_noise.cu:157-179)
    r = np.zeros((2, 2), np.complex128); r[1, 1] = 1
    o.orc_dm_channel(r.ctypes.data_as(P), 1, 3, 0, ctypes.c_double(1.0))
    assert abs(r[0, 0] - 1) < 1e-15 and abs(r[1, 1]) < 1e-15
    # amplitude damping gamma on |1><1|: population gamma moves to |0>
    r = np.zeros((2, 2), np.complex128); r[1, 1] = 1
    o.orc_dm_channel(r.ctypes.data_as(P), 1, 1, 0, ctypes.c_double(0.3))
    assert abs(r[0, 0] - 0.3) < 1e-15 and abs(r[1, 1] - 0.7) < 1e-15


def test_trajectory_average_matches_exact_channel():
    """Per-trajectory unravelling (oracle) averages to the exact channel (oracle DM) within sampling error."""
    o = H.oracle()
    n = 3
    g = H.gates([("H", 0), ("CNOT", 0, 1), ("CNOT", 1, 2)])
    ev = np.zeros(2 * n, H.CHANNEL_DTYPE)
    for q in range(n):
        ev[q] = (0, q, 0.05)          # depolarizing on every qubit
        ev[n + q] = (1, q, 0.1)       # amplitude damping on every qubit
    rho = np.zeros((8, 8), np.complex128); rho[0, 0] = 1
    assert o.orc_dm_run_schedule(rho.ctypes.data_as(P), n, g.ctypes.data_as(P), c_int64(len(g)),
                                 ev.ctypes.data_as(P), c_int64(len(ev))) == 0
    exact = np.real(np.diag(rho))
    T = 4000
    acc = np.zeros(8)
    for t in range(T):
        st = H.zero_state(n)
        assert o.orc_traj_run(st.ctypes.data_as(P), n, g.ctypes.data_as(P), c_int64(len(g)), ev.ctypes.data_as(P),
                              c_int64(len(ev)), c_uint(42), ctypes.c_uint64(t)) == 0
        assert abs(np.linalg.norm(st) - 1) < 1e-12
        acc += np.abs(st) ** 2
    acc /= T
    sigma = np.sqrt(np.maximum(exact * (1 - exact), 1e-6) / T)
    assert np.all(np.abs(acc - exact) < 5 * sigma + 1e-3), (acc, exact)
