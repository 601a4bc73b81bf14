"""Noisy trajectories and density matrices on the GPU against the oracle: exact channels (DM restatement),
per-trajectory unravelling with the same Philox draws, and the reference's own noise test expectations
(tests/test_noise.cu, tests/test_density_matrix.cu)."""
import ctypes
from ctypes import c_int64, c_uint

import numpy as np
import pytest

import cuda_quantum_simulator_b200 as q
import helpers as H

pytestmark = pytest.mark.gpu
P = H.P


def events_of(model: q.NoiseModel, n):
    ev = []
    for ch in model.get_channels():
        for qb in (ch.qubits or range(n)):
            ev.append((int(ch.type), qb, ch.probability))
    return np.array(ev, H.CHANNEL_DTYPE) if ev else np.zeros(0, H.CHANNEL_DTYPE)


def oracle_traj(n, g, ev, seed, traj, state=None):
    st = H.zero_state(n) if state is None else np.array(state, np.complex128)
    rc = H.oracle().orc_traj_run(st.ctypes.data_as(P), n, g.ctypes.data_as(P), c_int64(len(g)),
                                 ev.ctypes.data_as(P) if len(ev) else None, c_int64(len(ev)), c_uint(seed),
                                 ctypes.c_uint64(traj))
    assert rc == 0
    return st


# ---- batched trajectories ---------------------------------------------------------------------------

def test_batched_no_noise_is_ideal_for_every_gate_type():
    rng = np.random.default_rng(3)
    n = 7
    g = H.random_gates(n, 90, rng)
    sim = q.BatchedSimulator(n, 5)
    sim.run(q.Circuit(n).extend(g))
    want = H.oracle_run(n, g)
    for t in range(5):
        assert np.max(np.abs(sim.get_trajectory_state(t) - want)) < 1e-12
    assert np.max(np.abs(sim.get_average_probabilities() - np.abs(want) ** 2)) < 1e-12
    assert sim.get_total_memory_bytes() == 5 * (1 << n) * 16


def test_trajectories_match_oracle_unravelling_draw_for_draw():
    """Same Philox stream, same schedule: every trajectory's amplitudes agree with the CPU restatement."""
    n = 5
    rng = np.random.default_rng(8)
    g = H.random_gates(n, 25, rng)
    m = q.NoiseModel().add_depolarizing(0.15).add_amplitude_damping(0.2, [0, 3]).add_phase_damping(0.1, [1])
    m.add_bit_flip(0.1, [2]).add_phase_flip(0.2, [4]).add_bit_phase_flip(0.1, [0])
    ev = events_of(m, n)
    sim = q.BatchedSimulator(n, 64, m)
    sim.set_seed(1234)
    sim.run(q.Circuit(n).extend(g))
    worst = 0.0
    for t in range(64):
        want = oracle_traj(n, g, ev, 1234, t)
        worst = max(worst, float(np.max(np.abs(sim.get_trajectory_state(t) - want))))
    assert worst < 1e-12
    # a second run continues the noise stream (fresh draws) on top of the evolved states
    sim2 = q.BatchedSimulator(n, 4, m)
    sim2.set_seed(1234)
    sim2.run(q.Circuit(n).extend(g))
    a = sim2.get_trajectory_state(1)
    assert np.max(np.abs(a - sim.get_trajectory_state(1))) == 0.0     # same seed => identical (tests/test_noise.cu:345-377)


def test_config_c5_average_matches_exact_channel():
    """BASELINE config 5 at reduced width: GHZ + depolarizing 0.005 + amplitude damping 0.001 on every qubit after
    every gate; the trajectory average converges to the exact Kraus evolution within sampling error."""
    n, batch = 6, 20000
    g = q.create_ghz_circuit(n).gates
    m = q.NoiseModel().add_depolarizing(0.02).add_amplitude_damping(0.01)      # global channels = all qubits (D7)
    ev = events_of(m, n)
    assert len(ev) == 2 * n
    sim = q.BatchedSimulator(n, batch, m)
    sim.set_seed(42)
    sim.run(q.Circuit(n).extend(g))
    avg = sim.get_average_probabilities()
    rho = np.zeros((1 << n, 1 << n), np.complex128)
    rho[0, 0] = 1
    assert H.oracle().orc_dm_run_schedule(rho.ctypes.data_as(P), n, g.ctypes.data_as(P), c_int64(len(g)),
                                          ev.ctypes.data_as(P), c_int64(len(ev))) == 0
    exact = np.real(np.diag(rho))
    sigma = np.sqrt(np.maximum(exact * (1 - exact), 1e-7) / batch)
    assert np.all(np.abs(avg - exact) < 5 * sigma + 2e-4), np.max(np.abs(avg - exact) / sigma)
    assert abs(avg.sum() - 1) < 1e-10
    hist = sim.get_histogram(1)
    assert hist.sum() == batch                                                    # tests/test_noise.cu:313-330
    assert np.all(np.abs(hist / batch - exact) < 6 * sigma + 1e-3)


def test_config_c5_full_size_against_exact_channel():
    """BASELINE config 5 AT ITS STATED SIZE: BatchedSimulator(12, 65536), createGHZCircuit(12), depolarizing 0.005 +
    amplitude damping 0.001 on every qubit after every gate, seed 42.  The exact distribution is the
    committed fixture tests/golden/c5_exact_diag.npy (oracle Kraus evolution of the 4096 x 4096 density matrix,
    generator tests/golden/make_c5_golden.py); tolerance 4 sigma of the 65 536-trajectory sampling error."""
    n, batch = 12, 65536
    exact = np.load(H.GOLDEN + "/c5_exact_diag.npy")
    assert exact.shape == (1 << n,) and abs(exact.sum() - 1) < 1e-12
    m = q.NoiseModel().add_depolarizing(0.005).add_amplitude_damping(0.001)
    assert len(events_of(m, n)) == 2 * n
    sim = q.BatchedSimulator(n, batch, m)
    sim.set_seed(42)
    sim.run(q.create_ghz_circuit(n))
    avg = sim.get_average_probabilities()
    assert abs(avg.sum() - 1) < 1e-10
    # the average of |a|^2 over trajectories has per-entry variance <= p(1-p)/batch (it is a mean of values in [0,1] with mean
    # p): 4 sigma.  The 4096 cells are tested at once and the rare ones are Poisson, not Gaussian - a cell with p ~ 1e-6
    # is hit by 0, 1, 2, ... whole trajectories of weight up to 1 - so the bound also allows 8 full-weight hits (8 / batch).
    sigma = np.sqrt(np.maximum(exact * (1 - exact), 0) / batch)
    assert np.all(np.abs(avg - exact) < 4 * sigma + 8.0 / batch), float(np.max(np.abs(avg - exact) / (sigma + 8.0 / batch)))
    # the rare cells jointly: their total weight is one number with a Gaussian error
    rare = exact < 1e-4
    s_rare = np.sqrt(exact[rare].sum() / batch)
    assert abs(avg[rare].sum() - exact[rare].sum()) < 4 * s_rare + 1e-6, (avg[rare].sum(), exact[rare].sum())
    assert abs(avg[0] - exact[0]) < 4 * sigma[0] and abs(avg[-1] - exact[-1]) < 4 * sigma[-1]
    hist = sim.get_histogram(1)
    assert hist.sum() == batch                                                    # tests/test_noise.cu:313-330
    # one shot per trajectory is a multinomial draw of the exact distribution; the 4096 entries are tested jointly:
    # chi-square over the cells with an expected count >= 5 (the rest pooled)
    expct = exact * batch
    big = expct >= 5
    chi2 = float(np.sum((hist[big] - expct[big]) ** 2 / expct[big]))
    rest_e, rest_o = expct[~big].sum(), hist[~big].sum()
    if rest_e > 0:
        chi2 += (rest_o - rest_e) ** 2 / rest_e
    dof = int(big.sum())
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
    assert abs(hist[0] / batch - exact[0]) < 4 * np.sqrt(exact[0] * (1 - exact[0]) / batch)
    assert abs(hist[-1] / batch - exact[-1]) < 4 * np.sqrt(exact[-1] * (1 - exact[-1]) / batch)


def test_batched_sampling_matches_sequential_cdf():
    n, batch, shots = 6, 16, 5
    rng = np.random.default_rng(2)
    m = q.NoiseModel().add_depolarizing(0.1)
    sim = q.BatchedSimulator(n, batch, m)
    sim.set_seed(7)
    sim.run(q.Circuit(n).extend(H.random_gates(n, 30, rng)))
    out = sim.sample(shots)
    u = H.mt19937_uniforms(7, batch * shots)          # trajectory-major draws from the member engine
    for t in range(batch):
        probs = H.oracle_probs(sim.get_trajectory_state(t))
        want = H.oracle_sample(probs, u[t * shots:(t + 1) * shots])
        assert np.array_equal(out[:, t], want)


def test_reference_noise_expectations():
    """reference tests/test_noise.cu: p=0 exact (:106-122), bit-flip p=1 after X -> |0> (:157-179), phase flip keeps
    probabilities (:185-200), batched Bell on every trajectory (:283-311)."""
    sim = q.NoisySimulator(2, q.NoiseModel().add_depolarizing(0.0, [0, 1]))
    sim.run(q.create_bell_circuit())
    assert np.max(np.abs(sim.get_probabilities() - [0.5, 0, 0, 0.5])) < 1e-12
    sim = q.NoisySimulator(1, q.NoiseModel().add_bit_flip(1.0, [0]))
    sim.run(q.Circuit(1).x(0))
    assert abs(sim.get_probabilities()[0] - 1.0) < 1e-12
    sim = q.NoisySimulator(1, q.NoiseModel().add_phase_flip(1.0, [0]))
    sim.run(q.Circuit(1).h(0))
    assert np.max(np.abs(sim.get_probabilities() - 0.5)) < 1e-12
    b = q.BatchedSimulator(2, 10)
    b.run(q.create_bell_circuit())
    for t in range(10):
        assert np.max(np.abs(b.get_probabilities(t) - [0.5, 0, 0, 0.5])) < 1e-12
    with pytest.raises(q.OutOfRange):
        b.get_probabilities(10)
    with pytest.raises(q.InvalidArgument):
        b.run(q.Circuit(3).h(0))
    # amplitude damping on one qubit: ground-state population grows (tests/test_noise.cu:206-231)
    ground = 0
    for seed in range(100):
        s = q.NoisySimulator(1, q.NoiseModel().add_amplitude_damping(0.5, [0]))
        s.set_seed(seed)
        s.run(q.Circuit(1).x(0))
        ground += int(s.sample(1)[0] == 0)
    assert 25 < ground < 75


def test_noisy_simulator_wide_state_and_readout():
    """n > 13: single trajectory through the fused-pass engine with host-side draws from the same stream."""
    n = 15
    g = H.gates([("H", 0)] + [("CNOT", i, i + 1) for i in range(6)] + [("Ry", 9, 0.7)])
    m = q.NoiseModel().add_bit_flip(1.0, [2]).add_amplitude_damping(0.3, [9])
    sim = q.NoisySimulator(n, m)
    sim.set_seed(5)
    sim.run(q.Circuit(n).extend(g))
    want = oracle_traj(n, g, events_of(m, n), 5, 0)
    assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-12
    # seeded sampling / measurement: reference draw order (one mt19937 double per shot, then one per measurement)
    u = H.mt19937_uniforms(5, 11)
    probs = H.oracle_probs(want)
    assert np.array_equal(sim.sample(10), H.oracle_sample(probs, u[:10]))
    ref = want.copy()
    want_bit = H.oracle().orc_measure_noisy(ref.ctypes.data_as(P), n, 3, ctypes.c_double(u[10]))
    assert sim.measure_qubit(3) == want_bit
    assert np.max(np.abs(sim.get_state_vector() - ref)) < 1e-12


# ---- density matrix -------------------------------------------------------------------------------------

def test_density_matrix_unit_answers():
    """reference tests/test_density_matrix.cu:16-184."""
    dm = q.DensityMatrixSimulator(2)
    assert abs(dm.get_trace() - 1) < 1e-12 and abs(dm.get_purity() - 1) < 1e-12
    dm.run(q.create_bell_circuit())
    assert np.max(np.abs(dm.get_probabilities() - [0.5, 0, 0, 0.5])) < 1e-12
    assert abs(dm.get_purity() - 1) < 1e-12 and dm.is_valid()
    dm.init_maximally_mixed()
    assert abs(dm.get_purity() - 0.25) < 1e-12 and abs(dm.get_trace() - 1) < 1e-12
    for bad in (0, 15):
        with pytest.raises(q.InvalidArgument):
            q.DensityMatrixSimulator(bad)


def test_density_matrix_gates_match_oracle_all_types():
    rng = np.random.default_rng(4)
    n = 5
    psi = H.random_state(n, rng)
    g = H.random_gates(n, 60, rng)
    dm = q.DensityMatrixSimulator(n)
    dm.init_from_pure_state(psi)
    dm.run(q.Circuit(n).extend(g))
    want = H.oracle_run(n, g, psi)
    assert np.max(np.abs(dm.get_density_matrix() - np.outer(want, want.conj()))) < 1e-12
    assert abs(dm.get_purity() - 1) < 1e-11


def test_density_matrix_channels_match_exact_kraus():
    rng = np.random.default_rng(6)
    n = 4
    psi = H.random_state(n, rng)
    rho = np.ascontiguousarray(np.outer(psi, psi.conj()))
    dm = q.DensityMatrixSimulator(n)
    dm.init_from_pure_state(psi)
    for step in range(18):
        t, qb, p = step % 6, int(rng.integers(0, n)), float(rng.uniform(0.05, 0.4))
        dm.apply_channel(t, qb, p)
        assert H.oracle().orc_dm_channel(rho.ctypes.data_as(P), n, t, qb, ctypes.c_double(p)) == 0
        if step % 5 == 0:
            g = H.random_gates(n, 4, rng)
            dm.run(q.Circuit(n).extend(g))
            for i in range(len(g)):
                H.oracle().orc_dm_apply_gate(rho.ctypes.data_as(P), n, g[i:i + 1].ctypes.data_as(P))
    got = dm.get_density_matrix()
    assert np.max(np.abs(got - rho)) < 1e-12
    assert abs(dm.get_trace() - 1) < 1e-12 and dm.get_purity() < 1 and dm.is_valid()


def test_density_matrix_simulator_noise_schedule_and_measure():
    """Channels after each gate on the qubits it touched (reference src/DensityMatrix.cu:201-212)."""
    n = 3
    m = q.NoiseModel().add_depolarizing(0.1).add_amplitude_damping(0.05, [1])
    g = H.gates([("H", 0), ("CNOT", 0, 1), ("Ry", 2, 0.4), ("CZ", 1, 2)])
    dm = q.DensityMatrixSimulator(n, m)
    dm.run(q.Circuit(n).extend(g))
    rho = np.zeros((8, 8), np.complex128)
    rho[0, 0] = 1
    for i in range(len(g)):
        H.oracle().orc_dm_apply_gate(rho.ctypes.data_as(P), n, g[i:i + 1].ctypes.data_as(P))
        for qb in [int(g[i][k]) for k in ("q0", "q1", "q2") if int(g[i][k]) >= 0]:
            H.oracle().orc_dm_channel(rho.ctypes.data_as(P), n, 0, qb, ctypes.c_double(0.1))
            if qb == 1:
                H.oracle().orc_dm_channel(rho.ctypes.data_as(P), n, 1, qb, ctypes.c_double(0.05))
    assert np.max(np.abs(dm.get_density_matrix() - rho)) < 1e-12
    want = H.oracle().orc_dm_measure(rho.ctypes.data_as(P), n, 1, ctypes.c_double(0.3))
    assert dm.measure_qubit(1, 0.3) == want
    assert np.max(np.abs(dm.get_density_matrix() - rho)) < 1e-12
