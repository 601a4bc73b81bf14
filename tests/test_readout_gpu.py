"""Read-out parity: probabilities, bit-identical sampling and measurement (north_star: sampled indices
bit-identical given the same uniform draws; measurement outcomes bit-identical under the same seed)."""
import ctypes

import numpy as np
import pytest

import cuda_quantum_simulator_b200 as q
import helpers as H

pytestmark = pytest.mark.gpu


def prepared(n, g, state=None):
    sim = q.Simulator(n)
    if state is not None:
        sim.set_state(state)
    if len(g):
        sim.run(q.Circuit(n).extend(g))
    return sim


def test_probabilities_bit_exact_vs_std_norm():
    rng = np.random.default_rng(1)
    for n in (1, 3, 10, 16):
        st = H.random_state(n, rng)
        sim = prepared(n, [], st)
        assert np.array_equal(sim.get_probabilities(), H.oracle_probs(st))
        assert np.array_equal(sim.get_probabilities(3 % (1 << n), 1), H.oracle_probs(st)[3 % (1 << n):][:1])


def test_basis_state_sampling_is_deterministic():
    """reference tests/test_statevector.cu:101-123."""
    sim = q.Simulator(2)
    assert np.all(sim.sample(100) == 0)
    sim.init_basis(3)
    assert np.all(sim.sample(100) == 3)
    with pytest.raises(q.InvalidArgument):
        sim.sample(0)


@pytest.mark.parametrize("n", [1, 2, 5, 11, 12, 13, 16, 20])
def test_sampling_bit_identical_to_sequential_cdf(n):
    rng = np.random.default_rng(40 + n)
    st = H.random_state(n, rng)
    sim = prepared(n, [], st)
    probs = H.oracle_probs(st)
    u = np.concatenate([rng.random(4096), [0.0, 0.5, 1.0 - 2 ** -53, 0.999999999]])
    assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(probs, u))
    # draws placed exactly on / next to CDF steps (where a parallel scan would disagree)
    cum = np.cumsum(probs)
    picks = cum[rng.integers(0, len(cum), 512)]
    edge = np.concatenate([picks, np.nextafter(picks, 0), np.nextafter(picks, 2)])
    edge = edge[edge < 1.0]
    assert np.array_equal(sim.sample(0, uniforms=edge), H.oracle_sample(probs, edge))
    # total probability equals the sequential host sum bit for bit
    assert sim.get_total_probability() == H.oracle().orc_total_probability(probs.ctypes.data_as(H.P), H.c_int64(len(probs)))


def test_sampling_flat_and_sparse_states():
    n = 18
    sim = prepared(n, H.gates([("H", i) for i in range(n)]))       # flat, dyadic probabilities
    probs = H.oracle_probs(H.oracle_run(n, H.gates([("H", i) for i in range(n)])))
    u = np.random.default_rng(0).random(2048)
    assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(probs, u))
    g = H.gates([("H", 0)] + [("CNOT", i, i + 1) for i in range(n - 1)])   # GHZ: two spikes, c == 0.5 plateau
    sim = prepared(n, g)
    probs = H.oracle_probs(H.oracle_run(n, g))
    assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(probs, u))
    out = sim.sample(1000, seed=42)
    assert np.array_equal(out, H.oracle_sample(probs, H.mt19937_uniforms(42, 1000)))
    assert set(out.tolist()) <= {0, (1 << n) - 1}


@pytest.mark.parametrize("n", [14, 20])      # 14: the one-CTA small-state path (scan + exact replay inside the margin)
@pytest.mark.parametrize("shape", ["plateau_half", "noise_tail", "ties", "tiny_head", "over_one"])
def test_sampling_hard_cdf_shapes(shape, n):
    """States that stress the exact parallel CDF (csrc/kernels_readout.cu): the running sum parked on a power of two
    over many chunks, a long tail of ~1e-34 probabilities after the sum has reached ~1, terms that round on exact ties,
    a head of tiny probabilities (many binades per chunk), and a sum that passes 1.0.  All must equal the host's
    sequential algorithm bit for bit — and stay fast (the replay path is sequential)."""
    N = 1 << n
    rng = np.random.default_rng(77)
    amp = np.zeros(N, np.complex128)
    if shape == "plateau_half":
        amp[: N // 4] = np.sqrt(0.5 / (N // 4))                  # c reaches 0.5 (up to rounding) at N/4
        amp[N // 4: 3 * N // 4] = 1e-40 * rng.random(N // 2)     # ... and sits there for half the state
        amp[3 * N // 4:] = np.sqrt(0.5 / (N // 4))
    elif shape == "noise_tail":
        amp[: N // 2] = rng.normal(size=N // 2) + 1j * rng.normal(size=N // 2)
        amp[: N // 2] /= np.linalg.norm(amp[: N // 2])
        amp[N // 2:] = 1e-17 * (rng.normal(size=N // 2) + 1j * rng.normal(size=N // 2))
    elif shape == "ties":
        # p = 3 * 2^-54 style terms: half an ulp of the running sum once it is in [0.5, 1)
        amp[:] = np.sqrt(1.0 / N)
        amp[N // 2 + 5:: 97] = np.sqrt(3.0 * 2.0 ** -54)
    elif shape == "tiny_head":
        amp[: N // 2] = 10.0 ** rng.uniform(-160, -20, N // 2)
        amp[N // 2:] = np.sqrt(1.0 / (N // 2))
    else:
        amp[:] = np.sqrt(1.0000001 / N)                            # unnormalised on purpose: the sum crosses 1.0
    sim = prepared(n, [], amp)
    probs = H.oracle_probs(amp)
    cum = np.cumsum(probs)
    picks = cum[rng.integers(0, N, 256)]
    u = np.concatenate([rng.random(1024), picks, np.nextafter(picks, 0), np.nextafter(picks, 2), [0.0, 0.5, 0.25]])
    u = u[u < 1.0]
    assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(probs, u))
    assert sim.get_total_probability() == H.oracle().orc_total_probability(probs.ctypes.data_as(H.P), H.c_int64(len(probs)))


def test_measure_matches_reference_semantics():
    """StateVector::measure: bit n-1-q, r < p0 ? 0 : 1, collapse by 1/sqrt(p) (src/StateVector.cu:260-314)."""
    rng = np.random.default_rng(9)
    n = 9
    for trial in range(6):
        st = H.random_state(n, rng)
        qubit, r = int(rng.integers(0, n)), float(rng.random())
        sim = prepared(n, [], st)
        got = sim.measure_qubit(qubit, r)
        ref = st.copy()
        p0 = ctypes.c_double()
        want = H.oracle().orc_measure(ref.ctypes.data_as(H.P), n, n - 1 - qubit, ctypes.c_double(r), ctypes.byref(p0))
        assert got == want
        assert np.max(np.abs(sim.get_state_vector() - ref)) < 1e-14
        # the direct-bit variant reports the exact sequential p0
        sim2 = prepared(n, [], st)
        res, p0_gpu = sim2.measure_bit(n - 1 - qubit, r)
        assert res == want and p0_gpu == p0.value


def test_measure_edge_cases():
    sim = q.Simulator(2)
    assert sim.measure_qubit(0, 0.999) == 0          # |00>: always 0 (tests/test_statevector.cu:125-135)
    sim = q.Simulator(1)
    sim.apply_gate(q.GateType.X, 0)
    assert sim.measure_qubit(0, 0.0) == 1
    with pytest.raises(q.InvalidArgument):
        sim.measure_qubit(1, 0.5)
    with pytest.raises(q.InvalidArgument):
        sim.measure_qubit(-1, 0.5)
    # Bell correlation (tests/test_statevector.cu:137-172)
    sim = q.Simulator(2)
    sim.run(q.create_bell_circuit())
    a = sim.measure_qubit(0, 0.7)
    b = sim.measure_qubit(1, 0.2)
    assert a == b


@pytest.mark.parametrize("n", [1, 3, 9, 14, 20])
def test_marginal_probabilities(n):
    """Device-side marginals (qsim_sim_marginal): what a caller of the reference computes from getProbabilities()."""
    rng = np.random.default_rng(600 + n)
    st = H.random_state(n, rng)
    sim = prepared(n, [], st)
    probs = H.oracle_probs(st)
    idx = np.arange(1 << n)
    for k in sorted({0, 1, min(n, 2), min(n, 5), min(n, 12)}):
        qs = [int(x) for x in rng.permutation(n)[:k]]
        outcome = np.zeros(1 << n, np.int64)
        for i, qb in enumerate(qs):
            outcome |= ((idx >> qb) & 1) << i
        want = np.bincount(outcome, weights=probs, minlength=1 << k)
        got = sim.marginal(qs)
        assert got.shape == (1 << k,)
        assert np.max(np.abs(got - want)) < 1e-13, (n, k, qs)
        assert np.array_equal(got, sim.marginal(qs))          # fixed summation order: reproducible bit for bit
    with pytest.raises(q.InvalidArgument):
        sim.marginal([0, 0] if n > 1 else [5])
