"""Parity of the CUDA gate-application path (through the C ABI) against the CPU oracle, the reference's
known answers and the fixtures produced by the reference CPUSimulator.  Tolerance: 1e-10 max-abs per
amplitude (north_star; fusion reorders sums), most cases are checked at 1e-12."""
import numpy as np
import pytest

import cuda_quantum_simulator_b200 as q
import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-10


def run_gpu(n, g, state=None):
    sim = q.Simulator(n)
    if state is not None:
        sim.set_state(state)
    c = q.Circuit(n).extend(g) if len(g) else q.Circuit(n)
    sim.run(c)
    return sim.get_state_vector(), sim


def test_native_library_is_loaded():
    import os
    assert os.path.exists(q.LIB_PATH)
    sim = q.Simulator(3)
    sim.run(q.Circuit(3).h(0))
    assert sim.launch_count() > 0


@pytest.mark.parametrize("case", H.load_known_answers()["cases"], ids=lambda c: c["name"])
def test_known_answers(case):
    from test_oracle import _check_expect
    g = H.gates([tuple(x) for x in case["gates"]])
    st, _ = run_gpu(case["n"], g)
    _check_expect(st, case["expect"], case["tol"])


def test_reference_cpu_fixtures():
    z = np.load(H.GOLDEN + "/ref_cpu_states.npz")
    names = sorted(k[:-3] for k in z.files if k.endswith("__n"))
    worst = 0.0
    for name in names:
        n, g = int(z[name + "__n"]), np.ascontiguousarray(z[name + "__gates"], H.GATE_DTYPE)
        st, _ = run_gpu(n, g)
        err = np.max(np.abs(st - z[name + "__state"]))
        worst = max(worst, err)
        assert err < (1e-10 if "deep" in name else 1e-12), (name, err)
    print("worst fixture error", worst)


@pytest.mark.parametrize("n", list(range(1, 17)))
def test_fuzz_all_gate_types_from_random_state(n):
    rng = np.random.default_rng(100 + n)
    for trial in range(3):
        d = int(rng.integers(1, 160))
        g = H.random_gates(n, d, rng)
        st0 = H.random_state(n, rng)
        got, _ = run_gpu(n, g, st0)
        want = H.oracle_run(n, g, st0)
        assert np.max(np.abs(got - want)) < 1e-12, (n, d, trial)


@pytest.mark.parametrize("n", [2, 5, 12, 13, 14, 16, 19])
def test_trailing_flips_fold_into_the_store(n):
    """X / CNOT / Toffoli / SWAP runs at the end of a pass are folded into the final store's addressing."""
    rng = np.random.default_rng(700 + n)
    for trial in range(4):
        g = H.flip_heavy_gates(n, rng, body=int(rng.integers(0, 8)), tail=int(rng.integers(1, 40)))
        st0 = H.random_state(n, rng)
        got, _ = run_gpu(n, g, st0)
        want = H.oracle_run(n, g, st0)
        assert np.max(np.abs(got - want)) < 1e-12, (n, trial)
        gr = np.ascontiguousarray(g[::-1])            # leading flips: folded into the first sweep's load
        got, _ = run_gpu(n, gr, st0)
        want = H.oracle_run(n, gr, st0)
        assert np.max(np.abs(got - want)) < 1e-12, ("leading", n, trial)


@pytest.mark.parametrize("n", [1, 4, 12, 13, 15, 18, 21])
def test_run_from_a_recorded_basis_state(n):
    """reset() / init_basis() only record the basis state; the first pass of the next run generates its tiles on chip
    (fused_pass_kernel, PassParams::init_basis) — with deferred X gates on tile and non-tile qubits, folded flips,
    several passes, and a compiled program as well as run()."""
    rng = np.random.default_rng(4000 + n)
    for trial in range(4):
        idx = int(rng.integers(0, 1 << n))
        g = H.random_gates(n, int(rng.integers(1, 60)), rng)
        if n > 1:
            extra = H.flip_heavy_gates(n, rng, body=2, tail=int(rng.integers(1, 12)))
            g = np.concatenate([g, extra])
        st0 = np.zeros(1 << n, np.complex128)
        st0[idx] = 1.0
        want = H.oracle_run(n, g, st0)
        sim = q.Simulator(n)
        sim.init_basis(idx)
        c = q.Circuit(n).extend(g)
        if trial % 2:
            sim.execute(q.CompiledCircuit(c))
        else:
            sim.run(c)
        assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-12, (n, trial, idx)
        # and again after a plain reset, sampling without ever touching the amplitudes from the host
        sim.reset()
        sim.run(c)
        u = rng.random(64)
        want0 = H.oracle_run(n, g)
        assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(H.oracle_probs(want0), u)), (n, trial)


@pytest.mark.parametrize("n", [13, 14, 16, 18])
def test_leading_flips_with_a_plain_store(n):
    """Flips folded into a sweep's LOAD read other threads' slots; when the same sweep then stores in place (no folded
    trailing flips, no X frame) a barrier must separate the two.  Short passes make the window wide: a CNOT ladder that
    cannot slide to the end because a Hadamard on its last target follows."""
    rng = np.random.default_rng(8100 + n)
    for trial in range(12):
        qs = [int(x) for x in rng.permutation(n)[: int(rng.integers(3, min(n, 9)))]]
        lst = [("CNOT", qs[i], qs[i + 1]) for i in range(len(qs) - 1)]
        lst += [("H", qs[-1]), ("H", qs[0])]
        g = H.gates(lst)
        st0 = H.random_state(n, rng)
        want = H.oracle_run(n, g, st0)
        for rep in range(3):
            got, _ = run_gpu(n, g, st0)
            assert np.max(np.abs(got - want)) < 1e-12, (n, trial, rep)


def test_every_gate_on_every_qubit_18q():
    """Each gate type with its target on every bit position class (lane, register, warp, outside-tile)."""
    n = 18
    rng = np.random.default_rng(5)
    st0 = H.random_state(n, rng)
    for t in range(17):
        lst = []
        for q0 in range(n):
            q1, q2 = (q0 + 7) % n, (q0 + 11) % n
            g = [t, q0] + ([q1] if t >= 11 else []) + ([q2] if t == 16 else [])
            if t in (8, 9, 10, 13, 14):
                g.append(0.37 + 0.1 * q0)
            lst.append(tuple(g))
        g = H.gates(lst)
        got, _ = run_gpu(n, g, st0)
        want = H.oracle_run(n, g, st0)
        assert np.max(np.abs(got - want)) < 1e-12, H.NAMES[t]


def test_config_c1_20q_benchmark_circuit():
    """BASELINE config 1: benchmarks/benchmark_scaling.cu:68-75 (100 H + 20 CNOT at 20 qubits)."""
    g = H.bench_c1_gates(20)
    assert len(g) == 120
    got, sim = run_gpu(20, g)
    want = H.oracle_run(20, g)
    assert np.max(np.abs(got - want)) < TOL
    assert abs(sim.get_total_probability() - 1.0) < 1e-10


def test_run_composes_and_reset():
    """Simulator::run does not reset (reference src/Simulator.cu:28-36); reset() does."""
    sim = q.Simulator(5)
    c = q.Circuit(5).h(0).cnot(0, 3).rz(3, 0.4)
    sim.run(c)
    sim.run(c)
    g2 = np.concatenate([c.gates, c.gates])
    assert np.max(np.abs(sim.get_state_vector() - H.oracle_run(5, g2))) < 1e-12
    sim.reset()
    assert sim.get_state_vector()[0] == 1.0
    sim.apply_gate(q.GateType.X, 2)
    assert abs(sim.get_state_vector()[4] - 1.0) < 1e-15
    with pytest.raises(q.InvalidArgument):
        sim.run(q.Circuit(4).h(0))


def test_compiled_circuit_matches_run():
    rng = np.random.default_rng(2)
    n = 14
    g = H.random_gates(n, 200, rng)
    c = q.Circuit(n).extend(g)
    prog = q.CompiledCircuit(c)
    assert prog.n_passes < 200
    sim = q.Simulator(n)
    sim.execute(prog)
    assert np.max(np.abs(sim.get_state_vector() - H.oracle_run(n, g))) < 1e-11


def test_long_random_circuit_stays_normalised():
    """reference tests/test_boundary.cu:197-212: 1000-gate random circuit, |sum p - 1| <= 1e-10."""
    c = q.create_random_circuit(10, 1000, 7)
    sim = q.Simulator(10)
    sim.run(c)
    assert abs(sim.get_total_probability() - 1.0) <= 1e-10
    assert np.max(np.abs(sim.get_state_vector() - H.oracle_run(10, c.gates))) < 1e-10


@pytest.mark.parametrize("n", [24, 26])
def test_large_state_against_oracle(n):
    """Sizes where tiles have scattered high qubits and many CTAs per pass; oracle takes seconds."""
    c = q.create_random_circuit(n, 20, 42)
    sim = q.Simulator(n)
    sim.run(c)
    got = sim.get_state_vector()
    want = H.oracle_run(n, c.gates)
    assert np.max(np.abs(got - want)) < TOL


def test_30q_properties():
    """Full-size config C2 through size-independent properties: norm, known sparsity pattern, inverse."""
    n = 30
    c = q.create_random_circuit(n, 20, 42)
    sim = q.Simulator(n)
    sim.run(c)
    assert abs(sim.get_total_probability() - 1.0) < 1e-10
    # C2 leaves a uniform superposition over 2^k basis states: every sampled index has the same probability
    idx = np.unique(sim.sample(64, seed=1))
    pr = np.array([sim.get_probabilities(int(i), 1)[0] for i in idx])
    assert np.all(pr > 0) and np.allclose(pr, pr[0], rtol=1e-12)
    assert abs(np.log2(pr[0]) - round(np.log2(pr[0]))) < 1e-9
    # run the inverse circuit: back to a basis state (X/H/CNOT self-inverse, Rz(-theta))
    inv = q.Circuit(n)
    for g in c.gates[::-1]:
        t = int(g["type"])
        if t == 10:
            inv.rz(int(g["q0"]), -float(g["param"]))
        elif t == 11:
            inv.cnot(int(g["q0"]), int(g["q1"]))
        else:
            inv._add(t, int(g["q0"]))
    sim.run(inv)
    assert abs(sim.get_probabilities(0, 1)[0] - 1.0) < 1e-10


# ---- BASELINE config C3: GHZ + QFT-style H / CRZ / Rz layers -------------------------------------------

@pytest.mark.parametrize("n,window", [(20, None), (22, None), (24, None), (22, 8)])
def test_config_c3_qft_style_against_oracle(n, window):
    """The C3 generator at sizes the oracle finishes in seconds; the CRZ ladders must take the fused-diagonal path
    (OP_PHASE: one multiplicative phase polynomial per run) and still match the gate-by-gate oracle."""
    c = H.qft_style_circuit(n, window)
    prog = q.CompiledCircuit(c)
    assert "PHASE" in prog.describe(), "the CRZ/Rz runs were expected to fuse into OP_PHASE ops"
    assert prog.n_passes < c.get_gate_count() // 8
    sim = q.Simulator(n)
    sim.execute(prog)
    got = sim.get_state_vector()
    want = H.oracle_run(n, c.gates)
    assert np.max(np.abs(got - want)) < TOL
    # the uncompiled entry point (run from host gate records) takes the same path
    sim2 = q.Simulator(n)
    sim2.run(c)
    assert np.max(np.abs(sim2.get_state_vector() - want)) < TOL
    # and sampling on that state is bit-identical to the sequential CDF of the reference (src/Simulator.cu:164-185)
    u = np.random.default_rng(n).random(512)
    assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(H.oracle_probs(got), u))


def test_config_c3_30q_properties():
    """C3 at 30 qubits (16 GiB) through size-independent properties: analytic GHZ amplitudes, unit norm after the
    QFT-style layers, and the inverse layers bring the GHZ state back."""
    n = 30
    sim = q.Simulator(n)
    sim.run(q.create_ghz_circuit(n))
    last = (1 << n) - 1
    r = 1.0 / np.sqrt(2.0)
    assert abs(sim.get_probabilities(0, 1)[0] - 0.5) < 1e-12 and abs(sim.get_probabilities(last, 1)[0] - 0.5) < 1e-12
    assert abs(sim.get_total_probability() - 1.0) < 1e-12
    full = H.qft_style_circuit(n)
    layers = full.gates[n:]                                  # everything after the GHZ ladder
    sim.run(q.Circuit(n).extend(layers))
    assert abs(sim.get_total_probability() - 1.0) < 1e-10
    inv = q.Circuit(n)
    for g in layers[::-1]:
        t = int(g["type"])
        if t == 3:
            inv.h(int(g["q0"]))
        elif t == 10:
            inv.rz(int(g["q0"]), -float(g["param"]))
        else:
            assert t == 14
            inv.crz(int(g["q0"]), int(g["q1"]), -float(g["param"]))
    sim.run(inv)
    p0, p1 = sim.get_probabilities(0, 1)[0], sim.get_probabilities(last, 1)[0]
    assert abs(p0 - 0.5) < 1e-10 and abs(p1 - 0.5) < 1e-10
    m = sim.marginal([0, n - 1])                             # GHZ correlations: only 00 and 11
    assert np.allclose(m, [0.5, 0, 0, 0.5], atol=1e-10)
