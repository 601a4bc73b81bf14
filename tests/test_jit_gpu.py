"""The run-time specialised pass kernels (csrc/jit.cpp) against the CPU oracle.  By default only passes over >= 26 local
qubits are specialised (the 26/30-qubit tests elsewhere take that path); here the mode is forced to "always", so every pass
of every circuit - all gate kinds, controls in every home, partial tiles, folded flips, fused diagonals, basis-state input,
the redirected store of the fused exchange - runs through generated code, and a failed compile is an error."""
import numpy as np
import pytest

import cuda_quantum_simulator_b200 as q
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def jit_always():
    q.jit_set_mode("always")
    yield
    q.jit_set_mode("auto")


def run_gpu(n, g, state=None):
    sim = q.Simulator(n)
    if state is not None:
        sim.set_state(state)
    sim.run(q.Circuit(n).extend(g) if len(g) else q.Circuit(n))
    return sim.get_state_vector(), sim


def test_specialised_kernels_are_what_runs():
    before = q.jit_stats()
    n = 10
    g = H.random_gates(n, 40, np.random.default_rng(1))
    got, sim = run_gpu(n, g)
    after = q.jit_stats()
    assert after["mode"] == "always"
    assert after["launches"] > before["launches"] and after["failures"] == before["failures"]
    made = lambda st: st["compiles"] + st["cache_hits"] + st["disk_hits"]     # compiled now, or by an earlier run / process
    assert made(after) > made(before)
    assert np.max(np.abs(got - H.oracle_run(n, g))) < 1e-12
    # the same structure with other angles is a cache hit, not a compile
    g2 = g.copy()
    g2["param"] = np.where(g2["param"] != 0, g2["param"] * 0.37 + 0.1, 0.0)
    c0 = q.jit_stats()["compiles"]
    got2, _ = run_gpu(n, g2)
    assert np.max(np.abs(got2 - H.oracle_run(n, g2))) < 1e-12


@pytest.mark.parametrize("case", H.load_known_answers()["cases"], ids=lambda c: c["name"])
def test_known_answers(case):
    from test_oracle import _check_expect
    g = H.gates([tuple(x) for x in case["gates"]])
    st, _ = run_gpu(case["n"], g)
    _check_expect(st, case["expect"], case["tol"])


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 8, 9, 11, 12, 13, 14, 16])
def test_fuzz_all_gate_types_from_random_state(n):
    rng = np.random.default_rng(500 + n)
    g = H.random_gates(n, int(rng.integers(20, 90)), rng)
    st0 = H.random_state(n, rng)
    got, _ = run_gpu(n, g, st0)
    assert np.max(np.abs(got - H.oracle_run(n, g, st0))) < 1e-12


@pytest.mark.parametrize("n", [7, 13, 15, 18])
def test_flips_fold_into_loads_and_stores(n):
    rng = np.random.default_rng(40 + n)
    for body, tail in ((6, 25), (0, 30), (12, 8)):
        g = H.flip_heavy_gates(n, rng, body, tail)
        st0 = H.random_state(n, rng)
        got, _ = run_gpu(n, g, st0)
        assert np.max(np.abs(got - H.oracle_run(n, g, st0))) < 1e-12
        g2 = np.concatenate([g[::-1], g])          # flips leading a pass, dense gates, flips trailing it
        got, _ = run_gpu(n, g2, st0)
        assert np.max(np.abs(got - H.oracle_run(n, g2, st0))) < 1e-12


@pytest.mark.parametrize("n", [9, 14, 19])
def test_run_from_a_recorded_basis_state(n):
    rng = np.random.default_rng(70 + n)
    g = H.random_gates(n, 50, rng)
    sim = q.Simulator(n)
    idx = int(rng.integers(1 << n))
    sim.init_basis(idx)
    sim.run(q.Circuit(n).extend(g))
    st0 = np.zeros(1 << n, np.complex128)
    st0[idx] = 1
    assert np.max(np.abs(sim.get_state_vector() - H.oracle_run(n, g, st0))) < 1e-12


def test_every_gate_on_every_qubit_16q():
    n = 16
    rng = np.random.default_rng(9)
    st0 = H.random_state(n, rng)
    for t in range(17):
        lst = []
        for qb in range(n):
            others = [x for x in range(n) if x != qb]
            a, b = (int(x) for x in rng.choice(others, 2, replace=False))
            ang = float(rng.uniform(0, 6))
            if t <= 7: lst.append((t, qb))
            elif t <= 10: lst.append((t, qb, ang))
            elif t in (11, 12, 15): lst.append((t, a, qb))
            elif t in (13, 14): lst.append((t, a, qb, ang))
            else: lst.append((t, a, b, qb))
        g = H.gates(lst)
        got, _ = run_gpu(n, g, st0)
        assert np.max(np.abs(got - H.oracle_run(n, g, st0))) < 1e-12, H.NAMES[t]


def test_fused_diagonal_runs_c3_20q():
    n = 20
    c = H.qft_style_circuit(n)
    prog = q.CompiledCircuit(c)
    assert "PHASE" in prog.describe()
    sim = q.Simulator(n)
    sim.execute(prog)
    assert np.max(np.abs(sim.get_state_vector() - H.oracle_run(n, c.gates))) < 1e-10


def test_config_c1_and_c2_shapes():
    g = H.bench_c1_gates(20)
    got, _ = run_gpu(20, g)
    assert np.max(np.abs(got - H.oracle_run(20, g))) < 1e-10
    c = q.create_random_circuit(24, 200, 42)
    sim = q.Simulator(24)
    sim.run(c)
    assert np.max(np.abs(sim.get_state_vector() - H.oracle_run(24, c.gates))) < 1e-10


@pytest.mark.parametrize("nl,g_local", [(14, 13), (20, 17), (21, 0)])
def test_fused_exchange_store_redirect(nl, g_local):
    """The specialised kernel shares the skeleton's redirected store (fused qubit exchange): reuse the one-GPU check."""
    from test_sharded_gpu import test_fused_exchange_on_one_gpu
    test_fused_exchange_on_one_gpu(nl, g_local, False)


@pytest.mark.parametrize("mode", ["always", "off"])
def test_multi_pass_program_replays_as_a_cuda_graph(mode):
    """Small states are launch-bound: from the third run on the same amplitudes a multi-pass compiled circuit
    is one CUDA-graph launch (specialised or interpreter kernels alike).  Five runs must equal five oracle applications."""
    q.jit_set_mode(mode)
    n = 15
    rng = np.random.default_rng(77)
    g = H.random_gates(n, 150, rng)
    c = q.Circuit(n).extend(g)
    prog = q.CompiledCircuit(c, specialise=(mode == "always"))
    assert prog.n_passes >= 2
    sim = q.Simulator(n)
    st0 = H.random_state(n, rng)
    sim.set_state(st0)
    want = st0
    for k in range(5):
        sim.execute(prog)
        want = H.oracle_run(n, g, want)
        assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-11, k
    # another simulator, same program: captured again for the other amplitudes
    sim2 = q.Simulator(n)
    sim2.set_state(st0)
    want = st0
    for k in range(4):
        sim2.execute(prog)
        want = H.oracle_run(n, g, want)
    assert np.max(np.abs(sim2.get_state_vector() - want)) < 1e-11
    sim.execute(prog)      # and back on the first one
    assert sim.launch_count() > 0


def test_background_compilation_never_blocks_a_run():
    """Default mode: a run of a new pass structure launches the ahead-of-time kernel at once while a background thread
    compiles the specialised one; once it is ready (jit_wait) the next run uses it.  Both give the oracle's state."""
    q.jit_set_mode("auto", 10)                       # specialise from 10 qubits on (default 26), asynchronously
    try:
        n = 12
        rng = np.random.default_rng(4242)
        g = H.random_gates(n, 70, rng)
        g["param"] = np.where(g["param"] != 0, g["param"] + 0.123, 0.0)
        c = q.Circuit(n).extend(g)
        want = H.oracle_run(n, g)
        sim = q.Simulator(n)
        before = q.jit_stats()
        sim.run(c)                                   # interpreter kernels (unless an earlier process left the cubins on disk)
        first = q.jit_stats()
        assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-12
        q.jit_wait()
        ready = q.jit_stats()
        assert ready["failures"] == before["failures"]
        sim.reset()
        sim.run(c)                                   # specialised kernels now
        after = q.jit_stats()
        assert after["launches"] > first["launches"]
        assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-12
        made = lambda st: st["compiles"] + st["disk_hits"] + st["cache_hits"]
        assert made(after) > made(before)
    finally:
        q.jit_set_mode("auto", 26)


@pytest.mark.parametrize("n,depth,seed,from_state", [(20, 150, 1, True), (20, 60, 2, False), (21, 200, 3, True), (22, 120, 4, True)])
def test_two_warp_group_kernels(n, depth, seed, from_state):
    """The TWO-WARP-GROUP build of the specialised kernels (pass_kernel_body.inc QSIM_DUAL_GROUPS; what compute-heavy passes of
    large states get), forced for every pass that can take it: all gate kinds, folded flips (sweeps that permute the tile keep
    the first half in registers), fused diagonals (four factor copies), deferred X on outer bits (tile pairs across the two
    groups), several tiles per CTA.  Against the oracle, and bit-identical run to run (the groups' interleaving must not matter)."""
    q.jit_set_dual("always")
    try:
        rng = np.random.default_rng(7000 + seed)
        g = H.random_gates(n, depth, rng)
        if seed % 2 == 1:
            g = np.concatenate([g, H.gates([("X", n - 1), ("X", n - 2), ("X", 3)])])   # deferred X: partner tiles, local XOR
        st0 = H.random_state(n, rng) if from_state else None
        before = q.jit_stats()["launches"]
        got, sim = run_gpu(n, g, st0)
        assert q.jit_stats()["launches"] > before
        want = H.oracle_run(n, g, st0)
        assert np.max(np.abs(got - want)) < 1e-12
        prog = q.CompiledCircuit(q.Circuit(n).extend(g), specialise=True)
        outs = []
        for _ in range(3):
            s2 = q.Simulator(n)
            if st0 is not None:
                s2.set_state(st0)
            else:
                s2.set_state(H.zero_state(n))
            s2.execute(prog)
            outs.append(s2.get_state_vector())
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])
        assert np.max(np.abs(outs[0] - want)) < 1e-12
    finally:
        q.jit_set_dual("auto")


def test_two_warp_group_kernels_c3_shape():
    """C3-style layers (fused diagonal runs with factors that depend on bits outside the tile) through the two-group build."""
    q.jit_set_dual("always")
    try:
        n = 21
        c = H.qft_style_circuit(n)
        sim = q.Simulator(n)
        sim.run(c)
        want = H.oracle_run(n, c.gates)
        assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-12
    finally:
        q.jit_set_dual("auto")


@pytest.mark.parametrize("n,seed", [(14, 0), (16, 1), (17, 2), (20, 3), (21, 4)])
def test_conditional_flips_carried_to_the_store(n, seed):
    """CNOT / X-type flips whose controls sit in thread bits or outside the tile and whose target is a register bit are not
    executed as register moves by the specialised kernels: the thread carries an XOR mask over its slot index to the sweep's
    store, later 2x2 ops on the same bit take the X-conjugated matrix, diagonals the swapped factors, and ops that cannot live
    with a pending flip (register-bit controls, fused diagonal runs, a lane op on a lane bit that conditioned it) materialise it
    first.  CNOT-heavy random circuits over every gate kind, from a random state, both builds of the kernel."""
    rng = np.random.default_rng(8100 + seed)
    kinds = [0, 3, 3, 8, 9, 11, 11, 11, 11, 12, 13, 14, 15, 16, 5, 6]     # H, Rz..., many CNOTs, CZ, SWAP, CRY, CRZ, Toffoli
    g = H.random_gates(n, 160, rng, kinds=kinds)
    c = q.Circuit(n).extend(g)
    prog = q.CompiledCircuit(c)
    assert any("carried as a slot-index XOR" in prog.jit_source(i) for i in range(prog.n_passes)), "no deferred flip in this circuit"
    st0 = H.random_state(n, rng)
    want = H.oracle_run(n, g, st0)
    got, _ = run_gpu(n, g, st0)
    assert np.max(np.abs(got - want)) < 1e-12
    if n >= 20:
        q.jit_set_dual("always")
        try:
            got2, _ = run_gpu(n, g, st0)
            assert np.max(np.abs(got2 - want)) < 1e-12
        finally:
            q.jit_set_dual("auto")
