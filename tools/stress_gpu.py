"""Randomised GPU-vs-oracle stress beyond what the test suite affords (development aid): deep circuits over all gate
types at 13..22 qubits, from recorded basis states and from random states, run() and compiled programs, with sampling."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_quantum_simulator_b200 as q
import helpers as H

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
t0 = time.time()
n_cases, worst = 0, 0.0
while time.time() - t0 < budget:
    n = int(rng.integers(13, 23))
    d = int(rng.integers(20, 400))
    kinds = None if rng.random() < 0.5 else [0, 3, 3, 8, 9, 10, 11, 11, 12, 15, 16, 5]   # flip/H heavy mix
    g = H.random_gates(n, d, rng, kinds=kinds)
    sim = q.Simulator(n)
    mode = int(rng.integers(0, 3))
    if mode == 0:
        st0 = H.random_state(n, rng)
        sim.set_state(st0)
    else:
        idx = int(rng.integers(0, 1 << n)) if mode == 2 else 0
        st0 = np.zeros(1 << n, np.complex128)
        st0[idx] = 1.0
        sim.init_basis(idx) if mode == 2 else sim.reset()
    c = q.Circuit(n).extend(g)
    if rng.random() < 0.5:
        sim.run(c)
    else:
        sim.execute(q.CompiledCircuit(c))
    want = H.oracle_run(n, g, st0)
    u = rng.random(128)
    s = sim.sample(0, uniforms=u)
    got = sim.get_state_vector()
    err = float(np.max(np.abs(got - want)))
    worst = max(worst, err)
    assert err < 1e-10, (n, d, mode, err)
    assert np.array_equal(s, H.oracle_sample(H.oracle_probs(got), u)), (n, d, mode, "sample")
    n_cases += 1
print(f"stress ok: {n_cases} circuits, worst max|err| {worst:.2e}")
