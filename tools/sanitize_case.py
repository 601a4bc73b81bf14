"""A small tour of the round-2 kernels for compute-sanitizer (memcheck / racecheck): specialised pass kernels (forced for every
pass), CUDA-graph replay, fused diagonal runs with compact tables, the one-CTA sampler, the trajectory kernel with damping
runs and the batched sampler.  usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_quantum_simulator_b200 as q
import helpers as H

q.jit_set_mode("always")
rng = np.random.default_rng(1)
for n in (6, 13, 15):
    g = H.random_gates(n, 60, rng)
    sim = q.Simulator(n)
    sim.run(q.Circuit(n).extend(g))
    err = float(np.max(np.abs(sim.get_state_vector() - H.oracle_run(n, g))))
    assert err < 1e-12, err
    u = rng.random(64)
    assert np.array_equal(sim.sample(0, uniforms=u), H.oracle_sample(H.oracle_probs(sim.get_state_vector()), u))
c = H.qft_style_circuit(14)
prog = q.CompiledCircuit(c, specialise=True)
sim = q.Simulator(14)
for _ in range(4):
    sim.execute(prog)          # plain, plain, captured, replayed
want = H.zero_state(14)
for _ in range(4):
    want = H.oracle_run(14, c.gates, want)
assert float(np.max(np.abs(sim.get_state_vector() - want))) < 1e-10
noise = q.NoiseModel().add_depolarizing(0.05).add_amplitude_damping(0.1).add_phase_damping(0.05, [1, 9])
b = q.BatchedSimulator(10, 64, noise)
b.set_seed(7)
b.run(q.create_ghz_circuit(10))
assert abs(b.get_average_probabilities().sum() - 1) < 1e-10 and b.get_histogram(3).sum() == 192
print("sanitize tour ok", q.jit_stats())
