"""Runs the read-out and batched-trajectory kernels once each on realistic sizes (for an ncu launch list / per-kernel capture).
usage: profile_readout.py [n=30] [batch=65536]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_quantum_simulator_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
sim = q.Simulator(n)
sim.run(q.create_random_circuit(n, 200, 42))        # a dense state
u = np.random.default_rng(0).random(1024)
idx = sim.sample(0, uniforms=u)                     # chunk_approx / scan / classify / stitch / sample kernels
m = sim.marginal([0, 5, n - 1, n // 2, 7, n - 2])   # marginal kernels
tot = sim.get_total_probability()
res = sim.measure_qubit(3, 0.4)                     # exact p0 (sequential CDF) + collapse
print("readout", n, idx[:4], float(m.sum()), tot, res)
del sim
noise = q.NoiseModel().add_depolarizing(0.005).add_amplitude_damping(0.001)
b = q.BatchedSimulator(12, batch, noise)
b.set_seed(42)
b.run(q.create_ghz_circuit(12))                     # trajectory_kernel (average in the epilogue)
avg = b.get_average_probabilities()
hist = b.get_histogram(1)                           # batched_sample_kernel (histogram fused)
print("batched", batch, float(avg.sum()), int(hist.sum()))
