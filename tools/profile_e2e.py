"""The bench's e2e step (reset -> run -> sample 1024) twice at 30 qubits, for an ncu launch list of its kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_quantum_simulator_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sim = q.Simulator(n)
c = q.create_random_circuit(n, 20, 42)
u = np.random.default_rng(0).random(1024)
sim.run(c)
q.jit_wait()
for _ in range(2):
    sim.reset()
    sim.run(c)
    idx = sim.sample(0, uniforms=u)
print("e2e ok", idx[:4])
