"""Host-side timing breakdown of reset / run / sample at n qubits (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_quantum_simulator_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sim = q.Simulator(n)
c = q.create_random_circuit(n, 20, 42)
u = np.random.default_rng(0).random(1024)
def t(name, fn, reps=3):
    for _ in range(reps):
        sim.synchronize(); t0 = time.perf_counter(); r = fn(); sim.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:10s} {dt*1e3:9.2f} ms", flush=True)
    return r
t("reset", sim.reset)
t("run", lambda: sim.run(c))
t("sample", lambda: sim.sample(0, uniforms=u))
t("totalprob", sim.get_total_probability)
t("measure", lambda: sim.measure_bit(3, 0.3))
