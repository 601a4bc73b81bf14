"""Host-side timing breakdown of the bench's e2e step (reset / run / sample) at n qubits (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_quantum_simulator_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sim = q.Simulator(n)
c = q.create_random_circuit(n, 20, 42)
u = np.random.default_rng(0).random(1024)
acc = {}
def t(name, fn):
    sim.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); sim.synchronize(); t2 = time.perf_counter()
    acc.setdefault(name, []).append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    return r
for rep in range(6):
    t("reset", sim.reset)
    t("run", lambda: sim.run(c))
    t("sample", lambda: sim.sample(0, uniforms=u))
for k, v in acc.items():
    print(f"{k:8s} host-return {np.mean([x[0] for x in v[2:]]):8.3f} ms   complete {np.mean([x[1] for x in v[2:]]):8.3f} ms")
t0 = time.perf_counter()
for rep in range(5):
    sim.reset(); sim.run(c); sim.sample(0, uniforms=u)
sim.synchronize()
print(f"e2e step {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms")
prog = q.CompiledCircuit(c)
t0 = time.perf_counter()
for rep in range(20):
    q.CompiledCircuit(c)
print(f"compile+upload {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")

# read-out of a dense state (every amplitude non-zero)
d = q.Circuit(n)
for qb in range(n):
    d.h(qb)
for qb in range(0, n, 3):
    d.rz(qb, 0.37 * (qb + 1))
    d.ry(qb, 0.11 * (qb + 1))
sim.run(d)
sim.synchronize()
for rep in range(4):
    t0 = time.perf_counter(); sim.sample(0, uniforms=u); dt = time.perf_counter() - t0
print(f"dense-state sample {dt * 1e3:.3f} ms, total probability {sim.get_total_probability():.15f}")
