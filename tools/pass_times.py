"""Per-pass device times of a compiled circuit (CUDA events around every pass launch).
usage: pass_times.py c2|dense|c3|c3w [n] -> one JSON line (evidence for profiles/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_quantum_simulator_b200 as q
import helpers as H

name = sys.argv[1] if len(sys.argv) > 1 else "dense"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
circ = {"c2": lambda: q.create_random_circuit(n, 20, 42), "dense": lambda: q.create_random_circuit(n, 200, 42),
        "c3": lambda: H.qft_style_circuit(n), "c3w": lambda: H.qft_style_circuit(n, 8)}[name]()
prog = q.CompiledCircuit(circ)
sim = q.Simulator(n)
sim.execute(prog)          # from |0..0> (queues the specialised kernels for compilation)
q.jit_wait()
for _ in range(2):         # (a first pass that no longer starts from a basis state may ask for another build of its kernel;
    sim.execute(prog)      #  a pass that has two builds asks for the second one once the first is there)
    q.jit_wait()
for _ in range(4):         # the two builds of a heavy pass are timed against each other in passing; then the faster one stays
    sim.execute(prog)
sim.synchronize()
sim.set_timing(True)
reps = 3
for _ in range(reps):
    sim.execute(prog)
sim.synchronize()
t = sim.pass_times_ms().reshape(reps, -1).mean(axis=0)
byt = 2 * 16 * (1 << n)
print(json.dumps({"circuit": name, "qubits": n, "gates": circ.get_gate_count(), "passes": prog.n_passes, "ops": prog.n_ops,
                  "jit": q.jit_stats(), "pass_ms": [round(float(x), 3) for x in t], "total_ms": float(t.sum()),
                  "hbm_gbs_per_pass": [round(byt / (x * 1e-3) / 1e9) for x in t]}))
print(prog.describe(), file=sys.stderr)
