"""Per-pass device times of the fused-pass kernel for a few circuit shapes (development aid).
Tunables via env: QSIM_STAGES, QSIM_TILE_BITS, QSIM_MIN_LOW_BITS.   usage: pass_microbench.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cuda_quantum_simulator_b200 as q

if os.environ.get("QSIM_LIB"):   # development: load another build of the library
    q._lib.LIB_PATH = os.path.abspath(os.environ["QSIM_LIB"])

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = 5
sim = q.Simulator(n)
byt = 2 * 16 * (1 << n)


def run(name, circ):
    prog = q.CompiledCircuit(circ)
    for _ in range(2):
        sim.execute(prog)
    sim.synchronize()
    sim.set_timing(True)
    for _ in range(reps):
        sim.execute(prog)
    sim.synchronize()
    t = sim.pass_times_ms().reshape(reps, -1).mean(axis=0)
    sim.set_timing(False)
    print(f"{name:28s} passes={prog.n_passes} ops={prog.n_ops} sweeps={prog.n_sweeps} | " +
          " ".join(f"{x:7.3f}ms({byt / x / 1e6:6.0f}GB/s)" for x in t), flush=True)


C = q.Circuit
run("1 diag op, low tile", C(n).z(0))
run("1 H reg-bit (q5)", C(n).h(5))
run("1 H lane-bit (q0)", C(n).h(0))
run("4 H lane bits", C(n).h(0).h(1).h(2).h(3))
run("8 H (q0..7)", (lambda c: [c.h(i) for i in range(8)] and c)(C(n)))
run("12 H (q0..11) 2 sweeps", (lambda c: [c.h(i) for i in range(12)] and c)(C(n)))
run("1 H high (q n-1)", C(n).h(n - 1))
run("3 H high scattered", C(n).h(n - 1).h(n - 5).h(n - 9))
run("7 H high scattered", (lambda c: [c.h(n - 1 - 2 * i) for i in range(7)] and c)(C(n)))
run("X high (q n-1)", C(n).x(n - 1))
for k in (3, 4, 5, 6, 7):
    run(f"{k} X high scattered (L={12-k})", (lambda c: [c.x(n - 1 - 2 * i) for i in range(k)] and c)(C(n)))
run("7 Z high scattered (diag)", (lambda c: [c.z(n - 1 - 2 * i) for i in range(7)] and c)(C(n)))
run("12 X (q0..11) 2 sweeps", (lambda c: [c.x(i) for i in range(12)] and c)(C(n)))
run("9 X (q0..8) 1 sweep", (lambda c: [c.x(i) for i in range(9)] and c)(C(n)))
run("C2 random(20,42)", q.create_random_circuit(n, 20, 42))
run("random depth 200 seed 1", q.create_random_circuit(n, 200, 1))
