"""Runs one circuit shape a few times (for ncu).  usage: profile_case.py <case> [n] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cuda_quantum_simulator_b200 as q

case = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
C = q.Circuit
cases = {
    "c2": lambda: q.create_random_circuit(n, 20, 42),
    "h12": lambda: (lambda c: [c.h(i) for i in range(12)] and c)(C(n)),
    "h4lane": lambda: C(n).h(0).h(1).h(2).h(3),
    "z1": lambda: C(n).z(0),
    "x7high": lambda: (lambda c: [c.x(n - 1 - 2 * i) for i in range(7)] and c)(C(n)),
    "d200": lambda: q.create_random_circuit(n, 200, 1),
    "dense": lambda: q.create_random_circuit(n, 200, 42),
    "c3": lambda: __import__("helpers").qft_style_circuit(n),
}
sim = q.Simulator(n)
prog = q.CompiledCircuit(cases[case]())
print(prog.describe())
sim.execute(prog)
q.jit_wait()                   # the specialised kernels are ready before the profiled launches
sim.execute(prog)              # (the first pass, no longer fed a basis state, may want the other build of its kernel)
q.jit_wait()
for _ in range(reps):
    sim.execute(prog)
sim.synchronize()
print("done", case, n, "passes", prog.n_passes)
