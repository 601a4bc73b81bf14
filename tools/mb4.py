"""C2 anatomy: device time of createRandomCircuit(30,20,42) prefixes and compiler variants (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_quantum_simulator_b200 as q

if os.environ.get("QSIM_LIB"):   # development: load another build of the library
    q._lib.LIB_PATH = os.path.abspath(os.environ["QSIM_LIB"])

n = 30
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
    print(f"{name:34s} passes={prog.n_passes} ops={prog.n_ops} sweeps={prog.n_sweeps} | " +
          " ".join(f"{x:7.3f}ms" for x in t), flush=True)


full = q.create_random_circuit(n, 20, 42)
gates = full.gates
for env in ({}, {"QSIM_NO_TAIL": "1"}, {"QSIM_NO_PHASE": "1"}, {"QSIM_REG_BITS": "2"}):
    for k, v in env.items():
        os.environ[k] = v
    run(f"C2 {env}", full)
    for k in env:
        del os.environ[k]
for k in (2, 4, 6, 8, 10, 12, 14, 16, 18, 20):
    c = q.Circuit(n)
    c.extend(gates[:k])
    run(f"C2 first {k} gates", c)
    print(prog_desc := q.CompiledCircuit(c).describe().split("\n")[2] if k in (20,) else "", flush=True)
