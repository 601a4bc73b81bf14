"""Per-kind op cost: many ops of one kind in one pass (QSIM_NO_MERGE=1 keeps them separate)."""
import os, sys
os.environ["QSIM_NO_MERGE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_quantum_simulator_b200 as q
n=30; sim=q.Simulator(n); reps=3
def run(name, circ):
    prog=q.CompiledCircuit(circ)
    for _ in range(2): sim.execute(prog)
    sim.synchronize(); sim.set_timing(True)
    for _ in range(reps): sim.execute(prog)
    sim.synchronize()
    t=sim.pass_times_ms().reshape(reps,-1).mean(axis=0); sim.set_timing(False)
    print(f"{name:40s} ops={prog.n_ops:3d} sweeps={prog.n_sweeps} passes={prog.n_passes} | "+" ".join(f"{x:7.3f}" for x in t)+f"  per-op {(t.sum()-2.8)/max(prog.n_ops,1):.3f} ms", flush=True)
C=q.Circuit
def rep(fn, k):
    c=C(n)
    for i in range(k): fn(c,i)
    return c
K=24
run("reg MATREAL  (H on q5,6,7 x8)", rep(lambda c,i: c.h(5+i%3), K))
run("reg MAT      (Rx on q5,6,7)", rep(lambda c,i: c.rx(5+i%3,0.3), K))
run("lane MATREAL (H on q0,1,2)", rep(lambda c,i: c.h(i%3), K))
run("lane MAT     (Rx on q0,1,2)", rep(lambda c,i: c.rx(i%3,0.3), K))
run("diag uniform (Rz on q20)", rep(lambda c,i: c.rz(20+i%3,0.3), K))
run("diag reg     (Rz on q5,6,7)", rep(lambda c,i: c.rz(5+i%3,0.3), K))
run("CNOT reg tgt, thread ctrl (c=9+,t=5..7)", rep(lambda c,i: c.cnot(9+i%3,5+i%3), K))
run("CNOT lane tgt (c=9, t=0..2)", rep(lambda c,i: c.cnot(9,i%3), K))
run("CNOT outside ctrl (c=25,t=5..7)", rep(lambda c,i: c.cnot(25,5+i%3), K))
run("CZ (9,5..7)", rep(lambda c,i: c.cz(9,5+i%3), K))
run("CRZ outside (25,26)", rep(lambda c,i: c.crz(25+i%2,20+i%3,0.2), K))
run("Y reg (ADIAG)", rep(lambda c,i: c.y(5+i%3), K))
