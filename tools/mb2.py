import os, sys
sys.path.insert(0, '/root/repo'); 
import cuda_quantum_simulator_b200 as q
n=30; sim=q.Simulator(n); byt=2*16*(1<<n); reps=4
def run(name, circ):
    prog=q.CompiledCircuit(circ)
    for _ in range(2): sim.execute(prog)
    sim.synchronize(); sim.set_timing(True)
    for _ in range(reps): sim.execute(prog)
    sim.synchronize()
    t=sim.pass_times_ms().reshape(reps,-1).mean(axis=0); sim.set_timing(False)
    print(f"{name:34s} ops={prog.n_ops} sweeps={prog.n_sweeps} | "+" ".join(f"{x:7.3f}ms" for x in t), flush=True)
C=q.Circuit
def hs(qs):
    c=C(n)
    for x in qs: c.h(x)
    return c
run("H q29", hs([29]))
run("H q29,27", hs([29,27]))
run("H q29,27,25 (3 regs)", hs([29,27,25]))
run("H q29..23 (3 reg + 1 lane)", hs([29,27,25,23]))
run("H q29..21 (3 reg + 2 lane)", hs([29,27,25,23,21]))
run("H q29..19 (6: 2 sweeps)", hs([29,27,25,23,21,19]))
run("H 5,6,7 (3 reg low)", hs([5,6,7]))
run("H 5,6,7,8 ", hs([5,6,7,8]))
run("H 5..9 ", hs([5,6,7,8,9]))
run("H 5..10 (2 sweeps)", hs([5,6,7,8,9,10]))
run("H 3..10 (8)", hs(list(range(3,11))))
run("H 0,1,2", hs([0,1,2]))
run("Z 29..17 (7 diag)", (lambda c:[c.z(29-2*i) for i in range(7)] and c)(C(n)))
run("Rz 5", C(n).rz(5,0.3))
run("CNOT(29,27) CNOT(25,23) CNOT(21,19)", C(n).cnot(29,27).cnot(25,23).cnot(21,19))
