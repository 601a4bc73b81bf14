"""In-kernel timeline of the passes of the C1 circuit (QSIM_PASS_TIMELINE=1 must be set): where the fixed cost of a pass on a
small state goes.  usage: QSIM_PASS_TIMELINE=1 python tools/pass_timeline.py [n=20]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, ctypes
import cuda_quantum_simulator_b200 as q
from cuda_quantum_simulator_b200 import _lib
import helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = H.bench_c1_gates(n)
prog = q.CompiledCircuit(q.Circuit(n).extend(g), specialise=True)
sim = q.Simulator(n)
for _ in range(5): sim.execute(prog)
sim.synchronize()
out = (ctypes.c_uint64 * 512)(); k = ctypes.c_int64()
_lib.check(_lib.lib().qsim_sim_pass_timeline(sim._h, out, 512, ctypes.byref(k)))
v = np.array(out[:k.value], dtype=np.int64).reshape(-1, 8)
print("stamps (ns since kernel entry of CTA 0): set-up done, first loads issued, -, -, -, all tiles computed, stores complete; gap to next kernel entry")
for i, row in enumerate(v[-6:]):
    nxt = v[-6:][i + 1][0] - row[0] if i + 1 < 6 else -1
    print([int(x - row[0]) for x in row[1:]], "next kernel entry +", int(nxt))
