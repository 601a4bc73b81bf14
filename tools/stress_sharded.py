"""torchrun --nproc-per-node N tools/stress_sharded.py [seconds]: randomised multi-GPU parity against the CPU oracle
(development aid): dense random circuits over all gate types, with and without the free initial layout, repeated runs on
the same simulator (carried permutation and X frame), fused and separate exchanges."""
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_quantum_simulator_b200 as q
from cuda_quantum_simulator_b200.sharded import ShardedSimulator
import helpers as H

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
# second argument "native" / "native-inplace": the C++ driver (qsim::ShardedSimulator); -inplace forces the fused exchange in
# place (no second buffer) wherever a pass can carry it
NATIVE = len(sys.argv) > 2 and sys.argv[2].startswith("native")
if len(sys.argv) > 2 and sys.argv[2] == "native-inplace":
    os.environ["QSIM_FORCE_INPLACE_EXCHANGE"] = "1"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ng = int(math.log2(world))
rng = np.random.default_rng(99)          # same stream on every rank
t0 = time.time()
cases = swaps = fused = 0
worst = 0.0
go = torch.ones(1, device="cuda")
while True:
    go[0] = 1.0 if time.time() - t0 < budget else 0.0
    dist.broadcast(go, 0)
    if go.item() == 0.0:
        break
    n = int(rng.integers((17 if NATIVE else 14) + ng, (23 if NATIVE else 21) + ng))
    d = int(rng.integers(20, 250))
    kinds = None if rng.random() < 0.5 else [0, 3, 3, 8, 9, 10, 11, 11, 12, 15, 16, 5]
    g1, g2 = H.random_gates(n, d, rng, kinds=kinds), H.random_gates(n, int(rng.integers(5, 80)), rng, kinds=kinds)
    pristine = bool(rng.random() < 0.5)
    if cases < int(os.environ.get("START_CASE", "0")):
        cases += 1
        continue
    if NATIVE:
        from cuda_quantum_simulator_b200.sharded import NativeShardedSimulator
        sim = NativeShardedSimulator(n)
        sim.identity_layout_only(not pristine)
        for g in (g1, g2):
            cp = sim.compile(q.Circuit(n).extend(g))
            swaps += cp.n_swaps
            sim.execute(cp)
        fused += sim.fused_exchanges
        inplace = globals().get("inplace", 0) + sim.inplace_exchanges
        split = globals().get("split", 0) + sim.split_exchanges
    else:
        sim = ShardedSimulator(n)
        sim._pristine = pristine
        f0 = sim.engine.fused_exchanges
        for g in (g1, g2):                    # the second run starts from the carried permutation / frame
            cp = sim.compile(q.Circuit(n).extend(g))
            swaps += cp.n_swaps
            sim.execute(cp)
            if os.environ.get("SYNC_BETWEEN_RUNS"):
                torch.cuda.synchronize()
            if not os.environ.get("NO_RELEASE"):
                sim.release(cp)
        fused += sim.engine.fused_exchanges - f0
    got = sim.get_state_vector()
    want = H.oracle_run(n, g2, H.oracle_run(n, g1))
    err = float(np.max(np.abs(got - want)))
    worst = max(worst, err)
    assert err < 1e-10, (n, d, err)
    u = np.random.default_rng(cases).random(64)
    s = sim.sample(uniforms=u)
    assert np.all(np.abs(want[s]) ** 2 > 0)
    assert abs(sim.get_total_probability() - 1) < 1e-9
    sim.close()
    cases += 1
if rank == 0:
    print(f"sharded stress ok on {world} GPUs ({'C++ driver' if NATIVE else 'Python driver'}): {cases} cases, {swaps} exchanges ({fused} fused into a pass, "
          f"{globals().get('inplace', 0)} of them in place, {globals().get('split', 0)} split over two passes), worst max|err| {worst:.2e}")
dist.destroy_process_group()
