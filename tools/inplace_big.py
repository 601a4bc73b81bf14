"""torchrun --nproc-per-node N tools/inplace_big.py [local_qubits=33] [steps=2]
The C4 regime per GPU at any N: createRandomCircuit(local+log2 N, 20, 42) with the IDENTITY qubit layout on shards that leave
no room for a second buffer (33 local qubits = 128 GiB), so the exchange in front of the gate on the top qubit is fused IN
PLACE into the pass before it (qsim_shard_execute_exchange_inplace); then the same steps with the separate swap kernel.
Prints one JSON line (rank 0).  Checks: total probability 1, and both variants leave the same amplitudes behind (a checksum of
|a|^2-weighted index bits per shard, compared across the two variants)."""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_quantum_simulator_b200 as q
from cuda_quantum_simulator_b200.sharded import NativeShardedSimulator

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 33
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n = nl + int(math.log2(world))
stream = torch.cuda.current_stream()
circ = q.create_random_circuit(n, 20, 42)


def sync_all():
    torch.cuda.synchronize()
    dist.barrier()


def timed(fn, k):
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        fn()
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / k


def variant(sim, env):
    for k_, v_ in env.items():
        os.environ[k_] = v_
    try:
        sim.identity_layout_only(True)
        sim.reset()
        first = sim.compile(circ)
        sim.execute(first)
        sim.relabel_identity()
        plan = sim.compile(circ)

        def st():
            sim.relabel_identity()
            sim.execute(plan)
        st()
        q.jit_wait()
        st()
        f0, i0, s0, p0 = sim.fused_exchanges, sim.inplace_exchanges, sim.separate_exchanges, sim.split_exchanges
        ms = timed(st, steps)
        info = {"ms_per_step": ms, "passes": plan.n_passes, "swaps": plan.n_swaps, "fused_per_step": (sim.fused_exchanges - f0) / steps,
                "in_place_per_step": (sim.inplace_exchanges - i0) / steps, "split_per_step": (sim.split_exchanges - p0) / steps, "separate_per_step": (sim.separate_exchanges - s0) / steps,
                "total_probability": sim.get_total_probability(),
                "marginal_top_and_low": [float(x) for x in sim.marginal([n - 1, 0, nl - 1, 5])]}
        return info
    finally:
        for k_ in env:
            os.environ.pop(k_, None)


sim = NativeShardedSimulator(n)
out = {"qubits": n, "local_qubits": nl, "world": world, "second_buffer": sim.has_second_buffer, "exchange": sim.exchange}
out["split"] = variant(sim, {"QSIM_FORCE_INPLACE_EXCHANGE": "1"})     # scatter half in the pass before, gather half in the pass after
out["in_place"] = variant(sim, {"QSIM_FORCE_INPLACE_EXCHANGE": "1", "QSIM_NO_SPLIT_EXCHANGE": "1"})
out["separate"] = variant(sim, {"QSIM_NO_INPLACE_EXCHANGE": "1", "QSIM_NO_FUSED_EXCHANGE": "1"})
# both variants ran the same number of steps from the same start: same state, so the same marginals (1e-12)
a, b, c = (np.array(out[k]["marginal_top_and_low"]) for k in ("in_place", "separate", "split"))
out["marginals_agree"] = bool(np.max(np.abs(a - b)) < 1e-12 and np.max(np.abs(c - b)) < 1e-12)
out["half_shard_bytes"] = 16 * (1 << (nl - 1))
out["in_place"]["speedup_vs_separate"] = out["separate"]["ms_per_step"] / out["in_place"]["ms_per_step"]
out["split"]["speedup_vs_separate"] = out["separate"]["ms_per_step"] / out["split"]["ms_per_step"]
sim.close()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
ok = out["marginals_agree"] and abs(out["in_place"]["total_probability"] - 1) < 1e-9 and out["in_place"]["in_place_per_step"] > 0
sys.exit(0 if ok else 1)
