"""torchrun --nproc-per-node N tools/sharded_check.py [p2p|nccl]: multi-GPU parity against the CPU oracle and
NVLink exchange bandwidth (development / evidence script; also driven by tests/test_sharded_gpu.py)."""
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

exchange = sys.argv[1] if len(sys.argv) > 1 else "auto"
big = int(sys.argv[2]) if len(sys.argv) > 2 else 28
NATIVE = exchange.startswith("native")          # "native" / "native-nccl": the C++ driver (qsim::ShardedSimulator) over NCCL
INPLACE = exchange == "native-inplace"          # the fused exchange in place (no second buffer), forced for shards of any size
if INPLACE:
    os.environ["QSIM_FORCE_INPLACE_EXCHANGE"] = "1"
if NATIVE:
    exchange = "nccl" if exchange.endswith("nccl") else "auto"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ng = int(math.log2(world))

if NATIVE:
    from cuda_quantum_simulator_b200.sharded import NativeShardedSimulator
    worst = 0.0
    for seed, n, depth in [(1, 12, 80), (2, 16, 80), (3, 18, 120), (11, 22, 300)]:
        rng = np.random.default_rng(seed)
        g = H.random_gates(n, depth, rng) if n < 22 else q.create_random_circuit(n, depth, seed).gates
        sim = NativeShardedSimulator(n, exchange=exchange)
        c = q.Circuit(n).extend(g)
        sim.run(c)
        sim.run(c)
        got = sim.get_state_vector()
        want = H.oracle_run(n, g, H.oracle_run(n, g))
        err = float(np.max(np.abs(got - want)))
        worst = max(worst, err)
        assert abs(sim.get_total_probability() - 1) < 1e-10
        qs = [int(x) for x in np.random.default_rng(seed).permutation(n)[:5]]
        idx = np.arange(1 << n)
        outcome = np.zeros(1 << n, np.int64)
        for i, qb in enumerate(qs):
            outcome |= ((idx >> qb) & 1) << i
        assert np.max(np.abs(sim.marginal(qs) - np.bincount(outcome, weights=np.abs(want) ** 2, minlength=32))) < 1e-12
        # logical-order sampling after real exchanges: bit-identical to the reference's sequential CDF
        u = np.concatenate([np.random.default_rng(5).random(256), [0.0, 0.5]])
        assert np.array_equal(sim.sample(uniforms=u), H.oracle_sample(H.oracle_probs(got), u))
        assert np.array_equal(sim.get_state_vector(), got)
        # compiled plans, executed one per run; then a measurement
        sim.reset()
        plans = sim.compile_sequence(c, 2)
        for p_ in plans:
            sim.execute(p_)
        assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-10
        st = sim.get_state_vector()
        bit = n - 1
        p0 = float(np.sum((np.abs(st) ** 2)[((idx >> bit) & 1) == 0]))
        if min(p0, 1 - p0) > 1e-6:
            r = 0.5 if abs(p0 - 0.5) > 1e-6 else 0.3
            res = sim.measure_qubit(0, r)
            keep = ((idx >> bit) & 1) == res
            want_c = np.where(keep, st, 0) / np.sqrt(p0 if res == 0 else 1 - p0)
            assert res == (0 if r < p0 else 1) and np.max(np.abs(sim.get_state_vector() - want_c)) < 1e-10
        if rank == 0:
            print(f"native n={n} seed={seed} exchange={sim.exchange} swaps/run={plans[1].n_swaps} fused={sim.fused_exchanges} "
                  f"separate={sim.separate_exchanges} in-place={sim.inplace_exchanges} split={sim.split_exchanges} max|err|={err:.2e}", flush=True)
        assert not (INPLACE and n >= 22) or sim.inplace_exchanges > 0
        sim.close()
    assert worst < 1e-10, worst
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("sharded check ok", flush=True)
    sys.exit(0)

worst = 0.0
for seed, n in [(1, 12), (2, 16), (3, 18)]:
    rng = np.random.default_rng(seed)
    g = H.random_gates(n, 80, rng)
    sim = ShardedSimulator(n, exchange=exchange)
    c = q.Circuit(n).extend(g)
    sim.run(c)
    sim.run(c)
    got = sim.get_state_vector()
    want = H.oracle_run(n, g, H.oracle_run(n, g))
    err = float(np.max(np.abs(got - want)))
    worst = max(worst, err)
    u = np.random.default_rng(5).random(256)
    s = sim.sample(uniforms=u)
    assert np.all(np.abs(want[s]) ** 2 > 0)
    assert abs(sim.get_total_probability() - 1) < 1e-10
    qs = [int(x) for x in np.random.default_rng(seed).permutation(n)[:5]]
    idx = np.arange(1 << n)
    outcome = np.zeros(1 << n, np.int64)
    for i, qb in enumerate(qs):
        outcome |= ((idx >> qb) & 1) << i
    want_m = np.bincount(outcome, weights=np.abs(want) ** 2, minlength=32)
    assert np.max(np.abs(sim.marginal(qs) - want_m)) < 1e-12
    if rank == 0:
        print(f"n={n} seed={seed} exchange={sim.engine.exchange} swaps={sim.compile(c).n_swaps} "
              f"fused={sim.engine.fused_exchanges} max|err|={err:.2e}", flush=True)
    sim.close()
assert worst < 1e-10, worst

# the reference's generator at a size the oracle still handles, with a gate on the top (global) qubit
n = 24
c = q.create_random_circuit(n, 40, 7)
sim = ShardedSimulator(n, exchange=exchange)
sim.run(c)
got = sim.get_state_vector()
want = H.oracle_run(n, c.gates)
err = float(np.max(np.abs(got - want)))
if rank == 0:
    print(f"createRandomCircuit({n},40,7): swaps={sim.compile(c).n_swaps} fused={sim.engine.fused_exchanges} "
          f"max|err|={err:.2e}", flush=True)
assert err < 1e-10
sim.close()

# the same with the identity qubit layout forced (the free initial layout above avoids the exchange altogether), and a
# circuit that targets every qubit many times: global<->local swaps, fused into the preceding pass where possible
for label, n, depth, seed, free_layout in (("identity layout", 24, 40, 7, False), ("dense", 22, 300, 11, True)):
    c = q.create_random_circuit(n, depth, seed)
    sim = ShardedSimulator(n, exchange=exchange)
    sim._pristine = free_layout
    cp = sim.compile(c)
    sim.execute(cp)
    got = sim.get_state_vector()
    want = H.oracle_run(n, c.gates)
    err = float(np.max(np.abs(got - want)))
    if rank == 0:
        print(f"createRandomCircuit({n},{depth},{seed}) {label}: swaps={cp.n_swaps} fused={sim.engine.fused_exchanges} "
              f"max|err|={err:.2e}", flush=True)
    assert err < 1e-10
    sim.release(cp)
    sim.close()

# exchange bandwidth: swap the top global qubit with the top local qubit on 2^big-amplitude shards
n = big + ng
sim = ShardedSimulator(n, exchange=exchange)
eng = sim.engine
for _ in range(2):
    eng.swap(n - 1, big - 1)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 4
e0.record()
for _ in range(reps):
    eng.swap(n - 1, big - 1)
e1.record()
torch.cuda.synchronize(); dist.barrier()
ms = e0.elapsed_time(e1) / reps
half = 16 * (1 << (big - 1))
if rank == 0:
    print(f"swap of a {16 * (1 << big) / 2**30:.1f} GiB shard ({eng.exchange}): {ms:.2f} ms -> {half / ms / 1e6:.0f} GB/s per direction per GPU "
          f"(770 GB/s measured peer-copy reference)", flush=True)
sim.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("sharded check ok", flush=True)
