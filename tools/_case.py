import math, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_quantum_simulator_b200 as q
from cuda_quantum_simulator_b200.sharded import ShardedSimulator
import helpers as H
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ng = int(math.log2(world))
rng = np.random.default_rng(99)
case = 0
while True:
    n = int(rng.integers(14 + ng, 21 + ng)); d = int(rng.integers(20, 250))
    kinds = None if rng.random() < 0.5 else [0, 3, 3, 8, 9, 10, 11, 11, 12, 15, 16, 5]
    g1, g2 = H.random_gates(n, d, rng, kinds=kinds), H.random_gates(n, int(rng.integers(5, 80)), rng, kinds=kinds)
    pr = bool(rng.random() < 0.5)
    if (n, d) == (20, 237): break
    case += 1
sim = ShardedSimulator(n)
sim._pristine = pr
want = H.zero_state(n)
for gi, g in enumerate((g1, g2)):
    cp = sim.compile(q.Circuit(n).extend(g))
    if os.environ.get("STEPWISE"):
        # step by step, checking the stored (physical) state against the oracle after every step
        eng = sim.engine
        steps, progs = cp.plan.steps, cp.programs
        phys = H.zero_state(n) if gi == 0 else None
        i = 0
        while i < len(steps) and gi == 0:
            st = steps[i]
            fusedstep = False
            if st.kind == "gates":
                nxt = steps[i + 1] if i + 1 < len(steps) else None
                phys = H.oracle_run(n, st.gates, phys)
                if nxt is not None and nxt.kind == "swap" and eng.run_program_then_swap(progs[i], nxt.global_qubit, nxt.local_qubit):
                    fusedstep = True
                    a, b = nxt.global_qubit, nxt.local_qubit
                    idx = np.arange(1 << n, dtype=np.uint64)
                    ba, bb = (idx >> np.uint64(a)) & np.uint64(1), (idx >> np.uint64(b)) & np.uint64(1)
                    src = (idx & ~((np.uint64(1) << np.uint64(a)) | (np.uint64(1) << np.uint64(b)))) | (bb << np.uint64(a)) | (ba << np.uint64(b))
                    phys = phys[src.astype(np.int64)]
                    i += 1
                else:
                    eng.run_program(progs[i])
            else:
                eng.swap(st.global_qubit, st.local_qubit)
                a, b = st.global_qubit, st.local_qubit
                idx = np.arange(1 << n, dtype=np.uint64)
                ba, bb = (idx >> np.uint64(a)) & np.uint64(1), (idx >> np.uint64(b)) & np.uint64(1)
                src = (idx & ~((np.uint64(1) << np.uint64(a)) | (np.uint64(1) << np.uint64(b)))) | (bb << np.uint64(a)) | (ba << np.uint64(b))
                phys = phys[src.astype(np.int64)]
            i += 1
            local = eng.local_state()
            t = torch.from_numpy(local).cuda(); outs = [torch.empty_like(t) for _ in range(world)]; dist.all_gather(outs, t)
            stored = np.concatenate([o.cpu().numpy() for o in outs])
            # the stored state may carry an X frame on global bits between programs: compare up to that relabelling
            errs = [float(np.max(np.abs(stored.reshape(world, -1)[[r ^ fx for r in range(world)]].reshape(-1) - phys))) for fx in range(world)]
            if rank == 0:
                print("   step", i, st.kind, "fused" if fusedstep else "", "err(min over rank relabel)", min(errs), flush=True)
        break
    if os.environ.get("SYNC_AFTER") is not None and gi == 0:
        sync_after = {int(x) for x in os.environ["SYNC_AFTER"].split(",") if x}
        eng = sim.engine
        steps, progs = cp.plan.steps, cp.programs
        i = 0
        while i < len(steps):
            st = steps[i]
            if st.kind == "gates":
                nxt = steps[i + 1] if i + 1 < len(steps) else None
                if nxt is not None and nxt.kind == "swap" and eng.run_program_then_swap(progs[i], nxt.global_qubit, nxt.local_qubit):
                    i += 1
                else:
                    eng.run_program(progs[i])
            else:
                eng.swap(st.global_qubit, st.local_qubit)
            if i in sync_after:
                torch.cuda.synchronize(); dist.barrier()
            i += 1
        sim.perm = list(cp.plan.perm); sim.frame = cp.frame_after; sim._pristine = False
    else:
        sim.execute(cp)
    got = sim.get_state_vector()
    want = H.oracle_run(n, g, want)
    if rank == 0:
        print(os.environ.get("TAG", ""), "run", gi, "swaps", cp.n_swaps, "fused", sim.engine.fused_exchanges, "err", float(np.max(np.abs(got - want))), flush=True)
        for st in cp.plan.steps:
            print("   ", st.kind, len(st.gates) if st.kind == "gates" else (st.global_qubit, st.local_qubit), flush=True)
sim.close()
dist.destroy_process_group()
