"""Multi-rank path on CPU (gloo, world_size 2 and 4): the planner, the X-frame / permutation bookkeeping
and the exchange choreography of cuda_quantum_simulator_b200.sharded, with a TEST engine that executes
local segments through the kernel emulator and exchanges half shards over gloo.  The product engine
(CudaShardEngine) is exercised on GPUs by tests/test_sharded_gpu.py."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from cuda_quantum_simulator_b200 import Circuit
from cuda_quantum_simulator_b200.sharded import ShardedSimulator, plan_circuit


class EmulatorShardEngine:
    """TEST ONLY: local segments run in oracle/program_emulator (CPU), exchanges go over gloo."""

    def __init__(self, n, ng, rank, world):
        self.n, self.ng, self.nl, self.rank, self.world = n, ng, n - ng, rank, world
        self.shard = np.zeros(1 << self.nl, np.complex128)
        self.reset()

    def reset(self):
        self.shard[:] = 0
        if self.rank == 0:
            self.shard[0] = 1

    def _emu(self, gates, initial_xor, state):
        info = np.zeros(8, np.int64)
        err = ctypes.create_string_buffer(256)
        g = np.ascontiguousarray(gates, H.GATE_DTYPE)
        rc = H.emulator().emu_run_ex(self.n, self.ng, self.rank, g.ctypes.data_as(H.P) if len(g) else None,
                                     H.c_int64(len(g)), state.ctypes.data_as(H.P), 3, 6, 1, 1,
                                     ctypes.c_uint64(initial_xor), info.ctypes.data_as(H.P), err, 256)
        assert rc == 0, err.value
        return info

    def compile_gates(self, gates, initial_xor):
        info = self._emu(gates, initial_xor, self.shard.copy())       # dry run for the bookkeeping outputs
        return (np.array(gates), initial_xor), {"passes": int(info[0]), "ops": int(info[1]), "global_xor": int(info[3])}

    def run_program(self, handle):
        self._emu(handle[0], handle[1], self.shard)

    def free_program(self, handle):
        pass

    def swap(self, g, l):
        peer = self.rank ^ (1 << (g - self.nl))
        my_bit = (self.rank >> (g - self.nl)) & 1
        idx = np.arange(1 << self.nl)
        leaving = ((idx >> l) & 1) != my_bit
        send = torch.from_numpy(self.shard[leaving].copy().view(np.float64))
        recv = torch.empty_like(send)
        reqs = [dist.isend(send, peer), dist.irecv(recv, peer)]
        for r in reqs:
            r.wait()
        self.shard[leaving] = recv.numpy().view(np.complex128)

    fused_exchanges = 0

    def run_program_then_swap(self, handle, g, l):
        """Stands in for the fused pass + exchange of the CUDA engine (same contract: refuse, having done nothing, or do
        both): accepted for every other request so that both branches of ShardedSimulator.execute are exercised."""
        self._asked = getattr(self, "_asked", 0) + 1
        if self._asked % 2 == 0:
            return False
        self.run_program(handle)
        self.swap(g, l)
        self.fused_exchanges += 1
        return True

    def marginal(self, local_bits):
        probs = H.oracle_probs(self.shard)
        idx = np.arange(len(probs))
        outcome = np.zeros(len(probs), np.int64)
        for i, b in enumerate(local_bits):
            outcome |= ((idx >> b) & 1) << i
        return np.bincount(outcome, weights=probs, minlength=1 << len(local_bits))

    def cdf_prepare(self):
        """Staged sampling of the CUDA engine: this shard's approximate total (here simply its exact total)."""
        return float(np.sum(H.oracle_probs(self.shard)))

    def cdf_classify(self, approx_c_init):
        self._approx_c_init = approx_c_init   # only used by the CUDA engine to pick tentative binades

    def synchronize(self): pass
    def local_state(self): return self.shard.copy()
    def partial_probability(self, bit=-1):
        pr = np.abs(self.shard) ** 2
        if bit >= 0:
            pr = pr[((np.arange(len(pr)) >> bit) & 1) == 0]
        return float(np.sum(pr))

    def collapse(self, bit, outcome, scale):
        if bit >= 0:
            self.shard[((np.arange(len(self.shard)) >> bit) & 1) != outcome] = 0
        self.shard *= scale

    def shard_sample(self, c_init, first, u):
        probs = H.oracle_probs(self.shard)
        cum = np.empty(len(probs))
        c = c_init
        for i, p in enumerate(probs):         # sequential fp64, continuing from c_init
            c = c + p
            cum[i] = c
        out = np.searchsorted(cum, u, side="left").astype(np.int64)
        mine = (cum[-1] >= u) & ((c_init < u) | first)
        out[~mine] = -1
        return out, float(cum[-1])

    def allgather_float(self, v):
        t = torch.tensor([v], dtype=torch.float64)
        outs = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t)
        return [float(x.item()) for x in outs]

    def allreduce_max(self, arr):
        t = torch.from_numpy(arr.copy())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.numpy()

    def allgather_array(self, a):
        t = torch.from_numpy(a.view(np.float64).copy())
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t)
        return [o.numpy().view(np.complex128) for o in outs]

    def close(self): pass


def _worker(rank, world, port, n, seeds, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ng = world.bit_length() - 1
    try:
        worst = 0.0
        for seed in seeds:
            rng = np.random.default_rng(seed)
            g = H.random_gates(n, 60, rng)
            sim = ShardedSimulator(n, engine=EmulatorShardEngine(n, ng, rank, world), rank=rank, world=world)
            c = Circuit(n).extend(g)
            sim.run(c)
            got = sim.get_state_vector()
            want = H.oracle_run(n, g)
            worst = max(worst, float(np.max(np.abs(got - want))))
            # run() composes across calls with the carried permutation / frame
            sim.run(c)
            got2 = sim.get_state_vector()
            want2 = H.oracle_run(n, g, want)
            worst = max(worst, float(np.max(np.abs(got2 - want2))))
            assert abs(sim.get_total_probability() - 1) < 1e-10
            # marginals of a few logical qubits (local and rank bits mixed, after swaps and with a pending X frame)
            qs = [int(x) for x in np.random.default_rng(seed).permutation(n)[:4]]
            idx = np.arange(1 << n)
            outcome = np.zeros(1 << n, np.int64)
            for i, qb in enumerate(qs):
                outcome |= ((idx >> qb) & 1) << i
            want_m = np.bincount(outcome, weights=np.abs(want2) ** 2, minlength=16)
            assert np.max(np.abs(sim.marginal(qs) - want_m)) < 1e-12
            # sampling after real exchanges: the reference's sequential CDF in LOGICAL index order, bit for bit
            # (src/Simulator.cu:164-185) - the identity layout is restored first; the state itself is unchanged
            u = np.concatenate([np.random.default_rng(1).random(64), [0.0, 0.5]])
            s = sim.sample(uniforms=u)
            assert np.array_equal(s, H.oracle_sample(H.oracle_probs(got2), u))
            assert sim.perm == list(range(n))
            got3 = sim.get_state_vector()
            assert np.array_equal(got3, got2)            # a pure data movement
            # plans are tied to the layout they were compiled against
            cp = sim.compile(c)
            sim.execute(cp)
            if cp.plan.perm != cp.perm_before or cp.frame_after != cp.frame_before:
                with pytest.raises(ValueError):
                    sim.execute(cp)
            sim.release(cp)
            want3 = H.oracle_run(n, g, want2)
            worst = max(worst, float(np.max(np.abs(sim.get_state_vector() - want3))))
            # measureQubit: outcome r < p0 ? 0 : 1 on index bit n-1-q (reference src/StateVector.cu:87-89, 284-313), collapse
            for qb, r in ((0, 0.3), (n - 1, 0.8), (n // 2, 0.5)):
                bit = n - 1 - qb
                st = sim.get_state_vector()
                pr = np.abs(st) ** 2
                p0 = float(np.sum(pr[((np.arange(1 << n) >> bit) & 1) == 0]))
                if min(p0, 1 - p0) < 1e-9 or abs(r - p0) < 1e-9:
                    continue
                res = sim.measure_qubit(qb, r)
                assert res == (0 if r < p0 else 1)
                keep = ((np.arange(1 << n) >> bit) & 1) == res
                want_c = np.where(keep, st, 0) / np.sqrt(p0 if res == 0 else 1 - p0)
                worst = max(worst, float(np.max(np.abs(sim.get_state_vector() - want_c))))
                assert abs(sim.get_total_probability() - 1) < 1e-12
        # A circuit that leaves enough qubits without a non-diagonal target: from |0...0> the layout parks those in the
        # rank bits, no exchange happens, and — those bits being the same for every non-zero amplitude — sampling is
        # bit-identical to the single-device sequential CDF in LOGICAL index order.
        rng = np.random.default_rng(77)
        lst = []
        busy = [q_ for q_ in range(n) if q_ not in (2, n - 2)][: n - ng]
        for _ in range(50):
            k = str(rng.choice(["H", "Ry", "CNOT", "T", "X", "CZ"]))
            a, b = (int(x) for x in rng.choice(busy, 2, replace=False))
            if k == "CNOT":
                lst.append(("CNOT", int(rng.choice([2, n - 2, a])), b))       # parked qubits may control
            elif k == "CZ":
                lst.append(("CZ", 2, b))
            elif k == "Ry":
                lst.append(("Ry", a, float(rng.uniform(-3, 3))))
            elif k == "X":
                lst.append(("X", int(rng.choice([2, n - 2, a]))))              # ... and be flipped
            else:
                lst.append((k, a))
        g = H.gates(lst)
        sim = ShardedSimulator(n, engine=EmulatorShardEngine(n, ng, rank, world), rank=rank, world=world)
        cp = sim.compile(Circuit(n).extend(g))
        if ng <= 2:
            assert cp.n_swaps == 0
        sim.execute(cp)
        want = H.oracle_run(n, g)
        worst = max(worst, float(np.max(np.abs(sim.get_state_vector() - want))))
        u = np.concatenate([np.random.default_rng(3).random(200), [0.5, 0.25, 0.999]])
        if cp.n_swaps == 0:
            assert np.array_equal(sim.sample(uniforms=u), H.oracle_sample(H.oracle_probs(want), u))
        if rank == 0:
            q.put(worst)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_matches_oracle_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 8, [1, 2, 3], q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-12


def test_initial_layout_parks_untargeted_qubits_in_the_rank_bits():
    import cuda_quantum_simulator_b200 as qs
    from cuda_quantum_simulator_b200.sharded import choose_initial_layout
    for n, ng in ((31, 1), (33, 3), (36, 3)):
        c = qs.create_random_circuit(n, 20, 42)
        perm = choose_initial_layout(n, ng, c.gates)
        assert sorted(perm) == list(range(n))
        assert plan_circuit(n, ng, c.gates, perm).n_swaps == 0          # identity layout: one swap (next test)
        kept = [q for q in range(n) if perm[q] < n - ng]
        assert [perm[q] for q in kept] == list(range(n - ng))           # the others keep their relative order
    # nothing to gain: every qubit is a target early on -> the ones targeted last go global, layout still a permutation
    c = qs.Circuit(4).h(0).h(1).h(2).h(3).h(3)
    perm = choose_initial_layout(4, 1, c.gates)
    assert sorted(perm) == [0, 1, 2, 3] and perm[3] == 3


def test_planner_c4_needs_one_swap():
    """Config C4 (36 qubits over 8 GPUs): only H(35) needs an exchange; X(35) is a frame toggle."""
    import cuda_quantum_simulator_b200 as qs
    c = qs.create_random_circuit(36, 20, 42)
    plan = plan_circuit(36, 3, c.gates)
    assert plan.n_swaps == 1
    sw = [s for s in plan.steps if s.kind == "swap"][0]
    assert sw.global_qubit == 35 and sw.local_qubit < 33
    # diagonal gates and controls on global qubits never force a swap
    c2 = qs.Circuit(6).h(0).cz(5, 0).rz(5, 0.3).cnot(5, 1).crz(4, 5, 0.2).x(5).z(4)
    assert plan_circuit(6, 2, c2.gates).n_swaps == 0
    c3 = qs.Circuit(6).h(5).h(5).cnot(0, 5).ry(4, 0.3)
    p3 = plan_circuit(6, 2, c3.gates)
    assert p3.n_swaps == 2 and sorted(p3.perm) == list(range(6))


@pytest.mark.parametrize("n,ng,depth,seed", [(8, 1, 60, 1), (9, 2, 80, 2), (10, 3, 120, 3), (12, 2, 200, 4), (36, 3, 20, 42),
                                             (33, 3, 200, 42)])
def test_cpp_planner_equals_python_planner(n, ng, depth, seed):
    """csrc/sharded_plan.cpp (the planner of qsim::ShardedSimulator) against plan_circuit / choose_initial_layout above:
    same swaps at the same places, same gate records on physical positions, same final permutation."""
    import cuda_quantum_simulator_b200 as qs
    from cuda_quantum_simulator_b200.sharded import choose_initial_layout, plan_circuit_native
    if n <= 12:
        g = H.random_gates(n, depth, np.random.default_rng(seed))
    else:
        g = qs.create_random_circuit(n, depth, seed).gates
    for choose in (False, True):
        start = choose_initial_layout(n, ng, g) if choose else list(range(n))
        want = plan_circuit(n, ng, g, start)
        got, got_start = plan_circuit_native(n, ng, g, None, choose_layout=choose)
        assert got_start == start
        assert got.perm == want.perm and len(got.steps) == len(want.steps)
        for a, b in zip(got.steps, want.steps):
            assert a.kind == b.kind
            if a.kind == "swap":
                assert (a.global_qubit, a.local_qubit) == (b.global_qubit, b.local_qubit)
            else:
                assert np.array_equal(a.gates, b.gates)
    # a start permutation handed in explicitly
    rng = np.random.default_rng(seed)
    perm = [int(x) for x in rng.permutation(n)]
    want = plan_circuit(n, ng, g, perm)
    got, _ = plan_circuit_native(n, ng, g, perm)
    assert got.perm == want.perm and [s.kind for s in got.steps] == [s.kind for s in want.steps]
