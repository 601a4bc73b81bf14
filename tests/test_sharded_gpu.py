"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box).  Spawns torchrun on tools/sharded_check.py,
which compares the sharded simulators - the Python driver and the C++ driver qsim::ShardedSimulator ("native"), each with CUDA
IPC peer memory and with the NCCL fallback - with the CPU oracle."""
import os
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["p2p", "nccl", "native", "native-nccl", "native-inplace"])
def test_two_gpu_parity(exchange):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(H.ROOT, "tools", "sharded_check.py"), exchange, "24"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded check ok" in out.stdout


@pytest.mark.parametrize("nl,g_local,tail_x", [(14, 13, False), (16, 3, False), (20, 17, True), (21, 0, False), (20, 18, "victim"),
                                               (22, 19, "victim")])
def test_fused_exchange_on_one_gpu(nl, g_local, tail_x):
    """qsim_shard_execute_exchange with both 'ranks' of a 2-shard state living on this one GPU: the program's last pass
    stores the staying half into the rank's second buffer and the leaving half into the other rank's second buffer.
    Checked against the oracle: gates on the full state, then the global<->local qubit swap as an index permutation."""
    import ctypes
    from ctypes import byref, c_uint64, c_void_p

    import numpy as np
    import torch

    from cuda_quantum_simulator_b200 import _lib

    L = _lib.lib()
    n = nl + 1
    rng = np.random.default_rng(900 + nl)
    # gates on local qubits only, none on the victim at the end (so it is not a tile qubit of the last pass)
    others = [q for q in range(nl) if q != g_local]
    lst = []
    for _ in range(30):
        k = str(rng.choice(["H", "T", "CNOT", "Rz", "X"]))
        a, b = (int(x) for x in rng.choice(others, 2, replace=False))
        lst.append((k, a, b) if k == "CNOT" else (k, a, float(rng.uniform(-3, 3))) if k == "Rz" else (k, a))
    if tail_x == "victim":
        lst.append(("X", g_local))      # a deferred X on the exchanged qubit itself: partner tiles differ in that very bit
        lst.append(("X", others[-2]))
    elif tail_x:
        lst.append(("X", others[-1]))   # a deferred X on a high local qubit: the partner-tile store path
    g = H.gates(lst)
    full = H.random_state(n, rng)
    want_gates = H.oracle_run(n, g, full)
    idx = np.arange(1 << n, dtype=np.uint64)
    bg, bl = (idx >> np.uint64(nl)) & np.uint64(1), (idx >> np.uint64(g_local)) & np.uint64(1)
    src = (idx & ~((np.uint64(1) << np.uint64(nl)) | (np.uint64(1) << np.uint64(g_local)))) | (bl << np.uint64(nl)) | (bg << np.uint64(g_local))
    want = want_gates[src.astype(np.int64)]

    bufs = [[torch.empty(1 << nl, dtype=torch.complex128, device="cuda") for _ in range(2)] for _ in range(2)]
    sims = []
    for r in range(2):
        h = c_void_p()
        _lib.check(L.qsim_shard_create(n, 1, r, c_void_p(bufs[r][0].data_ptr()), byref(h)))
        shard = np.ascontiguousarray(full[r << nl:(r + 1) << nl])
        _lib.check(L.qsim_sim_set_state(h, shard.ctypes.data_as(c_void_p)))
        sims.append(h)
    prog = c_void_p()
    _lib.check(L.qsim_program_compile_ex(n, 1, _lib.gates_ptr(g), len(g), c_uint64(0), byref(prog)))
    mask = c_uint64()
    _lib.check(L.qsim_program_last_tile_mask(prog, byref(mask)))
    if (mask.value >> g_local) & 1:
        # the victim is a tile qubit of the last pass: the call must refuse and leave everything untouched
        rc = L.qsim_shard_execute_exchange(sims[0], prog, c_void_p(bufs[0][1].data_ptr()), c_void_p(bufs[1][1].data_ptr()), nl, g_local)
        assert rc != 0
    else:
        for r in range(2):
            _lib.check(L.qsim_shard_execute_exchange(sims[r], prog, c_void_p(bufs[r][1].data_ptr()),
                                                     c_void_p(bufs[1 - r][1].data_ptr()), nl, g_local))
        torch.cuda.synchronize()
        got = np.concatenate([bufs[r][1].cpu().numpy() for r in range(2)])
        assert np.max(np.abs(got - want)) < 1e-12
        # the simulators continue on their second buffers
        out = np.empty(1 << nl, np.complex128)
        _lib.check(L.qsim_sim_get_state(sims[1], out.ctypes.data_as(c_void_p)))
        assert np.array_equal(out, got[1 << nl:])
    L.qsim_program_destroy(prog)
    for h in sims:
        L.qsim_sim_destroy(h)


@pytest.mark.parametrize("nl,g_local,tail_x", [(18, 17, False), (18, 3, False), (18, 14, False), (18, 15, True), (19, 18, True), (18, 16, "victim"),
                                               (15, 14, False)])
def test_inplace_fused_exchange_on_one_gpu(nl, g_local, tail_x):
    """qsim_shard_execute_exchange_inplace with both 'ranks' of a 2-shard state on this one GPU, their kernels running
    concurrently on two streams: no second buffer - the staying half is stored in place, the leaving half straight over the
    other rank's leaving half, ordered tile by tile by the kernels' handshake words.  Two epochs back to back (the words are
    never reset).  Oracle: gates on the full state, then the qubit swap as an index permutation."""
    from ctypes import byref, c_int, c_uint64, c_void_p

    import numpy as np
    import torch

    from cuda_quantum_simulator_b200 import _lib

    L = _lib.lib()
    n = nl + 1
    rng = np.random.default_rng(1900 + nl + g_local)
    others = [q for q in range(nl) if q != g_local]
    lst = []
    for _ in range(30):
        k = str(rng.choice(["H", "T", "CNOT", "Rz", "X"]))
        a, b = (int(x) for x in rng.choice(others, 2, replace=False))
        lst.append((k, a, b) if k == "CNOT" else (k, a, float(rng.uniform(-3, 3))) if k == "Rz" else (k, a))
    if tail_x == "victim":
        lst.append(("X", g_local))
        lst.append(("X", others[-2]))
    elif tail_x:
        lst.append(("X", others[-1]))   # a deferred X on a high local qubit: tiles are stored to their pair partner's place
    g = H.gates(lst)
    full = H.random_state(n, rng)
    idx = np.arange(1 << n, dtype=np.uint64)
    bg, bl = (idx >> np.uint64(nl)) & np.uint64(1), (idx >> np.uint64(g_local)) & np.uint64(1)
    src = ((idx & ~((np.uint64(1) << np.uint64(nl)) | (np.uint64(1) << np.uint64(g_local)))) | (bl << np.uint64(nl)) | (bg << np.uint64(g_local))).astype(np.int64)

    bufs = [torch.empty(1 << nl, dtype=torch.complex128, device="cuda") for _ in range(2)]
    hs = [torch.zeros(1024 + 8, dtype=torch.int64, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    sims = []
    for r in range(2):
        h = c_void_p()
        _lib.check(L.qsim_shard_create(n, 1, r, c_void_p(bufs[r].data_ptr()), byref(h)))
        shard = np.ascontiguousarray(full[r << nl:(r + 1) << nl])
        _lib.check(L.qsim_sim_set_state(h, shard.ctypes.data_as(c_void_p)))
        _lib.check(L.qsim_sim_set_stream(h, c_void_p(streams[r].cuda_stream)))
        sims.append(h)
    torch.cuda.synchronize()
    prog = c_void_p()
    _lib.check(L.qsim_program_compile_ex(n, 1, _lib.gates_ptr(g), len(g), c_uint64(0), byref(prog)))
    ok = c_int()
    _lib.check(L.qsim_shard_inplace_exchange_possible(sims[0], prog, g_local, byref(ok)))

    def call(r, epoch):
        return L.qsim_shard_execute_exchange_inplace(sims[r], prog, c_void_p(bufs[1 - r].data_ptr()), nl, g_local,
                                                     c_void_p(hs[r].data_ptr()), c_void_p(hs[1 - r].data_ptr()), c_uint64(epoch << 32),
                                                     c_uint64(3_000_000_000), c_void_p(hs[r].data_ptr() + 1024 * 8))

    mask = c_uint64()
    _lib.check(L.qsim_program_last_tile_mask(prog, byref(mask)))
    if tail_x == "victim" or nl < 16 or ((mask.value >> g_local) & 1):
        # a deferred X pairs tiles across the exchanged qubit / fewer than 16 tiles / the qubit is a tile qubit of the last
        # pass: refused, nothing touched
        assert ok.value == 0
        assert call(0, 1) != 0
    else:
        assert ok.value == 1
        want = full
        for epoch in (1, 2):
            want = H.oracle_run(n, g, want)[src]
            for r in range(2):
                _lib.check(call(r, epoch))
            torch.cuda.synchronize()
            assert int(hs[0][1024].item()) == 0 and int(hs[1][1024].item()) == 0, "handshake timed out"
            got = np.concatenate([bufs[r].cpu().numpy() for r in range(2)])
            assert np.max(np.abs(got - want)) < 1e-12
    L.qsim_program_destroy(prog)
    for h in sims:
        L.qsim_sim_destroy(h)


@pytest.mark.parametrize("nl,g_local,tail_x", [(18, 17, False), (19, 16, False), (18, 14, True), (20, 19, False)])
def test_split_exchange_on_one_gpu(nl, g_local, tail_x):
    """qsim_shard_execute_exchange_half with both 'ranks' of a 2-shard state on this one GPU (kernels concurrent on two streams):
    the LAST pass of the program before the exchange scatters the half of the leaving tiles whose split bit is clear into the
    other rank's shard, the FIRST pass of the program after it - which has the exchanged qubit as its highest tile qubit - loads
    the leaving half of every tile whose split bit is set from there (TMA boxes from two GPUs into one tile) - no second buffer,
    no separate swap.  Oracle: gates A on the full state, the qubit swap as an index permutation, gates B."""
    from ctypes import byref, c_int, c_uint64, c_void_p

    import numpy as np
    import torch

    from cuda_quantum_simulator_b200 import _lib

    L = _lib.lib()
    n = nl + 1
    rng = np.random.default_rng(2900 + nl + g_local)
    # few distinct targets per program, so that enough qubits stay outside both tiles to split on
    pool = [q for q in range(nl) if q != g_local]
    targets = [int(x) for x in rng.choice(pool[:12], 7, replace=False)]

    def program_gates(k):
        lst = []
        for _ in range(k):
            kind = str(rng.choice(["H", "T", "CNOT", "Rz", "X"]))
            a, b = (int(x) for x in rng.choice(targets, 2, replace=False))
            lst.append((kind, a, b) if kind == "CNOT" else (kind, a, float(rng.uniform(-3, 3))) if kind == "Rz" else (kind, a))
        return lst
    la, lb = program_gates(24), program_gates(16)
    # the program after the exchange works on the exchanged qubit (that is why it was exchanged): it becomes the highest tile
    # qubit of its first pass, which the compile hint makes TMA instructions of its own enumerate
    lb = [("H", g_local)] + lb[:8] + [("CNOT", targets[1], g_local), ("Rz", g_local, 0.37)] + lb[8:] + [("H", g_local)]
    if tail_x:
        la.append(("X", targets[0]))
        lb.append(("X", targets[2]))
    ga, gb = H.gates(la), H.gates(lb)
    full = H.random_state(n, rng)
    idx = np.arange(1 << n, dtype=np.uint64)
    bg, bl = (idx >> np.uint64(nl)) & np.uint64(1), (idx >> np.uint64(g_local)) & np.uint64(1)
    src = ((idx & ~((np.uint64(1) << np.uint64(nl)) | (np.uint64(1) << np.uint64(g_local)))) | (bl << np.uint64(nl)) | (bg << np.uint64(g_local))).astype(np.int64)
    want = H.oracle_run(n, gb, H.oracle_run(n, ga, full)[src])

    bufs = [torch.empty(1 << nl, dtype=torch.complex128, device="cuda") for _ in range(2)]
    hs = [torch.zeros(1024 + 8, dtype=torch.int64, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    sims = []
    for r in range(2):
        h = c_void_p()
        _lib.check(L.qsim_shard_create(n, 1, r, c_void_p(bufs[r].data_ptr()), byref(h)))
        shard = np.ascontiguousarray(full[r << nl:(r + 1) << nl])
        _lib.check(L.qsim_sim_set_state(h, shard.ctypes.data_as(c_void_p)))
        _lib.check(L.qsim_sim_set_stream(h, c_void_p(streams[r].cuda_stream)))
        sims.append(h)
    torch.cuda.synchronize()
    pa, pb = c_void_p(), c_void_p()
    _lib.check(L.qsim_program_compile_ex(n, 1, _lib.gates_ptr(ga), len(ga), c_uint64(0), byref(pa)))
    _lib.check(L.qsim_program_compile_ex2(n, 1, _lib.gates_ptr(gb), len(gb), c_uint64(0), g_local, byref(pb)))
    w = c_int(-2)
    _lib.check(L.qsim_shard_split_exchange_possible(sims[0], pa, pb, g_local, byref(w)))
    assert w.value >= 0 and w.value != g_local, "no split bit for this case"

    def half(r, prog, which, epoch):
        return L.qsim_shard_execute_exchange_half(sims[r], prog, c_void_p(bufs[1 - r].data_ptr()), nl, g_local, w.value, which,
                                                  c_void_p(hs[r].data_ptr()), c_void_p(hs[1 - r].data_ptr()), c_uint64(epoch << 32),
                                                  c_uint64(3_000_000_000), c_void_p(hs[r].data_ptr() + 1024 * 8))
    for which, prog, epoch in ((1, pa, 1), (2, pb, 2)):
        for r in range(2):
            _lib.check(half(r, prog, which, epoch))
        torch.cuda.synchronize()          # (the ranks' barrier between the two halves)
        assert int(hs[0][1024].item()) == 0 and int(hs[1][1024].item()) == 0, "handshake failed"
    got = np.concatenate([bufs[r].cpu().numpy() for r in range(2)])
    assert np.max(np.abs(got - want)) < 1e-12
    # a split bit that is a tile qubit / the exchanged qubit itself is refused
    assert L.qsim_shard_execute_exchange_half(sims[0], pa, c_void_p(bufs[1].data_ptr()), nl, g_local, g_local, 1, c_void_p(hs[0].data_ptr()),
                                              c_void_p(hs[1].data_ptr()), c_uint64(9 << 32), c_uint64(0), c_void_p(hs[0].data_ptr() + 8192)) != 0
    for p_ in (pa, pb):
        L.qsim_program_destroy(p_)
    for h in sims:
        L.qsim_sim_destroy(h)


def test_staged_shard_sampling_on_one_gpu():
    """qsim_shard_cdf_prepare / _classify / qsim_shard_sample with both shards of a 2-shard state on this GPU: the chained
    result must equal the host's sequential CDF over the whole state, bit for bit."""
    from ctypes import byref, c_double, c_void_p

    import numpy as np

    from cuda_quantum_simulator_b200 import _lib

    L = _lib.lib()
    nl, n = 15, 16
    rng = np.random.default_rng(321)
    full = H.random_state(n, rng)
    full[: 1 << 13] *= 1e-9          # a tiny head, so the second shard starts far from where its own sum would
    full /= np.linalg.norm(full)
    probs = H.oracle_probs(full)
    u = np.concatenate([rng.random(2000), [0.0, 0.5]])
    want = H.oracle_sample(probs, u)
    sims = []
    for r in range(2):
        h = c_void_p()
        _lib.check(L.qsim_shard_create(n, 1, r, None, byref(h)))
        shard = np.ascontiguousarray(full[r << nl:(r + 1) << nl])
        _lib.check(L.qsim_sim_set_state(h, shard.ctypes.data_as(c_void_p)))
        sims.append(h)
    approx = []
    for h in sims:
        v = c_double()
        _lib.check(L.qsim_shard_cdf_prepare(h, byref(v)))
        approx.append(v.value)
    assert abs(sum(approx) - 1.0) < 1e-9
    _lib.check(L.qsim_shard_cdf_classify(sims[0], c_double(0.0)))
    _lib.check(L.qsim_shard_cdf_classify(sims[1], c_double(approx[0])))
    got = np.full(len(u), -1, np.int64)
    c = 0.0
    for r, h in enumerate(sims):
        out = np.empty(len(u), np.int64)
        c_end = c_double()
        _lib.check(L.qsim_shard_sample(h, c_double(c), int(r == 0), u.ctypes.data_as(c_void_p), len(u),
                                       out.ctypes.data_as(c_void_p), byref(c_end)))
        hit = out >= 0
        got[hit] = (r << nl) | out[hit]
        c = c_end.value
    assert np.array_equal(got, want)
    assert c == H.oracle().orc_total_probability(probs.ctypes.data_as(H.P), H.c_int64(len(probs)))
    for h in sims:
        L.qsim_sim_destroy(h)


def test_shard_run_rejects_x_on_a_rank_qubit():
    """qsim_sim_run / qsim_sim_apply_gate on a shard keep no X frame between calls: an uncontrolled X on a rank (global)
    qubit must fail loudly ("must be remapped") instead of being absorbed into a frame nobody applies; the same gate on
    a local qubit works, and a diagonal gate or a control on the rank qubit needs no remap."""
    from ctypes import byref, c_void_p

    import numpy as np

    import cuda_quantum_simulator_b200 as q
    from cuda_quantum_simulator_b200 import _lib

    L = _lib.lib()
    nl, n = 10, 11
    rng = np.random.default_rng(5)
    full = H.random_state(n, rng)
    sims = []
    for r in range(2):
        h = c_void_p()
        _lib.check(L.qsim_shard_create(n, 1, r, None, byref(h)))
        _lib.check(L.qsim_sim_set_state(h, np.ascontiguousarray(full[r << nl:(r + 1) << nl]).ctypes.data_as(c_void_p)))
        sims.append(h)
    bad = H.gates([("X", n - 1)])
    ok = H.gates([("X", 3), ("CNOT", n - 1, 2), ("Rz", n - 1, 0.3), ("H", 0), ("X", 0)])
    for h in sims:
        with pytest.raises(q.QsimError, match="remapped"):
            _lib.check(L.qsim_sim_run(h, n, _lib.gates_ptr(bad), len(bad)))
        with pytest.raises(q.QsimError, match="remapped"):
            _lib.check(L.qsim_sim_apply_gate(h, _lib.gates_ptr(bad)))
        _lib.check(L.qsim_sim_run(h, n, _lib.gates_ptr(ok), len(ok)))
        _lib.check(L.qsim_sim_apply_gate(h, _lib.gates_ptr(H.gates([("X", 5)]))))
    want = H.oracle_run(n, np.concatenate([ok, H.gates([("X", 5)])]), full)
    got = []
    for h in sims:
        out = np.empty(1 << nl, np.complex128)
        _lib.check(L.qsim_sim_get_state(h, out.ctypes.data_as(c_void_p)))
        got.append(out)
        L.qsim_sim_destroy(h)
    assert np.max(np.abs(np.concatenate(got) - want)) < 1e-12


def test_native_sharded_simulator_on_one_gpu():
    """qsim::ShardedSimulator (the C++ driver) with a world of one rank: no exchange can happen, but planning, compiling,
    plan validity, read-out (sampling in the reference's order, measurement on index bit n-1-q, marginals) all run through
    the same C++ code the multi-GPU runs use, here against the oracle."""
    import numpy as np

    import cuda_quantum_simulator_b200 as q
    from cuda_quantum_simulator_b200.sharded import NativeShardedSimulator

    n = 13
    rng = np.random.default_rng(31)
    g = H.random_gates(n, 120, rng)
    c = q.Circuit(n).extend(g)
    sim = NativeShardedSimulator(n)
    assert (sim.world, sim.nl, sim.ng, sim.exchange) == (1, n, 0, "none")
    sim.run(c)
    want = H.oracle_run(n, g)
    assert np.max(np.abs(sim.get_state_vector() - want)) < 1e-12
    plans = sim.compile_sequence(c, 2)
    for p_ in plans:
        sim.execute(p_)
    want = H.oracle_run(n, g, H.oracle_run(n, g, want))
    got = sim.get_state_vector()
    assert np.max(np.abs(got - want)) < 1e-11
    assert plans[0].n_swaps == 0 and plans[0].n_passes >= 1
    u = np.concatenate([rng.random(300), [0.0, 0.5]])
    assert np.array_equal(sim.sample(uniforms=u), H.oracle_sample(H.oracle_probs(got), u))
    qs = [0, 7, n - 1]
    idx = np.arange(1 << n)
    oc = ((idx >> 0) & 1) | (((idx >> 7) & 1) << 1) | (((idx >> (n - 1)) & 1) << 2)
    assert np.max(np.abs(sim.marginal(qs) - np.bincount(oc, weights=np.abs(got) ** 2, minlength=8))) < 1e-12
    assert abs(sim.get_total_probability() - 1) < 1e-10
    bit = n - 1 - 2
    p0 = float(np.sum((np.abs(got) ** 2)[((idx >> bit) & 1) == 0]))
    res = sim.measure_qubit(2, 0.5)
    assert res == (0 if 0.5 < p0 else 1)
    keep = ((idx >> bit) & 1) == res
    want_c = np.where(keep, got, 0) / np.sqrt(p0 if res == 0 else 1 - p0)
    assert np.max(np.abs(sim.get_state_vector() - want_c)) < 1e-11
    with pytest.raises(q.InvalidArgument):
        sim.run(q.Circuit(n - 1).h(0))
    sim.close()
