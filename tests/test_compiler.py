"""The circuit compiler (csrc/program.cpp) checked on the CPU: its output is executed by a
thread-level emulation of the fused-pass kernel (oracle/program_emulator.cpp) and compared with the
oracle, over every gate type, tile geometry and scheduling option."""
import numpy as np
import pytest

import helpers as H

CONFIGS = [(0, 0, 1, 1), (3, 6, 1, 1), (3, 5, 0, 0), (4, 9, 1, 0), (3, 12, 0, 1), (5, 10, 1, 1)]


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_all_gate_types(seed):
    rng = np.random.default_rng(seed)
    for _ in range(6):
        n, d = int(rng.integers(1, 15)), int(rng.integers(1, 140))
        g = H.random_gates(n, d, rng)
        st0 = H.random_state(n, rng)
        want = H.oracle_run(n, g, st0)
        for lmin, tmax, merge, reorder in CONFIGS:
            got, _ = H.emu_run(n, g, st0, lmin=lmin, tmax=tmax, merge=merge, reorder=reorder)
            assert np.max(np.abs(got - want)) < 1e-12, (n, d, lmin, tmax, merge, reorder)


def test_golden_circuits_through_compiler():
    z = np.load(H.GOLDEN + "/ref_cpu_states.npz")
    for name in sorted(k[:-3] for k in z.files if k.endswith("__n")):
        n, g = int(z[name + "__n"]), np.ascontiguousarray(z[name + "__gates"], H.GATE_DTYPE)
        got, _ = H.emu_run(n, g, H.zero_state(n))
        tol = 1e-10 if "deep" in name else 1e-12
        assert np.max(np.abs(got - z[name + "__state"])) < tol, name


def test_sharded_compile_matches_full_state():
    """Shards of a state whose top qubits are the rank id: diagonal gates and controls on global qubits
    need no data movement and must agree with the unsharded run."""
    rng = np.random.default_rng(11)
    n, ng = 9, 2
    nl = n - ng
    for trial in range(6):
        # only diagonal targets / controls may touch the global qubits
        lst = []
        for _ in range(60):
            t = int(rng.integers(0, 17))
            qs = rng.permutation(n)
            ang = float(rng.uniform(0, 6.28))
            diag = t in (2, 4, 5, 6, 7, 10, 12, 14)
            target_pos = 0 if t < 11 else (2 if t == 16 else 1)
            if not diag and qs[target_pos] >= nl:
                continue
            if t == 15 and (qs[0] >= nl or qs[1] >= nl):
                continue
            g = [t, int(qs[0])] + ([int(qs[1])] if t >= 11 else []) + ([int(qs[2])] if t == 16 else [])
            if t in (8, 9, 10, 13, 14):
                g.append(ang)
            lst.append(tuple(g))
        g = H.gates(lst)
        st0 = H.random_state(n, rng)
        want = H.oracle_run(n, g, st0)
        got = np.empty_like(st0)
        for rank in range(1 << ng):
            shard = st0[rank << nl:(rank + 1) << nl]
            out, _ = H.emu_run(n, g, shard, n_global=ng, rank=rank, tmax=6, lmin=3)
            got[rank << nl:(rank + 1) << nl] = out
        assert np.max(np.abs(got - want)) < 1e-12


def test_plans_for_benchmark_circuits():
    """Fusion does its job: passes << gates on the reference's workloads."""
    if H.reference() is None:
        pytest.skip("oracle/_ref not built")
    g = H.ref_random_circuit(30, 20, 42)
    _, info = (None, None)
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    assert H.emulator().emu_describe(30, 0, g.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(20), 0, buf, 1 << 16) == 0
    text = buf.value.decode()
    passes = int(text.split(" passes")[0].split()[-1])
    assert passes <= 2, text
    g1 = H.bench_c1_gates(20)
    assert H.emulator().emu_describe(20, 0, g1.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(g1)), 0, buf, 1 << 16) == 0
    passes = int(buf.value.decode().split(" passes")[0].split()[-1])
    assert passes <= 4


@pytest.mark.parametrize("seed", range(6))
def test_trailing_flips_fold_into_the_store(seed):
    """Controlled flips at the end of a pass become an affine map of the store index (program.hpp TailDyn): controls
    inside and outside the tile, controls on zero (after the X frame), Toffolis that must stay ops."""
    rng = np.random.default_rng(500 + seed)
    for n, tmax in ((3, 0), (7, 5), (13, 0), (14, 8), (15, 12)):
        g = H.flip_heavy_gates(n, rng, body=int(rng.integers(0, 6)), tail=int(rng.integers(1, 30)))
        st0 = H.random_state(n, rng)
        want = H.oracle_run(n, g, st0)
        got, _ = H.emu_run(n, g, st0, lmin=3, tmax=tmax)
        assert np.max(np.abs(got - want)) < 1e-12, (n, tmax, seed)
        # the same gates reversed: the flips LEAD the pass and fold into the first sweep's load (inverse affine map)
        gr = np.ascontiguousarray(g[::-1])
        want = H.oracle_run(n, gr, st0)
        got, _ = H.emu_run(n, gr, st0, lmin=3, tmax=tmax)
        assert np.max(np.abs(got - want)) < 1e-12, ("leading", n, tmax, seed)


def test_isolate_hint_makes_the_top_tile_qubit_a_tma_instruction_bit():
    """qsim_program_compile_ex2's hint (the program after a qubit exchange, second half of a split exchange): a pass whose highest
    tile qubit is the hinted one moves it with TMA instructions of its own - same tile, same ops, one more instruction bit."""
    import ctypes
    from ctypes import byref, c_uint64, c_void_p

    from cuda_quantum_simulator_b200 import _lib
    L = _lib.lib()
    n, v = 19, 17
    g = H.gates([("H", v), ("H", 3), ("CNOT", 5, v), ("Rz", v, 0.3), ("H", 7)])
    lines = {}
    for iso in (-1, v, 9):
        p = c_void_p()
        _lib.check(L.qsim_program_compile_ex2(n, 1, _lib.gates_ptr(g), len(g), c_uint64(0), iso, byref(p)))
        buf = ctypes.create_string_buffer(8000)
        L.qsim_program_describe(p, buf, 8000)
        lines[iso] = [l for l in buf.value.decode().split("\n") if l.startswith("  pass")]
        L.qsim_program_destroy(p)
    assert len(lines[-1]) == 1 and "tile_bits=[0,1,2,3,4,5,6,7,8,9,10,17]" in lines[-1][0] and "tma_instrs" not in lines[-1][0]
    assert "tma_instrs=2" in lines[v][0] and lines[v][0].replace(" tma_instrs=2", "") == lines[-1][0]
    assert lines[9] == lines[-1]          # a hint that is not the highest tile qubit changes nothing
