"""Build check of the run-time specialised pass kernels without a GPU: the generator's CUDA C++ must compile with NVRTC
for sm_100a for passes of every shape (all op kinds and target homes, controls everywhere, partial tiles, folded flips,
fused diagonal runs), and the result must be a cubin whose SASS holds the TMA tensor instructions."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import cuda_quantum_simulator_b200 as q
import helpers as H


def _nvrtc_present():
    return any(os.path.exists(p) for p in ("/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"))


pytestmark = pytest.mark.skipif(not _nvrtc_present(), reason="NVRTC not installed")


def test_generated_source_is_structural():
    """Matrix entries are run-time data: two circuits that differ only in their angles generate the same kernel."""
    a = q.Circuit(14).h(0).rz(3, 0.3).cnot(3, 9).ry(12, 1.0).crz(2, 13, 0.7).h(9)
    b = q.Circuit(14).h(0).rz(3, 1.3).cnot(3, 9).ry(12, 0.2).crz(2, 13, 0.1).h(9)
    pa, pb = q.CompiledCircuit(a), q.CompiledCircuit(b)
    assert pa.n_passes == pb.n_passes == 1
    assert pa.jit_source(0) == pb.jit_source(0)
    assert "sops[" in pa.jit_source(0) and "__shfl_xor_sync" in pa.jit_source(0)


@pytest.mark.parametrize("n,depth,seed", [(3, 30, 0), (7, 60, 1), (12, 80, 2), (16, 120, 3), (30, 60, 4)])
def test_every_pass_compiles_for_sm100a(n, depth, seed):
    g = H.random_gates(n, depth, np.random.default_rng(seed))
    prog = q.CompiledCircuit(q.Circuit(n).extend(g))
    for i in range(prog.n_passes):
        assert prog.jit_compile(i) > 10000


def test_c2_c3_kernels_and_their_sass(tmp_path):
    c2 = q.CompiledCircuit(q.create_random_circuit(30, 20, 42))
    assert c2.n_passes == 1
    cub = c2.jit_compile(0, want_cubin=True)
    c3 = q.CompiledCircuit(H.qft_style_circuit(26))
    assert "PHASE" in c3.describe()
    for i in range(c3.n_passes):
        assert c3.jit_compile(i) > 10000
    if shutil.which("cuobjdump"):
        path = tmp_path / "c2.cubin"
        path.write_bytes(cub)
        sass = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True).stdout
        assert "UTMALDG" in sass and "UTMASTG" in sass and "DFMA" in sass
        res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(path)], capture_output=True, text=True).stdout
        assert "STACK:0" in res and "LOCAL:0" in res, res   # the 9-op C2 pass fits the register file: no spills


def test_background_compilation_machinery(monkeypatch):
    """The default mode queues compiles on background threads (a run never waits for NVRTC).  Without a GPU: queue every pass
    of a circuit, wait, find them ready."""
    g = H.random_gates(15, 120, np.random.default_rng(99))
    g["param"] = np.where(g["param"] != 0, g["param"] + 0.5, 0.0)
    prog = q.CompiledCircuit(q.Circuit(15).extend(g))
    states = [prog.jit_request(i) for i in range(prog.n_passes)]
    assert set(states) <= {"ready", "compiling"}
    q.jit_wait()
    assert [prog.jit_request(i) for i in range(prog.n_passes)] == ["ready"] * prog.n_passes
    q.jit_wait()      # nothing queued: returns at once
    assert q.jit_stats()["failures"] == 0


def test_two_warp_group_build_compiles(tmp_path):
    """The two-warp-group build (QSIM_DUAL_GROUPS: 8 + 8 warps on two tiles at once) of heavy 30-qubit passes: generated,
    compiled for sm_100a, named barriers of 256 threads in the SASS; the policy picks it for the heavy passes only."""
    prog = q.CompiledCircuit(q.create_random_circuit(30, 200, 42))
    c2 = q.CompiledCircuit(q.create_random_circuit(30, 20, 42))
    assert "two warp groups" not in c2.jit_source(0)          # 9 ops: HBM-bound, one group
    assert "two warp groups" in prog.jit_source(0)             # 21 executed ops: compute-bound
    q.jit_set_dual("always")
    try:
        src = prog.jit_source(0)
        assert "QSIM_DUAL_GROUPS" in src and "bar.sync %0, 256" in src and "tid0 + 256u" in src
        cub = prog.jit_compile(0, want_cubin=True)
        assert len(cub) > 10000
        if shutil.which("cuobjdump"):
            path = tmp_path / "dual.cubin"
            path.write_bytes(cub)
            sass = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True).stdout
            assert "UTMALDG" in sass and "UTMASTG" in sass and "BAR.SYNC" in sass
        q.jit_set_dual("off")
        assert "QSIM_DUAL_GROUPS" not in prog.jit_source(0)
    finally:
        q.jit_set_dual("auto")


def test_conditional_flips_are_carried_to_the_store_in_generated_code():
    """A CNOT-heavy circuit: controlled flips on register bits become `fx_ ^= ...` (an XOR mask over the thread's slot index that
    the sweep's store addressing applies) instead of predicated register moves; later ops on the flipped bit select the
    X-conjugated matrix; the generated kernels still compile for sm_100a."""
    rng = np.random.default_rng(8103)
    g = H.random_gates(20, 160, rng, kinds=[0, 3, 3, 8, 9, 11, 11, 11, 11, 12, 13, 14, 15, 16, 5, 6])
    prog = q.CompiledCircuit(q.Circuit(20).extend(g))
    srcs = [prog.jit_source(i) for i in range(prog.n_passes)]
    assert any("carried as a slot-index XOR" in s and "fx_ ^=" in s for s in srcs)
    assert any("sw_ ? m3_ : m0_" in s for s in srcs)                       # a 2x2 on a bit with a pending flip
    assert any("if ((fx_ >>" in s and "sb_ ^=" in s for s in srcs)          # applied by the store addressing
    i = next(k for k, s in enumerate(srcs) if "fx_ ^=" in s)
    assert prog.jit_compile(i) > 10000
