"""Shared test helpers.  The ONLY place (with __graft_entry__.smoke and bench.py's cpu_baseline leg)
that loads anything under oracle/: the oracle is the checker, never the product."""
import ctypes
import json
import os
from ctypes import c_double, c_int, c_int64, c_uint, c_uint64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libqsim_oracle.so")
EMUL_SO = os.path.join(ROOT, "oracle", "_build", "libqsim_emul.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libqsim_ref.so")

GATE_DTYPE = np.dtype([("type", "<i4"), ("q0", "<i4"), ("q1", "<i4"), ("q2", "<i4"), ("param", "<f8")], align=True)
CHANNEL_DTYPE = np.dtype([("type", "<i4"), ("qubit", "<i4"), ("p", "<f8")], align=True)
NAMES = "X Y Z H S T Sdag Tdag Rx Ry Rz CNOT CZ CRY CRZ SWAP Toffoli".split()
G = {n: i for i, n in enumerate(NAMES)}
P = c_void_p


def _build_oracle():
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)


_cache = {}


def oracle():
    if "o" not in _cache:
        if not os.path.exists(ORACLE_SO):
            _build_oracle()
        L = ctypes.CDLL(ORACLE_SO, mode=ctypes.RTLD_LOCAL)
        L.orc_total_probability.restype = c_double
        L.orc_prob_zero.restype = c_double
        L.orc_dm_trace.restype = c_double
        L.orc_dm_purity.restype = c_double
        _cache["o"] = L
    return _cache["o"]


def emulator():
    if "e" not in _cache:
        if not os.path.exists(EMUL_SO):
            _build_oracle()
        _cache["e"] = ctypes.CDLL(EMUL_SO, mode=ctypes.RTLD_LOCAL)
    return _cache["e"]


def reference():
    """The unmodified reference built into oracle/_ref (None if it was never built)."""
    if "r" not in _cache:
        L = None
        if os.path.exists(REF_SO):
            L = ctypes.CDLL(REF_SO, mode=ctypes.RTLD_LOCAL)
            L.ref_cpu_run.restype = c_double
            L.ref_gpu_run.restype = c_double
            L.ref_gpu_batched_run.restype = c_double
            L.ref_circuit_depth.restype = c_int64
        _cache["r"] = L
    return _cache["r"]


def gates(lst):
    """[(name|type, q0[, q1[, q2]][, param])...] -> record array.  Angle goes last as a float."""
    out = np.zeros(len(lst), GATE_DTYPE)
    for i, g in enumerate(lst):
        t = G[g[0]] if isinstance(g[0], str) else int(g[0])
        rest = list(g[1:])
        param = 0.0
        if rest and isinstance(rest[-1], float):
            param = rest.pop()
        q = rest + [-1] * (3 - len(rest))
        out[i] = (t, q[0], q[1], q[2], param)
    return out


def flip_heavy_gates(n, rng, body, tail):
    """A few dense gates, then a long run of X / CNOT / Toffoli / SWAP: what the compiler folds into store addressing."""
    lst = []
    for _ in range(body):
        lst.append((str(rng.choice(["H", "T", "Ry"])), int(rng.integers(n))) if n else None)
        if lst[-1][0] == "Ry":
            lst[-1] = ("Ry", lst[-1][1], float(rng.uniform(-3, 3)))
    for _ in range(tail):
        kind = str(rng.choice(["X", "CNOT", "CNOT", "CNOT", "Toffoli", "SWAP"]))
        qs = [int(q) for q in rng.permutation(n)[:3]]
        if kind == "X" or n < 2:
            lst.append(("X", qs[0]))
        elif kind in ("CNOT", "SWAP") or n < 3:
            lst.append((kind if kind != "Toffoli" else "CNOT", qs[0], qs[1]))
        else:
            lst.append(("Toffoli", qs[0], qs[1], qs[2]))
    return gates(lst)


def random_gates(n, depth, rng, kinds=None):
    """Random circuit over ALL 17 gate types (the reference's generator only draws H/X/CNOT/Rz)."""
    out = np.zeros(depth, GATE_DTYPE)
    for i in range(depth):
        while True:
            t = int(rng.integers(0, 17)) if kinds is None else int(rng.choice(kinds))
            if t >= 11 and n < 2:
                continue
            if t == 16 and n < 3:
                continue
            break
        qs = rng.permutation(n)
        ang = float(rng.uniform(0, 2 * np.pi)) if t in (8, 9, 10, 13, 14) else 0.0
        out[i] = (t, qs[0], qs[1] if t >= 11 else -1, qs[2] if t == 16 else -1, ang)
    return out


def random_state(n, rng):
    a = rng.normal(size=(1 << n)) + 1j * rng.normal(size=(1 << n))
    return (a / np.linalg.norm(a)).astype(np.complex128)


def zero_state(n):
    a = np.zeros(1 << n, np.complex128)
    a[0] = 1.0
    return a


def oracle_run(n, g, state=None):
    st = zero_state(n) if state is None else np.array(state, np.complex128)
    rc = oracle().orc_run(st.ctypes.data_as(P), n, g.ctypes.data_as(P) if len(g) else None, c_int64(len(g)))
    assert rc == 0
    return st


def oracle_probs(state):
    n = int(np.log2(len(state)))
    p = np.empty(len(state))
    oracle().orc_probabilities(np.ascontiguousarray(state).ctypes.data_as(P), n, p.ctypes.data_as(P))
    return p


def oracle_sample(probs, uniforms):
    u = np.ascontiguousarray(uniforms, np.float64)
    out = np.empty(len(u), np.int64)
    oracle().orc_sample(np.ascontiguousarray(probs).ctypes.data_as(P), c_int64(len(probs)), u.ctypes.data_as(P),
                        c_int64(len(u)), out.ctypes.data_as(P))
    return out


def mt19937_uniforms(seed, count):
    out = np.empty(count)
    oracle().orc_mt19937_uniforms(c_uint(seed), c_int64(count), out.ctypes.data_as(P))
    return out


def emu_run(n, g, state, n_global=0, rank=0, lmin=0, tmax=0, merge=1, reorder=1):
    st = np.array(state, np.complex128)
    info = np.zeros(8, np.int64)
    err = ctypes.create_string_buffer(256)
    rc = emulator().emu_run(n, n_global, rank, g.ctypes.data_as(P) if len(g) else None, c_int64(len(g)),
                            st.ctypes.data_as(P), lmin, tmax, merge, reorder, info.ctypes.data_as(P), err, 256)
    assert rc == 0, err.value
    return st, info


def ref_cpu_run(n, g):
    out = np.zeros(1 << n, np.complex128)
    t = reference().ref_cpu_run(n, g.ctypes.data_as(P) if len(g) else None, c_int64(len(g)), out.ctypes.data_as(P))
    assert t >= 0
    return out


def ref_random_circuit(n, depth, seed):
    g = np.zeros(depth, GATE_DTYPE)
    k = reference().ref_random_circuit(n, depth, c_uint(seed), g.ctypes.data_as(P))
    assert k == depth
    return g


def bench_c1_gates(n=20):
    """benchmark_scaling workload (reference benchmarks/benchmark_scaling.cu:68-75): 100 H + 20 CNOT."""
    lst = []
    for i in range(100):
        lst.append(("H", i % n))
        if i % 5 == 0:
            lst.append(("CNOT", i % n, (i + 1) % n))
    return gates(lst)


def qft_style_circuit(n, window=None):
    """BASELINE config C3: createGHZCircuit(n) (reference src/Circuit.cpp:240-250) followed by QFT-style
    layers h(q); crz(j, q, pi/2^(j-q)) for j > q (optionally windowed), then rz(q, theta_q) with theta_q drawn from
    std::mt19937(42) / uniform_real_distribution(0, 2 pi).  Returns a package Circuit."""
    import cuda_quantum_simulator_b200 as q
    c = q.create_ghz_circuit(n)
    for qb in range(n):
        c.h(qb)
        hi = n if window is None else min(n, qb + 1 + window)
        for j in range(qb + 1, hi):
            c.crz(j, qb, float(np.pi / 2 ** (j - qb)))
    theta = mt19937_uniforms(42, n) * 2 * np.pi
    for qb in range(n):
        c.rz(qb, float(theta[qb]))
    return c


def load_known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)
