"""Generates tests/golden/ref_cpu_states.npz by running the UNMODIFIED reference CPUSimulator
(oracle/_ref/libqsim_ref.so, built from /root/reference by oracle/Makefile) on the circuits the
reference's own GPU-vs-CPU suite uses (tests/test_gpu_cpu_equivalence.cu:122-312) plus its
benchmark workload.  Run in the build container (the reference is not on the GPU box):

    make -C oracle && python tests/golden/make_golden.py

The fixture holds, per case, the gate records and the final state from |0...0>.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as H  # noqa: E402

PI = np.pi


def cases():
    out = {}
    # SingleQubitGates_AllTypes (:122-152): h(0).h(1).h(2) then each 1q gate on each qubit
    singles = [("X",), ("Y",), ("Z",), ("H",), ("S",), ("T",), ("Sdag",), ("Tdag",), ("Rx", PI / 3), ("Ry", PI / 5), ("Rz", PI / 7)]
    for gi, g in enumerate(singles):
        for q in range(3):
            lst = [("H", 0), ("H", 1), ("H", 2), (g[0], q) + tuple(float(x) for x in g[1:])]
            out[f"single_{g[0]}_q{q}"] = (3, H.gates(lst))
    # CNOT / CZ over all ordered pairs on 4 qubits (:158-190), SWAP on an asymmetric state (:192-206)
    for a in range(4):
        for b in range(4):
            if a != b:
                out[f"cnot_{a}_{b}"] = (4, H.gates([("H", 0), ("H", 1), ("H", 2), ("H", 3), ("T", a), ("CNOT", a, b)]))
                out[f"cz_{a}_{b}"] = (4, H.gates([("H", 0), ("H", 1), ("H", 2), ("H", 3), ("CZ", a, b)]))
    out["swap_asym"] = (3, H.gates([("H", 0), ("T", 0), ("X", 1), ("Rx", 2, 0.3), ("SWAP", 0, 2), ("SWAP", 1, 2)]))
    # GHZ 2..8 (:208-225)
    for n in range(2, 9):
        out[f"ghz_{n}"] = (n, H.gates([("H", 0)] + [("CNOT", i, i + 1) for i in range(n - 1)]))
    # Random circuits: small (:227-238), medium (:240-251), deep (:253-275) — the reference's generator
    for seed in range(20):
        n, d = 3 + seed % 3, 10 + seed % 20
        out[f"random_small_s{seed}"] = (n, H.ref_random_circuit(n, d, seed))
    for seed in range(10):
        n, d = 8 + seed % 4, 50 + seed % 50
        out[f"random_medium_s{seed}"] = (n, H.ref_random_circuit(n, d, seed))
    for seed in range(5):
        out[f"random_deep_s{seed}"] = (4, H.ref_random_circuit(4, 500, seed))
    # RotationGates_VariousAngles (:281-312)
    for theta in [0.0, PI / 8, PI / 4, PI / 3, PI / 2, 2 * PI / 3, PI, 3 * PI / 2, 2 * PI, 0.1, 0.7, 1.23, 2.5, 4.0, 5.5]:
        for r in ("Rx", "Ry", "Rz"):
            out[f"rot_{r}_{theta:.4f}"] = (2, H.gates([("H", 0), ("H", 1), (r, 0, float(theta)), ("CNOT", 0, 1)]))
    out["empty"] = (4, H.gates([]))
    # benchmark_scaling workload at 12 qubits (benchmarks/benchmark_scaling.cu:68-75; 20 q is checked live)
    out["bench_scaling_12"] = (12, H.bench_c1_gates(12))
    # the headline generator at a size that fits a fixture
    out["random_c2_like_12"] = (12, H.ref_random_circuit(12, 20, 42))
    return out


def main():
    assert H.reference() is not None, "build oracle/_ref first (make -C oracle)"
    blob = {}
    for name, (n, g) in cases().items():
        blob[name + "__n"] = np.int32(n)
        blob[name + "__gates"] = g
        blob[name + "__state"] = H.ref_cpu_run(n, g)
    path = os.path.join(H.GOLDEN, "ref_cpu_states.npz")
    np.savez_compressed(path, **blob)
    print(f"{len(blob) // 3} cases -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
