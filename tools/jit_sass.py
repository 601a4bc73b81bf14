"""Compile one pass of a circuit with the run-time kernel generator (no GPU needed) and print the opcode mix of the per-tile
compute region (between the tile's mbarrier wait and the TMA store).  usage: jit_sass.py c2|dense|c3 [pass]"""
import collections
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_quantum_simulator_b200 as q


def circuit(name):
    if name == "c2":
        return q.create_random_circuit(30, 20, 42)
    if name == "dense":
        return q.create_random_circuit(30, 200, 42)
    if name == "c3":
        import helpers as H
        return H.qft_style_circuit(33)
    raise SystemExit("unknown circuit")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    c = circuit(name)
    p = q.CompiledCircuit(c)
    print(p.describe())
    passes = range(p.n_passes) if len(sys.argv) < 3 else [int(sys.argv[2])]
    os.makedirs("/tmp/jit", exist_ok=True)
    for i in passes:
        t = time.time()
        cub = p.jit_compile(i, want_cubin=True)
        dt = time.time() - t
        path = f"/tmp/jit/{name}_{i}.cubin"
        open(path, "wb").write(cub)
        sass = subprocess.run(["nvdisasm", "-c", path], capture_output=True, text=True).stdout.split("\n")
        res = subprocess.run(["cuobjdump", "--dump-resource-usage", path], capture_output=True, text=True).stdout
        regs = re.search(r"REG:(\d+)", res).group(1)
        # compute region: first LDS.128 of amplitudes (non-uniform address) .. first UTMASTG
        start = next(k for k, l in enumerate(sass) if "TRYWAIT" in l and "R53" in l or ("LDS.128" in l and "[R" in l))
        end = next(k for k, l in enumerate(sass) if "UTMASTG" in l)
        mix = collections.Counter()
        for l in sass[start:end]:
            m = re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T] )?([A-Z0-9_]+)", l)
            if m:
                mix[m.group(1)] += 1
        tot = sum(mix.values())
        print(f"pass {i}: compile {dt:.2f}s regs {regs} cubin {len(cub)} B; ~{tot} instrs in the tile region:",
              ", ".join(f"{k} {v}" for k, v in mix.most_common(12)))


if __name__ == "__main__":
    main()
