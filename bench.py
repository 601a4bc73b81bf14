#!/usr/bin/env python
"""Benchmark of the gate-application hot path (BASELINE.json: gates/s & achieved HBM GB/s at 30 q fp64).

Workload (config.workload): `createRandomCircuit(n, 20, 42)` (reference src/Circuit.cpp:252-282) on an
n-qubit fp64 state, n = 30 on one GPU (BASELINE configs[1], 16 GiB state) and n = 30 + log2(N) on N GPUs
(weak scaling: every GPU keeps a 2^30-amplitude shard; non-diagonal gates on rank qubits go through NVLink exchanges,
fused into the preceding pass — but from |0...0> the qubit layout is chosen so that this circuit needs none; the
`dense_variant` object reports the depth-200 circuit of the same generator).

One "step" = one execution of the whole circuit on the resident state.
  value : gates * 2^(n-30) / s with the state and the compiled program already in HBM
          (= plain gates/s at 30 qubits; the 2^(n-30) factor makes the N-GPU number a whole-job aggregate).
  e2e   : the same through the public API from host inputs: reset -> run(circuit given as host gate records:
          compile + program upload) -> sample(1024 shots, host uniforms) -> indices back on the host.
  roofline : fused_pass_kernel, algorithmic bytes 2*16*2^n_local per launch / its mean CUDA-event duration,
          against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the reference's own CPUSimulator (oracle/_ref, built from the unmodified reference) on the same
          generator at a bounded size, single thread (the reference CPU path has no threading).

`--impl reference` times that CPU implementation alone (the reference arm).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "gate_throughput_30q_equiv"
UNIT = "gates/s (gates*2^(n-30)/s; =gates/s at 30 qubits)"
SHOTS = 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--local-qubits", type=int, default=30, help="qubits per GPU shard (30 = BASELINE configs[1])")
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-qubits", type=int, default=26, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the depth-200 variant")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"])
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(n, depth, seed, repeats=1):
    """Times the reference CPUSimulator::run (oracle/_ref) on createRandomCircuit(n, depth, seed)."""
    import helpers as H
    ref = H.reference()
    kind = "reference"
    g = None
    if ref is not None:
        g = H.ref_random_circuit(n, depth, seed)
        run = lambda: ref.ref_cpu_run(n, g.ctypes.data_as(H.P), H.c_int64(len(g)), None)
    else:  # the reference could not be built here: time the oracle port instead
        import cuda_quantum_simulator_b200 as q
        import numpy as np
        kind = "port"
        g = q.create_random_circuit(n, depth, seed).gates

        def run():
            st = H.zero_state(n)
            t0 = time.perf_counter()
            H.oracle().orc_run(st.ctypes.data_as(H.P), n, g.ctypes.data_as(H.P), H.c_int64(len(g)))
            return time.perf_counter() - t0
    best = min(run() for _ in range(repeats))
    return best, kind, len(g)


def reference_arm(args):
    """`--impl reference`: the reference's own CPU path (CPUSimulator) on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_total = args.local_qubits + int(math.log2(args.gpus))
    n_cpu = min(args.cpu_qubits, n_total)
    times = []
    kind, ng = "reference", args.depth
    for i in range(args.warmup + args.steps):
        t, kind, ng = cpu_reference_run(n_cpu, args.depth, args.seed)
        if i >= args.warmup:
            times.append(t)
    sec = sum(times) / len(times)
    # the bounded sample runs the same generator at n_cpu qubits; per-gate cost scales as 2^n (measured)
    value = ng * 2.0 ** (n_cpu - 30) / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"createRandomCircuit({n_total},{args.depth},{args.seed}) fp64 state vector",
                   "sample": f"same generator at {n_cpu} qubits, scaled by 2^({n_cpu}-30)", "gates": ng},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"CPUSimulator::run on createRandomCircuit({n_cpu},{args.depth},{args.seed}), "
                                   f"{sec:.2f} s per run, 1 of {os.cpu_count()} host cores (reference path is single-threaded)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import cuda_quantum_simulator_b200 as q

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert world & (world - 1) == 0, "GPU count must be a power of two"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_global = int(math.log2(world))
    n_local = args.local_qubits
    n = n_local + n_global
    circuit = q.create_random_circuit(n, args.depth, args.seed)
    n_gates = circuit.get_gate_count()
    stream = torch.cuda.current_stream()

    if world == 1:
        sim = q.Simulator(n)   # library-owned state (what a user of the public API gets)
        sim.set_stream(stream.cuda_stream)
        sim.reset()
        prog = q.CompiledCircuit(circuit)
        step = lambda: sim.execute(prog)
        sync_all = lambda: torch.cuda.synchronize()
        n_passes, n_swaps = prog.n_passes, 0
        runner = sim
    else:
        from cuda_quantum_simulator_b200.sharded import ShardedSimulator
        runner = ShardedSimulator(n, exchange=args.exchange)
        plan = runner.compile(circuit)
        step = lambda: runner.execute(plan)

        def sync_all():
            torch.cuda.synchronize()
            dist.barrier()
        n_passes, n_swaps = plan.n_passes, plan.n_swaps
        sim = runner.local

    def timed(fn, k):
        """K steps between barrier+sync; device time by CUDA events; max over ranks."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- kernel-resident throughput -------------------------------------------------------------
    step()   # untimed and not a warm-up step: the first pass after reset() generates |0..0> on chip; everything timed
             # below runs on the dense evolved state, loaded from and stored to HBM in full
    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sim.set_timing(True)
    launches0 = sim.launch_count()
    total_ms = timed(step, args.steps)
    launches = sim.launch_count() - launches0
    pass_ms, passes_timed = sim.pass_time_ms()
    sim.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n_gates * 2.0 ** (n - 30) / (ms_per_step * 1e-3)

    # ---- end to end through the public API, host buffers in, host results out ----------------------
    uniforms = torch.rand(SHOTS, dtype=torch.float64).pin_memory()
    u_np = uniforms.numpy()
    h2d = d2h = 0
    if world == 1:
        def e2e_step():
            sim.reset()
            sim.run(circuit)                       # compile + upload of the op records (H2D) + launches
            return sim.sample(0, uniforms=u_np)    # uniforms H2D, indices D2H
        h2d = prog.n_ops * 128 + SHOTS * 8
        d2h = SHOTS * 8 + 8
    else:
        def e2e_step():
            runner.reset()
            runner.run(circuit)
            return runner.sample(uniforms=u_np)
        h2d = plan.n_ops * 128 + SHOTS * 8
        d2h = SHOTS * 8 + 8 * world
    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    k_e2e = max(3, min(args.steps, 10))
    for _ in range(k_e2e):
        e2e_step()
    sync_all()
    e2e_sec = torch.tensor([(time.perf_counter() - t0) / k_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_sec, op=dist.ReduceOp.MAX)
    e2e_value = n_gates * 2.0 ** (n - 30) / float(e2e_sec.item())

    # ---- denser variant of the same generator: every qubit is targeted, exchanges cannot be avoided ----
    dense = None
    if not args.no_dense:
        dcirc = q.create_random_circuit(n, 200, args.seed)
        if world == 1:
            dprog = q.CompiledCircuit(dcirc)
            dstep = lambda: sim.execute(dprog)
            d_info = {"passes": dprog.n_passes, "global_qubit_swaps": 0}
        else:
            runner.reset()
            dplan = runner.compile(dcirc)
            f0 = runner.engine.fused_exchanges
            dstep = lambda: runner.execute(dplan)
            d_info = {"passes": dplan.n_passes, "global_qubit_swaps": dplan.n_swaps}
        dstep()
        d_ms = timed(dstep, 3) / 3
        if world > 1:
            d_info["exchanges_fused_into_a_pass"] = (runner.engine.fused_exchanges - f0) // 4
        dense = dict(workload=f"createRandomCircuit({n},200,{args.seed})", gates=dcirc.get_gate_count(), ms_per_step=d_ms,
                     value=dcirc.get_gate_count() * 2.0 ** (n - 30) / (d_ms * 1e-3), **d_info)

    # ---- NVLink leg of a global-qubit swap, timed on its own (two swaps = there and back) ---------------
    nvlink = None
    if world > 1:
        eng = runner.engine
        g, l = n - 1, n_local - 1
        eng.swap(g, l); eng.swap(g, l)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(4):
            eng.swap(g, l)
        e1.record(stream)
        sync_all()
        sw_ms = torch.tensor([e0.elapsed_time(e1) / 4], dtype=torch.float64, device="cuda")
        dist.all_reduce(sw_ms, op=dist.ReduceOp.MAX)
        half_bytes = 16 * (1 << (n_local - 1))
        nvlink = {"exchange": eng.exchange, "bytes_per_direction_per_gpu": half_bytes, "ms": float(sw_ms.item()),
                  "achieved": half_bytes / float(sw_ms.item()) / 1e6, "peak": 770.0, "unit": "GB/s",
                  "frac": half_bytes / float(sw_ms.item()) / 1e6 / 770.0,
                  "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md)"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    bytes_per_pass = 2 * 16 * (1 << n_local)
    avg_pass_ms = pass_ms / max(passes_timed, 1)
    achieved = bytes_per_pass / (avg_pass_ms * 1e-3) / 1e9 if passes_timed else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "pass_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"createRandomCircuit({n},{args.depth},{args.seed}) on a {n}-qubit fp64 state vector "
                               f"({16 * (1 << n_local) / 2**30:.0f} GiB per GPU)",
                   "gates": n_gates, "passes": n_passes, "global_qubit_swaps": n_swaps, "qubits": n,
                   "local_qubits": n_local,
                   "parallelism": f"shard {n_global} qubit(s) over {world} GPU(s); from |0..0> the layout puts qubits that are "
                                  f"never a non-diagonal target in the rank bits (no exchange for this circuit)"
                                  if world > 1 else "one GPU",
                   "l2": f"state ({16 * (1 << n_local) / 2**30:.0f} GiB/GPU) is far larger than the 126 MB L2: no flush needed",
                   "gates_per_s_raw": n_gates / (ms_per_step * 1e-3)},
        "roofline": {"bound": "hbm", "kernel": "fused_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_pass, "avg_launch_ms": avg_pass_ms,
                     "launches_timed": passes_timed,
                     "circuit_level_gbs": n_passes * bytes_per_pass / (ms_per_step * 1e-3) / 1e9},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(e2e_sec.item()) * 1e3,
                "what": "reset + run(host gate records: compile, upload, launch) + sample(1024 host uniforms) + indices to host; "
                        "reset is lazy: the next run zero-fills once (memset) and its first pass generates and processes only the "
                        "tile that holds |0..0>, instead of a memset sweep followed by a full load/compute/store sweep; "
                        "`value` is measured on the dense evolved state"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if nvlink:
        line["nvlink"] = nvlink
    if dense:
        line["dense_variant"] = dense
    if not args.no_cpu_baseline:
        n_cpu = min(args.cpu_qubits, n)
        sec, kind, ng = cpu_reference_run(n_cpu, args.depth, args.seed)
        line["cpu_baseline"] = {
            "value": ng * 2.0 ** (n_cpu - 30) / sec, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"CPUSimulator::run on createRandomCircuit({n_cpu},{args.depth},{args.seed}): {sec:.2f} s, scaled by "
                      f"2^({n_cpu}-30) to the 30-qubit-equivalent unit; 1 of {os.cpu_count()} host cores (the reference CPU path is single-threaded)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
