#!/usr/bin/env python
"""Benchmark of the gate-application hot path (BASELINE.json: gates/s & achieved HBM GB/s at 30 q fp64).

Workload (config.workload): `createRandomCircuit(n, 20, 42)` (reference src/Circuit.cpp:252-282) on an
n-qubit fp64 state, n = 30 on one GPU (BASELINE configs[1], 16 GiB state) and n = 30 + log2(N) on N GPUs
(weak scaling: every GPU keeps a 2^30-amplitude shard; non-diagonal gates on rank qubits go through NVLink exchanges,
fused into the preceding pass - but from |0...0> the qubit layout is chosen so that this circuit needs none).

One "step" = one execution of the whole circuit on the resident state.
  value : gates * 2^(n-30) / s with the state and the compiled program already in HBM
          (= plain gates/s at 30 qubits; the 2^(n-30) factor makes the N-GPU number a whole-job aggregate).
  e2e   : the same through the public API from host inputs: reset -> run(circuit given as host gate records:
          compile + program upload) -> sample(1024 shots, host uniforms) -> indices back on the host.
  roofline : the pass kernel (run-time specialised build at this size), algorithmic bytes 2*16*2^n_local per launch /
          its mean CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the reference's own CPUSimulator (oracle/_ref, built from the unmodified reference) on the same
          generator at a bounded size, single thread (the reference CPU path has no threading).
  dense_variant : createRandomCircuit(n, 200, seed), the depth-200 circuit of the same generator.
  N > 1 only (the C++ driver qsim::ShardedSimulator unless --sharded-driver python):
  parity_check : before anything is timed, the sharded engine against the CPU oracle on circuits that really exchange
          (fused and separate exchanges, both drivers): amplitudes, logical-order sampling, marginals, measurement;
          a mismatch ends the run with exit code 1.
  nvlink : one separate global<->local swap, alone.   forced_exchange : the headline circuit from the identity layout,
          1 pass + 1 exchange per step, and how well the two overlap.   c4 (N = 8) : the real 36-qubit circuit, 128 GiB shards.

`--impl reference` times the reference CPU implementation alone (the reference arm): the REAL 30-qubit circuit.
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
    ap.add_argument("--cpu-qubits", type=int, default=27, help="size of the bounded CPU-baseline sample of the GPU arm")
    ap.add_argument("--ref-qubits", type=int, default=30, help="--impl reference: largest state the host runs for real")
    ap.add_argument("--ref-runs", type=int, default=2, help="--impl reference: timed full-size runs (~50 s each at 30 qubits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the depth-200 variant")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the oracle parity check before timing")
    ap.add_argument("--no-c4", action="store_true", help="N = 8: skip the 36-qubit C4 leg (128 GiB shards)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--sharded-driver", default="native", choices=["native", "python"],
                    help="N > 1: qsim::ShardedSimulator (C++ over NCCL, default) or the Python driver")
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


def workload_name(n, depth, seed, n_local):
    return (f"createRandomCircuit({n},{depth},{seed}) on a {n}-qubit fp64 state vector "
            f"({16 * (1 << n_local) / 2**30:.0f} GiB per GPU)")


def reference_arm(args):
    """`--impl reference`: the reference's own CPU path (CPUSimulator::run, reference src/Simulator.cu:208-212) on the host
    cores; rank 0 only.  At N = 1 it runs the REAL workload - createRandomCircuit(30, depth, seed) on a 16 GiB host state
    vector - not an extrapolation: one circuit takes ~50 s on one core, so min(steps, 2) runs are timed and no warm-up run
    is spent at full size (a CPU loop over a freshly allocated vector has nothing to warm up; the 26-qubit cross-check
    below runs first and loads the code).  For N > 1 the 30 + log2(N)-qubit state (>= 32 GiB, up to 128 GiB) does not
    fit the host, and the metric's unit is 30-qubit equivalents (gates * 2^(n-30) / s), so the same 30-qubit run is the
    sample and the line says so."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_total = args.local_qubits + int(math.log2(args.gpus))
    n_run = min(n_total, args.ref_qubits)
    # cross-check sample (the round-1 figure): same generator at a small size, scaled by 2^(n-30)
    n_x = min(args.cpu_qubits, n_run)
    t_x, kind, ng_x = cpu_reference_run(n_x, args.depth, args.seed)
    runs = max(1, min(args.steps, args.ref_runs))
    times = []
    ng = args.depth
    for _ in range(runs):
        t, kind, ng = cpu_reference_run(n_run, args.depth, args.seed)
        times.append(t)
    sec = sum(times) / len(times)
    value = ng * 2.0 ** (n_run - 30) / sec
    exact = (n_run == n_total)
    sample = (f"CPUSimulator::run on createRandomCircuit({n_run},{args.depth},{args.seed}), the full workload, {runs} timed run(s) of "
              f"{sec:.1f} s each, no warm-up run at this size" if exact else
              f"CPUSimulator::run on createRandomCircuit({n_run},{args.depth},{args.seed}) ({runs} timed run(s) of {sec:.1f} s): the "
              f"{n_total}-qubit state does not fit the host; the unit is 30-qubit equivalents so this is the per-GPU-share sample")
    sample += f"; 1 of {os.cpu_count()} host cores (the reference CPU path is single-threaded)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n_total, args.depth, args.seed, args.local_qubits), "gates": ng, "qubits": n_total,
                   "local_qubits": args.local_qubits, "reference_qubits_run": n_run, "reference_runs_timed": runs,
                   "reference_warmup_runs_at_full_size": 0, "same_workload_as_gpu_arm": exact,
                   "cross_check": {"qubits": n_x, "seconds": t_x, "value_scaled_by_2^(n-30)": ng_x * 2.0 ** (n_x - 30) / t_x}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def make_sharded(n, exchange, driver):
    """The sharded simulator: the C++ driver qsim::ShardedSimulator (NCCL called from C++; torch.distributed only hands
    the NCCL unique id around) or the Python driver (same surface)."""
    from cuda_quantum_simulator_b200 import sharded
    return sharded.NativeShardedSimulator(n, exchange=exchange) if driver == "native" else sharded.ShardedSimulator(n, exchange=exchange)


def sharded_parity_check(exchange, rank, world, drivers=("native", "python")):
    """N > 1 only, before anything is timed: the sharded engine against the CPU oracle on circuits that REALLY exchange
    (24 and 22 qubits forced over `world` shards; identity layout so the gate on the top qubit needs a global<->local swap,
    and a 300-gate circuit that targets every qubit), once with the exchange fused into the preceding pass and once as
    a separate swap kernel; amplitudes <= 1e-10, sampling bit-identical to the reference's logical-order sequential CDF
    after the exchanges, marginals, and one measurement.  Any mismatch ends the run with a non-zero exit code."""
    import numpy as np
    import cuda_quantum_simulator_b200 as q
    import helpers as H
    out = {"drivers": list(drivers), "max_abs_err": 0.0, "exchanges": 0, "fused": 0, "in_place": 0, "split": 0, "cases": [], "sampling_bit_identical": True,
           "marginal_max_err": 0.0, "measure_ok": True, "tolerance": 1e-10}
    cases = (("createRandomCircuit(24,40,7) identity layout", 24, 40, 7, False),
             ("createRandomCircuit(22,300,11) free layout", 22, 300, 11, True))
    # third variant (first driver only): the fused exchange carried by run-time SPECIALISED pass kernels (forced here: these
    # shards are below the size at which they are used by default) - the 40-gate case only, a handful of compiles
    # fourth (C++ driver): the fused exchange IN PLACE - what shards too large for a second buffer get (C4's 128 GiB) - forced
    variants = [(d, m) for d in drivers for m in ("fused", "separate")] + [(drivers[0], "fused+specialised")]
    if "native" in drivers:
        variants.append(("native", "fused in place"))            # (split over the two passes around the exchange where it can be)
        variants.append(("native", "fused in place, unsplit"))
    for driver, mode in variants:
        if mode == "separate":
            os.environ["QSIM_NO_FUSED_EXCHANGE"] = "1"
        if mode.startswith("fused in place"):
            os.environ["QSIM_FORCE_INPLACE_EXCHANGE"] = "1"
        if mode.endswith("unsplit"):
            os.environ["QSIM_NO_SPLIT_EXCHANGE"] = "1"
        if mode == "fused+specialised":
            q.jit_set_mode("always")
        try:
            for label, n, depth, seed, free_layout in (cases[:1] if mode == "fused+specialised" else cases):
                c = q.create_random_circuit(n, depth, seed)
                sim = make_sharded(n, exchange, driver)
                sim.identity_layout_only(not free_layout)
                cp = sim.compile(c)
                f0 = sim.fused_exchanges
                sim.execute(cp)
                got = sim.get_state_vector()
                want = H.oracle_run(n, c.gates)
                err = float(np.max(np.abs(got - want)))
                u = np.concatenate([np.random.default_rng(seed).random(256), [0.0, 0.5]])
                samp_ok = bool(np.array_equal(sim.sample(uniforms=u), H.oracle_sample(H.oracle_probs(got), u)))
                qs = [0, n // 2, n - 1, n - 2]
                idx = np.arange(1 << n)
                oc = np.zeros(1 << n, np.int64)
                for i, qb in enumerate(qs):
                    oc |= ((idx >> qb) & 1) << i
                merr = float(np.max(np.abs(sim.marginal(qs) - np.bincount(oc, weights=np.abs(want) ** 2, minlength=16))))
                # measurement of the top qubit's index bit (Simulator::measureQubit(0) addresses bit n-1)
                pr = np.abs(got) ** 2
                p0 = float(np.sum(pr[((idx >> (n - 1)) & 1) == 0]))
                r = 0.25 if abs(p0 - 0.25) > 1e-6 else 0.6
                meas_ok = True
                if min(p0, 1 - p0) > 1e-9:
                    res = sim.measure_qubit(0, r)
                    keep = ((idx >> (n - 1)) & 1) == res
                    want_c = np.where(keep, got, 0) / np.sqrt(p0 if res == 0 else 1 - p0)
                    meas_ok = res == (0 if r < p0 else 1) and float(np.max(np.abs(sim.get_state_vector() - want_c))) < 1e-10
                out["cases"].append({"circuit": label, "driver": driver, "mode": mode, "exchange": sim.exchange, "swaps": cp.n_swaps,
                                     "fused_into_a_pass": sim.fused_exchanges - f0, "max_abs_err": err})
                out["max_abs_err"] = max(out["max_abs_err"], err)
                out["exchanges"] += cp.n_swaps
                out["fused"] += sim.fused_exchanges - f0
                out["in_place"] += getattr(sim, "inplace_exchanges", 0)
                out["split"] += getattr(sim, "split_exchanges", 0)
                out["sampling_bit_identical"] &= samp_ok
                out["marginal_max_err"] = max(out["marginal_max_err"], merr)
                out["measure_ok"] &= bool(meas_ok)
                sim.release(cp)
                sim.close()
        finally:
            os.environ.pop("QSIM_NO_FUSED_EXCHANGE", None)
            os.environ.pop("QSIM_FORCE_INPLACE_EXCHANGE", None)
            os.environ.pop("QSIM_NO_SPLIT_EXCHANGE", None)
            q.jit_set_mode("auto")
    out["passed"] = bool(out["max_abs_err"] < 1e-10 and out["sampling_bit_identical"] and out["marginal_max_err"] < 1e-12
                         and out["measure_ok"] and out["exchanges"] > 0 and ("native" not in drivers or out["in_place"] > 0))
    return out


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
        parity = None
        if not args.no_parity_check:
            parity = sharded_parity_check(args.exchange, rank, world, ("native", "python") if args.sharded_driver == "native" else ("python",))
            flag = torch.tensor([0.0 if parity["passed"] else 1.0], device="cuda")
            dist.all_reduce(flag)
            if flag.item() != 0:
                if rank == 0:
                    print(json.dumps({"parity_check": parity, "error": "sharded parity check failed"}), flush=True)
                dist.destroy_process_group()
                sys.exit(1)
        runner = make_sharded(n, args.exchange, args.sharded_driver)
        # one plan per step: a run may leave a different qubit layout / X frame behind than it started from, and a plan is
        # only valid for the layout it was compiled against (plans are shared when the layout repeats)
        plans = runner.compile_sequence(circuit, 2 + args.warmup + args.steps)   # (the first two steps are untimed: from |0..0>, and the one that lets late kernel builds finish)
        plan = plans[0]
        step_no = [0]

        def step():
            runner.execute(plans[step_no[0]])
            step_no[0] += 1

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
    q.jit_wait()   # the pass kernels specialised for this circuit are compiled in the background (a one-off cost per pass
                   # structure, cached on disk; jit.compile_seconds in the line): the timed steps run the steady state
    step()         # (a first pass that no longer starts from a basis state, and a compute-heavy pass that has two builds to
    q.jit_wait()   #  choose from, ask for another kernel at their second launch)
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
    q.jit_wait()
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
            dplans = runner.compile_sequence(dcirc, 10)
            f0 = runner.fused_exchanges
            d_no = [0]

            def dstep():
                runner.execute(dplans[d_no[0]])
                d_no[0] += 1
            d_info = {"passes_per_step": [p_.n_passes for p_ in dplans[7:]],
                      "global_qubit_swaps_per_step": [p_.n_swaps for p_ in dplans[7:]]}
        for _ in range(3):     # untimed: from |0..0>; kernel builds asked for at the second launch of a pass; ...
            dstep()
            q.jit_wait()
        for _ in range(4):     # ... and the two builds of a compute-heavy pass timed against each other in passing (the faster stays)
            dstep()
        f1 = runner.fused_exchanges if world > 1 else 0
        d_ms = timed(dstep, 3) / 3
        if world > 1:
            d_info["exchanges_fused_into_a_pass_in_the_timed_steps"] = runner.fused_exchanges - f1
        dense = dict(workload=f"createRandomCircuit({n},200,{args.seed})", gates=dcirc.get_gate_count(), ms_per_step=d_ms,
                     value=dcirc.get_gate_count() * 2.0 ** (n - 30) / (d_ms * 1e-3), **d_info)

    # ---- NVLink leg of a global-qubit swap, timed on its own (two swaps = there and back) ---------------
    nvlink = None
    if world > 1:
        eng = runner
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

    # ---- the same circuit with the IDENTITY layout: the gate on the top qubit forces a global<->local exchange -----------
    forced = None
    if world > 1:
        def forced_run(circ, sim_, k):
            """k timed steps of `circ`, every one from the IDENTITY qubit layout (no free initial layout), so that the gate
            on the top qubit needs its global<->local exchange every time: between steps the stored layout is declared to be
            the identity again (relabel_identity: bookkeeping only, no data moves; the state changes by a qubit permutation,
            which a dense state does not care about).  Device time."""
            sim_.identity_layout_only(True)
            sim_.reset()
            first = sim_.compile(circ)                   # from |0..0>, identity layout
            f_a = sim_.fused_exchanges
            sim_.execute(first)                          # untimed (basis-state input)
            f_b = sim_.fused_exchanges
            sim_.relabel_identity()
            plan_ = sim_.compile(circ)                   # against a dense state in the identity layout

            def st():
                sim_.relabel_identity()
                sim_.execute(plan_)
            st()
            q.jit_wait()
            st()
            f_c = sim_.fused_exchanges
            i_c = getattr(sim_, "inplace_exchanges", 0)
            p_c = getattr(sim_, "split_exchanges", 0)
            ms = timed(st, k) / k
            info = {"ms_per_step": ms, "passes_per_step": plan_.n_passes, "swaps_per_step": plan_.n_swaps,
                    "fused_into_a_pass_per_step": (sim_.fused_exchanges - f_c) / k,
                    "of_which_in_place_per_step": (getattr(sim_, "inplace_exchanges", 0) - i_c) / k,
                    "of_which_split_over_two_passes_per_step": (getattr(sim_, "split_exchanges", 0) - p_c) / k,
                    "fused_first_step_from_zero_state": f_b - f_a}
            sim_.identity_layout_only(False)
            return info, plan_
        k_f = 4
        f_info, fplan = forced_run(circuit, runner, k_f)
        pass_ms_1 = pass_ms / max(passes_timed, 1)
        link_ms = nvlink["ms"]
        # an exchange can only overlap the pass it is fused into (the next pass needs the exchanged data), so the ideal step is
        # max(pass, link) per exchange plus the remaining passes
        n_f = min(fplan.n_swaps, fplan.n_passes)
        ideal = n_f * max(pass_ms_1, link_ms) + (fplan.n_passes - n_f) * pass_ms_1 + (fplan.n_swaps - n_f) * link_ms
        fused_step_ms = (f_info["ms_per_step"] - (fplan.n_passes - n_f) * pass_ms_1) / max(n_f, 1)
        half_bytes = 16 * (1 << (n_local - 1))
        forced = dict(workload=f"createRandomCircuit({n},{args.depth},{args.seed}), identity qubit layout (QSIM_NO_LAYOUT semantics)",
                      exchange=runner.exchange, driver=args.sharded_driver, pass_ms=pass_ms_1, link_ms=link_ms,
                      ideal_ms_per_step=ideal, overlap_eff=ideal / f_info["ms_per_step"],
                      value=n_gates * 2.0 ** (n - 30) / (f_info["ms_per_step"] * 1e-3),
                      fused_pass_plus_exchange_ms=fused_step_ms,
                      nvlink_gbs_per_direction_in_the_fused_step=half_bytes / fused_step_ms / 1e6,
                      nvlink_frac_of_770=half_bytes / fused_step_ms / 1e6 / 770.0,
                      what="every timed step starts from the identity layout, so the gate on the top qubit needs its global<->local "
                           "exchange every time; the exchange is fused into the pass before it when the second buffer exists. "
                           "ideal = max(pass_ms, link_ms) per exchange + the remaining passes (an exchange can only overlap the pass "
                           "it rides on); overlap_eff = ideal / measured; fused_pass_plus_exchange_ms = measured minus the unfused "
                           "passes; pass_ms = the headline pass, link_ms = the separate-swap nvlink leg above", **f_info)
        if args.sharded_driver == "native" and runner.exchange == "p2p":
            # the same steps without a second buffer (what shards that fill the GPU get: C4): the exchange fused IN PLACE into
            # the pass before it, and SPLIT over the pass before (scatters half) and the pass after (gathers half)
            for leg, env in (("in_place", {"QSIM_FORCE_INPLACE_EXCHANGE": "1", "QSIM_NO_SPLIT_EXCHANGE": "1"}),
                             ("split", {"QSIM_FORCE_INPLACE_EXCHANGE": "1"})):
                os.environ.update(env)
                try:
                    ip_info, _ip = forced_run(circuit, runner, k_f)
                    runner.synchronize()
                finally:
                    for k_ in env:
                        os.environ.pop(k_, None)
                ip_step = (ip_info["ms_per_step"] - (fplan.n_passes - n_f) * pass_ms_1) / max(n_f, 1)
                forced[leg] = dict(overlap_eff=ideal / ip_info["ms_per_step"], fused_pass_plus_exchange_ms=ip_step,
                                   nvlink_gbs_per_direction_in_the_fused_step=half_bytes / ip_step / 1e6,
                                   nvlink_frac_of_770=half_bytes / ip_step / 1e6 / 770.0, **ip_info)
            forced["split"]["what"] = ("both passes carry half of the exchange, so the step can beat `ideal` (which lets an exchange "
                                       "overlap one pass only): ideal for the split is max(2 passes, link) = "
                                       f"{max(2 * pass_ms_1, link_ms):.2f} ms")

    # ---- BASELINE config C4 for real: createRandomCircuit(36,20,42) over 8 GPUs, 128 GiB shards --------------------------
    c4 = None
    if world == 8 and n_local == 30 and not args.no_c4:
        sync_all()
        runner.close()          # drops the 30-qubit shards (and their second buffers) on every rank
        sync_all()
        torch.cuda.empty_cache()
        n4 = 36
        c4circ = q.create_random_circuit(n4, 20, 42)
        free_b, _tot = torch.cuda.mem_get_info()
        fits = torch.tensor([1.0 if free_b > (128 << 30) + (6 << 30) else 0.0], device="cuda")
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    if world == 8 and n_local == 30 and not args.no_c4 and fits.item() == 0:
        c4 = {"skipped": f"a 128 GiB shard does not fit next to what is resident (free on rank 0: {free_b / 2**30:.0f} GiB)"}
    elif world == 8 and n_local == 30 and not args.no_c4:
        r4 = make_sharded(n4, args.exchange, args.sharded_driver)
        c4 = {"workload": "createRandomCircuit(36,20,42), 2^33 amplitudes (128 GiB) per GPU", "gates": c4circ.get_gate_count(),
              "second_buffer_for_fused_exchange": r4.has_second_buffer, "exchange": r4.exchange, "driver": args.sharded_driver}
        # (a) layout chosen from |0..0>
        seq = r4.compile_sequence(c4circ, 4)
        no = [0]

        def st4():
            r4.execute(seq[no[0]])
            no[0] += 1
        st4()
        q.jit_wait()
        st4()
        ms = timed(st4, 2) / 2
        c4["chosen_layout"] = {"ms_per_circuit": ms, "passes": seq[2].n_passes, "swaps_per_step": [p_.n_swaps for p_ in seq[2:]],
                               "gates_per_s": c4circ.get_gate_count() / (ms * 1e-3),
                               "hbm_gbs_per_gpu": seq[2].n_passes * 2 * 16 * (1 << 33) / (ms * 1e-3) / 1e9}
        tot = r4.get_total_probability()
        c4["total_probability"] = tot
        # (b) identity layout: H(35) needs the 64 GiB-each-way exchange
        # fused into the pass before it IN PLACE (no room for a second buffer); if that fails, and for comparison, the
        # separate swap kernel
        try:
            f_info, _fp = forced_run(c4circ, r4, 2)
            f_info["total_probability"] = r4.get_total_probability()   # (also where a failed handshake would surface, on every rank)
            c4["identity_layout"] = f_info
        except Exception as exc:   # noqa: BLE001 - reported in the line
            c4["identity_layout_error"] = str(exc)[:300]
        os.environ["QSIM_NO_INPLACE_EXCHANGE"] = "1"
        try:
            s_info, _sp = forced_run(c4circ, r4, 2)
            s_info["total_probability"] = r4.get_total_probability()
            c4["identity_layout_separate_swap"] = s_info
        finally:
            os.environ.pop("QSIM_NO_INPLACE_EXCHANGE", None)
        r4.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    bytes_per_pass = 2 * 16 * (1 << n_local)
    avg_pass_ms = pass_ms / max(passes_timed, 1)
    achieved = bytes_per_pass / (avg_pass_ms * 1e-3) / 1e9 if passes_timed else None
    # DRAM traffic per launch: from the committed `ncu --set full` capture of this very kernel and workload (a number taken
    # under a profiler is not re-measured inside a bench run); only meaningful for 30 local qubits
    traffic, traffic_src = None, None
    jit = q.jit_stats()
    specialised = jit["launches"] > 0
    tpath = os.path.join(ROOT, "profiles", "ncu_jit_pass_c2_30q_r02.json" if specialised else "pass_kernel_traffic.json")
    if os.path.exists(tpath) and n_local == 30 and args.depth == 20 and args.seed == 42:
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["launches"][0]["dram_bytes_per_launch"] if "launches" in tj else tj.get("dram_bytes_per_launch")
        traffic_src = f"committed ncu capture profiles/{os.path.basename(tpath)} (same kernel and workload; not measured in this run)"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, args.depth, args.seed, n_local),
                   "gates": n_gates, "passes": n_passes, "global_qubit_swaps": n_swaps, "qubits": n,
                   "local_qubits": n_local,
                   "sharded_driver": args.sharded_driver if world > 1 else None,
                   "parallelism": f"shard {n_global} qubit(s) over {world} GPU(s); from |0..0> the layout puts qubits that are "
                                  f"never a non-diagonal target in the rank bits (no exchange for this circuit)"
                                  if world > 1 else "one GPU",
                   "l2": f"state ({16 * (1 << n_local) / 2**30:.0f} GiB/GPU) is far larger than the 126 MB L2: no flush needed",
                   "gates_per_s_raw": n_gates / (ms_per_step * 1e-3)},
        "roofline": {"bound": "hbm",
                     "kernel": "qsim_jit_pass (the pass kernel specialised at run time for this pass, csrc/jit.cpp)" if specialised
                               else "fused_pass_kernel (the pass kernel's ahead-of-time interpreter build)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
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
        "jit": jit,
    }
    if nvlink:
        line["nvlink"] = nvlink
    if world > 1:
        line["parity_check"] = parity
        line["forced_exchange"] = forced
    if c4:
        line["c4"] = c4
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
