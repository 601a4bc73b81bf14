// TEST INFRASTRUCTURE ONLY — never linked or loaded by the product path.
//
// C-ABI shim around the UNMODIFIED reference implementation.  It is compiled
// (by oracle/Makefile) together with the reference's own sources where they
// lie under /root/reference/src; the output goes to oracle/_ref/libqsim_ref.so.
// Nothing from the reference is copied into this repository: this file only
// *calls* the reference's public C++ API (include/Simulator.hpp:53-112,
// include/Circuit.hpp:89-144, include/NoiseModel.cuh:141-297).
//
//  * ref_cpu_*   -> qsim::CPUSimulator  (src/Simulator.cu:195-345), host only,
//                  runs without a GPU: this is the parity oracle and the
//                  `cpu_baseline` / `--impl reference` arm of bench.py.
//  * ref_gpu_*   -> qsim::Simulator     (src/Simulator.cu:22-189) with the
//                  reference's naive kernels recompiled for sm_100a: second,
//                  GPU-side oracle (the only reference code that implements
//                  CRY/CRZ/Toffoli) and the "naive kernel" speed baseline.
//  * ref_random_circuit -> qsim::createRandomCircuit (src/Circuit.cpp:252-282).
#include "Simulator.hpp"
#include "Circuit.hpp"
#include "NoiseModel.cuh"

#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct ref_gate {            // must match qsim_gate_t in include/qsim_b200.h
    int32_t type;            // qsim::GateType order (include/Circuit.hpp:42-59)
    int32_t q0, q1, q2;
    double  param;
};

qsim::Circuit build(int n, const ref_gate* g, int64_t ng) {
    qsim::Circuit c(n);
    for (int64_t i = 0; i < ng; ++i) {
        const ref_gate& x = g[i];
        switch (static_cast<qsim::GateType>(x.type)) {
            case qsim::GateType::X:    c.x(x.q0); break;
            case qsim::GateType::Y:    c.y(x.q0); break;
            case qsim::GateType::Z:    c.z(x.q0); break;
            case qsim::GateType::H:    c.h(x.q0); break;
            case qsim::GateType::S:    c.s(x.q0); break;
            case qsim::GateType::T:    c.t(x.q0); break;
            case qsim::GateType::Sdag: c.sdag(x.q0); break;
            case qsim::GateType::Tdag: c.tdag(x.q0); break;
            case qsim::GateType::Rx:   c.rx(x.q0, x.param); break;
            case qsim::GateType::Ry:   c.ry(x.q0, x.param); break;
            case qsim::GateType::Rz:   c.rz(x.q0, x.param); break;
            case qsim::GateType::CNOT: c.cnot(x.q0, x.q1); break;
            case qsim::GateType::CZ:   c.cz(x.q0, x.q1); break;
            case qsim::GateType::CRY:  c.cry(x.q0, x.q1, x.param); break;
            case qsim::GateType::CRZ:  c.crz(x.q0, x.q1, x.param); break;
            case qsim::GateType::SWAP: c.swap(x.q0, x.q1); break;
            case qsim::GateType::Toffoli: c.toffoli(x.q0, x.q1, x.q2); break;
            default: throw std::runtime_error("bad gate type");
        }
    }
    return c;
}

void emit(const qsim::Circuit& c, ref_gate* out) {
    size_t i = 0;
    for (const auto& g : c.getGates()) {
        ref_gate r{static_cast<int32_t>(g.type), -1, -1, -1, g.parameter};
        if (g.qubits.size() > 0) r.q0 = g.qubits[0];
        if (g.qubits.size() > 1) r.q1 = g.qubits[1];
        if (g.qubits.size() > 2) r.q2 = g.qubits[2];
        out[i++] = r;
    }
}

}  // namespace

// Run `g[0..ng)` on |0..0> with the reference CPUSimulator; write 2^n (re,im)
// pairs to out (may be NULL for timing only).  Returns run() seconds, <0 on error.
REF_API double ref_cpu_run(int n, const ref_gate* g, int64_t ng, double* out) {
    try {
        qsim::Circuit c = build(n, g, ng);
        qsim::CPUSimulator sim(n);
        auto t0 = std::chrono::steady_clock::now();
        sim.run(c);
        auto t1 = std::chrono::steady_clock::now();
        if (out) {
            auto sv = sim.getStateVector();
            std::memcpy(out, sv.data(), sv.size() * sizeof(std::complex<double>));
        }
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (...) { return -1.0; }
}

// CPUSimulator::getProbabilities (src/Simulator.cu:319-325) after running the circuit.
REF_API double ref_cpu_probabilities(int n, const ref_gate* g, int64_t ng, double* probs) {
    try {
        qsim::Circuit c = build(n, g, ng);
        qsim::CPUSimulator sim(n);
        sim.run(c);
        auto p = sim.getProbabilities();
        std::memcpy(probs, p.data(), p.size() * sizeof(double));
        return 0.0;
    } catch (...) { return -1.0; }
}

// qsim::createRandomCircuit(n, depth, seed) -> gate records (out has `depth` slots).
REF_API int ref_random_circuit(int n, int depth, unsigned seed, ref_gate* out) {
    try {
        qsim::Circuit c = qsim::createRandomCircuit(n, depth, seed);
        emit(c, out);
        return static_cast<int>(c.getGateCount());
    } catch (...) { return -1; }
}

REF_API int ref_ghz_circuit(int n, ref_gate* out) {
    try {
        qsim::Circuit c = qsim::createGHZCircuit(n);
        emit(c, out);
        return static_cast<int>(c.getGateCount());
    } catch (...) { return -1; }
}

REF_API int64_t ref_circuit_depth(int n, const ref_gate* g, int64_t ng) {
    try { return static_cast<int64_t>(build(n, g, ng).getDepth()); } catch (...) { return -1; }
}

// Error-contract probe: 0 = accepted, 1 = std::invalid_argument, 2 = std::out_of_range, 3 = other.
REF_API int ref_circuit_error_code(int n, const ref_gate* g, int64_t ng) {
    try { build(n, g, ng); return 0; }
    catch (const std::invalid_argument&) { return 1; }
    catch (const std::out_of_range&) { return 2; }
    catch (...) { return 3; }
}

// ---- GPU-side reference (needs a device; used only on the GPU box) -------------
REF_API double ref_gpu_run(int n, const ref_gate* g, int64_t ng, double* out, int reps) {
    try {
        qsim::Circuit c = build(n, g, ng);
        qsim::Simulator sim(n);
        double best = 1e300;
        for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
            sim.reset();
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            sim.run(c);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            if (ms * 1e-3 < best) best = ms * 1e-3;
        }
        if (out) {
            auto sv = sim.getStateVector();
            std::memcpy(out, sv.data(), sv.size() * sizeof(std::complex<double>));
        }
        return best;
    } catch (...) { return -1.0; }
}

// Reference BatchedSimulator (depolarizing only; D7/D8 in SURVEY.md) for the C5 speed baseline.
REF_API double ref_gpu_batched_run(int n, int batch, const ref_gate* g, int64_t ng,
                                   const int* dep_qubits, int n_dep, double p_dep,
                                   unsigned seed, double* avg_probs) {
    try {
        qsim::Circuit c = build(n, g, ng);
        qsim::NoiseModel nm;
        if (n_dep > 0) nm.addDepolarizing(std::vector<int>(dep_qubits, dep_qubits + n_dep), p_dep);
        qsim::BatchedSimulator sim(n, batch, nm);
        sim.setSeed(seed);
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        sim.run(c);
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        if (avg_probs) {
            auto p = sim.getAverageProbabilities();
            std::memcpy(avg_probs, p.data(), p.size() * sizeof(double));
        }
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (...) { return -1.0; }
}
