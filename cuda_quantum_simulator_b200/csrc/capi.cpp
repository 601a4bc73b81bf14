// C ABI (include/qsim_b200.h) over the qsim C++ classes.  Exceptions become status codes:
// invalid_argument -> 1, out_of_range -> 2, anything else -> 3; the message is kept per thread.
#include "qsim_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "engine.hpp"
#include "jit.hpp"
#include "program.hpp"
#include "qsim/circuit.hpp"
#include "qsim/constants.hpp"
#include "qsim/density_matrix.cuh"
#include "qsim/noise_model.cuh"
#include "qsim/sharded_simulator.hpp"
#include "qsim/simulator.hpp"
#include "shard.cuh"
#include "sharded_plan.hpp"

using namespace qsim;

namespace {

thread_local std::string g_error;

template <class F>
qsim_status_t guarded(F&& body) {
    try {
        body();
        return QSIM_OK;
    } catch (const std::invalid_argument& e) {
        g_error = e.what();
        return QSIM_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range& e) {
        g_error = e.what();
        return QSIM_ERR_OUT_OF_RANGE;
    } catch (const std::exception& e) {
        g_error = e.what();
        return QSIM_ERR_RUNTIME;
    } catch (...) {
        g_error = "unknown error";
        return QSIM_ERR_RUNTIME;
    }
}

void require(bool cond, const char* what) {
    if (!cond) throw std::invalid_argument(what);
}

// Replays a gate record through the fluent builder so that validation is the builder's own.
void append(Circuit& c, const qsim_gate_t& g) {
    switch (g.type) {
        case QSIM_GATE_X: c.x(g.q0); break;
        case QSIM_GATE_Y: c.y(g.q0); break;
        case QSIM_GATE_Z: c.z(g.q0); break;
        case QSIM_GATE_H: c.h(g.q0); break;
        case QSIM_GATE_S: c.s(g.q0); break;
        case QSIM_GATE_T: c.t(g.q0); break;
        case QSIM_GATE_SDAG: c.sdag(g.q0); break;
        case QSIM_GATE_TDAG: c.tdag(g.q0); break;
        case QSIM_GATE_RX: c.rx(g.q0, g.param); break;
        case QSIM_GATE_RY: c.ry(g.q0, g.param); break;
        case QSIM_GATE_RZ: c.rz(g.q0, g.param); break;
        case QSIM_GATE_CNOT: c.cnot(g.q0, g.q1); break;
        case QSIM_GATE_CZ: c.cz(g.q0, g.q1); break;
        case QSIM_GATE_CRY: c.cry(g.q0, g.q1, g.param); break;
        case QSIM_GATE_CRZ: c.crz(g.q0, g.q1, g.param); break;
        case QSIM_GATE_SWAP: c.swap(g.q0, g.q1); break;
        case QSIM_GATE_TOFFOLI: c.toffoli(g.q0, g.q1, g.q2); break;
        default: throw std::runtime_error("Unknown gate type");
    }
}

Circuit build(int n, const qsim_gate_t* gates, int64_t ng) {
    Circuit c(n);
    for (int64_t i = 0; i < ng; ++i) append(c, gates[i]);
    return c;
}

void emit(const Circuit& c, qsim_gate_t* out) {
    size_t i = 0;
    for (const GateOp& g : c.getGates()) {
        qsim_gate_t r{static_cast<int32_t>(g.type), -1, -1, -1, g.parameter};
        if (g.qubits.size() > 0) r.q0 = g.qubits[0];
        if (g.qubits.size() > 1) r.q1 = g.qubits[1];
        if (g.qubits.size() > 2) r.q2 = g.qubits[2];
        out[i++] = r;
    }
}

}  // namespace

struct qsim_noisy { std::unique_ptr<NoisySimulator> sim; };
struct qsim_batched { std::unique_ptr<BatchedSimulator> sim; };
struct qsim_dm { std::unique_ptr<DensityMatrixSimulator> sim; };

namespace {

NoiseModel to_model(const qsim_noise_channel_t* ch, int n) {
    NoiseModel m;
    for (int i = 0; i < n; ++i) {
        const qsim_noise_channel_t& c = ch[i];
        require(c.type >= 0 && c.type <= QSIM_NOISE_BIT_PHASE_FLIP, "unknown noise type");
        const std::vector<int> q(c.qubits, c.qubits + (c.n_qubits > 0 ? c.n_qubits : 0));
        const bool all = c.n_qubits <= 0;
        switch (c.type) {
            case QSIM_NOISE_DEPOLARIZING: all ? m.addDepolarizing(c.probability) : m.addDepolarizing(q, c.probability); break;
            case QSIM_NOISE_AMPLITUDE_DAMPING: all ? m.addAmplitudeDamping(c.probability) : m.addAmplitudeDamping(q, c.probability); break;
            case QSIM_NOISE_PHASE_DAMPING: all ? m.addPhaseDamping(c.probability) : m.addPhaseDamping(q, c.probability); break;
            case QSIM_NOISE_BIT_FLIP: all ? m.addBitFlip(c.probability) : m.addBitFlip(q, c.probability); break;
            case QSIM_NOISE_PHASE_FLIP: all ? m.addPhaseFlip(c.probability) : m.addPhaseFlip(q, c.probability); break;
            default: all ? m.addBitPhaseFlip(c.probability) : m.addBitPhaseFlip(q, c.probability); break;
        }
    }
    return m;
}

template <class T>
void copy_out(const std::vector<T>& v, T* out) {
    std::memcpy(out, v.data(), v.size() * sizeof(T));
}

}  // namespace

struct qsim_program {
    b200::DeviceProgram dev;
    int n_global = 0;
};

struct qsim_sim {
    std::unique_ptr<Simulator> sim;
    int n_total = 0;
    int n_global = 0;
    int rank = 0;
    uint64_t hi_bits() const { return (uint64_t)rank << (n_total - n_global); }
};

extern "C" {

const char* qsim_last_error(void) { return g_error.c_str(); }
const char* qsim_version(void) { return "qsim_b200 0.1 (sm_100a fused-pass engine)"; }
int qsim_max_qubits(void) { return cuda_config::MAX_QUBITS; }

// ---- circuits ---------------------------------------------------------------------------------

qsim_status_t qsim_circuit_validate(int n, const qsim_gate_t* gates, int64_t ng) {
    return guarded([&] { build(n, gates, ng); });
}

qsim_status_t qsim_circuit_random(int n, int depth, unsigned seed, qsim_gate_t* out) {
    return guarded([&] { emit(createRandomCircuit(n, depth, seed), out); });
}

qsim_status_t qsim_circuit_ghz(int n, qsim_gate_t* out) {
    return guarded([&] { emit(createGHZCircuit(n), out); });
}

qsim_status_t qsim_circuit_depth(int n, const qsim_gate_t* gates, int64_t ng, int64_t* depth) {
    return guarded([&] { *depth = (int64_t)build(n, gates, ng).getDepth(); });
}

// ---- programs ---------------------------------------------------------------------------------

qsim_status_t qsim_program_compile(int n, int n_global, const qsim_gate_t* gates, int64_t ng, qsim_program_t** out) {
    return qsim_program_compile_ex(n, n_global, gates, ng, 0, out);
}

qsim_status_t qsim_program_compile_ex(int n, int n_global, const qsim_gate_t* gates, int64_t ng, uint64_t initial_xor,
                                      qsim_program_t** out) {
    return qsim_program_compile_ex2(n, n_global, gates, ng, initial_xor, -1, out);
}

qsim_status_t qsim_program_compile_ex2(int n, int n_global, const qsim_gate_t* gates, int64_t ng, uint64_t initial_xor,
                                       int isolate_qubit, qsim_program_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        require(n_global >= 0 && n_global < n, "n_global out of range");
        build(n, gates, ng);   // validation only
        auto p = std::make_unique<qsim_program>();
        p->n_global = n_global;
        b200::CompileOptions opt = b200::default_options();
        opt.n_global = n_global;
        opt.initial_xor = initial_xor;
        opt.isolate_bit = isolate_qubit;
        std::string err;
        if (!b200::compile(n, gates, ng, opt, p->dev.host, &err)) throw std::runtime_error(err);
        // the device copy is made now when a GPU is there, else at the first execute (compiling, describing and
        // inspecting the generated kernels need no device)
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) == cudaSuccess && n_dev > 0) p->dev.upload();
        else cudaGetLastError();
        *out = p.release();
    });
}

void qsim_program_destroy(qsim_program_t* p) { delete p; }

qsim_status_t qsim_program_info(const qsim_program_t* p, int64_t info[8]) {
    return guarded([&] {
        require(p != nullptr, "null program");
        const b200::Program& h = p->dev.host;
        int64_t sweeps = 0;
        for (auto& ps : h.passes) sweeps += ps.n_sweeps;
        info[0] = (int64_t)h.passes.size();
        info[1] = (int64_t)h.lops.size();
        info[2] = h.n_gates;
        info[3] = sweeps;
        info[4] = h.passes.empty() ? 0 : h.passes[0].t;
        info[5] = h.n_local;
        info[6] = (int64_t)h.global_xor;
        info[7] = 0;
    });
}

static size_t copy_out(const std::string& d, char* buf, size_t cap) {
    if (buf && cap) {
        const size_t k = d.size() < cap - 1 ? d.size() : cap - 1;
        std::memcpy(buf, d.data(), k);
        buf[k] = 0;
    }
    return d.size() + 1;
}

qsim_status_t qsim_jit_set_mode(int mode, int min_qubits) {
    return guarded([&] {
        require(mode >= 0 && mode <= 2, "mode must be 0 (off), 1 (auto) or 2 (always)");
        b200::jit_set_mode(static_cast<b200::JitMode>(mode), min_qubits);
    });
}

qsim_status_t qsim_jit_set_dual(int mode, int min_fp64) {
    return guarded([&] {
        require(mode >= -1 && mode <= 2, "mode must be -1 (keep), 0 (off), 1 (auto) or 2 (always)");
        b200::jit_set_dual(mode, min_fp64);
    });
}

// Queues the background compile of one pass (or finds its kernel ready) without a device: *state = 0 ready, 1 pending,
// 2 unavailable.  The GPU-less check of the background-compilation machinery (tests/test_jit_cpu.py).
qsim_status_t qsim_program_jit_request(const qsim_program_t* p, int pass, int* state) {
    return guarded([&] {
        require(p != nullptr && state != nullptr && pass >= 0 && pass < (int)p->dev.host.passes.size(), "bad argument");
        const b200::PassDesc& pd = p->dev.host.passes[pass];
        const auto rq = b200::jit_make_request(pd, p->dev.host.ops.data() + pd.op_offset);
        bool pending = false;
        const auto k = b200::jit_lookup(*rq, /*needs_device=*/false, /*async=*/true, &pending);
        *state = k ? 0 : (pending ? 1 : 2);
    });
}

qsim_status_t qsim_jit_shutdown(void) {
    return guarded([&] { b200::jit_shutdown(); });
}

qsim_status_t qsim_jit_wait(void) {
    return guarded([&] { b200::jit_wait_all(); });
}

qsim_status_t qsim_jit_stats(int64_t out[8]) {
    return guarded([&] {
        require(out != nullptr, "null output");
        const b200::JitStats st = b200::jit_stats();
        out[0] = st.compiles; out[1] = st.cache_hits; out[2] = st.launches; out[3] = st.failures;
        out[4] = (int64_t)(st.compile_seconds * 1e6); out[5] = st.disk_hits;
        out[6] = (int64_t)b200::jit_mode(); out[7] = b200::jit_min_qubits();
    });
}

qsim_status_t qsim_program_set_specialised(qsim_program_t* p, int on) {
    return guarded([&] {
        require(p != nullptr, "null program");
        p->dev.host.force_jit = on != 0;
        if (on) std::fill(p->dev.host.jit_tried.begin(), p->dev.host.jit_tried.end(), 0);
        if (p->dev.graph.exec) { cudaGraphExecDestroy(p->dev.graph.exec); p->dev.graph = b200::DeviceProgram::Graph(); }
    });
}

size_t qsim_program_jit_source(const qsim_program_t* p, int pass, int whole_unit, char* buf, size_t cap) {
    if (!p || pass < 0 || pass >= (int)p->dev.host.passes.size()) return 0;
    const b200::PassDesc& pd = p->dev.host.passes[pass];
    const b200::DevOp* ops = p->dev.host.ops.data() + pd.op_offset;
    const bool dual = b200::jit_dual_wanted(pd, ops);   // (QSIM_DUAL=always / off to look at either build)
    return copy_out(whole_unit ? b200::jit_translation_unit(pd, ops, dual) : b200::jit_generate_compute(pd, ops, dual), buf, cap);
}

qsim_status_t qsim_program_jit_compile(const qsim_program_t* p, int pass, int64_t* cubin_bytes, void* cubin_out, size_t cap) {
    return guarded([&] {
        require(p != nullptr && pass >= 0 && pass < (int)p->dev.host.passes.size(), "no such pass");
        const b200::PassDesc& pd = p->dev.host.passes[pass];
        const b200::JitMode saved = b200::jit_mode();
        const int saved_min = b200::jit_min_qubits();
        b200::jit_set_mode(b200::JitMode::Always, saved_min);   // a failure throws with the compiler's log
        std::shared_ptr<b200::JitKernel> k;
        const b200::DevOp* ops = p->dev.host.ops.data() + pd.op_offset;
        try { k = b200::jit_get_kernel(pd, ops, /*needs_device=*/false, b200::jit_dual_wanted(pd, ops)); }
        catch (...) { b200::jit_set_mode(saved, saved_min); throw; }
        b200::jit_set_mode(saved, saved_min);
        if (cubin_bytes) *cubin_bytes = k ? (int64_t)b200::jit_copy_cubin(*k, nullptr, 0) : 0;
        if (cubin_out && k) b200::jit_copy_cubin(*k, cubin_out, cap);
    });
}

size_t qsim_program_describe(const qsim_program_t* p, char* buf, size_t cap) {
    if (!p) return 0;
    return copy_out(p->dev.host.describe(), buf, cap);
}

// ---- simulator --------------------------------------------------------------------------------

qsim_status_t qsim_sim_create(int n, qsim_sim_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        auto s = std::make_unique<qsim_sim>();
        s->sim = std::make_unique<Simulator>(n);
        s->n_total = n;
        *out = s.release();
    });
}

qsim_status_t qsim_sim_create_external(int n, void* device_state, qsim_sim_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        auto s = std::make_unique<qsim_sim>();
        s->sim = std::make_unique<Simulator>(n, static_cast<cuDoubleComplex*>(device_state));
        s->n_total = n;
        *out = s.release();
    });
}

qsim_status_t qsim_shard_create(int n, int n_global, int rank, void* device_state, qsim_sim_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        require(isValidQubitCount(n), "Number of qubits out of range");
        require(n_global >= 0 && n_global < n, "n_global out of range");
        require(rank >= 0 && rank < (1 << n_global), "rank out of range");
        auto s = std::make_unique<qsim_sim>();
        const int nl = n - n_global;
        if (device_state) s->sim = std::make_unique<Simulator>(nl, static_cast<cuDoubleComplex*>(device_state));
        else s->sim = std::make_unique<Simulator>(nl);
        s->n_total = n;
        s->n_global = n_global;
        s->rank = rank;
        // |0...0> lives on rank 0 only.  Shards defer the write even on caller memory: every access goes through this
        // object, and ranks that are about to be read by a peer materialise first (qsim_sim_device_ptr).
        if (n_global > 0) s->sim->state().allowLazyExternal(true);
        if (rank != 0) s->sim->state().initializeAllZero();
        else if (device_state) s->sim->reset();
        *out = s.release();
    });
}

void qsim_sim_destroy(qsim_sim_t* s) { delete s; }

qsim_status_t qsim_sim_set_stream(qsim_sim_t* s, void* stream) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        s->sim->state().engine().setStream(static_cast<cudaStream_t>(stream));
    });
}

qsim_status_t qsim_sim_reset(qsim_sim_t* s) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        if (s->n_global && s->rank != 0) s->sim->state().initializeAllZero();
        else s->sim->reset();
    });
}

qsim_status_t qsim_sim_init_basis(qsim_sim_t* s, uint64_t idx) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->state().initializeBasis(idx); });
}

qsim_status_t qsim_sim_set_state(qsim_sim_t* s, const double* amps) {
    return guarded([&] {
        require(s != nullptr && amps != nullptr, "null argument");
        s->sim->state().setFromHost(reinterpret_cast<const std::complex<double>*>(amps));
    });
}

qsim_status_t qsim_sim_run(qsim_sim_t* s, int circuit_qubits, const qsim_gate_t* gates, int64_t ng) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        Circuit c = build(circuit_qubits, gates, ng);
        if (s->n_global == 0) { s->sim->run(c); return; }
        if (circuit_qubits != s->n_total) throw std::invalid_argument("Circuit qubit count doesn't match simulator");
        b200::Program prog;
        b200::CompileOptions opt = b200::default_options();
        opt.n_global = s->n_global;
        // no X frame on a shard through this entry point: an X on a rank qubit would be absorbed into the frame and left in
        // Program::global_xor, which only a sharded driver (qsim_program_compile_ex) carries between calls; without the
        // frame it is rejected by the compiler ("must be remapped") instead of being dropped
        opt.defer_x = false;
        std::string err;
        if (!b200::compile(s->n_total, gates, ng, opt, prog, &err)) throw std::runtime_error(err);
        s->sim->state().engine().execute(prog, s->sim->state().devicePtr(), s->hi_bits());
    });
}

qsim_status_t qsim_sim_apply_gate(qsim_sim_t* s, const qsim_gate_t* g) {
    return guarded([&] {
        require(s != nullptr && g != nullptr, "null argument");
        Circuit c = build(s->n_total, g, 1);
        if (s->n_global == 0) s->sim->applyGate(c.getGates()[0]);
        else {
            b200::Program prog;
            b200::CompileOptions opt = b200::default_options();
            opt.n_global = s->n_global;
            opt.defer_x = false;   // see qsim_sim_run
            std::string err;
            if (!b200::compile(s->n_total, g, 1, opt, prog, &err)) throw std::runtime_error(err);
            s->sim->state().engine().execute(prog, s->sim->state().devicePtr(), s->hi_bits());
        }
    });
}

qsim_status_t qsim_sim_execute(qsim_sim_t* s, const qsim_program_t* p) {
    return guarded([&] {
        require(s != nullptr && p != nullptr, "null argument");
        if (p->dev.host.n != s->n_total || p->n_global != s->n_global)
            throw std::invalid_argument("Circuit qubit count doesn't match simulator");
        if (!p->dev.d_ops && !p->dev.host.ops.empty()) const_cast<qsim_program_t*>(p)->dev.upload();   // compiled without a device
        if (s->n_global == 0) s->sim->execute(p->dev);
        else {
            StateVector& sv = s->sim->state();
            uint64_t basis = 0;
            if (!p->dev.host.passes.empty() && sv.takePendingBasis(&basis))   // first pass generates the shard on chip
                sv.engine().execute(p->dev, sv.rawDevicePtr(), s->hi_bits(), (int64_t)basis);
            else
                sv.engine().execute(p->dev, sv.devicePtr(), s->hi_bits());
        }
    });
}

qsim_status_t qsim_sim_synchronize(qsim_sim_t* s) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->synchronize(); });
}

qsim_status_t qsim_sim_get_state(const qsim_sim_t* s, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        s->sim->state().toHost(reinterpret_cast<std::complex<double>*>(out));
    });
}

qsim_status_t qsim_sim_get_probabilities(const qsim_sim_t* s, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        s->sim->state().getProbabilities(out, 0, s->sim->getStateSize());
    });
}

qsim_status_t qsim_sim_get_probability_range(const qsim_sim_t* s, uint64_t first, uint64_t count, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        s->sim->state().getProbabilities(out, first, count);
    });
}

qsim_status_t qsim_sim_total_probability(const qsim_sim_t* s, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        *out = s->sim->state().getTotalProbability();
    });
}

qsim_status_t qsim_sim_marginal(const qsim_sim_t* s, const int* qubits, int k, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr && (qubits != nullptr || k == 0), "null argument");
        require(k >= 0 && k <= 12, "marginal over 0..12 qubits");
        const std::vector<int> bits(qubits, qubits + k);
        const std::vector<double> m = s->sim->state().marginalProbabilities(bits);
        std::memcpy(out, m.data(), m.size() * sizeof(double));
    });
}

qsim_status_t qsim_sim_sample_uniforms(qsim_sim_t* s, const double* u, int64_t shots, int64_t* out) {
    return guarded([&] {
        require(s != nullptr && u != nullptr && out != nullptr, "null argument");
        auto v = s->sim->state().sampleWithUniforms(u, shots);
        std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    });
}

qsim_status_t qsim_sim_sample_seeded(qsim_sim_t* s, unsigned seed, int64_t shots, int64_t* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        auto v = s->sim->state().sampleSeeded(seed, shots);
        std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    });
}

qsim_status_t qsim_sim_measure(qsim_sim_t* s, int qubit, double r, int* outcome) {
    return guarded([&] {
        require(s != nullptr && outcome != nullptr, "null argument");
        *outcome = s->sim->state().measure(qubit, r);
    });
}

qsim_status_t qsim_sim_measure_bit(qsim_sim_t* s, int bit, double r, int* outcome, double* p0) {
    return guarded([&] {
        require(s != nullptr && outcome != nullptr, "null argument");
        *outcome = s->sim->state().measureBit(bit, r, p0);
    });
}

int qsim_sim_num_qubits(const qsim_sim_t* s) { return s ? s->n_total : 0; }
void* qsim_sim_device_ptr(qsim_sim_t* s) { return s ? s->sim->state().devicePtr() : nullptr; }
int64_t qsim_sim_launch_count(const qsim_sim_t* s) { return s ? s->sim->state().engine().launches() : 0; }

qsim_status_t qsim_sim_set_timing(qsim_sim_t* s, int enabled) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->state().engine().setTiming(enabled != 0); });
}

qsim_status_t qsim_sim_pass_time_ms(qsim_sim_t* s, double* total_ms, int64_t* n_passes) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        s->sim->state().engine().drainTiming(total_ms, n_passes);
    });
}

qsim_status_t qsim_sim_pass_times(qsim_sim_t* s, double* out_ms, int64_t cap, int64_t* n_out) {
    return guarded([&] {
        require(s != nullptr && n_out != nullptr, "null argument");
        std::vector<double> each;
        s->sim->state().engine().drainTiming(nullptr, nullptr, &each);
        *n_out = (int64_t)each.size();
        for (int64_t i = 0; i < (int64_t)each.size() && i < cap; ++i) out_ms[i] = each[i];
    });
}

qsim_status_t qsim_sim_pass_timeline(qsim_sim_t* s, uint64_t* out, int64_t cap, int64_t* n_out) {
    return guarded([&] {
        require(s != nullptr && n_out != nullptr, "null argument");
        const auto t = s->sim->state().engine().passTimeline();
        *n_out = (int64_t)t.size();
        for (int64_t i = 0; i < (int64_t)t.size() && i < cap; ++i) out[i] = t[i];
    });
}

// ---- shards -----------------------------------------------------------------------------------

qsim_status_t qsim_shard_swap_p2p(qsim_sim_t* s, void* peer_state, int global_qubit, int local_qubit) {
    return guarded([&] {
        require(s != nullptr && peer_state != nullptr, "null argument");
        const int nl = s->n_total - s->n_global;
        require(global_qubit >= nl && global_qubit < s->n_total, "global_qubit is not a global qubit");
        require(local_qubit >= 0 && local_qubit < nl, "local_qubit is not a local qubit");
        const int my_bit = (s->rank >> (global_qubit - nl)) & 1;
        auto& eng = s->sim->state().engine();
        b200::launch_swap_p2p(s->sim->state().devicePtr(), static_cast<cuDoubleComplex*>(peer_state), nl, local_qubit,
                              my_bit, eng.numSMs(), eng.stream());
        eng.countLaunch();
    });
}

qsim_status_t qsim_program_last_tile_mask(const qsim_program_t* p, uint64_t* mask_out) {
    return guarded([&] {
        require(p != nullptr && mask_out != nullptr, "null argument");
        uint64_t m = 0;
        if (!p->dev.host.passes.empty()) {
            const b200::PassDesc& pd = p->dev.host.passes.back();
            for (int j = 0; j < pd.t; ++j) m |= 1ULL << pd.tile_bits[j];
        }
        *mask_out = m;
    });
}

qsim_status_t qsim_shard_execute_exchange(qsim_sim_t* s, const qsim_program_t* p, void* alt_state, void* peer_alt_state,
                                          int global_qubit, int local_qubit) {
    return guarded([&] {
        require(s != nullptr && p != nullptr && alt_state != nullptr && peer_alt_state != nullptr, "null argument");
        if (p->dev.host.n != s->n_total || p->n_global != s->n_global)
            throw std::invalid_argument("Circuit qubit count doesn't match simulator");
        const int nl = s->n_total - s->n_global;
        require(global_qubit >= nl && global_qubit < s->n_total, "global_qubit is not a global qubit");
        require(local_qubit >= 0 && local_qubit < nl, "local_qubit is not a local qubit");
        require(!p->dev.host.passes.empty(), "the program has no pass to fuse the exchange into");
        const b200::PassDesc& pd = p->dev.host.passes.back();
        for (int j = 0; j < pd.t; ++j)
            require(pd.tile_bits[j] != local_qubit, "local_qubit is a tile qubit of the program's last pass");
        b200::StoreRedirect rd;
        rd.keep = static_cast<cuDoubleComplex*>(alt_state);
        rd.send = static_cast<cuDoubleComplex*>(peer_alt_state);
        rd.bit = local_qubit;
        rd.keep_value = (s->rank >> (global_qubit - nl)) & 1;
        StateVector& sv = s->sim->state();
        sv.engine().execute(p->dev, sv.devicePtr(), s->hi_bits(), -1, &rd);
        sv.rebindExternal(rd.keep);
    });
}

// Can this pass carry the exchange of local qubit v in place (w < 0), or one half of it split by index bit w?  (the launch-time
// conditions of launch_pass)
static bool pass_can_exchange_in_place(const b200::PassDesc& pd, int v, int w, int num_sms) {
    uint64_t vw = 1ULL << v;
    if (w >= 0) {
        if (w == v || w >= pd.n) return false;
        vw |= 1ULL << w;
    }
    if ((pd.tile_mask | pd.xdep) & vw) return false;   // a tile qubit, or deferred X gates pair tiles across it
    const uint64_t n_tiles = 1ULL << (pd.n - pd.t);
    const uint64_t grid = n_tiles < (uint64_t)num_sms ? n_tiles : (uint64_t)num_sms;
    return w >= 0 ? (grid >= 32 && n_tiles >= 64) : grid >= 16;
}

// Can this pass GATHER half of the exchange of local qubit v (split by index bit w)?  v must be its highest tile qubit, moved
// by TMA instructions of its own (qsim_program_compile_ex2's isolate hint); w a qubit outside the tile.
static bool pass_can_gather_half(const b200::PassDesc& pd, int v, int w, int num_sms) {
    if (w < 0 || w >= pd.n || w == v) return false;
    if (pd.t < 2 || pd.tma_instr_bits < 1 || (int)pd.tile_bits[pd.t - 1] != v) return false;
    if ((pd.tile_mask | pd.xdep) & (1ULL << w)) return false;
    const uint64_t n_tiles = 1ULL << (pd.n - pd.t);
    const uint64_t grid = n_tiles < (uint64_t)num_sms ? n_tiles : (uint64_t)num_sms;
    return grid >= 32 && n_tiles >= 64;
}

static bool inplace_exchange_possible(const qsim_program_t* p, int local_qubit, int num_sms) {
    if (p->dev.host.passes.empty()) return false;
    return pass_can_exchange_in_place(p->dev.host.passes.back(), local_qubit, -1, num_sms);
}

qsim_status_t qsim_shard_inplace_exchange_possible(qsim_sim_t* s, const qsim_program_t* p, int local_qubit, int* possible_out) {
    return guarded([&] {
        require(s != nullptr && p != nullptr && possible_out != nullptr, "null argument");
        *possible_out = inplace_exchange_possible(p, local_qubit, s->sim->state().engine().numSMs()) ? 1 : 0;
    });
}

qsim_status_t qsim_shard_execute_exchange_inplace(qsim_sim_t* s, const qsim_program_t* p, void* peer_state, int global_qubit,
                                                  int local_qubit, void* hs_local, void* hs_peer, uint64_t hs_base,
                                                  uint64_t timeout_ns, int* hs_error_dev) {
    return guarded([&] {
        require(s != nullptr && p != nullptr && peer_state != nullptr && hs_local != nullptr && hs_peer != nullptr &&
                hs_error_dev != nullptr, "null argument");
        if (p->dev.host.n != s->n_total || p->n_global != s->n_global)
            throw std::invalid_argument("Circuit qubit count doesn't match simulator");
        const int nl = s->n_total - s->n_global;
        require(global_qubit >= nl && global_qubit < s->n_total, "global_qubit is not a global qubit");
        require(local_qubit >= 0 && local_qubit < nl, "local_qubit is not a local qubit");
        StateVector& sv = s->sim->state();
        require(inplace_exchange_possible(p, local_qubit, sv.engine().numSMs()),
                "the program's last pass cannot carry this exchange in place (qsim_shard_inplace_exchange_possible)");
        b200::StoreRedirect rd;
        rd.keep = sv.devicePtr();   // (a lazily reset shard is written out here)
        rd.send = static_cast<cuDoubleComplex*>(peer_state);
        rd.bit = local_qubit;
        rd.keep_value = (s->rank >> (global_qubit - nl)) & 1;
        rd.in_place = true;
        rd.hs_local = static_cast<unsigned long long*>(hs_local);
        rd.hs_peer = static_cast<unsigned long long*>(hs_peer);
        rd.hs_base = hs_base;
        rd.hs_timeout_ns = timeout_ns ? timeout_ns : 10000000000ULL;
        rd.hs_error = hs_error_dev;
        sv.engine().execute(p->dev, rd.keep, s->hi_bits(), -1, &rd);
    });
}

qsim_status_t qsim_shard_split_exchange_possible(qsim_sim_t* s, const qsim_program_t* before, const qsim_program_t* after,
                                                 int local_qubit, int* split_bit_out) {
    return guarded([&] {
        require(s != nullptr && before != nullptr && after != nullptr && split_bit_out != nullptr, "null argument");
        *split_bit_out = -1;
        if (before->dev.host.passes.empty() || after->dev.host.passes.empty()) return;
        const int num_sms = s->sim->state().engine().numSMs();
        const b200::PassDesc& pa = before->dev.host.passes.back();
        const b200::PassDesc& pb = after->dev.host.passes.front();
        for (int w = pa.n - 1; w >= 0; --w)
            if (pass_can_exchange_in_place(pa, local_qubit, w, num_sms) && pass_can_gather_half(pb, local_qubit, w, num_sms)) {
                *split_bit_out = w;
                return;
            }
    });
}

qsim_status_t qsim_shard_execute_exchange_half(qsim_sim_t* s, const qsim_program_t* p, void* peer_state, int global_qubit,
                                               int local_qubit, int split_bit, int which, void* hs_local, void* hs_peer,
                                               uint64_t hs_base, uint64_t timeout_ns, int* hs_error_dev) {
    return guarded([&] {
        require(s != nullptr && p != nullptr && peer_state != nullptr && hs_local != nullptr && hs_peer != nullptr &&
                hs_error_dev != nullptr, "null argument");
        require(which == 1 || which == 2, "which must be 1 (scatter: the program's last pass) or 2 (gather: its first pass)");
        if (p->dev.host.n != s->n_total || p->n_global != s->n_global)
            throw std::invalid_argument("Circuit qubit count doesn't match simulator");
        const int nl = s->n_total - s->n_global;
        require(global_qubit >= nl && global_qubit < s->n_total, "global_qubit is not a global qubit");
        require(local_qubit >= 0 && local_qubit < nl, "local_qubit is not a local qubit");
        require(!p->dev.host.passes.empty(), "the program has no pass");
        StateVector& sv = s->sim->state();
        const b200::PassDesc& pd = which == 1 ? p->dev.host.passes.back() : p->dev.host.passes.front();
        require(which == 1 ? pass_can_exchange_in_place(pd, local_qubit, split_bit, sv.engine().numSMs())
                           : pass_can_gather_half(pd, local_qubit, split_bit, sv.engine().numSMs()),
                "that pass cannot carry half of this exchange (qsim_shard_split_exchange_possible)");
        b200::StoreRedirect rd;
        rd.keep = sv.devicePtr();   // (a lazily reset shard is written out here)
        rd.send = static_cast<cuDoubleComplex*>(peer_state);
        rd.bit = local_qubit;
        rd.keep_value = (s->rank >> (global_qubit - nl)) & 1;
        rd.in_place = true;
        rd.split = which;
        rd.split_bit = split_bit;
        rd.hs_local = static_cast<unsigned long long*>(hs_local);
        rd.hs_peer = static_cast<unsigned long long*>(hs_peer);
        rd.hs_base = hs_base;
        rd.hs_timeout_ns = timeout_ns ? timeout_ns : 10000000000ULL;
        rd.hs_error = hs_error_dev;
        sv.engine().execute(p->dev, rd.keep, s->hi_bits(), -1, &rd);
    });
}

static void chunk_range(int nl, int64_t chunk, int64_t n_chunks, uint64_t* jb, uint64_t* cnt) {
    const uint64_t pairs = 1ULL << (nl - 1);
    const uint64_t per = (pairs + (uint64_t)n_chunks - 1) / (uint64_t)n_chunks;
    *jb = per * (uint64_t)chunk;
    *cnt = *jb >= pairs ? 0 : (pairs - *jb < per ? pairs - *jb : per);
}

qsim_status_t qsim_shard_pack_half(qsim_sim_t* s, int local_qubit, int keep_bit, int64_t chunk, int64_t n_chunks,
                                   void* buf) {
    return guarded([&] {
        require(s != nullptr && buf != nullptr, "null argument");
        const int nl = s->n_total - s->n_global;
        require(local_qubit >= 0 && local_qubit < nl && n_chunks > 0 && chunk >= 0 && chunk < n_chunks, "bad argument");
        uint64_t jb, cnt;
        chunk_range(nl, chunk, n_chunks, &jb, &cnt);
        auto& eng = s->sim->state().engine();
        if (cnt) {
            b200::launch_pack_half(s->sim->state().devicePtr(), static_cast<cuDoubleComplex*>(buf), local_qubit,
                                   keep_bit ^ 1, jb, cnt, eng.numSMs(), eng.stream());
            eng.countLaunch();
        }
    });
}

qsim_status_t qsim_shard_unpack_half(qsim_sim_t* s, int local_qubit, int keep_bit, int64_t chunk, int64_t n_chunks,
                                     const void* buf) {
    return guarded([&] {
        require(s != nullptr && buf != nullptr, "null argument");
        const int nl = s->n_total - s->n_global;
        require(local_qubit >= 0 && local_qubit < nl && n_chunks > 0 && chunk >= 0 && chunk < n_chunks, "bad argument");
        uint64_t jb, cnt;
        chunk_range(nl, chunk, n_chunks, &jb, &cnt);
        auto& eng = s->sim->state().engine();
        if (cnt) {
            b200::launch_unpack_half(s->sim->state().devicePtr(), static_cast<const cuDoubleComplex*>(buf), local_qubit,
                                     keep_bit ^ 1, jb, cnt, eng.numSMs(), eng.stream());
            eng.countLaunch();
        }
    });
}

qsim_status_t qsim_ipc_get_handle(void* device_ptr, unsigned char handle_out[64], uint64_t* offset_out) {
    return guarded([&] {
        require(device_ptr != nullptr && handle_out != nullptr, "null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
        b200::require_device();
        // The driver API is reached through the runtime (no link-time dependency on libcuda.so.1, so the
        // library still loads on a GPU-less build box).
        using range_fn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) throw std::runtime_error("cuMemGetAddressRange unavailable");
        CUdeviceptr base = 0;
        size_t size = 0;
        if (reinterpret_cast<range_fn>(fn)(&base, &size, reinterpret_cast<CUdeviceptr>(device_ptr)) != CUDA_SUCCESS)
            throw std::runtime_error("cuMemGetAddressRange failed");
        cudaIpcMemHandle_t h;
        CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
        std::memcpy(handle_out, &h, 64);
        if (offset_out) *offset_out = (uint64_t)(reinterpret_cast<CUdeviceptr>(device_ptr) - base);
    });
}

qsim_status_t qsim_ipc_open_handle(const unsigned char handle[64], void** base_ptr_out) {
    return guarded([&] {
        require(handle != nullptr && base_ptr_out != nullptr, "null argument");
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle, 64);
        CUDA_CHECK(cudaIpcOpenMemHandle(base_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    });
}

qsim_status_t qsim_ipc_close_handle(void* base_ptr) {
    return guarded([&] { CUDA_CHECK(cudaIpcCloseMemHandle(base_ptr)); });
}

qsim_status_t qsim_shard_partial_probability(const qsim_sim_t* s, int bit, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        *out = s->sim->state().partialProbability(bit);
    });
}

qsim_status_t qsim_shard_collapse(qsim_sim_t* s, int bit, int outcome, double scale) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        s->sim->state().collapse(bit, outcome, scale);
    });
}

qsim_status_t qsim_shard_sample(qsim_sim_t* s, double c_init, int first_shard, const double* uniforms, int64_t n_shots,
                                int64_t* out, double* c_end) {
    return guarded([&] {
        require(s != nullptr && uniforms != nullptr && out != nullptr && c_end != nullptr, "null argument");
        *c_end = s->sim->state().sampleShard(c_init, first_shard != 0, uniforms, n_shots, out);
    });
}

qsim_status_t qsim_shard_cdf_prepare(qsim_sim_t* s, double* approx_total) {
    return guarded([&] {
        require(s != nullptr && approx_total != nullptr, "null argument");
        *approx_total = s->sim->state().sampleShardPrepare();
    });
}

qsim_status_t qsim_shard_cdf_classify(qsim_sim_t* s, double approx_c_init) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        s->sim->state().sampleShardClassify(approx_c_init);
    });
}

// ---- sharded simulator (C++ driver over NCCL, include/qsim/sharded_simulator.hpp) -----------------------------------------

struct qsim_sharded { std::unique_ptr<ShardedSimulator> sim; };
struct qsim_sharded_plan { std::shared_ptr<ShardedSimulator::CompiledPlan> plan; };

qsim_status_t qsim_sharded_unique_id(unsigned char out[128]) {
    return guarded([&] {
        require(out != nullptr, "null output");
        const auto id = ShardedSimulator::createUniqueId();
        std::memcpy(out, id.data(), id.size());
    });
}

qsim_status_t qsim_sharded_create(int n, int rank, int world, const unsigned char unique_id[128], int exchange, qsim_sharded_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        require(isValidQubitCount(n), "Number of qubits out of range");
        require(exchange >= 0 && exchange <= 2, "exchange must be 0 (auto), 1 (peer memory) or 2 (nccl)");
        auto h = std::make_unique<qsim_sharded>();
        h->sim = std::make_unique<ShardedSimulator>(n, rank, world, unique_id, static_cast<ShardedSimulator::Exchange>(exchange));
        *out = h.release();
    });
}

void qsim_sharded_destroy(qsim_sharded_t* h) { delete h; }

qsim_status_t qsim_sharded_reset(qsim_sharded_t* h) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->reset(); });
}

qsim_status_t qsim_sharded_run(qsim_sharded_t* h, int cq, const qsim_gate_t* gates, int64_t ng) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->run(build(cq, gates, ng)); });
}

qsim_status_t qsim_sharded_compile(qsim_sharded_t* h, int cq, const qsim_gate_t* gates, int64_t ng, int k, qsim_sharded_plan_t** out) {
    return guarded([&] {
        require(h != nullptr && out != nullptr && k >= 1, "bad argument");
        const Circuit c = build(cq, gates, ng);
        const auto plans = k == 1 ? std::vector<std::shared_ptr<ShardedSimulator::CompiledPlan>>{h->sim->compile(c)}
                                  : h->sim->compileSequence(c, k);
        for (int i = 0; i < k; ++i) out[i] = new qsim_sharded_plan{plans[(size_t)i]};
    });
}

void qsim_sharded_plan_destroy(qsim_sharded_plan_t* p) { delete p; }

qsim_status_t qsim_sharded_plan_info(const qsim_sharded_plan_t* p, int64_t info[8]) {
    return guarded([&] {
        require(p != nullptr && info != nullptr, "null argument");
        std::memset(info, 0, 8 * sizeof(int64_t));
        info[0] = ShardedSimulator::planPasses(*p->plan);
        info[1] = ShardedSimulator::planOps(*p->plan);
        info[2] = ShardedSimulator::planSwaps(*p->plan);
    });
}

qsim_status_t qsim_sharded_execute(qsim_sharded_t* h, const qsim_sharded_plan_t* p) {
    return guarded([&] { require(h != nullptr && p != nullptr, "null argument"); h->sim->execute(*p->plan); });
}

qsim_status_t qsim_sharded_sample(qsim_sharded_t* h, const double* uniforms, int64_t n_shots, int64_t* out) {
    return guarded([&] {
        require(h != nullptr && (n_shots == 0 || (uniforms != nullptr && out != nullptr)) && n_shots >= 0, "bad argument");
        const auto res = h->sim->sample(std::vector<double>(uniforms, uniforms + n_shots));
        std::copy(res.begin(), res.end(), out);
    });
}

qsim_status_t qsim_sharded_measure(qsim_sharded_t* h, int qubit, double uniform, int* result) {
    return guarded([&] { require(h != nullptr && result != nullptr, "null argument"); *result = h->sim->measureQubit(qubit, uniform); });
}

qsim_status_t qsim_sharded_measure_bit(qsim_sharded_t* h, int bit, double uniform, int* result, double* p0) {
    return guarded([&] { require(h != nullptr && result != nullptr, "null argument"); *result = h->sim->measureBit(bit, uniform, p0); });
}

qsim_status_t qsim_sharded_marginal(qsim_sharded_t* h, const int* qubits, int k, double* out) {
    return guarded([&] {
        require(h != nullptr && out != nullptr && k >= 0 && (k == 0 || qubits != nullptr), "bad argument");
        const auto m = h->sim->getMarginalProbabilities(std::vector<int>(qubits, qubits + k));
        std::copy(m.begin(), m.end(), out);
    });
}

qsim_status_t qsim_sharded_total_probability(qsim_sharded_t* h, double* out) {
    return guarded([&] { require(h != nullptr && out != nullptr, "null argument"); *out = h->sim->getTotalProbability(); });
}

qsim_status_t qsim_sharded_get_local_state(qsim_sharded_t* h, double* out) {
    return guarded([&] {
        require(h != nullptr && out != nullptr, "null argument");
        const auto st = h->sim->getLocalState();
        std::memcpy(out, st.data(), st.size() * sizeof(std::complex<double>));
    });
}

qsim_status_t qsim_sharded_set_local_state(qsim_sharded_t* h, const double* amps) {
    return guarded([&] {
        require(h != nullptr && amps != nullptr, "null argument");
        h->sim->setLocalState(reinterpret_cast<const std::complex<double>*>(amps));
    });
}

qsim_status_t qsim_sharded_restore_identity_layout(qsim_sharded_t* h) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->restoreIdentityLayout(); });
}

qsim_status_t qsim_sharded_layout(const qsim_sharded_t* h, int* perm_out, uint64_t* frame_out) {
    return guarded([&] {
        require(h != nullptr, "null simulator");
        if (perm_out) std::copy(h->sim->permutation().begin(), h->sim->permutation().end(), perm_out);
        if (frame_out) *frame_out = h->sim->frame();
    });
}

qsim_status_t qsim_sharded_set_identity_layout_only(qsim_sharded_t* h, int on) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->setIdentityLayoutOnly(on != 0); });
}

qsim_status_t qsim_sharded_relabel_identity(qsim_sharded_t* h) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->relabelIdentity(); });
}

qsim_status_t qsim_sharded_swap(qsim_sharded_t* h, int global_position, int local_position) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->swapQubits(global_position, local_position); });
}

qsim_status_t qsim_sharded_info(const qsim_sharded_t* h, int64_t info[8]) {
    return guarded([&] {
        require(h != nullptr && info != nullptr, "null argument");
        std::memset(info, 0, 8 * sizeof(int64_t));
        info[0] = h->sim->localQubits();
        info[1] = h->sim->getNumQubits() - h->sim->localQubits();
        info[2] = h->sim->fusedExchanges();
        info[3] = h->sim->separateExchanges();
        const std::string ex = h->sim->exchangeName();
        info[4] = ex == "p2p" ? 1 : (ex == "nccl" ? 2 : 0);
        info[5] = h->sim->rank();
        info[6] = h->sim->worldSize();
        info[7] = h->sim->hasSecondBuffer() ? 1 : 0;
    });
}

qsim_status_t qsim_sharded_inplace_exchanges(const qsim_sharded_t* h, int64_t* count_out) {
    return guarded([&] {
        require(h != nullptr && count_out != nullptr, "null argument");
        *count_out = h->sim->inPlaceExchanges();
    });
}

qsim_status_t qsim_sharded_split_exchanges(const qsim_sharded_t* h, int64_t* count_out) {
    return guarded([&] {
        require(h != nullptr && count_out != nullptr, "null argument");
        *count_out = h->sim->splitExchanges();
    });
}

qsim_sim_t* qsim_sharded_local(qsim_sharded_t* h) { return h ? static_cast<qsim_sim_t*>(h->sim->localHandle()) : nullptr; }

qsim_status_t qsim_sharded_set_stream(qsim_sharded_t* h, void* stream) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->setStream(static_cast<cudaStream_t>(stream)); });
}

qsim_status_t qsim_sharded_synchronize(qsim_sharded_t* h) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->synchronize(); });
}

qsim_status_t qsim_sharded_barrier(qsim_sharded_t* h) {
    return guarded([&] { require(h != nullptr, "null simulator"); h->sim->barrier(); });
}

qsim_status_t qsim_sharded_plan_circuit(int n, int n_global, const qsim_gate_t* gates, int64_t ng, const int* perm_in,
                                        int choose_layout, int64_t* steps_out, int64_t cap_steps, qsim_gate_t* gates_out,
                                        int* perm_start_out, int* perm_end_out, int64_t* n_steps_out) {
    return guarded([&] {
        require(n_steps_out != nullptr && n_global >= 0 && n_global < n, "bad argument");
        build(n, gates, ng);   // validation only
        std::vector<int> perm(n);
        for (int q = 0; q < n; ++q) perm[q] = perm_in ? perm_in[q] : q;
        if (choose_layout) perm = b200::shard_choose_initial_layout(n, n_global, gates, ng);
        if (perm_start_out) std::copy(perm.begin(), perm.end(), perm_start_out);
        const b200::ShardPlanRec plan = b200::shard_plan_circuit(n, n_global, gates, ng, perm);
        *n_steps_out = (int64_t)plan.steps.size();
        int64_t gi = 0;
        for (size_t i = 0; i < plan.steps.size(); ++i) {
            const auto& st = plan.steps[i];
            if (steps_out && (int64_t)i < cap_steps) {
                steps_out[3 * i] = st.is_swap ? 1 : 0;
                steps_out[3 * i + 1] = st.is_swap ? st.global_qubit : (int64_t)st.gates.size();
                steps_out[3 * i + 2] = st.is_swap ? st.local_qubit : 0;
            }
            if (gates_out)
                for (const auto& r : st.gates) gates_out[gi++] = r;
        }
        if (perm_end_out) std::copy(plan.perm.begin(), plan.perm.end(), perm_end_out);
    });
}

// ---- noisy ------------------------------------------------------------------------------------

qsim_status_t qsim_noisy_create(int n, const qsim_noise_channel_t* ch, int nch, qsim_noisy_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        auto h = std::make_unique<qsim_noisy>();
        h->sim = std::make_unique<NoisySimulator>(n, to_model(ch, nch));
        *out = h.release();
    });
}
void qsim_noisy_destroy(qsim_noisy_t* s) { delete s; }
qsim_status_t qsim_noisy_set_noise(qsim_noisy_t* s, const qsim_noise_channel_t* ch, int nch) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->setNoiseModel(to_model(ch, nch)); });
}
qsim_status_t qsim_noisy_set_seed(qsim_noisy_t* s, unsigned seed) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->setSeed(seed); });
}
qsim_status_t qsim_noisy_reset(qsim_noisy_t* s) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->reset(); });
}
qsim_status_t qsim_noisy_run(qsim_noisy_t* s, int cq, const qsim_gate_t* gates, int64_t ng) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->run(build(cq, gates, ng)); });
}
qsim_status_t qsim_noisy_apply_gate(qsim_noisy_t* s, const qsim_gate_t* g) {
    return guarded([&] {
        require(s != nullptr && g != nullptr, "null argument");
        s->sim->applyGate(build(s->sim->getNumQubits(), g, 1).getGates()[0]);
    });
}
qsim_status_t qsim_noisy_apply_noise(qsim_noisy_t* s, const qsim_noise_channel_t* ch) {
    return guarded([&] {
        require(s != nullptr && ch != nullptr, "null argument");
        for (const NoiseChannel& c : to_model(ch, 1).getChannels()) s->sim->applyNoise(c);
    });
}
qsim_status_t qsim_noisy_get_state(const qsim_noisy_t* s, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        copy_out(s->sim->getStateVector(), reinterpret_cast<std::complex<double>*>(out));
    });
}
qsim_status_t qsim_noisy_get_probabilities(const qsim_noisy_t* s, double* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); copy_out(s->sim->getProbabilities(), out); });
}
qsim_status_t qsim_noisy_sample(qsim_noisy_t* s, int n_shots, int32_t* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); copy_out(s->sim->sample(n_shots), out); });
}
qsim_status_t qsim_noisy_measure(qsim_noisy_t* s, int qubit, int* outcome) {
    return guarded([&] { require(s != nullptr && outcome != nullptr, "null argument"); *outcome = s->sim->measureQubit(qubit); });
}

// ---- batched ----------------------------------------------------------------------------------

qsim_status_t qsim_batched_create(int n, int batch, const qsim_noise_channel_t* ch, int nch, qsim_batched_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        auto h = std::make_unique<qsim_batched>();
        h->sim = std::make_unique<BatchedSimulator>(n, batch, to_model(ch, nch));
        *out = h.release();
    });
}
void qsim_batched_destroy(qsim_batched_t* s) { delete s; }
qsim_status_t qsim_batched_set_noise(qsim_batched_t* s, const qsim_noise_channel_t* ch, int nch) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->setNoiseModel(to_model(ch, nch)); });
}
qsim_status_t qsim_batched_set_seed(qsim_batched_t* s, unsigned seed) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->setSeed(seed); });
}
qsim_status_t qsim_batched_reset(qsim_batched_t* s) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->reset(); });
}
qsim_status_t qsim_batched_run(qsim_batched_t* s, int cq, const qsim_gate_t* gates, int64_t ng) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->run(build(cq, gates, ng)); });
}
qsim_status_t qsim_batched_average_probabilities(const qsim_batched_t* s, double* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); copy_out(s->sim->getAverageProbabilities(), out); });
}
qsim_status_t qsim_batched_get_probabilities(const qsim_batched_t* s, int traj, double* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); copy_out(s->sim->getProbabilities(traj), out); });
}
qsim_status_t qsim_batched_get_state(const qsim_batched_t* s, int traj, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        copy_out(s->sim->getTrajectoryState(traj), reinterpret_cast<std::complex<double>*>(out));
    });
}
qsim_status_t qsim_batched_sample(qsim_batched_t* s, int n_shots, int32_t* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        const auto v = s->sim->sample(n_shots);
        const size_t b = (size_t)s->sim->getBatchSize();
        for (size_t i = 0; i < v.size(); ++i) std::memcpy(out + i * b, v[i].data(), b * sizeof(int32_t));
    });
}
qsim_status_t qsim_batched_histogram(qsim_batched_t* s, int n_shots, int32_t* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); copy_out(s->sim->getHistogram(n_shots), out); });
}
size_t qsim_batched_total_memory_bytes(const qsim_batched_t* s) { return s ? s->sim->getTotalMemoryBytes() : 0; }

// ---- density matrix -----------------------------------------------------------------------------

qsim_status_t qsim_dm_create(int n, const qsim_noise_channel_t* ch, int nch, qsim_dm_t** out) {
    return guarded([&] {
        require(out != nullptr, "null output");
        auto h = std::make_unique<qsim_dm>();
        h->sim = std::make_unique<DensityMatrixSimulator>(n, to_model(ch, nch));
        *out = h.release();
    });
}
void qsim_dm_destroy(qsim_dm_t* s) { delete s; }
qsim_status_t qsim_dm_reset(qsim_dm_t* s) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->reset(); });
}
qsim_status_t qsim_dm_run(qsim_dm_t* s, int cq, const qsim_gate_t* gates, int64_t ng) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->run(build(cq, gates, ng)); });
}
qsim_status_t qsim_dm_apply_gate(qsim_dm_t* s, const qsim_gate_t* g) {
    return guarded([&] {
        require(s != nullptr && g != nullptr, "null argument");
        s->sim->applyGate(build(s->sim->getNumQubits(), g, 1).getGates()[0]);
    });
}
qsim_status_t qsim_dm_apply_channel(qsim_dm_t* s, int type, int qubit, double p) {
    return guarded([&] {
        require(s != nullptr, "null simulator");
        require(type >= 0 && type <= QSIM_NOISE_BIT_PHASE_FLIP, "unknown noise type");
        s->sim->applyChannel(static_cast<NoiseType>(type), qubit, p);
    });
}
qsim_status_t qsim_dm_init_pure(qsim_dm_t* s, const double* st) {
    return guarded([&] {
        require(s != nullptr && st != nullptr, "null argument");
        const auto* c = reinterpret_cast<const std::complex<double>*>(st);
        s->sim->densityMatrix().initFromPureState(std::vector<std::complex<double>>(c, c + s->sim->densityMatrix().getDimension()));
    });
}
qsim_status_t qsim_dm_init_maximally_mixed(qsim_dm_t* s) {
    return guarded([&] { require(s != nullptr, "null simulator"); s->sim->densityMatrix().initMaximallyMixed(); });
}
qsim_status_t qsim_dm_get_probabilities(const qsim_dm_t* s, double* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); copy_out(s->sim->getProbabilities(), out); });
}
qsim_status_t qsim_dm_get_matrix(const qsim_dm_t* s, double* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        copy_out(s->sim->getDensityMatrix(), reinterpret_cast<std::complex<double>*>(out));
    });
}
qsim_status_t qsim_dm_purity(const qsim_dm_t* s, double* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); *out = s->sim->getPurity(); });
}
qsim_status_t qsim_dm_trace(const qsim_dm_t* s, double* out) {
    return guarded([&] { require(s != nullptr && out != nullptr, "null argument"); *out = s->sim->getTrace(); });
}
qsim_status_t qsim_dm_is_valid(const qsim_dm_t* s, double tol, int* out) {
    return guarded([&] {
        require(s != nullptr && out != nullptr, "null argument");
        *out = const_cast<qsim_dm_t*>(s)->sim->densityMatrix().isValid(tol) ? 1 : 0;
    });
}
qsim_status_t qsim_dm_measure(qsim_dm_t* s, int qubit, double u, int* outcome) {
    return guarded([&] { require(s != nullptr && outcome != nullptr, "null argument"); *outcome = s->sim->measureQubit(qubit, u); });
}

}  // extern "C"
