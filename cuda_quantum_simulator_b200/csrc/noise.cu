// NoisySimulator and BatchedSimulator on the trajectory kernels (kernels_batched.cu) and, for single
// trajectories wider than shared memory, on the fused-pass engine.  See include/qsim/noise_model.cuh for
// the behavioural contract and the deliberate deviations from the reference (src/NoiseModel.cu).
#include "qsim/noise_model.cuh"

#include <cmath>
#include <numeric>
#include <stdexcept>
#include <string>

#include "batched.cuh"
#include "engine.hpp"
#include "program.hpp"
#include "qsim/circuit.hpp"
#include "qsim/constants.hpp"
#include "qsim/state_vector.cuh"
#include "qsim_b200.h"
#include "readout.cuh"

namespace qsim {

namespace {

qsim_gate_t to_record(const GateOp& g) {
    qsim_gate_t r{static_cast<int32_t>(g.type), -1, -1, -1, g.parameter};
    if (g.qubits.size() > 0) r.q0 = g.qubits[0];
    if (g.qubits.size() > 1) r.q1 = g.qubits[1];
    if (g.qubits.size() > 2) r.q2 = g.qubits[2];
    return r;
}

// every (channel, qubit) pair in model order; an empty qubit list means all qubits
std::vector<b200::TrajEvent> flatten(const std::vector<NoiseChannel>& channels, int n) {
    std::vector<b200::TrajEvent> ev;
    for (const NoiseChannel& c : channels) {
        if (c.qubits.empty()) {
            for (int q = 0; q < n; ++q) ev.push_back(b200::make_event(static_cast<int32_t>(c.type), q, c.probability));
        } else {
            for (int q : c.qubits) {
                if (q < 0 || q >= n) throw std::out_of_range("Noise channel qubit out of range");
                ev.push_back(b200::make_event(static_cast<int32_t>(c.type), q, c.probability));
            }
        }
    }
    return ev;
}

// gate list -> trajectory program: the gate's operator(s), then one noise block per gate
std::vector<b200::TrajItem> build_items(const std::vector<qsim_gate_t>& gates, bool with_noise) {
    std::vector<b200::TrajItem> items;
    std::vector<b200::LogicalOp> lops;
    for (size_t i = 0; i < gates.size(); ++i) {
        lops.clear();
        if (!b200::lower_gate(gates[i], lops, (int)i)) throw std::runtime_error("Unknown gate type");
        for (const b200::LogicalOp& op : lops) {
            b200::TrajItem it{};
            it.kind = 0;
            it.target = op.target;
            it.cmask = op.cmask;
            it.cval = op.cval;
            for (int k = 0; k < 8; ++k) it.m[k] = op.m[k];
            items.push_back(it);
        }
        if (with_noise) {
            b200::TrajItem nb{};
            nb.kind = 1;
            items.push_back(nb);
        }
    }
    return items;
}

// host copy of the device RNG (same counter layout as kernels_batched.cu)
void philox_uniforms(uint32_t seed, uint64_t traj, uint64_t event, double& u0, double& u1) {
    uint32_t c0 = (uint32_t)event, c1 = (uint32_t)(event >> 32), c2 = (uint32_t)traj, c3 = (uint32_t)(traj >> 32);
    uint32_t k0 = seed, k1 = 0x51534D42u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u0 = (double)((((uint64_t)c1 << 32) | c0) >> 11) * 0x1.0p-53;
    u1 = (double)((((uint64_t)c3 << 32) | c2) >> 11) * 0x1.0p-53;
}

void run_trajectories(cuDoubleComplex* states, int n, int64_t batch, const std::vector<b200::TrajItem>& items,
                      const std::vector<b200::TrajEvent>& events, uint32_t seed, uint64_t first_block, int num_sms,
                      cudaStream_t stream, double* d_avg = nullptr) {
    if (items.empty()) return;
    CudaMemory<b200::TrajItem> d_items(items.size());
    d_items.copyFromHost(items.data(), items.size());
    CudaMemory<b200::TrajEvent> d_events(events.size() ? events.size() : 1);
    if (!events.empty()) d_events.copyFromHost(events.data(), events.size());
    b200::launch_trajectories(states, n, batch, d_items.get(), (int)items.size(), d_events.get(), (int)events.size(), seed, 0,
                              first_block, num_sms, stream, d_avg);
    CUDA_CHECK(cudaStreamSynchronize(stream));   // the item / event buffers die with this scope
}

void validate_circuit_size(const Circuit& c, int n) {
    if (c.getNumQubits() != n) throw std::invalid_argument("Circuit qubit count doesn't match simulator");
}

}  // namespace

// =====================================================================================================
// NoisySimulator
// =====================================================================================================

NoisySimulator::NoisySimulator(int num_qubits, const NoiseModel& noise_model)
    : num_qubits_(num_qubits), noise_model_(noise_model), state_(std::make_unique<StateVector>(num_qubits)) {
    std::random_device rd;
    setSeed(rd());
}

NoisySimulator::NoisySimulator(int num_qubits) : NoisySimulator(num_qubits, NoiseModel()) {}
NoisySimulator::~NoisySimulator() noexcept = default;
NoisySimulator::NoisySimulator(NoisySimulator&&) noexcept = default;
NoisySimulator& NoisySimulator::operator=(NoisySimulator&&) noexcept = default;

void NoisySimulator::setSeed(unsigned int seed) {
    rng_.seed(seed);
    seed_ = seed;
    noise_block_ = 0;
}

void NoisySimulator::reset() { state_->initializeZero(); }

void NoisySimulator::run(const Circuit& circuit) {
    validate_circuit_size(circuit, num_qubits_);
    std::vector<qsim_gate_t> recs;
    for (const GateOp& g : circuit.getGates()) recs.push_back(to_record(g));
    if (recs.empty()) return;
    b200::Engine& eng = state_->engine();
    if (!noise_model_.hasNoise()) {   // ideal circuit: the whole thing goes through the fused-pass engine
        b200::Program prog;
        std::string err;
        if (!b200::compile(num_qubits_, recs.data(), (int64_t)recs.size(), b200::default_options(), prog, &err))
            throw std::runtime_error(err);
        eng.execute(prog, state_->devicePtr(), 0);
        return;
    }
    const auto events = flatten(noise_model_.getChannels(), num_qubits_);
    if (num_qubits_ <= b200::kTrajMaxQubits) {   // whole noisy circuit in one launch, state resident in shared memory
        const auto items = build_items(recs, true);
        run_trajectories(state_->devicePtr(), num_qubits_, 1, items, events, seed_, noise_block_, eng.numSMs(), eng.stream());
        eng.countLaunch();
        noise_block_ += recs.size();
        return;
    }
    for (const GateOp& g : circuit.getGates()) applyGate(g);
}

void NoisySimulator::applyGate(const GateOp& gate) {
    const qsim_gate_t rec = to_record(gate);
    b200::Program prog;
    std::string err;
    if (!b200::compile(num_qubits_, &rec, 1, b200::default_options(), prog, &err)) throw std::runtime_error(err);
    state_->engine().execute(prog, state_->devicePtr(), 0);
    if (noise_model_.hasNoise()) applyEvents(noise_model_.getChannels());
}

void NoisySimulator::applyNoise(const NoiseChannel& channel) { applyEvents({channel}); }

void NoisySimulator::applyNoiseToQubit(NoiseType type, int qubit, double probability) {
    applyEvents({NoiseChannel(type, {qubit}, probability)});
}

// One noise block (all given channels on their qubits) on the single resident trajectory.
void NoisySimulator::applyEvents(const std::vector<NoiseChannel>& channels) {
    const auto events = flatten(channels, num_qubits_);
    b200::Engine& eng = state_->engine();
    if (num_qubits_ <= b200::kTrajMaxQubits) {
        b200::TrajItem nb{};
        nb.kind = 1;
        run_trajectories(state_->devicePtr(), num_qubits_, 1, {nb}, events, seed_, noise_block_, eng.numSMs(), eng.stream());
        eng.countLaunch();
        ++noise_block_;
        return;
    }
    // wide state: decide on the host with the same Philox stream, apply through the fused-pass engine
    const double k = 0.70710678118654752440;
    (void)k;
    for (size_t e = 0; e < events.size(); ++e) {
        double u0, u1;
        philox_uniforms(seed_, 0, (noise_block_ << 16) | (uint64_t)e, u0, u1);
        const b200::TrajEvent& ev = events[e];
        b200::LogicalOp op{};
        op.target = ev.qubit;
        op.cmask = op.cval = 0;
        bool apply = false;
        auto pauli = [&](char which) {
            apply = true;
            for (double& x : op.m) x = 0.0;
            if (which == 'X') { op.m[2] = 1; op.m[4] = 1; }
            else if (which == 'Y') { op.m[3] = -1; op.m[5] = 1; }
            else { op.m[0] = 1; op.m[6] = -1; }
        };
        switch (static_cast<NoiseType>(ev.type)) {
            case NoiseType::Depolarizing: if (u0 < ev.p) pauli(u1 < 1.0 / 3.0 ? 'X' : (u1 < 2.0 / 3.0 ? 'Y' : 'Z')); break;
            case NoiseType::BitFlip: if (u0 < ev.p) pauli('X'); break;
            case NoiseType::PhaseFlip: if (u0 < ev.p) pauli('Z'); break;
            case NoiseType::BitPhaseFlip: if (u0 < ev.p) pauli('Y'); break;
            case NoiseType::AmplitudeDamping:
            case NoiseType::PhaseDamping: {
                const double total = state_->partialProbability(-1), p0 = state_->partialProbability(ev.qubit);
                const double P1 = total - p0, g = ev.p;
                apply = true;
                for (double& x : op.m) x = 0.0;
                if (u0 < g * P1) {
                    const double f = 1.0 / std::sqrt(P1);
                    if (static_cast<NoiseType>(ev.type) == NoiseType::AmplitudeDamping) op.m[2] = f;   // a0 <- a1 * f, a1 <- 0
                    else op.m[6] = f;                                                                    // a0 <- 0, a1 <- a1 * f
                } else {
                    const double f = 1.0 / std::sqrt(1.0 - g * P1);
                    op.m[0] = f;
                    op.m[6] = std::sqrt(1.0 - g) * f;
                }
                break;
            }
        }
        if (!apply) continue;
        b200::classify(op);
        b200::Program prog;
        std::string err;
        b200::CompileOptions opt = b200::default_options();
        opt.defer_x = false;
        if (!b200::compile_ops(num_qubits_, {op}, opt, prog, &err)) throw std::runtime_error(err);
        eng.execute(prog, state_->devicePtr(), 0);
    }
    ++noise_block_;
}

std::vector<std::complex<double>> NoisySimulator::getStateVector() const { return state_->toHost(); }
std::vector<double> NoisySimulator::getProbabilities() const { return state_->getProbabilities(); }

// One mt19937 draw per shot from the member engine (reference src/NoiseModel.cu:599-613).
std::vector<int> NoisySimulator::sample(int n_shots) {
    if (n_shots <= 0) return {};
    std::vector<double> u((size_t)n_shots);
    for (double& x : u) x = uniform_dist_(rng_);
    auto wide = state_->sampleWithUniforms(u.data(), n_shots);
    std::vector<int> out(wide.size());
    for (size_t i = 0; i < wide.size(); ++i) out[i] = static_cast<int>(wide[i]);
    return out;
}

// Bit `qubit` itself (not n-1-qubit), one draw, survivors divided by sqrt(sum of surviving |a|^2)
// (reference src/NoiseModel.cu:615-651).
int NoisySimulator::measureQubit(int qubit) {
    if (qubit < 0 || qubit >= num_qubits_) throw std::invalid_argument("Qubit index out of range");
    b200::SequentialCdf cdf(state_->devicePtr(), state_->getSize(), qubit, state_->engine());
    const double p0 = cdf.total();
    const double r = uniform_dist_(rng_);
    const int result = (r < p0) ? 0 : 1;
    // norm of the surviving half, summed in index order like the reference's host loop
    double norm2 = p0;
    if (result == 1) {
        b200::SequentialCdf ones(state_->devicePtr(), state_->getSize(), qubit | (1 << 8), state_->engine());
        norm2 = ones.total();
    }
    state_->collapse(qubit, result, 1.0 / std::sqrt(norm2));
    state_->engine().synchronize();
    return result;
}

// =====================================================================================================
// BatchedSimulator
// =====================================================================================================

BatchedSimulator::BatchedSimulator(int num_qubits, int batch_size)
    : num_qubits_(num_qubits), batch_size_(batch_size), state_size_(size_t(1) << (num_qubits > 0 ? num_qubits : 0)) {
    if (num_qubits < 1 || num_qubits > b200::kTrajMaxQubits)
        throw std::invalid_argument("BatchedSimulator keeps each trajectory in shared memory: 1 to " +
                                    std::to_string(b200::kTrajMaxQubits) + " qubits");
    if (batch_size < 1) throw std::invalid_argument("batch_size must be positive");
    b200::require_device();
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev));
    d_states_ = CudaMemory<cuDoubleComplex>(static_cast<size_t>(batch_size) * state_size_);
    d_avg_ = CudaMemory<double>(state_size_);
    std::random_device rd;
    rng_.seed(rd());
    seed_ = rng_();
    reset();
}

BatchedSimulator::BatchedSimulator(int num_qubits, int batch_size, const NoiseModel& noise_model)
    : BatchedSimulator(num_qubits, batch_size) {
    noise_model_ = noise_model;
}

BatchedSimulator::~BatchedSimulator() noexcept = default;
BatchedSimulator::BatchedSimulator(BatchedSimulator&&) noexcept = default;
BatchedSimulator& BatchedSimulator::operator=(BatchedSimulator&&) noexcept = default;

void BatchedSimulator::setSeed(unsigned int seed) {
    rng_.seed(seed);
    seed_ = seed;
    noise_block_ = 0;
}

void BatchedSimulator::reset() {
    b200::launch_batched_init(d_states_.get(), num_qubits_, batch_size_, num_sms_, nullptr);
    CUDA_CHECK(cudaStreamSynchronize(nullptr));
    avg_valid_ = false;
}

void BatchedSimulator::run(const Circuit& circuit) {
    validate_circuit_size(circuit, num_qubits_);
    std::vector<qsim_gate_t> recs;
    for (const GateOp& g : circuit.getGates()) recs.push_back(to_record(g));
    if (recs.empty()) return;
    const bool noisy = noise_model_.hasNoise();
    const auto events = noisy ? flatten(noise_model_.getChannels(), num_qubits_) : std::vector<b200::TrajEvent>{};
    const auto items = build_items(recs, noisy);
    // the kernel's epilogue leaves the average probabilities of the final states in d_avg_ (no second pass over the batch)
    run_trajectories(d_states_.get(), num_qubits_, batch_size_, items, events, seed_, noise_block_, num_sms_, nullptr, d_avg_.get());
    avg_valid_ = true;
    if (noisy) noise_block_ += recs.size();
}

std::vector<double> BatchedSimulator::getAverageProbabilities() const {
    if (!avg_valid_) {   // no run() since the states were last (re)set: one read of the batch
        b200::launch_batched_average(d_states_.get(), num_qubits_, batch_size_, const_cast<double*>(d_avg_.get()), num_sms_, nullptr);
        avg_valid_ = true;
    }
    std::vector<double> avg(state_size_);
    CUDA_CHECK(cudaStreamSynchronize(nullptr));
    d_avg_.copyToHost(avg.data(), state_size_);
    return avg;
}

std::vector<std::complex<double>> BatchedSimulator::getTrajectoryState(int trajectory_idx) const {
    if (trajectory_idx < 0 || trajectory_idx >= batch_size_) throw std::out_of_range("Invalid trajectory index");
    std::vector<std::complex<double>> st(state_size_);
    CUDA_CHECK(cudaMemcpy(st.data(), d_states_.get() + static_cast<size_t>(trajectory_idx) * state_size_,
                          state_size_ * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost));
    return st;
}

std::vector<double> BatchedSimulator::getProbabilities(int trajectory_idx) const {
    const auto st = getTrajectoryState(trajectory_idx);
    std::vector<double> p(state_size_);
    for (size_t i = 0; i < state_size_; ++i) p[i] = st[i].real() * st[i].real() + st[i].imag() * st[i].imag();
    return p;
}

// Draw order is trajectory-major, one mt19937 double per (trajectory, shot), as the reference
// (src/NoiseModel.cu:938-957); the CDF walk and lower_bound run on the device.
std::vector<int32_t> BatchedSimulator::sampleFlat(int n_shots, bool histogram_only, std::vector<int>* hist) {
    const size_t total = static_cast<size_t>(n_shots) * static_cast<size_t>(batch_size_);
    std::vector<double> u(total);
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    for (double& x : u) x = dist(rng_);
    CudaMemory<double> d_u(total);
    d_u.copyFromHost(u.data(), total);
    CudaMemory<int32_t> d_out(histogram_only ? 1 : total);
    CudaMemory<int32_t> d_hist(hist ? state_size_ : 1);
    if (hist) d_hist.zero();
    // one kernel: per trajectory the CDF in shared memory, per shot the index and (optionally) the histogram count
    b200::launch_batched_sample(d_states_.get(), num_qubits_, batch_size_, d_u.get(), n_shots, histogram_only ? nullptr : d_out.get(),
                                hist ? d_hist.get() : nullptr, num_sms_, nullptr);
    std::vector<int32_t> out;
    CUDA_CHECK(cudaStreamSynchronize(nullptr));
    if (hist) {
        std::vector<int32_t> h(state_size_);
        d_hist.copyToHost(h.data(), state_size_);
        hist->assign(h.begin(), h.end());
    }
    if (!histogram_only) {
        out.resize(total);
        d_out.copyToHost(out.data(), total);
    }
    return out;
}

std::vector<std::vector<int>> BatchedSimulator::sample(int n_shots) {
    if (n_shots <= 0) return {};
    const auto flat = sampleFlat(n_shots, false, nullptr);
    std::vector<std::vector<int>> out(static_cast<size_t>(n_shots), std::vector<int>(static_cast<size_t>(batch_size_)));
    for (int s = 0; s < n_shots; ++s)
        for (int t = 0; t < batch_size_; ++t) out[s][t] = flat[static_cast<size_t>(s) * batch_size_ + t];
    return out;
}

std::vector<int> BatchedSimulator::getHistogram(int n_shots) {
    std::vector<int> hist(state_size_, 0);
    if (n_shots <= 0) return hist;
    sampleFlat(n_shots, true, &hist);
    return hist;
}

}  // namespace qsim
