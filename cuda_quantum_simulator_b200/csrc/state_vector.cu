// StateVector: allocation, initialisation and read-out on top of the kernels in
// kernels_readout.cu.  Error contract follows the reference (src/StateVector.cu:130-342):
// invalid_argument for bad qubit counts / indices / n_shots <= 0, runtime_error for CUDA failures,
// a zero-probability outcome or an unnormalised state; destructors never throw.
#include "qsim/state_vector.cuh"

#include <cmath>
#include <cstring>
#include <random>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "engine.hpp"
#include "qsim/constants.hpp"
#include "qsim/cuda_memory.cuh"
#include "readout.cuh"

namespace qsim {

namespace {
void check_qubit_count(int n) {
    if (!isValidQubitCount(n))
        throw std::invalid_argument("Number of qubits must be between " + std::to_string(cuda_config::MIN_QUBITS) +
                                    " and " + std::to_string(cuda_config::MAX_QUBITS));
}
}  // namespace

StateVector::StateVector(int num_qubits) : num_qubits_(num_qubits) {
    check_qubit_count(num_qubits);
    size_ = size_t(1) << num_qubits;
    engine_ = std::make_unique<b200::Engine>();
    allocate();
    initializeZero();
}

StateVector::StateVector(int num_qubits, cuDoubleComplex* external) : num_qubits_(num_qubits) {
    check_qubit_count(num_qubits);
    if (!external) throw std::invalid_argument("external device memory must not be null");
    size_ = size_t(1) << num_qubits;
    engine_ = std::make_unique<b200::Engine>();
    d_state_ = external;
    owns_ = false;
}

StateVector::~StateVector() { deallocate(); }

StateVector::StateVector(StateVector&& o) noexcept
    : num_qubits_(o.num_qubits_), size_(o.size_), d_state_(o.d_state_), owns_(o.owns_), pending_basis_(o.pending_basis_),
      pending_idx_(o.pending_idx_), lazy_external_(o.lazy_external_), engine_(std::move(o.engine_)),
      prepared_cdf_(std::move(o.prepared_cdf_)) {
    o.d_state_ = nullptr;
    o.pending_basis_ = false;
    o.size_ = 0;
    o.num_qubits_ = 0;
}

StateVector& StateVector::operator=(StateVector&& o) noexcept {
    if (this != &o) {
        prepared_cdf_.reset();   // it references the engine and the amplitudes that are about to go
        deallocate();
        num_qubits_ = o.num_qubits_;
        size_ = o.size_;
        d_state_ = o.d_state_;
        owns_ = o.owns_;
        pending_basis_ = o.pending_basis_;
        pending_idx_ = o.pending_idx_;
        lazy_external_ = o.lazy_external_;
        engine_ = std::move(o.engine_);
        prepared_cdf_ = std::move(o.prepared_cdf_);
        o.d_state_ = nullptr;
        o.pending_basis_ = false;
        o.size_ = 0;
        o.num_qubits_ = 0;
    }
    return *this;
}

void StateVector::allocate() {
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_state_), size_ * sizeof(cuDoubleComplex)));
    owns_ = true;
}

void StateVector::deallocate() {
    if (d_state_ && owns_) {
        if (engine_) cudaStreamSynchronize(engine_->stream());
        cudaFree(d_state_);
    }
    d_state_ = nullptr;
}

void StateVector::initializeZero() { initializeBasis(0); }

void StateVector::initializeBasis(size_t basis_idx) {
    if (basis_idx >= size_) throw std::invalid_argument("Basis index out of range");
    if ((owns_ || lazy_external_) && !std::getenv("QSIM_EAGER_INIT")) {
        // nobody else can see this memory: defer the write (devicePtr() and every read-out materialise it)
        pending_basis_ = true;
        pending_idx_ = basis_idx;
        return;
    }
    pending_basis_ = false;
    b200::launch_init_basis(d_state_, size_, basis_idx, engine_->stream());
    engine_->countLaunch(2);
    engine_->synchronize();   // the reference synchronises here too (src/StateVector.cu:188-190)
}

void StateVector::initializeAllZero() {
    if ((owns_ || lazy_external_) && !std::getenv("QSIM_EAGER_INIT")) {
        pending_basis_ = true;
        pending_idx_ = kAllZero;
        return;
    }
    pending_basis_ = false;
    CUDA_CHECK(cudaMemsetAsync(d_state_, 0, size_ * sizeof(cuDoubleComplex), engine_->stream()));
    engine_->synchronize();
}

void StateVector::materialize() const {
    if (!pending_basis_) return;
    pending_basis_ = false;
    if (pending_idx_ == kAllZero) {
        CUDA_CHECK(cudaMemsetAsync(d_state_, 0, size_ * sizeof(cuDoubleComplex), engine_->stream()));
        engine_->countLaunch(1);
        return;
    }
    b200::launch_init_basis(d_state_, size_, pending_idx_, engine_->stream());
    engine_->countLaunch(2);
}

void StateVector::rebindExternal(cuDoubleComplex* external) {
    if (owns_) throw std::invalid_argument("rebindExternal needs a non-owning StateVector");
    if (!external) throw std::invalid_argument("external device memory must not be null");
    d_state_ = external;
}

bool StateVector::takePendingBasis(uint64_t* basis_idx) {
    if (!pending_basis_) return false;
    pending_basis_ = false;
    *basis_idx = pending_idx_;
    return true;
}

void StateVector::setFromHost(const std::complex<double>* amplitudes) {
    pending_basis_ = false;   // everything is overwritten
    CUDA_CHECK(cudaMemcpyAsync(d_state_, amplitudes, size_ * sizeof(cuDoubleComplex), cudaMemcpyHostToDevice,
                               engine_->stream()));
    engine_->synchronize();
}

void StateVector::toHost(std::complex<double>* out) const {
    static_assert(sizeof(std::complex<double>) == sizeof(cuDoubleComplex), "layout");
    CUDA_CHECK(cudaMemcpyAsync(out, devicePtr(), size_ * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost,
                               engine_->stream()));
    engine_->synchronize();
}

std::vector<std::complex<double>> StateVector::toHost() const {
    std::vector<std::complex<double>> v(size_);
    toHost(v.data());
    return v;
}

void StateVector::getProbabilities(double* out, uint64_t first, uint64_t count) const {
    if (first > size_ || count > size_ - first) throw std::invalid_argument("Probability range out of bounds");
    // stream the range through a bounded device buffer: never a second 2^n allocation
    const uint64_t cap = std::min<uint64_t>(count, uint64_t(1) << 24);
    double* buf = static_cast<double*>(engine_->scratch(1, cap * sizeof(double)));
    for (uint64_t done = 0; done < count; done += cap) {
        const uint64_t n = std::min<uint64_t>(cap, count - done);
        b200::launch_probabilities(devicePtr(), buf, first + done, n, engine_->numSMs(), engine_->stream());
        engine_->countLaunch();
        CUDA_CHECK(cudaMemcpyAsync(out + done, buf, n * sizeof(double), cudaMemcpyDeviceToHost, engine_->stream()));
        engine_->synchronize();
    }
}

std::vector<double> StateVector::getProbabilities() const {
    std::vector<double> p(size_);
    getProbabilities(p.data(), 0, size_);
    return p;
}

// Same value as the reference's index-order host loop (src/StateVector.cu:235-242), computed on
// the device by SequentialCdf.
double StateVector::getTotalProbability() const {
    b200::SequentialCdf cdf(devicePtr(), size_, -1, *engine_);
    engine_->countLaunch(cdf.launches());
    return cdf.total();
}

bool StateVector::isNormalized(double tolerance) const { return std::abs(getTotalProbability() - 1.0) <= tolerance; }

void StateVector::assertNormalized(double tolerance) const {
    const double total = getTotalProbability();
    if (std::abs(total - 1.0) > tolerance)
        throw std::runtime_error("State vector not normalized: total probability = " + std::to_string(total) +
                                 " (expected 1.0, tolerance = " + std::to_string(tolerance) + ")");
}

double StateVector::partialProbability(int bit) const {
    engine_->countLaunch(2);
    return b200::reduce_probability(devicePtr(), size_, bit, *engine_);
}

void StateVector::collapse(int bit, int outcome, double scale) {
    b200::launch_collapse(devicePtr(), size_, bit, outcome, scale, engine_->numSMs(), engine_->stream());
    engine_->countLaunch();
}

int StateVector::measureBit(int bit, double r, double* p0_out) {
    if (bit < 0 || bit >= num_qubits_)
        throw std::invalid_argument("Qubit index " + std::to_string(bit) + " out of range [0, " +
                                    std::to_string(num_qubits_ - 1) + "]");
    // p0 = index-order sum of the masked probabilities, exactly as the reference's host loop
    b200::SequentialCdf cdf(devicePtr(), size_, bit, *engine_);
    engine_->countLaunch(cdf.launches());
    const double p0 = cdf.total();
    if (p0_out) *p0_out = p0;
    const int result = (r < p0) ? 0 : 1;
    const double p_result = result == 0 ? p0 : 1.0 - p0;
    if (p_result < 1e-15)
        throw std::runtime_error("Measurement result " + std::to_string(result) +
                                 " has zero probability - state may be corrupted");
    collapse(bit, result, 1.0 / std::sqrt(p_result));
    engine_->synchronize();
    return result;
}

int StateVector::measure(int qubit, double r) {
    if (qubit < 0 || qubit >= num_qubits_)
        throw std::invalid_argument("Qubit index " + std::to_string(qubit) + " out of range [0, " +
                                    std::to_string(num_qubits_ - 1) + "]");
    // The reference's measure() addresses index bit n-1-qubit (src/StateVector.cu:87-89), unlike its
    // gates.  Kept for drop-in parity.
    return measureBit(num_qubits_ - 1 - qubit, r);
}

int StateVector::measure(int qubit) {
    std::random_device rd;
    std::mt19937 rng(rd());
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    return measure(qubit, dist(rng));
}

std::vector<int64_t> StateVector::sampleWithUniforms(const double* uniforms, int64_t n_shots) {
    if (n_shots <= 0) throw std::invalid_argument("n_shots must be positive");
    std::vector<int64_t> out((size_t)n_shots);
    if (b200::sample_small_state(devicePtr(), num_qubits_, uniforms, n_shots, out.data(), *engine_)) return out;
    b200::SequentialCdf cdf(devicePtr(), size_, -1, *engine_);
    cdf.sample(uniforms, n_shots, out.data());
    engine_->countLaunch(cdf.launches());
    return out;
}

std::vector<double> StateVector::marginalProbabilities(const std::vector<int>& bits) const {
    std::vector<double> out(size_t(1) << bits.size());
    b200::marginal_probabilities(devicePtr(), num_qubits_, bits.data(), (int)bits.size(), out.data(), *engine_);
    return out;
}

double StateVector::sampleShardPrepare() {
    prepared_cdf_ = std::make_unique<b200::SequentialCdf>(devicePtr(), size_, -1, *engine_, b200::SequentialCdf::Deferred{});
    return prepared_cdf_->approxTotal();
}

void StateVector::sampleShardClassify(double approx_c_init) {
    if (!prepared_cdf_) throw std::runtime_error("sampleShardClassify without sampleShardPrepare");
    prepared_cdf_->classify(approx_c_init);
}

double StateVector::sampleShard(double c_init, bool first_shard, const double* uniforms, int64_t n_shots, int64_t* out) {
    std::unique_ptr<b200::SequentialCdf> own = std::move(prepared_cdf_);
    if (own) own->stitch(c_init);
    else own = std::make_unique<b200::SequentialCdf>(devicePtr(), size_, -1, *engine_, c_init);
    b200::SequentialCdf& cdf = *own;
    const double c_end = cdf.total();
    if (n_shots > 0) {
        cdf.sample(uniforms, n_shots, out);
        for (int64_t i = 0; i < n_shots; ++i) {
            const double r = uniforms[i];
            const bool mine = (c_end >= r) && (c_init < r || first_shard);
            if (!mine) out[i] = -1;
        }
    }
    engine_->countLaunch(cdf.launches());
    return c_end;
}

std::vector<int64_t> StateVector::sampleSeeded(unsigned seed, int64_t n_shots) {
    if (n_shots <= 0) throw std::invalid_argument("n_shots must be positive");
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    std::vector<double> u((size_t)n_shots);
    for (auto& x : u) x = dist(rng);
    return sampleWithUniforms(u.data(), n_shots);
}

std::vector<int> StateVector::sample(int n_shots) {
    if (n_shots <= 0) throw std::invalid_argument("n_shots must be positive");
    std::random_device rd;
    auto wide = sampleSeeded(rd(), n_shots);
    std::vector<int> out(wide.size());
    for (size_t i = 0; i < wide.size(); ++i) out[i] = static_cast<int>(wide[i]);
    return out;
}

}  // namespace qsim
