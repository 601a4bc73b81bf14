// Density-matrix path: rho as a 2n-bit vector through the fused-pass engine.  See
// include/qsim/density_matrix.cuh for the model; gate matrices come from program.cpp's lower_gate, so
// they are the same operators the state-vector path applies.
#include "qsim/density_matrix.cuh"

#include <cmath>
#include <random>
#include <string>

#include "engine.hpp"
#include "program.hpp"
#include "qsim/constants.hpp"
#include "qsim_b200.h"
#include "readout.cuh"

namespace qsim {

// ---- helper kernels (grid-shape agnostic, one element per thread, like the reference's) ------------------
__global__ void dmComputeDiagonal(const cuDoubleComplex* rho, double* diag, size_t dim) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) diag[i] = rho[i * dim + i].x;
}

__global__ void dmComputeTrace(const cuDoubleComplex* rho, double* trace, size_t dim) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) atomicAdd(trace, rho[i * dim + i].x);
}

__global__ void dmInitPure(cuDoubleComplex* rho, const cuDoubleComplex* state, size_t dim) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= dim * dim) return;
    const cuDoubleComplex a = state[idx / dim], b = state[idx % dim];   // rho_rc = psi_r * conj(psi_c)
    rho[idx] = make_cuDoubleComplex(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

__global__ void dmInitMaxMixed(cuDoubleComplex* rho, size_t dim, double val) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) rho[i * dim + i] = make_cuDoubleComplex(val, 0.0);
}

__global__ void dmCollapseMeasurement(cuDoubleComplex* rho, int n_qubits, int target, int result, double norm_factor) {
    const size_t dim = size_t(1) << n_qubits;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= dim * dim) return;
    const size_t r = idx >> n_qubits, c = idx & (dim - 1);
    const bool keep = (int)((r >> target) & 1) == result && (int)((c >> target) & 1) == result;
    const cuDoubleComplex v = rho[idx];
    rho[idx] = keep ? make_cuDoubleComplex(v.x * norm_factor, v.y * norm_factor) : make_cuDoubleComplex(0.0, 0.0);
}

namespace {

__global__ void dm_marginal_one_kernel(const cuDoubleComplex* rho, size_t dim, int qubit, double* out) {
    // single block, fixed order: deterministic
    __shared__ double red[256];
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < dim; i += blockDim.x)
        if ((i >> qubit) & 1) acc += rho[i * dim + i].x;
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// ---- DensityMatrix -----------------------------------------------------------------------------------------

DensityMatrix::DensityMatrix(int n_qubits) : n_qubits_(n_qubits) {
    if (n_qubits < 1 || n_qubits > 14) throw std::invalid_argument("Density matrix supports 1 to 14 qubits");
    dim_ = size_t(1) << n_qubits;
    engine_ = std::make_unique<b200::Engine>();
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_rho_), dim_ * dim_ * sizeof(cuDoubleComplex)));
    reset();
}

DensityMatrix::DensityMatrix(int n_qubits, const std::vector<std::complex<double>>& pure_state) : DensityMatrix(n_qubits) {
    initFromPureState(pure_state);
}

DensityMatrix::~DensityMatrix() noexcept { release(); }

void DensityMatrix::release() noexcept {
    if (d_rho_) {
        if (engine_) cudaStreamSynchronize(engine_->stream());
        cudaFree(d_rho_);
    }
    d_rho_ = nullptr;
}

DensityMatrix::DensityMatrix(DensityMatrix&& o) noexcept
    : n_qubits_(o.n_qubits_), dim_(o.dim_), d_rho_(o.d_rho_), engine_(std::move(o.engine_)) {
    o.d_rho_ = nullptr;
    o.dim_ = 0;
    o.n_qubits_ = 0;
}

DensityMatrix& DensityMatrix::operator=(DensityMatrix&& o) noexcept {
    if (this != &o) {
        release();
        n_qubits_ = o.n_qubits_;
        dim_ = o.dim_;
        d_rho_ = o.d_rho_;
        engine_ = std::move(o.engine_);
        o.d_rho_ = nullptr;
        o.dim_ = 0;
        o.n_qubits_ = 0;
    }
    return *this;
}

void DensityMatrix::reset() {   // |0..0><0..0|
    b200::launch_init_basis(d_rho_, dim_ * dim_, 0, engine_->stream());
    engine_->countLaunch(2);
    engine_->synchronize();
}

void DensityMatrix::initFromPureState(const std::vector<std::complex<double>>& state) {
    if (state.size() != dim_) throw std::invalid_argument("State vector size doesn't match density matrix dimension");
    cuDoubleComplex* d_psi = static_cast<cuDoubleComplex*>(engine_->scratch(1, dim_ * sizeof(cuDoubleComplex)));
    CUDA_CHECK(cudaMemcpyAsync(d_psi, state.data(), dim_ * sizeof(cuDoubleComplex), cudaMemcpyHostToDevice, engine_->stream()));
    dmInitPure<<<blocks_for(dim_ * dim_), 256, 0, engine_->stream()>>>(d_rho_, d_psi, dim_);
    CUDA_CHECK_LAST_ERROR();
    engine_->countLaunch();
    engine_->synchronize();
}

void DensityMatrix::initMaximallyMixed() {
    CUDA_CHECK(cudaMemsetAsync(d_rho_, 0, dim_ * dim_ * sizeof(cuDoubleComplex), engine_->stream()));
    dmInitMaxMixed<<<blocks_for(dim_), 256, 0, engine_->stream()>>>(d_rho_, dim_, 1.0 / static_cast<double>(dim_));
    CUDA_CHECK_LAST_ERROR();
    engine_->countLaunch();
    engine_->synchronize();
}

std::vector<double> DensityMatrix::getProbabilities() const {
    double* d_diag = static_cast<double*>(engine_->scratch(1, dim_ * sizeof(double)));
    dmComputeDiagonal<<<blocks_for(dim_), 256, 0, engine_->stream()>>>(d_rho_, d_diag, dim_);
    CUDA_CHECK_LAST_ERROR();
    engine_->countLaunch();
    std::vector<double> p(dim_);
    CUDA_CHECK(cudaMemcpyAsync(p.data(), d_diag, dim_ * sizeof(double), cudaMemcpyDeviceToHost, engine_->stream()));
    engine_->synchronize();
    return p;
}

std::vector<std::complex<double>> DensityMatrix::getMatrix() const {
    std::vector<std::complex<double>> m(dim_ * dim_);
    CUDA_CHECK(cudaMemcpyAsync(m.data(), d_rho_, m.size() * sizeof(cuDoubleComplex), cudaMemcpyDeviceToHost, engine_->stream()));
    engine_->synchronize();
    return m;
}

double DensityMatrix::trace() const {
    double t = 0.0;
    for (double p : getProbabilities()) t += p;
    return t;
}

// sum_ij |rho_ij|^2 (= Tr rho^2 for Hermitian rho), the quantity the reference returns
// (src/DensityMatrix.cu:147-167), reduced on the device.
double DensityMatrix::purity() const {
    engine_->countLaunch(2);
    return b200::reduce_probability(d_rho_, dim_ * dim_, -1, *engine_);
}

bool DensityMatrix::isValid(double tolerance) const {
    if (std::abs(trace() - 1.0) > tolerance) return false;
    const double pur = purity(), min_purity = 1.0 / static_cast<double>(dim_);
    return !(pur < min_purity - tolerance || pur > 1.0 + tolerance);
}

// ---- DensityMatrixSimulator ----------------------------------------------------------------------------------

DensityMatrixSimulator::DensityMatrixSimulator(int n_qubits, const NoiseModel& noise)
    : n_qubits_(n_qubits), rho_(n_qubits), noise_model_(noise) {}

DensityMatrixSimulator::~DensityMatrixSimulator() noexcept = default;

void DensityMatrixSimulator::reset() { rho_.reset(); }

// U rho U^dagger: every controlled one-qubit operator of the gate acts on the row bits (q + n) as is and on the
// column bits (q) conjugated.
void DensityMatrixSimulator::lowerGate(const GateOp& gate, std::vector<b200::LogicalOp>& ops) const {
    qsim_gate_t rec{static_cast<int32_t>(gate.type), -1, -1, -1, gate.parameter};
    if (gate.qubits.size() > 0) rec.q0 = gate.qubits[0];
    if (gate.qubits.size() > 1) rec.q1 = gate.qubits[1];
    if (gate.qubits.size() > 2) rec.q2 = gate.qubits[2];
    for (int q : gate.qubits)
        if (q < 0 || q >= n_qubits_) throw std::out_of_range("Qubit index out of range");
    std::vector<b200::LogicalOp> one;
    if (!b200::lower_gate(rec, one, 0)) throw std::runtime_error("Unsupported gate type for density matrix simulation");
    const int n = n_qubits_;
    for (const b200::LogicalOp& op : one) {
        b200::LogicalOp row = op;
        row.target = op.target + n;
        row.cmask = op.cmask << n;
        row.cval = op.cval << n;
        b200::LogicalOp col = op;
        for (int k = 1; k < 8; k += 2) col.m[k] = -col.m[k];
        b200::classify(col);
        ops.push_back(row);
        ops.push_back(col);
    }
}

// One-qubit channel as a superoperator on (row bit a = q + n, column bit b = q).  With
// C = CNOT(a -> b) the pairs the channel mixes — (0,0)/(1,1) and (0,1)/(1,0) — become pairs that differ in
// bit a only, selected by b; so every channel is  C . [controlled real 2x2 on a, diagonal on b] . C.
// Consecutive channels on one qubit share their C's (the compiler merges C . C away).
void DensityMatrixSimulator::lowerChannel(NoiseType type, int q, double p, std::vector<b200::LogicalOp>& ops) const {
    if (q < 0 || q >= n_qubits_) throw std::out_of_range("Qubit index out of range");
    const int a = q + n_qubits_, b = q;
    auto make = [&](int target, uint64_t cmask, uint64_t cval, double m00, double m01, double m10, double m11) {
        b200::LogicalOp op{};
        op.target = target;
        op.cmask = cmask;
        op.cval = cval;
        op.m[0] = m00; op.m[2] = m01; op.m[4] = m10; op.m[6] = m11;
        op.first_gate = 0;
        op.n_gates = 0;
        b200::classify(op);
        ops.push_back(op);
    };
    const uint64_t A = 1ULL << a, B = 1ULL << b;
    auto cnot = [&] { make(b, A, A, 0, 1, 1, 0); };
    cnot();
    switch (type) {
        case NoiseType::Depolarizing:       // populations mix with 2p/3, coherences shrink by 1 - 4p/3
            make(a, B, 0, 1 - 2 * p / 3, 2 * p / 3, 2 * p / 3, 1 - 2 * p / 3);
            make(b, 0, 0, 1, 0, 0, 1 - 4 * p / 3);
            break;
        case NoiseType::AmplitudeDamping:   // rho00 += g rho11, rho11 *= 1-g, coherences *= sqrt(1-g)
            make(a, B, 0, 1, p, 0, 1 - p);
            make(b, 0, 0, 1, 0, 0, std::sqrt(1 - p));
            break;
        case NoiseType::PhaseDamping:
            make(b, 0, 0, 1, 0, 0, std::sqrt(1 - p));
            break;
        case NoiseType::BitFlip:            // (1-p) rho + p X rho X
            make(a, 0, 0, 1 - p, p, p, 1 - p);
            break;
        case NoiseType::PhaseFlip:          // coherences *= 1 - 2p
            make(b, 0, 0, 1, 0, 0, 1 - 2 * p);
            break;
        case NoiseType::BitPhaseFlip:       // (1-p) rho + p Y rho Y: populations mix with +p, coherences with -p
            make(a, B, 0, 1 - p, p, p, 1 - p);
            make(a, B, B, 1 - p, -p, -p, 1 - p);
            break;
    }
    cnot();
}

void DensityMatrixSimulator::lowerNoiseFor(const GateOp& gate, std::vector<b200::LogicalOp>& ops) const {
    if (!noise_model_.hasNoise()) return;
    for (int q : gate.qubits)
        for (const NoiseChannel& ch : noise_model_.getChannels())
            if (noise_model_.channelAppliesToQubit(ch, q)) lowerChannel(ch.type, q, ch.probability, ops);
}

void DensityMatrixSimulator::execute(std::vector<b200::LogicalOp>&& ops) {
    if (ops.empty()) return;
    b200::Program prog;
    std::string err;
    b200::CompileOptions opt = b200::default_options();
    if (!b200::compile_ops(2 * n_qubits_, std::move(ops), opt, prog, &err)) throw std::runtime_error(err);
    rho_.engine().execute(prog, rho_.getDevicePtr(), 0);
}

void DensityMatrixSimulator::run(const Circuit& circuit) {
    if (circuit.getNumQubits() != n_qubits_) throw std::invalid_argument("Circuit qubit count doesn't match simulator");
    std::vector<b200::LogicalOp> ops;
    for (const GateOp& g : circuit.getGates()) {
        lowerGate(g, ops);
        lowerNoiseFor(g, ops);
    }
    execute(std::move(ops));
}

void DensityMatrixSimulator::applyGate(const GateOp& gate) {
    std::vector<b200::LogicalOp> ops;
    lowerGate(gate, ops);
    lowerNoiseFor(gate, ops);
    execute(std::move(ops));
}

void DensityMatrixSimulator::applyChannel(NoiseType type, int qubit, double probability) {
    std::vector<b200::LogicalOp> ops;
    lowerChannel(type, qubit, probability, ops);
    execute(std::move(ops));
}

// result = (u < p1) ? 1 : 0 — compared against p1, as the reference does (src/DensityMatrix.cu:374-406);
// rho <- P rho P / p.
int DensityMatrixSimulator::measureQubit(int qubit, double u) {
    if (qubit < 0 || qubit >= n_qubits_) throw std::invalid_argument("Qubit index out of range");
    b200::Engine& eng = rho_.engine();
    double* d_p1 = static_cast<double*>(eng.scratch(1, sizeof(double)));
    dm_marginal_one_kernel<<<1, 256, 0, eng.stream()>>>(rho_.getDevicePtr(), rho_.getDimension(), qubit, d_p1);
    CUDA_CHECK_LAST_ERROR();
    double p1 = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&p1, d_p1, sizeof(double), cudaMemcpyDeviceToHost, eng.stream()));
    eng.synchronize();
    const int result = (u < p1) ? 1 : 0;
    const double prob = result ? p1 : 1.0 - p1;
    const size_t total = rho_.getNumElements();
    dmCollapseMeasurement<<<blocks_for(total), 256, 0, eng.stream()>>>(rho_.getDevicePtr(), n_qubits_, qubit, result, 1.0 / prob);
    CUDA_CHECK_LAST_ERROR();
    eng.countLaunch(2);
    eng.synchronize();
    return result;
}

int DensityMatrixSimulator::measureQubit(int qubit) {
    std::random_device rd;
    std::mt19937 gen(rd());
    std::uniform_real_distribution<> dis(0.0, 1.0);
    return measureQubit(qubit, dis(gen));
}

}  // namespace qsim
