// StateVector: move-only owner of one GPU's cuDoubleComplex[2^n] amplitudes plus the read-out
// operations on it.  Interface-compatible with the reference's include/StateVector.cuh:66-124.
//
// Index convention: gates treat qubit q as bit q of the amplitude index (reference
// src/Gates.cu:19-25; its header comment claims the opposite, the code and tests are the truth —
// SURVEY.md §0.1).  `measure(q)` is the one exception and measures index bit n-1-q exactly as the
// reference does (src/StateVector.cu:87-89, 110-112); measureBit() addresses a bit directly.
//
// Additions: read-out with caller-supplied uniforms or a seed, 64-bit sample
// indices, ranged probabilities, a non-owning constructor for caller-allocated device memory, and
// shard metadata for the multi-GPU layout.  Nothing here ever allocates a second 2^n buffer.
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <complex>
#include <cstdint>
#include <memory>
#include <vector>

namespace qsim {

namespace b200 { class Engine; class SequentialCdf; }

class StateVector {
public:
    explicit StateVector(int num_qubits);
    // Non-owning view of caller-allocated device memory (16 << num_qubits bytes).
    StateVector(int num_qubits, cuDoubleComplex* external_device_memory);
    ~StateVector();

    StateVector(const StateVector&) = delete;
    StateVector& operator=(const StateVector&) = delete;
    StateVector(StateVector&& other) noexcept;
    StateVector& operator=(StateVector&& other) noexcept;

    void initializeZero();
    void initializeBasis(size_t basis_idx);
    void initializeAllZero();                 // additive: the zero vector (a shard that does not hold |0...0>)
    // additive: let a non-owning view defer its initialisation like an owning one does (the caller promises to reach
    // the memory only through this object until devicePtr() has been called)
    void allowLazyExternal(bool on) { lazy_external_ = on; }

    int getNumQubits() const { return num_qubits_; }
    size_t getSize() const { return size_; }
    // (a pending lazy basis state, see takePendingBasis, is written out first)
    cuDoubleComplex* devicePtr() { materialize(); return d_state_; }
    const cuDoubleComplex* devicePtr() const { materialize(); return d_state_; }

    std::vector<std::complex<double>> toHost() const;
    std::vector<double> getProbabilities() const;
    double getTotalProbability() const;
    bool isNormalized(double tolerance = 1e-10) const;
    void assertNormalized(double tolerance = 1e-10) const;

    int measure(int qubit);                 // index bit n-1-qubit (reference behaviour), random_device draw
    std::vector<int> sample(int n_shots);   // random_device-seeded, 32-bit indices (reference signature)

    // ---- additive API -------------------------------------------------------------------------
    void setFromHost(const std::complex<double>* amplitudes);
    void toHost(std::complex<double>* out) const;
    void getProbabilities(double* out, uint64_t first, uint64_t count) const;
    int measure(int qubit, double uniform_draw);                         // bit n-1-qubit, injected draw
    int measureBit(int bit, double uniform_draw, double* p0_out = nullptr);
    std::vector<int64_t> sampleWithUniforms(const double* uniforms, int64_t n_shots);
    std::vector<int64_t> sampleSeeded(unsigned seed, int64_t n_shots);   // mt19937(seed) draws
    // one shard of a distributed CDF: continues the sequential sum from c_init; returns the running sum at the end
    double sampleShard(double c_init, bool first_shard, const double* uniforms, int64_t n_shots, int64_t* out);
    // the same in steps, so that all shards sweep their amplitudes at the same time (see b200::SequentialCdf):
    // sampleShardPrepare() -> this shard's approximate total; sampleShardClassify(approximate sum of the shards before);
    // then sampleShard(exact sum of the shards before, ...) finishes.  No other read-out in between.
    double sampleShardPrepare();
    void sampleShardClassify(double approx_c_init);
    // marginal distribution over k index bits (bits[i] -> bit i of the outcome), k <= 12; nothing of size 2^n is built
    std::vector<double> marginalProbabilities(const std::vector<int>& bits) const;
    double partialProbability(int bit) const;                            // sum |a|^2 with index bit == 0 (bit<0: all)
    void collapse(int bit, int outcome, double scale);

    b200::Engine& engine() const { return *engine_; }

    // initializeZero / initializeBasis on memory this object owns only RECORD the basis state; it is written by
    // whoever needs the amplitudes first.  Simulator::run takes it over: its first pass generates the tiles on chip
    // instead of loading them (no memset sweep, no load sweep).  Returns false when the memory is already valid.
    bool takePendingBasis(uint64_t* basis_idx);   // kAllZero: the zero vector
    static constexpr uint64_t kAllZero = 0x7fffffffffffffffULL;
    cuDoubleComplex* rawDevicePtr() { return d_state_; }
    // non-owning views only: continue on other caller memory of the same size (double-buffered qubit exchange)
    void rebindExternal(cuDoubleComplex* external_device_memory);   // no materialisation: only with takePendingBasis

private:
    int num_qubits_ = 0;
    size_t size_ = 0;
    cuDoubleComplex* d_state_ = nullptr;
    bool owns_ = true;
    mutable bool pending_basis_ = false;      // the memory does not hold the state yet: it is |pending_idx_>
    bool lazy_external_ = false;
    mutable uint64_t pending_idx_ = 0;
    std::unique_ptr<b200::Engine> engine_;
    std::unique_ptr<b200::SequentialCdf> prepared_cdf_;
    void materialize() const;

    void allocate();
    void deallocate();
};

// Legacy raw kernels of the reference's public header (include/StateVector.cuh:131-149); callers
// pick the grid (256-thread blocks), so they are plain one-element-per-thread kernels.
__global__ void initializeZeroKernel(cuDoubleComplex* state, size_t size);
__global__ void initializeBasisKernel(cuDoubleComplex* state, size_t size, size_t basis_idx);
__global__ void probabilityKernel(const cuDoubleComplex* state, double* probs, size_t size);
__global__ void sumReductionKernel(double* data, size_t size);
__global__ void qubitProbabilityKernel(const cuDoubleComplex* state, double* probs, size_t size, int num_qubits,
                                       int qubit);
__global__ void collapseStateKernel(cuDoubleComplex* state, size_t size, int num_qubits, int qubit, int result,
                                    double normalization_factor);

}  // namespace qsim
