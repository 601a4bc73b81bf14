// DensityMatrix / DensityMatrixSimulator of the qsim API (reference include/DensityMatrix.cuh:63-224).
//
// rho is stored row-major, element (r, c) at r * 2^n + c, i.e. a 2n-"qubit" vector whose low n index
// bits are the column and whose high n bits are the row.  A gate U acts as U on the row bit and
// conj(U) on the column bit; a one-qubit channel is a 4x4 superoperator on the (row bit, column bit)
// pair.  Both are lowered to the same controlled one-bit operators the state-vector path uses and run
// through the fused-pass kernel, so a whole noisy circuit is a handful of sweeps over rho instead of one
// 4^n sweep per gate and per channel (reference src/DensityMatrix.cu:214-296).
//
// Channels are the textbook ones the reference documents (include/DensityMatrix.cuh:250-264), not its
// kernels (whose Y gate, depolarizing and amplitude-damping kernels are defective — SURVEY D9).
// CRY, CRZ and Toffoli, which the reference rejects (src/DensityMatrix.cu:264-265), are supported.
// Noise schedule as the reference: after each gate, on the qubits that gate touched, every channel
// whose qubit list is empty or contains the qubit (src/DensityMatrix.cu:201-212, 269-296).
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <complex>
#include <memory>
#include <stdexcept>
#include <vector>

#include "circuit.hpp"
#include "noise_model.cuh"

namespace qsim {

namespace b200 { class Engine; struct LogicalOp; }

class DensityMatrix {
public:
    explicit DensityMatrix(int n_qubits);
    DensityMatrix(int n_qubits, const std::vector<std::complex<double>>& pure_state);
    ~DensityMatrix() noexcept;
    DensityMatrix(const DensityMatrix&) = delete;
    DensityMatrix& operator=(const DensityMatrix&) = delete;
    DensityMatrix(DensityMatrix&& other) noexcept;
    DensityMatrix& operator=(DensityMatrix&& other) noexcept;

    void reset();
    void initFromPureState(const std::vector<std::complex<double>>& state);
    void initMaximallyMixed();

    int getNumQubits() const { return n_qubits_; }
    size_t getDimension() const { return dim_; }
    size_t getNumElements() const { return dim_ * dim_; }
    size_t getMemoryBytes() const { return getNumElements() * sizeof(cuDoubleComplex); }

    std::vector<double> getProbabilities() const;
    std::vector<std::complex<double>> getMatrix() const;
    double trace() const;
    double purity() const;
    bool isValid(double tolerance = 1e-10) const;

    cuDoubleComplex* getDevicePtr() { return d_rho_; }
    const cuDoubleComplex* getDevicePtr() const { return d_rho_; }

    b200::Engine& engine() const { return *engine_; }

private:
    int n_qubits_ = 0;
    size_t dim_ = 0;
    cuDoubleComplex* d_rho_ = nullptr;
    std::unique_ptr<b200::Engine> engine_;
    void release() noexcept;
};

class DensityMatrixSimulator {
public:
    explicit DensityMatrixSimulator(int n_qubits, const NoiseModel& noise = NoiseModel());
    ~DensityMatrixSimulator() noexcept;

    void reset();
    void run(const Circuit& circuit);
    void applyGate(const GateOp& gate);

    std::vector<double> getProbabilities() const { return rho_.getProbabilities(); }
    std::vector<std::complex<double>> getDensityMatrix() const { return rho_.getMatrix(); }
    double getPurity() const { return rho_.purity(); }
    double getTrace() const { return rho_.trace(); }
    int measureQubit(int qubit);
    int getNumQubits() const { return n_qubits_; }

    // additive: measurement with an injected uniform draw; direct channel application
    int measureQubit(int qubit, double uniform_draw);
    void applyChannel(NoiseType type, int qubit, double probability);
    DensityMatrix& densityMatrix() { return rho_; }

private:
    int n_qubits_;
    DensityMatrix rho_;
    NoiseModel noise_model_;

    void lowerGate(const GateOp& gate, std::vector<b200::LogicalOp>& ops) const;
    void lowerChannel(NoiseType type, int qubit, double p, std::vector<b200::LogicalOp>& ops) const;
    void lowerNoiseFor(const GateOp& gate, std::vector<b200::LogicalOp>& ops) const;
    void execute(std::vector<b200::LogicalOp>&& ops);
};

// helper kernels of the reference's public header (include/DensityMatrix.cuh:266-272)
__global__ void dmComputeDiagonal(const cuDoubleComplex* rho, double* diag, size_t dim);
__global__ void dmComputeTrace(const cuDoubleComplex* rho, double* trace, size_t dim);
__global__ void dmInitPure(cuDoubleComplex* rho, const cuDoubleComplex* state, size_t dim);
__global__ void dmInitMaxMixed(cuDoubleComplex* rho, size_t dim, double val);
__global__ void dmCollapseMeasurement(cuDoubleComplex* rho, int n_qubits, int target, int result, double norm_factor);

}  // namespace qsim
