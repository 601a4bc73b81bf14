// Simulator: runs Circuits on a GPU-resident StateVector through the fused-pass engine.
// Interface-compatible with the reference's include/Simulator.hpp:53-112.  run() compiles the
// circuit (gate merging + pass/sweep planning, csrc/program.cpp) and launches one kernel per pass
// instead of one per gate; like the reference it is asynchronous until a read-out call.
#pragma once

#include <complex>
#include <cstdint>
#include <memory>
#include <vector>

#include "circuit.hpp"
#include "state_vector.cuh"

namespace qsim {

namespace b200 { struct DeviceProgram; }

class Simulator {
public:
    explicit Simulator(int num_qubits);
    Simulator(int num_qubits, cuDoubleComplex* external_device_memory);

    void reset();
    void run(const Circuit& circuit);
    void applyGate(const GateOp& gate);

    std::vector<std::complex<double>> getStateVector() const;
    std::vector<double> getProbabilities() const;
    std::vector<int> sample(int n_shots);
    int measureQubit(int qubit);

    int getNumQubits() const { return state_.getNumQubits(); }
    size_t getStateSize() const { return state_.getSize(); }

    // ---- additive API -------------------------------------------------------------------------
    StateVector& state() { return state_; }
    const StateVector& state() const { return state_; }
    void execute(const b200::DeviceProgram& program);   // pre-compiled circuit, no host work
    std::vector<int64_t> sampleSeeded(unsigned seed, int64_t n_shots) { return state_.sampleSeeded(seed, n_shots); }
    // marginal distribution of up to 12 qubits (qubits[i] -> bit i of the outcome), reduced on the device
    std::vector<double> getMarginalProbabilities(const std::vector<int>& qubits) const { return state_.marginalProbabilities(qubits); }
    int measureQubit(int qubit, double uniform_draw) { return state_.measure(qubit, uniform_draw); }
    void synchronize() const;

private:
    StateVector state_;
};

// Host-only reference implementation kept for API compatibility (reference include/Simulator.hpp:91-112).
// It is a separate, explicitly requested CPU class — no GPU path ever falls back to it.
// Unlike the reference's (src/Simulator.cu:214-220, 289-317) it also applies CRY, CRZ and Toffoli.
class CPUSimulator {
public:
    explicit CPUSimulator(int num_qubits);
    void reset();
    void run(const Circuit& circuit);
    void applyGate(const GateOp& gate);
    std::vector<std::complex<double>> getStateVector() const { return state_; }
    std::vector<double> getProbabilities() const;
    std::vector<int> sample(int n_shots);
    int getNumQubits() const { return num_qubits_; }

private:
    int num_qubits_;
    size_t size_;
    std::vector<std::complex<double>> state_;
};

}  // namespace qsim
