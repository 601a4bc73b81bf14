// qsim::CPUSimulator — host-only class of the public API (reference include/Simulator.hpp:91-112), kept for
// source compatibility (the reference's own tests compare Simulator against it).  It is a separate class a
// caller asks for explicitly; no GPU path falls back to it.  Every gate goes through one generic
// controlled-2x2 routine, so unlike the reference's (src/Simulator.cu:214-220, 289-317) it also applies
// CRY, CRZ and Toffoli.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>

#include "program.hpp"
#include "qsim/simulator.hpp"

namespace qsim {

CPUSimulator::CPUSimulator(int num_qubits) : num_qubits_(num_qubits), size_(size_t(1) << num_qubits), state_(size_) { reset(); }

void CPUSimulator::reset() {
    std::fill(state_.begin(), state_.end(), std::complex<double>(0.0, 0.0));
    state_[0] = 1.0;
}

void CPUSimulator::run(const Circuit& circuit) {
    for (const GateOp& g : circuit.getGates()) applyGate(g);
}

void CPUSimulator::applyGate(const GateOp& gate) {
    qsim_gate_t rec{static_cast<int32_t>(gate.type), -1, -1, -1, gate.parameter};
    if (gate.qubits.size() > 0) rec.q0 = gate.qubits[0];
    if (gate.qubits.size() > 1) rec.q1 = gate.qubits[1];
    if (gate.qubits.size() > 2) rec.q2 = gate.qubits[2];
    std::vector<b200::LogicalOp> ops;
    if (!b200::lower_gate(rec, ops, 0)) return;
    using C = std::complex<double>;
    for (const b200::LogicalOp& op : ops) {
        const C a(op.m[0], op.m[1]), b(op.m[2], op.m[3]), c(op.m[4], op.m[5]), d(op.m[6], op.m[7]);
        const size_t bit = size_t(1) << op.target;
        for (size_t i = 0; i < size_; ++i) {
            if ((i & bit) || (i & op.cmask) != op.cval) continue;
            const C x = state_[i], y = state_[i | bit];
            state_[i] = a * x + b * y;
            state_[i | bit] = c * x + d * y;
        }
    }
}

std::vector<double> CPUSimulator::getProbabilities() const {
    std::vector<double> p(size_);
    for (size_t i = 0; i < size_; ++i) p[i] = std::norm(state_[i]);
    return p;
}

std::vector<int> CPUSimulator::sample(int n_shots) {
    const std::vector<double> p = getProbabilities();
    std::vector<double> cdf(size_);
    std::partial_sum(p.begin(), p.end(), cdf.begin());
    std::random_device rd;
    std::mt19937 rng(rd());
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::vector<int> out(static_cast<size_t>(n_shots > 0 ? n_shots : 0));
    for (int& s : out) s = static_cast<int>(std::lower_bound(cdf.begin(), cdf.end(), u(rng)) - cdf.begin());
    return out;
}

}  // namespace qsim
