// Simulator: Circuit -> compiled Program -> fused-pass launches.  Replaces the reference's
// one-launch-per-gate dispatch (src/Simulator.cu:28-154 of the reference).
#include "qsim/simulator.hpp"

#include <stdexcept>
#include <string>

#include "engine.hpp"
#include "program.hpp"
#include "qsim/constants.hpp"
#include "qsim_b200.h"

namespace qsim {

namespace {

qsim_gate_t to_record(const GateOp& g) {
    qsim_gate_t r{static_cast<int32_t>(g.type), -1, -1, -1, g.parameter};
    if (g.qubits.size() > 0) r.q0 = g.qubits[0];
    if (g.qubits.size() > 1) r.q1 = g.qubits[1];
    if (g.qubits.size() > 2) r.q2 = g.qubits[2];
    return r;
}

// A state that was reset and not touched since is still only a recorded basis index (StateVector::takePendingBasis):
// the first pass then generates its tiles on chip instead of loading them.
template <class ProgramT>
void run_program(StateVector& sv, const ProgramT& prog, const b200::Program& host) {
    uint64_t basis = 0;
    if (!host.passes.empty() && sv.takePendingBasis(&basis))
        sv.engine().execute(prog, sv.rawDevicePtr(), 0, (int64_t)basis);
    else
        sv.engine().execute(prog, sv.devicePtr(), 0);
}
void run_program(StateVector& sv, const b200::Program& prog) { run_program(sv, prog, prog); }

void run_records(StateVector& sv, const std::vector<qsim_gate_t>& recs) {
    if (recs.empty()) return;
    b200::Program prog;
    std::string err;
    if (!b200::compile(sv.getNumQubits(), recs.data(), (int64_t)recs.size(), b200::default_options(), prog, &err))
        throw std::runtime_error(err);
    run_program(sv, prog);
}

}  // namespace

Simulator::Simulator(int num_qubits) : state_(num_qubits) {}
Simulator::Simulator(int num_qubits, cuDoubleComplex* external) : state_(num_qubits, external) {}

void Simulator::reset() { state_.initializeZero(); }

void Simulator::run(const Circuit& circuit) {
    if (circuit.getNumQubits() != state_.getNumQubits())
        throw std::invalid_argument("Circuit qubit count doesn't match simulator");
    std::vector<qsim_gate_t> recs;
    recs.reserve(circuit.getGateCount());
    for (const GateOp& g : circuit.getGates()) recs.push_back(to_record(g));
    run_records(state_, recs);
}

void Simulator::applyGate(const GateOp& gate) {
    const size_t k = gate.qubits.size();
    const int t = static_cast<int>(gate.type);
    // arity / type agreement, as the reference's three dispatch tables enforce (src/Simulator.cu:92,126,148)
    if (k == 1 && t > static_cast<int>(GateType::Rz)) throw std::runtime_error("Unknown single-qubit gate type");
    if (k == 2 && (t < static_cast<int>(GateType::CNOT) || t > static_cast<int>(GateType::SWAP)))
        throw std::runtime_error("Unknown two-qubit gate type");
    if (k == 3 && gate.type != GateType::Toffoli) throw std::runtime_error("Unknown three-qubit gate type");
    if (k < 1 || k > 3) return;   // the reference ignores other arities
    run_records(state_, {to_record(gate)});
}

void Simulator::execute(const b200::DeviceProgram& program) {
    if (program.host.n != state_.getNumQubits())
        throw std::invalid_argument("Circuit qubit count doesn't match simulator");
    run_program(state_, program, program.host);
}

void Simulator::synchronize() const { state_.engine().synchronize(); }

std::vector<std::complex<double>> Simulator::getStateVector() const { return state_.toHost(); }
std::vector<double> Simulator::getProbabilities() const { return state_.getProbabilities(); }

std::vector<int> Simulator::sample(int n_shots) {
    // The reference's Simulator::sample does not validate n_shots (src/Simulator.cu:164-185);
    // a non-positive count yields an empty result there (vector(n) with n == 0) or throws length_error.
    if (n_shots <= 0) return std::vector<int>(static_cast<size_t>(n_shots < 0 ? throw std::length_error("n_shots") : 0));
    return state_.sample(n_shots);
}

int Simulator::measureQubit(int qubit) { return state_.measure(qubit); }

}  // namespace qsim
