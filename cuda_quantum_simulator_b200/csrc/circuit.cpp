// Host-side circuit IR: validation, depth, printing and the seeded generators.
// Behavioural contract = reference src/Circuit.cpp:16-282 (exception types and messages'
// meaning, getDepth definition, createRandomCircuit draw order).
#include "qsim/circuit.hpp"

#include <algorithm>
#include <random>
#include <sstream>
#include <stdexcept>

#include "qsim/constants.hpp"

namespace qsim {

namespace {

void require_finite(double theta) {
    if (!std::isfinite(theta)) throw std::invalid_argument("Rotation angle must be a finite number");
}

const char* name_of(GateType t) {
    static const char* const names[] = {"X",  "Y",  "Z",  "H",    "S",  "T",   "Sdag", "Tdag",   "Rx",
                                        "Ry", "Rz", "CNOT", "CZ", "CRY", "CRZ", "SWAP", "Toffoli"};
    const int i = static_cast<int>(t);
    return (i >= 0 && i < 17) ? names[i] : "?";
}

bool is_parametric(GateType t) {
    return t == GateType::Rx || t == GateType::Ry || t == GateType::Rz || t == GateType::CRY || t == GateType::CRZ;
}

}  // namespace

Circuit::Circuit(int num_qubits) : num_qubits_(num_qubits) {
    if (!isValidQubitCount(num_qubits))
        throw std::invalid_argument("Number of qubits must be between " + std::to_string(cuda_config::MIN_QUBITS) +
                                    " and " + std::to_string(cuda_config::MAX_QUBITS));
}

void Circuit::validateQubit(int qubit) const {
    if (!isValidQubit(qubit, num_qubits_))
        throw std::out_of_range("Qubit index " + std::to_string(qubit) + " out of range [0, " +
                                std::to_string(num_qubits_ - 1) + "]");
}

void Circuit::validateQubitPair(int q1, int q2) const {
    validateQubit(q1);
    validateQubit(q2);
    if (q1 == q2) throw std::invalid_argument("Two-qubit gate requires distinct qubits");
}

void Circuit::validateQubitTriple(int q1, int q2, int q3) const {
    validateQubit(q1);
    validateQubit(q2);
    validateQubit(q3);
    if (q1 == q2 || q1 == q3 || q2 == q3) throw std::invalid_argument("Three-qubit gate requires three distinct qubits");
}

Circuit& Circuit::add1(GateType t, int q) {
    validateQubit(q);
    gates_.emplace_back(t, q);
    return *this;
}

Circuit& Circuit::add1p(GateType t, int q, double theta) {
    validateQubit(q);
    require_finite(theta);
    gates_.emplace_back(t, q, theta);
    return *this;
}

Circuit& Circuit::add2(GateType t, int a, int b) {
    validateQubitPair(a, b);
    gates_.emplace_back(t, a, b);
    return *this;
}

Circuit& Circuit::add2p(GateType t, int a, int b, double theta) {
    validateQubitPair(a, b);
    require_finite(theta);
    gates_.emplace_back(t, a, b, theta);
    return *this;
}

Circuit& Circuit::toffoli(int control1, int control2, int target) {
    validateQubitTriple(control1, control2, target);
    gates_.emplace_back(GateType::Toffoli, control1, control2, target);
    return *this;
}

// Depth = length of the longest chain of gates sharing a qubit (each gate occupies one time step on
// all of its qubits).
size_t Circuit::getDepth() const {
    std::vector<size_t> busy_until(static_cast<size_t>(num_qubits_), 0);
    size_t depth = 0;
    for (const GateOp& g : gates_) {
        size_t start = 0;
        for (int q : g.qubits) start = std::max(start, busy_until[static_cast<size_t>(q)]);
        for (int q : g.qubits) busy_until[static_cast<size_t>(q)] = start + 1;
        depth = std::max(depth, start + 1);
    }
    return depth;
}

std::string Circuit::toString() const {
    std::ostringstream out;
    out << "Circuit(" << num_qubits_ << " qubits, " << gates_.size() << " gates):\n";
    size_t index = 0;
    for (const GateOp& g : gates_) {
        out << "  " << index++ << ": " << name_of(g.type) << "(";
        const char* sep = "";
        for (int q : g.qubits) {
            out << sep << q;
            sep = ", ";
        }
        if (is_parametric(g.type)) out << ", " << g.parameter;
        out << ")\n";
    }
    return out.str();
}

Circuit createBellCircuit() {
    Circuit c(2);
    return c.h(0).cnot(0, 1);
}

Circuit createGHZCircuit(int num_qubits) {
    if (num_qubits < 2) throw std::invalid_argument("GHZ circuit requires at least 2 qubits");
    Circuit c(num_qubits);
    c.h(0);
    for (int q = 0; q + 1 < num_qubits; ++q) c.cnot(q, q + 1);
    return c;
}

// Seeded generator.  The draw order is the contract (it defines the benchmark circuits):
// per gate one draw of the gate kind in {H, X, CNOT, Rz}, one draw of the first qubit, then for CNOT
// draws of the second qubit until it differs, or for Rz one angle in [0, 2*pi).  libstdc++'s
// mt19937 / uniform_int_distribution / uniform_real_distribution define the actual numbers.
Circuit createRandomCircuit(int num_qubits, int depth, unsigned int seed) {
    std::mt19937 engine(seed);
    std::uniform_int_distribution<int> pick_qubit(0, num_qubits - 1);
    std::uniform_int_distribution<int> pick_kind(0, 3);
    std::uniform_real_distribution<double> pick_angle(0.0, constants::TWO_PI);

    Circuit c(num_qubits);
    for (int i = 0; i < depth; ++i) {
        const int kind = pick_kind(engine);
        const int a = pick_qubit(engine);
        if (kind == 0) {
            c.h(a);
        } else if (kind == 1) {
            c.x(a);
        } else if (kind == 2) {
            if (num_qubits < 2) {
                c.h(a);
            } else {
                int b = pick_qubit(engine);
                while (b == a) b = pick_qubit(engine);
                c.cnot(a, b);
            }
        } else {
            c.rz(a, pick_angle(engine));
        }
    }
    return c;
}

}  // namespace qsim
