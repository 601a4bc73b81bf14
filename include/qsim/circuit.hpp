// Circuit IR of the qsim API: GateType / GateOp / fluent Circuit builder and the three factory
// functions.  Interface-compatible with the reference's include/Circuit.hpp:42-144 (enum order and
// GateOp::qubits ordering are part of the contract: the enum value crosses the C ABI as
// qsim_gate_t::type).
#pragma once

#include <cmath>
#include <string>
#include <vector>

namespace qsim {

enum class GateType { X, Y, Z, H, S, T, Sdag, Tdag, Rx, Ry, Rz, CNOT, CZ, CRY, CRZ, SWAP, Toffoli };

struct GateOp {
    GateType type;
    std::vector<int> qubits;   // [target] | [control, target] | [q1, q2] (SWAP) | [c1, c2, target]
    double parameter;          // rotation angle in radians (Rx, Ry, Rz, CRY, CRZ), else 0

    GateOp(GateType t, int q) : type(t), qubits{q}, parameter(0.0) {}
    GateOp(GateType t, int q, double theta) : type(t), qubits{q}, parameter(theta) {}
    GateOp(GateType t, int qa, int qb) : type(t), qubits{qa, qb}, parameter(0.0) {}
    GateOp(GateType t, int qa, int qb, double theta) : type(t), qubits{qa, qb}, parameter(theta) {}
    GateOp(GateType t, int qa, int qb, int qc) : type(t), qubits{qa, qb, qc}, parameter(0.0) {}
};

class Circuit {
public:
    explicit Circuit(int num_qubits);

    Circuit& x(int qubit) { return add1(GateType::X, qubit); }
    Circuit& y(int qubit) { return add1(GateType::Y, qubit); }
    Circuit& z(int qubit) { return add1(GateType::Z, qubit); }
    Circuit& h(int qubit) { return add1(GateType::H, qubit); }
    Circuit& s(int qubit) { return add1(GateType::S, qubit); }
    Circuit& t(int qubit) { return add1(GateType::T, qubit); }
    Circuit& sdag(int qubit) { return add1(GateType::Sdag, qubit); }
    Circuit& tdag(int qubit) { return add1(GateType::Tdag, qubit); }
    Circuit& rx(int qubit, double theta) { return add1p(GateType::Rx, qubit, theta); }
    Circuit& ry(int qubit, double theta) { return add1p(GateType::Ry, qubit, theta); }
    Circuit& rz(int qubit, double theta) { return add1p(GateType::Rz, qubit, theta); }
    Circuit& cnot(int control, int target) { return add2(GateType::CNOT, control, target); }
    Circuit& cx(int control, int target) { return cnot(control, target); }
    Circuit& cz(int control, int target) { return add2(GateType::CZ, control, target); }
    Circuit& cry(int control, int target, double theta) { return add2p(GateType::CRY, control, target, theta); }
    Circuit& crz(int control, int target, double theta) { return add2p(GateType::CRZ, control, target, theta); }
    Circuit& swap(int qubit1, int qubit2) { return add2(GateType::SWAP, qubit1, qubit2); }
    Circuit& toffoli(int control1, int control2, int target);
    Circuit& ccx(int control1, int control2, int target) { return toffoli(control1, control2, target); }

    int getNumQubits() const { return num_qubits_; }
    const std::vector<GateOp>& getGates() const { return gates_; }
    size_t getDepth() const;
    size_t getGateCount() const { return gates_.size(); }
    void clear() { gates_.clear(); }
    std::string toString() const;

private:
    int num_qubits_;
    std::vector<GateOp> gates_;

    Circuit& add1(GateType t, int q);
    Circuit& add1p(GateType t, int q, double theta);
    Circuit& add2(GateType t, int a, int b);
    Circuit& add2p(GateType t, int a, int b, double theta);
    void validateQubit(int qubit) const;
    void validateQubitPair(int q1, int q2) const;
    void validateQubitTriple(int q1, int q2, int q3) const;
};

Circuit createBellCircuit();
Circuit createGHZCircuit(int num_qubits);
Circuit createRandomCircuit(int num_qubits, int depth, unsigned int seed = 42);

}  // namespace qsim
