// Circuit compiler for the fused-pass engine (host only, no CUDA).
//
// A circuit (the reference's gate list, include/Circuit.hpp:64-84 of the reference) is lowered to
//   ops    : controlled one-qubit operators (dense 2x2, real 2x2, anti-diagonal, bit-flip, diagonal),
//            after merging runs of gates that act on the same (target, controls);
//   passes : groups of consecutive ops whose non-diagonal targets all live inside one set of
//            `t` "tile qubits".  One pass = one kernel launch = one read + one write of the state
//            (2 * 16 * 2^n bytes), however many gates it carries;
//   sweeps : within a pass, the assignment of tile bits to lane / register / warp bits.  A sweep
//            can hit any target held in a lane bit (warp shuffle) or register bit (in-thread).
//
// The reference executes one kernel and one full sweep of HBM per gate
// (src/Simulator.cu:28-154 of the reference); this compiler is what replaces that loop.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "qsim_b200.h"
#include "pass_desc.h"

namespace qsim {
namespace b200 {

// Host-side logical op: controlled one-qubit operator on global qubits.
struct LogicalOp {
    uint8_t kind;
    int target;
    uint64_t cmask, cval;        // controls over global qubits (cval allows control-on-zero)
    double m[8];
    int first_gate, n_gates;     // provenance (for statistics)
};

struct CompileOptions {
    int min_low_bits = 5;        // every pass keeps at least this many low bits contiguous (512 B runs)
    int max_tile_bits = kMaxTileBits;
    bool merge = true;           // merge runs of gates on the same (target, controls)
    bool reorder = true;         // commute ops across passes when legal (fewer passes)
    int n_global = 0;            // qubits >= n - n_global live in the rank id (sharded state)
    bool fold_tail_flips = true; // bit flips that can slide to the end (start) of a pass become store (load) addressing
    bool fuse_diagonals = true;  // runs of >= 3 diagonal gates become one OP_PHASE
    bool defer_x = true;         // carry uncontrolled X gates as an index-XOR frame, folded into the last pass's addressing
    uint64_t initial_xor = 0;    // X frame inherited from earlier segments (sharded driver: pending flips of global qubits)
    int isolate_bit = -1;        // (sharded driver) a pass whose HIGHEST tile qubit this is moves it with TMA instructions of its
                                 // own, so that the half of every tile with that bit set / clear can come from another GPU
                                 // (second half of a split qubit exchange, PassParams::redirect == 4)
};

struct JitKernel;                // a run-time specialised pass kernel (jit.hpp)
struct JitRequest;               // its generated source + cache key

struct Program {
    int n = 0;                   // total qubits (local + global)
    int n_local = 0;
    std::vector<LogicalOp> lops; // after merging
    std::vector<DevOp> ops;      // encoded, pass-major
    std::vector<PassDesc> passes;
    std::vector<double> phase_tables;      // (re, im) pairs, kPhaseTableSize entries per OP_PHASE
    std::vector<PhaseTerm> phase_terms;
    int64_t n_gates = 0;
    uint64_t global_xor = 0;     // X frame left on global qubits (bit q - n_local): a rank relabelling the caller owns
    // per pass: the specialised kernel once it has been looked up (filled lazily by the engine; empty slot = not yet,
    // jit_tried marks passes for which the interpreter was chosen)
    mutable std::vector<std::shared_ptr<JitKernel>> jit;
    mutable std::vector<std::shared_ptr<JitRequest>> jit_req;   // kept while the kernel is being compiled in the background
    mutable std::vector<char> jit_tried;
    mutable std::vector<std::shared_ptr<struct DualTune>> jit_tune;   // per pass: measured choice between the two builds (kernels.cuh)
    bool force_jit = false;      // specialise every pass whatever the state size (a pre-compiled circuit that will run many times)
    std::string describe() const;
};

// Defaults, overridable for experiments by QSIM_TILE_BITS / QSIM_MIN_LOW_BITS / QSIM_NO_MERGE / QSIM_NO_REORDER.
CompileOptions default_options();

// Gate list -> matrices.  Returns false for an unknown gate type.
bool lower_gate(const qsim_gate_t& g, std::vector<LogicalOp>& out, int gate_index);

// Full compile.  Ops whose non-diagonal target is a global qubit (>= n_local) are rejected
// (return false): the sharded driver must remap those qubits before compiling a segment.
bool compile(int n, const qsim_gate_t* gates, int64_t n_gates, const CompileOptions& opt, Program& out,
             std::string* error = nullptr);

// Compile an explicit op list (used by the density-matrix and noise paths, which build
// non-unitary / conjugated operators directly).
bool compile_ops(int n, std::vector<LogicalOp> lops, const CompileOptions& opt, Program& out,
                 std::string* error = nullptr);

void classify(LogicalOp& op);    // picks the cheapest OpKind for op.m

}  // namespace b200
}  // namespace qsim
