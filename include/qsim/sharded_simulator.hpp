// qsim::ShardedSimulator — the multi-GPU Simulator (additive: the reference is single-GPU, README.md:367 lists multi-GPU
// as future work; the surface mirrors include/Simulator.hpp:53-85 of the reference).
//
// One process per GPU.  The top log2(P) qubits of the amplitude index are the rank; every rank holds
// cuDoubleComplex[2^(n - log2 P)].  Gates whose non-diagonal target is a local qubit run in the fused-pass engine on every
// shard; controls, diagonal gates and X on rank qubits move no data; any other gate on a rank qubit first swaps that qubit
// with a local one: a pairwise half-shard exchange over NVLink, fused into the preceding pass's store where possible
// (the pass kernel writes the leaving half straight into the partner GPU's memory), else a peer-memory swap kernel, else
// NCCL send/recv.  The swap is never undone: the logical->physical qubit permutation and an X frame on the rank bits are
// carried and resolved at read-out.  Host side: C++17 + NCCL (loaded at run time) + CUDA IPC; no Python, no torch.
//
// Bootstrap: rank 0 calls createUniqueId() and hands the 128 bytes to every rank by whatever channel the application has
// (MPI, a file, a socket, torch.distributed in the Python mirror); every rank then constructs the simulator.  All ranks
// must make the same calls in the same order (the calls contain collectives).
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <array>
#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "qsim/circuit.hpp"

namespace qsim {

// ---- planning (pure host logic, no device; exposed for tests and tools) ---------------------------------------------

struct ShardStep {
    bool is_swap = false;
    std::vector<GateOp> gates;          // !is_swap: gates on PHYSICAL qubit positions
    int global_qubit = -1, local_qubit = -1;   // is_swap: positions exchanged
};

struct ShardPlan {
    int num_qubits = 0, n_global = 0;
    std::vector<ShardStep> steps;
    std::vector<int> perm;              // logical qubit -> physical position once the plan has run
    int numSwaps() const;
};

// Logical qubit -> physical position for a circuit that starts from |0...0>: qubits that are never a non-diagonal target go
// to the rank bits (no exchange ever), failing that the ones targeted last; the others keep their relative order.
std::vector<int> chooseInitialLayout(int num_qubits, int n_global, const std::vector<GateOp>& gates);
// Split a circuit into local segments separated by global<->local swaps (Belady choice of the evicted local position).
ShardPlan planCircuit(int num_qubits, int n_global, const std::vector<GateOp>& gates, const std::vector<int>& perm);

// ---- the simulator --------------------------------------------------------------------------------------------------

class ShardedSimulator {
public:
    static constexpr size_t kUniqueIdBytes = 128;
    enum class Exchange { Auto = 0, PeerMemory = 1, Nccl = 2 };

    static std::array<unsigned char, kUniqueIdBytes> createUniqueId();   // rank 0; needs NCCL

    // world_size must be a power of two; the calling thread's current CUDA device holds the shard.
    ShardedSimulator(int num_qubits, int rank, int world_size, const unsigned char* unique_id, Exchange exchange = Exchange::Auto);
    ~ShardedSimulator();
    ShardedSimulator(const ShardedSimulator&) = delete;
    ShardedSimulator& operator=(const ShardedSimulator&) = delete;

    // Simulator surface (reference include/Simulator.hpp:53-85)
    void reset();
    void run(const Circuit& circuit);
    std::vector<int64_t> sample(const std::vector<double>& uniforms);   // LOGICAL indices, the reference's sequential CDF order
    int measureQubit(int qubit, double uniform);                         // index bit n-1-qubit, as the reference
    int measureBit(int bit, double uniform, double* p0_out = nullptr);
    double getTotalProbability();
    std::vector<double> getMarginalProbabilities(const std::vector<int>& qubits);   // qubits[i] -> bit i of the outcome
    int getNumQubits() const { return n_; }
    size_t getStateSize() const { return size_t(1) << n_; }

    // pre-compiled plans: valid for the qubit layout / X frame they were compiled against
    struct CompiledPlan;
    std::shared_ptr<CompiledPlan> compile(const Circuit& circuit);
    std::vector<std::shared_ptr<CompiledPlan>> compileSequence(const Circuit& circuit, int k);   // k consecutive runs
    void execute(const CompiledPlan& plan);
    static int planPasses(const CompiledPlan& p);
    static int planSwaps(const CompiledPlan& p);
    static int planOps(const CompiledPlan& p);

    // layout and raw access
    void restoreIdentityLayout();
    const std::vector<int>& permutation() const { return perm_; }
    uint64_t frame() const { return frame_; }
    void setIdentityLayoutOnly(bool on) { identity_only_ = on; }   // never choose a layout for |0...0> (measurement tools)
    // Declares the stored layout to be the identity WITHOUT moving data: the logical qubits are relabelled (the state changes
    // by a qubit permutation).  For benchmarks that need every step to start from the same layout.
    void relabelIdentity();
    bool hasSecondBuffer() const { return bufs_[1] != nullptr; }
    std::vector<std::complex<double>> getLocalState();             // this rank's shard, stored layout
    void setLocalState(const std::complex<double>* amplitudes);
    cuDoubleComplex* devicePtr();
    void swapQubits(int global_position, int local_position);      // one separate exchange (benchmarks)
    int rank() const { return rank_; }
    int worldSize() const { return world_; }
    int localQubits() const { return nl_; }
    int64_t fusedExchanges() const { return fused_exchanges_; }
    int64_t separateExchanges() const { return separate_exchanges_; }
    // of fusedExchanges(): those that ran IN PLACE (no second buffer; stores over the partner's live shard under the
    // cross-GPU handshake of qsim_shard_execute_exchange_inplace)
    int64_t inPlaceExchanges() const { return inplace_exchanges_; }
    // of inPlaceExchanges(): those split over the pass before and the pass after the exchange (scatter half, gather half)
    int64_t splitExchanges() const { return split_exchanges_; }
    const char* exchangeName() const;
    void setStream(cudaStream_t s);
    void synchronize();
    void barrier();                                                 // stream-ordered barrier across the ranks
    void* localHandle() const { return shard_; }                    // qsim_sim_t* of the shard (C ABI interop: timing, counters)

private:
    struct Comm;
    int n_ = 0, ng_ = 0, nl_ = 0, rank_ = 0, world_ = 1;
    std::unique_ptr<Comm> comm_;
    void* shard_ = nullptr;                 // qsim_sim_t*
    cuDoubleComplex* bufs_[2] = {nullptr, nullptr};
    int cur_ = 0;
    std::vector<std::array<cuDoubleComplex*, 2>> peer_ptr_;   // per rank bit: the partner's two buffers (peer-mapped)
    std::vector<void*> peer_base_;
    Exchange exchange_ = Exchange::Auto;
    cuDoubleComplex* bounce_[2] = {nullptr, nullptr};
    size_t bounce_amps_ = 0;
    std::vector<int> perm_;
    uint64_t frame_ = 0;
    bool pristine_ = true, order_preserving_ = true, identity_only_ = false;
    int64_t fused_exchanges_ = 0, separate_exchanges_ = 0;
    cudaStream_t stream_ = nullptr;
    // in-place fused exchange: 1024 handshake words + the error word, one small allocation every partner maps
    unsigned long long* hs_ = nullptr;
    std::vector<unsigned long long*> peer_hs_;      // per rank bit
    uint64_t hs_epoch_ = 0;
    int64_t inplace_exchanges_ = 0, split_exchanges_ = 0;
    bool hs_unchecked_ = false;
    void checkExchanges(bool collective);           // throws if a handshake of an in-place exchange timed out

    void openPeers();
    void swapSeparate(int g, int l);
    bool runThenSwap(void* program, int g, int l);
    bool runSwapSplit(void* before, void* after, int g, int l);
    void swapNccl(int peer, int g, int l);
    std::vector<double> allGather(double v);
};

}  // namespace qsim
