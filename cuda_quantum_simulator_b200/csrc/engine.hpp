// Executes compiled Programs on one GPU's amplitudes: uploads the op records, launches one
// fused-pass kernel per pass on the engine's stream, counts launches and (optionally) times the
// passes with CUDA events.  Internal to the library.
#pragma once

#include <cuda_runtime.h>
#include <cuComplex.h>

#include <cstdint>
#include <vector>

#include "program.hpp"

namespace qsim {
namespace b200 {

// Device-resident copy of a Program's op records (so a pre-compiled circuit launches with no
// host work beyond the kernel launches themselves).
struct DeviceProgram {
    Program host;
    DevOp* d_ops = nullptr;
    double* d_tables = nullptr;       // OP_PHASE tables
    PhaseTerm* d_terms = nullptr;     // OP_PHASE terms
    ~DeviceProgram();
    DeviceProgram() = default;
    DeviceProgram(const DeviceProgram&) = delete;
    DeviceProgram& operator=(const DeviceProgram&) = delete;
    void upload();   // throws std::runtime_error on CUDA failure
};

class Engine {
public:
    Engine();
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    void setStream(cudaStream_t s) { stream_ = s; }
    cudaStream_t stream() const { return stream_; }
    int numSMs() const { return num_sms_; }

    // Run a host Program (ops staged through a pinned buffer, asynchronous).
    void execute(const Program& p, cuDoubleComplex* state, uint64_t hi_bits);
    // Run a device-resident program.
    void execute(const DeviceProgram& p, cuDoubleComplex* state, uint64_t hi_bits);

    void synchronize() const;
    int64_t launches() const { return launches_; }
    void countLaunch(int64_t k = 1) { launches_ += k; }

    // Persistent device scratch (grow-only, two independent slots) for read-out temporaries, so the hot
    // read-out calls never cudaMalloc/cudaFree.
    void* scratch(int slot, size_t bytes);

    void setTiming(bool on);
    // Sum of pass-kernel device times since the last call, and the number of passes timed.
    void drainTiming(double* total_ms, int64_t* n_passes, std::vector<double>* each = nullptr);

private:
    cudaStream_t stream_ = nullptr;
    int num_sms_ = 0;
    int64_t launches_ = 0;
    void* scratch_[2] = {nullptr, nullptr};
    size_t scratch_cap_[2] = {0, 0};
    bool use_tensor_map_ = true;
    int stages_wanted_ = 0;   // 0 = as deep as shared memory allows
    // staging for execute(const Program&)
    DevOp* d_ops_ = nullptr;
    size_t d_cap_ = 0;
    DevOp* h_ops_ = nullptr;   // pinned
    size_t h_cap_ = 0;
    cudaEvent_t staged_ = nullptr;
    bool staged_pending_ = false;
    // timing
    bool timing_ = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events_;
    std::vector<cudaEvent_t> pool_;

    void launchAll(const Program& p, const DevOp* d_ops, const double* d_tables, const PhaseTerm* d_terms,
                   cuDoubleComplex* state, uint64_t hi_bits);
    // staging of the phase tables / terms of execute(const Program&)
    void* d_aux_ = nullptr;
    size_t d_aux_cap_ = 0;
    cudaEvent_t getEvent();
};

void require_device();   // throws std::runtime_error if no CUDA device is usable (no CPU fallback)

}  // namespace b200
}  // namespace qsim
