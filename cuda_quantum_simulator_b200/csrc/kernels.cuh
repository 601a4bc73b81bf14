// Device-side entry points shared by the host classes (internal header).
#pragma once

#include <cuda_runtime.h>
#include <cuComplex.h>

#include <cstdint>

#include "program.hpp"

namespace qsim {
namespace b200 {

size_t pass_smem_bytes(const PassDesc& pd, int stages);
int pick_stages(const PassDesc& pd, int wanted);
// Where the engine remembers, per program and pass, what the run-time specialisation has come to (all may be null:
// interpreter only).
struct JitSlots {
    // each points at TWO entries: [0] the one-group build of the pass kernel, [1] the two-warp-group build (jit.hpp)
    std::shared_ptr<JitKernel>* kernel = nullptr;    // the looked-up kernel
    std::shared_ptr<JitRequest>* request = nullptr;  // generated source + key, kept while a background compile is pending
    char* tried = nullptr;                           // 1: decided (kernel found, or the interpreter it is)
    bool force = false;                              // specialise whatever the state size, synchronously (pre-compiled circuits)
    struct DualTune* tune = nullptr;                 // which of the two builds is faster for this pass, measured (may be null)
};
// One-group or two-group build?  The FP64 estimate only says "worth trying": when both kernels are ready each is timed once
// in passing (CUDA events around one ordinary launch, read back later without blocking) and the faster one stays.
struct DualTune {
    cudaEvent_t ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int state[2] = {0, 0};     // 0 not timed, 1 events recorded, 2 measured
    float ms[2] = {0.f, 0.f};
    int choice = -1;           // -1 undecided, 0 one group, 1 two groups
    ~DualTune();
};
// host_ops: the pass's op records on the host (the key of the run-time specialised kernel).
cudaError_t launch_pass(const PassParams& params, int num_sms, cudaStream_t stream, const DevOp* host_ops = nullptr,
                        const JitSlots& jit = JitSlots());

}  // namespace b200
}  // namespace qsim
