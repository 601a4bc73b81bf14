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
};
// host_ops: the pass's op records on the host (the key of the run-time specialised kernel).
cudaError_t launch_pass(const PassParams& params, int num_sms, cudaStream_t stream, const DevOp* host_ops = nullptr,
                        const JitSlots& jit = JitSlots());

}  // namespace b200
}  // namespace qsim
