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
// host_ops: the pass's op records on the host (the key of the run-time specialised kernel); jit_slot / tried_slot: where
// the looked-up kernel is remembered between launches of the same program (all three may be null: interpreter only).
cudaError_t launch_pass(const PassParams& params, int num_sms, cudaStream_t stream, const DevOp* host_ops = nullptr,
                        std::shared_ptr<JitKernel>* jit_slot = nullptr, char* tried_slot = nullptr, bool force_jit = false);

}  // namespace b200
}  // namespace qsim
