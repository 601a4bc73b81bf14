// Device-side entry points shared by the host classes (internal header).
#pragma once

#include <cuda_runtime.h>
#include <cuComplex.h>

#include <cstdint>

#include "program.hpp"

namespace qsim {
namespace b200 {

constexpr int kPassThreads = kComputeThreads + 32;   // 8 compute warps + 1 TMA warp
constexpr int kMaxDynamicSmem = 227 * 1024;

struct PassParams {
    cuDoubleComplex* state;   // this GPU's amplitudes (2^pd.n of them)
    const DevOp* ops;         // device copy of this pass's ops
    uint64_t hi_bits;         // rank << n_local for a sharded state, else 0 (only used by controls)
    uint64_t n_tiles;         // 2^(pd.n - pd.t)
    PassDesc pd;
};
static_assert(sizeof(PassParams) <= 4000, "kernel parameter space");

size_t pass_smem_bytes(const PassDesc& pd);
cudaError_t launch_pass(const PassParams& params, int num_sms, cudaStream_t stream);

}  // namespace b200
}  // namespace qsim
