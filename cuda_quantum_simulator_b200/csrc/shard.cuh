// Shard exchange kernels (internal header).  See kernels_shard.cu.
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace qsim {
namespace b200 {

void launch_swap_p2p(cuDoubleComplex* mine, cuDoubleComplex* peer, int n_local, int local_bit, int my_global_bit,
                     int num_sms, cudaStream_t stream);
void launch_pack_half(const cuDoubleComplex* state, cuDoubleComplex* buf, int local_bit, int bit_value,
                      uint64_t j_begin, uint64_t count, int num_sms, cudaStream_t stream);
void launch_unpack_half(cuDoubleComplex* state, const cuDoubleComplex* buf, int local_bit, int bit_value,
                        uint64_t j_begin, uint64_t count, int num_sms, cudaStream_t stream);

}  // namespace b200
}  // namespace qsim
