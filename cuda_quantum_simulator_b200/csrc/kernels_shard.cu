// Multi-GPU support kernels: the qubit-remapping exchange between two shards.
//
// Layout: rank = the top n_global index bits, each GPU holds a contiguous
// cuDoubleComplex[2^n_local].  A non-diagonal gate on a global qubit g is made local by swapping g
// with a local qubit l: rank r and its partner r ^ (1 << (g - n_local)) exchange the half of their
// shards whose bit l differs from the owner's own value of g.  16 * 2^(n_local-1) bytes leave and
// enter each GPU, the NVLink roofline of the step.
//
//  * swap_p2p_kernel: in place over peer-mapped memory (CUDA IPC / peer access through NVSwitch).
//    Each of the two ranks moves half of the pairs, reading and writing the peer with 128-bit
//    accesses, so both link directions carry a quarter shard of reads and a quarter of writes.
//  * pack/unpack: the bounce-buffer variant used with NCCL send/recv.
#include "shard.cuh"

#include "qsim/constants.hpp"

namespace qsim {
namespace b200 {

namespace {

__device__ __forceinline__ uint64_t insert_bit(uint64_t j, int pos, uint64_t bit) {
    const uint64_t low = j & ((1ULL << pos) - 1);
    return low | (bit << pos) | ((j >> pos) << (pos + 1));
}

__global__ void swap_p2p_kernel(double2* __restrict__ mine, double2* __restrict__ peer, int local_bit, uint64_t my_bit,
                                uint64_t j_begin, uint64_t j_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nb = my_bit ^ 1;
    for (uint64_t j = j_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j_end; j += stride) {
        const uint64_t i_me = insert_bit(j, local_bit, nb), i_peer = insert_bit(j, local_bit, my_bit);
        const double2 x = mine[i_me];
        const double2 y = peer[i_peer];
        mine[i_me] = y;
        peer[i_peer] = x;
    }
}

__global__ void pack_half_kernel(const double2* __restrict__ state, double2* __restrict__ buf, int local_bit,
                                 uint64_t bit_value, uint64_t j_begin, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
        buf[k] = state[insert_bit(j_begin + k, local_bit, bit_value)];
}

__global__ void unpack_half_kernel(double2* __restrict__ state, const double2* __restrict__ buf, int local_bit,
                                   uint64_t bit_value, uint64_t j_begin, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
        state[insert_bit(j_begin + k, local_bit, bit_value)] = buf[k];
}

int grid_of(uint64_t n, int num_sms) {
    uint64_t b = (n + 255) / 256, cap = (uint64_t)num_sms * 16;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

void launch_swap_p2p(cuDoubleComplex* mine, cuDoubleComplex* peer, int n_local, int local_bit, int my_global_bit,
                     int num_sms, cudaStream_t stream) {
    const uint64_t pairs = 1ULL << (n_local - 1), half = pairs / 2;
    // the rank holding global bit 0 moves the first half of the pairs, its partner the second half
    const uint64_t jb = my_global_bit ? half : 0, je = my_global_bit ? pairs : (half ? half : pairs);
    if (pairs == 1 && my_global_bit) return;   // a single pair: rank with bit 0 moves it
    swap_p2p_kernel<<<grid_of(je - jb, num_sms), 256, 0, stream>>>(reinterpret_cast<double2*>(mine),
                                                                   reinterpret_cast<double2*>(peer), local_bit,
                                                                   (uint64_t)my_global_bit, jb, je);
    CUDA_CHECK_LAST_ERROR();
}

void launch_pack_half(const cuDoubleComplex* state, cuDoubleComplex* buf, int local_bit, int bit_value,
                      uint64_t j_begin, uint64_t count, int num_sms, cudaStream_t stream) {
    pack_half_kernel<<<grid_of(count, num_sms), 256, 0, stream>>>(reinterpret_cast<const double2*>(state),
                                                                  reinterpret_cast<double2*>(buf), local_bit,
                                                                  (uint64_t)bit_value, j_begin, count);
    CUDA_CHECK_LAST_ERROR();
}

void launch_unpack_half(cuDoubleComplex* state, const cuDoubleComplex* buf, int local_bit, int bit_value,
                        uint64_t j_begin, uint64_t count, int num_sms, cudaStream_t stream) {
    unpack_half_kernel<<<grid_of(count, num_sms), 256, 0, stream>>>(reinterpret_cast<double2*>(state),
                                                                    reinterpret_cast<const double2*>(buf), local_bit,
                                                                    (uint64_t)bit_value, j_begin, count);
    CUDA_CHECK_LAST_ERROR();
}

}  // namespace b200
}  // namespace qsim
