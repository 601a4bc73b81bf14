// Compile-time constants, limits and the CUDA error -> exception macros of the qsim API.
// Stands in for the reference's include/Constants.hpp (same names, same meaning); the only
// deliberate difference is MAX_QUBITS: 30 there (include/Constants.hpp:68), 36 here, because the
// B200 configurations are 33 qubits on one GPU and 36 over eight.
#pragma once

#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

namespace qsim {

namespace constants {
constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 2.0 * PI;
constexpr double HALF_PI = PI / 2.0;
constexpr double QUARTER_PI = PI / 4.0;
constexpr double SQRT2 = 1.41421356237309504880;
constexpr double INV_SQRT2 = 0.70710678118654752440;
constexpr double EPSILON = 1e-10;
constexpr double PROBABILITY_EPSILON = 1e-12;
}  // namespace constants

namespace cuda_config {
constexpr int DEFAULT_BLOCK_SIZE = 256;
constexpr int REDUCTION_BLOCK_SIZE = 256;
constexpr int MIN_QUBITS = 1;
constexpr int MAX_QUBITS = 36;            // 2^36 amplitudes = 1 TiB over 8 x B200; 33 fit one GPU
constexpr int MAX_QUBITS_SINGLE_GPU = 33; // 128 GiB of the 180 GB HBM3e
constexpr int TARGET_CC_MAJOR = 10;       // sm_100a
constexpr int TARGET_CC_MINOR = 0;
}  // namespace cuda_config

inline int calcBlocks(size_t n, int block_size = cuda_config::DEFAULT_BLOCK_SIZE) {
    return static_cast<int>((n + static_cast<size_t>(block_size) - 1) / static_cast<size_t>(block_size));
}
inline bool isValidQubit(int qubit, int num_qubits) { return qubit >= 0 && qubit < num_qubits; }
inline bool isValidQubitCount(int num_qubits) {
    return num_qubits >= cuda_config::MIN_QUBITS && num_qubits <= cuda_config::MAX_QUBITS;
}

}  // namespace qsim

#define CUDA_CHECK(call)                                                                              \
    do {                                                                                              \
        cudaError_t qsim_err_ = (call);                                                               \
        if (qsim_err_ != cudaSuccess)                                                                 \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(qsim_err_) +    \
                                     " at " + __FILE__ + ":" + std::to_string(__LINE__));             \
    } while (0)

#define CUDA_CHECK_LAST_ERROR()                                                                       \
    do {                                                                                              \
        cudaError_t qsim_err_ = cudaGetLastError();                                                   \
        if (qsim_err_ != cudaSuccess)                                                                 \
            throw std::runtime_error(std::string("CUDA kernel error: ") +                             \
                                     cudaGetErrorString(qsim_err_) + " at " + __FILE__ + ":" +        \
                                     std::to_string(__LINE__));                                       \
    } while (0)
