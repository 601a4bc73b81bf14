// "Optimized" kernel variants of the qsim API (reference include/OptimizedGates.cuh:65-166).  In the
// reference these are alternatives nobody but its tests and benchmarks launches; they are provided with the
// same names, launch conventions and results.  The real optimisation in this engine is elsewhere (fused passes).
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

namespace qsim {

constexpr int OPT_BLOCK_SIZE = 256;
constexpr int SHARED_MEM_QUBIT_THRESHOLD = 5;

// one thread per amplitude pair
__global__ void applyH_opt(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyX_opt(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyGate1Q_opt(cuDoubleComplex* state, int n_qubits, int target, cuDoubleComplex a, cuDoubleComplex b,
                                cuDoubleComplex c, cuDoubleComplex d);
__global__ void applyH_coalesced(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyGate1Q_coalesced(cuDoubleComplex* state, int n_qubits, int target, cuDoubleComplex a,
                                      cuDoubleComplex b, cuDoubleComplex c, cuDoubleComplex d);
// one thread per amplitude
__global__ void applyCNOT_opt(cuDoubleComplex* state, int n_qubits, int control, int target);
// one block per tile of 2 * blockDim.x amplitudes, 2 * blockDim.x * 16 bytes of dynamic shared memory
__global__ void applyH_shared(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyRotation_shared(cuDoubleComplex* state, int n_qubits, int target, double cos_half, double sin_half,
                                     bool is_rx);
// diagonal-only layer: amplitude *= gate_params[4q] (bit q = 0) or gate_params[4q + 3] (bit q = 1) for every
// qubit q in the mask; one thread per amplitude
__global__ void applyFusedSingleQubitLayer(cuDoubleComplex* state, int n_qubits, const cuDoubleComplex* gate_params,
                                           unsigned int active_qubits);

void applyHadamardOptimized(cuDoubleComplex* state, int n_qubits, int target, cudaStream_t stream = 0);
void applyCNOTOptimized(cuDoubleComplex* state, int n_qubits, int control, int target, cudaStream_t stream = 0);

}  // namespace qsim
