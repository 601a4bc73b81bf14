// Legacy per-gate kernels of the qsim API (reference include/Gates.cuh:58-111): one launch and one sweep of
// the state per gate, grid chosen by the caller (256-thread blocks; one thread per amplitude PAIR for the
// one-qubit kernels, one thread per AMPLITUDE for the controlled ones).  Kept so that code which launches
// them on StateVector::devicePtr() keeps compiling and linking; Simulator itself never uses them — it runs
// fused passes (csrc/kernels_pass.cu).
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <cstddef>

namespace qsim {

__global__ void applyX(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyY(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyZ(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyH(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyS(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyT(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applySdag(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyTdag(cuDoubleComplex* state, int n_qubits, int target);
__global__ void applyRx(cuDoubleComplex* state, int n_qubits, int target, double theta);
__global__ void applyRy(cuDoubleComplex* state, int n_qubits, int target, double theta);
__global__ void applyRz(cuDoubleComplex* state, int n_qubits, int target, double theta);
__global__ void applyCNOT(cuDoubleComplex* state, int n_qubits, int control, int target);
__global__ void applyCZ(cuDoubleComplex* state, int n_qubits, int control, int target);
__global__ void applyCRY(cuDoubleComplex* state, int n_qubits, int control, int target, double theta);
__global__ void applyCRZ(cuDoubleComplex* state, int n_qubits, int control, int target, double theta);
__global__ void applySWAP(cuDoubleComplex* state, int n_qubits, int qubit1, int qubit2);
__global__ void applyToffoli(cuDoubleComplex* state, int n_qubits, int control1, int control2, int target);

inline void getKernelConfig(size_t n_elements, int& blocks, int& threads) {
    threads = 256;
    blocks = static_cast<int>((n_elements + 255) / 256);
}

}  // namespace qsim
