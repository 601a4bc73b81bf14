// Legacy raw kernels of the qsim public headers (Gates.cuh, OptimizedGates.cuh, StateVector.cuh).
// One generic "controlled 2x2 on amplitude pairs" device routine serves all of them; the launch
// conventions (who picks the grid, one thread per pair or per amplitude) are the reference's
// (src/Gates.cu:31-410, src/OptimizedGates.cu:41-413, src/StateVector.cu:24-124).
#include <cmath>

#include "qsim/gates.cuh"
#include "qsim/optimized_gates.cuh"
#include "qsim/state_vector.cuh"

namespace qsim {

namespace {

struct M2 {   // 2x2 complex, row-major
    double ar, ai, br, bi, cr, ci, dr, di;
};

__device__ __forceinline__ size_t thread_index() { return (size_t)blockIdx.x * blockDim.x + threadIdx.x; }

__device__ __forceinline__ void mix(cuDoubleComplex& x, cuDoubleComplex& y, const M2& m) {
    const cuDoubleComplex u = x, v = y;
    x = make_cuDoubleComplex(m.ar * u.x - m.ai * u.y + m.br * v.x - m.bi * v.y, m.ar * u.y + m.ai * u.x + m.br * v.y + m.bi * v.x);
    y = make_cuDoubleComplex(m.cr * u.x - m.ci * u.y + m.dr * v.x - m.di * v.y, m.cr * u.y + m.ci * u.x + m.dr * v.y + m.di * v.x);
}

// thread p owns the pair whose indices differ in bit `target`
__device__ __forceinline__ void on_pair(cuDoubleComplex* s, int n, int target, const M2& m) {
    const size_t p = thread_index();
    if (p >= (size_t(1) << (n - 1))) return;
    const size_t low = (size_t(1) << target) - 1;
    const size_t i0 = (p & low) | ((p & ~low) << 1), i1 = i0 | (size_t(1) << target);
    mix(s[i0], s[i1], m);
}

// thread i owns amplitude i; acts when every control bit is 1 and the target bit is 0
__device__ __forceinline__ void on_controlled(cuDoubleComplex* s, int n, size_t cmask, int target, const M2& m) {
    const size_t i = thread_index();
    if (i >= (size_t(1) << n)) return;
    if ((i & cmask) != cmask || ((i >> target) & 1)) return;
    mix(s[i], s[i | (size_t(1) << target)], m);
}

__device__ __forceinline__ M2 real2(double a, double b, double c, double d) { return M2{a, 0, b, 0, c, 0, d, 0}; }
__device__ __forceinline__ M2 diag2(double dr0, double di0, double dr1, double di1) { return M2{dr0, di0, 0, 0, 0, 0, dr1, di1}; }

constexpr double kH = 0.70710678118654752440;
__device__ __forceinline__ M2 mat_x() { return real2(0, 1, 1, 0); }
__device__ __forceinline__ M2 mat_h() { return real2(kH, kH, kH, -kH); }
__device__ __forceinline__ M2 mat_ry(double t) { const double c = cos(t / 2), s = sin(t / 2); return real2(c, -s, s, c); }
__device__ __forceinline__ M2 mat_rz(double t) { const double c = cos(t / 2), s = sin(t / 2); return diag2(c, -s, c, s); }

// tile kernels: the block owns 2 * blockDim.x contiguous amplitudes staged in shared memory
template <class F>
__device__ __forceinline__ void tile_op(cuDoubleComplex* state, int n, int target, F&& pair_update) {
    extern __shared__ cuDoubleComplex tile[];
    const size_t total = size_t(1) << n, span = 2 * (size_t)blockDim.x, origin = (size_t)blockIdx.x * span;
    if (origin >= total) return;
    for (size_t k = threadIdx.x; k < span; k += blockDim.x)
        if (origin + k < total) tile[k] = state[origin + k];
    __syncthreads();
    const size_t stride = size_t(1) << target;
    if (target < 8) {   // same reach as the reference's tiled kernels
        const size_t p = threadIdx.x, low = stride - 1;
        const size_t l0 = (p & low) | ((p & ~low) << 1), l1 = l0 | stride;
        if (l1 < span && origin + l1 < total) pair_update(tile[l0], tile[l1]);
    }
    __syncthreads();
    for (size_t k = threadIdx.x; k < span; k += blockDim.x)
        if (origin + k < total) state[origin + k] = tile[k];
}

}  // namespace

// ---- Gates.cuh -----------------------------------------------------------------------------------------
__global__ void applyX(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, mat_x()); }
__global__ void applyY(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, M2{0, 0, 0, -1, 0, 1, 0, 0}); }
__global__ void applyZ(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, diag2(1, 0, -1, 0)); }
__global__ void applyH(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, mat_h()); }
__global__ void applyS(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, diag2(1, 0, 0, 1)); }
__global__ void applyT(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, diag2(1, 0, kH, kH)); }
__global__ void applySdag(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, diag2(1, 0, 0, -1)); }
__global__ void applyTdag(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, diag2(1, 0, kH, -kH)); }
__global__ void applyRx(cuDoubleComplex* s, int n, int t, double theta) {
    const double c = cos(theta / 2), si = sin(theta / 2);
    on_pair(s, n, t, M2{c, 0, 0, -si, 0, -si, c, 0});
}
__global__ void applyRy(cuDoubleComplex* s, int n, int t, double theta) { on_pair(s, n, t, mat_ry(theta)); }
__global__ void applyRz(cuDoubleComplex* s, int n, int t, double theta) { on_pair(s, n, t, mat_rz(theta)); }
__global__ void applyCNOT(cuDoubleComplex* s, int n, int c, int t) { on_controlled(s, n, size_t(1) << c, t, mat_x()); }
__global__ void applyCZ(cuDoubleComplex* s, int n, int c, int t) { on_controlled(s, n, size_t(1) << c, t, diag2(1, 0, -1, 0)); }
__global__ void applyCRY(cuDoubleComplex* s, int n, int c, int t, double theta) { on_controlled(s, n, size_t(1) << c, t, mat_ry(theta)); }
__global__ void applyCRZ(cuDoubleComplex* s, int n, int c, int t, double theta) { on_controlled(s, n, size_t(1) << c, t, mat_rz(theta)); }
__global__ void applyToffoli(cuDoubleComplex* s, int n, int c1, int c2, int t) {
    on_controlled(s, n, (size_t(1) << c1) | (size_t(1) << c2), t, mat_x());
}
__global__ void applySWAP(cuDoubleComplex* s, int n, int q1, int q2) {
    const size_t i = thread_index();
    if (i >= (size_t(1) << n)) return;
    if (((i >> q1) & 1) || !((i >> q2) & 1)) return;      // owner: q1 = 0, q2 = 1
    const size_t j = i ^ (size_t(1) << q1) ^ (size_t(1) << q2);
    const cuDoubleComplex tmp = s[i];
    s[i] = s[j];
    s[j] = tmp;
}

// ---- OptimizedGates.cuh ----------------------------------------------------------------------------------
__global__ void applyH_opt(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, mat_h()); }
__global__ void applyH_coalesced(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, mat_h()); }
__global__ void applyX_opt(cuDoubleComplex* s, int n, int t) { on_pair(s, n, t, mat_x()); }
__global__ void applyCNOT_opt(cuDoubleComplex* s, int n, int c, int t) { on_controlled(s, n, size_t(1) << c, t, mat_x()); }
__global__ void applyGate1Q_opt(cuDoubleComplex* s, int n, int t, cuDoubleComplex a, cuDoubleComplex b, cuDoubleComplex c,
                                cuDoubleComplex d) {
    on_pair(s, n, t, M2{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y});
}
__global__ void applyGate1Q_coalesced(cuDoubleComplex* s, int n, int t, cuDoubleComplex a, cuDoubleComplex b,
                                      cuDoubleComplex c, cuDoubleComplex d) {
    on_pair(s, n, t, M2{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y});
}
__global__ void applyH_shared(cuDoubleComplex* s, int n, int t) {
    const M2 h = mat_h();
    tile_op(s, n, t, [&](cuDoubleComplex& x, cuDoubleComplex& y) { mix(x, y, h); });
}
__global__ void applyRotation_shared(cuDoubleComplex* s, int n, int t, double cos_half, double sin_half, bool is_rx) {
    const M2 m = is_rx ? M2{cos_half, 0, 0, -sin_half, 0, -sin_half, cos_half, 0} : real2(cos_half, -sin_half, sin_half, cos_half);
    tile_op(s, n, t, [&](cuDoubleComplex& x, cuDoubleComplex& y) { mix(x, y, m); });
}
__global__ void applyFusedSingleQubitLayer(cuDoubleComplex* s, int n, const cuDoubleComplex* gate_params, unsigned int active) {
    const size_t i = thread_index();
    if (i >= (size_t(1) << n)) return;
    cuDoubleComplex a = s[i];
    for (int q = 0; q < n && q < 32; ++q) {
        if (!((active >> q) & 1u)) continue;
        const cuDoubleComplex f = gate_params[4 * q + (((i >> q) & 1) ? 3 : 0)];
        a = make_cuDoubleComplex(f.x * a.x - f.y * a.y, f.x * a.y + f.y * a.x);
    }
    s[i] = a;
}

void applyHadamardOptimized(cuDoubleComplex* state, int n_qubits, int target, cudaStream_t stream) {
    const size_t total = size_t(1) << n_qubits;
    if (target < SHARED_MEM_QUBIT_THRESHOLD && n_qubits <= 20) {
        const size_t span = 2 * OPT_BLOCK_SIZE;
        applyH_shared<<<(unsigned)((total + span - 1) / span), OPT_BLOCK_SIZE, span * sizeof(cuDoubleComplex), stream>>>(
            state, n_qubits, target);
    } else {
        applyH_coalesced<<<(unsigned)((total / 2 + OPT_BLOCK_SIZE - 1) / OPT_BLOCK_SIZE), OPT_BLOCK_SIZE, 0, stream>>>(
            state, n_qubits, target);
    }
}

void applyCNOTOptimized(cuDoubleComplex* state, int n_qubits, int control, int target, cudaStream_t stream) {
    const size_t total = size_t(1) << n_qubits;
    applyCNOT_opt<<<(unsigned)((total + OPT_BLOCK_SIZE - 1) / OPT_BLOCK_SIZE), OPT_BLOCK_SIZE, 0, stream>>>(state, n_qubits,
                                                                                                          control, target);
}

// ---- StateVector.cuh legacy kernels -----------------------------------------------------------------------
__global__ void initializeZeroKernel(cuDoubleComplex* state, size_t size) {
    const size_t i = thread_index();
    if (i < size) state[i] = make_cuDoubleComplex(i == 0 ? 1.0 : 0.0, 0.0);
}
__global__ void initializeBasisKernel(cuDoubleComplex* state, size_t size, size_t basis_idx) {
    const size_t i = thread_index();
    if (i < size) state[i] = make_cuDoubleComplex(i == basis_idx ? 1.0 : 0.0, 0.0);
}
__global__ void probabilityKernel(const cuDoubleComplex* state, double* probs, size_t size) {
    const size_t i = thread_index();
    if (i < size) probs[i] = __dadd_rn(__dmul_rn(state[i].x, state[i].x), __dmul_rn(state[i].y, state[i].y));
}
// in-place block reduction: data[blockIdx.x] = sum of this block's slice (dynamic shared memory: blockDim.x doubles)
__global__ void sumReductionKernel(double* data, size_t size) {
    extern __shared__ double slice[];
    const size_t i = thread_index();
    slice[threadIdx.x] = i < size ? data[i] : 0.0;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) slice[threadIdx.x] += slice[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) data[blockIdx.x] = slice[0];
}
// note the bit: measurement in this API addresses index bit n-1-qubit
__global__ void qubitProbabilityKernel(const cuDoubleComplex* state, double* probs, size_t size, int num_qubits, int qubit) {
    const size_t i = thread_index();
    if (i >= size) return;
    const bool one = (i >> (num_qubits - 1 - qubit)) & 1;
    probs[i] = one ? 0.0 : __dadd_rn(__dmul_rn(state[i].x, state[i].x), __dmul_rn(state[i].y, state[i].y));
}
__global__ void collapseStateKernel(cuDoubleComplex* state, size_t size, int num_qubits, int qubit, int result,
                                    double normalization_factor) {
    const size_t i = thread_index();
    if (i >= size) return;
    const int bit = (int)((i >> (num_qubits - 1 - qubit)) & 1);
    state[i] = bit == result ? make_cuDoubleComplex(state[i].x * normalization_factor, state[i].y * normalization_factor)
                             : make_cuDoubleComplex(0.0, 0.0);
}

}  // namespace qsim
