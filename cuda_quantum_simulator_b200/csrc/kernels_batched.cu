// Batched noisy trajectories: ONE trajectory per CTA, resident in shared memory for the whole circuit.
//
// The reference's BatchedSimulator (src/NoiseModel.cu:657-972) sweeps the whole [batch][2^n] array in HBM once
// per gate and once per (channel, qubit), keeps a 48-byte XORWOW state per amplitude pair, implements only
// X/Y/Z/H/CNOT and depolarizing noise, and draws one random number per amplitude pair.
// Here a trajectory (<= 2^13 amplitudes) is loaded once, every gate and every noise event is applied in
// shared memory, and it is written back once: HBM traffic is 2 * 16 * 2^n bytes per trajectory for the whole
// circuit.  Noise is a proper quantum-trajectory unravelling — one draw per (trajectory, gate, channel, qubit)
// from a counter-based Philox4x32-10 stream — whose average converges to the exact Kraus channel
// (checked against the oracle's density-matrix restatement, tests/test_noise_gpu.py).
#include "batched.cuh"

#include "qsim/constants.hpp"

namespace qsim {
namespace b200 {

namespace {

constexpr int kTrajThreads = 256;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniforms in [0,1) with 53 random bits each
__device__ __forceinline__ void traj_uniforms(uint32_t seed, uint64_t traj, uint64_t event, double& u0, double& u1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)event, (uint32_t)(event >> 32), (uint32_t)traj, (uint32_t)(traj >> 32), seed, 0x51534D42u, o);
    u0 = (double)((((uint64_t)o[1] << 32) | o[0]) >> 11) * 0x1.0p-53;
    u1 = (double)((((uint64_t)o[3] << 32) | o[2]) >> 11) * 0x1.0p-53;
}

__device__ __forceinline__ uint32_t insert_zero(uint32_t p, int t) {
    return (p & ((1u << t) - 1u)) | ((p >> t) << (t + 1));
}

// controlled 2x2 operator on the shared-memory state
__device__ __forceinline__ void apply_2x2(double2* s, int n, int t, uint32_t cmask, uint32_t cval, const double* m) {
    const uint32_t half = 1u << (n - 1), bit = 1u << t;
    const double m00r = m[0], m00i = m[1], m01r = m[2], m01i = m[3], m10r = m[4], m10i = m[5], m11r = m[6], m11i = m[7];
    for (uint32_t p = threadIdx.x; p < half; p += kTrajThreads) {
        const uint32_t i0 = insert_zero(p, t), i1 = i0 | bit;
        if ((i0 & cmask) != cval) continue;
        const double2 a = s[i0], b = s[i1];
        s[i0] = make_double2(m00r * a.x - m00i * a.y + m01r * b.x - m01i * b.y, m00r * a.y + m00i * a.x + m01r * b.y + m01i * b.x);
        s[i1] = make_double2(m10r * a.x - m10i * a.y + m11r * b.x - m11i * b.y, m10r * a.y + m10i * a.x + m11r * b.y + m11i * b.x);
    }
    __syncthreads();
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < kTrajThreads / 32) ? red[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
}

__device__ void noise_event(double2* s, int n, const TrajEvent& ev, double u0, double u1, double* red) {
    const int q = ev.qubit;
    const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0}, Y[8] = {0, 0, 0, -1, 0, 1, 0, 0}, Z[8] = {1, 0, 0, 0, 0, 0, -1, 0};
    switch (ev.type) {
        case 0:  // depolarizing: with probability p one of X, Y, Z (thirds)
            if (u0 < ev.p) apply_2x2(s, n, q, 0, 0, u1 < 1.0 / 3.0 ? X : (u1 < 2.0 / 3.0 ? Y : Z));
            break;
        case 3: if (u0 < ev.p) apply_2x2(s, n, q, 0, 0, X); break;   // bit flip
        case 4: if (u0 < ev.p) apply_2x2(s, n, q, 0, 0, Z); break;   // phase flip
        case 5: if (u0 < ev.p) apply_2x2(s, n, q, 0, 0, Y); break;   // bit-phase flip
        case 1:    // amplitude damping  K0 = diag(1, sqrt(1-g)), K1 = [[0, sqrt(g)], [0, 0]]
        case 2: {  // phase damping      K0 = diag(1, sqrt(1-g)), K1 = diag(0, sqrt(g))
            const uint32_t size = 1u << n, bit = 1u << q;
            double part = 0.0;
            for (uint32_t i = threadIdx.x; i < size; i += kTrajThreads)
                if (i & bit) { const double2 a = s[i]; part += a.x * a.x + a.y * a.y; }
            const double P1 = block_sum(part, red);
            const double g = ev.p;
            if (u0 < g * P1) {   // jump
                const double f = 1.0 / sqrt(P1);
                const uint32_t half = size >> 1;
                for (uint32_t p = threadIdx.x; p < half; p += kTrajThreads) {
                    const uint32_t i0 = insert_zero(p, q), i1 = i0 | bit;
                    const double2 b = s[i1];
                    if (ev.type == 1) { s[i0] = make_double2(b.x * f, b.y * f); s[i1] = make_double2(0.0, 0.0); }
                    else { s[i0] = make_double2(0.0, 0.0); s[i1] = make_double2(b.x * f, b.y * f); }
                }
            } else {             // no jump, renormalised
                const double f = 1.0 / sqrt(1.0 - g * P1), kf = sqrt(1.0 - g) * f;
                for (uint32_t i = threadIdx.x; i < size; i += kTrajThreads) {
                    const double w = (i & bit) ? kf : f;
                    const double2 a = s[i];
                    s[i] = make_double2(a.x * w, a.y * w);
                }
            }
            __syncthreads();
            break;
        }
        default: break;
    }
}

__global__ void __launch_bounds__(kTrajThreads) trajectory_kernel(cuDoubleComplex* __restrict__ states, int n, int64_t batch,
                                                                  const TrajItem* __restrict__ items, int n_items,
                                                                  const TrajEvent* __restrict__ events, int n_events,
                                                                  uint32_t seed, uint64_t traj_offset,
                                                                  uint64_t first_noise_block) {
    extern __shared__ __align__(16) unsigned char traj_smem[];
    double2* s = reinterpret_cast<double2*>(traj_smem);
    const uint32_t size = 1u << n;
    double* red = reinterpret_cast<double*>(s + size);
    for (int64_t traj = blockIdx.x; traj < batch; traj += gridDim.x) {
        double2* g = reinterpret_cast<double2*>(states) + (size_t)traj * size;
        for (uint32_t i = threadIdx.x; i < size; i += kTrajThreads) s[i] = g[i];
        __syncthreads();
        uint64_t block = first_noise_block;
        for (int it = 0; it < n_items; ++it) {
            const TrajItem& item = items[it];
            if (item.kind == 0) {
                apply_2x2(s, n, item.target, (uint32_t)item.cmask, (uint32_t)item.cval, item.m);
            } else {
                for (int e = 0; e < n_events; ++e) {
                    double u0, u1;
                    traj_uniforms(seed, (uint64_t)traj + traj_offset, block * (uint64_t)n_events + (uint64_t)e, u0, u1);
                    noise_event(s, n, events[e], u0, u1, red);
                }
                ++block;
            }
        }
        for (uint32_t i = threadIdx.x; i < size; i += kTrajThreads) g[i] = s[i];
        __syncthreads();
    }
}

__global__ void batched_init_kernel(cuDoubleComplex* states, uint64_t size_mask, uint64_t total) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        states[i] = make_cuDoubleComplex((i & size_mask) == 0 ? 1.0 : 0.0, 0.0);
}

// each block owns a slice of trajectories and adds its partial column sums with one atomic per basis state
__global__ void batched_average_kernel(const cuDoubleComplex* __restrict__ states, int n, int64_t batch, double inv_batch,
                                       double* __restrict__ avg) {
    const uint32_t size = 1u << n;
    const int64_t per = (batch + gridDim.y - 1) / gridDim.y;
    const int64_t t0 = (int64_t)blockIdx.y * per, t1 = (t0 + per < batch) ? t0 + per : batch;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int64_t t = t0; t < t1; ++t) {
            const cuDoubleComplex a = states[(size_t)t * size + i];
            acc += (a.x * a.x + a.y * a.y) * inv_batch;
        }
        atomicAdd(&avg[i], acc);
    }
}

// one warp per trajectory: sequential fp64 CDF (std::partial_sum order) in shared memory, then lower_bound per shot
__global__ void batched_sample_kernel(const cuDoubleComplex* __restrict__ states, int n, int64_t batch,
                                      const double* __restrict__ uniforms, int n_shots, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char samp_smem[];
    const uint32_t size = 1u << n;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* cum = reinterpret_cast<double*>(samp_smem) + (size_t)warp * size;
    for (int64_t traj = (int64_t)blockIdx.x * warps + warp; traj < batch; traj += (int64_t)gridDim.x * warps) {
        const cuDoubleComplex* a = states + (size_t)traj * size;
        double c = 0.0;
        for (uint32_t g = 0; g < size; g += 32) {
            double p = 0.0;
            if (g + lane < size) { const cuDoubleComplex v = a[g + lane]; p = __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)); }
            const uint32_t lim = (size - g) < 32u ? (size - g) : 32u;
            double mine = 0.0;
            for (uint32_t j = 0; j < lim; ++j) {
                c = __dadd_rn(c, __shfl_sync(0xffffffffu, p, j));
                if (j == (uint32_t)lane) mine = c;
            }
            if (g + lane < size) cum[g + lane] = mine;
        }
        __syncwarp();
        for (int shot = lane; shot < n_shots; shot += 32) {
            const double r = uniforms[(size_t)traj * n_shots + shot];
            uint32_t lo = 0, hi = size;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cum[mid] >= r) hi = mid; else lo = mid + 1; }
            out[(size_t)shot * batch + traj] = (int32_t)lo;
        }
        __syncwarp();
    }
}

__global__ void histogram_kernel(const int32_t* __restrict__ samples, int64_t count, uint32_t size, int32_t* __restrict__ hist) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const int32_t v = samples[i];
        if (v >= 0 && (uint32_t)v < size) atomicAdd(&hist[v], 1);
    }
}

}  // namespace

void launch_trajectories(cuDoubleComplex* states, int n, int64_t batch, const TrajItem* d_items, int n_items,
                         const TrajEvent* d_events, int n_events, uint32_t seed, uint64_t traj_offset,
                         uint64_t first_noise_block, int num_sms, cudaStream_t stream) {
    const size_t smem = ((size_t)16 << n) + 40 * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(trajectory_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(220 * 1024) / smem));
    int64_t grid = std::min<int64_t>(batch, (int64_t)num_sms * per_sm);
    trajectory_kernel<<<(unsigned)grid, kTrajThreads, smem, stream>>>(states, n, batch, d_items, n_items, d_events, n_events,
                                                                     seed, traj_offset, first_noise_block);
    CUDA_CHECK_LAST_ERROR();
}

void launch_batched_init(cuDoubleComplex* states, int n, int64_t batch, int num_sms, cudaStream_t stream) {
    const uint64_t total = (uint64_t)batch << n;
    batched_init_kernel<<<num_sms * 8, 256, 0, stream>>>(states, ((uint64_t)1 << n) - 1, total);
    CUDA_CHECK_LAST_ERROR();
}

void launch_batched_average(const cuDoubleComplex* states, int n, int64_t batch, double* d_avg, int num_sms,
                            cudaStream_t stream) {
    const uint32_t size = 1u << n;
    CUDA_CHECK(cudaMemsetAsync(d_avg, 0, size * sizeof(double), stream));
    const unsigned gx = (size + 255) / 256;
    unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)(num_sms * 8) / (int64_t)gx + 1));
    batched_average_kernel<<<dim3(gx, gy), 256, 0, stream>>>(states, n, batch, 1.0 / (double)batch, d_avg);
    CUDA_CHECK_LAST_ERROR();
}

void launch_batched_sample(const cuDoubleComplex* states, int n, int64_t batch, const double* d_uniforms, int n_shots,
                           int32_t* d_out, int num_sms, cudaStream_t stream) {
    const size_t per_warp = (size_t)8 << n;
    int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / per_warp));
    const size_t smem = per_warp * warps;
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(batched_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t grid = std::min<int64_t>((batch + warps - 1) / warps, (int64_t)num_sms * 4);
    batched_sample_kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(states, n, batch, d_uniforms, n_shots, d_out);
    CUDA_CHECK_LAST_ERROR();
}

void launch_histogram(const int32_t* d_samples, int64_t count, int n, int32_t* d_hist, int num_sms, cudaStream_t stream) {
    histogram_kernel<<<num_sms * 4, 256, 0, stream>>>(d_samples, count, 1u << n, d_hist);
    CUDA_CHECK_LAST_ERROR();
}

}  // namespace b200
}  // namespace qsim
