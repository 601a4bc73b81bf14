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

#include <algorithm>
#include <cstdlib>
#include <stdexcept>

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

// ---- a RUN of damping events (amplitude / phase damping on any qubits) in two sweeps instead of 1.5 sweeps and four
// barriers each.  While no event jumps, the run is a product of diagonal Kraus operators K0 = diag(1, sqrt(1-g)): with
// the unnormalised state a^(j) before event j, N_j = |a^(j)|^2 and S_j = the part of it with the event's bit set,
//     jump_j  <=>  u_j < g_j S_j / N_j,       N_{j+1} = N_j - g_j S_j,
// so ONE sweep yields every S_j (weighted sums: the weight of amplitude i for event j is |a_i|^2 times the (1 - g_j') of
// the earlier events whose bit it has), one block reduction adds them up, every thread replays the decisions, and ONE
// sweep applies the surviving prefix of the run, normalised.  A jump (rare: probability g S) ends the prefix; that
// event takes the one-event path above and the rest of the run starts over.
// Thread t owns amplitudes t + 256 s: index bits < 8 are thread bits (one value for all of the thread's amplitudes: scalar
// work), bits >= 8 select the slot s.
constexpr int kDampRun = 16;   // events per run (a run is cut when it gets longer)

template <int NS>
__device__ void damping_run(double2* s, int n, const TrajEvent* __restrict__ ev, int k, const double* u0, double* red,
                            double* stot) {
    const uint32_t size = 1u << n, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    while (k > 0) {
        double T[NS];
        double tot = 0.0;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const uint32_t i = tid + (uint32_t)sl * kTrajThreads;
            double t = 0.0;
            if (i < size) { const double2 a = s[i]; t = a.x * a.x + a.y * a.y; }
            T[sl] = t;
            tot += t;
        }
        double S[kDampRun + 1];
        S[kDampRun] = tot;        // N_0
        double f = 1.0;           // product of the thread-bit factors so far
#pragma unroll
        for (int j = 0; j < kDampRun; ++j) {
            S[j] = 0.0;
            if (j < k) {
                const int q = ev[j].qubit;
                const double g = ev[j].p, keep = 1.0 - g;
                if (q < 8) {
                    if ((tid >> q) & 1u) { S[j] = f * tot; f *= keep; }
                } else {
                    const int b = q - 8;
                    double part = 0.0;
#pragma unroll
                    for (int sl = 0; sl < NS; ++sl)
                        if ((sl >> b) & 1) { part += T[sl]; T[sl] *= keep; }
                    S[j] = f * part;
                    tot -= g * part;
                }
            }
        }
        // block reduction of S[0..k) and N_0: butterflies inside the warps, then a fixed-order sum over the warps
#pragma unroll
        for (int j = 0; j <= kDampRun; ++j) {
            if (j < k || j == kDampRun) {
                double v = S[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp * (kDampRun + 1) + j] = v;
            }
        }
        __syncthreads();
        if (tid <= (uint32_t)kDampRun && ((int)tid < k || tid == (uint32_t)kDampRun)) {
            double v = 0.0;
            for (int w = 0; w < kTrajThreads / 32; ++w) v += red[w * (kDampRun + 1) + tid];
            stot[tid] = v;
        }
        __syncthreads();
        // decisions, replayed by every thread on the same numbers
        double N = stot[kDampRun];
        int kk = k;               // events of the run that pass without a jump
        for (int j = 0; j < k; ++j) {
            const double gS = ev[j].p * stot[j];
            if (u0[j] * N < gS) { kk = j; break; }
            N -= gS;
        }
        // the prefix [0, kk): amplitude *= prod sqrt(1 - g_j) over its set bits, / sqrt(N)
        if (kk > 0) {
            double fthr = 1.0 / sqrt(N);
            double R[NS];
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) R[sl] = 1.0;
            for (int j = 0; j < kk; ++j) {
                const int q = ev[j].qubit;
                const double r = ev[j].sqrt_keep;
                if (q < 8) { if ((tid >> q) & 1u) fthr *= r; }
                else {
                    const int b = q - 8;
#pragma unroll
                    for (int sl = 0; sl < NS; ++sl)
                        if ((sl >> b) & 1) R[sl] *= r;
                }
            }
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                const uint32_t i = tid + (uint32_t)sl * kTrajThreads;
                if (i < size) { const double w = fthr * R[sl]; const double2 a = s[i]; s[i] = make_double2(a.x * w, a.y * w); }
            }
        }
        __syncthreads();
        if (kk == k) break;
        // event kk jumps (or sits on the threshold): the one-event path decides again on the renormalised state
        noise_event(s, n, ev[kk], u0[kk], 0.0, red);
        ev += kk + 1;
        u0 += kk + 1;
        k -= kk + 1;
    }
}

// event id of the Philox counter: (noise block << 16) | event index within the block - independent of how many events a
// block has, so run() (whole model) and applyNoise (a few events) never reuse a counter
__device__ __forceinline__ uint64_t event_id(uint64_t block, int e) { return (block << 16) | (uint64_t)(uint32_t)e; }

constexpr int kUniformChunk = kTrajThreads;   // uniforms are drawn kTrajThreads events at a time, one event per thread

template <int NS, int MINB>
__global__ void __launch_bounds__(kTrajThreads, MINB) trajectory_kernel(cuDoubleComplex* __restrict__ states, int n, int64_t batch,
                                                                  const TrajItem* __restrict__ items, int n_items,
                                                                  const TrajEvent* __restrict__ events, int n_events,
                                                                  uint32_t seed, uint64_t traj_offset,
                                                                  uint64_t first_noise_block, double* __restrict__ avg,
                                                                  double inv_batch) {
    extern __shared__ __align__(16) unsigned char traj_smem[];
    double2* s = reinterpret_cast<double2*>(traj_smem);
    const uint32_t size = 1u << n;
    double* red = reinterpret_cast<double*>(s + size);                       // kTrajThreads / 32 * (kDampRun + 1), >= 40
    double* stot = red + (kTrajThreads / 32) * (kDampRun + 1);               // kDampRun + 2
    double* u0s = stot + kDampRun + 2;                                       // kUniformChunk
    double* u1s = u0s + kUniformChunk;                                       // kUniformChunk
    const uint32_t tid = threadIdx.x;
    double acc[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) acc[sl] = 0.0;
    for (int64_t traj = blockIdx.x; traj < batch; traj += gridDim.x) {
        double2* g = reinterpret_cast<double2*>(states) + (size_t)traj * size;
        for (uint32_t i = tid; i < size; i += kTrajThreads) s[i] = g[i];
        __syncthreads();
        uint64_t block = first_noise_block;
        for (int it = 0; it < n_items; ++it) {
            const TrajItem& item = items[it];
            if (item.kind == 0) {
                apply_2x2(s, n, item.target, (uint32_t)item.cmask, (uint32_t)item.cval, item.m);
                continue;
            }
            for (int base = 0; base < n_events; base += kUniformChunk) {
                const int cnt = (n_events - base) < kUniformChunk ? (n_events - base) : kUniformChunk;
                if ((int)tid < cnt) {     // one Philox call per event, spread over the threads
                    double a, b;
                    traj_uniforms(seed, (uint64_t)traj + traj_offset, event_id(block, base + (int)tid), a, b);
                    u0s[tid] = a;
                    u1s[tid] = b;
                }
                __syncthreads();
                int e = 0;
                while (e < cnt) {
                    const TrajEvent& ev = events[base + e];
                    if (ev.type == 1 || ev.type == 2) {
                        int k = 1;
                        while (e + k < cnt && k < kDampRun && (events[base + e + k].type == 1 || events[base + e + k].type == 2)) ++k;
                        damping_run<NS>(s, n, events + base + e, k, u0s + e, red, stot);
                        e += k;
                    } else {
                        noise_event(s, n, ev, u0s[e], u1s[e], red);
                        ++e;
                    }
                }
                __syncthreads();          // the uniforms are overwritten by the next chunk / block
            }
            ++block;
        }
        for (uint32_t i = tid; i < size; i += kTrajThreads) g[i] = s[i];
        if (avg) {
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                const uint32_t i = tid + (uint32_t)sl * kTrajThreads;
                if (i < size) { const double2 a = s[i]; acc[sl] += (a.x * a.x + a.y * a.y) * inv_batch; }
            }
        }
        __syncthreads();
    }
    if (avg) {   // every thread owns the same basis states for all of the CTA's trajectories: one atomic per state and CTA
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const uint32_t i = tid + (uint32_t)sl * kTrajThreads;
            if (i < size && acc[sl] != 0.0) atomicAdd(&avg[i], acc[sl]);
        }
    }
}

__global__ void batched_init_kernel(cuDoubleComplex* states, uint64_t size_mask, uint64_t total) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        states[i] = make_cuDoubleComplex((i & size_mask) == 0 ? 1.0 : 0.0, 0.0);
}

// (fallback when no run() has filled the average in its epilogue) each block owns a slice of trajectories; its threads walk
// whole 2^n rows coalesced and keep one partial sum per owned basis state, then one atomic per basis state and block
__global__ void batched_average_kernel(const cuDoubleComplex* __restrict__ states, int n, int64_t batch, double inv_batch,
                                       double* __restrict__ avg) {
    const uint32_t size = 1u << n;
    const int64_t per = (batch + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = (t0 + per < batch) ? t0 + per : batch;
    for (uint32_t i0 = 0; i0 < size; i0 += blockDim.x * 4) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t t = t0; t < t1; ++t) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t i = i0 + r * blockDim.x + threadIdx.x;
                if (i < size) { const cuDoubleComplex a = states[(size_t)t * size + i]; acc[r] += (a.x * a.x + a.y * a.y) * inv_batch; }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t i = i0 + r * blockDim.x + threadIdx.x;
            if (i < size && t1 > t0) atomicAdd(&avg[i], acc[r]);
        }
    }
}

__device__ __forceinline__ uint32_t pad(uint32_t i) { return i + (i >> 4); }

// Sampling: one CTA per trajectory.  The reference draws from the SEQUENTIAL fp64 prefix sums of the probabilities
// (std::partial_sum + lower_bound, src/NoiseModel.cu:938-957).  A sum of m non-negative terms differs from the exact sum by
// at most m * 2^-53 * total whatever the order, so a block-wide tree-order scan (approximate CDF c~) decides every shot whose
// distance to the neighbouring c~ values exceeds tau = 2^-38 * total (> 2 * 8192 * 2^-53): there the sequential sums C
// satisfy C[k-1] < u <= C[k] as well.  The (rare) shots inside the margin replay the sequential sum from shared memory, so the
// result is the reference's index in every case.
__global__ void __launch_bounds__(kTrajThreads) batched_sample_kernel(const cuDoubleComplex* __restrict__ states, int n, int64_t batch,
                                                                      const double* __restrict__ uniforms, int n_shots,
                                                                      int32_t* __restrict__ out, int32_t* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char samp_smem[];
    const uint32_t size = 1u << n, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // (both arrays are indexed through pad(): one spare double per 16, so that the threads' contiguous 16-element segments
    //  start in different banks - without it the segment scans are 32-way bank conflicts)
    double* p = reinterpret_cast<double*>(samp_smem);   // probabilities, rounded as std::norm
    double* c = p + pad(size);                          // tree-order inclusive prefix sums
    __shared__ double warp_tot[kTrajThreads / 32];
    const uint32_t per = size >= (uint32_t)kTrajThreads ? size / kTrajThreads : 1u;
    for (int64_t traj = blockIdx.x; traj < batch; traj += gridDim.x) {
        const cuDoubleComplex* a = states + (size_t)traj * size;
        for (uint32_t i = tid; i < size; i += kTrajThreads) {
            const cuDoubleComplex v = a[i];
            p[pad(i)] = __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
        }
        __syncthreads();
        // thread t scans its contiguous segment, then the segment totals are scanned across the block
        const uint32_t i0 = tid * per;
        double run = 0.0;
        if (i0 < size)
            for (uint32_t j = 0; j < per; ++j) { run = __dadd_rn(run, p[pad(i0 + j)]); c[pad(i0 + j)] = run; }
        double incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        double offset = incl - run;
        for (int w = 0; w < warp; ++w) offset += warp_tot[w];
        double total = 0.0;
        for (int w = 0; w < kTrajThreads / 32; ++w) total += warp_tot[w];
        if (i0 < size && offset != 0.0)
            for (uint32_t j = 0; j < per; ++j) c[pad(i0 + j)] += offset;
        __syncthreads();
        const double tau = total * 0x1.0p-38;
        for (int shot = tid; shot < n_shots; shot += kTrajThreads) {
            const double r = uniforms[(size_t)traj * n_shots + shot];
            uint32_t lo = 0, hi = size;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (c[pad(mid)] >= r) hi = mid; else lo = mid + 1; }
            uint32_t k = lo;
            bool sure;
            if (k == 0) sure = (p[0] >= r);                               // C[0] = p[0] exactly
            else if (k == size) sure = (r - c[pad(size - 1)] > tau);
            else sure = (c[pad(k)] - r > tau) && (r - c[pad(k - 1)] > tau);
            if (!sure) {                                                  // the reference's own loop
                double C = 0.0;
                k = size;
                for (uint32_t i = 0; i < size; ++i) {
                    C = __dadd_rn(C, p[pad(i)]);
                    if (C >= r) { k = i; break; }
                }
            }
            if (out) out[(size_t)shot * batch + traj] = (int32_t)k;
            if (hist && k < size) atomicAdd(&hist[k], 1);
        }
        __syncthreads();
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

namespace {
template <int NS>
void launch_traj_ns(cuDoubleComplex* states, int n, int64_t batch, const TrajItem* d_items, int n_items, const TrajEvent* d_events,
                    int n_events, uint32_t seed, uint64_t traj_offset, uint64_t first_noise_block, double* d_avg, int num_sms,
                    cudaStream_t stream) {
    const size_t smem = ((size_t)16 << n) +
                        ((kTrajThreads / 32) * (kDampRun + 1) + (kDampRun + 2) + 2 * kUniformChunk) * sizeof(double);
    // resident CTAs per SM: what shared memory allows (3 at 12 qubits), unless the register budget that goes with it costs more
    // in spills than the occupancy brings (NS >= 16: 2 CTAs at 128 registers; QSIM_TRAJ_OCC=2|3 overrides, for measurements)
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(224 * 1024) / (smem + 1024)));
    int minb = NS >= 16 ? 2 : 3;
    if (const char* e = std::getenv("QSIM_TRAJ_OCC")) minb = std::atoi(e) >= 3 ? 3 : 2;
    if (minb == 2) per_sm = std::min(per_sm, 2);
    const int64_t grid = std::min<int64_t>(batch, (int64_t)num_sms * per_sm);
    if (minb == 3) {
        CUDA_CHECK(cudaFuncSetAttribute(trajectory_kernel<NS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trajectory_kernel<NS, 3><<<(unsigned)grid, kTrajThreads, smem, stream>>>(states, n, batch, d_items, n_items, d_events, n_events,
                                                                                seed, traj_offset, first_noise_block, d_avg,
                                                                                1.0 / (double)batch);
    } else {
        CUDA_CHECK(cudaFuncSetAttribute(trajectory_kernel<NS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        trajectory_kernel<NS, 2><<<(unsigned)grid, kTrajThreads, smem, stream>>>(states, n, batch, d_items, n_items, d_events, n_events,
                                                                                seed, traj_offset, first_noise_block, d_avg,
                                                                                1.0 / (double)batch);
    }
    CUDA_CHECK_LAST_ERROR();
}
}  // namespace

void launch_trajectories(cuDoubleComplex* states, int n, int64_t batch, const TrajItem* d_items, int n_items,
                         const TrajEvent* d_events, int n_events, uint32_t seed, uint64_t traj_offset,
                         uint64_t first_noise_block, int num_sms, cudaStream_t stream, double* d_avg) {
    if (n_events >= 65536) throw std::invalid_argument("too many noise events per gate (at most 65535)");
    if (d_avg) CUDA_CHECK(cudaMemsetAsync(d_avg, 0, sizeof(double) << n, stream));
#define QSIM_TRAJ_CASE(NS) launch_traj_ns<NS>(states, n, batch, d_items, n_items, d_events, n_events, seed, traj_offset, first_noise_block, d_avg, num_sms, stream)
    switch (n <= 8 ? 0 : n - 8) {
        case 0: QSIM_TRAJ_CASE(1); break;
        case 1: QSIM_TRAJ_CASE(2); break;
        case 2: QSIM_TRAJ_CASE(4); break;
        case 3: QSIM_TRAJ_CASE(8); break;
        case 4: QSIM_TRAJ_CASE(16); break;
        default: QSIM_TRAJ_CASE(32); break;
    }
#undef QSIM_TRAJ_CASE
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
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)num_sms * 8));
    batched_average_kernel<<<grid, 256, 0, stream>>>(states, n, batch, 1.0 / (double)batch, d_avg);
    CUDA_CHECK_LAST_ERROR();
}

void launch_batched_sample(const cuDoubleComplex* states, int n, int64_t batch, const double* d_uniforms, int n_shots,
                           int32_t* d_out, int32_t* d_hist, int num_sms, cudaStream_t stream) {
    const size_t smem = 2 * sizeof(double) * (((size_t)1 << n) + ((size_t)1 << n) / 16 + 2);
    CUDA_CHECK(cudaFuncSetAttribute(batched_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(224 * 1024) / (smem + 1024)));
    const int64_t grid = std::min<int64_t>(batch, (int64_t)num_sms * per_sm);
    batched_sample_kernel<<<(unsigned)grid, kTrajThreads, smem, stream>>>(states, n, batch, d_uniforms, n_shots, d_out, d_hist);
    CUDA_CHECK_LAST_ERROR();
}

void launch_histogram(const int32_t* d_samples, int64_t count, int n, int32_t* d_hist, int num_sms, cudaStream_t stream) {
    histogram_kernel<<<num_sms * 4, 256, 0, stream>>>(d_samples, count, 1u << n, d_hist);
    CUDA_CHECK_LAST_ERROR();
}

}  // namespace b200
}  // namespace qsim
