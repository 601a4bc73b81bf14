// Read-out kernels: probabilities, reductions, measurement collapse and *bit-exact* sampling.
//
// The reference computes its CDF on the host with a sequential fp64 std::partial_sum over 2^n
// probabilities and answers each shot with std::lower_bound (src/Simulator.cu:164-185,
// src/StateVector.cu:316-342); measurement sums the masked probabilities sequentially on the host
// too (src/StateVector.cu:280-287).  A parallel scan rounds differently, so shots that land within
// rounding distance of a CDF step would come out different.  SequentialCdf reproduces the
// *sequential* rounding exactly, in parallel, without ever materialising 2^n probabilities:
//
//   While the running sum c stays inside one binade [2^e, 2^(e+1)) it is a multiple of
//   u = 2^(e-52), and fl(c + p) = c + rn_u(p), where rn_u(p) (p rounded to a multiple of u) does
//   not depend on c — except for exact ties.  So for a chunk of 4096 probabilities known (from an
//   approximate prefix sum with a rigorous margin) to stay inside binade e, every thread can add
//   its 16 probabilities to the surrogate start 2^e instead of the unknown c: the increment
//   comes out identical, and increments are multiples of u that add exactly.  Chunks that may
//   cross a binade boundary, start at zero, or contain a tie are flagged and replayed
//   sequentially from the exact start by one warp (a few dozen chunks per sweep).
//
// p_i is computed as fl(fl(re*re) + fl(im*im)) — std::norm's rounding on the reference CPU path
// (src/Simulator.cu:319-325) — with explicit __dmul_rn/__dadd_rn so nvcc cannot contract it.
#include "readout.cuh"

#include <stdexcept>

#include "qsim/constants.hpp"

namespace qsim {
namespace b200 {

namespace {

constexpr int kBlock = 256;
constexpr double kMargin = 1.9073486328125e-06;  // 2^-19: twice the (N-1)*2^-53 bound (N <= 2^33) on |sequential - exact| / sum

__device__ __forceinline__ double prob_of(const cuDoubleComplex a) {
    return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));
}

// mask_bit < 0: every index counts.  Otherwise bits 0-7 hold an index bit position and bit 8 the value that
// bit must have for the amplitude to count (0: "qubit reads 0", the measurement sums of the reference).
__device__ __forceinline__ double masked_prob(const cuDoubleComplex* __restrict__ state, uint64_t i, int mask_bit) {
    if (mask_bit >= 0 && (int)((i >> (mask_bit & 0xff)) & 1) != ((mask_bit >> 8) & 1)) return 0.0;
    return prob_of(state[i]);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < blockDim.x / 32) ? red[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// ---- plain read-out ----------------------------------------------------------------------------

__global__ void probabilities_range_kernel(const cuDoubleComplex* __restrict__ state, double* __restrict__ out,
                                           uint64_t first, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = prob_of(state[first + i]);
}

__global__ void set_basis_kernel(cuDoubleComplex* state, uint64_t idx) { state[idx] = make_cuDoubleComplex(1.0, 0.0); }

// Deterministic two-level reduction of sum |a_i|^2 over indices with bit `mask_bit` == 0.
__global__ void partial_prob_kernel(const cuDoubleComplex* __restrict__ state, uint64_t n, int mask_bit,
                                    double* __restrict__ partial) {
    __shared__ double red[32];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc = __dadd_rn(acc, masked_prob(state, i, mask_bit));
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void final_sum_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc = __dadd_rn(acc, partial[i]);
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) *out = s;
}

__global__ void collapse_kernel(cuDoubleComplex* __restrict__ state, uint64_t n, int bit, int outcome, double scale) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        cuDoubleComplex a = state[i];
        if ((int)((i >> bit) & 1) != outcome) a = make_cuDoubleComplex(0.0, 0.0);
        else a = make_cuDoubleComplex(__dmul_rn(a.x, scale), __dmul_rn(a.y, scale));
        state[i] = a;
    }
}

// ---- exact sequential CDF ------------------------------------------------------------------------

// K1: approximate (tree-order) sum of each chunk.
__global__ void chunk_approx_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk,
                                    double* __restrict__ approx) {
    __shared__ double red[32];
    const uint64_t base = (uint64_t)blockIdx.x * chunk;
    double acc = 0.0;
    for (int j = threadIdx.x; j < chunk; j += blockDim.x) acc = __dadd_rn(acc, masked_prob(state, base + j, mask_bit));
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) approx[blockIdx.x] = s;
}

// K2: single-block exclusive scan of the approximate chunk sums: lo[k] = approx start of chunk k.
__global__ void chunk_scan_kernel(const double* __restrict__ approx, uint64_t m, double* __restrict__ lo, double c_init) {
    __shared__ double part[1024];
    const uint64_t per = (m + blockDim.x - 1) / blockDim.x;
    const uint64_t b = (uint64_t)threadIdx.x * per, e = (b + per < m) ? b + per : m;
    double acc = 0.0;
    for (uint64_t i = b; i < e; ++i) acc += approx[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = c_init;
        for (int i = 0; i < (int)blockDim.x; ++i) { double v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    double run = part[threadIdx.x];
    for (uint64_t i = b; i < e; ++i) { lo[i] = run; run += approx[i]; }
}

enum : uint8_t { CH_ZERO = 0, CH_FAST = 1, CH_SLOW = 2 };

// K3: per chunk, the sequential-order increment computed from the surrogate start 2^e.
__global__ void chunk_surrogate_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk,
                                       const double* __restrict__ approx, const double* __restrict__ lo,
                                       double* __restrict__ delta, double* __restrict__ bin_base,
                                       uint8_t* __restrict__ flag) {
    __shared__ double ps[4096 + 256];
    __shared__ double red[32];
    __shared__ int tie_any;
    const uint64_t k = blockIdx.x;
    const double a_lo = lo[k], a_sum = approx[k], a_hi = a_lo + a_sum;
    // classification (uniform across the block)
    uint8_t f = CH_SLOW;
    double base = 0.0;
    if (a_sum == 0.0) f = CH_ZERO;  // a tree sum of non-negative terms is 0 only if every term is 0: c unchanged
    else if (chunk == 4096 && a_lo > 1e-290) {
        int ex;
        frexp(a_lo, &ex);           // a_lo = m * 2^ex, m in [0.5, 1)
        base = ldexp(1.0, ex - 1);  // 2^e with 2^e <= a_lo < 2^(e+1)
        if (a_lo * (1.0 - kMargin) >= base && a_hi * (1.0 + kMargin) < 2.0 * base) f = CH_FAST;
    }
    if (f != CH_FAST) {
        if (threadIdx.x == 0) { flag[k] = f; delta[k] = 0.0; bin_base[k] = 0.0; }
        return;
    }
    if (threadIdx.x == 0) tie_any = 0;
    const uint64_t g0 = k * (uint64_t)chunk;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int i = threadIdx.x + j * kBlock;
        ps[i + (i >> 4)] = masked_prob(state, g0 + i, mask_bit);
    }
    __syncthreads();
    const double half_u = ldexp(base, -53);   // u/2 with u = 2^(e-52)
    double s = base;
    bool tie = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const double x = ps[17 * threadIdx.x + j];
        const double t = __dadd_rn(s, x);
        const double err = __dsub_rn(x, __dsub_rn(t, s));   // exact: s >= x (Fast2Sum)
        tie |= (fabs(err) == half_u);
        s = t;
    }
    if (tie) atomicOr(&tie_any, 1);
    const double d = block_sum(__dsub_rn(s, base), red);   // multiples of u below 2^(e+1): exact adds
    if (threadIdx.x == 0) {
        flag[k] = tie_any ? CH_SLOW : CH_FAST;
        delta[k] = d;
        bin_base[k] = base;
    }
}

// Sequential replay of one chunk by a full warp (all lanes carry the same running sum).
__device__ __forceinline__ double replay_chunk(const cuDoubleComplex* __restrict__ state, int mask_bit, uint64_t g0,
                                               int chunk, double c) {
    const int lane = threadIdx.x & 31;
    for (int g = 0; g < chunk; g += 32) {
        const double p = (g + lane < chunk) ? masked_prob(state, g0 + g + lane, mask_bit) : 0.0;
        const int lim = (chunk - g) < 32 ? (chunk - g) : 32;
        for (int j = 0; j < lim; ++j) c = __dadd_rn(c, __shfl_sync(0xffffffffu, p, j));
    }
    return c;
}

// K4: the exact running sum at every chunk start, in three launches.  The chunks are taken in groups of 32.
//  (a) one warp per group, whole grid: a group whose non-zero chunks are all FAST in one binade has increments that are
//      multiples of one ulp — they add exactly in any order, so a warp scan gives the group's total;
//  (b) one warp stitches the group totals in order — one exact addition and one binade check per group — and replays,
//      chunk by chunk, the few groups that need it;
//  (c) one warp per group, whole grid: per-chunk starts = group start + exact in-group prefix.
enum : uint8_t { G_ZERO = 0, G_SIMPLE = 1, G_COMPLEX = 2, G_DONE = 3 };

// in-group scan shared by (a) and (c): returns the group's kind, its total and the lane's exclusive prefix
__device__ __forceinline__ uint8_t group_scan(const double* __restrict__ delta, const double* __restrict__ bin_base,
                                              const uint8_t* __restrict__ flag, uint64_t m, uint64_t group, double& total,
                                              double& bb, double& excl) {
    const int lane = threadIdx.x & 31;
    const uint64_t kk = group * 32 + lane;
    const double d = kk < m ? delta[kk] : 0.0;
    const double b = kk < m ? bin_base[kk] : 0.0;
    const int f = kk < m ? (int)flag[kk] : (int)CH_ZERO;
    const unsigned fast_mask = __ballot_sync(0xffffffffu, f == (int)CH_FAST);
    const unsigned slow_mask = __ballot_sync(0xffffffffu, f == (int)CH_SLOW);
    total = 0.0; bb = 0.0; excl = 0.0;
    if (slow_mask != 0u) return G_COMPLEX;
    if (fast_mask == 0u) return G_ZERO;
    bb = __shfl_sync(0xffffffffu, b, __ffs(fast_mask) - 1);
    const bool is_fast = (fast_mask >> lane) & 1u;
    if (!__all_sync(0xffffffffu, !is_fast || b == bb)) return G_COMPLEX;
    const double dd = is_fast ? d : 0.0;
    double incl = dd;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = __dadd_rn(incl, up);
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    excl = __dsub_rn(incl, dd);
    return G_SIMPLE;
}

__global__ void group_summary_kernel(const double* __restrict__ delta, const double* __restrict__ bin_base,
                                     const uint8_t* __restrict__ flag, uint64_t m, uint64_t n_groups,
                                     double* __restrict__ g_total, double* __restrict__ g_bb, uint8_t* __restrict__ g_kind) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < n_groups; g += n_warps) {
        double total, bb, excl;
        const uint8_t kind = group_scan(delta, bin_base, flag, m, g, total, bb, excl);
        if ((threadIdx.x & 31) == 0) { g_kind[g] = kind; g_total[g] = total; g_bb[g] = bb; }
    }
}

__global__ void group_stitch_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk, uint64_t m,
                                    uint64_t n_groups, const double* __restrict__ delta,
                                    const double* __restrict__ bin_base, const uint8_t* __restrict__ flag,
                                    const double* __restrict__ g_total, const double* __restrict__ g_bb,
                                    uint8_t* __restrict__ g_kind, double* __restrict__ g_start,
                                    double* __restrict__ start, unsigned long long* __restrict__ n_slow, double c_init) {
    const int lane = threadIdx.x & 31;
    double c = c_init;
    unsigned long long slow = 0;
    // lane l holds the summary of group g0 + l; the next 32 are fetched while these are stitched
    uint64_t gl = lane;
    double t_nxt = gl < n_groups ? g_total[gl] : 0.0, b_nxt = gl < n_groups ? g_bb[gl] : 0.0;
    int k_nxt = gl < n_groups ? (int)g_kind[gl] : (int)G_ZERO;
    for (uint64_t g0 = 0; g0 < n_groups; g0 += 32) {
        const double t_cur = t_nxt, b_cur = b_nxt;
        const int k_cur = k_nxt;
        gl = g0 + 32 + lane;
        t_nxt = gl < n_groups ? g_total[gl] : 0.0;
        b_nxt = gl < n_groups ? g_bb[gl] : 0.0;
        k_nxt = gl < n_groups ? (int)g_kind[gl] : (int)G_ZERO;
        const int cnt = (n_groups - g0) < 32 ? (int)(n_groups - g0) : 32;
        double my_start = 0.0;
        bool my_done = false;
        for (int i = 0; i < cnt; ++i) {
            const int kind = __shfl_sync(0xffffffffu, k_cur, i);
            if (kind == G_ZERO) { if (lane == i) my_start = c; continue; }
            if (kind == G_SIMPLE) {
                const double bb = __shfl_sync(0xffffffffu, b_cur, i);
                const double c_end = __dadd_rn(c, __shfl_sync(0xffffffffu, t_cur, i));
                if (c >= bb && c_end < 2.0 * bb) { if (lane == i) my_start = c; c = c_end; continue; }   // exact binade check
            }
            // chunk by chunk
            const uint64_t k0 = (g0 + i) * 32, kk = k0 + lane;
            const double d = kk < m ? delta[kk] : 0.0;
            const double b = kk < m ? bin_base[kk] : 0.0;
            const int f = kk < m ? (int)flag[kk] : (int)CH_ZERO;
            const int lim = (m - k0) < 32 ? (int)(m - k0) : 32;
            for (int j = 0; j < lim; ++j) {
                const double dj = __shfl_sync(0xffffffffu, d, j), bj = __shfl_sync(0xffffffffu, b, j);
                const int fj = __shfl_sync(0xffffffffu, f, j);
                if (lane == 0) start[k0 + j] = c;
                if (fj == CH_ZERO) continue;
                const double cc = __dadd_rn(c, dj);
                if (fj == CH_FAST && c >= bj && cc < 2.0 * bj) c = cc;   // exact binade check on the true values
                else { c = replay_chunk(state, mask_bit, (k0 + j) * (uint64_t)chunk, chunk, c); ++slow; }
            }
            if (lane == i) my_done = true;
        }
        if (lane < cnt) {
            g_start[g0 + lane] = my_start;
            if (my_done) g_kind[g0 + lane] = G_DONE;
        }
    }
    if (lane == 0) { start[m] = c; *n_slow = slow; }
}

__global__ void group_write_kernel(const double* __restrict__ delta, const double* __restrict__ bin_base,
                                   const uint8_t* __restrict__ flag, uint64_t m, uint64_t n_groups,
                                   const uint8_t* __restrict__ g_kind, const double* __restrict__ g_start,
                                   double* __restrict__ start) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < n_groups; g += n_warps) {
        if (g_kind[g] == G_DONE) continue;
        double total, bb, excl;
        group_scan(delta, bin_base, flag, m, g, total, bb, excl);
        const uint64_t kk = g * 32 + (threadIdx.x & 31);
        if (kk < m) start[kk] = __dadd_rn(g_start[g], excl);
    }
}

// K5: one warp per shot: binary search over the exact chunk ends, then replay inside the chunk.
__global__ void sample_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk, uint64_t m,
                              const double* __restrict__ start, const double* __restrict__ uniforms, int64_t n_shots,
                              int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t shot = warp; shot < n_shots; shot += n_warps) {
        const double r = uniforms[shot];
        // smallest k in [0, m) with start[k+1] >= r
        uint64_t lo_k = 0, hi_k = m;
        while (lo_k < hi_k) {
            const uint64_t mid = (lo_k + hi_k) >> 1;
            if (start[mid + 1] >= r) hi_k = mid; else lo_k = mid + 1;
        }
        int64_t result = (int64_t)(m * (uint64_t)chunk);   // past the end: the reference returns 2^n here
        if (lo_k < m) {
            double c = start[lo_k];
            const uint64_t g0 = lo_k * (uint64_t)chunk;
            bool found = false;
            for (int g = 0; g < chunk && !found; g += 32) {
                const double p = (g + lane < chunk) ? masked_prob(state, g0 + g + lane, mask_bit) : 0.0;
                const int lim = (chunk - g) < 32 ? (chunk - g) : 32;
                for (int j = 0; j < lim; ++j) {
                    c = __dadd_rn(c, __shfl_sync(0xffffffffu, p, j));
                    if (c >= r) { result = (int64_t)(g0 + g + j); found = true; break; }
                }
            }
        }
        if (lane == 0) out[shot] = result;
    }
}

int grid_for(uint64_t n, int num_sms) {
    uint64_t blocks = (n + kBlock - 1) / kBlock;
    uint64_t cap = (uint64_t)num_sms * 8;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

// ---- host wrappers -------------------------------------------------------------------------------

void launch_probabilities(const cuDoubleComplex* state, double* out, uint64_t first, uint64_t count, int num_sms,
                          cudaStream_t stream) {
    if (!count) return;
    probabilities_range_kernel<<<grid_for(count, num_sms), kBlock, 0, stream>>>(state, out, first, count);
    CUDA_CHECK_LAST_ERROR();
}

void launch_init_basis(cuDoubleComplex* state, uint64_t n, uint64_t idx, cudaStream_t stream) {
    CUDA_CHECK(cudaMemsetAsync(state, 0, n * sizeof(cuDoubleComplex), stream));
    set_basis_kernel<<<1, 1, 0, stream>>>(state, idx);
    CUDA_CHECK_LAST_ERROR();
}

double reduce_probability(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng) {
    cudaStream_t stream = eng.stream();
    const int grid = grid_for(n, eng.numSMs());
    double* partial = static_cast<double*>(eng.scratch(0, ((size_t)grid + 1) * sizeof(double)));
    partial_prob_kernel<<<grid, kBlock, 0, stream>>>(state, n, mask_bit, partial);
    CUDA_CHECK_LAST_ERROR();
    final_sum_kernel<<<1, kBlock, 0, stream>>>(partial, grid, partial + grid);
    CUDA_CHECK_LAST_ERROR();
    double out = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&out, partial + grid, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return out;
}

void launch_collapse(cuDoubleComplex* state, uint64_t n, int bit, int outcome, double scale, int num_sms,
                     cudaStream_t stream) {
    collapse_kernel<<<grid_for(n, num_sms), kBlock, 0, stream>>>(state, n, bit, outcome, scale);
    CUDA_CHECK_LAST_ERROR();
}

SequentialCdf::SequentialCdf(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng, double c_init)
    : state_(state), n_(n), mask_bit_(mask_bit), stream_(eng.stream()), eng_(eng) {
    chunk_ = n >= 4096 ? 4096 : (int)n;
    m_ = n / (uint64_t)chunk_;
    const size_t m8 = (m_ + 2) * sizeof(double);
    unsigned char* arena = static_cast<unsigned char*>(eng.scratch(0, 5 * m8 + 16 + m_ + 64));
    approx_ = reinterpret_cast<double*>(arena);
    lo_ = reinterpret_cast<double*>(arena + m8);
    delta_ = reinterpret_cast<double*>(arena + 2 * m8);
    base_ = reinterpret_cast<double*>(arena + 3 * m8);
    start_ = reinterpret_cast<double*>(arena + 4 * m8);
    slow_ = reinterpret_cast<unsigned long long*>(arena + 5 * m8);
    flag_ = arena + 5 * m8 + 16;
    cudaStream_t stream = stream_;
    chunk_approx_kernel<<<(unsigned)m_, kBlock, 0, stream>>>(state, mask_bit, chunk_, approx_);
    CUDA_CHECK_LAST_ERROR();
    chunk_scan_kernel<<<1, 1024, 0, stream>>>(approx_, m_, lo_, c_init);
    CUDA_CHECK_LAST_ERROR();
    chunk_surrogate_kernel<<<(unsigned)m_, kBlock, 0, stream>>>(state, mask_bit, chunk_, approx_, lo_, delta_, base_, flag_);
    CUDA_CHECK_LAST_ERROR();
    // the approximate sums and lower bounds are dead now: their storage holds the group summaries
    const uint64_t n_groups = (m_ + 31) / 32;
    double *g_total = approx_, *g_bb = approx_ + n_groups, *g_start = approx_ + 2 * n_groups;
    uint8_t* g_kind = reinterpret_cast<uint8_t*>(lo_);
    const int g_grid = (int)std::min<uint64_t>((n_groups * 32 + kBlock - 1) / kBlock, (uint64_t)eng.numSMs() * 8);
    group_summary_kernel<<<g_grid, kBlock, 0, stream>>>(delta_, base_, flag_, m_, n_groups, g_total, g_bb, g_kind);
    CUDA_CHECK_LAST_ERROR();
    group_stitch_kernel<<<1, 32, 0, stream>>>(state, mask_bit, chunk_, m_, n_groups, delta_, base_, flag_, g_total, g_bb,
                                             g_kind, g_start, start_, slow_, c_init);
    CUDA_CHECK_LAST_ERROR();
    group_write_kernel<<<g_grid, kBlock, 0, stream>>>(delta_, base_, flag_, m_, n_groups, g_kind, g_start, start_);
    CUDA_CHECK_LAST_ERROR();
    launches_ = 6;
}

double SequentialCdf::total() const {
    double t = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&t, start_ + m_, sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    return t;
}

uint64_t SequentialCdf::slowChunks() const {
    unsigned long long s = 0;
    CUDA_CHECK(cudaMemcpyAsync(&s, slow_, sizeof(s), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    return s;
}

void SequentialCdf::sample(const double* uniforms_host, int64_t n_shots, int64_t* out_host) {
    if (n_shots <= 0) return;
    unsigned char* arena = static_cast<unsigned char*>(eng_.scratch(1, (size_t)n_shots * 16));
    double* d_u = reinterpret_cast<double*>(arena);
    int64_t* d_out = reinterpret_cast<int64_t*>(arena + (size_t)n_shots * 8);
    CUDA_CHECK(cudaMemcpyAsync(d_u, uniforms_host, (size_t)n_shots * sizeof(double), cudaMemcpyHostToDevice, stream_));
    int64_t blocks = (n_shots * 32 + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)eng_.numSMs() * 8;
    if (blocks > cap) blocks = cap;
    sample_kernel<<<(unsigned)blocks, kBlock, 0, stream_>>>(state_, mask_bit_, chunk_, m_, start_, d_u, n_shots, d_out);
    CUDA_CHECK_LAST_ERROR();
    ++launches_;
    CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, (size_t)n_shots * sizeof(int64_t), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}

}  // namespace b200
}  // namespace qsim
