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

#include <cstdlib>
#include <mutex>

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "qsim/constants.hpp"

namespace qsim {
namespace b200 {

namespace {

constexpr int kBlock = 256;

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
        if (bit >= 0 && (int)((i >> bit) & 1) != outcome) a = make_cuDoubleComplex(0.0, 0.0);   // bit < 0: scale everything
        else a = make_cuDoubleComplex(__dmul_rn(a.x, scale), __dmul_rn(a.y, scale));
        state[i] = a;
    }
}

// ---- exact sequential CDF ------------------------------------------------------------------------

// K1: approximate (tree-order) sum of each chunk and, in the same sweep, the chunk's exact sequential-order increment
// under each of kCand CANDIDATE binades [2^-j, 2^(1-j)), j = 0..kCand-1, of the running sum (see K2b for why that is a
// sum of independently rounded terms): almost every chunk of a normalised state starts in one of them, so the second
// sweep over the state (K3) only revisits the rest.
constexpr int kCand = 4;

__global__ void chunk_approx_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk,
                                    double* __restrict__ approx, double* __restrict__ cand_delta,
                                    uint8_t* __restrict__ cand_tie, uint64_t m) {
    __shared__ int tie_bits;
    const uint64_t base = (uint64_t)blockIdx.x * chunk;
    const bool cands = (chunk == 4096);
    if (threadIdx.x == 0) tie_bits = 0;
    __syncthreads();
    double acc = 0.0;
    double s[kCand];
    int tie = 0;
#pragma unroll
    for (int c = 0; c < kCand; ++c) s[c] = ldexp(1.0, -c);
    for (int j = threadIdx.x; j < chunk; j += blockDim.x) {
        const double x = masked_prob(state, base + j, mask_bit);
        acc = __dadd_rn(acc, x);
        if (cands && x != 0.0) {   // (adding zero changes nothing and cannot tie)
#pragma unroll
            for (int c = 0; c < kCand; ++c) {
                const double t = __dadd_rn(s[c], x);
                const double err = __dsub_rn(x, __dsub_rn(t, s[c]));   // exact while s >= x (Fast2Sum); else unused
                if (fabs(err) == ldexp(1.0, -c - 53)) tie |= 1 << c;
                s[c] = t;
            }
        }
    }
    // one reduction for the approximate sum and the candidate increments together (order inside a warp, then over
    // the warps: any order is fine for the approximate sum, and the increments add exactly)
    double v[kCand + 1];
    v[0] = acc;
#pragma unroll
    for (int c = 0; c < kCand; ++c) v[c + 1] = cands ? __dsub_rn(s[c], ldexp(1.0, -c)) : 0.0;
#pragma unroll
    for (int q = 0; q <= kCand; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] = __dadd_rn(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
    }
    __shared__ double part[kBlock / 32][kCand + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (tie) atomicOr(&tie_bits, tie);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q <= kCand; ++q) part[warp][q] = v[q];
    }
    __syncthreads();
    if (threadIdx.x <= kCand) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) t = __dadd_rn(t, part[w][threadIdx.x]);
        if (threadIdx.x == 0) approx[blockIdx.x] = t;
        else if (cands) cand_delta[(uint64_t)(threadIdx.x - 1) * m + blockIdx.x] = t;
    }
    if (cands && threadIdx.x == 0) cand_tie[blockIdx.x] = (uint8_t)tie_bits;
}

// K2: exclusive scan of the approximate chunk sums, lo[k] = approximate start of chunk k (any summation order will do:
// it only picks the tentative binade).  Three small launches: per-block totals, scan of the block totals, per-block scan.
constexpr int kScanBlock = 1024;

__global__ void scan_block_totals_kernel(const double* __restrict__ approx, uint64_t m, double* __restrict__ block_tot) {
    __shared__ double red[32];
    const uint64_t i = (uint64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const double t = block_sum(i < m ? approx[i] : 0.0, red);
    if (threadIdx.x == 0) block_tot[blockIdx.x] = t;
}

__global__ void scan_totals_kernel(double* __restrict__ block_tot, int n_blocks, double c_init) {
    // n_blocks <= 2^21 / 1024 = 2048 for the largest state: one thread walks them
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double run = c_init;
        for (int b = 0; b < n_blocks; ++b) { const double v = block_tot[b]; block_tot[b] = run; run += v; }
    }
}

__global__ void scan_blocks_kernel(const double* __restrict__ approx, uint64_t m, const double* __restrict__ block_tot,
                                   double* __restrict__ lo) {
    __shared__ double wsum[32];
    const uint64_t i = (uint64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double v = i < m ? approx[i] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = wsum[lane];
        double wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += up;
        }
        wsum[lane] = wi - w;   // exclusive prefix of the warp totals
    }
    __syncthreads();
    if (i < m) lo[i] = block_tot[blockIdx.x] + wsum[warp] + (incl - v);
}

enum : uint8_t { CH_ZERO = 0, CH_FAST = 1, CH_SLOW = 2, CH_PENDING = 3 };

// K2b: classify every chunk from the approximate running sum.  ZERO: all terms are zero.  Otherwise TENTATIVELY FAST
// in the binade [2^e, 2^(e+1)) of the approximate start: while the running sum c stays inside one binade,
// fl(c + p) = c + rn_u(p) with u = 2^(e-52) whatever c is, so the chunk's increment is a sum of independently rounded
// terms — taken from K1's candidates when 2^e is one of them, else left PENDING for K3; a term that rounds on an exact
// tie depends on c's parity and makes the chunk SLOW.  Whether c really stays inside the binade is decided later, by
// the stitch, on the EXACT running sum (c >= 2^e and c + increment < 2^(e+1)); chunks that fail are replayed.
__global__ void chunk_classify_kernel(int chunk, uint64_t m, const double* __restrict__ approx,
                                      const double* __restrict__ lo, const double* __restrict__ cand_delta,
                                      const uint8_t* __restrict__ cand_tie, double* __restrict__ delta,
                                      double* __restrict__ bin_base, uint8_t* __restrict__ flag,
                                      unsigned int* __restrict__ n_pending, unsigned int* __restrict__ pending) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
        const double a_lo = lo[k], a_sum = approx[k];
        uint8_t f = CH_SLOW;
        double base = 0.0, d = 0.0;
        if (a_sum == 0.0) f = CH_ZERO;  // a tree sum of non-negative terms is 0 only if every term is 0: c unchanged
        else if (chunk == 4096 && a_lo > 1e-290 && a_sum < a_lo) {
            int ex;
            frexp(a_lo, &ex);           // a_lo = m * 2^ex, m in [0.5, 1)
            base = ldexp(1.0, ex - 1);  // 2^e with 2^e <= a_lo < 2^(e+1)
            const int c = 1 - ex;       // candidate index of the binade starting at 2^-c
            if (c >= 0 && c < kCand) {
                d = cand_delta[(uint64_t)c * m + k];
                f = ((cand_tie[k] >> c) & 1) ? CH_SLOW : CH_FAST;
            } else {
                f = CH_PENDING;
                pending[atomicAdd(n_pending, 1u)] = (unsigned int)k;
            }
        }
        flag[k] = f;
        delta[k] = d;
        bin_base[k] = (f == CH_SLOW) ? 0.0 : base;
    }
}

// K3: the pending chunks' sequential-order increments, computed from the surrogate start 2^e (a second read of those
// chunks only).  Each thread adds 16 probabilities to 2^e; a term that rounds on an exact tie makes the chunk SLOW.
__global__ void chunk_surrogate_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk,
                                       const unsigned int* __restrict__ n_pending, const unsigned int* __restrict__ pending,
                                       double* __restrict__ delta, const double* __restrict__ bin_base,
                                       uint8_t* __restrict__ flag) {
    __shared__ double red[32];
    __shared__ int tie_any;
    const unsigned int n = *n_pending;
    for (unsigned int w = blockIdx.x; w < n; w += gridDim.x) {
        const uint64_t k = pending[w];
        const double base = bin_base[k];
        if (threadIdx.x == 0) tie_any = 0;
        __syncthreads();
        const uint64_t g0 = k * (uint64_t)chunk;
        const double half_u = ldexp(base, -53);   // u/2 with u = 2^(e-52)
        double s = base;
        bool tie = false;
        for (int j = threadIdx.x; j < chunk; j += blockDim.x) {
            const double x = masked_prob(state, g0 + j, mask_bit);
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
        }
        __syncthreads();
    }
}

constexpr bool kSequentialReplayOnly = false;   // (true: every replay walks the whole chunk; for A/B measurements)

// Sequential replay of one chunk by a full warp (all lanes carry the same running sum).  Only the additions form a
// dependent chain: the next 32 probabilities are loaded and the 32 broadcasts issued ahead of it.
__device__ __forceinline__ double replay_chunk(const cuDoubleComplex* __restrict__ state, int mask_bit, uint64_t g0,
                                               int chunk, double c) {
    const int lane = threadIdx.x & 31;
    double p_next = (lane < chunk) ? masked_prob(state, g0 + lane, mask_bit) : 0.0;
    for (int g = 0; g < chunk; g += 32) {
        const double p = p_next;
        const int gn = g + 32 + lane;
        p_next = (gn < chunk) ? masked_prob(state, g0 + gn, mask_bit) : 0.0;
        if (chunk - g >= 32) {
            double v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
            for (int j = 0; j < 32; ++j) c = __dadd_rn(c, v[j]);
        } else {
            for (int j = 0; j < chunk - g; ++j) c = __dadd_rn(c, __shfl_sync(0xffffffffu, p, j));
        }
    }
    return c;
}

// The same result much faster when the chunk is the usual 4096 amplitudes: while the running sum stays inside one binade
// [2^e, 2^(e+1)), fl(c + p) = c + rn_u(p) with u = 2^(e-52), whatever c is - so the increments of whole BLOCKS of the chunk
// (32 blocks of 128 elements, one per lane) add exactly and in parallel, a warp scan finds the first block in which the sum
// would reach 2^(e+1), the sum jumps to that block's start, ONLY that block is walked sequentially (it handles the crossing,
// and any tie inside it, exactly), and the walk continues behind it in the new binade.  A tie in a block that would have been
// skipped (a term exactly half an ulp: its rounding depends on the parity of the running sum) sends everything from that
// point on down the sequential path.  Typical replay (one binade crossing): two block sweeps and 128 dependent additions
// instead of 4096.
__device__ __noinline__ double replay_chunk_blocks(const cuDoubleComplex* __restrict__ state, int mask_bit, uint64_t g0, double c) {
    constexpr int kPer = 128;                      // elements per block, 32 blocks = 4096
    const int lane = threadIdx.x & 31;
    int b = 0;                                     // first block not yet accounted for
    while (b < 32) {
        const int E = (__double2hiint(c) >> 20) & 0x7ff;
        if (!(c > 0.0) || E == 0 || E >= 0x7fe) {  // nothing summed yet (or out of the normal range): walk this block
            c = replay_chunk(state, mask_bit, g0 + (uint64_t)b * kPer, kPer, c);
            ++b;
            continue;
        }
        const double B = __hiloint2double(E << 20, 0);        // 2^e <= c < 2^(e+1)
        const double lim = __hiloint2double((E + 1) << 20, 0);
        const double half_u = __hiloint2double((E - 53) > 0 ? (E - 53) << 20 : 0, 0);   // 2^(e-53)
        double L = 0.0;
        bool tie = (E - 53) <= 0;                  // (ulp not representable as a normal number: do not trust the model)
        if (lane >= b) {
            const uint64_t i0 = g0 + (uint64_t)lane * kPer;
            constexpr int kBatch = 16;               // loads in flight per lane (the block is read from HBM, not from a cache)
            for (int j0 = 0; j0 < kPer; j0 += kBatch) {
                double xs[kBatch];
#pragma unroll
                for (int j = 0; j < kBatch; ++j) xs[j] = masked_prob(state, i0 + j0 + j, mask_bit);
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const double x = xs[j];
                    const double t = __dadd_rn(B, x);
                    const double r = __dsub_rn(t, B);                 // rn_u(x): exact (0 for x = 0)
                    const double err = __dsub_rn(x, r);               // exact while B >= x; larger x cross anyway
                    if (fabs(err) == half_u || x >= B) tie = true;
                    L = __dadd_rn(L, r);                              // exact while the block stays below 2^(e+1)
                }
            }
        }
        double incl = L;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = __dadd_rn(incl, up);
        }
        const double excl = __dsub_rn(incl, L);
        const unsigned crossing = __ballot_sync(0xffffffffu, lane >= b && !(__dadd_rn(__dadd_rn(c, excl), L) < lim));
        const int x_blk = crossing ? (__ffs(crossing) - 1) : 32;                      // first block that reaches 2^(e+1)
        const unsigned ties = __ballot_sync(0xffffffffu, tie && lane >= b && lane < x_blk);
        if (ties) {                                // a skipped block holds a tie: the rest goes the sequential way
            const int first = __ffs(ties) - 1;
            c = __dadd_rn(c, __shfl_sync(0xffffffffu, excl, first));                  // the blocks before it are exact
            return replay_chunk(state, mask_bit, g0 + (uint64_t)first * kPer, (32 - first) * kPer, c);
        }
        if (x_blk == 32) return __dadd_rn(c, __shfl_sync(0xffffffffu, incl, 31));     // never left the binade: exact
        c = __dadd_rn(c, __shfl_sync(0xffffffffu, excl, x_blk));                      // exact: still below 2^(e+1)
        c = replay_chunk(state, mask_bit, g0 + (uint64_t)x_blk * kPer, kPer, c);     // the block with the crossing
        b = x_blk + 1;
    }
    return c;
}

// The shot lookup inside a chunk with the same trick: index (0..4095) of the first element at which the sequential running
// sum, started at c, reaches r; 4096 if it never does.  Blocks that neither reach r nor leave the binade are skipped with
// their exact increments; only the block that does one or the other is walked.
__device__ __noinline__ int locate_in_chunk_blocks(const cuDoubleComplex* __restrict__ state, int mask_bit, uint64_t g0, double c,
                                                   double r) {
    constexpr int kPer = 128;
    const int lane = threadIdx.x & 31;
    // sequential walk over blocks [blk, blk + nblk): first index whose running sum is >= r, else -1 (cc carries on)
    auto walk = [&](int blk, int nblk, double& cc) -> int {
        const uint64_t e0 = (uint64_t)blk * kPer;
        const int cnt = nblk * kPer;
        double p_next = masked_prob(state, g0 + e0 + lane, mask_bit);
        for (int g = 0; g < cnt; g += 32) {
            const double p = p_next;
            if (g + 32 < cnt) p_next = masked_prob(state, g0 + e0 + g + 32 + lane, mask_bit);
            for (int j = 0; j < 32; ++j) {
                cc = __dadd_rn(cc, __shfl_sync(0xffffffffu, p, j));
                if (cc >= r) return (int)(e0 + g + j);
            }
        }
        return -1;
    };
    int b = 0;
    while (b < 32) {
        const int E = (__double2hiint(c) >> 20) & 0x7ff;
        if (!(c > 0.0) || E == 0 || E >= 0x7fe) {
            const int idx = walk(b, 1, c);
            if (idx >= 0) return idx;
            ++b;
            continue;
        }
        const double B = __hiloint2double(E << 20, 0);
        const double lim = __hiloint2double((E + 1) << 20, 0);
        const double half_u = __hiloint2double((E - 53) > 0 ? (E - 53) << 20 : 0, 0);
        double L = 0.0;
        bool tie = (E - 53) <= 0;
        if (lane >= b) {
            const uint64_t i0 = g0 + (uint64_t)lane * kPer;
            constexpr int kBatch = 16;
            for (int j0 = 0; j0 < kPer; j0 += kBatch) {
                double xs[kBatch];
#pragma unroll
                for (int j = 0; j < kBatch; ++j) xs[j] = masked_prob(state, i0 + j0 + j, mask_bit);
#pragma unroll
                for (int j = 0; j < kBatch; ++j) {
                    const double x = xs[j];
                    const double t = __dadd_rn(B, x);
                    const double rr = __dsub_rn(t, B);
                    const double err = __dsub_rn(x, rr);
                    if (fabs(err) == half_u || x >= B) tie = true;
                    L = __dadd_rn(L, rr);
                }
            }
        }
        double incl = L;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = __dadd_rn(incl, up);
        }
        const double excl = __dsub_rn(incl, L);
        const double end = __dadd_rn(__dadd_rn(c, excl), L);          // exact for every block before the first crossing
        const unsigned crossing = __ballot_sync(0xffffffffu, lane >= b && !(end < lim));
        const unsigned reaching = __ballot_sync(0xffffffffu, lane >= b && end >= r);
        const int x_blk = crossing ? (__ffs(crossing) - 1) : 32;
        const int y_blk = reaching ? (__ffs(reaching) - 1) : 32;
        const int stop = x_blk < y_blk ? x_blk : y_blk;                // first block that reaches r or leaves the binade
        const unsigned ties = __ballot_sync(0xffffffffu, tie && lane >= b && lane < stop);
        if (ties) {
            const int first = __ffs(ties) - 1;
            c = __dadd_rn(c, __shfl_sync(0xffffffffu, excl, first));
            const int idx = walk(first, 32 - first, c);
            return idx >= 0 ? idx : 4096;
        }
        if (stop == 32) return 4096;
        c = __dadd_rn(c, __shfl_sync(0xffffffffu, excl, stop));
        const int idx = walk(stop, 1, c);
        if (idx >= 0) return idx;
        b = stop + 1;
    }
    return 4096;
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

// inclusive warp scan of exact addends (multiples of one ulp, or garbage that the caller's range check rejects)
__device__ __forceinline__ double warp_incl_scan(double v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = __dadd_rn(v, up);
    }
    return v;
}

__global__ void group_summary_kernel(const double* __restrict__ delta, const double* __restrict__ bin_base,
                                     const uint8_t* __restrict__ flag, int chunk, uint64_t m, uint64_t n_groups,
                                     const double* __restrict__ cand_delta, const uint8_t* __restrict__ cand_tie,
                                     double* __restrict__ g_total, double* __restrict__ g_bb, uint8_t* __restrict__ g_kind,
                                     double* __restrict__ g_cand, uint8_t* __restrict__ g_tie) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t g = warp; g < n_groups; g += n_warps) {
        double total, bb, excl;
        const uint8_t kind = group_scan(delta, bin_base, flag, m, g, total, bb, excl);
        if (lane == 0) { g_kind[g] = kind; g_total[g] = total; g_bb[g] = bb; }
        // the group's total and tie flags under each candidate binade of K1 (the stitch picks by the exact sum)
        const uint64_t kk = g * 32 + lane;
        unsigned tie = (chunk == 4096 && kk < m) ? cand_tie[kk] : 0xffu;
        if (kk >= m) tie = 0;
        if (chunk != 4096) tie = 0xffu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tie |= __shfl_xor_sync(0xffffffffu, tie, o);
#pragma unroll
        for (int c = 0; c < kCand; ++c) {
            const double d = (chunk == 4096 && kk < m) ? cand_delta[(uint64_t)c * m + kk] : 0.0;
            const double t = __shfl_sync(0xffffffffu, warp_incl_scan(d), 31);
            if (lane == 0) g_cand[(uint64_t)c * n_groups + g] = t;
        }
        if (lane == 0) g_tie[g] = (uint8_t)tie;
    }
}

__global__ void group_stitch_kernel(const cuDoubleComplex* __restrict__ state, int mask_bit, int chunk, uint64_t m,
                                    uint64_t n_groups, const double* __restrict__ delta,
                                    const double* __restrict__ bin_base, const uint8_t* __restrict__ flag,
                                    const double* __restrict__ g_total, const double* __restrict__ g_bb,
                                    uint8_t* __restrict__ g_kind, double* __restrict__ g_start,
                                    const double* __restrict__ cand_delta, const uint8_t* __restrict__ cand_tie,
                                    const double* __restrict__ g_cand, const uint8_t* __restrict__ g_tie,
                                    uint8_t* __restrict__ g_choice,
                                    double* __restrict__ start, unsigned long long* __restrict__ n_slow, double c_init) {
    const int lane = threadIdx.x & 31;
    double c = c_init;
    unsigned long long slow = 0;
    unsigned int n_cand = 0, n_simple = 0, n_chunkwise = 0;   // groups by the path they took (diagnostics, QSIM_DEBUG_CDF)
    // lane l holds the summary of group g0 + l.  The summaries of the next kAhead batches of 32 groups are in flight
    // while a batch is stitched (a batch of all-zero or candidate-path groups is shorter than one trip to L2, so a single
    // batch of look-ahead left the warp waiting for memory most of the time).
    constexpr int kAhead = 4;
    struct Batch {
        double t, b, ct[kCand];
        int k, tie;
    };
    auto load_batch = [&](uint64_t g_first, Batch& bt) {
        const uint64_t gl = g_first + lane;
        const bool in = gl < n_groups;
        bt.t = in ? g_total[gl] : 0.0;
        bt.b = in ? g_bb[gl] : 0.0;
        bt.k = in ? (int)g_kind[gl] : (int)G_ZERO;
        bt.tie = in ? (int)g_tie[gl] : 0xff;
#pragma unroll
        for (int cd = 0; cd < kCand; ++cd) bt.ct[cd] = in ? g_cand[(uint64_t)cd * n_groups + gl] : 0.0;
    };
    Batch ring[kAhead];
#pragma unroll
    for (int a = 0; a < kAhead; ++a) load_batch((uint64_t)a * 32, ring[a]);
    auto stitch_batch = [&](uint64_t g0, const Batch& cur_batch) {
        const double t_cur = cur_batch.t, b_cur = cur_batch.b;
        const int k_cur = cur_batch.k, tie_cur = cur_batch.tie;
        double ct_cur[kCand];
#pragma unroll
        for (int cd = 0; cd < kCand; ++cd) ct_cur[cd] = cur_batch.ct[cd];
        int my_choice = 0xff;
        // a whole batch of all-zero groups (long stretches of a sparse state): the running sum does not move
        if (__all_sync(0xffffffffu, k_cur == (int)G_ZERO)) {
            if (g0 + lane < n_groups) { g_start[g0 + lane] = c; g_choice[g0 + lane] = 0xff; }
            return;
        }
        const int cnt = (n_groups - g0) < 32 ? (int)(n_groups - g0) : 32;
        double my_start = 0.0;
        bool my_done = false;
        // The loop below is ONE warp walking the groups in order: its speed is the latency of the chain through c.  Group
        // i + 1's summary is broadcast (shuffles, independent of c) while group i's addition is in flight, and the binade of
        // c is read from its exponent bits instead of frexp / ldexp (3.3 ms -> see DESIGN 5.2 for a dense 30-qubit state).
        struct Summary { int kind, tie; double t, bb, ct[kCand]; };
        auto fetch = [&](int i) {
            Summary sm;
            sm.kind = __shfl_sync(0xffffffffu, k_cur, i & 31);
            sm.tie = __shfl_sync(0xffffffffu, tie_cur, i & 31);
            sm.t = __shfl_sync(0xffffffffu, t_cur, i & 31);
            sm.bb = __shfl_sync(0xffffffffu, b_cur, i & 31);
#pragma unroll
            for (int q = 0; q < kCand; ++q) sm.ct[q] = __shfl_sync(0xffffffffu, ct_cur[q], i & 31);
            return sm;
        };
        Summary nxt = fetch(0);
        for (int i = 0; i < cnt; ++i) {
            const Summary cur = nxt;
            nxt = fetch(i + 1);   // (i + 1 == 32 wraps to lane 0: fetched, never used)
            const int kind = cur.kind;
            if (kind == G_ZERO) { if (lane == i) my_start = c; continue; }
            // the whole group under the candidate binade the EXACT running sum is in
            if (c > 0.0) {
                const int E = (__double2hiint(c) >> 20) & 0x7ff;   // c = m * 2^(E - 1022), m in [0.5, 1)
                const int cd = 1023 - E;                             // c in [2^-cd, 2^(1-cd))
                if (cd >= 0 && cd < kCand && !((cur.tie >> cd) & 1)) {
                    double tot = cur.ct[0];
#pragma unroll
                    for (int q = 1; q < kCand; ++q) tot = (q == cd) ? cur.ct[q] : tot;
                    const double c_end = __dadd_rn(c, tot);
                    if (c_end < __hiloint2double((E + 1) << 20, 0)) {   // still below 2^(1-cd)
                        if (lane == i) { my_start = c; my_choice = cd; }
                        c = c_end;
                        ++n_cand;
                        continue;
                    }
                }
            }
            if (kind == G_SIMPLE) {
                const double bb = cur.bb;
                const double c_end = __dadd_rn(c, cur.t);
                if (c >= bb && c_end < 2.0 * bb) { if (lane == i) my_start = c; c = c_end; ++n_simple; continue; }   // exact binade check
            }
            ++n_chunkwise;
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
                if (fj == CH_FAST && c >= bj && cc < 2.0 * bj) { c = cc; continue; }   // exact binade check on the true values
                // the approximate start picked the wrong binade (c sits next to a power of two): try K1's candidate
                // for the binade the exact c is in
                if (chunk == 4096 && c > 0.0) {
                    int ex;
                    frexp(c, &ex);
                    const int cd = 1 - ex;
                    if (cd >= 0 && cd < kCand && !((cand_tie[k0 + j] >> cd) & 1)) {
                        const double c2 = __dadd_rn(c, cand_delta[(uint64_t)cd * m + k0 + j]);
                        if (c2 < ldexp(1.0, ex)) { c = c2; continue; }
                    }
                }
                c = (chunk == 4096 && !kSequentialReplayOnly) ? replay_chunk_blocks(state, mask_bit, (k0 + j) * (uint64_t)chunk, c)
                                                              : replay_chunk(state, mask_bit, (k0 + j) * (uint64_t)chunk, chunk, c);
                ++slow;
            }
            if (lane == i) my_done = true;
        }
        if (lane < cnt) {
            g_start[g0 + lane] = my_start;
            g_choice[g0 + lane] = (uint8_t)my_choice;
            if (my_done) g_kind[g0 + lane] = G_DONE;
        }
    };
    for (uint64_t g0 = 0; g0 < n_groups; g0 += 32 * kAhead) {
#pragma unroll
        for (int a = 0; a < kAhead; ++a) {   // (static ring slots: the look-ahead lives in registers)
            const uint64_t ga = g0 + (uint64_t)a * 32;
            if (ga < n_groups) {
                stitch_batch(ga, ring[a]);
                load_batch(ga + 32 * kAhead, ring[a]);
            }
        }
    }
    if (lane == 0) {
        start[m] = c;
        // chunks replayed sequentially in the low 24 bits; above them the number of groups stitched chunk by chunk
        // (20 bits) and through a candidate binade (20 bits)
        *n_slow = (slow & 0xffffffULL) | ((unsigned long long)(n_chunkwise & 0xfffffu) << 24) | ((unsigned long long)(n_cand & 0xfffffu) << 44);
        (void)n_simple;
    }
}

__global__ void group_write_kernel(const double* __restrict__ delta, const double* __restrict__ bin_base,
                                   const uint8_t* __restrict__ flag, uint64_t m, uint64_t n_groups,
                                   const uint8_t* __restrict__ g_kind, const double* __restrict__ g_start,
                                   const double* __restrict__ cand_delta, const uint8_t* __restrict__ g_choice,
                                   double* __restrict__ start) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g < n_groups; g += n_warps) {
        if (g_kind[g] == G_DONE) continue;
        const uint64_t kk = g * 32 + (threadIdx.x & 31);
        const int choice = g_choice[g];
        double excl;
        if (choice < kCand) {   // stitched under a candidate binade
            const double d = kk < m ? cand_delta[(uint64_t)choice * m + kk] : 0.0;
            excl = __dsub_rn(warp_incl_scan(d), d);
        } else {
            double total, bb;
            group_scan(delta, bin_base, flag, m, g, total, bb, excl);
        }
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
            if (chunk == 4096 && !kSequentialReplayOnly) {   // block-wise: two parallel sweeps of the chunk + one 128-element walk
                const int idx = locate_in_chunk_blocks(state, mask_bit, g0, c, r);
                if (idx < 4096) result = (int64_t)(g0 + (uint64_t)idx);
                found = true;
            }
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

namespace {

// Marginals.  Lanes always enumerate the five lowest index bits (coalesced 512-byte rows whatever the chosen qubits
// are); the chosen bits among them ("low bin bits") split a warp's lanes into sub-bins, reduced with butterflies over
// the other lane bits.  A block handles one value of the remaining chosen bits and one segment of the other indices;
// sums are formed in a fixed order throughout.
__global__ void marginal_partial_kernel(const cuDoubleComplex* __restrict__ state, uint64_t hole_mask, uint64_t keep,
                                        unsigned low_bin_mask, const uint64_t* __restrict__ high_bits_of,
                                        uint64_t others_per_seg, int n_seg, uint64_t step_dep,
                                        const uint64_t* __restrict__ start_dep, int n_sub, double* __restrict__ partial) {
    __shared__ double part[kBlock / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int mh = blockIdx.x / n_seg, seg = blockIdx.x % n_seg;
    // x runs over: (fixed high bin bits) | (other-index o deposited into the non-bin, non-lane bits) | lane
    uint64_t x = start_dep[(size_t)seg * n_warps + warp];
    const uint64_t fixed = high_bits_of[mh] | (uint64_t)lane;
    double acc = 0.0;
    for (uint64_t j = warp; j < others_per_seg; j += n_warps) {
        acc = __dadd_rn(acc, prob_of(state[x | fixed]));
        x = ((x | hole_mask) + step_dep) & keep;
    }
    // butterflies over the lane bits that are not bin bits: afterwards every lane holds its sub-bin's warp total
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const double other = __shfl_xor_sync(0xffffffffu, acc, 1 << b);
        if (!((low_bin_mask >> b) & 1u)) acc = __dadd_rn(acc, other);
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        // lane l with no non-bin lane bits set represents sub-bin pattern (l & low_bin_mask); sum the warps in order
        if ((lane & ~low_bin_mask & 31u) == 0) {
            double t = 0.0;
            for (int w = 0; w < n_warps; ++w) t = __dadd_rn(t, part[w][lane]);
            // compact the pattern's bits into the sub-bin number
            int sub = 0, jb = 0;
            for (int b = 0; b < 5; ++b)
                if ((low_bin_mask >> b) & 1u) { sub |= ((lane >> b) & 1) << jb; ++jb; }
            partial[((size_t)mh * n_sub + sub) * n_seg + seg] = t;
        }
    }
}

// out[m] = sum over segments, m assembled from (high part, sub-bin) through the bin table
__global__ void marginal_final_kernel(const double* __restrict__ partial, int n_seg, int n_rows, const int* __restrict__ row_to_bin,
                                      double* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double t = 0.0;
    for (int s2 = 0; s2 < n_seg; ++s2) t = __dadd_rn(t, partial[(size_t)r * n_seg + s2]);
    out[row_to_bin[r]] = t;
}

uint64_t deposit_bits(uint64_t v, uint64_t mask) {   // software pdep (host)
    uint64_t r = 0;
    int j = 0;
    for (int b = 0; b < 64; ++b)
        if ((mask >> b) & 1ULL) { r |= ((v >> j) & 1ULL) << b; ++j; }
    return r;
}

}  // namespace

void marginal_probabilities(const cuDoubleComplex* state, int n_bits, const int* bits, int k, double* host_out, Engine& eng) {
    if (k < 0 || k > 12 || k > n_bits) throw std::invalid_argument("marginal over 0..12 qubits");
    uint64_t bin_mask = 0;
    for (int i = 0; i < k; ++i) {
        if (bits[i] < 0 || bits[i] >= n_bits || ((bin_mask >> bits[i]) & 1ULL)) throw std::invalid_argument("Duplicate or invalid qubit in marginal");
        bin_mask |= 1ULL << bits[i];
    }
    const int n_bins = 1 << k;
    if (n_bits < 5) {   // tiny states: on the host from the probabilities
        std::vector<double> p((size_t)1 << n_bits);
        double* d = static_cast<double*>(eng.scratch(0, p.size() * sizeof(double)));
        launch_probabilities(state, d, 0, p.size(), eng.numSMs(), eng.stream());
        CUDA_CHECK(cudaMemcpyAsync(p.data(), d, p.size() * sizeof(double), cudaMemcpyDeviceToHost, eng.stream()));
        CUDA_CHECK(cudaStreamSynchronize(eng.stream()));
        for (int m = 0; m < n_bins; ++m) host_out[m] = 0.0;
        for (size_t x = 0; x < p.size(); ++x) {
            int m = 0;
            for (int i = 0; i < k; ++i) m |= (int)((x >> bits[i]) & 1) << i;
            host_out[m] += p[x];
        }
        eng.countLaunch(1);
        return;
    }
    const uint64_t full_mask = (1ULL << n_bits) - 1ULL;
    const unsigned low_bin_mask = (unsigned)(bin_mask & 31ULL);
    const uint64_t high_bin_mask = bin_mask & ~31ULL;
    const uint64_t hole_mask = high_bin_mask | 31ULL;             // bits the running index does not enumerate
    const uint64_t keep = full_mask & ~hole_mask;
    const int k_low = __builtin_popcount(low_bin_mask), k_high = k - k_low;
    const int n_sub = 1 << k_low, n_high = 1 << k_high;
    const uint64_t n_others = 1ULL << (n_bits - 5 - k_high);      // other-indices per high-bin value (lanes excluded)
    const int n_warps = kBlock / 32;
    int n_seg = 1;
    while ((uint64_t)n_high * n_seg < (uint64_t)eng.numSMs() * 8 && (uint64_t)n_seg * 2 * n_warps <= n_others) n_seg *= 2;
    const uint64_t others_per_seg = n_others / n_seg;
    // host tables: index bits of every high-bin value; deposited start of every (segment, warp); row -> outcome
    std::vector<uint64_t> h((size_t)n_high + (size_t)n_seg * n_warps);
    std::vector<int> hb, lb;   // positions (within `bits`) of the high / low chosen bits, ascending by index bit
    for (int b = 0; b < n_bits; ++b)
        for (int i = 0; i < k; ++i)
            if (bits[i] == b) (b < 5 ? lb : hb).push_back(i);
    for (int mh = 0; mh < n_high; ++mh) {
        uint64_t v = 0;
        for (int j = 0; j < k_high; ++j) v |= (uint64_t)((mh >> j) & 1) << bits[hb[j]];
        h[mh] = v;
    }
    for (int sg = 0; sg < n_seg; ++sg)
        for (int w = 0; w < n_warps; ++w) h[(size_t)n_high + (size_t)sg * n_warps + w] = deposit_bits((uint64_t)sg * others_per_seg + w, keep);
    std::vector<int> row_to_bin((size_t)n_high * n_sub);
    for (int mh = 0; mh < n_high; ++mh)
        for (int sub = 0; sub < n_sub; ++sub) {
            int m = 0;
            for (int j = 0; j < k_high; ++j) m |= ((mh >> j) & 1) << hb[j];
            for (int j = 0; j < k_low; ++j) m |= ((sub >> j) & 1) << lb[j];
            row_to_bin[(size_t)mh * n_sub + sub] = m;
        }
    const size_t tab_bytes = (h.size() * sizeof(uint64_t) + 63) & ~(size_t)63;
    const size_t map_bytes = (row_to_bin.size() * sizeof(int) + 63) & ~(size_t)63;
    const size_t part_bytes = ((size_t)n_bins * n_seg + n_bins) * sizeof(double);
    unsigned char* arena = static_cast<unsigned char*>(eng.scratch(0, tab_bytes + map_bytes + part_bytes + 64));
    uint64_t* d_tab = reinterpret_cast<uint64_t*>(arena);
    int* d_map = reinterpret_cast<int*>(arena + tab_bytes);
    double* d_part = reinterpret_cast<double*>(arena + tab_bytes + map_bytes);
    double* d_out = d_part + (size_t)n_bins * n_seg;
    cudaStream_t stream = eng.stream();
    CUDA_CHECK(cudaMemcpyAsync(d_tab, h.data(), h.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(d_map, row_to_bin.data(), row_to_bin.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));   // the host tables are pageable and go out of scope
    marginal_partial_kernel<<<n_high * n_seg, kBlock, 0, stream>>>(state, hole_mask, keep, low_bin_mask, d_tab, others_per_seg, n_seg,
                                                                  deposit_bits((uint64_t)n_warps, keep), d_tab + n_high, n_sub, d_part);
    CUDA_CHECK_LAST_ERROR();
    marginal_final_kernel<<<(n_bins + kBlock - 1) / kBlock, kBlock, 0, stream>>>(d_part, n_seg, n_bins, d_map, d_out);
    CUDA_CHECK_LAST_ERROR();
    CUDA_CHECK(cudaMemcpyAsync(host_out, d_out, (size_t)n_bins * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    eng.countLaunch(2);
}

SequentialCdf::SequentialCdf(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng, double c_init)
    : state_(state), n_(n), mask_bit_(mask_bit), stream_(eng.stream()), eng_(eng) {
    setup();
    prepare();
    classify(c_init);
    stitch(c_init);
}

SequentialCdf::SequentialCdf(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng, Deferred)
    : state_(state), n_(n), mask_bit_(mask_bit), stream_(eng.stream()), eng_(eng) {
    setup();
    prepare();
}

void SequentialCdf::setup() {
    chunk_ = n_ >= 4096 ? 4096 : (int)n_;
    m_ = n_ / (uint64_t)chunk_;
    n_groups_ = (m_ + 31) / 32;
    const size_t m8 = (m_ + 2) * sizeof(double);
    const size_t m4 = ((m_ + 2) * sizeof(unsigned int) + 7) & ~(size_t)7;
    const size_t m1 = (m_ + 8) & ~(size_t)7;
    const size_t g8 = (n_groups_ + 1) * sizeof(double), g1 = (n_groups_ + 8) & ~(size_t)7;
    unsigned char* arena = static_cast<unsigned char*>(
        eng_.scratch(0, (5 + kCand) * m8 + 16 + 2 * m1 + m4 + (3 + kCand) * g8 + 3 * g1 + 64));
    approx_ = reinterpret_cast<double*>(arena);
    lo_ = reinterpret_cast<double*>(arena + m8);
    delta_ = reinterpret_cast<double*>(arena + 2 * m8);
    base_ = reinterpret_cast<double*>(arena + 3 * m8);
    start_ = reinterpret_cast<double*>(arena + 4 * m8);
    cand_delta_ = reinterpret_cast<double*>(arena + 5 * m8);
    slow_ = reinterpret_cast<unsigned long long*>(arena + (5 + kCand) * m8);
    n_pending_ = reinterpret_cast<unsigned int*>(arena + (5 + kCand) * m8 + 8);
    flag_ = arena + (5 + kCand) * m8 + 16;
    cand_tie_ = flag_ + m1;
    pending_ = reinterpret_cast<unsigned int*>(cand_tie_ + m1);
    unsigned char* garena = reinterpret_cast<unsigned char*>(pending_) + m4;
    g_total_ = reinterpret_cast<double*>(garena);
    g_bb_ = g_total_ + (n_groups_ + 1);
    g_start_ = g_bb_ + (n_groups_ + 1);
    g_cand_ = g_start_ + (n_groups_ + 1);
    g_kind_ = garena + (3 + kCand) * g8;
    g_tie_ = g_kind_ + g1;
    g_choice_ = g_tie_ + g1;
}

void SequentialCdf::prepare() {
    CUDA_CHECK(cudaMemsetAsync(n_pending_, 0, sizeof(unsigned int), stream_));
    chunk_approx_kernel<<<(unsigned)m_, kBlock, 0, stream_>>>(state_, mask_bit_, chunk_, approx_, cand_delta_, cand_tie_, m_);
    CUDA_CHECK_LAST_ERROR();
    launches_ = 1;
}

double SequentialCdf::approxTotal() {
    // (g_start_ is not in use yet: one double of it receives the sum)
    final_sum_kernel<<<1, kBlock, 0, stream_>>>(approx_, (int)m_, g_start_);
    CUDA_CHECK_LAST_ERROR();
    ++launches_;
    double t = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&t, g_start_, sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    return t;
}

void SequentialCdf::classify(double approx_c_init) {
    cudaStream_t stream = stream_;
    {   // (start_ is only written by the stitch: until then its head holds the per-block totals)
        const int n_blocks = (int)((m_ + kScanBlock - 1) / kScanBlock);
        scan_block_totals_kernel<<<n_blocks, kScanBlock, 0, stream>>>(approx_, m_, start_);
        CUDA_CHECK_LAST_ERROR();
        scan_totals_kernel<<<1, 32, 0, stream>>>(start_, n_blocks, approx_c_init);
        CUDA_CHECK_LAST_ERROR();
        scan_blocks_kernel<<<n_blocks, kScanBlock, 0, stream>>>(approx_, m_, start_, lo_);
        CUDA_CHECK_LAST_ERROR();
    }
    const int c_grid = (int)std::min<uint64_t>((m_ + kBlock - 1) / kBlock, (uint64_t)eng_.numSMs() * 8);
    chunk_classify_kernel<<<c_grid, kBlock, 0, stream>>>(chunk_, m_, approx_, lo_, cand_delta_, cand_tie_, delta_, base_, flag_,
                                                        n_pending_, pending_);
    CUDA_CHECK_LAST_ERROR();
    chunk_surrogate_kernel<<<eng_.numSMs() * 8, kBlock, 0, stream>>>(state_, mask_bit_, chunk_, n_pending_, pending_, delta_,
                                                                    base_, flag_);
    CUDA_CHECK_LAST_ERROR();
    if (std::getenv("QSIM_DEBUG_CDF")) {
        std::vector<uint8_t> hf(m_);
        CUDA_CHECK(cudaStreamSynchronize(stream));
        cudaMemcpy(hf.data(), flag_, m_, cudaMemcpyDeviceToHost);
        uint64_t cnt[4] = {0, 0, 0, 0};
        for (uint64_t k = 0; k < m_; ++k) cnt[hf[k] & 3]++;
        fprintf(stderr, "[cdf] zero %llu fast %llu slow %llu pending %llu\n", (unsigned long long)cnt[0], (unsigned long long)cnt[1], (unsigned long long)cnt[2], (unsigned long long)cnt[3]);
    }
    const int g_grid = (int)std::min<uint64_t>((n_groups_ * 32 + kBlock - 1) / kBlock, (uint64_t)eng_.numSMs() * 8);
    group_summary_kernel<<<g_grid, kBlock, 0, stream>>>(delta_, base_, flag_, chunk_, m_, n_groups_, cand_delta_, cand_tie_, g_total_,
                                                       g_bb_, g_kind_, g_cand_, g_tie_);
    CUDA_CHECK_LAST_ERROR();
    launches_ += 6;
}

void SequentialCdf::stitch(double c_init) {
    cudaStream_t stream = stream_;
    const int g_grid = (int)std::min<uint64_t>((n_groups_ * 32 + kBlock - 1) / kBlock, (uint64_t)eng_.numSMs() * 8);
    group_stitch_kernel<<<1, 32, 0, stream>>>(state_, mask_bit_, chunk_, m_, n_groups_, delta_, base_, flag_, g_total_, g_bb_,
                                             g_kind_, g_start_, cand_delta_, cand_tie_, g_cand_, g_tie_, g_choice_, start_, slow_,
                                             c_init);
    CUDA_CHECK_LAST_ERROR();
    group_write_kernel<<<g_grid, kBlock, 0, stream>>>(delta_, base_, flag_, m_, n_groups_, g_kind_, g_start_, cand_delta_, g_choice_,
                                                     start_);
    CUDA_CHECK_LAST_ERROR();
    launches_ += 2;
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
    return s & 0xffffffULL;
}

void SequentialCdf::sample(const double* uniforms_host, int64_t n_shots, int64_t* out_host) {
    if (n_shots <= 0) return;
    if (std::getenv("QSIM_DEBUG_CDF")) {
        unsigned long long raw = 0;
        CUDA_CHECK(cudaMemcpyAsync(&raw, slow_, sizeof(raw), cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        fprintf(stderr, "[cdf] chunks %llu (groups %llu): replayed sequentially %llu; groups stitched chunk by chunk %llu, through a candidate binade %llu\n",
                (unsigned long long)m_, (unsigned long long)n_groups_, raw & 0xffffffULL, (raw >> 24) & 0xfffffULL, (raw >> 44) & 0xfffffULL);
    }
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


// ---- small states (n <= 14): the whole CDF in one CTA's shared memory -------------------------------------------------
// Below ~2^15 amplitudes the chunked machinery above is launch- and latency-bound (five kernels, two of them one warp
// wide).  Here one CTA computes a TREE-order inclusive scan of the probabilities into shared memory and answers every shot
// by binary search.  A sum of m non-negative terms differs from the exact sum by at most m * 2^-53 * total in ANY order, so
// wherever the shot is farther than tau = 2^-37 * total (> 2 * 2^14 * 2^-53) from the neighbouring scan values the
// reference's sequential sums C satisfy C[k-1] < u <= C[k] too; the (rare) shots inside that margin replay the sequential
// sum from the amplitudes.  The result is the reference's index in every case (src/Simulator.cu:164-185).
namespace {
constexpr int kSmallThreads = 1024;

__global__ void __launch_bounds__(kSmallThreads) small_cdf_sample_kernel(const cuDoubleComplex* __restrict__ state, uint32_t size,
                                                                         const double* __restrict__ uniforms, int64_t n_shots,
                                                                         int64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    double* c = reinterpret_cast<double*>(small_smem);   // indexed through pad(): the threads' segments start in different banks
    auto pad = [](uint32_t i) { return i + (i >> 4); };
    __shared__ double warp_tot[kSmallThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = size >= (uint32_t)kSmallThreads ? size / kSmallThreads : 1u;
    const uint32_t i0 = tid * per;
    double run = 0.0;
    if (i0 < size)
        for (uint32_t j = 0; j < per; ++j) { run = __dadd_rn(run, prob_of(state[i0 + j])); c[pad(i0 + j)] = run; }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    double offset = incl - run, total = 0.0;
    for (uint32_t w = 0; w < kSmallThreads / 32; ++w) {
        if (w < warp) offset += warp_tot[w];
        total += warp_tot[w];
    }
    if (i0 < size && offset != 0.0)
        for (uint32_t j = 0; j < per; ++j) c[pad(i0 + j)] += offset;
    __syncthreads();
    const double tau = total * 0x1.0p-37;
    for (int64_t shot = tid; shot < n_shots; shot += kSmallThreads) {
        const double r = uniforms[shot];
        uint32_t lo = 0, hi = size;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (c[pad(mid)] >= r) hi = mid; else lo = mid + 1; }
        uint32_t k = lo;
        bool sure;
        if (k == 0) sure = (c[0] >= r);                               // C[0] = p[0] exactly
        else if (k == size) sure = (r - c[pad(size - 1)] > tau);
        else sure = (c[pad(k)] - r > tau) && (r - c[pad(k - 1)] > tau);
        if (!sure) {                                                  // the reference's own loop
            double C = 0.0;
            k = size;
            for (uint32_t i = 0; i < size; ++i) {
                C = __dadd_rn(C, prob_of(state[i]));
                if (C >= r) { k = i; break; }
            }
        }
        out[shot] = (int64_t)k;
    }
}
}  // namespace

bool sample_small_state(const cuDoubleComplex* state, int n_qubits, const double* uniforms_host, int64_t n_shots,
                        int64_t* out_host, Engine& eng) {
    if (n_qubits > kSmallCdfMaxQubits || std::getenv("QSIM_NO_SMALL_CDF")) return false;
    const uint32_t size = 1u << n_qubits;
    const size_t smem = ((size_t)size + size / 16 + 2) * sizeof(double);
    static int configured_dev = -1;   // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    if (dev != configured_dev) {
        CUDA_CHECK(cudaFuncSetAttribute(small_cdf_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * ((1u << kSmallCdfMaxQubits) + (1u << kSmallCdfMaxQubits) / 16 + 2))));
        configured_dev = dev;
    }
    cudaStream_t stream = eng.stream();
    unsigned char* scratch = static_cast<unsigned char*>(eng.scratch(1, (size_t)n_shots * 16));
    double* d_u = reinterpret_cast<double*>(scratch);
    int64_t* d_out = reinterpret_cast<int64_t*>(scratch + (size_t)n_shots * 8);
    CUDA_CHECK(cudaMemcpyAsync(d_u, uniforms_host, (size_t)n_shots * 8, cudaMemcpyHostToDevice, stream));
    small_cdf_sample_kernel<<<1, kSmallThreads, smem, stream>>>(state, size, d_u, n_shots, d_out);
    CUDA_CHECK_LAST_ERROR();
    CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, (size_t)n_shots * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    eng.countLaunch();
    return true;
}

}  // namespace b200
}  // namespace qsim
