// Fused-pass kernel for sm_100a: ONE read and ONE write of the state vector per pass, however
// many gates the pass carries (the reference sweeps HBM once per gate, src/Simulator.cu:48-154 and
// src/Gates.cu:31-410 of the reference).
//
// Structure (persistent, one CTA per SM, kComputeWarps warps, every warp computes):
//   A tile is 2^t amplitudes (t <= 12: 64 KiB): up to five runs of contiguous index bits, moved by ONE
//   tensor-map TMA instruction per direction (cp.async.bulk.tensor.5d, SASS UTMALDG / UTMASTG) into a
//   ring of shared-memory stages, completion through mbarrier complete_tx.  Per sweep every thread pulls
//   kSlots amplitudes from shared memory into registers with 128-bit conflict-free loads, runs the
//   sweep's op list — register-resident targets in-thread, lane-resident targets with __shfl_xor
//   butterflies, diagonal gates as a single complex multiply wherever their qubits live — and writes back.
//   TMA issue is folded into warp 0 (lane q issues box q): after tile i has been computed it issues the
//   store of tile i, makes sure the store of tile i-1 has drained its stage and refills that stage with tile
//   i-1+S, so one or two tile loads are always in flight while the warps compute.  (A dedicated TMA
//   warp would be the seventeenth warp and, with the 4-warp register allocation granularity, cost
//   a quarter of the register file.)
//   Variants of the same kernel: basis-state input (PassParams::init_basis: no loads, zero fill + one generated
//   tile), out-of-place store for a fused qubit exchange between GPUs (PassParams::redirect), store-side index
//   maps (deferred X gates, folded trailing CNOTs).
//
// Algorithmic traffic: 2 * 16 * 2^n bytes per pass.
#include "kernels.cuh"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "jit.hpp"
#include "pass_device.cuh"

namespace qsim {
namespace b200 {

namespace {

// ---- op application on the thread's register file -------------------------------------------------
// Everything below is straight-line code over the kSlots register slots: no per-slot branches, so the
// independent dependency chains interleave (controls become selects, and ops without controls —
// the common case — carry no predicate at all).

// Every op is written OUT OF PLACE: it reads the register file x and writes all of y.  The interpreter loop
// alternates the two files (a -> b, b -> a), so no value has to survive in "its" register across the loop
// header — an in-place formulation makes the compiler copy the whole file once per op.

template <int J, int KIND, bool CTRL>
__device__ __forceinline__ void reg_pairs(const DevOp& op, uint32_t sm, const double (&xr_)[kSlots], const double (&xi_)[kSlots],
                                          double (&yr_)[kSlots], double (&yi_)[kSlots]) {
    const double m00r = op.m[0], m00i = op.m[1], m01r = op.m[2], m01i = op.m[3];
    const double m10r = op.m[4], m10i = op.m[5], m11r = op.m[6], m11i = op.m[7];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        if (k & (1 << J)) continue;   // compile-time: k enumerates the slots whose target bit is 0
        const int k1 = k | (1 << J);
        const double xr = xr_[k], xi = xi_[k], yr = xr_[k1], yi = xi_[k1];
        double n0r, n0i, n1r, n1i;
        if (KIND == OP_FLIP) {
            n0r = yr; n0i = yi; n1r = xr; n1i = xi;
        } else if (KIND == OP_ADIAG) {
            n0r = m01r * yr - m01i * yi; n0i = m01r * yi + m01i * yr;
            n1r = m10r * xr - m10i * xi; n1i = m10r * xi + m10i * xr;
        } else if (KIND == OP_MATREAL) {
            n0r = m00r * xr + m01r * yr; n0i = m00r * xi + m01r * yi;
            n1r = m10r * xr + m11r * yr; n1i = m10r * xi + m11r * yi;
        } else {
            n0r = m00r * xr - m00i * xi + m01r * yr - m01i * yi;
            n0i = m00r * xi + m00i * xr + m01r * yi + m01i * yr;
            n1r = m10r * xr - m10i * xi + m11r * yr - m11i * yi;
            n1i = m10r * xi + m10i * xr + m11r * yi + m11i * yr;
        }
        if (CTRL) {
            const bool on = (sm >> k) & 1;
            yr_[k] = on ? n0r : xr; yi_[k] = on ? n0i : xi; yr_[k1] = on ? n1r : yr; yi_[k1] = on ? n1i : yi;
        } else {
            yr_[k] = n0r; yi_[k] = n0i; yr_[k1] = n1r; yi_[k1] = n1i;
        }
    }
}

template <int KIND, bool CTRL>
__device__ __forceinline__ void lane_target(const DevOp& op, uint32_t sm, uint32_t tid, const double (&xr_)[kSlots],
                                            const double (&xi_)[kSlots], double (&yr_)[kSlots], double (&yi_)[kSlots]) {
    const int lm = 1 << op.tbit;
    if (KIND == OP_FLIP) {
        // pure data movement: every slot is fetched from the partner lane, or (controls not satisfied) from the
        // thread's own lane — the select is on the source lane, not on the data
        const int lane = (int)(tid & 31u);
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int src = (!CTRL || ((sm >> k) & 1)) ? (lane ^ lm) : lane;
            yr_[k] = __shfl_sync(0xffffffffu, xr_[k], src);
            yi_[k] = __shfl_sync(0xffffffffu, xi_[k], src);
        }
        return;
    }
    const bool b = (tid >> op.tbit) & 1;
    // coefficient of my own amplitude and of my partner's
    const double cor = b ? op.m[6] : op.m[0], coi = b ? op.m[7] : op.m[1];
    const double cpr = b ? op.m[4] : op.m[2], cpi = b ? op.m[5] : op.m[3];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        const double xr = xr_[k], xi = xi_[k];
        const double pr = shfl_xor_f64(xr, lm), pi = shfl_xor_f64(xi, lm);
        double nr, ni;
        if (KIND == OP_ADIAG) { nr = cpr * pr - cpi * pi; ni = cpr * pi + cpi * pr; }
        else if (KIND == OP_MATREAL) { nr = cor * xr + cpr * pr; ni = cor * xi + cpr * pi; }
        else {
            nr = cor * xr - coi * xi + cpr * pr - cpi * pi;
            ni = cor * xi + coi * xr + cpr * pi + cpi * pr;
        }
        if (CTRL) {
            const bool on = (sm >> k) & 1;
            yr_[k] = on ? nr : xr; yi_[k] = on ? ni : xi;
        } else {
            yr_[k] = nr; yi_[k] = ni;
        }
    }
}

template <bool CTRL>
__device__ __forceinline__ void diagonal(const DevOp& op, uint32_t sm, uint32_t tid, uint64_t gbase, const double (&xr_)[kSlots],
                                         const double (&xi_)[kSlots], double (&yr_)[kSlots], double (&yi_)[kSlots]) {
    const double d0r = op.m[0], d0i = op.m[1], d1r = op.m[6], d1i = op.m[7];
    // which diagonal entry a slot takes: by the slot (register-resident target) or the same for the whole thread
    // (target on a tid bit or outside the tile) — folded into ONE mask so there is a single multiply loop
    const bool bt = (op.thome == T_THREAD) ? ((tid & op.tmask_thr) != 0) : ((gbase & op.tmask_out) != 0);
    const uint32_t tsl = (op.thome == T_REG) ? (uint32_t)op.tslots : (bt ? 0xffffu : 0u);
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        const bool b = (tsl >> k) & 1;
        const double pr = b ? d1r : d0r, pi = b ? d1i : d0i;
        const double xr = xr_[k], xi = xi_[k];
        const double nr = xr * pr - xi * pi, ni = xr * pi + xi * pr;
        if (CTRL) { const bool on = (sm >> k) & 1; yr_[k] = on ? nr : xr; yi_[k] = on ? ni : xi; }
        else { yr_[k] = nr; yi_[k] = ni; }
    }
}

// A fused run of diagonal gates: amplitude(l) *= TABLE[l] * U * prod_{tile bits j of l} E_j.  op.cmask_thr says which of
// those factors exist at all (uniform), so a run whose qubits all lie outside the tile costs one complex multiply per
// amplitude and a run without outside partners only the table look-up.
__device__ __forceinline__ void phase_op(const DevOp& op, const double2* __restrict__ tables, const double2* eu,
                                         const SweepDesc& sd, uint32_t tid, uint32_t base_local, const double (&xr_)[kSlots],
                                         const double (&xi_)[kSlots], double (&yr_)[kSlots], double (&yi_)[kSlots]) {
    const uint32_t present = op.cmask_thr;
    const double2* e = eu + (size_t)op.tmask_out * 13;
    // thread-wide factor: U and the E_j of the tile bits held by tid bits
    double pr = 1.0, pi = 0.0;
    if (present & (1u << 12)) { pr = e[12].x; pi = e[12].y; }
    uint32_t reg_present = 0;
#pragma unroll
    for (int j = 0; j < kMaxRegBits; ++j)
        if (j < sd.r && ((present >> sd.reg_pos[j]) & 1u)) reg_present |= 1u << j;
    if (present & 0xfffu) {
#pragma unroll
        for (int b = 0; b < kMaxTileBits - kMaxRegBits; ++b) {
            if (b < sd.nthr && ((present >> sd.thr_pos[b]) & 1u) && ((tid >> b) & 1)) {   // (first two tests are uniform)
                const double2 v = e[sd.thr_pos[b]];
                const double r = pr * v.x - pi * v.y;
                pi = pr * v.y + pi * v.x;
                pr = r;
            }
        }
    }
    double2 f[kSlots];
    if (present & (1u << 13)) {
        const double2* tb = tables + op.cmask_out;
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const double2 t = __ldg(tb + (base_local ^ (uint32_t)sd.slot_off[k]));
            f[k].x = t.x * pr - t.y * pi;
            f[k].y = t.x * pi + t.y * pr;
        }
    } else {
        if (present & (1u << 14)) {   // constant table: fold it into the thread-wide factor
            const double2 t = __ldg(tables + op.cmask_out);
            const double r = t.x * pr - t.y * pi;
            pi = t.x * pi + t.y * pr;
            pr = r;
        }
#pragma unroll
        for (int k = 0; k < kSlots; ++k) { f[k].x = pr; f[k].y = pi; }
    }
    if (reg_present) {   // E_j of register-resident tile bits
#pragma unroll
        for (int j = 0; j < kMaxRegBits; ++j) {
            if (!((reg_present >> j) & 1u)) continue;
            const double2 v = e[sd.reg_pos[j]];
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                if (!((k >> j) & 1)) continue;
                const double r = f[k].x * v.x - f[k].y * v.y;
                f[k].y = f[k].x * v.y + f[k].y * v.x;
                f[k].x = r;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        const double xr = xr_[k], xi = xi_[k];
        yr_[k] = xr * f[k].x - xi * f[k].y;
        yi_[k] = xr * f[k].y + xi * f[k].x;
    }
}

#if QSIM_REG_BITS > 3
#define QSIM_PAIR_CASES_BIT3(KIND)                                                                          \
    case (KIND) * kOpcodesPerKind + 8: reg_pairs<3, KIND, false>(op, 0xffffu, xr, xi, yr, yi); break;      \
    case (KIND) * kOpcodesPerKind + 9: reg_pairs<3, KIND, true>(op, sm, xr, xi, yr, yi); break;
#else
#define QSIM_PAIR_CASES_BIT3(KIND)
#endif
#if QSIM_REG_BITS > 2
#define QSIM_PAIR_CASES_BIT2(KIND)                                                                          \
    case (KIND) * kOpcodesPerKind + 6: reg_pairs<2, KIND, false>(op, 0xffffu, xr, xi, yr, yi); break;      \
    case (KIND) * kOpcodesPerKind + 7: reg_pairs<2, KIND, true>(op, sm, xr, xi, yr, yi); break;
#else
#define QSIM_PAIR_CASES_BIT2(KIND)
#endif
#define QSIM_PAIR_CASES(KIND)                                                                                \
    case (KIND) * kOpcodesPerKind + 0: lane_target<KIND, false>(op, 0xffffu, tid, xr, xi, yr, yi); break;   \
    case (KIND) * kOpcodesPerKind + 1: lane_target<KIND, true>(op, sm, tid, xr, xi, yr, yi); break;         \
    case (KIND) * kOpcodesPerKind + 2: reg_pairs<0, KIND, false>(op, 0xffffu, xr, xi, yr, yi); break;       \
    case (KIND) * kOpcodesPerKind + 3: reg_pairs<0, KIND, true>(op, sm, xr, xi, yr, yi); break;             \
    case (KIND) * kOpcodesPerKind + 4: reg_pairs<1, KIND, false>(op, 0xffffu, xr, xi, yr, yi); break;       \
    case (KIND) * kOpcodesPerKind + 5: reg_pairs<1, KIND, true>(op, sm, xr, xi, yr, yi); break;             \
    QSIM_PAIR_CASES_BIT2(KIND)                                                                               \
    QSIM_PAIR_CASES_BIT3(KIND)

// one op, register file x -> register file y
__device__ __forceinline__ void apply_op(const DevOp& op, uint32_t opcode, uint32_t sm, uint32_t tid, uint64_t gbase,
                                         const double2* __restrict__ tables, const double2* eu, const SweepDesc& sd,
                                         uint32_t base_local, const double (&xr)[kSlots], const double (&xi)[kSlots],
                                         double (&yr)[kSlots], double (&yi)[kSlots]) {
    switch (opcode) {
        QSIM_PAIR_CASES(OP_MAT)
        QSIM_PAIR_CASES(OP_MATREAL)
        QSIM_PAIR_CASES(OP_ADIAG)
        QSIM_PAIR_CASES(OP_FLIP)
        case kOpcodeDiag + 0: case kOpcodeDiag + 2: diagonal<false>(op, 0xffffu, tid, gbase, xr, xi, yr, yi); break;
        case kOpcodeDiag + 1: case kOpcodeDiag + 3: diagonal<true>(op, sm, tid, gbase, xr, xi, yr, yi); break;
        case kOpcodePhase: phase_op(op, tables, eu, sd, tid, base_local, xr, xi, yr, yi); break;
        case kOpcodeCopy:
#pragma unroll
            for (int k = 0; k < kSlots; ++k) { yr[k] = xr[k]; yi[k] = xi[k]; }
            break;
        default: __builtin_unreachable();
    }
}
// The interpreter: applies the pass's sweeps to one tile in shared memory (QSIM_COMPUTE_TILE of pass_kernel_body.inc).
__device__ __forceinline__ void interp_compute_tile(const PassParams& P, unsigned char* tile, uint64_t gbase, uint32_t tid,
                                                    const DevOp* sops, const double2* eu, const uint16_t* base_tab) {
    const PassDesc& pd = P.pd;
    const uint32_t warp = tid >> 5;
    double ar[kSlots], ai[kSlots], br[kSlots], bi[kSlots];
#pragma unroll 1
    for (int sw = 0; sw < pd.n_sweeps; ++sw) {
        const SweepDesc& sd = pd.sweep[sw];
        if (sw > 0) __syncthreads();
        const uint32_t n_active = 1u << sd.nthr;
        const bool warp_active = (warp << 5) < n_active;
        const bool last_sweep = (sw + 1 == pd.n_sweeps);
        const uint32_t xl = last_sweep ? pd.xor_local : 0u;                 // deferred X gates, see the store
        const int n_tail = last_sweep ? pd.n_tail : 0;                      // trailing bit flips, see the store
        const bool mapped_load = (sw == 0 && pd.n_head > 0) || sd.n_head > 0;   // folded leading flips, see the load
        // a load or a store that reaches into other threads' slots: everybody must have loaded before anybody stores
        const bool permuted_store = (xl != 0u) || (n_tail > 0) || mapped_load;
        if (!warp_active) {
            if (permuted_store) __syncthreads();   // keep the barrier count equal across warps
            continue;
        }
        const bool active = tid < n_active;
        const int slots = 1 << sd.r;
        const uint32_t base_local = base_tab[sw * kComputeThreads + (int)tid];
        const uint32_t tile_u32 = smem_u32(tile);
        const bool full_sweep = (sd.r == kMaxRegBits) && (sd.nthr == kMaxTileBits - kMaxRegBits);
        if (mapped_load) {
            // folded leading flips: every slot is read from the pre-image of its index under the flips
            uint32_t lb;
            const uint16_t* loff;
            if (sw == 0 && pd.n_head > 0) {
                lb = base_tab[(pd.n_sweeps + 1) * kComputeThreads + (int)tid];
                for (int f = 0; f < pd.n_head_dyn; ++f)
                    if ((gbase & pd.head_dyn[f].cmask_out) == pd.head_dyn[f].cval_out) lb ^= pd.head_dyn[f].w;
                loff = pd.load_slot_off;
            } else {   // a later sweep (no table row: mapped on the fly)
                lb = sd.head_const;
                for (int j = 0; j < pd.t; ++j)
                    if ((base_local >> j) & 1u) lb ^= sd.head_lin[j];
                loff = sd.load_slot_off;
            }
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                ar[k] = 0.0;
                ai[k] = 0.0;
                if (active && k < slots) lds128(tile_u32 + (lb ^ (uint32_t)loff[k]) * 16u, ar[k], ai[k]);
            }
        } else {
            const uint32_t my_addr = tile_u32 + base_local * 16u;
            if (full_sweep) {
#pragma unroll
                for (int k = 0; k < kSlots; ++k) lds128(my_addr + (uint32_t)sd.slot_off[k] * 16u, ar[k], ai[k]);
            } else {
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    ar[k] = 0.0;
                    ai[k] = 0.0;
                    if (active && k < slots) lds128(my_addr + (uint32_t)sd.slot_off[k] * 16u, ar[k], ai[k]);
                }
            }
        }
        const uint32_t ops_u32 = smem_u32(sops);
        // the interpreter: fetch() finds the next op that applies to this tile (ops whose controls outside the
        // tile fail are skipped), apply_op() alternates between the two register files
        int o = sd.op_begin;
        const int o_end = sd.op_end;
        uint32_t opcode = 0, sm = 0;
        auto fetch = [&]() -> bool {
            if (o >= o_end) return false;
            // first 16 bytes of the record: kind|thome|tbit|opcode, slotmask|tslots, cmask_thr, cval_thr
            const uint4 hdr = lds_u4(ops_u32 + (uint32_t)o * (uint32_t)sizeof(DevOp));
            opcode = hdr.x >> 24;
            // an op whose controls outside the tile fail becomes a register-file copy (no extra control-flow
            // edge around apply_op: that edge is what made the compiler copy the file at every loop header)
            const bool skip = (opcode & 0x80u) && (gbase & sops[o].cmask_out) != sops[o].cval_out;
            opcode = skip ? kOpcodeCopy : (opcode & 0x7fu);
            sm = ((tid & hdr.z) == hdr.w) ? (hdr.y & 0xffffu) : 0u;
            return true;
        };
        bool in_b = false;
#pragma unroll 1
        for (;;) {
            if (!fetch()) break;
            apply_op(sops[o], opcode, sm, tid, gbase, reinterpret_cast<const double2*>(P.phase_tables), eu, sd, base_local, ar, ai, br, bi);
            ++o;
            if (!fetch()) { in_b = true; break; }
            apply_op(sops[o], opcode, sm, tid, gbase, reinterpret_cast<const double2*>(P.phase_tables), eu, sd, base_local, br, bi, ar, ai);
            ++o;
        }
        // The pass's index permutations ride on the last sweep's store: the folded trailing bit flips (an affine
        // map of the tile-local index, see TailDyn), then the deferred X gates (l ^= xor_local).  The targets are
        // other threads' slots — and with folded LEADING flips this sweep's load read other threads' slots — hence
        // the barrier: everybody has finished loading before anybody stores.
        if (permuted_store) __syncthreads();
        {
            uint32_t l[kSlots];
            if (last_sweep) {
                uint32_t sb = base_tab[pd.n_sweeps * kComputeThreads + (int)tid] ^ xl;
                for (int f = 0; f < pd.n_dyn; ++f)   // flips controlled from outside the tile: the same for the whole tile
                    if ((gbase & pd.dyn[f].cmask_out) == pd.dyn[f].cval_out) sb ^= pd.dyn[f].w;
#pragma unroll
                for (int k = 0; k < kSlots; ++k) l[k] = tile_u32 + (sb ^ (uint32_t)pd.store_slot_off[k]) * 16u;
            } else {
#pragma unroll
                for (int k = 0; k < kSlots; ++k) l[k] = tile_u32 + (base_local ^ (uint32_t)sd.slot_off[k]) * 16u;
            }
            // (the result sits in whichever register file the last op wrote)
            auto store = [&](const double (&sr)[kSlots], const double (&si)[kSlots]) {
                if (full_sweep) {
#pragma unroll
                    for (int k = 0; k < kSlots; ++k) sts128(l[k], sr[k], si[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < kSlots; ++k)
                        if (active && k < slots) sts128(l[k], sr[k], si[k]);
                }
            };
            if (in_b) store(br, bi); else store(ar, ai);
        }
    }
}

}  // namespace

#define QSIM_PASS_KERNEL fused_pass_kernel
#define QSIM_COMPUTE_TILE interp_compute_tile
#define QSIM_KERNEL_LINKAGE
#include "pass_kernel_body.inc"
#undef QSIM_PASS_KERNEL
#undef QSIM_COMPUTE_TILE
#undef QSIM_KERNEL_LINKAGE


// worst case with three full stages must fit (pick_stages never has to go below the three the partner-tile store needs)
static_assert(3 * (16 << kMaxTileBits) + (kMaxOpsPerPass + 1) * sizeof(DevOp) + 2 * kMaxPhaseOps * 13 * sizeof(double2) +
                      2 * 3 * sizeof(uint64_t) + (kMaxSweeps + 2) * kComputeThreads * sizeof(uint16_t) <=
                  (size_t)kMaxDynamicSmem,
              "shared-memory budget of a pass");

size_t pass_smem_bytes(const PassDesc& pd, int stages) {
    return (size_t)stages * ((size_t)16 << pd.t) + ((size_t)pd.n_ops + 1) * sizeof(DevOp) + 2 * (size_t)pd.n_phase * 13 * sizeof(double2) +
           2 * (size_t)stages * sizeof(uint64_t) + ((size_t)pd.n_sweeps + 2) * kComputeThreads * sizeof(uint16_t);
}

// Deepest ring that fits the 227 KiB of shared memory (at most kMaxStages).
int pick_stages(const PassDesc& pd, int wanted) {
    int st = wanted > 0 ? wanted : kMaxStages;
    if (st > kMaxStages) st = kMaxStages;
    if (st < 3) st = 3;   // the partner-tile store waits for the load of the next item: needs three stages
    while (st > 1 && pass_smem_bytes(pd, st) > (size_t)kMaxDynamicSmem) --st;
    return st;
}

namespace {

using encode_fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn get_encode() {
    static encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<encode_fn>(p);
    }();
    return fn;
}

// The state viewed as a 5-D tensor of doubles whose dimensions are the pass's runs of tile bits.
bool encode_tensor_map(const PassParams& prm, void* base, CUtensorMap* out) {
    encode_fn enc = get_encode();
    if (!enc) return false;
    const PassDesc& pd = prm.pd;
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estride[5] = {1, 1, 1, 1, 1};
    for (int d = 0; d < 5; ++d) {
        const TmaDim& td = pd.tma_dim[d];
        gdim[d] = (cuuint64_t)1 << td.range_bits;
        box[d] = 1u << td.box_bits;
        if (d == 0) { gdim[d] *= 2; box[d] *= 2; }
        else gstride[d - 1] = (cuuint64_t)16 << td.start_bit;
    }
    static const CUtensorMapL2promotion promo = [] {
        const char* e = std::getenv("QSIM_TMA_L2PROMO");   // (development aid) 0 / 64 / 128 / 256
        const int v = e ? std::atoi(e) : 256;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                      : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    }();
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The dynamic shared-memory opt-in is a per-device (per-context) function attribute: set it once on every device this
// process launches on.
cudaError_t ensure_smem_optin() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fused_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynamicSmem);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

}  // namespace

DualTune::~DualTune() {
    for (auto& pair : ev)
        for (cudaEvent_t e : pair)
            if (e) cudaEventDestroy(e);
}

cudaError_t launch_pass(const PassParams& params_in, int num_sms, cudaStream_t stream, const DevOp* host_ops, const JitSlots& jit) {
    PassParams params = params_in;
    alignas(64) CUtensorMap tmap, tmap_keep, tmap_send;
    std::memset(&tmap, 0, sizeof(tmap));
    if (params.use_tensor_map && !encode_tensor_map(params, params.state, &tmap)) return cudaErrorInvalidValue;
    tmap_keep = tmap;
    tmap_send = tmap;
    if (params.redirect) {
        if (!params.dst_keep || !params.dst_send || params.init_basis) return cudaErrorInvalidValue;
        for (int j = 0; j < params.pd.t && params.redirect != 4; ++j)   // (4: the gathering pass has it as its highest tile qubit)
            if (params.pd.tile_bits[j] == params.redirect_bit) return cudaErrorInvalidValue;
        if (params.use_tensor_map && (!encode_tensor_map(params, params.dst_keep, &tmap_keep) ||
                                      !encode_tensor_map(params, params.dst_send, &tmap_send)))
            return cudaErrorInvalidValue;
    }
    if (cudaError_t e = ensure_smem_optin()) return e;
    const size_t smem = pass_smem_bytes(params.pd, params.stages);
    if (params.stages < 1 || smem > (size_t)kMaxDynamicSmem) return cudaErrorInvalidValue;
    uint64_t grid = params.n_tiles < (uint64_t)num_sms ? params.n_tiles : (uint64_t)num_sms;
    params.send_ctas = 0;
    if (params.redirect && grid >= 16) {
        // enough senders that their compute on the leaving half keeps up with NVLink, few enough that each sender's share
        // of the link drains a stage about as fast as it computes one (measured on C2 at 31 q, 2 GPUs: 16 senders 43.8 ms
        // per circuit, 32: 24.9, 48: 18.6, 63: 19.6, 100: 19.2, no split: 22.4)
        int send = (int)(grid / 3);
        if (const char* e = std::getenv("QSIM_SEND_CTAS")) send = std::atoi(e);
        if (send > 0 && send < (int)grid) params.send_ctas = send;
    }
    params.mid_ctas = 0;
    if (params.redirect >= 2) {
        // the in-place exchange counts items per CTA across the two GPUs: it needs the sender / keeper split, the
        // handshake words, and a tile XOR (deferred X gates) that does not cross the exchanged bit
        if (!params.send_ctas || ((params.pd.xdep >> params.redirect_bit) & 1ULL) || !params.hs_local || !params.hs_peer ||
            !params.hs_error || params.dst_keep != params.state)
            return cudaErrorInvalidValue;
    }
    if (params.redirect >= 3) {
        if (!params.use_tensor_map || grid < 32 || params.n_tiles < 64 || params.split_bit < 0 || params.split_bit >= params.pd.n ||
            params.split_bit == params.redirect_bit)
            return cudaErrorInvalidValue;
        const uint64_t vbit = 1ULL << params.redirect_bit, wbit = 1ULL << params.split_bit;
        if (params.redirect == 3) {
            // first half of a split exchange: three CTA classes - the exchanged quarter of the tiles (NVLink-bound), the staying
            // half and the leaving quarter this pass does not move (both HBM-bound, 2 : 1 tiles)
            if ((params.pd.xdep | params.pd.tile_mask) & (vbit | wbit)) return cudaErrorInvalidValue;
            const int rest = (int)grid - params.send_ctas;
            params.mid_ctas = rest * 2 / 3;
            if (params.mid_ctas < 1 || params.send_ctas + params.mid_ctas >= (int)grid) return cudaErrorInvalidValue;
        } else {
            // second half: the exchanged qubit is this pass's highest tile qubit, moved by TMA instructions of its own; the tiles
            // with the split bit set load half of their boxes over NVLink (so twice the CTAs of a full-tile sender class)
            if (params.pd.tma_instr_bits < 1 || params.pd.tile_bits[params.pd.t - 1] != params.redirect_bit ||
                ((params.pd.xdep | params.pd.tile_mask) & wbit))
                return cudaErrorInvalidValue;
            int gath = (int)grid * 3 / 5;
            if (const char* e = std::getenv("QSIM_GATHER_CTAS")) gath = std::atoi(e);
            if (gath < 1 || gath >= (int)grid) return cudaErrorInvalidValue;
            params.send_ctas = gath;
        }
    }
    // A kernel specialised for this pass's structure (large states, pre-compiled circuits; see jit.hpp).  In the default
    // mode the compile runs on a background thread: until it is ready the interpreter kernel below does the pass.
    if (params.use_tensor_map && (host_ops || params.pd.n_ops == 0)) {
        // compute-heavy passes take the two-warp-group build (not for basis-state input or the in-place exchange, which
        // only the one-group skeleton implements); its factor tables of fused diagonal runs exist in four copies
        const size_t smem_dual = smem + 2 * (size_t)params.pd.n_phase * 13 * sizeof(double2);
        const bool dual_ok = !params.init_basis && params.redirect < 2 && params.stages == 3 && smem_dual <= (size_t)kMaxDynamicSmem &&
                             params.n_tiles >= (uint64_t)num_sms && jit_dual_wanted(params.pd, host_ops);
        int dual = dual_ok ? 1 : 0;
        // a pass that carries a fused exchange is bound by NVLink, not by the SM: one group (two tiles in flight), unless this
        // pass has been measured without the exchange and the two-group build won
        if (params.redirect && !(jit.tune && jit.tune->choice == 1)) dual = 0;
        // measured choice: once both builds of this pass are loaded, one ordinary launch of each is bracketed by events; the
        // times are collected at a later launch (cudaEventQuery: never blocks) and the faster build stays
        // (states of >= 26 qubits: a pass takes >= 0.4 ms there and one measurement is meaningful; below, the estimate decides)
        DualTune* tune = (dual_ok && !params.redirect && params.pd.n >= 26 && jit.tune && jit.kernel && jit_dual_autotune()) ? jit.tune : nullptr;
        const uint64_t grid_one = grid;
        int timing = -1;
        if (tune) {
            for (int v = 0; v < 2; ++v)
                if (tune->state[v] == 1 && cudaEventQuery(tune->ev[v][1]) == cudaSuccess) {
                    if (cudaEventElapsedTime(&tune->ms[v], tune->ev[v][0], tune->ev[v][1]) == cudaSuccess) tune->state[v] = 2;
                    else tune->state[v] = 0;
                }
            if (tune->choice < 0 && tune->state[0] == 2 && tune->state[1] == 2) {
                tune->choice = tune->ms[1] < 0.98f * tune->ms[0] ? 1 : 0;
                static const bool verbose = std::getenv("QSIM_DUAL_VERBOSE") != nullptr;
                if (verbose)
                    std::fprintf(stderr, "qsim_b200: pass over %d qubits, %d ops in %d sweeps (~%d FP64 instructions per thread): one warp "
                                 "group %.3f ms, two %.3f ms -> %s\n", params.pd.n, params.pd.n_ops, params.pd.n_sweeps,
                                 jit_fp64_estimate(params.pd, host_ops), tune->ms[0], tune->ms[1], tune->choice ? "two" : "one");
            }
            if (tune->choice >= 0) dual = tune->choice;
            else if (jit.kernel[0] && jit.kernel[1]) {   // both ready: time the one that has not been timed yet
                const int v = tune->state[1] == 0 ? 1 : (tune->state[0] == 0 ? 0 : -1);
                if (v >= 0) { dual = v; timing = v; }
            } else if (jit.kernel[1] && !jit.kernel[0] && !(jit.tried && jit.tried[0])) {
                dual = 0;   // the two-group build is there, the one-group build has not been asked for yet: ask (below)
            }
        }
        (void)cudaGetLastError();   // (cudaEventQuery's cudaErrorNotReady is not an error)
        // both groups of a CTA need a tile: a state of fewer than two tiles per SM runs on half as many CTAs as it has tiles
        if (dual && grid * 2 > params.n_tiles) grid = params.n_tiles / 2;
        std::shared_ptr<JitKernel> local_k;
        std::shared_ptr<JitRequest> local_r;
        std::shared_ptr<JitKernel>* slot = jit.kernel ? jit.kernel + dual : &local_k;
        std::shared_ptr<JitRequest>* req = jit.request ? jit.request + dual : &local_r;
        char* tried = jit.tried ? jit.tried + dual : nullptr;
        if (!*slot && !(tried && *tried)) {
            const JitMode mode = jit_mode();
            bool pending = false;
            if (mode != JitMode::Off && (jit.force || jit_wanted(params.pd))) {
                if (!*req) *req = jit_make_request(params.pd, host_ops, dual != 0);
                const bool async = !jit.force && mode == JitMode::Auto && jit_async_enabled() && jit.kernel != nullptr;
                *slot = jit_lookup(**req, /*needs_device=*/true, async, &pending);
            }
            if (!pending) {
                if (tried) *tried = 1;
                req->reset();
            }
        }
        // the build asked for is still being compiled, the other one is loaded: use that rather than the interpreter
        if (!*slot && jit.kernel && jit.kernel[1 - dual] && (dual == 0 ? dual_ok : true)) {
            dual = 1 - dual;
            slot = jit.kernel + dual;
            timing = -1;
            grid = grid_one;
            if (dual && grid * 2 > params.n_tiles) grid = params.n_tiles / 2;
        }
        if (!*slot) grid = grid_one;   // (the interpreter: one group)
        if (*slot) {
            const bool timed = tune && timing == dual;
            if (timed) {
                for (cudaEvent_t& e : tune->ev[dual])
                    if (!e && cudaEventCreate(&e) != cudaSuccess) return cudaGetLastError();
                if (cudaError_t e = cudaEventRecord(tune->ev[dual][0], stream)) return e;
            }
            const cudaError_t rc = jit_launch(**slot, params, &tmap, &tmap_keep, &tmap_send, (unsigned)grid, dual ? smem_dual : smem, stream);
            if (timed && rc == cudaSuccess) {
                if (cudaError_t e = cudaEventRecord(tune->ev[dual][1], stream)) return e;
                tune->state[dual] = 1;
            }
            return rc;
        }
    }
    fused_pass_kernel<<<(unsigned)grid, kComputeThreads, smem, stream>>>(params, tmap, tmap_keep, tmap_send);
    return cudaGetLastError();
}

}  // namespace b200
}  // namespace qsim
