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

namespace qsim {
namespace b200 {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}


__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

// tensor-map (tiled) variants: one instruction moves a whole 5-D box (SASS UTMALDG / UTMASTG)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, const int (&c)[5], uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const int (&c)[5], const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(map),
                 "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(smem_u32(smem_src))
                 : "memory");
}

// coordinates of the box that starts at global amplitude index g
__device__ __forceinline__ void tma_coords(const PassDesc& pd, uint64_t g, int (&c)[5]) {
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        const uint32_t rb = pd.tma_dim[d].range_bits;
        const uint64_t v = rb ? ((g >> pd.tma_dim[d].start_bit) & ((1ULL << rb) - 1)) : 0ULL;
        c[d] = (int)(d == 0 ? v * 2 : v);   // dimension 0 counts doubles
    }
}

__device__ __forceinline__ uint64_t instr_offset(const PassDesc& pd, uint32_t q) {
    uint64_t off = 0;
    const int box_bits = pd.t - pd.tma_instr_bits;
#pragma unroll 1
    for (int b = 0; b < pd.tma_instr_bits; ++b)
        if ((q >> b) & 1) off |= 1ULL << pd.tile_bits[box_bits + b];
    return off;
}

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void lds128(uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask) {
    return __shfl_xor_sync(0xffffffffu, v, lane_mask);
}

struct TileGeom {
    uint64_t base;  // global amplitude index of the tile's element 0
};

__device__ __forceinline__ uint64_t tile_base(const PassDesc& pd, uint64_t tau) {
    uint64_t b = 0;
#pragma unroll 1
    for (int s = 0; s < pd.n_segments; ++s) b |= ((tau >> pd.seg[s].src_shift) & pd.seg[s].mask) << pd.seg[s].dst_shift;
    return b;
}

__device__ __forceinline__ uint64_t run_offset(const PassDesc& pd, uint32_t run) {
    uint64_t off = 0;
#pragma unroll 1
    for (int b = 0; b < pd.n_high; ++b)
        if ((run >> b) & 1) off |= 1ULL << pd.tile_bits[pd.L + b];
    return off;
}

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

}  // namespace

__global__ void __launch_bounds__(kComputeThreads, 1)
fused_pass_kernel(const __grid_constant__ PassParams P, const __grid_constant__ CUtensorMap tmap,
                  const __grid_constant__ CUtensorMap tmap_keep, const __grid_constant__ CUtensorMap tmap_send) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const PassDesc& pd = P.pd;
    const uint32_t tile_bytes = 16u << pd.t;
    unsigned char* tiles = smem;
    const int n_stages = P.stages;
    DevOp* sops = reinterpret_cast<DevOp*>(smem + (size_t)n_stages * tile_bytes);
    double2* eu = reinterpret_cast<double2*>(sops + pd.n_ops + 1);       // (E_0..E_11, U) of every OP_PHASE, per tile
    uint64_t* full = reinterpret_cast<uint64_t*>(eu + (size_t)pd.n_phase * 13);
    // base_tab[sw][tid]: the tile-local index of the thread's slot 0 in sweep sw (tile-invariant)
    uint16_t* base_tab = reinterpret_cast<uint16_t*>(full + 2 * n_stages);

    const uint32_t tid = threadIdx.x, warp = tid >> 5;

    // stage this pass's op list in shared memory (broadcast reads later)
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.ops);
        uint4* dst = reinterpret_cast<uint4*>(sops);
        const int n16 = pd.n_ops * (int)(sizeof(DevOp) / 16);
        for (int i = tid; i < n16; i += kComputeThreads) dst[i] = src[i];
        if (tid < (int)(sizeof(DevOp) / 16)) dst[n16 + tid] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int sw = 0; sw < pd.n_sweeps; ++sw) {
        const SweepDesc& sd = pd.sweep[sw];
        uint32_t bl = 0;
        for (int b = 0; b < sd.nthr; ++b)
            if ((tid >> b) & 1) bl |= 1u << sd.thr_pos[b];
        base_tab[sw * kComputeThreads + (int)tid] = (uint16_t)bl;
        if (sw + 1 == pd.n_sweeps) {
            // extra row: where the final store puts the thread's slot 0 (the folded flips' affine map, see TailDyn)
            uint32_t sb = pd.tail_const;
            for (int j = 0; j < pd.t; ++j)
                if ((bl >> j) & 1) sb ^= pd.tail_lin[j];
            base_tab[pd.n_sweeps * kComputeThreads + (int)tid] = (uint16_t)sb;
        }
        if (sw == 0) {
            // and where the first load finds it (leading flips, inverse map)
            uint32_t lb = pd.head_const;
            for (int j = 0; j < pd.t; ++j)
                if ((bl >> j) & 1) lb ^= pd.head_lin[j];
            base_tab[(pd.n_sweeps + 1) * kComputeThreads + (int)tid] = (uint16_t)lb;
        }
    }
    if (tid == 0) {
        if (P.use_tensor_map) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t n_tiles = P.n_tiles;
    uint64_t first = blockIdx.x, stride = gridDim.x;
    // Work items of this CTA.  Without a tile XOR item i is tile first + i*stride.  With one, tiles are
    // handled in partner pairs (tau, tau ^ xor_tau): items 2j and 2j+1 are the two members of pair
    // first + j*stride (the pair id is the tile number with the pivot bit of xor_tau squeezed out), each is
    // written to the other's location, and a tile is only overwritten after its own contents have been loaded
    // (see the wait before the store below).
    const uint64_t xor_tau = pd.xor_tau;
    const int pivot = xor_tau ? 63 - __clzll((long long)xor_tau) : 0;
    uint64_t n_units = xor_tau ? n_tiles / 2 : n_tiles;
    // Global base index of a unit = its number deposited into the index bits that are neither tile bits nor the
    // pivot ("holes").  Stepping to the next unit is an add with the holes filled so carries pass through them.
    const uint64_t xdep = xor_tau ? tile_base(pd, xor_tau) : 0ULL;   // index XOR between the members of a pair
    const uint64_t pivot_dep = xor_tau ? tile_base(pd, 1ULL << pivot) : 0ULL;
    uint64_t holes = pivot_dep;
    for (int j = 0; j < pd.t; ++j) holes |= 1ULL << pd.tile_bits[j];
    // Fused exchange: the tiles that leave drain at NVLink speed, the ones that stay at HBM speed; mixing them in one
    // CTA's three-stage ring makes every third stage wait for a slow drain.  So the grid is split: CTAs [0, send_ctas)
    // take the leaving tiles, the rest the staying ones (the exchanged bit becomes one more hole with a fixed value).
    uint64_t class_bits = 0;
    if (P.redirect && P.send_ctas > 0 && P.send_ctas < (int)gridDim.x && !((xdep >> P.redirect_bit) & 1ULL)) {
        const bool sender = (int)blockIdx.x < P.send_ctas;
        first = sender ? blockIdx.x : blockIdx.x - P.send_ctas;
        stride = sender ? P.send_ctas : gridDim.x - P.send_ctas;
        n_units /= 2;
        holes |= 1ULL << P.redirect_bit;
        class_bits = (uint64_t)(sender ? (P.redirect_keep ^ 1) : P.redirect_keep) << P.redirect_bit;
    }
    const uint64_t my_units = first < n_units ? (n_units - first + stride - 1) / stride : 0;
    const uint64_t n_my = xor_tau ? 2 * my_units : my_units;
    const uint64_t keep = ((1ULL << pd.n) - 1ULL) & ~holes;
    auto deposit = [&](uint64_t x) -> uint64_t {
        uint64_t r = 0;
        int j = 0;
        for (int b = 0; b < pd.n; ++b)
            if ((keep >> b) & 1ULL) { r |= ((x >> j) & 1ULL) << b; ++j; }
        return r;
    };
    const uint64_t ustride = deposit(stride), ufirst = deposit(first) | class_bits;
    auto next_unit = [&](uint64_t ub) -> uint64_t { return (((ub | holes) + ustride) & keep) | class_bits; };
    const uint32_t n_runs = 1u << pd.n_high;
    const uint32_t run_bytes = 16u << pd.L;
    const uint32_t n_instr = 1u << pd.tma_instr_bits;
    const uint32_t box_bytes = tile_bytes >> pd.tma_instr_bits;
    unsigned char* const gstate = reinterpret_cast<unsigned char*>(P.state);

    // ---- TMA duties of the elected thread --------------------------------------------------------------
    // (executed by every lane of warp 0: lane q issues instruction q, q+32, ...)
    const uint32_t lane = tid & 31u;
    auto issue_load = [&](uint64_t i, uint64_t base) {
        const int s = (int)(i % n_stages);
        if (lane == 0) mbar_expect_tx(&full[s], tile_bytes);
        __syncwarp();
        if (P.use_tensor_map) {
            for (uint32_t q = lane; q < n_instr; q += 32) {
                int c[5];
                tma_coords(pd, base + instr_offset(pd, q), c);
                tma_load_5d(tiles + (size_t)s * tile_bytes + (size_t)q * box_bytes, &tmap, c, &full[s]);
            }
        } else {
            for (uint32_t run = lane; run < n_runs; run += 32) {
                const uint64_t g = base + run_offset(pd, run);
                tma_load_1d(tiles + (size_t)s * tile_bytes + (size_t)run * run_bytes, gstate + g * 16, run_bytes, &full[s]);
            }
        }
    };
    auto issue_store = [&](uint64_t i, uint64_t base) {
        const int s = (int)(i % n_stages);
        if (xor_tau && !(i & 1) && !P.init_basis) {
            // this tile goes to its partner's location: the partner (item i+1) must have been read first
            const uint64_t j = i + 1;
            mbar_wait(&full[(int)(j % n_stages)], (uint32_t)((j / n_stages) & 1));
        }
        // fused qubit exchange: the whole tile stays (other buffer, same index) or leaves (partner GPU, bit flipped)
        const CUtensorMap* smap = &tmap;
        unsigned char* sdst = gstate;
        if (P.redirect) {
            const bool keep = (int)((base >> P.redirect_bit) & 1ULL) == P.redirect_keep;
            smap = keep ? &tmap_keep : &tmap_send;
            sdst = reinterpret_cast<unsigned char*>(keep ? P.dst_keep : P.dst_send);
            if (!keep) base ^= 1ULL << P.redirect_bit;
        }
        if (P.use_tensor_map) {
            for (uint32_t q = lane; q < n_instr; q += 32) {
                int c[5];
                tma_coords(pd, base + instr_offset(pd, q), c);
                tma_store_5d(smap, c, tiles + (size_t)s * tile_bytes + (size_t)q * box_bytes);
            }
        } else {
            for (uint32_t run = lane; run < n_runs; run += 32) {
                const uint64_t g = base + run_offset(pd, run);
                tma_store_1d(sdst + g * 16, tiles + (size_t)s * tile_bytes + (size_t)run * run_bytes, run_bytes);
            }
        }
        tma_store_commit();   // every lane commits its own (possibly empty) bulk group
    };

    // bases of the next item to load (warp 0) and of the item being computed
    uint64_t load_unit = ufirst, cur_unit = ufirst;
    uint64_t n_loaded = 0;
    auto load_next = [&]() {
        const bool odd = xor_tau && (n_loaded & 1);
        issue_load(n_loaded, odd ? (load_unit ^ xdep) : load_unit);
        if (!xor_tau || odd) load_unit = next_unit(load_unit);
        ++n_loaded;
    };
    if (warp == 0 && !P.init_basis) {
        const uint64_t pre = n_my < (uint64_t)n_stages ? n_my : (uint64_t)n_stages;
        for (uint64_t i = 0; i < pre; ++i) load_next();
    }
    // Basis-state input (P.init_basis): nothing is loaded.  A tile that does not contain |init_index> is all zero
    // before and after any op: those locations are zero-filled linearly by the whole grid (the memset this pass
    // replaces, minus the one tile below); the tile that does contain it is generated in shared memory and takes the
    // normal path.
    uint64_t tile_mask = 0;
    for (int j = 0; j < pd.t; ++j) tile_mask |= 1ULL << pd.tile_bits[j];
    const uint64_t init_tile = P.init_index & ~tile_mask;
    uint32_t init_local = 0;
    for (int j = 0; j < pd.t; ++j) init_local |= (uint32_t)((P.init_index >> pd.tile_bits[j]) & 1ULL) << j;
    if (P.init_basis == 1) {   // (init_basis == 2: the caller has zero-filled the buffer, e.g. with cudaMemsetAsync)
        const uint64_t dest_tile = init_tile ^ xdep;   // where that tile is written (deferred X gates on outer bits)
        const uint64_t n_amps = 1ULL << pd.n;
        uint4* out = reinterpret_cast<uint4*>(P.state);
        if (n_amps >= 4096 && pd.t == kMaxTileBits) {
            // 64 KiB at a time, as 1-D bulk copies from an all-zero stage (one instruction per chunk); the few chunks
            // that contain rows of the special tile are written element by element around them
            uint4* z = reinterpret_cast<uint4*>(tiles);
            for (uint32_t e = tid; e < 65536 / 16; e += kComputeThreads) z[e] = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async();
            __syncthreads();
            const uint64_t n_chunks = n_amps >> 12;
            for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
                const bool has_special = ((((c << 12) ^ dest_tile) & ~tile_mask) >> 12) == 0;   // (uniform)
                if (has_special) {
                    for (uint32_t e = tid; e < 4096; e += kComputeThreads) {
                        const uint64_t idx = (c << 12) + e;
                        if ((idx & ~tile_mask) != dest_tile) out[idx] = make_uint4(0u, 0u, 0u, 0u);
                    }
                } else if (tid == 0) {
                    tma_store_1d(gstate + (c << 16), tiles, 65536);
                    tma_store_commit();
                }
            }
            if (tid == 0) tma_store_wait_all();   // the stage is reused below
            __syncthreads();
        } else {
            const uint64_t step = (uint64_t)gridDim.x * kComputeThreads;
            for (uint64_t idx = (uint64_t)blockIdx.x * kComputeThreads + tid; idx < n_amps; idx += step)
                if ((idx & ~tile_mask) != dest_tile) out[idx] = make_uint4(0u, 0u, 0u, 0u);
        }
    }

    double ar[kSlots], ai[kSlots], br[kSlots], bi[kSlots];
    // Basis-state input: only the item that holds |init_index> has anything to do — go straight to it
    uint64_t i_begin = 0, i_end = n_my;
    if (P.init_basis) {
        uint64_t ub = init_tile;   // its unit: the member of the pair whose pivot bit is clear
        const bool odd = xor_tau && (ub & pivot_dep);
        if (odd) ub ^= xdep;
        uint64_t u = 0;            // unit number = the base with the holes squeezed out
        int j = 0;
        for (int b = 0; b < pd.n; ++b)
            if ((keep >> b) & 1ULL) { u |= ((ub >> b) & 1ULL) << j; ++j; }
        i_end = 0;
        if (u < n_units && u >= first && (u - first) % stride == 0 && !((P.init_index >> pd.n) != 0)) {
            const uint64_t iu = (u - first) / stride;
            i_begin = xor_tau ? 2 * iu + (odd ? 1 : 0) : iu;
            i_end = i_begin + 1;
            cur_unit = ub;
        }
    }
    for (uint64_t i = i_begin; i < i_end; ++i) {
        const int s = (int)(i % n_stages);
        const uint32_t parity = (uint32_t)((i / n_stages) & 1);
        const bool odd_item = xor_tau && (i & 1);
        const uint64_t tbase = odd_item ? (cur_unit ^ xdep) : cur_unit;
        if (!xor_tau || odd_item) cur_unit = next_unit(cur_unit);
        const uint64_t gbase = tbase | P.hi_bits;
        unsigned char* tile = tiles + (size_t)s * tile_bytes;
        if (P.init_basis) {
            if (tbase != init_tile) continue;   // (uniform over the CTA) zero-filled above
            uint4* z = reinterpret_cast<uint4*>(tile);
            for (uint32_t e = tid; e < tile_bytes / 16; e += kComputeThreads) z[e] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
            if (tid == 0) reinterpret_cast<double2*>(z)[init_local] = make_double2(1.0, 0.0);
            __syncthreads();
        }
        if (pd.n_phase > 0) {
            // factors of the fused diagonal runs that depend on index bits outside the tile: one (op, factor) per thread
            __syncthreads();   // the previous tile's readers of eu are done
            for (int idx = (int)tid; idx < pd.n_ops * 13; idx += kComputeThreads) {
                const int o = idx / 13, e = idx - o * 13;
                const DevOp& op = sops[o];
                if (op.kind != OP_PHASE) continue;
                const uint16_t* starts = reinterpret_cast<const uint16_t*>(op.m);
                const PhaseTerm* terms = P.phase_terms + op.cval_out;
                double fr = 1.0, fi = 0.0;
                for (int k = starts[e]; k < starts[e + 1]; ++k) {
                    const PhaseTerm t = terms[k];
                    bool on = (gbase >> t.o) & 1;
                    if (t.kind == 2) on = on && ((gbase >> t.j) & 1);
                    if (on) { const double r = fr * t.fr - fi * t.fi; fi = fr * t.fi + fi * t.fr; fr = r; }
                }
                eu[(size_t)op.tmask_out * 13 + e] = make_double2(fr, fi);
            }
            __syncthreads();
        }
        if (!P.init_basis) mbar_wait(&full[s], parity);

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
                apply_op(sops[o], opcode, sm, tid, gbase, P.phase_tables, eu, sd, base_local, ar, ai, br, bi);
                ++o;
                if (!fetch()) { in_b = true; break; }
                apply_op(sops[o], opcode, sm, tid, gbase, P.phase_tables, eu, sd, base_local, br, bi, ar, ai);
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
        // make the generic-proxy writes visible to the bulk-copy engine; then the elected thread stores
        // this tile and refills the stage whose store (tile i-1) has drained
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            issue_store(i, tbase ^ xdep);
            if (!P.init_basis && i >= 1 && n_loaded < n_my) {
                tma_store_wait_read_1();   // all but this lane's newest store group have finished reading shared memory
                __syncwarp();              // ... for every lane: the stage of tile i-1 is free
                load_next();               // item (i - 1) + n_stages
            }
        }
    }
    if (warp == 0) {
        tma_store_wait_all();
        if (P.redirect) {
            // half of the tiles went to the partner GPU's memory through the async proxy: order those writes before
            // anything the partner is told after this kernel (the ranks' barrier follows in the stream)
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence_system();
        }
    }
}

// worst case with three full stages must fit (pick_stages never has to go below the three the partner-tile store needs)
static_assert(3 * (16 << kMaxTileBits) + (kMaxOpsPerPass + 1) * sizeof(DevOp) + kMaxPhaseOps * 13 * sizeof(double2) +
                      2 * 3 * sizeof(uint64_t) + (kMaxSweeps + 2) * kComputeThreads * sizeof(uint16_t) <=
                  (size_t)kMaxDynamicSmem,
              "shared-memory budget of a pass");

size_t pass_smem_bytes(const PassDesc& pd, int stages) {
    return (size_t)stages * ((size_t)16 << pd.t) + ((size_t)pd.n_ops + 1) * sizeof(DevOp) + (size_t)pd.n_phase * 13 * sizeof(double2) +
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
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

cudaError_t launch_pass(const PassParams& params_in, int num_sms, cudaStream_t stream) {
    PassParams params = params_in;
    alignas(64) CUtensorMap tmap, tmap_keep, tmap_send;
    std::memset(&tmap, 0, sizeof(tmap));
    if (params.use_tensor_map && !encode_tensor_map(params, params.state, &tmap)) return cudaErrorInvalidValue;
    tmap_keep = tmap;
    tmap_send = tmap;
    if (params.redirect) {
        if (!params.dst_keep || !params.dst_send || params.init_basis) return cudaErrorInvalidValue;
        for (int j = 0; j < params.pd.t; ++j)
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
    fused_pass_kernel<<<(unsigned)grid, kComputeThreads, smem, stream>>>(params, tmap, tmap_keep, tmap_send);
    return cudaGetLastError();
}

}  // namespace b200
}  // namespace qsim
