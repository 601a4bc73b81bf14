// Fused-pass kernel for sm_100a: ONE read and ONE write of the state vector per pass, however
// many gates the pass carries (the reference sweeps HBM once per gate, src/Simulator.cu:48-154 and
// src/Gates.cu:31-410 of the reference).
//
// Structure (persistent, one CTA per SM, 9 warps):
//   warp 8     TMA producer/consumer of global memory.  For its CTA's tiles it issues 1-D bulk
//              copies (cp.async.bulk, SASS UBLKCP) global -> shared into a 3-stage ring, signalled
//              through mbarrier complete_tx; after the compute warps release a stage it issues the
//              bulk copies shared -> global for that stage and immediately refills it.
//   warps 0-7  compute.  A tile is 2^t amplitudes (t <= 12: 64 KiB): the low L index bits (one
//              contiguous run of 16 << L bytes) x (t - L) arbitrary high "tile qubits".  Per sweep
//              every thread pulls 2^r amplitudes (r <= 4) from shared memory into registers with
//              128-bit conflict-free loads, runs the sweep's op list — register-resident targets
//              in-thread, lane-resident targets with __shfl_xor butterflies, diagonal gates as a
//              single complex multiply wherever their qubits live — and writes back.
//
// Algorithmic traffic: 2 * 16 * 2^n bytes per pass.
#include "kernels.cuh"

#include <cstdio>

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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void compute_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
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

template <int J, int KIND>
__device__ __forceinline__ void reg_pairs_k(const DevOp& op, uint32_t sm, double (&ar)[16], double (&ai)[16]) {
    const double m00r = op.m[0], m00i = op.m[1], m01r = op.m[2], m01i = op.m[3];
    const double m10r = op.m[4], m10i = op.m[5], m11r = op.m[6], m11i = op.m[7];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k & (1 << J)) continue;   // k enumerates the slots whose target bit is 0
        const int k1 = k | (1 << J);
        if (!((sm >> k) & 1)) continue;
        const double xr = ar[k], xi = ai[k], yr = ar[k1], yi = ai[k1];
        if (KIND == OP_FLIP) {
            ar[k] = yr; ai[k] = yi; ar[k1] = xr; ai[k1] = xi;
        } else if (KIND == OP_ADIAG) {
            ar[k] = m01r * yr - m01i * yi; ai[k] = m01r * yi + m01i * yr;
            ar[k1] = m10r * xr - m10i * xi; ai[k1] = m10r * xi + m10i * xr;
        } else if (KIND == OP_MATREAL) {
            ar[k] = m00r * xr + m01r * yr; ai[k] = m00r * xi + m01r * yi;
            ar[k1] = m10r * xr + m11r * yr; ai[k1] = m10r * xi + m11r * yi;
        } else {
            ar[k] = m00r * xr - m00i * xi + m01r * yr - m01i * yi;
            ai[k] = m00r * xi + m00i * xr + m01r * yi + m01i * yr;
            ar[k1] = m10r * xr - m10i * xi + m11r * yr - m11i * yi;
            ai[k1] = m10r * xi + m10i * xr + m11r * yi + m11i * yr;
        }
    }
}

template <int J>
__device__ __forceinline__ void reg_pairs(const DevOp& op, uint32_t sm, double (&ar)[16], double (&ai)[16]) {
    switch (op.kind) {
        case OP_FLIP: reg_pairs_k<J, OP_FLIP>(op, sm, ar, ai); break;
        case OP_ADIAG: reg_pairs_k<J, OP_ADIAG>(op, sm, ar, ai); break;
        case OP_MATREAL: reg_pairs_k<J, OP_MATREAL>(op, sm, ar, ai); break;
        default: reg_pairs_k<J, OP_MAT>(op, sm, ar, ai); break;
    }
}

__device__ __forceinline__ void lane_target(const DevOp& op, uint32_t sm, uint32_t tid, double (&ar)[16],
                                            double (&ai)[16]) {
    const int lm = 1 << op.tbit;
    const bool b = (tid >> op.tbit) & 1;
    // coefficient of my own amplitude and of my partner's
    const double cor = b ? op.m[6] : op.m[0], coi = b ? op.m[7] : op.m[1];
    const double cpr = b ? op.m[4] : op.m[2], cpi = b ? op.m[5] : op.m[3];
    const uint8_t kind = op.kind;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double pr = shfl_xor_f64(ar[k], lm), pi = shfl_xor_f64(ai[k], lm);
        if (!((sm >> k) & 1)) continue;
        const double xr = ar[k], xi = ai[k];
        if (kind == OP_FLIP) { ar[k] = pr; ai[k] = pi; }
        else if (kind == OP_ADIAG) { ar[k] = cpr * pr - cpi * pi; ai[k] = cpr * pi + cpi * pr; }
        else if (kind == OP_MATREAL) { ar[k] = cor * xr + cpr * pr; ai[k] = cor * xi + cpr * pi; }
        else {
            ar[k] = cor * xr - coi * xi + cpr * pr - cpi * pi;
            ai[k] = cor * xi + coi * xr + cpr * pi + cpi * pr;
        }
    }
}

__device__ __forceinline__ void diagonal(const DevOp& op, uint32_t sm, uint32_t tid, uint64_t gbase, double (&ar)[16],
                                         double (&ai)[16]) {
    uint32_t tsl;
    if (op.thome == T_REG) tsl = op.tslots;
    else {
        const bool b = (op.thome == T_THREAD) ? ((tid & op.tmask_thr) != 0) : ((gbase & op.tmask_out) != 0);
        tsl = b ? 0xffffu : 0u;
    }
    const double d0r = op.m[0], d0i = op.m[1], d1r = op.m[6], d1i = op.m[7];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (!((sm >> k) & 1)) continue;
        const bool b = (tsl >> k) & 1;
        const double pr = b ? d1r : d0r, pi = b ? d1i : d0i;
        const double xr = ar[k], xi = ai[k];
        ar[k] = xr * pr - xi * pi;
        ai[k] = xr * pi + xi * pr;
    }
}

}  // namespace

__global__ void __launch_bounds__(kPassThreads, 1) fused_pass_kernel(const __grid_constant__ PassParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const PassDesc& pd = P.pd;
    const uint32_t tile_bytes = 16u << pd.t;
    unsigned char* tiles = smem;
    DevOp* sops = reinterpret_cast<DevOp*>(smem + kStages * tile_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(sops + pd.n_ops);
    uint64_t* done = full + kStages;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // stage this pass's op list in shared memory (broadcast reads later)
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.ops);
        uint4* dst = reinterpret_cast<uint4*>(sops);
        const int n16 = pd.n_ops * (int)(sizeof(DevOp) / 16);
        for (int i = tid; i < n16; i += kPassThreads) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&done[s], kComputeWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t n_tiles = P.n_tiles;
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t n_my = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
    const uint32_t n_runs = 1u << pd.n_high;
    const uint32_t run_bytes = 16u << pd.L;
    unsigned char* const gstate = reinterpret_cast<unsigned char*>(P.state);

    if (warp == kComputeWarps) {
        // ===================== TMA warp =====================
        auto issue_load = [&](uint64_t i) {
            const int s = (int)(i % kStages);
            const uint64_t base = tile_base(pd, first + i * stride);
            if (lane == 0) mbar_expect_tx(&full[s], tile_bytes);
            __syncwarp();
            for (uint32_t run = lane; run < n_runs; run += 32) {
                const uint64_t g = base + run_offset(pd, run);
                tma_load_1d(tiles + (size_t)s * tile_bytes + (size_t)run * run_bytes, gstate + g * 16, run_bytes,
                            &full[s]);
            }
        };
        const uint64_t pre = n_my < (uint64_t)kStages ? n_my : (uint64_t)kStages;
        for (uint64_t i = 0; i < pre; ++i) issue_load(i);
        for (uint64_t i = 0; i < n_my; ++i) {
            const int s = (int)(i % kStages);
            const uint32_t parity = (uint32_t)((i / kStages) & 1);
            mbar_wait(&done[s], parity);
            const uint64_t base = tile_base(pd, first + i * stride);
            for (uint32_t run = lane; run < n_runs; run += 32) {
                const uint64_t g = base + run_offset(pd, run);
                tma_store_1d(gstate + g * 16, tiles + (size_t)s * tile_bytes + (size_t)run * run_bytes, run_bytes);
            }
            tma_store_commit();
            tma_store_wait_read();   // shared memory of this stage may be overwritten now
            __syncwarp();
            if (i + kStages < n_my) issue_load(i + kStages);
        }
        tma_store_wait_all();
    } else {
        // ===================== compute warps =====================
        double ar[16], ai[16];
        for (uint64_t i = 0; i < n_my; ++i) {
            const int s = (int)(i % kStages);
            const uint32_t parity = (uint32_t)((i / kStages) & 1);
            const uint64_t gbase = tile_base(pd, first + i * stride) | P.hi_bits;
            unsigned char* tile = tiles + (size_t)s * tile_bytes;
            mbar_wait(&full[s], parity);

#pragma unroll 1
            for (int sw = 0; sw < pd.n_sweeps; ++sw) {
                const SweepDesc& sd = pd.sweep[sw];
                if (sw > 0) compute_barrier();
                const uint32_t n_active = 1u << sd.nthr;
                const bool warp_active = (warp << 5) < n_active;
                if (!warp_active) continue;
                const bool active = tid < n_active;
                const int slots = 1 << sd.r;
                uint32_t base_local = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (b < sd.nthr && ((tid >> b) & 1)) base_local |= 1u << sd.thr_pos[b];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (active && k < slots) {
                        const double2 v =
                            *reinterpret_cast<const double2*>(tile + (size_t)(base_local + sd.slot_off[k]) * 16);
                        ar[k] = v.x;
                        ai[k] = v.y;
                    } else {
                        ar[k] = 0.0;
                        ai[k] = 0.0;
                    }
                }
#pragma unroll 1
                for (int o = sd.op_begin; o < sd.op_end; ++o) {
                    const DevOp& op = sops[o];
                    if ((gbase & op.cmask_out) != op.cval_out) continue;
                    const bool thr_ok = (tid & op.cmask_thr) == op.cval_thr;
                    const uint32_t sm = thr_ok ? (uint32_t)op.slotmask : 0u;
                    if (op.kind == OP_DIAG) {
                        diagonal(op, sm, tid, gbase, ar, ai);
                    } else if (op.thome == T_LANE) {
                        lane_target(op, sm, tid, ar, ai);
                    } else {
                        switch (op.tbit) {
                            case 0: reg_pairs<0>(op, sm, ar, ai); break;
                            case 1: reg_pairs<1>(op, sm, ar, ai); break;
                            case 2: reg_pairs<2>(op, sm, ar, ai); break;
                            default: reg_pairs<3>(op, sm, ar, ai); break;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (active && k < slots)
                        *reinterpret_cast<double2*>(tile + (size_t)(base_local + sd.slot_off[k]) * 16) =
                            make_double2(ar[k], ai[k]);
                }
            }
            // make the generic-proxy writes visible to the bulk-copy engine, then hand the stage over
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[s]);
        }
    }
}

size_t pass_smem_bytes(const PassDesc& pd) {
    return (size_t)kStages * ((size_t)16 << pd.t) + (size_t)pd.n_ops * sizeof(DevOp) + 2 * kStages * sizeof(uint64_t);
}

cudaError_t launch_pass(const PassParams& params, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e =
            cudaFuncSetAttribute(fused_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynamicSmem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const size_t smem = pass_smem_bytes(params.pd);
    if (smem > (size_t)kMaxDynamicSmem) return cudaErrorInvalidValue;
    uint64_t grid = params.n_tiles < (uint64_t)num_sms ? params.n_tiles : (uint64_t)num_sms;
    fused_pass_kernel<<<(unsigned)grid, kPassThreads, smem, stream>>>(params);
    return cudaGetLastError();
}

}  // namespace b200
}  // namespace qsim
