// Device-side building blocks of the pass kernels (mbarrier, TMA, shared-memory vector access, tile addressing), shared by
// the ahead-of-time interpreter kernel (kernels_pass.cu) and the run-time specialised kernels: jit.cpp hands this very text
// to NVRTC, so it must not include anything.
#pragma once

#ifdef __CUDACC_RTC__
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };   // cuda.h's type: 128 opaque bytes
#endif

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
__device__ __forceinline__ void tma_store_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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

// cross-GPU handshake words of the in-place fused exchange (PassParams::hs_*).  RELAXED system-scope accesses on purpose: the
// word only says "my copy of that tile has arrived in my shared memory" (the mbarrier has been observed), nothing the writer
// stored has to become visible with it, and a release store (MEMBAR.SYS) would first wait for every TMA store this SM has in
// flight over NVLink - one tile per round trip (measured: 280 GB/s instead of 700).
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
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

}  // namespace
}  // namespace b200
}  // namespace qsim
