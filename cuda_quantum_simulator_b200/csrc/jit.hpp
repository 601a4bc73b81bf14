// Run-time specialisation of the pass kernel (internal header).
//
// The ahead-of-time kernel (kernels_pass.cu) INTERPRETS a pass's op list: per op and tile it fetches the record from
// shared memory, decodes it, jumps through a table and selects on per-slot control masks — about half of the issued
// instructions of a heavy pass are that bookkeeping (ncu, profiles/ncu_fused_pass_c2_30q_r01.json).  For large states a
// pass is worth a kernel of its own: jit.cpp turns (PassDesc, op list) into straight-line CUDA C++ for the per-tile
// compute — layouts, targets, control masks and slot sets are literals, bit flips on register bits are renamings,
// controls become branches or static slot subsets — wraps it in the SAME skeleton (pass_kernel_body.inc: TMA ring, tile
// stepping, basis-state input, fused exchange) and compiles it with NVRTC for sm_100a.  Matrix entries stay run-time data
// (read from the staged op records), so the kernel depends only on the STRUCTURE of the pass: a variational loop that
// re-runs a circuit with new angles hits the cache.  Compiled kernels are also kept on disk ($QSIM_JIT_CACHE, default
// ~/.cache/qsim_b200/jit, "off" disables), keyed by the generated source and the embedded skeleton.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>

#include "pass_desc.h"

namespace qsim {
namespace b200 {

struct JitKernel;   // a loaded, specialised pass kernel

enum class JitMode : int { Off = 0, Auto = 1, Always = 2 };

// QSIM_JIT=off|auto|always (default auto), QSIM_JIT_MIN_QUBITS (default 26: a pass over 2^26 amplitudes takes ~0.4 ms,
// a compile ~0.5 s and is cached per pass structure).  Always: every pass, and a compile failure throws (tests).
JitMode jit_mode();
int jit_min_qubits();
void jit_set_mode(JitMode mode, int min_qubits);

// Should this pass run specialised?  (mode, size of the state, NVRTC present)
bool jit_wanted(const PassDesc& pd);

// CUDA C++ of the specialised per-tile compute for this pass (the generated part only).
// dual = true: the TWO-WARP-GROUP build of the kernel (pass_kernel_body.inc, QSIM_DUAL_GROUPS): the 16 warps work as two
// groups of 8 on two tiles at once, so that one group's shared-memory / shuffle / barrier phases overlap the other's FP64
// phases.  For compute-heavy passes (jit_dual_wanted); HBM-bound passes keep the one-group build, whose ring keeps two
// tiles in flight instead of one.
std::string jit_generate_compute(const PassDesc& pd, const DevOp* host_ops, bool dual = false);
// The whole translation unit handed to NVRTC (generated part + the three embedded sources).
std::string jit_translation_unit(const PassDesc& pd, const DevOp* host_ops, bool dual = false);
// Full 12-bit tiles swept by all threads in every sweep, and enough FP64 work per tile that the pass is bound by the SM, not
// by HBM (QSIM_DUAL=off|auto|always, QSIM_DUAL_MIN_FP64: FP64 instructions per tile and thread, default 350).
bool jit_dual_possible(const PassDesc& pd);
bool jit_dual_wanted(const PassDesc& pd, const DevOp* host_ops);
bool jit_dual_autotune();   // QSIM_DUAL_AUTOTUNE=0 turns the measured choice off (the estimate alone decides)
void jit_set_dual(int mode, int min_fp64);   // mode 0 off / 1 auto / 2 always, -1 keeps; min_fp64 < 0 keeps
int jit_fp64_estimate(const PassDesc& pd, const DevOp* host_ops);

// The generated source of a pass and its cache key (the expensive part of a lookup; kept per program and pass).
struct JitRequest;
std::shared_ptr<JitRequest> jit_make_request(const PassDesc& pd, const DevOp* host_ops, bool dual = false);
// The kernel if it is ready (process cache, on-disk cache).  Otherwise: async = false compiles now; async = true queues the
// compile on a background thread and returns nullptr with *pending = true - the caller launches the interpreter kernel this
// time and asks again at the next launch, so run() never waits for NVRTC (QSIM_JIT_ASYNC=0 turns this off).  nullptr with
// *pending = false: unavailable (NVRTC missing or the compile failed; logged once).  Mode always: synchronous, failures throw.
std::shared_ptr<JitKernel> jit_lookup(const JitRequest& rq, bool needs_device, bool async, bool* pending);
bool jit_async_enabled();
void jit_wait_all();   // blocks until every queued compile has finished
// Before the process exits: queued compiles are dropped, compiles in progress get to finish, no new ones start (a process
// must not reach the exit handlers of NVRTC while a background thread is still compiling).  Idempotent.
void jit_shutdown();

// Compile (or fetch from the process-wide cache) the kernel of this pass.  Returns nullptr when NVRTC is unavailable or
// the compile failed in Auto mode (logged once; the caller uses the interpreter kernel); throws in Always mode.
// needs_device = false only compiles to a cubin (used by the CPU-side build check); such a kernel cannot be launched.
std::shared_ptr<JitKernel> jit_get_kernel(const PassDesc& pd, const DevOp* host_ops, bool needs_device = true, bool dual = false);

// Launch with the interpreter kernel's parameters.
cudaError_t jit_launch(JitKernel& k, const PassParams& params, const void* tmap, const void* tmap_keep, const void* tmap_send,
                       unsigned grid, size_t smem, cudaStream_t stream);

// Copies the kernel's cubin (at most cap bytes); returns its size.
size_t jit_copy_cubin(const JitKernel& k, void* out, size_t cap);

struct JitStats {
    int64_t compiles = 0, cache_hits = 0, launches = 0, failures = 0, disk_hits = 0;
    double compile_seconds = 0;
    int64_t last_cubin_bytes = 0;
    int last_registers = 0;
};
JitStats jit_stats();
std::string jit_last_log();   // NVRTC log of the last compile (warnings / errors)

}  // namespace b200
}  // namespace qsim
