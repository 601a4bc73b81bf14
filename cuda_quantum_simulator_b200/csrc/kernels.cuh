// Device-side entry points shared by the host classes (internal header).
#pragma once

#include <cuda_runtime.h>
#include <cuComplex.h>

#include <cstdint>

#include "program.hpp"

namespace qsim {
namespace b200 {

constexpr int kMaxDynamicSmem = 227 * 1024;

struct PassParams {
    cuDoubleComplex* state;   // this GPU's amplitudes (2^pd.n of them)
    const DevOp* ops;         // device copy of this pass's ops
    const double2* phase_tables;    // this pass's OP_PHASE tables (kPhaseTableSize entries each)
    const PhaseTerm* phase_terms;   // this pass's OP_PHASE outside-bit terms
    uint64_t hi_bits;         // rank << n_local for a sharded state, else 0 (only used by controls)
    uint64_t n_tiles;         // 2^(pd.n - pd.t)
    int32_t stages;           // depth of the shared-memory ring
    int32_t use_tensor_map;   // 1: cp.async.bulk.tensor boxes (default); 0: one 1-D bulk copy per contiguous run
    int32_t init_basis;       // 1 / 2: the memory holds nothing yet / only zeros; the input state is the basis state |init_index>:
    int32_t pad;              //    tiles are generated on chip, all-zero tiles are stored without interpretation
    uint64_t init_index;
    // Fused qubit exchange (sharded states): this pass stores OUT OF PLACE.  Tiles whose index bit `redirect_bit`
    // equals `redirect_keep` go to dst_keep at the same index, the others to dst_send (the partner GPU's buffer,
    // peer-mapped) at index ^ (1 << redirect_bit).  redirect_bit is never a tile bit of such a pass.
    int32_t redirect;
    int32_t redirect_bit;
    int32_t redirect_keep;
    int32_t send_ctas;        // > 0: CTAs [0, send_ctas) take the tiles that leave, the others the tiles that stay
    cuDoubleComplex* dst_keep;
    cuDoubleComplex* dst_send;
    PassDesc pd;
};
static_assert(sizeof(PassParams) <= 4000, "kernel parameter space");

constexpr int kMaxStages = 8;
size_t pass_smem_bytes(const PassDesc& pd, int stages);
int pick_stages(const PassDesc& pd, int wanted);
cudaError_t launch_pass(const PassParams& params, int num_sms, cudaStream_t stream);

}  // namespace b200
}  // namespace qsim
