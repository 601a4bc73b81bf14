// Plain-data description of a compiled pass, shared VERBATIM by the host compiler (program.hpp), the ahead-of-time
// interpreter kernel (kernels_pass.cu) and the run-time specialised kernels (jit.cpp hands this very text to NVRTC, so
// the kernel-parameter layouts cannot drift apart).  No includes beyond fixed-width integers; no host-only types.
#pragma once

#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned long long size_t_rtc;
#else
#include <cstdint>
#endif

namespace qsim {
namespace b200 {

constexpr int kMaxTileBits = 12;         // 2^12 amplitudes * 16 B = 64 KiB per pipeline stage
#ifndef QSIM_REG_BITS
#define QSIM_REG_BITS 3                  // register bits per thread: 3 -> 16 warps x 8 amplitudes, 4 -> 8 warps x 16
#endif
constexpr int kMaxRegBits = QSIM_REG_BITS;
constexpr int kSlots = 1 << kMaxRegBits;                                  // amplitudes per thread
constexpr int kComputeWarps = 1 << (kMaxTileBits - kMaxRegBits - 5);      // a full tile is one sweep of all warps
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kMaxSweeps = 12;
constexpr int kMaxSegments = 14;
constexpr int kMaxOpsPerPass = 120;      // ops of one pass are staged in shared memory (with kMaxSweeps and kMaxPhaseOps this
                                         // keeps three 64 KiB stages + tables within the 227 KiB of shared memory)
constexpr int kMaxPhaseOps = 12;         // fused diagonal runs per pass (13 complex factors each in shared memory, two copies:
                                         // the next tile's factors are computed while the current tile's are in use)
constexpr int kPhaseTableSize = 1 << kMaxTileBits;   // one complex factor per tile-local index

enum OpKind : uint8_t {
    OP_MAT = 0,      // dense complex 2x2
    OP_MATREAL = 1,  // real 2x2 (H, Ry, ...): half the flops
    OP_ADIAG = 2,    // anti-diagonal [[0,b],[c,0]] (Y, X*diag)
    OP_FLIP = 3,     // bit flip (X / CNOT / Toffoli): pure data movement
    OP_DIAG = 4,     // diagonal diag(d0, d1): phase by target bit, target may live anywhere
    OP_PHASE = 5,    // a fused RUN of diagonal gates (<= 1 control each): one complex factor per amplitude,
                     //   f(idx) = TABLE[tile-local idx] * U(tile) * prod_{tile bits j set} E_j(tile)
};

enum TargetHome : uint8_t {
    T_LANE = 0,      // target is tid bit `tbit` < 5: partner amplitude comes by __shfl_xor
    T_REG = 1,       // target is register bit `tbit`: partner is another slot of the same thread
    T_THREAD = 2,    // (diagonal only) target is a tid bit, no partner needed
    T_OUTSIDE = 3,   // (diagonal only) target is not a tile bit: uniform for the whole tile
};

// Device-visible op record (128 bytes, read as broadcast from shared memory).
struct alignas(16) DevOp {
    uint8_t kind;
    uint8_t thome;
    uint8_t tbit;
    uint8_t opcode;       // dense dispatch code, see op_code(): (kind, where the target lives, has in-tile controls)
    uint16_t slotmask;    // register slots whose register-resident control bits are satisfied
    uint16_t tslots;      // OP_DIAG with T_REG: slots whose target bit is 1
    uint32_t cmask_thr;   // controls held in tid bits
    uint32_t cval_thr;
    uint32_t tmask_thr;   // OP_DIAG with T_THREAD: the tid bit of the target
    uint32_t has_out;     // 1 if the op depends on index bits outside the tile (cmask_out / tmask_out)
    uint64_t cmask_out;   // controls outside the tile, as a mask over the global amplitude index
    uint64_t cval_out;
    uint64_t tmask_out;   // OP_DIAG with T_OUTSIDE: global-index bit of the target
    double m[8];          // m00.re m00.im m01.re m01.im m10.re m10.im m11.re m11.im
    double pad3[2];
};
static_assert(sizeof(DevOp) == 128, "DevOp layout");

// Dense dispatch codes (the interpreter's switch is one jump table):
// opcode = kind * kOpcodesPerKind + home * 2 + ctrl for the pair-wise kinds (home: 0 = lane, 1 + j = register bit j),
//          kOpcodeDiag + home * 2 + ctrl for OP_DIAG (home: 0 = register-resident target, 1 = thread/outside target)
constexpr int kOpcodesPerKind = 2 * (1 + kMaxRegBits);
constexpr uint8_t kOpcodeDiag = 4 * kOpcodesPerKind;
constexpr uint8_t kOpcodePhase = kOpcodeDiag + 4;
constexpr uint8_t kOpcodeCopy = kOpcodeDiag + 5;   // not produced by the compiler: the kernel's stand-in for a skipped op
constexpr int kNumOpcodes = kOpcodeDiag + 6;
inline uint8_t op_code(uint8_t kind, uint8_t thome, uint8_t tbit, bool ctrl) {
    if (kind == OP_DIAG) return (uint8_t)(kOpcodeDiag + (thome == T_REG ? 0 : 2) + (ctrl ? 1 : 0));
    const int home = thome == T_LANE ? 0 : 1 + tbit;
    return (uint8_t)(kind * kOpcodesPerKind + home * 2 + (ctrl ? 1 : 0));
}

// OP_PHASE records reuse DevOp fields: cmask_out = first entry of the op's table in the pass's table blob (tmask_thr = the
// tile bits that table really depends on, cval_thr = first entry of the COMPACT copy indexed by just those bits),
// cval_out = first term in the pass's term array, tmask_out = slot of its (E_0..E_11, U) factors in shared
// memory, tslots = number of terms; m[] viewed as uint16[14]: term range start of E_0..E_11, U, end.
// A term multiplies one of those 13 factors when one (kind 0/1) or two (kind 2) index bits OUTSIDE the tile are set.
struct PhaseTerm {
    uint8_t kind;     // 0: E_j *= f if bit o;  1: U *= f if bit o;  2: U *= f if bits o and j
    uint8_t o;        // global index bit (may be a rank bit of a sharded state)
    uint8_t j;        // kind 0: tile-local bit; kind 2: the second global bit
    uint8_t pad[5];
    double fr, fi;
};
static_assert(sizeof(PhaseTerm) == 24, "PhaseTerm layout");

struct SweepDesc {
    uint16_t op_begin, op_end;   // indices into the pass's op array
    uint8_t r;                   // register bits in use (slots = 1 << r)
    uint8_t nthr;                // tid bits in use (active threads = 1 << nthr)
    uint8_t thr_pos[12];         // tile-local bit position held by tid bit i
    uint8_t reg_pos[4];          // tile-local bit position held by register bit j
    uint16_t slot_off[16];       // tile-local index offset of register slot k
    uint16_t pad;
    // Flips (without controls outside the tile) that lead a LATER sweep fold into that sweep's load like the pass's
    // leading flips fold into the first one (PassDesc::head_lin): inverse map, slot offsets already mapped.
    uint16_t n_head;
    uint16_t head_const;
    uint16_t head_lin[kMaxTileBits];
    uint16_t load_slot_off[16];
};

struct Segment {                 // tile number -> global base index, one contiguous run of outer bits
    uint64_t mask;               // applied after the shift
    uint8_t src_shift, dst_shift;
    uint8_t pad[6];
};

// One dimension of the pass's tensor map (cp.async.bulk.tensor): index bits [start_bit,
// start_bit + range_bits); the box covers the lowest box_bits of that range (the tile bits), the
// remaining bits of the range are coordinate (outer / per-instruction) bits.
struct TmaDim {
    uint8_t start_bit, range_bits, box_bits, pad;
};

// Controlled bit flips (X / CNOT) that can slide to the end of a pass are not executed as ops: together they are
// an affine map over GF(2) of the tile-local index, l -> A l ^ c, applied by the final store's addressing.  A flip
// whose control lies outside the tile is a translation that fires per tile (TailDyn); flips with two or more
// controls inside the tile are not affine and stay ops.
struct TailDyn {
    uint64_t cmask_out, cval_out;       // controls outside the tile (global index bits)
    uint16_t w;                         // tile-local index XOR when they match (already pushed through later flips)
    uint16_t pad[3];
};
constexpr int kMaxTailFlips = 16;
constexpr int kMaxTailDyn = 6;

struct PassDesc {
    int32_t n;                   // qubits held in this buffer (local qubits when sharded)
    int32_t t;                   // tile bits
    int32_t L;                   // low contiguous tile bits: runs of 16 << L bytes
    int32_t n_sweeps;
    int32_t n_ops;
    int32_t op_offset;           // into Program::ops
    int32_t n_segments;
    int32_t n_high;              // t - L
    int32_t n_phase;             // OP_PHASE ops in this pass
    int32_t phase_table_offset;  // into Program::phase_tables (entries)
    int32_t phase_term_offset;   // into Program::phase_terms
    int32_t n_tail;              // trailing flips folded into the final store (0: the map below is the identity)
    int32_t n_dyn;               // of which translations that depend on the tile
    uint8_t tile_bits[kMaxTileBits];   // global bit of tile-local bit i (ascending)
    uint32_t xor_local;          // tile-local index XOR applied by the pass's final store (deferred X gates)
    uint64_t xor_tau;            // tile-number XOR: the tile read from tau is written to tau ^ xor_tau
    uint8_t tma_instr_bits;      // the top tma_instr_bits tile bits are enumerated by separate TMA instructions
    uint8_t pad[3];
    TmaDim tma_dim[5];
    uint16_t tail_lin[kMaxTileBits];   // A: image of tile-local bit j
    uint16_t tail_const;               // c
    uint16_t pad3;
    uint16_t store_slot_off[16];       // A applied to the last sweep's slot_off
    TailDyn dyn[kMaxTailDyn];
    // Leading flips, folded the same way into the FIRST sweep's load: with F(x) = A x ^ c (^ translations) their
    // combined index map, the element that belongs at tile-local index l is read from F^-1(l).  The fields hold the
    // inverse map: head_lin = A^-1, head_const = A^-1 c, head_dyn[].w = A^-1 w.
    int32_t n_head;
    int32_t n_head_dyn;
    uint16_t head_lin[kMaxTileBits];
    uint16_t head_const;
    uint16_t pad4;
    uint16_t load_slot_off[16];        // A^-1 applied to the first sweep's slot_off
    TailDyn head_dyn[kMaxTailDyn];
    Segment seg[kMaxSegments];
    SweepDesc sweep[kMaxSweeps];
    // derived by the compiler so that the kernel's set-up does not walk the arrays above bit by bit
    uint64_t tile_mask;          // OR of 1 << tile_bits[j]
    uint64_t xdep;               // tile_base(xor_tau): index XOR between the two tiles of a pair
    uint64_t pivot_dep;          // tile_base(highest bit of xor_tau)
};


constexpr int kMaxDynamicSmem = 227 * 1024;
constexpr int kMaxStages = 8;

// Kernel parameters of one pass launch (both kernels).  Pointers are plain: double2 = cuDoubleComplex.
struct PassParams {
    void* state;              // this GPU's amplitudes (2^pd.n cuDoubleComplex)
    const DevOp* ops;         // device copy of this pass's ops
    const void* phase_tables; // this pass's OP_PHASE tables (kPhaseTableSize double2 entries each)
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
    void* dst_keep;
    void* dst_send;
    // redirect == 3 / 4: the in-place exchange SPLIT over the two passes around it.  `split_bit` (another index bit that is a
    // tile bit of neither pass) halves the leaving tiles: the pass before the exchange scatters the half with that bit clear
    // (3: as redirect == 2, restricted to those tiles; the other leaving tiles are stored in place), the pass after it GATHERS
    // the half with the bit set (4: those tiles are LOADED from the partner's live shard at index ^ (1 << redirect_bit) and
    // stored in place, under the same handshake: my store over a tile waits until the partner has loaded it).  Three CTA
    // classes: [0, send_ctas) the exchanged quarter, [send_ctas, send_ctas + mid_ctas) the staying half, the rest the
    // leaving quarter this pass does not move.
    int32_t split_bit;
    int32_t mid_ctas;
    // redirect == 2: the fused exchange IN PLACE (no second buffer: dst_keep is `state` itself, dst_send the partner's live
    // buffer).  A leaving tile lands on the partner's own leaving tile of the same item number (both ranks run the same
    // grid over the same item order), so CTA c may store its item i only after the partner's CTA c has LOADED its item i:
    // every sender CTA publishes hs_base + (items loaded) in the partner's hs array (hs_peer[c], a relaxed system-scope
    // store over NVLink) and polls its own (hs_local[c]).  hs_base grows from exchange to exchange, so the words never need a reset.
    // A poll that sees nothing for hs_timeout_ns sets *hs_error and goes on (wrong data, reported by the host: no hang).
    unsigned long long* hs_local;
    unsigned long long* hs_peer;
    unsigned long long hs_base;
    unsigned long long hs_timeout_ns;
    int32_t* hs_error;
    // development aid (QSIM_PASS_TIMELINE=1): CTA 0's %globaltimer at eight points of the kernel, null otherwise
    unsigned long long* timeline;
    unsigned long long* progress;   // (1024 words: per CTA and warp group, the item it has reached)
    PassDesc pd;
};
static_assert(sizeof(PassParams) <= 4000, "kernel parameter space");

}  // namespace b200
}  // namespace qsim
