// TEST INFRASTRUCTURE ONLY.  A thread-by-thread CPU emulation of what the fused-pass CUDA kernel
// (cuda_quantum_simulator_b200/csrc/kernels_pass.cu) does with a compiled Program: tiles, sweeps,
// register slots, lane shuffles, control masks.  It exists so that the circuit compiler
// (csrc/program.cpp — host code, linked here unchanged) can be validated against the oracle in the
// GPU-less container.  The product never loads this library.
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../cuda_quantum_simulator_b200/csrc/program.hpp"

using namespace qsim::b200;
using cplx = std::complex<double>;

namespace {

long g_tma_errors = 0;
long g_bank_conflicts = 0;

void run_pass(const Program& prog, const PassDesc& pd, uint64_t hi_bits, cplx* state) {
    const DevOp* ops = prog.ops.data() + pd.op_offset;
    const uint64_t n_tiles = 1ULL << (pd.n - pd.t);
    const uint32_t tile_amps = 1u << pd.t;
    std::vector<cplx> tile(tile_amps);
    std::vector<uint64_t> gidx(tile_amps);
    // the kernel reads tile tau and writes tile tau ^ xor_tau (pairs are ordered so every tile is read before
    // it is overwritten); emulate with a snapshot of the input
    std::vector<cplx> snapshot;
    const cplx* src_state = state;
    if (pd.xor_tau) { snapshot.assign(state, state + (1ULL << pd.n)); src_state = snapshot.data(); }
    for (uint64_t tau = 0; tau < n_tiles; ++tau) {
        uint64_t base = 0;
        for (int s = 0; s < pd.n_segments; ++s)
            base |= ((tau >> pd.seg[s].src_shift) & pd.seg[s].mask) << pd.seg[s].dst_shift;
        const uint64_t gbase = base | hi_bits;
        for (uint32_t l = 0; l < tile_amps; ++l) {
            uint64_t g = base;
            for (int i = 0; i < pd.t; ++i)
                if ((l >> i) & 1) g |= 1ULL << pd.tile_bits[i];
            gidx[l] = g;
            tile[l] = src_state[g];
        }
        // cross-check the tensor-map geometry: the kernel's TMA boxes must land exactly these elements
        {
            const int ib = pd.tma_instr_bits, box_bits = pd.t - ib;
            for (uint32_t q = 0; q < (1u << ib); ++q) {
                uint64_t g = base;
                for (int b = 0; b < ib; ++b)
                    if ((q >> b) & 1) g |= 1ULL << pd.tile_bits[box_bits + b];
                uint64_t coord[5];
                for (int d = 0; d < 5; ++d)
                    coord[d] = pd.tma_dim[d].range_bits ? (g >> pd.tma_dim[d].start_bit) & ((1ULL << pd.tma_dim[d].range_bits) - 1) : 0;
                for (uint32_t e = 0; e < (1u << box_bits); ++e) {
                    uint64_t addr = 0;
                    uint32_t rem = e;
                    for (int d = 0; d < 5; ++d) {
                        const uint32_t in_box = rem & ((1u << pd.tma_dim[d].box_bits) - 1);
                        rem >>= pd.tma_dim[d].box_bits;
                        if (coord[d] & ((1ULL << pd.tma_dim[d].box_bits) - 1)) { g_tma_errors++; }
                        addr += (coord[d] + in_box) << pd.tma_dim[d].start_bit;
                    }
                    if (rem != 0 || addr != gidx[(q << box_bits) + e]) g_tma_errors++;
                }
            }
        }
        // per-tile factors of the fused diagonal runs (what the kernel's tile prologue computes)
        std::vector<cplx> eu((size_t)pd.n_phase * 13, cplx(1, 0));
        for (int o = 0; o < pd.n_ops; ++o) {
            const DevOp& op = ops[o];
            if (op.kind != OP_PHASE) continue;
            uint16_t starts[14];
            std::memcpy(starts, op.m, sizeof(starts));
            const PhaseTerm* terms = prog.phase_terms.data() + pd.phase_term_offset + op.cval_out;
            for (int e = 0; e < 13; ++e)
                for (int k = starts[e]; k < starts[e + 1]; ++k) {
                    const PhaseTerm& t = terms[k];
                    bool on = (gbase >> t.o) & 1;
                    if (t.kind == 2) on = on && ((gbase >> t.j) & 1);
                    if (on) eu[op.tmask_out * 13 + e] *= cplx(t.fr, t.fi);
                }
        }
        if (pd.n_head > 0) {
            // folded leading flips: the element that belongs at tile-local index l is read from F^-1(l)
            std::vector<cplx> moved(tile_amps);
            uint32_t shift = pd.head_const;
            for (int f = 0; f < pd.n_head_dyn; ++f)
                if ((gbase & pd.head_dyn[f].cmask_out) == pd.head_dyn[f].cval_out) shift ^= pd.head_dyn[f].w;
            for (uint32_t l = 0; l < tile_amps; ++l) {
                uint32_t src = shift;
                for (int j = 0; j < pd.t; ++j) if ((l >> j) & 1) src ^= pd.head_lin[j];
                moved[l] = tile[src];
            }
            tile.swap(moved);
        }
        for (int sw = 0; sw < pd.n_sweeps; ++sw) {
            const SweepDesc& sd = pd.sweep[sw];
            if (sd.n_head > 0) {   // flips folded into this sweep's load
                std::vector<cplx> moved(tile_amps);
                for (uint32_t l = 0; l < tile_amps; ++l) {
                    uint32_t src = sd.head_const;
                    for (int j = 0; j < pd.t; ++j) if ((l >> j) & 1) src ^= sd.head_lin[j];
                    moved[l] = tile[src];
                }
                tile.swap(moved);
            }
            const int slots = 1 << sd.r;
            const uint32_t n_active = 1u << sd.nthr;
            for (uint32_t warp0 = 0; warp0 < n_active; warp0 += 32) {
                cplx reg[32][16];
                uint32_t base_local[32];
                bool active[32];
                for (int lane = 0; lane < 32; ++lane) {
                    uint32_t tid = warp0 + lane;
                    active[lane] = tid < n_active;
                    uint32_t bl = 0;
                    for (int i = 0; i < sd.nthr; ++i)
                        if ((tid >> i) & 1) bl |= 1u << sd.thr_pos[i];
                    base_local[lane] = bl;
                    for (int k = 0; k < 16; ++k)
                        reg[lane][k] = (active[lane] && k < slots) ? tile[bl + sd.slot_off[k]] : cplx(0, 0);
                }
                // shared-memory rows: the 8 lanes of a quarter warp must touch 8 different 16-byte columns
                if (sd.nthr >= 5 && pd.t >= 3)
                    for (int k = 0; k < slots; ++k)
                        for (int q = 0; q < 32; q += 8) {
                            unsigned cols = 0;
                            for (int lane = q; lane < q + 8; ++lane)
                                if (active[lane]) cols |= 1u << ((base_local[lane] + sd.slot_off[k]) & 7u);
                            if (active[q + 7] && cols != 0xffu) g_bank_conflicts++;
                        }
                for (int o = sd.op_begin; o < sd.op_end; ++o) {
                    const DevOp& op = ops[o];
                    if (op.kind != OP_PHASE && (gbase & op.cmask_out) != op.cval_out) continue;   // (PHASE reuses these fields)
                    const cplx m00(op.m[0], op.m[1]), m01(op.m[2], op.m[3]), m10(op.m[4], op.m[5]), m11(op.m[6], op.m[7]);
                    cplx nxt[32][16];
                    for (int lane = 0; lane < 32; ++lane) {
                        const uint32_t tid = warp0 + lane;
                        const bool thr_ok = (tid & op.cmask_thr) == op.cval_thr;
                        const uint32_t sm = thr_ok ? op.slotmask : 0;
                        for (int k = 0; k < 16; ++k) {
                            cplx own = reg[lane][k], res = own;
                            if (op.kind == OP_PHASE) {
                                const uint32_t l = base_local[lane] + sd.slot_off[k < slots ? k : 0];
                                const double* tb = prog.phase_tables.data() + 2 * ((size_t)pd.phase_table_offset + op.cmask_out + l);
                                cplx f = cplx(tb[0], tb[1]) * eu[op.tmask_out * 13 + 12];
                                for (int j = 0; j < pd.t; ++j)
                                    if ((l >> j) & 1) f *= eu[op.tmask_out * 13 + j];
                                res = f * own;
                            } else if ((sm >> k) & 1) {
                                if (op.kind == OP_DIAG) {
                                    bool b;
                                    if (op.thome == T_REG) b = (op.tslots >> k) & 1;
                                    else if (op.thome == T_THREAD) b = (tid & op.tmask_thr) != 0;
                                    else b = (gbase & op.tmask_out) != 0;
                                    res = (b ? m11 : m00) * own;
                                } else {
                                    bool b;
                                    cplx partner;
                                    if (op.thome == T_REG) { b = (k >> op.tbit) & 1; partner = reg[lane][k ^ (1 << op.tbit)]; }
                                    else { b = (tid >> op.tbit) & 1; partner = reg[lane ^ (1 << op.tbit)][k]; }
                                    cplx c_own = b ? m11 : m00, c_par = b ? m10 : m01;
                                    if (op.kind == OP_FLIP) res = partner;
                                    else if (op.kind == OP_ADIAG) res = c_par * partner;
                                    else res = c_own * own + c_par * partner;
                                }
                            }
                            nxt[lane][k] = res;
                        }
                    }
                    std::memcpy(reg, nxt, sizeof(reg));
                }
                for (int lane = 0; lane < 32; ++lane)
                    if (active[lane])
                        for (int k = 0; k < slots; ++k) tile[base_local[lane] + sd.slot_off[k]] = reg[lane][k];
            }
        }
        {
            uint64_t sbase = 0;
            const uint64_t stau = tau ^ pd.xor_tau;
            for (int s = 0; s < pd.n_segments; ++s)
                sbase |= ((stau >> pd.seg[s].src_shift) & pd.seg[s].mask) << pd.seg[s].dst_shift;
            // the final store's index map: the folded flips (affine map, per-tile translations), then the X frame
            std::vector<cplx> permuted(tile_amps);
            uint32_t shift = pd.tail_const ^ pd.xor_local;
            for (int f = 0; f < pd.n_dyn; ++f)
                if ((gbase & pd.dyn[f].cmask_out) == pd.dyn[f].cval_out) shift ^= pd.dyn[f].w;
            for (uint32_t l = 0; l < tile_amps; ++l) {
                uint32_t d = 0;
                for (int j = 0; j < pd.t; ++j) if ((l >> j) & 1) d ^= pd.tail_lin[j];
                permuted[d ^ shift] = tile[l];
            }
            for (uint32_t l = 0; l < tile_amps; ++l) state[(gidx[l] - base) + sbase] = permuted[l];
        }
    }
}

}  // namespace

extern "C" int emu_run_ex(int n, int n_global, int rank, const qsim_gate_t* gates, int64_t ng, double* state,
                          int min_low_bits, int max_tile_bits, int merge, int reorder, uint64_t initial_xor,
                          int64_t* info_out, char* err, int errcap);

// Compile `gates` and run the emulated kernel on `state` (2^(n - n_global) amplitudes of shard `rank`).
// info_out: [0]=passes [1]=ops [2]=sweeps.  Returns 0, or -1 with the compiler's message in err.
extern "C" __attribute__((visibility("default")))
int emu_run(int n, int n_global, int rank, const qsim_gate_t* gates, int64_t ng, double* state,
            int min_low_bits, int max_tile_bits, int merge, int reorder, int64_t* info_out, char* err, int errcap) {
    return emu_run_ex(n, n_global, rank, gates, ng, state, min_low_bits, max_tile_bits, merge, reorder, 0, info_out, err,
                      errcap);
}

// Same with an inherited X frame (see CompileOptions::initial_xor); info_out[3] = frame left on the global qubits.
extern "C" __attribute__((visibility("default")))
int emu_run_ex(int n, int n_global, int rank, const qsim_gate_t* gates, int64_t ng, double* state, int min_low_bits,
               int max_tile_bits, int merge, int reorder, uint64_t initial_xor, int64_t* info_out, char* err, int errcap) {
    CompileOptions opt;
    opt.n_global = n_global;
    opt.initial_xor = initial_xor;
    if (min_low_bits > 0) opt.min_low_bits = min_low_bits;
    if (max_tile_bits > 0) opt.max_tile_bits = max_tile_bits;
    opt.merge = merge != 0;
    opt.reorder = reorder != 0;
    Program prog;
    std::string e;
    if (!compile(n, gates, ng, opt, prog, &e)) {
        if (err && errcap > 0) { std::strncpy(err, e.c_str(), errcap - 1); err[errcap - 1] = 0; }
        return -1;
    }
    const uint64_t hi = (uint64_t)rank << prog.n_local;
    g_tma_errors = 0;
    g_bank_conflicts = 0;
    for (const PassDesc& pd : prog.passes) run_pass(prog, pd, hi, reinterpret_cast<cplx*>(state));
    if (g_bank_conflicts) {
        if (err && errcap > 0) std::snprintf(err, errcap, "shared-memory bank conflicts in a sweep's rows (%ld)", g_bank_conflicts);
        return -3;
    }
    if (g_tma_errors) {
        if (err && errcap > 0) std::snprintf(err, errcap, "tensor-map geometry mismatch (%ld)", g_tma_errors);
        return -2;
    }
    if (info_out) {
        info_out[0] = (int64_t)prog.passes.size();
        info_out[1] = (int64_t)prog.lops.size();
        int64_t sw = 0;
        for (auto& p : prog.passes) sw += p.n_sweeps;
        info_out[2] = sw;
        info_out[3] = (int64_t)prog.global_xor;
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
int emu_describe(int n, int n_global, const qsim_gate_t* gates, int64_t ng, int min_low_bits, char* buf, int cap) {
    CompileOptions opt;
    opt.n_global = n_global;
    if (min_low_bits > 0) opt.min_low_bits = min_low_bits;
    Program prog;
    std::string e;
    if (!compile(n, gates, ng, opt, prog, &e)) { std::strncpy(buf, e.c_str(), cap - 1); buf[cap - 1] = 0; return -1; }
    std::string d = prog.describe();
    std::strncpy(buf, d.c_str(), cap - 1);
    buf[cap - 1] = 0;
    return 0;
}
