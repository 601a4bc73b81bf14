// Run-time specialised pass kernels: code generation, NVRTC, cache, launch.  See jit.hpp.
#include "jit.hpp"

#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace qsim {
namespace b200 {

namespace {

// the three sources the specialised kernel shares with the ahead-of-time one, embedded at build time (Makefile: *.embed)
const char kSrcPassDesc[] =
#include "_build/pass_desc.h.embed"
    ;
const char kSrcPassDevice[] =
#include "_build/pass_device.cuh.embed"
    ;
const char kSrcKernelBody[] =
#include "_build/pass_kernel_body.inc.embed"
    ;

// ---- code generation ------------------------------------------------------------------------------------------------

struct Gen {
    std::ostringstream os;
    int indent = 1;
    void line(const std::string& s) {
        for (int i = 0; i < indent; ++i) os << "    ";
        os << s << "\n";
    }
    void open(const std::string& s) { line(s + " {"); ++indent; }
    void close() { --indent; line("}"); }
};

std::string hex(uint64_t v) {
    char b[32];
    std::snprintf(b, sizeof(b), "0x%llxULL", (unsigned long long)v);
    return b;
}
std::string hex32(uint32_t v) {
    char b[32];
    std::snprintf(b, sizeof(b), "0x%xu", v);
    return b;
}
std::string num(long long v) { return std::to_string(v); }

// names of the register variables that hold slot k's amplitude
struct Slots {
    std::vector<int> var;   // slot -> variable index (bit flips on register bits are renamings)
    std::string pfx;        // (two-group kernels: the first half's variables outlive its block)
    std::string r(int k) const { return pfx + "xr" + num(var[k]); }
    std::string i(int k) const { return pfx + "xi" + num(var[k]); }
};

bool is_one(const double* m) { return m[0] == 1.0 && m[1] == 0.0; }

// value classes that change the generated code (they are part of the kernel's identity, the values themselves are not)
struct DiagClass {
    bool d0_one, real;
};
DiagClass diag_class(const DevOp& op) {
    DiagClass c;
    c.d0_one = is_one(&op.m[0]);
    c.real = op.m[1] == 0.0 && op.m[7] == 0.0;
    return c;
}

void emit_cmul(Gen& g, const Slots& s, int k, const std::string& pr, const std::string& pi, bool real) {
    if (real) {
        g.line(s.r(k) + " *= " + pr + "; " + s.i(k) + " *= " + pr + ";");
    } else {
        g.line("{ const double t_ = " + s.r(k) + " * " + pr + " - " + s.i(k) + " * " + pi + "; " + s.i(k) + " = " + s.r(k) +
               " * " + pi + " + " + s.i(k) + " * " + pr + "; " + s.r(k) + " = t_; }");
    }
}

// Conditional bit flips on a register bit (X / CNOT whose controls sit in thread bits or outside the tile) are not executed
// as 32 predicated register moves: the thread carries them as an XOR mask over its slot index (`fx_`), applied by the sweep's
// store addressing (a register bit only permutes the thread's OWN slots, so no barrier is involved).  Ops in between that
// do not look at the flipped bit are unaffected; a 2x2 on that very bit takes the X-conjugated matrix, a diagonal on it the
// swapped pair of factors (selects); anything else (controls on register bits, fused diagonal runs) first materialises the
// pending flips the old way.
struct Defer {
    bool enabled = false;
    bool dirty[4] = {false, false, false, false};
    uint32_t dep_thr[4] = {0, 0, 0, 0};   // thread bits the pending flips of that register bit were conditioned on
    bool any() const { return dirty[0] || dirty[1] || dirty[2] || dirty[3]; }
};

bool jit_defer_flips() {
    static const bool on = [] {
        const char* e = std::getenv("QSIM_JIT_DEFER_FLIPS");
        return !(e && (std::string(e) == "0" || std::string(e) == "off"));
    }();
    return on;
}

bool flip_deferrable(const DevOp& op, int n_slots) {
    const uint32_t all = (n_slots >= 32) ? 0xffffffffu : ((1u << n_slots) - 1u);
    return op.kind == OP_FLIP && op.thome == T_REG && (op.cmask_out != 0 || op.cmask_thr != 0) && (op.slotmask & all) == all;
}

void emit_materialise(Gen& g, Slots& s, Defer& df, int n_slots, int only_bit = -1) {
    for (int b = 0; b < 4; ++b) {
        if (!df.dirty[b] || (only_bit >= 0 && b != only_bit)) continue;
        const int J = 1 << b;
        g.open("if ((fx_ >> " + num(b) + ") & 1u)");
        for (int k = 0; k < n_slots; ++k) {
            if (k & J) continue;
            const std::string ar = s.r(k), ai = s.i(k), br = s.r(k | J), bi = s.i(k | J);
            g.line("{ const double ar_ = " + ar + ", ai_ = " + ai + "; " + ar + " = " + br + "; " + ai + " = " + bi + "; " + br + " = ar_; " + bi + " = ai_; }");
        }
        g.line("fx_ &= ~" + hex32((uint32_t)J) + ";");
        g.close();
        df.dirty[b] = false;
        df.dep_thr[b] = 0;
    }
}

void emit_op(Gen& g, Slots& s, const PassDesc& pd, const SweepDesc& sd, const DevOp& op, int o, int n_slots, Defer* df = nullptr) {
    static const char* kn[] = {"MAT", "MATREAL", "ADIAG", "FLIP", "DIAG", "PHASE"};
    bool dirty_target = false;   // this op's register target bit carries a pending conditional flip
    if (df && df->enabled) {
        const uint32_t all_ = (n_slots >= 32) ? 0xffffffffu : ((1u << n_slots) - 1u);
        if (flip_deferrable(op, n_slots)) {
            g.line("// op " + num(o) + ": FLIP target=reg:" + num(op.tbit) + " (conditional: carried as a slot-index XOR to the store)");
            std::string cond;
            if (op.cmask_out) cond = "(gbase & " + hex(op.cmask_out) + ") == " + hex(op.cval_out);
            if (op.cmask_thr) cond += std::string(cond.empty() ? "" : " && ") + "(tid & " + hex32(op.cmask_thr) + ") == " + hex32(op.cval_thr);
            g.line("if (" + cond + ") fx_ ^= " + hex32(1u << op.tbit) + ";");
            df->dirty[op.tbit] = true;
            df->dep_thr[op.tbit] |= op.cmask_thr;
            return;
        }
        if (df->any()) {
            const bool full = (op.slotmask & all_) == all_;
            if (op.kind == OP_PHASE || !full) emit_materialise(g, s, *df, n_slots);
            else if (op.thome == T_LANE) {
                // the two lanes of a pair must agree on where their slots are: a pending flip conditioned on this very lane
                // bit differs between them
                for (int b = 0; b < 4; ++b)
                    if (df->dirty[b] && ((df->dep_thr[b] >> op.tbit) & 1u)) emit_materialise(g, s, *df, n_slots, b);
            } else if (op.thome == T_REG && op.kind != OP_FLIP && df->dirty[op.tbit]) dirty_target = true;
        }
    }
    static const char* hn[] = {"lane", "reg", "thread", "outside"};
    const std::string M = "reinterpret_cast<const double2*>(sops[" + num(o) + "].m)";
    const uint32_t all = (n_slots >= 32) ? 0xffffffffu : ((1u << n_slots) - 1u);

    if (op.kind == OP_PHASE) {
        // amplitude(l) *= TABLE[l] * U * prod_{tile bits j of l} E_j   (see phase_op of the interpreter)
        const uint32_t present = op.cmask_thr;
        g.line("// op " + num(o) + ": PHASE present=" + hex32(present));
        g.open("");
        g.line("const double2* e_ = eu + " + num((long long)op.tmask_out * 13) + ";");
        g.line("double pr_ = 1.0, pi_ = 0.0;");
        if (present & (1u << 12)) g.line("pr_ = e_[12].x; pi_ = e_[12].y;");
        for (int b = 0; b < sd.nthr; ++b) {
            const int pos = sd.thr_pos[b];
            if (!((present >> pos) & 1u)) continue;
            g.line("if ((tid >> " + num(b) + ") & 1u) { const double2 v_ = e_[" + num(pos) + "]; const double t_ = pr_ * v_.x - pi_ * v_.y; "
                   "pi_ = pr_ * v_.y + pi_ * v_.x; pr_ = t_; }");
        }
        if (!(present & (1u << 13)) && (present & (1u << 14)))
            g.line("{ const double2 t2_ = __ldg(tables + " + num((long long)op.cmask_out) + "); const double t_ = t2_.x * pr_ - t2_.y * pi_; "
                   "pi_ = t2_.x * pi_ + t2_.y * pr_; pr_ = t_; }");
        // the table look-up goes to the COMPACT copy (indexed by the tile bits the run really depends on, op.tmask_thr):
        // the thread's part of the index is gathered from base_local with literal shifts, the slot's part is a literal
        const uint32_t dep = op.tmask_thr;
        auto compact = [&](uint32_t l) {   // pext(l, dep)
            uint32_t v = 0;
            int c = 0;
            for (int j = 0; j < kMaxTileBits; ++j)
                if ((dep >> j) & 1u) { v |= ((l >> j) & 1u) << c; ++c; }
            return v;
        };
        if (present & (1u << 13)) {
            std::string expr;
            int c = 0;
            for (int j = 0; j < kMaxTileBits;) {
                if (!((dep >> j) & 1u)) { ++j; continue; }
                int len = 0;
                while (j + len < kMaxTileBits && ((dep >> (j + len)) & 1u)) ++len;
                const std::string term = "(((base_local >> " + num(j) + ") & " + hex32((1u << len) - 1u) + ") << " + num(c) + ")";
                expr += (expr.empty() ? "" : " | ") + term;
                c += len;
                j += len;
            }
            g.line("const double2* tb_ = tables + " + num((long long)op.cval_thr) + " + (" + (expr.empty() ? std::string("0u") : expr) + ");");
        }
        for (int k = 0; k < n_slots; ++k) {
            std::string fr = "fr" + num(k) + "_", fi = "fi" + num(k) + "_";
            if (present & (1u << 13)) {
                g.line("double " + fr + ", " + fi + "; { const double2 t2_ = __ldg(tb_ + " + num((long long)compact(sd.slot_off[k])) +
                       "); " + fr + " = t2_.x * pr_ - t2_.y * pi_; " + fi + " = t2_.x * pi_ + t2_.y * pr_; }");
            } else {
                g.line("double " + fr + " = pr_, " + fi + " = pi_;");
            }
        }
        for (int j = 0; j < sd.r; ++j) {
            if (!((present >> sd.reg_pos[j]) & 1u)) continue;
            g.line("{ const double2 v_ = e_[" + num(sd.reg_pos[j]) + "];");
            for (int k = 0; k < n_slots; ++k) {
                if (!((k >> j) & 1)) continue;
                std::string fr = "fr" + num(k) + "_", fi = "fi" + num(k) + "_";
                g.line("  { const double t_ = " + fr + " * v_.x - " + fi + " * v_.y; " + fi + " = " + fr + " * v_.y + " + fi +
                       " * v_.x; " + fr + " = t_; }");
            }
            g.line("}");
        }
        for (int k = 0; k < n_slots; ++k) emit_cmul(g, s, k, "fr" + num(k) + "_", "fi" + num(k) + "_", false);
        g.close();
        return;
    }

    const bool has_out_ctrl = op.cmask_out != 0;
    const bool has_thr_ctrl = op.cmask_thr != 0;
    const uint32_t slotset = op.slotmask & all;
    g.line("// op " + num(o) + ": " + kn[op.kind] + " target=" + hn[op.thome] + ":" + num(op.tbit) + " slots=" + hex32(slotset) +
           (has_thr_ctrl ? " thr-ctrl" : "") + (has_out_ctrl ? " out-ctrl" : ""));
    if (slotset == 0) return;

    // FLIP on a register bit without thread / outside controls: a renaming of the slot variables, no code at all
    if (op.kind == OP_FLIP && op.thome == T_REG && !has_out_ctrl && !has_thr_ctrl) {
        const int J = 1 << op.tbit;
        for (int k = 0; k < n_slots; ++k)
            if (!(k & J) && ((slotset >> k) & 1)) std::swap(s.var[k], s.var[k | J]);
        return;
    }

    int opened = 0;
    if (has_out_ctrl) { g.open("if ((gbase & " + hex(op.cmask_out) + ") == " + hex(op.cval_out) + ")"); ++opened; }
    else { g.open(""); ++opened; }
    // Controls held in tid bits: a branch around in-thread work (warp-uniform for warp bits; for lane bits the idle lanes
    // cost nothing extra).  Lane-target ops keep every lane in the shuffles (a partial-mask shuffle compiles to a
    // WARPSYNC / collective sequence per instruction) and select afterwards, as the interpreter does.
    const bool lane_sel = has_thr_ctrl && op.thome == T_LANE;
    if (has_thr_ctrl) {
        g.line("const bool pt_ = (tid & " + hex32(op.cmask_thr) + ") == " + hex32(op.cval_thr) + ";");
        if (!lane_sel) { g.open("if (pt_)"); ++opened; }
    }

    if (op.kind == OP_DIAG) {
        const DiagClass dc = diag_class(op);
        g.line("const double2 d0_ = " + M + "[0], d1_ = " + M + "[3];");
        if (op.thome == T_REG && dirty_target) {
            // the target bit carries a pending flip: the two factors trade places when it is set
            g.line("const bool sw_ = (fx_ >> " + num(op.tbit) + ") & 1u;");
            g.line("const double e0r_ = sw_ ? d1_.x : d0_.x, e0i_ = sw_ ? d1_.y : d0_.y, e1r_ = sw_ ? d0_.x : d1_.x, e1i_ = sw_ ? d0_.y : d1_.y;");
            for (int k = 0; k < n_slots; ++k) {
                if (!((slotset >> k) & 1)) continue;
                const bool b = (op.tslots >> k) & 1;
                emit_cmul(g, s, k, b ? "e1r_" : "e0r_", b ? "e1i_" : "e0i_", dc.real);
            }
        } else if (op.thome == T_REG) {
            for (int k = 0; k < n_slots; ++k) {
                if (!((slotset >> k) & 1)) continue;
                const bool b = (op.tslots >> k) & 1;
                if (!b && dc.d0_one) continue;
                emit_cmul(g, s, k, b ? "d1_.x" : "d0_.x", b ? "d1_.y" : "d0_.y", dc.real);
            }
        } else {
            const std::string bt = (op.thome == T_THREAD) ? "(tid & " + hex32(op.tmask_thr) + ") != 0u"
                                                          : "(gbase & " + hex(op.tmask_out) + ") != 0ULL";
            if (dc.d0_one) {
                g.open("if (" + bt + ")");
                for (int k = 0; k < n_slots; ++k)
                    if ((slotset >> k) & 1) emit_cmul(g, s, k, "d1_.x", "d1_.y", dc.real);
                g.close();
            } else {
                g.line("const bool bt_ = " + bt + ";");
                g.line("const double pr_ = bt_ ? d1_.x : d0_.x, pi_ = bt_ ? d1_.y : d0_.y;");
                for (int k = 0; k < n_slots; ++k)
                    if ((slotset >> k) & 1) emit_cmul(g, s, k, "pr_", "pi_", dc.real);
            }
        }
    } else if (op.thome == T_REG) {
        const int J = 1 << op.tbit;
        if (op.kind != OP_FLIP && dirty_target) {
            // the target bit carries a pending flip: X-conjugated matrix [[d, c], [b, a]] when it is set
            g.line("const bool sw_ = (fx_ >> " + num(op.tbit) + ") & 1u;");
            g.line("const double2 m0_ = " + M + "[0], m1_ = " + M + "[1], m2_ = " + M + "[2], m3_ = " + M + "[3];");
            g.line("const double2 ma_ = sw_ ? m3_ : m0_, mb_ = sw_ ? m2_ : m1_, mc_ = sw_ ? m1_ : m2_, md_ = sw_ ? m0_ : m3_;");
        } else if (op.kind != OP_FLIP) g.line("const double2 ma_ = " + M + "[0], mb_ = " + M + "[1], mc_ = " + M + "[2], md_ = " + M + "[3];");
        for (int k = 0; k < n_slots; ++k) {
            if ((k & J) || !((slotset >> k) & 1)) continue;
            const int k1 = k | J;
            const std::string ar = s.r(k), ai = s.i(k), br = s.r(k1), bi = s.i(k1);
            g.line("{ const double ar_ = " + ar + ", ai_ = " + ai + ", br_ = " + br + ", bi_ = " + bi + ";");
            if (op.kind == OP_FLIP) {
                g.line("  " + ar + " = br_; " + ai + " = bi_; " + br + " = ar_; " + bi + " = ai_; }");
            } else if (op.kind == OP_ADIAG) {
                g.line("  " + ar + " = mb_.x * br_ - mb_.y * bi_; " + ai + " = mb_.x * bi_ + mb_.y * br_;");
                g.line("  " + br + " = mc_.x * ar_ - mc_.y * ai_; " + bi + " = mc_.x * ai_ + mc_.y * ar_; }");
            } else if (op.kind == OP_MATREAL) {
                g.line("  " + ar + " = ma_.x * ar_ + mb_.x * br_; " + ai + " = ma_.x * ai_ + mb_.x * bi_;");
                g.line("  " + br + " = mc_.x * ar_ + md_.x * br_; " + bi + " = mc_.x * ai_ + md_.x * bi_; }");
            } else {
                g.line("  " + ar + " = ma_.x * ar_ - ma_.y * ai_ + mb_.x * br_ - mb_.y * bi_;");
                g.line("  " + ai + " = ma_.x * ai_ + ma_.y * ar_ + mb_.x * bi_ + mb_.y * br_;");
                g.line("  " + br + " = mc_.x * ar_ - mc_.y * ai_ + md_.x * br_ - md_.y * bi_;");
                g.line("  " + bi + " = mc_.x * ai_ + mc_.y * ar_ + md_.x * bi_ + md_.y * br_; }");
            }
        }
    } else {   // T_LANE
        const std::string lm = num(1 << op.tbit);
        if (op.kind != OP_FLIP) {
            g.line("const bool hb_ = (tid >> " + num(op.tbit) + ") & 1u;");
            // coefficient of my own amplitude and of my partner's
            g.line("const double2 ma_ = " + M + "[0], mb_ = " + M + "[1], mc_ = " + M + "[2], md_ = " + M + "[3];");
            g.line("const double cor_ = hb_ ? md_.x : ma_.x, coi_ = hb_ ? md_.y : ma_.y, cpr_ = hb_ ? mc_.x : mb_.x, cpi_ = hb_ ? mc_.y : mb_.y;");
        }
        if (op.kind == OP_FLIP && lane_sel) g.line("const int src_ = pt_ ? (int)((tid & 31u) ^ " + lm + "u) : (int)(tid & 31u);");
        for (int k = 0; k < n_slots; ++k) {
            if (!((slotset >> k) & 1)) continue;
            const std::string xr = s.r(k), xi = s.i(k);
            if (op.kind == OP_FLIP) {
                // pure data movement: the select is on the source lane, not on the data
                if (lane_sel) g.line(xr + " = __shfl_sync(0xffffffffu, " + xr + ", src_); " + xi + " = __shfl_sync(0xffffffffu, " + xi + ", src_);");
                else g.line(xr + " = __shfl_xor_sync(0xffffffffu, " + xr + ", " + lm + "); " + xi + " = __shfl_xor_sync(0xffffffffu, " + xi + ", " + lm + ");");
                continue;
            }
            g.line("{ const double pr_ = __shfl_xor_sync(0xffffffffu, " + xr + ", " + lm + "), pi_ = __shfl_xor_sync(0xffffffffu, " + xi + ", " + lm + ");");
            if (op.kind == OP_ADIAG)
                g.line("  const double nr_ = cpr_ * pr_ - cpi_ * pi_, ni_ = cpr_ * pi_ + cpi_ * pr_;");
            else if (op.kind == OP_MATREAL)
                g.line("  const double nr_ = cor_ * " + xr + " + cpr_ * pr_, ni_ = cor_ * " + xi + " + cpr_ * pi_;");
            else
                g.line("  const double nr_ = cor_ * " + xr + " - coi_ * " + xi + " + cpr_ * pr_ - cpi_ * pi_, ni_ = cor_ * " + xi + " + coi_ * " + xr +
                       " + cpr_ * pi_ + cpi_ * pr_;");
            if (lane_sel) g.line("  " + xr + " = pt_ ? nr_ : " + xr + "; " + xi + " = pt_ ? ni_ : " + xi + "; }");
            else g.line("  " + xr + " = nr_; " + xi + " = ni_; }");
        }
    }
    while (opened-- > 0) g.close();
}

}  // namespace

// Can this pass run as TWO WARP GROUPS (pass_kernel_body.inc, QSIM_DUAL_GROUPS)?  Full tiles swept by all 512 (virtual)
// threads only.
bool jit_dual_possible(const PassDesc& pd) {
    if (pd.t != kMaxTileBits || kMaxRegBits != 3 || pd.n_sweeps < 1) return false;
    for (int sw = 0; sw < pd.n_sweeps; ++sw)
        if (pd.sweep[sw].r != 3 || pd.sweep[sw].nthr != 9) return false;
    return true;
}

// Rough count of FP64 instructions per tile and thread (8 amplitudes), to tell compute-heavy passes from HBM-bound ones.
int jit_fp64_estimate(const PassDesc& pd, const DevOp* ops) {
    int total = 0;
    for (int o = 0; o < pd.n_ops; ++o) {
        const DevOp& op = ops[o];
        int w = 0;
        switch (op.kind) {
            case OP_MAT: w = 64; break;
            case OP_MATREAL: w = 32; break;
            case OP_ADIAG: w = 32; break;
            case OP_DIAG: w = 24; break;
            case OP_PHASE: w = 80; break;
            default: w = 0;
        }
        if (op.kind != OP_PHASE && op.kind != OP_DIAG && op.thome == T_REG && op.slotmask != 0 &&
            (op.slotmask & 0xff) != 0xff) w /= 2;   // a control on a register bit: half the slots
        total += w;
    }
    return total;
}

std::string jit_generate_compute(const PassDesc& pd, const DevOp* ops, bool dual) {
    Gen g;
    g.indent = 0;
    g.line("// generated by qsim_b200 jit.cpp: per-tile compute of one pass (n=" + num(pd.n) + ", t=" + num(pd.t) + ", " +
           num(pd.n_sweeps) + " sweep(s), " + num(pd.n_ops) + " op(s), ~" + num(jit_fp64_estimate(pd, ops)) + " FP64 instructions per thread)" +
           (dual ? ", two warp groups" : ""));
    // the pass's tensor-map geometry as literals: box coordinates of a tile and the offsets of its TMA instructions (the
    // elected warp computes these for every load and store, on the tile's critical path)
    g.line("#define QSIM_JIT_GEOMETRY 1");
    g.line("__device__ __forceinline__ void jit_tma_coords(uint64_t g, int (&c)[5]) {");
    for (int d = 0; d < 5; ++d) {
        const TmaDim& td = pd.tma_dim[d];
        if (td.range_bits == 0) { g.line("    c[" + num(d) + "] = 0;"); continue; }
        const std::string v = "((g >> " + num(td.start_bit) + ") & " + hex((td.range_bits >= 64 ? ~0ULL : ((1ULL << td.range_bits) - 1ULL))) + ")";
        g.line("    c[" + num(d) + "] = (int)(" + v + (d == 0 ? " * 2ULL" : "") + ");");
    }
    g.line("}");
    g.line("__device__ __forceinline__ uint64_t jit_instr_offset(uint32_t q) {");
    {
        std::string e = "0ULL";
        const int box_bits = pd.t - pd.tma_instr_bits;
        for (int b = 0; b < pd.tma_instr_bits; ++b)
            e += " | ((uint64_t)((q >> " + num(b) + ") & 1u) << " + num(pd.tile_bits[box_bits + b]) + ")";
        g.line("    return " + e + ";");
    }
    g.line("}");
    if (dual) {
        g.line("#define QSIM_DUAL_GROUPS 1");
        g.line("__device__ __forceinline__ void jit_compute_tile(const PassParams& P, unsigned char* tile, uint64_t gbase, uint32_t tid0,");
        g.line("                                                 const DevOp* sops, const double2* eu, const uint16_t* base_tab, uint32_t bar_id) {");
    } else {
        g.line("__device__ __forceinline__ void jit_compute_tile(const PassParams& P, unsigned char* tile, uint64_t gbase, uint32_t tid,");
        g.line("                                                 const DevOp* sops, const double2* eu, const uint16_t* base_tab) {");
    }
    g.indent = 1;
    g.line("const uint32_t tile_u32 = smem_u32(tile);");
    if (!dual) g.line("const uint32_t warp = tid >> 5;");
    g.line("const double2* tables = reinterpret_cast<const double2*>(P.phase_tables);");
    g.line(std::string(dual ? "" : "(void)warp; ") + "(void)tables; (void)eu; (void)sops; (void)gbase;");
    const std::string BAR = dual ? "asm volatile(\"bar.sync %0, 256;\" ::\"r\"(bar_id) : \"memory\");" : "__syncthreads();";
    const int T = kComputeThreads;
    for (int sw = 0; sw < pd.n_sweeps; ++sw) {
        const SweepDesc& sd = pd.sweep[sw];
        const int n_slots = 1 << sd.r;
        const uint32_t n_active = 1u << sd.nthr;
        const bool last = (sw + 1 == pd.n_sweeps);
        const uint32_t xl = last ? pd.xor_local : 0u;
        const int n_tail = last ? pd.n_tail : 0;
        const bool pass_head = (sw == 0 && pd.n_head > 0);
        const bool mapped_load = pass_head || sd.n_head > 0;
        const bool permuted_store = (xl != 0u) || (n_tail > 0) || mapped_load;
        const bool partial_warp = n_active < 32u;          // lanes beyond the tile inside the one active warp
        const bool some_warps_idle = n_active < (uint32_t)T;
        const bool plain_store = !last || (pd.n_tail == 0 && pd.n_dyn == 0 && xl == 0u);
        const std::string guard = partial_warp ? "if (active) " : "";
        // the three parts of a sweep for one (virtual) thread: addresses + loads, ops, stores; `px` prefixes the names that
        // must outlive the block they are set in (two-group kernels, first half of a sweep that permutes the tile)
        auto emit_decl = [&](Slots& s, const std::string& px) {
            s.var.resize(n_slots);
            s.pfx = px;
            for (int k = 0; k < n_slots; ++k) s.var[k] = k;
            std::string decl = "double ";
            for (int k = 0; k < n_slots; ++k) decl += (k ? ", " : "") + (px + "xr" + num(k) + " = 0.0, " + px + "xi" + num(k) + " = 0.0");
            g.line(decl + ";");
        };
        auto emit_load = [&](Slots& s) {
            if (partial_warp) g.line("const bool active = tid < " + num(n_active) + "u;");
            g.line("const uint32_t base_local = base_tab[" + num(sw * T) + " + (int)tid];");
            g.line("const uint32_t a0_ = tile_u32 + base_local * 16u;");
            if (mapped_load) {
                const uint16_t* loff;
                if (pass_head) {
                    g.line("uint32_t lb_ = base_tab[" + num((pd.n_sweeps + 1) * T) + " + (int)tid];");
                    for (int f = 0; f < pd.n_head_dyn; ++f)
                        g.line("if ((gbase & " + hex(pd.head_dyn[f].cmask_out) + ") == " + hex(pd.head_dyn[f].cval_out) + ") lb_ ^= " +
                               hex32(pd.head_dyn[f].w) + ";");
                    loff = pd.load_slot_off;
                } else {
                    g.line("uint32_t lb_ = " + hex32(sd.head_const) + ";");
                    for (int j = 0; j < pd.t; ++j)
                        if (sd.head_lin[j]) g.line("if ((base_local >> " + num(j) + ") & 1u) lb_ ^= " + hex32(sd.head_lin[j]) + ";");
                    loff = sd.load_slot_off;
                }
                for (int k = 0; k < n_slots; ++k)
                    g.line(guard + "lds128(tile_u32 + ((lb_ ^ " + hex32(loff[k]) + ") << 4), " + s.r(k) + ", " + s.i(k) + ");");
            } else {
                for (int k = 0; k < n_slots; ++k)
                    g.line(guard + "lds128(a0_ + " + num((long long)sd.slot_off[k] * 16) + "u, " + s.r(k) + ", " + s.i(k) + ");");
            }
        };
        // where the stores go: `dst` names a variable holding a0_ (plain) or sb_ (mapped)
        // does this sweep carry conditional register-bit flips to its store (Defer)?
        bool sweep_defers = false;
        for (int o = sd.op_begin; o < sd.op_end && jit_defer_flips(); ++o) sweep_defers = sweep_defers || flip_deferrable(ops[o], n_slots);
        const bool index_store = !plain_store || sweep_defers;   // stores address by tile-local index XOR instead of base + literal
        auto emit_store_base = [&](const std::string& dst, bool declare, const Defer& df) {
            const std::string lhs = (declare ? "const uint32_t " : "") + dst + " = ";
            if (!index_store) { g.line(lhs + "a0_;"); return; }
            if (plain_store) g.line("uint32_t sb_ = base_local;");
            else {
                g.line("uint32_t sb_ = (uint32_t)base_tab[" + num(pd.n_sweeps * T) + " + (int)tid] ^ " + hex32(xl) + ";");
                for (int f = 0; f < pd.n_dyn; ++f)
                    g.line("if ((gbase & " + hex(pd.dyn[f].cmask_out) + ") == " + hex(pd.dyn[f].cval_out) + ") sb_ ^= " + hex32(pd.dyn[f].w) + ";");
            }
            // pending conditional flips: slot k's amplitude goes where slot k ^ fx_ lives (the slot offsets are XOR-linear)
            for (int b = 0; b < sd.r; ++b)
                if (df.dirty[b])
                    g.line("if ((fx_ >> " + num(b) + ") & 1u) sb_ ^= " + hex32(plain_store ? sd.slot_off[1 << b] : pd.store_slot_off[1 << b]) + ";");
            g.line(lhs + "sb_;");
        };
        auto emit_store = [&](const Slots& s, const std::string& base) {
            for (int k = 0; k < n_slots; ++k) {
                if (!index_store) g.line(guard + "sts128(" + base + " + " + num((long long)sd.slot_off[k] * 16) + "u, " + s.r(k) + ", " + s.i(k) + ");");
                else g.line(guard + "sts128(tile_u32 + ((" + base + " ^ " + hex32(plain_store ? sd.slot_off[k] : pd.store_slot_off[k]) + ") << 4), " + s.r(k) + ", " + s.i(k) + ");");
            }
        };
        auto emit_ops = [&](Slots& s, Defer& df) {
            df.enabled = sweep_defers;
            if (sweep_defers) g.line("uint32_t fx_ = 0u;   // pending conditional flips of register bits (see Defer)");
            for (int o = sd.op_begin; o < sd.op_end; ++o) emit_op(g, s, pd, sd, ops[o], o, n_slots, &df);
        };
        g.line("// ---- sweep " + num(sw) + ": r=" + num(sd.r) + " nthr=" + num(sd.nthr));
        g.open("");
        if (sw > 0) g.line(BAR);
        if (!dual) {
            if (some_warps_idle) g.open("if ((warp << 5) < " + num(n_active) + "u)");
            else g.open("");
            Slots s;
            Defer df;
            emit_decl(s, "");
            emit_load(s);
            emit_ops(s, df);
            if (permuted_store) g.line(BAR);
            if (!index_store) emit_store(s, "a0_");
            else { emit_store_base("sbx_", true, df); emit_store(s, "sbx_"); }
            g.close();
            if (some_warps_idle && permuted_store) g.line("else { " + BAR + " }   // keep the barrier count equal across warps");
        } else if (!permuted_store) {
            // each thread does the work of virtual threads tid0 and tid0 + 256, one after the other
            for (int h = 0; h < 2; ++h) {
                g.open("");
                g.line("const uint32_t tid = tid0 + " + num(256 * h) + "u;");
                Slots s;
                Defer df;
                emit_decl(s, "");
                emit_load(s);
                emit_ops(s, df);
                if (!index_store) emit_store(s, "a0_");
                else { emit_store_base("sbx_", true, df); emit_store(s, "sbx_"); }
                g.close();
            }
        } else {
            // the sweep permutes the tile: both halves are loaded before anything is stored (the first half stays in registers)
            Slots s0, s1;
            emit_decl(s0, "h0");
            g.line("uint32_t h0sb_;");
            g.open("");
            g.line("const uint32_t tid = tid0;");
            Defer df0, df1;
            emit_load(s0);
            emit_ops(s0, df0);
            emit_store_base("h0sb_", false, df0);
            g.close();
            g.open("");
            g.line("const uint32_t tid = tid0 + 256u;");
            emit_decl(s1, "");
            emit_load(s1);
            emit_ops(s1, df1);
            g.line(BAR);
            if (!index_store) emit_store(s1, "a0_");
            else { emit_store_base("sbx_", true, df1); emit_store(s1, "sbx_"); }
            g.close();
            emit_store(s0, "h0sb_");
        }
        g.close();
    }
    g.indent = 0;
    g.line("}");
    return g.os.str();
}

std::string jit_translation_unit(const PassDesc& pd, const DevOp* ops, bool dual);   // defined below (needs tu_from_compute)

// ---- NVRTC (loaded on demand: the library itself does not link it) ----------------------------------------------------

namespace {

struct Nvrtc {
    void* handle = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetErrorString) error_string = nullptr;
    std::string why;   // why it is unavailable
    bool ok() const { return handle != nullptr; }
};

const Nvrtc& nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        std::vector<std::string> names;
        if (const char* e = std::getenv("QSIM_NVRTC_LIB")) names.push_back(e);
        // the toolkit's own NVRTC first (the build's CUDA version); a bare soname may resolve to an older copy that another
        // library of the process bundles (PyTorch ships 12.8)
        for (const char* s : {"/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12", "libnvrtc.so"})
            names.push_back(s);
        for (const std::string& nm : names) {
            r.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (r.handle) break;
        }
        if (!r.handle) { r.why = "libnvrtc.so.12 not found (set QSIM_NVRTC_LIB)"; return r; }
        auto sym = [&](const char* s) { return dlsym(r.handle, s); };
        r.create = reinterpret_cast<decltype(r.create)>(sym("nvrtcCreateProgram"));
        r.compile = reinterpret_cast<decltype(r.compile)>(sym("nvrtcCompileProgram"));
        r.cubin_size = reinterpret_cast<decltype(r.cubin_size)>(sym("nvrtcGetCUBINSize"));
        r.cubin = reinterpret_cast<decltype(r.cubin)>(sym("nvrtcGetCUBIN"));
        r.log_size = reinterpret_cast<decltype(r.log_size)>(sym("nvrtcGetProgramLogSize"));
        r.log = reinterpret_cast<decltype(r.log)>(sym("nvrtcGetProgramLog"));
        r.destroy = reinterpret_cast<decltype(r.destroy)>(sym("nvrtcDestroyProgram"));
        r.error_string = reinterpret_cast<decltype(r.error_string)>(sym("nvrtcGetErrorString"));
        if (!r.create || !r.compile || !r.cubin_size || !r.cubin || !r.log_size || !r.log || !r.destroy) {
            r.why = "libnvrtc lacks a required entry point";
            dlclose(r.handle);
            r.handle = nullptr;
        }
        return r;
    }();
    return n;
}

// Process-wide state.  It lives in heap objects that are never destroyed: background threads may still be compiling when the
// process exits, and a loaded kernel must not be unloaded while the CUDA runtime is being torn down.
std::mutex& g_mu = *new std::mutex();
JitMode g_mode = JitMode::Auto;
int g_min_qubits = 26;
bool g_mode_init = false;
JitStats& g_stats = *new JitStats();
std::string& g_last_log = *new std::string();
bool g_warned = false;

void init_mode_locked() {
    if (g_mode_init) return;
    g_mode_init = true;
    if (const char* e = std::getenv("QSIM_JIT")) {
        const std::string v = e;
        if (v == "off" || v == "0") g_mode = JitMode::Off;
        else if (v == "always" || v == "2") g_mode = JitMode::Always;
        else g_mode = JitMode::Auto;
    }
    if (const char* e = std::getenv("QSIM_JIT_MIN_QUBITS")) g_min_qubits = std::atoi(e);
}

// ---- on-disk cache of compiled kernels: $QSIM_JIT_CACHE (default ~/.cache/qsim_b200/jit; "off" disables) ----
// file = "QJ01" | u64 source length | generated source | cubin; the source is compared on load, so a hash collision or a
// stale file from another library version (the embedded skeleton is part of the key) can only cost a recompile.
std::string cache_dir() {
    static std::string dir = [] {
        std::string d;
        if (const char* e = std::getenv("QSIM_JIT_CACHE")) d = e;
        else if (const char* h = std::getenv("HOME")) d = std::string(h) + "/.cache/qsim_b200/jit";
        if (d == "off" || d == "0") d.clear();
        if (!d.empty()) {   // mkdir -p
            for (size_t i = 1; i <= d.size(); ++i)
                if (i == d.size() || d[i] == '/') ::mkdir(d.substr(0, i).c_str(), 0755);
        }
        return d;
    }();
    return dir;
}

std::string cache_path(uint64_t key) {
    char name[64];
    std::snprintf(name, sizeof(name), "/%016llx.qjit", (unsigned long long)key);
    return cache_dir() + name;
}

bool cache_load(uint64_t key, const std::string& source, std::vector<char>& cubin) {
    if (cache_dir().empty()) return false;
    FILE* f = std::fopen(cache_path(key).c_str(), "rb");
    if (!f) return false;
    bool ok = false;
    char magic[4];
    uint64_t len = 0;
    if (std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "QJ01", 4) == 0 && std::fread(&len, 8, 1, f) == 1 && len == source.size()) {
        std::string src(len, '\0');
        if (std::fread(&src[0], 1, len, f) == len && src == source) {
            const long pos = std::ftell(f);
            std::fseek(f, 0, SEEK_END);
            const long end = std::ftell(f);
            std::fseek(f, pos, SEEK_SET);
            if (end > pos) {
                cubin.resize((size_t)(end - pos));
                ok = std::fread(cubin.data(), 1, cubin.size(), f) == cubin.size();
            }
        }
    }
    std::fclose(f);
    return ok;
}

void cache_store(uint64_t key, const std::string& source, const std::vector<char>& cubin) {
    if (cache_dir().empty()) return;
    const std::string path = cache_path(key), tmp = path + ".tmp" + std::to_string((long)::getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return;
    const uint64_t len = source.size();
    const bool ok = std::fwrite("QJ01", 1, 4, f) == 4 && std::fwrite(&len, 8, 1, f) == 1 &&
                    std::fwrite(source.data(), 1, source.size(), f) == source.size() &&
                    std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    std::fclose(f);
    if (ok) std::rename(tmp.c_str(), path.c_str());
    else std::remove(tmp.c_str());
}

uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ULL;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
    return h;
}

}  // namespace

struct JitKernel {
    std::vector<char> cubin;
    cudaLibrary_t library = nullptr;
    cudaKernel_t kernel = nullptr;
    bool attr_set[64] = {};
    std::string source;
    ~JitKernel() {
        if (library) cudaLibraryUnload(library);
    }
};

namespace {
// keyed by the hash of the generated compute (never destroyed, see above)
std::unordered_map<uint64_t, std::shared_ptr<JitKernel>>& g_cache = *new std::unordered_map<uint64_t, std::shared_ptr<JitKernel>>();
std::unordered_map<uint64_t, bool>& g_failed = *new std::unordered_map<uint64_t, bool>();
}  // namespace

JitMode jit_mode() {
    std::lock_guard<std::mutex> lock(g_mu);
    init_mode_locked();
    return g_mode;
}
int jit_min_qubits() {
    std::lock_guard<std::mutex> lock(g_mu);
    init_mode_locked();
    return g_min_qubits;
}
void jit_set_mode(JitMode mode, int min_qubits) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_mode_init = true;
    g_mode = mode;
    if (min_qubits > 0) g_min_qubits = min_qubits;
}

bool jit_wanted(const PassDesc& pd) {
    const JitMode m = jit_mode();
    if (m == JitMode::Off) return false;
    if (m == JitMode::Always) return true;
    return pd.n >= jit_min_qubits();
}

JitStats jit_stats() {
    std::lock_guard<std::mutex> lock(g_mu);
    return g_stats;
}
std::string jit_last_log() {
    std::lock_guard<std::mutex> lock(g_mu);
    return g_last_log;
}

namespace {

std::string tu_from_compute(const std::string& compute) {
    std::string tu;
    tu.reserve(compute.size() + 1024);
    tu += "#define QSIM_REG_BITS " + std::to_string(QSIM_REG_BITS) + "\n";
    tu += "#include \"pass_desc.h\"\n#include \"pass_device.cuh\"\n";
    tu += "namespace qsim {\nnamespace b200 {\nnamespace {\n";
    tu += compute;
    tu += "}  // namespace\n";
    tu += "#define QSIM_PASS_KERNEL qsim_jit_pass\n#define QSIM_COMPUTE_TILE jit_compute_tile\n#define QSIM_KERNEL_LINKAGE extern \"C\"\n";
    tu += "#include \"pass_kernel_body.inc\"\n";
    tu += "}  // namespace b200\n}  // namespace qsim\n";
    return tu;
}

// NVRTC, no locks held: generated compute -> sm_100a cubin.  Returns an empty string on success, else what went wrong.
std::string compile_cubin(const std::string& compute, std::vector<char>& cubin, std::string& log) {
    const Nvrtc& rt = nvrtc();
    if (!rt.ok()) return rt.why;
    const std::string tu = tu_from_compute(compute);
    nvrtcProgram prog = nullptr;
    const char* headers[] = {kSrcPassDesc, kSrcPassDevice, kSrcKernelBody};
    const char* names[] = {"pass_desc.h", "pass_device.cuh", "pass_kernel_body.inc"};
    if (rt.create(&prog, tu.c_str(), "qsim_jit_pass.cu", 3, headers, names) != NVRTC_SUCCESS) return "nvrtcCreateProgram failed";
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    const nvrtcResult rc = rt.compile(prog, 4, opts);
    size_t log_n = 0;
    rt.log_size(prog, &log_n);
    log.assign(log_n, '\0');
    if (log_n > 1) rt.log(prog, &log[0]);
    if (rc != NVRTC_SUCCESS) {
        rt.destroy(&prog);
        if (std::getenv("QSIM_JIT_DUMP")) std::fprintf(stderr, "%s\n", tu.c_str());
        return std::string("compile failed: ") + (rt.error_string ? rt.error_string(rc) : "?") + "\n" + log.substr(0, 4000);
    }
    size_t cb = 0;
    rt.cubin_size(prog, &cb);
    cubin.resize(cb);
    rt.cubin(prog, cubin.data());
    rt.destroy(&prog);
    return "";
}

// (g_mu held) make the kernel launchable in this process: cubin -> library -> kernel handle
bool ensure_loaded_locked(JitKernel& k) {
    if (k.kernel) return true;
    cudaError_t e = cudaLibraryLoadData(&k.library, k.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k.kernel, k.library, "qsim_jit_pass");
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (k.library) { cudaLibraryUnload(k.library); k.library = nullptr; }
        k.kernel = nullptr;
        return false;
    }
    return true;
}

// ---- background compiles: run() never waits for NVRTC ----
// A few worker threads compile queued requests (CPU work only: the cubin is loaded into the CUDA context by the thread that
// launches, so the worker never touches a device).  State lives in a leaked heap object: no destructor races at exit.
struct Worker {
    std::mutex mu;                       // guards queue / inflight (g_mu guards the caches; never hold both)
    std::condition_variable cv_work, cv_done;
    std::deque<std::pair<uint64_t, std::string>> queue;
    std::unordered_map<uint64_t, bool> inflight;
    bool started = false;
    bool stopping = false;               // the process is exiting: no new compiles
    int running = 0;                     // compiles in progress
    int reregistered = 0;
};
Worker& worker() {
    static Worker* w = new Worker();
    return *w;
}

void worker_atexit();

void worker_main() {
    Worker& w = worker();
    for (;;) {
        std::pair<uint64_t, std::string> job;
        {
            std::unique_lock<std::mutex> lk(w.mu);
            w.cv_work.wait(lk, [&] { return !w.queue.empty() && !w.stopping; });
            job = std::move(w.queue.front());
            w.queue.pop_front();
            ++w.running;
        }
        const auto t0 = std::chrono::steady_clock::now();
        auto k = std::make_shared<JitKernel>();
        std::string log;
        const std::string err = compile_cubin(job.second, k->cubin, log);
        {
            std::lock_guard<std::mutex> lock(g_mu);
            g_last_log = log;
            if (err.empty()) {
                k->source = job.second;
                cache_store(job.first, job.second, k->cubin);
                ++g_stats.compiles;
                g_stats.compile_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                g_stats.last_cubin_bytes = (int64_t)k->cubin.size();
                g_cache[job.first] = k;
            } else {
                ++g_stats.failures;
                g_failed[job.first] = true;
                if (!g_warned) {
                    g_warned = true;
                    std::fprintf(stderr, "qsim_b200: run-time specialisation unavailable (%s); using the interpreter kernel\n", err.c_str());
                }
            }
        }
        {
            std::lock_guard<std::mutex> lk(w.mu);
            w.inflight.erase(job.first);
            --w.running;
            // (NVRTC registers exit handlers of its own lazily, inside its first compiles: ours has to be registered after
            // them to run before them)
            if (w.reregistered < 8) { ++w.reregistered; std::atexit(worker_atexit); }
        }
        w.cv_done.notify_all();
    }
}

// Registered with atexit when the first worker starts (after NVRTC and the CUDA runtime have been initialised, so it runs
// BEFORE their own exit handlers): a process must not run into the static destructors of those libraries while a detached
// worker is still inside nvrtcCompileProgram (seen as a segmentation fault at exit of a test process whose last circuits had
// queued compiles nobody waited for).  Queued jobs are dropped, compiles in progress get to finish.
void worker_atexit() {
    Worker& w = worker();
    std::unique_lock<std::mutex> lk(w.mu);
    w.stopping = true;
    w.queue.clear();
    w.cv_done.wait_for(lk, std::chrono::seconds(30), [&] { return w.running == 0; });
}

}  // namespace

struct JitRequest {
    uint64_t key = 0;
    std::string compute;
};

namespace {
struct DualPolicy {
    std::atomic<int> mode{1};        // 0 off, 1 auto, 2 always (whenever possible)
    std::atomic<int> min_fp64{350};
    DualPolicy() {
        if (const char* e = std::getenv("QSIM_DUAL")) {
            const std::string v(e);
            mode = (v == "off" || v == "0") ? 0 : ((v == "always" || v == "2") ? 2 : 1);
        }
        if (const char* e = std::getenv("QSIM_DUAL_MIN_FP64")) min_fp64 = std::atoi(e);
    }
};
DualPolicy& dual_policy() {
    static DualPolicy p;
    return p;
}
}  // namespace

bool jit_dual_autotune() {
    static const bool on = [] {
        const char* e = std::getenv("QSIM_DUAL_AUTOTUNE");
        return !(e && (std::string(e) == "0" || std::string(e) == "off"));
    }();
    return on && dual_policy().mode == 1;   // (mode always: every eligible pass takes the two-group build, the tests rely on it)
}

void jit_set_dual(int mode, int min_fp64) {
    if (mode >= 0 && mode <= 2) dual_policy().mode = mode;
    if (min_fp64 >= 0) dual_policy().min_fp64 = min_fp64;
}

bool jit_dual_wanted(const PassDesc& pd, const DevOp* ops) {
    const int mode = dual_policy().mode;
    if (mode == 0 || !ops || !jit_dual_possible(pd)) return false;
    return mode == 2 || jit_fp64_estimate(pd, ops) >= dual_policy().min_fp64;
}

std::shared_ptr<JitRequest> jit_make_request(const PassDesc& pd, const DevOp* ops, bool dual) {
    auto rq = std::make_shared<JitRequest>();
    rq->compute = jit_generate_compute(pd, ops, dual);
    static const uint64_t skeleton = fnv1a(kSrcPassDesc) * 31 + fnv1a(kSrcPassDevice) * 17 + fnv1a(kSrcKernelBody);
    rq->key = fnv1a(rq->compute) ^ ((uint64_t)QSIM_REG_BITS << 56) ^ (skeleton * 0x9E3779B97F4A7C15ULL);
    return rq;
}

bool jit_async_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("QSIM_JIT_ASYNC");
        return !(e && (std::string(e) == "0" || std::string(e) == "off"));
    }();
    return on;
}

std::shared_ptr<JitKernel> jit_lookup(const JitRequest& rq, bool needs_device, bool async, bool* pending) {
    if (pending) *pending = false;
    const JitMode mode = jit_mode();
    auto fail_locked = [&](const std::string& what) -> std::shared_ptr<JitKernel> {
        ++g_stats.failures;
        g_failed[rq.key] = true;
        if (mode == JitMode::Always) throw std::runtime_error("qsim_b200 jit: " + what);
        if (!g_warned) {
            g_warned = true;
            std::fprintf(stderr, "qsim_b200: run-time specialisation unavailable (%s); using the interpreter kernel\n", what.c_str());
        }
        return nullptr;
    };
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_cache.find(rq.key);
        if (it != g_cache.end() && it->second->source == rq.compute) {
            if (needs_device && !ensure_loaded_locked(*it->second)) return fail_locked("loading the compiled kernel failed");
            ++g_stats.cache_hits;
            return it->second;
        }
        if (g_failed.count(rq.key) && mode != JitMode::Always) return nullptr;
        // a kernel of this structure compiled by an earlier process?
        auto k = std::make_shared<JitKernel>();
        if (cache_load(rq.key, rq.compute, k->cubin)) {
            k->source = rq.compute;
            if (!needs_device || ensure_loaded_locked(*k)) {
                ++g_stats.disk_hits;
                g_stats.last_cubin_bytes = (int64_t)k->cubin.size();
                g_cache[rq.key] = k;
                return k;
            }
        }
    }
    if (async && mode != JitMode::Always) {
        if (!nvrtc().ok()) {
            std::lock_guard<std::mutex> lock(g_mu);
            return fail_locked(nvrtc().why);
        }
        Worker& w = worker();
        std::lock_guard<std::mutex> lk(w.mu);
        if (!w.started) {
            w.started = true;
            unsigned n_workers = std::thread::hardware_concurrency() / 2;   // the passes of a circuit compile side by side
            n_workers = n_workers < 1 ? 1 : (n_workers > 4 ? 4 : n_workers);
            if (const char* e = std::getenv("QSIM_JIT_THREADS")) n_workers = (unsigned)std::max(1, std::atoi(e));
            for (unsigned i = 0; i < n_workers; ++i) std::thread(worker_main).detach();
            std::atexit(worker_atexit);
        }
        if (!w.inflight.count(rq.key)) {
            // a caller that streams many one-off programs (gate-by-gate application on a large state) must not pile up
            // compiles nobody will wait for: beyond a backlog the pass simply stays on the interpreter
            if (w.queue.size() >= 64) return nullptr;
            w.inflight[rq.key] = true;
            w.queue.emplace_back(rq.key, rq.compute);
            w.cv_work.notify_one();
        }
        if (pending) *pending = true;
        return nullptr;
    }
    // synchronous compile (pre-compiled circuits with `specialise`, mode always, the GPU-less build check)
    const auto t0 = std::chrono::steady_clock::now();
    auto k = std::make_shared<JitKernel>();
    std::string log;
    const std::string err = compile_cubin(rq.compute, k->cubin, log);
    std::lock_guard<std::mutex> lock(g_mu);
    g_last_log = log;
    if (!err.empty()) return fail_locked(err);
    k->source = rq.compute;
    if (needs_device && !ensure_loaded_locked(*k)) return fail_locked("loading the compiled kernel failed");
    cache_store(rq.key, rq.compute, k->cubin);
    ++g_stats.compiles;
    g_stats.compile_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_stats.last_cubin_bytes = (int64_t)k->cubin.size();
    g_cache[rq.key] = k;
    return k;
}

void jit_shutdown() {
    if (worker().started) worker_atexit();
}

void jit_wait_all() {
    Worker& w = worker();
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv_done.wait(lk, [&] { return w.inflight.empty() || w.stopping; });
}

std::shared_ptr<JitKernel> jit_get_kernel(const PassDesc& pd, const DevOp* ops, bool needs_device, bool dual) {
    const auto rq = jit_make_request(pd, ops, dual);
    return jit_lookup(*rq, needs_device, /*async=*/false, nullptr);
}

std::string jit_translation_unit(const PassDesc& pd, const DevOp* ops, bool dual) { return tu_from_compute(jit_generate_compute(pd, ops, dual)); }

size_t jit_copy_cubin(const JitKernel& k, void* out, size_t cap) {
    if (out && cap) std::memcpy(out, k.cubin.data(), k.cubin.size() < cap ? k.cubin.size() : cap);
    return k.cubin.size();
}

cudaError_t jit_launch(JitKernel& k, const PassParams& params, const void* tmap, const void* tmap_keep, const void* tmap_send,
                       unsigned grid, size_t smem, cudaStream_t stream) {
    if (!k.kernel) return cudaErrorInvalidDeviceFunction;
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev)) return e;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (dev >= 0 && dev < 64 && !k.attr_set[dev]) {
            if (cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k.kernel),
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynamicSmem))
                return e;
            k.attr_set[dev] = true;
        }
        ++g_stats.launches;
    }
    void* args[] = {const_cast<PassParams*>(&params), const_cast<void*>(tmap), const_cast<void*>(tmap_keep),
                    const_cast<void*>(tmap_send)};
    return cudaLaunchKernel(reinterpret_cast<const void*>(k.kernel), dim3(grid), dim3(kComputeThreads), args, smem, stream);
}

}  // namespace b200
}  // namespace qsim
