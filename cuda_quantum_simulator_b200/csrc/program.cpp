// Circuit compiler: gate list -> merged ops -> passes -> sweeps -> device records.
// See program.hpp for the model.  Gate matrices follow SURVEY.md Appendix A, i.e. the
// reference kernels src/Gates.cu:31-410 and CPU path src/Simulator.cu:222-317.
#include "program.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace qsim {
namespace b200 {

namespace {

constexpr double kInvSqrt2 = 0.70710678118654752440;  // reference include/Constants.hpp:44

inline void set_m(LogicalOp& o, double a_re, double a_im, double b_re, double b_im, double c_re, double c_im,
                  double d_re, double d_im) {
    o.m[0] = a_re; o.m[1] = a_im; o.m[2] = b_re; o.m[3] = b_im;
    o.m[4] = c_re; o.m[5] = c_im; o.m[6] = d_re; o.m[7] = d_im;
}

LogicalOp make_op(int target, uint64_t cmask, int gate_index) {
    LogicalOp o{};
    o.kind = OP_MAT;
    o.target = target;
    o.cmask = cmask;
    o.cval = cmask;
    o.first_gate = gate_index;
    o.n_gates = 1;
    return o;
}

// 2x2 complex product out = b * a (a is applied first).
void matmul(const double* b, const double* a, double* out) {
    auto at = [](const double* m, int r, int c, int part) { return m[(r * 2 + c) * 2 + part]; };
    double tmp[8];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
            double re = 0, im = 0;
            for (int k = 0; k < 2; ++k) {
                double xr = at(b, r, k, 0), xi = at(b, r, k, 1), yr = at(a, k, c, 0), yi = at(a, k, c, 1);
                re += xr * yr - xi * yi;
                im += xr * yi + xi * yr;
            }
            tmp[(r * 2 + c) * 2] = re;
            tmp[(r * 2 + c) * 2 + 1] = im;
        }
    std::memcpy(out, tmp, sizeof(tmp));
}

inline bool is_diag_kind(uint8_t k) { return k == OP_DIAG || k == OP_PHASE; }

inline uint64_t qubits_of(const LogicalOp& o) { return o.cmask | (1ULL << o.target); }

// Two controlled one-qubit operators commute when each non-diagonal target is untouched by the
// other operator (diagonal operators and control projectors are all diagonal in the same basis).
bool commutes(const LogicalOp& a, const LogicalOp& b) {
    if (!is_diag_kind(a.kind) && (qubits_of(b) >> a.target) & 1) return false;
    if (!is_diag_kind(b.kind) && (qubits_of(a) >> b.target) & 1) return false;
    return true;
}

bool is_identity(const LogicalOp& o) {
    return o.m[0] == 1.0 && o.m[1] == 0.0 && o.m[2] == 0.0 && o.m[3] == 0.0 && o.m[4] == 0.0 && o.m[5] == 0.0 &&
           o.m[6] == 1.0 && o.m[7] == 0.0;
}

}  // namespace

CompileOptions default_options() {
    CompileOptions o;
    if (const char* e = std::getenv("QSIM_TILE_BITS")) { int v = std::atoi(e); if (v >= 3 && v <= kMaxTileBits) o.max_tile_bits = v; }
    if (const char* e = std::getenv("QSIM_MIN_LOW_BITS")) { int v = std::atoi(e); if (v >= 3 && v <= kMaxTileBits) o.min_low_bits = v; }
    if (std::getenv("QSIM_NO_MERGE")) o.merge = false;
    if (std::getenv("QSIM_NO_REORDER")) o.reorder = false;
    if (std::getenv("QSIM_NO_PHASE")) o.fuse_diagonals = false;
    if (std::getenv("QSIM_NO_TAIL")) o.fold_tail_flips = false;
    return o;
}

void classify(LogicalOp& o) {
    const double* m = o.m;
    const bool off_zero = m[2] == 0 && m[3] == 0 && m[4] == 0 && m[5] == 0;
    const bool on_zero = m[0] == 0 && m[1] == 0 && m[6] == 0 && m[7] == 0;
    const bool real = m[1] == 0 && m[3] == 0 && m[5] == 0 && m[7] == 0;
    if (off_zero) o.kind = OP_DIAG;
    else if (on_zero && real && m[2] == 1.0 && m[4] == 1.0) o.kind = OP_FLIP;
    else if (on_zero) o.kind = OP_ADIAG;
    else if (real) o.kind = OP_MATREAL;
    else o.kind = OP_MAT;
}

bool lower_gate(const qsim_gate_t& g, std::vector<LogicalOp>& out, int gi) {
    const double k = kInvSqrt2;
    const double c = std::cos(g.param / 2.0), s = std::sin(g.param / 2.0);
    auto one = [&](int target, uint64_t cmask) -> LogicalOp& {
        out.push_back(make_op(target, cmask, gi));
        return out.back();
    };
    auto flip = [&](int target, uint64_t cmask) { set_m(one(target, cmask), 0, 0, 1, 0, 1, 0, 0, 0); };
    switch (g.type) {
        case QSIM_GATE_X: flip(g.q0, 0); break;
        case QSIM_GATE_Y: set_m(one(g.q0, 0), 0, 0, 0, -1, 0, 1, 0, 0); break;
        case QSIM_GATE_Z: set_m(one(g.q0, 0), 1, 0, 0, 0, 0, 0, -1, 0); break;
        case QSIM_GATE_H: set_m(one(g.q0, 0), k, 0, k, 0, k, 0, -k, 0); break;
        case QSIM_GATE_S: set_m(one(g.q0, 0), 1, 0, 0, 0, 0, 0, 0, 1); break;
        case QSIM_GATE_T: set_m(one(g.q0, 0), 1, 0, 0, 0, 0, 0, k, k); break;
        case QSIM_GATE_SDAG: set_m(one(g.q0, 0), 1, 0, 0, 0, 0, 0, 0, -1); break;
        case QSIM_GATE_TDAG: set_m(one(g.q0, 0), 1, 0, 0, 0, 0, 0, k, -k); break;
        case QSIM_GATE_RX: set_m(one(g.q0, 0), c, 0, 0, -s, 0, -s, c, 0); break;
        case QSIM_GATE_RY: set_m(one(g.q0, 0), c, 0, -s, 0, s, 0, c, 0); break;
        case QSIM_GATE_RZ: set_m(one(g.q0, 0), c, -s, 0, 0, 0, 0, c, s); break;
        case QSIM_GATE_CNOT: flip(g.q1, 1ULL << g.q0); break;
        case QSIM_GATE_CZ: set_m(one(g.q1, 1ULL << g.q0), 1, 0, 0, 0, 0, 0, -1, 0); break;
        case QSIM_GATE_CRY: set_m(one(g.q1, 1ULL << g.q0), c, 0, -s, 0, s, 0, c, 0); break;
        case QSIM_GATE_CRZ: set_m(one(g.q1, 1ULL << g.q0), c, -s, 0, 0, 0, 0, c, s); break;
        case QSIM_GATE_SWAP:  // swap(a,b) = CX(a,b) CX(b,a) CX(a,b): three index permutations
            flip(g.q1, 1ULL << g.q0);
            flip(g.q0, 1ULL << g.q1);
            flip(g.q1, 1ULL << g.q0);
            break;
        case QSIM_GATE_TOFFOLI: flip(g.q2, (1ULL << g.q0) | (1ULL << g.q1)); break;
        default: return false;
    }
    for (size_t i = out.size(); i-- > 0 && out[i].first_gate == gi;) classify(out[i]);
    return true;
}

namespace {

// Merge `b` into an earlier op on the same (target, controls) when everything in between
// commutes with it; otherwise append.
void append_merged(std::vector<LogicalOp>& out, const LogicalOp& b, bool merge) {
    if (merge) {
        const size_t window = 96;
        for (size_t back = 0; back < window && back < out.size(); ++back) {
            LogicalOp& a = out[out.size() - 1 - back];
            if (a.target == b.target && a.cmask == b.cmask && a.cval == b.cval) {
                matmul(b.m, a.m, a.m);
                a.n_gates += b.n_gates;
                classify(a);
                return;
            }
            if (!commutes(a, b)) break;
        }
    }
    out.push_back(b);
}

struct PassPlan {
    std::vector<int> op_idx;   // indices into the logical op list, execution order
    uint64_t need = 0;         // qubits that must be tile bits (non-diagonal targets)
};

// Can a pass whose required qubit set is `need` be tiled with t tile bits, keeping lmin low bits?
bool tile_feasible(uint64_t need, int n_local, int t, int lmin) {
    if (t >= n_local) return true;
    int high = __builtin_popcountll(need >> lmin);
    return lmin + high <= t;
}

void choose_tile_bits(uint64_t need, int n_local, int t, int lmin, PassDesc& pd, int isolate_bit = -1) {
    // Largest L with L + |need ∩ [L, n)| <= t (f is non-decreasing in L).
    int L = std::min(t, n_local);
    if (t < n_local) {
        L = lmin;
        while (L + 1 <= t && (L + 1) + __builtin_popcountll(need >> (L + 1)) <= t) ++L;
    }
    std::vector<int> bits;
    for (int q = 0; q < L; ++q) bits.push_back(q);
    for (int q = L; q < n_local; ++q)
        if ((need >> q) & 1) bits.push_back(q);
    // pad with the lowest unused qubits (keeps runs long and the tile count a power of two)
    for (int q = L; q < n_local && (int)bits.size() < t; ++q)
        if (!((need >> q) & 1)) bits.push_back(q);
    std::sort(bits.begin(), bits.end());
    int LL = 0;
    while (LL < (int)bits.size() && bits[LL] == LL) ++LL;
    pd.t = (int)bits.size();
    pd.L = LL;
    pd.n_high = pd.t - pd.L;
    for (int i = 0; i < pd.t && i < kMaxTileBits; ++i) pd.tile_bits[i] = (uint8_t)bits[i];
    // Tensor-map geometry: contiguous runs of tile bits become box dimensions (dim 0 at most 7 bits =
    // 256 doubles, the others at most 8 bits); each dimension's coordinate range reaches up to the next
    // one.  Runs beyond the fifth dimension are enumerated by separate instructions.
    {
        std::vector<std::pair<int, int>> chunks;   // (first bit, length)
        for (size_t i = 0; i < bits.size();) {
            size_t j = i;
            const int cap = chunks.empty() ? 7 : 8;
            while (j + 1 < bits.size() && bits[j + 1] == bits[j] + 1 && (int)(j + 1 - i) < cap) ++j;
            chunks.emplace_back(bits[i], (int)(j - i + 1));
            i = j + 1;
        }
        // the highest tile bit on request in a chunk of its own that separate instructions enumerate (CompileOptions::isolate_bit)
        bool isolated = false;
        if (isolate_bit >= 0 && !bits.empty() && bits.back() == isolate_bit && chunks.size() >= 1 && bits.size() >= 2) {
            if (chunks.back().second > 1) {
                --chunks.back().second;
                chunks.emplace_back(isolate_bit, 1);
            }
            isolated = chunks.size() >= 2;
        }
        const int nd = std::min<int>(5, (int)chunks.size() - (isolated ? 1 : 0));
        int instr_bits = 0;
        for (size_t c = nd; c < chunks.size(); ++c) instr_bits += chunks[c].second;
        pd.tma_instr_bits = (uint8_t)instr_bits;
        for (int d = 0; d < 5; ++d) {
            TmaDim& td = pd.tma_dim[d];
            if (d < nd) {
                const int start = d == 0 ? 0 : chunks[d].first;
                const int end = (d + 1 < nd) ? chunks[d + 1].first : n_local;
                td.start_bit = (uint8_t)start;
                td.range_bits = (uint8_t)(end - start);
                td.box_bits = (uint8_t)chunks[d].second;
            } else {
                td.start_bit = (uint8_t)n_local;   // degenerate dimension of extent 1
                td.range_bits = 0;
                td.box_bits = 0;
            }
            td.pad = 0;
        }
    }
    // segments of outer (non-tile) bits
    pd.n_segments = 0;
    uint64_t tmask = 0;
    for (int b : bits) tmask |= 1ULL << b;
    int src = 0;
    for (int q = 0; q < n_local;) {
        if ((tmask >> q) & 1) { ++q; continue; }
        int q0 = q;
        while (q < n_local && !((tmask >> q) & 1)) ++q;
        Segment& sg = pd.seg[pd.n_segments++];
        int len = q - q0;
        sg.mask = (len >= 64) ? ~0ULL : ((1ULL << len) - 1);
        sg.src_shift = (uint8_t)src;
        sg.dst_shift = (uint8_t)q0;
        src += len;
    }
}

}  // namespace

namespace {

struct Cx {
    double r, i;
};
inline Cx cmul(Cx a, Cx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
inline Cx cdiv(Cx a, Cx b) {
    const double d = b.r * b.r + b.i * b.i;
    return {(a.r * b.r + a.i * b.i) / d, (a.i * b.r - a.r * b.i) / d};
}

// Turns a run of diagonal ops into the multiplicative degree-2 polynomial  f = C * prod A_q^{b_q} * prod B_pq^{b_p b_q}
// and splits it by where the bits live: inside-only terms go to a table over the tile-local index, terms that involve
// bits outside the tile become PhaseTerm records evaluated once per tile.
void encode_phase(Program& out, PassDesc& pd, const std::vector<int>& members, const int* local_of, int nl, DevOp& d) {
    Cx C{1, 0};
    std::vector<Cx> A(64, Cx{1, 0});
    std::vector<std::pair<std::pair<int, int>, Cx>> B;   // ((p, q), factor), p < q
    auto addB = [&](int p, int q, Cx f) {
        if (p > q) std::swap(p, q);
        for (auto& e : B)
            if (e.first.first == p && e.first.second == q) { e.second = cmul(e.second, f); return; }
        B.push_back({{p, q}, f});
    };
    for (int mi : members) {
        const LogicalOp& op = out.lops[mi];
        const Cx d0{op.m[0], op.m[1]}, d1{op.m[6], op.m[7]};
        const int t = op.target;
        if (op.cmask == 0) {                       // factor d0 * (d1/d0)^{b_t}
            C = cmul(C, d0);
            A[t] = cmul(A[t], cdiv(d1, d0));
        } else {
            const int c = __builtin_ctzll(op.cmask);
            if ((op.cval >> c) & 1) {              // active when b_c = 1: [d0 (d1/d0)^{b_t}]^{b_c}
                A[c] = cmul(A[c], d0);
                addB(c, t, cdiv(d1, d0));
            } else {                               // active when b_c = 0: d0 (d1/d0)^{b_t} * [d0^-1 (d0/d1)^{b_t}]^{b_c}
                C = cmul(C, d0);
                A[t] = cmul(A[t], cdiv(d1, d0));
                A[c] = cmul(A[c], cdiv(Cx{1, 0}, d0));
                addB(c, t, cdiv(d0, d1));
            }
        }
    }
    auto inside = [&](int q) { return q < nl && local_of[q] >= 0; };
    // table over the tile-local index: C, inside singles, inside-inside pairs
    const size_t base = out.phase_tables.size();
    out.phase_tables.resize(base + 2 * (size_t)kPhaseTableSize);
    const uint32_t n_entries = 1u << pd.t;
    for (uint32_t l = 0; l < (uint32_t)kPhaseTableSize; ++l) {
        Cx f{1, 0};
        if (l < n_entries) {
            f = C;
            for (int q = 0; q < 64; ++q)
                if (inside(q) && ((l >> local_of[q]) & 1) && (A[q].r != 1.0 || A[q].i != 0.0)) f = cmul(f, A[q]);
            for (auto& e : B) {
                const int p = e.first.first, q = e.first.second;
                if (inside(p) && inside(q) && ((l >> local_of[p]) & 1) && ((l >> local_of[q]) & 1)) f = cmul(f, e.second);
            }
        }
        out.phase_tables[base + 2 * l] = f.r;
        out.phase_tables[base + 2 * l + 1] = f.i;
    }
    // The table only depends on the tile bits that carry an inside single or take part in an inside-inside pair - often a
    // handful (a CRZ ladder inside a tile of which half the bits are idle low qubits).  A COMPACT copy indexed by just those
    // bits (entry e = full entry at pdep(e, dep_mask)) follows the full one: a few hundred bytes to 2 KiB that stay in L1,
    // where the full 64 KiB table is re-read from L2 for every tile (the specialised kernels index it with literals; the
    // interpreter and the emulator keep the full table).
    uint32_t dep_mask = 0;
    for (int q = 0; q < 64; ++q)
        if (inside(q) && (A[q].r != 1.0 || A[q].i != 0.0)) dep_mask |= 1u << local_of[q];
    for (auto& e : B) {
        const int p = e.first.first, q = e.first.second;
        if (inside(p) && inside(q) && (e.second.r != 1.0 || e.second.i != 0.0)) dep_mask |= (1u << local_of[p]) | (1u << local_of[q]);
    }
    const int dep_bits = __builtin_popcount(dep_mask);
    const size_t cbase = out.phase_tables.size();
    out.phase_tables.resize(cbase + 2 * ((size_t)1 << dep_bits));
    for (uint32_t e = 0; e < (1u << dep_bits); ++e) {
        uint32_t l = 0, rest = e;
        for (int j = 0; j < kMaxTileBits; ++j)
            if ((dep_mask >> j) & 1u) { l |= (rest & 1u) << j; rest >>= 1; }
        out.phase_tables[cbase + 2 * e] = out.phase_tables[base + 2 * (size_t)l];
        out.phase_tables[cbase + 2 * e + 1] = out.phase_tables[base + 2 * (size_t)l + 1];
    }
    d.tmask_thr = dep_mask;                                                          // tile bits the table depends on
    d.cval_thr = (uint32_t)((cbase / 2) - (size_t)pd.phase_table_offset);            // compact table, entry offset within the pass
    // terms with outside bits, grouped by the factor they feed: E_0..E_11, then U
    std::vector<std::vector<PhaseTerm>> groups(13);
    auto term = [&](int kind, int o, int j, Cx f) {
        PhaseTerm t{};
        t.kind = (uint8_t)kind; t.o = (uint8_t)o; t.j = (uint8_t)j; t.fr = f.r; t.fi = f.i;
        return t;
    };
    for (int q = 0; q < 64; ++q)
        if (!inside(q) && (A[q].r != 1.0 || A[q].i != 0.0)) groups[12].push_back(term(1, q, 0, A[q]));
    for (auto& e : B) {
        const int p = e.first.first, q = e.first.second;
        const bool ip = inside(p), iq = inside(q);
        if (ip && iq) continue;
        if (!ip && !iq) groups[12].push_back(term(2, p, q, e.second));
        else if (ip) groups[local_of[p]].push_back(term(0, q, local_of[p], e.second));
        else groups[local_of[q]].push_back(term(0, p, local_of[q], e.second));
    }
    const size_t first_term = out.phase_terms.size();
    uint16_t starts[14];
    size_t cursor = 0;
    for (int e = 0; e < 13; ++e) {
        starts[e] = (uint16_t)cursor;
        for (auto& t : groups[e]) out.phase_terms.push_back(t);
        cursor += groups[e].size();
    }
    starts[13] = (uint16_t)cursor;
    d.kind = OP_PHASE;
    d.opcode = kOpcodePhase;
    d.slotmask = 0xffff;
    d.tslots = (uint16_t)cursor;                                              // number of terms
    d.cmask_out = (uint64_t)(base / 2) - (uint64_t)pd.phase_table_offset;     // table entry offset within the pass
    d.cval_out = (uint64_t)first_term - (uint64_t)pd.phase_term_offset;      // first term within the pass
    d.tmask_out = (uint64_t)pd.n_phase;                                       // shared-memory slot of (E, U)
    // which factors are not identically 1: bit j (j < 12) = E_j has terms, bit 12 = U has terms, bit 13 = the table is
    // not constant (bit 14 = not even the constant 1)
    uint32_t present = 0;
    for (int e = 0; e < 13; ++e)
        if (!groups[e].empty()) present |= 1u << e;
    {
        const double* tb = out.phase_tables.data() + base;
        bool constant = true;
        for (uint32_t l = 1; l < n_entries && constant; ++l) constant = (tb[2 * l] == tb[0] && tb[2 * l + 1] == tb[1]);
        if (!constant) present |= 1u << 13;
        if (!constant || tb[0] != 1.0 || tb[1] != 0.0) present |= 1u << 14;
    }
    d.cmask_thr = present;
    std::memcpy(d.m, starts, sizeof(starts));
    pd.n_phase++;
}

}  // namespace

namespace {

using Fail = bool (*)(std::string*, const std::string&);
inline bool set_error(std::string* error, const std::string& s) { if (error) *error = s; return false; }

// Step 1 of compile_ops: the X frame and the merge of gate runs on the same (target, controls).
bool merge_with_x_frame(const std::vector<LogicalOp>& lops_in, const CompileOptions& opt, int nl, Program& out, uint64_t* xf_local_out,
                        std::string* error) {
    auto fail = [&](const std::string& s) { return set_error(error, s); };
    // 0. X frame: an uncontrolled X is not executed; it toggles a pending index-XOR mask through which
    //    later ops are conjugated (X on their target: swap the matrix's rows and columns; X on a control:
    //    the control fires on 0).  What is left at the end is applied by the last pass's addressing.
    uint64_t xf = opt.initial_xor;
    // 1. merge
    for (auto& op_in : lops_in) {
        LogicalOp op = op_in;
        if (opt.defer_x) {
            if (op.kind == OP_FLIP && op.cmask == 0) { xf ^= 1ULL << op.target; continue; }
            if ((xf >> op.target) & 1) {
                std::swap(op.m[0], op.m[6]); std::swap(op.m[1], op.m[7]);
                std::swap(op.m[2], op.m[4]); std::swap(op.m[3], op.m[5]);
                classify(op);
            }
            op.cval ^= (xf & op.cmask);
        }
        if (!is_diag_kind(op.kind) && op.target >= nl)
            return fail("non-diagonal gate on a global qubit must be remapped before compilation");
        append_merged(out.lops, op, opt.merge);
    }
    out.global_xor = nl < 64 ? (xf >> nl) : 0;
    const uint64_t xf_local = nl < 64 ? (xf & ((1ULL << nl) - 1)) : xf;
    {
        std::vector<LogicalOp> kept;
        for (auto& op : out.lops)
            if (!(is_diag_kind(op.kind) && is_identity(op))) kept.push_back(op);
        out.lops.swap(kept);
    }

    *xf_local_out = xf_local;
    return true;
}

// Step 2: list scheduling of the merged ops into passes.
bool schedule_passes(const CompileOptions& opt, int nl, int t, int lmin, uint64_t xf_local, const Program& out, std::vector<PassPlan>& plans,
                     std::string* error) {
    auto fail = [&](const std::string& s) { return set_error(error, s); };
    // 2. passes: list scheduling — an op joins the open pass if its target fits the tile and it
    //    commutes with every op already deferred to a later pass.
    std::vector<int> pending(out.lops.size());
    for (size_t i = 0; i < pending.size(); ++i) pending[i] = (int)i;
    while (!pending.empty()) {
        PassPlan plan;
        std::vector<int> deferred;
        // Sweep budget: a sweep can target five tile bits besides the lowest three, so every time a sixth distinct
        // target shows up among the ops taken in order, another sweep starts (the sweep scheduler below reorders with
        // commutation and only does better).  The pass is closed before that estimate exceeds what PassDesc holds.
        uint64_t sweep_targets = 0;
        int sweeps_est = 1;
        bool closed = false;
        for (int idx : pending) {
            const LogicalOp& op = out.lops[idx];
            uint64_t need = plan.need | (is_diag_kind(op.kind) ? 0ULL : (1ULL << op.target));
            bool ok = !closed && tile_feasible(need, nl, t, lmin) && (int)plan.op_idx.size() < kMaxOpsPerPass;
            if (ok && !is_diag_kind(op.kind) && !((sweep_targets >> op.target) & 1ULL)) {
                if (__builtin_popcountll(sweep_targets) >= 5) {
                    if (sweeps_est + 1 >= kMaxSweeps) { ok = false; closed = true; }
                    else { ++sweeps_est; sweep_targets = 0; }
                }
                if (ok) sweep_targets |= 1ULL << op.target;
            }
            if (ok && !deferred.empty()) {
                if (!opt.reorder) ok = false;
                else
                    for (int d : deferred)
                        if (!commutes(out.lops[d], op)) { ok = false; break; }
            }
            if (ok) { plan.op_idx.push_back(idx); plan.need = need; }
            else deferred.push_back(idx);
        }
        if (plan.op_idx.empty()) return fail("scheduler made no progress");
        plans.push_back(std::move(plan));
        pending.swap(deferred);
    }

    if (plans.empty() && xf_local) plans.emplace_back();   // nothing but a deferred X: one pure permutation pass

    return true;
}

// Step 3, once per pass: tile bits, folded flips, fused diagonal runs, sweeps, device records.
bool build_pass(PassPlan& plan, bool last_pass, const CompileOptions& opt, int nl, int t, int lmin, uint64_t xf_local, Program& out,
                std::string* error) {
    auto fail = [&](const std::string& s) { return set_error(error, s); };
    PassDesc pd{};
    pd.n = nl;
    choose_tile_bits(plan.need, nl, t, lmin, pd, opt.isolate_bit);
    int local_of[64];
    for (int q = 0; q < 64; ++q) local_of[q] = -1;
    for (int i = 0; i < pd.t; ++i) local_of[pd.tile_bits[i]] = i;

    const int r = pd.t > 5 + kMaxRegBits ? kMaxRegBits : std::max(0, pd.t - 5);
    const int nthr = pd.t - r;
    const int n_lane = std::min(5, nthr);
    const int cap = n_lane + r;  // targetable positions per sweep

    pd.op_offset = (int)out.ops.size();
    pd.n_ops = 0;
    pd.n_sweeps = 0;

    if (last_pass && xf_local) {
        // the deferred X frame rides on the last pass: tile bits -> XOR of the tile-local store index,
        // other bits -> the tile is written to the partner tile's location
        uint64_t tau_bit = 0;
        int src = 0;
        for (int q = 0; q < nl; ++q) {
            if (local_of[q] >= 0) { if ((xf_local >> q) & 1) pd.xor_local |= 1u << local_of[q]; }
            else { if ((xf_local >> q) & 1) tau_bit |= 1ULL << src; ++src; }
        }
        pd.xor_tau = tau_bit;
    }

    // Trailing / leading bit flips.  Walking backwards, a FLIP that commutes with every op that stays behind it can
    // be moved to the end of the pass, where it costs only index arithmetic in the final store; walking forwards,
    // one that commutes with every op that stays in front of it can be moved to the start, into the first sweep's
    // load.  Only flips that are affine in the tile-local index qualify (see TailDyn).
    struct AffineMap {
        uint16_t lin[kMaxTileBits];
        uint16_t cst = 0;
        TailDyn dyn[kMaxTailDyn];
        int n_dyn = 0, n = 0;
    };
    auto affine_ok = [&](const LogicalOp& op, int n_dyn_so_far, bool* is_dyn) {
        if (op.kind != OP_FLIP || op.target >= nl || local_of[op.target] < 0) return false;
        int n_in = 0, n_out = 0;
        for (int q = 0; q < 64; ++q)
            if ((op.cmask >> q) & 1) ((q < nl && local_of[q] >= 0) ? n_in : n_out)++;
        *is_dyn = n_out > 0;
        return *is_dyn ? (n_in == 0 && n_dyn_so_far < kMaxTailDyn) : (n_in <= 1);
    };
    // F = f_m o ... o f_1 for the flips in execution order: F(x) = lin x ^ cst ^ (fired translations)
    auto compose = [&](const std::vector<int>& flips, AffineMap& M) {
        for (int j = 0; j < kMaxTileBits; ++j) M.lin[j] = (uint16_t)(1u << j);
        for (int idx : flips) {
            const LogicalOp& op = out.lops[idx];
            const uint16_t et = (uint16_t)(1u << local_of[op.target]);
            uint64_t cm_out = 0, cv_out = 0;
            int cb = -1, cv = 1;
            for (int q = 0; q < 64; ++q) {
                if (!((op.cmask >> q) & 1)) continue;
                const uint64_t v = (op.cval >> q) & 1;
                if (q < nl && local_of[q] >= 0) { cb = local_of[q]; cv = (int)v; }
                else { cm_out |= 1ULL << q; cv_out |= v << q; }
            }
            if (cm_out) {
                TailDyn& d = M.dyn[M.n_dyn++];
                d.cmask_out = cm_out; d.cval_out = cv_out; d.w = et;
            } else if (cb < 0) {
                M.cst ^= et;
            } else {
                // l_t ^= l_cb (^ 1 for a control on zero), composed after everything folded so far
                for (int j = 0; j < kMaxTileBits; ++j) if ((M.lin[j] >> cb) & 1) M.lin[j] ^= et;
                for (int d = 0; d < M.n_dyn; ++d) if ((M.dyn[d].w >> cb) & 1) M.dyn[d].w ^= et;
                if ((M.cst >> cb) & 1) M.cst ^= et;
                if (!cv) M.cst ^= et;
            }
            ++M.n;
        }
    };
    auto apply_lin = [](const uint16_t (&lin)[kMaxTileBits], unsigned x) {
        unsigned v = 0;
        for (int j = 0; j < kMaxTileBits; ++j) if ((x >> j) & 1) v ^= lin[j];
        return v;
    };
    // inverse of a linear map over GF(2)^12 given by the images of the unit vectors (Gauss-Jordan on [A | I])
    auto invert_gf2 = [](const uint16_t (&lin)[kMaxTileBits], uint16_t (&inv)[kMaxTileBits]) {
        uint32_t rows[kMaxTileBits];   // row i: bits 0..11 = A[i][*], bits 16..27 = I[i][*]
        for (int i = 0; i < kMaxTileBits; ++i) {
            uint32_t r = 1u << (16 + i);
            for (int j = 0; j < kMaxTileBits; ++j) if ((lin[j] >> i) & 1) r |= 1u << j;   // A[i][j] = bit i of the image of e_j
            rows[i] = r;
        }
        for (int c = 0; c < kMaxTileBits; ++c) {
            int piv = -1;
            for (int i = c; i < kMaxTileBits; ++i) if ((rows[i] >> c) & 1) { piv = i; break; }
            if (piv < 0) return false;
            std::swap(rows[c], rows[piv]);
            for (int i = 0; i < kMaxTileBits; ++i) if (i != c && ((rows[i] >> c) & 1)) rows[i] ^= rows[c];
        }
        for (int j = 0; j < kMaxTileBits; ++j) {
            unsigned v = 0;
            for (int i = 0; i < kMaxTileBits; ++i) if ((rows[i] >> (16 + j)) & 1) v |= 1u << i;
            inv[j] = (uint16_t)v;
        }
        return true;
    };
    for (int j = 0; j < kMaxTileBits; ++j) pd.tail_lin[j] = pd.head_lin[j] = (uint16_t)(1u << j);
    if (opt.fold_tail_flips) {
        std::vector<int> stay, tail;   // both in reverse order
        int n_dyn = 0;
        for (size_t k = plan.op_idx.size(); k-- > 0;) {
            const int idx = plan.op_idx[k];
            const LogicalOp& op = out.lops[idx];
            bool dyn = false;
            bool movable = (int)tail.size() < kMaxTailFlips && affine_ok(op, n_dyn, &dyn);
            if (movable)
                for (int s2 : stay)
                    if (!commutes(op, out.lops[s2])) { movable = false; break; }
            if (movable && dyn) ++n_dyn;
            (movable ? tail : stay).push_back(idx);
        }
        if (!tail.empty()) {
            std::reverse(stay.begin(), stay.end());
            std::reverse(tail.begin(), tail.end());
            AffineMap M;
            compose(tail, M);
            std::memcpy(pd.tail_lin, M.lin, sizeof(M.lin));
            pd.tail_const = M.cst;
            pd.n_dyn = M.n_dyn;
            for (int d = 0; d < M.n_dyn; ++d) pd.dyn[d] = M.dyn[d];
            pd.n_tail = M.n;
            plan.op_idx.swap(stay);
        }
        // leading flips (of what is left)
        std::vector<int> keep, head;
        n_dyn = 0;
        for (int idx : plan.op_idx) {
            const LogicalOp& op = out.lops[idx];
            bool dyn = false;
            bool movable = (int)head.size() < kMaxTailFlips && affine_ok(op, n_dyn, &dyn);
            if (movable)
                for (int s2 : keep)
                    if (!commutes(op, out.lops[s2])) { movable = false; break; }
            if (movable && dyn) ++n_dyn;
            (movable ? head : keep).push_back(idx);
        }
        if (!head.empty()) {
            AffineMap M;
            compose(head, M);
            uint16_t inv[kMaxTileBits];
            if (!invert_gf2(M.lin, inv)) return fail("internal: folded flips are not invertible");
            std::memcpy(pd.head_lin, inv, sizeof(inv));
            pd.head_const = (uint16_t)apply_lin(pd.head_lin, M.cst);
            pd.n_head_dyn = M.n_dyn;
            for (int d = 0; d < M.n_dyn; ++d) {
                pd.head_dyn[d] = M.dyn[d];
                pd.head_dyn[d].w = (uint16_t)apply_lin(pd.head_lin, M.dyn[d].w);
            }
            pd.n_head = M.n;
            plan.op_idx.swap(keep);
        }
    }

    // Fuse runs of diagonal gates (each with at most one control and no zero entry) into OP_PHASE ops:
    // a diagonal op may slide back to the open run as long as nothing emitted since the run started has its
    // non-diagonal target among the op's qubits.
    std::vector<std::vector<int>> clusters;   // members (indices into out.lops) of each fused run
    if (opt.fuse_diagonals) {
        auto eligible = [&](const LogicalOp& op) {
            if (op.kind != OP_DIAG || __builtin_popcountll(op.cmask) > 1) return false;
            return (op.m[0] != 0.0 || op.m[1] != 0.0) && (op.m[6] != 0.0 || op.m[7] != 0.0);
        };
        std::vector<int> order;               // >= 0: op index, < 0: -(cluster id) - 1
        int open = -1;
        uint64_t blocked = 0;                 // non-diagonal targets emitted since the open run started
        for (int idx : plan.op_idx) {
            const LogicalOp& op = out.lops[idx];
            if (eligible(op)) {
                if (open >= 0 && (qubits_of(op) & blocked) == 0) { clusters[open].push_back(idx); continue; }
                if ((int)clusters.size() < kMaxPhaseOps) {
                    open = (int)clusters.size();
                    clusters.push_back({idx});
                    blocked = 0;
                    order.push_back(-open - 1);
                    continue;
                }
            }
            order.push_back(idx);
            if (!is_diag_kind(op.kind)) blocked |= 1ULL << op.target;
        }
        std::vector<int> rebuilt;
        for (int item : order) {
            if (item >= 0) { rebuilt.push_back(item); continue; }
            const std::vector<int>& mem = clusters[-item - 1];
            if (mem.size() < 3) { for (int m : mem) rebuilt.push_back(m); continue; }   // not worth a table
            LogicalOp ph{};
            ph.kind = OP_PHASE;
            uint64_t qs = 0;
            for (int m : mem) qs |= qubits_of(out.lops[m]);
            ph.target = __builtin_ctzll(qs);
            ph.cmask = qs & ~(1ULL << ph.target);   // qubits_of(ph) == every qubit of the run
            ph.cval = ph.cmask;
            ph.first_gate = -item - 1;              // cluster id
            ph.n_gates = (int)mem.size();
            out.lops.push_back(ph);
            rebuilt.push_back((int)out.lops.size() - 1);
        }
        plan.op_idx.swap(rebuilt);
    }
    pd.phase_table_offset = (int)(out.phase_tables.size() / 2);
    pd.phase_term_offset = (int)out.phase_terms.size();

    std::vector<int> todo = plan.op_idx;
    bool need_empty_sweep = todo.empty();
    while (!todo.empty() || need_empty_sweep) {
        need_empty_sweep = false;
        if (pd.n_sweeps >= kMaxSweeps) return fail("too many sweeps in one pass");
        // choose the targetable set greedily in order (commutation-aware deferral as above)
        uint32_t S = 0;
        for (int p = 0; p < std::min(3, pd.t); ++p) S |= 1u << p;
        std::vector<int> chosen, deferred;
        std::vector<int> uses(pd.t, 0);
        for (int idx : todo) {
            const LogicalOp& op = out.lops[idx];
            bool ok = true;
            uint32_t S2 = S;
            if (!is_diag_kind(op.kind)) {
                S2 |= 1u << local_of[op.target];
                ok = __builtin_popcount(S2) <= cap;
            }
            if (ok && !deferred.empty()) {
                if (!opt.reorder) ok = false;
                else
                    for (int d : deferred)
                        if (!commutes(out.lops[d], op)) { ok = false; break; }
            }
            if (ok) {
                S = S2;
                chosen.push_back(idx);
                if (!is_diag_kind(op.kind)) uses[local_of[op.target]]++;
            } else deferred.push_back(idx);
        }
        if (chosen.empty() && !plan.op_idx.empty()) return fail("sweep scheduler made no progress");

        // role assignment: tid bits 0..2 = tile bits 0..2; busiest other targets -> registers;
        // remaining targets -> lane bits 3,4; everything else -> leftover lane/warp bits.
        SweepDesc sd{};
        sd.r = (uint8_t)r;
        sd.nthr = (uint8_t)nthr;
        std::vector<int> cand;  // targetable positions beyond the fixed low lanes
        for (int p = std::min(3, pd.t); p < pd.t; ++p)
            if ((S >> p) & 1) cand.push_back(p);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return uses[a] > uses[b]; });
        std::vector<int> regs, lanes_extra, rest;
        for (int p : cand) {
            if ((int)regs.size() < r) regs.push_back(p);
            else lanes_extra.push_back(p);
        }
        std::vector<char> taken(pd.t, 0);
        for (int p = 0; p < std::min(3, pd.t); ++p) taken[p] = 1;
        for (int p : regs) taken[p] = 1;
        for (int p : lanes_extra) taken[p] = 1;
        // registers prefer high tile bits when free (keeps lane bits low => coalesced smem rows)
        for (int p = pd.t - 1; p >= 0 && (int)regs.size() < r; --p)
            if (!taken[p]) { regs.push_back(p); taken[p] = 1; }
        std::sort(regs.begin(), regs.end());
        for (int p = 0; p < pd.t; ++p)
            if (!taken[p]) rest.push_back(p);
        int ti = 0;
        for (int p = 0; p < std::min(3, pd.t); ++p) sd.thr_pos[ti++] = (uint8_t)p;
        for (int p : lanes_extra) sd.thr_pos[ti++] = (uint8_t)p;
        for (int p : rest) sd.thr_pos[ti++] = (uint8_t)p;
        if (ti != nthr) return fail("internal: thread-bit assignment");
        for (int j = 0; j < r; ++j) sd.reg_pos[j] = (uint8_t)regs[j];
        for (int k = 0; k < 16; ++k) {
            unsigned off = 0;
            for (int j = 0; j < r; ++j)
                if ((k >> j) & 1) off |= 1u << regs[j];
            sd.slot_off[k] = (uint16_t)off;
        }
        int role_tid[kMaxTileBits], role_reg[kMaxTileBits];
        for (int p = 0; p < pd.t; ++p) role_tid[p] = role_reg[p] = -1;
        for (int i = 0; i < nthr; ++i) role_tid[sd.thr_pos[i]] = i;
        for (int j = 0; j < r; ++j) role_reg[sd.reg_pos[j]] = j;

        // flips that lead a later sweep fold into its load (the first sweep's were taken at pass level above)
        for (int j = 0; j < kMaxTileBits; ++j) sd.head_lin[j] = (uint16_t)(1u << j);
        if (opt.fold_tail_flips && pd.n_sweeps > 0) {
            std::vector<int> keep, head;
            for (int idx : chosen) {
                const LogicalOp& op = out.lops[idx];
                bool dyn = false;
                bool movable = (int)head.size() < kMaxTailFlips && affine_ok(op, 0, &dyn) && !dyn;
                if (movable)
                    for (int s2 : keep)
                        if (!commutes(op, out.lops[s2])) { movable = false; break; }
                (movable ? head : keep).push_back(idx);
            }
            if (!head.empty()) {
                AffineMap M;
                compose(head, M);
                uint16_t inv[kMaxTileBits];
                if (!invert_gf2(M.lin, inv)) return fail("internal: folded flips are not invertible");
                std::memcpy(sd.head_lin, inv, sizeof(inv));
                sd.head_const = (uint16_t)apply_lin(sd.head_lin, M.cst);
                sd.n_head = (uint16_t)M.n;
                chosen.swap(keep);
            }
        }
        for (int k = 0; k < 16; ++k) sd.load_slot_off[k] = (uint16_t)apply_lin(sd.head_lin, sd.slot_off[k]);
        sd.op_begin = (uint16_t)pd.n_ops;
        for (int idx : chosen) {
            const LogicalOp& op = out.lops[idx];
            DevOp d{};
            d.kind = op.kind;
            if (op.kind == OP_PHASE) {
                encode_phase(out, pd, clusters[op.first_gate], local_of, nl, d);
                out.ops.push_back(d);
                pd.n_ops++;
                continue;
            }
            std::memcpy(d.m, op.m, sizeof(op.m));
            uint32_t reg_cmask = 0, reg_cval = 0;
            for (int q = 0; q < 64; ++q) {
                if (!((op.cmask >> q) & 1)) continue;
                uint64_t v = (op.cval >> q) & 1;
                int p = q < nl ? local_of[q] : -1;
                if (p < 0) { d.cmask_out |= 1ULL << q; d.cval_out |= v << q; }
                else if (role_reg[p] >= 0) { reg_cmask |= 1u << role_reg[p]; reg_cval |= (uint32_t)v << role_reg[p]; }
                else { d.cmask_thr |= 1u << role_tid[p]; d.cval_thr |= (uint32_t)v << role_tid[p]; }
            }
            for (int k = 0; k < 16; ++k)
                if (((uint32_t)k & reg_cmask) == reg_cval) d.slotmask |= (uint16_t)(1u << k);
            int p = op.target < nl ? local_of[op.target] : -1;
            if (p < 0) {
                if (!is_diag_kind(op.kind)) return fail("internal: non-diagonal target outside tile");
                d.thome = T_OUTSIDE;
                d.tmask_out = 1ULL << op.target;
            } else if (role_reg[p] >= 0) {
                d.thome = T_REG;
                d.tbit = (uint8_t)role_reg[p];
                for (int k = 0; k < 16; ++k)
                    if ((k >> role_reg[p]) & 1) d.tslots |= (uint16_t)(1u << k);
            } else {
                int tb = role_tid[p];
                if (is_diag_kind(op.kind)) { d.thome = T_THREAD; d.tbit = (uint8_t)tb; d.tmask_thr = 1u << tb; }
                else {
                    if (tb >= 5) return fail("internal: non-diagonal target on a warp bit");
                    d.thome = T_LANE;
                    d.tbit = (uint8_t)tb;
                }
            }
            d.has_out = (d.cmask_out != 0 || d.tmask_out != 0) ? 1u : 0u;
            d.opcode = (uint8_t)(op_code(d.kind, d.thome, d.tbit, d.cmask_thr != 0 || d.slotmask != 0xffffu) |
                                 (d.has_out ? 0x80u : 0u));   // bit 7: depends on index bits outside the tile
            out.ops.push_back(d);
            pd.n_ops++;
        }
        sd.op_end = (uint16_t)pd.n_ops;
        pd.sweep[pd.n_sweeps++] = sd;
        todo.swap(deferred);
    }
    {   // the final store's (first load's) slot offsets: the folded flips' linear part applied to the last (first) sweep's
        const SweepDesc& last = pd.sweep[pd.n_sweeps - 1];
        const SweepDesc& first = pd.sweep[0];
        for (int k = 0; k < 16; ++k) {
            unsigned v = 0, w = 0;
            for (int j = 0; j < kMaxTileBits; ++j) {
                if ((last.slot_off[k] >> j) & 1) v ^= pd.tail_lin[j];
                if ((first.slot_off[k] >> j) & 1) w ^= pd.head_lin[j];
            }
            pd.store_slot_off[k] = (uint16_t)v;
            pd.load_slot_off[k] = (uint16_t)w;
        }
    }
    {   // derived fields (see PassDesc)
        auto base_of = [&](uint64_t tau) {
            uint64_t b = 0;
            for (int sg = 0; sg < pd.n_segments; ++sg) b |= ((tau >> pd.seg[sg].src_shift) & pd.seg[sg].mask) << pd.seg[sg].dst_shift;
            return b;
        };
        pd.tile_mask = 0;
        for (int j = 0; j < pd.t; ++j) pd.tile_mask |= 1ULL << pd.tile_bits[j];
        pd.xdep = base_of(pd.xor_tau);
        pd.pivot_dep = pd.xor_tau ? base_of(1ULL << (63 - __builtin_clzll(pd.xor_tau))) : 0ULL;
    }
    out.passes.push_back(pd);
    return true;
}

}  // namespace

bool compile_ops(int n, std::vector<LogicalOp> lops_in, const CompileOptions& opt, Program& out, std::string* error) {
    auto fail = [&](const std::string& s) { if (error) *error = s; return false; };
    out = Program();
    out.n = n;
    out.n_local = n - opt.n_global;
    const int nl = out.n_local;
    if (nl < 1) return fail("no local qubits");
    const int t = std::min(opt.max_tile_bits, nl);
    const int lmin = std::min(std::max(opt.min_low_bits, 3), t);

    uint64_t xf_local = 0;
    if (!merge_with_x_frame(lops_in, opt, nl, out, &xf_local, error)) return false;
    std::vector<PassPlan> plans;
    if (!schedule_passes(opt, nl, t, lmin, xf_local, out, plans, error)) return false;
    for (size_t plan_i = 0; plan_i < plans.size(); ++plan_i)
        if (!build_pass(plans[plan_i], plan_i + 1 == plans.size(), opt, nl, t, lmin, xf_local, out, error)) return false;
    return true;
}

bool compile(int n, const qsim_gate_t* gates, int64_t n_gates, const CompileOptions& opt, Program& out,
             std::string* error) {
    std::vector<LogicalOp> lops;
    lops.reserve((size_t)n_gates + 8);
    for (int64_t i = 0; i < n_gates; ++i)
        if (!lower_gate(gates[i], lops, (int)i)) {
            if (error) *error = "Unknown gate type";
            return false;
        }
    bool ok = compile_ops(n, std::move(lops), opt, out, error);
    out.n_gates = n_gates;
    return ok;
}

std::string Program::describe() const {
    std::ostringstream os;
    size_t sweeps = 0;
    for (auto& p : passes) sweeps += p.n_sweeps;
    os << "program: " << n << " qubits (" << n_local << " local), " << n_gates << " gates -> " << lops.size()
       << " ops, " << passes.size() << " passes, " << sweeps << " sweeps\n";
    static const char* kn[] = {"MAT", "MATREAL", "ADIAG", "FLIP", "DIAG", "PHASE"};
    for (size_t i = 0; i < passes.size(); ++i) {
        const PassDesc& p = passes[i];
        os << "  pass " << i << ": t=" << p.t << " L=" << p.L << " tile_bits=[";
        for (int b = 0; b < p.t; ++b) os << (b ? "," : "") << (int)p.tile_bits[b];
        os << "] ops=" << p.n_ops << " sweeps=" << p.n_sweeps;
        if (p.n_head) os << " head_flips=" << p.n_head << "(" << p.n_head_dyn << " per-tile)";
        if (p.n_tail) os << " tail_flips=" << p.n_tail << "(" << p.n_dyn << " per-tile)";
        if (p.xor_local) os << " xor_local=0x" << std::hex << p.xor_local << std::dec;
        if (p.xor_tau) os << " xor_tau=0x" << std::hex << p.xor_tau << std::dec;
        if (p.tma_instr_bits) os << " tma_instrs=" << (1 << p.tma_instr_bits);
        os << "\n";
        for (int s = 0; s < p.n_sweeps; ++s) {
            const SweepDesc& sd = p.sweep[s];
            os << "    sweep " << s << ": regs=[";
            for (int j = 0; j < sd.r; ++j) os << (j ? "," : "") << (int)p.tile_bits[sd.reg_pos[j]];
            os << "] lanes=[";
            for (int j = 0; j < std::min<int>(5, sd.nthr); ++j) os << (j ? "," : "") << (int)p.tile_bits[sd.thr_pos[j]];
            os << "] ops:";
            for (int o = sd.op_begin; o < sd.op_end; ++o) os << " " << kn[ops[p.op_offset + o].kind];
            os << "\n";
        }
    }
    return os.str();
}

}  // namespace b200
}  // namespace qsim
