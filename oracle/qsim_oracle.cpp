// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's algorithm for the
// gate-application / read-out hot path.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load this library; the product
// (cuda_quantum_simulator_b200/csrc) never does and has no CPU fallback.
//
// Parity status: PINNED.  tests/test_oracle.py checks this file against
//   (1) the reference's own known-answer vectors transcribed from
//       tests/test_gates.cu:39-386 (tests/golden/known_answers.json), and
//   (2) the unmodified reference CPUSimulator built into oracle/_ref (bit-exact for
//       every gate CPUSimulator implements; fixtures in tests/golden/*.npz).
// CRY / CRZ / Toffoli are NOT implemented by the reference's CPU path
// (src/Simulator.cu:214-220, 289-317 silently skip them); they follow the GPU
// kernels src/Gates.cu:322-410 here and are pinned by tests/test_gates.cu:258-386
// and, on the GPU box, by the reference's own kernels (ref_gpu_run).
//
// Conventions: qubit q <-> bit q of the amplitude index.
// State layout: interleaved (re, im) doubles, 2^n amplitudes.
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

using cplx = std::complex<double>;

struct orc_gate {  // == qsim_gate_t (include/qsim_b200.h)
    int32_t type;  // GateType order, include/Circuit.hpp:42-59
    int32_t q0, q1, q2;
    double param;
};

enum : int {
    G_X, G_Y, G_Z, G_H, G_S, G_T, G_SDAG, G_TDAG, G_RX, G_RY, G_RZ,
    G_CNOT, G_CZ, G_CRY, G_CRZ, G_SWAP, G_TOFFOLI
};

static const double kInvSqrt2 = 0.70710678118654752440;  // include/Constants.hpp:44

// ---------------------------------------------------------------------------
// State-vector gates
// ---------------------------------------------------------------------------

// Visit every (i0, i1) pair that differs only in bit `t`; restates the index
// construction of src/Simulator.cu:226-229 (== getPairIndices, src/Gates.cu:19-25).
template <class F>
static void for_pairs(int n, int t, F&& body) {
    const uint64_t half = 1ULL << (n - 1), low = (1ULL << t) - 1;
    for (uint64_t k = 0; k < half; ++k) {
        uint64_t i0 = (k & low) | ((k & ~low) << 1);
        body(i0, i0 | (1ULL << t));
    }
}

// One-qubit gates: arithmetic forms follow src/Simulator.cu:234-285 term by term so the
// result is bit-identical to the reference CPU path under the same compiler.
static void one_qubit(cplx* a, int n, int type, int t, double theta) {
    const cplx I(0, 1);
    switch (type) {
        case G_X: for_pairs(n, t, [&](uint64_t i, uint64_t j) { std::swap(a[i], a[j]); }); break;
        case G_Y: for_pairs(n, t, [&](uint64_t i, uint64_t j) {
                      cplx u = a[i], v = a[j]; a[i] = cplx(0, -1) * v; a[j] = cplx(0, 1) * u; }); break;
        case G_Z: for_pairs(n, t, [&](uint64_t, uint64_t j) { a[j] = -a[j]; }); break;
        case G_H: for_pairs(n, t, [&](uint64_t i, uint64_t j) {
                      cplx u = a[i], v = a[j]; a[i] = (u + v) * kInvSqrt2; a[j] = (u - v) * kInvSqrt2; }); break;
        case G_S:    for_pairs(n, t, [&](uint64_t, uint64_t j) { a[j] = cplx(0, 1) * a[j]; }); break;
        case G_SDAG: for_pairs(n, t, [&](uint64_t, uint64_t j) { a[j] = cplx(0, -1) * a[j]; }); break;
        case G_T:    for_pairs(n, t, [&](uint64_t, uint64_t j) { a[j] = cplx(kInvSqrt2, kInvSqrt2) * a[j]; }); break;
        case G_TDAG: for_pairs(n, t, [&](uint64_t, uint64_t j) { a[j] = cplx(kInvSqrt2, -kInvSqrt2) * a[j]; }); break;
        case G_RX: for_pairs(n, t, [&](uint64_t i, uint64_t j) {
                       double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
                       cplx u = a[i], v = a[j];
                       a[i] = c * u - cplx(0, s) * v;
                       a[j] = -cplx(0, s) * u + c * v; }); break;
        case G_RY: for_pairs(n, t, [&](uint64_t i, uint64_t j) {
                       double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
                       cplx u = a[i], v = a[j];
                       a[i] = c * u - s * v;
                       a[j] = s * u + c * v; }); break;
        case G_RZ: for_pairs(n, t, [&](uint64_t i, uint64_t j) {
                       double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
                       a[i] = cplx(c, -s) * a[i];
                       a[j] = cplx(c, s) * a[j]; }); break;
        default: break;
    }
    (void)I;
}

// Two-/three-qubit gates.  CNOT/CZ/SWAP: src/Simulator.cu:289-317 (full index scan).
// CRY/CRZ/Toffoli: src/Gates.cu:322-410 (component-wise real arithmetic, no FMA here).
static void multi_qubit(cplx* a, int n, const orc_gate& g) {
    const uint64_t N = 1ULL << n;
    const int c = g.q0, t = g.q1;
    switch (g.type) {
        case G_CNOT:
            for (uint64_t i = 0; i < N; ++i)
                if (((i >> c) & 1) && !((i >> t) & 1)) std::swap(a[i], a[i ^ (1ULL << t)]);
            break;
        case G_CZ:
            for (uint64_t i = 0; i < N; ++i)
                if (((i >> c) & 1) && ((i >> t) & 1)) a[i] = -a[i];
            break;
        case G_SWAP:
            for (uint64_t i = 0; i < N; ++i)
                if (!((i >> c) & 1) && ((i >> t) & 1)) std::swap(a[i], a[i ^ (1ULL << c) ^ (1ULL << t)]);
            break;
        case G_CRY: {
            const double co = std::cos(g.param / 2.0), si = std::sin(g.param / 2.0);
            for (uint64_t i = 0; i < N; ++i)
                if (((i >> c) & 1) && !((i >> t) & 1)) {
                    uint64_t j = i ^ (1ULL << t);
                    cplx u = a[i], v = a[j];
                    a[i] = cplx(co * u.real() - si * v.real(), co * u.imag() - si * v.imag());
                    a[j] = cplx(si * u.real() + co * v.real(), si * u.imag() + co * v.imag());
                }
            break;
        }
        case G_CRZ: {
            const double h = g.param / 2.0, co = std::cos(h), si = std::sin(h);
            for (uint64_t i = 0; i < N; ++i)
                if ((i >> c) & 1) {
                    cplx u = a[i];
                    if ((i >> t) & 1) a[i] = cplx(co * u.real() - si * u.imag(), co * u.imag() + si * u.real());
                    else              a[i] = cplx(co * u.real() + si * u.imag(), co * u.imag() - si * u.real());
                }
            break;
        }
        case G_TOFFOLI: {
            const int c2 = g.q1, tt = g.q2;
            for (uint64_t i = 0; i < N; ++i)
                if (((i >> c) & 1) && ((i >> c2) & 1) && !((i >> tt) & 1)) std::swap(a[i], a[i ^ (1ULL << tt)]);
            break;
        }
        default: break;
    }
}

static bool gate_ok(int n, const orc_gate& g) {
    auto in = [&](int q) { return q >= 0 && q < n; };
    if (g.type < 0 || g.type > G_TOFFOLI) return false;
    if (g.type <= G_RZ) return in(g.q0);
    if (g.type == G_TOFFOLI) return in(g.q0) && in(g.q1) && in(g.q2) && g.q0 != g.q1 && g.q0 != g.q2 && g.q1 != g.q2;
    return in(g.q0) && in(g.q1) && g.q0 != g.q1;
}

ORC_API int orc_apply_gate(double* state, int n, const orc_gate* g) {
    if (!gate_ok(n, *g)) return -1;
    cplx* a = reinterpret_cast<cplx*>(state);
    if (g->type <= G_RZ) one_qubit(a, n, g->type, g->q0, g->param);
    else multi_qubit(a, n, *g);
    return 0;
}

ORC_API void orc_init_zero(double* state, int n) {
    std::memset(state, 0, sizeof(double) * 2 * (1ULL << n));
    state[0] = 1.0;
}

// Simulator::run semantics (src/Simulator.cu:28-36): gates compose on the current state.
ORC_API int orc_run(double* state, int n, const orc_gate* g, int64_t ng) {
    for (int64_t i = 0; i < ng; ++i)
        if (orc_apply_gate(state, n, g + i)) return -1;
    return 0;
}

// ---------------------------------------------------------------------------
// Read-out: probabilities, sampling, measurement
// ---------------------------------------------------------------------------

// p[i] = std::norm(a[i]) = re*re + im*im, two roundings then one (src/Simulator.cu:319-325).
ORC_API void orc_probabilities(const double* state, int n, double* probs) {
    const uint64_t N = 1ULL << n;
    for (uint64_t i = 0; i < N; ++i) {
        double re = state[2 * i], im = state[2 * i + 1];
        probs[i] = re * re + im * im;
    }
}

// Sequential index-order sum (StateVector::getTotalProbability, src/StateVector.cu:235-242).
ORC_API double orc_total_probability(const double* probs, int64_t N) {
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += probs[i];
    return s;
}

// std::partial_sum CDF + std::lower_bound per draw (src/Simulator.cu:164-185): the result is the
// smallest i with cum[i] >= r; N if r exceeds the rounded total (the reference does not clamp).
ORC_API void orc_sample(const double* probs, int64_t N, const double* uniforms, int64_t shots, int64_t* out) {
    std::vector<double> cum(N);
    std::partial_sum(probs, probs + N, cum.begin());
    for (int64_t s = 0; s < shots; ++s)
        out[s] = std::lower_bound(cum.begin(), cum.end(), uniforms[s]) - cum.begin();
}

// The uniforms the reference draws: std::uniform_real_distribution<double>(0,1) over std::mt19937(seed)
// (src/NoiseModel.cu:351-354, 605-610; libstdc++ consumes two 32-bit words per double).
ORC_API void orc_mt19937_uniforms(unsigned seed, int64_t count, double* out) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    for (int64_t i = 0; i < count; ++i) out[i] = dist(rng);
}

// P(bit `bitpos` == 0): index-order host sum of |a_i|^2 over matching indices
// (src/StateVector.cu:83-101, 280-287).
ORC_API double orc_prob_zero(const double* state, int n, int bitpos) {
    const uint64_t N = 1ULL << n;
    double p0 = 0.0;
    for (uint64_t i = 0; i < N; ++i) {
        double v = 0.0;
        if (!((i >> bitpos) & 1)) { double re = state[2 * i], im = state[2 * i + 1]; v = re * re + im * im; }
        p0 += v;
    }
    return p0;
}

// StateVector::measure given the bit position and the uniform draw r (src/StateVector.cu:260-314):
// result = r < p0 ? 0 : 1; p_result = result ? 1 - p0 : p0; error if < 1e-15; survivors scaled by
// 1/sqrt(p_result), others zeroed.  Note Simulator::measureQubit(q) uses bitpos = n-1-q.
// Returns the outcome, or -1 for the zero-probability error.
ORC_API int orc_measure(double* state, int n, int bitpos, double r, double* p0_out) {
    const uint64_t N = 1ULL << n;
    double p0 = orc_prob_zero(state, n, bitpos);
    if (p0_out) *p0_out = p0;
    int result = (r < p0) ? 0 : 1;
    double pr = result == 0 ? p0 : 1.0 - p0;
    if (pr < 1e-15) return -1;
    double f = 1.0 / std::sqrt(pr);
    for (uint64_t i = 0; i < N; ++i) {
        if (int((i >> bitpos) & 1) != result) { state[2 * i] = 0.0; state[2 * i + 1] = 0.0; }
        else { state[2 * i] *= f; state[2 * i + 1] *= f; }
    }
    return result;
}

// NoisySimulator::measureQubit (src/NoiseModel.cu:615-651): bit q itself; p0 summed over host
// probabilities; collapse divides by sqrt(sum of surviving |a|^2).
ORC_API int orc_measure_noisy(double* state, int n, int qubit, double r) {
    const uint64_t N = 1ULL << n;
    cplx* a = reinterpret_cast<cplx*>(state);
    double p0 = 0.0;
    for (uint64_t i = 0; i < N; ++i)
        if (!((i >> qubit) & 1)) p0 += std::norm(a[i]);
    int result = (r < p0) ? 0 : 1;
    double norm = 0.0;
    for (uint64_t i = 0; i < N; ++i) {
        if (int((i >> qubit) & 1) == result) norm += std::norm(a[i]);
        else a[i] = cplx(0, 0);
    }
    norm = std::sqrt(norm);
    for (uint64_t i = 0; i < N; ++i) a[i] /= norm;
    return result;
}

// ---------------------------------------------------------------------------
// Density matrices: exact gates and exact (textbook) Kraus channels.
// rho is row-major 2^n x 2^n (src/DensityMatrix.cu:23-32): element (r, c) at r*2^n + c.
// Channels follow the definitions the reference cites (include/NoiseModel.cuh:14-18,
// include/DensityMatrix.cuh:250-264), NOT its defective kernels.
// ---------------------------------------------------------------------------

static void gate_matrix_1q(int type, double theta, cplx U[2][2]) {
    const double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
    U[0][0] = U[1][1] = 1; U[0][1] = U[1][0] = 0;
    switch (type) {
        case G_X: U[0][0] = 0; U[0][1] = 1; U[1][0] = 1; U[1][1] = 0; break;
        case G_Y: U[0][0] = 0; U[0][1] = cplx(0, -1); U[1][0] = cplx(0, 1); U[1][1] = 0; break;
        case G_Z: U[1][1] = -1; break;
        case G_H: U[0][0] = U[0][1] = U[1][0] = kInvSqrt2; U[1][1] = -kInvSqrt2; break;
        case G_S: U[1][1] = cplx(0, 1); break;
        case G_SDAG: U[1][1] = cplx(0, -1); break;
        case G_T: U[1][1] = cplx(kInvSqrt2, kInvSqrt2); break;
        case G_TDAG: U[1][1] = cplx(kInvSqrt2, -kInvSqrt2); break;
        case G_RX: U[0][0] = c; U[0][1] = cplx(0, -s); U[1][0] = cplx(0, -s); U[1][1] = c; break;
        case G_RY: U[0][0] = c; U[0][1] = -s; U[1][0] = s; U[1][1] = c; break;
        case G_RZ: U[0][0] = cplx(c, -s); U[1][1] = cplx(c, s); break;
        default: break;
    }
}

// v' = (controlled) U on bit t of an m-bit vector, active where (i & cmask) == cmask.
static void apply_c1q(cplx* v, int m, int t, uint64_t cmask, const cplx U[2][2]) {
    const uint64_t half = 1ULL << (m - 1), low = (1ULL << t) - 1;
    for (uint64_t k = 0; k < half; ++k) {
        uint64_t i0 = (k & low) | ((k & ~low) << 1), i1 = i0 | (1ULL << t);
        if ((i0 & cmask) != cmask) continue;
        cplx x = v[i0], y = v[i1];
        v[i0] = U[0][0] * x + U[0][1] * y;
        v[i1] = U[1][0] * x + U[1][1] * y;
    }
}

// Unitary gate on rho viewed as a 2n-bit vector: U on the row bit (q+n), conj(U) on the column bit q.
static void dm_controlled(cplx* rho, int n, int t, uint64_t cmask, const cplx U[2][2]) {
    cplx Uc[2][2] = {{std::conj(U[0][0]), std::conj(U[0][1])}, {std::conj(U[1][0]), std::conj(U[1][1])}};
    apply_c1q(rho, 2 * n, t + n, cmask << n, U);
    apply_c1q(rho, 2 * n, t, cmask, Uc);
}

ORC_API void orc_dm_init_zero(double* rho, int n) {
    std::memset(rho, 0, sizeof(double) * 2 * (1ULL << (2 * n)));
    rho[0] = 1.0;
}

ORC_API void orc_dm_from_pure(double* rho, int n, const double* state) {
    const uint64_t D = 1ULL << n;
    const cplx* a = reinterpret_cast<const cplx*>(state);
    cplx* r = reinterpret_cast<cplx*>(rho);
    for (uint64_t i = 0; i < D; ++i)
        for (uint64_t j = 0; j < D; ++j) r[i * D + j] = a[i] * std::conj(a[j]);
}

ORC_API int orc_dm_apply_gate(double* rho_, int n, const orc_gate* g) {
    if (!gate_ok(n, *g)) return -1;
    cplx* rho = reinterpret_cast<cplx*>(rho_);
    cplx U[2][2];
    switch (g->type) {
        case G_CNOT: gate_matrix_1q(G_X, 0, U); dm_controlled(rho, n, g->q1, 1ULL << g->q0, U); break;
        case G_CZ:   gate_matrix_1q(G_Z, 0, U); dm_controlled(rho, n, g->q1, 1ULL << g->q0, U); break;
        case G_CRY:  gate_matrix_1q(G_RY, g->param, U); dm_controlled(rho, n, g->q1, 1ULL << g->q0, U); break;
        case G_CRZ:  gate_matrix_1q(G_RZ, g->param, U); dm_controlled(rho, n, g->q1, 1ULL << g->q0, U); break;
        case G_SWAP:
            gate_matrix_1q(G_X, 0, U);
            dm_controlled(rho, n, g->q1, 1ULL << g->q0, U);
            dm_controlled(rho, n, g->q0, 1ULL << g->q1, U);
            dm_controlled(rho, n, g->q1, 1ULL << g->q0, U);
            break;
        case G_TOFFOLI:
            gate_matrix_1q(G_X, 0, U);
            dm_controlled(rho, n, g->q2, (1ULL << g->q0) | (1ULL << g->q1), U);
            break;
        default: gate_matrix_1q(g->type, g->param, U); dm_controlled(rho, n, g->q0, 0, U); break;
    }
    return 0;
}

// NoiseType order: include/NoiseModel.cuh:46-53.
enum : int { N_DEPOL, N_AMPDAMP, N_PHASEDAMP, N_BITFLIP, N_PHASEFLIP, N_BITPHASEFLIP };

// rho' = sum_k K_k rho K_k^dagger on qubit q, evaluated per 2x2 block.
ORC_API int orc_dm_channel(double* rho_, int n, int type, int q, double p) {
    if (q < 0 || q >= n) return -1;
    cplx* rho = reinterpret_cast<cplx*>(rho_);
    std::vector<std::array<cplx, 4>> K;  // row-major 2x2 each
    const cplx I(0, 1);
    auto pauli = [&](double w, char which) {
        double s = std::sqrt(w);
        switch (which) {
            case 'I': K.push_back({s, 0, 0, s}); break;
            case 'X': K.push_back({0, s, s, 0}); break;
            case 'Y': K.push_back({0, -I * s, I * s, 0}); break;
            case 'Z': K.push_back({s, 0, 0, -s}); break;
        }
    };
    switch (type) {
        case N_DEPOL: pauli(1 - p, 'I'); pauli(p / 3, 'X'); pauli(p / 3, 'Y'); pauli(p / 3, 'Z'); break;
        case N_AMPDAMP: K.push_back({1, 0, 0, std::sqrt(1 - p)}); K.push_back({0, std::sqrt(p), 0, 0}); break;
        case N_PHASEDAMP: K.push_back({1, 0, 0, std::sqrt(1 - p)}); K.push_back({0, 0, 0, std::sqrt(p)}); break;
        case N_BITFLIP: pauli(1 - p, 'I'); pauli(p, 'X'); break;
        case N_PHASEFLIP: pauli(1 - p, 'I'); pauli(p, 'Z'); break;
        case N_BITPHASEFLIP: pauli(1 - p, 'I'); pauli(p, 'Y'); break;
        default: return -1;
    }
    const uint64_t D = 1ULL << n, bit = 1ULL << q;
    for (uint64_t r = 0; r < D; ++r) {
        if (r & bit) continue;
        for (uint64_t c = 0; c < D; ++c) {
            if (c & bit) continue;
            cplx B[2][2] = {{rho[r * D + c], rho[r * D + (c | bit)]},
                            {rho[(r | bit) * D + c], rho[(r | bit) * D + (c | bit)]}};
            cplx R[2][2] = {{0, 0}, {0, 0}};
            for (auto& k : K) {
                const cplx M[2][2] = {{k[0], k[1]}, {k[2], k[3]}};
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j)
                        for (int x = 0; x < 2; ++x)
                            for (int y = 0; y < 2; ++y) R[i][j] += M[i][x] * B[x][y] * std::conj(M[j][y]);
            }
            rho[r * D + c] = R[0][0]; rho[r * D + (c | bit)] = R[0][1];
            rho[(r | bit) * D + c] = R[1][0]; rho[(r | bit) * D + (c | bit)] = R[1][1];
        }
    }
    return 0;
}

ORC_API void orc_dm_probabilities(const double* rho, int n, double* probs) {
    const uint64_t D = 1ULL << n;
    for (uint64_t i = 0; i < D; ++i) probs[i] = rho[2 * (i * D + i)];
}

ORC_API double orc_dm_trace(const double* rho, int n) {
    const uint64_t D = 1ULL << n;
    double t = 0;
    for (uint64_t i = 0; i < D; ++i) t += rho[2 * (i * D + i)];
    return t;
}

// Purity as the reference computes it: sum of |rho_ij|^2 (src/DensityMatrix.cu:147-167).
ORC_API double orc_dm_purity(const double* rho, int n) {
    const uint64_t E = 1ULL << (2 * n);
    double s = 0;
    for (uint64_t i = 0; i < E; ++i) s += rho[2 * i] * rho[2 * i] + rho[2 * i + 1] * rho[2 * i + 1];
    return s;
}

// DensityMatrixSimulator::measureQubit with an injected uniform (src/DensityMatrix.cu:374-406):
// result = (u < p1) ? 1 : 0 (compared against p1, not p0); rho <- P rho P / p.
ORC_API int orc_dm_measure(double* rho_, int n, int q, double u) {
    cplx* rho = reinterpret_cast<cplx*>(rho_);
    const uint64_t D = 1ULL << n;
    double p1 = 0;
    for (uint64_t i = 0; i < D; ++i) if ((i >> q) & 1) p1 += rho[i * D + i].real();
    int result = (u < p1) ? 1 : 0;
    double p = result ? p1 : 1.0 - p1;
    for (uint64_t r = 0; r < D; ++r)
        for (uint64_t c = 0; c < D; ++c) {
            bool keep = int((r >> q) & 1) == result && int((c >> q) & 1) == result;
            rho[r * D + c] = keep ? rho[r * D + c] / p : cplx(0, 0);
        }
    return result;
}

// ---------------------------------------------------------------------------
// Quantum trajectories (per-trajectory unravelling, SURVEY.md Appendix B right column).
// The reference has no CPU trajectory code and its GPU kernels draw per amplitude pair (D6),
// so this section restates OUR documented schedule, driven by Philox4x32-10 (Salmon et al.,
// SC'11; the published algorithm, pinned by Random123's known-answer vectors in
// tests/test_oracle.py), so GPU trajectories can be checked amplitude-for-amplitude.
// ---------------------------------------------------------------------------

ORC_API void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n1 = uint32_t(p1);
        uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1, n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two uniforms in [0,1) with 53 bits each for noise event `event` of trajectory `traj`.
// counter = (event_lo, event_hi, traj_lo, traj_hi), key = (seed, 0x51534D42 "QSMB").
static void traj_uniforms(uint32_t seed, uint64_t traj, uint64_t event, double& u0, double& u1) {
    uint32_t ctr[4] = {uint32_t(event), uint32_t(event >> 32), uint32_t(traj), uint32_t(traj >> 32)};
    uint32_t key[2] = {seed, 0x51534D42u}, o[4];
    orc_philox4x32_10(ctr, key, o);
    u0 = double(((uint64_t(o[1]) << 32) | o[0]) >> 11) * 0x1.0p-53;
    u1 = double(((uint64_t(o[3]) << 32) | o[2]) >> 11) * 0x1.0p-53;
}

ORC_API void orc_traj_uniforms(uint32_t seed, uint64_t traj, uint64_t event, double* u) {
    traj_uniforms(seed, traj, event, u[0], u[1]);
}

struct orc_channel { int32_t type; int32_t qubit; double p; };  // one (channel, qubit) noise event

static void traj_noise_event(cplx* a, int n, const orc_channel& ch, double u0, double u1) {
    const uint64_t N = 1ULL << n, bit = 1ULL << ch.qubit;
    cplx U[2][2];
    auto pauli = [&](int g) { gate_matrix_1q(g, 0, U); apply_c1q(a, n, ch.qubit, 0, U); };
    switch (ch.type) {
        case N_DEPOL:
            if (u0 < ch.p) { if (u1 < 1.0 / 3.0) pauli(G_X); else if (u1 < 2.0 / 3.0) pauli(G_Y); else pauli(G_Z); }
            break;
        case N_BITFLIP: if (u0 < ch.p) pauli(G_X); break;
        case N_PHASEFLIP: if (u0 < ch.p) pauli(G_Z); break;
        case N_BITPHASEFLIP: if (u0 < ch.p) pauli(G_Y); break;
        case N_AMPDAMP:
        case N_PHASEDAMP: {
            double P1 = 0;
            for (uint64_t i = 0; i < N; ++i) if (i & bit) P1 += std::norm(a[i]);
            const double g = ch.p;
            if (u0 < g * P1) {  // jump
                const double f = 1.0 / std::sqrt(P1);
                for (uint64_t i = 0; i < N; ++i) {
                    if (i & bit) continue;
                    if (ch.type == N_AMPDAMP) { a[i] = a[i | bit] * f; a[i | bit] = 0; }
                    else { a[i] = 0; a[i | bit] = a[i | bit] * f; }
                }
            } else {            // no jump: K0 = diag(1, sqrt(1-g)), renormalised
                const double k = std::sqrt(1.0 - g), f = 1.0 / std::sqrt(1.0 - g * P1);
                for (uint64_t i = 0; i < N; ++i) a[i] = (i & bit) ? a[i] * (k * f) : a[i] * f;
            }
            break;
        }
        default: break;
    }
}

// One trajectory: after every gate, every event in `ev[0..nev)` in order (Noisy/Batched schedule,
// src/NoiseModel.cu:369-382, 815-831, with the empty-list-means-all fix D7 applied by the caller).
// Event counter: event = (gate_index << 16) | event_index (independent of the number of events per gate).
ORC_API int orc_traj_run(double* state, int n, const orc_gate* g, int64_t ng,
                         const orc_channel* ev, int64_t nev, uint32_t seed, uint64_t traj) {
    cplx* a = reinterpret_cast<cplx*>(state);
    for (int64_t i = 0; i < ng; ++i) {
        if (orc_apply_gate(state, n, g + i)) return -1;
        for (int64_t e = 0; e < nev; ++e) {
            double u0, u1;
            traj_uniforms(seed, traj, (uint64_t(i) << 16) | uint64_t(e), u0, u1);
            traj_noise_event(a, n, ev[e], u0, u1);
        }
    }
    return 0;
}

// Exact-channel counterpart with the same schedule, for the statistical check of trajectory averages.
ORC_API int orc_dm_run_schedule(double* rho, int n, const orc_gate* g, int64_t ng,
                                const orc_channel* ev, int64_t nev) {
    for (int64_t i = 0; i < ng; ++i) {
        if (orc_dm_apply_gate(rho, n, g + i)) return -1;
        for (int64_t e = 0; e < nev; ++e)
            if (orc_dm_channel(rho, n, ev[e].type, ev[e].qubit, ev[e].p)) return -1;
    }
    return 0;
}
