// Planning for sharded states: where the global<->local qubit swaps go (pure host logic, no device).
// The reference is single-GPU (README.md:367), so there is no reference counterpart; the contract is "same amplitudes as the
// single-device run" (SURVEY.md 8e).  The Python mirror (cuda_quantum_simulator_b200/sharded.py: plan_circuit,
// choose_initial_layout) implements the same rules and the CPU tests hold the two against each other step by step.
#include "sharded_plan.hpp"

#include <algorithm>
#include <set>
#include <stdexcept>

namespace qsim {
namespace b200 {

namespace {

bool is_diagonal(int t) {   // Z S T Sdag Tdag Rz CZ CRZ
    return t == 2 || t == 4 || t == 5 || t == 6 || t == 7 || t == 10 || t == 12 || t == 14;
}

int qubit_of(const qsim_gate_t& g, int slot) { return slot == 0 ? g.q0 : (slot == 1 ? g.q1 : g.q2); }

}  // namespace

int shard_target_slots(int gate_type, int slots[2]) {
    if (is_diagonal(gate_type)) return 0;
    if (gate_type <= 10) { slots[0] = 0; return 1; }
    if (gate_type == 15) { slots[0] = 0; slots[1] = 1; return 2; }   // SWAP: both are targets
    if (gate_type == 16) { slots[0] = 2; return 1; }
    slots[0] = 1;                                                     // CNOT, CRY
    return 1;
}

std::vector<int> shard_choose_initial_layout(int n, int n_global, const qsim_gate_t* gates, int64_t ng) {
    const int nl = n - n_global;
    const int64_t never = (int64_t)1 << 60;
    std::vector<int64_t> first_use(n, never);
    for (int64_t j = 0; j < ng; ++j) {
        if (gates[j].type == QSIM_GATE_X) continue;
        int slots[2];
        const int ns = shard_target_slots(gates[j].type, slots);
        for (int s = 0; s < ns; ++s) {
            const int q = qubit_of(gates[j], slots[s]);
            if (first_use[q] == never) first_use[q] = j;
        }
    }
    // best candidates for the global positions: latest first non-diagonal use, ties -> highest qubit (identity if possible)
    std::vector<int> order(n);
    for (int q = 0; q < n; ++q) order[q] = q;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (first_use[a] != first_use[b]) return first_use[a] > first_use[b];
        return a > b;
    });
    std::vector<int> glob(order.begin(), order.begin() + n_global);
    std::sort(glob.begin(), glob.end());
    std::vector<int> perm(n, 0);
    int pos = 0;
    for (int q = 0; q < n; ++q)
        if (!std::binary_search(glob.begin(), glob.end(), q)) perm[q] = pos++;
    for (int i = 0; i < n_global; ++i) perm[glob[i]] = nl + i;
    return perm;
}

ShardPlanRec shard_plan_circuit(int n, int n_global, const qsim_gate_t* gates, int64_t ng, const std::vector<int>& perm_in) {
    const int nl = n - n_global;
    if ((int)perm_in.size() != n) throw std::invalid_argument("permutation size does not match the qubit count");
    std::vector<int> perm = perm_in;
    ShardPlanRec plan;
    plan.n = n;
    plan.n_global = n_global;
    std::vector<qsim_gate_t> cur;
    auto flush = [&] {
        if (cur.empty()) return;
        ShardStepRec st;
        st.gates.swap(cur);
        plan.steps.push_back(std::move(st));
        cur.clear();
    };
    const int64_t never = (int64_t)1 << 60;
    for (int64_t idx = 0; idx < ng; ++idx) {
        const qsim_gate_t& g = gates[idx];
        if (g.type != QSIM_GATE_X) {   // an uncontrolled X on a global qubit is a frame toggle, handled by the compiler
            int slots[2];
            const int ns = shard_target_slots(g.type, slots);
            for (int s = 0; s < ns; ++s) {
                const int lq = qubit_of(g, slots[s]);
                if (perm[lq] < nl) continue;
                // tile qubits of the segment's last pass, roughly: the low contiguous run plus the most recent non-diagonal
                // targets.  A victim outside them lets the exchange ride on that pass's store.
                std::set<int> recent;
                for (int p = 0; p < std::min(5, nl); ++p) recent.insert(p);
                for (size_t r = cur.size(); r-- > 0;) {
                    if (recent.size() >= 12) break;
                    if (cur[r].type == QSIM_GATE_X) continue;
                    int rs[2];
                    const int rn = shard_target_slots(cur[r].type, rs);
                    for (int k = 0; k < rn; ++k) recent.insert(qubit_of(cur[r], rs[k]));
                }
                flush();
                // the local position to evict: the one whose next use as a non-diagonal target lies farthest ahead
                std::vector<int64_t> next_use(nl, never);
                std::vector<int> inv(n, 0);
                for (int q = 0; q < n; ++q) inv[perm[q]] = q;
                std::set<int> busy;
                for (int c = 0; c < 3; ++c)
                    if (qubit_of(g, c) >= 0) busy.insert(perm[qubit_of(g, c)]);
                for (int64_t j = idx; j < ng; ++j) {
                    int hs[2];
                    const int hn = shard_target_slots(gates[j].type, hs);
                    for (int k = 0; k < hn; ++k) {
                        const int p = perm[qubit_of(gates[j], hs[k])];
                        if (p < nl && next_use[p] == never) next_use[p] = j;
                    }
                }
                int victim = -1;
                for (int p = 0; p < nl; ++p) {
                    if (busy.count(p)) continue;
                    if (victim < 0) { victim = p; continue; }
                    const bool nr_p = !recent.count(p), nr_v = !recent.count(victim);
                    // max over (next_use, not in recent, position)
                    if (next_use[p] != next_use[victim] ? next_use[p] > next_use[victim]
                                                        : (nr_p != nr_v ? nr_p : p > victim))
                        victim = p;
                }
                if (victim < 0) throw std::runtime_error("no local qubit left to exchange with");
                const int gpos = perm[lq];
                ShardStepRec sw;
                sw.is_swap = true;
                sw.global_qubit = gpos;
                sw.local_qubit = victim;
                plan.steps.push_back(sw);
                const int other = inv[victim];
                perm[lq] = victim;
                perm[other] = gpos;
            }
        }
        qsim_gate_t rec = g;
        rec.q0 = g.q0 >= 0 ? perm[g.q0] : -1;
        rec.q1 = g.q1 >= 0 ? perm[g.q1] : -1;
        rec.q2 = g.q2 >= 0 ? perm[g.q2] : -1;
        cur.push_back(rec);
    }
    flush();
    plan.perm = perm;
    return plan;
}

}  // namespace b200
}  // namespace qsim
