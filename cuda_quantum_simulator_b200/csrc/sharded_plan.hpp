// Planning for sharded states (internal header, host only).  See sharded_plan.cpp.
#pragma once

#include <cstdint>
#include <vector>

#include "qsim_b200.h"

namespace qsim {
namespace b200 {

struct ShardStepRec {
    bool is_swap = false;
    std::vector<qsim_gate_t> gates;            // !is_swap: gate records on PHYSICAL qubit positions
    int global_qubit = -1, local_qubit = -1;   // is_swap
};

struct ShardPlanRec {
    int n = 0, n_global = 0;
    std::vector<ShardStepRec> steps;
    std::vector<int> perm;                     // logical qubit -> physical position after the plan
    int n_swaps() const {
        int k = 0;
        for (const auto& s : steps) k += s.is_swap ? 1 : 0;
        return k;
    }
};

// Qubit slots (q0/q1/q2) of a gate type that are non-diagonal targets; returns how many (0, 1 or 2).
int shard_target_slots(int gate_type, int slots[2]);
std::vector<int> shard_choose_initial_layout(int n, int n_global, const qsim_gate_t* gates, int64_t ng);
ShardPlanRec shard_plan_circuit(int n, int n_global, const qsim_gate_t* gates, int64_t ng, const std::vector<int>& perm);

}  // namespace b200
}  // namespace qsim
