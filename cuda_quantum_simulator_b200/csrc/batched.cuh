// Batched noisy-trajectory kernels (internal header).  See kernels_batched.cu.
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace qsim {
namespace b200 {

// One step of a trajectory program, executed by every trajectory in order.
struct TrajItem {
    int32_t kind;        // 0 = controlled 2x2 operator, 1 = "all noise events" (after a gate)
    int32_t target;
    uint64_t cmask, cval;
    double m[8];
};

struct TrajEvent {       // one (channel, qubit) noise event; NoiseType order of the qsim API
    int32_t type;
    int32_t qubit;
    double p;
    double sqrt_keep;    // sqrt(1 - p): the no-jump scaling of the damping channels (filled by the host, see make_event)
    double pad;
};
inline TrajEvent make_event(int type, int qubit, double p);

}  // namespace b200
}  // namespace qsim
#include <cmath>
namespace qsim {
namespace b200 {
inline TrajEvent make_event(int type, int qubit, double p) {
    TrajEvent e{};
    e.type = type; e.qubit = qubit; e.p = p; e.sqrt_keep = std::sqrt(1.0 - p); e.pad = 0.0;
    return e;
}

constexpr int kTrajMaxQubits = 13;   // 2^13 amplitudes = 128 KiB of shared memory per trajectory

// Runs `items` on trajectories [0, batch) whose states live at states + traj * 2^n (contiguous, the
// reference's [batch][2^n] layout).  Noise draws: Philox4x32-10, counter = (event, trajectory + traj_offset),
// key = (seed, "QSMB"), event = (index of the noise block << 16) | index in the block.
// d_avg (optional, 2^n doubles): the kernel's epilogue accumulates avg[i] = (1 / batch) * sum_traj |a_traj,i|^2 of the FINAL
// states there (zeroed first) - no separate pass over the batch for BatchedSimulator::getAverageProbabilities.
void launch_trajectories(cuDoubleComplex* states, int n, int64_t batch, const TrajItem* d_items, int n_items,
                         const TrajEvent* d_events, int n_events, uint32_t seed, uint64_t traj_offset,
                         uint64_t first_noise_block, int num_sms, cudaStream_t stream, double* d_avg = nullptr);
void launch_batched_init(cuDoubleComplex* states, int n, int64_t batch, int num_sms, cudaStream_t stream);
// avg[i] = (1 / batch) * sum_traj |a_traj,i|^2
void launch_batched_average(const cuDoubleComplex* states, int n, int64_t batch, double* d_avg, int num_sms,
                            cudaStream_t stream);
// out[shot * batch + traj] = first index whose sequential CDF >= uniforms[traj * n_shots + shot]
// d_out and d_hist are both optional: d_hist[outcome] += 1 per shot (outcomes >= 2^n are dropped, as the reference does)
void launch_batched_sample(const cuDoubleComplex* states, int n, int64_t batch, const double* d_uniforms, int n_shots,
                           int32_t* d_out, int32_t* d_hist, int num_sms, cudaStream_t stream);
// histogram[outcome] += 1 over out[0 .. count) (outcomes >= 2^n are dropped, as the reference does)
void launch_histogram(const int32_t* d_samples, int64_t count, int n, int32_t* d_hist, int num_sms, cudaStream_t stream);

}  // namespace b200
}  // namespace qsim
