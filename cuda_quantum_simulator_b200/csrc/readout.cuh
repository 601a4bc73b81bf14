// Read-out entry points (internal header).  See kernels_readout.cu.
#pragma once

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "engine.hpp"
#include "qsim/cuda_memory.cuh"

namespace qsim {
namespace b200 {

void launch_probabilities(const cuDoubleComplex* state, double* out, uint64_t first, uint64_t count, int num_sms,
                          cudaStream_t stream);
void launch_init_basis(cuDoubleComplex* state, uint64_t n, uint64_t idx, cudaStream_t stream);
// sum of |a_i|^2 over indices whose bit `mask_bit` is 0 (mask_bit < 0: all); deterministic tree order
double reduce_probability(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng);
// Marginal distribution over k index bits (bits[i] -> bit i of the outcome), 2^k doubles to the host: every amplitude is
// read once, sums are formed in a fixed tree order (deterministic), nothing of size 2^n is materialised.
void marginal_probabilities(const cuDoubleComplex* state, int n_bits, const int* bits, int k, double* host_out, Engine& eng);
void launch_collapse(cuDoubleComplex* state, uint64_t n, int bit, int outcome, double scale, int num_sms,
                     cudaStream_t stream);

// Small states: sampling in one launch of one CTA (tree-order scan in shared memory + exact replay inside the rounding
// margin); false if the state is too large for it (the caller uses SequentialCdf).  Host in / host out.
constexpr int kSmallCdfMaxQubits = 14;
bool sample_small_state(const cuDoubleComplex* state, int n_qubits, const double* uniforms_host, int64_t n_shots,
                        int64_t* out_host, Engine& eng);

// Exact sequential-order fp64 prefix sums of the (optionally masked) probabilities, kept as the
// running sum at every 4096-element chunk boundary.
class SequentialCdf {
public:
    // c_init: value the running sum starts from (0 for a whole state; the exact sum of the preceding shards
    // for one shard of a distributed state)
    SequentialCdf(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng, double c_init = 0.0);
    // The same in three steps, for a state sharded over several GPUs: every shard runs the sweep over its amplitudes
    // (prepare) and its classification (classify, from the APPROXIMATE sum of the shards before it) at the same time;
    // only stitch(c_init), which needs the EXACT sum of the shards before it, runs one shard after the other.
    // Nothing else may use the engine's scratch slot 0 between prepare and stitch.
    struct Deferred {};
    SequentialCdf(const cuDoubleComplex* state, uint64_t n, int mask_bit, Engine& eng, Deferred);
    double approxTotal();           // after prepare (the constructor above): tree-order sum of this shard
    void classify(double approx_c_init);
    void stitch(double c_init);
    double total() const;           // == the reference's index-order host sum, bit for bit
    uint64_t slowChunks() const;    // chunks that had to be replayed sequentially (diagnostics)
    // out[i] = smallest index whose CDF value >= uniforms[i] (n if none), host in / host out
    void sample(const double* uniforms_host, int64_t n_shots, int64_t* out_host);
    int launches() const { return launches_; }

private:
    const cuDoubleComplex* state_;
    uint64_t n_;
    int mask_bit_;
    cudaStream_t stream_;
    int chunk_ = 0;
    uint64_t m_ = 0;
    Engine& eng_;
    // slices of the engine's scratch slot 0
    double *approx_ = nullptr, *lo_ = nullptr, *delta_ = nullptr, *base_ = nullptr, *start_ = nullptr;
    unsigned long long* slow_ = nullptr;
    uint8_t* flag_ = nullptr;
    double *cand_delta_ = nullptr, *g_total_ = nullptr, *g_bb_ = nullptr, *g_start_ = nullptr, *g_cand_ = nullptr;
    uint8_t *cand_tie_ = nullptr, *g_kind_ = nullptr, *g_tie_ = nullptr, *g_choice_ = nullptr;
    unsigned int *n_pending_ = nullptr, *pending_ = nullptr;
    uint64_t n_groups_ = 0;
    void setup();
    void prepare();
    int launches_ = 0;
};

}  // namespace b200
}  // namespace qsim
