/* qsim_b200.h — C ABI of the B200-native state-vector engine.
 *
 * This is the drop-in boundary for the gate-application hot path of
 * rylanmalarchick/cuda-quantum-simulator.  The reference has no FFI layer: its boundary is the
 * C++ class surface of static library `quantum_sim_lib` (reference CMakeLists.txt:48-61).  The
 * same C++ surface is provided by include/qsim/ *.hpp (namespace qsim) on top of this library;
 * the functions below are what a foreign-language binding (ctypes / cgo / JNI) of that surface
 * would call.  Each entry cites the reference interface it stands in for (paths relative to the
 * reference repository).
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are caller-owned unless stated otherwise;
 *   - amplitudes are interleaved (re, im) IEEE doubles == cuDoubleComplex == std::complex<double>;
 *   - qubit q <-> bit q of the amplitude index (reference src/Gates.cu:19-25);
 *   - every function returns a qsim_status_t; the text of the last error of the calling thread
 *     is available from qsim_last_error().  The codes mirror the C++ exception the reference
 *     throws in the same situation (reference include/Constants.hpp:83-100, src/Circuit.cpp:16-56);
 *   - there is NO CPU fallback: functions that need the GPU return QSIM_ERR_RUNTIME when no
 *     device is present.
 */
#ifndef QSIM_B200_H
#define QSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define QSIM_API __attribute__((visibility("default")))
#else
#define QSIM_API
#endif

typedef enum {
    QSIM_OK = 0,
    QSIM_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument */
    QSIM_ERR_OUT_OF_RANGE = 2,     /* std::out_of_range     */
    QSIM_ERR_RUNTIME = 3           /* std::runtime_error (CUDA errors, unknown gate, zero-probability outcome) */
} qsim_status_t;

/* Gate types, same order and integer values as `enum class GateType`
 * (reference include/Circuit.hpp:42-59). */
enum {
    QSIM_GATE_X = 0, QSIM_GATE_Y, QSIM_GATE_Z, QSIM_GATE_H, QSIM_GATE_S, QSIM_GATE_T,
    QSIM_GATE_SDAG, QSIM_GATE_TDAG, QSIM_GATE_RX, QSIM_GATE_RY, QSIM_GATE_RZ,
    QSIM_GATE_CNOT, QSIM_GATE_CZ, QSIM_GATE_CRY, QSIM_GATE_CRZ, QSIM_GATE_SWAP, QSIM_GATE_TOFFOLI
};

/* One gate == `struct GateOp` (reference include/Circuit.hpp:64-84):
 * qubits are [target], [control, target] / [q1, q2], or [c1, c2, target]; unused slots are -1. */
typedef struct {
    int32_t type;
    int32_t q0, q1, q2;
    double param;
} qsim_gate_t;

/* Noise types, same order as `enum class NoiseType` (reference include/NoiseModel.cuh:46-53). */
enum {
    QSIM_NOISE_DEPOLARIZING = 0, QSIM_NOISE_AMPLITUDE_DAMPING, QSIM_NOISE_PHASE_DAMPING,
    QSIM_NOISE_BIT_FLIP, QSIM_NOISE_PHASE_FLIP, QSIM_NOISE_BIT_PHASE_FLIP
};

/* One noise channel == `struct NoiseChannel` (reference include/NoiseModel.cuh:58-66) flattened:
 * n_qubits == 0 means "all qubits" (reference include/NoiseModel.cuh:118-122). */
typedef struct {
    int32_t type;
    int32_t n_qubits;
    const int32_t* qubits;
    double probability;
} qsim_noise_channel_t;

typedef struct qsim_sim qsim_sim_t;           /* Simulator / StateVector      */
typedef struct qsim_program qsim_program_t;   /* a compiled (fused) circuit   */
typedef struct qsim_batched qsim_batched_t;   /* BatchedSimulator             */
typedef struct qsim_noisy qsim_noisy_t;       /* NoisySimulator               */
typedef struct qsim_dm qsim_dm_t;             /* DensityMatrixSimulator       */
typedef struct qsim_sharded qsim_sharded_t;   /* ShardedSimulator (multi-GPU) */
typedef struct qsim_sharded_plan qsim_sharded_plan_t;   /* a circuit planned + compiled for a sharded state */

QSIM_API const char* qsim_last_error(void);
QSIM_API const char* qsim_version(void);
QSIM_API int qsim_max_qubits(void);           /* cuda_config::MAX_QUBITS, raised from 30 */

/* ---- Circuit (reference include/Circuit.hpp:89-144, src/Circuit.cpp) -------------------------- */
/* Validation exactly as the fluent builder does it, gate by gate: out_of_range for a bad qubit
 * index (src/Circuit.cpp:26-31), invalid_argument for duplicate qubits (:33-48), non-finite angle
 * (:50-56) or a qubit count outside [1, MAX] (:16-24). */
QSIM_API qsim_status_t qsim_circuit_validate(int num_qubits, const qsim_gate_t* gates, int64_t n_gates);
/* createRandomCircuit(n, depth, seed) (src/Circuit.cpp:252-282): writes `depth` gates. */
QSIM_API qsim_status_t qsim_circuit_random(int num_qubits, int depth, unsigned seed, qsim_gate_t* out);
/* createGHZCircuit(n) (src/Circuit.cpp:240-250): writes n gates. */
QSIM_API qsim_status_t qsim_circuit_ghz(int num_qubits, qsim_gate_t* out);
/* Circuit::getDepth (src/Circuit.cpp:165-182). */
QSIM_API qsim_status_t qsim_circuit_depth(int num_qubits, const qsim_gate_t* gates, int64_t n_gates, int64_t* depth);

/* ---- Compiled circuits (new: the fusion pass; no reference counterpart) ----------------------- */
/* n_global > 0 compiles for one shard of a state whose top n_global qubits are the rank id.    */
QSIM_API qsim_status_t qsim_program_compile(int num_qubits, int n_global, const qsim_gate_t* gates,
                                            int64_t n_gates, qsim_program_t** out);
/* Same with an inherited X frame: bit q of initial_xor set means "an X on physical qubit q is still
 * pending" (the sharded driver threads the frame of the global qubits from one segment to the next). */
QSIM_API qsim_status_t qsim_program_compile_ex(int num_qubits, int n_global, const qsim_gate_t* gates,
                                               int64_t n_gates, uint64_t initial_xor, qsim_program_t** out);
/* The same with one more hint for a sharded driver: a pass whose HIGHEST tile qubit is local qubit `isolate_qubit` moves that
 * qubit's two halves of every tile with TMA instructions of their own, so that one half can be loaded from another GPU (the
 * program after a qubit exchange, second half of a split exchange: qsim_shard_execute_exchange_half).  -1: no hint. */
QSIM_API qsim_status_t qsim_program_compile_ex2(int n, int n_global, const qsim_gate_t* gates, int64_t n_gates, uint64_t initial_xor,
                                                int isolate_qubit, qsim_program_t** out);
QSIM_API void qsim_program_destroy(qsim_program_t* p);
/* info[0]=passes, [1]=ops after merging, [2]=gates, [3]=sweeps (total), [4]=tile bits of pass 0,
 * [5]=local qubits, [6]=X frame left on the global qubits (bit q - n_local): the caller owns it */
QSIM_API qsim_status_t qsim_program_info(const qsim_program_t* p, int64_t info[8]);
/* ---- Run-time specialised pass kernels (additive; the reference launches one fixed kernel per gate,
 * src/Simulator.cu:48-154).  For large states every pass gets a kernel generated from its op list and compiled with
 * NVRTC for sm_100a, cached by the STRUCTURE of the pass (matrix entries stay run-time data).
 * mode: 0 = never (always the interpreter kernel), 1 = passes over >= min_qubits local qubits (default, 26),
 * 2 = every pass and a failed compile is an error.  min_qubits <= 0 keeps the current threshold.
 * Environment: QSIM_JIT=off|auto|always, QSIM_JIT_MIN_QUBITS. */
QSIM_API qsim_status_t qsim_jit_set_mode(int mode, int min_qubits);
/* Specialised kernels come in two builds.  One warp group: all 16 warps sweep one tile together, two tiles in flight (HBM-bound
 * passes).  TWO warp groups: 8 + 8 warps on two tiles at once, each thread doing two threads' work in turn, so that one
 * group's shared-memory / shuffle / barrier phases overlap the other's FP64 phases (compute-heavy passes: many ops per tile).
 * mode: 0 = never, 1 = passes whose estimated FP64 instructions per tile and thread reach min_fp64 (default, 350),
 * 2 = every pass that can (full 12-qubit tiles, >= 148 tiles); -1 / min_fp64 < 0 keep the current value.
 * Environment: QSIM_DUAL=off|auto|always, QSIM_DUAL_MIN_FP64. */
QSIM_API qsim_status_t qsim_jit_set_dual(int mode, int min_fp64);
/* In mode 1 a run never waits for the compiler: the first launches of a new pass structure use the ahead-of-time kernel
 * while a background thread compiles (QSIM_JIT_ASYNC=0: compile in the calling thread instead).  qsim_jit_wait blocks until
 * every queued compile has finished (benchmarks, or before a long run). */
QSIM_API qsim_status_t qsim_jit_wait(void);
/* Call before the process exits if background compiles may still be running (the Python mirror registers it with atexit): queued
 * compiles are dropped, compiles in progress get to finish, none start after it.  A process must not run into NVRTC's own exit
 * handlers while a background thread is inside the compiler.  The library also registers it with atexit() itself (best effort:
 * handlers NVRTC registers later still run first). */
QSIM_API qsim_status_t qsim_jit_shutdown(void);
/* Queues the background compile of one pass of a compiled program, or finds its kernel ready: *state = 0 ready, 1 compiling,
 * 2 unavailable.  Needs no GPU (pre-warming the on-disk cache; the GPU-less test of the background machinery). */
QSIM_API qsim_status_t qsim_program_jit_request(const qsim_program_t* p, int pass, int* state);
/* out[0]=kernels compiled, [1]=cache hits, [2]=specialised launches, [3]=failed compiles, [4]=compile time (us),
 * [5]=kernels loaded from the on-disk cache ($QSIM_JIT_CACHE, default ~/.cache/qsim_b200/jit, "off" disables), [6]=mode,
 * [7]=min_qubits */
QSIM_API qsim_status_t qsim_jit_stats(int64_t out[8]);
/* Marks a compiled program for specialisation whatever the state size: a pre-compiled circuit that will run many times is
 * worth one NVRTC compile per pass even below min_qubits (ignored in mode 0).  Takes effect at the next execute. */
QSIM_API qsim_status_t qsim_program_set_specialised(qsim_program_t* p, int on);
/* CUDA C++ generated for pass `pass` of a compiled program (whole_unit != 0: the full translation unit handed to
 * NVRTC).  Returns bytes needed incl. NUL; needs no GPU. */
QSIM_API size_t qsim_program_jit_source(const qsim_program_t* p, int pass, int whole_unit, char* buf, size_t cap);
/* Compiles that pass's kernel to an sm_100a cubin without loading it (needs NVRTC, no GPU): the build check of the
 * generated code.  cubin_out (optional, cap bytes) receives the image. */
QSIM_API qsim_status_t qsim_program_jit_compile(const qsim_program_t* p, int pass, int64_t* cubin_bytes, void* cubin_out,
                                                size_t cap);
/* Human-readable plan (passes, tile qubits, sweeps).  Returns bytes needed incl. NUL. */
QSIM_API size_t qsim_program_describe(const qsim_program_t* p, char* buf, size_t cap);

/* ---- Simulator / StateVector (reference include/Simulator.hpp:53-85, include/StateVector.cuh:66-124) */
QSIM_API qsim_status_t qsim_sim_create(int num_qubits, qsim_sim_t** out);            /* Simulator(int) */
/* A simulator on caller-owned device memory of 16 << num_qubits bytes (e.g. a torch tensor).  Unlike Simulator(int)
 * the memory is ADOPTED AS IS - its contents are the state, nothing is initialised (so an existing state vector can be
 * handed over); call qsim_sim_reset for |0...0>. */
QSIM_API qsim_status_t qsim_sim_create_external(int num_qubits, void* device_state, qsim_sim_t** out);
QSIM_API void qsim_sim_destroy(qsim_sim_t* s);
QSIM_API qsim_status_t qsim_sim_set_stream(qsim_sim_t* s, void* cuda_stream);
/* On library-owned memory (and on shards) reset / init_basis only RECORD the basis state: the first pass of the next
 * run generates it on chip, and any other access (device_ptr, get_state, probabilities, sample, measure) writes it out
 * first.  Observable behaviour is the reference's; on caller-owned memory of a plain simulator the write is immediate. */
QSIM_API qsim_status_t qsim_sim_reset(qsim_sim_t* s);                                /* Simulator::reset */
QSIM_API qsim_status_t qsim_sim_init_basis(qsim_sim_t* s, uint64_t basis_index);     /* StateVector::initializeBasis */
QSIM_API qsim_status_t qsim_sim_set_state(qsim_sim_t* s, const double* host_amplitudes);
QSIM_API qsim_status_t qsim_sim_run(qsim_sim_t* s, int circuit_qubits, const qsim_gate_t* gates,
                                    int64_t n_gates);                                /* Simulator::run */
QSIM_API qsim_status_t qsim_sim_apply_gate(qsim_sim_t* s, const qsim_gate_t* gate);  /* Simulator::applyGate */
QSIM_API qsim_status_t qsim_sim_execute(qsim_sim_t* s, const qsim_program_t* p);     /* pre-compiled run */
QSIM_API qsim_status_t qsim_sim_synchronize(qsim_sim_t* s);
QSIM_API qsim_status_t qsim_sim_get_state(const qsim_sim_t* s, double* host_out);    /* getStateVector */
QSIM_API qsim_status_t qsim_sim_get_probabilities(const qsim_sim_t* s, double* host_out); /* getProbabilities */
/* Device-side read-out that never materialises 2^n values: probabilities of
 * `count` indices starting at `first`. */
QSIM_API qsim_status_t qsim_sim_get_probability_range(const qsim_sim_t* s, uint64_t first, uint64_t count,
                                                      double* host_out);
QSIM_API qsim_status_t qsim_sim_total_probability(const qsim_sim_t* s, double* out); /* getTotalProbability */
/* Marginal distribution of k <= 12 qubits (qubits[i] -> bit i of the outcome), 2^k doubles: what the reference's callers
 * get by summing getProbabilities() on the host (src/main.cpp:30-41), computed on the device without a 2^n vector. */
QSIM_API qsim_status_t qsim_sim_marginal(const qsim_sim_t* s, const int* qubits, int k, double* host_out);
/* Simulator::sample with caller-supplied uniforms in [0,1): out[i] = smallest index
 * whose sequential fp64 CDF value is >= uniforms[i] — bit-identical to std::partial_sum +
 * std::lower_bound (reference src/Simulator.cu:164-185).  64-bit indices. */
QSIM_API qsim_status_t qsim_sim_sample_uniforms(qsim_sim_t* s, const double* uniforms, int64_t n_shots,
                                                int64_t* out);
/* Same with uniforms drawn as the reference does: std::mt19937(seed) +
 * std::uniform_real_distribution<double>(0,1), one draw per shot. */
QSIM_API qsim_status_t qsim_sim_sample_seeded(qsim_sim_t* s, unsigned seed, int64_t n_shots, int64_t* out);
/* Simulator::measureQubit(q) with the uniform draw r supplied: measures index bit n-1-q exactly as
 * StateVector::measure does (reference src/StateVector.cu:87-89, 260-314). */
QSIM_API qsim_status_t qsim_sim_measure(qsim_sim_t* s, int qubit, double r, int* outcome);
/* Measurement of index bit `bit` (the NoisySimulator convention, reference src/NoiseModel.cu:615-651). */
QSIM_API qsim_status_t qsim_sim_measure_bit(qsim_sim_t* s, int bit, double r, int* outcome, double* p0);
QSIM_API int qsim_sim_num_qubits(const qsim_sim_t* s);
QSIM_API void* qsim_sim_device_ptr(qsim_sim_t* s);                                   /* StateVector::devicePtr */
/* Kernel launches issued by this simulator so far (bench.py's gpu_launches). */
QSIM_API int64_t qsim_sim_launch_count(const qsim_sim_t* s);
/* Average device time of the fused-pass kernel since the last call (CUDA events on the
 * simulator's stream), and how many passes were timed.  Enable with qsim_sim_set_timing. */
QSIM_API qsim_status_t qsim_sim_set_timing(qsim_sim_t* s, int enabled);
QSIM_API qsim_status_t qsim_sim_pass_time_ms(qsim_sim_t* s, double* total_ms, int64_t* n_passes);
/* Same, one duration per timed pass in launch order (at most `cap` written; *n_out = how many there were). */
QSIM_API qsim_status_t qsim_sim_pass_times(qsim_sim_t* s, double* out_ms, int64_t cap, int64_t* n_out);
/* Development aid, only with QSIM_PASS_TIMELINE=1 in the environment when the simulator is created: eight %globaltimer
 * stamps (ns) of CTA 0 for each of the last <= 64 pass launches - kernel entry, set-up done, first loads issued, first tile
 * arrived, first tile computed, its store issued, all tiles computed, all stores complete.  *n_out = values available. */
QSIM_API qsim_status_t qsim_sim_pass_timeline(qsim_sim_t* s, uint64_t* out, int64_t cap, int64_t* n_out);

/* ---- Sharded state: one process per GPU, top n_global qubits = rank -------------- */
/* The shard holds 2^(num_qubits - n_global) amplitudes; `rank` supplies the values of the global
 * qubits.  device_state may be NULL (library allocates). */
QSIM_API qsim_status_t qsim_shard_create(int num_qubits, int n_global, int rank, void* device_state,
                                         qsim_sim_t** out);
/* In-place half-shard exchange with the peer rank (rank ^ (1 << (global_qubit - n_local))):
 * swaps global qubit `global_qubit` with local qubit `local_qubit`.  `peer_state` is the peer's
 * shard mapped into this process (CUDA IPC / peer access).  Both ranks call it; each moves half of
 * the pairs.  Callers synchronise ranks before and after. */
QSIM_API qsim_status_t qsim_shard_swap_p2p(qsim_sim_t* s, void* peer_state, int global_qubit, int local_qubit);
/* The same exchange FUSED into the last pass of program `p` (one kernel does the gate pass and the NVLink transfer,
 * tile by tile): the pass stores out of place — amplitudes that stay go to `alt_state` (a second buffer of this
 * rank, same size as the shard), amplitudes that leave go straight into `peer_alt_state` (the partner's second
 * buffer, peer-mapped) — and the simulator continues on `alt_state`.  Nothing is written to the buffers any rank is
 * still reading, so no rank synchronisation is needed BEFORE the call; synchronise ranks after it.  Fails with
 * QSIM_ERR_INVALID_ARGUMENT, having done nothing, when the program has no pass or `local_qubit` is a tile qubit of
 * its last pass (qsim_program_last_tile_mask): run qsim_sim_execute + qsim_shard_swap_p2p instead. */
QSIM_API qsim_status_t qsim_shard_execute_exchange(qsim_sim_t* s, const qsim_program_t* p, void* alt_state,
                                                   void* peer_alt_state, int global_qubit, int local_qubit);
/* The fused exchange IN PLACE, for shards that leave no room for a second buffer (36 qubits on 8 x 180 GB): amplitudes
 * that stay are stored where they were, amplitudes that leave go straight over the partner's own leaving amplitudes in
 * `peer_state` (the partner's LIVE buffer, peer-mapped).  Both GPUs run the same grid over the same tile order, and a
 * tile is stored over the partner's tile only after the partner's kernel has loaded that tile: each sender CTA publishes
 * `hs_base` + its count of loaded tiles in the partner's handshake words (`hs_peer`, 1024 x u64, peer-mapped) and polls
 * its own (`hs_local`).  `hs_base` must grow by at least 2^32 from one exchange to the next on both ranks (the words are
 * never reset).  Ranks must be synchronised BEFORE the call (the partner's previous kernels have finished) and after it.
 * A poll that sees no progress for `timeout_ns` (0: 10 s) sets *hs_error_dev (device int, caller-zeroed) and the kernel
 * runs to its end without waiting any more: the state is then invalid, nothing hangs.  Needs
 * qsim_shard_inplace_exchange_possible; a program whose first pass would start from a lazily reset shard is written out first. */
QSIM_API qsim_status_t qsim_shard_execute_exchange_inplace(qsim_sim_t* s, const qsim_program_t* p, void* peer_state,
                                                           int global_qubit, int local_qubit, void* hs_local, void* hs_peer,
                                                           uint64_t hs_base, uint64_t timeout_ns, int* hs_error_dev);
/* The in-place exchange SPLIT over the two passes around it, so that each pass hides half of the NVLink time: another index
 * bit w that is a tile qubit of neither pass halves the leaving amplitudes.  which = 1: the LAST pass of `p` (the program before
 * the exchange) scatters the half with bit w clear into the partner's shard, like qsim_shard_execute_exchange_inplace;
 * synchronise ranks; which = 2: the FIRST pass of `p` (the program after the exchange) loads the half with bit w set from
 * `peer_state` (at index ^ (1 << local_qubit)) instead of from its own shard and stores it in place, a store over a tile
 * waiting until the partner has loaded that tile (same handshake words, a later epoch); synchronise ranks again.
 * qsim_shard_split_exchange_possible: *split_bit_out = a usable w, or -1 (then use the unsplit call). */
QSIM_API qsim_status_t qsim_shard_execute_exchange_half(qsim_sim_t* s, const qsim_program_t* p, void* peer_state, int global_qubit,
                                                        int local_qubit, int split_bit, int which, void* hs_local, void* hs_peer,
                                                        uint64_t hs_base, uint64_t timeout_ns, int* hs_error_dev);
QSIM_API qsim_status_t qsim_shard_split_exchange_possible(qsim_sim_t* s, const qsim_program_t* before, const qsim_program_t* after,
                                                          int local_qubit, int* split_bit_out);
/* *possible_out = 1 when the last pass of `p` can carry the exchange of `local_qubit` in place: the qubit is not one of its
 * tile qubits, the shard has at least 16 tiles, and no deferred X gate pairs tiles across that qubit. */
QSIM_API qsim_status_t qsim_shard_inplace_exchange_possible(qsim_sim_t* s, const qsim_program_t* p, int local_qubit,
                                                            int* possible_out);
/* Bit q set: local qubit q is a tile qubit of the program's last pass (0 when the program has no pass). */
QSIM_API qsim_status_t qsim_program_last_tile_mask(const qsim_program_t* p, uint64_t* mask_out);
/* Bounce-buffer variant for NCCL send/recv: pack the half shard that must leave into `buf`
 * (chunk `chunk` of `n_chunks`), and unpack a received chunk. */
QSIM_API qsim_status_t qsim_shard_pack_half(qsim_sim_t* s, int local_qubit, int keep_bit, int64_t chunk,
                                            int64_t n_chunks, void* buf);
QSIM_API qsim_status_t qsim_shard_unpack_half(qsim_sim_t* s, int local_qubit, int keep_bit, int64_t chunk,
                                              int64_t n_chunks, const void* buf);
/* CUDA IPC plumbing so Python can exchange handles over torch.distributed. */
/* handle_out identifies the allocation that contains device_ptr; *offset_out is device_ptr's byte
 * offset inside it (a torch tensor need not start its cudaMalloc block). */
QSIM_API qsim_status_t qsim_ipc_get_handle(void* device_ptr, unsigned char handle_out[64], uint64_t* offset_out);
/* Maps the peer allocation; returns its BASE pointer (add the peer's offset yourself). */
QSIM_API qsim_status_t qsim_ipc_open_handle(const unsigned char handle[64], void** base_ptr_out);
QSIM_API qsim_status_t qsim_ipc_close_handle(void* base_ptr);
/* Partial sums for distributed read-out: this shard's sum of |a|^2 (optionally restricted to
 * index bit `bit` == 0; bit < 0 means no restriction). */
QSIM_API qsim_status_t qsim_shard_partial_probability(const qsim_sim_t* s, int bit, double* out);
/* Collapse after a distributed measurement: amplitudes whose index bit `bit` != outcome become 0, the others are
 * multiplied by `scale` (1/sqrt(P(outcome)), reference src/StateVector.cu:110-124).  bit < 0: every amplitude is scaled
 * (the measured qubit is a rank bit: scale = 1/sqrt(P) on the ranks that hold the outcome, 0 on the others). */
QSIM_API qsim_status_t qsim_shard_collapse(qsim_sim_t* s, int bit, int outcome, double scale);
/* Distributed bit-exact sampling: the sequential CDF of this shard continued from c_init (the exact
 * running sum at the end of the previous shard, in logical rank order).  out[i] = local index of the
 * first amplitude whose CDF value is >= uniforms[i] if that index lies in this shard, else -1;
 * *c_end = exact running sum at the end of this shard.  `first_shard` != 0 makes r <= c_init (r == 0)
 * resolve to local index 0 as std::lower_bound does. */
QSIM_API qsim_status_t qsim_shard_sample(qsim_sim_t* s, double c_init, int first_shard, const double* uniforms,
                                         int64_t n_shots, int64_t* out, double* c_end);
/* Optional two-step preamble that lets all shards do the expensive part at the same time: qsim_shard_cdf_prepare sweeps
 * this shard's amplitudes (chunk sums and candidate increments) and returns its approximate total;
 * qsim_shard_cdf_classify takes the approximate sum of the shards before this one; the following qsim_shard_sample then
 * only stitches the chunk starts from the exact c_init and samples.  No other read-out call in between. */
QSIM_API qsim_status_t qsim_shard_cdf_prepare(qsim_sim_t* s, double* approx_total);
QSIM_API qsim_status_t qsim_shard_cdf_classify(qsim_sim_t* s, double approx_c_init);

/* ---- ShardedSimulator: the multi-GPU Simulator, driven from C++ over NCCL (include/qsim/sharded_simulator.hpp; additive,
 * the reference is single-GPU: README.md:367; surface after include/Simulator.hpp:53-85).  One process per GPU; all ranks make
 * the same calls in the same order (the calls contain collectives).  Bootstrap: rank 0 obtains a unique id and hands the 128
 * bytes to every rank (any channel: MPI, a file, torch.distributed in the Python mirror), then every rank calls create on its
 * current CUDA device.  exchange: 0 = auto (CUDA-IPC peer memory, else NCCL send/recv), 1 = peer memory, 2 = NCCL. */
QSIM_API qsim_status_t qsim_sharded_unique_id(unsigned char out[128]);
QSIM_API qsim_status_t qsim_sharded_create(int num_qubits, int rank, int world_size, const unsigned char unique_id[128],
                                           int exchange, qsim_sharded_t** out);
QSIM_API void qsim_sharded_destroy(qsim_sharded_t* h);
QSIM_API qsim_status_t qsim_sharded_reset(qsim_sharded_t* h);                                           /* Simulator::reset */
QSIM_API qsim_status_t qsim_sharded_run(qsim_sharded_t* h, int circuit_qubits, const qsim_gate_t* gates, int64_t n_gates);
/* Plans + compiles the circuit against the current qubit layout: out[0..k) receive plans for k consecutive runs (run i is
 * compiled against the layout run i-1 leaves; plans are shared when the layout repeats; destroy each handle). */
QSIM_API qsim_status_t qsim_sharded_compile(qsim_sharded_t* h, int circuit_qubits, const qsim_gate_t* gates, int64_t n_gates,
                                            int k, qsim_sharded_plan_t** out);
QSIM_API void qsim_sharded_plan_destroy(qsim_sharded_plan_t* p);
/* info[0]=passes, [1]=ops, [2]=global<->local swaps */
QSIM_API qsim_status_t qsim_sharded_plan_info(const qsim_sharded_plan_t* p, int64_t info[8]);
/* Fails with QSIM_ERR_INVALID_ARGUMENT when the state's layout is not the one the plan was compiled against. */
QSIM_API qsim_status_t qsim_sharded_execute(qsim_sharded_t* h, const qsim_sharded_plan_t* p);
/* LOGICAL basis-state indices in the reference's sequential-CDF order (src/Simulator.cu:164-185), identical on every rank. */
QSIM_API qsim_status_t qsim_sharded_sample(qsim_sharded_t* h, const double* uniforms, int64_t n_shots, int64_t* out);
/* Simulator::measureQubit: index bit n-1-qubit (src/StateVector.cu:87-89); outcome 0 iff uniform < P(bit = 0). */
QSIM_API qsim_status_t qsim_sharded_measure(qsim_sharded_t* h, int qubit, double uniform, int* result);
QSIM_API qsim_status_t qsim_sharded_measure_bit(qsim_sharded_t* h, int bit, double uniform, int* result, double* p0);
QSIM_API qsim_status_t qsim_sharded_marginal(qsim_sharded_t* h, const int* qubits, int n_qubits, double* out);
QSIM_API qsim_status_t qsim_sharded_total_probability(qsim_sharded_t* h, double* out);
/* This rank's shard in the STORED layout (see qsim_sharded_layout), 2^(local qubits) amplitudes. */
QSIM_API qsim_status_t qsim_sharded_get_local_state(qsim_sharded_t* h, double* out_amplitudes);
QSIM_API qsim_status_t qsim_sharded_set_local_state(qsim_sharded_t* h, const double* amplitudes);
/* Moves the amplitudes back to the identity qubit layout (exchanges + one local permutation pass). */
QSIM_API qsim_status_t qsim_sharded_restore_identity_layout(qsim_sharded_t* h);
/* perm_out[q] = physical index bit that holds logical qubit q; frame_out = pending X mask over physical bits (rank bits only). */
QSIM_API qsim_status_t qsim_sharded_layout(const qsim_sharded_t* h, int* perm_out, uint64_t* frame_out);
QSIM_API qsim_status_t qsim_sharded_set_identity_layout_only(qsim_sharded_t* h, int on);   /* never choose a layout for |0..0> */
/* Declares the stored layout to be the identity without moving data: a relabelling of the logical qubits (the state changes
 * by a qubit permutation).  For benchmarks that need every step to start from the same layout. */
QSIM_API qsim_status_t qsim_sharded_relabel_identity(qsim_sharded_t* h);
QSIM_API qsim_status_t qsim_sharded_swap(qsim_sharded_t* h, int global_position, int local_position);   /* one separate exchange */
/* info[0]=local qubits, [1]=rank qubits, [2]=exchanges fused into a pass, [3]=separate exchanges, [4]=1 peer memory / 2 NCCL,
 * [5]=rank, [6]=world size, [7]=1 if the second shard buffer of the fused exchange exists */
QSIM_API qsim_status_t qsim_sharded_info(const qsim_sharded_t* h, int64_t info[8]);
/* Of info[2]: the exchanges that ran in place (no second buffer; qsim_shard_execute_exchange_inplace). */
QSIM_API qsim_status_t qsim_sharded_inplace_exchanges(const qsim_sharded_t* h, int64_t* count_out);
/* Of those: the exchanges split over the pass before and the pass after (qsim_shard_execute_exchange_half). */
QSIM_API qsim_status_t qsim_sharded_split_exchanges(const qsim_sharded_t* h, int64_t* count_out);
QSIM_API qsim_sim_t* qsim_sharded_local(qsim_sharded_t* h);   /* the shard as a qsim_sim_t (timing, launch counters); borrowed */
QSIM_API qsim_status_t qsim_sharded_set_stream(qsim_sharded_t* h, void* cuda_stream);
QSIM_API qsim_status_t qsim_sharded_synchronize(qsim_sharded_t* h);
QSIM_API qsim_status_t qsim_sharded_barrier(qsim_sharded_t* h);   /* stream-ordered barrier across the ranks */
/* The planner alone (host logic, no device, no NCCL): steps_out[3*i] = 0 (gates: [3*i+1] = how many, taken in order from
 * gates_out, on PHYSICAL positions) or 1 (swap of global position [3*i+1] with local position [3*i+2]).  perm_in may be NULL
 * (identity); choose_layout != 0 picks the layout for a run from |0...0> first (returned in perm_start_out). */
QSIM_API qsim_status_t qsim_sharded_plan_circuit(int num_qubits, int n_global, const qsim_gate_t* gates, int64_t n_gates,
                                                 const int* perm_in, int choose_layout, int64_t* steps_out, int64_t cap_steps,
                                                 qsim_gate_t* gates_out, int* perm_start_out, int* perm_end_out,
                                                 int64_t* n_steps_out);

/* ---- NoisySimulator (reference include/NoiseModel.cuh:141-225) ---------------------------------- */
QSIM_API qsim_status_t qsim_noisy_create(int num_qubits, const qsim_noise_channel_t* channels, int n_channels,
                                         qsim_noisy_t** out);
QSIM_API void qsim_noisy_destroy(qsim_noisy_t* s);
QSIM_API qsim_status_t qsim_noisy_set_noise(qsim_noisy_t* s, const qsim_noise_channel_t* channels, int n_channels);
QSIM_API qsim_status_t qsim_noisy_set_seed(qsim_noisy_t* s, unsigned seed);                 /* setSeed */
QSIM_API qsim_status_t qsim_noisy_reset(qsim_noisy_t* s);
QSIM_API qsim_status_t qsim_noisy_run(qsim_noisy_t* s, int circuit_qubits, const qsim_gate_t* gates, int64_t n_gates);
QSIM_API qsim_status_t qsim_noisy_apply_gate(qsim_noisy_t* s, const qsim_gate_t* gate);
QSIM_API qsim_status_t qsim_noisy_apply_noise(qsim_noisy_t* s, const qsim_noise_channel_t* channel); /* applyNoise */
QSIM_API qsim_status_t qsim_noisy_get_state(const qsim_noisy_t* s, double* host_out);
QSIM_API qsim_status_t qsim_noisy_get_probabilities(const qsim_noisy_t* s, double* host_out);
/* sample(n_shots): one mt19937 draw per shot from the simulator's own engine (seeded by set_seed). */
QSIM_API qsim_status_t qsim_noisy_sample(qsim_noisy_t* s, int n_shots, int32_t* out);
/* measureQubit(q): index bit q, one draw (reference src/NoiseModel.cu:615-651). */
QSIM_API qsim_status_t qsim_noisy_measure(qsim_noisy_t* s, int qubit, int* outcome);

/* ---- BatchedSimulator (reference include/NoiseModel.cuh:236-297) --------------------------------- */
QSIM_API qsim_status_t qsim_batched_create(int num_qubits, int batch_size, const qsim_noise_channel_t* channels,
                                           int n_channels, qsim_batched_t** out);
QSIM_API void qsim_batched_destroy(qsim_batched_t* s);
QSIM_API qsim_status_t qsim_batched_set_noise(qsim_batched_t* s, const qsim_noise_channel_t* channels, int n_channels);
QSIM_API qsim_status_t qsim_batched_set_seed(qsim_batched_t* s, unsigned seed);
QSIM_API qsim_status_t qsim_batched_reset(qsim_batched_t* s);
QSIM_API qsim_status_t qsim_batched_run(qsim_batched_t* s, int circuit_qubits, const qsim_gate_t* gates, int64_t n_gates);
QSIM_API qsim_status_t qsim_batched_average_probabilities(const qsim_batched_t* s, double* host_out);   /* 2^n */
QSIM_API qsim_status_t qsim_batched_get_probabilities(const qsim_batched_t* s, int trajectory, double* host_out);
QSIM_API qsim_status_t qsim_batched_get_state(const qsim_batched_t* s, int trajectory, double* host_out);
/* sample(n_shots): out[shot * batch + trajectory]; draws are trajectory-major (src/NoiseModel.cu:938-957). */
QSIM_API qsim_status_t qsim_batched_sample(qsim_batched_t* s, int n_shots, int32_t* out);
QSIM_API qsim_status_t qsim_batched_histogram(qsim_batched_t* s, int n_shots, int32_t* out);             /* 2^n */
QSIM_API size_t qsim_batched_total_memory_bytes(const qsim_batched_t* s);

/* ---- DensityMatrixSimulator (reference include/DensityMatrix.cuh:63-224) -------------------------- */
QSIM_API qsim_status_t qsim_dm_create(int num_qubits, const qsim_noise_channel_t* channels, int n_channels, qsim_dm_t** out);
QSIM_API void qsim_dm_destroy(qsim_dm_t* s);
QSIM_API qsim_status_t qsim_dm_reset(qsim_dm_t* s);
QSIM_API qsim_status_t qsim_dm_run(qsim_dm_t* s, int circuit_qubits, const qsim_gate_t* gates, int64_t n_gates);
QSIM_API qsim_status_t qsim_dm_apply_gate(qsim_dm_t* s, const qsim_gate_t* gate);
QSIM_API qsim_status_t qsim_dm_apply_channel(qsim_dm_t* s, int noise_type, int qubit, double probability);
QSIM_API qsim_status_t qsim_dm_init_pure(qsim_dm_t* s, const double* host_state);        /* DensityMatrix::initFromPureState */
QSIM_API qsim_status_t qsim_dm_init_maximally_mixed(qsim_dm_t* s);
QSIM_API qsim_status_t qsim_dm_get_probabilities(const qsim_dm_t* s, double* host_out);  /* 2^n */
QSIM_API qsim_status_t qsim_dm_get_matrix(const qsim_dm_t* s, double* host_out);         /* 4^n complex, row-major */
QSIM_API qsim_status_t qsim_dm_purity(const qsim_dm_t* s, double* out);
QSIM_API qsim_status_t qsim_dm_trace(const qsim_dm_t* s, double* out);
QSIM_API qsim_status_t qsim_dm_is_valid(const qsim_dm_t* s, double tolerance, int* out);
/* measureQubit with the uniform supplied: outcome = (u < p1) ? 1 : 0 (src/DensityMatrix.cu:374-406). */
QSIM_API qsim_status_t qsim_dm_measure(qsim_dm_t* s, int qubit, double uniform, int* outcome);

#ifdef __cplusplus
}
#endif
#endif /* QSIM_B200_H */
