#include "engine.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"
#include "qsim/constants.hpp"

namespace qsim {
namespace b200 {

void require_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw std::runtime_error(
            "qsim_b200: no CUDA device available - this engine has no CPU fallback (the fused-pass kernels "
            "require an sm_100a GPU)");
    }
}

DeviceProgram::~DeviceProgram() {
    if (graph.exec) cudaGraphExecDestroy(graph.exec);
    if (d_ops) cudaFree(d_ops);
    if (d_tables) cudaFree(d_tables);
    if (d_terms) cudaFree(d_terms);
}

void DeviceProgram::upload() {
    require_device();
    if (d_ops) { cudaFree(d_ops); d_ops = nullptr; }
    if (host.ops.empty()) return;
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_ops), host.ops.size() * sizeof(DevOp)));
    CUDA_CHECK(cudaMemcpy(d_ops, host.ops.data(), host.ops.size() * sizeof(DevOp), cudaMemcpyHostToDevice));
    if (d_tables) { cudaFree(d_tables); d_tables = nullptr; }
    if (d_terms) { cudaFree(d_terms); d_terms = nullptr; }
    if (!host.phase_tables.empty()) {
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_tables), host.phase_tables.size() * sizeof(double)));
        CUDA_CHECK(cudaMemcpy(d_tables, host.phase_tables.data(), host.phase_tables.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (!host.phase_terms.empty()) {
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_terms), host.phase_terms.size() * sizeof(PhaseTerm)));
        CUDA_CHECK(cudaMemcpy(d_terms, host.phase_terms.data(), host.phase_terms.size() * sizeof(PhaseTerm), cudaMemcpyHostToDevice));
    }
}

Engine::Engine() {
    require_device();
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev));
    int major = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        throw std::runtime_error("qsim_b200: kernels are built for sm_100a only (found compute capability " +
                                 std::to_string(major) + ".x)");
    CUDA_CHECK(cudaEventCreateWithFlags(&staged_, cudaEventDisableTiming));
    if (const char* e = std::getenv("QSIM_STAGES")) stages_wanted_ = std::atoi(e);
    if (std::getenv("QSIM_TMA_1D")) use_tensor_map_ = false;
    if (const char* e = std::getenv("QSIM_GRAPH_MAX_QUBITS")) graph_max_qubits_ = std::atoi(e);
    if (std::getenv("QSIM_NO_GRAPH")) graph_max_qubits_ = 0;
    if (std::getenv("QSIM_PASS_TIMELINE")) {
        // pinned host memory the kernels write directly (system-scope stores): still readable after a kernel has faulted
        CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&d_timeline_), (64 * 8 + 1024) * sizeof(unsigned long long), cudaHostAllocMapped));
        std::memset(d_timeline_, 0, (64 * 8 + 1024) * sizeof(unsigned long long));
        graph_max_qubits_ = 0;   // (a replayed graph would stamp the slots it was captured with)
    }
}

std::vector<unsigned long long> Engine::passTimeline() {
    std::vector<unsigned long long> out;
    if (!d_timeline_) return out;
    cudaStreamSynchronize(stream_);   // (no check: the point is to see how far a faulting kernel got)
    const int64_t n = timeline_n_ < 64 ? timeline_n_ : 64;
    std::vector<unsigned long long> all(d_timeline_, d_timeline_ + 64 * 8);
    for (int64_t k = timeline_n_ - n; k < timeline_n_; ++k)
        for (int j = 0; j < 8; ++j) out.push_back(all[(size_t)(k % 64) * 8 + j]);
    if (std::getenv("QSIM_PASS_PROGRESS"))   // per-CTA progress words of the last two-group launch (development aid)
        out.insert(out.end(), d_timeline_ + 64 * 8, d_timeline_ + 64 * 8 + 1024);
    return out;
}

Engine::~Engine() {
    for (void* p : scratch_) if (p) cudaFree(p);
    if (d_aux_) cudaFree(d_aux_);
    if (d_timeline_) cudaFreeHost(d_timeline_);
    if (d_ops_) cudaFree(d_ops_);
    if (h_ops_) cudaFreeHost(h_ops_);
    if (staged_) cudaEventDestroy(staged_);
    if (capture_stream_) cudaStreamDestroy(capture_stream_);
    for (auto& e : events_) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto& e : pool_) cudaEventDestroy(e);
}

void* Engine::scratch(int slot, size_t bytes) {
    if (bytes > scratch_cap_[slot]) {
        if (scratch_[slot]) { CUDA_CHECK(cudaStreamSynchronize(stream_)); cudaFree(scratch_[slot]); scratch_[slot] = nullptr; }
        const size_t cap = bytes + bytes / 4 + 4096;
        CUDA_CHECK(cudaMalloc(&scratch_[slot], cap));
        scratch_cap_[slot] = cap;
    }
    return scratch_[slot];
}

void Engine::synchronize() const { CUDA_CHECK(cudaStreamSynchronize(stream_)); }

cudaEvent_t Engine::getEvent() {
    if (!pool_.empty()) { cudaEvent_t e = pool_.back(); pool_.pop_back(); return e; }
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    return e;
}

void Engine::setTiming(bool on) {
    timing_ = on;
    if (!on) {
        for (auto& e : events_) { pool_.push_back(e.first); pool_.push_back(e.second); }
        events_.clear();
    }
}

void Engine::drainTiming(double* total_ms, int64_t* n_passes, std::vector<double>* each) {
    double tot = 0;
    for (auto& e : events_) {
        CUDA_CHECK(cudaEventSynchronize(e.second));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
        if (each) each->push_back(ms);
        pool_.push_back(e.first);
        pool_.push_back(e.second);
    }
    if (total_ms) *total_ms = tot;
    if (n_passes) *n_passes = (int64_t)events_.size();
    events_.clear();
}

void Engine::launchAll(const Program& p, const DevOp* d_ops, const double* d_tables, const PhaseTerm* d_terms,
                       cuDoubleComplex* state, uint64_t hi_bits, int64_t init_basis, const StoreRedirect* redirect) {
    bool first = true;
    for (const PassDesc& pd : p.passes) {
        const bool last = (&pd == &p.passes.back());
        PassParams prm;
        prm.state = state;
        prm.ops = d_ops + pd.op_offset;
        prm.phase_tables = d_tables ? reinterpret_cast<const double2*>(d_tables) + pd.phase_table_offset : nullptr;   // (void*: pass_desc.h is plain data)
        prm.phase_terms = d_terms ? d_terms + pd.phase_term_offset : nullptr;
        prm.hi_bits = hi_bits;
        prm.n_tiles = 1ULL << (pd.n - pd.t);
        prm.pd = pd;
        prm.stages = pick_stages(pd, stages_wanted_);
        prm.use_tensor_map = use_tensor_map_ ? 1 : 0;
        prm.init_basis = 0;
        prm.pad = 0;
        prm.init_index = init_basis >= 0 ? (uint64_t)init_basis : 0;
        if (first && init_basis >= 0) {
            // Basis-state input: the driver's memset is the fastest zero fill (7.4 TB/s against 5.7 TB/s from the pass
            // kernel's own stores); the pass then only generates and processes the one tile that is not zero — and a
            // shard that holds no part of the basis state has nothing to process at all.
            first = false;
            if (redirect && (last || redirect->split == 2)) throw std::runtime_error("qsim_b200: basis-state input cannot be combined with a redirected store");
            if (std::getenv("QSIM_INIT_FILL_IN_KERNEL")) prm.init_basis = 1;
            else {
                CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(cuDoubleComplex) << pd.n, stream_));
                if ((uint64_t)init_basis >> pd.n) return;   // an all-zero shard stays all zero through every pass of a program
                prm.init_basis = 2;
            }
        }
        first = false;
        // the pass that carries the exchange: the program's last one, or (second half of a split exchange) its first
        const bool carries = redirect && (redirect->split == 2 ? (&pd == &p.passes.front()) : last);
        prm.redirect = carries ? (redirect->split == 2 ? 4 : (redirect->split == 1 ? 3 : (redirect->in_place ? 2 : 1))) : 0;
        prm.split_bit = redirect ? redirect->split_bit : -1;
        prm.mid_ctas = 0;
        prm.hs_local = carries ? redirect->hs_local : nullptr;
        prm.hs_peer = carries ? redirect->hs_peer : nullptr;
        prm.hs_base = redirect ? redirect->hs_base : 0;
        prm.hs_timeout_ns = redirect ? redirect->hs_timeout_ns : 0;
        prm.hs_error = carries ? redirect->hs_error : nullptr;
        prm.timeline = d_timeline_ ? d_timeline_ + (size_t)((timeline_n_++) % 64) * 8 : nullptr;
        prm.progress = d_timeline_ ? d_timeline_ + 64 * 8 : nullptr;
        prm.redirect_bit = redirect ? redirect->bit : 0;
        prm.redirect_keep = redirect ? redirect->keep_value : 0;
        prm.send_ctas = 0;   // chosen by launch_pass
        prm.dst_keep = carries ? redirect->keep : nullptr;
        prm.dst_send = carries ? redirect->send : nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timing_) {
            e0 = getEvent();
            e1 = getEvent();
            CUDA_CHECK(cudaEventRecord(e0, stream_));
        }
        const size_t pass_i = (size_t)(&pd - p.passes.data());
        // NVTX range per pass (visible in Nsight Systems / ncu --nvtx; a no-op costing nanoseconds when no tool is attached)
        char label[96];
        std::snprintf(label, sizeof(label), "qsim pass %zu/%zu: %d ops, %d sweeps, t=%d%s", pass_i + 1, p.passes.size(), pd.n_ops,
                      pd.n_sweeps, pd.t, prm.redirect ? (prm.redirect == 4 ? ", exchange gathered in place" : (prm.redirect >= 2 ? ", fused exchange in place" : ", fused exchange")) : (prm.init_basis ? ", basis-state input" : ""));
        struct NvtxScope { explicit NvtxScope(const char* l) { nvtxRangePushA(l); } ~NvtxScope() { nvtxRangePop(); } };
        if (p.jit.size() != 2 * p.passes.size()) {   // two builds per pass: one warp group / two (JitSlots)
            p.jit.assign(2 * p.passes.size(), nullptr);
            p.jit_req.assign(2 * p.passes.size(), nullptr);
            p.jit_tried.assign(2 * p.passes.size(), 0);
            p.jit_tune.assign(p.passes.size(), nullptr);
        }
        JitSlots slots;
        slots.kernel = &p.jit[2 * pass_i];
        slots.request = &p.jit_req[2 * pass_i];
        slots.tried = &p.jit_tried[2 * pass_i];
        slots.force = p.force_jit;
        if (stream_ != capture_stream_ || !capture_stream_) {   // (never inside a graph capture: events would become nodes)
            if (!p.jit_tune[pass_i]) p.jit_tune[pass_i] = std::make_shared<DualTune>();
            slots.tune = p.jit_tune[pass_i].get();
        }
        {
            NvtxScope range(label);
            CUDA_CHECK(launch_pass(prm, num_sms_, stream_, p.ops.data() + pd.op_offset, slots));
        }
        ++launches_;
        if (timing_) {
            CUDA_CHECK(cudaEventRecord(e1, stream_));
            events_.emplace_back(e0, e1);
        }
    }
}

void Engine::execute(const Program& p, cuDoubleComplex* state, uint64_t hi_bits, int64_t init_basis) {
    const size_t n = p.ops.size();
    if (p.passes.empty()) return;
    if (n == 0) {   // passes without ops exist: a pure index permutation (deferred X gates)
        launchAll(p, nullptr, nullptr, nullptr, state, hi_bits, init_basis);
        return;
    }
    if (staged_pending_) {           // the pinned buffer may still be in flight from the previous run
        CUDA_CHECK(cudaEventSynchronize(staged_));
        staged_pending_ = false;
    }
    if (n > h_cap_) {
        if (h_ops_) cudaFreeHost(h_ops_);
        h_ops_ = nullptr;
        size_t cap = n + n / 2 + 64;
        CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&h_ops_), cap * sizeof(DevOp), cudaHostAllocDefault));
        h_cap_ = cap;
    }
    if (n > d_cap_) {
        // Earlier launches on stream_ may still read the old buffer: free is stream-ordered.
        if (d_ops_) { CUDA_CHECK(cudaStreamSynchronize(stream_)); cudaFree(d_ops_); }
        d_ops_ = nullptr;
        size_t cap = n + n / 2 + 64;
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_ops_), cap * sizeof(DevOp)));
        d_cap_ = cap;
    }
    std::memcpy(h_ops_, p.ops.data(), n * sizeof(DevOp));
    CUDA_CHECK(cudaMemcpyAsync(d_ops_, h_ops_, n * sizeof(DevOp), cudaMemcpyHostToDevice, stream_));
    CUDA_CHECK(cudaEventRecord(staged_, stream_));
    staged_pending_ = true;
    // fused diagonal runs: tables and terms ride along (synchronous copy from pageable memory: the driver stages
    // it before returning; these programs are rare and the blob is small next to the state)
    const double* d_tables = nullptr;
    const PhaseTerm* d_terms = nullptr;
    const size_t tb = p.phase_tables.size() * sizeof(double), tt = p.phase_terms.size() * sizeof(PhaseTerm);
    if (tb + tt > 0) {
        const size_t tb_pad = (tb + 255) & ~size_t(255);
        if (tb_pad + tt > d_aux_cap_) {
            if (d_aux_) { CUDA_CHECK(cudaStreamSynchronize(stream_)); cudaFree(d_aux_); d_aux_ = nullptr; }
            d_aux_cap_ = (tb_pad + tt) * 2 + 4096;
            CUDA_CHECK(cudaMalloc(&d_aux_, d_aux_cap_));
        }
        if (tb) CUDA_CHECK(cudaMemcpyAsync(d_aux_, p.phase_tables.data(), tb, cudaMemcpyHostToDevice, stream_));
        if (tt) CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(d_aux_) + tb_pad, p.phase_terms.data(), tt, cudaMemcpyHostToDevice, stream_));
        d_tables = static_cast<const double*>(d_aux_);
        d_terms = tt ? reinterpret_cast<const PhaseTerm*>(static_cast<char*>(d_aux_) + tb_pad) : nullptr;
    }
    launchAll(p, d_ops_, d_tables, d_terms, state, hi_bits, init_basis);
}

void Engine::execute(const DeviceProgram& p, cuDoubleComplex* state, uint64_t hi_bits, int64_t init_basis,
                     const StoreRedirect* redirect) {
    if (p.host.passes.empty()) return;
    if (!p.d_ops && !p.host.ops.empty()) throw std::runtime_error("qsim_b200: program was not uploaded");
    // Launch-bound regime: replay the pass sequence as one CUDA graph.  The first run on these amplitudes goes through the
    // plain path (it also resolves the specialised kernels and sets their attributes), the second is captured, later ones
    // replay.  A different state pointer, rank or device captures again.
    const bool graphable = !timing_ && !redirect && init_basis < 0 && p.host.passes.size() >= 2 &&
                           p.host.n_local <= graph_max_qubits_;
    if (graphable) {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        DeviceProgram::Graph& g = p.graph;
        const bool same = g.state == state && g.hi_bits == hi_bits && g.device == dev;
        if (same && g.exec) {
            CUDA_CHECK(cudaGraphLaunch(g.exec, stream_));
            launches_ += (int64_t)p.host.passes.size();
            ++graph_launches_;
            return;
        }
        if (same && g.plain_runs >= 1) {
            if (!capture_stream_) CUDA_CHECK(cudaStreamCreateWithFlags(&capture_stream_, cudaStreamNonBlocking));
            cudaStream_t user = stream_;
            const int64_t launches_before = launches_;
            cudaGraph_t graph = nullptr;
            cudaError_t err = cudaStreamBeginCapture(capture_stream_, cudaStreamCaptureModeThreadLocal);
            if (err == cudaSuccess) {
                stream_ = capture_stream_;
                try {
                    launchAll(p.host, p.d_ops, p.d_tables, p.d_terms, state, hi_bits, init_basis, redirect);
                } catch (...) {
                    stream_ = user;
                    cudaStreamEndCapture(capture_stream_, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                    throw;
                }
                stream_ = user;
                err = cudaStreamEndCapture(capture_stream_, &graph);
            }
            launches_ = launches_before;   // nothing ran yet
            if (err == cudaSuccess && graph) {
                if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
                err = cudaGraphInstantiate(&g.exec, graph, 0);
                cudaGraphDestroy(graph);
            }
            if (err == cudaSuccess && g.exec) {
                CUDA_CHECK(cudaGraphLaunch(g.exec, stream_));
                launches_ += (int64_t)p.host.passes.size();
                ++graph_launches_;
                return;
            }
            cudaGetLastError();            // capture unavailable: stay on the plain path
            g.exec = nullptr;
            g.plain_runs = -(1 << 30);
        }
        if (!same) {
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            g.state = state; g.hi_bits = hi_bits; g.device = dev; g.plain_runs = 0;
        }
        ++g.plain_runs;
    }
    launchAll(p.host, p.d_ops, p.d_tables, p.d_terms, state, hi_bits, init_basis, redirect);
}

}  // namespace b200
}  // namespace qsim
