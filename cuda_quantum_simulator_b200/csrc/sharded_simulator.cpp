// qsim::ShardedSimulator: the multi-GPU driver in C++17 — planner (sharded_plan.cpp), layout / X-frame bookkeeping, CUDA-IPC
// peer memory, fused and separate exchanges, the buffer-flip invariant, distributed read-out — on NCCL, which is loaded at run
// time (dlopen) so that the library still loads on a box without it.  See include/qsim/sharded_simulator.hpp.
//
// Invariants every rank keeps (they replace any data-path handshake):
//   * all ranks make the same calls in the same order; every exchange is bracketed by stream-ordered barriers (an NCCL
//     all-reduce of one float on the engine's stream): nobody's later kernels start before everybody's earlier ones ended;
//   * a FUSED exchange writes only into the buffers nobody is reading (each rank's second buffer), so it needs one barrier,
//     after; every rank flips cur_ at the same steps, so "buffer cur_" names the same side everywhere.
#include "qsim/sharded_simulator.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "qsim_b200.h"
#include "sharded_plan.hpp"

namespace qsim {

namespace {

void chk(qsim_status_t st) {
    if (st == QSIM_OK) return;
    const std::string msg = qsim_last_error();
    if (st == QSIM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if (st == QSIM_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);
    throw std::runtime_error(msg);
}

void cuda_chk(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

// ---- NCCL, resolved at run time ----
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

const NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        std::vector<std::string> names;
        if (const char* e = std::getenv("QSIM_NCCL_LIB")) names.push_back(e);
        names.push_back("libnccl.so.2");   // (a copy some other library of the process has loaded under this soname is reused)
        names.push_back("libnccl.so");
        for (const auto& nm : names) {
            a.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) throw std::runtime_error("qsim_b200: libnccl.so.2 not found (set QSIM_NCCL_LIB): the sharded simulator needs NCCL");
        auto sym = [&](const char* s) {
            void* p = dlsym(a.handle, s);
            if (!p) throw std::runtime_error(std::string("qsim_b200: NCCL lacks ") + s);
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        return a;
    }();
    return api;
}

void nccl_chk(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r));
}

qsim_gate_t to_record(const GateOp& g) {
    qsim_gate_t r{};
    r.type = (int)g.type;
    r.q0 = g.qubits.size() > 0 ? g.qubits[0] : -1;
    r.q1 = g.qubits.size() > 1 ? g.qubits[1] : -1;
    r.q2 = g.qubits.size() > 2 ? g.qubits[2] : -1;
    r.param = g.parameter;
    return r;
}

GateOp from_record(const qsim_gate_t& r) {
    const GateType t = static_cast<GateType>(r.type);
    if (r.q2 >= 0) return GateOp(t, r.q0, r.q1, r.q2);
    if (r.q1 >= 0) return r.param != 0.0 || r.type == QSIM_GATE_CRY || r.type == QSIM_GATE_CRZ ? GateOp(t, r.q0, r.q1, r.param) : GateOp(t, r.q0, r.q1);
    return r.type >= QSIM_GATE_RX && r.type <= QSIM_GATE_RZ ? GateOp(t, r.q0, r.param) : GateOp(t, r.q0);
}

std::vector<qsim_gate_t> records_of(const std::vector<GateOp>& gates) {
    std::vector<qsim_gate_t> recs;
    recs.reserve(gates.size());
    for (const GateOp& g : gates) recs.push_back(to_record(g));
    return recs;
}

}  // namespace

// ---- planning API ---------------------------------------------------------------------------------------------------

int ShardPlan::numSwaps() const {
    int k = 0;
    for (const auto& s : steps) k += s.is_swap ? 1 : 0;
    return k;
}

std::vector<int> chooseInitialLayout(int num_qubits, int n_global, const std::vector<GateOp>& gates) {
    const auto recs = records_of(gates);
    return b200::shard_choose_initial_layout(num_qubits, n_global, recs.data(), (int64_t)recs.size());
}

ShardPlan planCircuit(int num_qubits, int n_global, const std::vector<GateOp>& gates, const std::vector<int>& perm) {
    const auto recs = records_of(gates);
    const b200::ShardPlanRec p = b200::shard_plan_circuit(num_qubits, n_global, recs.data(), (int64_t)recs.size(), perm);
    ShardPlan out;
    out.num_qubits = num_qubits;
    out.n_global = n_global;
    out.perm = p.perm;
    for (const auto& st : p.steps) {
        ShardStep s;
        s.is_swap = st.is_swap;
        s.global_qubit = st.global_qubit;
        s.local_qubit = st.local_qubit;
        for (const auto& r : st.gates) s.gates.push_back(from_record(r));
        out.steps.push_back(std::move(s));
    }
    return out;
}

// ---- communicator ---------------------------------------------------------------------------------------------------

struct ShardedSimulator::Comm {
    ncclComm_t comm = nullptr;
    float* flag = nullptr;          // barrier token
    unsigned char* stage = nullptr; // small device staging buffer for host-side collectives
    size_t stage_bytes = 0;
    int world = 1;
    ~Comm() {
        if (comm) nccl().CommDestroy(comm);
        if (flag) cudaFree(flag);
        if (stage) cudaFree(stage);
    }
    void ensure(size_t bytes, cudaStream_t s) {
        if (bytes <= stage_bytes) return;
        if (stage) { cudaStreamSynchronize(s); cudaFree(stage); stage = nullptr; }
        stage_bytes = std::max<size_t>(bytes * 2, 1 << 16);
        cuda_chk(cudaMalloc(reinterpret_cast<void**>(&stage), stage_bytes), "cudaMalloc(stage)");
    }
    // every rank contributes `bytes` bytes; returns world * bytes on the host
    std::vector<unsigned char> allGatherBytes(const void* mine, size_t bytes, cudaStream_t s) {
        std::vector<unsigned char> out(bytes * (size_t)world);
        if (world == 1) { std::memcpy(out.data(), mine, bytes); return out; }
        ensure(bytes * (size_t)(world + 1), s);
        cuda_chk(cudaMemcpyAsync(stage, mine, bytes, cudaMemcpyHostToDevice, s), "stage h2d");
        nccl_chk(nccl().AllGather(stage, stage + bytes, bytes, ncclUint8, comm, s), "ncclAllGather");
        cuda_chk(cudaMemcpyAsync(out.data(), stage + bytes, bytes * (size_t)world, cudaMemcpyDeviceToHost, s), "stage d2h");
        cuda_chk(cudaStreamSynchronize(s), "sync");
        return out;
    }
};

std::array<unsigned char, ShardedSimulator::kUniqueIdBytes> ShardedSimulator::createUniqueId() {
    static_assert(sizeof(ncclUniqueId) == kUniqueIdBytes, "ncclUniqueId size");
    ncclUniqueId id;
    nccl_chk(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::array<unsigned char, kUniqueIdBytes> out;
    std::memcpy(out.data(), &id, kUniqueIdBytes);
    return out;
}

// ---- construction ---------------------------------------------------------------------------------------------------

namespace {
constexpr size_t kHsWords = 1024;                       // one per CTA of a pass launch (grid <= number of SMs)
constexpr size_t kHsBytes = kHsWords * 8 + 64;          // + the error word
}  // namespace

struct ShardedSimulator::CompiledPlan {
    b200::ShardPlanRec plan;
    std::vector<qsim_program_t*> programs;   // per step: program or nullptr (swap)
    uint64_t frame_after = 0, frame_before = 0;
    std::vector<int> perm_before;
    bool from_pristine = false;
    int n_passes = 0, n_ops = 0;
    ~CompiledPlan() {
        for (qsim_program_t* p : programs)
            if (p) qsim_program_destroy(p);
    }
};

int ShardedSimulator::planPasses(const CompiledPlan& p) { return p.n_passes; }
int ShardedSimulator::planSwaps(const CompiledPlan& p) { return p.plan.n_swaps(); }
int ShardedSimulator::planOps(const CompiledPlan& p) { return p.n_ops; }

ShardedSimulator::ShardedSimulator(int num_qubits, int rank, int world_size, const unsigned char* unique_id, Exchange exchange)
    : n_(num_qubits), rank_(rank), world_(world_size), exchange_(exchange) {
    if (world_size < 1 || (world_size & (world_size - 1)) != 0) throw std::invalid_argument("world size must be a power of two");
    if (rank < 0 || rank >= world_size) throw std::invalid_argument("rank out of range");
    ng_ = 0;
    while ((1 << ng_) < world_size) ++ng_;
    nl_ = n_ - ng_;
    if (nl_ < 1) throw std::invalid_argument("more rank qubits than qubits");
    perm_.resize(n_);
    for (int q = 0; q < n_; ++q) perm_[q] = q;
    comm_ = std::make_unique<Comm>();
    comm_->world = world_;
    if (world_ > 1) {
        if (!unique_id) throw std::invalid_argument("unique_id must not be null");
        ncclUniqueId id;
        std::memcpy(&id, unique_id, kUniqueIdBytes);
        nccl_chk(nccl().CommInitRank(&comm_->comm, world_, id, rank_), "ncclCommInitRank");
    }
    cuda_chk(cudaMalloc(reinterpret_cast<void**>(&comm_->flag), sizeof(float)), "cudaMalloc(flag)");
    cuda_chk(cudaMemset(comm_->flag, 0, sizeof(float)), "memset(flag)");
    const size_t shard_bytes = sizeof(cuDoubleComplex) << nl_;
    cuda_chk(cudaMalloc(reinterpret_cast<void**>(&bufs_[0]), shard_bytes), "cudaMalloc(shard)");
    qsim_sim_t* h = nullptr;
    chk(qsim_shard_create(n_, ng_, rank_, bufs_[0], &h));
    shard_ = h;
    if (world_ > 1 && exchange_ != Exchange::Nccl) {
        // second buffer for the fused exchange, if every rank has room for it
        size_t free_b = 0, total_b = 0;
        cuda_chk(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        double want = (std::getenv("QSIM_NO_FUSED_EXCHANGE") == nullptr && free_b > shard_bytes + ((size_t)4 << 30)) ? 1.0 : 0.0;
        const auto all = allGather(want);
        if (std::all_of(all.begin(), all.end(), [](double v) { return v != 0.0; }))
            cuda_chk(cudaMalloc(reinterpret_cast<void**>(&bufs_[1]), shard_bytes), "cudaMalloc(second shard buffer)");
        cuda_chk(cudaMalloc(reinterpret_cast<void**>(&hs_), kHsBytes), "cudaMalloc(handshake words)");
        cuda_chk(cudaMemset(hs_, 0, kHsBytes), "memset(handshake words)");
        try {
            openPeers();
            exchange_ = Exchange::PeerMemory;
        } catch (const std::exception& e) {
            if (exchange == Exchange::PeerMemory) throw;
            std::fprintf(stderr, "qsim_b200 rank %d: CUDA-IPC peer memory unavailable (%s); global-qubit swaps fall back to NCCL "
                         "send/recv through bounce buffers\n", rank_, e.what());
            exchange_ = Exchange::Nccl;
            if (bufs_[1]) { cudaFree(bufs_[1]); bufs_[1] = nullptr; }
        }
    } else if (world_ > 1) {
        exchange_ = Exchange::Nccl;
    }
}

ShardedSimulator::~ShardedSimulator() {
    if (shard_) {
        qsim_sim_synchronize(static_cast<qsim_sim_t*>(shard_));
        qsim_sim_destroy(static_cast<qsim_sim_t*>(shard_));
    }
    for (void* b : peer_base_) cudaIpcCloseMemHandle(b);
    for (auto*& b : bounce_) if (b) cudaFree(b);
    for (auto*& b : bufs_) if (b) cudaFree(b);
    if (hs_) cudaFree(hs_);
}

const char* ShardedSimulator::exchangeName() const {
    return world_ == 1 ? "none" : (exchange_ == Exchange::PeerMemory ? "p2p" : "nccl");
}

void ShardedSimulator::setStream(cudaStream_t s) {
    stream_ = s;
    chk(qsim_sim_set_stream(static_cast<qsim_sim_t*>(shard_), s));
}

void ShardedSimulator::synchronize() {
    chk(qsim_sim_synchronize(static_cast<qsim_sim_t*>(shard_)));
    checkExchanges(false);
}

void ShardedSimulator::checkExchanges(bool collective) {
    if (!hs_unchecked_) return;
    int err = 0;
    cuda_chk(cudaMemcpyAsync(&err, reinterpret_cast<unsigned char*>(hs_) + kHsWords * 8, sizeof(int), cudaMemcpyDeviceToHost, stream_),
             "read handshake status");
    cuda_chk(cudaStreamSynchronize(stream_), "sync");
    if (collective) {   // every rank learns of any rank's failure, so that all of them leave the collective call sequence together
        for (double v : allGather((double)err))
            if (v != 0.0 && err == 0) err = (int)v;
        hs_unchecked_ = false;
    }
    if (err != 0)
        throw std::runtime_error(err == 1 ? "qsim_b200: an in-place qubit exchange timed out waiting for the partner GPU (the state is invalid)"
                                          : "qsim_b200: an in-place qubit exchange could not split its grid (the state is invalid)");
}

void ShardedSimulator::barrier() {
    if (world_ > 1)
        nccl_chk(nccl().AllReduce(comm_->flag, comm_->flag, 1, ncclFloat, ncclSum, comm_->comm, stream_), "ncclAllReduce(barrier)");
}

std::vector<double> ShardedSimulator::allGather(double v) {
    const auto bytes = comm_->allGatherBytes(&v, sizeof(double), stream_);
    std::vector<double> out(world_);
    std::memcpy(out.data(), bytes.data(), sizeof(double) * (size_t)world_);
    return out;
}

void ShardedSimulator::openPeers() {
    struct Info {
        unsigned char handle[3][64];   // the shard's buffer(s), then the handshake words
        uint64_t offset[3];
        int32_t n_bufs;
        int32_t pad;
    } mine{};
    mine.n_bufs = bufs_[1] ? 2 : 1;
    for (int b = 0; b < mine.n_bufs; ++b) chk(qsim_ipc_get_handle(bufs_[b], mine.handle[b], &mine.offset[b]));
    chk(qsim_ipc_get_handle(hs_, mine.handle[2], &mine.offset[2]));
    const auto all = comm_->allGatherBytes(&mine, sizeof(Info), stream_);
    peer_ptr_.assign(ng_, {nullptr, nullptr});
    peer_hs_.assign(ng_, nullptr);
    for (int b = 0; b < ng_; ++b) {
        const int peer = rank_ ^ (1 << b);
        Info pi;
        std::memcpy(&pi, all.data() + sizeof(Info) * (size_t)peer, sizeof(Info));
        {
            void* base = nullptr;
            chk(qsim_ipc_open_handle(pi.handle[2], &base));
            peer_base_.push_back(base);
            peer_hs_[b] = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(base) + pi.offset[2]);
        }
        for (int k = 0; k < pi.n_bufs && k < mine.n_bufs; ++k) {
            void* base = nullptr;
            chk(qsim_ipc_open_handle(pi.handle[k], &base));
            peer_base_.push_back(base);
            peer_ptr_[b][k] = reinterpret_cast<cuDoubleComplex*>(static_cast<unsigned char*>(base) + pi.offset[k]);
        }
    }
}

// ---- execution ------------------------------------------------------------------------------------------------------

void ShardedSimulator::reset() {
    chk(qsim_sim_reset(static_cast<qsim_sim_t*>(shard_)));
    for (int q = 0; q < n_; ++q) perm_[q] = q;
    frame_ = 0;
    pristine_ = true;
    order_preserving_ = true;
}

cuDoubleComplex* ShardedSimulator::devicePtr() {
    return static_cast<cuDoubleComplex*>(qsim_sim_device_ptr(static_cast<qsim_sim_t*>(shard_)));
}

std::vector<std::complex<double>> ShardedSimulator::getLocalState() {
    checkExchanges(false);
    std::vector<std::complex<double>> out(size_t(1) << nl_);
    chk(qsim_sim_get_state(static_cast<qsim_sim_t*>(shard_), reinterpret_cast<double*>(out.data())));
    return out;
}

void ShardedSimulator::setLocalState(const std::complex<double>* amplitudes) {
    chk(qsim_sim_set_state(static_cast<qsim_sim_t*>(shard_), reinterpret_cast<const double*>(amplitudes)));
    pristine_ = false;
}

std::shared_ptr<ShardedSimulator::CompiledPlan> ShardedSimulator::compile(const Circuit& circuit) {
    if (circuit.getNumQubits() != n_) throw std::invalid_argument("Circuit qubit count doesn't match simulator");
    const auto recs = records_of(circuit.getGates());
    auto cp = std::make_shared<CompiledPlan>();
    std::vector<int> start = perm_;
    if (pristine_ && ng_ > 0 && !identity_only_ && std::getenv("QSIM_NO_LAYOUT") == nullptr) {
        start = b200::shard_choose_initial_layout(n_, ng_, recs.data(), (int64_t)recs.size());   // carried by the plan
        cp->from_pristine = true;
    }
    cp->plan = b200::shard_plan_circuit(n_, ng_, recs.data(), (int64_t)recs.size(), start);
    cp->perm_before = start;
    cp->frame_before = frame_;
    uint64_t frame = frame_;
    int after_swap_of = -1;   // the local position of the exchange right before this segment (the gathering half of a split exchange)
    for (const auto& st : cp->plan.steps) {
        if (!st.is_swap) {
            qsim_program_t* prog = nullptr;
            chk(qsim_program_compile_ex2(n_, ng_, st.gates.data(), (int64_t)st.gates.size(), frame, after_swap_of, &prog));
            after_swap_of = -1;
            cp->programs.push_back(prog);
            int64_t info[8];
            chk(qsim_program_info(prog, info));
            cp->n_passes += (int)info[0];
            cp->n_ops += (int)info[1];
            frame = (uint64_t)info[6] << nl_;          // the local part was applied by the program itself
        } else {
            cp->programs.push_back(nullptr);
            after_swap_of = st.local_qubit;
            const int g = st.global_qubit, l = st.local_qubit;   // a pending X travels with its qubit
            const uint64_t bg = (frame >> g) & 1, bl = (frame >> l) & 1;
            frame = (frame & ~((1ULL << g) | (1ULL << l))) | (bl << g) | (bg << l);
        }
    }
    cp->frame_after = frame;
    return cp;
}

std::vector<std::shared_ptr<ShardedSimulator::CompiledPlan>> ShardedSimulator::compileSequence(const Circuit& circuit, int k) {
    // run i is compiled against the layout and X frame run i-1 leaves behind; plans are shared when the layout repeats
    struct Saved { std::vector<int> perm; uint64_t frame; bool pristine; } saved{perm_, frame_, pristine_};
    struct Entry { std::vector<int> perm; uint64_t frame; bool pristine; std::shared_ptr<CompiledPlan> plan; };
    std::vector<Entry> cache;
    std::vector<std::shared_ptr<CompiledPlan>> out;
    try {
        for (int i = 0; i < k; ++i) {
            std::shared_ptr<CompiledPlan> hit;
            for (const Entry& e : cache)
                if (e.perm == perm_ && e.frame == frame_ && e.pristine == pristine_) { hit = e.plan; break; }
            if (!hit) {
                hit = compile(circuit);
                cache.push_back(Entry{perm_, frame_, pristine_, hit});
            }
            out.push_back(hit);
            perm_ = hit->plan.perm;
            frame_ = hit->frame_after;
            pristine_ = false;
        }
    } catch (...) {
        perm_ = saved.perm; frame_ = saved.frame; pristine_ = saved.pristine;
        throw;
    }
    perm_ = saved.perm; frame_ = saved.frame; pristine_ = saved.pristine;
    return out;
}

void ShardedSimulator::swapSeparate(int g, int l) {
    qsim_sim_t* h = static_cast<qsim_sim_t*>(shard_);
    const int b = g - nl_, peer = rank_ ^ (1 << b);
    qsim_sim_device_ptr(h);   // a lazily reset shard is written out now: the peer is about to read it
    barrier();
    if (exchange_ == Exchange::PeerMemory) chk(qsim_shard_swap_p2p(h, peer_ptr_[b][cur_], g, l));
    else swapNccl(peer, g, l);
    barrier();
    ++separate_exchanges_;
}

void ShardedSimulator::swapQubits(int global_position, int local_position) {
    if (world_ == 1) throw std::invalid_argument("a single shard has no global qubits");
    swapSeparate(global_position, local_position);
}

void ShardedSimulator::swapNccl(int peer, int g, int l) {
    qsim_sim_t* h = static_cast<qsim_sim_t*>(shard_);
    const int my_bit = (rank_ >> (g - nl_)) & 1;
    const uint64_t pairs = 1ULL << (nl_ - 1);
    const uint64_t chunk = std::min<uint64_t>(pairs, 1ULL << 24);   // 256 MiB bounce buffers
    const int64_t n_chunks = (int64_t)(pairs / chunk);
    if (bounce_amps_ < chunk) {
        for (auto*& b : bounce_) { if (b) cudaFree(b); b = nullptr; }
        for (auto*& b : bounce_) cuda_chk(cudaMalloc(reinterpret_cast<void**>(&b), chunk * sizeof(cuDoubleComplex)), "cudaMalloc(bounce)");
        bounce_amps_ = chunk;
    }
    for (int64_t c = 0; c < n_chunks; ++c) {
        chk(qsim_shard_pack_half(h, l, my_bit, c, n_chunks, bounce_[0]));   // the half whose bit l differs from my value of g leaves
        nccl_chk(nccl().GroupStart(), "ncclGroupStart");
        nccl_chk(nccl().Send(bounce_[0], chunk * 2, ncclDouble, peer, comm_->comm, stream_), "ncclSend");
        nccl_chk(nccl().Recv(bounce_[1], chunk * 2, ncclDouble, peer, comm_->comm, stream_), "ncclRecv");
        nccl_chk(nccl().GroupEnd(), "ncclGroupEnd");
        chk(qsim_shard_unpack_half(h, l, my_bit, c, n_chunks, bounce_[1]));
    }
}

bool ShardedSimulator::runThenSwap(void* program, int g, int l) {
    if (exchange_ != Exchange::PeerMemory || std::getenv("QSIM_NO_FUSED_EXCHANGE") != nullptr) return false;
    qsim_program_t* prog = static_cast<qsim_program_t*>(program);
    if (!bufs_[1] || std::getenv("QSIM_FORCE_INPLACE_EXCHANGE") != nullptr) {
        // no room for a second buffer (36 qubits on 8 GPUs: 128 GiB shards): the same fusion IN PLACE, the stores over the
        // partner's live shard ordered tile by tile by the kernels' own handshake.  No barrier before: the partner's kernel
        // only publishes a tile as loaded from inside THIS step's pass, which its stream runs after everything earlier.
        if (std::getenv("QSIM_NO_INPLACE_EXCHANGE") != nullptr || !hs_) return false;
        int ok = 0;
        chk(qsim_shard_inplace_exchange_possible(static_cast<qsim_sim_t*>(shard_), prog, l, &ok));
        if (!ok) return false;   // (a property of the compiled pass: the same answer on every rank)
        const int b = g - nl_;
        ++hs_epoch_;
        chk(qsim_shard_execute_exchange_inplace(static_cast<qsim_sim_t*>(shard_), prog, peer_ptr_[b][cur_], g, l, hs_, peer_hs_[b],
                                                hs_epoch_ << 32, 0, reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(hs_) + kHsWords * 8)));
        hs_unchecked_ = true;
        ++fused_exchanges_;
        ++inplace_exchanges_;
        barrier();
        return true;
    }
    uint64_t mask = 0;
    chk(qsim_program_last_tile_mask(prog, &mask));
    if (mask == 0 || ((mask >> l) & 1)) return false;   // no pass to ride on, or the swapped qubit is one of its tile qubits
    const int b = g - nl_, alt = 1 - cur_;
    chk(qsim_shard_execute_exchange(static_cast<qsim_sim_t*>(shard_), prog, bufs_[alt], peer_ptr_[b][alt], g, l));
    cur_ = alt;
    ++fused_exchanges_;
    barrier();
    return true;
}

// A forced exchange between two programs, when there is no room for a second buffer: split over the last pass of the program
// before it (scatters half of the leaving tiles into the partner's shard) and the first pass of the program after it (gathers
// the other half from the partner's shard), so that each pass hides half of the NVLink time instead of one pass hiding what it
// can of all of it.  Both under the kernels' own handshake; a barrier after each half.
bool ShardedSimulator::runSwapSplit(void* before, void* after, int g, int l) {
    if (exchange_ != Exchange::PeerMemory || !hs_ || std::getenv("QSIM_NO_FUSED_EXCHANGE") != nullptr ||
        std::getenv("QSIM_NO_INPLACE_EXCHANGE") != nullptr || std::getenv("QSIM_NO_SPLIT_EXCHANGE") != nullptr)
        return false;
    if (bufs_[1] && std::getenv("QSIM_FORCE_INPLACE_EXCHANGE") == nullptr) return false;   // (room for the second buffer: that path)
    qsim_sim_t* h = static_cast<qsim_sim_t*>(shard_);
    qsim_program_t* pa = static_cast<qsim_program_t*>(before);
    qsim_program_t* pb = static_cast<qsim_program_t*>(after);
    int w = -1;
    chk(qsim_shard_split_exchange_possible(h, pa, pb, l, &w));
    if (w < 0) return false;   // (a property of the two compiled passes: the same answer on every rank)
    const int b = g - nl_;
    int* err = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(hs_) + kHsWords * 8);
    ++hs_epoch_;
    chk(qsim_shard_execute_exchange_half(h, pa, peer_ptr_[b][cur_], g, l, w, 1, hs_, peer_hs_[b], hs_epoch_ << 32, 0, err));
    barrier();   // the partner's half-finished shard is complete before anything is gathered from it
    ++hs_epoch_;
    chk(qsim_shard_execute_exchange_half(h, pb, peer_ptr_[b][cur_], g, l, w, 2, hs_, peer_hs_[b], hs_epoch_ << 32, 0, err));
    barrier();
    hs_unchecked_ = true;
    ++fused_exchanges_;
    ++inplace_exchanges_;
    ++split_exchanges_;
    return true;
}

void ShardedSimulator::execute(const CompiledPlan& cp) {
    const bool same = (cp.perm_before == perm_) && cp.frame_before == frame_;
    if (!same && !(cp.from_pristine && pristine_ && cp.frame_before == frame_))
        throw std::invalid_argument("plan compiled against a different qubit layout / X frame than the state now has (a plan that "
                                    "changes the layout cannot be executed twice in a row): reset(), or compile it against the "
                                    "current state (compile / compileSequence)");
    const auto& steps = cp.plan.steps;
    for (size_t i = 0; i < steps.size(); ++i) {
        if (!steps[i].is_swap) {
            // program, swap, program (and no further swap riding on that second program): the exchange split over both
            if (i + 2 < steps.size() && steps[i + 1].is_swap && !steps[i + 2].is_swap &&
                !(i + 3 < steps.size() && steps[i + 3].is_swap) &&
                runSwapSplit(cp.programs[i], cp.programs[i + 2], steps[i + 1].global_qubit, steps[i + 1].local_qubit)) {
                i += 2;
                continue;
            }
            if (i + 1 < steps.size() && steps[i + 1].is_swap &&
                runThenSwap(cp.programs[i], steps[i + 1].global_qubit, steps[i + 1].local_qubit)) {
                ++i;   // the program's last pass carried the exchange
                continue;
            }
            chk(qsim_sim_execute(static_cast<qsim_sim_t*>(shard_), cp.programs[i]));
        } else {
            swapSeparate(steps[i].global_qubit, steps[i].local_qubit);
        }
    }
    perm_ = cp.plan.perm;
    frame_ = cp.frame_after;
    order_preserving_ = (pristine_ ? true : order_preserving_) && cp.plan.n_swaps() == 0;
    pristine_ = false;
}

void ShardedSimulator::run(const Circuit& circuit) {
    auto cp = compile(circuit);
    execute(*cp);
}

// ---- layout ---------------------------------------------------------------------------------------------------------

void ShardedSimulator::relabelIdentity() {
    for (int q = 0; q < n_; ++q) perm_[q] = q;
    frame_ = 0;
    pristine_ = false;
    order_preserving_ = true;
}

void ShardedSimulator::restoreIdentityLayout() {
    std::vector<int> perm = perm_;
    uint64_t frame = frame_;
    auto swap = [&](int g, int l) {
        swapSeparate(g, l);
        int qg = -1, ql = -1;
        for (int q = 0; q < n_; ++q) { if (perm[q] == g) qg = q; if (perm[q] == l) ql = q; }
        perm[qg] = l;
        perm[ql] = g;
        const uint64_t bg = (frame >> g) & 1, bl = (frame >> l) & 1;
        frame = (frame & ~((1ULL << g) | (1ULL << l))) | (bl << g) | (bg << l);
    };
    for (int g = nl_; g < n_; ++g) {
        if (perm[g] == g) continue;
        if (perm[g] >= nl_) swap(perm[g], nl_ - 1);   // sits in another rank bit: bring it down to a local position first
        swap(g, perm[g]);
    }
    // local part: sort the positions with SWAP gates on PHYSICAL qubits (cycle sort) plus any X frame the exchanges moved
    // onto local bits: one program, applied by its addressing
    std::vector<int> inv(n_);
    for (int q = 0; q < n_; ++q) inv[perm[q]] = q;
    std::vector<qsim_gate_t> recs;
    for (int pos = 0; pos < nl_; ++pos) {
        while (inv[pos] != pos) {
            const int q = inv[pos];   // belongs at position q
            qsim_gate_t r{};
            r.type = QSIM_GATE_SWAP; r.q0 = pos; r.q1 = q; r.q2 = -1; r.param = 0.0;
            recs.push_back(r);
            inv[pos] = inv[q];
            inv[q] = q;
        }
    }
    if (!recs.empty() || (frame & ((1ULL << nl_) - 1))) {
        qsim_program_t* prog = nullptr;
        chk(qsim_program_compile_ex(n_, ng_, recs.empty() ? nullptr : recs.data(), (int64_t)recs.size(), frame, &prog));
        int64_t info[8];
        qsim_status_t st = qsim_program_info(prog, info);
        if (st == QSIM_OK) st = qsim_sim_execute(static_cast<qsim_sim_t*>(shard_), prog);
        qsim_program_destroy(prog);
        chk(st);
        frame = (uint64_t)info[6] << nl_;
    }
    for (int q = 0; q < n_; ++q) perm_[q] = q;
    frame_ = frame;
    order_preserving_ = true;
    pristine_ = false;
}

// ---- read-out -------------------------------------------------------------------------------------------------------

double ShardedSimulator::getTotalProbability() {
    checkExchanges(true);
    double part = 0.0;
    chk(qsim_shard_partial_probability(static_cast<qsim_sim_t*>(shard_), -1, &part));
    double tot = 0.0;
    for (double v : allGather(part)) tot += v;
    return tot;
}

int ShardedSimulator::measureBit(int bit, double uniform, double* p0_out) {
    checkExchanges(true);
    if (bit < 0 || bit >= n_) throw std::invalid_argument("Qubit index out of range");
    qsim_sim_t* h = static_cast<qsim_sim_t*>(shard_);
    const int pos = perm_[bit];
    const int fx = (int)(frame_ >> nl_);
    double part = 0.0;
    int mine = 0;
    if (pos < nl_) {
        chk(qsim_shard_partial_probability(h, pos, &part));
        if ((frame_ >> pos) & 1) { double all = 0.0; chk(qsim_shard_partial_probability(h, -1, &all)); part = all - part; }
    } else {
        mine = ((rank_ ^ fx) >> (pos - nl_)) & 1;
        if (mine == 0) chk(qsim_shard_partial_probability(h, -1, &part));
    }
    const auto parts = allGather(part);
    double p0 = 0.0;
    for (int r = 0; r < world_; ++r) p0 += parts[r ^ fx];   // frame-resolved order, the same on every rank
    if (p0_out) *p0_out = p0;
    const int outcome = uniform < p0 ? 0 : 1;
    const double prob = outcome == 0 ? p0 : 1.0 - p0;
    if (prob <= 0.0) throw std::runtime_error("Measurement outcome has zero probability");
    const double scale = 1.0 / std::sqrt(prob);
    if (pos < nl_) chk(qsim_shard_collapse(h, pos, outcome ^ (int)((frame_ >> pos) & 1), scale));
    else chk(qsim_shard_collapse(h, -1, 0, mine == outcome ? scale : 0.0));
    pristine_ = false;
    return outcome;
}

int ShardedSimulator::measureQubit(int qubit, double uniform) {
    if (qubit < 0 || qubit >= n_) throw std::invalid_argument("Qubit index out of range");
    return measureBit(n_ - 1 - qubit, uniform);   // the reference's Simulator::measureQubit addresses index bit n-1-q
}

std::vector<double> ShardedSimulator::getMarginalProbabilities(const std::vector<int>& qubits) {
    checkExchanges(true);
    const int k = (int)qubits.size();
    if (k > 12) throw std::invalid_argument("at most 12 qubits");
    std::vector<int> phys(k);
    for (int i = 0; i < k; ++i) {
        if (qubits[i] < 0 || qubits[i] >= n_) throw std::out_of_range("Qubit index out of range");
        phys[i] = perm_[qubits[i]];
        for (int j = 0; j < i; ++j)
            if (phys[j] == phys[i]) throw std::invalid_argument("Duplicate qubit in marginal");
    }
    std::vector<int> loc_i, loc_p, glob_i, glob_p;
    for (int i = 0; i < k; ++i) (phys[i] < nl_ ? loc_i : glob_i).push_back(i), (phys[i] < nl_ ? loc_p : glob_p).push_back(phys[i]);
    std::vector<double> mine(size_t(1) << loc_p.size());
    chk(qsim_sim_marginal(static_cast<qsim_sim_t*>(shard_), loc_p.empty() ? nullptr : loc_p.data(), (int)loc_p.size(), mine.data()));
    const auto all = comm_->allGatherBytes(mine.data(), mine.size() * sizeof(double), stream_);
    const int fx = (int)(frame_ >> nl_);
    std::vector<double> out(size_t(1) << k, 0.0);
    for (int r = 0; r < world_; ++r) {
        const double* share = reinterpret_cast<const double*>(all.data()) + (size_t)r * mine.size();
        const int rank_bits = r ^ fx;   // stored rank r holds frame-resolved rank r ^ fx
        size_t base = 0;
        for (size_t j = 0; j < glob_p.size(); ++j) base |= (size_t)((rank_bits >> (glob_p[j] - nl_)) & 1) << glob_i[j];
        for (size_t sub = 0; sub < mine.size(); ++sub) {
            size_t idx = base;
            for (size_t j = 0; j < loc_i.size(); ++j) idx |= ((sub >> j) & 1) << loc_i[j];
            out[idx] += share[sub];
        }
    }
    return out;
}

std::vector<int64_t> ShardedSimulator::sample(const std::vector<double>& uniforms) {
    checkExchanges(true);
    qsim_sim_t* h = static_cast<qsim_sim_t*>(shard_);
    bool identity = true;
    for (int q = 0; q < n_; ++q) identity = identity && perm_[q] == q;
    if (!order_preserving_ && !identity) restoreIdentityLayout();   // the reference's CDF runs in logical index order
    const int64_t shots = (int64_t)uniforms.size();
    const int fx = (int)(frame_ >> nl_);
    const int my_pos = rank_ ^ fx;   // position of my shard in the frame-resolved order
    if (world_ > 1) {
        // every shard sweeps its amplitudes now, at the same time; the chain below only stitches and samples
        double approx = 0.0;
        chk(qsim_shard_cdf_prepare(h, &approx));
        const auto totals = allGather(approx);   // indexed by physical rank
        double before = 0.0;
        for (int p = 0; p < my_pos; ++p) before += totals[p ^ fx];
        chk(qsim_shard_cdf_classify(h, before));
    }
    std::vector<int64_t> result(shots, -1), local(shots);
    double c = 0.0;
    for (int pos = 0; pos < world_; ++pos) {   // chain: shard `pos` continues from the exact sum so far
        double c_end = -1.0;
        if (pos == my_pos) {
            chk(qsim_shard_sample(h, c, pos == 0, uniforms.data(), shots, local.data(), &c_end));
            for (int64_t i = 0; i < shots; ++i)
                if (local[i] >= 0) result[i] = ((int64_t)pos << nl_) | local[i];
        }
        if (world_ > 1) {
            const auto ends = allGather(c_end);
            c = *std::max_element(ends.begin(), ends.end());
        } else c = c_end;
    }
    if (world_ > 1 && shots > 0) {   // every shot was claimed by exactly one shard (or none: past the end)
        comm_->ensure((size_t)shots * sizeof(int64_t), stream_);
        cuda_chk(cudaMemcpyAsync(comm_->stage, result.data(), (size_t)shots * sizeof(int64_t), cudaMemcpyHostToDevice, stream_), "h2d");
        nccl_chk(nccl().AllReduce(comm_->stage, comm_->stage, (size_t)shots, ncclInt64, ncclMax, comm_->comm, stream_), "ncclAllReduce(max)");
        cuda_chk(cudaMemcpyAsync(result.data(), comm_->stage, (size_t)shots * sizeof(int64_t), cudaMemcpyDeviceToHost, stream_), "d2h");
        cuda_chk(cudaStreamSynchronize(stream_), "sync");
    }
    // stored (frame-resolved) index -> logical index
    for (int64_t i = 0; i < shots; ++i) {
        if (result[i] < 0) { result[i] = (int64_t)1 << n_; continue; }   // past the end, as the reference's lower_bound
        const uint64_t ph = (uint64_t)result[i];
        uint64_t lg = 0;
        for (int q = 0; q < n_; ++q) lg |= ((ph >> perm_[q]) & 1ULL) << q;
        result[i] = (int64_t)lg;
    }
    return result;
}

}  // namespace qsim
