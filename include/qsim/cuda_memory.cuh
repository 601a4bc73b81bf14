// CudaMemory<T>: move-only RAII owner of a cudaMalloc allocation (the reference's
// include/CudaMemory.cuh:49-204 contract: get/size/bytes/empty, copyFromHost/copyToHost/zero,
// invalid_argument on over-long copies, runtime_error on CUDA failures, non-throwing destructor).
#pragma once

#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace qsim {

template <typename T>
class CudaMemory {
public:
    CudaMemory() = default;
    explicit CudaMemory(size_t count) : n_(count) {
        if (n_ == 0) return;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p_), n_ * sizeof(T));
        if (e != cudaSuccess) {
            p_ = nullptr;
            throw std::runtime_error(std::string("CUDA malloc failed: ") + cudaGetErrorString(e));
        }
    }
    ~CudaMemory() { release(); }
    CudaMemory(const CudaMemory&) = delete;
    CudaMemory& operator=(const CudaMemory&) = delete;
    CudaMemory(CudaMemory&& o) noexcept : n_(std::exchange(o.n_, 0)), p_(std::exchange(o.p_, nullptr)) {}
    CudaMemory& operator=(CudaMemory&& o) noexcept {
        if (this != &o) {
            release();
            n_ = std::exchange(o.n_, 0);
            p_ = std::exchange(o.p_, nullptr);
        }
        return *this;
    }

    T* get() { return p_; }
    const T* get() const { return p_; }
    size_t size() const { return n_; }
    size_t bytes() const { return n_ * sizeof(T); }
    bool empty() const { return p_ == nullptr; }

    void copyFromHost(const T* src, size_t count) {
        check_count(count);
        if (count) check(cudaMemcpy(p_, src, count * sizeof(T), cudaMemcpyHostToDevice), "CUDA memcpy H2D failed: ");
    }
    void copyToHost(T* dst, size_t count) const {
        check_count(count);
        if (count) check(cudaMemcpy(dst, p_, count * sizeof(T), cudaMemcpyDeviceToHost), "CUDA memcpy D2H failed: ");
    }
    void copyToHost(std::vector<T>& dst) const {
        dst.resize(n_);
        if (n_) copyToHost(dst.data(), n_);
    }
    void zero() {
        if (p_ && n_) check(cudaMemset(p_, 0, n_ * sizeof(T)), "CUDA memset failed: ");
    }

private:
    size_t n_ = 0;
    T* p_ = nullptr;

    void release() noexcept {
        if (p_) cudaFree(p_);
        p_ = nullptr;
        n_ = 0;
    }
    void check_count(size_t count) const {
        if (count > n_) throw std::invalid_argument("Copy count exceeds allocation size");
    }
    static void check(cudaError_t e, const char* what) {
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + cudaGetErrorString(e));
    }
};

}  // namespace qsim
