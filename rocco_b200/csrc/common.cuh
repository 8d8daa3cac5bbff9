// Shared helpers for the rocco_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/rocco_b200.h"

namespace rb {

// max / min of two doubles as ONE compare and a select.  fmax()/fmin() cost about eight instructions each on sm_100
// (DSETP.MAX + FSEL + SEL + NaN quieting + moves); these keep fmax/fmin's result whenever `b` is not NaN -- a NaN in
// `a` gives b, exactly like fmax/fmin -- which is how every call site uses them (b is a constant or an already clamped value).
#ifdef __CUDACC__
// (inline PTX: written as `a > b ? a : b` the compiler recognises the idiom and emits the fmax sequence again)
__device__ __forceinline__ double dmax(double a, double b)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double dmin(double a, double b)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
#endif

constexpr int ST_OK = 0, ST_NOMEM = -1, ST_INVALID = -2, ST_CUDA = -3, ST_NONFINITE = -4;

void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int k = 1) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }

#define RB_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::rb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? ::rb::ST_NOMEM : ::rb::ST_CUDA;       \
        }                                                                                    \
    } while (0)

#define RB_LAUNCH_CHECK()                                                                    \
    do {                                                                                     \
        ::rb::count_launch();                                                                \
        RB_CUDA(cudaGetLastError());                                                         \
    } while (0)

#define RB_TRY(expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s != 0) return _s;       \
    } while (0)

// Stream-ordered scratch: every buffer is freed (stream-ordered) when the arena dies.
// Memory pool private to host_stream() of the calling thread (nullptr for any other stream: default pool).  A private
// pool keeps a thread's scratch from being recycled into another thread's stream, which would chain the two
// streams together through the allocator's internal dependencies.
cudaMemPool_t pool_for_stream(cudaStream_t s);

class Arena {
  public:
    explicit Arena(cudaStream_t s) : stream_(s), pool_(pool_for_stream(s)) {}
    ~Arena() { release(); }
    Arena(const Arena &) = delete;
    Arena &operator=(const Arena &) = delete;
    template <typename T> int alloc(T **out, size_t count) {
        void *p = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        cudaError_t e = pool_ ? cudaMallocFromPoolAsync(&p, bytes, pool_, stream_) : cudaMallocAsync(&p, bytes, stream_);
        if (e != cudaSuccess) {
            set_error("cudaMallocAsync(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
            (void)cudaGetLastError();
            *out = nullptr;
            return e == cudaErrorMemoryAllocation ? ST_NOMEM : ST_CUDA;
        }
        ptrs_.push_back(p);
        *out = static_cast<T *>(p);
        return 0;
    }
    void release() {
        for (void *p : ptrs_) cudaFreeAsync(p, stream_);
        ptrs_.clear();
    }

  private:
    cudaStream_t stream_;
    cudaMemPool_t pool_;
    std::vector<void *> ptrs_;
};

// Pinned host staging that is released on scope exit.
template <typename T> class Pinned {
  public:
    Pinned() = default;
    ~Pinned() { if (p_) cudaFreeHost(p_); }
    int alloc(size_t count) {
        cudaError_t e = cudaMallocHost(&p_, (count ? count : 1) * sizeof(T));
        if (e != cudaSuccess) {
            set_error("cudaMallocHost -> %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
            p_ = nullptr;
            return ST_NOMEM;
        }
        return 0;
    }
    T *get() { return static_cast<T *>(p_); }
    T &operator[](size_t i) { return static_cast<T *>(p_)[i]; }

  private:
    void *p_ = nullptr;
};

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
// Disabled by default: a scope then costs one relaxed atomic load.
class ProfScope {
  public:
    ProfScope(const char *name, cudaStream_t st, double bytes = 0.0);
    ~ProfScope();
  private:
    int idx_;
    cudaStream_t st_;
};
#define RB_PROF(name, st, bytes) ::rb::ProfScope _prof_scope_##__LINE__(name, st, bytes)

int ensure_device();     // returns 0 or ST_CUDA when no usable device
// Stream used by the host-pointer entries: one non-blocking stream per host thread and device, so that two host
// threads driving different chromosomes overlap (one's H2D copy with the other's kernels) instead of queueing
// behind each other on the legacy default stream.
cudaStream_t host_stream();
// RAII lease of a (non-blocking stream, private memory pool) pair for one host-pointer call.  Pairs are recycled
// through a global free list, so the number that exist equals the highest concurrency seen, whatever the threads do.
// `urgent` leases carry the device's highest stream priority: the short launch chains of a solve or a mask->runs pass
// then get SM slots ahead of the bulk scoring kernels another host thread has in flight, instead of queueing behind them.
class HostScope {
  public:
    explicit HostScope(bool urgent = false, size_t scratch_hint = 0);
    ~HostScope();
    HostScope(const HostScope &) = delete;
    HostScope &operator=(const HostScope &) = delete;
    cudaStream_t stream() const { return st_; }
    // Pinned bounce buffer owned by the lease (grown on demand, kept for the next holder); nullptr when pinning fails.
    // A pageable upload is split by the driver into many small DMA commands, each of which waits its turn behind the
    // bulk copies other host threads have queued; one pinned transfer waits once.
    void *staging(size_t bytes);
  private:
    int slot_;
    cudaStream_t st_;
};
int sm_count();
// Kernel-driven upload from pinned (device-mapped) host memory; see runtime.cu.
int pull_from_pinned(void *d_dst, const void *h_pinned, size_t bytes, cudaStream_t st);

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    return __shfl_up_sync(0xffffffffu, v, delta);
}
__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __forceinline__ void st_release_i32(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_i32(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace rb

namespace rb {
// numpy.sum of a contiguous float64 vector, restated (pairwise halving down to blocks of <= 128 summed with 8
// accumulators): dp.py:110-111 builds the search bracket from numpy.sum(switch_costs).
double numpy_sum_f64(const double *a, size_t n);
double numpy_sum_const_f64(double value, size_t n);
}  // namespace rb
