// Library plumbing: error text, device selection, launch counter, numpy-compatible summation.
#include <thread>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"

#include <stdarg.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include <cstring>
#include <mutex>

namespace rb {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); rocco_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return ST_CUDA;
    }
    // keep stream-ordered scratch cached in the pool between calls (default threshold 0 hands every
    // buffer back to the driver at each synchronisation, which costs more than the kernels)
    static std::atomic<unsigned> tuned{0};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev < 32 && !(tuned.load() & (1u << dev))) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        tuned.fetch_or(1u << dev);
        (void)cudaGetLastError();
    }
    return 0;
}

struct HostCtx { int dev; cudaStream_t st; cudaMemPool_t pool; bool busy; int urgent; void *stage; size_t stage_bytes; };
static std::mutex g_ctx_mu;
static std::vector<HostCtx> g_ctx;

HostScope::HostScope(bool urgent, size_t scratch_hint) : slot_(-1), st_(nullptr)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    // Best fit on the pools' reserved sizes: growing a pool means creating and mapping physical memory (~50 ms per GB,
    // under driver-wide locks that stall every other host thread), so a call takes the free lease whose pool already
    // holds at least `scratch_hint` bytes and is the smallest such, else the largest one.
    int best = -1;
    unsigned long long best_res = 0;
    for (size_t k = 0; k < g_ctx.size(); ++k) {
        if (g_ctx[k].busy || g_ctx[k].dev != dev || g_ctx[k].urgent != (int)urgent) continue;
        unsigned long long res = 0;
        if (g_ctx[k].pool && cudaMemPoolGetAttribute(g_ctx[k].pool, cudaMemPoolAttrReservedMemCurrent, &res) != cudaSuccess) {
            (void)cudaGetLastError(); res = 0;
        }
        const bool fits = res >= scratch_hint, best_fits = best >= 0 && best_res >= scratch_hint;
        if (best < 0 || (fits && (!best_fits || res < best_res)) || (!fits && !best_fits && res > best_res)) { best = (int)k; best_res = res; }
    }
    if (best >= 0) { g_ctx[best].busy = true; slot_ = best; st_ = g_ctx[best].st; return; }
    HostCtx c{dev, nullptr, nullptr, true, (int)urgent, nullptr, 0};
    int least = 0, greatest = 0;
    if (urgent && cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { (void)cudaGetLastError(); greatest = 0; }
    if (cudaStreamCreateWithPriority(&c.st, cudaStreamNonBlocking, urgent ? greatest : 0) != cudaSuccess) { (void)cudaGetLastError(); return; }
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&c.pool, &props) == cudaSuccess) {
        unsigned long long thr = ~0ULL;
        cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else { (void)cudaGetLastError(); c.pool = nullptr; }
    g_ctx.push_back(c);
    slot_ = (int)g_ctx.size() - 1;
    st_ = c.st;
}

void *HostScope::staging(size_t bytes)
{
    if (slot_ < 0) return nullptr;
    void *cur = nullptr;
    size_t have = 0;
    { std::lock_guard<std::mutex> lk(g_ctx_mu); cur = g_ctx[slot_].stage; have = g_ctx[slot_].stage_bytes; }
    if (have >= bytes) return cur;
    if (cur) cudaFreeHost(cur);
    void *p = nullptr;
    // generous and sticky: (re)allocating pinned memory synchronises the device and stalls every other host thread, so a
    // lease's buffer is sized once for the largest mask a genome produces (64 MB) and only ever grows past that
    const size_t want = std::max<size_t>(bytes + bytes / 4, (size_t)64 << 20);
    if (cudaMallocHost(&p, want) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; }
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    g_ctx[slot_].stage = p;
    g_ctx[slot_].stage_bytes = p ? want : 0;
    return p;
}

HostScope::~HostScope()
{
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    g_ctx[slot_].busy = false;
}

cudaStream_t host_stream() { return nullptr; }      // legacy default stream (kept for callers without a lease)

// Every stream the library is called on gets its own memory pool (up to 16 caller streams; the legacy default
// stream keeps the device's default pool): scratch freed on one stream is then never recycled into another
// stream, which would chain the two streams together through the allocator's internal dependencies and
// serialise callers that drive different chromosomes from different host threads.
static std::vector<HostCtx> g_user_ctx;

cudaMemPool_t pool_for_stream(cudaStream_t s)
{
    if (s == nullptr) return nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (const HostCtx &c : g_ctx)
        if (c.st == s) return c.pool;
    for (const HostCtx &c : g_user_ctx)
        if (c.st == s && c.dev == dev) return c.pool;
    if (g_user_ctx.size() >= 16) return nullptr;
    HostCtx c{dev, s, nullptr, true, 0, nullptr, 0};
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&c.pool, &props) == cudaSuccess) {
        unsigned long long thr = ~0ULL;
        cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else { (void)cudaGetLastError(); c.pool = nullptr; }
    g_user_ctx.push_back(c);
    return c.pool;
}

// Upload from PINNED host memory by a kernel that reads it over PCIe (zero-copy) instead of a DMA command.  The copy
// engine serves H2D commands in submission order across all streams, so a 40 MB score vector submitted while other host
// threads have gigabytes of count matrices queued would wait for that whole backlog; SM loads do not queue behind it.
__global__ void __launch_bounds__(256) k_pull(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16,
                                              unsigned char *dst_tail, const unsigned char *src_tail, int tail)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

int pull_from_pinned(void *d_dst, const void *h_pinned, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(d_dst) | reinterpret_cast<uintptr_t>(h_pinned)) & 15) {
        RB_CUDA(cudaMemcpyAsync(d_dst, h_pinned, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    const size_t n16 = bytes / 16;
    const int tail = (int)(bytes - n16 * 16);
    const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256 + 1, (size_t)sm_count() * 8);
    k_pull<<<grid, 256, 0, st>>>(static_cast<uint4 *>(d_dst), static_cast<const uint4 *>(h_pinned), n16,
                                 static_cast<unsigned char *>(d_dst) + n16 * 16,
                                 static_cast<const unsigned char *>(h_pinned) + n16 * 16, tail);
    RB_LAUNCH_CHECK();
    return 0;
}

int sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

// ---------------------------------------------------------------- event profiler
struct ProfEntry { std::string name; cudaEvent_t a, b; double bytes; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;

static std::vector<cudaEvent_t> g_prof_pool;      // recycled events: creating them inside the timed region is not free

ProfScope::ProfScope(const char *name, cudaStream_t st, double bytes) : idx_(-1), st_(st)
{
    nvtxRangePushA(name);                         // every scope is also an NVTX range (a no-op unless a tool is attached)
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfEntry e;
    e.name = name; e.bytes = bytes;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (cudaEvent_t *ev : {&e.a, &e.b}) {
        if (!g_prof_pool.empty()) { *ev = g_prof_pool.back(); g_prof_pool.pop_back(); }
        else if (cudaEventCreate(ev) != cudaSuccess) { (void)cudaGetLastError(); return; }
    }
    cudaEventRecord(e.a, st);
    g_prof.push_back(e);
    idx_ = (int)g_prof.size() - 1;
}

ProfScope::~ProfScope()
{
    nvtxRangePop();
    if (idx_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (idx_ < (int)g_prof.size()) cudaEventRecord(g_prof[idx_].b, st_);
}

// numpy/_core/src/umath/loops_utils.h.src  DOUBLE_pairwise_sum, restated
static double pairwise(const double *a, size_t n)
{
    if (n < 8) {
        double r = 0.0;
        for (size_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        size_t i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise(a, n2) + pairwise(a + n2, n - n2);
}

double numpy_sum_f64(const double *a, size_t n)
{
    if (n == 0) return 0.0;
    return pairwise(a, n);
}

static double pairwise_const(double v, size_t n, std::map<size_t, double> &memo)
{
    if (n <= 128) {
        double buf[128];
        for (size_t i = 0; i < n; ++i) buf[i] = v;
        return pairwise(buf, n);
    }
    auto it = memo.find(n);
    if (it != memo.end()) return it->second;
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    const double r = pairwise_const(v, n2, memo) + pairwise_const(v, n - n2, memo);
    memo[n] = r;
    return r;
}

double numpy_sum_const_f64(double value, size_t n)
{
    if (n == 0) return 0.0;
    std::map<size_t, double> memo;
    return pairwise_const(value, n, memo);
}

}  // namespace rb

#define RB_API __attribute__((visibility("default")))
extern "C" {

RB_API const char *rocco_b200_version(void) { return "0.1.0"; }
RB_API const char *rocco_b200_last_error(void) { return rb::g_err; }

RB_API int rocco_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

RB_API int rocco_b200_set_device(int device)
{
    RB_TRY(rb::ensure_device());
    RB_CUDA(cudaSetDevice(device));
    return 0;
}

RB_API unsigned long long rocco_b200_kernel_launches(void) { return rb::g_launches.load(); }

/* Scratch comes from stream-ordered pools that keep what they have freed (growing a pool costs ~50 ms/GB under driver-wide
 * locks), so an idle library can sit on tens of GB that neither torch's caching allocator nor another workload in the
 * process can see.  This hands everything above `bytes_to_keep` per pool back to the driver (all pools of the current
 * device: the per-call leases, the per-caller-stream pools and the device's default pool).  Call it between workloads, or
 * after a MemoryError before retrying.  Returns the number of pools trimmed. */
RB_API int rocco_b200_trim_pools(size_t bytes_to_keep)
{
    if (rb::ensure_device() != 0) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    cudaDeviceSynchronize();                      // stream-ordered frees still in flight become trimmable
    int trimmed = 0;
    std::lock_guard<std::mutex> lk(rb::g_ctx_mu);
    auto trim = [&](cudaMemPool_t pool) {
        if (pool && cudaMemPoolTrimTo(pool, bytes_to_keep) == cudaSuccess) ++trimmed; else (void)cudaGetLastError();
    };
    for (const rb::HostCtx &c : rb::g_ctx) if (c.dev == dev && !c.busy) trim(c.pool);
    for (const rb::HostCtx &c : rb::g_user_ctx) if (c.dev == dev) trim(c.pool);
    cudaMemPool_t def = nullptr;
    if (cudaDeviceGetDefaultMemPool(&def, dev) == cudaSuccess) trim(def); else (void)cudaGetLastError();
    return trimmed;
}

RB_API int rocco_b200_profile_enable(int on)
{
    return rb::g_prof_on.exchange(on ? 1 : 0);
}

/* Writes one line per scope name: "<name> <total_ms> <count> <total_bytes>\n"; clears the log. */
RB_API int rocco_b200_profile_report(char *buf, size_t cap)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(rb::g_prof_mu);
    std::map<std::string, std::pair<double, std::pair<long long, double>>> agg;
    for (auto &e : rb::g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            auto &a = agg[e.name];
            a.first += ms; a.second.first += 1; a.second.second += e.bytes;
        }
        rb::g_prof_pool.push_back(e.a); rb::g_prof_pool.push_back(e.b);
    }
    (void)cudaGetLastError();
    rb::g_prof.clear();
    size_t off = 0;
    for (auto &kv : agg) {
        int w = snprintf(buf ? buf + off : nullptr, buf && cap > off ? cap - off : 0, "%s %.6f %lld %.0f\n",
                         kv.first.c_str(), kv.second.first, kv.second.second.first, kv.second.second.second);
        if (w < 0) break;
        off += (size_t)w;
        if (buf && off >= cap) { off = cap ? cap - 1 : 0; break; }
    }
    if (buf && cap) buf[off < cap ? off : cap - 1] = 0;
    return (int)off;
}

/* Writes one line per recorded scope in record order: "<name> <start_ms> <duration_ms>\n", start relative to the first
   scope's start (all scopes must be on one stream for the offsets to mean anything); does not clear the log. */
RB_API int rocco_b200_profile_timeline(char *buf, size_t cap)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(rb::g_prof_mu);
    size_t off = 0;
    for (auto &e : rb::g_prof) {
        float t0 = 0.f, ms = 0.f;
        if (cudaEventElapsedTime(&t0, rb::g_prof[0].a, e.a) != cudaSuccess || cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) continue;
        int w = snprintf(buf ? buf + off : nullptr, buf && cap > off ? cap - off : 0, "%s %.4f %.4f\n", e.name.c_str(), t0, ms);
        if (w < 0) break;
        off += (size_t)w;
        if (buf && off >= cap) { off = cap ? cap - 1 : 0; break; }
    }
    (void)cudaGetLastError();
    if (buf && cap) buf[off < cap ? off : cap - 1] = 0;
    return (int)off;
}

/* 1 when v[i+1]-v[i] is the same for all i (the reference's `len(set(np.diff(intervals))) > 1` test, rocco.py:170-172,
 * without materialising the differences); host helper. */
RB_API int rocco_b200_uniform_step_i64(const long long *v, size_t n)
{
    if (!v || n < 3) return 1;
    const long long step = v[1] - v[0];
    int ok = 1;
    for (size_t i = 2; i < n; ++i) ok &= ((v[i] - v[i - 1]) == step);
    return ok;
}

/* BED3 (or BED4 with chrom_start_end names) text of n records straight to a file: the host half of rocco.py:98-110
 * (_write_bed_records) without a Python loop per record.  names[name_idx[i]] (name_idx == NULL: names[0]). */
static const char kDigitPairs[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
static inline char *put_ll(char *p, long long v)
{
    unsigned long long u;
    if (v < 0) { *p++ = '-'; u = 0ULL - (unsigned long long)v; } else u = (unsigned long long)v;
    // digit count first, then two digits at a time from the back (genomic coordinates: 32-bit arithmetic almost always)
    int nd = 1;
    for (unsigned long long t = u; t >= 10; t /= 10) ++nd;
    char *q = p + nd;
    if (u <= 0xFFFFFFFFULL) {
        unsigned w = (unsigned)u;
        while (w >= 100) { const unsigned r = w % 100; w /= 100; q -= 2; q[0] = kDigitPairs[2 * r]; q[1] = kDigitPairs[2 * r + 1]; }
        if (w >= 10) { q -= 2; q[0] = kDigitPairs[2 * w]; q[1] = kDigitPairs[2 * w + 1]; } else *--q = (char)('0' + w);
    } else {
        while (u >= 100) { const unsigned r = (unsigned)(u % 100); u /= 100; q -= 2; q[0] = kDigitPairs[2 * r]; q[1] = kDigitPairs[2 * r + 1]; }
        if (u >= 10) { q -= 2; q[0] = kDigitPairs[2 * u]; q[1] = kDigitPairs[2 * u + 1]; } else *--q = (char)('0' + u);
    }
    return p + nd;
}

}  // extern "C"
namespace {
// BED text of n records, formatted by a few host threads (one contiguous slice of records each) into one buffer kept
// between calls (touching fresh pages costs more than the formatting itself).  `emit(ptr, bytes)` is called for the
// slices in record order while the buffer is held.
template <typename Emit>
static int format_bed3(const char *const *names, int n_names, const int *name_idx, const long long *starts, const long long *ends,
                       size_t n, int name_features, Emit emit)
{
    std::vector<size_t> len((size_t)n_names);
    size_t maxlen = 0;
    for (int k = 0; k < n_names; ++k) { len[k] = strlen(names[k]); maxlen = std::max(maxlen, len[k]); }
    const size_t per = (name_features ? 2 : 1) * (maxlen + 2 * 20 + 2) + 2;      // upper bound of one record's text
    for (size_t i = 0; i < n; ++i) {
        const int c = name_idx ? name_idx[i] : 0;
        if (c < 0 || c >= n_names) return rb::ST_INVALID;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>({(size_t)8, (size_t)(hw ? hw : 1), n / 8192 + 1}));
    static std::mutex text_mu;
    static char *text = nullptr;
    static size_t text_cap = 0;
    std::lock_guard<std::mutex> text_lock(text_mu);
    if (per * n + 1 > text_cap) {
        free(text);
        text_cap = per * n + 1 + (per * n) / 4;
        text = (char *)malloc(text_cap);
        if (!text) { text_cap = 0; return rb::ST_NOMEM; }
    }
    std::vector<size_t> used(nthreads, 0);
    auto work = [&](size_t t) {
        const size_t i0 = n * t / nthreads, i1 = n * (t + 1) / nthreads;
        char *const base = text + per * i0;
        char *p = base;
        for (size_t i = i0; i < i1; ++i) {
            const int c = name_idx ? name_idx[i] : 0;
            memcpy(p, names[c], len[c]); p += len[c];
            *p++ = '\t'; p = put_ll(p, starts[i]);
            *p++ = '\t'; p = put_ll(p, ends[i]);
            if (name_features) {
                *p++ = '\t';
                memcpy(p, names[c], len[c]); p += len[c];
                *p++ = '_'; p = put_ll(p, starts[i]);
                *p++ = '_'; p = put_ll(p, ends[i]);
            }
            *p++ = '\n';
        }
        used[t] = (size_t)(p - base);
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (size_t t = 0; t < nthreads; ++t)
        if (used[t] && !emit(text + per * (n * t / nthreads), used[t])) return rb::ST_INVALID;
    return 0;
}
}  // namespace
extern "C" {

RB_API int rocco_b200_write_bed3(const char *path, const char *const *names, int n_names, const int *name_idx,
                                 const long long *starts, const long long *ends, size_t n, int name_features)
{
    if (!path || !names || n_names <= 0 || (n && (!starts || !ends))) return rb::ST_INVALID;
    FILE *fh = fopen(path, "wb");
    if (!fh) { rb::set_error("cannot open %s for writing", path); return rb::ST_INVALID; }
    const int st = format_bed3(names, n_names, name_idx, starts, ends, n, name_features,
                               [&](const char *p, size_t bytes) { return fwrite(p, 1, bytes, fh) == bytes; });
    fclose(fh);
    return st;
}

/* The same text written INTO an existing (or new) file at byte `offset`, without truncating it: several processes that
 * know each other's text sizes assemble one BED file with no gather pass.  *bytes_written returns the text size. */
RB_API int rocco_b200_write_bed3_at(const char *path, long long offset, const char *const *names, int n_names, const int *name_idx,
                                    const long long *starts, const long long *ends, size_t n, int name_features,
                                    long long *bytes_written)
{
    if (!path || offset < 0 || !names || n_names <= 0 || (n && (!starts || !ends))) return rb::ST_INVALID;
    const int fd = open(path, O_WRONLY | O_CREAT, 0644);
    if (fd < 0) { rb::set_error("cannot open %s for writing", path); return rb::ST_INVALID; }
    long long pos = offset;
    const int st = format_bed3(names, n_names, name_idx, starts, ends, n, name_features, [&](const char *p, size_t bytes) {
        while (bytes) {
            const ssize_t w = pwrite(fd, p, bytes, (off_t)pos);
            if (w <= 0) return false;
            p += w; bytes -= (size_t)w; pos += w;
        }
        return true;
    });
    close(fd);
    if (bytes_written) *bytes_written = pos - offset;
    return st;
}

/* combine_chrom_results (rocco.py:194-240) for canonical BED text, without a Python object per record: read every file,
 * order records by (chrom bytes, start, end) -- bytewise order of UTF-8/ASCII names is Python's str order --, merge
 * records with start <= previous end on the same chrom, write BED3/BED4.  Returns the number of records written, or
 * -1 when any file is not canonical (the caller then takes the reference-faithful line-by-line reader, which owns the
 * error behaviour): canonical = every non-empty line is  <chrom>\t<digits>\t<digits>[\t...]  with ASCII chrom, no
 * leading/trailing whitespace, no carriage returns, at most 18 digits. */
namespace {
struct BedRec { int name; long long start, end; };
static bool is_space_py(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31) || c == 0x85 || c == 0xA0; }
static bool parse_digits(const char *&p, const char *end, long long *out)
{
    const char *q = p;
    unsigned long long v = 0;
    while (q < end && *q >= '0' && *q <= '9') { v = v * 10ULL + (unsigned long long)(*q - '0'); ++q; }
    const size_t nd = (size_t)(q - p);
    if (nd == 0 || nd > 18) return false;
    *out = (long long)v;
    p = q;
    return true;
}
}  // namespace

RB_API long long rocco_b200_combine_bed3(const char *const *paths, int n_paths, const char *out_path, int name_features,
                                         int *saw_extra_columns)
{
    if (!paths || n_paths < 0 || !out_path) return rb::ST_INVALID;
    std::vector<std::string> names;
    std::map<std::string, int> name_id;
    std::vector<BedRec> recs;
    int extra = 0;
    std::vector<char> buf;
    for (int f = 0; f < n_paths; ++f) {
        FILE *fh = fopen(paths[f], "rb");
        if (!fh) return -1;
        fseek(fh, 0, SEEK_END);
        const long sz = ftell(fh);
        fseek(fh, 0, SEEK_SET);
        if (sz < 0) { fclose(fh); return -1; }
        buf.resize((size_t)sz + 1);
        if (sz && fread(buf.data(), 1, (size_t)sz, fh) != (size_t)sz) { fclose(fh); return -1; }
        fclose(fh);
        const char *p = buf.data(), *end = p + sz;
        if (memchr(p, '\r', (size_t)sz)) return -1;
        int last_id = -1;
        const char *last_name = nullptr;
        size_t last_len = 0;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = nl ? nl : end;
            if (le == p) { p = le + 1; continue; }                                  // empty line
            if (is_space_py((unsigned char)*p) || is_space_py((unsigned char)le[-1]) || (unsigned char)le[-1] >= 0x80) return -1;
            const char *t1 = (const char *)memchr(p, '\t', (size_t)(le - p));
            if (!t1 || t1 == p) return -1;
            for (const char *q = p; q < t1; ++q) if ((unsigned char)*q >= 0x80) return -1;
            const char *q = t1 + 1;
            BedRec r;
            if (!parse_digits(q, le, &r.start) || q >= le || *q != '\t') return -1;
            ++q;
            if (!parse_digits(q, le, &r.end)) return -1;
            if (q < le) { if (*q != '\t') return -1; extra = 1; }
            const size_t nlen = (size_t)(t1 - p);
            if (last_id >= 0 && nlen == last_len && memcmp(p, last_name, nlen) == 0) r.name = last_id;
            else {
                std::string nm(p, nlen);
                auto it = name_id.find(nm);
                if (it == name_id.end()) { it = name_id.emplace(nm, (int)names.size()).first; names.push_back(nm); }
                r.name = it->second;
                last_id = r.name; last_name = names[(size_t)r.name].data(); last_len = nlen;
            }
            recs.push_back(r);
            p = le + 1;
        }
    }
    if (saw_extra_columns) *saw_extra_columns = extra;
    // rank of every name in sorted (bytewise) order; std::map iterates in that order
    std::vector<int> rank(names.size());
    std::vector<const std::string *> by_rank(names.size());
    { int k = 0; for (auto &kv : name_id) { rank[(size_t)kv.second] = k; by_rank[(size_t)k] = &kv.first; ++k; } }
    for (auto &r : recs) r.name = rank[(size_t)r.name];
    auto less = [](const BedRec &a, const BedRec &b) {
        if (a.name != b.name) return a.name < b.name;
        if (a.start != b.start) return a.start < b.start;
        return a.end < b.end;
    };
    if (!std::is_sorted(recs.begin(), recs.end(), less)) std::stable_sort(recs.begin(), recs.end(), less);
    size_t w = 0;
    for (size_t i = 0; i < recs.size(); ++i) {
        if (w && recs[i].name == recs[w - 1].name && recs[i].start <= recs[w - 1].end) {
            if (recs[i].end > recs[w - 1].end) recs[w - 1].end = recs[i].end;
        } else recs[w++] = recs[i];
    }
    recs.resize(w);
    FILE *fo = fopen(out_path, "wb");
    if (!fo) { rb::set_error("cannot open %s for writing", out_path); return rb::ST_INVALID; }
    size_t maxlen = 0;
    for (auto &nm : names) maxlen = std::max(maxlen, nm.size());
    const size_t per = 2 * maxlen + 4 * 21 + 8, chunk = 1 << 16;
    std::vector<char> ob(per * chunk);
    for (size_t i0 = 0; i0 < recs.size(); i0 += chunk) {
        char *p = ob.data();
        const size_t i1 = std::min(recs.size(), i0 + chunk);
        for (size_t i = i0; i < i1; ++i) {
            const std::string &nm = *by_rank[(size_t)recs[i].name];
            memcpy(p, nm.data(), nm.size()); p += nm.size();
            *p++ = '\t'; p = put_ll(p, recs[i].start);
            *p++ = '\t'; p = put_ll(p, recs[i].end);
            if (name_features) {
                *p++ = '\t';
                memcpy(p, nm.data(), nm.size()); p += nm.size();
                *p++ = '_'; p = put_ll(p, recs[i].start);
                *p++ = '_'; p = put_ll(p, recs[i].end);
            }
            *p++ = '\n';
        }
        if (fwrite(ob.data(), 1, (size_t)(p - ob.data()), fo) != (size_t)(p - ob.data())) { fclose(fo); return rb::ST_INVALID; }
    }
    fclose(fo);
    return (long long)recs.size();
}

RB_API int rocco_b200_pull_pinned(void *d_dst, const void *h_pinned, size_t bytes, void *cuda_stream)
{
    return rb::pull_from_pinned(d_dst, h_pinned, bytes, (cudaStream_t)cuda_stream);
}

RB_API double rocco_b200_numpy_sum_f64(const double *a, size_t n) { return rb::numpy_sum_f64(a, n); }
RB_API double rocco_b200_numpy_sum_const_f64(double value, size_t n) { return rb::numpy_sum_const_f64(value, n); }
}
