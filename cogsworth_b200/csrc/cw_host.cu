// cw_host.cu -- page-locked memory cache, pointer classification and the host staging thread pool (cw_host.h).
#include "cw_host.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <sched.h>
#include <thread>
#include <unordered_map>

#include <cuda_runtime.h>

#include "../../include/cogsworth_b200.h"

namespace cw {

// ---------------------------------------------------------------------------------------------
// page-locked memory cache
// ---------------------------------------------------------------------------------------------
namespace {

struct PinnedCache {
    std::mutex m;
    std::unordered_map<void *, size_t> live;          // blocks handed out
    std::multimap<size_t, void *> free_blocks;        // cached blocks by size
    size_t cached_bytes = 0;
    size_t cap_bytes = 0;
    PinnedCache()
    {
        const char *e = getenv("CW_PINNED_CACHE_MB");
        const long long mb = e ? atoll(e) : 4096;
        cap_bytes = (size_t)(mb > 0 ? mb : 0) << 20;
    }
};

PinnedCache &pinned_cache()
{
    static PinnedCache *c = new PinnedCache();        // never destroyed: blocks may outlive static destruction order
    return *c;
}

constexpr size_t PIN_GRAIN = (size_t)2 << 20;

} // namespace

// a cached block that fits `bytes`, or nullptr: never calls into the CUDA runtime
void *pinned_alloc_cached(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    const size_t want = (bytes + PIN_GRAIN - 1) / PIN_GRAIN * PIN_GRAIN;
    PinnedCache &c = pinned_cache();
    std::lock_guard<std::mutex> lk(c.m);
    auto it = c.free_blocks.lower_bound(want);
    // best fit, but never hand a block that wastes more than a quarter of itself (+ one grain)
    if (it != c.free_blocks.end() && it->first <= want + want / 4 + PIN_GRAIN) {
        void *p = it->second;
        const size_t sz = it->first;
        c.free_blocks.erase(it);
        c.cached_bytes -= sz;
        c.live[p] = sz;
        return p;
    }
    return nullptr;
}

void *pinned_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    const size_t want = (bytes + PIN_GRAIN - 1) / PIN_GRAIN * PIN_GRAIN;
    PinnedCache &c = pinned_cache();
    if (void *hit = pinned_alloc_cached(bytes)) return hit;
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pinned_trim();                                // give the cached blocks back and try once more
        e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    std::lock_guard<std::mutex> lk(c.m);
    c.live[p] = want;
    return p;
}

void pinned_free(void *p)
{
    if (!p) return;
    PinnedCache &c = pinned_cache();
    size_t sz = 0;
    {
        std::lock_guard<std::mutex> lk(c.m);
        auto it = c.live.find(p);
        if (it == c.live.end()) return;               // not ours
        sz = it->second;
        c.live.erase(it);
        if (c.cached_bytes + sz <= c.cap_bytes) {
            c.free_blocks.emplace(sz, p);
            c.cached_bytes += sz;
            return;
        }
    }
    cudaFreeHost(p);
}

void pinned_trim()
{
    PinnedCache &c = pinned_cache();
    std::multimap<size_t, void *> blocks;
    {
        std::lock_guard<std::mutex> lk(c.m);
        blocks.swap(c.free_blocks);
        c.cached_bytes = 0;
    }
    for (auto &b : blocks) cudaFreeHost(b.second);
}

bool pointer_is_dma_able(const void *p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------------------------------------
// staging thread pool
// ---------------------------------------------------------------------------------------------
namespace {

struct Job {
    std::vector<CopyTask> pieces;                      // copy job ...
    const std::function<void(size_t)> *fn = nullptr;   // ... or a parallel loop over n_pieces
    size_t n_pieces = 0;
    std::atomic<size_t> next{0};
    std::atomic<size_t> done{0};
    std::mutex m;
    std::condition_variable cv;
};

constexpr size_t PIECE_BYTES = (size_t)512 << 10;

void work_on(Job &j)
{
    const size_t n = j.n_pieces;
    for (;;) {
        const size_t i = j.next.fetch_add(1, std::memory_order_relaxed);
        if (i >= n) break;
        if (j.fn) {
            (*j.fn)(i);
        } else {
            const CopyTask &t = j.pieces[i];
            memcpy(t.dst, t.src, t.bytes);
        }
        if (j.done.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            std::lock_guard<std::mutex> lk(j.m);
            j.cv.notify_all();
        }
    }
}

} // namespace

struct HostPool::Impl {
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::shared_ptr<Job>> jobs;
    std::vector<std::thread> threads;
    bool stop = false;
    bool started = false;
    int n_workers = 0;

    void start()
    {
        std::lock_guard<std::mutex> lk(m);
        if (started) return;
        started = true;
        for (int i = 0; i < n_workers; i++)
            threads.emplace_back([this] {
                for (;;) {
                    std::shared_ptr<Job> j;
                    {
                        std::unique_lock<std::mutex> lk2(m);
                        cv.wait(lk2, [&] { return stop || !jobs.empty(); });
                        if (stop) return;
                        j = jobs.front();
                    }
                    work_on(*j);
                    std::lock_guard<std::mutex> lk2(m);
                    if (!jobs.empty() && jobs.front() == j) jobs.pop_front();   // exhausted: the next job moves up
                }
            });
    }
};

static void submit_and_wait(HostPool::Impl *p, const std::shared_ptr<Job> &j);

HostPool::HostPool() : p_(new Impl())
{
    int avail = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) avail = CPU_COUNT(&set);
    // staging copies saturate the memory system with a few threads, the compute-bound loops (kick rotation: four libm
    // calls per event) scale with every core: all but one of the cores this process may use, at most 15 workers
    int n = avail - 1;
    if (n > 15) n = 15;
    if (const char *e = getenv("CW_HOST_THREADS")) n = atoi(e) - 1;        // CW_HOST_THREADS counts the caller too
    if (n < 0) n = 0;
    p_->n_workers = n;
}

HostPool::~HostPool()
{
    {
        std::lock_guard<std::mutex> lk(p_->m);
        p_->stop = true;
    }
    p_->cv.notify_all();
    for (std::thread &t : p_->threads) t.join();
    delete p_;
}

HostPool &HostPool::get()
{
    static HostPool *pool = new HostPool();          // leaked on purpose: worker threads must not be joined at exit
    return *pool;
}

int HostPool::workers() const { return p_->n_workers; }

void HostPool::run(const std::vector<CopyTask> &tasks)
{
    size_t total = 0;
    for (const CopyTask &t : tasks) total += t.bytes;
    if (total == 0) return;
    if (p_->n_workers == 0 || total <= PIECE_BYTES) {
        for (const CopyTask &t : tasks)
            if (t.bytes) memcpy(t.dst, t.src, t.bytes);
        return;
    }
    auto j = std::make_shared<Job>();
    for (const CopyTask &t : tasks) {
        for (size_t off = 0; off < t.bytes; off += PIECE_BYTES) {
            CopyTask p;
            p.dst = (char *)t.dst + off;
            p.src = (const char *)t.src + off;
            p.bytes = (t.bytes - off < PIECE_BYTES) ? t.bytes - off : PIECE_BYTES;
            j->pieces.push_back(p);
        }
    }
    j->n_pieces = j->pieces.size();
    submit_and_wait(p_, j);
}

void HostPool::parallel_for(size_t n_pieces, const std::function<void(size_t)> &fn)
{
    if (n_pieces == 0) return;
    if (p_->n_workers == 0 || n_pieces == 1) {
        for (size_t i = 0; i < n_pieces; i++) fn(i);
        return;
    }
    auto j = std::make_shared<Job>();
    j->fn = &fn;
    j->n_pieces = n_pieces;
    submit_and_wait(p_, j);
}

static void submit_and_wait(HostPool::Impl *p, const std::shared_ptr<Job> &j)
{
    p->start();
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->jobs.push_back(j);
    }
    p->cv.notify_all();
    work_on(*j);
    {
        std::unique_lock<std::mutex> lk(j->m);
        j->cv.wait(lk, [&] { return j->done.load(std::memory_order_acquire) == j->n_pieces; });
    }
    std::lock_guard<std::mutex> lk(p->m);
    for (auto it = p->jobs.begin(); it != p->jobs.end(); ++it)
        if (*it == j) { p->jobs.erase(it); break; }
}

} // namespace cw

// get_kick_differential (cogsworth kicks.py:12-49) for K events on the host thread pool: BSE-frame systemic kick
// (km/s), orbital phase and inclination -> Galactocentric velocity change, times `scale` (km/s -> kpc/Myr).
// The vectorised numpy form of cogsworth_b200.batch.rotate_kicks, whose cost is four libm calls per event.
extern "C" void cw_host_rotate_kicks(int64_t K, const double *dv, const double *phase, const double *inc, double scale,
                                     double *out)
{
    if (K <= 0) return;
    const int64_t grain = 1 << 15;
    const size_t pieces = (size_t)((K + grain - 1) / grain);
    std::function<void(size_t)> fn = [&](size_t p) {
        const int64_t a = (int64_t)p * grain, b = std::min<int64_t>(K, a + grain);
        for (int64_t k = a; k < b; k++) {
            double st, ct, sp, cp;
            sincos(phase[k], &st, &ct);
            sincos(inc[k], &sp, &cp);
            const double vx = dv[3 * k], vy = dv[3 * k + 1], vz = dv[3 * k + 2];
            out[3 * k + 0] = (vx * ct - vy * st * cp + vz * st * sp) * scale;
            out[3 * k + 1] = (vx * st + vy * ct * cp - vz * ct * sp) * scale;
            out[3 * k + 2] = (vy * sp + vz * cp) * scale;
        }
    };
    cw::HostPool::get().parallel_for(pieces, fn);
}

extern "C" void *cw_host_alloc(size_t bytes) { return cw::pinned_alloc(bytes); }
extern "C" void *cw_host_alloc_cached(size_t bytes) { return cw::pinned_alloc_cached(bytes); }
extern "C" void cw_host_free(void *p) { cw::pinned_free(p); }
extern "C" void cw_host_trim(void) { cw::pinned_trim(); }
extern "C" int cw_host_is_pinned(const void *p) { return (p && cw::pointer_is_dma_able(p)) ? 1 : 0; }
