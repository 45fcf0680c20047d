// cw_host.h -- host side of the CW_MEM_HOST data path: page-locked memory cache, pointer classification and the pool of
// host threads that stages pageable caller buffers through page-locked memory.
//
// The reference moves arguments and results across its process pool by pickling them (cogsworth pop.py:1245-1266); here
// they cross the host link by DMA, and plain numpy arrays (pageable) are first copied into a page-locked ring by several
// host threads, chunk by chunk, while the kernels of the previous chunks run.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace cw {

// ---- page-locked memory cache (cudaHostAlloc, portable): blocks are handed back to a free list, not to the driver
void *pinned_alloc(size_t bytes);             // nullptr on failure
void *pinned_alloc_cached(size_t bytes);      // a block from the cache, or nullptr (no CUDA call)
void pinned_free(void *p);
void pinned_trim();                    // release every cached block
// true when p is page-locked host memory or device / managed memory known to the runtime (= DMA-able as it is)
bool pointer_is_dma_able(const void *p);

// ---- one piece of a staging copy
struct CopyTask {
    void *dst;
    const void *src;
    size_t bytes;
};

// Pool of host threads shared by every context of the process.  run() executes the tasks (large ones are cut into
// pieces) on the workers AND the calling thread and returns when all of them are done.  Several threads may call
// run() at the same time (one per device slot): their pieces share the workers.
class HostPool {
public:
    static HostPool &get();
    void run(const std::vector<CopyTask> &tasks);
    // fn(piece) for piece = 0 .. n_pieces-1, spread over the workers and the calling thread
    void parallel_for(size_t n_pieces, const std::function<void(size_t)> &fn);
    int workers() const;
    struct Impl;
private:
    HostPool();
    ~HostPool();
    Impl *p_;
};

} // namespace cw
