// cuda_emul.cpp -- fiber scheduler of the host emulation (TEST INFRASTRUCTURE ONLY)
#include "cuda_emul.h"
#include <sys/mman.h>

namespace emul {
Block *B = nullptr;

static void trampoline()
{
    B->body();
    me().done = true;
    swapcontext(&me().uc, &B->sched);
}

void launch(Idx3 grid, unsigned block_threads, size_t smem_bytes, std::function<void()> body)
{
    if (block_threads % 32) { fprintf(stderr, "emul: block size must be a multiple of 32\n"); abort(); }
    const size_t STACK = 512 * 1024;
    char *stacks = (char *)mmap(nullptr, STACK * block_threads, PROT_READ | PROT_WRITE,
                                MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == (char *)MAP_FAILED) { perror("mmap"); abort(); }
    std::vector<char> smem(smem_bytes + 64);
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk;
        B = &blk;
        blk.bidx = Idx3{bx, by, bz};
        blk.bdim = Idx3{block_threads, 1, 1};
        blk.gdim = grid;
        blk.smem = smem.data();
        blk.smem_bytes = smem_bytes;
        memset(blk.smem, 0xcd, smem_bytes);      // poison: uninitialised shared memory shows
        blk.body = body;
        blk.f.resize(block_threads);
        blk.w.resize(block_threads / 32);
        for (auto &w : blk.w) memset(&w, 0, sizeof(w));
        for (unsigned t = 0; t < block_threads; ++t) {
            Fiber &f = blk.f[t];
            f.tid = (int)t;
            getcontext(&f.uc);
            f.uc.uc_stack.ss_sp = stacks + STACK * t;
            f.uc.uc_stack.ss_size = STACK;
            f.uc.uc_link = nullptr;
            makecontext(&f.uc, (void (*)())trampoline, 0);
        }
        unsigned live = block_threads;
        uint64_t last_progress = 0;
        int idle_rounds = 0;
        while (live) {
            for (unsigned t = 0; t < block_threads; ++t) {
                if (blk.f[t].done) continue;
                blk.cur = (int)t;
                swapcontext(&blk.sched, &blk.f[t].uc);
                if (blk.f[t].done) --live;
            }
            if (blk.progress == last_progress && live) {
                if (++idle_rounds > 4) {
                    fprintf(stderr, "emul: deadlock (divergent collective?) in block %u, %u fibers live\n", bx, live);
                    abort();
                }
            } else {
                idle_rounds = 0;
                last_progress = blk.progress;
            }
        }
        B = nullptr;
    }
    munmap(stacks, STACK * block_threads);
}
}  // namespace emul
