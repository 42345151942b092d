// cuda_sim.cc - fiber scheduler behind cuda_sim.h.  TEST INFRASTRUCTURE ONLY (see cuda_sim.h).
#include "cuda_sim.h"

#include <sys/mman.h>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace cusim {

Block* g_blk = nullptr;
size_t g_stack_bytes = 192 * 1024;

static std::vector<char*> g_stacks;     // reused across blocks and launches
static std::vector<unsigned> g_warp_alive;
static unsigned long g_progress = 0;

static char* get_stack(size_t i) {
    while (g_stacks.size() <= i) {
        void* p = mmap(nullptr, g_stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("cusim mmap"); abort(); }
        g_stacks.push_back((char*)p);
    }
    return g_stacks[i];
}

void yield() {
    Block* b = g_blk;
    Fiber& f = b->fibers[b->cur];
    swapcontext(&f.ctx, &b->sched);
}

int lane_id() { return g_blk->fibers[g_blk->cur].lin & 31; }
int warp_id() { return g_blk->fibers[g_blk->cur].lin >> 5; }
uint64_t* warp_slots() { return g_blk->xchg.data() + (size_t)warp_id() * 32; }
unsigned active_mask() { return g_warp_alive[warp_id()]; }
unsigned char* dyn_smem() { return g_blk->smem; }

void block_barrier() {
    Block* b = g_blk;
    unsigned long gen = b->bar_gen;
    b->bar_count++;
    for (;;) {
        if (b->bar_gen != gen) return;
        if (b->bar_count >= b->alive) { b->bar_count = 0; b->bar_gen++; g_progress++; return; }
        yield();
    }
}

void warp_barrier(unsigned mask) {
    Block* b = g_blk;
    int w = warp_id();
    unsigned long gen = b->warp_gen[w];
    b->warp_count[w]++;
    for (;;) {
        if (b->warp_gen[w] != gen) return;
        int need = __builtin_popcount(mask & g_warp_alive[w]);
        if (b->warp_count[w] >= need) { b->warp_count[w] = 0; b->warp_gen[w]++; g_progress++; return; }
        yield();
    }
}

static void fiber_entry() {
    Block* b = g_blk;
    (*b->body)();
    Fiber& f = b->fibers[b->cur];
    f.done = true;
    b->alive--;
    g_warp_alive[f.lin >> 5] &= ~(1u << (f.lin & 31));
    g_progress++;
    // returning switches to uc_link (the scheduler)
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) { fprintf(stderr, "cusim: bad block size %d\n", nthreads); abort(); }
    const int nwarps = (nthreads + 31) / 32;
    Block blk;
    blk.nthreads = nthreads;
    blk.fibers.resize(nthreads);
    blk.warp_count.assign(nwarps, 0);
    blk.warp_gen.assign(nwarps, 0);
    blk.xchg.assign((size_t)nwarps * 32, 0);
    std::vector<unsigned char> smem(smem_bytes + 1024);
    // 1024-byte aligned dynamic shared memory, like the hardware's
    blk.smem = (unsigned char*)(((uintptr_t)smem.data() + 1023) & ~(uintptr_t)1023);
    blk.body = &body;
    Block* prev = g_blk;
    g_blk = &blk;
    blockDim = block;
    gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = uint3{bx, by, bz};
        blk.alive = nthreads;
        blk.bar_count = 0;
        std::fill(blk.warp_count.begin(), blk.warp_count.end(), 0);
        g_warp_alive.assign(nwarps, 0);
        for (int i = 0; i < nthreads; ++i) {
            Fiber& f = blk.fibers[i];
            f.done = false;
            f.lin = i;
            f.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
            g_warp_alive[i >> 5] |= 1u << (i & 31);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = get_stack(i);
            f.ctx.uc_stack.ss_size = g_stack_bytes;
            f.ctx.uc_link = &blk.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        int stalled_rounds = 0;
        while (blk.alive > 0) {
            unsigned long before = g_progress;
            for (int i = 0; i < nthreads; ++i) {
                Fiber& f = blk.fibers[i];
                if (f.done) continue;
                blk.cur = i;
                threadIdx = f.tid;
                swapcontext(&blk.sched, &f.ctx);
            }
            if (g_progress == before) {
                if (++stalled_rounds > 4) {
                    fprintf(stderr, "cusim: deadlock in block (%u,%u,%u): %d threads alive, barrier count %d\n",
                            bx, by, bz, blk.alive, blk.bar_count);
                    abort();
                }
            } else {
                stalled_rounds = 0;
            }
        }
    }
    g_blk = prev;
}

}  // namespace cusim
