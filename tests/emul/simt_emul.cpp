// simt_emul.cpp -- scheduler of the TEST-ONLY SIMT emulator (see simt_emul.h).
#include "simt_emul.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace fdb_emul {

BlockState* g_blk = nullptr;
unsigned char* g_dyn_smem = nullptr;

static const size_t kStack = 256 * 1024;

static void set_tid(unsigned t) {
    threadIdx.x = t % blockDim.x;
    threadIdx.y = (t / blockDim.x) % blockDim.y;
    threadIdx.z = t / (blockDim.x * blockDim.y);
}

void yield() {
    BlockState* b = g_blk;
    unsigned me = b->cur;
    swapcontext(&b->fibers[me].ctx, &b->sched);
    set_tid(me);
}

static void trampoline() {
    BlockState* b = g_blk;
    b->body();
    b->fibers[b->cur].done = true;
    b->live--;
    b->spins = 0;
    // uc_link returns to the scheduler
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    gridDim = grid;
    blockDim = block;
    unsigned nthreads = block.x * block.y * block.z;
    if (blockDim.y != 1 || blockDim.z != 1) {
        fprintf(stderr, "fdb_emul: only 1-D blocks are supported\n");
        abort();
    }
    unsigned char* dyn = (unsigned char*)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
    g_dyn_smem = dyn;
    std::vector<char*> stacks(nthreads);
    for (unsigned t = 0; t < nthreads; t++) stacks[t] = (char*)malloc(kStack);

    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx = uint3{bx, by, bz};
                BlockState blk;
                blk.nthreads = nthreads;
                blk.live = nthreads;
                blk.body = body;
                blk.fibers.resize(nthreads);
                blk.warps.resize((nthreads + 31) / 32);
                for (unsigned t = 0; t < nthreads; t++) blk.warps[t / 32].expected |= 1u << (t % 32);
                g_blk = &blk;
                for (unsigned t = 0; t < nthreads; t++) {
                    Fiber& f = blk.fibers[t];
                    getcontext(&f.ctx);
                    f.stack = stacks[t];
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &blk.sched;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                }
                while (blk.live > 0) {
                    for (unsigned t = 0; t < nthreads; t++) {
                        if (blk.fibers[t].done) continue;
                        blk.cur = t;
                        set_tid(t);
                        swapcontext(&blk.sched, &blk.fibers[t].ctx);
                    }
                }
                g_blk = nullptr;
            }
    for (unsigned t = 0; t < nthreads; t++) free(stacks[t]);
    free(dyn);
    g_dyn_smem = nullptr;
}

}  // namespace fdb_emul
