// cuda_sim.cpp — see cuda_sim.h (TEST INFRASTRUCTURE): fiber-per-CUDA-thread execution of the kernel sources on the CPU.
#include "cuda_sim.h"

#include <sys/mman.h>
#include <ucontext.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

thread_local sim_dim3 threadIdx, blockIdx;
sim_dim3 blockDim, gridDim;
thread_local sim_bar sim_block_barrier;
thread_local sim_warp_area* sim_warps = nullptr;
thread_local unsigned long long sim_progress = 0;

namespace {

// x86-64: a switch exchanges the callee-saved registers and the stack pointer (no signal-mask system call as in swapcontext)
#if defined(__x86_64__)
#define SIM_FIBER_ASM 1
extern "C" void sim_fiber_switch(void** save_sp, void* new_sp);
__asm__(
    ".text\n"
    ".globl sim_fiber_switch\n"
    ".type sim_fiber_switch,@function\n"
    "sim_fiber_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size sim_fiber_switch,.-sim_fiber_switch\n");
#endif

struct Fiber {
#ifdef SIM_FIBER_ASM
    void* sp = nullptr;
#else
    ucontext_t uc;
#endif
    char* stack = nullptr;
    unsigned tid = 0;
    bool done = false;
};

struct BlockRunner {
    std::vector<Fiber> fibers;
    char* slab = nullptr;
    size_t slab_bytes = 0, stack_bytes = 0;
    Fiber* cur = nullptr;
#ifdef SIM_FIBER_ASM
    void* sched_sp = nullptr;
#else
    ucontext_t sched;
#endif
    const std::function<void()>* body = nullptr;
};
thread_local BlockRunner* tl_runner = nullptr;

void fiber_main() {
    BlockRunner* r = tl_runner;
    Fiber* f = r->cur;
    (*r->body)();
    f->done = true;
    ++sim_progress;
#ifdef SIM_FIBER_ASM
    sim_fiber_switch(&f->sp, r->sched_sp);
    __builtin_trap();
#endif
}

void run_block(BlockRunner& R, unsigned block, unsigned b, const std::function<void()>& body) {
    tl_runner = &R;
    R.body = &body;
    const unsigned nw = (block + 31) / 32;
    std::vector<sim_warp_area> warps(nw);
    for (unsigned w = 0; w < nw; ++w) warps[w].nl = (w * 32 + 32 <= block) ? 32 : (block & 31);
    sim_warps = warps.data();
    sim_block_barrier = sim_bar();
    blockIdx.x = b;
    for (unsigned t = 0; t < block; ++t) {
        Fiber& f = R.fibers[t];
        f.tid = t; f.done = false;
        f.stack = R.slab + (size_t)t * R.stack_bytes;
#ifdef SIM_FIBER_ASM
        void** sp = (void**)(((uintptr_t)f.stack + R.stack_bytes) & ~(uintptr_t)15);
        *--sp = nullptr;
        *--sp = (void*)&fiber_main;
        for (int i = 0; i < 6; ++i) *--sp = nullptr;
        f.sp = sp;
#else
        getcontext(&f.uc);
        f.uc.uc_stack.ss_sp = f.stack;
        f.uc.uc_stack.ss_size = R.stack_bytes;
        f.uc.uc_link = &R.sched;
        makecontext(&f.uc, (void (*)())fiber_main, 0);
#endif
    }
    unsigned alive = block;
    unsigned long long last_progress = sim_progress, idle_rounds = 0;
    while (alive) {
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = R.fibers[t];
            if (f.done) continue;
            threadIdx.x = t;
            R.cur = &f;
#ifdef SIM_FIBER_ASM
            sim_fiber_switch(&R.sched_sp, f.sp);
#else
            swapcontext(&R.sched, &f.uc);
#endif
            if (f.done) --alive;
        }
        if (sim_progress == last_progress) {
            if (++idle_rounds > 4) {   // every live fiber is waiting at a barrier nobody else will reach
                fprintf(stderr, "cuda_sim: deadlock in block %u (%u fibers waiting at barriers that cannot complete: divergent collective?)\n", b, alive);
                abort();
            }
        } else { idle_rounds = 0; last_progress = sim_progress; }
    }
    sim_warps = nullptr;
    tl_runner = nullptr;
}

std::mutex g_launch_mu;   // one launch at a time (blockDim / gridDim are process-wide)

}  // namespace

void sim_yield() {
    BlockRunner* r = tl_runner;
    Fiber* f = r->cur;
#ifdef SIM_FIBER_ASM
    sim_fiber_switch(&f->sp, r->sched_sp);
#else
    swapcontext(&f->uc, &r->sched);
#endif
}

void sim_launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    gridDim.x = grid;
    blockDim.x = block;
    const char* e = getenv("RTK_SIM_THREADS");
    unsigned n_workers = e ? (unsigned)atoi(e) : std::thread::hardware_concurrency();
    n_workers = std::max(1u, std::min(std::min(n_workers, grid), 16u));
    const size_t stack_bytes = 256u << 10;
    std::atomic<unsigned> next(0);
    auto worker = [&] {
        BlockRunner R;
        R.stack_bytes = stack_bytes;
        R.slab_bytes = (size_t)block * stack_bytes;
        R.slab = (char*)mmap(nullptr, R.slab_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (R.slab == (char*)MAP_FAILED) { fprintf(stderr, "cuda_sim: cannot map fiber stacks\n"); abort(); }
        R.fibers.resize(block);
        for (;;) {
            const unsigned b = next.fetch_add(1);
            if (b >= grid) break;
            run_block(R, block, b, body);
        }
        munmap(R.slab, R.slab_bytes);
    };
    if (n_workers == 1) { worker(); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < n_workers; ++i) th.emplace_back(worker);
    for (auto& x : th) x.join();
}
