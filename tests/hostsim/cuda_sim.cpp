// cuda_sim.cpp — see cuda_sim.h (TEST INFRASTRUCTURE).
#include "cuda_sim.h"

thread_local sim_dim3 threadIdx, blockIdx;
sim_dim3 blockDim, gridDim;
pthread_barrier_t sim_block_barrier;
sim_warp_area* sim_warps = nullptr;

void sim_launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    gridDim.x = grid;
    blockDim.x = block;
    const unsigned nw = (block + 31) / 32;
    sim_warps = new sim_warp_area[nw];
    for (unsigned w = 0; w < nw; ++w) {
        const unsigned nl = (w * 32 + 32 <= block) ? 32 : (block & 31);
        pthread_barrier_init(&sim_warps[w].bar, nullptr, nl);
        for (int lg = 1; lg < 5; ++lg)
            for (unsigned f = 0; f < 32; f += (1u << lg)) pthread_barrier_init(&sim_warps[w].gbar[lg][f], nullptr, 1u << lg);
    }
    pthread_barrier_init(&sim_block_barrier, nullptr, block);
    // persistent worker threads: one per CUDA thread, looping over the blocks
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t) {
        th.emplace_back([&, t] {
            for (unsigned b = 0; b < grid; ++b) {
                threadIdx.x = t;
                blockIdx.x = b;
                body();
                pthread_barrier_wait(&sim_block_barrier);  // block boundary
            }
        });
    }
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&sim_block_barrier);
    for (unsigned w = 0; w < nw; ++w) pthread_barrier_destroy(&sim_warps[w].bar);
    delete[] sim_warps;
    sim_warps = nullptr;
}
