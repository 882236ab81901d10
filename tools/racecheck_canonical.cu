// racecheck_canonical.cu — the TEXTBOOK bulk-copy pipeline (one producer thread, N consumer warps, a 2-stage
// shared-memory ring, full/empty mbarriers; the same hand-off CUTLASS's PipelineTmaAsync uses: consumers release a
// stage with a plain mbarrier.arrive, the producer acquires it with try_wait and then issues cp.async.bulk).
// It exists to answer ONE question for profiles/: does `compute-sanitizer --tool racecheck` understand this
// synchronisation?  If it reports hazards HERE, its hazards on dt_walk_tile's ring (same pattern) are a tool limit,
// not a finding.  The kernel checks its own result, so a real race would also show up as a wrong sum.
#include "../distributed-decisiontrees_b200/csrc/dte_kernels.cuh"
#include <cstdio>
#include <vector>
using namespace dte;

constexpr int kStages = 2, kStageBytes = 8192, kConsumers = 4;

__global__ void __launch_bounds__(32 * (kConsumers + 1)) canon(const uint32_t* src, int steps, unsigned long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(sb + 8 * s, 1); mbar_init(sb + 8 * (kStages + s), kConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kConsumers) {
        if (lane == 0) {
            uint32_t slot = 0, par = 1;
            for (int i = 0; i < steps; ++i) {
                mbar_wait(sb + 8 * (kStages + slot), par);                                  // stage free (consumers arrived)
                mbar_arrive_expect_tx(sb + 8 * slot, kStageBytes);
                bulk_g2s(sb + 128 + slot * kStageBytes, src + (size_t)(i % 4) * (kStageBytes / 4), kStageBytes, sb + 8 * slot);
                if (++slot == kStages) { slot = 0; par ^= 1; }
            }
        }
        return;
    }
    unsigned long long acc = 0;
    uint32_t slot = 0, par = 0;
    for (int i = 0; i < steps; ++i) {
        mbar_wait(sb + 8 * slot, par);                                                      // bytes landed
        for (uint32_t k = lane + 32 * warp; k < kStageBytes / 4; k += 32 * kConsumers)
            acc += lds32(sb + 128 + slot * kStageBytes + 4 * k);
        __syncwarp();
        if (lane == 0) mbar_arrive(sb + 8 * (kStages + slot));                              // release the stage
        if (++slot == kStages) { slot = 0; par ^= 1; }
    }
    atomicAdd(out, acc);
}

int main() {
    const int steps = 64, words = 4 * kStageBytes / 4;
    std::vector<uint32_t> h(words);
    unsigned long long want = 0;
    for (int i = 0; i < words; ++i) h[i] = (uint32_t)(i * 2654435761u) >> 8;
    for (int i = 0; i < steps; ++i)
        for (int k = 0; k < kStageBytes / 4; ++k) want += h[(i % 4) * (kStageBytes / 4) + k];
    uint32_t* d; unsigned long long* o;
    cudaMalloc(&d, words * 4); cudaMalloc(&o, 8);
    cudaMemcpy(d, h.data(), words * 4, cudaMemcpyHostToDevice); cudaMemset(o, 0, 8);
    const int blocks = 4;
    canon<<<blocks, 32 * (kConsumers + 1), 128 + kStages * kStageBytes>>>(d, steps, o);
    unsigned long long got = 0;
    cudaError_t st = cudaMemcpy(&got, o, 8, cudaMemcpyDeviceToHost);
    printf("canonical bulk-copy pipeline: %s, sum %llu (want %llu) -> %s\n", cudaGetErrorString(st), got, want * blocks,
           got == want * blocks ? "CORRECT" : "WRONG");
    return got == want * blocks ? 0 : 1;
}
