// pipe_microbench.cu — what does ONE warp-wide access cost on the B200 shared-memory / L1 data pipe?
// Evidence for DESIGN.md §4/§8 and profiles/r02_summary.md: cycles per warp instruction (throughput, 11 warps per SM
// like the walk kernel, every SM busy) for the access patterns a tree walk can use for its node reads:
//   lds64 / lds32 / lds16 over R random records (R = nodes of one tree level), a broadcast, a warp shuffle pair
//   (VERDICT r1 item 4b "hold the top levels in registers and walk them with __shfl_sync"), and the 32-byte
//   one-sector-per-lane L2 gather of the bottom records.
// Addresses come from a per-lane LCG so consecutive accesses are independent (no latency chain); the ALU work per
// access (~4 instructions) is far below the issue limit at these rates.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(352, 1) k(uint32_t records, int iters, const uint4* gmem, uint32_t gmask, unsigned long long* sink,
                                            unsigned long long* cycles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    for (uint32_t i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    __syncthreads();
    uint32_t s = threadIdx.x * 9781u + blockIdx.x * 7919u + 1u;
    uint32_t acc = 0;
    const unsigned lane = threadIdx.x & 31;
    const long long t0 = clock64();
    const uint32_t rmask = records - 1;               // records is a power of two
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        if ((i & 3) == 0) lcg(s);
        uint32_t r = (s >> (4 + 5 * (i & 3))) & rmask;
        if (MODE == 9) { acc += r; continue; }        // the address arithmetic alone: the issue floor of this loop
        if (MODE == 0) {            // LDS.64, random record among `records` 8-byte records
            uint32_t a, b;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(sb + 8u * r));
            acc += a ^ b;
        } else if (MODE == 1) {     // LDS.32, random 4-byte record
            uint32_t a;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sb + 4u * r));
            acc += a;
        } else if (MODE == 2) {     // LDS.U16, random 2-byte record
            uint32_t a;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(a) : "r"(sb + 2u * r));
            acc += a;
        } else if (MODE == 3) {     // conflict-free LDS.32: row r, column = lane (the feature read of the walk)
            uint32_t a;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sb + 128u * (r & 255u) + 4u * lane));
            acc += a;
        } else if (MODE == 4) {     // two SHFL.IDX = one 8-byte node held in registers of lane (r % 32)
            acc += __shfl_sync(0xFFFFFFFFu, s, r & 31) ^ __shfl_sync(0xFFFFFFFFu, acc, r & 31);
        } else if (MODE == 5) {     // bottom-record gather: one 32-byte sector per lane from an L2-resident region
            uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
            const uint4* p = gmem + 2ull * ((s * 977u + (r << 7) + i) & gmask);
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "l"(p));
            acc += v0 ^ v7;
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
double run(uint32_t records, int iters, const uint4* g, uint32_t gmask, unsigned long long* sink, unsigned long long* cyc) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<MODE><<<148, 352, 65536>>>(records, iters, g, gmask, sink, cyc);
    k<MODE><<<148, 352, 65536>>>(records, iters, g, gmask, sink, cyc);
    cudaDeviceSynchronize();
    std::vector<unsigned long long> h(148);
    cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double m = 0;
    for (auto v : h) m += (double)v;
    return m / 148.0 / ((double)iters * 11.0);      // cycles per warp instruction per SM (11 warps issue `iters` each)
}

int main() {
    unsigned long long *sink, *cyc;
    uint4* g;
    const size_t gbytes = 32ull << 20;               // 32 MiB: L2-resident like the bottom records of cfg3
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&g, gbytes);
    cudaMemset(g, 1, gbytes);
    const uint32_t gmask = (uint32_t)(gbytes / 32 - 1);
    const int it = 20000;
    printf("{\"what\": \"cycles per warp-wide access per SM (throughput, 11 warps/SM, 148 SMs)\",\n");
    printf(" \"issue_floor_no_memory_op\": %.2f,\n", run<9>(256, it, g, gmask, sink, cyc));
    printf(" \"lds32_conflict_free_feature_read\": %.2f,\n", run<3>(256, it, g, gmask, sink, cyc));
    printf(" \"shfl_pair_64bit_node_in_registers\": %.2f,\n", run<4>(32, it, g, gmask, sink, cyc));
    printf(" \"ldg256_one_sector_per_lane_L2\": %.2f,\n", run<5>(1024, it / 4, g, gmask, sink, cyc));
    const uint32_t recs[] = {1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024};
    const char* names[] = {"lds64_8B_node", "lds32_4B_node", "lds16_2B"};
    for (int m = 0; m < 3; ++m) {
        printf(" \"%s_by_records_per_level\": {", names[m]);
        for (int i = 0; i < 11; ++i) {
            double c = m == 0 ? run<0>(recs[i], it, g, gmask, sink, cyc) : m == 1 ? run<1>(recs[i], it, g, gmask, sink, cyc) : run<2>(recs[i], it, g, gmask, sink, cyc);
            printf("\"%u\": %.2f%s", recs[i], c, i == 10 ? "" : ", ");
        }
        printf("}%s\n", m == 2 ? "" : ",");
    }
    printf("}\n");
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
