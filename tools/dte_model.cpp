// dte_model.cpp — closed-form throughput / device-count model, the B200 counterpart of the
// reference's profiler (profiler/profiler.cpp:51-118, profiler_performance_model.cpp:50-110).
// Same inputs (N_trees, Depth_tree, Size_tuple_Bytes), same structure (engine rate, capacity-bound
// minimum device count, host-link and network caps), B200 constants measured in round 1
// (profiles/r01_summary.md).  Prints the reference's Catapult law next to it for the same inputs.
//
//   dte_model <N_trees> <Depth_tree> <Size_tuple_Bytes> [n_gpus=1]
//   Depth_tree = comparison levels D (the profiler's "depth" counts the leaf level too: its depth = D+1).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s N_trees Depth_tree Size_tuple_Bytes [n_gpus]\n", argv[0]); return 2; }
    const double T = atof(argv[1]), D = atof(argv[2]), tuple_bytes = atof(argv[3]);
    const int G = argc > 4 ? atoi(argv[4]) : 1;

    // ---- B200 constants (measured, round 1) ----
    const double visits_per_s = 102.8e6 * 1024 * 12;     // node visits/s per GPU at cfg3 (dt_walk_tile<4,2,1,0,384>, profiles/r02_summary.md)
    const double pcie_Bps = 54.5e9;                       // host->device through dte_infer_host, pinned
    const double hbm_Bps = 6585.4e9;                      // MEASURED_PEAKS.json
    const double nvlink_Bps = 725e9;                      // 8-rank all-reduce bus bandwidth (B200_PROFILING.md)
    const double hbm_bytes = 180e9;
    const double smem_bytes = 227.0 * 1024;

    const double tree_bytes = 10.0 * std::pow(2.0, D);    // repacked tree, same as the reference's 10*2^D
    const double ens_bytes = T * tree_bytes;
    const double walk = visits_per_s / (T * D);           // tuples/s per GPU, walk-bound
    const double hbm = hbm_Bps / (tuple_bytes + 4);       // compulsory DRAM traffic: tuple in, score out
    const double pcie = pcie_Bps / tuple_bytes;           // when tuples come from the host
    const double per_gpu_resident = std::min(walk, hbm);
    const double per_gpu_host = std::min(per_gpu_resident, pcie);
    // feature-major tile: 32 tuples x Size_tuple_Bytes per warp must fit next to a 64 KiB ring
    const int warps = (int)std::min(8.0, std::floor((smem_bytes - 65536 - 128) / (32 * tuple_bytes)));
    const int min_gpus = (int)std::ceil(ens_bytes / (hbm_bytes * 0.5));

    printf("B200 model  (N_trees=%.0f, D=%.0f comparison levels, %.0f B tuples, %d GPU%s)\n", T, D, tuple_bytes, G, G > 1 ? "s" : "");
    printf("  ensemble            : %.1f MiB repacked (%s)\n", ens_bytes / 1048576.0, ens_bytes < 100e6 ? "L2-resident" : "HBM-resident");
    printf("  min GPUs (capacity) : %d\n", std::max(1, min_gpus));
    printf("  tile warps per SM   : %d%s\n", std::max(0, warps), warps < 1 ? "  (falls back to the generic kernel)" : "");
    printf("  walk-bound          : %.3e tuples/s per GPU  (%.3e visits/s)\n", walk, visits_per_s);
    printf("  HBM-bound           : %.3e tuples/s per GPU\n", hbm);
    printf("  PCIe-bound          : %.3e tuples/s per GPU (tuples streamed from the host)\n", pcie);
    printf("  data-sharded, %d GPU : %.3e tuples/s resident, %.3e from host\n", G, G * per_gpu_resident, G * per_gpu_host);
    {   // ensemble-sharded: T/G trees per GPU, every GPU walks every tuple, one reduce of 4 B/tuple/GPU
        const double w = visits_per_s / ((T / G) * D);
        const double red = nvlink_Bps / 4.0;
        printf("  ensemble-sharded    : %.3e tuples/s (walk %.3e, reduce cap %.3e)\n", std::min({w, hbm, red}), w, red);
    }
    // ---- the reference's own law for the same inputs (profiler/profiler.cpp:97-102), Catapult v1.2 ----
    const double f = 150e6, ncu = 8, npe = 8;             // RTL instantiates 8 x 8 PEs (DTEngine_Types.sv:25-27)
    const double fpga = f * ncu * npe / (D * T);
    const double fpga_nodes = ncu * npe * 8192.0;
    const int fpga_min = (int)std::ceil(T * std::pow(2.0, D + 1) / fpga_nodes);
    printf("Catapult model (reference law, 150 MHz, 64 PEs, modelled not measured)\n");
    printf("  engine              : %.3e tuples/s per FPGA, min FPGAs (capacity) %d, PCIe cap %.3e\n", fpga, std::max(1, fpga_min), 2.2e9 / tuple_bytes);
    return 0;
}
