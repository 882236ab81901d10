// dte_host.cpp — C++ host program over the C ABI (include/dte.h), with the reference's own
// parameter surface: the "User inputs" block of profiler/profiler.cpp:31-41
//     N_trees, Depth_tree, Size_tuple_Bytes
// Here Depth_tree counts COMPARISON levels (leaves one level below; DESIGN.md "depth convention").
// From those three numbers it derives the CSR values, builds a synthetic ensemble + tuple set with
// the same counter-based generator as the Python host and the device generator, then drives the
// engine exactly as a Catapult host would: register writes, `start`, one 128-bit line stream
// (weights | feature indexes | tuples), result lines back.  Prints tuples/s and a checksum the
// tests compare with the Python path.
//
//   dte_host <N_trees> <Depth_tree> <Size_tuple_Bytes> <N_tuples> [clusters=8] [mode=stream|host] [gpus=0] [partition=data|ensemble]
// gpus = comma list of CUDA ordinals, e.g. 0,1,2,3,4,5,6,7: ONE handle (dte_create_multi) then drives the whole ring:
//   data     = BASELINE configs[4]: ensemble replicated (broadcast_trees), tuple lines dealt in batches over the GPUs
//   ensemble = BASELINE configs[3]: N_trees cut into contiguous chunks (one per GPU), every tuple on every GPU,
//              partial scores combined in ring order (aggreg_enabled)
#include "../include/dte.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static uint64_t splitmix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u01(uint64_t z) { return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f); }

#define CHECK(call)                                                                   \
    do {                                                                              \
        int _rc = (call);                                                             \
        if (_rc) { fprintf(stderr, "%s -> %d: %s\n", #call, _rc, dte_last_error(e)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s N_trees Depth_tree Size_tuple_Bytes N_tuples [clusters] [stream|host]\n", argv[0]);
        return 2;
    }
    const uint32_t N_trees = (uint32_t)atoi(argv[1]), Depth_tree = (uint32_t)atoi(argv[2]);
    const uint32_t Size_tuple_Bytes = (uint32_t)atoi(argv[3]);
    const uint64_t N_tuples = strtoull(argv[4], nullptr, 10);
    const uint32_t clusters = argc > 5 ? (uint32_t)atoi(argv[5]) : 8;
    const bool stream_mode = !(argc > 6 && !strcmp(argv[6], "host"));
    std::vector<int> gpus;
    if (argc > 7) {
        for (const char* q = argv[7]; *q;) { gpus.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
    }
    if (gpus.empty()) gpus.push_back(0);
    const bool ensemble_part = argc > 8 && !strcmp(argv[8], "ensemble");
    const uint32_t G = (uint32_t)gpus.size();
    const uint32_t missing = 0xBF800000u, F = Size_tuple_Bytes / 4, D = Depth_tree;
    const uint64_t seed_e = 0xD7EE5, seed_t = 0x7091E5;

    uint64_t regs[8];
    dte_t* e = nullptr;
    if (dte_csr_from_profile(N_trees, D, Size_tuple_Bytes, clusters, missing, N_tuples, regs)) {
        fprintf(stderr, "parameters outside the engine's CSR field ranges\n");
        return 2;
    }
    if (dte_create_multi(&e, gpus.data(), (int)G)) { fprintf(stderr, "dte_create_multi failed (no CUDA device; there is no CPU fallback)\n"); return 3; }
    if (G > 1) {
        // the multi-device flags a Catapult host writes (EngineCSR.sv:194-216): host_node | multiple_nodes | pcie_receiver_enabled
        // + broadcast_trees (data dealt in batches of 1024 tuples)  or  broadcast_data | aggreg_enabled (tree chunks)
        const uint64_t t_cls = Size_tuple_Bytes / 16;
        uint64_t flags = 0x2 | 0x20 | 0x40;
        flags |= ensemble_part ? (0x4 | 0x10) : 0x8;
        regs[0] = flags | ((1024 * t_cls) << 32);
        regs[2] = (uint64_t)G << 32;                     // numDevs; chunk fields 0 = cut the trees evenly
        if (ensemble_part) {                             // S sized for one device's chunk
            const uint64_t per = (N_trees + G - 1) / G, S = (per + 8 * clusters - 1) / (8 * clusters);
            regs[4] = (regs[4] & ~(0xFFull << 36)) | (S << 36);
        }
    }
    for (int i = 0; i < 8; ++i) CHECK(dte_softreg_write(e, 201 + i, regs[i]));

    // ---- synthetic ensemble in the reference's stream layout (same law as layout.synth_ensemble) ----
    const uint32_t w_cls = (uint32_t)((regs[3] >> 16) & 0xFFFF), f_cls = (uint32_t)((regs[3] >> 32) & 0xFFFF);
    const uint32_t n_int = (1u << D) - 1, n_all = (2u << D) - 1;
    std::vector<uint32_t> wl((size_t)N_trees * w_cls * 4, 0);
    std::vector<uint16_t> fl((size_t)N_trees * f_cls * 8, 0);
    for (uint32_t t = 0; t < N_trees; ++t) {
        for (uint32_t i = 0; i < n_all; ++i) {
            float u = u01(splitmix64(seed_e, (uint64_t)t * n_all + i));
            float v = i < n_int ? u : (u * 2.0f - 1.0f + 0.02f) / (float)N_trees;
            wl[(size_t)t * w_cls * 4 + i] = fbits(v);
        }
        for (uint32_t i = 0; i < n_int; ++i) {
            uint64_t z = splitmix64(seed_e ^ 0x5EED, (uint64_t)t * n_int + i);
            fl[(size_t)t * f_cls * 8 + i] = (uint16_t)((z % F) | (((z >> 33) & 1) << 13));
        }
    }
    // DMA buffers from the "driver" (page-locked, every device of the handle can reach them at full PCIe speed)
    uint32_t* x = nullptr;
    float* scores = nullptr;
    CHECK(dte_host_alloc(e, (size_t)N_tuples * F * 4, reinterpret_cast<void**>(&x)));
    CHECK(dte_host_alloc(e, ((size_t)N_tuples + 4) * 4, reinterpret_cast<void**>(&scores)));
    for (uint64_t i = 0; i < (uint64_t)N_tuples * F; ++i) {
        uint64_t z = splitmix64(seed_t, i);
        x[i] = ((uint32_t)(z & 0xFFFFFFFFull) % 1000000u < 10000u) ? missing : fbits(u01(z));
    }
    memset(scores, 0, ((size_t)N_tuples + 4) * 4);
    auto t0 = std::chrono::steady_clock::now();
    if (stream_mode) {
        CHECK(dte_softreg_write(e, 200, 1));                                       // start
        CHECK(dte_stream_write(e, wl.data(), wl.size() / 4));                      // all weight lines
        CHECK(dte_stream_write(e, fl.data(), fl.size() / 8));                      // all feature-index lines
        t0 = std::chrono::steady_clock::now();
        CHECK(dte_stream_write(e, x, (size_t)N_tuples * F / 4));                   // tuple lines
        size_t got = 0, total = 0;
        do {
            CHECK(dte_stream_read(e, scores + total * 4, N_tuples / 4 - total, &got));
            total += got;
        } while (got && total < N_tuples / 4);
        int done = 0;
        CHECK(dte_process_done(e, &done));
        if (!done) fprintf(stderr, "warning: process_done not reached (%zu of %llu lines)\n", total, (unsigned long long)(N_tuples / 4));
    } else {
        CHECK(dte_load_ensemble(e, wl.data(), wl.size() / 4, fl.data(), fl.size() / 8, 0, 0));
        t0 = std::chrono::steady_clock::now();
        CHECK(dte_infer_host(e, x, N_tuples, scores, nullptr));
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t sum = 0;
    const uint64_t n_out = stream_mode ? (N_tuples / 4) * 4 : N_tuples;
    for (uint64_t i = 0; i < n_out; ++i) sum += fbits(scores[i]);
    uint64_t exec_ns = 0;
    dte_softreg_read(e, 223, &exec_ns);
    printf("{\"N_trees\": %u, \"Depth_tree\": %u, \"Size_tuple_Bytes\": %u, \"N_tuples\": %llu, \"mode\": \"%s\", "
           "\"gpus\": %u, \"partition\": \"%s\", "
           "\"tuples_per_s\": %.1f, \"exec_ns\": %llu, \"score_words_sum\": %llu, \"results\": %llu}\n",
           N_trees, D, Size_tuple_Bytes, (unsigned long long)N_tuples, stream_mode ? "stream" : "host", G,
           G == 1 ? "single" : (ensemble_part ? "ensemble" : "data"), N_tuples / dt,
           (unsigned long long)exec_ns, (unsigned long long)sum, (unsigned long long)n_out);
    dte_host_free(e, x);
    dte_host_free(e, scores);
    dte_destroy(e);
    return 0;
}
