// partition_check.cpp — brute-force check of dte_partition.hpp (test infrastructure, CPU only): deal N tuples in
// batches over G positions exactly as PCIeReceiver.sv:298-307 counts them and compare with Deal::locate / prefix.
#include "../distributed-decisiontrees_b200/csrc/dte_partition.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main() {
    unsigned long long seed = 12345;
    auto rnd = [&](unsigned long long m) { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (seed >> 33) % m; };
    for (int it = 0; it < 2000; ++it) {
        const uint64_t bt = 1 + rnd(9), G = 1 + rnd(8), N = rnd(400);
        dte::Deal deal{bt, G};
        // the reference's counters: currDCount counts lines of the current batch, currDevID rotates
        std::vector<std::vector<uint64_t>> held(G);
        uint64_t cur_dev = 0, cur_d = 0;
        for (uint64_t i = 0; i < N; ++i) {
            held[cur_dev].push_back(i);
            if (++cur_d == bt) { cur_d = 0; cur_dev = (cur_dev + 1 == G) ? 0 : cur_dev + 1; }
        }
        for (uint64_t i = 0; i < N; ++i) {
            uint64_t pos, loc;
            deal.locate(i, pos, loc);
            if (pos >= G || loc >= held[pos].size() || held[pos][loc] != i) { printf("locate wrong: bt=%llu G=%llu i=%llu\n", (unsigned long long)bt, (unsigned long long)G, (unsigned long long)i); return 1; }
        }
        // any per-position completed counts: the covered prefix is the largest p such that every tuple < p is completed
        for (int trial = 0; trial < 8; ++trial) {
            std::vector<uint64_t> cnt(G);
            std::vector<char> done(N + 1, 0);
            for (uint64_t r = 0; r < G; ++r) {
                cnt[r] = held[r].empty() ? 0 : rnd(held[r].size() + 1);
                for (uint64_t k = 0; k < cnt[r]; ++k) done[held[r][k]] = 1;
            }
            uint64_t want = 0;
            while (want < N && done[want]) ++want;
            uint64_t got = deal.prefix(cnt.data());
            if (got > N) got = N;              // positions that hold everything they will ever get report past the end
            if (got != want) { printf("prefix wrong: bt=%llu G=%llu N=%llu got %llu want %llu\n", (unsigned long long)bt, (unsigned long long)G, (unsigned long long)N, (unsigned long long)got, (unsigned long long)want); return 1; }
        }
    }
    printf("partition arithmetic ok\n");
    return 0;
}
