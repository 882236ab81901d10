// dte_partition.hpp — the index arithmetic of the data-sharded partition, free of CUDA so the CPU test tier can
// check it (tools/partition_check.cpp).  Data lines are dealt to the ring in batches of core_data_batch_cls lines,
// round-robin, starting at the host node (rtl/DTEngine/PCIeReceiver.sv:298-307); with bt tuples per batch and G
// ring positions, batch q goes to position q % G and is that position's (q / G)-th batch.  The engine exposes the
// results of all positions merged back in GLOBAL tuple order.
#pragma once
#include <stdint.h>

namespace dte {

struct Deal {
    uint64_t bt;   // tuples per batch (core_data_batch_cls / tuple_numcls)
    uint64_t G;    // ring positions (numDevs)

    // exposed (global) tuple i -> ring position and that position's local tuple index
    void locate(uint64_t i, uint64_t& pos, uint64_t& loc) const {
        const uint64_t q = i / bt;
        pos = q % G;
        loc = (q / G) * bt + i % bt;
    }
    // first global tuple that position r does NOT hold yet when it holds `cnt` local tuples
    uint64_t first_missing(uint64_t r, uint64_t cnt) const { return (cnt / bt * G + r) * bt + cnt % bt; }
    // number of leading global tuples that are covered when position r holds cnt[r] local tuples
    uint64_t prefix(const uint64_t* cnt) const {
        uint64_t best = ~0ull;
        for (uint64_t r = 0; r < G; ++r) {
            const uint64_t m = first_missing(r, cnt[r]);
            if (m < best) best = m;
        }
        return best;
    }
};

}  // namespace dte
