// dte_device.cuh — one GPU of an engine ("one FPGA of the ring"): the resident ensemble in its repacked
// layout, the launch planner, the walk launch, and the landing-slot pipeline (H2D | walk | D2H on three
// streams) that both the line-stream path and the host fast path feed.  Host code only; the kernels are in
// dte_kernels.cuh.  Reference behaviour cited per function (paths relative to the reference root).
#pragma once
#include "dte_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

namespace dte {

constexpr int kNumSlots = 3;          // landing slots per device: H2D of i+1, walk of i, D2H of i-1 overlap

enum { KERNEL_AUTO = 0, KERNEL_GENERIC = 1, KERNEL_TILE = 2, KERNEL_TILE_STAGED = 3 };

// experiment knobs, env DTE_TUNE="ilp=4,pair=2,stages=1,warps=10,phased=1,fill=0,chunk=65536"
// (ilp: trees per warp, pair: warps per tuple group, stages: ring depth, warps: consumer-warp cap,
//  phased: 0 off / k>=1 on with part A ending k levels earlier, fill=1: one bulk copy per tree,
//  chunk: tuples per landing slot)
struct Tune {
    int ilp = 0, stages = 0, warps = 0, pair = 0, fill = 0, phased = -1;
    size_t chunk = 0;
};

inline void parse_tune(Tune& t) {
    const char* s = getenv("DTE_TUNE");
    if (!s) return;
    std::string str(s);
    size_t pos = 0;
    while (pos < str.size()) {
        size_t comma = str.find(',', pos);
        std::string kv = str.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        size_t eq = kv.find('=');
        if (eq != std::string::npos) {
            std::string k = kv.substr(0, eq);
            long long v = atoll(kv.c_str() + eq + 1);
            if (k == "ilp") t.ilp = (int)v;
            else if (k == "stages") t.stages = (int)v;
            else if (k == "warps") t.warps = (int)v;
            else if (k == "pair") t.pair = (int)v;
            else if (k == "fill") t.fill = (int)v;
            else if (k == "phased") t.phased = (int)v;
            else if (k == "chunk") t.chunk = (size_t)v;
        }
        if (comma == std::string::npos) break;
        pos = comma + 1;
    }
}

// The geometry registers, decoded (EngineCSR.sv:218-233).
struct Geom {
    uint32_t D = 0, K = 0, S = 0, missing = 0, w_cls = 0, f_cls = 0, tuple_cls = 0;
    uint32_t F() const { return tuple_cls * 4; }
    size_t tuple_bytes() const { return (size_t)tuple_cls * 16; }
};

struct Plan {               // how the next walk will be launched
    int variant = KERNEL_GENERIC;
    int ilp = 8, pair = 1, nstages = 0, nwarps = 0;   // nwarps = consumer warps = groups * pair
    bool wide = false;
    size_t smem = 0;
    int threads() const { return variant == KERNEL_GENERIC ? 128 : 32 * (nwarps + (variant == KERNEL_TILE_STAGED ? 1 : 0)); }
    int thread_bound() const {                         // NT of the instantiation that will run (see dt_walk_tile)
        const int nt_max = pair == 4 ? 672 : (ilp == 8 ? 288 : 416);
        return (nt_max > 384 && threads() <= 384) ? 384 : nt_max;
    }
};

// The ensemble in the device layout, still in host memory (built once, uploaded to one or many devices).
//   top[t][2^(D-2)]    8-byte heap records {thr bits, fidx | (8 + 8*missing_right) << 16} for levels 0..D-3
//   bottom[t][2^(D-2)] one record per level-(D-2) node: that node, its two children and their four leaves;
//                      32 bytes (= one L2 sector) when every feature index < 512, 64 bytes otherwise
// Source layout: per tree W = heap array of 2^(D+1)-1 fp32 words padded to w_cls lines, FI = 2^D-1
// u16 padded to f_cls lines (DTPU.sv:282-338,579-596).
struct PackedEnsemble {
    Geom g;
    uint32_t T = 0, Tpad = 0, Dtop = 0, top_stride = 1, nb = 1;
    bool wide = false;
    std::vector<uint2> top;
    std::vector<uint4> bottom;
};

// returns 0 or a negative dte_status; *msg receives the reason
inline int pack_ensemble(const Geom& g, const unsigned char* wl, size_t n_wl, const unsigned char* fl, size_t n_fl,
                         uint32_t first, uint32_t count, PackedEnsemble& out, std::string& msg) {
    char buf[256];
    if (n_wl % g.w_cls) {
        snprintf(buf, sizeof buf, "weights stream: %zu lines is not a multiple of %u lines per tree", n_wl, g.w_cls);
        msg = buf; return -1;
    }
    const size_t T_all = n_wl / g.w_cls;
    if (n_fl != T_all * g.f_cls) {
        snprintf(buf, sizeof buf, "feature-index stream: %zu lines, expected %zu (= %zu trees x %u)", n_fl, T_all * g.f_cls, T_all, g.f_cls);
        msg = buf; return -1;
    }
    if (count == 0 && first == 0) count = (uint32_t)T_all;
    if ((size_t)first + count > T_all || count == 0) {
        snprintf(buf, sizeof buf, "tree chunk [%u, %u) outside the %zu trees of the stream", first, first + count, T_all);
        msg = buf; return -1;
    }
    const uint32_t D = g.D, F = g.F();
    const uint32_t Dk = std::max(D, 2u);
    const uint32_t Dtop = Dk - 2;
    const uint32_t nb = 1u << Dtop;
    const uint32_t top_stride = std::max(1u, 1u << Dtop);
    const uint32_t Tpad = (count + 7u) & ~7u;
    const size_t wstride = (size_t)g.w_cls * 4, fstride = (size_t)g.f_cls * 8;
    const uint32_t* Wall = reinterpret_cast<const uint32_t*>(wl);
    const uint16_t* Fall = reinterpret_cast<const uint16_t*>(fl);

    // pass 1: contract check + widest feature index.  Early leaves: bit 14 of a node's index word says "my children are
    // leaves" (DTPU.sv:596,661,712) — the walk then ends at the child cell W[2n+1+right], whatever D says.  (The RTL
    // freezes node_offset WITHOUT the direction bit while its memory address keeps advancing, so with >= 1 level left
    // it reads a wrong cell; the build-defined rule here is the evident intent, SURVEY R3 / DESIGN.md.)  Nodes below
    // an early leaf are unreachable: their words are don't-care and are not validated.
    const uint32_t n_int = (1u << D) - 1;
    uint32_t max_f = 0;
    bool any_early = false;
    std::vector<uint8_t> dead(n_int);
    for (uint32_t t = 0; t < count; ++t) {
        const uint16_t* fi = Fall + (size_t)(first + t) * fstride;
        std::fill(dead.begin(), dead.end(), 0);
        for (uint32_t i = 0; i < n_int; ++i) {
            const bool early = !dead[i] && (fi[i] & 0x4000u);
            if (dead[i] || early) {
                if (2 * i + 2 < n_int) dead[2 * i + 1] = dead[2 * i + 2] = 1;
                any_early |= early && (2 * i + 2 < n_int);
                if (dead[i]) continue;
            }
            const uint32_t f = fi[i] & 0x7FFu;
            if (f >= F) {
                snprintf(buf, sizeof buf, "tree %u node %u: feature index %u >= %u features", first + t, i, f, F);
                msg = buf; return -4;
            }
            max_f = std::max(max_f, f);
        }
    }
    const bool wide = max_f >= 512;
    const uint32_t BV = wide ? 4 : 2;

    out.g = g;
    out.T = count; out.Tpad = Tpad; out.Dtop = Dtop; out.top_stride = top_stride; out.nb = nb; out.wide = wide;
    out.top.assign((size_t)Tpad * top_stride, make_uint2(0u, 8u << 16));
    out.bottom.assign((size_t)Tpad * nb * BV, make_uint4(0, 0, 0, 0));
    std::vector<uint32_t> Wk((2u << Dk) - 1);
    std::vector<uint16_t> Fk((1u << Dk) - 1);
    for (uint32_t t = 0; t < count; ++t) {
        const uint32_t* W = Wall + (size_t)(first + t) * wstride;
        const uint16_t* FI = Fall + (size_t)(first + t) * fstride;
        if (D == 1) {
            // one comparison level: extend to two levels by giving both children of a dummy level
            // the same leaf — the same function of x (see DESIGN.md "D = 1")
            Wk[0] = W[0]; Wk[1] = 0; Wk[2] = 0;
            Wk[3] = W[1]; Wk[4] = W[1]; Wk[5] = W[2]; Wk[6] = W[2];
            Fk[0] = FI[0]; Fk[1] = 0; Fk[2] = 0;
        } else {
            std::copy(W, W + ((2u << D) - 1), Wk.begin());
            std::copy(FI, FI + ((1u << D) - 1), Fk.begin());
            if (any_early) {
                // expand early leaves into complete subtrees whose every leaf is the early leaf's value: the same
                // function of x on a complete tree, so the kernels need no early-exit path
                std::vector<uint8_t>& konst = dead;
                std::fill(konst.begin(), konst.end(), 0);
                for (uint32_t i = 0; i < n_int; ++i) {
                    const uint32_t l = 2 * i + 1, r = 2 * i + 2;
                    if (konst[i]) {                       // inside a constant subtree: value travels in Wk[i]
                        const uint32_t v = Wk[i];
                        Wk[l] = v; Wk[r] = v;
                        if (r < n_int) konst[l] = konst[r] = 1;
                        Wk[i] = 0; Fk[i] = 0;             // any threshold / feature 0: both ways lead to v
                    } else if ((Fk[i] & 0x4000u) && r < n_int) {
                        konst[l] = konst[r] = 1;          // children are leaves: W[l], W[r] already hold their values
                    }
                    Fk[i] &= (uint16_t)~0x4000u;
                }
            }
        }
        uint2* tp = out.top.data() + (size_t)t * top_stride;
        for (uint32_t n = 0; n + 1 < (1u << Dtop); ++n) {
            const uint32_t f = Fk[n] & 0x7FFu, mr = (Fk[n] >> 13) & 1u;
            tp[n] = make_uint2(Wk[n], f | ((8u + 8u * mr) << 16));
        }
        uint4* bp = out.bottom.data() + (size_t)t * nb * BV;
        for (uint32_t j = 0; j < nb; ++j) {
            const uint32_t n = nb - 1 + j, l = 2 * n + 1, r = 2 * n + 2;
            const uint32_t fp = Fk[n] & 0x7FFu, mp = (Fk[n] >> 13) & 1u;
            const uint32_t fl_ = Fk[l] & 0x7FFu, ml = (Fk[l] >> 13) & 1u;
            const uint32_t fr = Fk[r] & 0x7FFu, mr = (Fk[r] >> 13) & 1u;
            const uint4 leaves = make_uint4(Wk[2 * l + 1], Wk[2 * l + 2], Wk[2 * r + 1], Wk[2 * r + 2]);
            if (wide) {
                bp[j * 4 + 0] = make_uint4(Wk[n], Wk[l], Wk[r], fp | (mp << 16));
                bp[j * 4 + 1] = make_uint4(fl_ | (ml << 16), fr | (mr << 16), 0, 0);
                bp[j * 4 + 2] = leaves;
            } else {
                const uint32_t pack = (fp | (mp << 9)) | ((fl_ | (ml << 9)) << 10) | ((fr | (mr << 9)) << 20);
                bp[j * 2 + 0] = make_uint4(Wk[n], Wk[l], Wk[r], pack);
                bp[j * 2 + 1] = leaves;
            }
        }
    }
    return 0;
}

// One landing slot: a device buffer the tuple stream lands in, plus the scores (labels) of its tuples.
struct Slot {
    unsigned char* d_tup = nullptr;
    float* d_sc = nullptr;        // final scores of this slot (single / data-sharded / host node of a group)
    float* d_part = nullptr;      // group (ensemble-sharded) mode: this device's partial scores
    uint8_t* d_lb = nullptr;
    size_t fill = 0;              // bytes of tuple stream landed (or enqueued to land) in d_tup
    size_t walked = 0;            // whole tuples already submitted to the walk kernel
    cudaEvent_t ev_walk = nullptr;   // last walk reading d_tup / writing d_sc, d_part
    cudaEvent_t ev_d2h = nullptr;    // last D2H reading d_sc / d_lb
    cudaEvent_t ev_land = nullptr;   // last copy into d_tup issued on THIS device's copy stream
    cudaEvent_t ev_comb = nullptr;   // group mode, host device: last combine reading every device's d_part
};

struct Pending {              // one enqueued D2H of results: complete when `ev` fires
    cudaEvent_t ev;
    uint64_t upto;            // device-local count of result tuples complete once ev has fired
};

struct Dev {
    int ordinal = 0, sm_count = 0, smem_optin = 0;
    cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    bool timing_pending = false;

    // ---- resident ensemble ----
    Geom g;                           // geometry the resident ensemble was loaded with
    uint32_t T = 0, Tpad = 0, Dtop = 0, top_stride = 1, nb = 1;
    bool wide = false;
    uint2* d_top = nullptr;
    uint4* d_bottom = nullptr;
    uint64_t ensemble_bytes = 0;

    // ---- landing slots ----
    Slot slot[kNumSlots];
    size_t cap_tuples = 0;            // tuples per slot
    uint32_t slot_F = 0;
    bool slot_parts = false;          // d_part allocated
    int cur = 0;                      // slot being filled
    // result sink of the next submitted tuples: direct host pointers (fast path) or the pinned ring (line stream)
    float* sink_sc = nullptr;
    uint8_t* sink_lb = nullptr;
    float* h_ring = nullptr;          // pinned, device-local tuple order
    size_t ring_cap = 0;              // tuples
    uint64_t res_enq = 0;             // result tuples whose D2H has been enqueued (device-local count)
    uint64_t res_done = 0;            // ... whose D2H has completed
    std::deque<Pending> pend;
    std::vector<cudaEvent_t> ev_pool;

    // pageable sources: a small pinned staging ring filled by a few copy threads, DMA'd from there (the driver's own
    // pageable path stages through one thread and reaches ~10 GB/s; PCIe Gen5 takes 55)
    unsigned char* h_stage[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};
    int stage_next = 0;

    uint64_t kernel_launches = 0;
    double last_walk_ms = 0;
    uint64_t tuples_landed = 0;       // whole tuples landed on this device since `start`
};

// ---- launch planning ---------------------------------------------------------------------------
inline size_t tile_smem(uint32_t F, int groups, int trees_per_stage, int nstages, uint32_t top_stride) {
    return (size_t)kHdrBytes + (size_t)nstages * trees_per_stage * top_stride * 8 + (size_t)F * 32 * groups * 4;
}

inline Plan make_plan(const Dev& d, const Tune& tune, int want) {
    Plan p;
    p.wide = d.wide;
    const uint32_t F = d.g.F();
    const size_t budget = (size_t)d.smem_optin;
    // tuple groups (32 tuples, F*128 B of shared memory each) that fit next to the ring
    auto max_groups = [&](int ilp, int pair, int nstages) -> int {
        const size_t fixed = tile_smem(F, 0, ilp * pair, nstages, d.top_stride);
        if (fixed >= budget) return 0;
        const int warp_cap = (pair == 4) ? 20 : (ilp == 8 ? 8 : 12);   // thread bound of dt_walk_tile (+1 producer warp)
        int g = (int)std::min<size_t>((size_t)(warp_cap / pair), (budget - fixed) / ((size_t)F * 128));
        if (tune.warps) g = std::min(g, std::max(1, tune.warps / pair));
        return g;
    };
    const bool can_stage = d.Dtop >= 3;          // below that there is nothing worth staging
    // staged candidates {trees per warp, warps per tuple group, ring stages}.  Measured on B200
    // (profiles/r01_summary.md); the first candidate that fits with the most walks in flight wins.
    const int cand[6][3] = {{4, 2, 1}, {2, 4, 1}, {8, 1, 1}, {8, 1, 2}, {4, 1, 2}, {4, 1, 1}};
    Plan staged;
    if (can_stage) {
        int best = 0;
        for (auto& c : cand) {
            int ilp = tune.ilp ? tune.ilp : c[0];
            int pair = tune.pair ? tune.pair : c[1];
            int st = tune.stages ? tune.stages : c[2];
            if (pair == 4) ilp = 2; else if (pair == 2) ilp = (ilp == 2) ? 2 : 4; else { pair = 1; if (ilp != 4 && ilp != 8) ilp = 8; }
            st = std::max(1, std::min(st, 4));
            const int g = max_groups(ilp, pair, st);
            const int score = g * pair * ilp * 8 + (pair == 2 ? 4 : 0) - (pair == 4 ? 4 : 0) + st;
            if (g >= 1 && score > best) {
                best = score;
                staged.variant = KERNEL_TILE_STAGED;
                staged.ilp = ilp; staged.pair = pair; staged.nstages = st; staged.nwarps = g * pair; staged.wide = d.wide;
                staged.smem = tile_smem(F, g, ilp * pair, st, d.top_stride);
            }
            if (tune.ilp && tune.stages && tune.pair) break;
        }
    }
    Plan tile;
    {
        const int ilp = (tune.ilp == 4) ? 4 : 8;
        const int g = max_groups(ilp, 1, 0);
        if (g >= 1) {
            tile.variant = KERNEL_TILE; tile.ilp = ilp; tile.pair = 1; tile.nstages = 0; tile.nwarps = g; tile.wide = d.wide;
            tile.smem = tile_smem(F, g, ilp, 0, d.top_stride);
        }
    }
    if (want == KERNEL_TILE_STAGED && staged.nwarps >= 1) return staged;
    if (want == KERNEL_TILE && tile.nwarps >= 1) return tile;
    if (want == KERNEL_GENERIC) return p;
    // AUTO (or a forced variant that does not fit): staged > tile > generic
    if (staged.nwarps >= 2) return staged;
    if (tile.nwarps >= 1) return tile;
    return p;
}

// phased ring refill needs >= 3 staged levels and at most 4 ring stages (16 mbarriers in the header).
// Measured (profiles/r01_summary.md): +7 % at D = 12 (64 KiB stage), -4 % at D <= 10 (<= 16 KiB stage, the
// refill is already cheap there and the extra barrier hand-offs cost more than they hide).
inline uint32_t phased_level(const Dev& d, const Tune& tune, const Plan& pl) {
    if (pl.variant != KERNEL_TILE_STAGED) return 0xFFFFFFFFu;
    const bool phased = tune.phased >= 1 || (tune.phased == -1 && d.Dtop >= 10);
    // part A = levels 0..Lw; DTE_TUNE phased=k moves the split k-1 levels further up (smaller part A)
    const uint32_t up = tune.phased > 1 ? (uint32_t)tune.phased - 1 : 0;
    return (phased && d.Dtop >= 3 + up && pl.nstages <= 4) ? d.Dtop - 3 - up : 0xFFFFFFFFu;
}

template <int ILP, int P, bool STAGED, bool WIDE, int NT>
cudaError_t launch_tile_nt(const WalkParams& wp, int grid, int threads, size_t smem, cudaStream_t st) {
    auto k = dt_walk_tile<ILP, P, STAGED, WIDE, NT>;
    cudaError_t rc = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    k<<<grid, threads, smem, st>>>(wp);
    return cudaGetLastError();
}
// thread-bound classes (see dt_walk_tile): <= 12 warps -> the 168-register instantiation, else the wide one
template <int ILP, int P, bool STAGED, bool WIDE>
cudaError_t launch_tile(const WalkParams& wp, int grid, int threads, size_t smem, cudaStream_t st) {
    constexpr int NT_MAX = P == 4 ? 672 : (ILP == 8 ? 288 : 416);
    if constexpr (NT_MAX > 384) {
        if (threads <= 384) return launch_tile_nt<ILP, P, STAGED, WIDE, 384>(wp, grid, threads, smem, st);
    }
    return launch_tile_nt<ILP, P, STAGED, WIDE, NT_MAX>(wp, grid, threads, smem, st);
}

// Launch the walk of n tuples on device d (cudaSetDevice already done).  Returns a cudaError_t.
inline cudaError_t launch_walk(Dev& d, const Tune& tune, int want_variant, const void* d_tuples, size_t n, float* d_scores,
                               uint8_t* d_labels, cudaStream_t st, bool accumulate, const char** why) {
    *why = nullptr;
    if (n == 0) return cudaSuccess;
    const Plan pl = make_plan(d, tune, want_variant);
    WalkParams wp;
    wp.top = d.d_top;
    wp.bottom = d.d_bottom;
    wp.tuples = static_cast<const float*>(d_tuples);
    wp.scores = d_scores;
    wp.labels = d_labels;
    wp.n = n;
    wp.F = d.g.F();
    wp.Dtop = d.Dtop;
    wp.top_stride = d.top_stride;
    wp.nb = d.nb;
    // slots beyond S are never issued (DTPU.sv:519-531); groups past the last tree add exact zeros
    wp.groups = (uint32_t)std::min<uint64_t>((uint64_t)d.g.S * d.g.K, d.Tpad / 8);
    wp.K = d.g.K;
    wp.missing = d.g.missing;
    wp.nwarps = (uint32_t)pl.nwarps;
    wp.nstages = (uint32_t)pl.nstages;
    wp.accumulate = accumulate ? 1u : 0u;
    wp.wide_rows = (wp.F % 8 == 0 && (reinterpret_cast<uintptr_t>(d_tuples) & 31u) == 0) ? 1u : 0u;
    wp.fill_split = tune.fill ? 1u : 0u;      // DTE_TUNE fill=1: one bulk copy per tree instead of one per stage
    wp.Lw = phased_level(d, tune, pl);
    wp.tiles = 0;
    cudaError_t rc;
    if (pl.variant == KERNEL_GENERIC) {
        const int threads = 128;
        const unsigned long long blocks = (n + threads - 1) / threads;
        if (blocks > 0x7FFFFFFFull) { *why = "batch too large for one launch"; return cudaErrorInvalidValue; }
        if (pl.wide) dt_walk_generic<true><<<(unsigned)blocks, threads, 0, st>>>(wp);
        else dt_walk_generic<false><<<(unsigned)blocks, threads, 0, st>>>(wp);
        rc = cudaGetLastError();
    } else {
        const size_t M = 32ull * (pl.nwarps / pl.pair);
        const unsigned long long tiles = (n + M - 1) / M;
        if (tiles > 0xFFFFFFFFull) { *why = "batch too large for one launch"; return cudaErrorInvalidValue; }
        wp.tiles = (uint32_t)tiles;
        const int grid = (int)std::min<unsigned long long>(tiles, (unsigned long long)d.sm_count);
        const bool staged = pl.variant == KERNEL_TILE_STAGED;
        const int threads = pl.threads();
        const size_t sm = pl.smem;
#define DTE_LAUNCH(ILP_, P_, ST_) (pl.wide ? launch_tile<ILP_, P_, ST_, true>(wp, grid, threads, sm, st) \
                                           : launch_tile<ILP_, P_, ST_, false>(wp, grid, threads, sm, st))
        if (staged) {
            if (pl.ilp == 8) rc = DTE_LAUNCH(8, 1, true);
            else if (pl.pair == 4) rc = DTE_LAUNCH(2, 4, true);
            else if (pl.pair == 2 && pl.ilp == 2) rc = DTE_LAUNCH(2, 2, true);
            else if (pl.pair == 2) rc = DTE_LAUNCH(4, 2, true);
            else rc = DTE_LAUNCH(4, 1, true);
        } else {
            if (pl.ilp == 8) rc = DTE_LAUNCH(8, 1, false);
            else rc = DTE_LAUNCH(4, 1, false);
        }
#undef DTE_LAUNCH
    }
    if (rc == cudaSuccess) d.kernel_launches++;
    return rc;
}

inline void kernel_name(const Dev& d, const Tune& tune, int want, char* buf, size_t len) {
    const Plan pl = make_plan(d, tune, want);
    if (pl.variant == KERNEL_GENERIC) {
        snprintf(buf, len, "dt_walk_generic<%d>", pl.wide ? 1 : 0);
    } else {
        const bool staged = pl.variant == KERNEL_TILE_STAGED;
        snprintf(buf, len, "dt_walk_tile<%d, %d, %d, %d, %d> warps=%d stages=%d phased=%d threads=%d smem=%zu",
                 pl.ilp, pl.pair, staged ? 1 : 0, pl.wide ? 1 : 0, pl.thread_bound(), pl.nwarps, pl.nstages,
                 phased_level(d, tune, pl) != 0xFFFFFFFFu ? 1 : 0, pl.threads(), pl.smem);
    }
}

}  // namespace dte
