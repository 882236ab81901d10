// dte_engine.cu — host side of libdte.so: the C ABI of include/dte.h, the CSR file, the PCIe line
// stream framer, the ensemble repacker and the kernel launches.  No torch, no Python: plain CUDA
// runtime.  There is NO CPU compute path in this file — every score comes from a CUDA kernel.
//
// Reference behaviour mirrored here (paths relative to the reference root):
//   CSR decode ............... rtl/DTEngine/EngineCSR.sv:113-125,189-306
//   stream order / counting .. rtl/DTEngine/PCIeReceiver.sv:136-139,205-316
//   tree / tuple framing ..... rtl/DTEngine/InputDistributor.sv:248-296
//   result packing ........... rtl/DTEngine/ResultsCombiner.sv:132-162
//   completion ............... rtl/DTEngine/DTInference.sv:633-663
#include "../../include/dte.h"
#include "dte_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace dte;

namespace {

constexpr int kNumBuf = 3;   // device chunk buffers of the host-path pipeline

// experiment knobs, env DTE_TUNE="ilp=4,pair=2,stages=1,warps=10,phased=1,fill=0,chunk=65536"
// (ilp: trees per warp, pair: warps per tuple group, stages: ring depth, warps: consumer-warp cap,
//  phased: 0 off / k>=1 on with part A ending k levels earlier, fill=1: one bulk copy per tree, chunk: host-path tuples per chunk)
struct Tune {
    int ilp = 0, stages = 0, warps = 0, pair = 0, fill = 0, phased = -1;
    size_t chunk = 0;
};

struct Plan {               // how the next walk will be launched
    int variant = DTE_KERNEL_GENERIC;
    int ilp = 8, pair = 1, nstages = 0, nwarps = 0;   // nwarps = consumer warps = groups * pair
    bool wide = false;
    size_t smem = 0;
};

}  // namespace

struct dte_engine {
    int dev = 0;
    int sm_count = 0;
    int smem_optin = 0;
    cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    bool timing_pending = false;
    std::string err;

    // ---- CSR file (EngineCSR.sv) ----
    uint64_t regs[12] = {0};          // 200..211 as written
    // decoded at `start` / load
    uint32_t D = 0, K = 0, S = 0, missing = 0, w_cls = 0, f_cls = 0, tuple_cls = 0;

    // ---- stream state (PCIeReceiver.sv FSM) ----
    enum { ST_IDLE = 0, ST_TREES = 1, ST_WAIT = 2, ST_DATA = 3 } state = ST_IDLE;
    uint64_t lines_received = 0;
    uint32_t node_index = 0;                    // position in devices_list; 0 = host node (PCIeReceiver.sv:160-178)
    uint64_t cur_w = 0, cur_f = 0, cur_d = 0;   // currWCount / currFCount / currDCount
    uint32_t cur_dev = 0;                       // currDevID
    uint64_t local_w_lines = 0;                 // weight lines kept by this node
    std::vector<unsigned char> tree_lines;      // weights then findexes kept by this node, as received
    std::vector<unsigned char> tuple_partial;   // bytes of an incomplete tuple
    std::vector<float> result_words;            // scores not yet returned, tuple order
    size_t result_read_pos = 0;                 // words already handed out
    uint64_t result_lines_out = 0;              // lines produced since start (for process_done)
    uint32_t out_cl_count = 0;                  // pcie_out_cl_count (DTInference.sv:632-646)

    // ---- resident ensemble ----
    uint32_t T = 0;                   // trees resident
    uint32_t Tpad = 0;                // padded to a multiple of 8
    uint32_t Dtop = 0, top_stride = 1, nb = 1;
    bool wide = false;
    uint2* d_top = nullptr;
    uint4* d_bottom = nullptr;
    uint64_t ensemble_bytes = 0;

    // ---- host-path device buffers ----
    void* d_tup[kNumBuf] = {nullptr, nullptr, nullptr};
    float* d_sc[kNumBuf] = {nullptr, nullptr, nullptr};
    uint8_t* d_lb[kNumBuf] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d[kNumBuf], ev_comp[kNumBuf], ev_d2h[kNumBuf];
    size_t chunk_cap = 0;             // tuples per buffer
    uint32_t chunk_F = 0;

    // ---- counters ----
    int forced_variant = DTE_KERNEL_AUTO;
    Tune tune;
    uint64_t kernel_launches = 0;
    double prog_ns = 0, exec_ns = 0, last_walk_ms = 0;
    uint64_t tuples_in = 0, tuples_out = 0;
};

namespace {

int fail(dte_engine* e, int code, const char* fmt, ...) {
    if (e) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        e->err = buf;
    }
    return code;
}

#define CUDA_TRY(e, call)                                                                       \
    do {                                                                                        \
        cudaError_t _st = (call);                                                               \
        if (_st != cudaSuccess)                                                                 \
            return fail((e), DTE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

void parse_tune(Tune& t) {
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

// Decode the geometry registers (EngineCSR.sv:218-233) and check them.
int decode_csr(dte_engine* e) {
    const uint64_t r204 = e->regs[4], r205 = e->regs[5];
    e->w_cls = (uint32_t)((r204 >> 16) & 0xFFFF);
    e->f_cls = (uint32_t)((r204 >> 32) & 0xFFFF);
    e->tuple_cls = (uint32_t)((r204 >> 48) & 0xFFFF);
    e->missing = (uint32_t)(r205 & 0xFFFFFFFFu);
    e->D = (uint32_t)((r205 >> 32) & 0xF);
    e->S = (uint32_t)((r205 >> 36) & 0xFF);
    e->K = (uint32_t)((r205 >> 44) & 0xF);
    if (e->D < 1) return fail(e, DTE_ERR_CONFIG, "reg 205: num_levels_per_tree = 0");
    if (e->K < 1 || e->K > 8) return fail(e, DTE_ERR_CONFIG, "reg 205: num_clusters_per_tuple = %u, need 1..8", e->K);
    if (e->S < 1) return fail(e, DTE_ERR_CONFIG, "reg 205: num_trees_per_pu = 0");
    if (e->tuple_cls < 1 || e->tuple_cls > 512) return fail(e, DTE_ERR_CONFIG, "reg 204: tuple_numcls = %u, need 1..512", e->tuple_cls);
    if ((uint64_t)e->w_cls * 4 < (2ull << e->D) - 1)
        return fail(e, DTE_ERR_CONFIG, "reg 204: tree_weights_numcls = %u too small for %u levels", e->w_cls, e->D);
    if ((uint64_t)e->f_cls * 8 < (1ull << e->D) - 1)
        return fail(e, DTE_ERR_CONFIG, "reg 204: tree_feature_index_numcls = %u too small for %u levels", e->f_cls, e->D);
    return DTE_OK;
}

void free_ensemble(dte_engine* e) {
    if (e->d_top) cudaFree(e->d_top);
    if (e->d_bottom) cudaFree(e->d_bottom);
    e->d_top = nullptr;
    e->d_bottom = nullptr;
    e->T = e->Tpad = 0;
    e->ensemble_bytes = 0;
}

// Repack trees [first, first+count) of the two reference streams into the device layout and upload.
//   stream layout (R1 in SURVEY.md): per tree, W = heap array of 2^(D+1)-1 fp32 words padded to
//   w_cls lines; FI = 2^D-1 u16 padded to f_cls lines (DTPU.sv:282-338,579-596).
int load_ensemble(dte_engine* e, const unsigned char* wl, size_t n_wl, const unsigned char* fl, size_t n_fl,
                  uint32_t first, uint32_t count) {
    auto t_begin = std::chrono::steady_clock::now();
    int rc = decode_csr(e);
    if (rc) return rc;
    if (n_wl % e->w_cls) return fail(e, DTE_ERR_ARG, "weights stream: %zu lines is not a multiple of %u lines per tree", n_wl, e->w_cls);
    const size_t T_all = n_wl / e->w_cls;
    if (n_fl != T_all * e->f_cls)
        return fail(e, DTE_ERR_ARG, "feature-index stream: %zu lines, expected %zu (= %zu trees x %u)", n_fl, T_all * e->f_cls, T_all, e->f_cls);
    if (count == 0 && first == 0) count = (uint32_t)T_all;
    if ((size_t)first + count > T_all || count == 0)
        return fail(e, DTE_ERR_ARG, "tree chunk [%u, %u) outside the %zu trees of the stream", first, first + count, T_all);

    const uint32_t D = e->D, F = e->tuple_cls * 4;
    const uint32_t Dk = std::max(D, 2u);
    const uint32_t Dtop = Dk - 2;
    const uint32_t nb = 1u << Dtop;
    const uint32_t top_stride = std::max(1u, 1u << Dtop);
    const uint32_t Tpad = (count + 7u) & ~7u;
    const size_t wstride = (size_t)e->w_cls * 4, fstride = (size_t)e->f_cls * 8;
    const uint32_t* Wall = reinterpret_cast<const uint32_t*>(wl);
    const uint16_t* Fall = reinterpret_cast<const uint16_t*>(fl);

    // pass 1: contract check + widest feature index
    uint32_t max_f = 0;
    for (uint32_t t = 0; t < count; ++t) {
        const uint16_t* fi = Fall + (size_t)(first + t) * fstride;
        for (uint32_t i = 0; i + 1 < (1u << D); ++i) {
            const uint32_t f = fi[i] & 0x7FFu;
            if (f >= F) return fail(e, DTE_ERR_UNSUPPORTED, "tree %u node %u: feature index %u >= %u features", first + t, i, f, F);
            if (fi[i] & 0x4000u)
                return fail(e, DTE_ERR_UNSUPPORTED, "tree %u node %u: bit 14 (next-node-is-leaf) set; complete trees only", first + t, i);
            max_f = std::max(max_f, f);
        }
    }
    const bool wide = max_f >= 512;
    const uint32_t BV = wide ? 4 : 2;

    std::vector<uint2> top((size_t)Tpad * top_stride, make_uint2(0u, 8u << 16));
    std::vector<uint4> bot((size_t)Tpad * nb * BV, make_uint4(0, 0, 0, 0));
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
        }
        uint2* tp = top.data() + (size_t)t * top_stride;
        for (uint32_t n = 0; n + 1 < (1u << Dtop); ++n) {
            const uint32_t f = Fk[n] & 0x7FFu, mr = (Fk[n] >> 13) & 1u;
            tp[n] = make_uint2(Wk[n], f | ((8u + 8u * mr) << 16));
        }
        uint4* bp = bot.data() + (size_t)t * nb * BV;
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

    CUDA_TRY(e, cudaSetDevice(e->dev));
    free_ensemble(e);
    CUDA_TRY(e, cudaMalloc(&e->d_top, top.size() * sizeof(uint2)));
    CUDA_TRY(e, cudaMalloc(&e->d_bottom, bot.size() * sizeof(uint4)));
    CUDA_TRY(e, cudaMemcpy(e->d_top, top.data(), top.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(e->d_bottom, bot.data(), bot.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    e->T = count;
    e->Tpad = Tpad;
    e->Dtop = Dtop;
    e->top_stride = top_stride;
    e->nb = nb;
    e->wide = wide;
    e->ensemble_bytes = top.size() * sizeof(uint2) + bot.size() * sizeof(uint4);
    e->prog_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t_begin).count();
    return DTE_OK;
}

// ---- launch planning ---------------------------------------------------------------------------
size_t tile_smem(uint32_t F, int groups, int trees_per_stage, int nstages, uint32_t top_stride) {
    return (size_t)kHdrBytes + (size_t)nstages * trees_per_stage * top_stride * 8 + (size_t)F * 32 * groups * 4;
}

Plan make_plan(const dte_engine* e) {
    Plan p;
    p.wide = e->wide;
    const uint32_t F = e->tuple_cls * 4;
    const size_t budget = (size_t)e->smem_optin;
    // tuple groups (32 tuples, F*128 B of shared memory each) that fit next to the ring
    auto max_groups = [&](int ilp, int pair, int nstages) -> int {
        const size_t fixed = tile_smem(F, 0, ilp * pair, nstages, e->top_stride);
        if (fixed >= budget) return 0;
        const int warp_cap = (pair == 4) ? 20 : (ilp == 8 ? 8 : 12);   // __launch_bounds__ of dt_walk_tile (+1 producer warp)
        int g = (int)std::min<size_t>((size_t)(warp_cap / pair), (budget - fixed) / ((size_t)F * 128));
        if (e->tune.warps) g = std::min(g, std::max(1, e->tune.warps / pair));
        return g;
    };
    const int want = e->forced_variant;
    const bool can_stage = e->Dtop >= 3;          // below that there is nothing worth staging
    // staged candidates {trees per warp, warps per tuple group, ring stages}.  Measured on B200
    // (profiles/r01_summary.md); the first candidate that fits with the most walks in flight wins.
    const int cand[6][3] = {{4, 2, 1}, {2, 4, 1}, {8, 1, 1}, {8, 1, 2}, {4, 1, 2}, {4, 1, 1}};
    Plan staged;
    if (can_stage) {
        int best = 0;
        for (auto& c : cand) {
            int ilp = e->tune.ilp ? e->tune.ilp : c[0];
            int pair = e->tune.pair ? e->tune.pair : c[1];
            int st = e->tune.stages ? e->tune.stages : c[2];
            if (pair == 4) ilp = 2; else if (pair == 2) ilp = (ilp == 2) ? 2 : 4; else { pair = 1; if (ilp != 4 && ilp != 8) ilp = 8; }
            st = std::max(1, std::min(st, 4));
            const int g = max_groups(ilp, pair, st);
            const int score = g * pair * ilp * 8 + (pair == 2 ? 4 : 0) - (pair == 4 ? 4 : 0) + st;
            if (g >= 1 && score > best) {
                best = score;
                staged.variant = DTE_KERNEL_TILE_STAGED;
                staged.ilp = ilp; staged.pair = pair; staged.nstages = st; staged.nwarps = g * pair; staged.wide = e->wide;
                staged.smem = tile_smem(F, g, ilp * pair, st, e->top_stride);
            }
            if (e->tune.ilp && e->tune.stages && e->tune.pair) break;
        }
    }
    Plan tile;
    {
        const int ilp = (e->tune.ilp == 4) ? 4 : 8;
        const int g = max_groups(ilp, 1, 0);
        if (g >= 1) {
            tile.variant = DTE_KERNEL_TILE; tile.ilp = ilp; tile.pair = 1; tile.nstages = 0; tile.nwarps = g; tile.wide = e->wide;
            tile.smem = tile_smem(F, g, ilp, 0, e->top_stride);
        }
    }
    if (want == DTE_KERNEL_TILE_STAGED && staged.nwarps >= 1) return staged;
    if (want == DTE_KERNEL_TILE && tile.nwarps >= 1) return tile;
    if (want == DTE_KERNEL_GENERIC) return p;
    // AUTO (or a forced variant that does not fit): staged > tile > generic
    if (staged.nwarps >= 2) return staged;
    if (tile.nwarps >= 1) return tile;
    return p;
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

int launch_walk(dte_engine* e, const void* d_tuples, size_t n, float* d_scores, uint8_t* d_labels, cudaStream_t st,
                bool accumulate = false) {
    if (!e->d_top) return fail(e, DTE_ERR_STATE, "no ensemble loaded");
    if (n == 0) return DTE_OK;
    const Plan pl = make_plan(e);
    WalkParams wp;
    wp.top = e->d_top;
    wp.bottom = e->d_bottom;
    wp.tuples = static_cast<const float*>(d_tuples);
    wp.scores = d_scores;
    wp.labels = d_labels;
    wp.n = n;
    wp.F = e->tuple_cls * 4;
    wp.Dtop = e->Dtop;
    wp.top_stride = e->top_stride;
    wp.nb = e->nb;
    // slots beyond S are never issued (DTPU.sv:519-531); groups past the last tree add exact zeros
    wp.groups = (uint32_t)std::min<uint64_t>((uint64_t)e->S * e->K, e->Tpad / 8);
    wp.K = e->K;
    wp.missing = e->missing;
    wp.nwarps = (uint32_t)pl.nwarps;
    wp.nstages = (uint32_t)pl.nstages;
    wp.accumulate = accumulate ? 1u : 0u;
    wp.wide_rows = (wp.F % 8 == 0 && (reinterpret_cast<uintptr_t>(d_tuples) & 31u) == 0) ? 1u : 0u;
    wp.fill_split = e->tune.fill ? 1u : 0u;      // DTE_TUNE fill=1: one bulk copy per tree instead of one per stage
    // phased refill needs >= 3 staged levels and at most 4 ring stages (16 mbarriers in the header)
    // Measured (profiles/r01_summary.md): +7 % at D = 12 (64 KiB stage), -4 % at D <= 10 (<= 16 KiB stage, the
    // refill is already cheap there and the extra barrier hand-offs cost more than they hide).
    const bool phased = e->tune.phased >= 1 || (e->tune.phased == -1 && e->Dtop >= 10);
    // part A = levels 0..Lw; DTE_TUNE phased=k moves the split k-1 levels further up (smaller part A)
    const uint32_t up = e->tune.phased > 1 ? (uint32_t)e->tune.phased - 1 : 0;
    wp.Lw = (phased && e->Dtop >= 3 + up && pl.nstages <= 4) ? e->Dtop - 3 - up : 0xFFFFFFFFu;
    wp.tiles = 0;
    cudaError_t rc;
    if (pl.variant == DTE_KERNEL_GENERIC) {
        const int threads = 128;
        const unsigned long long blocks = (n + threads - 1) / threads;
        if (blocks > 0x7FFFFFFFull) return fail(e, DTE_ERR_ARG, "batch too large for one launch");
        if (pl.wide) dt_walk_generic<true><<<(unsigned)blocks, threads, 0, st>>>(wp);
        else dt_walk_generic<false><<<(unsigned)blocks, threads, 0, st>>>(wp);
        rc = cudaGetLastError();
    } else {
        const size_t M = 32ull * (pl.nwarps / pl.pair);
        const unsigned long long tiles = (n + M - 1) / M;
        if (tiles > 0xFFFFFFFFull) return fail(e, DTE_ERR_ARG, "batch too large for one launch");
        wp.tiles = (uint32_t)tiles;
        const int grid = (int)std::min<unsigned long long>(tiles, (unsigned long long)e->sm_count);
        const bool staged = pl.variant == DTE_KERNEL_TILE_STAGED;
        const int threads = 32 * (pl.nwarps + (staged ? 1 : 0));
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
    if (rc != cudaSuccess) return fail(e, DTE_ERR_CUDA, "walk kernel launch failed: %s", cudaGetErrorString(rc));
    e->kernel_launches++;
    return DTE_OK;
}

int ensure_chunk_buffers(dte_engine* e, size_t cap, uint32_t F) {
    if (e->chunk_cap >= cap && e->chunk_F == F) return DTE_OK;
    for (int b = 0; b < kNumBuf; ++b) {
        if (e->d_tup[b]) cudaFree(e->d_tup[b]);
        if (e->d_sc[b]) cudaFree(e->d_sc[b]);
        if (e->d_lb[b]) cudaFree(e->d_lb[b]);
        e->d_tup[b] = nullptr; e->d_sc[b] = nullptr; e->d_lb[b] = nullptr;
    }
    e->chunk_cap = 0;
    for (int b = 0; b < kNumBuf; ++b) {
        CUDA_TRY(e, cudaMalloc(&e->d_tup[b], cap * F * 4));
        CUDA_TRY(e, cudaMalloc(&e->d_sc[b], cap * 4));
        CUDA_TRY(e, cudaMalloc(&e->d_lb[b], cap));
    }
    e->chunk_cap = cap;
    e->chunk_F = F;
    return DTE_OK;
}

// Host buffers in, host buffers out: H2D of chunk i+1, walk of chunk i and D2H of chunk i-1 overlap.
int infer_host(dte_engine* e, const unsigned char* h_tuples, size_t n, float* h_scores, uint8_t* h_labels) {
    if (!e->d_top) return fail(e, DTE_ERR_STATE, "no ensemble loaded");
    if (n == 0) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    const uint32_t F = e->tuple_cls * 4;
    size_t chunk = e->tune.chunk ? e->tune.chunk : std::max<size_t>(4096, (64ull << 20) / (F * 4));
    chunk = std::min(chunk, n);
    int rc = ensure_chunk_buffers(e, chunk, F);
    if (rc) return rc;
    const size_t nchunks = (n + chunk - 1) / chunk;
    CUDA_TRY(e, cudaEventRecord(e->ev_t0, e->s_main));
    for (size_t i = 0; i < nchunks; ++i) {
        const int b = (int)(i % kNumBuf);
        const size_t off = i * chunk, cnt = std::min(chunk, n - off);
        if (i >= (size_t)kNumBuf) {
            CUDA_TRY(e, cudaStreamWaitEvent(e->s_h2d, e->ev_comp[b], 0));   // tuple buffer b free again
            CUDA_TRY(e, cudaStreamWaitEvent(e->s_main, e->ev_d2h[b], 0));   // score buffer b drained
        }
        CUDA_TRY(e, cudaMemcpyAsync(e->d_tup[b], h_tuples + off * F * 4, cnt * F * 4, cudaMemcpyHostToDevice, e->s_h2d));
        CUDA_TRY(e, cudaEventRecord(e->ev_h2d[b], e->s_h2d));
        CUDA_TRY(e, cudaStreamWaitEvent(e->s_main, e->ev_h2d[b], 0));
        rc = launch_walk(e, e->d_tup[b], cnt, e->d_sc[b], h_labels ? e->d_lb[b] : nullptr, e->s_main);
        if (rc) return rc;
        CUDA_TRY(e, cudaEventRecord(e->ev_comp[b], e->s_main));
        CUDA_TRY(e, cudaStreamWaitEvent(e->s_d2h, e->ev_comp[b], 0));
        CUDA_TRY(e, cudaMemcpyAsync(h_scores + off, e->d_sc[b], cnt * 4, cudaMemcpyDeviceToHost, e->s_d2h));
        if (h_labels) CUDA_TRY(e, cudaMemcpyAsync(h_labels + off, e->d_lb[b], cnt, cudaMemcpyDeviceToHost, e->s_d2h));
        CUDA_TRY(e, cudaEventRecord(e->ev_d2h[b], e->s_d2h));
    }
    CUDA_TRY(e, cudaEventRecord(e->ev_t1, e->s_main));
    CUDA_TRY(e, cudaStreamSynchronize(e->s_h2d));
    CUDA_TRY(e, cudaStreamSynchronize(e->s_main));
    CUDA_TRY(e, cudaStreamSynchronize(e->s_d2h));
    float ms = 0;
    CUDA_TRY(e, cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
    e->last_walk_ms = ms;
    e->exec_ns += (double)ms * 1e6;
    e->tuples_in += n;
    e->tuples_out += n;
    return DTE_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* dte_version(void) { return "dte-b200 0.1 (sm_100a)"; }

int dte_create(dte_t** engine, int gpu_ordinal) {
    if (!engine) return DTE_ERR_ARG;
    *engine = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DTE_ERR_CUDA;   // no CPU fallback
    if (gpu_ordinal < 0 || gpu_ordinal >= ndev) return DTE_ERR_ARG;
    dte_engine* e = new (std::nothrow) dte_engine();
    if (!e) return DTE_ERR_NOMEM;
    e->dev = gpu_ordinal;
    parse_tune(e->tune);
    bool ok = cudaSetDevice(gpu_ordinal) == cudaSuccess &&
              cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, gpu_ordinal) == cudaSuccess &&
              cudaDeviceGetAttribute(&e->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, gpu_ordinal) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->s_main, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&e->ev_t0) == cudaSuccess && cudaEventCreate(&e->ev_t1) == cudaSuccess;
    for (int b = 0; ok && b < kNumBuf; ++b)
        ok = cudaEventCreateWithFlags(&e->ev_h2d[b], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->ev_comp[b], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->ev_d2h[b], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        delete e;
        return DTE_ERR_CUDA;
    }
    *engine = e;
    return DTE_OK;
}

int dte_destroy(dte_t* e) {
    if (!e) return DTE_ERR_ARG;
    cudaSetDevice(e->dev);
    cudaDeviceSynchronize();
    free_ensemble(e);
    for (int b = 0; b < kNumBuf; ++b) {
        if (e->d_tup[b]) cudaFree(e->d_tup[b]);
        if (e->d_sc[b]) cudaFree(e->d_sc[b]);
        if (e->d_lb[b]) cudaFree(e->d_lb[b]);
        cudaEventDestroy(e->ev_h2d[b]);
        cudaEventDestroy(e->ev_comp[b]);
        cudaEventDestroy(e->ev_d2h[b]);
    }
    cudaEventDestroy(e->ev_t0);
    cudaEventDestroy(e->ev_t1);
    cudaStreamDestroy(e->s_main);
    cudaStreamDestroy(e->s_h2d);
    cudaStreamDestroy(e->s_d2h);
    delete e;
    return DTE_OK;
}

const char* dte_last_error(const dte_t* e) { return e ? e->err.c_str() : "null engine"; }

int dte_softreg_write(dte_t* e, uint32_t addr, uint64_t data) {
    if (!e) return DTE_ERR_ARG;
    if (addr < 200 || addr > 211) return DTE_OK;          // writes elsewhere are ignored (EngineCSR.sv:190)
    if (addr != 200) {
        e->regs[addr - 200] = data;
        return DTE_OK;
    }
    if (!(data & 1)) return DTE_OK;
    // `start` (EngineCSR.sv:191-193): Core FSM, counters and schedules reset (Core.sv:168-187);
    // the PU tree memories are NOT cleared (DTPU.sv:307-319) -> the resident ensemble stays.
    e->lines_received = 0;
    e->cur_w = e->cur_f = e->cur_d = 0;
    e->cur_dev = 0;
    e->local_w_lines = 0;
    e->tree_lines.clear();
    e->tuple_partial.clear();
    e->result_words.clear();
    e->result_read_pos = 0;
    e->result_lines_out = 0;
    e->out_cl_count = 0;
    const uint64_t r201 = e->regs[1];
    const bool data_distributed = r201 & 1, host_node = (r201 >> 1) & 1, rx_enabled = (r201 >> 6) & 1;
    e->state = dte_engine::ST_IDLE;
    if (rx_enabled) {                                       // PCIeReceiver.sv:218-227
        if (host_node) e->state = dte_engine::ST_TREES;
        else if (data_distributed) e->state = dte_engine::ST_DATA;
    }
    if (e->state == dte_engine::ST_DATA) {
        int rc = decode_csr(e);
        if (rc) { e->state = dte_engine::ST_IDLE; return rc; }
        if (!e->d_top) { e->state = dte_engine::ST_IDLE; return fail(e, DTE_ERR_STATE, "start in data-only mode without a resident ensemble"); }
    }
    return DTE_OK;
}

int dte_softreg_read(dte_t* e, uint32_t addr, uint64_t* data) {
    if (!e || !data) return DTE_ERR_ARG;
    switch (addr) {
        case 220: *data = (uint64_t)e->state; break;                          // pcie_receiver_fsm_state
        case 221: *data = e->lines_received & 0xFFFFFFFFull; break;           // pcie_numcls_received
        case 222: *data = (uint64_t)e->prog_ns; break;                        // progCycles (ns here)
        case 223: *data = (uint64_t)e->exec_ns; break;                        // execCycles (ns here)
        case 224: case 225: case 226: *data = 0; break;                       // SL3 tx counters: no ring traffic
        case 121: *data = e->tuples_in & 0xFFFFFFFFull; break;                // appStatus: tuples received
        case 122: *data = e->tuples_out & 0xFFFFFFFFull; break;               //            tuples emitted
        case 123: *data = e->result_lines_out & 0xFFFFFFFFull; break;         //            result lines
        case 124: *data = e->kernel_launches & 0xFFFFFFFFull; break;
        case 125: *data = e->T; break;
        case 126: *data = 0; break;                                            // res_lines_lost: never
        default: *data = 0xFFFFFFFFFFFFFFFFull; break;                         // EngineCSR.sv:123
    }
    return DTE_OK;
}

int dte_stream_write(dte_t* e, const void* cl128, size_t n_lines) {
    if (!e || (!cl128 && n_lines)) return DTE_ERR_ARG;
    const unsigned char* p = static_cast<const unsigned char*>(cl128);
    while (n_lines) {
        if (e->state == dte_engine::ST_TREES) {
            const uint64_t total = e->regs[2] & 0xFFFFFFFFull, wtotal = e->regs[2] >> 32;   // reg 202
            if (total == 0 || wtotal == 0 || wtotal >= total)
                return fail(e, DTE_ERR_CONFIG, "reg 202: total_num_trees_cls=%llu total_num_weights_cls=%llu",
                            (unsigned long long)total, (unsigned long long)wtotal);
            // Multi-node chunking (PCIeReceiver.sv:241-264): unless broadcast_trees, the weights stream
            // is cut every numcls_local_weights lines and the index stream every numcls_local_findexes
            // lines, chunk i going to devices_list[i % numDevs]; entry 0 is the host node itself
            // (:160-178).  Every process replays the SAME stream; engine g keeps the chunks of entry g.
            const uint64_t r201 = e->regs[1], r203 = e->regs[3];
            const bool multi = (r201 >> 5) & 1, bcast_trees = (r201 >> 3) & 1;
            const uint64_t chunk_w = (r203 & 0xFFFF) + 1, chunk_f = (r203 >> 16) & 0xFFFF;
            const uint32_t ndev = (uint32_t)std::max<uint64_t>(1, (r203 >> 32) & 0xFF);
            const bool deal = multi && !bcast_trees && ndev > 1;
            if (deal && (chunk_f == 0 || e->node_index >= ndev))
                return fail(e, DTE_ERR_CONFIG, "reg 203: numcls_local_findexes=%llu numDevs=%u node=%u",
                            (unsigned long long)chunk_f, ndev, e->node_index);
            const size_t take = (size_t)std::min<uint64_t>(n_lines, total - e->lines_received);
            if (!deal) {
                e->tree_lines.insert(e->tree_lines.end(), p, p + take * 16);
                e->local_w_lines += std::min<uint64_t>(take, wtotal > e->lines_received ? wtotal - e->lines_received : 0);
            } else {
                for (size_t i = 0; i < take; ++i) {
                    const bool is_w = e->lines_received + i < wtotal;
                    if (e->cur_dev == e->node_index) {
                        e->tree_lines.insert(e->tree_lines.end(), p + i * 16, p + i * 16 + 16);
                        if (is_w) e->local_w_lines++;
                    }
                    uint64_t& cnt = is_w ? e->cur_w : e->cur_f;
                    if (++cnt == (is_w ? chunk_w : chunk_f)) {
                        cnt = 0;
                        e->cur_dev = (e->cur_dev + 1 == ndev) ? 0 : e->cur_dev + 1;
                    }
                    if (is_w && e->lines_received + i + 1 == wtotal) { e->cur_dev = 0; e->cur_w = 0; }   // index stream restarts at the host
                }
            }
            e->lines_received += take;
            p += take * 16;
            n_lines -= take;
            if (e->lines_received == total) {
                // prog_mode = (numcls_received < total_num_weights_cls), PCIeReceiver.sv:136-139
                const size_t lw = (size_t)e->local_w_lines, lf = e->tree_lines.size() / 16 - lw;
                int rc = (lw && lf) ? load_ensemble(e, e->tree_lines.data(), lw, e->tree_lines.data() + lw * 16, lf, 0, 0)
                                    : fail(e, DTE_ERR_CONFIG, "node %u received no trees", e->node_index);
                e->tree_lines.clear();
                e->tree_lines.shrink_to_fit();
                if (rc) { e->state = dte_engine::ST_IDLE; return rc; }
                e->state = dte_engine::ST_DATA;                                 // WAIT_DATA -> RECEIVE_DATA
                e->cur_dev = 0;
                e->cur_d = 0;
            }
        } else if (e->state == dte_engine::ST_DATA) {
            const size_t tbytes = (size_t)e->tuple_cls * 16;
            // Data dealing (PCIeReceiver.sv:298-307): unless broadcast_data, batches of core_data_batch_cls
            // lines go round-robin over devices_list; this engine keeps the batches of its own entry.
            const uint64_t r201d = e->regs[1];
            const bool multi_d = (r201d >> 5) & 1, bcast_data = (r201d >> 2) & 1, distributed = r201d & 1;
            const uint32_t ndev_d = (uint32_t)std::max<uint64_t>(1, (e->regs[3] >> 32) & 0xFF);
            std::vector<unsigned char> mine;
            const size_t lines_in = n_lines;
            if (multi_d && !bcast_data && !distributed && ndev_d > 1) {
                const uint64_t batch = r201d >> 32;
                if (batch == 0 || batch % e->tuple_cls)
                    return fail(e, DTE_ERR_CONFIG, "reg 201: core_data_batch_cls=%llu must be a positive multiple of tuple_numcls=%u",
                                (unsigned long long)batch, e->tuple_cls);
                mine.reserve(n_lines * 16 / ndev_d + 16);
                for (size_t i = 0; i < n_lines; ++i) {
                    if (e->cur_dev == e->node_index) mine.insert(mine.end(), p + i * 16, p + i * 16 + 16);
                    if (++e->cur_d == batch) { e->cur_d = 0; e->cur_dev = (e->cur_dev + 1 == ndev_d) ? 0 : e->cur_dev + 1; }
                }
                p = mine.data();
                n_lines = mine.size() / 16;
            }
            // frame tuples by counting tuple_numcls lines (InputDistributor.sv:276-296)
            std::vector<unsigned char>& part = e->tuple_partial;
            const unsigned char* src = p;
            size_t nbytes = n_lines * 16;
            std::vector<unsigned char> joined;
            if (!part.empty()) {
                joined.reserve(part.size() + nbytes);
                joined.insert(joined.end(), part.begin(), part.end());
                joined.insert(joined.end(), p, p + nbytes);
                src = joined.data();
                nbytes = joined.size();
            }
            const size_t ntup = nbytes / tbytes;
            e->lines_received += lines_in;
            if (ntup) {
                const size_t old = e->result_words.size();
                e->result_words.resize(old + ntup);
                int rc = infer_host(e, src, ntup, e->result_words.data() + old, nullptr);
                if (rc) { e->result_words.resize(old); return rc; }
            }
            std::vector<unsigned char> rest(src + ntup * tbytes, src + nbytes);
            part.swap(rest);
            n_lines = 0;
        } else {
            return fail(e, DTE_ERR_STATE, "stream_write while the receiver is idle (write reg 200 first)");
        }
    }
    return DTE_OK;
}

int dte_stream_read_packets(dte_t* e, void* cl128, uint8_t* last_flags, size_t max_lines, size_t* got) {
    if (!e || !got || (!cl128 && max_lines)) return DTE_ERR_ARG;
    // 4 consecutive results per line, word j = tuple 4m+j (ResultsCombiner.sv:132-162)
    const size_t avail_lines = (e->result_words.size() - e->result_read_pos) / 4;
    const size_t n = std::min(avail_lines, max_lines);
    if (n) memcpy(cl128, e->result_words.data() + e->result_read_pos, n * 16);
    if (last_flags) {
        // `last` closes a PCIe packet every pcie_out_packet_numcls lines: reg 206[55:48], compared as the
        // 8-bit "minus one" copy (EngineCSR.sv:242, DTInference.sv:632-646,659-663)
        const uint32_t pkt_m1 = (uint32_t)(((e->regs[6] >> 48) & 0xFF) - 1) & 0xFF;
        for (size_t i = 0; i < n; ++i) {
            last_flags[i] = e->out_cl_count == pkt_m1;
            e->out_cl_count = (e->out_cl_count == pkt_m1) ? 0 : ((e->out_cl_count + 1) & 0xFF);
        }
    }
    e->result_read_pos += n * 4;
    e->result_lines_out += n;
    *got = n;
    if (e->result_read_pos > (1u << 20)) {              // compact the queue now and then
        e->result_words.erase(e->result_words.begin(), e->result_words.begin() + (ptrdiff_t)e->result_read_pos);
        e->result_read_pos = 0;
    }
    return DTE_OK;
}

int dte_stream_read(dte_t* e, void* cl128, size_t max_lines, size_t* got) {
    return dte_stream_read_packets(e, cl128, nullptr, max_lines, got);
}

int dte_process_done(dte_t* e, int* done) {
    if (!e || !done) return DTE_ERR_ARG;
    const uint64_t want = e->regs[7] & 0xFFFFFFFFull;                           // reg 207[31:0]
    const uint64_t produced = e->result_lines_out + (e->result_words.size() - e->result_read_pos) / 4;
    *done = (want != 0 && produced >= want) ? 1 : 0;
    return DTE_OK;
}

int dte_load_ensemble(dte_t* e, const void* weight_cls, size_t n_weight_cls, const void* findex_cls,
                      size_t n_findex_cls, uint32_t first_tree, uint32_t num_local_trees) {
    if (!e || !weight_cls || !findex_cls || !n_weight_cls || !n_findex_cls) return DTE_ERR_ARG;
    return load_ensemble(e, static_cast<const unsigned char*>(weight_cls), n_weight_cls,
                         static_cast<const unsigned char*>(findex_cls), n_findex_cls, first_tree, num_local_trees);
}

int dte_infer_device(dte_t* e, const void* d_tuples, size_t n, float* d_scores, uint8_t* d_labels, void* cuda_stream) {
    if (!e || (n && (!d_tuples || !d_scores))) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->s_main;
    CUDA_TRY(e, cudaEventRecord(e->ev_t0, st));
    int rc = launch_walk(e, d_tuples, n, d_scores, d_labels, st);
    if (rc) return rc;
    CUDA_TRY(e, cudaEventRecord(e->ev_t1, st));
    e->timing_pending = true;
    e->tuples_in += n;
    e->tuples_out += n;
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

int dte_infer_device_accumulate(dte_t* e, const void* d_tuples, size_t n, float* d_scores_accum, void* cuda_stream) {
    if (!e || (n && (!d_tuples || !d_scores_accum))) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->s_main;
    CUDA_TRY(e, cudaEventRecord(e->ev_t0, st));
    int rc = launch_walk(e, d_tuples, n, d_scores_accum, nullptr, st, true);
    if (rc) return rc;
    CUDA_TRY(e, cudaEventRecord(e->ev_t1, st));
    e->timing_pending = true;
    e->tuples_in += n;
    e->tuples_out += n;
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

// ---- peer-visible buffers (CUDA IPC): the combine target of the fused cross-device reduce ----
int dte_ipc_alloc(dte_t* e, size_t bytes, void** d_ptr, unsigned char handle_out[64]) {
    if (!e || !d_ptr || !handle_out || !bytes) return DTE_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CUDA_TRY(e, cudaSetDevice(e->dev));
    CUDA_TRY(e, cudaMalloc(d_ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t st = cudaIpcGetMemHandle(&h, *d_ptr);
    if (st != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        return fail(e, DTE_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(st));
    }
    memcpy(handle_out, &h, 64);
    return DTE_OK;
}

int dte_ipc_open(dte_t* e, const unsigned char handle[64], void** d_ptr) {
    if (!e || !handle || !d_ptr) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(e, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DTE_OK;
}

int dte_ipc_close(dte_t* e, void* d_ptr, int owner) {
    if (!e || !d_ptr) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    if (owner) CUDA_TRY(e, cudaFree(d_ptr));
    else CUDA_TRY(e, cudaIpcCloseMemHandle(d_ptr));
    return DTE_OK;
}

int dte_infer_host(dte_t* e, const void* h_tuples, size_t n, float* h_scores, uint8_t* h_labels) {
    if (!e || (n && (!h_tuples || !h_scores))) return DTE_ERR_ARG;
    return infer_host(e, static_cast<const unsigned char*>(h_tuples), n, h_scores, h_labels);
}

int dte_labels_device(dte_t* e, const float* d_scores, size_t n, uint8_t* d_labels, void* cuda_stream) {
    if (!e || (n && (!d_scores || !d_labels))) return DTE_ERR_ARG;
    if (!n) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->s_main;
    labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scores, d_labels, n);
    CUDA_TRY(e, cudaGetLastError());
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

int dte_ring_add_device(dte_t* e, const float* d_a, const float* d_b, float* d_out, size_t n, void* cuda_stream) {
    if (!e || (n && (!d_a || !d_b || !d_out))) return DTE_ERR_ARG;
    if (!n) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->s_main;
    ring_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, d_b, d_out, n);
    CUDA_TRY(e, cudaGetLastError());
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

int dte_csr_from_profile(uint32_t n_trees, uint32_t depth_levels, uint32_t tuple_bytes, uint32_t clusters,
                         uint32_t missing_value, uint64_t n_tuples, uint64_t regs_out[8]) {
    if (!regs_out || n_trees == 0 || depth_levels < 1 || depth_levels > 15) return DTE_ERR_ARG;
    if (tuple_bytes == 0 || tuple_bytes % 16 || tuple_bytes / 16 > 512) return DTE_ERR_ARG;
    if (!(clusters == 1 || clusters == 2 || clusters == 4 || clusters == 8)) return DTE_ERR_ARG;
    const uint64_t D = depth_levels, K = clusters;
    const uint64_t w_cls = ((2ull << D) - 1 + 3) / 4;           // 2^(D+1)-1 fp32 words, 4 per line
    const uint64_t f_cls = ((1ull << D) - 1 + 7) / 8;           // 2^D-1 u16 indexes, 8 per line
    const uint64_t t_cls = tuple_bytes / 16;
    const uint64_t S = (n_trees + 8 * K - 1) / (8 * K);         // slots per PU so that 8*K*S >= T
    if (S > 255 || w_cls > 0xFFFF) return DTE_ERR_CONFIG;
    uint64_t prog = 0;
    for (uint64_t b = 0; b < 8; b += K) prog |= 1ull << b;      // one replica group every K clusters
    const uint64_t proc = (1ull << K) - 1;
    const uint64_t wl = (uint64_t)n_trees * w_cls, fl = (uint64_t)n_trees * f_cls;
    regs_out[0] = 0x42ull | (t_cls << 32);                                        // 201: host_node | pcie_receiver_enabled
    regs_out[1] = ((wl + fl) & 0xFFFFFFFFull) | (wl << 32);                       // 202
    regs_out[2] = ((wl - 1) & 0xFFFF) | ((fl & 0xFFFF) << 16) | (1ull << 32);     // 203 (16-bit fields, unused single-device)
    regs_out[3] = prog | (proc << 8) | (w_cls << 16) | (f_cls << 32) | (t_cls << 48);   // 204
    regs_out[4] = (uint64_t)missing_value | (D << 32) | (S << 36) | (K << 44);   // 205
    regs_out[5] = (8ull << 16) | (8ull << 24) | (8ull << 32) | (8ull << 40) | (8ull << 48);   // 206 packet sizes
    regs_out[6] = ((n_tuples / 4) & 0xFFFFFFFFull) | (1ull << 32);                // 207
    regs_out[7] = 0;                                                              // 208 devices_list
    return DTE_OK;
}

int dte_get_info(dte_t* e, dte_info* info) {
    if (!e || !info) return DTE_ERR_ARG;
    if (e->timing_pending) {
        if (cudaEventSynchronize(e->ev_t1) == cudaSuccess) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1) == cudaSuccess) {
                e->last_walk_ms = ms;
                e->exec_ns += (double)ms * 1e6;
            }
        }
        e->timing_pending = false;
    }
    memset(info, 0, sizeof *info);
    info->num_trees = e->T;
    info->num_levels = e->D;
    info->num_features = e->tuple_cls * 4;
    info->clusters = e->K;
    info->trees_per_pu = e->S;
    info->sm_count = (uint32_t)e->sm_count;
    info->ensemble_bytes = e->ensemble_bytes;
    info->kernel_launches = e->kernel_launches;
    info->last_walk_ms = e->last_walk_ms;
    if (e->d_top) {
        Plan pl = make_plan(e);
        info->kernel_variant = (uint32_t)pl.variant;
        info->tuples_per_cta = pl.variant == DTE_KERNEL_GENERIC ? 128u : 32u * (uint32_t)(pl.nwarps / pl.pair);
    }
    return DTE_OK;
}

int dte_kernel_name(dte_t* e, char* buf, size_t len) {
    if (!e || !buf || !len) return DTE_ERR_ARG;
    if (!e->d_top) return fail(e, DTE_ERR_STATE, "no ensemble loaded");
    const Plan pl = make_plan(e);
    if (pl.variant == DTE_KERNEL_GENERIC) {
        snprintf(buf, len, "dt_walk_generic<%d>", pl.wide ? 1 : 0);
    } else {
        const bool staged = pl.variant == DTE_KERNEL_TILE_STAGED;
        const int threads = 32 * (pl.nwarps + (staged ? 1 : 0));
        const int nt_max = pl.pair == 4 ? 672 : (pl.ilp == 8 ? 288 : 416);
        const int nt = (nt_max > 384 && threads <= 384) ? 384 : nt_max;
        const bool phased = staged && (e->tune.phased >= 1 || (e->tune.phased == -1 && e->Dtop >= 10)) && e->Dtop >= 3;
        snprintf(buf, len, "dt_walk_tile<%d, %d, %d, %d, %d> warps=%d stages=%d phased=%d threads=%d smem=%zu",
                 pl.ilp, pl.pair, staged ? 1 : 0, pl.wide ? 1 : 0, nt, pl.nwarps, pl.nstages, phased ? 1 : 0, threads, pl.smem);
    }
    return DTE_OK;
}

int dte_set_node(dte_t* e, uint32_t node_index) {
    if (!e || node_index > 19) return DTE_ERR_ARG;             // devices_list has 20 entries (EngineCSR.sv:250-296)
    e->node_index = node_index;
    return DTE_OK;
}

int dte_set_kernel_variant(dte_t* e, int variant) {
    if (!e || variant < DTE_KERNEL_AUTO || variant > DTE_KERNEL_TILE_STAGED) return DTE_ERR_ARG;
    e->forced_variant = variant;
    return DTE_OK;
}

int dte_synth_tuples_device(dte_t* e, void* d_tuples, uint64_t first_tuple, uint64_t n, uint32_t num_features,
                            uint64_t seed, uint32_t missing_ppm, uint32_t missing_value, void* cuda_stream) {
    if (!e || (n && !d_tuples) || !num_features) return DTE_ERR_ARG;
    if (!n) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(e->dev));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->s_main;
    const unsigned long long ne = n * num_features;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((ne + 255) / 256, (unsigned long long)e->sm_count * 32);
    synth_tuples_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(d_tuples), first_tuple * num_features, ne, seed,
                                                missing_ppm, missing_value);
    CUDA_TRY(e, cudaGetLastError());
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

}  // extern "C"
