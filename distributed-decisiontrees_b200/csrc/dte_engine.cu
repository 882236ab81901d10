// dte_engine.cu — host side of libdte.so: the C ABI of include/dte.h.  One handle = the host node's view of a
// ring of 1..20 devices: ONE soft-register file, ONE PCIe line stream in, ONE result stream out; behind it the
// per-device pipelines of dte_device.cuh.  No torch, no Python, no NCCL link dependency (libnccl is dlopen'ed
// only when DTE_OPT_COMBINE = 1): plain CUDA runtime.  There is NO CPU compute path in this file — every score
// comes from a CUDA kernel.
//
// Reference behaviour mirrored here (paths relative to the reference root):
//   CSR decode ............... rtl/DTEngine/EngineCSR.sv:113-125,189-306
//   stream order / dealing ... rtl/DTEngine/PCIeReceiver.sv:136-139,205-316
//   tree / tuple framing ..... rtl/DTEngine/InputDistributor.sv:248-296, data broadcast :199-204
//   result packing / combine . rtl/DTEngine/ResultsCombiner.sv:132-162,292-311,359-391
//   completion / counters .... rtl/DTEngine/DTInference.sv:314-374,633-663
#include "../../include/dte.h"
#include "dte_device.cuh"
#include "dte_partition.hpp"

#include <atomic>
#include <chrono>
#include <dlfcn.h>
#include <mutex>
#include <new>
#include <thread>

using namespace dte;

namespace {
using Clock = std::chrono::steady_clock;
inline double ns_since(Clock::time_point t0, Clock::time_point t1) {
    return std::chrono::duration<double, std::nano>(t1 - t0).count();
}

// ---- NCCL through dlopen (only for DTE_OPT_COMBINE = 1; the default combine is our own ring-order kernel) ----
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // a copy the process already has (e.g. torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW);
        if (!h) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        Reduce = (decltype(Reduce))dlsym(h, "ncclReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Reduce) return false;
        lib = h;
        return true;
    }
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;        // ncclDataType_t / ncclRedOp_t values of nccl.h
}  // namespace

struct dte_engine {
    std::vector<Dev> devs;            // devs[d] = device ID d of devices_list = d-th ordinal given at create
    int pos2dev[kMaxRing];            // ring position -> index into devs (registers 208-210), decoded at start
    uint32_t node_index = 0;          // single-device handle: the ring position this engine stands for (dte_set_node)
    Tune tune;
    int forced_variant = KERNEL_AUTO;
    std::string err;

    // ---- CSR file (EngineCSR.sv) ----
    uint64_t regs[12] = {0};          // 200..211 as written
    Geom g;                           // geometry of the current run (committed by `start` / load on success only)

    // ---- options ----
    uint32_t cycle_mhz = 0;
    int combine_nccl = 0;
    size_t opt_queue_lines = 0, opt_chunk_tuples = 0;

    // ---- stream state (PCIeReceiver.sv FSM) ----
    enum { ST_IDLE = 0, ST_TREES = 1, ST_WAIT = 2, ST_DATA = 3 } state = ST_IDLE;
    uint64_t lines_received = 0;
    uint64_t cur_w = 0, cur_f = 0, cur_d = 0;   // currWCount / currFCount / currDCount
    uint32_t cur_dev = 0;                       // currDevID
    std::vector<std::vector<unsigned char>> tree_w, tree_f;   // tree lines kept per ring position (weights | indexes)

    // ---- partition of the current run (decided at `start` / load) ----
    uint32_t ndev_ring = 1;           // numDevs (reg 203[39:32])
    int prog_partition = -1;          // how the RESIDENT ensemble was placed: 0 one device, 1 replicated, 2 chunks in ring order
    int prog_pos2dev[kMaxRing];       // ... and on which devices (a run must use the same placement)
    bool group = false;               // ensemble-sharded lock-step group: data broadcast, partial scores ring-combined
    bool deal = false;                // data lines dealt round-robin in batches of deal_lines
    uint64_t deal_lines = 0;

    // ---- data / result accounting since `start` ----
    uint64_t data_lines_in = 0;       // data lines received
    uint64_t res_read = 0;            // results handed out (tuples, in the order this handle exposes them)
    uint64_t result_lines_out = 0;    // result lines handed out (process_done)
    uint32_t out_cl_count = 0;        // pcie_out_cl_count (DTInference.sv:632-646)
    uint64_t ring_cap = 0;            // result queue capacity in tuples
    uint64_t prog_lines_local = 0, sl3_lines_sent = 0, sl3_res_lines = 0;
    bool done_latched = false;

    // ---- cumulative counters (cluster counters are only cleared by a hardware reset, DTPUCluster.sv:85-97) ----
    uint64_t cum_c0 = 0, cum_c0_lines = 0;
    uint64_t tuples_in = 0;
    std::atomic<uint64_t> tuples_out{0};
    std::mutex err_mu;                // fail() may be reached from the per-device feeder threads of infer_host

    // ---- timing: progCycles / execCycles (DTInference.sv:330-357) ----
    Clock::time_point t_start, t_prog0, t_prog1, t_done;
    bool running = false, prog_seen = false, prog_open = false;
    double last_walk_ms = 0;

    // ---- NCCL (optional combine) ----
    NcclApi nccl;
    std::vector<void*> nccl_comms;

    bool multi() const { return devs.size() > 1; }
};

namespace {

int fail(dte_engine* e, int code, const char* fmt, ...) {
    if (e) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        std::lock_guard<std::mutex> lk(e->err_mu);
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
#define TRY(call)                  \
    do {                           \
        int _rc = (call);          \
        if (_rc) return _rc;       \
    } while (0)

// Decode the geometry registers (EngineCSR.sv:218-233) into `g` and check them; the engine is not touched.
int decode_geom(dte_engine* e, Geom& g) {
    const uint64_t r204 = e->regs[4], r205 = e->regs[5];
    g.w_cls = (uint32_t)((r204 >> 16) & 0xFFFF);
    g.f_cls = (uint32_t)((r204 >> 32) & 0xFFFF);
    g.tuple_cls = (uint32_t)((r204 >> 48) & 0xFFFF);
    g.missing = (uint32_t)(r205 & 0xFFFFFFFFu);
    g.D = (uint32_t)((r205 >> 32) & 0xF);
    g.S = (uint32_t)((r205 >> 36) & 0xFF);
    g.K = (uint32_t)((r205 >> 44) & 0xF);
    if (g.D < 1) return fail(e, DTE_ERR_CONFIG, "reg 205: num_levels_per_tree = 0");
    if (g.K < 1 || g.K > 8) return fail(e, DTE_ERR_CONFIG, "reg 205: num_clusters_per_tuple = %u, need 1..8", g.K);
    if (g.S < 1) return fail(e, DTE_ERR_CONFIG, "reg 205: num_trees_per_pu = 0");
    if (g.tuple_cls < 1 || g.tuple_cls > 512) return fail(e, DTE_ERR_CONFIG, "reg 204: tuple_numcls = %u, need 1..512", g.tuple_cls);
    if ((uint64_t)g.w_cls * 4 < (2ull << g.D) - 1)
        return fail(e, DTE_ERR_CONFIG, "reg 204: tree_weights_numcls = %u too small for %u levels", g.w_cls, g.D);
    if ((uint64_t)g.f_cls * 8 < (1ull << g.D) - 1)
        return fail(e, DTE_ERR_CONFIG, "reg 204: tree_feature_index_numcls = %u too small for %u levels", g.f_cls, g.D);
    return DTE_OK;
}

// A run may reuse the resident ensemble only with the geometry it was loaded with (F and D fix the device
// layout and the validated feature-index range; K, S and the missing pattern are free to change).
int check_resident(dte_engine* e, const Geom& g) {
    if (e->multi() && e->prog_partition >= 0) {
        const int now = e->group ? 2 : 1;
        if (now != e->prog_partition)
            return fail(e, DTE_ERR_CONFIG, "reg 201 selects the %s partition but the resident ensemble was programmed %s",
                        now == 2 ? "ensemble-sharded" : "data-sharded", e->prog_partition == 2 ? "in per-device chunks" : "replicated");
        if (now == 2 && memcmp(e->pos2dev, e->prog_pos2dev, sizeof(int) * e->devs.size()))
            return fail(e, DTE_ERR_CONFIG, "devices_list (registers 208-210) changed since the tree chunks were programmed: the ring order of the partial sums would change");
    }
    for (Dev& d : e->devs) {
        if (!d.d_top) return fail(e, DTE_ERR_STATE, "no ensemble loaded");
        if (d.g.tuple_cls != g.tuple_cls || d.g.D != g.D)
            return fail(e, DTE_ERR_CONFIG, "registers describe %u features / %u levels but the resident ensemble was loaded with %u / %u",
                        g.F(), g.D, d.g.F(), d.g.D);
    }
    return DTE_OK;
}

void free_ensemble(Dev& d) {
    cudaSetDevice(d.ordinal);
    if (d.d_top) cudaFree(d.d_top);
    if (d.d_bottom) cudaFree(d.d_bottom);
    d.d_top = nullptr;
    d.d_bottom = nullptr;
    d.T = d.Tpad = 0;
    d.ensemble_bytes = 0;
}

int upload_ensemble(dte_engine* e, Dev& d, const PackedEnsemble& pk) {
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    CUDA_TRY(e, cudaDeviceSynchronize());           // nothing may still be walking the old ensemble
    free_ensemble(d);
    CUDA_TRY(e, cudaMalloc(&d.d_top, pk.top.size() * sizeof(uint2)));
    CUDA_TRY(e, cudaMalloc(&d.d_bottom, pk.bottom.size() * sizeof(uint4)));
    CUDA_TRY(e, cudaMemcpy(d.d_top, pk.top.data(), pk.top.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(d.d_bottom, pk.bottom.data(), pk.bottom.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    d.g = pk.g;
    d.T = pk.T; d.Tpad = pk.Tpad; d.Dtop = pk.Dtop; d.top_stride = pk.top_stride; d.nb = pk.nb; d.wide = pk.wide;
    d.ensemble_bytes = pk.top.size() * sizeof(uint2) + pk.bottom.size() * sizeof(uint4);
    return DTE_OK;
}

int pack_or_fail(dte_engine* e, const Geom& g, const unsigned char* wl, size_t n_wl, const unsigned char* fl, size_t n_fl,
                 uint32_t first, uint32_t count, PackedEnsemble& pk) {
    std::string msg;
    const int rc = pack_ensemble(g, wl, n_wl, fl, n_fl, first, count, pk, msg);
    if (rc) return fail(e, rc == -4 ? DTE_ERR_UNSUPPORTED : DTE_ERR_ARG, "%s", msg.c_str());
    return DTE_OK;
}

// ---- partition of a run: registers 201 / 203 / 208-210 ---------------------------------------------
struct Flags {
    bool data_distributed, host_node, bcast_data, bcast_trees, aggreg, multiple, rx_enabled;
    uint64_t batch_cls, chunk_w, chunk_f;
    uint32_t ndev;
};
Flags decode_flags(const dte_engine* e) {
    const uint64_t r201 = e->regs[1], r203 = e->regs[3];
    Flags f;
    f.data_distributed = r201 & 1; f.host_node = (r201 >> 1) & 1; f.bcast_data = (r201 >> 2) & 1;
    f.bcast_trees = (r201 >> 3) & 1; f.aggreg = (r201 >> 4) & 1; f.multiple = (r201 >> 5) & 1;
    f.rx_enabled = (r201 >> 6) & 1;
    f.batch_cls = r201 >> 32;
    f.chunk_w = (r203 & 0xFFFF) + 1;                // numcls_local_weights, "subtract in SW" (EngineCSR.sv:213)
    f.chunk_f = (r203 >> 16) & 0xFFFF;
    f.ndev = (uint32_t)std::max<uint64_t>(1, (r203 >> 32) & 0xFF);
    return f;
}

// devices_list (registers 208-210): entry i, 5 bits in byte lane i%8 of register 208 + i/8 (EngineCSR.sv:250-296)
uint32_t devices_list_entry(const dte_engine* e, uint32_t i) { return (uint32_t)((e->regs[8 + i / 8] >> (8 * (i % 8))) & 0x1F); }

// Which device of this handle serves ring position `pos` (-1: none — another process owns it).
int owner_of(const dte_engine* e, uint32_t pos) {
    if (e->multi()) return pos < e->devs.size() ? e->pos2dev[pos] : -1;
    return pos == e->node_index ? 0 : -1;
}

// Decide how the run is partitioned.  Single-device handles keep the raw RTL filter semantics (one process per
// GPU replaying the same stream, dte_set_node); a multi-device handle supports exactly the two partitions of
// SURVEY 8(e) and refuses the rest.
int decide_partition(dte_engine* e, const Geom& g) {
    const Flags f = decode_flags(e);
    bool group = false, deal = false;
    int pos2dev[kMaxRing];
    const size_t G = e->devs.size();
    for (uint32_t i = 0; i < (uint32_t)kMaxRing; ++i) pos2dev[i] = (int)i;
    const bool split = f.multiple && f.ndev > 1;
    if (e->multi()) {
        if (!split) return fail(e, DTE_ERR_CONFIG, "a %zu-device handle needs multiple_nodes=1 (reg 201[5]) and numDevs=%zu (reg 203[39:32])", G, G);
        if (f.ndev != G) return fail(e, DTE_ERR_CONFIG, "reg 203: numDevs=%u but the handle drives %zu devices", f.ndev, G);
        bool any = false;
        for (uint32_t i = 0; i < G; ++i) any |= devices_list_entry(e, i) != 0;
        if (any) {                                   // programmed: must be a permutation of the device IDs 0..G-1
            std::vector<int> seen(G, 0);
            for (uint32_t i = 0; i < G; ++i) {
                const uint32_t id = devices_list_entry(e, i);
                if (id >= G || seen[id]++) return fail(e, DTE_ERR_CONFIG, "registers 208-210: devices_list[%u]=%u is not a permutation of 0..%zu", i, id, G - 1);
                pos2dev[i] = (int)id;
            }
        }
        if (f.bcast_trees && !f.bcast_data && !f.aggreg) {
            deal = true;
        } else if (!f.bcast_trees && f.bcast_data && f.aggreg) {
            group = true;
            for (size_t a = 0; a < G; ++a)
                for (size_t b = 0; b < G; ++b) {
                    int ok = 1;
                    if (e->devs[a].ordinal != e->devs[b].ordinal) cudaDeviceCanAccessPeer(&ok, e->devs[a].ordinal, e->devs[b].ordinal);
                    if (!ok) return fail(e, DTE_ERR_UNSUPPORTED, "GPU %d cannot access GPU %d peer-to-peer: the data broadcast and the ring combine need it",
                                         e->devs[a].ordinal, e->devs[b].ordinal);
                }
        } else {
            return fail(e, DTE_ERR_CONFIG, "reg 201 = 0x%llx: a multi-device handle supports broadcast_trees (data dealt, no aggregate) or "
                        "broadcast_data + aggreg_enabled (tree chunks, ring combine)", (unsigned long long)(e->regs[1] & 0xFF));
        }
    } else if (split && !f.bcast_data && !f.data_distributed) {
        deal = true;                                 // this process keeps the batches of its own ring position
    }
    if (deal && (f.batch_cls == 0 || f.batch_cls % g.tuple_cls))
        return fail(e, DTE_ERR_CONFIG, "reg 201: core_data_batch_cls=%llu must be a positive multiple of tuple_numcls=%u",
                    (unsigned long long)f.batch_cls, g.tuple_cls);
    e->ndev_ring = f.ndev;
    e->group = group;
    e->deal = deal;
    e->deal_lines = deal ? f.batch_cls : 0;
    memcpy(e->pos2dev, pos2dev, sizeof pos2dev);
    return DTE_OK;
}

// ---- landing slots ------------------------------------------------------------------------------
void free_slots(Dev& d) {
    cudaSetDevice(d.ordinal);
    for (Slot& s : d.slot) {
        if (s.d_tup) cudaFree(s.d_tup);
        if (s.d_sc) cudaFree(s.d_sc);
        if (s.d_part) cudaFree(s.d_part);
        if (s.d_lb) cudaFree(s.d_lb);
        s.d_tup = nullptr; s.d_sc = nullptr; s.d_part = nullptr; s.d_lb = nullptr;
        s.fill = s.walked = 0;
    }
    d.cap_tuples = 0;
    d.slot_parts = false;
}

int ensure_slots(dte_engine* e, Dev& d, size_t cap, uint32_t F, bool parts) {
    if (d.cap_tuples == cap && d.slot_F == F && (d.slot_parts || !parts)) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    CUDA_TRY(e, cudaDeviceSynchronize());
    free_slots(d);
    for (Slot& s : d.slot) {
        CUDA_TRY(e, cudaMalloc(&s.d_tup, cap * F * 4));
        CUDA_TRY(e, cudaMalloc(&s.d_sc, cap * 4));
        CUDA_TRY(e, cudaMalloc(&s.d_lb, cap));
        if (parts) CUDA_TRY(e, cudaMalloc(&s.d_part, cap * 4));
    }
    d.cap_tuples = cap;
    d.slot_F = F;
    d.slot_parts = parts;
    d.cur = 0;
    return DTE_OK;
}

size_t default_chunk(const dte_engine* e, uint32_t F) {
    // 64 MiB of tuples per landing buffer; a multi-device handle issues ~25 API calls per device per buffer from ONE
    // host thread, so it takes up to 4 x larger buffers to stay walk-bound instead of launch-bound (measured, 8 GPUs)
    const size_t bytes = (64ull << 20) * std::min<size_t>(4, e->devs.size());
    size_t c = e->opt_chunk_tuples ? e->opt_chunk_tuples : (e->tune.chunk ? e->tune.chunk : std::max<size_t>(4096, bytes / (F * 4)));
    return (std::max<size_t>(c, 4) + 3) & ~(size_t)3;      // whole result lines per slot
}

int ensure_ring(dte_engine* e, Dev& d, size_t cap_tuples) {
    if (d.ring_cap >= cap_tuples) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    if (d.h_ring) { cudaFreeHost(d.h_ring); d.h_ring = nullptr; d.ring_cap = 0; }
    CUDA_TRY(e, cudaHostAlloc(reinterpret_cast<void**>(&d.h_ring), cap_tuples * 4, cudaHostAllocPortable));
    d.ring_cap = cap_tuples;
    return DTE_OK;
}

cudaEvent_t get_event(Dev& d) {
    if (!d.ev_pool.empty()) {
        cudaEvent_t ev = d.ev_pool.back();
        d.ev_pool.pop_back();
        return ev;
    }
    cudaEvent_t ev = nullptr;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    return ev;
}

// Retire completed result copies of device d; block until its local result count reaches `need` (0 = only poll).
int poll_dev(dte_engine* e, Dev& d, uint64_t need) {
    if (!d.pend.empty()) cudaSetDevice(d.ordinal);
    while (!d.pend.empty()) {
        Pending& p = d.pend.front();
        cudaError_t st = (d.res_done < need) ? cudaEventSynchronize(p.ev) : cudaEventQuery(p.ev);
        if (st == cudaErrorNotReady) break;
        if (st != cudaSuccess) return fail(e, DTE_ERR_CUDA, "result copy failed: %s", cudaGetErrorString(st));
        d.res_done = p.upto;
        d.ev_pool.push_back(p.ev);
        d.pend.pop_front();
    }
    return DTE_OK;
}

// D2H of `cnt` results of slot s starting at slot-local tuple `from`, into the device's current sink.
int enqueue_results(dte_engine* e, Dev& d, Slot& s, size_t from, size_t cnt, bool labels) {
    CUDA_TRY(e, cudaStreamWaitEvent(d.s_d2h, s.ev_walk, 0));
    if (d.sink_sc) {
        CUDA_TRY(e, cudaMemcpyAsync(d.sink_sc, s.d_sc + from, cnt * 4, cudaMemcpyDeviceToHost, d.s_d2h));
        d.sink_sc += cnt;
        if (labels && d.sink_lb) {
            CUDA_TRY(e, cudaMemcpyAsync(d.sink_lb, s.d_lb + from, cnt, cudaMemcpyDeviceToHost, d.s_d2h));
            d.sink_lb += cnt;
        }
    } else {
        size_t pos = (size_t)(d.res_enq % d.ring_cap), left = cnt, off = from;
        while (left) {                               // the ring may wrap
            const size_t run = std::min(left, d.ring_cap - pos);
            CUDA_TRY(e, cudaMemcpyAsync(d.h_ring + pos, s.d_sc + off, run * 4, cudaMemcpyDeviceToHost, d.s_d2h));
            left -= run; off += run; pos = 0;
        }
    }
    CUDA_TRY(e, cudaEventRecord(s.ev_d2h, d.s_d2h));
    cudaEvent_t ev = get_event(d);
    CUDA_TRY(e, cudaEventRecord(ev, d.s_d2h));
    d.res_enq += cnt;
    d.pend.push_back({ev, d.res_enq});
    return DTE_OK;
}

// Walk the whole tuples that have landed in slot b of device d since the last submit and queue their results.
int submit_dev(dte_engine* e, Dev& d, int b) {
    Slot& s = d.slot[b];
    const size_t tb = d.g.tuple_bytes();
    if (!d.cap_tuples || !tb) return DTE_OK;
    const size_t whole = s.fill / tb, cnt = whole - s.walked;
    if (!cnt) return DTE_OK;
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    CUDA_TRY(e, cudaEventRecord(s.ev_land, d.s_h2d));
    CUDA_TRY(e, cudaStreamWaitEvent(d.s_main, s.ev_land, 0));
    CUDA_TRY(e, cudaStreamWaitEvent(d.s_main, s.ev_d2h, 0));        // score buffer of this slot drained
    const bool labels = d.sink_sc && d.sink_lb;
    const char* why = nullptr;
    cudaError_t rc = launch_walk(d, e->tune, e->forced_variant, s.d_tup + s.walked * tb, cnt, s.d_sc + s.walked,
                                 labels ? s.d_lb + s.walked : nullptr, d.s_main, false, &why);
    if (rc != cudaSuccess) return fail(e, why ? DTE_ERR_ARG : DTE_ERR_CUDA, "walk kernel launch failed: %s", why ? why : cudaGetErrorString(rc));
    CUDA_TRY(e, cudaEventRecord(s.ev_walk, d.s_main));
    TRY(enqueue_results(e, d, s, s.walked, cnt, labels));
    s.walked = whole;
    e->tuples_out += cnt;
    return DTE_OK;
}

// ---- host -> device copy of one piece of the stream ------------------------------------------------------
constexpr size_t kStageBytes = 16u << 20;
constexpr int kCopyThreads = 4;

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

void parallel_memcpy(unsigned char* dst, const unsigned char* src, size_t bytes) {
    if (bytes < (1u << 20)) { memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes + kCopyThreads - 1) / kCopyThreads + 63) & ~(size_t)63;
    std::thread th[kCopyThreads - 1];
    int nth = 0;
    for (int k = 1; k < kCopyThreads; ++k) {
        const size_t lo = (size_t)k * per;
        if (lo >= bytes) break;
        th[nth++] = std::thread([=] { memcpy(dst + lo, src + lo, std::min(per, bytes - lo)); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (int k = 0; k < nth; ++k) th[k].join();
}

// Pinned (or registered) memory is DMA'd in place.  Pageable memory is copied by kCopyThreads threads into a pinned
// staging ring and DMA'd from there, the copy of piece i+1 overlapping the DMA of piece i.
int h2d_piece(dte_engine* e, Dev& d, unsigned char* dst, const unsigned char* src, size_t bytes, bool pageable) {
    if (!pageable || bytes < (256u << 10)) {
        CUDA_TRY(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d.s_h2d));
        return DTE_OK;
    }
    if (!d.h_stage[0]) {
        for (int b = 0; b < 3; ++b) {
            CUDA_TRY(e, cudaHostAlloc(reinterpret_cast<void**>(&d.h_stage[b]), kStageBytes, cudaHostAllocDefault));
            CUDA_TRY(e, cudaEventCreateWithFlags(&d.ev_stage[b], cudaEventDisableTiming));
        }
    }
    for (size_t off = 0; off < bytes; off += kStageBytes) {
        const size_t len = std::min(kStageBytes, bytes - off);
        const int b = d.stage_next;
        d.stage_next = (b + 1) % 3;
        CUDA_TRY(e, cudaEventSynchronize(d.ev_stage[b]));             // the DMA that last read this staging buffer is done
        parallel_memcpy(d.h_stage[b], src + off, len);
        CUDA_TRY(e, cudaMemcpyAsync(dst + off, d.h_stage[b], len, cudaMemcpyHostToDevice, d.s_h2d));
        CUDA_TRY(e, cudaEventRecord(d.ev_stage[b], d.s_h2d));
    }
    return DTE_OK;
}

void advance_slot(Dev& d) {
    d.cur = (d.cur + 1) % kNumSlots;
    d.slot[d.cur].fill = 0;
    d.slot[d.cur].walked = 0;
}

// Land `bytes` of the tuple stream on device d (pinned memory is DMA'd in place, pageable memory goes through the
// engine's staging ring, h2d_piece).  Full slots are submitted as they fill.
int land_dev(dte_engine* e, Dev& d, const unsigned char* src, size_t bytes) {
    const size_t cap_bytes = d.cap_tuples * d.g.tuple_bytes();
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    const bool pageable = bytes >= (256u << 10) && is_pageable(src);
    while (bytes) {
        Slot& s = d.slot[d.cur];
        if (s.fill == 0) CUDA_TRY(e, cudaStreamWaitEvent(d.s_h2d, s.ev_walk, 0));   // previous walk done with this buffer
        const size_t take = std::min(bytes, cap_bytes - s.fill);
        TRY(h2d_piece(e, d, s.d_tup + s.fill, src, take, pageable));
        s.fill += take;
        src += take;
        bytes -= take;
        if (s.fill == cap_bytes) {
            TRY(submit_dev(e, d, d.cur));
            advance_slot(d);
        }
    }
    return DTE_OK;
}

// ---- ensemble-sharded group: every device sees every tuple, partial scores are ring-combined ---------
Dev& host_dev(dte_engine* e) { return e->devs[(size_t)e->pos2dev[0]]; }
const Dev& host_dev(const dte_engine* e) { return e->devs[(size_t)e->pos2dev[0]]; }

// one launch of the ring-order combine; vector path only when every pointer is line-aligned
cudaError_t launch_ring_combine(const RingParts& parts, int G, float* out, uint8_t* labels, size_t n, int sm_count, cudaStream_t st) {
    bool vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && (!labels || (reinterpret_cast<uintptr_t>(labels) & 3u) == 0);
    for (int g = 0; g < G; ++g) vec = vec && (reinterpret_cast<uintptr_t>(parts.p[g]) & 15u) == 0;
    const size_t work = vec ? n / 4 + 3 : n;
    const unsigned blocks = (unsigned)std::min<size_t>((work + 255) / 256, (size_t)sm_count * 8);
    ring_combine_kernel<<<std::max(1u, blocks), 256, 0, st>>>(parts, G, out, labels, n, vec ? 1 : 0);
    return cudaGetLastError();
}

int nccl_setup(dte_engine* e) {
    if (!e->nccl_comms.empty()) return DTE_OK;
    if (!e->nccl.load()) return fail(e, DTE_ERR_UNSUPPORTED, "DTE_OPT_COMBINE=1 needs libnccl.so.2 (dlopen failed)");
    const int G = (int)e->devs.size();
    std::vector<int> ords((size_t)G);
    for (int r = 0; r < G; ++r) ords[(size_t)r] = e->devs[(size_t)e->pos2dev[r]].ordinal;       // NCCL rank = ring position
    e->nccl_comms.assign((size_t)G, nullptr);
    const int rc = e->nccl.CommInitAll(e->nccl_comms.data(), G, ords.data());
    if (rc) {
        e->nccl_comms.clear();
        return fail(e, DTE_ERR_CUDA, "ncclCommInitAll failed: %s", e->nccl.GetErrorString ? e->nccl.GetErrorString(rc) : "?");
    }
    return DTE_OK;
}

int submit_group(dte_engine* e, int b) {
    const size_t G = e->devs.size();
    Dev& h = host_dev(e);
    const size_t tb = h.g.tuple_bytes();
    if (!h.cap_tuples || !tb) return DTE_OK;
    const size_t whole = h.slot[b].fill / tb, walked = h.slot[b].walked, cnt = whole - walked;
    if (!cnt) return DTE_OK;
    for (Dev& d : e->devs) {
        CUDA_TRY(e, cudaSetDevice(d.ordinal));
        CUDA_TRY(e, cudaEventRecord(d.slot[b].ev_land, d.s_h2d));
    }
    for (Dev& d : e->devs) {
        Slot& s = d.slot[b];
        CUDA_TRY(e, cudaSetDevice(d.ordinal));
        for (Dev& src : e->devs) CUDA_TRY(e, cudaStreamWaitEvent(d.s_main, src.slot[b].ev_land, 0));   // every piece has landed here
        CUDA_TRY(e, cudaStreamWaitEvent(d.s_main, h.slot[b].ev_comb, 0));    // the previous combine is done with our partials
        const char* why = nullptr;
        cudaError_t rc = launch_walk(d, e->tune, e->forced_variant, s.d_tup + walked * tb, cnt, s.d_part + walked, nullptr, d.s_main, false, &why);
        if (rc != cudaSuccess) return fail(e, why ? DTE_ERR_ARG : DTE_ERR_CUDA, "walk kernel launch failed: %s", why ? why : cudaGetErrorString(rc));
        CUDA_TRY(e, cudaEventRecord(s.ev_walk, d.s_main));
        s.walked = whole;
    }
    Slot& hs = h.slot[b];
    const bool labels = h.sink_sc && h.sink_lb;
    CUDA_TRY(e, cudaSetDevice(h.ordinal));
    CUDA_TRY(e, cudaStreamWaitEvent(h.s_main, hs.ev_d2h, 0));
    if (!e->combine_nccl) {
        for (Dev& d : e->devs) CUDA_TRY(e, cudaStreamWaitEvent(h.s_main, d.slot[b].ev_walk, 0));
        RingParts parts;
        for (size_t r = 0; r < G; ++r) parts.p[r] = e->devs[(size_t)e->pos2dev[r]].slot[b].d_part + walked;   // ring order, host first
        CUDA_TRY(e, launch_ring_combine(parts, (int)G, hs.d_sc + walked, labels ? hs.d_lb + walked : nullptr, cnt, h.sm_count, h.s_main));
        h.kernel_launches++;
    } else {
        TRY(nccl_setup(e));
        int rc = e->nccl.GroupStart();
        for (size_t r = 0; r < G && !rc; ++r) {
            Dev& d = e->devs[(size_t)e->pos2dev[r]];
            rc = e->nccl.Reduce(d.slot[b].d_part + walked, hs.d_sc + walked, cnt, kNcclFloat32, kNcclSum, 0, e->nccl_comms[r], d.s_main);
        }
        const int rc2 = e->nccl.GroupEnd();
        if (rc || rc2) return fail(e, DTE_ERR_CUDA, "ncclReduce failed: %s", e->nccl.GetErrorString ? e->nccl.GetErrorString(rc ? rc : rc2) : "?");
        CUDA_TRY(e, cudaSetDevice(h.ordinal));
        if (labels) {
            labels_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h.s_main>>>(hs.d_sc + walked, hs.d_lb + walked, cnt);
            CUDA_TRY(e, cudaGetLastError());
        }
        for (Dev& d : e->devs) {                     // the reduce read d_part on every device's stream
            CUDA_TRY(e, cudaSetDevice(d.ordinal));
            CUDA_TRY(e, cudaEventRecord(d.slot[b].ev_walk, d.s_main));
        }
        CUDA_TRY(e, cudaSetDevice(h.ordinal));
        for (Dev& d : e->devs) CUDA_TRY(e, cudaStreamWaitEvent(h.s_main, d.slot[b].ev_walk, 0));
    }
    CUDA_TRY(e, cudaEventRecord(hs.ev_comb, h.s_main));
    CUDA_TRY(e, cudaEventRecord(hs.ev_walk, h.s_main));           // results of this slot are final after the combine
    TRY(enqueue_results(e, h, hs, walked, cnt, labels));
    e->tuples_out += cnt;
    return DTE_OK;
}

// Broadcast landing (InputDistributor.sv:199-204 over NVLink instead of the SL3 ring): device g uploads the
// g-th piece of the block over ITS OWN PCIe link and pushes it to every peer with peer-to-peer copies.
int land_group(dte_engine* e, const unsigned char* src, size_t bytes) {
    const size_t G = e->devs.size();
    Dev& h = host_dev(e);
    const size_t cap_bytes = h.cap_tuples * h.g.tuple_bytes();
    static const size_t kMinPiece = 256u << 10;
    uint32_t rot = 0;
    const bool pageable = bytes >= kMinPiece && is_pageable(src);
    while (bytes) {
        const int b = h.cur;
        const size_t fill = h.slot[b].fill;
        if (fill == 0) {
            for (Dev& d : e->devs) {                 // every copy stream writes into every device's slot b
                CUDA_TRY(e, cudaSetDevice(d.ordinal));
                for (Dev& o : e->devs) CUDA_TRY(e, cudaStreamWaitEvent(d.s_h2d, o.slot[b].ev_walk, 0));
            }
        }
        const size_t take = std::min(bytes, cap_bytes - fill);
        const size_t npieces = std::max<size_t>(1, std::min(G, take / kMinPiece));
        size_t off = 0;
        for (size_t pc = 0; pc < npieces; ++pc) {
            const size_t len = (pc + 1 == npieces) ? take - off : ((take / npieces) & ~(size_t)15);
            Dev& d = e->devs[(pc + rot) % G];
            CUDA_TRY(e, cudaSetDevice(d.ordinal));
            unsigned char* mine = d.slot[b].d_tup + fill + off;
            TRY(h2d_piece(e, d, mine, src + off, len, pageable));
            for (Dev& o : e->devs) {
                if (&o == &d) continue;
                if (o.ordinal == d.ordinal)
                    CUDA_TRY(e, cudaMemcpyAsync(o.slot[b].d_tup + fill + off, mine, len, cudaMemcpyDeviceToDevice, d.s_h2d));
                else
                    CUDA_TRY(e, cudaMemcpyPeerAsync(o.slot[b].d_tup + fill + off, o.ordinal, mine, d.ordinal, len, d.s_h2d));
            }
            off += len;
        }
        if (npieces < G) rot = (rot + (uint32_t)npieces) % (uint32_t)G;
        for (Dev& d : e->devs) d.slot[b].fill = fill + take;
        src += take;
        bytes -= take;
        if (fill + take == cap_bytes) {
            TRY(submit_group(e, b));
            for (Dev& d : e->devs) advance_slot(d);
        }
    }
    return DTE_OK;
}

// ---- run-level helpers ----------------------------------------------------------------------------
int flush_all(dte_engine* e) {
    if (e->group) return submit_group(e, host_dev(e).cur);
    for (Dev& d : e->devs) TRY(submit_dev(e, d, d.cur));
    return DTE_OK;
}

// wait until the source buffers of every enqueued H2D may be reused by the caller
int sync_sources(dte_engine* e) {
    for (Dev& d : e->devs) {
        CUDA_TRY(e, cudaSetDevice(d.ordinal));
        CUDA_TRY(e, cudaStreamSynchronize(d.s_h2d));
    }
    return DTE_OK;
}

int drain_all(dte_engine* e) {
    TRY(flush_all(e));
    for (Dev& d : e->devs) {
        CUDA_TRY(e, cudaSetDevice(d.ordinal));
        CUDA_TRY(e, cudaStreamSynchronize(d.s_h2d));
        CUDA_TRY(e, cudaStreamSynchronize(d.s_main));
        CUDA_TRY(e, cudaStreamSynchronize(d.s_d2h));
        TRY(poll_dev(e, d, 0));
    }
    return DTE_OK;
}

bool fragment_pending(const dte_engine* e) {
    for (const Dev& d : e->devs)
        if (d.cap_tuples && d.g.tuple_cls && d.slot[d.cur].fill % d.g.tuple_bytes()) return true;
    return false;
}

void reset_pipeline(Dev& d) {
    for (Slot& s : d.slot) { s.fill = 0; s.walked = 0; }
    d.cur = 0;
    d.sink_sc = nullptr;
    d.sink_lb = nullptr;
    d.res_enq = d.res_done = 0;
    d.tuples_landed = 0;
}

// Results exposed by this handle, in order: tuple i lives on device dev at local index loc (dte_partition.hpp).
void locate_result(const dte_engine* e, uint64_t i, int& dev, uint64_t& loc) {
    if (e->multi() && e->deal) {
        const Deal deal{e->deal_lines / e->g.tuple_cls, e->devs.size()};
        uint64_t pos;
        deal.locate(i, pos, loc);
        dev = e->pos2dev[pos];
    } else {
        dev = e->group ? e->pos2dev[0] : 0;
        loc = i;
    }
}
// how many leading results (in exposed order) the per-device local counts stand for
uint64_t global_prefix(const dte_engine* e, bool completed) {
    auto cnt = [&](const Dev& d) { return completed ? d.res_done : d.res_enq; };
    if (e->multi() && e->deal) {
        const Deal deal{e->deal_lines / e->g.tuple_cls, e->devs.size()};
        uint64_t c[kMaxRing];
        for (size_t r = 0; r < e->devs.size(); ++r) c[r] = cnt(e->devs[(size_t)e->pos2dev[r]]);
        return deal.prefix(c);
    }
    return cnt(e->devs[(size_t)(e->group ? e->pos2dev[0] : 0)]);
}

int prepare_data_phase(dte_engine* e, bool need_ring) {
    const uint32_t F = e->g.F();
    const size_t cap = default_chunk(e, F);
    for (Dev& d : e->devs) {
        TRY(ensure_slots(e, d, cap, F, e->group));
        d.g.K = e->g.K; d.g.S = e->g.S; d.g.missing = e->g.missing;     // summation geometry of THIS run
    }
    if (need_ring) {
        // result queue: reg 207 says how many result lines the run produces; bounded either way
        uint64_t lines = e->opt_queue_lines ? e->opt_queue_lines : (e->regs[7] & 0xFFFFFFFFull) + 64;
        if (!e->opt_queue_lines) lines = std::max<uint64_t>(lines, 1u << 18);
        lines = std::min<uint64_t>(std::max<uint64_t>(lines, 16), 1u << 26);
        e->ring_cap = lines * 4;
        const uint64_t G = e->devs.size();
        const uint64_t bt = e->deal ? e->deal_lines / e->g.tuple_cls : 0;
        for (Dev& d : e->devs) {
            const bool has_results = !e->group || &d == &host_dev(e);
            if (!has_results) continue;
            const uint64_t per = (e->multi() && e->deal) ? e->ring_cap / G + 2 * bt + cap : e->ring_cap + cap;
            TRY(ensure_ring(e, d, per));
        }
    }
    return DTE_OK;
}

// Host buffers in, host buffers out (fast path, no line framing): the same landing pipeline with the caller's
// score buffer as the direct sink.  Multi-device: chunks round-robin over the devices (trees replicated) or every
// chunk broadcast to all of them (tree chunks + ring combine).
int infer_host(dte_engine* e, const unsigned char* h_tuples, size_t n, float* h_scores, uint8_t* h_labels) {
    Geom g;
    TRY(decode_geom(e, g));
    if (e->state == dte_engine::ST_TREES) return fail(e, DTE_ERR_STATE, "infer_host while the tree stream is being received");
    if (fragment_pending(e)) return fail(e, DTE_ERR_STATE, "a partial tuple is pending in the line stream");
    TRY(drain_all(e));
    if (e->multi()) TRY(decide_partition(e, g));
    TRY(check_resident(e, g));
    if (n == 0) return DTE_OK;
    e->g = g;
    for (Dev& d : e->devs)                           // every chunk of this call starts on a fresh landing buffer
        if (d.cap_tuples && d.slot[d.cur].fill) advance_slot(d);
    TRY(prepare_data_phase(e, false));
    const size_t tb = g.tuple_bytes();
    std::vector<uint64_t> enq0(e->devs.size());
    for (size_t i = 0; i < e->devs.size(); ++i) {
        Dev& d = e->devs[i];
        enq0[i] = d.res_enq;
        CUDA_TRY(e, cudaSetDevice(d.ordinal));
        CUDA_TRY(e, cudaEventRecord(d.ev_t0, d.s_main));
    }
    int rc = DTE_OK;
    if (e->group) {
        Dev& h = host_dev(e);
        h.sink_sc = h_scores;
        h.sink_lb = h_labels;
        rc = land_group(e, h_tuples, n * tb);
        if (!rc) rc = submit_group(e, h.cur);
        for (Dev& d : e->devs)
            if (d.slot[d.cur].fill) advance_slot(d);
    } else {
        // chunks round-robin over the devices; with several devices each one is fed by its own host thread (one
        // thread cannot issue the ~25 runtime calls per chunk fast enough for 8 PCIe links)
        const size_t chunk = e->devs[0].cap_tuples, G = e->devs.size();
        auto feed = [&](size_t g) -> int {
            Dev& d = e->devs[g];
            int r = DTE_OK;
            size_t i = g;
            for (size_t off = g * chunk; off < n && !r; off += G * chunk, i += G) {
                const size_t cnt = std::min(chunk, n - off);
                d.sink_sc = h_scores + off;
                d.sink_lb = h_labels ? h_labels + off : nullptr;
                r = land_dev(e, d, h_tuples + off * tb, cnt * tb);
                if (!r && d.slot[d.cur].fill) {          // a short last chunk
                    r = submit_dev(e, d, d.cur);
                    advance_slot(d);
                }
            }
            return r;
        };
        if (G == 1) {
            rc = feed(0);
        } else {
            std::vector<int> rcs(G, DTE_OK);
            std::vector<std::thread> th;
            for (size_t g = 1; g < G; ++g) th.emplace_back([&, g] { rcs[g] = feed(g); });
            rcs[0] = feed(0);
            for (auto& t : th) t.join();
            for (int r : rcs) if (r && !rc) rc = r;
        }
    }
    for (Dev& d : e->devs) {
        cudaSetDevice(d.ordinal);
        cudaEventRecord(d.ev_t1, d.s_main);
    }
    const int rc2 = drain_all(e);
    float ms_max = 0;
    for (size_t i = 0; i < e->devs.size(); ++i) {
        Dev& d = e->devs[i];
        float ms = 0;
        cudaSetDevice(d.ordinal);
        if (cudaEventElapsedTime(&ms, d.ev_t0, d.ev_t1) == cudaSuccess) d.last_walk_ms = ms;
        ms_max = std::max(ms_max, ms);
        d.sink_sc = nullptr; d.sink_lb = nullptr;
        d.res_enq = d.res_done = enq0[i];            // direct-sink results are not part of the line-stream queue
    }
    e->last_walk_ms = ms_max;
    if (rc || rc2) return rc ? rc : rc2;
    e->tuples_in += n;
    return DTE_OK;
}

cudaStream_t pick_stream(Dev& d, void* cuda_stream) { return cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : d.s_main; }

int single_only(dte_engine* e, const char* what) {
    if (e->multi()) return fail(e, DTE_ERR_UNSUPPORTED, "%s takes device pointers of ONE device: use a single-device handle (a multi-device handle takes host buffers)", what);
    return DTE_OK;
}

int init_dev(Dev& d, int ordinal) {
    d.ordinal = ordinal;
    bool ok = cudaSetDevice(ordinal) == cudaSuccess &&
              cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, ordinal) == cudaSuccess &&
              cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ordinal) == cudaSuccess &&
              cudaStreamCreateWithFlags(&d.s_main, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&d.ev_t0) == cudaSuccess && cudaEventCreate(&d.ev_t1) == cudaSuccess;
    for (Slot& s : d.slot)
        ok = ok && cudaEventCreateWithFlags(&s.ev_walk, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev_land, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev_comb, cudaEventDisableTiming) == cudaSuccess;
    return ok ? DTE_OK : DTE_ERR_CUDA;
}

void destroy_dev(Dev& d) {
    if (cudaSetDevice(d.ordinal) != cudaSuccess) return;
    cudaDeviceSynchronize();
    free_ensemble(d);
    free_slots(d);
    if (d.h_ring) cudaFreeHost(d.h_ring);
    for (int b = 0; b < 3; ++b) {
        if (d.h_stage[b]) cudaFreeHost(d.h_stage[b]);
        if (d.ev_stage[b]) cudaEventDestroy(d.ev_stage[b]);
    }
    for (Slot& s : d.slot) {
        if (s.ev_walk) cudaEventDestroy(s.ev_walk);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
        if (s.ev_land) cudaEventDestroy(s.ev_land);
        if (s.ev_comb) cudaEventDestroy(s.ev_comb);
    }
    for (Pending& p : d.pend) cudaEventDestroy(p.ev);
    for (cudaEvent_t ev : d.ev_pool) cudaEventDestroy(ev);
    if (d.ev_t0) cudaEventDestroy(d.ev_t0);
    if (d.ev_t1) cudaEventDestroy(d.ev_t1);
    if (d.s_main) cudaStreamDestroy(d.s_main);
    if (d.s_h2d) cudaStreamDestroy(d.s_h2d);
    if (d.s_d2h) cudaStreamDestroy(d.s_d2h);
}

// Program the devices from the tree lines gathered per ring position.
//   replicated: every device gets all trees;  chunked: ring position r gets the lines kept for it.
int program_devices(dte_engine* e, const Geom& g) {
    const Flags f = decode_flags(e);
    const bool replicate = f.bcast_trees || f.ndev <= 1 || !f.multiple;
    PackedEnsemble pk;
    if (e->multi()) {
        for (size_t r = 0; r < e->devs.size(); ++r) {
            const size_t src = replicate ? 0 : r;
            if (!replicate || r == 0) {
                const size_t lw = e->tree_w[src].size() / 16, lf = e->tree_f[src].size() / 16;
                if (!lw || !lf) return fail(e, DTE_ERR_CONFIG, "ring position %zu received no trees", r);
                TRY(pack_or_fail(e, g, e->tree_w[src].data(), lw, e->tree_f[src].data(), lf, 0, 0, pk));
            }
            TRY(upload_ensemble(e, e->devs[(size_t)e->pos2dev[r]], pk));
        }
    } else {
        const size_t src = replicate ? 0 : e->node_index;
        const size_t lw = e->tree_w[src].size() / 16, lf = e->tree_f[src].size() / 16;
        if (!lw || !lf) return fail(e, DTE_ERR_CONFIG, "node %u received no trees", e->node_index);
        TRY(pack_or_fail(e, g, e->tree_w[src].data(), lw, e->tree_f[src].data(), lf, 0, 0, pk));
        TRY(upload_ensemble(e, e->devs[0], pk));
    }
    e->prog_partition = !e->multi() ? 0 : (replicate ? 1 : 2);
    memcpy(e->prog_pos2dev, e->pos2dev, sizeof e->prog_pos2dev);
    return DTE_OK;
}

uint64_t cycles_of(const dte_engine* e, double ns) {
    return e->cycle_mhz ? (uint64_t)(ns * 1e-3 * e->cycle_mhz) : (uint64_t)ns;
}

// tuples served by cluster 0: tuple i uses clusters (i*K)%8 .. +K-1 (Core.sv:305-316)
uint64_t cluster0_tuples(const dte_engine* e) {
    const Dev& h = e->devs[(size_t)(e->multi() ? e->pos2dev[0] : 0)];
    return (h.tuples_landed * std::max(1u, e->g.K) + 7) / 8;
}

// process_done = total_pcie_out_cls_count == total_results_numcls (DTInference.sv:649-656); "produced" counts the
// lines handed out plus the complete ones still queued
int check_done(dte_engine* e, bool* done) {
    const uint64_t want = e->regs[7] & 0xFFFFFFFFull;                           // reg 207[31:0]
    if (e->done_latched) { *done = true; return DTE_OK; }
    *done = false;
    if (!want || !e->running || e->state != dte_engine::ST_DATA) return DTE_OK;
    TRY(flush_all(e));
    for (Dev& d : e->devs) TRY(poll_dev(e, d, 0));
    const uint64_t produced = e->result_lines_out + (global_prefix(e, true) - e->res_read) / 4;
    if (produced >= want) {
        *done = true;
        e->done_latched = true;
        e->t_done = Clock::now();
        e->running = false;
        e->state = dte_engine::ST_IDLE;                                         // PCIeReceiver.sv:289-292
        const uint64_t c0 = cluster0_tuples(e);
        e->cum_c0 += c0;
        e->cum_c0_lines += c0 * e->g.tuple_cls;
        for (Dev& d : e->devs) d.tuples_landed = 0;
    }
    return DTE_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* dte_version(void) { return "dte-b200 0.2 (sm_100a)"; }

int dte_create_multi(dte_t** engine, const int* gpu_ordinals, int n_gpus) {
    if (!engine) return DTE_ERR_ARG;
    *engine = nullptr;
    if (!gpu_ordinals || n_gpus < 1 || n_gpus > kMaxRing) return DTE_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DTE_ERR_CUDA;   // no CPU fallback
    // an ordinal may appear more than once: several ring positions then share one GPU (how the multi-device
    // logic is tested on a one-GPU box; the NCCL combine needs distinct GPUs)
    for (int i = 0; i < n_gpus; ++i)
        if (gpu_ordinals[i] < 0 || gpu_ordinals[i] >= ndev) return DTE_ERR_ARG;
    dte_engine* e = new (std::nothrow) dte_engine();
    if (!e) return DTE_ERR_NOMEM;
    parse_tune(e->tune);
    e->devs.resize((size_t)n_gpus);
    for (uint32_t i = 0; i < (uint32_t)kMaxRing; ++i) e->pos2dev[i] = (int)i;
    e->tree_w.resize(kMaxRing);
    e->tree_f.resize(kMaxRing);
    int rc = DTE_OK;
    for (int i = 0; i < n_gpus && !rc; ++i) rc = init_dev(e->devs[(size_t)i], gpu_ordinals[i]);
    if (!rc && n_gpus > 1) {
        // peer access between all devices of the ring: the data broadcast and the ring combine ride NVLink
        for (int a = 0; a < n_gpus; ++a)
            for (int b = 0; b < n_gpus; ++b) {
                if (gpu_ordinals[a] == gpu_ordinals[b]) continue;
                int ok = 0;
                cudaDeviceCanAccessPeer(&ok, gpu_ordinals[a], gpu_ordinals[b]);
                if (!ok) continue;                   // refused later, only if a run needs it
                cudaSetDevice(gpu_ordinals[a]);
                if (cudaDeviceEnablePeerAccess(gpu_ordinals[b], 0) != cudaSuccess) cudaGetLastError();   // "already enabled" is fine
            }
    }
    if (rc) {
        for (Dev& d : e->devs) destroy_dev(d);
        delete e;
        return rc;
    }
    *engine = e;
    return DTE_OK;
}

int dte_create(dte_t** engine, int gpu_ordinal) { return dte_create_multi(engine, &gpu_ordinal, 1); }

int dte_destroy(dte_t* e) {
    if (!e) return DTE_ERR_ARG;
    for (Dev& d : e->devs) {
        cudaSetDevice(d.ordinal);
        cudaDeviceSynchronize();
    }
    if (!e->nccl_comms.empty() && e->nccl.CommDestroy)
        for (void* c : e->nccl_comms)
            if (c) e->nccl.CommDestroy(c);
    for (Dev& d : e->devs) destroy_dev(d);
    delete e;
    return DTE_OK;
}

const char* dte_last_error(const dte_t* e) { return e ? e->err.c_str() : "null engine"; }

int dte_set_option(dte_t* e, int option, uint64_t value) {
    if (!e) return DTE_ERR_ARG;
    switch (option) {
        case DTE_OPT_CYCLE_MHZ: e->cycle_mhz = (uint32_t)value; break;
        case DTE_OPT_COMBINE:
            if (value > 1) return fail(e, DTE_ERR_ARG, "DTE_OPT_COMBINE: 0 (ring kernel) or 1 (NCCL reduce)");
            e->combine_nccl = (int)value;
            break;
        case DTE_OPT_RESULT_QUEUE_LINES: e->opt_queue_lines = (size_t)value; break;
        case DTE_OPT_CHUNK_TUPLES: e->opt_chunk_tuples = (size_t)value; break;
        default: return fail(e, DTE_ERR_ARG, "unknown option %d", option);
    }
    return DTE_OK;
}

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
    TRY(drain_all(e));                                    // results of an unfinished previous run are dropped
    if (e->running) {                                     // cluster counters survive `start` (DTPUCluster.sv:85-97)
        const uint64_t c0 = cluster0_tuples(e);
        e->cum_c0 += c0;
        e->cum_c0_lines += c0 * e->g.tuple_cls;
    }
    e->lines_received = 0;
    e->cur_w = e->cur_f = e->cur_d = 0;
    e->cur_dev = 0;
    for (auto& v : e->tree_w) v.clear();
    for (auto& v : e->tree_f) v.clear();
    e->data_lines_in = 0;
    e->res_read = 0;
    e->result_lines_out = 0;
    e->out_cl_count = 0;
    e->prog_lines_local = e->sl3_lines_sent = e->sl3_res_lines = 0;
    e->done_latched = false;
    e->prog_seen = e->prog_open = false;
    e->ring_cap = 0;
    for (Dev& d : e->devs) reset_pipeline(d);
    e->t_start = Clock::now();
    e->running = true;
    const Flags f = decode_flags(e);
    e->state = dte_engine::ST_IDLE;
    if (f.rx_enabled) {                                   // PCIeReceiver.sv:218-227
        if (f.host_node) e->state = dte_engine::ST_TREES;
        else if (f.data_distributed) e->state = dte_engine::ST_DATA;
    }
    if (e->state == dte_engine::ST_IDLE) { e->running = false; return DTE_OK; }
    // decode into locals; commit only when the whole configuration is accepted
    Geom g;
    int rc = decode_geom(e, g);
    if (!rc) rc = decide_partition(e, g);
    if (!rc && e->state == dte_engine::ST_DATA) {
        rc = check_resident(e, g);
        if (!rc) { e->g = g; rc = prepare_data_phase(e, true); }
    }
    if (rc) { e->state = dte_engine::ST_IDLE; e->running = false; return rc; }
    e->g = g;
    return DTE_OK;
}

int dte_softreg_read(dte_t* e, uint32_t addr, uint64_t* data) {
    if (!e || !data) return DTE_ERR_ARG;
    const Clock::time_point now = Clock::now();
    // appStatus (DTInference.sv:367-372): pairs of 32-bit debug counters, modelled from the run's counts
    const uint32_t S = std::max(1u, e->g.S), tcl = std::max(1u, e->g.tuple_cls);
    const Dev& h = e->devs[(size_t)(e->multi() ? e->pos2dev[0] : 0)];
    const uint64_t n_t = h.tuples_landed;
    const uint64_t c0 = e->running ? cluster0_tuples(e) : 0;
    auto pair32 = [](uint64_t hi, uint64_t lo) { return ((hi & 0xFFFFFFFFull) << 32) | (lo & 0xFFFFFFFFull); };
    switch (addr) {
        case 220: *data = (uint64_t)e->state; break;                          // pcie_receiver_fsm_state
        case 221: *data = e->lines_received & 0xFFFFFFFFull; break;           // pcie_numcls_received
        case 222: {                                                            // progCycles (DTInference.sv:347-357)
            double ns = 0;
            if (e->prog_seen) ns = ns_since(e->t_prog0, e->prog_open ? now : e->t_prog1);
            *data = cycles_of(e, ns) & 0xFFFFFFFFull;
            break;
        }
        case 223: {                                                            // execCycles (:330-345): start -> process_done
            double ns = 0;
            if (e->running) ns = ns_since(e->t_start, now);
            else if (e->done_latched) ns = ns_since(e->t_start, e->t_done);
            *data = cycles_of(e, ns) & 0xFFFFFFFFull;
            break;
        }
        case 224: *data = e->sl3_lines_sent & 0xFFFFFFFFull; break;           // num_sent_lines: lines that left for other devices
        case 225: {                                                            // num_sent_packets (data_packet_numcls, reg 206[31:24])
            const uint64_t pk = std::max<uint64_t>(1, (e->regs[6] >> 24) & 0xFF);
            *data = ((e->sl3_lines_sent + pk - 1) / pk) & 0xFFFFFFFFull;
            break;
        }
        case 226: *data = e->sl3_lines_sent & 0xFFFFFFFFull; break;           // packets_lines
        case 121: *data = pair32(c0, e->cum_c0 + c0); break;                   // {cluster_out_valids, cluster_tuples_res_out[0]}
        case 122: *data = pair32(n_t, (e->cum_c0 + c0) * S); break;            // {num_out_tuples, cluster_tree_res_out[0]}
        case 123: *data = pair32(n_t * tcl, e->prog_lines_local); break;       // {data_lines, prog_lines}
        case 124: *data = pair32(n_t, e->cum_c0 + c0); break;                  // {aggreg_tuples_in, cluster_reduce_tree_outs[0]}
        case 125: *data = pair32(e->cum_c0 + c0, e->sl3_res_lines); break;     // {cluster_reduce_tree_outs_valids[0], sl3_res_lines}
        case 126: *data = pair32(e->cum_c0 + c0, e->cum_c0_lines + c0 * tcl); break;   // {cluster_tuples_received[0], cluster_lines_received[0]}
        default: *data = 0xFFFFFFFFFFFFFFFFull; break;                         // EngineCSR.sv:123
    }
    return DTE_OK;
}

int dte_stream_write(dte_t* e, const void* cl128, size_t n_lines) {
    if (!e || (!cl128 && n_lines)) return DTE_ERR_ARG;
    const unsigned char* p = static_cast<const unsigned char*>(cl128);
    bool landed = false;
    while (n_lines) {
        if (e->state == dte_engine::ST_TREES) {
            const uint64_t total = e->regs[2] & 0xFFFFFFFFull, wtotal = e->regs[2] >> 32;   // reg 202
            if (total == 0 || wtotal == 0 || wtotal >= total)
                return fail(e, DTE_ERR_CONFIG, "reg 202: total_num_trees_cls=%llu total_num_weights_cls=%llu",
                            (unsigned long long)total, (unsigned long long)wtotal);
            if (!e->prog_seen) { e->prog_seen = e->prog_open = true; e->t_prog0 = Clock::now(); }
            // Chunking (PCIeReceiver.sv:241-264): unless broadcast_trees, the weights stream is cut every
            // numcls_local_weights lines and the index stream every numcls_local_findexes lines, chunk i going to
            // devices_list[i % numDevs]; entry 0 is the host node itself (:160-178).  currDevID keeps rotating
            // from the weights into the index stream, as in the RTL.
            Flags f = decode_flags(e);
            const bool cut = f.multiple && !f.bcast_trees && f.ndev > 1;
            if (cut && (e->regs[3] & 0xFFFFFFFFull) == 0 && wtotal % e->g.w_cls == 0) {
                // extension: both 16-bit chunk fields left at 0 = "cut the trees evenly over numDevs" — chunks of more
                // than 65535 lines (e.g. 1024 trees of depth 10 per device) cannot be written into reg 203
                const uint64_t per = (wtotal / e->g.w_cls + f.ndev - 1) / f.ndev;
                f.chunk_w = per * e->g.w_cls;
                f.chunk_f = per * e->g.f_cls;
            }
            if (cut && (f.chunk_f == 0 || (!e->multi() && e->node_index >= f.ndev)))
                return fail(e, DTE_ERR_CONFIG, "reg 203: numcls_local_findexes=%llu numDevs=%u node=%u",
                            (unsigned long long)f.chunk_f, f.ndev, e->node_index);
            const size_t take = (size_t)std::min<uint64_t>(n_lines, total - e->lines_received);
            if (!cut) {
                const uint64_t w_take = std::min<uint64_t>(take, wtotal > e->lines_received ? wtotal - e->lines_received : 0);
                e->tree_w[0].insert(e->tree_w[0].end(), p, p + w_take * 16);
                e->tree_f[0].insert(e->tree_f[0].end(), p + w_take * 16, p + take * 16);
                e->prog_lines_local += take;
                if (e->multi()) e->sl3_lines_sent += take;                       // broadcast to the ring as well
            } else {
                for (size_t i = 0; i < take; ++i) {
                    const bool is_w = e->lines_received + i < wtotal;
                    if (owner_of(e, e->cur_dev) >= 0) {
                        auto& v = is_w ? e->tree_w[e->cur_dev] : e->tree_f[e->cur_dev];
                        v.insert(v.end(), p + i * 16, p + i * 16 + 16);
                    }
                    if (e->cur_dev == (e->multi() ? 0u : e->node_index)) e->prog_lines_local++;
                    if (e->cur_dev != 0) e->sl3_lines_sent++;
                    uint64_t& cnt = is_w ? e->cur_w : e->cur_f;
                    if (++cnt == (is_w ? f.chunk_w : f.chunk_f)) {
                        cnt = 0;
                        e->cur_dev = (e->cur_dev + 1 == f.ndev) ? 0 : e->cur_dev + 1;
                    }
                }
            }
            e->lines_received += take;
            p += take * 16;
            n_lines -= take;
            if (e->lines_received == total) {
                // prog_mode = (numcls_received < total_num_weights_cls), PCIeReceiver.sv:136-139
                int rc = program_devices(e, e->g);
                for (auto& v : e->tree_w) { v.clear(); v.shrink_to_fit(); }
                for (auto& v : e->tree_f) { v.clear(); v.shrink_to_fit(); }
                if (!rc) rc = prepare_data_phase(e, true);
                e->t_prog1 = Clock::now();
                e->prog_open = false;
                if (rc) { e->state = dte_engine::ST_IDLE; e->running = false; return rc; }
                e->state = dte_engine::ST_DATA;                                 // WAIT_DATA -> RECEIVE_DATA
                e->cur_dev = 0;
                e->cur_d = 0;
            }
        } else if (e->state == dte_engine::ST_DATA) {
            // back-pressure: the result queue is bounded (pcie_full_out, PCIeReceiver.sv:126)
            const size_t tcl = e->g.tuple_cls;
            uint64_t in_engine = (e->data_lines_in + n_lines) / tcl;                         // results queued once these lines are in
            if (!e->multi() && e->deal) in_engine = in_engine / e->ndev_ring + e->deal_lines / tcl;   // this node keeps 1/numDevs of them
            in_engine = in_engine > e->res_read ? in_engine - e->res_read : 0;
            if (in_engine > e->ring_cap)
                return fail(e, DTE_ERR_BACKPRESSURE, "result queue full (%llu results would wait, capacity %llu): read result lines first",
                            (unsigned long long)in_engine, (unsigned long long)e->ring_cap);
            const size_t tb = e->g.tuple_bytes();
            const uint64_t lines_before = e->data_lines_in;
            if (e->group) {
                // data broadcast (InputDistributor.sv:199-204): every device receives every line
                TRY(land_group(e, p, n_lines * 16));
                for (Dev& d : e->devs) d.tuples_landed = (e->data_lines_in + n_lines) * 16 / tb;
                e->sl3_lines_sent += n_lines;
            } else if (!e->deal) {
                TRY(land_dev(e, e->devs[0], p, n_lines * 16));
                e->devs[0].tuples_landed = (e->data_lines_in + n_lines) * 16 / tb;
            } else {
                // data dealing (PCIeReceiver.sv:298-307): batches of core_data_batch_cls lines round-robin over the ring
                size_t left = n_lines;
                const unsigned char* q = p;
                while (left) {
                    const size_t run = (size_t)std::min<uint64_t>(left, e->deal_lines - e->cur_d);
                    const int dv = owner_of(e, e->cur_dev);
                    if (dv >= 0) {
                        Dev& d = e->devs[(size_t)dv];
                        TRY(land_dev(e, d, q, run * 16));
                        d.tuples_landed = d.slot[d.cur].fill / tb - d.slot[d.cur].walked + d.res_enq;   // whole tuples so far on this device
                    }
                    if (e->cur_dev != 0) e->sl3_lines_sent += run;
                    e->cur_d += run;
                    if (e->cur_d == e->deal_lines) { e->cur_d = 0; e->cur_dev = (e->cur_dev + 1 == e->ndev_ring) ? 0 : e->cur_dev + 1; }
                    q += run * 16;
                    left -= run;
                }
            }
            landed = true;
            e->data_lines_in += n_lines;
            e->lines_received += n_lines;
            e->tuples_in += (e->data_lines_in / tcl) - (lines_before / tcl);
            n_lines = 0;
        } else {
            return fail(e, DTE_ERR_STATE, "stream_write while the receiver is idle (write reg 200 first)");
        }
    }
    if (landed) TRY(sync_sources(e));                     // the caller may reuse its buffer when we return
    return DTE_OK;
}

int dte_stream_flush(dte_t* e) {
    if (!e) return DTE_ERR_ARG;
    if (e->state != dte_engine::ST_DATA) return DTE_OK;
    return flush_all(e);
}

int dte_stream_read_packets(dte_t* e, void* cl128, uint8_t* last_flags, size_t max_lines, size_t* got) {
    if (!e || !got || (!cl128 && max_lines)) return DTE_ERR_ARG;
    *got = 0;
    if (!e->ring_cap) return DTE_OK;                      // no run yet
    if (e->state == dte_engine::ST_DATA) TRY(flush_all(e));
    // 4 consecutive results per line, word j = tuple 4m+j (ResultsCombiner.sv:132-162).  Wait for the device work
    // of tuples already written, never for input that has not arrived.
    const uint64_t submitted = global_prefix(e, false);
    const uint64_t want_t = std::min<uint64_t>(e->res_read + (uint64_t)max_lines * 4, submitted);
    for (Dev& d : e->devs) TRY(poll_dev(e, d, 0));
    if (global_prefix(e, true) < want_t) {
        if (e->multi() && e->deal) {
            for (Dev& d : e->devs) TRY(poll_dev(e, d, d.res_enq));
        } else {
            TRY(poll_dev(e, e->devs[(size_t)(e->group ? e->pos2dev[0] : 0)], want_t));
        }
    }
    const uint64_t avail = std::min(global_prefix(e, true), want_t);
    const size_t n = avail > e->res_read ? (size_t)((avail - e->res_read) / 4) : 0;
    float* out = static_cast<float*>(cl128);
    uint64_t i = e->res_read;
    const uint64_t end = e->res_read + (uint64_t)n * 4;
    const int host_ix = e->multi() ? e->pos2dev[0] : 0;
    while (i < end) {
        int dv; uint64_t loc;
        locate_result(e, i, dv, loc);
        const Dev& d = e->devs[(size_t)dv];
        uint64_t run = end - i;
        if (e->multi() && e->deal) {
            const uint64_t bt = e->deal_lines / e->g.tuple_cls;
            run = std::min(run, bt - i % bt);
        }
        const size_t pos = (size_t)(loc % d.ring_cap);
        run = std::min<uint64_t>(run, d.ring_cap - pos);
        memcpy(out + (i - e->res_read), d.h_ring + pos, (size_t)run * 4);
        if (dv != host_ix) e->sl3_res_lines += (run + 3) / 4;
        i += run;
    }
    if (e->group) e->sl3_res_lines += n;                  // aggregate mode: every result line came back over the ring
    if (last_flags) {
        // `last` closes a PCIe packet every pcie_out_packet_numcls lines: reg 206[55:48], compared as the
        // 8-bit "minus one" copy (EngineCSR.sv:242, DTInference.sv:632-646,659-663)
        const uint32_t pkt_m1 = (uint32_t)(((e->regs[6] >> 48) & 0xFF) - 1) & 0xFF;
        for (size_t k = 0; k < n; ++k) {
            last_flags[k] = e->out_cl_count == pkt_m1;
            e->out_cl_count = (e->out_cl_count == pkt_m1) ? 0 : ((e->out_cl_count + 1) & 0xFF);
        }
    }
    e->res_read = end;
    e->result_lines_out += n;
    *got = n;
    bool done;
    return check_done(e, &done);
}

int dte_stream_read(dte_t* e, void* cl128, size_t max_lines, size_t* got) {
    return dte_stream_read_packets(e, cl128, nullptr, max_lines, got);
}

int dte_process_done(dte_t* e, int* done) {
    if (!e || !done) return DTE_ERR_ARG;
    bool d = false;
    TRY(check_done(e, &d));
    *done = d ? 1 : 0;
    return DTE_OK;
}

int dte_load_ensemble(dte_t* e, const void* weight_cls, size_t n_weight_cls, const void* findex_cls,
                      size_t n_findex_cls, uint32_t first_tree, uint32_t num_local_trees) {
    if (!e || !weight_cls || !findex_cls || !n_weight_cls || !n_findex_cls) return DTE_ERR_ARG;
    const Clock::time_point t0 = Clock::now();
    Geom g;
    TRY(decode_geom(e, g));
    const unsigned char* wl = static_cast<const unsigned char*>(weight_cls);
    const unsigned char* fl = static_cast<const unsigned char*>(findex_cls);
    PackedEnsemble pk;
    if (!e->multi()) {
        TRY(pack_or_fail(e, g, wl, n_weight_cls, fl, n_findex_cls, first_tree, num_local_trees, pk));
        TRY(upload_ensemble(e, e->devs[0], pk));
    } else {
        TRY(decide_partition(e, g));
        if (n_weight_cls % g.w_cls) return fail(e, DTE_ERR_ARG, "weights stream: %zu lines is not a multiple of %u lines per tree", n_weight_cls, g.w_cls);
        const uint32_t T_all = (uint32_t)(n_weight_cls / g.w_cls);
        if (num_local_trees == 0 && first_tree == 0) num_local_trees = T_all;
        const uint32_t G = (uint32_t)e->devs.size();
        if (!e->group) {                             // replicated
            TRY(pack_or_fail(e, g, wl, n_weight_cls, fl, n_findex_cls, first_tree, num_local_trees, pk));
            for (Dev& d : e->devs) TRY(upload_ensemble(e, d, pk));
        } else {
            // contiguous chunks in ring order (PCIeReceiver.sv:241-264): reg 203's numcls_local_weights if it is a
            // whole number of trees that covers the ensemble, else an even split
            const Flags f = decode_flags(e);
            const bool programmed = (e->regs[3] & 0xFFFF) != 0 && f.chunk_w % g.w_cls == 0 && (f.chunk_w / g.w_cls) * G >= num_local_trees;
            const uint32_t per = programmed ? (uint32_t)(f.chunk_w / g.w_cls) : (num_local_trees + G - 1) / G;
            for (uint32_t r = 0; r < G; ++r) {
                const uint32_t lo = std::min(num_local_trees, r * per), hi = std::min(num_local_trees, (r + 1) * per);
                if (hi == lo) return fail(e, DTE_ERR_CONFIG, "ring position %u would hold no trees (%u trees over %u devices)", r, num_local_trees, G);
                TRY(pack_or_fail(e, g, wl, n_weight_cls, fl, n_findex_cls, first_tree + lo, hi - lo, pk));
                TRY(upload_ensemble(e, e->devs[(size_t)e->pos2dev[r]], pk));
            }
        }
    }
    e->g = g;
    e->prog_partition = !e->multi() ? 0 : (e->group ? 2 : 1);
    memcpy(e->prog_pos2dev, e->pos2dev, sizeof e->prog_pos2dev);
    e->prog_seen = true; e->prog_open = false;
    e->t_prog0 = t0; e->t_prog1 = Clock::now();
    return DTE_OK;
}

static int infer_device_common(dte_t* e, const void* d_tuples, size_t n, float* d_scores, uint8_t* d_labels, void* cuda_stream,
                               bool accumulate, const char* what) {
    TRY(single_only(e, what));
    Geom g;
    TRY(decode_geom(e, g));
    TRY(check_resident(e, g));
    Dev& d = e->devs[0];
    d.g.K = g.K; d.g.S = g.S; d.g.missing = g.missing;
    e->g = g;
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    cudaStream_t st = pick_stream(d, cuda_stream);
    CUDA_TRY(e, cudaEventRecord(d.ev_t0, st));
    const char* why = nullptr;
    cudaError_t rc = launch_walk(d, e->tune, e->forced_variant, d_tuples, n, d_scores, d_labels, st, accumulate, &why);
    if (rc != cudaSuccess) return fail(e, why ? DTE_ERR_ARG : DTE_ERR_CUDA, "walk kernel launch failed: %s", why ? why : cudaGetErrorString(rc));
    CUDA_TRY(e, cudaEventRecord(d.ev_t1, st));
    d.timing_pending = true;
    e->tuples_in += n;
    e->tuples_out += n;
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

int dte_infer_device(dte_t* e, const void* d_tuples, size_t n, float* d_scores, uint8_t* d_labels, void* cuda_stream) {
    if (!e || (n && (!d_tuples || !d_scores))) return DTE_ERR_ARG;
    return infer_device_common(e, d_tuples, n, d_scores, d_labels, cuda_stream, false, "dte_infer_device");
}

int dte_infer_device_accumulate(dte_t* e, const void* d_tuples, size_t n, float* d_scores_accum, void* cuda_stream) {
    if (!e || (n && (!d_tuples || !d_scores_accum))) return DTE_ERR_ARG;
    return infer_device_common(e, d_tuples, n, d_scores_accum, nullptr, cuda_stream, true, "dte_infer_device_accumulate");
}

// ---- peer-visible buffers (CUDA IPC): partial-score buffers other processes' combine kernels read ----
int dte_ipc_alloc(dte_t* e, size_t bytes, void** d_ptr, unsigned char handle_out[64]) {
    if (!e || !d_ptr || !handle_out || !bytes) return DTE_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CUDA_TRY(e, cudaSetDevice(e->devs[0].ordinal));
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
    CUDA_TRY(e, cudaSetDevice(e->devs[0].ordinal));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CUDA_TRY(e, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DTE_OK;
}

int dte_ipc_close(dte_t* e, void* d_ptr, int owner) {
    if (!e || !d_ptr) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaSetDevice(e->devs[0].ordinal));
    if (owner) CUDA_TRY(e, cudaFree(d_ptr));
    else CUDA_TRY(e, cudaIpcCloseMemHandle(d_ptr));
    return DTE_OK;
}

int dte_infer_host(dte_t* e, const void* h_tuples, size_t n, float* h_scores, uint8_t* h_labels) {
    if (!e || (n && (!h_tuples || !h_scores))) return DTE_ERR_ARG;
    return infer_host(e, static_cast<const unsigned char*>(h_tuples), n, h_scores, h_labels);
}

int dte_host_alloc(dte_t* e, size_t bytes, void** h_ptr) {
    if (!e || !h_ptr || !bytes) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaSetDevice(e->devs[0].ordinal));
    CUDA_TRY(e, cudaHostAlloc(h_ptr, bytes, cudaHostAllocPortable));
    return DTE_OK;
}

int dte_host_free(dte_t* e, void* h_ptr) {
    if (!e || !h_ptr) return DTE_ERR_ARG;
    CUDA_TRY(e, cudaFreeHost(h_ptr));
    return DTE_OK;
}

int dte_labels_device(dte_t* e, const float* d_scores, size_t n, uint8_t* d_labels, void* cuda_stream) {
    if (!e || (n && (!d_scores || !d_labels))) return DTE_ERR_ARG;
    if (!n) return DTE_OK;
    Dev& d = e->devs[0];
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    cudaStream_t st = pick_stream(d, cuda_stream);
    labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scores, d_labels, n);
    CUDA_TRY(e, cudaGetLastError());
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

int dte_ring_add_device(dte_t* e, const float* d_a, const float* d_b, float* d_out, size_t n, void* cuda_stream) {
    if (!e || (n && (!d_a || !d_b || !d_out))) return DTE_ERR_ARG;
    if (!n) return DTE_OK;
    Dev& d = e->devs[0];
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    cudaStream_t st = pick_stream(d, cuda_stream);
    ring_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, d_b, d_out, n);
    CUDA_TRY(e, cudaGetLastError());
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

int dte_ring_combine_device(dte_t* e, const float* const* d_parts, int n_parts, size_t n, float* d_out, uint8_t* d_labels,
                            void* cuda_stream) {
    if (!e || !d_parts || n_parts < 1 || n_parts > kMaxRing || (n && !d_out)) return DTE_ERR_ARG;
    RingParts parts;
    for (int g = 0; g < n_parts; ++g) {
        if (!d_parts[g] || (reinterpret_cast<uintptr_t>(d_parts[g]) & 3u)) return fail(e, DTE_ERR_ARG, "part %d: null or not 4-byte aligned", g);
        parts.p[g] = d_parts[g];
    }
    if (!n) return DTE_OK;
    Dev& d = e->devs[0];
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    cudaStream_t st = pick_stream(d, cuda_stream);
    CUDA_TRY(e, launch_ring_combine(parts, n_parts, d_out, d_labels, n, d.sm_count, st));
    d.kernel_launches++;
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
    Dev& d0 = e->devs[0];
    if (d0.timing_pending) {
        cudaSetDevice(d0.ordinal);
        if (cudaEventSynchronize(d0.ev_t1) == cudaSuccess) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, d0.ev_t0, d0.ev_t1) == cudaSuccess) {
                d0.last_walk_ms = ms;
                e->last_walk_ms = ms;
            }
        }
        d0.timing_pending = false;
    }
    memset(info, 0, sizeof *info);
    info->num_levels = d0.d_top ? d0.g.D : e->g.D;
    info->num_features = d0.d_top ? d0.g.F() : e->g.F();
    info->clusters = e->g.K;
    info->trees_per_pu = e->g.S;
    info->sm_count = (uint32_t)d0.sm_count;
    for (Dev& d : e->devs) {
        info->ensemble_bytes += d.ensemble_bytes;
        info->kernel_launches += d.kernel_launches;
        if (e->group) info->num_trees += d.T;
    }
    if (!e->group) info->num_trees = d0.T;
    info->last_walk_ms = e->last_walk_ms;
    info->num_devices = (uint32_t)e->devs.size();
    info->partition = !e->multi() ? 0u : (e->group ? 2u : 1u);
    info->tuples_in = e->tuples_in;
    info->tuples_out = e->tuples_out;
    if (d0.d_top) {
        Plan pl = make_plan(d0, e->tune, e->forced_variant);
        info->kernel_variant = (uint32_t)pl.variant;
        info->tuples_per_cta = pl.variant == KERNEL_GENERIC ? 128u : 32u * (uint32_t)(pl.nwarps / pl.pair);
    }
    return DTE_OK;
}

int dte_kernel_name(dte_t* e, char* buf, size_t len) {
    if (!e || !buf || !len) return DTE_ERR_ARG;
    if (!e->devs[0].d_top) return fail(e, DTE_ERR_STATE, "no ensemble loaded");
    kernel_name(e->devs[0], e->tune, e->forced_variant, buf, len);
    return DTE_OK;
}

int dte_autotune(dte_t* e, size_t n_tuples, char* report, size_t report_len) {
    if (!e) return DTE_ERR_ARG;
    Dev& d = e->devs[0];
    if (!d.d_top) return fail(e, DTE_ERR_STATE, "no ensemble loaded");
    Geom g;
    TRY(decode_geom(e, g));
    TRY(check_resident(e, g));
    TRY(drain_all(e));
    const uint32_t F = d.g.F();
    n_tuples = std::max<size_t>(n_tuples, (size_t)d.sm_count * 192 * 4);
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    void* d_x = nullptr;
    float* d_s = nullptr;
    CUDA_TRY(e, cudaMalloc(&d_x, n_tuples * F * 4));
    if (cudaMalloc(&d_s, n_tuples * 4) != cudaSuccess) { cudaFree(d_x); return fail(e, DTE_ERR_NOMEM, "autotune: out of device memory"); }
    synth_tuples_kernel<<<(unsigned)d.sm_count * 8, 256, 0, d.s_main>>>(static_cast<uint32_t*>(d_x), 0, (unsigned long long)n_tuples * F, 0x7091E5ull, 10000u, d.g.missing);
    struct Cand { int ilp, pair, stages, phased, warps; };      // warps = 10: the 384-thread (168-register) instantiation
    const Cand cands[] = {{4, 2, 1, 1, 0}, {4, 2, 1, 0, 0}, {4, 2, 1, 1, 10}, {4, 2, 1, 0, 10}, {2, 4, 1, 1, 0}, {2, 4, 1, 0, 0}, {8, 1, 1, 0, 0},
                          {8, 1, 2, 0, 0}, {4, 1, 2, 0, 0}, {2, 2, 2, 0, 0}, {2, 2, 1, 0, 0}};
    const Tune saved = e->tune;
    Tune best = saved;
    float best_ms = 1e30f;
    std::string rep;
    char name[256], line[400];
    int rc = DTE_OK;
    for (const Cand& c : cands) {
        Tune t = saved;
        t.ilp = c.ilp; t.pair = c.pair; t.stages = c.stages; t.phased = c.phased; t.warps = c.warps;
        const Plan pl = make_plan(d, t, KERNEL_AUTO);
        if (pl.variant != KERNEL_TILE_STAGED || pl.ilp != c.ilp || pl.pair != c.pair || pl.nstages != c.stages) continue;   // does not fit
        if (c.warps) {                               // only meaningful when the cap actually removes a tuple group
            Tune u = t; u.warps = 0;
            if (make_plan(d, u, KERNEL_AUTO).nwarps <= c.warps) continue;
        }
        if (c.phased && phased_level(d, t, pl) == 0xFFFFFFFFu) continue;
        float ms = 0;
        for (int rep_i = 0; rep_i < 2 && !rc; ++rep_i) {                     // the second launch is the measurement
            const char* why = nullptr;
            cudaEventRecord(d.ev_t0, d.s_main);
            cudaError_t st = launch_walk(d, t, KERNEL_AUTO, d_x, n_tuples, d_s, nullptr, d.s_main, false, &why);
            cudaEventRecord(d.ev_t1, d.s_main);
            if (st != cudaSuccess || cudaEventSynchronize(d.ev_t1) != cudaSuccess) { rc = fail(e, DTE_ERR_CUDA, "autotune launch failed: %s", cudaGetErrorString(st)); break; }
            cudaEventElapsedTime(&ms, d.ev_t0, d.ev_t1);
        }
        if (rc) break;
        kernel_name(d, t, KERNEL_AUTO, name, sizeof name);
        snprintf(line, sizeof line, "%8.3f ms  %7.2f M tuples/s  %s\n", ms, n_tuples / (ms * 1e3), name);
        rep += line;
        if (ms < best_ms) { best_ms = ms; best = t; }
    }
    cudaFree(d_x);
    cudaFree(d_s);
    if (rc) { e->tune = saved; return rc; }
    e->tune = best;
    kernel_name(d, e->tune, e->forced_variant, name, sizeof name);
    rep += std::string("chosen: ") + name + "\n";
    if (report && report_len) snprintf(report, report_len, "%s", rep.c_str());
    return DTE_OK;
}

int dte_set_node(dte_t* e, uint32_t node_index) {
    if (!e || node_index >= (uint32_t)kMaxRing) return DTE_ERR_ARG;            // devices_list has 20 entries (EngineCSR.sv:250-296)
    if (e->multi()) return fail(e, DTE_ERR_STATE, "a multi-device handle owns every ring position");
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
    Dev& d = e->devs[0];
    CUDA_TRY(e, cudaSetDevice(d.ordinal));
    cudaStream_t st = pick_stream(d, cuda_stream);
    const unsigned long long ne = n * num_features;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((ne + 255) / 256, (unsigned long long)d.sm_count * 32);
    synth_tuples_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(d_tuples), first_tuple * num_features, ne, seed,
                                                missing_ppm, missing_value);
    CUDA_TRY(e, cudaGetLastError());
    if (!cuda_stream) CUDA_TRY(e, cudaStreamSynchronize(st));
    return DTE_OK;
}

}  // extern "C"
