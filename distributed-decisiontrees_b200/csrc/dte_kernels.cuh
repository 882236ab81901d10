// dte_kernels.cuh — sm_100a tree-walk kernels of the B200 decision-tree-ensemble engine.
//
// What they compute (per tuple x, complete trees, one device) is the reference's Core output
// (rtl/DTEngine/Core.sv:486-542 on top of core/DTPUCluster.sv:188-222 and core/DTPU.sv:579-761):
//
//   score(x) = SUM_seq{j<K}  SUM_seq{s<S}  tree8{p<8}( leaf(x, tree t = (s*K + j)*8 + p) )
//   leaf(x,t): n = 0; D times { right = (x[fi] == missing) ? fi.bit13 : !((int32)x[fi] < (int32)thr); n = 2n+1+right }
//   tree8(l)  = ((l0+l1)+(l2+l3))+((l4+l5)+(l6+l7)),  every add = add.rn.ftz.f32
//
// B200 mapping (nothing here resembles the FPGA pipeline):
//   * lane = tuple.  A warp owns 32 tuples and walks the SAME tree in all lanes, so the upper tree
//     levels are shared-memory broadcasts and the tuple tile can be stored feature-major
//     (xs[f][column]) which makes every per-lane feature fetch bank-conflict free.
//   * one lane accumulates all trees of its tuple in the reference's summation order, so scores
//     are bit-exact with the oracle.  Two warps serve each 32-tuple group (4 trees each of every
//     tree8 group, half-sums exchanged through shared memory in pairing order): the tile, not the
//     warp count, fills the SM, so this doubles the warps per SM for free.
//   * inside a warp the 4 (8) walks are issued in a skewed order — node load, feature load and
//     compare of different walks interleave — so an in-order warp always has independent work
//     between a shared-memory load and its first use.
//   * the ensemble is repacked on load: levels 0..D-3 as 8-byte heap records ("top"), and the two
//     last comparison levels plus their four leaves as ONE 32-byte sector-aligned record
//     ("bottom") fetched with one 256-bit load (LDG.E.256, new on sm_100) — the deep, divergent
//     part of a walk costs one L2 sector instead of three, and the fetch of step q is consumed
//     after the top walk of step q+1 (two register buffers).
//   * TILE_STAGED: a producer warp streams the top parts of the next trees into a shared-memory
//     ring with cp.async.bulk + mbarrier (TMA bulk copy engine) while the consumer warps walk.
//     A stage is refilled in two parts on separate barriers (levels 0..D-5 | D-4..D-3) that the
//     consumers hand back at different points of the walk, so ONE 64 KiB stage overlaps copy and
//     walk like a double buffer ("phased refill", used for deep trees).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dte {

struct WalkParams {
    const uint2* top;           // [trees_padded][top_stride] {thr bits, fidx | (8+8*missing_right)<<16}
    const uint4* bottom;        // [trees_padded][nb][BOT_VEC] see repack in dte_engine.cu
    const float* tuples;        // [n][F] row-major fp32 (the reference's tuple CLs)
    float* scores;              // [n]
    uint8_t* labels;            // [n] or nullptr
    unsigned long long n;
    uint32_t F;
    uint32_t Dtop;              // comparison levels held in `top` (= max(D,2) - 2)
    uint32_t top_stride;        // records per tree in `top` (= max(1, 2^Dtop))
    uint32_t nb;                // bottom records per tree (= 2^Dtop)
    uint32_t groups;            // tree8 groups to walk (= min(S*K, ceil(T/8)))
    uint32_t K;                 // clusters per tuple: group g accumulates into partial g % K
    uint32_t missing;           // raw missing-value pattern
    uint32_t tiles;             // ceil(n / tuples_per_cta)
    uint32_t nwarps;            // consumer warps per CTA (tuples_per_cta = 32 * nwarps)
    uint32_t nstages;           // ring depth (TILE_STAGED)
    uint32_t fill_split;        // 0: one bulk copy per ring stage; 1: one per tree
    uint32_t Lw;                // phased ring refill: level iteration at which part B is needed; 0xFFFFFFFF = off
    uint32_t accumulate;        // 1: scores[i] += partial with a system-scope reduction (fused cross-device combine)
    uint32_t wide_rows;         // 1: tuple rows are 32-byte aligned (F % 8 == 0, base aligned): 256-bit tile loads
};

// ---------------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fadd_ref(float a, float b) {
    // the reference adder: round-to-nearest-even, no subnormals (FPAdder_2cycles_latency.v:360-386)
    float r;
    asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// fire-and-forget fp32 add into memory that may belong to a peer GPU (NVLink): the ResultsCombiner
// hop (ResultsCombiner.sv:292-311) done by the walk kernel's own epilogue instead of a collective
__device__ __forceinline__ void red_add_sys(float* addr, float v) {
    // fp32 reductions on GLOBAL memory round to nearest even and flush subnormal inputs/results to zero (PTX ISA,
    // atom/red .add.f32) — the same arithmetic as add.rn.ftz.f32; the instruction takes no .ftz modifier.
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ uint2 ldg64_nc(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg128_nc(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 256-bit load (sm_100+: LDG.E.256): one instruction, one sector per lane for the bottom records
__device__ __forceinline__ void ldg256_nc(const void* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}
__device__ __forceinline__ void ldg256_nc_na(const void* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {   // tuples: read once, keep out of L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void ldg256_stream(const void* p, uint4& lo, uint4& hi) {   // tuples: one request per 32 B sector
    ldg256_nc_na(p, lo, hi);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// bulk async copy global -> shared of this CTA, completion counted on an mbarrier (TMA bulk engine)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// direction of one node step as a byte increment on the 8-byte heap offset: left 8, right 16
__device__ __forceinline__ uint32_t step_inc(uint32_t x, uint32_t thr, uint32_t meta, uint32_t missing) {
    uint32_t inc = ((int32_t)x < (int32_t)thr) ? 8u : 16u;   // go right iff !(x < thr), DTPU.sv:655-657
    if (x == missing) inc = meta >> 16;                       // DTPU.sv:653,659,667
    return inc;
}

// Same decision as step_inc(), phrased on predicates so that only {compare, 3-input predicate LUT,
// select} separate the feature value from the next node address: returns a_right if the walk goes
// right (feature missing ? missing-goes-right flag : !(x < thr)), else a_left.
__device__ __forceinline__ uint32_t step_select(uint32_t x, uint32_t thr, uint32_t meta, uint32_t missing,
                                                uint32_t a_left, uint32_t a_right) {
    uint32_t r;
    asm("{\n\t"
        ".reg .pred pge, pm, pmr, t1, t2, npm, pr;\n\t"
        "setp.ge.s32 pge, %1, %2;\n\t"          // !(x < thr), signed compare of the raw words (DTPU.sv:655)
        "setp.eq.u32 pm, %1, %3;\n\t"           // feature == missing pattern (DTPU.sv:653)
        "setp.ge.u32 pmr, %4, 1048576;\n\t"     // meta >> 16 == 16  <=>  missing goes right (DTPU.sv:659)
        "and.pred t1, pm, pmr;\n\t"
        "not.pred npm, pm;\n\t"
        "and.pred t2, npm, pge;\n\t"
        "or.pred pr, t1, t2;\n\t"
        "selp.u32 %0, %6, %5, pr;\n\t"
        "}"
        : "=r"(r) : "r"(x), "r"(thr), "r"(missing), "r"(meta), "r"(a_left), "r"(a_right));
    return r;
}

// select partial `j` out of a register file of 8 (j is warp-uniform)
__device__ __forceinline__ float acc_get(const float (&a)[8], uint32_t j) {
    float v = a[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) v = (j == (uint32_t)i) ? a[i] : v;
    return v;
}
__device__ __forceinline__ void acc_set(float (&a)[8], uint32_t j, float v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (j == (uint32_t)i) ? v : a[i];
}

// ---------------------------------------------------------------------------------------------
// bottom record: two comparison levels + four leaves.
//   compact (every feature index < 512), 32 B:  {thr_p, thr_l, thr_r, fpack} {leaf LL, LR, RL, RR}
//       fpack = fp | fl<<10 | fr<<20, each 10-bit field = fidx | missing_right<<9
//   wide (any index < 2048), 64 B: {thr_p, thr_l, thr_r, meta_p} {meta_l, meta_r, 0, 0} {leaves} {0}
//       meta = fidx | missing_right<<16
// FEAT(f) fetches feature f of this lane's tuple.
// ---------------------------------------------------------------------------------------------
template <bool WIDE> struct BotRec;
template <> struct BotRec<false> { uint4 a, b; };
template <> struct BotRec<true> { uint4 a, b, c; };
#ifndef DTE_BOT_NA
#define DTE_BOT_NA 1     // bottom records are touched once per (tuple, tree): do not allocate them in L1
#endif
__device__ __forceinline__ void bot_load(BotRec<false>& r, const uint4* rec) {
#if DTE_BOT_NA
    ldg256_nc_na(rec, r.a, r.b);
#else
    ldg256_nc(rec, r.a, r.b);
#endif
}
__device__ __forceinline__ void bot_load(BotRec<true>& r, const uint4* rec) {
    ldg256_nc(rec, r.a, r.b);
    r.c = ldg128_nc(rec + 2);
}
template <bool WIDE, class Feat>
__device__ __forceinline__ float bot_finish(const BotRec<WIDE>& r, uint32_t missing, Feat feat) {
    uint32_t fp, fl, fr, mp, ml, mr;
    uint4 leaves;
    if constexpr (WIDE) {
        fp = r.a.w & 0xFFFFu; mp = r.a.w >> 16;
        fl = r.b.x & 0xFFFFu; ml = r.b.x >> 16;
        fr = r.b.y & 0xFFFFu; mr = r.b.y >> 16;
        leaves = r.c;
    } else {
        uint32_t k = r.a.w;
        fp = k & 0x1FFu;         mp = (k >> 9) & 1u;
        fl = (k >> 10) & 0x1FFu; ml = (k >> 19) & 1u;
        fr = (k >> 20) & 0x1FFu; mr = (k >> 29) & 1u;
        leaves = r.b;
    }
    // all three features are fetched up front (conflict-free in the tile kernels): one dependent
    // round trip instead of two
    uint32_t xp = feat(fp), xl = feat(fl), xr = feat(fr);
    bool r1 = (xp == missing) ? (mp != 0) : !((int32_t)xp < (int32_t)r.a.x);
    uint32_t thr = r1 ? r.a.z : r.a.y;
    uint32_t xc = r1 ? xr : xl;
    uint32_t mc = r1 ? mr : ml;
    bool r2 = (xc == missing) ? (mc != 0) : !((int32_t)xc < (int32_t)thr);
    uint32_t lo = r2 ? leaves.y : leaves.x;
    uint32_t hi = r2 ? leaves.w : leaves.z;
    return __uint_as_float(r1 ? hi : lo);
}

// ---------------------------------------------------------------------------------------------
// Kernel 1 — generic: one thread per tuple, nodes AND features from global memory.
// Any F (up to 2047) and any D; used when the feature-major tile does not fit shared memory and as
// an independent cross-check of the tile kernels in the tests.
// ---------------------------------------------------------------------------------------------
template <bool WIDE>
__global__ void __launch_bounds__(128) dt_walk_generic(const WalkParams p) {
    constexpr int BV = WIDE ? 4 : 2;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t* x = reinterpret_cast<const uint32_t*>(p.tuples) + i * p.F;
    auto feat = [&](uint32_t f) { return __ldg(x + f); };
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
    for (uint32_t g = 0; g < p.groups; ++g) {
        float l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = 0.0f;
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const uint32_t t = g * 8 + q;
            const char* tp = reinterpret_cast<const char*>(p.top + (size_t)t * p.top_stride);
            uint32_t o = 0;
            for (uint32_t lvl = 0; lvl < p.Dtop; ++lvl) {
                uint2 nd = ldg64_nc(tp + o);
                uint32_t xv = feat(nd.y & 0xFFFFu);
                o = 2 * o + step_inc(xv, nd.x, nd.y, p.missing);
            }
            const uint32_t j = (o >> 3) - (p.nb - 1);
            BotRec<WIDE> br;
            bot_load(br, p.bottom + ((size_t)t * p.nb + j) * BV);
            float leaf = bot_finish<WIDE>(br, p.missing, feat);
            // static register index for l[q]
#pragma unroll
            for (int k = 0; k < 8; ++k) l[k] = (k == q) ? leaf : l[k];
        }
        float r = fadd_ref(fadd_ref(fadd_ref(l[0], l[1]), fadd_ref(l[2], l[3])),
                           fadd_ref(fadd_ref(l[4], l[5]), fadd_ref(l[6], l[7])));
        const uint32_t j = g % p.K;
        acc_set(acc, j, fadd_ref(r, acc_get(acc, j)));    // acc = r + acc, FPAggregator.v:79-131
    }
    float tot = 0.0f;
    for (uint32_t j = 0; j < p.K; ++j) tot = fadd_ref(acc_get(acc, j), tot);   // Core.sv:486-542
    if (p.accumulate) {
        red_add_sys(p.scores + i, tot);
    } else {
        p.scores[i] = tot;
        if (p.labels) p.labels[i] = tot > 0.0f ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2/3 — tile kernels.  Persistent CTAs (one per SM).  A CTA holds G tuple groups of 32
// tuples; each group is served by P consumer warps (P = 1 or 2) that share the group's tile
// columns and split every stage of ILP*P trees between them (+ 1 producer warp when STAGED).
// P = 2 doubles the warps per SM without costing shared memory — the tile, not the warp count,
// is what fills the SM — so the hardware scheduler, not a static interleave, hides the
// shared-memory latency.  Shared memory:
//   [0,128)                       mbarriers: fullA, emptyA, fullB, emptyB per ring stage (<= 4 stages)
//   [128, 128+2048)               P = 2: half-sum exchange, [group][step parity][lane] fp32
//   [kHdrBytes, + ring)           STAGED: nstages x (ILP*P) tree tops (top_stride * 8 B each)
//   [.., + F * M * 4)             xs[f][M] feature-major tuple tile, M = 32 * G columns
// Group g owns columns 32g..32g+31 exclusively, so tile reloads need no CTA-wide barrier.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kBarBytes = 128;
constexpr uint32_t kXchBytes = 2048;
constexpr uint32_t kHdrBytes = kBarBytes + kXchBytes;

__device__ __forceinline__ void group_barrier(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// NT = thread bound of the instantiation.  Registers are allocated per SM sub-partition (16 K each), so the
// per-thread limit is set by ceil(warps / 4): 9..12 warps -> 168 registers, 13..16 warps -> 128.  The planner
// picks the 384-thread instantiation whenever the plan fits in 12 warps (cfg3: 5 tuple groups x 2 + producer).
template <int ILP, int P, bool STAGED, bool WIDE, int NT>
__global__ void __launch_bounds__(NT, 1) dt_walk_tile(const WalkParams p) {
    static_assert(ILP * P == 8 || ILP * P == 4, "a step covers a tree8 group or half of one");
    constexpr int BV = WIDE ? 4 : 2;
    constexpr int SP = ILP * P;                         // trees per ring stage / per step
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t G = p.nwarps / P;                    // tuple groups per CTA
    const uint32_t M = 32u * G;
    const uint32_t tree_bytes = p.top_stride * 8u;
    const uint32_t stage_bytes = (uint32_t)SP * tree_bytes;
    const uint32_t ring_bytes = STAGED ? p.nstages * stage_bytes : 0u;
    const uint32_t xs_base = sbase + kHdrBytes + ring_bytes;
    const uint32_t row_bytes = M * 4u;
    const uint32_t steps = p.groups * (8 / SP);         // SP trees per step
    const uint32_t my_tiles = (p.tiles > blockIdx.x) ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (STAGED) {
        if (threadIdx.x == 0) {
            for (uint32_t s = 0; s < p.nstages; ++s) {
                mbar_init(sbase + 8 * s, 1);                        // full (part A): producer's arrive + tx bytes
                mbar_init(sbase + 8 * (p.nstages + s), p.nwarps);   // empty (part A): one arrive per consumer warp
                mbar_init(sbase + 8 * (2 * p.nstages + s), 1);      // full, part B  (phased refill only)
                mbar_init(sbase + 8 * (3 * p.nstages + s), p.nwarps);   // empty, part B
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (warp == p.nwarps) {
            // ---------------- producer: stream tree tops through the ring ----------------
            if (lane == 0) {
                const uint64_t total = (uint64_t)my_tiles * steps;
                const char* src0 = reinterpret_cast<const char*>(p.top);
                uint32_t slot = 0, par = 1;                          // fresh barrier: parity-1 wait passes
                uint32_t q = 0;
                // Phased refill (p.Lw set): a stage is refilled in two parts with their own barriers.
                //   part A = the first 8*2^(Lw+1) bytes of every tree top (levels 0..Lw and the first record of
                //   level Lw+1), part B = the rest (75 % of the bytes).  The consumers hand part A back before
                //   their last level and part B after it, and need part B only from level iteration Lw on, so
                //   ONE stage overlaps copy and walk: B of step q+1 streams in under levels 0..Lw of step q+1,
                //   A of step q+1 under the last level + bottom records + leaf sums of step q.
                const bool phased = p.Lw != 0xFFFFFFFFu;
                const uint32_t bytesA = phased ? (16u << p.Lw) : tree_bytes;
                const uint32_t bytesB = tree_bytes - bytesA;
                for (uint64_t it = 0; it < total; ++it) {
                    const uint32_t dst = sbase + kHdrBytes + slot * stage_bytes;
                    const char* src = src0 + (size_t)q * stage_bytes;
                    mbar_wait(sbase + 8 * (p.nstages + slot), par);
                    const uint32_t full = sbase + 8 * slot;
                    if (!phased) {
                        mbar_arrive_expect_tx(full, stage_bytes);
                        if (p.fill_split == 0) {
                            bulk_g2s(dst, src, stage_bytes, full);                 // the SP tree tops are contiguous: one copy
                        } else {
#pragma unroll
                            for (int c = 0; c < SP; ++c) bulk_g2s(dst + c * tree_bytes, src + (size_t)c * tree_bytes, tree_bytes, full);
                        }
                    } else {
                        mbar_arrive_expect_tx(full, (uint32_t)SP * bytesA);
#pragma unroll
                        for (int c = 0; c < SP; ++c) bulk_g2s(dst + c * tree_bytes, src + (size_t)c * tree_bytes, bytesA, full);
                        mbar_wait(sbase + 8 * (3 * p.nstages + slot), par);
                        const uint32_t fullB = sbase + 8 * (2 * p.nstages + slot);
                        mbar_arrive_expect_tx(fullB, (uint32_t)SP * bytesB);
#pragma unroll
                        for (int c = 0; c < SP; ++c)
                            bulk_g2s(dst + c * tree_bytes + bytesA, src + (size_t)c * tree_bytes + bytesA, bytesB, fullB);
                    }
                    if (++q == steps) q = 0;
                    if (++slot == p.nstages) { slot = 0; par ^= 1; }
                }
            }
            return;
        }
    }

    // ---------------- consumers ----------------
    const uint32_t grp = warp / P, sub = warp % P;      // tuple group, position inside the warp pair
    const uint32_t col = grp * 32u + lane;
    const uint32_t xcol = xs_base + col * 4u;
    // exchange slots: P=2: 2 x 128 B per group (step parity); P=4: 3 x 128 B per group (sender sub-1).
    // 6 groups x 256 B or 5 groups x 384 B both stay inside kXchBytes.
    static_assert(6 * 256 <= kXchBytes && 5 * 384 <= kXchBytes, "exchange area too small");
    const uint32_t xch = sbase + kBarBytes + grp * (P == 4 ? 384u : 256u) + lane * 4u;
    auto feat = [&](uint32_t f) { return lds32(xcol + f * row_bytes); };
    uint32_t slot = 0, par = 0;

    for (uint32_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        // ---- load this group's 32 tuples, transposing row-major global -> feature-major shared;
        //      the P warps of the group take alternate blocks of 16 features ----
        __syncwarp();
        {
            const unsigned long long m = (unsigned long long)tile * M + col;
            const bool live = m < p.n;
            const uint4* row = reinterpret_cast<const uint4*>(p.tuples + (live ? m : 0ull) * p.F);
            const uint32_t nvec = p.F >> 2;
            uint32_t v = 4 * sub;
            auto put4 = [&](uint32_t a, const uint4& d) {
                sts32(a, d.x); sts32(a + row_bytes, d.y); sts32(a + 2 * row_bytes, d.z); sts32(a + 3 * row_bytes, d.w);
            };
            if (p.wide_rows) {
                // rows are 32-byte aligned: one 256-bit request per sector (LDG.E.256) instead of two 128-bit ones
                for (; v + 4 <= nvec; v += 4 * P) {
                    uint4 d0, d1, d2, d3;
                    ldg256_stream(row + v, d0, d1);
                    ldg256_stream(row + v + 2, d2, d3);
                    if (!live) { d0 = d1 = d2 = d3 = make_uint4(0, 0, 0, 0); }
                    const uint32_t a = xcol + (4 * v) * row_bytes;
                    put4(a, d0); put4(a + 4 * row_bytes, d1); put4(a + 8 * row_bytes, d2); put4(a + 12 * row_bytes, d3);
                }
            } else {
                for (; v + 4 <= nvec; v += 4 * P) {
                    uint4 d0 = ldg128_stream(row + v), d1 = ldg128_stream(row + v + 1);
                    uint4 d2 = ldg128_stream(row + v + 2), d3 = ldg128_stream(row + v + 3);
                    if (!live) { d0 = d1 = d2 = d3 = make_uint4(0, 0, 0, 0); }
                    const uint32_t a = xcol + (4 * v) * row_bytes;
                    put4(a, d0); put4(a + 4 * row_bytes, d1); put4(a + 8 * row_bytes, d2); put4(a + 12 * row_bytes, d3);
                }
            }
            if (sub == 0) {                              // ragged tail (F/4 not a multiple of 4)
                for (v = nvec & ~3u; v < nvec; ++v) {
                    uint4 d0 = ldg128_stream(row + v);
                    if (!live) d0 = make_uint4(0, 0, 0, 0);
                    put4(xcol + (4 * v) * row_bytes, d0);
                }
            }
        }
        if (P > 1) group_barrier(1 + grp, 32 * P); else __syncwarp();

        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        float half0 = 0.0f;

        // finish one step: resolve the ILP bottom records into leaves, tree8-reduce, accumulate in the
        // reference's order.  Called one step late so the record fetch overlaps the next top walk.
        auto finish = [&](const BotRec<WIDE> (&rec)[ILP], uint32_t qf) {
            float l[ILP];
#pragma unroll
            for (int c = 0; c < ILP; ++c) l[c] = bot_finish<WIDE>(rec[c], p.missing, feat);
            // (1) this warp's share of the tree8 pairing ((l0+l1)+(l2+l3))+((l4+l5)+(l6+l7))
            float h;
            if constexpr (ILP == 8)
                h = fadd_ref(fadd_ref(fadd_ref(l[0], l[1]), fadd_ref(l[2], l[3])),
                             fadd_ref(fadd_ref(l[4], l[5]), fadd_ref(l[6], l[7])));
            else if constexpr (ILP == 4)
                h = fadd_ref(fadd_ref(l[0], l[1]), fadd_ref(l[2], l[3]));
            else
                h = fadd_ref(l[0], l[1]);
            // (2) combine the P warps of the group in pairing order; only sub 0 continues
            if constexpr (P == 2) {
                const uint32_t slot_x = xch + 128u * (qf & 1u);              // double-buffered by step parity
                if (sub == 1) sts32(slot_x, __float_as_uint(h));
                group_barrier(1 + grp, 64);
                if (sub == 0) h = fadd_ref(h, __uint_as_float(lds32(slot_x)));
            } else if constexpr (P == 4) {
                if (sub != 0) sts32(xch + 128u * (sub - 1u), __float_as_uint(h));
                group_barrier(1 + grp, 128);
                if (sub == 0) {
                    const float s1 = __uint_as_float(lds32(xch)), s2 = __uint_as_float(lds32(xch + 128u));
                    const float s3 = __uint_as_float(lds32(xch + 256u));
                    h = fadd_ref(fadd_ref(h, s1), fadd_ref(s2, s3));
                }
                group_barrier(1 + grp, 128);             // single exchange buffer: readers done before the next write
            }
            // (3) a step covers SP = ILP*P trees: a whole tree8 group, or half of one
            float r;
            bool group_done;
            if constexpr (SP == 8) {
                r = h;
                group_done = (sub == 0);
            } else {
                group_done = (sub == 0) && (qf & 1u) != 0;
                r = fadd_ref(half0, h);
                half0 = h;
            }
            if (group_done) {
                const uint32_t j = (qf / (8 / SP)) % p.K;
                acc_set(acc, j, fadd_ref(r, acc_get(acc, j)));
            }
        };

        // one step, first half: walk the top levels of this warp's ILP trees of stage q and ISSUE the
        // loads of their bottom records into `dst` (not consumed here)
        auto walk_top = [&](uint32_t q, BotRec<WIDE> (&dst)[ILP]) {
            const uint32_t t0 = q * SP + sub * ILP;
            uint32_t o[ILP];
#pragma unroll
            for (int c = 0; c < ILP; ++c) o[c] = 0;
            if (STAGED) {
                mbar_wait(sbase + 8 * slot, par);
                const uint32_t tb = sbase + kHdrBytes + slot * stage_bytes + sub * ILP * tree_bytes;
                // Absolute shared-memory addresses: with A = tb_c + o the child address is
                // A' = tb_c + 2o + 8 + 8*right = (2A + 8 - tb_c) + 8*right.  2A + (8 - tb_c) does not depend
                // on the feature, so only {compare, select, add} sit between the feature and the next node.
                uint32_t A[ILP], kb8[ILP];
#pragma unroll
                for (int c = 0; c < ILP; ++c) { A[c] = tb + c * tree_bytes; kb8[c] = 8u - A[c]; }
                // Skewed schedule inside the warp.  A walk step has three dependent stages:
                //   S1 node record  = LDS.64 [A]            S2 feature = LDS [xs + fidx*row]
                //   S3 compare + child address (ALU)
                // Issuing all S1, then all S2, then all S3 (lock-step) leaves two shared-memory
                // latencies per level exposed to an in-order warp.  Instead the ILP chains are split in
                // two halves half a level apart: tick k issues S3+S1 of chain k and S2 of chain k+-H, so
                // every load has H ticks (~H*12 instructions) of independent work before its first use.
                // (The loads are `asm volatile`, so this program order IS the issue order.)
                constexpr int H = ILP / 2;
                uint2 nd[ILP];
                uint32_t xv[ILP];
                auto S1 = [&](int c) { nd[c] = lds64(A[c]); };
                auto S2 = [&](int c) { xv[c] = feat(nd[c].y & 0xFFFFu); };
                auto S3 = [&](int c) {
                    const uint32_t a2 = A[c] + A[c] + kb8[c];
                    A[c] = step_select(xv[c], nd[c].x, nd[c].y, p.missing, a2, a2 + 8u);
                };
#pragma unroll
                for (int c = 0; c < ILP; ++c) S1(c);                         // level 0 nodes
#pragma unroll
                for (int c = 0; c < H; ++c) S2(c);                           // level 0 features, first half
                const bool phased = p.Lw != 0xFFFFFFFFu;
                for (uint32_t lvl = 0; lvl + 1 < p.Dtop; ++lvl) {
                    if (phased && lvl == p.Lw) mbar_wait(sbase + 8 * (2 * p.nstages + slot), par);   // part B landed
#pragma unroll
                    for (int k = 0; k < H; ++k) { S3(k); S1(k); S2(k + H); }         // S2(k+H) still at level lvl
#pragma unroll
                    for (int k = H; k < ILP; ++k) { S3(k); S1(k); S2(k - H); }       // S2(k-H) already at level lvl+1
                }
                if (phased) {                                                // part A is no longer read: hand it back early
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sbase + 8 * (p.nstages + slot));
                }
#pragma unroll
                for (int k = 0; k < H; ++k) { S3(k); S2(k + H); }            // last staged level: no further node
#pragma unroll
                for (int k = H; k < ILP; ++k) S3(k);
#pragma unroll
                for (int c = 0; c < ILP; ++c) o[c] = A[c] - (tb + c * tree_bytes);
                __syncwarp();
                if (lane == 0) mbar_arrive(sbase + 8 * ((phased ? 3 * p.nstages : p.nstages) + slot));
                if (++slot == p.nstages) { slot = 0; par ^= 1; }
            } else {
                const char* tp = reinterpret_cast<const char*>(p.top + (size_t)t0 * p.top_stride);
                for (uint32_t lvl = 0; lvl < p.Dtop; ++lvl) {
                    uint2 nd[ILP];
                    uint32_t xv[ILP];
#pragma unroll
                    for (int c = 0; c < ILP; ++c) nd[c] = ldg64_nc(tp + (size_t)c * tree_bytes + o[c]);
#pragma unroll
                    for (int c = 0; c < ILP; ++c) xv[c] = feat(nd[c].y & 0xFFFFu);
#pragma unroll
                    for (int c = 0; c < ILP; ++c) o[c] = 2 * o[c] + step_inc(xv[c], nd[c].x, nd[c].y, p.missing);
                }
            }
            // bottom: last two comparison levels + leaves, ONE 32 B (64 B) record per walk
#pragma unroll
            for (int c = 0; c < ILP; ++c) {
                const uint32_t j = (o[c] >> 3) - (p.nb - 1);
                bot_load(dst[c], p.bottom + ((size_t)(t0 + c) * p.nb + j) * BV);
            }
        };

        // Software pipeline with two register buffers (no copies): the records issued by step q are
        // consumed after the top walk of step q+1, so their L2 latency hides behind shared-memory work.
        BotRec<WIDE> recA[ILP], recB[ILP];
        if (steps > 0) {
            walk_top(0, recA);
            uint32_t q = 1;
            for (; q + 1 < steps; q += 2) {
                walk_top(q, recB);
                finish(recA, q - 1);
                walk_top(q + 1, recA);
                finish(recB, q);
            }
            if (q < steps) {
                walk_top(q, recB);
                finish(recA, q - 1);
                finish(recB, q);
            } else {
                finish(recA, q - 1);
            }
        }

        if (sub == 0) {
            float tot = 0.0f;
            for (uint32_t j = 0; j < p.K; ++j) tot = fadd_ref(acc_get(acc, j), tot);
            const unsigned long long m = (unsigned long long)tile * M + col;
            if (m < p.n) {
                if (p.accumulate) {
                    red_add_sys(p.scores + m, tot);
                } else {
                    p.scores[m] = tot;
                    if (p.labels) p.labels[m] = tot > 0.0f ? 1 : 0;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// small helpers: labels, one ring-add hop, synthetic tuples
// ---------------------------------------------------------------------------------------------
__global__ void labels_kernel(const float* __restrict__ s, uint8_t* __restrict__ l, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) l[i] = s[i] > 0.0f ? 1 : 0;
}
__global__ void ring_add_kernel(const float* a, const float* b, float* out, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fadd_ref(a[i], b[i]);       // ResultsCombiner.sv:292-311, one hop
}
// ResultsCombiner aggregate mode in ONE kernel (ResultsCombiner.sv:292-311,359-368): the host node injects its
// partial, every following node of the ring emits local + incoming, so
//     score = (((p[0] + p[1]) + p[2]) + ...) + p[G-1]          each add = add.rn.ftz.f32
// p[g] may live on a PEER GPU (mapped through peer access or CUDA IPC): the loads ride NVLink, the sum is
// formed in ring order in registers, so the result is bit-exact with the reference ring, unlike a library
// reduction whose order is free.  4 tuples per thread = one 128-bit result line (ResultsCombiner.sv:132-162).
constexpr int kMaxRing = 20;                 // devices_list has 20 entries (EngineCSR.sv:250-296)
struct RingParts { const float* p[kMaxRing]; };
__device__ __forceinline__ uint4 ldg128_peer(const void* p) {   // read-once peer data: do not pollute L1
    uint4 v;
    asm volatile("ld.global.relaxed.sys.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer32(const float* p) {
    float v;
    asm volatile("ld.global.relaxed.sys.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(256) ring_combine_kernel(const RingParts parts, int G, float* __restrict__ out,
                                                           uint8_t* __restrict__ labels, unsigned long long n, int vec) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    // vec = 1: every pointer is 16-byte aligned (labels 4-byte): whole result lines per thread
    const unsigned long long nline = vec ? n >> 2 : 0;
    for (unsigned long long i = tid; i < nline; i += stride) {
        uint4 a = ldg128_peer(reinterpret_cast<const uint4*>(parts.p[0]) + i);
        float s0 = __uint_as_float(a.x), s1 = __uint_as_float(a.y), s2 = __uint_as_float(a.z), s3 = __uint_as_float(a.w);
#pragma unroll 4
        for (int g = 1; g < G; ++g) {
            const uint4 b = ldg128_peer(reinterpret_cast<const uint4*>(parts.p[g]) + i);
            s0 = fadd_ref(__uint_as_float(b.x), s0); s1 = fadd_ref(__uint_as_float(b.y), s1);     // local + incoming
            s2 = fadd_ref(__uint_as_float(b.z), s2); s3 = fadd_ref(__uint_as_float(b.w), s3);
        }
        reinterpret_cast<float4*>(out)[i] = make_float4(s0, s1, s2, s3);
        if (labels) {
            const uint32_t l = (s0 > 0.0f ? 1u : 0u) | (s1 > 0.0f ? 0x100u : 0u) | (s2 > 0.0f ? 0x10000u : 0u) | (s3 > 0.0f ? 0x1000000u : 0u);
            reinterpret_cast<uint32_t*>(labels)[i] = l;
        }
    }
    // the rest one tuple at a time: the n % 4 tail, or everything when a pointer is not line-aligned (a partial
    // flush of the landing buffer can start at any tuple)
    for (unsigned long long i = (nline << 2) + tid; i < n; i += stride) {
        float s = ld_peer32(parts.p[0] + i);
        for (int g = 1; g < G; ++g) s = fadd_ref(ld_peer32(parts.p[g] + i), s);
        out[i] = s;
        if (labels) labels[i] = s > 0.0f ? 1 : 0;
    }
}
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void synth_tuples_kernel(uint32_t* out, unsigned long long first_elem, unsigned long long n_elem,
                                    unsigned long long seed, uint32_t missing_ppm, uint32_t missing_value) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n_elem; i += stride) {
        unsigned long long z = splitmix64(seed, first_elem + i);
        float v = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);      // 24 bits -> [0,1), exact
        uint32_t bits = __float_as_uint(v);
        if ((uint32_t)(z & 0xFFFFFFFFull) % 1000000u < missing_ppm) bits = missing_value;
        out[i] = bits;
    }
}

}  // namespace dte
