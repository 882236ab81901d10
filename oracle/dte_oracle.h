/*
 * dte_oracle.h — CPU ORACLE for the decision-tree-ensemble hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (distributed-decisiontrees_b200/libdte.so) never links, loads or calls anything in oracle/.
 *
 * It is a hand restatement in C of the reference's SystemVerilog semantics
 * (fpgasystems/Distributed-DecisionTrees @ 6b48669, paths relative to /root/reference):
 *   - tree walk / compare / leaf read .......... rtl/DTEngine/core/DTPU.sv:579-761
 *   - tree -> (cluster, PU, slot) placement ..... rtl/DTEngine/Core.sv:280-375
 *   - 8-input fp32 reduce tree ................. rtl/DTEngine/core/FPAddersReduceTree.sv:90-141
 *   - sequential fp32 accumulator .............. rtl/DTEngine/core/FPAggregator.v:79-131
 *   - cross-cluster sum ........................ rtl/DTEngine/Core.sv:486-542
 *   - fp32 adder (FloPoCo 8/23, RNE, no subnormals)  rtl/DTEngine/common/FPAdder_2cycles_latency.v:210-389
 *   - cross-device ring add, 4-per-line packing  rtl/DTEngine/ResultsCombiner.sv:132-162,292-311
 *   - stream framing (weights | findexes | data) rtl/DTEngine/PCIeReceiver.sv:136-139,
 *                                                rtl/DTEngine/InputDistributor.sv:248-296
 *
 * PARITY STATUS: **parity unpinned by the reference** — the reference ships no tests, no
 * golden vectors and no executable implementation of this path (the RTL cannot be
 * simulated here: no Verilog simulator, missing vendor IP).  The oracle is pinned instead by
 *   (1) hand-derived known-answer tests (tests/golden/kat_*.json, derivations cite RTL lines),
 *   (2) two independently written walkers (clean heap walker vs PU-memory address-literal
 *       walker) diffed on random ensembles,
 *   (3) two independently written adders (host-float RNE+FTZ model vs bit-level restatement
 *       of the FloPoCo netlist) diffed on random operands.
 */
#ifndef DTE_ORACLE_H
#define DTE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Run-time parameters exactly as the CSR file presents them (EngineCSR.sv:189-306). */
typedef struct {
    uint32_t num_levels;      /* D : reg 205[35:32]  comparison levels, leaves are level D */
    uint32_t clusters;        /* K : reg 205[47:44]  clusters that share one tuple */
    uint32_t trees_per_pu;    /* S : reg 205[43:36]  tree slots walked per PU */
    uint32_t missing_value;   /*     reg 205[31:0]   raw 32-bit pattern meaning "feature missing" */
    uint32_t tree_w_cls;      /*     reg 204[31:16]  128-bit lines per tree in the weights stream */
    uint32_t tree_f_cls;      /*     reg 204[47:32]  128-bit lines per tree in the feature-index stream */
    uint32_t tuple_cls;       /*     reg 204[63:48]  128-bit lines per tuple (F = 4*tuple_cls) */
    uint32_t num_trees;       /* T : trees present in the streams (total_num_weights_cls / tree_w_cls) */
} dteo_cfg;

/* --- fp32 adder, two independent restatements (operands/results are raw IEEE bit patterns) --- */
uint32_t dteo_fpadd(uint32_t a, uint32_t b);          /* host float RNE + flush-below-normal-to-+0 */
uint32_t dteo_fpadd_literal(uint32_t a, uint32_t b);  /* bit-level FloPoCo datapath, incl. in/out wrappers */

/* --- one tree, one tuple: returns the raw leaf word --- */
uint32_t dteo_leaf(const dteo_cfg* c, const uint32_t* w_tree, const uint16_t* fi_tree, const uint32_t* x);

/* --- full per-device score for n tuples.  weights: T*tree_w_cls lines, findex: T*tree_f_cls lines,
 *     tuples: n*tuple_cls lines, scores: n raw fp32 words.  literal_adder!=0 selects dteo_fpadd_literal.
 *     threads<=1 : single thread; otherwise that many pthreads, tuples split contiguously. Returns 0 or a negative error. --- */
int dteo_scores(const dteo_cfg* c, const void* weights_cls, const void* findex_cls,
                const void* tuple_cls, size_t n, uint32_t* scores, int literal_adder, int threads);

/* --- same result computed the way the hardware stores and addresses it: builds the 64 PU memory
 *     images in arrival order and walks them with DTPU.sv's address arithmetic.  Only defined inside
 *     the hardware limits (S*tree_w_cls <= 2048 lines, S*tree_f_cls <= 1024 lines, F <= 2047, tuple <= 512 lines).
 *     Returns 0, or -2 when the configuration exceeds those limits. --- */
int dteo_scores_literal(const dteo_cfg* c, const void* weights_cls, const void* findex_cls,
                        const void* tuple_cls, size_t n, uint32_t* scores);

/* --- ResultsCombiner: ring add of G partial score vectors in ring order (host first),
 *     ((p0 + p1) + p2) + ...   (ResultsCombiner.sv:292-311,359-368) --- */
void dteo_ring_combine(const uint32_t* const* partials, int G, size_t n, uint32_t* out);

/* --- build-defined label rule (the reference has no labels): label = score > 0.0f --- */
void dteo_labels(const uint32_t* scores, size_t n, uint8_t* labels);

/* --- the same scores with the loops interchanged (a block of tuples walks one tree8 group at a time, so the trees stay
 *     in cache); identical arithmetic and per-tuple summation order — bit-identical to dteo_scores(literal_adder=0).
 *     This is the loop bench.py times as the CPU baseline. --- */
int dteo_scores_blocked(const dteo_cfg* c, const void* weights_cls, const void* findex_cls,
                        const void* tuple_cls, size_t n, uint32_t* scores, int threads);

/* --- ResultsCombiner line packing: 4 results per 128-bit line, a trailing group of < 4 is not emitted
 *     (ResultsCombiner.sv:132-162).  lines must hold 4*(n/4) words; returns the number of lines. --- */
size_t dteo_result_lines(const uint32_t* scores, size_t n, uint32_t* lines);

/* --- host cores: online, and usable (online ∩ affinity mask ∩ cgroup CPU quota) --- */
int dteo_online_cpus(void);
int dteo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
