#!/usr/bin/env python3
"""Writes tests/golden/kats.json — HAND-DERIVED known-answer tests for the tree-walk path.

The reference ships no tests and no vectors, so every expected value below was
derived BY HAND from the cited RTL lines — none is computed by the oracle or by the engine.  The
script only serialises them (floats -> bit patterns).  Paths are relative to the reference root.

A case = trees (heap arrays W/FI, all with the same number of levels D), summation geometry (K, S),
the missing-value pattern, tuples, and the expected raw score word per tuple.
"""
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def f(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


MISS = 0xBF800000  # bits(-1.0f)
NEG0 = 0x80000000


def const_tree_d1(leaf):
    """D=1 tree whose two leaves are equal: the leaf does not depend on the tuple."""
    return {"W": [f(0.5), leaf, leaf], "FI": [0]}


cases = []

# 1 -- D=1, one tree: DTPU.sv:655-657 `right = !(x < thr)`, leaf = W[2n+1+right] (:710,731).
cases.append(dict(
    name="d1_basic_direction", D=1, K=1, S=1, missing=MISS, F=4,
    why="x<thr -> left leaf W[1]; x>thr -> right leaf W[2]; x==thr is 'not smaller' -> right (DTPU.sv:655-657)",
    trees=[{"W": [f(0.5), f(10.0), f(20.0)], "FI": [2]}],
    tuples=[[0, 0, f(0.25), 0], [0, 0, f(0.75), 0], [0, 0, f(0.5), 0]],
    expect=[f(10.0), f(20.0), f(20.0)],
))

# 2 -- missing value: DTPU.sv:653 raw equality with CSR 205[31:0]; :659 FI bit 13 picks the side.
cases.append(dict(
    name="missing_bit13", D=1, K=1, S=1, missing=MISS, F=4,
    why="feature == missing pattern: bit13=0 -> left even though -1.0 < 0.5 would also be left; "
        "bit13=1 -> right although the value compares smaller (DTPU.sv:653,659,667)",
    trees=[{"W": [f(0.5), f(1.0), f(2.0)], "FI": [1]},
           {"W": [f(0.5), f(4.0), f(8.0)], "FI": [1 | (1 << 13)]}] + [const_tree_d1(0)] * 6,
    tuples=[[0, MISS, 0, 0], [0, f(0.75), 0, 0]],
    # tuple 0: tree0 -> 1.0 (left), tree1 -> 8.0 (right): ((1+8)+(0+0))+((0+0)+(0+0)) = 9
    # tuple 1: 0.75 >= 0.5 -> right in both: 2 + 8 = 10
    expect=[f(9.0), f(10.0)],
))

# 3 -- the comparator is a SIGNED INT32 compare of the raw words (DTPU.sv:655), not IEEE.
cases.append(dict(
    name="both_negative_int32_compare", D=1, K=1, S=1, missing=MISS, F=4,
    why="v=-3.0 (0xC0400000) vs thr=-2.0 (0xC0000000): as int32 v > thr so NOT smaller -> right, although "
        "IEEE says -3 < -2; mirror case v=-2.0 thr=-3.0 -> smaller -> left (DTPU.sv:655)",
    trees=[{"W": [f(-2.0), f(1.0), f(2.0)], "FI": [0]},
           {"W": [f(-3.0), f(16.0), f(32.0)], "FI": [1]}] + [const_tree_d1(0)] * 6,
    tuples=[[f(-3.0), f(-2.0), 0, 0]],
    # tree0: right -> 2.0 ; tree1: int32(0xC0000000) < int32(0xC0400000) -> smaller -> left -> 16.0
    expect=[f(18.0)],
))

# 4 -- signed zeros through the integer comparator.
cases.append(dict(
    name="signed_zero_compare", D=1, K=1, S=1, missing=MISS, F=4,
    why="-0.0 = 0x80000000 = INT_MIN is smaller than +0.0 -> left; +0.0 vs thr -0.0 -> not smaller -> right "
        "(DTPU.sv:655; IEEE would call them equal)",
    trees=[{"W": [0, f(1.0), f(2.0)], "FI": [0]},
           {"W": [NEG0, f(16.0), f(32.0)], "FI": [1]}] + [const_tree_d1(0)] * 6,
    tuples=[[NEG0, 0, 0, 0]],
    expect=[f(33.0)],   # 1.0 (left) + 32.0 (right)
))

# 5 -- D=2 walk, heap addressing 2n+1 / 2n+2 (DTPU.sv:594-596,710-712).
cases.append(dict(
    name="d2_heap_addressing", D=2, K=1, S=1, missing=MISS, F=4,
    why="root f0 thr .5; node1 f1 thr .5; node2 f2 thr .5; leaves W[3..6] = 1,2,3,4",
    trees=[{"W": [f(0.5), f(0.5), f(0.5), f(1.0), f(2.0), f(3.0), f(4.0)], "FI": [0, 1, 2]}],
    tuples=[[f(0.1), f(0.1), f(0.9), 0], [f(0.1), f(0.9), f(0.1), 0], [f(0.9), f(0.9), f(0.1), 0], [f(0.9), f(0.1), f(0.9), 0]],
    expect=[f(1.0), f(2.0), f(3.0), f(4.0)],
))

# 6 -- tree8 pairing ((l0+l1)+(l2+l3))+((l4+l5)+(l6+l7)) (FPAddersReduceTree.sv:90-125).
cases.append(dict(
    name="tree8_pairing_order", D=1, K=1, S=1, missing=MISS, F=4,
    why="leaves [1e8, 1, -1e8, 1, 0,0,0,0]: 1e8+1 = 1e8 (ulp 8), -1e8+1 = -1e8 -> 0.0; a left-to-right sum "
        "would give 1.0",
    trees=[const_tree_d1(f(1e8)), const_tree_d1(f(1.0)), const_tree_d1(f(-1e8)), const_tree_d1(f(1.0))] + [const_tree_d1(0)] * 4,
    tuples=[[0, 0, 0, 0]],
    expect=[0],
))
cases.append(dict(
    name="tree8_pairing_order_b", D=1, K=1, S=1, missing=MISS, F=4,
    why="leaves [1e8, -1e8, 1, 1, 0...]: (0)+(2) = 2.0",
    trees=[const_tree_d1(f(1e8)), const_tree_d1(f(-1e8)), const_tree_d1(f(1.0)), const_tree_d1(f(1.0))] + [const_tree_d1(0)] * 4,
    tuples=[[0, 0, 0, 0]],
    expect=[f(2.0)],
))

# 7 -- slot/cluster order: tree t -> cluster (t/8)%K, slot t/(8K) (Core.sv:291-304); per cluster
#      acc = r_s + acc over slots (FPAggregator.v:79-131); then acc = part_c + acc over clusters (Core.sv:486-542).
def group(val):
    return [const_tree_d1(val)] + [const_tree_d1(0)] * 7

cases.append(dict(
    name="cluster_slot_order_k2s2", D=1, K=2, S=2, missing=MISS, F=4,
    why="group sums r0=1e8 (c0,s0) r1=1 (c1,s0) r2=-1e8 (c0,s1) r3=1 (c1,s1): acc_c0 = -1e8+1e8 = 0, "
        "acc_c1 = 1+1 = 2, total 2.0; stream-order summation ((r0+r1)+r2)+r3 would give 1.0",
    trees=group(f(1e8)) + group(f(1.0)) + group(f(-1e8)) + group(f(1.0)),
    tuples=[[0, 0, 0, 0]],
    expect=[f(2.0)],
))
cases.append(dict(
    name="cluster_slot_order_k1s4", D=1, K=1, S=4, missing=MISS, F=4,
    why="same trees, K=1: one cluster, 4 slots, sequential ((( r0)+r1)+r2)+r3 = ((1e8+1)-1e8)+1 = 1.0",
    trees=group(f(1e8)) + group(f(1.0)) + group(f(-1e8)) + group(f(1.0)),
    tuples=[[0, 0, 0, 0]],
    expect=[f(1.0)],
))

# 8 -- T not a multiple of 8K: unprogrammed PU slots contribute 0 (DTPU.sv:544,587,760).
cases.append(dict(
    name="empty_slots_add_zero", D=1, K=2, S=1, missing=MISS, F=4,
    why="11 trees of leaf 1.0 with K=2,S=1: cluster0 holds 8 trees (sum 8), cluster1 holds 3 (+5 empty): 11.0",
    trees=[const_tree_d1(f(1.0))] * 11,
    tuples=[[0, 0, 0, 0]],
    expect=[f(11.0)],
))
# 9 -- slots beyond S are never issued (DTPU.sv:519-531: curr_tree_index runs to num_trees_per_pu-1).
cases.append(dict(
    name="slots_beyond_S_ignored", D=1, K=1, S=1, missing=MISS, F=4,
    why="16 trees of leaf 1.0 but S=1,K=1: only the first 8 trees are walked: 8.0",
    trees=[const_tree_d1(f(1.0))] * 16,
    tuples=[[0, 0, 0, 0]],
    expect=[f(8.0)],
))
# 10 -- round to nearest even of the adder (FPAdder_2cycles_latency.v:360-369).
cases.append(dict(
    name="adder_rne_ties", D=1, K=1, S=1, missing=MISS, F=4,
    why="2^24 + 1 is a tie between 2^24 and 2^24+2 -> even mantissa 2^24; 2^24+2 plus 1 ties -> 2^24+4",
    trees=[const_tree_d1(f(16777216.0)), const_tree_d1(f(1.0)), const_tree_d1(f(16777218.0)), const_tree_d1(f(1.0))] + [const_tree_d1(0)] * 4,
    tuples=[[0, 0, 0, 0]],
    # (2^24 + 1 -> 2^24) + (2^24+2 + 1 -> 2^24+4) = 2^25 + 4 (exact, ulp of 2^25 is 4)
    expect=[f(33554436.0)],
))
# 11 -- exact cancellation gives +0 (FPAdder :376-386 eqdiffsign -> zero; FPAddersReduceTree.sv:141 -> word 0).
cases.append(dict(
    name="cancellation_plus_zero", D=1, K=1, S=1, missing=MISS, F=4,
    why="1.5 + (-1.5) = +0.0 (all-zero word), never -0.0",
    trees=[const_tree_d1(f(1.5)), const_tree_d1(f(-1.5))] + [const_tree_d1(0)] * 6,
    tuples=[[0, 0, 0, 0]],
    expect=[0],
))

# 12 -- trees and tuples that span SEVERAL 128-bit lines, with line padding between trees.
#       D=3: W = 15 words -> 4 lines (1 pad word), FI = 7 entries -> 1 line (1 pad entry); F = 8 -> 2 lines per tuple.
#       Word i of a line is bits [32i+31:32i] (PipelinedMUX.sv:63-65); a tree's arrays start on a line boundary
#       (InputDistributor.sv:276-296 counts whole lines per tree); feature f lives in line f/4, word f%4 (DTPU.sv:628).
def const_tree(D, leaf):
    return {"W": [f(0.5)] * ((1 << D) - 1) + [leaf] * (1 << D), "FI": [0] * ((1 << D) - 1)}

tree_a = {"W": [f(0.5)] * 7 + [f(2.0 ** i) for i in range(8)], "FI": [0, 1, 2, 3, 4, 5, 6]}            # node i tests feature i
tree_b = {"W": [f(0.5)] * 7 + [f(2.0 ** (8 + i)) for i in range(8)], "FI": [6, 5, 4, 3, 2, 1, 0]}      # node i tests feature 6-i
lo, hi = f(0.1), f(0.9)
cases.append(dict(
    name="multi_line_trees_and_tuples", D=3, K=1, S=1, missing=MISS, F=8,
    why="t0 all-left: A 0->1->3->leaf W[7]=1, B 0->1->3->W[7]=256: 257.  t1 x=[R,L,R,L,L,R,L,-]: A 0-R->2 (x2=R)->6 (x6=L)-> "
        "leaf W[13]=64 (4th line of W); B node0 tests x6=L->1, node1 tests x5=R->4, node4 tests x2=R-> leaf W[10]=2^11: 2112 "
        "(features 5 and 6 sit in the tuple's SECOND line)",
    trees=[tree_a, tree_b] + [const_tree(3, 0)] * 6,
    tuples=[[lo] * 8, [hi, lo, hi, lo, lo, hi, lo, 0]],
    expect=[f(257.0), f(2112.0)],
))

# 13 -- R10 stride decision: a PU's second tree starts TRUE-lines-per-tree after the first (the RTL wires the minus-one
#       copies into the stride, DTInference.sv:505-506 vs InputDistributor.sv:284-285 / DTPU.sv:522-523).
#       K=1, S=2: trees 0 and 8 share PU 0 (slot 0 and slot 1).  D=4: W = 31 words -> 8 lines, FI = 15 -> 2 lines.
def tree_d4(first_exp):
    return {"W": [f(0.5)] * 15 + [f(2.0 ** (first_exp + i)) for i in range(16)], "FI": [i % 4 for i in range(15)]}

cases.append(dict(
    name="second_slot_uses_true_line_stride", D=4, K=1, S=2, missing=MISS, F=4,
    why="x=[R,L,R,L]: 0-R->2 (f2=R)->6 (f2=R)->14 (f2=R)-> leaf W[30]: tree 0 gives 2^15, tree 8 gives 2^31; slot 0 sum 2^15, "
        "then acc = 2^31 + 2^15 (exact).  With a 7-line stride tree 8 would be read 4 words early (leaf 2^27).",
    trees=[tree_d4(0)] + [const_tree(4, 0)] * 7 + [tree_d4(16)],
    tuples=[[hi, lo, hi, lo]],
    expect=[f(2147516416.0)],
))

# 14 -- early leaves, bit 14 "next node is leaf" (DTPU.sv:596,661,712).  BUILD-DEFINED RULE (the RTL is broken here: it
#       freezes node_offset WITHOUT the direction bit while the read address keeps advancing, so an early leaf with levels
#       left reads a wrong cell): when a node carries bit 14 the walk ends in its child's cell W[2n+1+right].
#       D=3: node 1 carries bit 14 -> cells W[3], W[4] are leaves (10, 20); node 2's side is a normal 3-level walk.
#       Tree B: bit 14 on the ROOT -> a one-level tree, leaves W[1], W[2]; everything below is don't-care (index 99 > F).
early_a = {"W": [f(0.5), f(0.5), f(0.5), f(10.0), f(20.0), f(0.5), f(0.5)] + [f(float(100 + i)) for i in range(8)],
           "FI": [0, 1 | (1 << 14), 2, 99, 99, 3, 3]}
early_b = {"W": [f(0.5), f(1000.0), f(2000.0)] + [0] * 12, "FI": [1 | (1 << 14)] + [99] * 6}
cases.append(dict(
    name="early_leaf_bit14", D=3, K=1, S=1, missing=MISS, F=4, rtl_literal=False,
    why="t0 x=[L,R,-,-]: A 0-L->1 (bit14) x1=R -> cell W[4]=20; B root bit14 x1=R -> W[2]=2000: 2020.  t1 x=[L,L]: 10+1000=1010.  "
        "t2 x=[R,L,R,R]: A 0-R->2 (x2=R)->6 (x3=R)-> leaf W[14]=107; B x1=L -> 1000: 1107",
    trees=[early_a, early_b] + [const_tree(3, 0)] * 6,
    tuples=[[lo, hi, 0, 0], [lo, lo, 0, 0], [hi, lo, hi, hi]],
    expect=[f(2020.0), f(1010.0), f(1107.0)],
))

# result packing (ResultsCombiner.sv:132-162): 4 consecutive results per line, word j = tuple 4m+j; a trailing group
# of fewer than 4 results is never emitted (:153-155, curr_word only wraps on the 4th fill)
lines_kat = dict(
    name="result_lines_flush_on_fourth",
    why="6 results -> ONE line {r0,r1,r2,r3}; r4, r5 stay in the line register",
    scores=[f(1.0), f(2.0), f(3.0), f(4.0), f(5.0), f(6.0)],
    expect_lines=[[f(1.0), f(2.0), f(3.0), f(4.0)]],
)

# data dealing (PCIeReceiver.sv:298-307): currDCount counts the lines of the current batch up to core_data_batch_cls - 1,
# then currDevID moves to the next entry of devices_list and wraps at numDevs; the host node (entry 0) goes first.
deal_kat = dict(
    name="deal_batches_round_robin",
    why="11 data lines, core_data_batch_cls = 2, numDevs = 3: lines {0,1}->dev0 {2,3}->dev1 {4,5}->dev2 {6,7}->dev0 {8,9}->dev1 {10}->dev2",
    n_lines=11, batch_cls=2, num_devs=3,
    expect=[[[0, 2], [6, 2]], [[2, 2], [8, 2]], [[4, 2], [10, 1]]],      # per device: (first line, number of lines)
)
# tree chunking (PCIeReceiver.sv:241-264): the weights stream is cut every numcls_local_weights lines, the index stream every
# numcls_local_findexes lines, chunk i -> devices_list[i % numDevs]; contiguous trees per device, host node first.
chunk_kat = dict(
    name="contiguous_tree_chunks",
    why="7 trees over 3 devices in chunks of ceil(7/3) = 3 trees: dev0 trees 0-2, dev1 trees 3-5, dev2 tree 6",
    n_trees=7, num_devs=3, expect=[[0, 3], [3, 3], [6, 1]],
)

# ring combine (ResultsCombiner.sv:292-311,359-368): ((p_host + p_1) + p_2)
ring = dict(
    name="ring_order_3dev",
    why="partials 1e8, 1, -1e8 in device order: (1e8+1)-1e8 = 0.0; (p0+p2)+p1 would be 1.0",
    partials=[[f(1e8), f(1.0)], [f(1.0), f(2.0)], [f(-1e8), f(3.0)]],
    expect=[0, f(6.0)],
)

out = {"about": "hand-derived KATs; see make_kats.py for the derivations", "cases": cases, "ring": [ring], "lines": [lines_kat], "deal": [deal_kat],
       "chunks": [chunk_kat]}
with open(os.path.join(HERE, "kats.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", len(cases), "cases")
