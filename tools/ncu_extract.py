#!/usr/bin/env python3
"""ncu report -> the few numbers the bench line and profiles/r02_summary.md quote.

    python tools/ncu_extract.py gpurun_out/prof_walk.ncu-rep <tuples profiled> <trees> [tag]

Writes profiles/ncu_pipe.json + profiles/ncu_traffic.json (read by bench.py) and profiles/<tag>_ncu_raw.csv."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    tag = sys.argv[4] if len(sys.argv) > 4 else "r02"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(ROOT, "profiles", tag + "_ncu_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    units = rows[1]
    m = {}
    for h, u, v in zip(hdr, units, vals):
        m[h] = (v, u)

    def num(name, scale_units=True):
        v, u = m[name]
        x = float(v.replace(",", ""))
        if scale_units:
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u)
            if mult:
                x *= mult
        return x

    sms = 148
    lsu_per_sm = num("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg")
    lsu_shared_per_sm = num("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg")
    lsu_global_per_sm = num("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg")
    tma_bytes = num("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum")
    cycles = num("sm__cycles_elapsed.max")
    shared_ld = num("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum")
    shared_st = num("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum")
    conflicts_ld = num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum")
    glob_ld = num("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum")
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    insts = num("smsp__inst_executed.sum")
    dur_ms = num("gpu__time_duration.sum")
    warp_trees = n / 32.0 * T
    total_pipe = (lsu_per_sm + tma_bytes / sms / 128.0) * sms
    name = [v for h, v in zip(hdr, vals) if h == "Kernel Name"][0]
    regs = num("launch__registers_per_thread")
    out = {
        "from": os.path.basename(rep), "kernel": name, "tuples": n, "trees": T, "duration_ms": dur_ms,
        "sm_cycles_elapsed": cycles, "registers_per_thread": regs,
        "lsu_wavefronts_per_sm": lsu_per_sm, "lsu_wavefronts_shared_per_sm": lsu_shared_per_sm, "lsu_wavefronts_global_per_sm": lsu_global_per_sm,
        "tma_fill_bytes": tma_bytes, "tma_fill_wavefront_equiv_per_sm": tma_bytes / sms / 128.0,
        "pipe_frac_of_cycles": (lsu_per_sm + tma_bytes / sms / 128.0) / cycles,
        "pipe_frac_lsu_only": lsu_per_sm / cycles,
        "pipe_wavefronts_per_tuple": total_pipe / n,
        "per_warp_tree": {
            "shared_loads": shared_ld / warp_trees, "of_which_bank_conflicts": conflicts_ld / warp_trees,
            "shared_stores": shared_st / warp_trees, "global_loads_t_stage": glob_ld / warp_trees,
            "lsu_global_data_stage": lsu_global_per_sm * sms / warp_trees,
            "tma_ring_fill": tma_bytes / 128.0 / warp_trees, "total": total_pipe / warp_trees,
        },
        "instructions_per_warp_visit": insts / (warp_trees * 12.0),
        "issue_slots_pct": num("sm__inst_executed.sum.pct_of_peak_sustained_elapsed", False),
        "l2_sectors_pct": num("lts__t_sectors.sum.pct_of_peak_sustained_elapsed", False),
        "dram_bytes_per_tuple": dram / n,
        "dram_pct_of_peak": num("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", False),
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_pipe.json"), "w"), indent=1)
    json.dump({"dram_bytes_per_tuple": dram / n, "from": os.path.basename(rep), "tuples": n},
              open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
