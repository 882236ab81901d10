#!/usr/bin/env python3
"""bench.py — tuples/s of the tree-walk hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

ours: one "step" = one pass of the walk kernel (through the C ABI, libdte.so) over one batch of
  synthetic tuples already resident in HBM.  N=1 workload = BASELINE configs[2]
  ("1024 trees, depth 12, 256 features, 50M tuples, HBM-bound node walk").  N>1 = the data-sharded
  configuration (configs[4]): ensemble replicated, a fresh 50M-tuple shard per GPU, no collective,
  weak scaling; the ensemble-sharded configuration (configs[3], NCCL reduce of partial scores) is
  measured in the same run and reported under "ensemble_sharded".
reference: the reference has no CPU (or any software) implementation of this path — the RTL is the
  only definition — so the reference arm times the RTL-faithful oracle port (oracle/dte_oracle.c)
  on all host cores, on a bounded sample of the same workload.

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tuples/s at 1024 trees, depth 12, 256 feats"
T_TREES, DEPTH, FEATS, CLUSTERS = 1024, 12, 256, 8
N_FULL = 50_000_000
MISSING_PPM = 10000
SEED_TUPLES = 0x7091E5


def algorithmic_bytes_per_tuple(T, D, F):
    """SURVEY.md §8(d) byte model (B): every node visit reads threshold (4) + index (2) + feature (4),
    every tree one leaf (4), plus the tuple in and the score out."""
    return T * (D * 10 + 4) + 4 * F + 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi samples during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def load_oracle():
    from oracle import oracle as O
    return O


def build_ensemble(T, D, F):
    import ddt_b200 as ddt
    W, FI = ddt.layout.synth_ensemble(T, D, F)
    return ddt.layout.pack_streams(W, FI, D)


def oracle_throughput(T, D, F, K, S, seconds_target, threads, first_tuple=0, fixed_n=None):
    """tuples/s of the oracle port on `threads` host threads over a bounded sample."""
    import ddt_b200 as ddt
    O = load_oracle()
    L = ddt.layout
    wl, fl = build_ensemble(T, D, F)
    w_cls, f_cls = L.tree_cls(D)
    cfg = O.make_cfg(D, K, S, L.MISSING_DEFAULT, w_cls, f_cls, F // 4, T)
    if fixed_n is None:
        n0 = max(threads * 8, 256)
        x = L.synth_tuples(first_tuple, n0, F, seed=SEED_TUPLES, missing_ppm=MISSING_PPM)
        t = time.perf_counter(); O.scores(cfg, wl, fl, x, threads=threads); dt = time.perf_counter() - t
        n = int(min(max(n0, seconds_target * n0 / max(dt, 1e-6)), 2_000_000))
    else:
        n = fixed_n
    x = L.synth_tuples(first_tuple, n, F, seed=SEED_TUPLES, missing_ppm=MISSING_PPM)
    t = time.perf_counter(); s = O.scores(cfg, wl, fl, x, threads=threads); dt = time.perf_counter() - t
    return n / dt, n, dt, (cfg, wl, fl, x, s)


def run_reference(args):
    """--impl reference: the oracle port on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O = load_oracle()
    threads = O.max_threads()
    S = T_TREES // (8 * CLUSTERS)
    # size the per-step sample once (~8 s of CPU work), then time K steps after W warm-ups
    seconds = float(os.environ.get("DTE_BENCH_REF_SECONDS", "8"))       # CPU work per step (tests shrink it)
    _, n, _, pack = oracle_throughput(T_TREES, DEPTH, FEATS, CLUSTERS, S, seconds, threads)
    cfg, wl, fl, x, _ = pack
    for _ in range(min(args.warmup, 1)):
        O.scores(cfg, wl, fl, x, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.scores(cfg, wl, fl, x, threads=threads)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "tuples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg3: 1024 trees, D=12 comparison levels, 256 fp32 features; bounded sample of %d tuples per step" % n,
                   "trees": T_TREES, "depth_levels": DEPTH, "features": FEATS, "tuples_per_step": n},
        "cpu_baseline": {"value": val, "unit": "tuples/s", "cores": threads, "kind": "port",
                         "sample": "%d tuples of the cfg3 synthetic set per step, %d steps" % (n, args.steps)},
        "e2e": {"value": val, "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference ships no software implementation of this path (RTL only, not simulable here); "
                "this arm is the RTL-faithful oracle port on all host cores",
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)            # libraries that print to fd 1 (e.g. "NCCL version ...") now land on stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tuples", type=int, default=0, help="tuples per GPU per step (default: 50M, reduced only if a step would take > 30 s)")
    ap.add_argument("--e2e-tuples", type=int, default=4_000_000)
    ap.add_argument("--variant", type=int, default=0, help="force a kernel variant (dte_kernel_variant)")
    ap.add_argument("--no-ensemble-mode", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--trees", type=int, default=T_TREES)
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--features", type=int, default=FEATS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import ddt_b200 as ddt
    from ddt_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device; the engine has no CPU fallback"})
        return 1
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    T, D, F, K = args.trees, args.depth, args.features, CLUSTERS
    S = -(-T // (8 * K))
    L = ddt.layout
    wl, fl = build_ensemble(T, D, F)
    e = ddt.Engine(local)
    e.configure(T, D, 4 * F, clusters=K, missing_value=L.MISSING_DEFAULT, n_tuples=N_FULL)
    e.load_ensemble(wl, fl)
    if args.variant:
        e.set_kernel_variant(args.variant)
    st = torch.cuda.current_stream().cuda_stream

    # ---- size the step: calibrate on 1M tuples ----
    n_cal = 1 << 20
    d_cal = torch.empty((n_cal, F), dtype=torch.int32, device="cuda")
    s_cal = torch.empty(n_cal, dtype=torch.float32, device="cuda")
    e.synth_tuples_device(d_cal, 0, n_cal, F, SEED_TUPLES, MISSING_PPM, L.MISSING_DEFAULT, stream=st)
    e.infer_device(d_cal, n_cal, s_cal, None, stream=st)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); e.infer_device(d_cal, n_cal, s_cal, None, stream=st); ev1.record(); torch.cuda.synchronize()
    cal_rate = n_cal / (ev0.elapsed_time(ev1) * 1e-3)
    n = args.tuples or N_FULL
    free_b, _ = torch.cuda.mem_get_info()
    n = min(n, int((free_b - (8 << 30)) // (F * 4 + 8)))
    if not args.tuples and n / cal_rate > 30.0:
        n = max(1 << 22, int(cal_rate * 30.0))
    if dist is not None:                                   # every rank must time the same shard size
        tn = torch.tensor([n], dtype=torch.int64, device="cuda")
        dist.all_reduce(tn, op=dist.ReduceOp.MIN)
        n = int(tn.item())
    del d_cal, s_cal

    # ---- resident inputs: this rank's shard of the synthetic set (rank r owns tuples [r*n, (r+1)*n)) ----
    d_x = torch.empty((n, F), dtype=torch.int32, device="cuda")
    d_s = torch.empty(n, dtype=torch.float32, device="cuda")
    d_l = torch.empty(n, dtype=torch.uint8, device="cuda")
    e.synth_tuples_device(d_x, rank * n, n, F, SEED_TUPLES, MISSING_PPM, L.MISSING_DEFAULT, stream=st)
    torch.cuda.synchronize()

    def step():
        e.infer_device(d_x, n, d_s, d_l, stream=st)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = e.info()["kernel_launches"]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    launches = e.info()["kernel_launches"] - launches0
    if dist is not None:
        tms = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        total_ms = float(tms.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- spot parity inside the bench: a sample of the timed output against the oracle (rank 0) ----
    parity = None
    if rank == 0 and not args.no_cpu:
        O = load_oracle()
        idx = np.unique(np.concatenate([np.arange(64), np.arange(0, n, max(1, n // 192))]))[:256]
        xs = np.stack([L.synth_tuples(int(i), 1, F, seed=SEED_TUPLES, missing_ppm=MISSING_PPM)[0] for i in idx])
        w_cls, f_cls = L.tree_cls(D)
        want = O.scores(O.make_cfg(D, K, S, L.MISSING_DEFAULT, w_cls, f_cls, F // 4, T), wl, fl, xs, threads=O.max_threads())
        got = d_s[torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint32)
        lab = d_l[torch.from_numpy(idx).cuda()].cpu().numpy()
        parity = {"sampled": int(idx.size), "score_words_equal": int((got == want).sum()), "labels_equal": int((lab == O.labels(want)).sum())}

    # ---- e2e: host buffers through dte_infer_host (H2D + walk + D2H inside the timed region) ----
    n_e = min(args.e2e_tuples, n)
    h_x = torch.empty((n_e, F), dtype=torch.int32, pin_memory=True)
    h_x.copy_(d_x[:n_e])
    h_s = torch.empty(n_e, dtype=torch.float32, pin_memory=True)
    h_l = torch.empty(n_e, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        e.infer_host(h_x, out_scores=h_s, out_labels=h_l)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e.infer_host(h_x, out_scores=h_s, out_labels=h_l)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_dt = float(tt.item())
    e2e_val = world * n_e * e2e_steps / e2e_dt
    e2e_ok = bool(torch.equal(h_s.view(torch.int32), d_s[:n_e].cpu().view(torch.int32)))

    # ---- ensemble-sharded configuration (BASELINE configs[3]): 1024 trees per GPU, D=10, one NCCL reduce ----
    ens = None
    if world > 1 and not args.no_ensemble_mode:
        De, Te = 10, 1024 * world
        We, FIe = L.synth_ensemble(Te, De, F, seed=0xE5E)
        wle, fle = L.pack_streams(We, FIe, De)
        first, count = ddt.sharding.ensemble_chunk(Te, rank, world)
        Ke, Se = ddt.sharding.shard_geometry(Te, De, K, world)
        ee = ddt.Engine(local)
        ee.configure(count, De, 4 * F, clusters=Ke, missing_value=L.MISSING_DEFAULT)
        ee.softreg_write(205, (L.MISSING_DEFAULT) | (De << 32) | (Se << 36) | (Ke << 44))
        ee.load_ensemble(wle, fle, first_tree=first, num_local_trees=count)
        ne = min(n, 20_000_000)
        # every device sees every tuple (InputDistributor.sv:199-204): same seed/window on all ranks
        ee.synth_tuples_device(d_x, 0, ne, F, SEED_TUPLES, MISSING_PPM, L.MISSING_DEFAULT, stream=st)
        part = d_s[:ne]

        def estep():
            ee.infer_device(d_x, ne, part, None, stream=st)
            dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)          # ONE collective: ResultsCombiner ring -> NCCL reduce
            if rank == 0:
                ee.labels_device(part, ne, d_l, stream=st)

        for _ in range(2):
            estep()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            estep()
        b.record(); torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([a.elapsed_time(b) / 3], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # one GPU holding ALL trees on a slice of the same tuples -> time per tuple -> speed-up
        one_ms = None
        if rank == 0:
            e1 = ddt.Engine(local)
            K1, S1 = K, -(-Te // (8 * K))
            e1.configure(Te, De, 4 * F, clusters=K1, missing_value=L.MISSING_DEFAULT)
            e1.load_ensemble(wle, fle)
            n1 = min(ne, 4_000_000)
            tmp = torch.empty(n1, dtype=torch.float32, device="cuda")
            e1.infer_device(d_x, n1, tmp, None, stream=st); torch.cuda.synchronize()
            a.record(); e1.infer_device(d_x, n1, tmp, None, stream=st); b.record(); torch.cuda.synchronize()
            one_ms = a.elapsed_time(b) * (ne / n1)
            e1.close()
        # the same step with the combine FUSED into the walk epilogue (red.add over NVLink peer memory)
        fused_ms = None
        try:
            fc = ddt.sharding.FusedCombine(ee, dist, ne, dst=0)
            for _ in range(2):
                fc.step(d_x, ne, st)
            t0 = time.perf_counter()
            for _ in range(3):
                out = fc.step(d_x, ne, st)
                if rank == 0:
                    ee.labels_device(out, ne, d_l, stream=st)
            torch.cuda.synchronize()
            tf = torch.tensor([(time.perf_counter() - t0) / 3 * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            fused_ms = float(tf.item())
            fused_ok = None
            if rank == 0:
                estep()                                            # NCCL result into `part` for comparison
                torch.cuda.synchronize()
                fused_ok = bool(torch.allclose(out, part, rtol=1e-5, atol=1e-7))
            else:
                estep()
            fc.close()
        except Exception as ex:                                    # IPC not permitted in this container etc.
            fused_ms, fused_ok = None, "unavailable: %s" % ex
        ens = {"workload": "cfg4: %d trees split %d/GPU, D=10, 256 features, %d tuples on every GPU, one NCCL reduce(SUM) of fp32[%d] to rank 0"
                           % (Te, count, ne, ne),
               "ms_per_step": float(ms.item()), "tuples_per_s": ne / (float(ms.item()) * 1e-3),
               "one_gpu_all_trees_ms_extrapolated": one_ms,
               "speedup_vs_one_gpu": (one_ms / float(ms.item())) if one_ms else None,
               "fused_epilogue": {"what": "combine fused into the walk kernel (system-scope red.add into rank 0's IPC buffer over NVLink), no collective",
                                  "ms_per_step": fused_ms, "matches_nccl_within_1e-5": fused_ok,
                                  "speedup_vs_one_gpu": (one_ms / fused_ms) if (one_ms and fused_ms) else None}}
        ee.close()

    if rank == 0:
        bytes_b = algorithmic_bytes_per_tuple(T, D, F)
        peak, peak_src = measured_peaks()
        launch_ms = float(np.mean(per_launch_ms))
        achieved = (n * bytes_b / (launch_ms * 1e-3)) / 1e9
        traffic = None
        tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tj):
            try:
                tr = json.load(open(tj))
                traffic = tr.get("dram_bytes_per_tuple", 0) * n if tr.get("dram_bytes_per_tuple") else None
            except Exception:
                traffic = None
        info = e.info()
        cpu = None
        if not args.no_cpu:
            O = load_oracle()
            th = O.max_threads()
            v, ns, dt, _ = oracle_throughput(T, D, F, K, S, 12.0, th)
            v1, ns1, dt1, _ = oracle_throughput(T, D, F, K, S, 4.0, 1)
            # BASELINE configs[0]: 16 trees, depth 4, 32 features, 10k tuples, ONE host thread, median of 21 runs
            W1, FI1 = L.synth_ensemble(16, 4, 32)
            wl1, fl1 = L.pack_streams(W1, FI1, 4)
            x1 = L.synth_tuples(0, 10000, 32)
            w1c, f1c = L.tree_cls(4)
            c1 = O.make_cfg(4, 2, 1, L.MISSING_DEFAULT, w1c, f1c, 8, 16)
            t_cfg1 = []
            for _ in range(21):
                t0 = time.perf_counter(); O.scores(c1, wl1, fl1, x1, threads=1); t_cfg1.append(time.perf_counter() - t0)
            cfg1_rate = 10000 / float(np.median(t_cfg1))
            cpu = {"value": v, "unit": "tuples/s", "cores": th, "kind": "port",
                   "cfg1_single_thread": {"value": cfg1_rate, "unit": "tuples/s", "what": "BASELINE configs[0]: 16 trees, D=4, 32 features, 10k tuples, 1 thread, median of 21 runs"},
                   "sample": "%d tuples of the same synthetic set in %.1f s on %d threads (oracle/dte_oracle.c); single thread: %.0f tuples/s on %d tuples"
                             % (ns, dt, th, v1, ns1),
                   "single_thread_value": v1,
                   "reference_model_fpga": {"value": 150e6 * 8 * 8 / (D * T), "unit": "tuples/s",
                                            "what": "the reference's own analytical law f*Ncu*Npe/(depth*Ntrees) at 150 MHz, 8x8 PEs (profiler/profiler.cpp:97-102); modelled, Catapult v1.2"}}
        line = {
            "metric": METRIC, "value": value, "unit": "tuples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("cfg3" if world == 1 else "cfg5 (data-sharded, ensemble replicated, no collective)") +
                                   ": %d trees, D=%d comparison levels, %d fp32 features, %d tuples per GPU per step" % (T, D, F, n),
                       "trees": T, "depth_levels": D, "features": F, "tuples_per_gpu": n, "clusters": K, "trees_per_pu": S,
                       "missing_ppm": MISSING_PPM, "l2": "inputs (%.1f GB per GPU) far larger than L2; no flush needed" % (n * F * 4 / 1e9),
                       "kernel": E.KERNEL_NAMES.get(info["kernel_variant"], "?"), "tuples_per_cta": info["tuples_per_cta"],
                       "tune": os.environ.get("DTE_TUNE", "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_tuple": bytes_b,
                         "kernel": "dt_walk_tile" if info["kernel_variant"] != 1 else "dt_walk_generic",
                         "launch_ms": launch_ms, "tuples_per_launch": n,
                         "note": "byte model (B) of SURVEY 8d charges every node visit (10 B) and leaf (4 B) as HBM traffic; the "
                                 "engine serves them from shared memory / L2, so frac can exceed 1 — `traffic` is the DRAM bytes "
                                 "ncu measured for one launch (profiles/ncu_traffic.json), ~1049 B per tuple"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "tuples/s", "h2d_bytes_per_step": int(n_e * F * 4 * world),
                    "d2h_bytes_per_step": int(n_e * 5 * world), "tuples_per_step": int(n_e * world), "steps": e2e_steps,
                    "bit_equal_to_device_path": e2e_ok,
                    "api": "dte_infer_host (pinned host buffers; H2D, walk, D2H pipelined inside the call)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_spot_check": parity,
            "calibration_tuples_per_s": cal_rate,
        }
        if ens:
            line["ensemble_sharded"] = ens
        emit(line)
    e.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
