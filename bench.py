#!/usr/bin/env python
"""bench.py — k-mers tested/sec through count -> matrix -> chi2 -> filter (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] — 250 synthetic 4.3 Mbp C. difficile-shaped
FASTA assemblies, k=16, binary phenotype, weighted chi-square + p<0.05 (Bonferroni) filter.
One step = one full pass of the hot path over all 250 samples.

  value : U / step time with the FASTA text already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API with HOST (pinned) buffers: H2D of all text and D2H
          of the survivors inside the timed region
  N > 1 : strong scaling — the k-mer space is range-sharded over the ranks (phenotypeseeker_b200/dist.py)
  --impl reference : the reference's own CPU implementation (shipped GenomeTester4 binaries +
          restated modeling.py loop, oracle/ref_pipeline.py) on a bounded sample, all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOAD = "BASELINE.json configs[1]: 250 synthetic 4.3 Mbp assemblies, k=16, binary phenotype, GSC-like weights, chi2 + p<0.05 Bonferroni"
METRIC = "k-mers tested/sec (count+matrix+chi2+filter)"
UNIT = "k-mers/s"
K = 16
PVALUE = 0.05


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(args):
    from phenotypeseeker_b200 import synth
    t0 = time.time()
    ds = synth.config(1, n_samples=args.samples, genome_len=args.genome_len)
    return ds, time.time() - t0


# ---------------------------------------------------------------------------------------
def cpu_reference_run(ds, n_samples, genome_len, threads, steps=1, warmup=0):
    """The reference CPU path on a bounded sample: all N samples, genomes cut to genome_len."""
    import tempfile
    import shutil
    from oracle import build as obuild, ref_pipeline
    from phenotypeseeker_b200 import synth
    obuild.build_all()
    sub = synth.config(1, n_samples=n_samples, genome_len=genome_len)
    td = tempfile.mkdtemp(prefix="psbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        _, paths = sub.write(td)
        col = [None if np.isnan(v) else int(v) for v in sub.pheno[:, 0]]
        times, U, kind = [], 0, "reference"
        for it in range(warmup + steps):
            t0 = time.time()
            r = ref_pipeline.run(paths, sub.names, K, [col], True, list(sub.weights), 2, n_samples - 2, PVALUE,
                                 False, threads=threads)
            dt = time.time() - t0
            U = r["U"]
            if it >= warmup:
                times.append(dt)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    sec = float(np.mean(times))
    return {"value": U / sec, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"all {n_samples} samples of the workload with genomes cut to {genome_len} bp "
                      f"(U={U}); shipped GenomeTester4 binaries (oracle/_ref/bin) for stages 1-2 + restated "
                      f"modeling.py:677-858 Python loop for stage 3, {threads} processes, scratch on /dev/shm; "
                      f"{sec:.1f} s per run"}, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    base, sec = cpu_reference_run(None, args.samples, args.ref_genome_len, threads, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32 k-mers / f64 statistics", "data": "synthetic",
            "config": {"workload": WORKLOAD, "k": K, "n_samples": args.samples},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from phenotypeseeker_b200.pipeline import KmerAssociation
    from phenotypeseeker_b200 import dist as psdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=device)
    ds, t_gen = make_workload(args)
    N = ds.n_samples
    mine = list(psdist.sample_block(rank, world, N))
    # inputs: pinned host copies (e2e) and device-resident copies (value) of this rank's samples
    tot = sum(len(ds.files[s]) + 64 for s in mine)
    host = torch.empty(tot, dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    spans, off = {}, 0
    for s in mine:
        b = np.frombuffer(ds.files[s], dtype=np.uint8)
        hv[off:off + len(b)] = b
        spans[s] = (off, len(b))
        off += (len(b) + 63) // 64 * 64
    dev = host.to(device, non_blocking=False)
    host_bufs = {s: hv[o:o + n] for s, (o, n) in spans.items()}
    dev_bufs = {s: (dev.data_ptr() + o, n) for s, (o, n) in spans.items()}
    h2d_bytes = sum(n for _, n in spans.values())
    pheno = ds.pheno[:, :1]
    ka = KmerAssociation(device=local)
    stream = torch.cuda.ExternalStream(ka.ctx.stream(), device=device)
    kw = dict(min_samples=2, max_samples=N - 2, pvalue_cutoff=PVALUE, omit_b=False)

    def step(bufs):
        U, res, info = psdist.run_sharded(ka, bufs, N, K, pheno, True, ds.weights, rank, world, device,
                                          route=args.route, **kw)
        return U, res, info

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize(device)

    def timed(bufs, steps, warmup, profile=False):
        U = 0
        res = None
        for _ in range(warmup):
            U, res, info = step(bufs)
        if profile:
            ka.ctx.profile_reset()
            ka.ctx.profile(True)
        l0 = ka.ctx.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            U, res, info = step(bufs)
        e1.record(stream)
        barrier()
        wall = time.time() - t0
        ms = e0.elapsed_time(e1)
        if profile:
            ka.ctx.profile(False)
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item()) / steps, U, res, info, (ka.ctx.launch_count() - l0) // steps, wall / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, U, res, info, launches, _ = timed(dev_bufs, args.steps, args.warmup, profile=True)
    prof = ka.ctx.profile_table()
    ms_e2e, U2, res2, _, _, wall_e2e = timed(host_bufs, args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    assert U == U2
    n_surv = len(res[0].kmer) if res is not None else 0
    d2h_bytes = n_surv * (8 * 6 + 4 * 2 + ka.ctx.row_words() * 4) + 64
    if world > 1:
        tdist.barrier()
    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return
    peak, peak_src = peaks()
    # dominant kernel = whichever kernel name took the most time (the partition passes `part_pass` on config 2)
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom_name = max(prof, key=lambda k: prof[k]["ms"])
    dom = prof[dom_name]
    per_launch_ms = dom["ms"] / max(dom["launches"], 1)
    per_launch_bytes = dom["alg_bytes"] / max(dom["launches"], 1)
    achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("kernel") == dom_name and world == 1 and tj.get("n_samples") == N:
            traffic = tj.get("dram_bytes_per_launch")
    # whole-step algorithmic bytes by SURVEY.md 8d: B1 = sum(L/4 + 4 D_s), B2 = sum(4 D_s) + U (4 + R),
    # B3 = U R + survivors (4 + 24 + R); D_s ~ positions (assemblies: nearly every k-mer is distinct)
    n_rec = sum(v["alg_bytes"] for k_, v in prof.items() if k_ == "extract_direct") / (args.steps * (3.0 / 8 + 8)) \
        if "extract_direct" in prof else 0.0
    R = 4 * ((N + 31) // 32)
    step_bytes = (n_rec / 4 + 4 * n_rec) + (4 * n_rec + U * (4 + R)) + (U * R + n_surv * (28 + R))
    line = {
        "metric": METRIC, "value": U / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32 k-mers / f64 statistics", "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": K, "n_samples": N, "genome_len": ds.meta["genome_len"],
                   "union_kmers": U, "survivors": n_surv, "input_bytes": int(ds.total_bytes()),
                   "l2": "inputs (1.1 GB text, 6.4 GB of k-mer pairs) are far larger than the 126 MB L2",
                   "parallelism": f"kmer-range-shard x{world}"},
        "e2e": {"value": U / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "launches_per_step": dom["launches"] // args.steps,
                     "ms_per_launch": per_launch_ms, "alg_bytes_per_launch": per_launch_bytes,
                     "share_of_kernel_time": dom["ms"] / tot_ms,
                     "whole_step_alg_bytes": step_bytes if n_rec else None,
                     "whole_step_frac": (step_bytes / (ms_dev * 1e-3) / 1e9 / peak) if n_rec else None},
        "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] // args.steps,
                        "alg_GBps": (v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
        "clocks": clocks,
        "gen_seconds": t_gen,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            base, _ = cpu_reference_run(ds, N, args.ref_genome_len, os.cpu_count() or 1)
            line["cpu_baseline"] = base
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {type(e).__name__}: {e}"}
    print(json.dumps(line))
    if world > 1:
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=250)
    ap.add_argument("--genome-len", type=int, default=4_300_000)
    ap.add_argument("--ref-genome-len", type=int, default=60_000,
                    help="genome length of the bounded sample the CPU reference is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--route", default="auto", choices=["auto", "pages", "streams"],
                    help="multi-GPU exchange route (phenotypeseeker_b200/dist.py)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
