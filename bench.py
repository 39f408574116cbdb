#!/usr/bin/env python
"""bench.py -- Metropolis Monte-Carlo hot path: MC steps/s and atom-pair evaluations/s on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 2|3|4|5] [--chain-steps S]

One bench "step" = one pass of the hot path over one batch of synthetic input: every ligand chain of
the workload runs `chain_steps` (default 10 000) Metropolis iterations (metropolis.go:119-134) against
the synthetic 5 000-atom receptor:

  config 3 (default; BASELINE.json configs[2], the configuration north_star's 1000x / near-linear targets are quoted on):
           4096 ligands x 8 replica chains (numProcs = 8, metropolis.go:94-111) x 40 atoms = 32 768 chains, 3.3e8 MC steps
           per bench step, sharded over the N ranks by contiguous ligand blocks: STRONG scaling, no data-path collective.
  config 2: 1024 ligands x 40 atoms x 1 chain per GPU.  With N > 1 every rank runs its own 1024 ligands
           (ligand indices rank*1024 ...): weak scaling.
  config 4: large receptor, 50 000 atoms x 1024 ligands x 40 atoms per GPU (weak), with the cutoff extension
           (shifted Coulomb, 12 A, cell list; NO reference row -- the reference has no cutoff).  The line reports nominal
           pairs (nP*nL per step) and the pairs the cell list actually visited; --cutoff 0 runs the same receptor brute
           force with the reference's untruncated sum.
  config 5: virtual-screen sweep, 100 000 ligands of 20..80 atoms (1 chain each), sharded over the N ranks (strong),
           with the batched RMSD validation (CompareRMSD of every chain) device-resident: the line gains an "rmsd"
           object with the RMSD kernel's GB/s against the measured HBM peak.

`value`  = MC steps/s, whole job, chains resident in HBM (ddd_screen_*), max over ranks of the device time.
`e2e`    = the same metric through the reference-facing call (ddd_minimize_batch = SimulateMultipleLigands-
           Parallel, main.go:144) with pinned HOST buffers: H2D of the start poses and D2H of poses/energies
           inside the timed region; same seeds as the device-resident steps.
`roofline` = 12 flop/pair x pair evaluations / CUDA-event time of the chain kernel (measured on the library's
           own launch stream) against the FP32 FMA-pipe peak SMs*128*2*f_max (f_max = sm_max_mhz from
           MEASURED_PEAKS.json); the binding pipe is MUFU.RSQ (16/clk/SM = 75 % of that peak).
`cpu_baseline` = the C float64 restatement of the reference (oracle/, "port": NOT Go -- `go version` is probed on every
           run and fails on this image) timed on this box's host cores on a bounded sample of the same workload.  Its
           `value` is the build that squares through the restated Go math.Pow (what metropolis.go:253 executes;
           bit-identical results); the plain x*x build (faster than Go can be) is reported beside it, and the line carries
           `vs_cpu_gopow` and `vs_cpu_xx` (e2e / CPU) side by side.
--impl reference times that restatement alone (rank 0 only), same metric/config/units.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NP, NL, TEMPERATURE = 5000, 40, 310.15
FLOP_PER_PAIR = 12          # 3 sub, 3 mul + 2 add, rsqrt, clamp, mul + add
METRIC, UNIT = "mc_steps_per_sec", "MC steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5])
    ap.add_argument("--cutoff", type=float, default=-1.0, help="config 4: cutoff radius in A (default 12; 0 = brute force, reference sum)")
    ap.add_argument("--chain-steps", type=int, default=0, help="MC steps per chain (default 10000)")
    ap.add_argument("--ligands", type=int, default=0, help="override ligands per GPU (config 2, 4) / total (config 3, 5)")
    ap.add_argument("--chain-threads", type=int, default=0, help="one-CTA-per-chain engine with this CTA width (0 = SM-persistent engine)")
    ap.add_argument("--chain-slots", type=int, default=0, help="SM-persistent engine: chains resident per SM (0 = auto)")
    ap.add_argument("--e2e-calls", type=int, default=0, help="timed ddd_minimize_batch calls of the e2e leg (default: as many as fit ~60 s, at least 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "MEASURED_PEAKS.json"
    return {"sm_max_mhz": 1965.0, "hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def go_probe():
    """every run states whether the reference's own toolchain exists on this box."""
    exe = shutil.which("go")
    if not exe:
        return "absent (`go version`: command not found) -> CPU arm = C restatement of metropolis.go"
    try:
        out = subprocess.run([exe, "version"], capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:          # noqa: BLE001
        out = f"present but failed: {e}"
    return f"{out} (the overlay in go/ is not wired into this bench; CPU arm = C restatement)"


def workload(config, rank, world, n_override):
    """(first global ligand index, ligands of this rank, replicas, description)"""
    if config in (2, 4):
        n = n_override or 1024
        return rank * n, n, 1, f"synthetic {NP}-atom receptor x {n} ligands x {NL} atoms per GPU, 1 chain each"
    if config == 5:
        total = n_override or 100000
        lo, hi = rank * total // world, (rank + 1) * total // world
        if world > 1:
            # ligands of 20..80 atoms: the ranks' blocks are the library's own cost-balanced contiguous shards (ddd_shard_plan:
            # balanced by sum(nL + 1)), as a multi-device ddd_init would cut them -- an equal COUNT is up to 0.8 % off balance
            import ddd_loader
            m = ddd_loader.load()
            sizes = m.synth.sweep_sizes(total)
            plan = m.capi.shard_plan(np.concatenate([[0], np.cumsum(sizes)]), 1, world)
            lo, hi = int(plan[rank]), int(plan[rank + 1])
        return lo, hi - lo, 1, f"virtual-screen sweep: synthetic {NP}-atom receptor x {total} ligands of 20..80 atoms, 1 chain each, sharded; batched RMSD resident"
    total = n_override or 4096
    lo, hi = rank * total // world, (rank + 1) * total // world
    return lo, hi - lo, 8, f"synthetic {NP}-atom receptor x {total} ligands x {NL} atoms x 8 replica chains, sharded"


def config_dict(args, desc, n_lig, replicas, world, **extra):
    d = {"workload": desc, "baseline_config": args.config, "receptor_atoms": NP,
         "ligand_atoms": NL if args.config != 5 else "20..80", "ligands_per_gpu": n_lig, "replicas": replicas,
         "chain_steps": args.chain_steps, "rotate": True, "temperature": TEMPERATURE, "move": "REFERENCE",
         "parallelism": f"{world} x independent chain shards, no collective"}
    d.update(extra)
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        mhz, mx, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                mhz.append(float(f[1])); mx = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(mhz)}


def cpu_oracle_rate(prot, off, lig, seconds, variant="", cutoff=0.0):
    """Time the C restatement of the reference on all host cores; returns (steps/s, cores, sample text, pairs/s)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from _oracle import Oracle            # the one place bench.py executes oracle/: the CPU baseline
    orc = Oracle(variant)
    cores = os.cpu_count() or 1
    n_lig = min(len(off) - 1, max(cores, 8) * 2)
    o, x = off[:n_lig + 1], lig[:off[n_lig]]
    prm = orc.params(cutoff=cutoff)
    cal = 20 if len(prot) <= 10000 else 3
    sec, st, pr = orc.bench_chains(prot, o, x, 1, cal, True, TEMPERATURE, 1, cores, prm)          # calibration
    steps = max(cal, int(cal * seconds / max(sec, 1e-3)))
    sec, st, pr = orc.bench_chains(prot, o, x, 1, steps, True, TEMPERATURE, 1, cores, prm)
    return st / sec, cores, f"{n_lig} ligands x {steps} MC steps vs the {len(prot)}-atom receptor ({sec:.1f} s, {orc.build_info()})", pr / sec


class StdoutToStderr:
    """NCCL prints its version banner on fd 1 whenever NCCL_DEBUG is set: route fd 1 to fd 2 while NCCL and the bench
    run, so that NCCL_DEBUG=INFO can stay on (the driver reads the rank count from it) and stdout carries the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


T_START = time.perf_counter()


def main():
    args = parse_args()
    global NP
    if args.config == 4:
        NP = 50000
    if not args.chain_steps:
        args.chain_steps = 10000
    cutoff = (12.0 if args.cutoff < 0 else args.cutoff) if args.config == 4 else 0.0
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import ddd_loader
    ddd = ddd_loader.load()
    capi, synth = ddd.capi, ddd.synth
    peaks, peaks_src = measured_peaks()
    go = go_probe()

    first, n_lig, replicas, desc = workload(args.config, rank, world, args.ligands)
    prot = synth.make_receptor(NP)

    # ------------------------------------------------------------------ reference arm: CPU restatement only
    if args.impl == "reference":
        if rank != 0:
            return
        off, lig = synth.make_ligands([NL] * 64, prot)
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        rates = []
        for i in range(args.warmup + args.steps):
            r, cores, sample, pr = cpu_oracle_rate(prot, off, lig, per_step, "gopow", cutoff)
            if i >= args.warmup:
                rates.append((r, pr, sample))
        v = float(np.mean([r[0] for r in rates]))
        r_xx, _, sample_xx, _ = cpu_oracle_rate(prot, off, lig, min(per_step, 6.0), "", cutoff)
        _, n_lig0, _, _ = workload(args.config, 0, max(1, args.gpus), args.ligands)
        out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
               "scaling": "weak" if args.config in (2, 4) else "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "pair_evals_per_sec": float(np.mean([r[1] for r in rates])),
               "config": config_dict(args, desc, n_lig0, replicas, max(1, args.gpus)),
               "go_toolchain": go,
               "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": rates[-1][2],
                                "xx_variant_value": r_xx,
                                "note": "reference arm = C float64 restatement of metropolis.go on all host cores (no Go toolchain in this image), "
                                        "squares through the restated Go math.Pow as metropolis.go:253 does; each step is a bounded sample of the "
                                        "workload (a rate: the full workload is the sample scaled linearly in ligands and steps). "
                                        "xx_variant = the same restatement with x*x instead of math.Pow(x, 2): identical bits, faster than the Go program can be"},
               "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return

    # ------------------------------------------------------------------ our arm
    # synthetic ligands first (a process pool; CUDA is not initialised yet)
    if args.config == 5:
        total = args.ligands or 100000
        sizes = synth.sweep_sizes(total)[first:first + n_lig].tolist()
    else:
        sizes = [NL] * n_lig
    workers = max(1, (os.cpu_count() or 1) // max(1, world))
    off, lig = synth.make_ligands(sizes, prot, first_index=first, workers=workers)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    guard = StdoutToStderr()
    guard.__enter__()
    try:
        out = run_ours(args, capi, torch, dist, peaks, peaks_src, go, prot, off, lig, first, n_lig, replicas, desc, cutoff,
                       rank, local_rank, world)
    finally:
        guard.__exit__()
    if out is not None:
        print(json.dumps(out), flush=True)


def run_ours(args, capi, torch, dist, peaks, peaks_src, go, prot, off, lig, first, n_lig, replicas, desc, cutoff, rank, local_rank, world):
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.init([local_rank])
    info = capi.device_info(0)
    rec = capi.Receptor(prot)
    prm = capi.default_params(cutoff=cutoff)          # F32 pair sum, REFERENCE move, reference constants (+ config 4's cutoff)
    first_chain = first * replicas

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce(x, dist.ReduceOp.MAX)      # noqa: E731
    sum_over_ranks = lambda x: reduce(x, dist.ReduceOp.SUM)      # noqa: E731

    # ---- device-resident: value + roofline
    scr = capi.Screen(rec, off, lig, replicas, first_chain_id=first_chain)
    if args.chain_threads:
        scr.set_chain_threads(args.chain_threads)
    if args.chain_slots:
        scr.set_chain_slots(args.chain_slots)
    for w in range(args.warmup):
        scr.run(args.chain_steps, True, TEMPERATURE, prm, seed=100 + w)
        scr.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    kernel_ms, per_launch_ms = 0.0, []
    for k in range(args.steps):
        scr.run(args.chain_steps, True, TEMPERATURE, prm, seed=200 + k)
        scr.sync()                                    # device-side completion of the step (cudaStreamSynchronize)
        per_launch_ms.append(scr.last_kernel_ms())    # CUDA events on the library's launch stream
        kernel_ms += per_launch_ms[-1]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    _, st = scr.download()
    visited = float((scr.pair_units() * np.repeat(np.diff(off), replicas)).sum()) if cutoff > 0 else None   # pair terms the cell list kept
    rmsd_info = None
    if args.config == 5:                              # batched RMSD validation of the sweep (TestMethodRMSD shape), device-resident
        scr.rmsd()
        r_all, r_ms = scr.rmsd()
        r_bytes = 2.0 * 32.0 * float(off[-1]) * replicas + 8.0 * n_lig * replicas      # two 32-byte atoms per ligand atom + one result
        rmsd_info = {"pairs": int(n_lig * replicas), "kernel_ms": r_ms, "bytes": r_bytes, "finite": bool(np.all(np.isfinite(r_all)))}
    my_steps = float(st["steps"].sum()) * 1.0         # steps of the LAST bench step, all chains of this rank
    n_bailed = int(np.count_nonzero(st["status"] & capi.STATUS_RETRY_LIMIT))
    assert not np.any(st["status"] & ~capi.STATUS_RETRY_LIMIT), "chain status flags set"
    # a chain that needs more than max_retries (1024) collision retries in ONE step stops there (the reference would spin,
    # metropolis.go:173): counted below, and only the steps really executed enter the metric
    assert my_steps >= 0.99 * n_lig * replicas * args.chain_steps, "too many chains stopped early"
    width, slots, seg_steps = scr.chain_threads(), scr.chain_slots(), scr.segment_steps()
    scr.close()
    dev_s = max_over_ranks(kernel_ms * 1e-3)          # max over ranks of the device time of K steps
    wall_s = max_over_ranks(wall)
    total_steps = sum_over_ranks(my_steps) * args.steps
    my_pairs = float((st["steps"] * np.repeat(np.diff(off), replicas)).sum()) * NP      # sum_chains steps * nP * nL_chain
    total_pairs = sum_over_ranks(my_pairs) * args.steps
    visited_pairs = sum_over_ranks(visited) * args.steps if visited is not None else None
    n_bailed_all = int(sum_over_ranks(float(n_bailed)))
    value = total_steps / dev_s
    if rmsd_info is not None:
        rmsd_info["kernel_ms"] = max_over_ranks(rmsd_info["kernel_ms"])
        rmsd_info["bytes"] = sum_over_ranks(rmsd_info["bytes"]); rmsd_info["pairs"] = int(sum_over_ranks(rmsd_info["pairs"]))

    # ---- end to end through the reference-facing call with pinned host buffers: the SAME seeds as the device-resident
    #      steps (200, 201, ...), as many calls as fit the budget (all K when a launch is short)
    pin_lig = torch.empty(lig.shape, dtype=torch.float64).pin_memory()
    pin_lig.numpy()[:] = lig
    h_lig = pin_lig.numpy()
    launch_s = max_over_ranks(kernel_ms * 1e-3 / max(1, args.steps))
    # ~60 s of e2e calls, less when the device-resident steps already used most of the driver's per-run limit (870 s)
    e2e_budget = max(0.0, min(60.0, 780.0 - (time.perf_counter() - T_START) - 30.0))
    n_e2e = args.e2e_calls or int(max(1, min(args.steps, e2e_budget // max(launch_s, 1e-3))))
    iters = args.chain_steps * replicas               # SimulateEnergyMinimizationParallel runs iterations/numProcs per replica (:97)
    call = lambda seed: capi.minimize_batch(rec, off, h_lig, iters, True, TEMPERATURE, replicas, prm, seed=seed, first_ligand_index=first)   # noqa: E731
    if launch_s < 5.0:
        call(300)                                     # untimed warm call (skipped when one call costs many seconds: the library is warm)
    barrier()
    t0 = time.perf_counter()
    for k in range(n_e2e):
        out_pose, out_energy, picked = call(200 + k)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = sum_over_ranks(n_lig * replicas * args.chain_steps) * n_e2e / e2e_s * (my_steps / max(1.0, n_lig * replicas * args.chain_steps))
    h2d = int(sum_over_ranks(float(h_lig.nbytes + off.nbytes)))
    d2h = int(sum_over_ranks(float(h_lig.nbytes * replicas + n_lig * replicas * capi.STATS_DTYPE.itemsize)))
    dev_value_same_seeds = (total_steps / args.steps) * n_e2e / max_over_ranks(sum(per_launch_ms[:n_e2e]) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    fp32_peak = info["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12          # TFLOP/s per GPU
    work_pairs = visited_pairs if visited_pairs is not None else total_pairs           # what the kernel really evaluated
    achieved = FLOP_PER_PAIR * work_pairs / dev_s / 1e12 / world                       # per GPU
    cpu = None
    if not args.no_cpu_baseline:
        sub = (off[:min(n_lig, 64) + 1], lig[:off[min(n_lig, 64)]])
        r, cores, sample, pr = cpu_oracle_rate(prot, sub[0], sub[1], 7.0, "gopow", cutoff)
        rx, _, _, _ = cpu_oracle_rate(prot, sub[0], sub[1], 4.0, "", cutoff)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "pair_evals_per_sec": pr,
               "xx_variant_value": rx,
               "note": "C float64 restatement of metropolis.go, NOT Go (no Go toolchain here), one thread per chain on all host cores. "
                       "value: squares through the restated Go math.Pow (pow.go/frexp.go/ldexp.go, out of line), i.e. the arithmetic "
                       "metropolis.go:253 executes -- a cost MODEL of the Go program; xx_variant: x*x (same bits, faster than Go can be)"}

    # dram traffic of one launch (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum), kept per config in profiles/
    traffic, traffic_src = None, "not captured for this config"
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        t = json.loads(tfile.read_text()).get(str(args.config))
        if t and world == 1:
            traffic, traffic_src = t["bytes_per_launch"], t["source"]

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "weak" if args.config in (2, 4) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        **({"rmsd": {"pairs": rmsd_info["pairs"], "kernel_ms": rmsd_info["kernel_ms"], "algorithmic_bytes": rmsd_info["bytes"],
                     "achieved_gbs": rmsd_info["bytes"] / (rmsd_info["kernel_ms"] * 1e-3) / 1e9 / world, "peak_gbs": peaks.get("hbm_gbs"),
                     "frac": rmsd_info["bytes"] / (rmsd_info["kernel_ms"] * 1e-3) / 1e9 / world / peaks.get("hbm_gbs", 6540.8),
                     "bound": "hbm", "note": "CalculateRMSD of every chain (final vs start pose), Go []Atom layout (32 B/atom), device-resident"}}
           if rmsd_info else {}),
        "pair_evals_per_sec": total_pairs / dev_s,
        **({"cutoff": {"radius_A": cutoff, "nominal_pair_evals_per_sec": total_pairs / dev_s, "visited_pair_evals_per_sec": visited_pairs / dev_s,
                       "visited_fraction": visited_pairs / total_pairs,
                       "note": "EXTENSION without a reference row: shifted-Coulomb cutoff, cell list = Morton-ordered 32-atom units with bounding boxes; "
                               "roofline.achieved counts the visited pairs only; the CPU baseline runs the same truncated sum brute force"}}
           if visited_pairs is not None else {}),
        "config": config_dict(args, desc, n_lig, replicas, world,
                              pair_precision="fp32 terms, fp64 accumulation", chain_cta_threads=width, chains_per_sm=slots,
                              chains_stopped_at_retry_bound=n_bailed_all, chain_segment_steps=seg_steps,
                              engine="SM-persistent (one CTA per SM, speculative proposals)" if slots else "one CTA per chain",
                              l2=f"receptor ({NP * 16 // 1024} KB fp32) and poses are L1/L2-resident by design; the kernel is FP32/MUFU-bound, HBM traffic ~0, "
                                 "so no L2 flush applies"),
        "wall_ms_per_step": wall_s * 1e3 / args.steps,
        "launch_ms": {"min": min(per_launch_ms), "max": max(per_launch_ms), "spread": (max(per_launch_ms) - min(per_launch_ms)) / (kernel_ms / args.steps),
                      "note": "rank 0's CUDA-event time of every timed launch (seeds 200..): the seed-to-seed spread is the collision-retry tail"},
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": f"SMs*128 lanes*2*sm_max_mhz ({peaks['sm_max_mhz']:.0f} MHz from {peaks_src})",
                     "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": work_pairs / args.steps / world,
                     "kernel_ms_per_launch": dev_s * 1e3 / args.steps,
                     "mufu_bound_frac_of_peak": 0.75, "frac_of_mufu_bound": achieved / fp32_peak / 0.75},
        "cpu_baseline": cpu,
        "go_toolchain": go,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "ddd_minimize_batch (SimulateMultipleLigandsParallel, main.go:144) with pinned host buffers", "steps": n_e2e,
                "seeds": f"200..{199 + n_e2e} (the first {n_e2e} of the device-resident steps)",
                "device_resident_value_same_seeds": dev_value_same_seeds},
        **({"vs_cpu_gopow": e2e_value / cpu["value"], "vs_cpu_xx": e2e_value / cpu["xx_variant_value"], "host_cores": cpu["cores"]} if cpu else {}),
        "gpu_launches": args.steps * world,
        "clocks": clocks,
        "device": info["name"],
    }
    if world > 1:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
