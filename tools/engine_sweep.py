"""Development sweep of the chain-kernel engines on a B200 (not a test, not the bench):
python tools/engine_sweep.py <n_lig> <replicas> <steps> <spec> [<spec> ...]   spec = s<K> (SM-persistent, K slots; s0 = auto) | t<W> (CTA width)"""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
n_lig, replicas, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
specs = sys.argv[4:] or ["s0"]
NP = int(os.environ.get("DDD_NP", "5000")); NL = int(os.environ.get("DDD_NL", "40"))
capi.init([0])
prot = synth.make_receptor(NP)
off, lig = synth.make_ligands([NL] * n_lig, prot)
rec = capi.Receptor(prot)
scr = capi.Screen(rec, off, lig, replicas)
prm = capi.default_params()
for spec in specs:
    if spec[0] == "s":
        scr.set_chain_threads(0); scr.set_chain_slots(int(spec[1:]))
    else:
        scr.set_chain_slots(0); scr.set_chain_threads(int(spec[1:]))
    scr.run(max(1, steps // 4), 1, 310.15, prm, seed=1); scr.sync()
    best = 1e30
    for rep in range(2):
        scr.run(steps, 1, 310.15, prm, seed=2); scr.sync()
        best = min(best, scr.last_kernel_ms())
    _, st = scr.download()
    tot = float(st["steps"].sum()); pairs = tot * NP * NL
    print(f"{spec:5s} chains={n_lig * replicas} steps={steps} threads={scr.chain_threads()} slots={scr.chain_slots()}: {best:9.2f} ms  "
          f"{tot / best * 1e3:.3e} steps/s  {12 * pairs / best * 1e3 / 74.45e12 * 100:5.1f}% FP32  acc={st['accepted'].mean() / steps:.4f} "
          f"retries={int(st['retries'].sum())} maxretries={int(st['retries'].max())} status={int(st['status'].max())}", flush=True)
