"""Development timing of the chain kernel: python tools/quick_time.py <n_lig> <replicas> <steps> [width] (DDD_MC_LIB selects a build)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
n_lig, replicas, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
width = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rotate = int(sys.argv[5]) if len(sys.argv) > 5 else 1
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * n_lig, prot)
rec = capi.Receptor(prot)
scr = capi.Screen(rec, off, lig, replicas)
scr.set_chain_threads(width)
prm = capi.default_params()
scr.run(max(1, steps // 4), rotate, 310.15, prm, seed=1); scr.sync()
scr.run(steps, rotate, 310.15, prm, seed=2); scr.sync()
ms = scr.last_kernel_ms()
_, st = scr.download()
tot = float(st["steps"].sum()); pairs = tot * 5000 * 40
print(f"{os.environ.get('DDD_MC_LIB', 'default'):40s} chains={n_lig * replicas} steps={steps} width={scr.chain_threads()} rot={rotate}: {ms:9.2f} ms  "
      f"{tot / ms * 1e3:.3e} steps/s  {12 * pairs / ms * 1e3 / 74.45e12 * 100:5.1f}% FP32  acc={st['accepted'].mean() / steps:.4f} retries={int(st['retries'].sum())} "
      f"info={capi.chain_kernel_info()}")
