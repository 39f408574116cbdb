"""Development: kernel time of config 2 per seed (the retry tail is seed-dependent)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * 1024, prot)
rec = capi.Receptor(prot)
scr = capi.Screen(rec, off, lig, 1)
prm = capi.default_params()
scr.run(1000, 1, 310.15, prm, seed=1); scr.sync()
for seed in [int(x) for x in sys.argv[1:]] or [200, 201, 202, 203, 204, 300, 400, 401, 402]:
    scr.run(10000, 1, 310.15, prm, seed=seed); scr.sync()
    ms = scr.last_kernel_ms(); _, st = scr.download()
    t0 = time.perf_counter()
    capi.minimize_batch(rec, off, lig, 10000, True, 310.15, 1, prm, seed=seed)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"seed {seed}: kernel {ms:7.1f} ms  minimize_batch wall {wall:7.1f} ms  retries total {int(st['retries'].sum())} max {int(st['retries'].max())}", flush=True)
