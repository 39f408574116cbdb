"""The PRODUCT's multi-GPU path: ONE process, ddd_init with every visible device, BASELINE config 3 (4096 ligands x 8 replica chains x
10 000 steps) through ddd_minimize_batch (= SimulateMultipleLigandsParallel, main.go:144) from pinned-size host buffers -- what a Go
program linked against libddd_mc.so gets without torchrun.  Prints one JSON line: wall-clock MC steps/s of the call (H2D, kernels on
all devices, replica pick, D2H) and the slowest device's kernel time.
    python tools/in_process_scaling.py [calls]"""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
import torch
n_dev = torch.cuda.device_count()
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * 4096, prot, workers=16)
capi.init(list(range(n_dev)))
rec = capi.Receptor(prot)
prm = capi.default_params()
capi.minimize_batch(rec, off, lig, 8 * 300, True, 310.15, 8, prm, seed=1)            # warm: contexts, pools, kernels on every device
t = []
for k in range(calls):
    t0 = time.perf_counter()
    out, en, picked = capi.minimize_batch(rec, off, lig, 8 * 10000, True, 310.15, 8, prm, seed=200 + k)
    t.append(time.perf_counter() - t0)
steps = 4096 * 8 * 10000
print(json.dumps({"path": "in-process (ddd_init with all devices), ddd_minimize_batch, config 3", "devices": n_dev, "calls": calls,
                  "wall_s_per_call": t, "mc_steps_per_sec": steps / (sum(t) / len(t)), "finite": bool(np.all(np.isfinite(en)))}))
rec.close(); capi.shutdown()
