"""How often do the production (fp32 pair terms) chains take a different accept/reject decision than the float64 restatement,
and when?  Same Philox streams on both sides; a chain's traces are compared step by step.

    python tools/f32_divergence.py > profiles/r02_f32_trace_divergence.json

An fp32 chain diverges at the first near-tie newE ~ curE that fp32 summation resolves the other way; from there
on the trajectories differ (different accepted pose), but both are valid runs of the same sampler."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ddd_loader
from _oracle import Oracle

m = ddd_loader.load(); capi, synth = m.capi, m.synth
orc = Oracle()
capi.init([0])
out = {}
for name, n_rec, sizes, steps in (("test shape: 1500-atom receptor x 48 ligands x 24 atoms", 1500, [24] * 48, 250),
                                  ("config 2 shape: 5000-atom receptor x 96 ligands x 40 atoms", 5000, [40] * 96, 400)):
    prot = synth.make_receptor(n_rec, seed=2 if n_rec == 1500 else synth.RECEPTOR_SEED)
    off, lig = synth.make_ligands(sizes, prot)
    rec = capi.Receptor(prot)
    for rotate in (False, True):
        poses, stats, trace = capi.run_chains(rec, off, lig, 1, steps, rotate, 310.15, capi.default_params(), seed=17, want_trace=True)
        first, acc_o = [], []
        for c in range(len(sizes)):
            _, st_o, tr_o, _ = orc.chain(prot, lig[off[c]:off[c + 1]], steps, rotate, 310.15, orc.philox_stream(17, c), want_trace=True)
            d = np.nonzero(tr_o != trace[c])[0]
            first.append(int(d[0]) if len(d) else -1)
            acc_o.append(int(st_o["accepted"]))
        first = np.array(first)
        div = first[first >= 0]
        out[f"{name}, {steps} steps, rotate={rotate}"] = {
            "chains": len(sizes), "identical_traces": int((first < 0).sum()), "identical_fraction": float((first < 0).mean()),
            "first_divergence_step_histogram": {f"{lo}-{lo + steps // 5 - 1}": int(((div >= lo) & (div < lo + steps // 5)).sum()) for lo in range(0, steps, steps // 5)},
            "mean_accepted_gpu_f32": float(stats["accepted"].mean()), "mean_accepted_oracle_f64": float(np.mean(acc_o)),
        }
    rec.close()
print(json.dumps(out, indent=1))
