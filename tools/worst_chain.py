"""Development: time the chain with the most collision retries alone on the GPU (critical-path estimate).
python tools/worst_chain.py <n_lig> <steps> [spec ...]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
n_lig, steps = int(sys.argv[1]), int(sys.argv[2])
specs = sys.argv[3:] or ["s0"]
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * n_lig, prot)
rec = capi.Receptor(prot)
prm = capi.default_params()
scr = capi.Screen(rec, off, lig, 1)
scr.run(steps, 1, 310.15, prm, seed=2); scr.sync()
_, st = scr.download(); scr.close()
r = st["retries"]
order = np.argsort(r)
print("retries percentiles 50/90/99/max:", np.percentile(r, [50, 90, 99, 100]), "all-chain ms", None)
for name, idx in (("worst", order[-1]), ("p99", order[int(0.99 * n_lig)]), ("median", order[n_lig // 2]), ("best", order[0])):
    o1 = np.array([0, 40]); l1 = lig[off[idx]:off[idx + 1]]
    for spec in specs:
        s1 = capi.Screen(rec, o1, l1, 1, first_chain_id=int(idx))
        if spec[0] == "s": s1.set_chain_slots(int(spec[1:]))
        else: s1.set_chain_threads(int(spec[1:]))
        s1.run(steps, 1, 310.15, prm, seed=2); s1.sync()
        s1.run(steps, 1, 310.15, prm, seed=2); s1.sync()
        ms = s1.last_kernel_ms(); _, s = s1.download(); s1.close()
        att = steps + int(s["retries"][0])
        print(f"{name:7s} chain {idx:5d} {spec:5s}: retries={int(s['retries'][0]):7d} (batch run said {int(r[idx])}) {ms:9.2f} ms  => {ms * 1e3 / steps:7.2f} us/step, "
              f"{ms * 1e3 / att:6.3f} us/attempt-equivalent", flush=True)
