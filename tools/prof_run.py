"""Profiling driver: one warm launch, then ONE launch of the requested shape (capture it with `ncu -k regex:mc_sm -s 1 -c 1`).
    python tools/prof_run.py <receptor atoms> <ligands> <replicas> <steps> [cutoff] [slots] [segment steps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
n_rec, n_lig, replicas, steps = (int(x) for x in sys.argv[1:5])
cutoff = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
slots = int(sys.argv[6]) if len(sys.argv) > 6 else 0
prot = synth.make_receptor(n_rec)
off, lig = synth.make_ligands([40] * n_lig, prot, workers=8)
capi.init([0])
rec = capi.Receptor(prot)
scr = capi.Screen(rec, off, lig, replicas)
if slots:
    scr.set_chain_slots(slots)
if len(sys.argv) > 7:
    scr.set_segment_steps(int(sys.argv[7]))
prm = capi.default_params(cutoff=cutoff)
scr.run(max(1, steps // 10), True, 310.15, prm, seed=1); scr.sync()
scr.run(steps, True, 310.15, prm, seed=200); scr.sync()
ms = scr.last_kernel_ms()
_, st = scr.download()
tot = float(st["steps"].sum())
print(f"receptor {n_rec} chains {n_lig * replicas} steps {steps} cutoff {cutoff}: {ms:.2f} ms, {tot / ms * 1e3:.3e} MC steps/s, slots {scr.chain_slots()} segment {scr.segment_steps()}")
