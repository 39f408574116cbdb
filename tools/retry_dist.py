import sys
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
for steps in (2000, 5000, 10000):
    scr = capi.Screen(rec, off, lig, 1)
    scr.run(steps, True, 310.15, capi.default_params(), seed=2); scr.sync()
    ms = scr.last_kernel_ms()
    poses, st = scr.download()
    r = np.sort(st["retries"])
    print(f"steps={steps} ms={ms:.1f} retries: sum={r.sum()} median={np.median(r)} p90={r[int(.9*len(r))]} p99={r[int(.99*len(r))]} max={r[-1]} top5={r[-5:]} status={set(st['status'].tolist())}")
    # ideal time if perfectly balanced ~ proportional to steps + c*retries
    d = np.linalg.norm(poses[:, :3].reshape(1024, 40, 3) - prot[:, :3].mean(0), axis=2)
    print("   max |r - receptor centre| per chain: median", np.median(d.max(1)), "max", d.max())
    scr.close()
