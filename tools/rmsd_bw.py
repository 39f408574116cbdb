"""Development: bandwidth of the stand-alone RMSD kernel on a resident screen (config 5 shape: ligands of 20..80 atoms).
   python tools/rmsd_bw.py [n_ligands]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
prot = synth.make_receptor(5000)
for n in [int(x) for x in sys.argv[1:]] or [12500, 100000]:
    rng = np.random.default_rng(5)
    sizes = rng.integers(20, 81, size=n)
    off, lig = synth.make_ligands(list(sizes[:64]), prot)
    # poses do not matter for the traffic: tile the 64 synthesised ligands
    reps = (n + 63) // 64
    sizes = np.tile(sizes[:64], reps)[:n]
    big = np.concatenate([lig[off[i % 64]:off[i % 64 + 1]] for i in range(n)])
    boff = np.concatenate([[0], np.cumsum(sizes)])
    rec = capi.Receptor(prot)
    scr = capi.Screen(rec, boff, big, 1)
    scr.run(1, True, 310.15, capi.default_params(), seed=1); scr.sync()
    best = 1e9
    for _ in range(6):
        r, ms = scr.rmsd(); best = min(best, ms)
    nbytes = 2 * 32 * int(boff[-1])
    fr, _ = scr.rmsd(); fu = scr.fused_rmsd()
    print(f"rmsd_chains_kernel: {n} pairs, {nbytes / 1e6:.1f} MB algorithmic, best of 6 {best * 1e3:.1f} us = {nbytes / best / 1e6:.0f} GB/s "
          f"= {nbytes / best / 1e6 / 6540.8 * 100:.1f}% of the measured HBM peak; same bits as the fused epilogue: {np.array_equal(fr, fu, equal_nan=True)}", flush=True)
    scr.close(); rec.close()
