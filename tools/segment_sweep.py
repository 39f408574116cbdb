"""Development: kernel time of one rank's shard of config 3 (4096 ligands x 8 replicas x 10 000 steps over N GPUs) as a function of
the chain-segment length: python tools/segment_sweep.py <n_gpus_emulated> [seg ...]   (seg: -1 automatic, 0 whole chains)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
segs = [int(x) for x in sys.argv[2:]] or [0, -1, 256, 1024]
steps = 10000
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * (4096 // n), prot)
rec = capi.Receptor(prot)
scr = capi.Screen(rec, off, lig, 8)
import os, signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
if os.environ.get('SLOTS'): scr.set_chain_slots(int(os.environ['SLOTS']))
prm = capi.default_params()
scr.run(500, 1, 310.15, prm, seed=1); scr.sync()
ref = None
for seg in segs:
    scr.set_segment_steps(seg)
    for seed in (200, 201):
        scr.run(steps, 1, 310.15, prm, seed=seed); scr.sync()
        ms = scr.last_kernel_ms()
        out, st = scr.download()
        if seed == 200:
            if ref is None: ref = (out, st)
            same = np.array_equal(out, ref[0]) and np.array_equal(st, ref[1])
        tot = float(st["steps"].sum()); pairs = tot * 5000 * 40
        print(f"shard 1/{n}: chains={len(st)} slots={scr.chain_slots()} seg={scr.segment_steps():5d} seed={seed}: {ms:9.2f} ms  {tot / ms * 1e3:.4e} steps/s  "
              f"{12 * pairs / ms * 1e3 / 74.45e12 * 100:5.2f}% FP32  same_bits_as_first={same}", flush=True)
