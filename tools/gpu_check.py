"""Development check on a B200: parity of every C-ABI entry point against the CPU oracle on small
cases, then a timing of the chain kernel on the config-2 shape.  Not a test, not the bench."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ddd_loader  # noqa: E402
from _oracle import Oracle  # noqa: E402

m = ddd_loader.load()
capi, synth = m.capi, m.synth
orc = Oracle()


def main():
    n = capi.init([0])
    print("devices", n, capi.device_info(0), capi.version())
    print("chain kernel F32", capi.chain_kernel_info(capi.PRECISION_F32))
    print("chain kernel F64", capi.chain_kernel_info(capi.PRECISION_F64))
    prot = synth.make_receptor(5000)
    off, lig = synth.make_ligands([40] * 16 + [20, 33, 57, 80, 1, 2, 3], prot)
    rec = capi.Receptor(prot)

    # ---- energy parity
    eo, ao = orc.energy_batch(prot, off, lig, with_abs=True)
    e64, a64 = capi.energy_batch(rec, off, lig, capi.default_params(precision=capi.PRECISION_F64), with_abs=True)
    e32 = capi.energy_batch(rec, off, lig, capi.default_params(precision=capi.PRECISION_F32))
    print("energy f64: max |d|/abs_sum", np.nanmax(np.abs(e64 - eo) / ao), " max rel", np.nanmax(np.abs(e64 - eo) / np.abs(eo)),
          " abs_sum rel", np.nanmax(np.abs(a64 - ao) / ao))
    print("energy f32: max |d|/abs_sum", np.nanmax(np.abs(e32 - eo) / ao), " max rel", np.nanmax(np.abs(e32 - eo) / np.abs(eo)),
          " cond max", np.nanmax(ao / np.abs(eo)))

    # ---- replay parity, F64
    for rotate in (False, True):
        steps = 300
        p64 = capi.default_params(precision=capi.PRECISION_F64)
        poses, stats, trace = capi.run_chains(rec, off, lig, 2, steps, rotate, 310.15, p64, seed=7, want_trace=True)
        nbad = 0; maxpose = 0.0; maxe = 0.0
        for c in range((len(off) - 1) * 2):
            li, r = divmod(c, 2)
            pose_o, st_o, tr_o, _ = orc.chain(prot, lig[off[li]:off[li + 1]], steps, rotate, 310.15,
                                              orc.philox_stream(7, c), want_trace=True)
            same = np.array_equal(tr_o, trace[c])
            g = capi.chain_pose(off, 2, poses, li, r)
            if not same:
                nbad += 1
                print("  chain", c, "trace differs at", int(np.argmax(tr_o != trace[c])))
            else:
                maxpose = max(maxpose, float(np.nanmax(np.abs(g - pose_o))) if len(g) else 0.0)
                maxe = max(maxe, abs(stats[c]["final_energy"] - st_o["final_energy"]) / max(1.0, abs(st_o["final_energy"])))
            assert stats[c]["steps"] == st_o["steps"], (stats[c], st_o)
            if same:
                assert stats[c]["accepted"] == st_o["accepted"] and stats[c]["draws"] == st_o["draws"] and stats[c]["retries"] == st_o["retries"], (c, stats[c], st_o)
        print(f"replay F64 rotate={rotate}: chains with different accept trace {nbad}/{(len(off) - 1) * 2}, max pose diff {maxpose:.3e}, max rel final E diff {maxe:.3e}, mean acc {stats['accepted'].mean() / steps:.3f}")

    # ---- F32 statistics
    for rotate in (False, True):
        steps = 300
        poses, stats, _ = capi.run_chains(rec, off, lig, 2, steps, rotate, 310.15, capi.default_params(), seed=7)
        acc_o = []
        for c in range((len(off) - 1) * 2):
            li = c // 2
            _, st_o = orc.chain(prot, lig[off[li]:off[li + 1]], steps, rotate, 310.15, orc.philox_stream(7, c))
            acc_o.append(st_o["accepted"])
        print(f"F32 rotate={rotate}: mean acc gpu {stats['accepted'].mean():.2f} oracle {np.mean(acc_o):.2f}; status {set(stats['status'].tolist())}; "
              f"|chain_energy-final_energy|/|E| max {np.nanmax(np.abs(stats['chain_energy'] - stats['final_energy']) / np.abs(stats['final_energy'])):.2e}")

    # ---- rigid extension
    prm = capi.default_params(precision=capi.PRECISION_F64, move_mode=capi.MOVE_RIGID)
    poses, stats, trace = capi.run_chains(rec, off[:5], lig[:off[4]], 1, 200, True, 310.15, prm, seed=3, want_trace=True)
    po = orc.params(move_mode=1)
    bad = 0
    for c in range(4):
        pose_o, st_o, tr_o, _ = orc.chain(prot, lig[off[c]:off[c + 1]], 200, True, 310.15, orc.philox_stream(3, c), po, want_trace=True)
        bad += int(not np.array_equal(tr_o, trace[c]))
    print("rigid F64 replay: differing traces", bad, "/4")

    # ---- minimize
    out, en, picked, st = capi.minimize_batch(rec, off, lig, 800, True, 310.15, 4, capi.default_params(precision=capi.PRECISION_F64), seed=11, want_stats=True)
    rc, xo, eno, sto = orc.simulate_multiple_ligands(prot, off, lig, 800, True, 310.15, 4, 11)
    print("minimize F64: max rel energy diff", np.nanmax(np.abs(en - eno) / np.abs(eno)), "max pose diff", np.nanmax(np.abs(out - xo)), "picked", picked[:8])

    # ---- helpers
    d, li, pj = capi.closest_atom_batch(rec, off, lig)
    ok = all((orc.find_closest(lig[off[i]:off[i + 1]], prot) == (d[i], li[i], pj[i])) for i in range(len(off) - 1))
    print("closest exact match:", ok)
    far = lig.copy(); far[:, :3] += 40.0
    sh = capi.shift_closer_batch(rec, off, far, 5.0)
    sho = np.concatenate([orc.shift_closer(far[off[i]:off[i + 1]], prot, 5.0) for i in range(len(off) - 1)])
    print("shift max diff", np.nanmax(np.abs(sh - sho)))
    rz = capi.randomize_pose_batch(off, lig, seed=5)
    rzo = np.concatenate([orc.randomize_pose(lig[off[i]:off[i + 1]], orc.philox_stream(5, capi.STREAM_RANDOMIZE_BASE + i)) for i in range(len(off) - 1)])
    print("randomize max diff", np.nanmax(np.abs(rz - rzo)))
    r4 = capi.rmsd_batch(off, rz, lig, 4)
    r3 = capi.rmsd_batch(off, np.ascontiguousarray(rz[:, :3]), np.ascontiguousarray(lig[:, :3]), 3)
    ro = np.array([orc.rmsd(rz[off[i]:off[i + 1]], lig[off[i]:off[i + 1]]) for i in range(len(off) - 1)])
    print("rmsd max rel diff", np.nanmax(np.abs(r4 - ro) / ro), np.nanmax(np.abs(r3 - ro) / ro))

    # ---- timing, config-2 shape
    off2, lig2 = synth.make_ligands([40] * 1024, prot)
    scr = capi.Screen(rec, off2, lig2, 1)
    for prec, steps, widths in ((capi.PRECISION_F32, 1000, (0, 64, 128, 256)), (capi.PRECISION_F64, 20, (0,))):
        for w in widths:
            scr.set_chain_threads(w)
            prm = capi.default_params(precision=prec)
            scr.run(steps // 4, True, 310.15, prm, seed=1); scr.sync()
            scr.run(steps, True, 310.15, prm, seed=1); scr.sync()
            ms = scr.last_kernel_ms()
            _, st = scr.download()
            total_steps = int(st["steps"].sum())
            pairs = total_steps * 5000 * 40
            print(f"timing 1024 chains prec={prec} steps={steps} width={scr.chain_threads()}: kernel {ms:.2f} ms -> {total_steps / ms * 1e3:.3e} steps/s, "
                  f"{pairs / ms * 1e3:.3e} pairs/s, {12 * pairs / ms * 1e3 / 74.4e12 * 100:.1f}% of 74.4 TF; acc {st['accepted'].mean() / steps:.3f} retries {st['retries'].sum()}")
    scr.close()
    # 32768 chains (config 3 on 1 GPU), short
    off3, lig3 = synth.make_ligands([40] * 4096, prot)
    scr = capi.Screen(rec, off3, lig3, 8)
    for w in (0, 64, 128):
        scr.set_chain_threads(w)
        scr.run(50, True, 310.15, capi.default_params(), seed=1); scr.sync()
        scr.run(200, True, 310.15, capi.default_params(), seed=1); scr.sync()
        ms = scr.last_kernel_ms()
        print(f"timing 4096x8 chains x200 steps width={scr.chain_threads()}: kernel {ms:.2f} ms -> {32768 * 200 / ms * 1e3:.3e} steps/s, {32768 * 200 * 2e5 / ms * 1e3:.3e} pairs/s, {12 * 32768 * 200 * 2e5 / ms * 1e3 / 74.4e12 * 100:.1f}% of 74.4 TF")
    scr.close()


if __name__ == "__main__":
    main()
