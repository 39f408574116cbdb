"""Consumer side of go/harness/record_stream_test.go: replays a stream recorded from Go's math/rand on the GPU
(DDD_PRECISION_F64) and on the CPU oracle, and compares both with the pose the untouched Go reference produced.

    python tools/replay_go_record.py <dir with stream.f64 + chain.json> <223l_protein xyzq .npy or tests/golden/sample_mol2.npz>

Needs a machine with Go (to produce the two files) and a B200; neither exists in the build container, so this
script is the documented procedure rather than an executed test ("parity unpinned")."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ddd_loader  # noqa: E402
from _oracle import Oracle  # noqa: E402


def main():
    d = Path(sys.argv[1])
    z = np.load(sys.argv[2]) if len(sys.argv) > 2 else np.load(ROOT / "tests" / "golden" / "sample_mol2.npz")
    prot = z["223l_protein"]
    rec_stream = np.fromfile(d / "stream.f64", dtype="<f8")
    chain = json.loads((d / "chain.json").read_text())
    start, final_go = np.array(chain["start"]), np.array(chain["final"])
    orc = Oracle()
    pose_o, st_o = orc.chain(prot, start, chain["iterations"], chain["rotate"], chain["temperature"], orc.recorded_stream(rec_stream))
    print("oracle vs Go : max |dpose| =", np.abs(pose_o - final_go).max(), " final E rel diff =",
          abs(st_o["final_energy"] - chain["final_energy"]) / abs(chain["final_energy"]), " status", st_o["status"])
    capi = ddd_loader.load().capi
    capi.init([0])
    with capi.Receptor(prot) as rec:
        off = np.array([0, len(start)])
        pose_g, st_g, _ = capi.run_chains(rec, off, start, 1, chain["iterations"], chain["rotate"], chain["temperature"],
                                          capi.default_params(precision=capi.PRECISION_F64), recorded=rec_stream,
                                          recorded_offsets=np.array([0, len(rec_stream)]))
    print("GPU    vs Go : max |dpose| =", np.abs(pose_g - final_go).max(), " final E rel diff =",
          abs(st_g[0]["final_energy"] - chain["final_energy"]) / abs(chain["final_energy"]), " status", st_g[0]["status"])


if __name__ == "__main__":
    main()
