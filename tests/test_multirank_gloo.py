"""N > 1 path on CPU: world_size-2 `gloo` processes shard a ligand batch exactly the way bench.py / the Go host
do (contiguous ligand blocks, global chain ids, no data-path collective), run their shard, and gather on rank 0.
The CPU stand-in for the per-rank compute is the oracle (this is a test); what is under test is the host logic:
shard boundaries, stream-id bases (first_ligand_index) and the gather order reproduce the single-rank result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ddd_loader
    from _oracle import Oracle
    import bench
    ddd = ddd_loader.load()
    orc = Oracle()
    prot = ddd.synth.make_receptor(300, seed=4)
    sizes = [9, 14, 6, 11, 20, 7, 13]
    off, lig = ddd.synth.make_ligands(sizes, prot)
    # rank boundaries: the library's shard plan at ligand granularity (replicas = 1)
    b = ddd.capi.shard_plan(off, 1, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    o = off[lo:hi + 1] - off[lo]
    x = lig[off[lo]:off[hi]]
    # each rank minimises its block with GLOBAL ligand indices -> the same Philox streams as a single-rank run
    en = np.zeros(hi - lo); out = x.copy()
    for i in range(hi - lo):
        rc, pose, st, _, _ = orc.minimize_parallel(prot, x[o[i]:o[i + 1]], 120, True, 310.15, 3, seed=5, lig_index=lo + i)
        assert rc == 0
        out[o[i]:o[i + 1]] = pose
        en[i] = orc.energy(prot, pose)
    # timing contract of bench.py: barrier, then MAX over ranks
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    # host-side gather of per-ligand energies and poses (no collective on the data path itself)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, en, out))
    if rank == 0:
        gathered.sort(key=lambda g: g[0])
        assert [g[0] for g in gathered][0] == 0 and gathered[-1][1] == len(sizes)
        assert all(gathered[k][1] == gathered[k + 1][0] for k in range(world - 1))           # contiguous, in order
        np.savez(out_path, energy=np.concatenate([g[2] for g in gathered]), pose=np.concatenate([g[3] for g in gathered]))
        # bench.py's own sharding helper agrees with itself across ranks
        spans = [bench.workload(3, r, world, 7)[:2] for r in range(world)]
        assert spans[0][0] == 0 and sum(s[1] for s in spans) == 7
        assert [bench.workload(2, r, world, 0)[0] for r in range(world)] == [0, 1024]
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_reproduces_single_rank(tmp_path, orc, ddd):
    out_path = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    z = np.load(out_path)
    prot = ddd.synth.make_receptor(300, seed=4)
    off, lig = ddd.synth.make_ligands([9, 14, 6, 11, 20, 7, 13], prot)
    rc, xo, eno, _ = orc.simulate_multiple_ligands(prot, off, lig, 120, True, 310.15, 3, 5)
    assert rc == 0
    assert np.array_equal(z["energy"], eno) and np.array_equal(z["pose"], xo)
