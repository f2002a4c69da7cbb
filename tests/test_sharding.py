"""Multi-rank host logic on CPU: contiguous instance sharding with a world_size-2 gloo group.
Every rank evaluates its own range (here with the CPU emulation of the kernel bodies — there is no GPU in this
container), no collective on the data path; the optional result gather is the only communication."""
import os
import socket

import numpy as np
import pytest

from sofa.rigidbodydynamics_b200.sharding import shard_range


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 1000, 1 << 20, (1 << 23) + 5):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            for (a0, a1), (b0, b1) in zip(got, got[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, q_path, out_path):
    import ctypes as C
    import sys
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    import helpers
    from helpers import P, from_tiled, to_tiled
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200._capi import make_desc
    from sofa.rigidbodydynamics_b200.sharding import gather_results, shard_range as sr
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d = mdl.solo12()
    data = np.load(q_path)
    lo, hi = sr(n, rank, world)
    q, v, a = data["q"][lo:hi], data["v"][lo:hi], data["a"][lo:hi]
    emul = helpers.load_emul()
    cd, keep = make_desc(d)
    m = hi - lo
    tile = 32
    nt = (m + tile - 1) // tile
    tau = np.zeros((nt, d.nv, tile)); M = np.zeros((nt, d.nv * (d.nv + 1) // 2, tile))
    qs, vs, as_ = to_tiled(q, tile), to_tiled(v, tile), to_tiled(a, tile)
    assert emul.emul_eval(C.addressof(cd), m, tile, 0, 1, P(qs), P(vs), P(as_), None, None, P(M), P(tau)) == 0
    local = torch.from_numpy(from_tiled(tau, m))
    # equal shard sizes are required by all_gather: n is a multiple of world in this test
    parts = gather_results(local)
    if rank == 0:
        np.save(out_path, torch.cat(parts, dim=0).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_eval_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    d = mdl.solo12()
    n, world = 96, 2
    q, v, a = mdl.random_state(d, n, seed=77)
    q_path = str(tmp_path / "in.npz"); out_path = str(tmp_path / "tau.npy")
    np.savez(q_path, q=q, v=v, a=a)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, q_path, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    ref, _ = oracle.eval_batch(d, q, v, a, 1, want=("tau",))
    assert got.shape == (n, d.nv)
    assert np.max(np.abs(got - ref["tau"])) <= 1e-10 * max(1.0, np.max(np.abs(ref["tau"])))
