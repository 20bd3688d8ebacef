"""CPU tier, world_size 2 over gloo: the N>1 path of bench.py / the sharding contract.  Each rank
takes its bhw_shard_range slice of the flat batch, fills it (here: with the oracle, standing in for
the device), and the concatenation must equal the unsharded batch.  There is no data-path
collective in the product; gloo is used only to compare the shards and to take the max-over-ranks
time the way bench.py does."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import blackman_harris_win_b200 as bhw
    import harness as H
    import bench
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        descs = [bhw.variant_desc(v, pw, 16) for v in (1, 3, 6) for pw in (4, 7, 10, 5)]
        total = bhw.batch_total(descs)
        b, c = bhw.shard_range(total, rank, world)
        mine = H.orc_batch(descs, b, c)
        # gather variable-length shards
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([c], dtype=torch.int64))
        mx = int(max(s.item() for s in sizes))
        pad = torch.zeros(mx, dtype=torch.int64)
        pad[:c] = torch.from_numpy(mine)
        parts = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(parts, pad)
        full = np.concatenate([p[: int(s.item())].numpy() for p, s in zip(parts, sizes)])
        ok = np.array_equal(full, H.orc_batch(descs, 0, total))
        # bench.py's cross-rank timing reduction (max over ranks)
        t = bench.max_over_ranks(float(rank + 1))
        # bench.py's per-rank workload split
        wl = bench.rank_workload(rank, world, nwin_per_gpu=3)
        # bench.py's headline split: the win_selector sweep cut into cost-balanced contiguous slices (strong
        # scaling); every rank fills its slice from the windows it touches only, as bench.py plans them
        sweep = bench.sweep_descs(pw_max=11)
        sb, sc, first, touched, local = bench.rank_sweep(sweep, rank, world)
        sub = [sweep[i] for i in range(first, first + touched)]
        mine2 = H.orc_batch(sub, local, sc)
        whole = H.orc_batch(list(sweep), 0, bhw.batch_total(sweep))
        ok2 = np.array_equal(mine2, whole[sb:sb + sc])
        times = bench.gather_ranks(float(10 * (rank + 1)))
        q.put((rank, ok and ok2, t, b, c, wl, (sb, sc, int(bhw.batch_total(sweep))), times))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert all(r[2] == 2.0 for r in res)            # max over ranks
    assert res[0][3] == 0 and res[0][3] + res[0][4] == res[1][3]
    # weak scaling: global batch = world * per-GPU windows, rank r owns flat slice r
    (b0, c0, tot0), (b1, c1, tot1) = res[0][5], res[1][5]
    assert tot0 == tot1 and b0 == 0 and b0 + c0 == b1 and b1 + c1 == tot0 and c0 == c1
    # strong scaling of the sweep: the two cost-balanced slices tile the flat range in order
    (sb0, sc0, st0), (sb1, sc1, st1) = res[0][6], res[1][6]
    assert st0 == st1 and sb0 == 0 and sb0 + sc0 == sb1 and sb1 + sc1 == st0 and sc0 > 0 and sc1 > 0
    assert res[0][7] == [10.0, 20.0] and res[1][7] == [10.0, 20.0]      # per-rank times, gathered
