#!/usr/bin/env python
"""BASELINE config 5 across GPUs: all 10 variants x PHI_WIDTH 4..26 (1.34 G samples, 5.37 GB) sharded by
contiguous flat sample range over the ranks (bhw_shard_range / bhw_shard_windows), one process per
GPU, no data-path collective.  Strong scaling: the total work is fixed.  Launch with torchrun
(or plain python for 1 GPU); rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sweep_multi.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import cases  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    descs = bhw.desc_array([bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)])
    total = bhw.batch_total(descs)
    by_cost = "--by-samples" not in sys.argv       # default: cost-balanced cuts (bhw_shard_range_cost)
    b, c = bhw.shard_range_cost(descs, rank, world) if by_cost else bhw.shard_range(total, rank, world)
    first, touched, lb = bhw.shard_windows(descs, b, c)
    mine = descs[first:first + touched]                  # ctypes array slice -> list of descriptors
    plan = bhw.Plan(mine) if c else None                 # a rank whose cuts snapped onto one boundary has nothing to do
    out = torch.empty(c, dtype=torch.int32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = {}
    for cache in (False, True):
        bhw.set_table_cache(cache)
        for _ in range(3 if plan else 0):
            plan.execute(lb, c, out=out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps if plan else 0):
            plan.execute(lb, c, out=out)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        all_ms = [t.clone() for _ in range(world)]
        if world > 1:
            dist.all_gather(all_ms, t)
        per_rank = [round(float(x.item()), 4) for x in all_ms]
        res["tables_kept" if cache else "tables_rebuilt"] = {
            "ms_per_sweep_max_over_ranks": max(per_rank), "per_rank_ms": per_rank,
            "gsamples_per_s": round(total / max(per_rank) / 1e6, 1)}
    bhw.set_table_cache(True)
    if rank == 0:
        print(json.dumps({"config": "cfg5 sweep, 10 variants x PHI_WIDTH 4..26, sharded by flat sample range",
                          "n_gpus": world, "samples": total, "scaling": "strong",
                          "cuts": "cost-balanced (bhw_shard_range_cost)" if by_cost else "equal sample counts (bhw_shard_range)",
                          **res}))
    if plan:
        plan.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
