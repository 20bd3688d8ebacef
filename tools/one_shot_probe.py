#!/usr/bin/env python
"""One-shot cost of the sweep: bhw_generate_batch (transient plan: resolve + upload + build + synthesis + release) and
bhw_plan_create / bhw_plan_destroy, wall clock with a device synchronize, after a warm-up call.  N=1.
  python tools/one_shot_probe.py > gpurun_out/one_shot.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import bench  # noqa: E402

out = {}
for name, pw_max in (("sweep PHI_WIDTH 4..26 (230 windows)", 26), ("sweep PHI_WIDTH 4..20", 20), ("sweep PHI_WIDTH 4..14", 14)):
    descs = bench.sweep_descs(pw_max=pw_max)
    total = bhw.batch_total(descs)
    buf = torch.empty(total, dtype=torch.int32, device="cuda")
    for _ in range(2):
        bhw.generate_batch(descs, out=buf)
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        bhw.generate_batch(descs, out=buf)
    torch.cuda.synchronize()
    one_shot = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    for _ in range(reps):
        bhw.generate_batch(descs, 0, 1024, out=buf)          # planning of a tiny request: touches one window
    torch.cuda.synchronize()
    tiny = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    plans = [bhw.Plan(descs) for _ in range(3)]
    torch.cuda.synchronize()
    create = (time.perf_counter() - t0) / 3 * 1e3
    plan = plans[0]
    plan.execute(out=buf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        plan.execute(out=buf)
    torch.cuda.synchronize()
    execute = (time.perf_counter() - t0) / reps * 1e3
    for p in plans:
        p.destroy()
    out[name] = {"samples": total, "one_shot_ms": round(one_shot, 3), "tiny_request_ms": round(tiny, 3),
                 "plan_create_ms": round(create, 3), "plan_execute_ms": round(execute, 3)}
    del buf
print(json.dumps(out, indent=1))
