import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, blackman_harris_win_b200 as bhw, bench
descs = bench.sweep_descs()
total = bhw.batch_total(descs)
out = torch.empty(total, dtype=torch.int32, device="cuda")
L = bhw.lib()
bhw.set_table_cache(False)
def t(plan, reps=20):
    for _ in range(3): plan.execute(out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.execute(out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for mt in (5, 7, 99):
    L.bhw_debug_set_spread_min_terms(mt)
    plan = bhw.Plan(descs)
    for side in (0, 4):
        bhw.set_side_streams(side)
        print(json.dumps({"spread_min_terms": mt, "side": side, "us": round(t(plan), 1)}), flush=True)
    plan.destroy()
