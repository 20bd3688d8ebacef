import sys, os, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, blackman_harris_win_b200 as bhw, bench
descs = bench.sweep_descs()
total = bhw.batch_total(descs)
out = torch.empty(total, dtype=torch.int32, device="cuda")
plan = bhw.Plan(descs)
bhw.set_table_cache(False)
L = bhw.lib()
def t(reps=20):
    for _ in range(3): plan.execute(out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.execute(out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for side in (4, 6, 8):
    bhw.set_side_streams(side)
    for n in (0, 1, 2, 3, 4, 8):
        L.bhw_debug_set_defer_ctas(n)
        print(json.dumps({"side": side, "defer_ctas": n, "us": round(t(), 1)}), flush=True)
