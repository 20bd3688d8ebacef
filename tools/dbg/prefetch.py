import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, blackman_harris_win_b200 as bhw, bench
L = bhw.lib()
out = torch.empty(1 << 26, dtype=torch.int32, device="cuda")
def t(plan, n, reps=20):
    for _ in range(3): plan.execute(out=out[:n])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.execute(out=out[:n])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for v in (9, 10):
    for pw in (23, 24, 25, 26):
        plan = bhw.Plan([bhw.variant_desc(v, pw, bench.VARIANT_DW[v])])
        res = {}
        for pl in (0, 8, 16, 32):
            L.bhw_debug_set_prefetch_lines(pl)
            res[pl] = round(t(plan, 1 << pw), 1)
        print(json.dumps({"v": v, "pw": pw, "us_by_prefetch_lines": res}), flush=True)
        plan.destroy()
