import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw, cases
descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
plan = bhw.Plan(descs)
out = torch.empty(plan.total, dtype=torch.int32, device="cuda")
def t(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {"min_run": os.environ.get("BHW_MIN_RUN")}
for ss in (4, 0):
    bhw.set_side_streams(ss)
    res["kept_ss%d_ms" % ss] = round(t(lambda: plan.execute(out=out)), 4)
n0 = bhw.launch_count(); plan.execute(out=out); res["launches"] = bhw.launch_count() - n0
# small mixed batch: pw 4..14 of all variants
d2 = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 15)]
p2 = bhw.Plan(d2); o2 = torch.empty(p2.total, dtype=torch.int32, device="cuda")
bhw.set_side_streams(4)
res["small_mixed_pw4_14_us"] = round(t(lambda: p2.execute(out=o2)) * 1e3, 1)
# single small windows through a plan
for pw in (10, 12, 14):
    p3 = bhw.Plan([bhw.variant_desc(6, pw, 17)]); o3 = torch.empty(1 << pw, dtype=torch.int32, device="cuda")
    res["single_bh4_pw%d_us" % pw] = round(t(lambda: p3.execute(out=o3), 50) * 1e3, 1)
print(json.dumps(res))
