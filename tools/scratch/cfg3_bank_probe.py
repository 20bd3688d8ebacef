import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw, cases
cfg = cases.baseline_configs()
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_bh7_n1m_dw32_dds"
d = cfg[name]
descs = [d.copy(aa=[int(a) - (i % 7) if k == 0 else int(a) for k, a in enumerate(d.aa)]) for i in range(64)]
plan = bhw.Plan(descs)
out = torch.empty(plan.total, dtype=torch.int32, device="cuda")
for _ in range(4):
    plan.execute(out=out)
torch.cuda.synchronize()
