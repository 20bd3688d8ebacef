import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw, cases
descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
plan = bhw.Plan(descs)
out = torch.empty(plan.total, dtype=torch.int32, device="cuda")
bhw.set_side_streams(0)
for _ in range(2):
    plan.execute(out=out)
torch.cuda.synchronize()
