import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw
v, pw, dw = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = bhw.variant_desc(v, pw, dw)
plan = bhw.Plan([d])
out = torch.empty(1 << pw, dtype=torch.int32, device="cuda")
for _ in range(4):
    plan.execute(out=out)
torch.cuda.synchronize()
