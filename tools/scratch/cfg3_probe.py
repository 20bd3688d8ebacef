import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw, cases
cfg = cases.baseline_configs()
out = torch.empty(1 << 20, dtype=torch.int32, device="cuda")
for name in ("cfg3_bh7_n1m_dw32_dds48", "cfg3_bh7_n1m_dw32_dds"):
    for _ in range(3):
        bhw.generate(cfg[name], out=out)
torch.cuda.synchronize()
