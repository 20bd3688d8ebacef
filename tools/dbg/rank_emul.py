import sys, os, json, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C, numpy as np, torch, blackman_harris_win_b200 as bhw, bench, harness as H
descs = bench.sweep_descs()
world = int(sys.argv[1])
bhw.set_table_cache(False)
for rank in range(world):
    b, c, first, touched, local = bench.rank_sweep(descs, rank, world)
    print("rank", rank, b, c, first, touched, local, flush=True)
    mine = [bhw.BhwDesc.from_buffer_copy(bytes(descs[i])) for i in range(first, first + touched)]
    plan = bhw.Plan(mine)
    out = torch.empty(c, dtype=torch.int32, device="cuda")
    for _ in range(3):
        plan.execute(local, c, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.execute(local, c, out=out)
    e1.record(); torch.cuda.synchronize()
    # spot check the slice ends against the oracle
    offs = np.cumsum([0] + [1 << d.phi_width for d in mine])
    for pos in (local, local + c - 2048):
        wi = int(np.searchsorted(offs, pos, side="right") - 1)
        n0 = pos - offs[wi]
        cnt = min(2048, (1 << mine[wi].phi_width) - n0)
        got = out[pos - local: pos - local + cnt].cpu().numpy().astype(np.int64)
        assert np.array_equal(got, H.orc_window(mine[wi], int(n0), int(cnt))), (rank, pos)
    print("rank", rank, "us", round(e0.elapsed_time(e1) / 10 * 1e3, 1), flush=True)
    plan.destroy()
