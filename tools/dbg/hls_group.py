import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import blackman_harris_win_b200 as bhw, harness as H
for model in (1, 0):
  for nw in (16, 17, 12, 24):
    for v in (1, 3, 6):
        for pw in (9, 13, nw, nw + 2):
            d = bhw.variant_desc(v, pw, nw, model=model)
            if bhw.validate(d): continue
            got = bhw.generate_batch([d]).cpu().numpy().astype(np.int64)
            want = H.orc_window(d)
            bad = np.nonzero(got != want)[0]
            print(model, nw, v, pw, "bad", len(bad), (int(bad[0]), int(got[bad[0]]), int(want[bad[0]])) if len(bad) else "")
print("---- multi-window HLS batch")
hls = [bhw.variant_desc(v, pw, 17, model=bhw.MODEL_HLS) for v in (1, 3, 6) for pw in (9, 14, 17, 19)]
got = bhw.generate_batch(hls).cpu().numpy().astype(np.int64)
off = 0
for d in hls:
    n = 1 << d.phi_width
    want = H.orc_window(d)
    g = got[off:off + n]
    bad = np.nonzero(g != want)[0]
    print(d.win_type, d.phi_width, "bad", len(bad), (int(bad[0]), int(g[bad[0]]), int(want[bad[0]])) if len(bad) else "")
    off += n
