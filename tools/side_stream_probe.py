#!/usr/bin/env python
"""How many internal side streams (bhw_set_side_streams) the config-5 sweep wants: ms per step, tables rebuilt
every step (the bench's rule) and tables kept, for 0..8 side streams; every setting's output is compared with
the serial one.  One JSON line per setting.

  python tools/side_stream_probe.py [--steps 40]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--rounds", type=int, default=2)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    descs = bench.sweep_descs()
    plan = bhw.Plan(descs)
    total = plan.total
    out = torch.empty(total, dtype=torch.int32, device="cuda")
    bhw.set_side_streams(0)
    bhw.set_table_cache(False)
    plan.execute(out=out)
    torch.cuda.synchronize()
    ref = out.clone()
    for rnd in range(args.rounds):                # two rounds: the order of the settings must not matter
        for n in (4, 0, 1, 2, 3, 5, 6, 8, 4):
            bhw.set_side_streams(n)
            line = {"side_streams": n, "round": rnd}
            for keep in (False, True):
                bhw.set_table_cache(keep)
                out.zero_()
                ms = bench._time_loop(lambda: plan.execute(out=out), args.steps, warm=4)
                line["ms_tables_kept" if keep else "ms_tables_rebuilt"] = round(ms, 4)
                line["equal"] = bool(torch.equal(out, ref)) and line.get("equal", True)
            print(json.dumps(line), flush=True)
    bhw.set_side_streams(4)
    bhw.set_table_cache(True)
    plan.destroy()


if __name__ == "__main__":
    main()
