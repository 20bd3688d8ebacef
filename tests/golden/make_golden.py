#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference's own C++ compiled in this container.

Run here (needs /root/reference): `make -C oracle ref && python tests/golden/make_golden.py`.
Outputs
  reference_vectors.npz : full integer tables produced by the UNMODIFIED reference sources
                          (hls/windows/win_function.cpp, hls/cordic/cordic.cpp,
                          cpp/cordic_sincos.cpp) compiled by oracle/build_ref.sh
  reference_kat.json    : sha256 (one decimal per line) + first values of larger tables from the
                          same binaries, incl. the BASELINE widths
The RTL anchors (rtl_kat.json) are NOT produced here: no VHDL simulator exists in this image;
they are the [derived] known answers of SURVEY.md 8(c), written by an independent Python
restatement during the survey, kept verbatim as a regression anchor ("parity unpinned").
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402

TYPES = (1, 2, 3, 4, 5, 7)


def main():
    vec = {}
    kat = {"_how": "oracle/_ref binaries = unmodified reference C++ (ap_int stand-in: oracle/shim/ap_int.h)",
           "hls_win": [], "hls_cordic": [], "cpp": []}
    for (np_, nw) in H.ref_configs("hls_win"):
        n = 1 << np_
        for t in TYPES:
            w = H.ref_hls_window(np_, nw, t)
            if n <= 4096:
                vec[f"hls_win_np{np_}_nw{nw}_t{t}"] = w.astype(np.int64)
            kat["hls_win"].append({"nphase": np_, "nwidth": nw, "type": t, "sha256": H.sha_lines(w),
                                   "first": [int(x) for x in w[:4]], "mid": int(w[n // 2])})
    for (np_, nw) in H.ref_configs("hls_cordic"):
        s, c = H.ref_hls_cordic(np_, nw)
        if (1 << np_) <= 4096:
            vec[f"hls_cordic_np{np_}_nw{nw}"] = np.stack([s, c]).astype(np.int64)
        kat["hls_cordic"].append({"nphase": np_, "nwidth": nw, "sha256_s_c": H.sha_pairs(s, c),
                                  "cos_first": [int(x) for x in c[:4]]})
    for (pw, dw) in H.ref_configs("cpp"):
        s, c = H.ref_cpp_cordic(pw, dw)
        if (1 << pw) <= 16384:
            vec[f"cpp_pw{pw}_dw{dw}"] = np.stack([s, c]).astype(np.int32)
        kat["cpp"].append({"phase_width": pw, "data_width": dw, "sha256_s_c": H.sha_pairs(s, c),
                           "first": [[int(s[i]), int(c[i])] for i in range(2)]})
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **vec)
    with open(os.path.join(HERE, "reference_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", len(vec), "vectors;", sum(len(v) for k, v in kat.items() if k != "_how"), "hashes")


if __name__ == "__main__":
    main()
