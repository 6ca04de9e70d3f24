import sys, os
import numpy as np
sys.path.insert(0, "tests")
import orclib
from fec import pkg, QPSK_MODCOD_OF_RATE
from orclib import ALL_CODES, code_params
import test_gpu_parity as T
dec = pkg.DVBS2Decoder(max_batch=256, max_trials=25)
for short, rate in ALL_CODES:
    dec.setDemodParams(QPSK_MODCOD_OF_RATE[rate], bool(short), False)
    n = 7 if not short else 13
    llr = T.make_llrs(short, rate, n, 17 + rate + 40 * short)
    llr[1, ::97] = 0
    llr[2] = np.where(llr[2] < 0, -128, 127)
    want_post, want_it = T.oracle_ldpc(short, rate, llr, 25)
    got = llr.copy()
    it = dec.ldpc_decode(got, 25)
    p = code_params(short, rate)
    bad = np.argwhere(got != want_post)
    print(short, rate, "iters ok", np.array_equal(it, want_it), it.tolist(), "mismatches", len(bad))
    if len(bad):
        fr = sorted(set(bad[:, 0].tolist()))
        print("  frames", fr, "K", p["K"], "first cols", bad[:12, 1].tolist(), "last", bad[-5:, 1].tolist())
        for f in fr[:2]:
            cols = bad[bad[:, 0] == f][:, 1]
            par = cols[cols >= p["K"]] - p["K"]
            print("   frame", f, "n", len(cols), "data", int((cols < p["K"]).sum()), "parity j(min,max)", (par // p["q"]).min() if len(par) else None, (par // p["q"]).max() if len(par) else None,
                  "i set", sorted(set((par % p["q"]).tolist()))[:10])
