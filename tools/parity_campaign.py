#!/usr/bin/env python3
"""Large differential run: CUDA decode stage vs the CPU oracle on thousands of frames (run on the GPU box).

    python tools/parity_campaign.py [--frames-b4 4000] [--frames-other 200] [--out profiles/r01_parity_campaign.json]

For every LDPC/BCH code: random payloads -> in-tree transmitter -> AWGN int8 LLRs at several Es/N0 points
around the code's threshold -> (a) dvbs2fec_decode_batch, (b) oracle orc_decode_frame in worker processes.
Compared per frame: BBFRAME bytes, LDPC iteration count, BCH correction count.  Exits non-zero on any mismatch."""
import argparse
import ctypes as C
import importlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SNR = {0: -2.0, 1: -1.0, 2: 0.0, 3: 1.3, 4: 2.6, 5: 3.4, 6: 4.3, 7: 5.0, 8: 5.5, 10: 6.5, 11: 6.7}
QPSK_MODCOD_OF_RATE = {0: 1, 1: 2, 2: 3, 3: 4, 4: 5, 5: 6, 6: 7, 7: 8, 8: 9, 10: 10, 11: 11}


def _oracle_job(job):
    import orclib
    short, rate, llr, max_trials = job
    o = orclib.oracle()
    p = orclib.code_params(short, rate)
    out = []
    bb = np.zeros(p["kbch"] // 8, np.uint8)
    for i in range(len(llr)):
        it, co = C.c_int(), C.c_int()
        o.orc_decode_frame(short, rate, llr[i].copy(), max_trials, bb, C.byref(it), C.byref(co))
        out.append((it.value, co.value, bb.tobytes()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-b4", type=int, default=4000)
    ap.add_argument("--frames-other", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_parity_campaign.json"))
    args = ap.parse_args()
    pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
    import orclib
    dec = pkg.DVBS2Decoder(max_batch=1024, max_trials=25)
    rng = np.random.default_rng(20260117)
    procs = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(procs)
    report, total, bad = [], 0, 0
    t_start = time.time()
    for short, rate in orclib.ALL_CODES:
        n = args.frames_b4 if (short, rate) == (0, 3) else args.frames_other
        modcod = QPSK_MODCOD_OF_RATE[rate]
        dec.setDemodParams(modcod, bool(short), False, 25)
        info = pkg.modcod_info(modcod, bool(short))
        ncodes = 8
        codes = np.stack([pkg.encode_fecframe(modcod, bool(short), rng.integers(0, 256, info["kbch"] // 8, dtype=np.uint8))
                          for _ in range(ncodes)])
        a = 1 / np.sqrt(2.0)
        offs = np.array([-0.9, -0.5, -0.2, 0.1, 0.5, 1.5, 6.0])
        snr = SNR[rate] + (0.4 if short else 0.0) + offs[np.arange(n) % len(offs)]
        sigma2 = 1.0 / (2.0 * 10 ** (snr / 10.0))
        llr = np.zeros((n, info["nldpc"]), np.int8)
        for i in range(n):
            y = (1.0 - 2.0 * codes[i % ncodes].astype(np.float32)) * a + rng.normal(0, np.sqrt(sigma2[i]), info["nldpc"]).astype(np.float32)
            llr[i] = np.clip(np.rint(4.0 * 2.0 * a * y / sigma2[i]), -127, 127).astype(np.int8)
        bb, res = dec.decode_batch(llr)
        per = max(1, n // (procs * 2))
        jobs = [(short, rate, llr[k:k + per], 25) for k in range(0, n, per)]
        want = [x for part in pool.map(_oracle_job, jobs) for x in part]
        mism = 0
        for i, (it, co, wbb) in enumerate(want):
            if res["ldpc_iters"][i] != it or res["bch_corr"][i] != co or bb[i].tobytes() != wbb:
                mism += 1
        its = res["ldpc_iters"]
        report.append({"code": ("s" if short else "n") + pkg.RATE_NAMES[rate], "frames": n, "mismatches": mism,
                       "ldpc_failed": int((its < 0).sum()), "bch_failed": int((res["bch_corr"] < 0).sum()),
                       "bch_corrected_frames": int((res["bch_corr"] > 0).sum()), "iters_min": int(its[its >= 0].min()) if (its >= 0).any() else None,
                       "iters_max": int(its.max())})
        total += n
        bad += mism
        print(report[-1], flush=True)
    pool.close()
    summary = {"total_frames": total, "mismatches": bad, "seconds": round(time.time() - t_start, 1), "host_processes": procs,
               "compared": ["BBFRAME bytes", "ldpc_iters", "bch_corr"], "codes": report}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(summary, open(args.out, "w"), indent=1)
    print("TOTAL", total, "mismatches", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
