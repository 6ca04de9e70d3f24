#!/usr/bin/env python3
"""Minimal driver for ncu: a few device-resident decode passes of the bench workload, nothing else.
    ncu ... python tools/prof_run.py [--pool F] [--reps R] [--esn0 DB] [--modcod M] [--short]"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", type=int, default=592)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--esn0", type=float, default=2.2)
    ap.add_argument("--modcod", type=int, default=None, help="MODCOD instead of the bench workload's")
    ap.add_argument("--short", action="store_true")
    args = ap.parse_args()
    if args.modcod is not None:
        bench.MODCOD, bench.SHORT = args.modcod, args.short
    import torch
    pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
    dev = torch.device("cuda", 0)
    dec = pkg.DVBS2Decoder(devices=[0], max_batch=args.pool, max_trials=bench.MAX_TRIALS)
    dec.setDemodParams(bench.MODCOD, bench.SHORT, False, bench.MAX_TRIALS)
    codes = bench.make_codewords(pkg, 16, 1)
    pool = bench.make_pool_torch(torch, codes, args.pool, args.esn0, 100, dev)
    d_bb = torch.empty((args.pool, dec.kbch // 8), dtype=torch.uint8, device=dev)
    d_res = torch.empty((args.pool, 16), dtype=torch.uint8, device=dev)
    for _ in range(args.reps):
        dec.decode_batch_device(pool.data_ptr(), args.pool, d_bb.data_ptr(), d_res.data_ptr(), 0)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)
    print("mean iters", np.where(res["ldpc_iters"] < 0, 25, res["ldpc_iters"]).mean(), "fer", (res["bch_corr"] < 0).mean())


if __name__ == "__main__":
    main()
