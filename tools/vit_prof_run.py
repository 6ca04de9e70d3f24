#!/usr/bin/env python3
"""Minimal driver for ncu: the locked Viterbi decoder on device buffers (K11), a few calls of --blocks blocks at one rate.
    ncu ... python tools/vit_prof_run.py [--rate R] [--blocks N] [--reps K]"""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dvbs_stream
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
ap = argparse.ArgumentParser(); ap.add_argument("--rate", type=int, default=0); ap.add_argument("--blocks", type=int, default=2046); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
rng = np.random.default_rng(a.rate)
per_block = [4096, 5462, 6144, 6827, 7168][a.rate]
base = dvbs_stream.inner_softs(rng.integers(0, 2, per_block * 68, dtype=np.uint8), a.rate, rng, sigma=[14.0, 8.0, 8.0, 5.0, 5.0][a.rate])[:66 * 8192]
s = np.tile(base, a.blocks // 66)
d_in = torch.from_numpy(s).cuda(); d_out = torch.zeros(len(s), dtype=torch.uint8, device="cuda")
g = pkg.DVBSViterbi()
for _ in range(a.reps):
    n = g.process_device(d_in.data_ptr(), len(s), d_out.data_ptr())
torch.cuda.synchronize()
print("bits", n, "stats", g.stats(), "counters", g.counters())
