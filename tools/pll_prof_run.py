#!/usr/bin/env python3
"""Minimal driver for ncu: the payload phase loop (K8) on device buffers, eight normal QPSK frames of one stream.
    ncu ... python tools/pll_prof_run.py [--mode 0|2]"""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
from test_pll_oracle import frames_for
ap = argparse.ArgumentParser(); ap.add_argument("--mode", type=int, default=0); a = ap.parse_args()
rng = np.random.default_rng(1)
pls, fr = frames_for("qpsk", 360, False, 8, rng, 8.0, 2e-5, 0, modcod=4, short=False)
d_in = torch.from_numpy(fr.view(np.float32).copy()).cuda(); d_out = torch.zeros_like(d_in)
g = pkg.S2PLSyncBlock(360, False); g.pll_set_params(0.004, 4, False, False, 0); g.pll_set_sequential(a.mode)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    g.pll_process_device(d_in.data_ptr(), 8, g.raw_frame_size, d_out.data_ptr(), 0, st)
torch.cuda.synchronize()
print("rounds", g.pll_rounds())
