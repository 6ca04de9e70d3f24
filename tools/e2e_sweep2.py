#!/usr/bin/env python3
"""e2e (dvbs2fec_decode_batch from pinned host memory) against max_batch and frames per call; QPSK 1/2 normal, 2.2 dB."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
import torch
L = pkg.lib()
N, kbch = 64800, 32208
codes = bench.make_codewords(pkg, 16, 1)
pool = bench.make_pool_torch(torch, codes, 16384, 2.2, 3, torch.device("cuda", 0)).cpu().numpy()
n = len(pool)
h_in, h_bb, h_res = L.dvbs2fec_alloc_pinned(n * N), L.dvbs2fec_alloc_pinned(n * (kbch // 8)), L.dvbs2fec_alloc_pinned(n * 16)
C.memmove(h_in, pool.ctypes.data, n * N)
rows = []
for mb in (1024, 2048, 4096, 8192, 16384):
    dec = pkg.DVBS2Decoder(devices=[0], max_batch=mb, max_trials=25)
    dec.setDemodParams(4, False, False, 25)
    for frames in (4096, 16384):
        if frames < mb:
            continue
        for _ in range(2):
            dec.decode_batch_raw(h_in, frames, h_bb, h_res)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            dec.decode_batch_raw(h_in, frames, h_bb, h_res)
        dt = (time.perf_counter() - t0) / reps
        rows.append({"max_batch": mb, "frames_per_call": frames, "ms": round(dt * 1e3, 2), "gbit_s": round(frames * kbch / dt / 1e9, 2)})
        print(rows[-1], flush=True)
    dec.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "e2e_sweep2.json"), "w"), indent=1)
