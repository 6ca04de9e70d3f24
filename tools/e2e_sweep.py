#!/usr/bin/env python3
"""e2e throughput (pinned host buffers through dvbs2fec_decode_batch) vs chunk size; prints one line per setting."""
import ctypes as C
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
info = pkg.modcod_info(4, False)
N, kb = info["nldpc"], info["kbch"] // 8
dev = torch.device("cuda", 0)
codes = bench.make_codewords(pkg, 16, 1)
L = pkg.lib()
for frames, chunk in ((4096, 1024), (8192, 1024), (8192, 2048), (8192, 4096), (16384, 2048)):
    pool = bench.make_pool_torch(torch, codes, frames, 2.2, 5, dev)
    h_in, h_bb, h_res = L.dvbs2fec_alloc_pinned(frames * N), L.dvbs2fec_alloc_pinned(frames * kb), L.dvbs2fec_alloc_pinned(frames * 16)
    host = pool.cpu().numpy()
    C.memmove(h_in, host.ctypes.data, frames * N)
    dec = pkg.DVBS2Decoder(devices=[0], max_batch=chunk)
    dec.setDemodParams(4, False, False, 25)
    for _ in range(2):
        dec.decode_batch_raw(h_in, frames, h_bb, h_res)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        dec.decode_batch_raw(h_in, frames, h_bb, h_res)
    dt = (time.perf_counter() - t0) / reps
    print("frames %5d chunk %5d: %.2f ms  %.2f Gbit/s  (H2D alone would be %.2f ms at 50 GB/s)" % (
        frames, chunk, dt * 1e3, frames * info["kbch"] / dt / 1e9, frames * N / 50e9 * 1e3))
    dec.close()
    for p in (h_in, h_bb, h_res):
        L.dvbs2fec_free_pinned(p)
    del pool
