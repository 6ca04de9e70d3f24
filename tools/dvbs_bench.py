#!/usr/bin/env python3
"""K9 (DVB-S outer decoder) on device buffers against the reference's own classes over libcorrect on one host core:
frames of 8 x 204 bytes per second, algorithmic GB/s (1632 B in + 1504 B + 32 B out per frame)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
import torch
import dvbs_stream, orclib
from test_dvbs_oracle import OrcOuter, RefOuter

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--frames", type=int, default=4096); a = ap.parse_args()
res = []
base_ts, base = dvbs_stream.outer_stream(256, np.random.default_rng(1))
for name, span in (("clean", (0, 0)), ("0-8 byte errors per 204 channel bytes", (0, 8)), ("6-20 (some packets beyond the code)", (6, 20))):
    rng = np.random.default_rng(2)
    ch = np.tile(base, a.frames // 256)
    bad = dvbs_stream.add_errors(ch, rng, per_packet=span)
    n = a.frames
    d_in = torch.from_numpy(bad).cuda(); d_out = torch.zeros(n * 1504, dtype=torch.uint8, device="cuda"); d_err = torch.zeros(n * 8, dtype=torch.int32, device="cuda")
    g = pkg.DVBSOuterDecoder()
    st = torch.cuda.current_stream().cuda_stream
    g.process_device(d_in.data_ptr(), n, 1632, d_out.data_ptr(), d_err.data_ptr(), st); torch.cuda.synchronize()
    first = d_out.cpu().numpy().reshape(-1, 188).copy(); first_err = d_err.cpu().numpy().copy()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.process_device(d_in.data_ptr(), n, 1632, d_out.data_ptr(), d_err.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    cpu = RefOuter() if orclib.have_ref() and hasattr(orclib.ref(), "ref_dvbs_outer_create") else OrcOuter()
    nc = min(n, 512)
    t0 = time.perf_counter(); want, we = cpu.process(bad, nc); cpu_ms = (time.perf_counter() - t0) * 1e3 * n / nc
    ok = bool(np.array_equal(first[:nc * 8], want) and np.array_equal(first_err[:nc * 8], we))
    res.append(dict(case=name, frames=n, gpu_ms=round(ms, 4), frames_per_s=round(n / ms * 1e3), ts_mbit_s=round(n * 1504 * 8 / ms / 1e3, 1),
                    algorithmic_gb_s=round(n * (1632 + 1504 + 32) / ms / 1e6, 2), cpu_ms_1_core=round(cpu_ms, 2), cpu_kind="reference" if isinstance(cpu, RefOuter) else "oracle",
                    speedup=round(cpu_ms / ms, 1), equal_to_cpu_on_first_frames=ok, mean_errors_per_packet=float(first_err.mean())))
    print(res[-1], flush=True)
    g.close()
if a.out: json.dump(res, open(a.out, "w"), indent=1)
