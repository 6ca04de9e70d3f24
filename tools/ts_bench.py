#!/usr/bin/env python3
"""Throughput of the BBFRAME -> TS kernels (K6) on device-resident BBFRAMEs, against the HBM roofline, with the
reference parser (oracle/_ref, else the C oracle) timed on one host core beside it.  Run on the GPU box."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bbstream  # noqa: E402
import orclib  # noqa: E402


def make_frames(kbch, nframes, rng):
    kb, df = kbch // 8, kbch // 8 - 10
    stream = rng.integers(0, 256, nframes * df, dtype=np.uint8)
    frames = np.zeros((nframes, kb), np.uint8)
    hdrs = {}
    for f in range(nframes):
        syncd = ((-f * df) % 188) * 8
        if syncd not in hdrs:
            hdrs[syncd] = bbstream.bbheader(0xF0, 1504, df * 8, 0x47, syncd)
        frames[f, :10] = hdrs[syncd]
        frames[f, 10:] = stream[f * df:(f + 1) * df]
    return frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ts_bench.json"))
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
    dev = torch.device("cuda", 0)
    rows = []
    for name, kbch in (("n1/2", 32208), ("n9/10", 58192), ("s1/2", 7032)):
        n = args.frames if kbch > 10000 else args.frames * 4
        frames = make_frames(kbch, n, np.random.default_rng(1))
        kb = kbch // 8
        # several copies so that consecutive calls do not hit in L2 (126 MB)
        copies = max(2, int(300e6 // frames.nbytes) + 1)
        d_bb = [torch.from_numpy(frames).to(dev) for _ in range(copies)]
        d_out = [torch.zeros(n * kb + 188, dtype=torch.uint8, device=dev) for _ in range(copies)]
        d_n = torch.zeros(1, dtype=torch.int32, device=dev)
        g = pkg.BBFrameTSParser()
        g.setFrameSize(kbch)
        st = torch.cuda.current_stream()
        for k in range(copies):
            g.work_device(d_bb[k].data_ptr(), n, d_out[k].data_ptr(), d_out[k].numel(), d_n.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        produced = int(d_n.item())
        reps = 4 * copies
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(reps):
            g.work_device(d_bb[k % copies].data_ptr(), n, d_out[k % copies].data_ptr(), d_out[k % copies].numel(), d_n.data_ptr(),
                          st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        g.close()
        # CPU: the reference parser, one core
        if orclib.have_ref() and hasattr(orclib.ref(), "ref_ts_create"):
            r = orclib.ref()
            h = r.ref_ts_create(kbch)
            kind, work = "reference", lambda fr, out: r.ref_ts_work(h, fr, len(fr), out, len(out))
        else:
            o = orclib.oracle()
            h = o.orc_ts_create(kbch)
            kind, work = "port", lambda fr, out: o.orc_ts_work(h, fr, len(fr), out, len(out))
        out = np.zeros(n * kb + 4096, np.uint8)
        work(frames, out)
        t0 = time.perf_counter()
        creps = 5
        for _ in range(creps):
            cpu_n = work(frames, out)
        cpu_ms = (time.perf_counter() - t0) / creps * 1e3
        algo = frames.nbytes + produced
        row = {"code": name, "frames": n, "bytes_in": int(frames.nbytes), "bytes_out": produced, "gpu_ms": round(ms, 4),
               "gpu_frames_per_s": round(n / ms * 1e3), "gpu_gb_s": round(algo / ms / 1e6, 1),
               "cpu_kind": kind, "cpu_ms": round(cpu_ms, 3), "cpu_frames_per_s": round(n / cpu_ms * 1e3),
               "cpu_gb_s": round(algo / cpu_ms / 1e6, 2), "cpu_bytes_out": int(cpu_n)}
        rows.append(row)
        print(row, flush=True)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    json.dump({"note": "K6 plan+copy kernels, device-resident BBFRAMEs, inputs cycled through > L2; algorithmic bytes = BBFRAME bytes read + TS "
                       "bytes written", "measured_peaks": peaks, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
