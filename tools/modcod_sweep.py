#!/usr/bin/env python3
"""Device-resident throughput of the decode stage for every LDPC code (from LLRs) and for the four
constellations from PLFRAME symbols.  Prints a table and writes JSON.  Run on the GPU box."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SNR = {0: -2.0, 1: -1.0, 2: 0.0, 3: 1.3, 4: 2.6, 5: 3.4, 6: 4.3, 7: 5.0, 8: 5.5, 10: 6.5, 11: 6.7}
QPSK = {0: 1, 1: 2, 2: 3, 3: 4, 4: 5, 5: 6, 6: 7, 7: 8, 8: 9, 10: 10, 11: 11}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "modcod_sweep.json"))
    ap.add_argument("--margin-db", type=float, default=0.9)
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
    dev = torch.device("cuda", 0)
    rows = []
    rng = np.random.default_rng(3)
    for short in (False, True):
        for rate, modcod in QPSK.items():
            if short and rate == 11:
                continue
            info = pkg.modcod_info(modcod, short)
            n = 2048 if not short else 8192
            dec = pkg.DVBS2Decoder(devices=[0], max_batch=n, max_trials=25)
            dec.setDemodParams(modcod, short, False, 25)
            codes = np.stack([pkg.encode_fecframe(modcod, short, rng.integers(0, 256, info["kbch"] // 8, dtype=np.uint8)) for _ in range(8)])
            esn0 = SNR[rate] + (0.4 if short else 0.0) + args.margin_db
            a, sigma2 = 1 / np.sqrt(2.0), 1.0 / (2.0 * 10 ** (esn0 / 10.0))
            cw = torch.from_numpy(codes).to(dev)
            y = (1.0 - 2.0 * cw[torch.arange(n, device=dev) % 8].float()) * a
            gen = torch.Generator(device=dev)
            gen.manual_seed(1000 + 100 * int(short) + rate)   # same noise every run: results are comparable
            y += torch.randn(y.shape, device=dev, generator=gen) * float(np.sqrt(sigma2))
            llr = torch.clamp(torch.round(4.0 * 2.0 * a * y / sigma2), -127, 127).to(torch.int8)
            d_bb = torch.empty((n, info["kbch"] // 8), dtype=torch.uint8, device=dev)
            d_res = torch.empty((n, 16), dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream()
            for _ in range(2):
                dec.decode_batch_device(llr.data_ptr(), n, d_bb.data_ptr(), d_res.data_ptr(), st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            reps = 3
            for _ in range(reps):
                dec.decode_batch_device(llr.data_ptr(), n, d_bb.data_ptr(), d_res.data_ptr(), st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)
            it = res["ldpc_iters"].astype(np.int32)
            row = {"code": ("s" if short else "n") + pkg.RATE_NAMES[rate], "esn0_db": round(esn0, 2), "frames": n, "ms": round(ms, 3),
                   "frames_per_s": round(n / ms * 1e3), "gbit_s": round(n * info["kbch"] / ms / 1e6, 2),
                   "mean_iters": round(float(np.where(it < 0, 25, it).mean()), 2), "fer": round(float((res["bch_corr"] < 0).mean()), 4)}
            rows.append(row)
            print(row, flush=True)
            dec.close()
            del llr, y
    # symbols path: PLFRAME symbols in pinned host memory -> K1 demap + K2 + K3 ->
    # BBFRAMEs in pinned host memory through dvbs2fec_decode_plframes; symbols at the mapper's amplitude with
    # light noise (the reference's LUT scale only converges well above threshold, SURVEY note N7)
    import ctypes as C
    import time
    L = pkg.lib()
    sym_rows = []
    # two operating points per MODCOD: nearly noise-free (the demapper and PCIe are what is measured), and the code's
    # threshold + 1 dB, where the reference's LUT-scaled LLRs make every frame burn all 25 iterations:
    # the decode-bound end of the symbols path.  For the LUT constellations also the quantised path (2 B per symbol).
    thresholds = {4: 1.0, 12: 5.5, 18: 8.97, 28: 16.05}
    for modcod, name in ((4, "QPSK 1/2"), (12, "8PSK 3/5"), (18, "16APSK 2/3"), (28, "32APSK 9/10")):
        info = pkg.modcod_info(modcod, False)
        n = 1024
        dec = pkg.DVBS2Decoder(devices=[0], max_batch=512, max_trials=25)
        dec.setDemodParams(modcod, False, False, 25)
        base = np.stack([pkg.modulate(modcod, False, False, pkg.encode_fecframe(modcod, False, rng.integers(0, 256, info["kbch"] // 8, dtype=np.uint8)))
                         for _ in range(4)]).view(np.float32).reshape(4, -1)
        es = float((base[:, 180:] ** 2).sum() / (base[:, 180:].size / 2))   # mean symbol energy at the mapper's amplitude
        sym_bytes = base.shape[1] * 4
        nsym = info["nldpc"] // info["bits"]
        h_in, h_bb, h_res = L.dvbs2fec_alloc_pinned(n * sym_bytes), L.dvbs2fec_alloc_pinned(n * (info["kbch"] // 8)), L.dvbs2fec_alloc_pinned(n * 16)
        h_idx = L.dvbs2fec_alloc_pinned(n * nsym * 2)
        buf = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_float)), shape=(n, base.shape[1]))
        idx = np.ctypeslib.as_array(C.cast(h_idx, C.POINTER(C.c_uint8)), shape=(n, nsym * 2))
        for point, esn0 in (("clean", None), ("threshold+1dB", thresholds[modcod] + 1.0)):
            sigma = 0.04 if esn0 is None else float(np.sqrt(0.5 * es / 10 ** (esn0 / 10)))
            nrng = np.random.default_rng(modcod)
            for i in range(n):
                buf[i] = base[i % 4] + nrng.normal(0, sigma, base.shape[1]).astype(np.float32)
            paths = [("symbols (8 B/symbol)", lambda: L.dvbs2fec_decode_plframes(dec._h, h_in, n, h_bb, h_res), n * sym_bytes)]
            quant_us = None
            if modcod != 28:
                t0 = time.perf_counter()
                assert L.dvbs2fec_quantize_plframes(dec._h, h_in, n, h_idx) == 0
                quant_us = (time.perf_counter() - t0) / n * 1e6
                paths.append(("LUT coordinates (2 B/symbol)", lambda: L.dvbs2fec_decode_plframes_idx(dec._h, h_idx, n, h_bb, h_res), n * nsym * 2))
            ref_bb = None
            for pname, call, in_bytes in paths:
                dec.set_profiling(True)
                for _ in range(2):
                    assert call() == 0
                dec.kernel_times()
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    assert call() == 0
                dt = (time.perf_counter() - t0) / reps
                demap_ms, ldpc_ms, bch_ms, _ = dec.kernel_times()
                res = np.ctypeslib.as_array(C.cast(h_res, C.POINTER(C.c_uint8)), shape=(n * 16,)).view(pkg.RESULT_DTYPE).reshape(-1)
                bbv = np.ctypeslib.as_array(C.cast(h_bb, C.POINTER(C.c_uint8)), shape=(n * (info["kbch"] // 8),)).copy()
                if ref_bb is None:
                    ref_bb = (bbv, res.copy())
                else:   # the quantised path must give the very same frames
                    assert np.array_equal(bbv, ref_bb[0]) and np.array_equal(res["ldpc_iters"], ref_bb[1]["ldpc_iters"])
                it = res["ldpc_iters"].astype(np.int32)
                row = {"modcod": modcod, "name": name + " normal", "point": point, "esn0_db": esn0, "input": pname, "frames": n,
                       "e2e_ms": round(dt * 1e3, 3), "e2e_frames_per_s": round(n / dt), "e2e_gbit_s": round(n * info["kbch"] / dt / 1e9, 2),
                       "h2d_gb_s": round(in_bytes / dt / 1e9, 1),
                       "demap_ms": round(demap_ms / reps, 3), "ldpc_ms": round(ldpc_ms / reps, 3), "bch_ms": round(bch_ms / reps, 3),
                       "mean_iters": round(float(np.where(it < 0, 25, it).mean()), 2), "fer": round(float((res["bch_corr"] < 0).mean()), 4)}
                if pname.startswith("LUT"):
                    row["host_quantize_us_per_frame_1_core"] = round(quant_us, 1)
                sym_rows.append(row)
                print(row, flush=True)
        dec.close()
        for ptr in (h_in, h_bb, h_res, h_idx):
            L.dvbs2fec_free_pinned(ptr)
    json.dump({"note": "rows: device-resident LLR input, QPSK MODCOD of each rate, Es/N0 = threshold estimate + margin, fixed noise; "
                       "plframes: PLFRAMEs in pinned host memory through dvbs2fec_decode_plframes (8 B per symbol) and, for the LUT constellations, dvbs2fec_decode_plframes_idx (2 B per symbol, quantised on the host beforehand); H2D inside the time",
               "rows": rows, "plframes": sym_rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
