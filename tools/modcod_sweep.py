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
    json.dump({"note": "device-resident LLR input, QPSK MODCOD of each rate, Es/N0 = threshold estimate + margin", "rows": rows},
              open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
