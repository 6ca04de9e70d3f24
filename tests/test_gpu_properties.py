"""Size-independent properties at the benchmark's own sizes (normal frames, batches of hundreds), where
running the CPU oracle on everything would take minutes: round trips, idempotence, order independence,
plus an oracle spot check on a random sample."""
import ctypes as C

import numpy as np
import pytest

import orclib
from fec import pkg

pytestmark = pytest.mark.gpu


def make_batch(modcod, short, n, esn0, seed, ncodes=16):
    rng = np.random.default_rng(seed)
    info = pkg.modcod_info(modcod, short)
    payloads = rng.integers(0, 256, (ncodes, info["kbch"] // 8), dtype=np.uint8)
    codes = np.stack([pkg.encode_fecframe(modcod, short, p) for p in payloads])
    a = 1 / np.sqrt(2.0)
    sigma2 = 1.0 / (2.0 * 10 ** (esn0 / 10.0))
    idx = np.arange(n) % ncodes
    y = (1.0 - 2.0 * codes[idx].astype(np.float32)) * a + rng.normal(0, np.sqrt(sigma2), (n, info["nldpc"])).astype(np.float32)
    llr = np.clip(np.rint(4.0 * 2.0 * a * y / sigma2), -127, 127).astype(np.int8)
    return llr, payloads[idx]


def test_round_trip_qpsk_half_normal_batch():
    dec = pkg.DVBS2Decoder(max_batch=1024, max_trials=25)
    dec.setDemodParams(4, False, False)
    n = 1500  # > max_batch: exercises chunking and the double-buffered slots
    llr, payload = make_batch(4, False, n, 2.2, 1)
    bb, res = dec.decode_batch(llr)
    ok = res["bch_corr"] >= 0
    assert ok.mean() > 0.995
    assert np.array_equal(bb[ok], payload[ok])
    assert (res["ldpc_iters"][ok] >= 0).all() and res["ldpc_iters"].max() <= 25
    assert np.array_equal(res["tag"], np.arange(n))
    # oracle spot check, iteration counts included
    o = orclib.oracle()
    rng = np.random.default_rng(0)
    for i in rng.choice(n, 6, replace=False):
        want = np.zeros(dec.kbch // 8, np.uint8)
        it, co = C.c_int(), C.c_int()
        o.orc_decode_frame(0, 3, llr[i].copy(), 25, want, C.byref(it), C.byref(co))
        assert (res["ldpc_iters"][i], res["bch_corr"][i]) == (it.value, co.value)
        assert np.array_equal(bb[i], want)
    # order independence: a permuted batch gives permuted results (frames never interact)
    perm = rng.permutation(n)[:700]
    bb2, res2 = dec.decode_batch(llr[perm])
    assert np.array_equal(bb2, bb[perm])
    assert np.array_equal(res2["ldpc_iters"], res["ldpc_iters"][perm])
    dec.close()


def test_posterior_is_a_fixed_point():
    """decoding the posterior LLRs of a converged frame again takes 0 iterations and changes nothing"""
    dec = pkg.DVBS2Decoder(max_batch=256)
    dec.setDemodParams(13, False, False)   # 8PSK 2/3 -> code n2/3
    llr, _ = make_batch(13, False, 64, 4.2, 2)
    it = dec.ldpc_decode(llr, 25)
    conv = it >= 0
    assert conv.sum() > 32
    again = llr.copy()
    it2 = dec.ldpc_decode(again, 25)
    assert (it2[conv] == 0).all()
    assert np.array_equal(again[conv], llr[conv])
    dec.close()


def test_worst_case_all_frames_fail():
    """far below threshold: every frame burns max_trials and reports -1 / BCH failure, like the reference"""
    dec = pkg.DVBS2Decoder(max_batch=256, max_trials=10)
    dec.setDemodParams(4, False, False, 10)
    llr, _ = make_batch(4, False, 40, -1.0, 3)
    bb, res = dec.decode_batch(llr)
    assert (res["ldpc_iters"] == -1).all() and (res["bch_corr"] == -1).all() and (res["flags"] == 3).all()
    o = orclib.oracle()
    want = np.zeros(dec.kbch // 8, np.uint8)
    it, co = C.c_int(), C.c_int()
    o.orc_decode_frame(0, 3, llr[7].copy(), 10, want, C.byref(it), C.byref(co))
    assert (it.value, co.value) == (-1, -1) and np.array_equal(bb[7], want)
    dec.close()


def test_empty_and_single_frame_batches():
    dec = pkg.DVBS2Decoder(max_batch=16)
    dec.setDemodParams(4, True, False)
    bb, res = dec.decode_batch(np.zeros((0, dec.N), np.int8))
    assert bb.shape == (0, dec.kbch // 8) and len(res) == 0
    llr, payload = make_batch(4, True, 1, 6.0, 4, ncodes=1)
    bb, res = dec.decode_batch(llr)
    assert np.array_equal(bb[0], payload[0]) and res["bch_corr"][0] >= 0
    dec.close()


def test_in_process_multi_gpu_sharding_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    llr, payload = make_batch(4, True, 301, 3.0, 5)
    one = pkg.DVBS2Decoder(devices=[0], max_batch=64)
    one.setDemodParams(4, True, False)
    two = pkg.DVBS2Decoder(devices=[0, 1], max_batch=64)
    two.setDemodParams(4, True, False)
    bb1, r1 = one.decode_batch(llr)
    bb2, r2 = two.decode_batch(llr)
    assert np.array_equal(bb1, bb2) and np.array_equal(r1, r2)
    one.close()
    two.close()
