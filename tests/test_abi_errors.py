"""Error behaviour of the C ABI on a GPU box: negative codes, never a crash or an exception across the boundary."""
import ctypes as C

import numpy as np
import pytest

from fec import pkg

pytestmark = pytest.mark.gpu


def test_calls_before_set_modcod_are_rejected():
    L = pkg.lib()
    h = C.c_void_p()
    assert L.dvbs2fec_create(None, C.byref(h)) == 0
    buf = np.zeros(64800, np.int8)
    out = np.zeros(4026, np.uint8)
    assert L.dvbs2fec_decode_batch(h, buf.ctypes.data_as(C.c_void_p), 1, out.ctypes.data_as(C.c_void_p), None) == pkg.EINVAL
    assert L.dvbs2fec_kbch(h) == pkg.EINVAL
    assert L.dvbs2fec_submit_llr(h, buf.ctypes.data_as(C.c_void_p), 1) == pkg.EINVAL
    assert b"set_modcod" in L.dvbs2fec_last_error()
    L.dvbs2fec_destroy(h)


def test_bad_arguments():
    L = pkg.lib()
    assert L.dvbs2fec_create(None, None) == pkg.EINVAL
    cfg = pkg.Config()
    cfg.n_devices = 1
    cfg.devices[0] = 99
    h = C.c_void_p()
    assert L.dvbs2fec_create(C.byref(cfg), C.byref(h)) == pkg.EINVAL
    dec = pkg.DVBS2Decoder(max_batch=8)
    dec.setDemodParams(4, True, False)
    assert L.dvbs2fec_decode_batch(dec._h, None, 1, None, None) == pkg.EINVAL
    assert L.dvbs2fec_decode_batch(dec._h, np.zeros(16200, np.int8).ctypes.data_as(C.c_void_p), -1, None, None) == pkg.EINVAL
    assert L.dvbs2fec_submit_plframe(dec._h, np.zeros(10, np.float32).ctypes.data_as(C.c_void_p), 5, 0) == pkg.EINVAL
    assert L.dvbs2fec_set_modcod(dec._h, 9, 1, 0, 0) == 0        # short 5/6 exists
    assert L.dvbs2fec_set_modcod(dec._h, 11, 1, 0, 0) == pkg.EINVAL  # short 9/10 does not
    assert L.dvbs2fec_set_modcod(dec._h, 17, 1, 0, 0) == pkg.EINVAL  # 8PSK 9/10 short neither
    # a rejected reconfiguration leaves the previous one in force
    assert (dec_n := L.dvbs2fec_nldpc(dec._h)) == 16200 and L.dvbs2fec_kldpc(dec._h) == 13320
    dec.close()
    L.dvbs2fec_destroy(None)  # no-op


def test_results_only_and_payload_only_outputs():
    dec = pkg.DVBS2Decoder(max_batch=8)
    dec.setDemodParams(4, True, False)
    L = pkg.lib()
    rng = np.random.default_rng(1)
    pay = rng.integers(0, 256, dec.kbch // 8, dtype=np.uint8)
    llr = np.where(pkg.encode_fecframe(4, True, pay) > 0, -9, 9).astype(np.int8)
    res = np.zeros(1, pkg.RESULT_DTYPE)
    assert L.dvbs2fec_decode_batch(dec._h, llr.ctypes.data_as(C.c_void_p), 1, None, res.ctypes.data_as(C.c_void_p)) == 0
    assert res["ldpc_iters"][0] == 0 and res["bch_corr"][0] == 0
    bb = np.zeros(dec.kbch // 8, np.uint8)
    assert L.dvbs2fec_decode_batch(dec._h, llr.ctypes.data_as(C.c_void_p), 1, bb.ctypes.data_as(C.c_void_p), None) == 0
    assert np.array_equal(bb, pay)
    dec.close()
