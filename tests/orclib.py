"""ctypes access to the CPU oracle (oracle/libdvbs2_oracle.so) and, when it has been built, the
compiled reference (oracle/_ref/libdvbs2_ref.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# reference rate enum (dvbs2/dvbs2.h:11-25)
RATES = {"1/4": 0, "1/3": 1, "2/5": 2, "1/2": 3, "3/5": 4, "2/3": 5, "3/4": 6, "4/5": 7, "5/6": 8,
         "8/9": 10, "9/10": 11}
NORMAL_RATES = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11]
SHORT_RATES = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10]
ALL_CODES = [(0, r) for r in NORMAL_RATES] + [(1, r) for r in SHORT_RATES]

_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def _build_oracle():
    so = os.path.join(ORACLE_DIR, "libdvbs2_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_ldpc.c", "oracle_bch.c", "oracle_demap.c", "oracle_ts.c", "oracle_plsync.c",
                                                   "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return so


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(_build_oracle())
        ip = C.POINTER(C.c_int)
        lib.orc_code_params.argtypes = [C.c_int, C.c_int, ip, ip, ip, ip, ip, ip]
        lib.orc_ldpc_decode.argtypes = [C.c_int, C.c_int, _i8p, C.c_int]
        lib.orc_ldpc_encode_bits.argtypes = [C.c_int, C.c_int, _u8p, _u8p]
        lib.orc_ldpc_schedule.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_repack.argtypes = [_i8p, C.c_int, _u8p]
        lib.orc_bch_decode.argtypes = [C.c_int, C.c_int, _u8p]
        lib.orc_bch_encode.argtypes = [C.c_int, C.c_int, _u8p]
        lib.orc_descramble.argtypes = [C.c_int, C.c_int, _u8p]
        lib.orc_bbheader_crc8.argtypes = [_u8p]
        lib.orc_bbheader_crc8.restype = C.c_uint
        lib.orc_decode_frame.argtypes = [C.c_int, C.c_int, _i8p, C.c_int, _u8p, ip, ip]
        lib.orc_const_create.argtypes = [C.c_int, C.c_float, C.c_float]
        lib.orc_const_create.restype = C.c_void_p
        lib.orc_const_destroy.argtypes = [C.c_void_p]
        lib.orc_const_bits.argtypes = [C.c_void_p]
        lib.orc_const_lut.argtypes = [C.c_void_p]
        lib.orc_const_lut.restype = C.POINTER(C.c_int8)
        lib.orc_const_points.argtypes = [C.c_void_p, _f32p]
        lib.orc_demod_soft_calc.argtypes = [C.c_void_p, C.c_float, C.c_float, _i8p]
        lib.orc_demod_soft_lut.argtypes = [C.c_void_p, C.c_float, C.c_float, _i8p]
        lib.orc_mod.argtypes = [C.c_void_p, C.c_int, _f32p]
        lib.orc_deinterleave.argtypes = [C.c_int, C.c_int, C.c_int, _i8p, _i8p]
        lib.orc_bb_to_soft.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _f32p, _i8p]
        lib.orc_pl_rn.argtypes = [C.c_int, _u8p]
        lib.orc_pl_descramble.argtypes = [_u8p, _f32p, C.c_int, _f32p]
        lib.orc_pl_scramble.argtypes = [_u8p, _f32p, C.c_int, _f32p]
        lib.orc_ts_create.argtypes = [C.c_int]
        lib.orc_ts_create.restype = C.c_void_p
        lib.orc_ts_destroy.argtypes = [C.c_void_p]
        lib.orc_ts_work.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p, C.c_int]
        lib.orc_ts_stats.argtypes = [C.c_void_p, _u8p, ip, ip, ip, ip, ip, ip]
        lib.orc_bbheader_seal.argtypes = [_u8p]
        lib.orc_up_crc8.argtypes = [_u8p]
        lib.orc_up_crc8.restype = C.c_uint8
        lib.orc_raw_frame_size.argtypes = [C.c_int, C.c_int]
        lib.orc_plsync_create.argtypes = [C.c_int, C.c_int]
        lib.orc_plsync_create.restype = C.c_void_p
        lib.orc_plsync_destroy.argtypes = [C.c_void_p]
        lib.orc_plsync_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
        lib.orc_plsync_stats.argtypes = [C.c_void_p, ip, ip, C.POINTER(C.c_double)]
        lib.orc_plhdr_create.argtypes = [C.c_float]
        lib.orc_plhdr_create.restype = C.c_void_p
        lib.orc_plhdr_destroy.argtypes = [C.c_void_p]
        lib.orc_plhdr_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _i32p, _f32p]
        lib.orc_coarse_fed.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.orc_coarse_fed.restype = C.c_float
        lib.orc_plheader_symbols.argtypes = [C.c_int, _f32p]
        lib.orc_pls_codeword.argtypes = [C.c_int]
        lib.orc_pls_codeword.restype = C.c_uint64
        lib.orc_demod_phase_error.argtypes = [C.c_void_p, C.c_float, C.c_float]
        lib.orc_demod_phase_error.restype = C.c_float
        lib.orc_demod_phase_error_calc.argtypes = [C.c_void_p, C.c_float, C.c_float]
        lib.orc_demod_phase_error_calc.restype = C.c_float
        lib.orc_const_phase_lut.argtypes = [C.c_void_p]
        lib.orc_const_phase_lut.restype = C.POINTER(C.c_float)
        lib.orc_pll_create.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_pll_create.restype = C.c_void_p
        lib.orc_pll_destroy.argtypes = [C.c_void_p]
        lib.orc_pll_pilot_cnt.argtypes = [C.c_void_p]
        lib.orc_pll_process.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        lib.orc_dvbs_outer_create.restype = C.c_void_p
        lib.orc_dvbs_outer_destroy.argtypes = [C.c_void_p]
        lib.orc_dvbs_outer_process.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _u8p, _i32p]
        lib.orc_dvbs_outer_process.restype = None
        lib.orc_rs204_parity.argtypes = [_u8p, _u8p]
        lib.orc_rs204_parity.restype = None
        lib.orc_dvbs_deframer_create.restype = C.c_void_p
        lib.orc_dvbs_deframer_destroy.argtypes = [C.c_void_p]
        lib.orc_dvbs_deframer_work.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p]
        lib.orc_dvbs_deframer_stats.argtypes = [C.c_void_p, ip, ip]
        lib.orc_vit_create.argtypes = [C.c_float, C.c_int]
        lib.orc_vit_create.restype = C.c_void_p
        lib.orc_vit_destroy.argtypes = [C.c_void_p]
        lib.orc_vit_process.argtypes = [C.c_void_p, C.c_int, _i8p, _u8p]
        lib.orc_vit_stats.argtypes = [C.c_void_p, C.POINTER(C.c_float)] + [ip] * 5
        lib.orc_sts_create.restype = C.c_void_p
        lib.orc_sts_destroy.argtypes = [C.c_void_p]
        lib.orc_sts_process.argtypes = [C.c_void_p, C.c_int, _f32p, _i8p]
        _oracle = lib
    return _oracle


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libdvbs2_ref.so")


def have_ref():
    return os.path.exists(ref_path())


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(ref_path())
        lib.ref_ldpc_decode.argtypes = [C.c_int, C.c_int, _i8p, C.c_int]
        lib.ref_ldpc_decode_simd.argtypes = [C.c_int, C.c_int, _i8p, C.c_int]
        lib.ref_ldpc_encode_bits.argtypes = [C.c_int, C.c_int, _u8p, _u8p]
        lib.ref_ldpc_links.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_int]
        lib.ref_bch_decode.argtypes = [C.c_int, C.c_int, _u8p]
        lib.ref_bch_encode.argtypes = [C.c_int, C.c_int, _u8p]
        lib.ref_descramble.argtypes = [C.c_int, C.c_int, _u8p]
        lib.ref_deinterleave.argtypes = [C.c_int, C.c_int, C.c_int, _i8p, _i8p]
        lib.ref_interleave.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, _u8p]
        lib.ref_demap.argtypes = [C.c_int, C.c_float, C.c_float, _f32p, C.c_int, _i8p]
        lib.ref_demap_calc.argtypes = [C.c_int, C.c_float, C.c_float, _f32p, C.c_int, _i8p]
        lib.ref_mod.argtypes = [C.c_int, C.c_float, C.c_float, _u8p, C.c_int, _f32p]
        if hasattr(lib, "ref_pl_descramble"):
            lib.ref_pl_descramble.argtypes = [C.c_int, _f32p, C.c_int, _f32p, C.c_int]
        if hasattr(lib, "ref_plsync_create"):
            lib.ref_plsync_create.argtypes = [C.c_int, C.c_int]
            lib.ref_plsync_create.restype = C.c_void_p
            lib.ref_plsync_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
            lib.ref_plsync_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
            lib.ref_plhdr_create.argtypes = [C.c_float]
            lib.ref_plhdr_create.restype = C.c_void_p
            lib.ref_plhdr_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _i32p, _f32p]
            lib.ref_coarse_fed.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int]
            lib.ref_coarse_fed.restype = C.c_float
            lib.ref_plheader_symbols.argtypes = [C.c_int, _f32p]
            lib.ref_pls_codeword.argtypes = [C.c_int]
            lib.ref_pls_codeword.restype = C.c_uint64
        if hasattr(lib, "ref_pll_create"):
            lib.ref_pll_create.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
            lib.ref_pll_create.restype = C.c_void_p
            lib.ref_pll_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p]
            lib.ref_pll_pilot_cnt.argtypes = [C.c_void_p]
        if hasattr(lib, "ref_dvbs_outer_create"):
            lib.ref_dvbs_outer_create.restype = C.c_void_p
            lib.ref_dvbs_outer_process.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _u8p, _i32p]
            lib.ref_dvbs_outer_process.restype = None
            lib.ref_rs204_parity.argtypes = [_u8p, _u8p]
            lib.ref_rs204_parity.restype = None
        if hasattr(lib, "ref_dvbs_deframer_create"):
            lib.ref_dvbs_deframer_create.restype = C.c_void_p
            lib.ref_dvbs_deframer_work.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p]
            lib.ref_dvbs_deframer_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        if hasattr(lib, "ref_vit_create"):
            lib.ref_vit_create.argtypes = [C.c_float, C.c_int]
            lib.ref_vit_create.restype = C.c_void_p
            lib.ref_vit_destroy.argtypes = [C.c_void_p]
            lib.ref_vit_layout_ok.argtypes = [C.c_void_p]
            lib.ref_vit_process.argtypes = [C.c_void_p, C.c_int, _i8p, _u8p]
            lib.ref_vit_stats.argtypes = [C.c_void_p, C.POINTER(C.c_float)] + [C.POINTER(C.c_int)] * 5
            lib.ref_sts_create.restype = C.c_void_p
            lib.ref_sts_process.argtypes = [C.c_void_p, C.c_int, _f32p, _i8p]
        if hasattr(lib, "ref_ts_create"):
            lib.ref_ts_create.argtypes = [C.c_int]
            lib.ref_ts_create.restype = C.c_void_p
            lib.ref_ts_destroy.argtypes = [C.c_void_p]
            lib.ref_ts_work.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p, C.c_int]
            lib.ref_ts_stats.argtypes = [C.c_void_p, _i32p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _ref = lib
    return _ref


def code_params(short, rate):
    o = oracle()
    v = [C.c_int() for _ in range(6)]
    idx = o.orc_code_params(short, rate, *[C.byref(x) for x in v])
    if idx < 0:
        return None
    N, K, kbch, t, q, lt = [x.value for x in v]
    return dict(idx=idx, N=N, K=K, kbch=kbch, t=t, q=q, links=lt)


def encode_frame(short, rate, rng, payload=None):
    """random BBFRAME payload -> BCH -> LDPC, all through the oracle.  Returns (payload bytes, code bits)."""
    o = oracle()
    p = code_params(short, rate)
    K, kbch = p["K"], p["kbch"]
    frame = np.zeros(K // 8, np.uint8)
    if payload is None:
        payload = rng.integers(0, 256, kbch // 8, dtype=np.uint8)
    frame[: kbch // 8] = payload
    assert o.orc_bch_encode(short, rate, frame) == 0
    bits = np.unpackbits(frame)
    code = np.zeros(p["N"], np.uint8)
    assert o.orc_ldpc_encode_bits(short, rate, np.ascontiguousarray(bits), code) == 0
    return payload, code


def awgn_llr(code_bits, esn0_db, rng, scale=4.0):
    """BPSK-per-dimension AWGN LLRs in the 'L4' format of SURVEY.md 8d: clamp(rint(scale*2y/sigma^2), +-127),
    y = +-1/sqrt(2) per real dimension at Es = 1 (QPSK), bit 0 -> positive."""
    a = 1.0 / np.sqrt(2.0)
    sigma2 = 1.0 / (2.0 * 10 ** (esn0_db / 10.0))  # per real dimension, Es = 1
    y = (1.0 - 2.0 * code_bits.astype(np.float64)) * a + rng.normal(0.0, np.sqrt(sigma2), code_bits.shape)
    llr = scale * (2.0 * a * y / sigma2)  # scale x the true channel LLR
    return np.clip(np.rint(llr), -127, 127).astype(np.int8)
