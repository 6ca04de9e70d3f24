"""Transmit-side generator for PL-framed symbol streams (EN 302 307 5.5), test infrastructure: PLHEADER (SOF + PLS
code, pi/2-BPSK) + payload symbols (+ pilot blocks), PL-scrambled, with noise, a carrier offset and leading junk."""
import numpy as np

import orclib


def plheader(pls_code):
    out = np.zeros(180, np.float32)
    orclib.oracle().orc_plheader_symbols(pls_code, out)
    return out.view(np.complex64).copy()


def pl_rn(codenum):
    rn = np.zeros(131072, np.uint8)
    orclib.oracle().orc_pl_rn(codenum, rn)
    return rn


def plframe(pls_code, slots, pilots, rng, rn, bits=2):
    """one PLFRAME: 90 header symbols + slots * 90 payload symbols (+ 36-symbol pilot blocks after every 16 slots)"""
    n = slots * 90
    m = 1 << bits
    pts = np.exp(1j * (np.pi / m + 2 * np.pi * np.arange(m) / m)) if bits <= 3 else None
    pay = pts[rng.integers(0, m, n)]
    if pilots:
        out, at = [], 0
        while at < n:
            out.append(pay[at:at + 1440])
            at += 1440
            if at < n:
                out.append(np.full(36, (1 + 1j) / np.sqrt(2)))
        pay = np.concatenate(out)
    scr = np.array([1, 1j, -1, -1j])[rn[:len(pay)]]
    return np.concatenate([plheader(pls_code), (pay * scr)]).astype(np.complex64)


def stream(pls_code, slots, pilots, nframes, rng, esn0_db=12.0, lead=777, cfo=1e-4, phase=0.3, codenum=0, bits=2):
    rn = pl_rn(codenum)
    x = np.concatenate([plframe(pls_code, slots, pilots, rng, rn, bits) for _ in range(nframes)])
    junk = (rng.normal(size=lead) + 1j * rng.normal(size=lead)) / np.sqrt(2)
    x = np.concatenate([junk, x])
    n = np.arange(len(x))
    x = x * np.exp(1j * (phase + 2 * np.pi * cfo * n))
    sigma = np.sqrt(0.5 / 10 ** (esn0_db / 10))
    x = x + sigma * (rng.normal(size=len(x)) + 1j * rng.normal(size=len(x)))
    return x.astype(np.complex64)
