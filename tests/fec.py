"""Test helper: import the product package (its directory name has hyphens)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")

# (modcod, reference rate enum) for the QPSK MODCODs: every LDPC/BCH code appears once here
QPSK_MODCOD_OF_RATE = {0: 1, 1: 2, 2: 3, 3: 4, 4: 5, 5: 6, 6: 7, 7: 8, 8: 9, 10: 10, 11: 11}
# MODCOD -> (constellation index 0..3, oracle constellation type, rate enum, g1, g2)
MODCODS = {}
for _r, _m in QPSK_MODCOD_OF_RATE.items():
    MODCODS[_m] = (0, 1, _r, 0.0, 0.0)
for _mc, _r in zip(range(12, 18), (4, 5, 6, 8, 10, 11)):
    MODCODS[_mc] = (1, 3, _r, 0.0, 0.0)
for _mc, _r, _g in zip(range(18, 24), (5, 6, 7, 8, 10, 11), (3.15, 2.85, 2.75, 2.70, 2.60, 2.57)):
    MODCODS[_mc] = (2, 4, _r, _g, 0.0)
for _mc, _r, _g1, _g2 in zip(range(24, 29), (6, 7, 8, 10, 11), (2.84, 2.72, 2.64, 2.54, 2.53), (5.27, 4.87, 4.64, 4.33, 4.30)):
    MODCODS[_mc] = (3, 5, _r, _g1, _g2)
