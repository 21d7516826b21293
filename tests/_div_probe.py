"""Scratch helper (not a test): where do the straight-line division forms differ from IEEE a / b?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dcsmc_b200.engine import Engine

rng = np.random.default_rng(11)
def rand_operands(k, emin, emax, signed=True):
    mant = rng.integers(0, 1 << 52, k, dtype=np.uint64)
    expo = rng.integers(1023 + emin, 1023 + emax, k, dtype=np.uint64)
    sign = (rng.integers(0, 2, k, dtype=np.uint64) << np.uint64(63)) if signed else np.uint64(0)
    return (sign | (expo << np.uint64(52)) | mant).view(np.float64)

edge = np.array([1.0, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), np.nextafter(2.0, 0.0), 3.0, 1.0 / 3.0, 1e-60, 1e60,
                 0.1, 7.0, np.nextafter(4.0, 0.0), 1.5])
ea, eb = np.meshgrid(edge, edge)
with Engine(nParticles=8) as e:
    allones = (np.uint64(0x000FFFFFFFFFFFFF) | (rng.integers(1023 - 30, 1023 + 30, 2000, dtype=np.uint64) << np.uint64(52))).view(np.float64)
    for name, a, b in (("b mantissa all ones, random a", rand_operands(2000, -30, 30), allones),
                       ("edge grid", np.concatenate([ea.ravel(), -ea.ravel(), np.zeros(8)]), np.concatenate([eb.ravel(), eb.ravel(), edge[:8]])),
                       ("random +-200", rand_operands(400000, -200, 200), rand_operands(400000, -200, 200)),
                       ("random +-20", rand_operands(400000, -20, 20), rand_operands(400000, -20, 20)),
                       ("positive +-200", rand_operands(400000, -200, 200, False), rand_operands(400000, -200, 200, False))):
        x = np.empty(2 * len(a)); x[0::2], x[1::2] = a, b
        y = e.debug_eval(3, x)
        want = a / b
        y6 = e.debug_eval(6, x)
        for lab, got in (("div_by", y[0::2]), ("div_safe", y[1::2]), ("nvcc /", y6[0::2])):
            bad = np.nonzero(got != want)[0]
            print(name, lab, "mismatches", len(bad), "of", len(a))
            for i in bad[:6]:
                print("   a=%r b=%r got=%r want=%r  ulps=%d" % (a[i].hex(), b[i].hex(), got[i].hex(), want[i].hex(),
                      int(got[i].view(np.int64)) - int(want[i].view(np.int64))))
