"""Pins the oracle against the only golden vectors the reference holds for this path:
the recorded run in src/test/java/dc/Doc.java:282 (= README.md:257) and the exact logZ of
Doc.java:287, plus the JDK / Random123 known answers for the generators."""
import ctypes as C
import math

import numpy as np
import pytest

import dcsmc_b200 as d
import oracle_lib as o

TRANSITION = np.array([[0.8, 0.2], [0.2, 0.8]])  # Doc.java:228-231
PRIOR = np.array([0.5, 0.5])  # Doc.java:234


def exact_log_z(depth: int) -> float:
    """Sum-product on the perfect binary tree with all leaves observed = 0
    (TestUtilities.computeExactLogZ, TestUtilities.java:40-85)."""

    def up(level):
        if level == depth:
            return np.array([1.0, 0.0])
        m = up(level + 1)
        msg = TRANSITION @ m
        return msg * msg

    return float(np.log(np.sum(PRIOR * up(0))))


def test_exact_logz_matches_doc_comment():
    assert exact_log_z(3) == pytest.approx(-3.600962588536195, abs=1e-14)  # Doc.java:287


@pytest.mark.slow
def test_doc_recorded_run_bit_for_bit():
    """Doc.java:282: node=0, ESS=548622.6672945316, rESS=0.5486226672945316,
    logZ=-3.5950250310183574 (nParticles=1e6, seed 31, MULTINOMIAL, threshold 0.5)."""
    t = d.perfectBinaryTree(3).to_csr()
    r = o.run(t, 1_000_000, model=o.MODEL_MARKOV, masterRandomSeed=31, resamplingScheme=o.MULTINOMIAL,
              relativeEssThreshold=0.5, transition=TRANSITION, prior=PRIOR,
              rngMode=o.RNG_JAVA, sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM)
    assert r.logNormEstimate[0] == -3.5950250310183574
    assert r.ess[0] == 548622.6672945316
    assert r.relEss[0] == 0.5486226672945316
    # the reference's own assertion (Doc.java:289-290)
    exact = exact_log_z(3)
    assert abs(exact - r.logNormEstimate[0]) / abs(exact) < 0.01
    # level-1 nodes are the only ones below the 0.5 threshold
    names = [n.toString() for n in t.nodes]
    assert sorted(names[i] for i in np.nonzero(r.resampled)[0]) == ["0-0", "0-1"]
    # fdlibm elementary functions (StrictMath) give the same digits on this run
    r2 = o.run(t, 1_000_000, model=o.MODEL_MARKOV, masterRandomSeed=31, resamplingScheme=o.MULTINOMIAL,
               relativeEssThreshold=0.5, transition=TRANSITION, prior=PRIOR,
               rngMode=o.RNG_JAVA, sumMode=o.SUM_JAVA, elemMode=o.ELEM_FDLIBM)
    assert r2.logNormEstimate[0] == -3.5950250310183574
    assert r2.ess[0] == 548622.6672945316


def test_doc_model_threads_do_not_change_result():
    """Doc.java:268: given the seed, output is deterministic even if parallelised."""
    t = d.perfectBinaryTree(3).to_csr()
    kw = dict(model=o.MODEL_MARKOV, masterRandomSeed=31, relativeEssThreshold=0.5,
              transition=TRANSITION, prior=PRIOR)
    a = o.run(t, 20_000, **kw)
    b = o.run(t, 20_000, nThreads=4, maximumDistributionDepth=2, **kw)
    assert np.array_equal(a.logNormEstimate, b.logNormEstimate)
    assert np.array_equal(a.ess, b.ess)


def test_exact_sum_mode_agrees_with_java_order():
    t = d.perfectBinaryTree(3).to_csr()
    kw = dict(model=o.MODEL_MARKOV, masterRandomSeed=31, relativeEssThreshold=0.5,
              transition=TRANSITION, prior=PRIOR, dump=True)
    a = o.run(t, 100_000, **kw)
    b = o.run(t, 100_000, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, **kw)
    assert int((a.anc != b.anc).sum()) == 0
    assert np.allclose(a.logNormEstimate, b.logNormEstimate, rtol=1e-13, atol=0)
    assert np.allclose(a.ess, b.ess, rtol=1e-12, atol=0)


@pytest.mark.parametrize("depth,n", [(3, 40_000), (5, 20_000)])
def test_doc_model_statistical(depth, n):
    """DiscreteExample.java:29-82 style check: logZ estimate converges to sum-product."""
    t = d.perfectBinaryTree(depth).to_csr()
    exact = exact_log_z(depth)
    for scheme in (o.MULTINOMIAL, o.STRATIFIED):
        r = o.run(t, n, model=o.MODEL_MARKOV, masterRandomSeed=31, resamplingScheme=scheme,
                  relativeEssThreshold=0.5, transition=TRANSITION, prior=PRIOR)
        assert abs(exact - r.logNormEstimate[0]) / abs(exact) < 0.02


# ---- generators -------------------------------------------------------------------------

def test_java_random_known_answers():
    L = o.lib()
    s = C.c_uint64(L.orc_java_seed_scramble(42))
    assert L.orc_java_next(C.byref(s), 32) == -1170105035  # new Random(42).nextInt()
    s = C.c_uint64(L.orc_java_seed_scramble(0))
    assert L.orc_java_next_double(C.byref(s)) == 0.730967787376657  # new Random(0).nextDouble()
    s = C.c_uint64(L.orc_java_seed_scramble(0))
    assert [L.orc_java_next_int(C.byref(s), 10) for _ in range(5)] == [0, 8, 9, 7, 5]  # new Random(0).nextInt(10)


def test_java_hashing():
    L = o.lib()
    assert L.orc_java_string_hash(b"hello") == 99162322
    assert d.java_string_hash("hello") == 99162322
    assert d.java_list_hash(["0"]) == 31 + 48
    n = d.Node(("0", "1", "0"))
    assert n.hashCode() == d.tree._i32(31 + d.java_list_hash(["0", "1", "0"]))
    # seed = 31*(31*1+master) + hash in long arithmetic (DCRecursionTask.java:98-106)
    assert L.orc_node_seed(31, 110) == 31 * (31 + 31) + 110
    assert L.orc_node_seed(31, -5) == 31 * (31 + 31) - 5


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    L = o.lib()

    def ph(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        out = (C.c_uint32 * 4)()
        L.orc_philox4x32_10(c, k, out)
        return list(out)

    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert ph([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert ph([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = [L.orc_philox_uniform(31, 7, i) for i in range(1000)]
    assert all(0.0 <= x < 1.0 for x in u) and 0.4 < sum(u) / len(u) < 0.6


def test_fdlibm_elementary_functions_within_one_ulp_of_libm():
    L = o.lib()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-2, 2, 20000), rng.normal(0, 1e-3, 5000)])
    for x in xs:
        a, b = L.orc_elem_exp(o.ELEM_FDLIBM, float(x)), math.exp(float(x))
        assert abs(a - b) <= np.spacing(b) * 1.0, (x, a, b)
    ys = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2.0, 20000), rng.uniform(0, 1, 20000)])
    for y in ys:
        a, b = L.orc_elem_log(o.ELEM_FDLIBM, float(y)), math.log(float(y))
        assert abs(a - b) <= np.spacing(abs(b)) * 1.0 + 0.0, (y, a, b)
    assert L.orc_elem_log(o.ELEM_FDLIBM, 0.0) == -math.inf
    assert math.isnan(L.orc_elem_log(o.ELEM_FDLIBM, -1.0))
    assert L.orc_elem_log(o.ELEM_FDLIBM, 1.0) == 0.0
    assert L.orc_elem_exp(o.ELEM_FDLIBM, 0.0) == 1.0
    assert L.orc_elem_exp(o.ELEM_FDLIBM, -math.inf) == 0.0
    assert L.orc_elem_exp(o.ELEM_FDLIBM, -800.0) == 0.0
    assert L.orc_elem_log(o.ELEM_FDLIBM, 5e-324) == math.log(5e-324)
