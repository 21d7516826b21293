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


# ---- committed fixtures (tests/golden/, generated by tests/golden/make_golden.py) ---------------------------
import hashlib
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as fh:
        return json.load(fh)


def _unhex(xs):
    return np.array([float.fromhex(x) for x in xs])


def _digest(a):
    a = np.asarray(a)
    if a.dtype.kind == "f":
        a = np.where(np.isnan(a), np.nan, a)  # one NaN bit pattern (leaves carry variance = NaN)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_doc_recorded_run_fixture_is_what_the_tests_pin():
    g = _load("doc_recorded_run")
    assert float(g["root"]["logZ"]) == -3.5950250310183574 and float(g["root"]["ESS"]) == 548622.6672945316
    assert float(g["root"]["rESS"]) == 0.5486226672945316
    assert float(g["exactLogZ"]) == pytest.approx(exact_log_z(3), rel=1e-14)
    csr = d.perfectBinaryTree(3).to_csr()
    r = o.run(csr, g["nParticles"], model=o.MODEL_MARKOV, masterRandomSeed=g["masterRandomSeed"],
              resamplingScheme=o.MULTINOMIAL, relativeEssThreshold=g["relativeEssThreshold"], rngMode=o.RNG_JAVA,
              sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM, transition=np.array(g["transition"]), prior=np.array(g["prior"]))
    assert repr(float(r.logNormEstimate[0])) == g["root"]["logZ"]
    assert repr(float(r.ess[0])) == g["root"]["ESS"]
    assert repr(float(r.relEss[0])) == g["root"]["rESS"]


def test_oracle_reproduces_committed_multilevel_fixture():
    """Guards the oracle itself: any change to the restatement that moves a bit of the small seeded run
    (statistics, log weights, ancestors, states, root population) fails here, on CPU."""
    from dcsmc_b200 import synth

    g = _load("multilevel_small")
    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows())
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    assert [int(x) for x in csr.childOffsets] == g["childOffsets"] and [int(x) for x in csr.children] == g["children"]
    assert [int(x) for x in nT] == g["nTrials"] and [int(x) for x in nS] == g["nSuccesses"]
    for case in g["cases"]:
        r = o.run(csr, g["nParticles"], masterRandomSeed=g["masterRandomSeed"], resamplingScheme=case["resamplingScheme"],
                  relativeEssThreshold=float.fromhex(case["relativeEssThreshold"]), rngMode=o.RNG_PHILOX,
                  sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
        assert np.array_equal(r.logNormEstimate, _unhex(case["logNormEstimate"]))
        assert np.array_equal(r.ess, _unhex(case["ess"]))
        assert np.array_equal(r.logScaling, _unhex(case["logScaling"]))
        assert [int(x) for x in r.resampled] == case["resampled"]
        for n in range(csr.nNodes):
            assert _digest(r.anc[n].astype(np.int32)) == case["ancestors_sha256"][n], n
            assert _digest(r.logw[n]) == case["logw_sha256"][n], n
            assert _digest(r.state[n]) == case["state_sha256"][n], n
        assert _digest(r.rootState) == case["rootState_sha256"] and _digest(r.rootWeights) == case["rootWeights_sha256"]


def test_oracle_reproduces_committed_summary_and_underflow_fixtures():
    from dcsmc_b200 import synth

    g = _load("summaries_small")
    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=11))
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    for case in g["cases"]:
        r = o.run(csr, g["nParticles"], masterRandomSeed=31, resamplingScheme=case["resamplingScheme"], rngMode=o.RNG_PHILOX,
                  sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
        for i, want in case["nodes"].items():
            s = o.node_summaries(r, csr, int(i), levelCutOffForOutput=g["levelCutOffForOutput"])
            assert s["meanMean"] == float.fromhex(want["meanMean"]) and s["meanSD"] == float.fromhex(want["meanSD"])
            assert bool(s["hasVar"]) == want["hasVar"]
            assert s["varMean"] == float.fromhex(want["varMean"]) and s["varSD"] == float.fromhex(want["varSD"])
            assert {str(c): [float(a).hex(), float(b).hex()] for c, (a, b) in s["delta"].items()} == want["delta"]
    # summaries are statistically sane: on the logistic scale the root mean is a success probability
    root = g["cases"][0]["nodes"]["0"]
    assert 0.0 < float.fromhex(root["meanMean"]) < 1.0 and float.fromhex(root["meanSD"]) > 0.0
    g = _load("wide_underflow")
    N, C_ = g["nParticles"], g["nChildren"]
    rows = [["NY", "c0", "d0", f"s{k}", "2006", str(40 + (7 * k) % 50), str(10 + (3 * k) % 25)] for k in range(C_)]
    ds = d.MultiLevelDataset(rows=rows)
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    assert (1.0 / N) ** C_ == 0.0  # the reference's linear-space product underflows here
    for case in g["cases"]:
        r = o.run(csr, N, masterRandomSeed=31, resamplingScheme=g["resamplingScheme"],
                  relativeEssThreshold=float.fromhex(case["relativeEssThreshold"]), rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                  elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
        assert np.all(np.isfinite(r.logNormEstimate))
        assert np.array_equal(r.logNormEstimate, _unhex(case["logNormEstimate"])) and np.array_equal(r.ess, _unhex(case["ess"]))
    # the fallback is consistent with a run that does not underflow: same model at N = 100 (100^-110 = 1e-220)
    small = o.run(csr, 100, masterRandomSeed=31, resamplingScheme=g["resamplingScheme"], rngMode=o.RNG_PHILOX,
                  sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
    big = _unhex(g["cases"][0]["logNormEstimate"])[0]
    assert abs(small.logNormEstimate[0] - big) < 25.0  # 110-way merge: a few nats of Monte-Carlo error at N = 100


def test_oracle_gaussian_summary_sampler_moments():
    """orc_node_summaries on a population with a known message: without the transform the node statistics
    are the mean / SD of N(mu, sigma^2) draws (polar method on the summary stream)."""
    N = 200_000
    mu, var = 0.7, 2.25

    class R:  # the fields node_summaries reads
        state = np.zeros((1, 6, N))
        anc = np.arange(N, dtype=np.int64)[None, :]

    R.state[0][0][:] = mu
    R.state[0][1][:] = var
    csr = d.perfectBinaryTree(0).to_csr() if False else None
    import types

    one = types.SimpleNamespace(nNodes=1, childOffsets=np.array([0, 0], np.int32), children=np.zeros(0, np.int32),
                                parent=np.array([-1], np.int32), nChildren=lambda i: 0, depth=lambda: np.zeros(1, np.int32))
    s = o.node_summaries(R, one, 0, levelCutOffForOutput=1, useTransform=False)
    assert abs(s["meanMean"] - mu) < 5 * math.sqrt(var / N)
    assert abs(s["meanSD"] - math.sqrt(var)) < 5 * math.sqrt(var / (2 * N))


# ---- mathematical ground truth for the parts of the path the reference never tests (fixtures 2-3) ----

@pytest.mark.parametrize("C_", [1, 2, 3, 7])
@pytest.mark.parametrize("zero_child_variance", [True, False])
def test_brownian_combine_equals_the_closed_form_marginal_likelihood(C_, zero_child_variance):
    """BrownianModelCalculator.combine (BrownianModelCalculator.java:18-41, 86-115): child c carries the message
    N(x_c; m_c, s_c) and logLik L_c; the parent value x has a flat prior and x_c | x ~ N(x, v). Folding pairwise must
    give the posterior of x (precision-weighted mean, 1 / sum of precisions) and
    logLik = sum_c L_c + log INT prod_c N(m_c; x, s_c + v) dx, which has the closed form below."""
    L = o.lib()
    rng = np.random.default_rng(100 + C_)
    m = rng.normal(0, 2, C_)
    s_ = np.zeros(C_) if zero_child_variance else rng.uniform(0.05, 3.0, C_)
    ll = rng.normal(-3, 1, C_)
    v = 0.7
    out = np.zeros(3)
    dp = C.POINTER(C.c_double)
    for mode in (o.ELEM_LIBM, o.ELEM_FDLIBM):
        L.orc_brownian_combine(mode, C_, m.ctypes.data_as(dp), s_.ctypes.data_as(dp), ll.ctypes.data_as(dp), v, out.ctypes.data_as(dp))
        if C_ == 1:  # :23-27: the branch only adds its variance
            assert out[0] == m[0] and out[1] == s_[0] + v and out[2] == ll[0]
            continue
        sig2 = s_ + v
        P = np.sum(1.0 / sig2)
        mean = np.sum(m / sig2) / P
        logint = (-0.5 * np.sum(np.log(2 * np.pi * sig2)) + 0.5 * np.log(2 * np.pi / P)
                  - 0.5 * (np.sum(m * m / sig2) - np.sum(m / sig2) ** 2 / P))
        assert out[0] == pytest.approx(mean, rel=1e-12, abs=1e-13)
        assert out[1] == pytest.approx(1.0 / P, rel=1e-12)
        assert out[2] == pytest.approx(np.sum(ll) + logint, rel=1e-12, abs=1e-12)


def test_two_leaf_binomial_tree_logz_converges_to_quadrature():
    """One parent, two binomial leaves. In the reference's accounting (MultiLevelLeafProposal.java:40: leaf weight 0,
    the Beta normaliser is knowingly dropped) the estimator targets
        Z = E[ INT N(x_1; x, v) N(x_2; x, v) dx ],  x_c = logit p_c,  p_c ~ Beta(1+s_c, 1+n_c-s_c),  v ~ Exp(rate),
    and INT Exp(v; r) N(d; 0, 2v) dv = sqrt(r)/2 * exp(-sqrt(r)|d|) (a Gaussian scale mixture is a Laplace density),
    so log Z is a smooth 2-D integral, done here by Gauss-Legendre quadrature. The oracle's logZ must converge to it."""
    from math import lgamma

    rows = [["NY", "a", "40", "13"], ["NY", "b", "25", "17"]]  # the root merges two leaves directly
    ds = d.MultiLevelDataset(rows=rows)
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    assert csr.nNodes == 3 and csr.nChildren(0) == 2
    (n1, s1), (n2, s2) = [(int(nT[i]), int(nS[i])) for i in (1, 2)]
    rate = 1.0  # MultiLevelProposalFactory.variancePrior default
    x, w = np.polynomial.legendre.leggauss(1200)
    p = 0.5 * (x + 1.0)
    w = 0.5 * w

    def beta_pdf(p_, a, b):
        return np.exp(lgamma(a + b) - lgamma(a) - lgamma(b) + (a - 1) * np.log(p_) + (b - 1) * np.log1p(-p_))

    f1 = beta_pdf(p, 1 + s1, 1 + n1 - s1) * w
    f2 = beta_pdf(p, 1 + s2, 1 + n2 - s2) * w
    assert abs(f1.sum() - 1) < 1e-10 and abs(f2.sum() - 1) < 1e-10
    lg = np.log(p) - np.log1p(-p)
    dmat = np.abs(lg[:, None] - lg[None, :])
    Z = float(f1 @ (0.5 * np.sqrt(rate) * np.exp(-np.sqrt(rate) * dmat)) @ f2)
    want = np.log(Z)
    est = []
    for seed in (31, 32, 33, 34):
        r = o.run(csr, 200_000, masterRandomSeed=seed, resamplingScheme=o.MULTINOMIAL, rngMode=o.RNG_PHILOX,
                  sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
        est.append(r.logNormEstimate[0])
    assert abs(np.mean(est) - want) < 0.02, (est, want)


def test_impl1_and_impl2_accounting_agree_on_the_same_draws():
    """SURVEY §8c fixture 5. impl-1 (DivideConquerMCAlgorithm.java:360-361, 404-406: logZ_n = sum_c logZ_c +
    LSE(lw) - log N, always resample) and impl-2 (DCRecursion.java:55-73 + ParticlePopulation: logScaling with the
    children's 1/N weights folded into the log weights, logNormEstimate = logScaling - log N) are the same
    estimator; on identical node-keyed draws every node's logZ must agree to rounding."""
    from dcsmc_b200 import synth

    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows())
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    kw = dict(resamplingScheme=o.MULTINOMIAL, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
    r2 = o.run(csr, 5000, impl=2, **kw)
    r1 = o.run(csr, 5000, impl=1, **kw)
    assert np.max(np.abs(r2.logNormEstimate - r1.logNormEstimate)) < 1e-12 * max(1.0, abs(r2.logNormEstimate[0]))
    assert np.allclose(r1.ess, r2.ess, rtol=1e-9, atol=0)  # the constant -C log N shifts log weights before exp
