"""GPU <-> oracle parity through the C ABI (include/dcsmc.h). The oracle runs the same
numerics contract (fdlibm elementary functions = StrictMath, exact sums), so everything is
compared BIT-EXACT: particle states, log weights, ancestor indices, ESS, logZ.
Tolerance for fp64: 0 ulp vs the contract oracle; 1e-9 relative vs the java-order oracle
(BASELINE.json north_star)."""
import numpy as np
import pytest

import dcsmc_b200 as d
import oracle_lib as o
from dcsmc_b200 import synth
from dcsmc_b200.engine import Engine

pytestmark = pytest.mark.gpu

T = np.array([[0.8, 0.2], [0.2, 0.8]])
PRIOR = np.array([0.5, 0.5])


def small_dataset(seed=7):
    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=seed))
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    return csr, nT, nS


def compare_all(e, ref, csr, markov=False):
    st = e.all_node_stats()
    assert np.array_equal(st["resampled"], ref.resampled)
    assert np.array_equal(st["logScaling"], ref.logScaling)
    assert np.array_equal(st["logNormEstimate"], ref.logNormEstimate)
    assert np.array_equal(st["ess"], ref.ess)
    assert np.array_equal(st["relEss"], ref.relEss)
    for n in range(csr.nNodes):
        state, istate, lw, anc = e.debug_node(n)
        assert np.array_equal(lw, ref.logw[n]), f"logW node {n}"
        assert np.array_equal(anc, ref.anc[n]), f"ancestors node {n}"
        if markov:
            assert np.array_equal(istate, ref.istate[n]), f"istate node {n}"
        else:
            assert np.array_equal(state, ref.state[n], equal_nan=True), f"state node {n}"


@pytest.mark.parametrize("N", [1, 33, 1000, 5000, 20000, 20001, 150001])  # <= 16384: shared-memory node kernel; above: generic path (odd N: dart pairs straddle the region start)
@pytest.mark.parametrize("threshold", [1.0 + 1e-10, 0.5])
@pytest.mark.parametrize("scheme", [o.MULTINOMIAL, o.STRATIFIED])
def test_multilevel_philox_bit_exact(N, threshold, scheme):
    csr, nT, nS = small_dataset()
    ref = o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, relativeEssThreshold=threshold,
                rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
    with Engine(nParticles=N, masterRandomSeed=31, resamplingScheme=scheme, relativeEssThreshold=threshold,
                retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        compare_all(e, ref, csr)
        # root population == what getRootPopulation() returns
        rs, _, rw = e.population(0)
        assert np.array_equal(rs, ref.rootState, equal_nan=True)
        assert np.array_equal(rw, ref.rootWeights)


@pytest.mark.parametrize("scheme", [o.MULTINOMIAL, o.STRATIFIED])
@pytest.mark.parametrize("duplicates", [False, True])
def test_multilevel_injected_stream_bit_exact(scheme, duplicates):
    """north_star: both sides driven by an identical injected uniform stream."""
    csr, nT, nS = small_dataset(seed=11)
    N, maxAtt = 777, 8
    rng = np.random.default_rng(31)
    inj = []
    for n in range(csr.nNodes):
        S = o.slots_per_particle(csr, n, maxBetaAttempts=maxAtt)
        u = rng.random(N * S + N)
        if duplicates:  # collisions: repeated and clustered darts (all in one sort bucket)
            u[N * S + 10:N * S + 60] = u[N * S + 5]
            u[N * S + 100:N * S + 300] = 0.5 + 1e-9 * rng.random(200)
        inj.append(u)
    ref = o.run(csr, N, resamplingScheme=scheme, rngMode=o.RNG_INJECTED, sumMode=o.SUM_EXACT,
                elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, maxBetaAttempts=maxAtt, injected=inj, dump=True)
    with Engine(nParticles=N, resamplingScheme=scheme, rngMode=1, maxBetaAttempts=maxAtt, retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        for n in range(csr.nNodes):
            assert e.slots_per_particle(n) == o.slots_per_particle(csr, n, maxBetaAttempts=maxAtt)
            e.inject_uniforms(n, inj[n])
        e.run()
        compare_all(e, ref, csr)
    # and the java-order / libm oracle (the most literal restatement of the Java) agrees within 1e-9
    lit = o.run(csr, N, resamplingScheme=scheme, rngMode=o.RNG_INJECTED, sumMode=o.SUM_JAVA,
                elemMode=o.ELEM_LIBM, nTrials=nT, nSuccesses=nS, maxBetaAttempts=maxAtt, injected=inj, dump=True)
    assert int((lit.anc != ref.anc).sum()) == 0
    assert np.allclose(lit.logNormEstimate, ref.logNormEstimate, rtol=1e-9, atol=0)
    assert np.allclose(lit.logw, ref.logw, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("scheme", [o.MULTINOMIAL, o.STRATIFIED])
def test_markov_doc_model_philox_bit_exact(scheme):
    csr = d.perfectBinaryTree(3).to_csr()
    N = 20000
    ref = o.run(csr, N, model=o.MODEL_MARKOV, resamplingScheme=scheme, relativeEssThreshold=0.5,
                rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, transition=T, prior=PRIOR, dump=True)
    with Engine(nParticles=N, resamplingScheme=scheme, relativeEssThreshold=0.5, retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_markov_model(T, PRIOR)
        e.run()
        compare_all(e, ref, csr, markov=True)
        assert abs(e.node_stats(0).logNormEstimate - (-3.600962588536195)) / 3.6 < 0.02


@pytest.mark.parametrize("scheme", [o.MULTINOMIAL, o.STRATIFIED])
def test_gauss_binary_tree_bit_exact(scheme):
    """BASELINE config 4 shape (small): perfect binary tree with observed real leaves."""
    csr, leafKind, leafObs = synth.binary_gauss_tree(5)
    N = 2000
    ref = o.run(csr, N, resamplingScheme=scheme, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                elemMode=o.ELEM_FDLIBM, leafKind=leafKind, leafObs=leafObs, dump=True)
    with Engine(nParticles=N, resamplingScheme=scheme, retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(leafKind=leafKind, leafObs=leafObs)
        e.run()
        compare_all(e, ref, csr)


def test_populations_are_released_without_retain():
    csr, nT, nS = small_dataset()
    N = 3000
    ref = o.run(csr, N, resamplingScheme=o.STRATIFIED, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
    with Engine(nParticles=N, resamplingScheme=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        st = e.all_node_stats()
        assert np.array_equal(st["logNormEstimate"], ref.logNormEstimate)
        rs, _, rw = e.population(0)
        assert np.array_equal(rs, ref.rootState, equal_nan=True)
        with pytest.raises(Exception):
            e.population(1)  # consumed by the root (DCRecursionTask.java:113)
        # running twice gives the same answer (per-node streams, Doc.java:268)
        e.run()
        assert np.array_equal(e.all_node_stats()["logNormEstimate"], ref.logNormEstimate)


def test_error_paths():
    with Engine(nParticles=10) as e:
        with pytest.raises(Exception):
            e.run()  # no tree
        csr, nT, nS = small_dataset()
        e.set_tree(csr)
        with pytest.raises(Exception):
            e.run()  # no model
        bad = nS.copy()
        bad[np.argmax(nT > 0)] = 10 ** 6  # nSuccesses > nTrials: logBinomialPr throws (DivideConquerMCAlgorithm.java:595)
        with pytest.raises(Exception):
            e.set_multilevel_model(nT, bad)


@pytest.mark.parametrize("nTrials,nSucc", [(12, 0), (9, 9), (1, 1), (30, 11), (200, 3), (97, 60)])
def test_beta_sampler_single_leaf_many_particles(nTrials, nSucc):
    """Regression: the early-exit form of the Cheng BB/BC rejection loop was mis-compiled for
    sm_100a (about 1 wrong draw in 3e4). A single-leaf tree with 2e5 particles exercises every
    accept/reject path of both samplers; all draws must equal the oracle's."""
    tree = d.DirectedTree(d.Node(("R",)))
    csr = tree.to_csr()
    nT = np.array([nTrials], np.int32)
    nS = np.array([nSucc], np.int32)
    N = 200_000
    ref = o.run(csr, N, resamplingScheme=o.STRATIFIED, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
    with Engine(nParticles=N, resamplingScheme=1, retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        state, _, lw, anc = e.debug_node(0)
        assert np.array_equal(state, ref.state[0], equal_nan=True)
        assert np.array_equal(anc, ref.anc[0])
        assert e.node_stats(0).logNormEstimate == ref.logNormEstimate[0] == 0.0


def test_device_elementary_functions_equal_fdlibm_oracle():
    """dc_log / dc_exp on the device == the oracle's fdlibm restatement (StrictMath), bit for bit."""
    L = o.lib()
    rng = np.random.default_rng(5)
    with Engine(nParticles=8) as e:
        xs = np.concatenate([rng.uniform(-745, 709, 100000), rng.uniform(-2, 2, 100000), rng.normal(0, 1e-3, 20000),
                             rng.normal(0, 1e-9, 20000),
                             np.array([0.0, -0.0, 1e-300, 709.78, -745.13, -746, 710, np.inf, -np.inf, 0.34657359027997264,
                                       -0.34657359027997264, 1.0397207708399179, -1.0397207708399179])])
        yd = e.debug_eval(1, xs)
        yo = np.array([L.orc_elem_exp(o.ELEM_FDLIBM, float(x)) for x in xs])
        assert np.array_equal(yd, yo, equal_nan=True)
        ys = np.concatenate([np.exp(rng.uniform(-700, 700, 100000)), rng.uniform(0.5, 2.0, 100000), rng.uniform(0, 1, 100000),
                             1 + rng.normal(0, 1e-7, 20000), np.array([0.0, 1.0, 5e-324, 1e-310, 2.0, 0.5, np.inf, -1.0])])
        yd = e.debug_eval(0, ys)
        yo = np.array([L.orc_elem_log(o.ELEM_FDLIBM, float(x)) for x in ys])
        assert np.array_equal(yd, yo, equal_nan=True)
        idx = np.arange(0, 50000, dtype=np.float64)
        ud = e.debug_eval(2 | (7 << 8), idx)
        uo = np.array([L.orc_philox_uniform(31, 7, int(i)) for i in idx])
        assert np.array_equal(ud, uo)


def test_straight_line_divisions_and_elementary_functions_are_ieee_exact():
    """The hot kernels do not use nvcc's '/' everywhere: a quotient over a shared refined reciprocal (div_by /
    rcp_refined, the merge kernels' folds) and dc_div_safe (the leaf kernels' fdlibm) run the compiler's fast-path
    sequence without its range checks. Inside the operand range the kernels guarantee, both must return the correctly
    rounded IEEE quotient — checked here against numpy's division bit for bit on random and adversarial operands
    (mantissas of all ones / one ulp off a power of two, equal operands, operands 2^+-200 apart). Likewise the
    straight-line log / exp equal the general routines on every non-rare argument."""
    rng = np.random.default_rng(11)
    n = 400000

    def rand_operands(k, emin, emax):
        mant = rng.integers(0, 1 << 52, k, dtype=np.uint64)
        expo = rng.integers(1023 + emin, 1023 + emax, k, dtype=np.uint64)
        sign = rng.integers(0, 2, k, dtype=np.uint64) << np.uint64(63)
        return (sign | (expo << np.uint64(52)) | mant).view(np.float64)

    a = rand_operands(n, -200, 200)
    b = rand_operands(n, -200, 200)
    edge = np.array([1.0, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), np.nextafter(2.0, 0.0), 3.0, 1.0 / 3.0, 1e-60, 1e60,
                     0.1, 7.0, np.nextafter(4.0, 0.0), 1.5])
    ea, eb = np.meshgrid(edge, edge)
    a = np.concatenate([a, ea.ravel(), -ea.ravel(), np.zeros(8)])
    b = np.concatenate([b, eb.ravel(), eb.ravel(), edge[:8]])
    # the exact ties of the final correction step: power-of-two dividends over divisors whose mantissa is all ones
    # (1 / (1 - 2^-53) must round to 1 + 2^-52); the reciprocal seed's low word = 1, as in nvcc's own expansion, decides them
    ones = (np.uint64(0x000FFFFFFFFFFFFF) | (rng.integers(1023 - 150, 1023 + 150, 4000, dtype=np.uint64) << np.uint64(52))).view(np.float64)
    pow2 = np.ldexp(1.0, rng.integers(-150, 150, 4000)) * rng.choice([-1.0, 1.0], 4000)
    a = np.concatenate([a, pow2, rand_operands(4000, -150, 150), ones])
    b = np.concatenate([b, ones, ones, np.abs(pow2)])
    x = np.empty(2 * len(a))
    x[0::2], x[1::2] = a, b
    with Engine(nParticles=8) as e:
        y = e.debug_eval(3, x)
        want = a / b
        assert np.array_equal(y[0::2], want), "quotient over a shared refined reciprocal"
        assert np.array_equal(y[1::2], want), "dc_div_safe"
        # straight-line log / exp against the general routines (arguments away from the rare sets: |x - 1| >= 2^-20,
        # normal range; 2^-28 <= |x| < 708)
        xs = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(1e-12, 0.999, 200000), rng.uniform(1.001, 50, 100000)])
        hm = (xs.view(np.uint64) >> np.uint64(32)).astype(np.int64) & 0xFFFFF  # dc_log_is_rare's mantissa test
        xs = xs[((hm + 2) & 0xFFFFF) >= 3]
        assert np.array_equal(e.debug_eval(4, xs), e.debug_eval(0, xs))
        xe = np.concatenate([rng.uniform(-707, 707, 200000), rng.uniform(-3, 3, 200000)])
        xe = xe[np.abs(xe) >= 2.0 ** -28]
        assert np.array_equal(e.debug_eval(5, xe), e.debug_eval(1, xe))


@pytest.mark.parametrize("scheme", [o.MULTINOMIAL, o.STRATIFIED])
@pytest.mark.parametrize("threshold", [1.0 + 1e-10, 0.5])
def test_java_order_sums_bit_exact_against_java_order_oracle(scheme, threshold):
    """DCSMC_SUM_JAVA: the device runs the reference's own summation order (sequential fp64 S, W_i =
    e_i/S, compensated sum of squares, running-sum CDF, unscaled darts). Against the oracle in
    java-order mode everything is bit-identical — no 'device-order' caveat at all."""
    csr, nT, nS = small_dataset(seed=13)
    N = 3000
    ref = o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, relativeEssThreshold=threshold,
                rngMode=o.RNG_PHILOX, sumMode=o.SUM_JAVA, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
    with Engine(nParticles=N, masterRandomSeed=31, resamplingScheme=scheme, relativeEssThreshold=threshold,
                sumMode=1, retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        compare_all(e, ref, csr)
        rs, _, rw = e.population(0)
        assert np.array_equal(rs, ref.rootState, equal_nan=True)
        assert np.array_equal(rw, ref.rootWeights)


def test_gpu_reproduces_the_reference_recorded_run_to_every_digit():
    """The reference's only golden vector (Doc.java:282 = README.md:257), on the GPU:
    nParticles=1e6, masterRandomSeed=31, MULTINOMIAL, relativeEssThreshold=0.5, perfect binary tree
    of depth 3. java.util.Random regenerated on the device (DCSMC_RNG_JAVA) + the reference's
    summation order (DCSMC_SUM_JAVA) give
        ESS=548622.6672945316, rESS=0.5486226672945316, logZ=-3.5950250310183574."""
    csr = d.perfectBinaryTree(3).to_csr()
    with Engine(nParticles=1_000_000, masterRandomSeed=31, resamplingScheme=0, relativeEssThreshold=0.5,
                rngMode=2, sumMode=1) as e:
        e.set_tree(csr)
        e.set_markov_model(T, PRIOR)
        e.run()
        root = e.node_stats(0)
        assert root.logNormEstimate == -3.5950250310183574
        assert root.ess == 548622.6672945316
        assert root.relEss == 0.5486226672945316
        names = [n.toString() for n in csr.nodes]
        st = e.all_node_stats()
        assert sorted(names[i] for i in np.nonzero(st["resampled"])[0]) == ["0-0", "0-1"]


@pytest.mark.parametrize("threshold", [1.0 + 1e-10, 0.5])
def test_wide_node_weight_product_underflow(threshold):
    """The reference multiplies the children's weights in linear space (DCRecursion.java:55-63); with C
    children that is N^-C, which underflows to 0 for C*log10(N) > 323 (54 children at nParticles = 1e6,
    BASELINE config 3) and makes every log weight -inf. Engine and oracle fall back to the sum of logs
    exactly there (documented deviation): logZ stays finite, and both still agree bit for bit.
    threshold > 1: children resampled (per-node constant); 0.5: leaves arrive weighted (per particle)."""
    N, C = 2000, 110  # 2000^-110 = 1e-363
    rows = [["NY", "c0", "d0", f"s{k}", "2006", str(40 + (7 * k) % 50), str(10 + (3 * k) % 25)] for k in range(C)]
    ds = d.MultiLevelDataset(rows=rows)
    tree = ds.getTree()
    csr = tree.to_csr()
    nT, nS = ds.leaf_arrays(csr)
    wide = [i for i in range(csr.nNodes) if csr.nChildren(i) >= C]
    assert wide, "fixture must contain a node with >= 110 children"
    assert (1.0 / N) ** C == 0.0
    ref = o.run(csr, N, masterRandomSeed=31, resamplingScheme=o.STRATIFIED, relativeEssThreshold=threshold,
                rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
    assert np.all(np.isfinite(ref.logNormEstimate))
    with Engine(nParticles=N, masterRandomSeed=31, resamplingScheme=o.STRATIFIED, relativeEssThreshold=threshold,
                retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        compare_all(e, ref, csr)


def test_engine_matches_committed_golden_fixture():
    """tests/golden/multilevel_small.json (made by tests/golden/make_golden.py): the engine reproduces the
    committed vectors bit for bit — statistics, and SHA-256 of every node's log weights, ancestors, states."""
    import hashlib
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multilevel_small.json")))
    csr, nT, nS = small_dataset()
    def dig(a):
        a = np.asarray(a)
        if a.dtype.kind == "f":
            a = np.where(np.isnan(a), np.nan, a)  # one NaN bit pattern (leaves carry variance = NaN)
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    unhex = lambda xs: np.array([float.fromhex(x) for x in xs])
    for case in g["cases"]:
        with Engine(nParticles=g["nParticles"], masterRandomSeed=g["masterRandomSeed"], resamplingScheme=case["resamplingScheme"],
                    relativeEssThreshold=float.fromhex(case["relativeEssThreshold"]), retainPopulations=1) as e:
            e.set_tree(csr)
            e.set_multilevel_model(nT, nS)
            e.run()
            st = e.all_node_stats()
            assert np.array_equal(st["logNormEstimate"], unhex(case["logNormEstimate"]))
            assert np.array_equal(st["ess"], unhex(case["ess"]))
            assert np.array_equal(st["logScaling"], unhex(case["logScaling"]))
            assert [int(x) for x in st["resampled"]] == case["resampled"]
            for n in range(csr.nNodes):
                state, _, lw, anc = e.debug_node(n)
                assert dig(anc.astype(np.int32)) == case["ancestors_sha256"][n], n
                assert dig(lw) == case["logw_sha256"][n], n
                assert dig(state) == case["state_sha256"][n], n
            rs, _, rw = e.population(0)
            assert dig(rs) == case["rootState_sha256"] and dig(rw) == case["rootWeights_sha256"]


def ragged_school_rows(seed=11, counts=(1, 2, 3, 8, 9, 17, 40)):
    """R -> school -> leaf with ragged leaf counts: one child (combine's C == 1 branch), exactly one and more than one
    small and large fan-outs, more than one tile of child descriptors (32)."""
    rng = np.random.default_rng(seed)
    rows = []
    for s, c in enumerate(counts):
        for l in range(c):
            n = int(20 + rng.poisson(60))
            rows.append(["R", f"s{s}", f"l{l}", str(n), str(int(rng.binomial(n, rng.beta(2, 2))))])
    return rows


@pytest.mark.parametrize("scheme", [o.MULTINOMIAL, o.STRATIFIED])
def test_kernel_variants_are_bit_identical_to_the_oracle(scheme, monkeypatch):
    """Every selectable kernel variant (merge kernels for the level above the leaves incl. the shared-reciprocal fold
    with its literal fallback forced for most / all folds, leaf kernels, scans, dart kernels) against the oracle."""
    ds = d.MultiLevelDataset(rows=ragged_school_rows())
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    N = 20011  # above the shared-memory node kernels, not a multiple of the block size
    ref = o.run(csr, N, masterRandomSeed=5, resamplingScheme=scheme, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
    knobs = ("DCSMC_LEAFKIDS", "DCSMC_LK2_RANGE", "DCSMC_LEAF_V", "DCSMC_SCAN_V", "DCSMC_PARTITION_DARTS", "DCSMC_L2_PREFETCH")
    variants = [{}] + [{"DCSMC_LEAFKIDS": str(v)} for v in range(3)]
    variants += [{"DCSMC_LEAFKIDS": "2", "DCSMC_LK2_RANGE": "1"}, {"DCSMC_LEAFKIDS": "2", "DCSMC_LK2_RANGE": "0"}]
    variants += [{"DCSMC_LEAF_V": "1"}, {"DCSMC_LEAF_V": "2"}, {"DCSMC_SCAN_V": "1"}, {"DCSMC_PARTITION_DARTS": "0"}]
    variants += [{"DCSMC_L2_PREFETCH": "0"}, {"DCSMC_L2_PREFETCH": "0", "DCSMC_LEAFKIDS": "0"}]
    for env in variants:
        for k in knobs:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Engine(nParticles=N, masterRandomSeed=5, resamplingScheme=scheme, retainPopulations=1) as e:
            e.set_tree(csr)
            e.set_multilevel_model(nT, nS)
            e.run()
            compare_all(e, ref, csr)
