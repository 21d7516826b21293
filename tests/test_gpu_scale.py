"""Parity at sizes the oracle cannot reach in seconds, through size-independent properties, plus the
memory-constrained schedule (subtree batches) and determinism across schedules."""
import numpy as np
import pytest

import dcsmc_b200 as d
import oracle_lib as o
from dcsmc_b200 import synth
from dcsmc_b200.engine import Engine

pytestmark = pytest.mark.gpu


def dataset(rows):
    ds = d.MultiLevelDataset(rows=rows)
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    return csr, nT, nS


@pytest.mark.parametrize("scheme", [0, 1])
def test_small_memory_budget_forces_subtree_batches_same_result(scheme):
    """The static memory plan falls back to subtree batches when a level does not fit; results must
    not depend on the schedule (per-node streams, Doc.java:268) and must equal the oracle."""
    csr, nT, nS = dataset(synth.wide_rows(11, [4, 5, 6]))
    N = 4096
    ref = o.run(csr, N, resamplingScheme=scheme, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM,
                nTrials=nT, nSuccesses=nS)
    with Engine(nParticles=N, resamplingScheme=scheme) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        full = e.all_node_stats()
        fullBatches = e.run_info().levelBatches
        fullPool = e.run_info().poolBytes
    assert np.array_equal(full["logNormEstimate"], ref.logNormEstimate)
    # budget: about a quarter of "everything resident"
    with Engine(nParticles=N, resamplingScheme=scheme, memoryBudgetBytes=int(fullPool // 4)) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        small = e.all_node_stats()
        assert e.run_info().levelBatches > fullBatches
        assert e.run_info().poolBytes <= fullPool // 4
        rs, _, rw = e.population(0)
    for k in ("ess", "relEss", "logNormEstimate", "logScaling", "resampled"):
        assert np.array_equal(small[k], full[k]), k
    assert np.array_equal(rs, ref.rootState, equal_nan=True)
    # far too small: a clean error, not a crash
    with Engine(nParticles=N, resamplingScheme=scheme, memoryBudgetBytes=1 << 16) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        with pytest.raises(Exception) as ei:
            e.run()
        assert "pool" in str(ei.value)


@pytest.mark.parametrize("scheme", [0, 1])
def test_million_particles_invariants(scheme):
    """nParticles = 1e6 (BASELINE config 3 size) on a small tree: ancestors sorted and in range,
    offspring counts sum to N, ESS in (0, N], logZ matches an independent 1e5-particle run within
    Monte-Carlo error, and the exact-sum statistics are reproduced by numpy from the raw weights."""
    csr, nT, nS = dataset(synth.small_multilevel_rows(seed=5, nGroups=2, edge_cases=False))
    N = 1_000_000
    with Engine(nParticles=N, resamplingScheme=scheme, retainPopulations=1) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        st = e.all_node_stats()
        assert np.all(st["done"] == 1) and np.all(st["resampled"] == 1)
        assert np.all(st["ess"] > 0) and np.all(st["ess"] <= N * (1 + 1e-12))
        for n in range(csr.nNodes):
            state, _, lw, anc = e.debug_node(n)
            assert anc.min() >= 0 and anc.max() < N
            assert np.all(np.diff(anc) >= 0), "output must be in ascending-ancestor order (SMCUtils sweep)"
            cnt = np.bincount(anc, minlength=N)
            assert cnt.sum() == N
            if csr.nChildren(n) > 0:
                # logScaling = sum_c logScaling_c + LSE(logW)  (DCRecursion.java:68-73)
                m = lw.max()
                lse = m + np.log(np.sum(np.exp(lw - m)))
                base = sum(st["logScaling"][c] for c in csr.children[csr.childOffsets[n]:csr.childOffsets[n + 1]])
                assert st["logScaling"][n] == pytest.approx(base + lse, rel=1e-12)
                w = np.exp(lw - m)
                assert st["ess"][n] == pytest.approx(w.sum() ** 2 / np.sum(w * w), rel=1e-10)
                # heavy particles get offspring in proportion to their weight (5 sigma)
                p = w / w.sum()
                top = np.argsort(p)[-5:]
                for i in top:
                    assert abs(cnt[i] - N * p[i]) < 5 * np.sqrt(N * p[i]) + 2
            else:
                assert st["logNormEstimate"][n] == 0.0 and st["ess"][n] == N
        big = st["logNormEstimate"][0]
    with Engine(nParticles=100_000, resamplingScheme=scheme, masterRandomSeed=77) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        small = e.node_stats(0).logNormEstimate
    assert abs(big - small) < 0.15


def test_runs_are_deterministic_and_seed_sensitive():
    csr, nT, nS = dataset(synth.small_multilevel_rows(seed=9))
    outs = []
    for seed in (31, 31, 32):
        with Engine(nParticles=20_000, masterRandomSeed=seed) as e:
            e.set_tree(csr)
            e.set_multilevel_model(nT, nS)
            e.run()
            outs.append(e.all_node_stats()["logNormEstimate"].copy())
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


def test_nys_config1_matches_oracle_statistics():
    """BASELINE config 1 shape: NYS-shaped CSV, nParticles = 1000: every node's statistics equal the
    contract oracle's, and the java-order/libm oracle (the literal restatement) within 1e-9."""
    csr, nT, nS = dataset(synth.nys_shaped_rows())
    N = 1000
    with Engine(nParticles=N) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.run()
        st = e.all_node_stats()
    ref = o.run(csr, N, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS,
                nThreads=8, maximumDistributionDepth=2)
    assert np.array_equal(st["logNormEstimate"], ref.logNormEstimate)
    assert np.array_equal(st["ess"], ref.ess)
    lit = o.run(csr, N, rngMode=o.RNG_PHILOX, sumMode=o.SUM_JAVA, elemMode=o.ELEM_LIBM, nTrials=nT, nSuccesses=nS,
                nThreads=8, maximumDistributionDepth=2)
    assert np.allclose(st["logNormEstimate"], lit.logNormEstimate, rtol=1e-9, atol=1e-9)
    assert np.allclose(st["ess"], lit.ess, rtol=1e-9)


@pytest.mark.parametrize("scheme", [0, 1])
def test_observed_leaves_not_materialised_same_results(scheme):
    """Observed (Gaussian) leaves are N copies of one datum; without retainPopulations they get no
    slot and no kernel work. Every node statistic and the root population must still equal the
    oracle's (which does materialise and resample them) — also when binomial and observed leaves
    share a parent."""
    csr, leafKind, leafObs = synth.binary_gauss_tree(6)
    N = 3000
    ref = o.run(csr, N, resamplingScheme=scheme, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM,
                leafKind=leafKind, leafObs=leafObs)
    with Engine(nParticles=N, resamplingScheme=scheme) as e:
        e.set_tree(csr)
        e.set_multilevel_model(leafKind=leafKind, leafObs=leafObs)
        e.run()
        st = e.all_node_stats()
        for k in ("ess", "relEss", "logNormEstimate", "logScaling", "resampled"):
            assert np.array_equal(st[k], getattr(ref, k)), k
        rs, _, rw = e.population(0)
        assert np.array_equal(rs, ref.rootState, equal_nan=True)
    # mixed leaves
    csr, nT, nS = dataset(synth.small_multilevel_rows(seed=2))
    leafKind = np.zeros(csr.nNodes, np.int32)
    leafObs = np.zeros(csr.nNodes)
    leaves = [i for i in range(csr.nNodes) if csr.nChildren(i) == 0]
    for i in leaves[::2]:
        leafKind[i] = 1
        leafObs[i] = 0.1 * i - 0.7
    ref = o.run(csr, N, resamplingScheme=scheme, relativeEssThreshold=0.7, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, leafKind=leafKind, leafObs=leafObs)
    with Engine(nParticles=N, resamplingScheme=scheme, relativeEssThreshold=0.7) as e:
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS, leafKind, leafObs)
        e.run()
        st = e.all_node_stats()
        for k in ("ess", "logNormEstimate", "resampled"):
            assert np.array_equal(st[k], getattr(ref, k)), k
        rs, _, rw = e.population(0)
        assert np.array_equal(rs, ref.rootState, equal_nan=True)
        assert np.array_equal(rw, ref.rootWeights)


@pytest.mark.parametrize("scheme", [0, 1])
def test_graph_replay_and_chunked_streams_do_not_change_results(scheme, monkeypatch):
    """A cached plan is replayed as a CUDA graph from its second execution on; DCSMC_OVERLAP_CHUNKS
    splits every level over two streams. Neither may change a single bit (per-node streams, exact sums)."""
    csr, nT, nS = dataset(synth.wide_rows(7, [5, 6, 4]))
    N = 30000

    def run(env):
        for k in ("DCSMC_GRAPHS", "DCSMC_OVERLAP_CHUNKS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Engine(nParticles=N, resamplingScheme=scheme, relativeEssThreshold=0.7) as e:
            e.set_tree(csr)
            e.set_multilevel_model(nT, nS)
            out = []
            for _ in range(3):  # 1st: planned, 2nd: graph captured, 3rd: graph replayed
                e.run()
                st = e.all_node_stats()
                rs, _, rw = e.population(0)
                out.append((st, rs, rw, e.run_info().kernelLaunches))
            return out

    base = run({"DCSMC_GRAPHS": "0"})
    graphs = run({"DCSMC_GRAPHS": "1"})
    chunks = run({"DCSMC_OVERLAP_CHUNKS": "4"})
    for other in (graphs, chunks):
        for (st0, rs0, rw0, _), (st1, rs1, rw1, _) in zip(base, other):
            for k in st0:
                assert np.array_equal(st0[k], st1[k]), k
            assert np.array_equal(rs0, rs1, equal_nan=True) and np.array_equal(rw0, rw1)
    assert graphs[0][3] == graphs[2][3] == base[0][3] > 0  # launch accounting survives the replay


def test_engine_stream_timer_and_export_error_paths():
    from dcsmc_b200 import _ffi

    csr, nT, nS = dataset(synth.small_multilevel_rows())
    with Engine(nParticles=2048) as e:
        with pytest.raises(_ffi.DcsmcError):
            e.timer_end()  # no timer_begin yet
        e.set_tree(csr)
        e.set_multilevel_model(nT, nS)
        e.timer_begin()
        e.run()
        e.run()
        ms = e.timer_end()
        assert ms >= e.run_info().deviceMs > 0
        desc = e.export([0])  # the root is resident after a run
        assert len(desc) == _ffi.EXPORT_DESC_BYTES
        with pytest.raises(_ffi.DcsmcError):
            e.export([1])  # consumed by its parent: not resident
        with pytest.raises(_ffi.DcsmcError):
            e.import_peer([1], [3], desc)  # peer 3 was never mapped with dcsmc_peer_open
        handle, gen, nbytes = e.peer_pool_handle()
        assert len(handle) == _ffi.IPC_HANDLE_BYTES and gen >= 1 and nbytes == e.run_info().poolBytes
