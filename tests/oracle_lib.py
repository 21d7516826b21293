"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE: imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

RNG_JAVA, RNG_PHILOX, RNG_INJECTED = 0, 1, 2
SUM_JAVA, SUM_EXACT = 0, 1
ELEM_LIBM, ELEM_FDLIBM = 0, 1
MULTINOMIAL, STRATIFIED = 0, 1
MODEL_MULTILEVEL, MODEL_MARKOV = 0, 1
LEAF_BINOMIAL, LEAF_GAUSS_OBS = 0, 1


class OrcOptions(C.Structure):
    _fields_ = [
        ("masterRandomSeed", C.c_int64),
        ("nParticles", C.c_int32),
        ("resamplingScheme", C.c_int32),
        ("relativeEssThreshold", C.c_double),
        ("nThreads", C.c_int32),
        ("maximumDistributionDepth", C.c_int32),
        ("rngMode", C.c_int32),
        ("sumMode", C.c_int32),
        ("elemMode", C.c_int32),
        ("model", C.c_int32),
        ("impl", C.c_int32),
        ("maxBetaAttempts", C.c_int32),
        ("variancePriorRate", C.c_double),
        ("nStates", C.c_int32),
        ("transition", C.POINTER(C.c_double)),
        ("prior", C.POINTER(C.c_double)),
    ]


class OrcTree(C.Structure):
    _fields_ = [
        ("nNodes", C.c_int32),
        ("root", C.c_int32),
        ("childOffsets", C.POINTER(C.c_int32)),
        ("children", C.POINTER(C.c_int32)),
        ("javaHash", C.POINTER(C.c_int32)),
        ("leafKind", C.POINTER(C.c_int32)),
        ("nTrials", C.POINTER(C.c_int32)),
        ("nSuccesses", C.POINTER(C.c_int32)),
        ("leafObs", C.POINTER(C.c_double)),
    ]


class OrcOutputs(C.Structure):
    _fields_ = [
        ("ess", C.POINTER(C.c_double)),
        ("relEss", C.POINTER(C.c_double)),
        ("logNormEstimate", C.POINTER(C.c_double)),
        ("logScaling", C.POINTER(C.c_double)),
        ("resampled", C.POINTER(C.c_int32)),
        ("state", C.POINTER(C.c_double)),
        ("istate", C.POINTER(C.c_int32)),
        ("logw", C.POINTER(C.c_double)),
        ("anc", C.POINTER(C.c_int32)),
        ("rootState", C.POINTER(C.c_double)),
        ("rootIState", C.POINTER(C.c_int32)),
        ("rootWeights", C.POINTER(C.c_double)),
    ]


_lib = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return os.path.join(ORACLE_DIR, "liboracle.so")


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src_mtime = max(
            os.path.getmtime(os.path.join(ORACLE_DIR, f))
            for f in ("dcsmc_oracle.c", "dcsmc_oracle.h", "orc_elem.h")
        )
        if not os.path.exists(path) or os.path.getmtime(path) < src_mtime:
            build()
        L = C.CDLL(path)
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.POINTER(OrcOptions), C.POINTER(OrcTree), C.POINTER(C.POINTER(C.c_double)), C.POINTER(OrcOutputs)]
        L.orc_java_seed_scramble.restype = C.c_uint64
        L.orc_java_seed_scramble.argtypes = [C.c_int64]
        L.orc_java_next.restype = C.c_int32
        L.orc_java_next.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        L.orc_java_next_double.restype = C.c_double
        L.orc_java_next_double.argtypes = [C.POINTER(C.c_uint64)]
        L.orc_java_next_int.restype = C.c_int32
        L.orc_java_next_int.argtypes = [C.POINTER(C.c_uint64), C.c_int32]
        L.orc_java_string_hash.restype = C.c_int32
        L.orc_java_string_hash.argtypes = [C.c_char_p]
        L.orc_node_seed.restype = C.c_int64
        L.orc_node_seed.argtypes = [C.c_int64, C.c_int32]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniform.argtypes = [C.c_int64, C.c_int32, C.c_int64]
        L.orc_elem_log.restype = C.c_double
        L.orc_elem_log.argtypes = [C.c_int, C.c_double]
        L.orc_elem_exp.restype = C.c_double
        L.orc_elem_exp.argtypes = [C.c_int, C.c_double]
        L.orc_brownian_combine.restype = None
        L.orc_brownian_combine.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
        dp = C.POINTER(C.c_double)
        L.orc_node_summaries.restype = C.c_int
        L.orc_node_summaries.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, dp, dp, dp,
                                         C.c_int32, C.c_int32, dp, dp, dp]
        L.orc_node_sample_series.restype = C.c_int
        L.orc_node_sample_series.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, dp, dp, dp,
                                             C.c_int32, C.c_int32, dp, dp, dp, dp]
        L.orc_slots_per_particle.restype = C.c_int32
        L.orc_slots_per_particle.argtypes = [C.POINTER(OrcOptions), C.POINTER(OrcTree), C.c_int32]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


@dataclass
class OracleResult:
    ess: np.ndarray
    relEss: np.ndarray
    logNormEstimate: np.ndarray
    logScaling: np.ndarray
    resampled: np.ndarray
    state: Optional[np.ndarray]
    istate: Optional[np.ndarray]
    logw: Optional[np.ndarray]
    anc: Optional[np.ndarray]
    rootState: np.ndarray
    rootIState: np.ndarray
    rootWeights: np.ndarray


def run(
    csr,
    nParticles: int,
    *,
    model: int = MODEL_MULTILEVEL,
    masterRandomSeed: int = 31,
    resamplingScheme: int = MULTINOMIAL,
    relativeEssThreshold: float = 1.0 + 1e-10,
    nThreads: int = 1,
    maximumDistributionDepth: int = 2**31 - 1,
    rngMode: int = RNG_JAVA,
    sumMode: int = SUM_JAVA,
    elemMode: int = ELEM_LIBM,
    impl: int = 2,
    maxBetaAttempts: int = 32,
    variancePriorRate: float = 1.0,
    nTrials: Optional[np.ndarray] = None,
    nSuccesses: Optional[np.ndarray] = None,
    leafKind: Optional[np.ndarray] = None,
    leafObs: Optional[np.ndarray] = None,
    transition: Optional[np.ndarray] = None,
    prior: Optional[np.ndarray] = None,
    injected: Optional[List[Optional[np.ndarray]]] = None,
    dump: bool = False,
    root: int = 0,
) -> OracleResult:
    """`root` != 0 runs only the subtree under that node (DCRecursionTask.run(node, false)); the per-node outputs
    of the other nodes stay zero and rootState/rootWeights describe that node's population."""
    L = lib()
    nN = csr.nNodes
    N = nParticles
    keep = []

    def arr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a

    tr = arr(transition, np.float64)
    pr = arr(prior, np.float64)
    opt = OrcOptions(
        masterRandomSeed, N, resamplingScheme, relativeEssThreshold, nThreads, maximumDistributionDepth,
        rngMode, sumMode, elemMode, model, impl, maxBetaAttempts, variancePriorRate,
        0 if pr is None else len(pr), _p(tr, C.c_double), _p(pr, C.c_double),
    )
    off = arr(csr.childOffsets, np.int32)
    ch = arr(csr.children if len(csr.children) else np.zeros(1, np.int32), np.int32)
    jh = arr(csr.javaHash, np.int32)
    lk = arr(leafKind if leafKind is not None else np.zeros(nN, np.int32), np.int32)
    nt = arr(nTrials if nTrials is not None else np.zeros(nN, np.int32), np.int32)
    ns = arr(nSuccesses if nSuccesses is not None else np.zeros(nN, np.int32), np.int32)
    lo = arr(leafObs if leafObs is not None else np.zeros(nN, np.float64), np.float64)
    tree = OrcTree(nN, int(root), _p(off, C.c_int32), _p(ch, C.c_int32), _p(jh, C.c_int32), _p(lk, C.c_int32),
                   _p(nt, C.c_int32), _p(ns, C.c_int32), _p(lo, C.c_double))
    res = OracleResult(
        ess=np.zeros(nN), relEss=np.zeros(nN), logNormEstimate=np.zeros(nN), logScaling=np.zeros(nN),
        resampled=np.zeros(nN, np.int32),
        state=np.zeros((nN, 6, N)) if dump else None,
        istate=np.zeros((nN, N), np.int32) if dump else None,
        logw=np.zeros((nN, N)) if dump else None,
        anc=np.zeros((nN, N), np.int32) if dump else None,
        rootState=np.zeros((6, N)), rootIState=np.zeros(N, np.int32), rootWeights=np.zeros(N),
    )
    out = OrcOutputs(
        _p(res.ess, C.c_double), _p(res.relEss, C.c_double), _p(res.logNormEstimate, C.c_double),
        _p(res.logScaling, C.c_double), _p(res.resampled, C.c_int32), _p(res.state, C.c_double),
        _p(res.istate, C.c_int32), _p(res.logw, C.c_double), _p(res.anc, C.c_int32),
        _p(res.rootState, C.c_double), _p(res.rootIState, C.c_int32), _p(res.rootWeights, C.c_double),
    )
    inj_ptr = None
    if injected is not None:
        PD = C.POINTER(C.c_double)
        inj_arr = (PD * nN)()
        for i, a in enumerate(injected):
            if a is not None:
                a = arr(a, np.float64)
                inj_arr[i] = _p(a, C.c_double)
        keep.append(inj_arr)
        inj_ptr = C.cast(inj_arr, C.POINTER(PD))
    rc = L.orc_run(C.byref(opt), C.byref(tree), inj_ptr, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_run failed with code {rc}")
    return res


def slots_per_particle(csr, node: int, *, model=MODEL_MULTILEVEL, maxBetaAttempts=32, leafKind=None) -> int:
    nC = csr.nChildren(node)
    if model == MODEL_MARKOV:
        return 0 if nC == 0 else 1
    if nC > 0:
        return 1
    if leafKind is not None and leafKind[node] == LEAF_GAUSS_OBS:
        return 0
    return 2 * maxBetaAttempts


def node_summaries(ref: OracleResult, csr, node: int, *, levelCutOffForOutput: int, masterRandomSeed: int = 31,
                   maxBetaAttempts: int = 32, useTransform: bool = True, elemMode: int = ELEM_FDLIBM,
                   injected: Optional[np.ndarray] = None) -> Dict[str, object]:
    """impl-1 posterior summaries of `node` from a dumped oracle run (ref = run(..., dump=True)): the
    resampled population is rebuilt from the pre-resampling particles and the ancestor indices, the
    children's messages are aligned with it, and orc_node_summaries restates the reference's loops."""
    L = lib()
    N = ref.state.shape[2]
    depth = csr.depth()
    C_ = csr.nChildren(node)
    kids = [int(c) for c in csr.children[csr.childOffsets[node]:csr.childOffsets[node + 1]]]
    a = ref.anc[node]  # identity when the node was not resampled
    mean = np.ascontiguousarray(ref.state[node][0][a])
    msgVar = np.ascontiguousarray(ref.state[node][1][a])
    variance = np.ascontiguousarray(ref.state[node][5][a]) if C_ > 0 else None
    doDelta = 1 if (depth[node] + 1 < levelCutOffForOutput and C_ > 0) else 0
    cm = np.zeros((max(C_, 1), N))
    cv = np.zeros((max(C_, 1), N))
    for k, ch in enumerate(kids):
        ci = ref.anc[ch][a]
        cm[k] = ref.state[ch][0][ci]
        cv[k] = ref.state[ch][1][ci]
    out = np.zeros(4 + 2 * max(C_, 1))
    inj = None if injected is None else np.ascontiguousarray(injected, dtype=np.float64)
    # the per-particle series behind the statistics = the reference's `samples` table (logSamples :508-519)
    nSeries = 2 + (1 if C_ > 0 else 0) + (C_ if doDelta else 0)
    series = np.zeros((nSeries, N))
    rc = L.orc_node_sample_series(elemMode, masterRandomSeed, node, N, maxBetaAttempts, 1 if useTransform else 0,
                                  _p(inj, C.c_double), _p(mean, C.c_double), _p(msgVar, C.c_double),
                                  _p(variance, C.c_double), C_, doDelta, _p(cm, C.c_double), _p(cv, C.c_double),
                                  _p(out, C.c_double), _p(series, C.c_double))
    assert rc == 0
    res = dict(meanMean=out[0], meanSD=out[1], hasVar=C_ > 0, varMean=out[2], varSD=out[3], delta={},
               samples=dict(naturalParam=series[0], meanSample=series[1]), deltaSamples={})
    if C_ > 0:
        res["samples"]["varSample"] = series[2]
    if doDelta:
        for k, ch in enumerate(kids):
            res["delta"][ch] = (out[4 + 2 * k], out[5 + 2 * k])
            res["deltaSamples"][ch] = series[(3 if C_ > 0 else 2) + k]
    return res
