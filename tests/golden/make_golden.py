#!/usr/bin/env python
"""Generates the committed golden vectors of tests/golden/ (run from the repo root:
`python tests/golden/make_golden.py`).

* doc_recorded_run.json — NOT generated: the values the reference itself records for its only test
  (src/test/java/dc/Doc.java:282 = README.md:257, and the exact logZ of Doc.java:287). The oracle must
  reproduce them to every printed digit (tests/test_oracle_golden.py).
* multilevel_small.json, summaries_small.json, wide_underflow.json — produced by the CPU oracle
  (oracle/dcsmc_oracle.c, exact-sum contract, fdlibm, Philox seed 31) on small seeded inputs. The
  reference cannot run in this environment (no JVM, private Maven dependencies), so these pin the
  ENGINE and future versions of the oracle to the oracle as committed with them; they are not
  reference outputs, and the parity of the hierarchical-binomial model stays "unpinned" (DESIGN.md §7).
Floats are stored as float.hex() strings (exact); ancestor / log-weight arrays as SHA-256 digests."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import dcsmc_b200 as d  # noqa: E402
import oracle_lib as o  # noqa: E402
from dcsmc_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def digest(a):
    a = np.asarray(a)
    if a.dtype.kind == "f":
        a = np.where(np.isnan(a), np.nan, a)  # one NaN bit pattern (leaves carry variance = NaN)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_dataset():
    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows())
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    return ds, csr, nT, nS


def multilevel_small():
    ds, csr, nT, nS = small_dataset()
    N = 4096
    cases = []
    for scheme in (o.MULTINOMIAL, o.STRATIFIED):
        for thr in (1.0 + 1e-10, 0.5):
            r = o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, relativeEssThreshold=thr,
                      rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
            cases.append(dict(resamplingScheme=int(scheme), relativeEssThreshold=float(thr).hex(),
                              logNormEstimate=hx(r.logNormEstimate), ess=hx(r.ess), logScaling=hx(r.logScaling),
                              resampled=[int(x) for x in r.resampled],
                              ancestors_sha256=[digest(r.anc[n].astype(np.int32)) for n in range(csr.nNodes)],
                              logw_sha256=[digest(r.logw[n]) for n in range(csr.nNodes)],
                              state_sha256=[digest(r.state[n]) for n in range(csr.nNodes)],
                              rootState_sha256=digest(r.rootState), rootWeights_sha256=digest(r.rootWeights)))
    return dict(input="synth.small_multilevel_rows()", nNodes=csr.nNodes, nParticles=N, masterRandomSeed=31,
                rng="philox4x32-10", sums="exact", elementary="fdlibm", nTrials=[int(x) for x in nT],
                nSuccesses=[int(x) for x in nS], childOffsets=[int(x) for x in csr.childOffsets],
                children=[int(x) for x in csr.children], cases=cases)


def summaries_small():
    ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=11))
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    N, cut = 6000, 2
    out = []
    depth = csr.depth()
    for scheme in (o.MULTINOMIAL, o.STRATIFIED):
        r = o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT,
                  elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
        nodes = {}
        for i in range(csr.nNodes):
            if depth[i] >= cut:
                continue
            s = o.node_summaries(r, csr, i, levelCutOffForOutput=cut)
            nodes[str(i)] = dict(meanMean=float(s["meanMean"]).hex(), meanSD=float(s["meanSD"]).hex(), hasVar=bool(s["hasVar"]),
                                 varMean=float(s["varMean"]).hex(), varSD=float(s["varSD"]).hex(),
                                 delta={str(c): [float(a).hex(), float(b).hex()] for c, (a, b) in s["delta"].items()})
        out.append(dict(resamplingScheme=int(scheme), nodes=nodes))
    return dict(input="synth.small_multilevel_rows(seed=11)", nParticles=N, levelCutOffForOutput=cut, masterRandomSeed=31,
                tolerance="statistics are fp64 sums: compare at 1e-12 relative", cases=out)


def wide_underflow():
    N, C = 2000, 110
    rows = [["NY", "c0", "d0", f"s{k}", "2006", str(40 + (7 * k) % 50), str(10 + (3 * k) % 25)] for k in range(C)]
    ds = d.MultiLevelDataset(rows=rows)
    csr = ds.getTree().to_csr()
    nT, nS = ds.leaf_arrays(csr)
    cases = []
    for thr in (1.0 + 1e-10, 0.5):
        r = o.run(csr, N, masterRandomSeed=31, resamplingScheme=o.STRATIFIED, relativeEssThreshold=thr, rngMode=o.RNG_PHILOX,
                  sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS)
        cases.append(dict(relativeEssThreshold=float(thr).hex(), logNormEstimate=hx(r.logNormEstimate), ess=hx(r.ess)))
    return dict(input="one district with 110 schools (see tests/test_gpu_parity.py::test_wide_node_weight_product_underflow)",
                nParticles=N, nChildren=C, resamplingScheme=int(o.STRATIFIED), masterRandomSeed=31, cases=cases)


def main():
    for name, fn in (("multilevel_small", multilevel_small), ("summaries_small", summaries_small), ("wide_underflow", wide_underflow)):
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(fn(), fh, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
