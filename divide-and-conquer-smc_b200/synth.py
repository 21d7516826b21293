"""Synthetic inputs shaped like the reference's data (no network: the NYS CSV of
scripts/prepare-data.sh cannot be downloaded). Row format = PreprocessNYSchoolData.java:81:
    NY,<county>,<district>,<school>,<year>,<nTested>,<level3+4>
i.e. 5 path components (tree levels 0-4) then nTrials, nSuccesses."""
from __future__ import annotations

from typing import List

import numpy as np

from .tree import DirectedTree, MultiLevelDataset, Node


def nys_shaped_rows(seed: int = 20260101, nCounties: int = 5, nDistricts: int = 32, nSchools: int = 700,
                    years=(2006, 2007, 2008, 2009)) -> List[List[str]]:
    """~nSchools*len(years) leaves under NY -> county -> district -> school -> year.
    theta_child ~ N(theta_parent, sigma2), sigma2 ~ Exp(1); nTrials ~ 20 + Poisson(80);
    nSuccesses ~ Binomial(nTrials, logistic(theta))."""
    rng = np.random.default_rng(seed)
    rows: List[List[str]] = []
    # uneven split of districts over counties and schools over districts
    dist_county = np.sort(rng.integers(0, nCounties, nDistricts))
    dist_county[:nCounties] = np.arange(nCounties)
    dist_county = np.sort(dist_county)
    w = rng.gamma(2.0, 1.0, nDistricts)
    school_dist = np.sort(rng.choice(nDistricts, size=nSchools, p=w / w.sum()))
    school_dist[:nDistricts] = np.arange(nDistricts)
    school_dist = np.sort(school_dist)
    theta0 = 0.3
    th_county = theta0 + rng.normal(0, np.sqrt(rng.exponential(1.0)), nCounties)
    th_dist = th_county[dist_county] + rng.normal(0, np.sqrt(rng.exponential(1.0) * 0.5), nDistricts)
    th_school = th_dist[school_dist] + rng.normal(0, np.sqrt(rng.exponential(1.0) * 0.3), nSchools)
    for s in range(nSchools):
        d = int(school_dist[s])
        c = int(dist_county[d])
        for y in years:
            th = th_school[s] + rng.normal(0, 0.3)
            n = int(20 + rng.poisson(80))
            k = int(rng.binomial(n, 1.0 / (1.0 + np.exp(-th))))
            rows.append(["NY", f"county{c:02d}", f"district{d:03d}", f"school{s:04d}", str(y), str(n), str(k)])
    return rows


def write_csv(rows: List[List[str]], path: str):
    with open(path, "w") as fh:
        for r in rows:
            fh.write(",".join(r) + "\n")


def small_multilevel_rows(seed: int = 7, nGroups: int = 3, maxLeaves: int = 4, edge_cases: bool = True) -> List[List[str]]:
    """Tiny 3-level hierarchy for parity tests; includes s=0 and s=n leaves (Cheng BC branch)
    and a single-child internal node (BrownianModelCalculator.combine C==1 branch)."""
    rng = np.random.default_rng(seed)
    rows: List[List[str]] = []
    for g in range(nGroups):
        nl = int(rng.integers(2, maxLeaves + 1))
        for l in range(nl):
            n = int(rng.integers(5, 60))
            k = int(rng.integers(1, n))
            rows.append(["R", f"g{g}", f"l{l}", str(n), str(k)])
    if edge_cases:
        rows.append(["R", "gE", "zero", "12", "0"])
        rows.append(["R", "gE", "full", "9", "9"])
        rows.append(["R", "gE", "tiny", "1", "1"])
        rows.append(["R", "gS", "only", "30", "11"])  # single child
    return rows


def wide_rows(seed: int, fanouts, nTrialsMean: int = 80) -> List[List[str]]:
    """Balanced hierarchy with the given fan-outs per level, binomial leaves."""
    rng = np.random.default_rng(seed)
    rows: List[List[str]] = []

    def rec(path, level):
        if level == len(fanouts):
            n = int(20 + rng.poisson(nTrialsMean))
            k = int(rng.binomial(n, rng.beta(2, 2)))
            rows.append(path + [str(n), str(k)])
            return
        for i in range(fanouts[level]):
            rec(path + [f"n{level}_{i}"], level + 1)

    rec(["R"], 0)
    return rows


def binary_gauss_tree(depth: int, seed: int = 4):
    """BASELINE config 4: perfect binary tree (labels as TestUtilities.perfectBinaryTree),
    observed real leaves y ~ N(0,1)."""
    from .tree import perfectBinaryTree

    tree = perfectBinaryTree(depth)
    csr = tree.to_csr()
    rng = np.random.default_rng(seed)
    leafObs = np.zeros(csr.nNodes)
    leafKind = np.zeros(csr.nNodes, dtype=np.int32)
    for i in range(csr.nNodes):
        if csr.nChildren(i) == 0:
            leafObs[i] = rng.normal()
            leafKind[i] = 1
    return csr, leafKind, leafObs
