"""Static subtree -> GPU sharding: the B200 replacement of the reference's task graph
(dc/DistributedDC.java:165-175 populateInitialTasks; dc/DCRecursionTask.java:77-93
prepareNextTask). Pure host logic, no torch: unit-tested on CPU.

The reference cuts the tree at `maximumDistributionDepth`: every node at that depth (or a leaf
above it) is an initial task that recurses serially through its whole subtree; nodes above the
cut become tasks when their last child finishes and fetch the children populations from the
cluster map. Here the same cut yields (i) subtrees, assigned to ranks by longest-processing-time (or by a
contiguous partition when that is as well balanced: siblings then stay together) and run level-batched on one GPU each, and (ii) a short, fully static schedule for the nodes above
the cut: for each level, which populations move between which ranks and which rank runs which
node. Per-node random streams are keyed by node id, so the result does not depend on the
assignment (Doc.java:268).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


def initial_tasks(csr, maximumDistributionDepth: int) -> List[int]:
    """DistributedDC.populateInitialTasks (DistributedDC.java:165-175)."""
    out: List[int] = []
    stack = [(0, maximumDistributionDepth)]
    while stack:
        node, d = stack.pop()
        if csr.nChildren(node) == 0 or d == 0:
            out.append(node)
        else:
            ch = csr.children[csr.childOffsets[node]:csr.childOffsets[node + 1]]
            for c in reversed(ch):
                stack.append((int(c), d - 1))
    return out


def effective_cut_depth(csr, maximumDistributionDepth: int, world: int, oversubscribe: int = 4) -> int:
    """The reference's default (Integer.MAX_VALUE) makes every leaf a task; on GPUs coarse tasks are
    better, so an unbounded depth is replaced by the shallowest depth that yields at least
    oversubscribe*world subtrees (README.md:116-121 describes the same coarsening knob)."""
    depth = csr.depth()
    height = int(depth.max())
    if maximumDistributionDepth <= height:
        return maximumDistributionDepth
    for d in range(height + 1):
        if len(initial_tasks(csr, d)) >= oversubscribe * world:
            return d
    return height


def assign_lpt(tasks: List[int], weights: List[int], world: int) -> Dict[int, int]:
    """Longest-processing-time assignment of subtree roots to ranks (ties -> lowest rank)."""
    load = [0] * world
    owner: Dict[int, int] = {}
    for t, w in sorted(zip(tasks, weights), key=lambda x: (-x[1], x[0])):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[t] = r
        load[r] += w
    return owner


def assign_contiguous(tasks: List[int], weights: List[int], world: int) -> Dict[int, int]:
    """Contiguous partition of the subtree roots (they come in depth-first order, so siblings are adjacent)
    into at most `world` ranges with the smallest possible maximum load. Siblings then share a rank, and the
    nodes above the cut find most of their children locally instead of reading them over NVLink."""
    if not tasks:
        return {}
    def ranges(cap: int) -> List[List[int]]:
        out: List[List[int]] = []
        cur: List[int] = []
        load = 0
        for t, w in zip(tasks, weights):
            if cur and load + w > cap:
                out.append(cur)
                cur, load = [], 0
            cur.append(t)
            load += w
        out.append(cur)
        return out

    lo, hi = max(weights), sum(weights)
    while lo < hi:
        mid = (lo + hi) // 2
        if len(ranges(mid)) <= world:
            hi = mid
        else:
            lo = mid + 1
    owner: Dict[int, int] = {}
    for r, group in enumerate(ranges(lo)):
        for t in group:
            owner[t] = r
    return owner


def _makespan(owner: Dict[int, int], tasks: List[int], weights: List[int], world: int) -> int:
    load = [0] * world
    for t, w in zip(tasks, weights):
        load[owner[t]] += w
    return max(load)


def assign(tasks: List[int], weights: List[int], world: int, slack: float = 0.02) -> Dict[int, int]:
    """Longest-processing-time gives the best balance, a contiguous partition the best locality (fewest remote
    children above the cut). Balance decides the run time of the subtrees, locality only the short levels
    above the cut, so the contiguous partition is taken only when its slowest rank is within `slack` of LPT's:
    always on regular trees (BASELINE configs 4 and 5), not on the uneven NYS-shaped tree (4-16 % worse)."""
    lpt = assign_lpt(tasks, weights, world)
    if not tasks or world <= 1:
        return lpt
    contiguous = assign_contiguous(tasks, weights, world)
    if _makespan(contiguous, tasks, weights, world) <= (1.0 + slack) * _makespan(lpt, tasks, weights, world):
        return contiguous
    return lpt


@dataclass
class TopLevel:
    nodes: List[Tuple[int, int]] = field(default_factory=list)            # (node, owner rank)
    transfers: List[Tuple[int, int, int]] = field(default_factory=list)   # (child node, src, dst)


@dataclass
class ShardPlan:
    cut_depth: int
    tasks: List[int]
    task_owner: Dict[int, int]
    node_owner: Dict[int, int]       # owner rank of every task root and every node above the cut
    top_levels: List[TopLevel]

    def my_tasks(self, rank: int) -> List[int]:
        return [t for t in self.tasks if self.task_owner[t] == rank]


def make_plan(csr, maximumDistributionDepth: int, world: int) -> ShardPlan:
    cut = effective_cut_depth(csr, maximumDistributionDepth, world)
    tasks = initial_tasks(csr, cut)
    sizes = csr.subtree_sizes()
    owner = assign(tasks, [int(sizes[t]) for t in tasks], world)
    node_owner: Dict[int, int] = dict(owner)
    task_set = set(tasks)
    # nodes above the cut = ancestors of task roots
    top: List[int] = []
    seen = set()
    for t in tasks:
        p = int(csr.parent[t])
        while p >= 0 and p not in seen:
            seen.add(p)
            top.append(p)
            p = int(csr.parent[p])
    height = csr.height()
    top.sort(key=lambda n: (int(height[n]), n))
    levels: Dict[int, TopLevel] = {}
    for n in top:
        ch = [int(c) for c in csr.children[csr.childOffsets[n]:csr.childOffsets[n + 1]]]
        votes = [0] * world
        for c in ch:
            votes[node_owner[c]] += 1
        own = max(range(world), key=lambda k: (votes[k], -k))
        node_owner[n] = own
        lvl = levels.setdefault(int(height[n]), TopLevel())
        lvl.nodes.append((n, own))
        for c in ch:
            if node_owner[c] != own:
                lvl.transfers.append((c, node_owner[c], own))
    assert all(t in task_set or t in seen for t in node_owner)
    return ShardPlan(cut, tasks, owner, node_owner, [levels[h] for h in sorted(levels)])


def plan_memory(csr, plan: ShardPlan, budgetBytes: int = 0, **engine_options):
    """Per-rank schedule and memory plan of a sharded run (dcsmc_plan_query on each rank's subtrees): host only,
    no device needed. Returns one PlanInfo per rank (None for a rank without subtrees)."""
    from .engine import plan_query

    world = max(plan.task_owner.values(), default=-1) + 1
    out = []
    for r in range(world):
        mine = plan.my_tasks(r)
        out.append(plan_query(csr, roots=mine, budgetBytes=budgetBytes, **engine_options) if mine else None)
    return out
