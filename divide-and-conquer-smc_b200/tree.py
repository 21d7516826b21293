"""Tree topology and dataset loading: host-side mirror of the reference's L0 layer.

Mirrors (reference paths relative to /root/reference/src/main/java):
  * prototype/Node.java:14-76            -> Node (value hashCode, level(), toString)
  * bayonet DirectedTree [un-vendored]   -> DirectedTree (root, ordered children, parent)
  * src/test/java/dc/TestUtilities.java:20-38 -> perfectBinaryTree
  * prototype/io/MultiLevelDataset.java:51-66,102-117 and Datum.java -> MultiLevelDataset

Everything here is integer / string work on the host; the GPU engine consumes the CSR
arrays produced by DirectedTree.to_csr().
"""
from __future__ import annotations

import csv
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def _i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x >= (1 << 31) else x


def java_string_hash(s: str) -> int:
    """java.lang.String.hashCode (UTF-16 code units, int32 wrap)."""
    h = 0
    data = s.encode("utf-16-be")
    for k in range(0, len(data), 2):
        h = (31 * h + ((data[k] << 8) | data[k + 1])) & 0xFFFFFFFF
    return _i32(h)


def java_list_hash(items: Iterable[str]) -> int:
    """java.util.List.hashCode over Strings."""
    h = 1
    for e in items:
        h = (31 * h + java_string_hash(e)) & 0xFFFFFFFF
    return _i32(h)


class Node:
    """prototype/Node.java: a node is its path from the root; equality and hashCode are
    value-based so the per-node seed is reproducible across machines (Node.java:34-40)."""

    __slots__ = ("path",)

    def __init__(self, path: Sequence[str]):
        self.path = tuple(path)

    def child(self, childName: str) -> "Node":
        return Node(self.path + (childName,))

    @staticmethod
    def root(name: str) -> "Node":
        return Node((name,))

    def hashCode(self) -> int:
        return _i32(31 * 1 + java_list_hash(self.path))

    def level(self) -> int:
        return len(self.path) - 1

    def toString(self, separator: str = "-") -> str:
        return separator.join(self.path)

    def __hash__(self):
        return hash(self.path)

    def __eq__(self, other):
        return isinstance(other, Node) and self.path == other.path

    def __repr__(self):
        return self.toString()


class DirectedTree:
    """Minimal bayonet.graphs.DirectedTree: children keep insertion order
    (the fold order of BrownianModelCalculator.combine is fp-visible)."""

    def __init__(self, root):
        self._root = root
        self._children: Dict[object, List[object]] = {root: []}
        self._parent: Dict[object, Optional[object]] = {root: None}

    def getRoot(self):
        return self._root

    def addChild(self, parent, child):
        if parent not in self._children:
            raise RuntimeError("parent not in tree")
        if child in self._children:
            raise RuntimeError("child already in tree")
        self._children[parent].append(child)
        self._children[child] = []
        self._parent[child] = parent

    def getChildren(self, node) -> List[object]:
        return self._children[node]

    def getParent(self, node):
        return self._parent[node]

    def isLeaf(self, node) -> bool:
        return len(self._children[node]) == 0

    def getNodes(self) -> List[object]:
        return list(self._children.keys())

    def to_csr(self) -> "CsrTree":
        """Number nodes in pre-order DFS (root = 0) and emit CSR arrays."""
        order: List[object] = []
        stack = [self._root]
        while stack:
            n = stack.pop()
            order.append(n)
            stack.extend(reversed(self._children[n]))
        index = {n: i for i, n in enumerate(order)}
        nN = len(order)
        offsets = np.zeros(nN + 1, dtype=np.int32)
        children: List[int] = []
        parent = np.full(nN, -1, dtype=np.int32)
        for i, n in enumerate(order):
            ch = self._children[n]
            offsets[i + 1] = offsets[i] + len(ch)
            for c in ch:
                children.append(index[c])
                parent[index[c]] = i
        hashes = np.array(
            [n.hashCode() if hasattr(n, "hashCode") else _i32(hash(n)) for n in order], dtype=np.int32
        )
        return CsrTree(
            nodes=order,
            index=index,
            parent=parent,
            childOffsets=offsets,
            children=np.array(children, dtype=np.int32),
            javaHash=hashes,
        )


@dataclass
class CsrTree:
    nodes: List[object]
    index: Dict[object, int]
    parent: np.ndarray
    childOffsets: np.ndarray
    children: np.ndarray
    javaHash: np.ndarray

    @property
    def nNodes(self) -> int:
        return len(self.nodes)

    def nChildren(self, i: int) -> int:
        return int(self.childOffsets[i + 1] - self.childOffsets[i])

    def depth(self) -> np.ndarray:
        d = np.zeros(self.nNodes, dtype=np.int32)
        for i in range(1, self.nNodes):  # pre-order: parents precede children
            d[i] = d[self.parent[i]] + 1
        return d

    def height(self) -> np.ndarray:
        h = np.zeros(self.nNodes, dtype=np.int32)
        for i in range(self.nNodes - 1, 0, -1):
            p = self.parent[i]
            h[p] = max(h[p], h[i] + 1)
        return h

    def subtree_sizes(self) -> np.ndarray:
        s = np.ones(self.nNodes, dtype=np.int64)
        for i in range(self.nNodes - 1, 0, -1):
            s[self.parent[i]] += s[i]
        return s


def perfectBinaryTree(depth: int) -> DirectedTree:
    """TestUtilities.perfectBinaryTree (TestUtilities.java:20-38): root path ["0"], child k
    appends str(k)."""
    root = Node(("0",))
    tree = DirectedTree(root)

    def build(cur: Node, remaining: int):
        if remaining <= 0:
            return
        for i in range(2):
            ch = cur.child(str(i))
            tree.addChild(cur, ch)
            build(ch, remaining - 1)

    build(root, depth)
    return tree


@dataclass
class Datum:
    """prototype/io/Datum.java"""

    numberOfTrials: int
    numberOfSuccesses: int


class MultiLevelDataset:
    """prototype/io/MultiLevelDataset.java:51-66: CSV rows `path..., nTrials, nSuccesses`;
    the child order of a node is first-appearance order in the file."""

    def __init__(self, file=None, rows: Optional[Iterable[Sequence[str]]] = None):
        self.children: Dict[Node, List[Node]] = {}
        self._childset: Dict[Node, set] = {}
        self.data: Dict[Node, Datum] = {}
        self.root: Optional[Node] = None
        if file is not None:
            with open(file, newline="") as fh:
                rows = [r for r in csv.reader(fh) if r]
        for line in rows or []:
            line = [str(x) for x in line]
            path, datum = line[:-2], line[-2:]
            if self.root is None:
                self.root = Node.root(path[0])
            for i in range(len(path) - 1):
                par, ch = Node(path[: i + 1]), Node(path[: i + 2])
                if par not in self.children:
                    self.children[par] = []
                    self._childset[par] = set()
                if ch not in self._childset[par]:
                    self._childset[par].add(ch)
                    self.children[par].append(ch)
            self.data[Node(path)] = Datum(int(datum[0]), int(datum[1]))

    def getChildren(self, node: Node) -> List[Node]:
        return self.children.get(node, [])

    def getRoot(self) -> Node:
        return self.root

    def getDatum(self, node: Node) -> Datum:
        return self.data[node]

    def getTree(self) -> DirectedTree:  # MultiLevelDataset.java:102-117
        tree = DirectedTree(self.root)
        stack = [self.root]
        while stack:
            n = stack.pop()
            for c in self.getChildren(n):
                tree.addChild(n, c)
            stack.extend(reversed(self.getChildren(n)))
        return tree

    def postOrder(self) -> List[Node]:
        out: List[Node] = []

        def rec(n):
            for c in self.getChildren(n):
                rec(c)
            out.append(n)

        rec(self.root)
        return out

    def leaf_arrays(self, csr: CsrTree) -> Tuple[np.ndarray, np.ndarray]:
        nT = np.zeros(csr.nNodes, dtype=np.int32)
        nS = np.zeros(csr.nNodes, dtype=np.int32)
        for n, d in self.data.items():
            i = csr.index[n]
            nT[i] = d.numberOfTrials
            nS[i] = d.numberOfSuccesses
        return nT, nS
