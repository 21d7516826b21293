"""Thin object wrapper over the C ABI (include/dcsmc.h). One Engine == one dcsmc_engine handle."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi as F


def _ptr(a: Optional[np.ndarray], ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def plan_query(csr, *, model: int = F.MODEL_MULTILEVEL, leafKind=None, roots=None, budgetBytes: int = 0, **opts) -> F.PlanInfo:
    """dcsmc_plan_query: what the static memory planner would do for this tree (or for the subtrees `roots`,
    one rank's share of a sharded run) — host only, no CUDA device needed. `opts` are dcsmc_options fields
    (nParticles, resamplingScheme, retainPopulations, ...)."""
    L = F.lib()
    o = F.Options()
    L.dcsmc_options_default(C.byref(o))
    for k, v in opts.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    off = np.ascontiguousarray(csr.childOffsets, dtype=np.int32)
    ch = np.ascontiguousarray(csr.children if len(csr.children) else np.zeros(1, np.int32), dtype=np.int32)
    lk = None if leafKind is None else np.ascontiguousarray(leafKind, dtype=np.int32)
    info = F.PlanInfo()
    err = C.create_string_buffer(512)
    rt = None if roots is None else np.ascontiguousarray(roots, dtype=np.int32)
    rc = L.dcsmc_plan_query(C.byref(o), csr.nNodes, _ptr(off, C.c_int32), _ptr(ch, C.c_int32), _ptr(lk, C.c_int32),
                            model, _ptr(rt, C.c_int32), 0 if rt is None else len(rt), int(budgetBytes), C.byref(info), err, len(err))
    if rc != 0:
        raise F.DcsmcError(rc, err.value.decode())
    return info


def host_eval(op: int, x: np.ndarray) -> np.ndarray:
    """dcsmc_host_eval: the engine's own log / exp / Philox (csrc/dcsmc_math.cuh) compiled for the host."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    rc = F.lib().dcsmc_host_eval(op, _ptr(x, C.c_double), _ptr(y, C.c_double), len(x))
    if rc != 0:
        raise F.DcsmcError(rc, "dcsmc_host_eval: bad arguments")
    return y


class Engine:
    def __init__(self, **opts):
        L = F.lib()
        o = F.Options()
        L.dcsmc_options_default(C.byref(o))
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k}")
            setattr(o, k, v)
        self.options = o
        h = C.c_void_p()
        rc = L.dcsmc_create(C.byref(o), C.byref(h))
        if rc != 0:
            raise F.DcsmcError(rc, L.dcsmc_last_error(None).decode())
        self._h = h
        self._L = L
        self.nNodes = 0
        self.N = o.nParticles

    # -- helpers ---------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise F.DcsmcError(rc, self._L.dcsmc_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.dcsmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- topology / model --------------------------------------------------------------------------
    def set_tree(self, csr):
        off = np.ascontiguousarray(csr.childOffsets, dtype=np.int32)
        ch = np.ascontiguousarray(csr.children if len(csr.children) else np.zeros(1, np.int32), dtype=np.int32)
        self._ck(self._L.dcsmc_set_tree(self._h, csr.nNodes, 0, _ptr(off, C.c_int32), _ptr(ch, C.c_int32)))
        self.nNodes = csr.nNodes
        h = np.ascontiguousarray(csr.javaHash, dtype=np.int32)
        self._ck(self._L.dcsmc_set_node_seeds(self._h, _ptr(h, C.c_int32)))

    def set_multilevel_model(self, nTrials=None, nSuccesses=None, leafKind=None, leafObs=None, variancePrior=1.0):
        a = lambda x, dt: None if x is None else np.ascontiguousarray(x, dtype=dt)
        lk, nt, ns, lo = a(leafKind, np.int32), a(nTrials, np.int32), a(nSuccesses, np.int32), a(leafObs, np.float64)
        self._ck(self._L.dcsmc_set_multilevel_model(
            self._h, _ptr(lk, C.c_int32), _ptr(nt, C.c_int32), _ptr(ns, C.c_int32), _ptr(lo, C.c_double),
            float(variancePrior)))

    def set_markov_model(self, transition, prior):
        t = np.ascontiguousarray(transition, dtype=np.float64)
        p = np.ascontiguousarray(prior, dtype=np.float64)
        self._ck(self._L.dcsmc_set_markov_model(self._h, len(p), _ptr(t, C.c_double), _ptr(p, C.c_double)))

    def slots_per_particle(self, node: int) -> int:
        s = C.c_int32()
        self._ck(self._L.dcsmc_slots_per_particle(self._h, node, C.byref(s)))
        return s.value

    def inject_uniforms(self, node: int, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._ck(self._L.dcsmc_inject_uniforms(self._h, node, _ptr(u, C.c_double), len(u)))

    # -- execution -----------------------------------------------------------------------------------
    def run(self):
        self._ck(self._L.dcsmc_run(self._h))

    def run_subtrees(self, roots: Sequence[int]):
        r = np.ascontiguousarray(roots, dtype=np.int32)
        self._ck(self._L.dcsmc_run_subtrees(self._h, _ptr(r, C.c_int32), len(r)))

    def run_nodes(self, nodes: Sequence[int]):
        r = np.ascontiguousarray(nodes, dtype=np.int32)
        self._ck(self._L.dcsmc_run_nodes(self._h, _ptr(r, C.c_int32), len(r)))

    # -- results ---------------------------------------------------------------------------------------
    def node_stats(self, node: int) -> F.NodeStats:
        s = F.NodeStats()
        self._ck(self._L.dcsmc_node_stats(self._h, node, C.byref(s)))
        return s

    def node_summary(self, node: int) -> F.NodeSummary:
        s = F.NodeSummary()
        self._ck(self._L.dcsmc_node_summary(self._h, node, C.byref(s)))
        return s

    def node_samples(self, node: int, which: int) -> np.ndarray:
        """dcsmc_node_samples: one retained sample series of a summary node (F.SAMPLE_*)."""
        out = np.empty(self.N)
        self._ck(self._L.dcsmc_node_samples(self._h, node, which, _ptr(out, C.c_double)))
        return out

    def node_times(self) -> np.ndarray:
        """dcsmc_node_times: per-node share of its level batch's device time (ms); profiled runs only."""
        out = np.zeros(max(1, self.nNodes))
        self._ck(self._L.dcsmc_node_times(self._h, _ptr(out, C.c_double)))
        return out[:self.nNodes]

    def population_view(self, node: int) -> F.PopulationView:
        """dcsmc_population_view: borrowed device pointers (valid until the next run / reset / destroy)."""
        v = F.PopulationView()
        self._ck(self._L.dcsmc_population_view(self._h, node, C.byref(v)))
        return v

    def all_node_stats(self) -> Dict[str, np.ndarray]:
        # one array per field, filled by the library in one pass (no per-node struct walk or strided copy in Python)
        n = self.nNodes
        out = {k: np.empty(n) for k in ("ess", "relEss", "logNormEstimate", "logScaling")}
        out["resampled"] = np.empty(n, np.int32)
        out["done"] = np.empty(n, np.int32)
        self._ck(self._L.dcsmc_all_node_stats_columns(
            self._h, *[_ptr(out[k], C.c_double) for k in ("ess", "relEss", "logNormEstimate", "logScaling")],
            _ptr(out["resampled"], C.c_int32), _ptr(out["done"], C.c_int32)))
        return out

    def population(self, node: int, state=True, istate=False, weights=True, out=None):
        """Copy of the node's population after its node step. `out` = (state[6][N] f64, istate[N] i32,
        weights[N] f64) lets a caller reuse its host buffers (any entry may be None)."""
        N = self.N
        if out is not None:
            st, ist, w = out
            for a, shape, dt in ((st, (6, N), np.float64), (ist, (N,), np.int32), (w, (N,), np.float64)):
                if a is not None and (a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]):
                    raise ValueError(f"population(out=...): expected C-contiguous {shape} {dt}")
        else:
            st = np.empty((6, N)) if state else None
            ist = np.empty(N, np.int32) if istate else None
            w = np.empty(N) if weights else None
        self._ck(self._L.dcsmc_population_copy_to_host(
            self._h, node, _ptr(st, C.c_double), _ptr(ist, C.c_int32), _ptr(w, C.c_double)))
        return st, ist, w

    def debug_node(self, node: int):
        N = self.N
        st = np.empty((6, N))
        ist = np.empty(N, np.int32)
        lw = np.empty(N)
        anc = np.empty(N, np.int32)
        self._ck(self._L.dcsmc_debug_copy_node(
            self._h, node, _ptr(st, C.c_double), _ptr(ist, C.c_int32), _ptr(lw, C.c_double), _ptr(anc, C.c_int32)))
        return st, ist, lw, anc

    def debug_eval(self, op: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._ck(self._L.dcsmc_debug_eval(self._h, op, _ptr(x, C.c_double), _ptr(y, C.c_double), len(x)))
        return y

    def packed_bytes(self) -> int:
        return int(self._L.dcsmc_population_packed_bytes(self._h))

    def pack(self, node: int, device_ptr: int):
        self._ck(self._L.dcsmc_population_pack(self._h, node, C.c_void_p(device_ptr)))

    def unpack(self, node: int, device_ptr: int):
        self._ck(self._L.dcsmc_population_unpack(self._h, node, C.c_void_p(device_ptr)))

    # -- zero-copy sharing over NVLink peer memory (CUDA IPC) ---------------------------------------
    def peer_pool_handle(self):
        """(64-byte IPC handle of this engine's pool, pool generation, pool bytes)"""
        h = (C.c_char * F.IPC_HANDLE_BYTES)()
        gen, nbytes = C.c_int64(), C.c_int64()
        self._ck(self._L.dcsmc_peer_pool_handle(self._h, h, C.byref(gen), C.byref(nbytes)))
        return bytes(h), gen.value, nbytes.value

    def peer_open(self, peer: int, handle: bytes, generation: int, poolBytes: int):
        buf = (C.c_char * F.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        self._ck(self._L.dcsmc_peer_open(self._h, peer, buf, generation, poolBytes))

    def export(self, nodes: Sequence[int]) -> bytes:
        r = np.ascontiguousarray(nodes, dtype=np.int32)
        out = (C.c_char * (F.EXPORT_DESC_BYTES * max(1, len(r))))()
        self._ck(self._L.dcsmc_population_export(self._h, _ptr(r, C.c_int32), len(r), out))
        return bytes(out)[:F.EXPORT_DESC_BYTES * len(r)]

    def import_peer(self, nodes: Sequence[int], peers: Sequence[int], descs: bytes):
        r = np.ascontiguousarray(nodes, dtype=np.int32)
        q = np.ascontiguousarray(peers, dtype=np.int32)
        buf = (C.c_char * max(1, len(descs))).from_buffer_copy(descs if descs else b"\0")
        self._ck(self._L.dcsmc_population_import_peer(self._h, _ptr(r, C.c_int32), _ptr(q, C.c_int32), len(r), buf))

    # -- sharded runs without host round trips (include/dcsmc.h) ----------------------------------------
    def stream(self) -> int:
        """the engine's launching stream (cudaStream_t as an integer)"""
        p = C.c_void_p()
        self._ck(self._L.dcsmc_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def begin_deferred(self):
        self._ck(self._L.dcsmc_begin_deferred(self._h))

    def sync(self):
        self._ck(self._L.dcsmc_sync(self._h))

    def stats_export_device(self, nodes: Sequence[int], device_ptr: int):
        r = np.ascontiguousarray(nodes, dtype=np.int32)
        self._ck(self._L.dcsmc_stats_export_device(self._h, _ptr(r, C.c_int32), len(r), C.c_void_p(device_ptr)))

    def export_layout(self, nodes: Sequence[int]) -> bytes:
        r = np.ascontiguousarray(nodes, dtype=np.int32)
        out = (C.c_char * (F.EXPORT_DESC_BYTES * max(1, len(r))))()
        self._ck(self._L.dcsmc_population_export_layout(self._h, _ptr(r, C.c_int32), len(r), out))
        return bytes(out)[:F.EXPORT_DESC_BYTES * len(r)]

    def import_peer_device(self, nodes: Sequence[int], peers: Sequence[int], descs: bytes, device_stats_ptr: int,
                           stat_index: Sequence[int], n_records: int):
        r = np.ascontiguousarray(nodes, dtype=np.int32)
        q = np.ascontiguousarray(peers, dtype=np.int32)
        ix = np.ascontiguousarray(stat_index, dtype=np.int32)
        buf = (C.c_char * max(1, len(descs))).from_buffer_copy(descs if descs else b"\0")
        self._ck(self._L.dcsmc_population_import_peer_device(
            self._h, _ptr(r, C.c_int32), _ptr(q, C.c_int32), len(r), buf, C.c_void_p(device_stats_ptr), _ptr(ix, C.c_int32),
            int(n_records)))

    def stats_matrix_device(self, device_ptr: int):
        """dcsmc_stats_matrix_device: [6][nNodes] doubles (ess, relEss, logNormEstimate, logScaling, resampled,
        computedHere) of the nodes this engine computed, on the device."""
        self._ck(self._L.dcsmc_stats_matrix_device(self._h, C.c_void_p(device_ptr)))

    def reset_populations(self):
        self._ck(self._L.dcsmc_reset_populations(self._h))

    def release(self, node: int):
        self._ck(self._L.dcsmc_population_release(self._h, node))

    def run_info(self) -> F.RunInfo:
        r = F.RunInfo()
        self._ck(self._L.dcsmc_last_run_info(self._h, C.byref(r)))
        return r

    def timer_begin(self):
        self._ck(self._L.dcsmc_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_double()
        self._ck(self._L.dcsmc_timer_end(self._h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool = True):
        self._ck(self._L.dcsmc_profile_enable(self._h, 1 if on else 0))

    def profile(self) -> Dict[str, Dict[str, float]]:
        out = {}
        for k, name in enumerate(F.K_NAMES):
            ms, n = C.c_double(), C.c_int64()
            self._ck(self._L.dcsmc_profile_get(self._h, k, C.byref(ms), C.byref(n)))
            out[name] = {"ms": ms.value, "launches": n.value}
        return out
