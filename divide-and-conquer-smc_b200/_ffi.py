"""ctypes binding of include/dcsmc.h (libdcsmc.so, built in-tree from csrc/).

There is no CPU fallback: a missing library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DCSMC_LIB selects another build of the same sources (profiles/: timing-only experiment variants)
LIB_PATH = os.environ.get("DCSMC_LIB") or os.path.join(HERE, "csrc", "libdcsmc.so")

OK, EINVAL, ECUDA, ENOMEM, ESTATE = 0, 1, 2, 3, 4
MULTINOMIAL, STRATIFIED = 0, 1
RNG_PHILOX, RNG_INJECTED, RNG_JAVA = 0, 1, 2
MODEL_MULTILEVEL, MODEL_MARKOV = 0, 1
LEAF_BINOMIAL, LEAF_GAUSS_OBS = 0, 1
SUM_EXACT, SUM_JAVA = 0, 1
SAMPLE_NATURAL, SAMPLE_MEAN, SAMPLE_VAR, SAMPLE_DELTA0 = 0, 1, 2, 3
ABI_VERSION = 3
K_NAMES = ["leaf", "internal", "scan", "finalize", "darts", "resample", "gather", "markov"]


class Options(C.Structure):
    _fields_ = [
        ("masterRandomSeed", C.c_int64),
        ("nParticles", C.c_int32),
        ("resamplingScheme", C.c_int32),
        ("relativeEssThreshold", C.c_double),
        ("clusterSubGroup", C.c_int32),
        ("nThreadsPerNode", C.c_int32),
        ("minimumNumberOfClusterMembersToStart", C.c_int32),
        ("maximumTimeToWaitInMinutes", C.c_int32),
        ("maximumDistributionDepth", C.c_int32),
        ("indexInCluster", C.c_int32),
        ("device", C.c_int32),
        ("rngMode", C.c_int32),
        ("sumMode", C.c_int32),
        ("maxBetaAttempts", C.c_int32),
        ("retainPopulations", C.c_int32),
        ("levelCutOffForOutput", C.c_int32),
        ("memoryBudgetBytes", C.c_int64),
        ("useTransform", C.c_int32),
        ("retainSamples", C.c_int32),
    ]


class NodeStats(C.Structure):
    _fields_ = [
        ("ess", C.c_double),
        ("relEss", C.c_double),
        ("logNormEstimate", C.c_double),
        ("logScaling", C.c_double),
        ("resampled", C.c_int32),
        ("done", C.c_int32),
    ]


class NodeSummary(C.Structure):
    _fields_ = [
        ("meanMean", C.c_double), ("meanSD", C.c_double),
        ("varMean", C.c_double), ("varSD", C.c_double),
        ("deltaMean", C.c_double), ("deltaSD", C.c_double),
        ("hasMean", C.c_int32), ("hasVar", C.c_int32), ("hasDelta", C.c_int32), ("reserved", C.c_int32),
    ]


class PopulationView(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("istate", C.c_void_p), ("logWeights", C.c_void_p), ("ancestors", C.c_void_p),
        ("stride", C.c_int64), ("nParticles", C.c_int32), ("nColumns", C.c_int32), ("resampled", C.c_int32),
        ("onPeer", C.c_int32), ("constantObservation", C.c_int32), ("reserved", C.c_int32),
        ("observation", C.c_double), ("logScaling", C.c_double), ("maxLogWeight", C.c_double),
        ("sumExp", C.c_double), ("ess", C.c_double),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("poolBytesNeeded", C.c_int64), ("poolBytes", C.c_int64), ("peakBytes", C.c_int64),
        ("scratchBytes", C.c_int64), ("tableBytes", C.c_int64), ("nodesPlanned", C.c_int64),
        ("levelBatches", C.c_int32), ("maxBatchNodes", C.c_int32), ("feasible", C.c_int32), ("reserved", C.c_int32),
    ]


class RunInfo(C.Structure):
    _fields_ = [
        ("deviceMs", C.c_double),
        ("kernelLaunches", C.c_int64),
        ("particleUpdates", C.c_int64),
        ("nodesProcessed", C.c_int64),
        ("algorithmicBytes", C.c_int64),
        ("poolBytes", C.c_int64),
        ("h2dBytes", C.c_int64),
        ("d2hBytes", C.c_int64),
        ("levelBatches", C.c_int32),
        ("reserved", C.c_int32),
    ]


# every symbol include/dcsmc.h declares; tests check that the library exports all of them
SYMBOLS = [
    "dcsmc_abi_version", "dcsmc_options_default", "dcsmc_create", "dcsmc_destroy", "dcsmc_last_error",
    "dcsmc_set_tree", "dcsmc_set_node_seeds", "dcsmc_set_multilevel_model", "dcsmc_set_markov_model",
    "dcsmc_slots_per_particle", "dcsmc_inject_uniforms", "dcsmc_run", "dcsmc_run_subtrees",
    "dcsmc_run_nodes", "dcsmc_node_stats", "dcsmc_all_node_stats", "dcsmc_all_node_stats_columns", "dcsmc_population_copy_to_host",
    "dcsmc_debug_copy_node", "dcsmc_population_packed_bytes", "dcsmc_population_pack",
    "dcsmc_population_unpack", "dcsmc_population_release", "dcsmc_last_run_info",
    "dcsmc_profile_enable", "dcsmc_profile_get", "dcsmc_debug_eval", "dcsmc_reset_populations",
    "dcsmc_timer_begin", "dcsmc_timer_end",
    "dcsmc_node_samples", "dcsmc_node_times", "dcsmc_population_view",
    "dcsmc_stream", "dcsmc_begin_deferred", "dcsmc_sync", "dcsmc_stats_export_device", "dcsmc_population_export_layout",
    "dcsmc_population_import_peer_device", "dcsmc_stats_matrix_device",
    "dcsmc_node_summary", "dcsmc_plan_query", "dcsmc_host_eval", "dcsmc_peer_pool_handle", "dcsmc_peer_open", "dcsmc_population_export", "dcsmc_population_import_peer",
]
IPC_HANDLE_BYTES, EXPORT_DESC_BYTES, STATS_RECORD_BYTES = 64, 128, 64

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    i32, i64, dbl, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    P = C.POINTER
    L.dcsmc_abi_version.restype = i32
    L.dcsmc_options_default.argtypes = [P(Options)]
    L.dcsmc_create.argtypes = [P(Options), P(vp)]
    L.dcsmc_destroy.argtypes = [vp]
    L.dcsmc_last_error.argtypes = [vp]
    L.dcsmc_last_error.restype = C.c_char_p
    L.dcsmc_set_tree.argtypes = [vp, i32, i32, P(i32), P(i32)]
    L.dcsmc_set_node_seeds.argtypes = [vp, P(i32)]
    L.dcsmc_set_multilevel_model.argtypes = [vp, P(i32), P(i32), P(i32), P(dbl), dbl]
    L.dcsmc_set_markov_model.argtypes = [vp, i32, P(dbl), P(dbl)]
    L.dcsmc_slots_per_particle.argtypes = [vp, i32, P(i32)]
    L.dcsmc_inject_uniforms.argtypes = [vp, i32, P(dbl), i64]
    L.dcsmc_run.argtypes = [vp]
    L.dcsmc_run_subtrees.argtypes = [vp, P(i32), i32]
    L.dcsmc_run_nodes.argtypes = [vp, P(i32), i32]
    L.dcsmc_node_stats.argtypes = [vp, i32, P(NodeStats)]
    L.dcsmc_all_node_stats.argtypes = [vp, P(NodeStats)]
    L.dcsmc_all_node_stats_columns.argtypes = [vp] + [P(dbl)] * 4 + [P(i32)] * 2
    L.dcsmc_population_copy_to_host.argtypes = [vp, i32, P(dbl), P(i32), P(dbl)]
    L.dcsmc_debug_copy_node.argtypes = [vp, i32, P(dbl), P(i32), P(dbl), P(i32)]
    L.dcsmc_population_packed_bytes.argtypes = [vp]
    L.dcsmc_population_packed_bytes.restype = i64
    L.dcsmc_population_pack.argtypes = [vp, i32, vp]
    L.dcsmc_population_unpack.argtypes = [vp, i32, vp]
    L.dcsmc_population_release.argtypes = [vp, i32]
    L.dcsmc_reset_populations.argtypes = [vp]
    L.dcsmc_last_run_info.argtypes = [vp, P(RunInfo)]
    L.dcsmc_node_summary.argtypes = [vp, i32, P(NodeSummary)]
    L.dcsmc_node_samples.argtypes = [vp, i32, i32, P(dbl)]
    L.dcsmc_node_times.argtypes = [vp, P(dbl)]
    L.dcsmc_population_view.argtypes = [vp, i32, P(PopulationView)]
    L.dcsmc_plan_query.argtypes = [P(Options), i32, P(i32), P(i32), P(i32), i32, P(i32), i32, i64, P(PlanInfo), C.c_char_p, i32]
    L.dcsmc_peer_pool_handle.argtypes = [vp, vp, P(i64), P(i64)]
    L.dcsmc_peer_open.argtypes = [vp, i32, vp, i64, i64]
    L.dcsmc_population_export.argtypes = [vp, P(i32), i32, vp]
    L.dcsmc_population_import_peer.argtypes = [vp, P(i32), P(i32), i32, vp]
    L.dcsmc_stream.argtypes = [vp, P(vp)]
    L.dcsmc_begin_deferred.argtypes = [vp]
    L.dcsmc_sync.argtypes = [vp]
    L.dcsmc_stats_matrix_device.argtypes = [vp, vp]
    L.dcsmc_stats_export_device.argtypes = [vp, P(i32), i32, vp]
    L.dcsmc_population_export_layout.argtypes = [vp, P(i32), i32, vp]
    L.dcsmc_population_import_peer_device.argtypes = [vp, P(i32), P(i32), i32, vp, vp, P(i32), i32]
    L.dcsmc_timer_begin.argtypes = [vp]
    L.dcsmc_timer_end.argtypes = [vp, P(dbl)]
    L.dcsmc_profile_enable.argtypes = [vp, i32]
    L.dcsmc_profile_get.argtypes = [vp, i32, P(dbl), P(i64)]
    L.dcsmc_debug_eval.argtypes = [vp, i32, P(dbl), P(dbl), i64]
    L.dcsmc_host_eval.argtypes = [i32, P(dbl), P(dbl), i64]
    for s in SYMBOLS:
        f = getattr(L, s)
        if s not in ("dcsmc_last_error", "dcsmc_population_packed_bytes"):
            f.restype = i32
    _lib = L
    return L


class DcsmcError(RuntimeError):
    """The reference throws bare RuntimeExceptions (DCRecursion.java:43,50); the C ABI returns
    a status code + message, surfaced here."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"dcsmc error {code}: {msg}")
        self.code = code
