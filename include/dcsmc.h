/*
 * dcsmc.h — C ABI of the B200-native DC-SMC node engine (libdcsmc.so).
 *
 * This is the drop-in boundary for the per-tree-node DC-SMC step of
 * alexandrebouchard/divide-and-conquer-smc. The reference's FFI-free Java path is
 *     DistributedDC.start() -> DCRecursionTask.run() -> DCRecursion.dcRecurse()
 *         -> DCProposal.propose() per particle -> ParticlePopulation.resample()
 * A per-particle JVM callback cannot be a GPU boundary, so the boundary is one level up:
 * "run the node step of a known proposal family for a whole (sub)tree, level-batched".
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/src/main/java). INTEGRATION.md shows the JNI / Panama binding.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a
 * DCSMC_E* code, and dcsmc_last_error() returns the message (no exception crosses the
 * ABI; the reference throws bare RuntimeExceptions, e.g. DCRecursion.java:43,50).
 * One engine handle is driven by one host thread at a time; distinct handles are
 * independent (this lifts the reference's one-instance-per-JVM limit,
 * DistributedDC.java:63-67). The engine owns all device memory; particles never cross
 * the ABI unless the caller asks for a copy. There is NO CPU fallback: dcsmc_create
 * fails when no CUDA device is present.
 */
#ifndef DCSMC_H
#define DCSMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCSMC_ABI_VERSION 3

enum {
  DCSMC_OK = 0,
  DCSMC_EINVAL = 1,   /* bad argument / bad state (reference: RuntimeException) */
  DCSMC_ECUDA = 2,    /* CUDA runtime error */
  DCSMC_ENOMEM = 3,   /* population pool does not fit the memory budget */
  DCSMC_ESTATE = 4    /* call sequence violated (e.g. run before set_tree) */
};

/* bayonet.smc.ResamplingScheme as used by DCOptions.resamplingScheme (DCOptions.java:22) */
enum { DCSMC_MULTINOMIAL = 0, DCSMC_STRATIFIED = 1 };

/* source of the uniforms (engine extension; the reference has java.util.Random only) */
enum {
  DCSMC_RNG_PHILOX = 0,   /* counter-based Philox4x32-10 keyed by (masterRandomSeed, node) */
  DCSMC_RNG_INJECTED = 1, /* caller-supplied uniform stream per node (parity runs) */
  DCSMC_RNG_JAVA = 2      /* java.util.Random(seed(node)) reproduced with LCG jump-ahead;
                             only for proposals with fixed draw counts (internal nodes,
                             Markov model); binomial leaves are rejected */
};

/* proposal families = implementations of dc.DCProposalFactory shipped with the reference */
enum {
  DCSMC_MODEL_MULTILEVEL = 0, /* prototype.adaptor.MultiLevelProposalFactory */
  DCSMC_MODEL_MARKOV = 1      /* Doc.markovChainNaiveProposalFactory (src/test/java/dc/Doc.java:206) */
};

enum { DCSMC_LEAF_BINOMIAL = 0, DCSMC_LEAF_GAUSS_OBS = 1 };

/* summation contract of the normaliser / CDF (DESIGN.md "Numerics contract") */
enum {
  DCSMC_SUM_EXACT = 0, /* order-independent fixed-point sums (default, fast) */
  DCSMC_SUM_JAVA = 1   /* strictly sequential fp64 sums, the reference's own order */
};

/* dc.DCOptions (dc/DCOptions.java:15-45), field for field, then engine extensions. */
typedef struct dcsmc_options {
  int64_t masterRandomSeed;                     /* :16  default 31 */
  int32_t nParticles;                           /* :19  default 1000 */
  int32_t resamplingScheme;                     /* :22  default MULTINOMIAL */
  double relativeEssThreshold;                  /* :25  default 1 + THRESHOLD (always resample) */
  int32_t clusterSubGroup;                      /* :29  accepted, unused */
  int32_t nThreadsPerNode;                      /* :32  accepted, unused (the GPU is the pool) */
  int32_t minimumNumberOfClusterMembersToStart; /* :35  accepted, unused */
  int32_t maximumTimeToWaitInMinutes;           /* :38  accepted, unused */
  int32_t maximumDistributionDepth;             /* :41  cut depth for subtree sharding */
  int32_t indexInCluster;                       /* :45  accepted, unused */
  /* ---- engine extensions ---- */
  int32_t device;            /* CUDA device ordinal */
  int32_t rngMode;           /* DCSMC_RNG_* */
  int32_t sumMode;           /* DCSMC_SUM_* */
  int32_t maxBetaAttempts;   /* Beta rejection-sampler attempt cap; slot budget per leaf
                                particle = 2*maxBetaAttempts uniforms (default 32) */
  int32_t retainPopulations; /* 1: keep every node's population resident (debug/parity) */
  int32_t levelCutOffForOutput; /* DcSmcOptions.levelCutOffForOutput (prototype/smc/
                                   DivideConquerMCAlgorithm.java:56): posterior summaries for nodes
                                   with level() < this; 0 (default) = none, reference default 2 */
  int64_t memoryBudgetBytes; /* 0: use 80% of free device memory */
  int32_t useTransform;      /* MultiLevelModelOptions.useTransform (:47): summaries on the
                                logistic scale (default 1) */
  int32_t retainSamples;     /* 1: keep the per-particle sample series behind the summaries (the reference's
                                `samples` table, DivideConquerMCAlgorithm.java:508-519) for dcsmc_node_samples */
} dcsmc_options;

typedef struct dcsmc_engine dcsmc_engine;

/* ---- life cycle --------------------------------------------------------------------- */

/* new DCOptions() defaults (DCOptions.java:15-45) + engine defaults. */
int32_t dcsmc_options_default(dcsmc_options* opt);

/* DistributedDC.createInstance(options, proposalFactory, tree) is split into
 * dcsmc_create + dcsmc_set_tree + dcsmc_set_*_model (DistributedDC.java:58-70). */
int32_t dcsmc_create(const dcsmc_options* opt, dcsmc_engine** out);
int32_t dcsmc_destroy(dcsmc_engine* e);
/* message of the last failing call on this handle (or of dcsmc_create when e == NULL) */
const char* dcsmc_last_error(const dcsmc_engine* e);
int32_t dcsmc_abi_version(void);

/* ---- topology and model --------------------------------------------------------------- */

/* bayonet DirectedTree<N> passed to createInstance: CSR with children in the reference's
 * order (DCRecursionTask.java:51; the order is the fold order of
 * BrownianModelCalculator.combine and therefore visible in floating point). */
int32_t dcsmc_set_tree(dcsmc_engine* e, int32_t nNodes, int32_t root,
                       const int32_t* childOffsets, const int32_t* children);

/* N.hashCode() per node, feeding DCRecursionTask.getRandom (DCRecursionTask.java:98-106):
 * seed = 31*(31*1 + masterRandomSeed) + hashCode. Used by DCSMC_RNG_JAVA. */
int32_t dcsmc_set_node_seeds(dcsmc_engine* e, const int32_t* javaHashCodes);

/* prototype.adaptor.MultiLevelProposalFactory (MultiLevelProposalFactory.java:22-52):
 * per-node arrays (entries of internal nodes ignored); leafKind/leafObs may be NULL
 * (all leaves binomial). variancePrior is the Exponential rate (default 1.0). */
int32_t dcsmc_set_multilevel_model(dcsmc_engine* e, const int32_t* leafKind,
                                   const int32_t* nTrials, const int32_t* nSuccesses,
                                   const double* leafObs, double variancePrior);

/* Doc.markovChainNaiveProposalFactory (Doc.java:162-225): row-major transition matrix. */
int32_t dcsmc_set_markov_model(dcsmc_engine* e, int32_t nStates, const double* transition,
                               const double* prior);

/* Number of uniforms a particle of this node may draw in propose() (slot stride S); the
 * node's stream is [N*S proposal slots | N resampling darts]. */
int32_t dcsmc_slots_per_particle(const dcsmc_engine* e, int32_t node, int32_t* slots);

/* Replaces the `Random random` argument of DCProposal.propose / ParticlePopulation.resample
 * (DCProposal.java:28, DCRecursion.java:31) for DCSMC_RNG_INJECTED: host buffer of
 * N*S + N uniforms in [0,1) for this node; copied to the device. */
int32_t dcsmc_inject_uniforms(dcsmc_engine* e, int32_t node, const double* u, int64_t n);

/* ---- execution ------------------------------------------------------------------------ */

/* DistributedDC.start() (DistributedDC.java:87-92): run DCRecursion.dcRecurse for every
 * node, bottom-up, all sibling nodes of a level in one batch of launches. Blocking. */
int32_t dcsmc_run(dcsmc_engine* e);

/* DCRecursionTask.run(node, false) for several subtree roots at once
 * (DCRecursionTask.java:48-75; DistributedDC.populateInitialTasks :165-175): runs every
 * node under each root; the roots' populations stay resident. */
int32_t dcsmc_run_subtrees(dcsmc_engine* e, const int32_t* roots, int32_t nRoots);

/* DCRecursionTask.run(node, true) (:48-66) for nodes whose children populations are
 * already resident (computed here or installed with dcsmc_population_unpack). */
int32_t dcsmc_run_nodes(dcsmc_engine* e, const int32_t* nodes, int32_t nNodes);

/* ---- results -------------------------------------------------------------------------- */

/* What DCProcessor.process(populationBeforeResampling, ..) can read
 * (DCProcessor.java:18, DefaultProcessorFactory.java:41-54): getESS(), getRelativeESS(),
 * logNormEstimate(), plus logScaling and whether resample() ran (DCRecursion.java:29-31). */
typedef struct dcsmc_node_stats_t {
  double ess;
  double relEss;
  double logNormEstimate;
  double logScaling;
  int32_t resampled;
  int32_t done;
} dcsmc_node_stats_t;

int32_t dcsmc_node_stats(const dcsmc_engine* e, int32_t node, dcsmc_node_stats_t* out);
/* all nodes at once: out[nNodes] */
int32_t dcsmc_all_node_stats(const dcsmc_engine* e, dcsmc_node_stats_t* out);
/* the same table as one array per field (each [nNodes]; any pointer may be NULL): what a columnar consumer — the
 * `timing` / `logZ` rows of DefaultProcessorFactory over a whole tree (DefaultProcessorFactory.java:41-54) — reads
 * without a per-node struct walk */
int32_t dcsmc_all_node_stats_columns(const dcsmc_engine* e, double* ess, double* relEss, double* logNormEstimate,
                                     double* logScaling, int32_t* resampled, int32_t* done);

/* Posterior summaries of impl-1 (printNodeStatistics / printDeltaNodeStatistics,
 * DivideConquerMCAlgorithm.java:433-507), computed on the device right after the node step of every
 * node with level() < levelCutOffForOutput, from the resampled population:
 *   meanStats : mean / SD of inverseTransform(sampleValue) over the particles (:476-498)
 *   varStats  : mean / SD of the particles' variance (internal nodes, :488-505)
 *   deltaStats: for a node whose parent has level()+1 < levelCutOffForOutput, mean / SD of
 *               inverseTransform(child sample) - inverseTransform(parent sample) with the child drawn
 *               jointly (sampleChildrenJointly, :203-217) — reported under the CHILD's path (:460-463).
 * Mean and SD follow commons-math3 DescriptiveStatistics (two-pass mean, bias-corrected variance).
 * Gaussian draws: java.util.Random.nextGaussian's polar method on a separate slot-addressed uniform
 * stream of the node. A node that was not resampled reports nothing (the reference's
 * statistics assume equally weighted particles, :292). */
typedef struct dcsmc_node_summary_t {
  double meanMean, meanSD;
  double varMean, varSD;
  double deltaMean, deltaSD;
  int32_t hasMean, hasVar, hasDelta, reserved;
} dcsmc_node_summary_t;
int32_t dcsmc_node_summary(dcsmc_engine* e, int32_t node, dcsmc_node_summary_t* out);

/* The reference's `samples` table (logSamples, DivideConquerMCAlgorithm.java:508-519; rows node, variable,
 * iteration, value): the per-particle series behind the summaries of a node with level() < levelCutOffForOutput,
 * kept on the device when dcsmc_options.retainSamples = 1. `which`: DCSMC_SAMPLE_NATURAL ("naturalParam", :497),
 * _MEAN ("meanSample", :499), _VAR ("varSample", internal nodes, :504), _DELTA0 + c ("delta" of child c, logged
 * under the CHILD's path, :461). out[nParticles]. */
enum { DCSMC_SAMPLE_NATURAL = 0, DCSMC_SAMPLE_MEAN = 1, DCSMC_SAMPLE_VAR = 2, DCSMC_SAMPLE_DELTA0 = 3 };
int32_t dcsmc_node_samples(dcsmc_engine* e, int32_t node, int32_t which, double* out);

/* DefaultProcessorFactory's iterationProposalTime (DefaultProcessorFactory.java:41-50): per node, the device
 * time (ms) of the level batch it ran in divided by the sibling nodes of that batch. Filled by runs made with
 * dcsmc_profile_enable(e, 1); zeros otherwise. out[nNodes]. */
int32_t dcsmc_node_times(const dcsmc_engine* e, double* msPerNode);

/* ParticlePopulation without a copy: BORROWED device pointers to the node's population, valid until the next
 * run / reset / destroy on this engine (co-resident consumers: another CUDA library, a JNI direct buffer over
 * device memory, a plotting kernel). Particle j of the population is row ancestors[j] of the columns when the
 * node was resampled (ancestors != NULL), else row j with weight exp(logWeights[j] - maxLogWeight) / sumExp. */
typedef struct dcsmc_population_view_t {
  const double* state;        /* column c at state + c * stride; NULL for Markov / constant leaves */
  const int32_t* istate;      /* Markov particle */
  const double* logWeights;   /* raw log weights; NULL for leaves (all equal) */
  const int32_t* ancestors;   /* NULL unless resampled on this engine */
  int64_t stride;             /* elements between columns */
  int32_t nParticles;
  int32_t nColumns;           /* 6: mean, msgVar, logLik, descObs, descVar, variance; 2 (leaves): mean, descObs; 0 */
  int32_t resampled;
  int32_t onPeer;             /* the memory belongs to a peer engine's pool (mapped over NVLink) */
  int32_t constantObservation; /* observed leaf that is not materialised: every particle is (observation, 0, 0, 0, 0, NaN) */
  int32_t reserved;
  double observation;
  double logScaling, maxLogWeight, sumExp, ess;
} dcsmc_population_view_t;
int32_t dcsmc_population_view(dcsmc_engine* e, int32_t node, dcsmc_population_view_t* out);

/* DistributedDC.getRootPopulation() / ParticlePopulation accessors (DistributedDC.java:107;
 * SURVEY §8a row P): copies the node's population AFTER the node step (resampled if
 * resampling ran) to host buffers; any pointer may be NULL.
 *   state   [6][N]: mean, msgVar, logLik, descObs, descVar, variance (multilevel model)
 *   istate  [N]   : Markov particle
 *   weights [N]   : getNormalizedWeight(i)
 * Blocking. Populations above 4 MiB are copied device -> pinned staging in 4 MiB pieces and moved into the caller's
 * (pageable) arrays by up to four short-lived host threads while the next pieces are still on the bus. */
int32_t dcsmc_population_copy_to_host(dcsmc_engine* e, int32_t node, double* state,
                                      int32_t* istate, double* weights);

/* Parity/debug view of a node (needs retainPopulations=1): the population BEFORE
 * resampling, the raw log weights and the ancestor indices (identity if not resampled). */
int32_t dcsmc_debug_copy_node(dcsmc_engine* e, int32_t node, double* statePre,
                              int32_t* istatePre, double* logWeights, int32_t* ancestors);

/* Parity/debug: evaluate an elementary function of the numerics contract on the DEVICE for n host
 * inputs (op 0: log, 1: exp — the fdlibm algorithms StrictMath specifies; 2 | node << 8: Philox uniforms;
 * 3: the kernels' two straight-line division forms on the pairs (x[2k], x[2k+1]) -> (y[2k], y[2k+1]), which must equal
 * IEEE x[2k] / x[2k+1]; 4, 5: the straight-line log / exp of the leaf kernels, valid for non-rare arguments). */
int32_t dcsmc_debug_eval(dcsmc_engine* e, int32_t op, const double* x, double* y, int64_t n);
/* The same functions from the same header compiled for the host (no device needed): op 0 log, 1 exp,
 * 2 | node << 8: Philox uniform number (uint64)x[i] of `node`, seed 31. */
int32_t dcsmc_host_eval(int32_t op, const double* x, double* y, int64_t n);

/* ---- moving populations between engines (replaces Hazelcast populations.put/remove,
 * DCRecursionTask.java:37,113) ---------------------------------------------------------- */

/* bytes of one packed population (post-node-step, SoA) */
int64_t dcsmc_population_packed_bytes(const dcsmc_engine* e);
/* write the node's packed population into a DEVICE buffer of that size */
int32_t dcsmc_population_pack(dcsmc_engine* e, int32_t node, void* deviceDst);
/* install a packed population (DEVICE buffer) as `node`'s finished population */
int32_t dcsmc_population_unpack(dcsmc_engine* e, int32_t node, const void* deviceSrc);
/* release a node's population (the parent consumed it) */
int32_t dcsmc_population_release(dcsmc_engine* e, int32_t node);
/* forget every resident population and statistic (dcsmc_run does this itself; callers of
 * dcsmc_run_subtrees / dcsmc_run_nodes use it between independent executions) */
int32_t dcsmc_reset_populations(dcsmc_engine* e);

/* ---- zero-copy sharing between the engines of one NVLink/NVSwitch node ----------------------
 * Instead of pack -> send -> unpack, the consuming engine maps the producing engine's population
 * pool (CUDA IPC) and its merge kernel reads the child's columns — through the child's ancestor
 * indices, i.e. the gather of the resampled population is fused into the NVLink read — directly
 * from peer memory. Replaces the same Hazelcast put/remove (DCRecursionTask.java:37,113); only
 * DCSMC_EXPORT_DESC_BYTES per population cross the host. The producer must keep the population
 * resident (no release, no new run) until the consumer's run call has returned. */
#define DCSMC_IPC_HANDLE_BYTES 64
#define DCSMC_EXPORT_DESC_BYTES 128
/* IPC handle of this engine's pool (allocates it if needed), the pool's generation (changes
 * whenever the pool is re-allocated; peers must then re-open) and its size in bytes */
int32_t dcsmc_peer_pool_handle(dcsmc_engine* e, void* handle, int64_t* generation, int64_t* poolBytes);
/* map a peer engine's pool (no-op if that generation is already mapped); poolBytes bounds every descriptor
 * later imported from that peer */
int32_t dcsmc_peer_open(dcsmc_engine* e, int32_t peer, const void* handle, int64_t generation, int64_t poolBytes);
/* descriptors (n * DCSMC_EXPORT_DESC_BYTES host bytes) of finished, resident populations */
int32_t dcsmc_population_export(dcsmc_engine* e, const int32_t* nodes, int32_t n, void* descs);
/* install peer-resident populations as the finished populations of `nodes` (peers[k] = the id
 * given to dcsmc_peer_open for the owner of nodes[k]). Every descriptor is validated before anything changes:
 * same node, model, layout and nParticles, finished, and offset + size inside the mapped pool of that peer. */
int32_t dcsmc_population_import_peer(dcsmc_engine* e, const int32_t* nodes, const int32_t* peers,
                                     int32_t n, const void* descs);

/* ---- sharded runs without host round trips ---------------------------------------------------
 * The schedule above the cut is static (same tree, same plan, same pool offsets every run), so from the second run on
 * nothing but the 64-byte statistics of the exported populations has to move between ranks, and it can move on
 * the device: producers publish the records (dcsmc_stats_export_device) into a device buffer, ONE small
 * all-gather ordered on the engine's stream (dcsmc_stream; NCCL or any stream-ordered transport) carries them —
 * and is the "producers are done" barrier —, consumers install them with the cached layout descriptors
 * (dcsmc_population_import_peer_device). Between dcsmc_begin_deferred and dcsmc_sync the run calls only
 * enqueue work: the host synchronises once per run, not once per level (DCRecursionTask.java:77-93,108-115 is
 * the reference's per-node hand-over through the cluster map). */
#define DCSMC_STATS_RECORD_BYTES 64
/* the engine's launching stream (a cudaStream_t): order collectives after / before the engine's kernels on it */
int32_t dcsmc_stream(dcsmc_engine* e, void** cudaStream);
/* run calls (dcsmc_run_subtrees, dcsmc_run_nodes) return after enqueueing; statistics and dcsmc_last_run_info
 * (accumulated over the calls) become valid at dcsmc_sync */
int32_t dcsmc_begin_deferred(dcsmc_engine* e);
int32_t dcsmc_sync(dcsmc_engine* e);
/* write the statistics records (DCSMC_STATS_RECORD_BYTES each) of resident `nodes` into a DEVICE buffer, in order */
int32_t dcsmc_stats_export_device(dcsmc_engine* e, const int32_t* nodes, int32_t n, void* deviceDst);
/* descriptors (n * DCSMC_EXPORT_DESC_BYTES host bytes) of resident populations WITHOUT their statistics: where the
 * population lies in this engine's pool. Host bookkeeping only — valid while the populations are still being computed. */
int32_t dcsmc_population_export_layout(dcsmc_engine* e, const int32_t* nodes, int32_t n, void* descs);
/* dcsmc_population_import_peer with the statistics of nodes[k] read on the device: record statIndex[k] of the
 * DEVICE buffer deviceStats (nRecords records, e.g. the all-gathered exports of every rank). Same validation. */
int32_t dcsmc_population_import_peer_device(dcsmc_engine* e, const int32_t* nodes, const int32_t* peers, int32_t n,
                                            const void* descs, const void* deviceStats, const int32_t* statIndex,
                                            int32_t nRecords);

/* The per-node statistics of the nodes THIS engine computed since the last reset, zeros elsewhere, as a DEVICE
 * matrix of doubles [6][nNodes]: ess, relEss, logNormEstimate, logScaling, resampled, computedHere. Summed over the
 * ranks of a sharded run (an all-reduce ordered on dcsmc_stream) it is the whole tree's `timing` table
 * (DefaultProcessorFactory.java:45-52) without any per-node work on the host. */
int32_t dcsmc_stats_matrix_device(dcsmc_engine* e, void* deviceDst);

/* ---- planning without a device ---------------------------------------------------------
 * The schedule and the memory plan are a pure function of (options, tree, leaf kinds, pool size): this
 * entry point runs the planner on the host only — no CUDA device is needed or touched — and reports what
 * dcsmc_run would do. Use it to size a job (does BASELINE config 5 fit 180 GB per GPU? how many subtree
 * batches does a memory budget cost?) before allocating anything. */
typedef struct dcsmc_plan_info {
  int64_t poolBytesNeeded; /* every population resident at once (the pool dcsmc_run would like) */
  int64_t poolBytes;       /* the pool the plan was made for: min(needed, budget) */
  int64_t peakBytes;       /* high-water mark of the plan inside that pool */
  int64_t scratchBytes;    /* per-step scratch of the largest step: CDF, offspring counts, index, descriptors */
  int64_t tableBytes;      /* per-node device tables (descriptors, statistics, leaf constants, seeds) */
  int64_t nodesPlanned;
  int32_t levelBatches;    /* groups of launches (one per level when everything fits) */
  int32_t maxBatchNodes;   /* sibling nodes in the largest batch */
  int32_t feasible;        /* 0: even one node with its children does not fit the budget */
  int32_t reserved;
} dcsmc_plan_info;
/* leafKind: per node (DCSMC_LEAF_*), NULL = all binomial; model: DCSMC_MODEL_*; roots/nRoots: the
 * subtrees to plan as dcsmc_run_subtrees would (one rank's share of a sharded run), NULL/0 = the whole
 * tree as dcsmc_run; budgetBytes: 0 = unlimited. err (optional, errLen bytes) receives the message on failure. */
int32_t dcsmc_plan_query(const dcsmc_options* opt, int32_t nNodes, const int32_t* childOffsets,
                         const int32_t* children, const int32_t* leafKind, int32_t model,
                         const int32_t* roots, int32_t nRoots, int64_t budgetBytes, dcsmc_plan_info* out,
                         char* err, int32_t errLen);

/* ---- measurement ---------------------------------------------------------------------- */

enum {
  DCSMC_K_LEAF = 0,     /* leaf propose (MultiLevelLeafProposal.propose) */
  DCSMC_K_INTERNAL = 1, /* merge + propose + weight (DCRecursion.dcPropose) */
  DCSMC_K_SCAN = 2,     /* expNormalize + CDF + ESS partials */
  DCSMC_K_FINALIZE = 3, /* logScaling / ESS / resample decision per node */
  DCSMC_K_DARTS = 4,    /* multinomial: dart -> ancestor -> offspring count (no sort) */
  DCSMC_K_RESAMPLE = 5, /* ancestor search */
  DCSMC_K_GATHER = 6,   /* materialise a resampled population */
  DCSMC_K_MARKOV = 7,   /* Doc toy-model propose */
  DCSMC_K_COUNT = 8
};

typedef struct dcsmc_run_info {
  double deviceMs;         /* CUDA-event time of the last run call */
  int64_t kernelLaunches;  /* kernels of this library launched by the last run call */
  int64_t particleUpdates; /* nodes processed * nParticles */
  int64_t nodesProcessed;
  int64_t algorithmicBytes; /* sum over nodes of N*(40*C + 184 [+16 multinomial]) */
  int64_t poolBytes;        /* device bytes reserved for populations */
  int64_t h2dBytes;         /* host->device bytes copied by the last run call (tables, uniforms) */
  int64_t d2hBytes;         /* device->host bytes copied by the last run call (node stats) */
  int32_t levelBatches;
  int32_t reserved;
} dcsmc_run_info;

int32_t dcsmc_last_run_info(const dcsmc_engine* e, dcsmc_run_info* out);

/* CUDA events on the engine's launching stream, bracketing any sequence of calls on this engine
 * (bench.py: K whole runs incl. the gaps between them). deviceMs of dcsmc_run_info covers one call. */
int32_t dcsmc_timer_begin(dcsmc_engine* e);
int32_t dcsmc_timer_end(dcsmc_engine* e, double* ms);

/* per-kernel-class CUDA-event timing (adds an event pair around every launch) */
int32_t dcsmc_profile_enable(dcsmc_engine* e, int32_t enable);
int32_t dcsmc_profile_get(const dcsmc_engine* e, int32_t kernelClass, double* ms,
                          int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
