// engine.cu — host runtime of the DC-SMC node engine + the C ABI of include/dcsmc.h.
//
// Replaces, for the hot path only, the reference's L3/L4 machinery
// (dc/DistributedDC.java, dc/DCRecursionTask.java): instead of one Hazelcast task per node the
// tree is cut into level batches ("steps"); every step is a fixed sequence of kernel launches
// over all its sibling nodes. A static memory plan decides which populations are resident
// when; child populations are released as soon as their parent has consumed them
// (DCRecursionTask.java:113).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dcsmc.h"
#include "kernels.cuh"

using namespace dcsmc;

namespace {

thread_local std::string g_createError;

struct Step {
  int offset = 0;  // into stepNodes
  int count = 0;
  bool leaf = false;
  std::vector<int> releaseAfter;  // nodes whose slots are returned after this step
};

struct Slot {
  size_t offset = 0;
  size_t bytes = 0;
  char* external = nullptr;  // unpacked populations live in their own allocation, not in the pool
  bool peer = false;         // `external` points into a peer engine's pool (not owned, never recycled)
};

struct Plan {
  std::vector<Step> steps;
  std::vector<int32_t> stepNodes;
  // the same nodes in ascending id order: the order in which their statistics leave the device, so that the host
  // scatters them into its per-node tables front to back (in plan order — leaves first, level by level — the scatter
  // of a 131 071-node tree was 3 ms of cache misses per run)
  std::vector<int32_t> exportOrder;
};

// A plan is a pure function of (tree, model, options, roots, pool state at entry): cached together with the pool
// state it leaves behind, so repeated runs skip planning and descriptor uploads and replay a CUDA graph.
// Plans made on an EMPTY pool (dcsmc_run, a rank's subtrees) store the whole per-node tables; CONTEXTUAL plans
// (made with populations resident: the nodes above the cut of a sharded run, dcsmc_run_nodes) are keyed by the
// entry state of the allocator and store only the nodes they touch.
struct NodeDelta {
  int node;
  Slot slot;
  char resident, imported;
};
struct CachedPlan {
  std::vector<int> roots;
  Plan plan;
  bool contextual = false;
  std::map<size_t, size_t> entryFree;  // contextual: free holes / top of the pool when the plan was made
  size_t entryTop = 0;
  uint64_t entryKids = 0;              // contextual: residency of the roots' children (decides what gets planned)
  std::map<size_t, size_t> freeBySize;
  size_t poolTop = 0;
  std::vector<Slot> slot;
  std::vector<char> resident, imported;
  std::vector<NodeDelta> delta;        // contextual: state of every node the plan touched, after the plan
  std::vector<char> inPlan;  // per node: computed by this plan
  std::vector<int> finalResident;  // nodes still resident when the plan has run (its roots, unless retained)
  int32_t* dStepNodes = nullptr;   // device copy of plan.stepNodes
  std::vector<std::pair<int, int>> leafRanges;  // id ranges of the leaf entries the plan reads
  uint64_t leafVersion = 0;        // leaf data version those entries were last uploaded at
  bool onDevice = false;           // step list + descriptors of the plan's nodes are on the device, untouched since
  // the plan's whole launch sequence as a CUDA graph (captured on its first replay)
  mutable cudaGraphExec_t graphExec = nullptr;
  mutable uint64_t graphScratchGen = 0;
  mutable dcsmc_run_info graphInfo{};
};
constexpr size_t MAX_CACHED_PLANS = 24;

}  // namespace

struct dcsmc_engine {
  dcsmc_options opt{};
  mutable std::string err;
  cudaStream_t stream = nullptr;   // launching stream: propose kernels, copies, timing events
  cudaStream_t streamR = nullptr;  // higher priority: the normalise/resample chain of a level chunk
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> evPool;  // untimed events ordering the two streams
  size_t evNext = 0;
  int smCount = 148;
  int overlapChunks = 1;
  bool useGraphs = true;       // replay cached plans as CUDA graphs (DCSMC_GRAPHS=0 disables)
  uint64_t scratchGen = 1;     // bumped when a scratch buffer baked into captured launches moves            // level batches are split in this many chunks (1 = one stream)

  // topology / model (host)
  int nNodes = 0, root = 0;
  std::vector<int32_t> childOffsets, children, parent, height, javaHash, subtreeSize;
  bool preorder = false;  // nodes are numbered in pre-order (subtree of r = ids [r, r + subtreeSize[r]))
  int model = -1;
  std::vector<int32_t> leafKind, nTrials, nSuccesses;
  std::vector<double> leafObs;
  double rate = 1.0;
  int nStates = 0;
  std::vector<double> markovT, markovPrior;
  std::vector<std::vector<double>> injHost;  // kept only until uploaded
  std::vector<double*> injDev;

  // device tables
  NodeDev* dNodes = nullptr;
  NodeStats* dStats = nullptr;
  int32_t* dErr = nullptr;           // device-side consistency flag of deferred imports (read at dcsmc_sync)
  int32_t* hErr = nullptr;           // pinned
  LeafConst* dLeaf = nullptr;
  double* dLeafObs = nullptr;             // observation of every Gaussian leaf, by node id (8 B per node instead of a LeafConst)
  int32_t* dChildren = nullptr;
  int32_t* dStepNodesTmp = nullptr;  // step lists of the current uncached plan
  size_t stepNodesTmpCap = 0;
  int32_t* curStepNodes = nullptr;   // the one the running plan uses
  const int32_t* planStepNodes = nullptr;  // host copy of the same lists (valid while a plan executes)
  double** dInj = nullptr;
  uint64_t* dJavaState = nullptr;
  double *dMarkovT = nullptr, *dMarkovPrior = nullptr;
  bool tablesDirty = true;
  // the leaf data (trial / success counts, observations) is the per-run INPUT: a new set on the same tree and leaf
  // kinds bumps leafVersion; a plan uploads the entries of the leaves it reads (its subtrees' id ranges) when its
  // own copy is older — a rank of a sharded run never touches the other ranks' leaves
  uint64_t leafVersion = 1, leafFullVersion = 0;
  std::vector<uint64_t> leafVer;  // per node: version of its entry on the device
  size_t tableNodesCap = 0, tableEdgesCap = 0;  // sizes the device tables were allocated for

  // population pool
  char* pool = nullptr;
  size_t poolBytes = 0;
  bool poolIsBudget = false;  // pool was capped by the memory budget
  size_t poolNeed = 0;        // bytes of "everything resident at once" (0 = not computed yet)
  std::map<size_t, size_t> freeBySize;  // free holes below poolTop: offset -> bytes (coalesced)
  size_t poolTop = 0;
  size_t poolPeak = 0;           // high-water mark of poolTop (dcsmc_plan_query)
  std::vector<Slot> slot;        // per node
  std::vector<char> resident;    // per node
  std::vector<char> imported;    // per node: population installed by dcsmc_population_unpack
  std::vector<char*> importFree;  // recycled buffers of unpacked populations (all the same size)
  struct Undo { int node; Slot slot; char resident, imported; };
  std::vector<Undo> undo;         // per-node pool changes since the planner's last checkpoint
  bool undoOn = false;
  std::vector<NodeDev> hNodes;   // host mirror
  std::vector<dcsmc_node_stats_t> hStats;
  std::vector<NodeStatsOut> hOut;   // per node: last fetched record (incl. m, S for exports)
  std::vector<int> statNodes;       // nodes whose hStats entry was written since the last reset
  std::vector<int32_t> pendNodes;   // nodes of the last plan whose records still sit in hExport (flushStats)
  std::vector<int> externalNodes;   // nodes that got an external slot since the last reset
  struct PeerMap { char* base = nullptr; int64_t generation = -1; int64_t bytes = 0; };
  std::map<int, PeerMap> peers;     // peer engines' pools mapped through CUDA IPC
  // identifies one allocation of the pool: starts from a per-engine random base (a peer process that
  // destroys its engine and creates a new one must never collide with a mapping a consumer still holds)
  // and is bumped whenever the pool is (re)allocated
  int64_t poolGeneration = 0;
  char *dImp = nullptr, *hImp = nullptr;  // dcsmc_population_import_peer staging
  LeafConst* hLeafStage = nullptr;        // pinned staging of the leaf table
  double* hObsStage = nullptr;            // ... and of the observations
  std::vector<int32_t> binomPrefix;       // binomial leaves among the nodes [0, n): which id ranges need LeafConst entries
  size_t leafStageCap = 0;
  size_t impCap = 0;
  cudaEvent_t evImp = nullptr;            // the last upload out of hImp
  NodeStatsOut *dExport = nullptr, *hExport = nullptr;  // statistics of the last plan (device / pinned host)
  size_t exportCap = 0;
  cudaEvent_t evT0 = nullptr, evT1 = nullptr;           // dcsmc_timer_begin / dcsmc_timer_end
  int Npad = 0;

  // posterior summaries (levelCutOffForOutput > 0)
  std::vector<int32_t> depth;        // level() of every node
  NodeSummary* dSumm = nullptr;      // per node
  double* dSumScratch = nullptr;     // sample series of one summary batch (or of every summary node: retainSamples)
  size_t sumScratchCap = 0;          // in doubles
  std::vector<int64_t> sampleBase;   // retainSamples: first series of a node inside dSumScratch (-1: none)
  size_t sampleSeriesTotal = 0;
  std::vector<double> hNodeMs;       // per node: device time of its level batch / nodes of the batch (profiled runs)

  // scratch
  double* dCdf = nullptr;
  size_t cdfCap = 0;
  unsigned long long* dDesc = nullptr;  // descA | descB | ticket
  size_t descCap = 0;
  int32_t* dInvTable = nullptr;  // inverse-CDF index of the internal nodes of a step
  size_t invTableCap = 0;
  int invB = 1;
  size_t leafSmemSet[3] = {0, 0, 0};
  size_t leaf2SmemSet[3] = {0, 0, 0};
  size_t smallSmemSet[3] = {0, 0, 0};
  size_t smallSmemSet512[3] = {0, 0, 0};
  size_t fusedSmemSet[3] = {0, 0, 0};
  size_t fusedSmemSet512[3] = {0, 0, 0};
  bool fuseSmall = true;       // k_node_fused for small populations (DCSMC_FUSE_SMALL=0: propose + k_node_small)
  int fuseMinNodes = 148;      // ... for level batches of at least this many nodes (DCSMC_FUSE_MIN_NODES)
  std::vector<std::unique_ptr<CachedPlan>> planCache;
  // deferred mode (dcsmc_begin_deferred .. dcsmc_sync): run calls only enqueue; one host synchronisation at the end
  bool deferred = false;
  bool inFlight = false;           // statistics copies into hExport may still be running
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> callEvents;  // one timed pair per deferred call
  size_t callEventsUsed = 0;
  // device copies of small host arrays that repeat run after run (static schedule): keyed by content
  std::map<std::string, char*> blobCache;
  char *dStage = nullptr, *hStage = nullptr;  // dcsmc_population_copy_to_host staging (device / pinned host)
  size_t stageCap = 0;
  std::vector<cudaEvent_t> stageEvents;  // one per 4 MiB piece of a pipelined read-back
  int32_t* dOffspring = nullptr;  // multinomial: offspring counts
  size_t offspringCap = 0;
  // partitioned multinomial resampling of large populations (k_dart_partition / k_seg_resample)
  bool partitionDarts = true;     // DCSMC_PARTITION_DARTS=0: k_darts_count + k_expand
  int leafVariant = 3;            // DCSMC_LEAF_V=1: k_leaf_propose (r01), 2: k_leaf_propose2, 3: k_leaf_propose3
  int leafChunkMax = 4096;        // DCSMC_LEAF_CH: particles a warp of k_leaf_propose3 works through before it drains
  int leafKidsKernel = 2;         // DCSMC_LEAFKIDS=0: the generic merge kernel also for the level above the leaves;
                                  // 1: k_internal_propose_lk with the literal fold; 2: with the shared-reciprocal fold
  int l2Prefetch = 3;             // DCSMC_L2_PREFETCH: bit 0 prefetch hints in k_internal_propose, bit 1 in k_internal_propose_lk
  int lk2Range = 100;             // DCSMC_LK2_RANGE (tests): see RunCtx::lkRange
  int scanVariant = 2;            // DCSMC_SCAN_V=1: k_scan (r01) instead of k_scan2
  double* dDartBuf = nullptr;
  size_t dartBufCap = 0;
  int32_t* dChunkBin = nullptr;
  size_t chunkBinCap = 0;
  int32_t* dSegTotal = nullptr;
  size_t segTotalCap = 0;
  bool segSmemSet = false;

  // measurement
  dcsmc_run_info info{};
  int64_t pendingH2D = 0;  // table uploads since the last run, charged to the next run
  bool profile = false;
  double profMs[DCSMC_K_COUNT] = {0};
  int64_t profLaunches[DCSMC_K_COUNT] = {0};
  struct ProfEvent { int cls, step; cudaEvent_t a, b; };
  std::vector<ProfEvent> profEvents;
  int curStep = -1;  // level batch being launched (index into the plan's steps)

  int fail(int code, const char* fmt, ...) const {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      (void)cudaGetLastError(); /* reported here: do not leave it pending for the next call (e.g. a refused \
                                   cudaIpcOpenMemHandle must not poison the NCCL fallback) */  \
      return e->fail(DCSMC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),     \
                     __FILE__, __LINE__);                                                     \
    }                                                                                         \
  } while (0)

namespace {

int nChildrenOf(const dcsmc_engine* e, int n) { return e->childOffsets[n + 1] - e->childOffsets[n]; }

int nodeKind(const dcsmc_engine* e, int n) {
  if (nChildrenOf(e, n) > 0) return KIND_INTERNAL;
  if (e->model == DCSMC_MODEL_MARKOV) return KIND_LEAF_GAUSS;  // no proposal draws
  if (!e->leafKind.empty() && e->leafKind[n] == DCSMC_LEAF_GAUSS_OBS) return KIND_LEAF_GAUSS;
  return KIND_LEAF_BINOMIAL;
}

int slotsPerParticle(const dcsmc_engine* e, int n) {
  const int k = nodeKind(e, n);
  if (k == KIND_INTERNAL) return 1;
  if (k == KIND_LEAF_BINOMIAL) return 2 * e->opt.maxBetaAttempts;
  return 0;
}

size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Observed (Gaussian) leaves hold N copies of one datum: unless populations are retained for
// inspection (or the parity mode wants every array) they get no slot and no kernel work.
bool skipObservedLeaves(const dcsmc_engine* e) {
  return e->model == DCSMC_MODEL_MULTILEVEL && !e->opt.retainPopulations && e->opt.sumMode == DCSMC_SUM_EXACT;
}

// bytes of one node's resident population. Leaves carry no log-weight column (all equal).
size_t slotBytes(const dcsmc_engine* e, int n) {
  const size_t NP = (size_t)e->Npad;
  const bool internal = nodeKind(e, n) == KIND_INTERNAL;
  if (!internal && nodeKind(e, n) == KIND_LEAF_GAUSS && skipObservedLeaves(e)) return 0;
  size_t b = 0;
  if (e->model == DCSMC_MODEL_MARKOV) {
    b = NP * 4 /*istate*/ + NP * 4 /*anc*/ + (internal ? NP * 8 : 0) /*logw*/;
  } else {
    b = NP * 8 * (internal ? 6 : 2) + NP * 4 /*anc*/ + (internal ? NP * 8 : 0) /*logw*/;
  }
  return alignUp(b, 256);
}

void poolFree(dcsmc_engine* e, int n);

// bytes of an unpacked population: always the internal layout (6 columns + ancestors + logw)
size_t importBytes(const dcsmc_engine* e) {
  const size_t NP = (size_t)e->Npad;
  return alignUp(e->model == DCSMC_MODEL_MARKOV ? NP * 16 : NP * 60, 256);
}

// Population pool: a bump pointer (poolTop) plus coalesced free holes below it. Best fit over the
// holes first, then the hole touching poolTop is extended, then fresh space. All of this runs at
// plan time on the host; the device only ever sees final addresses.

bool poolAlloc(dcsmc_engine* e, int n) {
  poolFree(e, n);  // recomputing a node reuses nothing of its old population
  if (e->undoOn) e->undo.push_back({n, e->slot[n], e->resident[n], e->imported[n]});
  const size_t b = slotBytes(e, n);
  if (b == 0) {
    e->slot[n] = Slot{0, 0};
    e->resident[n] = 1;
    return true;
  }
  auto best = e->freeBySize.end();
  for (auto it = e->freeBySize.begin(); it != e->freeBySize.end(); ++it)
    if (it->second >= b && (best == e->freeBySize.end() || it->second < best->second)) best = it;
  size_t off;
  if (best != e->freeBySize.end()) {
    off = best->first;
    const size_t rest = best->second - b;
    e->freeBySize.erase(best);
    if (rest) e->freeBySize.emplace(off + b, rest);
  } else {
    // extend the hole that ends at poolTop, if any
    size_t start = e->poolTop;
    if (!e->freeBySize.empty()) {
      auto last = std::prev(e->freeBySize.end());
      if (last->first + last->second == e->poolTop) start = last->first;
    }
    if (start + b > e->poolBytes) return false;
    if (start != e->poolTop) e->freeBySize.erase(std::prev(e->freeBySize.end()));
    off = start;
    e->poolTop = start + b;
  }
  e->slot[n] = Slot{off, b};
  e->resident[n] = 1;
  return true;
}

void poolFree(dcsmc_engine* e, int n) {
  if (!e->resident[n]) return;
  if (e->undoOn) e->undo.push_back({n, e->slot[n], e->resident[n], e->imported[n]});
  e->resident[n] = 0;
  if (e->slot[n].external) {  // unpacked population: recycle its buffer (peer memory is not ours)
    if (!e->slot[n].peer) e->importFree.push_back(e->slot[n].external);
    e->slot[n].external = nullptr;
    e->slot[n].peer = false;
    e->imported[n] = 0;
    return;
  }
  size_t off = e->slot[n].offset, len = e->slot[n].bytes;
  if (len == 0) return;
  auto next = e->freeBySize.lower_bound(off);
  if (next != e->freeBySize.begin()) {
    auto prev = std::prev(next);
    if (prev->first + prev->second == off) {
      off = prev->first;
      len += prev->second;
      e->freeBySize.erase(prev);
    }
  }
  if (next != e->freeBySize.end() && off + len == next->first) {
    len += next->second;
    e->freeBySize.erase(next);
  }
  if (off + len == e->poolTop)
    e->poolTop = off;  // the hole touches the top: give it back
  else
    e->freeBySize.emplace(off, len);
}

void fillNodeDev(dcsmc_engine* e, int n) {
  NodeDev& d = e->hNodes[n];
  const size_t NP = (size_t)e->Npad;
  char* base = e->slot[n].external ? e->slot[n].external : e->pool + e->slot[n].offset;
  d.childBegin = e->childOffsets[n];
  d.nChildren = nChildrenOf(e, n);
  d.kind = nodeKind(e, n);
  d.pad = 0;
  // unpacked (imported == 1): 6 columns + logw; peer-mapped (imported == 2): the owner's native layout
  if (!e->imported.empty() && e->imported[n] == 1) d.kind = KIND_INTERNAL;
  const bool internal = d.kind == KIND_INTERNAL;
  if (e->model == DCSMC_MODEL_MARKOV) {
    d.state = nullptr;
    d.istate = reinterpret_cast<int32_t*>(base);
    d.anc = reinterpret_cast<int32_t*>(base + NP * 4);
    d.logw = internal ? reinterpret_cast<double*>(base + NP * 8) : nullptr;
  } else {
    const int cols = internal ? 6 : 2;
    d.state = reinterpret_cast<double*>(base);
    d.anc = reinterpret_cast<int32_t*>(base + NP * 8 * cols);
    d.logw = internal ? reinterpret_cast<double*>(base + NP * 8 * cols + NP * 4) : nullptr;
    d.istate = nullptr;
  }
}

void leafConstFor(const dcsmc_engine* e, int node, LeafConst& lc) {
  memset(&lc, 0, sizeof lc);
  if (nodeKind(e, node) != KIND_LEAF_BINOMIAL) return;  // observed leaves: their datum lives in dLeafObs
  // MultiLevelProposalFactory.build (MultiLevelProposalFactory.java:40-49): Beta(1+s, 1+(n-s)),
  // then the constants of Cheng's BB/BC samplers (commons-math3 BetaDistribution.sample()).
  const int32_t n = e->nTrials[node], s = e->nSuccesses[node];
  const double shapeA = 1 + s, shapeB = 1 + (n - s);
  lc.n = n;
  lc.s = s;
  lc.a0 = shapeA;
  const double mn = std::fmin(shapeA, shapeB), mx = std::fmax(shapeA, shapeB);
  lc.useBB = mn > 1;
  if (lc.useBB) {
    lc.a = mn;
    lc.b = mx;
    lc.alpha = lc.a + lc.b;
    lc.beta = std::sqrt((lc.alpha - 2.) / (2. * lc.a * lc.b - lc.alpha));
    lc.gamma = lc.a + 1. / lc.beta;
  } else {
    lc.a = mx;
    lc.b = mn;
    lc.alpha = lc.a + lc.b;
    lc.beta = 1. / lc.b;
    lc.delta = 1. + lc.a - lc.b;
    lc.k1 = lc.delta * (0.0138889 + 0.0416667 * lc.b) / (lc.a * lc.beta - 0.777778);
    lc.k2 = 0.25 + (0.5 + 0.25 / lc.delta) * lc.b;
  }
  lc.logAlpha = dc_log(lc.alpha);
  // SpecialFunctions.logBinomial(n, s): per-node constant, hoisted out of the particle loop
  lc.logBinom = std::lgamma((double)n + 1.0) - std::lgamma((double)s + 1.0) - std::lgamma((double)(n - s) + 1.0);
}

int ensureDevice(dcsmc_engine* e) {
  CK(cudaSetDevice(e->opt.device));
  return DCSMC_OK;
}

// (re)build the device-side tables that do not depend on the memory plan
void clearPlanCache(dcsmc_engine* e) {
  for (auto& cp : e->planCache) {
    if (cp->graphExec) cudaGraphExecDestroy(cp->graphExec);
    if (cp->dStepNodes) cudaFree(cp->dStepNodes);
  }
  e->planCache.clear();
  for (auto& kv : e->blobCache) cudaFree(kv.second);
  e->blobCache.clear();
}

// the device descriptor of `node` was rewritten (by another plan, an unpack, an import): cached plans that
// compute this node must upload their descriptors again before their next (graph) replay
void descWritten(dcsmc_engine* e, int node, const CachedPlan* except) {
  for (auto& cp : e->planCache)
    if (cp.get() != except && cp->onDevice && cp->inPlan[node]) cp->onDevice = false;
}
void allDescWritten(dcsmc_engine* e, const CachedPlan* except) {
  for (auto& cp : e->planCache)
    if (cp.get() != except) cp->onDevice = false;
}

// device copy of a small host array whose content repeats run after run (node lists and descriptors of the
// static schedule above the cut): uploaded once, found again by content
int cachedBlob(dcsmc_engine* e, const void* data, size_t bytes, char** out) {
  std::string key(static_cast<const char*>(data), bytes);
  auto it = e->blobCache.find(key);
  if (it != e->blobCache.end()) {
    *out = it->second;
    return DCSMC_OK;
  }
  if (e->blobCache.size() > 4096) {  // a caller whose lists never repeat: start over (after the device is done with them)
    CK(cudaStreamSynchronize(e->stream));
    for (auto& kv : e->blobCache) cudaFree(kv.second);
    e->blobCache.clear();
  }
  char* d = nullptr;
  CK(cudaMalloc((void**)&d, std::max<size_t>(bytes, 16)));
  CK(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, e->stream));  // pageable source: staged before the call returns
  e->pendingH2D += (int64_t)bytes;
  e->blobCache.emplace(std::move(key), d);
  *out = d;
  return DCSMC_OK;
}

// per-leaf constants of the model (the per-run INPUT DATA: trial / success counts or observations) for the nodes
// [lo, hi). Staged in pinned memory (entry n at offset n, so ranges never overlap); a previous copy out of the
// same entries has completed (every run ends with a sync).
int uploadLeafRange(dcsmc_engine* e, int lo, int hi) {
  const int nN = e->nNodes;
  if ((size_t)nN > e->leafStageCap) {
    if (e->hLeafStage) cudaFreeHost(e->hLeafStage);
    e->hLeafStage = nullptr;
    e->leafStageCap = 0;
    if (e->hObsStage) cudaFreeHost(e->hObsStage);
    e->hObsStage = nullptr;
    CK(cudaMallocHost((void**)&e->hLeafStage, sizeof(LeafConst) * (size_t)nN));
    CK(cudaMallocHost((void**)&e->hObsStage, sizeof(double) * (size_t)nN));
    e->leafStageCap = (size_t)nN;
  }
  if ((int)e->leafVer.size() != nN) e->leafVer.assign(nN, 0);
  for (int n = lo; n < hi; ++n) e->leafVer[n] = e->leafVersion;
  // binomial leaves carry a LeafConst (Cheng's constants, logBinomial); an observed leaf is one double. Ranges without
  // binomial leaves (BASELINE config 4: 65 536 observed leaves) move 8 bytes per node instead of 104.
  const bool multilevel = e->model == DCSMC_MODEL_MULTILEVEL;
  const bool anyBinom = multilevel && (int)e->binomPrefix.size() == nN + 1 && e->binomPrefix[hi] > e->binomPrefix[lo];
  const bool anyObs = multilevel && !e->leafObs.empty();
  if (anyBinom || !multilevel) {
    LeafConst* lc = e->hLeafStage;
    if (multilevel)
      for (int n = lo; n < hi; ++n) leafConstFor(e, n, lc[n]);
    else
      memset(lc + lo, 0, sizeof(LeafConst) * (size_t)(hi - lo));
    CK(cudaMemcpyAsync(e->dLeaf + lo, lc + lo, sizeof(LeafConst) * (size_t)(hi - lo), cudaMemcpyHostToDevice, e->stream));
    e->pendingH2D += (int64_t)(sizeof(LeafConst) * (size_t)(hi - lo));
  }
  if (anyObs) {
    memcpy(e->hObsStage + lo, e->leafObs.data() + lo, sizeof(double) * (size_t)(hi - lo));
    CK(cudaMemcpyAsync(e->dLeafObs + lo, e->hObsStage + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyHostToDevice, e->stream));
    e->pendingH2D += (int64_t)(sizeof(double) * (size_t)(hi - lo));
  }
  if (getenv("DCSMC_HOSTPROF"))
    fprintf(stderr, "[dcsmc hostprof] leaf entries [%d, %d) uploaded (version %llu, LeafConst %d, observations %d)\n", lo, hi,
            (unsigned long long)e->leafVersion, (int)(anyBinom || !multilevel), (int)anyObs);
  return DCSMC_OK;
}
int uploadLeafTable(dcsmc_engine* e) {
  int rc = uploadLeafRange(e, 0, e->nNodes);
  if (rc == DCSMC_OK) e->leafFullVersion = e->leafVersion;
  return rc;
}

// The leaf entries a plan reads: its own nodes and their children (an imported or still resident child's entry
// is read for its observation), as contiguous id ranges — a rank's subtrees are a handful of ranges when the nodes
// are numbered in pre-order, as tree.py and ddc_main.cpp do.
std::vector<std::pair<int, int>> leafRangesOf(const dcsmc_engine* e, const Plan& plan) {
  std::vector<int> ids(plan.stepNodes.begin(), plan.stepNodes.end());
  for (int n : plan.stepNodes)
    for (int c = e->childOffsets[n]; c < e->childOffsets[n + 1]; ++c) ids.push_back(e->children[c]);
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  std::vector<std::pair<int, int>> out;
  for (int id : ids) {
    if (!out.empty() && out.back().second == id)
      out.back().second = id + 1;
    else
      out.push_back({id, id + 1});
  }
  return out;
}

// before a plan runs: upload the entries of its ranges that are older than the current leaf data
int uploadLeafForPlan(dcsmc_engine* e, const Plan& plan, CachedPlan* cached) {
  if (e->leafFullVersion == e->leafVersion) return DCSMC_OK;
  if (cached && cached->leafVersion == e->leafVersion) return DCSMC_OK;
  if ((int)e->leafVer.size() != e->nNodes) e->leafVer.assign(e->nNodes, 0);
  std::vector<std::pair<int, int>> tmp;
  if (cached && cached->leafRanges.empty()) cached->leafRanges = leafRangesOf(e, plan);
  if (!cached) tmp = leafRangesOf(e, plan);
  const std::vector<std::pair<int, int>>& ranges = cached ? cached->leafRanges : tmp;
  int rc;
  for (const auto& r : ranges) {
    bool stale = false;  // another plan of this run may already have brought the range up to date (shared children)
    for (int n = r.first; n < r.second && !stale; ++n) stale = e->leafVer[n] != e->leafVersion;
    if (stale && (rc = uploadLeafRange(e, r.first, r.second))) return rc;
  }
  if (cached) cached->leafVersion = e->leafVersion;
  return DCSMC_OK;
}

int uploadTables(dcsmc_engine* e) {
  if (!e->tablesDirty) return DCSMC_OK;
  if (getenv("DCSMC_HOSTPROF")) fprintf(stderr, "[dcsmc hostprof] full table upload (%d nodes), plans dropped\n", e->nNodes);
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree has not been called");
  if (e->model < 0) return e->fail(DCSMC_ESTATE, "no proposal model set (dcsmc_set_*_model)");
  const int nN = e->nNodes;
  // tables are re-allocated only when the tree grows: cudaFree/cudaMalloc synchronise the device, and
  // a caller that re-sends the same topology every step (bench.py's end-to-end leg) pays them each time
  const bool keep = e->dNodes != nullptr && (size_t)nN <= e->tableNodesCap && e->children.size() <= e->tableEdgesCap;
  auto re = [&](void** p, size_t bytes) -> cudaError_t {
    if (keep && *p) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    return cudaMalloc(p, bytes ? bytes : 8);
  };
  CK(re((void**)&e->dNodes, sizeof(NodeDev) * nN));
  CK(re((void**)&e->dStats, sizeof(NodeStats) * nN));
  if (e->opt.levelCutOffForOutput > 0) {
    CK(re((void**)&e->dSumm, sizeof(NodeSummary) * nN));
    CK(cudaMemsetAsync(e->dSumm, 0, sizeof(NodeSummary) * nN, e->stream));
  }
  if (!(keep && e->dLeaf)) {  // entries of nodes that are not binomial leaves are never uploaded: keep them defined
    CK(re((void**)&e->dLeaf, sizeof(LeafConst) * nN));
    CK(cudaMemsetAsync(e->dLeaf, 0, sizeof(LeafConst) * nN, e->stream));
  }
  if (!(keep && e->dLeafObs)) {  // never read for a node that is not an observed leaf; zeros keep every byte defined
    CK(re((void**)&e->dLeafObs, sizeof(double) * nN));
    CK(cudaMemsetAsync(e->dLeafObs, 0, sizeof(double) * nN, e->stream));
  }
  CK(re((void**)&e->dChildren, sizeof(int32_t) * std::max<size_t>(1, e->children.size())));
  CK(re((void**)&e->dInj, sizeof(double*) * nN));
  CK(re((void**)&e->dJavaState, sizeof(uint64_t) * nN));
  if (!keep) {
    e->tableNodesCap = (size_t)nN;
    e->tableEdgesCap = e->children.size();
  }
  CK(cudaMemcpyAsync(e->dChildren, e->children.data(), sizeof(int32_t) * e->children.size(),
                     cudaMemcpyHostToDevice, e->stream));
  e->pendingH2D += (int64_t)(sizeof(int32_t) * e->children.size() + sizeof(uint64_t) * nN);
  int rcLeaf;
  e->leafVersion++;
  e->leafVer.clear();
  if ((rcLeaf = uploadLeafTable(e))) return rcLeaf;
  std::vector<uint64_t> js(nN, 0);
  for (int n = 0; n < nN; ++n)
    js[n] = java_scramble(java_node_seed(e->opt.masterRandomSeed, e->javaHash.empty() ? 0 : e->javaHash[n]));
  CK(cudaMemcpyAsync(e->dJavaState, js.data(), sizeof(uint64_t) * nN, cudaMemcpyHostToDevice, e->stream));
  if (e->model == DCSMC_MODEL_MARKOV) {
    auto reAlways = [&](void** p, size_t bytes) -> cudaError_t {  // sized by nStates, not by the tree
      if (*p) cudaFree(*p);
      *p = nullptr;
      return cudaMalloc(p, bytes ? bytes : 8);
    };
    CK(reAlways((void**)&e->dMarkovT, sizeof(double) * e->markovT.size()));
    CK(reAlways((void**)&e->dMarkovPrior, sizeof(double) * e->markovPrior.size()));
    CK(cudaMemcpyAsync(e->dMarkovT, e->markovT.data(), sizeof(double) * e->markovT.size(),
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->dMarkovPrior, e->markovPrior.data(), sizeof(double) * e->markovPrior.size(),
                       cudaMemcpyHostToDevice, e->stream));
  }
  CK(cudaMemsetAsync(e->dStats, 0, sizeof(NodeStats) * nN, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->hNodes.assign(nN, NodeDev{});
  e->hStats.assign(nN, dcsmc_node_stats_t{});
  e->hOut.assign(nN, NodeStatsOut{});
  e->pendNodes.clear();
  e->statNodes.clear();
  e->externalNodes.clear();
  for (Slot& sl : e->slot)  // buffers of unpacked populations (their size depends on the model)
    if (sl.external && !sl.peer) cudaFree(sl.external);
  for (char* p : e->importFree) cudaFree(p);
  e->importFree.clear();
  e->slot.assign(nN, Slot{});
  e->poolNeed = 0;
  clearPlanCache(e);
  e->resident.assign(nN, 0);
  e->imported.assign(nN, 0);
  e->freeBySize.clear();
  e->poolTop = 0;
  e->tablesDirty = false;
  return DCSMC_OK;
}

int uploadInjected(dcsmc_engine* e) {
  if (e->opt.rngMode != DCSMC_RNG_INJECTED) return DCSMC_OK;
  if ((int)e->injDev.size() != e->nNodes) e->injDev.assign(e->nNodes, nullptr);
  for (int n = 0; n < e->nNodes; ++n) {
    if (n < (int)e->injHost.size() && !e->injHost[n].empty()) {
      if (e->injDev[n]) cudaFree(e->injDev[n]);
      CK(cudaMalloc((void**)&e->injDev[n], sizeof(double) * e->injHost[n].size()));
      CK(cudaMemcpyAsync(e->injDev[n], e->injHost[n].data(), sizeof(double) * e->injHost[n].size(),
                         cudaMemcpyHostToDevice, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      e->pendingH2D += (int64_t)(sizeof(double) * e->injHost[n].size());
      e->injHost[n].clear();
      e->injHost[n].shrink_to_fit();
    }
  }
  CK(cudaMemcpyAsync(e->dInj, e->injDev.data(), sizeof(double*) * e->nNodes, cudaMemcpyHostToDevice, e->stream));
  return DCSMC_OK;
}

int ensurePool(dcsmc_engine* e) {
  // the pool never needs more than "everything resident at once"
  if (e->poolNeed == 0)
    for (int n = 0; n < e->nNodes; ++n) e->poolNeed += slotBytes(e, n);
  const size_t all = std::max<size_t>(e->poolNeed, 256);
  if (e->pool && (e->poolBytes >= all || e->poolIsBudget)) return DCSMC_OK;
  if (e->pool) {
    cudaFree(e->pool);
    e->pool = nullptr;
  }
  size_t freeB = 0, totalB = 0;
  CK(cudaMemGetInfo(&freeB, &totalB));
  // leave room for the per-step scratch (CDF: 8 B per particle of the largest step)
  const size_t budget = e->opt.memoryBudgetBytes > 0 ? (size_t)e->opt.memoryBudgetBytes : (size_t)(freeB * 0.80);
  e->poolBytes = std::min(budget, all);
  e->poolIsBudget = e->poolBytes < all;
  CK(cudaMalloc((void**)&e->pool, e->poolBytes));
  e->poolGeneration++;  // peers holding a mapping of the old pool must re-open (dcsmc_peer_pool_handle)
  return DCSMC_OK;
}

// ---- planning --------------------------------------------------------------------------------

void collectSubtree(const dcsmc_engine* e, int r, std::vector<int>& out) {
  std::vector<int> stack{r};
  while (!stack.empty()) {
    const int n = stack.back();
    stack.pop_back();
    if (e->resident[n] && n != r) continue;  // already computed (e.g. unpacked) — treat as done
    out.push_back(n);
    for (int c = e->childOffsets[n]; c < e->childOffsets[n + 1]; ++c)
      if (!e->resident[e->children[c]]) stack.push_back(e->children[c]);
  }
}

// Emit level batches for the union of the subtrees under `roots`; allocates slots as it goes.
// Returns false (and rolls nothing back) if the pool overflows.
bool emitLevels(dcsmc_engine* e, const std::vector<int>& roots, Plan& plan, bool dryRun,
                size_t& peakTop) {
  std::vector<int> nodes;
  for (int r : roots) collectSubtree(e, r, nodes);
  std::map<int, std::vector<int>> byHeight;
  for (int n : nodes) byHeight[e->height[n]].push_back(n);
  for (auto& kv : byHeight) {
    std::vector<int>& lvl = kv.second;
    std::sort(lvl.begin(), lvl.end());
    Step st;
    st.leaf = kv.first == 0;
    st.offset = (int)plan.stepNodes.size();
    st.count = (int)lvl.size();
    for (int n : lvl) {
      e->imported[n] = 0;
      if (!poolAlloc(e, n)) return false;
      peakTop = std::max(peakTop, e->poolTop);
      plan.stepNodes.push_back(n);
    }
    if (!e->opt.retainPopulations)
      for (int n : lvl)
        for (int c = e->childOffsets[n]; c < e->childOffsets[n + 1]; ++c) {
          st.releaseAfter.push_back(e->children[c]);
          poolFree(e, e->children[c]);
        }
    plan.steps.push_back(std::move(st));
  }
  (void)dryRun;
  return true;
}

// Planner checkpoint: the free-hole map is small and copied; per-node arrays are rolled back from
// the undo log (copying them per trial would cost O(nNodes) even for a one-node plan).
struct PoolState {
  std::map<size_t, size_t> freeBySize;
  size_t poolTop;
  size_t undoMark;
  bool undoWasOn;
};
PoolState savePool(dcsmc_engine* e) {
  PoolState s{e->freeBySize, e->poolTop, e->undo.size(), e->undoOn};
  e->undoOn = true;
  return s;
}
void restorePool(dcsmc_engine* e, const PoolState& s) {
  e->freeBySize = s.freeBySize;
  e->poolTop = s.poolTop;
  while (e->undo.size() > s.undoMark) {
    const dcsmc_engine::Undo& u = e->undo.back();
    if (u.slot.external && !u.slot.peer) {  // the trial recycled this unpacked population's buffer: take it back
      auto it = std::find(e->importFree.begin(), e->importFree.end(), u.slot.external);
      if (it != e->importFree.end()) e->importFree.erase(it);
    }
    e->slot[u.node] = u.slot;
    e->resident[u.node] = u.resident;
    e->imported[u.node] = u.imported;
    e->undo.pop_back();
  }
  e->undoOn = s.undoWasOn;
}
void commitPool(dcsmc_engine* e, const PoolState& s) {
  e->undoOn = s.undoWasOn;
  if (!e->undoOn) e->undo.clear();
}

// Memory-aware recursive planner: try the whole root set level by level; if the pool
// overflows, halve the root set, or descend below a single root.
int planRoots(dcsmc_engine* e, const std::vector<int>& roots, Plan& plan) {
  if (roots.empty()) return DCSMC_OK;
  {
    const PoolState saved = savePool(e);
    Plan trial;
    trial.stepNodes = plan.stepNodes;
    size_t peak = 0;
    if (emitLevels(e, roots, trial, false, peak)) {
      for (auto& s : trial.steps) plan.steps.push_back(std::move(s));
      plan.stepNodes = std::move(trial.stepNodes);
      commitPool(e, saved);
      e->poolPeak = std::max(e->poolPeak, peak);
      return DCSMC_OK;
    }
    restorePool(e, saved);
  }
  if (roots.size() > 1) {
    const size_t half = roots.size() / 2;
    std::vector<int> a(roots.begin(), roots.begin() + half), b(roots.begin() + half, roots.end());
    int rc = planRoots(e, a, plan);
    if (rc) return rc;
    return planRoots(e, b, plan);
  }
  const int r = roots[0];
  if (nChildrenOf(e, r) == 0)
    return e->fail(DCSMC_ENOMEM, "population pool (%zu bytes) cannot hold node %d", e->poolBytes, r);
  std::vector<int> kids;
  for (int c = e->childOffsets[r]; c < e->childOffsets[r + 1]; ++c)
    if (!e->resident[e->children[c]]) kids.push_back(e->children[c]);
  int rc = planRoots(e, kids, plan);
  if (rc) return rc;
  // all children resident now: the root alone
  Step st;
  st.leaf = false;
  st.offset = (int)plan.stepNodes.size();
  st.count = 1;
  e->imported[r] = 0;
  if (!poolAlloc(e, r))
    return e->fail(DCSMC_ENOMEM, "population pool (%zu bytes) too small for node %d and its children",
                   e->poolBytes, r);
  plan.stepNodes.push_back(r);
  e->poolPeak = std::max(e->poolPeak, e->poolTop);
  if (!e->opt.retainPopulations)
    for (int c = e->childOffsets[r]; c < e->childOffsets[r + 1]; ++c) {
      st.releaseAfter.push_back(e->children[c]);
      poolFree(e, e->children[c]);
    }
  plan.steps.push_back(std::move(st));
  return DCSMC_OK;
}

// ---- execution -------------------------------------------------------------------------------

struct ProfScope {
  dcsmc_engine* e;
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  ProfScope(dcsmc_engine* e_, int cls_, cudaStream_t s_) : e(e_), cls(cls_), s(s_) {
    e->info.kernelLaunches++;
    if (e->profile) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, s);
    }
  }
  ~ProfScope() {
    if (e->profile) {
      cudaEventRecord(b, s);
      e->profEvents.push_back({cls, e->curStep, a, b});
    }
  }
};

double uniformLeafLogW(const dcsmc_engine* e) {
  // DCRecursion.java:55,63: log(1.0) + proposal log weight; leaves propose with a constant log weight
  if (e->model == DCSMC_MODEL_MARKOV) return dc_log(1.0) + dc_log(e->markovPrior[0]);  // Doc.java:219
  return dc_log(1.0) + 0.0;  // MultiLevelLeafProposal.java:40
}

// Two-stream pipeline of a level batch. The propose kernels are fp64/issue bound and leave the
// memory system idle; the chain after them (scan, table, darts, expansion) is L2/latency bound and
// leaves the issue slots idle. A level is therefore cut into chunks of nodes: chunk k's chain runs
// on the high-priority stream while chunk k+1's propose kernel runs on the launching stream.
// Scratch (CDF, offspring counts, inverse tables, look-back descriptors) is indexed by the node's
// position in the level, so chunks use disjoint regions.
struct ChunkCtx {
  cudaStream_t sP, sR;   // propose stream, chain stream (== sP when not overlapping)
  int lo;                // first node of the chunk inside its level batch
  int idx;               // chunk number inside the level batch
};

cudaEvent_t nextEvent(dcsmc_engine* e) {
  if (e->evNext == e->evPool.size()) {
    cudaEvent_t ev = nullptr;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    e->evPool.push_back(ev);
  }
  return e->evPool[e->evNext++];
}

// look-back descriptor words per node of a step (scan: A and B limbs; expansion: one)
size_t descWordsPerNode(int N) {
  // an even count keeps every node's first descriptor pair 16-byte aligned (k_scan2 moves a pair at a time)
  static_assert(SCAN2_TILE == SCAN_TILE, "k_scan and k_scan2 share the descriptor budget");
  const size_t tilesS = (N + SCAN_TILE - 1) / SCAN_TILE, tilesE = (N + EXPAND_TILE - 1) / EXPAND_TILE;
  return (std::max(2 * tilesS, tilesE) + 1) / 2 * 2;
}
constexpr int MAX_CHUNKS = 16;

// Sample series a summary node writes: mean, variance (internal nodes), one delta per child when
// the node's children are still below the cut-off (DivideConquerMCAlgorithm.java:423-428).
bool summaryReportsDeltas(const dcsmc_engine* e, int n) {
  return e->depth[n] + 1 < e->opt.levelCutOffForOutput && nChildrenOf(e, n) > 0;
}
int summarySeries(const dcsmc_engine* e, int n) {
  return 1 + (nodeKind(e, n) == KIND_INTERNAL ? 1 : 0) + (summaryReportsDeltas(e, n) ? nChildrenOf(e, n) : 0) +
         (e->opt.retainSamples ? 1 : 0);
}
// uniforms of a node's summary stream (injected mode): 2 per attempt, maxAtt attempts, 2+C draws per particle
int64_t summaryUniforms(const dcsmc_engine* e, int n) {
  if (e->opt.levelCutOffForOutput <= 0 || e->model != DCSMC_MODEL_MULTILEVEL || e->depth.empty() ||
      e->depth[n] >= e->opt.levelCutOffForOutput)
    return 0;
  return (int64_t)e->opt.nParticles * (2 + nChildrenOf(e, n)) * e->opt.maxBetaAttempts * 2;
}

// Posterior summaries of the step's nodes above the output cut-off, right after their node step
// (their children are still resident). Batches of <= SUM_BATCH nodes, passed to the kernels by value.
template <int RNG>
int launchSummaries(dcsmc_engine* e, RunCtx& c, const Step& st, cudaStream_t s) {
  if (e->opt.levelCutOffForOutput <= 0 || e->model != DCSMC_MODEL_MULTILEVEL || e->dSumm == nullptr) return DCSMC_OK;
  const int* nodes = e->planStepNodes + st.offset;
  const int N = c.N;
  const int tilesP = (N + PROPOSE_THREADS - 1) / PROPOSE_THREADS;
  const size_t maxSeries = std::max<size_t>(64, ((size_t)1 << 28) / (size_t)c.Npad);  // <= 2 GiB of samples per batch
  int k = 0;
  while (k < st.count) {
    SumBatch b;
    memset(&b, 0, sizeof b);
    // retained samples: every summary node of the plan owns a fixed range of the store (sampleBase, set by
    // executePlan); otherwise the scratch is reused batch after batch
    const bool keepS = e->opt.retainSamples != 0;
    int series = 0, first = 0;
    for (; k < st.count && b.n < SUM_BATCH; ++k) {
      const int n = nodes[k];
      if (e->depth[n] >= e->opt.levelCutOffForOutput) continue;
      const int ns = summarySeries(e, n);
      if (!keepS && b.n > 0 && (size_t)(series + ns) > maxSeries) break;
      if (keepS && b.n == 0) first = (int)e->sampleBase[n];
      if (keepS && b.n > 0 && (int)e->sampleBase[n] != first + series) break;  // ranges of a batch must be contiguous
      b.node[b.n] = n;
      b.base[b.n] = first + series;
      b.delta[b.n] = summaryReportsDeltas(e, n) ? 1 : 0;
      series += ns;
      b.n++;
    }
    if (b.n == 0) continue;
    b.base[b.n] = first + series;
    const size_t need = keepS ? e->sampleSeriesTotal * (size_t)c.Npad : (size_t)series * c.Npad;
    if (need > e->sumScratchCap) {  // first (uncaptured) execution of a plan sizes the scratch
      if (e->dSumScratch) cudaFree(e->dSumScratch);
      e->dSumScratch = nullptr;
      e->sumScratchCap = 0;
      CK(cudaMalloc((void**)&e->dSumScratch, need * sizeof(double)));
      e->sumScratchCap = need;
      e->scratchGen++;
    }
    c.sumScratch = e->dSumScratch;
    {
      ProfScope p(e, DCSMC_K_GATHER, s);
      k_summary_samples<RNG><<<(unsigned)(b.n * tilesP), PROPOSE_THREADS, 0, s>>>(c, b, tilesP);
    }
    {
      ProfScope p(e, DCSMC_K_GATHER, s);
      k_summary_reduce<<<(unsigned)series, SUMRED_THREADS, 0, s>>>(c, b);
    }
    CK(cudaGetLastError());
  }
  return DCSMC_OK;
}

template <int RNG>
int launchStep(dcsmc_engine* e, const RunCtx& base, const Step& st, const ChunkCtx& ch) {
  RunCtx c = base;
  c.stepNodes = e->curStepNodes + st.offset;
  c.nStepNodes = st.count;
  const int N = c.N;
  const cudaStream_t sP = ch.sP, sR = ch.sR;
  // this chunk's scratch
  c.cdf = e->dCdf ? e->dCdf + (size_t)ch.lo * c.Npad : nullptr;
  c.invTable = e->dInvTable ? e->dInvTable + (size_t)ch.lo * (c.invB + 1) : nullptr;
  int32_t* const offspring = e->dOffspring ? e->dOffspring + (size_t)ch.lo * c.Npad : nullptr;
  unsigned long long* const desc = e->dDesc + (size_t)ch.lo * descWordsPerNode(N) + 2 * (size_t)ch.idx;
  const int tilesP = (N + PROPOSE_THREADS - 1) / PROPOSE_THREADS;
  const int tilesS = (N + SCAN_TILE - 1) / SCAN_TILE;
  const int tilesR = (dart_pairs(N) + RESAMPLE_THREADS - 1) / RESAMPLE_THREADS;  // two darts per thread
  const long long gridP = (long long)st.count * tilesP;
  if (gridP > 0x7fffffffLL) return e->fail(DCSMC_EINVAL, "step too large for one launch");
  // multilevel leaves: rESS == 1 exactly, so the resample decision is known on the host and the
  // leaf kernel does the dart work itself (k_finalize still records the statistics)
  const bool javaSums = e->opt.sumMode == DCSMC_SUM_JAVA;
  const bool leafFused = !javaSums && st.leaf && e->model == DCSMC_MODEL_MULTILEVEL && 1.0 < e->opt.relativeEssThreshold;
  // one block per node: pays when the level has enough nodes to fill the GPU (the few top levels of a tree keep
  // the two-kernel form, whose propose kernel spreads a node over many blocks)
  bool fusedSmall = !st.leaf && !javaSums && N <= SMALLN_MAX && e->model == DCSMC_MODEL_MULTILEVEL && e->fuseSmall && sR == sP &&
                    st.count >= e->fuseMinNodes;
  if (fusedSmall)
    for (int k = 0; k < st.count && fusedSmall; ++k) fusedSmall = nChildrenOf(e, e->planStepNodes[st.offset + k]) <= FUSED_CMAX;

  if (e->model == DCSMC_MODEL_MARKOV) {
    ProfScope p(e, DCSMC_K_MARKOV, sP);
    k_markov_propose<RNG><<<(unsigned)gridP, PROPOSE_THREADS, 0, sP>>>(c, tilesP);
  } else if (st.leaf) {
    // each warp refills its lanes from a chunk of CH particles (see k_leaf_propose). k_leaf_propose / 2 buffer the
    // chunk's candidates in shared memory (CH <= 512); k_leaf_propose3 does not: long chunks, the blocks of a node
    // carrying equal shares
    const int v = e->leafVariant;
    int CH, blocksPerNode;
    if (v == 3) {
      blocksPerNode = (N + LEAF3_WARPS * e->leafChunkMax - 1) / (LEAF3_WARPS * e->leafChunkMax);
      CH = (N + blocksPerNode * LEAF3_WARPS - 1) / (blocksPerNode * LEAF3_WARPS);
      CH = std::max(32, (CH + 31) / 32 * 32);
      blocksPerNode = (N + LEAF3_WARPS * CH - 1) / (LEAF3_WARPS * CH);
    } else {
      CH = (N + LEAF_WARPS - 1) / LEAF_WARPS;
      CH = std::min(512, std::max(32, (CH + 31) / 32 * 32));
      blocksPerNode = (N + LEAF_WARPS * CH - 1) / (LEAF_WARPS * CH);
    }
    if (leafFused && e->opt.resamplingScheme == DCSMC_MULTINOMIAL) {
      c.offspring = offspring;
      CK(cudaMemsetAsync(offspring, 0, (size_t)st.count * c.Npad * sizeof(int32_t), sP));
    }
    const int mode = !leafFused ? 0 : (e->opt.resamplingScheme == DCSMC_STRATIFIED ? 1 : 2);
    const bool v2 = v == 2;
    const size_t leafSmem = v == 3 ? 0 : LEAF_WARPS * (v2 ? leaf2_warp_smem_bytes(CH) : leaf_warp_smem_bytes(CH));
    size_t& smemSet = v2 ? e->leaf2SmemSet[RNG] : e->leafSmemSet[RNG];
    if (leafSmem > smemSet) {  // > 48 KB of dynamic shared memory needs the opt-in
      if (v2)
        CK(cudaFuncSetAttribute(k_leaf_propose2<RNG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leafSmem));
      else
        CK(cudaFuncSetAttribute(k_leaf_propose<RNG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leafSmem));
      smemSet = leafSmem;
    }
    ProfScope p(e, DCSMC_K_LEAF, sP);
    if (v == 3)
      k_leaf_propose3<RNG><<<(unsigned)(st.count * blocksPerNode), LEAF3_THREADS, 0, sP>>>(c, blocksPerNode, CH, mode);
    else if (v2)
      k_leaf_propose2<RNG><<<(unsigned)(st.count * blocksPerNode), PROPOSE_THREADS, leafSmem, sP>>>(c, blocksPerNode, CH, mode);
    else
      k_leaf_propose<RNG><<<(unsigned)(st.count * blocksPerNode), PROPOSE_THREADS, leafSmem, sP>>>(c, blocksPerNode, CH, mode);
  } else if (fusedSmall) {
    // the population fits in shared memory and the children's descriptors too: one block per node does the whole
    // node step (k_node_fused), the log weights never leave the SM
    const size_t smem = smalln_smem_bytes(c.Npad);
    const bool wideBlk = smalln_threads(c.Npad) == 1024;
    size_t& smemSet = wideBlk ? e->fusedSmemSet[RNG] : e->fusedSmemSet512[RNG];
    if (smem > smemSet) {
      if (wideBlk)
        CK(cudaFuncSetAttribute(k_node_fused<RNG, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else
        CK(cudaFuncSetAttribute(k_node_fused<RNG, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smemSet = smem;
    }
    {
      ProfScope p(e, DCSMC_K_INTERNAL, sP);
      if (wideBlk)
        k_node_fused<RNG, 1024><<<st.count, 1024, smem, sP>>>(c);
      else
        k_node_fused<RNG, 512><<<st.count, 512, smem, sP>>>(c);
    }
    CK(cudaGetLastError());
    return launchSummaries<RNG>(e, c, st, sP);
  } else {
    {
      ProfScope p(e, DCSMC_K_INTERNAL, sP);
      k_internal_prelude<<<(st.count + 3) / 4, 128, 0, sP>>>(c);
    }
    // can the linear-space product of the children's weights underflow anywhere in this step?
    int maxC = 0;
    for (int k = 0; k < st.count; ++k) maxC = std::max(maxC, nChildrenOf(e, e->planStepNodes[st.offset + k]));
    const bool wide = (double)maxC * std::log10((double)N) > 300.0;
    // the level above the leaves: every child of every node is a binomial leaf in its native layout, and leaves
    // always resample here (rESS == 1 < threshold), so nothing arrives weighted: k_internal_propose_lk
    bool leafKids = e->leafKidsKernel && !wide && 1.0 < e->opt.relativeEssThreshold;
    for (int k = 0; k < st.count && leafKids; ++k) {
      const int n = e->planStepNodes[st.offset + k];
      for (int cc = e->childOffsets[n]; cc < e->childOffsets[n + 1] && leafKids; ++cc) {
        const int ch = e->children[cc];
        leafKids = nodeKind(e, ch) == KIND_LEAF_BINOMIAL && !(!e->imported.empty() && e->imported[ch] == 1);
      }
    }
    ProfScope p(e, DCSMC_K_INTERNAL, sP);
    // 2-D grid (x = tile, y = node; the kernels are told by tilesPerNode == 0) whenever the step's node count fits
    // gridDim.y: spares every block the integer division of the flat index
    const bool grid2d = st.count <= 65535;
    const dim3 gridM = grid2d ? dim3((unsigned)tilesP, (unsigned)st.count) : dim3((unsigned)gridP);
    const int tpn = grid2d ? 0 : tilesP;
    if (leafKids && e->leafKidsKernel == 2)
      k_internal_propose_lk<RNG, true><<<gridM, PROPOSE_THREADS, 0, sP>>>(c, tpn);
    else if (leafKids)
      k_internal_propose_lk<RNG, false><<<gridM, PROPOSE_THREADS, 0, sP>>>(c, tpn);
    else if (wide)
      k_internal_propose<RNG, true><<<gridM, PROPOSE_THREADS, 0, sP>>>(c, tpn);
    else
      k_internal_propose<RNG, false><<<gridM, PROPOSE_THREADS, 0, sP>>>(c, tpn);
  }
  if (sR != sP) {  // the chain of this chunk starts when its propose kernel is done
    cudaEvent_t ev = nextEvent(e);
    CK(cudaEventRecord(ev, sP));
    CK(cudaStreamWaitEvent(sR, ev, 0));
  }
  const bool smallNode = !st.leaf && !javaSums && N <= SMALLN_MAX;
  if (smallNode) {
    // the population fits in shared memory: one block per node does everything after propose
    const size_t smem = smalln_smem_bytes(c.Npad);
    const bool wide = smalln_threads(c.Npad) == 1024;
    size_t& smemSet = wide ? e->smallSmemSet[RNG] : e->smallSmemSet512[RNG];
    if (smem > smemSet) {
      if (wide)
        CK(cudaFuncSetAttribute(k_node_small<RNG, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else
        CK(cudaFuncSetAttribute(k_node_small<RNG, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smemSet = smem;
    }
    {
      ProfScope p(e, DCSMC_K_SCAN, sR);
      if (wide)
        k_node_small<RNG, 1024><<<st.count, 1024, smem, sR>>>(c);
      else
        k_node_small<RNG, 512><<<st.count, 512, smem, sR>>>(c);
    }
    CK(cudaGetLastError());
    return launchSummaries<RNG>(e, c, st, sR);
  }
  if (st.leaf) {
    // all weights equal: the exact sums are known in closed form, no scan and no CDF
    ProfScope p(e, DCSMC_K_FINALIZE, sR);
    k_uniform_stats<<<(st.count + 127) / 128, 128, 0, sR>>>(c);
  }
  if (javaSums) {
    // parity mode: the reference's sequential sums (also finalises the node statistics)
    ProfScope p(e, DCSMC_K_SCAN, sR);
    k_scan_java<<<st.count, 32, 0, sR>>>(c);
  } else {
    if (!st.leaf) {
      const int tilesV = e->scanVariant == 1 ? tilesS : (N + SCAN2_TILE - 1) / SCAN2_TILE;
      const size_t needDesc = (size_t)st.count * tilesV;
      c.descA = desc;
      c.descB = desc + needDesc;  // k_scan only; k_scan2 keeps the (A, B) words of a tile next to each other
      CK(cudaMemsetAsync(desc, 0, 2 * needDesc * sizeof(unsigned long long), sR));
      ProfScope p(e, DCSMC_K_SCAN, sR);
      if (e->scanVariant == 1)
        k_scan<<<(unsigned)(st.count * tilesS), SCAN_THREADS, 0, sR>>>(c, tilesS);
      else
        k_scan2<<<(unsigned)(st.count * tilesV), SCAN_THREADS, 0, sR>>>(c, tilesV);
    }
    ProfScope p(e, DCSMC_K_FINALIZE, sR);
    k_finalize<<<(st.count + 127) / 128, 128, 0, sR>>>(c);
  }
  if (!st.leaf && !javaSums && e->opt.resamplingScheme == DCSMC_MULTINOMIAL && e->partitionDarts && e->dDartBuf != nullptr) {
    // darts grouped by CDF segment, then one block per segment works in shared memory (k_dart_partition /
    // k_seg_resample): no inverse index, no global offspring counters, no expansion kernel
    const int K = dp_segs(N), nCh = dp_chunks(N);
    c.dpSegs = K;
    c.dpChunks = nCh;
    c.dpStride = dp_stride(N);
    c.dartBuf = e->dDartBuf + (size_t)ch.lo * c.dpStride;
    c.chunkBin = e->dChunkBin + (size_t)ch.lo * nCh * (size_t)(K + 1);
    c.segTotal = e->dSegTotal + (size_t)ch.lo * K;
    CK(cudaMemsetAsync(c.segTotal, 0, (size_t)st.count * K * sizeof(int32_t), sR));
    const size_t smemA = sizeof(double) * ((size_t)K + DP_CHUNK) + sizeof(int) * ((size_t)K + 1);
    {
      ProfScope p(e, DCSMC_K_DARTS, sR);
      k_dart_partition<RNG><<<(unsigned)((size_t)st.count * nCh), DP_THREADS, smemA, sR>>>(c);
    }
    if (!e->segSmemSet) {
      CK(cudaFuncSetAttribute(k_seg_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_resample_smem_bytes()));
      e->segSmemSet = true;
    }
    {
      ProfScope p(e, DCSMC_K_RESAMPLE, sR);
      k_seg_resample<<<(unsigned)((size_t)st.count * K), SR_THREADS, seg_resample_smem_bytes(), sR>>>(c);
    }
    CK(cudaGetLastError());
    return launchSummaries<RNG>(e, c, st, sR);
  }
  if (!st.leaf || javaSums) {
    const int tilesT = (c.invB + 1 + RESAMPLE_THREADS - 1) / RESAMPLE_THREADS;
    ProfScope p(e, DCSMC_K_RESAMPLE, sR);
    k_build_inverse_table<<<(unsigned)(st.count * tilesT), RESAMPLE_THREADS, 0, sR>>>(c, tilesT);
  }
  if (e->opt.resamplingScheme == DCSMC_STRATIFIED) {
    if (!leafFused) {
      ProfScope p(e, DCSMC_K_RESAMPLE, sR);
      k_resample_stratified<RNG><<<(unsigned)(st.count * tilesR), RESAMPLE_THREADS, 0, sR>>>(c, tilesR);
    }
  } else {
    // multinomial: offspring counts by atomics, then scan + expansion (no dart sort)
    const int tilesE = (N + EXPAND_TILE - 1) / EXPAND_TILE;
    // segments per node: enough blocks to fill the GPU twice, as few look-back hops as possible
    const int wantBlocks = 2 * e->smCount * (2048 / EXPAND_THREADS);
    const int segs = std::max(1, std::min(tilesE, (wantBlocks + st.count - 1) / st.count));
    const int tilesPerSeg = (tilesE + segs - 1) / segs;
    const int segsPerNode = (tilesE + tilesPerSeg - 1) / tilesPerSeg;
    c.offspring = offspring;
    c.descA = desc;
    if (!leafFused) {
      CK(cudaMemsetAsync(offspring, 0, (size_t)st.count * c.Npad * sizeof(int32_t), sR));
      ProfScope p(e, DCSMC_K_DARTS, sR);
      k_darts_count<RNG><<<(unsigned)(st.count * tilesR), RESAMPLE_THREADS, 0, sR>>>(c, tilesR);
    }
    // the look-back descriptors of the scan are dead by now (stream order): reuse them
    if (segsPerNode > 1)
      CK(cudaMemsetAsync(desc, 0, (size_t)st.count * segsPerNode * sizeof(unsigned long long), sR));
    {
      ProfScope p(e, DCSMC_K_RESAMPLE, sR);
      k_expand<<<(unsigned)(st.count * segsPerNode), EXPAND_THREADS, 0, sR>>>(c, segsPerNode, tilesPerSeg);
    }
  }
  CK(cudaGetLastError());
  return launchSummaries<RNG>(e, c, st, sR);
}

int makeCtx(dcsmc_engine* e, RunCtx& c) {
  memset(&c, 0, sizeof c);
  const int N = e->opt.nParticles;
  c.nodes = e->dNodes;
  c.stats = e->dStats;
  c.leaf = e->dLeaf;
  c.leafObs = e->dLeafObs;
  c.children = e->dChildren;
  c.N = N;
  c.Npad = e->Npad;
  c.scheme = e->opt.resamplingScheme;
  c.maxAtt = e->opt.maxBetaAttempts;
  c.lkRange = e->lk2Range;
  c.l2Prefetch = (e->l2Prefetch & 1 && e->peers.empty() ? 1 : 0) | (e->l2Prefetch & 2);
  c.nStates = e->nStates;
  c.javaStepsInternal = e->model == DCSMC_MODEL_MARKOV ? 1 : 2;
  c.seed = (uint64_t)e->opt.masterRandomSeed;
  c.rk = philox_round_keys(c.seed);
  c.inj = e->dInj;
  c.javaState = e->dJavaState;
  c.markovT = e->dMarkovT;
  c.markovPrior = e->dMarkovPrior;
  c.rate = e->rate;
  c.logRate = dc_log(e->rate);
  c.invN = 1.0 / (double)N;
  c.logN = dc_log((double)N);
  c.dN = (double)N;
  c.threshold = e->opt.relativeEssThreshold;
  c.cdf = e->dCdf;
  c.invTable = e->dInvTable;
  c.invB = e->invB;
  c.javaSums = e->opt.sumMode == DCSMC_SUM_JAVA;
  c.skipObs = skipObservedLeaves(e);
  c.keepLogw = (e->opt.retainPopulations || !(1.0 < e->opt.relativeEssThreshold)) ? 1 : 0;
  c.uniformLogW = uniformLeafLogW(e);
  c.summ = e->dSumm;
  c.sumScratch = e->dSumScratch;
  c.useTransform = e->opt.useTransform;
  c.retainSamples = e->opt.retainSamples ? 1 : 0;
  return DCSMC_OK;
}

int validateForRun(dcsmc_engine* e) {
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree has not been called");
  if (e->model < 0) return e->fail(DCSMC_ESTATE, "no proposal model set (dcsmc_set_*_model)");
  if (e->opt.levelCutOffForOutput > 0 && e->model == DCSMC_MODEL_MULTILEVEL) {
    if (e->opt.rngMode == DCSMC_RNG_JAVA)
      return e->fail(DCSMC_EINVAL, "levelCutOffForOutput needs a slot-addressed stream (Philox or injected): "
                                   "nextGaussian consumes a data-dependent number of draws");
    if (e->opt.sumMode == DCSMC_SUM_JAVA)
      return e->fail(DCSMC_EINVAL, "levelCutOffForOutput is not available in the DCSMC_SUM_JAVA parity mode");
  }
  if (e->opt.rngMode == DCSMC_RNG_JAVA) {
    if (e->javaHash.empty()) return e->fail(DCSMC_ESTATE, "DCSMC_RNG_JAVA needs dcsmc_set_node_seeds");
    if (e->model == DCSMC_MODEL_MULTILEVEL)
      for (int n = 0; n < e->nNodes; ++n)
        if (nodeKind(e, n) == KIND_LEAF_BINOMIAL)
          return e->fail(DCSMC_EINVAL,
                         "DCSMC_RNG_JAVA cannot reproduce the variable-length Beta rejection stream of "
                         "binomial leaves (node %d)", n);
    if (e->model == DCSMC_MODEL_MARKOV && (e->nStates & (e->nStates - 1)) != 0)
      return e->fail(DCSMC_EINVAL, "DCSMC_RNG_JAVA needs a power-of-two nStates (Random.nextInt rejection loop)");
  }
  return DCSMC_OK;
}

// run a plan: upload step lists, execute, fetch stats
// an error left behind by an unchecked runtime call must not be blamed on the next checked one
#define CKPOINT(label)                                                                          \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return e->fail(DCSMC_ECUDA, "pending CUDA error at %s: %s", label, cudaGetErrorString(_e)); \
  } while (0)

// scatter the statistics of the last plan from the staging buffer into hStats / hOut (see executePlan)
void flushStats(const dcsmc_engine* ce) {
  dcsmc_engine* e = const_cast<dcsmc_engine*>(ce);
  if (e->pendNodes.empty()) return;
  if (e->inFlight) {  // deferred calls: the copies into hExport are ordered on the stream
    cudaSetDevice(e->opt.device);
    cudaStreamSynchronize(e->stream);
    e->inFlight = false;
  }
  for (size_t k = 0; k < e->pendNodes.size(); ++k) {
    const NodeStatsOut& h = e->hExport[k];
    const int n = e->pendNodes[k];
    dcsmc_node_stats_t& o = e->hStats[n];
    o.ess = h.ess;
    o.relEss = h.relEss;
    o.logNormEstimate = h.logZ;
    o.logScaling = h.logScaling;
    o.resampled = h.resampled == RES_WEIGHTED ? 0 : 1;
    o.done = h.done;
    e->hOut[n] = h;
  }
  e->statNodes.insert(e->statNodes.end(), e->pendNodes.begin(), e->pendNodes.end());
  e->pendNodes.clear();
}

double hostNowMs() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int executePlan(dcsmc_engine* e, Plan& plan, CachedPlan* cached = nullptr) {
  int rc;
  const int N = e->opt.nParticles;
  static const bool hostProf = getenv("DCSMC_HOSTPROF") != nullptr;  // stderr: host milliseconds per phase of this call
  const double tEntry = hostProf ? hostNowMs() : 0.0;
  double tLaunched = 0.0, tSynced = 0.0;
  CKPOINT("executePlan entry");
  // statistics of a previous plan of this session (dcsmc_run discards its predecessor's in
  // dcsmc_reset_populations); deferred calls append to the staging buffer instead and are flushed after dcsmc_sync
  if (!e->deferred) flushStats(e);
  // scratch sizing: CDF + look-back descriptors for internal steps, darts for multinomial
  size_t maxStep = 0, maxInternalStep = 0;
  for (auto& s : plan.steps) {
    maxStep = std::max<size_t>(maxStep, s.count);
    if (!s.leaf || e->opt.sumMode == DCSMC_SUM_JAVA) maxInternalStep = std::max<size_t>(maxInternalStep, s.count);
  }
  auto grow = [&](void** p, size_t& cap, size_t need, size_t elem) -> cudaError_t {
    if (need <= cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    cudaError_t r = cudaMalloc(p, need * elem);
    cap = r == cudaSuccess ? need : 0;
    e->scratchGen++;
    return r;
  };
  CK(grow((void**)&e->dCdf, e->cdfCap, std::max<size_t>(1, maxInternalStep * (size_t)e->Npad), sizeof(double)));
  CK(grow((void**)&e->dDesc, e->descCap, maxStep * descWordsPerNode(N) + 2 * MAX_CHUNKS + 2, sizeof(unsigned long long)));
  CKPOINT("scratch growth");
  e->invB = 1;
  while ((size_t)e->invB * 8 < (size_t)N) e->invB <<= 1;
  CK(grow((void**)&e->dInvTable, e->invTableCap, std::max<size_t>(1, maxInternalStep * ((size_t)e->invB + 1)), sizeof(int32_t)));
  if (e->opt.resamplingScheme == DCSMC_MULTINOMIAL)
    CK(grow((void**)&e->dOffspring, e->offspringCap, maxStep * (size_t)e->Npad, sizeof(int32_t)));
  if (e->opt.resamplingScheme == DCSMC_MULTINOMIAL && e->partitionDarts && N > SMALLN_MAX && e->opt.sumMode == DCSMC_SUM_EXACT &&
      maxInternalStep > 0) {
    CK(grow((void**)&e->dDartBuf, e->dartBufCap, maxInternalStep * (size_t)dp_stride(N), sizeof(double)));
    CK(grow((void**)&e->dChunkBin, e->chunkBinCap, maxInternalStep * (size_t)dp_chunks(N) * (size_t)(dp_segs(N) + 1), sizeof(int32_t)));
    CK(grow((void**)&e->dSegTotal, e->segTotalCap, maxInternalStep * (size_t)dp_segs(N), sizeof(int32_t)));
  }
  if (e->opt.retainSamples && e->opt.levelCutOffForOutput > 0 && e->model == DCSMC_MODEL_MULTILEVEL) {
    // every summary node of the plan owns a fixed range of the sample store, in launch order
    if ((int)e->sampleBase.size() != e->nNodes) e->sampleBase.assign(e->nNodes, -1);
    size_t total = 0;
    for (auto& s : plan.steps)
      for (int k = 0; k < s.count; ++k) {
        const int n = plan.stepNodes[s.offset + k];
        if (e->depth[n] >= e->opt.levelCutOffForOutput) continue;
        e->sampleBase[n] = (int64_t)total;
        total += (size_t)summarySeries(e, n);
      }
    e->sampleSeriesTotal = total;
    CK(grow((void**)&e->dSumScratch, e->sumScratchCap, std::max<size_t>(1, total * (size_t)e->Npad), sizeof(double)));
  }
  bool descriptorsOnDevice = cached != nullptr && cached->onDevice;
  if (plan.exportOrder.size() != plan.stepNodes.size()) {
    plan.exportOrder = plan.stepNodes;
    std::sort(plan.exportOrder.begin(), plan.exportOrder.end());
  }
  {
    // device list: [plan order | export order]
    int32_t* buf;
    if (cached) {
      if (cached->dStepNodes == nullptr) {
        CK(cudaMalloc((void**)&cached->dStepNodes, std::max<size_t>(1, 2 * plan.stepNodes.size()) * sizeof(int32_t)));
        descriptorsOnDevice = false;
      }
      buf = cached->dStepNodes;
    } else {
      if (2 * plan.stepNodes.size() > e->stepNodesTmpCap) {
        if (e->dStepNodesTmp) {
          CK(cudaStreamSynchronize(e->stream));  // deferred calls may still read the old list
          cudaFree(e->dStepNodesTmp);
        }
        e->dStepNodesTmp = nullptr;
        CK(cudaMalloc((void**)&e->dStepNodesTmp, 2 * plan.stepNodes.size() * sizeof(int32_t)));
        e->stepNodesTmpCap = 2 * plan.stepNodes.size();
      } else if (e->deferred && e->inFlight) {
        CK(cudaStreamSynchronize(e->stream));  // the one temporary list is about to be overwritten
      }
      buf = e->dStepNodesTmp;
    }
    e->curStepNodes = buf;
    if (!descriptorsOnDevice) {
      CK(cudaMemcpyAsync(buf, plan.stepNodes.data(), plan.stepNodes.size() * sizeof(int32_t),
                         cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(buf + plan.stepNodes.size(), plan.exportOrder.data(), plan.exportOrder.size() * sizeof(int32_t),
                         cudaMemcpyHostToDevice, e->stream));
    }
  }
  if (!descriptorsOnDevice) {
    // node descriptors for every node touched by the plan (a few nodes: one small copy each;
    // otherwise the whole table). Other cached plans that compute one of these nodes lose their claim.
    for (int n : plan.stepNodes) fillNodeDev(e, n);
    if (plan.stepNodes.size() * 16 < (size_t)e->nNodes) {
      for (int n : plan.stepNodes) {
        CK(cudaMemcpyAsync(e->dNodes + n, &e->hNodes[n], sizeof(NodeDev), cudaMemcpyHostToDevice, e->stream));
        descWritten(e, n, cached);
      }
    } else {
      CK(cudaMemcpyAsync(e->dNodes, e->hNodes.data(), sizeof(NodeDev) * e->nNodes, cudaMemcpyHostToDevice, e->stream));
      allDescWritten(e, cached);
    }
    if (cached != nullptr) cached->onDevice = true;
  }
  CKPOINT("descriptor upload");
  e->planStepNodes = plan.stepNodes.data();
  RunCtx c;
  makeCtx(e, c);
  const dcsmc_run_info before = e->info;  // deferred calls accumulate into one record (dcsmc_begin_deferred zeroes it)
  e->info = dcsmc_run_info{};
  e->info.poolBytes = (int64_t)e->poolBytes;
  e->info.h2dBytes = e->pendingH2D + (descriptorsOnDevice ? 0 : (int64_t)(plan.stepNodes.size() * (sizeof(int32_t) + sizeof(NodeDev))));
  e->pendingH2D = 0;
  cudaEvent_t evA = e->ev0, evB = e->ev1;
  if (e->deferred) {
    if (e->callEventsUsed == e->callEvents.size()) {
      cudaEvent_t a = nullptr, b = nullptr;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      e->callEvents.push_back({a, b});
    }
    evA = e->callEvents[e->callEventsUsed].first;
    evB = e->callEvents[e->callEventsUsed].second;
    e->callEventsUsed++;
  }
  // level batches, each cut into chunks whose chains overlap the next chunk's propose kernel
  const bool overlap = !e->profile && e->overlapChunks > 1 && e->streamR != nullptr;
  auto launchAll = [&]() -> int {
    // accumulators of the nodes about to be recomputed start from zero
    if (!plan.stepNodes.empty()) {
      const int n = (int)plan.stepNodes.size();
      k_reset_stats<<<(n + 255) / 256, 256, 0, e->stream>>>(e->dStats, e->dSumm, e->curStepNodes, n);
      CK(cudaGetLastError());
    }
    e->evNext = 0;
    cudaEvent_t chainDone = nullptr;  // last chain work of the previous level batch
    int stepIdx = 0;
    for (auto& s : plan.steps) {
      e->curStep = stepIdx++;
      int nCh = 1;
      if (overlap) nCh = std::max(1, std::min({e->overlapChunks, MAX_CHUNKS, s.count / 8}));
      if (chainDone) {  // children must be complete (and the scratch free) before this level starts
        CK(cudaStreamWaitEvent(e->stream, chainDone, 0));
        chainDone = nullptr;
      }
      for (int k = 0; k < nCh; ++k) {
        const int lo = (int)((long long)s.count * k / nCh), hi = (int)((long long)s.count * (k + 1) / nCh);
        Step sub;
        sub.offset = s.offset + lo;
        sub.count = hi - lo;
        sub.leaf = s.leaf;
        const ChunkCtx ch{e->stream, nCh > 1 ? e->streamR : e->stream, lo, k};
        int rc2;
        switch (e->opt.rngMode) {
          case DCSMC_RNG_PHILOX: rc2 = launchStep<RNG_PHILOX>(e, c, sub, ch); break;
          case DCSMC_RNG_INJECTED: rc2 = launchStep<RNG_INJECTED>(e, c, sub, ch); break;
          default: rc2 = launchStep<RNG_JAVA>(e, c, sub, ch); break;
        }
        if (rc2) return rc2;
      }
      if (nCh > 1) {
        chainDone = nextEvent(e);
        CK(cudaEventRecord(chainDone, e->streamR));
      }
      e->info.levelBatches++;
      e->info.nodesProcessed += s.count;
      for (int k = 0; k < s.count; ++k) {
        const int n = plan.stepNodes[s.offset + k];
        const int64_t C = nChildrenOf(e, n);
        e->info.algorithmicBytes +=
            (int64_t)N * (40 * C + 184 + (e->opt.resamplingScheme == DCSMC_MULTINOMIAL ? 16 : 0));
      }
    }
    if (chainDone) CK(cudaStreamWaitEvent(e->stream, chainDone, 0));
    return DCSMC_OK;
  };
  // A cached plan whose descriptors are already on the device is a fixed launch sequence: it is
  // captured once into a CUDA graph and replayed with one launch (no per-kernel launch gaps; this
  // matters for small populations and for the short per-GPU runs of a sharded tree).
  const bool useGraph = e->useGraphs && cached != nullptr && descriptorsOnDevice && !e->profile && !overlap;
  CK(cudaEventRecord(evA, e->stream));
  if (useGraph) {
    if (cached->graphExec != nullptr && cached->graphScratchGen != e->scratchGen) {
      cudaGraphExecDestroy(cached->graphExec);
      cached->graphExec = nullptr;
    }
    if (cached->graphExec == nullptr) {
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
      rc = launchAll();
      const cudaError_t endErr = cudaStreamEndCapture(e->stream, &g);
      if (rc) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      if (endErr != cudaSuccess) return e->fail(DCSMC_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(endErr));
      const cudaError_t instErr = cudaGraphInstantiate(&cached->graphExec, g, 0);
      cudaGraphDestroy(g);
      if (instErr != cudaSuccess) {
        cached->graphExec = nullptr;
        return e->fail(DCSMC_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(instErr));
      }
      cached->graphScratchGen = e->scratchGen;
      cached->graphInfo = e->info;
    } else {
      e->info.kernelLaunches = cached->graphInfo.kernelLaunches;
      e->info.levelBatches = cached->graphInfo.levelBatches;
      e->info.nodesProcessed = cached->graphInfo.nodesProcessed;
      e->info.algorithmicBytes = cached->graphInfo.algorithmicBytes;
    }
    CK(cudaGraphLaunch(cached->graphExec, e->stream));
  } else {
    if ((rc = launchAll())) return rc;
  }
  CK(cudaEventRecord(evB, e->stream));
  CKPOINT("level batches");
  if (hostProf) tLaunched = hostNowMs();
  // fetch the statistics of the nodes this plan computed: compact records, pinned staging
  const size_t nPlan = plan.stepNodes.size();
  const size_t off = e->pendNodes.size();  // 0 unless deferred calls are pending
  if (off + nPlan > e->exportCap) {
    const size_t want = std::max(off + nPlan, e->deferred ? (size_t)e->nNodes + 1024 : (size_t)0);
    if (e->inFlight) CK(cudaStreamSynchronize(e->stream));
    NodeStatsOut *nd = nullptr, *nh = nullptr;
    CK(cudaMalloc((void**)&nd, want * sizeof(NodeStatsOut)));
    CK(cudaMallocHost((void**)&nh, want * sizeof(NodeStatsOut)));
    if (off) memcpy(nh, e->hExport, off * sizeof(NodeStatsOut));
    if (e->dExport) cudaFree(e->dExport);
    if (e->hExport) cudaFreeHost(e->hExport);
    e->dExport = nd;
    e->hExport = nh;
    e->exportCap = want;
  }
  if (nPlan > 0) {
    k_export_stats<<<(unsigned)((nPlan + 255) / 256), 256, 0, e->stream>>>(e->dStats, e->curStepNodes + nPlan, (int)nPlan, e->dExport + off,
                                                                               e->dNodes, e->pool, (int32_t)e->poolGeneration);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(e->hExport + off, e->dExport + off, nPlan * sizeof(NodeStatsOut), cudaMemcpyDeviceToHost, e->stream));
  }
  e->info.d2hBytes = (int64_t)(nPlan * sizeof(NodeStatsOut));
  e->info.particleUpdates = e->info.nodesProcessed * (int64_t)N;
  if (e->deferred) {
    // no host synchronisation: the caller keeps enqueueing (the next level, a collective, the closing barrier)
    e->inFlight = true;
    e->pendNodes.insert(e->pendNodes.end(), plan.exportOrder.begin(), plan.exportOrder.end());
    e->info.kernelLaunches += before.kernelLaunches;
    e->info.particleUpdates += before.particleUpdates;
    e->info.nodesProcessed += before.nodesProcessed;
    e->info.algorithmicBytes += before.algorithmicBytes;
    e->info.h2dBytes += before.h2dBytes;
    e->info.d2hBytes += before.d2hBytes;
    e->info.levelBatches += before.levelBatches;
    e->curStep = -1;
    return DCSMC_OK;
  }
  CK(cudaStreamSynchronize(e->stream));
  if (hostProf) tSynced = hostNowMs();
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->info.deviceMs = ms;
  // the records stay in the pinned staging buffer; they are scattered into the per-node host tables when somebody
  // asks (flushStats) — a loop of back-to-back runs on a 131 071-node tree spent 2 ms per run in that scatter
  e->pendNodes.insert(e->pendNodes.end(), plan.exportOrder.begin(), plan.exportOrder.end());
  if (e->profile) {
    // per level batch: the device time of its launches, shared equally by the sibling nodes that ran in it
    // (what DefaultProcessorFactory reports as iterationProposalTime, DefaultProcessorFactory.java:41-50)
    std::vector<double> stepMs(plan.steps.size(), 0.0);
    for (auto& pe : e->profEvents) {
      float t = 0;
      cudaEventElapsedTime(&t, pe.a, pe.b);
      e->profMs[pe.cls] += t;
      e->profLaunches[pe.cls]++;
      if (pe.step >= 0 && pe.step < (int)stepMs.size()) stepMs[pe.step] += t;
      cudaEventDestroy(pe.a);
      cudaEventDestroy(pe.b);
    }
    e->profEvents.clear();
    if ((int)e->hNodeMs.size() != e->nNodes) e->hNodeMs.assign(e->nNodes, 0.0);
    for (size_t k = 0; k < plan.steps.size(); ++k)
      for (int j = 0; j < plan.steps[k].count; ++j)
        e->hNodeMs[plan.stepNodes[plan.steps[k].offset + j]] = stepMs[k] / std::max(1, plan.steps[k].count);
  }
  e->curStep = -1;
  if (hostProf)
    fprintf(stderr, "[dcsmc hostprof] plan %zu nodes: launch %.3f ms, wait %.3f ms, stats %.3f ms (device %.3f ms)\n",
            plan.stepNodes.size(), tLaunched - tEntry, tSynced - tLaunched, hostNowMs() - tSynced, e->info.deviceMs);
  return DCSMC_OK;
}

int prepare(dcsmc_engine* e) {
  int rc;
  if ((rc = ensureDevice(e))) return rc;
  CKPOINT("prepare entry");
  if ((rc = validateForRun(e))) return rc;
  e->Npad = (int)alignUp((size_t)e->opt.nParticles, 32);
  if ((rc = uploadTables(e))) return rc;
  CKPOINT("table upload");
  if ((rc = uploadInjected(e))) return rc;
  CKPOINT("injected uniforms");
  if ((rc = ensurePool(e))) return rc;
  CKPOINT("population pool");
  return DCSMC_OK;
}

// residency of the roots' children: with the allocator state it decides what the planner does
uint64_t kidsSignature(const dcsmc_engine* e, const std::vector<int>& roots) {
  uint64_t h = 1469598103934665603ULL;
  for (int r : roots)
    for (int c = e->childOffsets[r]; c < e->childOffsets[r + 1]; ++c) {
      h ^= (uint64_t)(uint32_t)e->children[c] * 2 + (e->resident[e->children[c]] ? 1 : 0);
      h *= 1099511628211ULL;
    }
  return h;
}

int runRoots(dcsmc_engine* e, const std::vector<int>& roots) {
  int rc;
  if ((rc = prepare(e))) return rc;
  bool emptyPool = e->poolTop == 0;
  if (emptyPool)
    for (char r : e->resident)
      if (r) {
        emptyPool = false;
        break;
      }
  const uint64_t kids = emptyPool ? 0 : kidsSignature(e, roots);
  for (auto& up : e->planCache) {
    CachedPlan& cp = *up;
    if (cp.roots != roots || cp.contextual == emptyPool) continue;
    if (emptyPool) {
      e->freeBySize = cp.freeBySize;
      e->poolTop = cp.poolTop;
      if (cp.onDevice && cp.finalResident.size() * 8 < (size_t)e->nNodes) {
        // the device already holds this plan's descriptors and the pool is empty: only the nodes that
        // stay resident need their host-side slot back (O(roots), not O(nNodes))
        for (int n : cp.finalResident) {
          e->slot[n] = cp.slot[n];
          e->resident[n] = 1;
        }
      } else {
        e->slot = cp.slot;
        e->resident = cp.resident;
        e->imported = cp.imported;
      }
      if ((rc = uploadLeafForPlan(e, cp.plan, &cp))) return rc;
      return executePlan(e, cp.plan, &cp);
    }
    // contextual plan: valid when the allocator and the children are in the state it was made for (the schedule
    // of a sharded run is static, so from the second run on this always holds)
    if (cp.entryTop != e->poolTop || cp.entryKids != kids || cp.entryFree != e->freeBySize) continue;
    for (const NodeDelta& d : cp.delta) {
      const Slot& cur = e->slot[d.node];
      if (cur.external && !cur.peer && cur.external != d.slot.external) e->importFree.push_back(cur.external);  // as poolFree would
      e->slot[d.node] = d.slot;
      e->resident[d.node] = d.resident;
      e->imported[d.node] = d.imported;
    }
    e->freeBySize = cp.freeBySize;
    e->poolTop = cp.poolTop;
    if ((rc = uploadLeafForPlan(e, cp.plan, &cp))) return rc;
    return executePlan(e, cp.plan, &cp);
  }
  Plan plan;
  const bool record = e->planCache.size() < MAX_CACHED_PLANS;
  std::unique_ptr<CachedPlan> cp;
  if (record) {
    cp.reset(new CachedPlan());
    cp->roots = roots;
    cp->contextual = !emptyPool;
    if (!emptyPool) {
      cp->entryFree = e->freeBySize;
      cp->entryTop = e->poolTop;
      cp->entryKids = kids;
      e->undo.clear();
      e->undoOn = true;  // the log names every node the planner touches
    }
  }
  rc = planRoots(e, roots, plan);
  std::vector<int> touched;
  if (record && !emptyPool) {
    for (const dcsmc_engine::Undo& u : e->undo) touched.push_back(u.node);
    e->undo.clear();
    e->undoOn = false;
  }
  if (rc) return rc;
  if (!record) {
    if ((rc = uploadLeafForPlan(e, plan, nullptr))) return rc;
    return executePlan(e, plan);
  }
  cp->plan = plan;
  cp->freeBySize = e->freeBySize;
  cp->poolTop = e->poolTop;
  cp->inPlan.assign(e->nNodes, 0);
  for (int n : plan.stepNodes) {
    cp->inPlan[n] = 1;
    if (e->resident[n]) cp->finalResident.push_back(n);
  }
  if (emptyPool) {
    cp->slot = e->slot;
    cp->resident = e->resident;
    cp->imported = e->imported;
  } else {
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    for (int n : touched) cp->delta.push_back({n, e->slot[n], e->resident[n], e->imported[n]});
  }
  e->planCache.push_back(std::move(cp));
  CachedPlan* held = e->planCache.back().get();
  if ((rc = uploadLeafForPlan(e, held->plan, held))) return rc;
  return executePlan(e, held->plan, held);
}

// Wire format of one exported population (DCSMC_EXPORT_DESC_BYTES host bytes).
struct ExportDesc {
  int32_t magic, N;
  int64_t offset;       // of the population inside the owner's pool
  int64_t generation;   // of that pool
  double m, S, ess, relEss, logScaling, logZ;
  int32_t resampled, done;
  int64_t slotBytes;    // bytes of the population: offset + slotBytes must lie inside the owner's pool
  int32_t model;        // DCSMC_MODEL_* — fixes the column layout together with `kind`
  int32_t kind;         // KIND_* of the node on the owner (internal: 6 columns + logw; leaf: 2 columns)
  int32_t node;         // node id on the owner (must equal the importing id: one tree, one numbering)
  int32_t skipObs;      // the owner does not materialise observed leaves
  char pad[DCSMC_EXPORT_DESC_BYTES - 104];
};
static_assert(sizeof(ExportDesc) == DCSMC_EXPORT_DESC_BYTES, "export descriptor size");

// validation shared by the two import entry points (every descriptor, before any state changes)
int validateImport(dcsmc_engine* e, const int32_t* nodes, const int32_t* peerIds, int32_t n, const ExportDesc* in,
                   bool layoutOnly) {
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    const ExportDesc& d = in[k];
    if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
    if (d.magic != PACK_MAGIC || d.N != e->opt.nParticles)
      return e->fail(DCSMC_EINVAL, "exported population of node %d does not match this engine (magic/N)", node);
    if (!(d.done == 1 || (layoutOnly && d.done == 2)))
      return e->fail(DCSMC_EINVAL, "exported population of node %d is not finished (done = %d)", node, d.done);
    if (d.node != node || d.model != e->model || d.kind != nodeKind(e, node) || d.skipObs != (skipObservedLeaves(e) ? 1 : 0))
      return e->fail(DCSMC_EINVAL, "exported population of node %d has another layout (node %d, model %d, kind %d, skipObs %d)",
                     node, d.node, d.model, d.kind, d.skipObs);
    if (!layoutOnly && d.resampled != RES_WEIGHTED && d.resampled != RES_ANCESTORS && d.resampled != RES_MATERIALISED)
      return e->fail(DCSMC_EINVAL, "exported population of node %d: bad resampling state %d", node, d.resampled);
    auto it = e->peers.find(peerIds[k]);
    if (it == e->peers.end() || !it->second.base || it->second.generation != d.generation)
      return e->fail(DCSMC_ESTATE, "pool of peer %d (generation %lld) is not mapped: call dcsmc_peer_open", peerIds[k],
                     (long long)d.generation);
    if (d.slotBytes != (int64_t)slotBytes(e, node) || d.offset < 0 || d.offset % 256 != 0 ||
        d.offset + d.slotBytes > it->second.bytes)
      return e->fail(DCSMC_EINVAL, "exported population of node %d lies outside the pool of peer %d (offset %lld, %lld bytes, pool %lld)",
                     node, peerIds[k], (long long)d.offset, (long long)d.slotBytes, (long long)it->second.bytes);
  }
  return DCSMC_OK;
}

}  // namespace

// =================================== C ABI ====================================================

extern "C" {

int32_t dcsmc_abi_version(void) { return DCSMC_ABI_VERSION; }

int32_t dcsmc_options_default(dcsmc_options* o) {
  if (!o) return DCSMC_EINVAL;
  memset(o, 0, sizeof *o);
  o->masterRandomSeed = 31;              // DCOptions.java:16
  o->nParticles = 1000;                  // :19
  o->resamplingScheme = DCSMC_MULTINOMIAL;  // :22
  o->relativeEssThreshold = 1.0 + 1e-10; // :25  1.0 + NumericalUtils.THRESHOLD (always resample)
  o->clusterSubGroup = 1;
  o->nThreadsPerNode = 1;
  o->minimumNumberOfClusterMembersToStart = 1;
  o->maximumTimeToWaitInMinutes = 1;
  o->maximumDistributionDepth = 0x7fffffff;  // Integer.MAX_VALUE, :41
  o->indexInCluster = 1;
  o->device = 0;
  o->rngMode = DCSMC_RNG_PHILOX;
  o->sumMode = DCSMC_SUM_EXACT;
  o->maxBetaAttempts = 32;
  o->retainPopulations = 0;
  o->levelCutOffForOutput = 0;  // reference default 2 (DivideConquerMCAlgorithm.java:56); opt-in here
  o->useTransform = 1;          // MultiLevelModelOptions.useTransform (:47)
  o->memoryBudgetBytes = 0;
  o->retainSamples = 0;
  return DCSMC_OK;
}

const char* dcsmc_last_error(const dcsmc_engine* e) { return e ? e->err.c_str() : g_createError.c_str(); }

int32_t dcsmc_create(const dcsmc_options* opt, dcsmc_engine** out) {
  if (!opt || !out) {
    g_createError = "dcsmc_create: null argument";
    return DCSMC_EINVAL;
  }
  *out = nullptr;
  if (opt->nParticles < 1 || opt->nParticles >= (1 << 24)) {
    g_createError = "nParticles must be in [1, 2^24)";
    return DCSMC_EINVAL;
  }
  if (opt->maxBetaAttempts < 1) {
    g_createError = "maxBetaAttempts must be >= 1";
    return DCSMC_EINVAL;
  }
  if (opt->sumMode != DCSMC_SUM_EXACT && opt->sumMode != DCSMC_SUM_JAVA) {
    g_createError = "sumMode must be DCSMC_SUM_EXACT or DCSMC_SUM_JAVA";
    return DCSMC_EINVAL;
  }
  int nDev = 0;
  cudaError_t ce = cudaGetDeviceCount(&nDev);
  if (ce != cudaSuccess || nDev <= 0 || opt->device >= nDev) {
    g_createError = std::string("no usable CUDA device (there is no CPU fallback): ") +
                    (ce != cudaSuccess ? cudaGetErrorString(ce) : "device ordinal out of range");
    return DCSMC_ECUDA;
  }
  dcsmc_engine* e = new dcsmc_engine();
  e->opt = *opt;
  {
    std::random_device rd;
    e->poolGeneration = (int64_t)(((uint64_t)rd() & 0x7fffffffULL) << 24);
  }
  int prLow = 0, prHigh = 0;
  if (cudaSetDevice(opt->device) == cudaSuccess) cudaDeviceGetStreamPriorityRange(&prLow, &prHigh);
  cudaDeviceGetAttribute(&e->smCount, cudaDevAttrMultiProcessorCount, opt->device);
  if (e->smCount <= 0) e->smCount = 148;
  if (const char* v = getenv("DCSMC_GRAPHS")) e->useGraphs = atoi(v) != 0;
  if (const char* v = getenv("DCSMC_FUSE_SMALL")) e->fuseSmall = atoi(v) != 0;
  if (const char* v = getenv("DCSMC_PARTITION_DARTS")) e->partitionDarts = atoi(v) != 0;
  if (const char* v = getenv("DCSMC_LEAF_V")) e->leafVariant = std::min(3, std::max(1, atoi(v)));
  if (const char* v = getenv("DCSMC_LEAF_CH")) e->leafChunkMax = std::min(65536, std::max(32, atoi(v) / 32 * 32));
  if (const char* v = getenv("DCSMC_LEAFKIDS")) e->leafKidsKernel = std::min(2, std::max(0, atoi(v)));
  if (const char* v = getenv("DCSMC_L2_PREFETCH")) e->l2Prefetch = atoi(v) & 3;
  if (const char* v = getenv("DCSMC_LK2_RANGE")) e->lk2Range = std::min(100, std::max(0, atoi(v)));
  if (const char* v = getenv("DCSMC_SCAN_V")) e->scanVariant = atoi(v) == 1 ? 1 : 2;
  e->fuseMinNodes = e->smCount;
  if (const char* v = getenv("DCSMC_FUSE_MIN_NODES")) e->fuseMinNodes = std::max(1, atoi(v));
  if (const char* v = getenv("DCSMC_OVERLAP_CHUNKS")) e->overlapChunks = std::max(1, atoi(v));
  if (cudaSetDevice(opt->device) != cudaSuccess ||
      cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prLow) != cudaSuccess ||
      cudaStreamCreateWithPriority(&e->streamR, cudaStreamNonBlocking, prHigh) != cudaSuccess ||
      cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess) {
    g_createError = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
    delete e;
    return DCSMC_ECUDA;
  }
  *out = e;
  return DCSMC_OK;
}

int32_t dcsmc_destroy(dcsmc_engine* e) {
  if (!e) return DCSMC_OK;
  const bool dbg = getenv("DCSMC_DEBUG") != nullptr;
  auto chk = [&](const char* what) {  // never leave an error behind for the next engine of this thread
    const cudaError_t r = cudaGetLastError();
    if (r != cudaSuccess && dbg) fprintf(stderr, "dcsmc_destroy: %s: %s\n", what, cudaGetErrorString(r));
  };
  cudaSetDevice(e->opt.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  chk("entry");
  for (double* p : e->injDev)
    if (p) cudaFree(p);
  chk("injected uniforms");
  for (Slot& sl : e->slot)
    if (sl.external && !sl.peer) cudaFree(sl.external);
  for (char* p : e->importFree) cudaFree(p);
  chk("imported populations");
  for (auto& kv : e->peers)
    if (kv.second.base) cudaIpcCloseMemHandle(kv.second.base);
  chk("peer mappings");
  if (e->hImp) cudaFreeHost(e->hImp);
  if (e->dImp) cudaFree(e->dImp);
  if (e->hErr) cudaFreeHost(e->hErr);
  if (e->dErr) cudaFree(e->dErr);
  if (e->hLeafStage) cudaFreeHost(e->hLeafStage);
  chk("import staging");
  if (e->hExport) cudaFreeHost(e->hExport);
  chk("export staging (host)");
  if (e->dExport) cudaFree(e->dExport);
  chk("export staging (device)");
  if (e->evImp) cudaEventDestroy(e->evImp);
  if (e->evT0) cudaEventDestroy(e->evT0);
  if (e->evT1) cudaEventDestroy(e->evT1);
  chk("timer events");
  for (cudaEvent_t ev : e->stageEvents) cudaEventDestroy(ev);
  if (e->hStage) cudaFreeHost(e->hStage);
  chk("population staging (host)");
  if (e->dStage) cudaFree(e->dStage);
  chk("population staging (device)");
  if (e->dSumm) cudaFree(e->dSumm);
  if (e->dSumScratch) cudaFree(e->dSumScratch);
  void* ptrs[] = {e->dNodes, e->dStats, e->dLeaf, e->dLeafObs, e->dChildren, e->dStepNodesTmp, e->dInj, e->dJavaState,
                  e->dMarkovT, e->dMarkovPrior, e->pool, e->dCdf, e->dDesc, e->dOffspring, e->dInvTable,
                  e->dDartBuf, e->dChunkBin, e->dSegTotal};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  chk("tables");
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->streamR) cudaStreamDestroy(e->streamR);
  for (cudaEvent_t ev : e->evPool) cudaEventDestroy(ev);
  clearPlanCache(e);
  for (auto& pr : e->callEvents) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  chk("streams, events, graphs");
  delete e;
  return DCSMC_OK;
}

int32_t dcsmc_set_tree(dcsmc_engine* e, int32_t nNodes, int32_t root, const int32_t* childOffsets,
                       const int32_t* children) {
  if (!e) return DCSMC_EINVAL;
  if (nNodes < 1 || !childOffsets || root < 0 || root >= nNodes)
    return e->fail(DCSMC_EINVAL, "dcsmc_set_tree: bad arguments");
  // The same topology again (a caller that re-sends its tree with every data set): device tables, node seeds,
  // memory plans and captured graphs all stay valid — only the model data that follows is new.
  if (e->nNodes == nNodes && e->root == root && (int)e->childOffsets.size() == nNodes + 1 &&
      memcmp(e->childOffsets.data(), childOffsets, sizeof(int32_t) * (size_t)(nNodes + 1)) == 0 &&
      (nNodes == 1 || (children && memcmp(e->children.data(), children, sizeof(int32_t) * (size_t)(nNodes - 1)) == 0)))
    return DCSMC_OK;
  if (childOffsets[0] != 0) return e->fail(DCSMC_EINVAL, "childOffsets[0] must be 0");
  const int nEdges = childOffsets[nNodes];
  if (nEdges != nNodes - 1) return e->fail(DCSMC_EINVAL, "a tree with %d nodes needs %d edges, got %d", nNodes, nNodes - 1, nEdges);
  if (nEdges > 0 && !children) return e->fail(DCSMC_EINVAL, "children is null");
  std::vector<int32_t> parent(nNodes, -1);
  for (int n = 0; n < nNodes; ++n) {
    if (childOffsets[n + 1] < childOffsets[n]) return e->fail(DCSMC_EINVAL, "childOffsets not monotone");
    for (int c = childOffsets[n]; c < childOffsets[n + 1]; ++c) {
      const int ch = children[c];
      if (ch < 0 || ch >= nNodes || ch == root || parent[ch] != -1)
        return e->fail(DCSMC_EINVAL, "node %d: invalid or duplicate child %d", n, ch);
      parent[ch] = n;
    }
  }
  // heights (leaf = 0) via reverse BFS order
  std::vector<int> order;
  order.reserve(nNodes);
  order.push_back(root);
  for (size_t k = 0; k < order.size(); ++k)
    for (int c = childOffsets[order[k]]; c < childOffsets[order[k] + 1]; ++c) order.push_back(children[c]);
  if ((int)order.size() != nNodes) return e->fail(DCSMC_EINVAL, "tree is not connected");
  std::vector<int32_t> height(nNodes, 0);
  for (int k = nNodes - 1; k > 0; --k) {
    const int n = order[k];
    height[parent[n]] = std::max(height[parent[n]], height[n] + 1);
  }
  // per-node host state of the previous tree is meaningless now (and possibly shorter than the new tree)
  for (Slot& sl : e->slot)
    if (sl.external && !sl.peer) e->importFree.push_back(sl.external);
  e->slot.clear();
  e->resident.clear();
  e->imported.clear();
  e->hStats.clear();
  e->hOut.clear();
  e->pendNodes.clear();
  e->statNodes.clear();
  e->externalNodes.clear();
  e->freeBySize.clear();
  e->poolTop = 0;
  e->nNodes = nNodes;
  e->root = root;
  e->childOffsets.assign(childOffsets, childOffsets + nNodes + 1);
  e->children.assign(children, children + nEdges);
  e->parent = parent;
  e->height = height;
  // subtree sizes; pre-order numbering (root 0, a subtree = the id range [r, r + size)) lets a rank upload and plan
  // only its own part of a large tree
  e->subtreeSize.assign(nNodes, 1);
  for (int k = nNodes - 1; k > 0; --k) e->subtreeSize[parent[order[k]]] += e->subtreeSize[order[k]];
  e->preorder = root == 0;
  for (int n = 0; n < nNodes && e->preorder; ++n) {
    int next = n + 1;
    for (int c = childOffsets[n]; c < childOffsets[n + 1]; ++c) {
      if (children[c] != next) { e->preorder = false; break; }
      next += e->subtreeSize[children[c]];
    }
  }
  e->depth.assign(nNodes, 0);  // Node.level() (prototype/Node.java:72-75): root = 0
  for (int k = 1; k < nNodes; ++k) e->depth[order[k]] = e->depth[parent[order[k]]] + 1;
  e->javaHash.clear();
  e->injHost.clear();
  for (double* p : e->injDev)
    if (p) cudaFree(p);
  e->injDev.clear();
  e->tablesDirty = true;
  return DCSMC_OK;
}

int32_t dcsmc_set_node_seeds(dcsmc_engine* e, const int32_t* h) {
  if (!e) return DCSMC_EINVAL;
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree first");
  if (!h) return e->fail(DCSMC_EINVAL, "null hash codes");
  if ((int)e->javaHash.size() == e->nNodes && memcmp(e->javaHash.data(), h, sizeof(int32_t) * (size_t)e->nNodes) == 0)
    return DCSMC_OK;
  e->javaHash.assign(h, h + e->nNodes);
  e->tablesDirty = true;
  return DCSMC_OK;
}

int32_t dcsmc_set_multilevel_model(dcsmc_engine* e, const int32_t* leafKind, const int32_t* nTrials,
                                   const int32_t* nSuccesses, const double* leafObs, double variancePrior) {
  if (!e) return DCSMC_EINVAL;
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree first");
  if (!(variancePrior > 0)) return e->fail(DCSMC_EINVAL, "variancePrior must be > 0");
  const int nN = e->nNodes;
  const bool sameFamily = e->model == DCSMC_MODEL_MULTILEVEL && e->rate == variancePrior && !e->tablesDirty;
  // validate before any state changes (DivideConquerMCAlgorithm.logBinomialPr :595-596 throws on bad counts)
  for (int n = 0; n < nN; ++n) {
    if (e->childOffsets[n + 1] != e->childOffsets[n]) continue;
    const int32_t kind = leafKind ? leafKind[n] : DCSMC_LEAF_BINOMIAL;
    if (kind == DCSMC_LEAF_BINOMIAL) {
      if (!nTrials || !nSuccesses) return e->fail(DCSMC_EINVAL, "binomial leaves need nTrials/nSuccesses");
      if (nTrials[n] < 0 || nSuccesses[n] < 0 || nSuccesses[n] > nTrials[n])
        return e->fail(DCSMC_EINVAL, "leaf %d: invalid (nTrials=%d, nSuccesses=%d)", n, nTrials[n], nSuccesses[n]);
    } else if (kind != DCSMC_LEAF_GAUSS_OBS) {
      return e->fail(DCSMC_EINVAL, "leaf %d: unknown leaf kind %d", n, kind);
    }
  }
  // the memory plan depends on the leaf KINDS only (observed leaves hold no population); new counts /
  // observations on the same kinds are new input data for the same plan
  bool sameKinds = (int)e->leafKind.size() == nN;
  if (sameKinds) {
    if (leafKind)
      sameKinds = memcmp(e->leafKind.data(), leafKind, sizeof(int32_t) * (size_t)nN) == 0;
    else
      for (int n = 0; n < nN && sameKinds; ++n) sameKinds = e->leafKind[n] == DCSMC_LEAF_BINOMIAL;
  }
  if (!sameKinds) {
    if (leafKind)
      e->leafKind.assign(leafKind, leafKind + nN);
    else
      e->leafKind.assign(nN, DCSMC_LEAF_BINOMIAL);
  }
  e->nTrials.resize(nN);
  e->nSuccesses.resize(nN);
  e->binomPrefix.resize((size_t)nN + 1);
  e->binomPrefix[0] = 0;
  for (int n = 0; n < nN; ++n) {  // counts of internal nodes and observed leaves are kept at 0 whatever the caller sent
    const bool binom = e->childOffsets[n + 1] == e->childOffsets[n] && e->leafKind[n] == DCSMC_LEAF_BINOMIAL;
    e->nTrials[n] = binom ? nTrials[n] : 0;
    e->nSuccesses[n] = binom ? nSuccesses[n] : 0;
    e->binomPrefix[n + 1] = e->binomPrefix[n] + (binom ? 1 : 0);
  }
  if (leafObs)
    e->leafObs.assign(leafObs, leafObs + nN);
  else
    e->leafObs.assign(nN, 0.0);
  e->model = DCSMC_MODEL_MULTILEVEL;
  e->rate = variancePrior;
  if (sameFamily && sameKinds)
    e->leafVersion++;
  else
    e->tablesDirty = true;
  return DCSMC_OK;
}

int32_t dcsmc_set_markov_model(dcsmc_engine* e, int32_t nStates, const double* transition, const double* prior) {
  if (!e) return DCSMC_EINVAL;
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree first");
  // TestUtilities.checkValidMarkovChain (TestUtilities.java:76-82)
  if (nStates < 1 || !transition || !prior) return e->fail(DCSMC_EINVAL, "bad Markov model");
  e->model = DCSMC_MODEL_MARKOV;
  e->nStates = nStates;
  e->markovT.assign(transition, transition + (size_t)nStates * nStates);
  e->markovPrior.assign(prior, prior + nStates);
  e->tablesDirty = true;
  return DCSMC_OK;
}

int32_t dcsmc_slots_per_particle(const dcsmc_engine* e, int32_t node, int32_t* slots) {
  if (!e || !slots) return DCSMC_EINVAL;
  if (node < 0 || node >= e->nNodes || e->model < 0) return e->fail(DCSMC_EINVAL, "bad node or no model");
  *slots = slotsPerParticle(e, node);
  return DCSMC_OK;
}

int32_t dcsmc_inject_uniforms(dcsmc_engine* e, int32_t node, const double* u, int64_t n) {
  if (!e) return DCSMC_EINVAL;
  if (node < 0 || node >= e->nNodes || e->model < 0) return e->fail(DCSMC_EINVAL, "bad node or no model");
  if (e->opt.rngMode != DCSMC_RNG_INJECTED) return e->fail(DCSMC_ESTATE, "engine was not created with DCSMC_RNG_INJECTED");
  const int64_t need = (int64_t)e->opt.nParticles * slotsPerParticle(e, node) + e->opt.nParticles + summaryUniforms(e, node);
  if (!u || n < need) return e->fail(DCSMC_EINVAL, "node %d needs %lld uniforms, got %lld", node, (long long)need, (long long)n);
  if ((int)e->injHost.size() != e->nNodes) e->injHost.resize(e->nNodes);
  e->injHost[node].assign(u, u + need);
  return DCSMC_OK;
}

int32_t dcsmc_run(dcsmc_engine* e) {
  if (!e) return DCSMC_EINVAL;
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree has not been called");
  // a fresh run recomputes everything
  dcsmc_reset_populations(e);
  return runRoots(e, std::vector<int>{e->root});
}

int32_t dcsmc_run_subtrees(dcsmc_engine* e, const int32_t* roots, int32_t nRoots) {
  if (!e) return DCSMC_EINVAL;
  if (!roots || nRoots < 0) return e->fail(DCSMC_EINVAL, "bad roots");
  std::vector<int> r(roots, roots + nRoots);
  for (int x : r)
    if (x < 0 || x >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad root %d", x);
  return runRoots(e, r);
}

int32_t dcsmc_run_nodes(dcsmc_engine* e, const int32_t* nodes, int32_t nNodes) {
  if (!e) return DCSMC_EINVAL;
  if (!nodes || nNodes < 0) return e->fail(DCSMC_EINVAL, "bad nodes");
  // DCRecursionTask.getChildrenPopulationsFromCluster (:108-115) fails when a child population is missing; a
  // silent recomputation of the child's subtree would hide a missed transfer
  for (int k = 0; k < nNodes; ++k) {
    const int n = nodes[k];
    if (n < 0 || n >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", n);
    for (int c = e->childOffsets[n]; c < e->childOffsets[n + 1]; ++c) {
      const int ch = e->children[c];
      bool inSet = false;
      for (int j = 0; j < nNodes && !inSet; ++j) inSet = nodes[j] == ch;
      if (!inSet && (e->resident.empty() || !e->resident[ch]))
        return e->fail(DCSMC_ESTATE, "node %d: the population of child %d is not resident (dcsmc_run_subtrees, "
                                      "dcsmc_population_unpack or dcsmc_population_import_peer first)", n, ch);
    }
  }
  // children resident => collectSubtree stops at them, so this is the same planner
  return dcsmc_run_subtrees(e, nodes, nNodes);
}

int32_t dcsmc_node_stats(const dcsmc_engine* e, int32_t node, dcsmc_node_stats_t* out) {
  if (!e || !out) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= (int)e->hStats.size()) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  *out = e->hStats[node];
  return DCSMC_OK;
}

int32_t dcsmc_all_node_stats(const dcsmc_engine* e, dcsmc_node_stats_t* out) {
  if (!e || !out) return DCSMC_EINVAL;
  flushStats(e);
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree has not been called");
  if ((int)e->hStats.size() != e->nNodes) {  // nothing computed on this engine yet: done == 0 everywhere
    memset(out, 0, sizeof(dcsmc_node_stats_t) * e->nNodes);
    return DCSMC_OK;
  }
  memcpy(out, e->hStats.data(), sizeof(dcsmc_node_stats_t) * e->nNodes);
  return DCSMC_OK;
}

int32_t dcsmc_all_node_stats_columns(const dcsmc_engine* e, double* ess, double* relEss, double* logNormEstimate,
                                     double* logScaling, int32_t* resampled, int32_t* done) {
  if (!e) return DCSMC_EINVAL;
  flushStats(e);
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree has not been called");
  const size_t nN = (size_t)e->nNodes;
  const bool have = e->hStats.size() == nN;  // else nothing was computed on this engine yet: done == 0 everywhere
  const dcsmc_node_stats_t zero{};
  for (size_t n = 0; n < nN; ++n) {
    const dcsmc_node_stats_t& s = have ? e->hStats[n] : zero;
    if (ess) ess[n] = s.ess;
    if (relEss) relEss[n] = s.relEss;
    if (logNormEstimate) logNormEstimate[n] = s.logNormEstimate;
    if (logScaling) logScaling[n] = s.logScaling;
    if (resampled) resampled[n] = s.resampled;
    if (done) done[n] = s.done;
  }
  return DCSMC_OK;
}

int32_t dcsmc_node_summary(dcsmc_engine* e, int32_t node, dcsmc_node_summary_t* out) {
  if (!e || !out) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  memset(out, 0, sizeof *out);
  if (e->opt.levelCutOffForOutput <= 0 || e->dSumm == nullptr) return DCSMC_OK;
  if ((int)e->hStats.size() != e->nNodes || !e->hStats[node].done) return e->fail(DCSMC_ESTATE, "node %d not computed", node);
  CK(cudaSetDevice(e->opt.device));
  NodeSummary h;
  CK(cudaMemcpyAsync(&h, e->dSumm + node, sizeof h, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  out->meanMean = h.meanMean;
  out->meanSD = h.meanSD;
  out->varMean = h.varMean;
  out->varSD = h.varSD;
  out->deltaMean = h.deltaMean;
  out->deltaSD = h.deltaSD;
  out->hasMean = (h.flags & 1) ? 1 : 0;
  out->hasVar = (h.flags & 2) ? 1 : 0;
  out->hasDelta = (h.flags & 4) ? 1 : 0;
  return DCSMC_OK;
}

int32_t dcsmc_node_samples(dcsmc_engine* e, int32_t node, int32_t which, double* out) {
  if (!e || !out) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  if (!e->opt.retainSamples || e->opt.levelCutOffForOutput <= 0 || e->dSumScratch == nullptr ||
      (int)e->sampleBase.size() != e->nNodes || e->sampleBase[node] < 0)
    return e->fail(DCSMC_ESTATE, "no retained samples for node %d (retainSamples = 1, level() < levelCutOffForOutput, "
                                 "computed by the last run)", node);
  if ((int)e->hStats.size() != e->nNodes || !e->hStats[node].done) return e->fail(DCSMC_ESTATE, "node %d not computed", node);
  if (!e->hStats[node].resampled) return e->fail(DCSMC_ESTATE, "node %d was not resampled: no samples (statistics assume equal weights)", node);
  const bool internal = nodeKind(e, node) == KIND_INTERNAL;
  const int nS = summarySeries(e, node);
  int series;  // layout of k_summary_samples: [mean][var][delta 0..C-1][natural]
  if (which == DCSMC_SAMPLE_NATURAL) series = nS - 1;
  else if (which == DCSMC_SAMPLE_MEAN) series = 0;
  else if (which == DCSMC_SAMPLE_VAR && internal) series = 1;
  else if (which >= DCSMC_SAMPLE_DELTA0 && summaryReportsDeltas(e, node) && which - DCSMC_SAMPLE_DELTA0 < nChildrenOf(e, node))
    series = (internal ? 2 : 1) + (which - DCSMC_SAMPLE_DELTA0);
  else return e->fail(DCSMC_EINVAL, "node %d has no sample series %d", node, which);
  CK(cudaSetDevice(e->opt.device));
  CK(cudaMemcpyAsync(out, e->dSumScratch + ((size_t)e->sampleBase[node] + (size_t)series) * (size_t)e->Npad,
                     sizeof(double) * (size_t)e->opt.nParticles, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return DCSMC_OK;
}

int32_t dcsmc_node_times(const dcsmc_engine* e, double* msPerNode) {
  if (!e || !msPerNode) return DCSMC_EINVAL;
  if (e->nNodes <= 0) return e->fail(DCSMC_ESTATE, "dcsmc_set_tree has not been called");
  if ((int)e->hNodeMs.size() != e->nNodes) {
    memset(msPerNode, 0, sizeof(double) * (size_t)e->nNodes);
    return DCSMC_OK;
  }
  memcpy(msPerNode, e->hNodeMs.data(), sizeof(double) * (size_t)e->nNodes);
  return DCSMC_OK;
}

int32_t dcsmc_population_view(dcsmc_engine* e, int32_t node, dcsmc_population_view_t* out) {
  if (!e || !out) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  if (node >= (int)e->resident.size() || node >= (int)e->hStats.size() || !e->resident[node] || !e->hStats[node].done)
    return e->fail(DCSMC_ESTATE, "population of node %d is not resident (consumed by its parent, or not computed)", node);
  memset(out, 0, sizeof *out);
  fillNodeDev(e, node);
  const NodeDev& d = e->hNodes[node];
  const NodeStatsOut& h = e->hOut[node];
  const bool constLeaf = skipObservedLeaves(e) && d.kind == KIND_LEAF_GAUSS;
  out->nParticles = e->opt.nParticles;
  out->stride = e->Npad;
  out->state = constLeaf ? nullptr : d.state;
  out->nColumns = d.state == nullptr || constLeaf ? 0 : (d.kind == KIND_INTERNAL ? 6 : 2);
  out->istate = d.istate;
  out->logWeights = d.logw;
  out->resampled = h.resampled == RES_WEIGHTED ? 0 : 1;
  out->ancestors = (h.resampled == RES_ANCESTORS && !constLeaf) ? d.anc : nullptr;
  out->onPeer = e->slot[node].peer ? 1 : 0;
  out->constantObservation = constLeaf ? 1 : 0;
  out->observation = (!e->leafObs.empty() && d.kind == KIND_LEAF_GAUSS) ? e->leafObs[node] : 0.0;
  out->logScaling = h.logScaling;
  out->maxLogWeight = h.m;
  out->sumExp = h.S;
  out->ess = h.ess;
  return DCSMC_OK;
}

int32_t dcsmc_population_copy_to_host(dcsmc_engine* e, int32_t node, double* state, int32_t* istate, double* weights) {
  if (!e) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  if (node >= (int)e->resident.size() || node >= (int)e->hStats.size() || !e->resident[node] || !e->hStats[node].done)
    return e->fail(DCSMC_ESTATE, "population of node %d is not resident (consumed by its parent, or not computed)", node);
  CK(cudaSetDevice(e->opt.device));
  const int N = e->opt.nParticles;
  // device staging + pinned host staging, both kept across calls: the copy runs at PCIe speed into
  // pinned memory and the caller's (pageable) arrays are filled by a plain memcpy
  const size_t bS = state ? sizeof(double) * 6 * (size_t)N : 0, bI = istate ? sizeof(int32_t) * (size_t)N : 0,
               bW = weights ? sizeof(double) * (size_t)N : 0;
  const size_t oI = alignUp(bS, 256), oW = oI + alignUp(bI, 256), total = oW + alignUp(bW, 256);
  if (total > e->stageCap) {
    if (e->dStage) cudaFree(e->dStage);
    if (e->hStage) cudaFreeHost(e->hStage);
    e->dStage = e->hStage = nullptr;
    e->stageCap = 0;
    CK(cudaMalloc((void**)&e->dStage, total));
    CK(cudaMallocHost((void**)&e->hStage, total));
    e->stageCap = total;
  }
  double* dS = state ? reinterpret_cast<double*>(e->dStage) : nullptr;
  int32_t* dI = istate ? reinterpret_cast<int32_t*>(e->dStage + oI) : nullptr;
  double* dW = weights ? reinterpret_cast<double*>(e->dStage + oW) : nullptr;
  RunCtx c;
  makeCtx(e, c);
  e->info.kernelLaunches++;
  k_materialise<<<(N + 255) / 256, 256, 0, e->stream>>>(c, node, dS, dI, dW, (size_t)N);
  CK(cudaGetLastError());
  // staging range [off, off + len) -> the caller's arrays (a range may straddle the three sections)
  auto copyOut = [&](size_t off, size_t len) {
    const struct { size_t at, bytes; char* dst; } sec[3] = {{0, bS, reinterpret_cast<char*>(state)},
                                                            {oI, bI, reinterpret_cast<char*>(istate)},
                                                            {oW, bW, reinterpret_cast<char*>(weights)}};
    for (const auto& q : sec) {
      if (!q.dst || q.bytes == 0) continue;
      const size_t lo = std::max(off, q.at), hi = std::min(off + len, q.at + q.bytes);
      if (lo < hi) memcpy(q.dst + (lo - q.at), e->hStage + lo, hi - lo);
    }
  };
  constexpr size_t PIPE_CHUNK = 4u << 20;
  if (total <= PIPE_CHUNK) {
    CK(cudaMemcpyAsync(e->hStage, e->dStage, total, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    copyOut(0, total);
    return DCSMC_OK;
  }
  // Large populations (cfg3's root: 56 MB): the device -> pinned copy is cut into 4 MiB pieces, each followed by an
  // event, and up to four host threads move finished pieces into the caller's arrays while the next ones are still on
  // the bus — the call costs max(PCIe, host memcpy) instead of their sum, and the host side is no longer one thread's
  // memcpy (r02d: the end-to-end step of the bench spent more time here than in all the host-side planning).
  const size_t nChunks = (total + PIPE_CHUNK - 1) / PIPE_CHUNK;
  while (e->stageEvents.size() < nChunks) {
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->stageEvents.push_back(ev);
  }
  for (size_t k = 0; k < nChunks; ++k) {
    const size_t off = k * PIPE_CHUNK, len = std::min(PIPE_CHUNK, total - off);
    CK(cudaMemcpyAsync(e->hStage + off, e->dStage + off, len, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaEventRecord(e->stageEvents[k], e->stream));
  }
  const int nThreads = (int)std::min<size_t>(4, nChunks);
  std::vector<cudaError_t> terr(nThreads, cudaSuccess);
  auto worker = [&](int t) {
    cudaSetDevice(e->opt.device);
    for (size_t k = (size_t)t; k < nChunks; k += (size_t)nThreads) {
      const cudaError_t r = cudaEventSynchronize(e->stageEvents[k]);
      if (r != cudaSuccess) {
        terr[t] = r;
        return;
      }
      copyOut(k * PIPE_CHUNK, std::min(PIPE_CHUNK, total - k * PIPE_CHUNK));
    }
  };
  std::vector<std::thread> pool;
  int started = 1;  // this thread is worker 0; a thread that cannot be created leaves its pieces to the sweep below
  try {
    for (int t = 1; t < nThreads; ++t) {
      pool.emplace_back(worker, t);
      ++started;
    }
  } catch (...) {
  }
  worker(0);
  for (auto& th : pool) th.join();
  for (int t = started; t < nThreads; ++t) worker(t);  // no exception may cross the C ABI
  for (int t = 0; t < nThreads; ++t) CK(terr[t]);
  CK(cudaStreamSynchronize(e->stream));
  return DCSMC_OK;
}

int32_t dcsmc_debug_copy_node(dcsmc_engine* e, int32_t node, double* statePre, int32_t* istatePre,
                              double* logWeights, int32_t* ancestors) {
  if (!e) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  if (!e->opt.retainPopulations) return e->fail(DCSMC_ESTATE, "dcsmc_debug_copy_node needs retainPopulations=1");
  if (node >= (int)e->resident.size() || node >= (int)e->hStats.size() || !e->resident[node] || !e->hStats[node].done)
    return e->fail(DCSMC_ESTATE, "node %d not computed", node);
  CK(cudaSetDevice(e->opt.device));
  const int N = e->opt.nParticles;
  const size_t NP = (size_t)e->Npad;
  const NodeDev& d = e->hNodes[node];
  if (statePre) {
    const double nanv = std::nan("");
    if (!d.state) {
      for (size_t k = 0; k < (size_t)6 * N; ++k) statePre[k] = 0.0;
    } else if (d.kind == KIND_INTERNAL) {
      for (int col = 0; col < 6; ++col)
        CK(cudaMemcpyAsync(statePre + (size_t)col * N, d.state + col * NP, sizeof(double) * N, cudaMemcpyDeviceToHost, e->stream));
    } else {
      CK(cudaMemcpyAsync(statePre, d.state, sizeof(double) * N, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaMemcpyAsync(statePre + (size_t)3 * N, d.state + NP, sizeof(double) * N, cudaMemcpyDeviceToHost, e->stream));
      for (int i = 0; i < N; ++i) {
        statePre[(size_t)1 * N + i] = 0.0;
        statePre[(size_t)2 * N + i] = 0.0;
        statePre[(size_t)4 * N + i] = 0.0;
        statePre[(size_t)5 * N + i] = nanv;
      }
    }
  }
  if (istatePre) {
    if (d.istate)
      CK(cudaMemcpyAsync(istatePre, d.istate, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, e->stream));
    else
      memset(istatePre, 0, sizeof(int32_t) * N);
  }
  if (logWeights) {
    if (d.logw) {
      CK(cudaMemcpyAsync(logWeights, d.logw, sizeof(double) * N, cudaMemcpyDeviceToHost, e->stream));
    } else {
      const double c0 = uniformLeafLogW(e);
      for (int i = 0; i < N; ++i) logWeights[i] = c0;
    }
  }
  CK(cudaStreamSynchronize(e->stream));
  if (ancestors) {
    if (e->hStats[node].resampled) {
      CK(cudaMemcpy(ancestors, d.anc, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
    } else {
      for (int i = 0; i < N; ++i) ancestors[i] = i;
    }
  }
  return DCSMC_OK;
}

int32_t dcsmc_debug_eval(dcsmc_engine* e, int32_t op, const double* x, double* y, int64_t n) {
  if (!e) return DCSMC_EINVAL;
  if (!x || !y || n < 0 || op < 0 || (op & 0xff) > 6) return e->fail(DCSMC_EINVAL, "dcsmc_debug_eval: bad arguments");
  if (n == 0) return DCSMC_OK;
  CK(cudaSetDevice(e->opt.device));
  double *dx = nullptr, *dy = nullptr;
  CK(cudaMalloc((void**)&dx, sizeof(double) * n));
  CK(cudaMalloc((void**)&dy, sizeof(double) * n));
  CK(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
  k_debug_eval<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(op, dx, dy, n);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(y, dy, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(dx);
  cudaFree(dy);
  return DCSMC_OK;
}

// The same header functions compiled for the HOST (csrc/dcsmc_math.cuh is host+device code): lets the CPU test
// suite check the engine's own fdlibm / Philox against the oracle without a GPU.
int32_t dcsmc_host_eval(int32_t op, const double* x, double* y, int64_t n) {
  if (!x || !y || n < 0 || op < 0 || (op & 0xff) > 2) return DCSMC_EINVAL;
  const PhiloxKeys rk = philox_round_keys(31ULL);
  for (int64_t i = 0; i < n; ++i) {
    if ((op & 0xff) == 2) {
      double u0, u1;
      const uint64_t idx = (uint64_t)x[i];
      philox_uniform_pair_rk(rk, op >> 8, 0u, idx >> 1, u0, u1);
      y[i] = (idx & 1) ? u1 : u0;
    } else {
      y[i] = op == 0 ? dc_log(x[i]) : dc_exp(x[i]);
    }
  }
  return DCSMC_OK;
}

int64_t dcsmc_population_packed_bytes(const dcsmc_engine* e) {
  if (!e) return 0;
  const size_t NP = alignUp((size_t)e->opt.nParticles, 32);
  return (int64_t)(sizeof(PackedHeader) + NP * 8 * 7 + NP * 4);
}

int32_t dcsmc_population_pack(dcsmc_engine* e, int32_t node, void* deviceDst) {
  if (!e) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes || !deviceDst) return e->fail(DCSMC_EINVAL, "dcsmc_population_pack: bad arguments");
  if (node >= (int)e->resident.size() || node >= (int)e->hStats.size() || !e->resident[node] || !e->hStats[node].done)
    return e->fail(DCSMC_ESTATE, "population of node %d is not resident", node);
  CK(cudaSetDevice(e->opt.device));
  RunCtx c;
  makeCtx(e, c);
  const int N = e->opt.nParticles;
  e->info.kernelLaunches++;
  k_pack<<<(N + 255) / 256, 256, 0, e->stream>>>(c, node, static_cast<char*>(deviceDst));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  return DCSMC_OK;
}

int32_t dcsmc_population_unpack(dcsmc_engine* e, int32_t node, const void* deviceSrc) {
  if (!e) return DCSMC_EINVAL;
  flushStats(e);
  if (node < 0 || node >= e->nNodes || !deviceSrc) return e->fail(DCSMC_EINVAL, "dcsmc_population_unpack: bad arguments");
  int rc;
  if ((rc = prepare(e))) return rc;
  PackedHeader h;
  CK(cudaMemcpy(&h, deviceSrc, sizeof h, cudaMemcpyDeviceToHost));
  if (h.magic != PACK_MAGIC || h.N != e->opt.nParticles || h.model != (e->model == DCSMC_MODEL_MARKOV ? 1 : 0))
    return e->fail(DCSMC_EINVAL, "packed population does not match this engine (magic/N/model)");
  poolFree(e, node);
  char* buf = nullptr;
  if (!e->importFree.empty()) {
    buf = e->importFree.back();
    e->importFree.pop_back();
  } else {
    CK(cudaMalloc((void**)&buf, importBytes(e)));
  }
  e->slot[node] = Slot{0, 0, buf};
  e->externalNodes.push_back(node);
  e->statNodes.push_back(node);
  e->resident[node] = 1;
  e->imported[node] = 1;
  fillNodeDev(e, node);
  descWritten(e, node, nullptr);
  CK(cudaMemcpyAsync(e->dNodes + node, &e->hNodes[node], sizeof(NodeDev), cudaMemcpyHostToDevice, e->stream));
  RunCtx c;
  makeCtx(e, c);
  const int N = e->opt.nParticles;
  e->info.kernelLaunches++;
  k_unpack<<<(N + 255) / 256, 256, 0, e->stream>>>(c, node, static_cast<const char*>(deviceSrc));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  dcsmc_node_stats_t& o = e->hStats[node];
  o.ess = h.ess;
  o.relEss = h.relEss;
  o.logNormEstimate = h.logZ;
  o.logScaling = h.logScaling;
  o.resampled = h.resampled;
  o.done = 1;
  return DCSMC_OK;
}

// ---- zero-copy sharing over NVLink peer memory ---------------------------------------------------

int32_t dcsmc_peer_pool_handle(dcsmc_engine* e, void* handle, int64_t* generation, int64_t* poolBytes) {
  if (!e || !handle || !generation || !poolBytes) return DCSMC_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == DCSMC_IPC_HANDLE_BYTES, "IPC handle size");
  int rc;
  if ((rc = prepare(e))) return rc;  // allocates the pool
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->pool));
  memcpy(handle, &h, sizeof h);
  *generation = e->poolGeneration;
  *poolBytes = (int64_t)e->poolBytes;
  return DCSMC_OK;
}

int32_t dcsmc_peer_open(dcsmc_engine* e, int32_t peer, const void* handle, int64_t generation, int64_t poolBytes) {
  if (!e || !handle || peer < 0 || poolBytes < 0) return DCSMC_EINVAL;
  CK(cudaSetDevice(e->opt.device));
  dcsmc_engine::PeerMap& pm = e->peers[peer];
  if (pm.base && pm.generation == generation) {
    pm.bytes = poolBytes;
    return DCSMC_OK;
  }
  if (pm.base) {
    cudaIpcCloseMemHandle(pm.base);
    pm.base = nullptr;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  pm.base = static_cast<char*>(p);
  pm.generation = generation;
  pm.bytes = poolBytes;
  return DCSMC_OK;
}

int32_t dcsmc_population_export(dcsmc_engine* e, const int32_t* nodes, int32_t n, void* descs) {
  if (!e || n < 0 || (n > 0 && (!nodes || !descs))) return DCSMC_EINVAL;
  flushStats(e);
  ExportDesc* out = static_cast<ExportDesc*>(descs);
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
    if (node >= (int)e->resident.size() || node >= (int)e->hStats.size() || !e->resident[node] || !e->hStats[node].done)
      return e->fail(DCSMC_ESTATE, "population of node %d is not resident", node);
    if (e->slot[node].external)
      return e->fail(DCSMC_ESTATE, "node %d is itself an imported population; only pool residents can be exported", node);
    const NodeStatsOut& h = e->hOut[node];
    ExportDesc d;
    memset(&d, 0, sizeof d);
    d.magic = PACK_MAGIC;
    d.N = e->opt.nParticles;
    d.offset = (int64_t)e->slot[node].offset;
    d.generation = e->poolGeneration;
    d.m = h.m;
    d.S = h.S;
    d.ess = h.ess;
    d.relEss = h.relEss;
    d.logScaling = h.logScaling;
    d.logZ = h.logZ;
    d.resampled = h.resampled;
    d.done = h.done;
    d.slotBytes = (int64_t)e->slot[node].bytes;
    d.model = e->model;
    d.kind = nodeKind(e, node);
    d.node = node;
    d.skipObs = skipObservedLeaves(e) ? 1 : 0;
    out[k] = d;
  }
  return DCSMC_OK;
}

int32_t dcsmc_population_import_peer(dcsmc_engine* e, const int32_t* nodes, const int32_t* peerIds, int32_t n,
                                     const void* descs) {
  if (!e || n < 0 || (n > 0 && (!nodes || !peerIds || !descs))) return DCSMC_EINVAL;
  if (n == 0) return DCSMC_OK;
  flushStats(e);
  int rc;
  if ((rc = prepare(e))) return rc;
  const ExportDesc* in = static_cast<const ExportDesc*>(descs);
  // pass 1: validate EVERY descriptor before any state changes — a descriptor is foreign input (another process
  // wrote it, a collective carried it): it must describe a finished population of the same model and layout
  // that lies inside the mapped pool of that peer
  if ((rc = validateImport(e, nodes, peerIds, n, in, false))) return rc;
  const size_t oNd = alignUp(sizeof(int32_t) * n, 256), oSt = oNd + alignUp(sizeof(NodeDev) * n, 256),
               total = oSt + alignUp(sizeof(NodeStats) * n, 256);
  // the pinned staging buffer is reused: the previous upload must have left it
  if (e->evImp) CK(cudaEventSynchronize(e->evImp));
  if (total > e->impCap) {
    if (e->dImp) cudaFree(e->dImp);
    if (e->hImp) cudaFreeHost(e->hImp);
    e->dImp = e->hImp = nullptr;
    e->impCap = 0;
    CK(cudaMalloc((void**)&e->dImp, total));
    CK(cudaMallocHost((void**)&e->hImp, total));
    e->impCap = total;
  }
  if (!e->evImp) CK(cudaEventCreateWithFlags(&e->evImp, cudaEventDisableTiming));
  int32_t* hIds = reinterpret_cast<int32_t*>(e->hImp);
  NodeDev* hNd = reinterpret_cast<NodeDev*>(e->hImp + oNd);
  NodeStats* hSt = reinterpret_cast<NodeStats*>(e->hImp + oSt);
  // pass 2: install
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    const ExportDesc& d = in[k];
    const dcsmc_engine::PeerMap& pm = e->peers[peerIds[k]];
    poolFree(e, node);
    Slot sl;
    sl.external = pm.base + d.offset;
    sl.peer = true;
    e->slot[node] = sl;
    e->externalNodes.push_back(node);
    e->statNodes.push_back(node);
    e->resident[node] = 1;
    e->imported[node] = 2;
    fillNodeDev(e, node);
    descWritten(e, node, nullptr);
    hIds[k] = node;
    hNd[k] = e->hNodes[node];
    NodeStats z;
    memset(&z, 0, sizeof z);
    z.maxKey = ordered_key(d.m);
    z.m = d.m;
    z.S = d.S;
    z.ess = d.ess;
    z.relEss = d.relEss;
    z.logScaling = d.logScaling;
    z.logZ = d.logZ;
    z.resampled = d.resampled;
    z.done = 2;  // a population computed elsewhere (k_stats_matrix counts done == 1 only)
    hSt[k] = z;
    dcsmc_node_stats_t& o = e->hStats[node];
    o.ess = d.ess;
    o.relEss = d.relEss;
    o.logNormEstimate = d.logZ;
    o.logScaling = d.logScaling;
    o.resampled = d.resampled == RES_WEIGHTED ? 0 : 1;
    o.done = 1;
    NodeStatsOut& ho = e->hOut[node];
    ho.ess = d.ess; ho.relEss = d.relEss; ho.logZ = d.logZ; ho.logScaling = d.logScaling;
    ho.m = d.m; ho.S = d.S; ho.resampled = d.resampled; ho.done = 1;
  }
  CK(cudaMemcpyAsync(e->dImp, e->hImp, total, cudaMemcpyHostToDevice, e->stream));
  CK(cudaEventRecord(e->evImp, e->stream));
  e->info.kernelLaunches++;
  k_scatter_import<<<(n + 127) / 128, 128, 0, e->stream>>>(e->dNodes, e->dStats, reinterpret_cast<int32_t*>(e->dImp),
                                                          reinterpret_cast<NodeDev*>(e->dImp + oNd),
                                                          reinterpret_cast<NodeStats*>(e->dImp + oSt), n);
  CK(cudaGetLastError());
  return DCSMC_OK;  // stream order: the next run on this engine sees the imported descriptors
}

// ---- sharded runs without host round trips -------------------------------------------------------------
int32_t dcsmc_stream(dcsmc_engine* e, void** stream) {
  if (!e || !stream) return DCSMC_EINVAL;
  *stream = (void*)e->stream;
  return DCSMC_OK;
}

int32_t dcsmc_begin_deferred(dcsmc_engine* e) {
  if (!e) return DCSMC_EINVAL;
  if (e->profile) return e->fail(DCSMC_ESTATE, "deferred calls cannot be profiled per launch (dcsmc_profile_enable)");
  int rc;
  if ((rc = prepare(e))) return rc;
  flushStats(e);
  if (!e->dErr) {
    CK(cudaMalloc((void**)&e->dErr, sizeof(int32_t)));
    CK(cudaMallocHost((void**)&e->hErr, sizeof(int32_t)));
  }
  CK(cudaMemsetAsync(e->dErr, 0, sizeof(int32_t), e->stream));
  e->deferred = true;
  e->callEventsUsed = 0;
  e->info = dcsmc_run_info{};
  return DCSMC_OK;
}

int32_t dcsmc_sync(dcsmc_engine* e) {
  if (!e) return DCSMC_EINVAL;
  if (!e->deferred) return e->fail(DCSMC_ESTATE, "dcsmc_sync without dcsmc_begin_deferred");
  CK(cudaSetDevice(e->opt.device));
  e->deferred = false;
  *e->hErr = 0;
  CK(cudaMemcpyAsync(e->hErr, e->dErr, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->inFlight = false;
  if (*e->hErr != 0)
    return e->fail(DCSMC_ESTATE, "deferred import of node %d: the owner's population is not where the cached layout "
                                  "descriptor says (or not finished) — the static schedule changed; results of this run are invalid",
                   *e->hErr - 1);
  double total = 0.0;
  for (size_t k = 0; k < e->callEventsUsed; ++k) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e->callEvents[k].first, e->callEvents[k].second));
    total += ms;
  }
  e->callEventsUsed = 0;
  e->info.deviceMs = total;
  e->info.poolBytes = (int64_t)e->poolBytes;
  return DCSMC_OK;
}

int32_t dcsmc_stats_export_device(dcsmc_engine* e, const int32_t* nodes, int32_t n, void* deviceDst) {
  if (!e || n < 0 || (n > 0 && (!nodes || !deviceDst))) return DCSMC_EINVAL;
  if (n == 0) return DCSMC_OK;
  static_assert(sizeof(NodeStatsOut) == DCSMC_STATS_RECORD_BYTES, "statistics record size");
  CK(cudaSetDevice(e->opt.device));
  for (int k = 0; k < n; ++k)
    if (nodes[k] < 0 || nodes[k] >= e->nNodes || nodes[k] >= (int)e->resident.size() || !e->resident[nodes[k]])
      return e->fail(DCSMC_ESTATE, "population of node %d is not resident", nodes[k]);
  char* dIds = nullptr;
  int rc;
  if ((rc = cachedBlob(e, nodes, sizeof(int32_t) * (size_t)n, &dIds))) return rc;
  e->info.kernelLaunches++;
  k_export_stats<<<(n + 127) / 128, 128, 0, e->stream>>>(e->dStats, reinterpret_cast<const int32_t*>(dIds), n,
                                                        static_cast<NodeStatsOut*>(deviceDst), e->dNodes, e->pool,
                                                        (int32_t)e->poolGeneration);
  CK(cudaGetLastError());
  return DCSMC_OK;
}

int32_t dcsmc_population_export_layout(dcsmc_engine* e, const int32_t* nodes, int32_t n, void* descs) {
  if (!e || n < 0 || (n > 0 && (!nodes || !descs))) return DCSMC_EINVAL;
  ExportDesc* out = static_cast<ExportDesc*>(descs);
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    if (node < 0 || node >= e->nNodes) return e->fail(DCSMC_EINVAL, "bad node %d", node);
    if (node >= (int)e->resident.size() || !e->resident[node])
      return e->fail(DCSMC_ESTATE, "population of node %d is not resident", node);
    if (e->slot[node].external)
      return e->fail(DCSMC_ESTATE, "node %d is itself an imported population; only pool residents can be exported", node);
    ExportDesc d;
    memset(&d, 0, sizeof d);
    d.magic = PACK_MAGIC;
    d.N = e->opt.nParticles;
    d.offset = (int64_t)e->slot[node].offset;
    d.generation = e->poolGeneration;
    d.done = 2;  // layout only: the statistics travel on the device (dcsmc_stats_export_device)
    d.slotBytes = (int64_t)e->slot[node].bytes;
    d.model = e->model;
    d.kind = nodeKind(e, node);
    d.node = node;
    d.skipObs = skipObservedLeaves(e) ? 1 : 0;
    out[k] = d;
  }
  return DCSMC_OK;
}


int32_t dcsmc_population_import_peer_device(dcsmc_engine* e, const int32_t* nodes, const int32_t* peerIds, int32_t n,
                                            const void* descs, const void* deviceStats, const int32_t* statIndex,
                                            int32_t nRecords) {
  if (!e || n < 0 || (n > 0 && (!nodes || !peerIds || !descs || !deviceStats || !statIndex))) return DCSMC_EINVAL;
  if (n == 0) return DCSMC_OK;
  int rc;
  if ((rc = prepare(e))) return rc;
  const ExportDesc* in = static_cast<const ExportDesc*>(descs);
  if ((rc = validateImport(e, nodes, peerIds, n, in, true))) return rc;
  for (int k = 0; k < n; ++k)
    if (statIndex[k] < 0 || statIndex[k] >= nRecords) return e->fail(DCSMC_EINVAL, "statistics record %d out of range", statIndex[k]);
  if (!e->deferred) return e->fail(DCSMC_ESTATE, "dcsmc_population_import_peer_device belongs between dcsmc_begin_deferred and dcsmc_sync");
  // host bookkeeping + one blob [ids | NodeDev[] | statIdx] found again by content on every later run
  const size_t oNd = alignUp(sizeof(int32_t) * n, 16), oIx = oNd + alignUp(sizeof(NodeDev) * n, 16),
               oEx = oIx + alignUp(sizeof(int32_t) * n, 16), total = oEx + sizeof(int2) * n;
  std::vector<char> blob(total, 0);
  int32_t* hIds = reinterpret_cast<int32_t*>(blob.data());
  NodeDev* hNd = reinterpret_cast<NodeDev*>(blob.data() + oNd);
  int32_t* hIx = reinterpret_cast<int32_t*>(blob.data() + oIx);
  int2* hEx = reinterpret_cast<int2*>(blob.data() + oEx);
  for (int k = 0; k < n; ++k) {
    const int node = nodes[k];
    const dcsmc_engine::PeerMap& pm = e->peers[peerIds[k]];
    poolFree(e, node);
    Slot sl;
    sl.external = pm.base + in[k].offset;
    sl.peer = true;
    e->slot[node] = sl;
    e->externalNodes.push_back(node);
    e->resident[node] = 1;
    e->imported[node] = 2;
    fillNodeDev(e, node);
    descWritten(e, node, nullptr);
    hIds[k] = node;
    hNd[k] = e->hNodes[node];
    hIx[k] = statIndex[k];
    hEx[k] = make_int2((int32_t)(in[k].offset >> 8), (int32_t)in[k].generation);
  }
  char* dBlob = nullptr;
  if ((rc = cachedBlob(e, blob.data(), total, &dBlob))) return rc;
  // the imported statistics also go to the host's staging buffer (scattered into hStats when somebody asks)
  const size_t off = e->pendNodes.size();
  if (off + (size_t)n > e->exportCap) {
    const size_t want = std::max<size_t>(off + (size_t)n, (size_t)e->nNodes + 1024);
    if (e->inFlight) CK(cudaStreamSynchronize(e->stream));
    NodeStatsOut *nd = nullptr, *nh = nullptr;
    CK(cudaMalloc((void**)&nd, want * sizeof(NodeStatsOut)));
    CK(cudaMallocHost((void**)&nh, want * sizeof(NodeStatsOut)));
    if (off) memcpy(nh, e->hExport, off * sizeof(NodeStatsOut));
    if (e->dExport) cudaFree(e->dExport);
    if (e->hExport) cudaFreeHost(e->hExport);
    e->dExport = nd;
    e->hExport = nh;
    e->exportCap = want;
  }
  e->info.kernelLaunches++;
  k_scatter_import_dev<<<(n + 127) / 128, 128, 0, e->stream>>>(
      e->dNodes, e->dStats, reinterpret_cast<const int32_t*>(dBlob), reinterpret_cast<const NodeDev*>(dBlob + oNd),
      reinterpret_cast<const int32_t*>(dBlob + oIx), reinterpret_cast<const int2*>(dBlob + oEx),
      static_cast<const NodeStatsOut*>(deviceStats), n, e->dExport + off, e->dErr);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(e->hExport + off, e->dExport + off, (size_t)n * sizeof(NodeStatsOut), cudaMemcpyDeviceToHost, e->stream));
  e->inFlight = true;
  e->pendNodes.insert(e->pendNodes.end(), nodes, nodes + n);
  e->statNodes.insert(e->statNodes.end(), nodes, nodes + n);
  return DCSMC_OK;
}

int32_t dcsmc_stats_matrix_device(dcsmc_engine* e, void* deviceDst) {
  if (!e || !deviceDst) return DCSMC_EINVAL;
  if (e->nNodes <= 0 || e->dStats == nullptr || e->tablesDirty) return e->fail(DCSMC_ESTATE, "nothing has run on this engine yet");
  CK(cudaSetDevice(e->opt.device));
  e->info.kernelLaunches++;
  k_stats_matrix<<<(e->nNodes + 255) / 256, 256, 0, e->stream>>>(e->dStats, e->nNodes, static_cast<double*>(deviceDst));
  CK(cudaGetLastError());
  return DCSMC_OK;
}

int32_t dcsmc_reset_populations(dcsmc_engine* e) {
  if (!e) return DCSMC_EINVAL;
  e->pendNodes.clear();  // statistics nobody asked for
  if (e->dStats != nullptr && !e->tablesDirty && e->nNodes > 0 && cudaSetDevice(e->opt.device) == cudaSuccess)
    cudaMemsetAsync(e->dStats, 0, sizeof(NodeStats) * (size_t)e->nNodes, e->stream);  // nothing is computed any more
  // O(nodes touched since the last reset), not O(nNodes): a rank of a sharded run only ever touches
  // its own part of a large tree
  for (int n : e->externalNodes)
    if (n < (int)e->slot.size() && e->slot[n].external) {
      if (!e->slot[n].peer) e->importFree.push_back(e->slot[n].external);
      e->slot[n].external = nullptr;
      e->slot[n].peer = false;
    }
  e->externalNodes.clear();
  if (e->statNodes.size() * 4 > e->hStats.size()) {
    std::fill(e->hStats.begin(), e->hStats.end(), dcsmc_node_stats_t{});
  } else {
    for (int n : e->statNodes)
      if (n < (int)e->hStats.size()) e->hStats[n] = dcsmc_node_stats_t{};
  }
  e->statNodes.clear();
  std::fill(e->resident.begin(), e->resident.end(), 0);
  std::fill(e->imported.begin(), e->imported.end(), 0);
  e->freeBySize.clear();
  e->poolTop = 0;
  return DCSMC_OK;
}

int32_t dcsmc_population_release(dcsmc_engine* e, int32_t node) {
  if (!e) return DCSMC_EINVAL;
  if (node < 0 || node >= e->nNodes || node >= (int)e->resident.size()) return e->fail(DCSMC_EINVAL, "bad node %d", node);
  poolFree(e, node);
  return DCSMC_OK;
}

int32_t dcsmc_plan_query(const dcsmc_options* opt, int32_t nNodes, const int32_t* childOffsets,
                         const int32_t* children, const int32_t* leafKind, int32_t model, const int32_t* roots,
                         int32_t nRoots, int64_t budgetBytes, dcsmc_plan_info* out, char* err, int32_t errLen) {
  auto say = [&](const char* msg) {
    if (err && errLen > 0) snprintf(err, (size_t)errLen, "%s", msg);
  };
  if (!opt || !out || nNodes < 1 || !childOffsets) {
    say("dcsmc_plan_query: bad arguments");
    return DCSMC_EINVAL;
  }
  if (opt->nParticles < 1 || opt->nParticles >= (1 << 24)) {
    say("nParticles must be in [1, 2^24)");
    return DCSMC_EINVAL;
  }
  memset(out, 0, sizeof *out);
  // a host-only engine object: no stream, no device memory, never passed to a function that touches CUDA
  std::unique_ptr<dcsmc_engine> holder(new dcsmc_engine());
  dcsmc_engine* e = holder.get();
  e->opt = *opt;
  int rc = dcsmc_set_tree(e, nNodes, 0, childOffsets, children);
  if (rc == DCSMC_OK) {
    if (model == DCSMC_MODEL_MARKOV) {
      e->model = DCSMC_MODEL_MARKOV;
      e->nStates = 2;
    } else {
      const std::vector<int32_t> zeros(nNodes, 0);  // leaf data do not influence the plan
      rc = dcsmc_set_multilevel_model(e, leafKind, zeros.data(), zeros.data(), nullptr, 1.0);
    }
  }
  if (rc != DCSMC_OK) {
    say(e->err.c_str());
    return rc;
  }
  const int N = opt->nParticles;
  e->Npad = (int)alignUp((size_t)N, 32);
  e->slot.assign(nNodes, Slot{});
  e->resident.assign(nNodes, 0);
  e->imported.assign(nNodes, 0);
  std::vector<int> rootSet;
  if (roots && nRoots > 0) {
    for (int k = 0; k < nRoots; ++k) {
      if (roots[k] < 0 || roots[k] >= nNodes) {
        say("dcsmc_plan_query: bad root");
        return DCSMC_EINVAL;
      }
      rootSet.push_back(roots[k]);
    }
  } else {
    rootSet.push_back(e->root);
  }
  size_t need = 0;  // every population of the planned subtrees resident at once
  {
    std::vector<int> nodes;
    for (int r : rootSet) collectSubtree(e, r, nodes);
    for (int n : nodes) need += slotBytes(e, n);
  }
  need = std::max<size_t>(need, 256);
  e->poolBytes = budgetBytes > 0 ? std::min<size_t>((size_t)budgetBytes, need) : need;
  e->poolIsBudget = e->poolBytes < need;
  Plan plan;
  rc = planRoots(e, rootSet, plan);
  out->poolBytesNeeded = (int64_t)need;
  out->poolBytes = (int64_t)e->poolBytes;
  out->feasible = rc == DCSMC_OK ? 1 : 0;
  if (rc != DCSMC_OK) {
    say(e->err.c_str());
    return rc == DCSMC_ENOMEM ? DCSMC_OK : rc;  // "does not fit" is an answer, not an error
  }
  size_t maxStep = 0, maxInternalStep = 0;
  for (auto& s : plan.steps) {
    maxStep = std::max<size_t>(maxStep, s.count);
    if (!s.leaf || opt->sumMode == DCSMC_SUM_JAVA) maxInternalStep = std::max<size_t>(maxInternalStep, s.count);
  }
  size_t invB = 1;
  while (invB * 8 < (size_t)N) invB <<= 1;
  const bool small = N <= SMALLN_MAX && opt->sumMode != DCSMC_SUM_JAVA;  // shared-memory node kernel: no CDF / index scratch
  size_t scratch = (maxStep * descWordsPerNode(N) + 2 * MAX_CHUNKS + 2) * sizeof(unsigned long long);
  if (!small) scratch += maxInternalStep * (size_t)e->Npad * sizeof(double) + maxInternalStep * (invB + 1) * sizeof(int32_t);
  if (opt->resamplingScheme == DCSMC_MULTINOMIAL) scratch += maxStep * (size_t)e->Npad * sizeof(int32_t);
  if (opt->resamplingScheme == DCSMC_MULTINOMIAL && !small && opt->sumMode == DCSMC_SUM_EXACT)  // partitioned darts
    scratch += maxInternalStep * ((size_t)dp_stride(N) * sizeof(double) +
                                  ((size_t)dp_chunks(N) * (size_t)(dp_segs(N) + 1) + (size_t)dp_segs(N)) * sizeof(int32_t));
  out->peakBytes = (int64_t)e->poolPeak;
  out->scratchBytes = (int64_t)scratch;
  out->tableBytes = (int64_t)((sizeof(NodeDev) + sizeof(NodeStats) + sizeof(LeafConst) + sizeof(uint64_t) + sizeof(double*) +
                               sizeof(NodeStatsOut)) * (size_t)nNodes + sizeof(int32_t) * (size_t)(2 * nNodes));
  out->nodesPlanned = (int64_t)plan.stepNodes.size();
  out->levelBatches = (int32_t)plan.steps.size();
  out->maxBatchNodes = (int32_t)maxStep;
  return DCSMC_OK;
}

int32_t dcsmc_last_run_info(const dcsmc_engine* e, dcsmc_run_info* out) {
  if (!e || !out) return DCSMC_EINVAL;
  *out = e->info;
  return DCSMC_OK;
}

int32_t dcsmc_timer_begin(dcsmc_engine* e) {
  if (!e) return DCSMC_EINVAL;
  CK(cudaSetDevice(e->opt.device));
  if (!e->evT0) {
    CK(cudaEventCreate(&e->evT0));
    CK(cudaEventCreate(&e->evT1));
  }
  CK(cudaEventRecord(e->evT0, e->stream));
  return DCSMC_OK;
}

int32_t dcsmc_timer_end(dcsmc_engine* e, double* ms) {
  if (!e || !ms) return DCSMC_EINVAL;
  if (!e->evT0) return e->fail(DCSMC_ESTATE, "dcsmc_timer_end without dcsmc_timer_begin");
  CK(cudaSetDevice(e->opt.device));
  CK(cudaEventRecord(e->evT1, e->stream));
  CK(cudaEventSynchronize(e->evT1));
  float t = 0;
  CK(cudaEventElapsedTime(&t, e->evT0, e->evT1));
  *ms = t;
  return DCSMC_OK;
}

int32_t dcsmc_profile_enable(dcsmc_engine* e, int32_t enable) {
  if (!e) return DCSMC_EINVAL;
  e->profile = enable != 0;
  for (int k = 0; k < DCSMC_K_COUNT; ++k) {
    e->profMs[k] = 0;
    e->profLaunches[k] = 0;
  }
  return DCSMC_OK;
}

int32_t dcsmc_profile_get(const dcsmc_engine* e, int32_t cls, double* ms, int64_t* launches) {
  if (!e) return DCSMC_EINVAL;
  if (cls < 0 || cls >= DCSMC_K_COUNT) return e->fail(DCSMC_EINVAL, "bad kernel class");
  if (ms) *ms = e->profMs[cls];
  if (launches) *launches = e->profLaunches[cls];
  return DCSMC_OK;
}

}  // extern "C"
