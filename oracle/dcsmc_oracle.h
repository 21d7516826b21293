/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * CPU oracle: plain-C restatement of the reference's DC-SMC node step
 * (alexandrebouchard/divide-and-conquer-smc). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, link or call it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - DCRecursion + ParticlePopulation + MULTINOMIAL + java.util.Random + Node.hashCode
 *     seeding: PINNED by the reference's one recorded run (Doc.java:282), reproduced by
 *     tests/test_oracle_golden.py to every printed digit.
 *   - hierarchical-binomial model (Beta leaf sampler, BrownianModelCalculator,
 *     Exponential proposal), STRATIFIED and the impl-1 posterior summaries
 *     (orc_node_summaries): PARITY UNPINNED — the reference holds no test, golden vector or
 *     fixture for them; tied to mathematics by this repo's own fixtures
 *     (tests/test_oracle_golden.py: closed-form Gaussian marginal likelihood for combine,
 *     2-D quadrature of a two-leaf binomial tree's logZ, moments of the Gaussian sampler)
 *     and to themselves by the committed vectors of tests/golden/.
 *   - ONE deliberate extension beyond the reference (DESIGN.md "Known deviations"): where the
 *     linear-space product of the children's weights underflows to exactly 0 (the reference
 *     then returns log 0 = -inf for every particle), the sum of the logs is used instead.
 */
#ifndef DCSMC_ORACLE_H
#define DCSMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RNG_JAVA = 0, ORC_RNG_PHILOX = 1, ORC_RNG_INJECTED = 2 };
enum { ORC_SUM_JAVA = 0, ORC_SUM_EXACT = 1 };
enum { ORC_ELEM_LIBM = 0, ORC_ELEM_FDLIBM = 1 };
enum { ORC_SCHEME_MULTINOMIAL = 0, ORC_SCHEME_STRATIFIED = 1 };
enum { ORC_MODEL_MULTILEVEL = 0, ORC_MODEL_MARKOV = 1 };
enum { ORC_LEAF_BINOMIAL = 0, ORC_LEAF_GAUSS_OBS = 1 };

/* Mirrors dc/DCOptions.java:15-45 (the fields that reach the node step) plus the
 * oracle's mode switches. */
typedef struct {
  int64_t masterRandomSeed;     /* DCOptions.java:16 */
  int32_t nParticles;           /* DCOptions.java:19 */
  int32_t resamplingScheme;     /* DCOptions.java:22 */
  double relativeEssThreshold;  /* DCOptions.java:25 */
  int32_t nThreads;             /* nThreadsPerNode, DCOptions.java:32: threads over
                                   independent subtrees (never over particles) */
  int32_t maximumDistributionDepth; /* DCOptions.java:41 */
  int32_t rngMode;
  int32_t sumMode;
  int32_t elemMode;
  int32_t model;
  int32_t impl;                 /* 2 = dc.* (per-node Random); 1 = prototype.smc
                                   DivideConquerMCAlgorithm.recurse (one global Random,
                                   always resample) */
  int32_t maxBetaAttempts;      /* slot budget per leaf particle = 2*maxBetaAttempts */
  double variancePriorRate;     /* MultiLevelProposalFactory.java:25-26 */
  /* Markov toy model (Doc.java:162-237): row-major transition[nStates*nStates], prior */
  int32_t nStates;
  const double* transition;
  const double* prior;
} orc_options;

typedef struct {
  int32_t nNodes;
  int32_t root;
  const int32_t* childOffsets; /* CSR, nNodes+1 */
  const int32_t* children;     /* children in reference order (DCRecursionTask.java:51) */
  const int32_t* javaHash;     /* Node.hashCode() per node (Node.java:34-40) */
  const int32_t* leafKind;     /* per node; ignored for internal nodes */
  const int32_t* nTrials;      /* per node (binomial leaves) */
  const int32_t* nSuccesses;
  const double* leafObs;       /* per node (gauss-obs leaves) */
} orc_tree;

/* Per-node outputs; any pointer may be NULL. Populations are dumped for every node
 * (tests use small trees): state is [nNodes][6][N] (mean,msgVar,logLik,descObs,descVar,
 * variance) pre-resampling; istate [nNodes][N] (Markov particle); logw [nNodes][N] raw
 * log weights before normalisation; anc [nNodes][N] (identity if not resampled). */
typedef struct {
  double* ess;
  double* relEss;
  double* logNormEstimate;
  double* logScaling;
  int32_t* resampled;
  double* state;
  int32_t* istate;
  double* logw;
  int32_t* anc;
  /* root population after the node step (what getRootPopulation() returns) */
  double* rootState;   /* [6][N] */
  int32_t* rootIState; /* [N] */
  double* rootWeights; /* [N] normalised weights (1/N if resampled) */
} orc_outputs;

/* injected[node] points at that node's uniform buffer (length N*S + N, S = slots per
 * particle: 2*maxBetaAttempts for binomial leaves, 0 for gauss-obs / Markov leaves,
 * 1 for internal nodes), or NULL when rngMode != INJECTED. */
int orc_run(const orc_options* opt, const orc_tree* tree, const double* const* injected,
            orc_outputs* out);

/* building blocks exposed for unit tests */
uint64_t orc_java_seed_scramble(int64_t seed);
int32_t orc_java_next(uint64_t* state, int bits);
double orc_java_next_double(uint64_t* state);
int32_t orc_java_next_int(uint64_t* state, int32_t bound);
int32_t orc_java_string_hash(const char* s);
int64_t orc_node_seed(int64_t masterRandomSeed, int32_t nodeHash);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_philox_uniform(int64_t seed, int32_t nodeId, int64_t idx);
double orc_elem_log(int mode, double x);
double orc_elem_exp(int mode, double x);
/* BrownianModelCalculator.combine (BrownianModelCalculator.java:18-41) on C children:
 * in: mean[C], var[C], ll[C]; out: 3 doubles */
void orc_brownian_combine(int elemMode, int C, const double* mean, const double* var,
                          const double* ll, double variance, double* out3);
/* impl-1 posterior summaries of ONE node (printNodeStatistics / printDeltaNodeStatistics /
 * sampleChildrenJointly, prototype/smc/DivideConquerMCAlgorithm.java:203-217,433-507) from its
 * RESAMPLED population: mean[N], msgVar[N], variance[N] (NULL for a leaf) and, for the deltas, the
 * children's messages aligned with the parent's particles, childMean/childVar [C][N].
 * Gaussian draws: java.util.Random.nextGaussian's polar method; attempt a of draw d uses the uniform
 * pair d*maxAtt + a of the node's summary stream — Philox counter word 3 = 1, or `injected`
 * (2 uniforms per pair) when not NULL. Draws per particle: 2 + C (0: node statistics, 1: joint root,
 * 2+c: child c). Mean / SD as commons-math3 DescriptiveStatistics (two-pass mean, bias-corrected SD).
 * out: meanMean, meanSD, varMean, varSD, then (deltaMean, deltaSD) per child.
 * PARITY UNPINNED: the reference holds no fixture for these; Normal.generate and logistic are
 * un-vendored bayonet code restated from recall (mean + sqrt(var)*nextGaussian; 1/(1+exp(-x))). */
int orc_node_summaries(int elemMode, int64_t seed, int32_t nodeId, int32_t N, int32_t maxAtt,
                       int32_t useTransform, const double* injected, const double* mean,
                       const double* msgVar, const double* variance, int32_t C, int32_t doDelta,
                       const double* childMean, const double* childVar, double* out);
/* same, and also the per-particle sample series the reference logs to its `samples` table (logSamples,
 * DivideConquerMCAlgorithm.java:508-519): series = [naturalParam][meanSample][varSample (variance != NULL)]
 * [delta of child 0..C-1 (doDelta)], each N long. */
int orc_node_sample_series(int elemMode, int64_t seed, int32_t nodeId, int32_t N, int32_t maxAtt,
                           int32_t useTransform, const double* injected, const double* mean,
                           const double* msgVar, const double* variance, int32_t C, int32_t doDelta,
                           const double* childMean, const double* childVar, double* out, double* series);
int32_t orc_slots_per_particle(const orc_options* opt, const orc_tree* tree, int32_t node);

#ifdef __cplusplus
}
#endif
#endif
