/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  (see dcsmc_oracle.h for the usage rule)
 *
 * Plain-C restatement of the reference's DC-SMC node step. All file:line citations are
 * into /root/reference/src/main/java/ unless stated otherwise.
 *
 * Un-vendored arithmetic restated here from its published/derived definition
 * (SURVEY.md §8c): ca.ubc.stat:bayonet:2.3.19 (ParticlePopulation, ResamplingScheme,
 * Multinomial.expNormalize, Exponential, SpecialFunctions) and org.apache.commons:
 * commons-math3 (BetaDistribution.sample = Cheng 1978 BB/BC as shipped in >= 3.4;
 * version unpinned by build.gradle). JDK java.util.Random per its javadoc.
 */
#define _GNU_SOURCE
#include "dcsmc_oracle.h"
#include "orc_elem.h"

#include <math.h>
#include <pthread.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

/* ------------------------------------------------------------------------------------
 * java.util.Random (JDK javadoc: LCG mod 2^48) and Java hashing
 * ---------------------------------------------------------------------------------- */
#define JMULT 0x5DEECE66DULL
#define JMASK ((1ULL << 48) - 1)

uint64_t orc_java_seed_scramble(int64_t seed) { return ((uint64_t)seed ^ JMULT) & JMASK; }

int32_t orc_java_next(uint64_t* state, int bits) {
  *state = (*state * JMULT + 0xBULL) & JMASK;
  return (int32_t)((int64_t)(*state) >> (48 - bits)); /* state < 2^48, so arithmetic == logical */
}

double orc_java_next_double(uint64_t* state) {
  int64_t hi = (int64_t)orc_java_next(state, 26);
  int64_t lo = (int64_t)orc_java_next(state, 27);
  return (double)((hi << 27) + lo) * 0x1.0p-53;
}

int32_t orc_java_next_int(uint64_t* state, int32_t bound) {
  int32_t r = orc_java_next(state, 31);
  int32_t m = bound - 1;
  if ((bound & m) == 0) {
    r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
  } else {
    int32_t u = r;
    /* for (int u = r; u - (r = u % bound) + m < 0; u = next(31)); — int32 wraparound */
    for (;;) {
      r = u % bound;
      int32_t t = (int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m);
      if (t >= 0) break;
      u = orc_java_next(state, 31);
    }
  }
  return r;
}

/* String.hashCode: s[0]*31^(n-1) + ... (int32 wrap) */
int32_t orc_java_string_hash(const char* s) {
  uint32_t h = 0;
  for (; *s; ++s) h = 31u * h + (uint32_t)(unsigned char)*s;
  return (int32_t)h;
}

/* DCRecursionTask.getRandom, dc/DCRecursionTask.java:98-106: long arithmetic, int hash
 * sign-extended. */
int64_t orc_node_seed(int64_t masterRandomSeed, int32_t nodeHash) {
  uint64_t seed = 1;
  seed = 31ULL * seed + (uint64_t)masterRandomSeed;
  seed = 31ULL * seed + (uint64_t)(int64_t)nodeHash;
  return (int64_t)seed;
}

/* ------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11 — Random123), the counter-based generator the
 * device uses. Restated from the paper; pinned by the Random123 known-answer vectors in
 * tests/test_oracle_units.py.
 * ---------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Uniform number idx of node nodeId: counter = (idx>>1, nodeId, 0), key = seed; each
 * Philox block yields two 53-bit doubles built like java.util.Random.nextDouble
 * (26 high bits, 27 low bits). */
double orc_philox_uniform(int64_t seed, int32_t nodeId, int64_t idx) {
  uint64_t blk = (uint64_t)idx >> 1;
  uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)nodeId, 0u};
  uint32_t key[2] = {(uint32_t)(uint64_t)seed, (uint32_t)((uint64_t)seed >> 32)};
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  uint32_t a = (idx & 1) ? o[2] : o[0];
  uint32_t b = (idx & 1) ? o[3] : o[1];
  uint64_t bits = ((uint64_t)(a >> 6) << 27) | (uint64_t)(b >> 5);
  return (double)bits * 0x1.0p-53;
}

/* ------------------------------------------------------------------------------------
 * RNG abstraction: JAVA = one sequential java.util.Random per node (the reference);
 * PHILOX / INJECTED = slot-addressed streams (uniform number `cursor` of this node).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int mode;
  uint64_t jstate;
  int64_t seed;
  int32_t nodeId;
  const double* inj;
  int64_t cursor;
} orc_rng;

static inline void rng_seek(orc_rng* r, int64_t slot) { r->cursor = slot; }

static inline double rng_next_double(orc_rng* r) {
  switch (r->mode) {
    case ORC_RNG_JAVA: return orc_java_next_double(&r->jstate);
    case ORC_RNG_PHILOX: return orc_philox_uniform(r->seed, r->nodeId, r->cursor++);
    default: return r->inj[r->cursor++];
  }
}

static inline int32_t rng_next_int(orc_rng* r, int32_t bound) {
  if (r->mode == ORC_RNG_JAVA) return orc_java_next_int(&r->jstate, bound);
  /* slot modes: one uniform, floor(u*bound) */
  int32_t v = (int32_t)(rng_next_double(r) * (double)bound);
  return v >= bound ? bound - 1 : v;
}

/* ------------------------------------------------------------------------------------
 * elementary functions
 * ---------------------------------------------------------------------------------- */
double orc_elem_log(int mode, double x) { return mode == ORC_ELEM_FDLIBM ? orc_fd_log(x) : log(x); }
double orc_elem_exp(int mode, double x) { return mode == ORC_ELEM_FDLIBM ? orc_fd_exp(x) : exp(x); }
#define LOG(x) orc_elem_log(ctx->opt->elemMode, (x))
#define EXP(x) orc_elem_exp(ctx->opt->elemMode, (x))

/* ------------------------------------------------------------------------------------
 * particles and populations (DivideConquerMCAlgorithm.java:150-194 Particle;
 * bayonet ParticlePopulation)
 * ---------------------------------------------------------------------------------- */
typedef struct {
  double mean, msgVar, logLik, descObs, descVar, variance;
  int32_t state; /* Markov toy particle (Doc.java: particles are Integers) */
} particle;

typedef struct {
  particle* p;
  double* w; /* normalised weights; NULL => all equal (after resampling) */
  double logScaling;
  int32_t n;
} population;

static void pop_free(population* q) {
  if (!q) return;
  free(q->p);
  free(q->w);
  free(q);
}

typedef struct {
  const orc_options* opt;
  const orc_tree* tree;
  const double* const* injected;
  orc_outputs* out;
  population** cache; /* populations of subtree roots computed by worker threads */
  uint64_t globalJava; /* impl-1: the single Random threaded through the recursion */
  double* impl1LogZ;   /* impl-1: logZEstimate per node */
  int error;
} orc_ctx;

static int n_children(const orc_tree* t, int32_t node) {
  return t->childOffsets[node + 1] - t->childOffsets[node];
}

int32_t orc_slots_per_particle(const orc_options* opt, const orc_tree* tree, int32_t node) {
  int C = n_children(tree, node);
  if (opt->model == ORC_MODEL_MARKOV) return C == 0 ? 0 : 1;
  if (C > 0) return 1;
  if (tree->leafKind && tree->leafKind[node] == ORC_LEAF_GAUSS_OBS) return 0;
  return 2 * opt->maxBetaAttempts;
}

/* ------------------------------------------------------------------------------------
 * BrownianModelCalculator (prototype/smc/BrownianModelCalculator.java)
 * ---------------------------------------------------------------------------------- */
static inline double log_normal_density(int elemMode, double x, double mean, double var) {
  /* :137-139  -0.5*(x-mean)*(x-mean)/var -0.5*Math.log(2*Math.PI * var) */
  return -0.5 * (x - mean) * (x - mean) / var - 0.5 * orc_elem_log(elemMode, 2 * M_PI * var);
}

/* calculate(l1,l2,v1,v2,peek=false), :86-115, nsites == 1 */
static inline void brownian_calculate(int elemMode, double mean1, double var1, double ll1,
                                      double mean2, double var2, double ll2, double v1,
                                      double v2, double* m, double* s, double* l) {
  double logl = 0;
  const double var = 1 / (var1 + v1) + 1 / (var2 + v2);
  const double newMessageVariance = 1 / var;
  const double message = ((mean1) / (var1 + v1) + (mean2) / (var2 + v2)) / var;
  const double cur = log_normal_density(elemMode, mean1 - mean2, 0, (v1 + var1 + v2 + var2));
  logl += cur;
  logl += ll1 + ll2;
  *m = message;
  *s = newMessageVariance;
  *l = logl;
}

/* ------------------------------------------------------------------------------------
 * impl-1 posterior summaries (DivideConquerMCAlgorithm.java:203-217, 433-507)
 * ---------------------------------------------------------------------------------- */
/* uniform pair `pair` of the node's summary stream */
static void summary_pair(int64_t seed, int32_t nodeId, const double* injected, uint64_t pair,
                         double* u0, double* u1) {
  if (injected) {
    *u0 = injected[2 * pair];
    *u1 = injected[2 * pair + 1];
    return;
  }
  uint32_t ctr[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)nodeId, 1u};
  uint32_t key[2] = {(uint32_t)(uint64_t)seed, (uint32_t)((uint64_t)seed >> 32)};
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  *u0 = (double)(((uint64_t)(o[0] >> 6) << 27) | (uint64_t)(o[1] >> 5)) * 0x1.0p-53;
  *u1 = (double)(((uint64_t)(o[2] >> 6) << 27) | (uint64_t)(o[3] >> 5)) * 0x1.0p-53;
}

/* java.util.Random.nextGaussian (JDK spec), without the cached second value */
static double summary_gauss(int elemMode, int64_t seed, int32_t nodeId, const double* injected,
                            uint64_t draw, int maxAtt) {
  double v1 = 0.0, s = 0.5;
  for (int a = 0; a < maxAtt; ++a) {
    double u0, u1;
    summary_pair(seed, nodeId, injected, draw * (uint64_t)maxAtt + (uint64_t)a, &u0, &u1);
    v1 = 2 * u0 - 1;
    const double v2 = 2 * u1 - 1;
    s = v1 * v1 + v2 * v2;
    if (s < 1 && s != 0) break;
  }
  const double multiplier = sqrt(-2 * orc_elem_log(elemMode, s) / s);
  return v1 * multiplier;
}

static double summary_transform(int elemMode, int useTransform, double x) {
  return useTransform ? 1.0 / (1.0 + orc_elem_exp(elemMode, -x)) : x; /* :610-616 */
}

/* commons-math3 DescriptiveStatistics.getMean / getStandardDeviation */
static void descriptive(const double* x, int n, double* mean, double* sd) {
  double sum = 0;
  for (int i = 0; i < n; ++i) sum += x[i];
  const double xbar = sum / n;
  double corr = 0;
  for (int i = 0; i < n; ++i) corr += x[i] - xbar;
  const double m = xbar + corr / n;
  double accum = 0, accum2 = 0;
  for (int i = 0; i < n; ++i) {
    const double dev = x[i] - m;
    accum += dev * dev;
    accum2 += dev;
  }
  *mean = m;
  *sd = n > 1 ? sqrt((accum - (accum2 * accum2 / n)) / (n - 1.0)) : 0.0;
}

/* series (optional): the per-particle sample series the reference logs to its `samples` table
 * (logSamples, DivideConquerMCAlgorithm.java:508-519): [naturalParam][meanSample][varSample (internal nodes)]
 * [delta of child 0 .. C-1 (if doDelta)], each N long. */
static int node_summaries(int elemMode, int64_t seed, int32_t nodeId, int32_t N, int32_t maxAtt,
                          int32_t useTransform, const double* injected, const double* mean,
                          const double* msgVar, const double* variance, int32_t C, int32_t doDelta,
                          const double* childMean, const double* childVar, double* out, double* series) {
  double* buf = (double*)malloc(sizeof(double) * (size_t)N);
  if (!buf) return 1;
  const uint64_t G = 2 + (uint64_t)C;
  int row = 0;
  /* printNodeStatistics :476-498 */
  for (int i = 0; i < N; ++i) {
    const double natural = summary_gauss(elemMode, seed, nodeId, injected, (uint64_t)i * G, maxAtt) * sqrt(msgVar[i]) + mean[i];
    buf[i] = summary_transform(elemMode, useTransform, natural);
    if (series) {
      series[i] = natural;
      series[(size_t)N + i] = buf[i];
    }
  }
  row = 2;
  descriptive(buf, N, &out[0], &out[1]);
  out[2] = out[3] = 0.0;
  if (variance) {
    descriptive(variance, N, &out[2], &out[3]);
    if (series) memcpy(series + (size_t)row * N, variance, sizeof(double) * (size_t)N);
    row++;
  }
  /* printDeltaNodeStatistics :433-469 with sampleChildrenJointly :203-217 */
  for (int c = 0; doDelta && c < C; ++c) {
    for (int i = 0; i < N; ++i) {
      const double root = summary_gauss(elemMode, seed, nodeId, injected, (uint64_t)i * G + 1, maxAtt) * sqrt(msgVar[i]) + mean[i];
      double um, us, ul;
      brownian_calculate(elemMode, mean[i], msgVar[i], 0.0, childMean[(size_t)c * N + i], childVar[(size_t)c * N + i], 0.0,
                         variance ? variance[i] : NAN, 0.0, &um, &us, &ul);
      const double child = summary_gauss(elemMode, seed, nodeId, injected, (uint64_t)i * G + 2 + c, maxAtt) * sqrt(us) + um;
      buf[i] = summary_transform(elemMode, useTransform, child) - summary_transform(elemMode, useTransform, root);
    }
    descriptive(buf, N, &out[4 + 2 * c], &out[5 + 2 * c]);
    if (series) memcpy(series + (size_t)(row + c) * N, buf, sizeof(double) * (size_t)N);
  }
  free(buf);
  return 0;
}

int orc_node_summaries(int elemMode, int64_t seed, int32_t nodeId, int32_t N, int32_t maxAtt,
                       int32_t useTransform, const double* injected, const double* mean,
                       const double* msgVar, const double* variance, int32_t C, int32_t doDelta,
                       const double* childMean, const double* childVar, double* out) {
  return node_summaries(elemMode, seed, nodeId, N, maxAtt, useTransform, injected, mean, msgVar, variance, C, doDelta,
                        childMean, childVar, out, NULL);
}

int orc_node_sample_series(int elemMode, int64_t seed, int32_t nodeId, int32_t N, int32_t maxAtt,
                           int32_t useTransform, const double* injected, const double* mean,
                           const double* msgVar, const double* variance, int32_t C, int32_t doDelta,
                           const double* childMean, const double* childVar, double* out, double* series) {
  return node_summaries(elemMode, seed, nodeId, N, maxAtt, useTransform, injected, mean, msgVar, variance, C, doDelta,
                        childMean, childVar, out, series);
}

/* combine(List children, double variance), :18-41 */
void orc_brownian_combine(int elemMode, int C, const double* mean, const double* var,
                          const double* ll, double variance, double* out3) {
  if (C == 1) { /* :23-27 */
    out3[0] = mean[0];
    out3[1] = var[0] + variance;
    out3[2] = ll[0];
    return;
  }
  double m, s, l;
  brownian_calculate(elemMode, mean[0], var[0], ll[0], mean[1], var[1], ll[1], variance,
                     variance, &m, &s, &l); /* :34 */
  for (int c = 2; c < C; ++c)               /* :35-39 */
    brownian_calculate(elemMode, m, s, l, mean[c], var[c], ll[c], 0.0, variance, &m, &s, &l);
  out3[0] = m;
  out3[1] = s;
  out3[2] = l;
}

/* ------------------------------------------------------------------------------------
 * Proposals
 * ---------------------------------------------------------------------------------- */

/* Per-leaf constants built once per node (MultiLevelProposalFactory.build :40-49 +
 * commons-math3 BetaDistribution's Cheng sampler set-up). */
typedef struct {
  double a0;       /* alpha = 1 + nSuccesses */
  double a, b;     /* BB: a=min,b=max ; BC: a=max,b=min (argument order of commons) */
  int useBB;
  double alpha, beta, gamma, logAlpha; /* BB */
  double delta, k1, k2;                /* BC */
  double logBinom;                     /* SpecialFunctions.logBinomial(n,s) */
  int32_t n, s;
} leaf_consts;

static void leaf_setup(const orc_ctx* ctx, int32_t node, leaf_consts* lc) {
  const int32_t n = ctx->tree->nTrials[node], s = ctx->tree->nSuccesses[node];
  const double shapeA = 1 + s, shapeB = 1 + (n - s);
  lc->n = n;
  lc->s = s;
  lc->a0 = shapeA;
  const double mn = fmin(shapeA, shapeB), mx = fmax(shapeA, shapeB);
  lc->useBB = mn > 1;
  if (lc->useBB) {
    lc->a = mn;
    lc->b = mx;
    lc->alpha = lc->a + lc->b;
    lc->beta = sqrt((lc->alpha - 2.) / (2. * lc->a * lc->b - lc->alpha));
    lc->gamma = lc->a + 1. / lc->beta;
  } else {
    lc->a = mx;
    lc->b = mn;
    lc->alpha = lc->a + lc->b;
    lc->beta = 1. / lc->b;
    lc->delta = 1. + lc->a - lc->b;
    lc->k1 = lc->delta * (0.0138889 + 0.0416667 * lc->b) / (lc->a * lc->beta - 0.777778);
    lc->k2 = 0.25 + (0.5 + 0.25 / lc->delta) * lc->b;
  }
  lc->logAlpha = LOG(lc->alpha);
  lc->logBinom = lgamma((double)n + 1.0) - lgamma((double)s + 1.0) - lgamma((double)(n - s) + 1.0);
}

/* commons-math3 (>= 3.4) BetaDistribution.sample(): Cheng (1978) BB / BC. Each attempt
 * consumes exactly two uniforms. log1p(-u1) is evaluated as log(1-u1) (one elementary
 * function for the whole engine). The attempt cap (maxBetaAttempts) exists so that slot-
 * addressed streams have a fixed per-particle budget; if it is hit the last attempt is
 * accepted. */
static double beta_sample(const orc_ctx* ctx, const leaf_consts* lc, orc_rng* rng) {
  const int maxAtt = ctx->opt->maxBetaAttempts;
  double w = 0;
  if (lc->useBB) {
    for (int att = 0; att < maxAtt; ++att) {
      const double u1 = rng_next_double(rng);
      const double u2 = rng_next_double(rng);
      const double v = lc->beta * (LOG(u1) - LOG(1.0 - u1));
      w = lc->a * EXP(v);
      const double z = u1 * u1 * u2;
      const double r = lc->gamma * v - 1.3862944;
      const double s = lc->a + r - w;
      if (s + 2.609438 >= 5 * z) break;
      const double t = LOG(z);
      if (s >= t) break;
      if (!(r + lc->alpha * (lc->logAlpha - LOG(lc->b + w)) < t)) break;
    }
  } else {
    for (int att = 0; att < maxAtt; ++att) {
      const double u1 = rng_next_double(rng);
      const double u2 = rng_next_double(rng);
      /* v, w are evaluated up front (commons evaluates them lazily; same values) */
      const double v = lc->beta * (LOG(u1) - LOG(1.0 - u1));
      w = lc->a * EXP(v);
      const double y = u1 * u2;
      const double z = u1 * y;
      if (u1 < 0.5) {
        if (0.25 * u2 + z - y >= lc->k1) continue;
      } else {
        if (z <= 0.25) break;
        if (z >= lc->k2) continue;
      }
      if (lc->alpha * (lc->logAlpha - LOG(lc->b + w) + v) - 1.3862944 >= LOG(z)) break;
    }
  }
  w = fmin(w, 1.7976931348623157e308);
  return (lc->a == lc->a0) ? w / (lc->b + w) : lc->b / (lc->b + w);
}

/* MultiLevelLeafProposal.propose, prototype/adaptor/MultiLevelLeafProposal.java:35-43
 * (== impl-1 _leafParticleApproximation, DivideConquerMCAlgorithm.java:557-591) */
static double propose_binomial_leaf(const orc_ctx* ctx, const leaf_consts* lc, orc_rng* rng,
                                    particle* out) {
  const double proposed = beta_sample(ctx, lc, rng);
  /* logBinomialPr, DivideConquerMCAlgorithm.java:593-600 ; logit = log(p) - log(1-p) */
  const double lp = LOG(proposed), lq = LOG(1.0 - proposed);
  const double logPi = lc->logBinom + lc->s * lp + (lc->n - lc->s) * lq;
  const double transformed = lp - lq;
  /* Particle(leaf, node, logPi), DivideConquerMCAlgorithm.java:181-184; observation():
   * BrownianModelCalculator.java:70-84 */
  out->mean = transformed;
  out->msgVar = 0.0;
  out->logLik = 0.0;
  out->descObs = logPi;
  out->descVar = 0.0;
  out->variance = NAN;
  out->state = 0;
  return 0.0; /* logWeight = 0.0, MultiLevelLeafProposal.java:40 */
}

static double propose_gauss_leaf(double y, particle* out) {
  out->mean = y;
  out->msgVar = 0.0;
  out->logLik = 0.0;
  out->descObs = 0.0;
  out->descVar = 0.0;
  out->variance = NAN;
  out->state = 0;
  return 0.0;
}

/* MultiLevelInternalProposal.propose, prototype/adaptor/MultiLevelInternalProposal.java:34-57
 * (== impl-1 loop DivideConquerMCAlgorithm.java:363-400) */
static double propose_internal(const orc_ctx* ctx, orc_rng* rng, population** kids, int C,
                               int32_t i, double* mean, double* var, double* ll,
                               particle* out) {
  const double rate = ctx->opt->variancePriorRate;
  /* bayonet Exponential.generate = -log(U)/rate ; logDensity = log(rate) - rate*x */
  const double variance = -LOG(rng_next_double(rng)) / rate;
  double descParticleObsLogl = 0.0;
  double descVar = LOG(rate) - rate * variance;
  for (int c = 0; c < C; ++c) {
    const particle* cp = &kids[c]->p[i];
    mean[c] = cp->mean;
    var[c] = cp->msgVar;
    ll[c] = cp->logLik;
    descParticleObsLogl += cp->descObs;
    descVar += cp->descVar;
  }
  double comb[3];
  orc_brownian_combine(ctx->opt->elemMode, C, mean, var, ll, variance, comb);
  double logWeight = comb[2];
  for (int c = 0; c < C; ++c) logWeight = logWeight - ll[c];
  out->mean = comb[0];
  out->msgVar = comb[1];
  out->logLik = comb[2];
  out->descObs = descParticleObsLogl;
  out->descVar = descVar;
  out->variance = variance;
  out->state = 0;
  return logWeight;
}

/* Doc.markovChainNaiveProposal, src/test/java/dc/Doc.java:162-195 and the factory
 * :206-225 */
static double propose_markov(const orc_ctx* ctx, orc_rng* rng, population** kids, int C,
                             int32_t i, particle* out) {
  const int nStates = ctx->opt->nStates;
  const double* T = ctx->opt->transition;
  const double* prior = ctx->opt->prior;
  particle z = {0};
  z.variance = NAN;
  if (C == 0) {
    *out = z;
    return LOG(prior[0]); /* Pair.of(Math.log(prior.get(0, 0)), 0) */
  }
  const int32_t proposal = rng_next_int(rng, nStates);
  double weightUpdate = nStates;
  weightUpdate *= prior[proposal];
  for (int c = 0; c < C; ++c) {
    weightUpdate *= T[proposal * nStates + kids[c]->p[i].state];
    weightUpdate /= prior[proposal];
  }
  z.state = proposal;
  *out = z;
  return LOG(weightUpdate);
}

/* ------------------------------------------------------------------------------------
 * Exact (order-independent) sums — the device contract ("SUM_EXACT"): every e in [0,1]
 * is split into two 38-bit fixed-point limbs, a = trunc(e*2^38), b = trunc((e*2^38-a)*2^38);
 * integer sums are exact for N <= 2^24, so prefix sums do not depend on the order or
 * blocking of the additions, and the double image
 *   val(A,B) = fl( fl(A)*2^-38 + fl(B)*2^-76 )
 * is monotone in the prefix.
 * ---------------------------------------------------------------------------------- */
#define TWO38 274877906944.0
static inline void exact_split(double e, uint64_t* a, uint64_t* b) {
  const double x = e * TWO38;
  const uint64_t ai = (uint64_t)x; /* truncation */
  const double r = x - (double)ai; /* exact */
  *a = ai;
  *b = (uint64_t)(r * TWO38);
}
static inline double exact_val(uint64_t A, uint64_t B) {
  return (double)A * 0x1.0p-38 + (double)B * 0x1.0p-76;
}

static int cmp_double(const void* x, const void* y) {
  const double a = *(const double*)x, b = *(const double*)y;
  return (a > b) - (a < b);
}

/* ------------------------------------------------------------------------------------
 * The node step: DCRecursion.dcRecurse / dcPropose, dc/DCRecursion.java:19-74
 * ---------------------------------------------------------------------------------- */
static population* node_step(orc_ctx* ctx, int32_t node, population** kids, int C) {
  const orc_options* opt = ctx->opt;
  const int32_t N = opt->nParticles;
  const int impl1 = (opt->impl == 1);
  const int isLeaf = (C == 0);
  const int64_t S = orc_slots_per_particle(opt, ctx->tree, node);

  /* DCRecursionTask.getRandom :98-106 (impl-2) ; impl-1 threads one Random through */
  orc_rng rng;
  rng.mode = opt->rngMode;
  rng.seed = opt->masterRandomSeed;
  rng.nodeId = node;
  rng.inj = ctx->injected ? ctx->injected[node] : NULL;
  rng.cursor = 0;
  rng.jstate = impl1 ? ctx->globalJava
                     : orc_java_seed_scramble(orc_node_seed(opt->masterRandomSeed,
                                                            ctx->tree->javaHash[node]));

  for (int c = 0; c < C; ++c) /* DCRecursion.java:48-50 */
    if (kids[c]->n != N) { ctx->error = 2; return NULL; }

  particle* parts = (particle*)malloc(sizeof(particle) * (size_t)N);
  double* logW = (double*)malloc(sizeof(double) * (size_t)N);
  double* mean = (double*)malloc(sizeof(double) * (size_t)(C + 1) * 3);
  double *var = mean + (C + 1), *ll = var + (C + 1);

  leaf_consts lc = {0};
  const int binomLeaf = isLeaf && opt->model == ORC_MODEL_MULTILEVEL &&
                        !(ctx->tree->leafKind && ctx->tree->leafKind[node] == ORC_LEAF_GAUSS_OBS);
  if (binomLeaf) leaf_setup(ctx, node, &lc);

  /* DCRecursion.java:52-66 */
  for (int32_t i = 0; i < N; ++i) {
    double childrenWeightProduct = 1.0;
    if (!impl1) /* impl-1 assumes resampling at each generation (DivideConquerMCAlgorithm.java:377-378) */
      for (int c = 0; c < C; ++c)
        childrenWeightProduct *= kids[c]->w ? kids[c]->w[i] : 1.0 / (double)N;
    rng_seek(&rng, (int64_t)i * S);
    double lw;
    if (opt->model == ORC_MODEL_MARKOV)
      lw = propose_markov(ctx, &rng, kids, C, i, &parts[i]);
    else if (binomLeaf)
      lw = propose_binomial_leaf(ctx, &lc, &rng, &parts[i]);
    else if (isLeaf)
      lw = propose_gauss_leaf(ctx->tree->leafObs[node], &parts[i]);
    else
      lw = propose_internal(ctx, &rng, kids, C, i, mean, var, ll, &parts[i]);
    double logProduct = LOG(childrenWeightProduct); /* DCRecursion.java:63 */
    if (!impl1 && C > 0 && childrenWeightProduct == 0.0) {
      /* EXTENSION (DESIGN.md "Known deviations"): the linear-space product underflowed to 0 — with
       * resampled children that is N^-C < 4.9e-324, e.g. 54 children at N = 1e6 — and the reference
       * returns log(0) = -inf for every particle here. The engine evaluates sum_c log W_c instead
       * (left to right); a particle with a genuinely zero child weight still gets -inf. */
      logProduct = 0.0;
      for (int c = 0; c < C; ++c) logProduct += LOG(kids[c]->w ? kids[c]->w[i] : 1.0 / (double)N);
    }
    logW[i] = impl1 ? lw : logProduct + lw;
  }
  free(mean);

  /* DCRecursion.java:68-71: DoubleStream.sum() (compensated in the JDK); the exact-sum
   * contract uses the plain left-to-right sum. */
  double base = 0.0;
  if (opt->sumMode == ORC_SUM_JAVA) {
    double sum = 0.0, comp = 0.0;
    for (int c = 0; c < C; ++c) {
      const double tmp = kids[c]->logScaling - comp;
      const double velvel = sum + tmp;
      comp = (velvel - sum) - tmp;
      sum = velvel;
    }
    base = sum - comp;
  } else {
    for (int c = 0; c < C; ++c) base += kids[c]->logScaling;
  }

  /* dumps (pre-normalisation) */
  orc_outputs* out = ctx->out;
  if (out && out->logw) memcpy(out->logw + (size_t)node * N, logW, sizeof(double) * (size_t)N);
  if (out && out->state)
    for (int32_t i = 0; i < N; ++i) {
      double* s = out->state + (size_t)node * 6 * N;
      s[0 * (size_t)N + i] = parts[i].mean;
      s[1 * (size_t)N + i] = parts[i].msgVar;
      s[2 * (size_t)N + i] = parts[i].logLik;
      s[3 * (size_t)N + i] = parts[i].descObs;
      s[4 * (size_t)N + i] = parts[i].descVar;
      s[5 * (size_t)N + i] = parts[i].variance;
    }
  if (out && out->istate)
    for (int32_t i = 0; i < N; ++i) out->istate[(size_t)node * N + i] = parts[i].state;

  /* ParticlePopulation.buildDestructivelyFromLogWeights -> Multinomial.expNormalize
   * (bayonet) */
  double m = -INFINITY;
  for (int32_t i = 0; i < N; ++i) m = fmax(m, logW[i]);
  double* W = logW; /* destructive, like the reference */
  double Ssum, ess, logNorm;
  double* cdf = NULL; /* exact mode keeps the unnormalised CDF */
  if (opt->sumMode == ORC_SUM_JAVA) {
    for (int32_t i = 0; i < N; ++i) W[i] = EXP(logW[i] - m);
    Ssum = 0.0;
    for (int32_t i = 0; i < N; ++i) Ssum += W[i];
    for (int32_t i = 0; i < N; ++i) W[i] /= Ssum;
    logNorm = m + LOG(Ssum);
    /* getESS = 1/sum w^2 (cf. SMCUtils.ess :16-22). The reference's recorded root ESS
     * 548622.6672945316 (Doc.java:282) is reproduced by a compensated sum of squares as
     * java.util.stream.DoubleStream.sum() performs (Kahan), and NOT by the naive loop
     * (...2943507), so the un-vendored bayonet getESS is restated with the stream sum. */
    double q = 0.0, comp = 0.0;
    for (int32_t i = 0; i < N; ++i) {
      const double tmp = W[i] * W[i] - comp;
      const double velvel = q + tmp;
      comp = (velvel - q) - tmp;
      q = velvel;
    }
    ess = 1.0 / (q - comp);
  } else {
    cdf = (double*)malloc(sizeof(double) * (size_t)N);
    uint64_t A = 0, B = 0, QA = 0, QB = 0;
    for (int32_t i = 0; i < N; ++i) {
      const double e = EXP(logW[i] - m);
      uint64_t a, b;
      exact_split(e, &a, &b);
      A += a;
      B += b;
      cdf[i] = exact_val(A, B);
      exact_split(e * e, &a, &b);
      QA += a;
      QB += b;
      W[i] = e;
    }
    Ssum = exact_val(A, B);
    const double Q = exact_val(QA, QB);
    for (int32_t i = 0; i < N; ++i) W[i] = W[i] / Ssum;
    logNorm = m + LOG(Ssum);
    ess = (Ssum * Ssum) / Q;
  }
  const double relEss = ess / (double)N;

  population* res = (population*)malloc(sizeof(population));
  res->p = parts;
  res->w = W;
  res->n = N;
  double logZ;
  if (impl1) { /* DivideConquerMCAlgorithm.java:360-361, 404-406 */
    logZ = 0.0;
    for (int c = 0; c < C; ++c) logZ += ctx->impl1LogZ[ctx->tree->children[ctx->tree->childOffsets[node] + c]];
    logZ += logNorm - LOG((double)N);
    ctx->impl1LogZ[node] = logZ;
    res->logScaling = logZ + LOG((double)N);
  } else {
    res->logScaling = base + logNorm;
    logZ = res->logScaling - LOG((double)N); /* logNormEstimate(), Appendix D fact 3 */
  }

  /* processors see the population before resampling (DCRecursion.java:27-28;
   * DefaultProcessorFactory.java:41-54) */
  if (out) {
    if (out->ess) out->ess[node] = ess;
    if (out->relEss) out->relEss[node] = relEss;
    if (out->logNormEstimate) out->logNormEstimate[node] = logZ;
    if (out->logScaling) out->logScaling[node] = res->logScaling;
  }

  /* DCRecursion.java:29-31 ; impl-1 always resamples (:419) */
  const int doResample = impl1 ? 1 : (relEss < opt->relativeEssThreshold);
  if (out && out->resampled) out->resampled[node] = doResample;
  int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  if (doResample) {
    double* darts = (double*)malloc(sizeof(double) * (size_t)N);
    rng_seek(&rng, (int64_t)N * S);
    const int scheme = impl1 ? ORC_SCHEME_MULTINOMIAL : opt->resamplingScheme;
    if (scheme == ORC_SCHEME_MULTINOMIAL) { /* SMCUtils.java:26-29 ; Appendix D fact 5 */
      for (int32_t j = 0; j < N; ++j) darts[j] = rng_next_double(&rng);
      qsort(darts, (size_t)N, sizeof(double), cmp_double);
    } else { /* bayonet STRATIFIED [recalled]: (j + U_j)/N */
      for (int32_t j = 0; j < N; ++j) darts[j] = ((double)j + rng_next_double(&rng)) / (double)N;
    }
    if (opt->sumMode == ORC_SUM_JAVA) { /* sweep, SMCUtils.java:31-49 */
      double sum = 0.0;
      int32_t nxt = 0;
      for (int32_t i = 0; i < N; ++i) {
        const double right = sum + W[i];
        while (nxt < N && darts[nxt] < right) anc[nxt++] = i;
        sum = right;
      }
      while (nxt < N) anc[nxt++] = N - 1; /* the reference throws here (:53) */
    } else {
      int32_t i = 0;
      for (int32_t j = 0; j < N; ++j) {
        const double t = darts[j] * Ssum;
        while (i < N - 1 && !(t < cdf[i])) ++i;
        anc[j] = i;
      }
    }
    free(darts);
    particle* np = (particle*)malloc(sizeof(particle) * (size_t)N);
    for (int32_t j = 0; j < N; ++j) np[j] = parts[anc[j]];
    free(parts);
    res->p = np;
    free(res->w);
    res->w = NULL; /* equally weighted, same logScaling */
  } else {
    for (int32_t j = 0; j < N; ++j) anc[j] = j;
  }
  if (out && out->anc) memcpy(out->anc + (size_t)node * N, anc, sizeof(int32_t) * (size_t)N);
  free(anc);
  free(cdf);
  if (impl1) ctx->globalJava = rng.jstate;
  return res;
}

/* DCRecursionTask.run(node, false), dc/DCRecursionTask.java:48-75 (serial DFS) */
static population* run_node(orc_ctx* ctx, int32_t node) {
  if (ctx->cache && ctx->cache[node]) {
    population* q = ctx->cache[node];
    ctx->cache[node] = NULL;
    return q;
  }
  const int C = n_children(ctx->tree, node);
  population** kids = (population**)calloc((size_t)(C > 0 ? C : 1), sizeof(population*));
  for (int c = 0; c < C; ++c) {
    kids[c] = run_node(ctx, ctx->tree->children[ctx->tree->childOffsets[node] + c]);
    if (!kids[c]) { ctx->error = ctx->error ? ctx->error : 3; }
  }
  population* res = ctx->error ? NULL : node_step(ctx, node, kids, C);
  for (int c = 0; c < C; ++c) pop_free(kids[c]); /* children are consumed (:113) */
  free(kids);
  return res;
}

/* DistributedDC.populateInitialTasks, dc/DistributedDC.java:165-175 */
static void initial_tasks(const orc_tree* t, int32_t node, int depth, int32_t* tasks, int* n) {
  if (n_children(t, node) == 0 || depth == 0) {
    tasks[(*n)++] = node;
    return;
  }
  for (int c = t->childOffsets[node]; c < t->childOffsets[node + 1]; ++c)
    initial_tasks(t, t->children[c], depth - 1, tasks, n);
}

/* The reference's executor (DistributedDC.java:221-224 + DCRecursionTask.prepareNextTask :77-93): the
 * initial tasks run whole subtrees serially; a node above the cut becomes a task the moment its last child
 * population is in the cluster map, and any free thread takes it. Nothing inside a node step is parallel. */
typedef struct {
  orc_ctx* ctx;
  int32_t* ready;    /* FIFO of runnable nodes */
  int head, tail;
  int32_t* pending;  /* per node above the cut: children still missing */
  char* isTask;      /* per node: an initial task (serial recursion through its subtree) */
  int32_t* parent;
  int remaining;     /* nodes still to finish (initial tasks + nodes above the cut) */
  pthread_mutex_t mu;
  pthread_cond_t cv;
} work_queue;

static void* worker(void* arg) {
  work_queue* q = (work_queue*)arg;
  for (;;) {
    pthread_mutex_lock(&q->mu);
    while (q->head == q->tail && q->remaining > 0) pthread_cond_wait(&q->cv, &q->mu);
    if (q->remaining <= 0 && q->head == q->tail) {
      pthread_mutex_unlock(&q->mu);
      break;
    }
    const int32_t node = q->ready[q->head++];
    pthread_mutex_unlock(&q->mu);
    orc_ctx local = *q->ctx; /* error flag is per-thread; outputs are disjoint per node */
    population* res;
    if (q->isTask[node]) {
      local.cache = NULL; /* DCRecursionTask.run(node, false): recurse through the whole subtree */
      res = run_node(&local, node);
    } else {
      res = run_node(&local, node); /* children are taken from the cache (getChildrenPopulationsFromCluster) */
    }
    pthread_mutex_lock(&q->mu);
    q->ctx->cache[node] = res;
    if (local.error) q->ctx->error = local.error;
    q->remaining--;
    const int32_t par = q->parent[node];
    if (par >= 0 && --q->pending[par] == 0) q->ready[q->tail++] = par;
    pthread_cond_broadcast(&q->cv);
    pthread_mutex_unlock(&q->mu);
  }
  return NULL;
}

int orc_run(const orc_options* opt, const orc_tree* tree, const double* const* injected,
            orc_outputs* out) {
  orc_ctx ctx;
  ctx.opt = opt;
  ctx.tree = tree;
  ctx.injected = injected;
  ctx.out = out;
  ctx.cache = NULL;
  ctx.error = 0;
  ctx.globalJava = orc_java_seed_scramble(opt->masterRandomSeed);
  ctx.impl1LogZ = NULL;
  if (opt->rngMode == ORC_RNG_INJECTED && !injected) return 4;
  if (opt->impl == 1) ctx.impl1LogZ = (double*)calloc((size_t)tree->nNodes, sizeof(double));

  if (opt->impl != 1 && opt->nThreads > 1) {
    /* nThreadsPerNode executor threads over the initial tasks (DistributedDC.java:221-224);
     * nothing inside a node step is parallel. */
    ctx.cache = (population**)calloc((size_t)tree->nNodes, sizeof(population*));
    const int32_t nN = tree->nNodes;
    int32_t* tasks = (int32_t*)malloc(sizeof(int32_t) * (size_t)nN);
    int nTasks = 0;
    initial_tasks(tree, tree->root, opt->maximumDistributionDepth, tasks, &nTasks);
    work_queue q;
    q.ctx = &ctx;
    q.ready = (int32_t*)malloc(sizeof(int32_t) * (size_t)nN);
    q.head = q.tail = 0;
    q.pending = (int32_t*)calloc((size_t)nN, sizeof(int32_t));
    q.isTask = (char*)calloc((size_t)nN, 1);
    q.parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)nN);
    for (int32_t n = 0; n < nN; ++n) q.parent[n] = -1;
    for (int32_t n = 0; n < nN; ++n)
      for (int c = tree->childOffsets[n]; c < tree->childOffsets[n + 1]; ++c) q.parent[tree->children[c]] = n;
    q.parent[tree->root] = -1;
    q.remaining = nTasks;
    for (int k = 0; k < nTasks; ++k) {
      q.isTask[tasks[k]] = 1;
      q.ready[q.tail++] = tasks[k];
    }
    for (int k = 0; k < nTasks; ++k) /* nodes above the cut = proper ancestors of the tasks (below `root`) */
      for (int32_t a = tasks[k] == tree->root ? -1 : q.parent[tasks[k]]; a >= 0 && q.pending[a] == 0; a = a == tree->root ? -1 : q.parent[a]) {
        q.pending[a] = n_children(tree, a);
        q.remaining++;
      }
    pthread_mutex_init(&q.mu, NULL);
    pthread_cond_init(&q.cv, NULL);
    const int nt = opt->nThreads;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, worker, &q);
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&q.mu);
    pthread_cond_destroy(&q.cv);
    free(th);
    free(tasks);
    free(q.ready);
    free(q.pending);
    free(q.isTask);
    free(q.parent);
  }

  population* rootPop = ctx.error ? NULL : run_node(&ctx, tree->root);
  if (rootPop && out) {
    const int32_t N = rootPop->n;
    for (int32_t i = 0; i < N; ++i) {
      if (out->rootState) {
        out->rootState[0 * (size_t)N + i] = rootPop->p[i].mean;
        out->rootState[1 * (size_t)N + i] = rootPop->p[i].msgVar;
        out->rootState[2 * (size_t)N + i] = rootPop->p[i].logLik;
        out->rootState[3 * (size_t)N + i] = rootPop->p[i].descObs;
        out->rootState[4 * (size_t)N + i] = rootPop->p[i].descVar;
        out->rootState[5 * (size_t)N + i] = rootPop->p[i].variance;
      }
      if (out->rootIState) out->rootIState[i] = rootPop->p[i].state;
      if (out->rootWeights) out->rootWeights[i] = rootPop->w ? rootPop->w[i] : 1.0 / (double)N;
    }
  }
  pop_free(rootPop);
  if (ctx.cache) {
    for (int32_t k = 0; k < tree->nNodes; ++k) pop_free(ctx.cache[k]);
    free(ctx.cache);
  }
  free(ctx.impl1LogZ);
  return ctx.error;
}
