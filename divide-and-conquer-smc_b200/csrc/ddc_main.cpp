// ddc_main.cpp — native command line with the reference's flags (drop-in for scripts/run-distributed.sh).
//
// Mirrors dc/DDCMain.java:12-36: option set "dc" = DCOptions (dc/DCOptions.java:15-45), option set
// "multiLevel" = MultiLevelProposalFactory (-dataFile, -variancePrior;
// prototype/adaptor/MultiLevelProposalFactory.java:22-26). The briefj runner accepts `-flag value`;
// so does this. Input = the CSV of prototype/io/MultiLevelDataset.java:51-66 (path..., nTrials,
// nSuccesses; children in first-appearance order). Output = what DefaultProcessorFactory writes
// (dc/DefaultProcessorFactory.java:41-62): one `timing` row per node on stdout and in
// <results>/timing.csv, the root estimate in <results>/logZ, total time in <results>/workTime.
//
//   dcsmc_ddc -dataFile data.csv -nParticles 100000 -resamplingScheme MULTINOMIAL
//             -relativeEssThreshold 1.0000000001 -masterRandomSeed 31 [-resultsFolder results/latest]
//   dcsmc_ddc -dataFile data.csv -dumpTree      # CSR + Node.hashCode() only, no GPU needed
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dcsmc.h"

namespace {

// java.lang.String.hashCode over UTF-16 code units (ASCII/BMP input assumed byte-wise for ASCII)
int32_t javaStringHash(const std::string& s) {
  uint32_t h = 0;
  for (unsigned char c : s) h = 31u * h + c;
  return (int32_t)h;
}

struct Tree {
  std::vector<std::vector<std::string>> path;  // node -> path (pre-order numbering, root = 0)
  std::vector<int32_t> childOffsets, children, javaHash, nTrials, nSuccesses;
};

std::vector<std::string> splitCsv(const std::string& line) {
  std::vector<std::string> out;
  std::string cur;
  bool quoted = false;
  for (char c : line) {
    if (c == '"') quoted = !quoted;
    else if (c == ',' && !quoted) { out.push_back(cur); cur.clear(); }
    else if (c != '\r') cur += c;
  }
  out.push_back(cur);
  return out;
}

std::string joinPath(const std::vector<std::string>& p, size_t n, char sep) {
  std::string s;
  for (size_t i = 0; i < n; ++i) {
    if (i) s += sep;
    s += p[i];
  }
  return s;
}

// MultiLevelDataset(File) + getTree(): insertion-ordered children, then pre-order numbering.
bool loadDataset(const std::string& file, Tree& t, std::string& err) {
  std::ifstream in(file);
  if (!in) { err = "cannot open " + file; return false; }
  std::unordered_map<std::string, int> id;                // key = path joined by '\x1f'
  std::vector<std::vector<std::string>> paths;            // insertion id -> path
  std::vector<std::vector<int>> kids;                     // insertion id -> children (first appearance)
  std::vector<int> nT, nS;
  auto intern = [&](const std::vector<std::string>& p, size_t n) {
    const std::string key = joinPath(p, n, '\x1f');
    auto it = id.find(key);
    if (it != id.end()) return it->second;
    const int k = (int)paths.size();
    id.emplace(key, k);
    paths.emplace_back(p.begin(), p.begin() + n);
    kids.emplace_back();
    nT.push_back(0);
    nS.push_back(0);
    return k;
  };
  std::string line;
  int root = -1;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    std::vector<std::string> f = splitCsv(line);
    if (f.size() < 3) { err = "bad CSV row: " + line; return false; }
    const size_t np = f.size() - 2;
    if (root < 0) root = intern(f, 1);
    for (size_t i = 0; i + 1 < np; ++i) {
      const int par = intern(f, i + 1);
      const bool isNew = id.find(joinPath(f, i + 2, '\x1f')) == id.end();
      const int ch = intern(f, i + 2);
      if (isNew) kids[par].push_back(ch);
      else {
        bool found = false;
        for (int c : kids[par]) found |= (c == ch);
        if (!found) kids[par].push_back(ch);
      }
    }
    const int leaf = intern(f, np);
    nT[leaf] = std::atoi(f[np].c_str());
    nS[leaf] = std::atoi(f[np + 1].c_str());
  }
  if (root < 0) { err = "empty data file"; return false; }
  // pre-order DFS numbering (root = 0), children in insertion order
  std::vector<int> order, newId(paths.size(), -1), stack{root};
  while (!stack.empty()) {
    const int n = stack.back();
    stack.pop_back();
    newId[n] = (int)order.size();
    order.push_back(n);
    for (auto it = kids[n].rbegin(); it != kids[n].rend(); ++it) stack.push_back(*it);
  }
  const int nN = (int)order.size();
  t.childOffsets.assign(nN + 1, 0);
  t.children.clear();
  t.path.resize(nN);
  t.javaHash.resize(nN);
  t.nTrials.resize(nN);
  t.nSuccesses.resize(nN);
  for (int i = 0; i < nN; ++i) {
    const int n = order[i];
    t.path[i] = paths[n];
    for (int c : kids[n]) t.children.push_back(newId[c]);
    t.childOffsets[i + 1] = (int32_t)t.children.size();
    // Node.hashCode = 31*1 + List<String>.hashCode (prototype/Node.java:34-40)
    uint32_t lh = 1;
    for (const std::string& s : paths[n]) lh = 31u * lh + (uint32_t)javaStringHash(s);
    t.javaHash[i] = (int32_t)(31u + lh);
    t.nTrials[i] = nT[n];
    t.nSuccesses[i] = nS[n];
  }
  return true;
}

void usage() {
  std::puts(
      "usage: dcsmc_ddc -dataFile <csv> [-nParticles 1000] [-masterRandomSeed 31]\n"
      "        [-resamplingScheme MULTINOMIAL|STRATIFIED] [-relativeEssThreshold 1.0000000001]\n"
      "        [-maximumDistributionDepth <int>] [-nThreadsPerNode 1] [-variancePrior 1.0]\n"
      "        [-clusterSubGroup 1] [-minimumNumberOfClusterMembersToStart 1] [-maximumTimeToWaitInMinutes 1]\n"
      "        [-indexInCluster 1] [-device 0] [-resultsFolder results/latest] [-dumpTree]\n"
      "        [-memoryBudgetGiB 0] [-planOnly]   (-planOnly: print the schedule and memory plan, no GPU needed)\n"
      "        [-levelCutOffForOutput 0] [-useTransform true]   (posterior summaries of impl-1: meanStats,\n"
      "        varStats, deltaStats for nodes with level < cut-off; the reference's default cut-off is 2)\n"
      "flags are the field names of dc.DCOptions / MultiLevelProposalFactory (dc/DDCMain.java:14-18)");
}

}  // namespace

int main(int argc, char** argv) {
  dcsmc_options opt;
  dcsmc_options_default(&opt);
  std::string dataFile, results = "results/latest";
  double variancePrior = 1.0;
  bool dumpTree = false, planOnly = false;
  long long memoryBudgetGiB = 0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    // briefj also accepts the option-set prefix, e.g. -dc.nParticles / -multiLevel.dataFile
    const size_t dot = a.find('.');
    if (dot != std::string::npos && (a.compare(0, dot, "dc") == 0 || a.compare(0, dot, "multiLevel") == 0)) a = a.substr(dot + 1);
    auto val = [&]() -> const char* {
      if (i + 1 >= argc) { std::fprintf(stderr, "missing value for -%s\n", a.c_str()); std::exit(2); }
      return argv[++i];
    };
    if (a == "help" || a == "h") { usage(); return 0; }
    else if (a == "dumpTree") dumpTree = true;
    else if (a == "planOnly") planOnly = true;
    else if (a == "memoryBudgetGiB") { memoryBudgetGiB = std::atoll(val()); opt.memoryBudgetBytes = memoryBudgetGiB << 30; }
    else if (a == "dataFile") dataFile = val();
    else if (a == "variancePrior") variancePrior = std::atof(val());
    else if (a == "resultsFolder") results = val();
    else if (a == "nParticles") opt.nParticles = std::atoi(val());
    else if (a == "masterRandomSeed") opt.masterRandomSeed = std::atoll(val());
    else if (a == "relativeEssThreshold") opt.relativeEssThreshold = std::atof(val());
    else if (a == "maximumDistributionDepth") opt.maximumDistributionDepth = std::atoi(val());
    else if (a == "nThreadsPerNode") opt.nThreadsPerNode = std::atoi(val());
    else if (a == "clusterSubGroup") opt.clusterSubGroup = std::atoi(val());
    else if (a == "minimumNumberOfClusterMembersToStart") opt.minimumNumberOfClusterMembersToStart = std::atoi(val());
    else if (a == "maximumTimeToWaitInMinutes") opt.maximumTimeToWaitInMinutes = std::atoi(val());
    else if (a == "indexInCluster") opt.indexInCluster = std::atoi(val());
    else if (a == "device") opt.device = std::atoi(val());
    // impl-1 output options (prototype/smc/DivideConquerMCAlgorithm.java:47,56)
    else if (a == "levelCutOffForOutput") opt.levelCutOffForOutput = std::atoi(val());
    else if (a == "useTransform") opt.useTransform = std::string(val()) != "false";
    else if (a == "resamplingScheme") {
      const std::string v = val();
      if (v == "MULTINOMIAL") opt.resamplingScheme = DCSMC_MULTINOMIAL;
      else if (v == "STRATIFIED") opt.resamplingScheme = DCSMC_STRATIFIED;
      else { std::fprintf(stderr, "unknown resamplingScheme %s\n", v.c_str()); return 2; }
    } else { std::fprintf(stderr, "unknown option -%s\n", a.c_str()); usage(); return 2; }
  }
  if (dataFile.empty()) { std::fprintf(stderr, "-dataFile is required\n"); usage(); return 2; }  // @Option(required = true)
  Tree t;
  std::string err;
  if (!loadDataset(dataFile, t, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
  const int nN = (int)t.path.size();
  if (dumpTree) {
    std::printf("nNodes %d\n", nN);
    for (int i = 0; i < nN; ++i) {
      std::printf("%d %s hash %d trials %d successes %d children", i, joinPath(t.path[i], t.path[i].size(), '-').c_str(),
                  t.javaHash[i], t.nTrials[i], t.nSuccesses[i]);
      for (int c = t.childOffsets[i]; c < t.childOffsets[i + 1]; ++c) std::printf(" %d", t.children[c]);
      std::printf("\n");
    }
    return 0;
  }
  if (planOnly) {  // the static memory planner on the host: sizes a job before anything is allocated
    dcsmc_plan_info pi;
    char msg[512] = "";
    const int rc = dcsmc_plan_query(&opt, nN, t.childOffsets.data(), t.children.empty() ? t.childOffsets.data() : t.children.data(),
                                    nullptr, DCSMC_MODEL_MULTILEVEL, nullptr, 0, opt.memoryBudgetBytes, &pi, msg, (int)sizeof msg);
    if (rc) { std::fprintf(stderr, "dcsmc_plan_query: %s\n", msg); return 1; }
    std::printf("plan: nodes=%d nParticles=%d feasible=%d levelBatches=%d maxBatchNodes=%d\n", nN, opt.nParticles, pi.feasible,
                pi.levelBatches, pi.maxBatchNodes);
    std::printf("plan: poolBytesNeeded=%lld poolBytes=%lld peakBytes=%lld scratchBytes=%lld tableBytes=%lld\n",
                (long long)pi.poolBytesNeeded, (long long)pi.poolBytes, (long long)pi.peakBytes, (long long)pi.scratchBytes,
                (long long)pi.tableBytes);
    if (!pi.feasible) std::printf("plan: %s\n", msg);
    return pi.feasible ? 0 : 3;
  }
  dcsmc_engine* e = nullptr;
  if (dcsmc_create(&opt, &e)) { std::fprintf(stderr, "dcsmc_create: %s\n", dcsmc_last_error(nullptr)); return 1; }
  auto ck = [&](int rc, const char* what) {
    if (rc) { std::fprintf(stderr, "%s: %s\n", what, dcsmc_last_error(e)); dcsmc_destroy(e); std::exit(1); }
  };
  const auto t0 = std::chrono::steady_clock::now();
  ck(dcsmc_set_tree(e, nN, 0, t.childOffsets.data(), t.children.empty() ? t.childOffsets.data() : t.children.data()), "dcsmc_set_tree");
  ck(dcsmc_set_node_seeds(e, t.javaHash.data()), "dcsmc_set_node_seeds");
  ck(dcsmc_set_multilevel_model(e, nullptr, t.nTrials.data(), t.nSuccesses.data(), nullptr, variancePrior), "dcsmc_set_multilevel_model");
  ck(dcsmc_run(e), "dcsmc_run");
  const double globalMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::vector<dcsmc_node_stats_t> st(nN);
  ck(dcsmc_all_node_stats(e, st.data()), "dcsmc_all_node_stats");
  dcsmc_run_info info;
  dcsmc_last_run_info(e, &info);
  // results folder (mkdir -p)
  {
    std::string acc;
    std::stringstream ss(results);
    std::string part;
    while (std::getline(ss, part, '/')) {
      acc += part + "/";
      if (!part.empty()) mkdir(acc.c_str(), 0755);
    }
  }
  std::ofstream timing(results + "/timing.csv");
  timing << "node,ESS,rESS,logZ,nWorkers,iterationProposalTime,globalTime\n";
  // the reference emits rows as nodes finish (children before parents): heights ascending here
  std::vector<int> height(nN, 0), parent(nN, -1);
  for (int i = 0; i < nN; ++i)
    for (int c = t.childOffsets[i]; c < t.childOffsets[i + 1]; ++c) parent[t.children[c]] = i;
  for (int i = nN - 1; i > 0; --i) height[parent[i]] = std::max(height[parent[i]], height[i] + 1);
  std::vector<int> order(nN);
  for (int i = 0; i < nN; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return height[a] < height[b]; });
  char buf[512];
  for (int i : order) {
    const std::string name = joinPath(t.path[i], t.path[i].size(), '-');
    std::snprintf(buf, sizeof buf, "%s,%.17g,%.17g,%.17g,1,%.3f,%.3f", name.c_str(), st[i].ess, st[i].relEss,
                  st[i].logNormEstimate, info.deviceMs / nN, globalMs);
    timing << buf << "\n";
    if (nN <= 64 || i == 0)
      std::printf("timing: node=%s, ESS=%.17g, rESS=%.17g, logZ=%.17g, nWorkers=1, iterationProposalTime=%.3f, globalTime=%.3f\n",
                  name.c_str(), st[i].ess, st[i].relEss, st[i].logNormEstimate, info.deviceMs / nN, globalMs);
  }
  if (opt.levelCutOffForOutput > 0) {
    // printNodeStatistics / printDeltaNodeStatistics tables (DivideConquerMCAlgorithm.java:433-507)
    std::ofstream meanStats(results + "/meanStats.csv"), varStats(results + "/varStats.csv"),
        deltaStats(results + "/deltaStats.csv");
    meanStats << "path,meanMean,meanSD\n";
    varStats << "path,varMean,varSD\n";
    deltaStats << "path,deltaMean,deltaSD\n";
    for (int i : order) {
      if ((int)t.path[i].size() - 1 > opt.levelCutOffForOutput) continue;  // deltas sit one level below the cut
      dcsmc_node_summary_t s;
      if (dcsmc_node_summary(e, i, &s) != DCSMC_OK) continue;
      const std::string name = joinPath(t.path[i], t.path[i].size(), '-');
      if (s.hasMean) {
        std::snprintf(buf, sizeof buf, "%s,%.17g,%.17g", name.c_str(), s.meanMean, s.meanSD);
        meanStats << buf << "\n";
      }
      if (s.hasVar) {
        std::snprintf(buf, sizeof buf, "%s,%.17g,%.17g", name.c_str(), s.varMean, s.varSD);
        varStats << buf << "\n";
      }
      if (s.hasDelta) {
        std::snprintf(buf, sizeof buf, "%s,%.17g,%.17g", name.c_str(), s.deltaMean, s.deltaSD);
        deltaStats << buf << "\n";
      }
    }
  }
  std::snprintf(buf, sizeof buf, "%.17g", st[0].logNormEstimate);
  std::ofstream(results + "/logZ") << buf;
  std::ofstream(results + "/workTime") << (long long)globalMs;
  std::printf("logZ estimate = %s  (%lld particle-node updates in %.3f ms on the device, %lld kernel launches)\n", buf,
              (long long)info.particleUpdates, info.deviceMs, (long long)info.kernelLaunches);
  dcsmc_destroy(e);
  return 0;
}
