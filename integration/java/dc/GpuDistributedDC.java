// GpuDistributedDC.java — the class a maintainer of alexandrebouchard/divide-and-conquer-smc would add next to
// src/main/java/dc/DistributedDC.java to run the DC-SMC node step on libdcsmc.so (Panama FFM, JDK 22+).
// Same public surface as dc.DistributedDC (DistributedDC.java:58-107): createInstance / addProcessorFactory /
// getInstance / start / getRootPopulation, so dc/DDCMain.java:21-30 changes by ONE token (the class name).
// SHIPPED UNCOMPILED: this build environment has no JDK and none of the reference's dependencies
// (bayonet, briefj, ...); the same call sequence is exercised through ctypes by
// divide-and-conquer-smc_b200/_ffi.py and tests/test_gpu_dc_api.py. Kept in sync with INTEGRATION.md §2 by
// tests/test_host_logic.py::test_java_shim_matches_the_abi.
package dc;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.util.*;
import static java.lang.foreign.ValueLayout.*;

import bayonet.graphs.DirectedTree;
import bayonet.smc.ParticlePopulation;
import bayonet.smc.ResamplingScheme;

/** GPU drop-in for DistributedDC&lt;P,N&gt; for the proposal families the kernels implement. */
public final class GpuDistributedDC<P, N> {
  private static final Linker L = Linker.nativeLinker();
  private static final SymbolLookup LIB = SymbolLookup.libraryLookup("libdcsmc.so", Arena.global());
  private static MethodHandle h(String name, FunctionDescriptor fd) { return L.downcallHandle(LIB.find(name).orElseThrow(), fd); }

  // struct dcsmc_options (include/dcsmc.h) — DCOptions field for field, then engine extensions
  static final StructLayout OPTIONS = MemoryLayout.structLayout(
      JAVA_LONG.withName("masterRandomSeed"), JAVA_INT.withName("nParticles"), JAVA_INT.withName("resamplingScheme"),
      JAVA_DOUBLE.withName("relativeEssThreshold"), JAVA_INT.withName("clusterSubGroup"), JAVA_INT.withName("nThreadsPerNode"),
      JAVA_INT.withName("minimumNumberOfClusterMembersToStart"), JAVA_INT.withName("maximumTimeToWaitInMinutes"),
      JAVA_INT.withName("maximumDistributionDepth"), JAVA_INT.withName("indexInCluster"),
      JAVA_INT.withName("device"), JAVA_INT.withName("rngMode"), JAVA_INT.withName("sumMode"), JAVA_INT.withName("maxBetaAttempts"),
      JAVA_INT.withName("retainPopulations"), JAVA_INT.withName("levelCutOffForOutput"), JAVA_LONG.withName("memoryBudgetBytes"),
      JAVA_INT.withName("useTransform"), JAVA_INT.withName("retainSamples"));   // ABI version 3
  static final StructLayout NODE_STATS = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("ess"), JAVA_DOUBLE.withName("relEss"), JAVA_DOUBLE.withName("logNormEstimate"),
      JAVA_DOUBLE.withName("logScaling"), JAVA_INT.withName("resampled"), JAVA_INT.withName("done"));

  static final MethodHandle ABI_VERSION = h("dcsmc_abi_version", FunctionDescriptor.of(JAVA_INT));
  static final MethodHandle OPT_DEFAULT = h("dcsmc_options_default", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle CREATE      = h("dcsmc_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle DESTROY     = h("dcsmc_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle LAST_ERROR  = h("dcsmc_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
  static final MethodHandle SET_TREE    = h("dcsmc_set_tree", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle SET_SEEDS   = h("dcsmc_set_node_seeds", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle SET_MULTI   = h("dcsmc_set_multilevel_model", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_DOUBLE));
  static final MethodHandle PROFILE     = h("dcsmc_profile_enable", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  static final MethodHandle RUN         = h("dcsmc_run", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle ALL_STATS   = h("dcsmc_all_node_stats", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle NODE_TIMES  = h("dcsmc_node_times", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle POP_COPY    = h("dcsmc_population_copy_to_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));

  /** Turns row i of the root population's columns (mean, msgVar, logLik, descObs, descVar, variance) into a P. */
  public interface ParticleDecoder<P> { P decode(double[][] columns, int i); }

  final DCOptions options;
  final DCProposalFactory<P, N> proposalFactory;
  final DirectedTree<N> tree;
  final List<DCProcessorFactory<P, N>> processorFactories = new ArrayList<>();
  private final List<N> nodes = new ArrayList<>();   // pre-order, root = 0
  private final Arena arena = Arena.ofConfined();
  private MemorySegment engine = MemorySegment.NULL;
  private ParticleDecoder<P> decoder = null;
  private ParticlePopulation<P> rootPopulation = null;
  private boolean started = false;
  @SuppressWarnings("rawtypes") private static GpuDistributedDC instance = null;

  /** DistributedDC.createInstance(options, proposalFactory, tree) (DistributedDC.java:58-70): same arguments, same
   *  one-instance-per-JVM rule (the DCProcessorFactoryContext of the processors resolves the singleton). */
  public static <P, N> GpuDistributedDC<P, N> createInstance(DCOptions options, DCProposalFactory<P, N> proposalFactory, DirectedTree<N> tree) {
    if (instance != null) throw new RuntimeException("Limitation: only one distributed DC computation at any given time in one JVM.");
    // registers the (never started: no Hazelcast) reference instance so that DCProcessorFactoryContext, whose
    // constructor calls DistributedDC.getInstance() (DCProcessorFactoryContext.java:12-16), keeps working unchanged
    DistributedDC.createInstance(options, proposalFactory, tree);
    GpuDistributedDC<P, N> r = new GpuDistributedDC<>(options, proposalFactory, tree);
    instance = r;
    return r;
  }

  /** DistributedDC.addProcessorFactory (DistributedDC.java:72-75). */
  public void addProcessorFactory(DCProcessorFactory<P, N> factory) { this.processorFactories.add(factory); }

  /** DistributedDC.getInstance (DistributedDC.java:77-85). */
  @SuppressWarnings("unchecked")
  public static <P, N> GpuDistributedDC<P, N> getInstance() {
    if (instance == null) throw new RuntimeException("Use createInstance(..) first.");
    return instance;
  }

  /** How getRootPopulation() materialises particles of type P (default: none — weights and logScaling only). */
  public void setParticleDecoder(ParticleDecoder<P> decoder) { this.decoder = decoder; }

  private GpuDistributedDC(DCOptions options, DCProposalFactory<P, N> proposalFactory, DirectedTree<N> tree) {
    this.options = options;
    this.proposalFactory = proposalFactory;
    this.tree = tree;
    // the default processor of the reference reads context.dc.cluster (DefaultProcessorFactory.java:43), which does
    // not exist here: the same `timing` / `logZ` / `workTime` outputs are written by start() itself
  }

  /** DistributedDC.start() (DistributedDC.java:87-92): blocking; throws if called twice (:177-182). */
  public void start() {
    if (started) throw new RuntimeException();
    started = true;
    try { run(); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
  }

  private void run() throws Throwable {
    if ((int) ABI_VERSION.invoke() != 3) throw new RuntimeException("libdcsmc.so: ABI version mismatch");
    if (!(proposalFactory instanceof prototype.adaptor.MultiLevelProposalFactory))
      throw new RuntimeException("GpuDistributedDC: no kernel family for " + proposalFactory.getClass().getName());
    prototype.adaptor.MultiLevelProposalFactory f = (prototype.adaptor.MultiLevelProposalFactory) proposalFactory;
    DCOptions o = options;
    MemorySegment opt = arena.allocate(OPTIONS);
    check((int) OPT_DEFAULT.invoke(opt), MemorySegment.NULL);
    opt.set(JAVA_LONG, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("masterRandomSeed")), o.masterRandomSeed);
    opt.set(JAVA_INT, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("nParticles")), o.nParticles);
    opt.set(JAVA_INT, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("resamplingScheme")),
            o.resamplingScheme == ResamplingScheme.MULTINOMIAL ? 0 : 1);
    opt.set(JAVA_DOUBLE, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("relativeEssThreshold")), o.relativeEssThreshold);
    opt.set(JAVA_INT, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("maximumDistributionDepth")), o.maximumDistributionDepth);
    MemorySegment out = arena.allocate(ADDRESS);
    check((int) CREATE.invoke(opt, out), MemorySegment.NULL);
    engine = out.get(ADDRESS, 0);
    final int nParticles = o.nParticles;

    // CSR in the tree's child order (DCRecursionTask.java:51) + Node.hashCode() per node (:98-106)
    Map<N, Integer> index = new HashMap<>();
    ArrayDeque<N> stack = new ArrayDeque<>(); stack.push(tree.getRoot());
    while (!stack.isEmpty()) { N n = stack.pop(); index.put(n, nodes.size()); nodes.add(n);
      List<N> ch = new ArrayList<>(tree.getChildren(n)); Collections.reverse(ch); ch.forEach(stack::push); }
    int nN = nodes.size();
    MemorySegment off = arena.allocate(JAVA_INT, nN + 1), chd = arena.allocate(JAVA_INT, Math.max(1, nN - 1)),
        hash = arena.allocate(JAVA_INT, nN), nT = arena.allocate(JAVA_INT, nN), nS = arena.allocate(JAVA_INT, nN);
    int e = 0;
    for (int i = 0; i < nN; i++) {
      off.setAtIndex(JAVA_INT, i, e);
      for (N c : tree.getChildren(nodes.get(i))) chd.setAtIndex(JAVA_INT, e++, index.get(c));
      hash.setAtIndex(JAVA_INT, i, nodes.get(i).hashCode());
      if (tree.isLeaf(nodes.get(i))) { prototype.io.Datum d = f.getDataset().getDatum((prototype.Node) nodes.get(i));
        nT.setAtIndex(JAVA_INT, i, d.numberOfTrials); nS.setAtIndex(JAVA_INT, i, d.numberOfSuccesses); }
    }
    off.setAtIndex(JAVA_INT, nN, e);
    check((int) SET_TREE.invoke(engine, nN, 0, off, chd), engine);
    check((int) SET_SEEDS.invoke(engine, hash), engine);
    check((int) SET_MULTI.invoke(engine, MemorySegment.NULL, nT, nS, MemorySegment.NULL, f.variancePrior), engine);

    long t0 = System.currentTimeMillis();
    check((int) PROFILE.invoke(engine, 1), engine);   // per-launch events -> iterationProposalTime per node
    check((int) RUN.invoke(engine), engine);
    long globalTime = System.currentTimeMillis() - t0;

    // DCRecursion.dcRecurse :26-28: every processor sees the population BEFORE resampling and the children populations
    MemorySegment stats = arena.allocate(NODE_STATS, nN), times = arena.allocate(JAVA_DOUBLE, nN);
    check((int) ALL_STATS.invoke(engine, stats), engine);
    check((int) NODE_TIMES.invoke(engine, times), engine);
    briefj.OutputManager output = new briefj.OutputManager();
    output.setOutputFolder(briefj.run.Results.getResultFolder());
    List<ParticlePopulation<P>> pops = new ArrayList<>(Collections.nCopies(nN, null));
    for (int i = nN - 1; i >= 0; i--) {              // reverse pre-order: children before parents
      final long base = NODE_STATS.byteSize() * i;
      double ess = stats.get(JAVA_DOUBLE, base), logScaling = stats.get(JAVA_DOUBLE, base + 24);
      boolean resampled = stats.get(JAVA_INT, base + 32) != 0;
      pops.set(i, StatsOnlyPopulation.of(nParticles, ess, logScaling));
      List<ParticlePopulation<P>> kids = new ArrayList<>();
      for (N c : tree.getChildren(nodes.get(i))) {
        int ci = index.get(c);                        // after resample(): null weights, ESS = N (DCRecursion.java:29-31)
        boolean cres = stats.get(JAVA_INT, NODE_STATS.byteSize() * ci + 32) != 0;
        kids.add(cres ? StatsOnlyPopulation.of(nParticles, nParticles, stats.get(JAVA_DOUBLE, NODE_STATS.byteSize() * ci + 24)) : pops.get(ci));
      }
      for (DCProcessorFactory<P, N> factory : processorFactories)
        factory.build(new DCProcessorFactoryContext<P, N>(nodes.get(i), tree)).process(pops.get(i), kids);
      output.printWrite("timing", "node", nodes.get(i), "ESS", ess, "rESS", ess / nParticles,          // DefaultProcessorFactory.java:45-52
          "logZ", logScaling - Math.log(nParticles), "nWorkers", 1, "iterationProposalTime", times.getAtIndex(JAVA_DOUBLE, i),
          "globalTime", globalTime);
      if (i == 0) briefj.BriefIO.write(briefj.run.Results.getFileInResultFolder("logZ"), "" + (logScaling - Math.log(nParticles)));
      if (i == 0 && resampled) pops.set(0, StatsOnlyPopulation.of(nParticles, nParticles, logScaling));
    }
    output.flush();
    for (DCProcessorFactory<P, N> factory : processorFactories) factory.close();               // DistributedDC.java:98-99
    briefj.BriefIO.write(briefj.run.Results.getFileInResultFolder("workTime"), "" + globalTime);   // DefaultProcessorFactory.java:59-62

    // for convenience, save the root population locally (DistributedDC.java:100-102)
    MemorySegment st = arena.allocate(JAVA_DOUBLE, 6L * nParticles), w = arena.allocate(JAVA_DOUBLE, nParticles);
    check((int) POP_COPY.invoke(engine, 0, st, MemorySegment.NULL, w), engine);
    double[][] cols = new double[6][];
    for (int c = 0; c < 6; c++) cols[c] = st.asSlice(8L * c * nParticles, 8L * nParticles).toArray(JAVA_DOUBLE);
    double[] weights = w.toArray(JAVA_DOUBLE);
    List<P> particles = new ArrayList<>(nParticles);
    for (int i = 0; i < nParticles; i++) particles.add(decoder == null ? null : decoder.decode(cols, i));
    // buildDestructivelyFromLogWeights(logW, particles, base): logScaling = base + LSE(logW) (DCRecursion.java:73), so
    // base = logScaling - LSE(log w) = logScaling (the weights are normalised: LSE = 0)
    double[] logW = new double[nParticles];
    for (int i = 0; i < nParticles; i++) logW[i] = Math.log(weights[i]);
    rootPopulation = ParticlePopulation.buildDestructivelyFromLogWeights(logW, particles, stats.get(JAVA_DOUBLE, 24));
    DESTROY.invoke(engine);
    arena.close();
    instance = null;
  }

  /** DistributedDC.getRootPopulation() (DistributedDC.java:104-107). */
  public ParticlePopulation<P> getRootPopulation() { return rootPopulation; }

  /** What a DCProcessor may read of a population that stays on the GPU: getESS / getRelativeESS / logNormEstimate /
   *  logScaling / nParticles (DefaultProcessorFactory.java:45-47). A maintainer adds the matching package-private
   *  factory to bayonet.smc.ParticlePopulation (statistics only, particles == null). */
  static final class StatsOnlyPopulation {
    static <P> ParticlePopulation<P> of(int nParticles, double ess, double logScaling) {
      return ParticlePopulation.statisticsOnly(nParticles, ess, logScaling);
    }
  }

  private static void check(int rc, MemorySegment eng) throws Throwable {
    if (rc != 0) {  // the reference throws bare RuntimeExceptions (DCRecursion.java:43,50)
      MemorySegment msg = (MemorySegment) LAST_ERROR.invoke(eng);
      throw new RuntimeException("dcsmc error " + rc + ": " + msg.reinterpret(512).getString(0));
    }
  }
}
