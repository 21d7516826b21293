// GpuDistributedDC.java — the class a maintainer of alexandrebouchard/divide-and-conquer-smc would add next to
// src/main/java/dc/DistributedDC.java to run the DC-SMC node step on libdcsmc.so (Panama FFM, JDK 22+).
// SHIPPED UNCOMPILED: this build environment has no JDK and none of the reference's dependencies
// (bayonet, briefj, ...); the same call sequence is exercised through ctypes by
// divide-and-conquer-smc_b200/_ffi.py and tests/test_gpu_dc_api.py. Kept in sync with INTEGRATION.md §2 by
// tests/test_host_logic.py::test_java_shim_matches_the_abi.
package dc;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

/** GPU drop-in for DistributedDC for the proposal families the kernels implement. */
public final class GpuDistributedDC<N> {
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
      JAVA_INT.withName("useTransform"), JAVA_INT.withName("reserved1"));   // ABI version 2
  static final StructLayout NODE_STATS = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("ess"), JAVA_DOUBLE.withName("relEss"), JAVA_DOUBLE.withName("logNormEstimate"),
      JAVA_DOUBLE.withName("logScaling"), JAVA_INT.withName("resampled"), JAVA_INT.withName("done"));

  static final MethodHandle OPT_DEFAULT = h("dcsmc_options_default", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle CREATE      = h("dcsmc_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle DESTROY     = h("dcsmc_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle LAST_ERROR  = h("dcsmc_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
  static final MethodHandle SET_TREE    = h("dcsmc_set_tree", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle SET_SEEDS   = h("dcsmc_set_node_seeds", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle SET_MULTI   = h("dcsmc_set_multilevel_model", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_DOUBLE));
  static final MethodHandle RUN         = h("dcsmc_run", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  static final MethodHandle NODE_STATS_F= h("dcsmc_node_stats", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle POP_COPY    = h("dcsmc_population_copy_to_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));

  private final Arena arena = Arena.ofConfined();
  private final MemorySegment engine;
  private final java.util.List<N> nodes = new java.util.ArrayList<>();   // pre-order, root = 0
  private final int nParticles;

  /** Same arguments as DistributedDC.createInstance(options, proposalFactory, tree). */
  public GpuDistributedDC(DCOptions o, prototype.adaptor.MultiLevelProposalFactory f, bayonet.graphs.DirectedTree<N> tree) throws Throwable {
    MemorySegment opt = arena.allocate(OPTIONS);
    check((int) OPT_DEFAULT.invoke(opt), MemorySegment.NULL);
    opt.set(JAVA_LONG, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("masterRandomSeed")), o.masterRandomSeed);
    opt.set(JAVA_INT, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("nParticles")), o.nParticles);
    opt.set(JAVA_INT, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("resamplingScheme")),
            o.resamplingScheme == bayonet.smc.ResamplingScheme.MULTINOMIAL ? 0 : 1);
    opt.set(JAVA_DOUBLE, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("relativeEssThreshold")), o.relativeEssThreshold);
    opt.set(JAVA_INT, OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("maximumDistributionDepth")), o.maximumDistributionDepth);
    MemorySegment out = arena.allocate(ADDRESS);
    check((int) CREATE.invoke(opt, out), MemorySegment.NULL);
    engine = out.get(ADDRESS, 0);
    nParticles = o.nParticles;

    // CSR in the tree's child order (DCRecursionTask.java:51) + Node.hashCode() per node (:98-106)
    java.util.Map<N, Integer> index = new java.util.HashMap<>();
    java.util.ArrayDeque<N> stack = new java.util.ArrayDeque<>(); stack.push(tree.getRoot());
    while (!stack.isEmpty()) { N n = stack.pop(); index.put(n, nodes.size()); nodes.add(n);
      java.util.List<N> ch = new java.util.ArrayList<>(tree.getChildren(n)); java.util.Collections.reverse(ch); ch.forEach(stack::push); }
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
  }

  /** DistributedDC.start(): blocking; processors are fed from dcsmc_node_stats afterwards. */
  public void start() throws Throwable { check((int) RUN.invoke(engine), engine); }

  /** ParticlePopulation.logNormEstimate() of the root (what DefaultProcessorFactory writes to `logZ`). */
  public double rootLogNormEstimate() throws Throwable {
    MemorySegment s = arena.allocate(NODE_STATS);
    check((int) NODE_STATS_F.invoke(engine, 0, s), engine);
    return s.get(JAVA_DOUBLE, NODE_STATS.byteOffset(MemoryLayout.PathElement.groupElement("logNormEstimate")));
  }

  /** getRootPopulation(): columns mean, msgVar, logLik, descObs, descVar, variance ([6][N]) + weights. */
  public double[][] rootPopulation() throws Throwable {
    MemorySegment st = arena.allocate(JAVA_DOUBLE, 6L * nParticles), w = arena.allocate(JAVA_DOUBLE, nParticles);
    check((int) POP_COPY.invoke(engine, 0, st, MemorySegment.NULL, w), engine);
    double[][] r = new double[7][];
    for (int c = 0; c < 6; c++) r[c] = st.asSlice(8L * c * nParticles, 8L * nParticles).toArray(JAVA_DOUBLE);
    r[6] = w.toArray(JAVA_DOUBLE);
    return r;
  }

  public void close() throws Throwable { DESTROY.invoke(engine); arena.close(); }

  private static void check(int rc, MemorySegment eng) throws Throwable {
    if (rc != 0) {  // the reference throws bare RuntimeExceptions (DCRecursion.java:43,50)
      MemorySegment msg = (MemorySegment) LAST_ERROR.invoke(eng);
      throw new RuntimeException("dcsmc error " + rc + ": " + msg.reinterpret(512).getString(0));
    }
  }
}
