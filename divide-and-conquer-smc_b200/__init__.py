"""dcsmc_b200 — B200-native DC-SMC node engine behind the reference's dc.* plugin surface.

Host-side mirror of the reference interfaces (Python over the C ABI in include/dcsmc.h);
all compute happens in the CUDA library built from csrc/.
"""
from .tree import (  # noqa: F401
    CsrTree,
    Datum,
    DirectedTree,
    MultiLevelDataset,
    Node,
    java_list_hash,
    java_string_hash,
    perfectBinaryTree,
)
from .dc import (  # noqa: F401,E402
    DCOptions,
    DCProcessor,
    DCProcessorFactory,
    DCProcessorFactoryContext,
    DCProposalFactory,
    DefaultProcessorFactory,
    PosteriorSummaryProcessorFactory,
    DistributedDC,
    GaussianLeafProposalFactory,
    MarkovChainNaiveProposalFactory,
    MultiLevelProposalFactory,
    NoOpProcessor,
    ParticlePopulation,
    ResamplingScheme,
)
