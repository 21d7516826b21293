import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, dcsmc_b200 as d
from dcsmc_b200.engine import Engine
tree=d.DirectedTree(d.Node(("R",))); csr=tree.to_csr()
nT=np.array([12],np.int32); nS=np.array([0],np.int32)
N=20000
idx=np.arange(N*64+N,dtype=np.float64)
with Engine(nParticles=10) as e2:
    u=e2.debug_eval(2,idx)
with Engine(nParticles=N,resamplingScheme=1,rngMode=1,retainPopulations=1) as e:
    e.set_tree(csr); e.set_multilevel_model(nT,nS); e.inject_uniforms(0,u); e.run()
    print('ran ok')
