import sys, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, dcsmc_b200 as d, oracle_lib as o
from dcsmc_b200 import _ffi
var=sys.argv[1]
_ffi.LIB_PATH=os.path.join(_ffi.HERE,'csrc','libdcsmc_%s.so'%var)
from dcsmc_b200.engine import Engine
tree=d.DirectedTree(d.Node(("R",))); csr=tree.to_csr()
for (n,s) in [(12,0),(9,9),(1,1),(30,11)]:
  nT=np.array([n],np.int32); nS=np.array([s],np.int32)
  N=200000
  ref=o.run(csr,N,resamplingScheme=1,rngMode=o.RNG_PHILOX,sumMode=1,elemMode=1,nTrials=nT,nSuccesses=nS,dump=True)
  with Engine(nParticles=N,resamplingScheme=1,retainPopulations=1) as e:
    e.set_tree(csr); e.set_multilevel_model(nT,nS); e.run()
    st,_,lw,anc=e.debug_node(0)
    bad=np.nonzero(st[0]!=ref.state[0][0])[0]
    print(var,(n,s),'bad',len(bad),bad[:8], [b%32 for b in bad[:8]])
