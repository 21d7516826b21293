import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, dcsmc_b200 as d, oracle_lib as o
from dcsmc_b200.engine import Engine
L=o.lib()
tree=d.DirectedTree(d.Node(("R",))); csr=tree.to_csr()
for (n,s) in [(1,1),(12,0),(9,9),(30,11)]:
  nT=np.array([n],np.int32); nS=np.array([s],np.int32)
  N=65536
  ref=o.run(csr,N,resamplingScheme=1,rngMode=o.RNG_PHILOX,sumMode=1,elemMode=1,nTrials=nT,nSuccesses=nS,dump=True)
  with Engine(nParticles=N,resamplingScheme=1,retainPopulations=1) as e:
    e.set_tree(csr); e.set_multilevel_model(nT,nS); e.run()
    st,_,lw,anc=e.debug_node(0)
    bad=np.nonzero(st[0]!=ref.state[0][0])[0]
    print((n,s),'philox-mode bad',len(bad),bad[:8], st[0][bad[:3]], ref.state[0][0][bad[:3]])
  idx=np.arange(N*64+N,dtype=np.float64)
  with Engine(nParticles=10) as e2:
    u=e2.debug_eval(2,idx)
  with Engine(nParticles=N,resamplingScheme=1,rngMode=1,retainPopulations=1) as e:
    e.set_tree(csr); e.set_multilevel_model(nT,nS); e.inject_uniforms(0,u); e.run()
    st,_,lw,anc=e.debug_node(0)
    bad=np.nonzero(st[0]!=ref.state[0][0])[0]
    print((n,s),'injected-mode bad',len(bad),bad[:8])
