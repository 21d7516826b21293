import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, dcsmc_b200 as d, oracle_lib as o
from dcsmc_b200 import synth
from dcsmc_b200.engine import Engine
ds = d.MultiLevelDataset(rows=synth.small_multilevel_rows(seed=7)); csr = ds.getTree().to_csr(); nT,nS=ds.leaf_arrays(csr)
scheme=1
for N in [2000,4096,5000]:
    ref = o.run(csr, N, masterRandomSeed=31, resamplingScheme=scheme, rngMode=o.RNG_PHILOX, sumMode=o.SUM_EXACT, elemMode=o.ELEM_FDLIBM, nTrials=nT, nSuccesses=nS, dump=True)
    with Engine(nParticles=N, masterRandomSeed=31, resamplingScheme=scheme, retainPopulations=1) as e:
        e.set_tree(csr); e.set_multilevel_model(nT,nS); e.run()
        for n in [13,14,15,16]:
            st,ist,lw,anc=e.debug_node(n)
            for col in range(6):
                bad=np.nonzero(~((st[col]==ref.state[n][col])|(np.isnan(st[col])&np.isnan(ref.state[n][col]))))[0]
                if len(bad): print(N,'node',n,'col',col,'bad',len(bad),bad[:6],st[col][bad[:3]],ref.state[n][col][bad[:3]])
            bad=np.nonzero(lw!=ref.logw[n])[0]
            if len(bad): print(N,'node',n,'lw bad',len(bad),bad[:6],lw[bad[:3]],ref.logw[n][bad[:3]])
