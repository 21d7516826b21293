import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, dcsmc_b200 as d, oracle_lib as o
from dcsmc_b200.engine import Engine
L=o.lib()
with Engine(nParticles=10) as e:
    for node in [16, 0, 3]:
        idx=np.arange(0, 400000, dtype=np.float64)
        yd=e.debug_eval(2|(node<<8), idx)
        yo=np.array([L.orc_philox_uniform(31,node,int(i)) for i in idx])
        bad=np.nonzero(yd!=yo)[0]
        print('node',node,'philox mismatches',len(bad),bad[:10])
