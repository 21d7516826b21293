import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, dcsmc_b200 as d, oracle_lib as o
from dcsmc_b200.engine import Engine
L=o.lib()
rng=np.random.default_rng(5)
with Engine(nParticles=10) as e:
    xs=np.concatenate([rng.uniform(-745,709,200000), rng.uniform(-2,2,200000), rng.normal(0,1e-3,50000), rng.normal(0,1e-9,50000), rng.uniform(-1.1,1.1,200000), np.array([0.0,-0.0,1e-300,-1e-300,709.78,-745.13,-746,710,np.inf,-np.inf, 0.34657359027997264, 0.3465735902799727, -0.34657359027997264, 1.0397207708399179,1.03972077083991,-1.0397207708399179])])
    yd=e.debug_eval(1,xs); yo=np.array([L.orc_elem_exp(1,float(x)) for x in xs])
    bad=np.nonzero(~((yd==yo)|(np.isnan(yd)&np.isnan(yo))))[0]
    print('exp mismatches',len(bad)); 
    for b in bad[:10]: print(repr(xs[b]),repr(yd[b]),repr(yo[b]))
    ys=np.concatenate([np.exp(rng.uniform(-700,700,200000)), rng.uniform(0.5,2.0,200000), rng.uniform(0,1,200000), 1+rng.normal(0,1e-7,50000),1+rng.normal(0,1e-3,50000), np.array([0.0,1.0,5e-324,1e-310,2.0,0.5,np.inf,-1.0,0.7071067811865476,1.4142135623730951,0.70710678,1.41421356])])
    yd=e.debug_eval(0,ys); yo=np.array([L.orc_elem_log(1,float(x)) for x in ys])
    bad=np.nonzero(~((yd==yo)|(np.isnan(yd)&np.isnan(yo))))[0]
    print('log mismatches',len(bad))
    for b in bad[:10]: print(repr(ys[b]),repr(yd[b]),repr(yo[b]))
