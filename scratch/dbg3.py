import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, math, oracle_lib as o
L=o.lib()
node=0; i=17099; maxAtt=32
a0=1.0; a=13.0; b=1.0; alpha=14.0; beta=1.0; delta=13.0
k1 = delta * (0.0138889 + 0.0416667 * b) / (a * beta - 0.777778); k2 = 0.25 + (0.5 + 0.25 / delta) * b
LOG=lambda x: L.orc_elem_log(1,x); EXP=lambda x: L.orc_elem_exp(1,x)
for att in range(32):
    idx=(i*maxAtt+att)*2
    u1=L.orc_philox_uniform(31,node,idx); u2=L.orc_philox_uniform(31,node,idx+1)
    v=beta*(LOG(u1)-LOG(1.0-u1)); w=a*EXP(v); y=u1*u2; z=u1*y
    if u1<0.5:
        path='A-reject' if 0.25*u2+z-y>=k1 else 'A-final'
    else:
        path='B-accept' if z<=0.25 else ('B-reject' if z>=k2 else 'B-final')
    fin = alpha*(LOG(alpha)-LOG(b+w)+v)-1.3862944 >= LOG(z)
    print(att,round(u1,4),round(u2,4),path,fin,'p',b/(b+w), 'logit', math.log(b/(b+w))-math.log(1-b/(b+w)))
    if path=='B-accept' or (path.endswith('final') and fin): break
