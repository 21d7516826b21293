import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, math, oracle_lib as o
L=o.lib()
target=1/(1+math.exp(-0.72177958)); w=target/(1-target); u1t=w/(w+2)
print('target p',target,'u1',u1t)
f=L.orc_philox_uniform
for node in [16]:
  for pair in range(0, 5000*32):
    u1=f(31,node,2*pair)
    if abs(u1-u1t)<3e-7: print('node',node,'pair',pair,'particle',pair//32,'att',pair%32,u1)
