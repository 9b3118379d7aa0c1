import sys, numpy as np
sys.path.insert(0,'.')
from ttcr_b200 import Grid3d
shape=tuple(int(a) for a in sys.argv[1:4]); opts=dict(kv.split('=') for kv in sys.argv[4:])
rng=np.random.default_rng(0)
x,y,z=(np.arange(m)*0.25 for m in shape)
s=rng.uniform(0.3,1.0,shape)
src=np.array([[min(3.3,x[-1]),min(2.2,y[-1]),min(9.1,z[-1])]])
out=[]
for k in (1,3):
    g=Grid3d(x,y,z,cell_slowness=0,tt_from_rp=False,weno=0,dtype=np.float32)
    g.set_option("kernel",k)
    if k==3:
        for kk,v in opts.items(): g.set_option(kk,float(v))
    g.raytrace(src,src,s); out.append(g.get_grid_traveltimes()); print(k,g.get_niter(),g.get_stats()['solve_ms'],flush=True)
print('equal',np.array_equal(out[0],out[1]), np.abs(out[0]-out[1]).max())
