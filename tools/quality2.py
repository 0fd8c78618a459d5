import sys,time
sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo/oracle')
from quality import *
tag=sys.argv[1]; n_sims=int(sys.argv[2]); neps=int(sys.argv[3])
om,z=load(tag)
cfgs=[("gs W=1",(1,1,1)),("gs 64x2..4096",(64,2,4096)),("gs 256x2..16384",(256,2,16384)),("gs 2048 flat",(2048,1,2048)),("gs 16384 flat",(16384,1,16384))]
for name,(w0,g,wm) in cfgs:
    t=time.time(); R=[episode_gs(om,z,1000+i,n_sims,w0,g,wm) for i in range(neps)]
    r=np.array([x[0] for x in R]); st=np.array([x[1] for x in R])
    print("%-16s return %.2f +- %.2f  steps %.1f  (%.1fs)"%(name,r.mean(),r.std()/np.sqrt(len(r)),st.mean(),time.time()-t),flush=True)
