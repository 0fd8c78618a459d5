import sys, time
sys.path.insert(0,'/root/repo/oracle')
import numpy as np, oracle
def load(tag):
    z=np.load('/root/repo/tests/golden/rocksample_%s.npz'%tag)
    return oracle.Model(int(z['rows']),int(z['cols']),z['cell'],int(z['start_cell']),z['thr53'],z['rewards']), z
def episode_gs(om, z, seed, n_sims, w0, growth, wmax, max_steps=200, discount=0.95):
    rng=np.random.default_rng(seed)
    K=om.n_rocks
    actual=int(rng.integers(0,1<<K))
    bag=(rng.integers(0,1<<K,size=2000).astype(np.uint32)<<8)|np.uint32(om.start_cell)
    g=oracle.GsPOMCP(om,actual,0,bag,seed=seed)
    cell=om.start_cell; rocks=int(bag[rng.integers(0,2000)]>>8); sampled=0
    ret=0.0; disc=1.0
    for mv in range(max_steps):
        g.search(n_sims,move=mv,w0=w0,growth=growth,wmax=wmax)
        a=g.greedy(move=mv)
        o=om.step(cell,rocks,a,int(rng.integers(0,1<<53)),actual,sampled)
        ret+=disc*o.reward; disc*=discount
        cell,rocks=o.next_cell,o.next_rocks
        if o.terminal or not o.legal: break
        if a==4: sampled|=1<<int(z['cell'][cell])
        u=g.update(a,o.obs_good,sampled,move=mv)
        if u['status']==2: print("update failed"); break
    return ret, mv+1
def episode_ref(om, z, seed, n_sims, max_steps=200, discount=0.95, quirks=(1,1,1)):
    rng=np.random.default_rng(seed)
    K=om.n_rocks
    actual=int(rng.integers(0,1<<K))
    bagr=rng.integers(0,1<<K,size=2000).astype(np.uint32)
    r=oracle.RefPOMCP(om,actual,0,bagr,seed=seed)
    r.set_quirks(*quirks)
    rocks=int(bagr[rng.integers(0,2000)]); cell=om.start_cell; sampled=0
    ret=0.0; disc=1.0
    for mv in range(max_steps):
        a=r.search(n_sims)
        o=om.step(cell,rocks,a,int(rng.integers(0,1<<53)),actual,sampled)
        ret+=disc*o.reward; disc*=discount
        cell,rocks=o.next_cell,o.next_rocks
        if o.terminal or not o.legal: break
        if a==4: sampled|=1<<int(z['cell'][cell])
        rc=r.update(a,o.obs_good,sampled)
        if rc<0: print("ref update rc",rc); break
    return ret, mv+1
if __name__=="__main__":
    tag=sys.argv[1]; n_sims=int(sys.argv[2]); neps=int(sys.argv[3])
    om,z=load(tag)
    for name,fn in [("ref",lambda s:episode_ref(om,z,s,n_sims)),
                    ("gs W=1",lambda s:episode_gs(om,z,s,n_sims,1,1,1)),
                    ("gs 8x2..64",lambda s:episode_gs(om,z,s,n_sims,8,2,64)),
                    ("gs 32 flat",lambda s:episode_gs(om,z,s,n_sims,32,1,32)),
                    ("gs 128 flat",lambda s:episode_gs(om,z,s,n_sims,128,1,128))]:
        t=time.time(); R=[fn(1000+i) for i in range(neps)]
        r=np.array([x[0] for x in R]); st=np.array([x[1] for x in R])
        print("%-12s return %.2f +- %.2f  steps %.1f  (%.1fs)"%(name,r.mean(),r.std()/np.sqrt(len(r)),st.mean(),time.time()-t),flush=True)
