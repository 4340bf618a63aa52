import numpy as np
rng=np.random.default_rng(1)
dp=1.0; h=np.sqrt(3)*dp; cs=2*h
# lattice of particles with small jitter in a 5x5x5-cell region; evaluate central cell's warps
n=int(np.ceil(5*cs/dp))
g=np.arange(n)*dp+0.01
P=np.stack(np.meshgrid(g,g,g,indexing='ij'),-1).reshape(-1,3)  # order: x slowest here; fix below
# id order: x fastest, then y, then z
ids=np.lexsort((P[:,0],P[:,1],P[:,2]))
P=P[ids]+rng.uniform(-0.15,0.15,(len(ids),3))*dp
cell=np.floor(P/cs).astype(int)
def cellparts(c):
    m=np.all(cell==c,axis=1); return np.nonzero(m)[0]   # in id order
def run(Q,policy,chunk=8):
    tot_rounds=0; tot_active=0; tot_hits=0; nwarps=0
    for c in [(2,2,2),(1,2,2),(2,1,2),(2,2,1)]:
        mine=cellparts(np.array(c))
        for w0 in range(0,len(mine)-31,32):
            lanes=mine[w0:w0+32]
            cnt=np.zeros(32,int); rounds=0; active=0
            cand=[]
            for dz in(-1,0,1):
                for dy in(-1,0,1):
                    for dx in(-1,0,1):
                        cand.append(cellparts(np.array(c)+np.array([dx,dy,dz])))
            cand=np.concatenate(cand)
            d2=((P[lanes][:,None,:]-P[cand][None,:,:])**2).sum(-1)
            hit=(d2<=cs*cs)&(d2>1e-12)   # [32,ncand]
            for i0 in range(0,hit.shape[1],chunk):
                cnt+=hit[:,i0:i0+chunk].sum(1)
                if policy=='one':
                    while cnt.max()>Q-chunk:
                        a=(cnt>0); rounds+=1; active+=a.sum(); cnt[a]-=1
                elif policy=='lowwater':
                    if cnt.max()>Q-chunk:
                        while cnt.max()>Q-chunk or (cnt>0).sum()>=28:
                            a=(cnt>0); rounds+=1; active+=a.sum(); cnt[a]-=1
                elif policy=='dense+full':
                    while (cnt>0).sum()>=30 or cnt.max()>Q-chunk:
                        a=(cnt>0); rounds+=1; active+=a.sum(); cnt[a]-=1
            while cnt.max()>0:
                a=(cnt>0); rounds+=1; active+=a.sum(); cnt[a]-=1
            tot_rounds+=rounds; tot_active+=active; tot_hits+=hit.sum(); nwarps+=1
    return tot_rounds/nwarps, tot_active/tot_rounds, tot_hits/nwarps/32
for Q in (8,16,32,64,128):
    for pol in ('one','lowwater','dense+full'):
        r,u,hh=run(Q,pol)
        print('Q=%3d %-11s rounds/warp %.0f  active lanes %.1f  hits/p1 %.0f'%(Q,pol,r,u,hh))
