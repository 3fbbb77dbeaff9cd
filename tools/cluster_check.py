import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
from helpers import so, make_problem, random_hypers
from gladsgp_b200 import ops, synthetic
for (m,q,pu) in [(64,3,2),(100,8,5),(512,8,3)]:
    pr = make_problem(m=m,q=q,pu=pu); num=pr['num']
    B=pu; beta,lamz,lamws,lamwos = random_hypers(num,B,seed=m); js=np.arange(B)%pu
    dadd = 1.0/(num.LamSim[js]*lamwos)+1.0/lamws
    W=np.stack([num.wv[j*m:(j+1)*m,0] for j in js])
    out=ops.loglik_batched(num.zt,W,beta,lamz,dadd,want_factor=True,want_u=True)
    torch.cuda.synchronize()
    ll=out['loglik'].cpu().numpy(); Lg=ops.factor_unpack(out['factor'],m).cpu().numpy(); ug=out['u'].cpu().numpy()[:,:m]
    for b in range(B):
        C=so.block_cov(num,beta[b],lamz[b],lamws[b],lamwos[b],js[b]); ref=so.do_loglik(C,W[b])
        L=np.linalg.cholesky(C); u=np.linalg.solve(L,W[b])
        print(m,b,'rel err',abs(ll[b]-ref)/abs(ref),'L',np.abs(Lg[b]-L).max(),'u',np.abs(ug[b]-u).max(), out['info'].cpu().numpy()[b])
# timing single matrix / 10 matrices at m=512
m,q=512,8; d=q+1
t=synthetic.design(m,q); X=np.concatenate([0.5*np.ones((m,1)),t.astype(np.float64)],axis=1)
rng=np.random.default_rng(0)
for B in (1,10,40):
    beta=np.exp(rng.uniform(np.log(0.05),np.log(3.0),size=(B,d))); lamz=rng.uniform(0.5,2.0,B); dadd=rng.uniform(1e-3,1e-2,B); W=rng.standard_normal((B,m))
    Xd,Wd,bd,ld,dd=[torch.as_tensor(a,device='cuda') for a in (X,W,beta,lamz,dadd)]
    ws=torch.empty((B,ops._lib.load().ggp_factor_doubles(m)),dtype=torch.float64,device='cuda')
    for _ in range(3): ops.loglik_batched(Xd,Wd,bd,ld,dd,factor_ws=ws)
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.loglik_batched(Xd,Wd,bd,ld,dd,factor_ws=ws)
    b.record(); torch.cuda.synchronize(); print('B',B,'ms per launch',a.elapsed_time(b)/10)
