import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import so, random_hypers
from gladsgp_b200 import ops, synthetic
# cfg5 shape: m=4096, q=16 (d=17); 2 matrices vs oracle
m,q=4096,16
t=synthetic.design(m,q)
X=np.concatenate([0.5*np.ones((m,1)), t.astype(np.float64)],axis=1)
rng=np.random.default_rng(0)
B=2
beta=np.exp(rng.uniform(np.log(0.05),np.log(1.0),size=(B,q+1))); lamz=np.array([0.9,1.4]); dadd=np.array([2e-3,5e-3])
W=rng.standard_normal((B,m))
out=ops.loglik_batched(X,W,beta,lamz,dadd); torch.cuda.synchronize()
t0=time.time(); out=ops.loglik_batched(X,W,beta,lamz,dadd); ll=out['loglik'].cpu().numpy(); print('gpu seconds (2 matrices, warm)',time.time()-t0, ll, out['info'].cpu().numpy())
import scipy.linalg
for b in range(B):
    D=((X[:,None,:]-X[None,:,:])**2)@beta[b]
    C=np.exp(-D)/lamz[b]; np.fill_diagonal(C,1/lamz[b]+dadd[b])
    L=scipy.linalg.cholesky(C,lower=True); u=scipy.linalg.solve_triangular(L,W[b],lower=True)
    ref=-np.sum(np.log(np.diag(L)))-0.5*u@u
    print(b, ll[b], ref, abs(ll[b]-ref)/abs(ref))
