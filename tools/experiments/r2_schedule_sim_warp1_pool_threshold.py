import heapq
nP=16
COV=16e3; TRSM=3e3; GK=75.0*32
FACT=29e3; INV=9.5e3; WS=7e3
def pair(K, trsm=True): return COV + (TRSM if trsm else 0) + GK*K
def sched(pool_n, j, free):
    h=[(free[i],i) for i in range(4)]; heapq.heapify(h)
    for _ in range(pool_n):
        t,i=heapq.heappop(h); heapq.heappush(h,(t+pair(j),i))
    return max(t for t,i in h)
def cur(T):
    tot=0
    for jc in range(nP):
        j=jc-1; npool=(28-2*j) if j>=0 else 0
        t0=pair(j) if j>=0 else 0
        tla=t0+pair(jc,False)
        w0=tla+FACT+WS
        tw1=tla
        if npool and npool>=T: npool-=1; tw1+=pair(j)
        tinv=max(tw1,tla+FACT)+INV
        tot+=sched(npool,j,[w0,tinv,0,0])
    return tot
for T in (0,1,2,4,6,8,12,100): print(T, '%.0fk'%(cur(T)/1e3))
