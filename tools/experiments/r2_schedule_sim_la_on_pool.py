import heapq
nP=16
COV=16e3; TRSM=3e3; GK=75.0*32
FACT=29e3; INV=9.5e3; WS=7e3
def pair(K, trsm=True): return COV + (TRSM if trsm else 0) + GK*K
def sched(pool_n, j, free):
    h=[(free[i],i) for i in range(4)]; heapq.heapify(h)
    for _ in range(pool_n):
        t,i=heapq.heappop(h); heapq.heappush(h,(t+pair(j),i))
    return max(t for t,i in h), sum(max(t for t,i in h)-t for t,i in h)
def cur():
    T=0; idle=0
    for jc in range(nP):
        j=jc-1; npool=(28-2*j) if j>=0 else 0
        t0=pair(j) if j>=0 else 0
        tla=t0+pair(jc,False)
        w0=tla+FACT+WS
        tw1=tla
        if npool: npool-=1; tw1+=pair(j)
        tinv=max(tw1,tla+FACT)+INV
        L,i=sched(npool,j,[w0,tinv,0,0]); T+=L; idle+=i
    return T, idle
def new():
    T=0; idle=0
    for jc in range(nP):
        j=jc-1; npool=(28-2*j) if j>=0 else 0
        if j<0:
            tla=pair(0,False); w0=tla+FACT+WS; tinv=tla+FACT+INV
            L,i=sched(0,0,[w0,tinv,tla,tla]); T+=L; idle+=i; continue
        tprio=pair(j)
        tla=max(tprio, GK*j)+GK*1+COV
        w0=tla+FACT+WS
        tw1=tprio
        # warp1 pool tasks until factor done: takes tasks while start < tla (one at a time)
        if npool: npool-=1; tw1+=pair(j)
        tinv=max(tw1,tla+FACT)+INV
        L,i=sched(npool,j,[w0,tinv,tla,tla]); T+=L; idle+=i
    return T, idle
a=cur(); b=new()
print('current  %.0fk idle %.0fk'%(a[0]/1e3,a[1]/1e3)); print('LA on pool warps %.0fk idle %.0fk'%(b[0]/1e3,b[1]/1e3))
