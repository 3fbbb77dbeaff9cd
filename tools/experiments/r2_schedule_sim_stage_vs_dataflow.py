import heapq, sys
nP=16
COV=16e3; TRSM=3e3; GK=75.0*32   # per panel of K
FACT=29e3; INV=9.5e3; WS=7e3
def pair(K, trsm=True): return COV + (TRSM if trsm else 0) + GK*K

def stage_sync():
    # current schedule: per stage barrier
    T=0
    for jc in range(nP):
        j=jc-1
        npool = (28-2*j) if j>=0 else 0
        # chain: warps 0,1
        t0 = pair(j) if j>=0 else 0      # priority
        tla = t0 + pair(jc, False)
        w0 = tla + FACT + WS
        # warp1: after LA takes one pool pair (if any) then inverse after factor
        pool = list(range(npool))
        free=[0,0,0,0]
        free[2]=free[3]=0
        free[0]=w0; 
        # warp 1: pool pair then inverse
        tw1 = tla
        if pool:
            pool.pop(0); tw1 += pair(j)
        tinv = max(tw1, tla+FACT) + INV
        free[1]=tinv
        # list-schedule remaining pool tasks on earliest free warp
        h=[(free[i],i) for i in range(4)]; heapq.heapify(h)
        while pool:
            t,i=heapq.heappop(h); pool.pop(0); heapq.heappush(h,(t+pair(j),i))
        T += max(t for t,i in h)
    return T

def dataflow(R):
    # event-driven: 4 warps; warps 0,1 chain-capable; tasks: pool (p, rp) ordered; deps as described
    # state
    import math
    rowdone = {}   # (rp) -> finish time of panel p for row pair : dict (rp,p)->time
    fin = {}       # (rp,p) -> finish time
    minv = {}      # block -> time inverse ready
    # We simulate with a time-stepped greedy: each warp has 'free at' time; chain progress is a state machine.
    # Simplify: chain handled by warps 0 and 1 symmetrically: the chain step for block b consists of
    #  prio (both warps in parallel, needs minv[b-1], rows b panels<b-1 done), LA (both), factor (w0), inverse (w1 after factor).
    # Pool tasks list in order.
    tasks=[(p,rp) for p in range(nP-1) for rp in range(2*(p+2), 2*nP)]
    nxt=0
    free=[0.0]*4
    chain_b=0
    chain_ready_time=0.0  # when chain warps are both done with the previous block's chain
    pool_panel_done={}    # p -> time all pool tasks of panel p complete
    cnt={p:0 for p in range(nP)}
    tot={p:len([1 for t in tasks if t[0]==p]) for p in range(nP)}
    lastfin={p:0.0 for p in range(nP)}
    def dep_time_pool(p,rp):
        t=minv.get(p, None)
        if t is None: return None
        if p>0:
            f=fin.get((rp,p-1))
            if f is None: return None
            t=max(t,f)
        return t
    def chain_dep(b):
        # rows of block b (rp 2b, 2b+1) in panels 0..b-2 done; minv[b-1]; ring: pool panel b-R complete
        t=0.0
        if b>0:
            if (b-1) not in minv: return None
            t=minv[b-1]
            if b>=2:
                for rp in (2*b,2*b+1):
                    f=fin.get((rp,b-2))
                    if f is None: return None
                    t=max(t,f)
        if b-R>=0 and tot[b-R]>0:
            if cnt[b-R]<tot[b-R]: return None
            t=max(t,lastfin[b-R])
        return t
    # Discrete event loop: repeatedly pick the warp with the earliest free time and assign work
    guard=0
    # chain state: we treat chain block as atomic two-warp job started when both warps 0,1 are free & deps ok
    while True:
        guard+=1
        if guard>100000: raise RuntimeError
        if chain_b>=nP and nxt>=len(tasks): break
        # try chain first if deps known
        progressed=False
        if chain_b<nP:
            d=chain_dep(chain_b)
            if d is not None:
                # both chain warps needed: start at max(free0, free1, d) -- but they may do pool tasks in between; decide: if a chain warp is free earlier than d by more than a pool task, let it take pool tasks (handled below by order of events)
                s=max(free[0],free[1],d)
                # check whether some pool task could be run by a chain warp before s: handled by greedy below: only commit chain if no warp is free before s-eps with a ready pool task
                cand=None
                if nxt<len(tasks):
                    p,rp=tasks[nxt]; dp=dep_time_pool(p,rp)
                    if dp is not None:
                        for i in range(4):
                            st=max(free[i],dp)
                            if i<2 and st+pair(p) > s+1e-9 and st < s: continue  # chain warp: don't start a pool task that would delay the chain
                            if st < s-1e-9 and (i>=2 or st+pair(p)<=s+1e-9):
                                if cand is None or st<cand[0]: cand=(st,i)
                if cand is None:
                    b=chain_b
                    t0=s + (pair(b-1) if b>0 else 0)
                    tla=t0+pair(b,False)
                    tf=tla+FACT
                    free[0]=tf+WS
                    # warp 1: may take one ready pool task between tla and tf
                    tw1=tla
                    if nxt<len(tasks):
                        p,rp=tasks[nxt]; dp=dep_time_pool(p,rp)
                        if dp is not None and dp<=tla:
                            nxt+=1; tw1=tla+pair(p); fin[(rp,p)]=tw1; cnt[p]+=1; lastfin[p]=max(lastfin[p],tw1)
                    ti=max(tw1,tf)+INV
                    free[1]=ti
                    minv[b]=ti
                    # priority rows: rows of block b at panel b-1 finished at t0; LA = panel b for rows b (diag)
                    if b>0:
                        fin[(2*b,b-1)]=t0; fin[(2*b+1,b-1)]=t0
                    chain_b+=1
                    progressed=True
                    continue
        if nxt<len(tasks):
            p,rp=tasks[nxt]; dp=dep_time_pool(p,rp)
            if dp is not None:
                # earliest warp; chain warps only if chain not ready
                best=None
                for i in range(4):
                    st=max(free[i],dp)
                    if best is None or st<best[0]: best=(st,i)
                st,i=best
                nxt+=1
                f=st+pair(p); free[i]=f; fin[(rp,p)]=f; cnt[p]+=1; lastfin[p]=max(lastfin[p],f)
                progressed=True
                continue
        if not progressed:
            raise RuntimeError('deadlock chain_b=%d nxt=%d'%(chain_b,nxt))
    return max(free)

W = sum(pair(p) for p in range(nP-1) for rp in range(2*(p+2),2*nP)) + sum((pair(b-1) if b>0 else 0)*2 + 2*pair(b,False) + FACT+WS+INV for b in range(nP))
print('work/4 = %.0fk'%(W/4e3))
print('stage-sync: %.0fk'%(stage_sync()/1e3))
for R in (1,2,3,4,5,6,8,16):
    print('dataflow R=%d: %.0fk'%(R, dataflow(R)/1e3))
