import sys
nP=16
COV=16e3; TRSM=3e3; GK=75.0*32
FACT=29e3; INV=9.5e3; WS=7e3
DT=500.0
def pair(K, trsm=True): return COV + (TRSM if trsm else 0) + GK*K

def dataflow(R, helpers=True, fact=FACT, nchain=2, verbose=False):
    tasks=[(p,rp) for p in range(nP-1) for rp in range(2*(p+2), 2*nP)]
    tot={p:sum(1 for t in tasks if t[0]==p) for p in range(nP)}
    cnt={p:0 for p in range(nP)}
    fin=set()            # completed (rp,p)
    minv=set()           # blocks with inverse ready
    factored=set()
    nxt=[0]
    # chain state per block b: step 0 prio, 1 LA, 2 factor/inverse
    cb=[0]; prio_done=[0]; la_done=[0]; stage=['prio']
    busy=[0.0]*4; act=[None]*4
    t=0.0
    idle=[0.0]*4
    def pool_ready():
        if nxt[0]>=len(tasks): return None
        p,rp=tasks[nxt[0]]
        if p not in minv: return None
        if p>0 and (rp,p-1) not in fin: return None
        return (p,rp)
    def chain_deps_ok(b):
        if b==0: return True
        if (b-1) not in minv: return False
        if b>=2 and (((2*b,b-2) not in fin) or ((2*b+1,b-2) not in fin)): return False
        if b-R>=0 and cnt[b-R]<tot[b-R]: return False
        return True
    # chain progress tracked with per-warp flags
    st={'b':0,'phase':'prio','w_done':[False,False],'w_started':[False,False],'fact_started':False,'fact_done':False,'inv_started':False,'w1_pool_taken':False}
    events=[]
    while True:
        # completion of actions
        for w in range(4):
            if act[w] is not None and busy[w]<=t+1e-9:
                a=act[w]; act[w]=None
                if a[0]=='pool': fin.add((a[2],a[1])); cnt[a[1]]+=1
                elif a[0]=='prio':
                    st['w_done'][w]=True
                    if all(st['w_done']):
                        b=st['b']
                        if b>0: fin.add((2*b,b-1)); fin.add((2*b+1,b-1))
                        st['phase']='la'; st['w_done']=[False,False]; st['w_started']=[False,False]
                elif a[0]=='la':
                    st['w_done'][w]=True
                    if all(st['w_done']):
                        st['phase']='fact'; st['fact_started']=False; st['fact_done']=False; st['inv_started']=False; st['w1_pool_taken']=False
                elif a[0]=='fact':
                    st['fact_done']=True
                elif a[0]=='inv':
                    minv.add(st['b']); st['b']+=1; st['phase']='prio'; st['w_done']=[False,False]; st['w_started']=[False,False]
                elif a[0]=='ws': pass
        if st['b']>=nP and nxt[0]>=len(tasks) and all(a is None for a in act): break
        for w in range(4):
            if act[w] is not None: continue
            done=False
            if w<2 and st['b']<nP:
                b=st['b']
                if st['phase']=='prio' and not st['w_started'][w] and chain_deps_ok(b):
                    st['w_started'][w]=True
                    dur=pair(b-1) if b>0 else 0.0
                    act[w]=('prio',); busy[w]=t+dur; done=True
                elif st['phase']=='la' and not st['w_started'][w]:
                    st['w_started'][w]=True; act[w]=('la',); busy[w]=t+pair(b,False); done=True
                elif st['phase']=='fact':
                    if w==0 and not st['fact_started']:
                        st['fact_started']=True; act[w]=('fact',); busy[w]=t+fact; done=True
                    elif w==1 and not st['inv_started']:
                        if st['fact_done']:
                            st['inv_started']=True; act[w]=('inv',); busy[w]=t+INV; done=True
                        elif not st['w1_pool_taken']:
                            pr=pool_ready()
                            st['w1_pool_taken']=True
                            if pr:
                                nxt[0]+=1; act[w]=('pool',pr[0],pr[1]); busy[w]=t+pair(pr[0]); done=True
                        else:
                            done=True; idle[w]+=DT   # waiting for factor
                    elif w==0 and st['fact_started'] and not st['fact_done']:
                        pass
            if done: continue
            # chain warp waiting for the other chain warp (barrier) -> cannot take long pool tasks? allow if helpers
            if w<2 and st['b']<nP:
                # waiting in chain: either deps not ok (help pool) or waiting partner/inverse
                b=st['b']
                waiting_partner = (st['phase'] in ('prio','la') and st['w_started'][w]) or (st['phase']=='fact' and w==0 and st['fact_done'] is False and st['fact_started'])
                if st['phase']=='fact' and w==0 and st['fact_done']:
                    # w0 after factor: w-solve then free to help pool until inverse finishes
                    pass
                if st['phase']=='prio' and not st['w_started'][w] and not chain_deps_ok(b):
                    if helpers:
                        pr=pool_ready()
                        if pr: nxt[0]+=1; act[w]=('pool',pr[0],pr[1]); busy[w]=t+pair(pr[0]); continue
                    idle[w]+=DT; continue
                if st['phase']=='fact' and w==0 and st['fact_done']:
                    if helpers:
                        pr=pool_ready()
                        if pr: nxt[0]+=1; act[w]=('pool',pr[0],pr[1]); busy[w]=t+pair(pr[0]); continue
                    idle[w]+=DT; continue
                idle[w]+=DT; continue
            pr=pool_ready()
            if pr: nxt[0]+=1; act[w]=('pool',pr[0],pr[1]); busy[w]=t+pair(pr[0])
            else: idle[w]+=DT
        t+=DT
        if t>1e7: raise RuntimeError('stuck b=%d nxt=%d'%(st['b'],nxt[0]))
    return t, idle
for R in (1,2,3,4,5,8,16):
    T,idle=dataflow(R)
    print('R=%2d: %.0fk idle %s'%(R,T/1e3,[int(x/1e3) for x in idle]))
T,idle=dataflow(16,fact=15e3); print('R=16 fast factor: %.0fk'%(T/1e3), [int(x/1e3) for x in idle])
print('no helpers')
for R in (2,3,5,16):
    T,idle=dataflow(R,helpers=False)
    print('R=%2d: %.0fk idle %s'%(R,T/1e3,[int(x/1e3) for x in idle]))
