import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parallel_eda_b200 import pfio, router
p=pfio.read_problem('tests/golden/toy_w64.pfp.xz'); p.opts['timing_analysis_enabled']=1
g=pfio.read_result('tests/golden/toy_w64.pfr.xz'); sta=router.replay_sta(g)
names=['SRC','SINK','IPIN','OPIN','CHX','CHY']
for trial in range(6):
    R=router.Router(p, router.default_config()); o=p.opts
    pres=float(o['first_iter_pres_fac']); crit=None; hist=[]
    prev_ov=None; same=0
    for it in range(1,51):
        st=R.route_iteration(pres,crit); R.reserve_locally_used_opins(pres,it!=1)
        if it==1: pres=float(o['initial_pres_fac']); acc=0.0
        else: pres=min(pres*float(o['pres_fac_mult']),1e25); acc=float(o['acc_fac'])
        over=R.pathfinder_update_cost(acc); hist.append((st.nets_routed,over))
        if over==0: break
        crit,_=sta(it,R.net_delay())
        if over<=2 and it>=20:
            res=R.result(); ov=np.nonzero(res.occ>p.capacity)[0]
            desc=[]
            for v in ov:
                users=[int(i) for i in p.routed_nets() if v in set(res.net_trace(int(i))[0].tolist())]
                desc.append((int(v),names[p.type[v]],int(p.xlow[v]),int(p.ylow[v]),int(p.xhigh[v]),int(p.yhigh[v]),int(res.occ[v]),users))
            print('  it',it,'nets',st.nets_routed,'overused',desc,flush=True)
    print('trial',trial,'iters',len(hist),hist[-12:],flush=True)
    R.close()
