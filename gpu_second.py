import sys, time, numpy as np
sys.path.insert(0,'.')
from parallel_eda_b200 import pfio, router, check_route

def run(p, cfg, sta=None, label=''):
    o=p.opts
    t0=time.time(); R=router.Router(p,cfg); t_create=time.time()-t0
    pres=float(o['first_iter_pres_fac']); crit=None; its=[]; t1=time.time()
    for it in range(1,int(o['max_router_iterations'])+1):
        ta=time.time(); st=R.route_iteration(pres,crit); tb=time.time()
        if it==1: R.total_wirelength()
        R.reserve_locally_used_opins(pres, it!=1)
        if it==1: pres=float(o['initial_pres_fac']); acc=0.0
        else: pres=min(pres*float(o['pres_fac_mult']),1e25); acc=float(o['acc_fac'])
        over=R.pathfinder_update_cost(acc); tc=time.time()
        its.append((st.nets_routed,over,tb-ta,tc-tb,st.heap_pops,st.edge_visits))
        if over==0: break
        if sta is not None:
            crit,_=sta(it,R.net_delay())
    t_route=time.time()-t1
    tm=R.timing(reset=True)
    t2=time.time(); res=R.result(); t_res=time.time()-t2
    res.success=int(its[-1][1]==0)
    ok='ok'
    try: check_route.check_route(p,res,require_legal=bool(res.success))
    except Exception as e: ok='CHECK FAILED %s'%e
    nr=sum(i[0] for i in its)
    print(label,'create %.3f route %.4f result %.3f | iters %d netroutes %d wl %d | kernel route %.2f ms (%d launches) update %.2f ms aux %.2f ms | nets/s %.0f | %s'%(
        t_create,t_route,t_res,len(its),nr,res.total_wirelength,tm.route_kernel_ms,tm.route_launches,tm.update_kernel_ms,tm.aux_kernel_ms,nr/t_route,ok),flush=True)
    print('    per-iter (nets,over,route_ms):',[(a,b,round(c*1e3,1)) for a,b,c,d,_,_ in its][:20],'pops',sum(i[4] for i in its),'visits',sum(i[5] for i in its),flush=True)
    R.close()
    return res

for name in ['mid_w200']:
    p=pfio.read_problem('tests/golden/%s.pfp.xz'%name); p.opts['timing_analysis_enabled']=0
    g=pfio.read_result('tests/golden/%s_nt.pfr.xz'%name)
    print('reference: iters',g.iterations,'wl',g.total_wirelength)
    run(p, router.default_config(), label='warmup')
    for ps in [0.0,0.25,0.5,1.0,2.0]:
        for mb in [4,16,32]:
            run(p, router.default_config(pop_slack=ps,max_batch=mb), label='slack=%.2f batch=%d'%(ps,mb))
    for div in [8,32]:
        run(p, router.default_config(pop_slack=0.5,max_batch=16,inflight_div=div), label='slack=.5 b16 div=%d'%div)
    p.opts['timing_analysis_enabled']=1
    gt=pfio.read_result('tests/golden/%s.pfr.xz'%name)
    print('reference timing: iters',gt.iterations,'wl',gt.total_wirelength)
    for ps in [0.0,0.5,1.0]:
        run(p, router.default_config(pop_slack=ps), sta=router.replay_sta(gt), label='timing slack=%.2f'%ps)
