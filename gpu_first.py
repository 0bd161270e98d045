import sys, time, numpy as np
sys.path.insert(0,'.')
from parallel_eda_b200 import pfio, router, check_route
for name in ['toy_w64','mid_w200']:
    p=pfio.read_problem('tests/golden/%s.pfp.xz'%name); p.opts['timing_analysis_enabled']=0
    g=pfio.read_result('tests/golden/%s_nt.pfr.xz'%name)
    for kw in [dict(), dict(num_slots=1), dict(reroute_all_iters=-1), dict(max_batch=1), dict(inflight_div=4)]:
        cfg=router.default_config(verbose=0, **kw)
        t=time.time(); r=router.try_timing_driven_route(p,cfg); dt=time.time()-t
        try:
            chk=check_route.check_route(p,r,require_legal=bool(r.success))
        except Exception as e:
            chk='CHECK FAILED: %s'%e
        print(name,kw,'time %.3f'%dt,'success',r.success,'iters',r.iterations,'wl',r.total_wirelength,'(ref %d it %d)'%(g.total_wirelength,g.iterations),
              'netroutes',int(r.iter_stats['nets_routed'].sum()),'pops',int(r.iter_stats['heap_pops'].sum()),'visits',int(r.iter_stats['edge_visits'].sum()), chk, flush=True)
        print('   overused',[int(x) for x in r.iter_stats['overused_nodes']][:30], flush=True)
