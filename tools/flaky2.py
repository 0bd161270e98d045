import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import pfio, router
G='tests/golden'
for name in ['hub_w90','toy_w64','mid_w200']:
    p=pfio.read_problem('%s/%s.pfp.xz'%(G,name)); p.opts['timing_analysis_enabled']=0; p.opts['max_router_iterations']=150
    g=pfio.read_result('%s/%s_nt.pfr.xz'%(G,name))
    for t in range(8 if name!='mid_w200' else 3):
        r=router.try_timing_driven_route(p, router.default_config())
        print(name,'success',r.success,'it',r.iterations,'(ref %d)'%g.iterations,'wl %.3f'%(r.total_wirelength/g.total_wirelength), 'over', [int(x) for x in r.iter_stats['overused_nodes']][-5:], flush=True)
