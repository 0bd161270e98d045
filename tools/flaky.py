import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parallel_eda_b200 import pfio, router, check_route
G='tests/golden'
import sys
for KN in (16,64,256):
  print('inflight_div',KN)
  for name in ['hub_w90','toy_w64','mid_w200']:
      p=pfio.read_problem('%s/%s.pfp.xz'%(G,name)); p.opts['timing_analysis_enabled']=1; p.opts['max_router_iterations']=150
      g=pfio.read_result('%s/%s.pfr.xz'%(G,name)); w=g.iter_crit[-1]
      for t in range(5 if name!='mid_w200' else 2):
          r=router.try_timing_driven_route(p, router.default_config(inflight_div=KN), sta=router.replay_sta(g))
          print(name,'success',r.success,'it',r.iterations,'(ref %d)'%g.iterations,'wl %.3f'%(r.total_wirelength/g.total_wirelength),'wdelay %.3f'%(float((w*r.net_delay).sum())/float((w*g.net_delay).sum())), 'over', [int(x) for x in r.iter_stats['overused_nodes']][-6:], flush=True)
