import sys, time, numpy as np
sys.path.insert(0,'.')
from parallel_eda_b200 import pfio, router, check_route
exec(open('gpu_second.py').read().split("for name in ['mid_w200']:")[0].split("from parallel_eda_b200 import pfio, router, check_route")[1])
import check_fast
for (nx,nn) in [(100,12500),(200,50000),(400,200000)]:
    t=time.time(); p=router.generate_grid_problem(nx=nx,ny=nx,W=100,num_nets=nn); print('gen %dx%d: %.1fs N=%d E=%d'%(nx,nx,time.time()-t,p.num_nodes,p.num_edges),flush=True)
    for kw in ([dict(), dict(num_slots=148*32, warps_per_block=8)] if nx<400 else [dict(), dict(num_slots=148*32, warps_per_block=8), dict(inflight_div=64)]):
        t=time.time()
        o=p.opts
        R=router.Router(p,router.default_config(verbose=1,**kw)); t_create=time.time()-t
        pres=float(o['first_iter_pres_fac']); its=[]; t1=time.time()
        for it in range(1,51):
            ta=time.time(); st=R.route_iteration(pres,None); tb=time.time()
            if it==1: wl=R.total_wirelength()
            if it==1: pres=float(o['initial_pres_fac']); acc=0.0
            else: pres=min(pres*float(o['pres_fac_mult']),1e25); acc=float(o['acc_fac'])
            over=R.pathfinder_update_cost(acc); tc=time.time()
            its.append((st.nets_routed,over,round((tb-ta)*1e3,1),st.heap_pops,st.edge_visits))
            if over==0: break
        t_route=time.time()-t1; tm=R.timing(reset=True)
        t2=time.time(); res=R.result(); t_res=time.time()-t2
        nr=sum(i[0] for i in its)
        print(kw,'create %.2f route %.3f result %.2f iters %d netroutes %d wl %d kernel %.1f ms nets/s %.0f'%(t_create,t_route,t_res,len(its),nr,res.total_wirelength,tm.route_kernel_ms,nr/t_route),flush=True)
        print('   ',[(a,b,c) for a,b,c,_,_ in its][:25],'pops',sum(i[3] for i in its),'visits',sum(i[4] for i in its),flush=True)
        res.success=int(its[-1][1]==0)
        t3=time.time()
        try: print('   check:',check_fast.check_fast(p,res),'%.1fs'%(time.time()-t3),flush=True)
        except Exception as e: print('   CHECK FAILED',e,flush=True)
        R.close()
