/*
 * vpr_adapter.cxx — the reference-side binding: makes the B200 router a drop-in for VPR's
 * try_timing_driven_route (reference vpr/SRC/route/route_timing.c:85, sole caller route_common.c:500).
 *
 * This is the file a maintainer of chinhau5/parallel_eda adds to the VPR build.  It is a C++
 * translation unit because every reference source is compiled as C++ and its headers use C++ types
 * (vpr_types.h:499,533); it includes VPR headers, reads VPR's globals (base/globals.c:48-97), flattens them
 * into the plain-array pf_problem of include/pf_types.h and calls the extern "C" ABI of libpf_router.so.
 * Nothing here routes: the adapter is glue.
 *
 *   route_common.c is compiled with  -Dtry_timing_driven_route=pf_adapter_try_timing_driven_route
 *   so the call at route_common.c:500 lands here; route_timing.c stays in the build unchanged (the placer's
 *   delay lookup and the packer use its helpers, place/timing_place_lookup.c:461, pack/cluster_legality.c:642).
 *
 * Between iterations the reference runs load_timing_graph_net_delays + do_timing_analysis +
 * get_critical_path_delay on the host (route_timing.c:295-309).  Here the timing graph (tnode[], tedge, levels,
 * constraints) is flattened once and the same analysis runs on the device (pf_try_timing_driven_route_sta; the
 * criticalities are bit-identical, tests/test_gpu_sta.py), so nothing crosses PCIe between iterations.  With
 * PF_HOST_STA=1 in the environment the reference's own STA is called back on the host instead.  Clock-to-flipflop override
 * constraints of an SDC file (g_sdc->cf_constraints) are resolved to (source domain, sink tnode) pairs and travel in the
 * timing graph (pf_timing_graph.override_*).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "vpr_types.h"
#include "globals.h"
#include "route_export.h"
#include "route_common.h"
#include "route_tree_timing.h"
#include "route_timing.h"
#include "path_delay.h"
#include "path_delay2.h"
#include "read_sdc.h"
#include "net_delay.h"

#include "pf_router.h"

extern struct s_bb *route_bb;                       /* route/route_common.c:59 */
void timing_driven_check_net_delays(float **net_delay);   /* route/route_timing.c:964 (not in its header) */
extern t_rr_node_route_inf *rr_node_route_inf;      /* route/route_common.c:57 */

namespace {

struct StaCtx {
	float **net_delay;
	t_slack *slacks;
	std::vector<int> net_ptr;
};

/* pf_sta_fn: the host step between iterations (route_timing.c:295-309) */
void sta_callback(void *user, int /*iters_done*/, const float *delay, float *crit, float *cpd) {
	StaCtx *c = (StaCtx *)user;
	for (int i = 0; i < num_nets; i++)
		for (int k = 1; k <= clb_net[i].num_sinks; k++) c->net_delay[i][k] = clb_net[i].is_global ? 0.f : delay[c->net_ptr[i] + k];
	load_timing_graph_net_delays(c->net_delay);
	do_timing_analysis(c->slacks, FALSE, FALSE, FALSE);
	*cpd = get_critical_path_delay();
	vpr_printf(TIO_MESSAGE_INFO, "Critical path: %g ns\n", *cpd);
	for (int i = 0; i < num_nets; i++)
		for (int k = 1; k <= clb_net[i].num_sinks; k++) crit[c->net_ptr[i] + k] = c->slacks->timing_criticality[i][k];
}

/* tnode[] / tnodes_at_level / g_sdc -> pf_timing_graph (include/pf_types.h); storage stays alive in `st` */
struct TimingGraphStore {
	std::vector<int32_t> eptr, eto, cdom, lptr, lnodes, drv;
	std::vector<float> etd, cdel, cons;
	std::vector<uint8_t> ty;
	std::vector<int32_t> ovr_d, ovr_t;
	std::vector<float> ovr_c;
};

void flatten_timing_graph(TimingGraphStore &st, pf_timing_graph &g) {
	memset(&g, 0, sizeof(g));
	st.eptr.assign(num_tnodes + 1, 0); st.cdom.resize(num_tnodes); st.cdel.resize(num_tnodes); st.ty.resize(num_tnodes);
	st.drv.assign(num_nets, -1);
	for (int i = 0; i < num_tnodes; i++) {
		st.eptr[i + 1] = st.eptr[i] + tnode[i].num_edges;
		for (int k = 0; k < tnode[i].num_edges; k++) { st.eto.push_back(tnode[i].out_edges[k].to_node); st.etd.push_back(tnode[i].out_edges[k].Tdel); }
		st.ty[i] = (uint8_t)tnode[i].type; st.cdom[i] = tnode[i].clock_domain; st.cdel[i] = tnode[i].clock_delay;
		if (tnode[i].type == TN_CB_OPIN) {           /* the tnodes that drive inter-block nets (path_delay.c:3342) */
			int iblk, inet;
			get_tnode_block_and_output_net(i, &iblk, &inet);
			if (inet >= 0 && inet < num_nets) st.drv[inet] = i;
		}
	}
	st.lptr.push_back(0);
	for (int lv = 0; lv < num_tnode_levels; lv++) {
		for (int k = 0; k < tnodes_at_level[lv].nelem; k++) st.lnodes.push_back(tnodes_at_level[lv].list[k]);
		st.lptr.push_back((int32_t)st.lnodes.size());
	}
	const int C = g_sdc->num_constrained_clocks;
	for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) st.cons.push_back(g_sdc->domain_constraint[i][j]);
	g.num_tnodes = num_tnodes; g.num_tedges = (int32_t)st.eto.size();
	g.edge_ptr = st.eptr.data(); g.edge_to = st.eto.data(); g.edge_Tdel = st.etd.data(); g.type = st.ty.data();
	g.clock_domain = st.cdom.data(); g.clock_delay = st.cdel.data();
	g.num_levels = num_tnode_levels; g.level_ptr = st.lptr.data(); g.level_nodes = st.lnodes.data();
	g.num_domains = C; g.constraint = st.cons.data();
	/* clock-to-flipflop override constraints (g_sdc->cf_constraints), resolved to (source domain, sink tnode): what
	 * find_cf_constraint(clock name, find_tnode_net_name(inode)) answers at every sink (timing/path_delay.c:2753-2768, :3667-3684,
	 * :3749-3767) — the first matching entry wins, as there */
	if (g_sdc && g_sdc->num_cf_constraints > 0) {
		for (int i = 0; i < num_tnodes; i++) {
			if (tnode[i].num_edges != 0 || (tnode[i].type != TN_FF_SINK && tnode[i].type != TN_OUTPAD_SINK)) continue;
			const char *name = block[tnode[i].block].pb->rr_node_to_pb_mapping[tnode[i].pb_graph_pin->pin_count_in_cluster]->name;
			for (int c = 0; c < g_sdc->num_constrained_clocks; c++) {
				int found = -1;
				for (int icf = 0; icf < g_sdc->num_cf_constraints && found < 0; icf++) {
					bool src = false, snk = false;
					for (int a = 0; a < g_sdc->cf_constraints[icf].num_source; a++) if (strcmp(g_sdc->cf_constraints[icf].source_list[a], g_sdc->constrained_clocks[c].name) == 0) src = true;
					for (int a = 0; src && a < g_sdc->cf_constraints[icf].num_sink; a++) if (strcmp(g_sdc->cf_constraints[icf].sink_list[a], name) == 0) snk = true;
					if (src && snk) found = icf;
				}
				if (found >= 0) { st.ovr_d.push_back(c); st.ovr_t.push_back(i); st.ovr_c.push_back(g_sdc->cf_constraints[found].constraint); }
			}
		}
	}
	g.num_overrides = (int32_t)st.ovr_t.size(); g.override_domain = st.ovr_d.data(); g.override_tnode = st.ovr_t.data(); g.override_constraint = st.ovr_c.data();
	g.num_nets = num_nets; g.net_driver = st.drv.data();
}

}  // namespace

static boolean route_on_b200(struct s_router_opts router_opts, float **net_delay, t_slack *slacks,
		t_ivec **clb_opins_used_locally, boolean timing_analysis_enabled, int router_algorithm);

/* route_common.c:500 (try_route, TIMING_DRIVEN / NO_TIMING) lands here */
boolean pf_adapter_try_timing_driven_route(struct s_router_opts router_opts, float **net_delay, t_slack *slacks,
		t_ivec **clb_opins_used_locally, boolean timing_analysis_enabled) {
	return route_on_b200(router_opts, net_delay, slacks, clb_opins_used_locally, timing_analysis_enabled, 0);
}

/* route_common.c:495 (try_route, BREADTH_FIRST; reference route_breadth_first.c:23) lands here when route_common.c
 * is also compiled with -Dtry_breadth_first_route=pf_adapter_try_breadth_first_route */
boolean pf_adapter_try_breadth_first_route(struct s_router_opts router_opts, t_ivec **clb_opins_used_locally, int /*width_fac*/) {
	return route_on_b200(router_opts, NULL, NULL, clb_opins_used_locally, FALSE, 1);
}

static boolean route_on_b200(struct s_router_opts router_opts, float **net_delay, t_slack *slacks,
		t_ivec **clb_opins_used_locally, boolean timing_analysis_enabled, int router_algorithm) {
	/* ---- flatten the globals (rr_node[], rr_indexed_data[], switch_inf[], clb_net[], net_rr_terminals, route_bb) */
	const int N = num_rr_nodes;
	long E = 0;
	for (int i = 0; i < N; i++) E += rr_node[i].num_edges;
	std::vector<int16_t> xl(N), yl(N), xh(N), yh(N), ptc(N), ci(N), cap(N);
	std::vector<uint8_t> ty(N), dir(N);
	std::vector<float> R(N), C(N);
	std::vector<int32_t> row(N + 1), to(E > 0 ? E : 1);
	std::vector<int16_t> sw(E > 0 ? E : 1);
	long e = 0;
	int nsw = 0;
	for (int i = 0; i < N; i++) {
		const t_rr_node &n = rr_node[i];
		xl[i] = n.xlow; yl[i] = n.ylow; xh[i] = n.xhigh; yh[i] = n.yhigh; ptc[i] = n.ptc_num; ci[i] = n.cost_index;
		cap[i] = n.capacity; ty[i] = (uint8_t)n.type; dir[i] = (uint8_t)n.direction; R[i] = n.R; C[i] = n.C;
		row[i] = (int32_t)e;
		for (int k = 0; k < n.num_edges; k++) { to[e] = n.edges[k]; sw[e] = n.switches[k]; if (sw[e] + 1 > nsw) nsw = sw[e] + 1; e++; }
	}
	row[N] = (int32_t)e;
	std::vector<pf_switch> sws(nsw);
	for (int s = 0; s < nsw; s++) {
		sws[s].buffered = switch_inf[s].buffered; sws[s].R = switch_inf[s].R; sws[s].Cin = switch_inf[s].Cin;
		sws[s].Cout = switch_inf[s].Cout; sws[s].Tdel = switch_inf[s].Tdel;
	}
	std::vector<pf_indexed> idx(num_rr_indexed_data);
	for (int i = 0; i < num_rr_indexed_data; i++) {
		idx[i].base_cost = rr_indexed_data[i].base_cost; idx[i].saved_base_cost = rr_indexed_data[i].saved_base_cost;
		idx[i].ortho_cost_index = rr_indexed_data[i].ortho_cost_index; idx[i].seg_index = rr_indexed_data[i].seg_index;
		idx[i].inv_length = rr_indexed_data[i].inv_length; idx[i].T_linear = rr_indexed_data[i].T_linear;
		idx[i].T_quadratic = rr_indexed_data[i].T_quadratic; idx[i].C_load = rr_indexed_data[i].C_load;
	}
	StaCtx ctx;
	ctx.net_delay = net_delay; ctx.slacks = slacks;
	ctx.net_ptr.assign(num_nets + 1, 0);
	for (int i = 0; i < num_nets; i++) ctx.net_ptr[i + 1] = ctx.net_ptr[i] + clb_net[i].num_sinks + 1;
	const int T = ctx.net_ptr[num_nets];
	std::vector<int32_t> term(T > 0 ? T : 1), bb(4 * (size_t)(num_nets > 0 ? num_nets : 1));
	std::vector<uint8_t> glob(num_nets > 0 ? num_nets : 1);
	for (int i = 0; i < num_nets; i++) {
		glob[i] = clb_net[i].is_global ? 1 : 0;
		for (int k = 0; k <= clb_net[i].num_sinks; k++) term[ctx.net_ptr[i] + k] = net_rr_terminals[i][k];
		bb[4 * i + 0] = route_bb[i].xmin; bb[4 * i + 1] = route_bb[i].xmax; bb[4 * i + 2] = route_bb[i].ymin; bb[4 * i + 3] = route_bb[i].ymax;
	}
	std::vector<int32_t> gsrc, gcnt;
	for (int b = 0; b < num_blocks; b++)
		for (int c = 0; c < block[b].type->num_class; c++)
			if (clb_opins_used_locally[b][c].nelem > 0) { gsrc.push_back(rr_blk_source[b][c]); gcnt.push_back(clb_opins_used_locally[b][c].nelem); }

	pf_problem p;
	memset(&p, 0, sizeof(p));
	p.nx = nx; p.ny = ny; p.num_nodes = N; p.num_edges = (int32_t)E;
	p.xlow = xl.data(); p.ylow = yl.data(); p.xhigh = xh.data(); p.yhigh = yh.data(); p.ptc_num = ptc.data();
	p.cost_index = ci.data(); p.capacity = cap.data(); p.type = ty.data(); p.direction = dir.data(); p.R = R.data(); p.C = C.data();
	p.row_ptr = row.data(); p.edge_to = to.data(); p.edge_sw = sw.data();
	p.num_switches = nsw; p.switches = sws.data(); p.num_indexed = num_rr_indexed_data; p.indexed = idx.data();
	p.num_nets = num_nets; p.num_terminals = T; p.net_ptr = ctx.net_ptr.data(); p.net_terminals = term.data();
	p.net_is_global = glob.data(); p.net_bb = bb.data();
	p.num_opin_groups = (int)gsrc.size(); p.opin_group_source = gsrc.data(); p.opin_group_count = gcnt.data();
	p.opts.first_iter_pres_fac = router_opts.first_iter_pres_fac; p.opts.initial_pres_fac = router_opts.initial_pres_fac;
	p.opts.pres_fac_mult = router_opts.pres_fac_mult; p.opts.acc_fac = router_opts.acc_fac; p.opts.bend_cost = router_opts.bend_cost;
	p.opts.astar_fac = router_opts.astar_fac; p.opts.max_criticality = router_opts.max_criticality;
	p.opts.criticality_exp = router_opts.criticality_exp; p.opts.max_router_iterations = router_opts.max_router_iterations;
	p.opts.timing_analysis_enabled = timing_analysis_enabled ? 1 : 0; p.opts.bb_factor = router_opts.bb_factor;
	p.opts.router_algorithm = router_algorithm;

	/* ---- route on the GPU */
	pf_config cfg;
	pf_config_default(&cfg);
	if (getenv("PF_DEVICE")) cfg.device = atoi(getenv("PF_DEVICE"));
	cfg.verbose = getenv("PF_VERBOSE") ? 1 : 0;
	pf_result res;
	int rc;
	const bool device_sta = timing_analysis_enabled && g_sdc && !getenv("PF_HOST_STA");
	if (device_sta) {
		TimingGraphStore st;
		pf_timing_graph tg;
		flatten_timing_graph(st, tg);
		rc = pf_try_timing_driven_route_sta(&p, &tg, &cfg, &res);
		if (rc == PF_OK)
			for (int i = 0; i < res.num_iter_stats - 1; i++) vpr_printf(TIO_MESSAGE_INFO, "Critical path: %g ns\n", res.iter_stats[i].crit_path_delay);
	} else {
		rc = pf_try_timing_driven_route(&p, &cfg, timing_analysis_enabled ? sta_callback : NULL, &ctx, &res);
	}
	if (rc != PF_OK) {
		/* reference style: message + exit (route_timing.c:482-489 prints "Routing failed" and returns FALSE only
		 * for an unroutable net) */
		vpr_printf(TIO_MESSAGE_ERROR, "pf_router: %s\n", pf_last_error());
		if (rc == PF_EUNROUTABLE) return FALSE;
		exit(1);
	}
	vpr_printf(TIO_MESSAGE_INFO, "%s after %d routing iterations (B200 router).\n",
			res.success ? "Successfully routed" : "Routing failed", res.iterations);

	/* ---- hand the routing back: s_trace lists in update_traceback's segment convention (route_common.c:638),
	 * net delays, occupancy.  The trace elements are plain mallocs we own; free_traceback only threads them
	 * onto the reference's free list, which is harmless. */
	for (int i = 0; i < num_nets; i++) {
		trace_head[i] = trace_tail[i] = NULL;
		for (int k = res.trace_ptr[i]; k < res.trace_ptr[i + 1]; k++) {
			struct s_trace *t = (struct s_trace *)malloc(sizeof(struct s_trace));
			t->index = res.trace_node[k]; t->iswitch = res.trace_switch[k]; t->next = NULL;
			if (trace_tail[i]) trace_tail[i]->next = t; else trace_head[i] = t;
			trace_tail[i] = t;
		}
		if (net_delay) for (int k = 1; k <= clb_net[i].num_sinks; k++) net_delay[i][k] = clb_net[i].is_global ? 0.f : res.net_delay[ctx.net_ptr[i] + k];
	}
	/* occupancy of the nets; the locally used OPINs are then re-reserved by the reference's own routine so
	 * that clb_opins_used_locally carries the node ids check_route expects (check_route.c:599) */
	std::vector<int> occ(N, 0);
	for (int i = 0; i < num_nets; i++) {
		bool after_sink = false;
		for (int k = res.trace_ptr[i]; k < res.trace_ptr[i + 1]; k++) {
			int v = res.trace_node[k];
			if (!(after_sink && k > res.trace_ptr[i])) occ[v]++;      /* the join element is not counted again */
			after_sink = rr_node[v].type == SINK;
		}
	}
	for (int i = 0; i < N; i++) {
		if (occ[i] > 32767) {                                /* rr_node[].occ is a short (vpr_types.h:965) */
			vpr_printf(TIO_MESSAGE_ERROR, "pf_router: occupancy %d of rr node %d does not fit rr_node[].occ\n", occ[i], i);
			exit(1);
		}
		rr_node[i].occ = (short)occ[i]; rr_node_route_inf[i].acc_cost = 1.;
		/* the present cost the reference keeps next to occ (route_common.c:563-568): reserve_locally_used_opins orders the
		 * OPINs of a class by base * acc * pres, so an OPIN a routed net already occupies must look more expensive than a
		 * free one — with a uniform 1 a class with equivalent outputs could re-reserve an occupied pin and overfill it */
		const float pf = router_opts.initial_pres_fac;
		rr_node_route_inf[i].pres_cost = occ[i] < rr_node[i].capacity ? 1.f : 1.f + (occ[i] + 1 - rr_node[i].capacity) * pf;
	}
	reserve_locally_used_opins(router_opts.initial_pres_fac, FALSE, clb_opins_used_locally);
	boolean ok = res.success ? TRUE : FALSE;
#ifdef DEBUG
	if (ok && net_delay) timing_driven_check_net_delays(net_delay);              /* route_timing.c:964: from-scratch Elmore cross-check */
#endif
	pf_result_free(&res);
	return ok;
}
