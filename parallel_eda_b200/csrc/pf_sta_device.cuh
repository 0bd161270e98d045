/*
 * pf_sta_device.cuh — per-element bodies of the device static timing analysis.
 *
 * What the reference runs on the host between router iterations (route_timing.c:295-309):
 *   load_timing_graph_net_delays        timing/path_delay.c:479-500
 *   do_timing_analysis                  timing/path_delay.c:2258-2522 (SLACK_DEFINITION 'R', no PATH_COUNTING)
 *   do_timing_analysis_for_constraint   :2571-2912
 *   update_slacks                       :3075-3160
 * restated level-synchronously: the reference pushes arrival times along out-edges in level order; here every
 * tnode of a level PULLS over its in-edges (max is order independent, so the floats are bit-identical) and the
 * levels are separated by barriers.  Shared by pf_kernels.cu and the test emulator backend.
 */
#ifndef PF_STA_DEVICE_CUH
#define PF_STA_DEVICE_CUH

#include "pf_device.cuh"

#define PF_STA_HUGE_POS 1.e30     /* HUGE_POSITIVE_FLOAT, vpr_types.h:84 (double constants, as in the reference) */
#define PF_STA_HUGE_NEG -1.e30
#define PF_STA_NEG_EPS -1.e-15    /* NEGATIVE_EPSILON */
#define PF_STA_TN_INPAD_SOURCE 0  /* e_tnode_type, vpr_types.h:305-323 */
#define PF_STA_TN_FF_SOURCE 12
#define PF_STA_TN_FF_CLOCK 13


#ifndef PF_EMU
/* float atomics by the sign-aware integer trick (IEEE order == integer order for >= 0, reversed for < 0) */
PF_DEV void pf_atomic_max_f(float *p, float v) {
	if (v >= 0.f) atomicMax((int *)p, __float_as_int(v)); else atomicMin((unsigned *)p, __float_as_uint(v));
}
PF_DEV void pf_atomic_min_f(float *p, float v) {
	if (v >= 0.f) atomicMin((int *)p, __float_as_int(v)); else atomicMax((unsigned *)p, __float_as_uint(v));
}
#endif

/* load_timing_graph_net_delays: the delay of net pin k is the delay of out-edge k-1 of the net's driver tnode */
PF_DEV void pf_sta_load_delay(const PfStaDev &S, int t, const float *net_delay) {
	const int e = S.term_edge[t];
	if (e >= 0) S.Tdel[e] = net_delay[t];
}

PF_DEV void pf_sta_reset_node(const PfStaDev &S, int n) {
	S.T_arr[n] = (float)PF_STA_HUGE_NEG;
	S.T_req[n] = (float)PF_STA_HUGE_POS;
}

/* forward sweep, one tnode of level lv (path_delay.c:2606-2690) */
PF_DEV void pf_sta_forward_node(const PfStaDev &S, int n, int lv, int src_domain, float *stat) {
	if (lv == 0) {
		if (S.clock_domain[n] == src_domain) {
			if (S.type[n] == PF_STA_TN_FF_SOURCE) S.T_arr[n] = S.clock_delay[n];
			else if (S.type[n] == PF_STA_TN_INPAD_SOURCE) S.T_arr[n] = 0.f;
		}
		return;
	}
	float best = (float)PF_STA_HUGE_NEG;
	int any = 0;
	for (int k = S.in_ptr[n]; k < S.in_ptr[n + 1]; k++) {
		const float ta = S.T_arr[S.in_rec[2 * k]];
		if (ta < PF_STA_NEG_EPS) continue;               /* the predecessor is not in this traversal */
		const float cand = ta + S.Tdel[S.in_rec[2 * k + 1]];
		if (cand > best) best = cand;                    /* set_and_balance_arrival_time, :3449 */
		any = 1;
	}
	if (any) {
		S.T_arr[n] = best;
		pf_atomic_max_f(&stat[0], best);                 /* max_Tarr */
	}
}

/* backward sweep, one tnode (path_delay.c:2693-2895) */
PF_DEV void pf_sta_backward_node(const PfStaDev &S, int n, int sink_domain, float constraint, float *stat) {
	const int e0 = S.edge_ptr[n], e1 = S.edge_ptr[n + 1];
	const float ta = S.T_arr[n];
	if (e0 == e1) {                                      /* sink */
		if (S.type[n] == PF_STA_TN_FF_CLOCK || ta < PF_STA_HUGE_NEG + 1) return;
		if (S.clock_domain[n] != sink_domain) return;
		if (S.num_overrides > 0) {                       /* find_cf_constraint, :2753-2768: an override of this sink for the source domain */
			int lo = 0, hi = S.num_overrides;
			while (lo < hi) {
				const int mid = (lo + hi) >> 1;
				if (S.ovr_tnode[mid] < n || (S.ovr_tnode[mid] == n && S.ovr_domain[mid] < S.src_domain)) lo = mid + 1; else hi = mid;
			}
			if (lo < S.num_overrides && S.ovr_tnode[lo] == n && S.ovr_domain[lo] == S.src_domain) {
				constraint = S.ovr_constraint[lo];
				if (constraint < PF_STA_NEG_EPS) return;     /* DO_NOT_ANALYSE for this particular sink */
			}
		}
		const float real = constraint + S.clock_delay[n], max_Tarr = stat[0];
		S.T_req[n] = (S.final_analysis || real > max_Tarr) ? real : max_Tarr;  /* T_req-relaxed slack except in the final analysis, :2786-2790 */
		pf_atomic_max_f(&stat[1], ta - S.clock_delay[n]);/* critical path delay of this constraint */
		return;
	}
	if (ta < PF_STA_HUGE_NEG + 1) return;
	int found = 0;
	for (int e = e0; e < e1 && !found; e++) if (S.T_req[S.edge_to[e]] < PF_STA_HUGE_POS) found = 1;
	if (!found) return;
	float tr_min = S.T_req[n];
	for (int e = e0; e < e1; e++) {
		const int to = S.edge_to[e];
		const float tr = S.T_req[to], cand = tr - S.Tdel[e];
		if (cand < tr_min) tr_min = cand;
		if (S.edge_ptr[to + 1] == S.edge_ptr[to] && S.clock_domain[to] == sink_domain) pf_atomic_min_f(&stat[2], tr - S.Tdel[e] - ta);
	}
	S.T_req[n] = tr_min;
}

/* update_slacks for one net pin (path_delay.c:3075-3160): criticality = 1 - slack / max(max_Tarr, constraint) */
PF_DEV void pf_sta_update_terminal(const PfStaDev &S, int t, float constraint, const float *stat, float *crit) {
	const int d = S.term_driver[t], e = S.term_edge[t];
	if (d < 0 || e < 0) return;
	if (!(S.T_arr[d] > PF_STA_HUGE_NEG + 1 && S.T_req[d] < PF_STA_HUGE_POS - 1)) return;
	const int to = S.edge_to[e];
	if (!(S.T_arr[to] > PF_STA_HUGE_NEG + 1 && S.T_req[to] < PF_STA_HUGE_POS - 1)) return;
	const float max_Tarr = stat[0];
	const float denom = max_Tarr > constraint ? max_Tarr : constraint;
	if (S.slack) {                                       /* update_slack == TRUE, :3117-3125 */
		const float slk = S.T_req[to] - S.T_arr[d] - S.Tdel[e];
		if (slk < S.slack[t]) S.slack[t] = slk;
	}
	const float tc = 1 - (S.T_req[to] - S.T_arr[d] - S.Tdel[e]) / denom;
	if (tc > crit[t]) crit[t] = tc;
}

#endif /* PF_STA_DEVICE_CUH */
