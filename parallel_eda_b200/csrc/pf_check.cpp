/*
 * pf_check.cpp — host side of check_route on the device (pf_check_route, include/pf_router.h): uploads a finished
 * routing, runs pf_check_net over every net and the occupancy comparison, and reads back the report.
 */
#include "pf_host.h"

extern "C" int pf_check_route(pf_router *r, const pf_result *res, pf_check_report *rep) {
	if (!r || !res || !rep) FAILF(PF_EINVAL, "null argument");
	const pf_problem *p = r->prob;
	if (res->num_nets != r->n || res->num_nodes != r->N || !res->trace_ptr || !res->occ) FAILF(PF_EINVAL, "result does not belong to this problem");
	/* the kernel indexes trace_node / trace_switch through trace_ptr: offsets must be monotone from 0 (any result may be handed in) */
	if (res->trace_ptr[0] != 0) FAILF(PF_EINVAL, "trace_ptr does not start at 0");
	for (int i = 0; i < r->n; i++) if (res->trace_ptr[i + 1] < res->trace_ptr[i]) FAILF(PF_EINVAL, "trace_ptr not monotone at net %d", i);
	if (res->trace_ptr[r->n] > 0 && (!res->trace_node || !res->trace_switch)) FAILF(PF_EINVAL, "result has no trace arrays");
	const size_t total = (size_t)res->trace_ptr[r->n];
	int *d_tp = (int *)pfb_alloc_raw(sizeof(int) * ((size_t)r->n + 1));
	int *d_tn = (int *)pfb_alloc_raw(sizeof(int) * std::max<size_t>(total, 1));
	short *d_ts = (short *)pfb_alloc_raw(sizeof(short) * std::max<size_t>(total, 1));
	int *d_occ = (int *)pfb_alloc_raw(sizeof(int) * (size_t)r->N);
	int *d_occ2 = (int *)pfb_alloc(sizeof(int) * (size_t)r->N);
	unsigned char *d_matched = (unsigned char *)pfb_alloc((size_t)std::max(r->T, 1));
	unsigned char *d_glob = (unsigned char *)pfb_alloc_raw((size_t)std::max(r->n, 1));
	int *d_rep = (int *)pfb_alloc(sizeof(int) * 8);
	unsigned long long *d_wl = (unsigned long long *)pfb_alloc(sizeof(unsigned long long) * 2);
	int h_rep[8] = { 0, 0x7fffffff, 0, 0, 0, 0, 0, 0 };
	unsigned long long h_wl[2] = { 0, 0 };
	int bad = !d_tp || !d_tn || !d_ts || !d_occ || !d_occ2 || !d_matched || !d_glob || !d_rep || !d_wl;
	bad = bad || pfb_h2d(d_tp, res->trace_ptr, sizeof(int) * ((size_t)r->n + 1)) || pfb_h2d(d_tn, res->trace_node, sizeof(int) * total)
		|| pfb_h2d(d_ts, res->trace_switch, sizeof(short) * total) || pfb_h2d(d_occ, res->occ, sizeof(int) * (size_t)r->N)
		|| pfb_h2d(d_glob, p->net_is_global, (size_t)r->n) || pfb_h2d(d_rep, h_rep, sizeof(h_rep))
		|| pfb_launch_check_route(r->nodes, r->edges, r->node_bits, r->N, r->n, r->net_ptr, r->net_term, d_glob, d_tp, d_tn, d_ts, d_matched, d_occ2, d_occ, d_rep, d_wl)
		|| pfb_d2h(h_rep, d_rep, sizeof(h_rep)) || pfb_d2h(h_wl, d_wl, sizeof(h_wl));
	pfb_free(d_tp); pfb_free(d_tn); pfb_free(d_ts); pfb_free(d_occ); pfb_free(d_occ2); pfb_free(d_matched); pfb_free(d_glob); pfb_free(d_rep); pfb_free(d_wl);
	if (bad) CUDA_FAIL();
	long long reserved = 0;
	for (int g = 0; g < p->num_opin_groups; g++) reserved += p->opin_group_count[g];
	memset(rep, 0, sizeof(*rep));
	rep->bad_nets = h_rep[0]; rep->first_bad_net = h_rep[0] ? h_rep[1] : -1; rep->first_bad_code = h_rep[0] ? h_rep[2] : 0;
	rep->occupancy_mismatch = h_rep[3] + ((long long)h_wl[1] != reserved ? 1 : 0);
	rep->overused_nodes = h_rep[4];
	rep->wirelength = (int64_t)h_wl[0]; rep->reserved_opins = (int64_t)h_wl[1];
	rep->ok = rep->bad_nets == 0 && rep->occupancy_mismatch == 0;
	return PF_OK;
}
