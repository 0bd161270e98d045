/*
 * pf_sta.cpp — host side of the device static timing analysis behind pf_sta_* (include/pf_router.h): the timing
 * graph is renumbered in level order, uploaded once, and every analysis is a fixed sequence of launches.  The
 * arithmetic is in pf_sta_device.cuh; pf_try_timing_driven_route_sta (pf_router.cpp) puts it into the router loop.
 */
#include "pf_host.h"

struct pf_sta {
	PfStaDev d;
	int num_domains, num_tedges;
	std::vector<float> constraint;
	std::vector<int> seg_begin, seg_end, seg_spread;     /* sweep plan over levels, ascending */
	float *stat;                                         /* [num_domains^2][4] on the device */
	float *scratch_delay, *scratch_crit;                 /* device staging for the host-buffer entry point */
	float *scratch_slack;                                /* pf_sta_analyze_final, allocated on first use */
	void *owned[32]; int num_owned;
};

extern "C" void pf_sta_destroy(pf_sta *s) {
	if (!s) return;
	for (int i = 0; i < s->num_owned; i++) pfb_free(s->owned[i]);
	delete s;
}

template <class T> static T *sta_upload(pf_sta *s, const std::vector<T> &v) {
	T *d = (T *)pfb_alloc(sizeof(T) * std::max<size_t>(v.size(), 1));
	if (!d) return NULL;
	s->owned[s->num_owned++] = d;
	if (!v.empty() && pfb_h2d(d, v.data(), sizeof(T) * v.size()) != 0) return NULL;
	return d;
}

extern "C" int pf_sta_create(const pf_timing_graph *g, const pf_problem *p, const pf_config *cfg, pf_sta **out) {
	char msg[256];
	if (!g || !p || !cfg || !out) FAILF(PF_EINVAL, "null argument");
	*out = NULL;
	if (g->num_nets != p->num_nets) FAILF(PF_EINVAL, "timing graph has %d nets, the problem %d", g->num_nets, p->num_nets);
	if (pf_timing_graph_check(g, p->net_ptr, msg, sizeof(msg)) != PF_OK) FAILF(PF_EINVAL, "invalid timing graph: %s", msg);
	if (pfb_init(cfg->device) != 0) CUDA_FAIL();
	const int N = g->num_tnodes, E = g->num_tedges, T = p->num_terminals;
	/* Renumber the tnodes in level order (new id = position in the level lists): a level is then a contiguous
	 * range, its structural reads are coalesced, and the kernels need no level list.  Out-edges keep the
	 * reference's order within a tnode, so "pin k of a net = out-edge k-1 of its driver" still holds. */
	std::vector<int> pos((size_t)N), r_eptr((size_t)N + 1, 0), r_eto((size_t)std::max(E, 1)), r_dom((size_t)N), new_edge((size_t)std::max(E, 1));
	std::vector<float> r_tdel((size_t)std::max(E, 1)), r_cdel((size_t)N);
	std::vector<unsigned char> r_type((size_t)N);
	for (int k = 0; k < N; k++) pos[(size_t)g->level_nodes[k]] = k;
	for (int k = 0; k < N; k++) {
		const int o = g->level_nodes[k];
		r_eptr[(size_t)k + 1] = r_eptr[(size_t)k] + (g->edge_ptr[o + 1] - g->edge_ptr[o]);
		r_dom[(size_t)k] = g->clock_domain[o]; r_cdel[(size_t)k] = g->clock_delay[o]; r_type[(size_t)k] = g->type[o];
		for (int e = g->edge_ptr[o], q = r_eptr[(size_t)k]; e < g->edge_ptr[o + 1]; e++, q++) {
			r_eto[(size_t)q] = pos[(size_t)g->edge_to[e]]; r_tdel[(size_t)q] = g->edge_Tdel[e]; new_edge[(size_t)e] = q;
		}
	}
	/* in-edges as (source, edge) pairs */
	std::vector<int> in_ptr((size_t)N + 1, 0), in_rec(2 * (size_t)std::max(E, 1));
	for (int e = 0; e < E; e++) in_ptr[(size_t)r_eto[(size_t)e] + 1]++;
	for (int n = 0; n < N; n++) in_ptr[(size_t)n + 1] += in_ptr[(size_t)n];
	{
		std::vector<int> fill(in_ptr.begin(), in_ptr.end() - 1);
		for (int n = 0; n < N; n++)
			for (int e = r_eptr[(size_t)n]; e < r_eptr[(size_t)n + 1]; e++) { int k = fill[(size_t)r_eto[(size_t)e]]++; in_rec[2 * (size_t)k] = n; in_rec[2 * (size_t)k + 1] = e; }
	}
	/* net pin -> (driver tnode, out-edge): pin k of net i is out-edge k-1 of its driver (path_delay.c:479-500) */
	std::vector<int> term_edge((size_t)std::max(T, 1), -1), term_driver((size_t)std::max(T, 1), -1);
	for (int i = 0; i < p->num_nets; i++) {
		if (g->net_driver[i] < 0) continue;
		const int d = pos[(size_t)g->net_driver[i]];
		for (int k = 1; k < p->net_ptr[i + 1] - p->net_ptr[i]; k++) { term_edge[(size_t)p->net_ptr[i] + k] = r_eptr[(size_t)d] + k - 1; term_driver[(size_t)p->net_ptr[i] + k] = d; }
	}
	(void)new_edge;
	pf_sta *s = new pf_sta();
	memset(&s->d, 0, sizeof(s->d));
	s->num_owned = 0; s->num_domains = g->num_domains; s->num_tedges = E; s->scratch_slack = NULL;
	s->constraint.assign(g->constraint, g->constraint + (size_t)g->num_domains * g->num_domains);
	/* sweep plan: a level wider than 4 K tnodes gets the whole GPU, runs of narrower ones share one CTA */
	for (int lv = 0; lv < g->num_levels;) {
		const int width = g->level_ptr[lv + 1] - g->level_ptr[lv];
		if (width > 4096) { s->seg_begin.push_back(lv); s->seg_end.push_back(lv + 1); s->seg_spread.push_back(width); lv++; continue; }
		int e = lv;
		while (e < g->num_levels && g->level_ptr[e + 1] - g->level_ptr[e] <= 4096) e++;
		s->seg_begin.push_back(lv); s->seg_end.push_back(e); s->seg_spread.push_back(0);
		lv = e;
	}
	PfStaDev &d = s->d;
	d.num_tnodes = N; d.num_terminals = T; d.num_levels = g->num_levels;
	bool ok = true;
	std::vector<int> ident((size_t)N), lptr(g->level_ptr, g->level_ptr + g->num_levels + 1);
	for (int k = 0; k < N; k++) ident[(size_t)k] = k;
	d.edge_ptr = sta_upload(s, r_eptr); d.edge_to = sta_upload(s, r_eto);
	d.in_ptr = sta_upload(s, in_ptr); d.in_rec = sta_upload(s, in_rec);
	d.clock_domain = sta_upload(s, r_dom);
	d.level_ptr = sta_upload(s, lptr); d.level_nodes = sta_upload(s, ident);
	d.term_edge = sta_upload(s, term_edge); d.term_driver = sta_upload(s, term_driver);
	d.Tdel = sta_upload(s, r_tdel); d.clock_delay = sta_upload(s, r_cdel); d.type = sta_upload(s, r_type);
	{ std::vector<float> v((size_t)N, 0.f); d.T_arr = sta_upload(s, v); d.T_req = sta_upload(s, v); }
	if (g->num_overrides > 0) {
		/* override constraints follow the renumbering; (tnode, domain) order is kept for the binary search at the sinks */
		std::vector<int> idx((size_t)g->num_overrides), ot((size_t)g->num_overrides), od((size_t)g->num_overrides);
		std::vector<float> oc((size_t)g->num_overrides);
		for (int k = 0; k < g->num_overrides; k++) idx[(size_t)k] = k;
		std::sort(idx.begin(), idx.end(), [&](int a, int b) {
			const int ta = pos[(size_t)g->override_tnode[a]], tb = pos[(size_t)g->override_tnode[b]];
			return ta != tb ? ta < tb : g->override_domain[a] < g->override_domain[b];
		});
		for (int k = 0; k < g->num_overrides; k++) {
			ot[(size_t)k] = pos[(size_t)g->override_tnode[idx[(size_t)k]]]; od[(size_t)k] = g->override_domain[idx[(size_t)k]]; oc[(size_t)k] = g->override_constraint[idx[(size_t)k]];
		}
		d.num_overrides = g->num_overrides;
		d.ovr_tnode = sta_upload(s, ot); d.ovr_domain = sta_upload(s, od); d.ovr_constraint = sta_upload(s, oc);
		ok = ok && d.ovr_tnode && d.ovr_domain && d.ovr_constraint;
	}
	{ std::vector<float> v((size_t)std::max(g->num_domains * g->num_domains, 1) * 4, 0.f); s->stat = sta_upload(s, v); }
	{ std::vector<float> v((size_t)std::max(T, 1), 0.f); s->scratch_delay = sta_upload(s, v); s->scratch_crit = sta_upload(s, v); }
	ok = ok && d.edge_ptr && d.edge_to && d.clock_domain && d.level_ptr && d.level_nodes && d.in_ptr && d.in_rec && d.term_edge && d.term_driver && d.Tdel && d.clock_delay && d.type && d.T_arr && d.T_req
		&& s->stat && s->scratch_delay && s->scratch_crit;
	if (!ok || pfb_sync() != 0) { pf_sta_destroy(s); CUDA_FAIL(); }
	*out = s;
	return PF_OK;
}

extern "C" int pf_sta_analyze_device(pf_sta *s, const void *dev_net_delay, void *dev_crit, float *cpd_ns) {
	if (!s || !dev_net_delay || !dev_crit) FAILF(PF_EINVAL, "null argument");
	const int C = s->num_domains, nseg = (int)s->seg_begin.size();
	CKB(pfb_sta_load(&s->d, (const float *)dev_net_delay));
	CKB(pfb_zero(dev_crit, sizeof(float) * (size_t)std::max(s->d.num_terminals, 1)));     /* path_delay.c:2403-2410 */
	for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) {
		const float constraint = s->constraint[(size_t)i * C + j];
		float *stat = s->stat + 4 * ((size_t)i * C + j);
		if (!(constraint > -1.e-15)) continue;                                                /* DO_NOT_ANALYSE */
		s->d.src_domain = i;
		CKB(pfb_sta_begin_pair(&s->d, stat));
		for (int k = 0; k < nseg; k++) CKB(pfb_sta_sweep(&s->d, 1, s->seg_begin[k], s->seg_end[k], s->seg_spread[k], i, constraint, stat));
		for (int k = nseg - 1; k >= 0; k--) CKB(pfb_sta_sweep(&s->d, 0, s->seg_begin[k], s->seg_end[k], s->seg_spread[k], j, constraint, stat));
		CKB(pfb_sta_update(&s->d, constraint, stat, (float *)dev_crit));
	}
	if (cpd_ns) return pf_sta_read_cpd(s, cpd_ns);
	return PF_OK;
}

extern "C" int pf_sta_read_cpd(pf_sta *s, float *cpd_ns) {
	if (!s || !cpd_ns) FAILF(PF_EINVAL, "null argument");
	const int C = s->num_domains;
	{
		/* get_critical_path_delay, path_delay.c:3791-3810: the cpd of the pair with the least slack */
		std::vector<float> h((size_t)std::max(C * C, 1) * 4, 0.f);
		CKB(pfb_d2h(h.data(), s->stat, sizeof(float) * h.size()));
		float least = (float)1.e30, cpd = -1.f;
		for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) {
			if (!(s->constraint[(size_t)i * C + j] > -1.e-15)) continue;
			const float *st = &h[4 * ((size_t)i * C + j)];
			if (least > st[2]) { least = st[2]; cpd = st[1]; }
		}
		*cpd_ns = (float)(cpd * 1e9);
	}
	return PF_OK;
}

extern "C" int pf_sta_analyze(pf_sta *s, const float *net_delay, float *crit, float *cpd_ns) {
	if (!s || !net_delay || !crit) FAILF(PF_EINVAL, "null argument");
	const size_t bytes = sizeof(float) * (size_t)s->d.num_terminals;
	CKB(pfb_h2d(s->scratch_delay, net_delay, bytes));
	int rc = pf_sta_analyze_device(s, s->scratch_delay, s->scratch_crit, cpd_ns);
	if (rc != PF_OK) return rc;
	CKB(pfb_d2h(crit, s->scratch_crit, bytes));
	return PF_OK;
}

/* The analysis of the finished routing: routing_stats (base/stats.c:155-178) runs do_timing_analysis(slacks, FALSE, FALSE,
 * TRUE) once the router is done — the same sweeps with the REAL required times at the sinks (path_delay.c:2786-2790; slacks
 * may be negative) and the least slack of every net pin kept over all analysed constraints (update_slacks :3117-3125; pins no
 * traversal reaches keep HUGE_POSITIVE_FLOAT, :2393).  crit may be NULL. */
extern "C" int pf_sta_analyze_final(pf_sta *s, const float *net_delay, float *slack, float *crit, float *cpd_ns) {
	if (!s || !net_delay || !slack) FAILF(PF_EINVAL, "null argument");
	const size_t T = (size_t)std::max(s->d.num_terminals, 1), bytes = sizeof(float) * (size_t)s->d.num_terminals;
	if (!s->scratch_slack) {
		if (s->num_owned >= (int)(sizeof(s->owned) / sizeof(s->owned[0]))) FAILF(PF_EINVAL, "pf_sta: allocation table full");
		s->scratch_slack = (float *)pfb_alloc(sizeof(float) * T);
		if (!s->scratch_slack) CUDA_FAIL();
		s->owned[s->num_owned++] = s->scratch_slack;
	}
	std::vector<float> init(T, (float)1.e30);                                                 /* HUGE_POSITIVE_FLOAT */
	CKB(pfb_h2d(s->scratch_slack, init.data(), sizeof(float) * T));
	CKB(pfb_h2d(s->scratch_delay, net_delay, bytes));
	s->d.final_analysis = 1; s->d.slack = s->scratch_slack;
	int rc = pf_sta_analyze_device(s, s->scratch_delay, s->scratch_crit, cpd_ns);
	s->d.final_analysis = 0; s->d.slack = NULL;
	if (rc != PF_OK) return rc;
	CKB(pfb_d2h(slack, s->scratch_slack, bytes));
	if (crit) CKB(pfb_d2h(crit, s->scratch_crit, bytes));
	return PF_OK;
}
