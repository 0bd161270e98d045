/* TEST / MEASUREMENT INFRASTRUCTURE — a multi-threaded CPU router for the "parallel CPU router on the box's host cores" line
 * of the report (CPU side (2)).  The reference's own parallel routers (TBB / MPI / Boost, SURVEY.md §2 rows
 * 17-19) cannot be built here, so this restates their common scheme on top of the bit-exact serial restatement
 * (pf_oracle.c, included below): every thread owns its search state (labels, heap, route tree), all threads share the
 * occupancy and cost arrays (occ updated with atomics, as the MPI router's sync_recalc_occ does between ranks,
 * parallel_route/spatial.cxx:3371-3383), nets are handed out in decreasing-fanout order from one atomic counter, and the
 * per-iteration steps (reserve_locally_used_opins, feasibility, pathfinder_update_cost) run on one thread behind a join.
 * Timing analysis off (criticality 0), like BASELINE configs[4].  Not bit-exact with anything: the order in which nets see
 * each other's occupancy depends on the schedule.  Never linked into the product. */
#define PF_ORACLE_PARALLEL 1
#include "pf_oracle.c"

#include <pthread.h>
#include <time.h>

typedef struct {
	oracle o;
	const int *net_index;
	int n;
	int *next;                /* shared atomic cursor into net_index */
	float pres_fac;
	const float *crit;
	float *net_delay;
	int rc;
} par_worker;

static int par_alloc(oracle *o, const pf_problem *p, const oracle *shared, int max_pins) {
	int N = p->num_nodes, i;
	memset(o, 0, sizeof(*o));
	o->p = p;
	o->occ = shared->occ; o->pres_cost = shared->pres_cost; o->acc_cost = shared->acc_cost;
	o->tr_node = shared->tr_node; o->tr_sw = shared->tr_sw; o->tr_n = shared->tr_n; o->tr_cap = shared->tr_cap;   /* a net's slot belongs to whoever routes it */
	o->path_cost = (float *)malloc(sizeof(float) * N);
	o->backward_path_cost = (float *)calloc(N, sizeof(float));
	o->prev_node = (int *)malloc(sizeof(int) * N);
	o->prev_edge = (short *)malloc(sizeof(short) * N);
	o->target_flag = (short *)calloc(N, sizeof(short));
	o->mod_list = (int *)malloc(sizeof(int) * N);
	o->rr_to_rt = (int *)malloc(sizeof(int) * N);
	o->base_cost = (float *)malloc(sizeof(float) * p->num_indexed);
	o->heap_size = p->nx * p->ny > 16 ? p->nx * p->ny : 16;
	o->heap = (heap_item *)malloc(sizeof(heap_item) * ((size_t)o->heap_size + 1));
	o->heap_tail = 1;
	o->tree_cap = 1024;
	o->tree = (rt_node *)malloc(sizeof(rt_node) * (size_t)o->tree_cap);
	o->pin_criticality = (float *)calloc((size_t)max_pins + 2, sizeof(float));
	o->sink_order = (int *)calloc((size_t)max_pins + 2, sizeof(int));
	o->rt_of_sink = (int *)calloc((size_t)max_pins + 2, sizeof(int));
	if (!o->path_cost || !o->backward_path_cost || !o->prev_node || !o->prev_edge || !o->target_flag || !o->mod_list || !o->rr_to_rt
			|| !o->base_cost || !o->heap || !o->tree || !o->pin_criticality || !o->sink_order || !o->rt_of_sink) return PF_ENOMEM;
	for (i = 0; i < N; i++) { o->prev_node[i] = PF_NO_PREVIOUS; o->prev_edge[i] = PF_NO_PREVIOUS; o->path_cost[i] = PF_HUGE_POSITIVE_FLOAT; o->rr_to_rt[i] = -1; }
	for (i = 0; i < p->num_indexed; i++) o->base_cost[i] = p->indexed[i].base_cost;
	return PF_OK;
}

static void par_free(oracle *o) {
	free(o->path_cost); free(o->backward_path_cost); free(o->prev_node); free(o->prev_edge); free(o->target_flag);
	free(o->mod_list); free(o->rr_to_rt); free(o->base_cost); free(o->heap); free(o->tree);
	free(o->pin_criticality); free(o->sink_order); free(o->rt_of_sink);
}

static void *par_route(void *arg) {
	par_worker *w = (par_worker *)arg;
	const pf_problem *p = w->o.p;
	for (;;) {
		int k = __atomic_fetch_add(w->next, 1, __ATOMIC_RELAXED), inet, rc;
		if (k >= w->n) return NULL;
		inet = w->net_index[k];
		if (p->net_is_global[inet]) continue;
		rc = timing_driven_route_net(&w->o, inet, w->pres_fac, w->crit + p->net_ptr[inet], w->net_delay + p->net_ptr[inet]);
		if (rc != PF_OK) { w->rc = rc; return NULL; }
		w->o.cur.nets_routed++;
	}
}

/* congested_only = 0: every net is re-routed in every iteration (the serial reference's policy, route_timing.c:152-187);
 * congested_only = 1: from the second iteration on only nets whose current route touches an overused rr node (the device
 * router's policy; reference precedent partitioning_multi_sink_delta_stepping_route.cxx:6241-6269) */
int pf_oracle_route_parallel(const pf_problem *p, int nthreads, int congested_only, int max_iters_override, pf_result *out, double *iter_seconds /* [max_iters] or NULL */) {
	oracle S;                                            /* shared state + the single-threaded steps */
	par_worker *W;
	pthread_t *th;
	int N = p->num_nodes, n = p->num_nets, T = p->num_terminals;
	int i, inet, itry, t, rc = PF_OK, max_pins = 0, success = 0, max_iters, next;
	int *net_index, *work, nwork;
	float *sinks, *crit, *net_delay, pres_fac;
	pf_iter_stats *stats;
	memset(out, 0, sizeof(*out));
	if (nthreads < 1) nthreads = 1;
	if (p->opts.router_algorithm != 0) return PF_EINVAL;
	memset(&S, 0, sizeof(S));
	S.p = p;
	S.occ = (int *)calloc(N, sizeof(int));
	S.pres_cost = (float *)malloc(sizeof(float) * N);
	S.acc_cost = (float *)malloc(sizeof(float) * N);
	for (i = 0; i < N; i++) { S.pres_cost[i] = 1.; S.acc_cost[i] = 1.; }
	S.tr_node = (int **)calloc(n, sizeof(int *));
	S.tr_sw = (short **)calloc(n, sizeof(short *));
	S.tr_n = (int *)calloc(n, sizeof(int));
	S.tr_cap = (int *)calloc(n, sizeof(int));
	for (inet = 0; inet < n; inet++)
		if (!p->net_is_global[inet] && p->net_ptr[inet + 1] - p->net_ptr[inet] > max_pins) max_pins = p->net_ptr[inet + 1] - p->net_ptr[inet];
	W = (par_worker *)calloc((size_t)nthreads, sizeof(par_worker));
	th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	for (t = 0; t < nthreads; t++) if ((rc = par_alloc(&W[t].o, p, &S, max_pins)) != PF_OK) return rc;
	/* worker 0's instance also runs reserve_locally_used_opins (it needs a heap and the per-group OPIN lists) */
	W[0].o.opin_list = (int **)calloc(p->num_opin_groups ? p->num_opin_groups : 1, sizeof(int *));
	for (i = 0; i < p->num_opin_groups; i++) W[0].o.opin_list[i] = (int *)calloc(p->opin_group_count[i] ? p->opin_group_count[i] : 1, sizeof(int));

	max_iters = max_iters_override > 0 ? max_iters_override : p->opts.max_router_iterations;
	stats = (pf_iter_stats *)calloc((size_t)max_iters + 1, sizeof(pf_iter_stats));
	sinks = (float *)malloc(sizeof(float) * (size_t)(n + 1));
	net_index = (int *)malloc(sizeof(int) * (size_t)(n + 1));
	for (i = 0; i < n; i++) { sinks[i] = p->net_ptr[i + 1] - p->net_ptr[i] - 1; net_index[i] = i; }
	pf_oracle_heapsort(net_index, sinks, n, 1);
	work = (int *)malloc(sizeof(int) * (size_t)(n + 1));
	crit = (float *)calloc((size_t)T + 1, sizeof(float));            /* timing analysis off: criticality 0 throughout */
	net_delay = (float *)calloc((size_t)T + 1, sizeof(float));
	pres_fac = p->opts.first_iter_pres_fac;

	for (itry = 1; itry <= max_iters; itry++) {
		struct timespec t0, t1;
		pf_iter_stats cur;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		memset(&cur, 0, sizeof(cur));
		cur.pres_fac = pres_fac;
		next = 0;
		/* this iteration's nets, in decreasing-fanout order */
		nwork = 0;
		for (i = 0; i < n; i++) {
			int k, hit = !(congested_only && itry > 1);
			inet = net_index[i];
			if (p->net_is_global[inet]) continue;
			for (k = 0; !hit && k < S.tr_n[inet]; k++) { int v = S.tr_node[inet][k]; hit = S.occ[v] > p->capacity[v]; }
			if (hit) work[nwork++] = inet;
		}
		for (t = 0; t < nthreads; t++) {
			memset(&W[t].o.cur, 0, sizeof(W[t].o.cur));
			W[t].net_index = work; W[t].n = nwork; W[t].next = &next; W[t].pres_fac = pres_fac; W[t].crit = crit; W[t].net_delay = net_delay; W[t].rc = PF_OK;
		}
		for (t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, par_route, &W[t]);
		par_route(&W[0]);
		for (t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
		for (t = 0; t < nthreads; t++) {
			if (W[t].rc != PF_OK) rc = W[t].rc;
			cur.nets_routed += W[t].o.cur.nets_routed; cur.heap_pushes += W[t].o.cur.heap_pushes;
			cur.heap_pops += W[t].o.cur.heap_pops; cur.edge_visits += W[t].o.cur.edge_visits;
		}
		if (rc != PF_OK) break;
		if (itry == 1) {                                        /* route_timing.c:189-225 wirelength sanity abort */
			long total = 0, avail = 0;
			for (i = 0; i < N; i++)
				if (p->type[i] == PF_CHANX || p->type[i] == PF_CHANY) avail += 1 + p->xhigh[i] - p->xlow[i] + p->yhigh[i] - p->ylow[i];
			for (inet = 0; inet < n; inet++)
				if (!p->net_is_global[inet] && p->net_ptr[inet + 1] - p->net_ptr[inet] - 1 != 0) total += trace_wirelength(p, S.tr_node[inet], S.tr_n[inet]);
			if ((float)total / (float)avail > PF_FIRST_ITER_WIRELENGTH_LIMIT) { cur.overused_nodes = count_overused(&W[0].o); stats[itry - 1] = cur; itry++; break; }
		}
		reserve_locally_used_opins(&W[0].o, pres_fac, itry != 1);
		cur.overused_nodes = count_overused(&W[0].o);
		stats[itry - 1] = cur;
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (iter_seconds) iter_seconds[itry - 1] = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
		if (cur.overused_nodes == 0) { success = 1; itry++; break; }
		if (itry == 1) { pres_fac = p->opts.initial_pres_fac; pathfinder_update_cost(&W[0].o, pres_fac, 0.); }
		else {
			pres_fac *= p->opts.pres_fac_mult;
			pres_fac = fminf(pres_fac, (float)(PF_HUGE_POSITIVE_FLOAT / 1e5));
			pathfinder_update_cost(&W[0].o, pres_fac, p->opts.acc_fac);
		}
	}
	itry--;
	if (rc == PF_OK) {
		int total = 0, wl = 0;
		out->success = success; out->iterations = itry; out->num_nets = n;
		out->trace_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
		out->trace_ptr[0] = 0;
		for (inet = 0; inet < n; inet++) { total += S.tr_n[inet]; out->trace_ptr[inet + 1] = total; }
		out->trace_node = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total ? total : 1));
		out->trace_switch = (int16_t *)malloc(sizeof(int16_t) * (size_t)(total ? total : 1));
		for (inet = 0; inet < n; inet++) {
			memcpy(out->trace_node + out->trace_ptr[inet], S.tr_node[inet], sizeof(int) * (size_t)S.tr_n[inet]);
			memcpy(out->trace_switch + out->trace_ptr[inet], S.tr_sw[inet], sizeof(short) * (size_t)S.tr_n[inet]);
			if (!p->net_is_global[inet] && p->net_ptr[inet + 1] - p->net_ptr[inet] - 1 != 0) wl += trace_wirelength(p, S.tr_node[inet], S.tr_n[inet]);
		}
		out->total_wirelength = wl;
		out->serial_num = pf_serial_num(p, out->trace_ptr, out->trace_node);
		out->num_terminals = T; out->net_delay = net_delay; net_delay = NULL;
		out->num_nodes = N;
		out->occ = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
		for (i = 0; i < N; i++) out->occ[i] = S.occ[i];
		out->num_iter_stats = itry; out->iter_stats = stats; stats = NULL;
	}
	for (t = 0; t < nthreads; t++) par_free(&W[t].o);
	for (i = 0; i < p->num_opin_groups; i++) free(W[0].o.opin_list[i]);
	free(W[0].o.opin_list);
	free(S.occ); free(S.pres_cost); free(S.acc_cost);
	for (inet = 0; inet < n; inet++) { free(S.tr_node[inet]); free(S.tr_sw[inet]); }
	free(S.tr_node); free(S.tr_sw); free(S.tr_n); free(S.tr_cap);
	free(W); free(th); free(sinks); free(net_index); free(work); free(crit); free(net_delay); free(stats);
	return rc;
}

#ifdef PF_ORACLE_PAR_MAIN
#include <unistd.h>
/* usage: pf_oracle_par_cli problem.pfp [--threads T] [--congested-only] [--result out.pfr] [--max_iters K] */
int main(int argc, char **argv) {
	const char *result_path = NULL;
	int threads = (int)sysconf(_SC_NPROCESSORS_ONLN), max_iters = -1, congested_only = 0, i, rc;
	pf_problem p;
	pf_result out;
	double *secs, total = 0;
	long routed = 0;
	if (argc < 2) { fprintf(stderr, "usage: %s problem.pfp [--threads T] [--result o.pfr] [--max_iters K]\n", argv[0]); return 2; }
	for (i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--result") && i + 1 < argc) result_path = argv[++i];
		else if (!strcmp(argv[i], "--congested-only")) congested_only = 1;
		else if (!strcmp(argv[i], "--max_iters") && i + 1 < argc) max_iters = atoi(argv[++i]);
		else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
	}
	if (pf_problem_read(argv[1], &p) != 0) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
	p.opts.timing_analysis_enabled = 0;
	secs = (double *)calloc((size_t)(max_iters > 0 ? max_iters : p.opts.max_router_iterations) + 1, sizeof(double));
	rc = pf_oracle_route_parallel(&p, threads, congested_only, max_iters, &out, secs);
	if (rc != 0) { fprintf(stderr, "PF_ORACLE_PAR route failed rc=%d\n", rc); return 1; }
	for (i = 0; i < out.iterations; i++) {
		total += secs[i]; routed += out.iter_stats[i].nets_routed;
		fprintf(stderr, "PF_ORACLE_PAR iter %d dt_s=%.6f overused=%d nets=%d\n", i + 1, secs[i], out.iter_stats[i].overused_nodes, out.iter_stats[i].nets_routed);
	}
	fprintf(stderr, "PF_ORACLE_PAR policy=%s\n", congested_only ? "congested-only" : "all-nets");
	fprintf(stderr, "PF_ORACLE_PAR threads=%d success=%d iterations=%d wirelength=%d route_time_s=%.6f nets_routed=%ld nets_per_s=%.1f\n",
			threads, out.success, out.iterations, out.total_wirelength, total, routed, total > 0 ? routed / total : 0.);
	if (result_path && pf_result_write(result_path, &out) != 0) { fprintf(stderr, "cannot write %s\n", result_path); return 2; }
	return out.success ? 0 : 1;
}
#endif
