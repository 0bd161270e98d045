/* TEST INFRASTRUCTURE: command-line driver for the CPU restatement (pf_oracle.c).
 * usage: pf_oracle_cli problem.pfp [--crit golden.pfr | --timing-graph g.pftg] [--result out.pfr] [--max_iters K] [--limit_nets M]
 *   --crit          timing-driven, replaying the criticalities the reference's STA produced (golden result)
 *   --timing-graph  timing-driven with the oracle's own restatement of the STA in the loop: router + analysis together
 *                   must then reproduce the reference's whole run */
#include "pf_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* crit_fn running pf_oracle_sta on the flat timing graph (user = this struct) */
struct sta_ctx { const pf_timing_graph *g; const pf_problem *p; };
static void sta_crit(void *user, int iters_done, const float *net_delay, float *crit, float *cpd) {
	struct sta_ctx *c = (struct sta_ctx *)user;
	float ns = 0.f;
	(void)iters_done;
	if (pf_oracle_sta(c->g, c->p->net_ptr, net_delay, crit, &ns) != 0) { fprintf(stderr, "PF_ORACLE sta failed\n"); exit(3); }
	*cpd = ns;
}

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
	const char *result_path = NULL, *crit_path = NULL, *tg_path = NULL;
	pf_timing_graph tg;
	struct sta_ctx sctx;
	int max_iters = -1, limit_nets = -1, i, rc;
	pf_problem p;
	pf_result golden, out;
	double t0, t1;
	long pushes = 0, pops = 0, visits = 0, nets = 0;
	if (argc < 2) { fprintf(stderr, "usage: %s problem.pfp [--crit g.pfr] [--result o.pfr] [--max_iters K] [--limit_nets M]\n", argv[0]); return 2; }
	for (i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--result") && i + 1 < argc) result_path = argv[++i];
		else if (!strcmp(argv[i], "--crit") && i + 1 < argc) crit_path = argv[++i];
		else if (!strcmp(argv[i], "--timing-graph") && i + 1 < argc) tg_path = argv[++i];
		else if (!strcmp(argv[i], "--max_iters") && i + 1 < argc) max_iters = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--limit_nets") && i + 1 < argc) limit_nets = atoi(argv[++i]);
		else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
	}
	if ((rc = pf_problem_read(argv[1], &p)) != 0) { fprintf(stderr, "cannot read %s (%d)\n", argv[1], rc); return 2; }
	if (limit_nets >= 0) for (i = limit_nets; i < p.num_nets; i++) p.net_is_global[i] = 1;
	memset(&golden, 0, sizeof(golden));
	if (crit_path) {
		if (pf_result_read(crit_path, &golden) != 0 || golden.num_terminals != p.num_terminals) { fprintf(stderr, "bad --crit %s\n", crit_path); return 2; }
		p.opts.timing_analysis_enabled = 1;
	} else if (tg_path) {
		char msg[256];
		if (pf_timing_graph_read(tg_path, &tg) != 0 || pf_timing_graph_check(&tg, p.net_ptr, msg, sizeof(msg)) != 0 || tg.num_nets != p.num_nets) {
			fprintf(stderr, "bad --timing-graph %s\n", tg_path); return 2;
		}
		sctx.g = &tg; sctx.p = &p;
		p.opts.timing_analysis_enabled = 1;
	} else {
		p.opts.timing_analysis_enabled = 0;
	}
	t0 = now_s();
	if (tg_path && !crit_path) rc = pf_oracle_route(&p, sta_crit, &sctx, max_iters, &out);
	else rc = pf_oracle_route(&p, crit_path ? pf_oracle_replay_crit : NULL, &golden, max_iters, &out);
	t1 = now_s();
	if (rc != 0) { fprintf(stderr, "PF_ORACLE route failed rc=%d\n", rc); return 3; }
	for (i = 0; i < out.num_iter_stats; i++) {
		pushes += out.iter_stats[i].heap_pushes; pops += out.iter_stats[i].heap_pops;
		visits += out.iter_stats[i].edge_visits; nets += out.iter_stats[i].nets_routed;
		fprintf(stderr, "PF_ORACLE iter %d overused=%d pushes=%ld pops=%ld visits=%ld\n", i + 1, out.iter_stats[i].overused_nodes,
				(long)out.iter_stats[i].heap_pushes, (long)out.iter_stats[i].heap_pops, (long)out.iter_stats[i].edge_visits);
	}
	fprintf(stderr, "PF_ORACLE route success=%d iterations=%d cookie=%d wirelength=%d route_time_s=%.6f nets_routed=%ld pushes=%ld pops=%ld visits=%ld\n",
			out.success, out.iterations, out.serial_num, out.total_wirelength, t1 - t0, nets, pushes, pops, visits);
	if (result_path && pf_result_write(result_path, &out) != 0) { fprintf(stderr, "cannot write %s\n", result_path); return 2; }
	pf_result_free(&out);
	if (crit_path) pf_result_free(&golden);
	pf_problem_free(&p);
	return 0;
}
