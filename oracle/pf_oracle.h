/*
 * pf_oracle.h — TEST INFRASTRUCTURE.  CPU restatement of the reference's serial timing-driven
 * PathFinder router (reference vpr/SRC/route/route_timing.c, route_common.c, route_tree_timing.c,
 * util/heapsort.c) over the flat pf_problem arrays.  See pf_oracle.c for the parity statement.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H

#include "../include/pf_file.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Stand-in for the host STA between iterations (route_timing.c:295-309): called after
 * `iters_done` iterations with the per-terminal net delays; must fill crit[num_terminals]
 * (timing_criticality, aligned with net_terminals) and *cpd (critical path delay, seconds). */
typedef void (*pf_oracle_crit_fn)(void *user, int iters_done, const float *net_delay, float *crit, float *cpd);

/* try_timing_driven_route (route_timing.c:85-343).  max_iters_override <= 0 uses
 * p->opts.max_router_iterations.  Returns PF_OK (0) and fills *out (free with pf_result_free)
 * whether or not a legal routing was found (out->success); PF_EUNROUTABLE if a net has no path. */
int pf_oracle_route(const pf_problem *p, pf_oracle_crit_fn crit_fn, void *user, int max_iters_override,
		pf_result *out);

/* crit_fn that replays golden->iter_crit (user = const pf_result *golden). */
void pf_oracle_replay_crit(void *user, int iters_done, const float *net_delay, float *crit, float *cpd);

/* heapsort (util/heapsort.c:13): indices of decreasing values, reference tie order. */
void pf_oracle_heapsort(int *sort_index, float *sort_values, int nelem, int start_index);

/* get_serial_num (route_common.c:224-254) over a flat trace. */
int pf_serial_num(const pf_problem *p, const int32_t *trace_ptr, const int32_t *trace_node);

/* do_timing_analysis(slacks, FALSE, FALSE, FALSE) as the router calls it between iterations (route_timing.c:295-309;
 * timing/path_delay.c:2258-2522 with SLACK_DEFINITION 'R', no PATH_COUNTING, not prepacked, no LUT rebalancing, not the
 * final analysis) preceded by load_timing_graph_net_delays (path_delay.c:479-500), over the flat timing graph.
 * net_ptr = pf_problem.net_ptr.  Fills crit[num_terminals] (timing_criticality; 0 where the reference leaves 0) and
 * *cpd_ns (get_critical_path_delay, path_delay.c:3791-3810).  scratch-free: allocates its own T_arr / T_req. */
int pf_oracle_sta(const pf_timing_graph *g, const int32_t *net_ptr, const float *net_delay, float *crit, float *cpd_ns);
/* the analysis of the finished routing (is_final_analysis = TRUE, base/stats.c:155-164): slack[num_terminals] out as well */
int pf_oracle_sta_final(const pf_timing_graph *g, const int32_t *net_ptr, const float *net_delay, float *slack, float *crit, float *cpd_ns);

#ifdef __cplusplus
}
#endif
#endif
