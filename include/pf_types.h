/*
 * pf_types.h — flat, VPR-free data model of one PathFinder routing problem and its result.
 *
 * Everything the `--route` hot path of chinhau5/parallel_eda reads lives in global AoS
 * structures (reference vpr/SRC/base/globals.c:48-97).  The drop-in boundary
 * flattens those globals into the plain arrays below; nothing here includes a VPR header.
 *
 *   reference global                         (file:line)                      field here
 *   ---------------------------------------------------------------------------------------
 *   nx, ny                                   globals.c:53-54                  nx, ny
 *   num_rr_nodes, rr_node[]                  globals.c:84-85, vpr_types.h:946 num_nodes, node SoA
 *   rr_node[i].edges[]/switches[]            vpr_types.h:977-978              row_ptr/edge_to/edge_sw (CSR,
 *                                                                             per-row order preserved:
 *                                                                             prev_edge is an index into it)
 *   switch_inf[]                             physical_types.h:744             switches[]
 *   num_rr_indexed_data, rr_indexed_data[]   vpr_types.h:1047                 indexed[]
 *   num_nets, clb_net[], net_rr_terminals    vpr_types.h:521, globals.c:78    net_ptr/net_terminals/net_is_global
 *   route_bb[]                               route_common.c:1065              net_bb
 *   clb_opins_used_locally + rr_blk_source   route_common.c:1435              opin_group_*
 *   struct s_router_opts                     vpr_types.h:742                  pf_router_opts
 *   trace_head[]/trace_tail[] (s_trace)      vpr_types.h:922, route_common.c:638  pf_result.trace_*
 *   net_delay[inet][ipin]                    route_tree_timing.c:515          pf_result.net_delay
 */
#ifndef PF_TYPES_H
#define PF_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rr node types: same numbering as t_rr_type (reference vpr/SRC/base/vpr_types.h:910). */
enum { PF_SOURCE = 0, PF_SINK = 1, PF_IPIN = 2, PF_OPIN = 3, PF_CHANX = 4, PF_CHANY = 5 };

/* cost-index rows: e_cost_indices (reference vpr_types.h:1087-1093). */
enum { PF_SOURCE_COST_INDEX = 0, PF_SINK_COST_INDEX = 1, PF_OPIN_COST_INDEX = 2,
       PF_IPIN_COST_INDEX = 3, PF_CHANX_COST_INDEX_START = 4 };

#define PF_NO_PREVIOUS (-1)            /* vpr_types.h:944 */
#define PF_OPEN (-1)                   /* iswitch of a SINK trace element */
#define PF_HUGE_POSITIVE_FLOAT 1.e30f  /* vpr_types.h:84 */
#define PF_HIGH_FANOUT_NET_LIM 64      /* vpr_types.h:91 */
#define PF_FIRST_ITER_WIRELENGTH_LIMIT 0.85f /* vpr_types.h:93 (FIRST_ITER_WIRELENTH_LIMIT) */

typedef struct pf_switch {      /* s_switch_inf, physical_types.h:744 */
	int32_t buffered;
	float R, Cin, Cout, Tdel;
} pf_switch;

typedef struct pf_indexed {     /* t_rr_indexed_data, vpr_types.h:1047 */
	float base_cost;
	float saved_base_cost;
	int32_t ortho_cost_index;
	int32_t seg_index;
	float inv_length;
	float T_linear;
	float T_quadratic;
	float C_load;
} pf_indexed;

typedef struct pf_router_opts { /* the s_router_opts fields the path reads, route_timing.c:140-170,283-292 */
	float first_iter_pres_fac;
	float initial_pres_fac;
	float pres_fac_mult;
	float acc_fac;
	float bend_cost;
	float astar_fac;
	float max_criticality;
	float criticality_exp;
	int32_t max_router_iterations;
	int32_t timing_analysis_enabled; /* boolean argument of try_timing_driven_route */
	int32_t bb_factor;               /* informational: net_bb is already expanded */
	int32_t router_algorithm; /* 0 = timing-driven / no-timing (route_timing.c), 1 = breadth-first (route_breadth_first.c) */
} pf_router_opts;

typedef struct pf_problem {
	int32_t nx, ny;
	int32_t num_nodes;
	int32_t num_edges;
	/* node SoA, [num_nodes] */
	int16_t *xlow, *ylow, *xhigh, *yhigh;
	int16_t *ptc_num;
	int16_t *cost_index;
	int16_t *capacity;
	uint8_t *type;       /* PF_SOURCE .. PF_CHANY */
	uint8_t *direction;  /* e_direction, vpr_types.h:855 (informational) */
	float *R, *C;
	/* CSR out-edges */
	int32_t *row_ptr;    /* [num_nodes+1] */
	int32_t *edge_to;    /* [num_edges] */
	int16_t *edge_sw;    /* [num_edges] index into switches[] */
	int32_t num_switches;
	pf_switch *switches;
	int32_t num_indexed;
	pf_indexed *indexed;
	/* nets: terminal 0 of a net is its SOURCE rr node, 1..num_sinks its SINK rr nodes */
	int32_t num_nets;
	int32_t num_terminals;   /* = net_ptr[num_nets] */
	int32_t *net_ptr;        /* [num_nets+1] */
	int32_t *net_terminals;  /* [num_terminals] */
	uint8_t *net_is_global;  /* [num_nets] global nets are never routed */
	int32_t *net_bb;         /* [num_nets][4] xmin,xmax,ymin,ymax (s_bb order, vpr_types.h:554) */
	/* reserve_locally_used_opins (route_common.c:1435): for each (block,class) that keeps
	 * `count` OPINs for intra-block use, the class SOURCE node whose out-edges are the OPINs */
	int32_t num_opin_groups;
	int32_t *opin_group_source;
	int32_t *opin_group_count;
	pf_router_opts opts;
} pf_problem;

/* One PathFinder iteration's counters (reference prints pushes route_timing.c:332; overuse is
 * what feasible_routing route_common.c:509 counts). */
typedef struct pf_iter_stats {
	int32_t overused_nodes;
	int32_t nets_routed;
	int64_t heap_pushes;      /* labels offered to the queue   (reference num_heap_pushes) */
	int64_t heap_pops;        /* labels taken from the queue                                */
	int64_t edge_visits;      /* out-edges examined in the expansion loop                   */
	float pres_fac;
	float crit_path_delay;    /* filled by the STA callback if any, else 0                  */
} pf_iter_stats;

typedef struct pf_result {
	int32_t success;          /* TRUE iff legal routing found within max_router_iterations */
	int32_t iterations;
	int32_t serial_num;       /* "magic cookie", route_common.c:224-254 */
	int32_t total_wirelength; /* stats.c:186-243 over CHANX/CHANY trace elements */
	/* traceback in the exact s_trace order of update_traceback (route_common.c:638):
	 * per net a concatenation of segments, each ending at a SINK (iswitch PF_OPEN); every
	 * segment after the first begins with the join node already in the routing. */
	int32_t num_nets;
	int32_t *trace_ptr;       /* [num_nets+1] */
	int32_t *trace_node;      /* [trace_ptr[num_nets]] */
	int16_t *trace_switch;    /* [trace_ptr[num_nets]] */
	int32_t num_terminals;
	float *net_delay;         /* [num_terminals], aligned with net_terminals (entry 0 of a net = 0) */
	int32_t num_nodes;
	int32_t *occ;             /* [num_nodes] final occupancy */
	int32_t num_iter_stats;
	pf_iter_stats *iter_stats;
	/* optional: timing criticalities used in each iteration (golden files of timing-driven
	 * reference runs): [iterations][num_terminals], iteration i (1-based) at (i-1)*num_terminals */
	int32_t num_crit_iters;
	float *iter_crit;
} pf_result;

/* ------------------------------------------------------------------ timing graph
 * Flat image of the reference's post-packing timing graph (tnode[] / tedge, vpr_types.h:300-380, built by
 * alloc_and_load_timing_graph, timing/path_delay.c:328) — what do_timing_analysis (path_delay.c:2258) reads when
 * the router calls it between iterations (route_timing.c:295-309).  Pin k (1-based) of net i is out-edge k-1 of
 * tnode net_driver[i] (path_delay.c:479-500); its delay comes from net_delay[net_ptr[i] + k] of the routing. */
#define PF_TN_FF_SINK 11      /* e_tnode_type values the analysis distinguishes (vpr_types.h:305-323) */
#define PF_TN_INPAD_SOURCE 0
#define PF_TN_OUTPAD_SINK 3
#define PF_TN_FF_SOURCE 12
#define PF_TN_FF_CLOCK 13
#define PF_TN_CONSTANT_GEN_SOURCE 14

typedef struct pf_timing_graph {
	int32_t num_tnodes, num_tedges;
	int32_t *edge_ptr;        /* [num_tnodes+1] CSR of tnode[].out_edges, the reference's order */
	int32_t *edge_to;         /* [num_tedges] tedge.to_node */
	float *edge_Tdel;         /* [num_tedges] tedge.Tdel; the edges of net drivers are overwritten from net_delay */
	uint8_t *type;            /* [num_tnodes] e_tnode_type */
	int32_t *clock_domain;    /* [num_tnodes] index into the constrained clocks, -1 = none */
	float *clock_delay;       /* [num_tnodes] */
	int32_t num_levels;
	int32_t *level_ptr;       /* [num_levels+1] tnodes_at_level (path_delay2.c) */
	int32_t *level_nodes;     /* [num_tnodes] */
	int32_t num_domains;      /* g_sdc->num_constrained_clocks */
	float *constraint;        /* [num_domains * num_domains] g_sdc->domain_constraint, < 0 = DO_NOT_ANALYSE */
	int32_t num_nets;         /* == pf_problem.num_nets */
	int32_t *net_driver;      /* [num_nets] f_net_to_driver_tnode */
	/* clock-to-flipflop override constraints (g_sdc->cf_constraints: set_max_delay / set_false_path / set_multicycle_path
	 * -from [get_clocks ..] -to <flip-flops or pads>), resolved to tnodes by the exporter: in the traversals whose source
	 * domain is override_domain[k], sink tnode override_tnode[k] takes override_constraint[k] (seconds; < 0 = this sink is not
	 * analysed) instead of the domain pair's constraint (find_cf_constraint, timing/path_delay.c:2753-2768).  Sorted by
	 * (tnode, domain), at most one entry per pair (the reference takes the first match); num_overrides == 0: none */
	int32_t num_overrides;
	int32_t *override_domain;
	int32_t *override_tnode;
	float *override_constraint;
} pf_timing_graph;

/* golden vectors of the reference's analysis: K calls, each net_delay in -> timing_criticality out */
typedef struct pf_sta_vectors {
	int32_t num_terminals, num_calls;
	float *net_delay;         /* [num_calls * num_terminals] */
	float *crit;              /* [num_calls * num_terminals] slacks->timing_criticality, 0 at terminal 0 of a net */
	float *cpd;               /* [num_calls] get_critical_path_delay() in ns */
} pf_sta_vectors;

#ifdef __cplusplus
}
#endif
#endif /* PF_TYPES_H */
