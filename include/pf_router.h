/*
 * pf_router.h — C-ABI of the B200 PathFinder router (libpf_router.so).
 *
 * This is the drop-in boundary for the `--route` hot path of chinhau5/parallel_eda (VPR 7.0).
 * The reference has no plugin API: the seam is the link-time symbol
 *
 *     boolean try_timing_driven_route(struct s_router_opts router_opts, float **net_delay,
 *             t_slack *slacks, t_ivec **clb_opins_used_locally, boolean timing_analysis_enabled);
 *             // decl vpr/SRC/route/route_timing.h:1, sole caller vpr/SRC/route/route_common.c:500
 *
 * whose inputs and outputs are global AoS structures.  A reference-side adapter (INTEGRATION.md,
 * integration/vpr_adapter.cxx) flattens those globals into a pf_problem (pf_types.h) and calls
 * the entry points below; every signature uses plain pointers and sizes only.
 *
 *   entry point                      replaces (reference file:line)
 *   -------------------------------------------------------------------------------------------
 *   pf_try_timing_driven_route       try_timing_driven_route           route/route_timing.c:85-343
 *   pf_router_create                 alloc_and_load_rr_node_route_structs route/route_common.c:1012,
 *                                    rr_node[] storage → device CSR     route/rr_graph.c:521,1532
 *   pf_route_iteration               the net loop calling timing_driven_route_net
 *                                                                       route/route_timing.c:161-183,399-563
 *   pf_reserve_opins                 reserve_locally_used_opins         route/route_common.c:1435-1491
 *   pf_update_costs                  feasible_routing + pathfinder_update_cost
 *                                                                       route/route_common.c:509-531,581-610
 *   pf_total_wirelength              first-iteration wirelength abort   route/route_timing.c:189-225
 *   pf_get_net_delay                 update_net_delays_from_route_tree  route/route_tree_timing.c:515-528
 *   pf_get_result                    trace_head[]/trace_tail[] lists    route/route_common.c:638-706
 *   pf_comm_events /                 MPI_Allreduce(occupancy) of the reference's MPI router
 *   pf_comm_apply_events             parallel_route/spatial.cxx:3371-3383 (sync_recalc_occ)
 *   pf_try_breadth_first_route       try_breadth_first_route            route/route_breadth_first.c:23-305
 *   pf_sta_create / pf_sta_analyze   load_timing_graph_net_delays + do_timing_analysis + get_critical_path_delay
 *   pf_try_timing_driven_route_sta                                       timing/path_delay.c:479,2258-2522,3791 (route_timing.c:295-309)
 *   pf_sta_analyze_final             do_timing_analysis(.., is_final_analysis TRUE) of routing_stats   base/stats.c:155-164
 *   pf_check_route                   check_route                        route/check_route.c:27-155
 *
 * All functions return PF_OK (0) or a negative PF_E* code (pf_file.h); none calls exit().
 * pf_last_error() describes the last failure of the calling process.  There is no CPU fallback:
 * without a visible sm_100 device pf_router_create fails with PF_ECUDA.
 */
#ifndef PF_ROUTER_H
#define PF_ROUTER_H

#include "pf_types.h"
#include "pf_file.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_router pf_router;

typedef struct pf_config {
	int32_t device;           /* CUDA device ordinal */
	int32_t rank, nranks;     /* net sharding (one process per GPU): nets are cut into nranks spatial stripes of
	                             equal total fanout along x; this process routes stripe `rank` */
	int32_t num_slots;        /* concurrent warps (nets in flight); 0 = 20 per SM (one full wave) */
	int32_t warps_per_block;  /* 0 = 4 (maximum) */
	int32_t label_log2;       /* (fixed) the regular slots keep 2^10 hot label entries in shared memory */
	int32_t label2_log2;      /* per-slot fallback label table in global memory, 2^n entries; 0 = 13, < 0 none */
	int32_t tree_cap;         /* per-warp route-tree entries; 0 = 2048 */
	int32_t far_cap;          /* per-warp far-list entries; 0 = 8192 */
	int32_t sink_cap;         /* nets with more sinks go to the big slots; 0 = 64 */
	int32_t big_slots;        /* warps with large scratch for big / overflowed nets; 0 = 64 */
	int32_t big_label_log2, big_tree_cap, big_far_cap;   /* 0 = sized from the problem */
	int32_t max_batch;        /* labels settled per step (delta bucket), 1..32; 0 = auto: 1 when nets >> warps
	                             (traffic-bound: strict best-first does the least work), 32 otherwise (latency-bound) */
	float pop_slack;          /* delta-stepping bucket width in units of the cheapest edge cost; <0 = auto (0 / 0.25) */
	float win_rel, win_abs;   /* near-set window; 0 = auto */
	int32_t verbose;
	int32_t reroute_all_iters;/* the first K iterations re-route every net (the serial reference re-routes
	                             every net in every iteration, route_timing.c:161-183); later iterations
	                             re-route only nets that touch an overused rr node, like "phase two" of the
	                             reference's parallel router (partitioning_multi_sink…cxx:6241-6269).
	                             0 = auto (1); < 0 = always every net */
	int32_t inflight_div;     /* nets in flight <= ceil(nets this iteration / inflight_div): bounds how
	                             stale the congestion seen by concurrent nets can be; 0 = auto (16 in the first iteration, then 2 if it
	                             has shown the channels to be under 40 % full, else 32) */
	int32_t min_slots;        /* lower bound for the above; 0 = auto (one net per 20 x 20 tiles of the grid, at least 1) */
	int32_t stall_iters;      /* overuse not down by 30 % over stall_iters+1 congested-only iterations => one
	                             iteration re-routes every net with 8x fewer nets in flight; 0 = auto (3); < 0 = never */
	int32_t history_window;   /* also re-route nets holding a node that was overused within the last K cost updates;
	                             0 = off (default) */
	int32_t keep_newcomer;    /* experimental: on an overused node, re-route every user except the one that committed it
	                             last; 0 = off (default) */
	int32_t defer_graph;      /* multi-GPU: allocate the device graph but do not pack and upload it from this process's
	                             host arrays — the caller fills it from another rank's copy over NVLink
	                             (pf_comm_graph_buffers, then pf_comm_graph_ready); 0 = upload here (default) */
	int32_t validate_commits; /* optimistic concurrency control for nets in flight together: a new path whose commit finds an
	                             rr node already full although its search saw it free lost a race against another net in
	                             flight; it is taken back and the sink searched again on the current occupancy (what the
	                             serial order would have seen), at most this many times per sink.  0 = auto (2), < 0 = off */
	int32_t ripple;           /* ripple re-routing inside an iteration: a net that knowingly shares a full rr node pushes the net
	                             holding it onto this iteration's work queue, so a chain of displacements is followed within one
	                             PathFinder iteration, as in the serial reference where every net is re-routed every iteration;
	                             0 = auto (on), < 0 = off */
	int32_t ripple_max_nets;  /* ripple only in iterations that re-route at most this many nets: displacement chains are followed one
	                             link after the other, which is what shortens the tail of the negotiation but would serialise an
	                             iteration with tens of thousands of nets; 0 = auto: max(1024, min(nets / 16, 4 * min_slots)) */
	int32_t polish;           /* > 0: when the routing first becomes legal, one more iteration re-routes EVERY net against the
	                             final congestion picture and the loop continues until legal again (pf_try_* loops only) */
	int32_t lazy_seed_min;    /* big slots: a search whose route tree holds at least this many entries labels only the seeds it can
	                             need and comes back for more (reference: add_route_tree_to_heap pushes the whole tree for every
	                             sink, route_timing.c:590-646 — a 1760-sink net then spends its time on seeds that are never
	                             popped); 0 = auto (256), < 0 = off */
} pf_config;

typedef struct pf_timing {    /* accumulated since create / last reset */
	double route_kernel_ms;   /* CUDA-event time of pf_route_kernel launches */
	double update_kernel_ms;  /* pf_update_cost_kernel */
	double aux_kernel_ms;     /* delta export, wirelength, OPIN reservation */
	int64_t route_launches, update_launches, aux_launches;
	int64_t h2d_bytes, d2h_bytes;
} pf_timing;

/* Host STA hook, called between iterations when opts.timing_analysis_enabled (the reference
 * calls load_timing_graph_net_delays + do_timing_analysis, route_timing.c:295-309):
 * in  net_delay[num_terminals]  out crit[num_terminals] (timing_criticality), *crit_path_delay */
typedef void (*pf_sta_fn)(void *user, int iters_done, const float *net_delay, float *crit, float *crit_path_delay);

void pf_config_default(pf_config *cfg);
const char *pf_last_error(void);
const char *pf_backend_name(void);
int pf_device_count(void);

int pf_router_create(const pf_problem *p, const pf_config *cfg, pf_router **out);
void pf_router_destroy(pf_router *r);
/* The same for a fabric described by generator parameters (pf_gen.h): `nets` comes from pf_gen_grid_nets (no node / edge
 * arrays) and the rr graph is built ON the device (rr_graph.c:385 build_rr_graph's role) — nothing but the
 * nets crosses PCIe.  Bit-identical to creating from pf_gen_grid_problem's arrays (pf_debug_graph_hash). */
struct pf_gen_params;
int pf_router_create_generated(const struct pf_gen_params *g, const pf_problem *nets, const pf_config *cfg, pf_router **out);
/* order-independent hashes of the device graph: node records, edge words, ptc numbers (tests) */
int pf_debug_graph_hash(pf_router *r, uint64_t out[3], int64_t *num_edges);
/* forget all routing and congestion history (occ = 0, acc_cost = 1): a fresh first iteration */
int pf_router_reset(pf_router *r);

/* One PathFinder iteration over this rank's nets: rip-up, route, commit (live occupancy).
 * crit: host array [num_terminals] or NULL to keep the criticalities already on the device. */
int pf_route_iteration(pf_router *r, float pres_fac, const float *crit, pf_iter_stats *stats);
/* The same in two steps, for multi-GPU sub-rounds: begin chooses this iteration's nets, route_part routes
 * slice `part` of `nparts` of them; the caller syncs occupancy (export delta / all-reduce / fold) after each part. */
int pf_iteration_begin(pf_router *r, const float *crit);
int pf_iteration_route_part(pf_router *r, float pres_fac, int part, int nparts, pf_iter_stats *stats);
int pf_reserve_opins(pf_router *r, float pres_fac, int rip_up_local_opins);
int pf_update_costs(pf_router *r, float acc_fac, int *overused_nodes);
int pf_total_wirelength(pf_router *r, int64_t *wirelength, int64_t *available);
int pf_get_net_delay(pf_router *r, float *net_delay);
/* Fills trace_*, net_delay, occ, serial_num, total_wirelength (free with pf_result_free). */
int pf_get_result(pf_router *r, pf_result *out);
int pf_get_timing(pf_router *r, pf_timing *t, int reset);
/* Device-side stopwatch: CUDA events recorded on the router's own stream (whole-step timing). */
int pf_timer_start(pf_router *r);
int pf_timer_stop(pf_router *r, double *elapsed_ms);
void *pf_stream(pf_router *r);   /* the cudaStream_t every kernel of this router is launched on */

/* Multi-GPU occupancy sync after a route part (routers created with nranks > 1).  While routing, every
 * occupancy change of this rank's nets (rip-up, commit, undo) is also appended to an event log in device
 * memory: uint32 per event, rr node id in bits 0..30, bit 31 set = decrement.
 *   1. pf_comm_events(r, &ptr, &n)          this rank's log of the last route part (device pointer, n events)
 *   2. all-gather the logs across ranks     (NCCL over NVLink; the caller owns the communicator)
 *   3. pf_comm_apply_events(r, ptr_k, n_k)  replay every OTHER rank's log on this rank's node records
 * afterwards all ranks hold the same occupancy, and pf_reserve_opins / pf_update_costs run as on one GPU.
 * pf_comm_net_delay_ptr returns the device float[num_terminals] delay vector; entries of nets
 * routed by other ranks are zero, so an all-reduce(sum) assembles the full vector in place. */
/* Multi-GPU create: every rank needs the same packed graph (rr node records, edge words, ptc numbers).  Rank 0
 * packs and uploads it over PCIe; the other ranks are created with cfg.defer_graph = 1 and receive it by an NCCL
 * broadcast over NVLink into the three device buffers named here, then call pf_comm_graph_ready.  (All ranks still
 * pass the same pf_problem: nets, options and the reset path use the host arrays.) */
int pf_comm_graph_buffers(pf_router *r, void *dev_ptrs[3], int64_t bytes[3]);
int pf_comm_graph_ready(pf_router *r);
/* the sharding decided at create: owner[num_nets] = rank routing the net, is_cut[num_nets] = 1 for nets whose bounding
 * box reaches across a stripe cut (routed in the second part of an iteration), 0 for stripe-interior nets */
int pf_comm_net_classes(pf_router *r, int32_t *owner, uint8_t *is_cut);
int pf_comm_events(pf_router *r, void **dev_events, int64_t *count);
int pf_comm_apply_events(pf_router *r, const void *dev_events, int64_t count);
void *pf_comm_net_delay_ptr(pf_router *r);

/* ---- The transport inside the library (SURVEY.md §8b `pf_comm_init`; the reference syncs in C as well,
 * parallel_route/spatial.cxx:3371-3383): one process per GPU on one node, every rank's exchange region (occupancy event
 * logs, published sink delays) mapped into the others through CUDA IPC and read over NVLink / NVSwitch by the kernels
 * themselves.  Bootstrap once per router:
 *   1. pf_comm_export(r, blob)            PF_COMM_HANDLE_BYTES describing this rank's region
 *   2. all-gather the blobs in rank order (the caller's communicator: MPI_Allgather, torch.distributed, ...)
 *   3. pf_comm_init(r, all_blobs)
 * afterwards pf_comm_exchange (after every route part) and pf_comm_gather_delays (before a timing analysis) are
 * stream-ordered device work — publish with a system-scope release, poll the peers with acquire loads, read their
 * payload out of their memory — and pf_route_run runs whole multi-GPU routings with one host synchronisation per
 * PathFinder iteration.  pf_comm_abort makes peers stop waiting for a rank that failed.
 * The transport outlives the router, like the reference's MPI communicator outlives one routing: the exchange region of a
 * destroyed router is kept by the process and given to the next router it fits, regions of peers stay mapped (steps 1-3 are
 * still run per router, but then cost microseconds: no cudaMalloc, no cudaIpcOpenMemHandle), sequence numbers continue.
 * pf_comm_release_cache() frees what no live router uses — call it on every rank after the last router of the job. */
#define PF_COMM_HANDLE_BYTES 128
int pf_comm_export(pf_router *r, void *handle);
int pf_comm_init(pf_router *r, const void *all_handles);
int pf_comm_exchange(pf_router *r);
int pf_comm_gather_delays(pf_router *r);
int pf_comm_abort(pf_router *r);
int pf_comm_release_cache(void);

/* try_timing_driven_route (route_timing.c:85-343) on an existing router, one GPU or — after pf_comm_init — one rank of
 * several: iterate until legal or out of iterations, with ONE host-device synchronisation per iteration.  dsta: device
 * timing analysis (or NULL), sta/user: host analysis callback (or NULL).  stats[stats_cap] receives one entry per
 * iteration (may be NULL); every rank returns the same *iterations and *success.  The routing stays in the router
 * (pf_get_result). */
int pf_route_run(pf_router *r, struct pf_sta *dsta, pf_sta_fn sta, void *user, pf_iter_stats *stats, int stats_cap,
		int *iterations, int *success);
/* the device float[num_terminals] criticality vector the next iteration reads (pf_iteration_begin with crit == NULL
 * keeps it): a device STA writes it in place (pf_sta_analyze_device) */
void *pf_comm_crit_ptr(pf_router *r);

/* ---- Device static timing analysis: what the reference runs on the host between iterations,
 * load_timing_graph_net_delays + do_timing_analysis + get_critical_path_delay (route_timing.c:295-309;
 * timing/path_delay.c:479, 2258, 3791), on the flat timing graph of pf_types.h.  The criticalities are
 * bit-identical to the reference's (tests/test_sta_golden.py, tests/test_gpu_sta.py). */
typedef struct pf_sta pf_sta;
int pf_sta_create(const pf_timing_graph *g, const pf_problem *p, const pf_config *cfg, pf_sta **out);
void pf_sta_destroy(pf_sta *s);
/* host buffers: net_delay[num_terminals] in, crit[num_terminals] out, *cpd_ns = critical path delay in ns */
int pf_sta_analyze(pf_sta *s, const float *net_delay, float *crit, float *cpd_ns);
/* device buffers (e.g. a router's own delay and criticality vectors: nothing crosses PCIe) */
int pf_sta_analyze_device(pf_sta *s, const void *dev_net_delay, void *dev_crit, float *cpd_ns);
/* the critical path delay (ns) of the last analysis — pf_sta_analyze_device with cpd_ns == NULL does not wait for it */
int pf_sta_read_cpd(pf_sta *s, float *cpd_ns);
/* do_timing_analysis(slacks, FALSE, FALSE, is_final_analysis = TRUE) as routing_stats runs it on the finished routing
 * (base/stats.c:155-164, timing/path_delay.c:2258-2522 with :2786-2790 and update_slack :3117-3125): host buffers of
 * num_terminals floats; slack[t] = least slack of net pin t over the analysed constraints (real required times, so it may be
 * negative; 1e30 = HUGE_POSITIVE_FLOAT where no analysed path passes), crit (may be NULL) = the criticalities of that analysis,
 * *cpd_ns the critical path delay (get_critical_path_delay, :3791) */
int pf_sta_analyze_final(pf_sta *s, const float *net_delay, float *slack, float *crit, float *cpd_ns);
/* try_timing_driven_route with the analysis on the device: no host callback, no per-iteration copies */
int pf_try_timing_driven_route_sta(const pf_problem *p, const pf_timing_graph *g, const pf_config *cfg, pf_result *out);

/* check_route (route/check_route.c:27-155) on the device, for ANY routing of the router's problem handed over as a
 * pf_result (the router's own, the oracle's, the reference's): traces start at the SOURCE, segments end at the
 * net's SINKs, consecutive elements are rr edges with the recorded switch, later segments branch off the net, every
 * pin is reached once; the occupancy recomputed from the traces explains result.occ up to the locally used OPINs
 * and respects capacity.  Never looks at the router's state — only at its copy of the rr graph. */
typedef struct pf_check_report {
	int32_t ok;                  /* 1 iff everything below is clean (overused_nodes may be > 0 for an unfinished routing) */
	int32_t bad_nets;            /* nets with a structural violation */
	int32_t first_bad_net;       /* lowest such net, -1 if none */
	int32_t first_bad_code;      /* 1 no trace, 2 not at SOURCE, 3 open segment, 4 join not in net, 5 no such rr edge,
	                                6 SINK carries a switch, 7 wrong sinks, 8 node id out of range */
	int32_t occupancy_mismatch;  /* rr nodes whose reported occupancy the traces (+ OPIN reservations) do not explain */
	int32_t overused_nodes;      /* reported occupancy above capacity */
	int64_t wirelength;          /* of the structurally valid nets */
	int64_t reserved_opins;      /* occupancy not due to traces (must equal the sum of opin_group_count) */
} pf_check_report;
int pf_check_route(pf_router *r, const pf_result *res, pf_check_report *rep);

/* try_breadth_first_route (route/route_breadth_first.c:23-91; `--router_algorithm breadth_first`): problems with
 * opts.router_algorithm == 1 are routed by one persistent maze wavefront per net (no lookahead, no delay term, the
 * new segment re-enters the wave at cost 0), acc_fac applies from the first iteration and there is no
 * first-iteration wirelength abort.  pf_try_timing_driven_route and the step functions dispatch on the same option. */
int pf_try_breadth_first_route(const pf_problem *p, const pf_config *cfg, pf_result *out);

/* The whole of try_timing_driven_route (single GPU): iterate until legal or out of iterations.
 * sta may be NULL when opts.timing_analysis_enabled == 0. */
int pf_try_timing_driven_route(const pf_problem *p, const pf_config *cfg, pf_sta_fn sta, void *user, pf_result *out);

#ifdef __cplusplus
}
#endif
#endif /* PF_ROUTER_H */
