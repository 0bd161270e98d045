/*
 * pf_backend.h — the seam between the host-side PathFinder driver (pf_router.cpp) and the device.
 * The product implementation is pf_kernels.cu (CUDA, sm_100a).  tests/emu/pf_backend_emu.cpp
 * implements the same functions over the fiber warp emulator so that the identical driver and
 * device source can be exercised on a CPU-only box; it is test infrastructure, built only by
 * tests/ into its own library, never linked into libpf_router.so.
 */
#ifndef PF_BACKEND_H
#define PF_BACKEND_H

#include <stddef.h>
#include "pf_layout.h"

struct PfLaunchTimes {            /* accumulated device time (CUDA events on the router's stream) */
	double route_ms, update_ms, aux_ms;
	long long route_launches, update_launches, aux_launches;
};

int pfb_init(int device);                         /* select device, create the stream; <0 on failure */
int pfb_device_count(void);
const char *pfb_name(void);                       /* "cuda:sm_100a" | "emu" */
const char *pfb_last_error(void);
void *pfb_alloc(size_t bytes);                    /* device memory, zero-filled; NULL on failure */
void *pfb_alloc_raw(size_t bytes);                /* device memory, uninitialised */
void *pfb_host_alloc(size_t bytes);               /* small pinned host buffer (control-block reads); NULL on failure */
void pfb_host_free(void *p);
void *pfb_pinned(size_t bytes);                   /* process-wide pinned staging buffer of at least `bytes`; NULL if unavailable */
void *pfb_pinned_upload(size_t bytes);            /* same, write-combined: host code must only WRITE it, sequentially */
int pfb_h2d_async(void *dst, const void *src, size_t bytes);   /* ordered on the router's stream; src must stay valid until pfb_sync */
int pfb_d2h_async(void *dst, const void *src, size_t bytes);
void pfb_free(void *p);
int pfb_h2d(void *dst, const void *src, size_t bytes);
int pfb_d2h(void *dst, const void *src, size_t bytes);
int pfb_d2d(void *dst, const void *src, size_t bytes);
int pfb_zero(void *dst, size_t bytes);
int pfb_fill(void *dst, int byte, size_t bytes);
int pfb_sync(void);
void pfb_times(PfLaunchTimes *out, int reset);
int pfb_num_sms(void);
int pfb_timer_start(void);                        /* CUDA events on the router's stream */
int pfb_timer_stop(double *ms);
void *pfb_stream(void);
/* phase marks: an event on the router's stream now; pfb_marks_read returns, after a sync, the device time in ms between
 * consecutive marks since the last read and forgets them (profiling aid of pf_route_run, PF_PHASES=1) */
int pfb_mark(void);
int pfb_marks_read(double *ms, int cap);

/* the warp-per-net router: num_slots warps, each looping over the work queue */
int pfb_launch_route(const PfParams *P, int num_slots, int warps_per_block);
/* pathfinder_update_cost + feasible_routing (route_common.c:581-610,509-531) in one pass; when
 * base/delta are non-NULL the pass first folds the all-reduced occupancy delta into the node
 * records: occ = base + delta, base = occ (multi-GPU iteration boundary) */
int pfb_launch_update_cost(PfNode *nodes, int num_nodes, float acc_fac, int *d_overused, unsigned char *last_over, int iter_tag,
		unsigned long long *d_wl_used);   /* d_wl_used (may be NULL) += occupancy x length over the CHANX / CHANY nodes */
/* delta[i] = nodes[i].occ - base[i]: what this GPU's nets changed since the last sync */
/* replay another rank's occupancy events (multi-GPU sync) */
int pfb_launch_apply_events(PfNode *nodes, const unsigned *events, long long count);
/* route trees → s_trace-ordered arrays on the device: pass 1 (trace_node == NULL) writes len[net];
 * pass 2 writes trace_node/trace_switch at tptr[net] and adds the wirelength into *d_wl */
int pfb_launch_build_traces(const PfTreeNode *pool, const PfNetLoc *loc, int num_nets, int *len, const int *tptr,
		int *trace_node, short *trace_switch, unsigned long long *d_wl, unsigned *trace_term, const short *ptc, int nx);
/* occ_out[i] = nodes[i].occ (compact copy for the host) */
int pfb_launch_extract_occ(const PfNode *nodes, int num_nodes, int *occ_out);
/* total wirelength of all trees in the route store (route_timing.c:189-225 sanity abort) */
int pfb_launch_wirelength(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, unsigned long long *d_out);
/* reserve_locally_used_opins (route_common.c:1435-1491): one thread per (block, class) group */
int pfb_launch_reserve_opins(PfNode *nodes, const uint32_t *edges, int node_bits, const PfIndexedDev *indexed,
		int num_groups, const int *group_source, const int *group_count, const int *group_off,
		int *chosen, int rip_up, float pres_fac);
/* work list of the next iteration: the nets of `all_nets` that touch an overused node (or every
 * net when force_all), split into the small/big slot classes by net_big[]; counts[0]/counts[1] */
int pfb_launch_select_nets(const PfNode *nodes, const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets,
		int num_all, const unsigned char *net_big, int force_all, int *list_small, int *list_big, int *counts,
		const unsigned char *last_over, int iter_tag, int window, const int *committer, int *scratch, int head_count,
		int *queued, int queued_tag,
		const int *pool_node, const unsigned char *over_now, int over_tag);   /* over_now != NULL: fast test (pf_net_is_congested_fast) */
		/* queued != NULL: queued[net] = queued_tag for every selected net (ripple re-routing) */
/* counts[0..1] = lengths of list_small / list_big; counts[2..3] = how many of each come from the first head_count
 * entries of all_nets (they are at the head of the lists: order is preserved) */
void pfb_bind_thread(void);                     /* make the router's device current in a helper thread */
size_t pfb_select_scratch_bytes(int num_all);   /* size of `scratch` (device memory) */
/* copy every live tree of `all_nets` from one log to another (garbage collection of the route store) */
int pfb_launch_compact(const PfTreeNode *src, PfTreeNode *dst, const int *src_node, int *dst_node, PfNetLoc *loc, const int *all_nets, int num_all,
		unsigned long long *dst_head);

/* owner[node] = a net of `all_nets` whose tree in the route store contains the node (ripple re-routing: who gets displaced) */
int pfb_launch_rebuild_owner(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, int *owner);

/* ---- rr graph built on the device (pf_gen_device.cuh) */
struct PfGenDev;
/* pass 1: out-degree of every node into row[0..N), exclusive prefix sum in place (row[v] = start of v's edge row), total in
 * *num_edges (host).  G->cb_inv is a HOST pointer here; the backend stages it. */
int pfb_gen_count(const PfGenDev *G, int *row, long long *num_edges);
/* pass 2: node records, packed edge words, ptc numbers; *avail_wl (host) = total wirelength of the CHANX / CHANY nodes */
int pfb_gen_fill(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc, long long *avail_wl);
/* the same pass in two halves: _begin launches it on a side stream ordered after everything issued so far (the 4 ms kernel of
 * cfg 4 then runs under the host work and the small uploads of the rest of pf_router_create), _end makes the router's stream
 * wait for it and returns the wirelength.  _end without a pending _begin is a no-op that leaves *avail_wl alone; nothing may
 * touch nodes / edges / ptc / row on the router's stream, or free them, between the two. */
int pfb_gen_fill_begin(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc);
int pfb_gen_fill_end(long long *avail_wl);
/* occ = 0, acc_cost = 1 on every node record: a fresh first iteration (pf_router_reset) */
int pfb_reset_nodes(PfNode *nodes, int num_nodes);
/* order-independent 64-bit hashes of the node records / edge words / ptc numbers (tests: generated == uploaded graph) */
int pfb_graph_hash(const PfNode *nodes, int num_nodes, const uint32_t *edges, long long num_edges, const short *ptc, unsigned long long out[3]);

/* ---- multi-GPU exchange over peer memory (PfXchgHeader, pf_layout.h) */
/* device memory other processes of the node can map: returns the pointer and fills a 64-byte handle; NULL on failure */
void *pfb_ipc_alloc(size_t bytes, void *handle64);
void *pfb_ipc_open(const void *handle64);          /* map another process's region; NULL on failure */
void pfb_ipc_close(void *p);
void pfb_ipc_free(void *p);
int pfb_ipc_clear_abort(void *region);                /* a cached region goes to a new router: PfXchgHeader.abort_flag = 0, nothing else touched */
/* publish this rank's event log of exchange `seq` (count read from *event_head on the device), wait for every peer's,
 * replay the peers' events on `nodes` — one launch, no host involvement */
int pfb_launch_xchg_events(PfNode *nodes, const PfPeers *peers, int me, int nranks, unsigned seq, const unsigned long long *event_head,
		long long event_cap, int *status, double timeout_s);
/* sink delays: copy the delays of this rank's nets into its published buffer, release it, wait for the peers' and copy the
 * delays of their nets into net_delay (term_owner[t] = rank that routes terminal t's net) */
int pfb_launch_xchg_delays(float *net_delay, const unsigned char *term_owner, int num_terminals, const PfPeers *peers, int me, int nranks,
		unsigned dseq, long long event_cap, int *status, double timeout_s);
int pfb_launch_xchg_abort(const PfPeers *peers, int me);

/* ---- device static timing analysis (pf_sta_device.cuh, pf_sta.cpp) */
struct PfStaDev;
int pfb_sta_load(const PfStaDev *S, const float *dev_net_delay);
/* T_arr / T_req of every tnode and the three statistics of a domain pair back to "unset" */
int pfb_sta_begin_pair(const PfStaDev *S, float *stat);
/* levels [lv_begin, lv_end) in ascending (forward) or descending order; spread != 0: ONE level of that many
 * tnodes, spread over many CTAs */
int pfb_sta_sweep(const PfStaDev *S, int forward, int lv_begin, int lv_end, int spread, int domain, float constraint, float *stat);
int pfb_sta_update(const PfStaDev *S, float constraint, const float *stat, float *dev_crit);

/* check_route over a finished routing (pf_check_net): report[0] bad nets, [1] lowest bad net (init INT_MAX), [2] its
 * code, [3] rr nodes whose occupancy is not explained by the traces, [4] overused rr nodes; wl_extra[0] wirelength,
 * [1] occupancy not accounted for by the traces (must equal the locally used OPINs) */
int pfb_launch_check_route(const PfNode *nodes, const uint32_t *edges, int node_bits, int num_nodes, int num_nets, const int *net_ptr, const int *net_term,
		const unsigned char *net_is_global, const int *trace_ptr, const int *trace_node, const short *trace_switch, unsigned char *matched,
		int *occ2, const int *occ_reported, int *report, unsigned long long *wl_extra);

#endif
