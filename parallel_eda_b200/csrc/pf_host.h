/*
 * pf_host.h — internals shared by the host-side translation units behind the C-ABI (pf_router.cpp, pf_sta.cpp,
 * pf_check.cpp): the error channel of pf_last_error(), the router handle, small host helpers.  Not installed.
 */
#ifndef PF_HOST_H
#define PF_HOST_H

#include "../../include/pf_router.h"
#include "pf_backend.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern char g_router_err[512];            /* pf_last_error() */
#define FAILF(code, ...) do { snprintf(g_router_err, sizeof(g_router_err), __VA_ARGS__); return (code); } while (0)
#define CUDA_FAIL() do { snprintf(g_router_err, sizeof(g_router_err), "%s", pfb_last_error()); return PF_ECUDA; } while (0)
#define CKB(x) do { if ((x) != 0) CUDA_FAIL(); } while (0)

struct SlotClass {
	int num_slots, label_log2, tree_cap, far_cap, sink_cap;
	uint64_t *hot; PfCold *cold; uint64_t *hot2; PfCold *cold2; int label2_log2; unsigned *epochs; PfTreeNode *tree; uint64_t *far; int *iscratch;
	int *work; int num_work; int *work_head;
};

struct pf_router {
	pf_config cfg;
	const pf_problem *prob;       /* caller-owned; must outlive the router */
	int N, E, T, n;
	int node_bits;                /* edge word / hot label layout (pf_layout.h) */
	PfNode *nodes; uint32_t *edges;
	PfSwitchDev *sw; PfIndexedDev *indexed;
	int *net_ptr, *net_term, *net_bb;
	float *crit, *net_delay;
	SlotClass small, big;
	PfTreeNode *pool[2]; int *pool_node[2]; PfNetLoc *loc; int cur;      /* pool[cur] is the live route-tree log */
	long long pool_cap; unsigned long long *pool_head;
	int *all_nets; int num_all; unsigned char *net_big; int *sel_counts; int *sel_scratch;
	short *ptc;                    /* rr_node[].ptc_num, only read when the result's serial number is assembled */
	std::vector<unsigned char> h_net_big;
	std::vector<int> net_rank;        /* position of a net in the fanout-sorted order */
	int iter_count;
	int best_overused, stall_count;   /* convergence watchdog, see pf_iteration_begin */
	std::vector<int> over_hist; int since_full;
	unsigned char *last_over; int cost_updates; int *committer;   /* per node: tag of the last cost update that found it overused */
	int cur_div, n_small, n_big; int *retry_work;
	char *ctl; unsigned long long h_pool_head;   /* device control block; host copy of the log head */
	int *status, *retry_list, *retry_count;
	PfStats *stats;
	int *d_overused; unsigned long long *d_wl;
	bool generated;               /* the rr graph was built on the device (pf_router_create_generated): no host node / edge arrays */
	int graph_ready;              /* 0 while a deferred graph has not been filled in */
	unsigned *events; long long event_cap; long long h_events;   /* multi-GPU only: this rank's occupancy event log */
	/* OPIN reservation */
	int num_groups; int *g_source, *g_count, *g_off, *g_chosen;
	long long avail_wl;
	int *gen_row;            /* device generator: edge-row offsets, alive until the fill pass has been joined (pfb_gen_fill_end) */
	double util;                  /* routed wirelength / available wirelength after the first iteration; < 0 = unknown */
	int div_explicit;             /* cfg.inflight_div was given by the caller */
	double t_mark[4];
	int64_t h2d_bytes, d2h_bytes;
	std::vector<int> work_small, work_big;
	std::vector<int> net_owner;       /* rank that routes each net (stripe sharding) */
	std::vector<unsigned char> net_cut;   /* 1 = the net's box reaches across a stripe cut (second route part) */
	std::vector<int> h_all;           /* host copy of all_nets: interior nets first, then cut nets, each in fanout order */
	int K1;                           /* number of interior nets at the head of all_nets */
	int n1_small, n1_big;             /* interior nets at the head of this iteration's two work lists */
	float win_abs_auto;
	int *vq[2]; int *queued; int vq_cap;   /* ripple re-routing: victim queues per slot class, per-net iteration tags */
	char *h_ctl;                          /* pinned host copy of the control block (fetch_ctl) */
	PfStats h_stats, h_stats_seen;        /* counters of the running iteration / what the step API has reported of them */
	bool sel_valid, sel_pending;          /* the next iteration's work lists and counts are already on the device / in h_ctl */
	std::vector<int> h_list_small, h_list_big;
	long long h_wl_used; unsigned xchg_seq;
	std::vector<float> crit_hist;        /* criticalities per iteration of the last pf_route_run with a host analysis */
	int comm_ready;                      /* pf_comm_init done */
	unsigned char *xreg; size_t xreg_bytes; unsigned char xhandle[64]; PfPeers peers;   /* region and mappings belong to the process-level transport cache (pf_router.cpp) */
	 unsigned char *term_owner; unsigned dseq;   /* exchange region (pf_layout.h) */
	bool owner_valid;                     /* committer[] reflects the route store (ripple re-routing) */
	bool force_all_once;                  /* the next iteration re-routes every net (polish pass) */
	bool iter_all;                        /* the running iteration re-routes every net */
};

extern "C" void pf_result_set_release_hook(int (*fn)(void *));   /* pf_file.c: who takes back result arrays that are not malloc memory */

static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
/* simple static-partition parallel loop for the host-side flattening of 10^7..10^8-element arrays */
template <class F> static inline void parallel_for(long long n, F f) {
	unsigned hw = std::thread::hardware_concurrency();
	int cap = 16;                                   /* memory-bound loops: more threads starve the DMA engine that drains the
	                                                 * staging buffer behind them (measured 8: 51, 16: 45, 32-64: 52 ms per
	                                                 * upload); PF_HOST_THREADS overrides */
	if (const char *e = getenv("PF_HOST_THREADS")) { int v = atoi(e); if (v > 0) cap = v; }
	int nt = (int)std::min<long long>(hw ? hw : 4, std::max<long long>(1, n / (1 << 16)));
	if (nt > cap) nt = cap;
	if (nt <= 1) { f(0, n); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < nt; t++) th.emplace_back([=]() { f(n * t / nt, n * (t + 1) / nt); });
	for (auto &t : th) t.join();
}

#endif /* PF_HOST_H */
