/*
 * pf_layout.h — device data layout of the B200 PathFinder router (plain structs shared by the
 * host driver pf_router.cpp, the CUDA kernels and the test emulator).  See DESIGN.md §3.
 */
#ifndef PF_LAYOUT_H
#define PF_LAYOUT_H

#include <stdint.h>

/* ------------------------------------------------------------------ device data layout */

/* rr node, 32 B = one HBM sector: everything one edge relaxation needs about its target.
 * Replaces the 104-byte AoS t_rr_node (vpr_types.h:946) + t_rr_node_route_inf (route_common_types.h:56). */
struct PfNode {
	short xlow, ylow, xhigh, yhigh;   /* 8  */
	float R, C;                       /* 8  */
	int occ;                          /* 4  mutable: atomics (rip-up / commit) */
	float acc_cost;                   /* 4  mutable: once per iteration */
	int edge_start;                   /* 4  row_ptr */
	unsigned short num_edges;         /* 2  */
	unsigned char type_ci;            /* 1  type | cost_index << 3 */
	unsigned char capacity;           /* 1  */
};

/* out-edge word: target node in the low `node_bits` bits, switch id above them.  node_bits is chosen per router
 * (PfParams.node_bits, 26 .. 30): 26 leaves 6 bits = 64 switch types and a 6-bit search tag in the hot label word;
 * graphs beyond 2^26 rr nodes (an 800 x 800 fabric has 7.3e7) take bits from both */
#define PF_MIN_NODE_BITS 26
#define PF_MAX_NODE_BITS 30
#define PF_MAX_SWITCHES 64
#define PF_MAX_INDEXED 32

struct PfSwitchDev { float R, Tdel; int buffered; };
struct PfIndexedDev { float base_cost, saved_base_cost, inv_length, T_linear, T_quadratic, C_load; int ortho; int pad; };

/* Search labels: an open-addressed hash table per warp, keyed by rr node, split in two.
 *   hot  (8 B/entry): total cost | search tag | node — everything a probe compares.  For the regular
 *        slots the hot table lives in SHARED memory (1024 entries = 8 KB per warp), so the probe that
 *        sits on the critical path of every edge relaxation costs a shared-memory access instead of an
 *        HBM round trip; big-net slots keep it in global memory.
 *   cold (32 B/entry, global, indexed by the same slot): what a settled label carries — backward cost,
 *        R_upstream, predecessor, entering switch and the node's CSR row.  Written once per accepted
 *        relaxation, read once per settled label; 32 KB per warp, so the whole set stays L2-resident.
 * The 6-bit search tag makes clearing free for 63 consecutive sink searches. */
#define PF_SMEM_HOT_LOG2 10
#define PF_SMEM_HOT_ENTRIES (1 << PF_SMEM_HOT_LOG2)
struct PfCold {
	float back; float R_up; int prev; int info;           /* prev >= 0: rr node; prev < 0: ~tree index (seed) */
	int edge_start; int pad0, pad1, pad2;                  /* info = entering switch | type << 8 | out-degree << 16;
	                                                          edge_start < 0: row unknown (seed), read the node record */
};

/* route-tree entry, 32 B.  Entries are appended in path order, so parent index < child index. */
struct PfTreeNode {
	int node;
	int parent;                    /* tree index, -1 at the root */
	float R_up, C_down, Tdel;
	short xlow, ylow, xhigh, yhigh;
	unsigned char sw;              /* switch parent→this */
	unsigned char type_ci;
	unsigned char flags;           /* bit0 re_expand, bit1 scratch mark */
	unsigned char pad;
};
#define PF_TF_REEXPAND 1
#define PF_TF_MARK 2

struct PfNetLoc { int off, count; };   /* a net's tree in the route store */

/* overridable for experiment builds (PF_EXTRA_NVCC_FLAGS, tools/ab_bench.sh): every pop scans the near set, every refill
 * fills it to PF_SH_REFILL — e.g. -DPF_SH_FRONTIER=64 -DPF_SH_REFILL=32 halves the scan and doubles the refills */
#ifndef PF_SH_FRONTIER
#define PF_SH_FRONTIER 128      /* near-set entries per warp in shared memory (a multiple of 32) */
#endif
#ifndef PF_SH_REFILL
#define PF_SH_REFILL 64         /* refill the near set to at most this many */
#endif
#define PF_MAX_BATCH 32         /* labels settled per step (one delta bucket) */

/* error/status bits written to PfParams.status[0] */
#define PF_ST_UNROUTABLE 1
#define PF_ST_POOL_OVERFLOW 2
#define PF_ST_INTERNAL 4
#define PF_ST_TWICE_TO_SINK_BF 8   /* breadth-first mode met a net with two pins on one SINK (route_breadth_first.c:208-256) */
#define PF_ST_BIG_OVERFLOW 16      /* a net outgrew the scratch of a big slot */
#define PF_ST_COMM_TIMEOUT 32      /* multi-GPU exchange: a peer never published its event log */
#define PF_ST_COMM_ABORT 64        /* multi-GPU exchange: a peer reported a failure */

struct PfStats { unsigned long long pops, pushes, visits, refills, nets, max_net_pops /* largest single net of the iteration */, stale; unsigned long long races; };

struct PfParams {
	PfNode *nodes;
	const uint32_t *edges;
	int num_nodes, nx, ny;
	int node_bits;         /* see PF_MIN_NODE_BITS */
	const PfSwitchDev *sw; int num_sw;
	const PfIndexedDev *indexed; int num_indexed;
	/* nets */
	const int *net_ptr; const int *net_term; const int *net_bb;   /* bb: xmin,xmax,ymin,ymax */
	const int *work; int num_work; int *work_head;
	const int *num_work_ptr;   /* non-NULL: the length of `work` is read from device memory (the retry launch: nets that
	                              overflowed a regular slot, counted by the launch before it) */
	const float *crit;     /* [num_terminals] timing criticality per terminal */
	float *net_delay;      /* [num_terminals] */
	/* options */
	float pres_fac, astar_fac, bend_cost, max_crit, crit_exp;
	float pop_slack;       /* delta-stepping bucket width in units of the cheapest edge cost for the
	                          sink's criticality: every label within it of the minimum is settled in one step */
	float win_rel, win_abs;/* near-set window: max(min*win_rel, win_abs) */
	int algorithm;         /* 0 = timing-driven (route_timing.c), 1 = breadth-first (route_breadth_first.c) */
	int max_batch;
	int skip_ripup;
	int validate;          /* > 0: a path whose commit finds a node full that its search saw free is taken back and the sink
	                          searched again, at most this many times per sink (optimistic concurrency control) */
	/* per-warp slot memory */
	uint64_t *hot;         /* global hot tables (NULL: the hot table lives in shared memory) */
	PfCold *cold; int label_log2;
	uint64_t *hot2; PfCold *cold2; int label2_log2;   /* per-slot fallback table in global memory: a sink search
	                                                      that outgrows the shared-memory table is restarted on it */
	unsigned *epochs;      /* [2 * slots]: search tags of the primary / fallback table */
	PfTreeNode *tree; int tree_cap;
	uint64_t *far; int far_cap;
	int lazy_seed_min;         /* big slots: trees of this many entries and more are seeded lazily (pf_search_sink); 0: never */
	int far_buckets;           /* 1: these slots keep the hybrid flat / bucketed far list (pf_device.cuh, frontier) */
	int *iscratch; int sink_cap;   /* per slot: 3 * (sink_cap+2) ints */
	/* route store: append-only log of route trees; loc[net] points at the net's current tree.
	 * A re-routed net appends its new tree and repoints loc; the log is compacted between
	 * iterations when it is more than half garbage. */
	PfTreeNode *pool; PfNetLoc *loc; unsigned long long *pool_head; long long pool_cap;
	int *pool_node;        /* [pool_cap] the rr node of every log entry again, 4 bytes each: what the per-iteration passes over ALL trees
	                          (congested-net selection, rip-up) read instead of the 32-byte entries */
	/* multi-GPU: every occupancy change this rank makes is also logged (node id, bit 31 = decrement) so that the
	 * other ranks can replay it; NULL on one GPU */
	unsigned *events; unsigned long long *event_head; long long event_cap;
	int *committer;        /* [num_nodes] net that committed this rr node last and still holds it, -1 none; may be NULL */
	/* ripple re-routing (NULL = off): a commit that knowingly shares a full rr node pushes the net holding it (committer[])
	 * onto the victim queue of that net's slot class, and the warps of this iteration's launches drain the queues after
	 * their own work list: the chain "A displaces B, B displaces C" is followed inside ONE PathFinder iteration, as in the
	 * serial reference, where every net is re-routed in every iteration (route_timing.c:161-183).  A net is queued at most
	 * once per iteration (queued[net] == iter_tag). */
	int *vq[2];            /* [vq_cap] victim queues of the regular / big slot class; -1 = empty entry */
	int *vq_ctl;           /* [0] head, [1] tail of vq[0]; [2] head, [3] tail of vq[1]; [4] warps routing a net right now */
	int vq_cap; int vq_class;   /* which queue this launch drains */
	int *queued; int iter_tag;
	unsigned char *net_big;    /* [num_nets] slot class of a net (written when a net moves to the big slots) */
	/* status */
	int *status;
	int *retry_list; int *retry_count;
	PfStats *stats;
};

/* ------------------------------------------------------------------ multi-GPU exchange region (one per rank)
 * A single allocation that the other ranks of the node map through CUDA IPC and read over NVLink / NVSwitch:
 *   PfXchgHeader | occupancy event log [2][event_cap] u32 | published sink delays [2][num_terminals] f32
 * Both payloads are double-buffered by the parity of their sequence number: a rank can only publish number s+2 after it
 * has consumed every peer's s+1, which the peers published after consuming its s — so buffer s & 1 is never rewritten
 * while somebody still reads it, without any acknowledgement traffic. */
#define PF_XCHG_MAX_RANKS 8
#define PF_XCHG_HEADER_BYTES 128
struct PfXchgHeader {
	unsigned seq[2];               /* seq[b] = number of the last exchange published in log buffer b (release-stored last) */
	unsigned dseq[2];              /* same for the delay buffers */
	unsigned abort_flag;           /* non-zero: this rank failed; peers stop waiting */
	unsigned pad0[3];
	unsigned long long count[2];   /* events in log buffer b */
	unsigned long long pad1[10];
};
struct PfPeers { unsigned char *base[PF_XCHG_MAX_RANKS]; };

/* device image of the timing graph for the static timing analysis (pf_sta_device.cuh, pf_sta.cpp) */
struct PfStaDev {
	int num_tnodes, num_terminals, num_levels;   /* tnodes are renumbered in level order: level l = [level_ptr[l], level_ptr[l+1]) */
	const int *edge_ptr, *edge_to;           /* out-edges (reference order within a tnode) */
	float *Tdel;                             /* [num_tedges] working copy: static delays + this call's net delays */
	const int *in_ptr, *in_rec;              /* in-edges: pairs (source tnode, index of the edge in the out-edge arrays) */
	const unsigned char *type;
	const int *clock_domain;
	const float *clock_delay;
	const int *level_ptr, *level_nodes;
	const int *term_edge;                    /* [num_terminals] out-edge index the terminal's delay belongs to, -1 none */
	const int *term_driver;                  /* [num_terminals] driver tnode of the terminal's net, -1 none */
	float *T_arr, *T_req;
	float *stat;                             /* per domain pair [3]: max_Tarr, cpd, least_slack */
	/* clock-to-flipflop override constraints (pf_timing_graph, pf_types.h), sorted by (renumbered tnode, source domain) */
	int num_overrides, src_domain;           /* src_domain: the source clock domain of the traversal in flight (set per domain pair) */
	const int *ovr_tnode, *ovr_domain;
	const float *ovr_constraint;
	/* the analysis of the finished routing (do_timing_analysis(..., is_final_analysis = TRUE), base/stats.c:155-164) */
	int final_analysis;                      /* required times at the sinks are the real ones, not relaxed to max_Tarr (path_delay.c:2786-2790) */
	float *slack;                            /* [num_terminals] or NULL: update_slacks keeps the least slack of every net pin (:3117-3125) */
};

#endif /* PF_LAYOUT_H */
