/*
 * pf_router.cpp — host side of the B200 PathFinder router behind the C-ABI of pf_router.h.
 * Owns device memory, flattens the pf_problem into the device layout (32-byte node records,
 * packed edge words), runs the PathFinder outer loop (reference route_timing.c:85-343) and turns
 * the device route store back into s_trace-ordered lists.  All device work goes through
 * pf_backend.h; this file contains no routing arithmetic.
 */
#include "pf_host.h"
#include "../../include/pf_gen.h"
#ifndef PF_DEV
#define PF_DEV static inline
#endif
#include "pf_gen_device.cuh"
extern "C" int pf_gen_dev_params(const pf_gen_params *gp, PfGenDev *G, short *cb_inv);

/* set by pf_router_create_generated for the pf_router_create call it makes: build the rr graph on the device */
static thread_local const pf_gen_params *g_generate = NULL;

char g_router_err[512] = "";

extern "C" const char *pf_last_error(void) { return g_router_err; }
extern "C" const char *pf_backend_name(void) { return pfb_name(); }
extern "C" int pf_device_count(void) { return pfb_device_count(); }

extern "C" void pf_config_default(pf_config *c) {
	memset(c, 0, sizeof(*c));
	c->nranks = 1;
	c->pop_slack = -1.f;
}

/* pf_problem_check (pf_file.c) is a single-threaded scan; on 10^8 edges that is a visible part of the call.
 * The range checks of the big arrays run inside the parallel flattening passes instead, and the serial
 * checker is only consulted for the error message. */
static bool problem_header_ok(const pf_problem *p) {
	if (p->nx <= 0 || p->ny <= 0 || p->num_nodes <= 0 || p->num_indexed < PF_CHANX_COST_INDEX_START) return false;
	if (!p->row_ptr || p->row_ptr[0] != 0 || p->row_ptr[p->num_nodes] != p->num_edges) return false;
	return true;
}

/* nets / tables part of pf_problem_check (small arrays).  The shape (monotonic net_ptr, boxes, small tables) is
 * checked before anything indexes with it; the terminal lookups — one random read of type[] per pin, 4 ms at
 * 800 k pins — run on a helper thread next to the graph upload (problem_terminals_ok). */
static bool problem_terminals_ok(const pf_problem *p) {
	for (int i = 0; i < p->num_nets; i++) {
		if (p->net_is_global[i]) continue;
		const int b = p->net_ptr[i], e = p->net_ptr[i + 1];
		for (int k = b; k < e; k++) {
			int n = p->net_terminals[k];
			if (n < 0 || n >= p->num_nodes || p->type[n] != (k == b ? PF_SOURCE : PF_SINK)) return false;
		}
	}
	return true;
}

static bool problem_nets_ok(const pf_problem *p) {
	if (p->net_ptr[0] != 0 || p->net_ptr[p->num_nets] != p->num_terminals) return false;
	for (int i = 0; i < p->num_nets; i++) {
		int b = p->net_ptr[i], e = p->net_ptr[i + 1];
		if (e <= b) return false;
		if (p->net_bb[4 * i] > p->net_bb[4 * i + 1] || p->net_bb[4 * i + 2] > p->net_bb[4 * i + 3]) return false;
	}
	for (int i = PF_CHANX_COST_INDEX_START; i < p->num_indexed; i++)
		if (p->indexed[i].ortho_cost_index < 0 || p->indexed[i].ortho_cost_index >= p->num_indexed) return false;
	for (int i = 0; i < p->num_opin_groups; i++) {
		int s = p->opin_group_source[i];
		if (s < 0 || s >= p->num_nodes || p->type[s] != PF_SOURCE || p->opin_group_count[i] < 0
				|| p->opin_group_count[i] > p->row_ptr[s + 1] - p->row_ptr[s]) return false;
	}
	return true;
}

/* ------------------------------------------------------------------ the device control block
 * One 512-byte block holds every small counter the host reads, so that a PathFinder iteration costs ONE read
 * (fetch_ctl) however many kernels it launched:
 *   iteration scope (zeroed by pf_iteration_begin)
 *     [0]   status[8]        error bits, failed-net count, offending net
 *     [32]  PfStats          pops / pushes / visits / nets ... accumulated over the iteration's launches
 *   written once per iteration by their kernels
 *     [96]  sel_counts[4]    lengths of the two work lists of the NEXT iteration (pf_select_*), interior heads
 *     [112] overused[4]      feasible_routing's count
 *     [128] wl[2]            [0] wirelength of this rank's trees, [1] wirelength in use on the fabric (from occupancy)
 *   part scope (zeroed before every route part)
 *     [160] retry_count[4]   nets that outgrew a regular slot in this part
 *     [176] heads of the regular / big / retry work queues (4 ints each)
 *     [224] victim-queue heads, tails, routing-warp count (8 ints)
 *     [256] event_head       multi-GPU: length of this part's occupancy event log
 *   persistent
 *     [272] pool_head[2]     route-store log head */
enum { CTL_STATUS = 0, CTL_STATS = 32, CTL_SEL = 96, CTL_OVER = 112, CTL_WL = 128, CTL_PART = 160, CTL_RETRY = 160, CTL_HEAD_SMALL = 176,
	CTL_HEAD_BIG = 192, CTL_HEAD_RETRY = 208, CTL_VQ = 224, CTL_EVENTS = 256, CTL_PART_END = 272, CTL_POOL = 272, CTL_BYTES = 512 };

static void bind_ctl(pf_router *r) {
	r->status = (int *)(r->ctl + CTL_STATUS);
	r->stats = (PfStats *)(r->ctl + CTL_STATS);
	r->sel_counts = (int *)(r->ctl + CTL_SEL);
	r->d_overused = (int *)(r->ctl + CTL_OVER);
	r->d_wl = (unsigned long long *)(r->ctl + CTL_WL);
	r->retry_count = (int *)(r->ctl + CTL_RETRY);
	r->small.work_head = (int *)(r->ctl + CTL_HEAD_SMALL); r->big.work_head = (int *)(r->ctl + CTL_HEAD_BIG);
	r->pool_head = (unsigned long long *)(r->ctl + CTL_POOL);
}

static int ceil_log2(long long v) { int l = 0; while ((1ll << l) < v) l++; return l; }

static void free_slot_class(SlotClass &s) {
	pfb_free(s.hot); pfb_free(s.cold); pfb_free(s.hot2); pfb_free(s.cold2); pfb_free(s.epochs); pfb_free(s.tree); pfb_free(s.far); pfb_free(s.iscratch);
	pfb_free(s.work); pfb_free(s.work_head);
	memset(&s, 0, sizeof(s));
}

static int alloc_slot_class(SlotClass &s, int max_work, bool hot_in_smem) {
	size_t cap = (size_t)1 << s.label_log2;
	s.hot = hot_in_smem ? NULL : (uint64_t *)pfb_alloc(sizeof(uint64_t) * cap * s.num_slots);
	s.cold = (PfCold *)pfb_alloc_raw(sizeof(PfCold) * cap * s.num_slots);
	s.hot2 = NULL; s.cold2 = NULL;
	if (hot_in_smem && s.label2_log2 > 0) {
		size_t cap2 = (size_t)1 << s.label2_log2;
		s.hot2 = (uint64_t *)pfb_alloc(sizeof(uint64_t) * cap2 * s.num_slots);
		s.cold2 = (PfCold *)pfb_alloc_raw(sizeof(PfCold) * cap2 * s.num_slots);
		if (!s.hot2 || !s.cold2) return -1;
	}
	s.epochs = (unsigned *)pfb_alloc(sizeof(unsigned) * 2 * s.num_slots);
	s.tree = (PfTreeNode *)pfb_alloc_raw(sizeof(PfTreeNode) * (size_t)s.tree_cap * s.num_slots);
	s.far = (uint64_t *)pfb_alloc_raw(sizeof(uint64_t) * (size_t)s.far_cap * s.num_slots);
	s.iscratch = (int *)pfb_alloc_raw(sizeof(int) * ((size_t)3 * (s.sink_cap + 2) + (size_t)2 * s.tree_cap) * s.num_slots);
	s.work = (int *)pfb_alloc_raw(sizeof(int) * (size_t)(max_work > 0 ? max_work : 1));
	s.work_head = (int *)pfb_alloc(sizeof(int) * 4);
	if ((!hot_in_smem && !s.hot) || !s.cold || !s.epochs || !s.tree || !s.far || !s.iscratch || !s.work || !s.work_head) return -1;
	return 0;
}

/* ---- process-level cache of the multi-GPU exchange region and of the peers' mappings.
 * The region is sized for the worst case of the route store (cfg 4: 426 MB), and cudaMalloc + cudaIpcGetMemHandle of it plus
 * one cudaIpcOpenMemHandle per peer cost ~65 ms per router at N = 4 — three times the rest of pf_router_create.  Like the
 * reference's MPI communicator (created once per process, parallel_route/spatial.cxx), the transport therefore outlives the
 * router: the region of a destroyed router is kept and handed to the next one that fits, peers' regions stay mapped (looked
 * up by their 64-byte IPC handle), and the exchange sequence numbers simply continue, so no header is ever reset while a
 * slow peer may still be polling it.  A region that has become too small is retired, not freed: a peer may still have it
 * mapped, and freeing exported memory under an importer is undefined.  pf_comm_release_cache() drops everything that no
 * live router uses (call it collectively, after the last router of the job). */
struct CommPeerMap { unsigned char handle[64]; void *ptr; };
struct CommRegion { unsigned char *reg; size_t bytes; unsigned char handle[64]; bool in_use; unsigned seq, dseq; /* last numbers published from it */ };
static struct CommCache {
	std::vector<CommRegion> regions;          /* one per multi-rank router alive at the same time in this process (usually one) */
	std::vector<CommPeerMap> peers;
	std::vector<void *> retired;
	long long n_alloc, n_reuse, n_open, n_open_reuse;     /* pf_debug_comm_cache */
} g_comm;
static std::mutex g_comm_mu;

/* the exchange region of a new router: a cached one that is free and large enough, else a new one */
static unsigned char *comm_region_acquire(size_t bytes, unsigned char handle[64], unsigned *seq, unsigned *dseq) {
	std::lock_guard<std::mutex> lk(g_comm_mu);
	CommRegion *pick = NULL, *small = NULL;
	for (CommRegion &c : g_comm.regions) {
		if (c.in_use) continue;
		if (c.bytes >= bytes) { if (!pick || c.bytes < pick->bytes) pick = &c; }
		else small = &c;
	}
	if (pick) {
		if (pfb_ipc_clear_abort(pick->reg) != 0) return NULL;
		g_comm.n_reuse++;
	} else {
		CommRegion c;
		memset(&c, 0, sizeof(c));
		const size_t want = bytes + bytes / 4;                      /* headroom: the next problem may be a little larger */
		c.reg = (unsigned char *)pfb_ipc_alloc(want, c.handle); c.bytes = want;
		if (!c.reg) { c.reg = (unsigned char *)pfb_ipc_alloc(bytes, c.handle); c.bytes = bytes; }
		if (!c.reg) return NULL;
		g_comm.n_alloc++;
		if (small) { g_comm.retired.push_back(small->reg); c.seq = small->seq; c.dseq = small->dseq; *small = c; pick = small; }   /* takes the place of the one that is too small */
		else { g_comm.regions.push_back(c); pick = &g_comm.regions.back(); }
	}
	pick->in_use = true;
	memcpy(handle, pick->handle, 64);
	*seq = pick->seq; *dseq = pick->dseq;
	return pick->reg;
}
static void comm_region_release(pf_router *r) {
	if (!r->xreg) return;
	std::lock_guard<std::mutex> lk(g_comm_mu);
	for (CommRegion &c : g_comm.regions) if (c.reg == r->xreg) { c.seq = r->xchg_seq; c.dseq = r->dseq; c.in_use = false; }
}
static void *comm_peer_open(const unsigned char handle[64]) {
	std::lock_guard<std::mutex> lk(g_comm_mu);
	for (const CommPeerMap &m : g_comm.peers) if (memcmp(m.handle, handle, 64) == 0) { g_comm.n_open_reuse++; return m.ptr; }
	void *p = pfb_ipc_open(handle);
	g_comm.n_open++;
	if (p) { CommPeerMap m; memcpy(m.handle, handle, 64); m.ptr = p; g_comm.peers.push_back(m); }
	return p;
}
/* diagnostics: { bytes of the cached regions, regions allocated, routers that re-used one, peer regions mapped, mappings re-used } */
extern "C" void pf_debug_comm_cache(int64_t out[5]) {
	std::lock_guard<std::mutex> lk(g_comm_mu);
	out[0] = 0;
	for (const CommRegion &c : g_comm.regions) out[0] += (int64_t)c.bytes;
	out[1] = g_comm.n_alloc; out[2] = g_comm.n_reuse; out[3] = g_comm.n_open; out[4] = g_comm.n_open_reuse;
}
extern "C" int pf_comm_release_cache(void) {
	std::lock_guard<std::mutex> lk(g_comm_mu);
	for (const CommRegion &c : g_comm.regions) if (c.in_use) FAILF(PF_EINVAL, "a router still uses an exchange region");
	pfb_sync();
	for (const CommPeerMap &m : g_comm.peers) pfb_ipc_close(m.ptr);
	g_comm.peers.clear();
	for (void *p : g_comm.retired) pfb_ipc_free(p);
	g_comm.retired.clear();
	for (const CommRegion &c : g_comm.regions) pfb_ipc_free(c.reg);
	g_comm.regions.clear();
	return PF_OK;
}

extern "C" void pf_router_destroy(pf_router *r) {
	if (!r) return;
	pfb_gen_fill_end(NULL);          /* a create that failed half way: the generator's side stream still writes the graph arrays */
	pfb_sync();
	pfb_free(r->gen_row);
	pfb_free(r->nodes); pfb_free(r->edges); pfb_free(r->sw); pfb_free(r->indexed);
	pfb_free(r->net_ptr); pfb_free(r->net_term); pfb_free(r->net_bb);
	pfb_free(r->crit); pfb_free(r->net_delay);
	if (r->ctl) { r->small.work_head = NULL; r->big.work_head = NULL; }
	free_slot_class(r->small); free_slot_class(r->big);
	pfb_free(r->pool[0]); pfb_free(r->pool[1]); pfb_free(r->pool_node[0]); pfb_free(r->pool_node[1]); pfb_free(r->loc);
	pfb_free(r->all_nets); pfb_free(r->net_big); pfb_free(r->retry_work); pfb_free(r->sel_scratch); pfb_free(r->ptc);
	pfb_free(r->ctl); pfb_host_free(r->h_ctl); pfb_free(r->retry_list); pfb_free(r->last_over); pfb_free(r->committer);
	comm_region_release(r);
	pfb_free(r->term_owner); pfb_free(r->vq[0]); pfb_free(r->vq[1]); pfb_free(r->queued);
	pfb_free(r->g_source); pfb_free(r->g_count); pfb_free(r->g_off); pfb_free(r->g_chosen);
	delete r;
}

/* rr node records → HBM.  With a pinned staging buffer the host threads flatten 4 MB pieces and each issues
 * the async copy of its own piece as soon as it is written, so packing and PCIe overlap; `bad` collects the
 * range checks of pf_problem_check for the node arrays (done here so the arrays are read once). */
#define PF_UPLOAD_PIECE (4u << 20)
/* stores to write-combined memory sit in the core's WC buffers until flushed: drain them before the DMA reads */
static inline void wc_fence() {
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_sfence();
#else
	std::atomic_thread_fence(std::memory_order_seq_cst);
#endif
}
static int upload_nodes(pf_router *r, void *staging, int *bad_out, long long *avail_wl, short *stage_ptc) {
	const pf_problem *p = r->prob;
	PfNode *stage = staging ? (PfNode *)staging : NULL;
	std::vector<PfNode> hv;
	if (!stage) { hv.resize((size_t)r->N); stage = hv.data(); }
	PfNode *h = stage;
	std::atomic<int> bad(0), fail(0);
	std::atomic<long long> wl(0);
	const long long piece = PF_UPLOAD_PIECE / sizeof(PfNode);
	parallel_for(r->N, [&](long long lo, long long hi) {
		if (staging) pfb_bind_thread();
		int b = 0;
		long long w = 0;
		for (long long c0 = lo; c0 < hi; c0 += piece) {
			const long long c1 = std::min(hi, c0 + piece);
			for (long long i = c0; i < c1; i++) {
				PfNode d;                                    /* built in registers, stored whole (the staging buffer is write-combined) */
				const int ne = p->row_ptr[i + 1] - p->row_ptr[i];
				b |= (ne < 0) | (ne > 32767) | (p->type[i] > PF_CHANY) | (p->cost_index[i] < 0) | (p->cost_index[i] >= p->num_indexed)
					| (p->xlow[i] > p->xhigh[i]) | (p->ylow[i] > p->yhigh[i]) | (p->xlow[i] < 0) | (p->ylow[i] < 0)
					| (p->xhigh[i] > p->nx + 1) | (p->yhigh[i] > p->ny + 1) | (p->capacity[i] < 0) | (p->capacity[i] > 255) | (p->cost_index[i] > 31);
				d.xlow = p->xlow[i]; d.ylow = p->ylow[i]; d.xhigh = p->xhigh[i]; d.yhigh = p->yhigh[i];
				d.R = p->R[i]; d.C = p->C[i];
				d.occ = 0; d.acc_cost = 1.f;                 /* alloc_and_load_rr_node_route_structs, route_common.c:1012-1034 */
				d.edge_start = p->row_ptr[i];
				d.num_edges = (unsigned short)ne;
				d.type_ci = (unsigned char)(p->type[i] | (p->cost_index[i] << 3));
				d.capacity = (unsigned char)p->capacity[i];
				h[i] = d;
				if (p->type[i] == PF_CHANX || p->type[i] == PF_CHANY) w += 1 + p->xhigh[i] - p->xlow[i] + p->yhigh[i] - p->ylow[i];
			}
			if (stage_ptc) memcpy(stage_ptc + c0, p->ptc_num + c0, sizeof(short) * (size_t)(c1 - c0));
			if (staging) wc_fence();
			if (staging && pfb_h2d_async(r->nodes + c0, h + c0, sizeof(PfNode) * (size_t)(c1 - c0)) != 0) fail = 1;
			if (stage_ptc) {
				if (pfb_h2d_async(r->ptc + c0, stage_ptc + c0, sizeof(short) * (size_t)(c1 - c0)) != 0) fail = 1;
			}
		}
		if (b) bad = 1;
		wl += w;
	});
	if (fail) return PF_ECUDA;
	if (!staging) CKB(pfb_h2d(r->nodes, h, sizeof(PfNode) * (size_t)r->N));
	if (avail_wl && !stage_ptc) CKB(pfb_h2d(r->ptc, p->ptc_num, sizeof(short) * (size_t)r->N));   /* first upload, no pinned staging */
	r->h2d_bytes += (int64_t)sizeof(PfNode) * r->N + (avail_wl ? (int64_t)sizeof(short) * r->N : 0);
	if (bad_out) *bad_out = bad;
	if (avail_wl) *avail_wl = wl;
	return PF_OK;
}

/* packed edge words → HBM, same scheme */
static int upload_edges(pf_router *r, uint32_t *staging, int *bad_out) {
	const pf_problem *p = r->prob;
	std::vector<uint32_t> ewv;
	uint32_t *ew = staging;
	if (!ew) { ewv.resize((size_t)std::max(r->E, 1)); ew = ewv.data(); }
	std::atomic<int> bad(0), fail(0);
	const long long piece = PF_UPLOAD_PIECE / sizeof(uint32_t);
	const unsigned N = (unsigned)p->num_nodes, S = (unsigned)p->num_switches;
	const int nb = r->node_bits;
	parallel_for(r->E, [&](long long lo, long long hi) {
		if (staging) pfb_bind_thread();
		int b = 0;
		for (long long c0 = lo; c0 < hi; c0 += piece) {
			const long long c1 = std::min(hi, c0 + piece);
			for (long long k = c0; k < c1; k++) {
				const unsigned to = (unsigned)p->edge_to[k], s = (unsigned)(int)p->edge_sw[k];
				b |= (to >= N) | (s >= S);
				ew[k] = to | (s << nb);
			}
			if (staging) wc_fence();
			if (staging && pfb_h2d_async(r->edges + c0, ew + c0, sizeof(uint32_t) * (size_t)(c1 - c0)) != 0) fail = 1;
		}
		if (b) bad = 1;
	});
	if (fail) return PF_ECUDA;
	if (!staging) CKB(pfb_h2d(r->edges, ew, sizeof(uint32_t) * (size_t)r->E));
	r->h2d_bytes += (int64_t)sizeof(uint32_t) * r->E;
	if (bad_out) *bad_out = bad;
	return PF_OK;
}

extern "C" int pf_router_create(const pf_problem *p, const pf_config *cfg_in, pf_router **out) {
	char msg[256];
	*out = NULL;
	if (!p || !cfg_in) FAILF(PF_EINVAL, "null argument");
	double t_a = now_s();
	/* small arrays now; the 10^7..10^8-element node and edge arrays are range-checked by the passes that
	 * flatten them (upload_nodes / upload_edges), so they are read once */
	const pf_gen_params *gen = g_generate;
	g_generate = NULL;
	PfGenDev Gd;
	std::vector<short> gen_inv(PF_GEN_MAX_W);
	if (gen) {
		/* the rr graph is a closed-form function of the generator parameters (pf_gen_device.cuh): no host arrays */
		if (pf_gen_dev_params(gen, &Gd, gen_inv.data()) != PF_OK) FAILF(PF_EINVAL, "invalid generator parameters");
		if (p->nx != Gd.nx || p->ny != Gd.ny || p->num_nodes != Gd.num_nodes || p->num_opin_groups != 0 || p->num_switches != 3 || p->num_indexed != 6)
			FAILF(PF_EINVAL, "the nets-only problem does not belong to these generator parameters (use pf_gen_grid_nets)");
		if (p->num_nets < 0 || !p->net_ptr || !problem_nets_ok(p)) FAILF(PF_EINVAL, "invalid problem (nets)");
	} else if (!problem_header_ok(p) || !problem_nets_ok(p)) {
		if (pf_problem_check(p, msg, sizeof(msg)) != PF_OK) FAILF(PF_EINVAL, "invalid problem: %s", msg);
		FAILF(PF_EINVAL, "invalid problem");
	}
	double t_b = now_s();
	std::atomic<int> terminals_bad(0);
	const PfGenDev *Gp = gen ? &Gd : NULL;
	std::thread terminal_check([p, Gp, &terminals_bad]() {
		if (!Gp) { if (!problem_terminals_ok(p)) terminals_bad = 1; return; }
		parallel_for(p->num_nets, [&](long long lo, long long hi) {      /* terminal types from the closed form */
			for (long long i = lo; i < hi; i++) {
				if (p->net_is_global[i]) continue;
				for (int k = p->net_ptr[i]; k < p->net_ptr[i + 1]; k++) {
					const int n = p->net_terminals[k];
					if (n < 0 || n >= Gp->num_nodes || pf_gen_decode(*Gp, n).type != (k == p->net_ptr[i] ? PF_SOURCE : PF_SINK)) { terminals_bad = 1; return; }
				}
			}
		});
	});
	struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{terminal_check};   /* every return path */
	/* edge word = target node | switch << node_bits; the bits above the node id are also the search tag of the hot labels */
	const int node_bits = gen ? Gd.node_bits : std::max(PF_MIN_NODE_BITS, ceil_log2((long long)p->num_nodes));
	if (node_bits > PF_MAX_NODE_BITS) FAILF(PF_EINVAL, "num_rr_nodes %d exceeds the %d-bit node field", p->num_nodes, PF_MAX_NODE_BITS);
	if (p->num_switches > PF_MAX_SWITCHES || p->num_switches > (1 << (32 - node_bits)))
		FAILF(PF_EINVAL, "%d switch types (max %d with %d rr nodes)", p->num_switches, std::min(PF_MAX_SWITCHES, 1 << (32 - node_bits)), p->num_nodes);
	if (p->num_indexed > PF_MAX_INDEXED) FAILF(PF_EINVAL, "%d rr_indexed_data rows (max %d)", p->num_indexed, PF_MAX_INDEXED);
	if (cfg_in->nranks < 1 || cfg_in->rank < 0 || cfg_in->rank >= cfg_in->nranks) FAILF(PF_EINVAL, "bad rank %d / nranks %d", cfg_in->rank, cfg_in->nranks);
	if (pfb_init(cfg_in->device) != 0) CUDA_FAIL();

	pf_router *r = new pf_router();
	memset((void *)&r->cfg, 0, sizeof(pf_config));
	r->cfg = *cfg_in;
	pf_config &c = r->cfg;
	r->node_bits = node_bits;
	r->prob = p; r->N = p->num_nodes; r->E = p->num_edges; r->generated = gen != NULL; r->T = p->num_terminals; r->n = p->num_nets;
	r->gen_row = NULL; r->nodes = NULL; r->edges = NULL; r->sw = NULL; r->indexed = NULL; r->net_ptr = r->net_term = r->net_bb = NULL;
	r->crit = r->net_delay = NULL; memset(&r->small, 0, sizeof(SlotClass)); memset(&r->big, 0, sizeof(SlotClass));
	r->pool[0] = r->pool[1] = NULL; r->pool_node[0] = r->pool_node[1] = NULL; r->loc = NULL; r->cur = 0; r->pool_head = NULL;
	r->all_nets = NULL; r->num_all = 0; r->net_big = NULL; r->sel_counts = NULL; r->sel_scratch = NULL; r->ptc = NULL; r->K1 = 0; r->n1_small = r->n1_big = 0; r->iter_count = 0;
	r->best_overused = 0x7fffffff; r->stall_count = 0; r->since_full = 0; r->last_over = NULL; r->cost_updates = 0; r->committer = NULL; r->cur_div = 32; r->n_small = r->n_big = 0; r->retry_work = NULL; r->ctl = NULL; r->h_pool_head = 0;
	r->status = r->retry_list = r->retry_count = NULL; r->stats = NULL; r->d_overused = NULL; r->d_wl = NULL;
	r->events = NULL; r->event_cap = 0; r->h_events = 0; r->graph_ready = 1; r->num_groups = 0; r->g_source = r->g_count = r->g_off = r->g_chosen = NULL;
	r->h2d_bytes = r->d2h_bytes = 0; r->vq[0] = r->vq[1] = NULL; r->queued = NULL; r->vq_cap = 0; r->iter_all = true; r->force_all_once = false; r->owner_valid = false; r->h_ctl = NULL; r->xreg = NULL; r->xreg_bytes = 0; r->term_owner = NULL; r->comm_ready = 0; r->dseq = 0; memset(&r->peers, 0, sizeof(r->peers)); memset(r->xhandle, 0, sizeof(r->xhandle)); r->sel_valid = r->sel_pending = false; r->xchg_seq = 0; r->h_wl_used = 0;
	memset(&r->h_stats, 0, sizeof(PfStats)); memset(&r->h_stats_seen, 0, sizeof(PfStats));

	int sms = pfb_num_sms();
	if (c.warps_per_block <= 0) c.warps_per_block = 4;
	if (c.num_slots <= 0) c.num_slots = (sms > 0 ? sms : 148) * 20;   /* 5 CTAs x 4 warps resident per SM */
	c.label_log2 = PF_SMEM_HOT_LOG2;                 /* regular slots: hot label table in shared memory */
	const bool bf = p->opts.router_algorithm == 1;   /* the breadth-first wave floods the net's bounding box: larger per-slot scratch */
	if (c.label2_log2 == 0) c.label2_log2 = bf ? 15 : 13;   /* per-slot fallback table in global memory; < 0: none */
	if (c.label2_log2 < 0) c.label2_log2 = 0;
	if (c.tree_cap <= 0) c.tree_cap = 2048;
	if (c.far_cap <= 0) c.far_cap = bf ? 32768 : 8192;      /* regular slots: one flat far list (pf_device.cuh, frontier) */
	if (c.sink_cap <= 0) c.sink_cap = 64;
	const bool auto_big_slots = c.big_slots <= 0;
	if (auto_big_slots) c.big_slots = 64;
	if (c.reroute_all_iters == 0) c.reroute_all_iters = 1;
	r->div_explicit = c.inflight_div > 0; r->util = -1.;
	if (c.inflight_div <= 0) c.inflight_div = 16;
	/* nets in flight: at least one per 20 x 20 tiles of fabric (the few hundred nets of a late iteration then go
	 * out together instead of sixteen rounds deep), and never fewer than one */
	if (c.min_slots <= 0) c.min_slots = std::max(1, (int)((long long)p->nx * p->ny / 400));
	if (c.stall_iters == 0) c.stall_iters = 3;
	if (c.history_window < 0) c.history_window = 0;   /* 0 = off (default) */
	if (c.keep_newcomer < 0) c.keep_newcomer = 0;     /* 0 = off (default): measured on B200 it costs ~8 % wirelength and does
	                                                    not shorten the tight-W tail */
	if (c.ripple == 0) c.ripple = 1;
	if (c.ripple < 0) c.ripple = 0;
	if (c.lazy_seed_min == 0) c.lazy_seed_min = 256;
	if (c.lazy_seed_min < 0) c.lazy_seed_min = 0;
	if (c.validate_commits == 0) c.validate_commits = 2;
	if (c.validate_commits < 0) c.validate_commits = 0;
	if (c.max_batch > PF_MAX_BATCH) c.max_batch = PF_MAX_BATCH;
	if (c.max_batch < 0) c.max_batch = 0;

	/* work lists: routed nets in decreasing-fanout order (route_timing.c:98-106), sharded by rank */
	std::vector<int> order;
	int max_sinks = 1;
	for (int i = 0; i < r->n; i++) {
		int ns = p->net_ptr[i + 1] - p->net_ptr[i] - 1;
		if (p->net_is_global[i] || ns < 1) continue;     /* SURVEY §8b edge case (i) */
		order.push_back(i);
		max_sinks = std::max(max_sinks, ns);
	}
	{	/* decreasing fanout, ties in netlist order: a stable counting sort (200 k nets of equal fanout need no comparison sort) */
		std::vector<int> cnt((size_t)max_sinks + 2, 0), sorted(order.size());
		for (int i : order) cnt[(size_t)(max_sinks - (p->net_ptr[i + 1] - p->net_ptr[i] - 1)) + 1]++;
		for (size_t k = 1; k < cnt.size(); k++) cnt[k] += cnt[k - 1];
		for (int i : order) sorted[(size_t)cnt[(size_t)(max_sinks - (p->net_ptr[i + 1] - p->net_ptr[i] - 1))]++] = i;
		order.swap(sorted);
	}
	r->net_rank.assign((size_t)std::max(r->n, 1), 0);
	for (size_t k = 0; k < order.size(); k++) r->net_rank[order[k]] = (int)k;
	/* Sharding over ranks.  The reference's MPI router deals nets round-robin / LPT over ranks
	 * (mpi_route_load_balanced…encoded.cxx:74-170) and pays for it with per-sink messages.  Here a rank only
	 * learns other ranks' routes at the occupancy sync, so nets are sharded SPATIALLY: sort by bounding-box
	 * centre along x and cut the grid into nranks stripes of equal total fanout.
	 *   interior nets  — bounding box inside one stripe, at least one maximum wire length away from the next
	 *                    stripe's interior nets: a search never leaves its bounding box (route_timing.c:622), so
	 *                    interior nets of different stripes cannot touch the same rr node.  They are routed
	 *                    first, all ranks at once, and nobody works on a stale view of anybody.
	 *   cut nets       — everything else (boxes that reach across a cut).  After the first occupancy sync the
	 *                    rank that owns the cut routes them with every interior route visible.
	 * The iteration therefore behaves like a one-GPU iteration with a particular net order. */
	std::vector<int> owner((size_t)std::max(r->n, 1), 0);
	std::vector<unsigned char> cut_net((size_t)std::max(r->n, 1), 0);
	/* the partition is host work over every net (5-8 ms per rank on cfg 4) that nothing on the device waits for: it runs on a helper
	 * thread under the device allocations and the generator passes below and is joined before the work lists are built.  It reads
	 * the problem and `order`, writes owner[] / cut_net[], and touches nothing of the router (which a failing create destroys) */
	const int part_nranks = c.nranks, part_verbose = c.verbose;
	auto partition_nets = [&, part_nranks, part_verbose]() {
		/* stable counting sort by xmin + xmax (0 .. 2 nx + 4) */
		std::vector<int> byx(order.size());
		{
			const int nkeys = 2 * (p->nx + 2) + 2;
			std::vector<int> start((size_t)nkeys + 1, 0);
			auto key = [&](int i) { return std::min(nkeys - 1, std::max(0, p->net_bb[4 * i] + p->net_bb[4 * i + 1])); };
			for (int i : order) start[(size_t)key(i) + 1]++;
			for (int k = 0; k < nkeys; k++) start[(size_t)k + 1] += start[(size_t)k];
			for (int i : order) byx[(size_t)start[(size_t)key(i)]++] = i;
		}
		long long total_f = 0;
		for (int i : byx) total_f += p->net_ptr[i + 1] - p->net_ptr[i];
		int lmax = 1;                                /* longest wire, in tiles: 1 / inv_length of the CHAN cost indices */
		for (int i = PF_CHANX_COST_INDEX_START; i < p->num_indexed; i++)
			if (p->indexed[i].inv_length > 0.f) lmax = std::max(lmax, (int)(1.f / p->indexed[i].inv_length + 0.5f));
		/* share[k]: the part of the total fanout (in the order of the box centres) that forms stripe k.  Equal shares do not give
		 * equal work: the nets across cut k all go to rank k, so rank 0 routes interior nets only and the last rank its stripe
		 * plus both sides of its cut (measured on 4 GPUs: 43.9 k / 50.0 k / 49.9 k / 56.2 k nets in iteration 1).  An iteration
		 * is the longest interior phase plus the longest cut phase (everybody waits at the exchange after each), so the shares
		 * are corrected until the INTERIOR fanout of every rank is level — rank 0, which owns no cut, then simply has less to do;
		 * every rank computes the same partition from the same problem. */
		std::vector<double> share((size_t)part_nranks, 1.0 / part_nranks), load((size_t)part_nranks, 0.0), load_cut((size_t)part_nranks, 0.0), best_share;
		std::vector<int> cut((size_t)part_nranks + 1, 0);
		/* the nets in box-centre order, packed (the correction rounds below read them a dozen times: 25 ms per router on cfg 4
		 * when every round walked net_bb / net_ptr through the permutation) */
		struct BoxF { int net; short xmin, xmax; int fan; };
		std::vector<BoxF> bx(byx.size());
		for (size_t k = 0; k < byx.size(); k++) {
			const int i = byx[k];
			bx[k] = BoxF{ i, (short)p->net_bb[4 * i], (short)p->net_bb[4 * i + 1], p->net_ptr[i + 1] - p->net_ptr[i] };
		}
		/* the partition the shares give, over every `step`-th net (the correction rounds look at a sample of <= 32 k nets, whose
		 * loads are within a per cent of the full ones; the final call, step 1, also writes owner[] and cut_net[]);
		 * returns the heaviest rank's load relative to the mean */
		auto partition = [&](size_t step, bool write) {
			/* cut[k] (k = 1..nranks-1): stripe k-1 holds boxes with xmax <= cut[k], stripe k boxes with xmin >= cut[k] + lmax */
			long long acc_f = 0, tot = 0;
			for (size_t k = 0; k < bx.size(); k += step) tot += bx[k].fan;
			int s = 0;
			double upto = share[0] * (double)tot;
			for (int k = 0; k <= part_nranks; k++) cut[(size_t)k] = p->nx + 2;
			std::vector<size_t> first((size_t)part_nranks + 1, bx.size());      /* first sampled position of every stripe */
			first[0] = 0;
			for (size_t k = 0; k < bx.size(); k += step) {
				while (s < part_nranks - 1 && (double)acc_f >= upto) { s++; first[(size_t)s] = k; cut[(size_t)s] = (bx[k].xmin + bx[k].xmax) / 2; upto += share[(size_t)s] * (double)tot; }
				acc_f += bx[k].fan;
			}
			for (int k = part_nranks - 1; k > 0; k--) if (first[(size_t)k] > first[(size_t)k + 1]) first[(size_t)k] = first[(size_t)k + 1];
			std::fill(load.begin(), load.end(), 0.0);
			std::fill(load_cut.begin(), load_cut.end(), 0.0);
			for (int st = 0; st < part_nranks; st++) {
				for (size_t k = first[(size_t)st]; k < first[(size_t)st + 1]; k += step) {
					const int xmin = bx[k].xmin, xmax = bx[k].xmax;
					const bool left_ok = st == 0 || xmin >= cut[(size_t)st] + lmax;
					const bool right_ok = st == part_nranks - 1 || xmax <= cut[(size_t)st + 1];
					const bool is_cut = !(left_ok && right_ok);
					/* a cut net goes to the rank that owns the violated cut (cut k belongs to rank k).  Sharing a cut's nets between its two
					 * neighbours was tried: two ranks then route overlapping nets on stale views of each other, and on a small fabric the
					 * negotiation oscillates for ever */
					const int own = is_cut ? (right_ok ? st : st + 1) : st;
					if (write) { owner[(size_t)bx[k].net] = own; cut_net[(size_t)bx[k].net] = is_cut ? 1 : 0; }
					(is_cut ? load_cut : load)[(size_t)own] += (double)bx[k].fan;
				}
			}
			/* an iteration is two phases with an exchange after each: its length is the longest interior phase plus the longest cut phase */
			double wi = 0.0, wc = 0.0;
			for (int k = 0; k < part_nranks; k++) { wi = std::max(wi, load[(size_t)k]); wc = std::max(wc, load_cut[(size_t)k]); }
			return (wi + wc) * part_nranks / std::max(1.0, (double)tot);
		};
		const size_t sample_step = std::max<size_t>(1, bx.size() / 32768);
		double best = 1e30;
		for (int round = 0; round < 12; round++) {
			const double worst = partition(sample_step, false);
			if (worst < best) { best = worst; best_share = share; }
			double sum = 0.0, mean_i = 0.0;
			for (int k = 0; k < part_nranks; k++) mean_i += load[(size_t)k] / part_nranks;
			for (int k = 0; k < part_nranks; k++) {
				const double want = std::max(1.0, mean_i) / std::max(1.0, load[(size_t)k]);
				share[(size_t)k] = std::min(4.0 / part_nranks, std::max(0.25 / part_nranks, share[(size_t)k] * pow(want, 0.5)));
				sum += share[(size_t)k];
			}
			for (int k = 0; k < part_nranks; k++) share[(size_t)k] /= sum;
		}
		share = best_share;
		partition(1, true);
		if (part_verbose) {
			fprintf(stderr, "pf_router: stripes of rank 0..%d: fanout routed (interior + across the rank's cut)", part_nranks - 1);
			for (int k = 0; k < part_nranks; k++) fprintf(stderr, " %.0f+%.0f", load[(size_t)k], load_cut[(size_t)k]);
			fprintf(stderr, "\n");
		}
	};
	std::thread part_thread;
	if (c.nranks > 1) part_thread = std::thread(partition_nets);
	Joiner part_joiner{part_thread};      /* every return path */

	/* device graph */
	r->nodes = (PfNode *)pfb_alloc_raw(sizeof(PfNode) * (size_t)r->N);
	int *gen_row = NULL;
	if (gen) {
		/* pass 1 of the device generator: out-degrees and their prefix sum give the number of edges */
		long long ne = 0;
		gen_row = (int *)pfb_alloc_raw(sizeof(int) * ((size_t)r->N + 1));
		r->gen_row = gen_row;          /* pf_router_destroy frees it (after joining the fill pass, should one be running) */
		if (!r->nodes || !gen_row || pfb_gen_count(&Gd, gen_row, &ne) != 0) { pf_router_destroy(r); CUDA_FAIL(); }
		if (ne >= 2147483647ll) { pf_router_destroy(r); FAILF(PF_EINVAL, "%lld rr edges exceed the 31-bit edge index", ne); }
		r->E = (int)ne;
	}
	r->edges = (uint32_t *)pfb_alloc_raw(sizeof(uint32_t) * (size_t)std::max(r->E, 1));
	r->ptc = (short *)pfb_alloc_raw(sizeof(short) * (size_t)r->N);
	r->sw = (PfSwitchDev *)pfb_alloc(sizeof(PfSwitchDev) * PF_MAX_SWITCHES);
	r->indexed = (PfIndexedDev *)pfb_alloc(sizeof(PfIndexedDev) * PF_MAX_INDEXED);
	r->net_ptr = (int *)pfb_alloc(sizeof(int) * ((size_t)r->n + 1));
	r->net_term = (int *)pfb_alloc(sizeof(int) * (size_t)std::max(r->T, 1));
	r->net_bb = (int *)pfb_alloc(sizeof(int) * 4 * (size_t)std::max(r->n, 1));
	r->crit = (float *)pfb_alloc(sizeof(float) * (size_t)std::max(r->T, 1));
	r->net_delay = (float *)pfb_alloc(sizeof(float) * (size_t)std::max(r->T, 1));
	if (!r->nodes || !r->edges || !r->ptc || !r->sw || !r->indexed || !r->net_ptr || !r->net_term || !r->net_bb || !r->crit || !r->net_delay) {
		pf_router_destroy(r); CUDA_FAIL();
	}
	{
		const size_t nbytes = sizeof(PfNode) * (size_t)r->N, ebytes = sizeof(uint32_t) * (size_t)std::max(r->E, 1);
		const size_t pbytes = sizeof(short) * (size_t)r->N;
		char *pin = (gen || (c.defer_graph && c.nranks > 1)) ? NULL : (char *)pfb_pinned_upload(nbytes + ebytes + pbytes + 768);
		void *stage_nodes = pin;
		uint32_t *stage_edges = pin ? (uint32_t *)(pin + ((nbytes + 255) & ~(size_t)255)) : NULL;
		short *stage_ptc = pin ? (short *)((char *)stage_edges + ((ebytes + 255) & ~(size_t)255)) : NULL;
		std::vector<PfSwitchDev> sw(PF_MAX_SWITCHES);
		for (int s = 0; s < p->num_switches; s++) { sw[s].R = p->switches[s].R; sw[s].Tdel = p->switches[s].Tdel; sw[s].buffered = p->switches[s].buffered; }
		std::vector<PfIndexedDev> ix(PF_MAX_INDEXED);
		float min_base = 0.f;
		for (int i = 0; i < p->num_indexed; i++) {
			ix[i].base_cost = p->indexed[i].base_cost; ix[i].saved_base_cost = p->indexed[i].saved_base_cost;
			ix[i].inv_length = p->indexed[i].inv_length; ix[i].T_linear = p->indexed[i].T_linear;
			ix[i].T_quadratic = p->indexed[i].T_quadratic; ix[i].C_load = p->indexed[i].C_load;
			ix[i].ortho = p->indexed[i].ortho_cost_index; ix[i].pad = 0;
			if (i >= PF_CHANX_COST_INDEX_START && (min_base == 0.f || p->indexed[i].base_cost < min_base)) min_base = p->indexed[i].base_cost;
		}
		r->win_abs_auto = 4.f * min_base;             /* a few wire hops of base cost */
		int bad_nodes = 0, bad_edges = 0;
		long long wl_avail = 0;
		r->t_mark[0] = now_s();
		if (gen) {
			/* pass 2: node records, edge words and ptc numbers written straight into HBM */
			/* PF_GEN_OVERLAP=1: the fill pass on a side stream under the allocations and uploads that follow (joined before the
			 * final sync below).  Measured on the B200 (profiles/r02H_e2e_phases.txt): create 11.4 -> 9.1 ms when one router is
			 * alive at a time, no gain next to a second live router (bench.py's e2e: 34.3 vs 34.0 ms), and one create in four
			 * took 57 ms in the allocator (stream-ordered pool memory used across two streams) — so it stays opt-in. */
			static const bool overlap = getenv("PF_GEN_OVERLAP") != NULL;
			if ((overlap ? pfb_gen_fill_begin(&Gd, gen_row, r->nodes, r->edges, r->ptc) : pfb_gen_fill(&Gd, gen_row, r->nodes, r->edges, r->ptc, &wl_avail)) != 0) { pf_router_destroy(r); CUDA_FAIL(); }
		} else if (c.defer_graph && c.nranks > 1) {
			/* the packed graph arrives from another rank (pf_comm_graph_buffers); only the available wirelength,
			 * which the first-iteration abort check needs, is computed from the host arrays here */
			r->graph_ready = 0;
			std::atomic<long long> wl(0);
			parallel_for(r->N, [&](long long lo, long long hi) {
				long long w = 0;
				for (long long i = lo; i < hi; i++)
					if (p->type[i] == PF_CHANX || p->type[i] == PF_CHANY) w += 1 + p->xhigh[i] - p->xlow[i] + p->yhigh[i] - p->ylow[i];
				wl += w;
			});
			wl_avail = wl;
		} else if (upload_edges(r, stage_edges, &bad_edges) != PF_OK || upload_nodes(r, stage_nodes, &bad_nodes, &wl_avail, stage_ptc) != PF_OK) { pf_router_destroy(r); return PF_ECUDA; }
		if (bad_nodes || bad_edges) {
			pfb_sync();
			pf_router_destroy(r);
			if (pf_problem_check(p, msg, sizeof(msg)) != PF_OK) FAILF(PF_EINVAL, "invalid problem: %s", msg);
			FAILF(PF_EINVAL, "invalid problem (capacity > 255 or cost_index > 31 on some rr node)");
		}
		r->avail_wl = wl_avail;
		if (part_thread.joinable()) part_thread.join();
		r->net_owner = owner; r->net_cut = cut_net;
		for (size_t k = 0; k < order.size(); k++) {
			if (owner[order[k]] != c.rank) continue;
			int i = order[k], ns = p->net_ptr[i + 1] - p->net_ptr[i] - 1;
			if (ns > c.sink_cap) r->work_big.push_back(i); else r->work_small.push_back(i);
		}
		if (c.big_label_log2 <= 0) {
			/* a search cannot label more rr nodes than its bounding box holds: size the big tables for twice the
			 * largest box (at the grid's average node density), capped by twice the whole graph */
			long long max_area = 1;
			for (int i : order) {
				const int *bb = &p->net_bb[4 * i];
				max_area = std::max<long long>(max_area, (long long)(bb[1] - bb[0] + 1) * (bb[3] - bb[2] + 1));
			}
			double density = (double)r->N / ((double)(p->nx + 2) * (p->ny + 2));
			long long est = (long long)(1.25 * density * (double)max_area) + 1024;
			c.big_label_log2 = std::min(22, std::max(c.label2_log2 + 1, ceil_log2(2 * std::min<long long>(est, r->N))));
		}
		if (c.big_tree_cap <= 0) c.big_tree_cap = std::max(1 << 16, 64 * max_sinks);
		if (c.big_far_cap <= 0) c.big_far_cap = std::min(1 << 21, 2 << c.big_label_log2);
		if (auto_big_slots) {
			/* the big slots take every search that outgrows a regular slot — at high pres_fac the timing-driven searches of a
			 * congested circuit flood tens of thousands of labels, thousands of nets end up here, and 64 warps on a 148-SM device
			 * made that class the whole run time (32 k-LUT stand-in: 13 s of 13.4 s).  Two warps per SM, within 32 GiB of scratch. */
			const double per_slot = (double)(sizeof(uint64_t) + sizeof(PfCold)) * (double)(1ll << c.big_label_log2) + 8.0 * c.big_far_cap
					+ (double)sizeof(PfTreeNode) * c.big_tree_cap + 1e6;
			const int by_mem = (int)std::min(32.0 * 1073741824.0 / per_slot, 1e6);
			c.big_slots = std::max(64, std::min(2 * (sms > 0 ? sms : 148), by_mem));
		}
		if (c.num_slots > (int)r->work_small.size() + 32) c.num_slots = std::max(32, (int)((r->work_small.size() + 31) / 32 * 32));

		r->t_mark[1] = now_s();
		if (pfb_h2d(r->sw, sw.data(), sizeof(PfSwitchDev) * PF_MAX_SWITCHES)
				|| pfb_h2d(r->indexed, ix.data(), sizeof(PfIndexedDev) * PF_MAX_INDEXED)
				|| pfb_h2d(r->net_ptr, p->net_ptr, sizeof(int) * ((size_t)r->n + 1))
				|| pfb_h2d(r->net_term, p->net_terminals, sizeof(int) * (size_t)r->T)
				|| pfb_h2d(r->net_bb, p->net_bb, sizeof(int) * 4 * (size_t)r->n)) { pf_router_destroy(r); CUDA_FAIL(); }
		r->h2d_bytes += (int64_t)sizeof(int) * (r->n + 1 + r->T + 4 * (int64_t)r->n);
	}
	/* initial criticalities (route_timing.c:116-128) */
	{
		std::vector<float> cr((size_t)std::max(r->T, 1), 0.f);
		float v = p->opts.timing_analysis_enabled ? 1.f : 0.f;
		for (int i = 0; i < r->n; i++)
			if (!p->net_is_global[i]) for (int k = p->net_ptr[i] + 1; k < p->net_ptr[i + 1]; k++) cr[k] = v;
		if (pfb_h2d(r->crit, cr.data(), sizeof(float) * (size_t)r->T)) { pf_router_destroy(r); CUDA_FAIL(); }
	}
	/* slots */
	r->small.num_slots = c.num_slots; r->small.label_log2 = c.label_log2; r->small.tree_cap = c.tree_cap;
	r->small.far_cap = c.far_cap; r->small.sink_cap = c.sink_cap; r->small.label2_log2 = c.label2_log2; r->big.label2_log2 = 0;
	r->big.num_slots = c.big_slots; r->big.label_log2 = c.big_label_log2; r->big.tree_cap = c.big_tree_cap;
	r->big.far_cap = c.big_far_cap; r->big.sink_cap = std::max(max_sinks, c.sink_cap);
	int nwork = (int)(r->work_small.size() + r->work_big.size());
	if (alloc_slot_class(r->small, nwork, true) || alloc_slot_class(r->big, nwork, false)) { pf_router_destroy(r); CUDA_FAIL(); }
	if (c.ripple_max_nets <= 0) c.ripple_max_nets = std::max(1024, std::min(nwork / 16, 4 * c.min_slots));   /* (256 until r02y: a 441-net
	                                                                                * circuit then rippled in its late iterations only; always on: wirelength +0.9 % instead of +4.4 %) */
	/* route store */
	/* live trees hold about one entry per used rr node; the log needs room for one iteration of re-routes on top */
	r->pool_cap = std::max<long long>(1 << 18, std::min<long long>(4ll * r->N + 96ll * r->T, 32ll * r->T + (1 << 20)));
	for (int k = 0; k < 2; k++) {
		r->pool[k] = (PfTreeNode *)pfb_alloc_raw(sizeof(PfTreeNode) * (size_t)r->pool_cap);
		r->pool_node[k] = (int *)pfb_alloc_raw(sizeof(int) * (size_t)r->pool_cap);
		if (!r->pool[k] || !r->pool_node[k]) { pf_router_destroy(r); CUDA_FAIL(); }
	}
	r->loc = (PfNetLoc *)pfb_alloc(sizeof(PfNetLoc) * (size_t)std::max(r->n, 1));
	{
		/* every net this rank owns: interior nets, then cut nets, each in the decreasing-fanout order of the
		 * reference's net loop (route_timing.c:98-106); the selection kernels keep that order in the work lists */
		std::vector<int> all;
		for (int pass = 0; pass < 2; pass++)
			for (size_t k = 0; k < order.size(); k++)
				if (owner[order[k]] == c.rank && cut_net[order[k]] == pass) all.push_back(order[k]);
		r->K1 = 0;
		for (int i : all) if (!cut_net[i]) r->K1++;
		r->h_all = all;
		r->num_all = (int)all.size();
		r->h_net_big.assign((size_t)std::max(r->n, 1), 0);
		for (int i : r->work_big) r->h_net_big[i] = 1;
		r->all_nets = (int *)pfb_alloc(sizeof(int) * (size_t)std::max(r->num_all, 1));
		r->net_big = (unsigned char *)pfb_alloc((size_t)std::max(r->n, 1));
		r->sel_scratch = (int *)pfb_alloc_raw(pfb_select_scratch_bytes(r->num_all));
		if (!r->loc || !r->all_nets || !r->net_big || !r->sel_scratch
				|| pfb_h2d(r->all_nets, all.data(), sizeof(int) * (size_t)r->num_all)
				|| pfb_h2d(r->net_big, r->h_net_big.data(), (size_t)r->n)) { pf_router_destroy(r); CUDA_FAIL(); }
	}
	/* one 256-byte control block holds every small counter the host polls, so an iteration needs one
	 * memset before and one 256-byte read after the route kernel instead of a handful of tiny copies:
	 *   [0] status[8] [32] retry_count[4] [48] sel_counts[4] [64] overused[4] [80] wl[2] [96] PfStats
	 *   [160] small.work_head[4] [176] big.work_head[4] | [192] pool_head[2] (not cleared per iteration)
	 *   [208] event_head (multi-GPU; cleared at the start of every route part) */
	r->ctl = (char *)pfb_alloc(CTL_BYTES);
	r->h_ctl = (char *)pfb_host_alloc(CTL_BYTES);
	if (!r->ctl || !r->h_ctl) { pf_router_destroy(r); CUDA_FAIL(); }
	memset(r->h_ctl, 0, CTL_BYTES);
	pfb_free(r->small.work_head); pfb_free(r->big.work_head);
	bind_ctl(r);
	r->retry_list = (int *)pfb_alloc_raw(sizeof(int) * (size_t)std::max(nwork, 1));
	r->last_over = (unsigned char *)pfb_alloc((size_t)r->N);
	r->committer = (c.keep_newcomer || c.ripple > 0) ? (int *)pfb_alloc_raw(sizeof(int) * (size_t)r->N) : NULL;
	if (r->committer && pfb_fill(r->committer, 0xff, sizeof(int) * (size_t)r->N)) { pf_router_destroy(r); CUDA_FAIL(); }
	if (c.ripple > 0) {
		r->vq_cap = std::max(nwork, 1);
		for (int k = 0; k < 2; k++) {
			r->vq[k] = (int *)pfb_alloc_raw(sizeof(int) * (size_t)r->vq_cap);
			if (!r->vq[k] || pfb_fill(r->vq[k], 0xff, sizeof(int) * (size_t)r->vq_cap)) { pf_router_destroy(r); CUDA_FAIL(); }
		}
		r->queued = (int *)pfb_alloc(sizeof(int) * (size_t)std::max(r->n, 1));
		if (!r->queued) { pf_router_destroy(r); CUDA_FAIL(); }
	}
	r->retry_work = (int *)pfb_alloc_raw(sizeof(int) * (size_t)std::max(nwork, 1));
	if (!r->retry_list || !r->retry_work || !r->last_over) { pf_router_destroy(r); CUDA_FAIL(); }
	if (c.nranks > 1) {
		/* one event per occupancy change: a route part rips up at most the live trees and commits at most what
		 * the route store can still take, so twice the store's capacity cannot overflow before the store does.
		 * The log lives in this rank's exchange region (PfXchgHeader, pf_layout.h), which the other ranks map through
		 * CUDA IPC (pf_comm_export / pf_comm_init) and read over NVLink. */
		if (c.nranks > PF_XCHG_MAX_RANKS) { pf_router_destroy(r); FAILF(PF_EINVAL, "at most %d ranks (one node)", PF_XCHG_MAX_RANKS); }
		r->event_cap = (2 * r->pool_cap + 3) & ~3ll;
		r->xreg_bytes = PF_XCHG_HEADER_BYTES + 8 * (size_t)r->event_cap + 2 * sizeof(float) * (size_t)std::max(r->T, 1);
		r->xreg = comm_region_acquire(r->xreg_bytes, r->xhandle, &r->xchg_seq, &r->dseq);   /* sequence numbers go on */
		if (!r->xreg) { pf_router_destroy(r); CUDA_FAIL(); }
		r->events = (unsigned *)(r->xreg + PF_XCHG_HEADER_BYTES);
		r->peers.base[c.rank] = r->xreg;
		std::vector<unsigned char> to((size_t)std::max(r->T, 1), 255);
		for (int i = 0; i < r->n; i++) for (int k = p->net_ptr[i]; k < p->net_ptr[i + 1]; k++) to[(size_t)k] = (unsigned char)owner[(size_t)i];
		r->term_owner = (unsigned char *)pfb_alloc((size_t)std::max(r->T, 1));
		if (!r->term_owner || pfb_h2d(r->term_owner, to.data(), (size_t)r->T)) { pf_router_destroy(r); CUDA_FAIL(); }
	}
	/* OPIN groups */
	r->num_groups = p->num_opin_groups;
	if (r->num_groups > 0) {
		std::vector<int> off((size_t)r->num_groups);
		int tot = 0;
		for (int g = 0; g < r->num_groups; g++) { off[g] = tot; tot += p->opin_group_count[g]; }
		r->g_source = (int *)pfb_alloc(sizeof(int) * (size_t)r->num_groups);
		r->g_count = (int *)pfb_alloc(sizeof(int) * (size_t)r->num_groups);
		r->g_off = (int *)pfb_alloc(sizeof(int) * (size_t)r->num_groups);
		r->g_chosen = (int *)pfb_alloc(sizeof(int) * (size_t)std::max(tot, 1));
		if (!r->g_source || !r->g_count || !r->g_off || !r->g_chosen
				|| pfb_h2d(r->g_source, p->opin_group_source, sizeof(int) * (size_t)r->num_groups)
				|| pfb_h2d(r->g_count, p->opin_group_count, sizeof(int) * (size_t)r->num_groups)
				|| pfb_h2d(r->g_off, off.data(), sizeof(int) * (size_t)r->num_groups)) { pf_router_destroy(r); CUDA_FAIL(); }
	}
	r->t_mark[2] = now_s();
	if (gen) {
		/* the fill pass ran on its side stream under the allocations and uploads above: join it, read the wirelength */
		long long wl = r->avail_wl;          /* stays as it is when the pass was not split */
		if (pfb_gen_fill_end(&wl) != 0) { pf_router_destroy(r); CUDA_FAIL(); }
		r->avail_wl = wl;
		pfb_free(r->gen_row); r->gen_row = NULL;
	}
	if (pfb_sync() != 0) { pf_router_destroy(r); CUDA_FAIL(); }
	terminal_check.join();
	if (terminals_bad) {
		pf_router_destroy(r);
		if (pf_problem_check(p, msg, sizeof(msg)) != PF_OK) FAILF(PF_EINVAL, "invalid problem: %s", msg);
		FAILF(PF_EINVAL, "invalid problem (net terminals)");
	}
	if (c.verbose) fprintf(stderr, "pf_router: create %.3f s (net check %.3f s, setup + device alloc %.3f s, flatten + upload issue %.3f s, scratch alloc %.3f s, drain %.3f s)\n",
			now_s() - t_a, t_b - t_a, r->t_mark[0] - t_b, r->t_mark[1] - r->t_mark[0], r->t_mark[2] - r->t_mark[1], now_s() - r->t_mark[2]);
	if (c.verbose)
		fprintf(stderr, "pf_router[%s] rank %d/%d: N=%d E=%d nets=%zu+%zu slots=%d(2^%d labels)+%d(2^%d) pool=%lld\n", pfb_name(), c.rank, c.nranks,
				r->N, r->E, r->work_small.size(), r->work_big.size(), c.num_slots, c.label_log2, c.big_slots, c.big_label_log2, r->pool_cap);
	*out = r;
	return PF_OK;
}

/* pf_router_create for a fabric described by generator parameters: the rr graph (node records, CSR edge words, ptc numbers)
 * is built ON the device from the closed forms of pf_gen_device.cuh — what crosses PCIe is the nets. */
extern "C" int pf_router_create_generated(const pf_gen_params *g, const pf_problem *nets, const pf_config *cfg, pf_router **out) {
	if (!g || !nets || !cfg || !out) FAILF(PF_EINVAL, "null argument");
	g_generate = g;
	const int rc = pf_router_create(nets, cfg, out);
	g_generate = NULL;
	return rc;
}

extern "C" int pf_debug_graph_hash(pf_router *r, uint64_t out[3], int64_t *num_edges) {
	if (!r || !out) FAILF(PF_EINVAL, "null argument");
	unsigned long long h[3] = { 0, 0, 0 };
	CKB(pfb_graph_hash(r->nodes, r->N, r->edges, r->E, r->ptc, h));
	out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
	if (num_edges) *num_edges = r->E;
	return PF_OK;
}

extern "C" int pf_router_reset(pf_router *r) {
	if (!r) FAILF(PF_EINVAL, "null router");
	CKB(pfb_reset_nodes(r->nodes, r->N));       /* occ = 0, acc_cost = 1: alloc_and_load_rr_node_route_structs' initial state */
	CKB(pfb_zero(r->loc, sizeof(PfNetLoc) * (size_t)std::max(r->n, 1)));
	CKB(pfb_zero(r->ctl, CTL_BYTES));
	r->sel_valid = r->sel_pending = false; r->iter_all = true; r->force_all_once = false; r->owner_valid = false;
	if (r->committer) CKB(pfb_fill(r->committer, 0xff, sizeof(int) * (size_t)r->N));
	if (r->queued) CKB(pfb_zero(r->queued, sizeof(int) * (size_t)std::max(r->n, 1)));
	r->h_pool_head = 0;
	r->iter_count = 0; r->best_overused = 0x7fffffff; r->stall_count = 0; r->over_hist.clear(); r->since_full = 0; r->cost_updates = 0; r->util = -1.;
	CKB(pfb_zero(r->last_over, (size_t)r->N));
	CKB(pfb_zero(r->net_delay, sizeof(float) * (size_t)std::max(r->T, 1)));
	{
		const pf_problem *p = r->prob;
		std::vector<float> cr((size_t)std::max(r->T, 1), 0.f);
		float v = p->opts.timing_analysis_enabled ? 1.f : 0.f;
		for (int i = 0; i < r->n; i++)
			if (!p->net_is_global[i]) for (int k = p->net_ptr[i] + 1; k < p->net_ptr[i + 1]; k++) cr[k] = v;
		CKB(pfb_h2d(r->crit, cr.data(), sizeof(float) * (size_t)r->T));
	}
	CKB(pfb_sync());
	return PF_OK;
}

/* Search granularity.  With many more nets than warps the kernel is bound by memory traffic, and strict
 * best-first order (one label per step, no bucket slack) does the least work; with few nets per warp the
 * latency of one search is what matters, and settling a whole delta bucket per step shortens it. */
static void tune_granularity(const pf_router *r, PfParams &P, int work, int slots, bool big_class = false) {
	const pf_config &c = r->cfg;
	/* a graph that does not fit in L2 pays DRAM round trips for every settled label: fewest labels wins there
	 * even with one net per warp */
	const bool in_l2 = (long long)r->N * (long long)sizeof(PfNode) + (long long)r->E * 4 < (96ll << 20);
	/* the big-slot class holds a few dozen warps: the GPU is mostly idle while they run, so what counts there is the
	 * latency of one net, never throughput (timing-driven 3.8 k-net fixture: slowest iteration 75 -> 59 ms) */
	const bool throughput = (slots > 0 && work >= 4 * slots && !big_class) || !in_l2;
	P.max_batch = c.max_batch > 0 ? c.max_batch : (throughput ? 1 : 32);
	P.pop_slack = c.pop_slack >= 0.f ? c.pop_slack : (throughput ? 0.f : 0.25f);
}

static int set_work(pf_router *r, SlotClass &s, const int *list, int count) {
	s.num_work = count;
	if (count > 0) { CKB(pfb_h2d_async(s.work, list, sizeof(int) * (size_t)count)); r->h2d_bytes += (int64_t)sizeof(int) * count; }
	return PF_OK;
}

static int slots_for(const pf_router *r, int total_work, int class_slots, int div) {
	int s = (total_work + div - 1) / div;
	s = std::max(s, r->cfg.min_slots);
	return std::max(1, std::min(s, class_slots));
}

static void fill_params(pf_router *r, PfParams &P, const SlotClass &s, float pres_fac) {
	const pf_problem *p = r->prob;
	const pf_config &c = r->cfg;
	memset(&P, 0, sizeof(P));
	P.nodes = r->nodes; P.edges = r->edges; P.num_nodes = r->N; P.nx = p->nx; P.ny = p->ny; P.node_bits = r->node_bits;
	P.sw = r->sw; P.num_sw = p->num_switches; P.indexed = r->indexed; P.num_indexed = p->num_indexed;
	P.net_ptr = r->net_ptr; P.net_term = r->net_term; P.net_bb = r->net_bb;
	P.work = s.work; P.num_work = s.num_work; P.work_head = s.work_head;
	P.crit = r->crit; P.net_delay = r->net_delay;
	P.pres_fac = pres_fac; P.astar_fac = p->opts.astar_fac; P.bend_cost = p->opts.bend_cost;
	P.max_crit = p->opts.max_criticality; P.crit_exp = p->opts.criticality_exp;
	P.pop_slack = 0.f;
	P.win_rel = c.win_rel > 0.f ? c.win_rel : 0.05f;
	P.win_abs = c.win_abs > 0.f ? c.win_abs : r->win_abs_auto;
	P.max_batch = 1;
	P.algorithm = p->opts.router_algorithm == 1 ? 1 : 0;
	P.skip_ripup = 0;
	P.validate = (c.num_slots > 1 || c.big_slots > 1) ? c.validate_commits : 0;   /* one warp cannot race */
	P.hot = s.hot; P.cold = s.cold; P.label_log2 = s.label_log2; P.epochs = s.epochs;
	P.hot2 = s.hot2; P.cold2 = s.cold2; P.label2_log2 = s.label2_log2;
	P.tree = s.tree; P.tree_cap = s.tree_cap; P.far = s.far; P.far_cap = s.far_cap; P.far_buckets = (&s == &r->big) ? 1 : 0; P.lazy_seed_min = c.lazy_seed_min;
	P.iscratch = s.iscratch; P.sink_cap = s.sink_cap;
	P.pool = r->pool[r->cur]; P.pool_node = r->pool_node[r->cur]; P.loc = r->loc; P.pool_head = r->pool_head; P.pool_cap = r->pool_cap;
	P.net_big = r->net_big;
	P.committer = r->cfg.keep_newcomer ? r->committer : NULL;
	if (r->vq[0] && !r->iter_all && r->n_small + r->n_big <= r->cfg.ripple_max_nets) {   /* (an iteration that re-routes every net displaces nobody unrouted) */
		P.committer = r->committer;          /* holders are tracked only while ripple re-routing is on (launch_routes rebuilds them) */
		P.vq[0] = r->vq[0]; P.vq[1] = r->vq[1]; P.vq_ctl = (int *)(r->ctl + CTL_VQ); P.vq_cap = r->vq_cap;
		P.vq_class = (&s == &r->big) ? 1 : 0; P.queued = r->queued; P.iter_tag = r->iter_count;
	}
	P.events = r->events ? r->events + (size_t)(r->xchg_seq & 1) * (size_t)r->event_cap : NULL;   /* double-buffered by exchange parity */
	P.event_head = (unsigned long long *)(r->ctl + CTL_EVENTS); P.event_cap = r->event_cap;
	P.status = r->status; P.retry_list = r->retry_list; P.retry_count = r->retry_count; P.stats = r->stats;
}

/* the select kernels: work lists of the iteration about to start (tag = its number) */
static int launch_select(pf_router *r, int force_all, bool right_behind_the_cost_update = false) {
	/* right behind the cost update the byte map it wrote (last_over[v] == tag <=> v is overused now) replaces the node records */
	const int tag = 1 + (r->cost_updates + 254) % 255;
	const bool fast = right_behind_the_cost_update && r->cfg.history_window <= 0 && !r->cfg.keep_newcomer
			&& r->cost_updates < 255;     /* tags are iteration numbers mod 255: beyond that an old mark could alias */
	CKB(pfb_launch_select_nets(r->nodes, r->pool[r->cur], r->loc, r->all_nets, r->num_all, r->net_big, force_all,
			r->small.work, r->big.work, r->sel_counts, r->cfg.history_window > 0 ? r->last_over : NULL,
			1 + (r->cost_updates + 254) % 255, r->cfg.history_window, r->cfg.keep_newcomer ? r->committer : NULL, r->sel_scratch, r->K1,
			r->queued, r->iter_count + 1, r->pool_node[r->cur], fast ? r->last_over : NULL, tag));
	return PF_OK;
}

/* ONE read of the control block: everything the host needs to know about the launches since the last read.
 * Raises the device-side error bits as PF_E* codes. */
static int fetch_ctl(pf_router *r) {
	CKB(pfb_d2h(r->h_ctl, r->ctl, CTL_BYTES));
	r->d2h_bytes += CTL_BYTES;
	const char *h = r->h_ctl;
	int h_status[8], h_retry[4];
	memcpy(h_status, h + CTL_STATUS, sizeof(h_status));
	memcpy(h_retry, h + CTL_RETRY, sizeof(h_retry));
	memcpy(&r->h_pool_head, h + CTL_POOL, sizeof(unsigned long long));
	{ unsigned long long ne = 0; memcpy(&ne, h + CTL_EVENTS, sizeof(ne)); r->h_events = (long long)ne; }
	memcpy(&r->h_stats, h + CTL_STATS, sizeof(PfStats));
	if (r->events && r->h_events > r->event_cap) FAILF(PF_EOVERFLOW, "occupancy event log overflow (%lld events, capacity %lld)", r->h_events, r->event_cap);
	if (h_status[0] & PF_ST_POOL_OVERFLOW) FAILF(PF_EOVERFLOW, "route store overflow (capacity %lld tree entries)", r->pool_cap);
	if (h_status[0] & PF_ST_BIG_OVERFLOW) FAILF(PF_EOVERFLOW, "net %d overflows the big slots (label 2^%d, tree %d, far %d)", h_status[2],
			r->big.label_log2, r->big.tree_cap, r->big.far_cap);
	if (h_status[0] & PF_ST_TWICE_TO_SINK_BF)
		FAILF(PF_EINVAL, "breadth-first router: net %d connects twice to one SINK; the reference's heap surgery for this case "
				"(route_breadth_first.c:208-256) is not supported — its own check_route rejects the routing it produces. "
				"Use the timing-driven / no-timing router (router_algorithm 0), which routes such nets", h_status[2]);
	if (h_status[0] & PF_ST_COMM_TIMEOUT) FAILF(PF_ECUDA, "multi-GPU exchange timed out waiting for a peer rank's event log");
	if (h_status[0] & PF_ST_COMM_ABORT) FAILF(PF_ECUDA, "a peer rank reported a failure during the occupancy exchange");
	if (h_status[0] & PF_ST_INTERNAL) FAILF(PF_ECUDA, "internal error in the device router (net %d)", h_status[2]);
	if (h_status[0] & PF_ST_UNROUTABLE) FAILF(PF_EUNROUTABLE, "net %d has no possible path (disconnected rr graph)", h_status[2]);
	if (h_retry[0] > 0) {
		/* nets that outgrew a regular slot were re-routed in the big slots by the retry launch and stay in that class:
		 * mirror the device's net_big[] for the host-built work lists of "every net" iterations */
		std::vector<int> lst((size_t)h_retry[0]);
		CKB(pfb_d2h(lst.data(), r->retry_list, sizeof(int) * lst.size()));
		for (int i : lst) r->h_net_big[(size_t)i] = 1;
		if (r->cfg.verbose) fprintf(stderr, "pf_router: %d nets moved to the big slots\n", h_retry[0]);
	}
	return PF_OK;
}

/* Start of one PathFinder iteration: garbage-collect the route store and choose the nets to re-route.
 * No host-device synchronisation on the common path: the counts of the work lists were computed by the select
 * kernels launched behind the previous iteration's cost update and arrived with that iteration's control block. */
extern "C" int pf_iteration_begin(pf_router *r, const float *crit) {
	if (!r) FAILF(PF_EINVAL, "null router");
	if (!r->graph_ready) FAILF(PF_EINVAL, "the router was created with defer_graph: fill the graph buffers and call pf_comm_graph_ready first");
	int rc;
	if (crit) { CKB(pfb_h2d(r->crit, crit, sizeof(float) * (size_t)r->T)); r->h2d_bytes += (int64_t)sizeof(float) * r->T; }
	/* garbage-collect the route-tree log when it is more than half full */
	if ((long long)r->h_pool_head > r->pool_cap / 2) {
		CKB(pfb_zero(r->pool_head, sizeof(unsigned long long) * 2));
		CKB(pfb_launch_compact(r->pool[r->cur], r->pool[r->cur ^ 1], r->pool_node[r->cur], r->pool_node[r->cur ^ 1], r->loc, r->all_nets, r->num_all, r->pool_head));
		r->cur ^= 1;
		r->h_pool_head = 0;       /* the live size arrives with the next control block; a log that is still too small then
		                           * overflows in the route kernel and is reported as such */
		r->sel_valid = false;     /* the lists are fine, but keep the rare path simple */
	}
	/* Congested-only re-routing negotiates locally; on a tight instance it can ping-pong between a few
	 * nets while every free resource nearby is held by legal nets.  When the overuse has not improved
	 * for stall_iters iterations, fall back to the serial reference's policy for one iteration — every
	 * net ripped up and re-routed (route_timing.c:161-183) — with few nets in flight. */
	/* stalled: the overuse has not dropped by 30 % over the last stall_iters+1 congested-only iterations
	 * (a healthy negotiation roughly halves it every iteration) */
	const int K = r->cfg.stall_iters + 1;
	const int H = (int)r->over_hist.size();
	const bool stalled = r->cfg.stall_iters > 0 && r->since_full >= K && H > K
			&& (double)r->over_hist[H - 1] > 0.7 * (double)r->over_hist[H - 1 - K];
	if (stalled) { r->stall_count = 0; if (r->cfg.verbose) fprintf(stderr, "pf_router: overuse stalled at %d, re-routing every net\n", r->best_overused); }
	const bool all = stalled || r->force_all_once || r->cfg.reroute_all_iters < 0 || r->iter_count < r->cfg.reroute_all_iters;
	r->force_all_once = false;
	/* how many nets may be in flight depends on how contested the fabric is: with the channels under 40 % full after
	 * the first iteration (BASELINE configs[4]: 29 %; the near-minimum-width fixtures: 52-63 %) twice as many nets
	 * in flight converge just as fast (cfg 4: 22.9 vs 24.9 ms measured), on a tight fabric they do not */
	/* ... and with the channels over 40 % full half as many: measured on the B200 over 6 runs each of the near-minimum-width
	 * fixtures (profiles/r02y_repeat.txt), divisor 32 instead of 16 narrows the run-to-run spread of the iteration count from
	 * 23-38 to 20-24 (heq, reference 25), 21-28 to 20-23 (toy), 21-30 to 20-22 (hub) and lowers the criticality-weighted delay
	 * by about 1 %: fewer nets commit against a congestion picture that is already stale */
	int base_div = r->cfg.inflight_div;
	/* (under 40 %: 16 / 8 = 2 — cfg 4 on the B200, divisor 8 / 4 / 2 / 1: 16.58 / 16.05 / 15.89 / 15.77 ms per routing, 6 iterations and
	 * the wirelength within 0.01 % each time) */
	if (!r->div_explicit && r->util >= 0.) base_div = r->util < 0.40 ? std::max(1, base_div / 8) : base_div * 2;
	/* several ranks: a rank holds 1 / nranks of the nets, and the same divisor would leave most of its warps without work
	 * (4 GPUs, cfg 4, iteration 3: 9.7 k nets on 1213 of 2960 warps, eight nets deep — as long as one GPU takes for all 37 k).
	 * The nets in flight per rank stay what one GPU has in flight; measured on one GPU, cfg 4 converges the same with every
	 * net in flight (inflight_div 8 / 4 / 2 / 1: 6 iterations, wirelength within 0.01 %). */
	r->cur_div = stalled ? r->cfg.inflight_div * 8 : std::max(1, base_div / std::max(1, r->cfg.nranks));
	r->iter_all = all;
	if (all) {
		/* host-built lists in the reference's net order (route_timing.c:98-106): interior nets, then cut nets */
		std::vector<int> &sm = r->h_list_small, &bg = r->h_list_big;
		sm.clear(); bg.clear();
		r->n1_small = r->n1_big = 0;
		for (int k = 0; k < r->num_all; k++) {
			const int i = r->h_all[(size_t)k];
			(r->h_net_big[i] ? bg : sm).push_back(i);
			if (k < r->K1) (r->h_net_big[i] ? r->n1_big : r->n1_small)++;
		}
		if ((rc = set_work(r, r->small, sm.data(), (int)sm.size())) != PF_OK) return rc;
		if ((rc = set_work(r, r->big, bg.data(), (int)bg.size())) != PF_OK) return rc;
		r->n_small = (int)sm.size(); r->n_big = (int)bg.size();
	} else {
		/* the work lists come back in the fanout order of the reference's net loop (route_timing.c:98-106): long
		 * nets start first and runs are reproducible */
		if (!r->sel_valid) {
			if ((rc = launch_select(r, 0)) != PF_OK) return rc;
			if ((rc = fetch_ctl(r)) != PF_OK) return rc;
		}
		int counts[4];
		memcpy(counts, r->h_ctl + CTL_SEL, sizeof(counts));
		r->n_small = counts[0]; r->n_big = counts[1];
		r->n1_small = counts[2]; r->n1_big = counts[3];
	}
	r->sel_valid = false;
	CKB(pfb_zero(r->ctl, CTL_SEL));           /* status and counters of the iteration */
	memset(&r->h_stats_seen, 0, sizeof(PfStats));
	r->iter_count++;
	r->since_full = all ? 0 : r->since_full + 1;
	return PF_OK;
}

/* Launches of one route part (no synchronisation): the big class (long nets first), the regular class, then the
 * retry launch in the big slots — nets that outgrew a regular slot (counted on the device) and, with ripple
 * re-routing, big-class victims displaced by regular nets. */
static int launch_routes(pf_router *r, float pres_fac, int part, int nparts) {
	CKB(pfb_zero(r->ctl + CTL_PART, CTL_PART_END - CTL_PART));
	const int div = r->cur_div;
	/* several ranks, two parts: interior nets, then cut nets (see pf_router_create); otherwise equal slices */
	const bool by_class = r->cfg.nranks > 1 && nparts == 2;
	auto slice = [&](int n, int n1, int &off, int &cnt) {
		if (by_class) { off = part ? n1 : 0; cnt = part ? n - n1 : n1; }
		else { off = (int)((long long)n * part / nparts); cnt = (int)((long long)n * (part + 1) / nparts) - off; }
	};
	int so, sc, bo, bc;
	slice(r->n_small, r->n1_small, so, sc); slice(r->n_big, r->n1_big, bo, bc);
	PfParams P;
	const int total = r->n_small + r->n_big;      /* the staleness bound is about the whole iteration's nets */
	const bool ripple = r->vq[0] && !r->iter_all && r->n_small + r->n_big <= r->cfg.ripple_max_nets;
	if (ripple && !r->owner_valid) {
		/* who holds which rr node: not tracked while ripple is off (one random 4-byte atomic per committed node is a fifth of the
		 * commit traffic of a 200 k-net iteration), rebuilt from the route store when it comes on */
		CKB(pfb_fill(r->committer, 0xff, sizeof(int) * (size_t)r->N));
		CKB(pfb_launch_rebuild_owner(r->pool[r->cur], r->loc, r->all_nets, r->num_all, r->committer));
		r->owner_valid = true;
	}
	if (!ripple && !r->cfg.keep_newcomer) r->owner_valid = false;
	if (bc > 0) {
		r->big.num_work = bc;
		fill_params(r, P, r->big, pres_fac);
		P.work = r->big.work + bo;
		const int sl = slots_for(r, total, std::min(r->big.num_slots, bc), div);
		tune_granularity(r, P, bc, sl, true);
		CKB(pfb_launch_route(&P, sl, 1));
	}
	if (sc > 0) {
		r->small.num_work = sc;
		fill_params(r, P, r->small, pres_fac);
		P.work = r->small.work + so;
		const int sl = slots_for(r, total, r->small.num_slots, div);
		tune_granularity(r, P, sc, sl);
		CKB(pfb_launch_route(&P, sl, r->cfg.warps_per_block));
	}
	if (sc > 0 || (ripple && bc > 0)) {
		/* nets whose scratch overflowed in a regular slot are re-routed in the big slots, and stay there (the kernel marks
		 * net_big[]); the length of the list is read on the device, so nobody waits for the host */
		r->big.num_work = 0;
		fill_params(r, P, r->big, pres_fac);
		P.work = r->retry_list; P.num_work_ptr = r->retry_count; P.work_head = (int *)(r->ctl + CTL_HEAD_RETRY);
		P.skip_ripup = 1;
		const int sl = r->big.num_slots;             /* (warps without work leave at once) */
		tune_granularity(r, P, sl, sl, true);
		CKB(pfb_launch_route(&P, sl, 1));
	}
	return PF_OK;
}

/* Route slice `part` of `nparts` of this iteration's nets (nparts > 1: multi-GPU sub-rounds with an
 * occupancy sync after each, so ranks see each other's routes several times per iteration). */
extern "C" int pf_iteration_route_part(pf_router *r, float pres_fac, int part, int nparts, pf_iter_stats *st) {
	if (!r || nparts < 1 || part < 0 || part >= nparts) FAILF(PF_EINVAL, "bad argument");
	int rc;
	if ((rc = launch_routes(r, pres_fac, part, nparts)) != PF_OK) return rc;
	if ((rc = fetch_ctl(r)) != PF_OK) return rc;
	if (st) {
		/* the device counters accumulate over the iteration: report this part's share */
		const PfStats &hs = r->h_stats, &seen = r->h_stats_seen;
		memset(st, 0, sizeof(*st));
		st->nets_routed = (int)(hs.nets - seen.nets); st->heap_pushes = (int64_t)(hs.pushes - seen.pushes); st->heap_pops = (int64_t)(hs.pops - seen.pops);
		st->edge_visits = (int64_t)(hs.visits - seen.visits); st->pres_fac = pres_fac;
	}
	r->h_stats_seen = r->h_stats;
	return PF_OK;
}

extern "C" int pf_route_iteration(pf_router *r, float pres_fac, const float *crit, pf_iter_stats *st) {
	int rc = pf_iteration_begin(r, crit);
	if (rc != PF_OK) return rc;
	return pf_iteration_route_part(r, pres_fac, 0, 1, st);
}

extern "C" int pf_reserve_opins(pf_router *r, float pres_fac, int rip_up) {
	if (!r) FAILF(PF_EINVAL, "null router");
	if (r->num_groups == 0) return PF_OK;
	r->sel_valid = false;
	CKB(pfb_launch_reserve_opins(r->nodes, r->edges, r->node_bits, r->indexed, r->num_groups, r->g_source, r->g_count, r->g_off, r->g_chosen, rip_up, pres_fac));
	return PF_OK;
}

/* feasible_routing + pathfinder_update_cost in one pass over the node records; the same pass sums the wirelength in
 * use on the fabric (occupancy x length of every CHANX / CHANY node: with several ranks that is the whole routing,
 * not only this rank's trees), and the select kernels for the NEXT iteration run right behind it, so that one read
 * of the control block ends the iteration. */
static int update_costs_async(pf_router *r, float acc_fac) {
	CKB(pfb_zero(r->ctl + CTL_OVER, CTL_PART - CTL_OVER));
	r->cost_updates++;
	CKB(pfb_launch_update_cost(r->nodes, r->N, acc_fac, r->d_overused, r->last_over, 1 + (r->cost_updates + 254) % 255, r->d_wl + 1));
	if (r->cfg.reroute_all_iters >= 0 && r->iter_count >= r->cfg.reroute_all_iters) {
		int rc = launch_select(r, 0, true);
		if (rc != PF_OK) return rc;
		r->sel_pending = true;
	}
	return PF_OK;
}

static int update_costs_finish(pf_router *r, int *overused) {
	int rc = fetch_ctl(r);
	if (rc != PF_OK) return rc;
	int h[4];
	memcpy(h, r->h_ctl + CTL_OVER, sizeof(h));
	unsigned long long wl[2];
	memcpy(wl, r->h_ctl + CTL_WL, sizeof(wl));
	r->h_wl_used = (long long)wl[1];
	if (r->iter_count <= 1 && r->avail_wl > 0) r->util = (double)wl[1] / (double)r->avail_wl;
	r->sel_valid = r->sel_pending; r->sel_pending = false;
	if (overused) *overused = h[0];
	if (h[0] < r->best_overused) { r->best_overused = h[0]; r->stall_count = 0; } else r->stall_count++;
	r->over_hist.push_back(h[0]);
	return PF_OK;
}

extern "C" int pf_update_costs(pf_router *r, float acc_fac, int *overused) {
	if (!r) FAILF(PF_EINVAL, "null router");
	int rc = update_costs_async(r, acc_fac);
	if (rc != PF_OK) return rc;
	return update_costs_finish(r, overused);
}

/* Multi-GPU occupancy sync.  Every occupancy change a rank makes while routing (rip-up, commit, undo) is
 * also appended to its event log; after a route part the ranks all-gather their logs and replay the others'.
 * What crosses NVLink is 4 bytes per changed rr node instead of a dense int32[num_rr_nodes] all-reduce, and no
 * pass over the node array is needed on either side. */
extern "C" int pf_comm_graph_buffers(pf_router *r, void *dev_ptrs[3], int64_t bytes[3]) {
	if (!r || !dev_ptrs || !bytes) FAILF(PF_EINVAL, "null argument");
	dev_ptrs[0] = r->nodes; bytes[0] = (int64_t)sizeof(PfNode) * r->N;
	dev_ptrs[1] = r->edges; bytes[1] = (int64_t)sizeof(uint32_t) * std::max(r->E, 1);
	dev_ptrs[2] = r->ptc; bytes[2] = (int64_t)sizeof(short) * r->N;
	CKB(pfb_sync());                    /* the sender's uploads have landed before anybody reads the buffers */
	return PF_OK;
}

extern "C" int pf_comm_graph_ready(pf_router *r) {
	if (!r) FAILF(PF_EINVAL, "null router");
	r->graph_ready = 1;
	return PF_OK;
}

/* introspection of the stripe sharding: for every net the rank that routes it and whether it is a cut net */
extern "C" int pf_comm_net_classes(pf_router *r, int32_t *owner, uint8_t *is_cut) {
	if (!r || !owner || !is_cut) FAILF(PF_EINVAL, "null argument");
	for (int i = 0; i < r->n; i++) { owner[i] = r->net_owner[(size_t)i]; is_cut[i] = r->net_cut[(size_t)i]; }
	return PF_OK;
}

/* ---- the transport inside the library: peer memory over NVLink / NVSwitch (one process per GPU, one node).
 * Bootstrap (once): every rank calls pf_comm_export, the caller all-gathers the PF_COMM_HANDLE_BYTES blobs with whatever
 * it has (MPI_Allgather in the reference's MPI router, torch.distributed here) and hands all of them to pf_comm_init.
 * From then on nothing crosses the host: pf_comm_exchange is one kernel behind the route kernels. */
struct CommHandle { unsigned char ipc[64]; uint64_t bytes; int64_t event_cap; int32_t rank, nranks, T, magic; uint32_t seq, dseq; };   /* seq / dseq: where this rank's
	                                * sequence numbers stand (the region and its numbers outlive the router, see the transport cache above) */
static_assert(sizeof(CommHandle) <= PF_COMM_HANDLE_BYTES, "handle blob too small");
#define PF_COMM_MAGIC 0x50465832   /* "PFX2" */

static double comm_timeout_s() { const char *e = getenv("PF_COMM_TIMEOUT_S"); double v = e ? atof(e) : 0.; return v > 0. ? v : 30.; }

extern "C" int pf_comm_export(pf_router *r, void *handle) {
	if (!r || !handle) FAILF(PF_EINVAL, "null argument");
	if (!r->xreg) FAILF(PF_EINVAL, "router was created with nranks == 1");
	CommHandle h;
	memset(&h, 0, sizeof(h));
	memcpy(h.ipc, r->xhandle, 64);
	h.bytes = r->xreg_bytes; h.event_cap = r->event_cap; h.rank = r->cfg.rank; h.nranks = r->cfg.nranks; h.T = r->T; h.magic = PF_COMM_MAGIC; h.seq = r->xchg_seq; h.dseq = r->dseq;
	memset(handle, 0, PF_COMM_HANDLE_BYTES);
	memcpy(handle, &h, sizeof(h));
	return PF_OK;
}

extern "C" int pf_comm_init(pf_router *r, const void *all_handles) {
	if (!r || !all_handles) FAILF(PF_EINVAL, "null argument");
	if (!r->xreg) FAILF(PF_EINVAL, "router was created with nranks == 1");
	if (r->comm_ready) return PF_OK;
	unsigned seq = r->xchg_seq, dseq = r->dseq;
	for (int k = 0; k < r->cfg.nranks; k++) {
		CommHandle h;
		memcpy(&h, (const unsigned char *)all_handles + (size_t)k * PF_COMM_HANDLE_BYTES, sizeof(h));
		if (h.magic != PF_COMM_MAGIC || h.rank != k || h.nranks != r->cfg.nranks) FAILF(PF_EINVAL, "handle %d is not rank %d's of %d", k, k, r->cfg.nranks);
		if (h.event_cap != r->event_cap || h.T != r->T || h.bytes != r->xreg_bytes) FAILF(PF_EINVAL, "rank %d routes a different problem", k);
		/* every rank starts from the furthest number any rank has published (they agree unless a rank joined late or
		 * failed half-way through a routing): the waits are "at least", so a stale header can only be behind */
		if ((int)(h.seq - seq) > 0) seq = h.seq;
		if ((int)(h.dseq - dseq) > 0) dseq = h.dseq;
		if (k == r->cfg.rank) continue;
		r->peers.base[k] = (unsigned char *)comm_peer_open(h.ipc);
		if (!r->peers.base[k]) CUDA_FAIL();
	}
	r->xchg_seq = seq; r->dseq = dseq;
	r->comm_ready = 1;
	return PF_OK;
}

/* After a route part: publish this rank's occupancy event log, replay every peer's.  Stream-ordered, no host wait. */
extern "C" int pf_comm_exchange(pf_router *r) {
	if (!r) FAILF(PF_EINVAL, "null router");
	if (!r->comm_ready) FAILF(PF_EINVAL, "pf_comm_init has not been called");
	r->xchg_seq++;                          /* the route kernels since the last exchange wrote log buffer (xchg_seq - 1) & 1 */
	r->sel_valid = false;
	CKB(pfb_launch_xchg_events(r->nodes, &r->peers, r->cfg.rank, r->cfg.nranks, r->xchg_seq, (const unsigned long long *)(r->ctl + CTL_EVENTS),
			r->event_cap, r->status, comm_timeout_s()));
	return PF_OK;
}

/* Sink delays of the nets other ranks route, into this rank's delay vector (before a timing analysis / the result). */
extern "C" int pf_comm_gather_delays(pf_router *r) {
	if (!r) FAILF(PF_EINVAL, "null router");
	if (!r->comm_ready) FAILF(PF_EINVAL, "pf_comm_init has not been called");
	r->dseq++;
	CKB(pfb_launch_xchg_delays(r->net_delay, r->term_owner, r->T, &r->peers, r->cfg.rank, r->cfg.nranks, r->dseq, r->event_cap, r->status, comm_timeout_s()));
	return PF_OK;
}

extern "C" int pf_comm_abort(pf_router *r) {
	if (!r || !r->xreg) return PF_OK;
	pfb_launch_xchg_abort(&r->peers, r->cfg.rank);
	pfb_sync();
	return PF_OK;
}

extern "C" int pf_comm_events(pf_router *r, void **dev_events, int64_t *count) {
	if (!r || !dev_events || !count) FAILF(PF_EINVAL, "null argument");
	if (!r->events) FAILF(PF_EINVAL, "router was created with nranks == 1");
	*dev_events = r->events + (size_t)(r->xchg_seq & 1) * (size_t)r->event_cap;
	*count = (int64_t)r->h_events;
	return PF_OK;
}

extern "C" int pf_comm_apply_events(pf_router *r, const void *dev_events, int64_t count) {
	if (!r || (!dev_events && count > 0)) FAILF(PF_EINVAL, "null argument");
	if (!r->events) FAILF(PF_EINVAL, "router was created with nranks == 1");
	r->sel_valid = false;
	CKB(pfb_launch_apply_events(r->nodes, (const unsigned *)dev_events, (long long)count));   /* stream-ordered; the caller keeps
	                                                                                          * the buffer alive until the next call */
	return PF_OK;
}

extern "C" void *pf_comm_net_delay_ptr(pf_router *r) { return r ? (void *)r->net_delay : NULL; }
extern "C" void *pf_comm_crit_ptr(pf_router *r) { return r ? (void *)r->crit : NULL; }

extern "C" int pf_total_wirelength(pf_router *r, int64_t *wl, int64_t *avail) {
	if (!r) FAILF(PF_EINVAL, "null router");
	CKB(pfb_zero(r->d_wl, sizeof(unsigned long long) * 2));
	CKB(pfb_launch_wirelength(r->pool[r->cur], r->loc, r->all_nets, r->num_all, r->d_wl));
	unsigned long long h[2];
	CKB(pfb_d2h(h, r->d_wl, sizeof(h)));
	r->d2h_bytes += 32;
	if (wl) *wl = (int64_t)h[0];
	if (avail) *avail = r->avail_wl;
	if (r->iter_count <= 1 && r->avail_wl > 0) r->util = (double)h[0] * r->cfg.nranks / (double)r->avail_wl;   /* ranks hold equal shares */
	return PF_OK;
}

extern "C" int pf_get_net_delay(pf_router *r, float *net_delay) {
	if (!r || !net_delay) FAILF(PF_EINVAL, "null argument");
	CKB(pfb_d2h(net_delay, r->net_delay, sizeof(float) * (size_t)r->T));
	r->d2h_bytes += (int64_t)sizeof(float) * r->T;
	return PF_OK;
}

extern "C" int pf_timer_start(pf_router *r) { if (!r) FAILF(PF_EINVAL, "null router"); CKB(pfb_timer_start()); return PF_OK; }
extern "C" int pf_timer_stop(pf_router *r, double *ms) { if (!r || !ms) FAILF(PF_EINVAL, "null argument"); CKB(pfb_timer_stop(ms)); return PF_OK; }
extern "C" void *pf_stream(pf_router *r) { (void)r; return pfb_stream(); }

extern "C" int pf_get_timing(pf_router *r, pf_timing *t, int reset) {
	if (!r || !t) FAILF(PF_EINVAL, "null argument");
	PfLaunchTimes lt;
	pfb_times(&lt, reset);
	t->route_kernel_ms = lt.route_ms; t->update_kernel_ms = lt.update_ms; t->aux_kernel_ms = lt.aux_ms;
	t->route_launches = lt.route_launches; t->update_launches = lt.update_launches; t->aux_launches = lt.aux_launches;
	t->h2d_bytes = r->h2d_bytes; t->d2h_bytes = r->d2h_bytes;
	if (reset) { r->h2d_bytes = 0; r->d2h_bytes = 0; }
	return PF_OK;
}

/* ---- result arrays: pinned host buffers, kept for the next result.
 * A result of cfg 4 is 120 MB (traces 6 bytes per element, occupancy 4 bytes per rr node).  Freshly malloc'ed arrays cost a
 * page fault per 4 KB when they are first written (measured: 8-9 ms for the copy out of the pinned staging buffer, against
 * 2.4 ms for the PCIe transfer itself); the arrays of pf_result therefore ARE pinned buffers the device copies into directly,
 * and pf_result_free (pf_file.c, through the release hook) returns them to this cache instead of the heap. */
struct HostBuf { void *p; size_t cap; bool in_use; };
static std::vector<HostBuf> g_hostbufs;
static std::mutex g_host_mu;
static const size_t HOSTBUF_KEEP_BYTES = (size_t)1 << 30;      /* free buffers kept: at most this much, at most 16 */
static void *host_take(size_t bytes) {
	if (bytes < 16) bytes = 16;
	std::lock_guard<std::mutex> lk(g_host_mu);
	HostBuf *best = NULL;
	int have_class = 0;           /* buffers of this size that exist, free or not */
	for (HostBuf &b : g_hostbufs) {
		if (b.cap < bytes || b.cap > 2 * bytes + (1 << 20)) continue;
		have_class++;
		if (!b.in_use && (!best || b.cap < best->cap)) best = &b;
	}
	if (best) { best->in_use = true; return best->p; }
	/* at most TWO pinned buffers per size (a caller that reads result k while it asks for result k + 1 alternates between them):
	 * pinning 120 MB costs tens of milliseconds (measured 30-160 ms for the arrays of one cfg 4 result), so a caller that holds
	 * more results than that gets malloc memory filled through the staging buffer instead (NULL here) — bounded pinned memory,
	 * and the pinning cost is paid twice per size at most */
	if (have_class >= 2) return NULL;
	const size_t cap = bytes + bytes / 16;
	void *p = pfb_host_alloc(cap);
	if (!p) return NULL;
	g_hostbufs.push_back(HostBuf{ p, cap, true });
	return p;
}
static int host_release(void *p) {
	std::lock_guard<std::mutex> lk(g_host_mu);
	size_t at = g_hostbufs.size();
	for (size_t i = 0; i < g_hostbufs.size(); i++) if (g_hostbufs[i].p == p) at = i;
	if (at == g_hostbufs.size()) return 0;                      /* not ours: plain malloc memory */
	g_hostbufs[at].in_use = false;
	size_t free_bytes = 0, free_count = 0;
	for (const HostBuf &b : g_hostbufs) if (!b.in_use) { free_bytes += b.cap; free_count++; }
	for (size_t i = 0; i < g_hostbufs.size() && (free_bytes > HOSTBUF_KEEP_BYTES || free_count > 16); ) {      /* oldest first */
		if (g_hostbufs[i].in_use) { i++; continue; }
		free_bytes -= g_hostbufs[i].cap; free_count--;
		pfb_host_free(g_hostbufs[i].p);
		g_hostbufs.erase(g_hostbufs.begin() + (long)i);
	}
	return 1;
}

/* Route store → s_trace-ordered lists (update_traceback, route_common.c:638-706): the first
 * segment runs SOURCE … SINK; every later segment starts with its join node (whose iswitch is
 * the switch into the first new node) and ends at a SINK (iswitch OPEN). */
extern "C" int pf_get_result(pf_router *r, pf_result *out) {
	if (!r || !out) FAILF(PF_EINVAL, "null argument");
	const pf_problem *p = r->prob;
	memset(out, 0, sizeof(*out));
	const int n = r->n;
	const double t_0 = now_s();
	/* the traces are assembled on the device (pf_build_traces_kernel), so what crosses PCIe is the final
	 * 6 bytes per trace element plus 4 bytes per rr node of occupancy — not the 32-byte tree entries */
	int *d_len = (int *)pfb_alloc_raw(sizeof(int) * ((size_t)n + 1));
	int *d_occ = (int *)pfb_alloc_raw(sizeof(int) * (size_t)r->N);
	if (!d_len || !d_occ) { pfb_free(d_len); pfb_free(d_occ); CUDA_FAIL(); }
	std::vector<int32_t> tptr((size_t)n + 1, 0);
	int bad = pfb_launch_build_traces(r->pool[r->cur], r->loc, n, d_len, NULL, NULL, NULL, NULL, NULL, NULL, 0) || pfb_d2h(tptr.data() + 1, d_len, sizeof(int) * (size_t)n);
	if (bad) { pfb_free(d_len); pfb_free(d_occ); CUDA_FAIL(); }
	for (int i = 0; i < n; i++) tptr[i + 1] += tptr[i];
	const size_t total = (size_t)tptr[n];
	int *d_tn = (int *)pfb_alloc_raw(sizeof(int) * std::max<size_t>(total, 1));
	short *d_ts = (short *)pfb_alloc_raw(sizeof(short) * std::max<size_t>(total, 1));
	unsigned *d_tt = (unsigned *)pfb_alloc_raw(sizeof(unsigned) * std::max<size_t>(total, 1));
	out->num_nets = n;
	pf_result_set_release_hook(host_release);
	/* each array: a pinned buffer of the cache the device copies into directly, or — cache busy / no pinned memory — malloc
	 * memory filled through the process-wide pinned staging buffer */
	const size_t b_tn = sizeof(int) * total, b_ts = sizeof(short) * total, b_tt = sizeof(unsigned) * total, b_occ = sizeof(int) * (size_t)r->N;
	const size_t b_nd = sizeof(float) * (size_t)std::max(r->T, 1);
	struct Arr { void **out; size_t bytes; const void *dev; bool direct; char *stage; };
	Arr arr[4] = { { (void **)&out->trace_node, b_tn, d_tn, false, NULL }, { (void **)&out->trace_switch, b_ts, d_ts, false, NULL },
			{ (void **)&out->occ, b_occ, d_occ, false, NULL }, { (void **)&out->net_delay, b_nd, r->net_delay, false, NULL } };
	out->trace_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
	size_t stage_bytes = 0;
	bool host_ok = out->trace_ptr != NULL;
	for (Arr &a : arr) {
		*a.out = host_take(std::max<size_t>(a.bytes, 16));
		a.direct = *a.out != NULL;
		if (!a.direct) { *a.out = malloc(std::max<size_t>(a.bytes, 16)); stage_bytes += (a.bytes + 255) & ~(size_t)255; }
		host_ok = host_ok && *a.out != NULL;
	}
	unsigned long long h_wl[2] = { 0, 0 };
	int serial_num = 0;
	bad = !d_tn || !d_ts || !d_tt || !host_ok;
	double t_1 = t_0, t_2 = t_0;
	if (!bad) {
		/* the serial-number terms: scratch of this call, after the staged arrays in the staging buffer */
		char *pin = (char *)pfb_pinned(stage_bytes + b_tt + 1024);
		std::vector<unsigned> ttv;
		if (!pin) ttv.resize(std::max<size_t>(total, 1));
		{
			char *q = pin;
			for (Arr &a : arr) if (!a.direct) { a.stage = pin ? q : NULL; if (pin) q += (a.bytes + 255) & ~(size_t)255; }
		}
		unsigned *h_tt = pin ? (unsigned *)(pin + stage_bytes) : ttv.data();
		/* the serial-number terms come first: the running remainder below is sequential by definition and runs on
		 * a helper thread while the traces and the occupancy cross PCIe */
		bad = pfb_h2d(d_len, tptr.data(), sizeof(int) * ((size_t)n + 1)) || pfb_zero(r->d_wl, sizeof(unsigned long long) * 2)
				|| pfb_launch_build_traces(r->pool[r->cur], r->loc, n, NULL, d_len, d_tn, d_ts, r->d_wl, d_tt, r->ptc, p->nx)
				|| pfb_d2h(h_tt, d_tt, b_tt);
		std::thread chain;
		if (!bad) {
			const unsigned *tt = h_tt;
			chain = std::thread([tt, total, &serial_num]() {
				/* get_serial_num, route_common.c:224-254: serial = (serial + a - b - c) % 2000000000 per trace element, in
				 * the reference's wrapping int arithmetic.  |x| < 2^31 < 2 * 2000000000, so C's truncating remainder is
				 * one conditional add or subtract. */
				const int M = 2000000000;
				int sv = 0;
				for (size_t k = 0; k < total; k++) {
					int x = (int)((unsigned)sv + tt[k]);
					sv = x >= M ? x - M : (x <= -M ? x + M : x);
				}
				serial_num = sv;
			});
			t_1 = now_s();
			bad = pfb_launch_extract_occ(r->nodes, r->N, d_occ) != 0;
			for (Arr &a : arr) if (!bad) bad = pfb_d2h_async(a.direct || !a.stage ? *a.out : (void *)a.stage, a.dev, a.bytes) != 0;
			if (!bad) bad = pfb_d2h(h_wl, r->d_wl, sizeof(h_wl)) != 0;
			for (Arr &a : arr)
				if (!bad && !a.direct && a.stage) {
					char *dst = (char *)*a.out; const char *src = a.stage;
					parallel_for((long long)a.bytes, [&](long long lo, long long hi) { memcpy(dst + lo, src + lo, (size_t)(hi - lo)); });
				}
			t_2 = now_s();
			chain.join();
		}
		r->d2h_bytes += (int64_t)(b_tn + b_ts + b_tt + b_occ) + (int64_t)sizeof(int) * n + (int64_t)sizeof(float) * r->T;
	}
	pfb_free(d_len); pfb_free(d_tn); pfb_free(d_ts); pfb_free(d_tt);
	pfb_free(d_occ);
	if (bad) { pf_result_free(out); if (!host_ok) FAILF(PF_ENOMEM, "out of host memory"); CUDA_FAIL(); }
	memcpy(out->trace_ptr, tptr.data(), sizeof(int32_t) * ((size_t)n + 1));
	out->num_terminals = r->T;
	out->num_nodes = r->N;
	out->total_wirelength = (int32_t)h_wl[0];
	out->serial_num = serial_num;
	if (r->cfg.verbose) fprintf(stderr, "pf_router: result %.3f s (trace build + serial terms to host %.3f s, traces + occupancy to host %.3f s, wait for serial number %.3f s)\n",
			now_s() - t_0, t_1 - t_0, t_2 - t_1, now_s() - t_2);
	return PF_OK;
}

/* try_timing_driven_route, reference route_timing.c:85-343, on an existing router: iterate until the routing is
 * legal or the iteration budget is spent.  ONE host-device synchronisation per PathFinder iteration (the read of the
 * control block behind the cost update); with several ranks (pf_comm_init) the occupancy exchange after each route
 * part is device-side as well.  The analysis between iterations is either the host callback `sta` (the reference's own
 * STA in the drop-in flow) or the device analysis `dsta`, which reads the router's delay vector and writes its
 * criticality vector in place. */
extern "C" int pf_route_run(pf_router *r, pf_sta *dsta, pf_sta_fn sta, void *user, pf_iter_stats *stats_out, int stats_cap,
		int *iterations_out, int *success_out) {
	if (!r) FAILF(PF_EINVAL, "null router");
	const pf_problem *p = r->prob;
	const pf_config *cfg = &r->cfg;
	const pf_router_opts &o = p->opts;
	if (cfg->nranks > 1 && !r->comm_ready) FAILF(PF_EINVAL, "pf_route_run on %d ranks needs pf_comm_init first", cfg->nranks);
	const int max_iters = o.max_router_iterations;
	const int nparts = cfg->nranks > 1 ? 2 : 1;
	std::vector<float> crit, delay;
	r->crit_hist.clear();
	if (sta && o.timing_analysis_enabled) {
		crit.assign((size_t)std::max(p->num_terminals, 1), 0.f); delay.assign((size_t)std::max(p->num_terminals, 1), 0.f);
		for (int i = 0; i < p->num_nets; i++)
			if (!p->net_is_global[i]) for (int k = p->net_ptr[i] + 1; k < p->net_ptr[i + 1]; k++) crit[k] = 1.f;
	}
	float pres_fac = o.first_iter_pres_fac;
	int success = 0, itry, rc = PF_OK, nstats = 0;
	bool have_crit = false, polished = false;
	const bool breadth_first = o.router_algorithm == 1;      /* try_breadth_first_route, route_breadth_first.c:23-91 */
	/* PF_PHASES=1: device time of every phase of every iteration (events on the router's stream), printed per rank at the end */
	const bool phases = getenv("PF_PHASES") != NULL;
	std::vector<int> phase_rows;      /* marks per iteration */
#define PHASE_MARK() do { if (phases) pfb_mark(); } while (0)
	for (itry = 1; itry <= max_iters; itry++) {
		pf_iter_stats st;
		memset(&st, 0, sizeof(st));
		PHASE_MARK();
		if (!crit.empty()) r->crit_hist.insert(r->crit_hist.end(), crit.begin(), crit.begin() + p->num_terminals);
		if ((rc = pf_iteration_begin(r, have_crit ? crit.data() : NULL)) != PF_OK) break;
		PHASE_MARK();
		for (int part = 0; part < nparts && rc == PF_OK; part++) {
			rc = launch_routes(r, pres_fac, part, nparts);
			PHASE_MARK();
			if (rc == PF_OK && cfg->nranks > 1) { rc = pf_comm_exchange(r); PHASE_MARK(); }
		}
		if (rc != PF_OK) break;
		if ((rc = pf_reserve_opins(r, pres_fac, itry != 1)) != PF_OK) break;
		float acc_fac;
		st.pres_fac = pres_fac;
		if (itry == 1) { pres_fac = o.initial_pres_fac; acc_fac = breadth_first ? o.acc_fac : 0.f; }   /* route_breadth_first.c:84 */
		else {
			pres_fac *= o.pres_fac_mult;
			pres_fac = fminf(pres_fac, (float)(PF_HUGE_POSITIVE_FLOAT / 1e5));
			acc_fac = o.acc_fac;
		}
		/* feasibility is decided on the occupancies before the cost update; when nothing is overused the update is a
		 * no-op, so one fused pass serves both — and carries the wirelength of the first-iteration abort test */
		if ((rc = update_costs_async(r, acc_fac)) != PF_OK) break;
		const bool analyse = o.timing_analysis_enabled && dsta;
		if (analyse) {
			/* enqueued before the control block is read: the analysis of the final routing comes for free */
			if (cfg->nranks > 1 && (rc = pf_comm_gather_delays(r)) != PF_OK) break;
			if ((rc = pf_sta_analyze_device(dsta, r->net_delay, r->crit, NULL)) != PF_OK) break;
		}
		PHASE_MARK();
		int overused = 0;
		if ((rc = update_costs_finish(r, &overused)) != PF_OK) break;
		if (phases) phase_rows.push_back(nparts);
		st.overused_nodes = overused;
		st.nets_routed = (int)r->h_stats.nets; st.heap_pushes = (int64_t)r->h_stats.pushes; st.heap_pops = (int64_t)r->h_stats.pops;
		st.edge_visits = (int64_t)r->h_stats.visits;
		if (analyse) { float cpd = 0.f; if ((rc = pf_sta_read_cpd(dsta, &cpd)) != PF_OK) break; st.crit_path_delay = cpd; }
		if (cfg->verbose) fprintf(stderr, "pf_router: iteration %d: %d nets routed (%llu lost races), %d rr nodes overused, pres_fac %g; %llu labels settled, largest net %llu, %llu edge visits, %llu refills, %llu stale pops, %llu pushes\n", itry, st.nets_routed,
				(unsigned long long)r->h_stats.races, overused, (double)st.pres_fac, (unsigned long long)r->h_stats.pops, (unsigned long long)r->h_stats.max_net_pops, (unsigned long long)r->h_stats.visits,
				(unsigned long long)r->h_stats.refills, (unsigned long long)r->h_stats.stale, (unsigned long long)r->h_stats.pushes);
		if (itry == 1 && !breadth_first && r->avail_wl > 0
				&& (float)r->h_wl_used / (float)r->avail_wl > PF_FIRST_ITER_WIRELENGTH_LIMIT) {   /* route_timing.c:189-225 */
			if (stats_out && nstats < stats_cap) stats_out[nstats] = st;
			nstats++; itry++;
			break;
		}
		if (stats_out && nstats < stats_cap) stats_out[nstats] = st;
		nstats++;
		if (overused == 0) {
			/* polish (off by default): the routing is legal, but under the congested-only policy a net routed around
			 * congestion that has since dissolved keeps its detour.  One more iteration re-routes every net against the
			 * final picture, then the negotiation carries on until the routing is legal again. */
			if (!polished && cfg->polish > 0 && itry < max_iters && itry > 1) { polished = true; r->force_all_once = true; }
			else { success = 1; itry++; break; }
		}
		if (o.timing_analysis_enabled && dsta) {
			have_crit = false;   /* already on the device */
		} else if (o.timing_analysis_enabled && sta) {
			if (cfg->nranks > 1 && (rc = pf_comm_gather_delays(r)) != PF_OK) break;
			if ((rc = pf_get_net_delay(r, delay.data())) != PF_OK) break;
			float cpd = 0.f;
			sta(user, itry, delay.data(), crit.data(), &cpd);
			if (stats_out && nstats - 1 < stats_cap) stats_out[nstats - 1].crit_path_delay = cpd;
			have_crit = true;
		} else if (!o.timing_analysis_enabled) {
			have_crit = false;   /* criticalities stay 0 on the device */
		}
	}
	itry--;
	if (phases) {
		/* per iteration: begin | route part k [| exchange k] ... | opins + cost update + select (+ analysis); the gaps between
		 * iterations (host reads the control block, decides, launches) show up in the first column of the next iteration */
		std::vector<double> ms(8192);
		const int n = pfb_marks_read(ms.data(), (int)ms.size());
		const int per = 1 + nparts * (cfg->nranks > 1 ? 2 : 1) + 1;      /* intervals per iteration, the first being the gap + begin */
		std::string out;                             /* one write per rank: several ranks share the terminal */
		char b[256];
		snprintf(b, sizeof(b), "PF_PHASES rank %d/%d: %d iterations; columns: [gap to previous iteration's last launch incl. host read] begin", cfg->rank, cfg->nranks, (int)phase_rows.size());
		out += b;
		for (int k = 0; k < nparts; k++) { snprintf(b, sizeof(b), cfg->nranks > 1 ? " route%d exchange%d" : " route%d", k, k); out += b; }
		out += " update(ms)\n";
		int at = 0;
		for (size_t it = 0; it < phase_rows.size(); it++) {
			snprintf(b, sizeof(b), "PF_PHASES rank %d it %2d:", cfg->rank, (int)it + 1); out += b;
			if (it > 0 && at < n) { snprintf(b, sizeof(b), " [%.3f]", ms[at++]); out += b; }
			for (int k = 0; k < per && at < n; k++) { snprintf(b, sizeof(b), " %.3f", ms[at++]); out += b; }
			out += "\n";
		}
		fputs(out.c_str(), stderr);
	}
	if (rc != PF_OK && cfg->nranks > 1) pf_comm_abort(r);      /* peers waiting in an exchange see it instead of timing out */
	if (rc != PF_OK) return rc;
	if (cfg->nranks > 1 && (rc = pf_comm_gather_delays(r)) != PF_OK) return rc;   /* the result carries every net's delays on every rank */
	if (iterations_out) *iterations_out = itry;
	if (success_out) *success_out = success;
	return PF_OK;
}

static int route_loop(const pf_problem *p, const pf_config *cfg, pf_sta_fn sta, void *user, pf_sta *dsta, pf_result *out) {
	if (!p || !cfg || !out) FAILF(PF_EINVAL, "null argument");
	if (cfg->nranks > 1) FAILF(PF_EINVAL, "pf_try_* route on one GPU; for several ranks create the router, call pf_comm_init and pf_route_run");
	pf_router *r = NULL;
	int rc = pf_router_create(p, cfg, &r);
	if (rc != PF_OK) return rc;
	const int cap = std::max(p->opts.max_router_iterations, 1);
	std::vector<pf_iter_stats> stats((size_t)cap);
	int iterations = 0, success = 0;
	rc = pf_route_run(r, dsta, sta, user, stats.data(), cap, &iterations, &success);
	if (rc == PF_OK) {
		rc = pf_get_result(r, out);
		if (rc == PF_OK) {
			const int ns = std::min(iterations, cap);
			out->success = success;
			out->iterations = iterations;
			out->num_iter_stats = ns;
			out->iter_stats = (pf_iter_stats *)malloc(sizeof(pf_iter_stats) * (size_t)std::max(ns, 1));
			memcpy(out->iter_stats, stats.data(), sizeof(pf_iter_stats) * (size_t)ns);
			/* the criticalities each iteration used: known on the host only when the host analysis ran (with the device
			 * analysis they never leave the GPU) */
			const std::vector<float> &ch = r->crit_hist;
			out->num_crit_iters = (p->num_terminals && !ch.empty()) ? (int)(ch.size() / (size_t)p->num_terminals) : 0;
			out->iter_crit = (float *)malloc(sizeof(float) * std::max<size_t>(ch.size(), 1));
			memcpy(out->iter_crit, ch.data(), sizeof(float) * ch.size());
		}
	}
	pf_router_destroy(r);
	return rc;
}

extern "C" int pf_try_timing_driven_route(const pf_problem *p, const pf_config *cfg, pf_sta_fn sta, void *user, pf_result *out) {
	return route_loop(p, cfg, sta, user, NULL, out);
}

/* try_breadth_first_route (route_breadth_first.c:23): the same loop with opts.router_algorithm == 1 */
extern "C" int pf_try_breadth_first_route(const pf_problem *p, const pf_config *cfg, pf_result *out) {
	if (!p || !cfg || !out) FAILF(PF_EINVAL, "null argument");
	if (p->opts.router_algorithm != 1) FAILF(PF_EINVAL, "opts.router_algorithm must be 1 (breadth-first)");
	return route_loop(p, cfg, NULL, NULL, NULL, out);
}

extern "C" int pf_try_timing_driven_route_sta(const pf_problem *p, const pf_timing_graph *g, const pf_config *cfg, pf_result *out) {
	if (!p || !g || !cfg || !out) FAILF(PF_EINVAL, "null argument");
	pf_sta *s = NULL;
	int rc = pf_sta_create(g, p, cfg, &s);
	if (rc != PF_OK) return rc;
	rc = route_loop(p, cfg, NULL, NULL, s, out);
	pf_sta_destroy(s);
	return rc;
}
