/*
 * TEST INFRASTRUCTURE: pf_backend.h over the fiber warp emulator (pf_emu.h).  "Device memory" is
 * host memory, a kernel launch runs the same device functions as the CUDA build on emulated
 * warps.  Built only by tests/ into tests/emu/_build/libpf_router_emu.so.
 */
#include "pf_backend.h"
#include "pf_device.cuh"
#include "pf_sta_device.cuh"
#include "pf_gen_device.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static char g_err[256] = "";
static PfLaunchTimes g_times;

long long pf_emu_bucket_refills = 0, pf_emu_lazy_seedings = 0;
extern "C" long long pfb_emu_lazy_seedings(void) { return pf_emu_lazy_seedings; }
extern "C" long long pfb_emu_bucket_refills(void) { return pf_emu_bucket_refills; }
int pfb_init(int) { return 0; }
int pfb_device_count(void) { return 1; }
const char *pfb_name(void) { return "emu"; }
const char *pfb_last_error(void) { return g_err; }
void *pfb_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 16); }
void *pfb_alloc_raw(size_t bytes) { return malloc(bytes ? bytes : 16); }
void *pfb_host_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 16); }
void pfb_host_free(void *p) { free(p); }
void *pfb_pinned(size_t) { return NULL; }
void *pfb_pinned_upload(size_t) { return NULL; }
int pfb_h2d_async(void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); return 0; }
int pfb_d2h_async(void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); return 0; }
void pfb_free(void *p) { free(p); }
int pfb_h2d(void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); return 0; }
int pfb_d2h(void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); return 0; }
int pfb_d2d(void *d, const void *s, size_t n) { if (n) memmove(d, s, n); return 0; }
int pfb_fill(void *d, int b, size_t n) { if (n) memset(d, b, n); return 0; }
int pfb_zero(void *d, size_t n) { if (n) memset(d, 0, n); return 0; }
int pfb_sync(void) { return 0; }
void pfb_times(PfLaunchTimes *out, int reset) { if (out) *out = g_times; if (reset) memset(&g_times, 0, sizeof(g_times)); }
#include <time.h>
static double g_tt0;
static double now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
int pfb_timer_start(void) { g_tt0 = now_ms(); return 0; }
int pfb_timer_stop(double *ms) { *ms = now_ms() - g_tt0; return 0; }
void *pfb_stream(void) { return NULL; }
static double g_mark_t[4096]; static int g_nmarks = 0;
int pfb_mark(void) { if (g_nmarks < 4096) g_mark_t[g_nmarks++] = now_ms(); return 0; }
int pfb_marks_read(double *ms, int cap) { int n = 0; for (int i = 1; i < g_nmarks && n < cap; i++) ms[n++] = g_mark_t[i] - g_mark_t[i - 1]; g_nmarks = 0; return n; }
int pfb_num_sms(void) { const char *e = getenv("PF_EMU_SMS"); return e ? atoi(e) : 1; }

struct RouteArg { const PfParams *P; std::vector<unsigned char> *smem; };
static void route_warp(void *arg, int warp_id) {
	RouteArg *a = (RouteArg *)arg;
	/* the emulator gives every warp its own copy of the per-CTA tables */
	unsigned char *base = a->smem->data() + (size_t)warp_id * (PF_SMEM_BLOCK_TABLES + PF_SMEM_PER_WARP + PF_SMEM_HOT_ENTRIES * 8);
	PfIndexedDev *idx = (PfIndexedDev *)base;
	PfSwitchDev *sw = (PfSwitchDev *)(base + PF_MAX_INDEXED * sizeof(PfIndexedDev));
	if (pf_lane() == 0) {
		for (int i = 0; i < a->P->num_indexed; i++) idx[i] = a->P->indexed[i];
		for (int i = 0; i < a->P->num_sw; i++) sw[i] = a->P->sw[i];
	}
	pf_syncwarp();
	const bool rip = a->P->vq_ctl != NULL || a->P->committer != NULL;   /* the same choice of variant as the CUDA launcher */
	unsigned char *sm = base + PF_SMEM_BLOCK_TABLES;
	const int bk = a->P->far_buckets ? 1 : 0, mode = a->P->algorithm == 1 ? 2 : (a->P->max_batch == 1 ? 1 : 0);
#define PF_EMU_RUN(M, R, B) if (mode == M && (rip ? 1 : 0) == R && bk == B) pf_warp_main<M, R, B>(a->P, warp_id, idx, sw, sm);
	PF_EMU_RUN(0, 0, 0) PF_EMU_RUN(0, 0, 1) PF_EMU_RUN(0, 1, 0) PF_EMU_RUN(0, 1, 1)
	PF_EMU_RUN(1, 0, 0) PF_EMU_RUN(1, 0, 1) PF_EMU_RUN(1, 1, 0) PF_EMU_RUN(1, 1, 1)
	PF_EMU_RUN(2, 0, 0) PF_EMU_RUN(2, 0, 1) PF_EMU_RUN(2, 1, 0) PF_EMU_RUN(2, 1, 1)
#undef PF_EMU_RUN
}

int pfb_launch_route(const PfParams *P, int num_slots, int) {
	std::vector<unsigned char> smem((size_t)num_slots * (PF_SMEM_BLOCK_TABLES + PF_SMEM_PER_WARP + PF_SMEM_HOT_ENTRIES * 8));
	RouteArg a = { P, &smem };
	pf_emu_launch(route_warp, &a, num_slots);
	g_times.route_launches++;
	return 0;
}

int pfb_launch_update_cost(PfNode *nodes, int num_nodes, float acc_fac, int *d_overused, unsigned char *last_over, int iter_tag,
		unsigned long long *d_wl_used) {
	int over = 0;
	unsigned long long wl = 0;
	for (int i = 0; i < num_nodes; i++) { over += pf_update_cost_one(nodes, i, acc_fac, last_over, iter_tag); wl += pf_node_wirelength_in_use(&nodes[i]); }
	*d_overused += over;
	if (d_wl_used) *d_wl_used += wl;
	g_times.update_launches++;
	return 0;
}

int pfb_launch_apply_events(PfNode *nodes, const unsigned *events, long long count) {
	for (long long i = 0; i < count; i++) nodes[events[i] & ~PF_EVENT_DEC].occ += (events[i] & PF_EVENT_DEC) ? -1 : 1;
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_wirelength(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, unsigned long long *d_out) {
	unsigned long long acc = 0;
	for (int k = 0; k < num_all; k++) {
		const PfNetLoc l = loc[all_nets[k]];
		for (int i = 0; i < l.count; i++) acc += pf_tree_wirelength_one(&pool[l.off + i]);
	}
	*d_out += acc;
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_reserve_opins(PfNode *nodes, const uint32_t *edges, int node_bits, const PfIndexedDev *indexed, int num_groups,
		const int *group_source, const int *group_count, const int *group_off, int *chosen, int rip_up, float pres_fac) {
	for (int g = 0; g < num_groups; g++)
		pf_reserve_opins_group(nodes, edges, node_bits, indexed, group_source[g], group_count[g], chosen + group_off[g], rip_up, pres_fac);
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_select_nets(const PfNode *nodes, const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets,
		int num_all, const unsigned char *net_big, int force_all, int *list_small, int *list_big, int *counts,
		const unsigned char *last_over, int iter_tag, int window, const int *committer, int *scratch, int head_count,
		int *queued, int queued_tag, const int *pool_node, const unsigned char *over_now, int over_tag) {
	(void)scratch;
	counts[0] = counts[1] = counts[2] = counts[3] = 0;
	for (int k = 0; k < num_all; k++) {          /* all_nets is in fanout order; so are the lists */
		int net = all_nets[k];
		const int hit = force_all ? 1 : over_now ? pf_net_is_congested_fast(pool_node, loc[net], over_now, over_tag)
				: pf_net_is_congested(nodes, pool, loc[net], last_over, iter_tag, window, committer, net);
		if (hit) {
			if (queued) queued[net] = queued_tag;
			if (net_big[net]) { list_big[counts[1]++] = net; if (k < head_count) counts[3]++; }
			else { list_small[counts[0]++] = net; if (k < head_count) counts[2]++; }
		}
	}
	g_times.aux_launches++;
	return 0;
}

void pfb_bind_thread(void) {}
size_t pfb_select_scratch_bytes(int num_all) { (void)num_all; return 16; }

int pfb_launch_compact(const PfTreeNode *src, PfTreeNode *dst, const int *src_node, int *dst_node, PfNetLoc *loc, const int *all_nets, int num_all,
		unsigned long long *dst_head) {
	for (int k = 0; k < num_all; k++) {
		int net = all_nets[k];
		PfNetLoc l = loc[net];
		if (l.count == 0) continue;
		unsigned long long off = *dst_head;
		*dst_head += (unsigned long long)l.count;
		memcpy(dst + off, src + l.off, sizeof(PfTreeNode) * (size_t)l.count);
		memcpy(dst_node + off, src_node + l.off, sizeof(int) * (size_t)l.count);
		loc[net].off = (int)off;
	}
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_rebuild_owner(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, int *owner) {
	for (int k = 0; k < num_all; k++) {
		const PfNetLoc l = loc[all_nets[k]];
		for (int i = 0; i < l.count; i++) owner[pool[l.off + i].node] = all_nets[k];
	}
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_extract_occ(const PfNode *nodes, int num_nodes, int *occ_out) {
	for (int i = 0; i < num_nodes; i++) occ_out[i] = nodes[i].occ;
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_build_traces(const PfTreeNode *pool, const PfNetLoc *loc, int num_nets, int *len, const int *tptr,
		int *trace_node, short *trace_switch, unsigned long long *d_wl, unsigned *trace_term, const short *ptc, int nx) {
	for (int i = 0; i < num_nets; i++) {
		PfNetLoc l = loc[i];
		if (!trace_node) { len[i] = pf_trace_of_net(pool + l.off, l.count, NULL, NULL, NULL, NULL, 0, 0u); continue; }
		*d_wl += (unsigned long long)pf_trace_of_net(pool + l.off, l.count, trace_node + tptr[i], trace_switch + tptr[i],
				trace_term ? trace_term + tptr[i] : NULL, ptc, nx, (unsigned)(i + 1));
	}
	g_times.aux_launches++;
	return 0;
}


/* ---- rr graph built "on the device": the same closed-form functions, plain loops */
int pfb_gen_count(const PfGenDev *G, int *row, long long *num_edges) {
	long long acc = 0;
	for (int v = 0; v < G->num_nodes; v++) { const PfGenNode nd = pf_gen_decode(*G, v); const int d = pf_gen_node_edges(*G, v, nd, NULL); row[v] = (int)acc; acc += d; }
	row[G->num_nodes] = (int)acc;
	*num_edges = acc;
	g_times.aux_launches++;
	return 0;
}
static long long g_fill_wl = 0; static int g_fill_pending = 0;
int pfb_gen_fill(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc, long long *avail_wl);
int pfb_gen_fill_begin(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc) {
	g_fill_pending = 1;
	return pfb_gen_fill(G, row, nodes, edges, ptc, &g_fill_wl);
}
int pfb_gen_fill_end(long long *avail_wl) {
	if (!g_fill_pending) return 0;
	g_fill_pending = 0;
	if (avail_wl) *avail_wl = g_fill_wl;
	return 0;
}
int pfb_gen_fill(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc, long long *avail_wl) {
	long long wl = 0;
	for (int v = 0; v < G->num_nodes; v++) {
		const PfGenNode nd = pf_gen_decode(*G, v);
		const int deg = pf_gen_node_edges(*G, v, nd, edges + row[v]);
		pf_gen_write_node(nd, row[v], deg, &nodes[v], &ptc[v]);
		if (nd.type == 4 || nd.type == 5) wl += 1 + nd.x1 - nd.x0 + nd.y1 - nd.y0;
	}
	*avail_wl = wl;
	g_times.aux_launches++;
	return 0;
}
int pfb_reset_nodes(PfNode *nodes, int num_nodes) { for (int i = 0; i < num_nodes; i++) { nodes[i].occ = 0; nodes[i].acc_cost = 1.f; } return 0; }
static unsigned long long emu_mix64(unsigned long long x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
int pfb_graph_hash(const PfNode *nodes, int num_nodes, const uint32_t *edges, long long num_edges, const short *ptc, unsigned long long out[3]) {
	unsigned long long h0 = 0, h1 = 0, h2 = 0;
	for (long long i = 0; i < num_nodes; i++) {
		unsigned long long w[4]; memcpy(w, &nodes[i], 32);
		h0 += emu_mix64(w[0] ^ emu_mix64((unsigned long long)i)) + emu_mix64(w[1] + 0x9e3779b97f4a7c15ull * (unsigned long long)i) + emu_mix64(w[2] ^ (unsigned long long)(3 * i + 1)) + emu_mix64(w[3] ^ (unsigned long long)(5 * i + 2));
		h2 += emu_mix64(((unsigned long long)(unsigned short)ptc[i] << 32) ^ (unsigned long long)i);
	}
	for (long long i = 0; i < num_edges; i++) h1 += emu_mix64(((unsigned long long)edges[i] << 32) ^ (unsigned long long)i);
	out[0] = h0; out[1] = h1; out[2] = h2;
	return 0;
}

/* ---- multi-rank exchange: the "peer memory" of the emulator is POSIX shared memory, the ranks are processes; the protocol
 * (payload, fence, release of the sequence number; acquire-polling consumers) is the one of pf_kernels.cu */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <map>
#include <string>
struct ShmRegion { std::string name; size_t bytes; bool owner; };
static std::map<void *, ShmRegion> g_shm;

/* regions kept by the transport cache of pf_router.cpp outlive their routers: unlink what this process still owns at exit */
static void emu_unlink_all(void) { for (auto &kv : g_shm) if (kv.second.owner) shm_unlink(kv.second.name.c_str()); }
void *pfb_ipc_alloc(size_t bytes, void *handle64) {
	static int counter = 0;
	char name[64];
	snprintf(name, sizeof(name), "/pf_emu_%d_%d", (int)getpid(), counter++);
	int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
	if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { snprintf(g_err, sizeof(g_err), "shm_open(%s) failed", name); if (fd >= 0) close(fd); return NULL; }
	void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (p == MAP_FAILED) { shm_unlink(name); snprintf(g_err, sizeof(g_err), "mmap of %s failed", name); return NULL; }
	memset(handle64, 0, 64);
	snprintf((char *)handle64, 64, "%s", name);
	static bool hooked = false;
	if (!hooked) { hooked = true; atexit(emu_unlink_all); }
	g_shm[p] = ShmRegion{ name, bytes, true };
	return p;
}
void *pfb_ipc_open(const void *handle64) {
	char name[65];
	memcpy(name, handle64, 64); name[64] = 0;
	int fd = shm_open(name, O_RDWR, 0600);
	if (fd < 0) { snprintf(g_err, sizeof(g_err), "shm_open(%s) failed", name); return NULL; }
	struct stat st;
	if (fstat(fd, &st) != 0) { close(fd); return NULL; }
	void *p = mmap(NULL, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (p == MAP_FAILED) return NULL;
	g_shm[p] = ShmRegion{ name, (size_t)st.st_size, false };
	return p;
}
int pfb_ipc_clear_abort(void *region) { ((PfXchgHeader *)region)->abort_flag = 0u; return 0; }
void pfb_ipc_close(void *p) { auto it = g_shm.find(p); if (it != g_shm.end()) { munmap(p, it->second.bytes); g_shm.erase(it); } }
void pfb_ipc_free(void *p) { auto it = g_shm.find(p); if (it != g_shm.end()) { munmap(p, it->second.bytes); shm_unlink(it->second.name.c_str()); g_shm.erase(it); } }

static int emu_wait_flag(const unsigned *flag, unsigned want, const PfXchgHeader *peer, int *status, double timeout_s) {
	const double t0 = now_ms();
	while ((int)(__atomic_load_n(flag, __ATOMIC_ACQUIRE) - want) < 0) {
		if (__atomic_load_n(&peer->abort_flag, __ATOMIC_RELAXED)) { *status |= PF_ST_COMM_ABORT; return 0; }
		if (now_ms() - t0 > timeout_s * 1e3) { *status |= PF_ST_COMM_TIMEOUT; return 0; }
		usleep(50);
	}
	return 1;
}

int pfb_launch_xchg_events(PfNode *nodes, const PfPeers *peers, int me, int nranks, unsigned seq, const unsigned long long *event_head,
		long long event_cap, int *status, double timeout_s) {
	const int buf = (int)((seq - 1u) & 1u);
	PfXchgHeader *mine = (PfXchgHeader *)peers->base[me];
	unsigned long long c = *event_head;
	if ((long long)c > event_cap) c = (unsigned long long)event_cap;
	mine->count[buf] = c;
	__atomic_store_n(&mine->seq[buf], seq, __ATOMIC_RELEASE);
	for (int d = 1; d < nranks; d++) {
		const int k = (me + d) % nranks;
		const PfXchgHeader *ph = (const PfXchgHeader *)peers->base[k];
		if (!emu_wait_flag(&ph->seq[buf], seq, ph, status, timeout_s)) return 0;
		const long long cnt = (long long)ph->count[buf];
		const volatile unsigned *log = (const volatile unsigned *)(peers->base[k] + PF_XCHG_HEADER_BYTES) + (size_t)buf * (size_t)event_cap;
		for (long long i = 0; i < cnt; i++) { const unsigned e = log[i]; nodes[e & ~PF_EVENT_DEC].occ += (e & PF_EVENT_DEC) ? -1 : 1; }
	}
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_xchg_delays(float *net_delay, const unsigned char *term_owner, int num_terminals, const PfPeers *peers, int me, int nranks,
		unsigned dseq, long long event_cap, int *status, double timeout_s) {
	const int buf = (int)((dseq - 1u) & 1u);
	const size_t off = PF_XCHG_HEADER_BYTES + 8 * (size_t)event_cap + sizeof(float) * (size_t)buf * (size_t)num_terminals;
	float *pub = (float *)(peers->base[me] + off);
	for (int t = 0; t < num_terminals; t++) if (term_owner[t] == me) pub[t] = net_delay[t];
	PfXchgHeader *mine = (PfXchgHeader *)peers->base[me];
	__atomic_store_n(&mine->dseq[buf], dseq, __ATOMIC_RELEASE);
	for (int d = 1; d < nranks; d++) {
		const PfXchgHeader *ph = (const PfXchgHeader *)peers->base[(me + d) % nranks];
		if (!emu_wait_flag(&ph->dseq[buf], dseq, ph, status, timeout_s)) return 0;
	}
	for (int t = 0; t < num_terminals; t++) {
		const int k = term_owner[t];
		if (k != me && k < nranks) net_delay[t] = *(const volatile float *)((const float *)(peers->base[k] + off) + t);
	}
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_xchg_abort(const PfPeers *peers, int me) {
	__atomic_store_n(&((PfXchgHeader *)peers->base[me])->abort_flag, 1u, __ATOMIC_RELEASE);
	return 0;
}

/* ---- static timing analysis: the same per-element bodies, run serially */
int pfb_sta_load(const PfStaDev *S, const float *net_delay) {
	for (int t = 0; t < S->num_terminals; t++) pf_sta_load_delay(*S, t, net_delay);
	g_times.aux_launches++;
	return 0;
}
int pfb_sta_begin_pair(const PfStaDev *S, float *stat) {
	stat[0] = (float)PF_STA_HUGE_NEG; stat[1] = (float)PF_STA_HUGE_NEG; stat[2] = (float)PF_STA_HUGE_POS;
	for (int n = 0; n < S->num_tnodes; n++) pf_sta_reset_node(*S, n);
	g_times.aux_launches++;
	return 0;
}
int pfb_sta_sweep(const PfStaDev *S, int forward, int lv_begin, int lv_end, int spread, int domain, float constraint, float *stat) {
	(void)spread;
	for (int s = 0; s < lv_end - lv_begin; s++) {
		const int lv = forward ? lv_begin + s : lv_end - 1 - s;
		for (int k = S->level_ptr[lv]; k < S->level_ptr[lv + 1]; k++) {
			const int n = S->level_nodes[k];
			if (forward) pf_sta_forward_node(*S, n, lv, domain, stat); else pf_sta_backward_node(*S, n, domain, constraint, stat);
		}
	}
	g_times.aux_launches++;
	return 0;
}
int pfb_sta_update(const PfStaDev *S, float constraint, const float *stat, float *crit) {
	for (int t = 0; t < S->num_terminals; t++) pf_sta_update_terminal(*S, t, constraint, stat, crit);
	g_times.aux_launches++;
	return 0;
}

int pfb_launch_check_route(const PfNode *nodes, const uint32_t *edges, int node_bits, int num_nodes, int num_nets, const int *net_ptr, const int *net_term,
		const unsigned char *net_is_global, const int *trace_ptr, const int *trace_node, const short *trace_switch, unsigned char *matched,
		int *occ2, const int *occ_reported, int *report, unsigned long long *wl_extra) {
	for (int i = 0; i < num_nets; i++) {
		if (net_is_global[i]) continue;
		const int t0 = net_ptr[i], ns = net_ptr[i + 1] - t0 - 1;
		unsigned w = 0;
		const int code = pf_check_net(nodes, edges, node_bits, num_nodes, net_term + t0, ns, trace_node + trace_ptr[i], trace_switch + trace_ptr[i],
				trace_ptr[i + 1] - trace_ptr[i], matched + t0, occ2, &w);
		if (code) { report[0]++; if (i < report[1]) { report[1] = i; report[2] = code; } }
		else wl_extra[0] += w;
	}
	for (int v = 0; v < num_nodes; v++) {
		const int d = occ_reported[v] - occ2[v];
		if (d < 0 || (d > 0 && (nodes[v].type_ci & 7) != 3)) report[3]++; else wl_extra[1] += (unsigned long long)d;
		if (occ_reported[v] > (int)nodes[v].capacity) report[4]++;
	}
	g_times.aux_launches++;
	return 0;
}
