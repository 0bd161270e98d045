/*
 * pf_kernels.cu — CUDA backend (sm_100a) of the PathFinder router: the __global__ entry points
 * around the device code in pf_device.cuh, launch wrappers, device memory and event timing.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false (see build.py).
 *
 * Kernels
 *   pf_route_kernel        persistent warp-per-net router (the hot path; see pf_device.cuh)
 *   pf_update_cost_kernel  pathfinder_update_cost + feasible_routing, optionally fused with the
 *                          fold-in of the all-reduced occupancy delta (multi-GPU)
 *   pf_apply_events_kernel replay of another GPU's occupancy changes (multi-GPU sync)
 *   pf_wirelength_kernel   first-iteration wirelength sanity sum
 *   pf_reserve_opins_kernel locally-used OPIN reservation
 */
#include "pf_backend.h"
#include "pf_device.cuh"
#include "pf_sta_device.cuh"
#include "pf_gen_device.cuh"

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[512] = "";
static cudaStream_t g_stream = 0;
static int g_device = -1;
static int g_sms = 0;
static PfLaunchTimes g_times;

struct PendingEvent { cudaEvent_t a, b; int kind; int slots, variant; };   /* slots / variant: route launches, for PF_LAUNCH_LOG */
static PendingEvent g_pending[64];
static int g_npending = 0;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return -1; } } while (0)

const char *pfb_last_error(void) { return g_err; }
const char *pfb_name(void) { return "cuda:sm_100a"; }

int pfb_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "no CUDA device / driver"); return 0; }
	return n;
}

int pfb_init(int device) {
	int n = pfb_device_count();
	if (n <= 0) { snprintf(g_err, sizeof(g_err), "pf_router needs a CUDA device (sm_100a); none is visible — there is no CPU fallback"); return -1; }
	if (device < 0 || device >= n) { snprintf(g_err, sizeof(g_err), "device %d out of range (%d visible)", device, n); return -1; }
	CK(cudaSetDevice(device));
	if (g_device != device || !g_stream) {
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		if (prop.major < 10) { snprintf(g_err, sizeof(g_err), "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return -1; }
		g_sms = prop.multiProcessorCount;
		if (g_stream) cudaStreamDestroy(g_stream);
		CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
		{
			cudaMemPool_t pool;
			unsigned long long keep = ~0ull;
			CK(cudaDeviceGetDefaultMemPool(&pool, device));
			CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
		}
		g_device = device;
		memset(&g_times, 0, sizeof(g_times));
	}
	return 0;
}

int pfb_num_sms(void) { return g_sms; }
void pfb_bind_thread(void) { if (g_device >= 0) cudaSetDevice(g_device); }

/* Stream-ordered allocation from the device's default memory pool with an unlimited release threshold:
 * the first router pays for mapping its ~GBs of scratch, later create/destroy cycles (the binary search
 * over channel widths calls the router repeatedly, place_and_route.c:557,664) reuse the cached blocks. */
void *pfb_alloc_raw(size_t bytes) {
	void *p = NULL;
	if (bytes == 0) bytes = 16;
	if (cudaMallocAsync(&p, bytes, g_stream) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaMallocAsync(%zu) failed", bytes); cudaGetLastError(); return NULL; }
	return p;
}
void *pfb_alloc(size_t bytes) {
	void *p = pfb_alloc_raw(bytes);
	if (bytes == 0) bytes = 16;
	if (p && cudaMemsetAsync(p, 0, bytes, g_stream) != cudaSuccess) { cudaFreeAsync(p, g_stream); return NULL; }
	return p;
}
void pfb_free(void *p) { if (p) cudaFreeAsync(p, g_stream); }

/* process-wide pinned staging buffer for host<->device transfers of the big arrays */
/* Two process-wide pinned staging buffers.  [0] download: ordinary cached memory, the host copies results out
 * of it.  [1] upload: write-combined — the flattening threads only ever write it front to back, the stores
 * bypass the cache hierarchy (no read-for-ownership traffic) and the DMA engine does not have to snoop. */
static void *g_pinned[2] = { NULL, NULL }; static size_t g_pinned_bytes[2] = { 0, 0 };
static void *pinned_get(int which, size_t bytes) {
	if (bytes <= g_pinned_bytes[which]) return g_pinned[which];
	if (g_pinned[which]) { cudaStreamSynchronize(g_stream); cudaFreeHost(g_pinned[which]); g_pinned[which] = NULL; g_pinned_bytes[which] = 0; }
	size_t want = bytes + (bytes >> 3);
	if (cudaHostAlloc(&g_pinned[which], want, which ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) {
		cudaGetLastError(); g_pinned[which] = NULL; return NULL;
	}
	g_pinned_bytes[which] = want;
	return g_pinned[which];
}
void *pfb_host_alloc(size_t bytes) { void *p = NULL; if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return NULL; } return p; }
void pfb_host_free(void *p) { if (p) cudaFreeHost(p); }
void *pfb_pinned(size_t bytes) { return pinned_get(0, bytes); }
void *pfb_pinned_upload(size_t bytes) { return pinned_get(1, bytes); }
int pfb_h2d_async(void *dst, const void *src, size_t bytes) { if (!bytes) return 0; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream)); return 0; }
int pfb_d2h_async(void *dst, const void *src, size_t bytes) { if (!bytes) return 0; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream)); return 0; }
int pfb_h2d(void *dst, const void *src, size_t bytes) { if (!bytes) return 0; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream)); CK(cudaStreamSynchronize(g_stream)); return 0; }
int pfb_d2h(void *dst, const void *src, size_t bytes) { if (!bytes) return 0; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream)); CK(cudaStreamSynchronize(g_stream)); return 0; }
int pfb_d2d(void *dst, const void *src, size_t bytes) { if (!bytes) return 0; CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream)); return 0; }
int pfb_fill(void *dst, int byte, size_t bytes) { if (!bytes) return 0; CK(cudaMemsetAsync(dst, byte, bytes, g_stream)); return 0; }
int pfb_zero(void *dst, size_t bytes) { if (!bytes) return 0; CK(cudaMemsetAsync(dst, 0, bytes, g_stream)); return 0; }

static int drain_events(void) {
	for (int i = 0; i < g_npending; i++) {
		float ms = 0.f;
		CK(cudaEventSynchronize(g_pending[i].b));
		CK(cudaEventElapsedTime(&ms, g_pending[i].a, g_pending[i].b));
		if (g_pending[i].kind == 0) {
			g_times.route_ms += ms; g_times.route_launches++;
			static const int log_launches = getenv("PF_LAUNCH_LOG") != NULL;   /* diagnostics: one line per route launch */
			if (log_launches) fprintf(stderr, "pf_launch: route %10.3f ms, %5d warps, mode %d ripple %d buckets %d\n", ms, g_pending[i].slots,
					g_pending[i].variant >> 2, (g_pending[i].variant >> 1) & 1, g_pending[i].variant & 1);
		}
		else if (g_pending[i].kind == 1) { g_times.update_ms += ms; g_times.update_launches++; }
		else { g_times.aux_ms += ms; g_times.aux_launches++; }
		cudaEventDestroy(g_pending[i].a);
		cudaEventDestroy(g_pending[i].b);
	}
	g_npending = 0;
	return 0;
}

int pfb_sync(void) { CK(cudaStreamSynchronize(g_stream)); return drain_events(); }

void pfb_times(PfLaunchTimes *out, int reset) {
	drain_events();
	if (out) *out = g_times;
	if (reset) memset(&g_times, 0, sizeof(g_times));
}

/* whole-step timer: device time between two points of the router's stream */
static cudaEvent_t g_t0 = 0, g_t1 = 0;
int pfb_timer_start(void) {
	if (!g_t0) { CK(cudaEventCreate(&g_t0)); CK(cudaEventCreate(&g_t1)); }
	CK(cudaEventRecord(g_t0, g_stream));
	return 0;
}
int pfb_timer_stop(double *ms) {
	float f = 0.f;
	CK(cudaEventRecord(g_t1, g_stream));
	CK(cudaEventSynchronize(g_t1));
	CK(cudaEventElapsedTime(&f, g_t0, g_t1));
	*ms = f;
	return 0;
}
void *pfb_stream(void) { return (void *)g_stream; }

static cudaEvent_t g_marks[4096];
static int g_nmarks = 0;
int pfb_mark(void) {
	if (g_nmarks >= 4096) return 0;
	CK(cudaEventCreate(&g_marks[g_nmarks]));
	CK(cudaEventRecord(g_marks[g_nmarks], g_stream));
	g_nmarks++;
	return 0;
}
int pfb_marks_read(double *ms, int cap) {
	int n = 0;
	if (g_nmarks > 0) cudaEventSynchronize(g_marks[g_nmarks - 1]);
	for (int i = 1; i < g_nmarks; i++) {
		float f = 0.f;
		cudaEventElapsedTime(&f, g_marks[i - 1], g_marks[i]);
		if (n < cap) ms[n++] = f;
	}
	for (int i = 0; i < g_nmarks; i++) cudaEventDestroy(g_marks[i]);
	g_nmarks = 0;
	return n;
}

static int ev_begin(int kind) {
	if (g_npending == 64 && drain_events() != 0) return -1;
	PendingEvent *p = &g_pending[g_npending];
	p->kind = kind;
	CK(cudaEventCreate(&p->a));
	CK(cudaEventCreate(&p->b));
	CK(cudaEventRecord(p->a, g_stream));
	return 0;
}
static int ev_end(void) {
	CK(cudaEventRecord(g_pending[g_npending].b, g_stream));
	g_npending++;
	CK(cudaGetLastError());
	return 0;
}

/* ------------------------------------------------------------------ kernels */
extern __shared__ __align__(16) unsigned char pf_smem[];

/* STRICT = 1: strict best-first search (one label settled per step; P.max_batch == 1), the throughput mode;
 * STRICT = 0: a delta bucket of up to P.max_batch labels per step, the latency mode for few nets per warp;
 * STRICT = 2: the breadth-first router (one persistent wavefront per net, route_breadth_first.c) */
template <int STRICT, int RIP, int BK> __global__ void __launch_bounds__(128, BK ? 2 : 5) pf_route_kernel(const __grid_constant__ PfParams P, int num_slots) {
	/* regular slots (BK = 0): 96 registers and 43.3 KB of shared memory per 4-warp CTA: 5 CTAs = 20 warps per SM; the big
	 * slots run a few hundred one-warp CTAs on the whole device and take the registers they like.
	 * shared memory: [switch + cost-index tables, one copy per CTA][per-warp regions] */
	const int warp_in_block = (int)(threadIdx.x >> 5);
	const int slot = (int)blockIdx.x * (int)(blockDim.x >> 5) + warp_in_block;
	PfIndexedDev *idx = (PfIndexedDev *)pf_smem;
	PfSwitchDev *sw = (PfSwitchDev *)(pf_smem + PF_MAX_INDEXED * sizeof(PfIndexedDev));
	PF_COLD_LOOP for (int i = (int)threadIdx.x; i < P.num_indexed; i += (int)blockDim.x) idx[i] = P.indexed[i];
	PF_COLD_LOOP for (int i = (int)threadIdx.x; i < P.num_sw; i += (int)blockDim.x) sw[i] = P.sw[i];
	__syncthreads();
	if (slot >= num_slots) return;
	const size_t per_warp = PF_SMEM_PER_WARP + (P.hot ? 0 : (size_t)PF_SMEM_HOT_ENTRIES * 8);
	pf_warp_main<STRICT, RIP, BK>(&P, slot, idx, sw, pf_smem + PF_SMEM_BLOCK_TABLES + (size_t)warp_in_block * per_warp);
}

__global__ void pf_update_cost_kernel(PfNode *nodes, int num_nodes, float acc_fac, int *d_overused,
		unsigned char *last_over, int iter_tag, unsigned long long *d_wl_used) {
	int over = 0;
	unsigned wl = 0;
	for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < num_nodes; i += (int)(gridDim.x * blockDim.x)) {
		over += pf_update_cost_one(nodes, i, acc_fac, last_over, iter_tag);
		wl += pf_node_wirelength_in_use(&nodes[i]);
	}
	over = __reduce_add_sync(0xffffffffu, over);
	wl = __reduce_add_sync(0xffffffffu, wl);
	if ((threadIdx.x & 31u) == 0) {
		if (over) atomicAdd(d_overused, over);
		if (wl && d_wl_used) atomicAdd(d_wl_used, (unsigned long long)wl);
	}
}

/* another rank's event log: one atomic per event on the node records */
__global__ void pf_apply_events_kernel(PfNode *nodes, const unsigned *events, long long count) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
		const unsigned e = events[i];
		atomicAdd(&nodes[e & ~PF_EVENT_DEC].occ, (e & PF_EVENT_DEC) ? -1 : 1);
	}
}

__global__ void pf_build_traces_kernel(const PfTreeNode *pool, const PfNetLoc *loc, int num_nets, int *len, const int *tptr,
		int *trace_node, short *trace_switch, unsigned long long *d_wl, unsigned *trace_term, const short *ptc, int nx) {
	int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= num_nets) return;
	PfNetLoc l = loc[i];
	if (!trace_node) { len[i] = pf_trace_of_net(pool + l.off, l.count, NULL, NULL, NULL, NULL, 0, 0u); return; }
	int wl = pf_trace_of_net(pool + l.off, l.count, trace_node + tptr[i], trace_switch + tptr[i], trace_term ? trace_term + tptr[i] : NULL,
			ptc, nx, (unsigned)(i + 1));
	if (wl) atomicAdd(d_wl, (unsigned long long)wl);
}

__global__ void pf_extract_occ_kernel(const PfNode *nodes, int num_nodes, int *occ_out) {
	for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < num_nodes; i += (int)(gridDim.x * blockDim.x)) occ_out[i] = nodes[i].occ;
}

/* wirelength of the LIVE route trees of this rank's nets (the route store is an append-only log: entries of
 * re-routed nets stay behind until the next compaction, so the log itself cannot be summed) */
__global__ void pf_wirelength_kernel(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all,
		unsigned long long *d_out) {
	unsigned acc = 0;
	for (int k = (int)(blockIdx.x * blockDim.x + threadIdx.x); k < num_all; k += (int)(gridDim.x * blockDim.x)) {
		const PfNetLoc l = loc[all_nets[k]];
		for (int i = 0; i < l.count; i++) acc += pf_tree_wirelength_one(&pool[l.off + i]);
	}
	acc = __reduce_add_sync(0xffffffffu, acc);
	if ((threadIdx.x & 31u) == 0 && acc) atomicAdd(d_out, (unsigned long long)acc);
}

__global__ void pf_reserve_opins_kernel(PfNode *nodes, const uint32_t *edges, int node_bits, const PfIndexedDev *indexed,
		int num_groups, const int *group_source, const int *group_count, const int *group_off,
		int *chosen, int rip_up, float pres_fac) {
	int g = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (g < num_groups)
		pf_reserve_opins_group(nodes, edges, node_bits, indexed, group_source[g], group_count[g], chosen + group_off[g], rip_up, pres_fac);
}

/* Congested-net selection, order-preserving: pass 1 flags the nets (all_nets is in the reference's
 * decreasing-fanout order) and counts per CTA, pass 2 turns the counts into offsets and scatters, so the two
 * work lists come out in that same order without a host-side sort. */
#define PF_SEL_BLOCK 256
__global__ void __launch_bounds__(PF_SEL_BLOCK) pf_select_flag_kernel(const PfNode *nodes, const PfTreeNode *pool, const PfNetLoc *loc,
		const int *all_nets, int num_all, const unsigned char *net_big, int force_all, unsigned char *flag, int *block_counts,
		const unsigned char *last_over, int iter_tag, int window, const int *committer, int head_count, int *counts,
		int *queued, int queued_tag, const int *pool_node, const unsigned char *over_now, int over_tag) {
	int k = (int)(blockIdx.x * PF_SEL_BLOCK + threadIdx.x);
	int f = 0;
	if (k < num_all) {
		int net = all_nets[k];
		const int hit = force_all ? 1 : over_now ? pf_net_is_congested_fast(pool_node, loc[net], over_now, over_tag)
				: pf_net_is_congested(nodes, pool, loc[net], last_over, iter_tag, window, committer, net);
		if (hit) f = net_big[net] ? 2 : 1;
		flag[k] = (unsigned char)f;
		if (f && queued) queued[net] = queued_tag;
	}
	int cs = __syncthreads_count(f == 1), cb = __syncthreads_count(f == 2);
	int hs = __syncthreads_count(f == 1 && k < head_count), hb = __syncthreads_count(f == 2 && k < head_count);
	if (threadIdx.x == 0) {
		block_counts[2 * blockIdx.x] = cs; block_counts[2 * blockIdx.x + 1] = cb;
		if (hs) atomicAdd(&counts[2], hs);
		if (hb) atomicAdd(&counts[3], hb);
	}
}

__global__ void __launch_bounds__(PF_SEL_BLOCK) pf_select_scatter_kernel(const int *all_nets, int num_all, const unsigned char *flag,
		const int *block_counts, int *list_small, int *list_big, int *counts) {
	__shared__ int s_red[2][PF_SEL_BLOCK / 32];
	__shared__ int s_warp[2][PF_SEL_BLOCK / 32];
	const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
	/* offset of this CTA = counts of all earlier CTAs */
	int ps = 0, pb = 0;
	for (int b = (int)threadIdx.x; b < (int)blockIdx.x; b += PF_SEL_BLOCK) { ps += block_counts[2 * b]; pb += block_counts[2 * b + 1]; }
	for (int o = 16; o > 0; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); pb += __shfl_xor_sync(0xffffffffu, pb, o); }
	int k = (int)(blockIdx.x * PF_SEL_BLOCK + threadIdx.x);
	int f = k < num_all ? flag[k] : 0;
	unsigned ms = __ballot_sync(0xffffffffu, f == 1), mb = __ballot_sync(0xffffffffu, f == 2);
	if (lane == 0) { s_red[0][warp] = ps; s_red[1][warp] = pb; s_warp[0][warp] = __popc(ms); s_warp[1][warp] = __popc(mb); }
	__syncthreads();
	int base_s = 0, base_b = 0, tot_s = 0, tot_b = 0;
	for (int w = 0; w < PF_SEL_BLOCK / 32; w++) {
		base_s += s_red[0][w]; base_b += s_red[1][w];
		if (w < warp) { base_s += s_warp[0][w]; base_b += s_warp[1][w]; }
		tot_s += s_warp[0][w]; tot_b += s_warp[1][w];
	}
	const unsigned below = (1u << lane) - 1u;
	if (f == 1) list_small[base_s + __popc(ms & below)] = all_nets[k];
	else if (f == 2) list_big[base_b + __popc(mb & below)] = all_nets[k];
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
		int all_s = tot_s, all_b = tot_b;
		for (int w = 0; w < PF_SEL_BLOCK / 32; w++) { all_s += s_red[0][w]; all_b += s_red[1][w]; }
		counts[0] = all_s; counts[1] = all_b;
	}
}

/* one warp per net: reserve space in the destination log, copy the tree with coalesced 32-byte entries */
__global__ void pf_compact_kernel(const PfTreeNode *src, PfTreeNode *dst, const int *src_node, int *dst_node, PfNetLoc *loc, const int *all_nets, int num_all,
		unsigned long long *dst_head) {
	int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31u);
	int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
	for (int k = warp; k < num_all; k += nwarps) {
		int net = all_nets[k];
		PfNetLoc l = loc[net];
		if (l.count == 0) continue;
		unsigned long long off = 0;
		if (lane == 0) off = atomicAdd(dst_head, (unsigned long long)l.count);
		off = __shfl_sync(0xffffffffu, off, 0);
		for (int i = lane; i < l.count; i += 32) { dst[off + i] = src[l.off + i]; dst_node[off + i] = src_node[l.off + i]; }
		if (lane == 0) loc[net].off = (int)off;
	}
}

/* one warp per net: owner[node] = net for every entry of the net's tree */
__global__ void pf_rebuild_owner_kernel(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, int *owner) {
	int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31u);
	int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
	for (int k = warp; k < num_all; k += nwarps) {
		const int net = all_nets[k];
		const PfNetLoc l = loc[net];
		for (int i = lane; i < l.count; i += 32) owner[pool[l.off + i].node] = net;
	}
}

/* ------------------------------------------------------------------ static timing analysis (pf_sta_device.cuh) */
__global__ void pf_sta_load_kernel(PfStaDev S, const float *net_delay) {
	for (int t = (int)(blockIdx.x * blockDim.x + threadIdx.x); t < S.num_terminals; t += (int)(gridDim.x * blockDim.x)) pf_sta_load_delay(S, t, net_delay);
}

__global__ void pf_sta_begin_pair_kernel(PfStaDev S, float *stat) {
	const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i == 0) { stat[0] = (float)PF_STA_HUGE_NEG; stat[1] = (float)PF_STA_HUGE_NEG; stat[2] = (float)PF_STA_HUGE_POS; }
	for (int n = i; n < S.num_tnodes; n += (int)(gridDim.x * blockDim.x)) pf_sta_reset_node(S, n);
}

/* A run of consecutive levels in ONE CTA: timing graphs are hundreds of levels deep and most levels hold a few
 * hundred tnodes, so a launch per level would be all launch latency; __syncthreads() is the level barrier.
 * tnodes are numbered in level order (pf_sta_create), so thread t of level l owns tnode level_ptr[l] + t: the
 * structural part of its work for the NEXT level (CSR bounds, first in-edge record / first out-edge target) does not
 * depend on this level's results and is loaded before the barrier; after the barrier only the T_arr / T_req gather
 * is left on the critical path (≈ 2.7 µs -> ≈ 1 µs per level). */
#define PF_STA_CTA 1024
static __device__ __forceinline__ void sta_forward_pre(const PfStaDev &S, int n, int lo, int hi, int2 rec0, float *stat) {
	float best = (float)PF_STA_HUGE_NEG;
	int any = 0;
	for (int k = lo; k < hi; k++) {
		const int2 rec = (k == lo) ? rec0 : ((const int2 *)S.in_rec)[k];
		const float ta = S.T_arr[rec.x];
		if (ta < PF_STA_NEG_EPS) continue;
		const float cand = ta + S.Tdel[rec.y];
		if (cand > best) best = cand;
		any = 1;
	}
	if (any) { S.T_arr[n] = best; pf_atomic_max_f(&stat[0], best); }
}

__global__ void __launch_bounds__(PF_STA_CTA) pf_sta_levels_kernel(PfStaDev S, int forward, int lv_begin, int lv_end, int domain,
		float constraint, float *stat) {
	const int tid = (int)threadIdx.x;
	if (forward) {
		int k0 = S.level_ptr[lv_begin] + tid, kend = S.level_ptr[lv_begin + 1];
		int lo = 0, hi = 0;
		int2 rec0 = make_int2(0, 0);
		if (k0 < kend && lv_begin > 0) { lo = S.in_ptr[k0]; hi = S.in_ptr[k0 + 1]; if (hi > lo) rec0 = ((const int2 *)S.in_rec)[lo]; }
		for (int lv = lv_begin; lv < lv_end; lv++) {
			const int c_k0 = k0, c_kend = kend, c_lo = lo, c_hi = hi;
			const int2 c_rec0 = rec0;
			if (lv + 1 < lv_end) {                   /* structure of this thread's tnode in the next level */
				k0 = kend + tid; kend = S.level_ptr[lv + 2];
				if (k0 < kend) { lo = S.in_ptr[k0]; hi = S.in_ptr[k0 + 1]; if (hi > lo) rec0 = ((const int2 *)S.in_rec)[lo]; }
			}
			if (c_k0 < c_kend) {
				if (lv == 0) pf_sta_forward_node(S, c_k0, 0, domain, stat);
				else sta_forward_pre(S, c_k0, c_lo, c_hi, c_rec0, stat);
			}
			for (int k = c_k0 + PF_STA_CTA; k < c_kend; k += PF_STA_CTA) pf_sta_forward_node(S, k, lv, domain, stat);
			__syncthreads();
		}
	} else {
		/* (preloading the out-edge structure the same way was measured: no gain — the required-time gather dominates) */
		for (int lv = lv_end - 1; lv >= lv_begin; lv--) {
			for (int k = S.level_ptr[lv] + tid; k < S.level_ptr[lv + 1]; k += PF_STA_CTA) pf_sta_backward_node(S, k, domain, constraint, stat);
			__syncthreads();
		}
	}
}

/* one wide level over the whole GPU */
__global__ void pf_sta_level_kernel(PfStaDev S, int forward, int lv, int domain, float constraint, float *stat) {
	for (int k = S.level_ptr[lv] + (int)(blockIdx.x * blockDim.x + threadIdx.x); k < S.level_ptr[lv + 1]; k += (int)(gridDim.x * blockDim.x)) {
		const int n = k;                                 /* tnodes are numbered in level order */
		if (forward) pf_sta_forward_node(S, n, lv, domain, stat); else pf_sta_backward_node(S, n, domain, constraint, stat);
	}
}

__global__ void pf_sta_update_kernel(PfStaDev S, float constraint, const float *stat, float *crit) {
	for (int t = (int)(blockIdx.x * blockDim.x + threadIdx.x); t < S.num_terminals; t += (int)(gridDim.x * blockDim.x)) pf_sta_update_terminal(S, t, constraint, stat, crit);
}

/* ------------------------------------------------------------------ check_route (pf_check_net) */
__global__ void pf_check_nets_kernel(const PfNode *nodes, const uint32_t *edges, int node_bits, int num_nodes, int num_nets, const int *net_ptr,
		const int *net_term, const unsigned char *net_is_global, const int *trace_ptr, const int *trace_node, const short *trace_switch,
		unsigned char *matched, int *occ2, int *report /* [0] bad nets [1] first bad net [2] its code */, unsigned long long *wl) {
	const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= num_nets || net_is_global[i]) return;
	const int t0 = net_ptr[i], ns = net_ptr[i + 1] - t0 - 1;
	unsigned w = 0;
	const int code = pf_check_net(nodes, edges, node_bits, num_nodes, net_term + t0, ns, trace_node + trace_ptr[i], trace_switch + trace_ptr[i],
			trace_ptr[i + 1] - trace_ptr[i], matched + t0, occ2, &w);
	if (code) {
		atomicAdd(&report[0], 1);
		if (atomicMin(&report[1], i) > i) report[2] = code;          /* code of the lowest-numbered bad net (last writer wins among equals) */
	} else if (w) atomicAdd(wl, (unsigned long long)w);
}

/* occupancy recomputed from the traces against the reported one: the difference may only be the locally used
 * OPINs (reserve_locally_used_opins), i.e. non-negative and on OPIN nodes */
__global__ void pf_check_occ_kernel(const PfNode *nodes, int num_nodes, const int *occ2, const int *occ_reported, int *report /* [3] mismatches [4] overused */,
		unsigned long long *extra) {
	int mism = 0, over = 0;
	unsigned ex = 0;
	for (int v = (int)(blockIdx.x * blockDim.x + threadIdx.x); v < num_nodes; v += (int)(gridDim.x * blockDim.x)) {
		const int d = occ_reported[v] - occ2[v];
		const int ty = nodes[v].type_ci & 7;
		if (d < 0 || (d > 0 && ty != 3)) mism++;
		else ex += (unsigned)d;
		if (occ_reported[v] > (int)nodes[v].capacity) over++;
	}
	mism = __reduce_add_sync(0xffffffffu, mism); over = __reduce_add_sync(0xffffffffu, over); ex = __reduce_add_sync(0xffffffffu, ex);
	if ((threadIdx.x & 31u) == 0) {
		if (mism) atomicAdd(&report[3], mism);
		if (over) atomicAdd(&report[4], over);
		if (ex) atomicAdd(extra, (unsigned long long)ex);
	}
}


/* ------------------------------------------------------------------ multi-GPU exchange over peer memory (NVLink / NVSwitch)
 * One process per GPU; every rank's exchange region (PfXchgHeader + event logs + published delays) is mapped into the
 * others through CUDA IPC.  A rank publishes by writing the payload, a system-scope fence and a release store of the
 * sequence number; consumers poll the sequence number with acquire loads over NVLink and then read the payload straight
 * out of the producer's memory (volatile loads: nothing of a previous round can be served from this SM's L1).  The host
 * never waits: the kernels are stream-ordered behind the route kernels that produce the logs. */
static __device__ __forceinline__ unsigned pf_ld_acquire_sys(const unsigned *p) {
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
static __device__ __forceinline__ void pf_st_release_sys(unsigned *p, unsigned v) {
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ unsigned long long pf_globaltimer(void) {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
static __device__ __forceinline__ uint4 pf_ld_volatile_u4(const void *p) {
	uint4 v;
	asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}
/* thread 0 of a CTA waits until *flag has reached `want` (wrap-safe), the peer aborted, or the timeout ran out */
static __device__ int pf_wait_flag(const unsigned *flag, unsigned want, const PfXchgHeader *peer, int *status, unsigned long long timeout_ns) {
	const unsigned long long t0 = pf_globaltimer();
	while ((int)(pf_ld_acquire_sys(flag) - want) < 0) {
		if (*(const volatile unsigned *)&peer->abort_flag) { atomicOr(status, PF_ST_COMM_ABORT); return 0; }
		if (pf_globaltimer() - t0 > timeout_ns) { atomicOr(status, PF_ST_COMM_TIMEOUT); return 0; }
		__nanosleep(64);
	}
	return 1;
}

__global__ void __launch_bounds__(256) pf_xchg_events_kernel(PfNode *nodes, const __grid_constant__ PfPeers peers, int me, int nranks, unsigned seq,
		const unsigned long long *event_head, long long event_cap, int *status, unsigned long long timeout_ns) {
	const int buf = (int)((seq - 1u) & 1u);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		PfXchgHeader *mine = (PfXchgHeader *)peers.base[me];
		unsigned long long c = *event_head;
		if ((long long)c > event_cap) c = (unsigned long long)event_cap;      /* the overflow itself is reported by the host */
		mine->count[buf] = c;
		__threadfence_system();
		pf_st_release_sys(&mine->seq[buf], seq);
	}
	__shared__ unsigned long long s_count;
	__shared__ int s_ok;
	for (int d = 1; d < nranks; d++) {
		const int k = (me + d) % nranks;
		const PfXchgHeader *ph = (const PfXchgHeader *)peers.base[k];
		if (threadIdx.x == 0) {
			const int ok = pf_wait_flag(&ph->seq[buf], seq, ph, status, timeout_ns);
			s_ok = ok;
			s_count = ok ? *(const volatile unsigned long long *)&ph->count[buf] : 0ull;
		}
		__syncthreads();
		if (!s_ok) return;
		const long long cnt = (long long)s_count;
		const unsigned *log = (const unsigned *)(peers.base[k] + PF_XCHG_HEADER_BYTES) + (size_t)buf * (size_t)event_cap;
		for (long long i = 4ll * ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < cnt; i += 4ll * (long long)gridDim.x * blockDim.x) {
			const uint4 e = pf_ld_volatile_u4(log + i);
			const unsigned ev[4] = { e.x, e.y, e.z, e.w };
#pragma unroll
			for (int q = 0; q < 4; q++)
				if (i + q < cnt) atomicAdd(&nodes[ev[q] & ~PF_EVENT_DEC].occ, (ev[q] & PF_EVENT_DEC) ? -1 : 1);
		}
		__syncthreads();
	}
}

/* step 1 of the delay exchange: this rank's sink delays into its published buffer */
__global__ void pf_xchg_delays_publish_kernel(const float *net_delay, const unsigned char *term_owner, int num_terminals, float *pub, int me) {
	for (int t = (int)(blockIdx.x * blockDim.x + threadIdx.x); t < num_terminals; t += (int)(gridDim.x * blockDim.x))
		if (term_owner[t] == me) pub[t] = net_delay[t];
}
/* step 2: release the buffer, wait for the peers', take the delays of the nets they route */
__global__ void __launch_bounds__(256) pf_xchg_delays_gather_kernel(float *net_delay, const unsigned char *term_owner, int num_terminals,
		const __grid_constant__ PfPeers peers, int me, int nranks, unsigned dseq, long long event_cap, int *status, unsigned long long timeout_ns) {
	const int buf = (int)((dseq - 1u) & 1u);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		PfXchgHeader *mine = (PfXchgHeader *)peers.base[me];
		__threadfence_system();
		pf_st_release_sys(&mine->dseq[buf], dseq);
	}
	__shared__ int s_ok;
	if (threadIdx.x == 0) {
		int ok = 1;
		for (int d = 1; d < nranks && ok; d++) {
			const PfXchgHeader *ph = (const PfXchgHeader *)peers.base[(me + d) % nranks];
			ok = pf_wait_flag(&ph->dseq[buf], dseq, ph, status, timeout_ns);
		}
		s_ok = ok;
	}
	__syncthreads();
	if (!s_ok) return;
	const size_t off = PF_XCHG_HEADER_BYTES + 8 * (size_t)event_cap + sizeof(float) * (size_t)buf * (size_t)num_terminals;
	for (int t = (int)(blockIdx.x * blockDim.x + threadIdx.x); t < num_terminals; t += (int)(gridDim.x * blockDim.x)) {
		const int k = term_owner[t];
		if (k != me && k < nranks) net_delay[t] = *(const volatile float *)((const float *)(peers.base[k] + off) + t);
	}
}

__global__ void pf_xchg_abort_kernel(PfXchgHeader *mine) { mine->abort_flag = 1u; __threadfence_system(); }

/* ------------------------------------------------------------------ rr graph built on the device (pf_gen_device.cuh) */
__global__ void pf_gen_degree_kernel(const __grid_constant__ PfGenDev G, int *row) {
	for (int v = (int)(blockIdx.x * blockDim.x + threadIdx.x); v < G.num_nodes; v += (int)(gridDim.x * blockDim.x)) {
		const PfGenNode nd = pf_gen_decode(G, v);
		row[v] = pf_gen_node_edges(G, v, nd, NULL);
	}
}
/* exclusive prefix sum of an int array in three passes: per-CTA totals, one CTA scans the totals, per-CTA local scan */
#define PF_SCAN_CTA 1024
__global__ void __launch_bounds__(PF_SCAN_CTA) pf_scan_totals_kernel(const int *a, int n, long long *totals) {
	__shared__ long long s[PF_SCAN_CTA / 32];
	const int i = (int)blockIdx.x * PF_SCAN_CTA + (int)threadIdx.x;
	long long v = i < n ? a[i] : 0;
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31u) == 0) s[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) { long long t = 0; for (int w = 0; w < PF_SCAN_CTA / 32; w++) t += s[w]; totals[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(PF_SCAN_CTA) pf_scan_offsets_kernel(long long *totals, int nblocks, long long *grand) {
	/* one CTA: exclusive scan of the CTA totals, chunk by chunk */
	__shared__ long long s[PF_SCAN_CTA];
	__shared__ long long carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < nblocks; base += PF_SCAN_CTA) {
		const int i = base + (int)threadIdx.x;
		const long long v = i < nblocks ? totals[i] : 0;
		s[threadIdx.x] = v;
		__syncthreads();
		for (int o = 1; o < PF_SCAN_CTA; o <<= 1) {
			const long long t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0;
			__syncthreads();
			s[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < nblocks) totals[i] = carry + s[threadIdx.x] - v;
		__syncthreads();
		if (threadIdx.x == PF_SCAN_CTA - 1) carry += s[PF_SCAN_CTA - 1];
		__syncthreads();
	}
	if (threadIdx.x == 0) *grand = carry;
}
__global__ void __launch_bounds__(PF_SCAN_CTA) pf_scan_apply_kernel(int *a, int n, const long long *totals) {
	__shared__ int s[PF_SCAN_CTA];
	const int i = (int)blockIdx.x * PF_SCAN_CTA + (int)threadIdx.x;
	const int v = i < n ? a[i] : 0;
	s[threadIdx.x] = v;
	__syncthreads();
	for (int o = 1; o < PF_SCAN_CTA; o <<= 1) {
		const int t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0;
		__syncthreads();
		s[threadIdx.x] += t;
		__syncthreads();
	}
	if (i < n) a[i] = (int)(totals[blockIdx.x] + s[threadIdx.x] - v);
}
/* pass 2.  A warp takes 32 consecutive nodes: their edge rows are one contiguous span of the edge array (row[] is the prefix sum
 * of the degrees, row[N] = E).  Every lane enumerates its node's edges into the warp's staging buffer in shared memory and the
 * warp then copies the span out with coalesced stores — 32 lanes writing 4-byte words into 32 different rows cost 3x the
 * enumeration itself (10.5 ms vs 3.1 ms for the degree pass on cfg 4). */
#define PF_GEN_STAGE 1024           /* words per warp: 32 nodes x degree <= 32 */
__global__ void __launch_bounds__(256) pf_gen_fill_kernel(const __grid_constant__ PfGenDev G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc, unsigned long long *avail_wl) {
	__shared__ uint32_t stage[8][PF_GEN_STAGE];
	const int lane = (int)(threadIdx.x & 31u), wib = (int)(threadIdx.x >> 5);
	const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
	unsigned wl = 0;
	for (int base = warp * 32; base < G.num_nodes; base += nwarps * 32) {
		const int v = base + lane;
		const int last = min(base + 32, G.num_nodes);
		const int s0 = row[base], span = row[last] - s0;
		const bool staged = span <= PF_GEN_STAGE;
		if (v < G.num_nodes) {
			const PfGenNode nd = pf_gen_decode(G, v);
			const int start = row[v];
			const int deg = pf_gen_node_edges(G, v, nd, staged ? &stage[wib][start - s0] : edges + start);
			pf_gen_write_node(nd, start, deg, &nodes[v], &ptc[v]);
			if (nd.type == 4 || nd.type == 5) wl += (unsigned)(1 + nd.x1 - nd.x0 + nd.y1 - nd.y0);
		}
		__syncwarp();
		if (staged) for (int i = lane; i < span; i += 32) edges[s0 + i] = stage[wib][i];
		__syncwarp();
	}
	wl = __reduce_add_sync(0xffffffffu, wl);
	if (lane == 0 && wl) atomicAdd(avail_wl, (unsigned long long)wl);
}
__global__ void pf_scan_close_kernel(int *a, int n, const long long *grand) { a[n] = (int)*grand; }
__global__ void pf_reset_nodes_kernel(PfNode *nodes, int num_nodes) {
	for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < num_nodes; i += (int)(gridDim.x * blockDim.x)) { nodes[i].occ = 0; nodes[i].acc_cost = 1.f; }
}
static __device__ __forceinline__ unsigned long long pf_mix64(unsigned long long x) {
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
	return x;
}
__global__ void pf_graph_hash_kernel(const PfNode *nodes, int num_nodes, const uint32_t *edges, long long num_edges, const short *ptc, unsigned long long *out) {
	unsigned long long h0 = 0, h1 = 0, h2 = 0;
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
	for (long long i = tid; i < num_nodes; i += step) {
		const unsigned long long *w = (const unsigned long long *)&nodes[i];
		h0 += pf_mix64(w[0] ^ pf_mix64((unsigned long long)i)) + pf_mix64(w[1] + 0x9e3779b97f4a7c15ull * (unsigned long long)i) + pf_mix64(w[2] ^ (unsigned long long)(3 * i + 1)) + pf_mix64(w[3] ^ (unsigned long long)(5 * i + 2));
		h2 += pf_mix64(((unsigned long long)(unsigned short)ptc[i] << 32) ^ (unsigned long long)i);
	}
	for (long long i = tid; i < num_edges; i += step) h1 += pf_mix64(((unsigned long long)edges[i] << 32) ^ (unsigned long long)i);
	atomicAdd(&out[0], h0); atomicAdd(&out[1], h1); atomicAdd(&out[2], h2);
}

/* ------------------------------------------------------------------ launchers */
int pfb_launch_route(const PfParams *P, int num_slots, int warps_per_block) {
	if (warps_per_block < 1) warps_per_block = 1;
	if (warps_per_block > 4) warps_per_block = 4;   /* __launch_bounds__(128, 5) */
	int blocks = (num_slots + warps_per_block - 1) / warps_per_block;
	size_t smem = PF_SMEM_BLOCK_TABLES + (size_t)warps_per_block * (PF_SMEM_PER_WARP + (P->hot ? 0 : (size_t)PF_SMEM_HOT_ENTRIES * 8));
	typedef void (*RouteKernel)(const PfParams, int);
	/* [algorithm / search mode][ripple][bucketed far list] */
	static const RouteKernel variants[3][2][2] = {
		{ { pf_route_kernel<0, 0, 0>, pf_route_kernel<0, 0, 1> }, { pf_route_kernel<0, 1, 0>, pf_route_kernel<0, 1, 1> } },
		{ { pf_route_kernel<1, 0, 0>, pf_route_kernel<1, 0, 1> }, { pf_route_kernel<1, 1, 0>, pf_route_kernel<1, 1, 1> } },
		{ { pf_route_kernel<2, 0, 0>, pf_route_kernel<2, 0, 1> }, { pf_route_kernel<2, 1, 0>, pf_route_kernel<2, 1, 1> } } };
	static size_t smem_set = 0;
	if (smem > smem_set) {
		for (int a = 0; a < 3; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < 2; c++)
			CK(cudaFuncSetAttribute((const void *)variants[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		smem_set = smem;
	}
	if (ev_begin(0) != 0) return -1;
	/* the ripple variant (victim queues, node ownership) only where the launch uses them, the bucketed far list only in the big
	 * slots: the throughput-bound launches of the first iterations run the variant that does not carry that code at all */
	const int rip = (P->vq_ctl != NULL || P->committer != NULL) ? 1 : 0;
	const int mode = P->algorithm == 1 ? 2 : (P->max_batch == 1 ? 1 : 0);
	variants[mode][rip][P->far_buckets ? 1 : 0]<<<blocks, warps_per_block * 32, smem, g_stream>>>(*P, num_slots);
	g_pending[g_npending].slots = num_slots; g_pending[g_npending].variant = mode << 2 | rip << 1 | (P->far_buckets ? 1 : 0);
	return ev_end();
}

static int stream_grid(long long n) {
	long long b = (n + 255) / 256;
	long long cap = (long long)(g_sms > 0 ? g_sms : 148) * 8;   /* a multiple of the SM count */
	if (b > cap) b = cap;
	if (b < 1) b = 1;
	return (int)b;
}

int pfb_launch_update_cost(PfNode *nodes, int num_nodes, float acc_fac, int *d_overused, unsigned char *last_over, int iter_tag,
		unsigned long long *d_wl_used) {
	if (ev_begin(1) != 0) return -1;
	pf_update_cost_kernel<<<stream_grid(num_nodes), 256, 0, g_stream>>>(nodes, num_nodes, acc_fac, d_overused, last_over, iter_tag, d_wl_used);
	return ev_end();
}

int pfb_launch_apply_events(PfNode *nodes, const unsigned *events, long long count) {
	if (count <= 0) return 0;
	if (ev_begin(2) != 0) return -1;
	pf_apply_events_kernel<<<stream_grid(count), 256, 0, g_stream>>>(nodes, events, count);
	return ev_end();
}

int pfb_launch_wirelength(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, unsigned long long *d_out) {
	if (num_all <= 0) return 0;
	if (ev_begin(2) != 0) return -1;
	pf_wirelength_kernel<<<stream_grid(num_all), 256, 0, g_stream>>>(pool, loc, all_nets, num_all, d_out);
	return ev_end();
}

int pfb_launch_reserve_opins(PfNode *nodes, const uint32_t *edges, int node_bits, const PfIndexedDev *indexed,
		int num_groups, const int *group_source, const int *group_count, const int *group_off,
		int *chosen, int rip_up, float pres_fac) {
	if (num_groups <= 0) return 0;
	if (ev_begin(2) != 0) return -1;
	pf_reserve_opins_kernel<<<(num_groups + 127) / 128, 128, 0, g_stream>>>(nodes, edges, node_bits, indexed, num_groups, group_source, group_count, group_off, chosen, rip_up, pres_fac);
	return ev_end();
}

int pfb_launch_select_nets(const PfNode *nodes, const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets,
		int num_all, const unsigned char *net_big, int force_all, int *list_small, int *list_big, int *counts,
		const unsigned char *last_over, int iter_tag, int window, const int *committer, int *scratch, int head_count,
		int *queued, int queued_tag, const int *pool_node, const unsigned char *over_now, int over_tag) {
	if (cudaMemsetAsync(counts, 0, sizeof(int) * 4, g_stream) != cudaSuccess) return -1;
	if (num_all <= 0) return 0;
	/* scratch: [2 ints per CTA][one flag byte per net] — pfb_select_scratch_bytes() */
	const int blocks = (num_all + PF_SEL_BLOCK - 1) / PF_SEL_BLOCK;
	unsigned char *flag = (unsigned char *)(scratch + 2 * (size_t)blocks);
	if (ev_begin(2) != 0) return -1;
	pf_select_flag_kernel<<<blocks, PF_SEL_BLOCK, 0, g_stream>>>(nodes, pool, loc, all_nets, num_all, net_big, force_all, flag, scratch,
			last_over, iter_tag, window, committer, head_count, counts, queued, queued_tag, pool_node, over_now, over_tag);
	pf_select_scatter_kernel<<<blocks, PF_SEL_BLOCK, 0, g_stream>>>(all_nets, num_all, flag, scratch, list_small, list_big, counts);
	return ev_end();
}

size_t pfb_select_scratch_bytes(int num_all) {
	const size_t blocks = ((size_t)std::max(num_all, 1) + PF_SEL_BLOCK - 1) / PF_SEL_BLOCK;
	return 8 * blocks + (size_t)std::max(num_all, 1) + 16;
}

int pfb_launch_compact(const PfTreeNode *src, PfTreeNode *dst, const int *src_node, int *dst_node, PfNetLoc *loc, const int *all_nets, int num_all,
		unsigned long long *dst_head) {
	if (num_all <= 0) return 0;
	if (ev_begin(2) != 0) return -1;
	pf_compact_kernel<<<stream_grid((long long)num_all * 32), 256, 0, g_stream>>>(src, dst, src_node, dst_node, loc, all_nets, num_all, dst_head);
	return ev_end();
}

int pfb_launch_rebuild_owner(const PfTreeNode *pool, const PfNetLoc *loc, const int *all_nets, int num_all, int *owner) {
	if (num_all <= 0) return 0;
	if (ev_begin(2) != 0) return -1;
	pf_rebuild_owner_kernel<<<stream_grid((long long)num_all * 32), 256, 0, g_stream>>>(pool, loc, all_nets, num_all, owner);
	return ev_end();
}

int pfb_launch_extract_occ(const PfNode *nodes, int num_nodes, int *occ_out) {
	if (ev_begin(2) != 0) return -1;
	pf_extract_occ_kernel<<<stream_grid(num_nodes), 256, 0, g_stream>>>(nodes, num_nodes, occ_out);
	return ev_end();
}

int pfb_launch_build_traces(const PfTreeNode *pool, const PfNetLoc *loc, int num_nets, int *len, const int *tptr,
		int *trace_node, short *trace_switch, unsigned long long *d_wl, unsigned *trace_term, const short *ptc, int nx) {
	if (num_nets <= 0) return 0;
	if (ev_begin(2) != 0) return -1;
	pf_build_traces_kernel<<<(num_nets + 127) / 128, 128, 0, g_stream>>>(pool, loc, num_nets, len, tptr, trace_node, trace_switch, d_wl, trace_term, ptc, nx);
	return ev_end();
}


/* ------------------------------------------------------------------ device graph generation launchers */
static int gen_stage(const PfGenDev *G, PfGenDev *Gd, short **d_inv) {
	*Gd = *G;
	*d_inv = (short *)pfb_alloc_raw(sizeof(short) * (size_t)G->W);
	if (!*d_inv) return -1;
	CK(cudaMemcpyAsync(*d_inv, G->cb_inv, sizeof(short) * (size_t)G->W, cudaMemcpyHostToDevice, g_stream));
	Gd->cb_inv = *d_inv;
	return 0;
}
int pfb_gen_count(const PfGenDev *G, int *row, long long *num_edges) {
	PfGenDev Gd; short *d_inv = NULL;
	if (gen_stage(G, &Gd, &d_inv) != 0) return -1;
	const int n = G->num_nodes, nb = (n + PF_SCAN_CTA - 1) / PF_SCAN_CTA;
	long long *totals = (long long *)pfb_alloc_raw(sizeof(long long) * ((size_t)nb + 1));
	if (!totals) { pfb_free(d_inv); return -1; }
	if (ev_begin(2) != 0) return -1;
	pf_gen_degree_kernel<<<stream_grid(n), 256, 0, g_stream>>>(Gd, row);
	pf_scan_totals_kernel<<<nb, PF_SCAN_CTA, 0, g_stream>>>(row, n, totals);
	pf_scan_offsets_kernel<<<1, PF_SCAN_CTA, 0, g_stream>>>(totals, nb, totals + nb);
	pf_scan_apply_kernel<<<nb, PF_SCAN_CTA, 0, g_stream>>>(row, n, totals);
	pf_scan_close_kernel<<<1, 1, 0, g_stream>>>(row, n, totals + nb);          /* row[N] = E */
	if (ev_end() != 0) return -1;
	CK(cudaMemcpyAsync(num_edges, totals + nb, sizeof(long long), cudaMemcpyDeviceToHost, g_stream));
	CK(cudaStreamSynchronize(g_stream));
	pfb_free(totals); pfb_free(d_inv);
	return 0;
}
int pfb_gen_fill(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc, long long *avail_wl) {
	PfGenDev Gd; short *d_inv = NULL;
	if (gen_stage(G, &Gd, &d_inv) != 0) return -1;
	unsigned long long *d_wl = (unsigned long long *)pfb_alloc(sizeof(unsigned long long));
	if (!d_wl) { pfb_free(d_inv); return -1; }
	if (ev_begin(2) != 0) return -1;
	pf_gen_fill_kernel<<<stream_grid(G->num_nodes), 256, 0, g_stream>>>(Gd, row, nodes, edges, ptc, d_wl);
	if (ev_end() != 0) return -1;
	unsigned long long h = 0;
	CK(cudaMemcpyAsync(&h, d_wl, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
	CK(cudaStreamSynchronize(g_stream));
	*avail_wl = (long long)h;
	pfb_free(d_wl); pfb_free(d_inv);
	return 0;
}
static cudaStream_t g_side = 0;
static cudaEvent_t g_side_fork = 0, g_side_join = 0;
static struct { unsigned long long *d_wl; short *d_inv; int pending; } g_fill = { NULL, NULL, 0 };
int pfb_gen_fill_begin(const PfGenDev *G, const int *row, PfNode *nodes, uint32_t *edges, short *ptc) {
	PfGenDev Gd; short *d_inv = NULL;
	if (g_fill.pending) { long long w; if (pfb_gen_fill_end(&w) != 0) return -1; }
	if (!g_side) {
		CK(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&g_side_fork, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&g_side_join, cudaEventDisableTiming));
	}
	if (gen_stage(G, &Gd, &d_inv) != 0) return -1;
	unsigned long long *d_wl = (unsigned long long *)pfb_alloc(sizeof(unsigned long long));
	if (!d_wl) { pfb_free(d_inv); return -1; }
	/* the allocations (stream-ordered, on the router's stream), the staged table and the zeroed counter come first */
	CK(cudaEventRecord(g_side_fork, g_stream));
	CK(cudaStreamWaitEvent(g_side, g_side_fork, 0));
	pf_gen_fill_kernel<<<stream_grid(G->num_nodes), 256, 0, g_side>>>(Gd, row, nodes, edges, ptc, d_wl);
	CK(cudaGetLastError());
	CK(cudaEventRecord(g_side_join, g_side));
	g_fill.d_wl = d_wl; g_fill.d_inv = d_inv; g_fill.pending = 1;
	g_times.aux_launches++;
	return 0;
}
int pfb_gen_fill_end(long long *avail_wl) {
	if (!g_fill.pending) return 0;
	g_fill.pending = 0;
	unsigned long long h = 0;
	CK(cudaStreamWaitEvent(g_stream, g_side_join, 0));
	CK(cudaMemcpyAsync(&h, g_fill.d_wl, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
	CK(cudaStreamSynchronize(g_stream));
	if (avail_wl) *avail_wl = (long long)h;
	pfb_free(g_fill.d_wl); pfb_free(g_fill.d_inv);
	g_fill.d_wl = NULL; g_fill.d_inv = NULL;
	return 0;
}
int pfb_reset_nodes(PfNode *nodes, int num_nodes) {
	if (ev_begin(2) != 0) return -1;
	pf_reset_nodes_kernel<<<stream_grid(num_nodes), 256, 0, g_stream>>>(nodes, num_nodes);
	return ev_end();
}
int pfb_graph_hash(const PfNode *nodes, int num_nodes, const uint32_t *edges, long long num_edges, const short *ptc, unsigned long long out[3]) {
	unsigned long long *d = (unsigned long long *)pfb_alloc(sizeof(unsigned long long) * 3);
	if (!d) return -1;
	pf_graph_hash_kernel<<<stream_grid(num_nodes), 256, 0, g_stream>>>(nodes, num_nodes, edges, num_edges, ptc, d);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out, d, sizeof(unsigned long long) * 3, cudaMemcpyDeviceToHost, g_stream));
	CK(cudaStreamSynchronize(g_stream));
	pfb_free(d);
	return 0;
}

/* ------------------------------------------------------------------ exchange launchers / IPC memory */
void *pfb_ipc_alloc(size_t bytes, void *handle64) {
	void *p = NULL;
	if (cudaMalloc(&p, bytes) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu) for the exchange region failed", bytes); cudaGetLastError(); return NULL; }
	if (cudaMemsetAsync(p, 0, PF_XCHG_HEADER_BYTES, g_stream) != cudaSuccess || cudaStreamSynchronize(g_stream) != cudaSuccess
			|| cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, p) != cudaSuccess) {
		snprintf(g_err, sizeof(g_err), "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
		cudaFree(p);
		return NULL;
	}
	return p;
}
void *pfb_ipc_open(const void *handle64) {
	void *p = NULL;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return NULL; }
	return p;
}
int pfb_ipc_clear_abort(void *region) {
	CK(cudaMemsetAsync((char *)region + offsetof(PfXchgHeader, abort_flag), 0, sizeof(unsigned), g_stream));
	return 0;
}
void pfb_ipc_close(void *p) { if (p) cudaIpcCloseMemHandle(p); }
void pfb_ipc_free(void *p) { if (p) { cudaStreamSynchronize(g_stream); cudaFree(p); } }

int pfb_launch_xchg_events(PfNode *nodes, const PfPeers *peers, int me, int nranks, unsigned seq, const unsigned long long *event_head,
		long long event_cap, int *status, double timeout_s) {
	if (ev_begin(2) != 0) return -1;
	const int grid = (g_sms > 0 ? g_sms : 148) * 2;
	pf_xchg_events_kernel<<<grid, 256, 0, g_stream>>>(nodes, *peers, me, nranks, seq, event_head, event_cap, status, (unsigned long long)(timeout_s * 1e9));
	return ev_end();
}

int pfb_launch_xchg_delays(float *net_delay, const unsigned char *term_owner, int num_terminals, const PfPeers *peers, int me, int nranks,
		unsigned dseq, long long event_cap, int *status, double timeout_s) {
	if (ev_begin(2) != 0) return -1;
	const int buf = (int)((dseq - 1u) & 1u);
	float *pub = (float *)(peers->base[me] + PF_XCHG_HEADER_BYTES + 8 * (size_t)event_cap) + (size_t)buf * (size_t)num_terminals;
	pf_xchg_delays_publish_kernel<<<stream_grid(num_terminals), 256, 0, g_stream>>>(net_delay, term_owner, num_terminals, pub, me);
	pf_xchg_delays_gather_kernel<<<(g_sms > 0 ? g_sms : 148) * 2, 256, 0, g_stream>>>(net_delay, term_owner, num_terminals, *peers, me, nranks, dseq,
			event_cap, status, (unsigned long long)(timeout_s * 1e9));
	return ev_end();
}

int pfb_launch_xchg_abort(const PfPeers *peers, int me) {
	pf_xchg_abort_kernel<<<1, 1, 0, g_stream>>>((PfXchgHeader *)peers->base[me]);
	CK(cudaGetLastError());
	return 0;
}

/* ------------------------------------------------------------------ static timing analysis launchers
 * (not individually timed: a wide graph issues ~1000 level launches per analysis, two event records each would
 * double their cost) */
int pfb_sta_load(const PfStaDev *S, const float *dev_net_delay) {
	if (S->num_terminals <= 0) return 0;
	pf_sta_load_kernel<<<stream_grid(S->num_terminals), 256, 0, g_stream>>>(*S, dev_net_delay);
	CK(cudaGetLastError());
	return 0;
}

int pfb_sta_begin_pair(const PfStaDev *S, float *stat) {
	pf_sta_begin_pair_kernel<<<stream_grid(S->num_tnodes), 256, 0, g_stream>>>(*S, stat);
	CK(cudaGetLastError());
	return 0;
}

int pfb_sta_sweep(const PfStaDev *S, int forward, int lv_begin, int lv_end, int spread, int domain, float constraint, float *stat) {
	if (lv_end <= lv_begin) return 0;
	if (spread) pf_sta_level_kernel<<<stream_grid(spread), 256, 0, g_stream>>>(*S, forward, lv_begin, domain, constraint, stat);   /* spread = level width */
	else pf_sta_levels_kernel<<<1, PF_STA_CTA, 0, g_stream>>>(*S, forward, lv_begin, lv_end, domain, constraint, stat);
	CK(cudaGetLastError());
	return 0;
}

int pfb_sta_update(const PfStaDev *S, float constraint, const float *stat, float *dev_crit) {
	if (S->num_terminals <= 0) return 0;
	pf_sta_update_kernel<<<stream_grid(S->num_terminals), 256, 0, g_stream>>>(*S, constraint, stat, dev_crit);
	CK(cudaGetLastError());
	return 0;
}

int pfb_launch_check_route(const PfNode *nodes, const uint32_t *edges, int node_bits, int num_nodes, int num_nets, const int *net_ptr, const int *net_term,
		const unsigned char *net_is_global, const int *trace_ptr, const int *trace_node, const short *trace_switch, unsigned char *matched,
		int *occ2, const int *occ_reported, int *report, unsigned long long *wl_extra) {
	if (ev_begin(2) != 0) return -1;
	if (num_nets > 0)
		pf_check_nets_kernel<<<(num_nets + 127) / 128, 128, 0, g_stream>>>(nodes, edges, node_bits, num_nodes, num_nets, net_ptr, net_term, net_is_global,
				trace_ptr, trace_node, trace_switch, matched, occ2, report, wl_extra);
	pf_check_occ_kernel<<<stream_grid(num_nodes), 256, 0, g_stream>>>(nodes, num_nodes, occ2, occ_reported, report, wl_extra + 1);
	return ev_end();
}
