/*
 * pf_device.cuh — the warp-per-net PathFinder net router (device code, sm_100a).
 *
 * One warp routes one net at a time: rip-up → criticality shaping + sink order → per-sink
 * A*-directed label-correcting search with a shared-memory frontier (near set) that spills to a
 * per-warp far list in HBM (two-level delta-stepping: the near set holds every label within a
 * cost window of the current minimum; the far list is re-bucketed when the near set runs dry) →
 * back-trace → Elmore route-tree update → occupancy commit.  It replaces, for this path,
 *   timing_driven_route_net            reference vpr/SRC/route/route_timing.c:399-563
 *   add_route_tree_to_heap             route_timing.c:565-601      (tree seeding)
 *   timing_driven_expand_neighbours    route_timing.c:603-691      (edge relaxation)
 *   get_timing_driven_expected_cost    route_timing.c:693-840      (A* lookahead)
 *   mark_node_expansion_by_bin         route_timing.c:867-960      (high-fanout window)
 *   node_to_heap/add_to_heap/get_heap_head  route_common.c:780-803,1142-1216 (→ frontier)
 *   pathfinder_update_one_cost         route_common.c:533-579      (→ atomics on occ)
 *   update_traceback                   route_common.c:638-706      (→ tree entries)
 *   update_route_tree & friends        route_tree_timing.c:155-456 (Elmore)
 *
 * The arithmetic of every cost expression keeps the reference's float/double mix (SURVEY.md
 * Appendix E); the file is compiled with -fmad=false so the device evaluates them with the same
 * roundings as the CPU reference.  What differs from the serial reference is the ORDER in which
 * equal-cost labels are settled (a heap pops one label at a time, a warp settles a batch), so
 * route trees can differ where costs tie — never the cost model.
 *
 * The code is written against a dozen warp primitives (pf_shfl_*, pf_ballot, pf_match_any, ...).
 * Under nvcc they are the hardware intrinsics below.  tests/emu/pf_emu.h provides the same names
 * on top of cooperative fibers so the identical source can be exercised on a CPU-only box; that
 * emulator is test infrastructure and is never part of libpf_router.so.
 */
#ifndef PF_DEVICE_CUH
#define PF_DEVICE_CUH

#include <stdint.h>

#ifndef PF_EMU
#include <cuda_runtime.h>
#define PF_DEV static __device__ __forceinline__
#define PF_WARP 32
typedef uint4 pf_u4;
PF_DEV int pf_lane(void) { return (int)(threadIdx.x & 31u); }
PF_DEV void pf_syncwarp(void) { __syncwarp(); }
PF_DEV unsigned pf_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
PF_DEV int pf_any(int pred) { return __any_sync(0xffffffffu, pred); }
PF_DEV int pf_shfl_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
PF_DEV float pf_shfl_f(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
PF_DEV uint64_t pf_shfl_u64(uint64_t v, int src) { return __shfl_sync(0xffffffffu, (unsigned long long)v, src); }
PF_DEV unsigned pf_match_any(int key) { return __match_any_sync(0xffffffffu, key); }
PF_DEV uint64_t pf_warp_min_u64(uint64_t v) {
	/* two 32-bit hardware reductions (REDUX): high word first, then the low word among the ties */
	unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
	unsigned mh = __reduce_min_sync(0xffffffffu, hi);
	unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
	return ((uint64_t)mh << 32) | ml;
}
PF_DEV float pf_warp_min_f(float v) {
	/* valid for non-negative floats and +inf: their bit patterns order like unsigned ints */
	return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}
PF_DEV int pf_warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }
PF_DEV int pf_warp_max_i(int v) { return __reduce_max_sync(0xffffffffu, v); }
PF_DEV int pf_popc(unsigned m) { return __popc(m); }
PF_DEV int pf_ffs(unsigned m) { return __ffs((int)m); }
PF_DEV unsigned pf_lanemask_lt(void) { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
PF_DEV int pf_atomic_add_i(int *p, int v) { return atomicAdd(p, v); }
PF_DEV unsigned long long pf_atomic_add_ull(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
PF_DEV void pf_atomic_max_ull(unsigned long long *p, unsigned long long v) { atomicMax(p, v); }
PF_DEV int pf_atomic_or_i(int *p, int v) { return atomicOr(p, v); }
PF_DEV int pf_atomic_min_i(int *p, int v) { return atomicMin(p, v); }   /* used on shared memory (ATOMS) */
PF_DEV int pf_atomic_exch_i(int *p, int v) { return atomicExch(p, v); }
PF_DEV int pf_atomic_cas_i(int *p, int cmp, int v) { return atomicCAS(p, cmp, v); }
PF_DEV int pf_ld_volatile_i(const int *p) { return *(const volatile int *)p; }
PF_DEV void pf_threadfence(void) { __threadfence(); }
PF_DEV void pf_spin_pause(void) { __nanosleep(200); }
PF_DEV pf_u4 pf_ld_cg_u4(const void *p) { return __ldcg((const uint4 *)p); }   /* L2-coherent: sees other SMs' atomics */
PF_DEV pf_u4 pf_ld_u4(const void *p) { return *(const uint4 *)p; }
struct pf_u8 { unsigned a, b, c, d, e, f, g, h; };
/* one 256-bit L2-coherent load (LDG.E.256, sm_100): a whole 32-byte node record in one request */
PF_DEV pf_u8 pf_ld_cg_u8(const void *p) {
	pf_u8 v;
	asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.a), "=r"(v.b), "=r"(v.c), "=r"(v.d), "=r"(v.e), "=r"(v.f), "=r"(v.g), "=r"(v.h) : "l"(p));
	return v;
}
PF_DEV void pf_st_u4(void *p, pf_u4 v) { *(uint4 *)p = v; }
PF_DEV void pf_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
PF_DEV float pf_int_as_float(int i) { return __int_as_float(i); }
PF_DEV int pf_float_as_int(float f) { return __float_as_int(f); }
PF_DEV double pf_ceil(double x) { return ceil(x); }
PF_DEV float pf_sqrtf(float x) { return sqrtf(x); }
PF_DEV float pf_ceilf(float x) { return ceilf(x); }
PF_DEV float pf_powf(float x, float y) { return powf(x, y); }
#else
#include "pf_emu.h"   /* tests/emu: fiber warp emulator, test builds only */
#endif

#include "pf_layout.h"

/* ------------------------------------------------------------------ per-warp context */
struct PfWarp {
	const PfParams *P;
	/* shared memory */
	uint64_t *fr;                 /* [PF_SH_FRONTIER] key = tot bits << 32 | node */
	uint64_t *b_key;              /* [PF_MAX_BATCH] */
	int *b_node; float *b_back; float *b_R; int *b_start; int *b_pre; int *b_type;
	float *base_cost;             /* [PF_MAX_INDEXED] per-net rescaled base costs */
	int *ticket;                  /* [PF_TICKETS] slot-write arbitration */
	PfIndexedDev *idx;            /* [PF_MAX_INDEXED] */
	PfSwitchDev *sw;              /* [PF_MAX_SWITCHES] */
	/* slot memory */
	uint64_t *hot; PfCold *cold; unsigned label_mask; int label_shift; int label_limit;
	uint64_t *hot_alt; PfCold *cold_alt; unsigned mask_alt; int shift_alt; int limit_alt; unsigned epoch_alt;   /* the other table */
	PfTreeNode *tree; uint64_t *far; int *iscratch;
	/* search state: warp-uniform */
	unsigned tag_mask; int nb;    /* search-tag mask and node-id width of this router (PfParams.node_bits), kept in registers */
	unsigned epoch; unsigned round; int n_labels; int sh_n; float T_hi; float far_min; float best;
	int st_n;                     /* entries in the staging / flat far list */
	int overflow;
	/* per-net constants */
	int bb_xmin, bb_xmax, bb_ymin, bb_ymax; int num_sinks; int cur_net;
	/* counters */
	unsigned pops, pushes, visits, refills, stale, races;   /* per launch and warp: 32 bits are plenty, and six registers fewer */
	unsigned max_net_pops;
};

/* per warp: fr 1024 + b_key 256 + base_cost 128 + 5 batch arrays 640 + b_pre 136 + tickets 256 = 2440 → 2448;
 * per CTA: cost-index table 1024 + switch table 768 */
#define PF_INFO_SEEN_FULL_BIT 11   /* label info word: switch (8) | type (3) | node was full when priced (1) | out-degree << 16 */
#define PF_OVF_LABELS 1   /* w.overflow bits */
#define PF_OVF_OTHER 2
#define PF_TICKETS 64
#define PF_SMEM_PER_WARP 2448                       /* + PF_SMEM_HOT_ENTRIES * 8 when the hot table is in shared memory */
#define PF_SMEM_BLOCK_TABLES (PF_MAX_INDEXED * 32 + PF_MAX_SWITCHES * 12)

PF_DEV float pf_key_tot(uint64_t k) { return pf_int_as_float((int)(k >> 32)); }
PF_DEV int pf_key_node(uint64_t k) { return (int)(uint32_t)k; }
/* frontier key: total cost | rr node.  Equal totals are settled oldest first (the near set keeps push order), so
 * the key carries no tie-break field. */
PF_DEV uint64_t pf_make_key(float tot, int node) {
	return ((uint64_t)(uint32_t)pf_float_as_int(tot) << 32) | (uint32_t)node;
}
#define PF_INF_F 3.0e38f
/* Cold per-net loops (rip-up, commit, undo, table wipes): nvcc unrolls them 4x and they are inlined at several sites, which
 * was a fifth of the kernel's code (tools/sass_by_line.py).  They are kept rolled: 7,432 -> 5,432 SASS instructions for the
 * strict kernel and 3 % less kernel time on BASELINE configs[4] (17.65 -> 17.05 ms, profiles/r02_ab_variants.txt);
 * -DPF_UNROLL_COLD=1 restores nvcc's default for A/B builds. */
#if !defined(PF_UNROLL_COLD) && defined(__CUDACC__)
#define PF_COLD_LOOP _Pragma("unroll 1")
#else
#define PF_COLD_LOOP
#endif
/* pf_refill's passes over the far list: unrolled 4x they are 1,170 SASS instructions in the middle of the search loop — a fifth
 * of the kernel's code for a function that runs once per few dozen settled labels */
#if !defined(PF_UNROLL_REFILL) && defined(__CUDACC__)
#define PF_REFILL_LOOP _Pragma("unroll 1")
#else
#define PF_REFILL_LOOP
#endif
#define PF_KEY_MAX 0xffffffffffffffffull

/* ------------------------------------------------------------------ A* lookahead
 * get_expected_segs_to_target + get_timing_driven_expected_cost, route_timing.c:693-840 */
#define PF_ROUND_UP(x) (pf_ceil((x) - 0.001))

PF_DEV float pf_expected_cost(const PfWarp &w, int type, int ci, int ixlow, int ixhigh, int iylow, int iyhigh,
		int target_x, int target_y, float criticality_fac, float R_upstream) {
	if (type == 4 || type == 5) {
		int num_segs_same_dir, num_segs_ortho_dir, no_need_to_pass_by_clb;
		const PfIndexedDev &I = w.idx[ci];
		int oci = I.ortho;
		const PfIndexedDev &O = w.idx[oci];
		float inv_length = I.inv_length, ortho_inv_length = O.inv_length;
		float ylow = iylow, yhigh = iyhigh, xlow = ixlow, xhigh = ixhigh;
		/* the reference spells the CHANX and the CHANY case out separately (route_timing.c:711-772); they are the
		 * same arithmetic with x and y exchanged, so one copy runs on (along-the-wire, across-the-wire)
		 * coordinates — a warp holds both kinds, and two branches would execute one after the other */
		const bool cx = (type == 4);
		const float a_low = cx ? xlow : ylow, a_high = cx ? xhigh : yhigh;   /* along the wire */
		const float o_low = cx ? ylow : xlow;                                /* across */
		const int t_a = cx ? target_x : target_y, t_o = cx ? target_y : target_x;
		if (o_low > t_o) {
			num_segs_ortho_dir = (int)(PF_ROUND_UP((o_low - t_o + 1.) * ortho_inv_length));
			no_need_to_pass_by_clb = 1;
		} else if (o_low < t_o - 1) {
			num_segs_ortho_dir = (int)(PF_ROUND_UP((t_o - o_low) * ortho_inv_length));
			no_need_to_pass_by_clb = 1;
		} else {
			num_segs_ortho_dir = 0;
			no_need_to_pass_by_clb = 0;
		}
		if (a_low > t_a + no_need_to_pass_by_clb)
			num_segs_same_dir = (int)(PF_ROUND_UP((a_low - no_need_to_pass_by_clb - t_a) * inv_length));
		else if (a_high < t_a - no_need_to_pass_by_clb)
			num_segs_same_dir = (int)(PF_ROUND_UP((t_a - no_need_to_pass_by_clb - a_high) * inv_length));
		else
			num_segs_same_dir = 0;
		float cong_cost = num_segs_same_dir * w.base_cost[ci] + num_segs_ortho_dir * w.base_cost[oci];
		cong_cost += w.base_cost[3] + w.base_cost[1];   /* IPIN_COST_INDEX, SINK_COST_INDEX */
		/* criticality 0 (timing analysis off, or a sink below the 1 - max_criticality cut): the delay term is
		 * multiplied by exactly 0 and (1. - 0) * cong_cost is exact, so the result IS cong_cost bit for bit */
		if (criticality_fac == 0.f) return cong_cost;
		float Tdel = num_segs_same_dir * I.T_linear + num_segs_ortho_dir * O.T_linear
				+ num_segs_same_dir * num_segs_same_dir * I.T_quadratic
				+ num_segs_ortho_dir * num_segs_ortho_dir * O.T_quadratic
				+ R_upstream * (num_segs_same_dir * I.C_load + num_segs_ortho_dir * O.C_load);
		Tdel += w.idx[3].T_linear;
		float expected_cost = criticality_fac * Tdel + (1. - criticality_fac) * cong_cost;
		return expected_cost;
	} else if (type == 2) { /* IPIN */
		return w.base_cost[1];
	}
	return 0.f;
}

/* ------------------------------------------------------------------ label table */
#ifdef PF_FIXED_NODE_BITS            /* A/B builds: the node-id width as a compile-time constant */
#define PF_NB(P) PF_FIXED_NODE_BITS
#else
#define PF_NB(P) ((P)->node_bits)
#endif
PF_DEV unsigned pf_node_mask(int node_bits) { return (1u << node_bits) - 1u; }
PF_DEV unsigned pf_tag_mask(const PfParams *P) { return (1u << (32 - PF_NB(P))) - 1u; }      /* search tag: the bits above the node id */
PF_DEV unsigned pf_hash(const PfWarp &w, int node) {
	return ((uint32_t)node * 2654435761u) >> w.label_shift;
}
PF_DEV uint64_t pf_hot_make(const PfWarp &w, float tot, int node) {
	return ((uint64_t)(uint32_t)pf_float_as_int(tot) << 32) | ((w.epoch & w.tag_mask) << w.nb) | (uint32_t)node;
}
PF_DEV int pf_hot_live(const PfWarp &w, uint64_t k) { return (((uint32_t)k) >> w.nb) == (w.epoch & w.tag_mask); }
PF_DEV int pf_hot_node(const PfWarp &w, uint64_t k) { return (int)((uint32_t)k & ~(w.tag_mask << w.nb)); }

/* Look up an existing label (used at settle time and in the back-trace).  Per-lane, no collectives. */
PF_DEV int pf_label_find(const PfWarp &w, int node) {
	unsigned h = pf_hash(w, node);
	for (;;) {
		uint64_t k = w.hot[h];
		if (!pf_hot_live(w, k)) return -1;
		if (pf_hot_node(w, k) == node) return (int)h;
		h = (h + 1) & w.label_mask;
	}
}

/* Warp-collective relax: every lane may offer one candidate label (valid != 0).  A candidate replaces an
 * existing label only if both its total and its backward cost are lower (the pop rule of
 * route_timing.c:511, applied at relax time).  Two lanes may target the same table slot in the same
 * round — the same node reached over two edges, or two nodes probing the same empty slot; a per-warp
 * shared-memory ticket array arbitrates (atomicMin: lowest lane wins): one writer per ticket and round,
 * the others re-probe in the next round and see the winner's label.  Returns 1 in lanes whose
 * candidate was written. */
PF_DEV int pf_label_relax(PfWarp &w, int valid, int node, float tot, float back, float R_up, int prev, int info,
		int edge_start) {
	const int lane = pf_lane();
	unsigned h = pf_hash(w, node);
	int pending = valid, written = 0;
	while (pf_any(pending)) {
		int want = 0;                                   /* 1: claim an empty slot, 2: improve my node's label */
		if (pending) {
			uint64_t k = w.hot[h];
			if (pf_hot_live(w, k)) {
				if (pf_hot_node(w, k) == node) {
					float otot = pf_int_as_float((int)(k >> 32));
					if (tot < otot && back < w.cold[h].back) want = 2; else pending = 0;
				} else {
					h = (h + 1) & w.label_mask;
				}
			} else {
				want = 1;
			}
		}
		/* lowest lane wins a contested ticket (deterministic); the round tag decreases every round so
		 * tickets never need clearing */
		if ((w.round & 0x3ffffffu) == 0x3ffffffu) {            /* the tag is about to wrap: start the tickets afresh */
			for (int i = lane; i < PF_TICKETS; i += PF_WARP) w.ticket[i] = 0x7fffffff;
			w.round++;
			pf_syncwarp();
		}
		const int tk = (int)(((0x3ffffffu - (w.round & 0x3ffffffu)) << 5) | (unsigned)lane);
		w.round++;
		if (want) pf_atomic_min_i(&w.ticket[h & (PF_TICKETS - 1)], tk);
		pf_syncwarp();
		int win = want && w.ticket[h & (PF_TICKETS - 1)] == tk;
		if (win) {
			pf_u4 c0, c1;
			c0.x = (unsigned)pf_float_as_int(back); c0.y = (unsigned)pf_float_as_int(R_up); c0.z = (unsigned)prev; c0.w = (unsigned)info;
			c1.x = (unsigned)edge_start; c1.y = c1.z = c1.w = 0;
			w.hot[h] = pf_hot_make(w, tot, node);
			pf_st_u4(&w.cold[h], c0); pf_st_u4((char *)&w.cold[h] + 16, c1);
			written = 1; pending = 0;
		}
		pf_syncwarp();                                  /* tickets reusable; label stores ordered before re-probes */
		w.n_labels += pf_popc(pf_ballot(win && want == 1));
	}
	if (w.n_labels > w.label_limit) w.overflow |= PF_OVF_LABELS;
	return written;
}

/* ------------------------------------------------------------------ frontier
 * Two levels, delta-stepping style.
 *   near set   shared memory, every label with total <= T_hi: the settle loop pops its exact minimum.
 *   far list   the slot's array in HBM / L2, cut into PF_BUCKETS cost buckets of width delta (= the near window at the start
 *              of the search) above `bk_base`: bucket b holds totals in [base + b delta, base + (b+1) delta); the last bucket
 *              is the catch-all for everything beyond, and for what does not fit its own bucket.  Lane b keeps the fill count
 *              of bucket b in a register, so no shared memory is spent on it.
 * A refill only touches the lowest non-empty bucket (and the catch-all when it holds something that belongs there), instead
 * of scanning the whole far list three times: on flooding searches — timing-driven nets at pres_fac >= 3 settle 10^4..10^5
 * labels — the flat list made a refill O(labels in flight) and the search quadratic (37 s of the 38 s of the 32 k-LUT
 * circuit were spent in 64 big slots doing that). */
#define PF_BUCKETS 32
#define PF_NREG 30                     /* regular buckets 0 .. 29 */
#define PF_SPILL 30                    /* labels whose own regular bucket was full */
#define PF_BEYOND 31                   /* labels beyond the range of the regular buckets (and what the spill bucket could not take) */
/* the frontier state the refill works on, handed to the out-of-line refill BY VALUE and returned (the warp context itself must
 * never have its address taken: it would move from registers to local memory for the whole kernel) */
struct PfFront { uint64_t *fr; uint64_t *far; int far_cap; float win_rel, win_abs; int sh_n, bk_n, st_n; float bk_base, bk_inv, spill_min, ovf_min, far_min, T_hi; int overflow; unsigned refills;
#ifdef PF_DIAG
	unsigned scanned, scanned_ovf;
#endif
};
/* the slot's far array, in units of cr = far_cap / 256 entries: [0, 30 cr) the regular buckets, [30 cr, 32 cr) the spill bucket,
 * [32 cr, 128 cr) the bucket of everything beyond the range; upper half = the staging list, which IS the far list of a search
 * that never needed buckets (PF_FLAT_MAX).
 * Pushes only APPEND to the staging list (one ballot, one store — most far labels of an A* search are never needed: cfg 4
 * pushes 40 M labels and settles 5 M); a refill deals the staged labels out to the buckets first.
 * The spill bucket is separate from the bucket of the beyond because a spilled label belongs to the band being drawn from: it
 * makes every refill of that band consult the list it sits in — measured with ONE catch-all for both, a flooding search
 * re-scanned 640 k far-future labels at each of its 2 k refills (1.4 G entries per iteration, 4.5 s for one net). */
PF_DEV int pf_bucket_cap(const PfFront &f, int b) { const int cr = f.far_cap >> 8; return b < PF_SPILL ? cr : (b == PF_SPILL ? 2 * cr : (f.far_cap >> 1) - 32 * cr); }
PF_DEV uint64_t *pf_bucket_ptr(const PfFront &f, int b) { return f.far + (size_t)(b < PF_BEYOND ? b : 32) * (size_t)(f.far_cap >> 8); }
PF_DEV int pf_stage_off(int far_cap) { return far_cap >> 1; }
PF_DEV int pf_stage_cap(int far_cap) { return far_cap - (far_cap >> 1) - 32; }      /* the last 32 words hold the bucket state */
PF_DEV int pf_bucket_of(const PfFront &f, float tot) {
	const float x = (tot - f.bk_base) * f.bk_inv;
	return x < 0.f ? 0 : (x >= (float)PF_NREG ? PF_BEYOND : (int)x);
}

#if defined(__CUDACC__) && !defined(PF_EMU) && !defined(PF_FAR_INLINE)
#define PF_NOINLINE static __device__ __noinline__
#elif defined(__CUDACC__) && !defined(PF_EMU)
#define PF_NOINLINE static __device__ __forceinline__
#else
#define PF_NOINLINE static inline
#endif
/* the scans of the out-of-line bucket code: four independent loads in flight per lane (a rolled loop waits a full L2 / HBM
 * round trip per 32 entries; code size is no concern outside the search loop) */
#if defined(__CUDACC__)
#define PF_BUCKET_LOOP _Pragma("unroll 4")
#else
#define PF_BUCKET_LOOP
#endif

/* Warp-collective insertion into the buckets: every active lane files one key. */
PF_DEV void pf_far_put(PfFront &f, int active, uint64_t key, float tot) {
	const int lane = pf_lane();
	int b = active ? pf_bucket_of(f, tot) : -1;
	unsigned rem = pf_ballot(active);
	int ovf = 0;
	while (rem) {
		const int bk = pf_shfl_i(b, pf_ffs(rem) - 1);
		const unsigned m = pf_ballot(b == bk);
		const int n0 = pf_shfl_i(f.bk_n, bk), cap = pf_bucket_cap(f, bk);
		const int pos = n0 + pf_popc(m & pf_lanemask_lt());
		int moved = 0;
		if (b == bk) {
			if (pos < cap) pf_bucket_ptr(f, bk)[pos] = key;
			else if (bk < PF_SPILL) { b = PF_SPILL; moved = 1; }        /* its own bucket is full */
			else if (bk == PF_SPILL) { b = PF_BEYOND; moved = 1; }
			else ovf = 1;                                               /* every list it could go to is full: the net is retried in a bigger slot */
		}
		if (lane == bk) f.bk_n = (n0 + pf_popc(m) < cap) ? n0 + pf_popc(m) : cap;
		rem = (rem & ~m) | pf_ballot(moved);
	}
	if (pf_any(ovf)) f.overflow = 1;
	const float fm = pf_warp_min_f(active ? tot : PF_INF_F);
	if (fm < f.far_min) f.far_min = fm;
	const float sm = pf_warp_min_f((active && b == PF_SPILL) ? tot : PF_INF_F);
	if (sm < f.spill_min) f.spill_min = sm;
	const float om = pf_warp_min_f((active && b == PF_BEYOND) ? tot : PF_INF_F);
	if (om < f.ovf_min) f.ovf_min = om;
}
/* a list of n keys dealt out to the buckets, four loads ahead of the filing */
PF_DEV void pf_far_deal(PfFront &f, const uint64_t *a, int n) {
	const int lane = pf_lane();
	for (int base = 0; base < n; base += 4 * PF_WARP) {
		uint64_t k[4];
#pragma unroll
		for (int j = 0; j < 4; j++) { const int i = base + j * PF_WARP + lane; k[j] = i < n ? a[i] : 0; }
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int i = base + j * PF_WARP + lane;
			if (base + j * PF_WARP < n) pf_far_put(f, i < n, k[j], pf_key_tot(k[j]));
		}
	}
}

/* One list of the refill: minimum key / entries with total <= T / move the first (PF_SH_REFILL - taken) of those to the near
 * set, close the gaps, return the minimum total that stays. */
PF_DEV uint64_t pf_blist_min(const uint64_t *a, int n) {
	uint64_t mk = PF_KEY_MAX;
	PF_BUCKET_LOOP for (int i = pf_lane(); i < n; i += PF_WARP) { const uint64_t k = a[i]; if (k < mk) mk = k; }
	return mk;
}
PF_DEV int pf_blist_count(const uint64_t *a, int n, float T) {
	int c = 0;
	PF_BUCKET_LOOP for (int i = pf_lane(); i < n; i += PF_WARP) if (pf_key_tot(a[i]) <= T) c++;
	return c;
}
PF_DEV float pf_blist_take(PfFront &f, uint64_t *a, int n, float T, int *taken_io, int *kept_out) {
	const int lane = pf_lane();
	int kept = 0, taken = *taken_io;
	float fmin = PF_INF_F;
	for (int sbase = 0; sbase < n; sbase += 4 * PF_WARP) {
		uint64_t kk[4];                                        /* read four chunks ahead: the in-place writes below stay behind them */
#pragma unroll
		for (int j = 0; j < 4; j++) { const int i = sbase + j * PF_WARP + lane; kk[j] = i < n ? a[i] : PF_KEY_MAX; }
		pf_syncwarp();
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int i = sbase + j * PF_WARP + lane;
			const uint64_t k = kk[j];
			const int q = (i < n) && pf_key_tot(k) <= T;
			const unsigned mq = pf_ballot(q);
			const int rank = taken + pf_popc(mq & pf_lanemask_lt());
			const int take = q && rank < PF_SH_REFILL;
			const unsigned mt = pf_ballot(take);
			const int keep = (i < n) && !take;
			const unsigned mkp = pf_ballot(keep);
			if (take) f.fr[taken + pf_popc(mt & pf_lanemask_lt())] = k;
			if (keep) {
				a[kept + pf_popc(mkp & pf_lanemask_lt())] = k;      /* kept <= the chunk's first index: behind everything still unread */
				const float t = pf_key_tot(k);
				if (t < fmin) fmin = t;
			}
			taken += pf_popc(mt);
			kept += pf_popc(mkp);
		}
	}
	*taken_io = taken; *kept_out = kept;
	return fmin;
}

/* Regular and spill buckets are empty: open a new range of buckets at the minimum of the beyond and deal its entries out. */
PF_DEV void pf_far_rebase(PfFront &f) {
	const int lane = pf_lane();
	const int n = pf_shfl_i(f.bk_n, PF_BEYOND);
	uint64_t *ovf = pf_bucket_ptr(f, PF_BEYOND);
	const float m = pf_key_tot(pf_warp_min_u64(pf_blist_min(ovf, n)));
	float win = m * f.win_rel;
	if (win < f.win_abs) win = f.win_abs;
	f.bk_base = m; f.bk_inv = 1.f / win;
	f.spill_min = PF_INF_F;                                /* (the spill bucket is empty: that is when a range is closed) */
	if (lane == PF_BEYOND) f.bk_n = 0;
	int kept = 0;
	float omin = PF_INF_F;
	for (int sbase = 0; sbase < n; sbase += 4 * PF_WARP) {
		uint64_t kk[4];
#pragma unroll
		for (int j = 0; j < 4; j++) { const int i = sbase + j * PF_WARP + lane; kk[j] = i < n ? ovf[i] : PF_KEY_MAX; }
		pf_syncwarp();
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int i = sbase + j * PF_WARP + lane;
			const uint64_t k = kk[j];
			int b = i < n ? pf_bucket_of(f, pf_key_tot(k)) : -1;
			/* regular buckets first (their regions are disjoint from the list being read) */
			unsigned rem = pf_ballot(b >= 0 && b < PF_NREG);
			float smin = PF_INF_F;
			while (rem) {
				const int bk = pf_shfl_i(b, pf_ffs(rem) - 1);
				const unsigned mm = pf_ballot(b == bk);
				const int n0 = pf_shfl_i(f.bk_n, bk), cap = pf_bucket_cap(f, bk);
				const int pos = n0 + pf_popc(mm & pf_lanemask_lt());
				int moved = 0;
				if (b == bk) {
					if (pos < cap) { pf_bucket_ptr(f, bk)[pos] = k; if (bk == PF_SPILL) smin = pf_key_tot(k); }
					else if (bk < PF_SPILL) { b = PF_SPILL; moved = 1; }       /* its bucket is full: the (empty) spill bucket takes it, */
					else b = PF_BEYOND;                                        /* or it stays where it is */
				}
				if (lane == bk) f.bk_n = (n0 + pf_popc(mm) < cap) ? n0 + pf_popc(mm) : cap;
				rem = (rem & ~mm) | pf_ballot(moved);
			}
			smin = pf_warp_min_f(smin);
			if (smin < f.spill_min) f.spill_min = smin;
			const unsigned mk2 = pf_ballot(b == PF_BEYOND);
			if (b == PF_BEYOND) {
				ovf[kept + pf_popc(mk2 & pf_lanemask_lt())] = k;   /* kept <= the chunk's first index: in place */
				const float t = pf_key_tot(k);
				if (t < omin) omin = t;
			}
			kept += pf_popc(mk2);
		}
	}
	if (lane == PF_BEYOND) f.bk_n = kept;
	f.ovf_min = pf_warp_min_f(omin);
	pf_syncwarp();
}

/* One list of the inline flat refill (rolled: it sits in the search loop of every kernel variant) */
PF_DEV uint64_t pf_list_min(const uint64_t *a, int n) {
	uint64_t mk = PF_KEY_MAX;
	PF_REFILL_LOOP for (int i = pf_lane(); i < n; i += PF_WARP) { const uint64_t k = a[i]; if (k < mk) mk = k; }
	return mk;
}
PF_DEV int pf_list_count(const uint64_t *a, int n, float T) {
	int c = 0;
	PF_REFILL_LOOP for (int i = pf_lane(); i < n; i += PF_WARP) if (pf_key_tot(a[i]) <= T) c++;
	return c;
}

/* Re-bucket: file the near set's leftovers and the staged labels, find the lowest non-empty bucket, open a window above its
 * minimum and pull every label inside the window into shared memory (at most PF_SH_REFILL; the window is narrowed until they
 * fit, down to exact ties of the minimum).  The spill bucket and the beyond are consulted only while they hold a total below
 * the upper bound of the bucket drawn from. */
/* A search whose far list stays short — every A* search of BASELINE configs[4]: 67 far labels per sink on average — never
 * builds buckets: its staging list is scanned directly, with the window relative to the current minimum, as in round 1. */
#ifndef PF_FLAT_MAX
#define PF_FLAT_MAX 768
#endif
#ifdef PF_EMU
extern long long pf_emu_bucket_refills;   /* test hooks of the emulator build: refills that went through the cost buckets, */
extern long long pf_emu_lazy_seedings;    /* searches seeded lazily (low word) and their returns for more seeds (high word) */
#endif
#define PF_LAZY_SEED_SPAN 4.f
#define PF_ROUND_BUCKETS 0x40000000u   /* bit of PfWarp.round: this search has opened its cost buckets (their state lives in the
                                       * last 256 bytes of the slot's far array, not in registers: it is touched once per refill) */
PF_DEV void pf_refill_body(PfFront &f) {
	const int lane = pf_lane();
	f.refills++;
#ifdef PF_DIAG
	f.scanned += (unsigned)(f.st_n + f.sh_n);
#endif
	if (f.bk_inv == 0.f) {                                /* the search has outgrown the flat list: open the buckets at the current minimum */
		uint64_t mk = pf_blist_min(f.far + pf_stage_off(f.far_cap), f.st_n);
		{ const uint64_t mn = pf_blist_min(f.fr, f.sh_n); if (mn < mk) mk = mn; }
		const float m = pf_key_tot(pf_warp_min_u64(mk));
		float win = m * f.win_rel;
		if (win < f.win_abs) win = f.win_abs;
		f.bk_base = m; f.bk_inv = 1.f / win;
	}
	/* 1. the near set's leftovers, then everything pushed since the last refill, into the buckets */
	pf_far_deal(f, f.fr, f.sh_n);
	f.sh_n = 0;
	pf_far_deal(f, f.far + pf_stage_off(f.far_cap), f.st_n);
	f.st_n = 0;
	pf_syncwarp();
	/* 2. the bucket to draw from */
	unsigned reg = pf_ballot(f.bk_n > 0 && lane < PF_NREG);
	int n_sp = pf_shfl_i(f.bk_n, PF_SPILL), n_by = pf_shfl_i(f.bk_n, PF_BEYOND);
	if (!reg && n_sp == 0) {
		if (n_by == 0) { f.far_min = PF_INF_F; return; }
#ifdef PF_DIAG
		f.scanned_ovf += (unsigned)(2 * n_by);
#endif
		pf_far_rebase(f);
		reg = pf_ballot(f.bk_n > 0 && lane < PF_NREG);
		n_by = pf_shfl_i(f.bk_n, PF_BEYOND);
	}
	/* (only the spill bucket left inside the range — or, when no regular bucket could take anything, only the beyond) */
	const int b = reg ? pf_ffs(reg) - 1 : (n_sp ? PF_SPILL : PF_BEYOND);
	const float ub = b < PF_SPILL ? f.bk_base + (float)(b + 1) / f.bk_inv : (b == PF_SPILL ? f.bk_base + (float)PF_NREG / f.bk_inv : PF_INF_F);
	uint64_t *lst = pf_bucket_ptr(f, b), *spl = pf_bucket_ptr(f, PF_SPILL), *byd = pf_bucket_ptr(f, PF_BEYOND);
	const int n_b = pf_shfl_i(f.bk_n, b);
	const int use_sp = b < PF_SPILL && n_sp > 0 && f.spill_min < ub;
	const int use_by = b != PF_BEYOND && n_by > 0 && f.ovf_min < ub;
#ifdef PF_DIAG
	f.scanned += (unsigned)(2 * n_b + (use_sp ? 2 * n_sp : 0)); if (use_by) f.scanned_ovf += (unsigned)(2 * n_by);
#endif
	/* 3. minimum and window */
	uint64_t mk = pf_blist_min(lst, n_b);
	if (use_sp) { const uint64_t mo = pf_blist_min(spl, n_sp); if (mo < mk) mk = mo; }
	if (use_by) { const uint64_t mo = pf_blist_min(byd, n_by); if (mo < mk) mk = mo; }
	mk = pf_warp_min_u64(mk);
	const float m = pf_key_tot(mk);
	float win = m * f.win_rel;
	if (win < f.win_abs) win = f.win_abs;
	float T = m + win;
	if (T > ub) T = ub;                                   /* what lies beyond belongs to later buckets */
	for (;;) {
		int c = pf_blist_count(lst, n_b, T);
		if (use_sp) c += pf_blist_count(spl, n_sp, T);
		if (use_by) c += pf_blist_count(byd, n_by, T);
#ifdef PF_DIAG
		f.scanned += (unsigned)(n_b + (use_sp ? n_sp : 0)); if (use_by) f.scanned_ovf += (unsigned)n_by;
#endif
		c = pf_warp_sum_i(c);
		if (c <= PF_SH_REFILL || T <= m) break;
		win *= 0.25f;
		T = m + win;
	}
	/* 4. move, compact, track the far minimum: exact for what stays in the lists drawn from, lower bounds for the rest */
	int taken = 0, kept = 0;
	float fmin = pf_warp_min_f(pf_blist_take(f, lst, n_b, T, &taken, &kept));
	if (lane == b) f.bk_n = kept;
	if (b == PF_SPILL) f.spill_min = fmin;
	if (b == PF_BEYOND) f.ovf_min = fmin;
	if (use_sp) {
		f.spill_min = pf_warp_min_f(pf_blist_take(f, spl, n_sp, T, &taken, &kept));
		if (lane == PF_SPILL) f.bk_n = kept;
	}
	if (use_by) {
		f.ovf_min = pf_warp_min_f(pf_blist_take(f, byd, n_by, T, &taken, &kept));
		if (lane == PF_BEYOND) f.bk_n = kept;
	}
	{
		const unsigned above = pf_ballot(f.bk_n > 0 && lane > b && lane < PF_NREG);
		if (above) { const float lb = f.bk_base + (float)(pf_ffs(above) - 1) / f.bk_inv; if (lb < fmin) fmin = lb; }
		if (pf_shfl_i(f.bk_n, PF_SPILL) > 0 && f.spill_min < fmin) fmin = f.spill_min;
		if (pf_shfl_i(f.bk_n, PF_BEYOND) > 0 && f.ovf_min < fmin) fmin = f.ovf_min;
	}
	f.sh_n = taken;
	f.far_min = fmin;
	f.T_hi = T;
	pf_syncwarp();
}


/* the out-of-line refill: everything above is inlined into this one function */
PF_NOINLINE PfFront pf_refill_impl(PfFront f) { pf_refill_body(f); return f; }
/* BK = 0: the far list is ONE flat list over the whole array of the slot (the regular slots: their searches are A* searches
 * with short far lists, and the kernel of the throughput launches is bound by instruction fetch — it does not carry the bucket
 * code at all); BK = 1: the hybrid flat / bucketed list (the big slots, where the flooding searches end up). */
template <int BK> PF_DEV int pf_list_off(int far_cap) { return BK ? pf_stage_off(far_cap) : 0; }
template <int BK> PF_DEV int pf_list_cap(int far_cap) { return BK ? pf_stage_cap(far_cap) : far_cap; }
template <int BK> PF_DEV void pf_refill(PfWarp &w) {
	if (!BK || (!(w.round & PF_ROUND_BUCKETS) && w.st_n + w.sh_n <= PF_FLAT_MAX)) {
		/* the common case, inline and on the warp context itself: a short flat far list (the round-1 algorithm: three passes
		 * over a few dozen entries) — a call costs more than it does */
		const int lane = pf_lane();
		uint64_t *lst = w.far + pf_list_off<BK>(w.P->far_cap);
		const int cap = pf_list_cap<BK>(w.P->far_cap);
		w.refills++;
		PF_REFILL_LOOP for (int base = 0; base < w.sh_n; base += PF_WARP) {
			const int i = base + lane;
			if (i < w.sh_n && w.st_n + i < cap) lst[w.st_n + i] = w.fr[i];
		}
		w.st_n += w.sh_n;
		w.sh_n = 0;
		if (w.st_n > cap) { w.st_n = cap; w.overflow |= PF_OVF_OTHER; }
		pf_syncwarp();
		const float m = pf_key_tot(pf_warp_min_u64(pf_list_min(lst, w.st_n)));
		float win = m * w.P->win_rel;
		if (win < w.P->win_abs) win = w.P->win_abs;
		float T = m + win;
		for (;;) {
			const int c = pf_warp_sum_i(pf_list_count(lst, w.st_n, T));
			if (c <= PF_SH_REFILL || T <= m) break;
			win *= 0.25f;
			T = m + win;
		}
		int kept = 0, taken = 0;
		float fmin = PF_INF_F;
		PF_REFILL_LOOP for (int base = 0; base < w.st_n; base += PF_WARP) {
			const int i = base + lane;
			const uint64_t k = (i < w.st_n) ? lst[i] : PF_KEY_MAX;
			const int q = (i < w.st_n) && pf_key_tot(k) <= T;
			const unsigned mq = pf_ballot(q);
			const int rank = taken + pf_popc(mq & pf_lanemask_lt());
			const int take = q && rank < PF_SH_REFILL;
			const unsigned mt = pf_ballot(take);
			const int keep = (i < w.st_n) && !take;
			const unsigned mkp = pf_ballot(keep);
			pf_syncwarp();                                      /* reads of this chunk are complete before the in-place writes */
			if (take) w.fr[taken + pf_popc(mt & pf_lanemask_lt())] = k;
			if (keep) {
				lst[kept + pf_popc(mkp & pf_lanemask_lt())] = k;    /* kept + rank <= i: in-place is safe after the ballots */
				const float t = pf_key_tot(k);
				if (t < fmin) fmin = t;
			}
			taken += pf_popc(mt);
			kept += pf_popc(mkp);
		}
		w.sh_n = taken;
		w.st_n = kept;
		w.far_min = pf_warp_min_f(fmin);
		w.T_hi = T;
		pf_syncwarp();
		return;
	}
	if (BK) {
		/* a search that has outgrown the flat list: cost buckets, out of line; their state is parked in the slot's memory */
		PfFront f;
#ifdef PF_EMU
		if (pf_lane() == 0) pf_emu_bucket_refills++;
#endif
		float *bst = (float *)(w.far + w.P->far_cap - 32);
		const int open = (w.round & PF_ROUND_BUCKETS) != 0;
		f.fr = w.fr; f.far = w.far; f.far_cap = w.P->far_cap; f.win_rel = w.P->win_rel; f.win_abs = w.P->win_abs;
		f.sh_n = w.sh_n; f.st_n = w.st_n; f.far_min = w.far_min; f.T_hi = w.T_hi; f.overflow = 0; f.refills = 0;
#ifdef PF_DIAG
		f.scanned = 0; f.scanned_ovf = 0;
#endif
		f.bk_n = open ? ((const int *)bst)[pf_lane()] : 0;
		f.bk_base = open ? bst[32] : 0.f; f.bk_inv = open ? bst[33] : 0.f; f.ovf_min = open ? bst[34] : PF_INF_F; f.spill_min = open ? bst[35] : PF_INF_F;
		f = pf_refill_impl(f);
		((int *)bst)[pf_lane()] = f.bk_n;
		if (pf_lane() == 0) { bst[32] = f.bk_base; bst[33] = f.bk_inv; bst[34] = f.ovf_min; bst[35] = f.spill_min; }
		w.round |= PF_ROUND_BUCKETS;
		w.sh_n = f.sh_n; w.st_n = f.st_n; w.far_min = f.far_min; w.T_hi = f.T_hi; w.refills += f.refills;
#ifdef PF_DIAG
		w.stale += f.scanned; w.pushes += f.scanned_ovf;     /* diagnostic build: far-list entries touched by bucket refills: regular buckets / catch-all */
#endif
		if (f.overflow) w.overflow |= PF_OVF_OTHER;
		pf_syncwarp();
	}
}

PF_DEV void pf_far_reset(PfWarp &w) {
	w.st_n = 0; w.far_min = PF_INF_F;
	w.round &= ~PF_ROUND_BUCKETS;          /* buckets closed: pf_refill opens them when the far list outgrows PF_FLAT_MAX */
}

/* Warp-collective push.  Labels inside the near window go to shared memory, the rest (and any
 * near-set overflow) to the far list; far_min guards the best-first order. */
template <int BK> PF_DEV void pf_push(PfWarp &w, int valid, float tot, int node, int edge_start) {
	uint64_t key = pf_make_key(tot, node);
	int to_sh = valid && tot <= w.T_hi;
	/* a label inside the near window is settled soon, and the first thing read then is its row of out-edges: start
	 * pulling it into L2 (labels parked in the far list are mostly never settled — prefetching those too cost 40 %
	 * more DRAM traffic for nothing) */
	if (to_sh && edge_start >= 0) pf_prefetch_l2(w.P->edges + edge_start);
	unsigned m1 = pf_ballot(to_sh);
	int pos = w.sh_n + pf_popc(m1 & pf_lanemask_lt());
	if (to_sh && pos < PF_SH_FRONTIER) w.fr[pos] = key;
	int spill = to_sh && pos >= PF_SH_FRONTIER;
	int cnt = pf_popc(m1);
	w.sh_n = (w.sh_n + cnt < PF_SH_FRONTIER) ? w.sh_n + cnt : PF_SH_FRONTIER;
	int to_far = (valid && !to_sh) || spill;
	unsigned m2 = pf_ballot(to_far);
	if (m2) {
		const int so = pf_list_off<BK>(w.P->far_cap), scap = pf_list_cap<BK>(w.P->far_cap);
		const int fpos = w.st_n + pf_popc(m2 & pf_lanemask_lt());
		if (to_far && fpos < scap) w.far[so + fpos] = key;
		w.st_n += pf_popc(m2);
		if (w.st_n > scap) { w.st_n = scap; w.overflow |= PF_OVF_OTHER; }
		const float fm = pf_warp_min_f(to_far ? tot : PF_INF_F);
		if (fm < w.far_min) w.far_min = fm;
	}
#ifndef PF_DIAG
	w.pushes += (unsigned)pf_popc(pf_ballot(valid));
#endif
	pf_syncwarp();
}

/* ------------------------------------------------------------------ node record access */
struct PfNodeView { int xlow, ylow, xhigh, yhigh; float R, C; int occ; float acc; int edge_start, num_edges, type, ci, cap; };

PF_DEV PfNodeView pf_load_node(const PfParams *P, int v) {
	pf_u8 r = pf_ld_cg_u8(&P->nodes[v]);
	PfNodeView n;
	n.xlow = (short)(r.a & 0xffffu); n.ylow = (short)(r.a >> 16);
	n.xhigh = (short)(r.b & 0xffffu); n.yhigh = (short)(r.b >> 16);
	n.R = pf_int_as_float((int)r.c); n.C = pf_int_as_float((int)r.d);
	n.occ = (int)r.e; n.acc = pf_int_as_float((int)r.f); n.edge_start = (int)r.g;
	n.num_edges = (int)(r.h & 0xffffu);
	int tc = (int)((r.h >> 16) & 0xffu);
	n.type = tc & 7; n.ci = tc >> 3; n.cap = (int)(r.h >> 24);
	return n;
}

/* ------------------------------------------------------------------ sink order: heapsort (util/heapsort.c:13-96)
 * run by one lane so that equal criticalities come out in the reference's order. */
PF_DEV void pf_heapsort_lane(int *heap /*[1..n]*/, const float *v /*[1..n]*/, int n) {
	for (int i = 1; i <= n; i++) {
		unsigned ifrom = i, ito = ifrom / 2;
		heap[i] = i;
		while (ito >= 1 && v[heap[ifrom]] < v[heap[ito]]) {
			int t = heap[ito]; heap[ito] = heap[ifrom]; heap[ifrom] = t;
			ifrom = ito; ito = ifrom / 2;
		}
	}
	for (int tail = n; tail >= 1; tail--) {
		int smallest = heap[1];
		heap[1] = heap[tail];
		unsigned heap_end = tail - 1, ifrom = 1, ito = 2;
		while (ito <= heap_end) {
			if (v[heap[ito + 1]] < v[heap[ito]]) ito++;   /* reads the vacated slot like the reference */
			if (v[heap[ito]] > v[heap[ifrom]]) break;
			int t = heap[ito]; heap[ito] = heap[ifrom]; heap[ifrom] = t;
			ifrom = ito; ito = 2 * ifrom;
		}
		heap[tail] = smallest;
	}
}

/* ------------------------------------------------------------------ one sink search */
/* Returns 1 when the target was reached (label present), 0 if the frontier was exhausted, -1 on
 * scratch overflow.  tree_n = current number of tree entries. */
template <int STRICT, int BK> PF_DEV int pf_search_sink(PfWarp &w, int tree_n, int target_node, float crit, int rlim) {
	const PfParams *P = w.P;
	const int lane = pf_lane();
	const float astar = P->astar_fac;
	PfNodeView tn = pf_load_node(P, target_node);
	const int tgt_xl = tn.xlow, tgt_yl = tn.ylow;       /* lookahead uses xlow/ylow (route_timing.c:764) */
	const int tgt_xh = tn.xhigh, tgt_yh = tn.yhigh;     /* pruning uses xhigh/yhigh (route_timing.c:622) */
	const int highfan = w.num_sinks >= 64;

	w.epoch++;
	if ((w.epoch & w.tag_mask) == 0) {                  /* tag wrapped: wipe the hot table, skip tag 0 (= never written) */
		PF_COLD_LOOP for (unsigned i = (unsigned)lane; i <= w.label_mask; i += PF_WARP) w.hot[i] = 0;
		w.epoch++;
		pf_syncwarp();
	}
	w.n_labels = 0; w.sh_n = 0; w.far_min = PF_INF_F; w.best = PF_INF_F;
	/* delta-stepping bucket width: a multiple of the cheapest possible edge for this criticality
	 * (an uncongested wire: (1-crit)*base_cost + crit*T_linear) */
	float slack;
	{
		float mb = PF_INF_F, mt = PF_INF_F;
		if (P->pop_slack != 0.f) PF_COLD_LOOP for (int i = 4; i < P->num_indexed; i++) {
			if (w.base_cost[i] < mb) mb = w.base_cost[i];
			if (w.idx[i].T_linear < mt) mt = w.idx[i].T_linear;
		}
		slack = (P->pop_slack != 0.f) ? P->pop_slack * ((1.f - crit) * mb + crit * mt) : 0.f;
	}

	/* ---- seed with the current route tree (add_route_tree_to_heap): pass 0 finds the cheapest seed (it sets
	 * the near-set window), pass 1 labels and pushes.  One loop body for every pass keeps a single inlined copy of
	 * the lookahead — and makes every pass compute bit-identical totals.
	 * Big slots (BK), trees of PfParams.lazy_seed_min entries and more: LAZY seeding.  The reference pushes the whole tree for every
	 * sink; a seed whose total lies above the cost of the connection found is never popped, and a net of 1760 sinks whose tree
	 * holds 20 k nodes spends 35 M label insertions per iteration on such seeds (that one net was 10 s of the 13.4 s of the
	 * 32 k-LUT stand-in).  Only the seeds with total <= seed_hi are labelled; the cheapest seed left out stands in the far
	 * minimum, so the settle loop comes back here — through the test it makes anyway before every pop — before it could pop
	 * anything dearer than a seed it has not seen.  Same labels settled, same costs; only the age order of equal totals moves. */
	const bool lazy = BK && P->lazy_seed_min > 0 && tree_n >= P->lazy_seed_min;
	float seed_lo = -PF_INF_F, seed_hi = PF_INF_F, seed_min = PF_INF_F;
	int pass = 0;
	for (;;) {
	{
		float smin = PF_INF_F;      /* pass 0: the cheapest seed; later passes: the cheapest seed still left out */
		for (int base = 0; base < tree_n; base += PF_WARP) {
			int i = base + lane;
			int valid = 0, node = 0; float tot = 0.f, back = 0.f, R_up = 0.f;
			if (i < tree_n) {
				PfTreeNode t = w.tree[i];
				if (t.flags & PF_TF_REEXPAND) {
					valid = 1; node = t.node; R_up = t.R_up;
					back = crit * t.Tdel;
					tot = back + astar * pf_expected_cost(w, t.type_ci & 7, t.type_ci >> 3, t.xlow, t.xhigh, t.ylow, t.yhigh, tgt_xl, tgt_yl, crit, t.R_up);
				}
			}
			if (pass == 0) {
				if (valid && tot < smin) smin = tot;
			} else {
				if (lazy) {
					if (valid && tot > seed_hi && tot < smin) smin = tot;
					valid = valid && tot > seed_lo && tot <= seed_hi;
				}
				int wr = pf_label_relax(w, valid, node, tot, back, R_up, ~i, 0, -1);
				pf_push<BK>(w, wr, tot, node, -1);
				if (w.overflow) return -1;
			}
		}
		if (pass == 0 || lazy) smin = pf_warp_min_f(smin);
		if (pass == 0) {
			if (!(smin < PF_INF_F)) return 0;
			float win = smin * P->win_rel;
			if (win < P->win_abs) win = P->win_abs;
			w.T_hi = smin + win;
			pf_far_reset(w);
			if (lazy) seed_hi = smin + PF_LAZY_SEED_SPAN * win;
#ifdef PF_EMU
			if (lazy && lane == 0) pf_emu_lazy_seedings++;      /* low word: searches seeded lazily */
#endif
			pass = 1;
			continue;
		}
		if (lazy) { seed_min = smin; if (seed_min < w.far_min) w.far_min = seed_min; }
	}
	int more_seeds = 0;

	/* ---- settle loop */
	for (;;) {
		if (w.overflow) return -1;
		float mtot = PF_INF_F;
		int first = 0x7fffffff;                 /* this lane's oldest label at its local minimum */
#if defined(PF_TIE_NODE)
		/* A/B build: ties between equal totals are broken by the rr node id instead of by age — the minimum is then ONE 64-bit
		 * warp reduction over the keys and the popped entry is replaced by the last one (no order to preserve) */
		uint64_t my_key = PF_KEY_MAX;
		if (STRICT) {
#pragma unroll
			for (int c = 0; c < PF_SH_FRONTIER / PF_WARP; c++) {
				const int i = lane + c * PF_WARP;
				const uint64_t k = i < w.sh_n ? w.fr[i] : PF_KEY_MAX;
				if (k < my_key) { my_key = k; first = i; }
			}
			mtot = pf_key_tot(my_key);
		} else
#endif
		{	/* branch-free: the near set holds at most PF_SH_FRONTIER / 32 entries per lane; only the total (high word) is read */
			const uint32_t *hi = (const uint32_t *)w.fr + 1;
#pragma unroll
			for (int c = 0; c < PF_SH_FRONTIER / PF_WARP; c++) {
				const int i = lane + c * PF_WARP;
				const float t = i < w.sh_n ? pf_int_as_float((int)hi[2 * i]) : PF_INF_F;
				if (t < mtot) { mtot = t; first = i; }
			}
		}
		const float my_min = mtot;
#if defined(PF_TIE_NODE)
		const uint64_t min_key = STRICT ? pf_warp_min_u64(my_key) : 0;
		mtot = STRICT ? (w.sh_n > 0 ? pf_key_tot(min_key) : PF_INF_F) : pf_warp_min_f(mtot);
#else
		mtot = pf_warp_min_f(mtot);
#endif
		if (w.far_min < mtot) {                 /* a cheaper label sits in the far list (or near set empty) */
			if (w.far_min >= w.best) break;
			if (lazy && seed_min <= w.far_min) { more_seeds = 1; break; }      /* ... or among the seeds left out */
			pf_refill<BK>(w);
			if (lazy && seed_min < w.far_min) w.far_min = seed_min;
			continue;
		}
		if (w.sh_n == 0) break;                 /* both exhausted */
		if (mtot >= w.best) break;              /* target settled: nothing cheaper remains */

		const float thr = mtot + slack;
		int taken, M;
		/* the label(s) being expanded: uniform registers in strict mode, per-lane views of the batch otherwise */
		int x_node = -1, x_start = 0, x_type = 0; float x_back = 0.f, x_R = 0.f;
		if (STRICT) {
			/* -- strict best-first: settle one label, the OLDEST near label within pop_slack of the minimum (what the
			 * batch selection below yields for max_batch == 1), and close the gap so the near set stays in push
			 * order.  Every lane then does the same label lookup (broadcast reads): nothing is staged. */
#if defined(PF_TIE_NODE)
			const uint64_t mk = min_key;
			(void)my_min; (void)thr;
			{	/* the lane that holds the minimum (keys are unique: a node re-enters only with a lower total) moves the last
				 * entry into its place */
				const unsigned own = pf_ballot(my_key == mk);
				const int idx = pf_shfl_i(first, pf_ffs(own) - 1);
				if (lane == 0) w.fr[idx] = w.fr[w.sh_n - 1];
				w.sh_n--;
				pf_syncwarp();
			}
#else
			int li = 0x7fffffff;
			if (slack == 0.f) { if (my_min == mtot) li = first; }      /* within 0 of the minimum = at the minimum */
			else for (int i = lane; i < w.sh_n; i += PF_WARP) if (pf_key_tot(w.fr[i]) <= thr) { li = i; break; }
			const int idx = -pf_warp_max_i(-li);
			const uint64_t mk = w.fr[idx];
			uint64_t keep[PF_SH_FRONTIER / PF_WARP];
			for (int c = 0; c < PF_SH_FRONTIER / PF_WARP; c++) { int i = lane + c * PF_WARP; keep[c] = (i > idx && i < w.sh_n) ? w.fr[i] : 0; }
			pf_syncwarp();
			for (int c = 0; c < PF_SH_FRONTIER / PF_WARP; c++) { int i = lane + c * PF_WARP; if (i > idx && i < w.sh_n) w.fr[i - 1] = keep[c]; }
			w.sh_n--;
			pf_syncwarp();
#endif
			taken = 1; M = 0;
			const int u = pf_key_node(mk);
			const int h = pf_label_find(w, u);
			int ok = 0;
			if (h >= 0) {
				uint64_t hk = w.hot[h];
				if (pf_int_as_float((int)(hk >> 32)) == pf_key_tot(mk)) {      /* else stale: the node was re-labelled cheaper */
					pf_u4 a = pf_ld_u4(&w.cold[h]), b = pf_ld_u4((const char *)&w.cold[h] + 16);
					x_start = (int)b.x; x_type = (int)((a.w >> 8) & 7u);
					M = (int)(a.w >> 16);
					if (x_start < 0) {                                      /* seed: row not cached in the label */
						PfNodeView un = pf_load_node(P, u);
						x_start = un.edge_start; x_type = un.type; M = un.num_edges;
					}
					x_node = u; x_back = pf_int_as_float((int)a.x); x_R = pf_int_as_float((int)a.y);
					ok = 1;
				}
			}
			w.pops += (unsigned)ok;
			w.stale += (unsigned)(1 - ok);
			if (!ok) continue;
		} else {
		/* -- select the batch: every near label within pop_slack of the minimum, at most max_batch */
		int kept = 0;
		taken = 0;
		for (int base = 0; base < w.sh_n; base += PF_WARP) {
			int i = base + lane;
			uint64_t k = (i < w.sh_n) ? w.fr[i] : PF_KEY_MAX;
			int q = (i < w.sh_n) && pf_key_tot(k) <= thr;
			unsigned mq = pf_ballot(q);
			int rank = taken + pf_popc(mq & pf_lanemask_lt());
			int take = q && rank < P->max_batch;
			unsigned mt = pf_ballot(take);
			int keep = (i < w.sh_n) && !take;
			unsigned mkp = pf_ballot(keep);
			pf_syncwarp();                                  /* every lane has its entry before slots of this chunk are overwritten */
			if (take) w.b_key[taken + pf_popc(mt & pf_lanemask_lt())] = k;
			if (keep) w.fr[kept + pf_popc(mkp & pf_lanemask_lt())] = k;
			taken += pf_popc(mt);
			kept += pf_popc(mkp);
		}
		w.sh_n = kept;
		pf_syncwarp();

		/* -- validate each settled label; its row (start, degree, type) travels in the label */
		int deg = 0, ok = 0;
		if (lane < taken) {
			uint64_t k = w.b_key[lane];
			int u = pf_key_node(k);
			int h = pf_label_find(w, u);
			if (h >= 0) {
				uint64_t hk = w.hot[h];
				if (pf_int_as_float((int)(hk >> 32)) == pf_key_tot(k)) {   /* else stale: the node was re-labelled cheaper */
					pf_u4 a = pf_ld_u4(&w.cold[h]), b = pf_ld_u4((const char *)&w.cold[h] + 16);
					int es = (int)b.x, ty = (int)((a.w >> 8) & 7u);
					deg = (int)(a.w >> 16);
					if (es < 0) {                                  /* seed: row not cached in the label */
						PfNodeView un = pf_load_node(P, u);
						es = un.edge_start; ty = un.type; deg = un.num_edges;
					}
					w.b_node[lane] = u; w.b_back[lane] = pf_int_as_float((int)a.x); w.b_R[lane] = pf_int_as_float((int)a.y);
					w.b_start[lane] = es; w.b_type[lane] = ty;
					ok = 1;
				}
			}
			if (!ok) { w.b_node[lane] = -1; w.b_start[lane] = 0; w.b_type[lane] = 0; w.b_back[lane] = 0.f; w.b_R[lane] = 0.f; deg = 0; }
		}
		{
			int nok = pf_popc(pf_ballot(ok));
			w.pops += (unsigned)nok;
			w.stale += (unsigned)(taken - nok);
		}
		/* exclusive prefix of the degrees over the first `taken` lanes */
		int incl = deg;
		for (int d = 1; d < PF_WARP; d <<= 1) {
			int o = pf_shfl_i(incl, lane - d);
			if (lane >= d) incl += o;
		}
		if (lane < taken) w.b_pre[lane] = incl - deg;
		M = pf_shfl_i(incl, taken - 1);
		if (lane == 0) w.b_pre[taken] = M;
		pf_syncwarp();
		}
		w.visits += (unsigned)M;

		/* -- relax every out-edge of the batch, 32 edges per pass */
		for (int base = 0; base < M; base += PF_WARP) {
			int e = base + lane;
			int valid = e < M;
			int to = 0, u = 0, isw = 0, info = 0, es = 0; float tot = 0.f, back = 0.f, R_up = 0.f;
			if (valid) {
				int eoff = e;
				if (!STRICT) {
					int j = 0, hi = taken - 1;                  /* owner: last j with b_pre[j] <= e */
					while (j < hi) { int mid = (j + hi + 1) >> 1; if (w.b_pre[mid] <= e) j = mid; else hi = mid - 1; }
					x_node = w.b_node[j]; x_start = w.b_start[j]; x_type = w.b_type[j]; x_back = w.b_back[j]; x_R = w.b_R[j];
					eoff = e - w.b_pre[j];
				}
				u = x_node;
				uint32_t ew = P->edges[x_start + eoff];
				to = (int)(ew & pf_node_mask(PF_NB(P))); isw = (int)(ew >> PF_NB(P));
				PfNodeView n = pf_load_node(P, to);
				if (n.xhigh < w.bb_xmin || n.xlow > w.bb_xmax || n.yhigh < w.bb_ymin || n.ylow > w.bb_ymax) valid = 0;
				if (valid && highfan && (n.xhigh < tgt_xh - rlim || n.xlow > tgt_xh + rlim || n.yhigh < tgt_yh - rlim || n.ylow > tgt_yh + rlim)) valid = 0;
				if (valid && n.type == 2 && (n.xhigh != tgt_xh || n.yhigh != tgt_yh)) valid = 0;
				if (valid) {
					/* get_rr_cong_cost with the present cost derived from the live occupancy:
					 * pres = occ < cap ? 1 : 1 + (occ + 1 - cap) * pres_fac  (route_common.c:563-568) */
					float pres;
					if (n.occ < n.cap) pres = 1.;
					else pres = 1. + (n.occ + 1 - n.cap) * P->pres_fac;
					float cong = w.base_cost[n.ci] * n.acc * pres;
					float old_back = x_back, Ru = x_R;
					float new_back = old_back + (1. - crit) * cong;
					float new_R;
					const PfSwitchDev S = w.sw[isw];
					if (S.buffered) new_R = S.R; else new_R = Ru + S.R;
					if (crit != 0.f) {                       /* crit == 0: adds exactly +0 */
						float Tdel = n.C * (new_R + 0.5 * n.R);
						Tdel += S.Tdel;
						new_back += crit * Tdel;
					}
					new_R += n.R;
					if (P->bend_cost != 0.) {
						int ft = x_type;
						if ((ft == 4 && n.type == 5) || (ft == 5 && n.type == 4)) new_back += P->bend_cost;
					}
					tot = new_back + astar * pf_expected_cost(w, n.type, n.ci, n.xlow, n.xhigh, n.ylow, n.yhigh, tgt_xl, tgt_yl, crit, new_R);
					back = new_back; R_up = new_R;
					info = isw | (n.type << 8) | ((n.occ >= n.cap) << PF_INFO_SEEN_FULL_BIT) | (n.num_edges << 16); es = n.edge_start;
				}
			}
			int wr = pf_label_relax(w, valid, to, tot, back, R_up, u, info, es);
			/* the target SINK is never expanded; only its best total matters */
			float tb = pf_warp_min_f((wr && to == target_node) ? tot : PF_INF_F);
			if (tb < w.best) w.best = tb;
			pf_push<BK>(w, wr && to != target_node, tot, to, es);
			if (w.overflow) return -1;
		}
	}
	if (!more_seeds) break;
#ifdef PF_EMU
	if (lane == 0) pf_emu_lazy_seedings += 1ll << 32;   /* high word: searches that came back for more seeds */
#endif
	{	/* the next span of seeds, opened at the cheapest one left out */
		float win = seed_min * P->win_rel;
		if (win < P->win_abs) win = P->win_abs;
		seed_lo = seed_hi; seed_hi = seed_min + PF_LAZY_SEED_SPAN * win;
	}
	}
	return (w.best < PF_INF_F) ? 1 : 0;
}

/* ------------------------------------------------------------------ high-fanout window
 * mark_node_expansion_by_bin, route_timing.c:867-960: only the root's direct children are
 * examined and re-flagged (kept as in the reference). */
PF_DEV int pf_highfanout_rlim(PfWarp &w, int tree_n, int target_node) {
	const PfParams *P = w.P;
	const int lane = pf_lane();
	if (w.num_sinks < 64) return 1;
	PfNodeView tn = pf_load_node(P, target_node);
	int target_x = tn.xlow, target_y = tn.ylow;
	float area = (float)((w.bb_xmax - w.bb_xmin) * (w.bb_ymax - w.bb_ymin));
	if (area <= 0) area = 1;
	int rlim = (int)(pf_ceilf(pf_sqrtf(area / (float)w.num_sinks)));
	int maxdim = (P->nx + 2 > P->ny + 2) ? P->nx + 2 : P->ny + 2;
	/* does the root have children? */
	int has_child = 0;
	PF_COLD_LOOP for (int i = lane; i < tree_n; i += PF_WARP) if (i > 0 && w.tree[i].parent == 0) has_child = 1;
	if (!pf_any(has_child)) return maxdim;
	for (;;) {
		int hit = 0;
		for (int i = lane; i < tree_n; i += PF_WARP) {
			if (i > 0) {
				PfTreeNode t = w.tree[i];
				int ty = t.type_ci & 7;
				if (t.parent == 0 && !(ty == 2 || ty == 1)
						&& t.xlow <= target_x + rlim && t.xhigh >= target_x - rlim
						&& t.ylow <= target_y + rlim && t.yhigh >= target_y - rlim) hit = 1;
			}
		}
		if (pf_any(hit)) { rlim += 4; break; }
		if (rlim > maxdim) return -1;
		rlim *= 2;
	}
	for (int i = lane; i < tree_n; i += PF_WARP) {
		if (i > 0) {
			PfTreeNode t = w.tree[i];
			int ty = t.type_ci & 7;
			if (t.parent == 0 && !(ty == 2 || ty == 1)) {
				int in = t.xlow <= target_x + rlim && t.xhigh >= target_x - rlim && t.ylow <= target_y + rlim && t.yhigh >= target_y - rlim;
				w.tree[i].flags = (unsigned char)((t.flags & ~PF_TF_REEXPAND) | (in ? PF_TF_REEXPAND : 0));
			}
		}
	}
	pf_syncwarp();
	return rlim;
}

/* Warp-collective occupancy change: the atomic on the node record (pathfinder_update_one_cost's occ += / -=,
 * route_common.c:542-578) and, on several GPUs, an entry in this rank's event log. */
#define PF_EVENT_DEC 0x80000000u
PF_DEV int pf_occ_change(const PfParams *P, int active, int v, int d) {
	int old = 0;
	if (active) old = pf_atomic_add_i(&P->nodes[v].occ, d);
	if (P->events) {
		const unsigned m = pf_ballot(active);
		if (m) {
			const int leader = pf_ffs(m) - 1;
			unsigned long long base = 0;
			if (pf_lane() == leader) base = pf_atomic_add_ull(P->event_head, (unsigned long long)pf_popc(m));
			base = pf_shfl_u64(base, leader);
			if (active) {
				const unsigned long long at = base + (unsigned long long)pf_popc(m & pf_lanemask_lt());
				if ((long long)at < P->event_cap) P->events[at] = (unsigned)v | (d < 0 ? PF_EVENT_DEC : 0u);
			}
		}
	}
	return old;
}

/* Ripple re-routing: `victim` holds an rr node that the running net has just knowingly overfilled; queue it for a
 * re-route in this same iteration (once per iteration).  Per-lane, no collectives. */
PF_DEV void pf_queue_victim(const PfParams *P, int victim) {
	if (pf_atomic_exch_i(&P->queued[victim], P->iter_tag) == P->iter_tag) return;    /* already (to be) routed in this iteration */
	const int cls = P->net_big[victim] ? 1 : 0;
	const int pos = pf_atomic_add_i(&P->vq_ctl[2 * cls + 1], 1);
	if (pos < P->vq_cap) P->vq[cls][pos] = victim;
}

/* ------------------------------------------------------------------ back-trace + Elmore + commit
 * update_traceback (route_common.c:638) and update_route_tree (route_tree_timing.c:181-456).
 * Returns the tree index of the new SINK entry, or -1 on tree overflow.
 *
 * validate != 0: optimistic concurrency control.  The search priced every node with the occupancy it saw at that
 * moment; a net in flight at the same time may have committed the same node since.  The commit's atomic returns the
 * occupancy it found: if that is already at capacity although the search saw the node free (PF_INFO_SEEN_FULL in the
 * label), the serial router would have seen the other net's commit and priced the node as congested.  The path is
 * then taken back (-3) and the caller searches this sink again on the current occupancy — what the serial order
 * "other net first, then this one" would have produced.  Knowingly shared nodes (seen full) never trigger it. */
template <int RIP> PF_DEV int pf_add_path(PfWarp &w, int *tree_n_io, int target_node, int validate) {
	const PfParams *P = w.P;
	const int lane = pf_lane();
	int tree_n = *tree_n_io;
	int *pathbuf = w.iscratch + 3 * (P->sink_cap + 2);     /* reversed path: node ids, then switches */
	const int pcap = P->tree_cap;                          /* iscratch holds 2*tree_cap ints behind the sink arrays */
	int L = 0, join = -1;
	if (lane == 0) {
		int v = target_node;
		for (;;) {
			int h = pf_label_find(w, v);
			if (h < 0) { L = -1; break; }
			pf_u4 b = pf_ld_u4(&w.cold[h]);
			int prev = (int)b.z;
			if (prev < 0 && v != target_node) { join = ~prev; break; }
			if (prev < 0) { L = -1; break; }               /* target itself is a seed: cannot happen (SINKs are never seeds) */
			if (L >= pcap) { L = -2; break; }
			pathbuf[L] = v; pathbuf[pcap + L] = (int)(b.w & 0xfffu);  /* switch used to enter v | seen-full flag (bit 11) */
			L++;
			v = prev;
		}
	}
	L = pf_shfl_i(L, 0); join = pf_shfl_i(join, 0);
	if (L < 0) { if (L == -1 && lane == 0) pf_atomic_or_i(P->status, PF_ST_INTERNAL); return -1; }
	if (tree_n + L > P->tree_cap) return -1;
	pf_syncwarp();
	/* materialise the entries in path order (join's child first, SINK last) */
	int raced = 0;
	for (int base = 0; base < L; base += PF_WARP) {
		const int i = base + lane;
		int v = 0, cap = 0x7fffffff, fcap = 0x7fffffff;
		if (i < L) {
			v = pathbuf[L - 1 - i];
			PfNodeView n = pf_load_node(P, v);
			PfTreeNode t;
			t.node = v; t.parent = (i == 0) ? join : tree_n + i - 1;
			t.R_up = n.R; t.C_down = n.C; t.Tdel = 0.f;     /* R_up/C_down temporarily hold the node's own R and C */
			t.xlow = (short)n.xlow; t.ylow = (short)n.ylow; t.xhigh = (short)n.xhigh; t.yhigh = (short)n.yhigh;
			t.sw = (unsigned char)(pathbuf[pcap + L - 1 - i] & 0xff);
			if (!((pathbuf[pcap + L - 1 - i] >> PF_INFO_SEEN_FULL_BIT) & 1)) cap = n.cap;   /* seen free: a full node now is a race */
			t.type_ci = (unsigned char)(n.type | (n.ci << 3));
			t.flags = (n.type == 2 || n.type == 1) ? 0 : PF_TF_REEXPAND;   /* IPIN / SINK are not re-expanded */
			t.pad = 0;
			w.tree[tree_n + i] = t;
			if (RIP && P->committer) fcap = n.cap;
		}
		const int old = pf_occ_change(P, i < L, v, 1);  /* commit: pathfinder_update_one_cost(+1) */
		if (validate) raced |= (i < L && old >= cap);
		if (RIP && i < L && old >= fcap) pathbuf[pcap + L - 1 - i] |= 1 << 12;   /* now overfull: its holder becomes a victim below */
	}
	if (validate && pf_any(raced)) {
		PF_COLD_LOOP for (int base = 0; base < L; base += PF_WARP) {
			const int i = base + lane;
			pf_occ_change(P, i < L, i < L ? pathbuf[L - 1 - i] : 0, -1);
		}
		pf_syncwarp();
		return -3;
	}
	if (RIP && P->committer) {
		/* this net now holds the path; whoever held a node that is overfull now is re-routed in this iteration */
		PF_COLD_LOOP for (int base = 0; base < L; base += PF_WARP) {
			const int i = base + lane;
			if (i < L) {
				const int prev = pf_atomic_exch_i(&P->committer[pathbuf[L - 1 - i]], w.cur_net);
				if (P->vq_ctl && ((pathbuf[pcap + L - 1 - i] >> 12) & 1) && prev >= 0 && prev != w.cur_net) pf_queue_victim(P, prev);
			}
		}
	}
	pf_syncwarp();
	if (lane == 0) {
		PfTreeNode *T = w.tree;
		const int a = tree_n, z = tree_n + L - 1;
		/* node R of each new entry is parked in R_up, node C in C_down */
		/* C_downstream, sink → join (add_path_to_route_tree :286-297) */
		float Cd = T[z].C_down;                          /* C of the SINK */
		/* keep the per-node R and C while rewriting: walk backwards for C_down */
		float Cnode;
		for (int i = z - 1; i >= a; i--) {
			Cnode = T[i].C_down;
			int isw = T[i + 1].sw;
			if (!w.sw[isw].buffered) Cd += Cnode; else Cd = Cnode;
			T[i].C_down = Cd;
		}
		/* R_upstream, join → sink (load_new_path_R_upstream :340-390) */
		{
			int isw = T[a].sw;
			float Rn = T[a].R_up;
			float Ru = w.sw[isw].R + Rn;
			if (!w.sw[isw].buffered) Ru += T[join].R_up;
			T[a].R_up = Ru;
			/* Tdel needs the node R afterwards: stash it in Tdel for now */
			T[a].Tdel = Rn;
			for (int i = a + 1; i <= z; i++) {
				isw = T[i].sw;
				Rn = T[i].R_up;
				if (w.sw[isw].buffered) Ru = w.sw[isw].R + Rn; else Ru += w.sw[isw].R + Rn;
				T[i].R_up = Ru;
				T[i].Tdel = Rn;
			}
		}
		/* ancestors (update_unbuffered_ancestors_C_downstream :393-417) */
		int rt = a;
		{
			float C_add = T[a].C_down;
			int parent = T[a].parent, isw = T[a].sw;
			while (parent != -1 && !w.sw[isw].buffered) {
				rt = parent;
				T[rt].C_down += C_add;
				parent = T[rt].parent;
				isw = T[rt].sw;
			}
		}
		/* Tdel of the affected subtree (load_rt_subtree_Tdel :419-456) */
		if (rt == a) {
			/* the usual case: a buffered switch isolates the new branch; only its own chain changes */
			float Tarr;
			{
				int isw = T[a].sw;
				Tarr = T[join].Tdel;
				Tarr += w.sw[isw].R * T[a].C_down;
				Tarr += w.sw[isw].Tdel;
			}
			for (int i = a; i <= z; i++) {
				float Rn = T[i].Tdel;                    /* stashed node R */
				float Td = Tarr + 0.5 * T[i].C_down * Rn;
				T[i].Tdel = Td;
				if (i < z) {
					int isw = T[i + 1].sw;
					Tarr = Td + w.sw[isw].R * T[i + 1].C_down;
					Tarr += w.sw[isw].Tdel;
				}
			}
		} else {
			/* generic sweep: entries are in topological order, so one forward pass over the
			 * descendants of rt (marked on the fly) recomputes every affected Tdel */
			float Tstart;
			if (T[rt].parent != -1) {
				int isw = T[rt].sw;
				Tstart = T[T[rt].parent].Tdel;
				Tstart += w.sw[isw].R * T[rt].C_down;
				Tstart += w.sw[isw].Tdel;
			} else {
				Tstart = 0.;
			}
			for (int i = rt; i <= z; i++) {
				int affected = (i == rt) || (T[i].parent >= rt && (T[T[i].parent].flags & PF_TF_MARK));
				if (!affected) continue;
				float Rn = (i >= a) ? T[i].Tdel : pf_load_node(P, T[i].node).R;
				float Tarr;
				if (i == rt) Tarr = Tstart;
				else {
					int isw = T[i].sw;
					Tarr = T[T[i].parent].Tdel + w.sw[isw].R * T[i].C_down;
					Tarr += w.sw[isw].Tdel;
				}
				T[i].Tdel = Tarr + 0.5 * T[i].C_down * Rn;
				T[i].flags |= PF_TF_MARK;
			}
			for (int i = rt; i <= z; i++) T[i].flags &= (unsigned char)~PF_TF_MARK;
		}
	}
	pf_syncwarp();
	*tree_n_io = tree_n + L;
	return tree_n + L - 1;
}

/* ------------------------------------------------------------------ breadth-first router (route_breadth_first.c)
 * breadth_first_route_net :93-171: ONE maze wavefront per net (no lookahead, no delay term).  The SOURCE enters the
 * frontier at its congestion cost; whenever an unreached SINK of the net is settled its path joins the tree and its
 * rr nodes re-enter the frontier at cost 0 (breadth_first_expand_trace_segment :173-257), and the same wave carries
 * on.  Labels persist for the whole net.  Returns 1 (all sinks connected), 0 (frontier exhausted: no path),
 * -1 (scratch overflow: retry in a bigger slot), -2 (two pins of the net on one SINK: not supported). */
template <int RIP, int BK> PF_DEV int pf_route_wave_bf(PfWarp &w, int *tree_n_io, int t0, int ns, int *sink_done, int *rt_of_sink) {
	const PfParams *P = w.P;
	const int lane = pf_lane();
	w.epoch++;
	if ((w.epoch & w.tag_mask) == 0) {
		PF_COLD_LOOP for (unsigned i = (unsigned)lane; i <= w.label_mask; i += PF_WARP) w.hot[i] = 0;
		w.epoch++;
		pf_syncwarp();
	}
	w.n_labels = 0; w.sh_n = 0; w.far_min = PF_INF_F; w.best = PF_INF_F;
	for (int k = 1 + lane; k <= ns; k += PF_WARP) sink_done[k] = 0;
	{	/* breadth_first_add_source_to_heap :294-305 */
		const int src = w.tree[0].node;
		PfNodeView n = pf_load_node(P, src);
		float pres;
		if (n.occ < n.cap) pres = 1.; else pres = 1. + (n.occ + 1 - n.cap) * P->pres_fac;
		const float c = w.base_cost[n.ci] * n.acc * pres;
		float win = c * P->win_rel;
		if (win < P->win_abs) win = P->win_abs;
		w.T_hi = c + win;
		pf_far_reset(w);
		int wr = pf_label_relax(w, lane == 0, src, c, c, 0.f, ~0, 0, -1);
		pf_push<BK>(w, wr, c, src, -1);
	}
	int remaining = ns;
	while (remaining > 0) {
		if (w.overflow) return -1;
		float mtot = PF_INF_F;
		int first = 0x7fffffff;
		for (int i = lane; i < w.sh_n; i += PF_WARP) { float t = pf_key_tot(w.fr[i]); if (t < mtot) { mtot = t; first = i; } }
		const float my_min = mtot;
		mtot = pf_warp_min_f(mtot);
		if (w.far_min < mtot) { pf_refill<BK>(w); continue; }
		if (w.sh_n == 0) return 0;                          /* heap empty: "no possible path" :124 */
		const int idx = -pf_warp_max_i(my_min == mtot ? -first : -0x7fffffff);
		const uint64_t mk = w.fr[idx];
		uint64_t keep[PF_SH_FRONTIER / PF_WARP];
		for (int c = 0; c < PF_SH_FRONTIER / PF_WARP; c++) { int i = lane + c * PF_WARP; keep[c] = (i > idx && i < w.sh_n) ? w.fr[i] : 0; }
		pf_syncwarp();
		for (int c = 0; c < PF_SH_FRONTIER / PF_WARP; c++) { int i = lane + c * PF_WARP; if (i > idx && i < w.sh_n) w.fr[i - 1] = keep[c]; }
		w.sh_n--;
		pf_syncwarp();
		/* the settled label (every lane does the same lookup) */
		const int u = pf_key_node(mk);
		const int h = pf_label_find(w, u);
		if (h < 0) { w.stale++; continue; }
		if (pf_int_as_float((int)(w.hot[h] >> 32)) != pf_key_tot(mk)) { w.stale++; continue; }   /* re-labelled cheaper */
		pf_u4 a = pf_ld_u4(&w.cold[h]), b = pf_ld_u4((const char *)&w.cold[h] + 16);
		int x_start = (int)b.x, x_type = (int)((a.w >> 8) & 7u), M = (int)(a.w >> 16);
		const float x_back = pf_int_as_float((int)a.x);
		if (x_start < 0) { PfNodeView un = pf_load_node(P, u); x_start = un.edge_start; x_type = un.type; M = un.num_edges; }
		w.pops++;

		if (x_type == 1) {                                  /* a SINK: one of ours, still unconnected? (target_flag, :132) */
			int pin = 0x7fffffff, dup = 0;
			for (int k = 1 + lane; k <= ns; k += PF_WARP) if (!sink_done[k] && P->net_term[t0 + k] == u) { dup++; if (k < pin) pin = k; }
			pin = -pf_warp_max_i(-pin);
			dup = pf_warp_sum_i(dup);
			if (dup > 1) return -2;
			if (dup == 1) {
				const int a0 = *tree_n_io;
				const int si = pf_add_path<RIP>(w, tree_n_io, u, 0);
				if (si < 0) { w.overflow |= PF_OVF_OTHER; return -1; }
				if (lane == 0) { rt_of_sink[pin] = si; sink_done[pin] = 1; }
				pf_syncwarp();
				remaining--;
				/* the new segment re-enters the frontier at cost 0 */
				for (int base = a0; base <= si; base += PF_WARP) {
					const int i = base + lane;
					const int valid = i <= si;
					const int node = valid ? w.tree[i].node : 0;
					int wr = pf_label_relax(w, valid, node, 0.f, 0.f, 0.f, ~i, 0, -1);
					pf_push<BK>(w, wr, 0.f, node, -1);
					if (w.overflow) return -1;
				}
				continue;
			}
		}
		/* breadth_first_expand_neighbours :259-292 */
		w.visits += (unsigned)M;
		for (int base = 0; base < M; base += PF_WARP) {
			const int e = base + lane;
			int valid = e < M, to = 0, info = 0, es = 0;
			float tot = 0.f;
			if (valid) {
				const uint32_t ew = P->edges[x_start + e];
				to = (int)(ew & pf_node_mask(PF_NB(P)));
				const int isw = (int)(ew >> PF_NB(P));
				PfNodeView n = pf_load_node(P, to);
				if (n.xhigh < w.bb_xmin || n.xlow > w.bb_xmax || n.yhigh < w.bb_ymin || n.ylow > w.bb_ymax) valid = 0;
				if (valid) {
					float pres;
					if (n.occ < n.cap) pres = 1.; else pres = 1. + (n.occ + 1 - n.cap) * P->pres_fac;
					tot = x_back + w.base_cost[n.ci] * n.acc * pres;
					if (P->bend_cost != 0.) { if ((x_type == 4 && n.type == 5) || (x_type == 5 && n.type == 4)) tot += P->bend_cost; }
					info = isw | (n.type << 8) | (n.num_edges << 16); es = n.edge_start;
				}
			}
			int wr = pf_label_relax(w, valid, to, tot, tot, 0.f, u, info, es);
			pf_push<BK>(w, wr, tot, to, es);
			if (w.overflow) return -1;
		}
	}
	return 1;
}

PF_DEV void pf_swap_tables(PfWarp &w) {
	uint64_t *h = w.hot; w.hot = w.hot_alt; w.hot_alt = h;
	PfCold *c = w.cold; w.cold = w.cold_alt; w.cold_alt = c;
	unsigned m = w.label_mask; w.label_mask = w.mask_alt; w.mask_alt = m;
	int s = w.label_shift; w.label_shift = w.shift_alt; w.shift_alt = s;
	int l = w.label_limit; w.label_limit = w.limit_alt; w.limit_alt = l;
	unsigned e = w.epoch; w.epoch = w.epoch_alt; w.epoch_alt = e;
}

/* ------------------------------------------------------------------ one net
 * timing_driven_route_net, route_timing.c:399-563 */
/* STRICT: 0 = delta buckets, 1 = strict best-first, 2 = breadth-first router (always strict order) */
template <int STRICT, int RIP, int BK> PF_DEV int pf_route_net(PfWarp &w, int inet, int ripup) {   /* 1: routed, 0: handed to a bigger slot / failed */
	const PfParams *P = w.P;
	const int lane = pf_lane();
	const int t0 = P->net_ptr[inet];
	const int ns = P->net_ptr[inet + 1] - t0 - 1;
	w.num_sinks = ns; w.cur_net = inet;
	w.bb_xmin = P->net_bb[4 * inet + 0]; w.bb_xmax = P->net_bb[4 * inet + 1];
	w.bb_ymin = P->net_bb[4 * inet + 2]; w.bb_ymax = P->net_bb[4 * inet + 3];
	w.overflow = 0;

	/* rip-up: pathfinder_update_one_cost(trace_head[inet], -1) — one atomic per tree entry */
	if (ripup) {
		PfNetLoc loc = P->loc[inet];
		PF_COLD_LOOP for (int base = 0; base < loc.count; base += PF_WARP) {
			const int i = base + lane;
			const int v = i < loc.count ? P->pool_node[loc.off + i] : 0;
			pf_occ_change(P, i < loc.count, v, -1);
			if (RIP && P->committer && i < loc.count) pf_atomic_cas_i(&P->committer[v], inet, -1);   /* no longer the holder */
		}
	}
	if (ns > P->sink_cap) { w.overflow = PF_OVF_OTHER; }

	float *pin_crit = (float *)w.iscratch;                   /* [1..ns] */
	int *sink_order = w.iscratch + (P->sink_cap + 2);        /* [1..ns] */
	int *rt_of_sink = w.iscratch + 2 * (P->sink_cap + 2);    /* [1..ns] */
	int tree_n = 0;
	int fail = 0;
	if (!w.overflow) {
		/* pin criticalities, route_timing.c:424-454 */
		for (int ipin = 1 + lane; ipin <= ns; ipin += PF_WARP) {
			float c = P->crit[t0 + ipin];
			double d = c - (1.0 - P->max_crit);
			c = (d > 0.0) ? d : 0.0;
			if (P->crit_exp != 1.0f) c = pf_powf(c, P->crit_exp);
			if (c > P->max_crit) c = P->max_crit;
			pin_crit[ipin] = c;
		}
		pf_syncwarp();
		if (lane == 0) pf_heapsort_lane(sink_order, pin_crit, ns);
		/* update_rr_base_costs, route_timing.c:842-864 */
		if (lane < P->num_indexed) {
			float bc = w.idx[lane].base_cost;
			if (lane >= 4) {
				float factor = pf_sqrtf((float)ns);
				bc = (w.idx[lane].T_quadratic > 0.) ? w.idx[lane].saved_base_cost * factor : w.idx[lane].saved_base_cost;
			}
			w.base_cost[lane] = bc;
		}
		/* init_route_tree_to_source, route_tree_timing.c:155-178 */
		if (lane == 0) {
			int src = P->net_term[t0];
			PfNodeView n = pf_load_node(P, src);
			PfTreeNode t;
			t.node = src; t.parent = -1; t.R_up = n.R; t.C_down = n.C; t.Tdel = 0.5 * n.R * n.C;
			t.xlow = (short)n.xlow; t.ylow = (short)n.ylow; t.xhigh = (short)n.xhigh; t.yhigh = (short)n.yhigh;
			t.sw = 0; t.type_ci = (unsigned char)(n.type | (n.ci << 3)); t.flags = PF_TF_REEXPAND; t.pad = 0;
			w.tree[0] = t;
		}
		pf_occ_change(P, lane == 0, P->net_term[t0], 1);     /* the SOURCE is the head of the first trace segment */
		tree_n = 1;
		pf_syncwarp();

		if (STRICT == 2) {
			/* breadth-first: no per-net base-cost rescaling (route_breadth_first.c never calls update_rr_base_costs),
			 * and the wave floods the bounding box, so it runs on the slot's table in global memory */
			if (lane < P->num_indexed) w.base_cost[lane] = w.idx[lane].base_cost;
			pf_syncwarp();
			const int swapped = w.hot_alt != NULL;
			if (swapped) pf_swap_tables(w);
			const int r = pf_route_wave_bf<RIP, BK>(w, &tree_n, t0, ns, sink_order /* reused: per-pin done flags */, rt_of_sink);
			if (swapped) pf_swap_tables(w);
			if (r == 0) fail = PF_ST_UNROUTABLE;
			else if (r == -2) fail = PF_ST_TWICE_TO_SINK_BF;
			else if (r < 0 && !w.overflow) w.overflow = PF_OVF_OTHER;
		} else
		for (int itarget = 1; itarget <= ns; itarget++) {
			int target_pin = sink_order[itarget];
			int target_node = P->net_term[t0 + target_pin];
			float crit = pin_crit[target_pin];
			int rlim = pf_highfanout_rlim(w, tree_n, target_node);
			if (rlim < 0) { fail = PF_ST_INTERNAL; break; }
			/* a search that outgrows the shared-memory label table runs again on this slot's fallback table in
			 * global memory; the tables are swapped back before the next sink.  (One call site each for the search
			 * and the back-trace: the kernel's instruction footprint matters, see DESIGN.md.) */
#ifdef PF_NO_VALIDATE
			int r, swapped = 0, si = 0, tries = 0;           /* A/B build: commit validation compiled out */
#else
			int r, swapped = 0, si = 0, tries = P->validate;
#endif
			for (;;) {
				r = pf_search_sink<STRICT == 2 ? 1 : STRICT, BK>(w, tree_n, target_node, crit, rlim);
				if (r < 0 && !swapped && w.overflow == PF_OVF_LABELS && w.hot_alt) {
					pf_swap_tables(w);
					w.overflow = 0;
					swapped = 1;
					continue;
				}
				if (r > 0) {
					si = pf_add_path<RIP>(w, &tree_n, target_node, tries > 0);
					if (si == -3) { tries--; w.races++; continue; }      /* lost a race for a node: search again on the current occupancy */
				}
				break;
			}
			if (swapped) pf_swap_tables(w);
			if (r < 0) break;                                 /* overflow: retry in a bigger slot */
			if (r == 0) { fail = PF_ST_UNROUTABLE; break; }
			if (si < 0) { w.overflow = PF_OVF_OTHER; break; }
			if (lane == 0) rt_of_sink[target_pin] = si;
			pf_syncwarp();
		}
	}

	if (w.overflow || fail) {
		/* undo this net's commits; it owns no routing until it is retried */
		PF_COLD_LOOP for (int base = 0; base < tree_n; base += PF_WARP) {
			const int i = base + lane;
			const int v = i < tree_n ? w.tree[i].node : 0;
			pf_occ_change(P, i < tree_n, v, -1);
			if (RIP && P->committer && i < tree_n) pf_atomic_cas_i(&P->committer[v], inet, -1);
		}
		if (lane == 0) {
			P->loc[inet].off = 0; P->loc[inet].count = 0;
			if (fail) { pf_atomic_add_i(P->status + 1, 1); pf_atomic_or_i(P->status, fail); P->status[2] = inet; }
			else if (P->hot) { pf_atomic_or_i(P->status, PF_ST_BIG_OVERFLOW); P->status[2] = inet; }   /* already in a big slot */
			else { int k = pf_atomic_add_i(P->retry_count, 1); P->retry_list[k] = inet; if (P->net_big) P->net_big[inet] = 1; }
		}
		pf_syncwarp();
		return 0;
	}
	/* update_net_delays_from_route_tree, route_tree_timing.c:515-528 */
	for (int ipin = 1 + lane; ipin <= ns; ipin += PF_WARP) P->net_delay[t0 + ipin] = w.tree[rt_of_sink[ipin]].Tdel;
	/* publish the tree in the route store */
	unsigned long long off = 0;
	if (lane == 0) off = pf_atomic_add_ull(P->pool_head, (unsigned long long)tree_n);
	off = pf_shfl_u64(off, 0);
	if ((long long)(off + tree_n) > P->pool_cap) {
		if (lane == 0) { pf_atomic_or_i(P->status, PF_ST_POOL_OVERFLOW); P->loc[inet].off = 0; P->loc[inet].count = 0; }
		PF_COLD_LOOP for (int base = 0; base < tree_n; base += PF_WARP) { const int i = base + lane; pf_occ_change(P, i < tree_n, i < tree_n ? w.tree[i].node : 0, -1); }
		pf_syncwarp();
		return 0;
	}
	for (int i = lane; i < tree_n; i += PF_WARP) { P->pool[off + i] = w.tree[i]; P->pool_node[off + i] = w.tree[i].node; }
	if (lane == 0) { P->loc[inet].off = (int)off; P->loc[inet].count = tree_n; }
	pf_syncwarp();
	return 1;
}

/* ------------------------------------------------------------------ warp main: persistent work loop */
template <int STRICT, int RIP, int BK> PF_DEV void pf_warp_main(const PfParams *P, int slot, PfIndexedDev *idx_tab, PfSwitchDev *sw_tab, unsigned char *smem_warp) {
	const int lane = pf_lane();
	PfWarp w;
	w.P = P;
	unsigned char *s = smem_warp;
	w.fr = (uint64_t *)s; s += PF_SH_FRONTIER * 8;
	w.b_key = (uint64_t *)s; s += PF_MAX_BATCH * 8;
	w.idx = idx_tab; w.sw = sw_tab;               /* filled by the caller (one copy per CTA) */
	w.base_cost = (float *)s; s += PF_MAX_INDEXED * 4;
	w.b_node = (int *)s; s += PF_MAX_BATCH * 4;
	w.b_back = (float *)s; s += PF_MAX_BATCH * 4;
	w.b_R = (float *)s; s += PF_MAX_BATCH * 4;
	w.b_start = (int *)s; s += PF_MAX_BATCH * 4;
	w.b_type = (int *)s; s += PF_MAX_BATCH * 4;
	w.b_pre = (int *)s; s += (PF_MAX_BATCH + 1 + 1) * 4;
	w.ticket = (int *)s; s += PF_TICKETS * 4;
	const long long cap = 1ll << P->label_log2;
	w.cold = P->cold + (long long)slot * cap;
	w.label_mask = (unsigned)(cap - 1);
	w.label_shift = 32 - P->label_log2;
	if (P->hot) {                                     /* big-net slots: hot table in global memory, half full at most */
		w.hot = P->hot + (long long)slot * cap;
		w.label_limit = (int)(cap >> 1);
	} else {                                          /* regular slots: hot table in shared memory, 3/4 full at most */
		w.hot = (uint64_t *)s; s += PF_SMEM_HOT_ENTRIES * 8;
		w.label_limit = (int)(cap - (cap >> 2));
		for (int i = lane; i < PF_SMEM_HOT_ENTRIES; i += PF_WARP) w.hot[i] = 0;
	}
	w.hot_alt = NULL; w.cold_alt = NULL; w.mask_alt = 0; w.shift_alt = 0; w.limit_alt = 0; w.epoch_alt = 0;
	if (!P->hot && P->hot2) {
		const long long cap2 = 1ll << P->label2_log2;
		w.hot_alt = P->hot2 + (long long)slot * cap2; w.cold_alt = P->cold2 + (long long)slot * cap2;
		w.mask_alt = (unsigned)(cap2 - 1); w.shift_alt = 32 - P->label2_log2; w.limit_alt = (int)(cap2 >> 1);
		w.epoch_alt = P->epochs[2 * slot + 1];
	}
	w.tree = P->tree + (long long)slot * P->tree_cap;
	w.far = P->far + (long long)slot * P->far_cap;
	w.iscratch = P->iscratch + (long long)slot * (3 * (P->sink_cap + 2) + 2 * P->tree_cap);
	/* the shared-memory table starts empty at every launch; a global table keeps its tags across launches */
	w.epoch = P->hot ? P->epochs[2 * slot] : 0;
	w.round = 0;
	w.nb = PF_NB(P); w.tag_mask = pf_tag_mask(P);
	for (int i = lane; i < PF_TICKETS; i += PF_WARP) w.ticket[i] = 0x7fffffff;
	w.pops = w.pushes = w.visits = w.refills = w.stale = w.races = 0; w.max_net_pops = 0;
	w.n_labels = 0; w.sh_n = 0; w.T_hi = 0.f; w.far_min = PF_INF_F; w.best = PF_INF_F; w.overflow = 0;
	w.st_n = 0;
	w.bb_xmin = w.bb_xmax = w.bb_ymin = w.bb_ymax = 0; w.num_sinks = 0;
	pf_syncwarp();
	unsigned long long nets = 0;
	/* work: this launch's list first, then — with ripple re-routing — the victim queue of the slot class, until no warp is
	 * routing any more (only a routing warp can produce victims) */
	int *const vctl = RIP ? P->vq_ctl : NULL;      /* RIP == 0: the kernel variant without any ripple code (big iterations) */
	int *const vq = vctl ? P->vq[P->vq_class] : NULL;
	for (;;) {
		int net = -1, ripup = !P->skip_ripup;
		if (lane == 0) {
			if (vctl) { pf_atomic_add_i(&vctl[4], 1); pf_threadfence(); }      /* counted as routing before the claim is visible */
			const int k = pf_atomic_add_i(P->work_head, 1);
			const int nw = P->num_work_ptr ? pf_ld_volatile_i(P->num_work_ptr) : P->num_work;
			if (k < nw) net = P->work[k];
			else if (vctl) { pf_atomic_add_i(&vctl[4], -1); net = -2; }
		}
		net = pf_shfl_i(net, 0);
		while (net == -2) {                                    /* collective spin: one shuffle per turn */
			int got = -2;
			if (lane == 0) {
				const int routing = pf_ld_volatile_i(&vctl[4]);
				pf_threadfence();
				const int h = pf_ld_volatile_i(&vctl[2 * P->vq_class]);
				int t = pf_ld_volatile_i(&vctl[2 * P->vq_class + 1]);
				if (t > P->vq_cap) t = P->vq_cap;
				if (h < t) {
					pf_atomic_add_i(&vctl[4], 1);
					if (pf_atomic_cas_i(&vctl[2 * P->vq_class], h, h + 1) == h) {
						int v;
						while ((v = pf_ld_volatile_i(&vq[h])) < 0) pf_spin_pause();   /* the producer's store is a few cycles behind its ticket */
						vq[h] = -1;
						got = v;
					} else pf_atomic_add_i(&vctl[4], -1);
				} else if (routing == 0) got = -1;
			}
			net = pf_shfl_i(got, 0);
			if (net == -2) pf_spin_pause();
			else if (net >= 0) ripup = 1;                       /* a victim still owns its old route */
		}
		if (net < 0) break;
		const unsigned pops0 = BK ? w.pops : 0u;           /* (the largest search of the launch: a diagnostic of the big slots only) */
#ifdef PF_DIAG
		unsigned long long t0_ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0_ns));
		const unsigned st0 = w.stale, rf0 = w.refills;
#endif
		nets += (unsigned long long)pf_route_net<STRICT, RIP, BK>(w, net, ripup);
		if (BK && w.pops - pops0 > w.max_net_pops) w.max_net_pops = w.pops - pops0;
#ifdef PF_DIAG
		if (BK && lane == 0) {      /* diagnostic build: 'races' = the slowest net of the iteration: ms << 44 | net << 24 | refills (saturated) */
			unsigned long long t1_ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_ns));
			unsigned rf = w.refills - rf0; if (rf > 0xffffffu) rf = 0xffffffu;
			pf_atomic_max_ull(&P->stats->races, ((t1_ns - t0_ns) / 1000000ull) << 44 | (unsigned long long)(unsigned)net << 24 | rf);
			(void)st0;
		}
#endif
		if (vctl && lane == 0) { pf_threadfence(); pf_atomic_add_i(&vctl[4], -1); }
	}
	if (lane == 0) {
		if (P->hot) P->epochs[2 * slot] = w.epoch;
		if (w.hot_alt) P->epochs[2 * slot + 1] = w.epoch_alt;
		pf_atomic_add_ull(&P->stats->pops, (unsigned long long)w.pops);
		pf_atomic_add_ull(&P->stats->pushes, (unsigned long long)w.pushes);
		pf_atomic_add_ull(&P->stats->visits, (unsigned long long)w.visits);
		pf_atomic_add_ull(&P->stats->refills, (unsigned long long)w.refills);
		pf_atomic_add_ull(&P->stats->stale, (unsigned long long)w.stale);
#ifndef PF_DIAG
		if (w.races) pf_atomic_add_ull(&P->stats->races, (unsigned long long)w.races);
#endif
		if (BK) pf_atomic_max_ull(&P->stats->max_net_pops, (unsigned long long)w.max_net_pops);
		pf_atomic_add_ull(&P->stats->nets, nets);
	}
}

/* ------------------------------------------------------------------ per-element bodies of the streaming kernels
 * (shared by pf_kernels.cu and the test emulator) */

/* pathfinder_update_cost (route_common.c:581-610) for one node; returns 1 if the node is overused
 * (feasible_routing, route_common.c:509-531).  pres_cost is not stored: it is a function of occ. */
PF_DEV int pf_update_cost_one(PfNode *nodes, int i, float acc_fac, unsigned char *last_over, int iter_tag) {
	PfNode *n = &nodes[i];
	int occ = n->occ;
	int cap = n->capacity;
	if (occ > cap) {
		n->acc_cost += (occ - cap) * acc_fac;
		if (last_over) last_over[i] = (unsigned char)iter_tag;   /* remembered for the re-route selection */
		return 1;
	}
	return 0;
}

/* wirelength this rr node contributes to the routing in place: occupancy x length for CHANX / CHANY (stats.c:355-409
 * counts a wire once per net using it) */
PF_DEV unsigned pf_node_wirelength_in_use(const PfNode *n) {
	const int ty = n->type_ci & 7;
	if ((ty == 4 || ty == 5) && n->occ > 0) return (unsigned)n->occ * (unsigned)(1 + n->xhigh - n->xlow + n->yhigh - n->ylow);
	return 0u;
}

PF_DEV unsigned pf_tree_wirelength_one(const PfTreeNode *t) {
	int ty = t->type_ci & 7;
	if (ty == 4 || ty == 5) return (unsigned)(1 + t->xhigh - t->xlow + t->yhigh - t->ylow);
	return 0u;
}

/* reserve_locally_used_opins (route_common.c:1435-1491) for one (block, class) group, executed
 * by one thread: rip up last iteration's picks, then take the `count` cheapest OPINs of the class
 * SOURCE in the order the reference's binary heap (route_common.c:1142-1216) would deliver them. */
#define PF_OPIN_HEAP_MAX 128
PF_DEV void pf_reserve_opins_group(PfNode *nodes, const uint32_t *edges, int node_bits, const PfIndexedDev *indexed,
		int source, int count, int *chosen, int rip_up, float pres_fac) {
	/* several GPUs: every rank makes the same reservation on the same synced occupancy, so it is not logged */
	if (rip_up) for (int k = 0; k < count; k++) pf_atomic_add_i(&nodes[chosen[k]].occ, -1);
	if (count == 0) return;
	float hc[PF_OPIN_HEAP_MAX + 2]; int hn[PF_OPIN_HEAP_MAX + 2];
	int tail = 1;
	int e0 = nodes[source].edge_start, ne = nodes[source].num_edges;
	if (ne > PF_OPIN_HEAP_MAX) ne = PF_OPIN_HEAP_MAX;
	for (int k = 0; k < ne; k++) {
		int to = (int)(edges[e0 + k] & pf_node_mask(node_bits));
		const PfNode *n = &nodes[to];
		float pres;
		if (n->occ < n->capacity) pres = 1.; else pres = 1. + (n->occ + 1 - n->capacity) * pres_fac;
		float cost = indexed[n->type_ci >> 3].base_cost * n->acc_cost * pres;
		/* add_to_heap */
		hc[tail] = cost; hn[tail] = to;
		int ifrom = tail, ito = ifrom / 2;
		tail++;
		while (ito >= 1 && hc[ifrom] < hc[ito]) {
			float tc = hc[ito]; hc[ito] = hc[ifrom]; hc[ifrom] = tc;
			int tn = hn[ito]; hn[ito] = hn[ifrom]; hn[ifrom] = tn;
			ifrom = ito; ito = ifrom / 2;
		}
	}
	for (int k = 0; k < count; k++) {
		if (tail == 1) { chosen[k] = chosen[k > 0 ? k - 1 : 0]; continue; }
		int pick = hn[1];
		/* get_heap_head */
		tail--;
		hc[1] = hc[tail]; hn[1] = hn[tail];
		int ifrom = 1, ito = 2;
		while (ito < tail) {
			if (hc[ito + 1] < hc[ito]) ito++;
			if (hc[ito] > hc[ifrom]) break;
			float tc = hc[ito]; hc[ito] = hc[ifrom]; hc[ifrom] = tc;
			int tn = hn[ito]; hn[ito] = hn[ifrom]; hn[ifrom] = tn;
			ifrom = ito; ito = 2 * ifrom;
		}
		pf_atomic_add_i(&nodes[pick].occ, 1);
		chosen[k] = pick;
	}
}

/* Route tree → s_trace order (update_traceback, route_common.c:638-706), one net per call.
 * pass 1 (out == NULL): returns the trace length; pass 2: writes node / switch at out[...] and returns the
 * net's wirelength (get_num_bends_and_length, base/stats.c:355-409); out_term (optional) receives each element's
 * contribution to the routing's serial number. */
PF_DEV unsigned pf_serial_term(const PfTreeNode &n, const short *ptc, int nx, unsigned net_mult) {
	/* get_serial_num, route_common.c:224-254: per trace element the reference adds (inet+1)*(xlow*(nx+1) - yhigh) and
	 * subtracts ptc_num*(inet+1)*10 and type*(inet+1)*100, all in wrapping 32-bit arithmetic */
	return net_mult * (unsigned)(n.xlow * (nx + 1) - n.yhigh - 10 * (int)ptc[n.node] - 100 * (int)(n.type_ci & 7));
}

PF_DEV int pf_trace_of_net(const PfTreeNode *t, int cnt, int *out_node, short *out_sw, unsigned *out_term, const short *ptc, int nx,
		unsigned net_mult) {
	if (cnt <= 1) return 0;
	if (!out_node) {
		int sinks = 0;
		for (int k = 0; k < cnt; k++) if ((t[k].type_ci & 7) == 1) sinks++;
		return cnt + (sinks > 0 ? sinks - 1 : 0);
	}
	int w = 0, wl = 0, k = 0;
	while (k < cnt) {
		int e = k;
		while (e < cnt - 1 && (t[e].type_ci & 7) != 1) e++;
		if (k > 0) {
			const PfTreeNode &j = t[t[k].parent];
			out_node[w] = j.node; out_sw[w] = (short)t[k].sw;
			if (out_term) out_term[w] = pf_serial_term(j, ptc, nx, net_mult);
			w++;
		}
		for (int q = k; q <= e; q++) {
			out_node[w] = t[q].node; out_sw[w] = q < e ? (short)t[q + 1].sw : (short)-1;
			if (out_term) out_term[w] = pf_serial_term(t[q], ptc, nx, net_mult);
			w++;
			int ty = t[q].type_ci & 7;
			if (ty == 4 || ty == 5) wl += 1 + t[q].xhigh - t[q].xlow + t[q].yhigh - t[q].ylow;
		}
		k = e + 1;
	}
	return wl;
}

/* Does this net touch an overused rr node?  (the test the reference's parallel router uses to pick
 * the nets of its "phase two", parallel_route/partitioning_multi_sink_delta_stepping_route.cxx:6241-6269) */
PF_DEV int pf_net_is_congested(const PfNode *nodes, const PfTreeNode *pool, PfNetLoc loc, const unsigned char *last_over,
		int iter_tag, int window, const int *committer, int net) {
	for (int i = 0; i < loc.count; i++) {
		int v = pool[loc.off + i].node;
		const PfNode *n = &nodes[v];
		/* On an overused node every user is re-routed EXCEPT the one that committed it last: the newcomer was
		 * squeezed onto it for lack of anything better, the incumbents usually have alternatives.  If both
		 * were ripped up, the newcomer (routed first, still seeing the incumbent) would move on to its next
		 * victim while the incumbent returns to the freed node, and the conflict would wander for ever. */
		if (n->occ > n->capacity && (!committer || committer[v] != net)) return 1;
		/* PathFinder's history cost is meant to be felt by every user of a contested resource, also by the
		 * net that currently holds it legally: nets on nodes that were overused within the last `window`
		 * iterations are re-routed too, so they can give way (tags are iteration numbers mod 255, 0 = never) */
		if (last_over) {
			int t = last_over[v];
			if (t != 0 && ((iter_tag - t + 255) % 255) <= window) return 1;
		}
	}
	return 0;
}

/* The same test right behind the cost update, from the 4-byte node list and the byte map the update pass just wrote
 * (over[v] == tag  <=>  v is overused now): 5 bytes per tree entry instead of 64. */
PF_DEV int pf_net_is_congested_fast(const int *pool_node, PfNetLoc loc, const unsigned char *over, int tag) {
	for (int i = 0; i < loc.count; i++) if (over[pool_node[loc.off + i]] == (unsigned char)tag) return 1;
	return 0;
}

/* ------------------------------------------------------------------ check_route (route/check_route.c:27-155)
 * One net of a finished routing, judged from its traceback alone: starts at the net's SOURCE; every segment ends
 * at a SINK (iswitch OPEN) that is a still-unmatched pin of the net; consecutive elements are joined by a real rr
 * edge carrying the recorded switch (check_adjacent, check_route.c:262 reduced to edge existence: the rr graph is
 * the authority); every later segment starts at a node already in the net; all pins are reached.  Adds the net's
 * occupancy to occ2 (recompute_occupancy_from_scratch, check_route.c:535-597: the join element of a segment is not
 * counted again) and returns 0 or the first violation. */
#define PF_CHK_NO_TRACE 1
#define PF_CHK_NOT_AT_SOURCE 2
#define PF_CHK_OPEN_SEGMENT 3
#define PF_CHK_JOIN_NOT_IN_NET 4
#define PF_CHK_NO_SUCH_EDGE 5
#define PF_CHK_SINK_HAS_SWITCH 6
#define PF_CHK_WRONG_SINKS 7
#define PF_CHK_BAD_NODE 8
PF_DEV int pf_check_net(const PfNode *nodes, const uint32_t *edges, int node_bits, int num_nodes, const int *term, int ns,
		const int *tn, const short *ts, int len, unsigned char *matched /*[ns+1], zeroed*/, int *occ2, unsigned *wl_out) {
	if (ns == 0) return 0;
	if (len == 0) return PF_CHK_NO_TRACE;
	for (int k = 0; k < len; k++) if (tn[k] < 0 || tn[k] >= num_nodes) return PF_CHK_BAD_NODE;
	if (tn[0] != term[0]) return PF_CHK_NOT_AT_SOURCE;
	int seg_start = 0, reached = 0;
	unsigned wl = 0;
	for (int k = 0; k < len; k++) {
		const int v = tn[k];
		const PfNode n = nodes[v];
		const int ty = n.type_ci & 7;
		const bool first_of_segment = (k == seg_start);
		if (first_of_segment && k > 0) {                    /* the join node: must already belong to the net */
			int found = 0;
			for (int q = 0; q < k && !found; q++) found = (tn[q] == v);
			if (!found) return PF_CHK_JOIN_NOT_IN_NET;
		} else {
			pf_atomic_add_i(&occ2[v], 1);
			if (ty == 4 || ty == 5) wl += (unsigned)(1 + n.xhigh - n.xlow + n.yhigh - n.ylow);
		}
		if (ty == 1 && !(first_of_segment && k > 0)) {      /* a SINK closes the segment */
			if (ts[k] != -1) return PF_CHK_SINK_HAS_SWITCH;
			int hit = 0;
			for (int q = 1; q <= ns && !hit; q++) if (!matched[q] && term[q] == v) { matched[q] = 1; hit = 1; }
			if (!hit) return PF_CHK_WRONG_SINKS;
			reached++;
			seg_start = k + 1;
			continue;
		}
		if (k + 1 >= len) return PF_CHK_OPEN_SEGMENT;       /* the traceback must end in a SINK */
		{	/* an rr edge v -> tn[k+1] with the recorded switch */
			const int to = tn[k + 1];
			int ok = 0;
			for (int e = 0; e < (int)n.num_edges && !ok; e++) {
				const uint32_t ew = edges[n.edge_start + e];
				ok = ((int)(ew & pf_node_mask(node_bits)) == to) && ((int)(ew >> node_bits) == (int)ts[k]);
			}
			if (!ok) return PF_CHK_NO_SUCH_EDGE;
		}
	}
	if (reached != ns) return PF_CHK_WRONG_SINKS;
	*wl_out = wl;
	return 0;
}

#endif /* PF_DEVICE_CUH */
