/*
 * pf_emu.h — TEST INFRASTRUCTURE: a lock-step warp emulator for the device code in
 * parallel_eda_b200/csrc/pf_device.cuh.  There is no GPU in the development container, so the
 * warp-synchronous router is written against the small primitive set below; the CUDA build maps
 * them to the hardware intrinsics (pf_device.cuh), this header maps them to 32 cooperative
 * fibers per warp (ucontext) that switch at every warp collective.  It is never compiled into
 * the product library (libpf_router.so has no CPU path and fails loudly without CUDA).
 *
 * Execution model: all fibers of all emulated warps are resumed round-robin; a fiber runs until
 * its next collective (shuffle / ballot / match / syncwarp), deposits its operand in a
 * double-buffered per-warp exchange slot and yields.  When it is resumed every other lane of its
 * warp has deposited the operand of the same collective, so the result can be formed.  Every
 * deposit carries the lane's collective sequence number and an op tag; a mismatch (= divergent
 * control flow around a collective, the classic warp-synchronous bug) aborts with a message.
 * PF_EMU_REVERSE=1 resumes lanes in descending order, which exposes missing pf_syncwarp()
 * between a shared/global store and a dependent load by another lane.
 */
#ifndef PF_EMU_H
#define PF_EMU_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define PF_DEV static inline
#define PF_WARP 32

struct pf_u4 { unsigned x, y, z, w; };

struct pf_emu_warp {
	uint64_t slot[2][PF_WARP];
	uint32_t seq[2][PF_WARP];
	uint32_t tag[2][PF_WARP];
	int done[PF_WARP];
	int warp_id;
};

struct pf_emu_lane {
	pf_emu_warp *warp;
	int lane;
	uint32_t seq;
};

extern pf_emu_lane *pf_emu_cur;          /* the running fiber */
void pf_emu_yield(void);                 /* switch to the scheduler */

/* launches `nwarps` emulated warps; fn(arg, warp_id) is executed by each of the 32 lanes */
typedef void (*pf_emu_warp_fn)(void *arg, int warp_id);
void pf_emu_launch(pf_emu_warp_fn fn, void *arg, int nwarps);

PF_DEV int pf_lane(void) { return pf_emu_cur->lane; }

static inline void pf_emu_exchange(uint64_t v, uint32_t tag, uint64_t out[PF_WARP]) {
	pf_emu_lane *me = pf_emu_cur;
	pf_emu_warp *w = me->warp;
	int b = me->seq & 1;
	w->slot[b][me->lane] = v;
	w->seq[b][me->lane] = me->seq;
	w->tag[b][me->lane] = tag;
	pf_emu_yield();
	for (int i = 0; i < PF_WARP; i++) {
		if (w->seq[b][i] != me->seq || w->tag[b][i] != tag) {
			fprintf(stderr, "pf_emu: warp %d divergent collective: lane %d at seq %u tag %u, lane %d at seq %u tag %u done=%d\n",
					w->warp_id, me->lane, me->seq, tag, i, w->seq[b][i], w->tag[b][i], w->done[i]);
			abort();
		}
		out[i] = w->slot[b][i];
	}
	me->seq++;
}

PF_DEV void pf_syncwarp(void) { uint64_t o[PF_WARP]; pf_emu_exchange(0, 1, o); }
PF_DEV unsigned pf_ballot(int pred) {
	uint64_t o[PF_WARP]; unsigned m = 0;
	pf_emu_exchange(pred ? 1 : 0, 2, o);
	for (int i = 0; i < PF_WARP; i++) if (o[i]) m |= 1u << i;
	return m;
}
PF_DEV int pf_any(int pred) { return pf_ballot(pred) != 0; }
PF_DEV int pf_shfl_i(int v, int src) { uint64_t o[PF_WARP]; pf_emu_exchange((uint64_t)(uint32_t)v, 3, o); return (int)(uint32_t)o[src & 31]; }
PF_DEV float pf_shfl_f(float v, int src) {
	uint64_t o[PF_WARP]; uint32_t u; memcpy(&u, &v, 4);
	pf_emu_exchange(u, 4, o); u = (uint32_t)o[src & 31]; memcpy(&v, &u, 4); return v;
}
PF_DEV uint64_t pf_shfl_u64(uint64_t v, int src) { uint64_t o[PF_WARP]; pf_emu_exchange(v, 5, o); return o[src & 31]; }
PF_DEV unsigned pf_match_any(int key) {
	uint64_t o[PF_WARP]; unsigned m = 0;
	pf_emu_exchange((uint64_t)(uint32_t)key, 6, o);
	for (int i = 0; i < PF_WARP; i++) if ((uint32_t)o[i] == (uint32_t)key) m |= 1u << i;
	return m;
}
PF_DEV uint64_t pf_warp_min_u64(uint64_t v) {
	uint64_t o[PF_WARP]; pf_emu_exchange(v, 7, o);
	uint64_t m = o[0]; for (int i = 1; i < PF_WARP; i++) if (o[i] < m) m = o[i];
	return m;
}
PF_DEV float pf_warp_min_f(float v) {
	uint64_t o[PF_WARP]; uint32_t u; memcpy(&u, &v, 4); pf_emu_exchange(u, 8, o);
	float m = v; for (int i = 0; i < PF_WARP; i++) { float t; u = (uint32_t)o[i]; memcpy(&t, &u, 4); if (t < m) m = t; }
	return m;
}
PF_DEV int pf_warp_sum_i(int v) {
	uint64_t o[PF_WARP]; pf_emu_exchange((uint64_t)(uint32_t)v, 9, o);
	int s = 0; for (int i = 0; i < PF_WARP; i++) s += (int)(uint32_t)o[i];
	return s;
}
PF_DEV int pf_warp_max_i(int v) {
	uint64_t o[PF_WARP]; pf_emu_exchange((uint64_t)(uint32_t)v, 10, o);
	int s = (int)(uint32_t)o[0]; for (int i = 1; i < PF_WARP; i++) if ((int)(uint32_t)o[i] > s) s = (int)(uint32_t)o[i];
	return s;
}
PF_DEV void pf_prefetch_l2(const void *) {}
PF_DEV int pf_popc(unsigned m) { return __builtin_popcount(m); }
PF_DEV int pf_ffs(unsigned m) { return __builtin_ffs((int)m); }   /* 1-based, 0 if none */
PF_DEV unsigned pf_lanemask_lt(void) { return (1u << pf_lane()) - 1u; }

PF_DEV int pf_atomic_add_i(int *p, int v) { int o = *p; *p = o + v; return o; }
PF_DEV unsigned long long pf_atomic_add_ull(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
PF_DEV void pf_atomic_max_ull(unsigned long long *p, unsigned long long v) { if (v > *p) *p = v; }
PF_DEV int pf_atomic_or_i(int *p, int v) { int o = *p; *p = o | v; return o; }
PF_DEV int pf_atomic_min_i(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
PF_DEV int pf_atomic_exch_i(int *p, int v) { int o = *p; *p = v; return o; }
PF_DEV int pf_atomic_cas_i(int *p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
PF_DEV int pf_ld_volatile_i(const int *p) { return *(const volatile int *)p; }
PF_DEV void pf_threadfence(void) {}
PF_DEV void pf_spin_pause(void) {}     /* spin loops in the device code contain a warp collective per turn: that is the yield */
PF_DEV void pf_atomic_max_f(float *p, float v) { if (v > *p) *p = v; }
PF_DEV void pf_atomic_min_f(float *p, float v) { if (v < *p) *p = v; }
PF_DEV pf_u4 pf_ld_cg_u4(const void *p) { pf_u4 v; memcpy(&v, p, 16); return v; }   /* L2-coherent load */
PF_DEV pf_u4 pf_ld_u4(const void *p) { pf_u4 v; memcpy(&v, p, 16); return v; }
struct pf_u8 { unsigned a, b, c, d, e, f, g, h; };
PF_DEV pf_u8 pf_ld_cg_u8(const void *p) { pf_u8 v; memcpy(&v, p, 32); return v; }
PF_DEV void pf_st_u4(void *p, pf_u4 v) { memcpy(p, &v, 16); }
PF_DEV float pf_int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
PF_DEV int pf_float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
PF_DEV double pf_ceil(double x) { return ceil(x); }
PF_DEV float pf_sqrtf(float x) { return sqrtf(x); }
PF_DEV float pf_ceilf(float x) { return ceilf(x); }
PF_DEV float pf_powf(float x, float y) { return powf(x, y); }

#endif
