/*
 * pf_gen_device.cuh — the synthetic k6_N10-style rr graph of pf_gen.cpp as CLOSED-FORM functions of the node id, so that
 * the graph can be built directly in HBM ("native rr-graph generator → CSR directly on device"): no
 * 104-byte AoS detour, no host arrays, no 1.4 GB over PCIe.  Role in the reference: build_rr_graph (route/rr_graph.c:385,
 * rr_graph2.c:741-1366) + alloc_and_load_rr_indexed_data.
 *
 * pf_gen.cpp builds the same graph on the host with lookup tables (first node of a tile, wire covering position p on track
 * t) and an edge callback run twice; here every lookup is arithmetic:
 *   node order   tiles (x outer, y inner): class nodes then pin nodes; then every CHANX channel y = 0..ny, inside a channel
 *                the wires by start position then track; then every CHANY channel x = 0..nx
 *   wire starts  track t has stagger s = (t/2) % L: its segments start at p = 1 (clipped) and at every p with
 *                (p - 1 - s) % L == 0, so the wires starting before position p in a channel are counted by pf_gen_pref
 *   out-edges    per node in exactly the order pf_gen.cpp emits them (per-row order is part of the graph: the traceback
 *                stores switch ids, and the parity tests compare a generated router with an uploaded one bit for bit)
 * The functions are PF_DEV like the rest of the device code: the CUDA kernels call them per node, the CPU warp emulator
 * calls them from plain loops (tests/test_generator.py checks the device-built graph against pf_gen.cpp's arrays).
 */
#ifndef PF_GEN_DEVICE_CUH
#define PF_GEN_DEVICE_CUH

#include <stddef.h>
#include "pf_layout.h"

#define PF_GEN_CLB_PINS 51
#define PF_GEN_CLB_CLASSES 12
#define PF_GEN_CLB_IN 40
#define PF_GEN_CLB_OUT 10
#define PF_GEN_CLB_NODES (PF_GEN_CLB_PINS + PF_GEN_CLB_CLASSES)
#define PF_GEN_MAX_W 512

/* x / d and x % d for 0 <= x < 2^31 by a divisor fixed per graph: one 32 x 32 -> 64-bit multiply and a shift instead of the ~25
 * instructions of a 32-bit division.  The generator is integer arithmetic on (W, L, wires per channel, ...) throughout — per
 * out-edge a dozen remainders — and was bound by exactly that (cfg 4: degree pass 3.1 ms, fill pass 10 ms for 1.4 GB of output).
 * m = ceil(2^(31+s) / d), s = ceil(log2 d): exact for every x below 2^31 (the error m d - 2^(31+s) is below d <= 2^s). */
struct PfFastDiv { unsigned m; int sh; int d; };
PF_DEV int pf_fdiv(int x, const PfFastDiv &D) { return (int)(((unsigned long long)D.m * (unsigned long long)(unsigned)x) >> D.sh); }
PF_DEV int pf_fmod(int x, const PfFastDiv &D) { return x - pf_fdiv(x, D) * D.d; }
#ifndef __CUDA_ARCH__
static inline PfFastDiv pf_fastdiv_make(int d) {
	PfFastDiv D;
	int s = 0;
	while ((1ll << s) < (long long)d) s++;
	D.d = d; D.sh = 31 + s;
	D.m = (unsigned)((((unsigned long long)1 << (31 + s)) + (unsigned long long)d - 1ull) / (unsigned long long)d);
	return D;
}
#endif

struct PfGenDev {
	int nx, ny, W, L, fc_in, fc_out, io_cap;
	PfFastDiv dW, dL, dHalf, dWpcX, dWpcY, dColInner, dIoNodes, dClbNodes;   /* W, L, W / 2, wpc_x, wpc_y, col_inner, io_nodes, PF_GEN_CLB_NODES */
	int m7, m28;                  /* 7 % W, 28 % W (connection-box stepping) */
	int io_nodes;                 /* nodes of an IO tile: 3 classes + 3 pins per pad */
	int col0, col_inner;          /* nodes of the x = 0 column, of an inner column */
	int wpc_x, wpc_y;             /* wires per CHANX / CHANY channel */
	int chanx0, chany0, num_nodes;
	int node_bits;                /* edge word: target | switch << node_bits (pf_layout.h) */
	float R_metal, C_per_tile;
	const short *cb_inv;          /* [W]: k with (k * W) / fc_in == d, or -1 — the connection-box pattern inverted */
};

/* wire types in the device node record */
#define PF_GEN_SW_WIRE 0
#define PF_GEN_SW_IPIN 1
#define PF_GEN_SW_DELAYLESS 2

PF_DEV int pf_gen_is_clb(const PfGenDev &G, int x, int y) { return x >= 1 && x <= G.nx && y >= 1 && y <= G.ny; }
PF_DEV int pf_gen_is_io(const PfGenDev &G, int x, int y) {
	const int ex = (x == 0 || x == G.nx + 1), ey = (y == 0 || y == G.ny + 1);
	return ex != ey;
}
/* first (class) node of tile (x, y); -1 for a corner */
PF_DEV int pf_gen_tile_base(const PfGenDev &G, int x, int y) {
	if (!pf_gen_is_clb(G, x, y) && !pf_gen_is_io(G, x, y)) return -1;
	if (x == 0) return (y - 1) * G.io_nodes;
	if (x == G.nx + 1) return G.col0 + G.nx * G.col_inner + (y - 1) * G.io_nodes;
	const int c = G.col0 + (x - 1) * G.col_inner;
	if (y == 0) return c;
	if (y == G.ny + 1) return c + G.io_nodes + G.ny * PF_GEN_CLB_NODES;
	return c + G.io_nodes + (y - 1) * PF_GEN_CLB_NODES;
}
/* tracks of stagger s among the W/2 direction pairs */
PF_DEV int pf_gen_groups_with(const PfGenDev &G, int s) { const int g = G.W / 2; return s < g ? pf_fdiv(g - 1 - s, G.dL) + 1 : 0; }
/* wires of one channel that start at a position < p (p >= 1) */
PF_DEV int pf_gen_pref(const PfGenDev &G, int p) {
	if (p <= 1) return 0;
	const int q = pf_fdiv(p - 2, G.dL), rem = p - 2 - q * G.L;
	int n = G.W + q * G.W;
	for (int j = 1; j <= rem; j++) n += 2 * pf_gen_groups_with(G, j);      /* j <= rem < L */
	return n;
}
/* (p - 1 - s) mod L with p >= 1, 0 <= s < L: p - 1 - s + L is positive */
PF_DEV int pf_gen_seg_start(const PfGenDev &G, int p, int s) { const int a = p - pf_fmod(p - 1 - s + G.L, G.dL); return a < 1 ? 1 : a; }
PF_DEV int pf_gen_seg_end(const PfGenDev &G, int p, int s, int P) { const int a = p - pf_fmod(p - 1 - s + G.L, G.dL); const int b = a + G.L - 1; return b > P ? P : b; }
/* the wire covering position p on track t of a channel: node id, and its span [a, b] */
PF_DEV int pf_gen_wire_at(const PfGenDev &G, int horiz, int chan, int p, int t, int *a_out, int *b_out) {
	const int grp = pf_fdiv(t >> 1, G.dL), s = (t >> 1) - grp * G.L, P = horiz ? G.nx : G.ny;
	const int a = pf_gen_seg_start(G, p, s);
	const int rank = (a == 1) ? t : 2 * grp + (t & 1);
	if (a_out) *a_out = a;
	if (b_out) *b_out = pf_gen_seg_end(G, p, s, P);
	return (horiz ? G.chanx0 + chan * G.wpc_x : G.chany0 + chan * G.wpc_y) + pf_gen_pref(G, a) + rank;
}

/* does the INC (even) / DEC (odd) wire of track t that covers position pos START there?  (an INC wire is entered at its low end
 * a, a DEC wire at its high end b; segments are clipped at 1 and P) */
PF_DEV int pf_gen_starts_here(const PfGenDev &G, int pos, int t, int P) {
	const int s = pf_fmod(t >> 1, G.dL);
	if (t & 1) return pos == P || pf_fmod(pos - s + G.L, G.dL) == 0;      /* pos >= 1, s < L: the arguments are positive */
	return pos == 1 || pf_fmod(pos - 1 - s + G.L, G.dL) == 0;
}

struct PfGenNode { int type, x0, y0, x1, y1, ptc, ci, cap; float R, C; int horiz, chan, t; /* wires */ int tx, ty, local, clb; /* tile nodes */ };

/* everything about node v but its edges */
PF_DEV PfGenNode pf_gen_decode(const PfGenDev &G, int v) {
	PfGenNode n;
	n.horiz = n.chan = n.t = 0; n.tx = n.ty = n.local = n.clb = 0; n.R = 0.f; n.C = 0.f;
	if (v >= G.chanx0) {
		const int horiz = v < G.chany0;
		const int idx0 = horiz ? v - G.chanx0 : v - G.chany0, wpc = horiz ? G.wpc_x : G.wpc_y, P = horiz ? G.nx : G.ny;
		const int chan = pf_fdiv(idx0, horiz ? G.dWpcX : G.dWpcY);
		int idx = idx0 - chan * wpc, a, t;
		if (idx < G.W) { a = 1; t = idx; }
		else {
			idx -= G.W;
			const int blk = pf_fdiv(idx, G.dW);
			int q = 2 + blk * G.L;
			idx -= blk * G.W;
			for (;; q++) { const int c = 2 * pf_gen_groups_with(G, pf_fmod(q - 1, G.dL)); if (idx < c) break; idx -= c; }
			a = q;
			t = 2 * ((idx / 2) * G.L + pf_fmod(a - 1, G.dL)) + (idx & 1);
		}
		const int b = pf_gen_seg_end(G, a, pf_fmod(t >> 1, G.dL), P), len = b - a + 1;
		n.type = horiz ? 4 : 5; n.ci = horiz ? 4 : 5; n.cap = 1; n.ptc = t;
		n.x0 = horiz ? a : chan; n.x1 = horiz ? b : chan; n.y0 = horiz ? chan : a; n.y1 = horiz ? chan : b;
		n.R = G.R_metal * len; n.C = G.C_per_tile * len;
		n.horiz = horiz; n.chan = chan; n.t = t;
		return n;
	}
	int x, y, local;
	if (v < G.col0) { const int q = pf_fdiv(v, G.dIoNodes); x = 0; y = 1 + q; local = v - q * G.io_nodes; }
	else {
		const int u = v - G.col0;
		if (u < G.nx * G.col_inner) {
			const int c = pf_fdiv(u, G.dColInner);
			x = 1 + c;
			const int w = u - c * G.col_inner;
			if (w < G.io_nodes) { y = 0; local = w; }
			else if (w < G.io_nodes + G.ny * PF_GEN_CLB_NODES) { const int q = pf_fdiv(w - G.io_nodes, G.dClbNodes); y = 1 + q; local = w - G.io_nodes - q * PF_GEN_CLB_NODES; }
			else { y = G.ny + 1; local = w - G.io_nodes - G.ny * PF_GEN_CLB_NODES; }
		} else { x = G.nx + 1; const int w = u - G.nx * G.col_inner; const int q = pf_fdiv(w, G.dIoNodes); y = 1 + q; local = w - q * G.io_nodes; }
	}
	n.tx = x; n.ty = y; n.local = local; n.clb = pf_gen_is_clb(G, x, y);
	n.x0 = n.x1 = x; n.y0 = n.y1 = y;
	if (n.clb) {
		if (local == 0) { n.type = 1; n.ptc = 0; n.ci = 1; n.cap = PF_GEN_CLB_IN; }
		else if (local <= PF_GEN_CLB_OUT) { n.type = 0; n.ptc = local; n.ci = 0; n.cap = 1; }
		else if (local == 11) { n.type = 1; n.ptc = 11; n.ci = 1; n.cap = 1; }
		else {
			const int p = local - PF_GEN_CLB_CLASSES;            /* pin number = ptc: inputs 0..39, outputs 40..49, clock 50 */
			const int is_out = p >= PF_GEN_CLB_IN && p < PF_GEN_CLB_IN + PF_GEN_CLB_OUT;
			n.type = is_out ? 3 : 2; n.ptc = p; n.ci = is_out ? 2 : 3; n.cap = 1;
		}
	} else {
		const int nc = 3 * G.io_cap;
		const int p = local < nc ? local : local - nc;
		if (local < nc) { n.type = (p % 3 == 1) ? 0 : 1; n.ci = (p % 3 == 1) ? 0 : 1; }
		else { n.type = (p % 3 == 1) ? 3 : 2; n.ci = (p % 3 == 1) ? 2 : 3; }
		n.ptc = p; n.cap = 1;
	}
	return n;
}

/* the two wire -> wire turns at switch box q of wire w (pf_gen.cpp: turns) */
PF_DEV int pf_gen_turns(const PfGenDev &G, int horiz, int chan, int q, int t, uint32_t *out, int n) {
	const int qx = horiz ? q : chan, qy = horiz ? chan : q;
	int P2;
	if (horiz) { if (qx < 0 || qx > G.nx) return n; P2 = G.ny; }
	else { if (qy < 0 || qy > G.ny) return n; P2 = G.nx; }
	const int pchan = horiz ? qx : qy;            /* the perpendicular channel */
	const int base = horiz ? qy : qx, g = t / 2, half = G.W / 2;
	if (base + 1 <= P2) {                         /* an INC wire starting at base + 1 */
		const int pos = base + 1;
		for (int k = 0; k < half; k++) {
			const int t2 = 2 * pf_fmod(g + qx + qy + k, G.dHalf);
			if (!pf_gen_starts_here(G, pos, t2, P2)) continue;
			if (out) out[n] = (uint32_t)pf_gen_wire_at(G, !horiz, pchan, pos, t2, NULL, NULL) | ((uint32_t)PF_GEN_SW_WIRE << G.node_bits);
			n++;
			break;
		}
	}
	if (base >= 1) {                              /* a DEC wire starting at base (its high end) */
		const int pos = base;
		for (int k = 0; k < half; k++) {
			const int t2 = 2 * pf_fmod(g + 2 * qx + qy + k, G.dHalf) + 1;
			if (!pf_gen_starts_here(G, pos, t2, P2)) continue;
			if (out) out[n] = (uint32_t)pf_gen_wire_at(G, !horiz, pchan, pos, t2, NULL, NULL) | ((uint32_t)PF_GEN_SW_WIRE << G.node_bits);
			n++;
			break;
		}
	}
	return n;
}

/* connection-box edges wire -> IPIN of the pins of tile (x, y) that face the wire's channel on `side` at position pos */
PF_DEV int pf_gen_cb_tile(const PfGenDev &G, int x, int y, int side, int pos, int t, uint32_t *out, int n) {
	const int clb = pf_gen_is_clb(G, x, y);
	if (!clb && !pf_gen_is_io(G, x, y)) return n;
	const int base = pf_gen_tile_base(G, x, y);
	const int npins = clb ? PF_GEN_CLB_PINS : 3 * G.io_cap, pin0 = base + (clb ? PF_GEN_CLB_CLASSES : 3 * G.io_cap);
	if (!clb) {
		const int io_side = (x == 0) ? 1 : (x == G.nx + 1) ? 3 : (y == 0) ? 0 : 2;
		if (io_side != side) return n;
	}
	/* the pattern t == (p * 7 + pos + (k * W) / fc_in) % W inverted: d = (t - pos - 7 p) mod W must be one of the fc_in offsets.
	 * Only the pins that face this side are visited, and d is stepped instead of recomputed. */
	int d = t - pf_fmod(pos, G.dW);               /* (t - pos) mod W, 0 <= t < W */
	if (d < 0) d += G.W;
	const int m7 = G.m7;
	if (clb) {
		for (int q = 0; q < side; q++) { d -= m7; if (d < 0) d += G.W; }
		const int m28 = G.m28;
		for (int p = side; p < PF_GEN_CLB_PINS; p += 4) {
			if (!(p >= PF_GEN_CLB_IN && p < PF_GEN_CLB_IN + PF_GEN_CLB_OUT) && G.cb_inv[d] >= 0) {
				if (out) out[n] = (uint32_t)(pin0 + p) | ((uint32_t)PF_GEN_SW_IPIN << G.node_bits);
				n++;
			}
			d -= m28; if (d < 0) d += G.W;
		}
	} else {
		for (int p = 0; p < npins; p++) {
			if (p % 3 != 1 && G.cb_inv[d] >= 0) { if (out) out[n] = (uint32_t)(pin0 + p) | ((uint32_t)PF_GEN_SW_IPIN << G.node_bits); n++; }
			d -= m7; if (d < 0) d += G.W;
		}
	}
	return n;
}

/* out-edges of node v in pf_gen.cpp's order; out == NULL: only the degree */
PF_DEV int pf_gen_node_edges(const PfGenDev &G, int v, const PfGenNode &nd, uint32_t *out) {
	int n = 0;
	if (v < G.chanx0) {
		const int base = pf_gen_tile_base(G, nd.tx, nd.ty);
		const int ncls = nd.clb ? PF_GEN_CLB_CLASSES : 3 * G.io_cap, pin0 = base + ncls;
		if (nd.type == 0) {                                   /* SOURCE -> its OPIN */
			const int pin = nd.clb ? PF_GEN_CLB_IN + (nd.local - 1) : nd.local;
			if (out) out[0] = (uint32_t)(pin0 + pin) | ((uint32_t)PF_GEN_SW_DELAYLESS << G.node_bits);
			return 1;
		}
		if (nd.type == 1) return 0;                           /* SINK */
		const int p = nd.ptc;
		if (nd.type == 2) {                                   /* IPIN -> its SINK */
			const int cls = nd.clb ? (p < PF_GEN_CLB_IN ? 0 : 11) : p;
			if (out) out[0] = (uint32_t)(base + cls) | ((uint32_t)PF_GEN_SW_DELAYLESS << G.node_bits);
			return 1;
		}
		/* OPIN -> the wires that start next to the tile */
		const int x = nd.tx, y = nd.ty;
		const int side = nd.clb ? (p & 3) : ((x == 0) ? 1 : (x == G.nx + 1) ? 3 : (y == 0) ? 0 : 2);
		const int horiz = (side == 0 || side == 2);
		const int chan = horiz ? (side == 0 ? y : y - 1) : (side == 1 ? x : x - 1);
		const int pos = horiz ? x : y, P = horiz ? G.nx : G.ny;
		if (pos < 1 || pos > P) return 0;
		int t = pf_fmod(p * 11 + pos * 3, G.dW);
		for (int k = 0; k < G.W && n < G.fc_out; k++) {
			if (pf_gen_starts_here(G, pos, t, P)) {
				if (out) out[n] = (uint32_t)pf_gen_wire_at(G, horiz, chan, pos, t, NULL, NULL) | ((uint32_t)PF_GEN_SW_WIRE << G.node_bits);
				n++;
			}
			if (++t == G.W) t = 0;
		}
		return n;
	}
	/* a wire: connection boxes first (pf_gen.cpp emits them in its tile loop: x outer, y inner), then the straight
	 * continuation and the turns at its switch boxes */
	const int horiz = nd.horiz, chan = nd.chan, t = nd.t;
	const int a = horiz ? nd.x0 : nd.y0, b = horiz ? nd.x1 : nd.y1;
	if (horiz) {
		for (int x = a; x <= b; x++) {
			n = pf_gen_cb_tile(G, x, chan, 0, x, t, out, n);          /* tile below the channel: its top side */
			n = pf_gen_cb_tile(G, x, chan + 1, 2, x, t, out, n);      /* tile above: its bottom side */
		}
	} else {
		/* tiles (chan, y) face the channel with their right side, tiles (chan + 1, y) with their left side; x outer */
		for (int y = a; y <= b; y++) n = pf_gen_cb_tile(G, chan, y, 1, y, t, out, n);
		for (int y = a; y <= b; y++) n = pf_gen_cb_tile(G, chan + 1, y, 3, y, t, out, n);
	}
	const int P = horiz ? G.nx : G.ny;
	if (!(t & 1)) {                                           /* INC: enters at a, switch boxes q = a .. b */
		if (b < P) { if (out) out[n] = (uint32_t)pf_gen_wire_at(G, horiz, chan, b + 1, t, NULL, NULL) | ((uint32_t)PF_GEN_SW_WIRE << G.node_bits); n++; }
		for (int q = a; q <= b; q++) n = pf_gen_turns(G, horiz, chan, q, t, out, n);
	} else {                                                  /* DEC: enters at b, switch boxes q = a-1 .. b-1 */
		if (a > 1) { if (out) out[n] = (uint32_t)pf_gen_wire_at(G, horiz, chan, a - 1, t, NULL, NULL) | ((uint32_t)PF_GEN_SW_WIRE << G.node_bits); n++; }
		for (int q = a - 1; q <= b - 1; q++) n = pf_gen_turns(G, horiz, chan, q, t, out, n);
	}
	return n;
}

/* node record + ptc of node v, given the start of its edge row */
PF_DEV void pf_gen_write_node(const PfGenNode &nd, int edge_start, int degree, PfNode *rec, short *ptc) {
	PfNode d;
	d.xlow = (short)nd.x0; d.ylow = (short)nd.y0; d.xhigh = (short)nd.x1; d.yhigh = (short)nd.y1;
	d.R = nd.R; d.C = nd.C; d.occ = 0; d.acc_cost = 1.f;
	d.edge_start = edge_start; d.num_edges = (unsigned short)degree;
	d.type_ci = (unsigned char)(nd.type | (nd.ci << 3)); d.capacity = (unsigned char)nd.cap;
	*rec = d;
	*ptc = (short)nd.ptc;
}

#endif /* PF_GEN_DEVICE_CUH */
