/*
 * pf_gen.cpp — native generator of a synthetic k6_N10-style routing problem directly in the flat
 * pf_problem form (no 104-byte AoS detour): nx × ny CLB grid with an IO ring, one unidirectional
 * length-L segment type, W tracks per channel, plus random multi-pin nets.
 *
 * Role in the reference: build_rr_graph (vpr/SRC/route/rr_graph.c:385, rr_graph2.c:741-1366) and
 * alloc_and_load_rr_indexed_data (rr_graph_indexed_data.c:41-319) always regenerate the graph from
 * the architecture; VPR 7 has no rr-graph file.  This generator follows the same conventions —
 * node order of alloc_and_load_rr_node_indices (rr_graph2.c:741-829: per tile the class nodes then
 * the pin nodes, then every CHANX, then every CHANY), SOURCE/SINK per pin class with
 * capacity = class size (rr_graph.c:1169-1222), unidirectional wires driven only at their start
 * (rr_graph.c:1388-1417), connection boxes of Fc_in / Fc_out tracks, switch points at every tile
 * along a wire (sb pattern 1 1 1 1 1), DELAY_NORMALIZED base costs (rr_graph_indexed_data.c:122-214)
 * — but it is an independent construction: track permutations are simple modular patterns, not
 * VPR's Wilton tables, so graphs are VPR-like (same node counts and degree profile as the
 * reference's dump of the same grid), not bit-identical.  Both arms of every comparison (CUDA
 * router, CPU oracle, the reference's own router through oracle/_ref inject mode) consume the
 * same generated problem, so parity never depends on that.
 *
 * Workload of BASELINE.json configs[4]: 400×400, W=100, 200,000 nets × 3 sinks,
 * sinks within ±16 tiles of the source, each CLB output drives at most one net, at most 40 sinks
 * per CLB, no net has two sinks in one CLB, MT19937 seed 20260921, timing analysis off.
 */
#include "../../include/pf_gen.h"
#ifndef PF_DEV
#define PF_DEV static inline
#endif
#include "pf_gen_device.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <random>
#include <vector>

namespace {

struct Gen {
	int nx, ny, W, L, fc_in, fc_out, io_cap;
	std::vector<int16_t> xlow, ylow, xhigh, yhigh, ptc, ci, cap;
	std::vector<uint8_t> type, dir;
	std::vector<float> R, C;
	std::vector<std::vector<int32_t>> adj_to;   /* built per node then flattened */
	std::vector<int32_t> row_ptr, edge_to;
	std::vector<int16_t> edge_sw;
	/* lookup */
	std::vector<int32_t> tile_class0, tile_pin0;   /* first class / pin node of tile (x,y) */
	std::vector<int32_t> chanx0, chany0;           /* first wire node of channel */
	std::vector<std::vector<int32_t>> chanx_wire, chany_wire;   /* [channel][pos*W + track] -> node */
	int add_node(int t, int x0, int y0, int x1, int y1, int p, int c, int cp, int d, float r, float cc) {
		xlow.push_back((int16_t)x0); ylow.push_back((int16_t)y0); xhigh.push_back((int16_t)x1); yhigh.push_back((int16_t)y1);
		ptc.push_back((int16_t)p); ci.push_back((int16_t)c); cap.push_back((int16_t)cp); type.push_back((uint8_t)t);
		dir.push_back((uint8_t)d); R.push_back(r); C.push_back(cc);
		return (int)type.size() - 1;
	}
	int tile(int x, int y) const { return x * (ny + 2) + y; }
	bool is_clb(int x, int y) const { return x >= 1 && x <= nx && y >= 1 && y <= ny; }
	bool is_io(int x, int y) const {
		bool ex = (x == 0 || x == nx + 1), ey = (y == 0 || y == ny + 1);
		return (ex != ey);
	}
};

/* CLB: pins 0..39 inputs (class 0, SINK cap 40), 40..49 outputs (classes 1..10), 50 clock (class 11) */
const int CLB_PINS = 51, CLB_CLASSES = 12, CLB_IN = 40, CLB_OUT = 10;
/* IO pad k: pins 3k outpad (in), 3k+1 inpad (out), 3k+2 clock; classes 3k, 3k+1, 3k+2 */

const float R_METAL = 101.f, C_PER_TILE = 2.7866e-14f;   /* metal + switch loading per tile, as the reference dump shows */

struct Edge { int32_t from, to; int16_t sw; };

/* wires of one channel: track t, group g = t/2, direction t%2 (0 INC, 1 DEC), stagger s = g % L.
 * segment boundaries along p in [1..P]: starts at a ≡ 1 + s (mod L), clipped. */
inline int seg_start(int p, int s, int L) { int a = p - ((p - 1 - s) % L + L) % L; return a < 1 ? 1 : a; }
inline int seg_end(int a_unclipped_p, int s, int L, int P) {
	int a = a_unclipped_p - ((a_unclipped_p - 1 - s) % L + L) % L;   /* unclipped start */
	int b = a + L - 1;
	return b > P ? P : b;
}

}  // namespace

extern "C" void pf_gen_params_default(pf_gen_params *g) {
	memset(g, 0, sizeof(*g));
	g->nx = 400; g->ny = 400; g->W = 100; g->L = 4;
	g->num_nets = 200000; g->sinks_per_net = 3; g->window = 16; g->seed = 20260921u;
	g->fc_in = 0.15f; g->fc_out = 0.10f; g->io_capacity = 8; g->bb_factor = 3;
}


/* closed-form layout of the generated graph (shared with the device generator, pf_gen_device.cuh) */
extern "C" int pf_gen_dev_params(const pf_gen_params *gp, PfGenDev *G, short *cb_inv /* [PF_GEN_MAX_W] */) {
	if (!gp || gp->nx < 2 || gp->ny < 2 || gp->W < 8 || (gp->W & 1) || gp->W > PF_GEN_MAX_W || gp->L < 1 || gp->nx > 30000 || gp->ny > 30000) return PF_EINVAL;
	memset(G, 0, sizeof(*G));
	G->nx = gp->nx; G->ny = gp->ny; G->W = gp->W; G->L = gp->L; G->io_cap = gp->io_capacity > 0 ? gp->io_capacity : 8;
	G->fc_in = (int)(gp->fc_in * G->W + 0.5f); if (G->fc_in < 2) G->fc_in = 2;
	G->fc_out = (int)(gp->fc_out * G->W + 0.5f); if (G->fc_out < 2) G->fc_out = 2;
	G->io_nodes = 6 * G->io_cap;
	G->col0 = G->ny * G->io_nodes;
	G->col_inner = 2 * G->io_nodes + G->ny * PF_GEN_CLB_NODES;
	G->dW = pf_fastdiv_make(G->W); G->dL = pf_fastdiv_make(G->L); G->dHalf = pf_fastdiv_make(G->W / 2);
	G->dColInner = pf_fastdiv_make(G->col_inner); G->dIoNodes = pf_fastdiv_make(G->io_nodes); G->dClbNodes = pf_fastdiv_make(PF_GEN_CLB_NODES);
	G->m7 = 7 % G->W; G->m28 = 28 % G->W;
	G->R_metal = R_METAL; G->C_per_tile = C_PER_TILE;
	G->wpc_x = pf_gen_pref(*G, G->nx + 1); G->wpc_y = pf_gen_pref(*G, G->ny + 1);
	G->dWpcX = pf_fastdiv_make(G->wpc_x); G->dWpcY = pf_fastdiv_make(G->wpc_y);
	const long long chanx0 = 2ll * G->col0 + (long long)G->nx * G->col_inner;
	const long long chany0 = chanx0 + (long long)(G->ny + 1) * G->wpc_x;
	const long long N = chany0 + (long long)(G->nx + 1) * G->wpc_y;
	if (N >= (1ll << 30)) return PF_EINVAL;
	G->chanx0 = (int)chanx0; G->chany0 = (int)chany0; G->num_nodes = (int)N;
	int nb = PF_MIN_NODE_BITS; while ((1ll << nb) < N) nb++;
	G->node_bits = nb;
	for (int d = 0; d < G->W; d++) cb_inv[d] = -1;
	for (int k = 0; k < G->fc_in; k++) cb_inv[(k * G->W) / G->fc_in] = (short)k;
	G->cb_inv = cb_inv;
	return PF_OK;
}

/* switch / cost-index tables, nets and router options of the generated problem (p.nx, p.ny are set by the caller) */
static int gen_tables_and_nets(const pf_gen_params *gp, const PfGenDev &G, pf_problem &p) {
	const int nx = G.nx, ny = G.ny, L = G.L;
	p.num_switches = 3;
	p.switches = (pf_switch *)calloc(3, sizeof(pf_switch));
	p.switches[0].buffered = 1; p.switches[0].R = 551.f; p.switches[0].Cin = .77e-15f; p.switches[0].Cout = 4e-15f; p.switches[0].Tdel = 58e-12f;
	p.switches[1].buffered = 1; p.switches[1].R = 0.f; p.switches[1].Cin = 596e-18f; p.switches[1].Cout = 0.f; p.switches[1].Tdel = 101.2e-12f;
	p.switches[2].buffered = 1;
	/* rr_indexed_data (rr_graph_indexed_data.c): buffered segment ⇒ T_linear = Tsw + Rsw*C + 0.5*R*C,
	 * T_quadratic = C_load = 0; DELAY_NORMALIZED: base = T_linear * inv_length for SOURCE/OPIN/CHAN,
	 * 0.95x for IPIN, 0 for SINK */
	p.num_indexed = 6;
	p.indexed = (pf_indexed *)calloc(6, sizeof(pf_indexed));
	{
		float Cw = C_PER_TILE * L, Rw = R_METAL * L;
		float T_lin = p.switches[0].Tdel + p.switches[0].R * Cw + 0.5f * Rw * Cw;
		float inv_len = 1.f / (float)(L < nx ? L : nx);
		float norm = T_lin * inv_len;
		for (int i = 0; i < 6; i++) {
			p.indexed[i].ortho_cost_index = -1; p.indexed[i].seg_index = -1; p.indexed[i].inv_length = -1.f;
			p.indexed[i].T_linear = -1.f; p.indexed[i].T_quadratic = -1.f; p.indexed[i].C_load = -1.f;
			p.indexed[i].base_cost = p.indexed[i].saved_base_cost = norm;
		}
		p.indexed[PF_SINK_COST_INDEX].base_cost = p.indexed[PF_SINK_COST_INDEX].saved_base_cost = 0.f;
		p.indexed[PF_IPIN_COST_INDEX].base_cost = p.indexed[PF_IPIN_COST_INDEX].saved_base_cost = 0.95f * norm;
		p.indexed[PF_IPIN_COST_INDEX].T_linear = p.switches[1].Tdel;
		for (int i = 4; i < 6; i++) {
			p.indexed[i].ortho_cost_index = (i == 4) ? 5 : 4; p.indexed[i].seg_index = 0; p.indexed[i].inv_length = inv_len;
			p.indexed[i].T_linear = T_lin; p.indexed[i].T_quadratic = 0.f; p.indexed[i].C_load = 0.f;
		}
	}

	/* ---- nets */
	const int n = gp->num_nets, spn = gp->sinks_per_net > 0 ? gp->sinks_per_net : 3;
	const int win = gp->window > 0 ? gp->window : 16, bbf = gp->bb_factor >= 0 ? gp->bb_factor : 3;
	p.num_nets = n; p.num_terminals = n * (spn + 1);
	p.net_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
	p.net_terminals = (int32_t *)malloc(sizeof(int32_t) * (size_t)(p.num_terminals ? p.num_terminals : 1));
	p.net_is_global = (uint8_t *)calloc((size_t)(n ? n : 1), 1);
	p.net_bb = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(n ? n : 1));
	p.opin_group_source = (int32_t *)malloc(4); p.opin_group_count = (int32_t *)malloc(4); p.num_opin_groups = 0;
	if (!p.net_ptr || !p.net_terminals || !p.net_is_global || !p.net_bb) { pf_problem_free(&p); return PF_ENOMEM; }
	if ((long long)n > (long long)nx * ny * CLB_OUT) { pf_problem_free(&p); return PF_EINVAL; }
	std::mt19937 rng(gp->seed);
	std::vector<uint16_t> out_used((size_t)(nx + 2) * (ny + 2), 0);
	std::vector<uint8_t> sinks_used((size_t)(nx + 2) * (ny + 2), 0);
	for (int i = 0; i < n; i++) {
		int sx, sy, o;
		for (;;) {
			sx = 1 + (int)(rng() % (unsigned)nx); sy = 1 + (int)(rng() % (unsigned)ny);
			uint16_t m = out_used[(size_t)sx * (ny + 2) + sy];
			if (m == (1u << CLB_OUT) - 1) continue;
			o = (int)(rng() % CLB_OUT);
			while (m & (1u << o)) o = (o + 1) % CLB_OUT;
			out_used[(size_t)sx * (ny + 2) + sy] = (uint16_t)(m | (1u << o));
			break;
		}
		int t0 = i * (spn + 1);
		p.net_ptr[i] = t0;
		p.net_terminals[t0] = pf_gen_tile_base(G, sx, sy) + 1 + o;
		int xmin = sx, xmax = sx, ymin = sy, ymax = sy;
		int cx[64], cy[64];
		for (int k = 0; k < spn; k++) {
			int tx, ty, tries = 0;
			for (;;) {
				int x0 = sx - win < 1 ? 1 : sx - win, x1 = sx + win > nx ? nx : sx + win;
				int y0 = sy - win < 1 ? 1 : sy - win, y1 = sy + win > ny ? ny : sy + win;
				tx = x0 + (int)(rng() % (unsigned)(x1 - x0 + 1)); ty = y0 + (int)(rng() % (unsigned)(y1 - y0 + 1));
				bool bad = (tx == sx && ty == sy) || sinks_used[(size_t)tx * (ny + 2) + ty] >= CLB_IN;
				for (int q = 0; q < k && !bad; q++) if (cx[q] == tx && cy[q] == ty) bad = true;
				if (!bad || ++tries > 1000) break;
			}
			cx[k & 63] = tx; cy[k & 63] = ty;
			sinks_used[(size_t)tx * (ny + 2) + ty]++;
			p.net_terminals[t0 + 1 + k] = pf_gen_tile_base(G, tx, ty);
			xmin = tx < xmin ? tx : xmin; xmax = tx > xmax ? tx : xmax;
			ymin = ty < ymin ? ty : ymin; ymax = ty > ymax ? ty : ymax;
		}
		/* load_route_bb, route_common.c:1065-1123 */
		xmin -= 1; ymin -= 1;
		p.net_bb[4 * i + 0] = xmin - bbf < 0 ? 0 : xmin - bbf;
		p.net_bb[4 * i + 1] = xmax + bbf > nx + 1 ? nx + 1 : xmax + bbf;
		p.net_bb[4 * i + 2] = ymin - bbf < 0 ? 0 : ymin - bbf;
		p.net_bb[4 * i + 3] = ymax + bbf > ny + 1 ? ny + 1 : ymax + bbf;
	}
	p.net_ptr[n] = n * (spn + 1);
	/* VPR defaults (SetupVPR.c:330-605), timing analysis off */
	p.opts.first_iter_pres_fac = 0.5f; p.opts.initial_pres_fac = 0.5f; p.opts.pres_fac_mult = 1.3f; p.opts.acc_fac = 1.f;
	p.opts.bend_cost = 0.f; p.opts.astar_fac = 1.2f; p.opts.max_criticality = 0.99f; p.opts.criticality_exp = 1.f;
	p.opts.max_router_iterations = 50; p.opts.timing_analysis_enabled = 0; p.opts.bb_factor = bbf;
	return PF_OK;
}

extern "C" int pf_gen_grid_nets(const pf_gen_params *gp, pf_problem *out) {
	memset(out, 0, sizeof(*out));
	PfGenDev G;
	std::vector<short> inv(PF_GEN_MAX_W);
	int rc = pf_gen_dev_params(gp, &G, inv.data());
	if (rc != PF_OK) return rc;
	out->nx = G.nx; out->ny = G.ny; out->num_nodes = G.num_nodes; out->num_edges = 0;
	return gen_tables_and_nets(gp, G, *out);
}

extern "C" int pf_gen_grid_problem(const pf_gen_params *gp, pf_problem *out) {
	memset(out, 0, sizeof(*out));
	if (!gp || gp->nx < 2 || gp->ny < 2 || gp->W < 8 || (gp->W & 1) || gp->L < 1 || gp->nx > 30000 || gp->ny > 30000) return PF_EINVAL;
	Gen G;
	G.nx = gp->nx; G.ny = gp->ny; G.W = gp->W; G.L = gp->L; G.io_cap = gp->io_capacity > 0 ? gp->io_capacity : 8;
	G.fc_in = (int)(gp->fc_in * G.W + 0.5f); if (G.fc_in < 2) G.fc_in = 2;
	G.fc_out = (int)(gp->fc_out * G.W + 0.5f); if (G.fc_out < 2) G.fc_out = 2;
	const int nx = G.nx, ny = G.ny, W = G.W, L = G.L;

	/* ---- nodes: tiles (x outer, y inner), classes then pins */
	G.tile_class0.assign((size_t)(nx + 2) * (ny + 2), -1);
	G.tile_pin0.assign((size_t)(nx + 2) * (ny + 2), -1);
	for (int x = 0; x <= nx + 1; x++)
		for (int y = 0; y <= ny + 1; y++) {
			if (G.is_clb(x, y)) {
				G.tile_class0[G.tile(x, y)] = (int)G.type.size();
				G.add_node(PF_SINK, x, y, x, y, 0, PF_SINK_COST_INDEX, CLB_IN, 2, 0.f, 0.f);
				for (int o = 0; o < CLB_OUT; o++) G.add_node(PF_SOURCE, x, y, x, y, 1 + o, PF_SOURCE_COST_INDEX, 1, 2, 0.f, 0.f);
				G.add_node(PF_SINK, x, y, x, y, 11, PF_SINK_COST_INDEX, 1, 2, 0.f, 0.f);
				G.tile_pin0[G.tile(x, y)] = (int)G.type.size();
				for (int p = 0; p < CLB_IN; p++) G.add_node(PF_IPIN, x, y, x, y, p, PF_IPIN_COST_INDEX, 1, 2, 0.f, 0.f);
				for (int o = 0; o < CLB_OUT; o++) G.add_node(PF_OPIN, x, y, x, y, CLB_IN + o, PF_OPIN_COST_INDEX, 1, 2, 0.f, 0.f);
				G.add_node(PF_IPIN, x, y, x, y, 50, PF_IPIN_COST_INDEX, 1, 2, 0.f, 0.f);
			} else if (G.is_io(x, y)) {
				G.tile_class0[G.tile(x, y)] = (int)G.type.size();
				for (int k = 0; k < G.io_cap; k++) {
					G.add_node(PF_SINK, x, y, x, y, 3 * k, PF_SINK_COST_INDEX, 1, 2, 0.f, 0.f);
					G.add_node(PF_SOURCE, x, y, x, y, 3 * k + 1, PF_SOURCE_COST_INDEX, 1, 2, 0.f, 0.f);
					G.add_node(PF_SINK, x, y, x, y, 3 * k + 2, PF_SINK_COST_INDEX, 1, 2, 0.f, 0.f);
				}
				G.tile_pin0[G.tile(x, y)] = (int)G.type.size();
				for (int k = 0; k < G.io_cap; k++) {
					G.add_node(PF_IPIN, x, y, x, y, 3 * k, PF_IPIN_COST_INDEX, 1, 2, 0.f, 0.f);
					G.add_node(PF_OPIN, x, y, x, y, 3 * k + 1, PF_OPIN_COST_INDEX, 1, 2, 0.f, 0.f);
					G.add_node(PF_IPIN, x, y, x, y, 3 * k + 2, PF_IPIN_COST_INDEX, 1, 2, 0.f, 0.f);
				}
			}
		}
	/* ---- wires.  wire_at[channel][(p-1)*W + t] = node covering position p on track t */
	G.chanx_wire.assign((size_t)ny + 1, std::vector<int32_t>());
	for (int y = 0; y <= ny; y++) {
		std::vector<int32_t> &wa = G.chanx_wire[y];
		wa.assign((size_t)nx * W, -1);
		for (int p = 1; p <= nx; p++)
			for (int t = 0; t < W; t++) {
				int s = (t / 2) % L;
				int a = seg_start(p, s, L);
				if (a != p) { wa[(size_t)(p - 1) * W + t] = wa[(size_t)(a - 1) * W + t]; continue; }
				int b = seg_end(p, s, L, nx);
				int len = b - a + 1;
				int id = G.add_node(PF_CHANX, a, y, b, y, t, PF_CHANX_COST_INDEX_START, 1, t & 1, R_METAL * len, C_PER_TILE * len);
				wa[(size_t)(p - 1) * W + t] = id;
			}
	}
	G.chany_wire.assign((size_t)nx + 1, std::vector<int32_t>());
	for (int x = 0; x <= nx; x++) {
		std::vector<int32_t> &wa = G.chany_wire[x];
		wa.assign((size_t)ny * W, -1);
		for (int p = 1; p <= ny; p++)
			for (int t = 0; t < W; t++) {
				int s = (t / 2) % L;
				int a = seg_start(p, s, L);
				if (a != p) { wa[(size_t)(p - 1) * W + t] = wa[(size_t)(a - 1) * W + t]; continue; }
				int b = seg_end(p, s, L, ny);
				int len = b - a + 1;
				int id = G.add_node(PF_CHANY, x, a, x, b, t, PF_CHANX_COST_INDEX_START + 1, 1, t & 1, R_METAL * len, C_PER_TILE * len);
				wa[(size_t)(p - 1) * W + t] = id;
			}
	}
	const int N = (int)G.type.size();
	if (N >= (1 << 30)) return PF_EINVAL;

	/* ---- edges, collected per source node in two passes (count, fill) through a callback */
	std::vector<int32_t> deg((size_t)N + 1, 0);
	std::vector<Edge> tmp;   /* reused small buffer per emitting loop */
	auto for_all_edges = [&](auto &&emit) {
		/* SOURCE→OPIN, IPIN→SINK (delayless switch 2) and pin ↔ channel connection boxes */
		for (int x = 0; x <= nx + 1; x++)
			for (int y = 0; y <= ny + 1; y++) {
				int c0 = G.tile_class0[G.tile(x, y)], p0 = G.tile_pin0[G.tile(x, y)];
				if (c0 < 0) continue;
				bool clb = G.is_clb(x, y);
				int npins = clb ? CLB_PINS : 3 * G.io_cap;
				for (int p = 0; p < npins; p++) {
					int pin = p0 + p;
					bool is_out = clb ? (p >= CLB_IN && p < CLB_IN + CLB_OUT) : (p % 3 == 1);
					int cls = clb ? (p < CLB_IN ? 0 : (p < CLB_IN + CLB_OUT ? 1 + (p - CLB_IN) : 11)) : p;
					if (is_out) emit(c0 + cls, pin, 2); else emit(pin, c0 + cls, 2);
					/* which channel does this pin face?  CLB pins are spread over the four sides */
					int side;   /* 0 top (CHANX y), 1 right (CHANY x), 2 bottom (CHANX y-1), 3 left (CHANY x-1) */
					if (clb) side = p & 3;
					else side = (x == 0) ? 1 : (x == nx + 1) ? 3 : (y == 0) ? 0 : 2;
					bool horiz = (side == 0 || side == 2);
					int chan = horiz ? (side == 0 ? y : y - 1) : (side == 1 ? x : x - 1);
					int pos = horiz ? x : y;                          /* position along the channel */
					const std::vector<int32_t> &wa = horiz ? G.chanx_wire[chan] : G.chany_wire[chan];
					int P = horiz ? nx : ny;
					if (pos < 1 || pos > P) continue;
					if (!is_out) {
						/* connection box: Fc_in tracks, evenly spaced, rotated by the pin number */
						for (int k = 0; k < G.fc_in; k++) {
							int t = (p * 7 + pos + (k * W) / G.fc_in) % W;
							emit(wa[(size_t)(pos - 1) * W + t], pin, 1);
						}
					} else {
						/* an output drives the muxes of wires that START next to the tile:
						 * INC wires with a == pos, DEC wires with b == pos */
						int made = 0;
						for (int k = 0; k < W && made < G.fc_out; k++) {
							int t = (p * 11 + pos * 3 + k) % W;
							int w = wa[(size_t)(pos - 1) * W + t];
							bool starts = horiz ? ((t & 1) ? G.xhigh[w] == pos : G.xlow[w] == pos)
							                    : ((t & 1) ? G.yhigh[w] == pos : G.ylow[w] == pos);
							if (starts) { emit(pin, w, 0); made++; }
						}
					}
				}
			}
		/* wire → wire: at every switch point along the wire (not its own start) one straight
		 * continuation at the end, and one turn onto each perpendicular direction */
		auto turns = [&](int w, bool horiz, int chan, int q, int t) {
			/* switch box (qx,qy): horiz wire in CHANX channel `chan` at SB column q; vertical wire in
			 * CHANY channel `chan` at SB row q */
			int qx = horiz ? q : chan, qy = horiz ? chan : q;
			/* perpendicular channel index and the position where a starting wire begins */
			const std::vector<int32_t> *pw; int P2;
			if (horiz) { if (qx < 0 || qx > nx) return; pw = &G.chany_wire[qx]; P2 = ny; }
			else { if (qy < 0 || qy > ny) return; pw = &G.chanx_wire[qy]; P2 = nx; }
			int base = horiz ? qy : qx;    /* SB coordinate along the perpendicular channel */
			int g = t / 2;
			/* INC wire starting at base+1 */
			if (base + 1 <= P2) {
				int pos = base + 1;
				for (int k = 0; k < W / 2; k++) {
					int t2 = 2 * ((g + qx + qy + k) % (W / 2));
					int w2 = (*pw)[(size_t)(pos - 1) * W + t2];
					bool starts = horiz ? G.ylow[w2] == pos : G.xlow[w2] == pos;
					if (starts) { emit(w, w2, 0); break; }
				}
			}
			/* DEC wire starting at base (its high end) */
			if (base >= 1) {
				int pos = base;
				for (int k = 0; k < W / 2; k++) {
					int t2 = 2 * ((g + 2 * qx + qy + k) % (W / 2)) + 1;
					int w2 = (*pw)[(size_t)(pos - 1) * W + t2];
					bool starts = horiz ? G.yhigh[w2] == pos : G.xhigh[w2] == pos;
					if (starts) { emit(w, w2, 0); break; }
				}
			}
		};
		for (int y = 0; y <= ny; y++) {
			const std::vector<int32_t> &wa = G.chanx_wire[y];
			for (int p = 1; p <= nx; p++)
				for (int t = 0; t < W; t++) {
					int w = wa[(size_t)(p - 1) * W + t];
					if (G.xlow[w] != p) continue;              /* visit each wire once, at its low end */
					int a = G.xlow[w], b = G.xhigh[w];
					if (!(t & 1)) {                            /* INC: enters at a, switch points q = a..b */
						if (b < nx) emit(w, wa[(size_t)b * W + t], 0);
						for (int q = a; q <= b; q++) turns(w, true, y, q, t);
					} else {                                   /* DEC: enters at b, switch points q = a-1..b-1 */
						if (a > 1) emit(w, wa[(size_t)(a - 2) * W + t], 0);
						for (int q = a - 1; q <= b - 1; q++) turns(w, true, y, q, t);
					}
				}
		}
		for (int x = 0; x <= nx; x++) {
			const std::vector<int32_t> &wa = G.chany_wire[x];
			for (int p = 1; p <= ny; p++)
				for (int t = 0; t < W; t++) {
					int w = wa[(size_t)(p - 1) * W + t];
					if (G.ylow[w] != p) continue;
					int a = G.ylow[w], b = G.yhigh[w];
					if (!(t & 1)) {
						if (b < ny) emit(w, wa[(size_t)b * W + t], 0);
						for (int q = a; q <= b; q++) turns(w, false, x, q, t);
					} else {
						if (a > 1) emit(w, wa[(size_t)(a - 2) * W + t], 0);
						for (int q = a - 1; q <= b - 1; q++) turns(w, false, x, q, t);
					}
				}
		}
	};
	long long E = 0;
	for_all_edges([&](int from, int to, int) { (void)to; deg[(size_t)from + 1]++; E++; });
	if (E >= 2147483647ll) return PF_EINVAL;
	for (int i = 0; i < N; i++) deg[(size_t)i + 1] += deg[i];
	std::vector<int32_t> fill(deg.begin(), deg.end() - 1);
	int32_t *edge_to = (int32_t *)malloc(sizeof(int32_t) * (size_t)(E ? E : 1));
	int16_t *edge_sw = (int16_t *)malloc(sizeof(int16_t) * (size_t)(E ? E : 1));
	if (!edge_to || !edge_sw) { free(edge_to); free(edge_sw); return PF_ENOMEM; }
	for_all_edges([&](int from, int to, int sw) { int k = fill[from]++; edge_to[k] = to; edge_sw[k] = (int16_t)sw; });

	/* ---- tables */
	pf_problem &p = *out;
	p.nx = nx; p.ny = ny; p.num_nodes = N; p.num_edges = (int32_t)E;
#define COPYV(dst, vec, T) do { dst = (T *)malloc(sizeof(T) * (vec).size()); if (!dst) { pf_problem_free(&p); return PF_ENOMEM; } memcpy(dst, (vec).data(), sizeof(T) * (vec).size()); } while (0)
	p.edge_to = edge_to; p.edge_sw = edge_sw;
	COPYV(p.xlow, G.xlow, int16_t); COPYV(p.ylow, G.ylow, int16_t); COPYV(p.xhigh, G.xhigh, int16_t); COPYV(p.yhigh, G.yhigh, int16_t);
	COPYV(p.ptc_num, G.ptc, int16_t); COPYV(p.cost_index, G.ci, int16_t); COPYV(p.capacity, G.cap, int16_t);
	COPYV(p.type, G.type, uint8_t); COPYV(p.direction, G.dir, uint8_t); COPYV(p.R, G.R, float); COPYV(p.C, G.C, float);
	COPYV(p.row_ptr, deg, int32_t);
	{
		PfGenDev Gd;
		std::vector<short> inv(PF_GEN_MAX_W);
		int rc = pf_gen_dev_params(gp, &Gd, inv.data());
		if (rc != PF_OK || Gd.num_nodes != N) { pf_problem_free(&p); return rc != PF_OK ? rc : PF_EINVAL; }   /* the closed forms and the tables agree */
		rc = gen_tables_and_nets(gp, Gd, p);
		if (rc != PF_OK) return rc;
	}
	return PF_OK;
}
