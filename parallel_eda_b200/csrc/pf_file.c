/*
 * pf_file.c — reader/writer for the flat routing problem / result containers (pf_file.h).
 * Plain C host code; no CUDA.  Compiled into libpf_router.so and also directly into the
 * test-side tools (oracle harness, CPU oracle driver) so they share one format definition.
 */
#include "pf_file.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char PROB_MAGIC[8] = { 'P', 'F', 'P', 'R', 'O', 'B', '0', '1' };
static const char RSLT_MAGIC[8] = { 'P', 'F', 'R', 'S', 'L', 'T', '0', '1' };

static int wr(FILE *f, const void *p, size_t bytes) {
	if (bytes == 0) return 0;
	return fwrite(p, 1, bytes, f) == bytes ? 0 : PF_EIO;
}

static int rd_alloc(FILE *f, void **pp, size_t bytes) {
	*pp = NULL;
	if (bytes == 0) return 0;
	*pp = malloc(bytes);
	if (!*pp) return PF_ENOMEM;
	return fread(*pp, 1, bytes, f) == bytes ? 0 : PF_EIO;
}

#define W(ptr, n) do { if ((rc = wr(f, (ptr), (size_t)(n) * sizeof(*(ptr)))) != 0) goto done; } while (0)
#define R(ptr, n) do { if ((rc = rd_alloc(f, (void **)&(ptr), (size_t)(n) * sizeof(*(ptr)))) != 0) goto done; } while (0)

int pf_problem_write(const char *path, const pf_problem *p) {
	int rc = 0;
	int32_t hdr[16];
	FILE *f = fopen(path, "wb");
	if (!f) return PF_EIO;
	memset(hdr, 0, sizeof(hdr));
	hdr[0] = p->nx; hdr[1] = p->ny; hdr[2] = p->num_nodes; hdr[3] = p->num_edges;
	hdr[4] = p->num_switches; hdr[5] = p->num_indexed; hdr[6] = p->num_nets;
	hdr[7] = p->num_terminals; hdr[8] = p->num_opin_groups;
	hdr[9] = (int32_t)sizeof(pf_router_opts);
	if ((rc = wr(f, PROB_MAGIC, 8)) != 0) goto done;
	if ((rc = wr(f, hdr, sizeof(hdr))) != 0) goto done;
	if ((rc = wr(f, &p->opts, sizeof(p->opts))) != 0) goto done;
	W(p->xlow, p->num_nodes); W(p->ylow, p->num_nodes);
	W(p->xhigh, p->num_nodes); W(p->yhigh, p->num_nodes);
	W(p->ptc_num, p->num_nodes); W(p->cost_index, p->num_nodes);
	W(p->capacity, p->num_nodes); W(p->type, p->num_nodes);
	W(p->direction, p->num_nodes); W(p->R, p->num_nodes); W(p->C, p->num_nodes);
	W(p->row_ptr, (size_t)p->num_nodes + 1);
	W(p->edge_to, p->num_edges); W(p->edge_sw, p->num_edges);
	W(p->switches, p->num_switches); W(p->indexed, p->num_indexed);
	W(p->net_ptr, (size_t)p->num_nets + 1);
	W(p->net_terminals, p->num_terminals);
	W(p->net_is_global, p->num_nets);
	W(p->net_bb, (size_t)p->num_nets * 4);
	W(p->opin_group_source, p->num_opin_groups);
	W(p->opin_group_count, p->num_opin_groups);
done:
	if (fclose(f) != 0 && rc == 0) rc = PF_EIO;
	return rc;
}

int pf_problem_read(const char *path, pf_problem *p) {
	int rc = 0;
	int32_t hdr[16];
	char magic[8];
	FILE *f = fopen(path, "rb");
	memset(p, 0, sizeof(*p));
	if (!f) return PF_EIO;
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, PROB_MAGIC, 8) != 0) { rc = PF_EFORMAT; goto done; }
	if (fread(hdr, 1, sizeof(hdr), f) != sizeof(hdr)) { rc = PF_EIO; goto done; }
	if (hdr[9] != (int32_t)sizeof(pf_router_opts)) { rc = PF_EFORMAT; goto done; }
	p->nx = hdr[0]; p->ny = hdr[1]; p->num_nodes = hdr[2]; p->num_edges = hdr[3];
	p->num_switches = hdr[4]; p->num_indexed = hdr[5]; p->num_nets = hdr[6];
	p->num_terminals = hdr[7]; p->num_opin_groups = hdr[8];
	if (p->num_nodes < 0 || p->num_edges < 0 || p->num_switches < 0 || p->num_indexed < 0
			|| p->num_nets < 0 || p->num_terminals < 0 || p->num_opin_groups < 0) {
		rc = PF_EFORMAT; goto done;
	}
	if (fread(&p->opts, 1, sizeof(p->opts), f) != sizeof(p->opts)) { rc = PF_EIO; goto done; }
	R(p->xlow, p->num_nodes); R(p->ylow, p->num_nodes);
	R(p->xhigh, p->num_nodes); R(p->yhigh, p->num_nodes);
	R(p->ptc_num, p->num_nodes); R(p->cost_index, p->num_nodes);
	R(p->capacity, p->num_nodes); R(p->type, p->num_nodes);
	R(p->direction, p->num_nodes); R(p->R, p->num_nodes); R(p->C, p->num_nodes);
	R(p->row_ptr, (size_t)p->num_nodes + 1);
	R(p->edge_to, p->num_edges); R(p->edge_sw, p->num_edges);
	R(p->switches, p->num_switches); R(p->indexed, p->num_indexed);
	R(p->net_ptr, (size_t)p->num_nets + 1);
	R(p->net_terminals, p->num_terminals);
	R(p->net_is_global, p->num_nets);
	R(p->net_bb, (size_t)p->num_nets * 4);
	R(p->opin_group_source, p->num_opin_groups);
	R(p->opin_group_count, p->num_opin_groups);
done:
	fclose(f);
	if (rc != 0) pf_problem_free(p);
	return rc;
}

void pf_problem_free(pf_problem *p) {
	free(p->xlow); free(p->ylow); free(p->xhigh); free(p->yhigh);
	free(p->ptc_num); free(p->cost_index); free(p->capacity); free(p->type);
	free(p->direction); free(p->R); free(p->C);
	free(p->row_ptr); free(p->edge_to); free(p->edge_sw);
	free(p->switches); free(p->indexed);
	free(p->net_ptr); free(p->net_terminals); free(p->net_is_global); free(p->net_bb);
	free(p->opin_group_source); free(p->opin_group_count);
	memset(p, 0, sizeof(*p));
}

#define FAIL(...) do { if (msg && msg_len > 0) snprintf(msg, (size_t)msg_len, __VA_ARGS__); return PF_EINVAL; } while (0)

int pf_problem_check(const pf_problem *p, char *msg, int msg_len) {
	int i, k;
	if (msg && msg_len > 0) msg[0] = 0;
	if (p->nx <= 0 || p->ny <= 0) FAIL("bad grid %d x %d", p->nx, p->ny);
	if (p->num_nodes <= 0) FAIL("no rr nodes");
	if (p->num_indexed < PF_CHANX_COST_INDEX_START) FAIL("num_indexed %d < 4", p->num_indexed);
	if (p->row_ptr[0] != 0 || p->row_ptr[p->num_nodes] != p->num_edges) FAIL("row_ptr ends do not match num_edges");
	for (i = 0; i < p->num_nodes; i++) {
		if (p->row_ptr[i + 1] < p->row_ptr[i]) FAIL("row_ptr not monotone at node %d", i);
		if (p->row_ptr[i + 1] - p->row_ptr[i] > 32767) FAIL("node %d has more than 32767 edges", i);
		if (p->type[i] > PF_CHANY) FAIL("node %d has type %d", i, p->type[i]);
		if (p->cost_index[i] < 0 || p->cost_index[i] >= p->num_indexed) FAIL("node %d cost_index %d", i, p->cost_index[i]);
		if (p->xlow[i] > p->xhigh[i] || p->ylow[i] > p->yhigh[i]) FAIL("node %d has inverted span", i);
		if (p->xlow[i] < 0 || p->ylow[i] < 0 || p->xhigh[i] > p->nx + 1 || p->yhigh[i] > p->ny + 1) FAIL("node %d outside grid", i);
		if (p->capacity[i] < 0) FAIL("node %d negative capacity", i);
	}
	for (k = 0; k < p->num_edges; k++) {
		if (p->edge_to[k] < 0 || p->edge_to[k] >= p->num_nodes) FAIL("edge %d targets node %d", k, p->edge_to[k]);
		if (p->edge_sw[k] < 0 || p->edge_sw[k] >= p->num_switches) FAIL("edge %d uses switch %d", k, p->edge_sw[k]);
	}
	for (i = 0; i < p->num_indexed; i++) {
		int o = p->indexed[i].ortho_cost_index;
		if (i >= PF_CHANX_COST_INDEX_START && (o < 0 || o >= p->num_indexed)) FAIL("indexed row %d ortho %d", i, o);
	}
	if (p->net_ptr[0] != 0 || p->net_ptr[p->num_nets] != p->num_terminals) FAIL("net_ptr ends do not match num_terminals");
	for (i = 0; i < p->num_nets; i++)                        /* completely, before net_terminals is indexed through it */
		if (p->net_ptr[i + 1] < p->net_ptr[i]) FAIL("net_ptr not monotone at net %d", i);
	for (i = 0; i < p->num_nets; i++) {
		int b = p->net_ptr[i], e = p->net_ptr[i + 1];
		if (e < b) FAIL("net_ptr not monotone at net %d", i);
		if (e == b) FAIL("net %d has no terminals", i);
		for (k = b; k < e; k++) {
			int n = p->net_terminals[k];
			if (p->net_is_global[i]) continue; /* global nets may carry OPEN terminals */
			if (n < 0 || n >= p->num_nodes) FAIL("net %d terminal %d is node %d", i, k - b, n);
			if (k == b && p->type[n] != PF_SOURCE) FAIL("net %d terminal 0 is not a SOURCE", i);
			if (k > b && p->type[n] != PF_SINK) FAIL("net %d terminal %d is not a SINK", i, k - b);
		}
		if (p->net_bb[4 * i + 0] > p->net_bb[4 * i + 1] || p->net_bb[4 * i + 2] > p->net_bb[4 * i + 3]) FAIL("net %d inverted bb", i);
	}
	for (i = 0; i < p->num_opin_groups; i++) {
		int s = p->opin_group_source[i];
		if (s < 0 || s >= p->num_nodes || p->type[s] != PF_SOURCE) FAIL("opin group %d source %d", i, s);
		if (p->opin_group_count[i] < 0 || p->opin_group_count[i] > p->row_ptr[s + 1] - p->row_ptr[s]) FAIL("opin group %d count", i);
	}
	return PF_OK;
}

int pf_result_write(const char *path, const pf_result *r) {
	int rc = 0;
	int32_t hdr[16];
	int32_t ntrace = r->trace_ptr ? r->trace_ptr[r->num_nets] : 0;
	FILE *f = fopen(path, "wb");
	if (!f) return PF_EIO;
	memset(hdr, 0, sizeof(hdr));
	hdr[0] = r->success; hdr[1] = r->iterations; hdr[2] = r->serial_num; hdr[3] = r->total_wirelength;
	hdr[4] = r->num_nets; hdr[5] = ntrace; hdr[6] = r->num_terminals; hdr[7] = r->num_nodes;
	hdr[8] = r->num_iter_stats; hdr[9] = r->num_crit_iters; hdr[10] = (int32_t)sizeof(pf_iter_stats);
	if ((rc = wr(f, RSLT_MAGIC, 8)) != 0) goto done;
	if ((rc = wr(f, hdr, sizeof(hdr))) != 0) goto done;
	W(r->trace_ptr, (size_t)r->num_nets + 1);
	W(r->trace_node, ntrace); W(r->trace_switch, ntrace);
	W(r->net_delay, r->num_terminals);
	W(r->occ, r->num_nodes);
	W(r->iter_stats, r->num_iter_stats);
	W(r->iter_crit, (size_t)r->num_crit_iters * (size_t)r->num_terminals);
done:
	if (fclose(f) != 0 && rc == 0) rc = PF_EIO;
	return rc;
}

int pf_result_read(const char *path, pf_result *r) {
	int rc = 0;
	int32_t hdr[16], ntrace;
	char magic[8];
	FILE *f = fopen(path, "rb");
	memset(r, 0, sizeof(*r));
	if (!f) return PF_EIO;
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, RSLT_MAGIC, 8) != 0) { rc = PF_EFORMAT; goto done; }
	if (fread(hdr, 1, sizeof(hdr), f) != sizeof(hdr)) { rc = PF_EIO; goto done; }
	if (hdr[10] != (int32_t)sizeof(pf_iter_stats)) { rc = PF_EFORMAT; goto done; }
	r->success = hdr[0]; r->iterations = hdr[1]; r->serial_num = hdr[2]; r->total_wirelength = hdr[3];
	r->num_nets = hdr[4]; ntrace = hdr[5]; r->num_terminals = hdr[6]; r->num_nodes = hdr[7];
	r->num_iter_stats = hdr[8]; r->num_crit_iters = hdr[9];
	if (r->num_nets < 0 || ntrace < 0 || r->num_terminals < 0 || r->num_nodes < 0
			|| r->num_iter_stats < 0 || r->num_crit_iters < 0) { rc = PF_EFORMAT; goto done; }
	R(r->trace_ptr, (size_t)r->num_nets + 1);
	R(r->trace_node, ntrace); R(r->trace_switch, ntrace);
	R(r->net_delay, r->num_terminals);
	R(r->occ, r->num_nodes);
	R(r->iter_stats, r->num_iter_stats);
	R(r->iter_crit, (size_t)r->num_crit_iters * (size_t)r->num_terminals);
	if (r->trace_ptr[r->num_nets] != ntrace) rc = PF_EFORMAT;
done:
	fclose(f);
	if (rc != 0) pf_result_free(r);
	return rc;
}

/* The router hands out result arrays from its cache of pinned host buffers (pf_router.cpp: the device copies straight into
 * them, and a buffer that has been faulted in once is used again); it registers this hook to get them back.  Everything else
 * (file readers, callers' own arrays) is plain malloc memory. */
static int (*g_release_hook)(void *) = 0;
void pf_result_set_release_hook(int (*fn)(void *)) { g_release_hook = fn; }
static void result_release(void *p) {
	if (!p) return;
	if (g_release_hook && g_release_hook(p)) return;
	free(p);
}

void pf_result_free(pf_result *r) {
	result_release(r->trace_ptr); result_release(r->trace_node); result_release(r->trace_switch);
	result_release(r->net_delay); result_release(r->occ); free(r->iter_stats); free(r->iter_crit);
	memset(r, 0, sizeof(*r));
}

/* ------------------------------------------------------------------ timing graph / STA vectors */
static const char TIMG_MAGIC[8] = { 'P', 'F', 'T', 'I', 'M', 'G', '0', '1' };
static const char STAV_MAGIC[8] = { 'P', 'F', 'S', 'T', 'A', 'V', '0', '1' };

int pf_timing_graph_write(const char *path, const pf_timing_graph *g) {
	int rc = 0;
	int32_t hdr[16];
	FILE *f = fopen(path, "wb");
	if (!f) return PF_EIO;
	memset(hdr, 0, sizeof(hdr));
	hdr[0] = g->num_tnodes; hdr[1] = g->num_tedges; hdr[2] = g->num_levels; hdr[3] = g->num_domains; hdr[4] = g->num_nets;
	hdr[5] = g->num_overrides;      /* 0 in files written before the field existed: the arrays follow at the end */
	if ((rc = wr(f, TIMG_MAGIC, 8)) != 0) goto done;
	if ((rc = wr(f, hdr, sizeof(hdr))) != 0) goto done;
	W(g->edge_ptr, (size_t)g->num_tnodes + 1);
	W(g->edge_to, g->num_tedges); W(g->edge_Tdel, g->num_tedges);
	W(g->type, g->num_tnodes); W(g->clock_domain, g->num_tnodes); W(g->clock_delay, g->num_tnodes);
	W(g->level_ptr, (size_t)g->num_levels + 1); W(g->level_nodes, g->num_tnodes);
	W(g->constraint, (size_t)g->num_domains * (size_t)g->num_domains);
	W(g->net_driver, g->num_nets);
	W(g->override_domain, g->num_overrides); W(g->override_tnode, g->num_overrides); W(g->override_constraint, g->num_overrides);
done:
	if (fclose(f) != 0 && rc == 0) rc = PF_EIO;
	return rc;
}

int pf_timing_graph_read(const char *path, pf_timing_graph *g) {
	int rc = 0;
	int32_t hdr[16];
	char magic[8];
	FILE *f = fopen(path, "rb");
	memset(g, 0, sizeof(*g));
	if (!f) return PF_EIO;
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, TIMG_MAGIC, 8) != 0) { rc = PF_EFORMAT; goto done; }
	if (fread(hdr, 1, sizeof(hdr), f) != sizeof(hdr)) { rc = PF_EIO; goto done; }
	g->num_tnodes = hdr[0]; g->num_tedges = hdr[1]; g->num_levels = hdr[2]; g->num_domains = hdr[3]; g->num_nets = hdr[4]; g->num_overrides = hdr[5];
	if (g->num_tnodes < 0 || g->num_tedges < 0 || g->num_levels < 0 || g->num_domains < 0 || g->num_nets < 0 || g->num_overrides < 0) { rc = PF_EFORMAT; goto done; }
	R(g->edge_ptr, (size_t)g->num_tnodes + 1);
	R(g->edge_to, g->num_tedges); R(g->edge_Tdel, g->num_tedges);
	R(g->type, g->num_tnodes); R(g->clock_domain, g->num_tnodes); R(g->clock_delay, g->num_tnodes);
	R(g->level_ptr, (size_t)g->num_levels + 1); R(g->level_nodes, g->num_tnodes);
	R(g->constraint, (size_t)g->num_domains * (size_t)g->num_domains);
	R(g->net_driver, g->num_nets);
	R(g->override_domain, g->num_overrides); R(g->override_tnode, g->num_overrides); R(g->override_constraint, g->num_overrides);
done:
	fclose(f);
	if (rc != 0) pf_timing_graph_free(g);
	return rc;
}

void pf_timing_graph_free(pf_timing_graph *g) {
	free(g->edge_ptr); free(g->edge_to); free(g->edge_Tdel); free(g->type); free(g->clock_domain); free(g->clock_delay);
	free(g->level_ptr); free(g->level_nodes); free(g->constraint); free(g->net_driver);
	free(g->override_domain); free(g->override_tnode); free(g->override_constraint);
	memset(g, 0, sizeof(*g));
}

#define TFAIL(...) do { if (msg) snprintf(msg, (size_t)msg_len, __VA_ARGS__); return PF_EINVAL; } while (0)
int pf_timing_graph_check(const pf_timing_graph *g, const int32_t *net_ptr, char *msg, int msg_len) {
	int i, k, lv;
	int32_t *level_of;
	if (g->num_tnodes <= 0 || !g->edge_ptr || g->edge_ptr[0] != 0 || g->edge_ptr[g->num_tnodes] != g->num_tedges) TFAIL("edge_ptr does not span the edges");
	if (g->num_levels <= 0 || g->level_ptr[0] != 0 || g->level_ptr[g->num_levels] != g->num_tnodes) TFAIL("level_ptr does not span the tnodes");
	/* both offset arrays completely before anything is indexed through them */
	for (lv = 0; lv < g->num_levels; lv++) if (g->level_ptr[lv + 1] < g->level_ptr[lv]) TFAIL("level_ptr not monotonic at level %d", lv);
	for (i = 0; i < g->num_tnodes; i++) if (g->edge_ptr[i + 1] < g->edge_ptr[i]) TFAIL("edge_ptr not monotonic at tnode %d", i);
	level_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->num_tnodes);
	if (!level_of) return PF_ENOMEM;
	for (i = 0; i < g->num_tnodes; i++) level_of[i] = -1;
	for (lv = 0; lv < g->num_levels; lv++) {
		if (g->level_ptr[lv + 1] < g->level_ptr[lv]) { free(level_of); TFAIL("level_ptr not monotonic at level %d", lv); }
		for (k = g->level_ptr[lv]; k < g->level_ptr[lv + 1]; k++) {
			int n = g->level_nodes[k];
			if (n < 0 || n >= g->num_tnodes || level_of[n] >= 0) { free(level_of); TFAIL("level list entry %d is not a fresh tnode", k); }
			level_of[n] = lv;
		}
	}
	for (i = 0; i < g->num_tnodes; i++) {
		if (g->edge_ptr[i + 1] < g->edge_ptr[i]) { free(level_of); TFAIL("edge_ptr not monotonic at tnode %d", i); }
		if (g->clock_domain[i] < -1 || g->clock_domain[i] >= g->num_domains) { free(level_of); TFAIL("tnode %d: clock domain %d", i, g->clock_domain[i]); }
		for (k = g->edge_ptr[i]; k < g->edge_ptr[i + 1]; k++) {
			int to = g->edge_to[k];
			if (to < 0 || to >= g->num_tnodes || level_of[to] <= level_of[i]) { free(level_of); TFAIL("tedge %d of tnode %d does not lead to a later level", k, i); }
		}
	}
	free(level_of);
	for (i = 0; i < g->num_nets; i++) {
		int d = g->net_driver[i];
		if (d < -1 || d >= g->num_tnodes) TFAIL("net %d: driver tnode %d", i, d);
		if (d >= 0 && net_ptr && g->edge_ptr[d + 1] - g->edge_ptr[d] != net_ptr[i + 1] - net_ptr[i] - 1)
			TFAIL("net %d: driver tnode %d has %d out-edges, the net %d sinks", i, d, g->edge_ptr[d + 1] - g->edge_ptr[d], net_ptr[i + 1] - net_ptr[i] - 1);
	}
	if (g->num_overrides < 0 || (g->num_overrides > 0 && (!g->override_domain || !g->override_tnode || !g->override_constraint))) TFAIL("override arrays missing");
	for (i = 0; i < g->num_overrides; i++) {
		int n = g->override_tnode[i], dmn = g->override_domain[i];
		if (n < 0 || n >= g->num_tnodes || g->edge_ptr[n + 1] != g->edge_ptr[n]) TFAIL("override %d: tnode %d is not a sink", i, n);
		if (dmn < 0 || dmn >= g->num_domains) TFAIL("override %d: source domain %d", i, dmn);
		if (i > 0 && (g->override_tnode[i - 1] > n || (g->override_tnode[i - 1] == n && g->override_domain[i - 1] >= dmn)))
			TFAIL("override %d: entries are not sorted by (tnode, domain) or a pair appears twice", i);
	}
	return PF_OK;
}

int pf_sta_vectors_write(const char *path, const pf_sta_vectors *v) {
	int rc = 0;
	int32_t hdr[16];
	FILE *f = fopen(path, "wb");
	if (!f) return PF_EIO;
	memset(hdr, 0, sizeof(hdr));
	hdr[0] = v->num_terminals; hdr[1] = v->num_calls;
	if ((rc = wr(f, STAV_MAGIC, 8)) != 0) goto done;
	if ((rc = wr(f, hdr, sizeof(hdr))) != 0) goto done;
	W(v->net_delay, (size_t)v->num_calls * (size_t)v->num_terminals);
	W(v->crit, (size_t)v->num_calls * (size_t)v->num_terminals);
	W(v->cpd, v->num_calls);
done:
	if (fclose(f) != 0 && rc == 0) rc = PF_EIO;
	return rc;
}

int pf_sta_vectors_read(const char *path, pf_sta_vectors *v) {
	int rc = 0;
	int32_t hdr[16];
	char magic[8];
	FILE *f = fopen(path, "rb");
	memset(v, 0, sizeof(*v));
	if (!f) return PF_EIO;
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, STAV_MAGIC, 8) != 0) { rc = PF_EFORMAT; goto done; }
	if (fread(hdr, 1, sizeof(hdr), f) != sizeof(hdr)) { rc = PF_EIO; goto done; }
	v->num_terminals = hdr[0]; v->num_calls = hdr[1];
	if (v->num_terminals < 0 || v->num_calls < 0) { rc = PF_EFORMAT; goto done; }
	R(v->net_delay, (size_t)v->num_calls * (size_t)v->num_terminals);
	R(v->crit, (size_t)v->num_calls * (size_t)v->num_terminals);
	R(v->cpd, v->num_calls);
done:
	fclose(f);
	if (rc != 0) pf_sta_vectors_free(v);
	return rc;
}

void pf_sta_vectors_free(pf_sta_vectors *v) {
	free(v->net_delay); free(v->crit); free(v->cpd);
	memset(v, 0, sizeof(*v));
}
