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

void pf_result_free(pf_result *r) {
	free(r->trace_ptr); free(r->trace_node); free(r->trace_switch);
	free(r->net_delay); free(r->occ); free(r->iter_stats); free(r->iter_crit);
	memset(r, 0, sizeof(*r));
}
