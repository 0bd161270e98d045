/*
 * harness.cxx — TEST INFRASTRUCTURE.  Driver + dump hooks around the UNMODIFIED reference
 * serial router (chinhau5/parallel_eda, VPR 7.0).  Linked with libvpr_ref.a, which is the
 * reference compiled from /root/reference by oracle/ref_build/Makefile.  No reference source
 * is copied here; this file only calls the reference's public functions and reads its globals.
 *
 *   vpr_ref flow <arch.xml> <circuit> [vpr options]
 *        the stock flow: vpr_init → [vpr_pack] → vpr_init_pre_place_and_route → place_and_route
 *        (reference vpr/SRC/base/vpr_api.c:164,390,241 and place_and_route.c:250; the fork's own
 *        place_and_route_new cannot reach the timing-driven router).
 *        env PF_DUMP_PROBLEM=<file>  write the flat problem seen by try_timing_driven_route
 *        env PF_DUMP_RESULT=<file>   write traces / delays / per-iteration criticalities
 *        env PF_DUMP_TGRAPH=<file>   write the flat timing graph (pf_timing_graph) do_timing_analysis runs on
 *        env PF_DUMP_NAMES=<file>    write net / block names, IO tiles, global-net pins (pf_names, include/pf_text.h)
 *        env PF_DUMP_NETLIST=<file>  write block[] / clb_net[] as read_netlist left them (text; golden of pf_net_read)
 *        env PF_ADAPTER_ROUTE_FILE=<file>  also write the .route file through integration/vpr_text_adapter.cxx
 *        env PF_DUMP_AT_SUCCESS=1    write PF_DUMP_RESULT as soon as the routing is legal (before the reference's DEBUG delay check)
 *        env PF_DUMP_STA=<file>      write every (net_delay in, timing_criticality out, cpd) of the run's STA calls
 *        env PF_DUMP_STA_FINAL=<file> write routing_stats' analysis of the finished routing (is_final_analysis): criticalities in <file>, slacks in <file>.slack
 *   vpr_ref inject <problem.pfp> [--result out.pfr] [--crit golden.pfr] [--max_iters K]
 *                  [--limit_nets M]
 *        load a flat problem into the reference's globals and call the reference's own
 *        try_timing_driven_route (route_timing.c:85) on it.  With --crit the recorded
 *        per-iteration criticalities are replayed in place of the STA (timing-driven mode);
 *        otherwise timing_analysis_enabled = FALSE (criticality 0, route_timing.c:116-120).
 *
 * Interposed symbols (see Makefile): try_timing_driven_route (call site route_common.c:500),
 * do_timing_analysis / load_timing_graph_net_delays / get_critical_path_delay /
 * feasible_routing (call sites route_timing.c:299-308,241).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <chrono>
#include <vector>

#include "vpr_types.h"
#include "vpr_api.h"
#include "globals.h"
#include "route_export.h"
#include "route_common.h"
#include "route_tree_timing.h"
#include "route_timing.h"
#include "path_delay.h"
#include "path_delay2.h"
#include "read_sdc.h"
#include "net_delay.h"
#include "place_and_route.h"
#include "stats.h"
#include "rr_graph.h"
#include "read_place.h"

#include <string>
#include "../../include/pf_file.h"
#include "../../include/pf_text.h"

/* symbols the reference's main.c normally provides (main.c:60-62,253,287) */
std::chrono::time_point<std::chrono::high_resolution_clock> program_start;
char *s_circuit_name = nullptr;
void print_context(int, int) {}
void get_mem_usage(unsigned long &vm, unsigned long &rss) { vm = 0; rss = 0; }

/* originals (declared under their real names by the reference headers) */
boolean try_timing_driven_route(struct s_router_opts router_opts, float **net_delay, t_slack *slacks,
		t_ivec **clb_opins_used_locally, boolean timing_analysis_enabled);
void do_timing_analysis(t_slack *slacks, boolean is_prepacked, boolean do_lut_input_balancing, boolean is_final_analysis);
void load_timing_graph_net_delays(float **net_delay);
float get_critical_path_delay(void);
boolean feasible_routing(void);
boolean try_breadth_first_route(struct s_router_opts router_opts, t_ivec **clb_opins_used_locally, int width_fac);
boolean pf_hook_try_breadth_first_route(struct s_router_opts router_opts, t_ivec **clb_opins_used_locally, int width_fac);
extern struct s_bb *route_bb;
extern t_rr_node_route_inf *rr_node_route_inf;
void alloc_route_static_structs(void);

static double now_s() {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------ capture state */
static bool g_inject = false;            /* inject mode: no timing graph exists */
static pf_result g_replay;               /* --crit: criticalities to replay */
static bool g_have_replay = false;
static int g_iter = 0;                   /* iterations completed (feasible_routing calls) */
static std::vector<pf_iter_stats> g_stats;
static std::vector<float> g_crit;        /* [iter][num_terminals] */
static std::vector<int> g_net_ptr;       /* terminals prefix */
static double g_t0 = 0;
static std::vector<double> g_iter_time;
static float g_last_cpd = 0;
static double g_sta_seconds = 0;
static std::vector<float> g_sta_delay, g_sta_crit, g_sta_cpd;   /* golden vectors of the reference's STA calls */

static void build_net_ptr() {
	g_net_ptr.assign(num_nets + 1, 0);
	for (int i = 0; i < num_nets; i++)
		g_net_ptr[i + 1] = g_net_ptr[i] + clb_net[i].num_sinks + 1;
}

static void record_crit(t_slack *slacks) {
	size_t base = g_crit.size();
	g_crit.resize(base + g_net_ptr[num_nets], 0.f);
	for (int i = 0; i < num_nets; i++) {
		if (clb_net[i].is_global) continue;
		for (int k = 1; k <= clb_net[i].num_sinks; k++)
			g_crit[base + g_net_ptr[i] + k] = slacks->timing_criticality[i][k];
	}
}

/* ------------------------------------------------------------------ hooks */
static float **g_net_delay_arg = NULL;   /* net_delay argument of the running try_timing_driven_route */
static void export_result(const char *path, boolean ok, float **net_delay);
boolean pf_hook_feasible_routing(void) {
	pf_iter_stats st;
	memset(&st, 0, sizeof(st));
	for (int i = 0; i < num_rr_nodes; i++)
		if (rr_node[i].occ > rr_node[i].capacity) st.overused_nodes++;
	for (int i = 0; i < num_nets; i++)
		if (!clb_net[i].is_global) st.nets_routed++;
	g_stats.push_back(st);
	g_iter_time.push_back(now_s() - g_t0);
	g_iter++;
	boolean ok = feasible_routing();
	/* PF_DUMP_AT_SUCCESS=1: write the result the moment the routing is legal — before the reference's own DEBUG
	 * cross-check timing_driven_check_net_delays (route_timing.c:246), which aborts the run on nets that connect twice
	 * to one SINK (its incremental and from-scratch delays disagree there: a limitation of the reference itself) */
	if (ok && getenv("PF_DUMP_AT_SUCCESS") && getenv("PF_DUMP_RESULT")) export_result(getenv("PF_DUMP_RESULT"), TRUE, g_net_delay_arg);
	return ok;
}

void pf_hook_load_timing_graph_net_delays(float **net_delay) {
	if (!g_inject) {
		load_timing_graph_net_delays(net_delay);
		size_t base = g_sta_delay.size();
		g_sta_delay.resize(base + g_net_ptr[num_nets], 0.f);
		for (int i = 0; i < num_nets; i++)
			for (int k = 1; k <= clb_net[i].num_sinks; k++) g_sta_delay[base + g_net_ptr[i] + k] = net_delay[i][k];
	}
}

static t_slack *g_slacks = NULL;
void pf_hook_do_timing_analysis(t_slack *slacks, boolean a, boolean b, boolean c) {
	if (!g_inject) {
		double t_sta = now_s();
		do_timing_analysis(slacks, a, b, c);
		g_sta_seconds += now_s() - t_sta;
		size_t base = g_sta_crit.size();          /* every net, global ones too: the analysis covers them */
		g_sta_crit.resize(base + g_net_ptr[num_nets], 0.f);
		for (int i = 0; i < num_nets; i++)
			for (int k = 1; k <= clb_net[i].num_sinks; k++) g_sta_crit[base + g_net_ptr[i] + k] = slacks->timing_criticality[i][k];
	} else if (g_have_replay) {
		/* criticalities for iteration g_iter+1 (g_iter iterations are complete) */
		int it = g_iter < g_replay.num_crit_iters ? g_iter : g_replay.num_crit_iters - 1;
		const float *src = g_replay.iter_crit + (size_t)it * g_replay.num_terminals;
		for (int i = 0; i < num_nets; i++)
			for (int k = 1; k <= clb_net[i].num_sinks; k++)
				slacks->timing_criticality[i][k] = src[g_net_ptr[i] + k];
	}
	record_crit(slacks);
}

/* base/stats.c is compiled with -Ddo_timing_analysis=pf_hook_final_timing_analysis -Dload_timing_graph_net_delays=
 * pf_hook_final_load_net_delays: routing_stats' analysis of the finished routing (is_final_analysis = TRUE) runs as always;
 * with PF_DUMP_STA_FINAL=<file> its input (net delays) and outputs are written as two one-call pf_sta_vectors containers:
 * <file> carries the criticalities, <file>.slack the slacks in the same slot; both carry the critical path delay. */
static std::vector<float> g_final_delay;
void pf_hook_final_load_net_delays(float **net_delay) {
	load_timing_graph_net_delays(net_delay);
	build_net_ptr();
	g_final_delay.assign(g_net_ptr[num_nets], 0.f);
	for (int i = 0; i < num_nets; i++)
		for (int k = 1; k <= clb_net[i].num_sinks; k++) g_final_delay[g_net_ptr[i] + k] = net_delay[i][k];
}
void pf_hook_final_timing_analysis(t_slack *slacks, boolean a, boolean b, boolean c) {
	do_timing_analysis(slacks, a, b, c);
	const char *path = getenv("PF_DUMP_STA_FINAL");
	if (!path || g_inject || g_final_delay.empty()) return;
	std::vector<float> crit(g_net_ptr[num_nets], 0.f), slk(g_net_ptr[num_nets], 0.f);
	for (int i = 0; i < num_nets; i++)
		for (int k = 1; k <= clb_net[i].num_sinks; k++) { crit[g_net_ptr[i] + k] = slacks->timing_criticality[i][k]; slk[g_net_ptr[i] + k] = slacks->slack[i][k]; }
	for (int i = 0; i < num_nets; i++) slk[g_net_ptr[i]] = 1.e30f;      /* the driver slot: as pf_sta_analyze_final leaves it */
	float cpd = get_critical_path_delay();
	pf_sta_vectors v;
	memset(&v, 0, sizeof(v));
	v.num_terminals = g_net_ptr[num_nets]; v.num_calls = 1; v.net_delay = g_final_delay.data(); v.cpd = &cpd;
	v.crit = crit.data();
	int rc = pf_sta_vectors_write(path, &v);
	std::string p2 = std::string(path) + ".slack";
	v.crit = slk.data();
	rc |= pf_sta_vectors_write(p2.c_str(), &v);
	fprintf(stderr, "PF_REF wrote final analysis %s (+ .slack): %d terminals, cpd %g ns, is_final %d (rc %d)\n", path, v.num_terminals, cpd, (int)c, rc);
}

float pf_hook_get_critical_path_delay(void) {
	g_last_cpd = g_inject ? 0.f : get_critical_path_delay();
	if (!g_inject) g_sta_cpd.push_back(g_last_cpd);
	if (!g_stats.empty()) g_stats.back().crit_path_delay = g_last_cpd;
	return g_last_cpd;
}

/* ------------------------------------------------------------------ export */
static void export_problem(const char *path, struct s_router_opts ro, boolean timing_enabled,
		t_ivec **clb_opins_used_locally) {
	pf_problem p;
	memset(&p, 0, sizeof(p));
	int N = num_rr_nodes;
	long E = 0;
	for (int i = 0; i < N; i++) E += rr_node[i].num_edges;
	p.nx = nx; p.ny = ny; p.num_nodes = N; p.num_edges = (int)E;
	std::vector<int16_t> xl(N), yl(N), xh(N), yh(N), ptc(N), ci(N), cap(N);
	std::vector<uint8_t> ty(N), dir(N);
	std::vector<float> R(N), C(N);
	std::vector<int32_t> row(N + 1), to(E);
	std::vector<int16_t> sw(E);
	long e = 0;
	for (int i = 0; i < N; i++) {
		xl[i] = rr_node[i].xlow; yl[i] = rr_node[i].ylow; xh[i] = rr_node[i].xhigh; yh[i] = rr_node[i].yhigh;
		ptc[i] = rr_node[i].ptc_num; ci[i] = rr_node[i].cost_index; cap[i] = rr_node[i].capacity;
		ty[i] = (uint8_t)rr_node[i].type; dir[i] = (uint8_t)rr_node[i].direction;
		R[i] = rr_node[i].R; C[i] = rr_node[i].C;
		row[i] = (int32_t)e;
		for (int k = 0; k < rr_node[i].num_edges; k++) { to[e] = rr_node[i].edges[k]; sw[e] = rr_node[i].switches[k]; e++; }
	}
	row[N] = (int32_t)e;
	p.xlow = xl.data(); p.ylow = yl.data(); p.xhigh = xh.data(); p.yhigh = yh.data();
	p.ptc_num = ptc.data(); p.cost_index = ci.data(); p.capacity = cap.data();
	p.type = ty.data(); p.direction = dir.data(); p.R = R.data(); p.C = C.data();
	p.row_ptr = row.data(); p.edge_to = to.data(); p.edge_sw = sw.data();

	/* number of switches: largest id used + 1 (det_routing_arch is not visible here) */
	int nsw = 0;
	for (long k = 0; k < E; k++) if (sw[k] + 1 > nsw) nsw = sw[k] + 1;
	std::vector<pf_switch> sws(nsw);
	for (int s = 0; s < nsw; s++) {
		sws[s].buffered = switch_inf[s].buffered; sws[s].R = switch_inf[s].R; sws[s].Cin = switch_inf[s].Cin;
		sws[s].Cout = switch_inf[s].Cout; sws[s].Tdel = switch_inf[s].Tdel;
	}
	p.num_switches = nsw; p.switches = sws.data();
	std::vector<pf_indexed> idx(num_rr_indexed_data);
	for (int i = 0; i < num_rr_indexed_data; i++) {
		idx[i].base_cost = rr_indexed_data[i].base_cost; idx[i].saved_base_cost = rr_indexed_data[i].saved_base_cost;
		idx[i].ortho_cost_index = rr_indexed_data[i].ortho_cost_index; idx[i].seg_index = rr_indexed_data[i].seg_index;
		idx[i].inv_length = rr_indexed_data[i].inv_length; idx[i].T_linear = rr_indexed_data[i].T_linear;
		idx[i].T_quadratic = rr_indexed_data[i].T_quadratic; idx[i].C_load = rr_indexed_data[i].C_load;
	}
	p.num_indexed = num_rr_indexed_data; p.indexed = idx.data();

	build_net_ptr();
	std::vector<int32_t> term(g_net_ptr[num_nets]), bb(4 * (size_t)num_nets);
	std::vector<uint8_t> glob(num_nets);
	for (int i = 0; i < num_nets; i++) {
		glob[i] = clb_net[i].is_global ? 1 : 0;
		for (int k = 0; k <= clb_net[i].num_sinks; k++) term[g_net_ptr[i] + k] = net_rr_terminals[i][k];
		bb[4 * i + 0] = route_bb[i].xmin; bb[4 * i + 1] = route_bb[i].xmax;
		bb[4 * i + 2] = route_bb[i].ymin; bb[4 * i + 3] = route_bb[i].ymax;
	}
	p.num_nets = num_nets; p.num_terminals = g_net_ptr[num_nets];
	p.net_ptr = g_net_ptr.data(); p.net_terminals = term.data(); p.net_is_global = glob.data(); p.net_bb = bb.data();

	std::vector<int32_t> gsrc, gcnt;
	if (clb_opins_used_locally) {
		for (int b = 0; b < num_blocks; b++)
			for (int c = 0; c < block[b].type->num_class; c++)
				if (clb_opins_used_locally[b][c].nelem > 0) {
					gsrc.push_back(rr_blk_source[b][c]);
					gcnt.push_back(clb_opins_used_locally[b][c].nelem);
				}
	}
	p.num_opin_groups = (int)gsrc.size(); p.opin_group_source = gsrc.data(); p.opin_group_count = gcnt.data();

	p.opts.first_iter_pres_fac = ro.first_iter_pres_fac; p.opts.initial_pres_fac = ro.initial_pres_fac;
	p.opts.pres_fac_mult = ro.pres_fac_mult; p.opts.acc_fac = ro.acc_fac; p.opts.bend_cost = ro.bend_cost;
	p.opts.astar_fac = ro.astar_fac; p.opts.max_criticality = ro.max_criticality;
	p.opts.criticality_exp = ro.criticality_exp; p.opts.max_router_iterations = ro.max_router_iterations;
	p.opts.timing_analysis_enabled = timing_enabled ? 1 : 0; p.opts.bb_factor = ro.bb_factor;
	p.opts.router_algorithm = ro.router_algorithm == BREADTH_FIRST ? 1 : 0;

	char msg[256];
	if (pf_problem_check(&p, msg, sizeof(msg)) != 0) { fprintf(stderr, "PF_REF problem check failed: %s\n", msg); exit(2); }
	if (pf_problem_write(path, &p) != 0) { fprintf(stderr, "PF_REF cannot write %s\n", path); exit(2); }
	fprintf(stderr, "PF_REF wrote problem %s: N=%d E=%ld nets=%d terminals=%d switches=%d indexed=%d opin_groups=%d\n",
			path, N, E, num_nets, p.num_terminals, nsw, num_rr_indexed_data, p.num_opin_groups);
}

/* names, IO tiles, block locations and global-net pins: what print_route (route_common.c:1322) and print_place
 * (read_place.c:266) read beyond the flat problem — pf_names of include/pf_text.h, built by the reference-side
 * binding integration/vpr_text_adapter.cxx (linked into this harness so that it is tested where the reference runs) */
int pf_adapter_build_names(pf_names *n);
void pf_adapter_print_route(char *route_file);
static void export_names(const char *path) {
	pf_names n;
	int rc = pf_adapter_build_names(&n);
	if (rc == 0) { rc = pf_names_write(path, &n); pf_names_free(&n); }
	fprintf(stderr, "PF_REF wrote names %s: %d nets, %d blocks (rc %d)\n", path, num_nets, num_blocks, rc);
}

/* PF_DUMP_NETLIST=<file>: block[] and clb_net[] as read_netlist (base/read_netlist.c:74) left them — the golden for the
 * native .net reader (pf_net_read, include/pf_text.h); the hook runs after placement, so post_place_sync's shift of the pins of
 * blocks at z > 0 (base/place_and_route.c:888-920) is taken out again.  Text: "b <name> <type> <pins per instance> <net per pin...>" per
 * block, "n <name> <is_global> <terminals> <block>:<pin>..." per net (driver first). */
static void export_netlist(const char *path) {
	FILE *f = fopen(path, "w");
	if (!f) { fprintf(stderr, "PF_REF cannot write %s\n", path); exit(2); }
	fprintf(f, "PFNETLIST 1\nblocks %d\n", num_blocks);
	for (int i = 0; i < num_blocks; i++) {
		const int per = block[i].type->num_pins / block[i].type->capacity;
		fprintf(f, "b %s %s %d", block[i].name, block[i].type->name, per);
		for (int k = 0; k < per; k++) fprintf(f, " %d", block[i].nets[k + block[i].z * per]);      /* post_place_sync undone */
		fprintf(f, "\n");
	}
	fprintf(f, "nets %d\n", num_nets);
	for (int i = 0; i < num_nets; i++) {
		fprintf(f, "n %s %d %d", clb_net[i].name, clb_net[i].is_global ? 1 : 0, clb_net[i].num_sinks + 1);
		for (int k = 0; k <= clb_net[i].num_sinks; k++) {
			const int b = clb_net[i].node_block[k];
			fprintf(f, " %d:%d", b, clb_net[i].node_block_pin[k] - block[b].z * (block[b].type->num_pins / block[b].type->capacity));
		}
		fprintf(f, "\n");
	}
	fclose(f);
	fprintf(stderr, "PF_REF wrote netlist %s: %d blocks, %d nets\n", path, num_blocks, num_nets);
}

/* base/place_and_route.c is compiled with -Dprint_route=pf_hook_print_route: the reference's own writer runs as
 * always; with PF_ADAPTER_ROUTE_FILE=<file> the adapter's native writer writes the same routing next to it */
void pf_hook_print_route(char *route_file) {
	double t0 = now_s();
	print_route(route_file);
	double t1 = now_s();
	if (getenv("PF_ADAPTER_ROUTE_FILE")) {
		pf_adapter_print_route((char *)getenv("PF_ADAPTER_ROUTE_FILE"));
		fprintf(stderr, "PF_REF print_route %.4f s, pf_adapter_print_route %.4f s\n", t1 - t0, now_s() - t1);
	}
}

static int serial_num_of_routing() { /* same arithmetic as get_serial_num, route_common.c:224-254 */
	int serial_num = 0;
	for (int inet = 0; inet < num_nets; inet++)
		for (struct s_trace *t = trace_head[inet]; t; t = t->next) {
			int inode = t->index;
			serial_num += (inet + 1) * (rr_node[inode].xlow * (nx + 1) - rr_node[inode].yhigh);
			serial_num -= rr_node[inode].ptc_num * (inet + 1) * 10;
			serial_num -= rr_node[inode].type * (inet + 1) * 100;
			serial_num %= 2000000000;
		}
	return serial_num;
}

static void export_result(const char *path, boolean ok, float **net_delay) {
	pf_result r;
	memset(&r, 0, sizeof(r));
	build_net_ptr();
	std::vector<int32_t> tptr(num_nets + 1, 0), tnode;
	std::vector<int16_t> tsw;
	int wl = 0;
	for (int i = 0; i < num_nets; i++) {
		for (struct s_trace *t = trace_head[i]; t; t = t->next) { tnode.push_back(t->index); tsw.push_back(t->iswitch); }
		tptr[i + 1] = (int32_t)tnode.size();
		if (!clb_net[i].is_global && clb_net[i].num_sinks != 0 && trace_head[i]) {
			int bends, len, segs;
			get_num_bends_and_length(i, &bends, &len, &segs);
			wl += len;
		}
	}
	std::vector<float> nd(g_net_ptr[num_nets], 0.f);
	for (int i = 0; i < num_nets; i++)
		for (int k = 1; k <= clb_net[i].num_sinks; k++) nd[g_net_ptr[i] + k] = net_delay ? net_delay[i][k] : 0.f;
	std::vector<int32_t> occ(num_rr_nodes);
	for (int i = 0; i < num_rr_nodes; i++) occ[i] = rr_node[i].occ;
	r.success = ok ? 1 : 0; r.iterations = g_iter; r.serial_num = serial_num_of_routing(); r.total_wirelength = wl;
	r.num_nets = num_nets; r.trace_ptr = tptr.data(); r.trace_node = tnode.data(); r.trace_switch = tsw.data();
	r.num_terminals = g_net_ptr[num_nets]; r.net_delay = nd.data();
	r.num_nodes = num_rr_nodes; r.occ = occ.data();
	r.num_iter_stats = (int)g_stats.size(); r.iter_stats = g_stats.data();
	r.num_crit_iters = r.num_terminals ? (int)(g_crit.size() / r.num_terminals) : 0; r.iter_crit = g_crit.data();
	if (pf_result_write(path, &r) != 0) { fprintf(stderr, "PF_REF cannot write %s\n", path); exit(2); }
	fprintf(stderr, "PF_REF wrote result %s: success=%d iterations=%d cookie=%d wirelength=%d crit_iters=%d\n",
			path, r.success, r.iterations, r.serial_num, r.total_wirelength, r.num_crit_iters);
}

static void report_times(boolean ok, double total) {
	fprintf(stderr, "PF_REF route success=%d iterations=%d route_time_s=%.6f\n", ok ? 1 : 0, g_iter, total);
	double prev = 0;
	for (size_t i = 0; i < g_iter_time.size(); i++) {
		fprintf(stderr, "PF_REF iter %zu t_end_s=%.6f dt_s=%.6f overused=%d nets=%d\n", i + 1, g_iter_time[i],
				g_iter_time[i] - prev, g_stats[i].overused_nodes, g_stats[i].nets_routed);
		prev = g_iter_time[i];
	}
}

/* the reference's timing graph, flattened (include/pf_types.h: pf_timing_graph) */
static void export_timing_graph(const char *path) {
	pf_timing_graph g;
	memset(&g, 0, sizeof(g));
	g.num_tnodes = num_tnodes;
	std::vector<int32_t> eptr(num_tnodes + 1, 0), eto, cdom(num_tnodes), lptr, lnodes, drv(num_nets, -1);
	std::vector<float> etd, cdel(num_tnodes), cons;
	std::vector<uint8_t> ty(num_tnodes);
	for (int i = 0; i < num_tnodes; i++) {
		eptr[i + 1] = eptr[i] + tnode[i].num_edges;
		for (int k = 0; k < tnode[i].num_edges; k++) { eto.push_back(tnode[i].out_edges[k].to_node); etd.push_back(tnode[i].out_edges[k].Tdel); }
		ty[i] = (uint8_t)tnode[i].type; cdom[i] = tnode[i].clock_domain; cdel[i] = tnode[i].clock_delay;
		if (tnode[i].type == TN_CB_OPIN) {
			int iblk, inet;
			get_tnode_block_and_output_net(i, &iblk, &inet);
			if (inet >= 0 && inet < num_nets) drv[inet] = i;
		}
	}
	lptr.push_back(0);
	for (int lv = 0; lv < num_tnode_levels; lv++) {
		for (int k = 0; k < tnodes_at_level[lv].nelem; k++) lnodes.push_back(tnodes_at_level[lv].list[k]);
		lptr.push_back((int32_t)lnodes.size());
	}
	const int C = g_sdc ? g_sdc->num_constrained_clocks : 0;
	for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) cons.push_back(g_sdc->domain_constraint[i][j]);
	std::vector<int32_t> ovr_d, ovr_t; std::vector<float> ovr_c;
	/* clock-to-flipflop override constraints (g_sdc->cf_constraints), resolved to (source domain, sink tnode): what
	 * find_cf_constraint(clock name, find_tnode_net_name(inode)) answers at every sink (timing/path_delay.c:2753-2768, :3667-3684,
	 * :3749-3767) — the first matching entry wins, as there */
	if (g_sdc && g_sdc->num_cf_constraints > 0) {
		for (int i = 0; i < num_tnodes; i++) {
			if (tnode[i].num_edges != 0 || (tnode[i].type != TN_FF_SINK && tnode[i].type != TN_OUTPAD_SINK)) continue;
			const char *name = block[tnode[i].block].pb->rr_node_to_pb_mapping[tnode[i].pb_graph_pin->pin_count_in_cluster]->name;
			for (int c = 0; c < g_sdc->num_constrained_clocks; c++) {
				int found = -1;
				for (int icf = 0; icf < g_sdc->num_cf_constraints && found < 0; icf++) {
					bool src = false, snk = false;
					for (int a = 0; a < g_sdc->cf_constraints[icf].num_source; a++) if (strcmp(g_sdc->cf_constraints[icf].source_list[a], g_sdc->constrained_clocks[c].name) == 0) src = true;
					for (int a = 0; src && a < g_sdc->cf_constraints[icf].num_sink; a++) if (strcmp(g_sdc->cf_constraints[icf].sink_list[a], name) == 0) snk = true;
					if (src && snk) found = icf;
				}
				if (found >= 0) { ovr_d.push_back(c); ovr_t.push_back(i); ovr_c.push_back(g_sdc->cf_constraints[found].constraint); }
			}
		}
	}
	g.num_overrides = (int32_t)ovr_t.size(); g.override_domain = ovr_d.data(); g.override_tnode = ovr_t.data(); g.override_constraint = ovr_c.data();
	g.num_tedges = (int32_t)eto.size();
	g.edge_ptr = eptr.data(); g.edge_to = eto.data(); g.edge_Tdel = etd.data(); g.type = ty.data();
	g.clock_domain = cdom.data(); g.clock_delay = cdel.data();
	g.num_levels = num_tnode_levels; g.level_ptr = lptr.data(); g.level_nodes = lnodes.data();
	g.num_domains = C; g.constraint = cons.data();
	g.num_nets = num_nets; g.net_driver = drv.data();
	char msg[256];
	int rc = pf_timing_graph_check(&g, g_net_ptr.data(), msg, sizeof(msg));
	if (rc != 0) { fprintf(stderr, "PF_REF timing graph export is inconsistent: %s\n", msg); exit(2); }
	rc = pf_timing_graph_write(path, &g);
	fprintf(stderr, "PF_REF wrote timing graph %s: %d override constraints, %d tnodes, %d tedges, %d levels, %d clock domains (rc %d)\n", path, g.num_overrides, g.num_tnodes, g.num_tedges,
			g.num_levels, g.num_domains, rc);
}

static void export_sta_vectors(const char *path) {
	pf_sta_vectors v;
	memset(&v, 0, sizeof(v));
	v.num_terminals = g_net_ptr[num_nets];
	v.num_calls = (int32_t)g_sta_cpd.size();
	if ((size_t)v.num_calls * v.num_terminals != g_sta_crit.size() || g_sta_crit.size() != g_sta_delay.size()) {
		fprintf(stderr, "PF_REF STA capture is inconsistent (%zu delays, %zu crits, %d calls)\n", g_sta_delay.size(), g_sta_crit.size(), v.num_calls);
		exit(2);
	}
	v.net_delay = g_sta_delay.data(); v.crit = g_sta_crit.data(); v.cpd = g_sta_cpd.data();
	int rc = pf_sta_vectors_write(path, &v);
	fprintf(stderr, "PF_REF wrote STA vectors %s: %d calls x %d terminals (rc %d); do_timing_analysis took %.3f ms per call\n", path, v.num_calls,
			v.num_terminals, rc, v.num_calls ? 1e3 * g_sta_seconds / v.num_calls : 0.);
}

boolean pf_hook_try_timing_driven_route(struct s_router_opts router_opts, float **net_delay, t_slack *slacks,
		t_ivec **clb_opins_used_locally, boolean timing_analysis_enabled) {
	const char *dp = getenv("PF_DUMP_PROBLEM"), *dr = getenv("PF_DUMP_RESULT");
	if (dp) export_problem(dp, router_opts, timing_analysis_enabled, clb_opins_used_locally);
	if (getenv("PF_DUMP_NAMES") && !g_inject) export_names(getenv("PF_DUMP_NAMES"));
	if (getenv("PF_DUMP_NETLIST") && !g_inject) export_netlist(getenv("PF_DUMP_NETLIST"));
	build_net_ptr();
	g_iter = 0; g_stats.clear(); g_crit.clear(); g_iter_time.clear();
	g_sta_delay.clear(); g_sta_crit.clear(); g_sta_cpd.clear(); g_sta_seconds = 0;
	if (getenv("PF_DUMP_TGRAPH") && timing_analysis_enabled && !g_inject) export_timing_graph(getenv("PF_DUMP_TGRAPH"));
	/* criticalities of iteration 1 (route_timing.c:116-128) */
	{
		float v = timing_analysis_enabled ? 1.f : 0.f;
		g_crit.assign(g_net_ptr[num_nets], 0.f);
		for (int i = 0; i < num_nets; i++)
			if (!clb_net[i].is_global)
				for (int k = 1; k <= clb_net[i].num_sinks; k++) g_crit[g_net_ptr[i] + k] = v;
	}
	g_slacks = slacks;
	g_net_delay_arg = net_delay;
	g_t0 = now_s();
	boolean ok = try_timing_driven_route(router_opts, net_delay, slacks, clb_opins_used_locally, timing_analysis_enabled);
	double total = now_s() - g_t0;
	if (!timing_analysis_enabled) {
		/* record the all-zero criticalities of later iterations too, for symmetry */
		size_t T = g_net_ptr[num_nets];
		g_crit.resize(T * (size_t)(g_iter > 0 ? g_iter : 1), 0.f);
	}
	report_times(ok, total);
	if (dr) export_result(dr, ok, net_delay);
	if (getenv("PF_DUMP_STA") && timing_analysis_enabled && !g_inject) export_sta_vectors(getenv("PF_DUMP_STA"));
	return ok;
}

/* base/place_and_route.c is also compiled with -Dread_place=pf_hook_read_place: with PF_ADAPTER_READ_PLACE=1 the adapter's
 * hashed reader (integration/vpr_text_adapter.cxx) reads the file first, then the reference's read_place; the two
 * placements must agree */
void pf_adapter_read_place(const char *place_file, const char *net_file, const char *arch_file, int L_nx, int L_ny, int L_num_blocks, struct s_block block_list[]);
void pf_hook_read_place(const char *place_file, const char *a2, const char *a3, int L_nx, int L_ny, int L_num_blocks, struct s_block block_list[]) {
	std::vector<int> ax, ay, az;
	double t0 = now_s(), t1 = t0;
	const bool both = getenv("PF_ADAPTER_READ_PLACE") != NULL;
	if (both) {
		pf_adapter_read_place(place_file, a2, a3, L_nx, L_ny, L_num_blocks, block_list);
		t1 = now_s();
		for (int b = 0; b < L_num_blocks; b++) { ax.push_back(block_list[b].x); ay.push_back(block_list[b].y); az.push_back(block_list[b].z); block_list[b].x = block_list[b].y = block_list[b].z = -7; }
	}
	read_place(place_file, a2, a3, L_nx, L_ny, L_num_blocks, block_list);
	if (both) {
		int diff = 0;
		for (int b = 0; b < L_num_blocks; b++) diff += ax[b] != block_list[b].x || ay[b] != block_list[b].y || az[b] != block_list[b].z;
		fprintf(stderr, "PF_REF read_place: adapter %.4f s, reference %.4f s, %d blocks, %d differ\n", t1 - t0, now_s() - t1, L_num_blocks, diff);
		if (diff) exit(3);
	}
}

/* ------------------------------------------------------------------ flow mode */
static int run_flow(int argc, char **argv) {
	static t_options Options;
	static t_arch Arch;
	static t_vpr_setup vpr_setup;
	memset(&Options, 0, sizeof(Options));
	program_start = std::chrono::high_resolution_clock::now();
	vpr_init(argc, argv, &Options, &vpr_setup, &Arch);
	if (vpr_setup.PackerOpts.doPacking) vpr_pack(vpr_setup, Arch);
	if (vpr_setup.PlacerOpts.doPlacement || vpr_setup.RouterOpts.doRouting) {
		vpr_init_pre_place_and_route(vpr_setup, Arch);
		place_and_route(vpr_setup.Operation, vpr_setup.PlacerOpts, vpr_setup.FileNameOpts.PlaceFile,
				vpr_setup.FileNameOpts.NetFile, vpr_setup.FileNameOpts.ArchFile, vpr_setup.FileNameOpts.RouteFile,
				vpr_setup.AnnealSched, vpr_setup.RouterOpts, vpr_setup.RoutingArch, vpr_setup.Segments,
				vpr_setup.Timing, Arch.Chans, Arch.models, Arch.Directs, Arch.num_directs);
	}
	fflush(stdout);
	return 0;
}

/* ------------------------------------------------------------------ inject mode */
static int run_inject(int argc, char **argv) {
	const char *prob_path = argv[0];
	const char *result_path = NULL, *crit_path = NULL, *route_file = NULL;
	int max_iters = -1, limit_nets = -1;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--result") && i + 1 < argc) result_path = argv[++i];
		else if (!strcmp(argv[i], "--route-file") && i + 1 < argc) route_file = argv[++i];
		else if (!strcmp(argv[i], "--crit") && i + 1 < argc) crit_path = argv[++i];
		else if (!strcmp(argv[i], "--max_iters") && i + 1 < argc) max_iters = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--limit_nets") && i + 1 < argc) limit_nets = atoi(argv[++i]);
		else { fprintf(stderr, "unknown inject option %s\n", argv[i]); return 2; }
	}
	pf_problem p;
	double tl0 = now_s();
	if (pf_problem_read(prob_path, &p) != 0) { fprintf(stderr, "cannot read %s\n", prob_path); return 2; }
	if (crit_path) {
		if (pf_result_read(crit_path, &g_replay) != 0 || g_replay.num_crit_iters < 1
				|| g_replay.num_terminals != p.num_terminals) { fprintf(stderr, "bad --crit file %s\n", crit_path); return 2; }
		g_have_replay = true;
	}
	g_inject = true;

	nx = p.nx; ny = p.ny;
	num_rr_nodes = p.num_nodes;
	rr_node = new t_rr_node[p.num_nodes];
	for (int i = 0; i < p.num_nodes; i++) {
		t_rr_node &n = rr_node[i];
		n.xlow = p.xlow[i]; n.ylow = p.ylow[i]; n.xhigh = p.xhigh[i]; n.yhigh = p.yhigh[i];
		n.ptc_num = p.ptc_num[i]; n.cost_index = p.cost_index[i]; n.occ = 0; n.capacity = p.capacity[i];
		n.type = (t_rr_type)p.type[i]; n.direction = (enum e_direction)p.direction[i];
		n.R = p.R[i]; n.C = p.C[i];
		n.num_edges = (short)(p.row_ptr[i + 1] - p.row_ptr[i]);
		n.edges = p.edge_to + p.row_ptr[i];        /* rows alias the flat CSR arrays */
		n.switches = p.edge_sw + p.row_ptr[i];
	}
	switch_inf = (struct s_switch_inf *)calloc(p.num_switches, sizeof(struct s_switch_inf));
	for (int s = 0; s < p.num_switches; s++) {
		switch_inf[s].buffered = p.switches[s].buffered ? TRUE : FALSE; switch_inf[s].R = p.switches[s].R;
		switch_inf[s].Cin = p.switches[s].Cin; switch_inf[s].Cout = p.switches[s].Cout; switch_inf[s].Tdel = p.switches[s].Tdel;
	}
	num_rr_indexed_data = p.num_indexed;
	rr_indexed_data = (t_rr_indexed_data *)calloc(p.num_indexed, sizeof(t_rr_indexed_data));
	for (int i = 0; i < p.num_indexed; i++) {
		rr_indexed_data[i].base_cost = p.indexed[i].base_cost; rr_indexed_data[i].saved_base_cost = p.indexed[i].saved_base_cost;
		rr_indexed_data[i].ortho_cost_index = p.indexed[i].ortho_cost_index; rr_indexed_data[i].seg_index = p.indexed[i].seg_index;
		rr_indexed_data[i].inv_length = p.indexed[i].inv_length; rr_indexed_data[i].T_linear = p.indexed[i].T_linear;
		rr_indexed_data[i].T_quadratic = p.indexed[i].T_quadratic; rr_indexed_data[i].C_load = p.indexed[i].C_load;
	}
	num_nets = p.num_nets;
	clb_net = new struct s_net[p.num_nets];
	net_rr_terminals = (int **)malloc(sizeof(int *) * p.num_nets);
	/* nets are called n<i> and the fabric gets VPR's IO ring (what pf_names_synthetic of include/pf_text.h assumes),
	 * so that the reference's own print_route can write the routing of a generated problem (--route-file) */
	std::vector<char> name_chars((size_t)p.num_nets * 12 + 16);
	static struct s_type_descriptor fake_io_type, fake_clb_type;
	IO_TYPE = &fake_io_type;
	grid = (struct s_grid_tile **)malloc(sizeof(struct s_grid_tile *) * (nx + 2));
	for (int x = 0; x <= nx + 1; x++) {
		grid[x] = (struct s_grid_tile *)calloc(ny + 2, sizeof(struct s_grid_tile));
		for (int y = 0; y <= ny + 1; y++) grid[x][y].type = (x == 0 || y == 0 || x == nx + 1 || y == ny + 1) ? &fake_io_type : &fake_clb_type;
	}
	for (int i = 0; i < p.num_nets; i++) {
		clb_net[i].name = &name_chars[(size_t)i * 12];
		sprintf(clb_net[i].name, "n%d", i);
		clb_net[i].num_sinks = p.net_ptr[i + 1] - p.net_ptr[i] - 1;
		clb_net[i].node_block = NULL; clb_net[i].node_block_port = NULL; clb_net[i].node_block_pin = NULL;
		clb_net[i].is_global = p.net_is_global[i] ? TRUE : FALSE;
		if (limit_nets >= 0 && i >= limit_nets) clb_net[i].is_global = TRUE; /* bounded sample: skip the rest */
		clb_net[i].is_const_gen = FALSE;
		net_rr_terminals[i] = p.net_terminals + p.net_ptr[i];
	}
	alloc_route_static_structs();              /* trace_head/tail, heap, route_bb (route_common.c:850) */
	for (int i = 0; i < p.num_nets; i++) {
		route_bb[i].xmin = p.net_bb[4 * i + 0]; route_bb[i].xmax = p.net_bb[4 * i + 1];
		route_bb[i].ymin = p.net_bb[4 * i + 2]; route_bb[i].ymax = p.net_bb[4 * i + 3];
		trace_tail[i] = NULL;
	}
	alloc_and_load_rr_node_route_structs();    /* route_common.c:1012 */

	/* locally used OPINs: one fabricated single-class block per group */
	t_ivec **opins = NULL;
	static struct s_type_descriptor fake_type;
	num_blocks = p.num_opin_groups;
	if (num_blocks > 0) {
		memset(&fake_type, 0, sizeof(fake_type));
		fake_type.num_class = 1;
		block = (struct s_block *)calloc(num_blocks, sizeof(struct s_block));
		rr_blk_source = (int **)malloc(sizeof(int *) * num_blocks);
		opins = (t_ivec **)malloc(sizeof(t_ivec *) * num_blocks);
		for (int b = 0; b < num_blocks; b++) {
			block[b].type = &fake_type;
			rr_blk_source[b] = (int *)malloc(sizeof(int));
			rr_blk_source[b][0] = p.opin_group_source[b];
			opins[b] = (t_ivec *)malloc(sizeof(t_ivec));
			opins[b][0].nelem = p.opin_group_count[b];
			opins[b][0].list = (int *)calloc(p.opin_group_count[b], sizeof(int));
		}
	}

	t_slack slacks;
	memset(&slacks, 0, sizeof(slacks));
	slacks.slack = (float **)malloc(sizeof(float *) * p.num_nets);
	slacks.timing_criticality = (float **)malloc(sizeof(float *) * p.num_nets);
	float **net_delay = (float **)malloc(sizeof(float *) * p.num_nets);
	for (int i = 0; i < p.num_nets; i++) {
		int n = clb_net[i].num_sinks + 1;
		slacks.slack[i] = (float *)calloc(n, sizeof(float));
		slacks.timing_criticality[i] = (float *)calloc(n, sizeof(float));
		net_delay[i] = (float *)calloc(n, sizeof(float));
	}

	struct s_router_opts ro;
	memset(&ro, 0, sizeof(ro));
	ro.first_iter_pres_fac = p.opts.first_iter_pres_fac; ro.initial_pres_fac = p.opts.initial_pres_fac;
	ro.pres_fac_mult = p.opts.pres_fac_mult; ro.acc_fac = p.opts.acc_fac; ro.bend_cost = p.opts.bend_cost;
	ro.astar_fac = p.opts.astar_fac; ro.max_criticality = p.opts.max_criticality;
	ro.criticality_exp = p.opts.criticality_exp;
	ro.max_router_iterations = max_iters > 0 ? max_iters : p.opts.max_router_iterations;
	ro.bb_factor = p.opts.bb_factor;
	fprintf(stderr, "PF_REF inject: loaded N=%d E=%d nets=%d in %.3f s\n", p.num_nodes, p.num_edges, p.num_nets, now_s() - tl0);

	if (result_path) setenv("PF_DUMP_RESULT", result_path, 1);
	unsetenv("PF_DUMP_PROBLEM");
	boolean timing = g_have_replay ? TRUE : FALSE;
	if (g_have_replay) {
		/* iteration-1 criticalities are set inside the router (1.0); the replay supplies the rest,
		 * and the hook's index 0 row is skipped */
	}
	boolean ok;
	if (p.opts.router_algorithm == 1) { ro.router_algorithm = BREADTH_FIRST; ok = pf_hook_try_breadth_first_route(ro, opins, 0); }
	else ok = pf_hook_try_timing_driven_route(ro, net_delay, &slacks, opins, timing);
	if (route_file) {
		double t0 = now_s();
		print_route((char *)route_file);                 /* the reference's own writer, route_common.c:1322 */
		fprintf(stderr, "PF_REF print_route %s: %.3f s\n", route_file, now_s() - t0);
	}
	return ok ? 0 : 1;
}

/* route_common.c:495 dispatches `--router_algorithm breadth_first` here (try_route), bound by
 * -Dtry_breadth_first_route=pf_hook_try_breadth_first_route; route_breadth_first.c itself is compiled with
 * -Dfeasible_routing=pf_hook_feasible_routing so the per-iteration overuse is recorded */
boolean pf_hook_try_breadth_first_route(struct s_router_opts router_opts, t_ivec **clb_opins_used_locally, int width_fac) {
	const char *dp = getenv("PF_DUMP_PROBLEM"), *dr = getenv("PF_DUMP_RESULT");
	router_opts.router_algorithm = BREADTH_FIRST;
	if (dp) export_problem(dp, router_opts, FALSE, clb_opins_used_locally);
	build_net_ptr();
	g_iter = 0; g_stats.clear(); g_crit.clear(); g_iter_time.clear();
	g_t0 = now_s();
	boolean ok = try_breadth_first_route(router_opts, clb_opins_used_locally, width_fac);
	double total = now_s() - g_t0;
	g_crit.assign((size_t)g_net_ptr[num_nets] * (size_t)(g_iter > 0 ? g_iter : 1), 0.f);
	report_times(ok, total);
	if (dr) export_result(dr, ok, NULL);
	return ok;
}

int main(int argc, char **argv) {
	if (argc >= 3 && !strcmp(argv[1], "flow")) {
		argv[1] = argv[0];
		return run_flow(argc - 1, argv + 1);
	}
	if (argc >= 3 && !strcmp(argv[1], "inject")) return run_inject(argc - 2, argv + 2);
	fprintf(stderr, "usage: vpr_ref flow <arch.xml> <circuit> [vpr options]\n"
			"       vpr_ref inject <problem.pfp> [--result out.pfr] [--route-file out.route] [--crit golden.pfr] [--max_iters K] [--limit_nets M]\n");
	return 2;
}
