/*
 * vpr_text_adapter.cxx — reference-side binding of the native text writers (include/pf_text.h).
 *
 * Second file a maintainer of chinhau5/parallel_eda adds to the VPR build: it reads VPR's
 * globals (clb_net[], block[], grid[][], rr_node[], trace_head[]; base/globals.c:48-97), fills the plain-array
 * pf_names / pf_problem / pf_result views and calls pf_route_write.  Glue only; needs pf_text.c (part of
 * libpf_router.so, no CUDA call on this path).
 *
 *   base/place_and_route.c is compiled with  -Dprint_route=pf_adapter_print_route -Dread_place=pf_adapter_read_place
 *   so its three print_route call sites (place_and_route.c:182,364,729) land here; the reference's print_route
 *   (route/route_common.c:1322-1417) stays in the build unchanged.  The file written is byte-identical
 *   (tests/test_text_formats.py::test_adapter_print_route_and_read_place_equal_the_reference), ten times sooner on a 200 k-net
 *   routing (0.33 s instead of 3.2 s for the 224 MB file).  The read_place call sites
 *   (place_and_route.c:85,284) land in pf_adapter_read_place: same checks and messages, but one hash lookup per line where
 *   read_place.c:108-114 runs strcmp down the whole block list ("O(B^2); 200 k-net fixtures ... expect
 *   minutes in the reader").
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "vpr_types.h"
#include "globals.h"

#include "pf_text.h"

/* The names side of the boundary: what print_route / print_place read beyond the flat problem.  Arrays are
 * malloc'ed; release with pf_names_free. */
int pf_adapter_build_names(pf_names *n) {
	memset(n, 0, sizeof(*n));
	n->nx = nx; n->ny = ny; n->num_nets = num_nets; n->num_blocks = num_blocks;
	size_t nchars = 0, bchars = 0, gpins = 0;
	for (int i = 0; i < num_nets; i++) { nchars += strlen(clb_net[i].name); if (clb_net[i].is_global) gpins += clb_net[i].num_sinks + 1; }
	for (int b = 0; b < num_blocks; b++) bchars += strlen(block[b].name);
	n->net_name_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)num_nets + 1));
	n->net_name_chars = (char *)malloc(nchars + 1);
	n->tile_is_io = (uint8_t *)malloc((size_t)(nx + 2) * (ny + 2));
	n->block_name_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)num_blocks + 1));
	n->block_name_chars = (char *)malloc(bchars + 1);
	n->block_x = (int32_t *)malloc(sizeof(int32_t) * ((size_t)num_blocks + 1));
	n->block_y = (int32_t *)malloc(sizeof(int32_t) * ((size_t)num_blocks + 1));
	n->block_z = (int32_t *)malloc(sizeof(int32_t) * ((size_t)num_blocks + 1));
	n->gpin_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)num_nets + 1));
	n->gpin_block = (int32_t *)malloc(sizeof(int32_t) * (gpins + 1));
	n->gpin_class = (int32_t *)malloc(sizeof(int32_t) * (gpins + 1));
	if (!n->net_name_ptr || !n->net_name_chars || !n->tile_is_io || !n->block_name_ptr || !n->block_name_chars || !n->block_x
			|| !n->block_y || !n->block_z || !n->gpin_ptr || !n->gpin_block || !n->gpin_class) { pf_names_free(n); return PF_ENOMEM; }
	size_t at = 0, g = 0;
	for (int i = 0; i < num_nets; i++) {
		size_t l = strlen(clb_net[i].name);
		n->net_name_ptr[i] = (int32_t)at; memcpy(n->net_name_chars + at, clb_net[i].name, l); at += l;
		n->gpin_ptr[i] = (int32_t)g;
		if (!clb_net[i].is_global) continue;
		for (int k = 0; k <= clb_net[i].num_sinks; k++) {            /* route_common.c:1399-1405 */
			int b = clb_net[i].node_block[k];
			n->gpin_block[g] = b;
			n->gpin_class[g] = block[b].type->pin_class[clb_net[i].node_block_pin[k]];
			g++;
		}
	}
	n->net_name_ptr[num_nets] = (int32_t)at; n->gpin_ptr[num_nets] = (int32_t)g;
	at = 0;
	for (int b = 0; b < num_blocks; b++) {
		size_t l = strlen(block[b].name);
		n->block_name_ptr[b] = (int32_t)at; memcpy(n->block_name_chars + at, block[b].name, l); at += l;
		n->block_x[b] = block[b].x; n->block_y[b] = block[b].y; n->block_z[b] = block[b].z;
	}
	n->block_name_ptr[num_blocks] = (int32_t)at;
	for (int x = 0; x <= nx + 1; x++)
		for (int y = 0; y <= ny + 1; y++) n->tile_is_io[(size_t)x * (ny + 2) + y] = grid[x][y].type == IO_TYPE ? 1 : 0;
	return PF_OK;
}

/* place_and_route.c:182,364,729 land here (-Dprint_route=pf_adapter_print_route) */
void pf_adapter_print_route(char *route_file) {
	/* the node and net fields print_route reads (route_common.c:1344-1386), and the s_trace lists as arrays */
	const int N = num_rr_nodes;
	std::vector<int16_t> xl(N), yl(N), xh(N), yh(N), ptc(N);
	std::vector<uint8_t> ty(N), glob(num_nets > 0 ? num_nets : 1);
	for (int i = 0; i < N; i++) {
		const t_rr_node &v = rr_node[i];
		xl[i] = v.xlow; yl[i] = v.ylow; xh[i] = v.xhigh; yh[i] = v.yhigh; ptc[i] = v.ptc_num; ty[i] = (uint8_t)v.type;
	}
	std::vector<int32_t> net_ptr(num_nets + 1, 0), tptr(num_nets + 1, 0), tnode;
	for (int i = 0; i < num_nets; i++) {
		net_ptr[i + 1] = net_ptr[i] + clb_net[i].num_sinks + 1;
		glob[i] = clb_net[i].is_global ? 1 : 0;
		if (!clb_net[i].is_global && clb_net[i].num_sinks != 0)
			for (struct s_trace *t = trace_head[i]; t; t = t->next) tnode.push_back(t->index);
		tptr[i + 1] = (int32_t)tnode.size();
	}
	if (tnode.empty()) tnode.push_back(0);
	pf_problem p;
	memset(&p, 0, sizeof(p));
	p.nx = nx; p.ny = ny; p.num_nodes = N; p.num_nets = num_nets; p.num_terminals = net_ptr[num_nets];
	p.xlow = xl.data(); p.ylow = yl.data(); p.xhigh = xh.data(); p.yhigh = yh.data(); p.ptc_num = ptc.data(); p.type = ty.data();
	p.net_ptr = net_ptr.data(); p.net_is_global = glob.data();
	pf_result r;
	memset(&r, 0, sizeof(r));
	r.num_nets = num_nets; r.trace_ptr = tptr.data(); r.trace_node = tnode.data();
	pf_names names;
	int rc = pf_adapter_build_names(&names);
	if (rc == PF_OK) {
		rc = pf_route_write(route_file, &p, &names, &r);
		pf_names_free(&names);
	}
	if (rc != PF_OK) {                                   /* reference style: message + exit (route_common.c:1376-1379) */
		vpr_printf(TIO_MESSAGE_ERROR, "in print_route: %s (%d)\n", pf_text_error(), rc);
		exit(1);
	}
}

/* place_and_route.c:85,284 land here (-Dread_place=pf_adapter_read_place).  Parameter order as in the reference's
 * definition (read_place.c:15): the second argument is compared with the netlist name in the file, the third with the
 * architecture name — the reference's parameter NAMES say the opposite, its callers pass (place, net, arch). */
void pf_adapter_read_place(const char *place_file, const char *net_file, const char *arch_file, int L_nx, int L_ny,
		int L_num_blocks, struct s_block block_list[]) {
	pf_names n;
	memset(&n, 0, sizeof(n));
	n.nx = L_nx; n.ny = L_ny; n.num_blocks = L_num_blocks;
	size_t chars = 0;
	for (int b = 0; b < L_num_blocks; b++) chars += strlen(block_list[b].name);
	std::vector<int32_t> ptr(L_num_blocks + 1), bx(L_num_blocks + 1), by(L_num_blocks + 1), bz(L_num_blocks + 1);
	std::vector<char> names(chars + 1);
	size_t at = 0;
	for (int b = 0; b < L_num_blocks; b++) {
		size_t l = strlen(block_list[b].name);
		ptr[b] = (int32_t)at; memcpy(&names[at], block_list[b].name, l); at += l;
		bx[b] = block_list[b].x; by[b] = block_list[b].y; bz[b] = block_list[b].z;
	}
	ptr[L_num_blocks] = (int32_t)at;
	n.block_name_ptr = ptr.data(); n.block_name_chars = names.data();
	n.block_x = bx.data(); n.block_y = by.data(); n.block_z = bz.data();
	int placed = 0;
	int rc = pf_place_read(place_file, net_file, arch_file, &n, &placed);
	if (rc != PF_OK) {                                   /* reference style: message + exit (read_place.c:49-120) */
		vpr_printf(TIO_MESSAGE_ERROR, "%s\n", rc == PF_EFORMAT ? pf_text_error() : "read_place: cannot read the placement file");
		exit(1);
	}
	for (int b = 0; b < L_num_blocks; b++) { block_list[b].x = bx[b]; block_list[b].y = by[b]; block_list[b].z = bz[b]; }
}
