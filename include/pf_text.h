/*
 * pf_text.h — VPR's text files either side of the --route path, without VPR.
 *
 *   entry point        replaces (reference file:line)                              notes
 *   ------------------------------------------------------------------------------------------------------
 *   pf_route_write     print_route        vpr/SRC/route/route_common.c:1322-1417   byte-identical .route file
 *   pf_route_read      (none: VPR 7 cannot read a .route back)                     .route -> pf_result traces,
 *                                                                                  so check_route / wirelength can
 *                                                                                  run on any router's output
 *   pf_place_write     print_place        vpr/SRC/base/read_place.c:266-293        byte-identical .place file
 *   pf_place_read      read_place         vpr/SRC/base/read_place.c:15-139         same checks, hashed block
 *                                                                                  lookup: O(B) instead of the
 *                                                                                  reference's O(B^2) strcmp scan
 *
 * The flat pf_problem (pf_types.h) holds no strings.  What the text files need beyond it — net and
 * block names, which tiles are IO pads, the pins of the (never routed) global nets — travels in
 * pf_names, exported by the reference-side adapter next to the problem or made up
 * by the generator for synthetic fabrics (pf_names_synthetic).
 *
 *   pf_net_read        read_netlist       vpr/SRC/base/read_netlist.c:74-244,      the packed netlist as place and
 *                      load_external_nets_and_cb                     :836-984    route see it — block[] and clb_net[] —
 *                                                                                  without the architecture: the file's
 *                                                                                  own port lists give the pin numbering
 *
 * Plain C host code, no CUDA; all functions return PF_OK or a negative PF_E* code (pf_file.h).
 */
#ifndef PF_TEXT_H
#define PF_TEXT_H

#include "pf_types.h"
#include "pf_file.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_names {
	int32_t nx, ny;
	/* clb_net[i].name (vpr_types.h:521): net i is net_name_chars[net_name_ptr[i] .. net_name_ptr[i+1]) (no NUL) */
	int32_t num_nets;
	int32_t *net_name_ptr;     /* [num_nets+1] */
	char *net_name_chars;
	/* grid[x][y].type == IO_TYPE (globals.c:60), index x * (ny + 2) + y: print_route writes "Pad:" there */
	uint8_t *tile_is_io;       /* [(nx+2) * (ny+2)] */
	/* block[] (vpr_types.h:451): name and location */
	int32_t num_blocks;
	int32_t *block_name_ptr;   /* [num_blocks+1] */
	char *block_name_chars;
	int32_t *block_x, *block_y, *block_z;   /* [num_blocks] */
	/* pins of global nets, clb_net[i].node_block[] and block[].type->pin_class[node_block_pin]
	 * (route_common.c:1399-1412); rows of routed nets are empty */
	int32_t *gpin_ptr;         /* [num_nets+1] */
	int32_t *gpin_block;       /* [gpin_ptr[num_nets]] */
	int32_t *gpin_class;       /* [gpin_ptr[num_nets]] */
} pf_names;

/* container "PFNAME01" | int32 header[16] | arrays in struct order */
int pf_names_write(const char *path, const pf_names *n);
int pf_names_read(const char *path, pf_names *n);
void pf_names_free(pf_names *n);
/* sizes against the problem, monotone offsets, block coordinates inside the grid, no white space in names */
int pf_names_check(const pf_names *n, const pf_problem *p, char *msg, int msg_len);
/* names for a generated fabric (pf_gen.h): nets "n<i>", no blocks, the IO ring of a VPR grid */
int pf_names_synthetic(const pf_problem *p, pf_names *n);

/* print_route.  A routed net with no trace (num_sinks == 0) gets the reference's "Used in local cluster only" text. */
int pf_route_write(const char *path, const pf_problem *p, const pf_names *n, const pf_result *r);
/* Parses a .route file written by print_route / pf_route_write.  Fills r->num_nets, trace_ptr, trace_node,
 * trace_switch (the switch of the first rr edge node[k] -> node[k+1]; PF_OPEN at SINKs), total_wirelength and
 * serial_num; every other field is zero.  Each "Node:" line is verified against the problem (type, coordinates,
 * ptc): a file that belongs to another rr graph is PF_EFORMAT with the line number in pf_text_error(). */
int pf_route_read(const char *path, const pf_problem *p, pf_result *r);

int pf_place_write(const char *path, const char *net_file, const char *arch_file, const pf_names *n);
/* Sets block_x/y/z of the blocks named in the file.  net_file / arch_file may be NULL to skip the reference's
 * file-name comparison (read_place.c:55-64); a block missing from n, a grid of another size or a malformed line is
 * PF_EFORMAT.  *placed = number of blocks the file positioned. */
int pf_place_read(const char *path, const char *net_file, const char *arch_file, pf_names *n, int *placed);

/* The packed netlist (.net) as read_netlist leaves it for place and route: block[] and clb_net[] (vpr_types.h:451, :521).
 * The reference gets there by instantiating the architecture's pb_type hierarchy; this reader needs only the file: the
 * <inputs> / <outputs> / <clocks> sections of a complex block list every port in pb_type order ("open" = unused pin), which is
 * the block's pin numbering (inputs, outputs, clocks: the order load_external_nets_and_cb asserts, read_netlist.c:850), and
 * the net on an output pin is found by following "child[i].port[b]->interconnect" down the nested blocks to a primitive.
 * Nets are numbered in the order of their first appearance (add_net_to_hash, :594), the driver is terminal 0 and the sinks
 * follow in block / pin order (:934-965).  is_global: the reference takes it from the architecture (type->is_global_pin);
 * here pins listed under <clocks> are global — inputs an architecture declares is_non_clock_global are not recognisable
 * from the file.  What is NOT rebuilt: the pb tree inside the clusters (t_pb, local nets, the intra-cluster rr graph) — the
 * packer's and the timing graph's business, not the router's. */
typedef struct pf_netlist {
	int32_t num_blocks;
	int32_t *block_name_ptr;   /* [num_blocks+1]; block[i].name = block_name_chars[ptr[i] .. ptr[i+1]) (no NUL) */
	char *block_name_chars;
	int32_t *block_type_ptr;   /* [num_blocks+1]; block[i].type->name ("clb", "io", "mult", ...) */
	char *block_type_chars;
	int32_t *block_pin_ptr;    /* [num_blocks+1]; pins of ONE instance of the type (type->num_pins / type->capacity) */
	int32_t *block_pin_net;    /* block[i].nets[pin]: net index or PF_OPEN */
	uint8_t *block_pin_kind;   /* 0 input (RECEIVER), 1 output (DRIVER), 2 clock (RECEIVER, global) */
	int32_t num_nets;
	int32_t *net_name_ptr;     /* [num_nets+1] */
	char *net_name_chars;
	int32_t *net_ptr;          /* [num_nets+1]; terminals of net i: net_ptr[i] (the driver) .. net_ptr[i+1] */
	int32_t *net_block;        /* clb_net[i].node_block[k] */
	int32_t *net_block_pin;    /* clb_net[i].node_block_pin[k] */
	uint8_t *net_is_global;    /* clb_net[i].is_global */
} pf_netlist;
/* PF_EFORMAT (text in pf_text_error(), with the reference's wording where read_netlist.c has one): malformed XML, a top-level
 * element that is not FPGA_packed_netlist[0], an output that cannot be followed to a primitive, a net with two drivers or
 * none, a net on both clock and non-clock pins. */
int pf_net_read(const char *path, pf_netlist *out);
void pf_netlist_free(pf_netlist *nl);

/* description of the last PF_EFORMAT / PF_EINVAL of this thread's pf_route_read / pf_place_read / pf_net_read */
const char *pf_text_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PF_TEXT_H */
