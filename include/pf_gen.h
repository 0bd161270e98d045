/*
 * pf_gen.h — native generator of synthetic k6_N10-style routing problems (C-ABI, host only).
 * Stands where the reference's build_rr_graph (vpr/SRC/route/rr_graph.c:385) + a random
 * clustered netlist would stand for BASELINE.json configs[4] ("synthetic 400×400 CLB k6_N10 grid,
 * 200k random 4-pin nets"); see parallel_eda_b200/csrc/pf_gen.cpp for the construction.
 */
#ifndef PF_GEN_H
#define PF_GEN_H

#include "pf_file.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_gen_params {
	int32_t nx, ny;          /* CLB grid (an IO ring is added around it) */
	int32_t W;               /* tracks per channel (even: unidirectional pairs) */
	int32_t L;               /* segment length */
	int32_t num_nets;
	int32_t sinks_per_net;
	int32_t window;          /* sinks are drawn within +-window tiles of the source */
	uint32_t seed;           /* MT19937 */
	float fc_in, fc_out;     /* connection-box flexibilities as fractions of W */
	int32_t io_capacity;
	int32_t bb_factor;
	int32_t reserved[4];
} pf_gen_params;

/* 400 x 400, W = 100, L = 4, 200000 nets x 3 sinks, window 16, seed 20260921 */
void pf_gen_params_default(pf_gen_params *g);
/* Allocates every array of *out with malloc (release with pf_problem_free). */
int pf_gen_grid_problem(const pf_gen_params *g, pf_problem *out);
/* The same problem WITHOUT the rr graph arrays (node SoA, CSR stay NULL; num_nodes is set, num_edges is 0): nets, bounding
 * boxes, switch / cost-index tables and router options — what pf_router_create_generated (pf_router.h) needs from the host
 * when the graph itself is built on the device. */
int pf_gen_grid_nets(const pf_gen_params *g, pf_problem *out);

#ifdef __cplusplus
}
#endif
#endif
