/*
 * pf_file.h — on-disk container for pf_problem / pf_result (little-endian, raw arrays).
 *
 * VPR 7 has no rr-graph file format (the graph is always regenerated, reference
 * vpr/SRC/route/rr_graph.c:385; its text echo dump_rr_graph rr_graph.c:2004 is write-only), so
 * the flat problem needs its own container: it is what the reference-side exporter writes
 * (INTEGRATION.md), what tests/golden/ holds, and what the stand-alone router CLI reads.
 *
 * Problem file:  "PFPROB01" | int32 header[16] | pf_router_opts | arrays in struct order
 * Result  file:  "PFRSLT01" | int32 header[16] | arrays in struct order
 * All functions return 0 on success, a negative PF_E* code otherwise; nothing calls exit().
 */
#ifndef PF_FILE_H
#define PF_FILE_H

#include "pf_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PF_OK 0
#define PF_EIO (-1)        /* cannot open / short read / short write */
#define PF_EFORMAT (-2)    /* bad magic or inconsistent header */
#define PF_ENOMEM (-3)
#define PF_EINVAL (-4)     /* argument check failed */
#define PF_ECUDA (-5)      /* CUDA runtime error (see pf_last_error) */
#define PF_EOVERFLOW (-6)  /* device scratch capacity exceeded even after growth */
#define PF_EUNROUTABLE (-7)/* a net has no path at all (disconnected rr graph), route_timing.c:482 */

int pf_problem_write(const char *path, const pf_problem *p);
/* Allocates every array with malloc; release with pf_problem_free. */
int pf_problem_read(const char *path, pf_problem *p);
void pf_problem_free(pf_problem *p);
/* Structural validation (ranges of every index, CSR monotonicity, terminal types). */
int pf_problem_check(const pf_problem *p, char *msg, int msg_len);

int pf_result_write(const char *path, const pf_result *r);
int pf_result_read(const char *path, pf_result *r);
void pf_result_free(pf_result *r);

/* Timing graph:  "PFTIMG01" | int32 header[16] | arrays in struct order
 * STA vectors:   "PFSTAV01" | int32 header[16] | net_delay | crit | cpd */
int pf_timing_graph_write(const char *path, const pf_timing_graph *g);
int pf_timing_graph_read(const char *path, pf_timing_graph *g);
void pf_timing_graph_free(pf_timing_graph *g);
/* ranges of every index, level lists a permutation with edges going to later levels, driver fan-outs equal to
 * the nets' sink counts (net_ptr = pf_problem.net_ptr, may be NULL to skip that part) */
int pf_timing_graph_check(const pf_timing_graph *g, const int32_t *net_ptr, char *msg, int msg_len);
int pf_sta_vectors_write(const char *path, const pf_sta_vectors *v);
int pf_sta_vectors_read(const char *path, pf_sta_vectors *v);
void pf_sta_vectors_free(pf_sta_vectors *v);

#ifdef __cplusplus
}
#endif
#endif
