/* Test infrastructure: link stubs for the two parallel routers that route_common.c names
 * (reference vpr/SRC/route/route_common.c:402,408) but whose sources need TBB/Boost/MPI. */
#include "vpr_types.h"
bool mpi_route_load_balanced_nonblocking_send_recv_encoded(t_router_opts *, struct s_det_routing_arch,
		t_direct_inf *, int, t_segment_inf *, t_timing_inf) { return false; }
bool partitioning_multi_sink_delta_stepping_route(const t_router_opts *) { return false; }
