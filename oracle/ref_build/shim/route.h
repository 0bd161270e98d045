/* Build shim: shadows parallel_route/route.h so that no TBB/Boost header is pulled in. */
#ifndef PF_SHIM_ROUTE_H
#define PF_SHIM_ROUTE_H
#include <vector>
struct net_t;
#endif
