/* Build shim: intentionally empty (shadows parallel_route/parallel_route_timing.h). */
