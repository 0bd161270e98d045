/* Build shim: intentionally empty (shadows parallel_route/advanced_parallel_route_timing.h). */
