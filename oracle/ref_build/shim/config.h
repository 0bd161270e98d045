/* Build shim: intentionally empty (shadows parallel_route/config.h). */
