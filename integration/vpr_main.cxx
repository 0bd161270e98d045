/* Minimal VPR driver for the drop-in demonstration binary oracle/_ref/vpr_b200: the stock flow
 * (vpr_init → [vpr_pack] → vpr_init_pre_place_and_route → place_and_route, reference base/vpr_api.c:164-454,
 * base/place_and_route.c:250) with the router call site bound to the B200 adapter.  Symbols the reference's
 * main.c normally provides are defined here (main.c:60-62,253,287). */
#include <chrono>
#include <string.h>
#include "vpr_types.h"
#include "vpr_api.h"
#include "globals.h"
#include "place_and_route.h"

std::chrono::time_point<std::chrono::high_resolution_clock> program_start;
char *s_circuit_name = nullptr;
void print_context(int, int) {}
void get_mem_usage(unsigned long &vm, unsigned long &rss) { vm = 0; rss = 0; }

int main(int argc, char **argv) {
	static t_options Options;
	static t_arch Arch;
	static t_vpr_setup vpr_setup;
	memset(&Options, 0, sizeof(Options));
	vpr_init(argc, argv, &Options, &vpr_setup, &Arch);
	if (vpr_setup.PackerOpts.doPacking) vpr_pack(vpr_setup, Arch);
	if (vpr_setup.PlacerOpts.doPlacement || vpr_setup.RouterOpts.doRouting) {
		vpr_init_pre_place_and_route(vpr_setup, Arch);
		place_and_route(vpr_setup.Operation, vpr_setup.PlacerOpts, vpr_setup.FileNameOpts.PlaceFile,
				vpr_setup.FileNameOpts.NetFile, vpr_setup.FileNameOpts.ArchFile, vpr_setup.FileNameOpts.RouteFile,
				vpr_setup.AnnealSched, vpr_setup.RouterOpts, vpr_setup.RoutingArch, vpr_setup.Segments,
				vpr_setup.Timing, Arch.Chans, Arch.models, Arch.Directs, Arch.num_directs);
	}
	return 0;
}
