/* TEST INFRASTRUCTURE: fiber scheduler of the lock-step warp emulator (see pf_emu.h). */
#include "pf_emu.h"

#include <vector>

/* Minimal x86-64 System V context switch (callee-saved registers + stack pointer): ucontext's
 * swapcontext costs two sigprocmask system calls per switch, which dominates an emulated warp. */
extern "C" void pf_emu_switch(void **save_sp, void *load_sp);
asm(R"(
	.text
	.globl pf_emu_switch
	.type pf_emu_switch,@function
pf_emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
	.size pf_emu_switch,.-pf_emu_switch
)");

pf_emu_lane *pf_emu_cur = NULL;

namespace {
struct Fiber {
	void *sp;
	pf_emu_lane lane;
	char *stack;
	bool finished;
};
void *g_sched_sp;
pf_emu_warp_fn g_fn;
void *g_arg;
Fiber *g_running;

void trampoline(void) {
	Fiber *f = g_running;
	g_fn(g_arg, f->lane.warp->warp_id);
	f->finished = true;
	f->lane.warp->done[f->lane.lane] = 1;
	pf_emu_switch(&f->sp, g_sched_sp);
	abort();   /* a finished fiber is never resumed */
}
}  // namespace

void pf_emu_yield(void) {
	Fiber *f = g_running;
	pf_emu_switch(&f->sp, g_sched_sp);
}

void pf_emu_launch(pf_emu_warp_fn fn, void *arg, int nwarps) {
	const size_t STACK = 256 * 1024;
	const char *rev = getenv("PF_EMU_REVERSE");
	bool reverse = rev && rev[0] == '1';
	std::vector<pf_emu_warp> warps(nwarps);
	std::vector<Fiber> fibers((size_t)nwarps * PF_WARP);
	g_fn = fn;
	g_arg = arg;
	for (int w = 0; w < nwarps; w++) {
		memset(&warps[w], 0, sizeof(pf_emu_warp));
		warps[w].warp_id = w;
		for (int b = 0; b < 2; b++)
			for (int l = 0; l < PF_WARP; l++) warps[w].seq[b][l] = 0xffffffffu;
		for (int l = 0; l < PF_WARP; l++) {
			Fiber &f = fibers[(size_t)w * PF_WARP + l];
			f.stack = (char *)malloc(STACK);
			f.finished = false;
			f.lane.warp = &warps[w];
			f.lane.lane = l;
			f.lane.seq = 0;
			/* initial frame: six callee-saved slots, then the entry address; after `ret` the stack
			 * must be 8 mod 16 as at any function entry */
			uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
			void **sp = (void **)(top - 8);
			*--sp = (void *)trampoline;
			for (int k = 0; k < 6; k++) *--sp = NULL;
			f.sp = (void *)sp;
		}
	}
	size_t live = fibers.size();
	while (live > 0) {
		live = 0;
		for (int w = 0; w < nwarps; w++) {
			for (int k = 0; k < PF_WARP; k++) {
				int l = reverse ? PF_WARP - 1 - k : k;
				Fiber &f = fibers[(size_t)w * PF_WARP + l];
				if (f.finished) continue;
				g_running = &f;
				pf_emu_cur = &f.lane;
				pf_emu_switch(&g_sched_sp, f.sp);
				if (!f.finished) live++;
			}
		}
	}
	pf_emu_cur = NULL;
	for (size_t i = 0; i < fibers.size(); i++) free(fibers[i].stack);
}
