// emu_rt.cpp — context switch for the CPU SIMT emulation (tests only; see simt_emu.h)
#ifdef RCZ_EMU
#if !defined(__x86_64__)
#error "the emulation harness is x86-64 only"
#endif
// void rcz_emu_switch(void** from_sp, void* to_sp): save callee-saved registers, swap stacks.
asm(R"(
.text
.globl rcz_emu_switch
.type rcz_emu_switch,@function
rcz_emu_switch:
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
.size rcz_emu_switch,.-rcz_emu_switch
)");
#endif

#ifdef RCZ_EMU
// RCZ_EMU_BACKTRACE=1: print a backtrace on SIGSEGV (debugging aid for kernels run on the emulator)
#include <execinfo.h>
#include <signal.h>
#include <stdlib.h>
#include <unistd.h>
static void rcz_emu_segv(int) {
    void* bt[48];
    const int n = backtrace(bt, 48);
    backtrace_symbols_fd(bt, n, 2);
    _exit(139);
}
__attribute__((constructor)) static void rcz_emu_install_segv() {
    if (!getenv("RCZ_EMU_BACKTRACE")) return;
    static char alt[1 << 16];
    stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof(alt); ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa; sa.sa_handler = rcz_emu_segv; sigemptyset(&sa.sa_mask); sa.sa_flags = SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
}
#endif
