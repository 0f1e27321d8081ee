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
