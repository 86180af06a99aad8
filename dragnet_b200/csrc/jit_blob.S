/* the relocatable kernel (fast_jit.cu) and fscan.cuh, as data for jit.cpp */
	.section .rodata
	.global dng_fast_jit_cubin
	.global dng_fast_jit_cubin_end
	.global dng_fscan_src
	.global dng_fscan_src_end
	.balign 16
dng_fast_jit_cubin:
#ifndef DNG_NO_JIT_CUBIN
	.incbin "build/fast_jit.cubin"
#endif
dng_fast_jit_cubin_end:
	.balign 16
dng_fscan_src:
	.incbin "fscan.cuh"
dng_fscan_src_end:
	.byte 0
	.section .note.GNU-stack,"",@progbits
