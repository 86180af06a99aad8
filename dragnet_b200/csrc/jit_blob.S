/* data for jit.cpp: the LTO-IR builds of the F kernel (fast_jit.cu, one per
 * lane-slice size), its ahead-of-time compiled rare paths (fast_jit_cold.cu)
 * and fscan.cuh as text */
	.section .rodata
	.macro blob name, file
	.global \name
	.global \name\()_end
	.balign 16
\name:
	.incbin "\file"
\name\()_end:
	.endm
	blob dng_jit_hot7, "build/fast_jit7.fatbin"
	blob dng_jit_hot9, "build/fast_jit9.fatbin"
	blob dng_jit_hot11, "build/fast_jit11.fatbin"
	blob dng_jit_hot13, "build/fast_jit13.fatbin"
	blob dng_jit_cold, "build/fast_jit_cold.cubin"
	blob dng_fscan_src, "fscan.cuh"
	.byte 0
	.section .note.GNU-stack,"",@progbits
