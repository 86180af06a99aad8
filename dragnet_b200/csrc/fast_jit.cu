/*
 * fast_jit.cu: scan_kernel_f with its matcher left open (fast_kernel.cuh
 * fscan_body<NSL, true>), built as LTO-IR, one fatbin per lane-slice size
 * (-DDNG_JIT_NSL=7|9|11|13), and embedded in the library.  At run time jit.cpp
 * compiles a scan's templates into dng_jmatch() with NVRTC (LTO-IR as well) and
 * nvJitLink optimises the two together: the matcher is inlined into the
 * kernel's record loop, registers are allocated across it, nothing is called.
 * The kernel's rare paths (fast_jit_cold.cu) are compiled ahead of time and
 * only linked.  Everything but the matcher is the very code of scan_kernel_f.
 */
#define DNG_NO_GENERAL_KERNELS
#define DNG_JIT_HOT
#include "fast_kernel.cuh"

using namespace dng;

#ifndef DNG_JIT_NSL
#define DNG_JIT_NSL 13
#endif

extern "C" __global__ void __launch_bounds__(DNG_F_NT, 1)
dng_scan_kernel_j(const FScanArgs a)
{
	fscan_body<DNG_JIT_NSL, true>(a);
}
