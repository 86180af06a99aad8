/*
 * fast_jit.cu: scan_kernel_f with its matcher left open (fast_kernel.cuh
 * fscan_body<NSL, true>): built as RELOCATABLE device code and embedded in the
 * library as a cubin; at run time jit.cpp compiles a scan's templates into
 * dng_jmatch() with NVRTC and links the two with nvJitLink.  Everything but
 * the matcher is the very code of scan_kernel_f.
 */
#define DNG_NO_GENERAL_KERNELS
#include "fast_kernel.cuh"

using namespace dng;

extern "C" __global__ void __launch_bounds__(DNG_NT, 1)
dng_scan_kernel_j7(const FScanArgs a)
{
	fscan_body<7, true>(a);
}

extern "C" __global__ void __launch_bounds__(DNG_NT, 1)
dng_scan_kernel_j9(const FScanArgs a)
{
	fscan_body<9, true>(a);
}

extern "C" __global__ void __launch_bounds__(DNG_NT, 1)
dng_scan_kernel_j11(const FScanArgs a)
{
	fscan_body<11, true>(a);
}

extern "C" __global__ void __launch_bounds__(DNG_NT, 1)
dng_scan_kernel_j13(const FScanArgs a)
{
	fscan_body<13, true>(a);
}
