/*
 * jit.h: run-time compilation of a scan's record templates (tmpl.h / fast.h)
 * into the matcher of scan_kernel_j -- the F kernel with its interpreted trie
 * walk replaced by straight-line code with the literals as immediates.
 *
 * What is compiled at run time is ONLY the matcher, dng_jmatch(): a few
 * hundred lines generated from the trie blob plus the wildcard scanners of
 * fscan.cuh (embedded as text).  NVRTC turns it into a relocatable cubin and
 * nvJitLink links that with the relocatable build of the kernel that ships
 * inside the library (fast_jit.cu), so everything else -- chunk pipeline,
 * newline index, stages, tally, miss handling -- is the code of scan_kernel_f,
 * compiled ahead of time.  Results cannot depend on this choice: the generated
 * matcher accepts exactly what fmatch() accepts (same trie, same scanners) and
 * everything it rejects is parsed by the general code.
 *
 * Kernels are cached per process by a hash of the generated source.  Both
 * libraries are dlopen()ed; without them (or with DNG_JIT=0) scans use the
 * interpreted matcher.
 */
#ifndef DNG_JIT_H
#define DNG_JIT_H

#include <cuda_runtime.h>

#include <atomic>
#include <memory>
#include <string>

#include "plan.h"

namespace dng {

struct JitKernels {
	cudaLibrary_t lib = nullptr;
	cudaKernel_t kern[4] = {};	/* lane slices of 7, 9, 11, 13 units */
	std::atomic<int> state{0};	/* 0 being built, 1 ready, 2 failed */
	bool ok = false;
	std::string err;
	double compile_ms = 0, link_ms = 0;
};

/* the CUDA source of dng_jmatch() for an F trie blob (tmpl_build, compact);
 * `prelude` replaces the device definitions the generated code builds on
 * (tests/hostcheck compiles the same code for the host with its own) */
std::string jit_source(const u8 *blob, size_t bytes,
    const char *prelude = nullptr);

/*
 * The kernels for this source on device `dev`, from the cache or built now:
 * on the caller's thread if `wait` (then state is 1 or 2 on return), else on a
 * worker thread (state 0 until it is done: the caller keeps using the
 * interpreted matcher meanwhile).  smem_max = dynamic shared memory the
 * kernels may be launched with.
 */
std::shared_ptr<JitKernels> jit_request(const std::string &source, int dev,
    int smem_max, bool wait);

/* only compile + link, to `cubin` (no device needed: tests) */
bool jit_build(const std::string &source, std::string &cubin, std::string &err,
    double *compile_ms, double *link_ms);

} /* namespace dng */

#endif
