/*
 * jit_load.cpp: the per-process cache of run-time compiled kernels (jit.h),
 * their compilation off the caller's thread and their loading into the CUDA
 * runtime.
 */
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "jit.h"

namespace dng {

namespace {

std::mutex g_mu;
std::condition_variable g_cv;
std::map<std::pair<int, u64>, std::shared_ptr<JitKernels>> g_cache;

/* compiler threads are joined when the library is unloaded */
struct Workers {
	std::vector<std::thread> th;
	~Workers() {
		for (auto &t : th)
			if (t.joinable())
				t.join();
	}
} g_workers;

void build_into(std::shared_ptr<JitKernels> k, std::string source, int nsl,
    int dev, int smem_max)
{
	std::string cubin, err;
	bool ok = jit_build(source, nsl, cubin, err, &k->compile_ms, &k->link_ms);
	if (ok) {
		cudaSetDevice(dev);
		cudaError_t e = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr,
		    nullptr, 0, nullptr, nullptr, 0);
		if (e == cudaSuccess)
			e = cudaLibraryGetKernel(&k->kern, k->lib, "dng_scan_kernel_j");
		if (e == cudaSuccess)
			e = cudaFuncSetAttribute((const void *)k->kern,
			    cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
		if (e != cudaSuccess) {
			ok = false;
			err = std::string("loading the linked kernel: ") +
			    cudaGetErrorString(e);
		}
	}
	{
		std::lock_guard<std::mutex> g(g_mu);
		k->err = err;
		k->ok = ok;
		k->state.store(ok ? 1 : 2);
	}
	g_cv.notify_all();
}

} /* namespace */

std::shared_ptr<JitKernels> jit_request(const std::string &source, int nsl,
    int dev, int smem_max, bool wait)
{
	u64 hsh = 1469598103934665603ull;
	for (unsigned char c : source)
		hsh = (hsh ^ c) * 1099511628211ull;
	hsh ^= (u64)source.size() << 40;
	hsh = (hsh ^ (u64)nsl) * 1099511628211ull;
	std::unique_lock<std::mutex> g(g_mu);
	std::shared_ptr<JitKernels> k;
	auto it = g_cache.find(std::make_pair(dev, hsh));
	if (it != g_cache.end()) {
		k = it->second;
	} else {
		/* (failures are cached too: not tried again for every scan) */
		k = std::make_shared<JitKernels>();
		g_cache[std::make_pair(dev, hsh)] = k;
		if (wait) {
			g.unlock();
			build_into(k, source, nsl, dev, smem_max);
			g.lock();
		} else {
			g_workers.th.emplace_back(build_into, k, source, nsl, dev,
			    smem_max);
		}
	}
	if (wait)
		g_cv.wait(g, [&] { return k->state.load() != 0; });
	return k;
}

} /* namespace dng */
