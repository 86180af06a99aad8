/*
 * jit_load.cpp: the per-process cache of run-time compiled kernels (jit.h),
 * their compilation off the caller's thread and their loading into the CUDA
 * runtime.
 */
#include <stdlib.h>

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

/*
 * Compiler threads are joined before the process goes away: by an atexit()
 * handler registered once the compiler libraries are loaded -- so that it runs
 * BEFORE their own exit handlers and static destructors (a compile in flight
 * while libnvrtc tears itself down ends in free(): invalid pointer) -- and, for
 * good measure, when this library's statics go.
 */
struct Workers {
	std::vector<std::thread> th;
	void join_all() {
		std::vector<std::thread> mine;
		{
			std::lock_guard<std::mutex> g(g_mu);
			mine.swap(th);
		}
		for (auto &t : mine)
			if (t.joinable())
				t.join();
	}
	~Workers() { join_all(); }
} g_workers;

bool g_exiting = false;

void join_workers_at_exit()
{
	{
		std::lock_guard<std::mutex> g(g_mu);
		g_exiting = true;
	}
	g_workers.join_all();
}

typedef std::pair<int, u64> CacheKey;

void build_into(std::shared_ptr<JitKernels> k, std::string source, int nsl,
    int dev, int smem_max, CacheKey key)
{
	/* one build at a time; by the time it is this one's turn the scan that
	 * asked may be gone (short scans, test suites), or the process on its
	 * way out: then there is nothing to build (and the next request for
	 * this source starts afresh) */
	static std::mutex turn;
	std::lock_guard<std::mutex> my_turn(turn);
	{
		std::lock_guard<std::mutex> g(g_mu);
		/* (the cache, this thread, the scan) */
		if (g_exiting || k.use_count() <= 2) {
			auto it = g_cache.find(key);
			if (it != g_cache.end() && it->second == k)
				g_cache.erase(it);
			k->err = "abandoned";
			k->state.store(2);
			g_cv.notify_all();
			return;
		}
	}
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
	const CacheKey key = std::make_pair(dev, hsh);
	auto it = g_cache.find(key);
	if (it != g_cache.end()) {
		k = it->second;
	} else {
		/* (failures are cached too: not tried again for every scan) */
		k = std::make_shared<JitKernels>();
		g_cache[key] = k;
		if (wait) {
			g.unlock();
			build_into(k, source, nsl, dev, smem_max, key);
			g.lock();
		} else {
			/* (the libraries first, here, and then the handler that
			 * waits for the thread: see Workers) */
			g.unlock();
			jit_prepare();
			static std::once_flag once;
			std::call_once(once, [] { atexit(join_workers_at_exit); });
			g.lock();
			g_workers.th.emplace_back(build_into, k, source, nsl, dev,
			    smem_max, key);
		}
	}
	if (wait)
		g_cv.wait(g, [&] { return k->state.load() != 0; });
	return k;
}

} /* namespace dng */
