/*
 * merge.cu: combining per-shard tallies -- the reference's Manta reduce phase
 * (lib/datasource-manta.js:202-219: `dn scan --points` over the map outputs,
 * i.e. sum `value` over identical `fields` tuples), restated for one process
 * per GPU:
 *
 *   1. every rank serialises the tuple dictionary of its local result,
 *   2. the dictionaries are all-gathered and unioned into one sorted global
 *      dictionary (identical on every rank),
 *   3. each rank scatters its counts into a dense uint64 vector indexed by the
 *      global dictionary (the scan counters ride at the end of the vector),
 *   4. ONE sum-reduce of that vector yields the merged tallies on the root.
 *
 * The dictionary/dense steps are plain host functions so that the same merge
 * can be driven over gloo in CPU tests; dng_merge_nccl() runs steps 2 and 4
 * over NCCL (ncclAllGather + one ncclReduce) on device buffers.  NCCL is
 * dlopen()ed so the library loads on boxes without it.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dragnet_gpu.h"
#include "result.h"
#include "plan.h"

/* api.cu: scratch buffers come from (and go back to) the per-process cache;
 * cudaMalloc/cudaFree per merge cost milliseconds and serialise across the
 * processes of a multi-GPU job */
cudaError_t dng_cached_alloc(int device, void **p, size_t n);
void dng_cached_free(void *p);

static size_t pow2_at_least(size_t n)
{
	size_t c = 4096;
	while (c < n)
		c <<= 1;
	return c;
}

namespace {

bool dict_parse(const void *buf, size_t len, std::vector<std::string> &keys)
{
	const unsigned char *p = (const unsigned char *)buf;
	if (len < 4)
		return false;
	uint32_t n;
	memcpy(&n, p, 4);
	size_t o = 4;
	for (uint32_t i = 0; i < n; i++) {
		uint32_t kl;
		if (o + 4 > len)
			return false;
		memcpy(&kl, p + o, 4);
		o += 4;
		if (o + kl > len)
			return false;
		keys.emplace_back((const char *)p + o, kl);
		o += kl;
	}
	return true;
}

std::string dict_serialize(const std::vector<std::string> &keys)
{
	std::string s;
	uint32_t n = (uint32_t)keys.size();
	s.append((const char *)&n, 4);
	for (auto &k : keys) {
		uint32_t kl = (uint32_t)k.size();
		s.append((const char *)&kl, 4);
		s.append(k);
	}
	return s;
}

} /* namespace */

extern "C" {

int dng_result_from_points(const dng_plan *plan, size_t npoints,
    const char *const *strs, const size_t *strlens, const double *numvals,
    const uint64_t *values, dng_result **out)
{
	if (!plan || !out || (npoints && !values))
		return DNG_EINVAL;
	dng_result *r = new dng_result();
	r->init_from_plan(plan);
	std::vector<std::pair<std::string, uint64_t>> rows;
	if (r->nmetrics != 1) {
		delete r;
		return DNG_EINVAL;	/* points of one metric only */
	}
	const int nc = r->ncols[0];
	for (size_t i = 0; i < npoints; i++) {
		std::string k;
		for (int j = 0; j < nc; j++) {
			size_t x = i * nc + j;
			if (r->col_kind[0][j] == dng::COL_DISCRETE) {
				size_t n = strlens[x];
				k += (char)(n & 0xff);
				k += (char)((n >> 8) & 0xff);
				k.append(strs[x], n);
			} else {
				double v = numvals[x], o;
				if (r->col_kind[0][j] == dng::COL_P2) {
					int e = 0;
					if (v != v || std::isinf(v))
						o = v;
					else if (v < 1)
						o = 0;
					else {
						std::frexp(v, &e);
						o = e;
					}
				} else {
					double q = v / r->col_step[0][j];
					o = (q != q) ? q : std::floor(q) + 0.0;
				}
				uint64_t b;
				if (o != o)
					b = 0x7ff8000000000000ull;
				else
					memcpy(&b, &o, 8);
				k += (char)0xFF;
				k += (char)0xFF;
				for (int t = 0; t < 8; t++)
					k += (char)((b >> (8 * t)) & 0xff);
			}
		}
		rows.emplace_back(std::move(k), values[i]);
	}
	std::sort(rows.begin(), rows.end());
	for (auto &kv : rows) {
		if (!r->keys.empty() && r->keys.back() == kv.first) {
			r->values.back() += kv.second;
		} else {
			r->keys.push_back(kv.first);
			r->values.push_back(kv.second);
		}
	}
	r->finalize();
	*out = r;
	return DNG_OK;
}

int dng_result_dict(const dng_result *cr, const void **buf, size_t *len)
{
	if (!cr || !buf || !len)
		return DNG_EINVAL;
	dng_result *r = const_cast<dng_result *>(cr);
	if (r->dict.empty())
		r->dict = dict_serialize(r->keys);
	*buf = r->dict.data();
	*len = r->dict.size();
	return DNG_OK;
}

int dng_dict_union(const void *const *bufs, const size_t *lens, size_t n,
    void **out, size_t *outlen)
{
	if (!bufs || !lens || !out || !outlen)
		return DNG_EINVAL;
	std::vector<std::string> all;
	for (size_t i = 0; i < n; i++)
		if (!dict_parse(bufs[i], lens[i], all))
			return DNG_EINVAL;
	std::sort(all.begin(), all.end());
	all.erase(std::unique(all.begin(), all.end()), all.end());
	std::string s = dict_serialize(all);
	void *m = malloc(s.size() ? s.size() : 1);
	if (!m)
		return DNG_ENOMEM;
	memcpy(m, s.data(), s.size());
	*out = m;
	*outlen = s.size();
	return DNG_OK;
}

void dng_buf_free(void *p)
{
	free(p);
}

size_t dng_dict_count(const void *dict, size_t len)
{
	if (!dict || len < 4)
		return 0;
	uint32_t n;
	memcpy(&n, dict, 4);
	return n;
}

int dng_result_dense(const dng_result *r, const void *dict, size_t dictlen,
    uint64_t *vec, size_t n)
{
	if (!r || !dict || !vec)
		return DNG_EINVAL;
	std::vector<std::string> keys;
	if (!dict_parse(dict, dictlen, keys) || keys.size() != n)
		return DNG_EINVAL;
	for (size_t i = 0; i < n; i++)
		vec[i] = 0;
	for (size_t i = 0; i < r->keys.size(); i++) {
		auto it = std::lower_bound(keys.begin(), keys.end(), r->keys[i]);
		if (it == keys.end() || *it != r->keys[i])
			return DNG_EINVAL;
		vec[it - keys.begin()] = r->values[i];
	}
	return DNG_OK;
}

int dng_result_from_dense(const dng_result *like, const void *dict,
    size_t dictlen, const uint64_t *vec, size_t n, dng_result **out)
{
	if (!like || !dict || !vec || !out)
		return DNG_EINVAL;
	std::vector<std::string> keys;
	if (!dict_parse(dict, dictlen, keys) || keys.size() != n)
		return DNG_EINVAL;
	dng_result *r = new dng_result();
	r->nmetrics = like->nmetrics;
	memcpy(r->ncols, like->ncols, sizeof (r->ncols));
	memcpy(r->col_kind, like->col_kind, sizeof (r->col_kind));
	memcpy(r->col_step, like->col_step, sizeof (r->col_step));
	for (size_t i = 0; i < n; i++) {
		r->keys.push_back(keys[i]);
		r->values.push_back(vec[i]);
	}
	r->finalize();
	*out = r;
	return DNG_OK;
}

} /* extern "C" */

/* ---- NCCL transport -------------------------------------------------------- */

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_SUM = 0 };

struct NcclApi {
	void *lib = nullptr;
	int (*GetUniqueId)(nccl_uid *) = nullptr;
	int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, nccl_comm,
	    cudaStream_t) = nullptr;
	int (*Reduce)(const void *, void *, size_t, int, int, int, nccl_comm,
	    cudaStream_t) = nullptr;
	int (*CommDestroy)(nccl_comm) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	bool ok = false;
};

NcclApi &nccl()
{
	static NcclApi api;
	if (api.lib)
		return api;
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char *n : names) {
		api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (api.lib)
			break;
	}
	if (!api.lib)
		return api;
	*(void **)&api.GetUniqueId = dlsym(api.lib, "ncclGetUniqueId");
	*(void **)&api.CommInitRank = dlsym(api.lib, "ncclCommInitRank");
	*(void **)&api.AllGather = dlsym(api.lib, "ncclAllGather");
	*(void **)&api.Reduce = dlsym(api.lib, "ncclReduce");
	*(void **)&api.CommDestroy = dlsym(api.lib, "ncclCommDestroy");
	*(void **)&api.GetErrorString = dlsym(api.lib, "ncclGetErrorString");
	api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather &&
	    api.Reduce && api.CommDestroy;
	return api;
}

} /* namespace */

struct dng_comm {
	nccl_comm comm = nullptr;
	int nranks = 1, rank = 0, device = 0;
	cudaStream_t stream = nullptr;
};

extern "C" {

int dng_comm_unique_id(void *id128)
{
	NcclApi &n = nccl();
	if (!n.ok || !id128)
		return DNG_ENCCL;
	nccl_uid id;
	if (n.GetUniqueId(&id) != 0)
		return DNG_ENCCL;
	memcpy(id128, &id, sizeof (id));
	return DNG_OK;
}

int dng_comm_init(dng_comm **out, int nranks, int rank, const void *id128,
    int device, char *err, size_t errlen)
{
	NcclApi &n = nccl();
	if (!n.ok) {
		if (err && errlen)
			snprintf(err, errlen, "NCCL (libnccl.so.2) not available");
		return DNG_ENCCL;
	}
	if (!out || !id128)
		return DNG_EINVAL;
	if (cudaSetDevice(device) != cudaSuccess)
		return DNG_ENODEV;
	dng_comm *c = new dng_comm();
	c->nranks = nranks;
	c->rank = rank;
	c->device = device;
	nccl_uid id;
	memcpy(&id, id128, sizeof (id));
	int rc = n.CommInitRank(&c->comm, nranks, id, rank);
	if (rc != 0) {
		if (err && errlen)
			snprintf(err, errlen, "ncclCommInitRank: %s",
			    n.GetErrorString ? n.GetErrorString(rc) : "error");
		delete c;
		return DNG_ENCCL;
	}
	cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	*out = c;
	return DNG_OK;
}

void dng_comm_destroy(dng_comm *c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	if (c->comm)
		nccl().CommDestroy(c->comm);
	if (c->stream)
		cudaStreamDestroy(c->stream);
	delete c;
}

int dng_merge_nccl(dng_scan *scan, dng_comm *c, int root, dng_result **out,
    dng_counters *counters)
{
	if (!scan || !c || !out)
		return DNG_EINVAL;
	NcclApi &n = nccl();
	*out = nullptr;
	dng_result *local = nullptr;
	int rc = dng_scan_finish(scan, &local);
	if (rc)
		return rc;
	dng_counters lc;
	rc = dng_scan_counters(scan, &lc);
	if (rc) {
		dng_result_destroy(local);
		return rc;
	}
	cudaSetDevice(c->device);
	const void *dict;
	size_t dlen;
	dng_result_dict(local, &dict, &dlen);

	/* (a) all-gather dictionary sizes, then the padded dictionaries */
	unsigned long long *d_sz = nullptr;
	dng_cached_alloc(c->device, (void **)&d_sz, 4096);
	unsigned long long mysz = dlen;
	cudaMemcpyAsync(d_sz + c->nranks, &mysz, 8, cudaMemcpyHostToDevice,
	    c->stream);
	int nrc = n.AllGather(d_sz + c->nranks, d_sz, 8, NCCL_UINT8, c->comm,
	    c->stream);
	std::vector<unsigned long long> sizes(c->nranks);
	cudaMemcpyAsync(sizes.data(), d_sz, 8 * c->nranks,
	    cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	size_t maxsz = 16;
	for (auto v : sizes)
		maxsz = std::max(maxsz, (size_t)v);
	maxsz = (maxsz + 15) & ~(size_t)15;
	unsigned char *d_all = nullptr, *d_mine = nullptr;
	dng_cached_alloc(c->device, (void **)&d_all,
	    pow2_at_least(maxsz * c->nranks));
	dng_cached_alloc(c->device, (void **)&d_mine, pow2_at_least(maxsz));
	cudaMemsetAsync(d_mine, 0, maxsz, c->stream);
	cudaMemcpyAsync(d_mine, dict, dlen, cudaMemcpyHostToDevice, c->stream);
	if (!nrc)
		nrc = n.AllGather(d_mine, d_all, maxsz, NCCL_UINT8, c->comm,
		    c->stream);
	std::vector<unsigned char> all(maxsz * c->nranks);
	cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost,
	    c->stream);
	cudaStreamSynchronize(c->stream);
	std::vector<const void *> bufs(c->nranks);
	std::vector<size_t> lens(c->nranks);
	for (int i = 0; i < c->nranks; i++) {
		bufs[i] = all.data() + (size_t)i * maxsz;
		lens[i] = (size_t)sizes[i];
	}
	void *gdict = nullptr;
	size_t glen = 0;
	if (!nrc)
		rc = dng_dict_union(bufs.data(), lens.data(), c->nranks, &gdict,
		    &glen);

	/* (b) dense tallies (+ counters) and the single sum-reduce */
	size_t G = rc || nrc ? 0 : dng_dict_count(gdict, glen);
	const size_t NC = sizeof (dng_counters) / sizeof (uint64_t);
	std::vector<uint64_t> vec(G + NC);
	if (!rc && !nrc)
		rc = dng_result_dense(local, gdict, glen, vec.data(), G);
	memcpy(vec.data() + G, &lc, sizeof (lc));
	uint64_t *d_vec = nullptr, *d_red = nullptr;
	dng_cached_alloc(c->device, (void **)&d_vec, pow2_at_least(vec.size() * 8));
	dng_cached_alloc(c->device, (void **)&d_red, pow2_at_least(vec.size() * 8));
	cudaMemcpyAsync(d_vec, vec.data(), vec.size() * 8,
	    cudaMemcpyHostToDevice, c->stream);
	if (!rc && !nrc)
		nrc = n.Reduce(d_vec, d_red, vec.size(), NCCL_UINT64, NCCL_SUM,
		    root, c->comm, c->stream);
	if (!rc && !nrc && c->rank == root) {
		cudaMemcpyAsync(vec.data(), d_red, vec.size() * 8,
		    cudaMemcpyDeviceToHost, c->stream);
		cudaStreamSynchronize(c->stream);
		rc = dng_result_from_dense(local, gdict, glen, vec.data(), G,
		    out);
		if (counters)
			memcpy(counters, vec.data() + G, sizeof (*counters));
	} else {
		cudaStreamSynchronize(c->stream);
	}
	dng_cached_free(d_sz);
	dng_cached_free(d_all);
	dng_cached_free(d_mine);
	dng_cached_free(d_vec);
	dng_cached_free(d_red);
	dng_buf_free(gdict);
	dng_result_destroy(local);
	if (nrc)
		return DNG_ENCCL;
	return rc;
}

} /* extern "C" */
