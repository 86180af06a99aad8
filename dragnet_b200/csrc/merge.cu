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
	/* allocated with the communicator, so that the agreement rounds of a
	 * merge never depend on an allocation succeeding: 2 x (nranks + 1)
	 * 64-bit words on the device, as many pinned on the host */
	unsigned long long *d_hdr = nullptr, *h_hdr = nullptr;
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
	const size_t hb = 2 * ((size_t)nranks + 1) * sizeof (unsigned long long);
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) !=
	    cudaSuccess || cudaMalloc((void **)&c->d_hdr, hb) != cudaSuccess ||
	    cudaMallocHost((void **)&c->h_hdr, hb) != cudaSuccess) {
		if (err && errlen)
			snprintf(err, errlen, "communicator scratch: %s",
			    cudaGetErrorString(cudaGetLastError()));
		dng_comm_destroy(c);
		return DNG_ECUDA;
	}
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
	if (c->d_hdr)
		cudaFree(c->d_hdr);
	if (c->h_hdr)
		cudaFreeHost(c->h_hdr);
	delete c;
}

/*
 * Every rank tells every rank a (size, status) pair: one all-gather over the
 * communicator's own scratch.  Returns this rank's status if it is non-zero,
 * else the first non-zero status of a peer, else 0 -- the SAME verdict
 * (zero or not) on every rank, so that all of them go on to the next
 * collective or none does.  DNG_ENCCL if the collective itself failed (then
 * nothing more can be agreed on).  sizes (nranks) may be null.
 */
static int merge_agree(dng_comm *c, unsigned long long size, int status,
    unsigned long long *sizes)
{
	NcclApi &n = nccl();
	const int R = c->nranks;
	unsigned long long *h = c->h_hdr;
	h[2 * R] = size;
	h[2 * R + 1] = (unsigned long long)(long long)status;
	if (cudaMemcpyAsync(c->d_hdr + 2 * R, h + 2 * R, 16,
	    cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
		return DNG_ECUDA;	/* (the copy engine is gone: so is NCCL) */
	if (n.AllGather(c->d_hdr + 2 * R, c->d_hdr, 16, NCCL_UINT8, c->comm,
	    c->stream) != 0)
		return DNG_ENCCL;
	if (cudaMemcpyAsync(h, c->d_hdr, 16 * (size_t)R, cudaMemcpyDeviceToHost,
	    c->stream) != cudaSuccess ||
	    cudaStreamSynchronize(c->stream) != cudaSuccess)
		return DNG_ECUDA;
	int peer = 0;
	for (int i = 0; i < R; i++) {
		if (sizes)
			sizes[i] = h[2 * i];
		if (!peer && h[2 * i + 1] != 0)
			peer = (int)(long long)h[2 * i + 1];
	}
	return status ? status : peer;
}

int dng_merge_nccl(dng_scan *scan, dng_comm *c, int root, dng_result **out,
    dng_counters *counters)
{
	if (!scan || !c || !out)
		return DNG_EINVAL;
	NcclApi &n = nccl();
	*out = nullptr;
	const int R = c->nranks;
	/*
	 * A failure on one rank (a full table, an unsupported record, a CUDA
	 * error, an allocation) must not leave its peers blocked in a
	 * collective it never enters: before each data collective the ranks
	 * exchange their status, and either all go on or all return.
	 */
	dng_result *local = nullptr;
	dng_counters lc;
	memset(&lc, 0, sizeof (lc));
	int st = dng_scan_finish(scan, &local);
	if (!st)
		st = dng_scan_counters(scan, &lc);
	if (cudaSetDevice(c->device) != cudaSuccess && !st)
		st = DNG_ECUDA;
	const void *dict = nullptr;
	size_t dlen = 0;
	if (!st)
		dng_result_dict(local, &dict, &dlen);

	unsigned char *d_all = nullptr, *d_mine = nullptr;
	uint64_t *d_vec = nullptr, *d_red = nullptr;
	void *gdict = nullptr;
	size_t glen = 0;
	std::vector<unsigned long long> sizes(R);
	std::vector<unsigned char> all;
	std::vector<uint64_t> vec;
	const size_t NC = sizeof (dng_counters) / sizeof (uint64_t);
	size_t G = 0, maxsz = 16;
	int rc;

	/* (a) dictionary sizes + status of the local scans */
	rc = merge_agree(c, dlen, st, sizes.data());
	if (rc)
		goto done;
	for (auto v : sizes)
		maxsz = std::max(maxsz, (size_t)v);
	maxsz = (maxsz + 15) & ~(size_t)15;
	if (dng_cached_alloc(c->device, (void **)&d_all,
	    pow2_at_least(maxsz * R)) != cudaSuccess ||
	    dng_cached_alloc(c->device, (void **)&d_mine,
	    pow2_at_least(maxsz)) != cudaSuccess)
		st = DNG_ENOMEM;
	try {
		all.resize(maxsz * R);
	} catch (...) {
		st = DNG_ENOMEM;
	}
	if (!st && (cudaMemsetAsync(d_mine, 0, maxsz, c->stream) != cudaSuccess ||
	    cudaMemcpyAsync(d_mine, dict, dlen, cudaMemcpyHostToDevice,
	    c->stream) != cudaSuccess))
		st = DNG_ECUDA;
	rc = merge_agree(c, 0, st, nullptr);
	if (rc)
		goto done;

	/* (b) the padded dictionaries */
	if (n.AllGather(d_mine, d_all, maxsz, NCCL_UINT8, c->comm, c->stream)) {
		rc = DNG_ENCCL;
		goto done;
	}
	if (cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost,
	    c->stream) != cudaSuccess ||
	    cudaStreamSynchronize(c->stream) != cudaSuccess)
		st = DNG_ECUDA;
	if (!st) {
		std::vector<const void *> bufs(R);
		std::vector<size_t> lens(R);
		for (int i = 0; i < R; i++) {
			bufs[i] = all.data() + (size_t)i * maxsz;
			lens[i] = (size_t)sizes[i];
		}
		st = dng_dict_union(bufs.data(), lens.data(), R, &gdict, &glen);
	}
	/* (c) dense tallies (+ counters): same length on every rank */
	if (!st) {
		G = dng_dict_count(gdict, glen);
		try {
			vec.resize(G + NC);
		} catch (...) {
			st = DNG_ENOMEM;
		}
	}
	if (!st)
		st = dng_result_dense(local, gdict, glen, vec.data(), G);
	if (!st) {
		memcpy(vec.data() + G, &lc, sizeof (lc));
		if (dng_cached_alloc(c->device, (void **)&d_vec,
		    pow2_at_least(vec.size() * 8)) != cudaSuccess ||
		    dng_cached_alloc(c->device, (void **)&d_red,
		    pow2_at_least(vec.size() * 8)) != cudaSuccess)
			st = DNG_ENOMEM;
		else if (cudaMemcpyAsync(d_vec, vec.data(), vec.size() * 8,
		    cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
			st = DNG_ECUDA;
	}
	rc = merge_agree(c, G, st, sizes.data());
	if (!rc)
		for (auto v : sizes)
			if (v != G)
				rc = DNG_ENCCL;	/* (cannot happen: same inputs) */
	if (rc)
		goto done;

	/* (d) the single sum-reduce */
	if (n.Reduce(d_vec, d_red, vec.size(), NCCL_UINT64, NCCL_SUM, root,
	    c->comm, c->stream)) {
		rc = DNG_ENCCL;
		goto done;
	}
	if (c->rank == root) {
		if (cudaMemcpyAsync(vec.data(), d_red, vec.size() * 8,
		    cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
		    cudaStreamSynchronize(c->stream) != cudaSuccess) {
			rc = DNG_ECUDA;
			goto done;
		}
		rc = dng_result_from_dense(local, gdict, glen, vec.data(), G,
		    out);
		if (!rc && counters)
			memcpy(counters, vec.data() + G, sizeof (*counters));
	} else if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
		rc = DNG_ECUDA;
	}
done:
	dng_cached_free(d_all);
	dng_cached_free(d_mine);
	dng_cached_free(d_vec);
	dng_cached_free(d_red);
	dng_buf_free(gdict);
	dng_result_destroy(local);
	return rc;
}

} /* extern "C" */
