/*
 * gen.cuh: deterministic synthetic input in the shape of the reference's
 * tools/mktestdata (:15-99 field configuration, :138-190 record builder),
 * used by bench.py and the parity tests.  The reference draws from unseeded
 * Math.random(); here every record j of a stream draws from splitmix64 keyed
 * by (seed, j), so the host generator and the CUDA generator kernel produce
 * byte-identical chunks and any record range can be produced independently.
 *
 * Key order matches JSON.stringify of the object mktestdata builds (see
 * tests/data/2014/05-01/one.log:1): time, host, req{method,url[,caller]},
 * operation, res{statusCode}, latency, dataLatency, dataSize.
 */
#ifndef DNG_GEN_CUH
#define DNG_GEN_CUH

#include "jsnum.cuh"
#include "../../include/dragnet_gpu.h"

namespace dng {

enum { GEN_MAXREC = 320 };

struct GenRng {
	uint64_t s;
	DNG_HD uint64_t next() {
		uint64_t z = (s += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	DNG_HD double uniform() {		/* [0, 1) like Math.random() */
		return (double)(next() >> 11) * (1.0 / 9007199254740992.0);
	}
	DNG_HD uint32_t pick(uint32_t n) {	/* floor(random() * n) */
		return (uint32_t)(uniform() * n);
	}
};

DNG_HD int gen_put(char *o, int n, const char *s)
{
	while (*s)
		o[n++] = *s++;
	return n;
}

DNG_HD int gen_putu(char *o, int n, uint64_t v, int width)
{
	char t[24];
	int k = 0;
	do {
		t[k++] = (char)('0' + v % 10);
		v /= 10;
	} while (v);
	while (k < width)
		t[k++] = '0';
	while (k)
		o[n++] = t[--k];
	return n;
}

/* mktestdata's probdist draw (:168-179): two Math.random() calls */
DNG_HD uint64_t gen_dist(GenRng &r)
{
	double rand = r.uniform();
	double lo, hi;
	if (0.4 > rand) { lo = 1; hi = 5; }
	else if (0.7 > rand) { lo = 20; hi = 30; }
	else if (0.7999999999999999 > rand) { lo = 100; hi = 200; }
	else { lo = 1024; hi = 4096; }
	return (uint64_t)floor(r.uniform() * (hi - lo) + lo + 0.5);
}

/* writes record j (with trailing '\n') into out; returns its length */
DNG_HD int gen_record(const dng_gen_params &p, uint64_t j, char *out)
{
	GenRng r;
	/* per-record stream: scramble (seed, j) so that neighbouring records
	 * do not walk overlapping stretches of the splitmix64 sequence */
	r.s = p.seed * 0xD1342543DE82EF95ull + j * 0xD6E8FEB86659FD93ull;
	r.s = r.next() ^ (j << 17);
	int n = 0;
	/* time: Math.round(j / nrecords * (max - min) + min) */
	double ts = floor((double)j / (double)p.total_records *
	    (double)(p.time_max_ms - p.time_min_ms) + (double)p.time_min_ms +
	    0.5);
	int64_t ms = (int64_t)ts;
	int64_t days = ms / 86400000, rem = ms % 86400000;
	if (rem < 0) {
		rem += 86400000;
		days--;
	}
	int64_t z = days + 719468;
	int64_t era = (z >= 0 ? z : z - 146096) / 146097;
	int64_t doe = z - era * 146097;
	int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
	int64_t y = yoe + era * 400;
	int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
	int64_t mp = (5 * doy + 2) / 153;
	int64_t d = doy - (153 * mp + 2) / 5 + 1;
	int64_t m = mp + (mp < 10 ? 3 : -9);
	y += m <= 2;
	n = gen_put(out, n, "{\"time\":\"");
	n = gen_putu(out, n, (uint64_t)y, 4);
	out[n++] = '-';
	n = gen_putu(out, n, (uint64_t)m, 2);
	out[n++] = '-';
	n = gen_putu(out, n, (uint64_t)d, 2);
	out[n++] = 'T';
	n = gen_putu(out, n, (uint64_t)(rem / 3600000), 2);
	out[n++] = ':';
	n = gen_putu(out, n, (uint64_t)(rem / 60000 % 60), 2);
	out[n++] = ':';
	n = gen_putu(out, n, (uint64_t)(rem / 1000 % 60), 2);
	out[n++] = '.';
	n = gen_putu(out, n, (uint64_t)(rem % 1000), 3);
	n = gen_put(out, n, "Z\",\"host\":\"");
	const char *hosts[5] = { "ralph", "janey", "kearney", "sherri",
	    "wendell" };
	n = gen_put(out, n, hosts[r.pick(5)]);
	const char *methods[4] = { "HEAD", "GET", "PUT", "DELETE" };
	uint32_t mi = r.pick(4);
	n = gen_put(out, n, "\",\"req\":{\"method\":\"");
	n = gen_put(out, n, methods[mi]);
	const char *ops[4][3] = {
	    { "headstorage", "headpublicstorage", "" },
	    { "getjoberrors", "getpublicstorage", "getstorage" },
	    { "putdirectory", "putpublicobject", "putobject" },
	    { "deletestorage", "deletepublicstorage", "" } };
	const char *op = ops[mi][r.pick((mi == 1 || mi == 2) ? 3 : 2)];
	n = gen_put(out, n, "\",\"url\":\"/random/url/number/");
	n = gen_putu(out, n, r.pick(500), 1);
	out[n++] = '"';
	uint32_t ci = r.pick(4);
	if (ci == 0)
		n = gen_put(out, n, ",\"caller\":\"admin\"");
	else if (ci == 1)
		n = gen_put(out, n, ",\"caller\":\"poseidon\"");
	else if (ci == 2)
		n = gen_put(out, n, ",\"caller\":null");
	n = gen_put(out, n, "},\"operation\":\"");
	n = gen_put(out, n, op);
	const uint32_t codes[7] = { 200, 204, 400, 404, 499, 500, 503 };
	n = gen_put(out, n, "\",\"res\":{\"statusCode\":");
	n = gen_putu(out, n, codes[r.pick(7)], 1);
	n = gen_put(out, n, "},\"latency\":");
	if (p.string_latency)
		out[n++] = '"';
	n = gen_putu(out, n, gen_dist(r), 1);
	if (p.string_latency)
		out[n++] = '"';
	n = gen_put(out, n, ",\"dataLatency\":");
	n = gen_putu(out, n, gen_dist(r), 1);
	n = gen_put(out, n, ",\"dataSize\":");
	r.uniform();		/* the unused `rand` of a one-entry probdist */
	n = gen_putu(out, n, (uint64_t)floor(r.uniform() * 1073741824.0 + 0.5),
	    1);
	out[n++] = '}';
	out[n++] = '\n';
	return n;
}

} /* namespace dng */
#endif
