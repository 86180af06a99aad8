#include "result.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

using namespace dng;

double dng_bucket_min(u8 kind, double step, double o)
{
	if (kind == COL_P2) {
		/* bucketMin(0) = 0, bucketMin(i) = 2^(i-1) */
		if (o != o)
			return o;
		if (o == 0)
			return 0.0;
		if (std::isinf(o))
			return o;
		return std::ldexp(1.0, (int)o - 1);
	}
	return o * step;	/* linear: i * step */
}

void dng_result::init_from_plan(const dng_plan *p)
{
	nmetrics = p->nmetrics;
	for (int m = 0; m < nmetrics; m++) {
		ncols[m] = p->ncols[m];
		for (int i = 0; i < ncols[m]; i++) {
			col_kind[m][i] = p->col_kind[m][i];
			col_step[m][i] = p->col_step[m][i];
		}
	}
}

void dng_result::finalize()
{
	/* a metric with zero decomps always emits exactly one point, even for
	 * empty input (tests/dn/local/tst.empty.sh.out:1-19) */
	for (int m = 0; m < nmetrics; m++) {
		if (ncols[m] != 0)
			continue;
		std::string k;
		if (nmetrics > 1) {
			k += (char)0xFD;
			k += (char)m;
		}
		if (std::find(keys.begin(), keys.end(), k) == keys.end()) {
			keys.push_back(k);
			values.push_back(0);
		}
	}
	std::vector<size_t> order(keys.size());
	std::iota(order.begin(), order.end(), 0);
	std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
		return keys[a] < keys[b];
	});
	std::vector<std::string> k2(keys.size());
	std::vector<uint64_t> v2(keys.size());
	for (size_t i = 0; i < order.size(); i++) {
		k2[i] = std::move(keys[order[i]]);
		v2[i] = values[order[i]];
	}
	keys.swap(k2);
	values.swap(v2);
	cells.clear();
	cell0.assign(keys.size(), 0);
	metric.assign(keys.size(), 0);
	for (size_t i = 0; i < keys.size(); i++) {
		const std::string &k = keys[i];
		size_t o = 0;
		int m = 0;
		if (nmetrics > 1 && k.size() >= 2) {
			m = (unsigned char)k[1];
			o = 2;
		}
		if (m >= nmetrics)
			m = 0;
		metric[i] = m;
		cell0[i] = cells.size();
		for (int j = 0; j < ncols[m]; j++) {
			Cell c{0, 0, 0.0, 0};
			if (o + 2 <= k.size()) {
				unsigned n = (unsigned char)k[o] |
				    ((unsigned char)k[o + 1] << 8);
				o += 2;
				if (n == 0xFFFF) {
					uint64_t b = 0;
					for (int x = 0; x < 8; x++)
						b |= (uint64_t)(unsigned char)
						    k[o + x] << (8 * x);
					double ord;
					memcpy(&ord, &b, 8);
					c.is_number = 1;
					c.num = dng_bucket_min(col_kind[m][j],
					    col_step[m][j], ord);
					o += 8;
				} else {
					c.off = o;
					c.len = n;
					o += n;
				}
			}
			cells.push_back(c);
		}
	}
}
