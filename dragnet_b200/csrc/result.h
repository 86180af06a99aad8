/*
 * result.h: host-side result set -- the skinner points a scan emits
 * (lib/stream-scan.js:81-84, `resultsAsPoints: true`).
 *
 * Group keys arrive from the device as encoded byte strings (record.cuh
 * process_record): per column either u16 len + bytes (discrete: the JS
 * String(value)) or 0xFFFF + 8 bytes of bucket ordinal (binary64).  Points
 * carry bucketMin(ordinal) for quantized columns (bin/dn:1020,1194; goldens
 * tst.scan_file.sh.out:307-314).
 */
#ifndef DNG_RESULT_H
#define DNG_RESULT_H

#include <stdint.h>
#include <string>
#include <vector>
#include "plan.h"

struct dng_result {
	int nmetrics = 1;
	int ncols[dng::MAX_METRICS] = {};
	dng::u8 col_kind[dng::MAX_METRICS][dng::MAX_COLS];
	double col_step[dng::MAX_METRICS][dng::MAX_COLS];
	/* sorted by encoded key */
	std::vector<std::string> keys;
	std::vector<uint64_t> values;
	/* decoded columns, filled by finalize() */
	struct Cell { size_t off, len; double num; uint8_t is_number; };
	std::vector<Cell> cells;	/* point i: cells[cell0[i] ...] */
	std::vector<size_t> cell0;
	std::vector<int> metric;	/* metric of point i */
	std::string dict;		/* serialised dictionary (lazy) */

	void init_from_plan(const dng_plan *p);
	/* sort, apply the "no breakdowns => exactly one point" rule, decode */
	void finalize();
};

double dng_bucket_min(dng::u8 kind, double step, double ordinal);

#endif
