/*
 * jsdate.cuh: Date.parse for the ECMAScript date-time string format
 * (ES5 15.9.1.15), as the reference's Datetime parser stage applies it to
 * `date` fields (lib/stream-synthetic.js:65: `parsed = Date.parse(val)`).
 *
 * Grammar: YYYY | YYYY-MM | YYYY-MM-DD | +-YYYYYY-..., optionally followed by
 * THH:mm | THH:mm:ss | THH:mm:ss.s+ and optionally Z | +-HH:mm.  No offset
 * means UTC.  A day of 29..31 in a shorter month carries into the next one,
 * as V8's parser lets it (its day check is 1..31; MakeDay does the rest).
 *
 * V8 falls back to a legacy free-form parser for everything else ("May 1,
 * 2014", "2014-05-01 12:00:00", RFC 2822 ...), which is not restated here.
 * A string that is not in the grammar is called NaN (the reference then counts
 * `baddate`) unless it looks like one of the forms that legacy parser exists
 * for (dng_date_maybe_legacy): such a string makes the scan fail loudly (the
 * `unsupported` counter -> DNG_EUNSUPPORTED) instead of silently dropping
 * records the reference may have kept.
 */
#ifndef DNG_JSDATE_CUH
#define DNG_JSDATE_CUH

#include "jsnum.cuh"

namespace dng {

DNG_HD bool dd2(const uint8_t *p, int &v)
{
	if (p[0] < '0' || p[0] > '9' || p[1] < '0' || p[1] > '9')
		return false;
	v = (p[0] - '0') * 10 + (p[1] - '0');
	return true;
}

/* days since 1970-01-01; 32-bit arithmetic (years are within +-999999), so the
 * divisions by constants compile to multiply-shift instead of the emulated
 * 64-bit divide */
DNG_HD int32_t days_from_civil(int32_t y, int m, int d)
{
	y -= m <= 2;
	int32_t era = (y >= 0 ? y : y - 399) / 400;
	int32_t yoe = y - era * 400;
	int32_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
	int32_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
	return era * 146097 + doe - 719468;
}

/* returns true and *ms on success; false for NaN */
DNG_HDN bool dng_date_parse(const uint8_t *p, int n, int64_t *ms)
{
	int i = 0;
	int32_t y = 0;
	int ysign = 0, ydig = 4;
	if (n >= 1 && (p[0] == '+' || p[0] == '-')) {
		ysign = p[0] == '-' ? -1 : 1;
		ydig = 6;
		i = 1;
	}
	if (n - i < ydig)
		return false;
	for (int k = 0; k < ydig; k++, i++) {
		if (p[i] < '0' || p[i] > '9')
			return false;
		y = y * 10 + (p[i] - '0');
	}
	if (ysign < 0) {
		if (y == 0)
			return false;	/* -000000 is not allowed */
		y = -y;
	}
	int mo = 1, dd = 1, hh = 0, mi = 0, ss = 0, msec = 0;
	if (i < n && p[i] == '-') {
		if (n - i < 3 || !dd2(p + i + 1, mo))
			return false;
		i += 3;
		if (i < n && p[i] == '-') {
			if (n - i < 3 || !dd2(p + i + 1, dd))
				return false;
			i += 3;
		}
	}
	int64_t off = 0;
	if (i < n && p[i] == 'T') {
		if (n - i < 6 || !dd2(p + i + 1, hh) || p[i + 3] != ':' ||
		    !dd2(p + i + 4, mi))
			return false;
		i += 6;
		if (i < n && p[i] == ':') {
			if (n - i < 3 || !dd2(p + i + 1, ss))
				return false;
			i += 3;
			if (i < n && p[i] == '.') {
				i++;
				int nd = 0;
				while (i < n && p[i] >= '0' && p[i] <= '9') {
					if (nd < 3)
						msec = msec * 10 + (p[i] - '0');
					nd++;
					i++;
				}
				if (nd == 0)
					return false;
				for (; nd < 3; nd++)
					msec *= 10;
			}
		}
		if (i < n && p[i] == 'Z') {
			i++;
		} else if (i < n && (p[i] == '+' || p[i] == '-')) {
			int oh, om;
			if (n - i < 6 || !dd2(p + i + 1, oh) ||
			    p[i + 3] != ':' || !dd2(p + i + 4, om))
				return false;
			if (oh > 23 || om > 59)
				return false;
			off = (int64_t)(oh * 60 + om) * 60000;
			if (p[i] == '-')
				off = -off;
			i += 6;
		}
	}
	if (i != n)
		return false;
	if (mo < 1 || mo > 12 || dd < 1 || dd > 31)
		return false;
	if (hh > 24 || mi > 59 || ss > 59)
		return false;
	if (hh == 24 && (mi || ss || msec))
		return false;
	int64_t t = (int64_t)days_from_civil(y, mo, dd) * 86400000ll +
	    (int64_t)(((hh * 60 + mi) * 60 + ss) * 1000 + msec) - off;
	if (t > 8640000000000000ll || t < -8640000000000000ll)
		return false;
	*ms = t;
	return true;
}

/*
 * Could V8's legacy parser make a date of this (non-ISO) string?  Not decided
 * here -- only whether it LOOKS like one of the forms that parser exists for:
 * a month name, two date separators between digits (2014/05/01, 5-1-2014,
 * 2014-13-45), or a clock time (12:00).  Everything else is called NaN.
 */
DNG_HD bool dng_date_maybe_legacy(const uint8_t *p, int n)
{
	int seps = 0;
	for (int i = 0; i + 2 < n; i++) {
		const bool d0 = p[i] >= '0' && p[i] <= '9';
		const bool d2 = p[i + 2] >= '0' && p[i + 2] <= '9';
		if (d0 && d2 && (p[i + 1] == '-' || p[i + 1] == '/'))
			seps++;
		if (d0 && d2 && p[i + 1] == ':')
			return true;
	}
	if (seps >= 2)
		return true;
	for (int i = 0; i + 2 < n; i++) {
		const uint32_t a = p[i] | 0x20, b = p[i + 1] | 0x20,
		    c = p[i + 2] | 0x20;
		if (i > 0 && (uint32_t)((p[i - 1] | 0x20) - 'a') < 26u)
			continue;		/* not the start of a word */
		const uint32_t w = a << 16 | b << 8 | c;
		if (w == 0x6a616e || w == 0x666562 || w == 0x6d6172 ||
		    w == 0x617072 || w == 0x6d6179 || w == 0x6a756e ||
		    w == 0x6a756c || w == 0x617567 || w == 0x736570 ||
		    w == 0x6f6374 || w == 0x6e6f76 || w == 0x646563)
			return true;
	}
	return false;
}

} /* namespace dng */
#endif
