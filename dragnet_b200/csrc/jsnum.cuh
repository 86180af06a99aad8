/*
 * jsnum.cuh: ECMAScript number <-> text conversions, exact, table-free, usable
 * from device code (record slow paths) and host code (plan compiler).
 *
 * The reference relies on V8 for these: JSON.parse of number literals
 * (lib/format-json.js:34 via vstream-json-parser), String(number) for group
 * keys (skinner aggregator, lib/dragnet-impl.js:48-51), ToNumber(string) in
 * krill's loose comparisons and in the bucketizers (lib/dragnet.js:61-70;
 * pinned by the golden `"latency":"26"` case, tst.scan_fileset.sh.out:292-318).
 *
 *   dng_parse_decimal()      correctly rounded decimal -> binary64
 *                            (Clinger fast path, else exact big-integer scaling)
 *   dng_number_to_string()   Number::toString (ECMA-262 7.1.12.1): shortest
 *                            round-trip digits (Burger & Dybvig free-format
 *                            with big integers), JS layout rules
 *   dng_string_to_number()   StringToNumber (ECMA-262 7.1.3.1, ES2015 grammar)
 */
#ifndef DNG_JSNUM_CUH
#define DNG_JSNUM_CUH

#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define DNG_HD __host__ __device__ __forceinline__
#define DNG_HDI __host__ __device__ inline
#define DNG_HDN __host__ __device__ __noinline__ inline
#else
#define DNG_HD inline
#define DNG_HDI inline
#define DNG_HDN inline
#endif

namespace dng {

enum { BIGW = 128 };

struct Big {
	uint32_t w[BIGW];
	int n;			/* used words (no leading zero words) */
	int overflow;
};

DNG_HDN void big_set(Big &b, uint64_t v)
{
	b.n = 0;
	b.overflow = 0;
	if (v) {
		b.w[b.n++] = (uint32_t)v;
		if (v >> 32)
			b.w[b.n++] = (uint32_t)(v >> 32);
	}
}

DNG_HDN void big_mul_small(Big &b, uint32_t m, uint32_t add)
{
	uint64_t carry = add;
	for (int i = 0; i < b.n; i++) {
		uint64_t t = (uint64_t)b.w[i] * m + carry;
		b.w[i] = (uint32_t)t;
		carry = t >> 32;
	}
	if (carry) {
		if (b.n < BIGW)
			b.w[b.n++] = (uint32_t)carry;
		else
			b.overflow = 1;
	}
}

DNG_HDN void big_mul_pow10(Big &b, int e)
{
	while (e >= 9) {
		big_mul_small(b, 1000000000u, 0);
		e -= 9;
	}
	uint32_t m = 1;
	while (e-- > 0)
		m *= 10;
	if (m > 1)
		big_mul_small(b, m, 0);
}

/* b /= d; returns remainder */
DNG_HDN uint32_t big_div_small(Big &b, uint32_t d)
{
	uint64_t rem = 0;
	for (int i = b.n - 1; i >= 0; i--) {
		uint64_t cur = (rem << 32) | b.w[i];
		b.w[i] = (uint32_t)(cur / d);
		rem = cur % d;
	}
	while (b.n > 0 && b.w[b.n - 1] == 0)
		b.n--;
	return (uint32_t)rem;
}

DNG_HDN void big_shl(Big &b, int k)
{
	if (b.n == 0 || k == 0)
		return;
	int ws = k >> 5, bs = k & 31;
	if (b.n + ws + 1 > BIGW) {
		b.overflow = 1;
		return;
	}
	if (bs == 0) {
		for (int i = b.n - 1; i >= 0; i--)
			b.w[i + ws] = b.w[i];
	} else {
		b.w[b.n + ws] = 0;
		for (int i = b.n - 1; i >= 0; i--) {
			b.w[i + ws + 1] |= b.w[i] >> (32 - bs);
			b.w[i + ws] = b.w[i] << bs;
		}
	}
	for (int i = 0; i < ws; i++)
		b.w[i] = 0;
	b.n += ws + (bs ? 1 : 0);
	while (b.n > 0 && b.w[b.n - 1] == 0)
		b.n--;
}

DNG_HDN int big_bitlen(const Big &b)
{
	if (b.n == 0)
		return 0;
	uint32_t top = b.w[b.n - 1];
	int l = 0;
	while (top) {
		l++;
		top >>= 1;
	}
	return (b.n - 1) * 32 + l;
}

DNG_HDN int big_cmp(const Big &a, const Big &b)
{
	if (a.n != b.n)
		return a.n < b.n ? -1 : 1;
	for (int i = a.n - 1; i >= 0; i--) {
		if (a.w[i] != b.w[i])
			return a.w[i] < b.w[i] ? -1 : 1;
	}
	return 0;
}

/* a -= b (requires a >= b) */
DNG_HDN void big_sub(Big &a, const Big &b)
{
	int64_t borrow = 0;
	for (int i = 0; i < a.n; i++) {
		int64_t t = (int64_t)a.w[i] - (i < b.n ? b.w[i] : 0) - borrow;
		borrow = t < 0;
		a.w[i] = (uint32_t)t;
	}
	while (a.n > 0 && a.w[a.n - 1] == 0)
		a.n--;
}

/* cmp(a + b, c) without materialising the sum beyond a temp */
DNG_HDN int big_cmp_sum(const Big &a, const Big &b, const Big &c, Big &tmp)
{
	int n = a.n > b.n ? a.n : b.n;
	uint64_t carry = 0;
	for (int i = 0; i < n; i++) {
		uint64_t t = carry + (i < a.n ? a.w[i] : 0) +
		    (i < b.n ? b.w[i] : 0);
		tmp.w[i] = (uint32_t)t;
		carry = t >> 32;
	}
	tmp.n = n;
	if (carry && n < BIGW)
		tmp.w[tmp.n++] = (uint32_t)carry;
	return big_cmp(tmp, c);
}

DNG_HD double bits_to_double(uint64_t b)
{
	union { uint64_t u; double d; } x;
	x.u = b;
	return x.d;
}

DNG_HD uint64_t double_to_bits(double d)
{
	union { uint64_t u; double d; } x;
	x.d = d;
	return x.u;
}

/*
 * Round (mant64 * 2^e2), mant64 normalised (bit 63 set), plus a sticky bit
 * for discarded lower-order value, to the nearest binary64 (ties to even).
 */
DNG_HDN double round_to_double(uint64_t mant64, int e2, int sticky)
{
	int E = e2 + 63;
	if (E > 1023)
		return bits_to_double(0x7ff0000000000000ull);
	int s = 11;
	if (E < -1022)
		s += -1022 - E;
	uint64_t m, rem, half;
	if (s >= 65)
		return 0.0;
	if (s == 64) {
		m = 0;
		rem = mant64;
		half = 1ull << 63;
	} else {
		m = mant64 >> s;
		rem = mant64 & ((1ull << s) - 1);
		half = 1ull << (s - 1);
	}
	if (rem > half || (rem == half && (sticky || (m & 1))))
		m++;
	if (E < -1022)
		return bits_to_double(m);	/* denormal (or min normal) */
	if (m >> 53) {
		m >>= 1;
		E++;
		if (E > 1023)
			return bits_to_double(0x7ff0000000000000ull);
	}
	return bits_to_double(((uint64_t)(E + 1023) << 52) |
	    (m & 0xfffffffffffffull));
}

/*
 * Slow, exact decimal -> double.  p[0..len) is [+-]? digits [. digits]
 * [(e|E) [+-] digits] with at least one digit in the mantissa (".5" and "5."
 * are accepted); the caller has validated the syntax.
 */
DNG_HDN double dng_parse_decimal_slow(const uint8_t *p, int len)
{
	int i = 0, neg = 0;
	if (i < len && (p[i] == '-' || p[i] == '+')) {
		neg = p[i] == '-';
		i++;
	}
	Big W;
	big_set(W, 0);
	int ndig = 0;		/* significant digits accumulated */
	int q = 0;		/* decimal exponent adjustment */
	int sticky = 0;
	int seen_dot = 0, seen_nz = 0;
	uint32_t chunk = 0, cm = 1;
	const int MAXD = 780;
	for (; i < len; i++) {
		uint8_t c = p[i];
		if (c == '.') {
			seen_dot = 1;
			continue;
		}
		if (c < '0' || c > '9')
			break;
		if (c != '0')
			seen_nz = 1;
		if (!seen_nz) {		/* leading zero */
			if (seen_dot)
				q--;
			continue;
		}
		if (ndig < MAXD) {
			chunk = chunk * 10 + (c - '0');
			cm *= 10;
			ndig++;
			if (cm == 1000000000u) {
				big_mul_small(W, cm, chunk);
				if (W.n == 0 && chunk)
					big_set(W, chunk);
				chunk = 0;
				cm = 1;
			}
			if (seen_dot)
				q--;
		} else {
			if (c != '0')
				sticky = 1;
			if (!seen_dot)
				q++;
		}
	}
	if (cm > 1) {
		if (W.n == 0)
			big_set(W, chunk);
		else
			big_mul_small(W, cm, chunk);
	}
	if (i < len && (p[i] == 'e' || p[i] == 'E')) {
		i++;
		int eneg = 0;
		if (i < len && (p[i] == '-' || p[i] == '+')) {
			eneg = p[i] == '-';
			i++;
		}
		int ex = 0;
		for (; i < len && p[i] >= '0' && p[i] <= '9'; i++) {
			if (ex < 100000)
				ex = ex * 10 + (p[i] - '0');
		}
		q += eneg ? -ex : ex;
	}
	double r;
	if (W.n == 0) {
		r = 0.0;
	} else {
		int mag = ndig + q;
		if (mag > 310) {
			r = bits_to_double(0x7ff0000000000000ull);
		} else if (mag < -330) {
			r = 0.0;
		} else {
			int binexp = 0;
			if (q >= 0) {
				big_mul_pow10(W, q);
			} else {
				int m = -q;
				int k = (m * 3322 + 999) / 1000 + 68 -
				    big_bitlen(W);
				if (k < 0)
					k = 0;
				big_shl(W, k);
				binexp = -k;
				while (m >= 9) {
					if (big_div_small(W, 1000000000u))
						sticky = 1;
					m -= 9;
				}
				uint32_t d = 1;
				while (m-- > 0)
					d *= 10;
				if (d > 1 && big_div_small(W, d))
					sticky = 1;
			}
			int L = big_bitlen(W);
			uint64_t mant = 0;
			int shift = L - 64;
			if (shift <= 0) {
				uint64_t v = 0;
				for (int j = W.n - 1; j >= 0; j--)
					v = (v << 32) | W.w[j];
				mant = v << (-shift);
			} else {
				/* top 64 bits of W */
				int ws = shift >> 5, bs = shift & 31;
				uint64_t lo = W.w[ws];
				uint64_t mid = ws + 1 < W.n ? W.w[ws + 1] : 0;
				uint64_t hi = ws + 2 < W.n ? W.w[ws + 2] : 0;
				if (bs == 0)
					mant = lo | (mid << 32);
				else
					mant = (lo >> bs) | (mid << (32 - bs)) |
					    (hi << (64 - bs));
				if (bs && (W.w[ws] & ((1u << bs) - 1)))
					sticky = 1;
				for (int j = 0; j < ws && !sticky; j++)
					if (W.w[j])
						sticky = 1;
			}
			r = round_to_double(mant, binexp + shift, sticky);
		}
	}
	return neg ? -r : r;
}

/* decimal -> double; fast path for <= 19 digits and small exponents */
DNG_HDN double dng_parse_decimal(const uint8_t *p, int len)
{
	int i = 0, neg = 0;
	if (i < len && (p[i] == '-' || p[i] == '+')) {
		neg = p[i] == '-';
		i++;
	}
	uint64_t w = 0;
	int nd = 0, q = 0, seen_dot = 0, ok = 1;
	for (; i < len; i++) {
		uint8_t c = p[i];
		if (c == '.') {
			seen_dot = 1;
			continue;
		}
		if (c < '0' || c > '9')
			break;
		if (w == 0 && c == '0') {
			if (seen_dot)
				q--;
			continue;
		}
		if (nd >= 19) {
			ok = 0;
			break;
		}
		w = w * 10 + (c - '0');
		nd++;
		if (seen_dot)
			q--;
	}
	if (ok && i < len) {		/* exponent */
		i++;
		int eneg = 0;
		if (i < len && (p[i] == '-' || p[i] == '+')) {
			eneg = p[i] == '-';
			i++;
		}
		int ex = 0;
		for (; i < len; i++) {
			if (ex < 10000)
				ex = ex * 10 + (p[i] - '0');
		}
		q += eneg ? -ex : ex;
	}
	if (ok && w <= (1ull << 53) && q >= -22 && q <= 22) {
		const double p10[] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7,
		    1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17,
		    1e18, 1e19, 1e20, 1e21, 1e22 };
		double d = (double)w;
		if (q < 0)
			d = d / p10[-q];
		else
			d = d * p10[q];
		return neg ? -d : d;
	}
	return dng_parse_decimal_slow(p, len);
}

/*
 * Shortest round-trip digits of a finite positive double.
 * digits[0..*nd) are '0'..'9', value = 0.d1d2... * 10^(*n).
 */
DNG_HDN void dng_shortest_digits(double v, char *digits, int *nd, int *n)
{
	uint64_t bits = double_to_bits(v);
	int be = (int)((bits >> 52) & 0x7ff);
	uint64_t f = bits & 0xfffffffffffffull;
	int e;
	if (be == 0) {
		e = -1074;
	} else {
		f |= 1ull << 52;
		e = be - 1075;
	}
	int even = !(f & 1);
	int flen = 0;
	for (uint64_t t = f; t; t >>= 1)
		flen++;
	Big r, s, mp, mm, tmp;
	int lower_closer = (f == (1ull << 52)) && be > 1;
	if (e >= 0) {
		big_set(r, f);
		big_shl(r, e + 1 + (lower_closer ? 1 : 0));
		big_set(s, lower_closer ? 4 : 2);
		big_set(mp, 1);
		big_shl(mp, e + (lower_closer ? 1 : 0));
		big_set(mm, 1);
		big_shl(mm, e);
	} else {
		big_set(r, f);
		big_shl(r, lower_closer ? 2 : 1);
		big_set(s, 1);
		big_shl(s, -e + 1 + (lower_closer ? 1 : 0));
		big_set(mp, lower_closer ? 2 : 1);
		big_set(mm, 1);
	}
	int k = (int)ceil((double)(e + flen - 1) * 0.30102999566398114 - 1e-10);
	if (k >= 0) {
		big_mul_pow10(s, k);
	} else {
		big_mul_pow10(r, -k);
		big_mul_pow10(mp, -k);
		big_mul_pow10(mm, -k);
	}
	int c = big_cmp_sum(r, mp, s, tmp);
	if (even ? c >= 0 : c > 0) {
		k++;
	} else {
		big_mul_small(r, 10, 0);
		big_mul_small(mp, 10, 0);
		big_mul_small(mm, 10, 0);
	}
	int count = 0;
	for (;;) {
		int d = 0;
		while (big_cmp(r, s) >= 0) {
			big_sub(r, s);
			d++;
		}
		int c1 = big_cmp(r, mm);
		int tc1 = even ? c1 <= 0 : c1 < 0;
		int c2 = big_cmp_sum(r, mp, s, tmp);
		int tc2 = even ? c2 >= 0 : c2 > 0;
		if (!tc1 && !tc2) {
			digits[count++] = (char)('0' + d);
			big_mul_small(r, 10, 0);
			big_mul_small(mp, 10, 0);
			big_mul_small(mm, 10, 0);
			if (count >= 24)
				break;
			continue;
		}
		if (tc1 && !tc2) {
			digits[count++] = (char)('0' + d);
		} else if (!tc1 && tc2) {
			digits[count++] = (char)('0' + d + 1);
		} else {
			/* both: pick the closer; an exact tie takes the even
			 * digit (ECMA-262 7.1.12.1 step 5 note) */
			big_shl(r, 1);
			int cc = big_cmp(r, s);
			int up = cc > 0 || (cc == 0 && (d & 1));
			digits[count++] = (char)('0' + (up ? d + 1 : d));
		}
		break;
	}
	/* a generated '9'+1 cannot occur: d+1 <= 9 is guaranteed by the
	 * termination conditions of the free-format algorithm */
	*nd = count;
	*n = k;
}

/* Number::toString(v) into out (>= 32 bytes); returns length */
DNG_HDN int dng_number_to_string(double v, char *out)
{
	int o = 0;
	if (v != v) {
		out[0] = 'N'; out[1] = 'a'; out[2] = 'N';
		return 3;
	}
	if (v == 0.0) {
		out[0] = '0';
		return 1;
	}
	if (v < 0) {
		out[o++] = '-';
		v = -v;
	}
	if (v > 1.7976931348623157e308) {
		const char *s = "Infinity";
		for (int i = 0; i < 8; i++)
			out[o++] = s[i];
		return o;
	}
	if (v < 9007199254740992.0 && v == floor(v)) {
		/* exact integer below 2^53: its decimal expansion is the
		 * shortest round-trip form */
		uint64_t u = (uint64_t)v;
		char tmp[20];
		int t = 0;
		while (u) {
			tmp[t++] = (char)('0' + u % 10);
			u /= 10;
		}
		while (t)
			out[o++] = tmp[--t];
		return o;
	}
	char dg[28];
	int k, n;
	dng_shortest_digits(v, dg, &k, &n);
	if (k <= n && n <= 21) {
		for (int i = 0; i < k; i++)
			out[o++] = dg[i];
		for (int i = k; i < n; i++)
			out[o++] = '0';
	} else if (0 < n && n <= 21) {
		for (int i = 0; i < n; i++)
			out[o++] = dg[i];
		out[o++] = '.';
		for (int i = n; i < k; i++)
			out[o++] = dg[i];
	} else if (-6 < n && n <= 0) {
		out[o++] = '0';
		out[o++] = '.';
		for (int i = 0; i < -n; i++)
			out[o++] = '0';
		for (int i = 0; i < k; i++)
			out[o++] = dg[i];
	} else {
		int e = n - 1;
		out[o++] = dg[0];
		if (k > 1) {
			out[o++] = '.';
			for (int i = 1; i < k; i++)
				out[o++] = dg[i];
		}
		out[o++] = 'e';
		out[o++] = e >= 0 ? '+' : '-';
		if (e < 0)
			e = -e;
		char tmp[6];
		int t = 0;
		do {
			tmp[t++] = (char)('0' + e % 10);
			e /= 10;
		} while (e);
		while (t)
			out[o++] = tmp[--t];
	}
	return o;
}

/* length in bytes of a JS WhiteSpace/LineTerminator char at p, else 0 */
DNG_HD int js_space_len(const uint8_t *p, int len)
{
	if (len <= 0)
		return 0;
	uint8_t c = p[0];
	if (c == ' ' || (c >= 9 && c <= 13))
		return 1;
	if (c == 0xC2 && len >= 2 && p[1] == 0xA0)
		return 2;
	if (len >= 3) {
		uint32_t cp = 0;
		if ((c & 0xF0) == 0xE0 && (p[1] & 0xC0) == 0x80 &&
		    (p[2] & 0xC0) == 0x80)
			cp = ((c & 0x0F) << 12) | ((p[1] & 0x3F) << 6) |
			    (p[2] & 0x3F);
		if (cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
		    cp == 0x2028 || cp == 0x2029 || cp == 0x202F ||
		    cp == 0x205F || cp == 0x3000 || cp == 0xFEFF)
			return 3;
	}
	return 0;
}

DNG_HD double dng_nan()
{
	return bits_to_double(0x7ff8000000000000ull);
}

/* ToNumber(string): p[0..len) is the string's UTF-8 */
DNG_HDN double dng_string_to_number(const uint8_t *p, int len)
{
	int a = 0, b = len, l;
	while (a < b && (l = js_space_len(p + a, b - a)) > 0)
		a += l;
	for (;;) {
		/* trailing whitespace: step back over 1-3 byte forms */
		int done = 1;
		for (int w = 1; w <= 3 && b - w >= a; w++) {
			if (js_space_len(p + b - w, w) == w) {
				b -= w;
				done = 0;
				break;
			}
		}
		if (done)
			break;
	}
	if (a == b)
		return 0.0;
	const uint8_t *s = p + a;
	int n = b - a;
	if (n > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X' ||
	    s[1] == 'o' || s[1] == 'O' || s[1] == 'b' || s[1] == 'B')) {
		int base = (s[1] | 0x20) == 'x' ? 16 :
		    (s[1] | 0x20) == 'o' ? 8 : 2;
		int bpd = base == 16 ? 4 : base == 8 ? 3 : 1;
		uint64_t m = 0;
		int extra = 0, sticky = 0;
		for (int i = 2; i < n; i++) {
			int c = s[i], d;
			if (c >= '0' && c <= '9')
				d = c - '0';
			else if ((c | 0x20) >= 'a' && (c | 0x20) <= 'f')
				d = (c | 0x20) - 'a' + 10;
			else
				return dng_nan();
			if (d >= base)
				return dng_nan();
			for (int bit = bpd - 1; bit >= 0; bit--) {
				int v = (d >> bit) & 1;
				if (m >> 63) {
					extra++;
					if (v)
						sticky = 1;
				} else {
					m = (m << 1) | v;
				}
			}
		}
		if (m == 0)
			return 0.0;
		int sh = 0;
		while (!(m >> 63)) {
			m <<= 1;
			sh++;
		}
		return round_to_double(m, extra - sh, sticky);
	}
	int i = 0;
	if (s[i] == '+' || s[i] == '-')
		i++;
	if (n - i == 8 && s[i] == 'I' && s[i+1] == 'n' && s[i+2] == 'f' &&
	    s[i+3] == 'i' && s[i+4] == 'n' && s[i+5] == 'i' &&
	    s[i+6] == 't' && s[i+7] == 'y') {
		double inf = bits_to_double(0x7ff0000000000000ull);
		return s[0] == '-' ? -inf : inf;
	}
	int nd = 0;
	while (i < n && s[i] >= '0' && s[i] <= '9') {
		i++;
		nd++;
	}
	if (i < n && s[i] == '.') {
		i++;
		while (i < n && s[i] >= '0' && s[i] <= '9') {
			i++;
			nd++;
		}
	}
	if (nd == 0)
		return dng_nan();
	if (i < n && (s[i] == 'e' || s[i] == 'E')) {
		i++;
		if (i < n && (s[i] == '+' || s[i] == '-'))
			i++;
		int ed = 0;
		while (i < n && s[i] >= '0' && s[i] <= '9') {
			i++;
			ed++;
		}
		if (ed == 0)
			return dng_nan();
	}
	if (i != n)
		return dng_nan();
	return dng_parse_decimal(s, n);
}

} /* namespace dng */
#endif
