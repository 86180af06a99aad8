/*
 * fscan.cuh: the F path's wildcard scanners: a JSON string body up to its
 * closing quote, and a bare scalar (number / true / false / null).  What they
 * accept is exactly what JSON.parse accepts at that place
 * (lib/format-json.js:34); what they store is a capture (fast.h DNG_FCAP).
 *
 * Shared by the interpreted matcher (fast.cuh fmatch) and the matchers
 * scan_kernel_j gets from the run-time compiler (jit.cpp): this file is also
 * embedded as text in the library and handed to NVRTC in front of the generated
 * code, so it includes nothing.  The includer supplies u32, DNG_HD, the T_*
 * value types, DNG_FCAP and the byte helpers is_hex / tm_isdigit /
 * nondigit_mask / low_flag_byte (record.cuh + tmpl.cuh here, jit.cpp's prelude
 * there).
 *
 * M supplies the record: cursor(off).next() = successive little-endian words
 * from byte `off` on, byte(off), word(off); a '\n' follows the record's last
 * byte and ends every scan.
 */
#ifndef DNG_FSCAN_CUH
#define DNG_FSCAN_CUH

/* 0x80 in (at least) the lowest byte of w that ends a plain run of string
 * body: '"', '\\' or a control byte */
DNG_HD u32 fstr_stop(u32 w)
{
	const u32 x1 = w ^ 0x22222222u;
	const u32 x2 = w ^ 0x5c5c5c5cu;
	const u32 x3 = w & 0xe0e0e0e0u;
	return (((x1 - 0x01010101u) & ~x1) | ((x2 - 0x01010101u) & ~x2) |
	    ((x3 - 0x01010101u) & ~x3)) & 0x80808080u;
}

/*
 * A string body from q (the byte after the opening quote).  True: q = its
 * closing quote, val = the capture (flag set if the body holds an escape).
 * Escapes are validated (\" \\ \/ \b \f \n \r \t \uXXXX) and stepped over.
 */
template <class M>
DNG_HD bool fscan_str(M &m, u32 &q, u32 &val)
{
	u32 e = q, esc = 0;
	bool ok;
	for (;;) {
		typename M::Cur c = m.cursor(e);
		u32 hit, w;
		/* eight bytes a round */
#pragma unroll 1
		for (;;) {
			w = c.next();
			hit = fstr_stop(w);
			const u32 w2 = c.next();
			const u32 hit2 = fstr_stop(w2);
			if (hit)
				break;
			e += 4;
			w = w2;
			hit = hit2;
			if (hit)
				break;
			e += 4;
		}
		const u32 b = low_flag_byte(hit);
		e += b;
		const u32 stop = (w >> (8 * b)) & 0xff;
		if (stop != '\\') {
			ok = stop == '"';
			break;
		}
		const u32 c1 = m.byte(e + 1);
		if (c1 == 'u') {
			ok = is_hex(m.byte(e + 2)) && is_hex(m.byte(e + 3)) &&
			    is_hex(m.byte(e + 4)) && is_hex(m.byte(e + 5));
			e += 6;
		} else {
			ok = c1 == '"' || c1 == '\\' || c1 == '/' || c1 == 'b' ||
			    c1 == 'f' || c1 == 'n' || c1 == 'r' || c1 == 't';
			e += 2;
		}
		esc = 1;
		if (!ok)
			break;
	}
	val = DNG_FCAP(T_STR, q, e - q, esc);
	q = e;
	return ok;
}

/*
 * A bare scalar at q.  True: q = the byte after it, val = the capture (for a
 * number, flag set if it is [-]digits with at most 15 digits and not "-0":
 * what the stages use without a conversion).
 */
template <class M>
DNG_HD bool fscan_bare(M &m, u32 &q, u32 &val)
{
	typename M::Cur c = m.cursor(q);
	u32 w = c.next();
	const u32 c0 = w & 0xff;
	bool ok;
	if (c0 - '0' > 9u && c0 != '-') {
		if (c0 == 't') {
			ok = w == 0x65757274u;
			val = DNG_FCAP(T_TRUE, q, 4, 0);
			q += 4;
		} else if (c0 == 'n') {
			ok = w == 0x6c6c756eu;
			val = DNG_FCAP(T_NULL, q, 4, 0);
			q += 4;
		} else {
			ok = w == 0x736c6166u && (c.next() & 0xff) == 'e';
			val = DNG_FCAP(T_FALSE, q, 5, 0);
			q += 5;
		}
		return ok;
	}
	/* -?(0|[1-9][0-9]*) word-wise; a fraction or an exponent continues
	 * byte-wise */
	const u32 neg = c0 == '-';
	u32 i = q + neg;
	if (neg) {
		c = m.cursor(i);
		w = c.next();
	}
	const u32 d0 = w & 0xff;
	u32 nd = 0, mk;
#pragma unroll 1
	while ((mk = nondigit_mask(w)) == 0) {
		nd += 4;
		w = c.next();
	}
	const u32 lb = low_flag_byte(mk);
	nd += lb;
	ok = nd > 0 && !(d0 == '0' && nd > 1);
	i += nd;
	u32 simple = 1;
	if (nd > 15 || (neg && nd == 1 && d0 == '0'))
		simple = 0;
	u32 ch = (w >> (8 * lb)) & 0xff;
	if (ok && (ch == '.' || (ch | 0x20) == 'e')) {
		simple = 0;
		if (ch == '.') {
			ch = m.byte(++i);
			ok = tm_isdigit(ch);
			while (tm_isdigit(ch))
				ch = m.byte(++i);
		}
		if (ok && (ch | 0x20) == 'e') {
			ch = m.byte(++i);
			if (ch == '+' || ch == '-')
				ch = m.byte(++i);
			ok = tm_isdigit(ch);
			while (tm_isdigit(ch))
				ch = m.byte(++i);
		}
	}
	val = DNG_FCAP(T_NUM, q, i - q, simple);
	q = i;
	return ok;
}

#endif
