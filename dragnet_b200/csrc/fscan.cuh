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
 * M supplies the record:
 *   byte(off), word(off)   one byte / four bytes (little endian) at any offset
 *   acursor(off)           ALIGNED words: c.next() returns, first, the word
 *                          that holds byte `off` -- c.k (0..3) of its low
 *                          bytes come before `off` -- then the words after it;
 *                          apos(c) = the record offset of byte 0 of the word
 *                          next() returned last
 * The scans run on aligned words (one load per word, nothing to shift) and
 * take what precedes `off` out of the first word with a mask.
 *
 * What follows the record's last byte: a '\n', and within the bytes readable
 * after it a '"'.  The '\n' ends a number; a string body is searched for its
 * quote or backslash only and the control bytes on the way are collected and
 * looked at once at the end, so a string the record does not close runs on
 * to the next quote (the next record's, or the sentinel's) and fails there on
 * the '\n' it has crossed.
 */
#ifndef DNG_FSCAN_CUH
#define DNG_FSCAN_CUH

/* how the two scanners are compiled: inlined, unless the includer says
 * otherwise (the run-time compiled matcher shares one copy of each among its
 * blocks: the kernel's loop has to fit the instruction cache) */
#ifndef DNG_FSCAN_FN
#define DNG_FSCAN_FN DNG_HD
#endif

/* what a scanner returns: ok << 63 | where the scan ended << 32 | the capture
 * (by value, in two registers: the shared copies are called) */
DNG_HD unsigned long long fscan_pack(bool ok, u32 end, u32 val)
{
	return ((unsigned long long)(end | (ok ? 0x80000000u : 0u)) << 32) | val;
}

/* bit 7 of (at least) the lowest byte of w that is '"' or '\\'; flags above it
 * may be spurious, flags never go missing */
DNG_HD u32 fstr_qb(u32 w)
{
	const u32 x1 = w ^ 0x22222222u;
	const u32 x2 = w ^ 0x5c5c5c5cu;
	return ((x1 - 0x01010101u) & ~x1) | ((x2 - 0x01010101u) & ~x2);
}

/* bit 7 of the lowest control byte (< 0x20) of w, likewise */
DNG_HD u32 fstr_ctl(u32 w)
{
	return (w - 0x20202020u) & ~w;
}

/*
 * A string body from q (the byte after the opening quote).  Ok: ends at its
 * closing quote, the capture has its flag set if the body holds an escape.
 * Escapes are validated (\" \\ \/ \b \f \n \r \t \uXXXX) and stepped over.
 */
template <class M>
DNG_FSCAN_FN unsigned long long fscan_str_p(M m, const u32 q)
{
	u32 e = q, esc = 0, ctl = 0;
	bool ok;
	for (;;) {
		typename M::ACur c = m.acursor(e);
		/* (what precedes e reads as 0xff: no quote, no control) */
		u32 w = c.next() | ~(0xffffffffu << (8 * c.k));
		u32 w2, h, h2;
		/* eight bytes a round */
#pragma unroll 1
		for (;;) {
			w2 = c.next();
			h = fstr_qb(w);
			h2 = fstr_qb(w2);
			if ((h | h2) & 0x80808080u)
				break;
			ctl |= fstr_ctl(w) | fstr_ctl(w2);
			w = c.next();
		}
		u32 at = m.apos(c);		/* of w2 */
		if (h & 0x80808080u) {
			at -= 4;
		} else {
			ctl |= fstr_ctl(w);
			w = w2;
			h = h2;
		}
		const u32 b = low_flag_byte(h & 0x80808080u);
		/* the control bytes in front of the stop */
		ctl |= fstr_ctl(w) & ~(0xffffffffu << (8 * b));
		e = at + b;
		const u32 stop = (w >> (8 * b)) & 0xff;
		if (stop != '\\') {
			ok = true;		/* the closing quote */
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
	return fscan_pack(ok && !(ctl & 0x80808080u), e,
	    DNG_FCAP(T_STR, q, e - q, esc));
}

template <class M>
DNG_HD bool fscan_str(M &m, u32 &q, u32 &val)
{
	const unsigned long long r = fscan_str_p(m, q);
	val = (u32)r;
	q = (u32)(r >> 32) & 0x7fffffffu;
	return (r >> 63) != 0;
}

/*
 * A bare scalar at q.  True: q = the byte after it, val = the capture (for a
 * number, flag set if it is [-]digits with at most 15 digits and not "-0":
 * what the stages use without a conversion).
 */
template <class M>
DNG_FSCAN_FN unsigned long long fscan_bare_p(M m, const u32 q)
{
	const u32 c0 = m.byte(q);
	bool ok;
	if (c0 - '0' > 9u && c0 != '-') {
		const u32 w = m.word(q);
		if (c0 == 't')
			return fscan_pack(w == 0x65757274u, q + 4,
			    DNG_FCAP(T_TRUE, q, 4, 0));
		if (c0 == 'n')
			return fscan_pack(w == 0x6c6c756eu, q + 4,
			    DNG_FCAP(T_NULL, q, 4, 0));
		return fscan_pack(w == 0x736c6166u && m.byte(q + 4) == 'e',
		    q + 5, DNG_FCAP(T_FALSE, q, 5, 0));
	}
	/* -?(0|[1-9][0-9]*) word-wise; a fraction or an exponent continues
	 * byte-wise */
	const u32 neg = c0 == '-';
	u32 i = q + neg;
	const u32 d0 = neg ? m.byte(i) : c0;
	typename M::ACur c = m.acursor(i);
	/* (what precedes i reads as digits) */
	const u32 keep = 0xffffffffu << (8 * c.k);
	u32 w = (c.next() & keep) | (0x30303030u & ~keep);
	u32 mk;
#pragma unroll 1
	while ((mk = nondigit_mask(w)) == 0)
		w = c.next();
	const u32 lb = low_flag_byte(mk);
	const u32 stop = m.apos(c) + lb;	/* the first non-digit */
	const u32 nd = stop - i;
	ok = nd > 0 && !(d0 == '0' && nd > 1);
	i = stop;
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
	return fscan_pack(ok, i, DNG_FCAP(T_NUM, q, i - q, simple));
}

template <class M>
DNG_HD bool fscan_bare(M &m, u32 &q, u32 &val)
{
	const unsigned long long r = fscan_bare_p(m, q);
	val = (u32)r;
	q = (u32)(r >> 32) & 0x7fffffffu;
	return (r >> 63) != 0;
}

#endif
