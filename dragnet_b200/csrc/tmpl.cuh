/*
 * tmpl.cuh: match one record against the scan's template trie (tmpl.h).
 *
 * One thread walks the trie for its own record: compare the node's literal
 * four bytes at a time, scan the wildcard that follows (a string body up to
 * its closing quote, or a bare scalar), store the value if the plan wants it,
 * move on.  A literal that differs or a malformed wildcard tries the node's
 * next sibling; there is no deeper backtracking, and any failure just means
 * "not templated": the caller parses the record with the byte automaton
 * instead, so this function only ever has to be right when it says yes.
 *
 * The values it stores are exactly what fast_finish() (record.cuh) makes of
 * the automaton's captures: strings as (T_STR, body span, VF_ESCAPED if the
 * body holds a backslash),
 * numbers with VF_SIMPLEINT decided by the same rule, containers as
 * (T_OBJ|T_ARR, start).
 *
 * M supplies memory access: the record (a '\n' must follow its last byte; the
 * scans rely on it as a terminator) and the trie blob.
 */
#ifndef DNG_TMPL_CUH
#define DNG_TMPL_CUH

#include "tmpl.h"
#include "record.cuh"

/* unroll factors of the matcher's two inner loops (measured: see DESIGN.md) */
#ifndef DNG_LIT_UNROLL
#define DNG_LIT_UNROLL 1
#endif
#ifndef DNG_STR_UNROLL
#define DNG_STR_UNROLL 2
#endif

/* unroll factors of the matcher's two inner loops (measured: see DESIGN.md) */
#ifndef DNG_LIT_UNROLL
#define DNG_LIT_UNROLL 1
#endif
#ifndef DNG_STR_UNROLL
#define DNG_STR_UNROLL 2
#endif

namespace dng {

static constexpr int TM_LIT_UNROLL = DNG_LIT_UNROLL;
static constexpr int TM_STR_UNROLL = DNG_STR_UNROLL;

struct TQuad { u32 x, y, z, w; };	/* one TNode as four words */

/* 0x80 in every byte of x that is zero, exact for the LOWEST such byte */
DNG_HD u32 tz_low(u32 x)
{
	return (x - 0x01010101u) & ~x & 0x80808080u;
}

DNG_HD u32 tm_isdigit(u32 c)
{
	return c - '0' <= 9u;
}

/* warp vote: the matcher's outer loop is uniform across the warp so that
 * lanes reconverge after every node (the scans inside diverge) */
#ifdef __CUDA_ARCH__
#define DNG_WARP_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define DNG_WARP_ANY(x) (x)
#endif

/* 0x80 in every byte of w that is not an ASCII digit */
DNG_HD u32 nondigit_mask(u32 w)
{
	const u32 x = w ^ 0x30303030u;
	return (((x & 0x7f7f7f7fu) + 0x76767676u) | x) & 0x80808080u;
}

/* index of the lowest byte flagged 0x80 in m (m != 0) */
DNG_HD u32 low_flag_byte(u32 m)
{
#ifdef __CUDA_ARCH__
	return (__ffs(m) - 8) >> 3;
#else
	u32 k = 0;
	while (!((m >> (8 * k + 7)) & 1))
		k++;
	return k;
#endif
}

/*
 * Which of a sibling group's nodes can match a record whose byte at the
 * group's dispatch offset is what it is (tmpl.h THdr): the node, or TN_NOALT.
 */
template <class M>
DNG_HD u32 tmpl_dispatch(M &m, u32 disp, u32 p, u32 len)
{
	const u32 head = m.pool32(4 * disp);
	const u32 k = head & 0xffff, n = head >> 16;
	if (p + k >= len)
		return TN_NOALT;
	const u32 b = m.byte(p + k);
	u32 node = TN_NOALT;
#pragma unroll 1
	for (u32 i = 0; i < n; i++) {
		const u32 e = m.pool32(4 * disp + 4 + 4 * i);
		if ((e & 0xff) == b)
			node = e >> 16;
	}
	return node;
}

/*
 * Every lane of the warp must call this (active = false for lanes without a
 * record): the node loop runs until no lane is active.
 */
template <class M>
DNG_HD bool tmpl_match(M &m, u32 len, RecState &R, bool active)
{
	u32 p = 0;
	bool matched = false;
	TQuad nd = m.node(0);
	while (DNG_WARP_ANY(active)) {
		if (active) {
		/* the likely successor is fetched while this node is matched */
		const u32 succ = nd.y & 0xffff;
		const TQuad nx = m.node(succ & TN_LEAF ? 0 : succ);
		const u32 lit = nd.x & 0xffff, L = nd.x >> 16;
		bool ok = p + L <= len;
		if (ok && L) {
			/* eight bytes a step against (value, mask) word pairs:
			 * the literal's padding is masked out, so no tail case */
			typename M::Cur c = m.cursor(p);
			u32 diff = 0;
#pragma unroll TM_LIT_UNROLL
			for (u32 k = 0; k < L; k += 8) {
				const TQuad lq = m.lit2(lit + 2 * k);
				const u32 d0 = c.next(), d1 = c.next();
				diff |= ((d0 ^ lq.x) & lq.y) | ((d1 ^ lq.z) & lq.w);
			}
			ok = diff == 0;
		}
		u32 q = p + L;
		u64 val = 0;
		const u32 kind = nd.z & 0xff;
		if (ok && kind == TK_STR) {
			/* first '"', '\\' or control byte from q on; an escape
			 * is checked and stepped over, and the scan goes on */
			u32 e = q, esc = 0;
			for (;;) {
				typename M::Cur c = m.cursor(e);
				u32 hit, w;
#pragma unroll TM_STR_UNROLL
				for (;;) {
					w = c.next();
					hit = tz_low(w ^ 0x22222222u) |
					    tz_low(w ^ 0x5c5c5c5cu) |
					    tz_low(w & 0xe0e0e0e0u);
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
				/* \" \\ \/ \b \f \n \r \t \uXXXX (what JSON.parse
				 * accepts); the record's '\n' ends a truncated one */
				const u32 c1 = m.byte(e + 1);
				if (c1 == 'u') {
					ok = is_hex(m.byte(e + 2)) &&
					    is_hex(m.byte(e + 3)) &&
					    is_hex(m.byte(e + 4)) &&
					    is_hex(m.byte(e + 5));
					e += 6;
				} else {
					ok = c1 == '"' || c1 == '\\' || c1 == '/' ||
					    c1 == 'b' || c1 == 'f' || c1 == 'n' ||
					    c1 == 'r' || c1 == 't';
					e += 2;
				}
				esc = VF_ESCAPED;
				if (!ok)
					break;
			}
			val = mkval(T_STR, q, e - q, esc);
			q = e;
		} else if (ok && kind == TK_BARE) {
			typename M::Cur c = m.cursor(q);
			u32 w = c.next();
			const u32 c0 = w & 0xff;
			if (c0 == 't') {
				ok = w == 0x65757274u;
				val = mkval(T_TRUE, q, 4, 0);
				q += 4;
			} else if (c0 == 'n') {
				ok = w == 0x6c6c756eu;
				val = mkval(T_NULL, q, 4, 0);
				q += 4;
			} else if (c0 == 'f') {
				ok = w == 0x736c6166u && (c.next() & 0xff) == 'e';
				val = mkval(T_FALSE, q, 5, 0);
				q += 5;
			} else {
				/* -?(0|[1-9][0-9]*) word-wise; a fraction or an
				 * exponent continues byte-wise */
				const u32 neg = c0 == '-';
				if (neg)	/* drop the sign: shift in one more byte */
					w = (w >> 8) | (m.byte(q + 4) << 24);
				const u32 d0 = w & 0xff;
				u32 i = q + neg, nd_ = 0, mk;
				while ((mk = nondigit_mask(w)) == 0) {
					nd_ += 4;
					w = neg ? m.word(i + nd_) : c.next();
				}
				nd_ += low_flag_byte(mk);
				ok = nd_ > 0 && !(d0 == '0' && nd_ > 1);
				i += nd_;
				u32 simple = VF_SIMPLEINT;
				if (nd_ > 15 || (neg && nd_ == 1 && d0 == '0'))
					simple = 0;
				u32 ch = (w >> (8 * low_flag_byte(mk))) & 0xff;
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
				val = mkval(T_NUM, q, i - q, simple);
				q = i;
			}
		}
		if (!ok) {
			/* this node is not it: its next sibling, same place */
			const u32 alt = nd.y >> 16;
			active = alt != TN_NOALT;
			nd = m.node(active ? alt : 0);
		} else {
			const u32 pc = (nd.z >> 16) & 0xff;
			const u32 cap = (nd.z >> 8) & 0xff;
			if (pc) {
				const u32 at = p + (nd.w & 0xffff);
				R.slots[pc - 1] = mkval(m.byte(at) == '{' ?
				    T_OBJ : T_ARR, at, 0, 0);
			}
			if (cap)
				R.slots[cap - 1] = val;
			p = q;
			const u32 disp = nd.w >> 16;
			if (succ & TN_LEAF) {
				active = false;
				if (p == len) {
					R.set_mask = m.leaf(succ & 0x7fff);
					R.flags = 0;
					matched = true;
				}
			}
			nd = nx;
			if (disp != TN_NODISP) {
				/* several children: straight to the one this
				 * record's byte selects */
				const u32 ch = tmpl_dispatch(m, disp, p, len);
				active = ch != TN_NOALT;
				nd = m.node(active ? ch : 0);
			}
		}
		}
	}
	return matched;
}

#ifndef __CUDACC__
/* host access (tests/hostcheck only): plain memory; past the record a '\n',
 * then quotes (fscan.cuh: what the kernel's buffers guarantee) */
struct TmplHostMem {
	const u8 *rec;
	u32 len;
	const u8 *blob;

	struct Cur {
		const TmplHostMem *m;
		u32 off;
		u32 next() { u32 w = m->word(off); off += 4; return w; }
	};
	struct ACur {
		const TmplHostMem *m;
		u32 off, k;
		u32 next() { u32 w = m->word(off); off += 4; return w; }
	};
	ACur acursor(u32 off) const {
		ACur c; c.m = this; c.k = off & 3; c.off = off - c.k; return c;
	}
	u32 apos(const ACur &c) const { return c.off - 4; }
	u32 byte(u32 off) const {
		return off < len ? rec[off] : off == len ? (u32)'\n' : (u32)'"';
	}
	u32 word(u32 off) const {
		return byte(off) | (byte(off + 1) << 8) | (byte(off + 2) << 16) |
		    (byte(off + 3) << 24);
	}
	Cur cursor(u32 off) const { Cur c; c.m = this; c.off = off; return c; }
	TQuad node(u32 i) const {
		TQuad v;
		memcpy(&v, blob + sizeof (THdr) + 16 * (size_t)i, 16);
		return v;
	}
	TQuad lit2(u32 off) const {
		TQuad v;
		memcpy(&v, blob + ((const THdr *)blob)->pool_off + off, 16);
		return v;
	}
	u32 leaf(u32 i) const {
		u32 v;
		memcpy(&v, blob + ((const THdr *)blob)->leaf_off + 4 * i, 4);
		return v;
	}
	u32 pool32(u32 off) const {
		u32 v;
		memcpy(&v, blob + ((const THdr *)blob)->pool_off + off, 4);
		return v;
	}
};
#endif

} /* namespace dng */
#endif
