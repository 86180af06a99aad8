/*
 * record.cuh: everything that happens to ONE record (one line of input):
 *
 *   parse_record()    strict JSON validation (what JSON.parse accepts,
 *                     lib/format-json.js:34) fused with capture of the values
 *                     at the plan's dotted paths (jsprim.pluck semantics)
 *   process_record()  datasource filter -> user filter -> synthetic date
 *                     fields -> time-bounds filter -> group key
 *                     (lib/stream-scan.js:56-86; krill-skinner-stream.js:29-52;
 *                     stream-synthetic.js:37-85; dragnet-impl.js:48-125)
 *
 * One GPU thread runs this per record, reading the record bytes from the
 * shared-memory tile the CTA staged with TMA (scan_kernel.cu).  The code is
 * __host__ __device__ only so that tests/hostcheck can unit-test the exact
 * same logic on the CPU against oracle/; libdragnet_gpu.so never runs it on
 * the host.
 */
#ifndef DNG_RECORD_CUH
#define DNG_RECORD_CUH

#include "plan.h"
#include "jsnum.cuh"
#include "jsdate.cuh"

namespace dng {

/* packed value: off(32) | len(24) | type(4) | flags(4) */
DNG_HD u64 mkval(u32 type, u32 off, u32 len, u32 flags)
{
	return (u64)off | ((u64)(len & 0xffffff) << 32) | ((u64)type << 56) |
	    ((u64)flags << 60);
}
DNG_HD u32 val_off(u64 v) { return (u32)v; }
DNG_HD u32 val_len(u64 v) { return (u32)(v >> 32) & 0xffffff; }
DNG_HD u32 val_type(u64 v) { return (u32)(v >> 56) & 0xf; }
DNG_HD u32 val_flags(u64 v) { return (u32)(v >> 60) & 0xf; }

struct RecState {
	u64 slots[MAX_SLOTS];
	double syn[MAX_SYN];
	u32 set_mask;
	u32 flags;
};

struct LocalCounters {
	u32 lines, invalid_json, invalid_point;
	u32 ds_filtered, ds_failedeval, user_filtered, user_failedeval;
	u32 synth_undef, synth_baddate, time_filtered, time_failedeval;
	u32 aggr, slow, unsupported;
};

enum { S_VALUE = 0, S_OBJ_FIRST, S_OBJ_KEY, S_COLON, S_ARR_FIRST, S_AFTER };

DNG_HD bool is_hex(u32 c)
{
	return (c >= '0' && c <= '9') || ((c | 0x20) >= 'a' && (c | 0x20) <= 'f');
}

DNG_HD u32 hexval(u32 c)
{
	return c <= '9' ? c - '0' : (c | 0x20) - 'a' + 10;
}

/*
 * Decode one unit of a JSON string body at s[i] into out[0..4); advances i.
 * The body has already been validated.  Surrogate pairs combine to one
 * 4-byte UTF-8 sequence; a lone surrogate becomes its 3-byte (WTF-8) form.
 */
DNG_HD int decode_unit(const u8 *s, u32 n, u32 &i, u8 *out)
{
	u32 c = s[i++];
	if (c != '\\') {
		out[0] = (u8)c;
		return 1;
	}
	u32 e = s[i++];
	switch (e) {
	case 'b': out[0] = 8; return 1;
	case 'f': out[0] = 12; return 1;
	case 'n': out[0] = 10; return 1;
	case 'r': out[0] = 13; return 1;
	case 't': out[0] = 9; return 1;
	case 'u': break;
	default: out[0] = (u8)e; return 1;	/* " \ / */
	}
	u32 cp = (hexval(s[i]) << 12) | (hexval(s[i + 1]) << 8) |
	    (hexval(s[i + 2]) << 4) | hexval(s[i + 3]);
	i += 4;
	if (cp >= 0xD800 && cp < 0xDC00 && i + 6 <= n && s[i] == '\\' &&
	    s[i + 1] == 'u') {
		u32 lo = (hexval(s[i + 2]) << 12) | (hexval(s[i + 3]) << 8) |
		    (hexval(s[i + 4]) << 4) | hexval(s[i + 5]);
		if (lo >= 0xDC00 && lo < 0xE000) {
			cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
			i += 6;
		}
	}
	if (cp < 0x80) {
		out[0] = (u8)cp;
		return 1;
	}
	if (cp < 0x800) {
		out[0] = (u8)(0xC0 | (cp >> 6));
		out[1] = (u8)(0x80 | (cp & 0x3F));
		return 2;
	}
	if (cp < 0x10000) {
		out[0] = (u8)(0xE0 | (cp >> 12));
		out[1] = (u8)(0x80 | ((cp >> 6) & 0x3F));
		out[2] = (u8)(0x80 | (cp & 0x3F));
		return 3;
	}
	out[0] = (u8)(0xF0 | (cp >> 18));
	out[1] = (u8)(0x80 | ((cp >> 12) & 0x3F));
	out[2] = (u8)(0x80 | ((cp >> 6) & 0x3F));
	out[3] = (u8)(0x80 | (cp & 0x3F));
	return 4;
}

/* escaped JSON string body == target bytes? */
DNG_HDN bool str_eq_escaped(const u8 *s, u32 n, const u8 *t, u32 tn)
{
	u32 i = 0, j = 0;
	u8 tmp[4];
	while (i < n) {
		int k = decode_unit(s, n, i, tmp);
		if (j + k > tn)
			return false;
		for (int x = 0; x < k; x++)
			if (tmp[x] != t[j + x])
				return false;
		j += k;
	}
	return j == tn;
}

/* decode an escaped body into out (cap bytes); returns length or -1 */
DNG_HDN int str_decode(const u8 *s, u32 n, u8 *out, u32 cap)
{
	u32 i = 0, j = 0;
	u8 tmp[4];
	while (i < n) {
		int k = decode_unit(s, n, i, tmp);
		if (j + k > cap)
			return -1;
		for (int x = 0; x < k; x++)
			out[j + x] = tmp[x];
		j += k;
	}
	return (int)j;
}

/*
 * Validate the line as JSON and capture the plan's paths.
 * Sets RF_INVALID in R.flags when JSON.parse would throw.
 */
DNG_HD void parse_record(const u8 *rec, u32 len, const DevPlan &P, RecState &R)
{
	u32 i = 0;
	int depth = 0;
	u64 stk = 0;
	int state = S_VALUE;
	int ctx = -1, ctx_depth = 0;
	int pend_term = -1, pend_child = -1;
	u32 set_mask = 0;
	u32 flags = 0;
	u32 c = 0;
	/* array-valued contexts (jsprim.pluck indexes arrays through
	 * hasOwnProperty: "0", "1", ... and "length") */
	u64 arrbits = 0;		/* bit 0: current context is an array */
	u32 arr_idx[MAX_LEVELS + 2];	/* next element index, per ctx depth */

	for (;;) {
		for (;;) {
			if (i >= len)
				goto done;
			c = rec[i];
			if (c != ' ' && c != '\t' && c != '\r' && c != '\n')
				break;
			i++;
		}

		switch (state) {
		case S_OBJ_FIRST:
			if (c == '}')
				goto close;
			/* FALLTHROUGH */
		case S_OBJ_KEY: {
			if (c != '"')
				goto invalid;
			i++;
			u32 kstart = i;
			u32 hash = 2166136261u;
			u32 esc = 0;
			for (;;) {
				if (i >= len)
					goto invalid;
				c = rec[i];
				if (c == '"')
					break;
				if (c < 0x20)
					goto invalid;
				if (c == '\\') {
					esc = 1;
					i++;
					if (i >= len)
						goto invalid;
					u32 e = rec[i];
					if (e == 'u') {
						if (i + 4 >= len ||
						    !is_hex(rec[i + 1]) ||
						    !is_hex(rec[i + 2]) ||
						    !is_hex(rec[i + 3]) ||
						    !is_hex(rec[i + 4]))
							goto invalid;
						i += 4;
					} else if (e != '"' && e != '\\' &&
					    e != '/' && e != 'b' && e != 'f' &&
					    e != 'n' && e != 'r' && e != 't') {
						goto invalid;
					}
				}
				hash = (hash ^ c) * 16777619u;
				i++;
			}
			u32 klen = i - kstart;
			i++;
			pend_term = -1;
			pend_child = -1;
			if (depth == ctx_depth && ctx >= 0) {
				const Ctx &cx = P.hot.ctx[ctx];
				if (esc) {
					flags |= RF_SLOW;
					for (u32 ci = cx.cand_begin;
					    ci < cx.cand_end; ci++) {
						const Cand &cd = P.cand[ci];
						if (str_eq_escaped(rec + kstart,
						    klen, (const u8 *)P.pool +
						    cd.off, cd.len)) {
							pend_term = cd.term_slot;
							pend_child = cd.child_ctx;
							break;
						}
					}
				} else if ((cx.bloom >> (hash & 63)) & 1) {
					for (u32 ci = cx.cand_begin;
					    ci < cx.cand_end; ci++) {
						const Cand &cd = P.cand[ci];
						if (cd.hash != hash ||
						    cd.len != klen)
							continue;
						const u8 *t = (const u8 *)
						    P.pool + cd.off;
						u32 x = 0;
						while (x < klen &&
						    rec[kstart + x] == t[x])
							x++;
						if (x == klen) {
							pend_term = cd.term_slot;
							pend_child = cd.child_ctx;
							break;
						}
					}
				}
			}
			state = S_COLON;
			continue;
		}
		case S_COLON:
			if (c != ':')
				goto invalid;
			i++;
			state = S_VALUE;
			continue;
		case S_ARR_FIRST:
			if (c == ']')
				goto close;
			/* FALLTHROUGH */
		case S_VALUE: {
			u32 vstart = i;
			u64 v;
			if ((arrbits & 1) && depth == ctx_depth && ctx >= 0) {
				/* element of an array context: its key is the
				 * decimal index */
				const Ctx &cx = P.hot.ctx[ctx];
				u32 lvl = cx.depth;
				u32 idx = arr_idx[lvl]++;
				char ib[12];
				u32 il = 0;
				do {
					ib[il++] = (char)('0' + idx % 10);
					idx /= 10;
				} while (idx);
				pend_term = -1;
				pend_child = -1;
				for (u32 ci = cx.cand_begin; ci < cx.cand_end;
				    ci++) {
					const Cand &cd = P.cand[ci];
					if (cd.len != il)
						continue;
					const u8 *t = (const u8 *)P.pool + cd.off;
					u32 x = 0;
					while (x < il && t[x] == (u8)ib[il - 1 - x])
						x++;
					if (x == il) {
						pend_term = cd.term_slot;
						pend_child = cd.child_ctx;
						break;
					}
				}
			}
			if (c == '"') {
				i++;
				u32 esc = 0;
				for (;;) {
					if (i >= len)
						goto invalid;
					c = rec[i];
					if (c == '"')
						break;
					if (c < 0x20)
						goto invalid;
					if (c == '\\') {
						esc = VF_ESCAPED;
						i++;
						if (i >= len)
							goto invalid;
						u32 e = rec[i];
						if (e == 'u') {
							if (i + 4 >= len ||
							    !is_hex(rec[i + 1]) ||
							    !is_hex(rec[i + 2]) ||
							    !is_hex(rec[i + 3]) ||
							    !is_hex(rec[i + 4]))
								goto invalid;
							i += 4;
						} else if (e != '"' &&
						    e != '\\' && e != '/' &&
						    e != 'b' && e != 'f' &&
						    e != 'n' && e != 'r' &&
						    e != 't') {
							goto invalid;
						}
					}
					i++;
				}
				v = mkval(T_STR, vstart + 1, i - vstart - 1, esc);
				i++;
			} else if (c == '{' || c == '[') {
				u32 isobj = c == '{';
				if (depth >= 64) {
					flags |= RF_UNSUPPORTED;
					goto invalid;
				}
				if (pend_term >= 0) {
					R.slots[pend_term] = mkval(isobj ?
					    T_OBJ : T_ARR, vstart, 0, 0);
					set_mask |= 1u << pend_term;
				}
				int enter = -1;
				if (pend_child >= 0) {
					set_mask &=
					    ~P.hot.ctx[pend_child].subtree_mask;
					enter = pend_child;
				} else if (depth == 0 && P.hot.nctx) {
					enter = 0;
				}
				if (enter >= 0 && (isobj ||
				    P.hot.ctx[enter].arraylike)) {
					ctx = enter;
					ctx_depth = depth + 1;
					arrbits = (arrbits << 1) | (isobj ? 0 : 1);
					if (!isobj)
						arr_idx[P.hot.ctx[enter].depth] = 0;
				}
				pend_term = -1;
				pend_child = -1;
				stk = (stk << 1) | isobj;
				depth++;
				i++;
				state = isobj ? S_OBJ_FIRST : S_ARR_FIRST;
				continue;
			} else if (c == '-' || (c >= '0' && c <= '9')) {
				u32 neg = c == '-';
				if (neg) {
					i++;
					if (i >= len)
						goto invalid;
					c = rec[i];
				}
				if (c == '0') {
					i++;
				} else if (c >= '1' && c <= '9') {
					do {
						i++;
					} while (i < len && rec[i] >= '0' &&
					    rec[i] <= '9');
				} else {
					goto invalid;
				}
				u32 simple = VF_SIMPLEINT;
				if (i < len && rec[i] == '.') {
					simple = 0;
					i++;
					if (i >= len || rec[i] < '0' ||
					    rec[i] > '9')
						goto invalid;
					while (i < len && rec[i] >= '0' &&
					    rec[i] <= '9')
						i++;
				}
				if (i < len && (rec[i] | 0x20) == 'e') {
					simple = 0;
					i++;
					if (i < len && (rec[i] == '+' ||
					    rec[i] == '-'))
						i++;
					if (i >= len || rec[i] < '0' ||
					    rec[i] > '9')
						goto invalid;
					while (i < len && rec[i] >= '0' &&
					    rec[i] <= '9')
						i++;
				}
				u32 nlen = i - vstart;
				if (nlen - neg > 15 ||
				    (neg && nlen == 2 && rec[vstart + 1] == '0'))
					simple = 0;
				v = mkval(T_NUM, vstart, nlen, simple);
			} else if (c == 't') {
				if (i + 4 > len || rec[i + 1] != 'r' ||
				    rec[i + 2] != 'u' || rec[i + 3] != 'e')
					goto invalid;
				i += 4;
				v = mkval(T_TRUE, vstart, 4, 0);
			} else if (c == 'f') {
				if (i + 5 > len || rec[i + 1] != 'a' ||
				    rec[i + 2] != 'l' || rec[i + 3] != 's' ||
				    rec[i + 4] != 'e')
					goto invalid;
				i += 5;
				v = mkval(T_FALSE, vstart, 5, 0);
			} else if (c == 'n') {
				if (i + 4 > len || rec[i + 1] != 'u' ||
				    rec[i + 2] != 'l' || rec[i + 3] != 'l')
					goto invalid;
				i += 4;
				v = mkval(T_NULL, vstart, 4, 0);
			} else {
				goto invalid;
			}
			/* a scalar value is complete */
			if (pend_term >= 0) {
				R.slots[pend_term] = v;
				set_mask |= 1u << pend_term;
			}
			if (pend_child >= 0)
				set_mask &= ~P.hot.ctx[pend_child].subtree_mask;
			pend_term = -1;
			pend_child = -1;
			state = S_AFTER;
			continue;
		}
		case S_AFTER:
			if (depth == 0)
				goto invalid;
			if (c == ',') {
				i++;
				state = (stk & 1) ? S_OBJ_KEY : S_VALUE;
				continue;
			}
			if (c == '}' || c == ']')
				goto close;
			goto invalid;
		}
close:
		/* c is '}' or ']' and depth >= 1 */
		if (((u32)(stk & 1)) != (u32)(c == '}'))
			goto invalid;
		if (depth == ctx_depth && ctx >= 0) {
			const Ctx &cx = P.hot.ctx[ctx];
			if (arrbits & 1) {
				/* arr.length */
				for (u32 ci = cx.cand_begin; ci < cx.cand_end;
				    ci++) {
					const Cand &cd = P.cand[ci];
					const u8 *t = (const u8 *)P.pool + cd.off;
					if (cd.len == 6 && t[0] == 'l' &&
					    t[1] == 'e' && t[2] == 'n' &&
					    t[3] == 'g' && t[4] == 't' &&
					    t[5] == 'h' && cd.term_slot >= 0) {
						R.slots[cd.term_slot] = mkval(
						    T_NUM, arr_idx[cx.depth], 0,
						    VF_INLINE);
						set_mask |= 1u << cd.term_slot;
					}
				}
			}
			arrbits >>= 1;
			ctx = cx.parent;
			ctx_depth--;
		}
		stk >>= 1;
		depth--;
		i++;
		state = S_AFTER;
	}
done:
	if (state != S_AFTER || depth != 0)
		goto invalid;
	R.set_mask = set_mask;
	R.flags = flags;
	return;
invalid:
	R.set_mask = 0;
	R.flags = flags | RF_INVALID;
}

/* ---- fast path: lock-step byte automaton (plan.h FastTab) ----------------- */

struct FastState {
	u32 state;
	u32 stk;	/* container types, innermost in bit 0 (1 = object), under a
			 * sentinel 1: stk == 1 means depth 0 */
	u32 ctx_stk;	/* value of stk while directly inside the context's
			 * container; 0 = no context */
	u32 arm, vstart, set_mask;
	int ctx, pend_term, pend_child;
};

DNG_HD void fast_init(FastState &s)
{
	s.state = FS_START;
	s.stk = 1;
	s.ctx_stk = 0;
	s.arm = FE_PUSH | FE_POP | FE_KEYHIT;
	s.vstart = 0;
	s.set_mask = 0;
	s.ctx = -1;
	s.pend_term = -1;
	s.pend_child = -1;
}

/* the value after a candidate key has ended at `end` (exclusive) */
DNG_HD void fast_capture(FastState &s, const HotPlan &P, u64 *slots, u32 end)
{
	if (s.pend_term >= 0) {
		slots[s.pend_term] = (u64)s.vstart | ((u64)end << 32);
		s.set_mask |= 1u << s.pend_term;
	}
	if (s.pend_child >= 0)
		s.set_mask &= ~P.ctx[s.pend_child].subtree_mask;
	s.pend_term = -1;
	s.pend_child = -1;
	s.arm &= ~(u32)(FE_VALSTART | FE_VALEND_INCL | FE_VALEND_EXCL);
}

/* the divergent part: runs only on bytes whose transition carries an armed
 * event flag (container open/close, a candidate key, the value after one) */
DNG_HD void fast_event(FastState &s, const HotPlan &P, u64 *slots, u32 f,
    u32 pos)
{
	const u32 VAL = FE_VALSTART | FE_VALEND_INCL | FE_VALEND_EXCL;
	if (f & (FE_PUSH | FE_POP)) {
		if (f & s.arm & FE_VALEND_EXCL)	/* 123} : value ends here */
			fast_capture(s, P, slots, pos);
		if (f & FE_PUSH) {
			u32 isobj = (f & FE_OBJ) ? 1u : 0u;
			if ((s.pend_term & s.pend_child) != -1 || s.stk == 1) {
				/* a container that matters: a captured value,
				 * a context to enter, or the top-level value */
				if (s.pend_term >= 0) {
					slots[s.pend_term] = (u64)pos; /* end 0 */
					s.set_mask |= 1u << s.pend_term;
				}
				int enter = -1;
				if (s.pend_child >= 0) {
					s.set_mask &=
					    ~P.ctx[s.pend_child].subtree_mask;
					enter = s.pend_child;
				} else if (s.stk == 1 && P.nctx) {
					enter = 0;
				}
				if (enter >= 0 && isobj) {
					s.ctx = enter;
					s.ctx_stk = (s.stk << 1) | 1u;
				}
				s.pend_term = -1;
				s.pend_child = -1;
				s.arm &= ~VAL;
			}
			if (s.stk >> 30)
				s.state = FS_FB;	/* too deep for 32 bits */
			else
				s.stk = (s.stk << 1) | isobj;
		} else {
			if (s.stk == s.ctx_stk) {	/* leaving the context */
				s.ctx = P.ctx[s.ctx].parent;
				s.ctx_stk = s.ctx >= 0 ? s.stk >> 1 : 0;
			}
			s.stk >>= 1;
			s.state = s.stk == 1 ? (u32)FS_DONE :
			    (u32)FS_AFTER_A - (s.stk & 1);
		}
		return;
	}
	if (f & FE_KEYHIT) {
		u32 g = s.state - P.fast.kc_base;
		s.state = FS_KC;
		if (s.stk == s.ctx_stk) {
			const u8 *cm = P.fast.candmap[s.ctx * FAST_MAXKEYS + g];
			if (cm[0] != 0xFF || cm[1] != 0xFF) {
				s.pend_term = (int8_t)cm[0];
				s.pend_child = (int8_t)cm[1];
				s.arm |= VAL;
			}
		}
		return;
	}
	if (f & s.arm & FE_VALSTART)
		s.vstart = pos;
	if (f & s.arm & (FE_VALEND_INCL | FE_VALEND_EXCL))
		fast_capture(s, P, slots, pos + ((f & FE_VALEND_INCL) ? 1u : 0u));
}

DNG_HD void fast_step(FastState &s, const HotPlan &P, u64 *slots, u32 c, u32 pos)
{
	u32 e = P.trans[s.state * P.fast.stride + P.fast.cls[c]];
	s.state = e & 0xff;
	u32 f = e >> 8;
	if (f & s.arm)
		fast_event(s, P, slots, f, pos);
}

/* turn the (start, end) spans the automaton captured into packed values */
DNG_HD void fast_finish(const u8 *rec, const FastState &s, RecState &R)
{
	R.set_mask = s.set_mask;
	R.flags = 0;
	u32 m = s.set_mask;
	while (m) {
		u32 slot = 0;
		while (!((m >> slot) & 1))
			slot++;
		m &= m - 1;
		u32 start = (u32)R.slots[slot], end = (u32)(R.slots[slot] >> 32);
		u32 c0 = rec[start];
		u64 v;
		if (end == 0) {
			v = mkval(c0 == '{' ? T_OBJ : T_ARR, start, 0, 0);
		} else if (c0 == '"') {
			u32 esc = 0;
			for (u32 i = start + 1; i + 1 < end; i++)
				if (rec[i] == '\\')
					esc = VF_ESCAPED;
			v = mkval(T_STR, start + 1, end - start - 2, esc);
		} else if (c0 == 't') {
			v = mkval(T_TRUE, start, 4, 0);
		} else if (c0 == 'f') {
			v = mkval(T_FALSE, start, 5, 0);
		} else if (c0 == 'n') {
			v = mkval(T_NULL, start, 4, 0);
		} else {
			u32 neg = c0 == '-', simple = VF_SIMPLEINT;
			u32 nlen = end - start;
			for (u32 i = start + neg; i < end; i++)
				if (rec[i] < '0' || rec[i] > '9')
					simple = 0;
			if (nlen - neg > 15 ||
			    (neg && nlen == 2 && rec[start + 1] == '0'))
				simple = 0;
			v = mkval(T_NUM, start, nlen, simple);
		}
		R.slots[slot] = v;
	}
}

/* ---- values ------------------------------------------------------------ */

struct V {
	u64 pk;
	double num;		/* when VF_BINARY */
};
enum : u32 { VF_BINARY = 4 };	/* VF_INLINE (plan.h) = 8 */

DNG_HD V get_src(const DevPlan &P, const RecState &R, Src s)
{
	V v;
	v.num = 0;
	v.pk = mkval(T_UNDEF, 0, 0, 0);
	if (s.kind == SRC_SYNTH) {
		v.pk = mkval(T_NUM, 0, 0, VF_BINARY);
		v.num = R.syn[s.idx];
	} else if (s.kind == SRC_PATH) {
		const PathInfo &pp = P.path[s.idx];
		for (u32 l = 0; l < pp.nlevels; l++) {
			u32 slot = pp.slot0 + l;
			if ((R.set_mask >> slot) & 1) {
				v.pk = R.slots[slot];
				break;
			}
		}
	}
	return v;
}

DNG_HD void put_bytes(u8 *out, u32 &o, u32 cap, const char *s, u32 n, u32 &ovf)
{
	if (o + n > cap) {
		ovf = 1;
		return;
	}
	for (u32 k = 0; k < n; k++)
		out[o + k] = (u8)s[k];
	o += n;
}

DNG_HD double simple_int(const u8 *p, u32 n)
{
	u32 i = 0, neg = 0;
	if (p[0] == '-') {
		neg = 1;
		i = 1;
	}
	u64 w = 0;
	for (; i < n; i++)
		w = w * 10 + (p[i] - '0');
	double d = (double)w;
	return neg ? -d : d;
}

DNG_HD double json_number(const u8 *rec, u64 pk)
{
	if (val_flags(pk) & VF_INLINE)
		return (double)val_off(pk);
	if (val_flags(pk) & VF_SIMPLEINT)
		return simple_int(rec + val_off(pk), val_len(pk));
	return dng_parse_decimal(rec + val_off(pk), (int)val_len(pk));
}

/*
 * ToString of an array value: Array.prototype.join(",") applied recursively
 * (null/undefined elements are empty; objects are "[object Object]").
 * `rec + off` points at '['.  The record is known to be valid JSON.
 */
DNG_HDN void stringify_array(const u8 *rec, u32 off, u8 *out, u32 &o, u32 cap,
    u32 &ovf)
{
	u32 i = off;
	int adepth = 0;
	for (;;) {
		u32 c = rec[i];
		if (c == ' ' || c == '\t' || c == '\r' || c == '\n') {
			i++;
		} else if (c == '[') {
			adepth++;
			i++;
		} else if (c == ']') {
			adepth--;
			i++;
			if (adepth == 0)
				return;
		} else if (c == ',') {
			put_bytes(out, o, cap, ",", 1, ovf);
			i++;
		} else if (c == '{') {
			put_bytes(out, o, cap, "[object Object]", 15, ovf);
			int d = 0;
			for (;;) {		/* skip the object */
				c = rec[i];
				if (c == '"') {
					i++;
					while (rec[i] != '"') {
						if (rec[i] == '\\')
							i++;
						i++;
					}
				} else if (c == '{' || c == '[') {
					d++;
				} else if (c == '}' || c == ']') {
					d--;
				}
				i++;
				if (d == 0)
					break;
			}
		} else if (c == '"') {
			u32 s = ++i;
			u32 esc = 0;
			while (rec[i] != '"') {
				if (rec[i] == '\\') {
					esc = 1;
					i++;
				}
				i++;
			}
			if (!esc) {
				put_bytes(out, o, cap, (const char *)rec + s,
				    i - s, ovf);
			} else {
				int k = str_decode(rec + s, i - s, out + o,
				    cap - o);
				if (k < 0)
					ovf = 1;
				else
					o += k;
			}
			i++;
		} else if (c == 't') {
			put_bytes(out, o, cap, "true", 4, ovf);
			i += 4;
		} else if (c == 'f') {
			put_bytes(out, o, cap, "false", 5, ovf);
			i += 5;
		} else if (c == 'n') {
			i += 4;		/* null -> "" */
		} else {
			u32 s = i;
			while ((rec[i] >= '0' && rec[i] <= '9') ||
			    rec[i] == '-' || rec[i] == '+' || rec[i] == '.' ||
			    (rec[i] | 0x20) == 'e')
				i++;
			char nb[32];
			int k = dng_number_to_string(dng_parse_decimal(rec + s,
			    (int)(i - s)), nb);
			put_bytes(out, o, cap, nb, k, ovf);
		}
	}
}

/*
 * ToString(v) appended to out at o.  Used for discrete group keys
 * (String(value) as a JS property key) and Date.parse arguments.
 * Out of line: everything but plain strings and plain integers.
 */
DNG_HDN void value_to_string_slow(const u8 *rec, u64 pk, double num, u8 *out,
    u32 &o, u32 cap, u32 &ovf)
{
	u32 t = val_type(pk), fl = val_flags(pk);
	switch (t) {
	case T_UNDEF: put_bytes(out, o, cap, "undefined", 9, ovf); break;
	case T_NULL: put_bytes(out, o, cap, "null", 4, ovf); break;
	case T_TRUE: put_bytes(out, o, cap, "true", 4, ovf); break;
	case T_FALSE: put_bytes(out, o, cap, "false", 5, ovf); break;
	case T_OBJ: put_bytes(out, o, cap, "[object Object]", 15, ovf); break;
	case T_NUM: {
		char nb[32];
		double d = (fl & VF_BINARY) ? num :
		    (fl & VF_INLINE) ? (double)val_off(pk) :
		    dng_parse_decimal(rec + val_off(pk), (int)val_len(pk));
		int k = dng_number_to_string(d, nb);
		put_bytes(out, o, cap, nb, k, ovf);
		break;
	}
	case T_STR: {
		int k = str_decode(rec + val_off(pk), val_len(pk), out + o,
		    cap - o);
		if (k < 0)
			ovf = 1;
		else
			o += k;
		break;
	}
	case T_ARR:
		stringify_array(rec, val_off(pk), out, o, cap, ovf);
		break;
	}
}

DNG_HD void value_to_string(const u8 *rec, const V &v, u8 *out, u32 &o, u32 cap,
    u32 &ovf, u32 &slow)
{
	u32 t = val_type(v.pk), fl = val_flags(v.pk);
	if ((t == T_STR && !(fl & VF_ESCAPED)) ||
	    (t == T_NUM && (fl & VF_SIMPLEINT))) {
		u32 n = val_len(v.pk), off = val_off(v.pk);
		if (o + n > cap) {
			ovf = 1;
			return;
		}
		for (u32 k = 0; k < n; k++)
			out[o + k] = rec[off + k];
		o += n;
		return;
	}
	if (t == T_STR || t == T_ARR || t == T_NUM)
		slow = 1;
	value_to_string_slow(rec, v.pk, v.num, out, o, cap, ovf);
}

/* ToNumber(v), out of line part; scratch is used for escaped strings/arrays */
DNG_HDN double value_to_number_slow(const u8 *rec, u64 pk, u8 *scratch, u32 cap,
    u32 &ovf)
{
	u32 t = val_type(pk), fl = val_flags(pk);
	switch (t) {
	case T_NULL:
	case T_FALSE:
		return 0.0;
	case T_TRUE:
		return 1.0;
	case T_NUM:
		if (fl & VF_INLINE)
			return (double)val_off(pk);
		return dng_parse_decimal(rec + val_off(pk), (int)val_len(pk));
	case T_STR:
		if (!(fl & VF_ESCAPED))
			return dng_string_to_number(rec + val_off(pk),
			    (int)val_len(pk));
		/* FALLTHROUGH */
	case T_ARR: {
		u32 o = 0;
		value_to_string_slow(rec, pk, 0.0, scratch, o, cap, ovf);
		return dng_string_to_number(scratch, (int)o);
	}
	default:
		return dng_nan();	/* undefined, objects */
	}
}

DNG_HD double value_to_number(const u8 *rec, const V &v, u8 *scratch, u32 cap,
    u32 &ovf, u32 &slow)
{
	u32 t = val_type(v.pk), fl = val_flags(v.pk);
	if (t == T_NUM) {
		if (fl & VF_BINARY)
			return v.num;
		if (fl & VF_SIMPLEINT)
			return simple_int(rec + val_off(v.pk), val_len(v.pk));
	}
	if (t == T_STR || t == T_ARR)
		slow = 1;
	return value_to_number_slow(rec, v.pk, scratch, cap, ovf);
}

/* relational order of two UTF-8 strings by UTF-16 code units */
DNG_HD int utf16_cmp(const u8 *a, u32 an, const u8 *b, u32 bn)
{
	u32 n = an < bn ? an : bn;
	for (u32 i = 0; i < n; i++) {
		u32 x = a[i], y = b[i];
		if (x == y)
			continue;
		/* supplementary planes (surrogates D800..DBFF) sort below
		 * U+E000..U+FFFF in UTF-16 */
		if (x >= 0xF0 && (y == 0xEE || y == 0xEF))
			return -1;
		if (y >= 0xF0 && (x == 0xEE || x == 0xEF))
			return 1;
		return x < y ? -1 : 1;
	}
	return an == bn ? 0 : (an < bn ? -1 : 1);
}

/* one krill leaf: 1 true, 0 false, -1 evaluation failed (field undefined) */
DNG_HD int eval_leaf(const u8 *rec, const DevPlan &P, const RecState &R,
    const Leaf &lf, u8 *scratch, u32 &ovf, u32 &slow)
{
	if (lf.op == OP_TRUE)
		return 1;
	V v = get_src(P, R, lf.src);
	u32 t = val_type(v.pk);
	if (t == T_UNDEF)
		return -1;
	const u8 *cb = (const u8 *)P.pool + lf.coff;
	const u8 *sb = nullptr;		/* value as string, if string-like */
	u32 sn = 0;
	if (t == T_STR && !(val_flags(v.pk) & VF_ESCAPED)) {
		sb = rec + val_off(v.pk);
		sn = val_len(v.pk);
	} else if (t == T_STR || t == T_ARR || t == T_OBJ) {
		/* ToPrimitive(object) is its string form */
		u32 o = 0;
		value_to_string(rec, v, scratch, o, KEY_MAX, ovf, slow);
		sb = scratch;
		sn = o;
	}
	if (lf.op == OP_EQ || lf.op == OP_NE) {
		int eq;
		if (t == T_NULL) {
			eq = 0;
		} else if (sb && lf.cstr) {
			eq = sn == lf.clen;
			for (u32 k = 0; eq && k < sn; k++)
				eq = sb[k] == cb[k];
		} else {
			double x = sb ? dng_string_to_number(sb, (int)sn) :
			    value_to_number(rec, v, scratch, KEY_MAX, ovf, slow);
			eq = x == lf.cnum;
		}
		return lf.op == OP_EQ ? eq : !eq;
	}
	if (sb && lf.cstr) {
		int c = utf16_cmp(sb, sn, cb, lf.clen);
		switch (lf.op) {
		case OP_LT: return c < 0;
		case OP_LE: return c <= 0;
		case OP_GT: return c > 0;
		default: return c >= 0;
		}
	}
	double x = sb ? dng_string_to_number(sb, (int)sn) :
	    value_to_number(rec, v, scratch, KEY_MAX, ovf, slow);
	double y = lf.cnum;
	switch (lf.op) {
	case OP_LT: return x < y;
	case OP_LE: return x <= y;
	case OP_GT: return x > y;
	default: return x >= y;
	}
}

DNG_HDN int eval_program(const u8 *rec, const DevPlan &P, const RecState &R,
    int entry, u8 *scratch, u32 &ovf, u32 &slow)
{
	int pc = entry;
	while (pc >= 0) {
		const Leaf &lf = P.code[pc];
		int r = eval_leaf(rec, P, R, lf, scratch, ovf, slow);
		if (r < 0)
			return -1;
		pc = r ? lf.jt : lf.jf;
	}
	return pc == -1;
}

DNG_HD double p2_ordinal(double x)
{
	if (x != x)
		return x;
	if (x < 1.0)
		return 0.0;
	u64 b = double_to_bits(x);
	int be = (int)((b >> 52) & 0x7ff);
	if (be == 0x7ff)
		return x;		/* +Infinity */
	return (double)(be - 1023 + 1);
}

DNG_HD double linear_ordinal(double x, double step)
{
	double q = x / step;
	if (q != q)
		return q;
	return floor(q) + 0.0;		/* -0 -> +0 */
}

/*
 * Stages shared by every metric of a record: the json-skinner envelope and the
 * datasource filter (lib/datasource-file.js:154-163 puts it on the parser, in
 * front of all StreamScans).  Returns 1 if the record goes on.
 */
DNG_HD int prepare_record(const u8 *rec, const DevPlan &P, RecState &R,
    LocalCounters &C, u8 *kbuf, u64 &weight)
{
	u32 ovf = 0, slow = 0;
	weight = 1;
	if (P.format == FMT_SKINNER) {
		/* the line is a point {fields:{..}, value:N}
		 * (lib/format-json.js:55-73); anything else is dropped */
		u64 fv = ((R.set_mask >> P.sk_fields_slot) & 1) ?
		    R.slots[P.sk_fields_slot] : 0;
		u64 wv = ((R.set_mask >> P.sk_value_slot) & 1) ?
		    R.slots[P.sk_value_slot] : 0;
		double w = -1;
		if (val_type(wv) == T_NUM)
			w = json_number(rec, wv);
		if (val_type(fv) != T_OBJ || !(w >= 0) ||
		    w > 9007199254740992.0 || w != floor(w)) {
			C.invalid_point++;
			return 0;
		}
		weight = (u64)w;
	}
	if (P.ds_entry >= 0) {
		int r = eval_program(rec, P, R, P.ds_entry, kbuf, ovf, slow);
		if (ovf)
			C.unsupported++;
		if (r < 0) {
			C.ds_failedeval++;
			return 0;
		}
		if (!r) {
			C.ds_filtered++;
			return 0;
		}
	}
	return 1;
}

/*
 * One metric's StreamScan on a prepared record: user filter -> synthetic
 * dates -> time bounds -> group key (lib/stream-scan.js:56-86).  Returns 1
 * when the record reaches the aggregator; then kbuf[0..klen) is its encoded
 * group key -- [0xFD, metric] when the plan fans out to several metrics, then
 * per column u16 len + bytes, or 0xFFFF + 8 bytes of ordinal -- zero padded to
 * a multiple of 8.
 */
DNG_HD int process_metric(const u8 *rec, const DevPlan &P, u32 mi, RecState &R,
    LocalCounters &C, u8 *kbuf, u32 &klen)
{
	const Metric &M = P.metric[mi];
	u32 ovf = 0, slow = 0;
	klen = 0;
	if (M.user_entry >= 0) {
		int r = eval_program(rec, P, R, M.user_entry, kbuf, ovf, slow);
		if (r < 0) {
			C.user_failedeval++;
			goto dropped;
		}
		if (!r) {
			C.user_filtered++;
			goto dropped;
		}
	}
	if (M.nsyn) {
		u32 nerr = 0;
		for (u32 j = M.syn0; j < (u32)M.syn0 + M.nsyn; j++) {
			V v = get_src(P, R, P.syn[j]);
			u32 t = val_type(v.pk);
			if (t == T_UNDEF) {
				if (!nerr)
					C.synth_undef++;
				nerr++;
				continue;
			}
			if (t == T_NUM) {
				R.syn[j] = (val_flags(v.pk) & VF_BINARY) ?
				    v.num : json_number(rec, v.pk);
				continue;
			}
			int64_t ms = 0;
			bool ok = false;
			if (t == T_STR && !(val_flags(v.pk) & VF_ESCAPED)) {
				ok = dng_date_parse(rec + val_off(v.pk),
				    (int)val_len(v.pk), &ms);
				/* (not ISO, but V8's legacy parser might know it:
				 * not ours to call NaN, jsdate.cuh) */
				if (!ok && dng_date_maybe_legacy(rec + val_off(v.pk),
				    (int)val_len(v.pk)))
					ovf = 1;
			} else if (t == T_STR || t == T_ARR) {
				u32 o = 0;
				value_to_string(rec, v, kbuf, o, KEY_MAX, ovf,
				    slow);
				ok = dng_date_parse(kbuf, (int)o, &ms);
				if (!ok && t == T_STR &&
				    dng_date_maybe_legacy(kbuf, (int)o))
					ovf = 1;
			}
			if (!ok) {
				if (!nerr)
					C.synth_baddate++;
				nerr++;
				continue;
			}
			/* Math.floor(parsed / 1000), in binary64 like JS */
			R.syn[j] = floor((double)ms / 1000.0);
		}
		if (nerr)
			goto dropped;
	}
	if (M.time_entry >= 0) {
		int r = eval_program(rec, P, R, M.time_entry, kbuf, ovf, slow);
		if (r < 0) {
			C.time_failedeval++;
			goto dropped;
		}
		if (!r) {
			C.time_filtered++;
			goto dropped;
		}
	}
	{
		u32 o = 0;
		if (P.nmetrics > 1) {
			kbuf[0] = 0xFD;
			kbuf[1] = (u8)mi;
			o = 2;
		}
		for (u32 j = M.col0; j < (u32)M.col0 + M.ncols; j++) {
			const Col &col = P.col[j];
			V v = get_src(P, R, col.src);
			if (o + 10 > KEY_MAX) {
				ovf = 1;
				break;
			}
			if (col.kind == COL_DISCRETE) {
				u32 lp = o;
				o += 2;
				value_to_string(rec, v, kbuf, o, KEY_MAX, ovf,
				    slow);
				u32 n = o - lp - 2;
				if (n >= 0xFFFF)
					ovf = 1;
				kbuf[lp] = (u8)n;
				kbuf[lp + 1] = (u8)(n >> 8);
			} else {
				/* the scratch for ToNumber lives past the key */
				double x = value_to_number(rec, v, kbuf + o + 10,
				    KEY_MAX - (o + 10), ovf, slow);
				double ord = col.kind == COL_P2 ? p2_ordinal(x) :
				    linear_ordinal(x, col.step);
				u64 b = ord != ord ? 0x7ff8000000000000ull :
				    double_to_bits(ord);
				kbuf[o] = 0xFF;
				kbuf[o + 1] = 0xFF;
				for (int k = 0; k < 8; k++)
					kbuf[o + 2 + k] = (u8)(b >> (8 * k));
				o += 10;
			}
		}
		klen = o;
		while (o & 7)
			kbuf[o++] = 0;
	}
	if (ovf) {
		C.unsupported++;
		return 0;
	}
	if (slow || (R.flags & RF_SLOW))
		C.slow++;
	C.aggr++;
	return 1;
dropped:
	if (ovf)
		C.unsupported++;
	return 0;
}

/* same hash, key given as aligned 64-bit words (little endian) */
DNG_HD u64 key_hash_words(const unsigned long long *kw, u32 klen)
{
	u64 h = 0x9E3779B97F4A7C15ull ^ ((u64)klen * 0xff51afd7ed558ccdull);
	u32 n = (klen + 7) >> 3;
	for (u32 i = 0; i < n; i++) {
		h ^= kw[i];
		h *= 0xff51afd7ed558ccdull;
		h ^= h >> 32;
	}
	h ^= h >> 29;
	h *= 0xc4ceb9fe1a85ec53ull;
	h ^= h >> 32;
	return h;
}

/* 64-bit hash of an encoded key (klen rounded up to 8, zero padded) */
DNG_HD u64 key_hash(const u8 *kbuf, u32 klen)
{
	u64 h = 0x9E3779B97F4A7C15ull ^ ((u64)klen * 0xff51afd7ed558ccdull);
	u32 n = (klen + 7) >> 3;
	for (u32 i = 0; i < n; i++) {
		u64 w = 0;
		for (int k = 0; k < 8; k++)
			w |= (u64)kbuf[i * 8 + k] << (8 * k);
		h ^= w;
		h *= 0xff51afd7ed558ccdull;
		h ^= h >> 32;
	}
	h ^= h >> 29;
	h *= 0xc4ceb9fe1a85ec53ull;
	h ^= h >> 32;
	return h;
}

} /* namespace dng */
#endif
