/*
 * fast.cpp: host side of the F path (fast.h): which plans it takes, their
 * trimmed form, and the path-indexed view of a template's captures.
 */
#include <string.h>

#include "fast.h"

namespace dng {

void fplan_build(const DevPlan &P, FPlan &F)
{
	memset(&F, 0, sizeof (F));
	memcpy(F.pool + FC_UNDEFINED, "undefined", 9);
	memcpy(F.pool + FC_NULL, "null", 4);
	memcpy(F.pool + FC_TRUE, "true", 4);
	memcpy(F.pool + FC_FALSE, "false", 5);
	if (P.format != FMT_JSON || P.nmetrics != 1)
		return;
	const Metric &M = P.metric[0];
	if (P.npaths > F_MAXPATHS || P.ncode > F_MAXCODE ||
	    M.ncols > F_MAXCOLS || M.nsyn > F_MAXSYN || M.syn0 != 0 ||
	    M.col0 != 0)
		return;
	u32 pool = FC_END;
	for (u32 i = 0; i < P.ncode; i++) {
		Leaf lf = P.code[i];
		if (lf.src.kind == SRC_SYNTH && lf.src.idx >= M.nsyn)
			return;
		if (lf.src.kind == SRC_PATH && lf.src.idx >= P.npaths)
			return;
		if (lf.cstr) {
			/* constants 4-byte aligned and zero padded: compared
			 * word-wise */
			const u32 need = (lf.clen + 3u) & ~3u;
			if (pool + need > F_POOL)
				return;
			memcpy(F.pool + pool, P.pool + lf.coff, lf.clen);
			lf.coff = (u16)pool;
			pool += need;
		} else {
			lf.coff = lf.clen = 0;
		}
		/* (plan.cpp only ever patches a jump to a later leaf; the
		 * unrolled evaluator of the link-time optimised build relies
		 * on it) */
		if ((lf.jt >= 0 && lf.jt <= (int)i) || (lf.jf >= 0 && lf.jf <= (int)i))
			return;
		F.code[i] = lf;
	}
	F.ncode = P.ncode;
	for (u32 j = 0; j < M.nsyn; j++) {
		const Src s = P.syn[j];
		if (s.kind == SRC_PATH && s.idx < P.npaths)
			F.syn_path[j] = s.idx;
		else if (s.kind == SRC_UNDEF)
			F.syn_path[j] = 0xff;
		else
			return;		/* a date field of a date field */
	}
	for (u32 j = 0; j < M.ncols; j++) {
		const Col &c = P.col[j];
		if (c.src.kind == SRC_SYNTH &&
		    (c.kind == COL_DISCRETE || c.src.idx >= M.nsyn))
			return;		/* Number::toString of a date: general path */
		if (c.src.kind == SRC_PATH && c.src.idx >= P.npaths)
			return;
		F.col[j] = c;
	}
	u32 rows = P.npaths;
	for (u32 j = 0; j < M.ncols; j++) {
		F.ord_row[j] = 0xff;
		if (P.col[j].kind != COL_DISCRETE) {
			F.ord_row[j] = (u8)rows;
			rows += 2;
		}
	}
	F.nrows = (u8)(rows ? rows : 1);
	F.ds_entry = P.ds_entry;
	F.user_entry = M.user_entry;
	F.time_entry = M.time_entry;
	F.nsyn = M.nsyn;
	F.ncols = M.ncols;
	F.npaths = P.npaths;
	F.ok = 1;
}

bool fplan_resolve(const DevPlan &P, const TResolved &in, TResolved &out)
{
	memset(&out, 0, sizeof (out));
	/* (a non-zero flags word makes tmpl_build skip the candidate) */
	out.flags = in.flags ? in.flags : (u32)RF_UNSUPPORTED;
	if (in.flags != 0)
		return false;
	for (u32 p = 0; p < P.npaths && p < (u32)F_MAXPATHS; p++) {
		const PathInfo &pi = P.path[p];
		for (u32 l = 0; l < pi.nlevels; l++) {
			const u32 slot = pi.slot0 + l;
			if (!((in.set_mask >> slot) & 1))
				continue;
			/* the first level that is set wins (record.cuh get_src) */
			const u64 v = in.slots[slot];
			const u32 type = (u32)(v >> 56) & 0xf;
			if (type == T_OBJ || type == T_ARR)
				return false;
			out.slots[p] = v;
			out.set_mask |= 1u << p;
			break;
		}
	}
	out.flags = 0;
	return true;
}

} /* namespace dng */
