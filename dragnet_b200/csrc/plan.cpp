/*
 * plan.cpp: host-side plan compiler (plan JSON -> DevPlan).
 *
 * Restates, for the device, what the reference does when it assembles a scan:
 *   lib/stream-scan.js:56-86      stage order; dn_ts synthetic + time filter
 *   lib/datasource-file.js:154-163 datasource filter in front
 *   lib/dragnet-impl.js:66-125    decomps = breakdown NAMES; time-bounds filter
 *   lib/stream-synthetic.js:37-85 synthetic fields are assigned onto the
 *                                 record, so later stages see them first
 *   jsprim.pluck                  whole key first, then split at FIRST dot
 */
#include "plan.h"
#include "jsnum.cuh"
#include "../../include/dragnet_gpu.h"

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>

using namespace dng;

namespace {

/* ---- a very small JSON DOM (host only; plans are tiny) ---------------- */
struct JVal {
	enum K { NUL, BOOL, NUM, STR, ARR, OBJ } k = NUL;
	bool b = false;
	double num = 0;
	std::string str;
	std::vector<JVal> arr;
	std::vector<std::pair<std::string, JVal>> obj;	/* insertion order */

	const JVal *get(const char *key) const {
		const JVal *r = nullptr;
		for (auto &kv : obj)
			if (kv.first == key)
				r = &kv.second;	/* last duplicate wins */
		return r;
	}
};

struct JParser {
	const char *p, *end;
	std::string err;

	void ws() {
		while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' ||
		    *p == '\r'))
			p++;
	}
	bool fail(const char *m) {
		if (err.empty())
			err = m;
		return false;
	}
	static void utf8(std::string &s, unsigned cp) {
		if (cp < 0x80) {
			s += (char)cp;
		} else if (cp < 0x800) {
			s += (char)(0xC0 | (cp >> 6));
			s += (char)(0x80 | (cp & 0x3F));
		} else if (cp < 0x10000) {
			s += (char)(0xE0 | (cp >> 12));
			s += (char)(0x80 | ((cp >> 6) & 0x3F));
			s += (char)(0x80 | (cp & 0x3F));
		} else {
			s += (char)(0xF0 | (cp >> 18));
			s += (char)(0x80 | ((cp >> 12) & 0x3F));
			s += (char)(0x80 | ((cp >> 6) & 0x3F));
			s += (char)(0x80 | (cp & 0x3F));
		}
	}
	bool hex4(unsigned &v) {
		if (end - p < 4)
			return fail("bad \\u escape");
		v = 0;
		for (int i = 0; i < 4; i++) {
			int c = *p++, d;
			if (c >= '0' && c <= '9') d = c - '0';
			else if ((c | 0x20) >= 'a' && (c | 0x20) <= 'f')
				d = (c | 0x20) - 'a' + 10;
			else return fail("bad \\u escape");
			v = v * 16 + d;
		}
		return true;
	}
	bool string(std::string &out) {
		if (p >= end || *p != '"')
			return fail("expected string");
		p++;
		while (p < end && *p != '"') {
			unsigned char c = *p++;
			if (c < 0x20)
				return fail("control character in string");
			if (c != '\\') {
				out += (char)c;
				continue;
			}
			if (p >= end)
				return fail("unterminated string");
			char e = *p++;
			switch (e) {
			case '"': out += '"'; break;
			case '\\': out += '\\'; break;
			case '/': out += '/'; break;
			case 'b': out += '\b'; break;
			case 'f': out += '\f'; break;
			case 'n': out += '\n'; break;
			case 'r': out += '\r'; break;
			case 't': out += '\t'; break;
			case 'u': {
				unsigned cp, lo;
				if (!hex4(cp))
					return false;
				if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 &&
				    p[0] == '\\' && p[1] == 'u') {
					const char *save = p;
					p += 2;
					if (!hex4(lo))
						return false;
					if (lo >= 0xDC00 && lo < 0xE000)
						cp = 0x10000 + ((cp - 0xD800) << 10) +
						    (lo - 0xDC00);
					else
						p = save;
				}
				utf8(out, cp);
				break;
			}
			default:
				return fail("bad escape");
			}
		}
		if (p >= end)
			return fail("unterminated string");
		p++;
		return true;
	}
	bool value(JVal &v, int depth) {
		if (depth > 64)
			return fail("plan nested too deeply");
		ws();
		if (p >= end)
			return fail("unexpected end of plan");
		char c = *p;
		if (c == '{') {
			v.k = JVal::OBJ;
			p++;
			ws();
			if (p < end && *p == '}') { p++; return true; }
			for (;;) {
				ws();
				std::string key;
				if (!string(key))
					return false;
				ws();
				if (p >= end || *p != ':')
					return fail("expected ':'");
				p++;
				JVal child;
				if (!value(child, depth + 1))
					return false;
				v.obj.emplace_back(std::move(key), std::move(child));
				ws();
				if (p < end && *p == ',') { p++; continue; }
				if (p < end && *p == '}') { p++; return true; }
				return fail("expected ',' or '}'");
			}
		}
		if (c == '[') {
			v.k = JVal::ARR;
			p++;
			ws();
			if (p < end && *p == ']') { p++; return true; }
			for (;;) {
				JVal child;
				if (!value(child, depth + 1))
					return false;
				v.arr.push_back(std::move(child));
				ws();
				if (p < end && *p == ',') { p++; continue; }
				if (p < end && *p == ']') { p++; return true; }
				return fail("expected ',' or ']'");
			}
		}
		if (c == '"') {
			v.k = JVal::STR;
			return string(v.str);
		}
		if (end - p >= 4 && !strncmp(p, "true", 4)) {
			v.k = JVal::BOOL; v.b = true; p += 4; return true;
		}
		if (end - p >= 5 && !strncmp(p, "false", 5)) {
			v.k = JVal::BOOL; v.b = false; p += 5; return true;
		}
		if (end - p >= 4 && !strncmp(p, "null", 4)) {
			v.k = JVal::NUL; p += 4; return true;
		}
		const char *s = p;
		if (p < end && *p == '-') p++;
		if (p >= end || *p < '0' || *p > '9')
			return fail("unexpected token in plan");
		if (*p == '0') p++;
		else while (p < end && *p >= '0' && *p <= '9') p++;
		if (p < end && *p == '.') {
			p++;
			if (p >= end || *p < '0' || *p > '9')
				return fail("bad number");
			while (p < end && *p >= '0' && *p <= '9') p++;
		}
		if (p < end && (*p == 'e' || *p == 'E')) {
			p++;
			if (p < end && (*p == '+' || *p == '-')) p++;
			if (p >= end || *p < '0' || *p > '9')
				return fail("bad number");
			while (p < end && *p >= '0' && *p <= '9') p++;
		}
		v.k = JVal::NUM;
		v.num = dng_parse_decimal((const uint8_t *)s, (int)(p - s));
		return true;
	}
};

/* ---- fast-path automaton ------------------------------------------------- */

struct FastBuilder {
	FastTab &F;
	u16 *TR;		/* FAST_MAXSTATES x FAST_NCLS while building */
	int nstates = 10;	/* FS_* are fixed */
	int ncls = 0;

	enum { C_OTHER = 0, C_WS, C_NL, C_CTRL, C_QUOTE, C_BSLASH, C_LBRACE,
	    C_RBRACE, C_LBRACK, C_RBRACK, C_COLON, C_COMMA, C_MINUS, C_PLUS,
	    C_DOT, C_ZERO, C_DIGIT, C_a, C_b, C_cd, C_e, C_f, C_l, C_n, C_r, C_s,
	    C_t, C_u, C_AF, C_E, C_SLASH, C_WSC, C_BASE_COUNT };

	FastBuilder(FastTab &f, u16 *tr) : F(f), TR(tr) {}

	int alloc() { return nstates < FAST_MAXSTATES ? nstates++ : -1; }
	void set(int st, int cls, int next, int flags = 0) {
		TR[st * FAST_NCLS + cls] = (u16)(next | (flags << 8));
	}
	void set_all(int st, int next) {
		for (int c = 0; c < FAST_NCLS; c++)
			set(st, c, next);
	}
	static int base_class(int b) {
		if (b == ' ') return C_WS;
		if (b == '\t' || b == '\r') return C_WSC;	/* ws, but not inside strings */
		if (b == '\n') return C_NL;
		if (b < 0x20) return C_CTRL;
		switch (b) {
		case '"': return C_QUOTE; case '\\': return C_BSLASH;
		case '{': return C_LBRACE; case '}': return C_RBRACE;
		case '[': return C_LBRACK; case ']': return C_RBRACK;
		case ':': return C_COLON; case ',': return C_COMMA;
		case '-': return C_MINUS; case '+': return C_PLUS;
		case '.': return C_DOT; case '0': return C_ZERO;
		case 'a': return C_a; case 'b': return C_b;
		case 'c': case 'd': return C_cd;
		case 'e': return C_e; case 'f': return C_f; case 'l': return C_l;
		case 'n': return C_n; case 'r': return C_r; case 's': return C_s;
		case 't': return C_t; case 'u': return C_u;
		case 'A': case 'B': case 'C': case 'D': case 'F': return C_AF;
		case 'E': return C_E; case '/': return C_SLASH;
		}
		if (b >= '1' && b <= '9') return C_DIGIT;
		return C_OTHER;
	}

	bool build(const std::vector<std::string> &keys,
	    const std::vector<std::vector<int>> &key_term,	/* [ctx][key] */
	    const std::vector<std::vector<int>> &key_child) {
		memset(&F, 0, sizeof (F));
		memset(TR, 0, sizeof (u16) * FAST_MAXSTATES * FAST_NCLS);
		if (keys.size() > FAST_MAXKEYS || key_term.size() > FAST_MAXCTX)
			return false;
		/* classes: base classes, then one per distinct key byte */
		int base_of[FAST_NCLS];
		for (int c = 0; c < C_BASE_COUNT; c++)
			base_of[c] = c;
		ncls = C_BASE_COUNT;
		for (int b = 0; b < 256; b++)
			F.cls[b] = (u8)base_class(b);
		for (auto &k : keys) {
			for (unsigned char b : k) {
				if (b < 0x20 || b == '"' || b == '\\')
					continue;	/* only reachable escaped */
				if (F.cls[b] >= C_BASE_COUNT)
					continue;
				if (ncls >= FAST_NCLS)
					return false;
				base_of[ncls] = F.cls[b];
				F.cls[b] = (u8)ncls++;
			}
		}
		/* states */
		int OKs = alloc(), V[2] = { alloc(), alloc() };	/* V_O, V_A */
		int S[2], SE[2], SU[2][4], NM[2], N0[2], NI[2], ND[2], NF[2],
		    NE[2], NS[2], NX[2], T[2][3], Fs[2][4], L[2][3];
		for (int x = 0; x < 2; x++) {
			S[x] = alloc(); SE[x] = alloc();
			for (int i = 0; i < 4; i++) SU[x][i] = alloc();
			NM[x] = alloc(); N0[x] = alloc(); NI[x] = alloc();
			ND[x] = alloc(); NF[x] = alloc(); NE[x] = alloc();
			NS[x] = alloc(); NX[x] = alloc();
			for (int i = 0; i < 3; i++) T[x][i] = alloc();
			for (int i = 0; i < 4; i++) Fs[x][i] = alloc();
			for (int i = 0; i < 3; i++) L[x][i] = alloc();
		}
		int KX = alloc();
		/* trie over the candidate keys */
		struct Node { std::map<int, int> kids; int key = -1; int st; };
		std::vector<Node> trie(1);
		trie[0].st = alloc();
		for (size_t g = 0; g < keys.size(); g++) {
			bool reachable = true;
			for (unsigned char b : keys[g])
				if (b < 0x20 || b == '"' || b == '\\')
					reachable = false;
			if (!reachable)
				continue;
			int t = 0;
			for (unsigned char b : keys[g]) {
				auto it = trie[t].kids.find(b);
				if (it == trie[t].kids.end()) {
					Node n;
					n.st = alloc();
					if (n.st < 0)
						return false;
					trie.push_back(n);
					trie[t].kids[b] = (int)trie.size() - 1;
					t = (int)trie.size() - 1;
				} else {
					t = it->second;
				}
			}
			trie[t].key = (int)g;
		}
		if (nstates + (int)keys.size() > 250 || KX < 0 ||
		    trie[0].st < 0)
			return false;
		F.kc_base = (u8)nstates;	/* ids only, no rows */
		F.nstates = (u8)nstates;
		F.nkeys = (u8)keys.size();

		for (int st = 0; st < nstates; st++)
			set_all(st, FS_ERR);
		set_all(FS_FIN, FS_FIN);
		set_all(FS_FB, FS_FB);
		/* START: only containers take the fast path */
		set_all(FS_START, FS_FB);
		set(FS_START, C_WS, FS_START);
		set(FS_START, C_NL, FS_ERR);
		set(FS_START, C_LBRACE, FS_OF, FE_PUSH | FE_OBJ);
		set(FS_START, C_LBRACK, FS_AF, FE_PUSH);
		set(FS_DONE, C_WS, FS_DONE);
		set(FS_DONE, C_NL, FS_FIN);
		/* keys */
		set(FS_OF, C_WS, FS_OF);
		set(FS_OF, C_QUOTE, trie[0].st);
		set(FS_OF, C_RBRACE, FS_AFTER_O, FE_POP | FE_OBJ);
		set(OKs, C_WS, OKs);
		set(OKs, C_QUOTE, trie[0].st);
		auto key_row = [&](int st) {
			for (int c = 0; c < ncls; c++)
				set(st, c, KX);
			set(st, C_CTRL, FS_ERR);
			set(st, C_NL, FS_ERR);
			set(st, C_BSLASH, FS_FB);
			set(st, C_QUOTE, FS_KC);
		};
		key_row(KX);
		for (auto &n : trie) {
			key_row(n.st);
			for (auto &kv : n.kids)
				set(n.st, F.cls[kv.first], trie[kv.second].st);
			if (n.key >= 0)
				set(n.st, C_QUOTE, F.kc_base + n.key, FE_KEYHIT);
		}
		set(FS_KC, C_WS, FS_KC);
		set(FS_KC, C_COLON, V[0]);
		/* values */
		for (int x = 0; x < 2; x++) {
			int after = x == 0 ? FS_AFTER_O : FS_AFTER_A;
			int rows[2] = { V[x], x == 1 ? FS_AF : -1 };
			for (int ri = 0; ri < 2; ri++) {
				int st = rows[ri];
				if (st < 0)
					continue;
				set(st, C_WS, st);
				set(st, C_QUOTE, S[x], FE_VALSTART);
				set(st, C_LBRACE, FS_OF, FE_PUSH | FE_OBJ);
				set(st, C_LBRACK, FS_AF, FE_PUSH);
				set(st, C_MINUS, NM[x], FE_VALSTART);
				set(st, C_ZERO, N0[x], FE_VALSTART);
				set(st, C_DIGIT, NI[x], FE_VALSTART);
				set(st, C_t, T[x][0], FE_VALSTART);
				set(st, C_f, Fs[x][0], FE_VALSTART);
				set(st, C_n, L[x][0], FE_VALSTART);
			}
			/* strings */
			for (int c = 0; c < ncls; c++)
				set(S[x], c, S[x]);
			set(S[x], C_CTRL, FS_ERR);
			set(S[x], C_NL, FS_ERR);
			set(S[x], C_BSLASH, SE[x]);
			set(S[x], C_QUOTE, after, FE_VALEND_INCL);
			const int esc[] = { C_QUOTE, C_BSLASH, C_SLASH, C_b, C_f, C_n,
			    C_r, C_t };
			for (int c : esc)
				set(SE[x], c, S[x]);
			set(SE[x], C_u, SU[x][0]);
			const int hex[] = { C_ZERO, C_DIGIT, C_a, C_b, C_cd, C_e, C_f,
			    C_AF, C_E };
			for (int i = 0; i < 4; i++)
				for (int c : hex)
					set(SU[x][i], c, i == 3 ? S[x] : SU[x][i + 1]);
			/* numbers */
			set(NM[x], C_ZERO, N0[x]);
			set(NM[x], C_DIGIT, NI[x]);
			auto term = [&](int st) {
				set(st, C_WS, after, FE_VALEND_EXCL);
				set(st, C_COMMA, x == 0 ? OKs : V[1], FE_VALEND_EXCL);
				if (x == 0)
					set(st, C_RBRACE, FS_AFTER_O,
					    FE_POP | FE_OBJ | FE_VALEND_EXCL);
				else
					set(st, C_RBRACK, FS_AFTER_O,
					    FE_POP | FE_VALEND_EXCL);
			};
			set(N0[x], C_DOT, ND[x]);
			set(N0[x], C_e, NE[x]); set(N0[x], C_E, NE[x]);
			term(N0[x]);
			set(NI[x], C_ZERO, NI[x]); set(NI[x], C_DIGIT, NI[x]);
			set(NI[x], C_DOT, ND[x]);
			set(NI[x], C_e, NE[x]); set(NI[x], C_E, NE[x]);
			term(NI[x]);
			set(ND[x], C_ZERO, NF[x]); set(ND[x], C_DIGIT, NF[x]);
			set(NF[x], C_ZERO, NF[x]); set(NF[x], C_DIGIT, NF[x]);
			set(NF[x], C_e, NE[x]); set(NF[x], C_E, NE[x]);
			term(NF[x]);
			set(NE[x], C_PLUS, NS[x]); set(NE[x], C_MINUS, NS[x]);
			set(NE[x], C_ZERO, NX[x]); set(NE[x], C_DIGIT, NX[x]);
			set(NS[x], C_ZERO, NX[x]); set(NS[x], C_DIGIT, NX[x]);
			set(NX[x], C_ZERO, NX[x]); set(NX[x], C_DIGIT, NX[x]);
			term(NX[x]);
			/* literals */
			set(T[x][0], C_r, T[x][1]); set(T[x][1], C_u, T[x][2]);
			set(T[x][2], C_e, after, FE_VALEND_INCL);
			set(Fs[x][0], C_a, Fs[x][1]); set(Fs[x][1], C_l, Fs[x][2]);
			set(Fs[x][2], C_s, Fs[x][3]);
			set(Fs[x][3], C_e, after, FE_VALEND_INCL);
			set(L[x][0], C_u, L[x][1]); set(L[x][1], C_l, L[x][2]);
			set(L[x][2], C_l, after, FE_VALEND_INCL);
		}
		set(FS_AF, C_RBRACK, FS_AFTER_O, FE_POP);
		set(FS_AFTER_O, C_WS, FS_AFTER_O);
		set(FS_AFTER_O, C_COMMA, OKs);
		set(FS_AFTER_O, C_RBRACE, FS_AFTER_O, FE_POP | FE_OBJ);
		set(FS_AFTER_A, C_WS, FS_AFTER_A);
		set(FS_AFTER_A, C_COMMA, V[1]);
		set(FS_AFTER_A, C_RBRACK, FS_AFTER_O, FE_POP);
		/* tab and CR are whitespace between tokens but control
		 * characters inside strings and keys */
		for (int st = 0; st < nstates; st++)
			TR[st * FAST_NCLS + C_WSC] = TR[st * FAST_NCLS + C_WS];
		for (int x = 0; x < 2; x++)
			set(S[x], C_WSC, FS_ERR);
		set(KX, C_WSC, FS_ERR);
		for (auto &n : trie)
			set(n.st, C_WSC, FS_ERR);
		/* own classes of key bytes behave like their base class wherever
		 * the trie did not claim them */
		for (int c = C_BASE_COUNT; c < ncls; c++) {
			for (int st = 0; st < nstates; st++) {
				bool trie_edge = false;
				for (auto &n : trie)
					if (n.st == st)
						for (auto &kv : n.kids)
							if (F.cls[kv.first] == c)
								trie_edge = true;
				if (!trie_edge)
					TR[st * FAST_NCLS + c] =
					    TR[st * FAST_NCLS + base_of[c]];
			}
		}
		/* candidate map */
		memset(F.candmap, 0xFF, sizeof (F.candmap));
		for (size_t c = 0; c < key_term.size(); c++) {
			for (size_t g = 0; g < keys.size(); g++) {
				F.candmap[c * FAST_MAXKEYS + g][0] =
				    (u8)key_term[c][g];
				F.candmap[c * FAST_MAXKEYS + g][1] =
				    (u8)key_child[c][g];
			}
		}
		/* compact the rows to the classes actually used */
		int stride = (ncls + 3) & ~3;
		for (int st = 0; st < nstates; st++)
			for (int c = 0; c < stride; c++)
				TR[st * stride + c] = c < ncls ?
				    TR[st * FAST_NCLS + c] : (u16)FS_ERR;
		F.stride = (u8)stride;
		F.ok = 1;
		return true;
	}
};


/* ---- compiler --------------------------------------------------------- */

struct Compiler {
	DevPlan &P;
	std::string err;
	int code = DNG_OK;
	size_t pool_used = 0;

	struct PathRec { std::string field; int slot0; int ncomp; };
	std::vector<PathRec> paths;		/* unique field strings */
	std::map<std::string, int> ctx_by_prefix;
	struct CandRec { std::string key; int term_slot; int child_ctx; };
	std::vector<std::vector<CandRec>> ctx_cands;
	std::vector<int> ctx_parent, ctx_depth;
	std::vector<std::pair<int, int>> pathinfo;	/* (path idx, nlevels) */
	std::vector<std::string> syn_names;
	int fields_ctx = 0;

	explicit Compiler(DevPlan &p) : P(p) {}

	bool fail(int c, const std::string &m) {
		if (code == DNG_OK) {
			code = c;
			err = m;
		}
		return false;
	}

	int pool_add(const std::string &s) {
		if (pool_used + s.size() > POOL_BYTES) {
			fail(DNG_ELIMIT, "plan constant pool exhausted");
			return 0;
		}
		int off = (int)pool_used;
		memcpy(P.pool + pool_used, s.data(), s.size());
		pool_used += s.size();
		return off;
	}

	int new_ctx(const std::string &prefix, int parent, int depth) {
		auto it = ctx_by_prefix.find(prefix);
		if (it != ctx_by_prefix.end())
			return it->second;
		int id = (int)ctx_cands.size();
		ctx_by_prefix[prefix] = id;
		ctx_cands.emplace_back();
		ctx_parent.push_back(parent);
		ctx_depth.push_back(depth);
		return id;
	}

	CandRec &cand(int ctx, const std::string &key) {
		for (auto &c : ctx_cands[ctx])
			if (c.key == key)
				return c;
		ctx_cands[ctx].push_back(CandRec{key, -1, -1});
		return ctx_cands[ctx].back();
	}

	/* register a JSON field path; returns index into `paths` */
	int add_path(const std::string &field) {
		for (size_t i = 0; i < paths.size(); i++)
			if (paths[i].field == field)
				return (int)i;
		std::vector<std::string> comps;
		size_t a = 0;
		for (;;) {
			size_t d = field.find('.', a);
			if (d == std::string::npos) {
				comps.push_back(field.substr(a));
				break;
			}
			comps.push_back(field.substr(a, d - a));
			a = d + 1;
		}
		int L = (int)comps.size();
		if (L > MAX_LEVELS) {
			fail(DNG_ELIMIT, "field \"" + field + "\" has too many "
			    "dot-separated components");
			return 0;
		}
		int slot0 = 0;
		for (auto &p : paths)
			slot0 += p.ncomp;
		if (slot0 + L > MAX_SLOTS - 2) {
			fail(DNG_ELIMIT, "too many distinct fields in one scan");
			return 0;
		}
		/* level l: inside context for prefix comps[0..l) */
		std::string prefix = "\x01";	/* fields root marker */
		int ctx = fields_ctx;
		size_t pos = 0;
		for (int l = 0; l < L; l++) {
			std::string suffix = field.substr(pos);
			cand(ctx, suffix).term_slot = slot0 + l;
			if (l < L - 1) {
				prefix += comps[l];
				prefix += '\x02';
				int child = new_ctx(prefix, ctx, ctx_depth[ctx] + 1);
				cand(ctx, comps[l]).child_ctx = child;
				ctx = child;
			}
			pos += comps[l].size() + 1;
		}
		paths.push_back(PathRec{field, slot0, L});
		return (int)paths.size() - 1;
	}

	int add_pathinfo(int pidx, int nlevels) {
		for (size_t i = 0; i < pathinfo.size(); i++)
			if (pathinfo[i].first == pidx &&
			    pathinfo[i].second == nlevels)
				return (int)i;
		pathinfo.emplace_back(pidx, nlevels);
		return (int)pathinfo.size() - 1;
	}

	/*
	 * Where does pluck(record, field) read from when the first `nsyn`
	 * synthetic fields have been assigned onto the record?
	 */
	Src resolve(const std::string &field, int nsyn) {
		Src s;
		s.kind = SRC_UNDEF;
		s.idx = 0;
		for (int j = nsyn - 1; j >= 0; j--) {
			if (syn_names[j] == field) {
				s.kind = SRC_SYNTH;
				s.idx = (u8)j;
				return s;
			}
		}
		int pidx = add_path(field);
		int nlevels = paths.size() ? paths[pidx].ncomp : 0;
		size_t d = field.find('.');
		if (d != std::string::npos) {
			std::string k1 = field.substr(0, d);
			for (int j = 0; j < nsyn; j++)
				if (syn_names[j] == k1)
					nlevels = 1;	/* number has no props */
		}
		s.kind = SRC_PATH;
		s.idx = (u8)add_pathinfo(pidx, nlevels);
		return s;
	}

	/* krill predicate -> jump code; returns entry index or -1 if trivial */
	struct Patch { int leaf; bool on_true; };

	bool emit(const JVal &pred, int nsyn, std::vector<Leaf> &prog,
	    std::vector<Patch> &tlist, std::vector<Patch> &flist) {
		if (pred.k != JVal::OBJ)
			return fail(DNG_EINVAL, "predicate is not an object");
		if (pred.obj.empty()) {
			Leaf lf;
			memset(&lf, 0, sizeof (lf));
			lf.op = OP_TRUE;
			prog.push_back(lf);
			tlist.push_back(Patch{(int)prog.size() - 1, true});
			flist.push_back(Patch{(int)prog.size() - 1, false});
			return true;
		}
		if (pred.obj.size() != 1)
			return fail(DNG_EINVAL, "predicate: expected exactly one key");
		const std::string &key = pred.obj[0].first;
		const JVal &args = pred.obj[0].second;
		if (key == "and" || key == "or") {
			bool is_and = key == "and";
			if (args.k != JVal::ARR || args.arr.empty())
				return fail(DNG_EINVAL, "predicate: \"" + key +
				    "\" requires a non-empty array");
			for (size_t i = 0; i < args.arr.size(); i++) {
				std::vector<Patch> t, f;
				int start = (int)prog.size();
				if (!emit(args.arr[i], nsyn, prog, t, f))
					return false;
				bool last = i + 1 == args.arr.size();
				(void)start;
				if (is_and) {
					/* false anywhere -> whole false */
					flist.insert(flist.end(), f.begin(), f.end());
					if (last)
						tlist.insert(tlist.end(), t.begin(), t.end());
					else
						patch(prog, t, (int)prog.size());
				} else {
					tlist.insert(tlist.end(), t.begin(), t.end());
					if (last)
						flist.insert(flist.end(), f.begin(), f.end());
					else
						patch(prog, f, (int)prog.size());
				}
			}
			return true;
		}
		static const char *ops[] = { "eq", "ne", "lt", "le", "gt", "ge" };
		int op = -1;
		for (int i = 0; i < 6; i++)
			if (key == ops[i])
				op = i;
		if (op < 0)
			return fail(DNG_EINVAL, "predicate: unknown operator \"" +
			    key + "\"");
		if (args.k != JVal::ARR || args.arr.size() != 2 ||
		    args.arr[0].k != JVal::STR)
			return fail(DNG_EINVAL, "predicate: \"" + key +
			    "\" requires [field, constant]");
		const JVal &c = args.arr[1];
		Leaf lf;
		memset(&lf, 0, sizeof (lf));
		lf.op = (u8)op;
		lf.src = resolve(args.arr[0].str, nsyn);
		if (c.k == JVal::STR) {
			lf.cstr = 1;
			lf.coff = (u16)pool_add(c.str);
			lf.clen = (u16)c.str.size();
			lf.cnum = dng_string_to_number(
			    (const uint8_t *)c.str.data(), (int)c.str.size());
		} else if (c.k == JVal::NUM) {
			lf.cnum = c.num;
		} else if (c.k == JVal::BOOL) {
			lf.cnum = c.b ? 1.0 : 0.0;
		} else {
			return fail(DNG_EINVAL, "predicate: constant must be a "
			    "string, number or boolean");
		}
		prog.push_back(lf);
		tlist.push_back(Patch{(int)prog.size() - 1, true});
		flist.push_back(Patch{(int)prog.size() - 1, false});
		return true;
	}

	static void patch(std::vector<Leaf> &prog, std::vector<Patch> &l,
	    int target) {
		for (auto &p : l) {
			if (p.on_true)
				prog[p.leaf].jt = (int16_t)target;
			else
				prog[p.leaf].jf = (int16_t)target;
		}
		l.clear();
	}

	int compile_pred(const JVal *pred, int nsyn, std::vector<Leaf> &prog) {
		if (!pred || pred->k == JVal::NUL)
			return -1;
		if (pred->k == JVal::OBJ && pred->obj.empty())
			return -1;	/* {} is always true */
		int entry = (int)prog.size();
		std::vector<Patch> t, f;
		if (!emit(*pred, nsyn, prog, t, f))
			return -1;
		patch(prog, t, -1);
		patch(prog, f, -2);
		return entry;
	}

	bool run(const JVal &root, dng_plan *out) {
		memset(&P, 0, sizeof (P));
		if (root.k != JVal::OBJ)
			return fail(DNG_EINVAL, "plan must be a JSON object");
		const JVal *fmt = root.get("format");
		P.format = FMT_JSON;
		if (fmt && fmt->k == JVal::STR) {
			if (fmt->str == "json-skinner")
				P.format = FMT_SKINNER;
			else if (fmt->str != "json")
				return fail(DNG_EINVAL, "unsupported format: \"" +
				    fmt->str + "\"");
		}
		P.sk_fields_slot = P.sk_value_slot = -1;
		if (P.format == FMT_SKINNER) {
			/* envelope {fields:{...}, value:N}: lib/format-json.js:55-73 */
			int env = new_ctx("", -1, 1);
			fields_ctx = new_ctx("\x01", env, 2);
			P.sk_fields_slot = MAX_SLOTS - 2;
			P.sk_value_slot = MAX_SLOTS - 1;
			CandRec &cf = cand(env, "fields");
			cf.term_slot = P.sk_fields_slot;
			cf.child_ctx = fields_ctx;
			cand(env, "value").term_slot = P.sk_value_slot;
		} else {
			fields_ctx = new_ctx("\x01", -1, 1);
		}
		P.root_ctx = (int8_t)fields_ctx;

		std::vector<Leaf> prog;
		P.ds_entry = (int16_t)compile_pred(root.get("ds_filter"), 0, prog);
		if (this->code != DNG_OK)
			return false;

		/* one metric (a plain scan: the plan object itself) or several
		 * ("metrics": [...], the fan-out of dn build / index-scan) */
		std::vector<const JVal *> mets;
		const JVal *ml = root.get("metrics");
		if (ml && ml->k == JVal::ARR) {
			for (auto &m : ml->arr)
				mets.push_back(&m);
			if (mets.empty() || mets.size() > MAX_METRICS)
				return fail(DNG_ELIMIT, "\"metrics\" must hold 1.."
				    + std::to_string((int)MAX_METRICS) + " entries");
		} else {
			mets.push_back(&root);
		}
		P.nmetrics = (u8)mets.size();
		out->nmetrics = (int)mets.size();
		int nsyn_total = 0, ncols_total = 0;
		for (size_t mi = 0; mi < mets.size(); mi++) {
			const JVal &mj = *mets[mi];
			if (mj.k != JVal::OBJ)
				return fail(DNG_EINVAL, "metric must be an object");
			Metric &M = P.metric[mi];
			/* synthetic names are private to a metric's StreamScan */
			syn_names.assign((size_t)nsyn_total, std::string("\x01"));
			M.user_entry = (int16_t)compile_pred(mj.get("filter"), 0,
			    prog);
			if (this->code != DNG_OK)
				return false;
			M.syn0 = (u8)nsyn_total;
			const JVal *syn = mj.get("synthetic");
			if (syn && syn->k == JVal::ARR) {
				for (auto &sj : syn->arr) {
					const JVal *n = sj.get("name"),
					    *f = sj.get("field");
					if (!n || !f || n->k != JVal::STR ||
					    f->k != JVal::STR)
						return fail(DNG_EINVAL, "synthetic: "
						    "need string \"name\" and "
						    "\"field\"");
					if (nsyn_total >= MAX_SYN)
						return fail(DNG_ELIMIT, "too many "
						    "synthetic fields");
					P.syn[nsyn_total] = resolve(f->str,
					    nsyn_total);
					syn_names.push_back(n->str);
					nsyn_total++;
				}
			}
			M.nsyn = (u8)(nsyn_total - M.syn0);

			M.time_entry = -1;
			const JVal *tb = mj.get("time_bounds");
			if (tb && tb->k == JVal::OBJ) {
				const JVal *f = tb->get("field"), *ge = tb->get("ge"),
				    *lt = tb->get("lt");
				if (!f || f->k != JVal::STR || !ge || !lt ||
				    ge->k != JVal::NUM || lt->k != JVal::NUM)
					return fail(DNG_EINVAL, "time_bounds: need "
					    "field, ge, lt");
				/* {and:[{ge:[f,a]},{lt:[f,b]}]}:
				 * lib/dragnet-impl.js:108-119 */
				M.time_entry = (int16_t)prog.size();
				Leaf a, b;
				memset(&a, 0, sizeof (a));
				memset(&b, 0, sizeof (b));
				a.op = OP_GE;
				a.src = resolve(f->str, nsyn_total);
				a.cnum = ge->num;
				a.jt = (int16_t)(prog.size() + 1);
				a.jf = -2;
				b.op = OP_LT;
				b.src = a.src;
				b.cnum = lt->num;
				b.jt = -1;
				b.jf = -2;
				prog.push_back(a);
				prog.push_back(b);
			}

			const JVal *bds = mj.get("breakdowns");
			if (!bds || bds->k != JVal::ARR)
				return fail(DNG_EINVAL, "plan: \"breakdowns\" must "
				    "be an array");
			if (bds->arr.size() > 12)
				return fail(DNG_ELIMIT, "too many breakdowns");
			M.col0 = (u8)ncols_total;
			int nc = 0;
			for (auto &bj : bds->arr) {
				const JVal *n = bj.get("name");
				if (!n || n->k != JVal::STR)
					return fail(DNG_EINVAL, "breakdown without "
					    "a name");
				if (ncols_total >= MAX_COLS)
					return fail(DNG_ELIMIT, "too many "
					    "breakdown columns in one scan");
				Col &c = P.col[ncols_total];
				/* the aggregator plucks the breakdown NAME
				 * (lib/dragnet-impl.js:79-80) */
				c.src = resolve(n->str, nsyn_total);
				c.kind = COL_DISCRETE;
				c.step = 0;
				const JVal *ag = bj.get("aggr");
				if (ag && ag->k == JVal::STR) {
					if (ag->str == "quantize") {
						c.kind = COL_P2;
					} else if (ag->str == "lquantize") {
						const JVal *st = bj.get("step");
						if (!st || st->k != JVal::NUM)
							return fail(DNG_EINVAL,
							    "aggr \"lquantize\" "
							    "requires \"step\"");
						c.kind = COL_LINEAR;
						c.step = st->num;
					} else {
						return fail(DNG_EINVAL,
						    "unsupported aggr: \"" +
						    ag->str + "\"");
					}
				}
				out->col_kind[mi][nc] = c.kind;
				out->col_step[mi][nc] = c.step;
				nc++;
				ncols_total++;
			}
			M.ncols = (u8)nc;
			out->ncols[mi] = nc;
		}
		if (this->code != DNG_OK)
			return false;

		/* ---- flatten ---- */
		if (prog.size() > MAX_CODE)
			return fail(DNG_ELIMIT, "filter too large");
		P.ncode = (u8)prog.size();
		for (size_t i = 0; i < prog.size(); i++)
			P.code[i] = prog[i];
		if (ctx_cands.size() > MAX_CTX)
			return fail(DNG_ELIMIT, "too many nested field contexts");
		P.hot.nctx = (u8)ctx_cands.size();
		int nc = 0;
		for (size_t c = 0; c < ctx_cands.size(); c++) {
			Ctx &x = P.hot.ctx[c];
			x.parent = (int8_t)ctx_parent[c];
			x.depth = (u8)ctx_depth[c];
			x.cand_begin = (u16)nc;
			x.bloom = 0;
			x.arraylike = 0;
			for (auto &cr : ctx_cands[c]) {
				bool idx = !cr.key.empty() &&
				    (cr.key == "0" || cr.key[0] != '0');
				for (unsigned char ch : cr.key)
					if (ch < '0' || ch > '9')
						idx = false;
				if (idx || cr.key == "length")
					x.arraylike = 1;
				if (nc >= MAX_CANDS)
					return fail(DNG_ELIMIT, "too many field "
					    "name candidates");
				Cand &d = P.cand[nc++];
				d.off = (u16)pool_add(cr.key);
				d.len = (u16)cr.key.size();
				u32 h = 2166136261u;
				for (unsigned char ch : cr.key)
					h = (h ^ ch) * 16777619u;
				d.hash = h;
				d.term_slot = (int8_t)cr.term_slot;
				d.child_ctx = (int8_t)cr.child_ctx;
				x.bloom |= 1ull << (h & 63);
			}
			x.cand_end = (u16)nc;
		}
		P.ncand = (u8)nc;
		/* subtree masks: slots reachable at or below each context */
		for (int c = (int)ctx_cands.size() - 1; c >= 0; c--) {
			u32 m = 0;
			for (auto &cr : ctx_cands[c]) {
				if (cr.term_slot >= 0)
					m |= 1u << cr.term_slot;
				if (cr.child_ctx >= 0)
					m |= P.hot.ctx[cr.child_ctx].subtree_mask;
			}
			P.hot.ctx[c].subtree_mask = m;
		}
		if (pathinfo.size() > MAX_PATHS)
			return fail(DNG_ELIMIT, "too many distinct fields");
		P.npaths = (u8)pathinfo.size();
		int nslots = 0;
		for (auto &p : paths)
			nslots += p.ncomp;
		P.nslots = (u8)nslots;
		for (size_t i = 0; i < pathinfo.size(); i++) {
			P.path[i].slot0 = (u8)paths[pathinfo[i].first].slot0;
			P.path[i].nlevels = (u8)pathinfo[i].second;
		}
		/* fast-path automaton (optional: plans it cannot express simply
		 * keep using the general parser) */
		{
			std::vector<std::string> keys;
			bool arraylike = false;
			for (size_t c = 0; c < ctx_cands.size(); c++) {
				if (P.hot.ctx[c].arraylike)
					arraylike = true;
				for (auto &cr : ctx_cands[c])
					if (std::find(keys.begin(), keys.end(),
					    cr.key) == keys.end())
						keys.push_back(cr.key);
			}
			std::vector<std::vector<int>> kt(ctx_cands.size()),
			    kc(ctx_cands.size());
			for (size_t c = 0; c < ctx_cands.size(); c++) {
				kt[c].assign(keys.size(), -1);
				kc[c].assign(keys.size(), -1);
				for (auto &cr : ctx_cands[c]) {
					size_t g = std::find(keys.begin(),
					    keys.end(), cr.key) - keys.begin();
					kt[c][g] = cr.term_slot;
					kc[c][g] = cr.child_ctx;
				}
			}
			FastBuilder fb(P.hot.fast, P.hot.trans);
			if (arraylike || !fb.build(keys, kt, kc))
				P.hot.fast.ok = 0;
		}
		return this->code == DNG_OK;
	}
};


} /* namespace */

int dng_plan_compile(const char *json, dng_plan *out, char *err,
    unsigned long errlen)
{
	JParser jp;
	jp.p = json;
	jp.end = json + strlen(json);
	JVal root;
	if (!jp.value(root, 0)) {
		if (err && errlen)
			snprintf(err, errlen, "invalid plan JSON: %s",
			    jp.err.c_str());
		return DNG_EINVAL;
	}
	jp.ws();
	if (jp.p != jp.end) {
		if (err && errlen)
			snprintf(err, errlen, "invalid plan JSON: trailing data");
		return DNG_EINVAL;
	}
	Compiler c(out->dev);
	if (!c.run(root, out)) {
		if (err && errlen)
			snprintf(err, errlen, "%s", c.err.c_str());
		return c.code != DNG_OK ? c.code : DNG_EINVAL;
	}
	if (err && errlen)
		err[0] = '\0';
	return DNG_OK;
}
