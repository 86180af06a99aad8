/*
 * dn_oracle.cpp: CPU oracle (C++17) -- a plain, DOM-based restatement of
 * dragnet's raw-data scan path.  TEST INFRASTRUCTURE AND CPU BASELINE ONLY:
 * nothing under dragnet_b200/ links, calls or executes this; only tests/,
 * __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) do.
 *
 * It follows the reference stage by stage (files under /root/reference):
 *   lib/format-json.js:26-98        lines -> JSON.parse -> {fields, value:1}
 *   lib/datasource-file.js:154-163  datasource filter first
 *   lib/stream-scan.js:56-86        user filter -> synthetic -> time filter
 *                                   -> aggregator
 *   lib/krill-skinner-stream.js:29-52  pass / nfilteredout / nfailedeval
 *   lib/stream-synthetic.js:37-85   date fields: number passthrough,
 *                                   floor(Date.parse/1000), undef / baddate
 *   lib/dragnet-impl.js:48-125      decomps = breakdown names; time bounds
 *   lib/dragnet.js:52-71            P2 / linear bucketizers
 * and restates the un-vendored npm modules it calls (lstream@0.0.4,
 * vstream-json-parser@1.0.0, krill@^1.0.0, jsprim@^1.3.0 pluck,
 * skinner#dragnet) as ordinary ECMAScript semantics.  Deliberately written
 * differently from the GPU code: it materialises every record as a tree
 * exactly like JSON.parse does, then plucks/compares/stringifies on that tree,
 * and leans on glibc (strtod, snprintf) for number conversions.
 *
 * Pinned by the reference's golden outputs through tests/test_oracle_cpp.py
 * (same harness as the Python oracle) and cross-checked against
 * oracle/dn_oracle.py on the edge-case corpus.
 *
 *   dn_oracle PLAN.json [--threads N] [--repeat R] FILE...
 * prints one JSON document: points, counters, and the best wall time of R
 * scans of the (already in memory) input.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <new>

/*
 * Per-line bump arena.  This restatement materialises every record as a tree
 * (as JSON.parse does); with the default allocator its threads spend their
 * time contending inside malloc/free and the "all host threads" baseline
 * swings by 5x from box to box.  While a line is being parsed and staged,
 * operator new hands out memory from a thread-local arena that is reset at the
 * next line (operator delete ignores arena pointers); tally-table nodes, which
 * outlive the line, are allocated with the arena switched off.
 */
namespace arena {
constexpr size_t BYTES = 4u << 20;
thread_local char *base = nullptr, *cur = nullptr, *end = nullptr;
thread_local bool on = false;
inline void begin()
{
	if (!base) {
		base = (char *)malloc(BYTES);
		end = base ? base + BYTES : nullptr;
	}
	cur = base;
	on = base != nullptr;
}
inline void off() { on = false; }
}

void *operator new(size_t n)
{
	if (arena::on) {
		const size_t r = (n + 15) & ~(size_t)15;
		if ((size_t)(arena::end - arena::cur) >= r) {
			void *p = arena::cur;
			arena::cur += r;
			return p;
		}
	}
	void *p = malloc(n ? n : 1);
	if (!p)
		throw std::bad_alloc();
	return p;
}

void operator delete(void *p) noexcept
{
	if (p >= (void *)arena::base && p < (void *)arena::end)
		return;
	free(p);
}

void operator delete(void *p, size_t) noexcept
{
	operator delete(p);
}

namespace {

/* ---- JS values ----------------------------------------------------------- */

struct JV;
typedef std::shared_ptr<JV> JP;

struct JV {
	enum T { UNDEF, NUL, BOOL, NUM, STR, OBJ, ARR } t = UNDEF;
	bool b = false;
	double num = 0;
	std::string str;
	std::vector<std::pair<std::string, JP>> obj;
	std::vector<JP> arr;
};

JP mk(JV::T t)
{
	JP p = std::make_shared<JV>();
	p->t = t;
	return p;
}

JP mknum(double d)
{
	JP p = mk(JV::NUM);
	p->num = d;
	return p;
}

const JP UNDEFINED = mk(JV::UNDEF);

/* ---- JSON.parse ----------------------------------------------------------- */

struct Parser {
	const unsigned char *p, *e;
	bool ok = true;

	void ws() {
		while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' ||
		    *p == '\r'))
			p++;
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
		if (e - p < 4)
			return false;
		v = 0;
		for (int i = 0; i < 4; i++) {
			int c = *p++, d;
			if (c >= '0' && c <= '9') d = c - '0';
			else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
			else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
			else return false;
			v = v * 16 + d;
		}
		return true;
	}
	bool string(std::string &out) {
		if (p >= e || *p != '"')
			return false;
		p++;
		for (;;) {
			if (p >= e)
				return false;
			unsigned char c = *p++;
			if (c == '"')
				return true;
			if (c < 0x20)
				return false;
			if (c != '\\') {
				out += (char)c;
				continue;
			}
			if (p >= e)
				return false;
			switch (*p++) {
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
				if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 6 &&
				    p[0] == '\\' && p[1] == 'u') {
					const unsigned char *save = p;
					p += 2;
					if (hex4(lo) && lo >= 0xDC00 &&
					    lo < 0xE000)
						cp = 0x10000 + ((cp - 0xD800) <<
						    10) + (lo - 0xDC00);
					else
						p = save;
				}
				utf8(out, cp);
				break;
			}
			default:
				return false;
			}
		}
	}
	JP value(int depth) {
		ws();
		if (p >= e || depth > 5000) {
			ok = false;
			return nullptr;
		}
		unsigned char c = *p;
		if (c == '{') {
			JP o = mk(JV::OBJ);
			p++;
			ws();
			if (p < e && *p == '}') {
				p++;
				return o;
			}
			for (;;) {
				ws();
				std::string key;
				if (!string(key)) { ok = false; return nullptr; }
				ws();
				if (p >= e || *p != ':') { ok = false; return nullptr; }
				p++;
				JP v = value(depth + 1);
				if (!ok)
					return nullptr;
				bool found = false;
				for (auto &kv : o->obj) {
					if (kv.first == key) {	/* last wins */
						kv.second = v;
						found = true;
						break;
					}
				}
				if (!found)
					o->obj.emplace_back(std::move(key), v);
				ws();
				if (p < e && *p == ',') { p++; continue; }
				if (p < e && *p == '}') { p++; return o; }
				ok = false;
				return nullptr;
			}
		}
		if (c == '[') {
			JP a = mk(JV::ARR);
			p++;
			ws();
			if (p < e && *p == ']') {
				p++;
				return a;
			}
			for (;;) {
				JP v = value(depth + 1);
				if (!ok)
					return nullptr;
				a->arr.push_back(v);
				ws();
				if (p < e && *p == ',') { p++; continue; }
				if (p < e && *p == ']') { p++; return a; }
				ok = false;
				return nullptr;
			}
		}
		if (c == '"') {
			JP s = mk(JV::STR);
			if (!string(s->str))
				ok = false;
			return s;
		}
		if (e - p >= 4 && !memcmp(p, "true", 4)) {
			JP b = mk(JV::BOOL);
			b->b = true;
			p += 4;
			return b;
		}
		if (e - p >= 5 && !memcmp(p, "false", 5)) {
			JP b = mk(JV::BOOL);
			p += 5;
			return b;
		}
		if (e - p >= 4 && !memcmp(p, "null", 4)) {
			p += 4;
			return mk(JV::NUL);
		}
		const unsigned char *s = p;
		if (p < e && *p == '-')
			p++;
		if (p >= e || *p < '0' || *p > '9') { ok = false; return nullptr; }
		if (*p == '0')
			p++;
		else
			while (p < e && *p >= '0' && *p <= '9')
				p++;
		if (p < e && *p == '.') {
			p++;
			if (p >= e || *p < '0' || *p > '9') { ok = false; return nullptr; }
			while (p < e && *p >= '0' && *p <= '9')
				p++;
		}
		if (p < e && (*p == 'e' || *p == 'E')) {
			p++;
			if (p < e && (*p == '+' || *p == '-'))
				p++;
			if (p >= e || *p < '0' || *p > '9') { ok = false; return nullptr; }
			while (p < e && *p >= '0' && *p <= '9')
				p++;
		}
		std::string txt((const char *)s, p - s);
		return mknum(strtod(txt.c_str(), nullptr));
	}
};

/* JSON.parse(line); nullptr when it throws */
JP json_parse(const unsigned char *s, size_t n)
{
	Parser ps;
	ps.p = s;
	ps.e = s + n;
	JP v = ps.value(0);
	if (!ps.ok)
		return nullptr;
	ps.ws();
	if (ps.p != ps.e)
		return nullptr;
	return v;
}

/* ---- ECMAScript conversions ----------------------------------------------- */

std::string num_to_string(double v)
{
	if (v != v)
		return "NaN";
	if (v == 0)
		return "0";
	if (std::isinf(v))
		return v > 0 ? "Infinity" : "-Infinity";
	std::string sign;
	if (v < 0) {
		sign = "-";
		v = -v;
	}
	/* shortest digits that round-trip */
	char buf[40];
	int prec = 1;
	for (; prec <= 17; prec++) {
		snprintf(buf, sizeof (buf), "%.*e", prec - 1, v);
		if (strtod(buf, nullptr) == v)
			break;
	}
	/* buf = d.ddddde[+-]XX */
	std::string digits;
	int exp10 = 0;
	{
		char *ep = strchr(buf, 'e');
		exp10 = atoi(ep + 1);
		for (char *q = buf; q < ep; q++)
			if (*q != '.')
				digits += *q;
		while (digits.size() > 1 && digits.back() == '0')
			digits.pop_back();
	}
	int k = (int)digits.size();
	int n = exp10 + 1;
	std::string out;
	if (k <= n && n <= 21) {
		out = digits + std::string(n - k, '0');
	} else if (0 < n && n <= 21) {
		out = digits.substr(0, n) + "." + digits.substr(n);
	} else if (-6 < n && n <= 0) {
		out = "0." + std::string(-n, '0') + digits;
	} else {
		int e = n - 1;
		out = digits.substr(0, 1);
		if (k > 1)
			out += "." + digits.substr(1);
		out += e >= 0 ? "e+" : "e-";
		out += std::to_string(e >= 0 ? e : -e);
	}
	return sign + out;
}

const double NaN = std::nan("");

size_t js_space(const std::string &s, size_t i)
{
	unsigned char c = s[i];
	if (c == ' ' || (c >= 9 && c <= 13))
		return 1;
	if (c == 0xC2 && i + 1 < s.size() && (unsigned char)s[i + 1] == 0xA0)
		return 2;
	if ((c & 0xF0) == 0xE0 && i + 2 < s.size()) {
		unsigned cp = ((c & 0x0F) << 12) |
		    (((unsigned char)s[i + 1] & 0x3F) << 6) |
		    ((unsigned char)s[i + 2] & 0x3F);
		if (cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
		    cp == 0x2028 || cp == 0x2029 || cp == 0x202F ||
		    cp == 0x205F || cp == 0x3000 || cp == 0xFEFF)
			return 3;
	}
	return 0;
}

double string_to_number(const std::string &in)
{
	size_t a = 0, b = in.size(), l;
	while (a < b && (l = js_space(in, a)) > 0)
		a += l;
	for (;;) {
		bool cut = false;
		for (size_t w = 1; w <= 3 && b >= a + w; w++) {
			if (js_space(in, b - w) == w) {
				b -= w;
				cut = true;
				break;
			}
		}
		if (!cut)
			break;
	}
	std::string s = in.substr(a, b - a);
	if (s.empty())
		return 0;
	if (s.size() > 2 && s[0] == '0' && strchr("xXoObB", s[1])) {
		int base = (s[1] | 0x20) == 'x' ? 16 : (s[1] | 0x20) == 'o' ? 8 : 2;
		long double acc = 0;	/* exact enough? no: do it in binary */
		uint64_t m = 0;
		int extra = 0, sticky = 0;
		int bpd = base == 16 ? 4 : base == 8 ? 3 : 1;
		(void)acc;
		for (size_t i = 2; i < s.size(); i++) {
			int c = (unsigned char)s[i], d;
			if (c >= '0' && c <= '9') d = c - '0';
			else if ((c | 0x20) >= 'a' && (c | 0x20) <= 'f')
				d = (c | 0x20) - 'a' + 10;
			else return NaN;
			if (d >= base)
				return NaN;
			for (int bit = bpd - 1; bit >= 0; bit--) {
				int v = (d >> bit) & 1;
				if (m >> 63) {
					extra++;
					sticky |= v;
				} else {
					m = (m << 1) | v;
				}
			}
		}
		if (extra == 0)
			return (double)m;	/* correctly rounded by the FPU */
		/* round m (64 bits) * 2^extra with sticky: to 53 bits */
		uint64_t top = m >> 11, rem = m & 0x7FF;
		if (rem > 0x400 || (rem == 0x400 && (sticky || (top & 1))))
			top++;
		return std::ldexp((double)top, 11 + extra);
	}
	size_t i = 0;
	if (s[i] == '+' || s[i] == '-')
		i++;
	if (s.compare(i, std::string::npos, "Infinity") == 0)
		return s[0] == '-' ? -INFINITY : INFINITY;
	size_t nd = 0;
	while (i < s.size() && isdigit((unsigned char)s[i])) { i++; nd++; }
	if (i < s.size() && s[i] == '.') {
		i++;
		while (i < s.size() && isdigit((unsigned char)s[i])) { i++; nd++; }
	}
	if (nd == 0)
		return NaN;
	if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
		i++;
		if (i < s.size() && (s[i] == '+' || s[i] == '-'))
			i++;
		size_t ed = 0;
		while (i < s.size() && isdigit((unsigned char)s[i])) { i++; ed++; }
		if (ed == 0)
			return NaN;
	}
	if (i != s.size())
		return NaN;
	return strtod(s.c_str(), nullptr);
}

std::string to_string(const JP &v)
{
	switch (v->t) {
	case JV::UNDEF: return "undefined";
	case JV::NUL: return "null";
	case JV::BOOL: return v->b ? "true" : "false";
	case JV::NUM: return num_to_string(v->num);
	case JV::STR: return v->str;
	case JV::OBJ: return "[object Object]";
	case JV::ARR: {
		std::string out;
		for (size_t i = 0; i < v->arr.size(); i++) {
			if (i)
				out += ",";
			const JP &x = v->arr[i];
			if (x->t != JV::NUL && x->t != JV::UNDEF)
				out += to_string(x);
		}
		return out;
	}
	}
	return "";
}

double to_number(const JP &v)
{
	switch (v->t) {
	case JV::UNDEF: return NaN;
	case JV::NUL: return 0;
	case JV::BOOL: return v->b ? 1 : 0;
	case JV::NUM: return v->num;
	case JV::STR: return string_to_number(v->str);
	default: return string_to_number(to_string(v));
	}
}

/* UTF-8 -> UTF-16 code units (lone surrogates pass through) */
std::vector<uint16_t> utf16(const std::string &s)
{
	std::vector<uint16_t> out;
	size_t i = 0;
	while (i < s.size()) {
		unsigned char c = s[i];
		unsigned cp;
		if (c < 0x80) { cp = c; i += 1; }
		else if ((c & 0xE0) == 0xC0 && i + 1 < s.size()) {
			cp = ((c & 0x1F) << 6) | (s[i + 1] & 0x3F);
			i += 2;
		} else if ((c & 0xF0) == 0xE0 && i + 2 < s.size()) {
			cp = ((c & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) |
			    (s[i + 2] & 0x3F);
			i += 3;
		} else if ((c & 0xF8) == 0xF0 && i + 3 < s.size()) {
			cp = ((c & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) |
			    ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
			i += 4;
		} else { cp = c; i += 1; }
		if (cp >= 0x10000) {
			cp -= 0x10000;
			out.push_back((uint16_t)(0xD800 + (cp >> 10)));
			out.push_back((uint16_t)(0xDC00 + (cp & 0x3FF)));
		} else {
			out.push_back((uint16_t)cp);
		}
	}
	return out;
}

/* ---- jsprim.pluck ----------------------------------------------------------- */

bool is_index(const std::string &k)
{
	if (k.empty() || (k.size() > 1 && k[0] == '0') || k.size() > 9)
		return false;
	for (char c : k)
		if (c < '0' || c > '9')
			return false;
	return true;
}

bool own(const JP &o, const std::string &k, JP &out)
{
	if (o->t == JV::OBJ) {
		for (auto &kv : o->obj)
			if (kv.first == k) {
				out = kv.second;
				return true;
			}
		return false;
	}
	if (o->t == JV::ARR) {
		if (k == "length") {
			out = mknum((double)o->arr.size());
			return true;
		}
		if (is_index(k)) {
			size_t i = (size_t)atol(k.c_str());
			if (i < o->arr.size()) {
				out = o->arr[i];
				return true;
			}
		}
	}
	return false;
}

JP pluck(const JP &o, const std::string &key)
{
	if (o->t != JV::OBJ && o->t != JV::ARR)
		return UNDEFINED;
	JP v;
	if (own(o, key, v))
		return v;
	size_t d = key.find('.');
	if (d == std::string::npos)
		return UNDEFINED;
	if (!own(o, key.substr(0, d), v))
		return UNDEFINED;
	return pluck(v, key.substr(d + 1));
}

/* a record: parsed fields + the synthetic own-properties assigned onto it */
struct Point {
	JP fields;
	std::vector<std::pair<std::string, JP>> synth;
	uint64_t value = 1;
};

JP pluck_point(const Point &pt, const std::string &key)
{
	if (pt.synth.empty() ||
	    (pt.fields->t != JV::OBJ && pt.fields->t != JV::ARR))
		return pluck(pt.fields, key);
	for (auto it = pt.synth.rbegin(); it != pt.synth.rend(); ++it)
		if (it->first == key)
			return it->second;
	JP v;
	if (own(pt.fields, key, v))
		return v;
	size_t d = key.find('.');
	if (d == std::string::npos)
		return UNDEFINED;
	std::string k1 = key.substr(0, d);
	for (auto it = pt.synth.rbegin(); it != pt.synth.rend(); ++it)
		if (it->first == k1)
			return pluck(it->second, key.substr(d + 1));
	if (!own(pt.fields, k1, v))
		return UNDEFINED;
	return pluck(v, key.substr(d + 1));
}

/* ---- krill ------------------------------------------------------------------ */

/* 1 true, 0 false, -1 threw */
int krill_eval(const JP &pred, const Point &pt)
{
	if (pred->obj.empty())
		return 1;
	const std::string &op = pred->obj[0].first;
	const JP &args = pred->obj[0].second;
	if (op == "and") {
		for (auto &sub : args->arr) {
			int r = krill_eval(sub, pt);
			if (r <= 0)
				return r;
		}
		return 1;
	}
	if (op == "or") {
		for (auto &sub : args->arr) {
			int r = krill_eval(sub, pt);
			if (r != 0)
				return r;
		}
		return 0;
	}
	JP val = pluck_point(pt, args->arr[0]->str);
	if (val->t == JV::UNDEF)
		return -1;
	JP c = args->arr[1];
	if (val->t == JV::OBJ || val->t == JV::ARR) {
		JP s = mk(JV::STR);
		s->str = to_string(val);
		val = s;
	}
	if (op == "eq" || op == "ne") {
		bool eq;
		if (val->t == JV::NUL)
			eq = false;
		else if (val->t == JV::STR && c->t == JV::STR)
			eq = val->str == c->str;
		else
			eq = to_number(val) == to_number(c);
		return (op == "eq") == eq;
	}
	if (val->t == JV::STR && c->t == JV::STR) {
		auto a = utf16(val->str), b = utf16(c->str);
		if (op == "lt") return a < b;
		if (op == "le") return a <= b;
		if (op == "gt") return a > b;
		return a >= b;
	}
	double x = to_number(val), y = to_number(c);
	if (op == "lt") return x < y;
	if (op == "le") return x <= y;
	if (op == "gt") return x > y;
	return x >= y;
}

/* ---- Date.parse (ES5 15.9.1.15 format; no offset = UTC) --------------------- */

int64_t days_from_civil(int64_t y, int m, int d)
{
	y -= m <= 2;
	int64_t era = (y >= 0 ? y : y - 399) / 400;
	int64_t yoe = y - era * 400;
	int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
	int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
	return era * 146097 + doe - 719468;
}

/* does a non-ISO string look like a form V8's legacy Date.parse exists for?
 * (dragnet_b200/csrc/jsdate.cuh dng_date_maybe_legacy) */
bool date_maybe_legacy(const std::string &s)
{
	int seps = 0;
	for (size_t i = 0; i + 2 < s.size(); i++) {
		bool d0 = isdigit((unsigned char)s[i]);
		bool d2 = isdigit((unsigned char)s[i + 2]);
		if (d0 && d2 && (s[i + 1] == '-' || s[i + 1] == '/'))
			seps++;
		if (d0 && d2 && s[i + 1] == ':')
			return true;
	}
	if (seps >= 2)
		return true;
	static const char *months[] = { "jan", "feb", "mar", "apr", "may", "jun",
	    "jul", "aug", "sep", "oct", "nov", "dec" };
	for (size_t i = 0; i + 2 < s.size(); i++) {
		if (i > 0 && isalpha((unsigned char)s[i - 1]))
			continue;
		if (!isalpha((unsigned char)s[i]) ||
		    !isalpha((unsigned char)s[i + 1]) ||
		    !isalpha((unsigned char)s[i + 2]))
			continue;
		char w[4] = { (char)tolower(s[i]), (char)tolower(s[i + 1]),
		    (char)tolower(s[i + 2]), 0 };
		for (const char *m : months)
			if (!strcmp(w, m))
				return true;
	}
	return false;
}

bool date_parse(const std::string &s, int64_t &ms)
{
	size_t i = 0, n = s.size();
	auto digits = [&](size_t cnt, int64_t &out) {
		if (i + cnt > n)
			return false;
		out = 0;
		for (size_t k = 0; k < cnt; k++) {
			if (!isdigit((unsigned char)s[i + k]))
				return false;
			out = out * 10 + (s[i + k] - '0');
		}
		i += cnt;
		return true;
	};
	int64_t y, mo = 1, d = 1, hh = 0, mi = 0, ss = 0, msec = 0, off = 0;
	if (n && (s[0] == '+' || s[0] == '-')) {
		i = 1;
		if (!digits(6, y))
			return false;
		if (s[0] == '-') {
			if (y == 0)
				return false;
			y = -y;
		}
	} else if (!digits(4, y)) {
		return false;
	}
	if (i < n && s[i] == '-') {
		i++;
		if (!digits(2, mo))
			return false;
		if (i < n && s[i] == '-') {
			i++;
			if (!digits(2, d))
				return false;
		}
	}
	if (i < n && s[i] == 'T') {
		i++;
		if (!digits(2, hh) || i >= n || s[i] != ':')
			return false;
		i++;
		if (!digits(2, mi))
			return false;
		if (i < n && s[i] == ':') {
			i++;
			if (!digits(2, ss))
				return false;
			if (i < n && s[i] == '.') {
				i++;
				size_t st = i;
				while (i < n && isdigit((unsigned char)s[i]))
					i++;
				if (i == st)
					return false;
				std::string f = s.substr(st, i - st) + "00";
				msec = atoi(f.substr(0, 3).c_str());
			}
		}
		if (i < n && s[i] == 'Z') {
			i++;
		} else if (i < n && (s[i] == '+' || s[i] == '-')) {
			bool neg = s[i] == '-';
			i++;
			int64_t oh, om;
			if (!digits(2, oh) || i >= n || s[i] != ':')
				return false;
			i++;
			if (!digits(2, om) || oh > 23 || om > 59)
				return false;
			off = (oh * 60 + om) * 60000;
			if (neg)
				off = -off;
		}
	}
	if (i != n)
		return false;
	/* (V8 takes any day 1..31 and lets MakeDay carry it over) */
	if (mo < 1 || mo > 12 || d < 1 || d > 31)
		return false;
	if (hh > 24 || mi > 59 || ss > 59 || (hh == 24 && (mi || ss || msec)))
		return false;
	int64_t t = days_from_civil(y, (int)mo, (int)d) * 86400000 +
	    ((hh * 60 + mi) * 60 + ss) * 1000 + msec - off;
	if (t > 8640000000000000ll || t < -8640000000000000ll)
		return false;
	ms = t;
	return true;
}

/* ---- the plan ---------------------------------------------------------------- */

struct Breakdown { std::string name; int kind; double step; };
struct Synth { std::string name, field; };

struct Plan {
	bool skinner = false;
	JP ds_filter, filter;
	std::vector<Synth> synth;
	bool has_bounds = false;
	std::string bounds_field;
	double ge = 0, lt = 0;
	std::vector<Breakdown> bds;
};

const JP *jget(const JP &o, const char *k)
{
	const JP *r = nullptr;
	for (auto &kv : o->obj)
		if (kv.first == k)
			r = &kv.second;
	return r;
}

bool load_plan(const std::string &text, Plan &pl)
{
	JP root = json_parse((const unsigned char *)text.data(), text.size());
	if (!root || root->t != JV::OBJ)
		return false;
	if (auto f = jget(root, "format"))
		pl.skinner = (*f)->t == JV::STR && (*f)->str == "json-skinner";
	if (auto f = jget(root, "ds_filter"))
		if ((*f)->t == JV::OBJ && !(*f)->obj.empty())
			pl.ds_filter = *f;
	if (auto f = jget(root, "filter"))
		if ((*f)->t == JV::OBJ && !(*f)->obj.empty())
			pl.filter = *f;
	if (auto f = jget(root, "synthetic"))
		if ((*f)->t == JV::ARR)
			for (auto &s : (*f)->arr)
				pl.synth.push_back(Synth{(*jget(s, "name"))->str,
				    (*jget(s, "field"))->str});
	if (auto f = jget(root, "time_bounds")) {
		if ((*f)->t == JV::OBJ) {
			pl.has_bounds = true;
			pl.bounds_field = (*jget(*f, "field"))->str;
			pl.ge = (*jget(*f, "ge"))->num;
			pl.lt = (*jget(*f, "lt"))->num;
		}
	}
	auto b = jget(root, "breakdowns");
	if (!b || (*b)->t != JV::ARR)
		return false;
	for (auto &x : (*b)->arr) {
		Breakdown bd;
		bd.name = (*jget(x, "name"))->str;
		bd.kind = 0;
		bd.step = 0;
		if (auto a = jget(x, "aggr")) {
			if ((*a)->str == "quantize") {
				bd.kind = 1;
			} else {
				bd.kind = 2;
				bd.step = (*jget(x, "step"))->num;
			}
		}
		pl.bds.push_back(bd);
	}
	return true;
}

/* ---- aggregation ------------------------------------------------------------- */

struct KeyPart {
	bool isnum;
	std::string s;
	uint64_t bits;		/* ordinal as binary64, canonical NaN, +0 */
	bool operator<(const KeyPart &o) const {
		if (isnum != o.isnum)
			return isnum < o.isnum;
		return isnum ? bits < o.bits : s < o.s;
	}
};
typedef std::vector<KeyPart> Key;

struct Counters {
	uint64_t lines = 0, invalid_json = 0, invalid_point = 0;
	uint64_t ds_filtered = 0, ds_failedeval = 0, user_filtered = 0,
	    user_failedeval = 0, synth_undef = 0, synth_baddate = 0,
	    time_filtered = 0, time_failedeval = 0, aggr = 0;
	void add(const Counters &o) {
		lines += o.lines; invalid_json += o.invalid_json;
		invalid_point += o.invalid_point;
		ds_filtered += o.ds_filtered; ds_failedeval += o.ds_failedeval;
		user_filtered += o.user_filtered;
		user_failedeval += o.user_failedeval;
		synth_undef += o.synth_undef; synth_baddate += o.synth_baddate;
		time_filtered += o.time_filtered;
		time_failedeval += o.time_failedeval; aggr += o.aggr;
	}
};

struct Tally {
	std::map<Key, uint64_t> table;
	uint64_t total = 0;
	Counters c;
};

uint64_t dbits(double d)
{
	if (d != d)
		return 0x7ff8000000000000ull;
	d += 0.0;
	uint64_t b;
	memcpy(&b, &d, 8);
	return b;
}

void scan_line(const Plan &pl, const unsigned char *s, size_t n, Tally &T)
{
	arena::begin();		/* (everything below dies with the line) */
	T.c.lines++;
	JP obj = json_parse(s, n);
	if (!obj) {
		T.c.invalid_json++;
		return;
	}
	Point pt;
	if (!pl.skinner) {
		pt.fields = obj;
	} else {
		JP f, w;
		bool okf = obj->t == JV::OBJ && own(obj, "fields", f) &&
		    f->t == JV::OBJ;
		bool okw = obj->t == JV::OBJ && own(obj, "value", w) &&
		    w->t == JV::NUM && w->num >= 0 &&
		    w->num <= 9007199254740992.0 && w->num == floor(w->num);
		if (!okf || !okw) {
			T.c.invalid_point++;
			return;
		}
		pt.fields = f;
		pt.value = (uint64_t)w->num;
	}
	if (pl.ds_filter) {
		int r = krill_eval(pl.ds_filter, pt);
		if (r < 0) { T.c.ds_failedeval++; return; }
		if (!r) { T.c.ds_filtered++; return; }
	}
	if (pl.filter) {
		int r = krill_eval(pl.filter, pt);
		if (r < 0) { T.c.user_failedeval++; return; }
		if (!r) { T.c.user_filtered++; return; }
	}
	if (!pl.synth.empty()) {
		int nerr = 0;
		bool container = pt.fields->t == JV::OBJ ||
		    pt.fields->t == JV::ARR;
		for (auto &sc : pl.synth) {
			JP v = pluck_point(pt, sc.field);
			if (v->t == JV::UNDEF) {
				if (!nerr)
					T.c.synth_undef++;
				nerr++;
				continue;
			}
			JP out;
			if (v->t == JV::NUM) {
				out = v;
			} else {
				int64_t ms;
				const std::string text = to_string(v);
				if (!date_parse(text, ms) && v->t == JV::STR &&
				    date_maybe_legacy(text)) {
					/* V8's legacy parser may know it: not
					 * restated, not ours to call NaN */
					fprintf(stderr, "dn_oracle: unsupported: "
					    "Date.parse(\"%s\") outside the ISO "
					    "format\n", text.c_str());
					exit(3);
				}
				if (!date_parse(text, ms)) {
					if (!nerr)
						T.c.synth_baddate++;
					nerr++;
					continue;
				}
				out = mknum(std::floor((double)ms / 1000.0));
			}
			if (container)
				pt.synth.emplace_back(sc.name, out);
		}
		if (nerr)
			return;
	}
	if (pl.has_bounds) {
		JP v = pluck_point(pt, pl.bounds_field);
		if (v->t == JV::UNDEF) { T.c.time_failedeval++; return; }
		double x = to_number(v);
		if (!(x >= pl.ge) || !(x < pl.lt)) { T.c.time_filtered++; return; }
	}
	T.c.aggr++;
	Key key;
	for (auto &b : pl.bds) {
		JP v = pluck_point(pt, b.name);
		KeyPart kp;
		if (b.kind == 0) {
			kp.isnum = false;
			kp.s = to_string(v);
			kp.bits = 0;
		} else {
			double x = to_number(v), o;
			if (b.kind == 1) {
				if (x != x) o = x;
				else if (x < 1) o = 0;
				else if (std::isinf(x)) o = x;
				else { int e; std::frexp(x, &e); o = e; }
			} else {
				double q = x / b.step;
				o = (q != q || std::isinf(q)) ? q : std::floor(q);
			}
			kp.isnum = true;
			kp.bits = dbits(o);
		}
		key.push_back(std::move(kp));
	}
	arena::off();		/* (a new table node outlives the line) */
	T.table[key] += pt.value;
	T.total += pt.value;
}

void scan_range(const Plan &pl, const unsigned char *d, size_t a, size_t b,
    size_t total, Tally &T)
{
	/* lines whose FIRST byte lies in [a, b); lstream emits every line
	 * (including empty ones) and a final unterminated non-empty tail */
	size_t pos = a;
	if (a > 0 && d[a - 1] != '\n') {
		const void *nl = memchr(d + a, '\n', total - a);
		if (!nl)
			return;
		pos = (const unsigned char *)nl - d + 1;
	}
	while (pos < b && pos < total) {
		const void *nl = memchr(d + pos, '\n', total - pos);
		size_t end = nl ? (size_t)((const unsigned char *)nl - d) : total;
		scan_line(pl, d + pos, end - pos, T);
		pos = end + 1;
	}
}

double bucket_min(const Breakdown &b, double o)
{
	if (b.kind == 1) {
		if (o != o || o == 0 || std::isinf(o))
			return o == 0 ? 0.0 : o;
		return std::ldexp(1.0, (int)o - 1);
	}
	return o * b.step;
}

} /* namespace */

int main(int argc, char **argv)
{
	if (argc < 2) {
		fprintf(stderr, "usage: dn_oracle PLAN.json [--threads N] "
		    "[--repeat R] [--min-seconds S] FILE...\n");
		return 2;
	}
	std::ifstream pf(argv[1], std::ios::binary);
	std::stringstream pss;
	pss << pf.rdbuf();
	Plan pl;
	if (!load_plan(pss.str(), pl)) {
		fprintf(stderr, "dn_oracle: bad plan\n");
		return 1;
	}
	int threads = 1, repeat = 1;
	double min_seconds = 0;
	std::string data;
	for (int i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--threads") && i + 1 < argc) {
			threads = atoi(argv[++i]);
		} else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) {
			repeat = atoi(argv[++i]);
		} else if (!strcmp(argv[i], "--min-seconds") && i + 1 < argc) {
			/* keep scanning until this much time has been measured
			 * (bench.py: a step long enough to be stable) */
			min_seconds = atof(argv[++i]);
		} else {
			std::ifstream f(argv[i], std::ios::binary);
			std::stringstream ss;
			ss << f.rdbuf();
			data += ss.str();
		}
	}
	if (threads < 1)
		threads = 1;
	const unsigned char *d = (const unsigned char *)data.data();
	size_t total = data.size();
	Tally result;
	double best = 1e300, spent = 0;
	int reps = 0;
	for (int rep = 0; rep < repeat || spent < min_seconds; rep++) {
		auto t0 = std::chrono::steady_clock::now();
		std::vector<Tally> parts(threads);
		std::vector<std::thread> th;
		for (int t = 0; t < threads; t++) {
			size_t a = total * t / threads, b = total * (t + 1) / threads;
			th.emplace_back([&, a, b, t]() {
				scan_range(pl, d, a, b, total, parts[t]);
			});
		}
		for (auto &x : th)
			x.join();
		Tally merged;
		for (auto &p : parts) {
			for (auto &kv : p.table)
				merged.table[kv.first] += kv.second;
			merged.total += p.total;
			merged.c.add(p.c);
		}
		auto t1 = std::chrono::steady_clock::now();
		const double dt = std::chrono::duration<double>(t1 - t0).count();
		best = std::min(best, dt);
		spent += dt;
		reps++;
		result = std::move(merged);
	}
	if (pl.bds.empty()) {
		result.table.clear();
		result.table[Key()] = result.total;
	}
	printf("{\"points\":[");
	bool first = true;
	for (auto &kv : result.table) {
		printf("%s{\"cols\":[", first ? "" : ",");
		first = false;
		for (size_t j = 0; j < kv.first.size(); j++) {
			const KeyPart &kp = kv.first[j];
			if (j)
				printf(",");
			if (kp.isnum) {
				double o;
				memcpy(&o, &kp.bits, 8);
				double m = bucket_min(pl.bds[j], o);
				uint64_t b;
				memcpy(&b, &m, 8);
				printf("{\"n\":\"%016llx\"}", (unsigned long long)b);
			} else {
				printf("{\"s\":\"");
				for (unsigned char ch : kp.s)
					printf("%02x", ch);
				printf("\"}");
			}
		}
		printf("],\"value\":%llu}", (unsigned long long)kv.second);
	}
	const Counters &c = result.c;
	printf("],\"counters\":{\"lines\":%llu,\"invalid_json\":%llu,"
	    "\"invalid_point\":%llu,\"ds_filtered\":%llu,\"ds_failedeval\":%llu,"
	    "\"user_filtered\":%llu,\"user_failedeval\":%llu,\"synth_undef\":%llu,"
	    "\"synth_baddate\":%llu,\"time_filtered\":%llu,"
	    "\"time_failedeval\":%llu,\"aggr\":%llu,\"unsupported\":0},"
	    "\"seconds\":%.6f,\"mean_seconds\":%.6f,\"reps\":%d,"
	    "\"threads\":%d,\"bytes\":%llu}\n",
	    (unsigned long long)c.lines, (unsigned long long)c.invalid_json,
	    (unsigned long long)c.invalid_point,
	    (unsigned long long)c.ds_filtered,
	    (unsigned long long)c.ds_failedeval,
	    (unsigned long long)c.user_filtered,
	    (unsigned long long)c.user_failedeval,
	    (unsigned long long)c.synth_undef,
	    (unsigned long long)c.synth_baddate,
	    (unsigned long long)c.time_filtered,
	    (unsigned long long)c.time_failedeval, (unsigned long long)c.aggr,
	    best, spent / (reps ? reps : 1), reps, threads,
	    (unsigned long long)total);
	return 0;
}
