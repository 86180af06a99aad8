/*
 * tmpl.h: record templates ("skeleton speculation") for the scan kernel.
 *
 * Machine-written logs repeat a handful of shapes: the same keys, in the same
 * order, with the same punctuation; only the scalar values change.  A template
 * is one such shape: the record's bytes with every scalar VALUE cut out,
 *
 *     {"time":"  ·  ","host":"  ·  ","req":{"method":"  ·  ", ... ,"latency":  ·  }
 *
 * i.e. a list of literal byte runs separated by wildcards (a string body, or a
 * bare scalar: number / true / false / null).  A record whose literal runs are
 * byte-identical to a template's and whose wildcards are well-formed is valid
 * JSON with exactly the template's structure, so everything the general parser
 * would have decided from structure (which key is a duplicate, which value a
 * dotted path plucks, lib/stream-synthetic.js:47 semantics) is decided once,
 * per template, when the template is built; per record only word-wise literal
 * compares and value scans remain.  Records that match no template are parsed
 * by the byte automaton / general parser (record.cuh) as before, so results
 * never depend on what was learned.
 *
 * The templates of one scan form a trie (shared literal prefixes are compared
 * once); the trie is a flat blob the kernel copies into shared memory:
 *
 *     THdr | TNode[nnodes] | u32 leaf_set_mask[nleaves] | literal pool
 */
#ifndef DNG_TMPL_H
#define DNG_TMPL_H

#include "plan.h"

namespace dng {

enum : u8 { TK_NONE = 0, TK_STR = 1, TK_BARE = 2 };
enum : u16 { TN_LEAF = 0x8000, TN_NOALT = 0xffff, TN_NODISP = 0xffff };

struct alignas(16) TNode {
	u16 lit;	/* literal: pool offset (16-byte aligned) of ceil(len/8)
			 * entries { value0, mask0, value1, mask1 }: eight literal
			 * bytes as two little-endian words, padding masked out */
	u16 len;	/* literal length (may be 0) */
	u16 next;	/* node after the wildcard, or TN_LEAF | leaf */
	u16 alt;	/* sibling to try if the literal differs, or TN_NOALT */
	u8 kind;	/* wildcard following the literal (TK_*) */
	u8 cap;		/* 1 + plan slot receiving the wildcard, or 0 */
	u8 poscap;	/* 1 + plan slot receiving the container that opens at
			 * byte `posoff` of the literal, or 0 */
	u8 pad;
	u16 posoff;
	u16 disp;	/* dispatch table of this node's children (pool offset
			 * / 4), or TN_NODISP: see THdr */
};

/*
 * Siblings (the children of one node, or the root's) are alternatives at the
 * same place in the record.  Where their literals all reach some byte offset k
 * and differ there, a dispatch table in the pool -- { k | n << 16 } then n
 * entries { byte | node << 16 } -- sends a record straight to the first
 * sibling that can match (its `alt` chain then holds only siblings with the
 * same byte at k); a byte without an entry matches none.  Matching starts at
 * node 0: the only top-level shape's first node, or, when there are several,
 * an empty node whose children they are.
 */
struct alignas(16) THdr {
	u32 bytes;	/* whole blob, multiple of 16 */
	u16 nnodes, nleaves;
	u16 leaf_off;	/* byte offset of the leaf table */
	u16 pool_off;	/* byte offset of the literal pool */
	u16 pad[2];
};

enum : u32 {
	TMPL_MAX_NODES = 448, TMPL_MAX_POOL = 12288, TMPL_MAX_LEAVES = 24,
	TMPL_MAX_LINE = 4096,		/* longer lines are never templated */
	TMPL_SAMPLE_BYTES = 256 * 1024
};

} /* namespace dng */

#include <string>
#include <vector>

namespace dng {

struct TSeg {
	std::string lit;	/* literal run */
	u8 kind;		/* wildcard after it */
	u32 woff, wlen;		/* the wildcard in the sample line */
};

/* what the record parser captured for a candidate's sample line */
struct TResolved {
	u32 flags, set_mask;
	u64 slots[MAX_SLOTS];
};

struct TCandidate {
	std::string sample;	/* one line, without its newline */
	std::vector<TSeg> segs;
	u32 count;
};

/* lexical skeleton of one line; false if it cannot be a JSON text */
bool tmpl_skeletonize(const u8 *s, u32 n, std::vector<TSeg> &out);

/* distinct skeletons of the complete lines in data[0, len), most frequent
 * first, at most maxk; *nlines = the lines looked at */
void tmpl_candidates(const u8 *data, size_t len, size_t maxk,
    std::vector<TCandidate> &out, size_t *nlines = nullptr);

/* the trie blob (at most max_bytes) for the candidates that resolved cleanly,
 * most frequent first; empty if none did.  *ntemplates = how many it holds.
 * compact: literals as plain little-endian words, zero padded, instead of
 * { value, mask } pairs (the F path, fast.cuh fmatch).  accepted[i] = 1 for
 * the candidates the blob holds. */
void tmpl_build(const std::vector<TCandidate> &cands,
    const std::vector<TResolved> &res, size_t max_bytes, std::vector<u8> &blob,
    u32 *ntemplates, bool compact = false,
    std::vector<u8> *accepted = nullptr);

} /* namespace dng */

#endif
