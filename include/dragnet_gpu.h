/*
 * dragnet_gpu.h: C ABI of libdragnet_gpu.so -- the B200-native replacement for
 * dragnet's raw-data scan path.
 *
 * This is the drop-in boundary.  The reference has no
 * FFI of its own (it is pure JavaScript), so the entry points below are what a
 * thin N-API addon binds to implement a `Datasource` whose scan() replaces
 *
 *     DatasourceFile.scan()            lib/datasource-file.js:72-108
 *       -> parserFor(format)           lib/dragnet-impl.js:131-140
 *          JsonLineStream              lib/format-json.js:26-46
 *       -> [Datasource filter]         lib/datasource-file.js:154-163
 *       -> new StreamScan(...)         lib/stream-scan.js:40-94
 *            KrillSkinnerStream        lib/krill-skinner-stream.js:29-52
 *            SyntheticTransformer      lib/stream-synthetic.js:37-85
 *            time-bounds filter        lib/dragnet-impl.js:94-125
 *            skinner aggregator        lib/dragnet-impl.js:48-89
 *
 * i.e. bytes of newline-delimited JSON in, skinner points out.  See
 * INTEGRATION.md for the reference-side binding (datasource-gpu.js + addon).
 *
 * Conventions: every function returns 0 on success or a negative DNG_E* code;
 * functions taking (err, errlen) also write a NUL-terminated message.  Plain
 * pointers and sizes only.  A dng_scan is not thread-safe (serialise calls
 * externally); different scans may be driven from different threads.  There is
 * no CPU fallback: without a CUDA device dng_scan_open fails with DNG_ENODEV.
 */
#ifndef DRAGNET_GPU_H
#define DRAGNET_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNG_OK            0
#define DNG_EINVAL       -1   /* bad argument / malformed plan               */
#define DNG_ENODEV       -2   /* no usable CUDA device                       */
#define DNG_ECUDA        -3   /* CUDA runtime error (see dng_scan_error)     */
#define DNG_ENOMEM       -4
#define DNG_EIO          -5   /* dng_scan_feed_file could not read           */
#define DNG_EUNSUPPORTED -6   /* input needed a path the device code lacks   */
#define DNG_ELIMIT       -7   /* a documented limit was exceeded             */
#define DNG_ENCCL        -8

typedef struct dng_plan dng_plan;
typedef struct dng_scan dng_scan;
typedef struct dng_result dng_result;
typedef struct dng_comm dng_comm;

/*
 * Per-stage counters, mirroring the vstream counters `dn scan --counters`
 * prints (bin/dn:911-916; goldens tests/dn/local/tst.scan_fileset.sh.out:2455-2660).
 * Stage ninputs/noutputs are derivable: e.g. json parser ninputs = lines,
 * noutputs = lines - invalid_json.
 */
typedef struct dng_counters {
	uint64_t lines;            /* json parser ninputs (every line, incl. empty) */
	uint64_t invalid_json;     /* json parser "invalid json"                    */
	uint64_t invalid_point;    /* json-skinner line that is not a point         */
	uint64_t ds_ninputs;       /* Datasource filter ninputs                     */
	uint64_t ds_filtered;      /* Datasource filter nfilteredout                */
	uint64_t ds_failedeval;    /* Datasource filter nfailedeval                 */
	uint64_t user_ninputs;     /* User filter ...                               */
	uint64_t user_filtered;
	uint64_t user_failedeval;
	uint64_t synth_ninputs;    /* Datetime parser ninputs                       */
	uint64_t synth_undef;      /* Datetime parser undef                         */
	uint64_t synth_baddate;    /* Datetime parser baddate                       */
	uint64_t time_ninputs;     /* Time filter ...                               */
	uint64_t time_filtered;
	uint64_t time_failedeval;
	uint64_t aggr_ninputs;     /* Aggregator ninputs (records aggregated)       */
	uint64_t slowpath_records; /* records that took a device slow path          */
	uint64_t long_records;     /* records longer than the in-tile window        */
	uint64_t unsupported;      /* records the device code could not decide      */
	uint64_t bytes;            /* input bytes consumed                          */
} dng_counters;

/* ---- plan ------------------------------------------------------------- */

/*
 * plan_json: the serialised QueryConfig + datasource properties
 * (lib/dragnet.js:28-77, lib/datasource-file.js:47-56):
 *   { "format": "json" | "json-skinner",
 *     "ds_filter": <krill predicate> | null,
 *     "filter":    <krill predicate> | null,
 *     "synthetic": [ { "name": s, "field": s }, ... ],   // qc_synthetic (+dn_ts)
 *     "time_bounds": { "field": "dn_ts", "ge": sec, "lt": sec } | null,
 *     "breakdowns": [ { "name": s, "field": s, ["date": true,]
 *                       ["aggr": "quantize" | "lquantize", "step": n] }, ... ] }
 * or, to scan several metrics in one pass over the data (dn build):
 *   { "format": ..., "ds_filter": ...,
 *     "metrics": [ { "filter", "synthetic", "time_bounds", "breakdowns" }, ... ] }
 */
int dng_plan_create(const char *plan_json, dng_plan **out,
    char *err, size_t errlen);
void dng_plan_destroy(dng_plan *plan);

/* ---- scan ------------------------------------------------------------- */

int dng_device_count(void);

int dng_scan_open(const dng_plan *plan, int device, dng_scan **out,
    char *err, size_t errlen);

/* Run this scan's kernels on a caller-owned CUDA stream (a cudaStream_t) instead
 * of the scan's private one, e.g. to bracket them with the caller's events.
 * Call before the first feed. */
int dng_scan_set_stream(dng_scan *scan, void *cuda_stream);

/*
 * Feed arbitrary byte chunks (lstream semantics: lines split on '\n', partial
 * trailing line carried to the next feed; lib/format-json.js:32-33).
 * dng_scan_feed: any host memory; the bytes are consumed (copied to a pinned
 *   staging ring) before it returns.
 * dng_scan_feed_pinned: page-locked host memory (dng_pinned_alloc or
 *   cudaHostRegister'd); DMA'd directly; the buffer must stay valid and
 *   unmodified until the next dng_scan_sync()/dng_scan_finish().
 * dng_scan_feed_device: bytes already resident in this device's HBM
 *   (16-byte aligned); scanned in place, no copy.  The kernels that read the
 *   buffer are only ENQUEUED when this returns (on the scan's stream, or the
 *   caller's after dng_scan_set_stream): like a pinned buffer, it must stay
 *   valid and unmodified until the next dng_scan_sync()/dng_scan_finish().
 * dng_scan_feed_file: read(2) the file into the pinned ring and feed it.
 */
int dng_scan_feed(dng_scan *scan, const void *buf, size_t len);
int dng_scan_feed_pinned(dng_scan *scan, const void *buf, size_t len);
int dng_scan_feed_device(dng_scan *scan, const void *devbuf, size_t len);
int dng_scan_feed_file(dng_scan *scan, const char *path);
int dng_scan_sync(dng_scan *scan);

/* End of input: flushes the final unterminated line, then emits the points. */
int dng_scan_finish(dng_scan *scan, dng_result **out);
int dng_scan_counters(dng_scan *scan, dng_counters *out);
/* counters of metric m's StreamScan (m = 0 is what dng_scan_counters gives) */
int dng_scan_counters_metric(dng_scan *scan, int metric, dng_counters *out);
const char *dng_scan_error(const dng_scan *scan);
void dng_scan_destroy(dng_scan *scan);
/* Destroyed scans keep their device and pinned buffers for the next scan of
 * this process (bounded by DNG_CACHE_BYTES per kind, default 2 GiB); this
 * returns them to the driver. */
void dng_release_cached(void);

/* Device time (ms, CUDA events on the scan stream) spent in scan kernels and
 * the number of kernel launches since open; for bench.py's roofline. */
int dng_scan_kernel_stats(dng_scan *scan, double *kernel_ms,
    uint64_t *launches, uint64_t *kernel_bytes);

/*
 * Record templates: the scan learns the few shapes (key order + punctuation)
 * the input's records repeat from the head of the first data it is fed, and
 * matches records against them before falling back to its byte automaton;
 * results never depend on it.  On by default (environment DNG_TEMPLATES=0
 * turns it off); set_templates must precede the first feed.  The stats give
 * the number of templates in use and how many records they accepted.
 */
int dng_scan_set_templates(dng_scan *scan, int enable);
int dng_scan_template_stats(dng_scan *scan, uint64_t *templates,
    uint64_t *templated_records);

/*
 * Which kernel the scan runs (chosen from the plan and from a sample of the
 * first data fed; environment DNG_KERNEL=tile|warp|fast forces one): 0 = CTA-wide
 * tiles (any line length), 1 = per-warp chunks (short lines), 2 = the F path
 * (templated input, plans it models: scan_kernel_f + the miss kernel).
 */
int dng_scan_kernel_kind(const dng_scan *scan);
/*
 * The F path's run-time compiled matcher (NVRTC + nvJitLink, cached per
 * process; environment DNG_JIT=0|async|sync, default async: scans use the
 * interpreted matcher until the compiled one is there).  *state: 0 none / being
 * built, 1 in use, 2 failed (message in err); *launches = scan launches that
 * ran the compiled kernel; compile and link times in milliseconds.
 */
int dng_scan_jit_stats(dng_scan *scan, int *state, uint64_t *launches,
    double *compile_ms, double *link_ms, char *err, size_t errlen);
/* Every kernel this scan has launched so far: scan kernels plus the small
 * ones around them (template resolution, newline search of device feeds,
 * result compaction). */
uint64_t dng_scan_launch_count(const dng_scan *scan);

void *dng_pinned_alloc(size_t len);
void dng_pinned_free(void *p);

/* ---- results: skinner points ---------------------------------------------
 * One point per observed tuple; with zero breakdowns exactly one point whose
 * value is the total (possibly 0); nothing for >=1 breakdown and no input
 * (tests/dn/local/tst.empty.sh.out).  Column j of point i is either a string
 * (discrete breakdown: the JS String(value), UTF-8, not NUL-terminated) or a
 * number (quantized breakdown: bucketMin(ordinal), may be NaN/Inf).
 * Points are sorted by encoded key bytes (deterministic; the reference's
 * emission order is unspecified).
 */
size_t dng_result_count(const dng_result *r);
size_t dng_result_ncols(const dng_result *r);
/* Fan-out plans ("metrics": [...], the reference's dn build / index-scan,
 * lib/datasource-file.js:386-432): every point belongs to one metric (the
 * reference tags it fields.__dn_metric) and has that metric's columns. */
size_t dng_result_nmetrics(const dng_result *r);
size_t dng_result_ncols_metric(const dng_result *r, int metric);
int dng_result_metric(const dng_result *r, size_t i);
int dng_result_get(const dng_result *r, size_t i, const char **strs,
    size_t *strlens, uint8_t *is_number, double *numvals, uint64_t *value);
void dng_result_destroy(dng_result *r);

/* ---- shard merge (the reference's Manta reduce phase,
 * lib/datasource-manta.js:202-219): sum values over identical tuples ------ */

/* Build a result by re-aggregating skinner points, as the reference's reduce
 * phase does (`dn scan --points` over json-skinner input,
 * lib/datasource-manta.js:212-219): string columns are taken as is, numeric
 * columns are bucket minima and are re-bucketized (idempotent).  Column j of
 * point i is strs[i*ncols+j] (strlens bytes) or numvals[i*ncols+j]. */
int dng_result_from_points(const dng_plan *plan, size_t npoints,
    const char *const *strs, const size_t *strlens, const double *numvals,
    const uint64_t *values, dng_result **out);

/* Serialised tuple dictionary of a result (keys only), for exchange. */
int dng_result_dict(const dng_result *r, const void **buf, size_t *len);
/* Union of n serialised dictionaries -> sorted global dictionary. */
int dng_dict_union(const void *const *bufs, const size_t *lens, size_t n,
    void **out, size_t *outlen);
void dng_buf_free(void *p);
size_t dng_dict_count(const void *dict, size_t len);
/* Scatter r's values into a dense vector indexed by the global dictionary. */
int dng_result_dense(const dng_result *r, const void *dict, size_t dictlen,
    uint64_t *vec, size_t n);
/* Rebuild a result from the global dictionary + summed dense vector. */
int dng_result_from_dense(const dng_result *like, const void *dict,
    size_t dictlen, const uint64_t *vec, size_t n, dng_result **out);

/* NCCL transport for the above: ncclAllGather of the dictionaries and ONE
 * ncclReduce(sum, uint64) of the dense tallies (+ counters) to `root`. */
int dng_comm_unique_id(void *id128);            /* rank 0; 128 bytes out */
int dng_comm_init(dng_comm **out, int nranks, int rank, const void *id128,
    int device, char *err, size_t errlen);
int dng_merge_nccl(dng_scan *scan, dng_comm *comm, int root,
    dng_result **out, dng_counters *counters);
void dng_comm_destroy(dng_comm *comm);

/* ---- synthetic input (tools/mktestdata shape, deterministic) ------------
 * Writes records [first, first+count) of stream `seed` as NDJSON.
 * dng_gen_host fills host memory; dng_gen_device runs the generator kernel
 * on `device`; both produce byte-identical output. Returns bytes written via
 * *len (or DNG_ELIMIT when cap is too small).
 */
typedef struct dng_gen_params {
	uint64_t seed;
	uint64_t total_records;   /* n in `time = round(j/n*(max-min)+min)`   */
	int64_t  time_min_ms;     /* default 2014-05-31T21:00:00Z             */
	int64_t  time_max_ms;     /* default 2014-05-31T23:59:59Z             */
	int      string_latency;  /* emit latency as a string (mktestdata)    */
} dng_gen_params;
void dng_gen_defaults(dng_gen_params *p);
int dng_gen_host(const dng_gen_params *p, uint64_t first, uint64_t count,
    void *buf, size_t cap, size_t *len);
int dng_gen_device(const dng_gen_params *p, int device, uint64_t first,
    uint64_t count, void *devbuf, size_t cap, size_t *len);

const char *dng_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DRAGNET_GPU_H */
