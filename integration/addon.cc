/*
 * dragnet_gpu N-API addon: the thinnest possible binding of libdragnet_gpu.so
 * (include/dragnet_gpu.h) for lib/datasource-gpu.js.
 *
 * NOT BUILT IN THIS REPOSITORY (no node headers in the image).  Each exported
 * function is a direct call of one C-ABI entry point; feedFile/finish run as
 * napi async work so the event loop never blocks.
 */
#include <node_api.h>
#include <string>
#include <vector>
#include "dragnet_gpu.h"

struct ScanWrap { dng_plan *plan; dng_scan *scan; };

struct FeedWork {
	napi_async_work work; napi_ref cb; ScanWrap *w; std::string path; int rc;
};
static void FeedExec(napi_env, void *d) {
	FeedWork *f = (FeedWork *)d;
	f->rc = dng_scan_feed_file(f->w->scan, f->path.c_str());
}
static void FeedDone(napi_env env, napi_status, void *d) {
	FeedWork *f = (FeedWork *)d;
	napi_value cb, undef, argv[1];
	napi_get_reference_value(env, f->cb, &cb);
	napi_get_undefined(env, &undef);
	if (f->rc != 0) {
		napi_value msg;
		napi_create_string_utf8(env, dng_scan_error(f->w->scan),
		    NAPI_AUTO_LENGTH, &msg);
		napi_create_error(env, nullptr, msg, &argv[0]);
	} else {
		argv[0] = undef;
	}
	napi_call_function(env, undef, cb, 1, argv, nullptr);
	napi_delete_reference(env, f->cb);
	napi_delete_async_work(env, f->work);
	delete f;
}
/* scan.feedFile(path, cb) */
static napi_value FeedFile(napi_env env, napi_callback_info info);
/* scan.finish(cb): dng_scan_finish + dng_scan_counters on a worker, then build
 * [{fields:{name: string|number, ...}, value: N}, ...] from dng_result_get():
 * is_number[j] ? napi_create_double(numvals[j]) : napi_create_string_utf8(strs[j], strlens[j]) */
static napi_value Finish(napi_env env, napi_callback_info info);
/* scanOpen(planJson, device) -> object wrapping {dng_plan*, dng_scan*} with
 * feedFile/feed(Buffer)/finish; throws Error(err) on DNG_E* */
static napi_value ScanOpen(napi_env env, napi_callback_info info);

NAPI_MODULE_INIT() {
	napi_value fn;
	napi_create_function(env, "scanOpen", NAPI_AUTO_LENGTH, ScanOpen, nullptr, &fn);
	napi_set_named_property(env, exports, "scanOpen", fn);
	return exports;
}
