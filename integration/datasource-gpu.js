/*
 * lib/datasource-gpu.js: Datasource backend that runs `scan` on a B200 through
 * libdragnet_gpu.so (via the dragnet_gpu N-API addon, integration/addon.cc).
 *
 * NOT EXECUTED IN THIS REPOSITORY: node is not installed in the build image.
 * It is the reference-side half of the drop-in boundary and mirrors
 * lib/datasource-file.js:31-108 of the reference; everything except scan()
 * delegates to the file backend.  Add to lib/dragnet.js:datasourceForConfig:
 *
 *     if (bename == 'gpu')
 *             return (mod_datasource_gpu.createDatasource(args));
 */

var mod_assertplus = require('assert-plus');
var mod_stream = require('stream');
var mod_vstream = require('vstream');
var VError = require('verror');

var mod_datasource_file = require('./datasource-file');
var mod_dragnet_impl = require('./dragnet-impl');
var dragnet_gpu = require('dragnet_gpu');	/* the N-API addon */

exports.createDatasource = function createDatasource(args)
{
	var filecfg, fileds;

	mod_assertplus.object(args.dsconfig);
	if (typeof (args.dsconfig.ds_backend_config.path) != 'string')
		return (new VError('expected datasource "path" to be a string'));

	/* build/query/index* stay on the file backend */
	filecfg = Object.create(args.dsconfig);
	filecfg.ds_backend = 'file';
	fileds = mod_datasource_file.createDatasource(
	    { 'dsconfig': filecfg, 'log': args.log });
	if (fileds instanceof Error)
		return (fileds);
	return (new DatasourceGpu(args, fileds));
};

function DatasourceGpu(args, fileds)
{
	this.ds_file = fileds;
	this.ds_format = args.dsconfig.ds_format;
	this.ds_timefield = args.dsconfig.ds_backend_config.timeField || null;
	this.ds_filter = args.dsconfig.ds_filter || null;
	this.ds_device = args.dsconfig.ds_backend_config.device || 0;
	this.ds_log = args.log;
}

[ 'close', 'build', 'query', 'indexScan', 'indexRead' ].forEach(function (m) {
	DatasourceGpu.prototype[m] = function () {
		return (this.ds_file[m].apply(this.ds_file, arguments));
	};
});

/*
 * scan(): same contract as DatasourceFile.scan (lib/datasource-file.js:72-108):
 * returns an object-mode Readable of skinner points, then 'end'.
 */
DatasourceGpu.prototype.scan = function (args)
{
	var self = this;
	var query = args.query;
	var scanctx, plan, scan, out, synthetic, bounds, tf;

	/* file enumeration, --dry-run and the timeField check are unchanged */
	scanctx = this.ds_file.scanInit({
	    'filter': null, 'dryRun': args.dryRun,
	    'timeBefore': query.qc_before, 'timeAfter': query.qc_after
	});
	if (scanctx instanceof Error)
		return (mod_dragnet_impl.asyncError(scanctx));
	if (args.dryRun)
		return (scanctx.outstream);

	/* the plan JSON is just the QueryConfig + datasource properties */
	synthetic = query.qc_synthetic.map(function (s) {
		return ({ 'name': s.name, 'field': s.field });
	});
	bounds = null;
	/* (as lib/stream-scan.js:62: either bound asks for the time filter) */
	if (query.qc_before !== null || query.qc_after !== null) {
		synthetic.push({ 'name': 'dn_ts', 'field': this.ds_timefield });
		tf = mod_dragnet_impl.queryTimeBoundsFilter(query, 'dn_ts');
		bounds = { 'field': 'dn_ts',
		    'ge': tf.and[0].ge[1], 'lt': tf.and[1].lt[1] };
	}
	plan = JSON.stringify({
	    'format': this.ds_format,
	    'ds_filter': this.ds_filter,
	    'filter': query.qc_filter,
	    'synthetic': synthetic,
	    'time_bounds': bounds,
	    'breakdowns': query.qc_breakdowns
	});

	out = mod_vstream.wrapStream(new mod_stream.PassThrough(
	    { 'objectMode': true, 'highWaterMark': 0 }), 'Aggregator');
	/* throws on error; the names label the point fields, in order */
	scan = dragnet_gpu.scanOpen(plan, this.ds_device,
	    query.qc_breakdowns.map(function (b) { return (b.name); }));

	/*
	 * Feed files as they are found.  feedFile() is an N-API async work item:
	 * read(2) into pinned buffers + cudaMemcpyAsync run on a worker thread,
	 * the event loop is never blocked.
	 */
	scanctx.findstream.on('data', function (fileinfo) {
		if (fileinfo.error)
			return;
		scanctx.findstream.pause();
		scan.feedFile(fileinfo.path, function (err) {
			if (err)
				out.emit('error', err);
			else
				scanctx.findstream.resume();
		});
	});
	scanctx.findstream.on('end', function () {
		scan.finish(function (err, points, counters) {
			if (err) {
				out.emit('error', err);
				return;
			}
			/* counters mirror the vstream per-stage counters */
			self.ds_counters = counters;
			points.forEach(function (p) { out.write(p); });
			out.end();
			scan.close();
		});
	});
	return (out);
};
