/*
 * gen_ndjson: the synthetic workload of bench.py / the parity tests as a
 * stand-alone program -- TEST / BENCH INFRASTRUCTURE, like the rest of oracle/.
 * The CPU reference arm writes its sample with this, so that it never loads
 * the product library.  The record builder is dragnet_b200/csrc/gen.cuh (the
 * reference's tools/mktestdata:15-99, 138-190 made deterministic), compiled
 * here for the host; tests check that the output is byte-identical to
 * dng_gen_host / dng_gen_device.
 *
 *   gen_ndjson OUT SEED TOTAL_RECORDS FIRST COUNT [THREADS] [STRING_LATENCY]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "../dragnet_b200/csrc/gen.cuh"

int main(int argc, char **argv)
{
	if (argc < 6) {
		fprintf(stderr, "usage: gen_ndjson OUT SEED TOTAL FIRST COUNT "
		    "[THREADS] [STRING_LATENCY]\n");
		return 2;
	}
	dng_gen_params p;
	memset(&p, 0, sizeof (p));
	p.seed = strtoull(argv[2], nullptr, 0);
	p.total_records = strtoull(argv[3], nullptr, 0);
	/* mktestdata's default window (tools/mktestdata:15-16) */
	p.time_min_ms = 1401570000000ll;
	p.time_max_ms = 1401580799000ll;
	p.string_latency = argc > 7 ? atoi(argv[7]) : 0;
	const uint64_t first = strtoull(argv[4], nullptr, 0);
	const uint64_t count = strtoull(argv[5], nullptr, 0);
	int nth = argc > 6 ? atoi(argv[6]) : 1;
	if (nth < 1)
		nth = 1;
	std::vector<std::string> parts(nth);
	std::vector<std::thread> th;
	for (int t = 0; t < nth; t++) {
		th.emplace_back([&, t]() {
			const uint64_t a = first + count * t / nth;
			const uint64_t b = first + count * (t + 1) / nth;
			std::string &o = parts[t];
			o.reserve((size_t)(b - a) * 232);
			char tmp[dng::GEN_MAXREC];
			for (uint64_t j = a; j < b; j++)
				o.append(tmp, (size_t)dng::gen_record(p, j, tmp));
		});
	}
	for (auto &x : th)
		x.join();
	FILE *f = fopen(argv[1], "wb");
	if (!f) {
		perror(argv[1]);
		return 1;
	}
	for (auto &o : parts)
		if (fwrite(o.data(), 1, o.size(), f) != o.size()) {
			perror("write");
			return 1;
		}
	fclose(f);
	return 0;
}
