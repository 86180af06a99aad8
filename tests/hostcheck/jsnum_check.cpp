/* TEST-ONLY: exposes dragnet_b200/csrc/jsnum.cuh + jsdate.cuh on stdin/stdout. */
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include "../../dragnet_b200/csrc/jsnum.cuh"
#include "../../dragnet_b200/csrc/jsdate.cuh"
using namespace dng;
int main()
{
	std::string line;
	while (std::getline(std::cin, line)) {
		if (line.size() < 2)
			continue;
		char op = line[0];
		std::string a = line.substr(2);
		if (op == 'p') {		/* parse decimal -> bits */
			double d = dng_parse_decimal((const uint8_t *)a.data(), (int)a.size());
			printf("%016llx\n", (unsigned long long)double_to_bits(d));
		} else if (op == 's') {		/* bits -> JS string */
			unsigned long long b = strtoull(a.c_str(), nullptr, 16);
			char out[40];
			int n = dng_number_to_string(bits_to_double(b), out);
			printf("%.*s\n", n, out);
		} else if (op == 'n') {		/* hex bytes -> ToNumber bits */
			std::string raw;
			for (size_t i = 0; i + 1 < a.size(); i += 2)
				raw += (char)strtoul(a.substr(i, 2).c_str(), nullptr, 16);
			double d = dng_string_to_number((const uint8_t *)raw.data(), (int)raw.size());
			printf("%016llx\n", (unsigned long long)double_to_bits(d));
		} else if (op == 'd') {		/* Date.parse */
			int64_t ms;
			if (dng_date_parse((const uint8_t *)a.data(), (int)a.size(), &ms))
				printf("%lld\n", (long long)ms);
			else
				printf("NaN\n");
		}
	}
	return 0;
}
