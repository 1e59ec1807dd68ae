// tests/cpu_probe/expand_probe.cpp -- the host-thread pool of the 2-bit Level-2 transfer (gpview_b200/csrc/gpv_expand.cpp) exercised the
// way gpv_voxelize_host uses it: begin, a few chunk submissions trickling in, end -- many calls in a row, from two client threads
// (the pool serialises them), with every output byte checked.  Built with -fsanitize=thread by tests/test_host_sanitizers.py.
#include "../../gpview_b200/csrc/gpv_internal.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace gpv { int fail(const std::string&) { return 1; } }

static int run_client(unsigned seed, int calls)
{
	std::mt19937 rng(seed);
	gpv::ExpandPool* pool = gpv::expand_pool_get();
	int bad = 0;
	for (int c = 0; c < calls; c++) {
		const int chunks = 1 + (int)(rng() % 6);
		std::vector<size_t> words(chunks);
		size_t total = 0;
		for (auto& w : words) { w = (rng() % 4 == 0) ? rng() % 7 : 512 + rng() % 60000; total += w; }
		std::vector<uint32_t> packed(total * 2 + 2);
		for (size_t i = 0; i < total; i++) { const uint32_t b = rng(); packed[2 * i + 1] = b; packed[2 * i] = rng() & ~b; }
		std::vector<uint8_t> out(total * 32 + 96, 0x5a);
		uint8_t* dst = out.data() + 32 + (rng() % 2) * 8; // aligned and unaligned destinations
		gpv::expand_pool_begin(pool);
		size_t at = 0;
		for (int k = 0; k < chunks; k++) {
			if (rng() % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 200)); // chunks land at their own pace
			gpv::expand_pool_submit(pool, packed.data() + 2 * at, dst + at * 32, words[k]);
			at += words[k];
		}
		gpv::expand_pool_end(pool);
		for (size_t i = 0; i < total && !bad; i++)
			for (int k = 0; k < 32; k++) {
				const uint8_t want = (packed[2 * i + 1] >> k & 1) ? 254 : (packed[2 * i] >> k & 1) ? 127 : 0;
				if (dst[i * 32 + k] != want) { bad = 1; fprintf(stderr, "call %d word %zu bit %d: got %d want %d\n", c, i, k, dst[i * 32 + k], want); break; }
			}
		for (uint8_t* p = out.data(); p < dst; p++) bad |= *p != 0x5a;
		for (uint8_t* p = dst + total * 32; p < out.data() + out.size(); p++) bad |= *p != 0x5a;
		if (bad) return 1;
	}
	return 0;
}

int main(int argc, char** argv)
{
	const int calls = argc > 1 ? atoi(argv[1]) : 60;
	int r1 = 0, r2 = 0;
	std::thread a([&] { r1 = run_client(1, calls); }), b([&] { r2 = run_client(2, calls); });
	a.join(); b.join();
	printf("%s (%d pool threads)\n", (r1 | r2) ? "FAILED" : "ok", gpv::expand_pool_threads(gpv::expand_pool_get()));
	return r1 | r2;
}
