// gpview_b200/csrc/gpv_expand.cpp -- host half of the 2-bit packed Level-2 transfer of gpv_voxelize_host (GPV_PACKED_L2).
//
// The reference moves Level2InOut to the host as 4 B per sub-voxel and converts it to bytes there (float -> uchar * 127,
// src/Object.cpp:2609, 3029-3051).  The file contract is 1 B per sub-voxel, and that byte has three values.  So the device sends
// 2 bits per sub-voxel -- one (inside mask, boundary mask) pair of 32-bit words per 32 consecutive sub-voxels of Level2InOut.raw,
// written by k_l2<.., L2_OUT_PACKED> -- a quarter of the PCIe bytes, and a pool of host threads expands every chunk into the
// caller's byte stream while the next chunk is still on the bus.  The pool's threads are started once per process, sleep
// between calls and spin only while a call is in flight (a condition-variable wake-up per chunk would cost as much as the chunk).
#include "gpv_internal.h"
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace gpv {

namespace {

// byte k of v[x] = bit k of x   (8 bits -> 8 bytes of 0 / 1)
struct Lut { uint64_t v[256]; Lut() { for (int x = 0; x < 256; x++) { uint64_t r = 0; for (int k = 0; k < 8; k++) if (x >> k & 1) r |= 1ull << (8 * k); v[x] = r; } } };
const Lut kLut;

void expand_scalar(const uint32_t* packed, uint8_t* out, size_t nWords)
{
	for (size_t w = 0; w < nWords; w++) {
		const uint32_t mi = packed[2 * w], mb = packed[2 * w + 1];
		for (int k = 0; k < 4; k++) {
			const uint64_t v = kLut.v[(mi >> (8 * k)) & 0xffu] * 127u + kLut.v[(mb >> (8 * k)) & 0xffu] * 254u;
			memcpy(out + w * 32 + k * 8, &v, 8);
		}
	}
}

#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw"))) void expand_avx512(const uint32_t* packed, uint8_t* out, size_t nWords)
{
	size_t w = 0;
	if (((uintptr_t)out & 63) == 32 && nWords) { expand_scalar(packed, out, 1); w = 1; } // 32-byte phase -> 64-byte lines
	const bool aligned = ((uintptr_t)(out + w * 32) & 63) == 0;
	const __m512i in = _mm512_set1_epi8(127), bd = _mm512_set1_epi8((char)254);
	for (; w + 2 <= nWords; w += 2) {
		const __mmask64 mi = (uint64_t)packed[2 * w] | ((uint64_t)packed[2 * w + 2] << 32);
		const __mmask64 mb = (uint64_t)packed[2 * w + 1] | ((uint64_t)packed[2 * w + 3] << 32);
		const __m512i v = _mm512_or_si512(_mm512_maskz_mov_epi8(mi, in), _mm512_maskz_mov_epi8(mb, bd));
		if (aligned) _mm512_stream_si512((__m512i*)(out + w * 32), v); // the stream is written once and read by someone else: no read-for-ownership
		else _mm512_storeu_si512((void*)(out + w * 32), v);
	}
	if (w < nWords) expand_scalar(packed + 2 * w, out + w * 32, nWords - w);
	_mm_sfence();
}

__attribute__((target("avx2"))) void expand_avx2(const uint32_t* packed, uint8_t* out, size_t nWords)
{
	const __m256i sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3);
	const __m256i bit = _mm256_set1_epi64x((long long)0x8040201008040201ull);
	const __m256i in = _mm256_set1_epi8(127), bd = _mm256_set1_epi8((char)254);
	const bool aligned = ((uintptr_t)out & 31) == 0;
	for (size_t w = 0; w < nWords; w++) {
		const __m256i a = _mm256_shuffle_epi8(_mm256_set1_epi32((int)packed[2 * w]), sel), b = _mm256_shuffle_epi8(_mm256_set1_epi32((int)packed[2 * w + 1]), sel);
		const __m256i ma = _mm256_cmpeq_epi8(_mm256_and_si256(a, bit), bit), mb = _mm256_cmpeq_epi8(_mm256_and_si256(b, bit), bit);
		const __m256i v = _mm256_or_si256(_mm256_and_si256(ma, in), _mm256_and_si256(mb, bd));
		if (aligned) _mm256_stream_si256((__m256i*)(out + w * 32), v);
		else _mm256_storeu_si256((__m256i*)(out + w * 32), v);
	}
	_mm_sfence();
}
#endif

using ExpandFn = void (*)(const uint32_t*, uint8_t*, size_t);
ExpandFn pick_expand()
{
	const char* force = getenv("GPV_EXPAND_ISA"); // "scalar" / "avx2": tests of the narrower paths on a wider machine
#if defined(__x86_64__)
	__builtin_cpu_init();
	if (force && !strcmp(force, "scalar")) return expand_scalar;
	if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") && !(force && !strcmp(force, "avx2"))) return expand_avx512;
	if (__builtin_cpu_supports("avx2")) return expand_avx2;
#endif
	(void)force;
	return expand_scalar;
}

} // namespace

void expand_packed(const void* packed, uint8_t* out, size_t nWords)
{
	static const ExpandFn fn = pick_expand();
	fn(static_cast<const uint32_t*>(packed), out, nWords);
}

// ---- the pool
struct ExpandPool {
	struct Task { const uint32_t* src; uint8_t* dst; size_t nWords; };
	std::mutex callMu;                       // one gpv_voxelize_host call at a time owns the pool
	std::mutex mu; std::condition_variable cv;
	std::vector<std::thread> workers;
	std::vector<Task> tasks;                 // ring: task i lives in tasks[i % size]; workers read entries below `published` without a lock
	std::atomic<size_t> published{ 0 }, next{ 0 }, done{ 0 }; // monotonic over the life of the pool (never reset: a worker may be between two loads)
	std::atomic<int> active{ 0 };            // > 0 while a call is in flight (workers spin), 0: sleep
	std::atomic<bool> quit{ false };
	unsigned long long generation = 0;

	explicit ExpandPool(int n)
	{
		tasks.resize(1 << 16);
		for (int k = 0; k < n; k++) workers.emplace_back([this] { run(); });
	}
	~ExpandPool()
	{
		{ std::lock_guard<std::mutex> g(mu); quit = true; }
		cv.notify_all();
		for (auto& t : workers) t.join();
	}
	void run()
	{
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [this] { return quit.load() || active.load() > 0; });
				if (quit.load()) return;
			}
			while (active.load(std::memory_order_acquire) > 0) { // a call is in flight: take tasks as they are published
				const size_t have = published.load(std::memory_order_acquire);
				size_t i = next.load(std::memory_order_relaxed);
				if (i < have && next.compare_exchange_weak(i, i + 1, std::memory_order_acq_rel)) {
					const Task t = tasks[i % tasks.size()];
					expand_packed(t.src, t.dst, t.nWords);
					done.fetch_add(1, std::memory_order_release);
				} else {
#if defined(__x86_64__)
					_mm_pause();
#endif
				}
			}
		}
	}
};

ExpandPool* expand_pool_get()
{
	static ExpandPool* pool = [] {
		int n = (int)std::thread::hardware_concurrency();
		if (const char* e = getenv("GPV_HOST_THREADS")) n = atoi(e) + 1;
		n = n > 1 ? n - 1 : 1;                 // the calling thread keeps a core for the CUDA events it waits on
		if (!getenv("GPV_HOST_THREADS") && n > 15) n = 15; // streaming stores saturate the memory channels well before 32 cores do; more spinners only get in the way
		if (n > 63) n = 63;
		return new ExpandPool(n);             // lives as long as the process (worker threads sleep between calls)
	}();
	return pool;
}

int expand_pool_threads(ExpandPool* p) { return (int)p->workers.size(); }

void expand_pool_begin(ExpandPool* p)
{
	p->callMu.lock();
	{ std::lock_guard<std::mutex> g(p->mu); p->active.store(1); }
	p->cv.notify_all();
}

// [packed, +nWords uint2) -> out; cut into pieces so that every worker gets a share of every chunk
void expand_pool_submit(ExpandPool* p, const void* packed, uint8_t* out, size_t nWords)
{
	const size_t nw = p->workers.size();
	size_t piece = (nWords + 2 * nw - 1) / (2 * nw);
	piece = (piece + 1) & ~(size_t)1;        // whole 64-byte lines
	if (piece < 2048) piece = 2048;         // >= 64 KB of output per task
	size_t at = p->published.load(std::memory_order_relaxed);
	for (size_t w = 0; w < nWords; w += piece) {
		if (at - p->done.load(std::memory_order_acquire) >= p->tasks.size() - 1) { // ring full (cannot happen with <= 16 chunks per call): expand on the spot
			expand_packed(static_cast<const uint32_t*>(packed) + 2 * w, out + w * 32, nWords - w);
			break;
		}
		p->tasks[at++ % p->tasks.size()] = { static_cast<const uint32_t*>(packed) + 2 * w, out + w * 32, std::min(piece, nWords - w) };
	}
	p->published.store(at, std::memory_order_release);
}

void expand_pool_end(ExpandPool* p)
{
	const size_t total = p->published.load();
	// the calling thread helps with what is left, then waits for the pieces in flight
	for (;;) {
		size_t i = p->next.load(std::memory_order_relaxed);
		if (i >= total) break;
		if (p->next.compare_exchange_weak(i, i + 1, std::memory_order_acq_rel)) {
			const ExpandPool::Task t = p->tasks[i % p->tasks.size()];
			expand_packed(t.src, t.dst, t.nWords);
			p->done.fetch_add(1, std::memory_order_release);
		}
	}
	while (p->done.load(std::memory_order_acquire) < total) {
#if defined(__x86_64__)
		_mm_pause();
#endif
	}
	p->active.store(0, std::memory_order_release);
	p->callMu.unlock();
}

} // namespace gpv
