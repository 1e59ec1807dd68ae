// gpview_b200/csrc/gpv_batch.cpp -- batched dataset generation (BASELINE.json config 5: thousands of small .off/.obj meshes
// -> the six ObjN* files each, for 3-D CNN training sets, README.md:22-27 of the reference).
//
// The reference does this one GLUT session per model.  Here a pool of host threads, each with its own gpv_ctx (device
// buffers and stream), pinned staging buffers and file I/O, pulls paths from a shared counter: parse -> gpv_voxelize_host
// -> gpv_save.  Contexts on the same device overlap each other's kernels, copies and host work; devices are assigned
// round-robin, so one call drives every GPU of the box (models are independent: no collective).
// Restartable: with skip_existing, a model whose six-file set is complete (gpv_check_voxels) is not recomputed.
#include "../../include/gpview_b200.h"
#include "gpv_internal.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
// Contexts are expensive to make (streams, events, ~50 pooled device buffers that grow with the first models) and cheap to keep:
// a worker takes an idle context of its device from this cache and hands it back when the batch is done, so that the second batch
// of a process -- and every later one -- starts with warm pools.  gpv_batch_release() frees them.
struct CtxCache {
	std::mutex mu;
	std::vector<std::pair<int, gpv_ctx*>> idle;
	gpv_ctx* take(int device)
	{
		std::lock_guard<std::mutex> g(mu);
		for (size_t k = 0; k < idle.size(); k++)
			if (idle[k].first == device) { gpv_ctx* c = idle[k].second; idle.erase(idle.begin() + (long)k); return c; }
		return nullptr;
	}
	void give(int device, gpv_ctx* c) { std::lock_guard<std::mutex> g(mu); idle.emplace_back(device, c); }
	void clear()
	{
		std::lock_guard<std::mutex> g(mu);
		for (auto& e : idle) gpv_destroy(e.second);
		idle.clear();
	}
};
CtxCache& ctx_cache() { static CtxCache c; return c; }

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Pinned {
	void* p = nullptr; int64_t cap = 0;
	bool ensure(int64_t bytes) { if (bytes <= cap) return true; gpv_free_host(p); p = gpv_alloc_host(bytes + bytes / 4 + 4096); cap = p ? bytes + bytes / 4 + 4096 : 0; return p != nullptr; }
	~Pinned() { gpv_free_host(p); }
};
}

extern "C" int gpv_voxelize_batch(const char* const* paths, int64_t n_paths, const gpv_params* params, const int* devices, int n_devices, int threads,
                                  const char* out_dir, int first_obj_id, int skip_existing, gpv_batch_stats* stats)
{
	if (!paths || n_paths <= 0 || !params || n_devices <= 0 || threads <= 0) return gpv::fail("gpv_voxelize_batch: bad arguments");
	std::atomic<int64_t> next(0), done(0), failed(0), skipped(0), resizes(0);
	std::mutex errMu;
	std::string firstErr;
	double tParse = 0, tGpu = 0, tSave = 0;
	const double t0 = now();
	auto worker = [&](int w) {
		const int device = devices[w % n_devices];
		gpv_ctx* ctx = ctx_cache().take(device);
		if (!ctx && gpv_create(device, &ctx)) { std::lock_guard<std::mutex> g(errMu); if (firstErr.empty()) firstErr = gpv_last_error(); return; }
		Pinned l1, pre, bi, l2, n1, n2;
		double parse = 0, gpu = 0, save = 0;
		const bool wantN = (params->flags & GPV_NORMALS) != 0, wantL2 = !(params->flags & GPV_NO_LEVEL2) && params->voxel_count2 > 0;
		for (;;) {
			const int64_t i = next.fetch_add(1);
			if (i >= n_paths) break;
			const int objID = first_obj_id + (int)i;
			// restart: a model is skipped only when its set is COMPLETE -- the config parses and every stream has the size it implies
			// (gpv_check_voxels; gpv_save_streams writes the config last, so a killed run leaves none)
			if (skip_existing && out_dir && gpv_check_voxels(out_dir, objID) == 0) { skipped++; continue; }
			auto bail = [&]() { failed++; std::lock_guard<std::mutex> g(errMu); if (firstErr.empty()) firstErr = std::string(paths[i]) + ": " + gpv_last_error(); };
			double a = now();
			gpv_mesh mesh;
			if (gpv_load_mesh_ex(paths[i], (params->flags & GPV_BATCH_TOLERANT_LOAD) ? GPV_LOAD_TOLERANT : 0u, &mesh)) { bail(); continue; }
			double b = now();
			parse += b - a;
			gpv_grid g;
			if (gpv_make_grid(mesh.bbox_min, mesh.bbox_max, mesh.max_model_size, params->voxel_count, wantL2 ? params->voxel_count2 : 1, &g)) { bail(); gpv_free_mesh(&mesh); continue; }
			const int64_t cells = (int64_t)g.num_div[0] * g.num_div[1] * g.num_div[2], n23 = (int64_t)g.n2 * g.n2 * g.n2;
			int64_t capB = std::max<int64_t>(1024, l2.cap / std::max<int64_t>(1, n23)); // boundary cells the Level-2 buffer can take
			bool ok = false;
			gpv_result res;
			for (int attempt = 0; attempt < 2 && !ok; attempt++) {
				if (!l1.ensure(cells) || !pre.ensure(cells * 4) || !bi.ensure(capB * 4) || (wantL2 && !l2.ensure(capB * n23)) || (wantN && !n1.ensure(cells * 3)) ||
				    (wantN && wantL2 && !n2.ensure(capB * n23 * 3))) break;
				gpv_host_streams h{ (uint8_t*)l1.p, (int32_t*)pre.p, (int32_t*)bi.p, wantL2 ? (uint8_t*)l2.p : nullptr, wantN ? (uint8_t*)n1.p : nullptr,
					                (wantN && wantL2) ? (uint8_t*)n2.p : nullptr, capB * n23, capB };
				if (gpv_voxelize_host(ctx, &mesh, params, gpv_stream(ctx), &res, &h) == 0) {
					ok = true;
					double c = now();
					gpu += c - b;
					if (out_dir && gpv_save_streams(&mesh, &res, &h, objID, out_dir, (params->flags & GPV_SAVE_COMPUTED_ONLY) != 0)) ok = false;
					save += now() - c;
				} else {
					// too small a Level-2 buffer: the call reports it after the Level-1 pass; size it from a Level-1-only run
					gpv_params p1 = *params; p1.flags |= GPV_NO_LEVEL2; p1.flags &= ~GPV_NORMALS;
					gpv_host_streams none{};
					if (gpv_voxelize_host(ctx, &mesh, &p1, gpv_stream(ctx), &res, &none)) break;
					capB = res.n_boundary + 16;
					resizes++;
				}
			}
			if (!ok) bail(); else done++;
			gpv_free_mesh(&mesh);
		}
		ctx_cache().give(device, ctx);
		std::lock_guard<std::mutex> g(errMu);
		tParse += parse; tGpu += gpu; tSave += save;
	};
	std::vector<std::thread> pool;
	for (int w = 0; w < threads; w++) pool.emplace_back(worker, w);
	for (auto& t : pool) t.join();
	if (stats) {
		stats->models_done = done; stats->models_failed = failed; stats->models_skipped = skipped; stats->level2_resizes = resizes;
		stats->seconds = now() - t0; stats->parse_seconds = tParse; stats->gpu_seconds = tGpu; stats->save_seconds = tSave;
	}
	if (failed > 0 || (done == 0 && skipped == 0)) return gpv::fail(firstErr.empty() ? "gpv_voxelize_batch: no model processed" : firstErr);
	return 0;
}

extern "C" void gpv_batch_release(void) { ctx_cache().clear(); }
