// tests/cpu_probe/sanitize_driver.cpp -- TEST-ONLY: the host half of the product (gpview_b200/csrc/gpv_host.cpp: loaders, writer,
// voxel-file reader) compiled with -fsanitize=address,undefined and driven over a corpus of valid, truncated, mutated and
// random files.  The readers take files from the outside world; they may refuse one, never read or write out of bounds.
//   sanitize_driver meshes DIR      every *.obj / *.off in DIR: strict and tolerant reader, 1 and 3 loader threads
//   sanitize_driver voxels DIR N    gpv_load_voxels on DIR/m0 .. DIR/m<N-1> (object id 5)
#include "../../include/gpview_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <string>
#include <vector>

namespace gpv { // gpv::fail lives in the CUDA half of the library
static thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }
}
extern "C" const char* gpv_last_error(void) { return gpv::g_err.c_str(); }

int main(int argc, char** argv)
{
	if (argc < 3) return 2;
	long ok = 0, refused = 0;
	if (!strcmp(argv[1], "meshes")) {
		std::vector<std::string> files;
		if (DIR* d = opendir(argv[2])) {
			while (dirent* e = readdir(d)) { std::string n = e->d_name; if (n.size() > 4) files.push_back(std::string(argv[2]) + "/" + n); }
			closedir(d);
		}
		for (const std::string& f : files) for (int threads = 1; threads <= 3; threads += 2) for (unsigned flags = 0; flags < 2; flags++) {
			setenv("GPV_LOAD_THREADS", threads == 1 ? "1" : "3", 1);
			gpv_mesh m;
			if (gpv_load_mesh_ex(f.c_str(), flags, &m) == 0) {
				volatile float s = 0;
				for (long i = 0; i < m.n_tri * 9; i++) s = s + m.tris[i]; // every float of the result is readable
				gpv_free_mesh(&m);
				ok++;
			} else refused++;
		}
	} else {
		const int n = argc > 3 ? atoi(argv[3]) : 0;
		for (int i = 0; i < n; i++) {
			const std::string d = std::string(argv[2]) + "/m" + std::to_string(i);
			gpv_voxel_file v;
			if (gpv_load_voxels(d.c_str(), 5, &v) == 0) {
				volatile long s = 0;
				for (long k = 0; k < v.cells; k++) s = s + v.level1_inout[k];
				if (v.level2_inout) for (long k = 0; k < v.n_boundary * v.n23; k++) s = s + v.level2_inout[k];
				if (v.prefix_sum) for (long k = 0; k < v.cells; k++) s = s + v.prefix_sum[k];
				gpv_free_voxels(&v);
				ok++;
			} else refused++;
		}
	}
	printf("ok %ld refused %ld\n", ok, refused);
	return 0;
}
