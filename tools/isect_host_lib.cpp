// isect_host_lib.cpp — TEST INFRASTRUCTURE (not part of the product): the device's scene_intersect
// (simple-spectral_b200/csrc/ssb_isect.cuh) compiled for the HOST behind a C entry point, so that the CPU test suite can
// run it on the same random scenes and ray families as the device fuzz test (tests/test_gpu_isect_fuzz.py) and compare
// with the checker's list scan (oracle/ssb_oracle.c; reference scene.cpp:433-445).  The packed-fp32 and rcp.approx
// instructions are emulated (exact reciprocal, fmaf): what this pins down is the LOGIC of the filter.
//   g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -o libisect_host.so tools/isect_host_lib.cpp
#include <cstdint>
#include <cstring>
#include <vector>

#include "../simple-spectral_b200/csrc/ssb_isect.cuh"

using namespace ssbk;

extern "C" int isect_host(const ssb_quad* quads, uint32_t nquads, const float* rays6, const int32_t* ignore, uint32_t list_scan, float eps,
                          float* out6, size_t n) {
	auto align_up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
	DevHeader hdr{};
	hdr.nquads = nquads;
	const float eye[3] = { 0, 0, 0 };
	const FilterTables ft = build_filter_tables(quads, nquads, eye);
	ft.fill_header(hdr);
	size_t off = align_up(sizeof(DevHeader), 16);
	hdr.off_quads = (uint32_t)off; off = align_up(off + nquads * sizeof(ssb_quad), 128);
	hdr.off_fpairs = (uint32_t)off; off = align_up(off + ft.pairs.size() * 4, 16);
	hdr.off_planes = (uint32_t)off; off = align_up(off + ft.planes.size() * 4, 16);
	hdr.off_entry_quad = (uint32_t)off; off = align_up(off + ft.entry_quad.size() * 4, 16);
	hdr.off_quad_mask = (uint32_t)off; off = align_up(off + ft.quad_mask.size() * 4, 16);
	hdr.off_chunks = (uint32_t)off; off = align_up(off + ft.chunks.size() * 4, 16);
	hdr.total_bytes = (uint32_t)off;
	std::vector<unsigned char> blob(off + 128, 0);
	unsigned char* p = blob.data();
	memcpy(p, &hdr, sizeof(hdr));
	memcpy(p + hdr.off_quads, quads, nquads * sizeof(ssb_quad));
	memcpy(p + hdr.off_fpairs, ft.pairs.data(), ft.pairs.size() * 4);
	memcpy(p + hdr.off_planes, ft.planes.data(), ft.planes.size() * 4);
	memcpy(p + hdr.off_entry_quad, ft.entry_quad.data(), ft.entry_quad.size() * 4);
	memcpy(p + hdr.off_quad_mask, ft.quad_mask.data(), ft.quad_mask.size() * 4);
	memcpy(p + hdr.off_chunks, ft.chunks.data(), ft.chunks.size() * 4);
	ssb_smem = p;
	const SceneView S;
	for (size_t r = 0; r < n; ++r) {
		Hit h;
		const float* q = rays6 + 6 * r;
		if (list_scan) scene_intersect_listscan(S, eps, ignore ? ignore[r] : -1, h, q[0], q[1], q[2], q[3], q[4], q[5]);
		else scene_intersect(S, eps, ignore ? ignore[r] : -1, h, q[0], q[1], q[2], q[3], q[4], q[5]);
		int32_t qd = h.quad, tr = h.tri;
		memcpy(out6 + 6 * r, &qd, 4); memcpy(out6 + 6 * r + 1, &tr, 4);
		out6[6 * r + 2] = h.dist; out6[6 * r + 3] = h.bx; out6[6 * r + 4] = h.by; out6[6 * r + 5] = h.bz;
	}
	ssb_smem = nullptr;
	return 0;
}
