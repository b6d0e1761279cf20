// ssb_blob.hpp — layout of the device "blob" (scene, materials, spectra, filter tables) that every kernel stages
// into shared memory, and the host-side construction of the conservative FILTER TABLES used by scene_intersect
// (ssb_isect.cuh).  Plain C++: included by the CUDA translation unit (ssb_capi.cu) and by the host-compiled
// check of the intersection logic (tools/isect_check.cpp).
//
// Filter entries.  The reference scans every primitive with the exact watertight test (Scene::intersect,
// scene.cpp:433-445; PrimQuad::intersect, geometry.cpp:128-139).  The device scan first runs a conservative filter
// over "entries":
//   * a planar quad is ONE entry: its plane, its bounding rectangle in two in-plane axes (enlarged by the margin),
//     and the side of the shared diagonal v00-v11 that tells tri0 (v00,v10,v11) from tri1 (v00,v11,v01);
//   * a non-planar quad (the red wall of the Cornell box is one: 3.2 units off its own plane) is TWO entries, one per
//     triangle, each with that triangle's own plane / rectangle / diagonal side — so that it no longer costs two
//     exact tests per ray;
//   * a degenerate quad is one all-zero entry, which the filter always keeps (both triangles).
// Entries are kept in list order (quad order, tri0 before tri1).  Two entries are packed per 128-byte record,
// component-interleaved, so that the device evaluates two entries with one packed-fp32 instruction (FFMA2).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/ssb200.h"

#if !defined(__CUDACC__)
struct uchar4;  // only pointers to them appear below
struct float4;
#endif

namespace ssbk {

struct DevSpectrum {  // _Spectrum (spectrum.hpp:12-70), data in the float pool
	uint32_t offset;    // index of the first sample in the pool
	uint32_t n_filter;  // n | (nearest ? 1u<<31 : 0)
	float low;
	float recip;        // _delta_lambda_recip = float(n-1)/(high-low), spectrum.cpp:22-25
};
struct DevMaterial {
	uint32_t kind, albedo_mode, texture, pad;
	DevSpectrum albedo, emission;
	float albedo_rgb[4], emission_rgb[4];  // RENDER_MODE_RGB constants (4th = 0)
};
struct DevTexture {
	const uchar4* rgba;  // RGB8 re-packed to RGBA8 at upload: one aligned 4-byte load per texel
	uint32_t width, height;
	const float4* coef;  // ssb_options.prebaked_textures: Jakob-Hanika coefficients (c0,c1,c2,-) per texel, or nullptr
};
struct DevHeader {
	uint32_t nquads, nmaterials, nlights, ntextures;
	uint32_t off_quads, off_materials, off_lights, off_textures, off_pool;  // byte offsets from blob start
	uint32_t total_bytes;  // multiple of 16
	// filter tables (see above); chunk = 32 entries = 16 pair records
	uint32_t nentries;        // padded to an even count
	uint32_t off_fpairs;      // [nentries/2] x 32 floats: {plane n.xyz, w | axis a.xyz/hu, -cu/hu | axis b.xyz/hv, -cv/hv | diagonal A,B,C, 0}, each component as {entry 2p, entry 2p+1}
	uint32_t off_planes;      // [nentries] float4: plane (n, w) again, un-interleaved (nearest-candidate pass)
	uint32_t off_entry_quad;  // [nentries] uint32: quad of the entry
	uint32_t off_quad_mask;   // [nquads] uint32: entries of the quad as a bit mask (only meaningful when nentries <= 32)
	uint32_t off_chunks;      // [ceil(nentries/32)] uint4: {entries that may yield tri0, entries that may yield tri1,
	                          //   tri0-entries whose quad's tri1-entry is the next entry, 0}
	float cull_margin;        // 1e-4 x scene extent
	float cull_margin_rneg;   // -1 / cull_margin
	float scene_centre[3];    // centre of the scene's bounding box
	float scene_radius;       // its half diagonal + margin
	float tmax_scale;         // 1.001 / cull_margin
	DevSpectrum xbar, ybar, zbar, basis_r, basis_g, basis_b;
	float srgb_lut[256];  // Color::srgb_to_lrgb(v/255) for v = 0..255 (color.hpp:91-97), built with the host libm
};

// ------------------------------------------------------------------ host side: build the filter tables
struct FilterTables {
	std::vector<float> pairs;           // 32 floats per pair record
	std::vector<float> planes;          // 4 floats per entry
	std::vector<uint32_t> entry_quad;   // per entry
	std::vector<uint32_t> quad_mask;    // per quad
	std::vector<uint32_t> chunks;       // 4 per chunk
	uint32_t nentries = 0;              // even
	float margin = 0.0f;
	float centre[3] = { 0, 0, 0 }, radius = 0.0f;

	void fill_header(DevHeader& hdr) const {
		hdr.nentries = nentries;
		hdr.cull_margin = margin;
		hdr.cull_margin_rneg = -1.0f / margin;
		for (int k = 0; k < 3; ++k) hdr.scene_centre[k] = centre[k];
		hdr.scene_radius = radius;
		hdr.tmax_scale = 1.001f / margin;
	}
};

namespace blob_detail {

struct Entry {
	float r[16];      // plane(4) | a/hu, -cu/hu | b/hv, -cv/hv | diagonal(3), 0
	uint32_t quad;
	bool tri0, tri1;  // which triangles the entry stands for
	bool pair_first;  // tri0-entry of a split quad (the tri1-entry follows)
};

inline Entry zero_entry(uint32_t quad, bool t0, bool t1) {
	Entry e;
	for (float& v : e.r) v = 0.0f;
	e.quad = quad; e.tri0 = t0; e.tri1 = t1; e.pair_first = false;
	return e;
}

// Plane through P[0],P[1],P[2] (double precision, from the vertices — not from the stored float normal); bounding
// rectangle of all `n` points in the axes a = dir(P[1]-P[0]), b = n x a.  Returns false for a degenerate triangle.
inline bool plane_and_rect(const double (*P)[3], int n, double diag, double margin, float* r, double* a, double* b, double& hu, double& hv, double& cu, double& cv) {
	double e1[3], e2[3], nn[3];
	for (int k = 0; k < 3; ++k) { e1[k] = P[1][k] - P[0][k]; e2[k] = P[2][k] - P[0][k]; }
	nn[0] = e1[1] * e2[2] - e1[2] * e2[1]; nn[1] = e1[2] * e2[0] - e1[0] * e2[2]; nn[2] = e1[0] * e2[1] - e1[1] * e2[0];
	const double nl = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
	const double e1l = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
	if (!(nl > 1e-12 * diag * diag) || !(e1l > 0)) return false;
	for (int k = 0; k < 3; ++k) nn[k] /= nl;
	const double w = nn[0] * P[0][0] + nn[1] * P[0][1] + nn[2] * P[0][2];
	for (int k = 0; k < 3; ++k) a[k] = e1[k] / e1l;
	b[0] = nn[1] * a[2] - nn[2] * a[1]; b[1] = nn[2] * a[0] - nn[0] * a[2]; b[2] = nn[0] * a[1] - nn[1] * a[0];
	double ulo = INFINITY, uhi = -INFINITY, vlo = INFINITY, vhi = -INFINITY;
	for (int v = 0; v < n; ++v) {
		const double u = a[0] * P[v][0] + a[1] * P[v][1] + a[2] * P[v][2], vv = b[0] * P[v][0] + b[1] * P[v][1] + b[2] * P[v][2];
		ulo = std::min(ulo, u); uhi = std::max(uhi, u); vlo = std::min(vlo, vv); vhi = std::max(vhi, vv);
	}
	hu = 0.5 * (uhi - ulo) + margin; hv = 0.5 * (vhi - vlo) + margin; cu = 0.5 * (uhi + ulo); cv = 0.5 * (vhi + vlo);
	r[0] = (float)nn[0]; r[1] = (float)nn[1]; r[2] = (float)nn[2]; r[3] = (float)w;
	r[4] = (float)(a[0] / hu); r[5] = (float)(a[1] / hu); r[6] = (float)(a[2] / hu); r[7] = (float)(-cu / hu);
	r[8] = (float)(b[0] / hv); r[9] = (float)(b[1] / hv); r[10] = (float)(b[2] / hv); r[11] = (float)(-cv / hv);
	r[12] = r[13] = r[14] = r[15] = 0.0f;
	return true;
}

// Signed distance to the line through D0,D1 in units of the margin, as a function of the scaled (u,v) of the entry:
// positive on the side of `pos_side`.  Returns false when the point is not clearly (> margin) off the line.
inline bool diagonal_record(const double* a, const double* b, double hu, double hv, double cu, double cv, double margin,
                            const double* D0, const double* D1, const double* side_pt, bool side_positive, float* r) {
	auto U = [&](const double* p) { return a[0] * p[0] + a[1] * p[1] + a[2] * p[2]; };
	auto V = [&](const double* p) { return b[0] * p[0] + b[1] * p[1] + b[2] * p[2]; };
	const double u0 = U(D0), v0 = V(D0), u1 = U(D1), v1 = V(D1);
	const double ex = u1 - u0, ey = v1 - v0, el = std::sqrt(ex * ex + ey * ey);
	if (!(el > 0)) return false;
	double nx = ey / el, ny = -ex / el;
	double ds = nx * (U(side_pt) - u0) + ny * (V(side_pt) - v0);
	if ((ds < 0) == side_positive) { nx = -nx; ny = -ny; ds = -ds; }
	if (!(std::fabs(ds) > margin)) return false;
	r[12] = (float)(nx * hu / margin); r[13] = (float)(ny * hv / margin);
	r[14] = (float)((nx * (cu - u0) + ny * (cv - v0)) / margin);
	return true;
}

}  // namespace blob_detail

// `eye`: a ray origin that is not on the scene's surfaces (the camera position), or nullptr.  The margin is relative to
// the magnitude of every coordinate the filter computes with.
inline FilterTables build_filter_tables(const ssb_quad* quads, size_t nquads, const float* eye) {
	using namespace blob_detail;
	FilterTables ft;
	float slo[3] = { INFINITY, INFINITY, INFINITY }, shi[3] = { -INFINITY, -INFINITY, -INFINITY };
	double maxabs = 0.0;
	for (size_t qi = 0; qi < nquads; ++qi)
		for (int t = 0; t < 2; ++t) for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) {
			const float c = quads[qi].tri[t].v[v].pos[k];
			slo[k] = std::min(slo[k], c); shi[k] = std::max(shi[k], c);
			maxabs = std::max(maxabs, (double)std::fabs(c));
		}
	if (eye) for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, (double)std::fabs(eye[k]));
	double diag = 0.0;
	if (nquads) for (int k = 0; k < 3; ++k) ft.centre[k] = 0.5f * (slo[k] + shi[k]);
	if (nquads) diag = std::sqrt((double)(shi[0] - slo[0]) * (shi[0] - slo[0]) + (double)(shi[1] - slo[1]) * (shi[1] - slo[1]) + (double)(shi[2] - slo[2]) * (shi[2] - slo[2]));
	// 1e-4 of the scene extent: orders of magnitude more than the rounding of the watertight test or of the filter
	// (both ~1e-7 of the coordinate magnitude, which is why the extent includes the distance from the origin)
	const double margin = 1e-4 * std::max(diag, maxabs) + 1e-6;
	ft.margin = (float)margin;
	ft.radius = (float)(0.5 * diag * 1.0001 + 2.0 * margin);

	std::vector<Entry> entries;
	for (size_t qi = 0; qi < nquads; ++qi) {
		const ssb_quad& q = quads[qi];
		double P[6][3];
		for (int t = 0; t < 2; ++t) for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) P[t * 3 + v][k] = q.tri[t].v[v].pos[k];
		// the two triangles share the diagonal v00-v11 in the reference's construction (geometry.hpp:93-95)
		bool shared = true;
		for (int k = 0; k < 3; ++k) shared = shared && q.tri[1].v[0].pos[k] == q.tri[0].v[0].pos[k] && q.tri[1].v[1].pos[k] == q.tri[0].v[2].pos[k];
		double a[3], b[3], hu, hv, cu, cv;
		Entry e = zero_entry((uint32_t)qi, true, true);
		bool ok = plane_and_rect(P, 6, diag, margin, e.r, a, b, hu, hv, cu, cv);
		bool planar = ok;
		if (ok) {
			for (int v = 0; v < 6; ++v)
				if (std::fabs((double)e.r[0] * P[v][0] + (double)e.r[1] * P[v][1] + (double)e.r[2] * P[v][2] - (double)e.r[3]) > 1e-6 * diag) planar = false;
		}
		if (planar) {
			// which triangle: only when the two really lie on opposite sides of the shared diagonal; else both stay candidates
			if (shared) {
				float keep[3] = { 0, 0, 0 };
				float tmp[16] = { 0 };
				if (diagonal_record(a, b, hu, hv, cu, cv, margin, P[0], P[2], P[1], true, tmp)) {
					// tri1's third vertex (v01) must be clearly on the other side
					const double sd01 = (double)tmp[12] * ((a[0] * P[5][0] + a[1] * P[5][1] + a[2] * P[5][2] - cu) / hu) +
					                    (double)tmp[13] * ((b[0] * P[5][0] + b[1] * P[5][1] + b[2] * P[5][2] - cv) / hv) + (double)tmp[14];
					if (sd01 < -1.0) { keep[0] = tmp[12]; keep[1] = tmp[13]; keep[2] = tmp[14]; }
				}
				e.r[12] = keep[0]; e.r[13] = keep[1]; e.r[14] = keep[2];
			}
			entries.push_back(e);
			continue;
		}
		// non-planar (or tri0 degenerate): one entry per triangle when both are proper triangles
		Entry e0 = zero_entry((uint32_t)qi, true, false), e1 = zero_entry((uint32_t)qi, false, true);
		double a0[3], b0[3], a1[3], b1[3], hu0, hv0, cu0, cv0, hu1, hv1, cu1, cv1;
		const bool ok0 = plane_and_rect(P, 3, diag, margin, e0.r, a0, b0, hu0, hv0, cu0, cv0);
		const bool ok1 = plane_and_rect(P + 3, 3, diag, margin, e1.r, a1, b1, hu1, hv1, cu1, cv1);
		if (!ok0 || !ok1) {  // a degenerate triangle: the all-zero entry keeps both triangles of the quad as candidates
			entries.push_back(zero_entry((uint32_t)qi, true, true));
			continue;
		}
		if (shared) {
			// tri0 = (v00,v10,v11): inside is the v10 side of v00-v11, reported as sd >= 0 (tested with sd < -1);
			// tri1 = (v00,v11,v01): inside is the v01 side, reported as sd <= 0 (tested with sd > 1)
			float tmp[16] = { 0 };
			if (diagonal_record(a0, b0, hu0, hv0, cu0, cv0, margin, P[0], P[2], P[1], true, tmp)) { e0.r[12] = tmp[12]; e0.r[13] = tmp[13]; e0.r[14] = tmp[14]; }
			float tmp1[16] = { 0 };
			if (diagonal_record(a1, b1, hu1, hv1, cu1, cv1, margin, P[3], P[4], P[5], false, tmp1)) { e1.r[12] = tmp1[12]; e1.r[13] = tmp1[13]; e1.r[14] = tmp1[14]; }
		}
		e0.pair_first = true;
		entries.push_back(e0);
		entries.push_back(e1);
	}
	// a split quad's two entries must not straddle a 32-entry chunk: pad with an inert entry
	std::vector<Entry> laid;
	for (size_t i = 0; i < entries.size(); ++i) {
		if (entries[i].pair_first && (laid.size() % 32) == 31) laid.push_back(zero_entry(entries[i].quad, false, false));
		laid.push_back(entries[i]);
	}
	if (laid.size() % 2) laid.push_back(zero_entry(laid.empty() ? 0u : laid.back().quad, false, false));
	const size_t R = laid.size();
	ft.nentries = (uint32_t)R;
	ft.pairs.assign(R / 2 * 32, 0.0f);
	ft.planes.assign(R * 4, 0.0f);
	ft.entry_quad.assign(R, 0u);
	ft.quad_mask.assign(nquads, 0u);
	ft.chunks.assign((R + 31) / 32 * 4, 0u);
	for (size_t i = 0; i < R; ++i) {
		const Entry& e = laid[i];
		float* rec = ft.pairs.data() + (i / 2) * 32;
		for (int c = 0; c < 16; ++c) rec[2 * c + (i & 1)] = e.r[c];
		for (int c = 0; c < 4; ++c) ft.planes[4 * i + c] = e.r[c];
		ft.entry_quad[i] = e.quad;
		const uint32_t bit = 1u << (i % 32);
		uint32_t* ch = ft.chunks.data() + (i / 32) * 4;
		if (e.tri0) ch[0] |= bit;
		if (e.tri1) ch[1] |= bit;
		if (e.pair_first) ch[2] |= bit;
		if (R <= 32 && (e.tri0 || e.tri1)) ft.quad_mask[e.quad] |= bit;
	}
	return ft;
}

}  // namespace ssbk
