// isect_check.cpp — TEST INFRASTRUCTURE (not part of the product): compiles the device's scene_intersect
// (simple-spectral_b200/csrc/ssb_isect.cuh: conservative packed filter + nearest-first exact tests) for the HOST and
// compares its hit record, bit for bit, with the reference's plain list scan (Scene::intersect, scene.cpp:433-445) on
// millions of rays: random, surface-to-surface, edge / corner / diagonal targeted, grazing, axis-aligned, tied
// (duplicated and coplanar quads), on the shipped scenes (quads passed in a file) and on synthetic ones (non-planar,
// degenerate, > 32 entries, tiny / huge / far-from-origin coordinates).
//
//   g++ -std=c++17 -O2 -ffp-contract=off -o isect_check tools/isect_check.cpp
//   isect_check [rays_per_scene] [quads.bin ...]        quads.bin = uint32 n, then n x ssb_quad
// Exit code 0 = all hit records identical.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#define SSB_ISECT_STATS 1
#include "../simple-spectral_b200/csrc/ssb_isect.cuh"

using namespace ssbk;

namespace {

std::mt19937_64 g_rng(12345);  // re-seeded from ISECT_SEED
double urand() { return std::uniform_real_distribution<double>(0.0, 1.0)(g_rng); }
double srand1() { return 2.0 * urand() - 1.0; }

struct V3 { float x, y, z; };
V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
V3 operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
V3 normalize(V3 a) { float l = std::sqrt(dot(a, a)); return l > 0 ? a * (1.0f / l) : V3{ 0, 0, 1 }; }
V3 vpos(const ssb_vertex& v) { return { v.pos[0], v.pos[1], v.pos[2] }; }
V3 rand_dir() {
	for (;;) {
		V3 d = { (float)srand1(), (float)srand1(), (float)srand1() };
		float l = dot(d, d);
		if (l > 1e-4f && l <= 1.0f) return normalize(d);
	}
}

ssb_quad make_quad(V3 v00, V3 v10, V3 v11, V3 v01) {  // geometry.hpp:93-95
	ssb_quad q{};
	auto set = [](ssb_vertex& v, V3 p) { v.pos[0] = p.x; v.pos[1] = p.y; v.pos[2] = p.z; v.st[0] = v.st[1] = 0; };
	set(q.tri[0].v[0], v00); set(q.tri[0].v[1], v10); set(q.tri[0].v[2], v11);
	set(q.tri[1].v[0], v00); set(q.tri[1].v[1], v11); set(q.tri[1].v[2], v01);
	for (int t = 0; t < 2; ++t) {
		V3 n = normalize(cross(vpos(q.tri[t].v[1]) - vpos(q.tri[t].v[0]), vpos(q.tri[t].v[2]) - vpos(q.tri[t].v[0])));
		q.tri[t].normal[0] = n.x; q.tri[t].normal[1] = n.y; q.tri[t].normal[2] = n.z;
	}
	return q;
}

struct Blob {
	std::vector<unsigned char> bytes;
	uint32_t nquads = 0;
};
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Blob build_blob(const std::vector<ssb_quad>& quads, const float* eye) {
	Blob b;
	b.nquads = (uint32_t)quads.size();
	DevHeader hdr{};
	hdr.nquads = b.nquads;
	const FilterTables ft = build_filter_tables(quads.data(), quads.size(), eye);
	ft.fill_header(hdr);
	size_t off = align_up(sizeof(DevHeader), 16);
	hdr.off_quads = (uint32_t)off; off = align_up(off + quads.size() * sizeof(ssb_quad), 128);
	hdr.off_fpairs = (uint32_t)off; off = align_up(off + ft.pairs.size() * 4, 16);
	hdr.off_planes = (uint32_t)off; off = align_up(off + ft.planes.size() * 4, 16);
	hdr.off_entry_quad = (uint32_t)off; off = align_up(off + ft.entry_quad.size() * 4, 16);
	hdr.off_quad_mask = (uint32_t)off; off = align_up(off + ft.quad_mask.size() * 4, 16);
	hdr.off_chunks = (uint32_t)off; off = align_up(off + ft.chunks.size() * 4, 16);
	hdr.total_bytes = (uint32_t)off;
	b.bytes.assign(off + 128, 0);
	unsigned char* p = b.bytes.data();
	memcpy(p, &hdr, sizeof(hdr));
	if (!quads.empty()) memcpy(p + hdr.off_quads, quads.data(), quads.size() * sizeof(ssb_quad));
	if (!ft.pairs.empty()) memcpy(p + hdr.off_fpairs, ft.pairs.data(), ft.pairs.size() * 4);
	if (!ft.planes.empty()) memcpy(p + hdr.off_planes, ft.planes.data(), ft.planes.size() * 4);
	if (!ft.entry_quad.empty()) memcpy(p + hdr.off_entry_quad, ft.entry_quad.data(), ft.entry_quad.size() * 4);
	if (!ft.quad_mask.empty()) memcpy(p + hdr.off_quad_mask, ft.quad_mask.data(), ft.quad_mask.size() * 4);
	if (!ft.chunks.empty()) memcpy(p + hdr.off_chunks, ft.chunks.data(), ft.chunks.size() * 4);
	return b;
}

struct Stats {
	unsigned long long rays = 0, hits = 0, mismatches = 0;
};

bool same(const Hit& a, const Hit& b) {
	if (a.quad != b.quad) return false;
	if (a.quad < 0) return true;
	return a.tri == b.tri && __float_as_uint(a.dist) == __float_as_uint(b.dist) && __float_as_uint(a.bx) == __float_as_uint(b.bx) &&
	       __float_as_uint(a.by) == __float_as_uint(b.by) && __float_as_uint(a.bz) == __float_as_uint(b.bz);
}

void check_ray(const char* scene, Stats& st, V3 o, V3 d, int ignore, float eps, Hit* out = nullptr) {
	const SceneView S;
	Hit a, b;
	scene_intersect(S, eps, ignore, a, o.x, o.y, o.z, d.x, d.y, d.z);
	scene_intersect_listscan(S, eps, ignore, b, o.x, o.y, o.z, d.x, d.y, d.z);
	st.rays++;
	if (b.quad >= 0) st.hits++;
	if (!same(a, b)) {
		if (st.mismatches < 10)
			fprintf(stderr, "MISMATCH %s: o=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g) ignore=%d: device quad %d tri %d dist %.9g | list scan quad %d tri %d dist %.9g\n",
			        scene, o.x, o.y, o.z, d.x, d.y, d.z, ignore, a.quad, a.tri, a.dist, b.quad, b.tri, b.dist);
		if (st.mismatches < 3 && b.quad >= 0 && getenv("ISECT_DEBUG")) {
			const DevHeader* H = S.hdr();
			for (uint32_t e = 0; e < H->nentries; ++e) {
				if ((int)S.entry_quad()[e] != b.quad) continue;
				const float* rec = reinterpret_cast<const float*>(S.fpairs()) + (e / 2) * 32;
				float r[16];
				for (int c = 0; c < 16; ++c) r[c] = rec[2 * c + (e & 1)];
				float nd = r[0] * d.x + r[1] * d.y + r[2] * d.z, no = r[0] * o.x + r[1] * o.y + r[2] * o.z;
				float tp = (r[3] - no) / nd;
				float px = o.x + tp * d.x, py = o.y + tp * d.y, pz = o.z + tp * d.z;
				float u = r[4] * px + r[5] * py + r[6] * pz + r[7], v = r[8] * px + r[9] * py + r[10] * pz + r[11];
				float sd = r[12] * u + r[13] * v + r[14];
				const uint32_t* ch = reinterpret_cast<const uint32_t*>(S.chunks()) + (e / 32) * 4;
				fprintf(stderr, "   entry %u: nd %g tp %g u %g v %g sd %g  q3 %g  chunk masks %08x %08x %08x margin %g\n", e, nd, tp, u, v, sd, tp * H->cull_margin_rneg, ch[0], ch[1], ch[2], H->cull_margin);
			}
		}
		st.mismatches++;
	}
	if (out) *out = b;
}

V3 quad_point(const ssb_quad& q, float s, float t) {  // bilinear over (v00, v10, v11, v01)
	V3 v00 = vpos(q.tri[0].v[0]), v10 = vpos(q.tri[0].v[1]), v11 = vpos(q.tri[0].v[2]), v01 = vpos(q.tri[1].v[2]);
	V3 a = v00 + (v10 - v00) * s, b = v01 + (v11 - v01) * s;
	return a + (b - a) * t;
}
float special01() {  // parameter values that land on corners, edges, the diagonal neighbourhood
	static const float sp[] = { 0.0f, 1.0f, 0.5f, 1e-7f, 1.0f - 1e-7f, 1e-4f, 1.0f - 1e-4f, 1e-3f, 0.25f, 0.75f };
	if (urand() < 0.5) return (float)urand();
	return sp[(int)(urand() * 10) % 10];
}

Stats run_scene(const char* name, const std::vector<ssb_quad>& quads, size_t nrays, float eps) {
	Stats st;
	const int nq = (int)quads.size();
	V3 lo = { 1e30f, 1e30f, 1e30f }, hi = { -1e30f, -1e30f, -1e30f };
	for (const ssb_quad& q : quads) for (int t = 0; t < 2; ++t) for (int v = 0; v < 3; ++v) {
		V3 p = vpos(q.tri[t].v[v]);
		lo = { std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z) };
		hi = { std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z) };
	}
	if (nq == 0) { lo = { 0, 0, 0 }; hi = { 0, 0, 0 }; }
	V3 ext = hi - lo;
	// the "camera": the farthest point a ray of this check starts from (the filter's margin is relative to it)
	const float eye[3] = { std::max(std::fabs(lo.x - 1.5f * ext.x), std::fabs(hi.x + 1.5f * ext.x)), std::max(std::fabs(lo.y - 1.5f * ext.y), std::fabs(hi.y + 1.5f * ext.y)),
	                       std::max(std::fabs(lo.z - 1.5f * ext.z), std::fabs(hi.z + 1.5f * ext.z)) };
	Blob blob = build_blob(quads, eye);
	ssb_smem = blob.bytes.data();
	if (nq == 0) { check_ray(name, st, { 0, 0, 0 }, { 0, 0, 1 }, -1, eps); return st; }
	auto rand_point = [&](float grow) {
		return V3{ lo.x + ext.x * (float)(urand() * (1 + 2 * grow) - grow), lo.y + ext.y * (float)(urand() * (1 + 2 * grow) - grow), lo.z + ext.z * (float)(urand() * (1 + 2 * grow) - grow) };
	};
	for (size_t r = 0; r < nrays; ++r) {
		static const bool paths_only = getenv("ISECT_PATHS") != nullptr;  // statistics of renderer-like rays: camera paths only
		const int kind = paths_only ? 3 : (int)(r % 8);
		V3 o, d;
		int ignore = -1;
		if (kind == 0) {  // anywhere, any direction
			o = rand_point(0.5f); d = rand_dir();
		} else if (kind == 1 || kind == 2) {  // surface to surface (targets include corners / edges / the diagonal s == t)
			const int qa = (int)(urand() * nq) % nq, qb = (int)(urand() * nq) % nq;
			o = quad_point(quads[qa], special01(), special01());
			float s = special01(), t = special01();
			if (urand() < 0.3) t = s;  // on the diagonal v00-v11
			V3 target = quad_point(quads[qb], s, t);
			d = normalize(target - o);
			if (!(dot(d, d) > 0.5f)) d = rand_dir();
			ignore = (kind == 1) ? qa : -1;
		} else if (kind == 3) {  // a path: bounce a few times like the renderer (origin = o + dist*d, ignore = hit quad)
			o = rand_point(0.2f); d = rand_dir();
			if (paths_only) { o = { lo.x + 0.5f * ext.x, lo.y + 0.5f * ext.y, lo.z - 1.4f * ext.z }; d = normalize(rand_point(0.0f) - o); }
			for (int bounce = 0; bounce < 6; ++bounce) {
				Hit h;
				check_ray(name, st, o, d, ignore, eps, &h);
				if (h.quad < 0) break;
				o = { o.x + h.dist * d.x, o.y + h.dist * d.y, o.z + h.dist * d.z };
				const ssb_tri& tr = quads[h.quad].tri[h.tri];
				V3 n = { tr.normal[0], tr.normal[1], tr.normal[2] };
				d = rand_dir();
				if (dot(d, n) < 0) d = d * -1.0f;
				ignore = h.quad;
			}
			continue;
		} else if (kind == 4) {  // grazing: direction (almost) in the plane of a quad
			const int qa = (int)(urand() * nq) % nq;
			const ssb_tri& tr = quads[qa].tri[(int)(urand() * 2) % 2];
			V3 n = { tr.normal[0], tr.normal[1], tr.normal[2] };
			V3 e = normalize(vpos(tr.v[1]) - vpos(tr.v[0])), f = cross(n, e);
			float ang = (float)(urand() * 6.2831853);
			V3 inpl = e * std::cos(ang) + f * std::sin(ang);
			static const float tilt[] = { 0.0f, 1e-7f, -1e-7f, 1e-6f, -2e-6f, 1e-5f, 1e-4f, -1e-3f, 0.01f, -0.03f, 0.049f, 0.051f, -0.05f, 0.1f };
			d = normalize(inpl + n * tilt[(int)(urand() * 14) % 14]);
			V3 target = quad_point(quads[qa], special01(), special01());
			float back = (float)urand() * std::min(ext.x, std::min(ext.y, ext.z)) * 1.4f;  // (stays inside the region `eye` covers)
			o = target - d * back;
			if (urand() < 0.3) { o = quad_point(quads[(int)(urand() * nq) % nq], (float)urand(), (float)urand()); }
		} else if (kind == 5) {  // axis-aligned directions (exact zeros) and rays along edges
			o = urand() < 0.5 ? rand_point(0.1f) : quad_point(quads[(int)(urand() * nq) % nq], special01(), special01());
			int ax = (int)(urand() * 3) % 3;
			float sgn = urand() < 0.5 ? 1.0f : -1.0f;
			d = { ax == 0 ? sgn : 0.0f, ax == 1 ? sgn : 0.0f, ax == 2 ? sgn : 0.0f };
			if (urand() < 0.3) { int a2 = (ax + 1) % 3; float v = 0.70710678f; d = { 0, 0, 0 }; (&d.x)[ax] = v * sgn; (&d.x)[a2] = v; }
		} else if (kind == 6) {  // from outside toward a quad point (camera-like)
			const int qb = (int)(urand() * nq) % nq;
			o = rand_point(1.5f);
			d = normalize(quad_point(quads[qb], special01(), special01()) - o);
			if (!(dot(d, d) > 0.5f)) d = rand_dir();
		} else {  // from a surface, random hemisphere direction, origin nudged by rounding-size offsets
			const int qa = (int)(urand() * nq) % nq;
			o = quad_point(quads[qa], (float)urand(), (float)urand());
			const ssb_tri& tr = quads[qa].tri[0];
			V3 n = { tr.normal[0], tr.normal[1], tr.normal[2] };
			d = rand_dir();
			if (dot(d, n) < 0) d = d * -1.0f;
			ignore = qa;
		}
		check_ray(name, st, o, d, ignore, eps);
	}
	const IsectStats is = g_isect_stats;
	g_isect_stats = IsectStats{ 0, 0, 0, 0, 0, 0 };
	printf("%-28s quads %3d entries %3u  rays %9llu  hits %9llu  mismatches %llu | nearest-first %.4f (extra tests/query %.4f) in-order %.4f  candidates/query %.2f exact tests/query %.3f\n",
	       name, nq, ((const DevHeader*)blob.bytes.data())->nentries, st.rays, st.hits, st.mismatches, (double)is.fast / is.queries, (double)is.fast_more / is.queries,
	       (double)is.inorder / is.queries, (double)is.candidates / is.queries, (double)is.exact_tests / is.queries);
	return st;
}

std::vector<ssb_quad> box_scene(float s, V3 c) {  // closed box + inner rotated block (shared edges, corners)
	std::vector<ssb_quad> q;
	auto P = [&](float x, float y, float z) { return V3{ c.x + s * x, c.y + s * y, c.z + s * z }; };
	q.push_back(make_quad(P(-1, -1, 1), P(-1, -1, -1), P(-1, 1, -1), P(-1, 1, 1)));
	q.push_back(make_quad(P(1, -1, -1), P(1, -1, 1), P(1, 1, 1), P(1, 1, -1)));
	q.push_back(make_quad(P(-1, -1, 1), P(1, -1, 1), P(1, -1, -1), P(-1, -1, -1)));
	q.push_back(make_quad(P(1, 1, 1), P(-1, 1, 1), P(-1, 1, -1), P(1, 1, -1)));
	q.push_back(make_quad(P(-1, -1, -1), P(1, -1, -1), P(1, 1, -1), P(-1, 1, -1)));
	q.push_back(make_quad(P(1, -1, 1), P(-1, -1, 1), P(-1, 1, 1), P(1, 1, 1)));
	const float top[4][2] = { { -0.3f, -0.5f }, { -0.6f, 0.1f }, { 0.0f, 0.4f }, { 0.3f, -0.2f } };
	q.push_back(make_quad(P(top[0][0], 0.2f, top[0][1]), P(top[1][0], 0.2f, top[1][1]), P(top[2][0], 0.2f, top[2][1]), P(top[3][0], 0.2f, top[3][1])));
	for (int k = 0; k < 4; ++k) {
		const float* a = top[k]; const float* b = top[(k + 1) % 4];
		q.push_back(make_quad(P(a[0], -1, a[1]), P(a[0], 0.2f, a[1]), P(b[0], 0.2f, b[1]), P(b[0], -1, b[1])));
	}
	return q;
}

std::vector<ssb_quad> random_scene(int n, float scale, V3 centre, double p_nonplanar, double p_degenerate, double p_duplicate) {
	std::vector<ssb_quad> q;
	while ((int)q.size() < n) {
		if (!q.empty() && urand() < p_duplicate) {  // exact duplicate or a coplanar overlapping copy: ties in distance
			ssb_quad c = q[(int)(urand() * q.size()) % q.size()];
			if (urand() < 0.5) {
				V3 e = vpos(c.tri[0].v[1]) - vpos(c.tri[0].v[0]);
				float sh = (float)(urand() * 0.5);
				for (int t = 0; t < 2; ++t) for (int v = 0; v < 3; ++v) { c.tri[t].v[v].pos[0] += sh * e.x; c.tri[t].v[v].pos[1] += sh * e.y; c.tri[t].v[v].pos[2] += sh * e.z; }
			}
			q.push_back(c);
			continue;
		}
		V3 o = { centre.x + scale * (float)srand1(), centre.y + scale * (float)srand1(), centre.z + scale * (float)srand1() };
		V3 e1 = rand_dir() * (scale * (float)(0.1 + urand())), e2 = rand_dir() * (scale * (float)(0.1 + urand()));
		if (urand() < 0.3) {  // axis-aligned rectangle
			int ax = (int)(urand() * 3) % 3;
			e1 = { 0, 0, 0 }; e2 = { 0, 0, 0 };
			(&e1.x)[(ax + 1) % 3] = scale * (float)(0.1 + urand()); (&e2.x)[(ax + 2) % 3] = scale * (float)(0.1 + urand());
		}
		V3 v00 = o, v10 = o + e1, v11 = o + e1 + e2, v01 = o + e2;
		if (urand() < 0.3) v11 = v11 + e1 * (float)(urand() * 0.5) + e2 * (float)(urand() * 0.5);  // not a parallelogram
		double r = urand();
		if (r < p_nonplanar) v01 = v01 + normalize(cross(e1, e2)) * (scale * (float)(srand1() * 0.2));
		else if (r < p_nonplanar + p_degenerate) {
			int kind = (int)(urand() * 4) % 4;
			if (kind == 0) v10 = v00;             // tri0 degenerate
			else if (kind == 1) v01 = v11;        // tri1 degenerate
			else if (kind == 2) { v10 = v00; v11 = v00; v01 = v00; }  // a point
			else v11 = v00 + (v10 - v00) * 2.0f;  // tri0 collinear
		}
		q.push_back(make_quad(v00, v10, v11, v01));
	}
	return q;
}

}  // namespace

int main(int argc, char** argv) {
	size_t nrays = argc > 1 ? (size_t)atoll(argv[1]) : 400000;
	if (const char* e = getenv("ISECT_SEED")) g_rng.seed((unsigned long long)atoll(e));
	unsigned long long bad = 0, total = 0;
	auto acc = [&](Stats s) { bad += s.mismatches; total += s.rays; };
	for (int a = 2; a < argc; ++a) {
		FILE* f = fopen(argv[a], "rb");
		if (!f) { fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
		uint32_t n = 0;
		if (fread(&n, 4, 1, f) != 1) { fclose(f); return 2; }
		std::vector<ssb_quad> quads(n);
		if (n && fread(quads.data(), sizeof(ssb_quad), n, f) != n) { fclose(f); return 2; }
		fclose(f);
		acc(run_scene(argv[a], quads, nrays * 4, 1e-3f));
	}
	acc(run_scene("box+block unit", box_scene(1.0f, { 0, 0, 0 }), nrays, 1e-3f));
	acc(run_scene("box+block x500 offset", box_scene(500.0f, { 300, 250, 280 }), nrays, 1e-3f));
	acc(run_scene("box+block x1e-2", box_scene(0.01f, { 0, 0, 0 }), nrays, 1e-7f));
	acc(run_scene("box+block far from origin", box_scene(1.0f, { 1000, -2000, 500 }), nrays, 1e-3f));
	acc(run_scene("random 12 planar", random_scene(12, 10.0f, { 0, 0, 0 }, 0, 0, 0), nrays, 1e-3f));
	acc(run_scene("random 16 with ties", random_scene(16, 10.0f, { 1, 2, 3 }, 0, 0, 0.4), nrays, 1e-3f));
	acc(run_scene("random 14 non-planar", random_scene(14, 5.0f, { 0, 0, 0 }, 0.5, 0, 0.1), nrays, 1e-3f));
	acc(run_scene("random 14 degenerate", random_scene(14, 5.0f, { 0, 0, 0 }, 0.2, 0.3, 0.1), nrays, 1e-3f));
	acc(run_scene("random 31 mixed", random_scene(31, 100.0f, { 50, 50, 50 }, 0.2, 0.1, 0.2), nrays, 1e-3f));
	acc(run_scene("random 45 (>32 entries)", random_scene(45, 20.0f, { 0, 0, 0 }, 0.3, 0.1, 0.2), nrays, 1e-3f));
	acc(run_scene("random 200 (>32 entries)", random_scene(200, 20.0f, { 0, 0, 0 }, 0.2, 0.05, 0.1), nrays / 4, 1e-3f));
	acc(run_scene("random 20 eps 1e-5", random_scene(20, 1.0f, { 0, 0, 0 }, 0.2, 0.1, 0.2), nrays, 1e-5f));
	acc(run_scene("single quad", random_scene(1, 1.0f, { 0, 0, 0 }, 0, 0, 0), nrays / 4, 1e-3f));
	acc(run_scene("empty", {}, 1, 1e-3f));
	printf("total rays %llu, mismatches %llu\n", total, bad);
	return bad ? 1 : 0;
}
