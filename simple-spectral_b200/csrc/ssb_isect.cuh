// ssb_isect.cuh — Scene::intersect (scene.cpp:433-445) + PrimQuad::intersect (geometry.cpp:128-139) +
// PrimTri::intersect (geometry.cpp:12-101) for the device, written so that the SAME source also compiles for the
// host (tools/isect_check.cpp runs it against the brute-force list scan on millions of rays: CPU test
// tests/test_isect_host.py).  Included by ssb_kernels.cuh inside namespace ssbk.
//
// The reference scans every primitive with the full watertight test.  Here the scan has three phases that give the
// SAME hit record, bit for bit:
//   1. FILTER, converged over all lanes: every filter entry (ssb_blob.hpp: a planar quad, or one triangle of a
//      non-planar quad) is tested conservatively — ray/plane point against the entry's bounding rectangle in the
//      plane's own axes, enlarged by 1e-4 of the scene extent, and the side of the quad's diagonal, again with that
//      margin, which tells which of the two triangles can be hit.  Rays (nearly) parallel to the plane and degenerate
//      entries always pass.  Two entries are evaluated per packed-fp32 instruction (FFMA2 / FMUL2 / FADD2 of sm_100);
//      the result is two bit masks per lane: entries whose tri0 / tri1 may be hit.  The filter arithmetic (fma,
//      approximate reciprocal) is not the reference's and never touches the hit record; it can only reject triangles
//      the exact test would reject.
//   2. NEAREST FIRST (when no quad has both triangles as candidates, i.e. almost always): with at most one candidate
//      triangle per quad the reference's result is simply the candidate with the smallest (distance, list position)
//      among those the exact test accepts — PrimQuad's "skip tri1 if tri0 was hit" rule cannot fire.  That minimum
//      does not depend on the order of evaluation, so the candidate with the nearest ray/plane distance is tested
//      first, and the others are tested only if their plane distance, less the margin, does not already exceed the
//      hit distance found (entries the ray grazes, |n.d| < 0.05, are never skipped: there the plane distance and the
//      watertight distance are both ill-conditioned).  Typically ONE exact test per ray instead of one per plane the
//      ray pierces inside a rectangle.
//   3. IN LIST ORDER (a quad with both triangles as candidates — near its diagonal, degenerate, or a scene with more
//      than 32 entries): the reference's own sequence, tri0 before tri1, quads in list order, running strict minimum.
#pragma once

#include <cstdint>

#include "ssb_blob.hpp"

#if defined(__CUDACC__)
#define SSB_ISECT_FN __device__ __forceinline__
#ifndef SSB_INLINE_ISECT
#define SSB_INLINE_ISECT 1  // scene_intersect inlined into its (single) call site per kernel: the Hit record it fills through a
                            // reference then lives in registers instead of local memory (+2.6 %, profiles/r5u_ab.txt)
#endif
#if SSB_INLINE_ISECT
#define SSB_ISECT_NOINLINE __device__ __forceinline__
#else
#define SSB_ISECT_NOINLINE __device__ __noinline__
#endif
#else
// ---- host build (logic check only): the few device intrinsics used below, with the same semantics
#include <cmath>
#include <cstring>
#define SSB_ISECT_FN inline
#define SSB_ISECT_NOINLINE inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = { x, y }; return r; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(uint32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float ssb_canonical_nan(float v) { return v != v ? __uint_as_float(0x7fffffffu) : v; }  // PTX arithmetic returns the canonical NaN
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(ssb_canonical_nan(a.x + b.x), ssb_canonical_nan(a.y + b.y)); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
#endif

namespace ssbk {

// ------------------------------------------------------------------ shared-memory view of the blob
// The blob lives in dynamic shared memory.  It is declared at namespace scope and reached through these accessors
// (not through generic pointers carried in a struct) so that every load — also inside the non-inlined helpers —
// is a shared-memory load (LDS) rather than a generic one (ncu on the first wavefront build: LD.E.128 + R2UR
// in the intersection scan).
#if defined(__CUDACC__)
extern __shared__ __align__(128) unsigned char ssb_smem[];
#else
static const unsigned char* ssb_smem = nullptr;  // host check: points at a blob image
#endif
struct SceneView {
	SSB_ISECT_FN const DevHeader* hdr() const { return reinterpret_cast<const DevHeader*>(ssb_smem); }
	SSB_ISECT_FN const ssb_quad* quads() const { return reinterpret_cast<const ssb_quad*>(ssb_smem + hdr()->off_quads); }
	SSB_ISECT_FN const DevMaterial* materials() const { return reinterpret_cast<const DevMaterial*>(ssb_smem + hdr()->off_materials); }
	SSB_ISECT_FN const uint32_t* lights() const { return reinterpret_cast<const uint32_t*>(ssb_smem + hdr()->off_lights); }
	SSB_ISECT_FN const DevTexture* textures() const { return reinterpret_cast<const DevTexture*>(ssb_smem + hdr()->off_textures); }
	SSB_ISECT_FN const float* pool() const { return reinterpret_cast<const float*>(ssb_smem + hdr()->off_pool); }
	SSB_ISECT_FN const float4* fpairs() const { return reinterpret_cast<const float4*>(ssb_smem + hdr()->off_fpairs); }
	SSB_ISECT_FN const float4* planes() const { return reinterpret_cast<const float4*>(ssb_smem + hdr()->off_planes); }
	SSB_ISECT_FN const uint32_t* entry_quad() const { return reinterpret_cast<const uint32_t*>(ssb_smem + hdr()->off_entry_quad); }
	SSB_ISECT_FN const uint32_t* quad_mask() const { return reinterpret_cast<const uint32_t*>(ssb_smem + hdr()->off_quad_mask); }
	SSB_ISECT_FN const uint4* chunks() const { return reinterpret_cast<const uint4*>(ssb_smem + hdr()->off_chunks); }
};

// ------------------------------------------------------------------ small helpers
SSB_ISECT_FN float rcp_approx(float x) {  // 1 ulp MUFU.RCP: culling only, never reference arithmetic
#if defined(__CUDA_ARCH__)
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
#else
	return 1.0f / x;
#endif
}
SSB_ISECT_FN float max3_nan_ignoring(float a, float b, float c) {  // FMNMX3: NaN operands are ignored unless all are NaN
#if defined(__CUDA_ARCH__)
	return fmaxf(fmaxf(a, b), c);
#else
	return std::fmax(std::fmax(a, b), c);
#endif
}
SSB_ISECT_FN float max_nan_ignoring(float a, float b) {
#if defined(__CUDA_ARCH__)
	return fmaxf(a, b);
#else
	return std::fmax(a, b);
#endif
}
SSB_ISECT_FN float sel3(float x, float y, float z, int i) { return i == 0 ? x : (i == 1 ? y : z); }

// ------------------------------------------------------------------ exact test
struct RayConst {  // per-ray constants of the watertight test (geometry.cpp:17-37), hoisted out of the scan
	int kx, ky, kz;
	float Sx, Sy, Sz;
	float okx, oky, okz;  // ray origin permuted
};
#ifndef SSB_RAY_SETUP_ROTATE
#define SSB_RAY_SETUP_ROTATE 1
#endif
SSB_ISECT_FN RayConst ray_setup(float ox, float oy, float oz, float dx, float dy, float dz) {
	RayConst rc;
	float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
	int kx, ky, kz;
	if (ax > ay) {
		if (ax > az) { kz = 0; kx = 1; ky = 2; } else { kz = 2; kx = 0; ky = 1; }
	} else {
		if (ay > az) { kz = 1; kx = 2; ky = 0; } else { kz = 2; kx = 0; ky = 1; }
	}
#if SSB_RAY_SETUP_ROTATE
	// (kx, ky, kz) is a rotation of (0, 1, 2) chosen by kz, then kx <-> ky if the dominant component is negative: rotate the
	// two triples once (two shared predicates) and swap, instead of seven independent three-way selections
	const bool z0 = kz == 0, z1 = kz == 1;
	float dkx = z0 ? dy : (z1 ? dz : dx), dky = z0 ? dz : (z1 ? dx : dy);
	const float dkz = z0 ? dx : (z1 ? dy : dz);
	float okx = z0 ? oy : (z1 ? oz : ox), oky = z0 ? oz : (z1 ? ox : oy);
	rc.okz = z0 ? ox : (z1 ? oy : oz);
	if (dkz < 0.0f) { int t = kx; kx = ky; ky = t; float f = dkx; dkx = dky; dky = f; f = okx; okx = oky; oky = f; }
	rc.kx = kx; rc.ky = ky; rc.kz = kz;
	rc.Sx = dkx / dkz;
	rc.Sy = dky / dkz;
	rc.Sz = 1.0f / dkz;
	rc.okx = okx; rc.oky = oky;
	return rc;
#else
	float dkz = sel3(dx, dy, dz, kz);
	if (dkz < 0.0f) { int t = kx; kx = ky; ky = t; }
	rc.kx = kx; rc.ky = ky; rc.kz = kz;
	rc.Sx = sel3(dx, dy, dz, kx) / dkz;
	rc.Sy = sel3(dx, dy, dz, ky) / dkz;
	rc.Sz = 1.0f / dkz;
	rc.okx = sel3(ox, oy, oz, kx); rc.oky = sel3(ox, oy, oz, ky); rc.okz = sel3(ox, oy, oz, kz);
	return rc;
#endif
}

struct Hit {
	int quad;  // -1: none
	int tri;
	float dist;
	float bx, by, bz;  // barycentrics (UVW * det_recip)
};

// PrimTri::intersect (geometry.cpp:12-101); true when the hit record was updated.  The reference accepts
// EPS <= dist < hitrec->dist (geometry.cpp:88); `allow_equal` additionally accepts dist == hit.dist — used only by the
// nearest-first order, for a candidate that precedes the current hit in the list (the reference would have met it first).
SSB_ISECT_FN bool tri_intersect(const ssb_tri& t, const RayConst& rc, float eps, Hit& hit, bool allow_equal) {
	// vertices relative to the ray origin, permuted: A = v0 - orig etc. (geometry.cpp:40-47)
	float Akx = t.v[0].pos[rc.kx] - rc.okx, Aky = t.v[0].pos[rc.ky] - rc.oky, Akz = t.v[0].pos[rc.kz] - rc.okz;
	float Bkx = t.v[1].pos[rc.kx] - rc.okx, Bky = t.v[1].pos[rc.ky] - rc.oky, Bkz = t.v[1].pos[rc.kz] - rc.okz;
	float Ckx = t.v[2].pos[rc.kx] - rc.okx, Cky = t.v[2].pos[rc.ky] - rc.oky, Ckz = t.v[2].pos[rc.kz] - rc.okz;
	float Ax = Akx - rc.Sx * Akz, Bx = Bkx - rc.Sx * Bkz, Cx = Ckx - rc.Sx * Ckz;
	float Ay = Aky - rc.Sy * Akz, By = Bky - rc.Sy * Bkz, Cy = Cky - rc.Sy * Ckz;
	// UVW = cross(ABCy, ABCx)
	float U = By * Cx - Bx * Cy;
	float V = Cy * Ax - Cx * Ay;
	float W = Ay * Bx - Ax * By;
	if (U != 0.0f && V != 0.0f && W != 0.0f) {
		if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
	} else {
		double Ud = (double)By * (double)Cx - (double)Bx * (double)Cy;
		double Vd = (double)Cy * (double)Ax - (double)Cx * (double)Ay;
		double Wd = (double)Ay * (double)Bx - (double)Ax * (double)By;
		if ((Ud < 0.0 || Vd < 0.0 || Wd < 0.0) && (Ud > 0.0 || Vd > 0.0 || Wd > 0.0)) return false;
		U = (float)Ud; V = (float)Vd; W = (float)Wd;
	}
	float det = (U + V) + W;
	if (!(fabsf(det) > eps)) return false;
	float T = (U * (rc.Sz * Akz) + V * (rc.Sz * Bkz)) + W * (rc.Sz * Ckz);
	if (((__float_as_uint(det) ^ __float_as_uint(T)) & 0x80000000u) != 0u) return false;
	float det_recip = 1.0f / det;
	float dist = T * det_recip;
	if (dist >= eps && (dist < hit.dist || (allow_equal && dist == hit.dist))) {
		hit.dist = dist;
		hit.bx = U * det_recip; hit.by = V * det_recip; hit.bz = W * det_recip;
		return true;
	}
	return false;
}

// ------------------------------------------------------------------ phase 1: the filter over one chunk (<= 16 pair records)
// Per entry: reject tri0 if  max(|u|, |v|, -tp/margin, -sd, z) > 1,  tri1 if  max(|u|, |v|, -tp/margin, +sd, z) > 1, where
// (u,v) are the ray/plane point's coordinates scaled to the enlarged rectangle, tp the ray/plane distance, sd the signed
// distance to the diagonal in units of the margin, and z = (|h| - |n.d| * tmax) / margin with h the height of the ray
// origin over the plane and tmax a bound on the distance from the origin to any point of the scene: z > 1 means the
// plane cannot be reached inside the scene at all.
// Conditioning: the ray/plane point is only trusted when |n.d| >= SSB_PAR (its error, ~4e-7 of the scene extent divided
// by |n.d|, is then a fraction of the margin).  Below that the reciprocal is replaced by NaN: u, v, tp, sd become NaN,
// the maxima ignore NaN operands, and z alone decides — a ray that runs (nearly) parallel to a plane keeps the entry
// (both triangles) exactly when its origin is close enough to that plane to reach it.  1 - NaN = NaN has a clear sign
// bit, so an all-NaN entry is kept too.  The sign bits of (1 - max) are shifted into the masks entry by entry, last
// entry first, so that entry j lands in bit j.
#define SSB_PAR 0.016f
SSB_ISECT_FN void filter_chunk(const float4* rec, int npairs, float margin_rneg, float tmax_k, float ox, float oy, float oz, float dx, float dy, float dz,
                               unsigned& keepA, unsigned& keepB) {
	unsigned rejA = 0u, rejB = 0u;
	const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
	const float2 ox2 = make_float2(ox, ox), oy2 = make_float2(oy, oy), oz2 = make_float2(oz, oz);
	const float2 one2 = make_float2(1.0f, 1.0f), mr2 = make_float2(margin_rneg, margin_rneg), tk2 = make_float2(tmax_k, tmax_k);
	const float nanv = __int_as_float(0x7fffffff);
#ifndef SSB_FILTER_UNROLL
#define SSB_FILTER_UNROLL 1
#endif
	constexpr int kUnroll = SSB_FILTER_UNROLL;
#pragma unroll kUnroll
	for (int i = npairs - 1; i >= 0; --i) {
		const float4* r = rec + 8 * i;
		const float4 a0 = r[0], a1 = r[1], a2 = r[2], a3 = r[3], a4 = r[4], a5 = r[5], a6 = r[6], a7 = r[7];
		const float2 plx = make_float2(a0.x, a0.y), ply = make_float2(a0.z, a0.w), plz = make_float2(a1.x, a1.y), plw = make_float2(a1.z, a1.w);
		const float2 uax = make_float2(a2.x, a2.y), uay = make_float2(a2.z, a2.w), uaz = make_float2(a3.x, a3.y), uaw = make_float2(a3.z, a3.w);
		const float2 vbx = make_float2(a4.x, a4.y), vby = make_float2(a4.z, a4.w), vbz = make_float2(a5.x, a5.y), vbw = make_float2(a5.z, a5.w);
		const float2 dgx = make_float2(a6.x, a6.y), dgy = make_float2(a6.z, a6.w), dgz = make_float2(a7.x, a7.y);
		const float2 nd = __ffma2_rn(plx, dx2, __ffma2_rn(ply, dy2, __fmul2_rn(plz, dz2)));
		const float2 no = __ffma2_rn(plx, ox2, __ffma2_rn(ply, oy2, __fmul2_rn(plz, oz2)));
		const float2 num = __fadd2_rn(plw, make_float2(-no.x, -no.y));
		float2 ri;
		ri.x = (fabsf(nd.x) >= SSB_PAR) ? rcp_approx(nd.x) : nanv;  // ill-conditioned (or all-zero entry): z decides
		ri.y = (fabsf(nd.y) >= SSB_PAR) ? rcp_approx(nd.y) : nanv;
		const float2 hk = __fmul2_rn(num, mr2), ndk = __fmul2_rn(nd, tk2);  // -h / margin, n.d * tmax / margin
		const float2 tp = __fmul2_rn(num, ri);
		const float2 px = __ffma2_rn(tp, dx2, ox2), py = __ffma2_rn(tp, dy2, oy2), pz = __ffma2_rn(tp, dz2, oz2);
		const float2 u = __ffma2_rn(uax, px, __ffma2_rn(uay, py, __ffma2_rn(uaz, pz, uaw)));
		const float2 v = __ffma2_rn(vbx, px, __ffma2_rn(vby, py, __ffma2_rn(vbz, pz, vbw)));
		const float2 sd = __ffma2_rn(dgx, u, __ffma2_rn(dgy, v, dgz));
		const float2 q3 = __fmul2_rn(tp, mr2);  // -tp / margin
		{
			const float g = max3_nan_ignoring(fabsf(u.y), fabsf(v.y), q3.y), z = fabsf(hk.y) - fabsf(ndk.y);
			const float2 w = __fadd2_rn(one2, make_float2(-max3_nan_ignoring(g, -sd.y, z), -max3_nan_ignoring(g, sd.y, z)));
			rejA = __funnelshift_l(__float_as_uint(w.x), rejA, 1);
			rejB = __funnelshift_l(__float_as_uint(w.y), rejB, 1);
		}
		{
			const float g = max3_nan_ignoring(fabsf(u.x), fabsf(v.x), q3.x), z = fabsf(hk.x) - fabsf(ndk.x);
			const float2 w = __fadd2_rn(one2, make_float2(-max3_nan_ignoring(g, -sd.x, z), -max3_nan_ignoring(g, sd.x, z)));
			rejA = __funnelshift_l(__float_as_uint(w.x), rejA, 1);
			rejB = __funnelshift_l(__float_as_uint(w.y), rejB, 1);
		}
	}
	keepA = ~rejA; keepB = ~rejB;
}

#ifndef SSB_NEAREST_FIRST
#define SSB_NEAREST_FIRST 1
#endif
#if defined(SSB_ISECT_STATS) && !defined(__CUDACC__)  // host check only: how often each phase runs
struct IsectStats { unsigned long long queries, fast, fast_more, inorder, exact_tests, candidates; };
static IsectStats g_isect_stats = { 0, 0, 0, 0, 0, 0 };
#define SSB_STAT(field, n) (g_isect_stats.field += (n))
#else
#define SSB_STAT(field, n) ((void)0)
#endif
#define SSB_GRAZE 0.05f  // |n.d| below which a candidate's plane distance is not used to skip its exact test

// approximate ray/plane distance of an entry, -inf when the ray grazes the plane (never skipped, tested first)
SSB_ISECT_FN float entry_plane_key(const float4 pl, float ox, float oy, float oz, float dx, float dy, float dz) {
	const float nd = __fmaf_rn(pl.x, dx, __fmaf_rn(pl.y, dy, pl.z * dz));
	const float no = __fmaf_rn(pl.x, ox, __fmaf_rn(pl.y, oy, pl.z * oz));
	const float tp = (pl.w - no) * rcp_approx(nd);
	return (fabsf(nd) >= SSB_GRAZE) ? tp : -__int_as_float(0x7f800000);
}

SSB_ISECT_NOINLINE void scene_intersect(const SceneView& S, float eps, int ignore, Hit& hit,
                                        float ox, float oy, float oz, float dx, float dy, float dz) {
	hit.quad = -1; hit.tri = 0; hit.dist = __int_as_float(0x7f800000);
	hit.bx = hit.by = hit.bz = 0.0f;
	const DevHeader* H = S.hdr();
	const int nent = (int)H->nentries;
	const float margin = H->cull_margin, margin_rneg = H->cull_margin_rneg;
	// bound on the distance from the ray origin to any point of the scene, in units of the margin (L1 distance to the
	// centre of the bounding box + its half diagonal: >= the Euclidean bound)
	const float tmax_k = ((fabsf(ox - H->scene_centre[0]) + fabsf(oy - H->scene_centre[1])) + (fabsf(oz - H->scene_centre[2]) + H->scene_radius)) * H->tmax_scale;
	const uint32_t* entry_quad = S.entry_quad();
	const RayConst rc = ray_setup(ox, oy, oz, dx, dy, dz);
	for (int base = 0; base < nent; base += 32) {
		const int cnt = (nent - base < 32) ? (nent - base) : 32;
		unsigned candA, candB;
		filter_chunk(S.fpairs() + 8 * (base >> 1), cnt >> 1, margin_rneg, tmax_k, ox, oy, oz, dx, dy, dz, candA, candB);
		const uint4 cm = S.chunks()[base >> 5];
		candA &= cm.x; candB &= cm.y;
#if SSB_NEAREST_FIRST
		if (nent <= 32) {
			if (ignore >= 0) { const unsigned im = ~S.quad_mask()[ignore]; candA &= im; candB &= im; }
			// a quad with both triangles as candidates: same entry, or the two entries of a split quad
			const unsigned both = (candA & candB) | (candA & cm.z & (candB >> 1));
			if (both == 0u) {
				// ---- phase 2: nearest first
				unsigned cand = candA | candB;
				SSB_STAT(queries, 1); SSB_STAT(fast, 1); SSB_STAT(candidates, __builtin_popcount(cand));
				if (cand == 0u) return;
				const float inf = __int_as_float(0x7f800000);
				float t1 = inf, t2 = inf;
				int e1 = __ffs(cand) - 1, e2 = -1;
				// (a single candidate needs no ordering: most shadow rays only have the light's entry)
				for (unsigned m = (cand & (cand - 1u)) ? cand : 0u; m != 0u; m &= m - 1u) {
					const int e = __ffs(m) - 1;
					const float key = entry_plane_key(S.planes()[e], ox, oy, oz, dx, dy, dz);
					const bool lt1 = key < t1, lt2 = key < t2;
					e2 = lt1 ? e1 : (lt2 ? e : e2);
					t2 = lt1 ? t1 : (lt2 ? key : t2);
					e1 = lt1 ? e : e1;
					t1 = lt1 ? key : t1;
				}
				// exact tests: the nearest entry e1; then, unless the second-nearest plane distance t2 already rules everything
				// else out, the second-nearest e2; then (rarely) the rest in list order, each unless its own plane distance rules it out
				int best_e = -1;  // entry of the current hit (-1: none, so that `e < best_e` never allows an equal distance)
				int e = e1;
				unsigned rest = cand & ~(1u << e1);
				int step = 0;
				for (;;) {
					const int q = (int)entry_quad[e], tt = (int)((candB >> e) & 1u);
					SSB_STAT(exact_tests, 1); if (step) SSB_STAT(fast_more, 1);
					if (tri_intersect(S.quads()[q].tri[tt], rc, eps, hit, e < best_e)) { hit.quad = q; hit.tri = tt; best_e = e; }
					if (rest == 0u) break;
					if (step < 2) {
						if (t2 - margin > hit.dist) break;  // every remaining entry has a plane distance >= t2
						if (step == 0 && e2 >= 0 && ((rest >> e2) & 1u)) { step = 1; e = e2; rest &= ~(1u << e2); continue; }
						step = 2;
					}
					bool found = false;
					while (rest != 0u) {
						e = __ffs(rest) - 1;
						rest &= rest - 1u;
						const float key = entry_plane_key(S.planes()[e], ox, oy, oz, dx, dy, dz);
						if (!(key - margin > hit.dist)) { found = true; break; }
					}
					if (!found) break;
				}
				return;
			}
		}
#endif
		// ---- phase 3: the reference's own order (one exact test per iteration keeps lanes with different candidates together)
		// (Entries that keep BOTH triangles are mostly planes the ray runs (nearly) parallel to with the origin close to the
		// plane: every shadow ray from the Cornell box's ceiling towards the coplanar light keeps the four other ceiling
		// pieces and the light this way.  They cannot be thinned out by where the ray's projection runs in the plane: for a ray
		// IN a triangle's plane the watertight test's U, V, W are all rounding noise around zero, and the reference accepts
		// such a "hit" wherever the noise happens to agree in sign — also far beside the triangle.  An in-plane pre-reject
		// tried in round 2 was caught by the device fuzz test on exactly such a ray; tests/test_gpu_isect_fuzz.py.)
		SSB_STAT(queries, base == 0 ? 1 : 0); SSB_STAT(inorder, base == 0 ? 1 : 0); SSB_STAT(candidates, __builtin_popcount(candA) + __builtin_popcount(candB));
		while (candA | candB) {
			const unsigned any = candA | candB;
			const unsigned bit = any & (0u - any);
			const int e = __ffs(bit) - 1;
			const int q = (int)entry_quad[base + e];
			const int tt = (candA & bit) ? 0 : 1;
			candA &= ~bit;
			if (tt == 1) candB &= ~bit;
			// tri1 is skipped when tri0 of the same quad was hit (PrimQuad::intersect, geometry.cpp:131-133): hit.quad == q can
			// only stem from this scan's tri0, a quad being visited once
			if (q == ignore || (tt == 1 && hit.quad == q)) continue;
			SSB_STAT(exact_tests, 1);
			if (tri_intersect(S.quads()[q].tri[tt], rc, eps, hit, false)) { hit.quad = q; hit.tri = tt; candB &= ~bit; }
		}
	}
}

// The reference's scan, verbatim (no filter): the yardstick of tools/isect_check.cpp
SSB_ISECT_FN void scene_intersect_listscan(const SceneView& S, float eps, int ignore, Hit& hit,
                                           float ox, float oy, float oz, float dx, float dy, float dz) {
	const RayConst rc = ray_setup(ox, oy, oz, dx, dy, dz);
	hit.quad = -1; hit.tri = 0; hit.dist = __int_as_float(0x7f800000);
	hit.bx = hit.by = hit.bz = 0.0f;
	const int nq = (int)S.hdr()->nquads;
	for (int q = 0; q < nq; ++q) {
		if (q == ignore) continue;
		const ssb_quad& quad = S.quads()[q];
		if (tri_intersect(quad.tri[0], rc, eps, hit, false)) { hit.quad = q; hit.tri = 0; }
		else if (tri_intersect(quad.tri[1], rc, eps, hit, false)) { hit.quad = q; hit.tri = 1; }
	}
}

#if defined(__CUDACC__)
SSB_ISECT_NOINLINE void scene_intersect_listscan_noinline(const SceneView& S, float eps, int ignore, Hit& hit,
                                                          float ox, float oy, float oz, float dx, float dy, float dz) {
	scene_intersect_listscan(S, eps, ignore, hit, ox, oy, oz, dx, dy, dz);
}
#endif

}  // namespace ssbk
