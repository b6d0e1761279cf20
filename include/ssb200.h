/* ssb200.h — C ABI of the B200-native spectral path-tracing integrator (libssb200.so).
 *
 * The reference (geometrian/simple-spectral) has NO plugin / FFI boundary: it is one C++17
 * binary whose hot path is Renderer::_render_pixel / _render_sample (src/renderer.cpp:103-308),
 * driven by Renderer::render_start / _render_threadwork (src/renderer.cpp:309-422).  This header
 * is the boundary a maintainer would bind instead of that per-pixel loop:
 * everything is plain-old-data, caller-owned HOST pointers that are copied at upload time,
 * `int` status codes and no C++/torch types.  INTEGRATION.md shows the reference-side stub.
 *
 * Status codes follow the reference's `throw int` convention:
 *    0  ok
 *   -1  data / IO / CUDA runtime error        (reference: spectrum.cpp:19,181; material.cpp:17)
 *   -2  invalid argument                      (reference: spectrum.cpp:198; main.cpp:78)
 *   -3  unsupported / unknown configuration   (reference: renderer.cpp:37; main.cpp:100)
 * ssb_last_error() returns a thread-local, human-readable message for the last failure.
 *
 * Threading: one context per GPU.  Calls on one context must be serialised by the caller;
 * different contexts may be used from different threads.
 */
#ifndef SSB200_H
#define SSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 4u /* 2: RGB render mode (ssb_material.*_rgb, ssb_options.render_mode), n_wavelengths;
                           * 3: ssb_options.prebaked_textures (was `reserved`);
                           * 4: ssb_options.band_* / scan_mode (appended), ssb_device_count, ssb_accum_merge,
                           *    ssb_debug_intersect */

#define SSB_OK 0
#define SSB_ERR_DATA (-1)
#define SSB_ERR_ARG (-2)
#define SSB_ERR_UNSUPPORTED (-3)

/* limits of the shared-memory-resident scene (cornell*: 19 quads, plane-srgb: 7) */
#define SSB_MAX_QUADS 256u
#define SSB_MAX_LIGHTS 64u
#define SSB_MAX_MATERIALS 64u
#define SSB_MAX_DEPTH 16u

typedef struct ssb_ctx ssb_ctx; /* opaque; owns device memory, a stream and the uploaded tables */

/* ---- geometry: replaces Vertex / PrimTri / PrimQuad (src/geometry.hpp:13-21,54-104) ---- */
typedef struct ssb_vertex {
	float pos[3];
	float st[2];
} ssb_vertex;

typedef struct ssb_tri {
	ssb_vertex v[3];
	float normal[3]; /* normalize(cross(v1-v0, v2-v0)), geometry.hpp:68 — supplied by the host */
} ssb_tri;

/* PrimQuad = tri0 (v00,v10,v11) then tri1 (v00,v11,v01), geometry.hpp:93-95.  List order is
 * part of the contract: it defines tie-breaks (strict `dist<best`, geometry.cpp:88) and the
 * identity used by `ignore` (scene.cpp:433-445). */
typedef struct ssb_quad {
	ssb_tri tri[2];
	uint32_t material; /* index into ssb_scene.materials */
	uint32_t is_light; /* MaterialBase::is_emissive(), material.cpp:100-106 */
} ssb_quad;

/* ---- spectra: replaces _Spectrum (src/spectrum.hpp:12-70) ---- */
#define SSB_FILTER_LINEAR 0u  /* _Spectrum::_sample_linear, spectrum.cpp:39-60 (the default) */
#define SSB_FILTER_NEAREST 1u /* _Spectrum::_sample_nearest, spectrum.cpp:29-38 */
typedef struct ssb_spectrum {
	const float* data; /* n samples, first at `low`, last at `high`, evenly spaced */
	uint32_t n;        /* >= 2 (spectrum.cpp:17-20) */
	float low, high;   /* nm */
	uint32_t filter;
} ssb_spectrum;

/* ---- materials: replaces MaterialLambertian / MaterialMirror (src/material.hpp:48-178) ---- */
#define SSB_MATERIAL_LAMBERT 0u
#define SSB_MATERIAL_MIRROR 1u
#define SSB_ALBEDO_CONSTANT 0u
#define SSB_ALBEDO_TEXTURE 1u
typedef struct ssb_material {
	uint32_t kind;
	uint32_t albedo_mode;
	ssb_spectrum albedo;   /* used when albedo_mode == CONSTANT */
	uint32_t texture;      /* index into ssb_scene.textures when albedo_mode == TEXTURE */
	ssb_spectrum emission; /* MaterialBase::emission (default: constant 0) */
	/* RENDER_MODE_RGB only (the `#else` branches of material.hpp:51-55,120-126): l-RGB triples used instead of
	 * the two spectra when ssb_options.render_mode == SSB_RENDER_RGB; ignored in spectral mode */
	float albedo_rgb[3];   /* RGB_Reflectance, default (1,1,1) */
	float emission_rgb[3]; /* RGB_Radiance, default (0,0,0) */
} ssb_material;

/* sRGB_ReflectanceTexture (material.hpp:14-44): RGB8, scanlines top-to-bottom */
typedef struct ssb_texture {
	const uint8_t* rgb8;
	uint32_t width, height;
} ssb_texture;

/* Scene::camera (scene.hpp:16-33).  pv_inv is column-major like glm::dmat4x4. */
typedef struct ssb_camera {
	double pv_inv[16];
	float pos[3];
	float dir[3];
} ssb_camera;

typedef struct ssb_scene {
	ssb_camera camera;
	const ssb_quad* quads;
	uint32_t nquads;
	const ssb_material* materials;
	uint32_t nmaterials;
	const ssb_texture* textures;
	uint32_t ntextures;
} ssb_scene;

/* ---- colour tables: replaces Color::data (src/util/color.hpp:22-69) ---- */
#define SSB_UPSAMPLE_OURS 1u /* RENDER_MODE_SPECTRAL_ALGNUM 1 (stdafx.hpp:66): basis, color.cpp:166-173 */
#define SSB_UPSAMPLE_MENG 2u /* Meng et al. 2015, color.cpp:174-201 */
#define SSB_UPSAMPLE_JH 3u   /* Jakob & Hanika 2019, color.cpp:202-232 */

typedef struct ssb_meng_tables { /* src/meng-et-al.-2015/spectra_xyz_5nm_380_780_0.97.h */
	const int32_t* grid;     /* ncells * 8: inside, num_points, idx[6] */
	uint32_t grid_w, grid_h; /* 12 x 14 */
	const float* points;     /* npoints * (3 + 2 + nsamples): xyz[3], uv[2], spectrum[] */
	uint32_t npoints, nsamples;
	float xy_to_uv[6];
	float sample_min, sample_max;
} ssb_meng_tables;

typedef struct ssb_color {
	ssb_spectrum xbar, ybar, zbar;          /* CIE standard observer (color.cpp:77-99) */
	ssb_spectrum basis_r, basis_g, basis_b; /* OURS only (color.cpp:122-141) */
	float xyz_to_lrgb[9];                   /* column-major glm::mat3x3 (color.cpp:146-154) */
	float d65_rad_Y;                        /* D65_rad_XYZ.y, MENG tonemap only (color.cpp:247) */
	const float* jh_scale;                  /* JH only: res floats (rgb2spec.c:36-43) */
	const float* jh_data;                   /* JH only: 3*res^3*3 floats */
	uint32_t jh_res;
	const ssb_meng_tables* meng;            /* MENG only */
} ssb_color;

/* RENDER_MODE_SPECTRAL vs RENDER_MODE_RGB (stdafx.hpp:62-90) */
#define SSB_RENDER_SPECTRAL 0u /* hero-wavelength spectral transport, CIE XYZ accumulator (the default) */
#define SSB_RENDER_RGB 1u      /* three-channel l-RGB transport: no wavelength sample, no upsampling, no colour tables;
                                * the accumulator holds sum(l-RGB, hit) and resolve is avg/spp -> lrgb_to_srgb
                                * (renderer.cpp:300-307) */

/* ---- render options: Renderer::Options (renderer.hpp:16-29) + the compile-time macros of
 * stdafx.hpp:44-90 exposed as runtime fields ---- */
typedef struct ssb_options {
	uint32_t width, height; /* image resolution */
	uint32_t spp;           /* samples per pixel of the whole job (averaging denominator) */
	/* work subset: pixels [x0,x1) x [y0,y1), samples [sample_begin, sample_end) of each pixel.
	 * 0/0 for x1/y1/sample_end means "all".  Row 0 is the BOTTOM row (framebuffer.hpp:24-26). */
	uint32_t x0, y0, x1, y1;
	uint32_t sample_begin, sample_end;
	uint32_t indirect_only;           /* --indirect-only (renderer.cpp:169,184) */
	uint32_t upsampling;              /* SSB_UPSAMPLE_* */
	float lambda_min, lambda_max;     /* LAMBDA_MIN/MAX: 380/780 (CIE 1931) or 390/830 (CIE 2006) */
	uint32_t max_depth;               /* MAX_DEPTH, 10 */
	uint32_t explicit_light_sampling; /* EXPLICIT_LIGHT_SAMPLING, 1 */
	uint32_t flat_field_correction;   /* FLAT_FIELD_CORRECTION, 1 */
	float eps;                        /* EPS, 1e-3f */
	uint64_t seed;                    /* per-sample seeding: PCG32.seed(mix(seed, sample index)) */
	uint32_t render_mode;             /* SSB_RENDER_* */
	uint32_t n_wavelengths;           /* SAMPLE_WAVELENGTHS (stdafx.hpp:90): 0 = the default 4; 2, 3 or 4 (the sizes of
	                                   * glm::vec the reference compiles with).  LAMBDA_STEP = (max-min)/n (stdafx.hpp:289) */
	uint32_t keep_accumulator;        /* 1: do not clear the accumulator although sample_begin == 0 (several pixel
	                                   * rectangles of one frame rendered by successive ssb_render calls) */
	uint32_t prebaked_textures;       /* JH upsampling only (ignored otherwise).  1: rgb2spec_fetch (rgb2spec.c:77-118) runs ONCE per
	                                   * texel into a 32-bit coefficient texture kept by the context, and shading only evaluates
	                                   * the polynomial — the pre-process Jakob & Hanika intend, which the reference describes but
	                                   * does not implement (color.cpp:204-216,222-223).  The coefficients are the same floats the
	                                   * per-lookup form computes, so results do not change by a bit; costs 16 B/texel of HBM. */
	/* Multi-GPU row bands (the reference's Framebuffer::Tile decomposition, framebuffer.hpp:14-21, renderer.cpp:399-406,
	 * dealt out statically): with band_count > 1 this call renders only the image rows j with
	 * (j / band_height) % band_count == band_index — interleaved bands of band_height rows, so that every GPU gets a
	 * share of the expensive and of the cheap rows.  Requires y0 == 0 and y1 == 0 or height.  Per-sample seeding makes
	 * every pixel's samples the ones a single-GPU render draws: the merged frame is bit-identical.  0/0/0 = off. */
	uint32_t band_height, band_count, band_index;
	uint32_t scan_mode;               /* SSB_SCAN_*: how Scene::intersect's linear scan is executed (same hit record either way) */
} ssb_options;
#define SSB_SCAN_FILTERED 0u /* conservative packed filter + exact tests on the candidates (the default) */
#define SSB_SCAN_LIST 1u     /* the reference's own loop: every quad, tri0 then tri1, exact test only (scene.cpp:433-445) —
                              * the yardstick the filtered scan is fuzzed against on the device */

typedef struct ssb_stats {
	uint64_t samples;      /* path samples traced by the last ssb_render */
	double device_ms;      /* CUDA-event time of the last ssb_render's kernels (render stream) */
	double trace_ms;       /* ... of the path-tracing kernel(s) alone */
	uint32_t launches;     /* kernel launches issued by the last ssb_render */
	uint32_t reserved;
} ssb_stats;

uint32_t ssb_abi_version(void);
const char* ssb_last_error(void);

/* fills *opt with the reference's defaults (stdafx.hpp:44-90) for a WxH image */
void ssb_default_options(ssb_options* opt, uint32_t width, uint32_t height, uint32_t spp);

int ssb_device_count(int* count); /* CUDA devices visible to this process (0 and SSB_ERR_DATA without a GPU) */
int ssb_create(int device, ssb_ctx** out);
void ssb_destroy(ssb_ctx* ctx);

/* Copy scene / colour tables to the device (replaces the pointer graph Scene -> PrimBase* ->
 * MaterialBase* -> _Spectrum that _render_sample walks, renderer.cpp:147-255). */
int ssb_upload_scene(ssb_ctx* ctx, const ssb_scene* scene);
int ssb_upload_color(ssb_ctx* ctx, const ssb_color* color);
/* Same as ssb_upload_scene, but the texel copies are only ENQUEUED (own copy stream): they overlap the camera-ray
 * stage of the next ssb_render, whose first shading stage waits for them.  The caller's ssb_texture.rgb8 buffers must
 * stay valid and unmodified until ssb_synchronize returns, or until a host-returning call (ssb_render_frame,
 * ssb_resolve, ssb_read_accum) that FOLLOWS an ssb_render of this scene returns; use pinned host memory for the copy
 * to be truly asynchronous.
 * Everything else (quads, materials, camera) is copied before the call returns, as in ssb_upload_scene. */
int ssb_upload_scene_async(ssb_ctx* ctx, const ssb_scene* scene);

/* Trace the requested samples and ADD each sample's float4 (X,Y,Z,hit)*0.001f, in sample order,
 * to the context's per-pixel double XYZA accumulator (renderer.cpp:292-295; RGB mode: the float4
 * (r,g,b,hit) unscaled, renderer.cpp:301-303).  The accumulator is cleared first when sample_begin == 0, unless
 * keep_accumulator is set, and by ssb_clear().  Device-resident; no host transfer. */
int ssb_render(ssb_ctx* ctx, const ssb_options* opt);
int ssb_clear(ssb_ctx* ctx);

/* Raw accumulator (sum of sample*0.001f): width*height*4 doubles.  `dst` is a HOST pointer for
 * ssb_read_accum, a DEVICE pointer for ssb_accum_device (borrowed, valid until the next
 * ssb_render with another resolution / ssb_destroy) — the latter is what a multi-GPU caller
 * hands to its collective. */
int ssb_read_accum(ssb_ctx* ctx, double* dst_host);
int ssb_write_accum(ssb_ctx* ctx, const double* src_host);
int ssb_accum_device(ssb_ctx* ctx, double** dptr, size_t* count);

/* Multi-GPU inside one process (the reference's Renderer uses every execution unit of the box by itself,
 * renderer.cpp:396-430): fold the accumulator of `src` into the accumulator of `dst` (same resolution; the contexts may
 * live on different GPUs).  `src_opt` = the options `src` rendered its share with:
 *   - a pixel subset (band_count > 1, or a pixel rectangle smaller than the image): those pixels are COPIED from src
 *     (disjoint tiles: a gather, bit-exact);
 *   - otherwise (a sample range of the whole frame): dst += src, element-wise in f64.
 * Ordered after everything already enqueued on either context's stream, asynchronous with respect to the host; src's
 * stream in turn waits until dst has consumed the data.  Uses direct peer access (NVLink) when the devices allow it,
 * else a peer copy into a staging buffer on dst's device. */
int ssb_accum_merge(ssb_ctx* dst, ssb_ctx* src, const ssb_options* src_opt);

/* Finish a frame: avg = accum * (1000/spp) (renderer.cpp:296), then
 * sRGBA = (ciexyz_to_srgb(float3(avg)), float(avg.a)) (renderer.cpp:298, color.cpp:237-257).
 * RGB mode: avg = accum / spp, sRGBA = (lrgb_to_srgb(float3(avg)), float(avg.a)) (renderer.cpp:304-306).
 * Either output may be NULL.  HOST pointers, width*height*4 elements, row 0 = bottom. */
int ssb_resolve(ssb_ctx* ctx, const ssb_options* opt, double* xyza_host, float* srgba_host);

/* Same, but the results stay on the device (borrowed pointers into context-owned buffers, valid until
 * the next resolve at another resolution / ssb_destroy); no host synchronisation. */
int ssb_resolve_device(ssb_ctx* ctx, const ssb_options* opt, double** xyza_dev, float** srgba_dev);

/* Issue all subsequent work of this context on the caller's CUDA stream (a cudaStream_t passed as
 * void*; NULL restores the context's own stream).  ssb_render / ssb_clear / ssb_resolve_device are then
 * asynchronous with respect to the host; entry points that return host data synchronise that stream. */
int ssb_set_stream(ssb_ctx* ctx, void* cuda_stream);

/* The whole seam in one call — what Renderer::render_start()+render_wait() do for a frame:
 * clear, render all samples, resolve, copy back. */
int ssb_render_frame(ssb_ctx* ctx, const ssb_options* opt, double* xyza_host, float* srgba_host);

int ssb_get_stats(ssb_ctx* ctx, ssb_stats* out);
int ssb_synchronize(ssb_ctx* ctx);

/* Test hooks (bit-exactness of the device maths against the host libm the reference links):
 * evaluates fn over n inputs ON THE GPU.  fn: 0 sinf, 1 cosf, 2 acosf, 3 powf(x, y=arg),
 * 4 / 5: the sin / cos result of the paired sincos the kernels use. */
int ssb_debug_eval_math(ssb_ctx* ctx, uint32_t fn, const float* x_host, float arg, float* out_host, size_t n);
/* Closest-hit queries of n caller-supplied rays against the uploaded scene, on the GPU, through the same
 * scene_intersect the render kernels call (scan_mode: SSB_SCAN_*).  rays: n x {origin[3], dir[3]}; ignore: n quad
 * indices or -1 (may be NULL = none); out: n x {quad (int bits, -1 = miss), tri (int bits), dist, bx, by, bz}.
 * Test hook of the conservative filter: tests/test_gpu_isect_fuzz.py compares SSB_SCAN_FILTERED with the oracle's scan. */
int ssb_debug_intersect(ssb_ctx* ctx, const float* rays6_host, const int32_t* ignore_host, uint32_t scan_mode, float eps,
                        float* out6_host, size_t n);
/* Per-sample outputs of one pixel (float4 per sample), for matched-seed debugging. */
int ssb_debug_trace_samples(ssb_ctx* ctx, const ssb_options* opt, uint32_t px, uint32_t py, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* SSB200_H */
