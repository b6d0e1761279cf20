/* ssb200_host.h — C wrappers over the C++ host layer (simple-spectral_b200/csrc/host): the same
 * Color::init / Scene::get_new_* / Renderer surface the reference exposes as C++ classes
 * (src/util/color.hpp:69-75, src/scene.hpp:49-59, src/renderer.hpp:13-82), flattened for FFI users
 * (ctypes in this repo's tests and bench).  Status codes and ssbh_last_error() as in ssb200.h. */
#ifndef SSB200_HOST_H
#define SSB200_HOST_H
#include "ssb200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssbh_color ssbh_color;       /* Color::data */
typedef struct ssbh_scene ssbh_scene;       /* Scene */
typedef struct ssbh_renderer ssbh_renderer; /* Renderer */

const char* ssbh_last_error(void);

/* Color::init() (color.cpp:72-155).  data_root contains the reference's "data/" directory.
 * observer: 1931 | 2006 (CIE_OBSERVER); upsampling: SSB_UPSAMPLE_* (RENDER_MODE_SPECTRAL_ALGNUM). */
int ssbh_color_init(const char* data_root, int observer, uint32_t upsampling, ssbh_color** out);
const ssb_color* ssbh_color_flat(const ssbh_color* color);
int ssbh_color_query(const ssbh_color* color, float* lambda_min_max2, float* d65_orig_xyz3, float* d65_rad_xyz3,
                     float* lrgb_to_xyz9, float* xyz_to_lrgb9);
int ssbh_color_spectrum(const ssbh_color* color, const char* name, ssb_spectrum* out); /* D65_orig, D65_rad, xbar, ... */
/* Color::round_trip_srgb (color.cpp:259-294, OURS tables only): sRGB -> spectrum -> XYZ under D65 -> sRGB */
int ssbh_color_round_trip_srgb(const ssbh_color* color, const float* srgb3, float* srgb_out3);
/* The reference's round-trip self-test (main.cpp:246-262) for the red levels [r_begin, r_end): running_max[k] = the
 * maximum |error| after level r_begin+k, starting from start_max.  Over 0..255 the reference documents 1.851469e-5. */
int ssbh_color_round_trip_running_max(const ssbh_color* color, uint32_t r_begin, uint32_t r_end, float start_max,
                                      float* running_max, uint32_t threads);
void ssbh_color_free(ssbh_color* color);

/* Scene::get_new_cornell / _cornell_srgb / _plane_srgb (scene.cpp:32-415); unknown name: -3 */
int ssbh_scene_new(const char* name, const char* data_root, const ssbh_color* color, int explicit_light_sampling, ssbh_scene** out);
const ssb_scene* ssbh_scene_flat(const ssbh_scene* scene);
int ssbh_scene_camera(const ssbh_scene* scene, double* matr_P16, double* matr_V16, double* matr_PV_inv16);
void ssbh_scene_free(ssbh_scene* scene);

/* texture decode (material.cpp:10-29) / image writers by extension (framebuffer.cpp:39-176) */
int ssbh_load_png_rgb8(const char* path, uint8_t** rgb8, uint32_t* width, uint32_t* height); /* free with ssbh_free */
void ssbh_free(void* p);
int ssbh_save_image(const char* path, const float* srgba, uint32_t width, uint32_t height);

/* Renderer (renderer.hpp:13-82): Options + the compile-time configuration as fields */
typedef struct ssbh_renderer_options {
	const char* scene_name;
	uint32_t width, height, spp;
	uint32_t indirect_only;
	const char* output_path; /* may be NULL: no file is written */
	int observer;
	uint32_t upsampling;
	uint32_t explicit_light_sampling, max_depth, flat_field_correction;
	uint64_t seed;
	int device;
	const char* data_root;
	uint32_t render_mode; /* SSB_RENDER_SPECTRAL | SSB_RENDER_RGB (RENDER_MODE_RGB, stdafx.hpp:62-90) */
	uint32_t n_wavelengths; /* SAMPLE_WAVELENGTHS (stdafx.hpp:90): 0 = 4; 2, 3 or 4 */
	uint32_t prebaked_textures; /* ssb_options.prebaked_textures (JH coefficient textures, color.cpp:204-216) */
	uint32_t progressive; /* 1: render in sample slices 1,1,2,4,... and refresh the framebuffer after each (the live
	                       * preview of the reference's window, main.cpp:316-323); the final image is unchanged */
	/* Multi-GPU (appended in ABI 4; the reference's Renderer uses every execution unit of the box by itself,
	 * renderer.cpp:396-430): ndevices > 0 renders on devices[0..ndevices) (one context + host thread each; `device` is
	 * then ignored), merged and resolved on devices[0].  shard: 0 = interleaved bands of band_height rows (0 = 8, the
	 * reference's tile edge; bit-identical to one GPU), 1 = sample ranges (differs by f64 summation order only). */
	const int* devices;
	uint32_t ndevices;
	uint32_t shard;
	uint32_t band_height;
} ssbh_renderer_options;
#define SSBH_SHARD_TILES 0u
#define SSBH_SHARD_SAMPLES 1u
int ssbh_renderer_new(const ssbh_renderer_options* options, ssbh_renderer** out);
int ssbh_renderer_render(ssbh_renderer* r); /* render_start() + render_wait() */
/* The reference's asynchronous life cycle (renderer.hpp:71-81; main.cpp:313-327): start returns at once, a caller polls
 * is_rendering / snapshot (what the reference's display loop does with framebuffer.draw()), stop asks the render to end
 * after the slice in flight, wait joins, reports the render's error and saves the image to output_path. */
int ssbh_renderer_start(ssbh_renderer* r);
void ssbh_renderer_stop(ssbh_renderer* r);
int ssbh_renderer_wait(ssbh_renderer* r);
int ssbh_renderer_is_rendering(const ssbh_renderer* r);
/* thread-safe copy of the current framebuffer (width*height*4 sRGBA; may be NULL) -> samples per pixel behind it */
uint32_t ssbh_renderer_snapshot(const ssbh_renderer* r, float* srgba);
const float* ssbh_renderer_framebuffer(const ssbh_renderer* r); /* width*height*4 sRGBA, row 0 = bottom */
const double* ssbh_renderer_xyza(const ssbh_renderer* r);
int ssbh_renderer_stats(const ssbh_renderer* r, ssb_stats* out);
void ssbh_renderer_free(ssbh_renderer* r);

#ifdef __cplusplus
}
#endif
#endif
