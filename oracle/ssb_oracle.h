/* ssb_oracle.h — TEST INFRASTRUCTURE ONLY (see ssb_oracle.c). */
#ifndef SSB_ORACLE_H
#define SSB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/ssb200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ssb_oracle_counters { /* path statistics, all uint64 */
	uint64_t samples, closest_queries, shadow_queries, unshadowed, bsdf_samples, texture_lookups, tri_tests, double_fallbacks;
} ssb_oracle_counters;
/* accum: width*height*4 doubles, ADDED to (raw sum of sample*0.001f, renderer.cpp:294); may be NULL.
 * samples_out: width*height*(sample_end-sample_begin)*4 floats or NULL. */
int ssb_oracle_render(const ssb_scene* scene, const ssb_color* color, const ssb_options* opt,
                      double* accum, float* samples_out, ssb_oracle_counters* counters);
int ssb_oracle_resolve(const ssb_color* color, const ssb_options* opt, const double* accum, double* xyza, float* srgba);
int ssb_oracle_intersect(const ssb_scene* scene, const float* rays6, const int32_t* ignore, float eps, float* out6, size_t n);
void ssb_oracle_eval_math(uint32_t fn, const float* x, float arg, float* out, size_t n);
#ifdef __cplusplus
}
#endif
#endif
