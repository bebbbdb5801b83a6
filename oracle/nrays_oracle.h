/*
 * nrays_oracle.h — C entry points of the CPU oracle (TEST INFRASTRUCTURE, not product).
 *
 * The oracle is a CPU restatement of the nrays render hot path (src/scene.rs:29-339 and the
 * ncollide3d 0.16 / nalgebra 0.15 behaviour at its call sites, SURVEY.md Appendix A+B).
 * PARITY UNPINNED: the reference has no tests, golden images or known-answer vectors, cannot
 * be compiled here (no rustc/cargo, deps un-vendored), and ncollide3d's source is absent — see
 * the header comment of nrays_oracle.cpp.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  It consumes the same flattened tables as the product (include/nrays_b200.h).
 */
#ifndef NRAYS_ORACLE_H
#define NRAYS_ORACLE_H

#include "../include/nrays_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct NroScene NroScene;

/* precision_bits: 64 = faithful (reference Scalar = f64), 32 = device-twin (all geometry in f32). */
int nro_scene_create(const NrbSceneDesc *desc, int precision_bits, NroScene **out);
void nro_scene_destroy(NroScene *s);

/* scene::render restated.  Renders pixels [first_pixel, first_pixel + n_pixels) of the row-major
 * image (whole image when n_pixels == 0) into out_rgb (always width*height*3 floats; untouched
 * pixels keep their value).  n_threads <= 0 -> hardware concurrency.  Threads own contiguous
 * ranges of npixels/T + 1 pixels exactly as src/scene.rs:58-63. */
int nro_render(NroScene *s, const NrbCamera *cam, int n_threads, uint64_t first_pixel, uint64_t n_pixels,
               float *out_rgb, NrbStats *stats);

/* SceneNode::cast (src/scene_node.rs:51-75) on node `node` with a world-space ray.
 * Returns 1 on hit.  uv_present receives 0/1. */
int nro_cast(NroScene *s, uint32_t node, const double o[3], const double d[3], double *toi, double normal[3],
             double uv[2], int *uv_present);

/* Scene::trace (src/scene.rs:163-193) for one ray; pixel/sample feed the RNG counter. */
int nro_trace(NroScene *s, const NrbCamera *cam, const double o[3], const double d[3], uint32_t pixel,
              uint32_t sample, float rgb[3]);

/* Scene::intersects_ray (src/scene.rs:147-161).  Returns 1 when Some(filter) (not occluded). */
int nro_intersects_ray(NroScene *s, const double o[3], const double d[3], double maxtoi, float filter[3]);

/* Texture2d::sample (src/texture2d.rs:207-256). */
int nro_texture_sample(NroScene *s, uint32_t texture, double u, double v, float rgba[4]);

/* AABB::toi_with_ray(identity, ray, solid).  Returns 1 on hit. */
int nro_aabb_toi(const double mins[3], const double maxs[3], const double o[3], const double d[3], int solid,
                 double *toi);

/* Primary ray of sample `sample` of pixel `pixel` (src/scene.rs:68-86). */
int nro_primary_ray(const NrbCamera *cam, uint32_t pixel, uint32_t sample, double o[3], double d[3]);

/* Philox4x32-10 block (the RNG shared by oracle and device). */
void nro_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

const char *nro_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
