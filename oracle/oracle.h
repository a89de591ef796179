/* ORACLE (test infrastructure, NOT product code): C API of the CPU restatement of waldheinz/bling's
 * path-integrator hot path. PARITY UNPINNED (no reference golden vectors exist; GHC absent).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it. */
#ifndef BLING_ORACLE_H
#define BLING_ORACLE_H
#include "../include/blingcu.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct oracle_ctx oracle_ctx;
int oracle_create(const blingcu_scene *ir, int build_kdtree, oracle_ctx **out);
void oracle_destroy(oracle_ctx *);
int oracle_trace_nearest(oracle_ctx *, const blingcu_ray *rays, size_t n, blingcu_hit *out, int mode,
                         uint64_t *nodes_traversed, uint64_t *intersections);
/* the kd-tree of KdTree.hs as a flat array (the form a host hands to blingcu_upload_kdtree) and per-ray dbgTraverse counters */
int oracle_export_kdtree(oracle_ctx *, blingcu_kdnode *nodes, uint32_t *n_nodes, uint32_t *leaf_prims, size_t *n_leaf_prims,
                         int32_t *root, float bounds[6]);   /* nodes / leaf_prims may be NULL: sizes only */
int oracle_trace_kd_stats(oracle_ctx *, const blingcu_ray *rays, size_t n, blingcu_hit *out, uint32_t *nodes_traversed, uint32_t *intersections);
int oracle_trace_occluded(oracle_ctx *, const blingcu_ray *rays, size_t n, uint8_t *out, int mode);
int oracle_sample_extent(oracle_ctx *, int32_t *x0, int32_t *x1, int32_t *y0, int32_t *y1);
int oracle_render_samples(oracle_ctx *, uint32_t pass, uint64_t seed, const int32_t *px, const int32_t *py,
                          const uint32_t *sample, size_t n, float *out_L, float *out_xy);
int oracle_render_slice(oracle_ctx *, uint32_t pass, uint64_t seed, uint32_t s_begin, uint32_t s_end, int nthreads);
int oracle_read_film(oracle_ctx *, float *wxyz);
int oracle_clear_film(oracle_ctx *);
int oracle_get_stats(oracle_ctx *, blingcu_stats *);
int oracle_reset_stats(oracle_ctx *);
/* Renderer/LightTracer.hs restated: photons [first, first + n) of a pass into the splat buffer [H][W]{X,Y,Z};
 * records = optional n*7 floats {photon, depth, px, py, X, Y, Z} in (photon, depth) order */
int oracle_light_trace(oracle_ctx *, uint32_t pass, uint64_t seed, uint64_t first, uint32_t n, float *records, size_t max_records, size_t *n_records);
int oracle_read_splat(oracle_ctx *, float *xyz);
int oracle_eval_texture(oracle_ctx *, int32_t texture, const float *p, const float *uv, size_t n, float *out);
/* single functions of oracle_shade.h at explicit arguments, for tests/test_third_statement.py (a pure-Python statement of the same
 * Haskell text held against this one). what: 1 frDielectric(etai, etat, cosi) -> 16; 2 frConductor(eta16, k16, cosi) -> 16;
 * 3 Blinn (e, wh3, wo3, wi3) -> D, pdf, mfG; 4 microfacet BxDF with frDielectric 1 1.5 (e, r16, wo3, wi3, u1, u2) -> eval16, pdf,
 * sampled f16, wi3, pdf; 5 plastic-like Bsdf [Lambertian kd, Microfacet Blinn e (frDielectric 1 1.5) ks] (kd16, ks16, e, s3, t3, n3,
 * ng3, woW3, wiW3, uComp, u1, u2) -> evalBsdf16, bsdfPdf, sample type, pdf, f16, wi3; 6 shape sampling (kind, p0..p5, pt3, u1, u2, wi3)
 * -> ps3, ns3, shapePdf(pt, wi) */
int oracle_debug_eval(int what, const float *in, float *out);
int oracle_add_sample_tile(oracle_ctx *, int wx0, int wx1, int wy0, int wy1, float sx, float sy, const float *L16,
                           float *out_tile, int *ox, int *oy, int *w, int *h);
#ifdef __cplusplus
}
#endif
#endif
