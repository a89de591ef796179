// api_impl.h -- the C ABI of include/blingcu.h, written once over Pipeline<Backend>.
// BL_DEFINE_API(blingcu, CudaBackend) in cuda_backend.cu produces the product entry points;
// tests/emu/emu.cpp instantiates the same text with prefix `blingemu` for the CPU kernel-body emulator.
#pragma once
#include "pipeline.h"
#include <new>

namespace bl {
template <class Backend> struct Ctx { Pipeline<Backend> p; };
static thread_local std::string g_createError;
}

#define BL_DEFINE_API(PFX, BACKEND)                                                                                              \
   struct PFX##_ctx { bl::Pipeline<BACKEND> p; };                                                                                \
   typedef PFX##_ctx PFX##_ctx_t;                                                                                                \
   extern "C" {                                                                                                                  \
   int PFX##_create(int device, PFX##_ctx_t **out) {                                                                             \
      if (!out) return BLINGCU_EINVAL;                                                                                           \
      *out = nullptr;                                                                                                            \
      PFX##_ctx_t *c = new (std::nothrow) PFX##_ctx_t();                                                                         \
      if (!c) return BLINGCU_EINVAL;                                                                                             \
      int rc = c->p.be.init(device, bl::g_createError);                                                                          \
      if (rc) { delete c; return rc; }                                                                                           \
      *out = c;                                                                                                                  \
      return 0;                                                                                                                  \
   }                                                                                                                             \
   void PFX##_destroy(PFX##_ctx_t *c) { if (c) { c->p.freeState(); c->p.freeScene(); c->p.be.shutdown(); delete c; } }           \
   const char *PFX##_last_error(const PFX##_ctx_t *c) { return c ? c->p.err.c_str() : bl::g_createError.c_str(); }               \
   int PFX##_upload_scene(PFX##_ctx_t *c, const blingcu_scene *ir) { return c ? c->p.be.guard(c->p.err, [&]() { return c->p.upload(ir); }) : BLINGCU_EINVAL; } \
   int PFX##_trace_nearest(PFX##_ctx_t *c, const blingcu_ray *r, size_t n, blingcu_hit *o) {                                     \
      if (!c || (n && (!r || !o))) return BLINGCU_EINVAL;                                                                        \
      return c->p.be.guard(c->p.err, [&]() { return c->p.traceBatch(r, n, o, nullptr, nullptr, nullptr); });                      \
   }                                                                                                                             \
   int PFX##_trace_occluded(PFX##_ctx_t *c, const blingcu_ray *r, size_t n, uint8_t *o) {                                        \
      if (!c || (n && (!r || !o))) return BLINGCU_EINVAL;                                                                        \
      return c->p.be.guard(c->p.err, [&]() { return c->p.traceBatch(r, n, nullptr, o, nullptr, nullptr); });                      \
   }                                                                                                                             \
   int PFX##_trace_stats(PFX##_ctx_t *c, const blingcu_ray *r, size_t n, blingcu_hit *o, uint32_t *nodes, uint32_t *prims) {     \
      if (!c || (n && (!r || !o || !nodes || !prims))) return BLINGCU_EINVAL;                                                    \
      return c->p.be.guard(c->p.err, [&]() { return c->p.traceBatch(r, n, o, nullptr, nodes, prims); });                          \
   }                                                                                                                             \
   int PFX##_render_slice(PFX##_ctx_t *c, uint32_t pass, uint64_t seed, uint32_t s0, uint32_t s1) {                              \
      return c ? c->p.be.guard(c->p.err, [&]() { return c->p.renderSlice(pass, seed, s0, s1); }) : BLINGCU_EINVAL;                \
   }                                                                                                                             \
   int PFX##_render_pass(PFX##_ctx_t *c, uint32_t pass, uint64_t seed) {                                                         \
      if (!c) return BLINGCU_EINVAL;                                                                                             \
      if (!c->p.uploaded) { c->p.err = "render before upload_scene"; return BLINGCU_ESTATE; }                                    \
      return PFX##_render_slice(c, pass, seed, 0, (uint32_t)(c->p.hs.nu * c->p.hs.nv));                                          \
   }                                                                                                                             \
   int PFX##_render_samples(PFX##_ctx_t *c, uint32_t pass, uint64_t seed, const int32_t *px, const int32_t *py,                  \
                            const uint32_t *s, size_t n, float *L, float *xy) {                                                  \
      if (!c || (n && (!px || !py || !s || !L || !xy))) return BLINGCU_EINVAL;                                                   \
      return c->p.be.guard(c->p.err, [&]() { return c->p.renderSamples(pass, seed, px, py, s, n, L, xy); });                      \
   }                                                                                                                             \
   int PFX##_eval_texture(PFX##_ctx_t *c, int32_t tex, const float *p, const float *uv, size_t n, float *out) {                  \
      if (!c || (n && (!p || !uv || !out))) return BLINGCU_EINVAL;                                                               \
      return c->p.be.guard(c->p.err, [&]() { return c->p.evalTexture(tex, p, uv, n, out); });                                    \
   }                                                                                                                             \
   int PFX##_read_film(PFX##_ctx_t *c, float *wxyz) {                                                                            \
      if (!c || !wxyz) return BLINGCU_EINVAL;                                                                                    \
      if (!c->p.uploaded) { c->p.err = "no scene"; return BLINGCU_ESTATE; }                                                      \
      return c->p.be.guard(c->p.err, [&]() { c->p.be.sync(); c->p.be.download(wxyz, c->p.film, sizeof(bl::F4) * (size_t)c->p.hs.W * c->p.hs.H); return 0; }); \
   }                                                                                                                             \
   int PFX##_clear_film(PFX##_ctx_t *c) {                                                                                        \
      if (!c) return BLINGCU_EINVAL;                                                                                             \
      if (!c->p.uploaded) { c->p.err = "no scene"; return BLINGCU_ESTATE; }                                                      \
      return c->p.be.guard(c->p.err, [&]() { c->p.be.waitReduced(); if (c->p.splat) c->p.be.zero(c->p.splat, sizeof(float) * 3 * (size_t)c->p.hs.W * c->p.hs.H); c->p.be.zero(c->p.film, sizeof(bl::F4) * (size_t)c->p.hs.W * c->p.hs.H); return 0; }); \
   }                                                                                                                             \
   int PFX##_film_add_host(PFX##_ctx_t *c, const float *wxyz) {                                                                  \
      if (!c || !wxyz) return BLINGCU_EINVAL;                                                                                    \
      if (!c->p.uploaded) { c->p.err = "no scene"; return BLINGCU_ESTATE; }                                                      \
      return c->p.be.guard(c->p.err, [&]() {                                                                                      \
         size_t n = (size_t)c->p.hs.W * c->p.hs.H;                                                                               \
         bl::F4 *tmp = (bl::F4 *)c->p.be.alloc(sizeof(bl::F4) * n);                                                              \
         c->p.be.upload(tmp, wxyz, sizeof(bl::F4) * n);                                                                          \
         c->p.be.waitReduced();                                                                                                  \
         c->p.be.run(bl::AddFilmBody{c->p.film, tmp}, (uint32_t)n);                                                              \
         c->p.be.sync(); c->p.be.free(tmp);                                                                                      \
         return 0; });                                                                                                           \
   }                                                                                                                             \
   int PFX##_film_device(PFX##_ctx_t *c, void **dptr, size_t *nf) {                                                              \
      if (!c || !dptr || !nf) return BLINGCU_EINVAL;                                                                             \
      if (!c->p.uploaded) { c->p.err = "no scene"; return BLINGCU_ESTATE; }                                                      \
      *dptr = c->p.film; *nf = (size_t)c->p.hs.W * c->p.hs.H * 4;                                                                \
      return 0;                                                                                                                  \
   }                                                                                                                             \
   int PFX##_light_trace(PFX##_ctx_t *c, uint32_t pass, uint64_t seed, uint64_t first, uint32_t n) {                             \
      return c ? c->p.be.guard(c->p.err, [&]() { return c->p.lightTrace(pass, seed, first, n, nullptr, 0, nullptr); }) : BLINGCU_EINVAL; \
   }                                                                                                                             \
   int PFX##_light_trace_records(PFX##_ctx_t *c, uint32_t pass, uint64_t seed, uint64_t first, uint32_t n, float *out, size_t maxr, size_t *nr) { \
      if (!c || !nr || (maxr && !out)) return BLINGCU_EINVAL;                                                                    \
      return c->p.be.guard(c->p.err, [&]() { return c->p.lightTrace(pass, seed, first, n, out, maxr, nr); });                     \
   }                                                                                                                             \
   int PFX##_read_splat(PFX##_ctx_t *c, float *xyz) {                                                                            \
      if (!c || !xyz) return BLINGCU_EINVAL;                                                                                     \
      if (!c->p.uploaded) { c->p.err = "no scene"; return BLINGCU_ESTATE; }                                                      \
      return c->p.be.guard(c->p.err, [&]() {                                                                                      \
         size_t n = (size_t)c->p.hs.W * c->p.hs.H * 3;                                                                           \
         if (!c->p.splat) { std::memset(xyz, 0, n * sizeof(float)); return 0; }                                                  \
         c->p.be.sync(); c->p.be.download(xyz, c->p.splat, n * sizeof(float)); return 0; });                                      \
   }                                                                                                                             \
   int PFX##_upload_kdtree(PFX##_ctx_t *c, const blingcu_kdnode *nodes, uint32_t nn, int32_t root, const uint32_t *leaf, size_t nl, const float *bounds) { \
      return c ? c->p.be.guard(c->p.err, [&]() { return c->p.uploadKd(nodes, nn, root, leaf, nl, bounds); }) : BLINGCU_EINVAL;       \
   }                                                                                                                             \
   int PFX##_trace_kdtree(PFX##_ctx_t *c, const blingcu_ray *r, size_t n, blingcu_hit *o, uint32_t *nodes, uint32_t *prims) {    \
      if (!c || (n && (!r || !o || !nodes || !prims))) return BLINGCU_EINVAL;                                                    \
      return c->p.be.guard(c->p.err, [&]() { return c->p.traceKd(r, n, o, nodes, prims); });                                      \
   }                                                                                                                             \
   int PFX##_host_alloc(PFX##_ctx_t *c, size_t bytes, void **out) {                                                              \
      if (!c || !out) return BLINGCU_EINVAL;                                                                                     \
      *out = nullptr;                                                                                                            \
      return c->p.be.guard(c->p.err, [&]() { *out = c->p.be.hostAlloc(bytes); return 0; });                                       \
   }                                                                                                                             \
   int PFX##_host_free(PFX##_ctx_t *c, void *p) { if (!c) return BLINGCU_EINVAL; return c->p.be.guard(c->p.err, [&]() { c->p.be.hostFree(p); return 0; }); } \
   int PFX##_comm_unique_id(uint8_t *id) {                                                                                       \
      if (!id) return BLINGCU_EINVAL;                                                                                            \
      return BACKEND::commUniqueId(id, bl::g_createError);                                                                       \
   }                                                                                                                             \
   int PFX##_comm_init(PFX##_ctx_t *c, int rank, int nranks, const uint8_t *id) {                                                \
      if (!c || (nranks > 1 && !id)) return BLINGCU_EINVAL;                                                                      \
      return c->p.be.guard(c->p.err, [&]() { return c->p.be.commInit(rank, nranks, id, c->p.err); });                             \
   }                                                                                                                             \
   int PFX##_comm_init_all(PFX##_ctx_t *const *cs, int n) {                                                                      \
      if (!cs || n < 1) return BLINGCU_EINVAL;                                                                                   \
      for (int i = 0; i < n; ++i) if (!cs[i]) return BLINGCU_EINVAL;                                                             \
      uint8_t id[BLINGCU_COMM_ID_BYTES] = {0};                                                                                   \
      int rc = 0;                                                                                                                \
      if (n > 1) { rc = BACKEND::commUniqueId(id, cs[0]->p.err); if (rc) return rc; rc = BACKEND::groupStart(cs[0]->p.err); if (rc) return rc; } \
      for (int i = 0; i < n && !rc; ++i) rc = cs[i]->p.be.guard(cs[i]->p.err, [&]() { return cs[i]->p.be.commInit(i, n, id, cs[i]->p.err); }); \
      if (n > 1) { int re = BACKEND::groupEnd(cs[0]->p.err); if (!rc) rc = re; }                                                  \
      return rc;                                                                                                                 \
   }                                                                                                                             \
   int PFX##_comm_destroy(PFX##_ctx_t *c) { if (!c) return BLINGCU_EINVAL; return c->p.be.guard(c->p.err, [&]() { c->p.be.commDestroy(); return 0; }); } \
   int PFX##_reduce_film(PFX##_ctx_t *c, int root) {                                                                             \
      return c ? c->p.be.guard(c->p.err, [&]() { return c->p.reduceFilm(root); }) : BLINGCU_EINVAL;                               \
   }                                                                                                                             \
   int PFX##_reduce_film_group(PFX##_ctx_t *const *cs, int n, int root) {                                                        \
      if (!cs || n < 1) return BLINGCU_EINVAL;                                                                                   \
      for (int i = 0; i < n; ++i) if (!cs[i]) return BLINGCU_EINVAL;                                                             \
      int rc = 0;                                                                                                                \
      if (n > 1) { rc = BACKEND::groupStart(cs[0]->p.err); if (rc) return rc; }                                                   \
      for (int i = 0; i < n && !rc; ++i) rc = cs[i]->p.be.guard(cs[i]->p.err, [&]() { return cs[i]->p.reduceFilm(root); });       \
      if (n > 1) { int re = BACKEND::groupEnd(cs[0]->p.err); if (!rc) rc = re; }                                                  \
      return rc;                                                                                                                 \
   }                                                                                                                             \
   int PFX##_comm_wait(PFX##_ctx_t *c) { if (!c) return BLINGCU_EINVAL; return c->p.be.guard(c->p.err, [&]() { c->p.be.waitReduced(); return 0; }); } \
   int PFX##_read_film_sum(PFX##_ctx_t *c, float *wxyz) {                                                                        \
      if (!c || !wxyz) return BLINGCU_EINVAL;                                                                                    \
      if (!c->p.uploaded || !c->p.filmSum) { c->p.err = "read_film_sum before reduce_film"; return BLINGCU_ESTATE; }             \
      return c->p.be.guard(c->p.err, [&]() { c->p.be.downloadOnComm(wxyz, c->p.filmSum, sizeof(bl::F4) * (size_t)c->p.hs.W * c->p.hs.H); return 0; }); \
   }                                                                                                                             \
   int PFX##_film_sum_device(PFX##_ctx_t *c, void **dptr, size_t *nf) {                                                          \
      if (!c || !dptr || !nf) return BLINGCU_EINVAL;                                                                             \
      if (!c->p.uploaded || !c->p.filmSum) { c->p.err = "film_sum_device before reduce_film"; return BLINGCU_ESTATE; }           \
      *dptr = c->p.filmSum; *nf = (size_t)c->p.hs.W * c->p.hs.H * 4;                                                             \
      return 0;                                                                                                                  \
   }                                                                                                                             \
   int PFX##_set_stream(PFX##_ctx_t *c, void *stream) { return c ? c->p.be.guard(c->p.err, [&]() { c->p.be.setStream(stream); return 0; }) : BLINGCU_EINVAL; } \
   int PFX##_synchronize(PFX##_ctx_t *c) { return c ? c->p.be.guard(c->p.err, [&]() { c->p.be.sync(); return 0; }) : BLINGCU_EINVAL; } \
   int PFX##_get_stats(PFX##_ctx_t *c, blingcu_stats *o) { return (c && o) ? c->p.be.guard(c->p.err, [&]() { return c->p.getStats(o); }) : BLINGCU_EINVAL; } \
   int PFX##_reset_stats(PFX##_ctx_t *c) { if (!c) return BLINGCU_EINVAL; return c->p.be.guard(c->p.err, [&]() { c->p.resetStats(); return 0; }); } \
   int PFX##_set_option(PFX##_ctx_t *c, const char *key, double v) {                                                             \
      if (!c || !key) return BLINGCU_EINVAL;                                                                                     \
      std::string k(key);                                                                                                        \
      if (k == "batch_samples") { if (!(v >= 1) || v > 2147483648.0) return BLINGCU_EINVAL; c->p.batchTarget = (uint32_t)v; return 0; }                  \
      if (k == "bvh_leaf") { if (v < 1 || v > 15) return BLINGCU_EINVAL; c->p.maxLeaf = (int)v; return 0; }                       \
      if (k == "bvh_trav_cost") { if (!(v >= 0 && v <= 64)) return BLINGCU_EINVAL; c->p.bvhTravCost = (float)v; return 0; }         \
      if (k == "bvh_collapse_cp") { if (!(v >= 0 && v <= 64)) return BLINGCU_EINVAL; c->p.bvhCollapseCp = (float)v; return 0; }      \
      if (k == "bvh_force_leaf") { c->p.bvhForceLeaf = v != 0 ? 1 : 0; return 0; }                                               \
      if (k == "fuse_resolve") { c->p.fuseResolveOpt = v < 0 ? -1 : (v > 0 ? 1 : 0); return 0; }                                  \
      if (c->p.be.setOption(k, v)) return 0;                                                                                     \
      c->p.err = "unknown option " + k;                                                                                          \
      return BLINGCU_EINVAL;                                                                                                     \
   }                                                                                                                             \
   int PFX##_kernel_times(PFX##_ctx_t *c, double *ms, uint64_t *launches, int n) {                                               \
      if (!c || !ms || !launches || n < 0) return BLINGCU_EINVAL;                                                               \
      return c->p.be.guard(c->p.err, [&]() { c->p.be.kernelTimes(ms, launches, n); return 0; });                                  \
   }                                                                                                                             \
   int PFX##_sample_extent(PFX##_ctx_t *c, int32_t *x0, int32_t *x1, int32_t *y0, int32_t *y1) {                                 \
      if (!c || !x0 || !x1 || !y0 || !y1) return BLINGCU_EINVAL;                                                                 \
      if (!c->p.uploaded) { c->p.err = "no scene"; return BLINGCU_ESTATE; }                                                      \
      *x0 = c->p.hs.ex0; *x1 = c->p.hs.ex1; *y0 = c->p.hs.ey0; *y1 = c->p.hs.ey1;                                                \
      return 0;                                                                                                                  \
   }                                                                                                                             \
   }
