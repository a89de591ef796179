/*
 * blingcu.h -- C ABI of the B200 path-tracing core for waldheinz/bling.
 *
 * This is the drop-in boundary for ONE hot path of the reference: the body of
 *   prender / tile            (src/lib/Graphics/Bling/Rendering.hs:111-150)
 * i.e. sampler -> fireRay -> surfaceLi (path integrator + NEE) -> addSample.
 * A Haskell `Renderer` instance (Rendering.hs:77-78) binds these entry points
 * with `foreign import ccall safe` (see INTEGRATION.md and haskell/).
 *
 * Conventions
 *   - every function returns 0 on success, a BLINGCU_E* code otherwise; the
 *     message is available from blingcu_last_error().
 *   - all input buffers are caller-owned HOST memory and are copied on upload.
 *   - all output buffers are caller-allocated HOST memory unless the name says
 *     `_device`.
 *   - POD structs, fixed-width types only; no C++ types, no torch types.
 *   - a context belongs to one host thread at a time and to ONE GPU; multi-GPU
 *     runs use one context per GPU (one process per GPU) and sum films
 *     (blingcu_film_device + an NCCL allreduce on the host side, or
 *     blingcu_film_add_host).
 *
 * The flat scene IR (`blingcu_scene`) is what the Haskell host produces where
 * the scene is built (the parser), because the reference's scene objects are
 * closures (Primitive.hs:21-27, Reflection.hs:42) and cannot be flattened later.
 */
#ifndef BLINGCU_H
#define BLINGCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLINGCU_BANDS 16 /* Spectrum.hs:34-36: 16 bands, 400-700nm */

enum {
   BLINGCU_OK = 0,
   BLINGCU_EINVAL = 1,  /* bad argument / malformed IR                    */
   BLINGCU_ECUDA = 2,   /* CUDA runtime error                             */
   BLINGCU_ENOGPU = 3,  /* no usable device: there is NO CPU fallback      */
   BLINGCU_ESTATE = 4,  /* call out of order (e.g. render before upload)  */
   BLINGCU_EUNSUPPORTED = 5 /* NCCL is not available to this process (multi-GPU film reduction only) */
};

/* Shape.hs:23-38 */
enum {
   BLINGCU_SHAPE_BOX = 0,      /* p[0..2]=pmin p[3..5]=pmax               */
   BLINGCU_SHAPE_CYLINDER = 1, /* p = radius zmin zmax phimax(rad)        */
   BLINGCU_SHAPE_DISK = 2,     /* p = height radius innerRadius phimax    */
   BLINGCU_SHAPE_QUAD = 3,     /* p = sx sy                               */
   BLINGCU_SHAPE_SPHERE = 4    /* p = radius                              */
};

/* Material.hs:32-127: the five materials the configs use, blackbody, and the rest of the module */
enum {
   BLINGCU_MAT_MATTE = 0,   /* tex[0]=kd          f[0]=sigma              */
   BLINGCU_MAT_GLASS = 1,   /* tex[0]=kr tex[1]=kt f[0]=ior               */
   BLINGCU_MAT_MIRROR = 2,  /* tex[0]=kr                                  */
   BLINGCU_MAT_PLASTIC = 3, /* tex[0]=kd tex[1]=ks f[0]=rough             */
   BLINGCU_MAT_METAL = 4,   /* tex[0]=eta tex[1]=k f[0]=rough             */
   BLINGCU_MAT_BLACKBODY = 5, /* Reflection.hs:337-338: no BxDFs          */
   /* SURVEY §8(f)2, first slice of the remaining materials (Material.hs:43-53,98-108): */
   BLINGCU_MAT_SHINYMETAL = 6, /* tex[0]=eta(ks) tex[1]=k(ks) tex[2]=eta(kr) tex3=k(kr) f[0]=rough: the HOST applies
                                  frApproxEta / frApproxK (Fresnel.hs:72-78) to the leaves of the ks / kr texture trees */
   BLINGCU_MAT_TRANSMATTE = 7, /* tex[0]=kr tex[1]=kt f[0]=sigma (translucentMatte)                                    */
   BLINGCU_MAT_SUBSTRATE = 8,  /* tex[0]=kd tex[1]=ks tex[2]=ka f[0]=urough f[1]=vrough f[2]=depth (mkSubstrate:
                                  FresnelBlend over an Anisotropic distribution, Microfacet.hs:56-108,140-192)          */
   BLINGCU_MAT_KINDS = 9
};

/* Texture.hs:159-414. Spectrum textures (0..15) and scalar textures (16..) share one table.
 * 2-D mappings (TextureMapping2d, Texture.hs:166-181) are encoded in s.v: s.v[0] = 0 uvMapping, s.v[1..4] = su sv ou ov;
 * s.v[0] = 1 planarMapping, s.v[1..3] = vu, s.v[4..6] = vv, s.v[7..8] = ou ov.
 * 3-D mappings (identityMapping3d, Texture.hs:150-152) are the 16 floats of the world-to-texture matrix in s.v. */
enum {
   BLINGCU_TEX_CONSTANT = 0,   /* s (as a gradient step: f[0] = position)  */
   BLINGCU_TEX_GRAPHPAPER = 1, /* f[0]=lineWidth, child[0]=paper child[1]=line; aux=0: uvMapping in f[1..4]=su sv ou ov;
                                  aux=1: 2-D mapping in s.v (see above)                                                  */
   BLINGCU_TEX_CHECKER = 2,    /* checkerBoard (Texture.hs:209-221): f[0..2]=scale, child[0]=tex1 child[1]=tex2, on dgP  */
   /* SURVEY §8(f)2, second slice: textures that COMPUTE a spectrum (Texture.hs:129-141,223-253) */
   BLINGCU_TEX_BLEND = 3,      /* spectrumBlend: child[0]=tex1 child[1]=tex2 aux=scalar texture f                        */
   BLINGCU_TEX_GRADIENT = 4,   /* gradient: aux=scalar texture, child[0]=index of the first step, child[1]=step count; the
                                  steps are consecutive CONSTANT entries sorted by position f[0] (mkGradient sorts)      */
   BLINGCU_TEX_IMAGE = 5,      /* imageTexture (Texture.hs:96-108,125-126): aux = index into images (3 channels),
                                  s.v = 2-D mapping; the pixel becomes a spectrum with rgbToSpectrumRefl (refl_basis)    */
   /* scalar textures (MaterialParser.hs:123-154) */
   BLINGCU_STEX_CONSTANT = 16, /* f[0]                                                                                   */
   BLINGCU_STEX_SCALE = 17,    /* scaleTexture a s t = a + s * t: f[0]=a f[1]=s child[0]=t                               */
   BLINGCU_STEX_PERLIN = 18,   /* noiseTexture (Texture.hs:343-380): s.v = 3-D mapping                                   */
   BLINGCU_STEX_FBM = 19,      /* fbm (Texture.hs:329-339): aux=octaves f[0]=omega, s.v = 3-D mapping                    */
   BLINGCU_STEX_CELLNOISE = 20,/* cellNoise (Texture.hs:255-303): aux = distance (0 euclidian, 1 euclidian2, 2 manhattan,
                                  3 chebyshev), s.v = 3-D mapping                                                        */
   BLINGCU_STEX_CRYSTAL = 21,  /* quasiCrystal (Texture.hs:305-326): aux=octaves, s.v = 2-D mapping                      */
   BLINGCU_STEX_IMAGE = 22     /* imageTexture over readImageScalarMap (Texture.hs:103-108,117-126): aux = index into
                                  images (1 channel), s.v = 2-D mapping                                                  */
};

/* Light.hs:31-45 */
enum {
   BLINGCU_LIGHT_INFINITE = 0,    /* env = index into envs                 */
   BLINGCU_LIGHT_DIRECTIONAL = 1, /* s = radiance, v = normalized normal   */
   BLINGCU_LIGHT_POINT = 2,       /* s = intensity, v = position           */
   BLINGCU_LIGHT_AREA = 3         /* s = radiance, shape = index into shapes */
};

enum {
   BLINGCU_ENV_CONSTANT = 0, /* s                                          */
   BLINGCU_ENV_RGBTABLE = 1, /* rgb[h][w][3], nearest, flipped (IO/Bitmap.hs:22-29) */
   BLINGCU_ENV_SUNSKY = 2    /* analytic Perez model (SunSky.hs:12-94)     */
};

enum { BLINGCU_CAM_PERSPECTIVE = 0, BLINGCU_CAM_ENVIRONMENT = 1 };
enum { BLINGCU_SAMPLER_STRATIFIED = 0, BLINGCU_SAMPLER_RANDOM = 1 };
/* Integrator/Path.hs:30-39 (max_depth, sample_depth) and, SURVEY §8(f)4, Integrator/DirectLighting.hs:13-21 (max_depth) */
enum { BLINGCU_INTEGRATOR_PATH = 0, BLINGCU_INTEGRATOR_DIRECT = 1,
       BLINGCU_INTEGRATOR_NORMALS = 2 /* `debug normals`: mkNormalMap (Integrator/Debug.hs:23-33), needs refl_basis */,
       BLINGCU_INTEGRATOR_BIDIR = 3 /* `bidir maxDepth sampleDepth`: mkBidirPathIntegrator (Integrator/BidirPath.hs:44-104) */ };

typedef struct blingcu_spectrum { float v[BLINGCU_BANDS]; } blingcu_spectrum;

/* One analytic shape wrapped by mkGeom (Primitive/Geometry.hs:14-36).
 * Matrices are row-major 4x4, m[r*4+c] (Transform.hs:40-42). */
typedef struct blingcu_shape {
   int32_t kind;
   int32_t material;  /* index into materials                              */
   int32_t light;     /* index into lights (area light) or -1              */
   int32_t prim_id;   /* position in the list handed to mkScene            */
   float p[8];
   float o2w[16];
   float w2o[16];
} blingcu_shape;

typedef struct blingcu_texture {
   int32_t kind;
   int32_t child[2];
   int32_t aux;       /* third reference / small integer parameter, see the kind list */
   float f[8];
   blingcu_spectrum s;
} blingcu_texture;

/* A decoded image (IO stays on the host: JuicyPixels in bling). channels == 3: (r, g, b) per pixel AFTER `fromIntegral x / 255`
 * and `unGamma` (Texture.hs:87-89, Spectrum.hs:118-120); channels == 1: `fromIntegral x / 255` (Texture.hs:103-104).
 * Row-major, row 0 first, as JuicyPixels' pixelAt i x y. */
typedef struct blingcu_image {
   int32_t width, height, channels, _pad;
   const float *data; /* width * height * channels */
} blingcu_image;

typedef struct blingcu_material {
   int32_t kind;
   int32_t tex[3];
   float f[3];
   int32_t tex3;   /* fourth texture (shinyMetal); same size and offsets as before for everything else */
   /* scalar textures. 0 = none, so a zero-filled tail keeps the constant-parameter meaning;
    * otherwise 1 + index into textures of a BLINGCU_STEX_* entry. */
   int32_t ftex[3]; /* replaces f[i] (sigma / ior / rough / urough / vrough / depth) by a ScalarTexture evaluated at the hit */
   int32_t bump;    /* bumpMapped d mat (Reflection.hs:344-377): displacement texture */
} blingcu_material;

typedef struct blingcu_light {
   int32_t kind;
   int32_t shape; /* AREA: index into shapes                               */
   int32_t env;   /* INFINITE: index into envs                             */
   int32_t _pad;
   float v[4];
   blingcu_spectrum s;
} blingcu_light;

/* SunSky.hs:27-36 SkyData + the precomputed sun radiance */
typedef struct blingcu_sunsky {
   float sun_dir[3];      /* initSky: normalize (worldToLocal basis (normalize sdw)) */
   float sun_theta;
   float sun_disc_dir[3]; /* mkSunSkyLight: normalize (worldToLocal basis sdw), used by sunSpectrum */
   float _pad;
   float perez_x[5], perez_y[5], perez_Y[5];
   float zenith_x, zenith_y, zenith_Y;
   float s0xyz[3], s1xyz[3], s2xyz[3]; /* Spectrum.hs:226-244 chromaticityToXYZ */
   blingcu_spectrum sun_radiance;      /* SunSky.hs:96-126 (black below horizon) */
} blingcu_sunsky;

/* The radiance map + Dist2D of mkInfiniteAreaLight (Light.hs:72-82,
 * Montecarlo.hs:34-104). nu,nv = texSize. */
typedef struct blingcu_envmap {
   int32_t kind;
   int32_t nu, nv;
   int32_t _pad;
   float w2l[16];             /* _infw2l                                    */
   float l2w[16];             /* inverse                                    */
   blingcu_spectrum s;        /* CONSTANT                                   */
   const float *rgb;          /* RGBTABLE: nv*nu*3                          */
   blingcu_sunsky sky;        /* SUNSKY                                     */
   const float *cond_func;    /* nv*nu      conditional distFunc            */
   const float *cond_cdf;     /* nv*(nu+1)  conditional cdf                 */
   const float *cond_int;     /* nv         conditional funcInt             */
   const float *marg_func;    /* nv                                         */
   const float *marg_cdf;     /* nv+1                                       */
   float marg_int;
   float _pad2;
} blingcu_envmap;

typedef struct blingcu_camera {
   int32_t kind;
   float raster2cam[16];      /* Camera.hs:105-116 _raster2cam (matrix)     */
   float cam2world[16];
   float lens_radius, focal_distance;
   float env_sx, env_sy;      /* Environment camera (Camera.hs:70-76)       */
   /* light tracer only (sampleCam, Camera.hs:78-103): world-to-raster `w2r` and the pixel area `ap` of mkProjective (:105-133), as the
    * host computed them; zero when the host does not light-trace */
   float world2raster[16];
   float pixel_area;
} blingcu_camera;

typedef struct blingcu_scene {
   /* triangles (TriangleMesh.hs, IO/WaveFront.hs) in world space */
   uint64_t n_triangles;
   const float *tri_verts;      /* n*9: p1 p2 p3                            */
   const float *tri_uvs;        /* n*6 (TriangleMesh.hs:119-120 default 0,0,1,0,1,1) */
   const float *tri_normals;    /* n*9 or NULL (flat shaded); a triangle with nine zeros is flat shaded (mesh without normals beside smooth ones) */
   const int32_t *tri_material; /* n                                        */
   const int32_t *tri_prim_id;  /* n, or NULL => prim id = prim_id_base + i */
   int32_t tri_prim_id_base;

   uint32_t n_shapes;    const blingcu_shape *shapes;
   uint32_t n_materials; const blingcu_material *materials;
   uint32_t n_textures;  const blingcu_texture *textures;
   uint32_t n_lights;    const blingcu_light *lights; /* sceneLights order, Scene.hs:38-42 */
   uint32_t n_envs;      const blingcu_envmap *envs;

   blingcu_camera camera;

   /* film + filter (Image.hs:40-61, Filter.hs:61-96) */
   int32_t width, height;
   float filter_w, filter_h;
   float filter_table[256];

   /* sampler + integrator (Sampling.hs:112-132, Integrator/Path.hs:30-39) */
   int32_t sampler_kind;
   int32_t nu, nv;       /* stratified nu nv; random: nu = spp, nv = 1      */
   int32_t max_depth, sample_depth;

   /* Spectrum.hs:337-347 (CIE tables as 16-band spectra) and :146-159 (illuminant basis r g b c m y w) */
   blingcu_spectrum cie_x, cie_y, cie_z;
   float cie_y_sum;
   blingcu_spectrum illum_basis[7];

   int32_t integrator_kind; /* BLINGCU_INTEGRATOR_*; 0 (path) keeps every older caller's meaning */

   /* image textures: decoded images and the reflectance basis r g b c m y w of rgbToSpectrumRefl
    * (Spectrum.hs:128-145); unused (0 / NULL) unless a BLINGCU_TEX_IMAGE / BLINGCU_STEX_IMAGE entry exists or the integrator
    * is BLINGCU_INTEGRATOR_NORMALS (basis only) */
   uint32_t n_images;
   const blingcu_image *images;
   blingcu_spectrum refl_basis[7];
} blingcu_scene;

typedef struct blingcu_ray { float o[3]; float tmin; float d[3]; float tmax; } blingcu_ray;

/* nearest-hit record; prim < 0 = miss */
typedef struct blingcu_hit { float t; int32_t prim; float b1, b2; } blingcu_hit;

/* counters; the traversal pair mirrors TraversalStats (Primitive/KdTree.hs:252-258) */
typedef struct blingcu_stats {
   uint64_t samples;
   uint64_t rays_camera, rays_extension, rays_mis, rays_shadow;
   uint64_t dropped_samples;   /* NaN/Inf samples skipped (Image.hs:253-256) */
   uint64_t nodes_traversed, intersections;  /* nearest-hit kernel totals while option "traversal_stats" = 1 */
   uint64_t rays_counted;      /* nearest-hit rays traced while "traversal_stats" = 1 (denominator of the two above) */
   uint64_t kernel_launches;
   uint64_t bvh_nodes, bvh_leaf_items;
   double last_pass_ms;        /* device time of the last render call        */
   uint64_t bvh_max_stack;     /* worst-case traversal stack entries of the uploaded tree */
   uint64_t rays_mis_culled;   /* BSDF-MIS rays of Scene.hs:71-82 NOT traced because they cannot reach the chosen light
                                  (miss its shape / delta light); rays_mis counts the traced ones only */
   uint64_t rays_ext_culled;   /* extension rays NOT traced: their vertex would have depth == maxDepth after a non-specular
                                  bounce, where neither a hit nor a miss contributes (Path.hs:43-51); rays_extension = traced */
   uint64_t rays_mis_any;      /* the part of rays_mis traced as any-hit queries (infinite lights: only hit/miss matters) */
   uint64_t photons, rays_light, rays_connect, splats;   /* light tracer: light paths started, nearest-hit rays (light + continuation), any-hit connection rays, splats that reached the image */
   uint64_t any_nodes_traversed, any_intersections, any_rays_counted; /* the traversal triple of the ANY-hit kernel while "traversal_stats" = 1 */
} blingcu_stats;

typedef struct blingcu_ctx blingcu_ctx;

/* create a context on one CUDA device. Fails with BLINGCU_ENOGPU if there is none. */
int blingcu_create(int device, blingcu_ctx **out);
void blingcu_destroy(blingcu_ctx *);
const char *blingcu_last_error(const blingcu_ctx *); /* ctx may be NULL: last create() error */

/* replaces mkScene (Scene.hs:37-43): copies the IR, builds the BVH, uploads. At most 2^26 primitives (triangles + shapes):
 * the hit reference keeps 26 index bits next to the 5-bit shade-queue slot. */
int blingcu_upload_scene(blingcu_ctx *, const blingcu_scene *ir);

/* replaces scIntersect / occluded (Scene.hs:45-51) on explicit ray batches -- parity check (a) */
int blingcu_trace_nearest(blingcu_ctx *, const blingcu_ray *rays, size_t n, blingcu_hit *out);
int blingcu_trace_occluded(blingcu_ctx *, const blingcu_ray *rays, size_t n, uint8_t *out);
/* Host batches are pipelined in chunks (option "trace_chunk", default 2^20 rays): copy-in, traversal and copy-out of
 * neighbouring chunks overlap. Page-locked buffers (blingcu_host_alloc, or memory the caller registered with CUDA) are
 * transferred in place; pageable ones go through the context's pinned staging ring (option "copy_threads" helper threads). */
int blingcu_host_alloc(blingcu_ctx *, size_t bytes, void **out);   /* page-locked host memory for ray / hit / film buffers */
int blingcu_host_free(blingcu_ctx *, void *p);
/* same as trace_nearest with per-ray node visits / primitive tests (dbgTraverse, KdTree.hs:260-281) */
int blingcu_trace_stats(blingcu_ctx *, const blingcu_ray *rays, size_t n, blingcu_hit *out,
                        uint32_t *nodes, uint32_t *prims);

/* replaces one `onePass` of prender (Rendering.hs:127-138): all nu*nv samples of every pixel of
 * the sample extent, accumulated into the device film. */
int blingcu_render_pass(blingcu_ctx *, uint32_t pass_index, uint64_t seed);
/* the sharding unit: sample indices [s_begin, s_end) of every pixel of pass `pass_index`.
 * render_pass == render_slice(0, nu*nv). */
int blingcu_render_slice(blingcu_ctx *, uint32_t pass_index, uint64_t seed, uint32_t s_begin, uint32_t s_end);

/* SURVEY 8(f)3 -- the HOST's own accelerator as an alternative input: bling's SAH kd-tree (Primitive/KdTree.hs:29-33) flattened
 * into an array. Interior: children `left` / `right`, split position and axis; Leaf: left = -1 and the primitive ids
 * leaf_prims[first .. first + count) in the leaf's own order. blingcu_trace_kdtree walks it exactly as `traverse` does
 * (KdTree.hs:223-242; entered through intersectAABB on `bounds` = lo xyz, hi xyz) on the uploaded scene's primitives and returns,
 * per ray, the hit and the two counters of dbgTraverse / TraversalStats (KdTree.hs:252-281) -- so a host can check the GPU against
 * its own tree node for node. The product's render path does not use it (it traverses the library's BVH). */
typedef struct blingcu_kdnode { int32_t left, right; float split; int32_t axis; uint32_t first, count; } blingcu_kdnode;
int blingcu_upload_kdtree(blingcu_ctx *, const blingcu_kdnode *nodes, uint32_t n_nodes, int32_t root,
                          const uint32_t *leaf_prims, size_t n_leaf_prims, const float bounds[6]);
int blingcu_trace_kdtree(blingcu_ctx *, const blingcu_ray *rays, size_t n, blingcu_hit *out,
                         uint32_t *nodes_traversed, uint32_t *intersections);

/* SURVEY 8(f)4 -- the light tracer (Renderer/LightTracer.hs:1-110) on the same traversal / BSDF kernels: `n_photons` light paths
 * (sampleLightRay, Light.sample' Light.hs:166-213), at every vertex a connection to the camera (sampleCam Camera.hs:78-103,
 * adjoint BSDF Reflection.hs:278-332), Russian roulette 0.8 beyond depth 3, splatted UNFILTERED into the splat buffer
 * `_imgS` = [H][W]{X, Y, Z} f32 (splatSample, Image.hs:201-221). One call = one `replicateM_ ppp oneRay` of a pass; the host
 * reports PassDone n img (1 / (n * ppp)). Photon i of pass p draws from the counter-based stream (seed, p, first_photon + i).
 * Needs a perspective camera with world2raster / pixel_area set. The splat sum is atomic-free and deterministic: the records of
 * one bounce are grouped by pixel with an integer counting sort and added in photon order. blingcu_clear_film clears both buffers. */
int blingcu_light_trace(blingcu_ctx *, uint32_t pass_index, uint64_t seed, uint64_t first_photon, uint32_t n_photons);
int blingcu_read_splat(blingcu_ctx *, float *xyz /* H*W*3 */);
/* debug/parity: the splats of individual photons, in (photon, depth) order: out = n_records * {photon, depth, px, py, X, Y, Z} as
 * floats (photon / depth exact below 2^24); returns the count in *n_records, at most max_records are written */
int blingcu_light_trace_records(blingcu_ctx *, uint32_t pass_index, uint64_t seed, uint64_t first_photon, uint32_t n_photons,
                                float *out, size_t max_records, size_t *n_records);

/* debug/parity: radiance of individual samples (no film). pixel coordinates are in sample-extent
 * space (may be negative, Image.hs:162-168); out_L = n*16 floats; out_xy = n*2 image positions. */
int blingcu_render_samples(blingcu_ctx *, uint32_t pass_index, uint64_t seed, const int32_t *px,
                           const int32_t *py, const uint32_t *sample, size_t n, float *out_L, float *out_xy);

/* debug/parity: `Texture a = DifferentialGeometry -> a` (Texture.hs:60-64) of texture-table entry `texture` at n explicit
 * points: p = n*3 world positions (dgP), uv = n*2 surface parameters (dgU, dgV) -- the only DG fields a texture in scope
 * reads. out = n*16 floats: the spectrum, or for a BLINGCU_STEX_* entry its value in out[16*i] (rest 0). */
int blingcu_eval_texture(blingcu_ctx *, int32_t texture, const float *p, const float *uv, size_t n, float *out);

/* film = Img._imgP layout [H][W]{weight, X*w, Y*w, Z*w} f32 (Image.hs:123-129,291-299) */
int blingcu_read_film(blingcu_ctx *, float *wxyz);
int blingcu_clear_film(blingcu_ctx *);
int blingcu_film_add_host(blingcu_ctx *, const float *wxyz); /* film += host buffer (resume / manual reduce) */
int blingcu_film_device(blingcu_ctx *, void **dptr, size_t *n_floats); /* zero-copy view for the host */
int blingcu_synchronize(blingcu_ctx *);
/* enqueue all further work of this context on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. the
 * stream the host's NCCL all-reduce of the film runs on; NULL restores the context's own stream. Render calls are
 * asynchronous with respect to the host; read_film / get_stats / synchronize wait. */
int blingcu_set_stream(blingcu_ctx *, void *cuda_stream);

/* ---- multi-GPU: the film sum over GPUs, the one exchange of the path. Replaces the sequential addTile merge
 * of prender (Rendering.hs:130-134, Image.hs:178-199). Every GPU holds a full scene replica and renders its share of the
 * sample indices (blingcu_render_slice) into a private film; blingcu_reduce_film sums the films with ncclAllReduce over
 * NVLink / NVSwitch (root >= 0: ncclReduce onto that rank only) into a second device buffer, `film_sum`, so the private films
 * keep accumulating. The library owns the communicator and a reduction stream; NCCL is bound at run time (libnccl.so.2).
 * The reduction is ASYNCHRONOUS: it starts once everything enqueued so far has rendered and overlaps the next render call up
 * to that call's film kernels. blingcu_read_film_sum waits for it.
 *   one process per GPU:  rank 0 calls blingcu_comm_unique_id and hands the 128 bytes to the other processes (MPI, a file,
 *                         torch.distributed ...); every process then calls blingcu_comm_init(ctx, rank, nranks, id);
 *   one process, n GPUs:  (a Haskell host) n contexts on n devices, blingcu_comm_init_all(ctxs, n); per pass
 *                         blingcu_render_slice on each context (asynchronous), then blingcu_reduce_film_group(ctxs, n, root).
 * Without a communicator (or nranks == 1) film_sum is a device copy of film. */
#define BLINGCU_COMM_ID_BYTES 128
int blingcu_comm_unique_id(uint8_t id[BLINGCU_COMM_ID_BYTES]);
int blingcu_comm_init(blingcu_ctx *, int rank, int nranks, const uint8_t id[BLINGCU_COMM_ID_BYTES]);
int blingcu_comm_init_all(blingcu_ctx *const *ctxs, int n);
int blingcu_comm_destroy(blingcu_ctx *);
int blingcu_reduce_film(blingcu_ctx *, int root);
int blingcu_reduce_film_group(blingcu_ctx *const *ctxs, int n, int root);
int blingcu_comm_wait(blingcu_ctx *);                       /* the context's render stream waits (on the device) for the reduction in flight */
int blingcu_read_film_sum(blingcu_ctx *, float *wxyz);      /* waits for the reduction, copies film_sum [H][W][4] to the host */
int blingcu_film_sum_device(blingcu_ctx *, void **dptr, size_t *n_floats);

int blingcu_get_stats(blingcu_ctx *, blingcu_stats *out);
int blingcu_reset_stats(blingcu_ctx *);

/* per-kernel-class device time, measured with CUDA events on the launching stream while option
 * "profile_kernels" = 1. Classes: see BLINGCU_KC_*. Arrays have n_classes entries (<= BLINGCU_KC_COUNT). */
enum {
   BLINGCU_KC_RAYGEN = 0, BLINGCU_KC_TRACE_NEAREST = 1, BLINGCU_KC_TRACE_ANY = 2, BLINGCU_KC_CLASSIFY = 3,
   BLINGCU_KC_SHADE = 4, BLINGCU_KC_RESOLVE = 5, BLINGCU_KC_FILM = 6, BLINGCU_KC_OTHER = 7, BLINGCU_KC_COUNT = 8
};
int blingcu_kernel_times(blingcu_ctx *, double *ms, uint64_t *launches, int n_classes);

/* tuning knobs (optional; returns EINVAL if unknown or out of range). Builder knobs take effect at the next upload_scene:
 *   "batch_samples"    paths per wavefront (1 .. 2^31)
 *   "bvh_leaf"         most items a leaf of the library's BVH may hold (1 .. 15; not set: 3 under the optimal collapse, 2 under the greedy one)
 *   "bvh_collapse_cp"  > 0: SAH-optimal collapse of the binary tree into 4-wide nodes with a primitive test costing this many node
 *                      visits; 0: greedy collapse ("open the largest child"). Not set: 0.5 above 4096 items, greedy below
 *   "bvh_trav_cost", "bvh_force_leaf"  SAH termination of the binary builder (measured flat, left at their defaults)
 *   "trace_variant"    0-3: traversal kernel (3 = the product), "fuse_resolve", "trace_chunk", "copy_threads", "traversal_stats",
 *                      "profile_kernels": see DESIGN.md */
int blingcu_set_option(blingcu_ctx *, const char *key, double value);

/* sample extent of the film: x0,x1,y0,y1 inclusive (Image.hs:162-168) */
int blingcu_sample_extent(blingcu_ctx *, int32_t *x0, int32_t *x1, int32_t *y0, int32_t *y1);

#ifdef __cplusplus
}
#endif
#endif
