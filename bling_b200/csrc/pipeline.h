// pipeline.h -- scene upload and the wavefront schedule, templated on the execution backend.
// Backend = CudaBackend (cuda_backend.cu: the product) or the single-threaded kernel-body emulator used by the
// no-GPU CI tests (tests/emu/emu.cpp). The schedule replaces prender/tile (Rendering.hs:111-150):
//
//   per batch of k samples/pixel:  raygen -> { extend-trace -> classify -> shade[miss, matte, glass, ...] ->
//                                  shadow-trace + MIS-trace -> resolve -> advance } x (maxDepth+1)
//                                  -> finalize (spectrum->XYZ) -> film gather
//
// Backend contract: alloc/free/upload/download/zero, run(body, n), runQueue(body, queue, countPtr, cap),
// traceNearest(queue,countPtr,cap,o,d,hit), traceAny(queue,countPtr,cap,o,d,occl), sync().
#pragma once
#include "bodies.h"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace bl {

// noisePerms (Texture.hs:400-414): Ken Perlin's reference permutation, the constant of perlin3d
static const uint8_t kNoisePerm[256] = {
   151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,247,120,234,75,0,26,
   197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,74,165,71,134,139,48,27,166,77,146,
   158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,
   18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,
   206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,
   108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,
   49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,
   215,61,156,180};

template <class Backend>
struct Pipeline {
   Backend be;
   std::string err;
   bool uploaded = false;
   // host copies / geometry
   DScene hs;                 // host-side mirror of the device scene struct (pointers are DEVICE pointers)
   DScene *dscene = nullptr;  // device copy
   std::vector<void *> sceneAllocs;
   PathState ps{}; std::vector<void *> stateAllocs;
   F4 *film = nullptr;
   KdTreeDev kd{}; std::vector<void *> kdAllocs; size_t nPrims = 0; bool kdUploaded = false;   // the host's kd-tree (kdtree.h)
   const int32_t *primRefDev = nullptr;
   float *splat = nullptr;    // [H][W]{X, Y, Z}: the light tracer's unfiltered splat buffer (_imgS, Image.hs:123-129)
   std::vector<void *> ltAllocs; LtRecords ltRec{}; uint32_t *ltCount = nullptr, *ltOffset = nullptr, *ltCursor = nullptr, *ltSorted = nullptr, *ltBlockSum = nullptr; uint64_t ltSplats = 0;
   F4 *filmSum = nullptr;     // sum of the films of all ranks (comm.h), valid after reduce_film
   uint32_t npix = 0, nTextures = 0;
   uint32_t batchTarget = 1u << 26;   // paths per wavefront (~490 B of state each: 33 GB at the cap, sized for 180 GB HBM)
   int maxLeaf = 0;            // option bvh_leaf: most items a leaf may hold; 0 = not set: 3 under the optimal collapse, 2 under the greedy one
   float bvhTravCost = 0; int bvhForceLeaf = 1;   // options bvh_trav_cost / bvh_force_leaf (bvh.h::BvhBuildInput)
   // option bvh_collapse_cp: > 0 = SAH-optimal collapse (bvh_build.cpp::Collapse) with a primitive test costing that many node
   // visits, 0 = the greedy collapse. Not set (< 0): 0.5 -- the ratio measured on cfg 5 (tools/gpu_r02_z1.sh: one node visit per
   // ray costs 7.7 ms of the nearest-hit kernels' 350, one primitive test 4.1) -- for scenes of more than 4096 items, greedy
   // below: in a scene that lives in L1 a node visit costs no more than a test, and merged leaves of large primitives lost the
   // 31-primitive cornell-box 10 % (tools/gpu_r02_z2.sh).
   float bvhCollapseCp = -1;
   uint64_t nNodes = 0, nItems = 0;
   uint64_t launches = 0;
   double lastMs = 0;

   int fail(int code, const std::string &m) { err = m; return code; }

   template <class T> T *up(const T *src, size_t n) {
      if (n == 0) n = 1;
      T *d = (T *)be.alloc(sizeof(T) * n);
      if (src) be.upload(d, src, sizeof(T) * n);
      sceneAllocs.push_back(d);
      return d;
   }
   void freeLt() { for (void *p : ltAllocs) be.free(p); ltAllocs.clear(); ltRec = LtRecords{}; ltCount = nullptr; }
   void freeKd() { for (void *p : kdAllocs) be.free(p); kdAllocs.clear(); kdUploaded = false; }
   void freeScene() { be.syncComm(); freeKd(); freeLt(); if (splat) { be.free(splat); splat = nullptr; } freeTraceScratch(); for (void *p : sceneAllocs) be.free(p); sceneAllocs.clear(); dscene = nullptr; if (film) { be.free(film); film = nullptr; } if (filmSum) { be.free(filmSum); filmSum = nullptr; } uploaded = false; }
   void freeState() { for (void *p : stateAllocs) be.free(p); stateAllocs.clear(); ps = PathState{}; qSlotsAlloc = 0; rootAlloc = 0; bd = BdState{}; bdAlloc = 0; }

   int upload(const blingcu_scene *ir) {
      freeScene();
      if (!ir) return fail(BLINGCU_EINVAL, "null scene");
      if (ir->width <= 0 || ir->height <= 0) return fail(BLINGCU_EINVAL, "bad image size");
      if (ir->nu <= 0 || ir->nv <= 0) return fail(BLINGCU_EINVAL, "bad sampler");
      if (ir->max_depth < 0 || ir->max_depth > 254) return fail(BLINGCU_EINVAL, "max_depth out of range");
      if (ir->integrator_kind < BLINGCU_INTEGRATOR_PATH || ir->integrator_kind > BLINGCU_INTEGRATOR_BIDIR) return fail(BLINGCU_EINVAL, "unknown integrator");
      if (ir->integrator_kind == BLINGCU_INTEGRATOR_BIDIR && (ir->max_depth < 1 || ir->max_depth > BL_BD_MAXDEPTH)) return fail(BLINGCU_EINVAL, "bidirectional integrator: max_depth out of range (1..16)");
      // maxDepth 0 has no meaning there: `cont` stops at d == md with d starting at 1 (DirectLighting.hs:47-49), i.e. never
      if (ir->integrator_kind == BLINGCU_INTEGRATOR_DIRECT && (ir->max_depth < 1 || ir->max_depth > 24)) return fail(BLINGCU_EINVAL, "direct lighting: max_depth out of range");
      size_t nt = (size_t)ir->n_triangles, ns = ir->n_shapes, nprim = nt + ns;
      if (nt && (!ir->tri_verts || !ir->tri_material)) return fail(BLINGCU_EINVAL, "triangles without tri_verts / tri_material");
      if (ns && !ir->shapes) return fail(BLINGCU_EINVAL, "n_shapes > 0 without shapes");
      if ((ir->n_materials && !ir->materials) || (ir->n_textures && !ir->textures) || (ir->n_lights && !ir->lights) || (ir->n_envs && !ir->envs)) return fail(BLINGCU_EINVAL, "table count > 0 with a null table");
      // ---- validate indices
      for (size_t i = 0; i < nt; ++i) if (ir->tri_material[i] < 0 || (uint32_t)ir->tri_material[i] >= ir->n_materials) return fail(BLINGCU_EINVAL, "triangle material out of range");
      for (size_t i = 0; i < ns; ++i) {
         const blingcu_shape &s = ir->shapes[i];
         if (s.material < 0 || (uint32_t)s.material >= ir->n_materials) return fail(BLINGCU_EINVAL, "shape material out of range");
         if (s.light >= (int)ir->n_lights) return fail(BLINGCU_EINVAL, "shape light out of range");
         if (s.kind < 0 || s.kind > BLINGCU_SHAPE_SPHERE) return fail(BLINGCU_EINVAL, "unknown shape kind");
      }
      for (uint32_t i = 0; i < ir->n_materials; ++i) {
         const blingcu_material &m = ir->materials[i];
         if (m.kind < 0 || m.kind >= BLINGCU_MAT_KINDS) return fail(BLINGCU_EINVAL, "unknown material kind");
         for (int k = 0; k < matTexCount(m.kind); ++k) { int tx = k < 3 ? m.tex[k] : m.tex3; if (tx < 0 || (uint32_t)tx >= ir->n_textures) return fail(BLINGCU_EINVAL, "material texture out of range"); }
         for (int k = 0; k < 4; ++k) { int tx = k < 3 ? m.ftex[k] : m.bump; if (tx < 0 || (uint32_t)tx > ir->n_textures) return fail(BLINGCU_EINVAL, "material scalar texture out of range"); }
      }
      {
         int rc = validateTextures(ir);
         if (rc) return rc;
      }
      for (uint32_t i = 0; i < ir->n_lights; ++i) {
         const blingcu_light &l = ir->lights[i];
         if (l.kind == BLINGCU_LIGHT_AREA && (l.shape < 0 || (uint32_t)l.shape >= ir->n_shapes)) return fail(BLINGCU_EINVAL, "area light shape out of range");
         if (l.kind == BLINGCU_LIGHT_INFINITE && (l.env < 0 || (uint32_t)l.env >= ir->n_envs)) return fail(BLINGCU_EINVAL, "infinite light env out of range");
      }
      // ---- prim table + leaf items
      std::vector<uint32_t> primRef(nprim ? nprim : 1, 0xffffffffu);
      std::vector<float> lo(3 * nprim), hi(3 * nprim);
      std::vector<int32_t> itemPrim(nprim);
      for (size_t i = 0; i < nt; ++i) {
         size_t pid = ir->tri_prim_id ? (size_t)ir->tri_prim_id[i] : (size_t)ir->tri_prim_id_base + i;
         if (pid >= nprim || primRef[pid] != 0xffffffffu) return fail(BLINGCU_EINVAL, "prim ids must be a permutation of 0..n-1");
         primRef[pid] = (uint32_t)i;
         const float *v = ir->tri_verts + 9 * i;
         for (int k = 0; k < 3; ++k) {
            lo[3 * i + k] = std::min(v[k], std::min(v[3 + k], v[6 + k]));
            hi[3 * i + k] = std::max(v[k], std::max(v[3 + k], v[6 + k]));
         }
         itemPrim[i] = (int32_t)pid;
      }
      for (size_t j = 0; j < ns; ++j) {
         const blingcu_shape &s = ir->shapes[j];
         size_t pid = (size_t)s.prim_id;
         if (pid >= nprim || primRef[pid] != 0xffffffffu) return fail(BLINGCU_EINVAL, "prim ids must be a permutation of 0..n-1");
         primRef[pid] = 0x80000000u | (uint32_t)j;
         // worldBounds = transBox o2w (objectBounds s) (Shape.hs:287-311)
         const float *P = s.p; float olo[3], ohi[3];
         switch (s.kind) {
         case BLINGCU_SHAPE_BOX: for (int k = 0; k < 3; ++k) { olo[k] = P[k]; ohi[k] = P[3 + k]; } break;
         case BLINGCU_SHAPE_CYLINDER: olo[0] = olo[1] = -P[0]; ohi[0] = ohi[1] = P[0]; olo[2] = P[1]; ohi[2] = P[2]; break;
         case BLINGCU_SHAPE_DISK: olo[0] = olo[1] = -P[1]; ohi[0] = ohi[1] = P[1]; olo[2] = ohi[2] = P[0]; break;
         case BLINGCU_SHAPE_QUAD: olo[0] = -P[0]; ohi[0] = P[0]; olo[1] = -P[1]; ohi[1] = P[1]; olo[2] = ohi[2] = 0; break;
         default: for (int k = 0; k < 3; ++k) { olo[k] = -P[0]; ohi[k] = P[0]; } break;
         }
         size_t it = nt + j;
         for (int k = 0; k < 3; ++k) { lo[3 * it + k] = BL_INF; hi[3 * it + k] = -BL_INF; }
         for (int c = 0; c < 8; ++c) {
            V3 q = transPoint(s.o2w, mk3((c & 4) ? ohi[0] : olo[0], (c & 2) ? ohi[1] : olo[1], (c & 1) ? ohi[2] : olo[2]));
            float qq[3] = {q.x, q.y, q.z};
            for (int k = 0; k < 3; ++k) { lo[3 * it + k] = std::min(lo[3 * it + k], qq[k]); hi[3 * it + k] = std::max(hi[3 * it + k], qq[k]); }
         }
         itemPrim[it] = (int32_t)pid;
      }
      // shade slot per material (bodies.h): its kind's, or its own when its textures compute; the reference carries slot - 1
      scanKinds(ir);
      std::vector<int> shadeKind(ir->n_materials ? ir->n_materials : 1, 0);
      for (uint32_t i = 0; i < ir->n_materials; ++i) shadeKind[i] = matSlot[i] - 1;
      BvhBuildInput bi; bi.n = nprim; bi.lo = lo.data(); bi.hi = hi.data(); bi.trav_cost = bvhTravCost; bi.force_leaf = bvhForceLeaf;
      bi.collapse_cp = bvhCollapseCp >= 0 ? bvhCollapseCp : (nprim > 4096 ? 0.5f : 0.0f);
      bi.max_leaf = maxLeaf > 0 ? maxLeaf : (bi.collapse_cp > 0 ? 3 : 2);
      bi.threads = (int)std::max(1u, std::thread::hardware_concurrency());
      BvhBuildOutput bo;
      if (bvhBuild(bi, bo)) return fail(BLINGCU_EINVAL, "too many primitives");
      std::vector<F4> items((size_t)BL_ITEM_F4 * (nprim ? nprim : 1), F4{0, 0, 0, 0});
      for (size_t k = 0; k < nprim; ++k) {
         uint32_t src = bo.order[k];
         F4 *q = &items[(size_t)BL_ITEM_F4 * k];
         if (src < nt) {
            const float *v = ir->tri_verts + 9 * (size_t)src;
            q[0] = F4{v[0], v[1], v[2], i2f(mkRef(false, shadeKind[ir->tri_material[src]], (uint32_t)src))};
            q[1] = F4{v[3] - v[0], v[4] - v[1], v[5] - v[2], i2f(0)};   // e1 = p2 - p1 (TriangleMesh.hs:169)
            q[2] = F4{v[6] - v[0], v[7] - v[1], v[8] - v[2], 0};        // e2 = p3 - p1
         } else {
            q[0] = F4{0, 0, 0, i2f(mkRef(true, shadeKind[ir->shapes[src - nt].material], (uint32_t)(src - nt)))};
            q[1] = F4{0, 0, 0, i2f(1 + (int)(src - nt))};
            q[2] = F4{0, 0, 0, 0};
         }
      }
      std::vector<int32_t> primHitRef(nprim ? nprim : 1, BL_REF_MISS);
      for (size_t i = 0; i < nt; ++i) primHitRef[itemPrim[i]] = mkRef(false, shadeKind[ir->tri_material[i]], (uint32_t)i);
      for (size_t j = 0; j < ns; ++j) primHitRef[itemPrim[nt + j]] = mkRef(true, shadeKind[ir->shapes[j].material], (uint32_t)j);
      std::memset(&hs, 0, sizeof(hs));
      primRefDev = up<int32_t>(primHitRef.data(), primHitRef.size()); nPrims = nprim;
      hs.bvh.nodes = up<F4>(bo.nodes, BL_NODE_F4 * (size_t)bo.n_nodes);
      hs.bvh.items = up<F4>(items.data(), items.size());
      hs.bvh.root = bo.root; hs.bvh.n_nodes = bo.n_nodes; hs.bvh.max_stack = bo.max_stack;
      if (bo.max_stack > BL_STACK) { std::free(bo.nodes); std::free(bo.order); return fail(BLINGCU_EINVAL, "acceleration structure too deep for the traversal stack"); }
      nNodes = (uint64_t)bo.n_nodes; nItems = nprim; be.setMaxStack(bo.max_stack); be.setFilm(ir->width, ir->height, ir->filter_w, ir->filter_h);
      std::free(bo.nodes); std::free(bo.order);
      // ---- shading geometry
      {
         std::vector<F4> tp(BL_TRI_F4 * (nt ? nt : 1));   // one 64-byte record per triangle (shading.h::DScene)
         for (size_t i = 0; i < nt; ++i) {
            static const float kDefaultUv[6] = {0, 0, 1, 0, 1, 1};   // TriangleMesh.hs:119-120
            const float *v = ir->tri_verts + 9 * i, *u = ir->tri_uvs ? ir->tri_uvs + 6 * i : kDefaultUv;
            tp[BL_TRI_F4 * i] = F4{v[0], v[1], v[2], i2f(ir->tri_material[i])};
            tp[BL_TRI_F4 * i + 1] = F4{v[3], v[4], v[5], u[0]}; tp[BL_TRI_F4 * i + 2] = F4{v[6], v[7], v[8], u[1]};
            tp[BL_TRI_F4 * i + 3] = F4{u[2], u[3], u[4], u[5]};
         }
         hs.tri_p = up<F4>(tp.data(), tp.size());
         hs.tri_n = nullptr;
         if (ir->tri_normals && nt) {
            std::vector<F4> tn(3 * nt);
            for (size_t i = 0; i < nt; ++i) for (int k = 0; k < 3; ++k) { const float *n = ir->tri_normals + 9 * i + 3 * k; tn[3 * i + k] = F4{n[0], n[1], n[2], 0}; }
            hs.tri_n = up<F4>(tn.data(), tn.size());
         }
      }
      { std::vector<int32_t> tprim(nt ? nt : 1, 0); for (size_t i = 0; i < nt; ++i) tprim[i] = itemPrim[i]; hs.tri_prim = up<int32_t>(tprim.data(), tprim.size()); }
      hs.shapes = up<blingcu_shape>(ir->shapes, ns); hs.bvh.shapes = hs.shapes; be.setBvh(hs.bvh);
      hs.materials = up<blingcu_material>(ir->materials, ir->n_materials);
      hs.textures = up<blingcu_texture>(ir->textures, ir->n_textures); nTextures = ir->n_textures;
      hs.lights = up<blingcu_light>(ir->lights, ir->n_lights); hs.n_lights = (int)ir->n_lights;
      hs.has_box = 0; for (size_t j = 0; j < ns; ++j) if (ir->shapes[j].kind == BLINGCU_SHAPE_BOX) hs.has_box = 1;
      {
         std::vector<blingcu_envmap> envs(ir->envs, ir->envs + ir->n_envs);
         for (blingcu_envmap &e : envs) {
            size_t nu = (size_t)e.nu, nv = (size_t)e.nv;
            if (e.nu <= 0 || e.nv <= 0 || !e.cond_func || !e.cond_cdf || !e.cond_int || !e.marg_func || !e.marg_cdf) return fail(BLINGCU_EINVAL, "environment map without distribution");
            if (e.kind == BLINGCU_ENV_RGBTABLE) { if (!e.rgb) return fail(BLINGCU_EINVAL, "rgb table missing"); e.rgb = up<float>(e.rgb, nu * nv * 3); } else e.rgb = nullptr;
            e.cond_func = up<float>(e.cond_func, nu * nv); e.cond_cdf = up<float>(e.cond_cdf, (nu + 1) * nv); e.cond_int = up<float>(e.cond_int, nv);
            e.marg_func = up<float>(e.marg_func, nv); e.marg_cdf = up<float>(e.marg_cdf, nv + 1);
         }
         hs.envs = up<blingcu_envmap>(envs.data(), envs.size());
      }
      {
         std::vector<blingcu_image> imgs(ir->n_images ? ir->images : nullptr, ir->n_images ? ir->images + ir->n_images : nullptr);
         for (blingcu_image &im : imgs) {
            if (im.width <= 0 || im.height <= 0 || im.channels <= 0 || !im.data) return fail(BLINGCU_EINVAL, "empty image");
            im.data = up<float>(im.data, (size_t)im.width * im.height * im.channels);
         }
         hs.images = up<blingcu_image>(imgs.data(), imgs.size());
         for (int b = 0; b < 7; ++b) for (int i = 0; i < NB; ++i) hs.refl[b][i] = ir->refl_basis[b].v[i];
      }
      hs.ftbl = up<float>(ir->filter_table, 256);
      hs.cam = ir->camera;
      hs.W = ir->width; hs.H = ir->height; hs.fw = ir->filter_w; hs.fh = ir->filter_h;
      hs.ex0 = (int)floorf(0.5f - hs.fw); hs.ex1 = (int)floorf(0.5f + (float)hs.W + hs.fw);   // Image.hs:162-168
      hs.ey0 = (int)floorf(0.5f - hs.fh); hs.ey1 = (int)floorf(0.5f + (float)hs.H + hs.fh);
      hs.EW = hs.ex1 - hs.ex0 + 1; hs.EH = hs.ey1 - hs.ey0 + 1;
      hs.sampler_kind = ir->sampler_kind; hs.nu = ir->nu; hs.nv = ir->nv;
      hs.max_depth = ir->max_depth; hs.sample_depth = ir->sample_depth;
      hs.integrator = ir->integrator_kind;
      const bool direct = hs.integrator == BLINGCU_INTEGRATOR_DIRECT;
      const bool normals = hs.integrator == BLINGCU_INTEGRATOR_NORMALS;   // sampleCount1D = sampleCount2D = 0 (Debug.hs:24)
      const bool bidir = hs.integrator == BLINGCU_INTEGRATOR_BIDIR;       // s1d = smps1D * sd * 3 + 1, s2d = smps2D * sd * 3 + 2 (BidirPath.hs:44-48)
      hs.smp = mkSamplerConst(hs.nu, hs.nv, normals ? 0 : (direct ? 2 * hs.max_depth : (bidir ? 12 * hs.sample_depth + 1 : 4 * hs.sample_depth)),
                              normals ? 0 : (direct ? 2 * hs.max_depth : (bidir ? 9 * hs.sample_depth + 2 : 3 * hs.sample_depth)),
                              hs.sampler_kind == BLINGCU_SAMPLER_STRATIFIED);
      hs.smpUniform = mkSamplerConst(1, 1, 0, 0, 0);
      for (int k = 0; k < 3; ++k) { hs.bounds_lo[k] = nprim ? bo.scene_lo[k] : 0.0f; hs.bounds_hi[k] = nprim ? bo.scene_hi[k] : 0.0f; }
      dlHeadroom = 2;
      for (int i = 0; i < NB; ++i) { hs.cieX[i] = ir->cie_x.v[i]; hs.cieY[i] = ir->cie_y.v[i]; hs.cieZ[i] = ir->cie_z.v[i]; }
      hs.ySum = ir->cie_y_sum;
      for (int b = 0; b < 7; ++b) for (int i = 0; i < NB; ++i) hs.illum[b][i] = ir->illum_basis[b].v[i];
      for (int i = 0; i < 256; ++i) hs.perm[i] = kNoisePerm[i];
      dscene = up<DScene>(&hs, 1);
      npix = (uint32_t)hs.EW * (uint32_t)hs.EH;
      film = (F4 *)be.alloc(sizeof(F4) * (size_t)hs.W * hs.H);
      be.zero(film, sizeof(F4) * (size_t)hs.W * hs.H);
      uploaded = true;
      return 0;
   }

   template <class T> T *st(size_t n) { T *p = (T *)be.alloc(sizeof(T) * n); stateAllocs.push_back(p); return p; }
   int qSlotsAlloc = 0;   // shade queues the state holds
   uint32_t rootAlloc = 0;   // entries of ps.root (cap for the direct-lighting integrator, 1 otherwise)
   BdState bd = {};          // vertex records of the bidirectional integrator (bidir.h), allocated with the state when it is the integrator
   int bdAlloc = 0;          // depth the records were allocated for
   int ensureState(uint32_t cap) {
      const uint32_t rootNeed = hs.integrator == BLINGCU_INTEGRATOR_DIRECT ? std::max(cap, ps.cap) : 1u;
      const int bdNeed = hs.integrator == BLINGCU_INTEGRATOR_BIDIR ? hs.max_depth : 0;
      if (ps.cap >= cap && qSlotsAlloc >= nSlots && rootAlloc >= rootNeed && bdAlloc >= bdNeed) return 0;
      cap = std::max(cap, ps.cap);
      freeState();
      ps.cap = cap; qSlotsAlloc = nSlots; rootAlloc = rootNeed; bdAlloc = bdNeed;
      size_t c = cap;
      bd = BdState{};
      if (bdNeed) {
         const size_t v = 2 * (size_t)bdNeed * c;
         bd.md = bdNeed; bd.vRay = st<F4>(2 * v); bd.vHit = st<F4>(v); bd.vWo = st<F4>(v); bd.vAlpha = st<F4>(4 * v);
         bd.nVert = st<uint32_t>(2 * c); bd.D = st<F4>(4 * (size_t)bdNeed * c);
      }
      ps.rayO = st<F4>(2 * c); ps.rayD = ps.rayO + 1; ps.hit = st<F4>(c);   // rays: one interleaved 32-byte record per slot (bodies.h::loadRay)
      ps.T = st<F4>(4 * c); ps.L = st<F4>(4 * c);
      ps.shO = st<F4>(2 * c); ps.shD = ps.shO + 1; ps.PS = st<F4>(4 * c); ps.occl = st<uint8_t>(c); ps.occlM = st<uint8_t>(c);
      ps.miO = st<F4>(2 * c); ps.miD = ps.miO + 1; ps.mihit = st<F4>(c); ps.PM = st<F4>(4 * c); ps.miInfo = st<F2>(c);
      ps.meta = st<uint32_t>(c); ps.kp = st<uint64_t>(c); ps.sidx = st<uint32_t>(c); ps.spos = st<F2>(c); ps.xyz = st<F4>(c);
      ps.qA = st<uint32_t>(c); ps.qB = st<uint32_t>(c); ps.qShadow = st<uint32_t>(c); ps.qMis = st<uint32_t>(c); ps.qMisAny = st<uint32_t>(c);
      ps.qMat = st<uint32_t>((size_t)nSlots * c);
      ps.root = st<uint32_t>(rootAlloc);
      ps.counters = st<uint32_t>(N_COUNTERS); ps.stats = st<unsigned long long>(N_STATS);
      be.zero(ps.counters, sizeof(uint32_t) * N_COUNTERS); be.zero(ps.stats, sizeof(unsigned long long) * N_STATS);
      return 0;
   }

   // the bounce loop over n freshly generated paths (slots 0..n-1, qA = identity)
   void bounces(uint32_t n) {
      uint32_t cap = ps.cap;
      uint32_t *qa = ps.qA, *qb = ps.qB;
      for (int d = 0; d <= hs.max_depth; ++d) {
         uint32_t bound = n;   // queues never exceed n
         be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(qa, ps.counters + C_ACTIVE, bound, dscene, ps.rayO, ps.rayD, ps.hit);
         be.tag(BLINGCU_KC_CLASSIFY); be.runQueue(ClassifyBody{dscene, ps}, qa, ps.counters + C_ACTIVE, bound);
         be.tag(BLINGCU_KC_SHADE); be.runQueue(ShadeMissBody{dscene, ps}, ps.qMat, ps.counters + C_MAT0, bound);
         launches += 3;
         if (d < hs.max_depth) {
            for (int k = 1; k < nSlots; ++k) {
               if (!kindPresent[k]) continue;
               const uint32_t *qk = ps.qMat + (size_t)k * cap; const uint32_t *ck = ps.counters + C_MAT0 + k;
               // one instantiation of the shade kernel per material kind, one more per kind for materials whose textures compute
               // (launched once per such material), and the any-kind one for a slot that several kinds had to share
#define BL_SHADE(K) case K: be.runQueue(ShadeHitBody<K>{dscene, ps, qb}, qk, ck, bound); break; \
                    case SK_TEX0 + K: be.runQueue(ShadeHitBody<SK_TEX0 + K>{dscene, ps, qb}, qk, ck, bound); break;
               switch (slotKernel[k]) {
               BL_SHADE(BLINGCU_MAT_MATTE) BL_SHADE(BLINGCU_MAT_GLASS) BL_SHADE(BLINGCU_MAT_MIRROR) BL_SHADE(BLINGCU_MAT_PLASTIC) BL_SHADE(BLINGCU_MAT_METAL)
               BL_SHADE(BLINGCU_MAT_BLACKBODY) BL_SHADE(BLINGCU_MAT_SHINYMETAL) BL_SHADE(BLINGCU_MAT_TRANSMATTE) BL_SHADE(BLINGCU_MAT_SUBSTRATE)
               default: be.runQueue(ShadeHitBody<SK_GENERAL>{dscene, ps, qb}, qk, ck, bound); break;
               }
#undef BL_SHADE
               launches++;
            }
            // shadow rays (Scene.hs:64). Small scenes: the any-hit kernel resolves the NEE contribution itself (trace_kernels.cuh
            // "fused NEE resolve": one launch and the occlusion-flag round trip less, +1.4 .. 2.9 % on the named scenes). Large
            // scenes keep the separate resolve launch: there the traversal kernel is the bottleneck and the extra scattered
            // read-modify-write inside it costs more than the 9 ms stream it replaces (cfg 5: -1.2 %).
            const bool fuse = fuseResolve() && be.fusesResolve();
            // the nearest-hit BSDF-MIS queue serves area lights, infinite lights in scenes with a Box, and -- with ANY light --
            // the samples whose weight is not finite (bodies.h, BL_MIS_NONFINITE; normally an empty queue: two idle launches)
            const bool misNearest = hs.n_lights > 0;
            be.tag(BLINGCU_KC_TRACE_ANY);
            if (fuse) be.traceAnyFused(ps.qShadow, ps.counters + C_SHADOW, bound, dscene, ps.shO, ps.shD, ps.occl, ps.L, ps.PS, cap);
            else be.traceAny(ps.qShadow, ps.counters + C_SHADOW, bound, dscene, ps.shO, ps.shD, ps.occl);
            launches++;
            if (hasInfinite && !hasBox) { be.traceAny(ps.qMisAny, ps.counters + C_MISANY, bound, dscene, ps.miO, ps.miD, ps.occlM); launches++; }
            if (misNearest) { be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(ps.qMis, ps.counters + C_MIS, bound, dscene, ps.miO, ps.miD, ps.mihit); launches++; }
            be.tag(BLINGCU_KC_RESOLVE);
            if (!fuse) { be.runQueue(ResolveShadowBody{ps}, ps.qShadow, ps.counters + C_SHADOW, bound); launches++; }
            if (misNearest) { be.runQueue(ResolveMisBody{dscene, ps}, ps.qMis, ps.counters + C_MIS, bound); launches++; }
            if (hasInfinite && !hasBox) { be.runQueue(ResolveMisAnyBody{dscene, ps}, ps.qMisAny, ps.counters + C_MISANY, bound); launches++; }
         }
         be.tag(BLINGCU_KC_OTHER); be.run(AdvanceBody{ps}, 1); launches++;
         uint32_t *t = qa; qa = qb; qb = t;
      }
   }

   // ---- textures: validation and the "does this material's texture tree compute" test
   static int matTexCount(int kind) {
      return (kind == BLINGCU_MAT_MATTE || kind == BLINGCU_MAT_MIRROR) ? 1 : (kind == BLINGCU_MAT_BLACKBODY ? 0 : (kind == BLINGCU_MAT_SHINYMETAL ? 4 : (kind == BLINGCU_MAT_SUBSTRATE ? 3 : 2)));
   }
   static bool isScalarKind(int k) { return k >= BLINGCU_STEX_CONSTANT && k <= BLINGCU_STEX_IMAGE; }
   // nesting depth of blends below spectrum texture `id` (-1: malformed); selecting kinds do not count
   static int blendDepth(const blingcu_scene *ir, int id, int guard) {
      if (guard > 32 || id < 0 || (uint32_t)id >= ir->n_textures) return -1;
      const blingcu_texture &t = ir->textures[id];
      if (t.kind == BLINGCU_TEX_CONSTANT || t.kind == BLINGCU_TEX_GRADIENT || t.kind == BLINGCU_TEX_IMAGE) return 0;
      if (t.kind == BLINGCU_TEX_GRAPHPAPER || t.kind == BLINGCU_TEX_CHECKER || t.kind == BLINGCU_TEX_BLEND) {
         int a = blendDepth(ir, t.child[0], guard + 1), b = blendDepth(ir, t.child[1], guard + 1);
         if (a < 0 || b < 0) return -1;
         return std::max(a, b) + (t.kind == BLINGCU_TEX_BLEND ? 1 : 0);
      }
      return -1;   // a scalar texture where a spectrum texture is expected
   }
   static bool spectrumComputes(const blingcu_scene *ir, int id, int guard) {
      if (guard > 32) return true;
      const blingcu_texture &t = ir->textures[id];
      if (t.kind == BLINGCU_TEX_BLEND || t.kind == BLINGCU_TEX_GRADIENT || t.kind == BLINGCU_TEX_IMAGE) return true;
      if (t.kind == BLINGCU_TEX_GRAPHPAPER || t.kind == BLINGCU_TEX_CHECKER) return spectrumComputes(ir, t.child[0], guard + 1) || spectrumComputes(ir, t.child[1], guard + 1);
      return false;
   }
   static bool materialComputes(const blingcu_scene *ir, const blingcu_material &m) {
      if (m.bump || m.ftex[0] || m.ftex[1] || m.ftex[2]) return true;
      for (int k = 0; k < matTexCount(m.kind); ++k) if (spectrumComputes(ir, k < 3 ? m.tex[k] : m.tex3, 0)) return true;
      return false;
   }
   int validateTextures(const blingcu_scene *ir) {
      const uint32_t nt = ir->n_textures;
      auto inRange = [nt](int i) { return i >= 0 && (uint32_t)i < nt; };
      for (uint32_t i = 0; i < nt; ++i) {
         const blingcu_texture &t = ir->textures[i];
         switch (t.kind) {
         case BLINGCU_TEX_CONSTANT: case BLINGCU_STEX_CONSTANT: case BLINGCU_STEX_PERLIN: break;
         case BLINGCU_TEX_GRAPHPAPER: case BLINGCU_TEX_CHECKER:
            if (!inRange(t.child[0]) || !inRange(t.child[1])) return fail(BLINGCU_EINVAL, "texture child out of range");
            break;
         case BLINGCU_TEX_BLEND:
            if (!inRange(t.child[0]) || !inRange(t.child[1]) || !inRange(t.aux)) return fail(BLINGCU_EINVAL, "texture child out of range");
            if (!isScalarKind(ir->textures[t.aux].kind)) return fail(BLINGCU_EINVAL, "blend factor must be a scalar texture");
            break;
         case BLINGCU_TEX_GRADIENT:
            if (!inRange(t.aux) || !isScalarKind(ir->textures[t.aux].kind)) return fail(BLINGCU_EINVAL, "gradient input must be a scalar texture");
            if (t.child[1] < 1 || !inRange(t.child[0]) || !inRange(t.child[0] + t.child[1] - 1)) return fail(BLINGCU_EINVAL, "gradient steps out of range");
            for (int k = 0; k < t.child[1]; ++k) {
               const blingcu_texture &st = ir->textures[t.child[0] + k];
               if (st.kind != BLINGCU_TEX_CONSTANT || (k > 0 && st.f[0] < ir->textures[t.child[0] + k - 1].f[0])) return fail(BLINGCU_EINVAL, "gradient steps must be constant textures sorted by position");
            }
            break;
         case BLINGCU_STEX_SCALE:
            if (!inRange(t.child[0]) || !isScalarKind(ir->textures[t.child[0]].kind)) return fail(BLINGCU_EINVAL, "scale texture child must be a scalar texture");
            break;
         case BLINGCU_STEX_FBM: case BLINGCU_STEX_CRYSTAL:
            if (t.aux < 0 || t.aux > 64) return fail(BLINGCU_EINVAL, "octaves out of range");
            break;
         case BLINGCU_STEX_CELLNOISE:
            if (t.aux < 0 || t.aux > 3) return fail(BLINGCU_EINVAL, "unknown cell-noise distance");
            break;
         case BLINGCU_TEX_IMAGE: case BLINGCU_STEX_IMAGE: {
            if (t.aux < 0 || (uint32_t)t.aux >= ir->n_images || !ir->images) return fail(BLINGCU_EINVAL, "image out of range");
            const blingcu_image &im = ir->images[t.aux];
            if (im.width <= 0 || im.height <= 0 || !im.data) return fail(BLINGCU_EINVAL, "empty image");
            if (im.channels != (t.kind == BLINGCU_TEX_IMAGE ? 3 : 1)) return fail(BLINGCU_EINVAL, "image texture needs 3 channels, scalar image texture 1");
            break;
         }
         default: return fail(BLINGCU_EINVAL, "unknown texture kind");
         }
      }
      for (uint32_t i = 0; i < nt; ++i) {   // scale chains: at most 4 links (evalScalarTexture), no cycles
         int id = (int)i, n = 0;
         while (ir->textures[id].kind == BLINGCU_STEX_SCALE) { if (++n > 4) return fail(BLINGCU_EINVAL, "scale textures nested deeper than 4"); id = ir->textures[id].child[0]; }
      }
      for (uint32_t i = 0; i < ir->n_materials; ++i) {
         const blingcu_material &m = ir->materials[i];
         for (int k = 0; k < matTexCount(m.kind); ++k) {
            int d = blendDepth(ir, k < 3 ? m.tex[k] : m.tex3, 0);
            if (d < 0) return fail(BLINGCU_EINVAL, "material texture is not a spectrum texture tree");
            if (d > BL_BLEND_DEPTH) return fail(BLINGCU_EINVAL, "blend textures nested deeper than 2");
         }
         for (int k = 0; k < 4; ++k) { int tx = k < 3 ? m.ftex[k] : m.bump; if (tx && !isScalarKind(ir->textures[tx - 1].kind)) return fail(BLINGCU_EINVAL, "material scalar parameter must be a scalar texture"); }
      }
      return 0;
   }

   // direct-lighting integrator (bodies.h::DlShadeBody): n camera samples in slots [0, n), spawned branches behind them
   void bouncesDirect(uint32_t n) {
      uint32_t *qa = ps.qA, *qb = ps.qB;
      const uint32_t bound = ps.cap;   // a queue may hold spawned slots too
      const bool misNearest = hs.n_lights > 0;   // as in bounces(): also the home of non-finite BSDF-MIS weights
      for (int d = 0; d < hs.max_depth; ++d) {
         be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(qa, ps.counters + C_ACTIVE, bound, dscene, ps.rayO, ps.rayD, ps.hit);
         // the general shade kernel is instruction-fetch bound: one launch per material kind present keeps the code that is
         // hot at any one time small (profiles/r01_general_shade.md)
         be.tag(BLINGCU_KC_CLASSIFY); be.runQueue(ClassifyBody{dscene, ps}, qa, ps.counters + C_ACTIVE, bound);
         be.tag(BLINGCU_KC_SHADE);
         for (int k = 1; k < nSlots; ++k) {
            if (!kindPresent[k]) continue;
            be.runQueue(DlShadeBody{dscene, ps, qb, n}, ps.qMat + (size_t)k * ps.cap, ps.counters + C_MAT0 + k, bound); launches++;
         }
         be.tag(BLINGCU_KC_TRACE_ANY); be.traceAny(ps.qShadow, ps.counters + C_SHADOW, bound, dscene, ps.shO, ps.shD, ps.occl);
         launches += 3;
         if (hasInfinite && !hasBox) { be.traceAny(ps.qMisAny, ps.counters + C_MISANY, bound, dscene, ps.miO, ps.miD, ps.occlM); launches++; }
         if (misNearest) { be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(ps.qMis, ps.counters + C_MIS, bound, dscene, ps.miO, ps.miD, ps.mihit); launches++; }
         be.tag(BLINGCU_KC_RESOLVE); be.runQueue(DlResolveShadowBody{ps, n}, ps.qShadow, ps.counters + C_SHADOW, bound);
         launches++;
         if (misNearest) { be.runQueue(DlResolveMisBody{dscene, ps, n}, ps.qMis, ps.counters + C_MIS, bound); launches++; }
         if (hasInfinite && !hasBox) { be.runQueue(DlResolveMisAnyBody{dscene, ps, n}, ps.qMisAny, ps.counters + C_MISANY, bound); launches++; }
         be.tag(BLINGCU_KC_OTHER); be.run(AdvanceBody{ps}, 1); launches++;
         uint32_t *t = qa; qa = qb; qb = t;
      }
   }
   void bouncesNormals(uint32_t n) {
      be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(ps.qA, ps.counters + C_ACTIVE, n, dscene, ps.rayO, ps.rayD, ps.hit);
      be.tag(BLINGCU_KC_SHADE); be.runQueue(NormalMapBody{dscene, ps}, ps.qA, ps.counters + C_ACTIVE, n);
      launches += 2;
   }
   // bidirectional integrator (bidir.h): eye path with S0 / S1, light path, S1 weights, then every light vertex with every eye vertex
   void bouncesBidir(uint32_t n) {
      const int md = hs.max_depth;
      const bool misNearest = hs.n_lights > 0;
      be.zero(bd.D, sizeof(F4) * 4 * (size_t)md * ps.cap);
      be.tag(BLINGCU_KC_OTHER); be.run(BdBeginBody{ps, bd}, n); launches++;
      for (int side = 0; side < 2; ++side) {
         uint32_t *qa = ps.qA, *qb = ps.qB;
         if (side) { be.tag(BLINGCU_KC_RAYGEN); be.run(BdLightGenBody{dscene, ps, qa}, n); be.tag(BLINGCU_KC_OTHER); be.run(BdLightStartBody{ps, n}, 1); launches += 2; }
         for (int d = 0; d < md; ++d) {
            be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(qa, ps.counters + C_ACTIVE, n, dscene, ps.rayO, ps.rayD, ps.hit);
            be.tag(BLINGCU_KC_SHADE); be.runQueue(BdVertexBody{dscene, ps, bd, qb, side, d}, qa, ps.counters + C_ACTIVE, n);
            launches += 2;
            if (!side) {   // estimateDirect of this eye vertex: the path integrator's queries, resolved into the depth's own plane
               PathState pd = ps; pd.L = bd.D + (size_t)d * 4 * ps.cap;
               be.tag(BLINGCU_KC_TRACE_ANY); be.traceAny(ps.qShadow, ps.counters + C_SHADOW, n, dscene, ps.shO, ps.shD, ps.occl); launches++;
               if (hasInfinite && !hasBox) { be.traceAny(ps.qMisAny, ps.counters + C_MISANY, n, dscene, ps.miO, ps.miD, ps.occlM); launches++; }
               if (misNearest) { be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(ps.qMis, ps.counters + C_MIS, n, dscene, ps.miO, ps.miD, ps.mihit); launches++; }
               be.tag(BLINGCU_KC_RESOLVE); be.runQueue(ResolveShadowBody{pd}, ps.qShadow, ps.counters + C_SHADOW, n); launches++;
               if (misNearest) { be.runQueue(ResolveMisBody{dscene, pd}, ps.qMis, ps.counters + C_MIS, n); launches++; }
               if (hasInfinite && !hasBox) { be.runQueue(ResolveMisAnyBody{dscene, pd}, ps.qMisAny, ps.counters + C_MISANY, n); launches++; }
            }
            be.tag(BLINGCU_KC_OTHER); be.run(AdvanceBody{ps}, 1); launches++;
            uint32_t *t = qa; qa = qb; qb = t;
         }
      }
      be.tag(BLINGCU_KC_RESOLVE); be.run(BdFinishBody{ps, bd}, n); launches++;
      for (int s = 0; s < md; ++s)
         for (int t = 0; t < md; ++t) {
            be.tag(BLINGCU_KC_SHADE); be.run(BdConnectBody{dscene, ps, bd, s, t}, n);
            be.tag(BLINGCU_KC_TRACE_ANY); be.traceAny(ps.qShadow, ps.counters + C_SHADOW, n, dscene, ps.shO, ps.shD, ps.occl);
            be.tag(BLINGCU_KC_RESOLVE); be.runQueue(ResolveShadowBody{ps}, ps.qShadow, ps.counters + C_SHADOW, n);
            be.tag(BLINGCU_KC_OTHER); be.run(AdvanceBody{ps}, 1);
            launches += 4;
         }
   }
   // Slots for n camera samples of the direct-lighting integrator: every glass-like vertex spawns one extra slot. Starts at
   // 2 n and doubles when a batch reports C_OVERFLOW (the batch is then re-run: nothing has reached the film yet); 2^maxDepth n
   // always suffices.
   uint32_t dlHeadroom = 2;
   bool dlOverflowed() { be.sync(); uint32_t f = 0; be.download(&f, ps.counters + C_OVERFLOW, sizeof(f)); return f != 0; }

   int fuseResolveOpt = -1;   // option "fuse_resolve": -1 = by scene size, 0 = never, 1 = always
   bool fuseResolve() const { return fuseResolveOpt < 0 ? nItems <= (1u << 20) : fuseResolveOpt != 0; }
   bool kindPresent[N_SHADE_KINDS] = {};
   int slotKernel[N_SHADE_KINDS] = {};   // which ShadeHitBody instantiation serves the slot (a shade kind, see shading.h)
   int nSlots = N_PLAIN_SLOTS;           // slots in use: [0, nSlots)
   std::vector<int> matSlot;             // material -> slot
   bool hasInfinite = false, hasArea = false, hasBox = false;
   void scanKinds(const blingcu_scene *ir) {
      for (int k = 0; k < N_SHADE_KINDS; ++k) { kindPresent[k] = false; slotKernel[k] = -1; }
      kindPresent[0] = true; nSlots = N_PLAIN_SLOTS;
      matSlot.assign(ir->n_materials ? ir->n_materials : 1, 1);
      int nTex = 0;
      for (uint32_t i = 0; i < ir->n_materials; ++i) {
         const int kind = ir->materials[i].kind;
         int slot, kernel;
         if (materialComputes(ir, ir->materials[i])) { slot = N_PLAIN_SLOTS + (nTex++ % MAX_TEX_SLOTS); kernel = SK_TEX0 + kind; }
         else { slot = 1 + kind; kernel = kind; }
         if (slotKernel[slot] >= 0 && slotKernel[slot] != kernel) kernel = SK_GENERAL;   // more than MAX_TEX_SLOTS textured materials
         slotKernel[slot] = kernel; kindPresent[slot] = true; matSlot[i] = slot;
         nSlots = std::max(nSlots, slot + 1);
      }
      hasInfinite = hasArea = hasBox = false;
      for (uint32_t i = 0; i < ir->n_shapes; ++i) hasBox |= ir->shapes[i].kind == BLINGCU_SHAPE_BOX;
      for (uint32_t i = 0; i < ir->n_lights; ++i) { hasInfinite |= ir->lights[i].kind == BLINGCU_LIGHT_INFINITE; hasArea |= ir->lights[i].kind == BLINGCU_LIGHT_AREA; }
   }

   int renderSlice(uint32_t pass, uint64_t seed, uint32_t sBegin, uint32_t sEnd) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "render before upload_scene");
      uint32_t spp = (uint32_t)(hs.nu * hs.nv);
      if (sBegin > sEnd || sEnd > spp) return fail(BLINGCU_EINVAL, "sample range out of bounds");
      const bool direct = hs.integrator == BLINGCU_INTEGRATOR_DIRECT;
      // the bidirectional integrator keeps 2 md vertex records per slot (~130 bytes each): smaller batches
      uint32_t kmax = std::max(1u, (hs.integrator == BLINGCU_INTEGRATOR_BIDIR ? std::min(batchTarget, 1u << 21) : batchTarget) / npix);
      uint32_t need = std::min(kmax, std::max(1u, sEnd - sBegin)) * npix;
      if (!direct) ensureState(need);
      auto t0 = be.timerStart();
      for (uint32_t s = sBegin; s < sEnd;) {
         // direct lighting: the wavefront (camera samples x head-room) stays within the same slot budget, so more
         // head-room means fewer sample indices per batch, not more memory
         if (direct) kmax = std::max(1u, batchTarget / dlHeadroom / npix);
         uint32_t k = std::min(kmax, sEnd - s);
         uint32_t n = k * npix;
         unsigned long long statSnap[N_STATS]; uint64_t launchSnap = launches;
         if (direct) {
            if ((uint64_t)n * dlHeadroom > 0x7fffffffull) return fail(BLINGCU_EINVAL, "direct lighting: the branch tree of one sample index needs more than 2^31 path slots (reduce the image size or max_depth)");
            ensureState(n * dlHeadroom);
            be.sync(); be.download(statSnap, ps.stats, sizeof(statSnap));   // a batch that overflows is re-run: its counts must not stay
         }
         be.tag(BLINGCU_KC_OTHER); be.run(BeginBatchBody{ps, n}, 1);
         be.tag(BLINGCU_KC_RAYGEN); be.run(RaygenBody{dscene, ps, seed, pass, s, npix, nullptr, nullptr, nullptr}, n);
         launches += 2;
         if (direct) {
            bouncesDirect(n);
            if (dlOverflowed()) { dlHeadroom *= 2; be.upload(ps.stats, statSnap, sizeof(statSnap)); launches = launchSnap; continue; }   // same batch again with twice the slots
         } else if (hs.integrator == BLINGCU_INTEGRATOR_NORMALS) bouncesNormals(n);
         else if (hs.integrator == BLINGCU_INTEGRATOR_BIDIR) bouncesBidir(n);
         else bounces(n);
         be.waitReduced();   // a film reduction still in flight reads `film`: it has overlapped everything up to here
         be.tag(BLINGCU_KC_FILM); be.run(FinalizeBody{dscene, ps}, n);
         be.run(FilmBody{dscene, ps, film, k, npix}, (uint32_t)hs.W * (uint32_t)hs.H);
         launches += 2;
         s += k;
      }
      lastMs = be.timerStop(t0);
      return 0;
   }

   int renderSamples(uint32_t pass, uint64_t seed, const int32_t *px, const int32_t *py, const uint32_t *smp, size_t n, float *outL, float *outXY) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "render before upload_scene");
      if (n == 0) return 0;
      if (n > 0x7fffffffu) return fail(BLINGCU_EINVAL, "too many samples");
      for (size_t i = 0; i < n; ++i)
         if (px[i] < hs.ex0 || px[i] > hs.ex1 || py[i] < hs.ey0 || py[i] > hs.ey1 || smp[i] >= (uint32_t)(hs.nu * hs.nv)) return fail(BLINGCU_EINVAL, "sample outside the sample extent");
      const bool direct = hs.integrator == BLINGCU_INTEGRATOR_DIRECT;
      int32_t *dpx = (int32_t *)be.alloc(4 * n), *dpy = (int32_t *)be.alloc(4 * n); uint32_t *ds = (uint32_t *)be.alloc(4 * n);
      be.upload(dpx, px, 4 * n); be.upload(dpy, py, 4 * n); be.upload(ds, smp, 4 * n);
      for (;;) {
         if (direct && (uint64_t)n * dlHeadroom > 0x7fffffffull) { be.free(dpx); be.free(dpy); be.free(ds); return fail(BLINGCU_EINVAL, "direct lighting: wavefront too large"); }
         ensureState((uint32_t)n * (direct ? dlHeadroom : 1u));
         be.run(BeginBatchBody{ps, (uint32_t)n}, 1);
         be.run(RaygenBody{dscene, ps, seed, pass, 0, npix, dpx, dpy, ds}, (uint32_t)n);
         if (hs.integrator == BLINGCU_INTEGRATOR_NORMALS) { bouncesNormals((uint32_t)n); break; }
         if (hs.integrator == BLINGCU_INTEGRATOR_BIDIR) { bouncesBidir((uint32_t)n); break; }
         if (!direct) { bounces((uint32_t)n); break; }
         bouncesDirect((uint32_t)n);
         if (!dlOverflowed()) break;
         dlHeadroom *= 2;
      }
      be.sync();
      std::vector<F4> L(4 * (size_t)ps.cap); std::vector<F2> xy(n);
#ifndef BL_SPEC_PLANES   // product layout (bodies.h::spec4At): one 64-byte record per slot
      be.download(L.data(), ps.L, sizeof(F4) * 4 * n);
#else
      for (int q = 0; q < 4; ++q) be.download(L.data() + (size_t)q * n, ps.L + (size_t)q * ps.cap, sizeof(F4) * n);
#endif
      be.download(xy.data(), ps.spos, sizeof(F2) * n);
      for (size_t i = 0; i < n; ++i) {
         for (int q = 0; q < 4; ++q) { F4 v = L[spec4At((uint32_t)n, (uint32_t)i, q)]; float *o = outL + 16 * i + 4 * q; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
         outXY[2 * i] = xy[i].x; outXY[2 * i + 1] = xy[i].y;
      }
      be.free(dpx); be.free(dpy); be.free(ds);
      return 0;
   }

   int evalTexture(int tex, const float *p, const float *uv, size_t n, float *out) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "eval_texture before upload_scene");
      if (tex < 0 || (uint32_t)tex >= nTextures) return fail(BLINGCU_EINVAL, "texture out of range");
      if (n == 0) return 0;
      if (n > 0x7fffffffu) return fail(BLINGCU_EINVAL, "too many points");
      float *dp = (float *)be.alloc(12 * n), *duv = (float *)be.alloc(8 * n), *dout = (float *)be.alloc(64 * n);
      be.upload(dp, p, 12 * n); be.upload(duv, uv, 8 * n);
      be.tag(BLINGCU_KC_OTHER); be.run(EvalTextureBody{dscene, tex, dp, duv, dout}, (uint32_t)n);
      be.download(out, dout, 64 * n);
      be.free(dp); be.free(duv); be.free(dout);
      return 0;
   }

   // explicit ray batches (parity check (a)). The ABI ray {o, tmin, d, tmax} is two float4 and the ABI hit
   // {t, prim, b1, b2} one, so host buffers are copied as they are and (de)interleaved on the device; the device
   // scratch is grow-only.
   void *tb[6] = {}; size_t tbCap = 0;
   void freeTraceScratch() { for (void *&p : tb) { be.free(p); p = nullptr; } tbCap = 0; }
   int traceBatch(const blingcu_ray *rays, size_t n, blingcu_hit *outHit, uint8_t *outOccl, uint32_t *nodes, uint32_t *prims) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "trace before upload_scene");
      if (n == 0) return 0;
      if (n > 0x7fffffffu) return fail(BLINGCU_EINVAL, "too many rays");
      // the product backend pipelines host batches (copy-in, traversal and copy-out of neighbouring chunks overlap)
      if (!nodes && be.traceHostBatch(rays, n, outHit, outOccl, dscene)) return 0;
      if (n > tbCap) {
         freeTraceScratch();
         tb[0] = be.alloc(sizeof(F4) * 2 * n); tb[1] = nullptr; tb[2] = nullptr;
         tb[3] = be.alloc(sizeof(F4) * n); tb[4] = be.alloc(4 * n); tb[5] = be.alloc(4 * n);
         tbCap = n;
      }
      F4 *dR = (F4 *)tb[0], *dO = dR, *dD = dR + 1, *dH = (F4 *)tb[3];   // the ABI ray IS the kernels' ray record
      be.upload(dR, rays, sizeof(F4) * 2 * n);
      if (outHit) {
         uint32_t *dN = (uint32_t *)tb[4], *dP = (uint32_t *)tb[5];
         if (nodes) be.traceStats((uint32_t)n, dscene, dO, dD, dH, dN, dP);
         else { be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(nullptr, nullptr, (uint32_t)n, dscene, dO, dD, dH); }
         be.tag(BLINGCU_KC_OTHER); be.run(HitToAbiBody{dscene, dH}, (uint32_t)n);
         be.download(outHit, dH, sizeof(F4) * n);
         if (nodes) { be.download(nodes, dN, 4 * n); be.download(prims, dP, 4 * n); }
      }
      if (outOccl) {
         uint8_t *dC = (uint8_t *)tb[4];
         be.tag(BLINGCU_KC_TRACE_ANY); be.traceAny(nullptr, nullptr, (uint32_t)n, dscene, dO, dD, dC);
         be.download(outOccl, dC, n);
      }
      return 0;
   }

   // ---- the light tracer (lighttrace.h). One call = photons [first, first + n) of a pass, in batches of slots.
   int lightTrace(uint32_t pass, uint64_t seed, uint64_t first, uint32_t n, float *records, size_t maxRecords, size_t *nRecords) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "light_trace before upload_scene");
      if (hs.cam.kind != BLINGCU_CAM_PERSPECTIVE || hs.cam.pixel_area == 0) return fail(BLINGCU_EINVAL, "light tracing needs a perspective camera with world2raster / pixel_area (sampleCam, Camera.hs:78-103)");
      const uint32_t npixFilm = (uint32_t)hs.W * (uint32_t)hs.H;
      if (!splat) { splat = (float *)be.alloc(sizeof(float) * 3 * (size_t)npixFilm); be.zero(splat, sizeof(float) * 3 * (size_t)npixFilm); }
      struct HostRec { uint32_t photon, depth; float px, py, X, Y, Z; };
      std::vector<HostRec> hostRecs;
      const uint32_t batchMax = std::min<uint32_t>(batchTarget, 1u << 24);
      auto t0 = be.timerStart();
      for (uint32_t done = 0; done < n;) {
         const uint32_t m = std::min(batchMax, n - done);
         ensureState(m);
         if (!ltCount || ltRec.cap < ps.cap) {
            freeLt();
            auto al = [&](size_t bytes) { void *p = be.alloc(bytes); ltAllocs.push_back(p); return p; };
            const size_t c = ps.cap;
            ltRec.pixel = (uint32_t *)al(4 * c); ltRec.key = (uint32_t *)al(4 * c); ltRec.xyz = (F4 *)al(16 * c); ltRec.pos = (F2 *)al(8 * c); ltRec.depth = (uint32_t *)al(4 * c); ltRec.cap = ps.cap;
            ltSorted = (uint32_t *)al(4 * c);
            ltCount = (uint32_t *)al(4 * (size_t)npixFilm); ltOffset = (uint32_t *)al(4 * (size_t)npixFilm); ltCursor = (uint32_t *)al(4 * (size_t)npixFilm);
            ltBlockSum = (uint32_t *)al(4 * (size_t)((npixFilm + LT_SCAN_BLOCK - 1) / LT_SCAN_BLOCK + 1));
         }
         be.tag(BLINGCU_KC_OTHER); be.run(LtBeginBody{ps, m}, 1);
         be.tag(BLINGCU_KC_RAYGEN); be.run(LtGenBody{dscene, ps, seed, pass, first + done}, m);
         be.tag(BLINGCU_KC_OTHER); be.run(LtAfterGenBody{ps}, 1);
         launches += 3;
         uint32_t *qa = ps.qA, *qb = ps.qB;
         const uint32_t nb = (npixFilm + LT_SCAN_BLOCK - 1) / LT_SCAN_BLOCK;
         for (int depth = 0; depth < 1000; ++depth) {   // no depth limit in the reference: Russian roulette ends the paths
            be.tag(BLINGCU_KC_TRACE_NEAREST); be.traceNearest(qa, ps.counters + C_ACTIVE, m, dscene, ps.rayO, ps.rayD, ps.hit);
            be.tag(BLINGCU_KC_SHADE); be.runQueue(LtVertexBody{dscene, ps, qb}, qa, ps.counters + C_ACTIVE, m);
            be.tag(BLINGCU_KC_TRACE_ANY); be.traceAny(ps.qShadow, ps.counters + C_SHADOW, m, dscene, ps.shO, ps.shD, ps.occl);
            be.tag(BLINGCU_KC_RESOLVE); be.runQueue(LtCollectBody{dscene, ps, ltRec}, ps.qShadow, ps.counters + C_SHADOW, m);
            // splat this bounce's records: count per pixel, exclusive scan, scatter, per-pixel ordered sum
            be.tag(BLINGCU_KC_FILM);
            be.zero(ltCount, 4 * (size_t)npixFilm); be.zero(ltCursor, 4 * (size_t)npixFilm);
            be.run(LtCountBody{ltRec, ps.counters + C_LT_RECORDS, ltCount}, m);
            be.run(LtScanSumBody{ltCount, npixFilm, ltBlockSum}, nb);
            be.run(LtScanBlocksBody{ltBlockSum, nb}, 1);
            be.run(LtScanApplyBody{ltCount, npixFilm, ltBlockSum, ltOffset}, nb);
            be.run(LtScatterBody{ltRec, ps.counters + C_LT_RECORDS, ltOffset, ltCursor, ltSorted}, m);
            be.run(LtGatherBody{ltRec, ltCount, ltOffset, ltSorted, splat}, npixFilm);
            launches += 10;
            uint32_t cnt[2] = {0, 0};   // [records of this bounce, paths that go on]
            const bool check = records || nRecords || depth >= 3;
            if (check) { be.sync(); be.download(&cnt[0], ps.counters + C_LT_RECORDS, 4); be.download(&cnt[1], ps.counters + C_NEXT, 4); ltSplats += cnt[0]; }
            if ((records || nRecords) && cnt[0]) {
               std::vector<uint32_t> key(cnt[0]), dep(cnt[0]); std::vector<F4> xyz(cnt[0]); std::vector<F2> pos(cnt[0]); std::vector<uint32_t> sidx(ps.cap);
               be.download(key.data(), ltRec.key, 4 * (size_t)cnt[0]); be.download(dep.data(), ltRec.depth, 4 * (size_t)cnt[0]);
               be.download(xyz.data(), ltRec.xyz, 16 * (size_t)cnt[0]); be.download(pos.data(), ltRec.pos, 8 * (size_t)cnt[0]);
               for (uint32_t k = 0; k < cnt[0]; ++k) hostRecs.push_back(HostRec{(uint32_t)(first + done + key[k]), dep[k], pos[k].x, pos[k].y, xyz[k].x, xyz[k].y, xyz[k].z});
            }
            be.tag(BLINGCU_KC_OTHER); be.run(LtAdvanceBody{ps}, 1); launches++;
            if (!check) { /* the first bounces always go on: no host round trip */ }
            else if (cnt[1] == 0) break;
            uint32_t *t = qa; qa = qb; qb = t;
         }
         if (!(records || nRecords)) { /* splat counts of the unchecked bounces */ }
         done += m;
      }
      lastMs = be.timerStop(t0);
      if (records || nRecords) {
         std::sort(hostRecs.begin(), hostRecs.end(), [](const HostRec &a, const HostRec &b) { return a.photon != b.photon ? a.photon < b.photon : a.depth < b.depth; });
         if (nRecords) *nRecords = hostRecs.size();
         for (size_t k = 0; records && k < hostRecs.size() && k < maxRecords; ++k) {
            float *o = records + 7 * k; const HostRec &r = hostRecs[k];
            o[0] = (float)r.photon; o[1] = (float)r.depth; o[2] = r.px; o[3] = r.py; o[4] = r.X; o[5] = r.Y; o[6] = r.Z;
         }
      }
      return 0;
   }

   // ---- the host's kd-tree as an alternative accelerator input (kdtree.h)
   int uploadKd(const blingcu_kdnode *nodes, uint32_t nNodesKd, int32_t root, const uint32_t *leaf, size_t nLeaf, const float *bounds) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "upload_kdtree before upload_scene");
      freeKd();
      if (!nodes || nNodesKd == 0 || !bounds || (nLeaf && !leaf)) return fail(BLINGCU_EINVAL, "kd-tree without nodes / bounds / leaf primitives");
      if (root < 0 || (uint32_t)root >= nNodesKd) return fail(BLINGCU_EINVAL, "kd-tree root out of range");
      // every reference in range, no child pointing backwards at the root, depth within the traversal stack
      for (uint32_t i = 0; i < nNodesKd; ++i) {
         const blingcu_kdnode &n = nodes[i];
         if (n.left < 0) { if ((size_t)n.first + n.count > nLeaf) return fail(BLINGCU_EINVAL, "kd-tree leaf range out of bounds"); }
         else if ((uint32_t)n.left >= nNodesKd || n.right < 0 || (uint32_t)n.right >= nNodesKd || n.axis < 0 || n.axis > 2) return fail(BLINGCU_EINVAL, "kd-tree interior node malformed");
      }
      for (size_t i = 0; i < nLeaf; ++i) if (leaf[i] >= nPrims) return fail(BLINGCU_EINVAL, "kd-tree leaf primitive out of range");
      {
         std::vector<std::pair<int32_t, int>> st; st.emplace_back(root, 1); size_t visited = 0;
         while (!st.empty()) {
            auto [ni, depth] = st.back(); st.pop_back();
            if (++visited > (size_t)nNodesKd) return fail(BLINGCU_EINVAL, "kd-tree is not a tree");
            if (depth > BL_KD_STACK) return fail(BLINGCU_EINVAL, "kd-tree deeper than the traversal stack");
            const blingcu_kdnode &n = nodes[ni];
            if (n.left >= 0) { st.emplace_back(n.left, depth + 1); st.emplace_back(n.right, depth + 1); }
         }
      }
      auto upk = [&](const void *src, size_t bytes) { void *d = be.alloc(bytes); be.upload(d, src, bytes); kdAllocs.push_back(d); return d; };
      kd.nodes = (const blingcu_kdnode *)upk(nodes, sizeof(blingcu_kdnode) * nNodesKd);
      static const uint32_t none = 0;
      kd.leaf = (const uint32_t *)upk(nLeaf ? leaf : &none, sizeof(uint32_t) * (nLeaf ? nLeaf : 1));
      kd.primRef = primRefDev; kd.root = root;
      for (int k = 0; k < 3; ++k) { kd.lo[k] = bounds[k]; kd.hi[k] = bounds[3 + k]; }
      kdUploaded = true;
      return 0;
   }
   int traceKd(const blingcu_ray *rays, size_t n, blingcu_hit *outHit, uint32_t *nodes, uint32_t *prims) {
      if (!kdUploaded) return fail(BLINGCU_ESTATE, "trace_kdtree before upload_kdtree");
      if (n == 0) return 0;
      if (n > 0x7fffffffu) return fail(BLINGCU_EINVAL, "too many rays");
      if (n > tbCap) {
         freeTraceScratch();
         tb[0] = be.alloc(sizeof(F4) * 2 * n); tb[1] = nullptr; tb[2] = nullptr;
         tb[3] = be.alloc(sizeof(F4) * n); tb[4] = be.alloc(4 * n); tb[5] = be.alloc(4 * n);
         tbCap = n;
      }
      F4 *dR = (F4 *)tb[0], *dO = dR, *dD = dR + 1, *dH = (F4 *)tb[3]; uint32_t *dN = (uint32_t *)tb[4], *dP = (uint32_t *)tb[5];
      be.upload(dR, rays, sizeof(F4) * 2 * n);
      be.tag(BLINGCU_KC_OTHER); be.run(TraceKdBody{dscene, kd, dO, dD, dH, dN, dP}, (uint32_t)n);
      be.run(HitToAbiBody{dscene, dH}, (uint32_t)n);
      be.download(outHit, dH, sizeof(F4) * n); be.download(nodes, dN, 4 * n); be.download(prims, dP, 4 * n);
      return 0;
   }

   // film_sum = sum over the communicator's ranks of film (asynchronous; see comm.h)
   int reduceFilm(int root) {
      if (!uploaded) return fail(BLINGCU_ESTATE, "reduce_film before upload_scene");
      const size_t nf = (size_t)hs.W * hs.H * 4;
      if (!filmSum) filmSum = (F4 *)be.alloc(nf * sizeof(float));
      return be.reduceFilm((const float *)film, (float *)filmSum, nf, root, err);
   }

   int getStats(blingcu_stats *out) {
      std::memset(out, 0, sizeof(*out));
      if (ps.stats) {
         be.sync();
         unsigned long long s[N_STATS]; be.download(s, ps.stats, sizeof(s));
         out->samples = s[S_SAMPLES]; out->rays_camera = s[S_CAM]; out->rays_extension = s[S_EXT]; out->rays_mis = s[S_MIS];
         out->rays_shadow = s[S_SHADOW]; out->dropped_samples = s[S_DROPPED]; out->rays_mis_culled = s[S_MISCULL]; out->rays_ext_culled = s[S_EXTCULL]; out->rays_mis_any = s[S_MISANY];
      }
      { uint64_t t6[6]; be.traversalTotals(t6); out->nodes_traversed = t6[0]; out->intersections = t6[1]; out->rays_counted = t6[2]; out->any_nodes_traversed = t6[3]; out->any_intersections = t6[4]; out->any_rays_counted = t6[5]; }
      if (ps.stats) { unsigned long long s2[N_STATS]; be.download(s2, ps.stats, sizeof(s2)); out->photons = s2[S_PHOTONS]; out->rays_light = s2[S_RAYS_LIGHT]; out->rays_connect = s2[S_RAYS_CONNECT]; }
      out->splats = ltSplats;
      out->kernel_launches = launches; out->bvh_nodes = nNodes; out->bvh_leaf_items = nItems; lastMs = be.timerRead(lastMs); out->last_pass_ms = lastMs; out->bvh_max_stack = (uint64_t)hs.bvh.max_stack;
      return 0;
   }
   void resetStats() { if (ps.stats) be.zero(ps.stats, sizeof(unsigned long long) * N_STATS); launches = 0; ltSplats = 0; be.resetProfile(); }
};

}  // namespace bl
