// trace_warpq.cuh -- variant 2 of K2 trace_nearest / K3 trace_any: persistent warps with a WARP-LEVEL LEAF QUEUE.
// Replaces the per-ray recursion of KdTree.hs:210-246 like variant 1 (trace_kernels.cuh), same nodes, same child order,
// same primitive tests (bit-identical t / b1 / b2).
//
// Why: ncu on variant 1 showed ~21.5 of 32 lanes active per instruction, and the warp model of tools/travsim.cpp on the real
// cfg-5 ray streams says where the rest goes: a majority-vote step runs the node code with the ~13 lanes that stand at a
// leaf idle, and the leaf code with only ~13.6 lanes busy. Here a lane NEVER waits at a leaf:
//   * a lane that reaches a leaf appends (lane, item) pairs to a per-warp ring in shared memory and pops on;
//   * as soon as 32 pairs are queued ALL 32 lanes test one pair each -- the ray comes from its owner lane by shuffle, the
//     shrunken tmax (nearest hit) goes back through a shared-memory atomicMin on an order-preserving key, the hit record
//     through a per-lane shared-memory slot, an any-hit through a per-lane flag;
//   * a ray whose stack runs dry waits (WAITING) until its queued pairs are through, then retires.
// Model (profiles/r02_travsim.md): node steps 21.4 -> 28.2 lanes, leaf steps 13.6 -> 31.9 lanes, ~1.35x fewer warp
// instructions per ray. The nearest hit found is independent of the order of the tests except for exact t-ties, where the
// pair queued LATER wins, as in the sequential order of variant 1 (t == tmax is accepted, TriangleMesh.hs:180).
#pragma once

namespace bl {

#define TQ_CAP 128        // ring entries per warp: < 32 left over + at most 2 per lane per trip
#ifndef TQ_FLUSH
#define TQ_FLUSH 24       // run a leaf pass once this many pairs are queued (<= 32). Re-tuned on the round-2 tree (a third fewer pairs per
                          // ray, so a ray whose stack has run dry waits longer for a full pass): 32 -> 24 any-hit -1.1 %, nearest-hit
                          // unchanged; 16: +3 % / +2 % (tools/gpu_r02_z3.sh)
#endif
#ifndef TQ_NEAR_BLOCKS
#define TQ_NEAR_BLOCKS TR_MINBLOCKS   // resident CTAs per SM the nearest-hit instantiation is compiled for (A/B: 7 = 72 registers)
#endif
#define TQ_WARPS (TR_THREADS / 32)
#define TQ_STACK_OFF (TQ_CAP * 4 + 32 * 4 + 32 * 16 + 32 * 32 + 32 * 4)   // per warp: ring, per-lane key / flag, per-lane hit record, per-lane ray (o, tmin)(d, tmax), per-lane slot ...
#define TQ_WARP_BYTES (TQ_STACK_OFF + TR_SS * 128)                     // ... and the warp's own stack [level][lane]
// The stack lives in the WARP's region and the stack pointer is a shared-window ADDRESS (spA: the next free entry of my
// column), so that every push and pop is one instruction with an immediate offset and neither a stack base nor a level count
// has to stay in a register: ncu's source view of the previous layout ([level][thread] behind the four warps' heads) showed
// the stack base spilled to local memory and re-read twice per trip, and -- with an L1 hit rate of 3 % -- 10 % of all warp
// samples waiting on those LDLs (profiles/r02_trace_warpq.md). The ray's slot moved to shared memory for the same reason.
#define TQ_LEVEL(spA_, sqA_) ((int)(((spA_) - (sqA_) - (uint32_t)TQ_STACK_OFF) >> 7))   // stack level of an entry address (lane * 4 < 128)

// order-preserving float <-> uint32 key (so that atomicMin on the key is a float min for any sign)
__device__ __forceinline__ uint32_t fkey(float f) { const uint32_t b = __float_as_uint(f); return b ^ ((uint32_t)((int)b >> 31) | 0x80000000u); }
__device__ __forceinline__ float fkeyInv(uint32_t k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }

// shared memory through 32-bit shared-window addresses: generic pointers cost an address conversion (S2UR + ULEA + LEA) per access
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory"); }
__device__ __forceinline__ F4 lds128(uint32_t a) { F4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory"); return v; }
// L2 prefetch of data a LATER trip will load (TQ_PREFETCH): the kernels wait on memory (long-scoreboard stalls are half of all
// warp samples, L2 hit rate 45-60 % on an 830 MB tree), and both the items a lane queues and the children it pushes are
// certain to be read -- there is no culling on pop
#ifndef TQ_PREFETCH
#define TQ_PREFETCH 0     // bit 0: leaf items when they are queued, bit 1: inner children when they are pushed (both into L2);
                          // bit 2: the node of the NEXT trip into L1 as soon as it is known (after the pop), bit 3: both of its sectors
#endif
__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetchL1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void sminU32(uint32_t a, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// L[slot] += P[slot], one quarter at a time: the fused NEE resolve of the any-hit kernels (trace_kernels.cuh). It is compiled
// only into the FUSE instantiation: the any-hit kernels live at 56 registers (9 CTAs per SM) and this once-per-ray tail takes
// part in the hot loop's register allocation -- with the record layout of the spectra (bodies.h::spec4At) it pushed loop-carried
// values of the traversal into local memory (12 -> 44 bytes of spill stores), +18 % kernel time on cfg 5, which does not even
// take this path (tools/gpu_r02_f.sh). An out-of-line function cured the spills but cost the small scenes, which DO take it for
// every unoccluded shadow ray, up to a third of their speed (tools/gpu_r02_g.sh). The fused instantiation keeps the spills and
// is still the faster way there (cornell-box any-hit 4.54 -> 3.30 ms).
__device__ __forceinline__ void fuseAddPending(F4 *__restrict__ L, const F4 *__restrict__ P, uint32_t cap, uint32_t slot) {
   for (int qq = 0; qq < 4; ++qq) {
      const size_t at = spec4At(cap, slot, qq);
      F4 l = L[at]; const F4 p_ = P[at];
      l.x += p_.x; l.y += p_.y; l.z += p_.z; l.w += p_.w;
      L[at] = l;
   }
}

template <bool ANY, bool SORTED, bool STATS, bool FUSE>
__global__ void __launch_bounds__(TR_THREADS, (ANY && !FUSE) ? TR_MINBLOCKS + 1 : (ANY ? TR_MINBLOCKS : TQ_NEAR_BLOCKS)) kTraceWarpQ(const uint32_t *__restrict__ q, const uint32_t *__restrict__ cnt, uint32_t n,
                                                                 const DScene *__restrict__ sc, const F4 *__restrict__ O, const F4 *__restrict__ D,
                                                                 F4 *__restrict__ hit, uint8_t *__restrict__ occl, uint32_t *__restrict__ work, F4 *__restrict__ fuseL, const F4 *__restrict__ fuseP, uint32_t fuseCap,
                                                                 const Bvh bvh,     // the accelerator's pointers as a kernel PARAMETER: constant-bank operands instead of six registers
                                                                 unsigned long long *__restrict__ totals) {   // STATS: node visits, primitive tests, rays of THIS kernel (TraversalStats, KdTree.hs:252-258)
   extern __shared__ int sstack[];
   const unsigned FULL = 0xffffffffu;
   const uint32_t total = cnt ? *cnt : n;
   // lane id and the two shared-window addresses the hot loop uses come out of volatile asm so that they stay in registers: left
   // to itself the compiler rematerialises them at every use (S2R SR_TID.X, S2UR SR_CgaCtaId, ULEA ...: the S2R latency
   // sits on the critical path of every push, pop and queue append)
   unsigned lane; asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
   const unsigned ltMask = (1u << lane) - 1u;
   uint32_t sqA;
   {
      const uint32_t sBase = (uint32_t)__cvta_generic_to_shared(sstack);
      asm volatile("mov.u32 %0, %1;" : "=r"(sqA) : "r"(sBase + (threadIdx.x >> 5) * (uint32_t)TQ_WARP_BYTES));     // my warp's ring
   }
   const uint32_t sownW = sqA + TQ_CAP * 4u;        // [lane] of my warp: tmax key (nearest) / occluded flag (any)
   const uint32_t shitW = sownW + 32u * 4u;         // [lane] of my warp: best hit so far (nearest)
   const uint32_t srayW = shitW + 32u * 16u;        // [lane] of my warp: (o, tmin)(d, tmax) for the lanes that test my pairs
   const uint32_t sslotW = srayW + 32u * 32u;       // [lane] of my warp: the slot of the ray (needed again only when it retires)
   const uint32_t SP_FAST = (uint32_t)TQ_STACK_OFF + TR_SS * 128u;   // spA - sqA below this: the entry lies in shared memory
   int tail_[BL_STACK - TR_SS];
   const int EMPTY = (int)0x80000000, WAITING = (int)0x80000001;   // leaf references are > WAITING (bvh.h: ~((first << 4) | count))
   int cur = EMPTY, li = 0;
   uint32_t spA = sqA + (uint32_t)TQ_STACK_OFF + lane * 4u;
   uint32_t qhead = 0, qtail = 0, lastSeq = 0;   // absolute ring positions (warp-uniform) and the position behind my last pair
   bool occ = false;
   Ray r; RayPre pre;
   bool exhausted = false;
   r.o = mk3(0, 0, 0); r.d = mk3(0, 0, 1); r.tmin = 0; r.tmax = 0; pre.idir = mk3(0, 0, 0);
   unsigned long long stN = 0, stP = 0, stR = 0;   // STATS only

   for (;;) {
      unsigned mBusy = __ballot_sync(FULL, cur != EMPTY);
      // ---- refill idle lanes (warp-uniform decision)
      if (!exhausted && __popc(mBusy) <= 32 - TR_REFILL) {
         const unsigned idle = ~mBusy;
         uint32_t base = 0;
         const int leader = __ffs(idle) - 1;
         if ((int)lane == leader) base = atomicAdd(work, (uint32_t)__popc(idle));
         base = __shfl_sync(FULL, base, leader);
         if (cur == EMPTY) {
            const uint32_t k = base + __popc(idle & ltMask);
            if (k < total) {
               const uint32_t slot = q ? q[k] : k;
               sts32(sslotW + lane * 4u, slot);
               if (STATS) stR++;
               r = loadRay(O, D, slot);
               pre = rayPre(r);
               sts128(srayW + lane * 32u, r.o.x, r.o.y, r.o.z, r.tmin); sts128(srayW + lane * 32u + 16u, r.d.x, r.d.y, r.d.z, r.tmax);
               if (ANY) sts32(sownW + lane * 4u, 0u);
               else { sts32(sownW + lane * 4u, fkey(r.tmax)); sts128(shitW + lane * 16u, 0.0f, 0.0f, 0.0f, i2f(-1)); }
               spA = sqA + (uint32_t)TQ_STACK_OFF + lane * 4u; li = 0; occ = false; lastSeq = qtail;
               cur = (bvh.root >= 0) ? bvh.root : WAITING;   // empty scene: nothing to wait for either
            }
         }
         if (base + (uint32_t)__popc(idle) >= total) exhausted = true;
         mBusy = __ballot_sync(FULL, cur != EMPTY);
      }
      if (mBusy == 0) { if (exhausted) break; continue; }
      int pop = 0;
      // ---- node step for every lane that stands at an inner node
      if (cur >= 0) {
         if (STATS) stN++;
         const F4 *np = bvh.nodes + BL_NODE_F4 * (size_t)cur;
         F4 n0, n1, n2, n3;
         if (BL_L2_POLICY & 2) { ld8Policy<true>(np, n0, n1); ld8Policy<true>(np + 2, n2, n3); }
         else if (BL_L1_POLICY & 2) { ld8KeepL1(np, n0, n1); ld8KeepL1(np + 2, n2, n3); }
         else { ld8(np, n0, n1); ld8(np + 2, n2, n3); }
         float tn[4];
         node4Near<!SORTED>(n0, n2, n3, r, pre, tn);   // unsorted (any-hit): tn = minus the length of the ray inside the child
         int c[4] = {f2i(n1.x), f2i(n1.y), f2i(n1.z), f2i(n1.w)};
         if (SORTED) {
            sort4(tn, c);   // nearest first, misses (+inf) last
            const bool h0 = tn[0] < BL_INF, h1 = tn[1] < BL_INF, h2 = tn[2] < BL_INF, h3 = tn[3] < BL_INF;
            const int nh = (int)h0 + (int)h1 + (int)h2 + (int)h3;
            const uint32_t top = spA + (uint32_t)((nh > 0) ? nh - 1 : 0) * 128u;   // the new stack pointer; c[1] goes right below it
            if (top - sqA < SP_FAST + 128u) { stsIf(top - 128u, c[1], h1); stsIf(top - 256u, c[2], h2); stsIf(top - 384u, c[3], h3); }
            else {
               const int l1 = TQ_LEVEL(top, sqA) - 1, l2 = l1 - 1, l3 = l1 - 2;
               if (h1) { if (l1 < TR_SS) stsIf(top - 128u, c[1], true); else tail_[l1 - TR_SS] = c[1]; }
               if (h2) { if (l2 < TR_SS) stsIf(top - 256u, c[2], true); else tail_[l2 - TR_SS] = c[2]; }
               if (h3) { if (l3 < TR_SS) stsIf(top - 384u, c[3], true); else tail_[l3 - TR_SS] = c[3]; }
            }
            if (TQ_PREFETCH & 2) {
               if (h1 && c[1] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[1]);
               if (h2 && c[2] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[2]);
               if (h3 && c[3] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[3]);
            }
            spA = top;
            cur = c[0];
            pop = h0 ? 0 : 1;
         } else {
            // any-hit (variant 3): the answer does not depend on the order, so there is no sort; but WHICH child is entered first
            // decides how soon an occluder is found. The warp model (tools/travsim.cpp, profiles/r02_travsim.md) on the cfg-5
            // shadow / BSDF-MIS rays: entering the nearest child or the first in slot order costs the same 44.5 node visits per
            // ray, entering the child the ray stays in LONGEST costs 37.9 (-15 %): a long stretch inside a box of the soup is a
            // likely hit. The others are pushed as they come.
            const bool h0 = tn[0] <= 0.0f, h1 = tn[1] <= 0.0f, h2 = tn[2] <= 0.0f, h3 = tn[3] <= 0.0f;
            const bool b01 = tn[1] < tn[0], b23 = tn[3] < tn[2];
            const float k01 = b01 ? tn[1] : tn[0], k23 = b23 ? tn[3] : tn[2];
            const bool bb = k23 < k01;                                       // a miss (> 0) never beats a hit (<= 0)
            const bool p0 = h0 && (b01 || bb), p1 = h1 && (!b01 || bb), p2 = h2 && (b23 || !bb), p3 = h3 && (!b23 || !bb);
            const uint32_t a0 = spA, a1 = a0 + (p0 ? 128u : 0u), a2 = a1 + (p1 ? 128u : 0u), a3 = a2 + (p2 ? 128u : 0u);
            if (a3 - sqA < SP_FAST) { stsIf(a0, c[0], p0); stsIf(a1, c[1], p1); stsIf(a2, c[2], p2); stsIf(a3, c[3], p3); }
            else {
               const int l0 = TQ_LEVEL(a0, sqA), l1 = TQ_LEVEL(a1, sqA), l2 = TQ_LEVEL(a2, sqA), l3 = TQ_LEVEL(a3, sqA);
               if (p0) { if (l0 < TR_SS) stsIf(a0, c[0], true); else tail_[l0 - TR_SS] = c[0]; }
               if (p1) { if (l1 < TR_SS) stsIf(a1, c[1], true); else tail_[l1 - TR_SS] = c[1]; }
               if (p2) { if (l2 < TR_SS) stsIf(a2, c[2], true); else tail_[l2 - TR_SS] = c[2]; }
               if (p3) { if (l3 < TR_SS) stsIf(a3, c[3], true); else tail_[l3 - TR_SS] = c[3]; }
            }
            if (TQ_PREFETCH & 2) {
               if (p0 && c[0] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[0]);
               if (p1 && c[1] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[1]);
               if (p2 && c[2] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[2]);
               if (p3 && c[3] >= 0) prefetchL2(bvh.nodes + BL_NODE_F4 * (size_t)c[3]);
            }
            spA = a3 + (p3 ? 128u : 0u);
            cur = bb ? (b23 ? c[3] : c[2]) : (b01 ? c[1] : c[0]);
            pop = (h0 || h1 || h2 || h3) ? 0 : 1;
         }
         li = 0;
      }
      // ---- lanes that stand at a leaf (entered just now, or popped last trip) queue up to two of its items and move on
      {
         const bool atLeaf = cur < 0 && cur > WAITING && pop == 0;
         const int enc = ~cur;
         const int rem = atLeaf ? (enc & 15) - li : 0;
         const unsigned b1 = __ballot_sync(FULL, rem >= 1);
         if (b1) {
            const unsigned b2 = __ballot_sync(FULL, rem >= 2);
            const uint32_t at = qtail + (uint32_t)__popc(b1 & ltMask) + (uint32_t)__popc(b2 & ltMask);
            const uint32_t ent = (lane << 27) | (uint32_t)((enc >> 4) + li);
            if (rem >= 1) sts32(sqA + ((at & (TQ_CAP - 1)) << 2), ent);
            if (rem >= 2) sts32(sqA + (((at + 1u) & (TQ_CAP - 1)) << 2), ent + 1u);
            if (TQ_PREFETCH & 1) {
               const F4 *ip = bvh.items + (size_t)BL_ITEM_F4 * (size_t)((enc >> 4) + li);
               if (rem >= 1) prefetchL2(ip);
               if (rem >= 2) prefetchL2(ip + BL_ITEM_F4);
            }
            qtail += (uint32_t)__popc(b1) + (uint32_t)__popc(b2);
            if (rem >= 1) { li += 2; lastSeq = qtail; }
         }
         if (atLeaf && rem <= 2) pop = 1;   // everything of this leaf is queued (or it is empty)
      }
      // ---- pop (predicated load); with an empty stack the ray waits for its queued pairs
      {
         const bool some = spA - sqA >= (uint32_t)TQ_STACK_OFF + 128u;   // my column holds an entry
         const bool doPop = pop != 0 && some;
         if (pop != 0) { li = 0; if (!some) cur = WAITING; }
         spA -= doPop ? 128u : 0u;
         if (doPop && spA - sqA >= SP_FAST) cur = tail_[TQ_LEVEL(spA, sqA) - TR_SS];
         else cur = ldsIf(spA, cur, doPop);
      }
      if (TQ_PREFETCH & 4) {
         if (cur >= 0) {
            const F4 *np = bvh.nodes + BL_NODE_F4 * (size_t)cur;
            prefetchL1(np);
            if (TQ_PREFETCH & 8) prefetchL1(np + 2);
         }
      }
      // ---- leaf passes: 32 pairs at a time; a partial batch only when nobody could do anything else
      {
         uint32_t qn = qtail - qhead;
         if (qn >= (uint32_t)TQ_FLUSH || (qn > 0u && !__any_sync(FULL, cur > WAITING))) {
            __syncwarp();   // ring writes of this trip
            do {
               const uint32_t nb = qn < 32u ? qn : 32u;
               const uint32_t e = lds32(sqA + (((qhead + lane) & (TQ_CAP - 1)) << 2));
               const bool valid = lane < nb;
               if (STATS && valid) stP++;
               const int owner = (int)(e >> 27); const int item = (int)(e & 0x07ffffffu);
               // the owner's ray from its shared-memory copy: two LDS.128 (bank = owner lane: conflict-free, equal owners broadcast)
               // instead of eight shuffles, and the direction does not have to live in the owner's registers at all
               Ray rr;
               { const F4 a = lds128(srayW + (uint32_t)owner * 32u), b = lds128(srayW + (uint32_t)owner * 32u + 16u);
                 rr.o = mk3(a.x, a.y, a.z); rr.tmin = a.w; rr.d = mk3(b.x, b.y, b.z); rr.tmax = b.w; }
               if (ANY) {
                  bool found = false;
                  if (valid) found = leafItemAny(bvh, item, rr);
                  if (found) sts32(sownW + (uint32_t)owner * 4u, 1u);
                  __syncwarp();
                  if (lds32(sownW + lane * 4u) != 0u && cur != EMPTY && !occ) { occ = true; spA = sqA + (uint32_t)TQ_STACK_OFF + lane * 4u; cur = WAITING; }   // my ray is occluded: drop the rest of its walk
               } else {
                  rr.tmax = fkeyInv(lds32(sownW + (uint32_t)owner * 4u));
                  HitRec hh; hh.t = 0; hh.prim = -1; hh.b1 = hh.b2 = 0;
                  bool found = false;
                  if (valid) found = leafItemNearest(bvh, item, rr, hh);
                  const unsigned mh = __ballot_sync(FULL, found);
                  if (mh) {
                     const uint32_t key = fkey(hh.t);
                     if (found) sminU32(sownW + (uint32_t)owner * 4u, key);
                     __syncwarp();
                     bool win = found && lds32(sownW + (uint32_t)owner * 4u) == key;
                     const unsigned mw = __ballot_sync(FULL, win);
                     if (win) { const unsigned peers = __match_any_sync(mw, owner); win = (31 - __clz((int)peers)) == (int)lane; }   // equal t: the pair queued last
                     if (win) sts128(shitW + (uint32_t)owner * 16u, hh.t, hh.b1, hh.b2, i2f(hh.prim));
                     __syncwarp();
                     r.tmax = fkeyInv(lds32(sownW + lane * 4u));
                  }
               }
               qhead += nb; qn -= nb;
            } while (qn >= (uint32_t)TQ_FLUSH);
         }
      }
      // ---- retire rays whose walk is over and whose pairs are all through
      if (cur == WAITING && (int)(qhead - lastSeq) >= 0) {
         const uint32_t slot = lds32(sslotW + lane * 4u);
         if (ANY) {
            if (FUSE) { if (!occ) fuseAddPending(fuseL, fuseP, fuseCap, slot); }   // fused NEE resolve (trace_kernels.cuh): L += pending
            else occl[slot] = occ ? 1 : 0;
         } else hit[slot] = lds128(shitW + lane * 16u);
         cur = EMPTY;
      }
   }
   if (STATS) {
      for (int o = 16; o > 0; o >>= 1) { stN += __shfl_down_sync(FULL, stN, o); stP += __shfl_down_sync(FULL, stP, o); stR += __shfl_down_sync(FULL, stR, o); }
      if (lane == 0 && stR) { atomicAdd(totals, stN); atomicAdd(totals + 1, stP); atomicAdd(totals + 2, stR); }
   }
}

static inline size_t traceWarpQSmemBytes(int) { return (size_t)TQ_WARPS * TQ_WARP_BYTES; }   // the deep part of a stack (> TR_SS levels) lives in local memory

}  // namespace bl
