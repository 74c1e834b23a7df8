// ptb_trace.cuh -- closest hit through the flat BVH (BASELINE.json configs[4]: 100 k objects; every scene with meshes).
// Included by ptb_kernels.cuh.  Replaces the reference's brute-force loop (src/pathtrace.cu:176-199) for large scenes; a
// hierarchical structure is an assignment ask (INSTRUCTION.md:123-140).
//
// ONE kernel serves every caller (fused pipeline: camera rays / the CTA-owned path segments; staged pipeline; first-bounce
// cache; the ptb_stage_intersect test entry): `trace_kernel<SRC, ...>`, a persistent-threads kernel in which every lane
// MULTIPLEXES R rays.
//
// Why.  A ray through the 100 k-object scene needs ~20 traversal steps, ~1 leaf primitive to prune and ~0.3 reference-exact
// tests, and it finishes after a number of steps that varies by an order of magnitude.  Round 1 kept one ray per lane and
// gathered the lanes of a warp before each of the three kinds of work; ncu on that kernel (profiles/r02l / r2a_trace*):
// 74 % of the warp instructions were traversal steps executed with 11.6 of 32 lanes, because a lane whose ray sat at a leaf
// (or had finished) idled until enough others joined it.  Here a lane that cannot step its ray steps its OTHER ray:
//   * every lane owns R ray SLOTS in shared memory (origin, direction, hit-record index, best distance, pending leaf,
//     candidate, traversal stack); the traversal state of ONE of them (the active slot) is cached in registers;
//   * a slot is IDLE (wants a ray), STEP (can take a traversal step), LEAF (a leaf primitive waits for the pruning stage)
//     or CAND (a primitive that survived pruning waits for the reference-exact test);
//   * the warp runs, in a loop: ray fetch when kFetchLanes lanes have an idle slot, the exact test when kExactLanes lanes hold
//     a candidate, pruning when kPruneLanes lanes hold a leaf -- each also whenever nothing else can run -- and then kSteps
//     traversal steps in which a lane whose active slot is blocked switches to its other slot (~25 instructions).
// The per-lane traversal stack lives in SHARED memory (kStackSmem entries per slot, [entry][slot][thread]: conflict-free);
// deeper entries -- only trees deeper than kStackSmem levels can produce them -- go to a global spill area sized from the
// depth of the tree (no local-memory array: ptxas reported a 304-byte stack frame for the round-1 `int stack[64]`).
// The accepted hit goes straight to the global hit record (hits.tn / hits.meta of the ray's index): it is written ~0.3 times
// per ray, so it does not need to live in the slot; ties in t (src/pathtrace.cu:192: lowest scene index wins) and the
// same-mesh ordering rule read it back.
//
// Nothing here computes an output value except exactTestCore (ptb_math.cuh): traversal and pruning use padded boxes,
// approximate reciprocals and slackened bounds, and only decide WHICH primitives get the exact test.
#pragma once

namespace ptb {

#ifndef PTB_TRACE_BLOCK
#define PTB_TRACE_BLOCK 256
#endif
#ifndef PTB_TRACE_MINBLOCKS
#define PTB_TRACE_MINBLOCKS 3
#endif
#ifndef PTB_TRACE_RAYS
#define PTB_TRACE_RAYS 2               // R: ray slots per lane
#endif
#ifndef PTB_TRACE_STACK
#define PTB_TRACE_STACK 12             // traversal-stack entries per slot kept in shared memory
#endif
#ifndef PTB_TRACE_STEPS
#define PTB_TRACE_STEPS 4              // traversal steps between two looks at the other stages
#endif
#ifndef PTB_TRACE_FETCH_LANES
#define PTB_TRACE_FETCH_LANES 8
#endif
#ifndef PTB_TRACE_PRUNE_LANES
#define PTB_TRACE_PRUNE_LANES 12
#endif
#ifndef PTB_TRACE_EXACT_LANES
#define PTB_TRACE_EXACT_LANES 6
#endif
constexpr int kTraceBlock = PTB_TRACE_BLOCK;
constexpr int kTraceRays = PTB_TRACE_RAYS;
constexpr int kStackSmem = PTB_TRACE_STACK;
constexpr int kTraceSteps = PTB_TRACE_STEPS;
constexpr int kFetchLanes = PTB_TRACE_FETCH_LANES;
constexpr int kPruneLanes = PTB_TRACE_PRUNE_LANES;
constexpr int kExactLanes = PTB_TRACE_EXACT_LANES;
constexpr int kMaxVirtualSegs = 64;    // CTA-owned path segments one persistent CTA serves at most (host: grid >= vgrid / 64)

// where the rays of a launch come from and where their hit records go
enum TraceSource {
    SRC_CAMERA   = 0,   // fused pipeline, bounce 0: camera rays generated in registers; hit record index = path index
    SRC_SEGMENTS = 1,   // fused pipeline, bounce >= 1: the CTA-owned segments of the path buffers (bounce_kernel's layout)
    SRC_PATHS    = 2,   // staged pipeline: paths 0 .. counts[depth] of p.in (dead paths -- no-compaction mode -- get a miss)
    SRC_PIXELS   = 3,   // first-bounce cache: the iteration-independent camera ray of every pixel (src/pathtrace.cu:122-143)
    SRC_RAYS     = 4    // ptb_stage_intersect: caller-supplied origins / directions (3 floats each)
};
struct TraceIO {        // SRC_RAYS only
    const float *origins;
    const float *dirs;
    int          n;
};

// slot words in shared memory (all 32-bit): [word][slot][thread]
enum { W_OX, W_OY, W_OZ, W_DX, W_DY, W_DZ, W_IX, W_IY, W_IZ, W_DL, W_IDX, W_BEST, W_NODE, W_SP, W_CAND, kSlotWords };
// what a slot needs next.  The state is IMPLICIT in (cand, node, sp): cand >= 0 -> S_CAND; node < -1 (an encoded leaf) -> S_LEAF;
// node >= 0, or node == -1 (nothing current) with a non-empty stack -> S_STEP; node == -1 and an empty stack -> S_IDLE (the ray
// is finished: a miss was recorded when it was fetched, a hit when it was accepted).
enum { S_IDLE = 0, S_STEP = 1, S_LEAF = 2, S_CAND = 3 };
constexpr int kNoNode = -1;

constexpr size_t traceSmemBytes(int smemNodes) {
    return (size_t)smemNodes * sizeof(BvhNode) + (size_t)(kSlotWords + kStackSmem) * kTraceRays * kTraceBlock * sizeof(uint32_t);
}

template <int SRC, bool EXT, bool MESH>
__global__ void __launch_bounds__(kTraceBlock, PTB_TRACE_MINBLOCKS) trace_kernel(const __grid_constant__ Params p, const int depth, const TraceIO io) {
    constexpr int R = kTraceRays;
    extern __shared__ float4 smem4[];
    __shared__ uint32_t sNextTicket, sTotal;
    __shared__ uint32_t sPrefix[kMaxVirtualSegs + 1];
    const int smemNodes = p.hdr->smemNodes;
    if (threadIdx.x == 0) sNextTicket = 0u;

    // ---- this CTA's rays.  Segment / camera sources: the virtual CTAs (segments) v = blockIdx.x, + gridDim.x, ... of the fused
    // pipeline's grid, served back to back as one queue; flat sources: one contiguous range.
    const uint32_t vgrid = p.vgrid;
    uint32_t flatLo = 0;
    if (SRC == SRC_CAMERA || SRC == SRC_SEGMENTS) {
        const uint32_t nv = vgrid > blockIdx.x ? (vgrid - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
        uint32_t cnt = 0;
        if (threadIdx.x < nv) {
            const uint32_t v = blockIdx.x + threadIdx.x * gridDim.x;
            if (SRC == SRC_CAMERA) {
                const uint32_t total = p.numPixels * (uint32_t)p.ctrl->batch;
                const uint32_t tiles = (total + kBlock - 1) / kBlock;
                cnt = (tiles > v ? (tiles - v + vgrid - 1) / vgrid : 0u) * kBlock;
            } else cnt = p.segCounts[((depth + 1) & 1) * vgrid + v];
        }
        if (threadIdx.x < kMaxVirtualSegs) sPrefix[threadIdx.x + 1] = threadIdx.x < nv ? cnt : 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            sPrefix[0] = 0u;
            for (uint32_t j = 1; j <= kMaxVirtualSegs; ++j) { run += sPrefix[j]; sPrefix[j] = run; }
            sTotal = run;
        }
    } else {
        const uint32_t n = SRC == SRC_PATHS ? p.ctrl->counts[depth] : (SRC == SRC_PIXELS ? p.numPixels : (uint32_t)io.n);
        const uint32_t chunk = (((n + gridDim.x - 1) / gridDim.x) + 31u) & ~31u;
        flatLo = min(n, blockIdx.x * chunk);
        if (threadIdx.x == 0) sTotal = min(chunk, n - flatLo);
    }
    stageShared(smem4, reinterpret_cast<const float4 *>(p.bvh), smemNodes * 4);      // ends with __syncthreads
    const BvhNode *snodes = reinterpret_cast<const BvhNode *>(smem4);
    uint32_t *sw = reinterpret_cast<uint32_t *>(smem4 + (size_t)smemNodes * 4);
    uint32_t *stk = sw + kSlotWords * R * kTraceBlock;
    const uint32_t nTickets = sTotal;
#define SW(w, r) sw[((w) * R + (r)) * kTraceBlock + threadIdx.x]
#define SWF(w, r) __uint_as_float(SW(w, r))
#define STK(k, r) stk[((k) * R + (r)) * kTraceBlock + threadIdx.x]

    CameraParams cam;
    if (SRC == SRC_CAMERA || SRC == SRC_PIXELS) cam = p.hdr->cam;
    int iterBase = 0, iterStride = 0;
    if (SRC == SRC_CAMERA && EXT) { iterBase = p.ctrl->iterBase; iterStride = p.ctrl->iterStride; }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t laneBelow = (1u << lane) - 1u;
    int *spill = p.stackSpill ? p.stackSpill + ((size_t)blockIdx.x * kTraceBlock + threadIdx.x) * R * p.spillCap : nullptr;

    // ---- per-lane state: the ACTIVE slot's traversal state in registers; the category of every parked slot in `parked`
    // (2 bits per slot; a parked slot does not change until the lane switches back to it: every stage works on the active slot)
    f3 o = mk3(0.f, 0.f, 0.f), invD = o;
    float dl = 0.f, bestT = FLT_MAX;
    int node = kNoNode, sp = 0, cand = -1;
    int act = 0;
    uint32_t parked = 0u;                                            // all S_IDLE
    bool drained = false;                                            // warp-uniform: the CTA's queue is empty
    auto activeCat = [&]() -> uint32_t {
        return cand >= 0 ? (uint32_t)S_CAND : (node < kNoNode ? (uint32_t)S_LEAF : ((node >= 0 || sp > 0) ? (uint32_t)S_STEP : (uint32_t)S_IDLE));
    };
    auto has = [&](uint32_t s) -> bool {                             // some slot of this lane is in category s
        bool f = activeCat() == s;
        if (R > 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) f = f || (r != act && ((parked >> (2 * r)) & 3u) == s);
        }
        return f;
    };
    // makes a slot of category s the active one (the caller checked has(s)); ~25 instructions when it has to switch
    auto activate = [&](uint32_t s) {
        if (R == 1 || activeCat() == s) return;
        int r = 0;
#pragma unroll
        for (int k = R - 1; k >= 0; --k) if (k != act && ((parked >> (2 * k)) & 3u) == s) r = k;
        // park the active slot ...
        parked = (parked & ~(3u << (2 * act))) | (activeCat() << (2 * act));
        SW(W_NODE, act) = (uint32_t)node; SW(W_SP, act) = (uint32_t)sp; SW(W_CAND, act) = (uint32_t)cand; SW(W_BEST, act) = __float_as_uint(bestT);
        // ... and load slot r
        o = mk3(SWF(W_OX, r), SWF(W_OY, r), SWF(W_OZ, r));
        invD = mk3(SWF(W_IX, r), SWF(W_IY, r), SWF(W_IZ, r));
        dl = SWF(W_DL, r); bestT = SWF(W_BEST, r);
        node = (int)SW(W_NODE, r); sp = (int)SW(W_SP, r); cand = (int)SW(W_CAND, r);
        act = r;
    };

    for (;;) {
        const unsigned mStep = __ballot_sync(0xffffffffu, has(S_STEP));
        unsigned mLeaf = __ballot_sync(0xffffffffu, has(S_LEAF));
        const unsigned mCand = __ballot_sync(0xffffffffu, has(S_CAND));
        const unsigned mIdle = drained ? 0u : __ballot_sync(0xffffffffu, has(S_IDLE));
        if (!(mStep | mLeaf | mCand | mIdle)) break;

        // ---- fetch: one new ray for every lane with an idle slot
        if (mIdle && ((int)__popc(mIdle) >= kFetchLanes || !(mStep | mLeaf | mCand))) {
            const uint32_t want = (uint32_t)__popc(mIdle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&sNextTicket, want);
            base = __shfl_sync(0xffffffffu, base, 0);
            drained = base + want >= nTickets;
            if (has(S_IDLE)) {
                const uint32_t k = base + (uint32_t)__popc(mIdle & laneBelow);
                bool have = k < nTickets;
                uint32_t idx = 0;
                f3 ro = mk3(0.f, 0.f, 0.f), rd = ro;
                if (have) {
                    if (SRC == SRC_CAMERA || SRC == SRC_SEGMENTS) {
                        // ticket -> (virtual segment j, offset): the segments of this CTA back to back
                        uint32_t j = 0;
#pragma unroll 1
                        for (uint32_t step = kMaxVirtualSegs / 2; step > 0; step >>= 1) if (sPrefix[j + step] <= k) j += step;
                        const uint32_t off = k - sPrefix[j];
                        const uint32_t v = blockIdx.x + j * gridDim.x;
                        if (SRC == SRC_CAMERA) {
                            idx = (v + (off / kBlock) * vgrid) * kBlock + (off % kBlock);
                            have = idx < p.numPixels * (uint32_t)p.ctrl->batch;
                            if (have) {
                                const uint32_t slot = idx / p.numPixels;
                                cameraRayIter<EXT>(cam, idx - slot * p.numPixels, iterBase + (int)slot * iterStride, ro, rd);
                            }
                        } else {
                            idx = v * p.segCap + off;
                            const float4 a = p.in.A[idx];
                            const float2 b = *reinterpret_cast<const float2 *>(&p.in.B[idx]);
                            ro = mk3(a.x, a.y, a.z); rd = mk3(a.w, b.x, b.y);
                        }
                    } else {
                        idx = flatLo + k;
                        if (SRC == SRC_PATHS) {
                            const float4 a = p.in.A[idx];
                            const float2 b = *reinterpret_cast<const float2 *>(&p.in.B[idx]);
                            ro = mk3(a.x, a.y, a.z); rd = mk3(a.w, b.x, b.y);
                            if (__float_as_uint(p.in.C[idx].y) & kDeadBit) { p.hits.meta[idx] = -1; have = false; }   // dead path in place (no compaction)
                        } else if (SRC == SRC_PIXELS) cameraRay(cam, idx, ro, rd);
                        else {
                            ro = mk3(io.origins[3 * (size_t)idx], io.origins[3 * (size_t)idx + 1], io.origins[3 * (size_t)idx + 2]);
                            rd = mk3(io.dirs[3 * (size_t)idx], io.dirs[3 * (size_t)idx + 1], io.dirs[3 * (size_t)idx + 2]);
                        }
                    }
                }
                if (have) {
                    p.hits.meta[idx] = -1;                            // a miss until a hit is accepted
                    // the ray's direction enters the traversal only through 1 / d and |d| (slackened)
                    const float dd = fmaf(rd.x, rd.x, fmaf(rd.y, rd.y, rd.z * rd.z));
                    const float rdl = dd * rsqrtApprox(dd) * 0.9999f;
                    const f3 rinv = mk3(rcpApprox(rd.x), rcpApprox(rd.y), rcpApprox(rd.z));
                    int r = act;                                      // the active slot when it is the idle one, else the lowest idle parked slot
                    if (R > 1 && activeCat() != S_IDLE) {
#pragma unroll
                        for (int q = R - 1; q >= 0; --q) if (q != act && ((parked >> (2 * q)) & 3u) == S_IDLE) r = q;
                    }
                    SW(W_DX, r) = __float_as_uint(rd.x); SW(W_DY, r) = __float_as_uint(rd.y); SW(W_DZ, r) = __float_as_uint(rd.z);
                    SW(W_IDX, r) = idx;
                    if (R > 1) {
                        SW(W_OX, r) = __float_as_uint(ro.x); SW(W_OY, r) = __float_as_uint(ro.y); SW(W_OZ, r) = __float_as_uint(ro.z);
                        SW(W_IX, r) = __float_as_uint(rinv.x); SW(W_IY, r) = __float_as_uint(rinv.y); SW(W_IZ, r) = __float_as_uint(rinv.z);
                        SW(W_DL, r) = __float_as_uint(rdl);
                    }
                    if (R == 1 || r == act) {
                        o = ro; invD = rinv; dl = rdl; bestT = FLT_MAX; node = 0; sp = 0; cand = -1;
                    } else {
                        SW(W_BEST, r) = __float_as_uint(FLT_MAX); SW(W_NODE, r) = 0u; SW(W_SP, r) = 0u; SW(W_CAND, r) = 0xffffffffu;
                        parked = (parked & ~(3u << (2 * r))) | ((uint32_t)S_STEP << (2 * r));
                    }
                }
            }
        }

        // ---- the reference-exact test of one candidate per lane
        if (mCand && ((int)__popc(mCand) >= kExactLanes || !mStep)) {
            if (has(S_CAND)) {
                activate(S_CAND);
                const uint32_t idx = SW(W_IDX, act);
                const f3 rd = mk3(SWF(W_DX, act), SWF(W_DY, act), SWF(W_DZ, act));
                const float4 *gp = reinterpret_cast<const float4 *>(p.geoms + cand);
                Affine inv, fwd, vtx;
                inv.r0 = __ldg(gp); inv.r1 = __ldg(gp + 1); inv.r2 = __ldg(gp + 2);
                fwd.r0 = __ldg(gp + 3); fwd.r1 = __ldg(gp + 4); fwd.r2 = __ldg(gp + 5);
                const int4 tag = __ldg(reinterpret_cast<const int4 *>(gp + 9));    // x type, z scene order, w object index
                vtx.r0 = vtx.r1 = vtx.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MESH && tag.x == 2) { vtx.r0 = __ldg(gp + 6); vtx.r1 = __ldg(gp + 7); vtx.r2 = __ldg(gp + 8); }
                const f3 qo = mulMVPoint(inv, o), qu = mulMV(inv, rd, 0.0f);
                int outside = 1; float a = 0.f, b = 0.f, c = 0.f;
                const float t = exactTestCore<MESH>(tag.x, fwd, vtx, o, qo, qu, outside, a, b, c);
                // src/pathtrace.cu:192 + the index tie rule (+ the same-mesh rule): the current best is re-read from its hit record
                // only when it matters (equal t; a triangle against a best of any kind)
                bool take = t > 0.0f && bestT > t;
                if (t > 0.0f && bestT < FLT_MAX && (bestT == t || (MESH && tag.x == 2))) {
                    const int bm = p.hits.meta[idx];
                    const int bslot = bm & 0x3fffffff;
                    const int2 bord = __ldg(reinterpret_cast<const int2 *>(reinterpret_cast<const float4 *>(p.geoms + bslot) + 9) + 1);
                    const bool sameMesh = MESH && tag.x == 2 && __ldg(reinterpret_cast<const int *>(p.geoms + bslot) + 36) == 2 && bord.y == tag.w;
                    const float bestA = sameMesh ? p.hits.tn[idx].y : 0.0f;
                    take = takesOver(sameMesh, t, a, tag.z, bestT, bestA, bord.x);
                }
                if (take) {
                    p.hits.tn[idx] = make_float4(t, a, b, c);
                    p.hits.meta[idx] = cand | (outside << 30);
                    bestT = t;
                }
                cand = -1;                                            // back to S_LEAF (more primitives in the leaf), S_STEP or S_IDLE
            }
            mLeaf = __ballot_sync(0xffffffffu, has(S_LEAF));
        }

        // ---- pruning: one primitive of a pending leaf per lane -- object-space ray + sure-miss filter
        if (mLeaf && ((int)__popc(mLeaf) >= kPruneLanes || !__ballot_sync(0xffffffffu, has(S_STEP)))) {
            if (has(S_LEAF)) {
                activate(S_LEAF);
                const uint32_t code = (uint32_t)~node;                // ref | count << 27
                const int slot = (int)(code & 0x7ffffffu);
                const uint32_t cnt = code >> 27;
                node = cnt > 1u ? ~(int)((uint32_t)(slot + 1) | ((cnt - 1u) << 27)) : kNoNode;
                const float4 *gp = reinterpret_cast<const float4 *>(p.geoms + slot);
                const int type = __ldg(reinterpret_cast<const int *>(gp + 9));
                Affine inv;
                inv.r0 = __ldg(gp); inv.r1 = __ldg(gp + 1); inv.r2 = __ldg(gp + 2);
                const f3 rd = mk3(SWF(W_DX, act), SWF(W_DY, act), SWF(W_DZ, act));
                const f3 qo = mulMVPoint(inv, o), qu = mulMV(inv, rd, 0.0f);
                bool miss = boxSureMiss(qo, qu);
                if (type == 0) miss = sphereSureMiss(qo, qu);
                if (MESH && type == 2) miss = false;                  // triangles go straight to the exact test
                if (!miss) cand = slot;
            }
        }

        // ---- traversal steps: lanes whose active slot cannot step switch to one that can, once per pass
        if (R > 1 && has(S_STEP)) activate(S_STEP);
#pragma unroll 1
        for (int rep = 0; rep < kTraceSteps; ++rep) {
            // nothing current: take the top of the stack (a node, or a leaf that then waits for the pruning stage)
            if (cand < 0 && node == kNoNode && sp > 0) {
                --sp;
                node = sp < kStackSmem ? (int)STK(sp, act) : spill[(size_t)act * p.spillCap + (sp - kStackSmem)];
            }
            const bool stepping = cand < 0 && node >= 0;
            if (!__any_sync(0xffffffffu, stepping)) break;
            if (stepping) {
                int next, far; bool haveNext, haveFar;
                bvhNodeStep(snodes, smemNodes, p.bvh, node, o, invD, dl, bestT, next, haveNext, far, haveFar);
                if (haveFar) {
                    if (sp < kStackSmem) STK(sp, act) = (uint32_t)far;
                    else spill[(size_t)act * p.spillCap + (sp - kStackSmem)] = far;
                    ++sp;
                }
                node = haveNext ? next : kNoNode;                     // an interior node, an encoded leaf (< -1), or nothing
            }
        }
    }
#undef SW
#undef SWF
#undef STK
}

// Raw hit records (t, a, b, c | geom slot, outside) -> what the staged pipeline, the first-bounce cache and the stage entry
// consume: (t, normal) + meta, the normal built by the same deferred exactNormal the fused bounce kernel uses.
__global__ void __launch_bounds__(kBlock) resolve_hits_kernel(const __grid_constant__ Params p, const uint32_t *countPtr, uint32_t countImm,
                                                              float4 *tnOut, int *metaOut) {
    const uint32_t n = countPtr ? *countPtr : countImm;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        const int meta = p.hits.meta[i];
        if (meta < 0) { tnOut[i] = make_float4(-1.0f, 0.f, 0.f, 0.f); metaOut[i] = -1; continue; }   // src/pathtrace.cu:203
        const float4 r = p.hits.tn[i];
        const GeomRec g = loadGeom(p.geoms + (meta & 0x3fffffff));
        const f3 nrm = exactNormal(g, (meta >> 30) & 1, r.y, r.z, r.w);
        tnOut[i] = make_float4(r.x, nrm.x, nrm.y, nrm.z);
        metaOut[i] = meta;
    }
}

} // namespace ptb
