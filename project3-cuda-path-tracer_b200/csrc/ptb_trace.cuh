// ptb_trace.cuh -- closest hit through the flat BVH (BASELINE.json configs[4]: 100 k objects; every scene with meshes).
// Included by ptb_kernels.cuh.  Replaces the reference's brute-force loop (src/pathtrace.cu:176-199) for large scenes; a
// hierarchical structure is an assignment ask (INSTRUCTION.md:123-140).
//
// ONE kernel serves every caller (fused pipeline: camera rays / the CTA-owned path segments; staged pipeline; first-bounce
// cache; the ptb_stage_intersect test entry): `trace_kernel<SRC, ...>`, persistent threads with dynamic ray fetch and
// SPECULATIVE traversal.
//
// Why.  A bounce ray through the 100 k-object scene takes ~35 traversal steps, reaches ~1 leaf primitive that has to be pruned in
// object space and gets ~0.3 reference-exact tests; it finishes after a number of steps that varies by an order of magnitude.
// The three kinds of work have to be entered by many lanes of a warp together or the warp runs them with one or two lanes.
// Round 1 gathered the lanes before each kind: a lane whose ray stood at a leaf (or had finished) idled until enough others
// joined it, and ncu showed the traversal steps -- 3/4 of the instructions -- running with 12-14 of 32 lanes.  Giving every
// lane a second ray parked in shared memory raised that to 22 lanes but paid for it in switching and bookkeeping
// (profiles/r2_trace_design.txt).  What ships keeps ONE ray per lane, all of its traversal state in registers, and lets it run ahead:
//   * a leaf the traversal reaches goes into a two-entry leaf queue of the lane and the traversal CONTINUES with the next node
//     (speculative: the hit that leaf may hold cannot cull yet); the lane only stops when both entries are taken;
//   * the pruning stage (object-space ray + sure-miss filter of one queued primitive per lane) runs when kPruneLanes lanes
//     have something queued, the reference-exact test when kExactLanes lanes hold a survivor, the ray fetch when kFetchLanes
//     lanes are done -- each also whenever nothing else can run;
//   * the closest hit is order-independent ((t, scene index)-lexicographic minimum, src/pathtrace.cu:192), so testing
//     primitives later than the front-to-back order would changes nothing but the amount of culling.
// Stack entries carry the entry distance of their box: an entry that can no longer beat the best hit is dropped when it is
// popped, without fetching its node.  The stack lives in SHARED memory (kStackSmem 8-byte entries per lane, [entry][thread]:
// conflict-free); deeper entries -- only trees deeper than kStackSmem levels can produce them -- go to a global spill area
// sized from the depth of the tree (no local-memory array: ptxas reported a 304-byte frame for round 1's `int stack[64]`).
// Nodes are read through L1 (the kernel's small shared-memory footprint leaves it ~170 KB): child links are pre-encoded on
// the host (interior >= 0; leaf = ~(first primitive | count << 27); an absent child has a NaN box that no ray enters).
// The accepted hit goes straight to the global hit record (hits.tn / hits.meta of the ray's index): it is written ~0.3 times
// per ray; ties in t and the same-mesh ordering rule read it back.  A ray is recorded as a miss when it is fetched.
//
// Nothing here computes an output value except exactTestCore (ptb_math.cuh): traversal and pruning use padded boxes,
// approximate reciprocals and slackened bounds, and only decide WHICH primitives get the exact test.
#pragma once

namespace ptb {

#ifndef PTB_TRACE_BLOCK
#define PTB_TRACE_BLOCK 256
#endif
#ifndef PTB_TRACE_MINBLOCKS
#define PTB_TRACE_MINBLOCKS 4
#endif
#ifndef PTB_TRACE_STACK
#define PTB_TRACE_STACK 16             // traversal-stack entries per lane kept in shared memory
#endif
#ifndef PTB_TRACE_STEPS
#define PTB_TRACE_STEPS 5              // traversal steps between two looks at the other stages
#endif
#ifndef PTB_TRACE_QUEUE
#define PTB_TRACE_QUEUE 0              // leaf-queue entries per lane: 0 = the traversal stops at every leaf until it is pruned; 1, 2 = run ahead (speculative)
#endif
#ifndef PTB_TRACE_POP_PTX
#define PTB_TRACE_POP_PTX 1
#endif
#ifndef PTB_TRACE_POP_ATTEMPTS
#define PTB_TRACE_POP_ATTEMPTS 1       // stack entries a lane may pop per step while they turn out culled
#endif
#ifndef PTB_TRACE_FETCH_LANES
#define PTB_TRACE_FETCH_LANES 16     // (8 until the warps claimed their rays themselves: a fetch is one shuffle + address arithmetic now; profiles/r6d .. r6f)
#endif
#ifndef PTB_TRACE_PRUNE_LANES
#define PTB_TRACE_PRUNE_LANES 8
#endif
#ifndef PTB_TRACE_EXACT_LANES
#define PTB_TRACE_EXACT_LANES 3
#endif
#ifndef PTB_TRACE_DYNAMIC
#define PTB_TRACE_DYNAMIC 1            // 1: the warps of a SRC_SEGMENTS launch CLAIM pieces of the path segments (Control::traceNext) instead of their CTA owning segments
#endif
#ifndef PTB_TRACE_PIECE
#define PTB_TRACE_PIECE 128            // rays per claimed piece at most (a power of two >= 32; launches with few rays take smaller pieces)
#endif
constexpr int kTraceBlock = PTB_TRACE_BLOCK;
constexpr int kTraceRays = 1;          // ray slots per lane (spill-area sizing on the host)
constexpr int kStackSmem = PTB_TRACE_STACK;
constexpr int kTraceSteps = PTB_TRACE_STEPS;
constexpr int kLeafQueue = PTB_TRACE_QUEUE;
constexpr int kPopAttempts = PTB_TRACE_POP_ATTEMPTS;
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

constexpr int kNoNode = -1;            // "nothing": no current node / empty queue entry / no candidate
// per-lane words in shared memory besides the stack, all of them only touched by the fetch / pruning / exact stages: the ray
// direction, the index of the ray's hit record, the best distance so far, 1 / |d| (slackened), the lane's entries in the spill area
enum { W_DX, W_DY, W_DZ, W_IDX, W_BEST, W_RDL, W_SPILLED, kLaneWords };

constexpr size_t traceSmemBytes() { return (size_t)(kLaneWords + 2 * (kStackSmem + 2)) * kTraceBlock * sizeof(uint32_t); }     // + dummy and sentinel entries

// One traversal step: slab-test both child boxes of `node` (planes paired per axis, ptb_kernels.cuh slabNearPair), keep the
// children whose entry parameter can still beat the best hit (bestTs: best distance in units of the ray parameter, slackened).
// near / far: encoded children in front-to-back order, kNoNode when absent; tNear / tFar their entry parameters.
// A node is fetched with TWO 256-bit loads (sm_100: ld.global.nc.v8): divergent lanes cost the L1 data pipe one pass per lane and
// load instruction, and with three 128-bit loads + one 64-bit load per node that pipe ran at 82 % of its peak while the issue slots
// were 56 % busy (profiles/r2n_trace_ncu.csv).
#ifndef PTB_TRACE_LD256
#define PTB_TRACE_LD256 1
#endif
#ifndef PTB_TRACE_NODE_POLICY
#define PTB_TRACE_NODE_POLICY ""            // L1 eviction-priority qualifier of the node loads (tuning: ".L1::evict_last")
#endif
PTB_DEV void ldNode(const BvhNode *n, float4 &X, float4 &Y, float4 &Z, int2 &m) {
#if PTB_TRACE_LD256
    float4 M;
    asm volatile("ld.global.nc" PTB_TRACE_NODE_POLICY ".v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(X.x), "=f"(X.y), "=f"(X.z), "=f"(X.w), "=f"(Y.x), "=f"(Y.y), "=f"(Y.z), "=f"(Y.w) : "l"(n));
    asm volatile("ld.global.nc" PTB_TRACE_NODE_POLICY ".v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(Z.x), "=f"(Z.y), "=f"(Z.z), "=f"(Z.w), "=f"(M.x), "=f"(M.y), "=f"(M.z), "=f"(M.w) : "l"(reinterpret_cast<const float4 *>(n) + 2));
    m = make_int2(__float_as_int(M.x), __float_as_int(M.y));
#else
    const float4 *np = reinterpret_cast<const float4 *>(n);
    X = __ldg(np); Y = __ldg(np + 1); Z = __ldg(np + 2);
    m = __ldg(reinterpret_cast<const int2 *>(np + 3));
#endif
}

PTB_DEV void nodeStep(const BvhNode *nodes, int node, f3 o, f3 invD, float bestTs, int &near, float &tNear, int &far, float &tFar) {
    float4 X, Y, Z; int2 m;
    ldNode(nodes + node, X, Y, Z, m);
    float tl, tr; bool hl, hr;
    slabNearPair(X, Y, Z, o, invD, tl, tr, hl, hr);
    const bool vl = hl && tl <= bestTs;
    const bool vr = hr && tr <= bestTs;
    const bool leftFirst = tl <= tr;
    const int cl = vl ? m.x : kNoNode, cr = vr ? m.y : kNoNode;
    const bool swap = vr && !(vl && leftFirst);                  // the right child goes first: it is the only one, or the nearer one
    near = swap ? cr : cl; tNear = swap ? tr : tl;
    far = (vl && vr) ? (swap ? cl : cr) : kNoNode; tFar = swap ? tl : tr;
}

// Primitive records are read once per pruning / exact test; sending them past L1 (ld.global.nc.L1::no_allocate) so that the BVH
// nodes keep it measured 1 % SLOWER on the 100 k-object scene (profiles/r2i_variants.txt: 3 311 against 3 274 us): off.
#ifndef PTB_TRACE_GEOM_NC
#define PTB_TRACE_GEOM_NC 0
#endif
PTB_DEV float4 ldGeom4(const float4 *q) {
#if PTB_TRACE_GEOM_NC
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(q));
    return v;
#else
    return __ldg(q);
#endif
}

// The traversal stack of a lane: 8-byte entries (child, entry parameter of its box) in shared memory, [entry][thread]
// (conflict-free), addressed through a 32-bit shared-memory POINTER to the top entry that the lane keeps instead of a count.
// Entry 0 is a dummy and entry 1 a SENTINEL (child kEmpty, entry parameter -inf: never culled); popping the sentinel makes the
// lane's node kEmpty = "stack empty", so neither pop nor push tests a count or builds an address: a pop is one load, one
// subtract and two selects, a push one store and one add (as counted branches they were 45 of the ~110 warp instructions of a
// traversal step, run by 4-6 lanes: profiles/r2n_step_loop.txt).
// Entries 2 .. kStackSmem + 1 hold the real stack.  Only a tree deeper than that can overflow it; then the OLDEST kSpillBlock
// entries move to a global spill area (sized from the tree's depth, ptb200.cu::ensureStackSpill) and come back when the sentinel is
// reached -- checked once per pass of kTraceSteps steps, outside the step loop.
constexpr int kEmpty = (int)0x80000000;                             // as a leaf code: 15 primitives from slot 0x7ffffff -- the builder never makes it
constexpr int kSpillBlock = kStackSmem / 2;
constexpr uint32_t kStackStride = (uint32_t)kTraceBlock * 8u;       // bytes between consecutive entries of one lane
static_assert(kStackSmem - kTraceSteps >= kSpillBlock && kSpillBlock >= 1, "a spill must find kSpillBlock live entries");
PTB_DEV uint2 ldsEntry(uint32_t addr) {
    uint2 e;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
    return e;
}
PTB_DEV void stsEntry(uint32_t addr, uint2 e) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(e.x), "r"(e.y) : "memory");
}
PTB_DEV uint2 *spillArea(const Params &p) {
    return reinterpret_cast<uint2 *>(p.stackSpill) + ((size_t)blockIdx.x * kTraceBlock + threadIdx.x) * p.spillCap;
}

template <int SRC, bool EXT, bool MESH>
__global__ void __launch_bounds__(kTraceBlock, PTB_TRACE_MINBLOCKS) trace_kernel(const __grid_constant__ Params p, const int depth, const TraceIO io) {
    extern __shared__ float4 smem4[];
    __shared__ uint32_t sNextTicket, sTotal;
    __shared__ uint32_t sPrefix[kMaxVirtualSegs + 1];
    // DYN (bounces >= 1 of the fused pipeline): the path segments are cut into PIECES of 32 .. kPieceMax consecutive rays and every WARP claims
    // its next piece with one global atomic (Control::traceNext[depth]) when the one it holds is used up.  With segments owned by CTAs
    // (blockIdx, + grid, ...) a CTA's share of the launch was fixed: the SMs were busy for 2.18 - 2.42 M of the 2.49 M cycles of a launch
    // and ran 28.6 of their 32 warps while they were (ncu, profiles/r5z_trace_ncu.csv) -- now everybody finishes within one piece of
    // everybody else.  Piece i of the launch is rays [q * piece, (q + 1) * piece) of segment v, q = i / vgrid, v = i % vgrid, for
    // q < ceil(longest segment / piece): pieces beyond the end of a shorter segment are empty and skipped.
    constexpr bool DYN = PTB_TRACE_DYNAMIC && SRC == SRC_SEGMENTS;
    // Piece size: 32 << pieceShift rays, the largest power of two up to kPieceMax that still leaves every warp of the grid ~8 pieces of the
    // counts[depth] rays of this launch -- the late bounces of a small wavefront have fewer rays than the grid has lanes, and with a fixed 512
    // a 4-iteration wavefront ran 36 % slower than with CTA-owned segments (profiles/r5e_pieces_ab.txt).
    constexpr uint32_t kPieceMax = PTB_TRACE_PIECE;
    uint32_t pieceShift = 0;
    if (DYN) {
        const uint32_t perWarp = p.ctrl->counts[depth] / (gridDim.x * (kTraceBlock / 32) * 8u);
        while ((64u << pieceShift) <= kPieceMax && (64u << pieceShift) <= perWarp) ++pieceShift;
    }
    __shared__ uint32_t sPiece[DYN ? kTraceBlock / 32 : 1][2];       // per warp: index of the next ray of its piece in the path buffers, rays left in it
    if (threadIdx.x == 0) sNextTicket = 0u;

    // ---- this CTA's rays.  Segment / camera sources: the virtual CTAs (segments) v = blockIdx.x, + gridDim.x, ... of the fused
    // pipeline's grid, served back to back as one queue; flat sources: one contiguous range.
    const uint32_t vgrid = p.vgrid;
    uint32_t flatLo = 0;
    if (DYN) {
        // the longest segment of the launch (sTotal): every CTA reads the vgrid counts (L2-resident, a few KB)
        if (threadIdx.x == 0) sTotal = 0u;
        if (threadIdx.x < 2 * (kTraceBlock / 32)) (&sPiece[0][0])[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t *cnts = p.segCounts + ((depth + 1) & 1) * vgrid;
        uint32_t mx = 0;
        for (uint32_t v = threadIdx.x; v < vgrid; v += kTraceBlock) mx = max(mx, cnts[v]);
        for (int s = 16; s > 0; s >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        if ((threadIdx.x & 31u) == 0u && mx > 0u) atomicMax(&sTotal, mx);
    } else if (SRC == SRC_CAMERA || SRC == SRC_SEGMENTS) {
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
    __syncthreads();
    uint32_t *sw = reinterpret_cast<uint32_t *>(smem4);
    uint2 *stk = reinterpret_cast<uint2 *>(sw + kLaneWords * kTraceBlock);
    const uint32_t nTickets = DYN ? 0u : sTotal;                      // DYN: sTotal holds the longest segment, read where pieces are claimed
#define SW(w) sw[(w) * kTraceBlock + threadIdx.x]
#define SWF(w) __uint_as_float(SW(w))
    // shared-memory address of entry k of this lane's stack (only outside the step loop)
#define STK_ADDR(k) ((uint32_t)__cvta_generic_to_shared(stk + threadIdx.x) + (uint32_t)(k) * kStackStride)

    CameraParams cam;
    if (SRC == SRC_CAMERA || SRC == SRC_PIXELS) cam = p.hdr->cam;
    int iterBase = 0, iterStride = 0;
    if (SRC == SRC_CAMERA && EXT) { iterBase = p.ctrl->iterBase; iterStride = p.ctrl->iterStride; }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t laneBelow = (1u << lane) - 1u;
    const BvhNode *nodes = p.bvh;
    const bool canSpill = p.stackSpill != nullptr;                   // uniform: the tree is deeper than the shared-memory stack
    {
        const uint2 sentinel = make_uint2((uint32_t)kEmpty, 0xff800000u);     // -inf: the sentinel is never culled
        stsEntry(STK_ADDR(0), sentinel); stsEntry(STK_ADDR(1), sentinel);
        SW(W_SPILLED) = 0u;
    }

    // ---- per-lane state, all in registers
    f3 o = mk3(0.f, 0.f, 0.f), invD = o;
    float bestTs = FLT_MAX;                                          // best distance in units of the ray parameter, slackened: the culling bound
    int node = kEmpty;                                               // interior node (>= 0) | kNoNode: pop next | leaf code (< kNoNode): waits for
                                                                     // the pruning stage | kEmpty: the stack ran empty, the ray is finished
    uint32_t top = STK_ADDR(1);                                      // shared-memory address of the top stack entry
    int cand = kNoNode;                                              // primitive that survived pruning
    bool drained = false;                                            // warp-uniform: the CTA's queue is empty
    uint32_t segCur = 0;                                             // warp-uniform: virtual segment of the last fetch's first ticket

    for (;;) {
        // ---- the stack against its shared-memory window, once per pass (a pass pushes at most kTraceSteps entries per lane)
        if (canSpill) {
            const int used = (int)(top - STK_ADDR(1)) / (int)kStackStride;        // -1 once the sentinel has been popped
            if (used > kStackSmem - kTraceSteps) {
                // the oldest kSpillBlock entries go to the spill area, the rest move down
                uint2 *area = spillArea(p);
                const uint32_t spilled = SW(W_SPILLED);
                for (int k = 0; k < kSpillBlock; ++k) area[spilled + k] = ldsEntry(STK_ADDR(2 + k));
                for (int k = 2; k + kSpillBlock <= used + 1; ++k) stsEntry(STK_ADDR(k), ldsEntry(STK_ADDR(k + kSpillBlock)));
                SW(W_SPILLED) = spilled + kSpillBlock;
                top -= kSpillBlock * kStackStride;
            } else if (node == kEmpty && SW(W_SPILLED) > 0u) {
                // the sentinel was reached with entries still parked in the spill area: the newest block comes back
                const uint2 *area = spillArea(p);
                const uint32_t spilled = SW(W_SPILLED) - kSpillBlock;
                for (int k = 0; k < kSpillBlock; ++k) stsEntry(STK_ADDR(2 + k), area[spilled + k]);
                SW(W_SPILLED) = spilled;
                top = STK_ADDR(1 + kSpillBlock);
                node = kNoNode;
            }
        }
        const bool canStep = node >= kNoNode;
        const bool hasLeaf = node < kNoNode && node != kEmpty && cand == kNoNode;
        const bool hasCand = cand != kNoNode;
        const bool finished = node == kEmpty && cand == kNoNode;
        const unsigned mStep = __ballot_sync(0xffffffffu, canStep);
        unsigned mLeaf = __ballot_sync(0xffffffffu, hasLeaf);
        unsigned mCand = __ballot_sync(0xffffffffu, hasCand);
        const unsigned mIdle = drained ? 0u : __ballot_sync(0xffffffffu, finished);
        if (!(mStep | mLeaf | mCand | mIdle)) break;

        // ---- fetch: a new ray for every finished lane
        if (mIdle && ((int)__popc(mIdle) >= kFetchLanes || !(mStep | mLeaf | mCand))) {
            const uint32_t want = (uint32_t)__popc(mIdle);
            bool have = false;
            uint32_t idx = 0;
            f3 ro = mk3(0.f, 0.f, 0.f), rd = ro;
            if (DYN) {
                // the idle lanes take the next rays of the warp's piece, in lane order; a used-up piece is replaced by the next unclaimed
                // one of the launch
                const uint32_t warp = threadIdx.x >> 5, rank = (uint32_t)__popc(mIdle & laneBelow);
                uint32_t next = sPiece[warp][0], left = sPiece[warp][1], given = 0;
                const uint32_t piece = 32u << pieceShift;
                const uint32_t nPieces = ((sTotal + piece - 1u) >> (5u + pieceShift)) * vgrid;
                for (;;) {
                    if (left == 0u) {
                        uint32_t i = 0;
                        if (lane == 0) i = atomicAdd(&p.ctrl->traceNext[depth], 1u);
                        i = __shfl_sync(0xffffffffu, i, 0);
                        if (i >= nPieces) { drained = true; break; }
                        const uint32_t q = i / vgrid, v = i - q * vgrid;
                        const uint32_t cnt = p.segCounts[((depth + 1) & 1) * vgrid + v], lo = q * piece;
                        if (lo >= cnt) continue;                      // beyond the end of this segment
                        left = min(piece, cnt - lo); next = v * p.segCap + lo;
                    }
                    const uint32_t take = min(want - given, left);
                    if (finished && rank - given < take) { have = true; idx = next + (rank - given); }     // (unsigned: rank < given is served already)
                    next += take; left -= take; given += take;
                    if (given == want) break;
                }
                __syncwarp();
                if (lane == 0) { sPiece[warp][0] = next; sPiece[warp][1] = left; }
                __syncwarp();
                if (have) {
                    const float4 a = p.in.A[idx];
                    const float2 b = *reinterpret_cast<const float2 *>(&p.in.B[idx]);
                    ro = mk3(a.x, a.y, a.z); rd = mk3(a.w, b.x, b.y);
                }
            }
            uint32_t base = 0;
            if (!DYN) {
                if (lane == 0) base = atomicAdd(&sNextTicket, want);
                base = __shfl_sync(0xffffffffu, base, 0);
                drained = base + want >= nTickets;
            }
            if (!DYN && (SRC == SRC_CAMERA || SRC == SRC_SEGMENTS)) {
                // tickets only grow: the segment of the warp's first new ticket is found from the one of the previous fetch
                while (segCur + 1 < kMaxVirtualSegs && sPrefix[segCur + 1] <= base) ++segCur;
            }
            if (!DYN && finished) {
                const uint32_t k = base + (uint32_t)__popc(mIdle & laneBelow);
                have = k < nTickets;
                if (have) {
                    if (SRC == SRC_CAMERA || SRC == SRC_SEGMENTS) {
                        // ticket -> (virtual segment j, offset): the segments of this CTA back to back
                        uint32_t j = segCur;
                        while (sPrefix[j + 1] <= k) ++j;              // k < nTickets = sPrefix[kMaxVirtualSegs]
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
            }
            {
                if (have) {
                    p.hits.meta[idx] = -1;                            // a miss until a hit is accepted
                    SW(W_DX) = __float_as_uint(rd.x); SW(W_DY) = __float_as_uint(rd.y); SW(W_DZ) = __float_as_uint(rd.z);
                    SW(W_IDX) = idx;
                    // the direction enters the traversal only through 1 / d and |d| (slackened: ray parameter * |d| * 0.9999 is a
                    // lower bound of the world distance)
                    const float dd = fmaf(rd.x, rd.x, fmaf(rd.y, rd.y, rd.z * rd.z));
                    SW(W_RDL) = __float_as_uint(rsqrtApprox(dd) * 1.0002f);                 // 1 / (|d| * 0.9999)
                    o = ro; invD = mk3(rcpApprox(rd.x), rcpApprox(rd.y), rcpApprox(rd.z));
                    SW(W_BEST) = __float_as_uint(FLT_MAX); bestTs = FLT_MAX;
                    node = 0; top = STK_ADDR(1);
                }
            }
        }

        // ---- pruning: the next primitive of the leaf every waiting lane stands at -- object-space ray + sure-miss filter
        if (mLeaf && ((int)__popc(mLeaf) >= kPruneLanes || !mStep)) {
            if (hasLeaf) {
                const uint32_t code = (uint32_t)~node;                // first primitive | count << 27
                const int slot = (int)(code & 0x7ffffffu);
                const uint32_t cnt = code >> 27;
                node = cnt > 1u ? ~(int)((uint32_t)(slot + 1) | ((cnt - 1u) << 27)) : kNoNode;
                Affine inv;
                const float4 *gp = reinterpret_cast<const float4 *>(p.geoms + slot);
                const int type = __ldg(reinterpret_cast<const int *>(gp + 9));
                inv.r0 = ldGeom4(gp); inv.r1 = ldGeom4(gp + 1); inv.r2 = ldGeom4(gp + 2);
                const f3 rd = mk3(SWF(W_DX), SWF(W_DY), SWF(W_DZ));
                const f3 qo = mulMVPoint(inv, o), qu = mulMV(inv, rd, 0.0f);
                bool miss = boxSureMiss(qo, qu);
                if (type == 0) miss = sphereSureMiss(qo, qu);
                if (MESH && type == 2) miss = false;                  // triangles go straight to the exact test
                if (!miss) cand = slot;
            }
            mCand = __ballot_sync(0xffffffffu, cand != kNoNode);
        }

        // ---- the reference-exact test of every lane's candidate
        if (mCand && ((int)__popc(mCand) >= kExactLanes || !mStep)) {
            if (cand != kNoNode) {
                const uint32_t idx = SW(W_IDX);
                const float bestT = SWF(W_BEST);
                const f3 rd = mk3(SWF(W_DX), SWF(W_DY), SWF(W_DZ));
                const float4 *gp = reinterpret_cast<const float4 *>(p.geoms + cand);
                Affine inv, fwd, vtx;
                inv.r0 = ldGeom4(gp); inv.r1 = ldGeom4(gp + 1); inv.r2 = ldGeom4(gp + 2);
                fwd.r0 = ldGeom4(gp + 3); fwd.r1 = ldGeom4(gp + 4); fwd.r2 = ldGeom4(gp + 5);
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
                    SW(W_BEST) = __float_as_uint(t);
                    bestTs = (t + 1e-4f) * SWF(W_RDL);                // box entry parameter beyond which nothing can win any more
                }
                cand = kNoNode;
            }
        }

        // ---- traversal steps: every lane that can pops (when it has no node) and steps, in lockstep
#pragma unroll 1
        for (int rep = 0; rep < kTraceSteps; ++rep) {
            // nothing current: the top of the stack, unless its box can no longer beat the best hit (the sentinel: kEmpty)
#pragma unroll
            for (int attempt = 0; attempt < kPopAttempts; ++attempt) {
#if PTB_TRACE_POP_PTX
                // straight-line by construction (the compiler otherwise branches around the load): 5 instructions
                asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b32 ex, ey;\n\t.reg .f32 fy;\n\t"
                             "setp.eq.s32 p, %0, -1;\n\t"
                             "ld.shared.v2.u32 {ex, ey}, [%1];\n\t"
                             "@p sub.u32 %1, %1, %3;\n\t"
                             "mov.b32 fy, ey;\n\t"
                             "setp.le.and.f32 q, fy, %2, p;\n\t"
                             "@q mov.b32 %0, ex;\n\t}"
                             : "+r"(node), "+r"(top) : "f"(bestTs), "n"(kStackStride));
#else
                const bool doPop = node == kNoNode;
                const uint2 e = ldsEntry(top);
                top -= doPop ? kStackStride : 0u;
                node = (doPop && __uint_as_float(e.y) <= bestTs) ? (int)e.x : node;
#endif
            }
            if (!__any_sync(0xffffffffu, node >= kNoNode)) break;
            if (node >= 0) {
                int near, far; float tNear, tFar;
                nodeStep(nodes, node, o, invD, bestTs, near, tNear, far, tFar);
                if (far != kNoNode) {
                    top += kStackStride;
                    stsEntry(top, make_uint2((uint32_t)far, __float_as_uint(tFar)));
                }
                node = near;                                          // an interior node, a leaf (the pruning stage is next), or nothing
            }
        }
    }
#undef SW
#undef SWF
#undef STK_ADDR
}

// ------------------------------------------------------------------------------------------------
// trace_pool_kernel: the same search with a POOL of 64 ray slots per warp in shared memory (PTB_TRACE_POOL).
//
// In trace_kernel a lane keeps its ray until it is finished, so whenever the ray waits (for the pruning stage, the exact test, a
// new ray) the lane idles: the traversal step runs with ~19 of 32 lanes.  Here the 64 rays of a warp live in shared memory
// (origin, direction, hit-record index, best distance, current node / pending leaf, stack pointer, candidate, stack) and no lane
// owns any of them: every pass the warp looks at all 64 slots (two per lane, ballots), picks ONE kind of work -- kPoolSteps
// traversal steps while at least kPoolStepMin slots can step, else whichever of fetch / prune / exact has the most slots waiting
// -- and hands the first 32 slots of that kind to its lanes.  A waiting ray costs nothing but its slot, and prune / exact passes
// fill up on their own while the warp keeps stepping.  Per pass: ~90 instructions of classification, list building and state
// load / store against ~100 per traversal step.
// ------------------------------------------------------------------------------------------------
#ifndef PTB_POOL_STEPS
#define PTB_POOL_STEPS 8
#endif
#ifndef PTB_POOL_STEP_MIN
#define PTB_POOL_STEP_MIN 24
#endif
#ifndef PTB_POOL_STACK
#define PTB_POOL_STACK 10
#endif
#ifndef PTB_POOL_MINBLOCKS
#define PTB_POOL_MINBLOCKS 3
#endif
constexpr int kPoolSteps = PTB_POOL_STEPS;
constexpr int kPoolStepMin = PTB_POOL_STEP_MIN;
constexpr int kPoolStack = PTB_POOL_STACK;
constexpr int kPoolSlots = 2 * kTraceBlock;                       // per CTA: 64 per warp
enum { P_OX, P_OY, P_OZ, P_DX, P_DY, P_DZ, P_IDX, P_BEST, P_NODE, P_SP, P_CAND, kPoolWords };
enum { T_NONE = 0, T_STEP = 1, T_LEAF = 2, T_CAND = 3, T_IDLE = 4 };
constexpr size_t poolSmemBytes() { return (size_t)(kPoolWords + 2 * kPoolStack) * kPoolSlots * sizeof(uint32_t) + (size_t)kPoolSlots; }

template <int SRC, bool EXT, bool MESH>
__global__ void __launch_bounds__(kTraceBlock, PTB_POOL_MINBLOCKS) trace_pool_kernel(const __grid_constant__ Params p, const int depth, const TraceIO io) {
    extern __shared__ float4 smem4[];
    __shared__ uint32_t sNextTicket, sTotal;
    __shared__ uint32_t sPrefix[kMaxVirtualSegs + 1];
    if (threadIdx.x == 0) sNextTicket = 0u;
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
    __syncthreads();
    const uint32_t nTickets = sTotal;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t laneBelow = (1u << lane) - 1u;
    uint32_t *sw = reinterpret_cast<uint32_t *>(smem4);
    uint2 *stk = reinterpret_cast<uint2 *>(sw + kPoolWords * kPoolSlots);
    uint8_t *lst = reinterpret_cast<uint8_t *>(stk + kPoolStack * kPoolSlots) + warp * 64;
#define PW(w, s) sw[(w) * kPoolSlots + warp * 64 + (s)]
#define PWF(w, s) __uint_as_float(PW(w, s))
#define PSTK(k, s) stk[(k) * kPoolSlots + warp * 64 + (s)]
    CameraParams cam;
    if (SRC == SRC_CAMERA || SRC == SRC_PIXELS) cam = p.hdr->cam;
    int iterBase = 0, iterStride = 0;
    if (SRC == SRC_CAMERA && EXT) { iterBase = p.ctrl->iterBase; iterStride = p.ctrl->iterStride; }
    int2 *spillBase = p.stackSpill ? reinterpret_cast<int2 *>(p.stackSpill) + ((size_t)blockIdx.x * kPoolSlots + warp * 64) * p.spillCap : nullptr;
    const BvhNode *nodes = p.bvh;

    // all 64 slots of the warp start idle
    for (int j = 0; j < 2; ++j) { PW(P_NODE, lane + 32 * j) = (uint32_t)kNoNode; PW(P_SP, lane + 32 * j) = 0u; PW(P_CAND, lane + 32 * j) = (uint32_t)kNoNode; }
    __syncwarp();
    bool drained = false;

    for (;;) {
        // ---- what does every slot need?  (two slots per lane)
        uint32_t cat[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int nd = (int)PW(P_NODE, lane + 32 * j), spj = (int)PW(P_SP, lane + 32 * j), cd = (int)PW(P_CAND, lane + 32 * j);
            cat[j] = cd != kNoNode ? T_CAND : (nd < kNoNode ? T_LEAF : ((nd >= 0 || spj > 0) ? T_STEP : T_IDLE));
        }
        const unsigned bS0 = __ballot_sync(0xffffffffu, cat[0] == T_STEP), bS1 = __ballot_sync(0xffffffffu, cat[1] == T_STEP);
        const unsigned bL0 = __ballot_sync(0xffffffffu, cat[0] == T_LEAF), bL1 = __ballot_sync(0xffffffffu, cat[1] == T_LEAF);
        const unsigned bC0 = __ballot_sync(0xffffffffu, cat[0] == T_CAND), bC1 = __ballot_sync(0xffffffffu, cat[1] == T_CAND);
        const unsigned bI0 = __ballot_sync(0xffffffffu, cat[0] == T_IDLE), bI1 = __ballot_sync(0xffffffffu, cat[1] == T_IDLE);
        const int nStep = __popc(bS0) + __popc(bS1), nLeaf = __popc(bL0) + __popc(bL1), nCand = __popc(bC0) + __popc(bC1);
        const int nIdle = drained ? 0 : __popc(bI0) + __popc(bI1);
        // ---- one kind of work for this pass
        uint32_t type = T_NONE;
        unsigned m0 = 0u, m1 = 0u;
        if (nStep >= kPoolStepMin) type = T_STEP;
        else {
            int best = 0;
            if (nIdle > best) { best = nIdle; type = T_IDLE; }
            if (nLeaf > best) { best = nLeaf; type = T_LEAF; }
            if (2 * nCand > best) { best = 2 * nCand; type = T_CAND; }      // a survivor's test is what lets its ray cull: served earlier
            if (type == T_NONE && nStep > 0) type = T_STEP;
        }
        if (type == T_NONE) break;
        if (type == T_STEP) { m0 = bS0; m1 = bS1; } else if (type == T_LEAF) { m0 = bL0; m1 = bL1; }
        else if (type == T_CAND) { m0 = bC0; m1 = bC1; } else { m0 = bI0; m1 = bI1; }
        // ---- the first 32 slots of that kind, one per lane
        const int n0 = __popc(m0);
        if (cat[0] == type) { const int r = __popc(m0 & laneBelow); if (r < 32) lst[r] = (uint8_t)lane; }
        if (cat[1] == type) { const int r = n0 + __popc(m1 & laneBelow); if (r < 32) lst[r] = (uint8_t)(lane + 32); }
        __syncwarp();
        const int nItems = min(32, n0 + __popc(m1));
        const bool have = (int)lane < nItems;
        const int s = have ? (int)lst[lane] : 0;
        __syncwarp();

        if (type == T_STEP) {
            // the lanes step in lockstep (votes at every step): with per-lane `break` / `continue` the warp never reconverged inside
            // the pass and the node step ran with 11 of 32 lanes (profiles/r2l)
            f3 o = mk3(0.f, 0.f, 0.f), invD = o;
            float bestTs = 0.f;
            int node = kNoNode, sp = 0;
            int2 *spill = nullptr;
            if (have) {
                o = mk3(PWF(P_OX, s), PWF(P_OY, s), PWF(P_OZ, s));
                const f3 rd = mk3(PWF(P_DX, s), PWF(P_DY, s), PWF(P_DZ, s));
                const float dd = fmaf(rd.x, rd.x, fmaf(rd.y, rd.y, rd.z * rd.z));
                const float rdl = rsqrtApprox(dd) * 1.0002f;
                invD = mk3(rcpApprox(rd.x), rcpApprox(rd.y), rcpApprox(rd.z));
                const float bestT = PWF(P_BEST, s);
                bestTs = bestT == FLT_MAX ? FLT_MAX : (bestT + 1e-4f) * rdl;
                node = (int)PW(P_NODE, s); sp = (int)PW(P_SP, s);
                spill = spillBase ? spillBase + (size_t)s * p.spillCap : nullptr;
            }
#pragma unroll 1
            for (int rep = 0; rep < kPoolSteps; ++rep) {
                if (node == kNoNode && sp > 0) {
                    --sp;
                    const uint2 e = sp < kPoolStack ? PSTK(sp, s) : *reinterpret_cast<const uint2 *>(&spill[sp - kPoolStack]);
                    if (__uint_as_float(e.y) <= bestTs) node = (int)e.x;
                }
                const bool stepping = node >= 0;
                if (!__any_sync(0xffffffffu, stepping || (node == kNoNode && sp > 0))) break;
                if (stepping) {
                    int near, far; float tNear, tFar;
                    nodeStep(nodes, node, o, invD, bestTs, near, tNear, far, tFar);
                    if (far != kNoNode) {
                        const uint2 e = make_uint2((uint32_t)far, __float_as_uint(tFar));
                        if (sp < kPoolStack) PSTK(sp, s) = e;
                        else *reinterpret_cast<uint2 *>(&spill[sp - kPoolStack]) = e;
                        ++sp;
                    }
                    node = near;                                      // an interior node, a leaf (< -1: the pruning stage is next), or nothing
                }
            }
            if (have) { PW(P_NODE, s) = (uint32_t)node; PW(P_SP, s) = (uint32_t)sp; }
        } else if (type == T_LEAF) {
            if (have) {
                const int leaf = (int)PW(P_NODE, s);
                const uint32_t code = (uint32_t)~leaf;                // first primitive | count << 27
                const int slot = (int)(code & 0x7ffffffu);
                const uint32_t cnt = code >> 27;
                PW(P_NODE, s) = cnt > 1u ? (uint32_t)~(int)((uint32_t)(slot + 1) | ((cnt - 1u) << 27)) : (uint32_t)kNoNode;
                const float4 *gp = reinterpret_cast<const float4 *>(p.geoms + slot);
                const int gtype = __ldg(reinterpret_cast<const int *>(gp + 9));
                Affine inv;
                inv.r0 = ldGeom4(gp); inv.r1 = ldGeom4(gp + 1); inv.r2 = ldGeom4(gp + 2);
                const f3 o = mk3(PWF(P_OX, s), PWF(P_OY, s), PWF(P_OZ, s));
                const f3 rd = mk3(PWF(P_DX, s), PWF(P_DY, s), PWF(P_DZ, s));
                const f3 qo = mulMVPoint(inv, o), qu = mulMV(inv, rd, 0.0f);
                bool miss = boxSureMiss(qo, qu);
                if (gtype == 0) miss = sphereSureMiss(qo, qu);
                if (MESH && gtype == 2) miss = false;
                if (!miss) PW(P_CAND, s) = (uint32_t)slot;
            }
        } else if (type == T_CAND) {
            if (have) {
                const int cand = (int)PW(P_CAND, s);
                const uint32_t idx = PW(P_IDX, s);
                const f3 o = mk3(PWF(P_OX, s), PWF(P_OY, s), PWF(P_OZ, s));
                const f3 rd = mk3(PWF(P_DX, s), PWF(P_DY, s), PWF(P_DZ, s));
                const float bestT = PWF(P_BEST, s);
                const float4 *gp = reinterpret_cast<const float4 *>(p.geoms + cand);
                Affine inv, fwd, vtx;
                inv.r0 = ldGeom4(gp); inv.r1 = ldGeom4(gp + 1); inv.r2 = ldGeom4(gp + 2);
                fwd.r0 = ldGeom4(gp + 3); fwd.r1 = ldGeom4(gp + 4); fwd.r2 = ldGeom4(gp + 5);
                const int4 tag = __ldg(reinterpret_cast<const int4 *>(gp + 9));
                vtx.r0 = vtx.r1 = vtx.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MESH && tag.x == 2) { vtx.r0 = __ldg(gp + 6); vtx.r1 = __ldg(gp + 7); vtx.r2 = __ldg(gp + 8); }
                const f3 qo = mulMVPoint(inv, o), qu = mulMV(inv, rd, 0.0f);
                int outside = 1; float a = 0.f, b = 0.f, c = 0.f;
                const float t = exactTestCore<MESH>(tag.x, fwd, vtx, o, qo, qu, outside, a, b, c);
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
                    PW(P_BEST, s) = __float_as_uint(t);
                }
                PW(P_CAND, s) = (uint32_t)kNoNode;
            }
        } else {
            // ---- fetch: new rays into idle slots
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&sNextTicket, (uint32_t)nItems);
            base = __shfl_sync(0xffffffffu, base, 0);
            drained = base + (uint32_t)nItems >= nTickets;
            const uint32_t k = base + lane;
            bool got = have && k < nTickets;
            uint32_t idx = 0;
            f3 ro = mk3(0.f, 0.f, 0.f), rd = ro;
            if (got) {
                if (SRC == SRC_CAMERA || SRC == SRC_SEGMENTS) {
                    uint32_t j = 0;
#pragma unroll 1
                    for (uint32_t step = kMaxVirtualSegs / 2; step > 0; step >>= 1) if (sPrefix[j + step] <= k) j += step;
                    const uint32_t off = k - sPrefix[j];
                    const uint32_t v = blockIdx.x + j * gridDim.x;
                    if (SRC == SRC_CAMERA) {
                        idx = (v + (off / kBlock) * vgrid) * kBlock + (off % kBlock);
                        got = idx < p.numPixels * (uint32_t)p.ctrl->batch;
                        if (got) {
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
                        if (__float_as_uint(p.in.C[idx].y) & kDeadBit) { p.hits.meta[idx] = -1; got = false; }
                    } else if (SRC == SRC_PIXELS) cameraRay(cam, idx, ro, rd);
                    else {
                        ro = mk3(io.origins[3 * (size_t)idx], io.origins[3 * (size_t)idx + 1], io.origins[3 * (size_t)idx + 2]);
                        rd = mk3(io.dirs[3 * (size_t)idx], io.dirs[3 * (size_t)idx + 1], io.dirs[3 * (size_t)idx + 2]);
                    }
                }
            }
            if (got) {
                p.hits.meta[idx] = -1;                                // a miss until a hit is accepted
                PW(P_OX, s) = __float_as_uint(ro.x); PW(P_OY, s) = __float_as_uint(ro.y); PW(P_OZ, s) = __float_as_uint(ro.z);
                PW(P_DX, s) = __float_as_uint(rd.x); PW(P_DY, s) = __float_as_uint(rd.y); PW(P_DZ, s) = __float_as_uint(rd.z);
                PW(P_IDX, s) = idx; PW(P_BEST, s) = __float_as_uint(FLT_MAX);
                PW(P_NODE, s) = 0u; PW(P_SP, s) = 0u;
            }
        }
        __syncwarp();
    }
#undef PW
#undef PWF
#undef PSTK
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
