// raster_warp.cuh -- the small-scene raster kernel: a CTA owns a few scenes (14, 6 or 4), its warps share
// their geometry phase by phase and their raster work through one queue; no colour tile in shared memory.
//
// Target: CartPole-class scenes (a few instances of small flat-shaded meshes, <= 36 triangle slots).
// A scene costs ~2.7 k warp instructions and ~2 % of its pixels are covered, so the design goals are
// (a) as many resident scenes per SM as possible -- the per-scene shared-memory footprint is ~6.2 KB
// (records, masks, parked vertices), 28 scenes share an SM and the whole 4096-scene batch is one wave;
// (b) no wasted memory traffic: the background (or the pre-rendered static layer) goes straight to
// out[scene] -- by TMA bulk stores from one shared-memory copy per CTA, or as 128-bit stores by the warps --
// and only covered pixels are patched afterwards (the lines are still dirty in L2, DRAM sees each byte
// once); (c) balance: scenes differ a lot in raster work, so the non-empty 8x8 blocks of the CTA's scenes go
// into one queue that all warps of the CTA drain; (d) one launch per frame, frames chained by programmatic
// dependent launch: the next frame's CTAs start on an SM as this frame's leave it.
//
// Per CTA (TMA build): lane 0 of the last helper warp loads the background image into shared memory and,
// when its CTA's turn on this SM has come, issues one bulk store per scene; it waits for them ahead of the
// pre-sweep barrier.  Otherwise idle lanes build two tables that are the same for every scene of the frame:
// the block table (packed block coordinates, "the static layer covers part of it") and the slot table (what a
// triangle slot means: corner vertices, draw id, node / instance / triangle); one warp asks the L2 for the
// per-scene inputs of the CTA 32 places behind this one.
// Geometry, by the worker warps of the CTA over (scene, item) pairs of all its scenes, a named barrier
// between the phases:
//   M  instances  pose channels (the caller's state) -> model matrix, or the matrix buffer     (node.py:110-178)
//   A  vertices   lanes = (scene, instance, unique vertex): clip = VP*(M*v), outcodes, project + snap,
//                 parked in shared memory                                               (basic.vert:24-43)
//   B  classify   lanes = (scene, triangle slot): trivial reject / needs-clip / back-face cull from the
//                 parked vertices; survivors are compacted into the CTA's live list
//   S  set-up     two lanes per survivor in different warps: integer edge equations + depth plane -> 64-byte
//                 record, binned into per-8x8-block 64-bit masks | flat shade (basic.frag:31-38) -> its colour
// then per scene, by its own warp:
//   B3 clip       triangles crossing the near plane / guard band: Sutherland-Hodgman, fan triangles appended
//                 to the spare record slots, then to the per-SM overflow pool
//   Q  queue      the scene's non-empty blocks into the CTA's queue (one atomic per scene)
// Per CTA:
//   D  raster     warps pull (scene, block) items from the shared queue; every lane owns 2 pixels of the
//                 block and keeps their depth key (32 bits when draw order allows, else depth|id) and colour
//                 in registers across the block's records; winners are written straight to out[scene]
#pragma once
#include "common.cuh"

namespace pbr {

constexpr int W_MAXREC = 48;     // records per scene (triangle slots that survive + clipped fans); < 64 (mask bits)
constexpr int W_MAXSLOT = 36;    // eligibility: leaves >= 12 spare records for clipped fans
constexpr int W_MAXVERT = 48;    // (instance, vertex) pairs per scene (< 256: the live list names a corner in a byte)
constexpr int W_MAXINST = 16;    // instances per scene (their 3x3 model matrices are parked in shared memory)
constexpr int W_GEOM_BYTES = W_MAXVERT * 32 + W_MAXINST * 64;    // parked vertices (clip + projected) and matrices
constexpr int W_MW = 2;          // mask words per block (64 record bits)
#ifndef W_WARPS
#define W_WARPS 4                // scenes (= warps) per CTA
#endif

// Record overflow.  A triangle clipped against the near plane and the guard band becomes a fan of up
// to 6 triangles, so a scene of S <= W_MAXSLOT slots can need up to 6 S records; shared memory holds
// W_MAXREC.  The rest goes to a pool entry in global memory (records + per-block masks) that the warp
// claims on its SM for the rest of the kernel: at most 32 warps of this kernel are resident per SM
// (64 registers x 1024 threads), so 32 entries per SM always suffice and the frame stays exact.
constexpr int W_OVF_MW = 6;                       // mask words per block for pool records
constexpr int W_OVF_MAXREC = 32 * W_OVF_MW;       // 192 >= 6 * W_MAXSLOT - W_MAXREC
constexpr int W_POOL_PER_SM = 32;
static_assert(6 * W_MAXSLOT - W_MAXREC <= W_OVF_MAXREC, "overflow pool too small for the worst case");

// shared memory of one scene
__host__ __device__ inline size_t warp_scene_bytes(int nblk) {
    return (size_t)W_GEOM_BYTES + (size_t)W_MAXREC * sizeof(Rec) + align16((size_t)nblk * W_MW * 4) +
           align16((size_t)nblk * 2) + W_MAXREC * 4 + 32;
}
// offset of the CTA's counters: behind the scene regions, the block queue, the block table, the slot table and the live list
__host__ __device__ inline size_t warp_qctr_offset(int nblk, int warps) {
    return (size_t)warps * warp_scene_bytes(nblk) + align16((size_t)warps * nblk * 4) + align16((size_t)nblk * 4) +
           align16((size_t)W_MAXSLOT * 12) + align16((size_t)warps * W_MAXSLOT * 12);
}
// shared memory of a CTA of `warps` scenes: scene regions + block queue + counters (+ mbarrier and the
// background image when the background is written by TMA)
__host__ __device__ inline size_t warp_smem_bytes(int nblk, int warps, size_t tma_tile_bytes = 0) {
    // (eight counter words -- [4], [5] are the mbarrier of the TMA build, [6], [7] are used by every build -- then the image)
    return warp_qctr_offset(nblk, warps) + 32 + (tma_tile_bytes ? align16(tma_tile_bytes) : 0);
}
#ifndef PBR_W_WARPS_TMA
#define PBR_W_WARPS_TMA 14
#endif
constexpr int W_WARPS_TMA = PBR_W_WARPS_TMA;      // scenes per CTA when the background goes through TMA (see kernel comment)
constexpr int W_WARPS_TMA_SMALL = 6;              // ... for tiles whose image leaves room for only one CTA of W_WARPS_TMA per SM
__host__ __device__ constexpr int w_helpers(int warps);                // helper warps (no scene of their own) of a TMA CTA with `warps` scenes

// views into one scene's shared-memory region (layout: warp_scene_bytes)
struct WScene {
    float4 *clipc;            // [W_MAXVERT] clip-space positions
    int4 *proj;               // [W_MAXVERT] snapped x, y, depth bits, flags
    float4 *minst;            // [W_MAXINST][4] columns of the model matrices
    Rec *recs;                // [W_MAXREC]
    unsigned *masks;          // [nblk][W_MW]
    unsigned short *blist;    // [nblk]
    unsigned *clipl;          // [W_MAXREC] packed slots of the triangles that need clipping
    int *ctr;                 // [4] overflow pool entry, clipped triangles, records drawn (S > 32), has int64 records
    unsigned char **out_slot; // out[scene], for the sweep
};
// (the parts whose size depends on the tile -- block masks, block list -- come last: everything else sits at a
// constant offset from the region's base)
constexpr int W_OFF_RECS = W_GEOM_BYTES;
constexpr int W_OFF_CLIPL = W_OFF_RECS + W_MAXREC * (int)sizeof(Rec);
constexpr int W_OFF_CTR = W_OFF_CLIPL + W_MAXREC * 4;
constexpr int W_OFF_OUT = W_OFF_CTR + 24;
constexpr int W_OFF_MASKS = W_OFF_CTR + 32;
__device__ __forceinline__ WScene wscene(unsigned char *base, int nblk) {
    WScene s;
    s.clipc = reinterpret_cast<float4 *>(base);
    s.proj = reinterpret_cast<int4 *>(base + (size_t)W_MAXVERT * 16);
    s.minst = reinterpret_cast<float4 *>(base + (size_t)W_MAXVERT * 32);
    s.recs = reinterpret_cast<Rec *>(base + W_OFF_RECS);
    s.clipl = reinterpret_cast<unsigned *>(base + W_OFF_CLIPL);
    s.ctr = reinterpret_cast<int *>(base + W_OFF_CTR);
    s.out_slot = reinterpret_cast<unsigned char **>(base + W_OFF_OUT);
    s.masks = reinterpret_cast<unsigned *>(base + W_OFF_MASKS);
    s.blist = reinterpret_cast<unsigned short *>(base + W_OFF_MASKS + align16((size_t)nblk * W_MW * 4));
    return s;
}


// named barrier 1 over the first THREADS threads of the CTA (the warps that share the geometry work)
template <int THREADS_>
__device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS_) : "memory");
}

struct WSlot {
    int ni, inst, tri;
};

__device__ __forceinline__ unsigned pack_slot(int ni, int inst, int tri) {
    return ((unsigned)ni << 26) | ((unsigned)inst << 13) | (unsigned)tri;
}
__device__ __forceinline__ WSlot unpack_slot(unsigned p) {
    WSlot s;
    s.ni = (int)(p >> 26); s.inst = (int)((p >> 13) & 8191u); s.tri = (int)(p & 8191u);
    return s;
}

__device__ __forceinline__ void load_mat(const float *m, float *M) {
    const float4 *m4 = reinterpret_cast<const float4 *>(m);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 a = __ldg(m4 + j);
        M[4 * j] = a.x; M[4 * j + 1] = a.y; M[4 * j + 2] = a.z; M[4 * j + 3] = a.w;
    }
}

// xform_normal with the matrix given as its columns (parked in shared memory by phase M)
__device__ __forceinline__ void xform_normal_cols(const float4 *c, float nx, float ny, float nz, float *r) {
    const float4 c0 = c[0], c1 = c[1], c2 = c[2];
    const float m[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
    xform_normal(m, nx, ny, nz, r);
}

// vertex flags
constexpr int VF_CLIP = 0x40, VF_PROJ = 0x80;

// fill [dst, dst+n) with byte value v: 128-bit stores on the aligned body
__device__ __forceinline__ void fill_bytes(unsigned char *dst, int n, unsigned v, int lane) {
    const unsigned v4 = v * 0x01010101u;
    const int head = min(n, (int)((16 - (reinterpret_cast<size_t>(dst) & 15)) & 15));
    for (int i = lane; i < head; i += 32) dst[i] = (unsigned char)v;
    const int n16 = (n - head) / 16;
    uint4 *p = reinterpret_cast<uint4 *>(dst + head);
    const uint4 q = make_uint4(v4, v4, v4, v4);
    // asm volatile keeps the 512-byte warp stores in ascending address order
    for (int i = lane; i < n16; i += 32)
        asm volatile("st.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p + i), "r"(v4) : "memory");
    for (int i = head + n16 * 16 + lane; i < n; i += 32) dst[i] = (unsigned char)v;
}

// background (or the pre-rendered static layer) of one scene, straight to global memory
__device__ __forceinline__ void write_background(const FrameDev &f, unsigned char *out_scene, int HW, int lane) {
    if (f.base_color != nullptr) {
        const int n = f.C * HW;
        if ((n & 15) == 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(f.base_color);
            uint4 *dst = reinterpret_cast<uint4 *>(out_scene);
            // 8 independent 128-bit loads in flight per lane, then 8 stores (a plain copy loop
            // serialises on the L2 latency of every load)
            const int n16 = n / 16;
            int i = lane;
            for (; i + 7 * 32 < n16; i += 8 * 32) {
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(src + i + k * 32);
#pragma unroll
                for (int k = 0; k < 8; ++k) dst[i + k * 32] = v[k];
            }
            for (; i < n16; i += 32) dst[i] = __ldg(src + i);
        } else {
            for (int i = lane; i < n; i += 32) out_scene[i] = __ldg(f.base_color + i);
        }
    } else if (((f.bg ^ (f.bg >> 8)) & (f.C == 4 ? 0xffffffu : 0xffffu)) == 0) {
        fill_bytes(out_scene, f.C * HW, f.bg & 255u, lane);       // grey background: one run
    } else {
        for (int c = 0; c < f.C; ++c) fill_bytes(out_scene + (size_t)c * HW, HW, (f.bg >> (8 * c)) & 255u, lane);
    }
}

// one third of the background copy (vectors [part*n16/3, (part+1)*n16/3) of the 16-byte-vector image):
// spreading the bursts between the geometry phases keeps the load/store queue from filling up
// (measured: background + geometry 13.7 -> 12.5 us).  Pushing parts into the raster sweep as well
// (patch lists, flush after a CTA barrier) was measured slower: 34.8 vs 28.3 us per frame -- the
// copy's load/store-queue stalls then hit the warps that should be sweeping.
__device__ __forceinline__ void write_background_part(const FrameDev &f, unsigned char *out_scene, int HW, int lane,
                                                      int part) {
    const int n16 = f.C * HW / 16;
    const int v0 = (int)((long long)n16 * part / 3), v1 = (int)((long long)n16 * (part + 1) / 3);
    const uint4 *src = reinterpret_cast<const uint4 *>(f.base_color);
    uint4 *dst = reinterpret_cast<uint4 *>(out_scene);
    int i = v0 + lane;
    for (; i + 7 * 32 < v1; i += 8 * 32) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(src + i + k * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[i + k * 32] = v[k];
    }
    for (; i < v1; i += 32) dst[i] = __ldg(src + i);
}

// One block of a scene that has records in the overflow pool: shared-memory records, then pool
// records, then the pixel patch.  Out of line and only run by the scene's own warp, so the main
// sweep loop carries no trace of the overflow machinery.
__device__ __noinline__ void overflow_block(const FrameDev &f, const Rec *srecs, const unsigned *smasks, int entry,
                                            int nblk, int packed, unsigned char *out_scene, int lane) {
    const int bx = packed & 255, by = (packed >> 8) & 255;
    const int HW = f.H * f.W;
    const int px = bx * 8 + (lane & 7), py0 = by * 8 + (lane >> 3);
    const bool ok0 = px < f.W && py0 < f.H, ok1 = px < f.W && py0 + 4 < f.H;
    const int b = by * f.nbx + bx;
    PixelState ps;
    ps.k0 = ps.k1 = KEY_CLEAR;
    if (f.base_flags != nullptr && __ldg(f.base_flags + b) != 0) {
        ps.k0 = __ldg(f.base_keys + (size_t)b * 64 + lane);
        ps.k1 = __ldg(f.base_keys + (size_t)b * 64 + 32 + lane);
    }
    ps.c0 = ps.c1 = 0u;
    const unsigned id0 = (unsigned)ps.k0, id1 = (unsigned)ps.k1;
    raster_block<W_MW, false>(srecs, smasks + b * W_MW, px, py0, ok0, ok1, ps, &f);
    raster_block<W_OVF_MW, false>(f.ovf_recs + (size_t)entry * W_OVF_MAXREC,
                                  f.ovf_masks + ((size_t)entry * nblk + b) * W_OVF_MW, px, py0, ok0, ok1, ps, &f);
    unsigned char *p = out_scene + py0 * f.W + px;
    if (key_changed(ps.k0, id0)) {
        p[0] = (unsigned char)(ps.c0 & 255u);
        p[HW] = (unsigned char)((ps.c0 >> 8) & 255u);
        p[2 * HW] = (unsigned char)((ps.c0 >> 16) & 255u);
        if (f.C == 4) p[3 * HW] = (unsigned char)(ps.c0 >> 24);
    }
    if (key_changed(ps.k1, id1)) {
        p += 4 * f.W;
        p[0] = (unsigned char)(ps.c1 & 255u);
        p[HW] = (unsigned char)((ps.c1 >> 8) & 255u);
        p[2 * HW] = (unsigned char)((ps.c1 >> 16) & 255u);
        if (f.C == 4) p[3 * HW] = (unsigned char)(ps.c1 >> 24);
    }
}

// Block lists of a scene with an overflow pool entry: blocks with pool records go to the back of
// blist (for overflow_block), the others with any record to the front (shared sweep).  A block is in
// at most one list, so the two cannot meet.  Returns front count | back count << 16.
__device__ __noinline__ int overflow_lists(const FrameDev &f, const unsigned *masks, unsigned short *blist, int entry,
                                           int nblk, int lane) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int nlist = 0, novf = 0;
#pragma unroll 1
    for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int b = b0 + lane;
        bool nz = false, ov = false;
        int packed = 0;
        if (b < nblk) {
            const unsigned *om = f.ovf_masks + ((size_t)entry * nblk + b) * W_OVF_MW;
#pragma unroll
            for (int k = 0; k < W_OVF_MW; ++k) ov |= om[k] != 0u;
            nz = !ov && (masks[b * W_MW] | masks[b * W_MW + 1]) != 0u;
            const int by = fast_div(b, f.nbx_magic);
            packed = (by << 8) | (b - by * f.nbx);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, nz), obal = __ballot_sync(0xffffffffu, ov);
        if (nz) blist[nlist + __popc(bal & lt_mask)] = (unsigned short)packed;
        if (ov) blist[nblk - 1 - novf - __popc(obal & lt_mask)] = (unsigned short)packed;
        nlist += __popc(bal);
        novf += __popc(obal);
    }
    return nlist | (novf << 16);
}

// the TMA thread of a TMA_BG CTA (lane 0 of its first helper warp), once the image has landed in shared memory (mbarrier at qctr + 4 ints, image
// at qctr + 8 ints): one bulk store per scene of the CTA, committed as one bulk group
__device__ __noinline__ void issue_bg_stores(const FrameDev &f, int *qctr, int warps, unsigned *pass_turn = nullptr) {
    mbar_wait(reinterpret_cast<unsigned long long *>(qctr + 4), 0);
    const size_t bytes = (size_t)f.C * f.H * f.W;
    const int first = f.scene_begin + (int)blockIdx.x * warps;
    const int n = min(warps, f.scene_begin + f.scene_count - first);
    for (int w = 0; w < n; ++w) tma_store(f.out + (size_t)(first + w) * bytes, qctr + 8, (unsigned)bytes);
    tma_commit();
    if (pass_turn != nullptr) atomicAdd(pass_turn, 1u);      // the next CTA of this SM may issue its stores
}

// TMA_BG: the background / static-layer image of every scene of the CTA is written by the TMA engine
// instead of by the warps: one bulk load brings the C*H*W image (L2 resident, shared by all scenes)
// into shared memory once per CTA, then one bulk store per scene sends it to out[scene]; the TMA
// thread (lane 0 of the last helper warp) waits for the stores ahead of the CTA barrier that precedes
// the sweep (they complete ~5.4 us after the CTA's entry, the slowest scene warp is ready at ~9 us; a flag checked
// before every warp's first pixel patch instead -- round 1 -- costs more under chained frames: 16.82 vs 16.54 us);
// the CTAs of an SM take turns at issuing their stores (PBR_W_BG_SERIAL).  The warps issue
// no background instruction at all (the copy was ~9 % of their instructions and the source of the
// lg_throttle stalls).  The image costs C*H*W bytes of shared memory per CTA, which is why this
// variant packs 14 scenes per CTA: 2 CTAs of 14 warps keep 28 scenes per SM resident, enough for the
// 4096-scene batch to stay one wave on 148 SMs (measured on one box: 7 scenes x 4 CTAs 23.6 us,
// 10 x 3 24.2, 14 x 2 22.8, 28 x 1 23.7 -- larger CTAs balance the shared sweep better and hold fewer
// copies of the image, until the barrier spans too many warps).
// Helper warps: the TMA build adds two warps per CTA that own no scene and only join the shared sweep
// (16 warps x 2 CTAs fill the SM's 1024 threads at 64 registers).  Measured on one box: 22.8 us
// without, 22.7 with one, 22.4 with two.  The first of them also drives the TMA engine (PBR_W_BG_HELPER;
// with the engine driven by a scene warp: 21.5 us against 20.5).
#ifndef PBR_W_HELPERS
#define PBR_W_HELPERS 2
#endif
// ... and with the 6-scene CTAs (three per SM): measured at 84x84 x 16,384 scenes 102.8 us with 2 helpers, 101.0 with 3,
// 99.5 with 4 (the 14-scene CTAs get slower with more: 15.30 / 15.79 / 15.70 us at 64x64 x 4096, 56 registers per thread)
#ifndef PBR_W_HELPERS_SMALL
#define PBR_W_HELPERS_SMALL 4
#endif
#ifndef PBR_W_BG_HELPER
#define PBR_W_BG_HELPER 1
#endif
// The CTAs that share an SM take turns with their bulk stores (per-SM ticket / done counters in global
// memory): the SM's write path into the L2 is the limit, so stores issued together also finish together
// (~9 us), while in turns the first CTA's are complete when its geometry is (~6 us) and its sweep starts
// 2.5 us earlier, beside the second CTA's stores.
// 0 = all CTAs issue at once; 1 = the turn passes when a CTA's stores are complete; 2 = when they are issued
// (the engine then works through them roughly in order without a bubble).  Measured on one box, kernel /
// step: 20.1 / 21.25 us (0), 19.6 / 21.2 (1), 19.0 / 20.6 (2); 84x84, 16384 scenes: 95.7, 90.6, 90.2 us.
// (Passing the turn after 7 / 10 / 12 of the 14 stores: 19.1 / 19.0 / 19.1 us -- no better.)
#ifndef PBR_W_BG_SERIAL
#define PBR_W_BG_SERIAL 2
#endif
// (Round 1 let the barrier before the sweep wait for the geometry only and had the TMA thread raise a flag behind it that
// every warp checked before its first pixel patch.  With the frames chained -- the CTAs of the next frame arrive while
// this one sweeps -- that no longer pays: 16.82 us per frame with the flag, 16.54 with the TMA thread simply waiting for
// its stores ahead of the barrier, 84x84 113.6 vs 111.7 us per 16,384 scenes.  Removed.)
// the sweep pops the item after the current one before it sweeps the current one (16.54 -> 16.51 us; 84x84: 111.5 -> 110.3)
// blocks with at least this many records are swept first (0: in list order)
// a CTA asks the L2 for the per-scene inputs of the CTA this many places behind it (0: off)
#ifndef PBR_W_PF_DIST
#define PBR_W_PF_DIST 32
#endif
#ifndef PBR_W_HEAVY
#define PBR_W_HEAVY 4
#endif
#ifndef PBR_W_POP_AHEAD
#define PBR_W_POP_AHEAD 1
#endif
// Programmatic launch chain.  The kernel is launched with programmatic stream serialisation and triggers its
// dependents (PBR_W_TRIGGER: 1 = at entry, 2 = behind the pre-sweep barrier, 3 = at the end of the sweep, 0 = never):
// when the next launch on the stream is another frame of this kernel -- the loop `renderer.step(state)` produces
// exactly that now that the pose is computed in phase A -- its CTAs are scheduled on each SM as soon as this
// frame's CTAs leave it, instead of after the slowest SM of this frame has finished and the launch latency has
// passed.  Nothing a frame reads is written by the frame before it (state, matrices, colours, the static layer are
// produced by ordinary launches, which complete before this kernel starts and never trigger early), so the only
// hazards are the writes: two frames in flight into overlapping `out` memory, or into the same out_mats.  The
// host knows the output ranges of the last two small-scene launches on the stream and sets f.sync_early when the
// new frame overlaps one of them (or writes matrices): every warp then executes griddepcontrol.wait before its
// first global write.  Two ranges suffice: frame n+2 cannot start before every CTA of frame n+1 is resident,
// i.e. before frame n has all but drained, and no CTA of frame n exits before frame n-1 has completed, because
// every thread ends with griddepcontrol.wait -- which also makes "this grid has completed" imply "every earlier
// frame has completed" for whatever follows on the stream (copies, torch kernels: ordinary, fully ordered launches).
#ifndef PBR_W_TRIGGER
#define PBR_W_TRIGGER 1
#endif
// phase S with two lanes per triangle (edges / shade) instead of one
#ifndef PBR_W_SPLIT_SETUP
#define PBR_W_SPLIT_SETUP 1
#endif
// A scene's non-empty blocks into the CTA's queue: blocks with many records from the front, the others from the back
// (qctr[0]: front count in the low half of the word, back count in the high half).  ROUNDS > 0: tiles of up to 32 x ROUNDS
// blocks -- every round's ballots stay in registers and ONE atomic per scene reserves both ranges (an atomic or two per
// 32 blocks made the 14 scene warps of a CTA queue up behind each other: 14.18 -> 13.83 us per frame); ROUNDS == 0: any
// tile, one atomic per 32 blocks.
template <int ROUNDS>
__device__ __forceinline__ void queue_blocks(const unsigned *masks, int nblk, int lane, unsigned lt_mask, int *qctr,
                                             unsigned *queue, const unsigned *btab, unsigned tag, int qcap) {
    auto classify = [&](int b, bool &nz, bool &heavy) {
        nz = false; heavy = true;
        if (b < nblk) {
            const uint2 m = *reinterpret_cast<const uint2 *>(masks + b * W_MW);
            nz = (m.x | m.y) != 0u;
            if (PBR_W_HEAVY > 0) heavy = __popc(m.x) + __popc(m.y) >= PBR_W_HEAVY;
        }
    };
    if (ROUNDS > 0) {
        constexpr int R = ROUNDS > 0 ? ROUNDS : 1;
        unsigned hbal[R], lbal[R];
        bool nz[R], hv[R];
        int nh = 0, nl = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            classify(r * 32 + lane, nz[r], hv[r]);
            hbal[r] = __ballot_sync(0xffffffffu, nz[r] && hv[r]);
            lbal[r] = PBR_W_HEAVY > 0 ? __ballot_sync(0xffffffffu, nz[r] && !hv[r]) : 0u;
            nh += __popc(hbal[r]);
            nl += __popc(lbal[r]);
        }
        if (nh + nl == 0) return;
        unsigned base = 0;
        if (lane == 0) base = (unsigned)atomicAdd(&qctr[0], nh | (nl << 16));
        base = __shfl_sync(0xffffffffu, base, 0);
        int hb = (int)(base & 0xffffu), lb = (int)(base >> 16);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (nz[r])
                queue[hv[r] ? hb + __popc(hbal[r] & lt_mask) : qcap - 1 - (lb + __popc(lbal[r] & lt_mask))] =
                    tag | btab[r * 32 + lane];
            hb += __popc(hbal[r]);
            lb += __popc(lbal[r]);
        }
    } else {
#pragma unroll 1
        for (int b0 = 0; b0 < nblk; b0 += 32) {
            const int b = b0 + lane;
            bool nz, heavy;
            classify(b, nz, heavy);
            const unsigned hbal = __ballot_sync(0xffffffffu, nz && heavy);
            const unsigned lbal = PBR_W_HEAVY > 0 ? __ballot_sync(0xffffffffu, nz && !heavy) : 0u;
            if ((hbal | lbal) == 0u) continue;
            unsigned base = 0;
            if (lane == 0) base = (unsigned)atomicAdd(&qctr[0], __popc(hbal) | (__popc(lbal) << 16));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (nz)
                queue[heavy ? (int)(base & 0xffffu) + __popc(hbal & lt_mask)
                            : qcap - 1 - ((int)(base >> 16) + __popc(lbal & lt_mask))] = tag | btab[b];
        }
    }
}

__host__ __device__ constexpr int w_helpers(int warps) { return warps <= W_WARPS_TMA_SMALL ? PBR_W_HELPERS_SMALL : PBR_W_HELPERS; }
template <int WARPS, bool TMA_BG>
__global__ void __launch_bounds__(32 * (WARPS + (TMA_BG ? w_helpers(WARPS) : 0)),
                                  (WARPS + (TMA_BG ? w_helpers(WARPS) : 0)) > 16 ? 2 : 32 / (WARPS + (TMA_BG ? w_helpers(WARPS) : 0)))
raster_warp_kernel(const __grid_constant__ FrameDev f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
#ifdef PBR_W_TIMING
    unsigned long long *tdump = reinterpret_cast<unsigned long long *>(f.ovf_recs) + ((size_t)blockIdx.x * 16 + warp) * 16;
#define W_STAMP(k) do { if (lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tdump[k] = t_; } } while (0)
    W_STAMP(0);
    if (lane == 0) {
        unsigned smid_, wid_;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid_));        // hardware warp slot
        tdump[7] = smid_ | ((unsigned long long)wid_ << 32);
    }
#else
#define W_STAMP(k) do { } while (0)
#endif
    if (PBR_W_TRIGGER == 1) asm volatile("griddepcontrol.launch_dependents;");
    const unsigned lt_mask = (1u << lane) - 1u;
    const int scene = f.scene_begin + (int)blockIdx.x * WARPS + warp;
    const bool helper = warp >= WARPS;                    // extra warps without a scene: they only sweep
    const bool active = !helper && scene < f.scene_begin + f.scene_count;
    const int nblk = f.nbx * f.nby;
    const int HW = f.H * f.W;
    const size_t scene_bytes_out = (size_t)f.C * HW;

    // ---- carve shared memory: one region per scene, then the CTA's block queue, live list and counters
    const unsigned region = (unsigned)f.w_region;       // == warp_scene_bytes(nblk), precomputed by the host
    const WScene me = wscene(smem_raw + warp * region, nblk);
    if (lane == 0 && !helper) *me.out_slot = f.out + (size_t)scene * scene_bytes_out;
    unsigned *queue = reinterpret_cast<unsigned *>(smem_raw + WARPS * region);   // [WARPS * nblk]
    // [nblk] what an item says about its block, the same for every scene: bits 0-7 bx, 8-15 by, bit 31 the static layer
    // covers part of it (a global load) -- looked up once per CTA instead of once per (scene, block)
    unsigned *const btab = queue + align16((size_t)WARPS * nblk * 4) / 4;
    // [3][W_MAXSLOT] what a triangle slot means, the same for every scene of the frame: the parked vertices of its
    // corners (a byte each) + the two-sided bit, its draw id, its (node, instance, triangle) -- decoded once per CTA
    // (node search, two divides, the load of the index triple) instead of once per (scene, slot) in phase B and again
    // in phase S
    unsigned *const stab_v = btab + align16((size_t)nblk * 4) / 4, *const stab_id = stab_v + W_MAXSLOT,
                   *const stab_slot = stab_id + W_MAXSLOT;
    int *qctr = reinterpret_cast<int *>(smem_raw + f.w_qctr_off);
    // live list, just below the counters: three words per surviving triangle -- (scene, slot, record index), the parked
    // vertices of its corners (a byte each) + the two-sided bit, the draw id: what the set-up's edge lanes need, so that they
    // neither search the node nor load the index triple again
    unsigned *livelist = reinterpret_cast<unsigned *>(qctr) - 3 * WARPS * W_MAXSLOT;
    unsigned *const live_v = livelist + WARPS * W_MAXSLOT, *const live_id = live_v + WARPS * W_MAXSLOT;
    // counters: [0] items queued from the front (low half) and from the back (high half) of the queue, [1] items popped,
    // [2] live triangles ([4], [5]: the mbarrier of the TMA build)
    // (the pop counter starts behind the warps' first items: warp w begins with item w -- sixteen first pops on one
    // counter would queue up behind each other for ~0.4 us at the start of every sweep)
    if (threadIdx.x == 0) { qctr[0] = 0; qctr[1] = (PBR_W_POP_AHEAD && WARPS > 1) ? (int)(blockDim.x >> 5) : 0; qctr[2] = 0; }
    __syncthreads();          // counters initialised (all warps arrive together: cheap)
    W_STAMP(8);
    // The thread that drives the TMA engine: lane 0 of the LAST helper warp when there is one.  Issuing the
    // CTA's bulk stores blocks the issuing thread for microseconds (time stamps: a scene warp that did it
    // reached the sweep 3.5 us after its 13 neighbours, and the whole CTA waited for it at the barrier), so
    // it must not be a warp that has geometry to do.  It loads the image, waits for it, issues the
    // stores and waits for them while the other warps do geometry.  (When the frame before this one may still
    // be writing the same memory -- f.sync_early -- it first waits for that grid to complete.)
    constexpr int HELP = TMA_BG ? w_helpers(WARPS) : 0;
    constexpr bool BG_HELP = TMA_BG && HELP > 0 && PBR_W_BG_HELPER != 0;
    constexpr int BG_T = BG_HELP ? (WARPS + HELP - 1) * 32 : 0;
    // the warps that share the geometry work: the scene warps and the helpers that do not drive the TMA engine
    constexpr int GW = BG_HELP ? WARPS + HELP - 1 : WARPS + HELP;
    const bool worker = warp < GW;
    unsigned bg_turn = 0, bg_sm = 0;
    if (TMA_BG && threadIdx.x == BG_T) {
        unsigned long long *bg_bar = reinterpret_cast<unsigned long long *>(qctr + 4);
        mbar_init(bg_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bg_bar, (unsigned)scene_bytes_out);
        tma_load(qctr + 8, f.base_color, (unsigned)scene_bytes_out, bg_bar);
        if (f.sync_early) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (BG_T != 0) {
            if (PBR_W_BG_SERIAL) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                bg_turn = atomicAdd(f.bg_ticket + smid, 1u);
                // every holder of an earlier ticket on this SM is resident and does not depend on us
                while ((int)(*reinterpret_cast<volatile unsigned *>(f.bg_done + smid) - bg_turn) < 0) __nanosleep(200);
                bg_sm = smid;
            }
            issue_bg_stores(f, qctr, WARPS, PBR_W_BG_SERIAL == 2 ? f.bg_done + bg_sm : nullptr);
        }
    }

    // blocks for the shared sweep; novf: 0 = no overflow pool entry claimed, else 1 + number of blocks
    // that have records in the pool (swept by overflow_block after the shared sweep)
    int nlist = 0, novf = 0;
    int nlight = 0;                          // PBR_W_HEAVY: blocks with few records, listed from the back of blist
    bool scene_slow = false;                 // the scene has int64 (slow-path) records: its items say so (bit 30)
    if (TMA_BG && BG_T == 0 && f.debug == 1 && threadIdx.x == 0) issue_bg_stores(f, qctr, WARPS);

    // =====================================================================================================
    // Geometry, shared by the CTA.  A scene has far fewer vertices / triangles than a warp has lanes
    // (CartPole: 16 vertices, 24 triangle slots of which ~11 face the camera), so a warp that works on its
    // own scene alone runs the expensive phases with a third of its lanes.  Instead the worker warps of the
    // CTA flatten (scene, item) pairs of all its scenes into full warps, phase by phase, with a named
    // barrier (bar.sync 1, the TMA warp is not part of it) between the phases:
    //   M  instances  lanes = (scene, instance): model matrix from the pose channels / matrix buffer
    //   A  vertices   lanes = (scene, instance, unique vertex): clip = VP*(M*v), outcodes, project + snap
    //   B  classify   lanes = (scene, triangle slot): reject / cull / clip from the parked vertices; survivors
    //                 go to the CTA's live list
    //   S  set-up     lanes = live list entries: edge equations, depth plane, flat shade -> record
    // and each scene's own warp finishes with binning, clipped fans and the block list.
    // =====================================================================================================
    const int first_scene = f.scene_begin + (int)blockIdx.x * WARPS;
    const int n_sc = min(WARPS, f.scene_begin + f.scene_count - first_scene);
    const bool geom = f.debug != 1;
    const int S = f.total_slots;
    const bool direct = S <= 32;             // record index = slot; else survivors draw an index per scene
    unsigned char *const out_scene = f.out + (size_t)scene * scene_bytes_out;
    const bool split_bg = !TMA_BG && f.base_color != nullptr && ((f.C * HW) & 15) == 0 && geom;
    if (active) {
        if (f.sync_early) asm volatile("griddepcontrol.wait;" ::: "memory");
        // ---- 0: background (warp-copy build only; three bursts between the phases)
        if (!TMA_BG) {
            if (split_bg) write_background_part(f, out_scene, HW, lane, 0);
            else write_background(f, out_scene, HW, lane);
        }
        if (geom) {
            {   // masks are 8 bytes per block, region padded to 16 bytes: clear with 128-bit stores
                uint4 *m4 = reinterpret_cast<uint4 *>(me.masks);
                const int n16 = (nblk * W_MW * 4 + 15) >> 4;
#pragma unroll 1
                for (int i = lane; i < n16; i += 32) m4[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (lane < 4) me.ctr[lane] = 0;
        }
    } else if (worker && f.sync_early && f.write_mats) {
        asm volatile("griddepcontrol.wait;" ::: "memory");       // helpers write matrices in phase M
    }
    if (worker && geom) {
        const int wl = warp * 32 + lane;     // this lane among the worker lanes

        // The per-scene inputs of a CTA -- pose channel values (the caller's state), VP rows, instance colours -- are
        // read once per frame and miss the L2 (50 MB of frames pass through it between two reads of a line): the first
        // load of phase M alone takes 1.5 us.  The CTAs of a frame enter the SMs over ~15 us in index order, so the last
        // worker warp of each CTA asks the L2 for the rows of the CTA PBR_W_PF_DIST places behind it, which enters a few
        // microseconds later: lane = (range, 128-byte line), the ranges listed by the host.  (L1 prefetch of the CTA's
        // own lines at entry was measured slower: 17.31 vs 16.78 us, the 8 KB of L1 beside 219 KB of shared memory do
        // not hold them.)
        constexpr int PW = GW > 1 ? GW - 1 : 0;          // the warp that runs the pose chain of phase M: the last worker
        if (PBR_W_PF_DIST > 0 && warp == (GW > 2 ? GW - 2 : 0) && (lane >> 3) < f.n_pf) {
            const int t_first = first_scene + PBR_W_PF_DIST * WARPS;
            const int t_n = min(WARPS, f.scene_begin + f.scene_count - t_first);
            if (t_n > 0) {
                const int r = lane >> 3;
                const size_t a0 = reinterpret_cast<size_t>(f.pf_ptr[r]) + (size_t)t_first * f.pf_row[r];
                const size_t a = (a0 & ~(size_t)127) + (size_t)(lane & 7) * 128;
                if (a < a0 + (size_t)t_n * f.pf_row[r]) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }

        // (filled by the worker lanes with the highest indices below the pose warp; the readers are behind the phase
        // barriers)
        for (int b = (PW > 0 ? PW : 1) * 32 - 1 - wl; b < nblk; b += (PW > 0 ? PW : 1) * 32) {
            if (b < 0) break;                             // (the pose warp)
            const int by = fast_div(b, f.nbx_magic);
            btab[b] = (unsigned)((by << 8) | (b - by * f.nbx)) |
                      ((f.base_flags != nullptr && __ldg(f.base_flags + b) != 0) ? 0x80000000u : 0u);
        }

        if (warp == (PW > 3 ? PW - 3 : 0)) {
#pragma unroll 1
            for (int s = lane; s < S && s < W_MAXSLOT; s += 32) {
                int ni = 0;
#pragma unroll 1
                for (int i = 1; i < f.n_nodes; ++i)
                    if (s >= f.nodes[i].slot_begin) ni = i;
                const NodeDev &nd = f.nodes[ni];
                const int local = s - nd.slot_begin;
                const int inst = fast_div(local, nd.tri_magic);
                const int tri = local - inst * nd.n_tris;
                const uint4 ti = __ldg(nd.tidx + tri);
                const int vb = nd.vert_begin + inst * nd.n_verts;
                stab_v[s] = (unsigned)(vb + ti.x) | ((unsigned)(vb + ti.y) << 8) | ((unsigned)(vb + ti.z) << 16) |
                            ((nd.flags & PBR_MESH_TWO_SIDED) ? 1u << 24 : 0u);
                stab_id[s] = (unsigned)(nd.id_begin + inst * nd.n_tris + tri) + 1u;
                stab_slot[s] = pack_slot(ni, inst, tri);
            }
        }

        // ---- M: instances.  A posed node's model matrix is computed here from its pose channels (the state
        // tensor of the caller: reference envs/cartpole/renderer.py:125-138 + shader_context.py:47-84), other
        // nodes' matrices are read from their matrix buffer; parked in shared memory for phases A and S.
        {
            const int TI = f.total_inst;
            // (work items start at the last worker warp -- with the TMA build a helper: no scene of its own to prepare, no
            // vertex in phase A whose loads it would issue first, and the block table and the prefetch are with the
            // warps before it: the pose chain is the longest thing between kernel entry and the first barrier)
            const int wlm = wl - PW * 32 + (wl < PW * 32 ? GW * 32 : 0);
#pragma unroll 1
            for (int it = wlm; it < n_sc * TI; it += GW * 32) {
                const int sl = fast_div(it, f.w_inst_magic);
                const int gi = it - sl * TI;
                int ni = 0;
#pragma unroll 1
                for (int i = 1; i < f.n_nodes; ++i)
                    if (gi >= f.nodes[i].inst_begin) ni = i;
                const NodeDev &nd = f.nodes[ni];
                const int inst = gi - nd.inst_begin;
                const size_t b = nd.shared ? (size_t)inst : (size_t)(first_scene + sl) * nd.inst + inst;
                float M[16];
                if (nd.pose_idx >= 0) pose_matrix(f.poses[nd.pose_idx], b, M);
                else load_mat(nd.mats + b * 16, M);
                float4 *mi = wscene(smem_raw + sl * region, nblk).minst + gi * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) mi[j] = make_float4(M[4 * j], M[4 * j + 1], M[4 * j + 2], M[4 * j + 3]);
                if (f.write_mats && nd.pose_idx >= 0) {
                    float4 *o = reinterpret_cast<float4 *>(f.poses[nd.pose_idx].out_mats + b * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = make_float4(M[4 * j], M[4 * j + 1], M[4 * j + 2], M[4 * j + 3]);
                }
            }
        }
        // ---- A: vertices (basic.vert:24-43).  The global loads of a lane's first vertex -- its scene's VP rows and the
        // object-space position -- are issued ahead of the barrier behind phase M: they are in flight while the pose
        // lanes finish (only the model matrix, which comes out of shared memory, depends on phase M)
        {
            const int TV = f.total_verts;
            const int n_a = n_sc * TV;
            auto decode = [&](int it, int &sl, int &v, int &mi_off, const float4 *&vp) {
                sl = fast_div(it, f.w_vert_magic);
                v = it - sl * TV;
                int ni = 0;
#pragma unroll 1
                for (int i = 1; i < f.n_nodes; ++i)
                    if (v >= f.nodes[i].vert_begin) ni = i;
                const NodeDev &nd = f.nodes[ni];
                const int local = v - nd.vert_begin;
                const int inst = fast_div(local, nd.vert_magic);
                mi_off = (nd.inst_begin + inst) * 4;
                vp = nd.vpos + (local - inst * nd.n_verts);
            };
            int sl = 0, v = 0, mi_off = 0;
            const float4 *vp = nullptr;
            float VP[16];
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (wl < n_a) {
                decode(wl, sl, v, mi_off, vp);
                load_mat(f.vp + (size_t)(first_scene + sl) * 16, VP);
                p = __ldg(vp);
            }
            group_sync<GW * 32>();
            W_STAMP(9);
#pragma unroll 1
            for (int it = wl; it < n_a; it += GW * 32) {
                if (it != wl) {
                    decode(it, sl, v, mi_off, vp);
                    load_mat(f.vp + (size_t)(first_scene + sl) * 16, VP);
                    p = __ldg(vp);
                }
                const WScene sc = wscene(smem_raw + sl * region, nblk);
                float M[16];
                {
                    const float4 *mi = sc.minst + mi_off;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 a = mi[j];
                        M[4 * j] = a.x; M[4 * j + 1] = a.y; M[4 * j + 2] = a.z; M[4 * j + 3] = a.w;
                    }
                }
                float world[4], c[4];
                mat_vec4(M, p.x, p.y, p.z, 1.0f, world);
                mat_vec4(VP, world[0], world[1], world[2], world[3], c);
                int flags = 0;
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) {
                    const float a = c[pl >> 1];
                    const bool out = (pl & 1) ? (a > c[3]) : (a < -c[3]);
                    flags |= out ? (1 << pl) : 0;
                }
                if (needs_clip(c)) flags |= VF_CLIP;
                int X = 0, Y = 0;
                float z = 0.0f;
                if (project_vertex(f, c, X, Y, z)) flags |= VF_PROJ;
                sc.clipc[v] = make_float4(c[0], c[1], c[2], c[3]);
                sc.proj[v] = make_int4(X, Y, __float_as_int(z), flags);
            }
        }
        W_STAMP(1);
        if (TMA_BG && BG_T == 0 && threadIdx.x == 0) issue_bg_stores(f, qctr, WARPS);      // the load has had phase A to arrive
        if (split_bg && active) write_background_part(f, out_scene, HW, lane, 1);
        group_sync<GW * 32>();
        W_STAMP(10);

        // ---- B: classify triangle slots; survivors go to the CTA's live list as (scene, slot, record index)
        {
#pragma unroll 1
            for (int base = warp * 32; base < n_sc * S; base += GW * 32) {
                const int it = base + lane;
                int cat = 0;            // 0 dead, 1 live, 2 clip
                int sl = 0, s = 0;
                unsigned pv = 0, pid = 0;
                if (it < n_sc * S) {
                    sl = fast_div(it, f.w_slot_magic);
                    s = it - sl * S;
                    pv = stab_v[s];
                    const int4 *pj = wscene(smem_raw + sl * region, nblk).proj;
                    const int4 q0 = pj[pv & 255u], q1 = pj[(pv >> 8) & 255u], q2 = pj[(pv >> 16) & 255u];
                    const int f_and = q0.w & q1.w & q2.w, f_or = q0.w | q1.w | q2.w;
                    if (f_and & 0x3f) {
                        cat = 0;
                    } else if (f_or & VF_CLIP) {
                        cat = 2;
                    } else if (f_and & VF_PROJ) {
                        const long long area2 = (long long)(q1.x - q0.x) * (q2.y - q0.y) -
                                                (long long)(q2.x - q0.x) * (q1.y - q0.y);
                        const bool two_sided = (pv >> 24) != 0u;
                        cat = (area2 < 0 || (two_sided && area2 > 0)) ? 1 : 0;
                        pid = stab_id[s];
                    }
                    if (cat == 2) {      // rare: the scene's own warp clips it later
                        const WScene sc = wscene(smem_raw + sl * region, nblk);
                        sc.clipl[atomicAdd(&sc.ctr[1], 1)] = stab_slot[s];
                    }
                }
                int j = s;
                if (!direct && cat == 1) j = atomicAdd(&wscene(smem_raw + sl * region, nblk).ctr[2], 1);
                const unsigned bl = __ballot_sync(0xffffffffu, cat == 1);
                int pos = 0;
                if (lane == 0 && bl) pos = atomicAdd(&qctr[2], __popc(bl));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (cat == 1) {
                    const int at = pos + __popc(bl & lt_mask);
                    livelist[at] = ((unsigned)sl << 16) | ((unsigned)s << 8) | (unsigned)j;
                    live_v[at] = pv;
                    live_id[at] = pid;
                }
            }
        }
        group_sync<GW * 32>();
        W_STAMP(11);

        // ---- S: set-up of the survivors.  Two lanes per triangle, in different warps: "edges" (integer edge
        // equations, depth plane -> 64-byte record, binned into the per-block masks) and "shade" (flat colour:
        // normal through the model matrix, ambient + Lambert, basic.frag:31-38 -> the record's colour word, which
        // the edge lane leaves alone).  The phase is a dependent chain per lane and only
        // ~150 of the CTA's 480 worker lanes have a triangle: splitting it shortens the chain by a third.
        {
            const int n_live = qctr[2];
            // (roles are contiguous ranges of the work items and the shade range starts at a warp boundary, so that no
            // warp runs both chains: the warp at the boundary used to reach the barrier 0.3 us after the others)
            const int shade0 = (n_live + 31) & ~31;
            const int n_work = PBR_W_SPLIT_SETUP ? shade0 + n_live : n_live;
#pragma unroll 1
            for (int it0 = wl; it0 < n_work; it0 += GW * 32) {
                const bool do_shade = !PBR_W_SPLIT_SETUP || it0 >= shade0;
                const bool do_edges = !PBR_W_SPLIT_SETUP || it0 < shade0;
                const int it = (PBR_W_SPLIT_SETUP && it0 >= shade0) ? it0 - shade0 : it0;
                if (PBR_W_SPLIT_SETUP && it >= n_live) continue;              // the gap between the two ranges
                const unsigned e = livelist[it];
                const int sl = (int)(e >> 16), s = (int)((e >> 8) & 255u), j = (int)(e & 255u);
                const WScene sc = wscene(smem_raw + sl * region, nblk);
                if (do_shade) {
                    const WSlot ws = unpack_slot(stab_slot[s]);
                    const NodeDev &nd = f.nodes[ws.ni];
                    const int inst = ws.inst, tri = ws.tri;
                    const size_t b = nd.shared ? (size_t)inst : (size_t)(first_scene + sl) * nd.inst + inst;
                    float n[3];
                    const float4 n0 = __ldg(nd.tn + 3 * tri);
                    xform_normal_cols(sc.minst + (nd.inst_begin + inst) * 4, n0.x, n0.y, n0.z, n);
                    sc.recs[j].col = shade(f, n, __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4)));
                }
                if (do_edges) {
                    const unsigned pv = live_v[it];
                    const unsigned id = live_id[it];
                    const bool two_sided = (pv >> 24) != 0u;
                    const int4 q0 = sc.proj[pv & 255u], q1 = sc.proj[(pv >> 8) & 255u], q2 = sc.proj[(pv >> 16) & 255u];
                    int X[3] = {q0.x, q1.x, q2.x}, Y[3] = {q0.y, q1.y, q2.y};
                    float z[3] = {__int_as_float(q0.z), __int_as_float(q1.z), __int_as_float(q2.z)};
                    Rec r;
                    BBox bb;
                    if (setup_snapped(f, X, Y, z, two_sided, id, 0, f.H, r, bb)) {
                        if (r.meta & M_SLOW) sc.ctr[3] = 1;
                        // (every field but the colour, which the shade lane of this triangle writes)
                        Rec *const d = sc.recs + j;
                        *reinterpret_cast<int4 *>(&d->e[0]) = make_int4(r.e[0], r.e[1], r.e[2], r.e[3]);
                        *reinterpret_cast<int4 *>(&d->e[4]) = make_int4(r.e[4], r.e[5], r.e[6], r.e[7]);
                        d->e[8] = r.e[8];
                        *reinterpret_cast<uint2 *>(&d->id) = make_uint2(r.id, r.meta);
                        *reinterpret_cast<float4 *>(&d->z0) = make_float4(r.z0, r.dz1, r.dz2, r.invA);
                        // into the masks of the 8x8 blocks its box touches (edge-function reject per block): the
                        // boxes are small (1-8 blocks), a loop per lane beats a pass with a lane per (record, block)
                        bin_record<W_MW>(r, bb, j, f.nbx, sc.masks);
                    }
                }
            }
        }
        W_STAMP(2);
        if (split_bg && active) write_background_part(f, out_scene, HW, lane, 2);
        group_sync<GW * 32>();
        W_STAMP(12);
    }

    // ---- per scene, by its own warp: binning, clipped fans, block list
    if (active && geom) {
        Rec *const recs = me.recs;
        unsigned *const masks = me.masks;
        unsigned short *const blist = me.blist;
        unsigned *const clipl = me.clipl;
        int *const ovf_entry = &me.ctr[0];
        const float4 *const clipc = me.clipc;
        const int nclip = me.ctr[1];
        const int nlive = direct ? S : me.ctr[2];
        int nrec = nlive;
        scene_slow = nclip > 0 || me.ctr[3] != 0;     // (clipped fans: not tracked, assume so)
        {
            // ---- B3: clipped triangles -> fan triangles in the spare record slots
            if (nclip > 0) {
                int entry = -1;                          // warp-uniform
                Rec *orecs = nullptr;
                unsigned *omasks = nullptr;
#pragma unroll 1
                for (int base = 0; base < nclip; base += 32) {
                    const int j = base + lane;
                    int cnt = 0;
                    CV poly[MAX_POLY];
                    float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
                    unsigned id = 0;
                    bool two_sided = false;
                    if (j < nclip) {
                        const WSlot ws = unpack_slot(clipl[j]);
                        const NodeDev &nd = f.nodes[ws.ni];
                        const uint4 ti = __ldg(nd.tidx + ws.tri);
                        const int vb = nd.vert_begin + ws.inst * nd.n_verts;
                        const size_t b = nd.shared ? (size_t)ws.inst : (size_t)scene * nd.inst + ws.inst;
                        float n[3];
                        const float4 n0 = __ldg(nd.tn + 3 * ws.tri);
                        xform_normal_cols(me.minst + (nd.inst_begin + ws.inst) * 4, n0.x, n0.y, n0.z, n);
                        col = __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4));
                        id = (unsigned)(nd.id_begin + ws.inst * nd.n_tris + ws.tri) + 1u;
                        two_sided = (nd.flags & PBR_MESH_TWO_SIDED) != 0;
                        CV v[3];
                        const unsigned vi[3] = {ti.x, ti.y, ti.z};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float4 c = clipc[vb + vi[k]];
                            v[k].c[0] = c.x; v[k].c[1] = c.y; v[k].c[2] = c.z; v[k].c[3] = c.w;
                            v[k].n[0] = n[0]; v[k].n[1] = n[1]; v[k].n[2] = n[2];
                        }
                        const int np = clip_poly(v, poly);
                        cnt = np >= 3 ? np - 2 : 0;
                    }
                    int incl = cnt;                      // inclusive prefix sum over the warp
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    const int start = nrec + incl - cnt;
                    if (entry < 0 && nrec + total > W_MAXREC) {
                        // claim a pool entry of this SM (lane 0), then clear its masks (all lanes)
                        unsigned smid;
                        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                        if (lane == 0) {
                            unsigned *busy = f.ovf_busy + smid;
                            while (entry < 0) {
                                const unsigned free_bits = ~atomicOr(busy, 0u);
                                if (free_bits == 0u) continue;
                                const int bit = __ffs(free_bits) - 1;
                                if (!(atomicOr(busy, 1u << bit) & (1u << bit))) entry = (int)smid * W_POOL_PER_SM + bit;
                            }
                            *ovf_entry = entry;
                            atomicOr(f.status, DEVSTAT_WARP_OVERFLOW);
                            *f.status_host = DEVSTAT_WARP_OVERFLOW;      // host prefers the general kernel from now on
                        }
                        entry = __shfl_sync(0xffffffffu, entry, 0);
                        novf = 1;
                        orecs = f.ovf_recs + (size_t)entry * W_OVF_MAXREC;
                        omasks = f.ovf_masks + (size_t)entry * nblk * W_OVF_MW;
                        for (int i = lane; i < nblk * W_OVF_MW; i += 32) omasks[i] = 0u;
                        __syncwarp();
                    }
#pragma unroll 1
                    for (int k = 0; k < cnt; ++k) {
                        const int idx = start + k;
                        int X[3], Y[3];
                        float z[3];
                        const bool ok = project_vertex(f, poly[0].c, X[0], Y[0], z[0]) &&
                                        project_vertex(f, poly[k + 1].c, X[1], Y[1], z[1]) &&
                                        project_vertex(f, poly[k + 2].c, X[2], Y[2], z[2]);
                        Rec r;
                        BBox bb;
                        if (ok && setup_snapped(f, X, Y, z, two_sided, id, 0, f.H, r, bb)) {
                            r.col = shade(f, poly[0].n, col);
                            if (idx < W_MAXREC) {
                                recs[idx] = r;
                                bin_record<W_MW>(r, bb, idx, f.nbx, masks);
                            } else {
                                orecs[idx - W_MAXREC] = r;
                                bin_record<W_OVF_MW>(r, bb, idx - W_MAXREC, f.nbx, omasks);
                            }
                        }
                    }
                    nrec += total;
                }
            }
            __syncwarp();     // records + masks complete; background stores ordered before patches
            W_STAMP(13);

            // ---- this scene's non-empty blocks: straight into the CTA's queue (several warps per CTA, no pool entry),
            // else into the scene's own list.  Item: bit 31 the static layer covers part of the block, bit 30 the scene
            // has int64 / clipped records, bit 29 sweep it with 64-bit keys (that, or the frame's draw order does not
            // allow 32-bit ones), bits 16.. the scene's warp, bits 8-15 by, 0-7 bx.  Blocks with many records go to the
            // front of the queue, the others fill it from the back: the queue hands out the expensive items first.
            if (f.debug != 2) {
                if (novf == 0 && WARPS > 1) {
                    const unsigned tag = ((unsigned)warp << 16) | (scene_slow ? 0x60000000u : 0u) |
                                         ((f.keys32 != 0 && direct) ? 0u : 0x20000000u);
                    // (separate copies for two and four rounds: skipping unused rounds at run time costs more than it saves)
                    if (nblk <= 64) queue_blocks<2>(masks, nblk, lane, lt_mask, qctr, queue, btab, tag, WARPS * nblk);
                    else if (nblk <= 128) queue_blocks<4>(masks, nblk, lane, lt_mask, qctr, queue, btab, tag, WARPS * nblk);
                    else queue_blocks<0>(masks, nblk, lane, lt_mask, qctr, queue, btab, tag, WARPS * nblk);
                } else if (novf == 0) {
#pragma unroll 1
                    for (int b0 = 0; b0 < nblk; b0 += 32) {
                        const int b = b0 + lane;
                        const bool nz = b < nblk && (masks[b * W_MW] | masks[b * W_MW + 1]) != 0u;
                        const unsigned bal = __ballot_sync(0xffffffffu, nz);
                        if (nz) blist[nlist + __popc(bal & lt_mask)] = (unsigned short)(btab[b] & 0xffffu);
                        nlist += __popc(bal);
                    }
                } else {
                    const int both = overflow_lists(f, masks, blist, *ovf_entry, nblk, lane);
                    nlist = both & 0xffff;
                    novf = 1 + (both >> 16);
                }
                __syncwarp();
            }
        }
    }

    // ---- D: raster.  One (scene, block) item at a time; with several warps per CTA the items of
    // all its scenes sit in one queue so that light scenes help heavy ones.
    W_STAMP(3);
    if (TMA_BG && threadIdx.x == BG_T) {
        tma_wait_all();                                  // background written before any pixel patch
        if (PBR_W_BG_SERIAL == 1 && BG_T != 0) atomicAdd(f.bg_done + bg_sm, 1u);      // next CTA of this SM: your turn
    }
    if (TMA_BG && warp == BG_T / 32) __syncwarp();       // lane 0 spun on the mbarrier / the bulk group: reconverge
    W_STAMP(4);
    if (WARPS > 1) {
        if (nlist > 0) {                                  // (a scene with a pool entry: the front of its own list)
            int qbase = 0;
            if (lane == 0) qbase = atomicAdd(&qctr[0], nlist) & 0xffff;
            qbase = __shfl_sync(0xffffffffu, qbase, 0);
            for (int i = lane; i < nlist; i += 32) {
                const int packed = me.blist[i];
                queue[qbase + i] = ((unsigned)warp << 16) | 0x60000000u | btab[(packed >> 8) * f.nbx + (packed & 255)];
            }
        }
        __syncthreads();                                  // every scene of the CTA is set up and queued
    }
    if (PBR_W_TRIGGER == 2) asm volatile("griddepcontrol.launch_dependents;");
    W_STAMP(5);
    const int nheavy = WARPS > 1 ? (qctr[0] & 0xffff) : nlist;       // items handed out from the front of the queue / list ...
    const int nitems = nheavy + (WARPS > 1 ? (int)((unsigned)qctr[0] >> 16) : nlight);     // ... then those from its back
    const int qlast = (WARPS > 1 ? WARPS * nblk : nblk) - 1 + nheavy;
    const int lx = lane & 7, ly = lane >> 3;
    const int tileW = f.W, tileH = f.H, tileH4 = f.H - 4, tileNbx = f.nbx;
    const bool rgba = f.C == 4;
    const unsigned pop_addr = smem_u32(&qctr[1]);
    const unsigned smem_base = smem_u32(smem_raw);
    int next = 0;
    int ahead = 0;                                        // PBR_W_POP_AHEAD: lane 0 holds the next item's index
    // (atom.inc with a bound below 2^32 - 1 stays one ATOMS.INC; ptxas rewrites a single-lane atom.add, and inc with
    // bound 0xffffffff, into leader election + popc + ATOMS.ADD, ~20 instructions.  A predicated atom in the asm
    // block instead of the branch: ptxas turns it back into the branch.)
#define PBR_W_POP(dst)                                                                                              \
    if (lane == 0) asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(dst) : "r"(pop_addr) : "memory")
    if (PBR_W_POP_AHEAD && WARPS > 1) ahead = warp;      // (lane 0's copy is the one that counts)
#pragma unroll 1
    while (true) {
        int i;
        unsigned item;
        if (PBR_W_POP_AHEAD && WARPS > 1) {
            // the pop of the item after this one is in flight while this one is swept (its result is first
            // read at the top of the next turn); the last pop of every warp runs past the end of the queue
            i = __shfl_sync(0xffffffffu, ahead, 0);
            if (i >= nitems) break;
            item = queue[(PBR_W_HEAVY > 0 && i >= nheavy) ? qlast - i : i];
            PBR_W_POP(ahead);
        } else if (WARPS > 1) {
            // (measured on one box: popping two items per atomic 27.1 us, static round robin without any
            // atomic 27.2 us, this single pop 25.3 us -- the dynamic balance is worth its instructions)
            i = 0;
            PBR_W_POP(i);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= nitems) break;
            item = queue[(PBR_W_HEAVY > 0 && i >= nheavy) ? qlast - i : i];
        } else {
            i = next++;
            if (i >= nitems) break;
            item = me.blist[(PBR_W_HEAVY > 0 && i >= nheavy) ? qlast - i : i] | (scene_slow ? 0x60000000u : 0u) | ((f.keys32 != 0 && direct) ? 0u : 0x20000000u);
            if (f.base_flags != nullptr && __ldg(f.base_flags + (int)((item >> 8) & 255u) * f.nbx + (int)(item & 255u)) != 0)
                item |= 0x80000000u;
        }
        const int w = (int)((item >> 16) & 0x1fffu);
        const int bx = (int)(item & 255u), by = (int)((item >> 8) & 255u);
        const unsigned char *sreg = smem_raw + w * region;
        const Rec *srecs = reinterpret_cast<const Rec *>(sreg + W_OFF_RECS);
        const unsigned *smasks = reinterpret_cast<const unsigned *>(sreg + W_OFF_MASKS);
        unsigned char *out_scene = *reinterpret_cast<unsigned char *const *>(sreg + W_OFF_OUT);

        const int px = bx * 8 + lx, py0 = by * 8 + ly;
        const bool ok0 = px < tileW && py0 < tileH, ok1 = px < tileW && py0 < tileH4;
        const int b = by * tileNbx + bx;
        // winners of this block: did this sweep win the lane's pixels, and with which colour
        bool won0, won1;
        unsigned c0, c1;
        if (!(item & 0x20000000u)) {
            // no clipped / int64 records in this scene, records in draw order, static layer drawn first:
            // 32-bit depth keys (see raster_block32)
            PixelState32 q;
            q.z0 = q.z1 = (unsigned)(KEY_CLEAR >> 32);
            if (item & 0x80000000u) {                     // static layer covers part of this block
                asm volatile("" ::: "memory");            // (a branch, not ten predicated-off instructions per item)
                const unsigned *bk = reinterpret_cast<const unsigned *>(f.base_keys + (size_t)b * 64);
                q.z0 = __ldg(bk + 2 * lane + 1);
                q.z1 = __ldg(bk + 2 * (32 + lane) + 1);
            }
            q.c0 = q.c1 = 0u;
            const unsigned zi0 = q.z0, zi1 = q.z1;
            raster_block32<W_MW>(smem_base + w * region + W_OFF_RECS, smem_base + w * region + W_OFF_MASKS + b * (W_MW * 4),
                                 px, py0, ok0, ok1, q);
            won0 = q.z0 != zi0; won1 = q.z1 != zi1;
            c0 = q.c0; c1 = q.c1;
        } else {
            PixelState ps;
            ps.k0 = ps.k1 = KEY_CLEAR;
            if (item & 0x80000000u) {
                ps.k0 = __ldg(f.base_keys + (size_t)b * 64 + lane);
                ps.k1 = __ldg(f.base_keys + (size_t)b * 64 + 32 + lane);
            }
            ps.c0 = ps.c1 = 0u;
            const unsigned id0 = (unsigned)ps.k0, id1 = (unsigned)ps.k1;
            if (WARPS > 1 && !(item & 0x40000000u))
                raster_block<W_MW, false, false>(srecs, smasks + b * W_MW, px, py0, ok0, ok1, ps, &f);
            else
                raster_block<W_MW, false, true>(srecs, smasks + b * W_MW, px, py0, ok0, ok1, ps, &f);
            won0 = key_changed(ps.k0, id0); won1 = key_changed(ps.k1, id1);
            c0 = ps.c0; c1 = ps.c1;
        }
#ifdef PBR_W_TIMING
        if (f.debug == 3) continue;                       // (profiling aid: sweep without the patches)
#endif
        // The lane's two pixels, three (four) planes each.  One 64-bit address per pixel; the planes are reached by
        // adding the plane size to it (tile width / plane size / channel count from locals: read through `f` they
        // are re-loaded after every byte store, which the compiler cannot prove not to alias the frame description).
        unsigned char *p = out_scene + (py0 * tileW + px);
        if (!rgba) {                                      // (one uniform branch, not a predicated fourth store per pixel)
            if (won0) {
                unsigned char *q = p;
                *q = (unsigned char)c0; q += HW;
                *q = (unsigned char)(c0 >> 8); q += HW;
                *q = (unsigned char)(c0 >> 16);
            }
            if (won1) {
                unsigned char *q = p + 4 * tileW;
                *q = (unsigned char)c1; q += HW;
                *q = (unsigned char)(c1 >> 8); q += HW;
                *q = (unsigned char)(c1 >> 16);
            }
        } else {
            if (won0) put_pixel(out_scene, HW, 4, tileW, px, py0, c0);
            if (won1) put_pixel(out_scene, HW, 4, tileW, px, py0 + 4, c1);
        }
    }
#undef PBR_W_POP
    W_STAMP(6);
    if (PBR_W_TRIGGER == 3) asm volatile("griddepcontrol.launch_dependents;");
    // blocks with records in the overflow pool: swept by the scene's own warp, then the entry is released
    if (novf > 0) {
        const int e = me.ctr[0];
#pragma unroll 1
        for (int i = 0; i + 1 < novf; ++i)
            overflow_block(f, me.recs, me.masks, e, nblk, me.blist[nblk - 1 - i], f.out + (size_t)scene * scene_bytes_out, lane);
        __syncwarp();
        if (lane == 0) atomicAnd(f.ovf_busy + e / W_POOL_PER_SM, ~(1u << (e % W_POOL_PER_SM)));
    }
    // no CTA leaves before the frame ahead of this one has completed (see "programmatic launch chain")
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace pbr
