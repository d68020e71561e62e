// raster_binned.cuh -- the large-scene path: per-8x8-block record lists, one warp per block.
//
// Large scenes (hundreds to tens of thousands of triangle slots per scene, tiles of 128^2 .. 256^2 and up:
// BASELINE configs 3 and 5, Steering-v0) go through six kernels per launch chunk:
//
//   cull_kernel       (raster_staged.cuh)  thread per (scene, instance): bounding sphere vs clip planes / pixel grid
//   bin_xform_kernel  thread per (scene, instance, unique vertex): clip = VP*(M*v), outcodes, project + snap
//                     (reference basic.vert:24-43) -- once per vertex, not once per triangle corner
//   bin_tri_kernel    thread per (scene, triangle slot): reject / cull / needs-clip from the three parked vertices;
//                     the survivors of the CTA are compacted in shared memory and set up on full warps (edge
//                     equations, depth plane, flat shade or per-pixel shading inputs) -> 64-byte record appended to
//                     the scene's list
//   bin_blocks_kernel<false>  thread per record: counts it in every 8x8 block it can touch
//   bin_scan_kernel   CTA per scene: exclusive prefix sum of the block counts = where each block's list starts
//   bin_blocks_kernel<true>   thread per record: its index into the list of each of those blocks
//   raster_binned_kernel  WARP per (scene, block): walks the block's list -- records gathered eight at a time into
//                     shared memory -- with the depth|id keys and colours of its 64 pixels in registers, then writes
//                     the finished block straight to out[scene]: no colour / depth tile in shared memory, no
//                     per-band rescans of the scene's records, no CTA barrier in the sweep.
//
// Measured against the band-based path it replaces (raster_staged.cuh: CTA per band of rows, 128-record chunks
// binned and swept behind CTA barriers, depth keys in shared memory) in profiles/README.md.
#pragma once
#include "raster_staged.cuh"

namespace pbr {

struct BinnedDev {
    float4 *vclip;           // [scenes_in_launch][total_verts] clip-space positions
    int4 *vproj;             // [scenes_in_launch][total_verts] snapped x, y, depth bits, flags (VF_*)
    // every block has two lists: [2*blk] big records (swept by the whole warp), [2*blk + 1] small ones (a lane each)
    int *blk_cnt;            // [scenes_in_launch][2 * nblk] records per list (count pass), then the fill cursors
    int *blk_off;            // [scenes_in_launch][2 * nblk + 1] start of each list in `pairs`
    unsigned *pairs;         // [scenes_in_launch][pairs_cap] record indices, block after block
    unsigned *pbox;          // [scenes_in_launch][cap] packed pixel box of small records (pack_pbox), else PBOX_NONE
    int pairs_cap;
    int total_verts;
};

constexpr int BV_CLIP = 0x40, BV_PROJ = 0x80;       // vertex flags (bits 0..5: outside clip plane p)
#ifndef PBR_B_THREADS
#define PBR_B_THREADS 256
#endif
constexpr int B_THREADS = PBR_B_THREADS;
// resident CTAs per SM the triangle and raster kernels are compiled for (register budget 65536 / (256 * n)); measured on
// configs 3 / 5 (ms per 1024 scenes): (4, 4) 0.352 / 2.325, (5, 4) 0.345 / 2.297, (4, 5) 0.362 / 2.400, (5, 5) 0.350 / 2.384,
// (6, 6) 0.375 / 2.530
#ifndef PBR_B_TRI_OCC
#define PBR_B_TRI_OCC 5
#endif
#ifndef PBR_B_RASTER_OCC
#define PBR_B_RASTER_OCC 32
#endif
#ifndef PBR_B_WPB
#define PBR_B_WPB 1
#endif
// (the raster warps are independent of each other: with eight per CTA a CTA's slot stays taken until its slowest warp is
// done -- measured on configs 3 / 5, ms per 1024 scenes: 8 warps 0.340 / 2.278, 4: 0.328 / 2.073, 2: 0.327 / 2.087,
// 1 (32 CTAs per SM): 0.318 / 2.025)
constexpr int B_WPB = PBR_B_WPB;                    // block-warps per CTA of the raster kernel
constexpr int B_GATHER = 8;                         // records staged per round (8 x 64 B = one 16-byte load per lane)

// ------------------------------------------------------------------------------------------------
// vertices
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(B_THREADS) bin_xform_kernel(const __grid_constant__ FrameDev f,
                                                              const __grid_constant__ StagedDev g,
                                                              const __grid_constant__ BinnedDev bd) {
    const int local_scene = blockIdx.y;
    const int scene = g.scene0 + local_scene;
    const int v = blockIdx.x * B_THREADS + threadIdx.x;
    if (v >= bd.total_verts) return;
    int ni = 0;
#pragma unroll 1
    for (int i = 1; i < f.n_nodes; ++i)
        if (v >= f.nodes[i].vert_begin) ni = i;
    const NodeDev &nd = f.nodes[ni];
    const int local = v - nd.vert_begin;
    const int inst = (bd.total_verts < 65536 && nd.n_verts < 65536) ? fast_div(local, nd.vert_magic) : local / nd.n_verts;
    const int vert = local - inst * nd.n_verts;
    if (!g.vis[(size_t)local_scene * f.total_inst + nd.inst_begin + inst]) return;
    const size_t b = nd.shared ? (size_t)inst : (size_t)scene * nd.inst + inst;
    float M[16], VP[16];
    const float4 *m4 = reinterpret_cast<const float4 *>(nd.mats + b * 16);
    const int vp_row = f.vp_scene_override >= 0 ? f.vp_scene_override : scene;
    const float4 *v4 = reinterpret_cast<const float4 *>(f.vp + (size_t)vp_row * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 a = __ldg(m4 + j), c = __ldg(v4 + j);
        M[4 * j] = a.x; M[4 * j + 1] = a.y; M[4 * j + 2] = a.z; M[4 * j + 3] = a.w;
        VP[4 * j] = c.x; VP[4 * j + 1] = c.y; VP[4 * j + 2] = c.z; VP[4 * j + 3] = c.w;
    }
    const float4 p = __ldg(nd.vpos + vert);
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
    if (needs_clip(c)) flags |= BV_CLIP;
    int X = 0, Y = 0;
    float z = 0.0f;
    if (project_vertex(f, c, X, Y, z)) flags |= BV_PROJ;
    const size_t o = (size_t)local_scene * bd.total_verts + v;
    bd.vclip[o] = make_float4(c[0], c[1], c[2], c[3]);
    bd.vproj[o] = make_int4(X, Y, __float_as_int(z), flags);
}

// ------------------------------------------------------------------------------------------------
// triangles
// ------------------------------------------------------------------------------------------------
// load_slot (raster_general.cuh) with the clip-space positions taken from the parked vertices
__device__ __forceinline__ void load_slot_parked(const FrameDev &f, int scene, int slot, int ni, int inst, int tri,
                                                 const float4 *vclip, SlotGeom &g) {
    const NodeDev &nd = f.nodes[ni];
    const int local = slot - nd.slot_begin;
    const size_t b = nd.shared ? (size_t)inst : (size_t)scene * nd.inst + inst;
    float M[16];
    const float4 *m4 = reinterpret_cast<const float4 *>(nd.mats + b * 16);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float4 a = __ldg(m4 + j);
        M[4 * j] = a.x; M[4 * j + 1] = a.y; M[4 * j + 2] = a.z; M[4 * j + 3] = a.w;
    }
    M[12] = M[13] = M[14] = M[15] = 0.0f;      // (the normal transform reads columns 0..2 only)
    g.col = __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4));
    g.two_sided = (nd.flags & PBR_MESH_TWO_SIDED) != 0;
    g.node = &nd;
    const bool textured = nd.tex != nullptr;
    g.id = (unsigned)(nd.id_begin + local) + 1u;
    const uint4 ti = __ldg(nd.tidx + tri);
    g.flat = ti.w != 0u;
    const unsigned vi[3] = {ti.x, ti.y, ti.z};
    const int vb = nd.vert_begin + inst * nd.n_verts;
    const float4 n0 = __ldg(nd.tn + 3 * tri);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 c = vclip[vb + vi[k]];
        g.v[k].c[0] = c.x; g.v[k].c[1] = c.y; g.v[k].c[2] = c.z; g.v[k].c[3] = c.w;
        const float4 n = (k == 0 || g.flat) ? n0 : __ldg(nd.tn + 3 * tri + k);
        xform_normal(M, n.x, n.y, n.z, g.v[k].n);
        float2 uv = make_float2(0.0f, 0.0f);
        if (textured && nd.tuv != nullptr) uv = __ldg(nd.tuv + 3 * tri + k);
        g.v[k].uv[0] = uv.x; g.v[k].uv[1] = uv.y;
    }
}

// append the record to the scene's list and count it in the blocks it can touch
__device__ __forceinline__ void bin_append(const FrameDev &f, const StagedDev &g, const BinnedDev &bd, int local_scene,
                                           const Rec &r, const BBox &bb, const CVT *vin, const SlotGeom &sg,
                                           const TriVary &tv) {
    const int idx = atomicAdd(&g.count[local_scene], 1);
    if (idx >= g.cap) {
        atomicOr(f.status, DEVSTAT_STAGED_OVERFLOW);
        f.status_host[1] = 1;            // host-mapped: the next large-scene call reports it and grows the lists
        return;
    }
    const size_t o = (size_t)local_scene * g.cap + idx;
    uint4 *dst = reinterpret_cast<uint4 *>(g.recs + o);
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    g.bbox[o] = (unsigned)bb.bx0 | ((unsigned)bb.by0 << 8) | ((unsigned)bb.bx1 << 16) | ((unsigned)bb.by1 << 24);
    if (g.srecs != nullptr && (r.meta & M_SMOOTH)) write_srec(f, g.srecs + o * (size_t)f.srec_stride, vin, sg, tv);
    // small pixel box and int32 edge functions: the raster kernel gives the record a lane of its own
    // (frames that shade per pixel only: measured on flat many-cubes frames the second list costs more than the
    // few small records save -- 0.43 vs 0.38 ms per 1024 scenes -- while mixed-mesh frames gain 2.59 -> 2.35 ms)
    const unsigned pb = ((r.meta & M_SLOW) || !f.smooth) ? PBOX_NONE : pack_pbox(bb);
    bd.pbox[o] = pb;
    // count it in the lists of the blocks it can touch; block boxes of more than 4 blocks are left to
    // bin_blocks_kernel<false>, whose warps walk them together
    const int bw = bb.bx1 - bb.bx0 + 1, bh = bb.by1 - bb.by0 + 1;
    if (bw * bh <= 4) {
        int *cnt = bd.blk_cnt + (size_t)local_scene * (2 * f.nbx * f.nby);
        const int kind = pb != PBOX_NONE ? 1 : 0;
        const bool small = (bw - 1) + (bh - 1) <= 1;       // <= 2 blocks: no reject test
        for (int by = bb.by0; by <= bb.by1; ++by)
            for (int bx = bb.bx0; bx <= bb.bx1; ++bx)
                if (small || block_hit(r, bb, bx, by)) atomicAdd(&cnt[2 * (by * f.nbx + bx) + kind], 1);
    }
}

__device__ __noinline__ void bin_clipped(const FrameDev &f, const StagedDev &g, const BinnedDev &bd, int local_scene,
                                         SlotGeom sg) {
    CVT poly[MAX_POLY];
    const int n = clip_poly(sg.v, poly);
    for (int k = 0; k + 2 < n; ++k) {
        CVT tri[3] = {poly[0], poly[k + 1], poly[k + 2]};
        Rec r;
        TriVary tv;
        BBox bb;
        if (setup_tri(f, tri, sg, 0, f.H, r, bb, tv)) bin_append(f, g, bd, local_scene, r, bb, tri, sg, tv);
    }
}

__global__ void __launch_bounds__(B_THREADS, PBR_B_TRI_OCC) bin_tri_kernel(const __grid_constant__ FrameDev f,
                                                            const __grid_constant__ StagedDev g,
                                                            const __grid_constant__ BinnedDev bd) {
    __shared__ unsigned s_live[B_THREADS], s_clip[B_THREADS];
    __shared__ int s_n[2];
    const int local_scene = blockIdx.y;
    const int scene = g.scene0 + local_scene;
    const int tid = threadIdx.x, lane = tid & 31;
    const int slot = blockIdx.x * B_THREADS + tid;
    if (tid < 2) s_n[tid] = 0;
    __syncthreads();
    const float4 *vclip = bd.vclip + (size_t)local_scene * bd.total_verts;
    const int4 *vproj = bd.vproj + (size_t)local_scene * bd.total_verts;

    // ---- classify this thread's slot from its three parked vertices
    int cat = 0;            // 0 dead, 1 live, 2 clip
    if (slot < f.total_slots) {
        int ni, inst, tri;
        locate_slot(f, slot, ni, inst, tri);
        const NodeDev &nd = f.nodes[ni];
        if (g.vis[(size_t)local_scene * f.total_inst + nd.inst_begin + inst]) {
            const uint4 ti = __ldg(nd.tidx + tri);
            const int vb = nd.vert_begin + inst * nd.n_verts;
            const int4 q0 = vproj[vb + ti.x], q1 = vproj[vb + ti.y], q2 = vproj[vb + ti.z];
            const int f_and = q0.w & q1.w & q2.w, f_or = q0.w | q1.w | q2.w;
            if (f_and & 0x3f) {
                cat = 0;                                  // all three outside one plane of the tile frustum
            } else if (f_or & BV_CLIP) {
                cat = 2;
            } else if (f_and & BV_PROJ) {
                const long long area2 = (long long)(q1.x - q0.x) * (q2.y - q0.y) - (long long)(q2.x - q0.x) * (q1.y - q0.y);
                const bool two_sided = (nd.flags & PBR_MESH_TWO_SIDED) != 0;
                cat = (area2 < 0 || (two_sided && area2 > 0)) ? 1 : 0;
                if (cat == 1) {     // no pixel centre inside the bounding box: the set-up would drop it anyway
                    const int xmin = min(q0.x, min(q1.x, q2.x)), xmax = max(q0.x, max(q1.x, q2.x));
                    const int ymin = min(q0.y, min(q1.y, q2.y)), ymax = max(q0.y, max(q1.y, q2.y));
                    const int i0 = max(0, (xmin - 128 + 255) >> 8), i1 = min(f.W - 1, (xmax - 128) >> 8);
                    const int j0 = max(0, (ymin - 128 + 255) >> 8), j1 = min(f.H - 1, (ymax - 128) >> 8);
                    if (i0 > i1 || j0 > j1) cat = 0;
                }
            }
        }
    }
    // ---- compact the survivors of the CTA
    {
        const unsigned lt = (1u << lane) - 1u;
        const unsigned bl = __ballot_sync(0xffffffffu, cat == 1), bc = __ballot_sync(0xffffffffu, cat == 2);
        int pl = 0, pc = 0;
        if (lane == 0) {
            if (bl) pl = atomicAdd(&s_n[0], __popc(bl));
            if (bc) pc = atomicAdd(&s_n[1], __popc(bc));
        }
        pl = __shfl_sync(0xffffffffu, pl, 0);
        pc = __shfl_sync(0xffffffffu, pc, 0);
        if (cat == 1) s_live[pl + __popc(bl & lt)] = (unsigned)slot;
        if (cat == 2) s_clip[pc + __popc(bc & lt)] = (unsigned)slot;
    }
    __syncthreads();
    // ---- set-up on full warps
    const int n_live = s_n[0], n_clip = s_n[1];
    for (int i = tid; i < n_live; i += B_THREADS) {
        const int s = (int)s_live[i];
        int ni, inst, tri;
        locate_slot(f, s, ni, inst, tri);
        SlotGeom sg;
        load_slot_parked(f, scene, s, ni, inst, tri, vclip, sg);
        Rec r;
        TriVary tv;
        BBox bb;
        if (setup_tri(f, sg.v, sg, 0, f.H, r, bb, tv)) bin_append(f, g, bd, local_scene, r, bb, sg.v, sg, tv);
    }
    for (int i = tid; i < n_clip; i += B_THREADS) {
        const int s = (int)s_clip[i];
        int ni, inst, tri;
        locate_slot(f, s, ni, inst, tri);
        SlotGeom sg;
        load_slot_parked(f, scene, s, ni, inst, tri, vclip, sg);
        bin_clipped(f, g, bd, local_scene, sg);
    }
}

// ------------------------------------------------------------------------------------------------
// block lists
// ------------------------------------------------------------------------------------------------
// CTA per scene: blk_off = exclusive prefix sum of blk_cnt; blk_cnt is zeroed (it becomes the fill cursor)
__global__ void __launch_bounds__(B_THREADS) bin_scan_kernel(const __grid_constant__ FrameDev f,
                                                             const __grid_constant__ BinnedDev bd) {
    __shared__ int s_warp[B_THREADS / 32];
    __shared__ int s_carry;
    const int local_scene = blockIdx.x;
    const int nblk = 2 * f.nbx * f.nby;                 // lists, two per block
    int *cnt = bd.blk_cnt + (size_t)local_scene * nblk;
    int *off = bd.blk_off + (size_t)local_scene * (nblk + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += B_THREADS) {
        const int i = base + tid;
        const int c = i < nblk ? cnt[i] : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (i < nblk) { off[i] = before + incl - c; cnt[i] = 0; }
        __syncthreads();
        if (tid == B_THREADS - 1) s_carry = before + incl;
        __syncthreads();
    }
    if (tid == 0) {
        off[nblk] = s_carry;
        if (s_carry > bd.pairs_cap) {        // lists do not fit: flagged like a record overflow (the host then
            atomicOr(f.status, DEVSTAT_STAGED_OVERFLOW);      // falls back to the band-based path with worst-case sizes)
            f.status_host[1] = 1;
        }
    }
}

// Thread per record, run twice: FILL = false counts the record in every 8x8 block it can touch (block box of the
// record + edge-function reject per block) -- only records whose block box exceeds 4 blocks, the others were counted
// by bin_tri_kernel when they were made; FILL = true -- after the scan -- writes every record's index into those
// blocks' lists.  Boxes of up to 4 blocks are walked by the record's own thread; larger ones (a near triangle can
// span the whole tile: 256 .. 1024 blocks) are handed to the warp, whose lanes stride over the box together.
template <bool FILL>
__global__ void __launch_bounds__(B_THREADS) bin_blocks_kernel(const __grid_constant__ FrameDev f,
                                                               const __grid_constant__ StagedDev g,
                                                               const __grid_constant__ BinnedDev bd) {
    const int local_scene = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int total = min(g.count[local_scene], g.cap);
    const int nlist = 2 * f.nbx * f.nby;
    int *cur = bd.blk_cnt + (size_t)local_scene * nlist;
    const int *off = bd.blk_off + (size_t)local_scene * (nlist + 1);
    unsigned *pairs = bd.pairs + (size_t)local_scene * bd.pairs_cap;
    auto visit = [&](const Rec &r, const BBox &bb, bool small, int bx, int by, int kind, unsigned ridx) {
        if (small || block_hit(r, bb, bx, by)) {
            const int l = 2 * (by * f.nbx + bx) + kind;
            const int p = atomicAdd(&cur[l], 1);
            if (FILL && off[l] + p < bd.pairs_cap) pairs[off[l] + p] = ridx;
        }
    };
    // (the grid covers a fraction of the list's capacity: the lists are usually far shorter than that)
#pragma unroll 1
    for (int cta_base = blockIdx.x * B_THREADS; cta_base < total; cta_base += gridDim.x * B_THREADS) {
    const int idx = cta_base + threadIdx.x;
    Rec r;
    BBox bb;
    bb.bx0 = bb.by0 = 0; bb.bx1 = bb.by1 = -1;
    int kind = 0;
    if (idx < total) {
        const size_t o = (size_t)local_scene * g.cap + idx;
        const uint4 *src = reinterpret_cast<const uint4 *>(g.recs + o);
        uint4 *dst = reinterpret_cast<uint4 *>(&r);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        const unsigned pb = g.bbox[o];
        bb.bx0 = (int)(pb & 255u); bb.by0 = (int)((pb >> 8) & 255u);
        bb.bx1 = (int)((pb >> 16) & 255u); bb.by1 = (int)(pb >> 24);
        kind = bd.pbox[o] != PBOX_NONE ? 1 : 0;
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) r.e[i] = 0;
        r.meta = 0;
    }
    const int bw = bb.bx1 - bb.bx0 + 1, bh = bb.by1 - bb.by0 + 1;
    const int area = bw * bh;
    if (FILL && area > 0 && area <= 4) {                    // (counted by bin_tri_kernel already)
        const bool small = (bw - 1) + (bh - 1) <= 1;
        for (int by = bb.by0; by <= bb.by1; ++by)
            for (int bx = bb.bx0; bx <= bb.bx1; ++bx) visit(r, bb, small, bx, by, kind, (unsigned)idx);
    }
    unsigned big = __ballot_sync(0xffffffffu, area > 4);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        Rec q;
#pragma unroll
        for (int i = 0; i < 9; ++i) q.e[i] = __shfl_sync(0xffffffffu, r.e[i], src);
        q.meta = __shfl_sync(0xffffffffu, r.meta, src);
        BBox qb;
        qb.bx0 = __shfl_sync(0xffffffffu, bb.bx0, src); qb.by0 = __shfl_sync(0xffffffffu, bb.by0, src);
        qb.bx1 = __shfl_sync(0xffffffffu, bb.bx1, src); qb.by1 = __shfl_sync(0xffffffffu, bb.by1, src);
        const int qkind = __shfl_sync(0xffffffffu, kind, src);
        const unsigned qidx = (unsigned)(cta_base + (threadIdx.x & ~31) + src);
        const int qw = qb.bx1 - qb.bx0 + 1, qa = qw * (qb.by1 - qb.by0 + 1);
        for (int k = lane; k < qa; k += 32) {
            const int yy = k / qw;
            visit(q, qb, false, qb.bx0 + k - yy * qw, qb.by0 + yy, qkind, qidx);
        }
    }
    }
}

// ------------------------------------------------------------------------------------------------
// raster: one warp per (scene, 8x8 block)
// ------------------------------------------------------------------------------------------------
template <bool SMOOTH>
__global__ void __launch_bounds__(B_WPB * 32, PBR_B_RASTER_OCC) raster_binned_kernel(const __grid_constant__ FrameDev f,
                                                                      const __grid_constant__ StagedDev g,
                                                                      const __grid_constant__ BinnedDev bd) {
    __shared__ __align__(16) Rec s_recs[B_WPB][2][B_GATHER];
    __shared__ unsigned long long s_key[B_WPB][64];        // small records: depth|id of the block's pixels
    __shared__ unsigned s_col[B_WPB][64];                  // ... colour of the current winner
    __shared__ unsigned s_tag[SMOOTH ? B_WPB : 1][64];     // ... its record index if it is shaded per pixel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int local_scene = blockIdx.y;
    const int scene = g.scene0 + local_scene;
    const int nblk = f.nbx * f.nby;
    const int blk = blockIdx.x * B_WPB + warp;
    if (blk >= nblk) return;                                  // (no CTA-wide barrier below)
    const int by = fast_div(blk, f.nbx_magic), bx = blk - by * f.nbx;
    const int px = bx * 8 + (lane & 7), py0 = by * 8 + (lane >> 3);
    const bool ok0 = px < f.W && py0 < f.H, ok1 = px < f.W && py0 + 4 < f.H;
    const int *off = bd.blk_off + (size_t)local_scene * (2 * nblk + 1);
    const int begin = min(off[2 * blk], bd.pairs_cap), end = min(off[2 * blk + 1], bd.pairs_cap);     // big records
    const int send = min(off[2 * blk + 2], bd.pairs_cap);                                             // small: [end, send)
    const unsigned *pairs = bd.pairs + (size_t)local_scene * bd.pairs_cap;
    const Rec *grecs = g.recs + (size_t)local_scene * g.cap;
    const unsigned char *gsrecs = SMOOTH ? g.srecs + (size_t)local_scene * g.cap * f.srec_stride : nullptr;

    PixelState ps;
    ps.k0 = ps.k1 = KEY_CLEAR;
    ps.c0 = ps.c1 = f.bg;
    // SMOOTH: records shaded per pixel are not shaded when they win (most wins are overwritten again, and every
    // one would cost a fragment-shader evaluation for the whole warp) -- the pixel remembers the record and the
    // fragment shader runs once, after the sweep, for the final winner
    constexpr unsigned NO_REC = 0xffffffffu;
    unsigned win0 = NO_REC, win1 = NO_REC;
    // Records travel list -> shared memory eight at a time (lane l fetches quarter (l & 3) of record (l >> 2): one
    // 16-byte load per lane per round), double buffered: the loads of round i + 1 are in flight while round i is
    // swept, so the two dependent L2 round trips (index, then record) are not on the warp's critical path.
    const int slot = lane >> 2;
    unsigned ridx = 0, ridx_next = 0;
    uint4 q_next = make_uint4(0u, 0u, 0u, 0u);
    auto fetch = [&](int base) {
        if (base + slot < end) {
            ridx_next = __ldg(pairs + base + slot);
            q_next = __ldg(reinterpret_cast<const uint4 *>(grecs + ridx_next) + (lane & 3));
        }
    };
    fetch(begin);
    int buf = 0;
#pragma unroll 1
    for (int base = begin; base < end; base += B_GATHER) {
        const int n = min(B_GATHER, end - base);
        Rec *mine = s_recs[warp][buf];
        ridx = ridx_next;
        if (slot < n) reinterpret_cast<uint4 *>(mine + slot)[lane & 3] = q_next;
        fetch(base + B_GATHER);
        buf ^= 1;
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const Rec &r = mine[k];
            const int4 ea = *reinterpret_cast<const int4 *>(&r.e[0]);
            const int4 eb = *reinterpret_cast<const int4 *>(&r.e[4]);
            const int4 ec = *reinterpret_cast<const int4 *>(&r.e[8]);
            const float4 zq = *reinterpret_cast<const float4 *>(&r.z0);
            bool w0, w1;
            if (!((unsigned)ec.w & M_SLOW)) {
                const FastCov c = fast_cover(ea, eb, ec, px, py0, ok0, ok1);
                if (!__any_sync(0xffffffffu, c.cov0 || c.cov1)) continue;
                depth_update(ps, (float)c.F1, (float)c.F2, (float)c.G1, (float)c.G2, c.cov0, c.cov1, zq, (unsigned)ec.z,
                             (unsigned)ec.y, w0, w1);
            } else {
                bool cov0, cov1;
                float f0a, f1a, f2a, f0b, f1b, f2b;
                const bool any = slow_cover(r, px, py0, ok0, ok1, cov0, cov1, f1a, f2a, f1b, f2b, f0a, f0b);
                if (!__any_sync(0xffffffffu, any)) continue;
                depth_update(ps, f1a, f2a, f1b, f2b, cov0, cov1, zq, (unsigned)ec.z, (unsigned)ec.y, w0, w1);
            }
            if (SMOOTH) {
                const unsigned ri = __shfl_sync(0xffffffffu, ridx, k * 4);
                const unsigned tag = ((unsigned)ec.w & M_SMOOTH) ? ri : NO_REC;
                win0 = w0 ? tag : win0;
                win1 = w1 ? tag : win1;
            }
        }
    }
    // ---- small records (pixel box <= 4 x 4, int32 edge functions): ONE LANE PER RECORD.  A triangle that can cover
    // a handful of pixels would keep a whole warp busy for ~60 instructions in the sweep above; here 32 of them are
    // evaluated at once, each lane walking the pixels of its record's box that lie in this block.  Visibility is
    // resolved through the block's depth|id keys in shared memory (64-bit atomicMin: smaller (depth, id) wins,
    // whatever the order), then the lanes whose key survived write their colour, and the warp merges the result
    // with the keys of the big records in its registers.
    if (send > end) {
        s_key[warp][lane] = KEY_CLEAR;
        s_key[warp][lane + 32] = KEY_CLEAR;
        __syncwarp();
        const int ox = bx * 8, oy = by * 8;
#pragma unroll 1
        for (int base = end; base < send; base += 32) {
            const int i = base + lane;
            const bool valid = i < send;
            unsigned ridx = 0, hits = 0;
            int4 ea = make_int4(0, 0, 0, 0), eb = ea, ec = ea;
            float4 zq = make_float4(0.f, 0.f, 0.f, 0.f);
            int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
            if (valid) {
                ridx = __ldg(pairs + i);
                const unsigned pb = __ldg(bd.pbox + (size_t)local_scene * g.cap + ridx);
                const Rec *rp = grecs + ridx;
                ea = __ldg(reinterpret_cast<const int4 *>(&rp->e[0]));
                eb = __ldg(reinterpret_cast<const int4 *>(&rp->e[4]));
                ec = __ldg(reinterpret_cast<const int4 *>(&rp->e[8]));
                zq = __ldg(reinterpret_cast<const float4 *>(&rp->z0));
                const int bx0 = (int)(pb & 2047u), by0 = (int)((pb >> 11) & 2047u);
                x0 = max(bx0, ox); x1 = min(bx0 + (int)((pb >> 22) & 31u), ox + 7);
                y0 = max(by0, oy); y1 = min(by0 + (int)(pb >> 27), oy + 7);
            }
            const unsigned meta = (unsigned)ec.w;
            const int nb1 = meta_nb1(meta), nb2 = meta_nb2(meta);
            // biased edge values + depth of the sample at pixel (xx, yy): same integer / float operations as fast_cover
            // and depth_update
            auto sample = [&](int xx, int yy, unsigned long long &key) -> bool {
                const unsigned rx = (unsigned)xx, ry = (unsigned)yy;
                const int F0 = (int)((unsigned)ea.x + (unsigned)ea.w * rx + (unsigned)eb.z * ry);
                const int F1 = (int)((unsigned)ea.y + (unsigned)eb.x * rx + (unsigned)eb.w * ry);
                const int F2 = (int)((unsigned)ea.z + (unsigned)eb.y * rx + (unsigned)ec.x * ry);
                if ((F0 | F1 | F2) < 0) return false;
                const float z = fmaf((float)(F2 + nb2) * zq.w, zq.z, fmaf((float)(F1 + nb1) * zq.w, zq.y, zq.x));
                key = make_key(z, (unsigned)ec.z);
                return true;
            };
            {
                unsigned bit = 1u;
                for (int yy = y0; yy <= y1; ++yy)
                    for (int xx = x0; xx <= x1; ++xx, bit <<= 1) {
                        unsigned long long key;
                        if (sample(xx, yy, key)) {
                            atomicMin(&s_key[warp][(yy - oy) * 8 + (xx - ox)], key);
                            hits |= bit;
                        }
                    }
            }
            __syncwarp();
            if (hits) {
                unsigned bit = 1u;
                for (int yy = y0; yy <= y1; ++yy)
                    for (int xx = x0; xx <= x1; ++xx, bit <<= 1)
                        if (hits & bit) {
                            unsigned long long key;
                            sample(xx, yy, key);
                            const int p = (yy - oy) * 8 + (xx - ox);
                            if (s_key[warp][p] == key) {
                                s_col[warp][p] = (unsigned)ec.y;
                                if (SMOOTH) s_tag[warp][p] = (meta & M_SMOOTH) ? ridx : NO_REC;
                            }
                        }
            }
            __syncwarp();
        }
        // merge with the big records' keys (this lane's pixels are entries lane and lane + 32 of the block)
        const unsigned long long k0 = s_key[warp][lane], k1 = s_key[warp][lane + 32];
        if (ok0 && k0 < ps.k0) { ps.k0 = k0; ps.c0 = s_col[warp][lane]; if (SMOOTH) win0 = s_tag[warp][lane]; }
        if (ok1 && k1 < ps.k1) { ps.k1 = k1; ps.c1 = s_col[warp][lane + 32]; if (SMOOTH) win1 = s_tag[warp][lane + 32]; }
    }
    if (SMOOTH && __any_sync(0xffffffffu, win0 != NO_REC || win1 != NO_REC)) {
        // fragment shader of the final winners (basic.frag:31-38 with interpolated normal / uv): every lane fetches
        // the record and the shading inputs of its own two pixels
        auto shade_final = [&](unsigned ri, bool second) -> unsigned {
            const Rec *rp = grecs + ri;
            const int4 ea = __ldg(reinterpret_cast<const int4 *>(&rp->e[0]));
            const int4 eb = __ldg(reinterpret_cast<const int4 *>(&rp->e[4]));
            const int4 ec = __ldg(reinterpret_cast<const int4 *>(&rp->e[8]));
            const float invA = __ldg(&rp->invA);
            const SRec *sr = reinterpret_cast<const SRec *>(gsrecs + (size_t)ri * f.srec_stride);
            float f0, f1, f2;
            if (!((unsigned)ec.w & M_SLOW)) {
                const FastCov c = fast_cover(ea, eb, ec, px, py0, true, true);
                f0 = (float)(second ? c.G0 : c.F0); f1 = (float)(second ? c.G1 : c.F1); f2 = (float)(second ? c.G2 : c.F2);
            } else {
                Rec r;
                *reinterpret_cast<int4 *>(&r.e[0]) = ea; *reinterpret_cast<int4 *>(&r.e[4]) = eb;
                *reinterpret_cast<int4 *>(&r.e[8]) = ec;
                bool c0, c1;
                float f0a, f1a, f2a, f0b, f1b, f2b;
                slow_cover(r, px, py0, true, true, c0, c1, f1a, f2a, f1b, f2b, f0a, f0b);
                f0 = second ? f0b : f0a; f1 = second ? f1b : f1a; f2 = second ? f2b : f2a;
            }
            return shade_pixel(f, *sr, ((unsigned)ec.w & M_TEX) != 0, f0, f1, f2, invA);
        };
        if (win0 != NO_REC) ps.c0 = shade_final(win0, false);
        if (win1 != NO_REC) ps.c1 = shade_final(win1, true);
    }
    // the finished block, straight to out[scene] (background where nothing was drawn)
    const int HW = f.H * f.W;
    unsigned char *p = f.out + (size_t)scene * f.C * HW + (size_t)py0 * f.W + px;
    if (ok0) {
        p[0] = (unsigned char)(ps.c0 & 255u);
        p[HW] = (unsigned char)((ps.c0 >> 8) & 255u);
        p[2 * HW] = (unsigned char)((ps.c0 >> 16) & 255u);
        if (f.C == 4) p[3 * HW] = (unsigned char)(ps.c0 >> 24);
    }
    if (ok1) {
        p += 4 * f.W;
        p[0] = (unsigned char)(ps.c1 & 255u);
        p[HW] = (unsigned char)((ps.c1 >> 8) & 255u);
        p[2 * HW] = (unsigned char)((ps.c1 >> 16) & 255u);
        if (f.C == 4) p[3 * HW] = (unsigned char)(ps.c1 >> 24);
    }
}

}  // namespace pbr
