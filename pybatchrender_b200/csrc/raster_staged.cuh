// raster_staged.cuh -- the two-kernel path for large scenes (hundreds to thousands of triangle
// slots per scene, tiles that need several bands):
//
//   geom_kernel           the batched vertex / instance transform kernel: one thread per
//                         (scene, triangle slot).  clip = VP*(M*v) (basic.vert:24-43), trivial
//                         reject, near-plane / guard-band clipping into fan triangles, projection,
//                         snap, back-face cull, integer edge setup, flat shading (basic.frag:31-38).
//                         Survivors are appended to the scene's record list in global memory
//                         (64-byte records + a packed block bounding box), warp-aggregated.
//   raster_staged_kernel  one CTA per (scene, band of rows).  The scene's records are staged into
//                         shared memory 128 at a time with TMA bulk copies (cp.async.bulk +
//                         mbarrier, double buffered: the next chunk lands while the current one is
//                         binned and rasterised), binned into per-8x8-block bitmasks, swept by the
//                         warps (depth|id keys of the band in shared memory), and the finished band
//                         is written with 128-bit streaming stores into out[scene].
//
// Compared with raster_general_kernel (which redoes the geometry in every band and every pass) the
// geometry is done exactly once per frame.  Depth is order independent (key = depth bits | draw id)
// so the arbitrary order of the appended records does not matter.
#pragma once
#include "common.cuh"
#include "raster_general.cuh"

namespace pbr {

constexpr int G_THREADS = 256;

struct StagedDev {
    Rec *recs;               // [scenes_in_launch][cap]
    unsigned char *srecs;    // [scenes_in_launch][cap] x srec_stride bytes, or NULL (all triangles flat)
    unsigned *bbox;          // [scenes_in_launch][cap]  bx0 | by0 << 8 | bx1 << 16 | by1 << 24 (tile blocks)
    int *count;              // [scenes_in_launch]
    unsigned char *vis;      // [scenes_in_launch][total_inst] 1 = instance may touch a pixel (cull_kernel)
    // per-band index lists (tiles with several bands): which records touch the band, so that a band's
    // CTA stages only those instead of scanning the whole scene.  NULL = one list per scene.
    int *bcount;             // [scenes_in_launch][nbands]
    unsigned *bidx;          // [scenes_in_launch][nbands][cap] indices into the scene's records
    int cap;                 // records per scene (multiple of 4)
    int scene0;              // first scene of this launch (index into vp / out / per-scene rows)
};

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void append_record(const StagedDev &g, const FrameDev &f, int local_scene, const Rec &r,
                                              const BBox &bb, const CVT *vin, const SlotGeom &sg, const TriVary &tv) {
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
    if (g.bidx != nullptr) {
        const int b0 = bb.by0 / f.nby, b1 = bb.by1 / f.nby;             // bands are f.nby block rows high
        for (int band = b0; band <= b1; ++band) {
            const size_t l = (size_t)local_scene * f.nbands + band;
            const int pos = atomicAdd(&g.bcount[l], 1);                 // < cap: a record enters a band's list once
            g.bidx[l * g.cap + pos] = (unsigned)idx;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// instance culling: one thread per (scene, instance).  The mesh's bounding sphere is pushed through
// M and VP with interval arithmetic; an instance is dropped when every point of the sphere is outside
// one clip plane (all its triangles would be rejected by trivially_outside) or when the pixel
// rectangle that bounds its projection contains no pixel centre (no triangle inside it can cover a
// sample).  Both tests are conservative (margins far above fp32 rounding and the 1/256 px snap), so
// the image is unchanged; what is saved is the per-triangle transform / setup of instances that
// are off screen or smaller than the pixel grid (distant obstacles of Steering-v0).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cull_kernel(const __grid_constant__ FrameDev f,
                                                   const __grid_constant__ StagedDev g) {
    const int local_scene = blockIdx.y;
    const int scene = g.scene0 + local_scene;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= f.total_inst) return;
    int ni = 0;
#pragma unroll 1
    for (int i = 1; i < f.n_nodes; ++i)
        if (k >= f.nodes[i].inst_begin) ni = i;
    const NodeDev &nd = f.nodes[ni];
    const int inst = k - nd.inst_begin;
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
    float world[4], cc[4];
    mat_vec4(M, nd.bsphere.x, nd.bsphere.y, nd.bsphere.z, 1.0f, world);
    mat_vec4(VP, world[0], world[1], world[2], world[3], cc);
    // radius in world space: largest stretch of A = mat3(M) is sqrt(lambda_max(A^T A)), bounded by
    // the largest absolute row sum of A^T A (exact for rotation * uniform scale)
    float gmax = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float rowsum = 0.0f;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            rowsum += fabsf(fmaf(M[4 * i], M[4 * j], fmaf(M[4 * i + 1], M[4 * j + 1], M[4 * i + 2] * M[4 * j + 2])));
        gmax = fmaxf(gmax, rowsum);
    }
    const float rw = nd.bsphere.w * sqrtf(gmax) * 1.0001f;
    float e[4];                        // how far each clip coordinate can move inside the sphere
    bool finite = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float row = sqrtf(fmaf(VP[i], VP[i], fmaf(VP[4 + i], VP[4 + i], VP[8 + i] * VP[8 + i])));
        e[i] = fmaf(rw, row, 1e-4f * (fabsf(cc[i]) + rw * row)) + 1e-30f;
        finite &= (fabsf(cc[i]) < 1e30f) && (e[i] < 1e30f);
    }
    bool visible = true;
    if (finite && world[3] == 1.0f) {
        // clip planes: a >= -w and a <= w for x, y, z
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if ((cc[3] + cc[a]) + (e[3] + e[a]) < 0.0f) visible = false;
            if ((cc[3] - cc[a]) + (e[3] + e[a]) < 0.0f) visible = false;
        }
        const float wl = cc[3] - e[3], wh = cc[3] + e[3];
        if (visible && wl > 1e-20f) {
            const float xl = cc[0] - e[0], xh = cc[0] + e[0], yl = cc[1] - e[1], yh = cc[1] + e[1];
            const float nxl = xl / (xl >= 0.0f ? wh : wl), nxh = xh / (xh >= 0.0f ? wl : wh);
            const float nyl = yl / (yl >= 0.0f ? wh : wl), nyh = yh / (yh >= 0.0f ? wl : wh);
            const float mx = 0.0625f + 1e-4f * (float)f.W, my = 0.0625f + 1e-4f * (float)f.H;
            const float pxl = fmaf(nxl, f.hw, f.hw) - mx, pxh = fmaf(nxh, f.hw, f.hw) + mx;
            const float pyl = fmaf(-nyh, f.hh, f.hh) - my, pyh = fmaf(-nyl, f.hh, f.hh) + my;
            // pixel i is sampled at i + 0.5
            const float ilo = fmaxf(ceilf(pxl - 0.5f), 0.0f), ihi = fminf(floorf(pxh - 0.5f), (float)(f.W - 1));
            const float jlo = fmaxf(ceilf(pyl - 0.5f), 0.0f), jhi = fminf(floorf(pyh - 0.5f), (float)(f.H - 1));
            if (ilo > ihi || jlo > jhi) visible = false;
        }
    }
    g.vis[(size_t)local_scene * f.total_inst + k] = visible ? 1 : 0;
}

// The rare path (triangle crosses the near plane or the guard band), out of line and with the slot
// passed by value: its arrays live in local memory, and keeping them out of geom_kernel's body keeps
// the common path's vertices in registers.
__device__ __noinline__ void geom_clipped(const FrameDev &f, const StagedDev &g, int local_scene, SlotGeom sg) {
    CVT poly[MAX_POLY];
    const int n = clip_poly(sg.v, poly);
    for (int k = 0; k + 2 < n; ++k) {
        CVT tri[3] = {poly[0], poly[k + 1], poly[k + 2]};
        Rec r;
        TriVary tv;
        BBox bb;
        if (setup_tri(f, tri, sg, 0, f.H, r, bb, tv)) append_record(g, f, local_scene, r, bb, tri, sg, tv);
    }
}

__global__ void __launch_bounds__(G_THREADS) geom_kernel(const __grid_constant__ FrameDev f,
                                                         const __grid_constant__ StagedDev g) {
    const int local_scene = blockIdx.y;
    const int scene = g.scene0 + local_scene;
    const int slot = blockIdx.x * G_THREADS + threadIdx.x;
    if (slot >= f.total_slots) return;
    int ni, inst, tri;
    locate_slot(f, slot, ni, inst, tri);
    if (!g.vis[(size_t)local_scene * f.total_inst + f.nodes[ni].inst_begin + inst]) return;
    SlotGeom sg;
    const int st = load_slot(f, scene, slot, ni, inst, tri, sg);
    if (st == SLOT_SKIP) return;
    Rec r;
    TriVary tv;
    BBox bb;
    if (st == SLOT_OK) {
        if (setup_tri(f, sg.v, sg, 0, f.H, r, bb, tv)) append_record(g, f, local_scene, r, bb, sg.v, sg, tv);
        return;
    }
    geom_clipped(f, g, local_scene, sg);
}

// ------------------------------------------------------------------------------------------------
// raster
// ------------------------------------------------------------------------------------------------
struct StagedSmem {
    unsigned char *color;          // [C][plane_stride]
    unsigned long long *ktile;     // [nblk*64]
    Rec *recs[2];                  // 2 x [CH]
    unsigned char *srecs[2];       // 2 x [CH] x srec_stride bytes (per-pixel shading frames only)
    unsigned *bbox[2];             // 2 x [CH]
    unsigned *masks;               // [nblk*MW]
    unsigned short *blist;         // [nblk]
    int *ctr;                      // nlist, next
    unsigned long long *bar;       // 2 mbarriers
};

__host__ __device__ inline size_t staged_smem_bytes(int C, int plane_stride, int nblk, int srec_bytes) {
    return 2 * (size_t)CH * srec_bytes + align16((size_t)C * plane_stride) + (size_t)nblk * 64 * 8 + 2 * (size_t)CH * sizeof(Rec) + 2 * (size_t)CH * 4 +
           align16((size_t)nblk * MW * 4) + align16((size_t)nblk * 2) + 16 + 16;
}

__device__ __forceinline__ StagedSmem staged_carve(unsigned char *base, int C, int plane_stride, int nblk, int srec_bytes) {
    StagedSmem s;
    s.color = base; base += align16((size_t)C * plane_stride);
    s.ktile = reinterpret_cast<unsigned long long *>(base); base += (size_t)nblk * 64 * 8;
    s.recs[0] = reinterpret_cast<Rec *>(base); base += (size_t)CH * sizeof(Rec);
    s.recs[1] = reinterpret_cast<Rec *>(base); base += (size_t)CH * sizeof(Rec);
    s.srecs[0] = base; base += (size_t)CH * srec_bytes;
    s.srecs[1] = base; base += (size_t)CH * srec_bytes;
    s.bbox[0] = reinterpret_cast<unsigned *>(base); base += (size_t)CH * 4;
    s.bbox[1] = reinterpret_cast<unsigned *>(base); base += (size_t)CH * 4;
    s.masks = reinterpret_cast<unsigned *>(base); base += align16((size_t)nblk * MW * 4);
    s.blist = reinterpret_cast<unsigned short *>(base); base += align16((size_t)nblk * 2);
    s.ctr = reinterpret_cast<int *>(base); base += 16;
    s.bar = reinterpret_cast<unsigned long long *>(base);
    return s;
}

template <bool SMOOTH>
__global__ void __launch_bounds__(THREADS) raster_staged_kernel(const __grid_constant__ FrameDev f,
                                                                const __grid_constant__ StagedDev g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int local_scene = (int)(blockIdx.x / f.nbands);
    const int scene = g.scene0 + local_scene;
    const int band = (int)(blockIdx.x % f.nbands);
    const int band_y0 = band * f.BH;
    const int band_h = min(f.BH, f.H - band_y0);
    const int band_by0 = band_y0 / 8;
    const int nblk = f.nbx * f.nby;
    const StagedSmem s = staged_carve(smem_raw, f.C, f.plane_stride, nblk, SMOOTH ? f.srec_stride : 0);

    const bool gather = g.bidx != nullptr;
    const size_t blist_row = (size_t)local_scene * f.nbands + band;
    const int total = gather ? min(g.bcount[blist_row], g.cap) : min(g.count[local_scene], g.cap);
    const Rec *grecs = g.recs + (size_t)local_scene * g.cap;
    const unsigned *gbbox = g.bbox + (size_t)local_scene * g.cap;
    const unsigned *glist = gather ? g.bidx + blist_row * g.cap : nullptr;
    const int nchunks = (total + CH - 1) / CH;

    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // stage chunk c into buffer c & 1.  One list per scene: thread 0 issues three bulk copies of contiguous
    // ranges.  Per-band lists: thread i gathers record list[c * CH + i] with its own 64-byte bulk copy
    // (thread 0 arms the mbarrier with the total; completions that overtake it only drive the pending
    // byte count negative for a moment, the phase cannot complete before its arrival).
    auto issue = [&](int c) {
        const int cnt = min(CH, total - c * CH);
        const unsigned rb = (unsigned)cnt * (unsigned)sizeof(Rec);
        const unsigned sb = SMOOTH ? (unsigned)cnt * (unsigned)f.srec_stride : 0u;
        unsigned long long *bar = &s.bar[c & 1];
        if (gather) {
            if (tid == 0) mbar_expect_tx(bar, rb + sb);
            if (tid < cnt) {
                const unsigned ridx = glist[(size_t)c * CH + tid];
                tma_load(&s.recs[c & 1][tid], grecs + ridx, (unsigned)sizeof(Rec), bar);
                if (SMOOTH)
                    tma_load(s.srecs[c & 1] + (size_t)tid * f.srec_stride,
                             g.srecs + ((size_t)local_scene * g.cap + ridx) * f.srec_stride, (unsigned)f.srec_stride, bar);
                s.bbox[c & 1][tid] = __ldg(gbbox + ridx);       // read back by this thread only
            }
        } else if (tid == 0) {
            const unsigned bb = (unsigned)align16((size_t)cnt * 4);
            mbar_expect_tx(bar, rb + sb + bb);
            tma_load(s.recs[c & 1], grecs + (size_t)c * CH, rb, bar);
            if (SMOOTH) tma_load(s.srecs[c & 1], g.srecs + ((size_t)local_scene * g.cap + (size_t)c * CH) * f.srec_stride, sb, bar);
            tma_load(s.bbox[c & 1], gbbox + (size_t)c * CH, bb, bar);
        }
    };
    if (nchunks > 0) issue(0);

    clear_color(f, s.color, tid, THREADS);
    {
        uint4 *kt = reinterpret_cast<uint4 *>(s.ktile);
        const uint4 clr = make_uint4(0u, 0x3F800000u, 0u, 0x3F800000u);
        for (int i = tid; i < nblk * 32; i += THREADS) kt[i] = clr;
    }

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        __syncthreads();                         // everyone is done sweeping the previous chunk
        for (int i = tid; i < nblk * MW; i += THREADS) s.masks[i] = 0;
        if (tid == 0) { s.ctr[0] = 0; s.ctr[1] = 0; }
        __syncthreads();                         // masks cleared; buffer buf^1 no longer read by anyone
        const int cnt = min(CH, total - c * CH);
        if (c + 1 < nchunks) issue(c + 1);
        // one lane per warp polls the mbarrier, then the warp reconverges: lanes leaving a spin loop
        // at different times would reach the aligned __syncthreads below diverged
        if (lane == 0) mbar_wait(&s.bar[buf], (unsigned)((c >> 1) & 1));
        __syncwarp();
        const Rec *recs = s.recs[buf];

        // bin: one record per thread, only the block rows of this band
        if (tid < cnt) {
            const unsigned pb = s.bbox[buf][tid];
            BBox bb;
            bb.bx0 = (int)(pb & 255u); bb.by0 = (int)((pb >> 8) & 255u);
            bb.bx1 = (int)((pb >> 16) & 255u); bb.by1 = (int)(pb >> 24);
            const int y0 = max(bb.by0, band_by0), y1 = min(bb.by1, band_by0 + f.nby - 1);
            if (y0 <= y1) {
                const Rec &r = recs[tid];
                const unsigned bit = rec_bit(tid);
                const int word = tid >> 5;
                const bool small = (bb.bx1 - bb.bx0) + (bb.by1 - bb.by0) <= 1;
                for (int by = y0; by <= y1; ++by)
                    for (int bx = bb.bx0; bx <= bb.bx1; ++bx)
                        if (small || block_hit(r, bb, bx, by))
                            atomicOr(&s.masks[((by - band_by0) * f.nbx + bx) * MW + word], bit);
            }
        }
        __syncthreads();

        // list of non-empty blocks, then the warps sweep them
        for (int b = tid; b < nblk; b += THREADS) {
            unsigned any = 0;
#pragma unroll
            for (int w = 0; w < MW; ++w) any |= s.masks[b * MW + w];
            if (any) s.blist[atomicAdd(&s.ctr[0], 1)] = (unsigned short)b;
        }
        __syncthreads();
        const int nlist = s.ctr[0];
        while (true) {
            int i = 0;
            if (lane == 0) i = atomicAdd(&s.ctr[1], 1);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= nlist) break;
            const int b = s.blist[i];
            const int byl = fast_div(b, f.nbx_magic), bxl = b - byl * f.nbx;
            const int px = bxl * 8 + (lane & 7);
            const int py0 = band_y0 + byl * 8 + (lane >> 3), py1 = py0 + 4;      // tile-global rows
            const bool ok0 = px < f.W && py0 < band_y0 + band_h, ok1 = px < f.W && py1 < band_y0 + band_h;
            PixelState ps;
            ps.k0 = s.ktile[b * 64 + lane];
            ps.k1 = s.ktile[b * 64 + 32 + lane];
            ps.c0 = ps.c1 = 0;
            const unsigned id0 = (unsigned)ps.k0, id1 = (unsigned)ps.k1;
            raster_block<MW, SMOOTH>(recs, s.masks + b * MW, px, py0, ok0, ok1, ps, &f, s.srecs[buf]);
            if (key_changed(ps.k0, id0)) {
                s.ktile[b * 64 + lane] = ps.k0;
                put_pixel(s.color, f.plane_stride, f.C, f.W, px, py0 - band_y0, ps.c0);
            }
            if (key_changed(ps.k1, id1)) {
                s.ktile[b * 64 + 32 + lane] = ps.k1;
                put_pixel(s.color, f.plane_stride, f.C, f.W, px, py1 - band_y0, ps.c1);
            }
        }
        // the __syncthreads at the top of the next iteration orders these reads before the TMA
        // write into this buffer two chunks later
    }
    __syncthreads();
    store_band(f, s.color, scene, band_y0, band_h, tid, THREADS);
}

}  // namespace pbr
