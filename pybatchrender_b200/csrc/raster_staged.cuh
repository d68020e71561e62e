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
    SRec *srecs;             // [scenes_in_launch][cap] smooth-shading companions, or NULL (all meshes flat)
    unsigned *bbox;          // [scenes_in_launch][cap]  bx0 | by0 << 8 | bx1 << 16 | by1 << 24 (tile blocks)
    int *count;              // [scenes_in_launch]
    int cap;                 // records per scene (multiple of 4)
    int scene0;              // first scene of this launch (index into vp / out / per-scene rows)
};

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void append_record(const StagedDev &g, const FrameDev &f, int local_scene, const Rec &r,
                                              const BBox &bb, const SRec &sr) {
    const int idx = atomicAdd(&g.count[local_scene], 1);
    if (idx >= g.cap) {
        atomicOr(f.status, DEVSTAT_STAGED_OVERFLOW);
        return;
    }
    const size_t o = (size_t)local_scene * g.cap + idx;
    uint4 *dst = reinterpret_cast<uint4 *>(g.recs + o);
    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    g.bbox[o] = (unsigned)bb.bx0 | ((unsigned)bb.by0 << 8) | ((unsigned)bb.bx1 << 16) | ((unsigned)bb.by1 << 24);
    if (g.srecs != nullptr && (r.meta & M_SMOOTH)) {
        uint4 *sd = reinterpret_cast<uint4 *>(g.srecs + o);
        const uint4 *ss = reinterpret_cast<const uint4 *>(&sr);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(SRec) / 16); ++i) sd[i] = ss[i];
    }
}

__global__ void __launch_bounds__(G_THREADS) geom_kernel(const __grid_constant__ FrameDev f,
                                                         const __grid_constant__ StagedDev g) {
    const int local_scene = blockIdx.y;
    const int scene = g.scene0 + local_scene;
    const int slot = blockIdx.x * G_THREADS + threadIdx.x;
    if (slot >= f.total_slots) return;
    SlotGeom sg;
    const int st = load_slot(f, scene, slot, sg);
    if (st == SLOT_SKIP) return;
    Rec r;
    SRec sr;
    BBox bb;
    if (st == SLOT_OK) {
        if (setup_tri(f, sg.v, sg, 0, f.H, r, bb, &sr)) append_record(g, f, local_scene, r, bb, sr);
        return;
    }
    CVT poly[MAX_POLY];
    const int n = clip_poly(sg.v, poly);
    for (int k = 0; k + 2 < n; ++k) {
        CVT tri[3] = {poly[0], poly[k + 1], poly[k + 2]};
        if (setup_tri(f, tri, sg, 0, f.H, r, bb, &sr)) append_record(g, f, local_scene, r, bb, sr);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (sm_90+ PTX)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); completion is
// signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void tma_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// raster
// ------------------------------------------------------------------------------------------------
struct StagedSmem {
    unsigned char *color;          // [C][plane_stride]
    unsigned long long *ktile;     // [nblk*64]
    Rec *recs[2];                  // 2 x [CH]
    SRec *srecs[2];                // 2 x [CH] (smooth frames only)
    unsigned *bbox[2];             // 2 x [CH]
    unsigned *masks;               // [nblk*MW]
    unsigned short *blist;         // [nblk]
    int *ctr;                      // nlist, next
    unsigned long long *bar;       // 2 mbarriers
};

__host__ __device__ inline size_t staged_smem_bytes(int C, int plane_stride, int nblk, bool smooth) {
    return (smooth ? 2 * (size_t)CH * sizeof(SRec) : 0) + align16((size_t)C * plane_stride) + (size_t)nblk * 64 * 8 + 2 * (size_t)CH * sizeof(Rec) + 2 * (size_t)CH * 4 +
           align16((size_t)nblk * MW * 4) + align16((size_t)nblk * 2) + 16 + 16;
}

__device__ __forceinline__ StagedSmem staged_carve(unsigned char *base, int C, int plane_stride, int nblk, bool smooth) {
    StagedSmem s;
    s.color = base; base += align16((size_t)C * plane_stride);
    s.ktile = reinterpret_cast<unsigned long long *>(base); base += (size_t)nblk * 64 * 8;
    s.recs[0] = reinterpret_cast<Rec *>(base); base += (size_t)CH * sizeof(Rec);
    s.recs[1] = reinterpret_cast<Rec *>(base); base += (size_t)CH * sizeof(Rec);
    s.srecs[0] = reinterpret_cast<SRec *>(base); base += smooth ? (size_t)CH * sizeof(SRec) : 0;
    s.srecs[1] = reinterpret_cast<SRec *>(base); base += smooth ? (size_t)CH * sizeof(SRec) : 0;
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
    const StagedSmem s = staged_carve(smem_raw, f.C, f.plane_stride, nblk, SMOOTH);

    const int total = min(g.count[local_scene], g.cap);
    const Rec *grecs = g.recs + (size_t)local_scene * g.cap;
    const unsigned *gbbox = g.bbox + (size_t)local_scene * g.cap;
    const int nchunks = (total + CH - 1) / CH;

    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {       // thread 0: stage chunk c into buffer c & 1
        const int cnt = min(CH, total - c * CH);
        const unsigned rb = (unsigned)cnt * (unsigned)sizeof(Rec), bb = (unsigned)align16((size_t)cnt * 4);
        const unsigned sb = SMOOTH ? (unsigned)cnt * (unsigned)sizeof(SRec) : 0u;
        mbar_expect_tx(&s.bar[c & 1], rb + sb + bb);
        tma_load(s.recs[c & 1], grecs + (size_t)c * CH, rb, &s.bar[c & 1]);
        if (SMOOTH) tma_load(s.srecs[c & 1], g.srecs + (size_t)local_scene * g.cap + (size_t)c * CH, sb, &s.bar[c & 1]);
        tma_load(s.bbox[c & 1], gbbox + (size_t)c * CH, bb, &s.bar[c & 1]);
    };
    if (tid == 0 && nchunks > 0) issue(0);

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
        if (tid == 0 && c + 1 < nchunks) issue(c + 1);
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
                const unsigned bit = 1u << (tid & 31);
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
            const int bxl = b % f.nbx, byl = b / f.nbx;
            const int px = bxl * 8 + (lane & 7);
            const int py0 = band_y0 + byl * 8 + (lane >> 3), py1 = py0 + 4;      // tile-global rows
            const bool ok0 = px < f.W && py0 < band_y0 + band_h, ok1 = px < f.W && py1 < band_y0 + band_h;
            PixelState ps;
            ps.k0 = s.ktile[b * 64 + lane];
            ps.k1 = s.ktile[b * 64 + 32 + lane];
            ps.c0 = ps.c1 = 0;
            ps.ch0 = ps.ch1 = false;
            raster_block<MW, SMOOTH>(recs, s.masks + b * MW, px, py0, ok0, ok1, ps, &f, s.srecs[buf]);
            if (ps.ch0) {
                s.ktile[b * 64 + lane] = ps.k0;
                put_pixel(s.color, f.plane_stride, f.C, f.W, px, py0 - band_y0, ps.c0);
            }
            if (ps.ch1) {
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
