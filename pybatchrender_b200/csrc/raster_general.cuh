// raster_general.cuh -- the general raster kernel: one CTA (128 threads) per (scene, band of rows).
//
// Any number of triangle slots (128 per pass), near-plane / guard-band clipping, depth + id tiles in
// shared memory so that passes compose.  Used for scenes the small-scene kernel (raster_warp.cuh)
// does not take: many instances, large tiles, smooth meshes (SMOOTH instantiation: per-pixel
// normal interpolation, basic.frag:33-38, from a second 64-byte record per triangle).
//
// Per pass: (1) one triangle slot per thread: transform (reference basic.vert:24-56), trivial
// reject, project, snap, cull, edge setup, flat shade (basic.frag:31-38) -> shared-memory record;
// (2) the same thread bins its record into per-8x8-block bitmasks; (3) warps pull non-empty blocks
// from a shared counter and sweep them (depth/id/colour of two pixels per lane in registers);
// clipped triangles are queued and run as extra passes of 16 slots x 8 fan triangles;
// (4) the finished band is written with 128-bit streaming stores into out[scene].
#pragma once
#include "common.cuh"

namespace pbr {

constexpr int THREADS = 128;
constexpr int CH = 128;              // triangle records per pass (one per thread)
constexpr int MW = CH / 32;          // mask words per 8x8 block

enum { SLOT_SKIP = 0, SLOT_OK = 1, SLOT_CLIP = 2 };

struct SlotGeom {
    CVT v[3];
    float4 col;
    const NodeDev *node;  // texture source (node->tex may be NULL)
    bool flat;
    bool two_sided;
    unsigned id;          // 1 + draw index
};

// triangle slot -> (node, instance of the node, triangle of the mesh)
__device__ __forceinline__ void locate_slot(const FrameDev &f, int slot, int &ni, int &inst, int &tri) {
    ni = 0;
#pragma unroll 1
    for (int i = 1; i < f.n_nodes; ++i)
        if (slot >= f.nodes[i].slot_begin) ni = i;
    const int local = slot - f.nodes[ni].slot_begin;
    // (an integer divide here was 8 % of bin_tri_kernel's instructions; the multiply is exact while x * n < 2^32)
    inst = (f.total_slots < 65536 && f.nodes[ni].n_tris < 65536) ? fast_div(local, f.nodes[ni].tri_magic)
                                                                 : local / f.nodes[ni].n_tris;
    tri = local - inst * f.nodes[ni].n_tris;
}

__device__ __forceinline__ int load_slot(const FrameDev &f, int scene, int slot, int ni, int inst, int tri, SlotGeom &g) {
    const NodeDev &nd = f.nodes[ni];
    const int local = slot - nd.slot_begin;
    const size_t b = nd.shared ? (size_t)inst : (size_t)scene * nd.inst + inst;

    float M[16], VP[16];
    const float4 *m4 = reinterpret_cast<const float4 *>(nd.mats + b * 16);
    const int vp_row = f.vp_scene_override >= 0 ? f.vp_scene_override : scene;
    const float4 *v4 = reinterpret_cast<const float4 *>(f.vp + (size_t)vp_row * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 a = __ldg(m4 + j), c = __ldg(v4 + j);
        M[4 * j] = a.x; M[4 * j + 1] = a.y; M[4 * j + 2] = a.z; M[4 * j + 3] = a.w;
        VP[4 * j] = c.x; VP[4 * j + 1] = c.y; VP[4 * j + 2] = c.z; VP[4 * j + 3] = c.w;
    }
    g.col = __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4));
    g.two_sided = (nd.flags & PBR_MESH_TWO_SIDED) != 0;
    g.node = &nd;
    const bool textured = nd.tex != nullptr;
    g.id = (unsigned)(nd.id_begin + local) + 1u;

    const float4 p0 = __ldg(nd.tp + 3 * tri);
    g.flat = __float_as_int(p0.w) != 0;
    float4 n0 = __ldg(nd.tn + 3 * tri);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float4 p = (k == 0) ? p0 : __ldg(nd.tp + 3 * tri + k);
        float world[4];
        mat_vec4(M, p.x, p.y, p.z, 1.0f, world);
        mat_vec4(VP, world[0], world[1], world[2], world[3], g.v[k].c);
        float4 n = (k == 0 || g.flat) ? n0 : __ldg(nd.tn + 3 * tri + k);
        xform_normal(M, n.x, n.y, n.z, g.v[k].n);
        float2 uv = make_float2(0.0f, 0.0f);
        if (textured && nd.tuv != nullptr) uv = __ldg(nd.tuv + 3 * tri + k);
        g.v[k].uv[0] = uv.x; g.v[k].uv[1] = uv.y;
    }
    if (trivially_outside(g.v[0].c, g.v[1].c, g.v[2].c)) return SLOT_SKIP;
    const bool clip = needs_clip(g.v[0].c) || needs_clip(g.v[1].c) || needs_clip(g.v[2].c);
    return clip ? SLOT_CLIP : SLOT_OK;
}

__device__ __forceinline__ int load_slot(const FrameDev &f, int scene, int slot, SlotGeom &g) {
    int ni, inst, tri;
    locate_slot(f, slot, ni, inst, tri);
    return load_slot(f, scene, slot, ni, inst, tri, g);
}

// what setup_tri hands to write_srec
struct TriVary {
    float rw[3];
    bool swapped;    // vertices 1 and 2 were exchanged (two-sided back face)
};

// flat triangles are shaded here; the others get M_SMOOTH (| M_TEX) and need write_srec afterwards
__device__ __forceinline__ bool setup_tri(const FrameDev &f, const CVT *vin, const SlotGeom &g, int band_y0,
                                          int band_h, Rec &r, BBox &bb, TriVary &tv) {
    const float4 col = g.col;
    const bool two_sided = g.two_sided;
    const unsigned id = g.id;
    const NodeDev &nd = *g.node;
    const bool flat = g.flat && nd.tex == nullptr;        // textured triangles shade per pixel
    int X[3], Y[3];
    float z[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (!project_vertex(f, vin[i].c, X[i], Y[i], z[i], tv.rw[i])) return false;
    if (!setup_snapped(f, X, Y, z, two_sided, id, band_y0, band_h, r, bb, &tv.swapped)) return false;
    if (flat) {
        r.col = shade(f, vin[0].n, col);
    } else {
        r.col = 0;
        r.meta |= M_SMOOTH | (nd.tex != nullptr ? M_TEX : 0u);
    }
    return true;
}

// The per-pixel shading inputs of a M_SMOOTH record, written as 16-byte stores to shared or global
// memory (dst is 16-byte aligned, f.srec_stride bytes long).  Static indexing only: vin stays in
// registers.
__device__ __forceinline__ void write_srec(const FrameDev &f, void *dst, const CVT *vin, const SlotGeom &g,
                                           const TriVary &tv) {
    const bool s = tv.swapped;
    const CVT &a = vin[0];
    float n1[3], n2[3], uv1[2], uv2[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        n1[k] = s ? vin[2].n[k] : vin[1].n[k];
        n2[k] = s ? vin[1].n[k] : vin[2].n[k];
    }
    const float rw1 = s ? tv.rw[2] : tv.rw[1], rw2 = s ? tv.rw[1] : tv.rw[2];
    float4 *q = reinterpret_cast<float4 *>(dst);
    q[0] = make_float4(a.n[0], a.n[1], a.n[2], n1[0]);
    q[1] = make_float4(n1[1], n1[2], n2[0], n2[1]);
    q[2] = make_float4(n2[2], tv.rw[0], rw1, rw2);
    q[3] = g.col;
    if (f.srec_stride == SREC_TEXTURED) {
        const NodeDev &nd = *g.node;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uv1[k] = s ? vin[2].uv[k] : vin[1].uv[k];
            uv2[k] = s ? vin[1].uv[k] : vin[2].uv[k];
        }
        q[4] = make_float4(a.uv[0], a.uv[1], uv1[0], uv1[1]);
        const unsigned long long tp = reinterpret_cast<unsigned long long>(nd.tex);
        q[5] = make_float4(uv2[0], uv2[1], __uint_as_float((unsigned)tp), __uint_as_float((unsigned)(tp >> 32)));
        q[6] = make_float4(__int_as_float(nd.tw), __int_as_float(nd.th), nd.use_tex, 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------------
struct Smem {
    unsigned char *color;      // [C][plane_stride]
    unsigned long long *ktile; // [nblk*64] block-major (depth bits << 32) | id
    Rec *recs;                 // [CH]
    unsigned char *srecs;      // [CH] x srec_stride bytes when the frame shades per pixel, else unused
    unsigned *masks;           // [nblk*MW]
    unsigned short *blist;     // [nblk]
    unsigned short *cliplist;  // [CH]
    int *ctr;                  // nlist, next, nclip
};

__host__ __device__ inline size_t general_smem_bytes(int C, int plane_stride, int nblk, int srec_bytes) {
    size_t n = (size_t)CH * srec_bytes;
    n += align16((size_t)C * plane_stride);
    n += 2 * (size_t)nblk * 64 * 4;
    n += (size_t)CH * sizeof(Rec);
    n += align16((size_t)nblk * MW * 4);
    n += align16((size_t)nblk * 2);
    n += align16((size_t)CH * 2);
    n += 16;
    return n;
}

__device__ __forceinline__ Smem carve(unsigned char *base, int C, int plane_stride, int nblk, int srec_bytes) {
    Smem s;
    s.color = base; base += align16((size_t)C * plane_stride);
    s.ktile = reinterpret_cast<unsigned long long *>(base); base += (size_t)nblk * 64 * 8;
    s.recs = reinterpret_cast<Rec *>(base); base += (size_t)CH * sizeof(Rec);
    s.srecs = base; base += (size_t)CH * srec_bytes;
    s.masks = reinterpret_cast<unsigned *>(base); base += align16((size_t)nblk * MW * 4);
    s.blist = reinterpret_cast<unsigned short *>(base); base += align16((size_t)nblk * 2);
    s.cliplist = reinterpret_cast<unsigned short *>(base); base += align16((size_t)CH * 2);
    s.ctr = reinterpret_cast<int *>(base);
    return s;
}

template <bool SMOOTH>
__device__ __forceinline__ void general_block(const FrameDev &f, const Smem &s, int b, int band_h, int lane) {
    const int by = fast_div(b, f.nbx_magic), bx = b - by * f.nbx;
    const int px = bx * 8 + (lane & 7);
    const int py0 = by * 8 + (lane >> 3), py1 = py0 + 4;
    const bool ok0 = px < f.W && py0 < band_h, ok1 = px < f.W && py1 < band_h;
    PixelState ps;
    ps.k0 = s.ktile[b * 64 + lane];
    ps.k1 = s.ktile[b * 64 + 32 + lane];
    ps.c0 = ps.c1 = 0;
    const unsigned id0 = (unsigned)ps.k0, id1 = (unsigned)ps.k1;
    raster_block<MW, SMOOTH>(s.recs, s.masks + b * MW, px, py0, ok0, ok1, ps, &f, s.srecs);
    if (key_changed(ps.k0, id0)) {
        s.ktile[b * 64 + lane] = ps.k0;
        put_pixel(s.color, f.plane_stride, f.C, f.W, px, py0, ps.c0);
    }
    if (key_changed(ps.k1, id1)) {
        s.ktile[b * 64 + 32 + lane] = ps.k1;
        put_pixel(s.color, f.plane_stride, f.C, f.W, px, py1, ps.c1);
    }
}

// Called by all threads after records + masks of this pass are complete (and synchronised).
template <bool SMOOTH>
__device__ __forceinline__ void raster_pass(const FrameDev &f, const Smem &s, int nblk, int band_h) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { s.ctr[0] = 0; s.ctr[1] = 0; }
    __syncthreads();
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
        general_block<SMOOTH>(f, s, s.blist[i], band_h, lane);
    }
    __syncthreads();
}

__device__ __forceinline__ void zero_masks(const Smem &s, int nblk) {
    for (int i = threadIdx.x; i < nblk * MW; i += THREADS) s.masks[i] = 0;
}

// write a finished colour tile / band: 128-bit streaming stores straight into the caller's tensor
__device__ __forceinline__ void store_band(const FrameDev &f, const unsigned char *color, int scene, int band_y0,
                                           int band_h, int tid, int nthreads) {
    const size_t scene_bytes = (size_t)f.C * f.H * f.W;
    unsigned char *dst_scene = f.out + (size_t)scene * scene_bytes;
    if (f.linear) {
        const int n16 = (int)(scene_bytes / 16);
        const uint4 *src = reinterpret_cast<const uint4 *>(color);
        uint4 *dst = reinterpret_cast<uint4 *>(dst_scene);
        for (int i = tid; i < n16; i += nthreads) __stcs(dst + i, src[i]);
        for (int i = n16 * 16 + tid; i < (int)scene_bytes; i += nthreads) dst_scene[i] = color[i];
    } else {
        const int nbytes = band_h * f.W;
        for (int c = 0; c < f.C; ++c) {
            unsigned char *dst = dst_scene + ((size_t)c * f.H + band_y0) * f.W;
            const unsigned char *src = color + (size_t)c * f.plane_stride;
            if ((reinterpret_cast<size_t>(dst) & 15) == 0) {
                const int n16 = nbytes / 16;
                for (int i = tid; i < n16; i += nthreads)
                    __stcs(reinterpret_cast<uint4 *>(dst) + i, reinterpret_cast<const uint4 *>(src)[i]);
                for (int i = n16 * 16 + tid; i < nbytes; i += nthreads) dst[i] = src[i];
            } else {
                for (int i = tid; i < nbytes; i += nthreads) dst[i] = src[i];
            }
        }
    }
}

__device__ __forceinline__ void clear_color(const FrameDev &f, unsigned char *color, int tid, int nthreads) {
    const int n16 = f.plane_stride / 16;
    for (int c = 0; c < f.C; ++c) {
        const unsigned v = ((f.bg >> (8 * c)) & 255u) * 0x01010101u;
        uint4 *p = reinterpret_cast<uint4 *>(color + (size_t)c * f.plane_stride);
        const uint4 v4 = make_uint4(v, v, v, v);
        for (int i = tid; i < n16; i += nthreads) p[i] = v4;
    }
}

template <bool SMOOTH>
__global__ void __launch_bounds__(THREADS) raster_general_kernel(const __grid_constant__ FrameDev f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int scene = f.scene_begin + (int)(blockIdx.x / f.nbands);
    const int band = (int)(blockIdx.x % f.nbands);
    const int band_y0 = band * f.BH;
    const int band_h = min(f.BH, f.H - band_y0);
    const int nblk = f.nbx * f.nby;
    const Smem s = carve(smem_raw, f.C, f.plane_stride, nblk, SMOOTH ? f.srec_stride : 0);

    clear_color(f, s.color, tid, THREADS);
    {
        uint4 *kt = reinterpret_cast<uint4 *>(s.ktile);
        const uint4 clr = make_uint4(0u, 0x3F800000u, 0u, 0x3F800000u);
        for (int i = tid; i < nblk * 32; i += THREADS) kt[i] = clr;
    }

    const int S = f.total_slots;
#pragma unroll 1
    for (int chunk = 0; chunk < S; chunk += CH) {
        zero_masks(s, nblk);
        if (tid == 0) s.ctr[2] = 0;
        __syncthreads();
        {
            Rec r;
            r.meta = 0;
            const int slot = chunk + tid;
            if (slot < S) {
                SlotGeom g;
                const int st = load_slot(f, scene, slot, g);
                if (st == SLOT_OK) {
                    BBox bb;
                    TriVary tv;
                    if (setup_tri(f, g.v, g, band_y0, band_h, r, bb, tv)) {
                        bin_record<MW>(r, bb, tid, f.nbx, s.masks);
                        if (SMOOTH && (r.meta & M_SMOOTH)) write_srec(f, s.srecs + (size_t)tid * f.srec_stride, g.v, g, tv);
                    } else {
                        r.meta = 0;
                    }
                } else if (st == SLOT_CLIP) {
                    s.cliplist[atomicAdd(&s.ctr[2], 1)] = (unsigned short)tid;
                }
            }
            s.recs[tid] = r;
        }
        __syncthreads();
        const int nclip = s.ctr[2];
        raster_pass<SMOOTH>(f, s, nblk, band_h);

        // clipped triangles: 16 slots x 8 fan triangles per pass
#pragma unroll 1
        for (int q0 = 0; q0 < nclip; q0 += CH / FAN) {
            zero_masks(s, nblk);
            __syncthreads();
            Rec r;
            r.meta = 0;
            const int q = q0 + tid / FAN, k = tid % FAN;
            if (q < nclip) {
                const int slot = chunk + s.cliplist[q];
                SlotGeom g;
                load_slot(f, scene, slot, g);
                CVT poly[MAX_POLY];
                const int n = clip_poly(g.v, poly);
                if (k + 2 < n) {
                    CVT tri[3] = {poly[0], poly[k + 1], poly[k + 2]};
                    BBox bb;
                    TriVary tv;
                    if (setup_tri(f, tri, g, band_y0, band_h, r, bb, tv)) {
                        bin_record<MW>(r, bb, tid, f.nbx, s.masks);
                        if (SMOOTH && (r.meta & M_SMOOTH)) write_srec(f, s.srecs + (size_t)tid * f.srec_stride, tri, g, tv);
                    } else {
                        r.meta = 0;
                    }
                }
            }
            s.recs[tid] = r;
            __syncthreads();
            raster_pass<SMOOTH>(f, s, nblk, band_h);
        }
    }
    __syncthreads();
    store_band(f, s.color, scene, band_y0, band_h, tid, THREADS);

    // static-layer pass: also publish the depth|id keys and which blocks have any coverage
    if (f.base_keys_out != nullptr) {
        const size_t blk0 = (size_t)band * (f.BH / 8) * f.nbx;
        for (int i = tid; i < nblk * 64; i += THREADS) f.base_keys_out[blk0 * 64 + i] = s.ktile[i];
        for (int b = tid; b < nblk; b += THREADS) {
            bool any = false;
            for (int k = 0; k < 64; ++k) any |= s.ktile[b * 64 + k] != KEY_CLEAR;
            f.base_flags_out[blk0 + b] = any ? 1 : 0;
        }
    }
}

}  // namespace pbr
