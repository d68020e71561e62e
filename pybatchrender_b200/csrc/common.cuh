// common.cuh -- frame description, triangle records and the fp32 / integer math shared by the
// raster kernels.  Every float operation here mirrors oracle/pbr_oracle.c one to one (same order,
// explicit fmaf only; the library is compiled with -fmad=false so nothing else fuses).
#pragma once
#include "../../include/pbr_b200.h"

#include <cuda_runtime.h>

namespace pbr {

constexpr int MAX_POLY = 10;         // vertices of a clipped polygon (3 + 5 planes, with slack)
constexpr int FAN = 8;               // max fan triangles of a clipped polygon

// ------------------------------------------------------------------------------------------------
// device-side frame description (kernel parameter, < 4 KB)
// ------------------------------------------------------------------------------------------------
// pose channels of a node whose matrices are computed in the kernel (pbr_node_desc.pose)
struct PoseDev {
    pbr_channel pos[3], hpr[3], scale;
    float *out_mats;      // [B,16]: written only when the frame asks for it (PBR_FRAME_WRITE_MATS)
};
constexpr int MAX_FRAME_PF = 4;      // prefetch ranges per frame (32 lanes = 4 ranges x 8 lines)
constexpr int MAX_FRAME_POSES = 4;   // posed nodes per frame handled in-kernel; more are materialised by compose_kernel

struct NodeDev {
    const float4 *tp;     // [T*3] triangle soup: xyz = position, w of corner 0 = flat flag (int bits)
    const float4 *tn;     // [T*3] xyz = corner normal
    const float4 *vpos;   // [V]   unique positions (de-duplicated)
    const uint4 *tidx;    // [T]   (i0, i1, i2, flat) into vpos
    const float2 *tuv;    // [T*3] texture coordinates of the triangle corners, or NULL
    const uchar4 *tex;    // [th, tw] RGBA8 texels, row 0 = v 0 (bottom), or NULL = untextured
    int tw, th;
    float use_tex;        // clamp(useTexture): mix(1, texel, use_tex)
    float4 bsphere;       // object-space bounding sphere of the mesh (centre xyz, radius w)
    int inst_begin;       // index of this node's first instance among the scene's instances
    const float *mats;    // [B,16] column packed
    const float *cols;    // [B,4]
    int n_tris;
    int n_verts;          // unique positions
    int inst;             // instances per scene
    int shared;
    int slot_begin;       // first triangle slot of this node inside a scene (compact, active nodes only)
    int vert_begin;       // first (instance, vertex) pair of this node inside a scene
    unsigned flags;
    int id_begin;         // draw index of this node's first triangle in the full frame (skipped nodes count)
    unsigned tri_magic;   // floor(2^32 / n_tris) + 1   (x / n_tris == __umulhi(x, magic) for x * n_tris < 2^32)
    unsigned vert_magic;  // floor(2^32 / n_verts) + 1
    int pose_idx;         // >= 0: matrices come from FrameDev::poses[pose_idx] instead of mats
};

struct Rec;

struct FrameDev {
    const float *vp;
    unsigned char *out;
    int *status;            // device word: sticky PBR_DEVSTAT_* bits
    volatile int *status_host;   // the same bits in host-mapped pinned memory (read by the host without a sync)
    // static layer (see pbr_base_t): inputs of a frame that starts from it ...
    const unsigned char *base_color;         // [C,H,W]
    const unsigned long long *base_keys;     // [nblk,64] block-major depth|id keys
    const unsigned char *base_flags;         // [nblk] 1 = block has base coverage
    // ... and outputs of the pass that renders it (general kernel, one scene)
    unsigned long long *base_keys_out;
    unsigned char *base_flags_out;
    // small-scene kernel: where records beyond its shared-memory slots go (see raster_warp.cuh)
    Rec *ovf_recs;               // [n_sm * 32][W_OVF_MAXREC]
    unsigned *ovf_masks;         // [n_sm * 32][nblk * W_OVF_MW]
    unsigned *ovf_busy;          // [n_sm] one bit per pool entry of that SM
    unsigned *bg_ticket, *bg_done;   // [n_sm] each: order of the CTAs' bulk-store phases on an SM (PBR_W_BG_SERIAL)
    int vp_scene_override;  // >= 0: use this row of vp for every scene (base pass)
    int scene_begin, scene_count;
    int W, H, C;
    int n_nodes, total_slots, total_verts, total_inst;
    int BH, nbands;         // band height (multiple of 8) and bands per tile
    int nbx, nby;           // 8x8 blocks per band
    unsigned nbx_magic;     // floor(2^32 / nbx) + 1
    int w_region;           // small-scene kernel: bytes of one scene's shared-memory region
    int w_qctr_off;         // ... and offset of the CTA's queue counters (after the regions, the queue and the live list)
    unsigned w_inst_magic, w_vert_magic, w_slot_magic;   // div_magic of total_inst / total_verts / total_slots
    int plane_stride;       // bytes between colour planes in shared memory (multiple of 16)
    int linear;             // 1: the shared colour tile is a byte image of out[scene]
    int smooth;             // 1: some triangles are shaded per pixel (SMOOTH kernel instantiations)
    int srec_stride;        // bytes per SRec: SREC_PLAIN, or SREC_TEXTURED when a node is textured
    int debug;              // profiling aid: 1 = stop after the background, 2 = stop after setup
    float hw, hh;
    unsigned bg;            // packed RGBA8 clear colour
    float amb[3], dcol[3], ldir[3];
    float s, oms;           // clamp(strength), 1 - clamp(strength)
    int write_mats;         // small-scene kernel: posed nodes also write their matrices to out_mats
    int keys32;             // small-scene kernel: the static layer (if any) was drawn before every node of this frame,
                            // so scenes without clipped / int64 records may use 32-bit depth keys (raster_block32)
    int sync_early;         // small-scene kernel: wait for the previous grid before the first write to `out`
                            // (the previous launch on this stream may still be writing the same buffer)
    // small-scene kernel: per-scene inputs a CTA asks the L2 for on behalf of a later CTA (see PBR_W_PF_DIST):
    // base address of scene 0's row and bytes per scene
    int n_pf;
    const unsigned char *pf_ptr[MAX_FRAME_PF];
    int pf_row[MAX_FRAME_PF];
    PoseDev poses[MAX_FRAME_POSES];
    NodeDev nodes[PBR_MAX_NODES];
};
static_assert(sizeof(FrameDev) <= 8192, "FrameDev is a kernel parameter");

// ------------------------------------------------------------------------------------------------
// pose -> column-packed model matrix (reference shader_context.py:47-84: R = Rz(h) Ry(p) Rx(r);
// node.py:116-126: M = [R*s | t]).  One function for compose_kernel and for the small-scene raster
// kernel, so that a pose folded into the frame gives the same bits as the materialised matrices.
// An angle channel that is the constant 0 skips sincosf (sin = 0, cos = 1 exactly, as sincosf returns).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pose_chan(const pbr_channel &c, size_t b) {
    return c.ptr ? __ldg(c.ptr + b * (size_t)c.stride) : c.constant;
}
// one out-of-line copy of sincosf per kernel: inlined three times (with its large-argument path) it was 600 SASS
// instructions of straight-line code that a CTA runs once -- the geometry phases pay for instruction fetch, not issue
__device__ __noinline__ float2 pose_sincos(float x) {
    float sn, cs;
    sincosf(x, &sn, &cs);
    return make_float2(sn, cs);
}
__device__ __forceinline__ void pose_matrix(const PoseDev &d, size_t b, float *M) {
    // every channel value first -- seven independent loads in flight together -- then the trigonometry (the calls are
    // out of line: a load behind one of them would wait for its own round trip)
    const float ah = pose_chan(d.hpr[0], b), ap = pose_chan(d.hpr[1], b), ar = pose_chan(d.hpr[2], b);
    const float s = pose_chan(d.scale, b);
    const float tx = pose_chan(d.pos[0], b), ty = pose_chan(d.pos[1], b), tz = pose_chan(d.pos[2], b);
    float sh = 0.0f, ch = 1.0f, sp = 0.0f, cp = 1.0f, sr = 0.0f, cr = 1.0f;
    if (d.hpr[0].ptr != nullptr || d.hpr[0].constant != 0.0f) { const float2 v = pose_sincos(ah); sh = v.x; ch = v.y; }
    if (d.hpr[1].ptr != nullptr || d.hpr[1].constant != 0.0f) { const float2 v = pose_sincos(ap); sp = v.x; cp = v.y; }
    if (d.hpr[2].ptr != nullptr || d.hpr[2].constant != 0.0f) { const float2 v = pose_sincos(ar); sr = v.x; cr = v.y; }
    const float r00 = ch * cp, r01 = ch * sp * sr - sh * cr, r02 = ch * sp * cr + sh * sr;
    const float r10 = sh * cp, r11 = sh * sp * sr + ch * cr, r12 = sh * sp * cr - ch * sr;
    const float r20 = -sp, r21 = cp * sr, r22 = cp * cr;
    M[0] = r00 * s; M[1] = r10 * s; M[2] = r20 * s; M[3] = 0.0f;
    M[4] = r01 * s; M[5] = r11 * s; M[6] = r21 * s; M[7] = 0.0f;
    M[8] = r02 * s; M[9] = r12 * s; M[10] = r22 * s; M[11] = 0.0f;
    M[12] = tx; M[13] = ty; M[14] = tz; M[15] = 1.0f;
}

constexpr int DEVSTAT_WARP_OVERFLOW = 1;   // small-scene kernel ran out of record slots
constexpr int DEVSTAT_STAGED_OVERFLOW = 2; // geometry pre-pass ran out of per-scene record capacity

// record meta bits
// byte 0: flags; bytes 1..3: 1 if edge 0 / 1 / 2 is not a top-left edge (its stored edge value carries a bias of
// -1) -- a byte each so that the sweep gets them with one PRMT / shift instead of shift + mask
constexpr unsigned M_VALID = 1u, M_SLOW = 2u, M_SMOOTH = 4u;
constexpr unsigned M_TEX = 8u;              // the record's SRec carries a texture part
constexpr unsigned M_NB0 = 1u << 8, M_NB1 = 1u << 16, M_NB2 = 1u << 24;
__device__ __forceinline__ int meta_nb0(unsigned meta) { return (int)__byte_perm(meta, 0u, 0x4441u); }
__device__ __forceinline__ int meta_nb1(unsigned meta) { return (int)__byte_perm(meta, 0u, 0x4442u); }
__device__ __forceinline__ int meta_nb2(unsigned meta) { return (int)(meta >> 24); }

struct __align__(16) Rec {
    int e[9];        // fast: Eo[3], A[3], B[3]        slow: X0,Y0,X1,Y1,X2,Y2,-,-,-
    unsigned col;    // packed RGBA8 (flat shading)
    unsigned id;     // 1 + draw index
    unsigned meta;   // flags | edge bias bytes (M_*)
    float z0, dz1, dz2, invA;   // 16-byte aligned: one LDS.128
};
static_assert(sizeof(Rec) == 64, "Rec must be 64 bytes");

// companion of a Rec for triangles with per-vertex normals (smooth shading): what the fragment
// shader interpolates (reference basic.vert:53-54 -> basic.frag:33-38)
// Stored with a per-frame stride: 64 bytes (first four fields) when no node is textured, 128 bytes
// otherwise; the texture part is only read for records flagged M_TEX.
struct __align__(16) SRec {
    float n[3][3];   // world-space unit normals of the three vertices (after a two-sided swap)
    float rw[3];     // 1 / w_clip of the three vertices (perspective-correct weights)
    float col[4];    // instance RGBA
    // ---- texture part
    float uv[3][2];  // texture coordinates of the three vertices
    const uchar4 *tex;
    int tw, th;
    float use_tex;
    float pad[5];
};
static_assert(sizeof(SRec) == 128 && offsetof(SRec, uv) == 64, "SRec layout");
constexpr int SREC_PLAIN = 64, SREC_TEXTURED = 128;

struct CV {
    float c[4];      // clip-space position
    float n[3];      // world-space unit normal
};
// vertex of the kernels that can texture: CV + texture coordinates
struct CVT : CV {
    float uv[2];
};
__device__ __forceinline__ void lerp_extra(CV &, const CV &, const CV &, float) {}
__device__ __forceinline__ void lerp_extra(CVT &w, const CVT &vi, const CVT &vo, float t) {
    w.uv[0] = fmaf(t, vo.uv[0] - vi.uv[0], vi.uv[0]);
    w.uv[1] = fmaf(t, vo.uv[1] - vi.uv[1], vi.uv[1]);
}

struct BBox {
    int bx0, by0, bx1, by1;   // 8x8-pixel blocks, inclusive
    int px0, py0, px1, py1;   // pixels whose centres lie inside the triangle's bounding box, inclusive
};
// pixel box of a small triangle in one word: x0 (11 bits) | y0 (11) | width - 1 (5) | height - 1 (5); boxes of more
// than PBOX_AREA pixels or longer than PBOX_SIDE (and anything the lane-per-record raster path must not take) are
// PBOX_NONE (the lane's hit mask has one bit per pixel of the box: at most 32).
constexpr unsigned PBOX_NONE = 0xffffffffu;
#ifndef PBR_PBOX_AREA
#define PBR_PBOX_AREA 16
#endif
constexpr int PBOX_AREA = PBR_PBOX_AREA, PBOX_SIDE = 16;
__device__ __forceinline__ unsigned pack_pbox(const BBox &bb) {
    const int w = bb.px1 - bb.px0 + 1, h = bb.py1 - bb.py0 + 1;
    if (w > PBOX_SIDE || h > PBOX_SIDE || w * h > PBOX_AREA) return PBOX_NONE;
    return (unsigned)bb.px0 | ((unsigned)bb.py0 << 11) | ((unsigned)(w - 1) << 22) | ((unsigned)(h - 1) << 27);
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// exact x / d for x * d < 2^32 (d >= 2), via a precomputed magic = floor(2^32 / d) + 1; d == 1 -> magic 0
__host__ inline unsigned div_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)(0x100000000ull / d) + 1u; }
__device__ __forceinline__ int fast_div(int x, unsigned magic) { return magic ? (int)__umulhi((unsigned)x, magic) : x; }

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
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store(void *dst_gmem, const void *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread complete: their global writes are performed
__device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// math shared with the oracle (same operation order)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat_vec4(const float *m, float x, float y, float z, float w, float *r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fmaf(m[i], x, fmaf(m[4 + i], y, fmaf(m[8 + i], z, m[12 + i] * w)));
}

__device__ __forceinline__ void xform_normal(const float *m, float nx, float ny, float nz, float *r) {
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = fmaf(m[i], nx, fmaf(m[4 + i], ny, m[8 + i] * nz));
    float l2 = fmaf(t[2], t[2], fmaf(t[1], t[1], t[0] * t[0]));
    float inv = 1.0f / sqrtf(l2);
    r[0] = t[0] * inv; r[1] = t[1] * inv; r[2] = t[2] * inv;
}

__device__ __forceinline__ unsigned unorm8(float c) {
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (unsigned)__float2int_rz(c * 255.0f + 0.5f);
}

// ambient + Lambert with the strength blend (reference basic.frag:33-37), n unit length
__device__ __forceinline__ unsigned shade(const FrameDev &f, const float *n, const float4 col) {
    float ndl = fmaxf(fmaf(n[2], f.ldir[2], fmaf(n[1], f.ldir[1], n[0] * f.ldir[0])), 0.0f);
    float cc[3] = {col.x, col.y, col.z};
    unsigned out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float light = fmaf(ndl, f.dcol[c], f.amb[c]);
        float l = fmaf(light, f.s, f.oms);
        out |= unorm8(cc[c] * l) << (8 * c);
    }
    out |= unorm8(col.w) << 24;
    return out;
}

__device__ __forceinline__ float plane_dist4(const float *c, int p) {
    const float G = 1024.0f;
    switch (p) {
    case 0: return c[2] + c[3];
    case 1: return G * c[3] + c[0];
    case 2: return G * c[3] - c[0];
    case 3: return G * c[3] + c[1];
    default: return G * c[3] - c[1];
    }
}
__device__ __forceinline__ float plane_dist(const CV &v, int p) { return plane_dist4(v.c, p); }

// all three vertices outside one plane of the tile frustum?
__device__ __forceinline__ bool trivially_outside(const float *c0, const float *c1, const float *c2) {
    bool rej = false;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const int a = p >> 1;
        const bool o0 = (p & 1) ? (c0[a] > c0[3]) : (c0[a] < -c0[3]);
        const bool o1 = (p & 1) ? (c1[a] > c1[3]) : (c1[a] < -c1[3]);
        const bool o2 = (p & 1) ? (c2[a] > c2[3]) : (c2[a] < -c2[3]);
        rej |= (o0 && o1 && o2);
    }
    return rej;
}

__device__ __forceinline__ bool needs_clip(const float *c) {
    bool need = false;
#pragma unroll
    for (int p = 0; p < 5; ++p) need |= plane_dist4(c, p) < 0.0f;
    return need;
}

// Sutherland-Hodgman against near + guard band; returns vertex count (0 = nothing left).
template <class V>
__device__ __noinline__ int clip_poly(const V *in3, V *a) {
    V b[MAX_POLY];
    int n = 3;
    for (int i = 0; i < 3; ++i) a[i] = in3[i];
    for (int p = 0; p < 5 && n >= 3; ++p) {
        bool any_out = false;
        for (int i = 0; i < n; ++i) any_out |= plane_dist(a[i], p) < 0.0f;
        if (!any_out) continue;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const V &u = a[i];
            const V &v = a[(i + 1) % n];
            float du = plane_dist(u, p), dv = plane_dist(v, p);
            bool iu = !(du < 0.0f), iv = !(dv < 0.0f);
            if (iu) b[m++] = u;
            if (iu != iv) {
                const V &vi = iu ? u : v;
                const V &vo = iu ? v : u;
                float di = iu ? du : dv, dout = iu ? dv : du;
                float t = di / (di - dout);
                V w;
                for (int k = 0; k < 4; ++k) w.c[k] = fmaf(t, vo.c[k] - vi.c[k], vi.c[k]);
                for (int k = 0; k < 3; ++k) w.n[k] = fmaf(t, vo.n[k] - vi.n[k], vi.n[k]);
                lerp_extra(w, vi, vo, t);
                b[m++] = w;
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) a[i] = b[i];
    }
    return n < 3 ? 0 : n;
}

// perspective divide + viewport (image orientation, y down) + snap to 1/256 px
__device__ __forceinline__ bool project_vertex(const FrameDev &f, const float *c, int &X, int &Y, float &z, float &rw) {
    if (!(c[3] > 0.0f)) return false;
    rw = 1.0f / c[3];
    const float xs = fmaf(c[0] * rw, f.hw, f.hw);
    const float ys = fmaf(-(c[1] * rw), f.hh, f.hh);
    z = fmaf(0.5f, c[2] * rw, 0.5f);
    const float fx = xs * 256.0f, fy = ys * 256.0f;
    if (!(fabsf(fx) < 1073741824.0f) || !(fabsf(fy) < 1073741824.0f)) return false;
    X = __float2int_rn(fx);
    Y = __float2int_rn(fy);
    return true;
}
__device__ __forceinline__ bool project_vertex(const FrameDev &f, const float *c, int &X, int &Y, float &z) {
    float rw;
    return project_vertex(f, c, X, Y, z, rw);
}

// ------------------------------------------------------------------------------------------------
// triangle setup from snapped vertices: cull, edge equations, depth plane -> record + block bbox.
// r.col is left for the caller (flat colour is only worth computing for surviving triangles).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool setup_snapped(const FrameDev &f, int *X, int *Y, float *z, bool two_sided,
                                              unsigned id, int band_y0, int band_h, Rec &r, BBox &bb,
                                              bool *swapped = nullptr) {
    if (swapped) *swapped = false;
    long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 == 0) return false;
    if (area2 > 0) {                 // clockwise in GL's y-up window space: back face
        if (!two_sided) return false;
        int t = X[1]; X[1] = X[2]; X[2] = t;
        t = Y[1]; Y[1] = Y[2]; Y[2] = t;
        float q = z[1]; z[1] = z[2]; z[2] = q;
        area2 = -area2;
        if (swapped) *swapped = true;
    }
    const long long A2 = -area2;

    // band-local coordinates (edge functions are translation invariant)
    const int yshift = band_y0 * 256;
#pragma unroll
    for (int i = 0; i < 3; ++i) Y[i] -= yshift;

    const int xmin = min(X[0], min(X[1], X[2])), xmax = max(X[0], max(X[1], X[2]));
    const int ymin = min(Y[0], min(Y[1], Y[2])), ymax = max(Y[0], max(Y[1], Y[2]));
    const int i0 = max(0, (xmin - 128 + 255) >> 8), i1 = min(f.W - 1, (xmax - 128) >> 8);
    const int j0 = max(0, (ymin - 128 + 255) >> 8), j1 = min(band_h - 1, (ymax - 128) >> 8);
    if (i0 > i1 || j0 > j1) return false;
    bb.bx0 = i0 >> 3; bb.bx1 = i1 >> 3; bb.by0 = j0 >> 3; bb.by1 = j1 >> 3;
    bb.px0 = i0; bb.px1 = i1; bb.py0 = j0; bb.py1 = j1;

    r.invA = 1.0f / (float)A2;
    r.z0 = z[0];
    r.dz1 = z[1] - z[0];
    r.dz2 = z[2] - z[0];
    r.id = id;
    r.col = 0;

    unsigned meta = M_VALID;
    int dx[3], dy[3], bias[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        dx[i] = X[b] - X[a];
        dy[i] = Y[b] - Y[a];
        // top-left rule in image space: samples exactly on an edge belong to left / top edges only
        bias[i] = (dy[i] > 0 || (dy[i] == 0 && dx[i] < 0)) ? 0 : -1;
    }
    if (bias[0]) meta |= M_NB0;
    if (bias[1]) meta |= M_NB1;
    if (bias[2]) meta |= M_NB2;

    // int32 fast path iff every |F| over hull(triangle, touched blocks) stays below 2^30
    const long long rx0 = (long long)bb.bx0 * 2048 + 128, rx1 = (long long)bb.bx1 * 2048 + 7 * 256 + 128;
    const long long ry0 = (long long)bb.by0 * 2048 + 128, ry1 = (long long)bb.by1 * 2048 + 7 * 256 + 128;
    const long long spanx = max((long long)xmax, rx1) - min((long long)xmin, rx0);
    const long long spany = max((long long)ymax, ry1) - min((long long)ymin, ry0);
    const bool fast = spanx < (1ll << 30) && spany < (1ll << 30) && spanx * spany < (1ll << 29);
    if (fast) {
        // biased edge value at the centre of pixel (0, 0), modulo 2^32: the raster loop adds A*px + B*py in
        // unsigned arithmetic, and the true value at every sample it visits fits an int (|F| < 2^30 over
        // the hull), so intermediate wrap-around is harmless and no per-record origin has to be unpacked
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int a = (i + 1) % 3;
            r.e[i] = (int)((unsigned)dy[i] * (unsigned)(128 - X[a]) - (unsigned)dx[i] * (unsigned)(128 - Y[a]) + (unsigned)bias[i]);
            r.e[3 + i] = dy[i] * 256;
            r.e[6 + i] = -dx[i] * 256;
        }
    } else {
        meta |= M_SLOW;
        r.e[0] = X[0]; r.e[1] = Y[0]; r.e[2] = X[1]; r.e[3] = Y[1]; r.e[4] = X[2]; r.e[5] = Y[2];
        r.e[6] = r.e[7] = r.e[8] = 0;
    }
    r.meta = meta;
    return true;
}

// exact (biased) edge value of a slow-path record at sample (px,py) in 1/256 px units
__device__ __forceinline__ long long slow_edge(const Rec &r, int i, int px, int py) {
    const int a = (i + 1) % 3, b = (i + 2) % 3;
    const int xa = r.e[2 * a], ya = r.e[2 * a + 1], xb = r.e[2 * b], yb = r.e[2 * b + 1];
    const long long F = (long long)(yb - ya) * (px - xa) - (long long)(xb - xa) * (py - ya);
    const unsigned nb = (r.meta >> (8 + 8 * i)) & 1u;
    return F - (long long)nb;
}

// can the triangle of this record touch block (bx,by)?  (conservative: max of each edge function
// over the block's pixel centres must be non-negative)
__device__ __forceinline__ bool block_hit(const Rec &r, const BBox &bb, int bx, int by) {
    bool hit = true;
    if (!(r.meta & M_SLOW)) {
        const int ox = bx * 8, oy = by * 8;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int A = r.e[3 + i], B = r.e[6 + i];
            const unsigned cx = (unsigned)(ox + (A > 0 ? 7 : 0)), cy = (unsigned)(oy + (B > 0 ? 7 : 0));
            const int fmax = (int)((unsigned)r.e[i] + (unsigned)A * cx + (unsigned)B * cy);
            hit &= fmax >= 0;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int a = (i + 1) % 3, b = (i + 2) % 3;
            const int ddx = r.e[2 * b] - r.e[2 * a], ddy = r.e[2 * b + 1] - r.e[2 * a + 1];
            const int cx = (bx * 8 + (ddy > 0 ? 7 : 0)) * 256 + 128;      // dF/dpx = ddy
            const int cy = (by * 8 + (ddx < 0 ? 7 : 0)) * 256 + 128;      // dF/dpy = -ddx
            hit &= slow_edge(r, i, cx, cy) >= 0;
        }
    }
    return hit;
}

// set record t's bit in every block its triangle can touch.  masks: [nblk][MWORDS]
// Record t of a mask word is bit 31 - (t & 31): the record that is drawn first is the HIGHEST set bit, which one FLO
// finds (ffs is BREV + FLO, both on the quarter-rate XU pipe), and the bit index times the record size is subtracted
// from the address of the word's record 31.
__device__ __forceinline__ unsigned rec_bit(int t) { return 0x80000000u >> (t & 31); }
// pops the first record of a mask word: returns 31 - (its index in the word)
__device__ __forceinline__ int pop_rec(unsigned &m) {
    int p;
    asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(m));      // (31 - __clz(m) is rewritten into clz arithmetic again)
    unsigned below;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(below) : "r"(0), "r"(p));      // bits 0 .. p-1
    m &= below;
    return p;
}

template <int MWORDS>
__device__ __forceinline__ void bin_record(const Rec &r, const BBox &bb, int t, int nbx, unsigned *masks) {
    const unsigned bit = rec_bit(t);
    const int word = t >> 5;
    const bool small = (bb.bx1 - bb.bx0) + (bb.by1 - bb.by0) <= 1;       // <= 2 blocks: no reject test
    if (!small && !(r.meta & M_SLOW)) {
        // block_hit() for every block of the box, stepped: the largest value of each edge function over a block's
        // pixel centres is taken at a fixed corner, so it moves by 8 A per block to the right and 8 B per block down
        // (the same wrapping 32-bit arithmetic as block_hit: identical bits)
        unsigned row[3], sx[3], sy[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int A = r.e[3 + i], B = r.e[6 + i];
            row[i] = (unsigned)r.e[i] + (unsigned)A * (unsigned)(bb.bx0 * 8 + (A > 0 ? 7 : 0)) +
                     (unsigned)B * (unsigned)(bb.by0 * 8 + (B > 0 ? 7 : 0));
            sx[i] = 8u * (unsigned)A;
            sy[i] = 8u * (unsigned)B;
        }
        unsigned *mrow = masks + (bb.by0 * nbx + bb.bx0) * MWORDS + word;
#pragma unroll 1
        for (int by = bb.by0; by <= bb.by1; ++by) {
            unsigned f0 = row[0], f1 = row[1], f2 = row[2];
            unsigned *m = mrow;
#pragma unroll 1
            for (int bx = bb.bx0; bx <= bb.bx1; ++bx) {
                if ((int)(f0 | f1 | f2) >= 0) atomicOr(m, bit);
                f0 += sx[0]; f1 += sx[1]; f2 += sx[2];
                m += MWORDS;
            }
            row[0] += sy[0]; row[1] += sy[1]; row[2] += sy[2];
            mrow += nbx * MWORDS;
        }
        return;
    }
    for (int by = bb.by0; by <= bb.by1; ++by)
        for (int bx = bb.bx0; bx <= bb.bx1; ++bx)
            if (small || block_hit(r, bb, bx, by)) atomicOr(&masks[(by * nbx + bx) * MWORDS + word], bit);
}

// ------------------------------------------------------------------------------------------------
// one 8x8 block: every lane owns pixels (lx, ly) and (lx, ly+4); loop over the block's records
// ------------------------------------------------------------------------------------------------
struct PixelState {
    unsigned long long k0, k1;   // (depth bits << 32) | id   -- smaller wins (LESS, ties to the earlier draw)
    unsigned c0, c1;             // packed RGBA8 of the current winner (valid once the key's id changed)
};

// Did a record of this sweep win the pixel?  Every triangle slot has its own id and the fan
// triangles of one slot never cover the same sample, so "the winner changed" == "the id changed".
__device__ __forceinline__ bool key_changed(unsigned long long key, unsigned id_before) {
    return (unsigned)key != id_before;
}

constexpr unsigned long long KEY_CLEAR = 0x3F80000000000000ull;   // depth 1.0, id 0

__device__ __forceinline__ unsigned long long make_key(float z, unsigned id) {
    return ((unsigned long long)__float_as_uint(z) << 32) | id;
}

// exact int64 coverage of one slow-path record at this lane's two pixels (rare: huge triangles)
__device__ __noinline__ bool slow_cover(const Rec &r, int px, int py0, bool ok0, bool ok1, bool &cov0, bool &cov1,
                                        float &f1a, float &f2a, float &f1b, float &f2b, float &f0a, float &f0b) {
    const unsigned meta = r.meta;
    const int spx = px * 256 + 128, spy0 = py0 * 256 + 128, spy1 = (py0 + 4) * 256 + 128;
    const long long F0 = slow_edge(r, 0, spx, spy0), F1 = slow_edge(r, 1, spx, spy0), F2 = slow_edge(r, 2, spx, spy0);
    const long long G0 = slow_edge(r, 0, spx, spy1), G1 = slow_edge(r, 1, spx, spy1), G2 = slow_edge(r, 2, spx, spy1);
    cov0 = ok0 && ((F0 | F1 | F2) >= 0);
    cov1 = ok1 && ((G0 | G1 | G2) >= 0);
    const long long nb0 = meta_nb0(meta), nb1 = meta_nb1(meta), nb2 = meta_nb2(meta);
    f1a = (float)(F1 + nb1); f2a = (float)(F2 + nb2);
    f1b = (float)(G1 + nb1); f2b = (float)(G2 + nb2);
    f0a = (float)(F0 + nb0); f0b = (float)(G0 + nb0);
    return cov0 || cov1;
}

// int32 fast-path coverage of one record at this lane's two pixels; F* (G*) are the unbiased edge
// values of pixel 0 (pixel 1): F1, F2 weight the depth, F0 is only needed for smooth shading
struct FastCov {
    int F0, F1, F2, G0, G1, G2;
    bool cov0, cov1;
};

__device__ __forceinline__ FastCov fast_cover(const int4 &ea, const int4 &eb, const int4 &ec, int px, int py0,
                                              bool ok0, bool ok1) {
    const unsigned meta = (unsigned)ec.w;
    const unsigned rx = (unsigned)px, ry = (unsigned)py0;       // record edge values are relative to pixel (0, 0)
    const int B2 = ec.x;
    const int F0 = (int)((unsigned)ea.x + (unsigned)ea.w * rx + (unsigned)eb.z * ry);
    const int F1 = (int)((unsigned)ea.y + (unsigned)eb.x * rx + (unsigned)eb.w * ry);
    const int F2 = (int)((unsigned)ea.z + (unsigned)eb.y * rx + (unsigned)B2 * ry);
    const int G0 = (int)((unsigned)F0 + 4u * (unsigned)eb.z);
    const int G1 = (int)((unsigned)F1 + 4u * (unsigned)eb.w);
    const int G2 = (int)((unsigned)F2 + 4u * (unsigned)B2);
    FastCov c;
    c.cov0 = ok0 && ((F0 | F1 | F2) >= 0);
    c.cov1 = ok1 && ((G0 | G1 | G2) >= 0);
    const int nb0 = meta_nb0(meta), nb1 = meta_nb1(meta), nb2 = meta_nb2(meta);
    c.F0 = F0 + nb0; c.F1 = F1 + nb1; c.F2 = F2 + nb2;
    c.G0 = G0 + nb0; c.G1 = G1 + nb1; c.G2 = G2 + nb2;
    return c;
}

// depth from the edge values + LESS test (ties to the earlier draw) for both pixels of the lane
__device__ __forceinline__ void depth_update(PixelState &ps, float f1a, float f2a, float f1b, float f2b, bool cov0,
                                             bool cov1, const float4 &zq, unsigned id, unsigned col, bool &w0,
                                             bool &w1) {
    const float za = fmaf(f2a * zq.w, zq.z, fmaf(f1a * zq.w, zq.y, zq.x));
    const float zc = fmaf(f2b * zq.w, zq.z, fmaf(f1b * zq.w, zq.y, zq.x));
    const unsigned long long ka = make_key(za, id), kc = make_key(zc, id);
    w0 = cov0 & (ka < ps.k0);
    w1 = cov1 & (kc < ps.k1);
    ps.k0 = w0 ? ka : ps.k0; ps.c0 = w0 ? col : ps.c0;
    ps.k1 = w1 ? kc : ps.k1; ps.c1 = w1 ? col : ps.c1;
}

// GL_REPEAT + GL_LINEAR lookup of an RGBA8 texture, fp32, texel centres at (i + 0.5) / size
__device__ __forceinline__ void sample_bilinear(const uchar4 *tex, int tw, int th, float u, float v, float *rgb) {
    u -= floorf(u);
    v -= floorf(v);
    const float x = fmaf(u, (float)tw, -0.5f), y = fmaf(v, (float)th, -0.5f);
    const float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    int x0 = (int)xf, y0 = (int)yf;
    x0 = x0 < 0 ? x0 + tw : (x0 >= tw ? x0 - tw : x0);
    y0 = y0 < 0 ? y0 + th : (y0 >= th ? y0 - th : y0);
    const int x1 = x0 + 1 >= tw ? 0 : x0 + 1, y1 = y0 + 1 >= th ? 0 : y0 + 1;
    const uchar4 c00 = __ldg(tex + (size_t)y0 * tw + x0), c10 = __ldg(tex + (size_t)y0 * tw + x1);
    const uchar4 c01 = __ldg(tex + (size_t)y1 * tw + x0), c11 = __ldg(tex + (size_t)y1 * tw + x1);
    const float k = 255.0f;
    const float a00[3] = {c00.x / k, c00.y / k, c00.z / k}, a10[3] = {c10.x / k, c10.y / k, c10.z / k};
    const float a01[3] = {c01.x / k, c01.y / k, c01.z / k}, a11[3] = {c11.x / k, c11.y / k, c11.z / k};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lo = fmaf(fx, a10[c] - a00[c], a00[c]);
        const float hi = fmaf(fx, a11[c] - a01[c], a01[c]);
        rgb[c] = fmaf(fy, hi - lo, lo);
    }
}

// fragment shader for a smooth triangle at one pixel: perspective-correct normal (weights b_i / w_i;
// their normalisation is dropped because the normal is re-normalised), ambient + Lambert
__device__ __forceinline__ unsigned shade_pixel(const FrameDev &f, const SRec &sr, bool textured, float f0, float f1,
                                                float f2, float invA) {
    const float p0 = (f0 * invA) * sr.rw[0], p1 = (f1 * invA) * sr.rw[1], p2 = (f2 * invA) * sr.rw[2];
    float n[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) n[k] = fmaf(p2, sr.n[2][k], fmaf(p1, sr.n[1][k], p0 * sr.n[0][k]));
    const float l2 = fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0]));
    const float inv = 1.0f / sqrtf(l2);
    n[0] *= inv; n[1] *= inv; n[2] *= inv;
    float4 col = make_float4(sr.col[0], sr.col[1], sr.col[2], sr.col[3]);
    if (textured) {
        // basic.frag:31-32: base = mix(1, texture(uv).rgb, useTexture); colour = base * v_color
        const float sum = (p0 + p1) + p2;
        const float u = fmaf(p2, sr.uv[2][0], fmaf(p1, sr.uv[1][0], p0 * sr.uv[0][0])) / sum;
        const float v = fmaf(p2, sr.uv[2][1], fmaf(p1, sr.uv[1][1], p0 * sr.uv[0][1])) / sum;
        float t[3];
        sample_bilinear(sr.tex, sr.tw, sr.th, u, v, t);
        const float a = sr.use_tex, oma = 1.0f - sr.use_tex;
        col.x *= fmaf(t[0], a, oma);
        col.y *= fmaf(t[1], a, oma);
        col.z *= fmaf(t[2], a, oma);
    }
    return shade(f, n, col);
}

// one record, any path.  SMOOTH: records flagged M_SMOOTH shade the pixels they win per pixel.
// CHECK_SLOW = false: the caller knows that none of the records is a slow-path (int64) record
template <bool SMOOTH, bool CHECK_SLOW = true>
__device__ __forceinline__ void raster_one(const FrameDev &f, const Rec &r, const SRec *sr, const int4 &ea,
                                           const int4 &eb, const int4 &ec, const float4 &zq, int px, int py0,
                                           bool ok0, bool ok1, PixelState &ps) {
    bool w0, w1;
    float f0a = 0.f, f1a, f2a, f0b = 0.f, f1b, f2b;
    if (!CHECK_SLOW || !((unsigned)ec.w & M_SLOW)) {
        const FastCov c = fast_cover(ea, eb, ec, px, py0, ok0, ok1);
        if (!__any_sync(0xffffffffu, c.cov0 || c.cov1)) return;
        f1a = (float)c.F1; f2a = (float)c.F2; f1b = (float)c.G1; f2b = (float)c.G2;
        if (SMOOTH) { f0a = (float)c.F0; f0b = (float)c.G0; }
        depth_update(ps, f1a, f2a, f1b, f2b, c.cov0, c.cov1, zq, (unsigned)ec.z, (unsigned)ec.y, w0, w1);
    } else {
        bool cov0, cov1;
        const bool any = slow_cover(r, px, py0, ok0, ok1, cov0, cov1, f1a, f2a, f1b, f2b, f0a, f0b);
        if (!__any_sync(0xffffffffu, any)) return;
        depth_update(ps, f1a, f2a, f1b, f2b, cov0, cov1, zq, (unsigned)ec.z, (unsigned)ec.y, w0, w1);
    }
    if (SMOOTH) {
        if (((unsigned)ec.w & M_SMOOTH) && __any_sync(0xffffffffu, w0 || w1)) {
            const bool tex = ((unsigned)ec.w & M_TEX) != 0;
            const unsigned ca = shade_pixel(f, *sr, tex, f0a, f1a, f2a, zq.w);
            const unsigned cb = shade_pixel(f, *sr, tex, f0b, f1b, f2b, zq.w);
            ps.c0 = w0 ? ca : ps.c0;
            ps.c1 = w1 ? cb : ps.c1;
        }
    }
}

// All records of one block, in draw order.  (A two-records-per-iteration variant was measured:
// slower -- register pressure and the wasted second evaluation on odd counts outweigh the ILP.)
template <int MWORDS, bool SMOOTH = false, bool CHECK_SLOW = true>
__device__ __forceinline__ void raster_block(const Rec *recs, const unsigned *bmask, int px, int py0, bool ok0,
                                             bool ok1, PixelState &ps, const FrameDev *f = nullptr,
                                             const unsigned char *srecs = nullptr) {
#pragma unroll 1
    for (int w = 0; w < MWORDS; ++w) {
        unsigned m = bmask[w];
#pragma unroll 1
        while (m) {
            const int t = w * 32 + 31 - pop_rec(m);
            const Rec &r = recs[t];
            // the whole record in four 128-bit shared loads, issued back to back
            const int4 ea = *reinterpret_cast<const int4 *>(&r.e[0]);         // Eo0 Eo1 Eo2 A0
            const int4 eb = *reinterpret_cast<const int4 *>(&r.e[4]);         // A1 A2 B0 B1
            const int4 ec = *reinterpret_cast<const int4 *>(&r.e[8]);         // B2 col id meta
            const float4 zq = *reinterpret_cast<const float4 *>(&r.z0);       // z0 dz1 dz2 invA
            raster_one<SMOOTH, CHECK_SLOW>(*f, r, SMOOTH ? reinterpret_cast<const SRec *>(srecs + (size_t)t * f->srec_stride) : nullptr,
                               ea, eb, ec, zq, px, py0, ok0, ok1, ps);
        }
    }
}

// The same sweep with 32-bit depth keys, for callers that can promise (a) no int64 records and no per-pixel
// shading among `recs`, (b) the records are visited in draw order (ascending record index == ascending draw id) and
// (c) whatever the pixel state was initialised from was drawn before all of them.  Then "smaller (depth, id) wins"
// is "strictly smaller depth wins": the id never has to be compared or carried, and a pixel was won by this sweep
// iff its depth changed.  Saves 4 of the 10 compare / select instructions per covered (block, record) pair -- they
// all go to the ALU pipe, which is what bounds the small-scene kernel.
// skip the depth part of a (record, block) pair when no lane is covered?  The block reject test of the binning leaves
// few such pairs: the vote and its branch cost more than they save
#ifndef PBR_SWEEP_EARLY_OUT
#define PBR_SWEEP_EARLY_OUT 1
#endif
struct PixelState32 {
    unsigned z0, z1;             // depth bits (non-negative floats order like unsigned integers)
    unsigned c0, c1;             // packed RGBA8 of the current winner
};

// one record of the 32-bit sweep: depth of both pixels, strict less-than, select
__device__ __forceinline__ void update32(PixelState32 &ps, const FastCov &c, const int4 &ec, const float4 &zq) {
    const float za = fmaf((float)c.F2 * zq.w, zq.z, fmaf((float)c.F1 * zq.w, zq.y, zq.x));
    const float zc = fmaf((float)c.G2 * zq.w, zq.z, fmaf((float)c.G1 * zq.w, zq.y, zq.x));
    const unsigned ka = __float_as_uint(za), kc = __float_as_uint(zc);
    const bool w0 = c.cov0 & (ka < ps.z0), w1 = c.cov1 & (kc < ps.z1);
    ps.z0 = w0 ? ka : ps.z0; ps.c0 = w0 ? (unsigned)ec.y : ps.c0;
    ps.z1 = w1 ? kc : ps.z1; ps.c1 = w1 ? (unsigned)ec.y : ps.c1;
}

template <int MWORDS>
__device__ __forceinline__ void raster_block32(unsigned recs_saddr, unsigned bmask_saddr, int px, int py0, bool ok0,
                                               bool ok1, PixelState32 &ps) {
    // (measured slower, see profiles/README.md: two records per iteration, and loading the next record's words
    // while this one is evaluated -- the loop is bound by what it issues, not by its shared-memory loads)
    static_assert(MWORDS == 2, "the block's two mask words are read with one 64-bit load");
    // records and masks by their 32-bit shared-memory addresses: the loads take a register plus an immediate
    unsigned m, m_next;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(m), "=r"(m_next) : "r"(bmask_saddr));
    unsigned last = recs_saddr + 31u * (unsigned)sizeof(Rec);                                // record of bit 0
#pragma unroll 1
    while (true) {
#pragma unroll 1
        while (m) {
            const unsigned a = last + (unsigned)(pop_rec(m) * -(int)sizeof(Rec));
            int4 ea, eb, ec;                                                  // Eo0 Eo1 Eo2 A0 | A1 A2 B0 B1 | B2 col id meta
            asm("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(ea.x), "=r"(ea.y), "=r"(ea.z), "=r"(ea.w) : "r"(a));
            asm("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4+16];" : "=r"(eb.x), "=r"(eb.y), "=r"(eb.z), "=r"(eb.w) : "r"(a));
            asm("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4+32];" : "=r"(ec.x), "=r"(ec.y), "=r"(ec.z), "=r"(ec.w) : "r"(a));
            const FastCov c = fast_cover(ea, eb, ec, px, py0, ok0, ok1);
#if PBR_SWEEP_EARLY_OUT
            if (!__any_sync(0xffffffffu, c.cov0 || c.cov1)) continue;
#endif
            float4 zq;                                                        // z0 dz1 dz2 invA
            asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+48];" : "=f"(zq.x), "=f"(zq.y), "=f"(zq.z), "=f"(zq.w) : "r"(a));
            update32(ps, c, ec, zq);
        }
        if (m_next == 0u) break;
        m = m_next; m_next = 0u;
        last += 32u * (unsigned)sizeof(Rec);
    }
}

__device__ __forceinline__ void put_pixel(unsigned char *color, int plane_stride, int C, int W, int px, int py,
                                          unsigned c) {
    unsigned char *p = color + py * W + px;
    p[0] = (unsigned char)(c & 255u);
    p[plane_stride] = (unsigned char)((c >> 8) & 255u);
    p[2 * plane_stride] = (unsigned char)((c >> 16) & 255u);
    if (C == 4) p[3 * plane_stride] = (unsigned char)(c >> 24);
}

}  // namespace pbr
