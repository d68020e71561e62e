// raster_warp.cuh -- the small-scene raster kernel: ONE WARP per scene, no block-level barriers.
//
// Target: CartPole-class scenes (a few instances of small flat-shaded meshes, <= 48 triangle slots,
// tile up to ~128x128).  A scene costs only a few thousand warp instructions, so the design goal is
// to keep every issue slot busy: a CTA is a single warp that owns its scene end to end, 10+ such
// CTAs are resident per SM, and there is nothing to wait for except the warp's own memory traffic.
//
//   A  vertices   lanes = (instance, unique vertex): clip = VP*(M*v), outcodes, project + snap,
//                 parked in shared memory (aliasing the not-yet-cleared colour tile)
//   B1 classify   lanes = triangle slots: trivial reject / needs-clip / back-face cull from the
//                 parked vertices; survivors are compacted with ballots
//   B2 setup      lanes = surviving triangles: integer edge equations, depth plane, flat shade ->
//                 64-byte record in shared memory; binned into per-8x8-block 64-bit masks
//   B3 clip       lanes = triangles crossing the near plane / guard band: Sutherland-Hodgman, fan
//                 triangles appended to the spare record slots
//   C  clear      colour tile = background (128-bit shared stores)
//   D  raster     non-empty blocks, one after the other; every lane owns 2 pixels of the block and
//                 keeps their depth / id / colour in registers across the block's records
//   E  store      the finished tile goes out with 128-bit streaming stores into out[scene]
#pragma once
#include "common.cuh"
#include "raster_general.cuh"   // store_band, clear_color

namespace pbr {

constexpr int W_MAXREC = 64;     // records per scene (triangle slots that survive + clipped fans)
constexpr int W_MAXSLOT = 48;    // eligibility: leaves >= 16 spare records for clipped fans
constexpr int W_MAXVERT = 96;    // (instance, vertex) pairs per scene
constexpr int W_MW = W_MAXREC / 32;

__host__ __device__ inline size_t warp_smem_bytes(int C, int plane_stride, int nblk) {
    size_t color = align16((size_t)C * plane_stride);
    const size_t scratch = (size_t)W_MAXVERT * 32;
    if (color < scratch) color = scratch;
    return color + (size_t)W_MAXREC * sizeof(Rec) + align16((size_t)nblk * W_MW * 4) + align16((size_t)nblk * 2) +
           2 * W_MAXREC * 4;
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

// vertex flags
constexpr int VF_CLIP = 0x40, VF_PROJ = 0x80;

__global__ void __launch_bounds__(32) raster_warp_kernel(const __grid_constant__ FrameDev f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int scene = f.scene_begin + (int)blockIdx.x;
    const int nblk = f.nbx * f.nby;

    // ---- carve shared memory
    size_t color_bytes = align16((size_t)f.C * f.plane_stride);
    if (color_bytes < (size_t)W_MAXVERT * 32) color_bytes = (size_t)W_MAXVERT * 32;
    unsigned char *color = smem_raw;
    float4 *clipc = reinterpret_cast<float4 *>(smem_raw);                      // [W_MAXVERT] (aliases colour)
    int4 *proj = reinterpret_cast<int4 *>(smem_raw + (size_t)W_MAXVERT * 16);  // [W_MAXVERT]
    Rec *recs = reinterpret_cast<Rec *>(smem_raw + color_bytes);
    unsigned *masks = reinterpret_cast<unsigned *>(recs + W_MAXREC);
    unsigned short *blist = reinterpret_cast<unsigned short *>(masks + align16((size_t)nblk * W_MW * 4) / 4);
    unsigned *live = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(blist) + align16((size_t)nblk * 2));
    unsigned *clipl = live + W_MAXREC;

    for (int i = lane; i < nblk * W_MW; i += 32) masks[i] = 0u;

    // ---- A: vertices
    {
        float VP[16];
        load_mat(f.vp + (size_t)scene * 16, VP);
#pragma unroll 1
        for (int v = lane; v < f.total_verts; v += 32) {
            int ni = 0;
#pragma unroll 1
            for (int i = 1; i < f.n_nodes; ++i)
                if (v >= f.nodes[i].vert_begin) ni = i;
            const NodeDev &nd = f.nodes[ni];
            const int local = v - nd.vert_begin;
            const int inst = local / nd.n_verts;
            const int vert = local - inst * nd.n_verts;
            const size_t b = nd.shared ? (size_t)inst : (size_t)scene * nd.inst + inst;
            float M[16];
            load_mat(nd.mats + b * 16, M);
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
            if (needs_clip(c)) flags |= VF_CLIP;
            int X = 0, Y = 0;
            float z = 0.0f;
            if (project_vertex(f, c, X, Y, z)) flags |= VF_PROJ;
            clipc[v] = make_float4(c[0], c[1], c[2], c[3]);
            proj[v] = make_int4(X, Y, __float_as_int(z), flags);
        }
    }
    __syncwarp();

    // ---- B1: classify triangle slots
    int nlive = 0, nclip = 0;
    const int S = f.total_slots;
#pragma unroll 1
    for (int base = 0; base < S; base += 32) {
        const int s = base + lane;
        int cat = 0;            // 0 dead, 1 live, 2 clip
        unsigned packed = 0;
        if (s < S) {
            int ni = 0;
#pragma unroll 1
            for (int i = 1; i < f.n_nodes; ++i)
                if (s >= f.nodes[i].slot_begin) ni = i;
            const NodeDev &nd = f.nodes[ni];
            const int local = s - nd.slot_begin;
            const int inst = local / nd.n_tris;
            const int tri = local - inst * nd.n_tris;
            packed = pack_slot(ni, inst, tri);
            const uint4 ti = __ldg(nd.tidx + tri);
            const int vb = nd.vert_begin + inst * nd.n_verts;
            const int4 q0 = proj[vb + ti.x], q1 = proj[vb + ti.y], q2 = proj[vb + ti.z];
            const int f_and = q0.w & q1.w & q2.w, f_or = q0.w | q1.w | q2.w;
            if (f_and & 0x3f) {
                cat = 0;
            } else if (f_or & VF_CLIP) {
                cat = 2;
            } else if (f_and & VF_PROJ) {
                const long long area2 = (long long)(q1.x - q0.x) * (q2.y - q0.y) - (long long)(q2.x - q0.x) * (q1.y - q0.y);
                const bool two_sided = (nd.flags & PBR_MESH_TWO_SIDED) != 0;
                cat = (area2 < 0 || (two_sided && area2 > 0)) ? 1 : 0;
            }
        }
        const unsigned bl = __ballot_sync(0xffffffffu, cat == 1);
        const unsigned bc = __ballot_sync(0xffffffffu, cat == 2);
        if (cat == 1) live[nlive + __popc(bl & lt_mask)] = packed;
        if (cat == 2) clipl[nclip + __popc(bc & lt_mask)] = packed;
        nlive += __popc(bl);
        nclip += __popc(bc);
    }
    __syncwarp();

    // ---- B2: setup + bin the surviving triangles (record index = position in the live list)
#pragma unroll 1
    for (int base = 0; base < nlive; base += 32) {
        const int j = base + lane;
        if (j < nlive) {
            const WSlot ws = unpack_slot(live[j]);
            const NodeDev &nd = f.nodes[ws.ni];
            const uint4 ti = __ldg(nd.tidx + ws.tri);
            const int vb = nd.vert_begin + ws.inst * nd.n_verts;
            const int4 q0 = proj[vb + ti.x], q1 = proj[vb + ti.y], q2 = proj[vb + ti.z];
            int X[3] = {q0.x, q1.x, q2.x}, Y[3] = {q0.y, q1.y, q2.y};
            float z[3] = {__int_as_float(q0.z), __int_as_float(q1.z), __int_as_float(q2.z)};
            const unsigned id = (unsigned)(nd.slot_begin + ws.inst * nd.n_tris + ws.tri) + 1u;
            Rec r;
            BBox bb;
            if (setup_snapped(f, X, Y, z, (nd.flags & PBR_MESH_TWO_SIDED) != 0, id, 0, f.H, r, bb)) {
                const size_t b = nd.shared ? (size_t)ws.inst : (size_t)scene * nd.inst + ws.inst;
                float M[16], n[3];
                load_mat(nd.mats + b * 16, M);
                const float4 n0 = __ldg(nd.tn + 3 * ws.tri);
                xform_normal(M, n0.x, n0.y, n0.z, n);
                r.col = shade(f, n, __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4)));
                recs[j] = r;
                bin_record<W_MW>(r, bb, j, f.nbx, masks);
            }
        }
    }
    int nrec = nlive;

    // ---- B3: clipped triangles -> fan triangles in the spare record slots
    bool overflow = false;
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
            float M[16], n[3];
            load_mat(nd.mats + b * 16, M);
            const float4 n0 = __ldg(nd.tn + 3 * ws.tri);
            xform_normal(M, n0.x, n0.y, n0.z, n);
            col = __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4));
            id = (unsigned)(nd.slot_begin + ws.inst * nd.n_tris + ws.tri) + 1u;
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
        // exclusive prefix sum of cnt over the warp
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int start = nrec + incl - cnt;
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            const int idx = start + k;
            if (idx >= W_MAXREC) { overflow = true; break; }
            int X[3], Y[3];
            float z[3];
            const bool ok = project_vertex(f, poly[0].c, X[0], Y[0], z[0]) &&
                            project_vertex(f, poly[k + 1].c, X[1], Y[1], z[1]) &&
                            project_vertex(f, poly[k + 2].c, X[2], Y[2], z[2]);
            Rec r;
            BBox bb;
            if (ok && setup_snapped(f, X, Y, z, two_sided, id, 0, f.H, r, bb)) {
                r.col = shade(f, poly[0].n, col);
                recs[idx] = r;
                bin_record<W_MW>(r, bb, idx, f.nbx, masks);
            }
        }
        nrec = min(nrec + total, W_MAXREC);
    }
    if (__any_sync(0xffffffffu, overflow) && lane == 0) atomicOr(f.status, DEVSTAT_WARP_OVERFLOW);
    __syncwarp();

    // ---- C: clear the colour tile (the vertex scratch is dead now)
    clear_color(f, color, lane, 32);
    __syncwarp();

    // ---- D: raster the non-empty blocks
    int nlist = 0;
#pragma unroll 1
    for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int b = b0 + lane;
        bool nz = false;
        if (b < nblk) nz = (masks[b * W_MW] | masks[b * W_MW + 1]) != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (nz) blist[nlist + __popc(bal & lt_mask)] = (unsigned short)b;
        nlist += __popc(bal);
    }
    __syncwarp();
    const int lx = lane & 7, ly = lane >> 3;
#pragma unroll 1
    for (int i = 0; i < nlist; ++i) {
        const int b = blist[i];
        const int by = b / f.nbx, bx = b - by * f.nbx;
        const int px = bx * 8 + lx, py0 = by * 8 + ly;
        const bool ok0 = px < f.W && py0 < f.H, ok1 = px < f.W && py0 + 4 < f.H;
        PixelState ps;
        ps.zb0 = ps.zb1 = 0x3F800000u;
        ps.id0 = ps.id1 = 0u;
        ps.c0 = ps.c1 = 0u;
        ps.ch0 = ps.ch1 = false;
        raster_block<W_MW>(recs, masks + b * W_MW, px, py0, ok0, ok1, ps);
        if (ps.ch0) put_pixel(color, f.plane_stride, f.C, f.W, px, py0, ps.c0);
        if (ps.ch1) put_pixel(color, f.plane_stride, f.C, f.W, px, py0 + 4, ps.c1);
    }
    __syncwarp();

    // ---- E: store
    store_band(f, color, scene, 0, f.H, lane, 32);
}

}  // namespace pbr
