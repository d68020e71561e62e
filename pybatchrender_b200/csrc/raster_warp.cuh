// raster_warp.cuh -- the small-scene raster kernel: ONE WARP per scene, no block-level barriers and
// no colour tile in shared memory.
//
// Target: CartPole-class scenes (a few instances of small flat-shaded meshes, <= 48 triangle slots).
// A scene costs only a few thousand warp instructions and ~2 % of its pixels are covered, so the
// design goals are (a) as many resident warps per SM as possible -- the per-scene shared-memory
// footprint is ~7 KB (records, masks, parked vertices), which lets 24+ single-warp CTAs share an SM
// -- and (b) no wasted memory traffic: the background goes straight to out[scene] as 128-bit
// stores issued first (so HBM/L2 work overlaps the geometry phase), and only covered pixels are
// patched afterwards (the lines are still dirty in L2, so DRAM sees each byte once).
//
//   0  background 128-bit stores of the clear colour over out[scene]           (renderer.py:262-264)
//   A  vertices   lanes = (instance, unique vertex): clip = VP*(M*v), outcodes, project + snap,
//                 parked in shared memory                                        (basic.vert:24-43)
//   B1 classify   lanes = triangle slots: trivial reject / needs-clip / back-face cull from the
//                 parked vertices; survivors are compacted with ballots
//   B2 setup      lanes = surviving triangles: integer edge equations, depth plane, flat shade
//                 (basic.frag:31-38) -> 64-byte record; binned into per-8x8-block 64-bit masks
//   B3 clip       lanes = triangles crossing the near plane / guard band: Sutherland-Hodgman, fan
//                 triangles appended to the spare record slots
//   D  raster     non-empty blocks one after the other; every lane owns 2 pixels of the block and
//                 keeps their (depth|id) key and colour in registers across the block's records;
//                 winners are written straight to out[scene] (byte stores, merged in L2)
#pragma once
#include "common.cuh"

namespace pbr {

constexpr int W_MAXREC = 48;     // records per scene (triangle slots that survive + clipped fans); < 64 (mask bits)
constexpr int W_MAXSLOT = 36;    // eligibility: leaves >= 12 spare records for clipped fans
constexpr int W_MAXVERT = 48;    // (instance, vertex) pairs per scene
constexpr int W_MW = 2;            // mask words per block (64 record bits)

__host__ __device__ inline size_t warp_smem_bytes(int nblk) {
    return (size_t)W_MAXVERT * 32 + (size_t)W_MAXREC * sizeof(Rec) + align16((size_t)nblk * W_MW * 4) +
           align16((size_t)nblk * 2) + 2 * W_MAXREC * 4;
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

// fill [dst, dst+n) with byte value v: 128-bit stores on the aligned body
__device__ __forceinline__ void fill_bytes(unsigned char *dst, int n, unsigned v, int lane) {
    const unsigned v4 = v * 0x01010101u;
    const int head = min(n, (int)((16 - (reinterpret_cast<size_t>(dst) & 15)) & 15));
    for (int i = lane; i < head; i += 32) dst[i] = (unsigned char)v;
    const int n16 = (n - head) / 16;
    uint4 *p = reinterpret_cast<uint4 *>(dst + head);
    const uint4 q = make_uint4(v4, v4, v4, v4);
    for (int i = lane; i < n16; i += 32) p[i] = q;
    for (int i = head + n16 * 16 + lane; i < n; i += 32) dst[i] = (unsigned char)v;
}

#ifndef W_MINB
#define W_MINB 32
#endif
__global__ void __launch_bounds__(32, W_MINB) raster_warp_kernel(const __grid_constant__ FrameDev f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int scene = f.scene_begin + (int)blockIdx.x;
    const int nblk = f.nbx * f.nby;
    const int HW = f.H * f.W;
    unsigned char *out_scene = f.out + (size_t)scene * f.C * HW;

    // ---- prefetch this scene's (cold) rows before the output burst occupies the load/store queue
    if (lane < f.n_nodes) {
        const NodeDev &nd = f.nodes[lane];
        const int ninst = min(nd.inst, 8);
        for (int i = 0; i < ninst; ++i) {
            const size_t b = nd.shared ? (size_t)i : (size_t)scene * nd.inst + i;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nd.mats + b * 16));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nd.cols + b * 4));
        }
    } else if (lane == 31) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(f.vp + (size_t)scene * 16));
    }

    // ---- background (or the pre-rendered static layer) straight to global memory.  Issued AFTER the
    // raster phase: scenes finish their (variable amount of) raster work at different times, so the
    // bandwidth-bound output bursts of some warps overlap the issue-bound raster loops of others;
    // covered pixels wait in a small shared-memory patch list until the background is out.
    auto write_background = [&]() {
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
    };
    if (f.debug == 1) { write_background(); return; }
    // ---- carve shared memory
    float4 *clipc = reinterpret_cast<float4 *>(smem_raw);                      // [W_MAXVERT]
    int4 *proj = reinterpret_cast<int4 *>(smem_raw + (size_t)W_MAXVERT * 16);  // [W_MAXVERT]
    Rec *recs = reinterpret_cast<Rec *>(smem_raw + (size_t)W_MAXVERT * 32);
    unsigned *masks = reinterpret_cast<unsigned *>(recs + W_MAXREC);
    unsigned short *blist = reinterpret_cast<unsigned short *>(masks + align16((size_t)nblk * W_MW * 4) / 4);
    unsigned *live = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(blist) + align16((size_t)nblk * 2));
    unsigned *clipl = live + W_MAXREC;

    {   // masks are 8 bytes per block and the region is padded to 16 bytes: clear with 128-bit stores
        uint4 *m4 = reinterpret_cast<uint4 *>(masks);
        const int n16 = (int)(align16((size_t)nblk * W_MW * 4) / 16);
        for (int i = lane; i < n16; i += 32) m4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // which blocks does the static layer cover?  one bit per block, two words for up to 64 blocks
    unsigned bf0 = 0u, bf1 = 0u;
    const bool base_bits = f.base_flags != nullptr && nblk <= 64;
    if (base_bits) {
        bf0 = __ballot_sync(0xffffffffu, lane < nblk && __ldg(f.base_flags + lane) != 0);
        bf1 = __ballot_sync(0xffffffffu, lane + 32 < nblk && __ldg(f.base_flags + lane + 32) != 0);
    }

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
            const int inst = fast_div(local, nd.vert_magic);
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

    // setup + shade + bin one surviving triangle into record j
    auto setup_live = [&](const NodeDev &nd, int inst, int tri, const int4 &q0, const int4 &q1, const int4 &q2, int j) {
        int X[3] = {q0.x, q1.x, q2.x}, Y[3] = {q0.y, q1.y, q2.y};
        float z[3] = {__int_as_float(q0.z), __int_as_float(q1.z), __int_as_float(q2.z)};
        const unsigned id = (unsigned)(nd.id_begin + inst * nd.n_tris + tri) + 1u;
        Rec r;
        BBox bb;
        if (setup_snapped(f, X, Y, z, (nd.flags & PBR_MESH_TWO_SIDED) != 0, id, 0, f.H, r, bb)) {
            const size_t b = nd.shared ? (size_t)inst : (size_t)scene * nd.inst + inst;
            float M[16], n[3];
            load_mat(nd.mats + b * 16, M);
            const float4 n0 = __ldg(nd.tn + 3 * tri);
            xform_normal(M, n0.x, n0.y, n0.z, n);
            r.col = shade(f, n, __ldg(reinterpret_cast<const float4 *>(nd.cols + b * 4)));
            recs[j] = r;
            bin_record<W_MW>(r, bb, j, f.nbx, masks);
        }
    };

    // ---- B1: classify triangle slots.  With <= 32 slots every lane keeps its own slot and goes
    // straight to setup (record index = slot); otherwise survivors are compacted first so that the
    // expensive setup runs on full warps.
    int nlive = 0, nclip = 0;
    const int S = f.total_slots;
    const bool direct = S <= 32;
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
            const int inst = fast_div(local, nd.tri_magic);
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
            if (direct && cat == 1) setup_live(nd, inst, tri, q0, q1, q2, s);
        }
        const unsigned bc = __ballot_sync(0xffffffffu, cat == 2);
        if (cat == 2) clipl[nclip + __popc(bc & lt_mask)] = packed;
        nclip += __popc(bc);
        if (!direct) {
            const unsigned bl = __ballot_sync(0xffffffffu, cat == 1);
            if (cat == 1) live[nlive + __popc(bl & lt_mask)] = packed;
            nlive += __popc(bl);
        }
    }
    __syncwarp();

    // ---- B2 (only when slots were compacted): record index = position in the live list
    if (!direct) {
#pragma unroll 1
        for (int base = 0; base < nlive; base += 32) {
            const int j = base + lane;
            if (j < nlive) {
                const WSlot ws = unpack_slot(live[j]);
                const NodeDev &nd = f.nodes[ws.ni];
                const uint4 ti = __ldg(nd.tidx + ws.tri);
                const int vb = nd.vert_begin + ws.inst * nd.n_verts;
                setup_live(nd, ws.inst, ws.tri, proj[vb + ti.x], proj[vb + ti.y], proj[vb + ti.z], j);
            }
        }
    } else {
        nlive = S;
    }
    int nrec = nlive;

    // ---- B3: clipped triangles -> fan triangles in the spare record slots
    if (nclip > 0) {
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
    }
    __syncwarp();     // records + masks visible to the whole warp; background stores ordered before patches

    if (f.debug == 2) return;
    // ---- D: raster the non-empty blocks
    int nlist = 0;
#pragma unroll 1
    for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int b = b0 + lane;
        bool nz = false;
        int packed = 0;
        if (b < nblk) {
            nz = (masks[b * W_MW] | masks[b * W_MW + 1]) != 0u;
            const int by = fast_div(b, f.nbx_magic);
            packed = (by << 8) | (b - by * f.nbx);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (nz) blist[nlist + __popc(bal & lt_mask)] = (unsigned short)packed;
        nlist += __popc(bal);
    }
    __syncwarp();
    const int lx = lane & 7, ly = lane >> 3;
    uint2 *plist = reinterpret_cast<uint2 *>(smem_raw);          // (pixel offset, RGBA8): aliases the dead vertex scratch
    constexpr int PCAP = W_MAXVERT * 32 / 8;
    int npatch = 0;
    bool bg_done = false;
    auto put = [&](unsigned off, unsigned c) {
        unsigned char *p = out_scene + off;
        p[0] = (unsigned char)(c & 255u);
        p[HW] = (unsigned char)((c >> 8) & 255u);
        p[2 * HW] = (unsigned char)((c >> 16) & 255u);
        if (f.C == 4) p[3 * HW] = (unsigned char)(c >> 24);
    };
    auto flush = [&]() {
        __syncwarp();
        for (int i = lane; i < npatch; i += 32) { const uint2 e = plist[i]; put(e.x, e.y); }
        npatch = 0;
    };
#pragma unroll 1
    for (int i = 0; i < nlist; ++i) {
        const int pk = blist[i];
        const int bx = pk & 255, by = pk >> 8;
        const int px = bx * 8 + lx, py0 = by * 8 + ly;
        const bool ok0 = px < f.W && py0 < f.H, ok1 = px < f.W && py0 + 4 < f.H;
        PixelState ps;
        ps.k0 = ps.k1 = KEY_CLEAR;
        if (f.base_flags != nullptr) {
            const int b = by * f.nbx + bx;
            const bool covered = base_bits ? (((b < 32 ? bf0 : bf1) >> (b & 31)) & 1u) != 0u
                                           : __ldg(f.base_flags + b) != 0;
            if (covered) {                          // the static layer covers part of this block
                ps.k0 = __ldg(f.base_keys + (size_t)b * 64 + lane);
                ps.k1 = __ldg(f.base_keys + (size_t)b * 64 + 32 + lane);
            }
        }
        ps.c0 = ps.c1 = 0u;
        ps.ch0 = ps.ch1 = false;
        raster_block<W_MW>(recs, masks + (by * f.nbx + bx) * W_MW, px, py0, ok0, ok1, ps);
        if (f.debug == 3) continue;
        const unsigned off0 = (unsigned)(py0 * f.W + px), off1 = off0 + 4u * (unsigned)f.W;
        const unsigned b0 = __ballot_sync(0xffffffffu, ps.ch0), b1 = __ballot_sync(0xffffffffu, ps.ch1);
        const int cnt = __popc(b0) + __popc(b1);
        if (!bg_done && npatch + cnt > PCAP) {       // patch list full: background now, direct writes from here on
            write_background();
            flush();
            __syncwarp();
            bg_done = true;
        }
        if (bg_done) {
            if (ps.ch0) put(off0, ps.c0);
            if (ps.ch1) put(off1, ps.c1);
        } else {
            if (ps.ch0) plist[npatch + __popc(b0 & lt_mask)] = make_uint2(off0, ps.c0);
            if (ps.ch1) plist[npatch + __popc(b0) + __popc(b1 & lt_mask)] = make_uint2(off1, ps.c1);
            npatch += cnt;
        }
    }
    if (!bg_done) {
        write_background();
        flush();
    }
}

}  // namespace pbr
