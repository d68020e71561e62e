// pbr_b200.cu -- C ABI (include/pbr_b200.h) and host-side launch logic of libpbr_b200.so.
//
// What the library replaces in the reference (dolphin-in-a-coma/pybatchrender):
//   pybatchrender/shaders/basic.vert:24-56   instance decode, clip = VP*(M*v), normal, colour
//   OpenGL fixed function  near clip, divide, viewport, coverage, depth LESS
//   pybatchrender/shaders/basic.frag:20-38   tile scissor, ambient + Lambert
//   renderer/frame_grabber.py:85-106, renderer/renderer.py:352-363   readback, flip, un-tiling
//   renderer/node.py:116-154                 transform composition + upload
//
// Kernels (sm_100a only):
//   raster_warp.cuh      raster_warp_kernel      one warp per scene, small flat-shaded scenes
//   raster_general.cuh   raster_general_kernel   one CTA per (scene, band), everything else
//   transforms.cuh       pack_transforms_kernel / compose_kernel   instance transforms
//
// Arithmetic contract (shared with oracle/pbr_oracle.c, which tests compare against bit for bit):
// IEEE fp32, round-to-nearest, -fmad=false so that only explicit fmaf() calls fuse; coverage is
// exact integer arithmetic on coordinates snapped to 1/256 pixel with a top-left fill rule.
#include "../../include/pbr_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "raster_general.cuh"
#include "raster_warp.cuh"
#include "raster_staged.cuh"
#include "raster_binned.cuh"
#include "transforms.cuh"

using namespace pbr;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};     // kernels of this library enqueued (or captured) by this process
#define COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(PBR_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

unsigned host_unorm8(float c) {
    if (!(c == c)) c = 0.0f;         // fmaxf(NaN, 0) = 0
    c = c < 0.0f ? 0.0f : c;
    c = c > 1.0f ? 1.0f : c;
    volatile float v = c * 255.0f;
    volatile float w = v + 0.5f;
    return (unsigned)(int)w;
}

// Buffers that a frame in flight owns until its kernels have run: one set per (device, stream), so
// that frames issued on different streams do not share them.  (A CUDA graph that captured a frame keeps
// using the set of its capture stream; a later, larger frame on that stream re-allocates it.)
struct StreamScratch {
    // scratch of the geometry pre-pass: per-scene record lists (grown on demand, reused per launch)
    Rec *g_recs = nullptr;
    unsigned char *g_srecs = nullptr;
    size_t g_srec_cap = 0;
    unsigned char *g_vis = nullptr;
    size_t g_vis_cap = 0;
    int *g_bcount = nullptr;
    unsigned *g_bidx = nullptr;
    size_t g_bcount_cap = 0, g_bidx_cap = 0;
    unsigned *g_bbox = nullptr;
    int *g_count = nullptr;
    // block-list path (raster_binned.cuh): parked vertices, per-block counts / offsets, (block, record) pairs
    float4 *b_vclip = nullptr;
    int4 *b_vproj = nullptr;
    int *b_cnt = nullptr, *b_off = nullptr;
    unsigned *b_pairs = nullptr, *b_pbox = nullptr;
    size_t b_vert_cap = 0, b_blk_cap = 0, b_pair_cap = 0, b_pbox_cap = 0;
    size_t g_rec_cap = 0;            // records allocated (all scenes of one launch)
    size_t g_scene_cap = 0;          // scenes allocated in g_count
    // clear-colour image [C,H,W]: copy source of the small-scene kernel's background when there is
    // no static layer (one code path for both; the 12 KB tile stays in L1/L2)
    unsigned char *bgtile = nullptr;
    size_t bgtile_bytes = 0;
    unsigned bg_sig[4] = {0, 0, 0, 0};   // W, H, C, packed colour of the current contents
    // output range of the last small-scene launch on this stream, if nothing else of ours followed it: the
    // next small-scene launch may start while that one drains, and must order itself behind it when the two
    // write overlapping memory (see raster_warp.cuh, "programmatic launch chain")
    const unsigned char *last_out_lo[2] = {nullptr, nullptr}, *last_out_hi[2] = {nullptr, nullptr};
};

struct DeviceState {
    int max_smem_optin = 0;
    int sm_count = 0;
    bool attr_general = false, attr_warp = false, attr_staged = false;
    int *status = nullptr;           // device word with sticky DEVSTAT_* bits
    volatile int *status_host = nullptr;   // host-mapped copy (pinned, zero-copy): polled without a sync
    int *status_host_dev = nullptr;        // device alias of status_host
    std::map<void *, StreamScratch> per_stream;
    // small-scene kernel's record overflow pool: sm_count * W_POOL_PER_SM entries
    Rec *ovf_recs = nullptr;
    unsigned *ovf_masks = nullptr;
    unsigned *ovf_busy = nullptr;
    size_t ovf_mask_words = 0;           // mask words per entry currently allocated
    bool cap_worst_case = false;         // staged path: a frame overflowed its record lists -> size them for 6 per slot
    bool attr_binned = false;
};
// one lock per device: calls only enqueue work, and calls for different GPUs (one host thread per device is the
// expected multi-GPU driver inside one process) do not wait for each other
std::mutex g_mu[64];
DeviceState g_dev[64];

int device_state(int device, DeviceState **out) {
    DeviceState &st = g_dev[device & 63];
    if (st.max_smem_optin == 0) {
        CUDA_TRY(cudaDeviceGetAttribute(&st.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        CUDA_TRY(cudaDeviceGetAttribute(&st.sm_count, cudaDevAttrMultiProcessorCount, device));
        CUDA_TRY(cudaMalloc(&st.status, sizeof(int)));
        {   // %smid ranges over [0, %nsmid), which may exceed the SM count: size per-SM tables by it
            nsmid_kernel<<<1, 1>>>(st.status);
            int nsmid = 0;
            CUDA_TRY(cudaMemcpy(&nsmid, st.status, sizeof(int), cudaMemcpyDeviceToHost));
            if (nsmid > st.sm_count) st.sm_count = nsmid;
        }
        CUDA_TRY(cudaMemset(st.status, 0, sizeof(int)));
        int *hp = nullptr;
        CUDA_TRY(cudaHostAlloc(&hp, 2 * sizeof(int), cudaHostAllocMapped));     // [0] small-scene overflow, [1] staged overflow
        hp[0] = hp[1] = 0;
        CUDA_TRY(cudaHostGetDevicePointer(&st.status_host_dev, hp, 0));
        st.status_host = hp;
    }
    *out = &st;
    return PBR_OK;
}

}  // namespace

struct pbr_mesh_s {
    int device;
    int n_tris;
    int n_verts;        // unique positions
    int all_flat;
    unsigned flags;
    float4 *tp;         // [T*3]
    float4 *tn;         // [T*3]
    float4 *vpos;       // [V]
    uint4 *tidx;        // [T]
    float2 *tuv;        // [T*3] or NULL
    float bsphere[4];   // bounding sphere (centre, radius) of the positions
};

struct pbr_texture_s {
    int device;
    int w, h;
    uchar4 *texels;     // [h, w], row 0 = v 0
};

extern "C" {

int pbr_version(void) { return PBR_B200_VERSION; }

const char *pbr_last_error(void) { return g_err; }

int pbr_mesh_create(const float *pos, const float *nrm, const float *uv, int32_t n_verts, const uint32_t *idx,
                    int32_t n_tris, int32_t device, uint32_t flags, pbr_mesh_t *out) {
    if (!out) return fail(PBR_EINVAL, "pbr_mesh_create: out is NULL");
    *out = nullptr;
    if (!pos || !nrm || !idx || n_verts <= 0 || n_tris <= 0)
        return fail(PBR_EINVAL, "pbr_mesh_create: empty or NULL geometry (n_verts=%d n_tris=%d)", n_verts, n_tris);
    std::vector<float4> tp((size_t)n_tris * 3), tn((size_t)n_tris * 3), vpos;
    std::vector<uint4> tidx((size_t)n_tris);
    std::vector<float2> tuv(uv ? (size_t)n_tris * 3 : 0);
    struct Key {
        uint32_t a, b, c;
        bool operator<(const Key &o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : c < o.c); }
    };
    std::map<Key, uint32_t> uniq;
    int all_flat = 1;
    for (int t = 0; t < n_tris; ++t) {
        bool flat = true;
        uint32_t ui[3];
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = idx[3 * t + k];
            if (v >= (uint32_t)n_verts) return fail(PBR_EINVAL, "pbr_mesh_create: index %u out of range", v);
            tp[3 * t + k] = make_float4(pos[3 * v], pos[3 * v + 1], pos[3 * v + 2], 0.0f);
            tn[3 * t + k] = make_float4(nrm[3 * v], nrm[3 * v + 1], nrm[3 * v + 2], 0.0f);
            if (uv) tuv[3 * t + k] = make_float2(uv[2 * (size_t)v], uv[2 * (size_t)v + 1]);
            if (memcmp(nrm + 3 * (size_t)v, nrm + 3 * (size_t)idx[3 * t], 12) != 0) flat = false;
            Key key;
            memcpy(&key, pos + 3 * (size_t)v, 12);
            auto it = uniq.find(key);
            if (it == uniq.end()) {
                it = uniq.emplace(key, (uint32_t)vpos.size()).first;
                vpos.push_back(make_float4(pos[3 * v], pos[3 * v + 1], pos[3 * v + 2], 1.0f));
            }
            ui[k] = it->second;
        }
        const int flag = flat ? 1 : 0;
        memcpy(&tp[3 * t].w, &flag, 4);
        tidx[t] = make_uint4(ui[0], ui[1], ui[2], (unsigned)flag);
        if (!flat) all_flat = 0;
    }
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(device));
    pbr_mesh_s *m = new (std::nothrow) pbr_mesh_s();
    if (!m) { cudaSetDevice(prev); return fail(PBR_ENOMEM, "pbr_mesh_create: host allocation failed"); }
    memset(m, 0, sizeof(*m));
    {   // bounding sphere: centre of the bounding box, radius to the farthest vertex (rounded up)
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int v = 0; v < n_verts; ++v)
            for (int a = 0; a < 3; ++a) {
                lo[a] = pos[3 * v + a] < lo[a] ? pos[3 * v + a] : lo[a];
                hi[a] = pos[3 * v + a] > hi[a] ? pos[3 * v + a] : hi[a];
            }
        float c[3];
        for (int a = 0; a < 3; ++a) c[a] = (float)(0.5 * (lo[a] + hi[a]));
        double r2 = 0.0;
        for (int v = 0; v < n_verts; ++v) {
            double d2 = 0.0;
            for (int a = 0; a < 3; ++a) d2 += ((double)pos[3 * v + a] - c[a]) * ((double)pos[3 * v + a] - c[a]);
            r2 = d2 > r2 ? d2 : r2;
        }
        m->bsphere[0] = c[0]; m->bsphere[1] = c[1]; m->bsphere[2] = c[2];
        m->bsphere[3] = (float)(sqrt(r2) * 1.00001) + 1e-30f;
    }
    m->device = device; m->n_tris = n_tris; m->n_verts = (int)vpos.size(); m->all_flat = all_flat; m->flags = flags;
    const size_t tb = (size_t)n_tris * 3 * sizeof(float4), vb = vpos.size() * sizeof(float4), ib = (size_t)n_tris * sizeof(uint4);
    cudaError_t e = cudaMalloc(&m->tp, tb);
    if (e == cudaSuccess) e = cudaMalloc(&m->tn, tb);
    if (e == cudaSuccess) e = cudaMalloc(&m->vpos, vb);
    if (e == cudaSuccess) e = cudaMalloc(&m->tidx, ib);
    if (e == cudaSuccess) e = cudaMemcpy(m->tp, tp.data(), tb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->tn, tn.data(), tb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->vpos, vpos.data(), vb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->tidx, tidx.data(), ib, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && uv) e = cudaMalloc(&m->tuv, tuv.size() * sizeof(float2));
    if (e == cudaSuccess && uv) e = cudaMemcpy(m->tuv, tuv.data(), tuv.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaFree(m->tp); cudaFree(m->tn); cudaFree(m->vpos); cudaFree(m->tidx); cudaFree(m->tuv);
        delete m;
        return fail(e == cudaErrorMemoryAllocation ? PBR_ENOMEM : PBR_ECUDA, "pbr_mesh_create: %s", cudaGetErrorString(e));
    }
    *out = m;
    return PBR_OK;
}

int pbr_mesh_destroy(pbr_mesh_t m) {
    if (!m) return PBR_OK;
    cudaFree(m->tp); cudaFree(m->tn); cudaFree(m->vpos); cudaFree(m->tidx); cudaFree(m->tuv);
    delete m;
    return PBR_OK;
}

int pbr_texture_create(const uint8_t *rgba, int32_t width, int32_t height, int32_t device, pbr_texture_t *out) {
    if (!out) return fail(PBR_EINVAL, "pbr_texture_create: out is NULL");
    *out = nullptr;
    if (!rgba || width < 1 || height < 1 || width > 16384 || height > 16384)
        return fail(PBR_EINVAL, "pbr_texture_create: bad image (%p, %dx%d)", (const void *)rgba, width, height);
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(device));
    pbr_texture_s *t = new (std::nothrow) pbr_texture_s();
    if (!t) { cudaSetDevice(prev); return fail(PBR_ENOMEM, "pbr_texture_create: host allocation failed"); }
    t->device = device; t->w = width; t->h = height; t->texels = nullptr;
    const size_t bytes = (size_t)width * height * 4;
    cudaError_t e = cudaMalloc(&t->texels, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(t->texels, rgba, bytes, cudaMemcpyHostToDevice);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaFree(t->texels);
        delete t;
        return fail(e == cudaErrorMemoryAllocation ? PBR_ENOMEM : PBR_ECUDA, "pbr_texture_create: %s", cudaGetErrorString(e));
    }
    *out = t;
    return PBR_OK;
}

int pbr_texture_destroy(pbr_texture_t t) {
    if (!t) return PBR_OK;
    cudaFree(t->texels);
    delete t;
    return PBR_OK;
}

int pbr_mesh_info(pbr_mesh_t m, int32_t *n_tris, int32_t *all_flat, int32_t *device) {
    if (!m) return fail(PBR_EINVAL, "pbr_mesh_info: NULL mesh");
    if (n_tris) *n_tris = m->n_tris;
    if (all_flat) *all_flat = m->all_flat;
    if (device) *device = m->device;
    return PBR_OK;
}

// Launch with programmatic stream serialisation: the kernel may begin once the preceding kernel in
// the stream has triggered launch_dependents (only compose_kernel does; after any other kernel this
// is an ordinary launch) and orders itself with griddepcontrol.wait.
static cudaError_t launch_dependent(void (*kernel)(FrameDev), unsigned grid, unsigned block, size_t smem, void *stream,
                                    const FrameDev &f) {
    static const bool off = getenv("PBR_B200_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, f);
}

static void pose_to_dev(const pbr_pose_desc &p, PoseDev &o) {
    for (int k = 0; k < 3; ++k) { o.pos[k] = p.pos[k]; o.hpr[k] = p.hpr[k]; }
    o.scale = p.scale;
    o.out_mats = p.out_mats;
}

static int launch_compose(const pbr_pose_desc *const *poses, int n_poses, void *stream) {
    for (int base = 0; base < n_poses; base += MAX_POSES) {
        PoseBatch pb;
        pb.n = n_poses - base < MAX_POSES ? n_poses - base : MAX_POSES;
        int max_n = 0;
        for (int i = 0; i < pb.n; ++i) {
            const pbr_pose_desc &p = *poses[base + i];
            if (!p.out_mats || (reinterpret_cast<size_t>(p.out_mats) & 15))
                return fail(PBR_EINVAL, "pbr_compose_transforms: pose %d out_mats NULL or not 16-byte aligned", base + i);
            if (p.n_instances < 0) return fail(PBR_EINVAL, "pbr_compose_transforms: pose %d n_instances < 0", base + i);
            pose_to_dev(p, pb.p[i]);
            pb.n_inst[i] = p.n_instances;
            if (p.n_instances > max_n) max_n = p.n_instances;
        }
        if (max_n == 0) continue;
        dim3 grid((max_n + 127) / 128, pb.n);
        compose_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(pb);
        COUNT_LAUNCH();
        CUDA_TRY(cudaGetLastError());
    }
    return PBR_OK;
}

// write the matrices of every posed node of the frame to its pose->out_mats (paths that read matrix buffers)
static int materialise_poses(const pbr_frame_desc *d, void *stream) {
    std::vector<const pbr_pose_desc *> ptrs;
    for (int i = 0; i < d->n_nodes; ++i)
        if (d->nodes[i].pose && d->nodes[i].instances_per_scene > 0) ptrs.push_back(d->nodes[i].pose);
    return ptrs.empty() ? PBR_OK : launch_compose(ptrs.data(), (int)ptrs.size(), stream);
}

enum NodeMode { NODES_ALL = 0, NODES_SKIP_BASE = 1, NODES_SHARED_ONLY = 2 };

struct NodeStats {
    long long slots = 0, verts = 0, insts = 0;
    bool any_smooth = false, any_textured = false, warp_ok = true;
    int skipped = 0;
    long long skipped_id_end = 0;         // draw index past the last triangle of the skipped (static layer) nodes
    long long active_id_begin = 1ll << 40;   // draw index of the first triangle of the active nodes
    int n_poses = 0;          // posed nodes among the active ones (all of them in f.poses when <= MAX_FRAME_POSES)
    bool poses_in_frame = false;
};



// Fill f.nodes from the host descriptors.  Draw indices (id_begin) always count every node so
// that a frame split into static layer + per-scene part numbers its triangles like the full frame.
static int fill_nodes(const pbr_frame_desc *d, int device, NodeMode mode, FrameDev &f, NodeStats &st) {
    long long ids = 0;
    f.n_nodes = 0;
    st = NodeStats();
    for (int i = 0; i < d->n_nodes; ++i) {
        const pbr_node_desc &n = d->nodes[i];
        if (!n.mesh) return fail(PBR_EINVAL, "pbr_render: node %d has no mesh", i);
        if (n.mesh->device != device) return fail(PBR_EINVAL, "pbr_render: node %d mesh lives on device %d, current device is %d", i, n.mesh->device, device);
        const float *mats = n.pose ? n.pose->out_mats : n.mats;
        if (!mats || !n.cols) return fail(PBR_EINVAL, "pbr_render: node %d mats%s / cols is NULL", i, n.pose ? " (pose->out_mats)" : "");
        if ((reinterpret_cast<size_t>(mats) & 15) || (reinterpret_cast<size_t>(n.cols) & 15))
            return fail(PBR_EINVAL, "pbr_render: node %d mats / cols must be 16-byte aligned", i);
        if (n.pose) {
            const long long B = n.shared ? (long long)n.instances_per_scene : (long long)d->num_scenes * n.instances_per_scene;
            if (n.pose->n_instances != B)
                return fail(PBR_EINVAL, "pbr_render: node %d pose has %d instances, node has %lld", i, n.pose->n_instances, B);
        }
        if (n.instances_per_scene < 0) return fail(PBR_EINVAL, "pbr_render: node %d instances_per_scene < 0", i);
        if (n.texture && n.texture->device != device) return fail(PBR_EINVAL, "pbr_render: node %d texture lives on device %d, current device is %d", i, n.texture->device, device);
        const long long ntri = (long long)n.instances_per_scene * n.mesh->n_tris;
        const long long id_begin = ids;
        ids += ntri;
        if (ids > (1ll << 30)) return fail(PBR_EUNSUPPORTED, "pbr_render: more than 2^30 triangles per scene");
        if (n.instances_per_scene == 0) continue;
        const bool in_base = (n.flags & PBR_NODE_IN_BASE) != 0;
        if ((mode == NODES_SKIP_BASE && in_base) || (mode == NODES_SHARED_ONLY && !(n.shared && in_base))) {
            st.skipped++;
            if (ids > st.skipped_id_end) st.skipped_id_end = ids;
            continue;
        }
        if (id_begin < st.active_id_begin) st.active_id_begin = id_begin;
        NodeDev &nd = f.nodes[f.n_nodes++];
        nd.tp = n.mesh->tp; nd.tn = n.mesh->tn; nd.vpos = n.mesh->vpos; nd.tidx = n.mesh->tidx;
        // basic.frag:31-32: base = mix(1, texel, clamp(useTexture)); without a bound image the node is untextured
        float ut = n.use_texture;
        ut = !(ut == ut) || ut < 0.0f ? 0.0f : (ut > 1.0f ? 1.0f : ut);
        const bool textured = n.texture != nullptr && ut > 0.0f;
        nd.tuv = textured ? n.mesh->tuv : nullptr;
        nd.tex = textured ? n.texture->texels : nullptr;
        nd.tw = textured ? n.texture->w : 0; nd.th = textured ? n.texture->h : 0;
        nd.use_tex = textured ? ut : 0.0f;
        if (textured) { st.any_smooth = true; st.any_textured = true; }      // per-pixel shading path
        nd.mats = mats; nd.cols = n.cols;
        nd.pose_idx = -1;
        if (n.pose) {
            if (st.n_poses < MAX_FRAME_POSES) {
                nd.pose_idx = st.n_poses;
                pose_to_dev(*n.pose, f.poses[st.n_poses]);
            }
            st.n_poses++;
        }
        nd.n_tris = n.mesh->n_tris; nd.n_verts = n.mesh->n_verts;
        nd.inst = n.instances_per_scene; nd.shared = n.shared ? 1 : 0;
        nd.slot_begin = (int)st.slots; nd.vert_begin = (int)st.verts; nd.flags = n.mesh->flags;
        nd.id_begin = (int)id_begin;
        nd.bsphere = make_float4(n.mesh->bsphere[0], n.mesh->bsphere[1], n.mesh->bsphere[2], n.mesh->bsphere[3]);
        nd.inst_begin = (int)st.insts;
        st.insts += n.instances_per_scene;
        nd.tri_magic = div_magic((unsigned)n.mesh->n_tris);
        nd.vert_magic = div_magic((unsigned)n.mesh->n_verts);
        st.slots += ntri;
        st.verts += (long long)n.instances_per_scene * n.mesh->n_verts;
        if (!n.mesh->all_flat) st.any_smooth = true;
        if (n.instances_per_scene >= 8192 || n.mesh->n_tris >= 8192) st.warp_ok = false;
    }
    st.poses_in_frame = st.n_poses > 0 && st.n_poses <= MAX_FRAME_POSES;
    if (!st.poses_in_frame)
        for (int i = 0; i < f.n_nodes; ++i) f.nodes[i].pose_idx = -1;
    f.smooth = st.any_smooth ? 1 : 0;
    f.srec_stride = st.any_textured ? SREC_TEXTURED : SREC_PLAIN;
    if (st.insts > 0x7fffffffll) return fail(PBR_EUNSUPPORTED, "pbr_render: more than 2^31 instances per scene");
    f.total_inst = (int)st.insts;
    f.total_slots = (int)st.slots;
    f.total_verts = (int)(st.verts > 0x7fffffff ? 0x7fffffff : st.verts);
    return PBR_OK;
}

struct pbr_base_s {
    int device;
    int W, H, C;            // layout of the last successful pbr_base_render (0 = none yet)
    size_t color_bytes, key_blocks;
    unsigned char *color;
    unsigned long long *keys;
    unsigned char *flags;
};

static int check_frame(const pbr_frame_desc *d, const char *who, bool need_out) {
    if (!d) return fail(PBR_EINVAL, "%s: NULL frame", who);
    if (d->tile_w < 1 || d->tile_h < 1 || d->tile_w > PBR_MAX_TILE || d->tile_h > PBR_MAX_TILE)
        return fail(PBR_EINVAL, "%s: tile %dx%d outside [1,%d]", who, d->tile_w, d->tile_h, PBR_MAX_TILE);
    if (d->channels != 3 && d->channels != 4) return fail(PBR_EINVAL, "%s: channels must be 3 or 4, got %d", who, d->channels);
    if (d->scene_begin < 0 || d->scene_count < 0 || (long long)d->scene_begin + d->scene_count > d->num_scenes)
        return fail(PBR_EINVAL, "%s: scene window [%d,+%d) outside [0,%d)", who, d->scene_begin, d->scene_count, d->num_scenes);
    if (!d->vp || (need_out && !d->out)) return fail(PBR_EINVAL, "%s: out / vp is NULL", who);
    if ((need_out && (reinterpret_cast<size_t>(d->out) & 15)) || (reinterpret_cast<size_t>(d->vp) & 15))
        return fail(PBR_EINVAL, "%s: out and vp must be 16-byte aligned", who);
    if (d->n_nodes < 0 || d->n_nodes > PBR_MAX_NODES) return fail(PBR_EUNSUPPORTED, "%s: %d nodes (max %d)", who, d->n_nodes, PBR_MAX_NODES);
    if (d->n_nodes > 0 && !d->nodes) return fail(PBR_EINVAL, "%s: nodes is NULL", who);
    return PBR_OK;
}

static void fill_uniforms(const pbr_frame_desc *d, FrameDev &f) {
    memset(&f, 0, sizeof(f));
    f.vp = d->vp; f.out = d->out;
    f.vp_scene_override = -1;
    f.scene_begin = d->scene_begin; f.scene_count = d->scene_count;
    f.W = d->tile_w; f.H = d->tile_h; f.C = d->channels;
    f.hw = 0.5f * (float)d->tile_w; f.hh = 0.5f * (float)d->tile_h;
    f.bg = host_unorm8(d->bg[0]) | (host_unorm8(d->bg[1]) << 8) | (host_unorm8(d->bg[2]) << 16) | (host_unorm8(d->bg[3]) << 24);
    // normalize(dirLightDir) and clamp(strength): same fp32 operations as the oracle's make_light
    volatile float x = d->dir_dir[0], y = d->dir_dir[1], z = d->dir_dir[2];
    float l2 = fmaf(z, z, fmaf(y, y, x * x));
    float inv = 1.0f / sqrtf(l2);
    f.ldir[0] = x * inv; f.ldir[1] = y * inv; f.ldir[2] = z * inv;
    float s = d->strength;
    if (!(s == s)) s = 0.0f;
    s = s < 0.0f ? 0.0f : s;
    s = s > 1.0f ? 1.0f : s;
    f.s = s; f.oms = 1.0f - s;
    for (int c = 0; c < 3; ++c) { f.amb[c] = d->ambient[c]; f.dcol[c] = d->dir_col[c]; }
}

// general kernel: one CTA per (scene, band of rows).  plan = choose the band height.
static int plan_general(FrameDev &f, DeviceState *st, size_t *smem_out) {
    const int W = f.W, H = f.H;
    const int nbx = (W + 7) / 8;
    const int H8 = ((H + 7) / 8) * 8;
    auto bytes_for = [&](int BH) {
        const int nby = (BH + 7) / 8;
        const int ps = (int)align16((size_t)BH * W);
        return general_smem_bytes(f.C, ps, nbx * nby, f.smooth ? f.srec_stride : 0);
    };
    static const size_t budget_kb = getenv("PBR_B200_BAND_KB") ? (size_t)atoi(getenv("PBR_B200_BAND_KB")) : 56;   // band-height experiments
    const size_t budget = budget_kb * 1024;
    int BH = H8;
    while (BH > 8 && bytes_for(BH) > budget) BH -= 8;
    if (bytes_for(BH) > (size_t)st->max_smem_optin)
        return fail(PBR_EUNSUPPORTED, "pbr_render: tile width %d needs %zu bytes of shared memory per 8-row band", W, bytes_for(BH));
    f.BH = BH;
    f.nbands = (H + BH - 1) / BH;
    f.nbx = nbx;
    f.nbx_magic = div_magic((unsigned)nbx);
    f.nby = BH / 8;
    if (f.nbx > 256 || f.nby > 256) return fail(PBR_EUNSUPPORTED, "pbr_render: more than 256 blocks per band side");
    f.plane_stride = (int)align16((size_t)BH * W);
    f.linear = (f.nbands == 1 && f.plane_stride == H * W) ? 1 : 0;
    *smem_out = bytes_for(BH);
    return PBR_OK;
}

static int launch_general(FrameDev &f, DeviceState *st, void *stream) {
    size_t smem = 0;
    if (f.BH == 0)
        if (int rc = plan_general(f, st, &smem)) return rc;
    smem = general_smem_bytes(f.C, f.plane_stride, f.nbx * f.nby, f.smooth ? f.srec_stride : 0);
    if (!st->attr_general) {
        CUDA_TRY(cudaFuncSetAttribute(raster_general_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(raster_general_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem_optin));
        st->attr_general = true;
    }
    const long long grid = (long long)f.scene_count * f.nbands;
    if (grid > 0x7fffffffll) return fail(PBR_EUNSUPPORTED, "pbr_render: grid too large");
    if (f.smooth)
        raster_general_kernel<true><<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(f);
    else
        raster_general_kernel<false><<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(f);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return PBR_OK;
}

// geometry pre-pass + TMA-staged raster, in launches of as many scenes as the scratch budget holds
static int launch_staged(FrameDev &f, DeviceState *st, void *stream) {
    StreamScratch *ss = &st->per_stream[stream];
    // A scene's record list holds 1.5 records per triangle slot (+ 64): only clipped triangles need more than
    // one, up to 6.  If an earlier frame ran out (geom_kernel raised the flag in host-mapped memory and dropped
    // records) that frame is wrong: say so once, and size the lists for the worst case from now on.
    if (st->status_host[1] != 0) {
        st->status_host[1] = 0;
        st->cap_worst_case = true;
        return fail(PBR_EOVERFLOW, "pbr_render: an earlier large-scene frame on this device dropped triangles (more than 1.5 "
                                   "records per triangle slot after clipping); record capacity is now sized for the worst case -- render again");
    }
    const size_t cap = st->cap_worst_case ? (((size_t)f.total_slots * 6 + 64 + 3) & ~(size_t)3)
                                          : (((size_t)f.total_slots + (size_t)f.total_slots / 2 + 64 + 3) & ~(size_t)3);
    static const size_t budget_mb = getenv("PBR_B200_SCRATCH_MB") ? (size_t)atoll(getenv("PBR_B200_SCRATCH_MB")) : 4096;
    const bool smooth = f.smooth != 0;
    // per-band index lists pay when a tile has several bands (each band CTA would otherwise scan
    // all of the scene's records)
    static const int band_lists_min = getenv("PBR_B200_BAND_LISTS_MIN") ? atoi(getenv("PBR_B200_BAND_LISTS_MIN")) : 2;
    const bool band_lists = f.nbands >= band_lists_min;
    const size_t per_scene = cap * (sizeof(Rec) + 4 + (smooth ? (size_t)f.srec_stride : 0) + (band_lists ? (size_t)f.nbands * 4 : 0));
    size_t per_launch = (budget_mb << 20) / per_scene;
    if (per_launch < 1) per_launch = 1;
    if (per_launch > (size_t)f.scene_count) per_launch = (size_t)f.scene_count;
    if (per_launch > 65535) per_launch = 65535;                     // gridDim.y of the geometry kernel
    if (per_launch * cap > ss->g_rec_cap || per_launch > ss->g_scene_cap || (smooth && per_launch * cap * (size_t)f.srec_stride > ss->g_srec_cap)) {
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        cudaFree(ss->g_recs); cudaFree(ss->g_bbox); cudaFree(ss->g_count); cudaFree(ss->g_srecs);
        ss->g_recs = nullptr; ss->g_bbox = nullptr; ss->g_count = nullptr; ss->g_srecs = nullptr;
        ss->g_rec_cap = 0; ss->g_scene_cap = 0; ss->g_srec_cap = 0;
        CUDA_TRY(cudaMalloc(&ss->g_recs, per_launch * cap * sizeof(Rec)));
        CUDA_TRY(cudaMalloc(&ss->g_bbox, per_launch * cap * 4 + 64));
        CUDA_TRY(cudaMalloc(&ss->g_count, per_launch * sizeof(int)));
        ss->g_rec_cap = per_launch * cap; ss->g_scene_cap = per_launch;
        if (smooth) {
            CUDA_TRY(cudaMalloc(&ss->g_srecs, per_launch * cap * (size_t)f.srec_stride));
            ss->g_srec_cap = per_launch * cap * (size_t)f.srec_stride;
        }
    }
    if (per_launch * (size_t)f.total_inst > ss->g_vis_cap) {
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        cudaFree(ss->g_vis);
        ss->g_vis = nullptr; ss->g_vis_cap = 0;
        CUDA_TRY(cudaMalloc(&ss->g_vis, per_launch * (size_t)f.total_inst));
        ss->g_vis_cap = per_launch * (size_t)f.total_inst;
    }
    if (band_lists && (per_launch * (size_t)f.nbands > ss->g_bcount_cap || per_launch * (size_t)f.nbands * cap > ss->g_bidx_cap)) {
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        cudaFree(ss->g_bcount); cudaFree(ss->g_bidx);
        ss->g_bcount = nullptr; ss->g_bidx = nullptr; ss->g_bcount_cap = ss->g_bidx_cap = 0;
        CUDA_TRY(cudaMalloc(&ss->g_bcount, per_launch * (size_t)f.nbands * sizeof(int)));
        CUDA_TRY(cudaMalloc(&ss->g_bidx, per_launch * (size_t)f.nbands * cap * sizeof(unsigned)));
        ss->g_bcount_cap = per_launch * (size_t)f.nbands; ss->g_bidx_cap = per_launch * (size_t)f.nbands * cap;
    }
    const size_t smem = staged_smem_bytes(f.C, f.plane_stride, f.nbx * f.nby, smooth ? f.srec_stride : 0);
    if (smem > (size_t)st->max_smem_optin) return launch_general(f, st, stream);
    if (!st->attr_staged) {
        CUDA_TRY(cudaFuncSetAttribute(raster_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(raster_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem_optin));
        st->attr_staged = true;
    }
    StagedDev g;
    g.vis = ss->g_vis;
    g.bcount = band_lists ? ss->g_bcount : nullptr;
    g.bidx = band_lists ? ss->g_bidx : nullptr;
    g.recs = ss->g_recs; g.srecs = smooth ? ss->g_srecs : nullptr; g.bbox = ss->g_bbox; g.count = ss->g_count; g.cap = (int)cap;
    const int first = f.scene_begin, last = f.scene_begin + f.scene_count;
    for (int s0 = first; s0 < last; s0 += (int)per_launch) {
        const int n = (int)((size_t)(last - s0) < per_launch ? (size_t)(last - s0) : per_launch);
        g.scene0 = s0;
        CUDA_TRY(cudaMemsetAsync(ss->g_count, 0, (size_t)n * sizeof(int), (cudaStream_t)stream));
        if (band_lists) CUDA_TRY(cudaMemsetAsync(ss->g_bcount, 0, (size_t)n * f.nbands * sizeof(int), (cudaStream_t)stream));
        dim3 cgrid((unsigned)((f.total_inst + 255) / 256), (unsigned)n);
        static const bool no_cull = getenv("PBR_B200_NO_CULL") != nullptr;          // A/B timing aid
        if (no_cull) {
            CUDA_TRY(cudaMemsetAsync(ss->g_vis, 1, (size_t)n * f.total_inst, (cudaStream_t)stream));
        } else {
            cull_kernel<<<cgrid, 256, 0, (cudaStream_t)stream>>>(f, g);
            COUNT_LAUNCH();
            CUDA_TRY(cudaGetLastError());
        }
        dim3 ggrid((unsigned)((f.total_slots + G_THREADS - 1) / G_THREADS), (unsigned)n);
        geom_kernel<<<ggrid, G_THREADS, 0, (cudaStream_t)stream>>>(f, g);
        COUNT_LAUNCH();
        CUDA_TRY(cudaGetLastError());
        const long long grid = (long long)n * f.nbands;
        if (grid > 0x7fffffffll) return fail(PBR_EUNSUPPORTED, "pbr_render: grid too large");
        if (smooth)
            raster_staged_kernel<true><<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(f, g);
        else
            raster_staged_kernel<false><<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(f, g);
        COUNT_LAUNCH();
        CUDA_TRY(cudaGetLastError());
    }
    return PBR_OK;
}

// Large scenes, block-list path (raster_binned.cuh): cull, vertices, triangles + block counts, scan, fill, one warp
// per block -- in launches of as many scenes as the scratch budget holds.
static int launch_binned(FrameDev &f, DeviceState *st, void *stream) {
    StreamScratch *ss = &st->per_stream[stream];
    cudaStream_t cs = (cudaStream_t)stream;
    if (st->status_host[1] != 0) {
        st->status_host[1] = 0;
        st->cap_worst_case = true;
        return fail(PBR_EOVERFLOW, "pbr_render: an earlier large-scene frame on this device dropped triangles (record or block "
                                   "lists too small); later frames take the band-based path with worst-case sizes -- render again");
    }
    const int H8 = ((f.H + 7) / 8) * 8;
    f.BH = H8; f.nbands = 1; f.nbx = (f.W + 7) / 8; f.nby = H8 / 8;       // blocks of the whole tile
    f.nbx_magic = div_magic((unsigned)f.nbx);
    f.plane_stride = f.H * f.W; f.linear = 1;
    const size_t nblk = (size_t)f.nbx * f.nby;
    const size_t cap = ((size_t)f.total_slots + (size_t)f.total_slots / 2 + 64 + 3) & ~(size_t)3;
    const size_t pairs_cap = 4 * cap + 4 * nblk;
    const size_t tv = (size_t)f.total_verts;
    const bool smooth = f.smooth != 0;
    static const size_t budget_mb = getenv("PBR_B200_SCRATCH_MB") ? (size_t)atoll(getenv("PBR_B200_SCRATCH_MB")) : 4096;
    const size_t per_scene = cap * (sizeof(Rec) + 8 + (smooth ? (size_t)f.srec_stride : 0)) + tv * 32 + nblk * 16 + 4 + pairs_cap * 4 +
                             (size_t)f.total_inst + 4;
    size_t per_launch = (budget_mb << 20) / per_scene;
    if (per_launch < 1) per_launch = 1;
    if (per_launch > (size_t)f.scene_count) per_launch = (size_t)f.scene_count;
    if (per_launch > 65535) per_launch = 65535;                     // gridDim.y
    auto grow = [&](void **p, size_t *have, size_t need_elems, size_t elem) -> cudaError_t {
        if (need_elems <= *have) return cudaSuccess;
        cudaError_t e = cudaStreamSynchronize(cs);
        if (e != cudaSuccess) return e;
        cudaFree(*p);
        *p = nullptr; *have = 0;
        e = cudaMalloc(p, need_elems * elem);
        if (e == cudaSuccess) *have = need_elems;
        return e;
    };
    if (per_launch * cap > ss->g_rec_cap || per_launch > ss->g_scene_cap || (smooth && per_launch * cap * (size_t)f.srec_stride > ss->g_srec_cap)) {
        CUDA_TRY(cudaStreamSynchronize(cs));
        cudaFree(ss->g_recs); cudaFree(ss->g_bbox); cudaFree(ss->g_count); cudaFree(ss->g_srecs);
        ss->g_recs = nullptr; ss->g_bbox = nullptr; ss->g_count = nullptr; ss->g_srecs = nullptr;
        ss->g_rec_cap = 0; ss->g_scene_cap = 0; ss->g_srec_cap = 0;
        CUDA_TRY(cudaMalloc(&ss->g_recs, per_launch * cap * sizeof(Rec)));
        CUDA_TRY(cudaMalloc(&ss->g_bbox, per_launch * cap * 4 + 64));
        CUDA_TRY(cudaMalloc(&ss->g_count, per_launch * sizeof(int)));
        ss->g_rec_cap = per_launch * cap; ss->g_scene_cap = per_launch;
        if (smooth) {
            CUDA_TRY(cudaMalloc(&ss->g_srecs, per_launch * cap * (size_t)f.srec_stride));
            ss->g_srec_cap = per_launch * cap * (size_t)f.srec_stride;
        }
    }
    CUDA_TRY(grow((void **)&ss->g_vis, &ss->g_vis_cap, per_launch * (size_t)f.total_inst, 1));
    if (per_launch * tv > ss->b_vert_cap) {
        CUDA_TRY(cudaStreamSynchronize(cs));
        cudaFree(ss->b_vclip); cudaFree(ss->b_vproj);
        ss->b_vclip = nullptr; ss->b_vproj = nullptr; ss->b_vert_cap = 0;
        CUDA_TRY(cudaMalloc(&ss->b_vclip, per_launch * tv * sizeof(float4)));
        CUDA_TRY(cudaMalloc(&ss->b_vproj, per_launch * tv * sizeof(int4)));
        ss->b_vert_cap = per_launch * tv;
    }
    if (per_launch * (2 * nblk + 1) > ss->b_blk_cap) {      // two lists per block
        CUDA_TRY(cudaStreamSynchronize(cs));
        cudaFree(ss->b_cnt); cudaFree(ss->b_off);
        ss->b_cnt = nullptr; ss->b_off = nullptr; ss->b_blk_cap = 0;
        CUDA_TRY(cudaMalloc(&ss->b_cnt, per_launch * (2 * nblk + 1) * sizeof(int)));
        CUDA_TRY(cudaMalloc(&ss->b_off, per_launch * (2 * nblk + 1) * sizeof(int)));
        ss->b_blk_cap = per_launch * (2 * nblk + 1);
    }
    CUDA_TRY(grow((void **)&ss->b_pairs, &ss->b_pair_cap, per_launch * pairs_cap, sizeof(unsigned)));
    CUDA_TRY(grow((void **)&ss->b_pbox, &ss->b_pbox_cap, per_launch * cap, sizeof(unsigned)));

    StagedDev g;
    g.vis = ss->g_vis; g.bcount = nullptr; g.bidx = nullptr;
    g.recs = ss->g_recs; g.srecs = smooth ? ss->g_srecs : nullptr; g.bbox = ss->g_bbox; g.count = ss->g_count; g.cap = (int)cap;
    BinnedDev bd;
    bd.vclip = ss->b_vclip; bd.vproj = ss->b_vproj; bd.blk_cnt = ss->b_cnt; bd.blk_off = ss->b_off; bd.pairs = ss->b_pairs; bd.pbox = ss->b_pbox;
    bd.pairs_cap = (int)pairs_cap; bd.total_verts = f.total_verts;
    const int first = f.scene_begin, last = f.scene_begin + f.scene_count;
    for (int s0 = first; s0 < last; s0 += (int)per_launch) {
        const int n = (int)((size_t)(last - s0) < per_launch ? (size_t)(last - s0) : per_launch);
        g.scene0 = s0;
        CUDA_TRY(cudaMemsetAsync(ss->g_count, 0, (size_t)n * sizeof(int), cs));
        CUDA_TRY(cudaMemsetAsync(ss->b_cnt, 0, (size_t)n * 2 * nblk * sizeof(int), cs));
        cull_kernel<<<dim3((unsigned)((f.total_inst + 255) / 256), (unsigned)n), 256, 0, cs>>>(f, g);
        COUNT_LAUNCH();
        bin_xform_kernel<<<dim3((unsigned)((tv + B_THREADS - 1) / B_THREADS), (unsigned)n), B_THREADS, 0, cs>>>(f, g, bd);
        COUNT_LAUNCH();
        bin_tri_kernel<<<dim3((unsigned)((f.total_slots + B_THREADS - 1) / B_THREADS), (unsigned)n), B_THREADS, 0, cs>>>(f, g, bd);
        COUNT_LAUNCH();
        // a quarter of the lists' capacity (grid-stride loop inside): they are rarely more than a third full
        const dim3 bgrid((unsigned)((cap / 4 + B_THREADS - 1) / B_THREADS), (unsigned)n);
        bin_blocks_kernel<false><<<bgrid, B_THREADS, 0, cs>>>(f, g, bd);
        COUNT_LAUNCH();
        bin_scan_kernel<<<(unsigned)n, B_THREADS, 0, cs>>>(f, bd);
        COUNT_LAUNCH();
        bin_blocks_kernel<true><<<bgrid, B_THREADS, 0, cs>>>(f, g, bd);
        COUNT_LAUNCH();
        const dim3 rgrid((unsigned)((nblk + B_WPB - 1) / B_WPB), (unsigned)n);
        if (smooth)
            raster_binned_kernel<true><<<rgrid, B_WPB * 32, 0, cs>>>(f, g, bd);
        else
            raster_binned_kernel<false><<<rgrid, B_WPB * 32, 0, cs>>>(f, g, bd);
        COUNT_LAUNCH();
        CUDA_TRY(cudaGetLastError());
    }
    return PBR_OK;
}

int pbr_render(const pbr_frame_desc *d, void *stream) {
    if (int rc = check_frame(d, "pbr_render", true)) return rc;
    if (d->scene_count == 0) return PBR_OK;
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));

    FrameDev f;
    fill_uniforms(d, f);
    std::lock_guard<std::mutex> lock(g_mu[device & 63]);
    DeviceState *st = nullptr;
    if (int rc = device_state(device, &st)) return rc;
    f.status = st->status;
    f.status_host = st->status_host_dev;
    StreamScratch *ss = &st->per_stream[stream];

    const int W = f.W, H = f.H;
    const int nbx = (W + 7) / 8;
    const int H8 = ((H + 7) / 8) * 8;
    const size_t warp_smem = warp_smem_bytes(nbx * (H8 / 8), W_WARPS);
    auto warp_eligible = [&](const NodeStats &ns) {
        // once a scene overflowed the small-scene kernel's record slots (too many clipped fan
        // triangles; sticky flag in host-mapped memory) this device keeps to the general kernel
        return ns.warp_ok && !ns.any_smooth && st->status_host[0] == 0 && !(d->flags & PBR_FRAME_FORCE_GENERAL) && ns.slots <= W_MAXSLOT && ns.verts <= W_MAXVERT && ns.insts <= W_MAXINST &&
               nbx <= 256 && H8 / 8 <= 256 && warp_smem <= 100 * 1024 && warp_smem <= (size_t)st->max_smem_optin;
    };

    // a usable static layer?  (same device and layout; only the small-scene kernel consumes it)
    const pbr_base_s *base = d->base;
    if (base && (base->device != device || base->W != W || base->H != H || base->C != f.C))
        return fail(PBR_EINVAL, "pbr_render: base layer was rendered for %dx%dx%d on device %d", base->W, base->H, base->C, base->device);

    NodeStats ns;
    bool use_base = false;
    if (base) {
        if (int rc = fill_nodes(d, device, NODES_SKIP_BASE, f, ns)) return rc;
        use_base = ns.skipped > 0 && warp_eligible(ns);
    }
    if (!use_base)
        if (int rc = fill_nodes(d, device, NODES_ALL, f, ns)) return rc;

    if (warp_eligible(ns)) {
        f.BH = H8; f.nbands = 1; f.nbx = nbx; f.nby = H8 / 8;
        f.nbx_magic = div_magic((unsigned)nbx);
        f.w_region = (int)warp_scene_bytes(nbx * (H8 / 8));
        f.w_inst_magic = div_magic((unsigned)f.total_inst);
        f.w_vert_magic = div_magic((unsigned)f.total_verts);
        f.w_slot_magic = div_magic((unsigned)f.total_slots);
        // w_qctr_off is set where the kernel variant (scenes per CTA) is chosen
        f.plane_stride = H * W;
        f.linear = 1;
        f.debug = (int)((d->flags >> 8) & 3u);
        // posed nodes: computed in the kernel when they fit f.poses, else materialised first
        if (ns.n_poses > 0 && !ns.poses_in_frame)
            if (int rc = materialise_poses(d, stream)) return rc;
        f.write_mats = (ns.poses_in_frame && (d->flags & PBR_FRAME_WRITE_MATS)) ? 1 : 0;
        // 32-bit depth keys are enough when ties against the static layer always go to the layer, i.e. when
        // everything in it was drawn before the first node of this frame (CartPole: the rail is node 0)
        static const bool no_k32 = getenv("PBR_B200_NO_KEYS32") != nullptr;       // A/B timing aid
        f.keys32 = (!no_k32 && (!use_base || ns.skipped_id_end <= ns.active_id_begin)) ? 1 : 0;
        {   // per-scene inputs worth an L2 prefetch by an earlier CTA: the pose channels (rows of the caller's state
            // tensor -- channels that point into the same rows count once), the VP rows, the instance colours
            f.n_pf = 0;
            auto add = [&](const void *ptr, long long row_bytes) {
                if (ptr == nullptr || row_bytes <= 0 || row_bytes > 64) return;
                const unsigned char *q = static_cast<const unsigned char *>(ptr);
                for (int k = 0; k < f.n_pf; ++k)
                    if (f.pf_row[k] == (int)row_bytes && q - f.pf_ptr[k] < row_bytes && f.pf_ptr[k] - q < row_bytes) {
                        if (q < f.pf_ptr[k]) f.pf_ptr[k] = q;
                        return;
                    }
                if (f.n_pf < MAX_FRAME_PF) { f.pf_ptr[f.n_pf] = q; f.pf_row[f.n_pf] = (int)row_bytes; ++f.n_pf; }
            };
            for (int i = 0; i < f.n_nodes; ++i) {
                const NodeDev &nd = f.nodes[i];
                if (nd.shared || nd.pose_idx < 0) continue;
                const pbr_channel *ch = f.poses[nd.pose_idx].pos;       // pos[3], hpr[3], scale: seven in a row
                for (int k = 0; k < 7; ++k) add(ch[k].ptr, (long long)nd.inst * ch[k].stride * 4);
            }
            if (f.vp_scene_override < 0) add(f.vp, 64);
            for (int i = 0; i < f.n_nodes; ++i)
                if (!f.nodes[i].shared) add(f.nodes[i].cols, (long long)f.nodes[i].inst * 16);
        }
        {   // programmatic launch chain: this launch may start while the previous small-scene launches on the
            // stream drain; it orders itself behind them when it writes memory they write (see raster_warp.cuh)
            const unsigned char *lo = f.out + (size_t)f.scene_begin * f.C * H * W;
            const unsigned char *hi = lo + (size_t)f.scene_count * f.C * H * W;
            bool overlap = false;
            for (int k = 0; k < 2; ++k)
                overlap |= ss->last_out_lo[k] != nullptr && lo < ss->last_out_hi[k] && ss->last_out_lo[k] < hi;
            f.sync_early = (overlap || f.write_mats) ? 1 : 0;
            ss->last_out_lo[1] = ss->last_out_lo[0]; ss->last_out_hi[1] = ss->last_out_hi[0];
            ss->last_out_lo[0] = lo; ss->last_out_hi[0] = hi;
        }
        if (use_base) {
            f.base_color = base->color; f.base_keys = base->keys; f.base_flags = base->flags;
        } else if ((((size_t)f.C * H * W) & 15) == 0) {
            const size_t need = (size_t)f.C * H * W;
            if (need > ss->bgtile_bytes) {
                CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
                cudaFree(ss->bgtile);
                ss->bgtile = nullptr; ss->bgtile_bytes = 0; ss->bg_sig[0] = 0;
                CUDA_TRY(cudaMalloc(&ss->bgtile, need));
                ss->bgtile_bytes = need;
            }
            if (ss->bg_sig[0] != (unsigned)W || ss->bg_sig[1] != (unsigned)H || ss->bg_sig[2] != (unsigned)f.C || ss->bg_sig[3] != f.bg) {
                fill_planes_kernel<<<(unsigned)((need + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ss->bgtile, H * W, f.C, f.bg);
                COUNT_LAUNCH();
                CUDA_TRY(cudaGetLastError());
                ss->bg_sig[0] = (unsigned)W; ss->bg_sig[1] = (unsigned)H; ss->bg_sig[2] = (unsigned)f.C; ss->bg_sig[3] = f.bg;
            }
            f.base_color = ss->bgtile;
        }
        {   // record overflow pool (claimed per SM inside the kernel, see raster_warp.cuh)
            const size_t entries = (size_t)st->sm_count * W_POOL_PER_SM;
            if (!st->ovf_recs) {
                CUDA_TRY(cudaMalloc(&st->ovf_recs, entries * W_OVF_MAXREC * sizeof(Rec)));
                CUDA_TRY(cudaMalloc(&st->ovf_busy, 3 * (size_t)st->sm_count * sizeof(unsigned)));     // busy bits, tickets, done counts
                CUDA_TRY(cudaMemset(st->ovf_busy, 0, 3 * (size_t)st->sm_count * sizeof(unsigned)));
            }
            const size_t words = (size_t)f.nbx * f.nby * W_OVF_MW;
            if (words > st->ovf_mask_words) {
                CUDA_TRY(cudaDeviceSynchronize());             // shared by every stream of the device
                cudaFree(st->ovf_masks);
                st->ovf_masks = nullptr; st->ovf_mask_words = 0;
                CUDA_TRY(cudaMalloc(&st->ovf_masks, entries * words * sizeof(unsigned)));
                st->ovf_mask_words = words;
            }
            f.ovf_recs = st->ovf_recs; f.ovf_masks = st->ovf_masks; f.ovf_busy = st->ovf_busy;
            f.bg_ticket = st->ovf_busy + st->sm_count; f.bg_done = st->ovf_busy + 2 * st->sm_count;
        }
        if (!st->attr_warp) {
            CUDA_TRY(cudaFuncSetAttribute(raster_warp_kernel<W_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(raster_warp_kernel<W_WARPS, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CUDA_TRY(cudaFuncSetAttribute(raster_warp_kernel<W_WARPS_TMA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem_optin));
            CUDA_TRY(cudaFuncSetAttribute(raster_warp_kernel<W_WARPS_TMA, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CUDA_TRY(cudaFuncSetAttribute(raster_warp_kernel<W_WARPS_TMA_SMALL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem_optin));
            CUDA_TRY(cudaFuncSetAttribute(raster_warp_kernel<W_WARPS_TMA_SMALL, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            st->attr_warp = true;
        }
        // background by TMA (scenes + 2 helper warps per CTA + the image in shared memory): 14 scenes per CTA while two
        // such CTAs fit an SM, else 6 scenes per CTA (three CTAs per SM at 84x84: a single CTA of 14 leaves the SM
        // idle at every barrier between its shared geometry phases -- measured 143 vs 113 us per 16,384 scenes)
        static const int tma_mode = getenv("PBR_B200_WARP_TMA") ? atoi(getenv("PBR_B200_WARP_TMA")) : -1;   // 0 / 1 force, else auto
        const size_t tile_bytes = (size_t)f.C * H * W;
        const size_t two_per_sm = ((size_t)st->max_smem_optin + 1024 - 2 * 1024) / 2;
        const size_t tma_big = warp_smem_bytes(nbx * (H8 / 8), W_WARPS_TMA, tile_bytes);
        const size_t tma_small = warp_smem_bytes(nbx * (H8 / 8), W_WARPS_TMA_SMALL, tile_bytes);
        const bool big = tma_big <= two_per_sm;
        const size_t tma_smem = big ? tma_big : tma_small;
        const bool tma_ok = f.base_color != nullptr && (tile_bytes & 15) == 0 && tma_smem <= (size_t)st->max_smem_optin;
        // measured: faster for 64x64 and 84x84, slower for 32x32 (the copy is cheap, the wide barrier is not) and
        // 128x128 (the image crowds out the scenes)
        const bool use_tma = tma_ok && (tma_mode == 1 || (tma_mode != 0 && tma_smem <= two_per_sm && tile_bytes >= 8192 && tile_bytes <= 32768));
        if (use_tma && big) {
            const unsigned wgrid = (unsigned)((f.scene_count + W_WARPS_TMA - 1) / W_WARPS_TMA);
            f.w_qctr_off = (int)warp_qctr_offset(nbx * (H8 / 8), W_WARPS_TMA);
            CUDA_TRY(launch_dependent(raster_warp_kernel<W_WARPS_TMA, true>, wgrid, 32 * (W_WARPS_TMA + w_helpers(W_WARPS_TMA)), tma_smem, stream, f));
            COUNT_LAUNCH();
        } else if (use_tma) {
            const unsigned wgrid = (unsigned)((f.scene_count + W_WARPS_TMA_SMALL - 1) / W_WARPS_TMA_SMALL);
            f.w_qctr_off = (int)warp_qctr_offset(nbx * (H8 / 8), W_WARPS_TMA_SMALL);
            CUDA_TRY(launch_dependent(raster_warp_kernel<W_WARPS_TMA_SMALL, true>, wgrid, 32 * (W_WARPS_TMA_SMALL + w_helpers(W_WARPS_TMA_SMALL)), tma_smem, stream, f));
            COUNT_LAUNCH();
        } else {
            static const size_t smem_pad = getenv("PBR_B200_WARP_SMEM_PAD") ? (size_t)atoi(getenv("PBR_B200_WARP_SMEM_PAD")) : 0;   // occupancy experiments
            const unsigned wgrid = (unsigned)((f.scene_count + W_WARPS - 1) / W_WARPS);
            f.w_qctr_off = (int)warp_qctr_offset(nbx * (H8 / 8), W_WARPS);
            CUDA_TRY(launch_dependent(raster_warp_kernel<W_WARPS, false>, wgrid, 32 * W_WARPS, warp_smem + smem_pad, stream, f));
            COUNT_LAUNCH();
        }
        CUDA_TRY(cudaGetLastError());
        return PBR_OK;
    }
    // large scenes: geometry once per frame into per-scene record lists, then TMA-staged raster
    // (ordinary launches: they start after everything before them on the stream has completed)
    ss->last_out_lo[0] = ss->last_out_lo[1] = nullptr;
    if (ns.n_poses > 0)
        if (int rc = materialise_poses(d, stream)) return rc;
    size_t smem_fused = 0;
    if (int rc = plan_general(f, st, &smem_fused)) return rc;
    if (!(d->flags & PBR_FRAME_FORCE_FUSED) && (ns.slots > CH || f.nbands > 1)) {
        // Tiles of several bands: per-block record lists (raster_binned.cuh) -- measured 1.3x (config 3) and 2.0x
        // (config 5) faster than the band-based path, which rescans and re-bins the scene's records in every band.
        // Single-band tiles (Steering-v0 at 64x64) stay on the band-based path: its one CTA per scene needs three
        // launches per frame instead of seven.  After a list overflow: band-based path with worst-case capacity.
        // PBR_B200_LARGE=staged|binned and PBR_FRAME_FORCE_STAGED / _BINNED override the choice (tests, A/B).
        static const char *want = getenv("PBR_B200_LARGE");
        const bool fits = ns.verts <= 0x7fffffffll / 64 && d->tile_w <= 2048 && d->tile_h <= 2048 && !st->cap_worst_case;
        bool binned = f.nbands > 1 && ns.slots <= 20000;       // (Steering-v0, ~50k slots per scene, mostly culled:
                                                               //  1.03 ms band-based vs 1.20 ms per 1024 scenes)
        if (want != nullptr) binned = strcmp(want, "binned") == 0;
        if (d->flags & PBR_FRAME_FORCE_STAGED) binned = false;
        if (d->flags & PBR_FRAME_FORCE_BINNED) binned = true;
        if (binned && fits) return launch_binned(f, st, stream);
        return launch_staged(f, st, stream);
    }
    return launch_general(f, st, stream);
}

#ifdef PBR_W_TIMING
// timing build only: copy the head of the overflow pool (where the kernel dumps its time stamps) to the host
int pbr_debug_pool(void *dst, size_t bytes) {
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lock(g_mu[device & 63]);
    DeviceState *st = nullptr;
    if (int rc = device_state(device, &st)) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(dst, st->ovf_recs, bytes, cudaMemcpyDeviceToHost));
    return PBR_OK;
}
#endif

int pbr_base_create(int32_t device, pbr_base_t *out) {
    if (!out) return fail(PBR_EINVAL, "pbr_base_create: out is NULL");
    pbr_base_s *b = new (std::nothrow) pbr_base_s();
    if (!b) return fail(PBR_ENOMEM, "pbr_base_create: host allocation failed");
    memset(b, 0, sizeof(*b));
    b->device = device;
    *out = b;
    return PBR_OK;
}

int pbr_base_destroy(pbr_base_t b) {
    if (!b) return PBR_OK;
    cudaFree(b->color); cudaFree(b->keys); cudaFree(b->flags);
    delete b;
    return PBR_OK;
}

int pbr_base_render(pbr_base_t b, const pbr_frame_desc *d, void *stream) {
    if (!b) return fail(PBR_EINVAL, "pbr_base_render: NULL base");
    if (int rc = check_frame(d, "pbr_base_render", false)) return rc;
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    if (device != b->device) return fail(PBR_EINVAL, "pbr_base_render: base belongs to device %d, current device is %d", b->device, device);
    if (d->scene_begin >= d->num_scenes) return fail(PBR_EINVAL, "pbr_base_render: scene_begin outside vp");

    FrameDev f;
    fill_uniforms(d, f);
    std::lock_guard<std::mutex> lock(g_mu[device & 63]);
    DeviceState *st = nullptr;
    if (int rc = device_state(device, &st)) return rc;
    f.status = st->status;
    f.status_host = st->status_host_dev;
    NodeStats ns;
    if (int rc = fill_nodes(d, device, NODES_SHARED_ONLY, f, ns)) return rc;
    {
        StreamScratch *ss = &st->per_stream[stream];
        ss->last_out_lo[0] = ss->last_out_lo[1] = nullptr;
    }
    if (ns.n_poses > 0)
        if (int rc = materialise_poses(d, stream)) return rc;

    // (re)allocate the layer: colour image + keys for every 8x8 block of every band
    size_t smem_unused = 0;
    if (int rc = plan_general(f, st, &smem_unused)) return rc;
    const size_t color_bytes = align16((size_t)f.C * f.H * f.W);
    const size_t key_blocks = (size_t)f.nbx * f.nby * f.nbands;           // the last band may overhang the tile
    if (color_bytes > b->color_bytes || key_blocks > b->key_blocks) {
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        cudaFree(b->color); cudaFree(b->keys); cudaFree(b->flags);
        b->color = nullptr; b->keys = nullptr; b->flags = nullptr; b->color_bytes = b->key_blocks = 0; b->W = 0;
        CUDA_TRY(cudaMalloc(&b->color, color_bytes));
        CUDA_TRY(cudaMalloc(&b->keys, key_blocks * 64 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMalloc(&b->flags, key_blocks));
        b->color_bytes = color_bytes; b->key_blocks = key_blocks;
    }
    f.out = b->color;
    f.base_keys_out = b->keys;
    f.base_flags_out = b->flags;
    f.vp_scene_override = d->scene_begin;
    f.scene_begin = 0;
    f.scene_count = 1;
    if (int rc = launch_general(f, st, stream)) { b->W = 0; return rc; }
    b->W = f.W; b->H = f.H; b->C = f.C;
    return PBR_OK;
}

int pbr_device_status(int32_t device, int32_t *status_bits, int32_t clear) {
    if (!status_bits) return fail(PBR_EINVAL, "pbr_device_status: NULL output");
    std::lock_guard<std::mutex> lock(g_mu[device & 63]);
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(device));
    DeviceState *st = nullptr;
    int rc = device_state(device, &st);
    if (rc == PBR_OK) {
        int v = 0;
        cudaError_t e = cudaMemcpy(&v, st->status, sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && clear) { e = cudaMemset(st->status, 0, sizeof(int)); st->status_host[0] = 0; st->status_host[1] = 0; st->cap_worst_case = false; }
        if (e != cudaSuccess) rc = fail(PBR_ECUDA, "pbr_device_status: %s", cudaGetErrorString(e));
        *status_bits = v;
    }
    cudaSetDevice(prev);
    return rc;
}

int pbr_pack_transforms(float *transforms_b44, const float *rot_b33, const float *scale_b, float *out_mats,
                        int32_t n, void *stream) {
    if (n < 0) return fail(PBR_EINVAL, "pbr_pack_transforms: n < 0");
    if (n == 0) return PBR_OK;
    if (!transforms_b44 || !rot_b33 || !scale_b || !out_mats) return fail(PBR_EINVAL, "pbr_pack_transforms: NULL pointer");
    if (reinterpret_cast<size_t>(out_mats) & 15) return fail(PBR_EINVAL, "pbr_pack_transforms: out_mats must be 16-byte aligned");
    pack_transforms_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(transforms_b44, rot_b33, scale_b, out_mats, n);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return PBR_OK;
}

int pbr_compose_transforms(const pbr_pose_desc *poses, int32_t n_poses, void *stream) {
    if (n_poses < 0 || (n_poses > 0 && !poses)) return fail(PBR_EINVAL, "pbr_compose_transforms: bad arguments");
    std::vector<const pbr_pose_desc *> ptrs((size_t)n_poses);
    for (int i = 0; i < n_poses; ++i) ptrs[i] = poses + i;
    return launch_compose(ptrs.data(), n_poses, stream);
}

unsigned long long pbr_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int pbr_device_status_nosync(int32_t device, int32_t *status_bits) {
    if (!status_bits) return fail(PBR_EINVAL, "pbr_device_status_nosync: NULL output");
    if (device < 0 || device >= 64) return fail(PBR_EINVAL, "pbr_device_status_nosync: bad device %d", device);
    std::lock_guard<std::mutex> lock(g_mu[device & 63]);
    const DeviceState &st = g_dev[device];
    *status_bits = st.status_host ? ((st.status_host[0] ? DEVSTAT_WARP_OVERFLOW : 0) | (st.status_host[1] ? DEVSTAT_STAGED_OVERFLOW : 0)) : 0;
    return PBR_OK;
}

}  // extern "C"
