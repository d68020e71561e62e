/*
 * pbr_b200.h -- C ABI of libpbr_b200.so: the B200 (sm_100a) replacement for the pixel path of
 * dolphin-in-a-coma/pybatchrender.
 *
 * The reference has no FFI seam of its own (it is pure Python on top of Panda3D + OpenGL).  Its
 * de-facto seam is the *buffer contract* between the host classes and the two GLSL programs
 * (SURVEY.md section 8b): per node `matbuf` (4 RGBA32F texels per instance = the columns of the
 * model matrix) and `colbuf` (1 texel per instance), per camera `viewbuf` (4 texels per scene = the
 * columns of VP) and the light uniforms.  Every entry point below consumes exactly those buffers,
 * as plain device pointers, and replaces the reference calls cited next to it.
 *
 * Conventions
 *   - plain C, no torch types; every function returns PBR_OK (0) or a negative pbr_status and
 *     never throws; pbr_last_error() returns a thread-local description of the last failure.
 *   - all device work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed
 *     as void*; NULL = the legacy default stream); no hidden synchronisation, no allocation on the
 *     per-frame path.  The caller keeps every buffer alive until the stream work completes.
 *   - matrices are "column packed": 16 floats, element [4*j + i] = row i, column j (what
 *     reference shader_context.py:42-45 `_pack_columns` + `tobytes()` produce).
 */
#ifndef PBR_B200_H
#define PBR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBR_B200_VERSION 101            /* major*10000 + minor*100 + patch */
#define PBR_MAX_NODES 24                /* nodes per frame (kernel parameter space) */
#define PBR_MAX_TILE 2048               /* max tile width / height in pixels */

typedef enum {
    PBR_OK = 0,
    PBR_EINVAL = -1,        /* bad argument (null pointer, misaligned buffer, bad size) */
    PBR_ECUDA = -2,         /* a CUDA runtime call failed; see pbr_last_error() */
    PBR_ENOMEM = -3,
    PBR_EUNSUPPORTED = -4,  /* valid request outside what this build implements */
    PBR_EOVERFLOW = -5      /* an EARLIER frame dropped triangles (record lists too small); nothing was launched by
                               this call, capacity has been raised, call again */
} pbr_status;

#define PBR_MESH_TWO_SIDED 1u           /* do not cull back faces of this mesh */

typedef struct pbr_mesh_s *pbr_mesh_t;  /* device-resident static geometry */
typedef struct pbr_texture_s *pbr_texture_t;   /* device-resident RGBA8 image (p3d_Texture0) */

/* Pose channels: value of instance b = ptr ? ptr[b*stride] : constant (device pointer, stride in floats). */
typedef struct {
    const float *ptr;           /* device pointer or NULL */
    int32_t stride;             /* in floats */
    float constant;
} pbr_channel;

/* One node's pose: position xyz, Euler H/P/R in radians (R = Rz(H) Ry(P) Rx(R), reference
 * shader_context.py:47-84) and uniform scale -> column-packed matrices.  Fuses what the reference
 * does in set_positions + set_hprs + set_scales + upload (node.py:128-154) into one pass.  Used two
 * ways: pbr_compose_transforms writes the matrices of several nodes in one launch; a pose attached
 * to a pbr_node_desc makes pbr_render compute them inside the raster kernel (no matrix buffer is
 * read or written on the small-scene path). */
typedef struct {
    pbr_channel pos[3];
    pbr_channel hpr[3];
    pbr_channel scale;
    float *out_mats;            /* device [n_instances,16] */
    int32_t n_instances;
} pbr_pose_desc;

/* One PBRNode: replaces the `matbuf` / `colbuf` / `instancesPerScene` / `shareAcrossScenes`
 * shader inputs of reference node.py:85-91 and the instanced draw of node.py:68. */
typedef struct {
    pbr_mesh_t mesh;
    const float *mats;          /* device [B,16] column packed  (== matbuf) */
    const float *cols;          /* device [B,4] RGBA            (== colbuf) */
    int32_t instances_per_scene;/* I */
    int32_t shared;             /* 1: B = I (same instances in every scene), 0: B = num_scenes*I,
                                   row = scene*I + inst   (reference basic.vert:25-28) */
    float use_texture;          /* the `useTexture` shader input (reference node.py:285-287, basic.frag:31-32):
                                   base = mix(1, texture(uv).rgb, clamp(use_texture, 0, 1)) */
    uint32_t flags;             /* PBR_NODE_* */
    pbr_texture_t texture;      /* image sampled when use_texture > 0; NULL = white (node stays untextured) */
    const pbr_pose_desc *pose;  /* optional (host pointer, copied during the call).  When set, the node's model
                                   matrices are DEFINED by these channels at the time the frame executes -- the
                                   reference's per-step set_positions / set_hprs (envs/cartpole/renderer.py:125-138)
                                   folded into the frame.  `mats` is then ignored; pose->out_mats (required,
                                   [B,16]) receives the matrices whenever the library has to materialise them
                                   (large-scene paths, static layer, PBR_FRAME_WRITE_MATS) and is left untouched
                                   by the small-scene kernel otherwise; pose->n_instances must equal B. */
} pbr_node_desc;

#define PBR_NODE_IN_BASE 1u             /* already rendered into frame->base: skipped by pbr_render,
                                           but its triangles keep their draw indices */

/* Static layer: the image (colour + depth|id keys) of nodes that are identical in every scene --
 * `shared` nodes seen through a camera whose VP is the same for all scenes (CartPole's rail:
 * reference envs/cartpole/renderer.py:33-39,91-93).  Rendered once by pbr_base_render, after which
 * every scene of pbr_render starts from it instead of the clear colour.  Results are bit-identical
 * to rendering those nodes in every scene. */
typedef struct pbr_base_s *pbr_base_t;

/* One frame: replaces taskMgr.step() + grab_pixels() + _rearrange_img() of reference
 * renderer.py:377-389, i.e. basic.vert + GL rasterisation + basic.frag + readback + flip +
 * un-tiling, for scenes [scene_begin, scene_begin+scene_count). */
typedef struct {
    int32_t num_scenes;         /* K: rows of vp / per-scene node buffers / out */
    int32_t scene_begin;        /* shard / chunk window */
    int32_t scene_count;
    int32_t tile_w, tile_h;     /* per-scene resolution */
    int32_t channels;           /* 3 (RGB) or 4 (RGBA) */
    const float *vp;            /* device [K,16] column packed (== viewbuf, reference camera.py:146-148) */
    float bg[4];                /* clear colour (reference renderer.py:262-264) */
    float ambient[3];           /* reference light.py:11-14 / basic.frag:15-18 */
    float dir_dir[3];
    float dir_col[3];
    float strength;
    int32_t n_nodes;
    const pbr_node_desc *nodes; /* host array, copied during the call */
    uint8_t *out;               /* device [K,C,H,W] contiguous uint8; row 0 = top of the image */
    uint32_t flags;             /* PBR_FRAME_* */
    pbr_base_t base;            /* optional static layer (NULL: start from the clear colour) */
} pbr_frame_desc;

#define PBR_FRAME_FORCE_GENERAL 1u      /* skip the small-scene fast kernel (testing / debugging) */
#define PBR_FRAME_WRITE_MATS 4u          /* posed nodes: also write their matrices to pose->out_mats on the
                                           small-scene path (tests; costs the overlap between frames) */
#define PBR_FRAME_FORCE_STAGED 8u        /* large scenes: band-based path (geometry pre-pass + TMA-staged raster per band) */
#define PBR_FRAME_FORCE_BINNED 16u       /* large scenes: block-list path (per-block record lists, one warp per block) */
#define PBR_FRAME_FORCE_FUSED 2u        /* general path: keep geometry fused into the raster kernel
                                           instead of the geometry pre-pass + TMA-staged raster */

int pbr_version(void);
const char *pbr_last_error(void);

/* Upload static geometry.  pos/nrm: host [n_verts,3] float32 (object space, already baked the way
 * reference node.py:61-72 flattens scale/HPR/pivot into vertices); uv: host [n_verts,2] or NULL; idx: host
 * [n_tris,3].  Replaces loader.loadModel + flattenStrong's vertex data living in GL buffers. */
int pbr_mesh_create(const float *pos_xyz, const float *nrm_xyz, const float *uv, int32_t n_verts,
                    const uint32_t *idx, int32_t n_tris, int32_t device, uint32_t flags, pbr_mesh_t *out);
int pbr_mesh_destroy(pbr_mesh_t mesh);

/* Upload an image for `use_texture` nodes.  rgba: host [height, width, 4] uint8, row 0 = v 0 (the
 * bottom row of the picture, as GL stores it).  Sampling: GL_REPEAT, GL_LINEAR, fp32.  Replaces
 * loader.loadTexture + NodePath.setTexture of reference node.py:277-287. */
int pbr_texture_create(const uint8_t *rgba, int32_t width, int32_t height, int32_t device, pbr_texture_t *out);
int pbr_texture_destroy(pbr_texture_t texture);
int pbr_mesh_info(pbr_mesh_t mesh, int32_t *n_tris, int32_t *all_flat, int32_t *device);

/* Render one frame (asynchronous on `stream`). */
int pbr_render(const pbr_frame_desc *frame, void *stream);

/* Static layer life cycle.  pbr_base_render draws the `shared` nodes of `frame` (all other nodes
 * are skipped but keep their draw indices) for ONE scene using row `frame->scene_begin` of vp;
 * frame->out is ignored.  The caller re-renders the base whenever those nodes, the camera, the
 * light, the clear colour or the tile size change. */
int pbr_base_create(int32_t device, pbr_base_t *out);
int pbr_base_destroy(pbr_base_t base);
int pbr_base_render(pbr_base_t base, const pbr_frame_desc *frame, void *stream);

/* Instance-transform kernel: out_mats[b] = column-packed [ R_b * s_b | t_b ; 0 0 0 1 ] and the
 * 3x3 block of transforms_b44 is refreshed in place -- the device version of reference
 * node.py:116-126 `_upload_current_transforms` (transforms[:, :3, :3] = rot * scale; pack; upload).
 *   transforms: device [B,4,4] row-major (translation read from [:, :3, 3])
 *   rot: device [B,3,3]; scale: device [B] ; out_mats: device [B,16]. */
int pbr_pack_transforms(float *transforms_b44, const float *rot_b33, const float *scale_b,
                        float *out_mats, int32_t n_instances, void *stream);

/* Pose kernel: out_mats of every pose <- matrices of its channels (one launch per 8 poses).  Same
 * arithmetic, bit for bit, as a pose attached to a node of pbr_render. */
int pbr_compose_transforms(const pbr_pose_desc *poses, int32_t n_poses, void *stream);

/* Sticky per-device status bits written by the kernels (diagnostics; synchronises the device).
 * bit 0: the small-scene kernel ran out of shared-memory record slots for clipped triangles in some
 *        scene (the frame is still exact: the surplus went to the overflow pool; later frames on this
 *        device take the general kernel until the bit is cleared);
 * bit 1: the geometry pre-pass of a large scene ran out of per-scene record capacity and DROPPED
 *        triangles -- the frame is wrong; the next large-scene pbr_render returns PBR_EOVERFLOW once and later
 *        frames use worst-case capacities (clear = 1 also returns the device to the normal capacities). */
int pbr_device_status(int32_t device, int32_t *status_bits, int32_t clear);

/* The same bits without a device synchronisation: read from host-mapped memory the kernels also write
 * (may lag behind frames still in flight). */
int pbr_device_status_nosync(int32_t device, int32_t *status_bits);

/* Number of kernels this library has enqueued (or captured into a CUDA graph) in this process so far:
 * lets a benchmark count the launches of its timed region instead of assuming them. */
unsigned long long pbr_kernel_launches(void);

#ifdef PBR_W_TIMING
/* timing builds only (profiles/kernel_timestamps.py): copies the time stamps the small-scene kernel dumped */
int pbr_debug_pool(void *dst, size_t bytes);
#endif

#ifdef __cplusplus
}
#endif
#endif /* PBR_B200_H */
