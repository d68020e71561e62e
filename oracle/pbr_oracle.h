/*
 * pbr_oracle.h -- CPU oracle for the PyBatchRender pixel path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of what the reference computes between "buffer
 * textures uploaded" and "uint8 [N,C,H,W] tensor returned":
 *   pybatchrender/shaders/basic.vert:24-56   (instance decode, clip = VP*(M*v), tile remap, normal, colour)
 *   pybatchrender/shaders/basic.frag:20-38   (tile scissor, ambient + Lambert, output colour)
 *   the OpenGL fixed-function rules the driver applies in between:
 *   near-plane clip, perspective divide, viewport, pixel-centre sampling, depth LESS with clear 1.0,
 *   back-face cull, RGBA8 unorm conversion
 *   pybatchrender/renderer/frame_grabber.py:105 (.flip(0)) and renderer.py:352-363 (_rearrange_img):
 *   folded in by rasterising each scene in image orientation (row 0 = top) straight into out[scene].
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pybatchrender_b200) never does.
 *
 * Parity status: PINNED against the reference's own golden output -- the 16 CartPole tiles embedded
 * in examples/notebooks/cartpole_benchmark.ipynb (raw line 473) and the three printed states
 * (raw lines 422-424); see tests/test_oracle_golden.py and tests/golden/make_golden.py.
 * Unpinned corners (no reference output exists in the tree): tie-break orientation of the
 * top-left rule, exact depth ties, smooth-normal meshes, near-plane clipping, C=4,
 * texture filtering (GL leaves the filter arithmetic to the implementation; here: GL_REPEAT, GL_LINEAR, fp32).
 */
#ifndef PBR_ORACLE_H
#define PBR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MESH_TWO_SIDED 1u   /* disable back-face culling for this mesh */

typedef struct {
    const float *pos;      /* [n_verts,3] object-space positions (already baked: node.py:61-72) */
    const float *nrm;      /* [n_verts,3] object-space normals */
    const uint32_t *idx;   /* [n_tris,3] */
    int n_verts;
    int n_tris;
    uint32_t flags;
    const float *uv;       /* [n_verts,2] texture coordinates, or NULL (basic.vert v_uv) */
} orc_mesh;

typedef struct {
    orc_mesh mesh;
    const float *mats;     /* [B,16]: texel j of instance b = column j of M_b  (== matbuf, node.py:110-126) */
    const float *cols;     /* [B,4] RGBA                                       (== colbuf, node.py:156-161) */
    int instances_per_scene;
    int shared;            /* shareAcrossScenes: B = I when set, else B = K*I  (node.py:42-46) */
    float use_texture;     /* useTexture shader input (node.py:285-287) */
    const uint8_t *tex;    /* [tex_h, tex_w, 4] RGBA8, row 0 = v 0; NULL = untextured */
    int tex_w, tex_h;
} orc_node;

typedef struct {
    int num_scenes;        /* K: rows in vp / per-scene node buffers / out */
    int scene_begin;       /* render scenes [scene_begin, scene_begin+scene_count) */
    int scene_count;
    int tile_w, tile_h, channels;   /* channels 3 or 4 */
    const float *vp;       /* [K,16]: texel j of scene k = column j of VP_k (== viewbuf, camera.py:146-148) */
    float bg[4];           /* clear colour (renderer.py:262-264) */
    float ambient[3];      /* light.py:11-14 */
    float dir_dir[3];
    float dir_col[3];
    float strength;
    int n_nodes;
    const orc_node *nodes;
    uint8_t *out;          /* [K,C,H,W] contiguous uint8 */
    int n_threads;         /* >=1: scenes are split across this many pthreads */
} orc_frame;

/* Returns 0 on success, -1 on invalid arguments. */
int orc_render(const orc_frame *f);

/* Per-scene debug planes (single scene, single thread): depth [H*W] float (1.0 = clear) and
 * prim id [H*W] (0 = background, else 1 + draw index of the winning triangle). */
int orc_render_scene_debug(const orc_frame *f, int scene, uint8_t *out_chw, float *depth, uint32_t *prim);

int orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
