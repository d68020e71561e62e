/*
 * pbr_oracle.c -- CPU oracle for the PyBatchRender pixel path.  TEST INFRASTRUCTURE ONLY
 * (see pbr_oracle.h for who may load it and for the parity status: PINNED on the notebook goldens).
 *
 * Build: make -C oracle   (gcc -O3 -ffp-contract=off; every fused multiply-add below is an explicit
 * fmaf(), so the float32 results are reproducible and are what the CUDA path is compared against
 * bit for bit).
 *
 * Structure is deliberately naive -- for each scene, for each node, instance and triangle in draw
 * order: transform, clip, project, snap to 1/256 px, exact int64 edge functions over the clamped
 * bounding box, depth LESS, shade.  No binning, no hierarchy, no fast paths.
 *
 * Reference lines restated by each function are cited inline as  [ref: file:line].
 */
#include "pbr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_VERSION 1
#define SUBPIX 256          /* 8 sub-pixel bits (GL_SUBPIXEL_BITS of the reference's GPU class) */
#define GUARD 1024.0f       /* guard band, in tile NDC units; coordinates beyond are clipped */
#define MAX_POLY 10

int orc_version(void) { return ORC_VERSION; }

typedef struct {
    float c[4];   /* clip-space x y z w */
    float n[3];   /* world-space unit normal (v_normal of basic.vert:53) */
    float uv[2];  /* v_uv */
} cvert;

/* what a textured draw samples */
typedef struct {
    const uint8_t *tex;   /* NULL = untextured */
    int w, h;
    float use;            /* clamp(useTexture, 0, 1) */
} texture_t;

/* texture(p3d_Texture0, uv).rgb with GL_REPEAT / GL_LINEAR, fp32, texel centres at (i + 0.5) / size */
static void sample_bilinear(const texture_t *T, float u, float v, float rgb[3]) {
    u -= floorf(u);
    v -= floorf(v);
    const float x = fmaf(u, (float)T->w, -0.5f), y = fmaf(v, (float)T->h, -0.5f);
    const float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    int x0 = (int)xf, y0 = (int)yf;
    x0 = x0 < 0 ? x0 + T->w : (x0 >= T->w ? x0 - T->w : x0);
    y0 = y0 < 0 ? y0 + T->h : (y0 >= T->h ? y0 - T->h : y0);
    const int x1 = x0 + 1 >= T->w ? 0 : x0 + 1, y1 = y0 + 1 >= T->h ? 0 : y0 + 1;
    const uint8_t *c00 = T->tex + 4 * ((size_t)y0 * T->w + x0), *c10 = T->tex + 4 * ((size_t)y0 * T->w + x1);
    const uint8_t *c01 = T->tex + 4 * ((size_t)y1 * T->w + x0), *c11 = T->tex + 4 * ((size_t)y1 * T->w + x1);
    for (int c = 0; c < 3; ++c) {
        const float a00 = (float)c00[c] / 255.0f, a10 = (float)c10[c] / 255.0f;
        const float a01 = (float)c01[c] / 255.0f, a11 = (float)c11[c] / 255.0f;
        const float lo = fmaf(fx, a10 - a00, a00);
        const float hi = fmaf(fx, a11 - a01, a01);
        rgb[c] = fmaf(fy, hi - lo, lo);
    }
}

/* r = Mcols * v, Mcols = 16 floats, texel j = column j  [ref: basic.vert:30-43, GLSL mat4*vec4] */
static inline void mat_vec4(const float *m, const float v[4], float r[4]) {
    for (int i = 0; i < 4; ++i)
        r[i] = fmaf(m[i], v[0], fmaf(m[4 + i], v[1], fmaf(m[8 + i], v[2], m[12 + i] * v[3])));
}

/* r = normalize(mat3(M) * n)  [ref: basic.vert:53] */
static inline void xform_normal(const float *m, const float n[3], float r[3]) {
    float t[3];
    for (int i = 0; i < 3; ++i) t[i] = fmaf(m[i], n[0], fmaf(m[4 + i], n[1], m[8 + i] * n[2]));
    float l2 = fmaf(t[2], t[2], fmaf(t[1], t[1], t[0] * t[0]));
    float inv = 1.0f / sqrtf(l2);
    r[0] = t[0] * inv; r[1] = t[1] * inv; r[2] = t[2] * inv;
}

static inline uint8_t unorm8(float c) {
    /* RGBA8 render target write: clamp, scale, round half up */
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint8_t)(int)(c * 255.0f + 0.5f);
}

typedef struct {
    float amb[3], dcol[3], ldir[3];   /* ldir = normalize(dirLightDir) */
    float s;                          /* clamp(lightingStrength, 0, 1) */
} light_t;

/* [ref: basic.frag:31-38] with useTexture = 0 (base = 1) and n already unit length */
static inline void shade(const light_t *L, const float n[3], const float col[4], uint8_t rgba[4]) {
    float ndl = fmaxf(fmaf(n[2], L->ldir[2], fmaf(n[1], L->ldir[1], n[0] * L->ldir[0])), 0.0f);
    for (int c = 0; c < 3; ++c) {
        float light = fmaf(ndl, L->dcol[c], L->amb[c]);
        float l = fmaf(light, L->s, 1.0f - L->s);          /* mix(1, light, s) */
        rgba[c] = unorm8(col[c] * l);
    }
    rgba[3] = unorm8(col[3]);
}

typedef struct {
    int W, H, C;
    uint8_t *out;      /* [C][H][W] of this scene */
    float *depth;      /* [H*W] */
    uint32_t *prim;    /* [H*W] */
} target_t;

/* distance to clip plane p  (>= 0 is inside)  [GL clip volume; x/y use the guard band] */
static inline float plane_dist(const cvert *v, int p) {
    switch (p) {
    case 0: return v->c[2] + v->c[3];                 /* near: z >= -w */
    case 1: return GUARD * v->c[3] + v->c[0];
    case 2: return GUARD * v->c[3] - v->c[0];
    case 3: return GUARD * v->c[3] + v->c[1];
    default: return GUARD * v->c[3] - v->c[1];
    }
}

static int clip_poly(const cvert *in3, cvert *poly) {
    cvert a[MAX_POLY], b[MAX_POLY];
    int n = 3;
    memcpy(a, in3, 3 * sizeof(cvert));
    for (int p = 0; p < 5 && n >= 3; ++p) {
        int any_out = 0;
        for (int i = 0; i < n; ++i) if (plane_dist(&a[i], p) < 0.0f) any_out = 1;
        if (!any_out) continue;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const cvert *u = &a[i], *v = &a[(i + 1) % n];
            float du = plane_dist(u, p), dv = plane_dist(v, p);
            int iu = !(du < 0.0f), iv = !(dv < 0.0f);
            if (iu) b[m++] = *u;
            if (iu != iv) {
                /* always interpolate from the inside vertex towards the outside one */
                const cvert *vi = iu ? u : v, *vo = iu ? v : u;
                float di = iu ? du : dv, dout = iu ? dv : du;
                float t = di / (di - dout);
                cvert w;
                for (int k = 0; k < 4; ++k) w.c[k] = fmaf(t, vo->c[k] - vi->c[k], vi->c[k]);
                for (int k = 0; k < 3; ++k) w.n[k] = fmaf(t, vo->n[k] - vi->n[k], vi->n[k]);
                for (int k = 0; k < 2; ++k) w.uv[k] = fmaf(t, vo->uv[k] - vi->uv[k], vi->uv[k]);
                b[m++] = w;
            }
        }
        n = m;
        memcpy(a, b, (size_t)n * sizeof(cvert));
    }
    if (n < 3) return 0;
    memcpy(poly, a, (size_t)n * sizeof(cvert));
    return n;
}

/* Rasterise one (already clipped) triangle. */
static void raster_tri(const target_t *T, const light_t *L, const cvert v_in[3], const float col[4],
                       int flat, int two_sided, uint32_t id, const texture_t *tx) {
    if (tx->tex) flat = 0;                 /* textured triangles are shaded per pixel */
    cvert v[3];
    memcpy(v, v_in, sizeof(v));
    const float hw = 0.5f * (float)T->W, hh = 0.5f * (float)T->H;
    int32_t X[3], Y[3];
    float z[3], rw[3];
    for (int i = 0; i < 3; ++i) {
        if (!(v[i].c[3] > 0.0f)) return;
        rw[i] = 1.0f / v[i].c[3];
        float xs = fmaf(v[i].c[0] * rw[i], hw, hw);            /* viewport, tile-local, x right */
        float ys = fmaf(-(v[i].c[1] * rw[i]), hh, hh);         /* image orientation: y down [ref: frame_grabber.py:105 flip] */
        z[i] = fmaf(0.5f, v[i].c[2] * rw[i], 0.5f);            /* depth range [0,1] */
        float fx = xs * (float)SUBPIX, fy = ys * (float)SUBPIX;
        if (!(fabsf(fx) < 1073741824.0f) || !(fabsf(fy) < 1073741824.0f)) return;
        X[i] = (int32_t)rintf(fx);
        Y[i] = (int32_t)rintf(fy);
    }
    int64_t area2 = (int64_t)(X[1] - X[0]) * (Y[2] - Y[0]) - (int64_t)(X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 == 0) return;
    if (area2 > 0) {                       /* clockwise in GL's y-up window space: back face */
        if (!two_sided) return;
        int32_t t; float f; cvert cv;
        t = X[1]; X[1] = X[2]; X[2] = t;  t = Y[1]; Y[1] = Y[2]; Y[2] = t;
        f = z[1]; z[1] = z[2]; z[2] = f;  f = rw[1]; rw[1] = rw[2]; rw[2] = f;
        cv = v[1]; v[1] = v[2]; v[2] = cv;
        area2 = -area2;
    }
    const int64_t A2 = -area2;
    const float invA = 1.0f / (float)A2;
    const float dz1 = z[1] - z[0], dz2 = z[2] - z[0];

    /* edge i is opposite vertex i: (1->2), (2->0), (0->1); F_i >= 0 inside */
    int64_t dx[3], dy[3], xa[3], ya[3];
    int bias[3];
    for (int i = 0; i < 3; ++i) {
        int a = (i + 1) % 3, b = (i + 2) % 3;
        dx[i] = (int64_t)X[b] - X[a];
        dy[i] = (int64_t)Y[b] - Y[a];
        xa[i] = X[a]; ya[i] = Y[a];
        /* top-left rule in image space (y down): a sample exactly on the edge belongs to the
         * triangle only for left edges (dy > 0) and top edges (dy == 0, dx < 0) */
        bias[i] = (dy[i] > 0 || (dy[i] == 0 && dx[i] < 0)) ? 0 : -1;
    }

    int32_t xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int i = 1; i < 3; ++i) {
        if (X[i] < xmin) xmin = X[i];
        if (X[i] > xmax) xmax = X[i];
        if (Y[i] < ymin) ymin = Y[i];
        if (Y[i] > ymax) ymax = Y[i];
    }
    /* pixel (i,j) is sampled at (256 i + 128, 256 j + 128) */
    int64_t i0 = ((int64_t)xmin - 128 + 255) >> 8, i1 = ((int64_t)xmax - 128) >> 8;
    int64_t j0 = ((int64_t)ymin - 128 + 255) >> 8, j1 = ((int64_t)ymax - 128) >> 8;
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (i1 > T->W - 1) i1 = T->W - 1;
    if (j1 > T->H - 1) j1 = T->H - 1;

    uint8_t flat_rgba[4];
    if (flat) shade(L, v[0].n, col, flat_rgba);

    for (int64_t j = j0; j <= j1; ++j) {
        for (int64_t i = i0; i <= i1; ++i) {
            int64_t px = 256 * i + 128, py = 256 * j + 128;
            int64_t F[3];
            int inside = 1;
            for (int e = 0; e < 3; ++e) {
                F[e] = dy[e] * (px - xa[e]) - dx[e] * (py - ya[e]);
                if (F[e] + bias[e] < 0) inside = 0;
            }
            if (!inside) continue;
            float b1 = (float)F[1] * invA, b2 = (float)F[2] * invA;
            float zp = fmaf(b2, dz2, fmaf(b1, dz1, z[0]));
            size_t o = (size_t)j * T->W + (size_t)i;
            /* depth LESS against clear 1.0; exact ties go to the earlier draw (sequential LESS) */
            uint32_t zb, zo;
            memcpy(&zb, &zp, 4);
            memcpy(&zo, &T->depth[o], 4);
            uint64_t key = ((uint64_t)zb << 32) | id, old = ((uint64_t)zo << 32) | T->prim[o];
            if (!(key < old)) continue;
            T->depth[o] = zp;
            T->prim[o] = id;
            uint8_t rgba[4];
            if (flat) {
                memcpy(rgba, flat_rgba, 4);
            } else {
                /* perspective-correct varyings: weights b_i / w_i (normalisation by their sum is
                 * dropped because the normal is re-normalised  [ref: basic.frag:33]) */
                float b0 = (float)F[0] * invA;
                float p0 = b0 * rw[0], p1 = b1 * rw[1], p2 = b2 * rw[2];
                float n[3];
                for (int k = 0; k < 3; ++k)
                    n[k] = fmaf(p2, v[2].n[k], fmaf(p1, v[1].n[k], p0 * v[0].n[k]));
                float l2 = fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0]));
                float inv = 1.0f / sqrtf(l2);
                n[0] *= inv; n[1] *= inv; n[2] *= inv;
                float pc[4] = {col[0], col[1], col[2], col[3]};
                if (tx->tex) {
                    /* [ref: basic.frag:31-32,37] base = mix(1, texel, useTexture); col = base * v_color * l */
                    const float sum = (p0 + p1) + p2;
                    const float tu = fmaf(p2, v[2].uv[0], fmaf(p1, v[1].uv[0], p0 * v[0].uv[0])) / sum;
                    const float tv = fmaf(p2, v[2].uv[1], fmaf(p1, v[1].uv[1], p0 * v[0].uv[1])) / sum;
                    float t3[3];
                    sample_bilinear(tx, tu, tv, t3);
                    const float a = tx->use, oma = 1.0f - tx->use;
                    for (int c = 0; c < 3; ++c) pc[c] *= fmaf(t3[c], a, oma);
                }
                shade(L, n, pc, rgba);
            }
            for (int c = 0; c < T->C; ++c) T->out[(size_t)c * T->H * T->W + o] = rgba[c];
        }
    }
}

static void render_scene(const orc_frame *f, int scene, const light_t *L, target_t *T) {
    const int W = T->W, H = T->H, C = T->C;
    uint8_t bg[4];
    for (int c = 0; c < 4; ++c) bg[c] = unorm8(f->bg[c]);       /* [ref: renderer.py:262-264] */
    for (int c = 0; c < C; ++c) memset(T->out + (size_t)c * H * W, bg[c], (size_t)H * W);
    for (int i = 0; i < H * W; ++i) { T->depth[i] = 1.0f; T->prim[i] = 0; }

    const float *VP = f->vp + (size_t)scene * 16;
    uint32_t slot = 0;
    for (int ni = 0; ni < f->n_nodes; ++ni) {
        const orc_node *nd = &f->nodes[ni];
        const orc_mesh *me = &nd->mesh;
        const int I = nd->instances_per_scene;
        const int two_sided = (me->flags & ORC_MESH_TWO_SIDED) != 0;
        texture_t tx = {NULL, 0, 0, 0.0f};
        {
            float ut = nd->use_texture;
            ut = !(ut == ut) || ut < 0.0f ? 0.0f : (ut > 1.0f ? 1.0f : ut);
            if (nd->tex && ut > 0.0f) { tx.tex = nd->tex; tx.w = nd->tex_w; tx.h = nd->tex_h; tx.use = ut; }
        }
        for (int inst = 0; inst < I; ++inst) {
            /* [ref: basic.vert:25-28]  id = shared ? inst : scene*I + inst   (evident intent) */
            size_t b = nd->shared ? (size_t)inst : (size_t)scene * I + inst;
            const float *M = nd->mats + b * 16;
            const float *col = nd->cols + b * 4;
            for (int t = 0; t < me->n_tris; ++t, ++slot) {
                cvert v[3];
                int flat = 1;
                const float *n0 = me->nrm + 3 * (size_t)me->idx[3 * t];
                for (int k = 0; k < 3; ++k) {
                    uint32_t vi = me->idx[3 * t + k];
                    const float *p = me->pos + 3 * (size_t)vi, *n = me->nrm + 3 * (size_t)vi;
                    float obj[4] = {p[0], p[1], p[2], 1.0f}, world[4];
                    mat_vec4(M, obj, world);
                    mat_vec4(VP, world, v[k].c);                 /* clip = VP * (M * v) */
                    xform_normal(M, n, v[k].n);
                    v[k].uv[0] = (tx.tex && me->uv) ? me->uv[2 * (size_t)vi] : 0.0f;
                    v[k].uv[1] = (tx.tex && me->uv) ? me->uv[2 * (size_t)vi + 1] : 0.0f;
                    if (memcmp(n, n0, 12) != 0) flat = 0;
                }
                /* trivial reject against the tile frustum (fragments outside the tile are
                 * discarded anyway  [ref: basic.frag:22-29]; z planes are the GL clip volume) */
                int rej = 0;
                for (int p = 0; p < 6 && !rej; ++p) {
                    int all_out = 1;
                    for (int k = 0; k < 3; ++k) {
                        float a = v[k].c[p >> 1], w = v[k].c[3];
                        int out = (p & 1) ? (a > w) : (a < -w);
                        if (!out) all_out = 0;
                    }
                    if (all_out) rej = 1;
                }
                if (rej) continue;
                int need_clip = 0;
                for (int k = 0; k < 3; ++k)
                    for (int p = 0; p < 5; ++p)
                        if (plane_dist(&v[k], p) < 0.0f) need_clip = 1;
                const uint32_t id = slot + 1;
                if (!need_clip) {
                    raster_tri(T, L, v, col, flat, two_sided, id, &tx);
                } else {
                    cvert poly[MAX_POLY];
                    int n = clip_poly(v, poly);
                    for (int k = 1; k + 1 < n; ++k) {
                        cvert tri[3] = {poly[0], poly[k], poly[k + 1]};
                        raster_tri(T, L, tri, col, flat, two_sided, id, &tx);
                    }
                }
            }
        }
    }
}

static int check_frame(const orc_frame *f) {
    if (!f || !f->out || !f->vp) return -1;
    if (f->tile_w < 1 || f->tile_h < 1 || f->tile_w > 2048 || f->tile_h > 2048) return -1;
    if (f->channels != 3 && f->channels != 4) return -1;
    if (f->scene_begin < 0 || f->scene_count < 0 || f->scene_begin + f->scene_count > f->num_scenes) return -1;
    if (f->n_nodes < 0 || (f->n_nodes > 0 && !f->nodes)) return -1;
    for (int i = 0; i < f->n_nodes; ++i) {
        const orc_node *n = &f->nodes[i];
        if (!n->mats || !n->cols || n->instances_per_scene < 0) return -1;
        if (n->mesh.n_tris > 0 && (!n->mesh.pos || !n->mesh.nrm || !n->mesh.idx)) return -1;
    }
    return 0;
}

static void make_light(const orc_frame *f, light_t *L) {
    for (int c = 0; c < 3; ++c) { L->amb[c] = f->ambient[c]; L->dcol[c] = f->dir_col[c]; }
    const float *d = f->dir_dir;                                   /* [ref: basic.frag:34] normalize(dirLightDir) */
    float l2 = fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]));
    float inv = 1.0f / sqrtf(l2);
    for (int c = 0; c < 3; ++c) L->ldir[c] = d[c] * inv;
    L->s = fminf(fmaxf(f->strength, 0.0f), 1.0f);
}

/* ---- scene-parallel driver: a persistent pool of pthreads (created on first use, grown on demand) that pull
 * chunks of scenes from a shared counter.  Persistent + dynamic so that a timed run (bench.py's CPU baseline)
 * measures rasterisation, not 32 pthread_create calls per frame or the slowest static slice. */
#define ORC_CHUNK 16
#define ORC_MAX_THREADS 256

static struct {
    pthread_mutex_t mu;
    pthread_cond_t go, done;
    pthread_t th[ORC_MAX_THREADS];
    int n_threads;            /* workers created */
    int want;                 /* workers taking part in the current frame */
    unsigned long generation; /* bumped per frame */
    int running;              /* participants that have not finished the current frame */
    const orc_frame *f;
    int next;                 /* next scene to hand out (atomic) */
    int end;
} pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, 0, 0, 0, NULL, 0, 0};

static void render_chunks(const orc_frame *f) {
    light_t L;
    make_light(f, &L);
    target_t T;
    T.W = f->tile_w; T.H = f->tile_h; T.C = f->channels;
    size_t npx = (size_t)T.W * T.H;
    T.depth = (float *)malloc(npx * sizeof(float));
    T.prim = (uint32_t *)malloc(npx * sizeof(uint32_t));
    for (;;) {
        int s0 = __atomic_fetch_add(&pool.next, ORC_CHUNK, __ATOMIC_RELAXED);
        if (s0 >= pool.end) break;
        int s1 = s0 + ORC_CHUNK < pool.end ? s0 + ORC_CHUNK : pool.end;
        for (int s = s0; s < s1; ++s) {
            T.out = f->out + (size_t)s * T.C * npx;
            render_scene(f, s, &L, &T);
        }
    }
    free(T.depth);
    free(T.prim);
}

static void *pool_worker(void *arg) {
    const int me = (int)(size_t)arg;
    unsigned long seen = 0;
    pthread_mutex_lock(&pool.mu);
    for (;;) {
        while (pool.generation == seen || me >= pool.want) {
            if (pool.generation != seen) seen = pool.generation;     /* a frame this worker sits out */
            pthread_cond_wait(&pool.go, &pool.mu);
        }
        seen = pool.generation;
        const orc_frame *f = pool.f;
        pthread_mutex_unlock(&pool.mu);
        render_chunks(f);
        pthread_mutex_lock(&pool.mu);
        if (--pool.running == 0) pthread_cond_signal(&pool.done);
    }
    return NULL;
}

static pthread_mutex_t frame_mu = PTHREAD_MUTEX_INITIALIZER;    /* one frame at a time through the pool */

static int render_locked(const orc_frame *f) {
    int nt = f->n_threads < 1 ? 1 : f->n_threads;
    if (nt > ORC_MAX_THREADS) nt = ORC_MAX_THREADS;
    const int chunks = (f->scene_count + ORC_CHUNK - 1) / ORC_CHUNK;
    if (nt > chunks) nt = chunks > 0 ? chunks : 1;
    pthread_mutex_lock(&pool.mu);
    pool.f = f;
    pool.next = f->scene_begin;
    pool.end = f->scene_begin + f->scene_count;
    if (nt == 1) {
        pthread_mutex_unlock(&pool.mu);
        render_chunks(f);
        return 0;
    }
    while (pool.n_threads < nt - 1) {             /* the caller is the nt-th participant */
        if (pthread_create(&pool.th[pool.n_threads], NULL, pool_worker, (void *)(size_t)pool.n_threads) != 0) break;
        pthread_detach(pool.th[pool.n_threads]);
        pool.n_threads++;
    }
    pool.want = nt - 1 < pool.n_threads ? nt - 1 : pool.n_threads;
    pool.running = pool.want;
    pool.generation++;
    pthread_cond_broadcast(&pool.go);
    pthread_mutex_unlock(&pool.mu);
    render_chunks(f);
    pthread_mutex_lock(&pool.mu);
    while (pool.running > 0) pthread_cond_wait(&pool.done, &pool.mu);
    pthread_mutex_unlock(&pool.mu);
    return 0;
}

int orc_render(const orc_frame *f) {
    if (check_frame(f)) return -1;
    pthread_mutex_lock(&frame_mu);
    const int rc = render_locked(f);
    pthread_mutex_unlock(&frame_mu);
    return rc;
}

int orc_render_scene_debug(const orc_frame *f, int scene, uint8_t *out_chw, float *depth, uint32_t *prim) {
    if (check_frame(f) || scene < 0 || scene >= f->num_scenes || !out_chw || !depth || !prim) return -1;
    light_t L;
    make_light(f, &L);
    target_t T;
    T.W = f->tile_w; T.H = f->tile_h; T.C = f->channels;
    T.out = out_chw; T.depth = depth; T.prim = prim;
    render_scene(f, scene, &L, &T);
    return 0;
}
