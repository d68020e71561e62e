// transforms.cuh -- instance-transform kernels (reference node.py:116-154, shader_context.py:47-84).
#pragma once
#include "common.cuh"

namespace pbr {

// ------------------------------------------------------------------------------------------------
// instance-transform kernels
// ------------------------------------------------------------------------------------------------
// plane c of a [C, hw] byte image = byte c of the packed colour
__global__ void nsmid_kernel(int *out) {
    unsigned n;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
    *out = (int)n;
}

__global__ void fill_planes_kernel(unsigned char *dst, int hw, int C, unsigned rgba) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hw * C) dst[i] = (unsigned char)((rgba >> (8 * (i / hw))) & 255u);
}

__global__ void pack_transforms_kernel(float *__restrict__ tr, const float *__restrict__ rot,
                                       const float *__restrict__ scale, float *__restrict__ out, int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    float *T = tr + (size_t)b * 16;
    const float *R = rot + (size_t)b * 9;
    const float s = scale[b];
    float m[16];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            m[4 * i + j] = R[3 * i + j] * s;
            T[4 * i + j] = m[4 * i + j];
        }
        m[4 * i + 3] = T[4 * i + 3];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) m[12 + j] = T[12 + j];
    float4 *o = reinterpret_cast<float4 *>(out + (size_t)b * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(m[j], m[4 + j], m[8 + j], m[12 + j]);   // column j
}

constexpr int MAX_POSES = 8;
struct PoseBatch {
    pbr_pose_desc p[MAX_POSES];
    int n;
};

__device__ __forceinline__ float chan(const pbr_channel &c, int b) {
    return c.ptr ? __ldg(c.ptr + (size_t)b * c.stride) : c.constant;
}

__global__ void compose_kernel(const __grid_constant__ PoseBatch pb) {
    // programmatic dependent launch: the raster kernel that follows may start its prologue (background
    // copy, mask clearing) now; it executes griddepcontrol.wait before it reads the matrices written here
    asm volatile("griddepcontrol.launch_dependents;");
    const pbr_pose_desc &d = pb.p[blockIdx.y];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.n_instances) return;
    const float x = chan(d.pos[0], b), y = chan(d.pos[1], b), z = chan(d.pos[2], b);
    const float h = chan(d.hpr[0], b), p = chan(d.hpr[1], b), r = chan(d.hpr[2], b);
    const float s = chan(d.scale, b);
    float sh, ch, sp, cp, sr, cr;
    sincosf(h, &sh, &ch);
    sincosf(p, &sp, &cp);
    sincosf(r, &sr, &cr);
    // R = Rz(h) Ry(p) Rx(r)   (reference shader_context.py:47-84)
    const float r00 = ch * cp, r01 = ch * sp * sr - sh * cr, r02 = ch * sp * cr + sh * sr;
    const float r10 = sh * cp, r11 = sh * sp * sr + ch * cr, r12 = sh * sp * cr - ch * sr;
    const float r20 = -sp, r21 = cp * sr, r22 = cp * cr;
    float4 *o = reinterpret_cast<float4 *>(d.out_mats + (size_t)b * 16);
    o[0] = make_float4(r00 * s, r10 * s, r20 * s, 0.0f);
    o[1] = make_float4(r01 * s, r11 * s, r21 * s, 0.0f);
    o[2] = make_float4(r02 * s, r12 * s, r22 * s, 0.0f);
    o[3] = make_float4(x, y, z, 1.0f);
}


}  // namespace pbr
