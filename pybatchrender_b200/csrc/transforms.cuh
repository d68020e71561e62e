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
    PoseDev p[MAX_POSES];
    int n_inst[MAX_POSES];
    int n;
};

__global__ void compose_kernel(const __grid_constant__ PoseBatch pb) {
    const PoseDev &d = pb.p[blockIdx.y];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.n_inst[blockIdx.y]) return;
    float M[16];
    pose_matrix(d, (size_t)b, M);
    float4 *o = reinterpret_cast<float4 *>(d.out_mats + (size_t)b * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(M[4 * j], M[4 * j + 1], M[4 * j + 2], M[4 * j + 3]);
}

}  // namespace pbr
