// t4d_activate.cu -- fused parameter activations of params2rendervar (reference helpers.py:91-112), sm_100a.
//
//     'rotations': torch.nn.functional.normalize(params['unnorm_rotations'])     x / max(||x||_2, 1e-12)
//     'opacities': torch.sigmoid(params['logit_opacities'])
//     'scales':    torch.exp(params['log_scales'])
// PyTorch runs these as 5 small kernels forward and ~12 backward on every iteration; at 8 k Gaussians each is pure launch
// latency.  One kernel forward, one backward, thread = Gaussian, 16-byte loads/stores for the quaternion.
#include <cuda_runtime.h>
#include "../../include/topo4d_b200.h"

namespace {

__global__ void __launch_bounds__(256) activate_fwd_kernel(const float4* __restrict__ q_in, const float* __restrict__ logit,
                                                           const float* __restrict__ log_s, int n, float4* __restrict__ q_out,
                                                           float* __restrict__ opac, float* __restrict__ scales)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = q_in[i];
    const float inv = 1.0f / fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    q_out[i] = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
    opac[i] = 1.0f / (1.0f + expf(-logit[i]));
    #pragma unroll
    for (int k = 0; k < 3; k++) scales[3 * i + k] = expf(log_s[3 * i + k]);
}

__global__ void __launch_bounds__(256) activate_bwd_kernel(const float4* __restrict__ q_in, const float* __restrict__ opac,
                                                           const float* __restrict__ scales, const float4* __restrict__ g_q,
                                                           const float* __restrict__ g_o, const float* __restrict__ g_s, int n,
                                                           float4* __restrict__ d_q, float* __restrict__ d_logit,
                                                           float* __restrict__ d_log_s)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (d_q) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g_q) {
            const float4 q = q_in[i], g = g_q[i];
            const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
            if (nrm > 1e-12f) {
                const float inv = 1.0f / nrm;
                const float hx = q.x * inv, hy = q.y * inv, hz = q.z * inv, hw = q.w * inv;
                const float dot = hx * g.x + hy * g.y + hz * g.z + hw * g.w;
                r = make_float4((g.x - hx * dot) * inv, (g.y - hy * dot) * inv, (g.z - hz * dot) * inv, (g.w - hw * dot) * inv);
            } else {
                r = make_float4(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f, g.w * 1e12f);     // clamped branch: x / eps
            }
        }
        d_q[i] = r;
    }
    if (d_logit) {
        const float o = opac[i];
        d_logit[i] = g_o ? g_o[i] * o * (1.0f - o) : 0.f;
    }
    if (d_log_s) {
        #pragma unroll
        for (int k = 0; k < 3; k++) d_log_s[3 * i + k] = g_s ? g_s[3 * i + k] * scales[3 * i + k] : 0.f;
    }
}

}  // namespace

extern "C" int t4d_activate(const float* unnorm_rotations, const float* logit_opacities, const float* log_scales, int32_t N,
                            float* rotations, float* opacities, float* scales, gs_stream_t stream)
{
    if (N < 0) return GS_E_BAD_ARGS;
    if (N == 0) return 0;
    if (!unnorm_rotations || !logit_opacities || !log_scales || !rotations || !opacities || !scales) return GS_E_BAD_ARGS;
    if ((((uintptr_t)unnorm_rotations | (uintptr_t)rotations) & 15u) != 0) return GS_E_BAD_ARGS;
    activate_fwd_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)unnorm_rotations, logit_opacities,
                                                                          log_scales, N, (float4*)rotations, opacities, scales);
    return cudaGetLastError() == cudaSuccess ? 0 : GS_E_CUDA;
}

extern "C" int t4d_activate_backward(const float* unnorm_rotations, const float* opacities, const float* scales,
                                     const float* dL_drotations, const float* dL_dopacities, const float* dL_dscales, int32_t N,
                                     float* dL_dunnorm_rotations, float* dL_dlogit_opacities, float* dL_dlog_scales,
                                     gs_stream_t stream)
{
    if (N < 0) return GS_E_BAD_ARGS;
    if (N == 0) return 0;
    if (!unnorm_rotations || !opacities || !scales) return GS_E_BAD_ARGS;
    if ((((uintptr_t)unnorm_rotations | (uintptr_t)dL_drotations | (uintptr_t)dL_dunnorm_rotations) & 15u) != 0) return GS_E_BAD_ARGS;
    activate_bwd_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)unnorm_rotations, opacities, scales, (const float4*)dL_drotations, dL_dopacities, dL_dscales, N,
        (float4*)dL_dunnorm_rotations, dL_dlogit_opacities, dL_dlog_scales);
    return cudaGetLastError() == cudaSuccess ? 0 : GS_E_CUDA;
}
