// t4d_optim.cu -- fused multi-tensor Adam step with per-tensor learning rates and pinned rows, sm_100a.
//
// Replaces the optimiser tail of one Topo4D iteration (reference train.py:672-700):
//     optimizer.step()            torch.optim.Adam(param_groups, lr=0.0, eps=1e-15), one group per named parameter
//                                 with its own lr (train.py:272-297), betas (0.9, 0.999), no weight decay / amsgrad
//     params[name][mask] = const  ~12 boolean-mask overwrites that freeze facial regions after every step
//                                 (train.py:676-700)
// PyTorch issues a foreach kernel chain per group plus one index_put per overwrite; here ONE launch walks all
// tensors: m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2, p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps),
// then rows whose pin mask is set are overwritten with their pinned values (the caller merges the reference's
// overwrite list into one mask + value table per tensor, once per timestep).  Streaming, HBM-bound: 16 B read +
// 12 B written per element, 16-byte vector accesses when the segment allows.
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/topo4d_b200.h"

namespace {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_ELEMS_PER_BLOCK = ADAM_THREADS * 4;

struct AdamSeg {
    float* param; const float* grad; float* m; float* v;
    const uint8_t* pin_mask; const float* pin_values;
    long long count; int row_width;
    float step_size, inv_sqrt_bc2;      // host-computed bias terms (step_dev == nullptr)
    float lr;
    const float* lr_dev;                // capturable mode: the group's learning rate, re-read at every replay
    int* step_dev;                      // capturable mode: completed steps live on the device, bias terms computed here
    int first_block;
};
struct AdamParams {
    AdamSeg seg[T4D_ADAM_MAX_SEGMENTS];
    int nseg;
    float beta1, beta2, eps;
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float step_size, float inv_sqrt_bc2, float b1, float b2, float eps)
{
    m = fmaf(b1, m, (1.0f - b1) * g);
    v = fmaf(b2, v, (1.0f - b2) * g * g);
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    return p - step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamParams P)
{
    int si = 0;
    #pragma unroll 1
    for (int k = 1; k < P.nseg; k++) if ((int)blockIdx.x >= P.seg[k].first_block) si = k;
    const AdamSeg& s = P.seg[si];
    float step_size = s.step_size, inv_sqrt_bc2 = s.inv_sqrt_bc2;
    if (s.step_dev) {
        // capturable mode (the launch sits in a CUDA graph): t = completed steps + 1 is read from device memory; the
        // counters are advanced by adam_advance_kernel right behind this launch
        __shared__ float s_bias[2];
        if (threadIdx.x == 0) {
            const double t = (double)(*s.step_dev + 1);
            const double lr = s.lr_dev ? (double)*s.lr_dev : (double)s.lr;
            s_bias[0] = (float)(lr / (1.0 - pow((double)P.beta1, t)));
            s_bias[1] = (float)(1.0 / sqrt(1.0 - pow((double)P.beta2, t)));
        }
        __syncthreads();
        step_size = s_bias[0]; inv_sqrt_bc2 = s_bias[1];
    }
    const long long base = (long long)(blockIdx.x - s.first_block) * ADAM_ELEMS_PER_BLOCK + threadIdx.x * 4;
    if (base >= s.count) return;
    const bool vec = base + 4 <= s.count && ((((uintptr_t)s.param | (uintptr_t)s.grad | (uintptr_t)s.m | (uintptr_t)s.v) & 15u) == 0);
    float p[4], g[4], m[4], v[4];
    const int n = (int)min(4LL, s.count - base);
    if (vec) {
        const float4 p4 = *reinterpret_cast<const float4*>(s.param + base), g4 = *reinterpret_cast<const float4*>(s.grad + base);
        const float4 m4 = *reinterpret_cast<const float4*>(s.m + base), v4 = *reinterpret_cast<const float4*>(s.v + base);
        p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w; g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
        m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w; v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
    } else {
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool in = k < n;
            p[k] = in ? s.param[base + k] : 0.f; g[k] = in ? s.grad[base + k] : 0.f;
            m[k] = in ? s.m[base + k] : 0.f; v[k] = in ? s.v[base + k] : 0.f;
        }
    }
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        p[k] = adam_one(p[k], g[k], m[k], v[k], step_size, inv_sqrt_bc2, P.beta1, P.beta2, P.eps);
        if (s.pin_mask && k < n) {
            const long long e = base + k, row = e / s.row_width;
            if (s.pin_mask[row]) p[k] = s.pin_values ? s.pin_values[e] : 0.f;
        }
    }
    if (vec) {
        *reinterpret_cast<float4*>(s.param + base) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(s.m + base) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(s.v + base) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        #pragma unroll
        for (int k = 0; k < 4; k++) if (k < n) { s.param[base + k] = p[k]; s.m[base + k] = m[k]; s.v[base + k] = v[k]; }
    }
}

__global__ void adam_advance_kernel(const __grid_constant__ AdamParams P)
{
    const int k = threadIdx.x;
    if (k >= P.nseg || !P.seg[k].step_dev) return;
    for (int j = 0; j < k; j++) if (P.seg[j].step_dev == P.seg[k].step_dev) return;     // shared counter: advance once
    *P.seg[k].step_dev += 1;
}

}  // namespace

extern "C" int t4d_adam_step(const T4dAdamSegment* segs, int32_t nseg, float beta1, float beta2, float eps, gs_stream_t stream)
{
    if (!segs || nseg < 1 || nseg > T4D_ADAM_MAX_SEGMENTS) return GS_E_BAD_ARGS;
    AdamParams P;
    P.nseg = 0; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps;
    long long blocks = 0;
    bool any_dev = false;
    for (int i = 0; i < nseg; i++) {
        const T4dAdamSegment& a = segs[i];
        if (a.count < 0 || (!a.step_device && a.step < 1) || a.row_width < 1) return GS_E_BAD_ARGS;
        if (a.count == 0) continue;
        if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq) return GS_E_BAD_ARGS;
        if (a.pin_mask && a.count % a.row_width != 0) return GS_E_BAD_ARGS;
        AdamSeg& s = P.seg[P.nseg++];
        s.param = a.param; s.grad = a.grad; s.m = a.exp_avg; s.v = a.exp_avg_sq;
        s.pin_mask = a.pin_mask; s.pin_values = a.pin_values; s.count = a.count; s.row_width = a.row_width;
        // torch.optim.Adam: step_size = lr / (1 - beta1^t); denom = sqrt(v) / sqrt(1 - beta2^t) + eps  (bias terms in fp64)
        s.lr = a.lr; s.step_dev = a.step_device; s.lr_dev = a.step_device ? a.lr_device : NULL;
        s.step_size = 0.f; s.inv_sqrt_bc2 = 1.f;
        if (!a.step_device) {
            const double bc1 = 1.0 - pow((double)beta1, (double)a.step), bc2 = 1.0 - pow((double)beta2, (double)a.step);
            s.step_size = (float)((double)a.lr / bc1);
            s.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
        }
        any_dev = any_dev || a.step_device != NULL;
        s.first_block = (int)blocks;
        blocks += (a.count + ADAM_ELEMS_PER_BLOCK - 1) / ADAM_ELEMS_PER_BLOCK;
        if (blocks > 0x7fffffffLL) return GS_E_UNSUPPORTED;
    }
    if (P.nseg == 0) return 0;
    adam_kernel<<<(unsigned)blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(P);
    if (any_dev) adam_advance_kernel<<<1, T4D_ADAM_MAX_SEGMENTS, 0, (cudaStream_t)stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : GS_E_CUDA;
}
