// t4d_loss.cu -- fused per-iteration image loss of Topo4D's optimisation loop, forward AND backward, sm_100a.
//
// Replaces, for the rendered image of one iteration (reference train.py:307-317):
//     im   = exp(cam_m[id])[:, None, None] * im + cam_c[id][:, None, None]              (train.py:310)
//     loss = 0.8 * l1_loss_v1(im, gt) + 0.2 * (1.0 - calc_ssim(im, gt))                  (train.py:317)
// with l1_loss_v1 = mean |x - y| (helpers.py:115-116) and calc_ssim/_ssim the 11x11 Gaussian-window SSIM
// (sigma 1.5, zero padding, per-channel "groups" convolution, mean over C*H*W; external.py:71-116), plus the whole
// autograd backward of that expression down to dL/d(rendered image), dL/dcam_m, dL/dcam_c.  PyTorch runs this as
// five grouped 11x11 conv2d + ~20 elementwise kernels forward and as many again backward; here it is two tile
// kernels and a tiny finalize:
//   ssim_fwd_kernel   one CTA per 32x32 tile of one (view, channel) plane: 42x42 halo tile of x = affine(render) and
//                     y = target into shared memory, SEPARABLE window (11 + 11 taps instead of 121) on the five
//                     moments x, y, x^2, y^2, xy with register sliding windows, SSIM map value, and the three partial
//                     derivative maps dS/dmu1, dS/dE[x^2], dS/dE[xy] written once to HBM; per-CTA partial sums of
//                     SSIM and |x - y| (deterministic two-level reduction, no atomics).
//   ssim_bwd_kernel   same tiling over the three derivative maps: separable window again (the window is symmetric,
//                     so the adjoint of the zero-padded correlation is the same correlation), then
//                     dSSIM/dx = G1 + 2 x G2 + y G3, the L1 sign term, the affine chain rule, dL/drender out, and
//                     per-CTA partial sums for dL/dcam_m, dL/dcam_c.
//   finalize          fp64 sums of the partials -> per-view {l1, ssim, total} and the camera-affine gradients.
// Because both reductions are plain means, dL/dim does not depend on the loss VALUE: forward and backward are
// enqueued back to back with no host round trip.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../include/topo4d_b200.h"

namespace {

constexpr int TX = 32, TY = 32, RAD = 5, WIN = 2 * RAD + 1;
constexpr int IN = TX + 2 * RAD;        // 42 input rows / columns per tile
constexpr int PITCH = 44;               // padded row pitch of the input planes (16-byte aligned strips)
constexpr int THREADS = 256;
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct Win { float w[WIN]; unsigned long long ww[WIN]; };   // taps, and the same taps as packed (w, w) fp32 pairs

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2: one issue slot for two FMAs) ----
// The five SSIM moments pair up naturally -- (x, y), (x^2, y^2) and xy alone -- so the window passes carry pairs in
// 64-bit registers: 4 instructions per tap instead of 7 in the horizontal pass, 3 instead of 5 in the vertical one.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct LossParams {
    const float* render; const float* target; const float* cam_m; const float* cam_c;
    int V, H, W, nbx, nby;
    float* maps;            // [V*3][3][H*W]
    float* part_a;          // [V*3][nblk][2]  ssim sum, |x-y| sum
    float* part_b;          // [V*3][nblk][2]  sum dL/dim, sum dL/dim * (x - b)
    float* d_render;
    float* loss; float* d_cam_m; float* d_cam_c;
    float w_l1, w_ssim;
    Win win;
};

__device__ __forceinline__ float2 block_sum2(float a, float b, float2* s_red)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_red[w] = make_float2(a, b);
    __syncthreads();
    float2 r = make_float2(0.f, 0.f);
    if (w == 0) {
        r = lane < THREADS / 32 ? s_red[lane] : make_float2(0.f, 0.f);
        #pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            r.x += __shfl_xor_sync(0xffffffffu, r.x, o);
            r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
        }
    }
    return r;       // valid in thread 0
}

__global__ void __launch_bounds__(THREADS) ssim_fwd_kernel(const LossParams p)
{
    __shared__ __align__(16) f32x2 s_xy[IN][PITCH];          // (x, y) pairs
    __shared__ __align__(16) f32x2 s_h01[IN][TX];            // horizontal pass of (x, y)
    __shared__ __align__(16) f32x2 s_h23[IN][TX];            //                    (x^2, y^2)
    __shared__ __align__(16) float s_h4[IN][TX];             //                    xy
    __shared__ float2 s_red[THREADS / 32];
    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const size_t HW = (size_t)p.H * p.W;
    const float a = p.cam_m ? expf(p.cam_m[plane]) : 1.0f, b = p.cam_c ? p.cam_c[plane] : 0.0f;
    const float* __restrict__ rp = p.render + (size_t)plane * HW;
    const float* __restrict__ tp = p.target + (size_t)plane * HW;
    const int tx = tid & 31, ty = tid >> 5;

    // the convolution pads the AFFINE image with zeros (conv2d padding=5 on `im`): outside pixels are x = y = 0.
    // All of a thread's global loads are issued before the first shared-memory store (one exposed latency, not six);
    // tiles whose halo lies inside the image (9 of 10 at 1080p) take a path without per-element bounds tests.
    {
        constexpr int NR = (IN + THREADS / 32 - 1) / (THREADS / 32);      // 6 row slots per thread
        float xr[NR][2], yr[NR][2];
        const bool interior = x0 >= RAD && y0 >= RAD && x0 - RAD + IN <= p.W && y0 - RAD + IN <= p.H;   // CTA-uniform
        const int base = (y0 - RAD + ty) * p.W + (x0 - RAD + tx);          // 32-bit offsets inside one plane
        if (interior) {
            #pragma unroll
            for (int i = 0; i < NR; i++) {
                #pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const bool ok = (ty + (THREADS / 32) * i < IN) && (tx + 32 * cc < IN);
                    const int off = base + (THREADS / 32) * i * p.W + 32 * cc;
                    xr[i][cc] = ok ? fmaf(a, __ldg(rp + off), b) : 0.f;
                    yr[i][cc] = ok ? __ldg(tp + off) : 0.f;
                }
            }
        } else {
            #pragma unroll
            for (int i = 0; i < NR; i++) {
                const int r = ty + (THREADS / 32) * i, gy = y0 - RAD + r;
                #pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const int c = tx + 32 * cc, gx = x0 - RAD + c;
                    const bool ok = r < IN && c < IN && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
                    const int off = base + (THREADS / 32) * i * p.W + 32 * cc;
                    xr[i][cc] = ok ? fmaf(a, __ldg(rp + off), b) : 0.f;
                    yr[i][cc] = ok ? __ldg(tp + off) : 0.f;
                }
            }
        }
        #pragma unroll
        for (int i = 0; i < NR; i++) {
            const int r = ty + (THREADS / 32) * i;
            #pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int c = tx + 32 * cc;
                if (r < IN && c < IN) s_xy[r][c] = pk2(xr[i][cc], yr[i][cc]);
            }
        }
    }
    __syncthreads();

    // horizontal pass: strips of 4 adjacent outputs share a 14-wide register window of (x, y) pairs
    for (int s = tid; s < IN * (TX / 4); s += THREADS) {
        const int r = s >> 3, c0 = (s & 7) * 4;
        f32x2 v[14];
        #pragma unroll
        for (int q = 0; q < 7; q++) {
            const ulonglong2 f = *reinterpret_cast<const ulonglong2*>(&s_xy[r][c0 + 2 * q]);
            v[2 * q] = f.x; v[2 * q + 1] = f.y;
        }
        f32x2 h01[4], h23[4];
        float h4[4];
        #pragma unroll
        for (int o = 0; o < 4; o++) {
            f32x2 a01 = 0ull, a23 = 0ull;                       // (+0, +0)
            float hxy = 0.f;
            #pragma unroll
            for (int k = 0; k < WIN; k++) {
                const f32x2 pxy = v[o + k];
                const f32x2 wp = mul2(p.win.ww[k], pxy);       // (w x, w y)
                a01 = fma2(p.win.ww[k], pxy, a01);
                a23 = fma2(wp, pxy, a23);
                float wx, wy, px, py;
                upk2(wp, wx, wy); upk2(pxy, px, py);
                hxy = fmaf(wx, py, hxy);
            }
            h01[o] = a01; h23[o] = a23; h4[o] = hxy;
        }
        *reinterpret_cast<ulonglong2*>(&s_h01[r][c0]) = make_ulonglong2(h01[0], h01[1]);
        *reinterpret_cast<ulonglong2*>(&s_h01[r][c0 + 2]) = make_ulonglong2(h01[2], h01[3]);
        *reinterpret_cast<ulonglong2*>(&s_h23[r][c0]) = make_ulonglong2(h23[0], h23[1]);
        *reinterpret_cast<ulonglong2*>(&s_h23[r][c0 + 2]) = make_ulonglong2(h23[2], h23[3]);
        *reinterpret_cast<float4*>(&s_h4[r][c0]) = make_float4(h4[0], h4[1], h4[2], h4[3]);
    }
    __syncthreads();

    // vertical pass: thread (tx, ty) owns 4 consecutive rows of column tx
    float m[5][4];
    {
        f32x2 v[14];
        #pragma unroll
        for (int i = 0; i < 14; i++) v[i] = s_h01[ty * 4 + i][tx];
        #pragma unroll
        for (int o = 0; o < 4; o++) {
            f32x2 acc = 0ull;
            #pragma unroll
            for (int k = 0; k < WIN; k++) acc = fma2(p.win.ww[k], v[o + k], acc);
            upk2(acc, m[0][o], m[1][o]);
        }
        #pragma unroll
        for (int i = 0; i < 14; i++) v[i] = s_h23[ty * 4 + i][tx];
        #pragma unroll
        for (int o = 0; o < 4; o++) {
            f32x2 acc = 0ull;
            #pragma unroll
            for (int k = 0; k < WIN; k++) acc = fma2(p.win.ww[k], v[o + k], acc);
            upk2(acc, m[2][o], m[3][o]);
        }
        float u[14];
        #pragma unroll
        for (int i = 0; i < 14; i++) u[i] = s_h4[ty * 4 + i][tx];
        #pragma unroll
        for (int o = 0; o < 4; o++) {
            float acc = 0.f;
            #pragma unroll
            for (int k = 0; k < WIN; k++) acc = fmaf(p.win.w[k], u[o + k], acc);
            m[4][o] = acc;
        }
    }

    float ssim_sum = 0.f, l1_sum = 0.f;
    float* __restrict__ mp = p.maps + (size_t)plane * 3 * HW;
    #pragma unroll
    for (int o = 0; o < 4; o++) {
        const int gy = y0 + ty * 4 + o, gx = x0 + tx;
        if (gy < p.H && gx < p.W) {
            const float mu1 = m[0][o], mu2 = m[1][o];
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s11 = m[2][o] - mu1_sq, s22 = m[3][o] - mu2_sq, s12 = m[4][o] - mu12;
            const float A1 = 2.f * mu12 + SSIM_C1, A2 = 2.f * s12 + SSIM_C2;
            const float B1 = mu1_sq + mu2_sq + SSIM_C1, B2 = s11 + s22 + SSIM_C2;
            const float inv = __fdividef(1.0f, B1 * B2);      // rcp.approx: 1 ulp, far inside the 1e-5 budget
            const float S = (A1 * A2) * inv;
            ssim_sum += S;
            float cx, cy;
            upk2(s_xy[ty * 4 + o + RAD][tx + RAD], cx, cy);
            l1_sum += fabsf(cx - cy);
            if (p.d_render) {
                const size_t pix = (size_t)gy * p.W + gx;
                mp[pix] = inv * (2.f * mu2 * (A2 - A1) - 2.f * mu1 * S * (B2 - B1));       // dS/dmu1
                mp[HW + pix] = -S * B1 * inv;                                               // dS/dE[x^2]
                mp[2 * HW + pix] = 2.f * A1 * inv;                                          // dS/dE[xy]
            }
        }
    }
    const float2 tot = block_sum2(ssim_sum, l1_sum, s_red);
    if (tid == 0) {
        const size_t blk = (size_t)plane * p.nbx * p.nby + (size_t)blockIdx.y * p.nbx + blockIdx.x;
        reinterpret_cast<float2*>(p.part_a)[blk] = tot;
    }
}

__global__ void __launch_bounds__(THREADS) ssim_bwd_kernel(const LossParams p)
{
    __shared__ __align__(16) f32x2 s_d01[IN][PITCH];         // (dS/dmu1, dS/dE[x^2]) pairs
    __shared__ __align__(16) float s_d2[IN][PITCH];          // dS/dE[xy]
    __shared__ __align__(16) f32x2 s_g01[IN][TX];
    __shared__ __align__(16) float s_g2[IN][TX];
    __shared__ float2 s_red[THREADS / 32];
    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const size_t HW = (size_t)p.H * p.W;
    const float a = p.cam_m ? expf(p.cam_m[plane]) : 1.0f, b = p.cam_c ? p.cam_c[plane] : 0.0f;
    const float* __restrict__ mp = p.maps + (size_t)plane * 3 * HW;
    const int tx = tid & 31, ty = tid >> 5;

    {
        constexpr int NR = (IN + THREADS / 32 - 1) / (THREADS / 32);
        float d[NR][2][3];
        const bool interior = x0 >= RAD && y0 >= RAD && x0 - RAD + IN <= p.W && y0 - RAD + IN <= p.H;   // CTA-uniform
        const int base = (y0 - RAD + ty) * p.W + (x0 - RAD + tx);
        const float* __restrict__ m1 = mp + HW;
        const float* __restrict__ m2 = mp + 2 * HW;
        #pragma unroll
        for (int i = 0; i < NR; i++) {
            const int r = ty + (THREADS / 32) * i, gy = y0 - RAD + r;
            #pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int c = tx + 32 * cc, gx = x0 - RAD + c;
                const bool ok = r < IN && c < IN && (interior || (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W));
                const int off = base + (THREADS / 32) * i * p.W + 32 * cc;
                d[i][cc][0] = ok ? __ldg(mp + off) : 0.f;
                d[i][cc][1] = ok ? __ldg(m1 + off) : 0.f;
                d[i][cc][2] = ok ? __ldg(m2 + off) : 0.f;
            }
        }
        #pragma unroll
        for (int i = 0; i < NR; i++) {
            const int r = ty + (THREADS / 32) * i;
            #pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int c = tx + 32 * cc;
                if (r < IN && c < IN) { s_d01[r][c] = pk2(d[i][cc][0], d[i][cc][1]); s_d2[r][c] = d[i][cc][2]; }
            }
        }
    }
    __syncthreads();

    for (int s = tid; s < IN * (TX / 4); s += THREADS) {
        const int r = s >> 3, c0 = (s & 7) * 4;
        f32x2 v[14];
        float u[16];
        #pragma unroll
        for (int q = 0; q < 7; q++) {
            const ulonglong2 f = *reinterpret_cast<const ulonglong2*>(&s_d01[r][c0 + 2 * q]);
            v[2 * q] = f.x; v[2 * q + 1] = f.y;
        }
        #pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 f = *reinterpret_cast<const float4*>(&s_d2[r][c0 + 4 * q]);
            u[4 * q] = f.x; u[4 * q + 1] = f.y; u[4 * q + 2] = f.z; u[4 * q + 3] = f.w;
        }
        f32x2 h01[4];
        float h2[4];
        #pragma unroll
        for (int o = 0; o < 4; o++) {
            f32x2 a01 = 0ull;
            float a2 = 0.f;
            #pragma unroll
            for (int k = 0; k < WIN; k++) {
                a01 = fma2(p.win.ww[k], v[o + k], a01);
                a2 = fmaf(p.win.w[k], u[o + k], a2);
            }
            h01[o] = a01; h2[o] = a2;
        }
        *reinterpret_cast<ulonglong2*>(&s_g01[r][c0]) = make_ulonglong2(h01[0], h01[1]);
        *reinterpret_cast<ulonglong2*>(&s_g01[r][c0 + 2]) = make_ulonglong2(h01[2], h01[3]);
        *reinterpret_cast<float4*>(&s_g2[r][c0]) = make_float4(h2[0], h2[1], h2[2], h2[3]);
    }
    __syncthreads();

    float g[3][4];
    {
        f32x2 v[14];
        float u[14];
        #pragma unroll
        for (int i = 0; i < 14; i++) { v[i] = s_g01[ty * 4 + i][tx]; u[i] = s_g2[ty * 4 + i][tx]; }
        #pragma unroll
        for (int o = 0; o < 4; o++) {
            f32x2 acc = 0ull;
            float a2 = 0.f;
            #pragma unroll
            for (int k = 0; k < WIN; k++) {
                acc = fma2(p.win.ww[k], v[o + k], acc);
                a2 = fmaf(p.win.w[k], u[o + k], a2);
            }
            upk2(acc, g[0][o], g[1][o]);
            g[2][o] = a2;
        }
    }

    const float inv_n = 1.0f / (3.0f * (float)p.H * (float)p.W);
    const float* __restrict__ rp = p.render + (size_t)plane * HW;
    const float* __restrict__ tp = p.target + (size_t)plane * HW;
    float* __restrict__ dp = p.d_render + (size_t)plane * HW;
    float sum_c = 0.f, sum_m = 0.f;
    #pragma unroll
    for (int o = 0; o < 4; o++) {
        const int gy = y0 + ty * 4 + o, gx = x0 + tx;
        if (gy < p.H && gx < p.W) {
            const size_t pix = (size_t)gy * p.W + gx;
            const float rv = __ldg(rp + pix);
            const float x = fmaf(a, rv, b), y = __ldg(tp + pix);
            const float dssim = g[0][o] + 2.f * x * g[1][o] + y * g[2][o];
            const float df = x - y;
            const float sgn = (df > 0.f ? 1.f : 0.f) - (df < 0.f ? 1.f : 0.f);
            const float dim = inv_n * (p.w_l1 * sgn - p.w_ssim * dssim);      // dL/d(affine image)
            dp[pix] = a * dim;
            sum_c += dim;
            sum_m += dim * (x - b);                                           // d im / d cam_m = exp(cam_m) * render
        }
    }
    const float2 tot = block_sum2(sum_c, sum_m, s_red);
    if (tid == 0) {
        const size_t blk = (size_t)plane * p.nbx * p.nby + (size_t)blockIdx.y * p.nbx + blockIdx.x;
        reinterpret_cast<float2*>(p.part_b)[blk] = tot;
    }
}

// one CTA per view: fp64 sums of the per-CTA partials (fixed order: deterministic)
__global__ void __launch_bounds__(THREADS) loss_finalize_kernel(const LossParams p)
{
    __shared__ double s_acc[THREADS / 32][4];
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nblk = p.nbx * p.nby;
    double ssim = 0.0, l1 = 0.0;
    double dc[3] = {0.0, 0.0, 0.0}, dm[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < 3; c++) {
        const size_t base = ((size_t)v * 3 + c) * nblk;
        for (int i = tid; i < nblk; i += THREADS) {
            const float2 pa = reinterpret_cast<const float2*>(p.part_a)[base + i];
            ssim += pa.x; l1 += pa.y;
            if (p.d_render) {
                const float2 pb = reinterpret_cast<const float2*>(p.part_b)[base + i];
                dc[c] += pb.x; dm[c] += pb.y;
            }
        }
    }
    double vals[8] = {ssim, l1, dc[0], dc[1], dc[2], dm[0], dm[1], dm[2]};
    #pragma unroll
    for (int round = 0; round < 2; round++) {
        double* x = vals + 4 * round;
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
        }
        if (lane == 0) { s_acc[w][0] = x[0]; s_acc[w][1] = x[1]; s_acc[w][2] = x[2]; s_acc[w][3] = x[3]; }
        __syncthreads();
        if (tid == 0) {
            double t[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < THREADS / 32; i++) { t[0] += s_acc[i][0]; t[1] += s_acc[i][1]; t[2] += s_acc[i][2]; t[3] += s_acc[i][3]; }
            x[0] = t[0]; x[1] = t[1]; x[2] = t[2]; x[3] = t[3];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double n = 3.0 * (double)p.H * (double)p.W;
        const double l1m = vals[1] / n, ssm = vals[0] / n;
        p.loss[4 * v] = (float)l1m;
        p.loss[4 * v + 1] = (float)ssm;
        p.loss[4 * v + 2] = (float)((double)p.w_l1 * l1m + (double)p.w_ssim * (1.0 - ssm));
        p.loss[4 * v + 3] = 0.f;
        if (p.d_render) {
            for (int c = 0; c < 3; c++) {
                if (p.d_cam_c) p.d_cam_c[3 * v + c] = (float)vals[2 + c];
                if (p.d_cam_m) p.d_cam_m[3 * v + c] = (float)vals[5 + c];
            }
        }
    }
}

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" size_t t4d_image_loss_workspace_bytes(int32_t V, int32_t H, int32_t W)
{
    if (V < 1 || H < 1 || W < 1) return 0;
    const size_t nblk = (size_t)((W + TX - 1) / TX) * ((H + TY - 1) / TY);
    return al256((size_t)V * 9 * H * W * 4) + 2 * al256((size_t)V * 3 * nblk * 8);
}

extern "C" int t4d_image_loss(const T4dImageLoss* q, gs_stream_t stream)
{
    if (!q || q->V < 1 || q->H < 1 || q->W < 1) return GS_E_BAD_ARGS;
    if (!q->render || !q->target || !q->loss || !q->workspace) return GS_E_BAD_ARGS;
    if ((q->cam_m == NULL) != (q->cam_c == NULL)) return GS_E_BAD_ARGS;
    if (!q->dL_drender && (q->dL_dcam_m || q->dL_dcam_c)) return GS_E_BAD_ARGS;
    if (q->workspace_bytes < t4d_image_loss_workspace_bytes(q->V, q->H, q->W)) return GS_E_WORKSPACE_SMALL;
    if (((uintptr_t)q->workspace & 255u) != 0) return GS_E_BAD_ARGS;
    if ((long long)q->V * 3 > 65535) return GS_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    LossParams p;
    p.render = q->render; p.target = q->target; p.cam_m = q->cam_m; p.cam_c = q->cam_c;
    p.V = q->V; p.H = q->H; p.W = q->W;
    p.nbx = (q->W + TX - 1) / TX; p.nby = (q->H + TY - 1) / TY;
    const size_t nblk = (size_t)p.nbx * p.nby;
    char* ws = (char*)q->workspace;
    p.maps = (float*)ws;
    p.part_a = (float*)(ws + al256((size_t)q->V * 9 * q->H * q->W * 4));
    p.part_b = (float*)((char*)p.part_a + al256((size_t)q->V * 3 * nblk * 8));
    p.d_render = q->dL_drender; p.loss = q->loss; p.d_cam_m = q->dL_dcam_m; p.d_cam_c = q->dL_dcam_c;
    p.w_l1 = q->w_l1; p.w_ssim = q->w_ssim;
    // gaussian(11, 1.5) of external.py:71-73, normalised in fp32 like torch.Tensor(...) / sum
    {
        float g[WIN], sum = 0.f;
        for (int i = 0; i < WIN; i++) { g[i] = (float)exp(-(double)((i - RAD) * (i - RAD)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
        for (int i = 0; i < WIN; i++) {
            p.win.w[i] = g[i] / sum;
            uint32_t bits;
            memcpy(&bits, &p.win.w[i], 4);
            p.win.ww[i] = ((unsigned long long)bits << 32) | bits;
        }
    }
    const dim3 grid(p.nbx, p.nby, q->V * 3);
    ssim_fwd_kernel<<<grid, THREADS, 0, s>>>(p);
    if (p.d_render) ssim_bwd_kernel<<<grid, THREADS, 0, s>>>(p);
    loss_finalize_kernel<<<q->V, THREADS, 0, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : GS_E_CUDA;
}
