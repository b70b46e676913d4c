// gs_blend.cu -- forward (K6) and backward (K7) alpha-blend kernels, sm_100a.
//
// Replaces upstream renderCUDA forward/backward of the un-vendored rasterizer behind
// GaussianRasterizer (reference call sites train.py:307,388 forward; train.py:667,738 backward).
// Same per-pixel arithmetic (alpha = min(0.99, o*exp(power)), skip alpha<1/255,
// stop at T(1-alpha)<1e-4, straight-through cap in backward), different machine mapping:
//   * persistent CTAs walk (view, tile) pairs; a CTA owns one 16x16 binning tile at a time, each of
//     its 8 warps an 8x4-pixel block (32-byte row segments per plane);
//   * the tile's depth-sorted 48-byte records are contiguous in HBM (gs_binning.cu), so a chunk of
//     128 records is ONE cp.async.bulk (SASS UBLKCP) into shared memory, double-buffered on two
//     mbarriers: no per-thread gather, no register staging;
//   * forward: a warp leaves the chunk loop as soon as all 32 of its pixels are saturated;
//   * backward: starts at the tile's deepest contributor (block-max n_contrib), skips records no
//     pixel of the warp touches (ballot), reduces the 10 per-Gaussian partials across the warp with a
//     halving butterfly (16 shuffles instead of 50), accumulates the 8 warps in shared memory and
//     issues three 16-byte vector REDs per (tile, Gaussian) instead of 10 scalar atomics per pixel.
#include "gs_common.cuh"

namespace {

constexpr int BLEND_THREADS = 256;
constexpr int CHUNK = 128;                    // records per bulk copy (6 KB)
constexpr uint32_t REC_BYTES = 48;

// Identical bits in forward and backward: the backward pass must re-derive exactly the alpha the
// forward pass blended, so the contraction pattern is pinned with explicit intrinsics.
__device__ __forceinline__ float splat_power(float cA, float cB, float cC, float dx, float dy)
{
    const float t1 = __fmul_rn(__fmul_rn(cA, dx), dx);
    const float t2 = __fmaf_rn(__fmul_rn(cC, dy), dy, t1);
    const float t3 = __fmul_rn(__fmul_rn(cB, dx), dy);
    return __fmaf_rn(-0.5f, t2, -t3);
}
__device__ __forceinline__ float splat_alpha(float opacity, float G) { return fminf(GS_ALPHA_CAP, __fmul_rn(opacity, G)); }
// exp flavour: GS_PRECISE_EXP=1 -> expf (~1 ulp, what upstream's exp() compiles to);
//              default          -> __expf (ex2.approx of x*log2e; ~2+|1.17x| ulp, 2 instructions)
#ifndef GS_PRECISE_EXP
#define GS_PRECISE_EXP 0
#endif
__device__ __forceinline__ float splat_exp(float power)
{
#if GS_PRECISE_EXP
    return expf(power);
#else
    return __expf(power);
#endif
}

struct TileCtx {
    int v, px, py;
    bool inside;
    unsigned long long start;
    int n;
};

__device__ __forceinline__ TileCtx tile_ctx(const GsParams& p, long long tg, int lx, int ly)
{
    TileCtx c;
    c.v = (int)(tg / p.tiles);
    const int t = (int)(tg - (long long)c.v * p.tiles);
    c.px = (t % p.tiles_x) * GS_TILE + lx;
    c.py = (t / p.tiles_x) * GS_TILE + ly;
    c.inside = c.px < p.W && c.py < p.H;
    unsigned long long s = p.tile_start[tg], e = p.tile_start[tg + 1];
    if (e > (unsigned long long)p.cap) e = (unsigned long long)p.cap;
    c.start = s;
    c.n = e > s ? (int)(e - s) : 0;
    return c;
}

__global__ void __launch_bounds__(BLEND_THREADS)
blend_fwd_kernel(const GsParams p, float* __restrict__ out_color, float* __restrict__ out_depth,
                 float* __restrict__ out_alpha)
{
    __shared__ __align__(128) float4 s_rec[2][CHUNK * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phases = 0u;                       // bit b = parity to wait for on s_bar[b]
    const size_t HW = (size_t)p.H * p.W;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);

    for (long long tg = blockIdx.x; tg < p.total_tiles; tg += gridDim.x) {
        const TileCtx tc = tile_ctx(p, tg, lx, ly);
        const float pxf = (float)tc.px, pyf = (float)tc.py;
        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, A = 0.f;
        uint32_t last = 0;
        bool done = !tc.inside;
        const int nchunks = (tc.n + CHUNK - 1) / CHUNK;
        const float4* __restrict__ src = p.sorted_rec + tc.start * 3;
        if (nchunks > 0 && tid == 0) {
            const uint32_t bytes = (uint32_t)min(tc.n, CHUNK) * REC_BYTES;
            mbar_expect_tx(&s_bar[0], bytes);
            bulk_g2s(s_rec[0], src, bytes, &s_bar[0]);
        }
        for (int c = 0; c < nchunks; c++) {
            const int cur = c & 1;
            const bool have_next = c + 1 < nchunks;
            if (have_next && tid == 0) {
                const uint32_t bytes = (uint32_t)min(tc.n - (c + 1) * CHUNK, CHUNK) * REC_BYTES;
                mbar_expect_tx(&s_bar[cur ^ 1], bytes);
                bulk_g2s(s_rec[cur ^ 1], src + (size_t)(c + 1) * CHUNK * 3, bytes, &s_bar[cur ^ 1]);
            }
            mbar_wait(&s_bar[cur], (phases >> cur) & 1u);
            phases ^= 1u << cur;
            const int cnt = min(tc.n - c * CHUNK, CHUNK);
            if (!__all_sync(0xffffffffu, done)) {
                const float4* __restrict__ rec = s_rec[cur];
                for (int j = 0; j < cnt && !done; j++) {
                    const float4 r0 = rec[j * 3], r1 = rec[j * 3 + 1];
                    const float dx = r0.x - pxf, dy = r0.y - pyf;
                    const float power = splat_power(r0.z, r0.w, r1.x, dx, dy);
                    if (power > 0.0f) continue;
                    const float alpha = splat_alpha(r1.y, splat_exp(power));
                    if (alpha < GS_ALPHA_MIN) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < GS_T_MIN) { done = true; break; }
                    const float4 r2 = rec[j * 3 + 2];
                    const float w = alpha * T;
                    C0 += r2.x * w; C1 += r2.y * w; C2 += r2.z * w;
                    D += r1.z * w;
                    A += w;
                    T = test_T;
                    last = (uint32_t)(c * CHUNK + j + 1);
                }
            }
            const int ndone = __syncthreads_count(done);
            if (ndone == BLEND_THREADS) {
                if (have_next) { mbar_wait(&s_bar[cur ^ 1], (phases >> (cur ^ 1)) & 1u); phases ^= 1u << (cur ^ 1); }   // drain prefetch
                break;
            }
        }
        if (tc.inside) {
            const float* __restrict__ bg = p.cams + (size_t)tc.v * GS_CAM_FLOATS + GS_CAM_BG;
            const size_t pix = (size_t)tc.py * p.W + tc.px;
            const size_t vb = (size_t)tc.v * HW;
            p.final_T[vb + pix] = T;
            p.n_contrib[vb + pix] = last;
            out_color[vb * 3 + pix] = C0 + T * bg[0];
            out_color[vb * 3 + HW + pix] = C1 + T * bg[1];
            out_color[vb * 3 + 2 * HW + pix] = C2 + T * bg[2];
            out_depth[vb + pix] = D;
            out_alpha[vb + pix] = A;
        }
    }
}

// slot s of the butterfly -> float index inside the 12-float grad2d record
__device__ __forceinline__ int slot_to_float(int s) { return s < 7 ? s : s + 1; }

__global__ void __launch_bounds__(BLEND_THREADS)
blend_bwd_kernel(const GsParams p, const float* __restrict__ g_color, const float* __restrict__ g_depth,
                 const float* __restrict__ g_alpha)
{
    __shared__ __align__(128) float4 s_rec[2][CHUNK * 3];
    __shared__ __align__(16) float s_acc[CHUNK * GS_REC_FLOATS];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_max[BLEND_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phases = 0u;                       // bit b = parity to wait for on s_bar[b]
    const size_t HW = (size_t)p.H * p.W;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const bool hi4 = lane & 16, hi3 = lane & 8, hi2 = lane & 4, hi1 = lane & 2;
    const int my_slot = lane >> 1;

    for (long long tg = blockIdx.x; tg < p.total_tiles; tg += gridDim.x) {
        const TileCtx tc = tile_ctx(p, tg, lx, ly);
        if (tc.n == 0) continue;                               // uniform per CTA
        const float pxf = (float)tc.px, pyf = (float)tc.py;
        const size_t vb = (size_t)tc.v * HW;
        const size_t pix = (size_t)tc.py * p.W + tc.px;
        float T_final = 0.f, gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gd = 0.f, ga = 0.f;
        uint32_t last = 0;
        if (tc.inside) {
            T_final = p.final_T[vb + pix];
            last = p.n_contrib[vb + pix];
            gc0 = g_color[vb * 3 + pix]; gc1 = g_color[vb * 3 + HW + pix]; gc2 = g_color[vb * 3 + 2 * HW + pix];
            if (g_depth) gd = g_depth[vb + pix];
            if (g_alpha) ga = g_alpha[vb + pix];
        }
        uint32_t m = last;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_max[warp] = m;
        __syncthreads();
        uint32_t nmax = 0;
        #pragma unroll
        for (int w = 0; w < BLEND_THREADS / 32; w++) nmax = max(nmax, s_max[w]);
        nmax = min(nmax, (uint32_t)tc.n);
        __syncthreads();                                       // s_max is rewritten by the next tile
        if (nmax == 0) continue;                               // uniform

        const float* __restrict__ bg = p.cams + (size_t)tc.v * GS_CAM_FLOATS + GS_CAM_BG;
        const float bg_dot = bg[0] * gc0 + bg[1] * gc1 + bg[2] * gc2;
        float T = T_final, last_alpha = 0.f;
        float rc0 = 0.f, rc1 = 0.f, rc2 = 0.f, rd = 0.f, ra = 0.f;      // "colour behind" recursions
        float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;

        const int nchunks = ((int)nmax + CHUNK - 1) / CHUNK;
        const float4* __restrict__ src = p.sorted_rec + tc.start * 3;
        float4* __restrict__ gbase = p.grad2d + (size_t)tc.v * p.N * 3;
        if (tid == 0) {
            const int c = nchunks - 1;
            const uint32_t bytes = (uint32_t)min((int)nmax - c * CHUNK, CHUNK) * REC_BYTES;
            mbar_expect_tx(&s_bar[0], bytes);
            bulk_g2s(s_rec[0], src + (size_t)c * CHUNK * 3, bytes, &s_bar[0]);
        }
        for (int k = 0; k < nchunks; k++) {
            const int c = nchunks - 1 - k, cur = k & 1;
            if (c > 0 && tid == 0) {
                const uint32_t bytes = (uint32_t)CHUNK * REC_BYTES;          // every earlier chunk is full
                mbar_expect_tx(&s_bar[cur ^ 1], bytes);
                bulk_g2s(s_rec[cur ^ 1], src + (size_t)(c - 1) * CHUNK * 3, bytes, &s_bar[cur ^ 1]);
            }
            for (int t = tid; t < CHUNK * 3; t += BLEND_THREADS)
                reinterpret_cast<float4*>(s_acc)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
            mbar_wait(&s_bar[cur], (phases >> cur) & 1u);
            phases ^= 1u << cur;
            const int cnt = min((int)nmax - c * CHUNK, CHUNK);
            const float4* __restrict__ rec = s_rec[cur];
            for (int j = cnt - 1; j >= 0; j--) {
                const uint32_t idx = (uint32_t)(c * CHUNK + j);
                const float4 r0 = rec[j * 3], r1 = rec[j * 3 + 1];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                const float power = splat_power(r0.z, r0.w, r1.x, dx, dy);
                const float G = splat_exp(power);
                const float alpha = splat_alpha(r1.y, G);
                const bool contrib = (idx < last) && (power <= 0.0f) && (alpha >= GS_ALPHA_MIN);
                if (!__any_sync(0xffffffffu, contrib)) continue;
                float r[16];
                #pragma unroll
                for (int s = 0; s < 16; s++) r[s] = 0.f;
                if (contrib) {
                    const float4 r2 = rec[j * 3 + 2];
                    T = T / (1.0f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha;
                    rc0 = last_alpha * lc0 + (1.f - last_alpha) * rc0; lc0 = r2.x;
                    rc1 = last_alpha * lc1 + (1.f - last_alpha) * rc1; lc1 = r2.y;
                    rc2 = last_alpha * lc2 + (1.f - last_alpha) * rc2; lc2 = r2.z;
                    dL_dalpha = (r2.x - rc0) * gc0 + (r2.y - rc1) * gc1 + (r2.z - rc2) * gc2;
                    rd = last_alpha * ld + (1.f - last_alpha) * rd; ld = r1.z;
                    dL_dalpha += (r1.z - rd) * gd;
                    ra = last_alpha + (1.f - last_alpha) * ra;
                    dL_dalpha += (1.f - ra) * ga;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = r1.y * dL_dalpha;          // straight-through the 0.99 cap
                    const float gdx = G * dx, gdy = G * dy;
                    r[0] = dL_dG * (-gdx * r0.z - gdy * r0.w);     // d/dpix.x
                    r[1] = dL_dG * (-gdy * r1.x - gdx * r0.w);     // d/dpix.y
                    r[2] = -0.5f * gdx * dx * dL_dG;               // d/dconA
                    r[3] = -gdx * dy * dL_dG;                      // d/dconB (true, not halved)
                    r[4] = -0.5f * gdy * dy * dL_dG;               // d/dconC
                    r[5] = G * dL_dalpha;                          // d/dopacity
                    r[6] = w * gd;                                 // d/ddepth
                    r[7] = w * gc0; r[8] = w * gc1; r[9] = w * gc2;
                }
                // halving butterfly: after it, lane pair (2s, 2s+1) holds the warp total of slot s
                #pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float send = hi4 ? r[i] : r[8 + i];
                    const float keep = hi4 ? r[8 + i] : r[i];
                    r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                #pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float send = hi3 ? r[i] : r[4 + i];
                    const float keep = hi3 ? r[4 + i] : r[i];
                    r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                #pragma unroll
                for (int i = 0; i < 2; i++) {
                    const float send = hi2 ? r[i] : r[2 + i];
                    const float keep = hi2 ? r[2 + i] : r[i];
                    r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                {
                    const float send = hi1 ? r[0] : r[1];
                    const float keep = hi1 ? r[1] : r[0];
                    r[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
                r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
                if (!(lane & 1) && my_slot < 10) atomicAdd(&s_acc[j * GS_REC_FLOATS + slot_to_float(my_slot)], r[0]);
            }
            __syncthreads();
            for (int t = tid; t < cnt * 3; t += BLEND_THREADS) {
                const float4 a = reinterpret_cast<const float4*>(s_acc)[t];
                if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) {
                    const int j = t / 3, part = t - j * 3;
                    const int id = __float_as_int(rec[j * 3 + 1].w);
                    red_add_v4(gbase + (size_t)id * 3 + part, a);
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace

static unsigned persistent_grid(long long total_tiles, int num_sms, int ctas_per_sm)
{
    long long blocks = total_tiles;
    const long long maxb = (long long)num_sms * ctas_per_sm;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

void gs_launch_blend_fwd(const GsParams& p, float* color, float* depth, float* alpha, int num_sms, cudaStream_t s)
{
    blend_fwd_kernel<<<persistent_grid(p.total_tiles, num_sms, 8), BLEND_THREADS, 0, s>>>(p, color, depth, alpha);
}

void gs_launch_blend_bwd(const GsParams& p, const float* g_color, const float* g_depth, const float* g_alpha,
                         int num_sms, cudaStream_t s)
{
    blend_bwd_kernel<<<persistent_grid(p.total_tiles, num_sms, 6), BLEND_THREADS, 0, s>>>(p, g_color, g_depth, g_alpha);
}
