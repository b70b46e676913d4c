// gs_preprocess.cu -- forward preprocess (K1), tile scatter (K3) and markVisible (K10).
//
// COMPILED WITH --fmad=false: every expression here is IEEE fp32, left to right, unfused,
// with IEEE division / sqrtf, so that radius, tile rectangle, depth key -- and therefore every
// tile/bin index downstream -- are bit-identical to the CPU oracle.
//
// Replaces upstream preprocessCUDA / duplicateWithKeys of the un-vendored rasterizer that
// GaussianRasterizer.forward (reference call sites train.py:307,388) dispatches to.
// Design differences (B200-first, not a port):
//   * one launch covers all V views (thread = (view, Gaussian)); per-Gaussian outputs are packed
//     into ONE 48-byte record written with three 16-byte stores (geom[], layout in gs_common.cuh);
//   * no per-Gaussian prefix sum: tiles are counted with atomics here, scanned per TILE, and the
//     per-tile order is restored later by a (depth,index) sort, which is the same total order a
//     stable (tile|depth) radix sort produces -- so no host sync is needed to size buffers.
#include "gs_common.cuh"

namespace {

__device__ __constant__ float kC0 = 0.28209479177387814f;
__device__ __constant__ float kC1 = 0.4886025119029199f;
__device__ __constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                        -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                        0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                        -0.5900435899266435f};

struct Rect { int minx, miny, maxx, maxy; };
constexpr int SH_ROW = 52;          // floats per staged SH row (48 + 4 padding: conflict-free 16-byte reads at lane stride)

// tile rectangle of a splat: C truncation toward zero, then clamp to the grid (max exclusive)
__device__ __forceinline__ Rect tile_rect(float pix_x, float pix_y, float radius, int gx, int gy)
{
    Rect r;
    r.minx = (int)((pix_x - radius) / (float)GS_TILE);
    r.miny = (int)((pix_y - radius) / (float)GS_TILE);
    r.maxx = (int)((pix_x + radius + (float)(GS_TILE - 1)) / (float)GS_TILE);
    r.maxy = (int)((pix_y + radius + (float)(GS_TILE - 1)) / (float)GS_TILE);
    r.minx = min(gx, max(0, r.minx)); r.miny = min(gy, max(0, r.miny));
    r.maxx = min(gx, max(0, r.maxx)); r.maxy = min(gy, max(0, r.maxy));
    return r;
}

// The kernel is latency-bound (a thread's work is a chain of dependent global loads and ~700 unfused fp32 operations; at 24 views
// it issued at a fifth of the issue rate): the camera block of the (at most two) views a CTA touches is staged in shared memory,
// and every per-Gaussian input -- the 12 16-byte SH loads included -- is requested at the top, before the projection and
// covariance arithmetic that used to sit between a thread's loads, so ~20 loads per thread are in flight at once.  Registers
// (up to 128 at 128-thread CTAs) are cheaper here than exposed L2 round trips.  Results are unchanged (same operations, same order).
#ifndef GS_PRE_MINB
#define GS_PRE_MINB 6
#endif
__global__ void __launch_bounds__(128, GS_PRE_MINB) preprocess_kernel(const GsParams p, int32_t* __restrict__ radii)
{
    __shared__ float s_cam[2][GS_CAM_FLOATS];
    __shared__ __align__(16) float s_sh[4][32 * SH_ROW];
    // (view, Gaussian) index arithmetic in 32 bits: V * N < 2^31 is validated on the host, and a 64-bit division costs ~100 instructions
    const unsigned Nu = (unsigned)p.N, total = (unsigned)p.V * Nu;
    const unsigned gid0 = blockIdx.x * blockDim.x;
    const int v_first = (int)(gid0 / Nu);
    for (int k = threadIdx.x; k < 2 * GS_CAM_FLOATS; k += blockDim.x) {
        const int vv = v_first + k / GS_CAM_FLOATS;
        s_cam[k / GS_CAM_FLOATS][k % GS_CAM_FLOATS] = vv < p.V ? p.cams[(size_t)vv * GS_CAM_FLOATS + k % GS_CAM_FLOATS] : 0.f;
    }
    __syncthreads();
    const unsigned gid_raw = gid0 + threadIdx.x;
    const bool in_range = gid_raw < total;                                       // out-of-range threads still help stage the SH rows
    const unsigned gid = in_range ? gid_raw : 0u;
    const int v = (int)(gid / Nu), i = (int)(gid - (unsigned)v * Nu);
    const float* __restrict__ cam = s_cam[v - v_first];       // a 128-thread CTA spans at most two views (N >= 128) ...
    if (v - v_first > 1) cam = p.cams + (size_t)v * GS_CAM_FLOATS;   // ... or reads global memory for the rest (tiny N)
    const float* V = cam + GS_CAM_VIEW;
    const float* P = cam + GS_CAM_PROJ;
    if (in_range) radii[gid] = 0;

    // ---- all per-Gaussian inputs up front ----
    const float px = p.means3D[3 * i], py = p.means3D[3 * i + 1], pz = p.means3D[3 * i + 2];
    const float opac = p.opac[i];
    float in_s[3] = {0.f, 0.f, 0.f}, in_c3[6];
    float4 in_q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (p.cov3D) {
        #pragma unroll
        for (int k = 0; k < 6; k++) in_c3[k] = p.cov3D[6 * (size_t)i + k];
    } else {
        in_s[0] = p.scales[3 * i]; in_s[1] = p.scales[3 * i + 1]; in_s[2] = p.scales[3 * i + 2];
        in_q = reinterpret_cast<const float4*>(p.rots)[i];
    }
    float in_col[3] = {0.f, 0.f, 0.f};
    // SH rows: thread i owns 192 contiguous bytes, so per-thread 16-byte loads touch 32 different lines per instruction
    // (ncu: the L1 data pipe at 55 % with 384 wavefronts per warp just for these).  The 32 rows of a warp are contiguous
    // in memory (consecutive Gaussians), so the warp copies them with fully coalesced 16-byte loads into a padded
    // shared-memory tile (row stride 52 floats: conflict-free 16-byte reads) and every thread reads its own row from there.
    const bool sh_fast = p.shs != nullptr && p.M == 16 && ((reinterpret_cast<uintptr_t>(p.shs) & 15) == 0);
    if (sh_fast) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int nf4 = ((p.deg + 1) * (p.deg + 1) * 3 + 3) / 4;                  // 16-byte chunks of a row that are needed
        const unsigned g0 = gid0 + warp * 32;                                    // first (view, Gaussian) of this warp
        const int i0 = (int)(g0 % Nu);
        int r = lane / 12, q = lane - r * 12;                                    // chunk c = 32 k + lane = row r, 16-byte column q
        const bool wraps = i0 + 32 > p.N;                                         // the warp straddles a view boundary (or N < 32)
        #pragma unroll
        for (int k = 0; k < 12; k++) {
            int ir = i0 + r;
            if (wraps && ir >= p.N) ir %= p.N;
            if (q < nf4)
                *reinterpret_cast<float4*>(&s_sh[warp][r * SH_ROW + 4 * q]) = __ldg(reinterpret_cast<const float4*>(p.shs + (size_t)ir * 48) + q);
            q += 8; r += 2;                                                       // c += 32 = 2 rows + 8 columns
            if (q >= 12) { q -= 12; r++; }
        }
        __syncwarp();
    } else if (!p.shs) {
        in_col[0] = p.colors[3 * i]; in_col[1] = p.colors[3 * i + 1]; in_col[2] = p.colors[3 * i + 2];
    }
    if (!in_range) return;
    const float tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
    const float ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
    const float tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
    if (!(tz > GS_NEAR_CULL)) return;

    const float hx = P[0] * px + P[4] * py + P[8] * pz + P[12];
    const float hy = P[1] * px + P[5] * py + P[9] * pz + P[13];
    const float hw = P[3] * px + P[7] * py + P[11] * pz + P[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ppx = hx * pw, ppy = hy * pw;

    float c3[6];
    if (p.cov3D) {
        #pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = in_c3[k];
    } else {
        const float sx = p.mod * in_s[0], sy = p.mod * in_s[1], sz = p.mod * in_s[2];
        const float4 q = in_q;
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
        const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
        const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
        const float A00 = R00 * sx, A01 = R01 * sy, A02 = R02 * sz;
        const float A10 = R10 * sx, A11 = R11 * sy, A12 = R12 * sz;
        const float A20 = R20 * sx, A21 = R21 * sy, A22 = R22 * sz;
        c3[0] = A00 * A00 + A01 * A01 + A02 * A02;
        c3[1] = A00 * A10 + A01 * A11 + A02 * A12;
        c3[2] = A00 * A20 + A01 * A21 + A02 * A22;
        c3[3] = A10 * A10 + A11 * A11 + A12 * A12;
        c3[4] = A10 * A20 + A11 * A21 + A12 * A22;
        c3[5] = A20 * A20 + A21 * A21 + A22 * A22;
    }

    const float tanx = cam[GS_CAM_TANFOVX], tany = cam[GS_CAM_TANFOVY];
    const float fx = (float)p.W / (2.0f * tanx), fy = (float)p.H / (2.0f * tany);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float txtz = tx / tz, tytz = ty / tz;
    const float cx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    const float cy = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float J00 = fx / tz, J02 = -(fx * cx) / (tz * tz);
    const float J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
    float T0[3], T1[3];
    #pragma unroll
    for (int j = 0; j < 3; j++) {
        T0[j] = J00 * V[4 * j + 0] + J02 * V[4 * j + 2];
        T1[j] = J11 * V[4 * j + 1] + J12 * V[4 * j + 2];
    }
    const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float X0[3], X1[3];
    #pragma unroll
    for (int j = 0; j < 3; j++) {
        X0[j] = T0[0] * S[0][j] + T0[1] * S[1][j] + T0[2] * S[2][j];
        X1[j] = T1[0] * S[0][j] + T1[1] * S[1][j] + T1[2] * S[2][j];
    }
    const float a = (X0[0] * T0[0] + X0[1] * T0[1] + X0[2] * T0[2]) + GS_LOWPASS;
    const float b = X0[0] * T1[0] + X0[1] * T1[1] + X0[2] * T1[2];
    const float c = (X1[0] * T1[0] + X1[1] * T1[1] + X1[2] * T1[2]) + GS_LOWPASS;

    const float det = a * c - b * b;
    if (!(det != 0.0f)) return;
    const float det_inv = 1.0f / det;
    const float cA = c * det_inv, cB = -b * det_inv, cC = a * det_inv;
    const float mid = 0.5f * (a + c);
    const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float l1 = mid + sq, l2 = mid - sq;
    const float radius = ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
    if (!(radius < 1.0e9f)) return;
    const float pix_x = ((ppx + 1.0f) * (float)p.W - 1.0f) * 0.5f;
    const float pix_y = ((ppy + 1.0f) * (float)p.H - 1.0f) * 0.5f;
    if (!(fabsf(pix_x) < 1.0e9f) || !(fabsf(pix_y) < 1.0e9f)) return;

    const Rect rc = tile_rect(pix_x, pix_y, radius, p.tiles_x, p.tiles_y);
    if ((rc.maxx - rc.minx) * (rc.maxy - rc.miny) == 0) return;

    float rgb[3];
    unsigned clampbits = 0;
    if (p.shs) {
        float sh[48];
        const int nf = (p.deg + 1) * (p.deg + 1) * 3;
        if (sh_fast) {
            const float4* __restrict__ row = reinterpret_cast<const float4*>(&s_sh[threadIdx.x >> 5][(threadIdx.x & 31) * SH_ROW]);
            #pragma unroll
            for (int q = 0; q < 12; q++) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q * 4 < nf) t = row[q];
                sh[4 * q] = t.x; sh[4 * q + 1] = t.y; sh[4 * q + 2] = t.z; sh[4 * q + 3] = t.w;
            }
        } else {
            const float* __restrict__ shg = p.shs + (size_t)i * p.M * 3;
            #pragma unroll
            for (int k = 0; k < 48; k++) sh[k] = k < nf ? __ldg(shg + k) : 0.f;
        }
        const float dx = px - cam[GS_CAM_CAMPOS], dy = py - cam[GS_CAM_CAMPOS + 1], dz = pz - cam[GS_CAM_CAMPOS + 2];
        const float len = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx / len, y = dy / len, z = dz / len;
        #pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float res = kC0 * sh[0 * 3 + ch];
            if (p.deg > 0) {
                res = res - kC1 * y * sh[1 * 3 + ch] + kC1 * z * sh[2 * 3 + ch] - kC1 * x * sh[3 * 3 + ch];
                if (p.deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    res = res + kC2[0] * xy * sh[4 * 3 + ch] + kC2[1] * yz * sh[5 * 3 + ch]
                              + kC2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + ch]
                              + kC2[3] * xz * sh[7 * 3 + ch] + kC2[4] * (xx - yy) * sh[8 * 3 + ch];
                    if (p.deg > 2) {
                        res = res + kC3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + ch]
                                  + kC3[1] * xy * z * sh[10 * 3 + ch]
                                  + kC3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + ch]
                                  + kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + ch]
                                  + kC3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + ch]
                                  + kC3[5] * z * (xx - yy) * sh[14 * 3 + ch]
                                  + kC3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + ch];
                    }
                }
            }
            res += 0.5f;
            if (res < 0.0f) { clampbits |= 1u << ch; res = 0.0f; }
            rgb[ch] = res;
        }
    } else {
        rgb[0] = in_col[0]; rgb[1] = in_col[1]; rgb[2] = in_col[2];
    }

    radii[gid] = (int)radius;
    p.clamped[gid] = (uint8_t)clampbits;
    float4* g = p.geom + (size_t)gid * 3;
    g[0] = make_float4(pix_x, pix_y, cA, cB);
    // thr: the blend kernels skip a pixel without evaluating exp() when power < thr.  alpha >= 1/255 needs
    // power >= -ln(255*opacity); the 1e-3 margin (0.1 % in alpha) dwarfs any fp32 rounding of power or exp,
    // so the prefilter can only pass extra pixels (which then fail the exact alpha test), never drop one.
    const float thr = opac > 0.0f ? fmaxf(-logf(255.0f * opac) - 1.0e-3f, -1.0e20f) : __int_as_float(0x7f800000);   // >= -1e20: see PARKED_Y in gs_blend.cu
    g[1] = make_float4(cC, opac, tz, thr);
    g[2] = make_float4(rgb[0], rgb[1], rgb[2], __int_as_float(i));

    // count this splat in every tile of its rectangle
    uint32_t* cnt = p.tile_count + (size_t)v * p.tiles;
    for (int y = rc.miny; y < rc.maxy; y++)
        for (int x = rc.minx; x < rc.maxx; x++) atomicAdd(cnt + y * p.tiles_x + x, 1u);
}

// K3: write (depth_bits<<32 | id) into the tile's segment.  Slot order inside a tile is arbitrary
// (atomic cursor); the per-tile sort restores (depth, id) order, a total order on distinct keys.
__global__ void __launch_bounds__(256) scatter_kernel(const GsParams p, const int32_t* __restrict__ radii)
{
    gs_pdl_wait();                                                                // tile_start comes from the scan kernel
    gs_pdl_trigger();
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;                  // V * N < 2^31 (validated on the host)
    if (gid >= (unsigned)p.V * (unsigned)p.N) return;
    const int rad = radii[gid];
    if (rad <= 0) return;
    const int v = (int)(gid / (unsigned)p.N), i = (int)(gid - (unsigned)v * (unsigned)p.N);
    const float4 g0 = p.geom[(size_t)gid * 3], g1 = p.geom[(size_t)gid * 3 + 1];
    const Rect rc = tile_rect(g0.x, g0.y, (float)rad, p.tiles_x, p.tiles_y);
    const unsigned long long pair = ((unsigned long long)__float_as_uint(g1.z) << 32) | (unsigned)i;
    const size_t tbase = (size_t)v * p.tiles;
    for (int y = rc.miny; y < rc.maxy; y++)
        for (int x = rc.minx; x < rc.maxx; x++) {
            const size_t t = tbase + y * p.tiles_x + x;
            const unsigned long long pos = (unsigned long long)p.tile_start[t] + atomicAdd(p.tile_fill + t, 1u);
            if (pos < (unsigned long long)p.cap) p.pairs[pos] = pair;
        }
}

__global__ void mark_visible_kernel(int N, const float* __restrict__ means3D, const float* __restrict__ cam,
                                    uint8_t* __restrict__ visible)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* V = cam + GS_CAM_VIEW;
    const float tz = V[2] * means3D[3 * i] + V[6] * means3D[3 * i + 1] + V[10] * means3D[3 * i + 2] + V[14];
    visible[i] = tz > GS_NEAR_CULL;
}

}  // namespace

void gs_launch_preprocess(const GsParams& p, int32_t* radii, cudaStream_t s)
{
    const long long n = (long long)p.V * p.N;
    if (n == 0) return;
    preprocess_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, radii);
}

void gs_launch_scatter(const GsParams& p, const int32_t* radii, cudaStream_t s)
{
    const long long n = (long long)p.V * p.N;
    if (n == 0) return;
    gs_launch_dependent(scatter_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, p, radii);
}

void gs_launch_mark_visible(int N, const float* means3D, const float* cam, uint8_t* visible, cudaStream_t s)
{
    if (N == 0) return;
    mark_visible_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, means3D, cam, visible);
}
