// f3d_render.cu -- face3d `render_colors` triangle rasterizer for the 8K texture bake, sm_100a.
//
// Replaces _render_colors_core (reference face3d/mesh/cython/mesh_core.cpp:169-234, with its
// helpers isPointInTri :23-50 and get_point_weight :53-82), reached from helpers.py:956 through
// face3d/mesh/render.py:52-86.  COMPILED WITH --fmad=false: the inside test and barycentric
// weights repeat the reference's fp32 operation sequence unfused so coverage is bit-identical.
//
// The reference is a serial painter: triangles in index order, a pixel is overwritten when
// p_depth > depth_buffer (strict).  That is order-free equivalent to: per pixel, the winner is
// the triangle maximising (p_depth, -index) among those accepting the pixel, drawn iff its
// p_depth > the initial depth.  Two passes:
//   stage 0/1 (triangle-parallel) 32-bit atomicMax of orderable(p_depth), then of ~index among the
//           triangles at that depth, per accepted bbox pixel; a warp takes 32 triangles, small boxes are
//           walked by their own lane, large boxes cooperatively by the whole warp; stage 0 is skipped
//           when every vertex z is 0 (the texture bake);
//   shade   (pixel-parallel, coalesced) decode the winner, recompute its weights (same bits),
//           depth-test, write c channels (+ depth) -- optionally filling the background and / or
//           converting to uint8 in the same pass.
// The image is resolved in bands of rows whose key planes stay in L2.  Roofline: HBM, h*w*4c bytes written
// (+ the caller's depth plane read and written when the in-place depth contract is used).
#include "gs_common.cuh"

namespace {

struct Tri {
    float x0, y0, z0, x1, y1, z1, x2, y2, z2;
    int xmin, xmax, ymin, ymax;
};

// reference get_point_weight / isPointInTri arithmetic (mesh_core.cpp:23-82), op for op.  The part that does not depend on the
// pixel (edge vectors, their dot products, the inverse Gram determinant) is computed once per triangle (TriSetup); the per-pixel
// part repeats the reference's remaining operations in the reference's order, so the bits are those of the fused function.
struct TriSetup { float v0x, v0y, v1x, v1y, dot00, dot01, dot11, inv; };

__device__ __forceinline__ TriSetup tri_setup(const Tri& t)
{
    TriSetup s;
    s.v0x = t.x2 - t.x0; s.v0y = t.y2 - t.y0;
    s.v1x = t.x1 - t.x0; s.v1y = t.y1 - t.y0;
    s.dot00 = s.v0x * s.v0x + s.v0y * s.v0y;
    s.dot01 = s.v0x * s.v1x + s.v0y * s.v1y;
    s.dot11 = s.v1x * s.v1x + s.v1y * s.v1y;
    const float den = s.dot00 * s.dot11 - s.dot01 * s.dot01;
    s.inv = (den == 0.0f) ? 0.0f : 1.0f / den;
    return s;
}

__device__ __forceinline__ void bary_px(float px, float py, const Tri& t, const TriSetup& s, float& w0, float& w1, float& w2, bool& inside)
{
    const float v2x = px - t.x0, v2y = py - t.y0;
    const float dot02 = s.v0x * v2x + s.v0y * v2y;
    const float dot12 = s.v1x * v2x + s.v1y * v2y;
    const float u = (s.dot11 * dot02 - s.dot01 * dot12) * s.inv;
    const float v = (s.dot00 * dot12 - s.dot01 * dot02) * s.inv;
    inside = (u >= 0.0f) && (v >= 0.0f) && (u + v < 1.0f);
    w0 = 1.0f - u - v; w1 = v; w2 = u;
}

__device__ __forceinline__ void bary(float px, float py, const Tri& t, float& w0, float& w1, float& w2, bool& inside)
{
    bary_px(px, py, t, tri_setup(t), w0, w1, w2, inside);
}

__device__ __forceinline__ bool accepts(int x, int y, int h, int w, bool inside)
{
    return x < 2 || x > w - 3 || y < 2 || y > h - 3 || inside;                // mesh_core.cpp:211 (integer pixel coordinates)
}

// order-preserving map fp32 -> u32 (NaN never reaches it); +0 and -0 compare equal in the reference's strict '>' test
__device__ __forceinline__ unsigned orderable(float depth)
{
    unsigned b = __float_as_uint(depth);
    if (depth == 0.0f) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ int clamp_to_int(float f)
{
    f = fminf(fmaxf(f, -1.0e9f), 1.0e9f);                    // NaN -> -1e9 (fmaxf returns the non-NaN)
    return (int)f;
}

__device__ __forceinline__ bool load_tri(const float* __restrict__ vertices, const int* __restrict__ triangles,
                                         int i, int nver, int h, int w, int y_lo, int y_hi, Tri& t)
{
    const int i0 = triangles[3 * i], i1 = triangles[3 * i + 1], i2 = triangles[3 * i + 2];
    if ((unsigned)i0 >= (unsigned)nver || (unsigned)i1 >= (unsigned)nver || (unsigned)i2 >= (unsigned)nver) return false;
    t.x0 = vertices[3 * i0]; t.y0 = vertices[3 * i0 + 1]; t.z0 = vertices[3 * i0 + 2];
    t.x1 = vertices[3 * i1]; t.y1 = vertices[3 * i1 + 1]; t.z1 = vertices[3 * i1 + 2];
    t.x2 = vertices[3 * i2]; t.y2 = vertices[3 * i2 + 1]; t.z2 = vertices[3 * i2 + 2];
    t.xmin = max(clamp_to_int(ceilf(fminf(t.x0, fminf(t.x1, t.x2)))), 0);
    t.xmax = min(clamp_to_int(floorf(fmaxf(t.x0, fmaxf(t.x1, t.x2)))), w - 1);
    t.ymin = max(max(clamp_to_int(ceilf(fminf(t.y0, fminf(t.y1, t.y2)))), 0), y_lo);          // clipped to the band of rows
    t.ymax = min(min(clamp_to_int(floorf(fmaxf(t.y0, fmaxf(t.y1, t.y2)))), h - 1), y_hi);     // this launch resolves
    return !(t.xmax < t.xmin || t.ymax < t.ymin);
}

// Per-pixel winner = the accepting triangle maximising (p_depth, -index).  With 32-bit atomics only:
//   STAGE 0  dmax[pix] = max orderable(p_depth)                     (skipped when every vertex z is 0: all depths are +-0)
//   STAGE 1  imax[pix] = max (~index) over the accepting triangles whose depth equals dmax[pix]
// (the r01 kernel used one 64-bit atomicMax on depth << 32 | ~index: twice the key traffic, half the atomic rate)
template <int STAGE>
__device__ __forceinline__ void test_pixel(const Tri& t, const TriSetup& ts, int idx, int x, int y, int h, int w, int y_lo, bool flat,
                                           unsigned* __restrict__ dmax, unsigned* __restrict__ imax)
{
    float w0, w1, w2; bool inside;
    bary_px((float)x, (float)y, t, ts, w0, w1, w2, inside);
    if (!accepts(x, y, h, w, inside)) return;
    const float d = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
    if (d != d) return;                                      // NaN never passes '>' in the reference
    const size_t pix = (size_t)(y - y_lo) * w + x;
    if (STAGE == 0) atomicMax(dmax + pix, orderable(d));
    else if (flat || dmax[pix] == orderable(d)) atomicMax(imax + pix, 0xffffffffu - (unsigned)idx);
}

#ifndef F3D_SPLIT_MAX_LOG2
#define F3D_SPLIT_MAX_LOG2 4
#endif
constexpr int SMALL_BOX = 48;   // bbox pixels a single lane walks by itself

// The texture bake's case -- every vertex z is 0 and the bbox keeps clear of the 2-pixel image border -- needs no depth and no border
// rule: a pixel accepts the triangle iff it is inside (then u, v are finite, the interpolated depth is +-0, never NaN), and the winner is
// the lowest index.  inside_px is bary_px's inside test with the products that depend on one coordinate only passed in: (v0x v2x),
// (v1x v2x) are computed once per column by the callers, so a pixel costs 17 fp32 operations instead of ~50 instructions of the general
// test_pixel.  Same operations on the same operands in the same order as the reference (mesh_core.cpp:23-50): same bits.
__device__ __forceinline__ bool inside_px(const TriSetup& s, float ax, float bx, float v2y)
{
    const float dot02 = ax + s.v0y * v2y;
    const float dot12 = bx + s.v1y * v2y;
    const float u = (s.dot11 * dot02 - s.dot01 * dot12) * s.inv;
    const float v = (s.dot00 * dot12 - s.dot01 * dot02) * s.inv;
    return (u >= 0.0f) && (v >= 0.0f) && (u + v < 1.0f);
}
__device__ __forceinline__ bool clear_of_border(int xmin, int xmax, int ymin, int ymax, int h, int w)
{
    return xmin >= 2 && xmax <= w - 3 && ymin >= 2 && ymax <= h - 3;
}

template <int STAGE>
__global__ void __launch_bounds__(256)
f3d_tri_kernel(const float* __restrict__ vertices, const int* __restrict__ triangles, int nver, int ntri, int h, int w,
               int y_lo, int y_hi, const int* __restrict__ nonflat, unsigned* __restrict__ dmax, unsigned* __restrict__ imax, int split_log2)
{
    const bool flat = *nonflat == 0;
    if (STAGE == 0 && flat) return;
    const int lane = threadIdx.x & 31;
    // 2^split_log2 warps share a group of 32 triangles: all of them load the group (one triangle per lane), each walks the boxes whose
    // lane index is congruent to its part.  With one warp per group a 120 k-triangle bake was a single wave of 3 752 warps, each walking
    // its 32 boxes (~1 100 pixels each) one after the other on its own dependent-latency clock: 0.38 ms at an IPC of 0.4, and as long
    // as the slowest warp (the rows of triangles along the image border take the general test).  The host picks the split from the
    // mean box size; meshes of pixel-sized triangles (one lane per box, 32 boxes per warp at once) keep one warp per group.
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int split = 1 << split_log2, part = (int)(warp_global & (split - 1));
    const long long ngroups = (((long long)gridDim.x * blockDim.x) >> 5) >> split_log2;
    for (long long base = (warp_global >> split_log2) * 32; base < ntri; base += ngroups * 32) {
        const int i = (int)(base + lane);
        Tri t;
        bool ok = false;
        if (i < ntri) ok = load_tri(vertices, triangles, i, nver, h, w, y_lo, y_hi, t);
        int bw = 0, area = 0;
        if (ok) { bw = t.xmax - t.xmin + 1; const long long a = (long long)bw * (t.ymax - t.ymin + 1); area = a > 0x7fffffffLL ? 0x7fffffff : (int)a; }
        const bool fast = STAGE == 1 && flat && ok && clear_of_border(t.xmin, t.xmax, t.ymin, t.ymax, h, w);
        if (ok && area <= SMALL_BOX && (lane & (split - 1)) == part) {
            const TriSetup ts = tri_setup(t);
            if (fast) {
                for (int x = t.xmin; x <= t.xmax; x++) {
                    const float v2x = (float)x - t.x0;
                    const float ax = ts.v0x * v2x, bx = ts.v1x * v2x;
                    for (int y = t.ymin; y <= t.ymax; y++)
                        if (inside_px(ts, ax, bx, (float)y - t.y0)) atomicMax(imax + (size_t)(y - y_lo) * w + x, 0xffffffffu - (unsigned)i);
                }
            } else {
                for (int y = t.ymin; y <= t.ymax; y++)
                    for (int x = t.xmin; x <= t.xmax; x++) test_pixel<STAGE>(t, ts, i, x, y, h, w, y_lo, flat, dmax, imax);
            }
        }
        unsigned big = __ballot_sync(0xffffffffu, ok && area > SMALL_BOX && (lane & (split - 1)) == part);
        const unsigned big_fast = __ballot_sync(0xffffffffu, fast);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            Tri s;
            s.x0 = __shfl_sync(0xffffffffu, t.x0, src); s.y0 = __shfl_sync(0xffffffffu, t.y0, src); s.z0 = __shfl_sync(0xffffffffu, t.z0, src);
            s.x1 = __shfl_sync(0xffffffffu, t.x1, src); s.y1 = __shfl_sync(0xffffffffu, t.y1, src); s.z1 = __shfl_sync(0xffffffffu, t.z1, src);
            s.x2 = __shfl_sync(0xffffffffu, t.x2, src); s.y2 = __shfl_sync(0xffffffffu, t.y2, src); s.z2 = __shfl_sync(0xffffffffu, t.z2, src);
            s.xmin = __shfl_sync(0xffffffffu, t.xmin, src); s.ymin = __shfl_sync(0xffffffffu, t.ymin, src);
            const int sbw = __shfl_sync(0xffffffffu, bw, src);
            const int sarea = __shfl_sync(0xffffffffu, area, src);
            const int sidx = (int)base + src;
            const TriSetup ts = tri_setup(s);                  // once per triangle (every lane computes the same values)
            if (STAGE == 1 && ((big_fast >> src) & 1u)) {
                // the warp covers the box in steps of (32 / cw) rows x cw columns and moves right by cw columns per block; a lane keeps its
                // column -- the column products are computed once per block -- and walks down the rows.  cw = the power of two in 8..32
                // that wastes the fewest lanes on the last block (a 34-pixel box: 8 -> 85 % of the lanes busy, 32 -> 53 %), boxes
                // narrower than 8 take the next power of two
                int sh;
                if (sbw <= 8) sh = sbw > 4 ? 3 : sbw > 2 ? 2 : 1;
                else {
                    const int w32 = ((sbw + 31) & ~31) - sbw, w16 = ((sbw + 15) & ~15) - sbw, w8 = ((sbw + 7) & ~7) - sbw;
                    sh = (w32 <= w16 && w32 <= w8) ? 5 : (w16 <= w8 ? 4 : 3);
                }
                const int cw = 1 << sh, col = lane & (cw - 1), rof = lane >> sh, rstep = 32 >> sh;
                const int symax = s.ymin + sarea / sbw - 1;
                const unsigned key = 0xffffffffu - (unsigned)sidx;
                for (int x0 = col; x0 < sbw; x0 += cw) {
                    const int x = s.xmin + x0;
                    const float v2x = (float)x - s.x0;
                    const float ax = ts.v0x * v2x, bx = ts.v1x * v2x;
                    unsigned* __restrict__ dst = imax + (size_t)(s.ymin + rof - y_lo) * w + x;
                    for (int y = s.ymin + rof; y <= symax; y += rstep, dst += (size_t)rstep * w)
                        if (inside_px(ts, ax, bx, (float)y - s.y0)) atomicMax(dst, key);
                }
                continue;
            }
            // lanes walk the box in row-major order, 32 pixels per step; (xx, yy) advance incrementally (no division)
            int xx = lane, yy = 0;
            while (xx >= sbw) { xx -= sbw; yy++; }
            for (int k = lane; k < sarea; k += 32) {
                test_pixel<STAGE>(s, ts, sidx, s.xmin + xx, s.ymin + yy, h, w, y_lo, flat, dmax, imax);
                xx += 32;
                while (xx >= sbw) { xx -= sbw; yy++; }
            }
        }
    }
}

__global__ void __launch_bounds__(256) f3d_clear_if_nonflat_kernel(unsigned* __restrict__ plane, size_t n, const int* __restrict__ nonflat)
{
    if (*nonflat == 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) plane[i] = 0u;
}

__global__ void __launch_bounds__(256) f3d_flat_kernel(const float* __restrict__ vertices, int nver, int* __restrict__ nonflat)
{
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nver; i += (long long)gridDim.x * blockDim.x)
        bad = bad || !(vertices[3 * i + 2] == 0.0f);
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(nonflat, 1);
}

// Per-triangle record for the shade pass (7 x 16 bytes): origin, edge vectors, dot products, inverse determinant, depths and the
// three vertex colours -- everything bary_px and the interpolation need, written once per triangle instead of re-derived per
// pixel from 3 indices + 18 dependent vertex / colour loads and an IEEE division (ncu: 180 instructions per shaded pixel).
// Same operations in the same order as the per-pixel path, so the bits do not change.  Used when c == 3 and the workspace
// has room (ntri <= F3D_REC_MAX_TRIS); larger meshes shade from the raw arrays.
constexpr int F3D_REC_F4 = 7;
#ifndef F3D_REC_MAX_TRIS
#define F3D_REC_MAX_TRIS (2 << 20)
#endif

__global__ void __launch_bounds__(256)
f3d_setup_kernel(const float* __restrict__ vertices, const int* __restrict__ triangles, const float* __restrict__ colors, int nver, int ntri,
                 float4* __restrict__ rec)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntri) return;
    const int i0 = triangles[3 * i], i1 = triangles[3 * i + 1], i2 = triangles[3 * i + 2];
    float4* __restrict__ r = rec + (size_t)i * F3D_REC_F4;
    if ((unsigned)i0 >= (unsigned)nver || (unsigned)i1 >= (unsigned)nver || (unsigned)i2 >= (unsigned)nver) return;   // never a winner
    Tri t;
    t.x0 = vertices[3 * i0]; t.y0 = vertices[3 * i0 + 1]; t.z0 = vertices[3 * i0 + 2];
    t.x1 = vertices[3 * i1]; t.y1 = vertices[3 * i1 + 1]; t.z1 = vertices[3 * i1 + 2];
    t.x2 = vertices[3 * i2]; t.y2 = vertices[3 * i2 + 1]; t.z2 = vertices[3 * i2 + 2];
    const TriSetup s = tri_setup(t);
    const float* __restrict__ a0 = colors + 3 * (size_t)i0; const float* __restrict__ a1 = colors + 3 * (size_t)i1;
    const float* __restrict__ a2 = colors + 3 * (size_t)i2;
    r[0] = make_float4(t.x0, t.y0, s.v0x, s.v0y);
    r[1] = make_float4(s.v1x, s.v1y, s.dot00, s.dot01);
    r[2] = make_float4(s.dot11, s.inv, t.z0, t.z1);
    r[3] = make_float4(t.z2, a0[0], a0[1], a0[2]);
    r[4] = make_float4(a1[0], a1[1], a1[2], a2[0]);
    r[5] = make_float4(a2[1], a2[2], 0.f, 0.f);
    r[6] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool FILL, bool U8>
__global__ void __launch_bounds__(256)
f3d_shade_rec_kernel(float* __restrict__ image, uint8_t* __restrict__ image_u8, const float4* __restrict__ rec, float* __restrict__ depth,
                     float depth_init, int ntri, int h, int w, int y_lo, int y_hi, const unsigned* __restrict__ imax)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    for (int y = y_lo + blockIdx.y; y <= y_hi; y += gridDim.y) {
        const size_t pix = (size_t)y * w + x;
        const unsigned key = imax[(size_t)(y - y_lo) * w + x];
        bool drawn = false;
        float col[3] = {0.f, 0.f, 0.f};
        if (key != 0u) {
            const int idx = (int)(0xffffffffu - key);
            if (idx >= 0 && idx < ntri) {
                const float4* __restrict__ r = rec + (size_t)idx * F3D_REC_F4;
                const float4 q0 = __ldg(r), q1 = __ldg(r + 1), q2 = __ldg(r + 2), q3 = __ldg(r + 3), q4 = __ldg(r + 4), q5 = __ldg(r + 5);
                Tri t;
                t.x0 = q0.x; t.y0 = q0.y; t.z0 = q2.z; t.z1 = q2.w; t.z2 = q3.x;
                TriSetup s;
                s.v0x = q0.z; s.v0y = q0.w; s.v1x = q1.x; s.v1y = q1.y; s.dot00 = q1.z; s.dot01 = q1.w; s.dot11 = q2.x; s.inv = q2.y;
                float w0, w1, w2; bool inside;
                bary_px((float)x, (float)y, t, s, w0, w1, w2, inside);
                const float d = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
                drawn = d > (depth ? depth[pix] : depth_init);
                if (drawn) {
                    if (depth) depth[pix] = d;
                    col[0] = w0 * q3.y + w1 * q4.x + w2 * q4.w;
                    col[1] = w0 * q3.z + w1 * q4.y + w2 * q5.x;
                    col[2] = w0 * q3.w + w1 * q4.z + w2 * q5.y;
                }
            }
        }
        if (drawn || FILL) {
            if (U8) {
                uint8_t* __restrict__ o = image_u8 + pix * 3;
                o[0] = (unsigned char)(int)(col[0] * 255.0f); o[1] = (unsigned char)(int)(col[1] * 255.0f); o[2] = (unsigned char)(int)(col[2] * 255.0f);
            } else {
                float* __restrict__ o = image + pix * 3;
                o[0] = col[0]; o[1] = col[1]; o[2] = col[2];
            }
        }
    }
}

// The same resolve with FOUR consecutive pixels of a row per thread (w % 4 == 0, 16-byte aligned planes): one 16-byte key load, the
// triangle record is fetched once per run of equal winners (inside a triangle all four pixels share it), and the twelve colour values
// leave as three 16-byte stores (fp32) or three 4-byte stores (uint8) instead of twelve scalar ones.  Same per-pixel arithmetic.
struct RecCache { int idx; float4 q0, q1, q2, q3, q4, q5; };

template <bool FILL, bool U8>
__device__ __forceinline__ void shade4_row(float* __restrict__ image, uint8_t* __restrict__ image_u8, const float4* __restrict__ rec,
                                           float* __restrict__ depth, float depth_init, int ntri, int w, int x4, int y, const uint4 k4, RecCache& rc)
{
    int& cached = rc.idx;
    float4 &q0 = rc.q0, &q1 = rc.q1, &q2 = rc.q2, &q3 = rc.q3, &q4 = rc.q4, &q5 = rc.q5;
    const size_t pix = (size_t)y * w + x4;
    const unsigned key[4] = {k4.x, k4.y, k4.z, k4.w};
    float col[12];
    bool drawn[4];
    #pragma unroll
    for (int e = 0; e < 4; e++) {
        drawn[e] = false;
        col[3 * e] = col[3 * e + 1] = col[3 * e + 2] = 0.f;
        if (key[e] == 0u) continue;
        const int idx = (int)(0xffffffffu - key[e]);
        if (idx < 0 || idx >= ntri) continue;
        if (idx != cached) {
            const float4* __restrict__ r = rec + (size_t)idx * F3D_REC_F4;
            q0 = __ldg(r); q1 = __ldg(r + 1); q2 = __ldg(r + 2); q3 = __ldg(r + 3); q4 = __ldg(r + 4); q5 = __ldg(r + 5);
            cached = idx;
        }
        Tri t;
        t.x0 = q0.x; t.y0 = q0.y; t.z0 = q2.z; t.z1 = q2.w; t.z2 = q3.x;
        TriSetup s;
        s.v0x = q0.z; s.v0y = q0.w; s.v1x = q1.x; s.v1y = q1.y; s.dot00 = q1.z; s.dot01 = q1.w; s.dot11 = q2.x; s.inv = q2.y;
        float w0, w1, w2; bool inside;
        bary_px((float)(x4 + e), (float)y, t, s, w0, w1, w2, inside);
        const float d = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
        drawn[e] = d > (depth ? depth[pix + e] : depth_init);
        if (drawn[e]) {
            if (depth) depth[pix + e] = d;
            col[3 * e] = w0 * q3.y + w1 * q4.x + w2 * q4.w;
            col[3 * e + 1] = w0 * q3.z + w1 * q4.y + w2 * q5.x;
            col[3 * e + 2] = w0 * q3.w + w1 * q4.z + w2 * q5.y;
        }
    }
    if (FILL || (drawn[0] && drawn[1] && drawn[2] && drawn[3])) {
        if (U8) {
            unsigned b[3];
            #pragma unroll
            for (int q = 0; q < 3; q++)
                b[q] = (unsigned)(unsigned char)(int)(col[4 * q] * 255.0f) | ((unsigned)(unsigned char)(int)(col[4 * q + 1] * 255.0f) << 8) |
                       ((unsigned)(unsigned char)(int)(col[4 * q + 2] * 255.0f) << 16) | ((unsigned)(unsigned char)(int)(col[4 * q + 3] * 255.0f) << 24);
            unsigned* __restrict__ o = reinterpret_cast<unsigned*>(image_u8 + pix * 3);
            o[0] = b[0]; o[1] = b[1]; o[2] = b[2];
        } else {
            float4* __restrict__ o = reinterpret_cast<float4*>(image + pix * 3);
            o[0] = make_float4(col[0], col[1], col[2], col[3]);
            o[1] = make_float4(col[4], col[5], col[6], col[7]);
            o[2] = make_float4(col[8], col[9], col[10], col[11]);
        }
    } else {
        #pragma unroll
        for (int e = 0; e < 4; e++) {
            if (!drawn[e]) continue;
            if (U8) {
                uint8_t* __restrict__ o = image_u8 + (pix + e) * 3;
                o[0] = (unsigned char)(int)(col[3 * e] * 255.0f); o[1] = (unsigned char)(int)(col[3 * e + 1] * 255.0f);
                o[2] = (unsigned char)(int)(col[3 * e + 2] * 255.0f);
            } else {
                float* __restrict__ o = image + (pix + e) * 3;
                o[0] = col[3 * e]; o[1] = col[3 * e + 1]; o[2] = col[3 * e + 2];
            }
        }
    }
}

// F3D_SHADE_ROWS rows per trip: all their 16-byte key loads are requested before any row is shaded (the key is the first link of a
// key -> record -> store chain of DRAM / L2 round trips), and the record fetched for one row usually serves the next.
#ifndef F3D_SHADE_ROWS
#define F3D_SHADE_ROWS 2
#endif
template <bool FILL, bool U8>
__global__ void __launch_bounds__(256)
f3d_shade_rec4_kernel(float* __restrict__ image, uint8_t* __restrict__ image_u8, const float4* __restrict__ rec, float* __restrict__ depth,
                      float depth_init, int ntri, int h, int w, int y_lo, int y_hi, const unsigned* __restrict__ imax)
{
    constexpr int R = F3D_SHADE_ROWS;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= w) return;
    RecCache rc;
    rc.idx = -1;
    rc.q0 = rc.q1 = rc.q2 = rc.q3 = rc.q4 = rc.q5 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = y_lo + R * blockIdx.y; y <= y_hi; y += R * gridDim.y) {
        const uint4* __restrict__ kp = reinterpret_cast<const uint4*>(imax + (size_t)(y - y_lo) * w + x4);
        uint4 k[R];
        #pragma unroll
        for (int r = 0; r < R; r++) k[r] = y + r <= y_hi ? kp[(size_t)r * (w / 4)] : make_uint4(0u, 0u, 0u, 0u);
        #pragma unroll
        for (int r = 0; r < R; r++)
            if (y + r <= y_hi) shade4_row<FILL, U8>(image, image_u8, rec, depth, depth_init, ntri, w, x4, y + r, k[r], rc);
    }
}

// Pixel-parallel resolve of rows [y_lo, y_hi]: decode the winner, recompute its weights (same bits), depth-test against the
// caller's depth buffer -- or against the constant `depth_init` when the caller has none (face3d/mesh/render.py:68 creates its
// own, filled with -999999, and throws it away: 268 MB of reads and 268 MB of writes at 8192^2 for nothing) -- and write.
// FILL: the image is uninitialised; uncovered pixels receive zeros here (render.py:66 np.zeros) instead of a separate 805 MB
// clear.  U8: write (uint8)(value * 255) (helpers.py:959) instead of fp32: a quarter of the bytes.
template <bool FILL, bool U8>
__global__ void __launch_bounds__(256)
f3d_shade_kernel(float* __restrict__ image, uint8_t* __restrict__ image_u8, const float* __restrict__ vertices,
                 const int* __restrict__ triangles, const float* __restrict__ colors, float* __restrict__ depth, float depth_init,
                 int nver, int ntri, int h, int w, int c, int y_lo, int y_hi, const unsigned* __restrict__ imax)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    for (int y = y_lo + blockIdx.y; y <= y_hi; y += gridDim.y) {
        const size_t pix = (size_t)y * w + x;
        const unsigned key = imax[(size_t)(y - y_lo) * w + x];
        bool drawn = false;
        float col[3] = {0.f, 0.f, 0.f};
        int i0 = 0, i1 = 0, i2 = 0;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f;
        if (key != 0u) {
            const int idx = (int)(0xffffffffu - key);
            if (idx >= 0 && idx < ntri) {
                i0 = triangles[3 * idx]; i1 = triangles[3 * idx + 1]; i2 = triangles[3 * idx + 2];
                Tri t;
                t.x0 = vertices[3 * i0]; t.y0 = vertices[3 * i0 + 1]; t.z0 = vertices[3 * i0 + 2];
                t.x1 = vertices[3 * i1]; t.y1 = vertices[3 * i1 + 1]; t.z1 = vertices[3 * i1 + 2];
                t.x2 = vertices[3 * i2]; t.y2 = vertices[3 * i2 + 1]; t.z2 = vertices[3 * i2 + 2];
                bool inside;
                bary((float)x, (float)y, t, w0, w1, w2, inside);
                const float d = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
                drawn = d > (depth ? depth[pix] : depth_init);
                if (drawn && depth) depth[pix] = d;
            }
        }
        if (c == 3) {
            if (drawn) {
                const float* __restrict__ a0 = colors + 3 * (size_t)i0; const float* __restrict__ a1 = colors + 3 * (size_t)i1;
                const float* __restrict__ a2 = colors + 3 * (size_t)i2;
                col[0] = w0 * a0[0] + w1 * a1[0] + w2 * a2[0];
                col[1] = w0 * a0[1] + w1 * a1[1] + w2 * a2[1];
                col[2] = w0 * a0[2] + w1 * a1[2] + w2 * a2[2];
            }
            if (drawn || FILL) {
                if (U8) {
                    uint8_t* __restrict__ o = image_u8 + pix * 3;
                    o[0] = (unsigned char)(int)(col[0] * 255.0f); o[1] = (unsigned char)(int)(col[1] * 255.0f); o[2] = (unsigned char)(int)(col[2] * 255.0f);
                } else {
                    float* __restrict__ o = image + pix * 3;
                    o[0] = col[0]; o[1] = col[1]; o[2] = col[2];
                }
            }
        } else if (drawn || FILL) {
            for (int k = 0; k < c; k++) {
                float v = 0.f;
                if (drawn) v = w0 * colors[(size_t)c * i0 + k] + w1 * colors[(size_t)c * i1 + k] + w2 * colors[(size_t)c * i2 + k];
                if (U8) image_u8[pix * c + k] = (unsigned char)(int)(v * 255.0f); else image[pix * c + k] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256) f3d_to_u8_kernel(const float* __restrict__ image, uint8_t* __restrict__ out, long long n)
{
    // (image*255).astype(np.uint8): C truncation toward zero of the float product (helpers.py:959)
    const long long n4 = n >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(image)[q];
        uchar4 o;
        o.x = (unsigned char)(int)(v.x * 255.0f); o.y = (unsigned char)(int)(v.y * 255.0f);
        o.z = (unsigned char)(int)(v.z * 255.0f); o.w = (unsigned char)(int)(v.w * 255.0f);
        reinterpret_cast<uchar4*>(out)[q] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long k = n4 << 2; k < n; k++) out[k] = (unsigned char)(int)(image[k] * 255.0f);
}


}  // namespace

// Workspace: a 256-byte header (the "some vertex z != 0" flag) and two u32 key planes for one BAND of rows; with all z = 0
// (the texture bake: helpers.py:945-950) only one plane is used.  Bands exist so that images beyond the budget still resolve
// with bounded scratch; at 8192^2 the whole image is ONE band.  (Measured, r02: resolving 8192^2 in L2-sized bands of 768 rows
// was 3x SLOWER -- 3.5 ms vs 1.1 ms -- because every band re-walks the triangle list and, triangles being stored row by row,
// only the ~9 % of the warps whose 32 triangles touch the band have work.)
#ifndef F3D_BAND_BYTES
#define F3D_BAND_BYTES (1ll << 30)
#endif
static thread_local long long g_band_bytes = F3D_BAND_BYTES;
extern "C" void f3d_set_band_bytes(int64_t bytes) { g_band_bytes = bytes > 0 ? bytes : F3D_BAND_BYTES; }

static int f3d_band_rows(int ntri, int h, int w)
{
    (void)ntri;
    const long long budget = g_band_bytes;                        // bytes per key plane and band
    long long rows = budget / ((long long)w * 4);
    if (rows < 16) rows = 16;
    if (rows > h) rows = h;
    return (int)rows;
}

extern "C" size_t f3d_workspace_bytes(int32_t ntri, int32_t h, int32_t w)
{
    if (h < 1 || w < 1) return 0;
    const size_t recs = (ntri > 0 && ntri <= F3D_REC_MAX_TRIS) ? (size_t)ntri * F3D_REC_F4 * sizeof(float4) : 0;
    return 256 + gs_align_up(2 * (size_t)f3d_band_rows(ntri, h, w) * (size_t)w * sizeof(unsigned), 16) + recs;
}

static int f3d_run(float* image, uint8_t* image_u8, bool fill, const float* vertices, const int32_t* triangles, const float* colors,
                   float* depth_buffer, float depth_init, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c,
                   void* workspace, size_t workspace_bytes, cudaStream_t s)
{
    if ((!image && !image_u8) || h < 1 || w < 1 || c < 1 || nver < 0 || ntri < 0) return F3D_E_BAD_ARGS;
    if (ntri > 0 && (!vertices || !triangles || !colors)) return F3D_E_BAD_ARGS;
    if (!workspace || workspace_bytes < f3d_workspace_bytes(ntri, h, w)) return F3D_E_WORKSPACE;
    if (ntri == 0 && !fill) return F3D_OK;
    const int band = f3d_band_rows(ntri, h, w);
    int* nonflat = (int*)workspace;
    unsigned* imax = (unsigned*)((char*)workspace + 256);
    unsigned* dmax = imax + (size_t)band * w;
    float4* rec = (c == 3 && ntri > 0 && ntri <= F3D_REC_MAX_TRIS)
                      ? (float4*)((char*)workspace + 256 + gs_align_up(2 * (size_t)band * w * sizeof(unsigned), 16)) : nullptr;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaMemsetAsync(nonflat, 0, 256, s) != cudaSuccess) return F3D_E_CUDA;
    if (nver > 0) f3d_flat_kernel<<<sms * 2, 256, 0, s>>>(vertices, nver, nonflat);
    if (rec) f3d_setup_kernel<<<(ntri + 255) / 256, 256, 0, s>>>(vertices, triangles, colors, nver, ntri, rec);
    // warps per group of 32 triangles: 16 when a triangle covers hundreds of pixels (the 8K bake of a 120 k-triangle head: ~560), 1 for
    // pixel-sized triangles (measured at 9.6 M triangles: 1.46 ms with 1, 1.95 / 2.86 / 4.30 ms with 2 / 4 / 8)
    const long long px_per_tri = ntri > 0 ? (long long)h * w / ntri : 0;
    const int split_log2 = px_per_tri >= 256 ? F3D_SPLIT_MAX_LOG2 : px_per_tri >= 96 ? 2 : px_per_tri >= 48 ? 1 : 0;
    const long long warps_needed = (((long long)ntri + 31) / 32) << split_log2;
    long long blocks1 = (warps_needed + 7) / 8;
    if (blocks1 > (long long)sms * 64) blocks1 = (long long)sms * 64;
    if (blocks1 < 1) blocks1 = 1;
    {   // a group's warps are consecutive global warp indices: the grid must hold whole groups (8 warps per block)
        const long long m = (1ll << split_log2) > 8 ? (1ll << split_log2) / 8 : 1;
        blocks1 = (blocks1 + m - 1) / m * m;
    }
    for (int y_lo = 0; y_lo < h; y_lo += band) {
        const int y_hi = (y_lo + band < h ? y_lo + band : h) - 1, rows = y_hi - y_lo + 1;
        if (cudaMemsetAsync(imax, 0, (size_t)band * w * sizeof(unsigned), s) != cudaSuccess) return F3D_E_CUDA;
        f3d_clear_if_nonflat_kernel<<<sms * 4, 256, 0, s>>>(dmax, (size_t)band * w, nonflat);      // the depth plane only when it will be used
        if (ntri > 0) {
            f3d_tri_kernel<0><<<(unsigned)blocks1, 256, 0, s>>>(vertices, triangles, nver, ntri, h, w, y_lo, y_hi, nonflat, dmax, imax, split_log2);
            f3d_tri_kernel<1><<<(unsigned)blocks1, 256, 0, s>>>(vertices, triangles, nver, ntri, h, w, y_lo, y_hi, nonflat, dmax, imax, split_log2);
        }
        const dim3 grid2((unsigned)((w + 255) / 256), (unsigned)(rows < 65535 ? rows : 65535));
        const bool vec4 = rec && (w & 3) == 0 && (((uintptr_t)image | (uintptr_t)image_u8 | (uintptr_t)imax) & 15u) == 0;
        if (vec4) {
            const dim3 grid4((unsigned)((w / 4 + 255) / 256), (unsigned)((rows + F3D_SHADE_ROWS - 1) / F3D_SHADE_ROWS < 65535 ? (rows + F3D_SHADE_ROWS - 1) / F3D_SHADE_ROWS : 65535));
            if (fill && image_u8) f3d_shade_rec4_kernel<true, true><<<grid4, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
            else if (fill) f3d_shade_rec4_kernel<true, false><<<grid4, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
            else if (image_u8) f3d_shade_rec4_kernel<false, true><<<grid4, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
            else f3d_shade_rec4_kernel<false, false><<<grid4, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
        } else if (rec) {
            if (fill && image_u8) f3d_shade_rec_kernel<true, true><<<grid2, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
            else if (fill) f3d_shade_rec_kernel<true, false><<<grid2, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
            else if (image_u8) f3d_shade_rec_kernel<false, true><<<grid2, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
            else f3d_shade_rec_kernel<false, false><<<grid2, 256, 0, s>>>(image, image_u8, rec, depth_buffer, depth_init, ntri, h, w, y_lo, y_hi, imax);
        } else if (fill && image_u8)
            f3d_shade_kernel<true, true><<<grid2, 256, 0, s>>>(image, image_u8, vertices, triangles, colors, depth_buffer, depth_init, nver, ntri, h, w, c, y_lo, y_hi, imax);
        else if (fill)
            f3d_shade_kernel<true, false><<<grid2, 256, 0, s>>>(image, image_u8, vertices, triangles, colors, depth_buffer, depth_init, nver, ntri, h, w, c, y_lo, y_hi, imax);
        else if (image_u8)
            f3d_shade_kernel<false, true><<<grid2, 256, 0, s>>>(image, image_u8, vertices, triangles, colors, depth_buffer, depth_init, nver, ntri, h, w, c, y_lo, y_hi, imax);
        else
            f3d_shade_kernel<false, false><<<grid2, 256, 0, s>>>(image, image_u8, vertices, triangles, colors, depth_buffer, depth_init, nver, ntri, h, w, c, y_lo, y_hi, imax);
    }
    return cudaGetLastError() == cudaSuccess ? F3D_OK : F3D_E_CUDA;
}

extern "C" int f3d_render_colors(float* image, const float* vertices, const int32_t* triangles, const float* colors,
                                 float* depth_buffer, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c,
                                 void* workspace, size_t workspace_bytes, gs_stream_t stream)
{
    if (!image || !depth_buffer) return F3D_E_BAD_ARGS;
    return f3d_run(image, nullptr, false, vertices, triangles, colors, depth_buffer, 0.f, nver, ntri, h, w, c, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

extern "C" int f3d_bake_colors(float* image, uint8_t* image_u8, const float* vertices, const int32_t* triangles, const float* colors,
                               float depth_init, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c,
                               void* workspace, size_t workspace_bytes, gs_stream_t stream)
{
    if ((image != nullptr) == (image_u8 != nullptr)) return F3D_E_BAD_ARGS;       // exactly one output format
    return f3d_run(image, image_u8, true, vertices, triangles, colors, nullptr, depth_init, nver, ntri, h, w, c, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

// Host-pointer entry with the calling convention of the reference's Cython shim (mesh_core_cython.pyx:64-77): every pointer
// is HOST memory, `image` [h,w,c] and `depth_buffer` [h,w] are read, updated in place and complete on return.  The one entry
// point of the library that owns device memory (stream-ordered allocations, released before returning).
extern "C" int f3d_render_colors_host(float* image, const float* vertices, const int32_t* triangles, const float* colors,
                                      float* depth_buffer, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c)
{
    if (!image || !depth_buffer || h < 1 || w < 1 || c < 1 || nver < 0 || ntri < 0) return F3D_E_BAD_ARGS;
    if (ntri > 0 && (!vertices || !triangles || !colors)) return F3D_E_BAD_ARGS;
    if (ntri == 0) return F3D_OK;
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return F3D_E_CUDA;
    const size_t P = (size_t)h * w, nb_img = P * c * 4, nb_dep = P * 4, nb_v = (size_t)nver * 12, nb_t = (size_t)ntri * 12,
                 nb_c = (size_t)nver * c * 4, nb_ws = f3d_workspace_bytes(ntri, h, w);
    void *d_img = nullptr, *d_dep = nullptr, *d_v = nullptr, *d_t = nullptr, *d_c = nullptr, *d_ws = nullptr;
    int rc = F3D_E_CUDA;
    if (cudaMallocAsync(&d_img, nb_img, s) == cudaSuccess && cudaMallocAsync(&d_dep, nb_dep, s) == cudaSuccess &&
        cudaMallocAsync(&d_v, nb_v, s) == cudaSuccess && cudaMallocAsync(&d_t, nb_t, s) == cudaSuccess &&
        cudaMallocAsync(&d_c, nb_c, s) == cudaSuccess && cudaMallocAsync(&d_ws, nb_ws, s) == cudaSuccess &&
        cudaMemcpyAsync(d_img, image, nb_img, cudaMemcpyHostToDevice, s) == cudaSuccess &&
        cudaMemcpyAsync(d_dep, depth_buffer, nb_dep, cudaMemcpyHostToDevice, s) == cudaSuccess &&
        cudaMemcpyAsync(d_v, vertices, nb_v, cudaMemcpyHostToDevice, s) == cudaSuccess &&
        cudaMemcpyAsync(d_t, triangles, nb_t, cudaMemcpyHostToDevice, s) == cudaSuccess &&
        cudaMemcpyAsync(d_c, colors, nb_c, cudaMemcpyHostToDevice, s) == cudaSuccess) {
        rc = f3d_run((float*)d_img, nullptr, false, (const float*)d_v, (const int32_t*)d_t, (const float*)d_c, (float*)d_dep, 0.f,
                     nver, ntri, h, w, c, d_ws, nb_ws, s);
        if (rc == F3D_OK && (cudaMemcpyAsync(image, d_img, nb_img, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                             cudaMemcpyAsync(depth_buffer, d_dep, nb_dep, cudaMemcpyDeviceToHost, s) != cudaSuccess)) rc = F3D_E_CUDA;
    }
    void* all[] = {d_img, d_dep, d_v, d_t, d_c, d_ws};
    for (void* q : all) if (q) cudaFreeAsync(q, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = F3D_E_CUDA;
    cudaStreamDestroy(s);
    return rc;
}

extern "C" int f3d_image_to_u8(const float* image, uint8_t* out_u8, int64_t count, gs_stream_t stream)
{
    if (!image || !out_u8 || count < 0) return F3D_E_BAD_ARGS;
    if (count == 0) return F3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    f3d_to_u8_kernel<<<148 * 16, 256, 0, s>>>(image, out_u8, (long long)count);
    return cudaGetLastError() == cudaSuccess ? F3D_OK : F3D_E_CUDA;
}
