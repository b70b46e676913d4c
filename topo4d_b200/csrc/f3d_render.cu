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
//   pass 1  (triangle-parallel) atomicMax of a 64-bit key (orderable(p_depth)<<32 | ~index) per
//           accepted bbox pixel; a warp takes 32 triangles, small boxes are walked by their own
//           lane, large boxes cooperatively by the whole warp;
//   pass 2  (pixel-parallel, coalesced) decode the winner, recompute its weights (same bits),
//           depth-test against the caller's depth buffer, write c channels + depth in place.
// Roofline: HBM -- h*w*(8 key + 8 key re-read + 4c image + 8 depth) bytes; no data reuse worth smem.
#include "gs_common.cuh"

namespace {

struct Tri {
    float x0, y0, z0, x1, y1, z1, x2, y2, z2;
    int xmin, xmax, ymin, ymax;
};

// reference get_point_weight / isPointInTri arithmetic (mesh_core.cpp:23-82), op for op
__device__ __forceinline__ void bary(float px, float py, const Tri& t, float& w0, float& w1, float& w2, bool& inside)
{
    const float v0x = t.x2 - t.x0, v0y = t.y2 - t.y0;
    const float v1x = t.x1 - t.x0, v1y = t.y1 - t.y0;
    const float v2x = px - t.x0, v2y = py - t.y0;
    const float dot00 = v0x * v0x + v0y * v0y;
    const float dot01 = v0x * v1x + v0y * v1y;
    const float dot02 = v0x * v2x + v0y * v2y;
    const float dot11 = v1x * v1x + v1y * v1y;
    const float dot12 = v1x * v2x + v1y * v2y;
    const float den = dot00 * dot11 - dot01 * dot01;
    const float inv = (den == 0.0f) ? 0.0f : 1.0f / den;
    const float u = (dot11 * dot02 - dot01 * dot12) * inv;
    const float v = (dot00 * dot12 - dot01 * dot02) * inv;
    inside = (u >= 0.0f) && (v >= 0.0f) && (u + v < 1.0f);
    w0 = 1.0f - u - v; w1 = v; w2 = u;
}

__device__ __forceinline__ bool accepts(int x, int y, int h, int w, bool inside)
{
    const float fx = (float)x, fy = (float)y;
    return fx < 2.0f || fx > (float)(w - 3) || fy < 2.0f || fy > (float)(h - 3) || inside;
}

__device__ __forceinline__ unsigned long long make_key(float depth, int index)
{
    unsigned b = __float_as_uint(depth);
    if (depth == 0.0f) b = 0u;                               // -0 == +0 for the strict '>' test
    const unsigned ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)ord << 32) | (unsigned long long)(0xffffffffu - (unsigned)index);
}

__device__ __forceinline__ int clamp_to_int(float f)
{
    f = fminf(fmaxf(f, -1.0e9f), 1.0e9f);                    // NaN -> -1e9 (fmaxf returns the non-NaN)
    return (int)f;
}

__device__ __forceinline__ bool load_tri(const float* __restrict__ vertices, const int* __restrict__ triangles,
                                         int i, int nver, int h, int w, Tri& t)
{
    const int i0 = triangles[3 * i], i1 = triangles[3 * i + 1], i2 = triangles[3 * i + 2];
    if ((unsigned)i0 >= (unsigned)nver || (unsigned)i1 >= (unsigned)nver || (unsigned)i2 >= (unsigned)nver) return false;
    t.x0 = vertices[3 * i0]; t.y0 = vertices[3 * i0 + 1]; t.z0 = vertices[3 * i0 + 2];
    t.x1 = vertices[3 * i1]; t.y1 = vertices[3 * i1 + 1]; t.z1 = vertices[3 * i1 + 2];
    t.x2 = vertices[3 * i2]; t.y2 = vertices[3 * i2 + 1]; t.z2 = vertices[3 * i2 + 2];
    t.xmin = max(clamp_to_int(ceilf(fminf(t.x0, fminf(t.x1, t.x2)))), 0);
    t.xmax = min(clamp_to_int(floorf(fmaxf(t.x0, fmaxf(t.x1, t.x2)))), w - 1);
    t.ymin = max(clamp_to_int(ceilf(fminf(t.y0, fminf(t.y1, t.y2)))), 0);
    t.ymax = min(clamp_to_int(floorf(fmaxf(t.y0, fmaxf(t.y1, t.y2)))), h - 1);
    return !(t.xmax < t.xmin || t.ymax < t.ymin);
}

__device__ __forceinline__ void test_pixel(const Tri& t, int idx, int x, int y, int h, int w,
                                           unsigned long long* __restrict__ keys)
{
    float w0, w1, w2; bool inside;
    bary((float)x, (float)y, t, w0, w1, w2, inside);
    if (!accepts(x, y, h, w, inside)) return;
    const float d = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
    if (d != d) return;                                      // NaN never passes '>' in the reference
    atomicMax(keys + (size_t)y * w + x, make_key(d, idx));
}

constexpr int SMALL_BOX = 48;   // bbox pixels a single lane walks by itself

__global__ void __launch_bounds__(256)
f3d_pass1_kernel(const float* __restrict__ vertices, const int* __restrict__ triangles, int nver, int ntri, int h, int w,
                 unsigned long long* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp_global * 32; base < ntri; base += nwarps * 32) {
        const int i = (int)(base + lane);
        Tri t;
        bool ok = false;
        if (i < ntri) ok = load_tri(vertices, triangles, i, nver, h, w, t);
        int bw = 0, area = 0;
        if (ok) { bw = t.xmax - t.xmin + 1; const long long a = (long long)bw * (t.ymax - t.ymin + 1); area = a > 0x7fffffffLL ? 0x7fffffff : (int)a; }
        if (ok && area <= SMALL_BOX) {
            for (int y = t.ymin; y <= t.ymax; y++)
                for (int x = t.xmin; x <= t.xmax; x++) test_pixel(t, i, x, y, h, w, keys);
        }
        unsigned big = __ballot_sync(0xffffffffu, ok && area > SMALL_BOX);
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
            // lanes walk the box in row-major order, 32 pixels per step; (xx, yy) advance incrementally (no division)
            int xx = lane, yy = 0;
            while (xx >= sbw) { xx -= sbw; yy++; }
            for (int k = lane; k < sarea; k += 32) {
                test_pixel(s, sidx, s.xmin + xx, s.ymin + yy, h, w, keys);
                xx += 32;
                while (xx >= sbw) { xx -= sbw; yy++; }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
f3d_pass2_kernel(float* __restrict__ image, const float* __restrict__ vertices, const int* __restrict__ triangles,
                 const float* __restrict__ colors, float* __restrict__ depth, int nver, int ntri, int h, int w, int c,
                 const unsigned long long* __restrict__ keys)
{
    // one thread per pixel, rows from blockIdx.y (grid-strided): no integer division, coalesced key/depth/image rows
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const size_t pix = (size_t)y * w + x;
        const unsigned long long key = keys[pix];
        if (key == 0ull) continue;
        const int idx = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
        if (idx < 0 || idx >= ntri) continue;
        const int i0 = triangles[3 * idx], i1 = triangles[3 * idx + 1], i2 = triangles[3 * idx + 2];
        Tri t;
        t.x0 = vertices[3 * i0]; t.y0 = vertices[3 * i0 + 1]; t.z0 = vertices[3 * i0 + 2];
        t.x1 = vertices[3 * i1]; t.y1 = vertices[3 * i1 + 1]; t.z1 = vertices[3 * i1 + 2];
        t.x2 = vertices[3 * i2]; t.y2 = vertices[3 * i2 + 1]; t.z2 = vertices[3 * i2 + 2];
        float w0, w1, w2; bool inside;
        bary((float)x, (float)y, t, w0, w1, w2, inside);
        const float d = w0 * t.z0 + w1 * t.z1 + w2 * t.z2;
        if (!(d > depth[pix])) continue;
        if (c == 3) {
            const float* __restrict__ a0 = colors + 3 * (size_t)i0; const float* __restrict__ a1 = colors + 3 * (size_t)i1;
            const float* __restrict__ a2 = colors + 3 * (size_t)i2;
            float* __restrict__ o = image + pix * 3;
            o[0] = w0 * a0[0] + w1 * a1[0] + w2 * a2[0];
            o[1] = w0 * a0[1] + w1 * a1[1] + w2 * a2[1];
            o[2] = w0 * a0[2] + w1 * a1[2] + w2 * a2[2];
        } else {
            for (int k = 0; k < c; k++) {
                const float c0 = colors[(size_t)c * i0 + k], c1 = colors[(size_t)c * i1 + k], c2 = colors[(size_t)c * i2 + k];
                image[pix * c + k] = w0 * c0 + w1 * c1 + w2 * c2;
            }
        }
        depth[pix] = d;
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

extern "C" size_t f3d_workspace_bytes(int32_t ntri, int32_t h, int32_t w)
{
    (void)ntri;
    if (h < 1 || w < 1) return 0;
    return (size_t)h * (size_t)w * sizeof(unsigned long long);
}

extern "C" int f3d_render_colors(float* image, const float* vertices, const int32_t* triangles, const float* colors,
                                 float* depth_buffer, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c,
                                 void* workspace, size_t workspace_bytes, gs_stream_t stream)
{
    if (!image || !depth_buffer || h < 1 || w < 1 || c < 1 || nver < 0 || ntri < 0) return F3D_E_BAD_ARGS;
    if (ntri > 0 && (!vertices || !triangles || !colors)) return F3D_E_BAD_ARGS;
    if (!workspace || workspace_bytes < f3d_workspace_bytes(ntri, h, w)) return F3D_E_WORKSPACE;
    if (ntri == 0) return F3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* keys = (unsigned long long*)workspace;
    cudaError_t e = cudaMemsetAsync(keys, 0, (size_t)h * w * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return F3D_E_CUDA;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long warps_needed = ((long long)ntri + 31) / 32;
    long long blocks1 = (warps_needed + 7) / 8;
    if (blocks1 > (long long)sms * 64) blocks1 = (long long)sms * 64;
    f3d_pass1_kernel<<<(unsigned)blocks1, 256, 0, s>>>(vertices, triangles, nver, ntri, h, w, keys);
    const dim3 grid2((unsigned)((w + 255) / 256), (unsigned)(h < 65535 ? h : 65535));
    f3d_pass2_kernel<<<grid2, 256, 0, s>>>(image, vertices, triangles, colors, depth_buffer, nver, ntri, h, w, c, keys);
    e = cudaGetLastError();
    return e == cudaSuccess ? F3D_OK : F3D_E_CUDA;
}

extern "C" int f3d_image_to_u8(const float* image, uint8_t* out_u8, int64_t count, gs_stream_t stream)
{
    if (!image || !out_u8 || count < 0) return F3D_E_BAD_ARGS;
    if (count == 0) return F3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    f3d_to_u8_kernel<<<148 * 16, 256, 0, s>>>(image, out_u8, (long long)count);
    return cudaGetLastError() == cudaSuccess ? F3D_OK : F3D_E_CUDA;
}
