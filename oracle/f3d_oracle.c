/*
 * f3d_oracle.c -- CPU restatement ("port") of face3d's render_colors painter.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows the reference
 *   _render_colors_core   face3d/mesh/cython/mesh_core.cpp:169-234
 *   isPointInTri          face3d/mesh/cython/mesh_core.cpp:23-50
 *   get_point_weight      face3d/mesh/cython/mesh_core.cpp:53-82
 * as bound by render_colors_core (face3d/mesh/cython/mesh_core_cython.pyx:64-77).
 * Parity PINNED: this port is checked bit-for-bit against the reference's own C++ compiled from
 * /root/reference (oracle/_ref/libf3d_ref.so) and against tests/golden/face3d_small.npz, which that
 * reference build generated (tests/golden/make_golden.py).
 * Build with -ffp-contract=off: coverage decisions depend on unfused fp32 arithmetic.
 */
#include <math.h>

static void weights(float px, float py, const float* p0, const float* p1, const float* p2, float* w, int* inside)
{
    /* v0 = p2 - p0, v1 = p1 - p0, v2 = p - p0 (mesh_core.cpp:27-30, 57-60) */
    float v0x = p2[0] - p0[0], v0y = p2[1] - p0[1];
    float v1x = p1[0] - p0[0], v1y = p1[1] - p0[1];
    float v2x = px - p0[0], v2y = py - p0[1];
    float dot00 = v0x * v0x + v0y * v0y;
    float dot01 = v0x * v1x + v0y * v1y;
    float dot02 = v0x * v2x + v0y * v2y;
    float dot11 = v1x * v1x + v1y * v1y;
    float dot12 = v1x * v2x + v1y * v2y;
    float inv;
    if (dot00 * dot11 - dot01 * dot01 == 0) inv = 0;
    else inv = 1 / (dot00 * dot11 - dot01 * dot01);
    float u = (dot11 * dot02 - dot01 * dot12) * inv;
    float v = (dot00 * dot12 - dot01 * dot02) * inv;
    *inside = (u >= 0) && (v >= 0) && (u + v < 1);
    w[0] = 1 - u - v; w[1] = v; w[2] = u;
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

void f3d_port_render_colors(float* image, const float* vertices, const int* triangles, const float* colors,
                            float* depth_buffer, int nver, int ntri, int h, int w, int c)
{
    (void)nver;
    for (int i = 0; i < ntri; i++) {
        const int i0 = triangles[3 * i], i1 = triangles[3 * i + 1], i2 = triangles[3 * i + 2];
        const float* p0 = vertices + 3 * i0; const float* p1 = vertices + 3 * i1; const float* p2 = vertices + 3 * i2;
        int x_min = imax((int)ceilf(fminf(p0[0], fminf(p1[0], p2[0]))), 0);
        int x_max = imin((int)floorf(fmaxf(p0[0], fmaxf(p1[0], p2[0]))), w - 1);
        int y_min = imax((int)ceilf(fminf(p0[1], fminf(p1[1], p2[1]))), 0);
        int y_max = imin((int)floorf(fmaxf(p0[1], fmaxf(p1[1], p2[1]))), h - 1);
        if (x_max < x_min || y_max < y_min) continue;
        for (int y = y_min; y <= y_max; y++)
            for (int x = x_min; x <= x_max; x++) {
                float wt[3]; int inside;
                weights((float)x, (float)y, p0, p1, p2, wt, &inside);
                /* 2-px border rule (mesh_core.cpp:211) */
                if (!((float)x < 2 || (float)x > w - 3 || (float)y < 2 || (float)y > h - 3 || inside)) continue;
                float d = wt[0] * p0[2] + wt[1] * p1[2] + wt[2] * p2[2];
                if (d > depth_buffer[y * w + x]) {
                    for (int k = 0; k < c; k++)
                        image[((long)y * w + x) * c + k] = wt[0] * colors[c * i0 + k] + wt[1] * colors[c * i1 + k] + wt[2] * colors[c * i2 + k];
                    depth_buffer[y * w + x] = d;
                }
            }
    }
}
