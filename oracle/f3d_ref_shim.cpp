// f3d_ref_shim.cpp -- extern "C" trampoline around the REFERENCE's own _render_colors_core
// (face3d/mesh/cython/mesh_core.cpp:169-234), which oracle/Makefile compiles from where it lies
// under /root/reference into oracle/_ref/libf3d_ref.so.  No reference code is copied here: this file
// only declares the reference symbol (mesh_core.h) and forwards to it, so ctypes can call it without
// Cython.  TEST INFRASTRUCTURE / CPU baseline only.
#include "mesh_core.h"

extern "C" void f3d_ref_render_colors(float* image, float* vertices, int* triangles, float* colors,
                                      float* depth_buffer, int nver, int ntri, int h, int w, int c)
{
    _render_colors_core(image, vertices, triangles, colors, depth_buffer, nver, ntri, h, w, c);
}
