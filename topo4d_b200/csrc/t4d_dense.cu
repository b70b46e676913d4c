// t4d_dense.cu -- dense Gaussian-mesh attribute interpolation on the device, sm_100a.
//
// Replaces compute_vertex_attribute_by_weight_2 (reference helpers.py:237-253), which update_dense_states calls once
// per frame (train.py:498-508) as  GPU tensor -> .cpu().numpy() -> NumPy fancy indexing -> torch.from_numpy -> .cuda():
// the first n_base rows of the dense attribute are the base-mesh rows, every further row i is the bilinear blend
//     sum_j  weight[i][j] * attribute[ quad_faces[ vertex_father[i] ][j] ],   j = 0..3
// of the four corners of its father quad.  NumPy evaluates it in float64 (the weights are float64), products first,
// then a left-to-right sum over j, and the caller casts to float32; the kernel does exactly that (unfused
// __dmul_rn / __dadd_rn, one final __double2float_rn) so the result is bit-identical.  One thread per output element;
// reads are gathers from an L2-resident base table, writes are coalesced.
#include <cuda_runtime.h>
#include "../../include/topo4d_b200.h"

namespace {

__global__ void __launch_bounds__(256) dense_attribute_kernel(const float* __restrict__ attr, int n_base, int ch,
                                                              const int32_t* __restrict__ quad_faces,
                                                              const int32_t* __restrict__ father,
                                                              const double* __restrict__ weight, long long total,
                                                              float* __restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long row = e / ch;
    const int c = (int)(e - row * ch);
    if (row < n_base) { out[e] = attr[e]; return; }
    const long long i = row - n_base;
    const int32_t* __restrict__ q = quad_faces + 4 * (size_t)father[i];
    const double* __restrict__ w = weight + 4 * i;
    double acc = __dmul_rn((double)attr[(size_t)q[0] * ch + c], w[0]);
    #pragma unroll
    for (int j = 1; j < 4; j++) acc = __dadd_rn(acc, __dmul_rn((double)attr[(size_t)q[j] * ch + c], w[j]));
    out[e] = __double2float_rn(acc);
}

}  // namespace

extern "C" int t4d_dense_attribute(const float* attribute, int32_t n_base, int32_t channels, const int32_t* quad_faces,
                                   const int32_t* vertex_father, const double* weight, int32_t n_new, float* dense_out,
                                   gs_stream_t stream)
{
    if (n_base < 0 || n_new < 0 || channels < 1) return GS_E_BAD_ARGS;
    const long long total = ((long long)n_base + n_new) * channels;
    if (total == 0) return 0;
    if (!dense_out || (n_base > 0 && !attribute)) return GS_E_BAD_ARGS;
    if (n_new > 0 && (!quad_faces || !vertex_father || !weight || !attribute)) return GS_E_BAD_ARGS;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return GS_E_UNSUPPORTED;
    dense_attribute_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(attribute, n_base, channels, quad_faces,
                                                                               vertex_father, weight, total, dense_out);
    return cudaGetLastError() == cudaSuccess ? 0 : GS_E_CUDA;
}
