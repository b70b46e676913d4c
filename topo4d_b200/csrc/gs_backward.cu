// gs_backward.cu -- backward preprocess (K8+K9 fused): per-Gaussian 2D gradients -> dL/d{means3D,
// scales, rotations, SH | colours, opacities} (+ means2D in NDC-scaled units, + cov3D_precomp).
//
// Replaces upstream computeCov2DCUDA + preprocessCUDA (backward) behind GaussianRasterizer's
// autograd backward (reference: loss.backward() at train.py:667,738).  Four adjacent lanes own one
// Gaussian and split the V views of the batch (lane q takes views q, q+4, ...), re-deriving the cheap
// forward intermediates (cov3D, J, T) instead of storing them and accumulating in registers; two
// shuffle steps combine the four partial sums and one lane writes, so every output element is written
// exactly once with no atomics -- the caller may point the outputs into one flat buffer that a single
// NCCL all-reduce then sums across view-parallel ranks.
// Conventions: frustum-clamped t.x/t.y are constants, SH max(0,.) kills the
// gradient where clamped, depth = view-space z, mean2D gradient reported x(0.5W, 0.5H).
#include "gs_common.cuh"


#ifndef GS_PBWD_PREFETCH
#define GS_PBWD_PREFETCH 1
#endif

namespace {

__device__ __constant__ float bC0 = 0.28209479177387814f;
__device__ __constant__ float bC1 = 0.4886025119029199f;
__device__ __constant__ float bC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                        -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float bC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                        0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                        -0.5900435899266435f};

// SH basis values and their direction derivatives for degree <= DEG
template <int K>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* b, float* bx, float* by, float* bz)
{
    #pragma unroll
    for (int k = 0; k < K; k++) { b[k] = bx[k] = by[k] = bz[k] = 0.f; }
    b[0] = bC0;
    if constexpr (K > 1) {
        b[1] = -bC1 * y; by[1] = -bC1;
        b[2] = bC1 * z;  bz[2] = bC1;
        b[3] = -bC1 * x; bx[3] = -bC1;
    }
    if constexpr (K > 4) {
        const float xx = x * x, yy = y * y, zz = z * z;
        b[4] = bC2[0] * x * y; bx[4] = bC2[0] * y; by[4] = bC2[0] * x;
        b[5] = bC2[1] * y * z; by[5] = bC2[1] * z; bz[5] = bC2[1] * y;
        b[6] = bC2[2] * (2.f * zz - xx - yy); bx[6] = -2.f * bC2[2] * x; by[6] = -2.f * bC2[2] * y; bz[6] = 4.f * bC2[2] * z;
        b[7] = bC2[3] * x * z; bx[7] = bC2[3] * z; bz[7] = bC2[3] * x;
        b[8] = bC2[4] * (xx - yy); bx[8] = 2.f * bC2[4] * x; by[8] = -2.f * bC2[4] * y;
    }
    if constexpr (K > 9) {
        const float xx = x * x, yy = y * y, zz = z * z;
        b[9] = bC3[0] * y * (3.f * xx - yy); bx[9] = bC3[0] * 6.f * x * y; by[9] = bC3[0] * (3.f * xx - 3.f * yy);
        b[10] = bC3[1] * x * y * z; bx[10] = bC3[1] * y * z; by[10] = bC3[1] * x * z; bz[10] = bC3[1] * x * y;
        b[11] = bC3[2] * y * (4.f * zz - xx - yy); bx[11] = bC3[2] * -2.f * x * y; by[11] = bC3[2] * (4.f * zz - xx - 3.f * yy); bz[11] = bC3[2] * 8.f * y * z;
        b[12] = bC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); bx[12] = bC3[3] * -6.f * x * z; by[12] = bC3[3] * -6.f * y * z; bz[12] = bC3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
        b[13] = bC3[4] * x * (4.f * zz - xx - yy); bx[13] = bC3[4] * (4.f * zz - 3.f * xx - yy); by[13] = bC3[4] * -2.f * x * y; bz[13] = bC3[4] * 8.f * x * z;
        b[14] = bC3[5] * z * (xx - yy); bx[14] = bC3[5] * 2.f * x * z; by[14] = bC3[5] * -2.f * y * z; bz[14] = bC3[5] * (xx - yy);
        b[15] = bC3[6] * x * (xx - 3.f * yy); bx[15] = bC3[6] * (3.f * xx - 3.f * yy); by[15] = bC3[6] * -6.f * x * y;
    }
}

template <int K, int LPG>   // K = (deg+1)^2 active SH coefficients, 0 = colors_precomp
__global__ void __launch_bounds__(128, 3)
preprocess_bwd_kernel(const GsParams p, const GsBackwardIO io)
{
    gs_pdl_wait();                                           // launched as a dependent of the blend backward (grad2d)
    const int t_global = blockIdx.x * blockDim.x + threadIdx.x;
    const int i_raw = t_global / LPG, vq = t_global % LPG;  // Gaussian, view sub-lane (LPG lanes share a Gaussian's views: 4, or 1 when V <= 2)
    const bool act = i_raw < p.N;                            // inactive lanes still take part in the shuffles
    const int i = act ? i_raw : 0;
    const float px = p.means3D[3 * i], py = p.means3D[3 * i + 1], pz = p.means3D[3 * i + 2];
    const bool use_cov = p.cov3D != nullptr;

    // view-independent: cov3D and rotation
    float c3[6], sc[3] = {0.f, 0.f, 0.f}, R[3][3];
    float qr = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
    if (use_cov) {
        #pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = p.cov3D[6 * (size_t)i + k];
    } else {
        sc[0] = p.mod * p.scales[3 * i]; sc[1] = p.mod * p.scales[3 * i + 1]; sc[2] = p.mod * p.scales[3 * i + 2];
        const float4 q = reinterpret_cast<const float4*>(p.rots)[i];
        qr = q.x; qx = q.y; qy = q.z; qz = q.w;
        R[0][0] = 1.f - 2.f * (qy * qy + qz * qz); R[0][1] = 2.f * (qx * qy - qr * qz); R[0][2] = 2.f * (qx * qz + qr * qy);
        R[1][0] = 2.f * (qx * qy + qr * qz); R[1][1] = 1.f - 2.f * (qx * qx + qz * qz); R[1][2] = 2.f * (qy * qz - qr * qx);
        R[2][0] = 2.f * (qx * qz - qr * qy); R[2][1] = 2.f * (qy * qz + qr * qx); R[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
        float A[3][3];
        #pragma unroll
        for (int a = 0; a < 3; a++) { A[a][0] = R[a][0] * sc[0]; A[a][1] = R[a][1] * sc[1]; A[a][2] = R[a][2] * sc[2]; }
        c3[0] = A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2];
        c3[1] = A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2];
        c3[2] = A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2];
        c3[3] = A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2];
        c3[4] = A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2];
        c3[5] = A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2];
    }
    const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};

    float gm[3] = {0.f, 0.f, 0.f}, gm2[2] = {0.f, 0.f}, gop = 0.f, gcol[3] = {0.f, 0.f, 0.f};
    float G3[3][3];
    #pragma unroll
    for (int a = 0; a < 3; a++) { G3[a][0] = G3[a][1] = G3[a][2] = 0.f; }
    constexpr int KK = K > 0 ? K : 1;
    float gsh[KK * 3];
    #pragma unroll
    for (int k = 0; k < KK * 3; k++) gsh[k] = 0.f;

#if GS_PBWD_PREFETCH
    // the per-(view, Gaussian) inputs of the NEXT view of this lane are requested before the arithmetic of the current one (the kernel
    // waits on these loads: long-scoreboard 3.2 per issue at 12 warps per SM)
    int v_next = vq;
    int rad_n = 0;
    unsigned cl_n = 0u;
    float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0, n2 = n0;
    if (act && v_next < p.V) {
        const size_t g = (size_t)v_next * p.N + i;
        rad_n = io.radii[g]; n0 = p.grad2d[g * 3]; n1 = p.grad2d[g * 3 + 1]; n2 = p.grad2d[g * 3 + 2];
        if constexpr (K > 0) cl_n = p.clamped[g];
    }
    while (act && v_next < p.V) {
        const int v = v_next;
        const int rad = rad_n;
        const float4 a0 = n0, a1 = n1, a2 = n2;
        [[maybe_unused]] const unsigned cl = cl_n;
        v_next += LPG;
        if (v_next < p.V) {
            const size_t g = (size_t)v_next * p.N + i;
            rad_n = io.radii[g]; n0 = p.grad2d[g * 3]; n1 = p.grad2d[g * 3 + 1]; n2 = p.grad2d[g * 3 + 2];
            if constexpr (K > 0) cl_n = p.clamped[g];
        }
        if (rad <= 0) continue;
        const float* __restrict__ cam = p.cams + (size_t)v * GS_CAM_FLOATS;
        const float* V = cam + GS_CAM_VIEW;
        const float* P = cam + GS_CAM_PROJ;
        const float g_px = a0.x, g_py = a0.y, gA = a0.z, gB = a0.w, gC = a1.x, g_o = a1.y, g_d = a1.z;
        gop += g_o;

        if constexpr (K > 0) {
#else
    for (int v = vq; v < p.V && act; v += LPG) {
        const size_t gid = (size_t)v * p.N + i;
        if (io.radii[gid] <= 0) continue;
        const float* __restrict__ cam = p.cams + (size_t)v * GS_CAM_FLOATS;
        const float* V = cam + GS_CAM_VIEW;
        const float* P = cam + GS_CAM_PROJ;
        const float4 a0 = p.grad2d[gid * 3], a1 = p.grad2d[gid * 3 + 1], a2 = p.grad2d[gid * 3 + 2];
        const float g_px = a0.x, g_py = a0.y, gA = a0.z, gB = a0.w, gC = a1.x, g_o = a1.y, g_d = a1.z;
        gop += g_o;

        if constexpr (K > 0) {
            const unsigned cl = p.clamped[gid];
#endif
            const float dx = px - cam[GS_CAM_CAMPOS], dy = py - cam[GS_CAM_CAMPOS + 1], dz = pz - cam[GS_CAM_CAMPOS + 2];
            const float inv_len = rsqrtf(dx * dx + dy * dy + dz * dz);
            const float x = dx * inv_len, y = dy * inv_len, z = dz * inv_len;
            float b[KK], bx[KK], by[KK], bz[KK];
            sh_basis<KK>(x, y, z, b, bx, by, bz);
            const float gr[3] = {(cl & 1u) ? 0.f : a2.x, (cl & 2u) ? 0.f : a2.y, (cl & 4u) ? 0.f : a2.z};
            const float* __restrict__ sh = p.shs + (size_t)i * p.M * 3;
            float gdx = 0.f, gdy = 0.f, gdz = 0.f;
            if (((p.M * 3) & 3) == 0 && (KK * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(p.shs) & 15) == 0) {
                const float4* __restrict__ sh4 = reinterpret_cast<const float4*>(sh);
                #pragma unroll
                for (int q4 = 0; q4 < (KK * 3) / 4; q4++) {
                    const float4 tq = __ldg(sh4 + q4);
                    const float tv[4] = {tq.x, tq.y, tq.z, tq.w};
                    #pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int idx = 4 * q4 + e, k = idx / 3, ch = idx % 3;
                        gsh[idx] += b[k] * gr[ch];
                        const float sg = tv[e] * gr[ch];
                        gdx += bx[k] * sg; gdy += by[k] * sg; gdz += bz[k] * sg;
                    }
                }
            } else {
                #pragma unroll
                for (int k = 0; k < KK; k++) {
                    #pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        gsh[k * 3 + ch] += b[k] * gr[ch];
                        const float sg = __ldg(sh + k * 3 + ch) * gr[ch];
                        gdx += bx[k] * sg; gdy += by[k] * sg; gdz += bz[k] * sg;
                    }
                }
            }
            const float dot = gdx * x + gdy * y + gdz * z;
            gm[0] += (gdx - dot * x) * inv_len; gm[1] += (gdy - dot * y) * inv_len; gm[2] += (gdz - dot * z) * inv_len;
        } else {
            gcol[0] += a2.x; gcol[1] += a2.y; gcol[2] += a2.z;
        }

        // forward intermediates
        const float tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
        const float ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
        const float tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
        const float tanx = cam[GS_CAM_TANFOVX], tany = cam[GS_CAM_TANFOVY];
        const float fx = (float)p.W / (2.0f * tanx), fy = (float)p.H / (2.0f * tany);
        const float limx = 1.3f * tanx, limy = 1.3f * tany;
        const float itz = 1.0f / tz;
        const float txtz = tx * itz, tytz = ty * itz;
        const float mx = (txtz < -limx || txtz > limx) ? 0.f : 1.f, my = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float cx = fminf(limx, fmaxf(-limx, txtz)) * tz, cy = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float itz2 = itz * itz, itz3 = itz2 * itz;
        const float J00 = fx * itz, J02 = -fx * cx * itz2, J11 = fy * itz, J12 = -fy * cy * itz2;
        float T[2][3], TS[2][3];
        #pragma unroll
        for (int j = 0; j < 3; j++) {
            T[0][j] = J00 * V[4 * j + 0] + J02 * V[4 * j + 2];
            T[1][j] = J11 * V[4 * j + 1] + J12 * V[4 * j + 2];
        }
        #pragma unroll
        for (int r = 0; r < 2; r++)
            #pragma unroll
            for (int j = 0; j < 3; j++) TS[r][j] = T[r][0] * S[0][j] + T[r][1] * S[1][j] + T[r][2] * S[2][j];
        const float ca = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2] + GS_LOWPASS;
        const float cb = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
        const float cc = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2] + GS_LOWPASS;
        const float det = ca * cc - cb * cb;
        const float d2 = 1.0f / (det * det + 0.0000001f);
        const float da = d2 * (-cc * cc * gA + cb * cc * gB - cb * cb * gC);
        const float db = d2 * (2.f * cb * cc * gA - (det + 2.f * cb * cb) * gB + 2.f * ca * cb * gC);
        const float dc = d2 * (-cb * cb * gA + ca * cb * gB - ca * ca * gC);
        const float g2[2][2] = {{da, 0.5f * db}, {0.5f * db, dc}};
        // dL/dSigma (full symmetric) += T^T g2 T
        float g2T[2][3];
        #pragma unroll
        for (int r = 0; r < 2; r++)
            #pragma unroll
            for (int j = 0; j < 3; j++) g2T[r][j] = g2[r][0] * T[0][j] + g2[r][1] * T[1][j];
        #pragma unroll
        for (int r = 0; r < 3; r++)
            #pragma unroll
            for (int c = 0; c < 3; c++) G3[r][c] += T[0][r] * g2T[0][c] + T[1][r] * g2T[1][c];
        // dL/dT = 2 g2 T S -> dL/dJ -> dL/dt
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
        #pragma unroll
        for (int j = 0; j < 3; j++) {
            const float dT0 = 2.f * (g2[0][0] * TS[0][j] + g2[0][1] * TS[1][j]);
            const float dT1 = 2.f * (g2[1][0] * TS[0][j] + g2[1][1] * TS[1][j]);
            dJ00 += dT0 * V[4 * j + 0]; dJ02 += dT0 * V[4 * j + 2];
            dJ11 += dT1 * V[4 * j + 1]; dJ12 += dT1 * V[4 * j + 2];
        }
        const float dtx = mx * (-fx * itz2) * dJ02;
        const float dty = my * (-fy * itz2) * dJ12;
        float dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + 2.f * fx * cx * itz3 * dJ02 + 2.f * fy * cy * itz3 * dJ12;
        dtz += g_d;
        // perspective divide of the full projection
        const float hx = P[0] * px + P[4] * py + P[8] * pz + P[12];
        const float hy = P[1] * px + P[5] * py + P[9] * pz + P[13];
        const float hw = P[3] * px + P[7] * py + P[11] * pz + P[15];
        const float pw = 1.0f / (hw + 0.0000001f);
        const float gxn = g_px * 0.5f * (float)p.W, gyn = g_py * 0.5f * (float)p.H;
        gm2[0] += gxn; gm2[1] += gyn;
        #pragma unroll
        for (int j = 0; j < 3; j++) {
            gm[j] += V[4 * j + 0] * dtx + V[4 * j + 1] * dty + V[4 * j + 2] * dtz
                   + (P[4 * j + 0] * pw - P[4 * j + 3] * hx * pw * pw) * gxn
                   + (P[4 * j + 1] * pw - P[4 * j + 3] * hy * pw * pw) * gyn;
        }
    }

    // combine the four view sub-lanes (lanes 4j..4j+3 of a warp), then lane 0 of the quad writes
    #pragma unroll
    for (int o = 1; o < LPG; o <<= 1) {
        #pragma unroll
        for (int k = 0; k < 3; k++) { gm[k] += __shfl_xor_sync(0xffffffffu, gm[k], o); gcol[k] += __shfl_xor_sync(0xffffffffu, gcol[k], o); }
        gm2[0] += __shfl_xor_sync(0xffffffffu, gm2[0], o); gm2[1] += __shfl_xor_sync(0xffffffffu, gm2[1], o);
        gop += __shfl_xor_sync(0xffffffffu, gop, o);
        #pragma unroll
        for (int a = 0; a < 3; a++)
            #pragma unroll
            for (int c = 0; c < 3; c++) G3[a][c] += __shfl_xor_sync(0xffffffffu, G3[a][c], o);
        if constexpr (K > 0) {
            #pragma unroll
            for (int k = 0; k < KK * 3; k++) gsh[k] += __shfl_xor_sync(0xffffffffu, gsh[k], o);
        }
    }
    if (!act || vq != 0) return;
    io.dL_dmeans3D[3 * i] = gm[0]; io.dL_dmeans3D[3 * i + 1] = gm[1]; io.dL_dmeans3D[3 * i + 2] = gm[2];
    io.dL_dmeans2D[3 * i] = gm2[0]; io.dL_dmeans2D[3 * i + 1] = gm2[1]; io.dL_dmeans2D[3 * i + 2] = 0.f;
    io.dL_dopacities[i] = gop;
    if constexpr (K > 0) {
        float* __restrict__ o = io.dL_dshs + (size_t)i * p.M * 3;
        #pragma unroll
        for (int k = 0; k < KK * 3; k++) o[k] = gsh[k];
        for (int k = KK * 3; k < p.M * 3; k++) o[k] = 0.f;
    } else {
        io.dL_dcolors[3 * i] = gcol[0]; io.dL_dcolors[3 * i + 1] = gcol[1]; io.dL_dcolors[3 * i + 2] = gcol[2];
    }
    if (use_cov) {
        float* __restrict__ o = io.dL_dcov3D + 6 * (size_t)i;
        o[0] = G3[0][0]; o[1] = 2.f * G3[0][1]; o[2] = 2.f * G3[0][2];
        o[3] = G3[1][1]; o[4] = 2.f * G3[1][2]; o[5] = G3[2][2];
    } else {
        // Sigma = R D R^T, D = diag((mod*scale)^2)
        float GR[3][3];
        #pragma unroll
        for (int a = 0; a < 3; a++)
            #pragma unroll
            for (int k = 0; k < 3; k++) GR[a][k] = G3[a][0] * R[0][k] + G3[a][1] * R[1][k] + G3[a][2] * R[2][k];
        float Hm[3][3];
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            const float rgr = R[0][k] * GR[0][k] + R[1][k] * GR[1][k] + R[2][k] * GR[2][k];
            io.dL_dscales[3 * i + k] = 2.f * sc[k] * p.mod * rgr;
            #pragma unroll
            for (int a = 0; a < 3; a++) Hm[a][k] = 2.f * GR[a][k] * sc[k] * sc[k];
        }
        float4 gq;
        gq.x = 2.f * (-qz * Hm[0][1] + qy * Hm[0][2] + qz * Hm[1][0] - qx * Hm[1][2] - qy * Hm[2][0] + qx * Hm[2][1]);
        gq.y = 2.f * (qy * Hm[0][1] + qz * Hm[0][2] + qy * Hm[1][0] - 2.f * qx * Hm[1][1] - qr * Hm[1][2] + qz * Hm[2][0] + qr * Hm[2][1] - 2.f * qx * Hm[2][2]);
        gq.z = 2.f * (-2.f * qy * Hm[0][0] + qx * Hm[0][1] + qr * Hm[0][2] + qx * Hm[1][0] + qz * Hm[1][2] - qr * Hm[2][0] + qz * Hm[2][1] - 2.f * qy * Hm[2][2]);
        gq.w = 2.f * (-2.f * qz * Hm[0][0] - qr * Hm[0][1] + qx * Hm[0][2] + qr * Hm[1][0] - 2.f * qz * Hm[1][1] + qy * Hm[1][2] + qx * Hm[2][0] + qy * Hm[2][1]);
        reinterpret_cast<float4*>(io.dL_drotations)[i] = gq;
    }
}

}  // namespace

template <int LPG>
static void launch_pbwd(const GsParams& p, const GsBackwardIO& io, cudaStream_t s)
{
    const int threads = 128, blocks = (int)(((long long)p.N * LPG + threads - 1) / threads);
    if (!p.shs) { gs_launch_dependent(preprocess_bwd_kernel<0, LPG>, dim3(blocks), dim3(threads), 0, s, p, io); return; }
    switch (p.deg) {
        case 0: gs_launch_dependent(preprocess_bwd_kernel<1, LPG>, dim3(blocks), dim3(threads), 0, s, p, io); break;
        case 1: gs_launch_dependent(preprocess_bwd_kernel<4, LPG>, dim3(blocks), dim3(threads), 0, s, p, io); break;
        case 2: gs_launch_dependent(preprocess_bwd_kernel<9, LPG>, dim3(blocks), dim3(threads), 0, s, p, io); break;
        default: gs_launch_dependent(preprocess_bwd_kernel<16, LPG>, dim3(blocks), dim3(threads), 0, s, p, io); break;
    }
}

void gs_launch_preprocess_bwd(const GsParams& p, const GsBackwardIO& io, cudaStream_t s)
{
    if (p.N == 0) return;
    // four lanes per Gaussian split the views of a multi-view launch; with one or two views (the reference renders one view per
    // step, train.py:307) three of them would idle, so a Gaussian gets a single lane
    if (p.V >= 3) launch_pbwd<4>(p, io, s); else launch_pbwd<1>(p, io, s);
}
