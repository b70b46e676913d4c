/*
 * gs_oracle.c -- CPU restatement of the differentiable Gaussian-splatting rasterizer
 *                that Topo4D calls through `diff_gaussian_rasterization`.
 *
 * THIS FILE IS TEST INFRASTRUCTURE (the parity checker and the reported CPU baseline).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (topo4d_b200/) never does.
 *
 * PARITY UNPINNED: the algorithm lives in an un-vendored third-party dependency of the
 * reference -- `ashawkey/diff-gaussian-rasterization` (fork of
 * graphdeco-inria/diff-gaussian-rasterization, "+ depth, alpha rendering"), cloned at an
 * unpinned default-branch HEAD by the reference's README.md:22-24 and absent from
 * /root/reference (empty dir diff-gaussian-rasterization-w-depth/, .SUBMODULES.json:8).
 * The reference ships no tests or golden images for this path.  This file therefore
 * restates the *published* algorithm (Kerbl et al. 2023 tile-based EWA splatting with
 * the fork's depth/alpha outputs) and is anchored on the
 * reference's own call sites and conventions:
 *   - settings tuple / matrix layout .......... helpers.py:63-88   (setup_camera)
 *   - op inputs, activations outside the op ... helpers.py:91-112  (params2rendervar[_dense])
 *   - outputs consumed (color, radii) ......... train.py:307-311, 374-376, 388-390
 *   - quaternion (w,x,y,z) -> R convention .... external.py:26-43  (build_rotation)
 *   - SH constants / polynomial order ......... helpers.py:836-922 (C0..C3, eval_sh)
 * Known-answer vectors derived from those reference functions are committed under
 * tests/golden/ (see tests/golden/make_golden.py) and checked by tests/test_oracle_*.py.
 * The arithmetic is additionally cross-checked against an independent dense fp64
 * autograd formulation (oracle/gs_dense_ref.py) and central finite differences.
 *
 * Arithmetic contract (what makes tile/bin indices bit-exact vs the CUDA path):
 *   every expression in preprocess_one() is evaluated in IEEE fp32, left to right, with
 *   NO fused multiply-add (compile with -ffp-contract=off; the CUDA preprocess kernel is
 *   compiled with --fmad=false), IEEE division and sqrtf.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define NEAR_CULL 0.2f
#define LOWPASS 0.3f
#define ALPHA_CAP 0.99f
#define ALPHA_MIN (1.0f / 255.0f)
#define T_MIN 0.0001f

/* SH constants: helpers.py:836-853 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    /* problem */
    int N, H, W, gx, gy, M, deg;
    int use_sh, use_cov_precomp;
    float tanfovx, tanfovy, mod;
    float view[16], proj[16], campos[3], bg[3];
    /* copies of the inputs (needed by backward) */
    float *means, *shs, *colors_in, *opac, *scales, *rots, *cov_in;
    /* geometry state, SURVEY.md A.9 */
    float *depth;      /* [N] view-space z */
    float *xy;         /* [N,2] pixel centre */
    float *conic_o;    /* [N,4] conic A,B,C + opacity */
    float *rgb;        /* [N,3] */
    float *cov3d;      /* [N,6] */
    int *radii;        /* [N] */
    int *rect;         /* [N,4] minx,miny,maxx,maxy (tile units, max exclusive) */
    uint32_t *tiles_touched; /* [N] */
    uint8_t *clamped;  /* [N,3] */
    /* binning state */
    int64_t I;
    uint64_t *keys;    /* [I] sorted */
    uint32_t *vals;    /* [I] sorted gaussian ids */
    uint32_t *ranges;  /* [tiles,2] */
    /* image state */
    float *final_T;    /* [H*W] */
    uint32_t *n_contrib; /* [H*W] */
} GsoState;

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* --- A.3 preprocess for one Gaussian.  Returns 1 if it produces instances. --- */
static int preprocess_one(GsoState* s, int i)
{
    const float* V = s->view; const float* P = s->proj;
    s->radii[i] = 0; s->tiles_touched[i] = 0;
    s->rect[4*i+0] = s->rect[4*i+1] = s->rect[4*i+2] = s->rect[4*i+3] = 0;
    const float px = s->means[3*i], py = s->means[3*i+1], pz = s->means[3*i+2];

    /* view-space point, A.2 (row-vector convention: element [4*col+row]) */
    float tx = V[0]*px + V[4]*py + V[8]*pz + V[12];
    float ty = V[1]*px + V[5]*py + V[9]*pz + V[13];
    float tz = V[2]*px + V[6]*py + V[10]*pz + V[14];
    if (!(tz > NEAR_CULL)) return 0;                     /* near cull (also drops NaN) */

    float hx = P[0]*px + P[4]*py + P[8]*pz + P[12];
    float hy = P[1]*px + P[5]*py + P[9]*pz + P[13];
    float hw = P[3]*px + P[7]*py + P[11]*pz + P[15];
    float pw = 1.0f / (hw + 0.0000001f);
    float ppx = hx * pw, ppy = hy * pw;

    /* 3D covariance (6 upper-triangular floats) */
    float c3[6];
    if (s->use_cov_precomp) {
        for (int k = 0; k < 6; k++) c3[k] = s->cov_in[6*i+k];
    } else {
        float sx = s->mod * s->scales[3*i], sy = s->mod * s->scales[3*i+1], sz = s->mod * s->scales[3*i+2];
        float r = s->rots[4*i], x = s->rots[4*i+1], y = s->rots[4*i+2], z = s->rots[4*i+3];
        /* external.py:34-42 convention, no renormalisation inside the op */
        float R[3][3] = {
            {1.f - 2.f*(y*y + z*z), 2.f*(x*y - r*z),       2.f*(x*z + r*y)},
            {2.f*(x*y + r*z),       1.f - 2.f*(x*x + z*z), 2.f*(y*z - r*x)},
            {2.f*(x*z - r*y),       2.f*(y*z + r*x),       1.f - 2.f*(x*x + y*y)}};
        float A[3][3];
        for (int a = 0; a < 3; a++) { A[a][0] = R[a][0]*sx; A[a][1] = R[a][1]*sy; A[a][2] = R[a][2]*sz; }
        c3[0] = A[0][0]*A[0][0] + A[0][1]*A[0][1] + A[0][2]*A[0][2];
        c3[1] = A[0][0]*A[1][0] + A[0][1]*A[1][1] + A[0][2]*A[1][2];
        c3[2] = A[0][0]*A[2][0] + A[0][1]*A[2][1] + A[0][2]*A[2][2];
        c3[3] = A[1][0]*A[1][0] + A[1][1]*A[1][1] + A[1][2]*A[1][2];
        c3[4] = A[1][0]*A[2][0] + A[1][1]*A[2][1] + A[1][2]*A[2][2];
        c3[5] = A[2][0]*A[2][0] + A[2][1]*A[2][1] + A[2][2]*A[2][2];
    }
    for (int k = 0; k < 6; k++) s->cov3d[6*i+k] = c3[k];

    /* 2D covariance: EWA Jacobian with frustum clamp at 1.3*tanfov */
    const float fx = (float)s->W / (2.0f * s->tanfovx), fy = (float)s->H / (2.0f * s->tanfovy);
    const float limx = 1.3f * s->tanfovx, limy = 1.3f * s->tanfovy;
    float txtz = tx / tz, tytz = ty / tz;
    float cx = fminf_(limx, fmaxf_(-limx, txtz)) * tz;
    float cy = fminf_(limy, fmaxf_(-limy, tytz)) * tz;
    float J00 = fx / tz, J02 = -(fx * cx) / (tz * tz);
    float J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
    /* Wr[i][j] = V[4*j+i] ; T = J * Wr (2x3) */
    float T0[3], T1[3];
    for (int j = 0; j < 3; j++) {
        T0[j] = J00 * V[4*j+0] + J02 * V[4*j+2];
        T1[j] = J11 * V[4*j+1] + J12 * V[4*j+2];
    }
    float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float X0[3], X1[3];
    for (int j = 0; j < 3; j++) {
        X0[j] = T0[0]*S[0][j] + T0[1]*S[1][j] + T0[2]*S[2][j];
        X1[j] = T1[0]*S[0][j] + T1[1]*S[1][j] + T1[2]*S[2][j];
    }
    float a = (X0[0]*T0[0] + X0[1]*T0[1] + X0[2]*T0[2]) + LOWPASS;
    float b =  X0[0]*T1[0] + X0[1]*T1[1] + X0[2]*T1[2];
    float c = (X1[0]*T1[0] + X1[1]*T1[1] + X1[2]*T1[2]) + LOWPASS;

    float det = a * c - b * b;
    if (!(det != 0.0f)) return 0;                         /* det==0 or NaN */
    float det_inv = 1.0f / det;
    float cA = c * det_inv, cB = -b * det_inv, cC = a * det_inv;
    float mid = 0.5f * (a + c);
    float sq = sqrtf(fmaxf_(0.1f, mid * mid - det));
    float l1 = mid + sq, l2 = mid - sq;
    float radius = ceilf(3.0f * sqrtf(fmaxf_(l1, l2)));
    if (!(radius < 1.0e9f)) return 0;                     /* inf/NaN guard */
    float pix_x = ((ppx + 1.0f) * (float)s->W - 1.0f) * 0.5f;
    float pix_y = ((ppy + 1.0f) * (float)s->H - 1.0f) * 0.5f;
    if (!(fabsf(pix_x) < 1.0e9f) || !(fabsf(pix_y) < 1.0e9f)) return 0;

    /* tile rect: C truncation toward zero, then clamp */
    int minx = (int)((pix_x - radius) / (float)TILE);
    int miny = (int)((pix_y - radius) / (float)TILE);
    int maxx = (int)((pix_x + radius + (float)(TILE - 1)) / (float)TILE);
    int maxy = (int)((pix_y + radius + (float)(TILE - 1)) / (float)TILE);
    minx = minx < 0 ? 0 : (minx > s->gx ? s->gx : minx);
    miny = miny < 0 ? 0 : (miny > s->gy ? s->gy : miny);
    maxx = maxx < 0 ? 0 : (maxx > s->gx ? s->gx : maxx);
    maxy = maxy < 0 ? 0 : (maxy > s->gy ? s->gy : maxy);
    if ((maxx - minx) * (maxy - miny) == 0) return 0;

    /* colour */
    float rgb[3];
    if (s->use_sh) {
        const float* sh = s->shs + (size_t)i * s->M * 3;
        float dx = px - s->campos[0], dy = py - s->campos[1], dz = pz - s->campos[2];
        float len = sqrtf(dx*dx + dy*dy + dz*dz);
        float x = dx / len, y = dy / len, z = dz / len;
        for (int ch = 0; ch < 3; ch++) {
            float res = SH_C0 * sh[0*3+ch];
            if (s->deg > 0) {
                res = res - SH_C1 * y * sh[1*3+ch] + SH_C1 * z * sh[2*3+ch] - SH_C1 * x * sh[3*3+ch];
                if (s->deg > 1) {
                    float xx = x*x, yy = y*y, zz = z*z, xy = x*y, yz = y*z, xz = x*z;
                    res = res + SH_C2[0] * xy * sh[4*3+ch] + SH_C2[1] * yz * sh[5*3+ch]
                              + SH_C2[2] * (2.0f*zz - xx - yy) * sh[6*3+ch]
                              + SH_C2[3] * xz * sh[7*3+ch] + SH_C2[4] * (xx - yy) * sh[8*3+ch];
                    if (s->deg > 2) {
                        res = res + SH_C3[0] * y * (3.0f*xx - yy) * sh[9*3+ch]
                                  + SH_C3[1] * xy * z * sh[10*3+ch]
                                  + SH_C3[2] * y * (4.0f*zz - xx - yy) * sh[11*3+ch]
                                  + SH_C3[3] * z * (2.0f*zz - 3.0f*xx - 3.0f*yy) * sh[12*3+ch]
                                  + SH_C3[4] * x * (4.0f*zz - xx - yy) * sh[13*3+ch]
                                  + SH_C3[5] * z * (xx - yy) * sh[14*3+ch]
                                  + SH_C3[6] * x * (xx - 3.0f*yy) * sh[15*3+ch];
                    }
                }
            }
            res += 0.5f;
            s->clamped[3*i+ch] = (res < 0.0f);
            rgb[ch] = res < 0.0f ? 0.0f : res;
        }
    } else {
        for (int ch = 0; ch < 3; ch++) { rgb[ch] = s->colors_in[3*i+ch]; s->clamped[3*i+ch] = 0; }
    }

    s->depth[i] = tz;
    s->radii[i] = (int)radius;
    s->xy[2*i] = pix_x; s->xy[2*i+1] = pix_y;
    s->conic_o[4*i] = cA; s->conic_o[4*i+1] = cB; s->conic_o[4*i+2] = cC; s->conic_o[4*i+3] = s->opac[i];
    s->rgb[3*i] = rgb[0]; s->rgb[3*i+1] = rgb[1]; s->rgb[3*i+2] = rgb[2];
    s->rect[4*i] = minx; s->rect[4*i+1] = miny; s->rect[4*i+2] = maxx; s->rect[4*i+3] = maxy;
    s->tiles_touched[i] = (uint32_t)((maxx - minx) * (maxy - miny));
    return 1;
}

/* stable LSD radix sort of (key,val) on the low `bits` bits */
static void radix_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int bits)
{
    if (n <= 1) return;
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    uint32_t* v2 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint64_t *ka = keys, *kb = k2; uint32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < bits; shift += 8) {
        int64_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; i++) cnt[((ka[i] >> shift) & 0xff) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d+1] += cnt[d];
        for (int64_t i = 0; i < n; i++) { int64_t p = cnt[(ka[i] >> shift) & 0xff]++; kb[p] = ka[i]; vb[p] = va[i]; }
        uint64_t* tk = ka; ka = kb; kb = tk; uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, sizeof(uint64_t) * (size_t)n); memcpy(vals, va, sizeof(uint32_t) * (size_t)n); }
    free(k2); free(v2);
}

/* --- A.5 forward blend of one tile --- */
static void blend_tile(const GsoState* s, int tile, float* out_color, float* out_depth, float* out_alpha)
{
    const int tx0 = (tile % s->gx) * TILE, ty0 = (tile / s->gx) * TILE;
    const uint32_t beg = s->ranges[2*tile], end = s->ranges[2*tile+1];
    const size_t HW = (size_t)s->H * s->W;
    for (int ly = 0; ly < TILE; ly++) for (int lx = 0; lx < TILE; lx++) {
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= s->W || y >= s->H) continue;
        const float pxf = (float)x, pyf = (float)y;       /* integer pixel coords, no +0.5 */
        float T = 1.0f, C[3] = {0, 0, 0}, D = 0.0f, Wt = 0.0f;
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = beg; k < end; k++) {
            contributor++;
            const uint32_t g = s->vals[k];
            const float dx = s->xy[2*g] - pxf, dy = s->xy[2*g+1] - pyf;
            const float* co = s->conic_o + 4*g;
            const float power = -0.5f * (co[0]*dx*dx + co[2]*dy*dy) - co[1]*dx*dy;
            if (power > 0.0f) continue;
            const float alpha = fminf_(ALPHA_CAP, co[3] * expf(power));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < T_MIN) break;                    /* this Gaussian is NOT blended */
            const float w = alpha * T;
            C[0] += s->rgb[3*g] * w; C[1] += s->rgb[3*g+1] * w; C[2] += s->rgb[3*g+2] * w;
            D += s->depth[g] * w;
            Wt += w;
            T = test_T;
            last = contributor;
        }
        const size_t pix = (size_t)y * s->W + x;
        s->final_T[pix] = T; s->n_contrib[pix] = last;
        out_color[0*HW + pix] = C[0] + T * s->bg[0];
        out_color[1*HW + pix] = C[1] + T * s->bg[1];
        out_color[2*HW + pix] = C[2] + T * s->bg[2];
        out_depth[pix] = D;
        out_alpha[pix] = Wt;
    }
}

static float* dupf(const float* p, size_t n) { if (!p) return NULL; float* q = (float*)malloc(n * sizeof(float) + 16); memcpy(q, p, n * sizeof(float)); return q; }

void gso_free(GsoState* s)
{
    if (!s) return;
    free(s->means); free(s->shs); free(s->colors_in); free(s->opac); free(s->scales); free(s->rots); free(s->cov_in);
    free(s->depth); free(s->xy); free(s->conic_o); free(s->rgb); free(s->cov3d); free(s->radii); free(s->rect);
    free(s->tiles_touched); free(s->clamped); free(s->keys); free(s->vals); free(s->ranges); free(s->final_T); free(s->n_contrib);
    free(s);
}

/*
 * Forward.  Interface mirrors the op's kwargs (helpers.py:91-100) + settings (helpers.py:73-86).
 * shs is [N,M,3] or NULL; colors_precomp [N,3] or NULL; scales/rots or cov3D_precomp [N,6].
 * Outputs: color [3,H,W], depth [H,W], alpha [H,W], radii [N].  Returns an opaque state.
 */
GsoState* gso_forward(int N, int M, int deg, int H, int W,
                      const float* means3D, const float* shs, const float* colors_precomp,
                      const float* opacities, const float* scales, const float* rotations, const float* cov3D_precomp,
                      float scale_modifier, const float* viewmatrix, const float* projmatrix, const float* campos,
                      float tanfovx, float tanfovy, const float* bg,
                      float* out_color, float* out_depth, float* out_alpha, int* out_radii)
{
    GsoState* s = (GsoState*)calloc(1, sizeof(GsoState));
    s->N = N; s->H = H; s->W = W; s->M = M; s->deg = deg;
    s->gx = (W + TILE - 1) / TILE; s->gy = (H + TILE - 1) / TILE;
    s->use_sh = shs != NULL; s->use_cov_precomp = cov3D_precomp != NULL;
    s->tanfovx = tanfovx; s->tanfovy = tanfovy; s->mod = scale_modifier;
    memcpy(s->view, viewmatrix, 64); memcpy(s->proj, projmatrix, 64); memcpy(s->campos, campos, 12); memcpy(s->bg, bg, 12);
    s->means = dupf(means3D, (size_t)N*3); s->shs = dupf(shs, (size_t)N*M*3); s->colors_in = dupf(colors_precomp, (size_t)N*3);
    s->opac = dupf(opacities, N); s->scales = dupf(scales, (size_t)N*3); s->rots = dupf(rotations, (size_t)N*4);
    s->cov_in = dupf(cov3D_precomp, (size_t)N*6);
    size_t n1 = (size_t)(N > 0 ? N : 1);
    s->depth = (float*)calloc(n1, 4); s->xy = (float*)calloc(n1*2, 4); s->conic_o = (float*)calloc(n1*4, 4);
    s->rgb = (float*)calloc(n1*3, 4); s->cov3d = (float*)calloc(n1*6, 4); s->radii = (int*)calloc(n1, 4);
    s->rect = (int*)calloc(n1*4, 4); s->tiles_touched = (uint32_t*)calloc(n1, 4); s->clamped = (uint8_t*)calloc(n1*3, 1);
    const int tiles = s->gx * s->gy;
    const size_t HW = (size_t)H * W;
    s->ranges = (uint32_t*)calloc((size_t)tiles*2 + 2, 4);
    s->final_T = (float*)calloc(HW + 1, 4); s->n_contrib = (uint32_t*)calloc(HW + 1, 4);

    #pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) preprocess_one(s, i);

    /* A.4 binning: inclusive scan, duplicate with keys (y outer, x inner), stable sort, ranges */
    int64_t I = 0;
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * n1);
    for (int i = 0; i < N; i++) { offs[i] = I; I += s->tiles_touched[i]; }
    s->I = I;
    s->keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(I > 0 ? I : 1));
    s->vals = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(I > 0 ? I : 1));
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        if (s->radii[i] <= 0) continue;
        int64_t o = offs[i];
        for (int y = s->rect[4*i+1]; y < s->rect[4*i+3]; y++)
            for (int x = s->rect[4*i]; x < s->rect[4*i+2]; x++) {
                uint64_t key = (uint64_t)(y * s->gx + x);
                key = (key << 32) | f2u(s->depth[i]);
                s->keys[o] = key; s->vals[o] = (uint32_t)i; o++;
            }
    }
    free(offs);
    int tbits = 0; while ((1 << tbits) < tiles) tbits++;
    radix_sort_pairs(s->keys, s->vals, I, 32 + tbits + 1);
    for (int64_t k = 0; k < I; k++) {
        uint32_t t = (uint32_t)(s->keys[k] >> 32);
        if (k == 0 || t != (uint32_t)(s->keys[k-1] >> 32)) s->ranges[2*t] = (uint32_t)k;
        if (k == I - 1 || t != (uint32_t)(s->keys[k+1] >> 32)) s->ranges[2*t+1] = (uint32_t)(k + 1);
    }

    if (N == 0) {   /* upstream returns zero images (not bg) for an empty scene */
        memset(out_color, 0, HW*3*4); memset(out_depth, 0, HW*4); memset(out_alpha, 0, HW*4);
        for (size_t p = 0; p < HW; p++) s->final_T[p] = 1.0f;
    } else {
        #pragma omp parallel for schedule(dynamic, 4)
        for (int t = 0; t < tiles; t++) blend_tile(s, t, out_color, out_depth, out_alpha);
    }
    if (out_radii) memcpy(out_radii, s->radii, sizeof(int) * (size_t)N);
    return s;
}

/* ---- state accessors (index structures for the bit-exact checks) ---- */
int64_t gso_num_rendered(const GsoState* s) { return s->I; }
void gso_get_geometry(const GsoState* s, float* depth, float* xy, float* conic_o, float* rgb, float* cov3d,
                      int* rect, uint32_t* tiles_touched, uint8_t* clamped)
{
    size_t N = (size_t)s->N;
    if (depth) memcpy(depth, s->depth, N*4);
    if (xy) memcpy(xy, s->xy, N*8);
    if (conic_o) memcpy(conic_o, s->conic_o, N*16);
    if (rgb) memcpy(rgb, s->rgb, N*12);
    if (cov3d) memcpy(cov3d, s->cov3d, N*24);
    if (rect) memcpy(rect, s->rect, N*16);
    if (tiles_touched) memcpy(tiles_touched, s->tiles_touched, N*4);
    if (clamped) memcpy(clamped, s->clamped, N*3);
}
void gso_get_binning(const GsoState* s, uint64_t* keys, uint32_t* vals, uint32_t* ranges)
{
    if (keys) memcpy(keys, s->keys, (size_t)s->I * 8);
    if (vals) memcpy(vals, s->vals, (size_t)s->I * 4);
    if (ranges) memcpy(ranges, s->ranges, (size_t)s->gx * s->gy * 8);
}
void gso_get_image_state(const GsoState* s, float* final_T, uint32_t* n_contrib)
{
    size_t HW = (size_t)s->H * s->W;
    if (final_T) memcpy(final_T, s->final_T, HW*4);
    if (n_contrib) memcpy(n_contrib, s->n_contrib, HW*4);
}

/* --- A.6 backward blend of one tile, REFERENCE-FAITHFUL fp32 replay (every value a float, as the CUDA
 *     reference computes it; only the cross-pixel accumulation is double).  Used to measure how far the fp32
 *     algorithm itself sits from the accurate gradient (blend_tile_bwd below), i.e. the noise floor any fp32
 *     implementation of this path has in ill-conditioned regimes (alpha at the 0.99 cap). */
static void blend_tile_bwd_f32(const GsoState* s, int tile, const float* gC, const float* gD, const float* gA, double* acc)
{
    const int tx0 = (tile % s->gx) * TILE, ty0 = (tile / s->gx) * TILE;
    const uint32_t beg = s->ranges[2*tile];
    const size_t HW = (size_t)s->H * s->W;
    for (int ly = 0; ly < TILE; ly++) for (int lx = 0; lx < TILE; lx++) {
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= s->W || y >= s->H) continue;
        const size_t pix = (size_t)y * s->W + x;
        const float pxf = (float)x, pyf = (float)y;
        const float T_final = s->final_T[pix];
        const uint32_t last = s->n_contrib[pix];
        const float g_c[3] = {gC[pix], gC[HW + pix], gC[2*HW + pix]};
        const float g_d = gD[pix], g_a = gA[pix];
        const float bg_dot = s->bg[0]*g_c[0] + s->bg[1]*g_c[1] + s->bg[2]*g_c[2];
        float T = T_final;
        float rec_c[3] = {0, 0, 0}, rec_d = 0, rec_a = 0;
        float last_alpha = 0, last_c[3] = {0, 0, 0}, last_d = 0;
        for (int64_t k = (int64_t)beg + last - 1; k >= (int64_t)beg; k--) {
            const uint32_t g = s->vals[k];
            const float dx = s->xy[2*g] - pxf, dy = s->xy[2*g+1] - pyf;
            const float* co = s->conic_o + 4*g;
            const float power = -0.5f * (co[0]*dx*dx + co[2]*dy*dy) - co[1]*dx*dy;
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf_(ALPHA_CAP, co[3] * G);
            if (alpha < ALPHA_MIN) continue;
            T = T / (1.0f - alpha);
            const float w = alpha * T;
            double* a = acc + (size_t)g * 10;
            float dL_dalpha = 0.0f;
            for (int ch = 0; ch < 3; ch++) {
                const float c = s->rgb[3*g+ch];
                rec_c[ch] = last_alpha * last_c[ch] + (1.0f - last_alpha) * rec_c[ch];
                last_c[ch] = c;
                dL_dalpha += (c - rec_c[ch]) * g_c[ch];
                a[6+ch] += (double)(w * g_c[ch]);
            }
            const float dep = s->depth[g];
            rec_d = last_alpha * last_d + (1.0f - last_alpha) * rec_d;
            last_d = dep;
            dL_dalpha += (dep - rec_d) * g_d;
            a[9] += (double)(w * g_d);
            rec_a = last_alpha + (1.0f - last_alpha) * rec_a;
            dL_dalpha += (1.0f - rec_a) * g_a;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
            const float dL_dG = co[3] * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            a[0] += (double)(dL_dG * (-gdx * co[0] - gdy * co[1]));
            a[1] += (double)(dL_dG * (-gdy * co[2] - gdx * co[1]));
            a[2] += (double)(-0.5f * gdx * dx * dL_dG);
            a[3] += (double)(-gdx * dy * dL_dG);
            a[4] += (double)(-0.5f * gdy * dy * dL_dG);
            a[5] += (double)(G * dL_dalpha);
        }
    }
}

/* --- A.6 backward blend of one tile into per-thread double accumulators acc[N][10]:
 *     0,1: dL/dpix (pixel units)  2,3,4: dL/d(conic A, B(true), C)  5: dL/dopacity
 *     6,7,8: dL/drgb  9: dL/ddepth */
static void blend_tile_bwd(const GsoState* s, int tile, const float* gC, const float* gD, const float* gA, double* acc)
{
    const int tx0 = (tile % s->gx) * TILE, ty0 = (tile / s->gx) * TILE;
    const uint32_t beg = s->ranges[2*tile];
    const size_t HW = (size_t)s->H * s->W;
    for (int ly = 0; ly < TILE; ly++) for (int lx = 0; lx < TILE; lx++) {
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= s->W || y >= s->H) continue;
        const size_t pix = (size_t)y * s->W + x;
        const float pxf = (float)x, pyf = (float)y;
        const float T_final = s->final_T[pix];
        const uint32_t last = s->n_contrib[pix];
        const float g_c[3] = {gC[pix], gC[HW + pix], gC[2*HW + pix]};
        const float g_d = gD[pix], g_a = gA[pix];
        const float bg_dot = s->bg[0]*g_c[0] + s->bg[1]*g_c[1] + s->bg[2]*g_c[2];
        /* Which Gaussians take part is decided exactly as the fp32 forward decided it (float power /
         * alpha tests, n_contrib); the VALUES of the replay are carried in double so that the oracle is
         * the accurate gradient of that forward and not one more fp32 rounding of it (the T = T/(1-alpha)
         * recovery amplifies fp32 noise by alpha/(1-alpha) per layer, up to 100x at the 0.99 cap). */
        double T = (double)T_final;
        double rec_c[3] = {0, 0, 0}, rec_d = 0, rec_a = 0;
        double last_alpha = 0, last_c[3] = {0, 0, 0}, last_d = 0;
        for (int64_t k = (int64_t)beg + last - 1; k >= (int64_t)beg; k--) {
            const uint32_t g = s->vals[k];
            const float dxf = s->xy[2*g] - pxf, dyf = s->xy[2*g+1] - pyf;
            const float* co = s->conic_o + 4*g;
            const float power_f = -0.5f * (co[0]*dxf*dxf + co[2]*dyf*dyf) - co[1]*dxf*dyf;
            if (power_f > 0.0f) continue;
            const float alpha_f = fminf_(ALPHA_CAP, co[3] * expf(power_f));
            if (alpha_f < ALPHA_MIN) continue;
            const double dx = (double)s->xy[2*g] - (double)pxf, dy = (double)s->xy[2*g+1] - (double)pyf;
            const double cA = co[0], cB = co[1], cC = co[2], op = co[3];
            const double G = exp(-0.5 * (cA*dx*dx + cC*dy*dy) - cB*dx*dy);
            const double alpha = fmin((double)ALPHA_CAP, op * G);
            T = T / (1.0 - alpha);
            const double w = alpha * T;
            double* a = acc + (size_t)g * 10;
            double dL_dalpha = 0.0;
            for (int ch = 0; ch < 3; ch++) {
                const double c = s->rgb[3*g+ch];
                rec_c[ch] = last_alpha * last_c[ch] + (1.0 - last_alpha) * rec_c[ch];
                last_c[ch] = c;
                dL_dalpha += (c - rec_c[ch]) * g_c[ch];
                a[6+ch] += w * g_c[ch];
            }
            const double dep = s->depth[g];
            rec_d = last_alpha * last_d + (1.0 - last_alpha) * rec_d;
            last_d = dep;
            dL_dalpha += (dep - rec_d) * g_d;
            a[9] += w * g_d;
            rec_a = last_alpha + (1.0 - last_alpha) * rec_a;
            dL_dalpha += (1.0 - rec_a) * g_a;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-(double)T_final / (1.0 - alpha)) * bg_dot;
            /* straight-through the 0.99 cap (A.7): no clamp mask */
            const double dL_dG = op * dL_dalpha;
            const double gdx = G * dx, gdy = G * dy;
            a[0] += dL_dG * (-gdx * cA - gdy * cB);
            a[1] += dL_dG * (-gdy * cC - gdx * cB);
            a[2] += -0.5 * gdx * dx * dL_dG;
            a[3] += -gdx * dy * dL_dG;
            a[4] += -0.5 * gdy * dy * dL_dG;
            a[5] += G * dL_dalpha;
        }
    }
}

/* --- A.8 backward preprocess for one Gaussian, in double --- */
static void preprocess_bwd_one(const GsoState* s, int i, const double* a,
                               float* g_means3D, float* g_means2D, float* g_shs, float* g_colors,
                               float* g_opac, float* g_scales, float* g_rots, float* g_cov3D)
{
    if (s->radii[i] <= 0) return;
    const float* V = s->view; const float* P = s->proj;
    const double px = s->means[3*i], py = s->means[3*i+1], pz = s->means[3*i+2];
    double gm[3] = {0, 0, 0};
    /* colour */
    if (s->use_sh) {
        const float* sh = s->shs + (size_t)i * s->M * 3;
        float* gsh = g_shs + (size_t)i * s->M * 3;
        double dx = px - s->campos[0], dy = py - s->campos[1], dz = pz - s->campos[2];
        double len = sqrt(dx*dx + dy*dy + dz*dz);
        double x = dx/len, y = dy/len, z = dz/len;
        double gdir[3] = {0, 0, 0};
        for (int ch = 0; ch < 3; ch++) {
            double g = s->clamped[3*i+ch] ? 0.0 : a[6+ch];
            double b[16], bx[16], by[16], bz[16];
            for (int k = 0; k < 16; k++) { b[k] = bx[k] = by[k] = bz[k] = 0; }
            b[0] = SH_C0;
            if (s->deg > 0) {
                b[1] = -SH_C1*y; by[1] = -SH_C1;
                b[2] = SH_C1*z;  bz[2] = SH_C1;
                b[3] = -SH_C1*x; bx[3] = -SH_C1;
            }
            if (s->deg > 1) {
                double xx = x*x, yy = y*y, zz = z*z;
                b[4] = SH_C2[0]*x*y; bx[4] = SH_C2[0]*y; by[4] = SH_C2[0]*x;
                b[5] = SH_C2[1]*y*z; by[5] = SH_C2[1]*z; bz[5] = SH_C2[1]*y;
                b[6] = SH_C2[2]*(2*zz - xx - yy); bx[6] = SH_C2[2]*-2*x; by[6] = SH_C2[2]*-2*y; bz[6] = SH_C2[2]*4*z;
                b[7] = SH_C2[3]*x*z; bx[7] = SH_C2[3]*z; bz[7] = SH_C2[3]*x;
                b[8] = SH_C2[4]*(xx - yy); bx[8] = SH_C2[4]*2*x; by[8] = SH_C2[4]*-2*y;
            }
            if (s->deg > 2) {
                double xx = x*x, yy = y*y, zz = z*z;
                b[9] = SH_C3[0]*y*(3*xx - yy); bx[9] = SH_C3[0]*6*x*y; by[9] = SH_C3[0]*(3*xx - 3*yy);
                b[10] = SH_C3[1]*x*y*z; bx[10] = SH_C3[1]*y*z; by[10] = SH_C3[1]*x*z; bz[10] = SH_C3[1]*x*y;
                b[11] = SH_C3[2]*y*(4*zz - xx - yy); bx[11] = SH_C3[2]*-2*x*y; by[11] = SH_C3[2]*(4*zz - xx - 3*yy); bz[11] = SH_C3[2]*8*y*z;
                b[12] = SH_C3[3]*z*(2*zz - 3*xx - 3*yy); bx[12] = SH_C3[3]*-6*x*z; by[12] = SH_C3[3]*-6*y*z; bz[12] = SH_C3[3]*(6*zz - 3*xx - 3*yy);
                b[13] = SH_C3[4]*x*(4*zz - xx - yy); bx[13] = SH_C3[4]*(4*zz - 3*xx - yy); by[13] = SH_C3[4]*-2*x*y; bz[13] = SH_C3[4]*8*x*z;
                b[14] = SH_C3[5]*z*(xx - yy); bx[14] = SH_C3[5]*2*x*z; by[14] = SH_C3[5]*-2*y*z; bz[14] = SH_C3[5]*(xx - yy);
                b[15] = SH_C3[6]*x*(xx - 3*yy); bx[15] = SH_C3[6]*(3*xx - 3*yy); by[15] = SH_C3[6]*-6*x*y;
            }
            int K = (s->deg + 1) * (s->deg + 1);
            for (int k = 0; k < K; k++) {
                gsh[k*3+ch] += (float)(b[k] * g);
                gdir[0] += bx[k] * sh[k*3+ch] * g; gdir[1] += by[k] * sh[k*3+ch] * g; gdir[2] += bz[k] * sh[k*3+ch] * g;
            }
        }
        /* through normalisation: (I - d d^T)/len */
        double dot = gdir[0]*x + gdir[1]*y + gdir[2]*z;
        gm[0] += (gdir[0] - dot*x) / len; gm[1] += (gdir[1] - dot*y) / len; gm[2] += (gdir[2] - dot*z) / len;
    } else {
        for (int ch = 0; ch < 3; ch++) g_colors[3*i+ch] += (float)a[6+ch];
    }
    g_opac[i] += (float)a[5];

    /* recompute forward intermediates in double from the fp32 inputs */
    double tx = V[0]*px + V[4]*py + V[8]*pz + V[12];
    double ty = V[1]*px + V[5]*py + V[9]*pz + V[13];
    double tz = V[2]*px + V[6]*py + V[10]*pz + V[14];
    const double fx = s->W / (2.0 * s->tanfovx), fy = s->H / (2.0 * s->tanfovy);
    const double limx = 1.3 * s->tanfovx, limy = 1.3 * s->tanfovy;
    double txtz = tx/tz, tytz = ty/tz;
    double mx = (txtz < -limx || txtz > limx) ? 0.0 : 1.0, my = (tytz < -limy || tytz > limy) ? 0.0 : 1.0;
    double cx = fmin(limx, fmax(-limx, txtz)) * tz, cy = fmin(limy, fmax(-limy, tytz)) * tz;
    double J00 = fx/tz, J02 = -fx*cx/(tz*tz), J11 = fy/tz, J12 = -fy*cy/(tz*tz);
    double Wr[3][3]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Wr[r][c] = V[4*c + r];
    double T[2][3];
    for (int j = 0; j < 3; j++) { T[0][j] = J00*Wr[0][j] + J02*Wr[2][j]; T[1][j] = J11*Wr[1][j] + J12*Wr[2][j]; }
    const float* c3 = s->cov3d + 6*i;
    double S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    double TS[2][3];
    for (int r = 0; r < 2; r++) for (int j = 0; j < 3; j++) TS[r][j] = T[r][0]*S[0][j] + T[r][1]*S[1][j] + T[r][2]*S[2][j];
    double ca = TS[0][0]*T[0][0] + TS[0][1]*T[0][1] + TS[0][2]*T[0][2] + LOWPASS;
    double cb = TS[0][0]*T[1][0] + TS[0][1]*T[1][1] + TS[0][2]*T[1][2];
    double cc = TS[1][0]*T[1][0] + TS[1][1]*T[1][1] + TS[1][2]*T[1][2] + LOWPASS;
    double det = ca*cc - cb*cb;
    double d2 = 1.0 / (det*det + 0.0000001);          /* upstream epsilon kept */
    const double gA = a[2], gB = a[3], gC = a[4];
    double da = d2 * (-cc*cc*gA + cb*cc*gB - cb*cb*gC);
    double db = d2 * (2*cb*cc*gA - (det + 2*cb*cb)*gB + 2*ca*cb*gC);
    double dc = d2 * (-cb*cb*gA + ca*cb*gB - ca*ca*gC);
    double g2[2][2] = {{da, 0.5*db}, {0.5*db, dc}};
    /* G3 = T^T g2 T : full symmetric dL/dSigma */
    double G3[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double v = 0; for (int p = 0; p < 2; p++) for (int q = 0; q < 2; q++) v += T[p][r]*g2[p][q]*T[q][c];
        G3[r][c] = v;
    }
    /* dL/dT = 2 g2 T S */
    double dT[2][3];
    for (int r = 0; r < 2; r++) for (int j = 0; j < 3; j++) dT[r][j] = 2.0*(g2[r][0]*TS[0][j] + g2[r][1]*TS[1][j]);
    double dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
    for (int j = 0; j < 3; j++) { dJ00 += dT[0][j]*Wr[0][j]; dJ02 += dT[0][j]*Wr[2][j]; dJ11 += dT[1][j]*Wr[1][j]; dJ12 += dT[1][j]*Wr[2][j]; }
    double itz = 1.0/tz, itz2 = itz*itz, itz3 = itz2*itz;
    double dtx = mx * (-fx*itz2) * dJ02;
    double dty = my * (-fy*itz2) * dJ12;
    double dtz = -fx*itz2*dJ00 - fy*itz2*dJ11 + 2*fx*cx*itz3*dJ02 + 2*fy*cy*itz3*dJ12;
    dtz += a[9];                                       /* depth = view-space z (A.7) */
    for (int j = 0; j < 3; j++) gm[j] += Wr[0][j]*dtx + Wr[1][j]*dty + Wr[2][j]*dtz;

    /* mean2D -> mean3D through the perspective divide of the full projection */
    double hx = P[0]*px + P[4]*py + P[8]*pz + P[12];
    double hy = P[1]*px + P[5]*py + P[9]*pz + P[13];
    double hw = P[3]*px + P[7]*py + P[11]*pz + P[15];
    double pw = 1.0 / (hw + 0.0000001);
    double gx_ndc = a[0] * 0.5 * s->W, gy_ndc = a[1] * 0.5 * s->H;   /* NDC-scaled units (API contract) */
    for (int j = 0; j < 3; j++) {
        gm[j] += (P[4*j+0]*pw - P[4*j+3]*hx*pw*pw) * gx_ndc + (P[4*j+1]*pw - P[4*j+3]*hy*pw*pw) * gy_ndc;
    }
    g_means2D[3*i] += (float)gx_ndc; g_means2D[3*i+1] += (float)gy_ndc;
    for (int j = 0; j < 3; j++) g_means3D[3*i+j] += (float)gm[j];

    if (s->use_cov_precomp) {
        g_cov3D[6*i+0] += (float)G3[0][0]; g_cov3D[6*i+1] += (float)(2*G3[0][1]); g_cov3D[6*i+2] += (float)(2*G3[0][2]);
        g_cov3D[6*i+3] += (float)G3[1][1]; g_cov3D[6*i+4] += (float)(2*G3[1][2]); g_cov3D[6*i+5] += (float)G3[2][2];
        return;
    }
    /* Sigma = R D R^T, D = diag((mod*scale)^2) */
    double sc[3] = {s->mod * (double)s->scales[3*i], s->mod * (double)s->scales[3*i+1], s->mod * (double)s->scales[3*i+2]};
    double r = s->rots[4*i], x = s->rots[4*i+1], y = s->rots[4*i+2], z = s->rots[4*i+3];
    double R[3][3] = {
        {1 - 2*(y*y + z*z), 2*(x*y - r*z),     2*(x*z + r*y)},
        {2*(x*y + r*z),     1 - 2*(x*x + z*z), 2*(y*z - r*x)},
        {2*(x*z - r*y),     2*(y*z + r*x),     1 - 2*(x*x + y*y)}};
    double GR[3][3];   /* G3 * R */
    for (int p = 0; p < 3; p++) for (int k = 0; k < 3; k++) GR[p][k] = G3[p][0]*R[0][k] + G3[p][1]*R[1][k] + G3[p][2]*R[2][k];
    for (int k = 0; k < 3; k++) {
        double rgr = R[0][k]*GR[0][k] + R[1][k]*GR[1][k] + R[2][k]*GR[2][k];
        g_scales[3*i+k] += (float)(2.0 * sc[k] * s->mod * rgr);
    }
    double Hm[3][3];
    for (int p = 0; p < 3; p++) for (int k = 0; k < 3; k++) Hm[p][k] = 2.0 * GR[p][k] * sc[k]*sc[k];
    g_rots[4*i+0] += (float)(2*(-z*Hm[0][1] + y*Hm[0][2] + z*Hm[1][0] - x*Hm[1][2] - y*Hm[2][0] + x*Hm[2][1]));
    g_rots[4*i+1] += (float)(2*( y*Hm[0][1] + z*Hm[0][2] + y*Hm[1][0] - 2*x*Hm[1][1] - r*Hm[1][2] + z*Hm[2][0] + r*Hm[2][1] - 2*x*Hm[2][2]));
    g_rots[4*i+2] += (float)(2*(-2*y*Hm[0][0] + x*Hm[0][1] + r*Hm[0][2] + x*Hm[1][0] + z*Hm[1][2] - r*Hm[2][0] + z*Hm[2][1] - 2*y*Hm[2][2]));
    g_rots[4*i+3] += (float)(2*(-2*z*Hm[0][0] - r*Hm[0][1] + x*Hm[0][2] + r*Hm[1][0] - 2*z*Hm[1][1] + y*Hm[1][2] + x*Hm[2][0] + y*Hm[2][1]));
}

/*
 * Backward.  Incoming grads: dL/dcolor [3,H,W], dL/ddepth [H,W], dL/dalpha [H,W].
 * Outputs are ACCUMULATED INTO (caller zero-fills): means3D [N,3], means2D [N,3] (NDC-scaled, z=0),
 * shs [N,M,3] or colors [N,3], opacities [N], scales [N,3], rotations [N,4] (or cov3D [N,6]).
 * Optional raw 2D accumulators out (acc2d [N,10], see blend_tile_bwd) for stage-wise checks.
 * f32_replay != 0 selects the reference-faithful fp32 blend replay (noise-floor measurement).
 */
void gso_backward(const GsoState* s, const float* g_color, const float* g_depth, const float* g_alpha,
                  float* g_means3D, float* g_means2D, float* g_shs, float* g_colors, float* g_opac,
                  float* g_scales, float* g_rots, float* g_cov3D, double* acc2d_out, int f32_replay)
{
    const int N = s->N, tiles = s->gx * s->gy;
    if (N == 0) return;
    int nth = 1;
#ifdef _OPENMP
    nth = omp_get_max_threads();
#endif
    double* accs = (double*)calloc((size_t)nth * N * 10, sizeof(double));
    #pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double* acc = accs + (size_t)tid * N * 10;
        #pragma omp for schedule(dynamic, 4)
        for (int t = 0; t < tiles; t++) {
            if (f32_replay) blend_tile_bwd_f32(s, t, g_color, g_depth, g_alpha, acc);
            else blend_tile_bwd(s, t, g_color, g_depth, g_alpha, acc);
        }
    }
    for (int t = 1; t < nth; t++) {
        const double* src = accs + (size_t)t * N * 10;
        #pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < (int64_t)N * 10; k++) accs[k] += src[k];
    }
    if (acc2d_out) memcpy(acc2d_out, accs, sizeof(double) * (size_t)N * 10);
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++)
        preprocess_bwd_one(s, i, accs + (size_t)i * 10, g_means3D, g_means2D, g_shs, g_colors, g_opac, g_scales, g_rots, g_cov3D);
    free(accs);
}

/* markVisible (K10): view-space z > 0.2 */
void gso_mark_visible(int N, const float* means3D, const float* viewmatrix, uint8_t* visible)
{
    for (int i = 0; i < N; i++) {
        const float* p = means3D + 3*i;
        float tz = viewmatrix[2]*p[0] + viewmatrix[6]*p[1] + viewmatrix[10]*p[2] + viewmatrix[14];
        visible[i] = tz > NEAR_CULL;
    }
}

int gso_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
