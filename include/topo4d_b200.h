/*
 * topo4d_b200.h -- C ABI of libtopo4d_b200.so (sm_100a).
 *
 * The drop-in boundary for Topo4D's two native hot paths.  Plain pointers and sizes only:
 * no torch / numpy / C++ types.  The caller owns every byte of memory (device or host as
 * stated); the library never allocates device memory and keeps no global mutable state, so
 * one process per GPU (torchrun) and one thread per GPU are both safe.  All entry points
 * enqueue on the given CUDA stream and return 0, or a negative GS_E_* / F3D_E_* code
 * (never throw); `*_last_error()` returns a static description of a code.
 *
 * What each entry point replaces in the reference:
 *   gs_forward / gs_backward ... `_C.rasterize_gaussians[_backward]` of the un-vendored
 *       `diff_gaussian_rasterization` extension that `GaussianRasterizer.forward/backward`
 *       dispatch to; reference call sites train.py:307,388,463,484 (render),
 *       train.py:667,738 (loss.backward()), settings built at helpers.py:73-86,
 *       inputs built at helpers.py:91-112.
 *   gs_mark_visible ........... `_C.mark_visible` behind `GaussianRasterizer.markVisible`
 *       (API completeness; Topo4D never calls it).
 *   t4d_image_loss ............ the per-iteration image loss and its autograd backward:
 *       `exp(cam_m)*im + cam_c` (train.py:310), `0.8*l1_loss_v1 + 0.2*(1 - calc_ssim)` (train.py:317;
 *       helpers.py:115-116, external.py:71-116), and loss.backward() down to dL/d(rendered image).
 *   t4d_adam_step ............. `optimizer.step()` of torch.optim.Adam(param_groups, lr=0.0, eps=1e-15)
 *       (train.py:272-297, 672) fused with the boolean-mask overwrites that follow it (train.py:676-700).
 *   t4d_activate[_backward] ... the activations of `params2rendervar` (helpers.py:91-112) and their autograd backward.
 *   t4d_dense_attribute ....... `compute_vertex_attribute_by_weight_2` (helpers.py:237-253; train.py:498-508).
 *   f3d_render_colors[_host] .. `_render_colors_core` (face3d/mesh/cython/mesh_core.cpp:169-234)
 *       as bound by `render_colors_core` (face3d/mesh/cython/mesh_core_cython.pyx:64-77) and
 *       reached through face3d/mesh/render.py:52-86 from helpers.py:956.
 */
#ifndef TOPO4D_B200_H_
#define TOPO4D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gs_stream_t;          /* a cudaStream_t */

/* ---- error codes ---- */
#define GS_OK                 0
#define GS_E_BAD_ARGS        -1     /* NULL / inconsistent arguments (both or neither of shs|colors, ...) */
#define GS_E_WORKSPACE_SMALL -2     /* workspace_bytes < gs_workspace_bytes(...) */
#define GS_E_CUDA            -3     /* a CUDA runtime call failed; see gs_last_cuda_error() */
#define GS_E_OVERFLOW        -4     /* (only from gs_read_status) instances needed > cap_instances */
#define GS_E_UNSUPPORTED     -5

/* ---- per-view camera block: GS_CAM_FLOATS consecutive floats on the DEVICE per view ----
 * Mirrors GaussianRasterizationSettings (helpers.py:73-86).  Matrices are the 16 consecutive
 * floats the reference passes: element [4*col+row] of the mathematical matrix
 * (viewmatrix = w2c^T, projmatrix = (P w2c)^T in PyTorch row-major terms). */
#define GS_CAM_FLOATS   48
#define GS_CAM_VIEW      0          /* [16] viewmatrix   */
#define GS_CAM_PROJ     16          /* [16] projmatrix   */
#define GS_CAM_CAMPOS   32          /* [3]  campos       */
#define GS_CAM_BG       35          /* [3]  bg           */
#define GS_CAM_TANFOVX  38
#define GS_CAM_TANFOVY  39          /* 40..47 reserved, must be 0 */

#define GS_HINT_UNKNOWN      0
#define GS_HINT_SHORT_LISTS  1
#define GS_HINT_LONG_LISTS   2
#define GS_HINT_SORT_MASK    3

/* ---- problem description shared by forward and backward ---- */
typedef struct GsProblem {
    int32_t N;                      /* Gaussians                                              */
    int32_t V;                      /* camera views rendered by this call (the reference: 1)  */
    int32_t H, W;                   /* image size, shared by all V views                      */
    int32_t sh_degree;              /* active SH degree 0..3 (settings.sh_degree)             */
    int32_t sh_coeffs;              /* M: coefficients stored per Gaussian in `shs` [N,M,3]   */
    float   scale_modifier;
    int32_t debug;                  /* !=0: synchronise + check after every stage             */
    int32_t blend_px;               /* tuning hint: pixels per thread in the blend kernels: 4 (many non-empty
                                       tiles: fewest instructions), 2 or 1 (few tiles: more warps per tile,
                                       lower latency); 0 or any other value = library default (4).  Never
                                       changes results.                                                   */
    int32_t hints;                  /* tuning hints, never change results.  Bits 0-1, per-tile list lengths the caller expects (from
                                       GsStatus.max_tile_instances of an earlier call of the same shape): GS_HINT_UNKNOWN (0),
                                       GS_HINT_SHORT_LISTS (no list longer than 2048: skip the long-list sort kernel),
                                       GS_HINT_LONG_LISTS.  Other bits must be 0.                                      */
    int64_t cap_instances;          /* capacity (tile,Gaussian) instances of the workspace    */
    /* inputs, DEVICE pointers, fp32, contiguous.  Exactly one of shs|colors_precomp and
       exactly one of (scales,rotations)|cov3D_precomp must be non-NULL. */
    const float* means3D;           /* [N,3] */
    const float* shs;               /* [N,M,3] */
    const float* colors_precomp;    /* [N,3]   */
    const float* opacities;         /* [N] (the reference passes [N,1]) */
    const float* scales;            /* [N,3]   */
    const float* rotations;         /* [N,4] (w,x,y,z), used as given (not renormalised) */
    const float* cov3D_precomp;     /* [N,6]   */
    const float* cameras;           /* [V,GS_CAM_FLOATS] */
    /* state kept between forward and backward, DEVICE, caller-owned */
    void*   workspace;
    size_t  workspace_bytes;
} GsProblem;

typedef struct GsForwardOut {       /* DEVICE pointers */
    float*   color;                 /* [V,3,H,W] */
    float*   depth;                 /* [V,1,H,W] un-normalised sum(alpha_i T_i z_i), no bg */
    float*   alpha;                 /* [V,1,H,W] sum(alpha_i T_i) */
    int32_t* radii;                 /* [V,N] screen radius in px, 0 = culled */
} GsForwardOut;

typedef struct GsBackwardIO {       /* DEVICE pointers */
    const float* dL_dcolor;         /* [V,3,H,W] */
    const float* dL_ddepth;         /* [V,1,H,W] or NULL (= zeros) */
    const float* dL_dalpha;         /* [V,1,H,W] or NULL (= zeros) */
    const int32_t* radii;           /* [V,N] as returned by forward */
    /* outputs: OVERWRITTEN with the sum over the V views (never accumulated into). */
    float* dL_dmeans3D;             /* [N,3] */
    float* dL_dmeans2D;             /* [N,3] NDC-scaled screen-space gradient, z = 0 */
    float* dL_dshs;                 /* [N,M,3]  (when shs given)            */
    float* dL_dcolors;              /* [N,3]    (when colors_precomp given) */
    float* dL_dopacities;           /* [N]   */
    float* dL_dscales;              /* [N,3] (when scales/rotations given)  */
    float* dL_drotations;           /* [N,4] */
    float* dL_dcov3D;               /* [N,6] (when cov3D_precomp given)     */
} GsBackwardIO;

typedef struct GsStatus {           /* host copy of the device status block */
    int64_t num_instances;          /* `num_rendered`: instances the V views need            */
    int64_t cap_instances;
    int32_t overflow;               /* 1 if num_instances > cap_instances (outputs invalid)  */
    int32_t max_tile_instances;     /* longest per-tile list                                 */
    int32_t num_active_tiles;       /* non-empty (view, tile) pairs: what blend_px should be chosen from */
    int32_t reserved0;
} GsStatus;

/* Byte size of the workspace for (N, V, H, W, cap_instances).  Pure host arithmetic. */
size_t gs_workspace_bytes(int32_t N, int32_t V, int32_t H, int32_t W, int64_t cap_instances);

/* Forward for V views: preprocess -> tile-bin -> per-tile depth sort + record gather ->
 * front-to-back blend.  Asynchronous.  Fills the device status block inside the workspace. */
int gs_forward(const GsProblem* p, const GsForwardOut* out, gs_stream_t stream);

/* Backward for the same V views; needs the workspace exactly as forward left it. */
int gs_backward(const GsProblem* p, const GsBackwardIO* io, gs_stream_t stream);

/* The same pipelines one stage at a time (bit mask, stages run in pipeline order), so a caller can put
 * CUDA events between kernels on its own stream (bench.py's per-kernel roofline timing does).
 * gs_forward == gs_forward_stages(GS_FWD_ALL), gs_backward == gs_backward_stages(GS_BWD_ALL). */
#define GS_FWD_PREPROCESS 1u        /* status/histogram reset + preprocess + tile scan */
#define GS_FWD_SCATTER    2u        /* (depth,id) pairs into tile segments */
#define GS_FWD_SORT       4u        /* per-tile sort + record gather */
#define GS_FWD_BLEND      8u        /* front-to-back blend */
#define GS_FWD_ALL       15u
#define GS_BWD_BLEND      1u        /* grad2d reset + back-to-front blend backward */
#define GS_BWD_PREPROCESS 2u        /* per-Gaussian chain rule, sums the V views */
#define GS_BWD_ALL        3u
int gs_forward_stages(const GsProblem* p, const GsForwardOut* out, uint32_t stages, gs_stream_t stream);
int gs_backward_stages(const GsProblem* p, const GsBackwardIO* io, uint32_t stages, gs_stream_t stream);

/* Copies the status block to the host (synchronises `stream`).  Returns GS_E_OVERFLOW if the
 * last forward needed more instances than the workspace holds (grow cap_instances, retry). */
int gs_read_status(const GsProblem* p, GsStatus* status_host, gs_stream_t stream);

/* Runs only preprocess + tile count + scan and returns the exact instance count (synchronises).
 * Lets a caller size `cap_instances` before the first real forward. */
int gs_count_instances(const GsProblem* p, int64_t* num_instances_host, gs_stream_t stream);

/* visible[i] = view-space z of means3D[i] > 0.2 for camera block `camera` (one view). */
int gs_mark_visible(int32_t N, const float* means3D, const float* camera, uint8_t* visible, gs_stream_t stream);

/* Introspection for the bit-exact index tests: device pointers into the workspace. */
typedef struct GsWorkspaceView {
    const uint32_t* tile_start;     /* [V*tiles+1] exclusive scan: tile t of view v owns [start[v*tiles+t], start[..+1]) */
    const uint32_t* sorted_ids;     /* [cap] Gaussian index per instance, tile-major, (depth, index) ascending */
    const float*    sorted_records; /* [cap,12] 48-byte records (x,y,-conA/2,-conB | -conC/2,opacity,depth,thr | r,g,b,id + reach mask << 24) */
    const float*    geom_records;   /* [V*N,12] same layout, per (view, Gaussian) */
    const float*    final_T;        /* [V,H,W] */
    const uint32_t* n_contrib;      /* [V,H,W] */
    const float*    grad2d;         /* [V*N,12] dL/d(pix.x,pix.y,conA,conB) | (conC,opacity,depth,_) | (r,g,b,_) after backward */
    int32_t tiles_x, tiles_y;
} GsWorkspaceView;
int gs_workspace_view(const GsProblem* p, GsWorkspaceView* view);

const char* gs_last_error(int code);
const char* gs_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * face3d render_colors
 * Semantics of _render_colors_core (mesh_core.cpp:169-234): painter with strict `>` depth test
 * (larger z = nearer, first triangle wins ties), integer bbox clipped to the image, the 2-px
 * border rule, barycentric colour interpolation; `image` [h,w,c] and `depth_buffer` [h,w] are
 * updated IN PLACE (image keeps its previous content where nothing is drawn).
 * ---------------------------------------------------------------------------------------- */
#define F3D_OK            0
#define F3D_E_BAD_ARGS   -1
#define F3D_E_CUDA       -3
#define F3D_E_WORKSPACE  -2

/* Device-pointer path.  workspace: f3d_workspace_bytes(ntri,h,w) bytes of device scratch. */
size_t f3d_workspace_bytes(int32_t ntri, int32_t h, int32_t w);
int f3d_render_colors(float* image, const float* vertices, const int32_t* triangles, const float* colors,
                      float* depth_buffer, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c,
                      void* workspace, size_t workspace_bytes, gs_stream_t stream);
/* Scratch budget per key plane (calling thread; <= 0 restores the default of 1 GiB = the whole image up to 16384^2): larger
 * images are resolved in bands of rows.  Call before f3d_workspace_bytes.  Results never depend on it. */
void f3d_set_band_bytes(int64_t bytes);
/* The bake as face3d/mesh/render.py:52-86 performs it (fresh zero image, private depth buffer filled with `depth_init` =
 * -999999 and discarded): writes EVERY pixel of `image` [h,w,c] fp32 or -- fused epilogue of helpers.py:959 -- of
 * `image_u8` [h,w,c] = (uint8)(value*255); exactly one of the two is non-NULL; no depth plane is read or written. */
int f3d_bake_colors(float* image, uint8_t* image_u8, const float* vertices, const int32_t* triangles, const float* colors,
                    float depth_init, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c,
                    void* workspace, size_t workspace_bytes, gs_stream_t stream);
/* HOST-pointer path, the calling convention of `render_colors_core` (face3d/mesh/cython/mesh_core_cython.pyx:64-77):
 * C-contiguous host arrays, image [h,w,c] and depth_buffer [h,w] updated in place, complete on return.  Copies and
 * device scratch are handled inside (the only entry point that allocates device memory: stream-ordered, freed on return). */
int f3d_render_colors_host(float* image, const float* vertices, const int32_t* triangles, const float* colors,
                           float* depth_buffer, int32_t nver, int32_t ntri, int32_t h, int32_t w, int32_t c);
/* Optional fused epilogue of helpers.py:959: out_u8[h,w,c] = (uint8)(image*255) (C truncation). */
int f3d_image_to_u8(const float* image, uint8_t* out_u8, int64_t count, gs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimisation-loop tail: image loss forward+backward, fused Adam.
 * ---------------------------------------------------------------------------------------- */
typedef struct T4dImageLoss {
    int32_t V, H, W;                /* views in this call (the reference: 1), image size                       */
    float   w_l1, w_ssim;           /* train.py:317: 0.8 and 0.2 (times any outer loss weight)                 */
    const float* render;            /* [V,3,H,W] rasterizer output, BEFORE the per-camera affine               */
    const float* target;            /* [V,3,H,W] ground-truth frame (curr_data['im'])                          */
    const float* cam_m;             /* [V,3] or NULL: im = exp(cam_m) * render + cam_c (train.py:310); both or neither */
    const float* cam_c;             /* [V,3] or NULL                                                           */
    float*  loss;                   /* out [V,4]: mean|im-gt|, mean SSIM, w_l1*l1 + w_ssim*(1-ssim), 0          */
    float*  dL_drender;             /* out [V,3,H,W]: d(sum over views of loss[v][2]) / d render; NULL = forward only */
    float*  dL_dcam_m;              /* out [V,3] or NULL                                                       */
    float*  dL_dcam_c;              /* out [V,3] or NULL                                                       */
    void*   workspace;              /* t4d_image_loss_workspace_bytes(V,H,W) bytes, 256-byte aligned, device   */
    size_t  workspace_bytes;
} T4dImageLoss;
size_t t4d_image_loss_workspace_bytes(int32_t V, int32_t H, int32_t W);
int t4d_image_loss(const T4dImageLoss* p, gs_stream_t stream);

#define T4D_ADAM_MAX_SEGMENTS 24
typedef struct T4dAdamSegment {    /* one named parameter = one torch param group (train.py:290-294)            */
    float*  param;                  /* [count] updated in place                                                 */
    const float* grad;              /* [count]                                                                  */
    float*  exp_avg;                /* [count] Adam first moment (state['exp_avg'])                             */
    float*  exp_avg_sq;             /* [count] second moment                                                    */
    const uint8_t* pin_mask;        /* [count/row_width] or NULL: rows overwritten after the step (train.py:676-700) */
    const float* pin_values;        /* [count] values for pinned rows; NULL = zeros                             */
    int64_t count;
    int32_t row_width;              /* elements per row (3 for means3D / colours / scales, 4 rotations, 1 opacities) */
    int32_t step;                   /* 1-based step count of this parameter (state['step'] after increment)     */
    float   lr;                     /* the group's current learning rate (update_optimizer, helpers.py:801-804) */
    int32_t* step_device;           /* NULL, or a DEVICE counter of completed steps: the launch then derives the bias
                                       corrections from it and advances it (CUDA-graph replay safe); `step` is ignored */
    const float* lr_device;         /* with step_device only: NULL, or a DEVICE float holding the learning rate (so an
                                       update_optimizer() between replays needs no re-capture); `lr` is then ignored */
} T4dAdamSegment;
/* `segments` is a HOST array; betas/eps as in torch.optim.Adam (reference: 0.9, 0.999, 1e-15). */
int t4d_adam_step(const T4dAdamSegment* segments, int32_t nseg, float beta1, float beta2, float eps, gs_stream_t stream);

/* Fused parameter activations of params2rendervar (helpers.py:91-112): rotations = normalize(unnorm_rotations) [N,4]
 * (x / max(||x||, 1e-12), as torch.nn.functional.normalize), opacities = sigmoid(logit_opacities) [N,1],
 * scales = exp(log_scales) [N,3]; and their backward.  Gradient inputs / outputs of the backward may be NULL
 * (NULL input = zero gradient, NULL output = not wanted).  Quaternion arrays must be 16-byte aligned. */
int t4d_activate(const float* unnorm_rotations, const float* logit_opacities, const float* log_scales, int32_t N,
                 float* rotations, float* opacities, float* scales, gs_stream_t stream);
int t4d_activate_backward(const float* unnorm_rotations, const float* opacities, const float* scales,
                          const float* dL_drotations, const float* dL_dopacities, const float* dL_dscales, int32_t N,
                          float* dL_dunnorm_rotations, float* dL_dlogit_opacities, float* dL_dlog_scales, gs_stream_t stream);

/* Dense Gaussian-mesh attribute interpolation: compute_vertex_attribute_by_weight_2
 * (helpers.py:237-253, called per frame from update_dense_states, train.py:498-508).  dense_out[n_base + n_new, channels]:
 * rows < n_base copy `attribute`; row n_base + i = sum_j weight[i][j] * attribute[quad_faces[vertex_father[i]][j]],
 * evaluated in float64 in NumPy's order and rounded once to float32 (bit-identical to the reference + `.float()`).
 * All pointers are device pointers; quad_faces [F,4] int32, vertex_father [n_new] int32, weight [n_new,4] float64. */
int t4d_dense_attribute(const float* attribute, int32_t n_base, int32_t channels, const int32_t* quad_faces,
                        const int32_t* vertex_father, const double* weight, int32_t n_new, float* dense_out,
                        gs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TOPO4D_B200_H_ */
