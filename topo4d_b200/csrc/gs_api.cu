// gs_api.cu -- C-ABI entry points of the Gaussian rasterizer (declared in include/topo4d_b200.h).
// Host-side sequencing only: validates arguments, carves the caller's workspace, enqueues kernels.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gs_common.cuh"

static thread_local char g_cuda_err[256] = "";

static int record_cuda(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return 0;
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return GS_E_CUDA;
}
#define CK(call) do { int _r = record_cuda((call), #call); if (_r) return _r; } while (0)
#define CK_LAUNCH(name) do { int _r = record_cuda(cudaGetLastError(), name); if (_r) return _r; \
    if (p->debug) { _r = record_cuda(cudaStreamSynchronize(s), name " (debug sync)"); if (_r) return _r; } } while (0)

int gs_pdl_enabled()
{
    static int cached = -1;                      // benign race: every thread computes the same value
    if (cached < 0) {
        const char* e = getenv("TOPO4D_B200_PDL");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached;
}

static int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount; cached_dev = dev;
    }
    return cached;
}

static int validate(const GsProblem* p)
{
    if (!p) return GS_E_BAD_ARGS;
    if (p->N < 0 || p->V < 1 || p->H < 1 || p->W < 1) return GS_E_BAD_ARGS;
    if (p->N > 0x00ffffff) return GS_E_UNSUPPORTED;      // 24-bit Gaussian index inside the tile-sorted records
    if (p->cap_instances < 0 || p->cap_instances > 0x7fffffffLL) return GS_E_BAD_ARGS;
    if ((long long)p->V * p->N > 0x7fffffffLL) return GS_E_BAD_ARGS;
    if ((long long)p->V * ((p->W + GS_TILE - 1) / GS_TILE) * ((p->H + GS_TILE - 1) / GS_TILE) > 0x7fffffffLL) return GS_E_BAD_ARGS;
    if (!p->workspace || !p->cameras) return GS_E_BAD_ARGS;
    if (p->N > 0) {
        if (!p->means3D || !p->opacities) return GS_E_BAD_ARGS;
        if ((p->shs != NULL) == (p->colors_precomp != NULL)) return GS_E_BAD_ARGS;
        const bool sr = p->scales != NULL && p->rotations != NULL;
        if (sr == (p->cov3D_precomp != NULL)) return GS_E_BAD_ARGS;
        if (!sr && (p->scales != NULL || p->rotations != NULL)) return GS_E_BAD_ARGS;
        if (p->shs) {
            if (p->sh_degree < 0 || p->sh_degree > 3) return GS_E_UNSUPPORTED;
            if (p->sh_coeffs < (p->sh_degree + 1) * (p->sh_degree + 1)) return GS_E_BAD_ARGS;
        }
    }
    // the kernels read rotations (and write their gradient) as 16-byte vectors
    if (p->N > 0 && p->rotations && ((uintptr_t)p->rotations & 15u) != 0) return GS_E_BAD_ARGS;
    if (p->workspace_bytes < gs_workspace_bytes(p->N, p->V, p->H, p->W, p->cap_instances)) return GS_E_WORKSPACE_SMALL;
    if (((uintptr_t)p->workspace & 255u) != 0) return GS_E_BAD_ARGS;
    return 0;
}

static GsParams make_params(const GsProblem* p, const GsLayout& L)
{
    GsParams q;
    memset(&q, 0, sizeof(q));
    char* ws = (char*)p->workspace;
    q.N = p->N; q.V = p->V; q.H = p->H; q.W = p->W; q.deg = p->sh_degree; q.M = p->sh_coeffs;
    q.tiles_x = L.tiles_x; q.tiles_y = L.tiles_y; q.tiles = L.tiles; q.total_tiles = L.total_tiles;
    q.cap = p->cap_instances; q.mod = p->scale_modifier;
    q.means3D = p->means3D; q.shs = p->shs; q.colors = p->colors_precomp; q.opac = p->opacities;
    q.scales = p->scales; q.rots = p->rotations; q.cov3D = p->cov3D_precomp; q.cams = p->cameras;
    q.status = (GsStatusDev*)(ws + L.off_status);
    q.tile_count = (uint32_t*)(ws + L.off_tile_count);
    q.tile_start = (uint32_t*)(ws + L.off_tile_start);
    q.tile_fill = (uint32_t*)(ws + L.off_tile_fill);
    q.active_tiles = (uint32_t*)(ws + L.off_active);
    q.block_sums = (uint32_t*)(ws + L.off_block_sums);
    q.clamped = (uint8_t*)(ws + L.off_clamped);
    q.geom = (float4*)(ws + L.off_geom);
    q.pairs = (unsigned long long*)(ws + L.off_pairs);
    q.sorted_ids = (uint32_t*)(ws + L.off_sorted_ids);
    q.sorted_rec = (float4*)(ws + L.off_sorted_rec);
    q.final_T = (float*)(ws + L.off_final_T);
    q.n_contrib = (uint32_t*)(ws + L.off_n_contrib);
    q.grad2d = (float4*)(ws + L.off_grad2d);
    q.scan_blocks = L.scan_blocks;
    q.blend_px = p->blend_px;
    q.sort_long = (p->hints & GS_HINT_SORT_MASK) != GS_HINT_SHORT_LISTS;       // unknown -> run the long-list kernel too
    return q;
}

extern "C" size_t gs_workspace_bytes(int32_t N, int32_t V, int32_t H, int32_t W, int64_t cap)
{
    if (N < 0 || V < 1 || H < 1 || W < 1 || cap < 0) return 0;
    return gs_make_layout(N, V, H, W, cap).total;
}

// preprocess + tile histogram + scan: everything up to knowing the instance count
static int run_front(const GsProblem* p, const GsParams& q, const GsLayout& L, int32_t* radii, cudaStream_t s)
{
    // ONE memset: the status block (work-queue cursors included) and the tile histogram are adjacent in the workspace
    static_assert(sizeof(GsStatusDev) <= 256, "status block must fit its 256-byte slot");
    if (L.off_status != 0 || L.off_tile_count != 256) return GS_E_BAD_ARGS;
    CK(cudaMemsetAsync(q.status, 0, L.off_tile_count + 4 * (size_t)(L.total_tiles + 1), s));
    gs_launch_preprocess(q, radii, s);
    CK_LAUNCH("preprocess_kernel");
    gs_launch_tile_scan(q, sm_count(), s);
    CK_LAUNCH("tile_scan");
    return 0;
}

extern "C" int gs_forward(const GsProblem* p, const GsForwardOut* out, gs_stream_t stream)
{
    return gs_forward_stages(p, out, GS_FWD_ALL, stream);
}

extern "C" int gs_forward_stages(const GsProblem* p, const GsForwardOut* out, uint32_t stages, gs_stream_t stream)
{
    int r = validate(p);
    if (r) return r;
    if (!out || !out->color || !out->depth || !out->alpha || (p->N > 0 && !out->radii)) return GS_E_BAD_ARGS;
    cudaStream_t s = (cudaStream_t)stream;
    const GsLayout L = gs_make_layout(p->N, p->V, p->H, p->W, p->cap_instances);
    const GsParams q = make_params(p, L);
    const size_t P = (size_t)p->V * p->H * p->W;
    if (p->N == 0) {
        // empty scene: zero images (not bg), T = 1 (SURVEY.md 8b "N = 0 -> zero image")
        CK(cudaMemsetAsync(out->color, 0, P * 3 * 4, s));
        CK(cudaMemsetAsync(out->depth, 0, P * 4, s));
        CK(cudaMemsetAsync(out->alpha, 0, P * 4, s));
        CK(cudaMemsetAsync(q.n_contrib, 0, P * 4, s));
        CK(cudaMemsetAsync(q.status, 0, sizeof(GsStatusDev), s));
        CK(cudaMemsetAsync(q.tile_start, 0, 4 * (size_t)(L.total_tiles + 1), s));
        return 0;
    }
    if (stages & GS_FWD_PREPROCESS) {
        r = run_front(p, q, L, out->radii, s);
        if (r) return r;
    }
    if (stages & GS_FWD_SCATTER) {
        gs_launch_scatter(q, out->radii, s);
        CK_LAUNCH("scatter_kernel");
    }
    const int sms = sm_count();
    if (stages & GS_FWD_SORT) {
        gs_launch_sort_gather(q, sms, s);
        CK_LAUNCH("sort_gather_kernel");
    }
    if (stages & GS_FWD_BLEND) {
        gs_launch_blend_fwd(q, out->color, out->depth, out->alpha, sms, s);
        CK_LAUNCH("blend_fwd_kernel");
    }
    return 0;
}

extern "C" int gs_backward(const GsProblem* p, const GsBackwardIO* io, gs_stream_t stream)
{
    return gs_backward_stages(p, io, GS_BWD_ALL, stream);
}

extern "C" int gs_backward_stages(const GsProblem* p, const GsBackwardIO* io, uint32_t stages, gs_stream_t stream)
{
    int r = validate(p);
    if (r) return r;
    if (!io || !io->dL_dcolor) return GS_E_BAD_ARGS;
    if (p->N == 0) return 0;
    if (!io->radii || !io->dL_dmeans3D || !io->dL_dmeans2D || !io->dL_dopacities) return GS_E_BAD_ARGS;
    if (p->shs ? !io->dL_dshs : !io->dL_dcolors) return GS_E_BAD_ARGS;
    if (p->cov3D_precomp ? !io->dL_dcov3D : (!io->dL_dscales || !io->dL_drotations)) return GS_E_BAD_ARGS;
    if (io->dL_drotations && ((uintptr_t)io->dL_drotations & 15u) != 0) return GS_E_BAD_ARGS;       // written as float4
    cudaStream_t s = (cudaStream_t)stream;
    const GsLayout L = gs_make_layout(p->N, p->V, p->H, p->W, p->cap_instances);
    const GsParams q = make_params(p, L);
    if (stages & GS_BWD_BLEND) {
        CK(cudaMemsetAsync(q.grad2d, 0, 48 * (size_t)p->V * p->N, s));
        gs_launch_blend_bwd(q, *io, sm_count(), s);
        CK_LAUNCH("blend_bwd_kernel");
    }
    if (stages & GS_BWD_PREPROCESS) {
        gs_launch_preprocess_bwd(q, *io, s);
        CK_LAUNCH("preprocess_bwd_kernel");
    }
    return 0;
}

extern "C" int gs_read_status(const GsProblem* p, GsStatus* st, gs_stream_t stream)
{
    if (!p || !st || !p->workspace) return GS_E_BAD_ARGS;
    cudaStream_t s = (cudaStream_t)stream;
    GsStatusDev d;
    CK(cudaMemcpyAsync(&d, p->workspace, sizeof(d), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    st->num_instances = (int64_t)d.num_instances;
    st->cap_instances = (int64_t)d.cap_instances;
    st->overflow = d.overflow;
    st->max_tile_instances = d.max_tile_instances;
    st->num_active_tiles = (int32_t)(d.num_long + d.num_short);
    st->reserved0 = 0;
    return d.overflow ? GS_E_OVERFLOW : 0;
}

extern "C" int gs_count_instances(const GsProblem* p, int64_t* n_host, gs_stream_t stream)
{
    int r = validate(p);
    if (r) return r;
    if (!n_host) return GS_E_BAD_ARGS;
    *n_host = 0;
    if (p->N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const GsLayout L = gs_make_layout(p->N, p->V, p->H, p->W, p->cap_instances);
    const GsParams q = make_params(p, L);
    // radii is an output of the real forward; this counting pass borrows the grad2d area (48*V*N bytes) for it
    int32_t* radii = (int32_t*)q.grad2d;
    r = run_front(p, q, L, radii, s);
    if (r) return r;
    GsStatus st;
    r = gs_read_status(p, &st, stream);
    if (r && r != GS_E_OVERFLOW) return r;
    *n_host = st.num_instances;
    return 0;
}

extern "C" int gs_mark_visible(int32_t N, const float* means3D, const float* camera, uint8_t* visible, gs_stream_t stream)
{
    if (N < 0 || (N > 0 && (!means3D || !camera || !visible))) return GS_E_BAD_ARGS;
    cudaStream_t s = (cudaStream_t)stream;
    gs_launch_mark_visible(N, means3D, camera, visible, s);
    return record_cuda(cudaGetLastError(), "mark_visible_kernel");
}

extern "C" int gs_workspace_view(const GsProblem* p, GsWorkspaceView* v)
{
    if (!p || !v || !p->workspace) return GS_E_BAD_ARGS;
    const GsLayout L = gs_make_layout(p->N, p->V, p->H, p->W, p->cap_instances);
    const GsParams q = make_params(p, L);
    v->tile_start = q.tile_start; v->sorted_ids = q.sorted_ids; v->sorted_records = (const float*)q.sorted_rec;
    v->geom_records = (const float*)q.geom; v->final_T = q.final_T; v->n_contrib = q.n_contrib;
    v->grad2d = (const float*)q.grad2d; v->tiles_x = L.tiles_x; v->tiles_y = L.tiles_y;
    return 0;
}

extern "C" const char* gs_last_error(int code)
{
    switch (code) {
        case GS_OK: return "ok";
        case GS_E_BAD_ARGS: return "bad arguments (NULL/inconsistent pointers, sizes, rotations / dL_drotations not 16-byte aligned, or not exactly one of shs|colors_precomp / (scales,rotations)|cov3D_precomp)";
        case GS_E_WORKSPACE_SMALL: return "workspace smaller than gs_workspace_bytes()";
        case GS_E_CUDA: return "CUDA runtime error (see gs_last_cuda_error)";
        case GS_E_OVERFLOW: return "instance capacity exceeded: grow cap_instances to status.num_instances and retry";
        case GS_E_UNSUPPORTED: return "unsupported configuration (sh_degree must be 0..3)";
        default: return "unknown error code";
    }
}

extern "C" const char* gs_last_cuda_error(void) { return g_cuda_err; }
