// gs_common.cuh -- shared declarations of the sm_100a Gaussian rasterizer kernels.
//
// Data layout in HBM (all caller-owned, carved out of GsProblem.workspace by ws_layout()):
//   status        GsStatusDev                       device status block
//   tile_count    u32 [V*T + 1]                     instances per (view, tile)
//   tile_start    u32 [V*T + 1]                     exclusive scan of tile_count; [V*T] = I
//   tile_fill     u32 [V*T]                         scatter cursors
//   active_tiles  u32 [V*T]                         non-empty (view,tile) ids: long lists from the front, short ones
//                                                   from the back (work queues hand out long tiles first)
//   block_sums    u32 [scan blocks]                 scan scratch
//   clamped       u8  [V*N]                         SH clamp bits (r,g,b)
//   geom          48-byte record [V*N]              (x,y,conA,conB | conC,opacity,depth,thr | r,g,b,id)
//                                                   thr = conservative lower bound of `power` for alpha >= 1/255
//   pairs         u64 [cap]                         (depth_bits<<32 | id), tile-major, unsorted
//   sorted_ids    u32 [cap]
//   sorted_rec    48-byte record [cap]              tile-sorted copy of geom: ONE contiguous bulk copy per chunk; conic pre-scaled
//                                                   (x,y,-conA/2,-conB | -conC/2,opacity,depth,thr | r,g,b,id + reach mask << 24)
//   final_T       f32 [V*H*W]
//   n_contrib     u32 [V*H*W]
//   grad2d        48-byte record [V*N]              (dpix.x,dpix.y,dconA,dconB | dconC,dopacity,ddepth,_ | dr,dg,db,_)
// T = tiles per view = ceil(W/16)*ceil(H/16).  The binning tile is 16x16 (reference contract).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/topo4d_b200.h"

#define GS_TILE 16
#ifndef GS_PRESCALE
#define GS_PRESCALE 1
#endif
#define GS_REC_FLOATS 12            // 48-byte records
#define GS_NEAR_CULL 0.2f
#define GS_LOWPASS 0.3f
#define GS_ALPHA_CAP 0.99f
#define GS_ALPHA_MIN (1.0f / 255.0f)
#define GS_T_MIN 0.0001f

struct GsStatusDev {
    unsigned long long num_instances;
    unsigned long long cap_instances;
    int overflow;
    int max_tile_instances;
    // work queues of the persistent blend kernels (device-side dynamic scheduling)
    unsigned int num_long;              // non-empty tiles with >= GS_LONG_TILE instances: active_tiles[0 .. num_long)
    unsigned int q_fwd_heavy;           // next index into active_tiles[] (forward)
    unsigned int q_fwd_fill;            // next group of GS_FILL_GROUP tiles to background-fill (forward)
    unsigned int q_bwd_heavy;           // next index into active_tiles[] (backward)
    unsigned int q_sort;                // next index into active_tiles[] (sort + gather)
    unsigned int num_short;             // the other non-empty tiles: active_tiles[total_tiles-1 .. ] downwards
    // (the first 48 bytes above are mirrored by the host layer; everything below is device-only)
    unsigned int scan_done;             // blocks of the fused tile scan that have published their total
    unsigned int done_sort, done_fwd, done_bwd;   // workers that have left a queue: the last one rewinds its cursor(s)
    unsigned int q_sort_long, done_sort_long;     // queue of the long-list sort kernel (walks active_tiles[0 .. num_long))
};
#ifndef GS_LONG_TILE
#define GS_LONG_TILE 256                // longest-first work order: long lists are handed out before short ones (measured at config 2,
                                        // blend_bwd: 128 -> 0.766 ms, 192 -> 0.756, 256 -> 0.751, 384 -> 0.759, 512 -> 0.767)
#endif
#define GS_FILL_GROUP 16

struct GsLayout {
    size_t off_status, off_tile_count, off_tile_start, off_tile_fill, off_active, off_block_sums, off_clamped, off_geom,
           off_pairs, off_sorted_ids, off_sorted_rec, off_final_T, off_n_contrib, off_grad2d, total;
    int tiles_x, tiles_y, tiles;        // per view
    long long total_tiles;              // V * tiles
    int scan_blocks;
    int blend_px;                       // pixels per thread of the blend kernels (1, 2, 4; see gs_blend.cu)
};

#define GS_SCAN_ELEMS_PER_BLOCK 4096    // 1024 threads x 4

static inline size_t gs_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline GsLayout gs_make_layout(int N, int V, int H, int W, long long cap)
{
    GsLayout L;
    L.tiles_x = (W + GS_TILE - 1) / GS_TILE;
    L.tiles_y = (H + GS_TILE - 1) / GS_TILE;
    L.tiles = L.tiles_x * L.tiles_y;
    L.total_tiles = (long long)V * L.tiles;
    L.scan_blocks = (int)((L.total_tiles + 1 + GS_SCAN_ELEMS_PER_BLOCK - 1) / GS_SCAN_ELEMS_PER_BLOCK);
    const size_t A = 256;
    size_t o = 0;
    const size_t VN = (size_t)V * (size_t)(N > 0 ? N : 1);
    const size_t P = (size_t)V * H * W;
    const size_t C = (size_t)(cap > 0 ? cap : 1);
    L.off_status = o;      o = gs_align_up(o + sizeof(GsStatusDev), A);
    L.off_tile_count = o;  o = gs_align_up(o + 4 * (size_t)(L.total_tiles + 1), A);
    L.off_tile_start = o;  o = gs_align_up(o + 4 * (size_t)(L.total_tiles + 1), A);
    L.off_tile_fill = o;   o = gs_align_up(o + 4 * (size_t)(L.total_tiles + 1), A);
    L.off_active = o;      o = gs_align_up(o + 4 * (size_t)(L.total_tiles + 1), A);
    L.off_block_sums = o;  o = gs_align_up(o + 4 * (size_t)(L.scan_blocks + 1), A);
    L.off_clamped = o;     o = gs_align_up(o + VN, A);
    L.off_geom = o;        o = gs_align_up(o + 48 * VN, A);
    L.off_pairs = o;       o = gs_align_up(o + 8 * C, A);
    L.off_sorted_ids = o;  o = gs_align_up(o + 4 * C, A);
    L.off_sorted_rec = o;  o = gs_align_up(o + 48 * C, A);
    L.off_final_T = o;     o = gs_align_up(o + 4 * P, A);
    L.off_n_contrib = o;   o = gs_align_up(o + 4 * P, A);
    L.off_grad2d = o;      o = gs_align_up(o + 48 * VN, A);
    L.total = o;
    return L;
}

// Everything a kernel needs, passed by value (__grid_constant__-sized, < 400 bytes).
struct GsParams {
    int N, V, H, W, deg, M;
    int tiles_x, tiles_y, tiles;
    long long total_tiles;
    long long cap;
    float mod;
    const float *means3D, *shs, *colors, *opac, *scales, *rots, *cov3D, *cams;
    GsStatusDev* status;
    uint32_t *tile_count, *tile_start, *tile_fill, *active_tiles, *block_sums;
    uint8_t* clamped;
    float4* geom;
    unsigned long long* pairs;
    uint32_t* sorted_ids;
    float4* sorted_rec;
    float* final_T;
    uint32_t* n_contrib;
    float4* grad2d;
    int scan_blocks;
    int blend_px;                       // pixels per thread of the blend kernels (1, 2, 4; see gs_blend.cu)
    int sort_long;                      // 1: sort_gather_long_kernel takes the lists of 2049 .. 16384 keys (sort_gather_kernel skips them)
};

// launchers implemented in the kernel translation units
void gs_launch_preprocess(const GsParams& p, int32_t* radii, cudaStream_t s);
void gs_launch_scatter(const GsParams& p, const int32_t* radii, cudaStream_t s);
void gs_launch_mark_visible(int N, const float* means3D, const float* cam, uint8_t* visible, cudaStream_t s);
void gs_launch_tile_scan(const GsParams& p, int num_sms, cudaStream_t s);
void gs_launch_sort_gather(const GsParams& p, int num_sms, cudaStream_t s);
void gs_launch_blend_fwd(const GsParams& p, float* color, float* depth, float* alpha, int num_sms, cudaStream_t s);
void gs_launch_blend_bwd(const GsParams& p, const GsBackwardIO& io, int num_sms, cudaStream_t s);
void gs_launch_preprocess_bwd(const GsParams& p, const GsBackwardIO& io, cudaStream_t s);

#ifdef __CUDACC__
// ---- programmatic dependent launch (sm_90+) ----
// The seven kernels of a step depend on each other in stream order, and between two dependent launches the GPU idles for the 2-4 us it
// takes to drain one grid and bring up the next: nothing for a 24-view step, 10 % of the reference's one-view step.  A kernel launched
// through gs_launch_dependent may become resident while its predecessor still runs -- as soon as every CTA of the predecessor has
// executed gs_pdl_trigger() -- and parks at gs_pdl_wait(), its first statement, until the predecessor grid has completed and its memory
// operations are visible (which transitively covers everything earlier in the stream).  TOPO4D_B200_PDL=0 restores plain launches.
int gs_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t gs_launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = gs_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void gs_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gs_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// i-th work item of the non-empty-tile queue (long tiles first), or -1 past the end
__device__ __forceinline__ long long gs_active_tile(const GsParams& p, unsigned i)
{
    const unsigned nl = p.status->num_long, ns = p.status->num_short;
    if (i < nl) return (long long)p.active_tiles[i];
    const unsigned j = i - nl;
    if (j < ns) return (long long)p.active_tiles[p.total_tiles - 1 - j];
    return -1;
}

// Self-cleaning work queues: a worker calls this once, after its last fetch came back empty.  The last of the
// `workers` to arrive rewinds the cursor (and the arrival counter), so the next launch on the same workspace --
// a repeated backward, a staged re-run -- starts from zero without a memset in the stream.
__device__ __forceinline__ void gs_queue_release(unsigned int* cursor, unsigned int* done, unsigned int workers, unsigned int* cursor2 = nullptr)
{
    __threadfence();
    if (atomicAdd(done, 1u) == workers - 1u) {
        *cursor = 0u;
        if (cursor2) *cursor2 = 0u;
        *done = 0u;
    }
}

// ---- mbarrier + bulk async copy (TMA-family, SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared::cta bulk copy, completion counted in bytes on `bar`.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 16-byte vector reduction to global memory (sm_90+): one RED per 4 floats.
__device__ __forceinline__ void red_add_v4(float4* addr, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
#endif
