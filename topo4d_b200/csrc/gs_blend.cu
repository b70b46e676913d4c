// gs_blend.cu -- forward (K6) and backward (K7) alpha-blend kernels, sm_100a.
//
// Replaces upstream renderCUDA forward/backward of the un-vendored rasterizer behind
// GaussianRasterizer (reference call sites train.py:307,388 forward; train.py:667,738 backward).
// Same per-pixel semantics (alpha = min(0.99, o*exp(power)), skip alpha < 1/255,
// stop at T(1-alpha) < 1e-4, straight-through cap in backward); the machine mapping is new:
//
//   * work = (view, 16x16 binning tile).  Persistent CTAs pull tiles from device-side queues (atomic counters in
//     the status block): the compacted list of NON-EMPTY tiles, and -- forward only -- groups of 16 tiles whose
//     empty members just receive the background.  One worker warp in eight starts on the fill queue, the others on
//     the heavy queue (each falls over to the other queue when its own runs dry), so the issue-bound blending and the
//     HBM-bound background stream overlap on every SM and nobody idles on a static tile->CTA map.
//   * a thread owns a vertical comb of PX pixels (x, y+4k): lane (lx,ly) of a warp sits at column lx of the
//     warp's 8 columns, rows ly+4k.  The column terms of the quadratic form are shared by the PX pixels (3 FP
//     ops per pixel for `power`), and pixel slot k of a warp is one COMPACT 8x4 block, so the divergent blend
//     code of a slot runs only when the splat reaches that block and with a dense lane mask.
//   * warp-level compaction: every record carries the mask of the 8x4 blocks it can reach (gs_binning.cu); lane l
//     tests record l of a 32-record group, a ballot yields the records that can touch THIS warp's blocks (about
//     half of them at PX = 4) and only those are walked.
//   * a conservative per-Gaussian threshold `thr` (stored in the record) rejects a pixel without touching
//     exp(); ~88 % of (pixel, Gaussian) pairs leave after 3 instructions.  Finished / out-of-image pixels are
//     parked at y = 1e18, which fails that same test, so the walk carries no per-pixel state checks.
//   * the tile's depth-sorted 48-byte records are contiguous in HBM (gs_binning.cu): a chunk of 64
//     records is ONE cp.async.bulk (SASS UBLKCP) into shared memory, double-buffered on two mbarriers.
//   * backward replays back-to-front from the tile's deepest contributor with T_i = T_{i+1}/(1-alpha_i).
//     alpha is re-derived by the SAME inlined code as in the forward (identical bits), so the division
//     undoes the forward's multiplication to within ulps and nothing is amplified by 1/(1-alpha); the
//     colour/depth/alpha "behind" terms collapse into one scalar suffix sum per pixel (see Q below).
//     Per (pixel, Gaussian) only the ten MOMENTS t, t dx, t dy, t dx^2, t dx dy, t dy^2, w g_* are accumulated
//     (t = G dL/dalpha); they are pre-added over a thread's pixels, reduced across the warp through a
//     shared-memory transpose, parked in a per-warp slot (no atomics), turned into the conic / position /
//     opacity gradients ONCE per (tile, Gaussian) at flush time and added to HBM with three 16-byte vector REDs.
#include "gs_common.cuh"

namespace {

// PX = pixels per thread (4, 2 or 1).  A tile is always 256 pixels, so a CTA has 256/PX threads = 8/PX warps;
// warp w sits at columns 8*(w&1).. and rows (w>>1)*4*PX.., lane (lx,ly) owns pixels (8*(w&1)+lx, base+ly+4k).
// PX = 4 minimises instructions per (pixel, Gaussian) pair and is used when there are enough non-empty tiles
// to fill the GPU; PX = 2 / 1 trade instructions for 2x / 4x more warps per tile when there are few tiles
// (small scenes, one view per GPU): the per-tile latency, not the throughput, bounds those launches.
#ifndef GS_CHUNK
#define GS_CHUNK 32
#endif
constexpr int CHUNK = GS_CHUNK;               // records per bulk copy (32 -> 1.5 KB) = one ballot of the compaction
static_assert(CHUNK == 32, "one 32-bit ballot per chunk");
// Warps are independent workers (own work item, own record ring, own mbarriers); a CTA is just a container.
#ifndef GS_WPC
#define GS_WPC 4
#endif
constexpr int WPC = GS_WPC;
// Minimum resident CTAs per SM asked of ptxas: after the warp-level compaction the record loop is a dependent
// ffs -> address -> LDS -> FMA chain, so the kernels want warps more than registers (measured at 64-thread CTAs:
// bwd 12 -> 72 regs, fwd 14 -> 70 regs, no spills, 5-8 % faster than the unconstrained build).
// One worker in GS_FILL_EVERY starts on the background-fill queue (the others start blending); a few warps are enough
// to keep the HBM write stream busy, the rest hide the blend path's latency.
#ifndef GS_FILL_EVERY
#define GS_FILL_EVERY 8
#endif
#ifndef GS_BWD_MINB
#define GS_BWD_MINB (24 / GS_WPC)
#endif
#ifndef GS_FWD_MINB
#define GS_FWD_MINB (28 / GS_WPC)
#endif
constexpr uint32_t REC_BYTES = 48;
[[maybe_unused]] constexpr float LOG2E = 1.4426950408889634f;


// ---- per-pair arithmetic shared by forward and backward (identical bits in both) ----
struct ColTerms { float hC, u, v; };          // power(dy) = dy*(hC*dy + u) + v for a fixed pixel column
__device__ __forceinline__ ColTerms col_terms(float cA, float cB, float cC, float dx)
{
    ColTerms r;
    r.hC = __fmul_rn(-0.5f, cC);
    r.u = -__fmul_rn(cB, dx);
    r.v = __fmul_rn(__fmul_rn(-0.5f, __fmul_rn(cA, dx)), dx);
    return r;
}
__device__ __forceinline__ float splat_power(const ColTerms& r, float dy)
{
    return __fmaf_rn(dy, __fmaf_rn(r.hC, dy, r.u), r.v);
}
// exp flavour (GS_EXP_MODE): 2 (default) = ex2.approx on a two-term product x*log2(e) with first-order correction,
// ~2 ulp, 6 instructions; 1 = expf (<= 1 ulp, ~13 instructions); 0 = bare ex2.approx(x*log2e) (~8 ulp at |x| = 5,
// 3 instructions -- breaks the 1e-3 gradient bar at opacity 1, kept for experiments only).
#ifndef GS_EXP_MODE
#define GS_EXP_MODE 2
#endif
__device__ __forceinline__ float splat_exp(float power)
{
#if GS_EXP_MODE == 1
    return expf(power);
#elif GS_EXP_MODE == 0
    float y = __fmul_rn(power, LOG2E), g;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(y));
    return g;
#else
    const float L2E_HI = 1.4426950216293335f, L2E_LO = 1.9259629911266175e-8f, LN2 = 0.6931471805599453f;
    const float y = __fmul_rn(power, L2E_HI);
    float r = __fmaf_rn(power, L2E_HI, -y);               // exact rounding error of the product
    r = __fmaf_rn(power, L2E_LO, r);
    float g;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(y));
    return __fmaf_rn(g, __fmul_rn(r, LN2), g);            // 2^(y+r) = 2^y (1 + r ln2 + O(r^2)), |r| < 1e-6
#endif
}
__device__ __forceinline__ float splat_alpha(float opacity, float G) { return fminf(GS_ALPHA_CAP, __fmul_rn(opacity, G)); }
__device__ __forceinline__ float next_T(float T, float alpha) { return __fmul_rn(T, __fsub_rn(1.0f, alpha)); }

// bits (within the record's 8-bit block-reach mask, bit b = 8x4 block row b>>1, column b&1) of the blocks warp `warp`
// of a PX-pixels-per-thread CTA owns
template <int PX>
__device__ __forceinline__ unsigned warp_blocks(int warp)
{
    constexpr unsigned rows = PX == 4 ? 0x55u : (PX == 2 ? 0x5u : 0x1u);       // PX consecutive block rows of one column
    return rows << (((warp >> 1) * PX) * 2 + (warp & 1));
}

// A pixel that takes no further part (finished, or outside the image) is parked here: its `power` against any
// record is about -1e36 * conic.C, below every thr (gs_preprocess.cu clamps thr at -1e20), so the record walk
// needs no per-pixel state test.
constexpr float PARKED_Y = 1.0e18f;

// 32-bit shared-memory addresses that the compiler must keep in a register instead of re-deriving them from
// threadIdx.x inside the hot loop (it does, under the register cap: +15 integer instructions per reduction)
__device__ __forceinline__ uint32_t pinned_smem_addr(const void* ptr)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(ptr);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float x)
{
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(x) : "memory");
}
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

// ---- work queues ----
// A work item is one warp's share of a non-empty tile: sub-block `sub` of the 8 / PX a tile splits into (PX = 4: the
// left / right 8 x 16 half).  Items are encoded tile * 8 + sub; a background-fill group as -(g + 2).
constexpr long long ITEM_DONE = -1;
template <int NWT>
__device__ __forceinline__ long long fetch_heavy(const GsParams& p, unsigned int* cursor)
{
    const unsigned q = atomicAdd(cursor, 1u);
    const long long t = gs_active_tile(p, q / NWT);
    return t >= 0 ? t * 8 + (long long)(q % NWT) : ITEM_DONE;
}
template <int NWT>
__device__ __forceinline__ long long fetch_fwd(const GsParams& p, bool prefer_fill, unsigned n_groups)
{
    GsStatusDev* st = p.status;
    #pragma unroll
    for (int attempt = 0; attempt < 2; attempt++) {
        const bool fill = (attempt == 0) == prefer_fill;
        if (fill) {
            const unsigned g = atomicAdd(&st->q_fwd_fill, 1u);
            if (g < n_groups) return -((long long)g + 2);
        } else {
            const long long it = fetch_heavy<NWT>(p, &st->q_fwd_heavy);
            if (it >= 0) return it;
        }
    }
    return ITEM_DONE;
}

struct TileCtx {
    int v, tx0, ty0;       // view, first pixel of the tile
    unsigned long long start;
    int n;
};
__device__ __forceinline__ TileCtx tile_ctx(const GsParams& p, long long tg)
{
    TileCtx c;
    c.v = (int)((unsigned)tg / (unsigned)p.tiles);            // V*tiles < 2^31 (validated on the host)
    const int t = (int)tg - c.v * p.tiles;
    c.tx0 = (t % p.tiles_x) * GS_TILE;
    c.ty0 = (t / p.tiles_x) * GS_TILE;
    unsigned long long s = p.tile_start[tg], e = p.tile_start[tg + 1];
    if (e > (unsigned long long)p.cap) e = (unsigned long long)p.cap;
    c.start = s;
    c.n = e > s ? (int)(e - s) : 0;
    return c;
}

// background fill uses a 4x1 strip per thread: one 16-byte store per plane when the row allows it
__device__ __forceinline__ void store4(float* __restrict__ plane, size_t pix0, float v, int valid, bool vec)
{
    if (vec && valid == 4) { *reinterpret_cast<float4*>(plane + pix0) = make_float4(v, v, v, v); return; }
    #pragma unroll
    for (int k = 0; k < 4; k++) if (k < valid) plane[pix0 + k] = v;
}

template <int PX>
__global__ void __launch_bounds__(WPC * 32, GS_FWD_MINB)
blend_fwd_kernel(const GsParams p, float* __restrict__ out_color, float* __restrict__ out_depth,
                 float* __restrict__ out_alpha)
{
    constexpr int NWT = 8 / PX;                                 // warps (work items) per tile
    __shared__ __align__(128) float4 s_rec[WPC][2][CHUNK * 3];
    __shared__ __align__(8) uint64_t s_bar[WPC][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 (*const ring)[CHUNK * 3] = s_rec[warp];
    uint64_t* const bar = s_bar[warp];
    constexpr unsigned ALL = (1u << PX) - 1u;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phases = 0u;                       // bit b = parity to wait for on bar[b]
    const size_t HW = (size_t)p.H * p.W;
    const bool vec = (p.W & 3) == 0;
    const unsigned n_groups = (unsigned)((p.total_tiles + GS_FILL_GROUP - 1) / GS_FILL_GROUP);
    const bool prefer_fill = ((blockIdx.x * WPC + warp) % GS_FILL_EVERY) == GS_FILL_EVERY - 1;

    for (;;) {
        long long item = 0;
        if (lane == 0) item = fetch_fwd<NWT>(p, prefer_fill, n_groups);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item == ITEM_DONE) break;

        if (item < 0) {
            // ---- background fill of the empty tiles of one group (work item = 4x1 pixel strip of one tile) ----
            const long long t0 = (-item - 2) * GS_FILL_GROUP;
            for (int it = lane; it < GS_FILL_GROUP * 64; it += 32) {
                const long long tg = t0 + (it >> 6);
                if (tg >= p.total_tiles) break;
                if (p.tile_start[tg + 1] != p.tile_start[tg]) continue;          // non-empty: heavy items own it
                const TileCtx tc = tile_ctx(p, tg);
                const int x0 = tc.tx0 + 4 * (it & 3), y = tc.ty0 + ((it & 63) >> 2);
                const int valid = y < p.H ? min(4, p.W - x0) : 0;
                if (valid <= 0) continue;
                const float* __restrict__ bg = p.cams + (size_t)tc.v * GS_CAM_FLOATS + GS_CAM_BG;
                const size_t vb = (size_t)tc.v * HW, pix0 = (size_t)y * p.W + x0;
                store4(out_color + vb * 3, pix0, bg[0], valid, vec);
                store4(out_color + vb * 3 + HW, pix0, bg[1], valid, vec);
                store4(out_color + vb * 3 + 2 * HW, pix0, bg[2], valid, vec);
                store4(out_depth + vb, pix0, 0.f, valid, vec);
                store4(out_alpha + vb, pix0, 0.f, valid, vec);
                store4(p.final_T + vb, pix0, 1.f, valid, vec);
                store4(reinterpret_cast<float*>(p.n_contrib) + vb, pix0, 0.f, valid, vec);   // bit pattern 0
            }
            __syncwarp();
            continue;
        }

        // ---- this warp's share of one non-empty tile ----
        const int sub = (int)(item & 7);
        const TileCtx tc = tile_ctx(p, item >> 3);
        const int cx = (sub & 1) * 8 + (lane & 7), cy = (sub >> 1) * (4 * PX) + (lane >> 3);
        const unsigned my_blocks = warp_blocks<PX>(sub);
        const float pxf = (float)(tc.tx0 + cx);
        float pyf[PX], T[PX], C0[PX], C1[PX], C2[PX], D[PX], A[PX];
        uint32_t last[PX];
        unsigned done = 0;                                      // bit k: pixel k finished (or outside)
        #pragma unroll
        for (int k = 0; k < PX; k++) {
            T[k] = 1.f; C0[k] = C1[k] = C2[k] = D[k] = A[k] = 0.f; last[k] = 0u;
            pyf[k] = (float)(tc.ty0 + cy + 4 * k);
            if (tc.tx0 + cx >= p.W || tc.ty0 + cy + 4 * k >= p.H) { done |= 1u << k; pyf[k] = PARKED_Y; }
        }
        const unsigned outside = done;
        const int nchunks = (tc.n + CHUNK - 1) / CHUNK;
        const float4* __restrict__ src = p.sorted_rec + tc.start * 3;
        if (lane == 0 && nchunks > 0) {
            const uint32_t bytes = (uint32_t)min(tc.n, CHUNK) * REC_BYTES;
            mbar_expect_tx(&bar[0], bytes);
            bulk_g2s(ring[0], src, bytes, &bar[0]);
        }
        for (int c = 0; c < nchunks; c++) {
            const int cur = c & 1;
            const bool have_next = c + 1 < nchunks;
            __syncwarp();                                       // every lane is done reading ring[cur ^ 1]
            if (have_next && lane == 0) {
                const uint32_t bytes = (uint32_t)min(tc.n - (c + 1) * CHUNK, CHUNK) * REC_BYTES;
                mbar_expect_tx(&bar[cur ^ 1], bytes);
                bulk_g2s(ring[cur ^ 1], src + (size_t)(c + 1) * CHUNK * 3, bytes, &bar[cur ^ 1]);
            }
            mbar_wait(&bar[cur], (phases >> cur) & 1u);
            phases ^= 1u << cur;
            const int cnt = min(tc.n - c * CHUNK, CHUNK);
            const float4* __restrict__ rec = ring[cur];
            const uint32_t* __restrict__ recw = reinterpret_cast<const uint32_t*>(ring[cur]);
            // Warp-level compaction: lane l looks at the reach mask of record l, a ballot turns the 32 answers into the
            // list of records that can touch this warp's pixels at all (~half of them), and only those are walked.
            // Structured per-record body (no break/continue out of divergent code) closed by __syncwarp(): the warp
            // re-converges every record.  Leaving the loop from inside the divergent blend block makes the compiler
            // re-converge only at loop exit, which serialises the 32 lanes (measured: 12x slower).
            unsigned live = __ballot_sync(0xffffffffu, lane < cnt && ((recw[lane * 12 + 11] >> 24) & my_blocks) != 0u);
            while (live) {                                                      // warp-uniform
                const int j = __ffs(live) - 1;
                live &= live - 1u;
                const float4 r0 = rec[j * 3], r1 = rec[j * 3 + 1];
                const ColTerms ct = col_terms(r0.z, r0.w, r1.x, __fsub_rn(r0.x, pxf));
                float pw[PX];
                bool hit[PX], any = false;
                #pragma unroll
                for (int k = 0; k < PX; k++) {
                    pw[k] = splat_power(ct, __fsub_rn(r0.y, pyf[k]));
                    hit[k] = pw[k] >= r1.w;                                      // the (rare) power > 0 skip is tested on the blend path
                    any = any || hit[k];
                }
                if (any) {
                    const float4 r2 = rec[j * 3 + 2];
                    #pragma unroll
                    for (int k = 0; k < PX; k++) {
                        if (hit[k]) {
                            const float alpha = splat_alpha(r1.y, splat_exp(pw[k]));
                            const float test_T = next_T(T[k], alpha);
                            const bool visible = alpha >= GS_ALPHA_MIN && pw[k] <= 0.0f;
                            const bool blend = visible && !(test_T < GS_T_MIN);
                            if (visible && !blend) { done |= 1u << k; pyf[k] = PARKED_Y; }
                            if (blend) {
                                const float w = __fmul_rn(alpha, T[k]);
                                C0[k] = __fmaf_rn(r2.x, w, C0[k]);
                                C1[k] = __fmaf_rn(r2.y, w, C1[k]);
                                C2[k] = __fmaf_rn(r2.z, w, C2[k]);
                                D[k] = __fmaf_rn(r1.z, w, D[k]);
                                A[k] = __fadd_rn(A[k], w);
                                T[k] = test_T;
                                last[k] = (uint32_t)(c * CHUNK + j + 1);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (__all_sync(0xffffffffu, done == ALL)) {
                if (have_next) { mbar_wait(&bar[cur ^ 1], (phases >> (cur ^ 1)) & 1u); phases ^= 1u << (cur ^ 1); }   // drain prefetch
                break;
            }
        }
        {
            const float* __restrict__ bg = p.cams + (size_t)tc.v * GS_CAM_FLOATS + GS_CAM_BG;
            const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
            const size_t vb = (size_t)tc.v * HW;
            #pragma unroll
            for (int k = 0; k < PX; k++) {
                if (outside & (1u << k)) continue;
                const size_t pix = (size_t)(tc.ty0 + cy + 4 * k) * p.W + (tc.tx0 + cx);
                p.final_T[vb + pix] = T[k];
                p.n_contrib[vb + pix] = last[k];
                out_color[vb * 3 + pix] = __fmaf_rn(T[k], b0, C0[k]);
                out_color[vb * 3 + HW + pix] = __fmaf_rn(T[k], b1, C1[k]);
                out_color[vb * 3 + 2 * HW + pix] = __fmaf_rn(T[k], b2, C2[k]);
                out_depth[vb + pix] = D[k];
                out_alpha[vb + pix] = A[k];
            }
        }
        __syncwarp();
    }
    if (lane == 0) gs_queue_release(&p.status->q_fwd_heavy, &p.status->done_fwd, gridDim.x * WPC, &p.status->q_fwd_fill);
}

template <int PX>
__global__ void __launch_bounds__(WPC * 32, GS_BWD_MINB)
blend_bwd_kernel(const GsParams p, const GsBackwardIO io)
{
    constexpr int NWT = 8 / PX;
    __shared__ __align__(128) float4 s_rec[WPC][2][CHUNK * 3];
    __shared__ __align__(16) float s_acc[WPC][CHUNK * GS_REC_FLOATS];   // reduced moments of the chunk's records
    __shared__ __align__(16) float s_tr[WPC][32 * GS_REC_FLOATS];       // transpose scratch of the reduction
    __shared__ __align__(8) uint64_t s_bar[WPC][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 (*const ring)[CHUNK * 3] = s_rec[warp];
    uint64_t* const bar = s_bar[warp];
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phases = 0u;
    const size_t HW = (size_t)p.H * p.W;
    // reduction addresses (see below): where this lane parks its partial record, the column it sums, its result slot
    const int red_c = lane & 15, red_h = lane >> 4;
    const bool red_on = red_c < GS_REC_FLOATS;
    const uint32_t a_park = pinned_smem_addr(&s_tr[warp][lane * GS_REC_FLOATS]);
    const uint32_t a_col = pinned_smem_addr(&s_tr[warp][red_h * GS_REC_FLOATS + (red_on ? red_c : 0)]);
    const uint32_t a_out = pinned_smem_addr(&s_acc[warp][red_on ? red_c : 0]);
    const float4* __restrict__ my_acc = reinterpret_cast<const float4*>(s_acc[warp]);

    for (;;) {
        long long item = 0;
        if (lane == 0) item = fetch_heavy<NWT>(p, &p.status->q_bwd_heavy);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item == ITEM_DONE) break;

        const int sub = (int)(item & 7);
        const TileCtx tc = tile_ctx(p, item >> 3);
        const int cx = (sub & 1) * 8 + (lane & 7), cy = (sub >> 1) * (4 * PX) + (lane >> 3);
        const unsigned my_blocks = warp_blocks<PX>(sub);
        const size_t vb = (size_t)tc.v * HW;
        const float* __restrict__ bg = p.cams + (size_t)tc.v * GS_CAM_FLOATS + GS_CAM_BG;
        const float b0 = bg[0], b1 = bg[1], b2 = bg[2];

        // per-pixel state: T (recovered back-to-front), loss gradients g, and
        // Q[k] = g . (colour, depth, alpha accumulated BEHIND the current splat) + T_final (bg . g): because the loss
        // gradients g are per-pixel constants the five suffix sums collapse into this one scalar.
        const float pxf = (float)(tc.tx0 + cx);
        float pyf[PX], T[PX], Q[PX], g0[PX], g1[PX], g2[PX], gd[PX], ga[PX];
        uint32_t last[PX];
        #pragma unroll
        for (int k = 0; k < PX; k++) {
            const int px = tc.tx0 + cx, py = tc.ty0 + cy + 4 * k;
            pyf[k] = (float)py;
            T[k] = 0.f; last[k] = 0u; g0[k] = g1[k] = g2[k] = gd[k] = ga[k] = 0.f;
            if (px < p.W && py < p.H) {
                const size_t pix = (size_t)py * p.W + px;
                T[k] = p.final_T[vb + pix];
                last[k] = p.n_contrib[vb + pix];
                g0[k] = io.dL_dcolor[vb * 3 + pix]; g1[k] = io.dL_dcolor[vb * 3 + HW + pix]; g2[k] = io.dL_dcolor[vb * 3 + 2 * HW + pix];
                if (io.dL_ddepth) gd[k] = io.dL_ddepth[vb + pix];
                if (io.dL_dalpha) ga[k] = io.dL_dalpha[vb + pix];
            }
            Q[k] = T[k] * (b0 * g0[k] + b1 * g1[k] + b2 * g2[k]);                // T_final * (bg . dL/dC)
        }
        uint32_t m = 0u;
        #pragma unroll
        for (int k = 0; k < PX; k++) m = max(m, last[k]);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        const int nmax = (int)min(m, (uint32_t)tc.n);           // deepest contributor of this warp's pixels
        if (nmax == 0) continue;                                // warp-uniform

        const int nchunks = (nmax + CHUNK - 1) / CHUNK;
        const float4* __restrict__ src = p.sorted_rec + tc.start * 3;
        float4* __restrict__ gbase = p.grad2d + (size_t)tc.v * p.N * 3;
        if (lane == 0) {
            const int c = nchunks - 1;
            const uint32_t bytes = (uint32_t)(nmax - c * CHUNK) * REC_BYTES;
            mbar_expect_tx(&bar[0], bytes);
            bulk_g2s(ring[0], src + (size_t)c * CHUNK * 3, bytes, &bar[0]);
        }
        for (int kc = 0; kc < nchunks; kc++) {
            const int c = nchunks - 1 - kc, cur = kc & 1;
            __syncwarp();                                       // every lane is done with ring[cur ^ 1] (walk and flush)
            if (c > 0 && lane == 0) {
                const uint32_t bytes = (uint32_t)CHUNK * REC_BYTES;              // every earlier chunk is full
                mbar_expect_tx(&bar[cur ^ 1], bytes);
                bulk_g2s(ring[cur ^ 1], src + (size_t)(c - 1) * CHUNK * 3, bytes, &bar[cur ^ 1]);
            }
            mbar_wait(&bar[cur], (phases >> cur) & 1u);
            phases ^= 1u << cur;
            const int cnt = min(nmax - c * CHUNK, CHUNK);
            const float4* __restrict__ rec = ring[cur];
            const uint32_t* __restrict__ recw = reinterpret_cast<const uint32_t*>(ring[cur]);
            // warp-level compaction (see the forward): only records whose reach mask meets this warp's 8x4 blocks are
            // walked, back to front
            unsigned live = __ballot_sync(0xffffffffu, lane < cnt && ((recw[lane * 12 + 11] >> 24) & my_blocks) != 0u);
            unsigned touched = 0u;                              // warp-uniform: records whose moments sit in s_acc
            while (live) {                                      // warp-uniform
                const int j = 31 - __clz(live);
                live ^= 1u << j;
                const uint32_t idx = (uint32_t)(c * CHUNK + j);
                const float4 r0 = rec[j * 3], r1 = rec[j * 3 + 1];
                const float dx = __fsub_rn(r0.x, pxf);
                const ColTerms ct = col_terms(r0.z, r0.w, r1.x, dx);
                float dyr[PX], pw[PX];
                bool hit[PX], any = false;
                #pragma unroll
                for (int k = 0; k < PX; k++) {
                    dyr[k] = __fsub_rn(r0.y, pyf[k]);
                    pw[k] = splat_power(ct, dyr[k]);
                    hit[k] = pw[k] >= r1.w && idx < last[k];           // the (rare) power > 0 skip is tested on the blend path
                    any = any || hit[k];
                }
                if (!__any_sync(0xffffffffu, any)) continue;
                // moments of t = G dL/dalpha over this thread's pixels: {t dx, t dy, t dx^2, t dx dy | t dy^2, t, w g_d, - | w g_rgb}
                float r[10];
                #pragma unroll
                for (int s = 0; s < 10; s++) r[s] = 0.f;
                if (any) {
                    const float4 r2 = rec[j * 3 + 2];
                    #pragma unroll
                    for (int k = 0; k < PX; k++) {
                        if (hit[k]) {
                            const float G = splat_exp(pw[k]);
                            const float alpha = splat_alpha(r1.y, G);
                            if (alpha >= GS_ALPHA_MIN && pw[k] <= 0.0f) {
                                // the recovery T_i = T_{i+1} * (1/(1-alpha)) is replayed once per contributing layer (thousands
                                // for very deep lists): a bare approximate reciprocal is biased and drifts past 1e-3, a
                                // Newton-refined one does not (and __frcp_rn costs 15 % of the kernel)
                                const float om = __fsub_rn(1.0f, alpha);             // in [0.01, 1]
                                float ra;
                                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(om));
                                ra = __fmaf_rn(__fmaf_rn(-om, ra, 1.0f), ra, ra);    // one Newton step: <= 1 ulp, unbiased
                                const float Tk = __fmul_rn(T[k], ra);            // undoes the forward's T*(1-alpha)
                                const float w = __fmul_rn(alpha, Tk);
                                // dL/dalpha_i = T_i (g . x_i) - (g . suffix_i + T_final bg.g) / (1 - alpha_i),  x_i = (rgb, depth, 1)
                                const float Pk = __fmaf_rn(g0[k], r2.x, __fmaf_rn(g1[k], r2.y, __fmaf_rn(g2[k], r2.z, __fmaf_rn(gd[k], r1.z, ga[k]))));
                                const float dLda = Tk * Pk - ra * Q[k];
                                Q[k] = __fmaf_rn(w, Pk, Q[k]);
                                // dL/dG * G = opacity * t (straight-through the 0.99 cap); the opacity, the conic entries and
                                // the -1/2 factors are constants of the record and are applied once, after the reduction
                                const float t = G * dLda;
                                const float tx = t * dx, ty = t * dyr[k];
                                r[0] += tx;
                                r[1] += ty;
                                r[2] = __fmaf_rn(tx, dx, r[2]);
                                r[3] = __fmaf_rn(tx, dyr[k], r[3]);
                                r[4] = __fmaf_rn(ty, dyr[k], r[4]);
                                r[5] += t;
                                r[6] = __fmaf_rn(w, gd[k], r[6]);
                                r[7] = __fmaf_rn(w, g0[k], r[7]); r[8] = __fmaf_rn(w, g1[k], r[8]); r[9] = __fmaf_rn(w, g2[k], r[9]);
                                T[k] = Tk;
                            }
                        }
                    }
                }
                // warp reduction through shared memory: every lane parks its 12-float partial record (three 16-byte
                // stores, conflict-free), then lane (c, h) = (lane & 15, lane >> 4) sums column c over rows 2i+h
                // (16 conflict-free loads) and one shuffle joins the halves: ~40 instructions instead of a 16-shuffle /
                // 32-select butterfly (~70).
                {
                    sts_v4(a_park, r[0], r[1], r[2], r[3]);
                    sts_v4(a_park + 16, r[4], r[5], r[6], 0.f);
                    sts_v4(a_park + 32, r[7], r[8], r[9], 0.f);
                    __syncwarp();
                    float acc = 0.f;
                    if (red_on) {
                        constexpr int RS = 2 * GS_REC_FLOATS * 4;            // byte stride of two rows
                        acc = ((lds_f32<0 * RS>(a_col) + lds_f32<1 * RS>(a_col)) + (lds_f32<2 * RS>(a_col) + lds_f32<3 * RS>(a_col))) +
                              ((lds_f32<4 * RS>(a_col) + lds_f32<5 * RS>(a_col)) + (lds_f32<6 * RS>(a_col) + lds_f32<7 * RS>(a_col)));
                        acc += ((lds_f32<8 * RS>(a_col) + lds_f32<9 * RS>(a_col)) + (lds_f32<10 * RS>(a_col) + lds_f32<11 * RS>(a_col))) +
                               ((lds_f32<12 * RS>(a_col) + lds_f32<13 * RS>(a_col)) + (lds_f32<14 * RS>(a_col) + lds_f32<15 * RS>(a_col)));
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (red_h == 0 && red_on) sts_f32(a_out + (uint32_t)j * (GS_REC_FLOATS * 4), acc);
                    __syncwarp();
                }
                touched |= 1u << j;
            }
            // flush the chunk: moments -> gradients of the 2-D record (pix.x, pix.y, conic A, B, C, opacity, depth, rgb),
            // conic B being the true (not halved) derivative; one 16-byte vector RED per touched (record, part)
            __syncwarp();
            for (int t = lane; t < cnt * 3; t += 32) {
                const int j = t / 3, part = t - j * 3;
                if (!((touched >> j) & 1u)) continue;
                float4 a = my_acc[t];
                const float4 q0 = rec[j * 3], q1 = rec[j * 3 + 1];
                const float o = q1.y;
                if (part == 0) a = make_float4(-o * (q0.z * a.x + q0.w * a.y), -o * (q1.x * a.y + q0.w * a.x), -0.5f * o * a.z, -o * a.w);
                else if (part == 1) a.x = -0.5f * o * a.x;
                const int id = __float_as_int(rec[j * 3 + 2].w) & 0x00ffffff;
                red_add_v4(gbase + (size_t)id * 3 + part, a);
            }
        }
        __syncwarp();
    }
    if (lane == 0) gs_queue_release(&p.status->q_bwd_heavy, &p.status->done_bwd, gridDim.x * WPC);
}

int resident_ctas(const void* kernel, int block, int num_sms, int fallback_per_sm)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = fallback_per_sm;
    return per_sm * num_sms;
}

}  // namespace

// pixels per thread for this launch: the caller's hint, else 4 (throughput-optimal for many tiles)
static int pick_px(const GsParams& p) { return (p.blend_px == 1 || p.blend_px == 2) ? p.blend_px : 4; }

template <int PX>
static void launch_fwd(const GsParams& p, float* color, float* depth, float* alpha, int num_sms, cudaStream_t s)
{
    static thread_local int grid = 0, grid_sms = 0;
    if (grid == 0 || grid_sms != num_sms) { grid = resident_ctas((const void*)blend_fwd_kernel<PX>, WPC * 32, num_sms, 4); grid_sms = num_sms; }
    blend_fwd_kernel<PX><<<grid, WPC * 32, 0, s>>>(p, color, depth, alpha);
}
template <int PX>
static void launch_bwd(const GsParams& p, const GsBackwardIO& io, int num_sms, cudaStream_t s)
{
    static thread_local int grid = 0, grid_sms = 0;
    if (grid == 0 || grid_sms != num_sms) { grid = resident_ctas((const void*)blend_bwd_kernel<PX>, WPC * 32, num_sms, 4); grid_sms = num_sms; }
    blend_bwd_kernel<PX><<<grid, WPC * 32, 0, s>>>(p, io);
}

void gs_launch_blend_fwd(const GsParams& p, float* color, float* depth, float* alpha, int num_sms, cudaStream_t s)
{
    switch (pick_px(p)) {
        case 1: launch_fwd<1>(p, color, depth, alpha, num_sms, s); break;
        case 2: launch_fwd<2>(p, color, depth, alpha, num_sms, s); break;
        default: launch_fwd<4>(p, color, depth, alpha, num_sms, s); break;
    }
}

void gs_launch_blend_bwd(const GsParams& p, const GsBackwardIO& io, int num_sms, cudaStream_t s)
{
    switch (pick_px(p)) {
        case 1: launch_bwd<1>(p, io, num_sms, s); break;
        case 2: launch_bwd<2>(p, io, num_sms, s); break;
        default: launch_bwd<4>(p, io, num_sms, s); break;
    }
}
