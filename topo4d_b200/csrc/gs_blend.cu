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
// to keep the HBM write stream busy, the rest hide the blend path's latency.  Re-measured with the r02q kernels (blend_fwd, 24 views):
// 3 -> 0.543 ms, 4 -> 0.500, 5 -> 0.476, 6 -> 0.470, 8 -> 0.480, 12 -> 0.557.
#ifndef GS_FILL_EVERY
#define GS_FILL_EVERY 6
#endif
#ifndef GS_BWD_MINB
#define GS_BWD_MINB (24 / GS_WPC)
#endif
#ifndef GS_FWD_MINB
#define GS_FWD_MINB (28 / GS_WPC)
#endif
// PX = 2 backward: 1 = warp reduction on the tensor pipe (blend_bwd_mma_kernel), 0 = shared-memory transpose (blend_bwd_kernel<2>)
#ifndef GS_BWD_MMA
#define GS_BWD_MMA 1
#endif
constexpr uint32_t REC_BYTES = 48;
[[maybe_unused]] constexpr float LOG2E = 1.4426950408889634f;


// ---- per-pair arithmetic shared by forward and backward (identical bits in both) ----
struct ColTerms { float hC, u, v; };          // power(dy) = dy*(hC*dy + u) + v for a fixed pixel column
// The tile-sorted records carry the conic PRE-SCALED (hA = -A/2, nB = -B, hC = -C/2; gs_binning.cu writes them that way): scaling by a
// power of two and negation commute with rounding, so these three products have the bits of -0.5*(A*dx)*dx, -(B*dx), -0.5*C.
__device__ __forceinline__ ColTerms col_terms(float hA, float nB, float hC, float dx)
{
    ColTerms r;
#if GS_PRESCALE
    r.hC = hC;
    r.u = __fmul_rn(nB, dx);
    r.v = __fmul_rn(__fmul_rn(hA, dx), dx);
#else
    r.hC = __fmul_rn(-0.5f, hC);
    r.u = -__fmul_rn(nB, dx);
    r.v = __fmul_rn(__fmul_rn(-0.5f, __fmul_rn(hA, dx)), dx);
#endif
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
// two fp32 lanes in one 64-bit register: shared-memory load / store and the packed add (SASS FADD2, sm_100)
template <int OFF>
__device__ __forceinline__ unsigned long long lds_b64(uint32_t addr)
{
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ void sts_b64(uint32_t addr, unsigned long long v)
{
    asm volatile("st.shared.b64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long shfl_xor_b64(unsigned long long v, int m)
{
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}
// index of the most significant set bit (x != 0): one FLO instead of the clz / 31-x pair __clz compiles to
__device__ __forceinline__ int msb_index(unsigned x)
{
    int j;
    asm("bfind.u32 %0, %1;" : "=r"(j) : "r"(x));
    return j;
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
    gs_pdl_wait();                                              // launched as a dependent of sort_gather_kernel
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
    gs_pdl_trigger();                                           // preprocess_bwd_kernel may become resident (it waits for this grid)
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phases = 0u;
    const size_t HW = (size_t)p.H * p.W;
    // reduction addresses (see below): where this lane parks its partial record, the column PAIR it sums over the rows
    // q, q + 4, ..., its result slot
    const int red_cp = lane & 7, red_q = lane >> 3;           // (conflict-free for 8-byte loads: each half-warp reads 24 consecutive words)
    const bool red_on = red_cp < GS_REC_FLOATS / 2;
    const uint32_t a_park = pinned_smem_addr(&s_tr[warp][lane * GS_REC_FLOATS]);
    const uint32_t a_col = pinned_smem_addr(&s_tr[warp][red_q * GS_REC_FLOATS + (red_on ? 2 * red_cp : 0)]);
    const uint32_t a_out = pinned_smem_addr(&s_acc[warp][red_on ? 2 * red_cp : 0]);
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
            unsigned touched = live;                            // warp-uniform: records whose moments will sit in s_acc
            int lastc[PX];                                      // pixel k contributes to records j < lastc[k] of this chunk
            #pragma unroll
            for (int k = 0; k < PX; k++) lastc[k] = (int)last[k] - c * CHUNK;
            while (live) {                                      // warp-uniform
                const int j = msb_index(live);
                const unsigned jbit = 1u << j;
                live ^= jbit;
                const float4 r0 = rec[j * 3], r1 = rec[j * 3 + 1];
                const float dx = __fsub_rn(r0.x, pxf);
                const ColTerms ct = col_terms(r0.z, r0.w, r1.x, dx);
                float dyr[PX], pw[PX];
                bool hit[PX], any = false;
                #pragma unroll
                for (int k = 0; k < PX; k++) {
                    dyr[k] = __fsub_rn(r0.y, pyf[k]);
                    pw[k] = splat_power(ct, dyr[k]);
                    hit[k] = pw[k] >= r1.w && j < lastc[k];            // the (rare) power > 0 skip is tested on the blend path
                    any = any || hit[k];
                }
                if (!__any_sync(0xffffffffu, any)) { touched ^= jbit; continue; }
                // moments of t = G dL/dalpha over this thread's pixels: {t dx, t dy, t dx^2, t dx dy | t dy^2, t, w g_d, - | w g_rgb}
                float r[10];
                #pragma unroll
                for (int s = 0; s < 10; s++) r[s] = 0.f;
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
                // warp reduction through shared memory: every lane parks its 12-float partial record (three 16-byte
                // stores, conflict-free), then lane (cp, q) = (lane & 7, lane >> 3) sums the column pair cp over the rows
                // q + 4i with eight 8-byte loads and packed adds (FADD2), and two packed shuffle steps join the four row
                // classes: ~26 instructions per record instead of ~41 for scalar column sums (or ~70 for a shuffle
                // butterfly over ten values).  Lanes with cp > 5 repeat pair 0 and do not store.
                {
                    sts_v4(a_park, r[0], r[1], r[2], r[3]);
                    sts_v4(a_park + 16, r[4], r[5], r[6], 0.f);
                    sts_v4(a_park + 32, r[7], r[8], r[9], 0.f);
                    __syncwarp();
                    constexpr int RS = 4 * GS_REC_FLOATS * 4;                // byte stride of four rows
                    unsigned long long acc =
                        add_f32x2(add_f32x2(add_f32x2(lds_b64<0 * RS>(a_col), lds_b64<1 * RS>(a_col)), add_f32x2(lds_b64<2 * RS>(a_col), lds_b64<3 * RS>(a_col))),
                                  add_f32x2(add_f32x2(lds_b64<4 * RS>(a_col), lds_b64<5 * RS>(a_col)), add_f32x2(lds_b64<6 * RS>(a_col), lds_b64<7 * RS>(a_col))));
                    acc = add_f32x2(acc, shfl_xor_b64(acc, 8));
                    acc = add_f32x2(acc, shfl_xor_b64(acc, 16));
                    if (red_q == 0 && red_on) sts_b64(a_out + (uint32_t)j * (GS_REC_FLOATS * 4), acc);
                    __syncwarp();
                }
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
                // q0.z, q0.w, q1.x = -A/2, -B, -C/2 (pre-scaled conic)
#if GS_PRESCALE
                if (part == 0) a = make_float4(o * __fmaf_rn(2.0f * q0.z, a.x, q0.w * a.y), o * __fmaf_rn(2.0f * q1.x, a.y, q0.w * a.x), -0.5f * o * a.z, -o * a.w);
#else
                if (part == 0) a = make_float4(-o * (q0.z * a.x + q0.w * a.y), -o * (q1.x * a.y + q0.w * a.x), -0.5f * o * a.z, -o * a.w);
#endif
                else if (part == 1) a.x = -0.5f * o * a.x;
                const int id = __float_as_int(rec[j * 3 + 2].w) & 0x00ffffff;
                red_add_v4(gbase + (size_t)id * 3 + part, a);
            }
        }
        __syncwarp();
    }
    if (lane == 0) gs_queue_release(&p.status->q_bwd_heavy, &p.status->done_bwd, gridDim.x * WPC);
}

// ---- PX = 2 backward with the warp reduction on the (otherwise idle) tensor pipe --------------------------------------
// The shared-memory transpose reduction of blend_bwd_kernel moves 32 lanes x 48 B in and out of shared memory per
// (record, warp) visit: ncu showed the LSU data pipe at 84 % of its peak (176 M shared-memory wavefronts per 24-view launch),
// i.e. that kernel is bound by shared-memory bandwidth, not by instruction issue.  Here the lane -> pixel map is
//     lane = 4 g + t:   pixel 0 = (x, y) = (g, t),  pixel 1 = (g, t + 4)        inside the warp's 8 x 8 pixel block
// (a vertical comb like blend_bwd_kernel's, so the two pixels still share the column terms of the quadratic form and pixel
// slot k is still one compact 8 x 4 block), which is exactly the fragment geometry of mma.sync.m16n8k8 (TF32): the four lanes
// of a quad and the two k-halves are the contraction index (= y, 0..7), g is a free index (= x).  The six geometric moments of
// a record are polynomial in the pixel coordinates: with tt = G dL/dalpha per pixel and coordinates centred on the block
// (x', y' = -3.5 .. 3.5, which keeps the later re-centring on the splat well conditioned),
//     S_n(x) = sum_y tt(x, y) y'^n  (n = 0, 1, 2)     is ONE mma with B[k][.] = (1, k', k'^2),
// and M1 = sum_x S_0, My = sum_x S_1, Mx = sum_x x' S_0, Mxy = sum_x x' S_1, Myy = sum_x S_2, Mxx = sum_x x'^2 S_0 follow by a per-lane
// scale (1, x', 1, x'^2 by lane-in-quad; B repeats the S_0 / S_1 columns so every lane receives the pair it scales).  The four
// colour / depth moments sum_pixels w g_c take two more mma (high and low halves) with a 0/1 selector as B.  TF32 keeps 10
// mantissa bits, so every fp32 operand is split exactly into hi (upper bits) + lo (remainder): the monomials and selectors are
// small integers (exact), products are exact, accumulation is fp32 -- measured agreement with an fp32 shuffle reduction: 5e-6
// worst case, 1e-7 typical (tools/probes/mma_quadsum.cu).  What reaches shared memory is one 12-float row per QUAD (8 rows, 3
// wavefronts) instead of one per lane (32 rows, 12 wavefronts), and the column sums read 8 rows instead of 32.
// The moments are taken about the block centre (half-integer coordinates and their squares: exact in TF32); the flush
// re-centres them on the splat (ex = mean.x - Xc, ey = mean.y - Yc) once per (chunk, record).  About the block CORNER the
// second moments lost 5x accuracy to cancellation (conic gradients 6.5e-4 instead of 1.3e-4 of the oracle, tools/probe_grad2d.py).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// exact split of an fp32 into a TF32-representable high part (bit pattern) and the remainder
__device__ __forceinline__ void split_tf32(float v, unsigned& hi, unsigned& lo)
{
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(__fsub_rn(v, __uint_as_float(hi)));
}
__device__ __forceinline__ void sts_v2(uint32_t addr, float x, float y)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(addr), "f"(x), "f"(y) : "memory");
}
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

constexpr int QROW = 8;                       // floats per parked quad row (per part)
constexpr int QW_OFF = 8 * QROW + 16;         // second part (colour / depth moments) starts 16 banks away from the first

// Minimum-CTAs hint 6 (<= 80 registers allowed): ptxas still allocates 72 registers, so 7 CTAs x 4 warps stay resident per SM, but it
// schedules the record loop differently from the hint-7 build -- measured with the r02r kernel 0.745 ms vs 0.752 (hint 7) vs 0.770
// (hint 5).  (r02e, older kernel: 0.789 ms at hint 7 vs 0.807 at 6 / 0.800 at 8 with 64 registers + stack.)
#ifndef GS_BWDM_MINB
#define GS_BWDM_MINB (24 / GS_WPC)
#endif
__global__ void __launch_bounds__(WPC * 32, GS_BWDM_MINB)
blend_bwd_mma_kernel(const GsParams p, const GsBackwardIO io)
{
    constexpr int PX = 2, NWT = 8 / PX;
    __shared__ __align__(128) float4 s_rec[WPC][2][CHUNK * 3];
    __shared__ __align__(16) float s_acc[WPC][CHUNK * GS_REC_FLOATS];   // block-origin moments of the chunk's records
    __shared__ __align__(16) float s_q[WPC][QW_OFF + 8 * QROW];         // per-quad partial rows of the record being reduced
    __shared__ __align__(8) uint64_t s_bar[WPC][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 (*const ring)[CHUNK * 3] = s_rec[warp];
    uint64_t* const bar = s_bar[warp];
    gs_pdl_trigger();                                           // preprocess_bwd_kernel may become resident (it waits for this grid)
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phases = 0u;
    const size_t HW = (size_t)p.H * p.W;
    const int g = lane >> 2, t = lane & 3;                       // mma fragment coordinates = pixel column, pixel row (and + 4)
    // B operands (TF32 bit patterns of half-integers).  Geometry: columns (S0, S1, S0, S1, S2, 0, S0, 0), lane-in-quad t receives
    // columns 2t, 2t+1 and scales them by (1, x', 1, x'^2).  Selector: k < 4 -> columns 0 and 4, k >= 4 -> columns 2 and 6.
    const float ft = (float)t - 3.5f, ft4 = (float)t + 0.5f;
    const bool c_s0 = g == 0 || g == 2 || g == 6, c_s1 = g == 1 || g == 3, c_s2 = g == 4;
    unsigned bt0 = __float_as_uint(c_s0 ? 1.f : c_s1 ? ft : c_s2 ? ft * ft : 0.f);
    unsigned bt1 = __float_as_uint(c_s0 ? 1.f : c_s1 ? ft4 : c_s2 ? ft4 * ft4 : 0.f);
    unsigned bw0 = __float_as_uint((g == 0 || g == 4) ? 1.f : 0.f), bw1 = __float_as_uint((g == 2 || g == 6) ? 1.f : 0.f);
    const float fxc = (float)g - 3.5f;
    float fsel = t == 0 ? 1.f : t == 1 ? fxc : t == 2 ? 1.f : fxc * fxc;
    asm volatile("" : "+r"(bt0), "+r"(bt1), "+r"(bw0), "+r"(bw1), "+f"(fsel));         // keep them in registers (no rematerialisation in the loop)
    const uint32_t a_parkA = pinned_smem_addr(&s_q[warp][QROW * g + 2 * t]);
    const uint32_t a_parkW = pinned_smem_addr(&s_q[warp][QW_OFF + QROW * g + 2 * (t & 1)]);
    // column sums: lane (c, h) = (lane & 15, lane >> 4) adds rows h, h + 2, h + 4, h + 6 of column c (12 live columns)
    const int red_c = lane & 15, red_h = lane >> 4;
    const bool red_on = red_c < GS_REC_FLOATS;
    const uint32_t a_col = pinned_smem_addr(&s_q[warp][red_c < 8 ? red_h * QROW + red_c : (red_on ? QW_OFF + red_h * QROW + red_c - 8 : 0)]);
    const uint32_t a_out = pinned_smem_addr(&s_acc[warp][red_on ? red_c : 0]);
    const bool red_tag = red_c == 5;                             // column 5 is identically zero: its slot carries the record index
    const float4* __restrict__ my_acc = reinterpret_cast<const float4*>(s_acc[warp]);

    for (;;) {
        long long item = 0;
        if (lane == 0) item = fetch_heavy<NWT>(p, &p.status->q_bwd_heavy);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item == ITEM_DONE) break;

        const int sub = (int)(item & 7);
        const TileCtx tc = tile_ctx(p, item >> 3);
        const int bx0 = tc.tx0 + (sub & 1) * 8, by0 = tc.ty0 + (sub >> 1) * 8;      // origin of this warp's 8 x 8 block
        const unsigned my_blocks = warp_blocks<PX>(sub);
        const size_t vb = (size_t)tc.v * HW;
        const float* __restrict__ bg = p.cams + (size_t)tc.v * GS_CAM_FLOATS + GS_CAM_BG;
        const float b0 = bg[0], b1 = bg[1], b2 = bg[2];

        const int px = bx0 + g;
        const float pxf = (float)px;
        float pyf[PX], T[PX], Q[PX], g0[PX], g1[PX], g2[PX], gd[PX], ga[PX];
        uint32_t last[PX];
        #pragma unroll
        for (int k = 0; k < PX; k++) {
            const int py = by0 + t + 4 * k;
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
        uint32_t m = max(last[0], last[1]);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        const int nmax = (int)min(m, (uint32_t)tc.n);           // deepest contributor of this warp's pixels
        if (nmax == 0) continue;                                // warp-uniform

        const int nchunks = (nmax + CHUNK - 1) / CHUNK;
        const float4* __restrict__ src = p.sorted_rec + tc.start * 3;
        float4* __restrict__ gbase = p.grad2d + (size_t)tc.v * p.N * 3;
        const float fxc0 = (float)bx0 + 3.5f, fyc0 = (float)by0 + 3.5f;               // block centre
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
            unsigned live = __ballot_sync(0xffffffffu, lane < cnt && ((recw[lane * 12 + 11] >> 24) & my_blocks) != 0u);
            int nt = 0;                                         // warp-uniform: records reduced so far (rows of s_acc in use)
            int lastc[PX];                                      // pixel k contributes to records j < lastc[k] of this chunk
            #pragma unroll
            for (int k = 0; k < PX; k++) lastc[k] = (int)last[k] - c * CHUNK;
            while (live) {                                      // warp-uniform
                const int j = msb_index(live);
                live ^= 1u << j;
                const float4 r0 = rec[j * 3], r1 = rec[j * 3 + 1];
                const ColTerms ct = col_terms(r0.z, r0.w, r1.x, __fsub_rn(r0.x, pxf));
                float pw[PX];
                bool hit[PX], any = false;
                #pragma unroll
                for (int k = 0; k < PX; k++) {
                    pw[k] = splat_power(ct, __fsub_rn(r0.y, pyf[k]));                   // same expression, same bits as the forward
                    hit[k] = pw[k] >= r1.w && j < lastc[k];            // the (rare) power > 0 skip is tested on the blend path
                    any = any || hit[k];
                }
                if (!__any_sync(0xffffffffu, any)) continue;
                float tt[PX], wq[4];                            // tt = G dL/dalpha per pixel; wq = sum_k w (g_depth, g_r, g_g, g_b)
                tt[0] = tt[1] = 0.f; wq[0] = wq[1] = wq[2] = wq[3] = 0.f;
                const float4 r2 = rec[j * 3 + 2];
                #pragma unroll
                for (int k = 0; k < PX; k++) {
                    if (hit[k]) {
                        const float G = splat_exp(pw[k]);
                        const float alpha = splat_alpha(r1.y, G);
                        if (alpha >= GS_ALPHA_MIN && pw[k] <= 0.0f) {
                            // T_i = T_{i+1} / (1 - alpha_i) with a Newton-refined reciprocal (see blend_bwd_kernel)
                            const float om = __fsub_rn(1.0f, alpha);             // in [0.01, 1]
                            float ra;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(om));
                            ra = __fmaf_rn(__fmaf_rn(-om, ra, 1.0f), ra, ra);
                            const float Tk = __fmul_rn(T[k], ra);
                            const float w = __fmul_rn(alpha, Tk);
                            const float Pk = __fmaf_rn(g0[k], r2.x, __fmaf_rn(g1[k], r2.y, __fmaf_rn(g2[k], r2.z, __fmaf_rn(gd[k], r1.z, ga[k]))));
                            const float dLda = Tk * Pk - ra * Q[k];
                            Q[k] = __fmaf_rn(w, Pk, Q[k]);
                            tt[k] = G * dLda;
                            wq[0] = __fmaf_rn(w, gd[k], wq[0]);
                            wq[1] = __fmaf_rn(w, g0[k], wq[1]); wq[2] = __fmaf_rn(w, g1[k], wq[2]); wq[3] = __fmaf_rn(w, g2[k], wq[3]);
                            T[k] = Tk;
                        }
                    }
                }
                {
                    unsigned th[PX], tl[PX], wh[4], wl[4];
                    split_tf32(tt[0], th[0], tl[0]); split_tf32(tt[1], th[1], tl[1]);
                    #pragma unroll
                    for (int i = 0; i < 4; i++) split_tf32(wq[i], wh[i], wl[i]);
                    // geometric moments: A = rows (x = g | hi, x = g | lo), k = y: a0 = hi(y = t), a1 = lo(y = t), a2 = hi(y = t + 4), a3 = lo(y = t + 4)
                    float dt[4] = {0.f, 0.f, 0.f, 0.f}, dw[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_tf32_16x8x8(dt, th[0], tl[0], th[1], tl[1], bt0, bt1);
                    // colour / depth moments: a0 = wq0 (row g, k < 4), a1 = wq2 (row g + 8, k < 4), a2 = wq1 (row g, k >= 4), a3 = wq3 (row g + 8, k >= 4)
                    mma_tf32_16x8x8(dw, wl[0], wl[2], wl[1], wl[3], bw0, bw1);
                    mma_tf32_16x8x8(dw, wh[0], wh[2], wh[1], wh[3], bw0, bw1);
                    // lane t of quad g now holds (S_{col 2t}, S_{col 2t+1}) of column x = g (hi in c0/c1, lo in c2/c3) and, for t < 2, the
                    // quad sums (wq0, wq2) [t = 0] or (wq1, wq3) [t = 1] in dw[0], dw[2]
                    sts_v2(a_parkA, fsel * (dt[0] + dt[2]), fsel * (dt[1] + dt[3]));
                    if (t < 2) sts_v2(a_parkW, dw[0], dw[2]);
                    __syncwarp();
                    constexpr int RS = 2 * QROW * 4;                     // byte stride of two rows
                    float acc = (lds_f32<0 * RS>(a_col) + lds_f32<1 * RS>(a_col)) + (lds_f32<2 * RS>(a_col) + lds_f32<3 * RS>(a_col));
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (red_h == 0 && red_on) sts_f32(a_out + (uint32_t)nt * (GS_REC_FLOATS * 4), red_tag ? __int_as_float(j) : acc);
                    nt++;
                    __syncwarp();
                }
            }
            // flush the chunk.  s_acc row r (r-th reduced record, compact) = (M1, My, Mx, Mxy | Myy, [record index], Mxx, 0 |
            // sum w g_d, sum w g_g, sum w g_r, sum w g_b): moments of tt about the block centre.  Re-centre on the splat (dx = ex - x',
            // dy = ey - y'), then gradients of the 2-D record (pix.x, pix.y, conic A, B | C, opacity, depth, _ | rgb), conic B being
            // the true (not halved) derivative; one 16-byte RED per (record, part).
            __syncwarp();
            // one lane per reduced record (nt <= 32 rows): three 16-byte row reads (48-byte stride: conflict-free), the record, three REDs --
            // ONE pass per chunk (the (record, part) -> lane mapping took 1.6 passes of the same length)
            if (lane < nt) {
                const float4 m0 = my_acc[lane * 3], m1 = my_acc[lane * 3 + 1], mw = my_acc[lane * 3 + 2];
                const int j = __float_as_int(m1.y);
                const float4 q0 = rec[j * 3], q1 = rec[j * 3 + 1];
                const int id = __float_as_int(rec[j * 3 + 2].w) & 0x00ffffff;
                const float o = q1.y;
                const float ex = __fsub_rn(q0.x, fxc0), ey = __fsub_rn(q0.y, fyc0);
                const float M1 = m0.x, My = m0.y, Mx = m0.z, Mxy = m0.w, Myy = m1.x, Mxx = m1.z;
                const float Sx = __fmaf_rn(ex, M1, -Mx), Sy = __fmaf_rn(ey, M1, -My);      // sum tt dx, sum tt dy
                const float Sxx = __fmaf_rn(ex, Sx - Mx, Mxx);                              // ex^2 M1 - 2 ex Mx + Mxx
                const float Sxy = __fmaf_rn(ex, Sy, __fmaf_rn(-ey, Mx, Mxy));                // ex ey M1 - ex My - ey Mx + Mxy
                const float Syy = __fmaf_rn(ey, Sy - My, Myy);
                float4* __restrict__ dst = gbase + (size_t)id * 3;
                // q0.z, q0.w, q1.x = -A/2, -B, -C/2 (pre-scaled conic): dL/dpix = -o (A Sx + B Sy), -o (C Sy + B Sx)
#if GS_PRESCALE
                red_add_v4(dst, make_float4(o * __fmaf_rn(2.0f * q0.z, Sx, q0.w * Sy), o * __fmaf_rn(2.0f * q1.x, Sy, q0.w * Sx), -0.5f * o * Sxx, -o * Sxy));
#else
                red_add_v4(dst, make_float4(-o * (q0.z * Sx + q0.w * Sy), -o * (q1.x * Sy + q0.w * Sx), -0.5f * o * Sxx, -o * Sxy));
#endif
                red_add_v4(dst + 1, make_float4(-0.5f * o * Syy, M1, mw.x, 0.f));
                red_add_v4(dst + 2, make_float4(mw.z, mw.y, mw.w, 0.f));
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
    gs_launch_dependent(blend_fwd_kernel<PX>, dim3(grid), dim3(WPC * 32), 0, s, p, color, depth, alpha);
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
        case 2: {
#if GS_BWD_MMA
            static thread_local int grid = 0, grid_sms = 0;
            if (grid == 0 || grid_sms != num_sms) { grid = resident_ctas((const void*)blend_bwd_mma_kernel, WPC * 32, num_sms, 4); grid_sms = num_sms; }
            blend_bwd_mma_kernel<<<grid, WPC * 32, 0, s>>>(p, io);
#else
            launch_bwd<2>(p, io, num_sms, s);
#endif
            break;
        }
        default: launch_bwd<4>(p, io, num_sms, s); break;
    }
}
