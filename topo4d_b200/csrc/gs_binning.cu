// gs_binning.cu -- tile scan (K2) and per-tile depth sort + record gather (K4).
//
// Replaces upstream's InclusiveSum + duplicateWithKeys + 64-bit DeviceRadixSort + identifyTileRanges
// (SURVEY.md 2.2 K2-K5).  Instead of ~7 radix passes over 12-byte pairs in HBM, instances are
// bucketed by tile with atomics (gs_preprocess.cu), the tile histogram is scanned here (tile
// ranges fall out directly), and every tile's short list is sorted by (depth_bits, index) inside
// shared memory by one CTA.  Keys are distinct 64-bit integers, so the result is exactly the order
// a stable radix sort on (tile | depth_bits) of index-ordered duplicates gives: bit-exact lists.
// The same CTA then gathers the 48-byte per-Gaussian records into tile order, so both blend passes
// stream each chunk with one contiguous bulk-async copy.
#include "gs_common.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;   // SCAN_THREADS * SCAN_ITEMS == GS_SCAN_ELEMS_PER_BLOCK

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// phase A: per-block totals (also the per-tile maximum for the status block)
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const GsParams p)
{
    __shared__ uint32_t s_sum[32];
    __shared__ uint32_t s_max[32];
    const long long n = p.total_tiles;
    const long long base = (long long)blockIdx.x * GS_SCAN_ELEMS_PER_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t sum = 0, mx = 0;
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const long long e = base + k;
        const uint32_t c = e < n ? p.tile_count[e] : 0u;
        sum += c; mx = max(mx, c);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_sum[w] = sum; s_max[w] = mx; }
    __syncthreads();
    if (w == 0) {
        sum = s_sum[lane]; mx = s_max[lane];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            p.block_sums[blockIdx.x] = sum;
            atomicMax(&p.status->max_tile_instances, (int)mx);
        }
    }
}

// phase B: exclusive scan of every block's chunk, offset by the sum of the preceding blocks.
// FUSED = true: phases A and B in ONE launch.  Every block publishes its total, then waits until all blocks have
// (a grid-wide counter in the status block); legal because the launcher only picks this variant when the whole
// grid is co-resident (scan_blocks <= SM count, one 1024-thread block per SM always fits).
template <bool FUSED>
__global__ void __launch_bounds__(SCAN_THREADS) scan_write_kernel(const GsParams p)
{
    __shared__ unsigned long long s_red[32];
    __shared__ uint32_t s_warp[32];
    __shared__ unsigned long long s_prefix;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    gs_pdl_trigger();                                   // scatter_kernel may become resident now (it waits for this grid's completion)
    if (FUSED) {
        __shared__ uint32_t s_sum[32];
        __shared__ uint32_t s_max[32];
        const long long n0 = p.total_tiles;
        const long long base0 = (long long)blockIdx.x * GS_SCAN_ELEMS_PER_BLOCK + threadIdx.x * SCAN_ITEMS;
        uint32_t sum = 0, mx = 0;
        #pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const long long e = base0 + k;
            const uint32_t c = e < n0 ? p.tile_count[e] : 0u;
            sum += c; mx = max(mx, c);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) { s_sum[w] = sum; s_max[w] = mx; }
        __syncthreads();
        if (w == 0) {
            sum = s_sum[lane]; mx = s_max[lane];
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (lane == 0) {
                p.block_sums[blockIdx.x] = sum;
                atomicMax(&p.status->max_tile_instances, (int)mx);
                __threadfence();
                atomicAdd(&p.status->scan_done, 1u);
                volatile unsigned int* done = &p.status->scan_done;
                while (*done < gridDim.x) { }
                __threadfence();
            }
        }
        __syncthreads();
    }
    // prefix of preceding blocks (64-bit so an overflowing total is detected, not wrapped)
    unsigned long long pre = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += SCAN_THREADS) pre += __ldcg(p.block_sums + b);
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if (lane == 0) s_red[w] = pre;
    __syncthreads();
    if (w == 0) {
        pre = s_red[lane];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
        if (lane == 0) s_prefix = pre;
    }
    __syncthreads();
    pre = s_prefix;

    const long long n = p.total_tiles;   // tile_start has n+1 entries
    const long long base = (long long)blockIdx.x * GS_SCAN_ELEMS_PER_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS], tsum = 0;
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { const long long e = base + k; c[k] = e < n ? p.tile_count[e] : 0u; tsum += c[k]; }
    const uint32_t incl = warp_incl_scan(tsum, lane);
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) { uint32_t x = s_warp[lane]; x = warp_incl_scan(x, lane); s_warp[lane] = x; }
    __syncthreads();
    unsigned long long run = pre + (w > 0 ? s_warp[w - 1] : 0u) + (incl - tsum);
    uint32_t nact = 0;                                   // packed: long count << 16 | short count (<= 4096 per block)
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const long long e = base + k;
        if (e <= n) p.tile_start[e] = (uint32_t)(run > 0xffffffffull ? 0xffffffffull : run);
        if (e < n) p.tile_fill[e] = 0u;
        if (e < n && c[k] > 0u) nact += c[k] >= GS_LONG_TILE ? (1u << 16) : 1u;
        run += c[k];
    }
    // compact the non-empty tiles into active_tiles[]: long lists from the front, short ones from the back (one
    // atomic pair per block; the order inside each class is arbitrary, which only changes the order tiles are
    // worked on, never a result)
    __shared__ uint32_t s_awarp[32];
    __shared__ uint32_t s_abase[2];
    const uint32_t aincl = warp_incl_scan(nact, lane);
    __syncthreads();
    if (lane == 31) s_awarp[w] = aincl;
    __syncthreads();
    if (w == 0) { uint32_t x = s_awarp[lane]; x = warp_incl_scan(x, lane); s_awarp[lane] = x; }
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) {
        s_abase[0] = atomicAdd(&p.status->num_long, s_awarp[31] >> 16);
        s_abase[1] = atomicAdd(&p.status->num_short, s_awarp[31] & 0xffffu);
    }
    __syncthreads();
    const uint32_t excl = (w > 0 ? s_awarp[w - 1] : 0u) + (aincl - nact);
    uint32_t pos_long = s_abase[0] + (excl >> 16), pos_short = s_abase[1] + (excl & 0xffffu);
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const long long e = base + k;
        if (e < n && c[k] > 0u) {
            if (c[k] >= GS_LONG_TILE) p.active_tiles[pos_long++] = (uint32_t)e;
            else p.active_tiles[n - 1 - (long long)(pos_short++)] = (uint32_t)e;
        }
    }
    if (base <= n && n < base + SCAN_ITEMS) {   // the thread that owns element n holds the grand total
        unsigned long long total = pre + (w > 0 ? s_warp[w - 1] : 0u) + (incl - tsum);
        for (int k = 0; k < SCAN_ITEMS && base + k < n; k++) total += c[k];
        p.status->num_instances = total;
        p.status->cap_instances = (unsigned long long)p.cap;
        p.status->overflow = total > (unsigned long long)p.cap ? 1 : 0;
    }
}

// ---- K4 ----
#ifndef GS_SORT_UNROLL
#define GS_SORT_UNROLL 1
#endif
constexpr int SORT_UNROLL = GS_SORT_UNROLL;     // instances in flight per thread in the gather epilogue
constexpr int SORT_THREADS = 128;
constexpr int SORT_REG_KEYS = 2048;    // lists up to here are sorted in registers (<= 16 keys per thread) ...
constexpr int SORT_LONG_THREADS = 1024;
constexpr int SORT_LONG_KEYS = 16384;  // ... up to here by a 1024-thread CTA in 128 KB of shared memory (sort_gather_long_kernel: the reference's
                                       // texture regime -- millions of pixel-sized splats -- has thousands of entries per tile); beyond, in place in HBM/L2

// Generic bitonic network in the "flip / disperse" form on a key array in memory (used only for lists longer than
// SORT_REG_KEYS): every compare-exchange moves the smaller key to the lower index, so virtual +inf padding above `n`
// never moves and arbitrary n needs no real padding.
template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort_mem(KeyPtr keys, int n, int tid, int nthreads)
{
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    const int half = n2 >> 1;
    for (int lk = 1; (1 << lk) <= n2; lk++) {
        const int k = 1 << lk, hk = k >> 1;
        for (int t = tid; t < half; t += nthreads) {
            const int blk = t >> (lk - 1), off = t & (hk - 1);
            const int i = (blk << lk) + off, j = (blk << lk) + (k - 1 - off);
            if (j < n) {
                const unsigned long long a = keys[i], b = keys[j];
                if (a > b) { keys[i] = b; keys[j] = a; }
            }
        }
        __syncthreads();
        for (int lj = lk - 2; lj >= 0; lj--) {
            const int jj = 1 << lj;
            for (int t = tid; t < half; t += nthreads) {
                const int i = ((t >> lj) << (lj + 1)) + (t & (jj - 1)), j = i + jj;
                if (j < n) {
                    const unsigned long long a = keys[i], b = keys[j];
                    if (a > b) { keys[i] = b; keys[j] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// Register-resident bitonic sort of 128 * KPT keys by one 128-thread CTA: thread t owns positions [t*KPT, (t+1)*KPT).
// Partner distances < KPT are compare-exchanges between a thread's own registers, distances < 32*KPT one
// shuffle-xor per key, and only the two largest distances (3 of the ~50 steps) cross warps through shared memory.
// Keys past the list end are +inf in registers, so the textbook network (direction bit = position & block size)
// needs no memory padding.  Fully unrolled: every distance / direction test is an immediate.
__host__ __device__ constexpr int ilog2_c(int x) { return x <= 1 ? 0 : 1 + ilog2_c(x >> 1); }

__device__ __forceinline__ void cex(unsigned long long& a, unsigned long long& b, bool asc)
{
    const bool sw = (a > b) == asc;
    const unsigned long long lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}

template <int KPT>
__device__ __forceinline__ void bitonic_sort_regs(unsigned long long (&key)[KPT], int tid, unsigned long long* s_x)
{
    constexpr int LOGK = ilog2_c(KPT), LOGN = LOGK + 7;
    #pragma unroll
    for (int lk = 1; lk <= LOGN; lk++) {
        #pragma unroll
        for (int lj = lk - 1; lj >= 0; lj--) {
            if (lj < LOGK) {
                #pragma unroll
                for (int a = 0; a < KPT; a++) {
                    if ((a & (1 << lj)) == 0) {
                        bool asc;
                        if (lk == LOGN) asc = true;
                        else if (lk < LOGK) asc = ((a >> lk) & 1) == 0;
                        else asc = ((tid >> (lk - LOGK)) & 1) == 0;
                        cex(key[a], key[a | (1 << lj)], asc);
                    }
                }
            } else {
                const int m = 1 << (lj - LOGK);                       // partner thread = tid ^ m
                const bool asc = lk == LOGN ? true : ((tid >> (lk - LOGK)) & 1) == 0;
                const bool take_min = asc == ((tid & m) == 0);
                if (lj < LOGK + 5) {
                    #pragma unroll
                    for (int a = 0; a < KPT; a++) {
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key[a], m);
                        key[a] = ((o < key[a]) == take_min) ? o : key[a];
                    }
                } else {
                    #pragma unroll
                    for (int a = 0; a < KPT; a++) s_x[a * SORT_THREADS + tid] = key[a];
                    __syncthreads();
                    #pragma unroll
                    for (int a = 0; a < KPT; a++) {
                        const unsigned long long o = s_x[a * SORT_THREADS + (tid ^ m)];
                        key[a] = ((o < key[a]) == take_min) ? o : key[a];
                    }
                    __syncthreads();
                }
            }
        }
    }
}

template <int KPT>
__device__ __forceinline__ void sort_tile_regs(const unsigned long long* __restrict__ gk, int n, int tid, unsigned long long* s_keys)
{
    unsigned long long key[KPT];
    #pragma unroll
    for (int a = 0; a < KPT; a++) {
        const int i = tid * KPT + a;
        key[a] = i < n ? gk[i] : ~0ull;
    }
    bitonic_sort_regs<KPT>(key, tid, s_keys);
    #pragma unroll
    for (int a = 0; a < KPT; a++) {
        const int i = tid * KPT + a;
        if (i < n) s_keys[i] = key[a];
    }
    __syncthreads();
}

// which of the tile's eight 8x4 pixel blocks can this splat reach at all?  Exact up to a safety margin: the minimum
// of q(d) = 1/2 d^T conic d over the block's (continuous) rectangle against tau = -thr (thr already carries its own
// margin).  q is convex, so the minimum is 0 if the centre is inside, else it lies on one of the four edges (a 1-D
// clamped parabola each).  The tile has only 4 distinct vertical and 8 distinct horizontal block edges: the per-line
// terms are computed once and shared by the blocks along the line.  The blend kernels skip a block whose bit is
// clear without evaluating a single pixel; a clear bit can never hide a contributor.
__device__ __forceinline__ unsigned block_reach_mask(const float4 g0, const float4 g1, float tx0, float ty0)
{
    const float A = g0.z, B = g0.w, C = g1.x, tau = -g1.w;
    if (!(A > 0.0f && C > 0.0f && A * C - B * B > 0.0f && tau < 3.0e38f)) return 0xffu;
    if (!(tau >= 0.0f)) return 0u;
    const float lim = tau * 1.0001f + 2.0e-3f;
    const float nBC = -B / C, nBA = -B / A, hA = 0.5f * A, hC = 0.5f * C;
    // d = centre - pixel.  Column c of blocks spans dx in [X[2c+1], X[2c]], row r spans dy in [Y[2r+1], Y[2r]].
    float X[4], Y[8], vq0[4], vq1[4], vs[4], hq0[8], hq1[8], hs[8];
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        X[i] = g0.x - (tx0 + (float)(8 * (i >> 1) + 7 * (i & 1)));
        vq0[i] = hA * X[i] * X[i]; vq1[i] = B * X[i]; vs[i] = nBC * X[i];      // q(X, t) = vq0 + t (hC t + vq1), argmin t = vs
    }
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        Y[i] = g0.y - (ty0 + (float)(4 * (i >> 1) + 3 * (i & 1)));
        hq0[i] = hC * Y[i] * Y[i]; hq1[i] = B * Y[i]; hs[i] = nBA * Y[i];      // q(t, Y) = hq0 + t (hA t + hq1), argmin t = hs
    }
    unsigned mask = 0u;
    #pragma unroll
    for (int b = 0; b < 8; b++) {
        const int c = b & 1, r = b >> 1;
        const float dx1 = X[2 * c], dx0 = X[2 * c + 1], dy1 = Y[2 * r], dy0 = Y[2 * r + 1];
        float qmin = 3.0e38f;
        #pragma unroll
        for (int e = 0; e < 2; e++) {
            const int vi = 2 * c + e, hi = 2 * r + e;
            const float ty = fminf(fmaxf(vs[vi], dy0), dy1);
            qmin = fminf(qmin, fmaf(ty, fmaf(hC, ty, vq1[vi]), vq0[vi]));
            const float tx = fminf(fmaxf(hs[hi], dx0), dx1);
            qmin = fminf(qmin, fmaf(tx, fmaf(hA, tx, hq1[hi]), hq0[hi]));
        }
        if (dx0 <= 0.0f && dx1 >= 0.0f && dy0 <= 0.0f && dy1 >= 0.0f) qmin = 0.0f;
        if (qmin <= lim) mask |= 1u << b;
    }
    return mask;
}

__global__ void __launch_bounds__(SORT_THREADS, 8) sort_gather_kernel(const GsParams p)
{
    __shared__ __align__(16) unsigned long long s_keys[SORT_REG_KEYS];
    __shared__ long long s_tile;
    const int tid = threadIdx.x;
    gs_pdl_wait();
    gs_pdl_trigger();
    // non-empty tiles come from the device-side queue the scan kernel filled (dynamic load balance; asking for the next tile ahead of
    // time was measured: it hides two L2 round trips per tile but hands the leftover tiles to the CTAs that hold the longest lists --
    // 0.034 -> 0.044 ms at 3 views, nothing at 24)
    for (;;) {
        if (tid == 0) {
            s_tile = gs_active_tile(p, atomicAdd(&p.status->q_sort, 1u));
        }
        __syncthreads();
        const long long tg = s_tile;
        __syncthreads();
        if (tg < 0) break;
        unsigned long long start = p.tile_start[tg], end = p.tile_start[tg + 1];
        if (end > (unsigned long long)p.cap) end = (unsigned long long)p.cap;
        if (start >= end) continue;
        const int n = (int)(end - start);
        const int v = (int)((unsigned)tg / (unsigned)p.tiles);             // V * tiles < 2^31 (validated on the host)
        unsigned long long* gk = p.pairs + start;
        unsigned long long* sorted = s_keys;
        if (n <= 128 * 2) sort_tile_regs<2>(gk, n, tid, s_keys);
        else if (n <= 128 * 4) sort_tile_regs<4>(gk, n, tid, s_keys);
        else if (n <= 128 * 8) sort_tile_regs<8>(gk, n, tid, s_keys);
        else if (n <= 128 * 16) sort_tile_regs<16>(gk, n, tid, s_keys);
        else if (p.sort_long && n <= SORT_LONG_KEYS) continue;     // sort_gather_long_kernel owns this list (block-uniform)
        else {
            bitonic_sort_mem(gk, n, tid, SORT_THREADS);     // in place in L2-resident global memory
            sorted = gk;
        }
        const float4* __restrict__ geom = p.geom + (size_t)v * p.N * 3;
        float4* __restrict__ rec = p.sorted_rec + start * 3;
        const int tl = (int)(tg - (long long)v * p.tiles);
        const float tx0 = (float)((tl % p.tiles_x) * GS_TILE), ty0 = (float)((tl / p.tiles_x) * GS_TILE);
        // one thread per instance: the three 16-byte parts of its record are requested together (ONE L2 round trip per instance, and
        // two instances in flight per thread), then the exact block-reach mask, then the record leaves -- conic pre-scaled for the blend
        // kernels (-A/2, -B, -C/2: exact), reach mask in bits 24..31 of the index word (indices are < 2^24, validated on the host)
        #pragma unroll SORT_UNROLL
        for (int k = tid; k < n; k += SORT_THREADS) {
            const uint32_t id = (uint32_t)sorted[k] & 0x00ffffffu;
            const float4* __restrict__ gr = geom + (size_t)id * 3;
            float4 g0 = __ldg(gr), g1 = __ldg(gr + 1), g2 = __ldg(gr + 2);
            p.sorted_ids[start + k] = id;
            const unsigned mask = block_reach_mask(g0, g1, tx0, ty0);
#if GS_PRESCALE
            g0.z *= -0.5f; g0.w = -g0.w; g1.x *= -0.5f;
#endif
            g2.w = __uint_as_float(id | (mask << 24));
            float4* __restrict__ o = rec + (size_t)k * 3;
            o[0] = g0; o[1] = g1; o[2] = g2;
        }
        __syncthreads();   // s_keys is reused by the next tile
    }
    if (tid == 0) gs_queue_release(&p.status->q_sort, &p.status->done_sort, gridDim.x);
}

// Lists of 2049 .. 16384 keys: one 1024-thread CTA per tile, keys in (dynamic) shared memory, textbook bitonic network with
// +inf padding up to the next power of two, then the same mask + gather epilogue as sort_gather_kernel.  Walks only the
// front of active_tiles[] (the lists >= GS_LONG_TILE).  Launched when the host expects such lists (GsProblem.hints) or does
// not know; sort_gather_kernel leaves them alone when this kernel runs and falls back to the in-place global network otherwise.
__global__ void __launch_bounds__(SORT_LONG_THREADS, 1) sort_gather_long_kernel(const GsParams p)
{
    extern __shared__ __align__(16) unsigned long long s_long[];
    __shared__ long long s_tile;
    const int tid = threadIdx.x;
    gs_pdl_wait();
    gs_pdl_trigger();
    for (;;) {
        if (tid == 0) {
            const unsigned q = atomicAdd(&p.status->q_sort_long, 1u);
            s_tile = q < p.status->num_long ? (long long)p.active_tiles[q] : -1;
        }
        __syncthreads();
        const long long tg = s_tile;
        __syncthreads();
        if (tg < 0) break;
        unsigned long long start = p.tile_start[tg], end = p.tile_start[tg + 1];
        if (end > (unsigned long long)p.cap) end = (unsigned long long)p.cap;
        if (end <= start + SORT_REG_KEYS || end > start + SORT_LONG_KEYS) continue;
        const int n = (int)(end - start);
        int n2 = SORT_REG_KEYS * 2;
        while (n2 < n) n2 <<= 1;
        const unsigned long long* __restrict__ gk = p.pairs + start;
        for (int k = tid; k < n2; k += SORT_LONG_THREADS) s_long[k] = k < n ? gk[k] : ~0ull;
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (n2 >> 1); t += SORT_LONG_THREADS) {
                    const int i = 2 * t - (t & (j - 1)), l = i + j;
                    const unsigned long long a = s_long[i], b = s_long[l];
                    if ((a > b) == ((i & k) == 0)) { s_long[i] = b; s_long[l] = a; }
                }
                __syncthreads();
            }
        const int v = (int)((unsigned)tg / (unsigned)p.tiles);             // V * tiles < 2^31 (validated on the host)
        const float4* __restrict__ geom = p.geom + (size_t)v * p.N * 3;
        float4* __restrict__ rec = p.sorted_rec + start * 3;
        const int tl = (int)(tg - (long long)v * p.tiles);
        const float tx0 = (float)((tl % p.tiles_x) * GS_TILE), ty0 = (float)((tl / p.tiles_x) * GS_TILE);
        #pragma unroll 2
        for (int k = tid; k < n; k += SORT_LONG_THREADS) {          // as in sort_gather_kernel: record in, reach mask, record out
            const uint32_t id = (uint32_t)s_long[k] & 0x00ffffffu;
            const float4* __restrict__ gr = geom + (size_t)id * 3;
            float4 g0 = __ldg(gr), g1 = __ldg(gr + 1), g2 = __ldg(gr + 2);
            p.sorted_ids[start + k] = id;
            const unsigned mask = block_reach_mask(g0, g1, tx0, ty0);
#if GS_PRESCALE
            g0.z *= -0.5f; g0.w = -g0.w; g1.x *= -0.5f;
#endif
            g2.w = __uint_as_float(id | (mask << 24));
            float4* __restrict__ o = rec + (size_t)k * 3;
            o[0] = g0; o[1] = g1; o[2] = g2;
        }
        __syncthreads();
    }
    if (tid == 0) gs_queue_release(&p.status->q_sort_long, &p.status->done_sort_long, gridDim.x);
}

}  // namespace

void gs_launch_tile_scan(const GsParams& p, int num_sms, cudaStream_t s)
{
    // FUSED = one launch whose blocks meet at a grid-wide counter: only legal when the whole grid is co-resident.  A single
    // block needs no barrier; several blocks are launched COOPERATIVELY, which makes the runtime guarantee co-residency (or
    // refuse: MPS / green-context SM limits, other persistent kernels holding the SMs) instead of assuming it from the SM count;
    // anything else takes the two-launch path.
    if (p.scan_blocks == 1) { scan_write_kernel<true><<<1, SCAN_THREADS, 0, s>>>(p); return; }
    static thread_local int coop_blocks = -1, coop_sms = 0;
    if (coop_blocks < 0 || coop_sms != num_sms) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)scan_write_kernel<true>, SCAN_THREADS, 0) != cudaSuccess) per_sm = 0;
        coop_blocks = per_sm * num_sms;
        coop_sms = num_sms;
    }
    if (p.scan_blocks <= coop_blocks) {
        GsParams q = p;
        void* args[] = {(void*)&q};
        if (cudaLaunchCooperativeKernel((const void*)scan_write_kernel<true>, dim3(p.scan_blocks), dim3(SCAN_THREADS), args, 0, s) == cudaSuccess) return;
        (void)cudaGetLastError();                  // refused (resources held elsewhere): fall through to the two-launch path
    }
    scan_reduce_kernel<<<p.scan_blocks, SCAN_THREADS, 0, s>>>(p);
    scan_write_kernel<false><<<p.scan_blocks, SCAN_THREADS, 0, s>>>(p);
}

void gs_launch_sort_gather(const GsParams& p, int num_sms, cudaStream_t s)
{
    if (p.sort_long) {
        static thread_local bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute((const void*)sort_gather_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_LONG_KEYS * 8);
            attr_set = true;
        }
        gs_launch_dependent(sort_gather_long_kernel, dim3(num_sms), dim3(SORT_LONG_THREADS), SORT_LONG_KEYS * 8, s, p);
    }
    long long blocks = p.total_tiles;
    const long long maxb = (long long)num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    gs_launch_dependent(sort_gather_kernel, dim3((unsigned)blocks), dim3(SORT_THREADS), 0, s, p);
}
