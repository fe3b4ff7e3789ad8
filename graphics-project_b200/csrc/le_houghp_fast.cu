// le_houghp_fast.cu -- batched exact PPHT (a6) for numangle <= 192 (one thread per angle).
//
// Same sequential semantics as cv2 (and as k_ppht in le_houghp.cu), restructured so that the
// dependent global-memory round trips are paid once per BATCH of picks instead of once per pick:
//   * thread 0 replays cv::RNG and the swap-remove list for 32 candidate picks at a time; warp 0 then
//     reads the 32 mask bytes in one round and keeps the "alive" candidates as a bit mask (refreshed
//     after every line walk, the only event that clears mask pixels);
//   * up to PF_K alive picks are voted together: every thread (= angle) loads its PF_K accumulator
//     cells at once, derives the value each pick WOULD see (cell + 1 + earlier picks of the batch
//     in the same cell), and the PF_K arg-max reductions (REDUX + one barrier) run side by side;
//   * the first pick whose maximum reaches the threshold ends the batch: votes are committed up to
//     and including it (fire-and-forget RED), later picks of the batch are simply re-evaluated after
//     the walk -- nothing is ever rolled back, so the result is bit-identical to the one-by-one loop.
// Accumulator cells are u16 biased by 0x4040 and packed two per 32-bit word: +-1 on either half never
// carries into the other, so commits and un-votes are RED.ADD on the word (cells go "negative" in cv2
// when a never-voted pixel on an accepted line is un-voted; the bias absorbs that).
#include "le_common.cuh"

#define PF_THREADS 192
#define PF_K 8
#define PF_NZ_SMEM 12288
#define PF_MAXT_WORDS 512
#define PF_BIAS 0x4040

struct PFShared {
    u32 nz[PF_NZ_SMEM];
    u32 walk_bits[2][PF_MAXT_WORDS];
    u32 red[2][PF_THREADS / 32][PF_K];
    u32 cand[32];
    u32 alive;
    int ncand;
    int t_end[2];
    int end_xy[2][2];
    int nlines, ntrace;
};

static __device__ __forceinline__ u32 pf_rng_next(u64 &st)
{
    st = (u64)(u32)st * 4164903690ULL + (st >> 32);
    return (u32)st;
}

__global__ void __launch_bounds__(PF_THREADS)
k_ppht_fast(u8 *__restrict__ mask, u32 *__restrict__ nz_g, u32 *__restrict__ accum, const float *__restrict__ trig,
            int *__restrict__ counters, int32_t *__restrict__ segs, int32_t *__restrict__ counts, int max_segs,
            int h, int w, int numangle, int stride, int threshold, int line_length, int line_gap, int32_t *trace,
            int trace_cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PFShared &S = *reinterpret_cast<PFShared *>(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t npx = (size_t)h * w;
    u8 *mk = mask + (size_t)f * npx;
    const int hw = stride >> 1;  // words per accumulator row
    u32 *acc = accum + (size_t)f * numangle * hw;
    int count = counters[f * LE_C_STRIDE + LE_C_AUX0];
    const int *rlo_g = (const int *)(trig + 2 * numangle);
    u32 *nz = nz_g + (size_t)f * npx;
    if (count <= PF_NZ_SMEM) {
        for (int i = tid; i < count; i += PF_THREADS) S.nz[i] = nz[i];
        nz = S.nz;
    }
    if (tid == 0) { S.nlines = 0; S.ntrace = 0; }
    const bool mine = tid < numangle;
    float tc = 0.f, ts = 0.f;
    int rlo = 0;
    if (mine) {
        tc = trig[tid];
        ts = trig[numangle + tid];
        rlo = rlo_g[tid];
    }
    u32 *myrow = acc + (size_t)(mine ? tid : 0) * hw;
    u64 rng = ~0ULL;
    int par = 0;
    const int shift = 16;
#ifdef LE_PPHT_PROFILE  // phase cycle counters of frame 0 (scripts/dbg_ppht_time.py); off in the product build
    long long dbg[6] = {0, 0, 0, 0, 0, 0}, t_a = 0;
#define PF_TICK(i) do { if (trace && tid == 0) { long long t_ = clock64(); dbg[i] += t_ - t_a; t_a = t_; } } while (0)
#else
#define PF_TICK(i) do { } while (0)
#endif
    __syncthreads();
#ifdef LE_PPHT_PROFILE
    if (trace && tid == 0) t_a = clock64();
#endif

    for (;;) {
        // ---- snapshot of the next (up to) 32 candidate picks: the visiting order does not depend on the
        // image, only on the number of edge pixels
        if (tid == 0) {
            int m = 0;
            while (m < 32 && count > 0) {
                int idx = (int)(pf_rng_next(rng) % (u32)count);
                u32 pt = nz[idx];
                nz[idx] = nz[count - 1];
                count--;
                S.cand[m++] = pt;
            }
            S.ncand = m;
        }
        __syncthreads();
        const int ncand = S.ncand;
        if (ncand == 0) break;
        if (wid == 0) {
            bool a = false;
            if (lane < ncand) {
                u32 pt = S.cand[lane];
                a = mk[(size_t)(pt >> 16) * w + (pt & 0xffff)] != 0;
            }
            u32 b = __ballot_sync(~0u, a);
            if (lane == 0) S.alive = b;
        }
        __syncthreads();
        PF_TICK(0);
        u32 alive = S.alive;  // every thread tracks the consumption identically

        while (alive) {
            // ---- vote up to PF_K alive picks together
            const int npk = min(__popc(alive), PF_K);
            u32 pts[PF_K], wordv[PF_K];
            int cellb[PF_K];
            {
                u32 tmp = alive;
#pragma unroll
                for (int k = 0; k < PF_K; k++) {
                    int pos = tmp ? __ffs(tmp) - 1 : 0;
                    tmp &= tmp - 1;
                    pts[k] = S.cand[pos];
                }
            }
#pragma unroll
            for (int k = 0; k < PF_K; k++) {
                cellb[k] = 0;
                wordv[k] = 0;
                if (k < npk && mine) {
                    float x = (float)(pts[k] & 0xffff), y = (float)(pts[k] >> 16);
                    int b = __float2int_rn(__fadd_rn(__fmul_rn(x, tc), __fmul_rn(y, ts))) - rlo;
                    cellb[k] = b;
                    wordv[k] = __ldcg(myrow + (b >> 1));
                }
            }
            u32 keys[PF_K];
#pragma unroll
            for (int k = 0; k < PF_K; k++) {
                int dups = 0;
#pragma unroll
                for (int j = 0; j < k; j++) dups += (j < npk && cellb[j] == cellb[k]) ? 1 : 0;
                int val = (int)((wordv[k] >> ((cellb[k] & 1) * 16)) & 0xffff) - PF_BIAS + 1 + dups;
                u32 key = (k < npk && mine) ? (((u32)(val + 0x8000) << 8) | (u32)(255 - tid)) : 0u;
                key = __reduce_max_sync(~0u, key);
                keys[k] = key;
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < PF_K; k++) S.red[par][wid][k] = keys[k];
            }
            __syncthreads();
            int m = npk, max_val = 0, max_n = 0;
#pragma unroll
            for (int k = 0; k < PF_K; k++) {
                u32 best = 0;
#pragma unroll
                for (int q = 0; q < PF_THREADS / 32; q++) best = max(best, S.red[par][q][k]);
                int v = (int)(best >> 8) - 0x8000;
                if (k < npk && m == npk && v >= threshold) {
                    m = k;
                    max_val = v;
                    max_n = 255 - (int)(best & 255);
                }
            }
            par ^= 1;
            // commit the votes of picks 0..m (inclusive when m is a real pick)
            const int ncommit = m < npk ? m + 1 : npk;
#pragma unroll
            for (int k = 0; k < PF_K; k++)
                if (k < ncommit && mine) atomicAdd(myrow + (cellb[k] >> 1), 1u << ((cellb[k] & 1) * 16));
            // drop the consumed picks from the alive set
            for (int k = 0; k < ncommit; k++) alive &= alive - 1;
            PF_TICK(1);
#ifdef LE_PPHT_PROFILE
            if (trace && tid == 0) { dbg[4] += 1; dbg[5] += ncommit; }
#endif
            if (m == npk) continue;  // no pick of this batch produced a line

            // ---- line walk for pick m (identical to the one-by-one kernel)
            const int j = (int)(pts[m] & 0xffff), i = (int)(pts[m] >> 16);
            const float a = -trig[numangle + max_n], b = trig[max_n];
            int x0 = j, y0 = i, dx0, dy0;
            bool xflag;
            if (fabsf(a) > fabsf(b)) {
                xflag = true;
                dx0 = a > 0 ? 1 : -1;
                dy0 = __float2int_rn(__fdiv_rn(__fmul_rn(b, 65536.f), fabsf(a)));
                y0 = (y0 << shift) + (1 << (shift - 1));
            } else {
                xflag = false;
                dy0 = b > 0 ? 1 : -1;
                dx0 = __float2int_rn(__fdiv_rn(__fmul_rn(a, 65536.f), fabsf(b)));
                x0 = (x0 << shift) + (1 << (shift - 1));
            }
            if (wid < 2) {
                const int k = wid;
                const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
                int gap = 0, t_end = 0, ex = j, ey = i;
                bool stop = false;
                for (int base = 0; !stop; base += 32) {
                    int t = base + lane;
                    int x = x0 + t * dx, y = y0 + t * dy;
                    int j1 = xflag ? x : (x >> shift), i1 = xflag ? (y >> shift) : y;
                    bool inb = j1 >= 0 && j1 < w && i1 >= 0 && i1 < h;
                    bool on = inb && mk[(size_t)i1 * w + j1] != 0;
                    u32 bm = __ballot_sync(~0u, on);
                    u32 below = bm & ((1u << lane) - 1);
                    int run = below ? lane - (31 - __clz(below)) : gap + lane + 1;
                    bool st = !inb || (!on && run > line_gap);
                    u32 bs = __ballot_sync(~0u, st);
                    u32 valid = bs ? ((1u << (__ffs(bs) - 1)) - 1) : ~0u;
                    u32 bmv = bm & valid;
                    if ((base >> 5) < PF_MAXT_WORDS && lane == 0) S.walk_bits[k][base >> 5] = bmv;
                    if (bmv) {
                        int hb = 31 - __clz(bmv);
                        t_end = base + hb;
                        ex = __shfl_sync(~0u, j1, hb);
                        ey = __shfl_sync(~0u, i1, hb);
                        gap = 31 - hb;
                    } else gap += 32;
                    stop = bs != 0;
                }
                if (lane == 0) {
                    S.t_end[k] = t_end;
                    S.end_xy[k][0] = ex;
                    S.end_xy[k][1] = ey;
                }
            }
            __syncthreads();
            PF_TICK(2);
            const int e0x = S.end_xy[0][0], e0y = S.end_xy[0][1], e1x = S.end_xy[1][0], e1y = S.end_xy[1][1];
            const bool good = abs(e1x - e0x) >= line_length || abs(e1y - e0y) >= line_length;
            if (trace && f == 0 && tid == 0) {
                int tp = S.ntrace++;
                if (tp < trace_cap) {
                    int32_t *o = trace + tp * 10;
                    o[0] = j; o[1] = i; o[2] = max_val; o[3] = max_n; o[4] = e0x; o[5] = e0y; o[6] = e1x; o[7] = e1y;
                    o[8] = good; o[9] = dx0 * 4 + dy0 % 4;
                }
            }
            // second walk: clear the mask; un-vote (RED, no round trip) if the segment is accepted
            for (int k = 0; k < 2; k++) {
                const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
                const int nwords = (S.t_end[k] >> 5) + 1;
                for (int wd = 0; wd < nwords; wd++) {
                    u32 bits = S.walk_bits[k][wd];
                    if (k == 1 && wd == 0) bits &= ~1u;  // the start pixel was handled by direction 0
                    while (bits) {
                        int bpos = __ffs(bits) - 1;
                        bits &= bits - 1;
                        int t = wd * 32 + bpos;
                        int x = x0 + t * dx, y = y0 + t * dy;
                        int j1 = xflag ? x : (x >> shift), i1 = xflag ? (y >> shift) : y;
                        if (good && mine) {
                            int bb = __float2int_rn(__fadd_rn(__fmul_rn((float)j1, tc), __fmul_rn((float)i1, ts))) - rlo;
                            atomicAdd(myrow + (bb >> 1), 0u - (1u << ((bb & 1) * 16)));
                        }
                        if (tid == 0) mk[(size_t)i1 * w + j1] = 0;
                    }
                }
            }
            if (good && tid == 0) {
                int pos = S.nlines++;
                if (pos < max_segs) {
                    int32_t *o = segs + ((size_t)f * max_segs + pos) * 4;
                    o[0] = e0x; o[1] = e0y; o[2] = e1x; o[3] = e1y;
                }
            }
            __syncthreads();  // mask writes of thread 0 are visible to warp 0 below
            // refresh the alive bits of the not yet consumed candidates (the walk may have erased them)
            if (wid == 0) {
                bool al = false;
                if ((alive >> lane) & 1u) {
                    u32 pt = S.cand[lane];
                    al = mk[(size_t)(pt >> 16) * w + (pt & 0xffff)] != 0;
                }
                u32 bb = __ballot_sync(~0u, al);
                if (lane == 0) S.alive = bb;
            }
            __syncthreads();
            alive = S.alive;
            PF_TICK(3);
        }
    }
    __syncthreads();
    if (tid == 0) counts[f] = S.nlines;
#ifdef LE_PPHT_PROFILE
    if (trace && tid == 0 && f == 0 && trace_cap >= 2) {
        long long *o = (long long *)(trace + (size_t)(trace_cap - 2) * 10);
        for (int i = 0; i < 6; i++) o[i] = dbg[i];
    }
#endif
}

int le_ppht_fast_launch(le_ctx *ctx, int n, int h, int w, int numangle, int stride, int threshold, int min_len,
                        int max_gap, int32_t *segs, int32_t *counts, int max_segs, int32_t *trace, int trace_cap,
                        cudaStream_t st)
{
    // accumulator: u16 cells biased by 0x4040 (byte pattern 0x40), two per word
    size_t bytes = sizeof(u16) * (size_t)numangle * stride * n;
    LE_CUDA(ctx, cudaMemsetAsync(ctx->ppht_accum.p, 0x40, bytes, st));
    static le_dev_once attr;
    if (attr.needs(ctx->device, sizeof(PFShared))) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_ppht_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PFShared)));
        attr.done(ctx->device, sizeof(PFShared));
    }
    LE_T1(ctx, LE_K_PPHT_PREP, st);
    LE_T0(ctx, LE_K_PPHT, st);
    k_ppht_fast<<<n, PF_THREADS, sizeof(PFShared), st>>>((u8 *)ctx->ppht_mask.p, (u32 *)ctx->ppht_nz.p,
                                                        (u32 *)ctx->ppht_accum.p, (const float *)ctx->trig.p,
                                                        (int *)ctx->counters.p, segs, counts, max_segs, h, w, numangle,
                                                        stride, threshold, min_len, max_gap, trace, trace_cap);
    LE_T1(ctx, LE_K_PPHT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
