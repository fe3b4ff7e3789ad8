// le_houghp_rounds.cu -- exact PPHT (a6) in rounds of up to 64 candidate picks, for numangle <= 192.
//
// cv2's HoughLinesProbabilistic is sequential (every accepted segment erases pixels and votes that later picks
// see), but two facts make most of it parallel without changing a bit of the result:
//
//  (1) THE VISITING ORDER DOES NOT DEPEND ON THE IMAGE.  cv2 draws idx = rng() % count, takes nz[idx], moves the
//      last element into the hole and decrements count -- whether or not the pixel is still set.  The raw outputs
//      of cv::RNG(-1) are one fixed sequence (a table made once on the host); the swap-remove shuffle depends
//      only on the number of edge pixels.  k_ppht_order computes the whole order of a frame before any vote: a
//      warp does 32 steps of the shuffle at once (the list sits in shared memory when it fits) -- steps whose
//      reads could see a write of an earlier step of the same group are detected with match_any / a lane-bit OR
//      and start the next group instead.  The shuffle runs in place as swap(A[idx], A[last]): pick k ends up at
//      A[count - 1 - k], so no second list is needed.  (Round 1 replayed the RNG on thread 0 of the vote kernel,
//      two dependent global round trips per pick once the list outgrew 48 KB of shared memory: 4K frames.)
//
//  (2) VOTES COMMUTE; ONLY A TRIGGER CHANGES THE STATE OTHER PICKS SEE.  A round takes the next 64 candidates,
//      reads their mask bytes in one round trip, and every thread (= angle) issues the votes of all alive
//      candidates as atomicAdd WITH RETURN, in pick order: atomics of one thread to one address are performed in
//      program order, so the returned value + 1 is exactly what cv2's `++adata[r]` yields at that pick -- all
//      (up to 64) atomics of a thread are in flight together, one L2 round trip per round instead of one per 8
//      picks.  The first pick whose value reaches the threshold in any angle (warp REDUX.MIN + one shared
//      atomicMin) is the trigger: votes of later picks of the round are taken back (fire-and-forget RED), the
//      line is walked, and the next round starts behind the trigger.  Nothing else is ever rolled back.
//
// The walk reads the mask speculatively, 192 steps per direction per round trip (three warps per direction, two
// 32-step chunks each); the gap rule is then resolved from the ballot words in shared memory.  The second pass
// (erase, un-vote) is fire-and-forget.  Accumulator cells are u16 biased by 0x4040, two per 32-bit word (see
// le_houghp_fast.cu, which remains as the reference implementation behind LE_PPHT_ROUNDS=0 and for tracing).
#include "le_common.cuh"
#include <algorithm>
#include <vector>

#define PR_THREADS 192
#define PR_MAXH 4            // 32-candidate groups per round: template parameter NH <= PR_MAXH (LE_PPHT_B = 32 NH)
#define PR_WIN 512           // order entries staged in shared memory at a time
#define PR_MAXT_WORDS 512    // walk bitmap: 16384 steps per direction
#define PR_BIAS 0x4040
#define PR_CHUNKS 6          // 32-step chunks read per direction and trip (3 warps x 2)

// ------------------------------------------------------------------------------------------ cv::RNG(-1) table
// state = (u32)state * 4164903690 + (state >> 32), output (u32)state: the first `n` outputs, on the host.
int le_ppht_rng_table(le_ctx *ctx, size_t n, cudaStream_t st)
{
    if (ctx->ppht_rng_n >= n && ctx->ppht_rng.p) return LE_OK;
    n = (n + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    std::vector<u32> tab(n);
    u64 s = ~0ULL;
    for (size_t i = 0; i < n; i++) {
        s = (u64)(u32)s * 4164903690ULL + (s >> 32);
        tab[i] = (u32)s;
    }
    int rc = le_ensure(ctx, &ctx->ppht_rng, sizeof(u32) * n);
    if (rc) return rc;
    LE_CUDA(ctx, cudaMemcpyAsync(ctx->ppht_rng.p, tab.data(), sizeof(u32) * n, cudaMemcpyHostToDevice, st));
    LE_CUDA(ctx, cudaStreamSynchronize(st));  // `tab` goes out of scope
    ctx->ppht_rng_n = n;
    return LE_OK;
}

// ------------------------------------------------------------------------------------------ visiting order
// One warp per frame.  A = the frame's row-major edge list (shared copy when count <= cap, else in place in
// global memory).  Step k: c = count - k, idx = r[k] % c, swap(A[idx], A[c - 1]).  Lanes take steps k0 .. k0+31;
// lane l conflicts if an earlier lane of the group writes what it reads: idx_l' == idx_l or idx_l' == last_l
// (l' < l).  Lanes below the first conflicting lane are done in parallel (all reads, barrier, all writes), the
// next group starts at the conflicting lane.
#define PO_RNG 1024  // raw RNG outputs staged in shared memory at a time
__global__ void __launch_bounds__(32) k_ppht_order(u32 *__restrict__ nz_g, const int *__restrict__ counters,
                                                   const u32 *__restrict__ rng, size_t npx, int cap)
{
    extern __shared__ u32 s_list[];  // [cap] list, then [PO_RNG] staged RNG outputs
    u32 *s_rng = s_list + cap;
    const int f = blockIdx.x, lane = threadIdx.x;
    const int count = counters[f * LE_C_STRIDE + LE_C_AUX0];
    u32 *g = nz_g + (size_t)f * npx;
    const bool in_smem = count <= cap;
    u32 *A = in_smem ? s_list : g;
    if (in_smem) {
        for (int i = lane; i < count; i += 32) s_list[i] = g[i];
        __syncwarp();
    }
    int k0 = 0, rbase = -PO_RNG;
    while (k0 < count) {
        if (k0 + 32 > rbase + PO_RNG) {  // one coalesced round trip per ~1000 steps instead of one per group
            __syncwarp();
            rbase = k0;
            for (int i = lane; i < PO_RNG; i += 32) s_rng[i] = rbase + i < count ? __ldg(rng + rbase + i) : 0u;
            __syncwarp();
        }
        const int k = k0 + lane;
        const bool valid = k < count;
        const u32 c = valid ? (u32)(count - k) : 1u;
        const u32 idx = valid ? s_rng[k - rbase] % c : 0xffffff00u + lane;  // dummies never match
        const u32 last = c - 1;
        const u32 last0 = (u32)(count - k0) - 1;  // last of lane 0; last_l = last0 - l
        // (a) an earlier lane has the same idx
        const u32 same = __match_any_sync(~0u, idx);
        bool conf = (same & ((1u << lane) - 1)) != 0;
        // (b) an earlier lane l' writes the slot this lane reads as `last`: idx_l' == last0 - l  <=>  l == last0 - idx_l'
        const u32 t = last0 - idx;  // lane whose `last` this lane's idx is (if < 32 and above this lane)
        const u32 hit = (valid && t < 32u && t > (u32)lane) ? (1u << t) : 0u;
        const u32 hits = __reduce_or_sync(~0u, hit);
        conf = conf || ((hits >> lane) & 1u);
        const u32 cb = __ballot_sync(~0u, conf && valid);
        const int nok = cb ? __ffs(cb) - 1 : 32;  // lanes [0, nok) are conflict-free (nok >= 1: lane 0 never conflicts)
        const bool act = valid && lane < nok;
        u32 a = 0, b = 0;
        if (act) {
            a = A[idx];
            b = A[last];
        }
        __syncwarp();
        if (act) {
            A[idx] = b;
            A[last] = a;
        }
        __syncwarp();
        k0 += nok;
    }
    if (in_smem) {
        __syncwarp();
        for (int i = lane; i < count; i += 32) g[i] = s_list[i];
    }
}

// ------------------------------------------------------------------------------------------ rounds
struct PRShared {
    u32 win[PR_WIN];                       // order window: picks wbase .. wbase + PR_WIN - 1
    u32 alive[PR_MAXH];                    // alive bits of the round's candidates
    int kstar;                             // first triggering candidate of the round (PR_B = none)
    u32 best;                              // (val + 0x8000) << 8 | (255 - angle) of the trigger
    u32 walk_on[2][PR_MAXT_WORDS];         // raw "mask set" ballots, then the accepted bits, per direction
    u32 walk_inb[2][PR_CHUNKS];            // in-bounds ballots of the current trip
    int gap[2], t_end[2], stop[2];
    int end_xy[2][2];
    int nlines;
};

// barrier of the three warps of walk direction k (ids 1 and 2; immediate operands keep the other barriers free)
static __device__ __forceinline__ void pr_bar(int k)
{
    if (k == 0) asm volatile("bar.sync 1, 96;" ::: "memory");
    else asm volatile("bar.sync 2, 96;" ::: "memory");
}

template <int NH>
__global__ void __launch_bounds__(PR_THREADS)
k_ppht_rounds(u8 *__restrict__ mask, const u32 *__restrict__ nz_g, u32 *__restrict__ accum,
              const float *__restrict__ trig, const int *__restrict__ counters, int32_t *__restrict__ segs,
              int32_t *__restrict__ counts, int max_segs, int h, int w, int numangle, int stride, int threshold,
              int line_length, int line_gap, int round_cap, int flags)
{
    constexpr int PR_B = 32 * NH;
    __shared__ PRShared S;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t npx = (size_t)h * w;
    u8 *mk = mask + (size_t)f * npx;
    const int hw = stride >> 1;  // words per accumulator row
    const int count = counters[f * LE_C_STRIDE + LE_C_AUX0];
    const u32 *ord = nz_g + (size_t)f * npx;  // pick k = ord[count - 1 - k]
    const int *rlo_g = (const int *)(trig + 2 * numangle);
    const bool mine = tid < numangle;
    float tc = 0.f, ts = 0.f;
    int rlo = 0;
    if (mine) {
        tc = trig[tid];
        ts = trig[numangle + tid];
        rlo = rlo_g[tid];
    }
    u32 *myrow = accum + ((size_t)f * numangle + (mine ? tid : 0)) * hw;
    const int shift = 16;
    if (tid == 0) S.nlines = 0;
    int pos = 0;          // next pick
    int wbase = -PR_WIN;  // first pick of the staged window
    // candidates per round adapt to the trigger rate: halved after two rounds in a row that triggered (the votes
    // behind a trigger are issued and taken back), doubled after a round that did not.  Cube frames trigger in
    // ~40 % of the rounds and stay at the cap; cave-style frames trigger in ~90 % and settle at 8
    int cap_now = round_cap, streak = 0;
    // (A second group of 192 threads taking one direction of the erase / un-vote pass was also measured: no gain at
    // 128 x 4K, 256 x 1080p 5.5 -> 5.9 ms, 1024 x 1080p 26.9 -> 25.2 ms -- that pass is bound by the request rate too.)
    // (Fetching the NEXT round's mask bytes while the votes are in flight -- valid when the round does not trigger --
    // was measured and dropped: the extra scattered loads cost more than the round trip they hide, 256 x 1080p
    // 5.5 -> 6.9 ms, profiles/r2_ppht_variants.txt.)
#ifdef LE_PPHT_PROFILE  // phase cycle counters of frame 0 (variant build, scripts/ppht_profile.sh)
    long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cnt_[6] = {0, 0, 0, 0, 0, 0}, t_a = clock64();
#define PR_TICK(i) do { if (tid == 0) { long long t_ = clock64(); dbg[i] += t_ - t_a; t_a = t_; } } while (0)
#define PR_CNT(i, v) do { if (tid == 0) cnt_[i] += (v); } while (0)
#else
#define PR_TICK(i) do { } while (0)
#define PR_CNT(i, v) do { } while (0)
#endif
    auto cell_of = [&](u32 pt) -> int {
        const float x = (float)(pt & 0xffff), y = (float)(pt >> 16);
        return __float2int_rn(__fadd_rn(__fmul_rn(x, tc), __fmul_rn(y, ts))) - rlo;
    };

    while (pos < count) {
        // ---- candidates of the round and their mask bytes (one round trip)
        const int nb = min(cap_now, count - pos);
        if (pos + nb > wbase + PR_WIN || pos < wbase) {
            __syncthreads();  // readers of the previous window are done
            wbase = pos;
            for (int j = tid; j < PR_WIN; j += PR_THREADS)
                S.win[j] = wbase + j < count ? __ldcg(ord + (count - 1 - (wbase + j))) : 0u;
        }
        if (tid == 0) { S.kstar = PR_B; S.best = 0; }
        __syncthreads();
        const u32 *cand = S.win + (pos - wbase);
        if (wid < NH) {
            const int i = wid * 32 + lane;
            bool a = false;
            if (i < nb) {
                const u32 pt = cand[i];
                a = __ldcg(mk + (size_t)(pt >> 16) * w + (pt & 0xffff)) != 0;
            }
            const u32 b = __ballot_sync(~0u, a);
            if (lane == 0) S.alive[wid] = b;
        }
        __syncthreads();
        PR_TICK(0);
        PR_CNT(0, 1);
        u32 al[NH], any_alive = 0;
#pragma unroll
        for (int half = 0; half < NH; half++) {
            al[half] = S.alive[half];
            any_alive |= al[half];
        }
        if (any_alive == 0) {  // every candidate already erased
            pos += nb;
            continue;
        }
        // ---- votes of all alive candidates, in pick order, as atomics with return (all in flight together)
        int my_first = PR_B, my_val = 0;  // this angle's first candidate at or above the threshold, and its value
        if (mine) {
#pragma unroll
            for (int half = 0; half < NH; half++) {
                u32 am = al[half];
                while (am) {
                    u32 old[16];
                    int cb[16], ci[16];
                    int m = 0;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        cb[j] = 0;
                        ci[j] = PR_B;
                        old[j] = 0;
                        if (am) {
                            const int i = __ffs(am) - 1;
                            am &= am - 1;
                            const int b = cell_of(cand[half * 32 + i]);
                            cb[j] = b;
                            ci[j] = half * 32 + i;
                            old[j] = atomicAdd(myrow + (b >> 1), 1u << ((b & 1) * 16));
                            m = j + 1;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if (j < m) {
                            const int val = (int)((old[j] >> ((cb[j] & 1) * 16)) & 0xffff) - PR_BIAS + 1;
                            if (val >= threshold && ci[j] < my_first) {
                                my_first = ci[j];
                                my_val = val;
                            }
                        }
                    }
                }
            }
        }
        {
            const int wmin = __reduce_min_sync(~0u, my_first);
            if (lane == 0 && wmin < PR_B) atomicMin(&S.kstar, wmin);
        }
        __syncthreads();
        PR_TICK(1);
#ifdef LE_PPHT_PROFILE
        for (int half = 0; half < NH; half++) PR_CNT(1, __popc(al[half]));
#endif
        const int ks = S.kstar;
        if (ks == PR_B) {  // no trigger: every vote stands
            pos += nb;
            cap_now = min(round_cap, cap_now * 2);
            streak = 0;
            continue;
        }
        if ((flags & 1) && ++streak >= 2) cap_now = max(min(8, round_cap), cap_now >> 1);
        // ---- trigger at candidate ks: arg-max over the angles (highest value, then lowest angle)
        {
            const u32 key = (my_first == ks) ? (((u32)(my_val + 0x8000) << 8) | (u32)(255 - tid)) : 0u;
            const u32 wmax = __reduce_max_sync(~0u, key);
            if (lane == 0 && wmax) atomicMax(&S.best, wmax);
        }
        // take back the votes of the candidates behind the trigger (they are voted again next round)
        if (mine) {
#pragma unroll
            for (int half = 0; half < NH; half++) {
                u32 am = al[half];
                if (ks >= half * 32) am &= (ks - half * 32 >= 31) ? 0u : ~((2u << (ks - half * 32)) - 1);
                while (am) {
                    const int i = __ffs(am) - 1;
                    am &= am - 1;
                    const int b = cell_of(cand[half * 32 + i]);
                    atomicAdd(myrow + (b >> 1), 0u - (1u << ((b & 1) * 16)));
                }
            }
        }
        __syncthreads();
        PR_TICK(2);
        PR_CNT(2, 1);
        const u32 best = S.best;
        const int max_n = 255 - (int)(best & 255);
        const u32 ptk = cand[ks];
        pos += ks + 1;

        // ---- line walk from the trigger pixel (cv2's Q16 walk), both directions at once
        const int j = (int)(ptk & 0xffff), i = (int)(ptk >> 16);
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
        {
            const int k = wid / 3, wsub = wid % 3;  // direction, warp of the direction
            const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
            if (wsub == 0 && lane == 0) { S.gap[k] = 0; S.t_end[k] = 0; S.stop[k] = 0; S.end_xy[k][0] = j; S.end_xy[k][1] = i; }
            for (int trip = 0;; trip++) {
                // read: chunks trip * 6 + wsub * 2 + {0, 1}
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int chunk = trip * PR_CHUNKS + wsub * 2 + q;
                    const int t = chunk * 32 + lane;
                    const int x = x0 + t * dx, y = y0 + t * dy;
                    const int j1 = xflag ? x : (x >> shift), i1 = xflag ? (y >> shift) : y;
                    const bool inb = j1 >= 0 && j1 < w && i1 >= 0 && i1 < h && chunk < PR_MAXT_WORDS;
                    const bool on = inb && __ldcg(mk + (size_t)i1 * w + j1) != 0;
                    const u32 bm = __ballot_sync(~0u, on), ib = __ballot_sync(~0u, inb);
                    if (lane == 0 && chunk < PR_MAXT_WORDS) {
                        S.walk_on[k][chunk] = bm;
                        S.walk_inb[k][wsub * 2 + q] = ib;
                    }
                }
                pr_bar(k);
                // resolve the gap rule over the six chunks (one warp per direction)
                if (wsub == 0) {
                    int gap = S.gap[k], t_end = S.t_end[k];
                    bool stop = false;
                    for (int q = 0; q < PR_CHUNKS && !stop; q++) {
                        const int chunk = trip * PR_CHUNKS + q;
                        if (chunk >= PR_MAXT_WORDS) { stop = true; break; }
                        const u32 bm = S.walk_on[k][chunk], ib = S.walk_inb[k][q];
                        const bool on = (bm >> lane) & 1u, inb = (ib >> lane) & 1u;
                        const u32 below = bm & ((1u << lane) - 1);
                        const int run = below ? lane - (31 - __clz(below)) : gap + lane + 1;
                        const bool st = !inb || (!on && run > line_gap);
                        const u32 bs = __ballot_sync(~0u, st);
                        const u32 valid = bs ? ((1u << (__ffs(bs) - 1)) - 1) : ~0u;
                        const u32 bmv = bm & valid;
                        if (lane == 0) S.walk_on[k][chunk] = bmv;
                        if (bmv) {
                            const int hb = 31 - __clz(bmv);
                            t_end = chunk * 32 + hb;
                            gap = 31 - hb;
                        } else gap += 32;
                        stop = bs != 0;
                    }
                    if (lane == 0) {
                        S.gap[k] = gap;
                        S.t_end[k] = t_end;
                        S.stop[k] = stop;
                        if (stop) {
                            const int x = x0 + t_end * dx, y = y0 + t_end * dy;
                            S.end_xy[k][0] = xflag ? x : (x >> shift);
                            S.end_xy[k][1] = xflag ? (y >> shift) : y;
                        }
                    }
                }
                pr_bar(k);
                if (S.stop[k]) break;
            }
        }
        __syncthreads();
        PR_TICK(3);
        const int e0x = S.end_xy[0][0], e0y = S.end_xy[0][1], e1x = S.end_xy[1][0], e1y = S.end_xy[1][1];
        const bool good = abs(e1x - e0x) >= line_length || abs(e1y - e0y) >= line_length;
        // second walk: clear the mask; un-vote (RED, no round trip) if the segment is accepted
        // Consecutive pixels of the walk fall into the same or the neighbouring cell of a row (|step| <= 1.42 bins):
        // the decrements of one 32-bit word are summed in a register and leave as ONE RED -- the SM's rate of
        // scattered L2 requests, not L2 itself, bounds this kernel (a 300-pixel segment would issue 54 000 of them).
        int q = 0;
        int uw = -1;   // accumulator word of the pending decrement
        u32 ud = 0;    // pending decrement (both halves)
        for (int k = 0; k < 2; k++) {
            const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
            const int nwords = (S.t_end[k] >> 5) + 1;
            for (int wd = 0; wd < nwords; wd++) {
                u32 bits = S.walk_on[k][wd];
                if (wd == nwords - 1) bits &= (2u << (S.t_end[k] & 31)) - 1;  // chunks past the stop hold raw ballots
                if (k == 1 && wd == 0) bits &= ~1u;  // the start pixel was handled by direction 0
                // four pixels per trip: their cells are independent chains of ~10 dependent instructions each (one
                // warp per scheduler has nothing else to issue), only the run-length combining is sequential
                while (bits) {
                    int tt[4], bbv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        tt[u] = bits ? wd * 32 + __ffs(bits) - 1 : -1;
                        bits &= bits - 1;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int t = max(tt[u], 0);
                        const int x = x0 + t * dx, y = y0 + t * dy;
                        const int j1 = xflag ? x : (x >> shift), i1 = xflag ? (y >> shift) : y;
                        bbv[u] = __float2int_rn(__fadd_rn(__fmul_rn((float)j1, tc), __fmul_rn((float)i1, ts))) - rlo;
                        if (tt[u] >= 0 && q == tid) __stcg(mk + (size_t)i1 * w + j1, (u8)0);
                        if (tt[u] >= 0) q = q + 1 == PR_THREADS ? 0 : q + 1;
                    }
                    if (good && mine) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (tt[u] < 0) continue;
                            if ((bbv[u] >> 1) != uw) {
                                if (ud) atomicAdd(myrow + uw, 0u - ud);
                                uw = bbv[u] >> 1;
                                ud = 0;
                            }
                            ud += 1u << ((bbv[u] & 1) * 16);
                        }
                    }
                }
            }
        }
        if (ud) atomicAdd(myrow + uw, 0u - ud);
        if (good && tid == 0) {
            const int p = S.nlines++;
            if (p < max_segs) {
                int32_t *o = segs + ((size_t)f * max_segs + p) * 4;
                o[0] = e0x; o[1] = e0y; o[2] = e1x; o[3] = e1y;
            }
        }
        __syncthreads();  // mask stores are ordered before the next round's mask loads (same CTA, L2 accesses)
        PR_TICK(4);
        PR_CNT(3, good ? 1 : 0);
        PR_CNT(4, S.t_end[0] + S.t_end[1]);
    }
    __syncthreads();
    if (tid == 0) counts[f] = S.nlines;
#ifdef LE_PPHT_PROFILE
    if (tid == 0 && f == 0)
        printf("PPHTPROF count %d rounds %lld alive-voted %lld triggers %lld good %lld walk-steps %lld | cycles: cand+mask %lld "
               "votes %lld trigger+rollback %lld walk %lld erase+unvote %lld\n", count, cnt_[0], cnt_[1], cnt_[2], cnt_[3],
               cnt_[4], dbg[0], dbg[1], dbg[2], dbg[3], dbg[4]);
#endif
}

int le_ppht_rounds_launch(le_ctx *ctx, int n, int h, int w, int numangle, int stride, int threshold, int min_len,
                          int max_gap, int32_t *segs, int32_t *counts, int max_segs, cudaStream_t st)
{
    const size_t npx = (size_t)h * w;
    int rc = le_ppht_rng_table(ctx, npx, st);
    if (rc) return rc;
    // accumulator: u16 cells biased by 0x4040 (byte pattern 0x40), two per word
    LE_CUDA(ctx, cudaMemsetAsync(ctx->ppht_accum.p, 0x40, sizeof(u16) * (size_t)numangle * stride * n, st));
    // order: the list in shared memory when it fits (most of the SM's 227 KB; one warp per frame, one CTA per SM
    // is plenty -- a 4K cube frame takes ~0.1 ms), else in place in global memory
    // list capacity in shared memory: 1 / 96 of the frame (cube and voxel frames have 0.2 - 0.6 % edge pixels,
    // cave frames ~1 %), at least 12 K entries, at most what one SM offers -- 86 KB and two CTAs per SM at 1080p,
    // one CTA per SM at 4K; a longer list is shuffled in place in global memory
    const int cap_max = (ctx->smem_optin - 1024) / (int)sizeof(u32) - PO_RNG;
    const int cap = (int)std::min((size_t)cap_max, std::max((size_t)12288, npx / 96));
    const size_t osmem = ((size_t)cap + PO_RNG) * sizeof(u32);
    static le_dev_once attr;
    if (attr.needs(ctx->device, osmem)) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_ppht_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osmem));
        attr.done(ctx->device, osmem);
    }
    k_ppht_order<<<n, 32, osmem, st>>>((u32 *)ctx->ppht_nz.p, (const int *)ctx->counters.p,
                                       (const u32 *)ctx->ppht_rng.p, npx, cap);
    LE_LAUNCH_CHECK(ctx);
    LE_T1(ctx, LE_K_PPHT_PREP, st);
    LE_T0(ctx, LE_K_PPHT, st);
    // candidates per round: measured on 128 x 4K cube frames 32 -> 8.1 ms, 64 -> 9.2, 128 -> 13.1 (votes behind a
    // trigger are issued and taken back: the SM's scattered-request rate pays for them twice)
    // With more frames than SMs the CTAs of an SM share that rate and wasted votes cost more than round trips:
    // 256 x 1080p 16 -> 6.0 ms (32: 6.7), 1024 x 1080p 8 -> 35.8 ms (16: 38.9, 32: 49.4)
    const int round_dflt = n <= ctx->sm_count ? 32 : (n <= 2 * ctx->sm_count ? 16 : 8);
    const int knob = LE_KNOB_INT("LE_PPHT_B", 0);  // development knob
    const int round_cap = std::max(1, std::min(128, knob ? knob : round_dflt));
    const int nh = (round_cap + 31) / 32;
    // development knob: LE_PPHT_NOADAPT=1 keeps the round size fixed (128 cave frames: 22.7 -> 34.5 ms)
    const int flags = LE_KNOB_SET("LE_PPHT_NOADAPT") ? 0 : 1;
#define LE_ROUNDS(NH)                                                                                             \
    k_ppht_rounds<NH><<<n, PR_THREADS, 0, st>>>((u8 *)ctx->ppht_mask.p, (const u32 *)ctx->ppht_nz.p,              \
                                                (u32 *)ctx->ppht_accum.p, (const float *)ctx->trig.p,             \
                                                (const int *)ctx->counters.p, segs, counts, max_segs, h, w,       \
                                                numangle, stride, threshold, min_len, max_gap, round_cap, flags)
    if (nh == 1) LE_ROUNDS(1);
    else if (nh == 2) LE_ROUNDS(2);
    else LE_ROUNDS(4);
#undef LE_ROUNDS
    LE_T1(ctx, LE_K_PPHT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
