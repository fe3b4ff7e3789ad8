// le_merge.cu -- combineParallelLines: the greedy merge of nearly parallel, nearby segments
// that opens every pipeline of the reference (util.py:141-159; callers opencv.py:225, graph.py:116,
// opencv_renewed.py:56-60).
//
// The reference rescans all pairs from (0, 1) after every merge (a recursive call per merge, O(n^3)
// Python: 150-590 ms on demo_scarce.png, minutes on the Minecraft fixtures).  Its result is "repeat: take
// the lexicographically first mergeable pair (i < j), replace i by the combined segment, delete j".
// Here one CTA owns a frame and keeps the mergeable-pair relation as a bit matrix (row a = bits b > a):
// built once in parallel, then per merge only row i and column i are re-evaluated; deletions are a mask
// (alive bits), the first pair is a parallel find-first-set + block minimum.  All arithmetic is float64 in
// the reference's operation order, compiled without FMA contraction (Makefile: -fmad=false).
#include "le_common.cuh"

#define MG_THREADS 1024
#define MG_MAX 4096  // segments per frame (bit matrix of n x n/32 words in global scratch)

struct Seg {
    double x1, y1, x2, y2;
};

// util.py:27-52 getEdgeProjection, then the distance of the point to it (util.py:106-107)
static __device__ double pt_dist(double px, double py, double ax, double ay, double bx, double by)
{
    const double ex = bx - ax, ey = by - ay;
    const double els = ex * ex + ey * ey;
    double qx, qy;
    if (els <= 0.001) { qx = bx; qy = by; }
    else {
        const double t = (ex * (px - ax) + ey * (py - ay)) / els;
        if (t < 0) { qx = ax; qy = ay; }
        else if (t > 1) { qx = bx; qy = by; }
        else { qx = t * ex + ax; qy = t * ey + ay; }
    }
    const double dx = px - qx, dy = py - qy;
    return sqrt(dx * dx + dy * dy);
}

// util.py:145-146: |cos| of the angle above cos(max_angle) and edgeDistance below max_distance
static __device__ bool mergeable(const Seg &a, const Seg &b, double max_distance, double cos_angle)
{
    const double eax = a.x2 - a.x1, eay = a.y2 - a.y1, ebx = b.x2 - b.x1, eby = b.y2 - b.y1;
    const double d = fabs(eax * ebx + eay * eby);
    const double na = sqrt(eax * eax + eay * eay), nb = sqrt(ebx * ebx + eby * eby);
    if (!(d / (na * nb) > cos_angle)) return false;
    double m = pt_dist(a.x1, a.y1, b.x1, b.y1, b.x2, b.y2);
    m = fmin(m, pt_dist(a.x2, a.y2, b.x1, b.y1, b.x2, b.y2));
    m = fmin(m, pt_dist(b.x1, b.y1, a.x1, a.y1, a.x2, a.y2));
    m = fmin(m, pt_dist(b.x2, b.y2, a.x1, a.y1, a.x2, a.y2));
    return m < max_distance;
}

// util.py:115-139 combineEdges
static __device__ Seg combine(const Seg &a, const Seg &b)
{
    double p3x = b.x1, p3y = b.y1, p4x = b.x2, p4y = b.y2;
    const double e1x = a.x2 - a.x1, e1y = a.y2 - a.y1;
    double e2x = p4x - p3x, e2y = p4y - p3y;
    if (e1x * e2x + e1y * e2y < 0) {
        e2x = -e2x; e2y = -e2y;
        double tx = p3x, ty = p3y;
        p3x = p4x; p3y = p4y; p4x = tx; p4y = ty;
    }
    const double ax = e1x + e2x, ay = e1y + e2y;
    double ox = a.x1, oy = a.y1;
    if ((p3x - ox) * ax + (p3y - oy) * ay < 0) { ox = p3x; oy = p3y; }
    const double den = ax * ax + ay * ay;
    const double l2 = ((a.x2 - ox) * ax + (a.y2 - oy) * ay) / den, l4 = ((p4x - ox) * ax + (p4y - oy) * ay) / den;
    const double len = l2 > l4 ? l2 : (l4 > l2 ? l4 : l2);  // Python max(): first of equals, NaN never wins
    Seg r;
    r.x1 = ox; r.y1 = oy; r.x2 = ox + len * ax; r.y2 = oy + len * ay;
    return r;
}

// One CTA per frame.  segs_out doubles as the working list.  pmat: [frames][max_k][nw] words.
__global__ void __launch_bounds__(MG_THREADS)
k_combine_parallel(const double *__restrict__ segs_in, const int *__restrict__ counts_in, int max_k, double md0,
                   double ca0, double md1, double ca1, double *__restrict__ segs_out, int *__restrict__ src_out,
                   int *__restrict__ counts_out, u32 *__restrict__ pmat, int nw)
{
    __shared__ u32 s_alive[MG_MAX / 32];
    __shared__ u32 s_mergedf[MG_MAX / 32];
    __shared__ u32 s_key;
    __shared__ int s_scan[MG_THREADS / 32];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = min(counts_in ? counts_in[f] : max_k, max_k);
    Seg *S = (Seg *)(segs_out + (size_t)f * max_k * 4);
    const Seg *Sin = (const Seg *)(segs_in + (size_t)f * max_k * 4);
    u32 *P = pmat + (size_t)f * max_k * nw;
    for (int i = tid; i < n; i += MG_THREADS) S[i] = Sin[i];
    for (int i = tid; i < MG_MAX / 32; i += MG_THREADS) {
        int lo = i * 32;
        s_alive[i] = lo + 32 <= n ? ~0u : (lo < n ? ((1u << (n - lo)) - 1u) : 0u);
        s_mergedf[i] = 0;
    }
    __syncthreads();
    const int nwn = (n + 31) >> 5;  // words in use

    // (re)build rows [r0, r1) of the relation for columns b > row, one warp per 32 columns
    auto build_rows = [&](int r0, int r1, double md, double ca) {
        const int warps = MG_THREADS / 32, wid = tid >> 5;
        for (long long t = wid; t < (long long)(r1 - r0) * nwn; t += warps) {
            const int a = r0 + (int)(t / nwn), wi = (int)(t % nwn);
            if (wi < (a >> 5)) continue;  // columns <= a only
            const int b = wi * 32 + lane;
            bool pr = false;
            if (b > a && b < n && ((s_alive[wi] >> lane) & 1u) && ((s_alive[a >> 5] >> (a & 31)) & 1u))
                pr = mergeable(S[a], S[b], md, ca);
            const u32 word = __ballot_sync(~0u, pr);
            if (lane == 0) P[(size_t)a * nw + wi] = word;
        }
    };
    build_rows(0, n, md0, ca0);
    __syncthreads();

    bool first = true;
    for (;;) {
        // lexicographically first alive pair
        if (tid == 0) s_key = 0xffffffffu;
        __syncthreads();
        u32 key = 0xffffffffu;
        for (int a = tid; a < n && key == 0xffffffffu; a += MG_THREADS) {
            if (!((s_alive[a >> 5] >> (a & 31)) & 1u)) continue;
            for (int wi = a >> 5; wi < nwn; wi++) {
                u32 wv = P[(size_t)a * nw + wi] & s_alive[wi];
                if (wi == (a >> 5)) wv &= (a & 31) == 31 ? 0u : ~0u << ((a & 31) + 1);
                if (wv) { key = (u32)a * MG_MAX + (u32)(wi * 32 + __ffs(wv) - 1); break; }
            }
        }
        // rows are visited in increasing a per thread, so a thread's first hit is its smallest key
        if (key != 0xffffffffu) atomicMin(&s_key, key);
        __syncthreads();
        const u32 best = s_key;
        if (best == 0xffffffffu) break;
        const int i = (int)(best / MG_MAX), j = (int)(best % MG_MAX);
        __syncthreads();
        if (tid == 0) {
            S[i] = combine(S[i], S[j]);
            s_alive[j >> 5] &= ~(1u << (j & 31));
            s_mergedf[i >> 5] |= 1u << (i & 31);
        }
        __syncthreads();
        if (first && (md0 != md1 || ca0 != ca1)) {
            // the reference recurses with its DEFAULT thresholds: every pair is judged again
            build_rows(0, n, md1, ca1);
        } else {
            build_rows(i, i + 1, md1, ca1);  // row i: pairs (i, b > i)
            // column i: pairs (a < i, i)
            for (int a = tid; a < i; a += MG_THREADS) {
                if (!((s_alive[a >> 5] >> (a & 31)) & 1u)) continue;
                const bool pr = mergeable(S[a], S[i], md1, ca1);
                u32 *wp = &P[(size_t)a * nw + (i >> 5)];
                if (pr) atomicOr(wp, 1u << (i & 31));
                else atomicAnd(wp, ~(1u << (i & 31)));
            }
        }
        first = false;
        __syncthreads();
    }

    // ordered compaction of the surviving slots (each thread holds its elements while the list is rewritten)
    constexpr int PER = MG_MAX / MG_THREADS;
    Seg mine[PER];
    int alive_me[PER], cnt = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int idx = tid * PER + k;
        alive_me[k] = idx < n && ((s_alive[idx >> 5] >> (idx & 31)) & 1u);
        if (alive_me[k]) mine[k] = S[idx];
        cnt += alive_me[k];
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(~0u, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int v = tid < MG_THREADS / 32 ? s_scan[tid] : 0, inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(~0u, inc2, o);
            if (lane >= o) inc2 += u;
        }
        if (tid < MG_THREADS / 32) s_scan[tid] = inc2 - v;
        if (tid == 31) counts_out[f] = inc2;
    }
    __syncthreads();
    int pos = s_scan[tid >> 5] + incl - cnt;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        if (alive_me[k]) {
            const int idx = tid * PER + k;
            S[pos] = mine[k];
            if (src_out) src_out[(size_t)f * max_k + pos] = idx | (((s_mergedf[idx >> 5] >> (idx & 31)) & 1u) ? 0x40000000 : 0);
            pos++;
        }
    }
}

extern "C" int le_combine_parallel_lines(le_ctx *ctx, const double *segs, const int32_t *counts, int n, int max_k,
                                         double max_distance, double cos_max_angle, double default_distance,
                                         double default_cos, double *out_segs, int32_t *out_src, int32_t *out_counts,
                                         void *stream)
{
    LE_ON_DEVICE(ctx);
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, segs && out_segs && out_counts, "null pointer");
    LE_REQUIRE(ctx, n > 0 && max_k > 0, "n and max_k must be positive");
    LE_REQUIRE(ctx, segs != out_segs, "le_combine_parallel_lines is not in-place");
    if (max_k > MG_MAX) return le_fail(ctx, LE_ERR_UNSUPPORTED, "at most %d segments per frame (got %d)", MG_MAX, max_k);
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = (max_k + 31) / 32;
    int rc = le_ensure(ctx, &ctx->misc, sizeof(u32) * (size_t)n * max_k * nw);
    if (rc) return rc;
    LE_T0(ctx, LE_K_MISC, st);
    k_combine_parallel<<<n, MG_THREADS, 0, st>>>(segs, counts, max_k, max_distance, cos_max_angle, default_distance,
                                                 default_cos, out_segs, out_src, out_counts, (u32 *)ctx->misc.p, nw);
    LE_T1(ctx, LE_K_MISC, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
