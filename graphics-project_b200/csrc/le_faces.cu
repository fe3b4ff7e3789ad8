// le_faces.cu -- get_faces_from_pairs: the quad search that follows classifyEdges / smoothEdges
// in the reference's `lines` and `mixed` pipelines (opencv_renewed.py:62-112; callers :176-178, :268-270,
// :278-280, :310-312), with util.py segments_distance :198-209, get_segments_intersection :212-231,
// point_segment_distance :242-267, lineIntersection :54-68, pointInConvexPolygon :162-168 and
// faceCircumference :170-171.
//
// The reference nests four Python loops (i over edges1, j over edges2, k > i, l > j) and evaluates
// segments_distance up to four times in the innermost one.  All four tests read the SAME n1 x n2 relation
// near[i][j] = segments_distance(edges1[i], edges2[j]) <= threshold (the function is symmetric in its two
// segments), so one CTA per problem
//   (1) builds near[][] once, in parallel, as a bit matrix;
//   (2) counts, for every set (i, j), the quads it opens:  sum over k > i with near[k][j] of
//       popcount(row_i & row_k & bits above j);  an exclusive scan over (i, j) in row-major order then gives
//       every quad its position in the reference's (i, j, k, l) lexicographic order;
//   (3) writes the quads and their four corner intersections at those positions;
//   (4) applies the reference's duplicate rule, which (its ban list never matches) reduces to an
//       independent test per face: dropped iff a LATER face shares its (i, k) or (j, l) pair, overlaps it and
//       has a smaller circumference;
//   (5) orients the kept faces with the reference's own (quirky) sign test.
// float64 throughout in the reference's operation order, compiled without FMA contraction (-fmad=false).
#include "le_common.cuh"

#define FC_THREADS 256
#define FC_MAX 1024  // segments per list

struct P2 {
    double x, y;
};

// util.py:242-267
static __device__ double pt_seg_distance(P2 p, P2 a1, P2 a2)
{
    double dx = a2.x - a1.x, dy = a2.y - a1.y;
    if (dx == 0 && dy == 0) return hypot(p.x - a1.x, p.y - a1.y);
    const double t = ((p.x - a1.x) * dx + (p.y - a1.y) * dy) / (dx * dx + dy * dy);
    if (t < 0) { dx = p.x - a1.x; dy = p.y - a1.y; }
    else if (t > 1) { dx = p.x - a2.x; dy = p.y - a2.y; }
    else { dx = p.x - (a1.x + t * dx); dy = p.y - (a1.y + t * dy); }
    return hypot(dx, dy);
}

// util.py:198-239
static __device__ bool segments_near(P2 a1, P2 a2, P2 b1, P2 b2, double threshold)
{
    const double dx1 = a2.x - a1.x, dy1 = a2.y - a1.y, dx2 = b2.x - b1.x, dy2 = b2.y - b1.y;
    const double delta = dx2 * dy1 - dy2 * dx1;
    if (delta != 0) {
        const double s = (dx1 * (b1.y - a1.y) + dy1 * (a1.x - b1.x)) / delta;
        const double t = (dx2 * (a1.y - b1.y) + dy2 * (b1.x - a1.x)) / (-delta);
        if (0 <= s && s <= 1 && 0 <= t && t <= 1) return 0 <= threshold;  // distance 0
    }
    double m = pt_seg_distance(a1, b1, b2);
    m = fmin(m, pt_seg_distance(a2, b1, b2));
    m = fmin(m, pt_seg_distance(b1, a1, a2));
    m = fmin(m, pt_seg_distance(b2, a1, a2));
    return m <= threshold;
}

// util.py:54-68 (None -> NaN)
static __device__ P2 line_intersection(P2 a1, P2 a2, P2 b1, P2 b2)
{
    const double xd0 = a1.x - a2.x, xd1 = b1.x - b2.x, yd0 = a1.y - a2.y, yd1 = b1.y - b2.y;
    const double div = xd0 * yd1 - xd1 * yd0;
    if (div == 0) return P2{nan(""), nan("")};
    const double d0 = a1.x * a2.y - a1.y * a2.x, d1 = b1.x * b2.y - b1.y * b2.x;
    return P2{(d0 * xd1 - d1 * xd0) / div, (d0 * yd1 - d1 * yd0) / div};
}

static __device__ double side(P2 p, P2 e0, P2 e1) { return (p.x - e0.x) * (e1.y - e0.y) - (p.y - e0.y) * (e1.x - e0.x); }

// util.py:162-168: the closing edge (poly[3], poly[0]) is never tested
static __device__ bool in_convex(P2 p, const P2 *poly)
{
    const double sign = side(p, poly[0], poly[1]);
    if (sign == 0) return true;
    return 0 <= sign * side(p, poly[1], poly[2]) && 0 <= sign * side(p, poly[2], poly[3]);
}

static __device__ P2 seg_a(const double *s) { return P2{s[0], s[1]}; }
static __device__ P2 seg_b(const double *s) { return P2{s[2], s[3]}; }

__global__ void __launch_bounds__(FC_THREADS)
k_face_quads(const double *__restrict__ segs1, const int32_t *__restrict__ counts1, int max1,
             const double *__restrict__ segs2, const int32_t *__restrict__ counts2, int max2, double threshold,
             int32_t *__restrict__ quads, double *__restrict__ faces, uint8_t *__restrict__ keep,
             int32_t *__restrict__ counts, int max_out, u32 *__restrict__ bits_all, int *__restrict__ cnt_all,
             double *__restrict__ aux_all)
{
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n1 = counts1 ? min(max(counts1[f], 0), max1) : max1, n2 = counts2 ? min(max(counts2[f], 0), max2) : max2;
    const int nw = (max2 + 31) / 32;
    const double *s1 = segs1 + (size_t)f * max1 * 4, *s2 = segs2 + (size_t)f * max2 * 4;
    u32 *bits = bits_all + (size_t)f * max1 * nw;
    int *cnt = cnt_all + (size_t)f * max1 * max2;
    double *aux = aux_all + (size_t)f * max_out * 3;  // centroid x, y, circumference
    int32_t *q = quads + (size_t)f * max_out * 4;
    double *fc = faces + (size_t)f * max_out * 8;
    uint8_t *kp = keep + (size_t)f * max_out;
    __shared__ int s_part[FC_THREADS];
    __shared__ int s_total;
    const int np = n1 * n2;

    for (int i = tid; i < n1 * nw; i += FC_THREADS) bits[i] = 0;
    __syncthreads();
    // (1) the relation
    for (int p = tid; p < np; p += FC_THREADS) {
        const int i = p / n2, j = p - i * n2;
        if (segments_near(seg_a(s1 + 4 * i), seg_b(s1 + 4 * i), seg_a(s2 + 4 * j), seg_b(s2 + 4 * j), threshold))
            atomicOr(&bits[i * nw + (j >> 5)], 1u << (j & 31));
    }
    __syncthreads();
    // (2) quads opened by every (i, j)
    const int w2 = (n2 + 31) / 32;
    for (int p = tid; p < np; p += FC_THREADS) {
        const int i = p / n2, j = p - i * n2;
        int c = 0;
        if ((bits[i * nw + (j >> 5)] >> (j & 31)) & 1) {
            for (int k = i + 1; k < n1; k++) {
                if (!((bits[k * nw + (j >> 5)] >> (j & 31)) & 1)) continue;
                for (int w = j >> 5; w < w2; w++) {
                    u32 m = bits[i * nw + w] & bits[k * nw + w];
                    if (w == (j >> 5)) m &= (j & 31) == 31 ? 0u : ~0u << ((j & 31) + 1);
                    c += __popc(m);
                }
            }
        }
        cnt[p] = c;
    }
    __syncthreads();
    // exclusive scan of cnt over p (row-major (i, j) = the reference's outer loop order)
    {
        const int chunk = (np + FC_THREADS - 1) / FC_THREADS;
        const int lo = min(tid * chunk, np), hi = min(lo + chunk, np);
        int s = 0;
        for (int p = lo; p < hi; p++) s += cnt[p];
        s_part[tid] = s;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int t = 0; t < FC_THREADS; t++) {
                const int v = s_part[t];
                s_part[t] = acc;
                acc += v;
            }
            s_total = acc;
            counts[f] = acc;
        }
        __syncthreads();
        int acc = s_part[tid];
        for (int p = lo; p < hi; p++) {
            const int v = cnt[p];
            cnt[p] = acc;
            acc += v;
        }
    }
    __syncthreads();
    // (3) quads and corners in order
    for (int p = tid; p < np; p += FC_THREADS) {
        const int i = p / n2, j = p - i * n2;
        if (!((bits[i * nw + (j >> 5)] >> (j & 31)) & 1)) continue;
        int o = cnt[p];
        if (o >= max_out) continue;
        for (int k = i + 1; k < n1 && o < max_out; k++) {
            if (!((bits[k * nw + (j >> 5)] >> (j & 31)) & 1)) continue;
            for (int w = j >> 5; w < w2 && o < max_out; w++) {
                u32 m = bits[i * nw + w] & bits[k * nw + w];
                if (w == (j >> 5)) m &= (j & 31) == 31 ? 0u : ~0u << ((j & 31) + 1);
                while (m && o < max_out) {
                    const int l = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    q[4 * o] = i; q[4 * o + 1] = j; q[4 * o + 2] = k; q[4 * o + 3] = l;
                    const P2 e1a = seg_a(s1 + 4 * i), e1b = seg_b(s1 + 4 * i), e2a = seg_a(s2 + 4 * j), e2b = seg_b(s2 + 4 * j);
                    const P2 e3a = seg_a(s1 + 4 * k), e3b = seg_b(s1 + 4 * k), e4a = seg_a(s2 + 4 * l), e4b = seg_b(s2 + 4 * l);
                    P2 c[4] = {line_intersection(e1a, e1b, e2a, e2b), line_intersection(e2a, e2b, e3a, e3b),
                               line_intersection(e3a, e3b, e4a, e4b), line_intersection(e4a, e4b, e1a, e1b)};
                    double sx = 0, sy = 0, circ = 0;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        fc[8 * o + 2 * v] = c[v].x;
                        fc[8 * o + 2 * v + 1] = c[v].y;
                        sx = sx + c[v].x;
                        sy = sy + c[v].y;
                    }
                    // util.py:170-171: |f3 - f0| + |f0 - f1| + |f1 - f2| + |f2 - f3|, summed left to right
#pragma unroll
                    for (int v = -1; v < 3; v++) {
                        const P2 a = c[(v + 4) & 3], b = c[v + 1];
                        const double dx = a.x - b.x, dy = a.y - b.y;
                        circ = circ + sqrt(dx * dx + dy * dy);
                    }
                    aux[3 * o] = sx / 4;
                    aux[3 * o + 1] = sy / 4;
                    aux[3 * o + 2] = circ;
                    o++;
                }
            }
        }
    }
    __syncthreads();
    // (4) duplicate rule
    const int nf = min(s_total, max_out);
    for (int a = tid; a < nf; a += FC_THREADS) {
        const int ai = q[4 * a], aj = q[4 * a + 1], ak = q[4 * a + 2], al = q[4 * a + 3];
        P2 fa[4];
        for (int v = 0; v < 4; v++) fa[v] = P2{fc[8 * a + 2 * v], fc[8 * a + 2 * v + 1]};
        const P2 ca = P2{aux[3 * a], aux[3 * a + 1]};
        const double circ_a = aux[3 * a + 2];
        bool ok = true;
        for (int b = a + 1; b < nf; b++) {
            const bool share = (q[4 * b] == ai && q[4 * b + 2] == ak) || (q[4 * b + 1] == aj && q[4 * b + 3] == al);
            if (!share) continue;
            P2 fb[4];
            for (int v = 0; v < 4; v++) fb[v] = P2{fc[8 * b + 2 * v], fc[8 * b + 2 * v + 1]};
            if (in_convex(ca, fb) || in_convex(P2{aux[3 * b], aux[3 * b + 1]}, fa)) {
                if (circ_a > aux[3 * b + 2]) { ok = false; break; }
            }
        }
        kp[a] = ok ? 1 : 0;
    }
    __syncthreads();
    // (5) orientation (opencv_renewed.py:103-111): keep if e1.x * e2.x - e1.y * e2.x < 0, else reverse
    for (int a = tid; a < nf; a += FC_THREADS) {
        double *c = fc + 8 * a;
        const double e1x = c[2] - c[0], e1y = c[3] - c[1], e2x = c[4] - c[2];
        if (!(e1x * e2x - e1y * e2x < 0)) {
            double t;
            t = c[0]; c[0] = c[6]; c[6] = t;
            t = c[1]; c[1] = c[7]; c[7] = t;
            t = c[2]; c[2] = c[4]; c[4] = t;
            t = c[3]; c[3] = c[5]; c[5] = t;
        }
    }
}

extern "C" int le_face_quads(le_ctx *ctx, const double *segs1, const int32_t *counts1, int max1, const double *segs2,
                             const int32_t *counts2, int max2, int n, double threshold, int32_t *quads, double *faces,
                             uint8_t *keep, int32_t *counts, int max_out, void *stream)
{
    LE_ON_DEVICE(ctx);
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, segs1 && segs2 && quads && faces && keep && counts, "null pointer");
    LE_REQUIRE(ctx, n > 0 && max1 > 0 && max2 > 0 && max_out > 0, "n, max1, max2, max_out must be positive");
    if (max1 > FC_MAX || max2 > FC_MAX)
        return le_fail(ctx, LE_ERR_UNSUPPORTED, "at most %d segments per list (got %d, %d)", FC_MAX, max1, max2);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nw = (max2 + 31) / 32;
    const size_t b_bits = (sizeof(u32) * (size_t)n * max1 * nw + 255) & ~(size_t)255;
    const size_t b_cnt = (sizeof(int) * (size_t)n * max1 * max2 + 255) & ~(size_t)255;
    const size_t b_aux = sizeof(double) * (size_t)n * max_out * 3;
    int rc = le_ensure(ctx, &ctx->misc, b_bits + b_cnt + b_aux);
    if (rc) return rc;
    char *base = (char *)ctx->misc.p;
    LE_T0(ctx, LE_K_MISC, st);
    k_face_quads<<<n, FC_THREADS, 0, st>>>(segs1, counts1, max1, segs2, counts2, max2, threshold, quads, faces, keep,
                                           counts, max_out, (u32 *)base, (int *)(base + b_bits),
                                           (double *)(base + b_bits + b_cnt));
    LE_T1(ctx, LE_K_MISC, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
