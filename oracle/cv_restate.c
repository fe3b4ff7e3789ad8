/*
 * oracle/cv_restate.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement, in plain C, of the OpenCV arithmetic that the reference
 * (pklito/Graphics-Project) calls on its line-detection hot path.  The reference
 * itself is pure Python; the arithmetic lives in the third-party wheel
 * `opencv-python==4.10.0.84` (reference requirements.txt:2), whose source is NOT
 * under /root/reference.  The algorithms below restate the published OpenCV
 * algorithms (imgproc: color, smooth, canny, hough, lsd) and are PINNED against
 * the cv2 4.13.0 wheel that ships in this image (tests/test_oracle.py::test_oracle_vs_live_cv2)
 * and against golden vectors generated from the reference's own fixture images
 * (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Reference call sites each function stands for:
 *   or_bgr2gray      cv.cvtColor(BGR2GRAY)   opencv.py:22,205,213  opencv_fit_color.py:74
 *   or_gauss5        cv.GaussianBlur((5,5),0) opencv.py:25          opencv_fit_color.py:77
 *   or_normalize     cv.normalize(MINMAX)    opencv.py:26
 *   or_canny         cv.Canny                opencv.py:26,214      opencv_fit_color.py:78
 *   or_hough_*       cv.HoughLines           opencv.py:88,118      opencv_fit_color.py:83
 *   or_hough_lines_p cv.HoughLinesP          opencv.py:35,217
 *   or_lsd           cv.createLineSegmentDetector(...).detect  opencv.py:206-207
 *   or_f32_*         the same cvtColor / GaussianBlur / normalize on float32 frames
 *                    _fboToImage opencv.py:101-108 -> doCanny opencv.py:21-28
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: cv2's scalar paths are unfused f32/f64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef unsigned char u8;

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}
static inline int clampi(int p, int lo, int hi) { return p < lo ? lo : (p > hi ? hi : p); }

/* ---------------------------------------------------------------- a1: gray */
void or_bgr2gray(const u8 *bgr, u8 *gray, int h, int w)
{
    for (long i = 0; i < (long)h * w; i++) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (u8)((3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15);
    }
}

/* ------------------------------------------------- separable fixed-point blur
 * taps[] is a Q8 kernel (sum 256) of odd length k.  Horizontal pass exact in
 * Q8, vertical pass exact in Q16, out = (sum + 32768) >> 16, REFLECT_101.    */
void or_sepblur_q8(const u8 *src, u8 *dst, int h, int w, const int *taps, int k)
{
    int r = k / 2;
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int t = -r; t <= r; t++)
                s += (uint32_t)taps[t + r] * src[(size_t)y * w + reflect101(x + t, w)];
            tmp[(size_t)y * w + x] = s;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int t = -r; t <= r; t++)
                s += (uint32_t)taps[t + r] * tmp[(size_t)reflect101(y + t, h) * w + x];
            dst[(size_t)y * w + x] = (u8)((s + 32768u) >> 16);
        }
    free(tmp);
}

/* a2: GaussianBlur((5,5),0) == Q8 taps [16 64 96 64 16] */
void or_gauss5(const u8 *src, u8 *dst, int h, int w)
{
    static const int taps[5] = {16, 64, 96, 64, 16};
    or_sepblur_q8(src, dst, h, w, taps, 5);
}

/* a3: normalize(src,None,0,255,NORM_MINMAX) u8->u8 (then astype(uint8) is a no-op) */
void or_normalize_minmax(const u8 *src, u8 *dst, long n)
{
    int mn = 255, mx = 0;
    for (long i = 0; i < n; i++) {
        if (src[i] < mn) mn = src[i];
        if (src[i] > mx) mx = src[i];
    }
    double smin = mn, smax = mx;
    double scale = (255.0 - 0.0) * ((smax - smin) > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0);
    double shift = 0.0 - smin * scale;
    /* convertTo u8->u8 runs in f32 with a FUSED multiply-add (cv2's AVX2/FMA3 dispatch of
     * convert_scale; verified: the unfused f32 and the f64 evaluation both differ from cv2). */
    float fs = (float)scale, fb = (float)shift;
    for (long i = 0; i < n; i++) {
        float v = fmaf((float)src[i], fs, fb);
        long iv = lrintf(v);
        dst[i] = (u8)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
    }
}

/* ------------------------------------------------- float32 frame-buffer ingest
 * The simulator path: _fboToImage (opencv.py:101-108) hands doCanny (opencv.py:21-28) a float32 BGR
 * view, so cvtColor / GaussianBlur / normalize run cv2's CV_32F kernels.  Summation orders below are
 * the ones of cv2 4.13's SIMD (FMA3) dispatch, found by differential testing against the wheel
 * (tests/test_oracle.py::test_f32_chain_vs_cv2): every variant with another order differs from cv2
 * in 2-10 % of the pixels.  cv2's scalar tail (last w mod 16 columns of GaussianBlur) uses another
 * order; the restatement follows the vector order everywhere and is pinned on widths % 16 == 0.
 * src is addressed with element strides (the reference's view has negative row and channel strides). */
void or_f32_bgr2gray(const float *src, long sy, long sx, long sc, float *gray, int h, int w)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float *p = src + (long)y * sy + (long)x * sx;
            float b = p[0], g = p[sc], r = p[2 * sc];
            gray[(size_t)y * w + x] = fmaf(r, 0.299f, fmaf(b, 0.114f, g * 0.587f));
        }
}

void or_f32_gauss5(const float *src, float *dst, int h, int w)
{
    const float k0 = 0.375f, k1 = 0.25f, k2 = 0.0625f;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)h * w);
    for (int y = 0; y < h; y++) {
        const float *r = src + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float xm2 = r[reflect101(x - 2, w)], xm1 = r[reflect101(x - 1, w)], x0 = r[x];
            float x1 = r[reflect101(x + 1, w)], x2 = r[reflect101(x + 2, w)];
            tmp[(size_t)y * w + x] = fmaf(x0, k0, (xm1 + x1) * k1) + (xm2 + x2) * k2;
        }
    }
    for (int y = 0; y < h; y++) {
        const float *rm2 = tmp + (size_t)reflect101(y - 2, h) * w, *rm1 = tmp + (size_t)reflect101(y - 1, h) * w;
        const float *r0 = tmp + (size_t)y * w;
        const float *r1 = tmp + (size_t)reflect101(y + 1, h) * w, *r2 = tmp + (size_t)reflect101(y + 2, h) * w;
        for (int x = 0; x < w; x++) dst[(size_t)y * w + x] = (r0[x] * k0 + (rm1[x] + r1[x]) * k1) + (rm2[x] + r2[x]) * k2;
    }
    free(tmp);
}

/* normalize(src f32, None, 0, 255, NORM_MINMAX) -> f32 (dst_f32) and the reference's .astype(np.uint8)
 * truncation (dst_u8); either output may be NULL */
void or_f32_normalize_minmax(const float *src, float *dst_f32, u8 *dst_u8, long n)
{
    float mn = src[0], mx = src[0];
    for (long i = 1; i < n; i++) {
        if (src[i] < mn) mn = src[i];
        if (src[i] > mx) mx = src[i];
    }
    double smin = mn, smax = mx;
    double scale = (255.0 - 0.0) * ((smax - smin) > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0);
    /* cv::normalize, rtype == CV_32F: the scale is rounded to float BEFORE the shift is formed, and the
     * product smin*scale is rounded to float too */
    scale = (double)(float)scale;
    double shift = (double)(float)0.0 - (double)(float)(smin * scale);
    float fs = (float)scale, fb = (float)shift;
    for (long i = 0; i < n; i++) {
        float v = fmaf(src[i], fs, fb);
        if (dst_f32) dst_f32[i] = v;
        if (dst_u8) {
            int iv = (int)v; /* numpy astype: truncation toward zero */
            dst_u8[i] = (u8)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
        }
    }
}

/* ---------------------------------------------------------------- a4: Canny
 * nms_out (optional): 0 = none, 1 = weak candidate, 2 = strong candidate.    */
/* cn = 1: a gray image.  cn = 3: cv2.Canny on an interleaved 3-channel image ("Canny accepts
 * 3-channel input") -- the Sobel derivatives are taken per channel and every pixel keeps the channel with the
 * largest |dx| + |dy| (the first one on ties); NMS and hysteresis then run on that single (dx, dy, magnitude). */
static void canny_cn(const u8 *src, int cn, u8 *dst, int h, int w, double lo_d, double hi_d, u8 *nms_out)
{
    if (lo_d > hi_d) { double t = lo_d; lo_d = hi_d; hi_d = t; }
    int lo = (int)floor(lo_d), hi = (int)floor(hi_d);
    size_t n = (size_t)h * w;
    short *dx = (short *)malloc(n * sizeof(short));
    short *dy = (short *)malloc(n * sizeof(short));
    int *mag = (int *)calloc((size_t)(h + 2) * (w + 2), sizeof(int));
    u8 *map = (u8 *)calloc(n, 1);
    const int ms = w + 2;
    for (int y = 0; y < h; y++) {
        int ym = clampi(y - 1, 0, h - 1), yp = clampi(y + 1, 0, h - 1);
        for (int x = 0; x < w; x++) {
            int xm = clampi(x - 1, 0, w - 1), xp = clampi(x + 1, 0, w - 1);
            int bgx = 0, bgy = 0, bm = -1;
            for (int c = 0; c < cn; c++) {
#define P(yy, xx) ((int)src[((size_t)(yy) * w + (xx)) * cn + c])
                int gx = (P(ym, xp) + 2 * P(y, xp) + P(yp, xp)) - (P(ym, xm) + 2 * P(y, xm) + P(yp, xm));
                int gy = (P(yp, xm) + 2 * P(yp, x) + P(yp, xp)) - (P(ym, xm) + 2 * P(ym, x) + P(ym, xp));
#undef P
                int m = abs(gx) + abs(gy);
                if (m > bm) { bm = m; bgx = gx; bgy = gy; }
            }
            dx[(size_t)y * w + x] = (short)bgx;
            dy[(size_t)y * w + x] = (short)bgy;
            mag[(size_t)(y + 1) * ms + x + 1] = bm;
        }
    }
    const int TG22 = 13573;
    int *stack = (int *)malloc(n * sizeof(int));
    size_t sp = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int *m = mag + (size_t)(y + 1) * ms + x + 1;
            int mv = m[0];
            if (mv <= lo) continue;
            int gx = dx[(size_t)y * w + x], gy = dy[(size_t)y * w + x];
            int ax = abs(gx), ay = abs(gy) << 15;
            int tg22x = ax * TG22;
            int keep;
            if (ay < tg22x) keep = (mv > m[-1] && mv >= m[1]);
            else {
                int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = (mv > m[-ms] && mv >= m[ms]);
                else {
                    int s = ((gx ^ gy) < 0) ? -1 : 1;
                    keep = (mv > m[-ms - s] && mv > m[ms + s]);
                }
            }
            if (!keep) continue;
            if (mv > hi) { map[(size_t)y * w + x] = 2; stack[sp++] = y * w + x; }
            else map[(size_t)y * w + x] = 1;
        }
    if (nms_out) memcpy(nms_out, map, n);
    /* hysteresis: 8-connected flood from strong pixels through weak ones */
    while (sp) {
        int p = stack[--sp];
        int y = p / w, x = p % w;
        for (int yy = y - 1; yy <= y + 1; yy++)
            for (int xx = x - 1; xx <= x + 1; xx++) {
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                size_t q = (size_t)yy * w + xx;
                if (map[q] == 1) { map[q] = 2; stack[sp++] = (int)q; }
            }
    }
    for (size_t i = 0; i < n; i++) dst[i] = (map[i] == 2) ? 255 : 0;
    free(dx); free(dy); free(mag); free(map); free(stack);
}

void or_canny(const u8 *src, u8 *dst, int h, int w, double lo_d, double hi_d, u8 *nms_out)
{
    canny_cn(src, 1, dst, h, w, lo_d, hi_d, nms_out);
}
void or_canny_c3(const u8 *src, u8 *dst, int h, int w, double lo_d, double hi_d, u8 *nms_out)
{
    canny_cn(src, 3, dst, h, w, lo_d, hi_d, nms_out);
}

/* ------------------------------------------------------ a5: standard Hough */
static int hough_numangle(double min_theta, double max_theta, double theta)
{
    int numangle = (int)floor((max_theta - min_theta) / theta) + 1;
    if (numangle > 1 && fabs(M_PI - (numangle - 1) * theta) < theta / 2) --numangle;
    return numangle;
}
int or_hough_numangle(double theta) { return hough_numangle(0.0, M_PI, theta); }
int or_hough_numrho(int h, int w, double rho) { return (int)lrint(((w + h) * 2 + 1) / rho); }

/* trig table of HoughLinesStandard: the angle ACCUMULATES in float32 */
void or_hough_trig_std(float rho, float theta, int numangle, float *tab_sin, float *tab_cos)
{
    float irho = 1.f / rho;
    float ang = 0.f;
    for (int n = 0; n < numangle; ang += theta, n++) {
        tab_sin[n] = (float)(sin((double)ang) * irho);
        tab_cos[n] = (float)(cos((double)ang) * irho);
    }
}

/* accum: (numangle+2) x (numrho+2) int32, caller-zeroed */
void or_hough_accum(const u8 *edges, int h, int w, float rho, float theta, int *accum)
{
    int numangle = hough_numangle(0.0, M_PI, theta);
    int numrho = or_hough_numrho(h, w, rho);
    float *ts = (float *)malloc(sizeof(float) * numangle), *tc = (float *)malloc(sizeof(float) * numangle);
    or_hough_trig_std(rho, theta, numangle, ts, tc);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (!edges[(size_t)y * w + x]) continue;
            for (int n = 0; n < numangle; n++) {
                float fx = (float)x * tc[n];
                float fy = (float)y * ts[n];
                int r = (int)lrintf(fx + fy);
                r += (numrho - 1) / 2;
                accum[(size_t)(n + 1) * (numrho + 2) + r + 1]++;
            }
        }
    free(ts); free(tc);
}

static const int *g_sort_accum;
static int hough_cmp(const void *a, const void *b)
{
    int l1 = *(const int *)a, l2 = *(const int *)b;
    int v1 = g_sort_accum[l1], v2 = g_sort_accum[l2];
    if (v1 != v2) return v1 > v2 ? -1 : 1;
    return l1 < l2 ? -1 : (l1 > l2 ? 1 : 0);
}

/* returns the number of lines found (may exceed max_lines; only max_lines are written).
 * lines: (rho,theta) f32 pairs; votes: int per line; accum_out optional. */
int or_hough_lines(const u8 *edges, int h, int w, float rho, float theta, int threshold,
                   float *lines, int *votes, int max_lines, int *accum_out)
{
    int numangle = hough_numangle(0.0, M_PI, theta);
    int numrho = or_hough_numrho(h, w, rho);
    size_t asz = (size_t)(numangle + 2) * (numrho + 2);
    int *accum = (int *)calloc(asz, sizeof(int));
    or_hough_accum(edges, h, w, rho, theta, accum);
    int *buf = (int *)malloc(sizeof(int) * (size_t)numangle * numrho);
    int total = 0;
    for (int r = 0; r < numrho; r++)
        for (int n = 0; n < numangle; n++) {
            int base = (n + 1) * (numrho + 2) + r + 1;
            if (accum[base] > threshold && accum[base] > accum[base - 1] && accum[base] >= accum[base + 1] &&
                accum[base] > accum[base - numrho - 2] && accum[base] >= accum[base + numrho + 2])
                buf[total++] = base;
        }
    g_sort_accum = accum;
    qsort(buf, total, sizeof(int), hough_cmp);
    double scale = 1. / (numrho + 2);
    for (int i = 0; i < total && i < max_lines; i++) {
        int idx = buf[i];
        int n = (int)floor(idx * scale) - 1;
        int r = idx - (n + 1) * (numrho + 2) - 1;
        lines[2 * i] = (r - (numrho - 1) * 0.5f) * rho;
        lines[2 * i + 1] = 0.f + n * theta;
        if (votes) votes[i] = accum[idx];
    }
    if (accum_out) memcpy(accum_out, accum, asz * sizeof(int));
    free(buf); free(accum);
    return total;
}

/* ------------------------------------------- a6: progressive probabilistic */
static inline unsigned rng_next(uint64_t *st)
{
    *st = (uint64_t)(unsigned)*st * 4164903690U + (unsigned)(*st >> 32);
    return (unsigned)*st;
}

static int *g_pp_trace = 0;
static int g_pp_trace_cap = 0, g_pp_trace_n = 0;
void or_ppht_trace(int *buf, int cap) { g_pp_trace = buf; g_pp_trace_cap = cap; g_pp_trace_n = 0; }
int or_ppht_trace_count(void) { return g_pp_trace_n; }

/* cv2 seeds cv::RNG((uint64)-1) on every call; another seed gives another visiting order -- used only to measure
 * how far the result depends on that order (tests/test_ppht_order.py, scripts/ppht_order_sensitivity.py). */
static uint64_t g_pp_seed = (uint64_t)-1;
void or_ppht_seed(uint64_t seed) { g_pp_seed = seed; }

int or_hough_lines_p(const u8 *edges, int h, int w, float rho, float theta, int threshold,
                     int line_length, int line_gap, int *segs, int max_segs)
{
    const int width = w, height = h;
    float irho = 1 / rho;
    uint64_t rng = g_pp_seed;
    int numangle = hough_numangle(0.0, M_PI, theta);
    int numrho = (int)lrint(((width + height) * 2 + 1) / rho);
    int *accum = (int *)calloc((size_t)numangle * numrho, sizeof(int));
    u8 *mask = (u8 *)malloc((size_t)h * w);
    float *ttab = (float *)malloc(sizeof(float) * 2 * numangle);
    for (int n = 0; n < numangle; n++) {
        ttab[n * 2] = (float)(cos((double)n * theta) * irho);
        ttab[n * 2 + 1] = (float)(sin((double)n * theta) * irho);
    }
    int *nz = (int *)malloc(sizeof(int) * 2 * (size_t)h * w);
    int count = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (edges[(size_t)y * w + x]) { mask[(size_t)y * w + x] = 1; nz[2 * count] = x; nz[2 * count + 1] = y; count++; }
            else mask[(size_t)y * w + x] = 0;
        }
    int nlines = 0;
    for (; count > 0; count--) {
        int idx = (int)(rng_next(&rng) % (unsigned)count);
        int max_val = threshold - 1, max_n = 0;
        int j = nz[2 * idx], i = nz[2 * idx + 1];
        int line_end[2][2] = {{0, 0}, {0, 0}}; /* [k] = {x, y} */
        float a, b;
        int k, x0, y0, dx0, dy0, xflag, good_line;
        const int shift = 16;
        nz[2 * idx] = nz[2 * (count - 1)];
        nz[2 * idx + 1] = nz[2 * (count - 1) + 1];
        if (!mask[(size_t)i * width + j]) continue;
        int *adata = accum;
        for (int n = 0; n < numangle; n++, adata += numrho) {
            float fx = (float)j * ttab[n * 2];
            float fy = (float)i * ttab[n * 2 + 1];
            int r = (int)lrintf(fx + fy);
            r += (numrho - 1) / 2;
            int val = ++adata[r];
            if (max_val < val) { max_val = val; max_n = n; }
        }
        if (max_val < threshold) continue;
        a = -ttab[max_n * 2 + 1];
        b = ttab[max_n * 2];
        x0 = j; y0 = i;
        if (fabsf(a) > fabsf(b)) {
            xflag = 1;
            dx0 = a > 0 ? 1 : -1;
            dy0 = (int)lrintf(b * (float)(1 << shift) / fabsf(a));
            y0 = (y0 << shift) + (1 << (shift - 1));
        } else {
            xflag = 0;
            dy0 = b > 0 ? 1 : -1;
            dx0 = (int)lrintf(a * (float)(1 << shift) / fabsf(b));
            x0 = (x0 << shift) + (1 << (shift - 1));
        }
        for (k = 0; k < 2; k++) {
            int gap = 0, x = x0, y = y0, dx = dx0, dy = dy0;
            if (k > 0) { dx = -dx; dy = -dy; }
            for (;; x += dx, y += dy) {
                int i1, j1;
                if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
                if (j1 < 0 || j1 >= width || i1 < 0 || i1 >= height) break;
                if (mask[(size_t)i1 * width + j1]) { gap = 0; line_end[k][1] = i1; line_end[k][0] = j1; }
                else if (++gap > line_gap) break;
            }
        }
        good_line = abs(line_end[1][0] - line_end[0][0]) >= line_length ||
                    abs(line_end[1][1] - line_end[0][1]) >= line_length;
        if (g_pp_trace && g_pp_trace_n < g_pp_trace_cap) {
            int *o = g_pp_trace + 10 * g_pp_trace_n++;
            o[0] = j; o[1] = i; o[2] = max_val; o[3] = max_n; o[4] = line_end[0][0]; o[5] = line_end[0][1];
            o[6] = line_end[1][0]; o[7] = line_end[1][1]; o[8] = good_line; o[9] = dx0 * 4 + dy0 % 4;
        }
        for (k = 0; k < 2; k++) {
            int x = x0, y = y0, dx = dx0, dy = dy0;
            if (k > 0) { dx = -dx; dy = -dy; }
            for (;; x += dx, y += dy) {
                int i1, j1;
                if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
                u8 *md = mask + (size_t)i1 * width + j1;
                if (*md) {
                    if (good_line) {
                        adata = accum;
                        for (int n = 0; n < numangle; n++, adata += numrho) {
                            float fx = (float)j1 * ttab[n * 2];
                            float fy = (float)i1 * ttab[n * 2 + 1];
                            int r = (int)lrintf(fx + fy);
                            r += (numrho - 1) / 2;
                            adata[r]--;
                        }
                    }
                    *md = 0;
                }
                if (i1 == line_end[k][1] && j1 == line_end[k][0]) break;
            }
        }
        if (good_line) {
            if (nlines < max_segs) {
                segs[4 * nlines] = line_end[0][0]; segs[4 * nlines + 1] = line_end[0][1];
                segs[4 * nlines + 2] = line_end[1][0]; segs[4 * nlines + 3] = line_end[1][1];
            }
            nlines++;
        }
    }
    free(accum); free(mask); free(ttab); free(nz);
    return nlines;
}

/* ------------------------------------------------------------------ a7: LSD */
#define NOTDEF (-1024.0)
#define M_3_2_PI (3 * M_PI / 2)
#define M_2__PI (2 * M_PI)
static const double DEG_TO_RADS = M_PI / 180;

float or_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* Q8 Gaussian taps, error-diffused (OpenCV getGaussianKernelFixedPoint_ED). */
void or_gauss_taps_q8(double sigma, int k, int *taps)
{
    double *ker = (double *)malloc(sizeof(double) * k);
    double sum = 0, scale2x = -0.5 / (sigma * sigma);
    for (int i = 0; i < k; i++) {
        double x = i - (k - 1) * 0.5;
        ker[i] = exp(scale2x * x * x);
        sum += ker[i];
    }
    for (int i = 0; i < k; i++) ker[i] /= sum;
    double err = 0;
    long tot = 0;
    for (int i = 0; i < k / 2; i++) {
        double adj = ker[i] * 256.0 + err;
        long v0 = lrint(adj);
        err = adj - (double)v0;
        taps[i] = taps[k - 1 - i] = (int)v0;
        tot += v0;
    }
    taps[k / 2] = (int)(256 - 2 * tot);
    free(ker);
}

/* resize INTER_LINEAR_EXACT u8 (fixed point, Q8 weights per axis). */
void or_resize_linear_exact(const u8 *src, int h, int w, u8 *dst, int dh, int dw, double inv_scale)
{
    int *xi = (int *)malloc(sizeof(int) * dw), *xw = (int *)malloc(sizeof(int) * dw);
    int *yi = (int *)malloc(sizeof(int) * dh), *yw = (int *)malloc(sizeof(int) * dh);
    /* resize(src, Size(), fx, fy): dsize is derived from fx, fy and the sampling step stays 1/fx
     * (NOT src/dst size ratio) -- matters for sizes that are not multiples of 1/scale. */
    double sx = 1.0 / inv_scale, sy = 1.0 / inv_scale;
    for (int d = 0; d < dw; d++) {
        double pos = (d + 0.5) * sx - 0.5;
        int i = (int)floor(pos);
        double f = pos - i;
        if (i < 0) { i = 0; f = 0; }
        if (i >= w - 1) { i = w - 1; f = 0; }
        xi[d] = i; xw[d] = (int)lrint(f * 256.0);
    }
    for (int d = 0; d < dh; d++) {
        double pos = (d + 0.5) * sy - 0.5;
        int i = (int)floor(pos);
        double f = pos - i;
        if (i < 0) { i = 0; f = 0; }
        if (i >= h - 1) { i = h - 1; f = 0; }
        yi[d] = i; yw[d] = (int)lrint(f * 256.0);
    }
    for (int y = 0; y < dh; y++) {
        const u8 *r0 = src + (size_t)yi[y] * w;
        const u8 *r1 = src + (size_t)(yi[y] + 1 < h ? yi[y] + 1 : h - 1) * w;
        for (int x = 0; x < dw; x++) {
            int x0 = xi[x], x1 = x0 + 1 < w ? x0 + 1 : w - 1;
            uint32_t a = (uint32_t)(256 - xw[x]) * r0[x0] + (uint32_t)xw[x] * r0[x1];
            uint32_t b = (uint32_t)(256 - xw[x]) * r1[x0] + (uint32_t)xw[x] * r1[x1];
            uint32_t v = (uint32_t)(256 - yw[y]) * a + (uint32_t)yw[y] * b;
            dst[(size_t)y * dw + x] = (u8)((v + 32768u) >> 16);
        }
    }
    free(xi); free(xw); free(yi); free(yw);
}

/* scaled image of LSD: Gaussian (sigma = sigma_scale/scale) then exact-linear resize */
void or_lsd_scaled_dims(int h, int w, double scale, int *dh, int *dw)
{
    *dw = (int)lrint(w * scale);
    *dh = (int)lrint(h * scale);
}
void or_lsd_scale_image(const u8 *gray, int h, int w, double scale, double sigma_scale, u8 *dst)
{
    const double sigma = (scale < 1) ? (sigma_scale / scale) : sigma_scale;
    const unsigned hh = (unsigned)ceil(sigma * sqrt(2 * 3.0 * log(10.0)));
    int k = 1 + 2 * (int)hh;
    int *taps = (int *)malloc(sizeof(int) * k);
    or_gauss_taps_q8(sigma, k, taps);
    u8 *blur = (u8 *)malloc((size_t)h * w);
    or_sepblur_q8(gray, blur, h, w, taps, k);
    int dh, dw;
    or_lsd_scaled_dims(h, w, scale, &dh, &dw);
    or_resize_linear_exact(blur, h, w, dst, dh, dw, scale);
    free(taps); free(blur);
}

typedef struct { int x, y; double angle, modgrad; } RegPt;
typedef struct {
    int W, H;
    double *angles, *modgrad;
    u8 *used;
} LsdState;

static int lsd_aligned(const LsdState *s, int x, int y, double theta, double prec)
{
    if (x < 0 || y < 0 || x >= s->W || y >= s->H) return 0;
    double a = s->angles[(size_t)y * s->W + x];
    if (a == NOTDEF) return 0;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > M_3_2_PI) {
        n_theta -= M_2__PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}

static double angle_diff_signed(double a, double b)
{
    double diff = a - b;
    while (diff <= -M_PI) diff += M_2__PI;
    while (diff > M_PI) diff -= M_2__PI;
    return diff;
}

static int lsd_region_grow(LsdState *s, int sx, int sy, RegPt *reg, double *reg_angle_out, double prec)
{
    int W = s->W, H = s->H, n = 0;
    double reg_angle = s->angles[(size_t)sy * W + sx];
    reg[n].x = sx; reg[n].y = sy; reg[n].angle = reg_angle; reg[n].modgrad = s->modgrad[(size_t)sy * W + sx];
    n++;
    float sumdx = (float)cos(reg_angle);
    float sumdy = (float)sin(reg_angle);
    s->used[(size_t)sy * W + sx] = 1;
    for (int i = 0; i < n; i++) {
        int px = reg[i].x, py = reg[i].y;
        int xx_min = px - 1 > 0 ? px - 1 : 0, xx_max = px + 1 < W - 1 ? px + 1 : W - 1;
        int yy_min = py - 1 > 0 ? py - 1 : 0, yy_max = py + 1 < H - 1 ? py + 1 : H - 1;
        for (int yy = yy_min; yy <= yy_max; ++yy)
            for (int xx = xx_min; xx <= xx_max; ++xx) {
                size_t q = (size_t)yy * W + xx;
                if (s->used[q] != 1 && lsd_aligned(s, xx, yy, reg_angle, prec)) {
                    double angle = s->angles[q];
                    s->used[q] = 1;
                    reg[n].x = xx; reg[n].y = yy; reg[n].angle = angle; reg[n].modgrad = s->modgrad[q];
                    n++;
                    sumdx += cosf((float)angle);
                    sumdy += sinf((float)angle);
                    reg_angle = or_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
                }
            }
    }
    *reg_angle_out = reg_angle;
    return n;
}

typedef struct { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; } LsdRect;

static double lsd_get_theta(const RegPt *reg, int n, double x, double y, double reg_angle, double prec)
{
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (int i = 0; i < n; i++) {
        double dx = (double)reg[i].x - x, dy = (double)reg[i].y - y, wgt = reg[i].modgrad;
        Ixx += dy * dy * wgt;
        Iyy += dx * dx * wgt;
        Ixy -= dx * dy * wgt;
    }
    double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)or_fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                           : (double)or_fast_atan2((float)Ixy, (float)(lambda - Iyy));
    theta *= DEG_TO_RADS;
    if (fabs(angle_diff_signed(theta, reg_angle)) > prec) theta += M_PI;
    return theta;
}

static void lsd_region2rect(const RegPt *reg, int n, double reg_angle, double prec, double p, LsdRect *rec)
{
    double x = 0, y = 0, sum = 0;
    for (int i = 0; i < n; i++) {
        double wgt = reg[i].modgrad;
        x += (double)reg[i].x * wgt;
        y += (double)reg[i].y * wgt;
        sum += wgt;
    }
    x /= sum; y /= sum;
    double theta = lsd_get_theta(reg, n, x, y, reg_angle, prec);
    double dx = cos(theta), dy = sin(theta);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (int i = 0; i < n; i++) {
        double rdx = (double)reg[i].x - x, rdy = (double)reg[i].y - y;
        double l = rdx * dx + rdy * dy;
        double wv = -rdx * dy + rdy * dx;
        if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
        if (wv > w_max) w_max = wv; else if (wv < w_min) w_min = wv;
    }
    rec->x1 = x + l_min * dx; rec->y1 = y + l_min * dy;
    rec->x2 = x + l_max * dx; rec->y2 = y + l_max * dy;
    rec->width = w_max - w_min;
    rec->x = x; rec->y = y; rec->theta = theta; rec->dx = dx; rec->dy = dy; rec->prec = prec; rec->p = p;
    if (rec->width < 1.0) rec->width = 1.0;
}

static inline double dist2d(double x1, double y1, double x2, double y2)
{
    return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}

/* refine=LSD_REFINE_STD pieces */
static int lsd_region_grow_from(LsdState *s, int sx, int sy, RegPt *reg, double *reg_angle, double prec)
{
    return lsd_region_grow(s, sx, sy, reg, reg_angle, prec);
}

static int lsd_reduce_region_radius(LsdState *s, RegPt *reg, int *pn, double *reg_angle, double prec, double p,
                                    LsdRect *rec, double density, double density_th)
{
    int n = *pn;
    double xc = (double)reg[0].x, yc = (double)reg[0].y;
    double radSq1 = (xc - rec->x1) * (xc - rec->x1) + (yc - rec->y1) * (yc - rec->y1);
    double radSq2 = (xc - rec->x2) * (xc - rec->x2) + (yc - rec->y2) * (yc - rec->y2);
    double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
    while (density < density_th) {
        radSq *= 0.75 * 0.75;
        for (int i = 0; i < n; i++) {
            double dx = (double)reg[i].x - xc, dy = (double)reg[i].y - yc;
            if (dx * dx + dy * dy > radSq) {
                s->used[(size_t)reg[i].y * s->W + reg[i].x] = 0;
                RegPt t = reg[i]; reg[i] = reg[n - 1]; reg[n - 1] = t;
                n--; i--;
            }
        }
        if (n < 2) { *pn = n; return 0; }
        lsd_region2rect(reg, n, *reg_angle, prec, p, rec);
        density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
    }
    *pn = n;
    return 1;
}

static int lsd_refine(LsdState *s, RegPt *reg, int *pn, double *reg_angle, double prec, double p, LsdRect *rec,
                      double density_th)
{
    int n = *pn;
    double density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
    if (density >= density_th) return 1;
    double xc = (double)reg[0].x, yc = (double)reg[0].y;
    double ang_c = reg[0].angle;
    double sum = 0, s_sum = 0;
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        s->used[(size_t)reg[i].y * s->W + reg[i].x] = 0;
        if (dist2d(xc, yc, (double)reg[i].x, (double)reg[i].y) < rec->width) {
            double ang_d = angle_diff_signed(reg[i].angle, ang_c);
            sum += ang_d;
            s_sum += ang_d * ang_d;
            ++cnt;
        }
    }
    double mean_angle = sum / (double)cnt;
    double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
    n = lsd_region_grow_from(s, reg[0].x, reg[0].y, reg, reg_angle, tau);
    *pn = n;
    if (n < 2) return 0;
    lsd_region2rect(reg, n, *reg_angle, prec, p, rec);
    density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
    if (density < density_th) return lsd_reduce_region_radius(s, reg, pn, reg_angle, prec, p, rec, density, density_th);
    return 1;
}


/* ---- refine = LSD_REFINE_ADV: number of false alarms of a rectangle.
 * nfa() and the two log-gamma approximations are the published LSD formulas; they are pinned exactly: every NFA
 * value cv2 4.13 returns for an unimproved rectangle inverts to an integer pair (points, aligned points).
 * The rectangle's pixel enumeration below is what those pairs identify, and it reproduces ALL of them (2333 of
 * 2333 unimproved rectangles of the 19 cube fixtures at both scales; then every refine-2 output of the 21
 * fixtures, tests/test_lsd2_oracle.py): the topmost vertex (ties: the left one) heads a left and a right chain
 * of two edges each; an edge that crosses no integer row has slope 0; rows y = ceil(min vertex y) ..
 * ceil(max vertex y); in a row the pixels ceil(left) .. floor(right), both limits evaluated directly from a
 * vertex and a slope (nothing is accumulated); the right limit changes edge at the first row at or below its
 * vertex, the LEFT limit one row later (that row still extrapolates the first left edge -- for a
 * near-horizontal rectangle this adds up to a row's worth of pixels, and cv2 does the same). */
static double lsd_log_gamma(double x)
{
    if (x > 15.0) /* Windschitl */
        return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479,
                                83.8676043424, 2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
    double b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}

static int lsd_double_equal(double a, double b)
{
    if (a == b) return 1;
    double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}

double or_lsd_nfa(int n, int k, double p, double LOG_NT)
{
    if (n == 0 || k == 0) return -LOG_NT;
    if (n == k) return -LOG_NT - (double)n * log10(p);
    double p_term = p / (1 - p);
    double log1term = lsd_log_gamma((double)n + 1) - lsd_log_gamma((double)k + 1) - lsd_log_gamma((double)(n - k) + 1) +
                      (double)k * log(p) + (double)(n - k) * log(1.0 - p);
    double term = exp(log1term);
    if (lsd_double_equal(term, 0)) {
        if (k > n * p) return -log1term / M_LN10 - LOG_NT;
        else return -LOG_NT;
    }
    double bin_tail = term;
    double tolerance = 0.1;
    for (int i = k + 1; i <= n; ++i) {
        double bin_term = (double)(n - i + 1) / (double)i;
        double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1) {
            double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
            if (err < tolerance * fabs(-log10(bin_tail) - LOG_NT) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - LOG_NT;
}

static int g_last_n, g_last_k;
static double lsd_rect_nfa(const LsdState *s, const LsdRect *rec, double LOG_NT)
{
    int total_pts = 0, alg_pts = 0;
    const double hw = rec->width / 2.0, dyhw = rec->dy * hw, dxhw = rec->dx * hw;
    /* the four vertices in rotation order; m = the topmost one, its two neighbours head the left and right chains */
    const double vx[4] = {rec->x1 - dyhw, rec->x2 - dyhw, rec->x2 + dyhw, rec->x1 + dyhw};
    const double vy[4] = {rec->y1 + dxhw, rec->y2 + dxhw, rec->y2 - dxhw, rec->y1 - dxhw};
    int m = 0; /* the topmost vertex; of two at the same height the left one */
    for (int i = 1; i < 4; i++)
        if (vy[i] < vy[m] || (vy[i] == vy[m] && vx[i] < vx[m])) m = i;
    const int a1 = (m + 1) & 3, b1 = (m + 3) & 3, op = (m + 2) & 3;
    const int ha = vy[m] == vy[a1], hb = vy[m] == vy[b1];
    const double ta = ha ? 0.0 : (vx[m] - vx[a1]) / (vy[m] - vy[a1]), tb = hb ? 0.0 : (vx[m] - vx[b1]) / (vy[m] - vy[b1]);
    const int a_left = (ha || hb) ? vx[a1] < vx[b1] : ta < tb;
    const int L1 = a_left ? a1 : b1, R1 = a_left ? b1 : a1;
    /* an edge that crosses no integer row counts as horizontal: its slope is taken as 0 */
#define LSD_SLOPE(p, q) ((ceil(vy[p]) == ceil(vy[q])) ? 0.0 : (vx[p] - vx[q]) / (vy[p] - vy[q]))
    const double fl = LSD_SLOPE(m, L1), sl = LSD_SLOPE(L1, op), fr = LSD_SLOPE(m, R1), sr = LSD_SLOPE(R1, op);
#undef LSD_SLOPE
    const double ya_d = ceil(vy[m]), yb_d = ceil(vy[op]);
    int ya = (int)fmax(fmin(ya_d, 1e9), -1e9), yb = (int)fmax(fmin(yb_d, 1e9), -1e9);
    if (yb > s->H - 1) yb = s->H - 1;
    for (int y = ya < 0 ? 0 : ya; y <= yb; ++y) {
        const int l2 = (double)(y - 1) >= vy[L1]; /* the left limit changes edge one row late */
        const int r2 = (double)y >= vy[R1];
        const double lx = l2 ? vx[L1] + ((double)y - vy[L1]) * sl : vx[m] + ((double)y - vy[m]) * fl;
        const double rx = r2 ? vx[R1] + ((double)y - vy[R1]) * sr : vx[m] + ((double)y - vy[m]) * fr;
        const double a_d = ceil(lx), b_d = floor(rx);
        const int a = !(a_d > 0) ? 0 : (int)fmin(a_d, 1e9);
        const int b = !(b_d < s->W - 1) ? s->W - 1 : (int)fmax(b_d, -1e9);
        for (int x = a; x <= b; ++x) {
            ++total_pts;
            if (lsd_aligned(s, x, y, rec->theta, rec->prec)) ++alg_pts;
        }
    }
    g_last_n = total_pts; g_last_k = alg_pts;
    return or_lsd_nfa(total_pts, alg_pts, rec->p, LOG_NT);
}

static double lsd_rect_improve(const LsdState *s, LsdRect *rec, double LOG_NT, double LOG_EPS)
{
    const double delta = 0.5, delta_2 = delta / 2.0;
    double log_nfa = lsd_rect_nfa(s, rec, LOG_NT);
    if (log_nfa > LOG_EPS) return log_nfa;
    LsdRect r = *rec; /* finer precision */
    for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        r.prec = r.p * M_PI;
        double v = lsd_rect_nfa(s, &r, LOG_NT);
        if (v > log_nfa) { log_nfa = v; *rec = r; }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = *rec; /* reduce width */
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            r.width -= delta;
            double v = lsd_rect_nfa(s, &r, LOG_NT);
            if (v > log_nfa) { *rec = r; log_nfa = v; }
        }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = *rec; /* reduce one side */
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
            r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
            r.width -= delta;
            double v = lsd_rect_nfa(s, &r, LOG_NT);
            if (v > log_nfa) { *rec = r; log_nfa = v; }
        }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = *rec; /* reduce the other side */
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
            r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
            r.width -= delta;
            double v = lsd_rect_nfa(s, &r, LOG_NT);
            if (v > log_nfa) { *rec = r; log_nfa = v; }
        }
    }
    if (log_nfa > LOG_EPS) return log_nfa;
    r = *rec; /* even finer precision */
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            r.p /= 2;
            r.prec = r.p * M_PI;
            double v = lsd_rect_nfa(s, &r, LOG_NT);
            if (v > log_nfa) { log_nfa = v; *rec = r; }
        }
    }
    return log_nfa;
}

static double *g_nfa_out = 0, *g_rect_out = 0;
/* optional debug output: 12 doubles per emitted segment (x1 y1 x2 y2 width dx dy theta prec p, then the point
 * count and aligned count of the rectangle's last NFA evaluation) */
void or_lsd_set_rect_out(double *p) { g_rect_out = p; }
/* optional NFA output of the next or_lsd call(s) with refine = 2 (one double per segment) */
void or_lsd_set_nfa_out(double *p) { g_nfa_out = p; }

/* LSD on a u8 gray image, refine in {0,1,2}.  lines: 4 f32 per segment; width/prec f64 (optional).
 * Optional debug outputs: scaled_out (dh*dw u8), angles_out (dh*dw f64).
 * Returns number of segments (only max_segs written). */
int or_lsd(const u8 *gray, int h, int w, int refine, double scale, double sigma_scale, double quant, double ang_th,
           double log_eps, double density_th, int n_bins, float *lines, double *widths, double *precs, int max_segs,
           u8 *scaled_out, double *angles_out)
{
    const double prec = M_PI * ang_th / 180;
    const double p = ang_th / 180;
    const double rho = quant / sin(prec);
    int W = w, H = h;
    u8 *img;
    if (scale != 1) {
        or_lsd_scaled_dims(h, w, scale, &H, &W);
        img = (u8 *)malloc((size_t)H * W);
        or_lsd_scale_image(gray, h, w, scale, sigma_scale, img);
    } else {
        img = (u8 *)malloc((size_t)H * W);
        memcpy(img, gray, (size_t)H * W);
    }
    if (scaled_out) memcpy(scaled_out, img, (size_t)H * W);
    size_t N = (size_t)H * W;
    LsdState st;
    st.W = W; st.H = H;
    st.angles = (double *)malloc(sizeof(double) * N);
    st.modgrad = (double *)calloc(N, sizeof(double));
    st.used = (u8 *)calloc(N, 1);
    for (size_t i = 0; i < N; i++) st.angles[i] = NOTDEF;
    double max_grad = -1;
    for (int y = 0; y < H - 1; y++)
        for (int x = 0; x < W - 1; x++) {
            int DA = img[(size_t)(y + 1) * W + x + 1] - img[(size_t)y * W + x];
            int BC = img[(size_t)y * W + x + 1] - img[(size_t)(y + 1) * W + x];
            int gx = DA + BC, gy = DA - BC;
            double norm = sqrt((gx * gx + gy * gy) / 4.0);
            st.modgrad[(size_t)y * W + x] = norm;
            if (norm <= rho) st.angles[(size_t)y * W + x] = NOTDEF;
            else {
                st.angles[(size_t)y * W + x] = or_fast_atan2((float)gx, (float)-gy) * DEG_TO_RADS;
                if (norm > max_grad) max_grad = norm;
            }
        }
    if (angles_out) memcpy(angles_out, st.angles, sizeof(double) * N);
    /* descending-bin, row-major-within-bin order (counting sort) */
    double bin_coef = (max_grad > 0) ? (double)(n_bins - 1) / max_grad : 0;
    int *cnt = (int *)calloc((size_t)n_bins + 1, sizeof(int));
    size_t NP = (size_t)(H - 1) * (W - 1);
    int *bins = (int *)malloc(sizeof(int) * (NP ? NP : 1));
    size_t q = 0;
    for (int y = 0; y < H - 1; y++)
        for (int x = 0; x < W - 1; x++) {
            int b = (int)(st.modgrad[(size_t)y * W + x] * bin_coef);
            if (b > n_bins - 1) b = n_bins - 1;
            bins[q++] = b;
            cnt[b]++;
        }
    int *start = (int *)malloc(sizeof(int) * n_bins);
    int acc = 0;
    for (int b = n_bins - 1; b >= 0; b--) { start[b] = acc; acc += cnt[b]; }
    int *order = (int *)malloc(sizeof(int) * (NP ? NP : 1));
    q = 0;
    for (int y = 0; y < H - 1; y++)
        for (int x = 0; x < W - 1; x++) { order[start[bins[q]]++] = y * W + x; q++; }

    const double LOG_NT = 5 * (log10((double)W) + log10((double)H)) / 2 + log10(11.0);
    const size_t min_reg_size = (size_t)(-LOG_NT / log10(p));
    RegPt *reg = (RegPt *)malloc(sizeof(RegPt) * N);
    int nseg = 0;
    for (size_t i = 0; i < NP; i++) {
        int pidx = order[i];
        if (st.used[pidx] || st.angles[pidx] == NOTDEF) continue;
        double reg_angle;
        int n = lsd_region_grow(&st, pidx % W, pidx / W, reg, &reg_angle, prec);
        if ((size_t)n < min_reg_size) continue;
        LsdRect rec;
        double log_nfa = -1;
        lsd_region2rect(reg, n, reg_angle, prec, p, &rec);
        if (refine > 0) {
            if (!lsd_refine(&st, reg, &n, &reg_angle, prec, p, &rec, density_th)) continue;
            if (refine >= 2) {
                log_nfa = lsd_rect_improve(&st, &rec, LOG_NT, log_eps);
                if (log_nfa <= log_eps) continue;
            }
        }
        if (g_rect_out && nseg < max_segs) {
            double *o = g_rect_out + 12 * nseg;
            o[0] = rec.x1; o[1] = rec.y1; o[2] = rec.x2; o[3] = rec.y2; o[4] = rec.width; o[5] = rec.dx; o[6] = rec.dy;
            o[7] = rec.theta; o[8] = rec.prec; o[9] = rec.p; o[10] = g_last_n; o[11] = g_last_k;
        }
        rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
        if (scale != 1) { rec.x1 /= scale; rec.y1 /= scale; rec.x2 /= scale; rec.y2 /= scale; rec.width /= scale; }
        if (nseg < max_segs) {
            lines[4 * nseg] = (float)rec.x1; lines[4 * nseg + 1] = (float)rec.y1;
            lines[4 * nseg + 2] = (float)rec.x2; lines[4 * nseg + 3] = (float)rec.y2;
            if (widths) widths[nseg] = rec.width;
            if (precs) precs[nseg] = rec.p;
            if (g_nfa_out) g_nfa_out[nseg] = log_nfa;
        }
        nseg++;
    }
    free(img); free(st.angles); free(st.modgrad); free(st.used);
    free(cnt); free(bins); free(start); free(order); free(reg);
    return nseg;
}
