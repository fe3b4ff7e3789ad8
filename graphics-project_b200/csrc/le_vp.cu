// le_vp.cu -- vanishing-point stage (a9-a12): polar conversion, batched loss, Gaussian-sphere
// histogram of pairwise interpretation-plane intersections, (rho,theta) pair intersections.
// All f64 (the reference computes in numpy float64); inputs are a few hundred segments, so these
// kernels are latency-, not bandwidth-bound; they exist so the whole fit runs without leaving the GPU.
#include "le_common.cuh"

// a10  edges_to_polar_lines (opencv_renewed.py:16-17)
__global__ void k_seg_polar(const double *__restrict__ segs, int m, double *__restrict__ polar)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double ax = segs[4 * i], ay = segs[4 * i + 1], bx = segs[4 * i + 2], by = segs[4 * i + 3];
    double al = atan2(bx - ax, by - ay);
    double sg = (al > 0) - (al < 0);
    double nrm = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
    polar[2 * i] = sg * (ay * bx - by * ax) / nrm;
    polar[2 * i + 1] = fmod(-al + M_PI, M_PI);
}

// a11 + a12  sum_loss(phi, theta, lines) for k candidates (opencv_fit_color.py:14-26, 95-101)
__global__ void __launch_bounds__(128) k_vp_loss(const double *__restrict__ lines, int m,
                                                 const double *__restrict__ cand, double width, double height,
                                                 double *__restrict__ loss)
{
    const int c = blockIdx.x;
    const double phi = cand[2 * c], theta = cand[2 * c + 1];
    const double tphi = tan(phi), sphi = sin(phi), tth = tan(theta);
    double px[3], py[3];
    px[0] = height / 2 * (1 / (tth * sphi)) + width / 2;
    py[0] = height / 2 * (1 / tphi) + height / 2;
    px[1] = height / 2 * 0.0 + width / 2;
    py[1] = height / 2 * (-tphi) + height / 2;
    px[2] = height / 2 * (-tth / sphi) + width / 2;
    py[2] = height / 2 * (1 / tphi) + height / 2;
    double acc = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double rho = lines[2 * i], ph = lines[2 * i + 1];
        double s = sin(ph), co = cos(ph);
        double best = INFINITY;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double d = py[k] * s + px[k] * co - rho;
            d = d * d;
            // python min(): first minimum wins, NaN never replaces
            if (d < best) best = d;
            else if (k == 0) best = d;
        }
        acc += best;
    }
    __shared__ double red[4];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) loss[c] = (red[0] + red[1] + red[2] + red[3]) / (double)m;
}

// Gaussian sphere: segment -> interpretation-plane normal n = K^-1 a x K^-1 b; pair (i<j) ->
// direction v = n_i x n_j folded to v_z >= 0; bin by azimuth atan2(v_y, v_x) and elevation asin(v_z).
__global__ void k_sphere_hist(const double *__restrict__ segs, int m, double width, double height,
                              int32_t *__restrict__ hist, int n_az, int n_el)
{
    const int i = blockIdx.x;
    const double f = height / 2, cx = width / 2, cy = height / 2;
    auto normal = [&](int s, double &nx, double &ny, double &nz) {
        double ax = (segs[4 * s] - cx) / f, ay = (segs[4 * s + 1] - cy) / f;
        double bx = (segs[4 * s + 2] - cx) / f, by = (segs[4 * s + 3] - cy) / f;
        nx = ay - by;
        ny = bx - ax;
        nz = ax * by - ay * bx;
        double l = sqrt(nx * nx + ny * ny + nz * nz);
        if (l > 0) { nx /= l; ny /= l; nz /= l; }
    };
    double ix, iy, iz;
    normal(i, ix, iy, iz);
    for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
        double jx, jy, jz;
        normal(j, jx, jy, jz);
        double vx = iy * jz - iz * jy, vy = iz * jx - ix * jz, vz = ix * jy - iy * jx;
        double l = sqrt(vx * vx + vy * vy + vz * vz);
        if (!(l > 1e-9)) continue;  // (nearly) the same interpretation plane
        vx /= l; vy /= l; vz /= l;
        if (vz < 0) { vx = -vx; vy = -vy; vz = -vz; }
        double az = atan2(vy, vx);  // [-pi, pi]
        double el = asin(fmin(1.0, vz));  // [0, pi/2]
        int ba = (int)floor((az + M_PI) / (2 * M_PI) * n_az);
        int be = (int)floor(el / (M_PI / 2) * n_el);
        ba = min(max(ba, 0), n_az - 1);
        be = min(max(be, 0), n_el - 1);
        atomicAdd(&hist[be * n_az + ba], 1);
    }
}

// a9  _getIntersections (opencv.py:171-189): ordered compaction over i<j
static __device__ __forceinline__ bool pair_ok(const float *__restrict__ lines, int i, int j)
{
    float r1 = lines[2 * i], t1 = lines[2 * i + 1], r2 = lines[2 * j], t2 = lines[2 * j + 1];
    double dr = (double)fabsf(__fsub_rn(r1, r2)), dt = (double)fabsf(__fsub_rn(t1, t2));
    const double d10 = 10.0 * (M_PI / 180.0), d80 = 80.0 * (M_PI / 180.0);
    if ((dr < 40.0 && dt < d10) || dt > d80) return false;
    // np.linalg.solve raises on an exactly singular matrix (identical angles): skipped by the reference
    double c1 = cos((double)t1), s1 = sin((double)t1), c2 = cos((double)t2), s2 = sin((double)t2);
    return __dsub_rn(__dmul_rn(c1, s2), __dmul_rn(c2, s1)) != 0.0;  // unfused: identical angles give exactly 0
}
__global__ void k_pair_count(const float *__restrict__ lines, int m, int *__restrict__ rowcnt)
{
    int i = blockIdx.x;
    int c = 0;
    for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x) c += pair_ok(lines, i, j);
    __shared__ int red[32];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(~0u, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += red[k];
        rowcnt[i] = t;
    }
}
__global__ void k_pair_scan(int *rowcnt, int m, int32_t *count)
{
    // single-thread exclusive scan (m <= tens of thousands)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < m; i++) {
            int c = rowcnt[i];
            rowcnt[i] = acc;
            acc += c;
        }
        *count = acc;
    }
}
__global__ void __launch_bounds__(32) k_pair_write(const float *__restrict__ lines, int m,
                                                   const int *__restrict__ rowoff, double *__restrict__ out, int max_out)
{
    const int i = blockIdx.x, lane = threadIdx.x;
    int base = rowoff[i];
    for (int j0 = i + 1; j0 < m; j0 += 32) {
        int j = j0 + lane;
        bool ok = j < m && pair_ok(lines, i, j);
        u32 b = __ballot_sync(~0u, ok);
        if (ok) {
            int pos = base + __popc(b & ((1u << lane) - 1));
            if (pos < max_out) {
                double r1 = lines[2 * i], t1 = lines[2 * i + 1], r2 = lines[2 * j], t2 = lines[2 * j + 1];
                double c1 = cos(t1), s1 = sin(t1), c2 = cos(t2), s2 = sin(t2);
                double det = c1 * s2 - c2 * s1;
                out[2 * pos] = (r1 * s2 - r2 * s1) / det;
                out[2 * pos + 1] = (c1 * r2 - c2 * r1) / det;
            }
        }
        base += __popc(b);
    }
}

static int vp_common(le_ctx *ctx)
{
    if (!ctx) return LE_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return LE_OK;
}

extern "C" int le_segments_to_polar(le_ctx *ctx, const double *segs, int m, double *polar, void *stream)
{
    int rc = vp_common(ctx);
    if (rc) return rc;
    LE_REQUIRE(ctx, segs && polar && m > 0, "bad arguments");
    k_seg_polar<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(segs, m, polar);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_vp_loss(le_ctx *ctx, const double *lines, int m, const double *cand, int k, double width,
                          double height, double *loss, void *stream)
{
    int rc = vp_common(ctx);
    if (rc) return rc;
    LE_REQUIRE(ctx, lines && cand && loss && m > 0 && k > 0, "bad arguments");
    k_vp_loss<<<k, 128, 0, (cudaStream_t)stream>>>(lines, m, cand, width, height, loss);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_vp_sphere_hist(le_ctx *ctx, const double *segs, int m, double width, double height, int32_t *hist,
                                 int n_az, int n_el, void *stream)
{
    int rc = vp_common(ctx);
    if (rc) return rc;
    LE_REQUIRE(ctx, segs && hist && m >= 0 && n_az > 0 && n_el > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    LE_CUDA(ctx, cudaMemsetAsync(hist, 0, sizeof(int32_t) * (size_t)n_az * n_el, st));
    if (m > 1) {
        k_sphere_hist<<<m - 1, 128, 0, st>>>(segs, m, width, height, hist, n_az, n_el);
        LE_LAUNCH_CHECK(ctx);
    }
    return LE_OK;
}

extern "C" int le_pair_intersections(le_ctx *ctx, const float *lines, int m, double *out, int32_t *count, int max_out,
                                     void *stream)
{
    int rc = vp_common(ctx);
    if (rc) return rc;
    LE_REQUIRE(ctx, lines && out && count && m >= 0 && max_out > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (m < 2) {
        LE_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int32_t), st));
        return LE_OK;
    }
    rc = le_ensure(ctx, &ctx->misc, sizeof(int) * (size_t)m + 256);
    if (rc) return rc;
    int *rowcnt = (int *)((char *)ctx->misc.p + 256);
    k_pair_count<<<m, 128, 0, st>>>(lines, m, rowcnt);
    LE_LAUNCH_CHECK(ctx);
    k_pair_scan<<<1, 32, 0, st>>>(rowcnt, m, count);
    LE_LAUNCH_CHECK(ctx);
    k_pair_write<<<m, 32, 0, st>>>(lines, m, rowcnt, out, max_out);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
