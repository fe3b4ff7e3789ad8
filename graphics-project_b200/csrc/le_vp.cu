// le_vp.cu -- vanishing-point stage (a9-a12): polar conversion, batched loss, Gaussian-sphere
// histogram of pairwise interpretation-plane intersections, (rho,theta) pair intersections.
// All f64 (the reference computes in numpy float64); inputs are a few hundred segments, so these
// kernels are latency-, not bandwidth-bound; they exist so the whole fit runs without leaving the GPU.
#include "le_common.cuh"

// a10  edges_to_polar_lines (opencv_renewed.py:16-17); one body for the single-list and the batched kernel
static __device__ __forceinline__ void seg_polar(double ax, double ay, double bx, double by, double &rho, double &phi)
{
    double al = atan2(bx - ax, by - ay);
    double sg = (al > 0) - (al < 0);
    double nrm = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
    rho = sg * (ay * bx - by * ax) / nrm;
    phi = fmod(-al + M_PI, M_PI);
}
__global__ void k_seg_polar(const double *__restrict__ segs, int m, double *__restrict__ polar)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    seg_polar(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3], polar[2 * i], polar[2 * i + 1]);
}

// a11  get_focal_points (opencv_fit_color.py:95-101): the three vanishing points of (phi, theta) in pixels
static __device__ __forceinline__ void vp_points(double phi, double theta, double width, double height, double *px,
                                                 double *py)
{
    const double tphi = tan(phi), sphi = sin(phi), tth = tan(theta);
    px[0] = height / 2 * (1 / (tth * sphi)) + width / 2;
    py[0] = height / 2 * (1 / tphi) + height / 2;
    px[1] = height / 2 * 0.0 + width / 2;
    py[1] = height / 2 * (-tphi) + height / 2;
    px[2] = height / 2 * (-tth / sphi) + width / 2;
    py[2] = height / 2 * (1 / tphi) + height / 2;
}
// min_loss of one line (opencv_fit_color.py:14-23): s, co = sin / cos of the line's phi
static __device__ __forceinline__ double vp_line_loss(const double *px, const double *py, double rho, double s,
                                                      double co)
{
    double best = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double d = py[k] * s + px[k] * co - rho;
        d = d * d;
        // python min(): first minimum wins, NaN never replaces
        if (d < best) best = d;
        else if (k == 0) best = d;
    }
    return best;
}

// a11 + a12  sum_loss(phi, theta, lines) for k candidates (opencv_fit_color.py:14-26, 95-101)
__global__ void __launch_bounds__(128) k_vp_loss(const double *__restrict__ lines, int m,
                                                 const double *__restrict__ cand, double width, double height,
                                                 double *__restrict__ loss)
{
    const int c = blockIdx.x;
    double px[3], py[3];
    vp_points(cand[2 * c], cand[2 * c + 1], width, height, px, py);
    double acc = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double rho = lines[2 * i], ph = lines[2 * i + 1];
        acc += vp_line_loss(px, py, rho, sin(ph), cos(ph));
    }
    __shared__ double red[4];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(~0u, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) loss[c] = (red[0] + red[1] + red[2] + red[3]) / (double)m;
}

// Gaussian sphere: segment -> interpretation-plane normal n = K^-1 a x K^-1 b; pair (i<j) ->
// direction v = n_i x n_j folded to v_z >= 0; bin by azimuth atan2(v_y, v_x) and elevation asin(v_z).
static __device__ __forceinline__ void sphere_normal(double sax, double say, double sbx, double sby, double cx,
                                                     double cy, double f, double &nx, double &ny, double &nz)
{
    double ax = (sax - cx) / f, ay = (say - cy) / f;
    double bx = (sbx - cx) / f, by = (sby - cy) / f;
    nx = ay - by;
    ny = bx - ax;
    nz = ax * by - ay * bx;
    double l = sqrt(nx * nx + ny * ny + nz * nz);
    if (l > 0) { nx /= l; ny /= l; nz /= l; }
}
static __device__ __forceinline__ void sphere_vote(double ix, double iy, double iz, double jx, double jy, double jz,
                                                   int32_t *__restrict__ hist, int n_az, int n_el)
{
    double vx = iy * jz - iz * jy, vy = iz * jx - ix * jz, vz = ix * jy - iy * jx;
    double l = sqrt(vx * vx + vy * vy + vz * vz);
    if (!(l > 1e-9)) return;  // (nearly) the same interpretation plane
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
__global__ void k_sphere_hist(const double *__restrict__ segs, int m, double width, double height,
                              int32_t *__restrict__ hist, int n_az, int n_el)
{
    const int i = blockIdx.x;
    const double f = height / 2, cx = width / 2, cy = height / 2;
    double ix, iy, iz;
    sphere_normal(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3], cx, cy, f, ix, iy, iz);
    for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
        double jx, jy, jz;
        sphere_normal(segs[4 * j], segs[4 * j + 1], segs[4 * j + 2], segs[4 * j + 3], cx, cy, f, jx, jy, jz);
        sphere_vote(ix, iy, iz, jx, jy, jz, hist, n_az, n_el);
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

// ---------------------------------------------------------------------------------------------------------------
// Batched `lines` pipeline (BASELINE config 5; a14 'lsd' branch, opencv_fit_color.py:85-93, for n frames at once):
// LSD segments -> length filter + polar lines -> Gaussian-sphere histogram -> (phi, theta) fit, one CTA per frame
// and no host round trip between the steps.

// the list comprehension of opencv_fit_color.py:88 / edges_to_polar_lines: ordered compaction of the segments
// longer than min_len (the reference tests np.linalg.norm(b - a) on float32 endpoints), polar form in f64
__global__ void __launch_bounds__(256) k_polar_batch(const float *__restrict__ segs, const int32_t *__restrict__ counts,
                                                     int max_segs, float min_len, double *__restrict__ polar,
                                                     int32_t *__restrict__ pcounts)
{
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = min(max(counts[f], 0), max_segs);
    const float4 *sg = (const float4 *)segs + (size_t)f * max_segs;
    double *out = polar + (size_t)f * max_segs * 2;
    __shared__ int s_w[8];
    int base = 0;
    for (int i0 = 0; i0 < m; i0 += 256) {
        const int i = i0 + tid;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (i < m) {
            s = sg[i];
            const float dx = __fsub_rn(s.z, s.x), dy = __fsub_rn(s.w, s.y);
            keep = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) > min_len;
        }
        const u32 b = __ballot_sync(~0u, keep);
        __syncthreads();  // s_w of the previous trip has been read
        if (lane == 0) s_w[warp] = __popc(b);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            off += (k < warp) ? s_w[k] : 0;
            tot += s_w[k];
        }
        if (keep) {
            const int pos = base + off + __popc(b & ((1u << lane) - 1));
            seg_polar((double)s.x, (double)s.y, (double)s.z, (double)s.w, out[2 * pos], out[2 * pos + 1]);
        }
        base += tot;
    }
    if (tid == 0) pcounts[f] = base;
}

// per line: (rho, sin phi, cos phi) for the fit and the interpretation-plane normal of the line drawn through the
// image (the segment p0 -+ 100 d that vp.sphere_seeds builds) for the histogram
__global__ void __launch_bounds__(256) k_vpb_prep(const double *__restrict__ polar, const int32_t *__restrict__ pcounts,
                                                  int max_lines, double width, double height, double *__restrict__ rsc,
                                                  double *__restrict__ nrm)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(pcounts[f], max_lines)) return;
    const size_t o = (size_t)f * max_lines + i;
    const double rho = polar[2 * o], phi = polar[2 * o + 1];
    const double s = sin(phi), c = cos(phi);
    rsc[3 * o] = rho;
    rsc[3 * o + 1] = s;
    rsc[3 * o + 2] = c;
    if (nrm) {
        const double p0x = rho * c, p0y = rho * s, dx = -s * 100.0, dy = c * 100.0;
        sphere_normal(p0x - dx, p0y - dy, p0x + dx, p0y + dy, width / 2, height / 2, height / 2, nrm[3 * o],
                      nrm[3 * o + 1], nrm[3 * o + 2]);
    }
}

__global__ void __launch_bounds__(128) k_sphere_hist_batch(const double *__restrict__ nrm,
                                                           const int32_t *__restrict__ pcounts, int max_lines,
                                                           int32_t *__restrict__ hist, int n_az, int n_el)
{
    const int f = blockIdx.y;
    const int m = min(pcounts[f], max_lines);
    const double *nf = nrm + (size_t)f * max_lines * 3;
    int32_t *hf = hist + (size_t)f * n_az * n_el;
    for (int i = blockIdx.x; i < m - 1; i += gridDim.x) {
        const double ix = nf[3 * i], iy = nf[3 * i + 1], iz = nf[3 * i + 2];
        for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x)
            sphere_vote(ix, iy, iz, nf[3 * j], nf[3 * j + 1], nf[3 * j + 2], hf, n_az, n_el);
    }
}

// regress_lines (opencv_fit_color.py:28-47) as a deterministic search, one CTA per frame: thread = candidate,
// the frame's lines are broadcast from shared memory; block arg-min with the lower candidate index winning ties
#define VPF_THREADS 512
#define VPF_CAP 1536  // lines kept in shared memory (36 KB); longer lists are read from the scratch copy
#define VPF_MAX_TOP 16

struct vpf_best {
    double loss;
    int idx;
};
static __device__ __forceinline__ vpf_best vpf_min(vpf_best a, vpf_best b)
{
    return (b.loss < a.loss || (b.loss == a.loss && b.idx < a.idx)) ? b : a;
}
static __device__ vpf_best vpf_block_min(vpf_best v, vpf_best *s_red)
{
    for (int o = 16; o; o >>= 1) {
        vpf_best w;
        w.loss = __shfl_xor_sync(~0u, v.loss, o);
        w.idx = __shfl_xor_sync(~0u, v.idx, o);
        v = vpf_min(v, w);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    vpf_best r = s_red[0];
    for (int k = 1; k < VPF_THREADS / 32; k++) r = vpf_min(r, s_red[k]);
    return r;
}
// numpy's  lo + (arange(n) + 0.5) / n * (hi - lo)  in its operation order, never fused
static __device__ __forceinline__ double vpf_grid(double lo, double hi, int i, int n)
{
    return __dadd_rn(lo, __dmul_rn(__ddiv_rn((double)i + 0.5, (double)n), __dsub_rn(hi, lo)));
}

__global__ void __launch_bounds__(VPF_THREADS) k_vp_fit(const double *__restrict__ rsc_g,
                                                        const int32_t *__restrict__ pcounts, int max_lines,
                                                        double width, double height, int n_phi, int n_th, int n_ref,
                                                        int rounds, double ref_area,
                                                        const int32_t *__restrict__ hist, int n_az, int n_el, int top,
                                                        double *__restrict__ fit)
{
    __shared__ double s_rsc[3 * VPF_CAP];
    __shared__ vpf_best s_red[VPF_THREADS / 32];
    __shared__ unsigned long long s_key[VPF_THREADS / 32];
    __shared__ int s_top[VPF_MAX_TOP];
    __shared__ double s_seed[2 * VPF_MAX_TOP][2];
    __shared__ int s_nseed;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int m = min(pcounts[f], max_lines);
    double *out = fit + 3 * (size_t)f;
    if (m <= 0) {  // the reference's initial best; regress_lines itself divides by zero here
        if (tid == 0) { out[0] = 0.0; out[1] = 0.0; out[2] = 100000.0; }
        return;
    }
    const double *rg = rsc_g + (size_t)f * max_lines * 3;
    if (m <= VPF_CAP)
        for (int i = tid; i < 3 * m; i += VPF_THREADS) s_rsc[i] = rg[i];
    const double *L = (m <= VPF_CAP) ? (const double *)s_rsc : rg;
    if (tid == 0) s_nseed = 0;

    // seeds: the `top` highest bins (count desc, index asc) read as the 'x' or 'z' vanishing point
    if (hist && m >= 2 && top > 0) {
        const int nb = n_az * n_el;
        const int32_t *hf = hist + (size_t)f * nb;
        unsigned long long prev = ~0ull;
        for (int r = 0; r < top; r++) {
            unsigned long long best = 0;
            for (int b = tid; b < nb; b += VPF_THREADS) {
                unsigned long long key = ((unsigned long long)(u32)hf[b] << 32) | (u32)(0xffffffffu - (u32)b);
                if (key < prev && key > best) best = key;
            }
            for (int o = 16; o; o >>= 1) {
                unsigned long long w = __shfl_xor_sync(~0u, best, o);
                best = w > best ? w : best;
            }
            __syncthreads();
            if ((tid & 31) == 0) s_key[tid >> 5] = best;
            __syncthreads();
            best = s_key[0];
            for (int k = 1; k < VPF_THREADS / 32; k++) best = s_key[k] > best ? s_key[k] : best;
            if (tid == 0) s_top[r] = best ? (int)(0xffffffffu - (u32)(best & 0xffffffffu)) : -1;
            prev = best ? best : 0;
        }
        __syncthreads();
        if (tid == 0) {
            int ns = 0;
            for (int r = 0; r < top; r++) {
                const int b = s_top[r];
                if (b < 0) continue;
                const int be = b / n_az, ba = b % n_az;
                const double az = ((double)ba + 0.5) / n_az * 2 * M_PI - M_PI;
                const double el = ((double)be + 0.5) / n_el * (M_PI / 2);
                const double vx = cos(el) * cos(az), vy = cos(el) * sin(az), vz = sin(el);
                if (vz < 1e-6) continue;
                const double x = vx / vz, y = vy / vz;
                double ph = atan2(1.0, y);  // 1 / tan(phi) = y
                if (ph >= M_PI) ph -= M_PI;
                const double sp = sin(ph);
                if (fabs(x) < 1e-9 || fabs(sp) < 1e-9) continue;
                const double th[2] = {atan(1.0 / (x * sp)), atan(-x * sp)};  // x = 1/(tan t sin p);  x = -tan t / sin p
                for (int k = 0; k < 2; k++)
                    if (th[k] >= 0 && th[k] < M_PI / 2) {
                        s_seed[ns][0] = ph;
                        s_seed[ns][1] = th[k];
                        ns++;
                    }
            }
            s_nseed = ns;
        }
    }
    __syncthreads();

    auto eval = [&](double phi, double theta) -> double {
        double px[3], py[3];
        vp_points(phi, theta, width, height, px, py);
        double acc = 0;
        if (m <= VPF_CAP) {
            for (int i = 0; i < m; i++) acc += vp_line_loss(px, py, L[3 * i], L[3 * i + 1], L[3 * i + 2]);
        } else {
            for (int i = 0; i < m; i++) acc += vp_line_loss(px, py, __ldg(L + 3 * i), __ldg(L + 3 * i + 1), __ldg(L + 3 * i + 2));
        }
        acc /= (double)m;
        return (acc != acc) ? INFINITY : acc;  // np.where(np.isnan(ls), inf, ls)
    };

    double best_loss = 100000.0, best_phi = 0.0, best_th = 0.0;
    // phase 1: the grid over [0, pi) x [0, pi/2) (candidate c = i_phi * n_th + i_th) followed by the seeds
    {
        const int ngrid = n_phi * n_th, ncand = ngrid + s_nseed;
        vpf_best mine;
        mine.loss = INFINITY;
        mine.idx = 0x7fffffff;
        for (int c = tid; c < ncand; c += VPF_THREADS) {
            double phi, th;
            if (c < ngrid) {
                phi = vpf_grid(0.0, M_PI, c / n_th, n_phi);
                th = vpf_grid(0.0, M_PI / 2, c % n_th, n_th);
            } else {
                phi = s_seed[c - ngrid][0];
                th = s_seed[c - ngrid][1];
            }
            vpf_best v;
            v.loss = eval(phi, th);
            v.idx = c;
            mine = vpf_min(mine, v);
        }
        const vpf_best r = vpf_block_min(mine, s_red);
        if (r.idx != 0x7fffffff && r.loss < best_loss) {
            best_loss = r.loss;
            if (r.idx < ngrid) {
                best_phi = vpf_grid(0.0, M_PI, r.idx / n_th, n_phi);
                best_th = vpf_grid(0.0, M_PI / 2, r.idx % n_th, n_th);
            } else {
                best_phi = s_seed[r.idx - ngrid][0];
                best_th = s_seed[r.idx - ngrid][1];
            }
        }
    }
    // phase 2: shrinking n_ref x n_ref boxes around the best so far (the reference makes one pass of side ref_area)
    double side = ref_area;
    for (int round = 0; round < rounds; round++) {
        const double plo = __dsub_rn(best_phi, __ddiv_rn(side, 2.0)), phi_hi = __dadd_rn(best_phi, __ddiv_rn(side, 2.0));
        const double tlo = __dsub_rn(best_th, __ddiv_rn(side, 2.0)), thi = __dadd_rn(best_th, __ddiv_rn(side, 2.0));
        vpf_best mine;
        mine.loss = INFINITY;
        mine.idx = 0x7fffffff;
        for (int c = tid; c < n_ref * n_ref; c += VPF_THREADS) {
            vpf_best v;
            v.loss = eval(vpf_grid(plo, phi_hi, c / n_ref, n_ref), vpf_grid(tlo, thi, c % n_ref, n_ref));
            v.idx = c;
            mine = vpf_min(mine, v);
        }
        const vpf_best r = vpf_block_min(mine, s_red);
        if (r.idx != 0x7fffffff && r.loss < best_loss) {
            best_loss = r.loss;
            best_phi = vpf_grid(plo, phi_hi, r.idx / n_ref, n_ref);
            best_th = vpf_grid(tlo, thi, r.idx % n_ref, n_ref);
        }
        side = __dmul_rn(side, __ddiv_rn(4.0, (double)n_ref));
    }
    if (tid == 0) { out[0] = best_phi; out[1] = best_th; out[2] = best_loss; }
}

static int vp_common(le_ctx *ctx)
{
    if (!ctx) return LE_ERR_INVALID;
    return LE_OK;
}

extern "C" int le_segments_to_polar(le_ctx *ctx, const double *segs, int m, double *polar, void *stream)
{
    LE_ON_DEVICE(ctx);
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
    LE_ON_DEVICE(ctx);
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
    LE_ON_DEVICE(ctx);
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
    LE_ON_DEVICE(ctx);
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

extern "C" int le_polar_batch(le_ctx *ctx, const float *segs, const int32_t *counts, int n, int max_segs,
                              double min_len, double *polar, int32_t *pcounts, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = vp_common(ctx);
    if (rc) return rc;
    LE_REQUIRE(ctx, segs && counts && polar && pcounts && n > 0 && max_segs > 0, "bad arguments");
    LE_REQUIRE(ctx, ((size_t)segs & 15) == 0, "segs must be 16-byte aligned");
    k_polar_batch<<<n, 256, 0, (cudaStream_t)stream>>>(segs, counts, max_segs, (float)min_len, polar, pcounts);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_vp_fit_batch(le_ctx *ctx, const double *polar, const int32_t *pcounts, int n, int max_lines,
                               double width, double height, int n_phi, int n_th, int n_ref, int rounds,
                               double ref_area, int32_t *hist, int n_az, int n_el, int top, double *fit, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = vp_common(ctx);
    if (rc) return rc;
    LE_REQUIRE(ctx, polar && pcounts && fit && n > 0 && max_lines > 0, "bad arguments");
    LE_REQUIRE(ctx, n_phi > 0 && n_th > 0 && n_ref > 0 && rounds >= 0 && (long long)n_phi * n_th < (1ll << 30) &&
                        n_ref < 32768, "bad search grid");
    LE_REQUIRE(ctx, !hist || (n_az > 0 && n_el > 0 && top >= 0 && top <= VPF_MAX_TOP), "bad histogram arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = sizeof(double) * 3 * (size_t)n * max_lines;
    rc = le_ensure(ctx, &ctx->misc, 2 * per + 256);
    if (rc) return rc;
    double *rsc = (double *)ctx->misc.p, *nrm = hist ? (double *)((char *)ctx->misc.p + per) : nullptr;
    k_vpb_prep<<<dim3((max_lines + 255) / 256, n), 256, 0, st>>>(polar, pcounts, max_lines, width, height, rsc, nrm);
    LE_LAUNCH_CHECK(ctx);
    if (hist) {
        LE_CUDA(ctx, cudaMemsetAsync(hist, 0, sizeof(int32_t) * (size_t)n * n_az * n_el, st));
        k_sphere_hist_batch<<<dim3(64, n), 128, 0, st>>>(nrm, pcounts, max_lines, hist, n_az, n_el);
        LE_LAUNCH_CHECK(ctx);
    }
    k_vp_fit<<<n, VPF_THREADS, 0, st>>>(rsc, pcounts, max_lines, width, height, n_phi, n_th, n_ref, rounds, ref_area,
                                        hist, n_az, n_el, top, fit);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
