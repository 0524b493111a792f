/*
 * ORACLE (test infrastructure, NOT product code) -- CPU restatement of the four DIB-R
 * rasterizer kernels that Self6D++ calls through
 *   kaolin.graphics.dib_renderer.cuda.rasterizer.{forward,backward}
 * (reference call sites: lib/dr_utils/dib_renderer_x/rasterizer/rasterizer.py:152-172 and
 * :249-269).  The kernel sources are NOT in /root/reference: they live in the un-vendored
 * third-party wheel kaolin v0.1 (pinned only by .gitignore:61 "external/kaolin-0.1"),
 * files kaolin/graphics/dib_renderer/cuda/rasterizer_cuda.cu and rasterizer_cuda_back.cu.
 * What follows restates their published algorithm (DIB-R, Chen et al., NeurIPS 2019) as
 * recalled; SURVEY.md section 8(a) rows a6,a7,a9,a10 list the same steps.
 *
 * PARITY UNPINNED: the reference ships no test, fixture or golden vector for this path,
 * so this oracle is pinned only by the known-answer tests in tests/test_oracle_kat.py
 * (hand algebra), by finite differences of its own float64 forward, and by an independent
 * float64 PyTorch transcription (oracle/torch_oracle.py).
 *
 * This body is included twice: REAL=float (kernel operation order, used for the bit-exact
 * face-index / visibility comparison) and REAL=double (used for the 1e-5 tolerance checks).
 *
 * Frozen choices (the items SURVEY.md marks VERIFY):
 *   - eps is the double literal 1e-15, so in the float build "k3 + eps" and the division
 *     happen in double and are rounded once to float (C usual arithmetic conversions).
 *   - bbox tests are half-open: xmin <= x0 < xmax, ymin <= y0 < ymax.
 *   - K1 culls faces with normalz < 0; K2 (soft silhouette) does NOT cull back faces.
 *   - argmin over the 6 squared distances keeps the FIRST minimum (strict '>').
 *   - nvcc's default -fmad=true contraction is modelled with explicit FMA() in the order
 *     LLVM's left-to-right fold produces: x*y - z*w -> fma(x,y,-(z*w));
 *     a*b + c*d + e*f -> fma(e,f,fma(a,b,c*d)).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

static inline REAL FN(pix_x)(int w, int width, int multiplier) {
    /* "1.0 * multiplier / width * (2 * wididx + 1 - width)": evaluated in double, rounded once */
    return (REAL)(1.0 * multiplier / width * (2 * w + 1 - width));
}
static inline REAL FN(pix_y)(int h, int height, int multiplier) {
    return (REAL)(1.0 * multiplier / height * (height - 2 * h - 1));
}

/* K1: dr_cuda_forward_render_batch */
void FN(dibr_oracle_forward_render)(
    const REAL *points3d_bxfx9, const REAL *points2d_bxfx6, const REAL *pointsdirect_bxfx1,
    const REAL *pointsbbox_bxfx4, const REAL *features_bxfx3d,
    REAL *imidx_bxhxwx1, REAL *imdep_bxhxwx1, REAL *imwei_bxhxwx3, REAL *im_bxhxwxd,
    int bnum, int height, int width, int fnum, int dnum, int multiplier)
{
    const long npix = (long)bnum * height * width;
#pragma omp parallel for schedule(dynamic, 256)
    for (long pix = 0; pix < npix; pix++) {
        const int wididx = (int)(pix % width);
        const int heiidx = (int)((pix / width) % height);
        const int bidx = (int)(pix / ((long)width * height));
        const long totalidx1 = pix, totalidx3 = pix * 3, totalidxd = pix * dnum;
        const REAL x0 = FN(pix_x)(wididx, width, multiplier);
        const REAL y0 = FN(pix_y)(heiidx, height, multiplier);

        for (int f = 0; f < fnum; f++) {
            const long shift1 = (long)bidx * fnum + f;
            const long shift4 = shift1 * 4, shift6 = shift1 * 6, shift9 = shift1 * 9;
            const long shift3d = shift1 * 3 * dnum;

            if (pointsdirect_bxfx1[shift1] < 0) continue;          /* back face */

            const REAL xmin = pointsbbox_bxfx4[shift4 + 0], ymin = pointsbbox_bxfx4[shift4 + 1];
            const REAL xmax = pointsbbox_bxfx4[shift4 + 2], ymax = pointsbbox_bxfx4[shift4 + 3];
            if (x0 < xmin || x0 >= xmax || y0 < ymin || y0 >= ymax) continue;

            const REAL ax = points2d_bxfx6[shift6 + 0], ay = points2d_bxfx6[shift6 + 1];
            const REAL bx = points2d_bxfx6[shift6 + 2], by = points2d_bxfx6[shift6 + 3];
            const REAL cx = points2d_bxfx6[shift6 + 4], cy = points2d_bxfx6[shift6 + 5];

            const REAL m = bx - ax, p = by - ay;
            const REAL n = cx - ax, q = cy - ay;
            const REAL s = x0 - ax, t = y0 - ay;

            const REAL k1 = FMA(s, q, -(n * t));
            const REAL k2 = FMA(m, t, -(s * p));
            const REAL k3 = FMA(m, q, -(n * p));

            const REAL w1 = (REAL)((double)k1 / ((double)k3 + 1e-15));
            const REAL w2 = (REAL)((double)k2 / ((double)k3 + 1e-15));
            const REAL w0 = ((REAL)1 - w1) - w2;
            if (w0 < 0 || w1 < 0 || w2 < 0) continue;

            const REAL az = points3d_bxfx9[shift9 + 2];
            const REAL bz = points3d_bxfx9[shift9 + 5];
            const REAL cz = points3d_bxfx9[shift9 + 8];
            const REAL z0 = FMA(w2, cz, FMA(w0, az, w1 * bz));
            const REAL znow = imdep_bxhxwx1[totalidx1];
            if (z0 <= znow) continue;                               /* strict '>' : first face wins ties */

            imidx_bxhxwx1[totalidx1] = (REAL)(f + 1.0);
            imdep_bxhxwx1[totalidx1] = z0;
            imwei_bxhxwx3[totalidx3 + 0] = w0;
            imwei_bxhxwx3[totalidx3 + 1] = w1;
            imwei_bxhxwx3[totalidx3 + 2] = w2;
            for (int d = 0; d < dnum; d++) {
                const REAL r0 = features_bxfx3d[shift3d + d];
                const REAL r1 = features_bxfx3d[shift3d + dnum + d];
                const REAL r2 = features_bxfx3d[shift3d + dnum + dnum + d];
                im_bxhxwxd[totalidxd + d] = FMA(w2, r2, FMA(w0, r0, w1 * r1));
            }
        }
    }
}

/* squared distances of pixel (x0,y0) to the 3 edges and 3 vertices of one face; returns the
 * first-min index in *edgeid (0..2 edge i->i+1, 3..5 vertex i-3).  Shared by K2 and the test
 * helpers so the "case" rule is stated once. */
static inline REAL FN(face_min_dis)(const REAL *p6, REAL x0, REAL y0, int multiplier, int *edgeid)
{
    REAL pdis[6];
    for (int i = 0; i < 3; i++) {
        const REAL x1 = p6[i * 2 + 0], y1 = p6[i * 2 + 1];
        const REAL x2 = p6[((i + 1) % 3) * 2 + 0], y2 = p6[((i + 1) % 3) * 2 + 1];
        const REAL A = y2 - y1;
        const REAL B = x1 - x2;
        const REAL C = FMA(x2, y1, -(x1 * y2));
        const REAL up = FMA(A, x0, B * y0) + C;
        const REAL down = FMA(A, A, B * B);
        /* foot of the perpendicular */
        REAL x3 = FMA(-A, C, FMA(B * B, x0, -((A * B) * y0)));
        REAL y3 = FMA(-B, C, FMA(A * A, y0, -((A * B) * x0)));
        x3 = (REAL)((double)x3 / ((double)down + 1e-15));
        y3 = (REAL)((double)y3 / ((double)down + 1e-15));
        const REAL direct = FMA(x3 - x1, x3 - x2, (y3 - y1) * (y3 - y2));
        if (direct > 0) {
            pdis[i] = (REAL)(4 * multiplier * multiplier);          /* foot outside the segment */
        } else {
            pdis[i] = (REAL)((double)(up * up) / ((double)down + 1e-15));
        }
    }
    for (int i = 0; i < 3; i++) {
        const REAL x1 = p6[i * 2 + 0], y1 = p6[i * 2 + 1];
        pdis[i + 3] = FMA(x0 - x1, x0 - x1, (y0 - y1) * (y0 - y1));
    }
    int eid = 0;
    REAL dissquare = pdis[0];
    for (int i = 1; i < 6; i++) {
        if (dissquare > pdis[i]) { dissquare = pdis[i]; eid = i; }
    }
    *edgeid = eid;
    return dissquare;
}

/* K2: dr_cuda_forward_prob_batch */
void FN(dibr_oracle_forward_prob)(
    const REAL *points2d_bxfx6, const REAL *pointsbbox2_bxfx4, const REAL *pointsdep_bxfx1,
    const REAL *imidx_bxhxwx1,
    REAL *probface_bxhxwxk, REAL *probcase_bxhxwxk, REAL *probdis_bxhxwxk, REAL *probdep_bxhxwxk,
    REAL *improb_bxhxwx1,
    int bnum, int height, int width, int fnum, int knum, int multiplier, int sigmainv)
{
    const long npix = (long)bnum * height * width;
#pragma omp parallel for schedule(dynamic, 256)
    for (long pix = 0; pix < npix; pix++) {
        const int wididx = (int)(pix % width);
        const int heiidx = (int)((pix / width) % height);
        const int bidx = (int)(pix / ((long)width * height));
        const long totalidx1 = pix, totalidxk = pix * knum;

        const int fidxcover = (int)(imidx_bxhxwx1[totalidx1] + 0.5) - 1;
        if (fidxcover >= 0) { improb_bxhxwx1[totalidx1] = (REAL)1.0; continue; }

        const REAL x0 = FN(pix_x)(wididx, width, multiplier);
        const REAL y0 = FN(pix_y)(heiidx, height, multiplier);
        int kid = 0;
        for (int f = 0; f < fnum && kid < knum; f++) {
            const long shift1 = (long)bidx * fnum + f;
            const long shift4 = shift1 * 4, shift6 = shift1 * 6;
            const REAL xmin = pointsbbox2_bxfx4[shift4 + 0], ymin = pointsbbox2_bxfx4[shift4 + 1];
            const REAL xmax = pointsbbox2_bxfx4[shift4 + 2], ymax = pointsbbox2_bxfx4[shift4 + 3];
            if (x0 < xmin || x0 >= xmax || y0 < ymin || y0 >= ymax) continue;

            int edgeid;
            const REAL dissquare = FN(face_min_dis)(points2d_bxfx6 + shift6, x0, y0, multiplier, &edgeid);
            const REAL z = (REAL)sigmainv * dissquare / (REAL)multiplier / (REAL)multiplier;
            const REAL prob = (REAL)EXP(-z);

            probface_bxhxwxk[totalidxk + kid] = (REAL)(f + 1.0);
            probcase_bxhxwxk[totalidxk + kid] = (REAL)(edgeid + 1.0);
            probdis_bxhxwxk[totalidxk + kid] = prob;
            probdep_bxhxwxk[totalidxk + kid] = pointsdep_bxfx1[shift1];
            kid++;
        }
        REAL allprob = (REAL)1.0;
        for (int i = 0; i < kid; i++) {
            const REAL prob = probdis_bxhxwxk[totalidxk + i];
            allprob = (REAL)((double)allprob * (1.0 - (double)prob));
        }
        improb_bxhxwx1[totalidx1] = (REAL)(1.0 - (double)allprob);
    }
}

/* K3: dr_cuda_backward_color_batch.  The reference scatters with fp32
 * atomicAdd in undefined order; the oracle accumulates in ACC (double) in pixel order and
 * rounds once, which is the value any summation order converges to. */
void FN(dibr_oracle_backward_color)(
    const REAL *grad_im_bxhxwxd, const REAL *imidx_bxhxwx1, const REAL *imwei_bxhxwx3,
    const REAL *points2d_bxfx6, const REAL *features_bxfx3d,
    double *grad_points2d_bxfx6, double *grad_features_bxfx3d,
    int bnum, int height, int width, int fnum, int dnum, int multiplier)
{
    const long npix = (long)bnum * height * width;
    for (long pix = 0; pix < npix; pix++) {
        const int wididx = (int)(pix % width);
        const int heiidx = (int)((pix / width) % height);
        const int bidx = (int)(pix / ((long)width * height));
        const long totalidx1 = pix, totalidx3 = pix * 3, totalidxd = pix * dnum;
        const int fidxint = (int)(imidx_bxhxwx1[totalidx1] + 0.5) - 1;
        if (fidxint < 0) continue;

        const REAL x0 = FN(pix_x)(wididx, width, multiplier);
        const REAL y0 = FN(pix_y)(heiidx, height, multiplier);
        const long shift1 = (long)bidx * fnum + fidxint;
        const long shift6 = shift1 * 6, shift3d = shift1 * 3 * dnum;

        for (int i = 0; i < 3; i++) {
            const REAL w = imwei_bxhxwx3[totalidx3 + i];
            for (int d = 0; d < dnum; d++)
                grad_features_bxfx3d[shift3d + i * dnum + d] += (double)(grad_im_bxhxwxd[totalidxd + d] * w);
        }

        const REAL ax = points2d_bxfx6[shift6 + 0], ay = points2d_bxfx6[shift6 + 1];
        const REAL bx = points2d_bxfx6[shift6 + 2], by = points2d_bxfx6[shift6 + 3];
        const REAL cx = points2d_bxfx6[shift6 + 4], cy = points2d_bxfx6[shift6 + 5];
        const REAL m = bx - ax, p = by - ay, n = cx - ax, q = cy - ay, s = x0 - ax, t = y0 - ay;
        const REAL k1 = s * q - n * t, k2 = m * t - s * p, k3 = m * q - n * p;

        const REAL dk1dm = 0, dk1dn = -t, dk1dp = 0, dk1dq = s, dk1ds = q, dk1dt = -n;
        const REAL dk2dm = t, dk2dn = 0, dk2dp = -s, dk2dq = 0, dk2ds = -p, dk2dt = m;
        const REAL dk3dm = q, dk3dn = -p, dk3dp = -n, dk3dq = m, dk3ds = 0, dk3dt = 0;

        /* w1 = k1/k3, w2 = k2/k3; the common 1/k3^2 is applied in dldI below */
        const REAL dw1dm = dk1dm * k3 - dk3dm * k1, dw1dn = dk1dn * k3 - dk3dn * k1;
        const REAL dw1dp = dk1dp * k3 - dk3dp * k1, dw1dq = dk1dq * k3 - dk3dq * k1;
        const REAL dw1ds = dk1ds * k3 - dk3ds * k1, dw1dt = dk1dt * k3 - dk3dt * k1;
        const REAL dw2dm = dk2dm * k3 - dk3dm * k2, dw2dn = dk2dn * k3 - dk3dn * k2;
        const REAL dw2dp = dk2dp * k3 - dk3dp * k2, dw2dq = dk2dq * k3 - dk3dq * k2;
        const REAL dw2ds = dk2ds * k3 - dk3ds * k2, dw2dt = dk2dt * k3 - dk3dt * k2;

        const REAL dw1[6] = { -(dw1dm + dw1dn + dw1ds), -(dw1dp + dw1dq + dw1dt), dw1dm, dw1dp, dw1dn, dw1dq };
        const REAL dw2[6] = { -(dw2dm + dw2dn + dw2ds), -(dw2dp + dw2dq + dw2dt), dw2dm, dw2dp, dw2dn, dw2dq };

        for (int d = 0; d < dnum; d++) {
            const REAL c0 = features_bxfx3d[shift3d + d];
            const REAL c1 = features_bxfx3d[shift3d + dnum + d];
            const REAL c2 = features_bxfx3d[shift3d + dnum + dnum + d];
            const REAL dldI = (REAL)((double)((REAL)multiplier * grad_im_bxhxwxd[totalidxd + d]) / ((double)(k3 * k3) + 1e-15));
            for (int j = 0; j < 6; j++) {
                const REAL dIdp = (c1 - c0) * dw1[j] + (c2 - c0) * dw2[j];
                grad_points2d_bxfx6[shift6 + j] += (double)(dldI * dIdp);
            }
        }
    }
}

/* K4: dr_cuda_backward_prob_batch */
void FN(dibr_oracle_backward_prob)(
    const REAL *grad_improb_bxhxwx1, const REAL *improb_bxhxwx1, const REAL *imidx_bxhxwx1,
    const REAL *probface_bxhxwxk, const REAL *probcase_bxhxwxk, const REAL *probdis_bxhxwxk,
    const REAL *points2d_bxfx6, double *grad_points2dprob_bxfx6,
    int bnum, int height, int width, int fnum, int knum, int multiplier, int sigmainv)
{
    const long npix = (long)bnum * height * width;
    for (long pix = 0; pix < npix; pix++) {
        const int wididx = (int)(pix % width);
        const int heiidx = (int)((pix / width) % height);
        const int bidx = (int)(pix / ((long)width * height));
        const long totalidx1 = pix, totalidxk = pix * knum;
        const int fidxcover = (int)(imidx_bxhxwx1[totalidx1] + 0.5) - 1;
        if (fidxcover >= 0) continue;

        const REAL x0 = FN(pix_x)(wididx, width, multiplier);
        const REAL y0 = FN(pix_y)(heiidx, height, multiplier);
        const REAL dLdp = grad_improb_bxhxwx1[totalidx1];
        const REAL allprob = improb_bxhxwx1[totalidx1];

        for (int kid = 0; kid < knum; kid++) {
            const int fidxint = (int)(probface_bxhxwxk[totalidxk + kid] + 0.5) - 1;
            if (fidxint < 0) break;
            const long shift1 = (long)bidx * fnum + fidxint;
            const long shift6 = shift1 * 6;
            const REAL prob = probdis_bxhxwxk[totalidxk + kid];
            const REAL dLdz = (REAL)(-1.0 * sigmainv * dLdp * (1.0 - allprob) / (1.0 - prob + 1e-15) * prob);
            const int edgeid = (int)(probcase_bxhxwxk[totalidxk + kid] + 0.5) - 1;

            if (edgeid >= 3) {
                const long pshift = shift6 + (edgeid - 3) * 2;
                const REAL x1 = points2d_bxfx6[pshift + 0], y1 = points2d_bxfx6[pshift + 1];
                const REAL dLdx1 = dLdz * 2 * (x1 - x0);
                const REAL dLdy1 = dLdz * 2 * (y1 - y0);
                grad_points2dprob_bxfx6[pshift + 0] += (double)(dLdx1 / multiplier);
                grad_points2dprob_bxfx6[pshift + 1] += (double)(dLdy1 / multiplier);
            } else {
                const long pshift = shift6 + edgeid * 2;
                const long pshift2 = shift6 + ((edgeid + 1) % 3) * 2;
                const REAL x1 = points2d_bxfx6[pshift + 0], y1 = points2d_bxfx6[pshift + 1];
                const REAL x2 = points2d_bxfx6[pshift2 + 0], y2 = points2d_bxfx6[pshift2 + 1];
                const REAL A = y2 - y1, B = x1 - x2, C = x2 * y1 - x1 * y2;
                const REAL up = A * x0 + B * y0 + C;
                const REAL down = A * A + B * B;
                const REAL dissquare = (REAL)((double)(up * up) / ((double)down + 1e-15));
                const REAL dzdA = (REAL)((double)(2 * (x0 * up - dissquare * A)) / ((double)down + 1e-15));
                const REAL dzdB = (REAL)((double)(2 * (y0 * up - dissquare * B)) / ((double)down + 1e-15));
                const REAL dzdC = (REAL)((double)(2 * up) / ((double)down + 1e-15));
                const REAL dLdx1 = dLdz * (dzdB - y2 * dzdC);
                const REAL dLdy1 = dLdz * (x2 * dzdC - dzdA);
                const REAL dLdx2 = dLdz * (y1 * dzdC - dzdB);
                const REAL dLdy2 = dLdz * (dzdA - x1 * dzdC);
                grad_points2dprob_bxfx6[pshift + 0] += (double)(dLdx1 / multiplier);
                grad_points2dprob_bxfx6[pshift + 1] += (double)(dLdy1 / multiplier);
                grad_points2dprob_bxfx6[pshift2 + 0] += (double)(dLdx2 / multiplier);
                grad_points2dprob_bxfx6[pshift2 + 1] += (double)(dLdy2 / multiplier);
            }
        }
    }
}

/* Work counters for the FP32-side roofline: N_cov = sum over front faces of
 * the pixels whose centre lies in the face bbox; N_soft = sum over faces of UNCOVERED pixels in
 * the expanded bbox (before the K cap). */
void FN(dibr_oracle_work_counts)(
    const REAL *pointsdirect_bxfx1, const REAL *pointsbbox_bxfx4, const REAL *pointsbbox2_bxfx4,
    const REAL *imidx_bxhxwx1, int bnum, int height, int width, int fnum, int multiplier,
    long long *n_cov, long long *n_soft)
{
    long long cov = 0, soft = 0;
    const long npix = (long)bnum * height * width;
#pragma omp parallel for reduction(+ : cov, soft) schedule(dynamic, 256)
    for (long pix = 0; pix < npix; pix++) {
        const int wididx = (int)(pix % width);
        const int heiidx = (int)((pix / width) % height);
        const int bidx = (int)(pix / ((long)width * height));
        const REAL x0 = FN(pix_x)(wididx, width, multiplier);
        const REAL y0 = FN(pix_y)(heiidx, height, multiplier);
        const int covered = ((int)(imidx_bxhxwx1[pix] + 0.5) - 1) >= 0;
        for (int f = 0; f < fnum; f++) {
            const long shift1 = (long)bidx * fnum + f, shift4 = shift1 * 4;
            if (pointsdirect_bxfx1[shift1] >= 0) {
                const REAL *bb = pointsbbox_bxfx4 + shift4;
                if (!(x0 < bb[0] || x0 >= bb[2] || y0 < bb[1] || y0 >= bb[3])) cov++;
            }
            if (!covered) {
                const REAL *bb = pointsbbox2_bxfx4 + shift4;
                if (!(x0 < bb[0] || x0 >= bb[2] || y0 < bb[1] || y0 >= bb[3])) soft++;
            }
        }
    }
    *n_cov = cov;
    *n_soft = soft;
}

/* Restatement of the vertex shader + face set-up in a FIXED fp32 operation order, used as the
 * spec of the fused B200 path (reference: renderer/vertex_shaders/perpsective.py:71-111 --
 * view transform, 4x4 projection, divide by w, per-face gather, un-normalised face normal --
 * and rasterizer/rasterizer.py:36-70 prepare_tfpoints).  The reference does the view transform
 * with torch.matmul whose accumulation order is unspecified; the order fixed here is
 *   pc_j = fma(Rc[j][2], d2, fma(Rc[j][1], d1, Rc[j][0]*d0)),  d = v - cam_pos
 *   clip_c = fma(pc2, P[2][c], fma(pc1, P[1][c], pc0*P[0][c])) + P[3][c]
 * proj is the row-major 4x4 used as  [pc,1] @ proj.
 * Outputs are in the reference's operator-seam layout (points3d_fx9, points2d_fx6 un-scaled,
 * normalz_fx1, normal_fx3 un-normalised). */
void FN(dibr_oracle_project)(
    const REAL *verts_px3, int pnum, const int *faces_fx3, int fnum,
    const REAL *cam_rot_3x3, const REAL *cam_pos_3, const REAL *cam_proj_4x4,
    REAL *points3d_fx9, REAL *points2d_fx6, REAL *normalz_fx1, REAL *normal_fx3)
{
    (void)pnum;
    for (int f = 0; f < fnum; f++) {
        REAL pc[3][3];
        for (int c = 0; c < 3; c++) {
            const REAL *v = verts_px3 + (long)faces_fx3[f * 3 + c] * 3;
            const REAL d0 = v[0] - cam_pos_3[0], d1 = v[1] - cam_pos_3[1], d2 = v[2] - cam_pos_3[2];
            for (int j = 0; j < 3; j++) {
                const REAL *r = cam_rot_3x3 + j * 3;
                pc[c][j] = FMA(r[2], d2, FMA(r[1], d1, r[0] * d0));
            }
            REAL clip[4];
            for (int k = 0; k < 4; k++) {
                clip[k] = FMA(pc[c][2], cam_proj_4x4[2 * 4 + k],
                              FMA(pc[c][1], cam_proj_4x4[1 * 4 + k], pc[c][0] * cam_proj_4x4[0 * 4 + k]))
                          + cam_proj_4x4[3 * 4 + k];
            }
            points3d_fx9[f * 9 + c * 3 + 0] = pc[c][0];
            points3d_fx9[f * 9 + c * 3 + 1] = pc[c][1];
            points3d_fx9[f * 9 + c * 3 + 2] = pc[c][2];
            points2d_fx6[f * 6 + c * 2 + 0] = clip[0] / clip[3];
            points2d_fx6[f * 6 + c * 2 + 1] = clip[1] / clip[3];
        }
        const REAL e1x = pc[1][0] - pc[0][0], e1y = pc[1][1] - pc[0][1], e1z = pc[1][2] - pc[0][2];
        const REAL e2x = pc[2][0] - pc[0][0], e2y = pc[2][1] - pc[0][1], e2z = pc[2][2] - pc[0][2];
        const REAL nx = FMA(e1y, e2z, -(e1z * e2y));
        const REAL ny = FMA(e1z, e2x, -(e1x * e2z));
        const REAL nz = FMA(e1x, e2y, -(e1y * e2x));
        normal_fx3[f * 3 + 0] = nx;
        normal_fx3[f * 3 + 1] = ny;
        normal_fx3[f * 3 + 2] = nz;
        normalz_fx1[f] = nz;
    }
}

/* Camera set-up in a FIXED fp32 operation order (spec of the fused path's pose mode).  Restates
 * renderer/base.py:169-170 (cam_view_R = diag(1,-1,-1) R, cam_view_pos = -(R^T t)) and
 * utils/perspective.py:96-129 (projectiveprojection_real with x0 = y0 = 0).  The reference does the
 * R^T t product with torch.matmul (order unspecified); fixed here as fma(R2k,t2,fma(R1k,t1,R0k*t0)). */
void FN(dibr_oracle_camera)(const REAL *R_3x3, const REAL *t_3, const REAL *K_3x3, int width, int height,
                            double znear, double zfar, REAL *cam_rot_3x3, REAL *cam_pos_3, REAL *cam_proj_4x4)
{
    for (int k = 0; k < 3; k++) {
        cam_rot_3x3[k] = R_3x3[k];
        cam_rot_3x3[3 + k] = -R_3x3[3 + k];
        cam_rot_3x3[6 + k] = -R_3x3[6 + k];
        cam_pos_3[k] = -FMA(R_3x3[6 + k], t_3[2], FMA(R_3x3[3 + k], t_3[1], R_3x3[k] * t_3[0]));
    }
    const REAL w = (REAL)width, h = (REAL)height;
    for (int i = 0; i < 16; i++) cam_proj_4x4[i] = 0;
    cam_proj_4x4[0] = ((REAL)2 * K_3x3[0]) / w;
    cam_proj_4x4[4] = ((REAL)-2 * K_3x3[1]) / w;
    cam_proj_4x4[5] = ((REAL)2 * K_3x3[4]) / h;
    cam_proj_4x4[8] = (((REAL)-2 * K_3x3[2]) + w) / w;
    cam_proj_4x4[9] = (((REAL)2 * K_3x3[5]) - h) / h;
    cam_proj_4x4[10] = (REAL)(-(zfar + znear) / (zfar - znear));
    cam_proj_4x4[14] = (REAL)(-2.0 * (zfar * znear) / (zfar - znear));
    cam_proj_4x4[11] = (REAL)-1;
}

#undef CAT_
#undef CAT
#undef FN
