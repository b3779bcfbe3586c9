// Branch-light float P3P used ONLY as the rejection prefilter of the sampling stage (hyp.cu).
//
// Same mathematics as p3p_lambdas<> in esac_geom.cuh (pencil of the two distance conics, one degenerate member,
// two lines, line-conic intersections), specialised for throughput: one real root of the cubic (any real root is
// enough whenever the P3P has a real solution), no loops over dynamically sized sets, no dynamically indexed arrays
// (3-way selects instead), reciprocal/rsqrt approximations.  Every decision that is numerically borderline returns
// "may pass", i.e. hands the try to the exact fp64 path; the invariant "an accepted try is never rejected here" is
// what tests/test_host_geom.py::test_float_prefilter_never_rejects_an_accepted_try checks.
#pragma once
#include "esac_geom.cuh"
#ifndef DBG
#define DBG(...)
#endif

namespace esacb200 {

ESAC_HD float sel3(int k, float a0, float a1, float a2) { return k == 0 ? a0 : (k == 1 ? a1 : a2); }

struct FastLine { float l0, l1, l2; };

// Intersections of the line {lam : l.lam = 0} with the conic lam^T D lam = 0 (D symmetric: d00 d01 d02 d11 d12 d22).
// Writes two direction vectors; returns false when there is no real intersection.  `unc` is raised on borderline signs.
ESAC_HD bool fast_line_conic(const FastLine& L, float d00, float d01, float d02, float d11, float d12, float d22,
                             float o[2][3], bool& unc) {
    using N = Num<float>;
    const float a0 = fabsf(L.l0), a1 = fabsf(L.l1), a2 = fabsf(L.l2);
    const int k = (a0 >= a1 && a0 >= a2) ? 0 : (a1 >= a2 ? 1 : 2);
    // permuted coordinates (i, j, k) = (k+1, k+2, k)
    const float li = sel3(k, L.l1, L.l2, L.l0), lj = sel3(k, L.l2, L.l0, L.l1), lk = sel3(k, L.l0, L.l1, L.l2);
    if (!(fabsf(lk) > 0.f)) { unc = true; return false; }
    const float Dii = sel3(k, d11, d22, d00), Djj = sel3(k, d22, d00, d11), Dkk = sel3(k, d00, d11, d22);
    const float Dij = sel3(k, d12, d02, d01), Dik = sel3(k, d01, d12, d02), Djk = sel3(k, d02, d01, d12);
    const float ilk = N::div_(1.f, lk);
    const float r = li * ilk, s = lj * ilk;
    const float A = Dii - 2.f * r * Dik + r * r * Dkk;
    const float B = Dij - s * Dik - r * Djk + r * s * Dkk;
    const float C = Djj - 2.f * s * Djk + s * s * Dkk;
    float disc = B * B - A * C;
    const float mag = B * B + fabsf(A * C);
    if (fabsf(disc) < N::kUncertain * mag) unc = true;
    if (disc < -N::kDiscTol * mag) return false;
    disc = disc < 0.f ? 0.f : disc;
    const float sq = N::sqrt_(disc);
    const float q = -(B + (B >= 0.f ? sq : -sq));
    float al[2], be[2];
    if (fabsf(A) >= fabsf(C)) {
        if (!(fabsf(A) > 0.f)) { unc = true; return false; }
        al[0] = N::div_(q, A); al[1] = (q != 0.f) ? N::div_(C, q) : al[0];
        be[0] = 1.f; be[1] = 1.f;
    } else {
        be[0] = N::div_(q, C); be[1] = (q != 0.f) ? N::div_(A, q) : be[0];
        al[0] = 1.f; al[1] = 1.f;
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float vi = al[t], vj = be[t], vk = -r * al[t] - s * be[t];
        o[t][0] = sel3(k, vk, vj, vi);
        o[t][1] = sel3(k, vi, vk, vj);
        o[t][2] = sel3(k, vj, vi, vk);
    }
    return true;
}

// Returns false only when the try certainly fails the 4-point gate.
constexpr float kPrefilterMargin = 2.f;    // 4th-point error band, in units of tau, inside which the exact path decides (MC: no false reject down to 1.5)
constexpr float kPrefilterNeedle = 0.02f;  // shortest / longest squared side below which the triangle goes to the exact path

ESAC_HD bool p3p_may_pass_fast(const float obj[4][3], const float img[4][2], float f, float ppx, float ppy, float tau,
                               float margin = kPrefilterMargin, float needle = kPrefilterNeedle) {
    using N = Num<float>;
    // ---- bearings, recentred scene points --------------------------------------------------------------------
    float y[3][3], x1[3], x2[3], x3[3];
    const float ifx = N::div_(1.f, f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float bx = (img[i][0] - ppx) * ifx, by = (img[i][1] - ppy) * ifx;
        const float n = rsqrtf(bx * bx + by * by + 1.f);
        y[i][0] = bx * n; y[i][1] = by * n; y[i][2] = n;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { x1[c] = obj[1][c] - obj[0][c]; x2[c] = obj[2][c] - obj[0][c]; x3[c] = obj[3][c] - obj[0][c]; }
    const float a12 = x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2];
    const float a13 = x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2];
    const float e0 = x1[0] - x2[0], e1 = x1[1] - x2[1], e2 = x1[2] - x2[2];
    const float a23 = e0 * e0 + e1 * e1 + e2 * e2;
    const float amax = fmaxf(a12, fmaxf(a13, a23)), amin = fminf(a12, fminf(a13, a23));
    if (!(amax > 0.f) || !(amax < 1e30f)) return !(amax == amax) ? true : (amax > 0.f);  // all-zero: certain reject; NaN/huge: exact path
    const float c12 = y[0][0] * y[1][0] + y[0][1] * y[1][1] + y[0][2] * y[1][2];
    const float c13 = y[0][0] * y[2][0] + y[0][1] * y[2][1] + y[0][2] * y[2][2];
    const float c23 = y[1][0] * y[2][0] + y[1][1] * y[2][1] + y[1][2] * y[2][2];
    const float cm = fmaxf(fabsf(c12), fmaxf(fabsf(c13), fabsf(c23)));
    if (amin < needle * amax || cm > 0.9999f) return true;  // needle triangle / nearly parallel bearings
    const float iam = N::div_(1.f, amax);
    const float s12 = a12 * iam, s13 = a13 * iam, s23 = a23 * iam;
    // ---- the two conics (symmetric storage 00 01 02 11 12 22) ------------------------------------------------
    const float p00 = s23, p01 = -s23 * c12, p02 = 0.f, p11 = s23 - s12, p12 = s12 * c23, p22 = -s12;     // D1
    const float q00 = s23, q01 = 0.f, q02 = -s23 * c13, q11 = -s13, q12 = s13 * c23, q22 = s23 - s13;     // D2
    // adjugates (symmetric)
    const float P00 = p11 * p22 - p12 * p12, P01 = p02 * p12 - p01 * p22, P02 = p01 * p12 - p02 * p11;
    const float P11 = p00 * p22 - p02 * p02, P12 = p01 * p02 - p00 * p12, P22 = p00 * p11 - p01 * p01;
    const float Q00 = q11 * q22 - q12 * q12, Q01 = q02 * q12 - q01 * q22, Q02 = q01 * q12 - q02 * q11;
    const float Q11 = q00 * q22 - q02 * q02, Q12 = q01 * q02 - q00 * q12, Q22 = q00 * q11 - q01 * q01;
    const float k0 = p00 * P00 + p01 * P01 + p02 * P02;
    const float k3 = q00 * Q00 + q01 * Q01 + q02 * Q02;
    const float k1 = P00 * q00 + P11 * q11 + P22 * q22 + 2.f * (P01 * q01 + P02 * q02 + P12 * q12);  // tr(adj(D1) D2)
    const float k2 = Q00 * p00 + Q11 * p11 + Q22 * p22 + 2.f * (Q01 * p01 + Q02 * p02 + Q12 * p12);  // tr(D1 adj(D2))
    const float ksc = fabsf(k3) + fabsf(k2) + fabsf(k1) + fabsf(k0);
    if (!(fabsf(k3) > 1e-4f * ksc)) return true;  // (nearly) degenerate cubic: exact path
    // ---- one real root of k3 g^3 + k2 g^2 + k1 g + k0 ---------------------------------------------------------
    const float ik3 = N::div_(1.f, k3);
    const float a = k2 * ik3, b = k1 * ik3, c = k0 * ik3;
    // Depressed form t^3 - 3 Q t + 2 R = 0 (t = g + a/3).  Newton from outside the root bound |t| <= 2 max(sqrt|Q|, cbrt|R|),
    // on the side of the inflection point where the cubic still has to cross zero: monotone convergence to a real root,
    // with no dependence on the (cancellation-prone) sign of the discriminant and no divergent branches.
    const float Qq = (a * a - 3.f * b) * (1.f / 9.f), Rr = (2.f * a * a * a - 9.f * a * b + 27.f * c) * (1.f / 54.f);
    const float m3 = fmaxf(N::sqrt_(fabsf(Qq)), cbrtf(fabsf(Rr)));
    float t = (Rr > 0.f ? -2.002f : 2.002f) * m3;
#pragma unroll
    for (int it = 0; it < 10; ++it) {
        const float fv = (t * t - 3.f * Qq) * t + 2.f * Rr, dv = 3.f * (t * t - Qq);
        t -= (fabsf(dv) > 0.f) ? N::div_(fv, dv) : 0.f;
    }
    float g = t - a * (1.f / 3.f);
    {   // one polishing step on the original monic cubic
        const float fv = ((g + a) * g + b) * g + c, dv = (3.f * g + 2.f * a) * g + b;
        g -= (fabsf(dv) > 0.f) ? N::div_(fv, dv) : 0.f;
    }
    DBG("root g=%g\n", g);
    if (!(g == g)) return true;
    // ---- degenerate member D0 and the other conic -------------------------------------------------------------
    const bool small = fabsf(g) <= 1.f;
    const float w1 = small ? 1.f : N::div_(1.f, g), w2 = small ? g : 1.f;
    const float d00 = w1 * p00 + w2 * q00, d01 = w1 * p01 + w2 * q01, d02 = w1 * p02 + w2 * q02;
    const float d11 = w1 * p11 + w2 * q11, d12 = w1 * p12 + w2 * q12, d22 = w1 * p22 + w2 * q22;
    const float o00 = small ? q00 : p00, o01 = small ? q01 : p01, o02 = small ? q02 : p02;
    const float o11 = small ? q11 : p11, o12 = small ? q12 : p12, o22 = small ? q22 : p22;
    // adj(D0) = -p p^T for a real line pair
    const float B00 = d11 * d22 - d12 * d12, B01 = d02 * d12 - d01 * d22, B02 = d01 * d12 - d02 * d11;
    const float B11 = d00 * d22 - d02 * d02, B12 = d01 * d02 - d00 * d12, B22 = d00 * d11 - d01 * d01;
    const float b0 = fabsf(B00), b1 = fabsf(B11), b2 = fabsf(B22);
    const int ib = (b0 >= b1 && b0 >= b2) ? 0 : (b1 >= b2 ? 1 : 2);
    const float bii = sel3(ib, B00, B11, B22);
    const float nD = fabsf(d00) + fabsf(d11) + fabsf(d22) + 2.f * (fabsf(d01) + fabsf(d02) + fabsf(d12));
    // D0 must be (numerically) singular; it is not when the cubic has a near-multiple root and the float root is off
    if (fabsf(d00 * B00 + d01 * B01 + d02 * B02) > 1e-4f * nD * nD * nD) return true;
    DBG("detD0=%g nD=%g bii=%g ib=%d\n", d00 * B00 + d01 * B01 + d02 * B02, nD, bii, ib);
    if (fabsf(bii) < N::kUncertain * nD * nD) return true;  // rank deficiency / sign of bii not trustworthy
    if (!(bii < 0.f)) return false;                          // complex line pair: no real P3P solution
    const float isq = rsqrtf(-bii);
    const float pv0 = sel3(ib, B00, B01, B02) * isq, pv1 = sel3(ib, B01, B11, B12) * isq, pv2 = sel3(ib, B02, B12, B22) * isq;
    // C = D0 + [p]x = 2 m l^T (rank 1): take the row and the column through its largest entry
    const float C[3][3] = {{d00, d01 - pv2, d02 + pv1}, {d01 + pv2, d11, d12 - pv0}, {d02 - pv1, d12 + pv0, d22}};
    float best = -1.f;
    FastLine L = {0, 0, 0}, Mline = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const float v = fabsf(C[r][cc]);
            if (v > best) {
                best = v;
                L.l0 = C[r][0]; L.l1 = C[r][1]; L.l2 = C[r][2];
                Mline.l0 = C[0][cc]; Mline.l1 = C[1][cc]; Mline.l2 = C[2][cc];
            }
        }
    if (!(best > 0.f)) return true;
    // ---- up to four depth directions ---------------------------------------------------------------------------
    float dir[4][3];
    bool unc = false;
    const bool okL = fast_line_conic(L, o00, o01, o02, o11, o12, o22, &dir[0], unc);
    const bool okM = fast_line_conic(Mline, o00, o01, o02, o11, o12, o22, &dir[2], unc);
    DBG("okL=%d okM=%d unc=%d\n", (int)okL, (int)okM, (int)unc);
    if (unc) return true;
    if (!okL && !okM) return false;
    // ---- 4th point in the frame of the scene triangle ---------------------------------------------------------
    float nrm[3];
    cross3(x1, x2, nrm);
    const float g12 = x1[0] * x2[0] + x1[1] * x2[1] + x1[2] * x2[2];
    const float nn = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
    const float det = a12 * a13 - g12 * g12;
    if (!(det > 1e-6f * a12 * a13) || !(nn > 0.f)) return true;
    const float v1 = x3[0] * x1[0] + x3[1] * x1[1] + x3[2] * x1[2], v2 = x3[0] * x2[0] + x3[1] * x2[1] + x3[2] * x2[2];
    const float idet = N::div_(1.f, det);
    const float al = (v1 * a13 - v2 * g12) * idet, be = (v2 * a12 - v1 * g12) * idet;
    const float ga = N::div_(x3[0] * nrm[0] + x3[1] * nrm[1] + x3[2] * nrm[2], nn);
    const float lim = margin * tau, lim2 = lim * lim;
    const float sa = N::sqrt_(amax);
    bool pass = false;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const bool live = s < 2 ? okL : okM;
        float l0 = dir[s][0], l1 = dir[s][1], l2 = dir[s][2];
        const float q12 = l0 * l0 + l1 * l1 - 2.f * c12 * l0 * l1;
        const float q13 = l0 * l0 + l2 * l2 - 2.f * c13 * l0 * l2;
        const float q23 = l1 * l1 + l2 * l2 - 2.f * c23 * l1 * l2;
        float sc2;
        if (q12 >= q13 && q12 >= q23) sc2 = N::div_(s12, q12);
        else if (q13 >= q23) sc2 = N::div_(s13, q13);
        else sc2 = N::div_(s23, q23);
        const bool scale_ok = (sc2 > 0.f) && (sc2 < 1e30f);
        float sc = N::sqrt_(sc2);
        const float mx = fmaxf(fabsf(l0), fmaxf(fabsf(l1), fabsf(l2)));
        const bool near0 = fabsf(l0) < N::kUncertain * mx || fabsf(l1) < N::kUncertain * mx || fabsf(l2) < N::kUncertain * mx;
        const int pos = (l0 > 0.f) + (l1 > 0.f) + (l2 > 0.f), neg = (l0 < 0.f) + (l1 < 0.f) + (l2 < 0.f);
        if (neg == 3) sc = -sc;
        const bool signs_ok = (pos == 3) || (neg == 3);
        if (live && scale_ok && near0) pass = true;  // a depth changes sign within the error band: exact path
        DBG("dir %d live=%d (%g %g %g) scale_ok=%d signs_ok=%d\n", s, (int)live, l0, l1, l2, (int)scale_ok, (int)signs_ok);
        if (!(live && scale_ok && signs_ok)) continue;
        l0 *= sc; l1 *= sc; l2 *= sc;
        const float res = fabsf(l0 * l0 + l1 * l1 - 2.f * c12 * l0 * l1 - s12) + fabsf(l0 * l0 + l2 * l2 - 2.f * c13 * l0 * l2 - s13) +
                          fabsf(l1 * l1 + l2 * l2 - 2.f * c23 * l1 * l2 - s23);
        if (!(res < 1e-4f * (l0 * l0 + l1 * l1 + l2 * l2 + 1.f))) pass = true;  // inaccurate candidate: not trusted
        float P0[3], u1[3], u2[3], m[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            P0[cc] = l0 * sa * y[0][cc];
            u1[cc] = l1 * sa * y[1][cc] - P0[cc];
            u2[cc] = l2 * sa * y[2][cc] - P0[cc];
        }
        cross3(u1, u2, m);
        const float xc = P0[0] + al * u1[0] + be * u2[0] + ga * m[0];
        const float yc = P0[1] + al * u1[1] + be * u2[1] + ga * m[1];
        const float zc = P0[2] + al * u1[2] + be * u2[2] + ga * m[2];
        const float iz = N::div_(1.f, zc);
        const float du = ppx + f * xc * iz - img[3][0], dv = ppy + f * yc * iz - img[3][1];
        const float e = du * du + dv * dv;
        DBG("cand %d lam=(%g %g %g) res=%g e4=%g zc=%g\n", s, l0, l1, l2, res, sqrtf(e), zc);
        if (!(e > lim2)) pass = true;  // close enough (or NaN/inf): let the exact path decide
    }
    return pass;
}

}  // namespace esacb200
