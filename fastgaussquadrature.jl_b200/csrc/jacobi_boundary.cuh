// jacobi_boundary.cuh — the boundary (Bessel-type) asymptotics of Gauss-Jacobi on the device:
//   asy2, feval_asy2            src/gaussjacobi.jl:363-468
//   asy2_higherterms            src/gaussjacobi.jl:470-568
//   bary, bessel_taylor         src/gaussjacobi.jl:601-654   (nck_mat :656-667 -> binomials directly)
// One 640-thread block per end of the interval: 10 nodes x 64 Bessel orders.  Every Newton sweep the
// block evaluates all J_{alpha+nu}(rho t_p), nu = -kmax..kmax, in parallel (double-double ascending
// series, bessel.cuh); threads 0..9 then assemble the Taylor-shifted Bessel function, the expansion and
// the Newton update of their node; thread 0 applies the inf-norm stopping rule (:383).
#ifndef FGQ_JACOBI_BOUNDARY_CUH_
#define FGQ_JACOBI_BOUNDARY_CUH_

#include "bessel.cuh"
#include "jacobi_plan.h"
#include "jacobi_tables.h"

namespace fgq {

__constant__ double c_cumsummat10[10][10] = FGQ_CUMSUMMAT_10_INIT;
__constant__ double c_diffmat10[10][10] = FGQ_DIFFMAT_10_INIT;

#define EPS 2.220446049250313e-16
#define PI 3.141592653589793

// ---- boundary: :470-568 -------------------------------------------------------------------------------
struct HigherTerms {
    double t[10], v[10], tB1[10], A2[10], tB2[10], A3[10];
};

__device__ inline void matvec10(const double (*M)[10], double scale, const double* x, double* y) {
    for (int i = 0; i < 10; ++i) {
        double s = 0.0;
        for (int j = 0; j < 10; ++j) s += (scale * M[i][j]) * x[j];
        y[i] = s;
    }
}

__device__ inline void asy2_higherterms(double al, double be, const double* theta, int npts, HigherTerms* H) {
    const int N = 10;
    const double A = 0.25 - al * al, B = 0.25 - be * be;
    double c = 0.5;
    for (int i = 0; i < npts; ++i) c = fmax(c, theta[i]);
    double* t = H->t;
    double* v = H->v;
    for (int i = 0; i < N; ++i) {
        t[i] = 0.5 * c * (sin(PI * (double)(-(N - 1) + 2 * i) / (2.0 * (N - 1))) + 1);
        v[i] = 1.0;
    }
    v[0] = 0.5;
    for (int i = 1; i < N; i += 2) v[i] = -1.0;
    v[N - 1] *= 0.5;
    double g[10], gp[10], gpp[10], B0[10], B0p[10], A1[10], A1p[10], A1p_t[10], f[10];
    for (int i = 0; i < N; ++i) {
        const double ti = t[i];
        const double cot = 1.0 / tan(ti / 2), csc = 1.0 / sin(ti / 2), sec = 1.0 / cos(ti / 2);
        g[i] = A * (cot - 2 / ti) - B * tan(ti / 2);
        gp[i] = A * (2 / (ti * ti) - 0.5 * csc * csc) - 0.5 * (0.25 - be * be) * sec * sec;
        const double s2 = sin(ti / 2), cs = 1.0 / sin(ti);
        gpp[i] = A * (-4 / (ti * ti * ti) + 0.25 * sin(ti) * (csc * csc) * (csc * csc)) - 4 * B * (s2 * s2) * (s2 * s2) * (cs * cs * cs);
    }
    g[0] = 0;
    gp[0] = -A / 6 - 0.5 * B;
    gpp[0] = 0;
    for (int i = 0; i < N; ++i) {
        B0[i] = 0.25 * g[i] / t[i];
        B0p[i] = 0.25 * (gp[i] / t[i] - g[i] / (t[i] * t[i]));
    }
    B0[0] = 0.25 * (-A / 6 - 0.5 * B);
    B0p[0] = 0;
    const double A10 = al * (A + 3 * B) / 24;
    for (int i = 0; i < N; ++i) {
        A1[i] = 0.125 * gp[i] - (1 + 2 * al) / 2 * B0[i] - g[i] * g[i] / 32 - A10;
        A1p[i] = 0.125 * gpp[i] - (1 + 2 * al) / 2 * B0p[i] - gp[i] * g[i] / 16;
        A1p_t[i] = A1p[i] / t[i];
    }
    A1p_t[0] = -A / 720 - A * A / 576 - A * B / 96 - B * B / 64 - B / 48 + al * (A / 720 + B / 48);
    for (int i = 0; i < N; ++i) {
        const double ti = t[i], t2 = ti * ti;
        const double ch = 2 * cos(ti / 2);
        const double fcos = B / (ch * ch);
        double fi;
        if (ti > 0.5) {
            const double sh = 2 * sin(ti / 2);
            fi = A * (1 / t2 - 1 / (sh * sh));
        } else {
            const double t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4, t10 = t8 * t2, t12 = t6 * t6, t14 = t12 * t2, t16 = t8 * t8;
            fi = -A * (1.0 / 12 + t2 / 240 + t4 / 6048 + t6 / 172800 + t8 / 5322240 + 691 * t10 / 118879488000.0 +
                       t12 / 5748019200.0 + 3617 * t14 / 711374856192000.0 + 43867 * t16 / 300534953951232000.0);
        }
        f[i] = fi - fcos;
    }
    const double cs = 0.5 * c, dsca = 2 / c;
    double I[10], J[10], tmp[10], K[10], B1[10], A2p[10], A2p_t[10], B2[10], w[9], D1[10], C1[10], C2v[10];
    matvec10(c_cumsummat10, cs, A1p_t, I);
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * A1[i];
    matvec10(c_cumsummat10, cs, tmp, J);
    double* tB1 = H->tB1;
    for (int i = 0; i < N; ++i) tB1[i] = -0.5 * A1p[i] - (0.5 + al) * I[i] + 0.5 * J[i];
    tB1[0] = 0;
    for (int i = 0; i < N; ++i) B1[i] = tB1[i] / t[i];
    B1[0] = A / 720 + A * A / 576 + A * B / 96 + B * B / 64 + B / 48 + al * (A * A / 576 + B * B / 64 + A * B / 96) -
            al * al * (A / 720 + B / 48);
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * tB1[i];
    matvec10(c_cumsummat10, cs, tmp, K);
    matvec10(c_diffmat10, dsca, tB1, D1);
    double* A2 = H->A2;
    for (int i = 0; i < N; ++i) A2[i] = 0.5 * D1[i] - (0.5 + al) * B1[i] - 0.5 * K[i];
    const double A20 = A2[0];
    for (int i = 0; i < N; ++i) A2[i] -= A20;
    matvec10(c_diffmat10, dsca, A2, A2p);
    const double A2p0 = A2p[0];
    for (int i = 0; i < N; ++i) A2p[i] -= A2p0;
    for (int i = 0; i < N; ++i) A2p_t[i] = A2p[i] / t[i];
    for (int i = 0; i < N - 1; ++i) w[i] = PI / 2 - t[i + 1];
    for (int i = 1; i < N - 1; i += 2) w[i] = -w[i];
    w[N - 2] = 0.5 * w[N - 2];
    double sw = 0.0, swa = 0.0;
    for (int i = 0; i < N - 1; ++i) {
        sw += w[i];
        swa += w[i] * A2p_t[i + 1];
    }
    A2p_t[0] = swa / sw;
    matvec10(c_cumsummat10, cs, A2p_t, C1);
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * A2[i];
    matvec10(c_cumsummat10, cs, tmp, C2v);
    double* tB2 = H->tB2;
    for (int i = 0; i < N; ++i) tB2[i] = -0.5 * A2p[i] - (0.5 + al) * C1[i] + 0.5 * C2v[i];
    for (int i = 0; i < N; ++i) B2[i] = tB2[i] / t[i];
    double swb = 0.0;
    for (int i = 0; i < N - 1; ++i) swb += w[i] * B2[i + 1];
    B2[0] = swb / sw;
    for (int i = 0; i < N; ++i) tmp[i] = f[i] * tB2[i];
    matvec10(c_cumsummat10, cs, tmp, K);
    matvec10(c_diffmat10, dsca, tB2, D1);
    double* A3 = H->A3;
    for (int i = 0; i < N; ++i) A3[i] = 0.5 * D1[i] - (0.5 + al) * B2[i] - 0.5 * K[i];
    const double A30 = A3[0];
    for (int i = 0; i < N; ++i) A3[i] -= A30;
}

// :601-622
__device__ inline double bary(double x, const double* fvals, const double* xk, const double* vk) {
    double num = 0.0, den = 0.0;
    for (int k = 0; k < 10; ++k) {
        const double xx = vk[k] / (x - xk[k]);
        num += xx * fvals[k];
        den += xx;
    }
    double r = num / den;
    if (r != r) {
        for (int k = 0; k < 10; ++k)
            if (x == xk[k]) return fvals[k];
    }
    return r;
}


// :624-654 — J_a(z + t) by a Taylor expansion about z, given Bjk[nu + 30] = J_{a+nu}(z), |nu| <= kmax
__device__ inline double bessel_taylor_node(double t, int kmax, const double* Bjk) {
    double acc = Bjk[30];
    double fact = 1.0, tp = 1.0, two_k = 1.0;
    for (int k = 1; k <= kmax; ++k) {
        double aa = 0.0, sgn = 1.0, binom = 1.0;
        for (int l = 0; l <= k; ++l) {
            aa = aa + sgn * binom * Bjk[30 + 2 * l - k];
            sgn = -sgn;
            binom = binom * (double)(k - l) / (double)(l + 1);
        }
        fact = k * fact;
        two_k *= 2.0;
        aa /= two_k * fact;
        tp *= t;
        acc += aa * tp;
    }
    return acc;
}

// :405-468 for one node, Bessel values supplied
__device__ inline void feval_asy2_node(const JacobiBoundaryPlan& P, const HigherTerms& H, double ti, double Ja,
                                       double Jb, double Jbb, double Jab, double& val, double& der) {
    const double nn = (double)P.n, al = P.al, be = P.be;
    const double rho = P.rho, rho2 = P.rho2;
    const double A = 0.25 - al * al, B = 0.25 - be * be;
    const double A10 = al * (A + 3 * B) / 24;
    const double cot = 1.0 / tan(ti / 2), csc = 1.0 / sin(ti / 2), sec = 1.0 / cos(ti / 2);
    const double gt = A * (cot - (2 / ti)) - B * tan(ti / 2);
    const double gtdx = A * (2 / (ti * ti) - 0.5 * csc * csc) - 0.5 * B * sec * sec;
    const double tB0 = 0.25 * gt;
    const double A1 = gtdx / 8 - (1 + 2 * al) / 8 * gt / ti - gt * gt / 32 - A10;
    const double tB1t = bary(ti, H.tB1, H.t, H.v), A2t = bary(ti, H.A2, H.t, H.v);
    const double tB2t = bary(ti, H.tB2, H.t, H.v), A3t = bary(ti, H.A3, H.t, H.v);
    const double r2 = rho * rho, r3 = r2 * rho, r4 = r2 * r2, r5 = r4 * rho, r6 = r3 * r3;
    const double q2 = rho2 * rho2, q3 = q2 * rho2, q4 = q2 * q2, q5 = q4 * rho2, q6 = q3 * q3;
    double v = Ja + Jb * tB0 / rho + Ja * A1 / r2 + Jb * tB1t / r3 + Ja * A2t / r4;
    double v2 = Jab + Jbb * tB0 / rho2 + Jab * A1 / q2 + Jbb * tB1t / q3 + Jab * A2t / q4;
    v += Jb * tB2t / r5 + Ja * A3t / r6;
    v2 += Jbb * tB2t / q5 + Jab * A3t / q6;
    const double valstmp = P.C * v;
    const double denom = pow(sin(ti / 2), al + 0.5) * pow(cos(ti / 2), be + 0.5);
    val = sqrt(ti) * valstmp / denom;
    double d = (nn * (al - be - (2 * nn + al + be) * cos(ti)) * valstmp + 2 * (nn + al) * (nn + be) * P.C2 * v2) /
               (2 * nn + al + be);
    d *= sqrt(ti) / (denom * sin(ti));
    der = d;
}

constexpr int JB_THREADS = 640;

// asy2 (:363-398) for one end; writes the 10 nodes/weights straight into the rule.
// side 0: right end (x -> +1), global indices n-10..n-1 ascending; side 1: left end, mirrored.
// One block per end (blockIdx.x selects the plan): the two ends are independent and run side by side.  Dynamic shared
// memory: the reciprocals 1 / ((k+1)(nu+k+1)) of the Bessel series' term ratios for the 62 orders, interleaved
// [k][order] (JB_TABLE_BYTES), filled by the block itself before the first sweep — every sweep then runs 640
// series whose dependent chain is two double-double products per term instead of a double-double quotient.
struct JacobiBoundaryPlans {
    JacobiBoundaryPlan end[2];
};
constexpr size_t JB_TABLE_BYTES = (size_t)BESSEL_INV_DEN * 64 * sizeof(ddv);

__global__ void __launch_bounds__(JB_THREADS)
jacobi_boundary_kernel(const __grid_constant__ JacobiBoundaryPlans plans, double* x, double* w) {
    const JacobiBoundaryPlan& P = plans.end[blockIdx.x];
    extern __shared__ __align__(16) unsigned char jb_dyn_smem[];
    ddv* inv_den = reinterpret_cast<ddv*>(jb_dyn_smem);  // [BESSEL_INV_DEN][64]
    for (int e = threadIdx.x; e < BESSEL_INV_DEN * 64; e += JB_THREADS) {
        const int k = e >> 6, oo = e & 63;
        if (oo <= 61) inv_den[e] = dd_div({1.0, 0.0}, besselj_term_den(P.ord[oo].nu, k));
    }
    __shared__ double t[JAC_NBDY], vals[JAC_NBDY], ders[JAC_NBDY];
    __shared__ double Bjk[JAC_NBDY][64];
    __shared__ HigherTerms H;
    __shared__ int s_kmax, s_phase;  // phase 0: iterate, 1: final evaluation, 2: done
    const int tid = threadIdx.x;
    const int p = tid >> 6, o = tid & 63;
    if (tid == 0) {
        const double r = (double)P.n + 0.5 * (P.al + P.be + 1);
        for (int i = 0; i < JAC_NBDY; ++i) {
            const double phik = P.jk[i] / r;
            t[i] = phik + ((P.al * P.al - 0.25) * (1 - phik * (1.0 / tan(phik))) / (2 * phik) -
                           0.25 * (P.al * P.al - P.be * P.be) * tan(0.5 * phik)) / (r * r);
        }
        asy2_higherterms(P.al, P.be, t, JAC_NBDY, &H);
        s_phase = 0;
    }
    __syncthreads();
    for (int it = 0; it < 11; ++it) {
        if (tid == 0) {
            double tinf = 0.0;
            for (int i = 0; i < JAC_NBDY; ++i) tinf = fmax(tinf, fabs(t[i]));
            int kmax = (int)ceil(fabs(log(EPS) / log(tinf)));
            s_kmax = kmax > 30 ? 30 : (kmax < 0 ? 0 : kmax);
        }
        __syncthreads();
        const int kmax = s_kmax;
        // o in 0..60: order alpha + (o - 30) at rho t_p (needed for |o - 30| <= max(kmax, 1));
        // o == 62: J_{alpha+1}(rho2 t_p)
        {
            const int nu = o - 30;
            const int need = kmax > 1 ? kmax : 1;
            if (o <= 60 && nu >= -need && nu <= need) Bjk[p][o] = besselj_series_dd_tab(P.ord[o], P.rho * t[p], inv_den + o, 64);
            if (o == 62) Bjk[p][62] = besselj_series_dd_tab(P.ord[31], P.rho2 * t[p], inv_den + 31, 64);
        }
        __syncthreads();
        if (tid < JAC_NBDY) {
            const double Jab = bessel_taylor_node(-t[tid], kmax, Bjk[tid]);
            feval_asy2_node(P, H, t[tid], Bjk[tid][30], Bjk[tid][31], Bjk[tid][62], Jab, vals[tid], ders[tid]);
        }
        __syncthreads();
        if (tid == 0) {
            double mx = 0.0;
            bool nan = false;
            for (int i = 0; i < JAC_NBDY; ++i) {
                const double dt = vals[i] / ders[i];
                t[i] += dt;
                if (dt != dt) nan = true;
                mx = fmax(mx, fabs(dt));
            }
            if (s_phase == 1) s_phase = 2;                                     // the extra evaluation (:388-390)
            else if ((!nan && mx < sqrt(EPS) / 200) || it == 9) s_phase = 1;    // :383, or 10 iterations used up
        }
        __syncthreads();
        if (s_phase == 2) break;
    }
    if (tid < JAC_NBDY) {
        // flip (:392-396): node i (ascending towards the end point) uses t[9 - i]
        const int r_i = JAC_NBDY - 1 - tid;
        const double xv = cos(t[r_i]);
        const double wv = (1.0 / (ders[r_i] * ders[r_i])) * P.wc;
        int64_t ig;
        double xo;
        if (P.side == 0) {
            ig = P.n - JAC_NBDY + tid;
            xo = xv;
        } else {  // x[bdyidx2] = -xbdy (:189-190): ascending from -1
            ig = JAC_NBDY - 1 - tid;
            xo = -xv;
        }
        if (ig >= P.lo && ig < P.hi) {
            x[ig - P.lo] = xo;
            w[ig - P.lo] = wv;
        }
    }
}

#undef EPS
#undef PI

}  // namespace fgq
#endif
