// cv_b200/csrc/geom.cu -- batched geometric verification on sm_100a (f64).
//
// Every model hypothesis (eight-point / P3P minimal solve) and every (hypothesis, datum) residual runs on
// the GPU, one thread per hypothesis resp. per (hypothesis, datum) pair; ARRSAC's inherently sequential
// bookkeeping (likelihood-ratio test over hypotheses, sort / truncate, RNG draws) stays on the host and
// consumes bit-packed inlier masks.  Reference lines are cited per function (paths relative to /root/reference).
// The linear algebra that lives in nalgebra upstream (symmetric eigen, SVD, from_matrix_eps) is implemented
// here as cyclic Jacobi / closed forms; f64 results are held to 1e-6 relative (BASELINE north_star) in the parity tests.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "c2c_filter.cuh"

namespace {

// ------------------------------------------------------------------------------------------ device math
// One cyclic Jacobi rotation, written with explicit roundings so that the one-thread and the lane-cooperative eigensolvers
// (below) produce the same bits whatever the compiler would contract.
__device__ __forceinline__ void jacobi_cs(double app, double aqq, double apq, double &c, double &s) {
    const double theta = __ddiv_rn(__dsub_rn(aqq, app), __dmul_rn(2.0, apq));
    const double t = __ddiv_rn(theta >= 0.0 ? 1.0 : -1.0, __dadd_rn(fabs(theta), __dsqrt_rn(__dadd_rn(__dmul_rn(theta, theta), 1.0))));
    c = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(t, t), 1.0)));
    s = __dmul_rn(t, c);
}
__device__ __forceinline__ void jacobi_rot(double &x, double &y, double c, double s) {
    const double a = x, b = y;
    x = __dsub_rn(__dmul_rn(c, a), __dmul_rn(s, b));
    y = __dadd_rn(__dmul_rn(s, a), __dmul_rn(c, b));
}

template <int N>
__device__ bool sym_eigen(const double *Ain, double eps, int max_sweeps, double *d, double *V) {
    double A[N * N];
#pragma unroll
    for (int i = 0; i < N * N; i++) A[i] = Ain[i];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) V[i * N + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; i++) {
            diag = __dadd_rn(diag, __dmul_rn(A[i * N + i], A[i * N + i]));
            for (int j = i + 1; j < N; j++) off = __dadd_rn(off, __dmul_rn(A[i * N + j], A[i * N + j]));
        }
        if (off <= eps * eps * diag || off == 0.0) {
            for (int i = 0; i < N; i++) d[i] = A[i * N + i];
            return true;
        }
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                double c, s;
                jacobi_cs(A[p * N + p], A[q * N + q], apq, c, s);
                for (int k = 0; k < N; k++) jacobi_rot(A[k * N + p], A[k * N + q], c, s);
                for (int k = 0; k < N; k++) jacobi_rot(A[p * N + k], A[q * N + k], c, s);
                for (int k = 0; k < N; k++) jacobi_rot(V[k * N + p], V[k * N + q], c, s);
            }
    }
    for (int i = 0; i < N; i++) d[i] = A[i * N + i];
    return false;
}

// The 9x9 eigensolver of the eight-point estimator: Jacobi in round-robin (tournament) order, restated on the checker's side as
// ref_sym_eigen9_rr.  A sweep is nine rounds; round r rotates the four DISJOINT index pairs {(r + k) mod 9, (r - k) mod 9}, k = 1..4:
// the angles come from the matrix in front of the round, then all column rotations (A and V), then all row rotations.  Disjoint pairs
// touch disjoint columns / rows, so JL = 1, 2 or 4 lanes (`mask` = those lanes, `lane` = 0..JL-1; A and V in shared memory, or thread
// local for JL = 1) each take their share of a round's pairs and produce the bits of the sequential order.  A sweep costs nine
// dependent rotation set-ups instead of 36, and the set-up itself is quotient-free: with d = aqq - app, h = 2 apq,
// w = |d| + sqrt(d^2 + h^2), n = sqrt(w^2 + h^2): c = w / n, s = +-|h| / n (two square roots and one level of division on the chain).
__device__ __forceinline__ void jacobi_cs_rr(double app, double aqq, double apq, double &c, double &s) {
    const double d = __dsub_rn(aqq, app), h = __dmul_rn(2.0, apq);
    const double w = __dadd_rn(fabs(d), __dsqrt_rn(__dadd_rn(__dmul_rn(d, d), __dmul_rn(h, h))));
    const double n = __dsqrt_rn(__dadd_rn(__dmul_rn(w, w), __dmul_rn(h, h)));
    const bool pos = d == 0.0 || ((d > 0.0) == (h > 0.0));
    c = __ddiv_rn(w, n);
    s = __ddiv_rn(pos ? fabs(h) : -fabs(h), n);
}
template <int JL>
__device__ bool sym_eigen9_rr(double *A, double *V, int lane, unsigned mask, double eps, int max_sweeps) {
    // JL = 1, 2, 4: a lane owns 4 / JL pairs of a round and all nine rows / columns of their updates;
    // JL = 8, 16: SUB = JL / 4 lanes share a pair (each forms the rotation itself) and split the nine rows / columns
    constexpr int N = 9, SUB = JL > 4 ? JL / 4 : 1, PL = JL >= 4 ? 1 : 4 / JL, PSTEP = JL / SUB;
    const int lp = lane / SUB, sub = lane % SUB;
    for (int e = lane; e < N * N; e += JL) V[e] = (e / N == e % N) ? 1.0 : 0.0;
    if (JL > 1) __syncwarp(mask);
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; i++) {
            diag = __dadd_rn(diag, __dmul_rn(A[i * N + i], A[i * N + i]));
            for (int j = i + 1; j < N; j++) off = __dadd_rn(off, __dmul_rn(A[i * N + j], A[i * N + j]));
        }
        if (off <= eps * eps * diag || off == 0.0) return true;
        for (int r = 0; r < N; r++) {
            int P[PL], Q[PL];
            bool act[PL];
            double Cc[PL], Ss[PL];
#pragma unroll
            for (int i = 0; i < PL; i++) {
                const int k = 1 + lp + i * PSTEP;
                int a = r + k, b = r + N - k;
                if (a >= N) a -= N;
                if (b >= N) b -= N;
                P[i] = min(a, b); Q[i] = max(a, b);
                const double apq = A[P[i] * N + Q[i]];
                act[i] = apq != 0.0;
                if (act[i]) jacobi_cs_rr(A[P[i] * N + P[i]], A[Q[i] * N + Q[i]], apq, Cc[i], Ss[i]);
            }
            // every lane has read its pivot block before a lane of the same pair rewrites it (lanes of OTHER pairs never touch it:
            // they write their own pairs' columns only)
            if (SUB > 1) __syncwarp(mask);
#pragma unroll
            for (int i = 0; i < PL; i++) {
                if (!act[i]) continue;
                for (int k = sub; k < N; k += SUB) {
                    jacobi_rot(A[k * N + P[i]], A[k * N + Q[i]], Cc[i], Ss[i]);
                    jacobi_rot(V[k * N + P[i]], V[k * N + Q[i]], Cc[i], Ss[i]);
                }
            }
            if (JL > 1) __syncwarp(mask);
#pragma unroll
            for (int i = 0; i < PL; i++) {
                if (!act[i]) continue;
                for (int k = sub; k < N; k += SUB) jacobi_rot(A[P[i] * N + k], A[Q[i] * N + k], Cc[i], Ss[i]);
            }
            if (JL > 1) __syncwarp(mask);
        }
    }
    return false;
}

__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double norm3(const double *a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
    const double r0 = a[1] * b[2] - a[2] * b[1], r1 = a[2] * b[0] - a[0] * b[2], r2 = a[0] * b[1] - a[1] * b[0];
    o[0] = r0; o[1] = r1; o[2] = r2;
}
__device__ void mat3_mul(const double *a, const double *b, double *o) {
    double r[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
    for (int i = 0; i < 9; i++) o[i] = r[i];
}
__device__ __forceinline__ double det3(const double *m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// sorted SVD of a 3x3 matrix through the eigen-decomposition of MtM; u3 = u1 x u2 (its sign is normalised by
// the det(U) > 0 rule of essential.rs:139-143 anyway)
__device__ __noinline__ bool svd3(const double *M, double eps, int iters, double *U, double *Vt) {
    double MtM[9], d[3], V[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) MtM[i * 3 + j] = M[i] * M[j] + M[3 + i] * M[3 + j] + M[6 + i] * M[6 + j];
    if (!sym_eigen<3>(MtM, eps, iters, d, V)) return false;
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (d[ord[j]] > d[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double v[3][3], u[3][3], s[3];
    for (int k = 0; k < 3; k++) {
        for (int r = 0; r < 3; r++) v[k][r] = V[r * 3 + ord[k]];
        s[k] = sqrt(d[ord[k]] > 0.0 ? d[ord[k]] : 0.0);
    }
    const double tiny = 1e-12 * (s[0] > 0.0 ? s[0] : 1.0);
    for (int k = 0; k < 2; k++) {
        if (!(s[k] > tiny)) return false;
        for (int r = 0; r < 3; r++) u[k][r] = (M[r * 3] * v[k][0] + M[r * 3 + 1] * v[k][1] + M[r * 3 + 2] * v[k][2]) / s[k];
    }
    cross3(u[0], u[1], u[2]);
    const double nn = norm3(u[2]);
    if (!(nn > 0.0)) return false;
    for (int r = 0; r < 3; r++) u[2][r] /= nn;
    for (int k = 0; k < 3; k++)
        for (int r = 0; r < 3; r++) { U[r * 3 + k] = u[k][r]; Vt[k * 3 + r] = v[k][r]; }
    return true;
}

// eight-point/src/lib.rs:11-24,43-58 (incl. b / a.z) + cv-pinhole/src/essential.rs:114-162,217-231
__device__ __forceinline__ void eight_point_row(const double *a, const double *b, uint32_t id, double *row /* 9 */) {
    const double *pa = a + 3 * (size_t)id, *pb = b + 3 * (size_t)id;
    const double ap[3] = {pa[0] / pa[2], pa[1] / pa[2], pa[2] / pa[2]};
    const double bp[3] = {pb[0] / pa[2], pb[1] / pa[2], pb[2] / pa[2]};
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) row[3 * j + k] = __dmul_rn(ap[j], bp[k]);
}
__device__ __forceinline__ double eight_point_gram(const double *D /* [8][9] */, int r, int c) {
    double s = 0.0;
    for (int i = 0; i < 8; i++) s = __dadd_rn(s, __dmul_rn(D[i * 9 + r], D[i * 9 + c]));
    return s;
}
// the four poses from the eigenvectors (V column-stacked, d = diagonal after convergence)
__device__ int eight_point_poses(const double *d, const double *V, cvb_pose *out) {
    int best = 0;
    for (int i = 1; i < 9; i++)
        if (d[i] < d[best]) best = i;
    double E[9];
    for (int k = 0; k < 9; k++) E[(k % 3) * 3 + (k / 3)] = V[k * 9 + best];   // Matrix3::from_iterator is column-major
    double U[9], Vt[9];
    if (!svd3(E, 1e-12, 1000, U, Vt)) return 0;
    if (det3(U) < 0.0) for (int r = 0; r < 3; r++) U[r * 3 + 2] *= -1.0;
    if (det3(Vt) < 0.0) for (int c = 0; c < 3; c++) Vt[6 + c] *= -1.0;
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double UW[9], Ra[9], Rb[9];
    mat3_mul(U, W, UW); mat3_mul(UW, Vt, Ra);
    mat3_mul(U, Wt, UW); mat3_mul(UW, Vt, Rb);
    const double t[3] = {U[2], U[5], U[8]};
    for (int k = 0; k < 4; k++) {
        for (int i = 0; i < 9; i++) out[k].r[i] = (k & 1) ? Rb[i] : Ra[i];
        for (int r = 0; r < 3; r++) out[k].t[r] = (k & 2) ? -t[r] : t[r];
    }
    return 4;
}
__device__ int eight_point(const double *a, const double *b, const uint32_t *idx, cvb_pose *out) {
    double D[72], EtE[81], d[9], V[81];
    for (int i = 0; i < 8; i++) eight_point_row(a, b, idx[i], D + 9 * i);
    for (int r = 0; r < 9; r++)
        for (int c = 0; c < 9; c++) EtE[r * 9 + c] = eight_point_gram(D, r, c);
    if (!sym_eigen9_rr<1>(EtE, V, 0, 0u, 1e-12, 1000)) return 0;
    for (int i = 0; i < 9; i++) d[i] = EtE[i * 9 + i];
    return eight_point_poses(d, V, out);
}
// JL lanes per hypothesis; sh = 163 doubles of shared memory of this hypothesis (A | V, the design matrix lives in V's place first)
#define EIGHT_SH 163
template <int JL>
__device__ int eight_point_lanes(const double *a, const double *b, const uint32_t *idx, cvb_pose *out, double *sh, int lane, unsigned mask) {
    double *A = sh, *V = sh + 81;
    for (int i = lane; i < 8; i += JL) eight_point_row(a, b, idx[i], V + 9 * i);
    __syncwarp(mask);
    for (int e = lane; e < 81; e += JL) A[e] = eight_point_gram(V, e / 9, e % 9);
    __syncwarp(mask);
    const bool ok = sym_eigen9_rr<JL>(A, V, lane, mask, 1e-12, 1000);
    __syncwarp(mask);
    int n = 0;
    if (ok && lane == 0) {
        double d[9];
        for (int i = 0; i < 9; i++) d[i] = A[i * 9 + i];
        n = eight_point_poses(d, V, out);
    }
    return n;                                                    // valid on lane 0
}

__device__ __forceinline__ void from_homogeneous(double *p) {
    if (signbit(p[3])) { p[0] = -p[0]; p[1] = -p[1]; p[2] = -p[2]; p[3] = -p[3]; }
    const double n = norm3(p);
    p[0] /= n; p[1] /= n; p[2] /= n; p[3] /= n;
}
__device__ __forceinline__ void pose_apply(const cvb_pose &P, const double *x, double *o) {
    for (int r = 0; r < 3; r++) o[r] = dot3(P.r + 3 * r, x) + P.t[r] * x[3];
    o[3] = x[3];
}
__device__ void design_add(const double *R, const double *t, const double *b, double *D) {
    double M[3][4], T[3][4];
    for (int r = 0; r < 3; r++) { M[r][0] = R[3 * r]; M[r][1] = R[3 * r + 1]; M[r][2] = R[3 * r + 2]; M[r][3] = t[r]; }
    for (int c = 0; c < 4; c++) {
        const double btP = b[0] * M[0][c] + b[1] * M[1][c] + b[2] * M[2][c];
        for (int r = 0; r < 3; r++) T[r][c] = M[r][c] - b[r] * btP;
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) D[i * 4 + j] += T[0][i] * T[0][j] + T[1][i] * T[1][j] + T[2][i] * T[2][j];
}

// cv-core/src/pose.rs:249-296: two-view linear-eigen triangulation inside the residual
__device__ double residual_c2c(const cvb_pose &P, const double *a, const double *b) {
    double D[16], d[4], V[16];
    for (int i = 0; i < 16; i++) D[i] = 0.0;
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z[3] = {0, 0, 0};
    design_add(I, z, a, D);
    design_add(P.r, P.t, b, D);
    if (!sym_eigen<4>(D, 1e-12, 1024, d, V)) return 2.0;
    int best = 0;
    for (int i = 1; i < 4; i++)
        if (fabs(d[i]) < fabs(d[best])) best = i;
    double p[4] = {V[best], V[4 + best], V[8 + best], V[12 + best]};
    from_homogeneous(p);
    for (int i = 0; i < 4; i++)
        if (!isfinite(p[i])) return 2.0;
    double q[4];
    pose_apply(P, p, q);
    from_homogeneous(q);
    return 0.5 * (1.0 - dot3(a, p) + 1.0 - dot3(b, q));
}

// cv-core/src/pose.rs:194-202
__device__ double residual_w2c(const cvb_pose &P, const double *bearing, const double *world) {
    double q[4];
    pose_apply(P, world, q);
    from_homogeneous(q);
    return 1.0 - dot3(bearing, q);
}

// ---- lambda twist (lambda-twist/src/lib.rs:110-317, 361-554)
__device__ void root2real(double b, double c, double *r1, double *r2) {
    const double disc = b * b - 4.0 * c;
    if (disc < 0.0) { *r1 = *r2 = 0.5 * b; }
    else if (b < 0.0) { const double y = sqrt(disc); *r1 = 0.5 * (-b + y); *r2 = 0.5 * (-b - y); }
    else { const double y = sqrt(disc); *r1 = 2.0 * c / (-b + y); *r2 = 2.0 * c / (-b - y); }
}
__device__ double cube_root(double b, double c, double d) {
    double r0;
    if (b * b >= 3.0 * c) {
        const double v = sqrt(b * b - 3.0 * c);
        const double t1 = (-b - v) / 3.0;
        double k = ((t1 + b) * t1 + c) * t1 + d;
        if (k > 0.0) r0 = t1 - sqrt(-k / (3.0 * t1 + b));
        else {
            const double t2 = (-b + v) / 3.0;
            k = ((t2 + b) * t2 + c) * t2 + d;
            r0 = t2 + sqrt(-k / (3.0 * t2 + b));
        }
    } else {
        r0 = -b / 3.0;
        if (fabs((3.0 * r0 + 2.0 * b) * r0 + c) < 1e-4) r0 += 1.0;
    }
    for (int i = 0; i < 7; i++) {
        const double fx = ((r0 + b) * r0 + c) * r0 + d, fpx = (3.0 * r0 + 2.0 * b) * r0 + c;
        r0 -= fx / fpx;
    }
    for (int i = 0; i < 43; i++) {
        const double fx = ((r0 + b) * r0 + c) * r0 + d;
        if (fabs(fx) > 1e-13) { const double fpx = (3.0 * r0 + 2.0 * b) * r0 + c; r0 -= fx / fpx; }
        else break;
    }
    return r0;
}
__device__ void eigen_decomposition_singular(const double *x, double *Ev, double *ev) {
    const double m11 = x[0], m12 = x[1], m13 = x[2], m21 = x[3], m22 = x[4], m23 = x[5], m31 = x[6], m32 = x[7], m33 = x[8];
    double v3[3] = {m21 * m32 - m31 * m22, m31 * m12 - m32 * m11, m22 * m11 - m21 * m12};
    const double n = norm3(v3);
    for (int i = 0; i < 3; i++) v3[i] /= n;
    const double x12_sqr = m12 * m12;
    const double b = -m11 - m22 - m33;
    const double c = -x12_sqr - m13 * m13 - m23 * m23 + m11 * (m22 + m33) + m22 * m33;
    double e1, e2;
    root2real(b, c, &e1, &e2);
    if (fabs(e1) < fabs(e2)) { const double t = e1; e1 = e2; e2 = t; }
    ev[0] = e1; ev[1] = e2; ev[2] = 0.0;
    const double mx0011 = -m11 * m22, prec_0 = m12 * m23 - m13 * m22, prec_1 = m12 * m13 - m11 * m23;
    const double es[2] = {e1, e2};
    double v[2][3];
    for (int k = 0; k < 2; k++) {
        const double e = es[k];
        const double tmp = 1.0 / (e * (m11 + m22) + mx0011 - e * e + x12_sqr);
        const double a1 = -(e * m13 + prec_0) * tmp, a2 = -(e * m23 + prec_1) * tmp;
        const double rnorm = 1.0 / sqrt(a1 * a1 + a2 * a2 + 1.0);
        v[k][0] = a1 * rnorm; v[k][1] = a2 * rnorm; v[k][2] = rnorm;
    }
    for (int r = 0; r < 3; r++) { Ev[r * 3] = v[0][r]; Ev[r * 3 + 1] = v[1][r]; Ev[r * 3 + 2] = v3[r]; }
}
__device__ __forceinline__ double l1n(const double *v) { return fabs(v[0]) + fabs(v[1]) + fabs(v[2]); }
__device__ void gn_residual(const double *l, double a12, double a13, double a23, double b12, double b13, double b23, double *r) {
    r[0] = l[0] * l[0] + l[1] * l[1] + b12 * l[0] * l[1] - a12;
    r[1] = l[0] * l[0] + l[2] * l[2] + b13 * l[0] * l[2] - a13;
    r[2] = l[1] * l[1] + l[2] * l[2] + b23 * l[1] * l[2] - a23;
}
__device__ void gauss_newton_refine_lambda(double *l, int iterations, double a12, double a13, double a23, double b12, double b13, double b23) {
    double res[3];
    gn_residual(l, a12, a13, a23, b12, b13, b23, res);
    for (int it = 0; it < iterations; it++) {
        if (l1n(res) < 1e-10) break;
        const double l1 = l[0], l2 = l[1], l3 = l[2];
        const double dr1dl1 = 2.0 * l1 + b12 * l2, dr1dl2 = 2.0 * l2 + b12 * l1, dr2dl1 = 2.0 * l1 + b13 * l3;
        const double dr2dl3 = 2.0 * l3 + b13 * l1, dr3dl2 = 2.0 * l2 + b23 * l3, dr3dl3 = 2.0 * l3 + b23 * l2;
        const double det = 1.0 / (-dr1dl1 * dr2dl3 * dr3dl2 - dr1dl2 * dr2dl1 * dr3dl3);
        const double J[9] = {-dr2dl3 * dr3dl2, -dr1dl2 * dr3dl3, dr1dl2 * dr2dl3,
                             -dr2dl1 * dr3dl3, dr1dl1 * dr3dl3, -dr1dl1 * dr2dl3,
                             dr2dl1 * dr3dl2, -dr1dl1 * dr3dl2, -dr1dl2 * dr2dl1};
        double ln[3], rn[3];
        for (int r = 0; r < 3; r++) ln[r] = l[r] - det * dot3(J + 3 * r, res);
        gn_residual(ln, a12, a13, a23, b12, b13, b23, rn);
        if (l1n(rn) > l1n(res)) break;
        for (int r = 0; r < 3; r++) { l[r] = ln[r]; res[r] = rn[r]; }
    }
}
__device__ bool inv3(const double *m, double *o) {
    const double d = det3(m);
    if (d == 0.0) return false;
    const double id = 1.0 / d;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return true;
}
// nalgebra Rotation3::from_matrix_eps(m, eps, max_iter, identity)
__device__ void rotation_from_matrix_eps(const double *m, double eps, int max_iter, double *rot) {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int it = 0; it < max_iter; it++) {
        double axis[3] = {0, 0, 0}, denom = 0.0;
        for (int c = 0; c < 3; c++) {
            const double rc[3] = {R[c], R[3 + c], R[6 + c]}, mc[3] = {m[c], m[3 + c], m[6 + c]};
            double x[3];
            cross3(rc, mc, x);
            for (int k = 0; k < 3; k++) axis[k] += x[k];
            denom += dot3(rc, mc);
        }
        const double sc = fabs(denom) + 2.220446049250313e-16;
        const double aa[3] = {axis[0] / sc, axis[1] / sc, axis[2] / sc};
        const double angle = norm3(aa);
        if (!(angle > eps)) break;
        const double u[3] = {aa[0] / angle, aa[1] / angle, aa[2] / angle};
        const double s = sin(angle), c = cos(angle), omc = 1.0 - c;
        const double Q[9] = {u[0] * u[0] + (1 - u[0] * u[0]) * c, u[0] * u[1] * omc - u[2] * s, u[0] * u[2] * omc + u[1] * s,
                             u[0] * u[1] * omc + u[2] * s, u[1] * u[1] + (1 - u[1] * u[1]) * c, u[1] * u[2] * omc - u[0] * s,
                             u[0] * u[2] * omc - u[1] * s, u[1] * u[2] * omc + u[0] * s, u[2] * u[2] + (1 - u[2] * u[2]) * c};
        mat3_mul(Q, R, R);
    }
    for (int i = 0; i < 9; i++) rot[i] = R[i];
}
__device__ int p3p(const double *bearings, const double *world, const uint32_t *idx, cvb_pose *out) {
    double wp[3][3];
    const double *y[3];
    for (int i = 0; i < 3; i++) {
        const double *w = world + 4 * (size_t)idx[i];
        if (w[3] == 0.0) return 0;
        for (int k = 0; k < 3; k++) wp[i][k] = w[k] / w[3];
        y[i] = bearings + 3 * (size_t)idx[i];
    }
    double d12[3], d13[3], d23[3], d12xd13[3];
    for (int k = 0; k < 3; k++) { d12[k] = wp[0][k] - wp[1][k]; d13[k] = wp[0][k] - wp[2][k]; d23[k] = wp[1][k] - wp[2][k]; }
    cross3(d12, d13, d12xd13);
    const double a12 = dot3(d12, d12), a13 = dot3(d13, d13), a23 = dot3(d23, d23);
    const double c12 = dot3(y[0], y[1]), c23 = dot3(y[1], y[2]), c31 = dot3(y[2], y[0]);
    const double blob = c12 * c23 * c31 - 1.0;
    const double s12_sqr = 1.0 - c12 * c12, s23_sqr = 1.0 - c23 * c23, s31_sqr = 1.0 - c31 * c31;
    const double b12 = -2.0 * c12, b13 = -2.0 * c31, b23 = -2.0 * c23;
    const double p3 = a13 * (a23 * s31_sqr - a13 * s23_sqr);
    const double p2 = 2.0 * blob * a23 * a13 + a13 * (2.0 * a12 + a13) * s23_sqr + a23 * (a23 - a12) * s31_sqr;
    const double p1 = a23 * (a13 - a23) * s12_sqr - a12 * a12 * s23_sqr - 2.0 * a12 * (blob * a23 + a13 * s23_sqr);
    const double p0 = a12 * (a12 * s23_sqr - a23 * s12_sqr);
    const double g = cube_root(p2 / p3, p1 / p3, p0 / p3);
    const double d0_00 = a23 * (1.0 - g), d0_01 = -(a23 * c12), d0_02 = a23 * c31 * g, d0_11 = a23 - a12 + a13 * g;
    const double d0_12 = -c23 * (a13 * g - a12), d0_22 = g * (a13 - a23) - a12;
    const double D0[9] = {d0_00, d0_01, d0_02, d0_01, d0_11, d0_12, d0_02, d0_12, d0_22};
    double Ev[9], ev[3];
    eigen_decomposition_singular(D0, Ev, ev);
    double lambdas[4][3];
    int nl = 0;
    const double eigen_ratio = sqrt(fmax(0.0, -ev[1] / ev[0]));
    for (int sgn = 0; sgn < 2; sgn++) {
        const double ratio = sgn ? -eigen_ratio : eigen_ratio;
        const double w2 = 1.0 / (ratio * Ev[1] - Ev[0]);
        const double w0 = w2 * (Ev[3] - ratio * Ev[4]);
        const double w1 = w2 * (Ev[6] - ratio * Ev[7]);
        const double a = 1.0 / ((a13 - a12) * w1 * w1 - a12 * b13 * w1 - a12);
        const double b = a * (a13 * b12 * w1 - a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13));
        const double c = a * ((a13 - a12) * w0 * w0 + a13 * b12 * w0 + a13);
        if (b * b - 4.0 * c >= 0.0) {
            double tau[2];
            root2real(b, c, &tau[0], &tau[1]);
            for (int k = 0; k < 2; k++) {
                if (tau[k] > 0.0) {
                    const double d = a23 / (tau[k] * (b23 + tau[k]) + 1.0);
                    if (d > 0.0) {
                        const double l2 = sqrt(d), l3 = tau[k] * l2, l1 = w0 * l2 + w1 * l3;
                        if (l1 >= 0.0 && nl < 4) { lambdas[nl][0] = l1; lambdas[nl][1] = l2; lambdas[nl][2] = l3; nl++; }
                    }
                }
            }
        }
    }
    const double X[9] = {d12[0], d13[0], d12xd13[0], d12[1], d13[1], d12xd13[1], d12[2], d13[2], d12xd13[2]};
    double Xi[9];
    if (!inv3(X, Xi)) return 0;
    for (int s = 0; s < nl; s++) {
        double l[3] = {lambdas[s][0], lambdas[s][1], lambdas[s][2]};
        gauss_newton_refine_lambda(l, 5, a12, a13, a23, b12, b13, b23);
        double ry1[3], ry2[3], ry3[3], yd1[3], yd2[3], yx[3];
        for (int k = 0; k < 3; k++) { ry1[k] = l[0] * y[0][k]; ry2[k] = l[1] * y[1][k]; ry3[k] = l[2] * y[2][k]; }
        for (int k = 0; k < 3; k++) { yd1[k] = ry1[k] - ry2[k]; yd2[k] = ry1[k] - ry3[k]; }
        cross3(yd1, yd2, yx);
        const double Y[9] = {yd1[0], yd2[0], yx[0], yd1[1], yd2[1], yx[1], yd1[2], yd2[2], yx[2]};
        double rot[9];
        mat3_mul(Y, Xi, rot);
        for (int k = 0; k < 3; k++) out[s].t[k] = ry1[k] - dot3(rot + 3 * k, wp[0]);
        rotation_from_matrix_eps(rot, 1e-12, 100, out[s].r);
    }
    return nl;
}

// ---- five-point (nister-stewenius/src/lib.rs:50-330).  nalgebra's full_piv_lu / complex_eigenvalues / try_svd are
// implemented as complete-pivoting elimination, Hessenberg + Francis double-shift QR, and one-sided Jacobi.
// `row0`: first eigenvector row used as (x, y, z, 1).  The reference takes rows 5..8 (`fixed_rows::<4>(5)`, lib.rs:229)
// although the monomial basis puts (x, y, z, 1) in rows 6..9, so its essentials violate the cubic constraints; row0 = 5
// reproduces the reference, row0 = 6 is the mathematically correct solver.
constexpr int FPN = 10;
// __noinline__: with every helper inlined into one five-point frame, nvcc 12.9 -O3 produced a wrong complete-pivoting
// elimination on sm_100a (the same text is right stand-alone and on the host); separate frames are bit-identical to the CPU.
__device__ __noinline__ bool lu_full_pivot_solve(const double *Ain, const double *Bin, double *X) {
    double A[FPN][FPN], B[FPN][FPN];
    int colperm[FPN];
    for (int i = 0; i < FPN; i++) for (int j = 0; j < FPN; j++) { A[i][j] = Ain[i * FPN + j]; B[i][j] = Bin[i * FPN + j]; }
    for (int i = 0; i < FPN; i++) colperm[i] = i;
    for (int k = 0; k < FPN; k++) {
        int pr = k, pc = k; double best = -1.0;
        for (int i = k; i < FPN; i++) for (int j = k; j < FPN; j++) if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
        if (best == 0.0) return false;
        if (pr != k) for (int j = 0; j < FPN; j++) { double t = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = t; t = B[k][j]; B[k][j] = B[pr][j]; B[pr][j] = t; }
        if (pc != k) { for (int i = 0; i < FPN; i++) { double t = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = t; } int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t; }
        for (int i = k + 1; i < FPN; i++) {
            const double f = A[i][k] / A[k][k];
            if (f == 0.0) continue;
            for (int j = k; j < FPN; j++) A[i][j] -= f * A[k][j];
            for (int j = 0; j < FPN; j++) B[i][j] -= f * B[k][j];
        }
    }
    double Y[FPN][FPN];
    for (int c = 0; c < FPN; c++)
        for (int i = FPN - 1; i >= 0; i--) {
            double v = B[i][c];
            for (int j = i + 1; j < FPN; j++) v -= A[i][j] * Y[j][c];
            Y[i][c] = v / A[i][i];
        }
    for (int i = 0; i < FPN; i++) for (int c = 0; c < FPN; c++) X[colperm[i] * FPN + c] = Y[i][c];
    return true;
}

__device__ __forceinline__ double fp_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }
__device__ __noinline__ bool real_eigenvalues10(const double *Ain, double *wr, double *wi) {
    const int n = FPN;
    double a[FPN][FPN];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i][j] = Ain[i * n + j];
    for (int m = 1; m < n - 1; m++) {
        double x = 0.0; int i = m;
        for (int j = m; j < n; j++) if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; i = j; }
        if (i != m) {
            for (int j = m - 1; j < n; j++) { double t = a[i][j]; a[i][j] = a[m][j]; a[m][j] = t; }
            for (int j = 0; j < n; j++) { double t = a[j][i]; a[j][i] = a[j][m]; a[j][m] = t; }
        }
        if (x != 0.0)
            for (i = m + 1; i < n; i++) {
                double y = a[i][m - 1];
                if (y != 0.0) {
                    y /= x; a[i][m - 1] = y;
                    for (int j = m; j < n; j++) a[i][j] -= y * a[m][j];
                    for (int j = 0; j < n; j++) a[j][m] += y * a[j][i];
                }
            }
    }
    for (int i = 2; i < n; i++) for (int j = 0; j < i - 1; j++) a[i][j] = 0.0;
    int nn = n - 1, l, its;
    double p = 0, q = 0, r = 0, s, t = 0.0, u, v, w, x, y, z, anorm = 0.0;
    for (int i = 0; i < n; i++) for (int j = (i > 0 ? i - 1 : 0); j < n; j++) anorm += fabs(a[i][j]);
    while (nn >= 0) {
        its = 0;
        do {
            for (l = nn; l >= 1; l--) {
                s = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
                if (s == 0.0) s = anorm;
                if (fabs(a[l][l - 1]) + s == s) { a[l][l - 1] = 0.0; break; }
            }
            x = a[nn][nn];
            if (l == nn) { wr[nn] = x + t; wi[nn--] = 0.0; }
            else {
                y = a[nn - 1][nn - 1]; w = a[nn][nn - 1] * a[nn - 1][nn];
                if (l == nn - 1) {
                    p = 0.5 * (y - x); q = p * p + w; z = sqrt(fabs(q)); x += t;
                    if (q >= 0.0) {
                        z = p + fp_sign(z, p);
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else { wr[nn - 1] = wr[nn] = x + p; wi[nn] = z; wi[nn - 1] = -z; }
                    nn -= 2;
                } else {
                    if (its == 60) return false;
                    if (its == 10 || its == 20) {
                        t += x;
                        for (int i = 0; i <= nn; i++) a[i][i] -= x;
                        s = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]);
                        y = x = 0.75 * s; w = -0.4375 * s * s;
                    }
                    ++its;
                    int m;
                    for (m = nn - 2; m >= l; m--) {
                        z = a[m][m]; r = x - z; s = y - z;
                        p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
                        q = a[m + 1][m + 1] - z - r - s;
                        r = a[m + 2][m + 1];
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
                        v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    for (int i = m + 2; i <= nn; i++) { a[i][i - 2] = 0.0; if (i != m + 2) a[i][i - 3] = 0.0; }
                    for (int k = m; k <= nn - 1; k++) {
                        if (k != m) {
                            p = a[k][k - 1]; q = a[k + 1][k - 1]; r = 0.0;
                            if (k != nn - 1) r = a[k + 2][k - 1];
                            if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        if ((s = fp_sign(sqrt(p * p + q * q + r * r), p)) != 0.0) {
                            if (k == m) { if (l != m) a[k][k - 1] = -a[k][k - 1]; }
                            else a[k][k - 1] = -s * x;
                            p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
                            for (int j = k; j <= nn; j++) {
                                p = a[k][j] + q * a[k + 1][j];
                                if (k != nn - 1) { p += r * a[k + 2][j]; a[k + 2][j] -= p * z; }
                                a[k + 1][j] -= p * y; a[k][j] -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; i++) {
                                p = x * a[i][k] + y * a[i][k + 1];
                                if (k != nn - 1) { p += z * a[i][k + 2]; a[i][k + 2] -= p * r; }
                                a[i][k + 1] -= p * q; a[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return true;
}

__device__ __noinline__ bool min_right_singular_vector10(const double *Min, double eps, int max_sweeps, double *vec, double *smin) {
    double U[FPN][FPN], V[FPN][FPN];
    for (int i = 0; i < FPN; i++) for (int j = 0; j < FPN; j++) { U[i][j] = Min[i * FPN + j]; V[i][j] = i == j ? 1.0 : 0.0; }
    bool converged = false;
    for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
        converged = true;
        for (int p = 0; p < FPN - 1; p++)
            for (int q = p + 1; q < FPN; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < FPN; i++) { alpha += U[i][p] * U[i][p]; beta += U[i][q] * U[i][q]; gamma += U[i][p] * U[i][q]; }
                if (gamma == 0.0 || fabs(gamma) <= eps * sqrt(alpha * beta)) continue;
                converged = false;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < FPN; i++) {
                    const double up = U[i][p], uq = U[i][q];
                    U[i][p] = c * up - s * uq; U[i][q] = s * up + c * uq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                }
            }
    }
    if (!converged) return false;
    int best = 0; double bn = -1.0;
    for (int j = 0; j < FPN; j++) {
        double nn = 0; for (int i = 0; i < FPN; i++) nn += U[i][j] * U[i][j];
        if (bn < 0.0 || nn < bn) { bn = nn; best = j; }
    }
    *smin = sqrt(bn);
    for (int i = 0; i < FPN; i++) vec[i] = V[i][best];
    return true;
}

enum { BXXX = 0, BXXY, BXYY, BYYY, BXXZ, BXYZ, BYYZ, BXZZ, BYZZ, BZZZ, BXX, BXY, BYY, BXZ, BYZ, BZZ, BX, BY, BZ, B1 };
__device__ void fp_o1(const double *a, const double *b, double *r) {
    for (int i = 0; i < 20; i++) r[i] = 0.0;
    r[BXX] = a[0] * b[0]; r[BXY] = a[0] * b[1] + a[1] * b[0]; r[BXZ] = a[0] * b[2] + a[2] * b[0];
    r[BYY] = a[1] * b[1]; r[BYZ] = a[1] * b[2] + a[2] * b[1]; r[BZZ] = a[2] * b[2];
    r[BX] = a[0] * b[3] + a[3] * b[0]; r[BY] = a[1] * b[3] + a[3] * b[1]; r[BZ] = a[2] * b[3] + a[3] * b[2]; r[B1] = a[3] * b[3];
}
__device__ void fp_o2(const double *a, const double *b, double *r) {
    r[BXXX] = a[BXX] * b[0];
    r[BXXY] = a[BXX] * b[1] + a[BXY] * b[0];
    r[BXXZ] = a[BXX] * b[2] + a[BXZ] * b[0];
    r[BXYY] = a[BXY] * b[1] + a[BYY] * b[0];
    r[BXYZ] = a[BXY] * b[2] + a[BYZ] * b[0] + a[BXZ] * b[1];
    r[BXZZ] = a[BXZ] * b[2] + a[BZZ] * b[0];
    r[BYYY] = a[BYY] * b[1];
    r[BYYZ] = a[BYY] * b[2] + a[BYZ] * b[1];
    r[BYZZ] = a[BYZ] * b[2] + a[BZZ] * b[1];
    r[BZZZ] = a[BZZ] * b[2];
    r[BXX] = a[BXX] * b[3] + a[BX] * b[0];
    r[BXY] = a[BXY] * b[3] + a[BX] * b[1] + a[BY] * b[0];
    r[BXZ] = a[BXZ] * b[3] + a[BX] * b[2] + a[BZ] * b[0];
    r[BYY] = a[BYY] * b[3] + a[BY] * b[1];
    r[BYZ] = a[BYZ] * b[3] + a[BY] * b[2] + a[BZ] * b[1];
    r[BZZ] = a[BZZ] * b[3] + a[BZ] * b[2];
    r[BX] = a[BX] * b[3] + a[B1] * b[0];
    r[BY] = a[BY] * b[3] + a[B1] * b[1];
    r[BZ] = a[BZ] * b[3] + a[B1] * b[2];
    r[B1] = a[B1] * b[3];
}

// essential matrix -> the four candidate poses (cv-pinhole/src/essential.rs:114-162,217-231)
__device__ __noinline__ int essential_poses(const double *E, cvb_pose *out) {
    double U[9], Vt[9];
    if (!svd3(E, 1e-12, 1000, U, Vt)) return 0;
    if (det3(U) < 0.0) for (int r = 0; r < 3; r++) U[r * 3 + 2] *= -1.0;
    if (det3(Vt) < 0.0) for (int c = 0; c < 3; c++) Vt[6 + c] *= -1.0;
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double UW[9], Ra[9], Rb[9];
    mat3_mul(U, W, UW); mat3_mul(UW, Vt, Ra);
    mat3_mul(U, Wt, UW); mat3_mul(UW, Vt, Rb);
    const double t[3] = {U[2], U[5], U[8]};
    for (int k = 0; k < 4; k++) {
        for (int i = 0; i < 9; i++) out[k].r[i] = (k & 1) ? Rb[i] : Ra[i];
        for (int r = 0; r < 3; r++) out[k].t[r] = (k & 2) ? -t[r] : t[r];
    }
    return 4;
}

__device__ int five_point(const double *a, const double *b, const uint32_t *idx, int row0, cvb_pose *out) {
    double A[5][9], EE[81], d[9], V[81];
    for (int i = 0; i < 5; i++) {
        const double *pa = a + 3 * (size_t)idx[i], *pb = b + 3 * (size_t)idx[i];
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) A[i][3 * j + k] = pa[j] * pb[k];
    }
    for (int r = 0; r < 9; r++)
        for (int c = 0; c < 9; c++) { double s = 0; for (int i = 0; i < 5; i++) s += A[i][r] * A[i][c]; EE[r * 9 + c] = s; }
    if (!sym_eigen<9>(EE, 1e-12, 1000, d, V)) return 0;
    int src[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
    for (int i = 1; i < 9; i++) { int x = src[i], j = i; while (j > 0 && d[src[j - 1]] > d[x]) { src[j] = src[j - 1]; j--; } src[j] = x; }
    int nullity = -1;
    for (int i = 0; i < 9; i++) if (d[src[i]] > 1e-12) { nullity = i; break; }
    if (nullity != 4) return 0;
    double eb[9][4];
    for (int c = 0; c < 4; c++) for (int r = 0; r < 9; r++) eb[r][c] = V[r * 9 + src[c]];
    double ep[3][3][4];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 4; k++) ep[i][j][k] = eb[3 * i + j][k];
    double M[10][20], t1[20], t2[20], t3[20], acc[20];
    {
        const int ia[3][2] = {{1, 2}, {2, 0}, {0, 1}};
        for (int k = 0; k < 20; k++) acc[k] = 0.0;
        for (int c = 0; c < 3; c++) {
            const int p = ia[c][0], q = ia[c][1];
            fp_o1(ep[0][p], ep[1][q], t1); fp_o1(ep[0][q], ep[1][p], t2);
            for (int k = 0; k < 20; k++) t1[k] -= t2[k];
            fp_o2(t1, ep[2][c], t3);
            for (int k = 0; k < 20; k++) acc[k] += t3[k];
        }
        for (int k = 0; k < 20; k++) M[0][k] = acc[k];
    }
    double eet[3][3][20], L[3][3][20];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            if (i <= j) {
                fp_o1(ep[i][0], ep[j][0], t1); fp_o1(ep[i][1], ep[j][1], t2); fp_o1(ep[i][2], ep[j][2], t3);
                for (int k = 0; k < 20; k++) eet[i][j][k] = t1[k] + t2[k] + t3[k];
            } else for (int k = 0; k < 20; k++) eet[i][j][k] = eet[j][i][k];
        }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 20; k++) L[i][j][k] = eet[i][j][k];
    for (int k = 0; k < 20; k++) {
        const double tr = 0.5 * (eet[0][0][k] + eet[1][1][k] + eet[2][2][k]);
        for (int i = 0; i < 3; i++) L[i][i][k] -= tr;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            fp_o2(L[i][0], ep[0][j], t1); fp_o2(L[i][1], ep[1][j], t2); fp_o2(L[i][2], ep[2][j], t3);
            for (int k = 0; k < 20; k++) M[1 + i * 3 + j][k] = t1[k] + t2[k] + t3[k];
        }
    double Cl[100], Cr[100], X[100];
    for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) { Cl[i * 10 + j] = M[i][j]; Cr[i * 10 + j] = M[i][10 + j]; }
    if (!lu_full_pivot_solve(Cl, Cr, X)) return 0;
    double At[100];
    for (int i = 0; i < 100; i++) At[i] = 0.0;
    for (int j = 0; j < 10; j++) {
        At[0 * 10 + j] = X[0 * 10 + j]; At[1 * 10 + j] = X[1 * 10 + j]; At[2 * 10 + j] = X[2 * 10 + j];
        At[3 * 10 + j] = X[4 * 10 + j]; At[4 * 10 + j] = X[5 * 10 + j]; At[5 * 10 + j] = X[7 * 10 + j];
    }
    At[6 * 10 + 0] = -1.0; At[7 * 10 + 1] = -1.0; At[8 * 10 + 3] = -1.0; At[9 * 10 + 6] = -1.0;
    double wr[10], wi[10];
    if (!real_eigenvalues10(At, wr, wi)) return 0;
    int n = 0;
    for (int i = 0; i < 10; i++) {
        if (wi[i] != 0.0) continue;
        double *Mx = Cl;   // reuse
        double vec[10], smin;
        for (int k = 0; k < 100; k++) Mx[k] = At[k];
        for (int k = 0; k < 10; k++) Mx[k * 10 + k] -= wr[i];
        if (!min_right_singular_vector10(Mx, 1e-15, 1000, vec, &smin)) continue;
        if (!(smin < 1e-12)) continue;
        double ev[9], E[9];
        for (int r = 0; r < 9; r++) ev[r] = eb[r][0] * vec[row0] + eb[r][1] * vec[row0 + 1] + eb[r][2] * vec[row0 + 2] + eb[r][3] * vec[row0 + 3];
        for (int k = 0; k < 9; k++) E[(k % 3) * 3 + (k / 3)] = ev[k];
        n += essential_poses(E, out + n);
    }
    return n;
}

// ------------------------------------------------------------------------------------------ kernels
template <int KIND>
__global__ void __launch_bounds__(128) k_estimate(const double *__restrict__ a, const double *__restrict__ b,
                                                  const uint32_t *__restrict__ samples, uint32_t H, cvb_pose *poses,
                                                  uint8_t *nposes, int row0) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    if (KIND == 2) {   // five-point: up to 40 poses per sample, written straight to global memory
        nposes[h] = (uint8_t)five_point(a, b, samples + (size_t)h * 5, row0, poses + (size_t)h * 40);
        return;
    }
    cvb_pose out[4];
    int n;
    if (KIND == 0) n = eight_point(a, b, samples + (size_t)h * 8, out);
    else n = p3p(a, b, samples + (size_t)h * 3, out);
    for (int k = 0; k < n; k++) poses[(size_t)h * 4 + k] = out[k];
    nposes[h] = (uint8_t)n;
}

// one thread per (pose, datum).  MODE 0: write residuals; MODE 1: write bit-packed inlier masks
// (residual < thr) over data [i0, i1): mask word w of pose p at masks[p * words + w].
template <int KIND, int MODE>
__global__ void __launch_bounds__(256) k_residuals(const cvb_pose *__restrict__ poses, uint32_t m,
                                                   const double *__restrict__ a, const double *__restrict__ b,
                                                   uint32_t i0, uint32_t i1, double thr, double *__restrict__ out,
                                                   uint32_t out_stride, uint32_t *__restrict__ masks, uint32_t words) {
    const uint32_t p = blockIdx.y;
    const uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < i1;
    double r = 3.0;
    if (in) {
        const cvb_pose P = poses[p];
        r = KIND == 0 ? residual_c2c(P, a + 3 * (size_t)i, b + 3 * (size_t)i) : residual_w2c(P, a + 3 * (size_t)i, b + 4 * (size_t)i);
    }
    if (MODE == 0) { if (in) out[(size_t)p * out_stride + i] = r; }
    else {
        const unsigned bits = __ballot_sync(0xffffffffu, in && r < thr);
        // warps behind the data count own no mask word (a row has cdiv(i1 - i0, 32) words, the CTA covers 8)
        if ((threadIdx.x & 31) == 0 && ((i - i0) >> 5) < words) masks[(size_t)p * words + ((i - i0) >> 5)] = bits;
    }
}

// cv-geom/src/triangulation.rs:82-130: n >= 2 observations (WorldToCamera pose, bearing) -> homogeneous world point
__device__ bool triangulate_linear_eigen(const cvb_pose *poses, const double *bearings, uint32_t n, double *p) {
    if (n < 2) return false;
    double A[16], d[4], V[16];
    for (int i = 0; i < 16; i++) A[i] = 0.0;
    for (uint32_t i = 0; i < n; i++) design_add(poses[i].r, poses[i].t, bearings + 3 * (size_t)i, A);
    if (!sym_eigen<4>(A, 1e-12, 1000, d, V)) return false;
    int best = 0;
    for (int i = 1; i < 4; i++)
        if (d[i] < d[best]) best = i;
    p[0] = V[best]; p[1] = V[4 + best]; p[2] = V[8 + best]; p[3] = V[12 + best];
    from_homogeneous(p);
    if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]) && isfinite(p[3]))) return false;
    for (uint32_t i = 0; i < n; i++) {
        const double *bb = bearings + 3 * (size_t)i, *R = poses[i].r;
        const double wb[3] = {R[0] * bb[0] + R[3] * bb[1] + R[6] * bb[2], R[1] * bb[0] + R[4] * bb[1] + R[7] * bb[2],
                              R[2] * bb[0] + R[5] * bb[1] + R[8] * bb[2]};
        if (signbit(dot3(wb, p))) return false;
    }
    return true;
}
// one thread per landmark
__global__ void __launch_bounds__(128) k_triangulate(const cvb_pose *__restrict__ poses, const double *__restrict__ bearings,
                                                     const uint32_t *__restrict__ offsets, uint32_t L, double *__restrict__ xyzw,
                                                     uint8_t *__restrict__ ok) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const uint32_t o0 = offsets[l], o1 = offsets[l + 1];
    double p[4] = {0, 0, 0, 0};
    const bool good = triangulate_linear_eigen(poses + o0, bearings + 3 * (size_t)o0, o1 - o0, p);
    ok[l] = good ? 1 : 0;
    for (int i = 0; i < 4; i++) xyzw[(size_t)l * 4 + i] = good ? p[i] : 0.0;
}

// cv-sfm keeps calibrated bearings per feature (CameraModel::calibrate, cv-pinhole/src/lib.rs:108-116, on
// akaze::KeyPoint's ImagePoint, akaze/src/lib.rs:95-99); a FeatureMatch is the bearing pair of a match (cv-sfm/src/lib.rs:1400).
// One thread per match: both bearings in f64 with the reference's operation order (no distortion: k1 = 0).
__device__ __forceinline__ void calibrate_px(const cvb_intrinsics K, double px, double py, double *o) {
    const double y = (py - K.cy) / K.fy;
    const double x = (px - K.cx - K.skew * y) / K.fx;
    const double n = sqrt(x * x + y * y + 1.0);
    o[0] = x / n; o[1] = y / n; o[2] = 1.0 / n;
}
__global__ void __launch_bounds__(256) k_pair_bearings(const cvb_keypoint *__restrict__ kpa, const cvb_keypoint *__restrict__ kpb,
                                                       const uint32_t *__restrict__ pairs, const uint32_t *__restrict__ npairs,
                                                       uint32_t cap, cvb_intrinsics K, double *__restrict__ a, double *__restrict__ b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*npairs, cap)) return;
    const cvb_keypoint ka = kpa[pairs[2 * i]], kb = kpb[pairs[2 * i + 1]];
    calibrate_px(K, (double)ka.x, (double)ka.y, a + 3 * (size_t)i);
    calibrate_px(K, (double)kb.x, (double)kb.y, b + 3 * (size_t)i);
}

#include "arrsac_dev.cuh"

// ------------------------------------------------------------------------------------------ post-consensus refinement
// cv-optimize single_view_simple_optimize_l2 / three_view_{simple,adaptive}_optimize_l2 with cv-geom's epipolar gradients,
// and cv-sfm's robustness checks.  One CTA per problem iterates to completion on the device: every iteration the threads
// evaluate the per-landmark tangents, a fixed-shape reduction tree adds them (warp shuffles, then the warps in order; the
// reference adds them in landmark order, so sums agree to rounding, not bit for bit), thread 0 replays the reference's
// bookkeeping (patience counter, pose update) and publishes the pose for the next iteration.
__device__ __forceinline__ void rotv(const double *R, const double *v, double *o) { for (int r = 0; r < 3; r++) o[r] = dot3(R + 3 * r, v); }
__device__ __forceinline__ bool any_nan3(const double *v) { return isnan(v[0]) || isnan(v[1]) || isnan(v[2]); }
__device__ __forceinline__ void normalize3(const double *v, double *o) { const double n = norm3(v); o[0] = v[0] / n; o[1] = v[1] / n; o[2] = v[2] / n; }
// Se3TangentSpace::new (cv-core/src/so3.rs:23-34)
__device__ __forceinline__ void tangent_new(double *t, double *r) {
    if (any_nan3(t)) t[0] = t[1] = t[2] = 0.0;
    if (any_nan3(r)) r[0] = r[1] = r[2] = 0.0;
}
// nalgebra Rotation3::from_scaled_axis
__device__ void rot_from_scaled_axis(const double *v, double *R) {
    const double angle = norm3(v);
    if (angle == 0.0) { for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
    const double ux = v[0] / angle, uy = v[1] / angle, uz = v[2] / angle;
    const double sqx = ux * ux, sqy = uy * uy, sqz = uz * uz, sn = sin(angle), c = cos(angle), omc = 1.0 - c;
    R[0] = sqx + (1.0 - sqx) * c; R[1] = ux * uy * omc - uz * sn; R[2] = ux * uz * omc + uy * sn;
    R[3] = ux * uy * omc + uz * sn; R[4] = sqy + (1.0 - sqy) * c; R[5] = uy * uz * omc - ux * sn;
    R[6] = ux * uz * omc - uy * sn; R[7] = uy * uz * omc + ux * sn; R[8] = sqz + (1.0 - sqz) * c;
}
// pose <- Se3TangentSpace{trans, rot}.isometry() * pose (so3.rs:57-60)
__device__ void apply_delta(const double *trans, const double *rot, cvb_pose *P) {
    double Rd[9], td[3], Rn[9], tn[3];
    rot_from_scaled_axis(rot, Rd);
    rotv(Rd, trans, td);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Rn[3 * r + c] = Rd[3 * r] * P->r[c] + Rd[3 * r + 1] * P->r[3 + c] + Rd[3 * r + 2] * P->r[6 + c];
    rotv(Rd, P->t, tn);
    for (int r = 0; r < 3; r++) tn[r] = td[r] + tn[r];
    for (int i = 0; i < 9; i++) P->r[i] = Rn[i];
    for (int i = 0; i < 3; i++) P->t[i] = tn[i];
}
__device__ void pose_inverse(const cvb_pose &P, cvb_pose *o) {
    const double nt[3] = {-P.t[0], -P.t[1], -P.t[2]};
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = P.r[3 * c + r];
    rotv(R, nt, o->t);
    for (int i = 0; i < 9; i++) o->r[i] = R[i];
}
// cv-geom/src/epipolar.rs:193-198
__device__ void world_pose_gradient(const double *translation, const double *b, double *tg, double *rg) {
    const double d = dot3(translation, b);
    double nt[3];
    for (int i = 0; i < 3; i++) tg[i] = d * b[i] - translation[i];
    normalize3(translation, nt);
    cross3(nt, b, rg);
    tangent_new(tg, rg);
}
// cv-optimize/src/single_view_optimizer.rs:4-14
__device__ bool landmark_delta(const cvb_pose &P, const double *bearing, const double *world, double *tg, double *rg) {
    double q[4];
    pose_apply(P, world, q);
    from_homogeneous(q);
    if (q[3] == 0.0) return false;
    const double p[3] = {q[0] / q[3], q[1] / q[3], q[2] / q[3]};
    world_pose_gradient(p, bearing, tg, rg);
    return true;
}
// epipolar.rs:8-50
__device__ bool sine_l1_point(const double *t, const double *a_in, const double *b_in, double *p) {
    double ca[3], cb[3], na[3], nb[3], a[3], b[3];
    cross3(a_in, t, ca); const double can = norm3(ca); for (int i = 0; i < 3; i++) na[i] = ca[i] / can;
    cross3(b_in, t, cb); const double cbn = norm3(cb); for (int i = 0; i < 3; i++) nb[i] = cb[i] / cbn;
    for (int i = 0; i < 3; i++) { a[i] = a_in[i]; b[i] = b_in[i]; }
    if (can < cbn) { const double d = dot3(a_in, nb); double v[3]; for (int i = 0; i < 3; i++) v[i] = a_in[i] - d * nb[i]; normalize3(v, a); }
    else { const double d = dot3(b_in, na); double v[3]; for (int i = 0; i < 3; i++) v[i] = b_in[i] - d * na[i]; normalize3(v, b); }
    double z[3], tb[3];
    cross3(a, b, z); cross3(t, b, tb);
    double q[4] = {a[0], a[1], a[2], dot3(z, z) / dot3(z, tb)};
    from_homogeneous(q);
    for (int i = 0; i < 4; i++) if (!isfinite(q[i])) return false;
    if (signbit(dot3(q, a)) || signbit(dot3(q, b))) return false;
    if (q[3] == 0.0) return false;
    for (int i = 0; i < 3; i++) p[i] = q[i] / q[3];
    return true;
}
// epipolar.rs:53-71
__device__ void rotation_gradient(const double *t, const double *a, const double *b, double *o) {
    double ca[3], cb[3], na[3], nb[3];
    cross3(a, t, ca); cross3(b, t, cb);
    normalize3(ca, na); normalize3(cb, nb);
    cross3(nb, na, o);
}
// epipolar.rs:85-176: out = [first.t, first.r, second.t, second.r]
__device__ void three_view_gradients(const double *c, const double *f, const double *ftoc, const double *s, const double *stoc, double *out) {
    double stof[3], rcf[3], rcs[3], rfs[3], p[3], q[3], tf[3] = {0, 0, 0}, ts[3] = {0, 0, 0}, tc[3] = {0, 0, 0}, neg[3];
    for (int i = 0; i < 3; i++) stof[i] = stoc[i] - ftoc[i];
    rotation_gradient(ftoc, c, f, rcf); rotation_gradient(stoc, c, s, rcs); rotation_gradient(stof, f, s, rfs);
    double *ft = out, *fr = out + 3, *st = out + 6, *sr = out + 9;
    for (int i = 0; i < 3; i++) {
        fr[i] = rcf[i] * (2.0 / 3.0) + (-rfs[i]) * (1.0 / 3.0);
        sr[i] = rcs[i] * (2.0 / 3.0) + rfs[i] * (1.0 / 3.0);
    }
    for (int i = 0; i < 3; i++) neg[i] = -stoc[i];
    if (sine_l1_point(neg, c, s, p)) { for (int i = 0; i < 3; i++) q[i] = p[i] - ftoc[i]; const double d = dot3(q, f); for (int i = 0; i < 3; i++) tf[i] = q[i] - d * f[i]; }
    for (int i = 0; i < 3; i++) neg[i] = -ftoc[i];
    if (sine_l1_point(neg, c, f, p)) { for (int i = 0; i < 3; i++) q[i] = p[i] - stoc[i]; const double d = dot3(q, s); for (int i = 0; i < 3; i++) ts[i] = q[i] - d * s[i]; }
    for (int i = 0; i < 3; i++) neg[i] = -stof[i];
    if (sine_l1_point(neg, f, s, p)) { for (int i = 0; i < 3; i++) q[i] = p[i] + ftoc[i]; const double d = dot3(q, c); for (int i = 0; i < 3; i++) tc[i] = d * c[i] - q[i]; }
    for (int i = 0; i < 3; i++) {
        ft[i] = tf[i] * (2.0 / 3.0) + tc[i] * (1.0 / 3.0);
        st[i] = ts[i] * (2.0 / 3.0) + tc[i] * (1.0 / 3.0);
    }
    tangent_new(ft, fr); tangent_new(st, sr);
}
// epipolar.rs:200-232
__device__ double epipolar_loss(const double *t, const double *a, const double *b) {
    double ca[3], cb[3];
    cross3(a, t, ca); cross3(b, t, cb);
    const double na2 = dot3(ca, ca), nb2 = dot3(cb, cb);
    double res;
    if (na2 < nb2) { const double sc = 1.0 / sqrt(nb2); const double v[3] = {cb[0] * sc, cb[1] * sc, cb[2] * sc}; res = fabs(dot3(a, v)); }
    else { const double sc = 1.0 / sqrt(na2); const double v[3] = {ca[0] * sc, ca[1] * sc, ca[2] * sc}; res = fabs(dot3(b, v)); }
    if (isnan(res) || signbit(dot3(a, b))) return 1.0;
    return res;
}

constexpr int OPT_NT = 512, OPT_WARPS = OPT_NT / 32;
// sums acc[0..NV) over the CTA into s_out[0..NV) (valid for thread 0 after the call); fixed tree -> run-to-run reproducible
template <int NV>
__device__ __forceinline__ void block_sum(double *acc, double *s_red) {
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < NV; k++) s_red[w * NV + k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < NV; k++) {
            double s = s_red[k];
            for (int ww = 1; ww < OPT_WARPS; ww++) s += s_red[ww * NV + k];
            acc[k] = s;
        }
}

// single_view_optimizer.rs:80-135, one CTA per (pose, landmark list)
__global__ void __launch_bounds__(OPT_NT) k_single_view_opt(const cvb_pose *__restrict__ poses_in, const double *__restrict__ bearings,
                                                            const double *__restrict__ world, const uint32_t *__restrict__ offsets,
                                                            double rate, uint32_t iterations, cvb_pose *__restrict__ poses_out,
                                                            uint32_t *__restrict__ updates_out) {
    __shared__ cvb_pose P;
    __shared__ double s_red[OPT_WARPS * 6];
    __shared__ int s_stop;
    const uint32_t b = blockIdx.x, o0 = offsets[b], n = offsets[b + 1] - o0;
    if (threadIdx.x == 0) { P = poses_in[b]; s_stop = 0; }
    __syncthreads();
    double best_t = INFINITY, best_r = INFINITY;
    uint32_t no_improve = 0, updates = 0;
    const double inv_len = 1.0 / (double)n;
    if (n > 0)
        for (uint32_t it = 0; it < iterations; it++) {
            double acc[6] = {0, 0, 0, 0, 0, 0}, tg[3], rg[3];
            const cvb_pose Pl = P;
            for (uint32_t i = threadIdx.x; i < n; i += OPT_NT)
                if (landmark_delta(Pl, bearings + 3 * (size_t)(o0 + i), world + 4 * (size_t)(o0 + i), tg, rg))
                    for (int k = 0; k < 3; k++) { acc[k] += tg[k]; acc[3 + k] += rg[k]; }
            block_sum<6>(acc, s_red);
            if (threadIdx.x == 0) {
                double dt[3], dr[3];
                for (int k = 0; k < 3; k++) { dt[k] = (acc[k] * inv_len) * rate; dr[k] = (acc[3 + k] * inv_len) * rate; }
                no_improve++;
                const double t = norm3(acc), r = norm3(acc + 3);
                if (best_t > t) { best_t = t; no_improve = 0; }
                if (best_r > r) { best_r = r; no_improve = 0; }
                if (no_improve >= 50) s_stop = 1;
                else {
                    apply_delta(dt, dr, &P); updates++;
                    if (it == iterations - 1) s_stop = 1;
                }
            }
            __syncthreads();
            if (s_stop) break;
        }
    if (threadIdx.x == 0) { poses_out[b] = P; updates_out[b] = updates; }
}

// three_view_optimizer.rs:126-272, one CTA per (pose pair, observation triples); obs = [centre, first, second] bearings
__global__ void __launch_bounds__(OPT_NT) k_three_view_opt(const cvb_pose *__restrict__ poses_in, const double *__restrict__ obs,
                                                           const uint32_t *__restrict__ offsets, int adaptive, double rate,
                                                           uint32_t iterations, cvb_pose *__restrict__ poses_out,
                                                           uint32_t *__restrict__ updates_out) {
    __shared__ cvb_pose P[2];
    __shared__ double s_red[OPT_WARPS * 16];
    __shared__ int s_stop;
    const uint32_t b = blockIdx.x, o0 = offsets[b], n = offsets[b + 1] - o0;
    if (threadIdx.x == 0) {
        if (n > 0) { pose_inverse(poses_in[2 * b], &P[0]); pose_inverse(poses_in[2 * b + 1], &P[1]); }
        s_stop = 0;
    }
    __syncthreads();
    double best[2][2] = {{INFINITY, INFINITY}, {INFINITY, INFINITY}};
    uint32_t no_improve = 0, updates = 0;
    const double inv_len = 1.0 / (double)n;
    if (n > 0)
        for (uint32_t it = 0; it < iterations; it++) {
            double acc[16], g[12];
            for (int k = 0; k < 16; k++) acc[k] = 0.0;
            const cvb_pose P0 = P[0], P1 = P[1];
            for (uint32_t i = threadIdx.x; i < n; i += OPT_NT) {
                const double *o = obs + 9 * (size_t)(o0 + i);
                double f[3], s[3];
                rotv(P0.r, o + 3, f); rotv(P1.r, o + 6, s);
                three_view_gradients(o, f, P0.t, s, P1.t, g);
                for (int k = 0; k < 12; k++) acc[k] += g[k];
                if (adaptive) { acc[12] += norm3(g); acc[13] += norm3(g + 3); acc[14] += norm3(g + 6); acc[15] += norm3(g + 9); }
            }
            block_sum<16>(acc, s_red);
            if (threadIdx.x == 0) {
                double d[12];
                bool stop = false;
                if (!adaptive) {
                    const double sc = inv_len * rate;
                    for (int k = 0; k < 12; k++) d[k] = acc[k] * sc;
                    no_improve++;
                    for (int v = 0; v < 2; v++) {
                        const double t = norm3(acc + 6 * v), r = norm3(acc + 6 * v + 3);
                        if (best[v][0] > t) { best[v][0] = t; no_improve = 0; }
                        if (best[v][1] > r) { best[v][1] = r; no_improve = 0; }
                    }
                    stop = no_improve >= 50;
                } else {
                    for (int v = 0; v < 2; v++) {
                        double l2[6];
                        for (int k = 0; k < 6; k++) l2[k] = acc[6 * v + k] * inv_len;
                        const double tstd = acc[12 + 2 * v] * inv_len, rstd = acc[13 + 2 * v] * inv_len;
                        double trate = norm3(l2) / tstd, rrate = norm3(l2 + 3) / rstd;
                        if (!isfinite(trate)) trate = 0.0;
                        if (!isfinite(rrate)) rrate = 0.0;
                        for (int k = 0; k < 3; k++) { d[6 * v + k] = l2[k] * trate; d[6 * v + 3 + k] = l2[3 + k] * rrate; }
                    }
                }
                if (stop) s_stop = 1;
                else {
                    apply_delta(d, d + 3, &P[0]); apply_delta(d + 6, d + 9, &P[1]); updates++;
                    if (it == iterations - 1) s_stop = 1;
                }
            }
            __syncthreads();
            if (s_stop) break;
        }
    if (threadIdx.x == 0) {
        if (n > 0) { pose_inverse(P[0], &poses_out[2 * b]); pose_inverse(P[1], &poses_out[2 * b + 1]); }
        else { poses_out[2 * b] = poses_in[2 * b]; poses_out[2 * b + 1] = poses_in[2 * b + 1]; }
        updates_out[b] = updates;
    }
}

__device__ void pose_mul(const cvb_pose &A, const cvb_pose &B, cvb_pose *o) {
    for (int i = 0; i < 3; i++)
        for (int c = 0; c < 3; c++) o->r[3 * i + c] = A.r[3 * i] * B.r[c] + A.r[3 * i + 1] * B.r[3 + c] + A.r[3 * i + 2] * B.r[6 + c];
    double sh[3];
    rotv(A.r, B.t, sh);
    for (int i = 0; i < 3; i++) o->t[i] = A.t[i] + sh[i];
}
__device__ double transformed_cosine_distance(const cvb_pose &P, const double *point_h, const double *bearing) {
    double q[4];
    pose_apply(P, point_h, q);
    from_homogeneous(q);
    return 1.0 - dot3(q, bearing);
}
// cv-sfm/src/lib.rs:2570-2620 observation_loss of every observation; one thread per landmark
__global__ void __launch_bounds__(128) k_observation_losses(const cvb_pose *__restrict__ poses, const double *__restrict__ bearings,
                                                            const uint32_t *__restrict__ offsets, uint32_t L, double *__restrict__ loss) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const uint32_t o0 = offsets[l], n = offsets[l + 1] - o0;
    const cvb_pose *P = poses + o0;
    const double *B = bearings + 3 * (size_t)o0;
    if (n == 0) return;
    if (n == 1) { loss[o0] = 2.0; return; }
    if (n == 2) {
        cvb_pose inv, tot;
        double fb[3];
        pose_inverse(P[0], &inv); pose_mul(P[1], inv, &tot);
        rotv(tot.r, B, fb);
        const double v = 1.0 - cos(asin(epipolar_loss(tot.t, fb, B + 3)));
        loss[o0] = v; loss[o0 + 1] = v;
        return;
    }
    double p[4];
    const bool ok = triangulate_linear_eigen(P, B, n, p);
    for (uint32_t i = 0; i < n; i++) loss[o0 + i] = ok ? transformed_cosine_distance(P[i], p, B + 3 * (size_t)i) : 2.0;
}
// cv-sfm/src/lib.rs:1320-1360 is_tri_landmark_robust; one thread per (centre, first, second) observation triple of one pose pair
__global__ void __launch_bounds__(128) k_tri_landmark_robust(cvb_pose first, cvb_pose second, const double *__restrict__ obs, uint32_t n,
                                                             double max_cos, double inc_min_cos, uint8_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *c = obs + 9 * (size_t)i, *f = c + 3, *s = c + 6;
    cvb_pose P[3];
    for (int k = 0; k < 9; k++) P[0].r[k] = (k % 4 == 0) ? 1.0 : 0.0;
    P[0].t[0] = P[0].t[1] = P[0].t[2] = 0.0;
    P[1] = first; P[2] = second;
    double p[4];
    if (!triangulate_linear_eigen(P, c, 3, p)) { out[i] = 0; return; }
    from_homogeneous(p);   // CameraPoint::from_homogeneous(p.0)
    double fc[3], sc[3];
    for (int k = 0; k < 3; k++) {
        fc[k] = first.r[k] * f[0] + first.r[3 + k] * f[1] + first.r[6 + k] * f[2];
        sc[k] = second.r[k] * s[0] + second.r[3 + k] * s[1] + second.r[6 + k] * s[2];
    }
    const bool cosine_ok = 1.0 - dot3(p, c) < max_cos && transformed_cosine_distance(first, p, f) < max_cos
        && transformed_cosine_distance(second, p, s) < max_cos;
    const bool incidence_ok = 1.0 - dot3(c, fc) > inc_min_cos || 1.0 - dot3(c, sc) > inc_min_cos || 1.0 - dot3(fc, sc) > inc_min_cos;
    out[i] = cosine_ok && incidence_ok;
}

// ------------------------------------------------------------------------------------------ host side
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(cvb_ctx *ctx, size_t need) {
        if (need <= bytes && p) return 0;
        if (p) { cvb_wait(ctx, ctx->stream); cudaFree(p); p = nullptr; bytes = 0; }
        size_t n = std::max<size_t>(need, 256);
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return cvb_set_error(ctx, CVB_ENOMEM, "cudaMalloc(%zu): %s", n, cudaGetErrorString(e));
        bytes = n;
        return 0;
    }
};

}  // namespace

// device-resident ARRSAC (arrsac_dev.cuh): buffers + page-locked staging of one context
struct ArsWorkspace {
    DevBuf ctl, raw, samples0, poses0, nposes0, masks0, vm, pass_id, pass_inl, tposes, tinl, tmasks, newposes, nposes_new, newmask,
        pool, samples_new, res, queue;
    uint32_t *h_raw = nullptr;      // page-locked: raw draws + ArrsacCtl header
    size_t h_raw_cap = 0;
    unsigned char *h_res = nullptr; // page-locked result block
    cudaEvent_t up_done = nullptr;  // the staging buffer may be rewritten once this has fired
    std::vector<cvb_rng> snaps;     // generator state every ARS_SNAP draws of the staged stream
    cvb_rng rng0;                   // generator state at draw 0 of the staged stream
    uint32_t nraw = 0;
    bool pending = false;           // a run whose draw count has not been committed to the caller's generator yet
    // CUDA graphs of the whole run (two copies, ~20 + 3 per data block kernels, one copy back), keyed by everything the enqueue
    // depends on; a key is captured the second time it is seen (a one-off call does not pay the instantiation)
    struct GraphKey {
        ArrsacParams P; int kind, row0; const void *a, *b, *n_dev; uint32_t n_host, nmax, cap, nb; const void *model, *inl, *ninl, *found;
        const void *ws[21];
    };
    struct GraphEntry { GraphKey key; cudaGraphExec_t exec; uint64_t launches; bool loop; };
    bool last_loop = false;         // the pending run went through a WHILE-node graph (its body's launches are counted at commit)
    std::vector<GraphEntry> graphs;
    std::vector<GraphKey> seen;
    int use_graph = -1;             // CVB_NO_GRAPH=1 / CVB_ARS_NO_GRAPH=1 disable
    // the block loop as a WHILE node of the graph (body: score, book, estimate; k_ars_book clears the condition at the loop's
    // end) instead of one unrolled body per possible data block.  CVB_ARS_WHILE=0 keeps the unrolled graph.
    int use_while = -1;
    cudaStream_t body_stream = nullptr;
};
struct GeomWorkspace {
    DevBuf a, b, samples, poses, nposes, out, masks, offsets, ok;
    ArsWorkspace *ars = nullptr;
};
void geom_workspace_free(GeomWorkspace *g) {
    if (!g) return;
    DevBuf *bufs[] = {&g->a, &g->b, &g->samples, &g->poses, &g->nposes, &g->out, &g->masks, &g->offsets, &g->ok};
    for (DevBuf *d : bufs) if (d->p) cudaFree(d->p);
    if (g->ars) {
        ArsWorkspace *w = g->ars;
        DevBuf *ab[] = {&w->ctl, &w->raw, &w->samples0, &w->poses0, &w->nposes0, &w->masks0, &w->vm, &w->pass_id, &w->pass_inl, &w->tposes,
                        &w->tinl, &w->tmasks, &w->newposes, &w->nposes_new, &w->newmask, &w->pool, &w->samples_new, &w->res, &w->queue};
        for (DevBuf *d : ab) if (d->p) cudaFree(d->p);
        if (w->h_raw) cudaFreeHost(w->h_raw);
        if (w->h_res) cudaFreeHost(w->h_res);
        if (w->up_done) cudaEventDestroy(w->up_done);
        if (w->body_stream) cudaStreamDestroy(w->body_stream);
        for (auto &ge : w->graphs) cudaGraphExecDestroy(ge.exec);
        delete w;
    }
    delete g;
}

namespace {

// estimator kinds: 0 EightPoint, 1 LambdaTwist (P3P), 2 NisterStewenius (five-point)
inline uint32_t kind_K(int kind) { return kind == 0 ? 8u : (kind == 1 ? 3u : 5u); }      // Estimator::MIN_SAMPLES
inline uint32_t kind_M(int kind) { return kind == 2 ? 40u : 4u; }                         // ModelIter capacity
inline int kind_res(int kind) { return kind == 1 ? 1 : 0; }                               // residual: 0 CameraToCamera, 1 WorldToCamera

GeomWorkspace *gws(cvb_ctx *ctx) {
    if (!ctx->geom) ctx->geom = new GeomWorkspace();
    return ctx->geom;
}

int upload(cvb_ctx *ctx, DevBuf &d, const void *src, size_t bytes) {
    int rc = d.ensure(ctx, bytes);
    if (rc) return rc;
    if (bytes) CVB_CUDA(ctx, cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// data stay resident in the workspace (a: n x 3; b: n x 3 or n x 4)
int upload_data(cvb_ctx *ctx, int kind, const double *a, const double *b, uint32_t n) {
    GeomWorkspace *g = gws(ctx);
    int rc = upload(ctx, g->a, a, sizeof(double) * 3 * (size_t)n);
    if (rc) return rc;
    return upload(ctx, g->b, b, sizeof(double) * (kind_res(kind) == 0 ? 3 : 4) * (size_t)n);
}

// estimate H minimal samples (host index lists) -> device poses (H x 4) + host counts
int estimate_dev(cvb_ctx *ctx, int kind, const uint32_t *samples, uint32_t H, std::vector<uint8_t> &nposes, int row0 = 5) {
    GeomWorkspace *g = gws(ctx);
    const uint32_t K = kind_K(kind), M = kind_M(kind);
    int rc = upload(ctx, g->samples, samples, sizeof(uint32_t) * K * (size_t)H);
    if (rc) return rc;
    if ((rc = g->poses.ensure(ctx, sizeof(cvb_pose) * M * (size_t)H))) return rc;
    if ((rc = g->nposes.ensure(ctx, H))) return rc;
    CVB_CUDA(ctx, cudaMemsetAsync(g->poses.p, 0, sizeof(cvb_pose) * M * (size_t)H, ctx->stream));
    {
        CVB_PROF(ctx, kind == 0 ? "k_estimate_eight_point" : (kind == 1 ? "k_estimate_p3p" : "k_estimate_five_point"), 0);
        const double *a = (const double *)g->a.p, *b = (const double *)g->b.p;
        const uint32_t *sp = (const uint32_t *)g->samples.p;
        if (kind == 0) k_estimate<0><<<cdiv(H, 128), 128, 0, ctx->stream>>>(a, b, sp, H, (cvb_pose *)g->poses.p, (uint8_t *)g->nposes.p, row0);
        else if (kind == 1) k_estimate<1><<<cdiv(H, 128), 128, 0, ctx->stream>>>(a, b, sp, H, (cvb_pose *)g->poses.p, (uint8_t *)g->nposes.p, row0);
        else k_estimate<2><<<cdiv(H, 128), 128, 0, ctx->stream>>>(a, b, sp, H, (cvb_pose *)g->poses.p, (uint8_t *)g->nposes.p, row0);
        CVB_LAUNCH_CHECK(ctx);
    }
    nposes.resize(H);
    CVB_CUDA(ctx, cudaMemcpyAsync(nposes.data(), g->nposes.p, H, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    return 0;
}

// inlier masks of m device poses over data [i0, i1) -> host (m x words)
int masks_dev(cvb_ctx *ctx, int kind, const cvb_pose *poses_dev, uint32_t m, uint32_t i0, uint32_t i1, double thr,
              std::vector<uint32_t> &masks, uint32_t *words_out) {
    GeomWorkspace *g = gws(ctx);
    const uint32_t cnt = i1 - i0, words = cdiv(cnt, 32);
    *words_out = words;
    masks.assign((size_t)m * words, 0u);
    if (m == 0 || cnt == 0) return 0;
    int rc = g->masks.ensure(ctx, sizeof(uint32_t) * (size_t)m * words);
    if (rc) return rc;
    for (uint32_t p0 = 0; p0 < m; p0 += 65535) {   // gridDim.y limit
        const uint32_t pm = std::min<uint32_t>(65535, m - p0);
        dim3 grid(cdiv(cnt, 256), pm);
        CVB_PROF(ctx, kind_res(kind) == 0 ? "k_residuals_c2c" : "k_residuals_w2c", (kind_res(kind) == 0 ? 48.0 : 56.0) * pm * cnt);
        if (kind_res(kind) == 0)
            k_residuals<0, 1><<<grid, 256, 0, ctx->stream>>>(poses_dev + p0, pm, (const double *)g->a.p, (const double *)g->b.p, i0, i1, thr, nullptr, 0,
                                                             (uint32_t *)g->masks.p + (size_t)p0 * words, words);
        else
            k_residuals<1, 1><<<grid, 256, 0, ctx->stream>>>(poses_dev + p0, pm, (const double *)g->a.p, (const double *)g->b.p, i0, i1, thr, nullptr, 0,
                                                             (uint32_t *)g->masks.p + (size_t)p0 * words, words);
        CVB_LAUNCH_CHECK(ctx);
    }
    CVB_CUDA(ctx, cudaMemcpyAsync(masks.data(), g->masks.p, sizeof(uint32_t) * (size_t)m * words, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    return 0;
}

inline uint32_t popc_range(const uint32_t *row, uint32_t lo, uint32_t hi) {   // bits [lo, hi)
    uint32_t c = 0;
    for (uint32_t i = lo; i < hi; i++) c += (row[i >> 5] >> (i & 31)) & 1u;
    return c;
}

uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

struct Hyp { cvb_pose m; uint32_t inliers; std::vector<uint32_t> mask; };   // mask over ALL n data

void sort_hyps(std::vector<Hyp> &H) {
    std::stable_sort(H.begin(), H.end(), [](const Hyp &x, const Hyp &y) { return x.inliers > y.inliers; });
}

void populate_samples(cvb_rng *rng, uint32_t k, uint32_t len, uint32_t *out) {
    for (uint32_t c = 0; c < k;) {
        uint32_t s = cvb_rng_next_u32(rng) % len;
        bool dup = false;
        for (uint32_t j = 0; j < c; j++) dup |= out[j] == s;
        if (!dup) out[c++] = s;
    }
}

// arrsac::Arrsac::model_inliers, restated (external crate arrsac 0.10.0; see DESIGN.md for what is and is not pinned):
// initialisation with an adaptive SPRT over the first blocks, then block-wise scoring / halving / re-estimation
// from the inliers of the current best.  All residuals come from the GPU as bit masks.
int arrsac_run(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, int kind, const double *a, const double *b, uint32_t n, cvb_rng *rng,
               cvb_pose *model_out, uint32_t *inliers_out, uint32_t cap, uint32_t *n_inliers, int32_t *found, int row0 = 5) {
    const uint32_t K = kind_K(kind), MM = kind_M(kind);
    *found = 0;
    if (n_inliers) *n_inliers = 0;
    if (n < K) return 0;
    if (cfg->block_size == 0 || cfg->initialization_blocks == 0) return cvb_set_error(ctx, CVB_EINVAL, "block_size / initialization_blocks must be > 0");
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload_data(ctx, kind, a, b, n);
    if (rc) return rc;
    GeomWorkspace *g = gws(ctx);
    const double thr = cfg->inlier_threshold;
    const uint32_t nwords = cdiv(n, 32);
    // ---- initialisation: all minimal samples are drawn first (the draws do not depend on any result)
    const uint32_t H0 = cfg->initialization_hypotheses;
    if (H0 == 0) return 0;
    std::vector<uint32_t> samples((size_t)H0 * K);
    for (uint32_t h = 0; h < H0; h++) populate_samples(rng, K, n, samples.data() + (size_t)h * K);
    std::vector<uint8_t> nposes;
    if ((rc = estimate_dev(ctx, kind, samples.data(), H0, nposes, row0))) return rc;
    // inlier masks of every candidate model over the initialisation blocks only (what the SPRT looks at)
    const uint32_t init_n = std::min<uint32_t>(cfg->block_size * cfg->initialization_blocks, n);
    std::vector<uint32_t> masks;
    uint32_t words = 0;
    if ((rc = masks_dev(ctx, kind, (const cvb_pose *)g->poses.p, H0 * MM, 0, init_n, thr, masks, &words))) return rc;
    std::vector<cvb_pose> poses_host((size_t)H0 * MM);
    CVB_CUDA(ctx, cudaMemcpy(poses_host.data(), g->poses.p, sizeof(cvb_pose) * MM * (size_t)H0, cudaMemcpyDeviceToHost));
    float epsilon = cfg->initial_epsilon, delta = cfg->initial_delta;
    uint32_t best_inliers = 0;
    uint64_t rej_inliers = 0, rej_tested = 0;
    std::vector<Hyp> H;
    for (uint32_t h = 0; h < H0; h++)
        for (uint32_t mi = 0; mi < nposes[h]; mi++) {
            const uint32_t *row = masks.data() + ((size_t)h * MM + mi) * words;
            const float pos = delta / epsilon, neg = (1.0f - delta) / (1.0f - epsilon);
            float ratio = 1.0f;
            uint32_t inl = 0, tested = 0;
            bool pass = true;
            for (uint32_t i = 0; i < init_n; i++) {
                tested++;
                if ((row[i >> 5] >> (i & 31)) & 1u) { inl++; ratio *= pos; }
                else ratio *= neg;
                if (ratio > cfg->likelihood_ratio_threshold) { pass = false; break; }
            }
            if (pass) {
                Hyp hy; hy.m = poses_host[(size_t)h * MM + mi]; hy.inliers = inl; hy.mask.assign(nwords, 0u);
                for (uint32_t w = 0; w < words; w++) hy.mask[w] = row[w];
                H.push_back(std::move(hy));
                if (inl > best_inliers) {
                    best_inliers = inl;
                    const float e = (float)inl / (float)init_n;
                    if (e > epsilon && e < 1.0f) epsilon = e; else if (e >= 1.0f) epsilon = 0.999f;
                }
            } else {
                rej_inliers += inl; rej_tested += tested;
                const float d = (float)rej_inliers / (float)rej_tested;
                if (d > 0.0f && d < epsilon) delta = d;
            }
        }
    sort_hyps(H);
    if (H.size() > cfg->max_candidate_hypotheses) H.resize(cfg->max_candidate_hypotheses);
    // the surviving candidates are scored on the remaining data in one launch
    if (init_n < n && !H.empty()) {
        std::vector<cvb_pose> surv(H.size());
        for (size_t i = 0; i < H.size(); i++) surv[i] = H[i].m;
        if ((rc = upload(ctx, g->poses, surv.data(), sizeof(cvb_pose) * surv.size()))) return rc;
        if ((rc = masks_dev(ctx, kind, (const cvb_pose *)g->poses.p, (uint32_t)surv.size(), init_n, n, thr, masks, &words))) return rc;
        for (size_t i = 0; i < H.size(); i++) {
            const uint32_t *row = masks.data() + i * words;
            for (uint32_t j = init_n; j < n; j++)
                if ((row[(j - init_n) >> 5] >> ((j - init_n) & 31)) & 1u) H[i].mask[j >> 5] |= 1u << (j & 31);
        }
    }
    sort_hyps(H);
    if (H.size() > cfg->max_candidate_hypotheses) H.resize(cfg->max_candidate_hypotheses);
    // ---- main loop over further blocks
    std::vector<uint32_t> pool, idx((size_t)std::max<uint32_t>(cfg->estimations_per_block, 1) * K);
    for (uint32_t start = init_n; start < n && H.size() > 1; start += cfg->block_size) {
        const uint32_t end = std::min<uint32_t>(start + cfg->block_size, n);
        for (Hyp &h : H) h.inliers += popc_range(h.mask.data(), start, end);
        sort_hyps(H);
        H.resize(std::max<size_t>(H.size() / 2, 1));
        pool.clear();
        for (uint32_t i = 0; i < end; i++)
            if ((H[0].mask[i >> 5] >> (i & 31)) & 1u) pool.push_back(i);
        if (pool.size() >= K && cfg->estimations_per_block > 0) {
            const uint32_t worst = H.back().inliers;
            const uint32_t G = cfg->estimations_per_block;
            for (uint32_t gi = 0; gi < G; gi++) {
                uint32_t loc[8];
                populate_samples(rng, K, (uint32_t)pool.size(), loc);   // K <= 8
                for (uint32_t k = 0; k < K; k++) idx[(size_t)gi * K + k] = pool[loc[k]];
            }
            if ((rc = estimate_dev(ctx, kind, idx.data(), G, nposes, row0))) return rc;
            if ((rc = masks_dev(ctx, kind, (const cvb_pose *)g->poses.p, G * MM, 0, n, thr, masks, &words))) return rc;
            poses_host.resize((size_t)G * MM);
            CVB_CUDA(ctx, cudaMemcpy(poses_host.data(), g->poses.p, sizeof(cvb_pose) * MM * (size_t)G, cudaMemcpyDeviceToHost));
            for (uint32_t gi = 0; gi < G; gi++)
                for (uint32_t mi = 0; mi < nposes[gi]; mi++) {
                    const uint32_t *row = masks.data() + ((size_t)gi * MM + mi) * words;
                    const uint32_t inl = popc_range(row, 0, end);
                    if (inl > worst) {
                        Hyp hy; hy.m = poses_host[(size_t)gi * MM + mi]; hy.inliers = inl; hy.mask.assign(row, row + words);
                        H.push_back(std::move(hy));
                    }
                }
            sort_hyps(H);
            if (H.size() > cfg->max_candidate_hypotheses) H.resize(cfg->max_candidate_hypotheses);
        }
    }
    if (H.empty()) return 0;
    sort_hyps(H);
    *model_out = H[0].m;
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; i++)
        if ((H[0].mask[i >> 5] >> (i & 31)) & 1u) { if (inliers_out && c < cap) inliers_out[c] = i; c++; }
    if (n_inliers) *n_inliers = c;
    *found = 1;
    (void)nwords;
    if (inliers_out && c > cap) return cvb_set_error(ctx, CVB_ECAP, "inlier capacity %u too small (%u needed)", cap, c);
    return 0;
}


// ---- device-resident ARRSAC driver (kernels: arrsac_dev.cuh).  Enqueues everything on the context stream and returns;
// no host synchronisation between the first kernel and the result.
#define ARS_SNAP 4096u
ArsWorkspace *arsws(cvb_ctx *ctx) {
    GeomWorkspace *g = gws(ctx);
    if (!g->ars) g->ars = new ArsWorkspace();
    return g->ars;
}

int arrsac_run_dev(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, int kind, const double *a_dev, const double *b_dev, const uint32_t *n_dev,
                   uint32_t n_host, uint32_t nmax, const cvb_rng *rng, cvb_pose *model_dev, uint32_t *inl_dev, uint32_t cap,
                   uint32_t *ninl_dev, int32_t *found_dev, int row0) {
    if (cfg->block_size == 0 || cfg->initialization_blocks == 0) return cvb_set_error(ctx, CVB_EINVAL, "block_size / initialization_blocks must be > 0");
    ArrsacParams P;
    memset(&P, 0, sizeof(P));
    P.K = kind_K(kind); P.MM = kind_M(kind); P.kind = (uint32_t)kind;
    P.H0 = cfg->initialization_hypotheses; P.ib = cfg->initialization_blocks; P.bs = cfg->block_size;
    P.max_cand = cfg->max_candidate_hypotheses; P.G = cfg->estimations_per_block;
    P.NMAX = std::max<uint32_t>(nmax, 1);
    P.W0 = cdiv(P.bs * P.ib, 32); P.NW = cdiv(P.NMAX, 32);
    P.rows = P.max_cand + P.G * P.MM;
    P.lr_thr = cfg->likelihood_ratio_threshold; P.eps0 = cfg->initial_epsilon; P.delta0 = cfg->initial_delta;
    P.thr = cfg->inlier_threshold; P.row0 = row0;
    // two-stage initial scoring only where a predicate is expensive (CameraToCamera residual); CVB_ARS_EAGER=1 scores everything up front
    { const char *env = getenv("CVB_ARS_EAGER"); const bool eager = (env && env[0] == '1') || kind_res(kind) == 1;
      P.prefix = eager ? P.H0 : std::min<uint32_t>(P.H0, 64); P.cmin = 2;
      if (const char *cm = getenv("CVB_ARS_CMIN")) P.cmin = (uint32_t)std::max(0, atoi(cm)); }
    if (P.max_cand == 0 || P.rows > ARS_SORT_CAP)
        return cvb_set_error(ctx, CVB_EUNSUPPORTED, "max_candidate_hypotheses + estimations_per_block * %u must be in 1..%u", P.MM, ARS_SORT_CAP);
    if ((uint64_t)P.bs * P.ib + 1 > 2ull * ARS_SORT_CAP) return cvb_set_error(ctx, CVB_EUNSUPPORTED, "block_size * initialization_blocks too large");
    if (P.NMAX >= (1u << 20)) return cvb_set_error(ctx, CVB_EUNSUPPORTED, "more than 2^20 data");
    ArsWorkspace *w = arsws(ctx);
    const uint32_t nb_max = cdiv(P.NMAX, P.bs) + 1;
    const uint32_t nraw = P.H0 * P.K + P.H0 * P.K / 4 + 64 + nb_max * (P.G * P.K + P.G * P.K / 4 + 64);
    const size_t nmodels0 = (size_t)P.H0 * P.MM, nnew = (size_t)std::max<uint32_t>(P.G, 1) * P.MM;
    int rc;
    if ((rc = w->ctl.ensure(ctx, sizeof(ArrsacCtl)))) return rc;
    if ((rc = w->raw.ensure(ctx, sizeof(uint32_t) * (size_t)nraw))) return rc;
    if ((rc = w->samples0.ensure(ctx, sizeof(uint32_t) * (size_t)std::max<uint32_t>(P.H0, 1) * P.K))) return rc;
    if ((rc = w->poses0.ensure(ctx, sizeof(cvb_pose) * std::max<size_t>(nmodels0, 1)))) return rc;
    if ((rc = w->nposes0.ensure(ctx, std::max<uint32_t>(P.H0, 1)))) return rc;
    if ((rc = w->masks0.ensure(ctx, sizeof(uint32_t) * std::max<size_t>(nmodels0, 1) * P.W0))) return rc;
    if ((rc = w->vm.ensure(ctx, sizeof(uint32_t) * std::max<size_t>(nmodels0, ARS_SORT_CAP)))) return rc;
    if ((rc = w->pass_id.ensure(ctx, sizeof(uint32_t) * std::max<size_t>(nmodels0, 1)))) return rc;
    if ((rc = w->pass_inl.ensure(ctx, sizeof(uint32_t) * std::max<size_t>(nmodels0, 1)))) return rc;
    if ((rc = w->tposes.ensure(ctx, sizeof(cvb_pose) * 2 * (size_t)P.rows))) return rc;
    if ((rc = w->tinl.ensure(ctx, sizeof(uint32_t) * 2 * (size_t)P.rows))) return rc;
    if ((rc = w->tmasks.ensure(ctx, sizeof(uint32_t) * 2 * (size_t)P.rows * P.NW))) return rc;
    if ((rc = w->newposes.ensure(ctx, sizeof(cvb_pose) * nnew))) return rc;
    if ((rc = w->nposes_new.ensure(ctx, std::max<uint32_t>(P.G, 1)))) return rc;
    if ((rc = w->newmask.ensure(ctx, sizeof(uint32_t) * nnew * P.NW))) return rc;
    if ((rc = w->pool.ensure(ctx, sizeof(uint32_t) * (size_t)P.NMAX))) return rc;
    if ((rc = w->samples_new.ensure(ctx, sizeof(uint32_t) * (size_t)std::max<uint32_t>(P.G, 1) * P.K))) return rc;
    if ((rc = w->queue.ensure(ctx, sizeof(uint2) * 2 * (size_t)ARS_QCAP))) return rc;
    // page-locked staging: [ArrsacCtl header | raw draws]
    const size_t hdr = (sizeof(ArrsacCtl) + 15) / 16 * 16, stage_bytes = hdr + sizeof(uint32_t) * (size_t)nraw;
    if (w->h_raw_cap < stage_bytes) {
        if (w->h_raw) { cvb_wait(ctx, ctx->stream); cudaFreeHost(w->h_raw); w->h_raw = nullptr; w->h_raw_cap = 0; }
        if (cudaHostAlloc((void **)&w->h_raw, stage_bytes, cudaHostAllocDefault) != cudaSuccess) return cvb_set_error(ctx, CVB_ENOMEM, "page-locked staging");
        w->h_raw_cap = stage_bytes;
    }
    if (!w->h_res && cudaHostAlloc((void **)&w->h_res, sizeof(ArrsacCtl) + sizeof(cvb_pose) + 64, cudaHostAllocDefault) != cudaSuccess)
        return cvb_set_error(ctx, CVB_ENOMEM, "page-locked staging");
    if (!w->up_done) CVB_CUDA(ctx, cudaEventCreateWithFlags(&w->up_done, cudaEventDisableTiming | cudaEventBlockingSync));
    else CVB_CUDA(ctx, cudaEventSynchronize(w->up_done));       // previous upload has left the staging buffer
    {
        cvb_rng g = *rng;
        w->rng0 = g;
        w->snaps.clear();
        uint32_t *raw = (uint32_t *)((unsigned char *)w->h_raw + hdr);
        for (uint32_t i = 0; i < nraw; i++) {
            if (i % ARS_SNAP == 0) w->snaps.push_back(g);
            raw[i] = cvb_rng_next_u32(&g);
        }
        ArrsacCtl *h = (ArrsacCtl *)w->h_raw;
        memset(h, 0, sizeof(*h));
        h->nraw = nraw; h->gen = g; h->gen_pos = nraw; h->rng_pos = 0;
        w->nraw = nraw;
    }
    cudaStream_t st = ctx->stream;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_ars_book, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ARS_BOOK_SMEM);
        cudaFuncSetAttribute(k_ars_sprt<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k_ars_sprt<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k_ars_score<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_ars_score<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    bool capturing = false, while_loop = false;
    // everything the stream sees, from the upload of the draw stream to the copy of the control block back
    auto enqueue = [&]() -> int {
        CVB_CUDA(ctx, cudaMemcpyAsync(w->ctl.p, w->h_raw, sizeof(ArrsacCtl), cudaMemcpyHostToDevice, st));
        CVB_CUDA(ctx, cudaMemcpyAsync(w->raw.p, (unsigned char *)w->h_raw + hdr, sizeof(uint32_t) * (size_t)nraw, cudaMemcpyHostToDevice, st));
        if (!capturing) CVB_CUDA(ctx, cudaEventRecord(w->up_done, st));      // eager: the staging buffer is free as soon as the two copies are done
        ArrsacCtl *ctl = (ArrsacCtl *)w->ctl.p;
        const uint32_t *raw = (const uint32_t *)w->raw.p;
        const int res = kind_res(kind);
        {
            CVB_PROF(ctx, "k_ars_begin", 0);
            k_ars_begin<<<1, ARS_BOOK_NT, 0, st>>>(ctl, P, n_dev, n_host, raw, (uint32_t *)w->samples0.p);
            CVB_LAUNCH_CHECK(ctx);
        }
        auto estimate = [&](int phase, uint32_t H, const uint32_t *samples, cvb_pose *poses, uint8_t *nposes) -> int {
            if (H == 0) return 0;
            CVB_PROF(ctx, phase == 0 ? "k_ars_estimate_init" : "k_ars_estimate_block", 0);
            // the big initial batch is bound by FP64 issue (4 lanes per hypothesis waste the fewest slots); a block's 64 hypotheses
            // are a latency chain in front of the next scoring (16 lanes: the shortest chain)
            // (8 lanes, or 16 lanes at 64 registers, measured the same 0.137-0.140 ms for the initial batch)
            if (kind == 0 && phase == 0) k_ars_estimate8<4, 5><<<cdiv(H, 128 / 4), 128, 0, st>>>(ctl, phase, H, a_dev, b_dev, samples, poses, nposes);
            else if (kind == 0) k_ars_estimate8<16, 5><<<cdiv(H, 128 / 16), 128, 0, st>>>(ctl, phase, H, a_dev, b_dev, samples, poses, nposes);
            else if (kind == 1) k_ars_estimate<1><<<cdiv(H, 128), 128, 0, st>>>(ctl, phase, H, a_dev, b_dev, samples, poses, nposes, row0);
            else k_ars_estimate<2><<<cdiv(H, 128), 128, 0, st>>>(ctl, phase, H, a_dev, b_dev, samples, poses, nposes, row0);
            CVB_LAUNCH_CHECK(ctx);
            return 0;
        };
        // the initial scoring fills the machine (4 CTAs per SM).  A block scores ~200 k predicates; measured on B200 per pair (76 launches,
    // most of them idle because the loop is over): 1.23 ms on 48 CTAs, 0.88 ms on 148, 0.79 ms on 296 for ONE context -- but 16
    // pipelined contexts reach 1 678 / 1 660 / 1 629 frames/s: the small grid costs the least SM time (more units per warp, the
    // exact-fallback stragglers amortised), and the step is bound by SM time, not by a pair's latency.  CVB_ARS_SGRID overrides.
    const uint32_t sgrid_full = (uint32_t)ctx->num_sms * 4;
        uint32_t sgrid_block = 48;
        if (const char *e = getenv("CVB_ARS_SGRID")) sgrid_block = (uint32_t)std::max(1, atoi(e));
        // CVB_ARS_SCORE_SMEM: unused dynamic shared memory per scoring CTA, a residency limiter.  The kernel needs 128 registers per
        // thread, so two CTAs take an SM's whole register file and nothing of another context can run beside them although they
        // only use the FP64 pipe; > half of the SM's shared memory leaves one CTA per SM and half the registers to other kernels.
        size_t score_smem = 0;
        if (const char *e = getenv("CVB_ARS_SCORE_SMEM")) score_smem = (size_t)std::max(0, atoi(e));
        auto score = [&](int phase) -> int {
            const uint32_t sgrid = phase == 1 ? sgrid_block : sgrid_full;
            CVB_PROF(ctx, phase == 1 ? "k_ars_score_block" : "k_ars_score_init", 0);
            if (res == 0)
                k_ars_score<0><<<sgrid, 256, score_smem, st>>>(ctl, (uint2 *)w->queue.p, P, phase, a_dev, b_dev, (const cvb_pose *)w->poses0.p, (const uint8_t *)w->nposes0.p,
                                                      (uint32_t *)w->masks0.p, (const cvb_pose *)w->tposes.p, (uint32_t *)w->tmasks.p,
                                                      (const cvb_pose *)w->newposes.p, (const uint8_t *)w->nposes_new.p, (uint32_t *)w->newmask.p);
            else
                k_ars_score<1><<<sgrid, 256, score_smem, st>>>(ctl, (uint2 *)w->queue.p, P, phase, a_dev, b_dev, (const cvb_pose *)w->poses0.p, (const uint8_t *)w->nposes0.p,
                                                      (uint32_t *)w->masks0.p, (const cvb_pose *)w->tposes.p, (uint32_t *)w->tmasks.p,
                                                      (const cvb_pose *)w->newposes.p, (const uint8_t *)w->nposes_new.p, (uint32_t *)w->newmask.p);
            CVB_LAUNCH_CHECK(ctx);
            return 0;
        };
        if ((rc = estimate(0, P.H0, (const uint32_t *)w->samples0.p, (cvb_pose *)w->poses0.p, (uint8_t *)w->nposes0.p))) return rc;
        auto resolve = [&](int stage) -> int {
            CVB_PROF(ctx, "k_ars_resolve", 0);
            k_ars_resolve<<<sgrid_full, 256, 0, st>>>(ctl, (const uint2 *)w->queue.p, stage, P, a_dev, b_dev, (const cvb_pose *)w->poses0.p, (uint32_t *)w->masks0.p);
            CVB_LAUNCH_CHECK(ctx);
            return 0;
        };
        if ((rc = score(0))) return rc;
        if (res == 0 && (rc = resolve(0))) return rc;
        if (P.prefix < P.H0 && P.W0 > 1) {
            if ((rc = score(2))) return rc;
            if (res == 0 && (rc = resolve(1))) return rc;
        }
        {
            CVB_PROF(ctx, "k_ars_sprt", 0);
            if (res == 0)
                k_ars_sprt<0><<<1, ARS_BOOK_NT, 8 * ARS_SORT_CAP, st>>>(ctl, P, a_dev, b_dev, (const cvb_pose *)w->poses0.p, (const uint8_t *)w->nposes0.p, (uint32_t *)w->masks0.p,
                                                        (uint32_t *)w->vm.p, (uint32_t *)w->pass_id.p, (uint32_t *)w->pass_inl.p, (cvb_pose *)w->tposes.p,
                                                        (uint32_t *)w->tinl.p, (uint32_t *)w->tmasks.p);
            else
                k_ars_sprt<1><<<1, ARS_BOOK_NT, 8 * ARS_SORT_CAP, st>>>(ctl, P, a_dev, b_dev, (const cvb_pose *)w->poses0.p, (const uint8_t *)w->nposes0.p, (uint32_t *)w->masks0.p,
                                                        (uint32_t *)w->vm.p, (uint32_t *)w->pass_id.p, (uint32_t *)w->pass_inl.p, (cvb_pose *)w->tposes.p,
                                                        (uint32_t *)w->tinl.p, (uint32_t *)w->tmasks.p);
            CVB_LAUNCH_CHECK(ctx);
        }
        // block loop: the number of launches follows the data count when the host knows it, the capacity otherwise;
        // kernels behind the loop's end return at once (ctl->done)
        const uint32_t n_bound = n_dev ? P.NMAX : std::min(n_host, P.NMAX);
        const uint32_t init_n = std::min(P.bs * P.ib, n_bound);
        const uint32_t nb = n_bound > init_n ? cdiv(n_bound - init_n, P.bs) : 0;
        auto book = [&](unsigned long long cond) -> int {
            CVB_PROF(ctx, "k_ars_book", 0);
            k_ars_book<<<1, ARS_BOOK_NT, ARS_BOOK_SMEM, st>>>(ctl, P, raw, (cvb_pose *)w->tposes.p, (uint32_t *)w->tinl.p, (uint32_t *)w->tmasks.p,
                                                              (const cvb_pose *)w->newposes.p, (const uint8_t *)w->nposes_new.p,
                                                              (const uint32_t *)w->newmask.p, (uint32_t *)w->pool.p, (uint32_t *)w->samples_new.p, cond);
            CVB_LAUNCH_CHECK(ctx);
            return 0;
        };
        if (capturing && while_loop) {
            // device-side loop: one WHILE node whose body is one block iteration; k_ars_book ends it (block nb at the latest: lo >= n)
            cudaStreamCaptureStatus cs;
            cudaGraph_t g = nullptr;
            const cudaGraphNode_t *deps = nullptr;
            size_t ndeps = 0;
            CVB_CUDA(ctx, cudaStreamGetCaptureInfo(st, &cs, nullptr, &g, &deps, &ndeps));
            cudaGraphConditionalHandle cond;
            CVB_CUDA(ctx, cudaGraphConditionalHandleCreate(&cond, g, 1, cudaGraphCondAssignDefault));
            cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
            np.conditional.handle = cond;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            CVB_CUDA(ctx, cudaGraphAddNode(&node, g, deps, ndeps, &np));
            cudaGraph_t body = np.conditional.phGraph_out[0];
            const cudaStream_t outer = st;
            CVB_CUDA(ctx, cudaStreamBeginCaptureToGraph(w->body_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            st = w->body_stream;                     // the launch helpers below capture into the body (restored before any return)
            rc = score(1);
            if (!rc) rc = book(cond);
            if (!rc && nb && P.G) rc = estimate(1, P.G, (const uint32_t *)w->samples_new.p, (cvb_pose *)w->newposes.p, (uint8_t *)w->nposes_new.p);
            cudaGraph_t same_body = nullptr;
            const cudaError_t be = cudaStreamEndCapture(st, &same_body);
            st = outer;
            if (rc) return rc;
            CVB_CUDA(ctx, be);
            CVB_CUDA(ctx, cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
        } else
        for (uint32_t it = 0; it <= nb; it++) {
            if ((rc = score(1))) return rc;
            if ((rc = book(0))) return rc;
            if (it < nb && P.G)
                if ((rc = estimate(1, P.G, (const uint32_t *)w->samples_new.p, (cvb_pose *)w->newposes.p, (uint8_t *)w->nposes_new.p))) return rc;
        }
        {
            CVB_PROF(ctx, "k_ars_final", 0);
            if (res == 0) k_ars_final<0><<<1, ARS_BOOK_NT, 0, st>>>(ctl, P, a_dev, b_dev, model_dev, inl_dev, cap, ninl_dev, found_dev);
            else k_ars_final<1><<<1, ARS_BOOK_NT, 0, st>>>(ctl, P, a_dev, b_dev, model_dev, inl_dev, cap, ninl_dev, found_dev);
            CVB_LAUNCH_CHECK(ctx);
        }
        CVB_CUDA(ctx, cudaMemcpyAsync(w->h_res, w->ctl.p, sizeof(ArrsacCtl), cudaMemcpyDeviceToHost, st));
        return 0;
    };
    if (w->use_graph < 0) {
        const char *e1 = getenv("CVB_NO_GRAPH"), *e2 = getenv("CVB_ARS_NO_GRAPH");
        w->use_graph = ((e1 && e1[0] == '1') || (e2 && e2[0] == '1')) ? 0 : 1;
    }
    if (!w->use_graph || ctx->prof) {
        if ((rc = enqueue())) return rc;
        w->pending = true;
        return 0;
    }
    ArsWorkspace::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.P = P; key.kind = kind; key.row0 = row0; key.a = a_dev; key.b = b_dev; key.n_dev = n_dev; key.n_host = n_host; key.nmax = nmax; key.cap = cap;
    key.model = model_dev; key.inl = inl_dev; key.ninl = ninl_dev; key.found = found_dev;
    {
        const DevBuf *bufs[] = {&w->ctl, &w->raw, &w->samples0, &w->poses0, &w->nposes0, &w->masks0, &w->vm, &w->pass_id, &w->pass_inl, &w->tposes,
                                &w->tinl, &w->tmasks, &w->newposes, &w->nposes_new, &w->newmask, &w->pool, &w->samples_new, &w->queue};
        int i = 0;
        for (const DevBuf *d : bufs) key.ws[i++] = d->p;
        key.ws[i++] = w->h_raw; key.ws[i++] = w->h_res; key.ws[i++] = (const void *)(uintptr_t)nraw;
    }
    auto same = [](const ArsWorkspace::GraphKey &x, const ArsWorkspace::GraphKey &y) { return memcmp(&x, &y, sizeof(x)) == 0; };
    for (auto &ge : w->graphs)
        if (same(ge.key, key)) {
            CVB_CUDA(ctx, cudaGraphLaunch(ge.exec, st));
            CVB_CUDA(ctx, cudaEventRecord(w->up_done, st));       // replay: the staging buffer is free when the run is over
            ctx->launches += ge.launches;
            w->last_loop = ge.loop;
            w->pending = true;
            return 0;
        }
    bool seen = false;
    for (auto &k : w->seen) seen = seen || same(k, key);
    if (!seen) {                                                   // first time: run eagerly, capture if it comes back
        if (w->seen.size() >= 64) w->seen.erase(w->seen.begin());
        w->seen.push_back(key);
        if ((rc = enqueue())) return rc;
        w->pending = true;
        return 0;
    }
    const uint64_t l0 = ctx->launches;
    if (w->use_while < 0) {
        const char *e = getenv("CVB_ARS_WHILE");
        w->use_while = (e && e[0] == '0') ? 0 : 1;
    }
    if (w->use_while && !w->body_stream && cudaStreamCreateWithFlags(&w->body_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        w->use_while = 0;
    }
    cudaGraphExec_t exec = nullptr;
    cudaError_t ce = cudaSuccess;
    for (int attempt = 0; attempt < 2 && !exec; attempt++) {
        ctx->launches = l0;
        capturing = true;
        while_loop = w->use_while != 0;
        CVB_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        rc = enqueue();
        cudaGraph_t graph = nullptr;
        ce = cudaStreamEndCapture(st, &graph);
        capturing = false;
        if (!rc && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (exec || !while_loop) break;
        cudaGetLastError();                                        // the WHILE node is not available: capture the unrolled loop instead
        w->use_while = 0;
        exec = nullptr;
    }
    if (rc) return rc;
    if (ce != cudaSuccess || !exec) {                              // capture not possible: eager from now on
        cudaGetLastError();
        w->use_graph = 0;
        ctx->launches = l0;
        if ((rc = enqueue())) return rc;
        w->pending = true;
        return 0;
    }
    if (w->graphs.size() >= 16) { cudaGraphExecDestroy(w->graphs.front().exec); w->graphs.erase(w->graphs.begin()); }
    w->graphs.push_back({key, exec, ctx->launches - l0, while_loop});
    w->last_loop = while_loop;
    CVB_CUDA(ctx, cudaGraphLaunch(exec, st));
    CVB_CUDA(ctx, cudaEventRecord(w->up_done, st));
    w->pending = true;
    return 0;
}

// after the stream has drained: advance the caller's generator by the draws the run consumed (state in == state the
// reference would hold after model_inliers)
int arrsac_commit_rng(cvb_ctx *ctx, cvb_rng *rng, ArrsacCtl *stats_out = nullptr) {
    ArsWorkspace *w = arsws(ctx);
    if (!w->pending) return cvb_set_error(ctx, CVB_EINVAL, "no device ARRSAC run to commit");
    const ArrsacCtl *h = (const ArrsacCtl *)w->h_res;
    if (stats_out) *stats_out = *h;
    if (w->last_loop) ctx->launches += 3ull * h->iters;           // the WHILE body ran iters + 1 times, the capture counted it once
    w->last_loop = false;
    w->pending = false;
    if (!rng) return 0;
    const uint64_t used = h->rng_pos;
    if (used >= w->nraw) {
        if (h->gen_pos != std::max<uint64_t>(used, w->nraw)) return cvb_set_error(ctx, CVB_ECUDA, "generator position mismatch");
        *rng = h->gen;
        return 0;
    }
    cvb_rng g = w->snaps[used / ARS_SNAP];
    for (uint64_t i = used / ARS_SNAP * ARS_SNAP; i < used; i++) cvb_rng_next_u32(&g);
    *rng = g;
    return 0;
}

// host-pointer entry: upload, run on the device, one synchronisation at the end
int arrsac_host(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, int kind, const double *a, const double *b, uint32_t n, cvb_rng *rng,
                cvb_pose *model_out, uint32_t *inliers_out, uint32_t cap, uint32_t *n_inliers, int32_t *found, int row0 = 5) {
    *found = 0;
    if (n_inliers) *n_inliers = 0;
    if (n < kind_K(kind) || cfg->initialization_hypotheses == 0) return 0;
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload_data(ctx, kind, a, b, n);
    if (rc) return rc;
    GeomWorkspace *g = gws(ctx);
    ArsWorkspace *w = arsws(ctx);
    const size_t res_bytes = sizeof(cvb_pose) + 16 + sizeof(uint32_t) * (size_t)n;
    if ((rc = w->res.ensure(ctx, res_bytes))) return rc;
    unsigned char *rd = (unsigned char *)w->res.p;
    cvb_pose *model_dev = (cvb_pose *)rd;
    uint32_t *ninl_dev = (uint32_t *)(rd + sizeof(cvb_pose));
    int32_t *found_dev = (int32_t *)(rd + sizeof(cvb_pose) + 4);
    uint32_t *inl_dev = (uint32_t *)(rd + sizeof(cvb_pose) + 16);
    if ((rc = arrsac_run_dev(ctx, cfg, kind, (const double *)g->a.p, (const double *)g->b.p, nullptr, n, n, rng, model_dev, inl_dev, n, ninl_dev,
                             found_dev, row0))) return rc;
    unsigned char *hs = (unsigned char *)cvb_pinned(ctx, res_bytes);
    if (!hs) return cvb_set_error(ctx, CVB_ENOMEM, "page-locked scratch");
    CVB_CUDA(ctx, cudaMemcpyAsync(hs, rd, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    if ((rc = arrsac_commit_rng(ctx, rng))) return rc;
    const uint32_t c = *(const uint32_t *)(hs + sizeof(cvb_pose));
    *found = *(const int32_t *)(hs + sizeof(cvb_pose) + 4);
    if (!*found) return 0;
    memcpy(model_out, hs, sizeof(cvb_pose));
    if (n_inliers) *n_inliers = c;
    if (inliers_out) memcpy(inliers_out, hs + sizeof(cvb_pose) + 16, sizeof(uint32_t) * std::min(c, cap));
    if (inliers_out && c > cap) return cvb_set_error(ctx, CVB_ECAP, "inlier capacity %u too small (%u needed)", cap, c);
    return 0;
}

bool arrsac_on_host() {     // CVB_ARRSAC_HOST=1: round-1 driver (bookkeeping on the host) kept for A/B tests
    const char *e = getenv("CVB_ARRSAC_HOST");
    return e && e[0] == '1';
}

int residuals_host(cvb_ctx *ctx, int kind, const cvb_pose *poses, uint32_t m, const double *a, const double *b, uint32_t n, double *out) {
    if (!ctx) return CVB_EINVAL;
    if ((m && !poses) || (n && (!a || !b)) || (m && n && !out)) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    if (m == 0 || n == 0) return 0;
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload_data(ctx, kind, a, b, n);
    if (rc) return rc;
    GeomWorkspace *g = gws(ctx);
    if ((rc = upload(ctx, g->poses, poses, sizeof(cvb_pose) * (size_t)m))) return rc;
    if ((rc = g->out.ensure(ctx, sizeof(double) * (size_t)m * n))) return rc;
    for (uint32_t p0 = 0; p0 < m; p0 += 65535) {
        const uint32_t pm = std::min<uint32_t>(65535, m - p0);
        dim3 grid(cdiv(n, 256), pm);
        CVB_PROF(ctx, kind == 0 ? "k_residuals_c2c" : "k_residuals_w2c", (kind == 0 ? 48.0 : 56.0) * pm * n);
        if (kind == 0)
            k_residuals<0, 0><<<grid, 256, 0, ctx->stream>>>((const cvb_pose *)g->poses.p + p0, pm, (const double *)g->a.p, (const double *)g->b.p, 0, n, 0.0,
                                                             (double *)g->out.p + (size_t)p0 * n, n, nullptr, 0);
        else
            k_residuals<1, 0><<<grid, 256, 0, ctx->stream>>>((const cvb_pose *)g->poses.p + p0, pm, (const double *)g->a.p, (const double *)g->b.p, 0, n, 0.0,
                                                             (double *)g->out.p + (size_t)p0 * n, n, nullptr, 0);
        CVB_LAUNCH_CHECK(ctx);
    }
    CVB_CUDA(ctx, cudaMemcpyAsync(out, g->out.p, sizeof(double) * (size_t)m * n, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    return 0;
}

int estimate_host(cvb_ctx *ctx, int kind, const double *a, const double *b, uint32_t n, const uint32_t *samples, uint32_t H,
                  cvb_pose *poses_out, uint8_t *nposes_out, int row0 = 5) {
    if (!ctx) return CVB_EINVAL;
    if (!a || !b || (H && (!samples || !poses_out || !nposes_out))) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    if (H == 0) return 0;
    const uint32_t K = kind_K(kind);
    for (size_t i = 0; i < (size_t)H * K; i++)
        if (samples[i] >= n) return cvb_set_error(ctx, CVB_EINVAL, "sample index %u out of range (n = %u)", samples[i], n);
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = upload_data(ctx, kind, a, b, n);
    if (rc) return rc;
    std::vector<uint8_t> np;
    if ((rc = estimate_dev(ctx, kind, samples, H, np, row0))) return rc;
    memcpy(nposes_out, np.data(), H);
    CVB_CUDA(ctx, cudaMemcpy(poses_out, gws(ctx)->poses.p, sizeof(cvb_pose) * kind_M(kind) * (size_t)H, cudaMemcpyDeviceToHost));
    return 0;
}

// poses (g->out) and per-problem update counts (g->ok) through page-locked scratch to the caller's (possibly pageable) arrays
int download_poses_updates(cvb_ctx *ctx, GeomWorkspace *g, cvb_pose *poses_out, size_t nposes, uint32_t *updates_out, uint32_t B) {
    const size_t pb = sizeof(cvb_pose) * nposes, ub = sizeof(uint32_t) * (size_t)B;
    unsigned char *hs = (unsigned char *)cvb_pinned(ctx, pb + ub);
    if (!hs) return cvb_set_error(ctx, CVB_ENOMEM, "page-locked scratch");
    CVB_CUDA(ctx, cudaMemcpyAsync(hs, g->out.p, pb, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cudaMemcpyAsync(hs + pb, g->ok.p, ub, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    memcpy(poses_out, hs, pb);
    if (updates_out) memcpy(updates_out, hs + pb, ub);
    return 0;
}

}  // namespace

extern "C" {

void cvb_arrsac_default_cfg(cvb_arrsac_cfg *c, double inlier_threshold) {
    if (!c) return;
    c->inlier_threshold = inlier_threshold;
    c->initialization_hypotheses = 256; c->initialization_blocks = 4; c->max_candidate_hypotheses = 64;
    c->estimations_per_block = 64; c->block_size = 64;
    c->likelihood_ratio_threshold = 1e3f; c->initial_epsilon = 0.1f; c->initial_delta = 0.05f;
}

void cvb_rng_seed_xoshiro256pp(cvb_rng *r, uint64_t seed) {   // SplitMix64 expansion
    if (!r) return;
    r->kind = 0;
    for (int i = 0; i < 4; i++) {
        seed += 0x9e3779b97f4a7c15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        r->s[i] = z ^ (z >> 31);
    }
}

void cvb_rng_seed_pcg64(cvb_rng *r, const uint8_t seed[32]) {   // Lcg128Xsl64::from_seed
    if (!r || !seed) return;
    r->kind = 1;
    uint64_t w[4];
    memcpy(w, seed, 32);
    unsigned __int128 state = (unsigned __int128)w[0] | ((unsigned __int128)w[1] << 64);
    unsigned __int128 incr = ((unsigned __int128)w[2] | ((unsigned __int128)w[3] << 64)) | 1;
    const unsigned __int128 MUL = ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
    state = state + incr;
    state = state * MUL + incr;
    r->s[0] = (uint64_t)state; r->s[1] = (uint64_t)(state >> 64); r->s[2] = (uint64_t)incr; r->s[3] = (uint64_t)(incr >> 64);
}

uint32_t cvb_rng_next_u32(cvb_rng *r) {
    if (!r) return 0;
    if (r->kind == 0) {
        uint64_t *s = r->s;
        const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
        return (uint32_t)(result >> 32);
    }
    unsigned __int128 state = (unsigned __int128)r->s[0] | ((unsigned __int128)r->s[1] << 64);
    const unsigned __int128 incr = (unsigned __int128)r->s[2] | ((unsigned __int128)r->s[3] << 64);
    const unsigned __int128 MUL = ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
    state = state * MUL + incr;
    r->s[0] = (uint64_t)state; r->s[1] = (uint64_t)(state >> 64);
    const uint32_t rot = (uint32_t)(state >> 122);
    const uint64_t xsl = (uint64_t)(state >> 64) ^ (uint64_t)state;
    return (uint32_t)((xsl >> rot) | (xsl << ((64 - rot) & 63)));
}

int cvb_eight_point_batch(cvb_ctx *ctx, const double *a, const double *b, uint32_t n, const uint32_t *samples, uint32_t H,
                          cvb_pose *poses_out, uint8_t *nposes_out) {
    return estimate_host(ctx, 0, a, b, n, samples, H, poses_out, nposes_out);
}
int cvb_p3p_batch(cvb_ctx *ctx, const double *bearings, const double *world, uint32_t n, const uint32_t *samples, uint32_t H,
                  cvb_pose *poses_out, uint8_t *nposes_out) {
    return estimate_host(ctx, 1, bearings, world, n, samples, H, poses_out, nposes_out);
}
int cvb_five_point_batch(cvb_ctx *ctx, const double *a, const double *b, uint32_t n, const uint32_t *samples, uint32_t H,
                         int32_t eigenvector_row0, cvb_pose *poses_out, uint8_t *nposes_out) {
    if (ctx && eigenvector_row0 != 5 && eigenvector_row0 != 6) return cvb_set_error(ctx, CVB_EINVAL, "eigenvector_row0 must be 5 (reference) or 6 (corrected)");
    return estimate_host(ctx, 2, a, b, n, samples, H, poses_out, nposes_out, eigenvector_row0);
}
int cvb_residuals_camera_to_camera(cvb_ctx *ctx, const cvb_pose *poses, uint32_t m, const double *a, const double *b, uint32_t n, double *out) {
    return residuals_host(ctx, 0, poses, m, a, b, n, out);
}
int cvb_residuals_world_to_camera(cvb_ctx *ctx, const cvb_pose *poses, uint32_t m, const double *bearings, const double *world, uint32_t n,
                                  double *out) {
    return residuals_host(ctx, 1, poses, m, bearings, world, n, out);
}

int cvb_triangulate_linear_eigen(cvb_ctx *ctx, const cvb_pose *poses, const double *bearings, const uint32_t *offsets, uint32_t L,
                                 double *xyzw_out, uint8_t *ok_out) {
    if (!ctx) return CVB_EINVAL;
    if (L == 0) return 0;
    if (!poses || !bearings || !offsets || !xyzw_out || !ok_out) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    for (uint32_t l = 0; l < L; l++)
        if (offsets[l + 1] < offsets[l]) return cvb_set_error(ctx, CVB_EINVAL, "offsets must be non-decreasing");
    const uint32_t nobs = offsets[L];
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    GeomWorkspace *g = gws(ctx);
    int rc;
    if ((rc = upload(ctx, g->poses, poses, sizeof(cvb_pose) * (size_t)nobs))) return rc;
    if ((rc = upload(ctx, g->a, bearings, sizeof(double) * 3 * (size_t)nobs))) return rc;
    if ((rc = upload(ctx, g->offsets, offsets, sizeof(uint32_t) * ((size_t)L + 1)))) return rc;
    if ((rc = g->out.ensure(ctx, sizeof(double) * 4 * (size_t)L))) return rc;
    if ((rc = g->ok.ensure(ctx, L))) return rc;
    {
        CVB_PROF(ctx, "k_triangulate", 120.0 * nobs);
        k_triangulate<<<cdiv(L, 128), 128, 0, ctx->stream>>>((const cvb_pose *)g->poses.p, (const double *)g->a.p, (const uint32_t *)g->offsets.p, L,
                                                             (double *)g->out.p, (uint8_t *)g->ok.p);
        CVB_LAUNCH_CHECK(ctx);
    }
    CVB_CUDA(ctx, cudaMemcpyAsync(xyzw_out, g->out.p, sizeof(double) * 4 * (size_t)L, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cudaMemcpyAsync(ok_out, g->ok.p, L, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    return 0;
}

int cvb_arrsac_eight_point(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, const double *a, const double *b, uint32_t n, cvb_rng *rng,
                           cvb_pose *model_out, uint32_t *inliers_out, uint32_t cap, uint32_t *n_inliers, int32_t *found) {
    if (!ctx) return CVB_EINVAL;
    if (!cfg || !rng || !model_out || !found || (n && (!a || !b))) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    if (arrsac_on_host()) return arrsac_run(ctx, cfg, 0, a, b, n, rng, model_out, inliers_out, cap, n_inliers, found);
    return arrsac_host(ctx, cfg, 0, a, b, n, rng, model_out, inliers_out, cap, n_inliers, found);
}
int cvb_arrsac_five_point(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, const double *a, const double *b, uint32_t n, cvb_rng *rng,
                          int32_t eigenvector_row0, cvb_pose *model_out, uint32_t *inliers_out, uint32_t cap, uint32_t *n_inliers,
                          int32_t *found) {
    if (!ctx) return CVB_EINVAL;
    if (!cfg || !rng || !model_out || !found || (n && (!a || !b))) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    if (eigenvector_row0 != 5 && eigenvector_row0 != 6) return cvb_set_error(ctx, CVB_EINVAL, "eigenvector_row0 must be 5 (reference) or 6 (corrected)");
    if (arrsac_on_host()) return arrsac_run(ctx, cfg, 2, a, b, n, rng, model_out, inliers_out, cap, n_inliers, found, eigenvector_row0);
    return arrsac_host(ctx, cfg, 2, a, b, n, rng, model_out, inliers_out, cap, n_inliers, found, eigenvector_row0);
}
int cvb_arrsac_p3p(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, const double *bearings, const double *world, uint32_t n, cvb_rng *rng,
                   cvb_pose *model_out, uint32_t *inliers_out, uint32_t cap, uint32_t *n_inliers, int32_t *found) {
    if (!ctx) return CVB_EINVAL;
    if (!cfg || !rng || !model_out || !found || (n && (!bearings || !world))) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    if (arrsac_on_host()) return arrsac_run(ctx, cfg, 1, bearings, world, n, rng, model_out, inliers_out, cap, n_inliers, found);
    return arrsac_host(ctx, cfg, 1, bearings, world, n, rng, model_out, inliers_out, cap, n_inliers, found);
}

// ---- device-resident entry points (asynchronous on the context stream) -----------------------------------------------------
int cvb_pair_bearings_dev(cvb_ctx *ctx, const cvb_keypoint *kp_a_dev, const cvb_keypoint *kp_b_dev, const uint32_t *pairs_dev,
                          const uint32_t *n_pairs_dev, uint32_t cap, const cvb_intrinsics *intrinsics, double *a_out_dev, double *b_out_dev) {
    if (!ctx) return CVB_EINVAL;
    if (!kp_a_dev || !kp_b_dev || !pairs_dev || !n_pairs_dev || !intrinsics || !a_out_dev || !b_out_dev) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    if (cap == 0) return 0;
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    CVB_PROF(ctx, "k_pair_bearings", 0);
    k_pair_bearings<<<cdiv(cap, 256), 256, 0, ctx->stream>>>(kp_a_dev, kp_b_dev, pairs_dev, n_pairs_dev, cap, *intrinsics, a_out_dev, b_out_dev);
    CVB_LAUNCH_CHECK(ctx);
    return 0;
}

static int arrsac_dev_entry(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, int kind, const double *a_dev, const double *b_dev, const uint32_t *n_dev,
                            uint32_t n_max, const cvb_rng *rng, cvb_pose *model_out_dev, uint32_t *inliers_out_dev, uint32_t cap,
                            uint32_t *n_inliers_dev, int32_t *found_dev) {
    if (!ctx) return CVB_EINVAL;
    if (!cfg || !rng || !a_dev || !b_dev || !n_dev || !model_out_dev || !n_inliers_dev || !found_dev) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    return arrsac_run_dev(ctx, cfg, kind, a_dev, b_dev, n_dev, 0, n_max, rng, model_out_dev, inliers_out_dev, cap, n_inliers_dev, found_dev, 5);
}
int cvb_arrsac_eight_point_dev(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, const double *a_dev, const double *b_dev, const uint32_t *n_dev,
                               uint32_t n_max, const cvb_rng *rng, cvb_pose *model_out_dev, uint32_t *inliers_out_dev, uint32_t cap,
                               uint32_t *n_inliers_dev, int32_t *found_dev) {
    return arrsac_dev_entry(ctx, cfg, 0, a_dev, b_dev, n_dev, n_max, rng, model_out_dev, inliers_out_dev, cap, n_inliers_dev, found_dev);
}
int cvb_arrsac_p3p_dev(cvb_ctx *ctx, const cvb_arrsac_cfg *cfg, const double *bearings_dev, const double *world_dev, const uint32_t *n_dev,
                       uint32_t n_max, const cvb_rng *rng, cvb_pose *model_out_dev, uint32_t *inliers_out_dev, uint32_t cap,
                       uint32_t *n_inliers_dev, int32_t *found_dev) {
    return arrsac_dev_entry(ctx, cfg, 1, bearings_dev, world_dev, n_dev, n_max, rng, model_out_dev, inliers_out_dev, cap, n_inliers_dev, found_dev);
}
int cvb_arrsac_commit_rng(cvb_ctx *ctx, cvb_rng *rng, uint32_t *stats_out) {
    if (!ctx) return CVB_EINVAL;
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    ArrsacCtl h;
    int rc = arrsac_commit_rng(ctx, rng, &h);
    if (rc) return rc;
    if (getenv("CVB_ARS_DEBUG"))
        fprintf(stderr, "[arrsac] n %u models %u pass %u chunks %u turns %u repairs %u lazy %u | sprt us: order %u walk %u commit %u | block iterations %u\n", h.n, h.Mv, h.npass,
                h.stat_chunks, h.stat_turns, h.stat_repairs, h.stat_lazy, h.stat_perm_us, h.stat_walk_us, h.stat_commit_us, h.iters);
    if (stats_out) {   // 16 words (13..15 reserved): n, valid initial models, SPRT passes, SPRT commit rounds, block iterations, draws, inliers, found, 32-datum units
                       // scored in stage 1 / stage 2, predicates resolved exactly from the queues, mask words computed by the SPRT itself, SPRT repairs
        stats_out[0] = h.n; stats_out[1] = h.Mv; stats_out[2] = h.npass; stats_out[3] = h.stat_chunks; stats_out[4] = h.iters;
        stats_out[5] = (uint32_t)h.rng_pos; stats_out[6] = h.n_inliers; stats_out[7] = h.found;
        stats_out[8] = h.stat_units0; stats_out[9] = h.stat_units2; stats_out[10] = h.q_count + h.q_count2; stats_out[11] = h.stat_lazy;
        stats_out[12] = h.stat_repairs; stats_out[13] = h.stat_pad /* data walked by the box walks */; stats_out[14] = h.stat_walk_us; stats_out[15] = h.stat_commit_us;
    }
    return 0;
}

int cvb_single_view_optimize_l2(cvb_ctx *ctx, const cvb_pose *poses, uint32_t B, double optimization_rate, uint32_t iterations,
                                const double *bearings, const double *world, const uint32_t *offsets, cvb_pose *poses_out,
                                uint32_t *updates_out) {
    if (!ctx) return CVB_EINVAL;
    if (B == 0) return 0;
    if (!poses || !offsets || !poses_out) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    for (uint32_t b = 0; b < B; b++)
        if (offsets[b + 1] < offsets[b]) return cvb_set_error(ctx, CVB_EINVAL, "offsets must be non-decreasing");
    const uint32_t n = offsets[B];
    if (n && (!bearings || !world)) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    GeomWorkspace *g = gws(ctx);
    int rc;
    if ((rc = upload(ctx, g->poses, poses, sizeof(cvb_pose) * (size_t)B))) return rc;
    if ((rc = upload(ctx, g->a, bearings, sizeof(double) * 3 * (size_t)n))) return rc;
    if ((rc = upload(ctx, g->b, world, sizeof(double) * 4 * (size_t)n))) return rc;
    if ((rc = upload(ctx, g->offsets, offsets, sizeof(uint32_t) * ((size_t)B + 1)))) return rc;
    if ((rc = g->out.ensure(ctx, sizeof(cvb_pose) * (size_t)B))) return rc;
    if ((rc = g->ok.ensure(ctx, sizeof(uint32_t) * (size_t)B))) return rc;
    {
        CVB_PROF(ctx, "k_single_view_opt", 0.0);
        k_single_view_opt<<<B, OPT_NT, 0, ctx->stream>>>((const cvb_pose *)g->poses.p, (const double *)g->a.p, (const double *)g->b.p,
                                                         (const uint32_t *)g->offsets.p, optimization_rate, iterations, (cvb_pose *)g->out.p,
                                                         (uint32_t *)g->ok.p);
        CVB_LAUNCH_CHECK(ctx);
    }
    return download_poses_updates(ctx, g, poses_out, (size_t)B, updates_out, B);
}

int cvb_three_view_optimize_l2(cvb_ctx *ctx, const cvb_pose *poses, uint32_t B, int32_t adaptive, double optimization_rate,
                               uint32_t iterations, const double *observations, const uint32_t *offsets, cvb_pose *poses_out,
                               uint32_t *updates_out) {
    if (!ctx) return CVB_EINVAL;
    if (B == 0) return 0;
    if (!poses || !offsets || !poses_out) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    for (uint32_t b = 0; b < B; b++)
        if (offsets[b + 1] < offsets[b]) return cvb_set_error(ctx, CVB_EINVAL, "offsets must be non-decreasing");
    const uint32_t n = offsets[B];
    if (n && !observations) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    GeomWorkspace *g = gws(ctx);
    int rc;
    if ((rc = upload(ctx, g->poses, poses, sizeof(cvb_pose) * 2 * (size_t)B))) return rc;
    if ((rc = upload(ctx, g->a, observations, sizeof(double) * 9 * (size_t)n))) return rc;
    if ((rc = upload(ctx, g->offsets, offsets, sizeof(uint32_t) * ((size_t)B + 1)))) return rc;
    if ((rc = g->out.ensure(ctx, sizeof(cvb_pose) * 2 * (size_t)B))) return rc;
    if ((rc = g->ok.ensure(ctx, sizeof(uint32_t) * (size_t)B))) return rc;
    {
        CVB_PROF(ctx, "k_three_view_opt", 0.0);
        k_three_view_opt<<<B, OPT_NT, 0, ctx->stream>>>((const cvb_pose *)g->poses.p, (const double *)g->a.p, (const uint32_t *)g->offsets.p,
                                                        adaptive ? 1 : 0, optimization_rate, iterations, (cvb_pose *)g->out.p, (uint32_t *)g->ok.p);
        CVB_LAUNCH_CHECK(ctx);
    }
    return download_poses_updates(ctx, g, poses_out, 2 * (size_t)B, updates_out, B);
}

int cvb_observation_losses(cvb_ctx *ctx, const cvb_pose *poses, const double *bearings, const uint32_t *offsets, uint32_t L,
                           double *loss_out) {
    if (!ctx) return CVB_EINVAL;
    if (L == 0) return 0;
    if (!poses || !bearings || !offsets || !loss_out) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    for (uint32_t l = 0; l < L; l++)
        if (offsets[l + 1] < offsets[l]) return cvb_set_error(ctx, CVB_EINVAL, "offsets must be non-decreasing");
    const uint32_t nobs = offsets[L];
    if (nobs == 0) return 0;
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    GeomWorkspace *g = gws(ctx);
    int rc;
    if ((rc = upload(ctx, g->poses, poses, sizeof(cvb_pose) * (size_t)nobs))) return rc;
    if ((rc = upload(ctx, g->a, bearings, sizeof(double) * 3 * (size_t)nobs))) return rc;
    if ((rc = upload(ctx, g->offsets, offsets, sizeof(uint32_t) * ((size_t)L + 1)))) return rc;
    if ((rc = g->out.ensure(ctx, sizeof(double) * (size_t)nobs))) return rc;
    {
        CVB_PROF(ctx, "k_observation_losses", 128.0 * nobs);
        k_observation_losses<<<cdiv(L, 128), 128, 0, ctx->stream>>>((const cvb_pose *)g->poses.p, (const double *)g->a.p,
                                                                    (const uint32_t *)g->offsets.p, L, (double *)g->out.p);
        CVB_LAUNCH_CHECK(ctx);
    }
    CVB_CUDA(ctx, cudaMemcpyAsync(loss_out, g->out.p, sizeof(double) * (size_t)nobs, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    return 0;
}

int cvb_tri_landmarks_robust(cvb_ctx *ctx, const cvb_pose *first_pose, const cvb_pose *second_pose, const double *observations, uint32_t n,
                             double maximum_cosine_distance, double incidence_minimum_cosine_distance, uint8_t *robust_out) {
    if (!ctx) return CVB_EINVAL;
    if (n == 0) return 0;
    if (!first_pose || !second_pose || !observations || !robust_out) return cvb_set_error(ctx, CVB_EINVAL, "null argument");
    CVB_CUDA(ctx, cudaSetDevice(ctx->device));
    GeomWorkspace *g = gws(ctx);
    int rc;
    if ((rc = upload(ctx, g->a, observations, sizeof(double) * 9 * (size_t)n))) return rc;
    if ((rc = g->ok.ensure(ctx, n))) return rc;
    {
        CVB_PROF(ctx, "k_tri_landmark_robust", 72.0 * n);
        k_tri_landmark_robust<<<cdiv(n, 128), 128, 0, ctx->stream>>>(*first_pose, *second_pose, (const double *)g->a.p, n, maximum_cosine_distance,
                                                                     incidence_minimum_cosine_distance, (uint8_t *)g->ok.p);
        CVB_LAUNCH_CHECK(ctx);
    }
    CVB_CUDA(ctx, cudaMemcpyAsync(robust_out, g->ok.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    CVB_CUDA(ctx, cvb_wait(ctx, ctx->stream));
    return 0;
}

}  // extern "C"
