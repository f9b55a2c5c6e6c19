/* oracle/ref_optimize.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the steps that follow the consensus stage in cv-sfm:
 *   - cv-optimize/src/single_view_optimizer.rs:4-14,80-135   single_view_simple_optimize_l2
 *   - cv-optimize/src/three_view_optimizer.rs:7-21,126-272    three_view_simple_optimize_l2, three_view_adaptive_optimize_l2
 *   - cv-geom/src/epipolar.rs:8-50,53-71,85-176,193-232       sine-L1 two-view point, rotation gradient, three_view_gradients,
 *                                                             world_pose_gradient, loss
 *   - cv-core/src/so3.rs:16-100                               Se3TangentSpace (NaN -> zero, isometry(), scale)
 *   - cv-sfm/src/lib.rs:1306-1360,2570-2655                   is_bi_landmark_robust, is_tri_landmark_robust, observation_loss,
 *                                                             is_observation_consistent
 * nalgebra pieces restated from their documented formulas: Rotation3::from_scaled_axis (Rodrigues, axis = v/|v|, identity
 * when |v| == 0), IsometryMatrix3 product t = a.t + a.R b.t, inverse t' = R^T (-t).  Sums run in landmark order like the
 * reference's `for` loops.  The reference holds no test for these functions: parity unpinned beyond this restatement.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file. */
#include "ref_geom.h"
#include <math.h>
#include <string.h>

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double *a) { return sqrt(dot3(a, a)); }
static void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static void rotv(const double *R, const double *v, double *o) { for (int r = 0; r < 3; r++) o[r] = dot3(R + 3 * r, v); }
static void rotTv(const double *R, const double *v, double *o) {
    for (int c = 0; c < 3; c++) o[c] = R[c] * v[0] + R[3 + c] * v[1] + R[6 + c] * v[2];
}
static int any_nan3(const double *v) { return isnan(v[0]) || isnan(v[1]) || isnan(v[2]); }
static void normalize3(const double *v, double *o) { double n = norm3(v); o[0] = v[0] / n; o[1] = v[1] / n; o[2] = v[2] / n; }

/* Projective::from_homogeneous (cv-core/src/point.rs:20-25) */
static void from_homogeneous(double *p) {
    if (signbit(p[3])) for (int i = 0; i < 4; i++) p[i] = -p[i];
    double n = norm3(p);
    for (int i = 0; i < 4; i++) p[i] /= n;
}

/* nalgebra Rotation3::from_scaled_axis -> from_axis_angle */
static void rot_from_scaled_axis(const double *v, double *R) {
    const double angle = norm3(v);
    if (angle == 0.0) { memset(R, 0, 72); R[0] = R[4] = R[8] = 1.0; return; }
    const double ux = v[0] / angle, uy = v[1] / angle, uz = v[2] / angle;
    const double sqx = ux * ux, sqy = uy * uy, sqz = uz * uz, s = sin(angle), c = cos(angle), omc = 1.0 - c;
    R[0] = sqx + (1.0 - sqx) * c; R[1] = ux * uy * omc - uz * s; R[2] = ux * uz * omc + uy * s;
    R[3] = ux * uy * omc + uz * s; R[4] = sqy + (1.0 - sqy) * c; R[5] = uy * uz * omc - ux * s;
    R[6] = ux * uz * omc - uy * s; R[7] = uy * uz * omc + ux * s; R[8] = sqz + (1.0 - sqz) * c;
}
/* pose <- Se3TangentSpace{trans, rot}.isometry() * pose   (so3.rs:57-60) */
static void apply_delta(const double *trans, const double *rot, ref_pose *P) {
    double Rd[9], td[3], Rn[9], tn[3];
    rot_from_scaled_axis(rot, Rd);
    rotv(Rd, trans, td);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Rn[3 * r + c] = Rd[3 * r] * P->R[c] + Rd[3 * r + 1] * P->R[3 + c] + Rd[3 * r + 2] * P->R[6 + c];
    rotv(Rd, P->t, tn);
    for (int r = 0; r < 3; r++) tn[r] = td[r] + tn[r];
    memcpy(P->R, Rn, 72); memcpy(P->t, tn, 24);
}
static void pose_inverse(const ref_pose *P, ref_pose *o) {
    double nt[3] = {-P->t[0], -P->t[1], -P->t[2]}, R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = P->R[3 * c + r];
    rotv(R, nt, o->t); memcpy(o->R, R, 72);
}
/* Se3TangentSpace::new: a vector holding a NaN becomes zero */
static void tangent_new(double *t, double *r) {
    if (any_nan3(t)) t[0] = t[1] = t[2] = 0.0;
    if (any_nan3(r)) r[0] = r[1] = r[2] = 0.0;
}

/* epipolar.rs:193-198 */
void ref_world_pose_gradient(const double *translation, const double *b, double *tg, double *rg) {
    const double d = dot3(translation, b);
    double nt[3];
    for (int i = 0; i < 3; i++) tg[i] = d * b[i] - translation[i];
    normalize3(translation, nt);
    cross3(nt, b, rg);
    tangent_new(tg, rg);
}

/* single_view_optimizer.rs:4-14: None when the transformed point has w == 0 */
static int landmark_delta(const ref_pose *P, const double *bearing, const double *world, double *tg, double *rg) {
    double q[4];
    for (int r = 0; r < 3; r++) q[r] = dot3(P->R + 3 * r, world) + P->t[r] * world[3];
    q[3] = world[3];
    from_homogeneous(q);
    if (q[3] == 0.0) return 0;
    double p[3] = {q[0] / q[3], q[1] / q[3], q[2] / q[3]};
    ref_world_pose_gradient(p, bearing, tg, rg);
    return 1;
}

/* single_view_optimizer.rs:80-135; returns the number of pose updates applied */
uint32_t ref_single_view_optimize_l2(ref_pose *pose, double rate, uint32_t iterations, const double *bearings, const double *world, uint32_t n) {
    if (n == 0) return 0;
    double best_t = INFINITY, best_r = INFINITY;
    uint32_t no_improve = 0, updates = 0;
    const double inv_len = 1.0 / (double)n;
    for (uint32_t it = 0; it < iterations; it++) {
        double st[3] = {0, 0, 0}, sr[3] = {0, 0, 0}, tg[3], rg[3];
        for (uint32_t i = 0; i < n; i++)
            if (landmark_delta(pose, bearings + 3 * (size_t)i, world + 4 * (size_t)i, tg, rg))
                for (int k = 0; k < 3; k++) { st[k] += tg[k]; sr[k] += rg[k]; }
        double dt[3], dr[3];
        for (int k = 0; k < 3; k++) { dt[k] = (st[k] * inv_len) * rate; dr[k] = (sr[k] * inv_len) * rate; }
        no_improve++;
        const double t = norm3(st), r = norm3(sr);
        if (best_t > t) { best_t = t; no_improve = 0; }
        if (best_r > r) { best_r = r; no_improve = 0; }
        if (no_improve >= 50) break;
        apply_delta(dt, dr, pose); updates++;
        if (it == iterations - 1) break;
    }
    return updates;
}

/* epipolar.rs:8-50: `t` goes from B to A, the point has A as origin; returns 0 for None */
static int sine_l1_point(const double *t, const double *a_in, const double *b_in, double *p) {
    double ca[3], cb[3], na[3], nb[3], a[3], b[3];
    cross3(a_in, t, ca); const double can = norm3(ca); for (int i = 0; i < 3; i++) na[i] = ca[i] / can;
    cross3(b_in, t, cb); const double cbn = norm3(cb); for (int i = 0; i < 3; i++) nb[i] = cb[i] / cbn;
    memcpy(a, a_in, 24); memcpy(b, b_in, 24);
    if (can < cbn) { double d = dot3(a_in, nb), v[3]; for (int i = 0; i < 3; i++) v[i] = a_in[i] - d * nb[i]; normalize3(v, a); }
    else { double d = dot3(b_in, na), v[3]; for (int i = 0; i < 3; i++) v[i] = b_in[i] - d * na[i]; normalize3(v, b); }
    double z[3], tb[3];
    cross3(a, b, z); cross3(t, b, tb);
    double q[4] = {a[0], a[1], a[2], dot3(z, z) / dot3(z, tb)};
    from_homogeneous(q);
    for (int i = 0; i < 4; i++) if (!isfinite(q[i])) return 0;
    if (signbit(dot3(q, a)) || signbit(dot3(q, b))) return 0;
    if (q[3] == 0.0) return 0;
    for (int i = 0; i < 3; i++) p[i] = q[i] / q[3];
    return 1;
}
/* epipolar.rs:53-71 */
static void rotation_gradient(const double *t, const double *a, const double *b, double *o) {
    double ca[3], cb[3], na[3], nb[3];
    cross3(a, t, ca); cross3(b, t, cb);
    normalize3(ca, na); normalize3(cb, nb);
    cross3(nb, na, o);
}
/* epipolar.rs:85-176: out = [first.t, first.r, second.t, second.r] */
void ref_three_view_gradients(const double *c, const double *f, const double *ftoc, const double *s, const double *stoc, double *out) {
    double stof[3], rcf[3], rcs[3], rfs[3], p[3], q[3], tf[3] = {0, 0, 0}, ts[3] = {0, 0, 0}, tc[3] = {0, 0, 0}, neg[3];
    for (int i = 0; i < 3; i++) stof[i] = stoc[i] - ftoc[i];
    rotation_gradient(ftoc, c, f, rcf); rotation_gradient(stoc, c, s, rcs); rotation_gradient(stof, f, s, rfs);
    double *ft = out, *fr = out + 3, *st = out + 6, *sr = out + 9;
    for (int i = 0; i < 3; i++) {
        fr[i] = rcf[i] * (2.0 / 3.0) + (-rfs[i]) * (1.0 / 3.0);
        sr[i] = rcs[i] * (2.0 / 3.0) + rfs[i] * (1.0 / 3.0);
    }
    for (int i = 0; i < 3; i++) neg[i] = -stoc[i];
    if (sine_l1_point(neg, c, s, p)) { for (int i = 0; i < 3; i++) q[i] = p[i] - ftoc[i]; double d = dot3(q, f); for (int i = 0; i < 3; i++) tf[i] = q[i] - d * f[i]; }
    for (int i = 0; i < 3; i++) neg[i] = -ftoc[i];
    if (sine_l1_point(neg, c, f, p)) { for (int i = 0; i < 3; i++) q[i] = p[i] - stoc[i]; double d = dot3(q, s); for (int i = 0; i < 3; i++) ts[i] = q[i] - d * s[i]; }
    for (int i = 0; i < 3; i++) neg[i] = -stof[i];
    if (sine_l1_point(neg, f, s, p)) { for (int i = 0; i < 3; i++) q[i] = p[i] + ftoc[i]; double d = dot3(q, c); for (int i = 0; i < 3; i++) tc[i] = d * c[i] - q[i]; }
    for (int i = 0; i < 3; i++) {
        ft[i] = tf[i] * (2.0 / 3.0) + tc[i] * (1.0 / 3.0);
        st[i] = ts[i] * (2.0 / 3.0) + tc[i] * (1.0 / 3.0);
    }
    tangent_new(ft, fr); tangent_new(st, sr);
}
/* three_view_optimizer.rs:7-21: poses are the inverted (first -> centre, second -> centre) isometries; obs = [c, f, s] */
static void landmark_gradients(const ref_pose *P, const double *obs, double *out) {
    double f[3], s[3];
    rotv(P[0].R, obs + 3, f); rotv(P[1].R, obs + 6, s);
    ref_three_view_gradients(obs, f, P[0].t, s, P[1].t, out);
}
/* three_view_optimizer.rs:126-201 (adaptive = 0) and :203-272 (adaptive = 1; `rate` unused); returns the pose updates applied */
uint32_t ref_three_view_optimize_l2(ref_pose poses[2], int adaptive, double rate, uint32_t iterations, const double *obs, uint32_t n) {
    if (n == 0) return 0;
    const double inv_len = 1.0 / (double)n;
    ref_pose P[2];
    pose_inverse(&poses[0], &P[0]); pose_inverse(&poses[1], &P[1]);
    double best[2][2] = {{INFINITY, INFINITY}, {INFINITY, INFINITY}};
    uint32_t no_improve = 0, updates = 0;
    for (uint32_t it = 0; it < iterations; it++) {
        double sum[12] = {0}, tv[2] = {0, 0}, rv[2] = {0, 0}, g[12];
        for (uint32_t i = 0; i < n; i++) {
            landmark_gradients(P, obs + 9 * (size_t)i, g);
            for (int k = 0; k < 12; k++) sum[k] += g[k];
            if (adaptive) for (int v = 0; v < 2; v++) { tv[v] += norm3(g + 6 * v); rv[v] += norm3(g + 6 * v + 3); }
        }
        double d[12];
        if (!adaptive) {
            const double sc = inv_len * rate;
            for (int k = 0; k < 12; k++) d[k] = sum[k] * sc;
            no_improve++;
            for (int v = 0; v < 2; v++) {
                const double t = norm3(sum + 6 * v), r = norm3(sum + 6 * v + 3);
                if (best[v][0] > t) { best[v][0] = t; no_improve = 0; }
                if (best[v][1] > r) { best[v][1] = r; no_improve = 0; }
            }
            if (no_improve >= 50) break;
        } else {
            for (int v = 0; v < 2; v++) {
                double l2[6];
                for (int k = 0; k < 6; k++) l2[k] = sum[6 * v + k] * inv_len;
                const double tstd = tv[v] * inv_len, rstd = rv[v] * inv_len;
                double trate = norm3(l2) / tstd, rrate = norm3(l2 + 3) / rstd;
                if (!isfinite(trate)) trate = 0.0;
                if (!isfinite(rrate)) rrate = 0.0;
                for (int k = 0; k < 3; k++) { d[6 * v + k] = l2[k] * trate; d[6 * v + 3 + k] = l2[3 + k] * rrate; }
            }
        }
        apply_delta(d, d + 3, &P[0]); apply_delta(d + 6, d + 9, &P[1]); updates++;
        if (it == iterations - 1) break;
    }
    pose_inverse(&P[0], &poses[0]); pose_inverse(&P[1], &poses[1]);
    return updates;
}

/* epipolar.rs:200-232: |sine| of the angle between the epipolar planes, 1.0 on NaN or failed cheirality */
double ref_epipolar_loss(const double *t, const double *a, const double *b) {
    double ca[3], cb[3];
    cross3(a, t, ca); cross3(b, t, cb);
    const double na2 = dot3(ca, ca), nb2 = dot3(cb, cb);
    double res;
    if (na2 < nb2) { const double sc = 1.0 / sqrt(nb2); double v[3] = {cb[0] * sc, cb[1] * sc, cb[2] * sc}; res = fabs(dot3(a, v)); }
    else { const double sc = 1.0 / sqrt(na2); double v[3] = {ca[0] * sc, ca[1] * sc, ca[2] * sc}; res = fabs(dot3(b, v)); }
    if (isnan(res) || signbit(dot3(a, b))) return 1.0;
    return res;
}

static void pose_mul(const ref_pose *A, const ref_pose *B, ref_pose *o) { /* A * B */
    ref_pose r;
    for (int i = 0; i < 3; i++) for (int c = 0; c < 3; c++) r.R[3 * i + c] = A->R[3 * i] * B->R[c] + A->R[3 * i + 1] * B->R[3 + c] + A->R[3 * i + 2] * B->R[6 + c];
    double sh[3]; rotv(A->R, B->t, sh);
    for (int i = 0; i < 3; i++) r.t[i] = A->t[i] + sh[i];
    *o = r;
}
static double transformed_cosine_distance(const ref_pose *P, const double *point_h, const double *bearing) {
    double q[4];
    for (int r = 0; r < 3; r++) q[r] = dot3(P->R + 3 * r, point_h) + P->t[r] * point_h[3];
    q[3] = point_h[3];
    from_homogeneous(q);
    return 1.0 - dot3(q, bearing);
}

/* cv-sfm/src/lib.rs:2570-2620 observation_loss for every observation of one landmark (poses are WorldToCamera) */
void ref_observation_losses(const ref_pose *poses, const double *bearings, uint32_t n, double *loss) {
    if (n == 1) { loss[0] = 2.0; return; }
    if (n == 2) {
        ref_pose inv, tot; double fb[3];
        pose_inverse(&poses[0], &inv); pose_mul(&poses[1], &inv, &tot);
        rotv(tot.R, bearings, fb);
        const double l = 1.0 - cos(asin(ref_epipolar_loss(tot.t, fb, bearings + 3)));
        loss[0] = loss[1] = l;
        return;
    }
    double p[4];
    if (!ref_triangulate_linear_eigen(poses, bearings, (int)n, p)) { for (uint32_t i = 0; i < n; i++) loss[i] = 2.0; return; }
    for (uint32_t i = 0; i < n; i++) loss[i] = transformed_cosine_distance(&poses[i], p, bearings + 3 * (size_t)i);
}

/* cv-sfm/src/lib.rs:1320-1360 (poses are CameraToCamera centre -> first / second) */
int ref_is_tri_landmark_robust(const ref_pose *first, const ref_pose *second, const double *c, const double *f, const double *s,
                               double maximum_cosine_distance, double incidence_minimum_cosine_distance) {
    ref_pose P[3]; double B[9], p[4];
    memset(&P[0], 0, sizeof(ref_pose)); P[0].R[0] = P[0].R[4] = P[0].R[8] = 1.0;
    P[1] = *first; P[2] = *second;
    memcpy(B, c, 24); memcpy(B + 3, f, 24); memcpy(B + 6, s, 24);
    if (!ref_triangulate_linear_eigen(P, B, 3, p)) return 0;
    from_homogeneous(p);   /* CameraPoint::from_homogeneous(p.0) */
    double fc[3], sc[3];
    rotTv(first->R, f, fc); rotTv(second->R, s, sc);
    const int cosine_ok = 1.0 - dot3(p, c) < maximum_cosine_distance
        && transformed_cosine_distance(first, p, f) < maximum_cosine_distance
        && transformed_cosine_distance(second, p, s) < maximum_cosine_distance;
    const int incidence_ok = 1.0 - dot3(c, fc) > incidence_minimum_cosine_distance || 1.0 - dot3(c, sc) > incidence_minimum_cosine_distance
        || 1.0 - dot3(fc, sc) > incidence_minimum_cosine_distance;
    return cosine_ok && incidence_ok;
}
