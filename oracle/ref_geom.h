/* oracle/ref_geom.h -- TEST INFRASTRUCTURE (CPU oracle), not product code. See ref_geom.c. */
#ifndef REF_GEOM_H
#define REF_GEOM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* IsometryMatrix3<f64>: rotation row-major + translation */
typedef struct { double R[9]; double t[3]; } ref_pose;
typedef struct { int kind; uint64_t s[4]; } ref_rng;   /* kind 0 xoshiro256++, 1 pcg64 */
typedef struct {
    double inlier_threshold;
    uint32_t initialization_hypotheses, initialization_blocks, max_candidate_hypotheses, estimations_per_block, block_size;
    float likelihood_ratio_threshold, initial_epsilon, initial_delta;
} ref_arrsac_cfg;

int ref_sym_eigen(int n, const double *A, double eps, int max_sweeps, double *d, double *V);
int ref_sym_eigen9_rr(const double *A, double eps, int max_sweeps, double *d, double *V);
int ref_svd3(const double *M, double eps, int max_iter, double *U, double *s, double *Vt);
int ref_eight_point_essential(const double *a, const double *b, double eps, int iters, double *E);
int ref_essential_poses(const double *E, double eps, int iters, ref_pose out[4]);
int ref_eight_point(const double *a, const double *b, ref_pose out[4]);
int ref_real_eigenvalues10(const double *A, double *wr, double *wi);
int ref_five_point_essentials(const double *a, const double *b, double *Es);
int ref_five_point(const double *a, const double *b, ref_pose *out);
void ref_five_point_set_row0(int r0);
void ref_fp_o1(const double *a, const double *b, double *r);   /* degree-1 x degree-1 -> 20-term basis */
void ref_fp_o2(const double *a, const double *b, double *r);   /* degree-2 (20-term layout) x degree-1 */
double ref_essential_residual(const double *E, const double *a, const double *b);
double ref_residual_c2c(const ref_pose *P, const double *a, const double *b);
double ref_residual_w2c(const ref_pose *P, const double *bearing, const double *world);
int ref_p3p(const double *bearings, const double *world, ref_pose out[4]);
int ref_triangulate_linear_eigen(const ref_pose *poses, const double *bearings, int n, double *out);
void ref_calibrate(double fx, double fy, double cx, double cy, double skew, double px, double py, double *bearing);
void ref_rng_seed_xoshiro(ref_rng *r, uint64_t seed);
void ref_rng_seed_pcg64(ref_rng *r, const uint8_t seed[32]);
uint32_t ref_rng_next_u32(ref_rng *r);
void ref_arrsac_default_cfg(ref_arrsac_cfg *c, double inlier_threshold);
int ref_arrsac(const ref_arrsac_cfg *cfg, int kind, const double *a, const double *b, uint32_t n, ref_rng *rng,
               ref_pose *model_out, uint32_t *inliers_out, uint32_t *n_inliers);
/* ref_optimize.c: post-consensus refinement and robustness checks */
void ref_world_pose_gradient(const double *translation, const double *b, double *tg, double *rg);
uint32_t ref_single_view_optimize_l2(ref_pose *pose, double rate, uint32_t iterations, const double *bearings, const double *world, uint32_t n);
void ref_three_view_gradients(const double *c, const double *f, const double *ftoc, const double *s, const double *stoc, double *out);
uint32_t ref_three_view_optimize_l2(ref_pose poses[2], int adaptive, double rate, uint32_t iterations, const double *obs, uint32_t n);
double ref_epipolar_loss(const double *t, const double *a, const double *b);
void ref_observation_losses(const ref_pose *poses, const double *bearings, uint32_t n, double *loss);
int ref_is_tri_landmark_robust(const ref_pose *first, const ref_pose *second, const double *c, const double *f, const double *s,
                               double maximum_cosine_distance, double incidence_minimum_cosine_distance);
#ifdef __cplusplus
}
#endif
#endif
