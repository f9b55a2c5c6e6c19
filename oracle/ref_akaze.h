/* oracle/ref_akaze.h -- TEST INFRASTRUCTURE (CPU oracle), not product code. See ref_akaze.c. */
#ifndef REF_AKAZE_H
#define REF_AKAZE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* mirrors akaze::Akaze (akaze/src/lib.rs:109-142); maximum_features < 0 == usize::MAX */
typedef struct {
    int64_t maximum_features;
    uint32_t num_sublevels;
    uint32_t max_octave_evolution;
    double base_scale_offset;
    double initial_contrast; /* never read by the reference (lib.rs:123,176) */
    double contrast_percentile;
    uint64_t contrast_factor_num_bins;
    double derivative_factor;
    double detector_threshold;
    uint64_t descriptor_channels;
    uint64_t descriptor_pattern_size;
} ref_akaze_cfg;

/* mirrors akaze::KeyPoint (akaze/src/lib.rs:71-93) */
typedef struct {
    float x, y;
    float response;
    float size;
    float angle;
    uint32_t octave;
    uint32_t class_id;
} ref_keypoint;

struct ref_akaze;
void ref_akaze_default_cfg(ref_akaze_cfg *c);
struct ref_akaze *ref_akaze_create(const ref_akaze_cfg *cfg);
void ref_akaze_destroy(struct ref_akaze *A);
int ref_akaze_extract(struct ref_akaze *A, const float *image, int w, int h);
int ref_akaze_num_evolutions(const struct ref_akaze *A);
int ref_akaze_evolution_info(const struct ref_akaze *A, int i, int *w, int *h, uint32_t *octave, double *esigma,
                             int *ntau, double *tau);
double ref_akaze_contrast_factor(const struct ref_akaze *A);
const float *ref_akaze_plane(const struct ref_akaze *A, int i, int plane);
int ref_akaze_stage(const struct ref_akaze *A, int stage, const ref_keypoint **out);
const uint8_t *ref_akaze_descriptors(const struct ref_akaze *A);

void ref_horizontal_filter(const float *in, int w, int h, const float *k, int ks, float *out);
void ref_vertical_filter(const float *in, int w, int h, const float *k, int ks, float *out);
void ref_gaussian_kernel(float r, int ks, float *out);
void ref_half_size(const float *in, int w, int h, float *out);
int ref_fed_tau(double T, double tau_max, double *out, int cap);
void ref_set_num_threads(int n);
float ref_sinf(float x);
float ref_cosf(float x);
float ref_atan2f(float y, float x);
float ref_fast_atan2_equiv(float y, float x);

/* ref_match.c: space::LinearKnn + bitarray::Hamming restatement */
void ref_hamming_knn(const uint8_t *q, uint32_t n, const uint8_t *db, uint32_t m, uint32_t k, uint32_t *idx_out,
                     uint32_t *dist_out);
#ifdef __cplusplus
}
#endif
#endif
