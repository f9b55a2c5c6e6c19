/* Test infrastructure: the product's exact-predicate filter (cv_b200/csrc/c2c_filter.cuh) compiled for the host so that the CPU
 * tests can compare every decision with the oracle's exact residual.  Not part of the product (the library never runs this). */
#include <stddef.h>
#include <stdint.h>
#include "../../cv_b200/csrc/c2c_filter.cuh"

/* poses: m x 12 doubles (R row-major, t); out[p*n + i] = 1 / 0 / -1 */
void c2c_filter_batch(const double *poses, uint32_t m, const double *a, const double *b, uint32_t n, double thr, int8_t *out) {
    for (uint32_t p = 0; p < m; p++)
        for (uint32_t i = 0; i < n; i++)
            out[(size_t)p * n + i] = (int8_t)c2c_inlier_filter(poses + 12 * (size_t)p, poses + 12 * (size_t)p + 9, a + 3 * (size_t)i, b + 3 * (size_t)i, thr);
}
