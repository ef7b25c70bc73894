/* CPU FP64 oracle, C part.  TEST INFRASTRUCTURE ONLY (same rules as nle_oracle.py: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it).
 *
 * Plain-C restatement of the reference's affinity loop, /root/reference/src/filter.cpp:104-112
 * (negativeWeightedDistance), :128-129 (weights 1/hx^2, 1/hy^2) and :130-145 (fill, then exp), written so that
 * the streaming oracle (nle_oracle.train_streaming) can afford the configurations the dense reference cannot hold
 * in host RAM (BASELINE configs[2..4]).  The NumPy formulation of the same loop (nle_oracle.affinity_block) builds
 * five p x tile temporaries on one core; this one evaluates each entry once and uses libm's scalar exp() (< 1 ulp).
 * It is single-threaded per call (libgomp is not in the image): nle_oracle.py calls it on disjoint sample-row
 * ranges from a thread pool (ctypes releases the GIL).  The arithmetic expression and its evaluation order are
 * the reference's:
 *
 *     K(i,j) = exp(-sw * double(int((ri-rj)^2 + (ci-cj)^2)) - pw * ((yi-yj) * (yi-yj)))
 *
 * Compile with -ffp-contract=off so that no FMA contraction changes the rounding of the argument
 * (oracle/build_c.py does).  tests/test_oracle.py pins it against nle_oracle.affinity_block.
 */
#include <math.h>
#include <stdint.h>

/* out[i * nb + j] = K(sample i, pixel j) for rows i0 <= i < i1 of the na x nb row-major block `out`.
 * (ra, ca, ya)[i] = row, column, luminance of sample i; (rb, cb, yb)[j] = the same for pixel j
 * (rows/columns are the divmod of the raster index by the image width, utils.hpp:11-19). */
void nle_oracle_affinity_block(const int64_t* ra, const int64_t* ca, const double* ya, int64_t i0, int64_t i1,
                               const int64_t* rb, const int64_t* cb, const double* yb, int64_t nb, double sw,
                               double pw, double* out) {
    for (int64_t i = i0; i < i1; ++i) {
        const int64_t ri = ra[i], ci = ca[i];
        const double yi = ya[i];
        double* o = out + i * nb;
        for (int64_t j = 0; j < nb; ++j) {
            const int64_t d2i = (ri - rb[j]) * (ri - rb[j]) + (ci - cb[j]) * (ci - cb[j]); /* filter.cpp:109 (int) */
            const double d2 = (double)d2i;
            const double dy = yi - yb[j];
            const double dy2 = dy * dy;                                        /* filter.cpp:110 */
            o[j] = exp(-sw * d2 - pw * dy2);                                   /* filter.cpp:111, :144-145 */
        }
    }
}

int nle_oracle_c_version(void) { return 2; }
