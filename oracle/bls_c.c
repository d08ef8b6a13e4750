/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the binned
 * "fast" Box-Least-Squares search that lightkurve reaches through
 *   astropy.timeseries.BoxLeastSquares(t, y, dy).power(period, duration)
 * at /root/reference/src/lightkurve/periodogram.py:1161-1169.
 *
 * The algorithm lives in astropy (>=5.0, timeseries/periodograms/bls/bls.c,
 * not vendored under /root/reference and not installable here); this file
 * restates its published structure: per trial period, histogram the samples
 * into bins of width min(duration)/oversample (index rule below), wrap-pad
 * `oversample` bins, inclusive prefix sum, then for every duration (in bins)
 * and every start bin evaluate the in/out-of-transit means and keep the first
 * strict maximum of the objective with y_out >= y_in.
 * PARITY UNPINNED for values (see oracle/__init__.py); pinned behaviourally by
 * the reference's tests/test_periodogram.py:264-361,434-442 ported in tests/.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC, NO -ffast-math:
 * the bin-index rule must be evaluated in strict IEEE fp64).
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* obj_flag: 0 = likelihood (astropy default), 1 = snr */
static void objective_terms(double y_in, double y_out, double ivar_in, double ivar_out,
                            double* depth, double* depth_err, double* depth_snr, double* log_like)
{
    *depth = y_out - y_in;
    *depth_err = sqrt(1.0 / ivar_in + 1.0 / ivar_out);
    *depth_snr = *depth / *depth_err;
    *log_like = 0.5 * ivar_in * (y_out - y_in) * (y_out - y_in);
}

/* Returns 0 ok, 1 bad period, 2 bad duration, -2 alloc failure. */
int oracle_bls_fast(int N, const double* t, const double* y, const double* ivar,
                    int n_periods, const double* periods,
                    int n_durations, const double* durations,
                    int oversample, int obj_flag,
                    double* best_objective, double* best_depth, double* best_depth_err,
                    double* best_duration, double* best_phase, double* best_depth_snr,
                    double* best_log_like,
                    int* best_bin /* optional [n_periods][2]: start bin n, duration-in-bins; may be NULL */)
{
    double max_period = periods[0], min_period = periods[0];
    for (int k = 1; k < n_periods; ++k) {
        if (periods[k] < min_period) min_period = periods[k];
        if (periods[k] > max_period) max_period = periods[k];
    }
    if (min_period < DBL_EPSILON) return 1;

    double min_duration = durations[0], max_duration = durations[0];
    for (int k = 1; k < n_durations; ++k) {
        if (durations[k] < min_duration) min_duration = durations[k];
        if (durations[k] > max_duration) max_duration = durations[k];
    }
    if ((max_duration > min_period) || (min_duration < DBL_EPSILON)) return 2;

    double bin_duration = min_duration / ((double)oversample);
    int max_n_bins = (int)(ceil(max_period / bin_duration)) + oversample;

    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    double* mean_y_0 = (double*)malloc((size_t)nthreads * (max_n_bins + 1) * sizeof(double));
    double* mean_ivar_0 = (double*)malloc((size_t)nthreads * (max_n_bins + 1) * sizeof(double));
    if (!mean_y_0 || !mean_ivar_0) { free(mean_y_0); free(mean_ivar_0); return -2; }

    double min_t = INFINITY, sum_y = 0.0, sum_ivar = 0.0;
    for (int n = 0; n < N; ++n) {
        min_t = fmin(min_t, t[n]);
        sum_y += y[n] * ivar[n];
        sum_ivar += ivar[n];
    }

#pragma omp parallel for schedule(dynamic, 8)
    for (int p = 0; p < n_periods; ++p) {
        int ithread = 0;
#ifdef _OPENMP
        ithread = omp_get_thread_num();
#endif
        double* mean_y = mean_y_0 + (size_t)ithread * (max_n_bins + 1);
        double* mean_ivar = mean_ivar_0 + (size_t)ithread * (max_n_bins + 1);
        double period = periods[p];
        int n_bins = (int)(ceil(period / bin_duration)) + oversample;

        for (int n = 0; n <= n_bins; ++n) { mean_y[n] = 0.0; mean_ivar[n] = 0.0; }

        for (int n = 0; n < N; ++n) {
            int ind = (int)(fabs(fmod(t[n] - min_t, period)) / bin_duration) + 1;
            mean_y[ind] += y[n] * ivar[n];
            mean_ivar[ind] += ivar[n];
        }

        for (int n = 1, ind = n_bins - oversample; n <= oversample; ++n, ++ind) {
            mean_y[ind] = mean_y[n];
            mean_ivar[ind] = mean_ivar[n];
        }

        for (int n = 1; n <= n_bins; ++n) {
            mean_y[n] += mean_y[n - 1];
            mean_ivar[n] += mean_ivar[n - 1];
        }

        best_objective[p] = -INFINITY;
        best_depth[p] = 0.0; best_depth_err[p] = 0.0; best_duration[p] = 0.0;
        best_phase[p] = 0.0; best_depth_snr[p] = 0.0; best_log_like[p] = 0.0;
        if (best_bin) { best_bin[2 * p] = -1; best_bin[2 * p + 1] = -1; }

        for (int k = 0; k < n_durations; ++k) {
            int dur = (int)(round(durations[k] / bin_duration));
            int n_max = n_bins - dur;
            for (int n = 0; n <= n_max; ++n) {
                double y_in = mean_y[n + dur] - mean_y[n];
                double ivar_in = mean_ivar[n + dur] - mean_ivar[n];
                double y_out = sum_y - y_in;
                double ivar_out = sum_ivar - ivar_in;
                if ((ivar_in < DBL_EPSILON) || (ivar_out < DBL_EPSILON)) continue;
                y_in /= ivar_in;
                y_out /= ivar_out;
                double depth, depth_err, depth_snr, log_like;
                objective_terms(y_in, y_out, ivar_in, ivar_out, &depth, &depth_err, &depth_snr, &log_like);
                double objective = obj_flag ? depth_snr : log_like;
                if (y_out >= y_in && objective > best_objective[p]) {
                    best_objective[p] = objective;
                    best_depth[p] = depth;
                    best_depth_err[p] = depth_err;
                    best_depth_snr[p] = depth_snr;
                    best_log_like[p] = log_like;
                    best_duration[p] = dur * bin_duration;
                    best_phase[p] = fmod(n * bin_duration + 0.5 * best_duration[p] + min_t, period);
                    if (best_bin) { best_bin[2 * p] = n; best_bin[2 * p + 1] = dur; }
                }
            }
        }
    }
    free(mean_y_0);
    free(mean_ivar_0);
    return 0;
}

/* Bin indices only (for the bit-exact index parity test). */
void oracle_bls_bin_index(int N, const double* t, double min_t, double period,
                          double bin_duration, int* ind_out)
{
    for (int n = 0; n < N; ++n)
        ind_out[n] = (int)(fabs(fmod(t[n] - min_t, period)) / bin_duration) + 1;
}
