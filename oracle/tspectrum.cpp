// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see oracle/README.md).
//
// CPU restatement of ROOT's TSpectrum::Search / TSpectrum::SearchHighRes as they are
// configured by the reference at /root/reference/TEST_2.C:187-188:
//     TSpectrum spec(12); spec.Search(hMF, 2, "nobackground,nodraw", 0.02)
// ROOT (hist/spectrum/src/TSpectrum.cxx, v6.30) is NOT vendored under /root/reference and
// is not installed here, so this file restates the published algorithm (M. Morhac et al.,
// NIM A 443 (2000) 108).  PARITY UNPINNED: the reference ships no
// golden vectors for this call and ROOT cannot be run in this environment.
//
// Arithmetic is kept in the order the ROOT source performs it (one IEEE op per source-level
// op, no FMA contraction: build with -ffp-contract=off) so the CUDA kernel can be compared
// bit-for-bit, intermediates included.
#include "npswf_oracle.h"
#include "det_exp.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace {

inline double exp_sel(double x, int libm) { return libm ? std::exp(x) : oracle_det_exp(x); }

}  // namespace

// TSpectrum::SearchHighRes(source, dest, ssize, sigma, threshold[%], backgroundRemove=false,
//                          deconIterations, markov=true, averWindow)
// Returns number of peaks (<= max_peaks); pos_x[] = fPositionX (fractional channel, sorted by
// raw height, descending).  Optional debug outputs: smoothed[ssize+2*shift] (W1 after the
// Markov step), decon[ssize] (destVector).
extern "C" int oracle_search_highres(const double *source, int ssize, double sigma, double threshold,
                                     int decon_iterations, int aver_window, int max_peaks,
                                     double *pos_x, double *smoothed_out, double *decon_out, int use_libm_exp)
{
    int i, j, number_iterations = (int)(7 * sigma + 0.5);
    double a, b;
    int k, lindex, posit, imin, imax, jmin, jmax, lh_gold, priz;
    double lda, ldb, ldc, area, maximum, maximum_decon;
    int xmin, xmax, l, peak_index = 0, size_ext = ssize + 2 * number_iterations, shift = number_iterations;
    double maxch, nom, nip, nim, sp, sm, plocha = 0;
    double m0low = 0, m1low = 0, m2low = 0, l0low = 0, l1low = 0, detlow;
    if (sigma < 1 || threshold <= 0 || threshold >= 100) return 0;
    if ((int)(5.0 * sigma + 0.5) >= 100 / 2) return 0;  // PEAK_WINDOW/2, "Too large sigma"
    if (aver_window <= 0) return 0;

    // edge slope of the first k channels, clamped to <= 0
    k = (int)(2 * sigma + 0.5);
    if (k >= 2) {
        for (i = 0; i < k; i++) {
            a = i; b = source[i];
            m0low += 1; m1low += a; m2low += a * a; l0low += b; l1low += a * b;
        }
        detlow = m0low * m2low - m1low * m1low;
        if (detlow != 0) l1low = (-l0low * m1low + l1low * m0low) / detlow;
        else l1low = 0;
        if (l1low > 0) l1low = 0;
    } else {
        l1low = 0;
    }

    std::vector<double> ws((size_t)7 * size_ext, 0.0);
    double *W = ws.data();
    const int S = size_ext;
    // extension
    for (i = 0; i < S; i++) {
        if (i < shift) {
            a = i - shift;
            W[i + S] = source[0] + l1low * a;
            if (W[i + S] < 0) W[i + S] = 0;
        } else if (i >= ssize + shift) {
            W[i + S] = source[ssize - 1];
            if (W[i + S] < 0) W[i + S] = 0;
        } else {
            W[i + S] = source[i - shift];
        }
    }
    // (backgroundRemove == false: skipped)
    for (i = 0; i < S; i++) W[i + 6 * S] = W[i + S];  // raw copy, used for thresholds + ordering

    // Markov smoothing (markov == true)
    {
        for (j = 0; j < S; j++) W[2 * S + j] = W[S + j];
        xmin = 0; xmax = S - 1;
        for (i = 0, maxch = 0; i < S; i++) {
            W[i] = 0;
            if (maxch < W[2 * S + i]) maxch = W[2 * S + i];
            plocha += W[2 * S + i];
        }
        if (maxch == 0) return 0;
        nom = 1;
        W[xmin] = 1;
        for (i = xmin; i < xmax; i++) {
            nip = W[2 * S + i] / maxch;
            nim = W[2 * S + i + 1] / maxch;
            sp = 0; sm = 0;
            for (l = 1; l <= aver_window; l++) {
                if ((i + l) > xmax) a = W[2 * S + xmax] / maxch;
                else a = W[2 * S + i + l] / maxch;
                b = a - nip;
                if (a + nip <= 0) a = 1;
                else a = std::sqrt(a + nip);
                b = b / a;
                b = exp_sel(b, use_libm_exp);
                sp = sp + b;
                if ((i - l + 1) < xmin) a = W[2 * S + xmin] / maxch;
                else a = W[2 * S + i - l + 1] / maxch;
                b = a - nim;
                if (a + nim <= 0) a = 1;
                else a = std::sqrt(a + nim);
                b = b / a;
                b = exp_sel(b, use_libm_exp);
                sm = sm + b;
            }
            a = sp / sm;
            a = W[i + 1] = W[i] * a;
            nom = nom + a;
        }
        for (i = xmin; i <= xmax; i++) W[i] = W[i] / nom;
        for (j = 0; j < S; j++) W[S + j] = W[j] * plocha;
        for (j = 0; j < S; j++) W[2 * S + j] = W[S + j];
    }
    if (smoothed_out) std::memcpy(smoothed_out, W + S, sizeof(double) * S);

    // deconvolution: response vector
    area = 0; lh_gold = -1; posit = 0; maximum = 0;
    for (i = 0; i < S; i++) {
        lda = (double)i - 3 * sigma;
        lda = lda * lda / (2 * sigma * sigma);
        j = (int)(1000 * std::exp(-lda));  // integer-valued response; libm exp, input-independent
        lda = j;
        if (lda != 0) lh_gold = i + 1;
        W[i] = lda;
        area = area + lda;
        if (lda > maximum) { maximum = lda; posit = i; }
    }
    for (i = 0; i < S; i++) W[2 * S + i] = std::fabs(W[S + i]);
    // matrix At*A
    i = lh_gold - 1;
    if (i > S) i = S;
    imin = -i; imax = i;
    for (i = imin; i <= imax; i++) {
        lda = 0;
        jmin = 0;
        if (i < 0) jmin = -i;
        jmax = lh_gold - 1 - i;
        if (jmax > (lh_gold - 1)) jmax = lh_gold - 1;
        for (j = jmin; j <= jmax; j++) {
            ldb = W[j]; ldc = W[i + j];
            lda = lda + ldb * ldc;
        }
        W[S + i - imin] = lda;
    }
    // vector p = At*y
    i = lh_gold - 1;
    imin = -i; imax = S + i - 1;
    for (i = imin; i <= imax; i++) {
        lda = 0;
        for (j = 0; j <= (lh_gold - 1); j++) {
            ldb = W[j];
            k = i + j;
            if (k >= 0 && k < S) {
                ldc = W[2 * S + k];
                lda = lda + ldb * ldc;
            }
        }
        W[4 * S + i - imin] = lda;
    }
    for (i = imin; i <= imax; i++) W[2 * S + i - imin] = W[4 * S + i - imin];  // spills into W3[0..]
    for (i = 0; i < S; i++) W[i] = 1;
    // Gold iterations
    for (lindex = 0; lindex < decon_iterations; lindex++) {
        for (i = 0; i < S; i++) {
            if (std::fabs(W[2 * S + i]) > 0.00001 && std::fabs(W[i]) > 0.00001) {
                lda = 0;
                jmin = lh_gold - 1;
                if (jmin > i) jmin = i;
                jmin = -jmin;
                jmax = lh_gold - 1;
                if (jmax > (S - 1 - i)) jmax = S - 1 - i;
                for (j = jmin; j <= jmax; j++) {
                    ldb = W[j + lh_gold - 1 + S];
                    ldc = W[i + j];
                    lda = lda + ldb * ldc;
                }
                ldb = W[2 * S + i];
                if (lda != 0) lda = ldb / lda;
                else lda = 0;
                ldb = W[i];
                lda = lda * ldb;
                W[3 * S + i] = lda;
            }
        }
        for (i = 0; i < S; i++) W[i] = W[3 * S + i];
    }
    // shift resulting spectrum
    for (i = 0; i < S; i++) {
        lda = W[i];
        j = i + posit;
        j = j % S;
        W[S + j] = lda;
    }
    // write back
    maximum = 0; maximum_decon = 0;
    j = lh_gold - 1;
    for (i = 0; i < S - j; i++) {
        if (i >= shift && i < ssize + shift) {
            W[i] = area * W[S + i + j];
            if (maximum_decon < W[i]) maximum_decon = W[i];
            if (maximum < W[6 * S + i]) maximum = W[6 * S + i];
        } else {
            W[i] = 0;
        }
    }
    lda = 1;
    if (lda > threshold) lda = threshold;
    lda = lda / 100;
    // peak search in the deconvolved spectrum
    for (i = 1; i < S - 1; i++) {
        if (W[i] > W[i - 1] && W[i] > W[i + 1]) {
            if (i >= shift && i < ssize + shift) {
                if (W[i] > lda * maximum_decon && W[6 * S + i] > threshold * maximum / 100.0) {
                    for (j = i - 1, a = 0, b = 0; j <= i + 1; j++) {
                        a += (double)(j - shift) * W[j];
                        b += W[j];
                    }
                    a = a / b;
                    if (a < 0) a = 0;
                    if (a >= ssize) a = ssize - 1;
                    if (peak_index == 0) {
                        pos_x[0] = a;
                        peak_index = 1;
                    } else {
                        for (j = 0, priz = 0; j < peak_index && priz == 0; j++) {
                            if (W[6 * S + shift + (int)a] > W[6 * S + shift + (int)pos_x[j]]) priz = 1;
                        }
                        if (priz == 0) {
                            if (j < max_peaks) pos_x[j] = a;
                        } else {
                            for (k = peak_index; k >= j; k--) {
                                if (k < max_peaks) pos_x[k] = pos_x[k - 1];
                            }
                            pos_x[j - 1] = a;
                        }
                        if (peak_index < max_peaks) peak_index += 1;
                    }
                }
            }
        }
    }
    if (decon_out)
        for (i = 0; i < ssize; i++) decon_out[i] = W[i + shift];
    return peak_index;
}

// TSpectrum::Search(hin, sigma, "nobackground,nodraw", threshold) on a TH1F with `nbins`
// bins on [0,nbins): source[i] = float bin content; PositionX = bin centre of bin
// 1+Int_t(a+0.5); PositionY = (float) bin content there.   (TSpectrum.cxx Search(), 1-D branch)
extern "C" int oracle_tspectrum_search(const float *hist, int nbins, double sigma, double threshold_frac,
                                       int max_peaks, double *position_x, double *position_y, int use_libm_exp)
{
    if (threshold_frac <= 0 || threshold_frac >= 1) threshold_frac = 0.05;
    std::vector<double> source(nbins);
    for (int i = 0; i < nbins; i++) source[i] = hist[i];
    if (sigma < 1) {
        sigma = nbins / max_peaks;
        if (sigma < 1) sigma = 1;
        if (sigma > 8) sigma = 8;
    }
    const int fgIterations = 3, fgAverageWindow = 3;  // TSpectrum static defaults
    std::vector<double> px(max_peaks > 0 ? max_peaks : 1, 0.0);
    int npeaks = oracle_search_highres(source.data(), nbins, sigma, 100 * threshold_frac, fgIterations,
                                       fgAverageWindow, max_peaks, px.data(), nullptr, nullptr, use_libm_exp);
    for (int i = 0; i < npeaks; i++) {
        int bin = 1 + (int)(px[i] + 0.5);
        position_x[i] = (double)(bin - 1) + 0.5;  // GetBinCenter(bin), unit-width bins from 0
        position_y[i] = hist[bin - 1];            // GetBinContent(bin), Float_t storage
    }
    return npeaks;
}
