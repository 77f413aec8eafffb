// Lossless binary64 -> int16 transport packing, see host_pack.hpp.  Plain host code (no CUDA).
#include "host_pack.hpp"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define NPSWF_PACK_AVX2 1
#endif

namespace npswf {

namespace {

const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: (v + MAGIC) - MAGIC = v rounded to an integer, |v| < 2^51

bool pack_portable(const double *__restrict__ x, int16_t *__restrict__ out, size_t n, double lsb, double inv_lsb)
{
    uint64_t bad = 0;
    for (size_t i = 0; i < n; i++) {
        const double xi = x[i];
        const volatile double shifted = xi * inv_lsb + MAGIC;   // volatile: the rounding step must not be folded away
        const double k = shifted - MAGIC;
        const double back = k * lsb;                             // what widen_counts_kernel computes on the device
        const bool in_range = (k >= -32767.0) & (k <= 32767.0);  // false for NaN
        bad |= ((back == xi) & in_range) ? 0ull : 1ull;          // equal as real numbers (-0.0 travels as +0.0)
        out[i] = (int16_t)(int32_t)(in_range ? k : 0.0);
    }
    return bad == 0;
}

#ifdef NPSWF_PACK_AVX2
__attribute__((target("avx2"))) bool pack_avx2(const double *__restrict__ x, int16_t *__restrict__ out, size_t n, double lsb,
                                                double inv_lsb)
{
    const __m256d vinv = _mm256_set1_pd(inv_lsb), vlsb = _mm256_set1_pd(lsb), vmagic = _mm256_set1_pd(MAGIC);
    const __m256d vhi = _mm256_set1_pd(32767.0), vlo = _mm256_set1_pd(-32767.0);
    __m256d inr = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));   // all ones while every sample is reproduced and in range
    size_t i = 0;
    // the staging buffer is written once and read by the copy engine only: streaming stores (no read-for-ownership)
    const bool stream = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    for (; i + 8 <= n; i += 8) {
        const __m256d a = _mm256_loadu_pd(x + i), b = _mm256_loadu_pd(x + i + 4);
        const __m256d ka = _mm256_sub_pd(_mm256_add_pd(_mm256_mul_pd(a, vinv), vmagic), vmagic);
        const __m256d kb = _mm256_sub_pd(_mm256_add_pd(_mm256_mul_pd(b, vinv), vmagic), vmagic);
        inr = _mm256_and_pd(inr, _mm256_and_pd(_mm256_cmp_pd(_mm256_mul_pd(ka, vlsb), a, _CMP_EQ_OQ),
                                               _mm256_cmp_pd(_mm256_mul_pd(kb, vlsb), b, _CMP_EQ_OQ)));
        inr = _mm256_and_pd(inr, _mm256_and_pd(_mm256_cmp_pd(ka, vhi, _CMP_LE_OQ), _mm256_cmp_pd(ka, vlo, _CMP_GE_OQ)));
        inr = _mm256_and_pd(inr, _mm256_and_pd(_mm256_cmp_pd(kb, vhi, _CMP_LE_OQ), _mm256_cmp_pd(kb, vlo, _CMP_GE_OQ)));
        // out-of-range lanes convert to INT_MIN and saturate: harmless, the chunk is rejected anyway
        const __m128i ia = _mm256_cvtpd_epi32(ka), ib = _mm256_cvtpd_epi32(kb);
        if (stream) _mm_stream_si128(reinterpret_cast<__m128i *>(out + i), _mm_packs_epi32(ia, ib));
        else _mm_storeu_si128(reinterpret_cast<__m128i *>(out + i), _mm_packs_epi32(ia, ib));
    }
    if (stream) _mm_sfence();
    bool ok = _mm256_movemask_pd(inr) == 0xf;
    if (i < n) ok = pack_portable(x + i, out + i, n - i, lsb, inv_lsb) && ok;
    return ok;
}
#endif

}  // namespace

bool pack_counts_range(const double *x, int16_t *out, size_t n, double lsb, double inv_lsb)
{
#ifdef NPSWF_PACK_AVX2
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return pack_avx2(x, out, n, lsb, inv_lsb);
#endif
    return pack_portable(x, out, n, lsb, inv_lsb);
}

}  // namespace npswf
