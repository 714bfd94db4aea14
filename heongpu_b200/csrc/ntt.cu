// Batched negacyclic NTT / INTT: contiguous polynomials (gpuntt::GPU_NTT / GPU_INTT / *_Modulus_Ordered).
#include "ntt_impl.cuh"

namespace heon {

void launch_ntt(const Context& c, const u64* src, u64* dst, long long n_polys, const PrimeList& pl,
                bool inverse, cudaStream_t st, bool col_only)
{
    MapContig m{src, dst, pl, c.logn};
    const long long w = n_polys << c.logn;
    run_ntt(c, m, n_polys, inverse, Extent{src, w, dst, w, src, w}, st, col_only);
}

} // namespace heon
