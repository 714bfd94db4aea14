// Divide-and-round stage one fused into the forward NTT (MapDivRoundOne).
#include "ntt_impl.cuh"

namespace heon {

void launch_divround1_ntt(const Context& c, const u64* src, long long bstride, long long cstride,
                          u64* out, int Lout, u64 half, u64 plast, const u64* d_half_mod,
                          long long batch, cudaStream_t st, bool col_only)
{
    MapDivRoundOne m{src, out, bstride, cstride, Lout, c.logn, half, plast, d_half_mod};
    const long long wo = (batch * 2 * Lout) << c.logn;
    Extent e{out, wo, out, wo};
    if ((bstride & 255) == 0 && (cstride & 255) == 0)
    {
        e.col_in_base = src;
        e.col_in_words = (batch - 1) * bstride + cstride + (1ll << c.logn);
    }
    run_ntt(c, m, batch * 2 * Lout, false, e, st, col_only);
}

} // namespace heon
