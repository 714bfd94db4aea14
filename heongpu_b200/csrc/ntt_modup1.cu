// Method-I mod-up fused into the forward NTT (MapModUpI).
#include "ntt_impl.cuh"

namespace heon {

void launch_modup1_ntt(const Context& c, const u64* coef, long long coef_bstride, u64* out, int L,
                       int depth, long long batch, cudaStream_t st, bool col_only)
{
    const int Qpl = L + c.P_size;
    MapModUpI m{coef, out, coef_bstride, L, Qpl, depth, c.logn, c.d_pc};
    const long long wo = (batch * L * Qpl) << c.logn;
    Extent e{out, wo, out, wo};
    if ((coef_bstride & 255) == 0)
    {
        e.col_in_base = coef;
        e.col_in_words = (batch - 1) * coef_bstride + ((long long) L << c.logn);
    }
    run_ntt(c, m, batch * L * Qpl, false, e, st, col_only);
}

} // namespace heon
