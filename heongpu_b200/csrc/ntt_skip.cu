// Forward NTT of the Method-II key-switch buffer without the digits' own limbs (MapDigitSkip).
#include "ntt_impl.cuh"

namespace heon {

void launch_ntt_digit_skip(const Context& c, u64* tmp, int d, const int* I_loc, const int* I_j, int L, int depth,
                           long long batch, cudaStream_t st, bool col_only, unsigned long long dbl_mask)
{
    const int Qpl = L + c.P_size;
    if (d > 64)
        throw std::invalid_argument("too many key-switch digits");
    MapDigitSkip m;
    m.base = tmp;
    m.d = d;
    m.Qpl = Qpl;
    m.L = L;
    m.depth = depth;
    m.logn = c.logn;
    m.dbl_mask = dbl_mask;
    int acc = 0;
    for (int i = 0; i < d; ++i)
    {
        m.prefix[i] = (short) acc;
        m.I_loc[i] = (short) I_loc[i];
        m.I_j[i] = (short) I_j[i];
        acc += Qpl - I_j[i];
    }
    m.prefix[d] = (short) acc;
    m.per_b = acc;
    const long long w = (batch * d * Qpl) << c.logn;
    run_ntt(c, m, batch * acc, false, Extent{tmp, w, tmp, w, tmp, w}, st, col_only);
}

} // namespace heon
