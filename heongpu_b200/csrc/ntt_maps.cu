// Batched NTT over scattered / strided polynomial sets (GPU_NTT_Poly_Ordered_Inplace and the strided operator calls).
#include "ntt_impl.cuh"

namespace heon {

void launch_ntt_scattered(const Context& c, u64* base, const long long* d_offsets, int n_polys,
                          int prime, bool inverse, long long extent_words, bool aligned, cudaStream_t st)
{
    MapScatter m{base, d_offsets, prime};
    // offsets that are not multiples of 16 words cannot be addressed in 128-byte lines
    Extent e{aligned ? base : base + 1, extent_words, aligned ? base : base + 1, extent_words};
    // arbitrary offsets: the column tiles need 2 KiB-aligned polynomials, not guaranteed here
    run_ntt(c, m, n_polys, inverse, e, st);
}

void launch_ntt_strided(const Context& c, u64* base, long long bstride, int per_batch, int first,
                        long long batch, const PrimeList& pl, bool inverse, cudaStream_t st)
{
    MapStrided m{base, bstride, per_batch, first, pl, c.logn};
    const long long w = (batch - 1) * bstride + ((long long) (first + per_batch) << c.logn);
    const u64* b0 = (bstride & 15) ? base + 1 : base; // odd strides: no line addressing -> LSU path
    Extent e{b0, w, b0, w};
    if ((bstride & 255) == 0)
    {
        e.col_in_base = base;
        e.col_in_words = w;
    }
    run_ntt(c, m, batch * per_batch, inverse, e, st);
}

void launch_ntt_strided_copy(const Context& c, const u64* src, long long src_bstride, u64* dst,
                             int per_batch, long long batch, const PrimeList& pl, bool inverse,
                             cudaStream_t st)
{
    MapStridedCopy m{src, dst, src_bstride, per_batch, pl, c.logn};
    const long long wi = (batch - 1) * src_bstride + ((long long) per_batch << c.logn);
    const long long wo = (batch * per_batch) << c.logn;
    const u64* s0 = (src_bstride & 15) ? src + 1 : src;
    Extent e{s0, wi, dst, wo};
    if ((src_bstride & 255) == 0)
    {
        e.col_in_base = src;
        e.col_in_words = wi;
    }
    run_ntt(c, m, batch * per_batch, inverse, e, st);
}

} // namespace heon
