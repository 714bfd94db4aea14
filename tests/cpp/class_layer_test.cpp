// Exercises the heongpu:: class layer (heongpu_b200/include/heongpu/heongpu.hpp)
// the way the reference's tests use it (test/test_ckks_relinearization.cpp):
// context -> ciphertexts -> multiply -> relinearize_inplace -> rescale_inplace
// -> rotate_rows, compared word for word with direct C-ABI calls, plus the
// reference's state-flag exceptions.  Built by tests/test_class_layer.py.
#include <heongpu/heongpu.hpp>
#include <cstdio>
#include <cstring>

using namespace heongpu;
constexpr Scheme S = Scheme::CKKS;

static std::vector<Data64> words(const std::vector<Modulus64>& primes, int polys_per_prime_set, int limbs, int n, Data64 seed)
{
    std::vector<Data64> v((size_t) polys_per_prime_set * limbs * n);
    Data64 s = seed;
    for (int c = 0; c < polys_per_prime_set; ++c)
        for (int y = 0; y < limbs; ++y)
            for (int i = 0; i < n; ++i)
            {
                s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                v[((size_t) c * limbs + y) * n + i] = (s >> 3) % primes[y].value;
            }
    return v;
}

int main(int argc, char** argv)
{
    const bool compile_only = argc > 1 && !strcmp(argv[1], "--no-gpu");
    if (compile_only)
    {
        std::puts("OK (compiled)");
        return 0;
    }
    for (int method = 1; method <= 2; ++method)
    {
        HEContext<S> context = GenHEContext<S>(sec_level_type::none);
        context->set_poly_modulus_degree(4096);
        if (method == 1)
            context->set_coeff_modulus_bit_sizes({40, 30, 30}, {40});
        else
            context->set_coeff_modulus_bit_sizes({40, 30, 30, 30}, {40, 40});
        context->generate();
        const int n = context->n, Q = context->Q_size, Qp = context->Q_prime_size;
        HEArithmeticOperator<S> operators(context);

        auto a = words(context->prime_vector_, 2, Q, n, 1), b = words(context->prime_vector_, 2, Q, n, 2);
        Ciphertext<S> C1(context, a), C2(context, b), C3;
        Relinkey<S> relin_key(context);
        relin_key.set_data(words(context->prime_vector_, context->digit_count(0) * 2, Qp, n, 3));

        operators.multiply(C1, C2, C3);
        bool threw = false;
        try { operators.multiply(C3, C1, C2); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) { std::puts("FAIL: multiply of a size-3 ciphertext must throw"); return 1; }
        operators.relinearize_inplace(C3, relin_key);
        operators.rescale_inplace(C3);
        std::vector<Data64> got;
        C3.get_data(got);

        // same through the C ABI
        DeviceVector<Data64> da(a), db(b), dc((size_t) 3 * Q * n);
        heon_ckks_multiply(context->handle(), da.data(), 0, db.data(), 0, dc.data(), 0, 0, 1, nullptr);
        heon_ckks_relinearize(context->handle(), dc.data(), 0, relin_key.data(), 0, 1, nullptr);
        heon_ckks_rescale(context->handle(), dc.data(), 0, 0, 1, nullptr);
        std::vector<Data64> want(got.size());
        cudaMemcpy(want.data(), dc.data(), want.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (got != want) { std::printf("FAIL: class layer differs from the C ABI (method %d)\n", method); return 1; }
        if (C3.depth() != 1 || C3.size() != 2 || C3.rescale_required() || C3.relinearization_required())
        { std::puts("FAIL: ciphertext metadata"); return 1; }

        // rotation
        Galoiskey<S> galois_key(context, std::vector<int>{1});
        const int elt = galois_key.galois_elt[1];
        galois_key.set_key(elt, words(context->prime_vector_, context->digit_count(0) * 2, Qp, n, 4));
        Ciphertext<S> R;
        operators.rotate_rows(C1, R, galois_key, 1);
        std::vector<Data64> r1;
        R.get_data(r1);
        DeviceVector<Data64> dr((size_t) 2 * Q * n);
        heon_ckks_apply_galois(context->handle(), da.data(), 0, dr.data(), 0, galois_key.device_location_[elt].data(),
                               (uint32_t) elt, 0, 1, nullptr);
        std::vector<Data64> r2(r1.size());
        cudaMemcpy(r2.data(), dr.data(), r2.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (r1 != r2) { std::puts("FAIL: rotate_rows differs from the C ABI"); return 1; }
        threw = false;
        try { operators.rotate_rows(C1, R, galois_key, 5); } catch (const std::logic_error&) { threw = true; }
        if (!threw) { std::puts("FAIL: missing galois key must throw"); return 1; }
    }
    std::puts("OK");
    return 0;
}
