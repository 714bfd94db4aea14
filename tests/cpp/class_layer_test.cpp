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

        // in-place forms: (C1 + C2) - C2 == C1, -(-C1) == C1
        {
            Ciphertext<S> T1(context, a);
            operators.add_inplace(T1, C2);
            operators.sub_inplace(T1, C2);
            operators.negate_inplace(T1);
            operators.negate_inplace(T1);
            std::vector<Data64> t1;
            T1.get_data(t1);
            if (t1 != a) { std::puts("FAIL: in-place add/sub/negate"); return 1; }
        }

        // hoisted rotations == stand-alone rotations
        {
            Galoiskey<S> gk2(context, std::vector<int>{1, 2, -1});
            int seed = 20;
            for (auto& kv : gk2.galois_elt)
                gk2.set_key(kv.second, words(context->prime_vector_, context->digit_count(0) * 2, Qp, n, seed++));
            std::vector<int> shifts{2, -1, 1};
            auto hs = operators.rotate_rows_hoisted(C1, gk2, shifts);
            for (size_t r = 0; r < shifts.size(); ++r)
            {
                Ciphertext<S> one;
                operators.rotate_rows(C1, one, gk2, shifts[r]);
                std::vector<Data64> x1, x2;
                hs[r].get_data(x1);
                one.get_data(x2);
                if (x1 != x2) { std::puts("FAIL: hoisted rotation differs from rotate_rows"); return 1; }
            }
        }

        // plaintext operands, keyswitch, conjugate
        auto pw = words(context->prime_vector_, 1, Q, n, 5);
        Plaintext<S> P1(context, pw);
        Ciphertext<S> M, Ad;
        operators.multiply_plain(C1, P1, M);
        operators.add_plain(C1, P1, Ad);
        operators.sub_plain_inplace(Ad, P1); // (C1 + P) - P == C1
        std::vector<Data64> m1, a1;
        M.get_data(m1);
        Ad.get_data(a1);
        DeviceVector<Data64> dp(pw), dm((size_t) 2 * Q * n);
        heon_ckks_multiply_plain(context->handle(), da.data(), 0, dp.data(), 0, dm.data(), 0, 2, 0, 1, nullptr);
        std::vector<Data64> m2(m1.size());
        cudaMemcpy(m2.data(), dm.data(), m2.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (m1 != m2 || a1 != a) { std::puts("FAIL: plaintext operators"); return 1; }
        Switchkey<S> swk(context);
        swk.set_data(words(context->prime_vector_, context->digit_count(0) * 2, Qp, n, 6));
        Ciphertext<S> K1, K2;
        operators.keyswitch(C1, K1, swk);
        galois_key.set_conjugate_key(words(context->prime_vector_, context->digit_count(0) * 2, Qp, n, 7));
        operators.conjugate(C1, K2, galois_key);
        std::vector<Data64> k1, k2, k3(2 * (size_t) Q * n);
        K1.get_data(k1);
        K2.get_data(k2);
        heon_ckks_keyswitch(context->handle(), da.data(), 0, dr.data(), 0, swk.data(), 0, 1, nullptr);
        cudaMemcpy(k3.data(), dr.data(), k3.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (k1 != k3) { std::puts("FAIL: keyswitch differs from the C ABI"); return 1; }
        heon_ckks_apply_galois(context->handle(), da.data(), 0, dr.data(), 0, galois_key.c_data(), (uint32_t) (2 * n - 1), 0, 1, nullptr);
        cudaMemcpy(k3.data(), dr.data(), k3.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (k2 != k3) { std::puts("FAIL: conjugate differs from apply_galois(2N-1)"); return 1; }
    }
    // BFV twins (test_bfv_multiplication.cpp parameters: N = 4096, {36,36}/{37}, t = 1032193)
    {
        constexpr Scheme B = Scheme::BFV;
        HEContext<B> context = GenHEContext<B>(sec_level_type::none);
        context->set_poly_modulus_degree(4096);
        context->set_coeff_modulus_bit_sizes({36, 36}, {37});
        context->set_plain_modulus(1032193);
        context->generate();
        const int n = context->n, Q = context->Q_size, Qp = context->Q_prime_size;
        HEArithmeticOperator<B> operators(context);
        auto a = words(context->prime_vector_, 2, Q, n, 11), b = words(context->prime_vector_, 2, Q, n, 12);
        Ciphertext<B> C1(context, a), C2(context, b), C3;
        Relinkey<B> relin_key(context);
        relin_key.set_data(words(context->prime_vector_, context->digit_count() * 2, Qp, n, 13));
        operators.multiply(C1, C2, C3);
        operators.relinearize_inplace(C3, relin_key);
        std::vector<Data64> got;
        C3.get_data(got);
        DeviceVector<Data64> da(a), db(b), dc((size_t) 3 * Q * n);
        heon_bfv_multiply(context->handle(), da.data(), 0, db.data(), 0, dc.data(), 0, 1, nullptr);
        heon_bfv_relinearize(context->handle(), dc.data(), 0, relin_key.data(), 1, nullptr);
        std::vector<Data64> want(got.size());
        cudaMemcpy(want.data(), dc.data(), want.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (got != want || C3.size() != 2 || C3.relinearization_required()) { std::puts("FAIL: BFV multiply+relinearize"); return 1; }
        Galoiskey<B> gk(context, std::vector<int>{1});
        gk.set_key(gk.galois_elt[1], words(context->prime_vector_, context->digit_count() * 2, Qp, n, 14));
        gk.set_key(gk.galois_elt_zero, words(context->prime_vector_, context->digit_count() * 2, Qp, n, 15));
        Ciphertext<B> R1, R2, Sm;
        operators.rotate_rows(C1, R1, gk, 1);
        operators.rotate_columns(C1, R2, gk);
        operators.add(R1, R2, Sm);
        operators.sub(Sm, R2, Sm);
        std::vector<Data64> r1, s1;
        R1.get_data(r1);
        Sm.get_data(s1);
        DeviceVector<Data64> dr((size_t) 2 * Q * n);
        heon_bfv_apply_galois(context->handle(), da.data(), 0, dr.data(), 0, gk.device_location_[gk.galois_elt[1]].data(),
                              (uint32_t) gk.galois_elt[1], 1, nullptr);
        std::vector<Data64> r2(r1.size());
        cudaMemcpy(r2.data(), dr.data(), r2.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (r1 != r2 || s1 != r1) { std::puts("FAIL: BFV rotate / add / sub"); return 1; }
        // plaintext operands and keyswitch
        std::vector<Data64> pw(n);
        for (int i = 0; i < n; ++i) pw[i] = (Data64) ((i * 7919u + 13u) % 1032193u);
        Plaintext<B> P1(context, pw);
        Ciphertext<B> Ap, Mp, Kp;
        operators.add_plain(C1, P1, Ap);
        operators.sub_plain_inplace(Ap, P1); // (C1 + P) - P == C1
        operators.multiply_plain(C1, P1, Mp);
        std::vector<Data64> a1, m1;
        Ap.get_data(a1);
        Mp.get_data(m1);
        DeviceVector<Data64> dp(pw), dm((size_t) 2 * Q * n);
        heon_bfv_multiply_plain(context->handle(), da.data(), 0, dp.data(), 0, dm.data(), 0, 1, nullptr);
        std::vector<Data64> m2(m1.size());
        cudaMemcpy(m2.data(), dm.data(), m2.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (a1 != a || m1 != m2) { std::puts("FAIL: BFV plaintext operators"); return 1; }
        Switchkey<B> swk(context);
        swk.set_data(words(context->prime_vector_, context->digit_count() * 2, Qp, n, 16));
        operators.keyswitch(C1, Kp, swk);
        std::vector<Data64> k1, k2(2 * (size_t) Q * n);
        Kp.get_data(k1);
        heon_bfv_keyswitch(context->handle(), da.data(), 0, dr.data(), 0, swk.data(), 1, nullptr);
        cudaMemcpy(k2.data(), dr.data(), k2.size() * sizeof(Data64), cudaMemcpyDeviceToHost);
        if (k1 != k2) { std::puts("FAIL: BFV keyswitch"); return 1; }
    }
    std::puts("OK");
    return 0;
}
