// BSGS diagonal matrix-vector products through the class layer with real keys: the single-hoisting form
// (HEOperator<CKKS>::multiply_matrix, ckks/operator.cu:2803-2895), its less-memory variant (:3398-3496) and the
// double-hoisting form in PQ_l (multiply_matrix_v2, :2898-3390), each against the plain computation
//   y = sum_j rot_{G_j}( sum_k diag_jk * rot_{b_k}(x) ).
#include <heongpu/heongpu.hpp>
#include <cmath>
#include <cstdio>
#include <random>

#define CHECK(c)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        if (!(c))                                                                                  \
        {                                                                                          \
            std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c);                              \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

using namespace heongpu;
constexpr Scheme S = Scheme::CKKS;

static std::vector<double> rot(const std::vector<double>& v, int s)
{
    const int n = (int) v.size();
    std::vector<double> r(n);
    for (int i = 0; i < n; ++i)
        r[i] = v[((i + s) % n + n) % n];
    return r;
}

int main()
{
    HEContext<S> ctx = GenHEContext<S>(sec_level_type::none);
    ctx->set_poly_modulus_degree(8192);
    ctx->set_coeff_modulus_bit_sizes({50, 40, 40, 40, 40}, {50, 50, 50});
    ctx->generate();
    // the same primes as one long Q chain: its encoder produces plaintexts over PQ_0
    std::vector<Data64> all;
    for (auto& m : ctx->prime_vector_)
        all.push_back(m.value);
    HEContext<S> spare = GenHEContext<S>(sec_level_type::none);
    spare->set_poly_modulus_degree(8192);
    spare->set_coeff_modulus_bit_sizes({45}, {46});
    spare->generate();
    HEContext<S> wide = GenHEContext<S>(sec_level_type::none);
    wide->set_poly_modulus_degree(8192);
    wide->set_coeff_modulus_values(all, {spare->prime_vector_[1].value});
    wide->generate();

    HEKeyGenerator<S> kg(ctx);
    kg.set_seed(21);
    Secretkey<S> sk(ctx);
    kg.generate_secret_key(sk);
    Publickey<S> pk(ctx);
    kg.generate_public_key(pk, sk);
    std::vector<int> shifts = {1, 2, 3, 4, 8};
    Galoiskey<S> gk(ctx, shifts);
    kg.generate_galois_key(gk, sk);
    HEEncoder<S> enc(ctx), wenc(wide);
    HEEncryptor<S> cry(ctx, pk);
    HEDecryptor<S> dec(ctx, sk);
    HEArithmeticOperator<S> op(ctx, enc);

    const int slots = ctx->n / 2, n = ctx->n, L = ctx->Q_size, K = ctx->P_size;
    const double scale = std::pow(2.0, 40);
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> x(slots);
    for (auto& v : x)
        v = U(rng);
    Plaintext<S> px(ctx);
    enc.encode(px, x, scale);
    Ciphertext<S> cx(ctx);
    cry.encrypt(cx, px);

    // baby steps {0,1,2,3}; giant steps {0,4,8}; full groups for the single-hoisting form
    std::vector<std::vector<std::vector<int>>> diags = {{{0, 1, 2, 3}, {4, 5, 6, 7}, {8, 9, 10, 11}}};
    std::vector<std::vector<int>> rot_n1 = {{0, 4, 8}}, rot_n2 = {{3, 1, 0, 2}};
    std::vector<std::vector<std::vector<int>>> chains = {{{0}, {4}, {8}}};
    std::vector<double> want(slots, 0.0);
    size_t terms = 12;
    std::vector<DeviceVector<Data64>> mat_q(1), mat_pq(1);
    mat_q[0] = DeviceVector<Data64>(terms * (size_t) L * n);
    mat_pq[0] = DeviceVector<Data64>(terms * (size_t) (L + K) * n);
    size_t t = 0;
    for (size_t j = 0; j < diags[0].size(); ++j)
    {
        std::vector<double> inner(slots, 0.0);
        for (int dg : diags[0][j])
        {
            std::vector<double> p(slots);
            for (auto& v : p)
                v = U(rng);
            Plaintext<S> pq(ctx), ppq(wide);
            enc.encode(pq, p, scale);
            wenc.encode(ppq, p, scale);
            cudaMemcpy(mat_q[0].data() + t * (size_t) L * n, pq.data(), (size_t) L * n * 8, cudaMemcpyDeviceToDevice);
            cudaMemcpy(mat_pq[0].data() + t * (size_t) (L + K) * n, ppq.data(), (size_t) (L + K) * n * 8,
                       cudaMemcpyDeviceToDevice);
            ++t;
            const std::vector<double> r = rot(x, dg - rot_n1[0][j]);
            for (int i = 0; i < slots; ++i)
                inner[i] += p[i] * r[i];
        }
        const std::vector<double> r = rot(inner, rot_n1[0][j]);
        for (int i = 0; i < slots; ++i)
            want[i] += r[i];
    }
    auto max_err = [&](Ciphertext<S>& c) {
        Plaintext<S> p(ctx);
        dec.decrypt(p, c);
        std::vector<double> got;
        enc.decode(got, p);
        double e = 0;
        for (int i = 0; i < slots; ++i)
            e = std::max(e, std::fabs(got[i] - want[i]));
        return e;
    };
    op.set_matrix_scale(scale);
    // single hoisting: baby shifts = diags[m][0], giant shift of group j = diags[m][j][0]
    Ciphertext<S> y1 = op.multiply_matrix(cx, mat_q, diags, gk);
    CHECK(y1.depth() == 1);
    const double e1 = max_err(y1);
    std::printf("single hoisting   max err %.3e\n", e1);
    CHECK(e1 < 1e-4);
    Ciphertext<S> y2 = op.multiply_matrix_less_memory(cx, mat_q, diags, chains, gk);
    const double e2 = max_err(y2);
    std::printf("less memory       max err %.3e\n", e2);
    CHECK(e2 < 1e-4);
    // double hoisting in PQ_l
    Ciphertext<S> y3 = op.multiply_matrix_v2(cx, mat_pq, diags, rot_n1, rot_n2, gk);
    CHECK(y3.depth() == 1);
    y3.scale_ = y1.scale_; // the reference multiplies by prime_vector_[L] there (:3384); decode at the true scale
    const double e3 = max_err(y3);
    std::printf("double hoisting   max err %.3e\n", e3);
    CHECK(e3 < 1e-4);
    // a shift without a key
    std::vector<std::vector<int>> bad_n1 = {{0, 4, 16}};
    std::vector<std::vector<std::vector<int>>> bad = {{{0, 1, 2, 3}, {4, 5, 6, 7}, {16, 17, 18, 19}}};
    bool threw = false;
    try
    {
        op.multiply_matrix_v2(cx, mat_pq, bad, bad_n1, rot_n2, gk);
    }
    catch (const std::logic_error&)
    {
        threw = true;
    }
    CHECK(threw);
    std::printf("BSGS OK\n");
    return 0;
}
