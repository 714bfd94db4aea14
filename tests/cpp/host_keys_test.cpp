// Evaluation keys in host memory (ExecutionOptions::set_storage_type(HOST), store_in_host / store_in_device,
// src/include/heongpu/host/{ckks,bfv}/evaluationkey.cuh): every operator that takes a key must give the same words
// whether the key lives on the device or on the host (the reference stages the key per call,
// ckks/operator.cu:3117-3131).
#include <heongpu/heongpu.hpp>
#include <cmath>
#include <cstdio>
#include <sstream>

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

template <class Ct> static std::vector<Data64> words_of(Ct& c)
{
    std::vector<Data64> h(c.memory_size());
    cudaMemcpy(h.data(), c.data(), h.size() * 8, cudaMemcpyDeviceToHost);
    return h;
}

static int ckks()
{
    constexpr Scheme S = Scheme::CKKS;
    HEContext<S> ctx = GenHEContext<S>(sec_level_type::none);
    ctx->set_poly_modulus_degree(8192);
    ctx->set_coeff_modulus_bit_sizes({50, 40, 40, 40}, {50, 50});
    ctx->generate();
    HEKeyGenerator<S> kg(ctx);
    kg.set_seed(3);
    Secretkey<S> sk(ctx);
    kg.generate_secret_key(sk);
    Publickey<S> pk(ctx);
    kg.generate_public_key(pk, sk);
    Relinkey<S> rk(ctx);
    kg.generate_relin_key(rk, sk);
    Galoiskey<S> gk(ctx); // default keys: +-2^i, i < 8, and the conjugation key
    kg.generate_galois_key(gk, sk);
    HEEncoder<S> enc(ctx);
    HEEncryptor<S> cry(ctx, pk);
    HEArithmeticOperator<S> op(ctx, enc);
    std::vector<double> m(ctx->n / 2);
    for (size_t i = 0; i < m.size(); ++i)
        m[i] = std::cos(0.02 * (double) i);
    Plaintext<S> p(ctx);
    enc.encode(p, m, std::pow(2.0, 40));
    Ciphertext<S> c1(ctx), c2(ctx);
    cry.encrypt(c1, p);
    cry.encrypt(c2, p);

    auto run = [&](std::vector<std::vector<Data64>>& out) {
        Ciphertext<S> prod(ctx), rot(ctx), rot2(ctx), conj(ctx);
        op.multiply(c1, c2, prod);
        op.relinearize_inplace(prod, rk);
        out.push_back(words_of(prod));
        op.rotate_rows(c1, rot, gk, 4);
        out.push_back(words_of(rot));
        op.rotate_rows(c1, rot2, gk, 3); // a chain of two keys
        out.push_back(words_of(rot2));
        op.conjugate(c1, conj, gk);
        out.push_back(words_of(conj));
        std::vector<int> hs = {1, 2};
        auto h = op.rotate_rows_hoisted(c1, gk, hs);
        out.push_back(words_of(h[0]));
        out.push_back(words_of(h[1]));
    };
    std::vector<std::vector<Data64>> dev, host, back;
    run(dev);
    rk.store_in_host();
    gk.store_in_host();
    CHECK(!rk.is_on_device() && !gk.is_on_device() && rk.data() == nullptr && gk.device_location_.empty());
    run(host);
    CHECK(dev == host);
    // serialization of a host-stored key, and a key generated straight into host memory
    std::stringstream s1;
    gk.save(s1);
    Galoiskey<S> gk2(ctx);
    gk2.load(s1);
    Ciphertext<S> r1(ctx);
    op.rotate_rows(c1, r1, gk2, 4);
    CHECK(words_of(r1) == dev[1]);
    Relinkey<S> rk_h(ctx);
    HEKeyGenerator<S> kg2(ctx);
    kg2.set_seed(3);
    Secretkey<S> sk_b(ctx);
    kg2.generate_secret_key(sk_b);
    Publickey<S> pk_b(ctx);
    kg2.generate_public_key(pk_b, sk_b);
    kg2.generate_relin_key(rk_h, sk_b, ExecutionOptions().set_storage_type(storage_type::HOST));
    CHECK(!rk_h.is_on_device() && rk_h.host_location_.size() == rk.host_location_.size());
    CHECK(std::equal(rk_h.host_location_.data(), rk_h.host_location_.data() + rk_h.host_location_.size(), rk.host_location_.data()));
    rk.store_in_device();
    gk.store_in_device();
    CHECK(rk.is_on_device() && gk.is_on_device() && gk.host_location_.empty());
    run(back);
    CHECK(dev == back);
    return 0;
}

static int bfv()
{
    constexpr Scheme S = Scheme::BFV;
    HEContext<S> ctx = GenHEContext<S>(sec_level_type::none);
    ctx->set_poly_modulus_degree(4096);
    ctx->set_coeff_modulus_bit_sizes({36, 36}, {37});
    ctx->set_plain_modulus(1032193);
    ctx->generate();
    HEKeyGenerator<S> kg(ctx);
    kg.set_seed(4);
    Secretkey<S> sk(ctx);
    kg.generate_secret_key(sk);
    Publickey<S> pk(ctx);
    kg.generate_public_key(pk, sk);
    Relinkey<S> rk(ctx);
    kg.generate_relin_key(rk, sk);
    Galoiskey<S> gk(ctx);
    kg.generate_galois_key(gk, sk);
    HEEncoder<S> enc(ctx);
    HEEncryptor<S> cry(ctx, pk);
    HEArithmeticOperator<S> op(ctx, enc);
    std::vector<uint64_t> m(ctx->n);
    for (size_t i = 0; i < m.size(); ++i)
        m[i] = i % 1000;
    Plaintext<S> p(ctx);
    enc.encode(p, m);
    Ciphertext<S> c1(ctx);
    cry.encrypt(c1, p);
    auto run = [&](std::vector<std::vector<Data64>>& out) {
        Ciphertext<S> prod(ctx), rot(ctx), col(ctx);
        op.multiply(c1, c1, prod);
        op.relinearize_inplace(prod, rk);
        out.push_back(words_of(prod));
        op.rotate_rows(c1, rot, gk, 4);
        out.push_back(words_of(rot));
        op.rotate_columns(c1, col, gk);
        out.push_back(words_of(col));
    };
    std::vector<std::vector<Data64>> dev, host;
    run(dev);
    rk.store_in_host();
    gk.store_in_host();
    run(host);
    CHECK(dev == host);
    return 0;
}

int main()
{
    if (ckks() || bfv())
        return 1;
    std::printf("host keys OK\n");
    return 0;
}
