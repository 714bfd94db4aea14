// Round trip of every serializable object of the class layer through save/load, heongpu::serializer
// (zlib) and a file, and an operator run on the reloaded objects (the shape of the reference's
// example/basic/14_ckks_serialization.cpp and 13_bfv_serialization.cpp).
#include <heongpu/heongpu.hpp>
#include <cmath>
#include <cstdio>
#include <sstream>

template <heongpu::Scheme S> static bool same_words(const Data64* a, const Data64* b, size_t n)
{
    std::vector<Data64> ha(n), hb(n);
    cudaMemcpy(ha.data(), a, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), b, n * 8, cudaMemcpyDeviceToHost);
    return ha == hb;
}
#define CHECK(c)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        if (!(c))                                                                                  \
        {                                                                                          \
            std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c);                              \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

int main()
{
    using namespace heongpu;
    constexpr Scheme S = Scheme::CKKS;
    HEContext<S> ctx = GenHEContext<S>(sec_level_type::none);
    ctx->set_poly_modulus_degree(4096);
    ctx->set_coeff_modulus_bit_sizes({40, 30, 30, 30}, {40, 40});
    ctx->generate();
    HEKeyGenerator<S> kg(ctx);
    kg.set_seed(11);
    Secretkey<S> sk(ctx);
    kg.generate_secret_key(sk);
    Publickey<S> pk(ctx);
    kg.generate_public_key(pk, sk);
    Relinkey<S> rk(ctx);
    kg.generate_relin_key(rk, sk);
    std::vector<int> shifts = {1, 3};
    Galoiskey<S> gk(ctx, shifts);
    kg.generate_galois_key(gk, sk);
    const size_t kw = (size_t) ctx->digit_count(0) * 2 * ctx->Q_prime_size * ctx->n;

    // save / load through a stream
    std::stringstream s1, s2, s3, s4;
    sk.save(s1);
    Secretkey<S> sk2(ctx);
    sk2.load(s1);
    CHECK(same_words<S>(sk.data(), sk2.data(), (size_t) ctx->Q_prime_size * ctx->n));
    // serializer: compressed buffer and file
    auto buf = serializer::serialize(pk);
    Publickey<S> pk2(ctx);
    serializer::deserialize(buf, pk2);
    CHECK(same_words<S>(pk.data(), pk2.data(), (size_t) 2 * ctx->Q_prime_size * ctx->n));
    serializer::save_to_file(rk, "/tmp/heon_rk.bin");
    Relinkey<S> rk2(ctx);
    serializer::load_from_file("/tmp/heon_rk.bin", rk2);
    CHECK(same_words<S>(rk.data(), rk2.data(), kw));
    gk.save(s2);
    Galoiskey<S> gk2(ctx);
    gk2.load(s2);
    CHECK(gk2.device_location_.size() == gk.device_location_.size());
    for (auto& kv : gk.device_location_)
        CHECK(same_words<S>(kv.second.data(), gk2.device_location_.at(kv.first).data(), kw));
    CHECK(same_words<S>(gk.c_data(), gk2.c_data(), kw));

    HEEncoder<S> enc(ctx);
    HEEncryptor<S> cry(ctx, pk2);
    HEDecryptor<S> dec(ctx, sk2);
    HEArithmeticOperator<S> op(ctx, enc);
    std::vector<double> m(ctx->n / 2);
    for (size_t i = 0; i < m.size(); ++i)
        m[i] = std::sin(0.01 * (double) i);
    Plaintext<S> p(ctx);
    enc.encode(p, m, std::pow(2.0, 30));
    p.save(s3);
    Plaintext<S> p2(ctx);
    p2.load(s3);
    Ciphertext<S> c(ctx);
    cry.encrypt(c, p2);
    c.save(s4);
    Ciphertext<S> c2(ctx);
    c2.load(s4);
    CHECK(same_words<S>(c.data(), c2.data(), (size_t) 2 * ctx->Q_size * ctx->n));
    // operators on the reloaded objects: (c2 * c2 -> relinearize -> rescale), rotate by 3
    Ciphertext<S> sq(ctx);
    op.multiply(c2, c2, sq);
    op.relinearize_inplace(sq, rk2);
    op.rescale_inplace(sq);
    Plaintext<S> out(ctx);
    dec.decrypt(out, sq);
    std::vector<double> got;
    enc.decode(got, out);
    for (size_t i = 0; i < m.size(); ++i)
        CHECK(std::fabs(got[i] - m[i] * m[i]) < 1e-3);
    Ciphertext<S> rot(ctx);
    op.rotate_rows(c2, rot, gk2, 3);
    dec.decrypt(out, rot);
    enc.decode(got, out);
    for (size_t i = 0; i + 3 < m.size(); ++i)
        CHECK(std::fabs(got[i] - m[i + 3]) < 1e-3);

    // BFV
    constexpr Scheme B = Scheme::BFV;
    HEContext<B> bctx = GenHEContext<B>(sec_level_type::none);
    bctx->set_poly_modulus_degree(4096);
    bctx->set_coeff_modulus_bit_sizes({36, 36}, {37});
    bctx->set_plain_modulus(1032193);
    bctx->generate();
    HEKeyGenerator<B> bkg(bctx);
    Secretkey<B> bsk(bctx);
    bkg.generate_secret_key(bsk);
    Publickey<B> bpk(bctx);
    bkg.generate_public_key(bpk, bsk);
    HEEncoder<B> benc(bctx);
    HEEncryptor<B> bcry(bctx, bpk);
    HEDecryptor<B> bdec(bctx, bsk);
    std::vector<uint64_t> bm(4096);
    for (size_t i = 0; i < bm.size(); ++i)
        bm[i] = (i * 7919) % 1032193;
    Plaintext<B> bp(bctx);
    benc.encode(bp, bm);
    Ciphertext<B> bc(bctx);
    bcry.encrypt(bc, bp);
    auto bbuf = serializer::serialize(bc);
    Ciphertext<B> bc2(bctx);
    serializer::deserialize(bbuf, bc2);
    Plaintext<B> bout(bctx);
    bdec.decrypt(bout, bc2);
    std::vector<uint64_t> bgot;
    benc.decode(bgot, bout);
    CHECK(bgot == bm);
    std::printf("serialization round trips OK (compressed public key: %zu bytes)\n", buf.size());
    return 0;
}
