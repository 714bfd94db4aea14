// heongpu_client.hpp -- the client side of the class layer: Secretkey, Publickey, HEKeyGenerator,
// HEEncoder, HEEncryptor, HEDecryptor, with the reference's names, argument order and exception
// behaviour, on top of the heon_keygen_* / heon_encrypt / heon_*_decrypt / heon_*_encode entry points
// (include/heon_b200.h, "client side").  Included by heongpu.hpp.
//
// reference: src/include/heongpu/host/{ckks,bfv}/{secretkey,publickey,keygenerator,encoder,encryptor,
// decryptor}.cuh.  With these the reference's own test/*.cpp and benchmark/*.cpp sources compile against
// this header unchanged (tests/cpp/build_reference_tests.sh).
#pragma once
#include <complex>
#include <random>

typedef std::complex<double> Complex64;

namespace heongpu {

namespace detail {
inline int digits0(const HEContextImpl<Scheme::CKKS>& c) { return c.digit_count(0); }
inline int digits0(const HEContextImpl<Scheme::BFV>& c) { return c.digit_count(); }
template <Scheme S> size_t evk_words(const HEContext<S>& c)
{
    return (size_t) digits0(*c) * 2 * c->Q_prime_size * c->n;
}
} // namespace detail

// ---- Secretkey<S>: [Q'][N] NTT-domain words (secretkey.cu) -------------------------------------------
template <Scheme S> class Secretkey : public detail::Storable {
  public:
    Secretkey() = default; // filled by load() (serializer::deserialize / load_from_file)
    explicit Secretkey(HEContext<S> ctx) : context_(ctx), hamming_weight_(ctx->n >> 1)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_prime_size;
    }
    Secretkey(HEContext<S> ctx, int hamming_weight) : context_(ctx), hamming_weight_(hamming_weight)
    {
        if (hamming_weight <= 0 || hamming_weight > ctx->n)
            throw std::invalid_argument("hamming weight has to be in range 0 to ring size.");
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_prime_size;
    }
    int coeff_modulus_count() const { return coeff_modulus_count_; }
    void save(std::ostream& os) const;
    void load(std::istream& is);
    HEContext<S> context_;
    int ring_size_ = 0, coeff_modulus_count_ = 0;
    int hamming_weight_ = 0;
    bool in_ntt_domain_ = false, secret_key_generated_ = false;
};

// ---- Publickey<S>: [2][Q'][N] NTT-domain words (publickey.cu) ----------------------------------------
template <Scheme S> class Publickey : public detail::Storable {
  public:
    Publickey() = default;
    explicit Publickey(HEContext<S> ctx) : context_(ctx)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_prime_size;
    }
    void save(std::ostream& os) const;
    void load(std::istream& is);
    HEContext<S> context_;
    int ring_size_ = 0, coeff_modulus_count_ = 0;
    bool in_ntt_domain_ = false, public_key_generated_ = false;
};

// ---- HEKeyGenerator<S> (keygenerator.cu) -------------------------------------------------------------
// Keys are deterministic in the generator's seed (set_seed); the default seed comes from
// std::random_device like the reference's (keygenerator.cu:16-22).
template <Scheme S> class HEKeyGenerator {
  public:
    explicit HEKeyGenerator(HEContext<S> ctx) : context_(ctx)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        std::random_device rd;
        seed_ = ((uint64_t) rd() << 32) ^ rd();
    }
    void set_seed(uint64_t seed) { seed_ = seed, counter_ = 0; }

    void generate_secret_key(Secretkey<S>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (sk.secret_key_generated_)
            throw std::logic_error("Secretkey is already generated!");
        DeviceVector<Data64> mem((size_t) context_->Q_prime_size * context_->n, opt.stream_);
        detail::check(heon_keygen_secret(context_->handle(), next(), sk.hamming_weight_, mem.data(), opt.stream_));
        sk.memory_set(std::move(mem));
        sk.in_ntt_domain_ = true;
        sk.secret_key_generated_ = true;
        detail::output_storage(sk, opt);
    }
    void generate_public_key(Publickey<S>& pk, Secretkey<S>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!sk.secret_key_generated_)
            throw std::logic_error("Secretkey is not generated!");
        if (pk.public_key_generated_)
            throw std::logic_error("Publickey is already generated!");
        detail::InputGuard<Secretkey<S>> g(sk, opt, false);
        DeviceVector<Data64> mem((size_t) 2 * context_->Q_prime_size * context_->n, opt.stream_);
        detail::check(heon_keygen_public(context_->handle(), sk.data(), next(), mem.data(), opt.stream_));
        pk.memory_set(std::move(mem));
        pk.in_ntt_domain_ = true;
        pk.public_key_generated_ = true;
        detail::output_storage(pk, opt);
    }
    void generate_relin_key(Relinkey<S>& rk, Secretkey<S>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!sk.secret_key_generated_)
            throw std::logic_error("Secretkey is not generated!");
        if (rk.relin_key_generated_)
            throw std::logic_error("Relinkey is already generated!");
        detail::InputGuard<Secretkey<S>> g(sk, opt, false);
        DeviceVector<Data64> mem(detail::evk_words(context_), opt.stream_);
        detail::check(heon_keygen_relin(context_->handle(), sk.data(), next(), mem.data(), opt.stream_));
        rk.device_location_ = std::move(mem);
        rk.relin_key_generated_ = true;
        if (opt.storage_ == storage_type::HOST) // output_storage_manager of the reference
            rk.store_in_host(opt.stream_);
    }
    // every shift of the key's table plus the conjugation / column-rotation element 2N-1
    void generate_galois_key(Galoiskey<S>& gk, Secretkey<S>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!sk.secret_key_generated_)
            throw std::logic_error("Secretkey is not generated!");
        if (gk.galois_key_generated_)
            throw std::logic_error("Galoiskey is already generated!");
        detail::InputGuard<Secretkey<S>> g(sk, opt, false);
        for (auto& kv : gk.galois_elt)
        {
            if (kv.second == 0)
                throw std::invalid_argument("Galois Key can not be generated, Step count too large");
            if (gk.device_location_.count(kv.second))
                continue;
            DeviceVector<Data64> mem(detail::evk_words(context_), opt.stream_);
            detail::check(heon_keygen_galois(context_->handle(), sk.data(), (uint32_t) kv.second, next(), mem.data(), opt.stream_));
            gk.device_location_[kv.second] = std::move(mem);
        }
        for (uint32_t e : gk.custom_galois_elt)
        {
            if (gk.device_location_.count((int) e))
                continue;
            DeviceVector<Data64> mem(detail::evk_words(context_), opt.stream_);
            detail::check(heon_keygen_galois(context_->handle(), sk.data(), e, next(), mem.data(), opt.stream_));
            gk.device_location_[(int) e] = std::move(mem);
        }
        const int zero = 2 * context_->n - 1;
        DeviceVector<Data64> mem(detail::evk_words(context_), opt.stream_);
        detail::check(heon_keygen_galois(context_->handle(), sk.data(), (uint32_t) zero, next(), mem.data(), opt.stream_));
        gk.set_zero_key(zero, std::move(mem));
        gk.galois_key_generated_ = true;
        if (opt.storage_ == storage_type::HOST)
            gk.store_in_host(opt.stream_);
    }
    void generate_switch_key(Switchkey<S>& swk, Secretkey<S>& new_sk, Secretkey<S>& old_sk,
                             const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!new_sk.secret_key_generated_ || !old_sk.secret_key_generated_)
            throw std::logic_error("Secretkey is not generated!");
        if (swk.switch_key_generated_)
            throw std::logic_error("Switchkey is already generated!");
        detail::InputGuard<Secretkey<S>> g1(new_sk, opt, false), g2(old_sk, opt, false);
        DeviceVector<Data64> mem(detail::evk_words(context_), opt.stream_);
        detail::check(heon_keygen_switch(context_->handle(), new_sk.data(), old_sk.data(), next(), mem.data(), opt.stream_));
        swk.device_location_ = std::move(mem);
        swk.switch_key_generated_ = true;
        if (opt.storage_ == storage_type::HOST)
            swk.store_in_host(opt.stream_);
    }

  private:
    uint64_t next() { return seed_ * 0x9E3779B97F4A7C15ull + (++counter_); }
    HEContext<S> context_;
    uint64_t seed_ = 0, counter_ = 0;
};

// ---- HEEncoder<CKKS> (ckks/encoder.cu) ---------------------------------------------------------------
template <> class HEEncoder<Scheme::CKKS> {
  public:
    explicit HEEncoder(HEContext<Scheme::CKKS> ctx) : context_(ctx)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        slot_count_ = ctx->n >> 1;
    }
    int slot_count() const { return slot_count_; }

    void encode(Plaintext<Scheme::CKKS>& plain, const std::vector<double>& message, double scale,
                const ExecutionOptions& opt = ExecutionOptions())
    {
        std::vector<double> z(2 * message.size(), 0.0);
        for (size_t i = 0; i < message.size(); ++i)
            z[2 * i] = message[i];
        encode_slots(plain, z.data(), (int) message.size(), scale, opt);
    }
    void encode(Plaintext<Scheme::CKKS>& plain, const std::vector<Complex64>& message, double scale,
                const ExecutionOptions& opt = ExecutionOptions())
    {
        encode_slots(plain, reinterpret_cast<const double*>(message.data()), (int) message.size(), scale, opt);
    }
    // one value in every slot (encode_ckks(double), encoder.cu)
    void encode(Plaintext<Scheme::CKKS>& plain, double message, double scale, const ExecutionOptions& opt = ExecutionOptions())
    {
        encode(plain, std::vector<double>((size_t) slot_count_, message), scale, opt);
    }
    void encode(Plaintext<Scheme::CKKS>& plain, int message, double scale, const ExecutionOptions& opt = ExecutionOptions())
    {
        encode(plain, (double) message, scale, opt);
    }
    void decode(std::vector<double>& message, Plaintext<Scheme::CKKS>& plain, const ExecutionOptions& opt = ExecutionOptions())
    {
        std::vector<double> z;
        decode_slots(z, plain, opt);
        message.resize((size_t) slot_count_);
        for (int i = 0; i < slot_count_; ++i)
            message[i] = z[2 * i];
    }
    void decode(std::vector<Complex64>& message, Plaintext<Scheme::CKKS>& plain, const ExecutionOptions& opt = ExecutionOptions())
    {
        std::vector<double> z;
        decode_slots(z, plain, opt);
        message.resize((size_t) slot_count_);
        for (int i = 0; i < slot_count_; ++i)
            message[i] = Complex64(z[2 * i], z[2 * i + 1]);
    }

  private:
    void encode_slots(Plaintext<Scheme::CKKS>& plain, const double* z, int count, double scale, const ExecutionOptions& opt)
    {
        if (count > slot_count_)
            throw std::invalid_argument("Vector size can not be higher than slot count!");
        if (!(scale > 0))
            throw std::invalid_argument("Scale can not be negative or zero");
        const int depth = plain.depth_;
        DeviceVector<Data64> mem((size_t) (context_->Q_size - depth) * context_->n, opt.stream_);
        detail::check(heon_ckks_encode(context_->handle(), z, count, scale, depth, mem.data(), opt.stream_));
        plain.context_ = context_;
        plain.memory_set(std::move(mem));
        plain.plain_size_ = (context_->Q_size - depth) * context_->n;
        plain.scale_ = scale;
        plain.in_ntt_domain_ = true;
        plain.plaintext_generated_ = true;
        detail::output_storage(plain, opt);
    }
    void decode_slots(std::vector<double>& z, Plaintext<Scheme::CKKS>& plain, const ExecutionOptions& opt)
    {
        detail::InputGuard<Plaintext<Scheme::CKKS>> g(plain, opt, false);
        z.assign(2 * (size_t) slot_count_, 0.0);
        detail::check(heon_ckks_decode(context_->handle(), plain.data(), plain.depth_, plain.scale_, z.data(), slot_count_, opt.stream_));
    }
    HEContext<Scheme::CKKS> context_;
    int slot_count_ = 0;
};

// ---- HEEncoder<BFV> (bfv/encoder.cu): batching --------------------------------------------------------
template <> class HEEncoder<Scheme::BFV> {
  public:
    explicit HEEncoder(HEContext<Scheme::BFV> ctx) : context_(ctx)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        slot_count_ = ctx->n;
    }
    int slot_count() const { return slot_count_; }
    void encode(Plaintext<Scheme::BFV>& plain, const std::vector<uint64_t>& message, const ExecutionOptions& opt = ExecutionOptions())
    {
        encode_words(plain, message, opt);
    }
    void encode(Plaintext<Scheme::BFV>& plain, const std::vector<int64_t>& message, const ExecutionOptions& opt = ExecutionOptions())
    {
        const int64_t t = (int64_t) context_->plain_modulus_;
        std::vector<uint64_t> m(message.size());
        for (size_t i = 0; i < message.size(); ++i)
        {
            int64_t r = message[i] % t;
            m[i] = (uint64_t) (r < 0 ? r + t : r);
        }
        encode_words(plain, m, opt);
    }
    void decode(std::vector<uint64_t>& message, Plaintext<Scheme::BFV>& plain, const ExecutionOptions& opt = ExecutionOptions())
    {
        detail::InputGuard<Plaintext<Scheme::BFV>> g(plain, opt, false);
        message.assign((size_t) slot_count_, 0);
        detail::check(heon_bfv_decode(context_->handle(), plain.data(), message.data(), slot_count_, opt.stream_));
    }
    // signed form: values above t/2 come back negative (decode_kernel_bfv's signed twin)
    void decode(std::vector<int64_t>& message, Plaintext<Scheme::BFV>& plain, const ExecutionOptions& opt = ExecutionOptions())
    {
        std::vector<uint64_t> m;
        decode(m, plain, opt);
        const uint64_t t = context_->plain_modulus_;
        message.resize(m.size());
        for (size_t i = 0; i < m.size(); ++i)
            message[i] = m[i] > t / 2 ? (int64_t) m[i] - (int64_t) t : (int64_t) m[i];
    }

  private:
    void encode_words(Plaintext<Scheme::BFV>& plain, const std::vector<uint64_t>& m, const ExecutionOptions& opt)
    {
        if ((int) m.size() > slot_count_)
            throw std::invalid_argument("Vector size can not be higher than slot count!");
        DeviceVector<Data64> mem((size_t) context_->n, opt.stream_);
        detail::check(heon_bfv_encode(context_->handle(), m.data(), (int) m.size(), mem.data(), opt.stream_));
        plain.context_ = context_;
        plain.memory_set(std::move(mem));
        plain.plain_size_ = context_->n;
        plain.in_ntt_domain_ = false;
        plain.plaintext_generated_ = true;
        detail::output_storage(plain, opt);
    }
    HEContext<Scheme::BFV> context_;
    int slot_count_ = 0;
};

// ---- HEEncryptor<S>(context, public_key) (encryptor.cu) -----------------------------------------------
template <Scheme S> class HEEncryptor {
  public:
    HEEncryptor(HEContext<S> ctx, Publickey<S>& pk) : context_(ctx), public_key_(&pk)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        if (!pk.public_key_generated_)
            throw std::invalid_argument("Publickey is not generated!");
        pk.store_in_device();
        std::random_device rd;
        seed_ = ((uint64_t) rd() << 32) ^ rd();
    }
    void set_seed(uint64_t seed) { seed_ = seed, counter_ = 0; }
    void encrypt(Ciphertext<S>& ct, Plaintext<S>& pt, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!pt.plaintext_generated_)
            throw std::invalid_argument("Plaintext is not generated!");
        if constexpr (S == Scheme::CKKS)
        {
            if (pt.depth_ != 0)
                throw std::invalid_argument("A plaintext has to be at depth 0 to be encrypted");
        }
        detail::InputGuard<Plaintext<S>> g(pt, opt, false);
        DeviceVector<Data64> mem((size_t) 2 * context_->Q_size * context_->n, opt.stream_);
        detail::check(heon_encrypt(context_->handle(), public_key_->data(), pt.data(), seed_ * 0x9E3779B97F4A7C15ull + (++counter_),
                                   mem.data(), opt.stream_));
        ct.context_ = context_;
        ct.ring_size_ = context_->n;
        ct.coeff_modulus_count_ = context_->Q_size;
        ct.cipher_size_ = 2;
        ct.relinearization_required_ = false;
        ct.ciphertext_generated_ = true;
        if constexpr (S == Scheme::CKKS)
        {
            ct.depth_ = 0;
            ct.scale_ = pt.scale_;
            ct.rescale_required_ = false;
            ct.in_ntt_domain_ = true;
        }
        else
            ct.in_ntt_domain_ = false;
        ct.memory_set(std::move(mem));
        detail::output_storage(ct, opt);
    }

  private:
    HEContext<S> context_;
    Publickey<S>* public_key_;
    uint64_t seed_ = 0, counter_ = 0;
};

// ---- HEDecryptor<S>(context, secret_key) (decryptor.cu) -----------------------------------------------
template <Scheme S> class HEDecryptor {
  public:
    HEDecryptor(HEContext<S> ctx, Secretkey<S>& sk) : context_(ctx), secret_key_(&sk)
    {
        if (!ctx || !ctx->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        if (!sk.secret_key_generated_)
            throw std::invalid_argument("Secretkey is not generated!");
        sk.store_in_device();
    }
    void decrypt(Plaintext<S>& pt, Ciphertext<S>& ct, const ExecutionOptions& opt = ExecutionOptions())
    {
        detail::InputGuard<Ciphertext<S>> g(ct, opt, false);
        if constexpr (S == Scheme::CKKS)
        {
            const int L = context_->Q_size - ct.depth_;
            DeviceVector<Data64> mem((size_t) L * context_->n, opt.stream_);
            detail::check(heon_ckks_decrypt(context_->handle(), secret_key_->data(), ct.data(), ct.cipher_size_, ct.depth_,
                                            mem.data(), opt.stream_));
            pt.context_ = context_;
            pt.memory_set(std::move(mem));
            pt.plain_size_ = L * context_->n;
            pt.depth_ = ct.depth_;
            pt.scale_ = ct.scale_;
            pt.in_ntt_domain_ = true;
        }
        else
        {
            DeviceVector<Data64> mem((size_t) context_->n, opt.stream_);
            detail::check(heon_bfv_decrypt(context_->handle(), secret_key_->data(), ct.data(), ct.cipher_size_, mem.data(),
                                           opt.stream_));
            pt.context_ = context_;
            pt.memory_set(std::move(mem));
            pt.plain_size_ = context_->n;
            pt.in_ntt_domain_ = false;
        }
        pt.plaintext_generated_ = true;
        detail::output_storage(pt, opt);
    }

    // HEDecryptor<BFV>::remainder_noise_budget (bfv/decryptor.cu)
    int remainder_noise_budget(Ciphertext<S>& ct, const ExecutionOptions& opt = ExecutionOptions())
    {
        static_assert(S == Scheme::BFV, "noise budget is a BFV notion");
        detail::InputGuard<Ciphertext<S>> g(ct, opt, false);
        int bits = 0;
        detail::check(heon_bfv_noise_budget(context_->handle(), secret_key_->data(), ct.data(), ct.cipher_size_, &bits, opt.stream_));
        return bits;
    }

  private:
    HEContext<S> context_;
    Secretkey<S>* secret_key_;
};

} // namespace heongpu
