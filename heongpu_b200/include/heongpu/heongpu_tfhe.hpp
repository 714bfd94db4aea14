// heongpu:: TFHE classes over the C ABI (heon_tfhe_*, include/heon_b200.h), source-compatible with the
// reference's public interface for Scheme::TFHE so that test/test_tfhe_gate_boot.cpp and
// example/basic/15_basic_tfhe.cpp compile unmodified:
//   HEContext<TFHE> / GenHEContext<TFHE>()                 src/include/heongpu/host/tfhe/context.cuh
//   Secretkey<TFHE>, Bootstrappingkey<TFHE>                 .../tfhe/secretkey.cuh, evaluationkey.cuh
//   Ciphertext<TFHE>                                        .../tfhe/ciphertext.cuh
//   HEKeyGenerator<TFHE>, HEEncryptor<TFHE>, HEDecryptor<TFHE>   .../tfhe/{keygenerator,encryptor,decryptor}.cuh
//   HELogicOperator<TFHE>::{NAND,AND,NOR,OR,XNOR,XOR,NOT,MUX}    .../tfhe/operator.cuh:29-812
// Included by heongpu.hpp.
#pragma once
#include <random>

namespace heongpu {

template <> class HEContextImpl<Scheme::TFHE> {
  public:
    explicit HEContextImpl(sec_level_type = sec_level_type::sec128, int device = 0)
    {
        detail::check(heon_tfhe_create(device, &h_));
        int p[7];
        detail::check(heon_tfhe_params(h_, p));
        n_ = p[0], N_ = p[1], k_ = p[2], bk_l_ = p[3], bk_bg_bit_ = p[4], ks_base_bit_ = p[5], ks_length_ = p[6];
        const double s = std::sqrt(2.0 / 3.14159265358979323846);
        ks_stdev_ = (1.0 / 32768.0) * s;
        bk_stdev_ = 9e-9 * s;
        max_stdev_ = (1.0 / 64.0) * s;
    }
    ~HEContextImpl() { heon_tfhe_destroy(h_); }
    HEContextImpl(const HEContextImpl&) = delete;
    HEContextImpl& operator=(const HEContextImpl&) = delete;
    heon_tfhe_t handle() const { return h_; }
    int n_, N_, k_, bk_l_, bk_bg_bit_, ks_base_bit_, ks_length_;
    double ks_stdev_, bk_stdev_, max_stdev_;

  private:
    heon_tfhe_t h_ = nullptr;
};

template <> class Secretkey<Scheme::TFHE> {
  public:
    explicit Secretkey(HEContext<Scheme::TFHE> ctx) : context_(ctx)
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
        n_ = ctx->n_;
        lwe_alpha_min = ctx->ks_stdev_;
        lwe_alpha_max = ctx->max_stdev_;
    }
    HEContext<Scheme::TFHE> context_;
    int n_;
    double lwe_alpha_min, lwe_alpha_max;
    DeviceVector<int32_t> lwe_key_device_location_, tlwe_key_device_location_;
    bool secret_key_generated_ = false;
    storage_type storage_type_ = storage_type::DEVICE;
};

template <Scheme S> class Bootstrappingkey;
template <> class Bootstrappingkey<Scheme::TFHE> {
  public:
    explicit Bootstrappingkey(HEContext<Scheme::TFHE> ctx) : context_(ctx)
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
        bk_k_ = ctx->k_, bk_base_bit_ = ctx->bk_bg_bit_, bk_length_ = ctx->bk_l_, bk_stdev_ = ctx->bk_stdev_;
        ks_base_bit_ = ctx->ks_base_bit_, ks_length_ = ctx->ks_length_;
    }
    HEContext<Scheme::TFHE> context_;
    int bk_k_, bk_base_bit_, bk_length_, ks_base_bit_, ks_length_;
    double bk_stdev_;
    DeviceVector<Data64> boot_key_device_location_; // [n][k+1][l][k+1][N], NTT domain
    DeviceVector<int32_t> switch_key_device_location_a_, switch_key_device_location_b_;
    std::vector<double> switch_key_variances_;
    bool boot_key_generated_ = false;
    storage_type storage_type_ = storage_type::DEVICE;
};

template <> class Ciphertext<Scheme::TFHE> {
  public:
    Ciphertext() = default;
    explicit Ciphertext(HEContext<Scheme::TFHE> ctx, const ExecutionOptions& = ExecutionOptions())
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
        n_ = ctx->n_;
        alpha_min_ = ctx->ks_stdev_;
        alpha_max_ = ctx->max_stdev_;
    }
    int size() const { return shape_; }
    int n_ = 0, shape_ = 0;
    double alpha_min_ = 0, alpha_max_ = 0;
    DeviceVector<int32_t> a_device_location_, b_device_location_;
    std::vector<double> variances_;
    bool ciphertext_generated_ = false;
    storage_type storage_type_ = storage_type::DEVICE;
};

template <> class HEKeyGenerator<Scheme::TFHE> {
  public:
    explicit HEKeyGenerator(HEContext<Scheme::TFHE> ctx) : context_(ctx)
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
        std::random_device rd;
        seed_ = ((uint64_t) rd() << 32) | rd();
    }
    void set_seed(uint64_t s) { seed_ = s; }
    void generate_secret_key(Secretkey<Scheme::TFHE>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (sk.secret_key_generated_)
            throw std::runtime_error("Secretkey is already generated!");
        sk.lwe_key_device_location_.resize(context_->n_, opt.stream_);
        sk.tlwe_key_device_location_.resize((size_t) context_->k_ * context_->N_, opt.stream_);
        detail::check(heon_tfhe_keygen_secret(context_->handle(), seed_++, sk.lwe_key_device_location_.data(),
                                              sk.tlwe_key_device_location_.data(), opt.stream_));
        sk.secret_key_generated_ = true;
    }
    void generate_bootstrapping_key(Bootstrappingkey<Scheme::TFHE>& bk, Secretkey<Scheme::TFHE>& sk,
                                    const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!sk.secret_key_generated_)
            throw std::logic_error("Secretkey is not generated!");
        if (bk.boot_key_generated_)
            throw std::logic_error("Bootkey is already generated!");
        auto& c = *context_;
        const size_t rows = (size_t) c.k_ * c.N_ * c.ks_length_ * ((1 << c.ks_base_bit_) - 1);
        bk.boot_key_device_location_.resize((size_t) c.n_ * (c.k_ + 1) * c.bk_l_ * (c.k_ + 1) * c.N_, opt.stream_);
        bk.switch_key_device_location_a_.resize(rows * c.n_, opt.stream_);
        bk.switch_key_device_location_b_.resize(rows, opt.stream_);
        detail::check(heon_tfhe_keygen_boot(c.handle(), sk.lwe_key_device_location_.data(), sk.tlwe_key_device_location_.data(),
                                            seed_++, bk.boot_key_device_location_.data(),
                                            bk.switch_key_device_location_a_.data(), bk.switch_key_device_location_b_.data(),
                                            opt.stream_));
        bk.switch_key_variances_.assign(rows, c.ks_stdev_ * c.ks_stdev_);
        bk.boot_key_generated_ = true;
    }

  private:
    HEContext<Scheme::TFHE> context_;
    uint64_t seed_;
};

template <> class HEEncryptor<Scheme::TFHE> {
  public:
    HEEncryptor(HEContext<Scheme::TFHE> ctx, Secretkey<Scheme::TFHE>& sk) : context_(ctx), sk_(&sk)
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
        if (!sk.secret_key_generated_)
            throw std::runtime_error("Secretkey was not generated!");
        std::random_device rd;
        seed_ = ((uint64_t) rd() << 32) | rd();
    }
    // true -> +1/8, false -> -1/8 on the torus (tfhe/encryptor.cuh: encrypt)
    void encrypt(Ciphertext<Scheme::TFHE>& ct, const std::vector<bool>& messages, const ExecutionOptions& opt = ExecutionOptions())
    {
        const int32_t mu = encode_to_torus32(1, 8);
        std::vector<int32_t> enc(messages.size());
        for (size_t i = 0; i < messages.size(); ++i)
            enc[i] = messages[i] ? mu : -mu;
        auto& c = *context_;
        ct.shape_ = (int) messages.size();
        ct.n_ = c.n_;
        DeviceVector<int32_t> dm(enc, opt.stream_);
        ct.a_device_location_.resize((size_t) ct.shape_ * c.n_, opt.stream_);
        ct.b_device_location_.resize(ct.shape_, opt.stream_);
        detail::check(heon_tfhe_encrypt(c.handle(), sk_->lwe_key_device_location_.data(), dm.data(), seed_++,
                                        ct.a_device_location_.data(), ct.b_device_location_.data(), ct.shape_, opt.stream_));
        detail::cuda(cudaStreamSynchronize(opt.stream_)); // `enc` is on this stack frame
        ct.variances_.assign(ct.shape_, c.ks_stdev_ * c.ks_stdev_);
        ct.ciphertext_generated_ = true;
    }
    static int32_t encode_to_torus32(uint32_t mu, uint32_t m_size)
    {
        const uint64_t interval = ((1ULL << 63) / m_size) * 2;
        return (int32_t) ((mu * interval) >> 32);
    }

  private:
    HEContext<Scheme::TFHE> context_;
    Secretkey<Scheme::TFHE>* sk_;
    uint64_t seed_;
};

template <> class HEDecryptor<Scheme::TFHE> {
  public:
    HEDecryptor(HEContext<Scheme::TFHE> ctx, Secretkey<Scheme::TFHE>& sk) : context_(ctx), sk_(&sk)
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
    }
    void decrypt(Ciphertext<Scheme::TFHE>& ct, std::vector<bool>& messages, const ExecutionOptions& opt = ExecutionOptions())
    {
        DeviceVector<int32_t> phase(ct.shape_, opt.stream_);
        detail::check(heon_tfhe_phase(context_->handle(), sk_->lwe_key_device_location_.data(), ct.a_device_location_.data(),
                                      ct.b_device_location_.data(), phase.data(), ct.n_, ct.shape_, opt.stream_));
        std::vector<int32_t> h(ct.shape_);
        detail::cuda(cudaMemcpyAsync(h.data(), phase.data(), h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, opt.stream_));
        detail::cuda(cudaStreamSynchronize(opt.stream_));
        messages.resize(ct.shape_);
        for (int i = 0; i < ct.shape_; ++i)
            messages[i] = h[i] > 0;
    }

  private:
    HEContext<Scheme::TFHE> context_;
    Secretkey<Scheme::TFHE>* sk_;
};

template <Scheme S> class HELogicOperator;
template <> class HELogicOperator<Scheme::TFHE> {
    using Ct = Ciphertext<Scheme::TFHE>;
    using Bk = Bootstrappingkey<Scheme::TFHE>;

  public:
    explicit HELogicOperator(HEContext<Scheme::TFHE> ctx) : context_(ctx)
    {
        if (!ctx)
            throw std::invalid_argument("HEContext is not set!");
    }
    void NAND(Ct& a, Ct& b, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_NAND, a, &b, nullptr, out, &bk, o); }
    void AND(Ct& a, Ct& b, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_AND, a, &b, nullptr, out, &bk, o); }
    void NOR(Ct& a, Ct& b, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_NOR, a, &b, nullptr, out, &bk, o); }
    void OR(Ct& a, Ct& b, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_OR, a, &b, nullptr, out, &bk, o); }
    void XNOR(Ct& a, Ct& b, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_XNOR, a, &b, nullptr, out, &bk, o); }
    void XOR(Ct& a, Ct& b, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_XOR, a, &b, nullptr, out, &bk, o); }
    void NOT(Ct& a, Ct& out, const ExecutionOptions& o = ExecutionOptions()) { gate(HEON_TFHE_NOT, a, nullptr, nullptr, out, nullptr, o); }
    // MUX(input1, input2, control): control ? input1 : input2 (operator.cuh:688-812)
    void MUX(Ct& in1, Ct& in2, Ct& control, Ct& out, Bk& bk, const ExecutionOptions& o = ExecutionOptions())
    {
        if (in1.shape_ != control.shape_)
            throw std::runtime_error("Ciphertexts size should be equal!");
        gate(HEON_TFHE_MUX, in1, &in2, &control, out, &bk, o);
    }

  private:
    void gate(int code, Ct& a, Ct* b, Ct* c, Ct& out, Bk* bk, const ExecutionOptions& o)
    {
        if (b && a.shape_ != b->shape_)
            throw std::runtime_error(c ? "Ciphertexts size should be equal!" : "Both ciphertexts size should be equal!");
        if (!a.ciphertext_generated_ || (b && !b->ciphertext_generated_))
            throw std::runtime_error("One or the inputs are generated!");
        if (bk && !bk->boot_key_generated_)
            throw std::runtime_error("Bootkey is not generated!");
        auto& x = *context_;
        DeviceVector<int32_t> oa((size_t) a.shape_ * x.n_, o.stream_), ob(a.shape_, o.stream_);
        detail::check(heon_tfhe_gate(x.handle(), code, a.a_device_location_.data(), a.b_device_location_.data(),
                                     b ? b->a_device_location_.data() : nullptr, b ? b->b_device_location_.data() : nullptr,
                                     c ? c->a_device_location_.data() : nullptr, c ? c->b_device_location_.data() : nullptr,
                                     oa.data(), ob.data(), bk ? bk->boot_key_device_location_.data() : nullptr,
                                     bk ? bk->switch_key_device_location_a_.data() : nullptr,
                                     bk ? bk->switch_key_device_location_b_.data() : nullptr, a.shape_, o.stream_));
        out.n_ = a.n_;
        out.shape_ = a.shape_;
        out.variances_ = a.variances_;
        out.alpha_min_ = a.alpha_min_;
        out.alpha_max_ = a.alpha_max_;
        out.a_device_location_ = std::move(oa);
        out.b_device_location_ = std::move(ob);
        out.ciphertext_generated_ = true;
        out.storage_type_ = storage_type::DEVICE;
    }
    HEContext<Scheme::TFHE> context_;
};

} // namespace heongpu
