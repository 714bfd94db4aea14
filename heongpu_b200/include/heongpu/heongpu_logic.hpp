// heongpu::HELogicOperator<BFV> / <CKKS>: bit-wise gates on ciphertexts whose slots hold 0 / 1, composed from the
// arithmetic operators exactly as the reference composes them
// (src/include/heongpu/host/bfv/operator.cuh:1324-2250, src/include/heongpu/host/ckks/operator.cuh: HELogicOperator;
// constant-one plaintexts: bfv/operator.cu:1520-1543, ckks/operator.cu:7329-7346; one_minus_cipher: bfv :1545-1575,
// ckks :8195-8225):
//   NOT a = 1 - a;  a AND b = a*b;  a OR b = (a + b) - a*b;  a XOR b = (a + b) - 2*a*b;  NAND / NOR / XNOR = 1 - (...).
// Ciphertext x ciphertext gates relinearise (and, for CKKS, rescale) the product; plaintext forms use multiply_plain.
// As in the reference, the CKKS sum (a + b) stays one level above the rescaled product, so OR / XOR / NOR / XNOR of
// two ciphertexts at the same level throw "Ciphertexts leveled are not equal" unless the caller has dropped the sum's
// operands a level first; the plaintext forms take a plaintext one level down (operators.mod_drop_inplace(P)).
// Included by heongpu.hpp.  (The CKKS bit / gate bootstrapping members of the reference class are out of scope.)
#pragma once

namespace heongpu {

template <> class HELogicOperator<Scheme::BFV> : public HEOperator<Scheme::BFV> {
    using Ct = Ciphertext<Scheme::BFV>;
    using Pt = Plaintext<Scheme::BFV>;
    using Rk = Relinkey<Scheme::BFV>;
    using Opt = ExecutionOptions;

  public:
    HELogicOperator(HEContext<Scheme::BFV> context, HEEncoder<Scheme::BFV>&) : HEOperator<Scheme::BFV>(context)
    {
        std::vector<Data64> one(context_->n, 0); // the batching encoding of the all-ones vector is the constant 1
        one[0] = 1;
        one_ = Pt(context_, one);
    }
    void NOT(Ct& a, Ct& out, const Opt& o = Opt()) { one_minus(a, out, o); }
    void NOT_inplace(Ct& a, const Opt& o = Opt()) { one_minus(a, a, o); }

    void AND(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        Ct t(context_);
        multiply(a, b, t, o);
        relinearize_inplace(t, rk, o);
        out = std::move(t);
    }
    void AND_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { AND(a, b, a, rk, o); }
    void AND(Ct& a, Pt& p, Ct& out, const Opt& o = Opt()) { multiply_plain(a, p, out, o); }
    void AND_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { AND(a, p, a, o); }

    void OR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        AND(a, b, prod, rk, o);
        add(a, b, sum, o);
        sub(sum, prod, out, o);
    }
    void OR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { OR(a, b, a, rk, o); }
    void OR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        multiply_plain(a, p, prod, o);
        add_plain(a, p, sum, o);
        sub(sum, prod, out, o);
    }
    void OR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { OR(a, p, a, o); }

    void XOR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        AND(a, b, prod, rk, o);
        add(prod, prod, prod, o);
        add(a, b, sum, o);
        sub(sum, prod, out, o);
    }
    void XOR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { XOR(a, b, a, rk, o); }
    void XOR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        multiply_plain(a, p, prod, o);
        add(prod, prod, prod, o);
        add_plain(a, p, sum, o);
        sub(sum, prod, out, o);
    }
    void XOR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { XOR(a, p, a, o); }

    void NAND(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        AND(a, b, out, rk, o);
        one_minus(out, out, o);
    }
    void NAND_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { NAND(a, b, a, rk, o); }
    void NAND(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        AND(a, p, out, o);
        one_minus(out, out, o);
    }
    void NAND_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { NAND(a, p, a, o); }

    void NOR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        OR(a, b, out, rk, o);
        one_minus(out, out, o);
    }
    void NOR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { NOR(a, b, a, rk, o); }
    void NOR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        OR(a, p, out, o);
        one_minus(out, out, o);
    }
    void NOR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { NOR(a, p, a, o); }

    void XNOR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        XOR(a, b, out, rk, o);
        one_minus(out, out, o);
    }
    void XNOR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { XNOR(a, b, a, rk, o); }
    void XNOR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        XOR(a, p, out, o);
        one_minus(out, out, o);
    }
    void XNOR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { XNOR(a, p, a, o); }

  private:
    // 1 - a  (the reference negates its INPUT in place before adding the constant; the input is left alone here)
    void one_minus(Ct& a, Ct& out, const Opt& o)
    {
        Ct t(context_);
        negate(a, t, o);
        add_plain(t, one_, out, o);
    }
    Pt one_;
};

template <> class HELogicOperator<Scheme::CKKS> : public HEOperator<Scheme::CKKS> {
    using Ct = Ciphertext<Scheme::CKKS>;
    using Pt = Plaintext<Scheme::CKKS>;
    using Rk = Relinkey<Scheme::CKKS>;
    using Opt = ExecutionOptions;

  public:
    HELogicOperator(HEContext<Scheme::CKKS> context, HEEncoder<Scheme::CKKS>&, double scale = 0.0)
        : HEOperator<Scheme::CKKS>(context), scale_(scale)
    {
        if (scale == 0.0)
            throw std::invalid_argument("Scale can not be zero for CKKS Scheme");
        // the constant 1.0 at `scale`: the constant polynomial round(scale), i.e. the same residue in every NTT slot
        // (quick_ckks_encoder_constant_double, ckks/operator.cu:2588-2598)
        const int n = context_->n, Q = context_->Q_size;
        const unsigned __int128 v = (unsigned __int128) std::llround(scale);
        std::vector<Data64> words((size_t) Q * n);
        for (int i = 0; i < Q; ++i)
        {
            const Data64 r = (Data64) (v % context_->prime_vector_[i].value);
            std::fill(words.begin() + (size_t) i * n, words.begin() + (size_t) (i + 1) * n, r);
        }
        one_words_ = DeviceVector<Data64>(words);
    }
    void NOT(Ct& a, Ct& out, const Opt& o = Opt()) { one_minus(a, out, o); }
    void NOT_inplace(Ct& a, const Opt& o = Opt()) { one_minus(a, a, o); }

    void AND(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        Ct t(context_);
        multiply(a, b, t, o);
        relinearize_inplace(t, rk, o);
        rescale_inplace(t, o);
        out = std::move(t);
    }
    void AND_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { AND(a, b, a, rk, o); }
    void AND(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        Ct t(context_);
        multiply_plain(a, p, t, o);
        rescale_inplace(t, o);
        out = std::move(t);
    }
    void AND_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { AND(a, p, a, o); }

    void OR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        AND(a, b, prod, rk, o);
        add(a, b, sum, o);
        sub(sum, prod, out, o);
    }
    void OR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { OR(a, b, a, rk, o); }
    void OR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        AND(a, p, prod, o);
        add_plain(a, p, sum, o);
        sub(sum, prod, out, o);
    }
    void OR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { OR(a, p, a, o); }

    void XOR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        AND(a, b, prod, rk, o);
        add(prod, prod, prod, o);
        add(a, b, sum, o);
        sub(sum, prod, out, o);
    }
    void XOR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { XOR(a, b, a, rk, o); }
    void XOR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        Ct prod(context_), sum(context_);
        AND(a, p, prod, o);
        add(prod, prod, prod, o);
        add_plain(a, p, sum, o);
        sub(sum, prod, out, o);
    }
    void XOR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { XOR(a, p, a, o); }

    void NAND(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        AND(a, b, out, rk, o);
        one_minus(out, out, o);
    }
    void NAND_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { NAND(a, b, a, rk, o); }
    void NAND(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        AND(a, p, out, o);
        one_minus(out, out, o);
    }
    void NAND_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { NAND(a, p, a, o); }

    void NOR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        OR(a, b, out, rk, o);
        one_minus(out, out, o);
    }
    void NOR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { NOR(a, b, a, rk, o); }
    void NOR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        OR(a, p, out, o);
        one_minus(out, out, o);
    }
    void NOR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { NOR(a, p, a, o); }

    void XNOR(Ct& a, Ct& b, Ct& out, Rk& rk, const Opt& o = Opt())
    {
        XOR(a, b, out, rk, o);
        one_minus(out, out, o);
    }
    void XNOR_inplace(Ct& a, Ct& b, Rk& rk, const Opt& o = Opt()) { XNOR(a, b, a, rk, o); }
    void XNOR(Ct& a, Pt& p, Ct& out, const Opt& o = Opt())
    {
        XOR(a, p, out, o);
        one_minus(out, out, o);
    }
    void XNOR_inplace(Ct& a, Pt& p, const Opt& o = Opt()) { XNOR(a, p, a, o); }

  private:
    // 1 - a at the ciphertext's depth: the first L limbs of the constant are the constant one level down
    void one_minus(Ct& a, Ct& out, const Opt& o)
    {
        const int L = context_->Q_size - a.depth_, n = context_->n;
        std::vector<Data64> none;
        Pt one(context_);
        DeviceVector<Data64> w((size_t) L * n, o.stream_);
        detail::cuda(cudaMemcpyAsync(w.data(), one_words_.data(), (size_t) L * n * sizeof(Data64), cudaMemcpyDeviceToDevice,
                                     o.stream_));
        one.memory_set(std::move(w));
        one.depth_ = a.depth_;
        one.scale_ = a.scale_; // the reference adds the words without a scale check
        one.plain_size_ = L * n;
        one.plaintext_generated_ = true;
        Ct t(context_);
        negate(a, t, o);
        add_plain(t, one, out, o);
    }
    double scale_;
    DeviceVector<Data64> one_words_;
};

} // namespace heongpu
