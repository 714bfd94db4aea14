// heongpu_serial.hpp -- save / load of the class layer's objects in the reference's on-wire layout
// (field order and widths of Ciphertext::save ckks/ciphertext.cu:171-232, Plaintext::save
// ckks/plaintext.cu, Secretkey / Publickey::save, Relinkey::save and Galoiskey::save
// ckks/evaluationkey.cu:102-140 ff.; the BFV classes write the same fields without depth / scale) and the
// heongpu::serializer helpers (src/include/heongpu/util/serializer.h:72-131: serialize / deserialize with
// zlib compression, save_to_file / load_from_file).  Included by heongpu.hpp.
#pragma once
#include <fstream>
#include <istream>
#include <ostream>
#include <sstream>

namespace heongpu {
namespace detail {
template <class T> void put(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <class T> void get(std::istream& is, T& v)
{
    is.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!is)
        throw std::runtime_error("Invalid binary: unexpected end of stream");
}
inline uint8_t scheme_tag(Scheme s) { return s == Scheme::BFV ? 0x1 : s == Scheme::CKKS ? 0x2 : 0x0; } // scheme_type
inline void put_words(std::ostream& os, const Storable& o, size_t words)
{
    std::vector<Data64> h(words);
    if (o.is_on_device())
    {
        cuda(cudaMemcpy(h.data(), o.data(), words * sizeof(Data64), cudaMemcpyDeviceToHost));
    }
    else
        std::copy(o.host_data(), o.host_data() + words, h.begin());
    os.write(reinterpret_cast<const char*>(h.data()), (std::streamsize) (words * sizeof(Data64)));
}
inline void get_words(std::istream& is, Storable& o, size_t words)
{
    std::vector<Data64> h(words);
    is.read(reinterpret_cast<char*>(h.data()), (std::streamsize) (words * sizeof(Data64)));
    if (!is)
        throw std::runtime_error("Invalid binary: unexpected end of stream");
    DeviceVector<Data64> d(words);
    cuda(cudaMemcpy(d.data(), h.data(), words * sizeof(Data64), cudaMemcpyHostToDevice));
    o.memory_set(std::move(d));
}
inline void put_dev(std::ostream& os, const Data64* dev, size_t words)
{
    std::vector<Data64> h(words);
    cuda(cudaMemcpy(h.data(), dev, words * sizeof(Data64), cudaMemcpyDeviceToHost));
    os.write(reinterpret_cast<const char*>(h.data()), (std::streamsize) (words * sizeof(Data64)));
}
inline DeviceVector<Data64> get_dev(std::istream& is, size_t words)
{
    std::vector<Data64> h(words);
    is.read(reinterpret_cast<char*>(h.data()), (std::streamsize) (words * sizeof(Data64)));
    if (!is)
        throw std::runtime_error("Invalid binary: unexpected end of stream");
    DeviceVector<Data64> d(words);
    cuda(cudaMemcpy(d.data(), h.data(), words * sizeof(Data64), cudaMemcpyHostToDevice));
    cuda(cudaDeviceSynchronize());
    return d;
}
} // namespace detail

// ---- Ciphertext -----------------------------------------------------------------------------------------
inline void save(const Ciphertext<Scheme::CKKS>& c, std::ostream& os)
{
    if (!c.ciphertext_generated_)
        throw std::runtime_error("Ciphertext is not generated so can not be serialized!");
    detail::put(os, detail::scheme_tag(Scheme::CKKS));
    detail::put(os, (int) c.ring_size_);
    detail::put(os, (int) c.coeff_modulus_count_);
    detail::put(os, (int) c.cipher_size_);
    detail::put(os, (int) c.depth_);
    detail::put(os, (bool) c.in_ntt_domain_);
    detail::put(os, (uint8_t) c.storage_type_);
    detail::put(os, (double) c.scale_);
    detail::put(os, (uint8_t) 0); // encoding::SLOT
    detail::put(os, (bool) c.rescale_required_);
    detail::put(os, (bool) c.relinearization_required_);
    detail::put(os, (bool) c.ciphertext_generated_);
    const uint32_t words = (uint32_t) ((size_t) c.cipher_size_ * (c.coeff_modulus_count_ - c.depth_) * c.ring_size_);
    detail::put(os, words);
    detail::put_words(os, c, words);
}
inline void load(Ciphertext<Scheme::CKKS>& c, std::istream& is)
{
    if (c.ciphertext_generated_)
        throw std::runtime_error("Ciphertext has been already exist!");
    uint8_t tag, st, enc;
    bool b;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(Scheme::CKKS))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, c.ring_size_);
    detail::get(is, c.coeff_modulus_count_);
    detail::get(is, c.cipher_size_);
    detail::get(is, c.depth_);
    detail::get(is, b);
    c.in_ntt_domain_ = b;
    detail::get(is, st);
    detail::get(is, c.scale_);
    detail::get(is, enc);
    detail::get(is, b);
    c.rescale_required_ = b;
    detail::get(is, b);
    c.relinearization_required_ = b;
    detail::get(is, b);
    uint32_t words;
    detail::get(is, words);
    if (words != (uint32_t) ((size_t) c.cipher_size_ * c.ring_size_ * (c.coeff_modulus_count_ - c.depth_)))
        throw std::runtime_error("Invalid ciphertext size!");
    detail::get_words(is, c, words);
    c.ciphertext_generated_ = true;
}
inline void save(const Ciphertext<Scheme::BFV>& c, std::ostream& os)
{
    if (!c.ciphertext_generated_)
        throw std::runtime_error("Ciphertext is not generated so can not be serialized!");
    detail::put(os, detail::scheme_tag(Scheme::BFV));
    detail::put(os, (int) c.ring_size_);
    detail::put(os, (int) c.coeff_modulus_count_);
    detail::put(os, (int) c.cipher_size_);
    detail::put(os, (bool) c.in_ntt_domain_);
    detail::put(os, (uint8_t) c.storage_type_);
    detail::put(os, (bool) c.relinearization_required_);
    detail::put(os, (bool) c.ciphertext_generated_);
    const uint32_t words = (uint32_t) ((size_t) c.cipher_size_ * c.coeff_modulus_count_ * c.ring_size_);
    detail::put(os, words);
    detail::put_words(os, c, words);
}
inline void load(Ciphertext<Scheme::BFV>& c, std::istream& is)
{
    if (c.ciphertext_generated_)
        throw std::runtime_error("Ciphertext has been already exist!");
    uint8_t tag, st;
    bool b;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(Scheme::BFV))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, c.ring_size_);
    detail::get(is, c.coeff_modulus_count_);
    detail::get(is, c.cipher_size_);
    detail::get(is, b);
    c.in_ntt_domain_ = b;
    detail::get(is, st);
    detail::get(is, b);
    c.relinearization_required_ = b;
    detail::get(is, b);
    uint32_t words;
    detail::get(is, words);
    if (words != (uint32_t) ((size_t) c.cipher_size_ * c.ring_size_ * c.coeff_modulus_count_))
        throw std::runtime_error("Invalid ciphertext size!");
    detail::get_words(is, c, words);
    c.ciphertext_generated_ = true;
}

// ---- Plaintext ------------------------------------------------------------------------------------------
inline void save(const Plaintext<Scheme::CKKS>& p, std::ostream& os)
{
    if (!p.plaintext_generated_)
        throw std::runtime_error("Plaintext is not generated so can not be serialized!");
    detail::put(os, detail::scheme_tag(Scheme::CKKS));
    detail::put(os, (int) p.plain_size_);
    detail::put(os, (int) p.depth_);
    detail::put(os, (double) p.scale_);
    detail::put(os, (bool) p.in_ntt_domain_);
    detail::put(os, (uint8_t) 0); // encoding::SLOT
    detail::put(os, (bool) p.plaintext_generated_);
    detail::put(os, (uint8_t) p.storage_type_);
    detail::put(os, (int) p.plain_size_);
    detail::put_words(os, p, (size_t) p.plain_size_);
}
inline void load(Plaintext<Scheme::CKKS>& p, std::istream& is)
{
    if (p.plaintext_generated_)
        throw std::runtime_error("Plaintext has been already exist!");
    uint8_t tag, st, enc;
    bool b;
    int again;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(Scheme::CKKS))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, p.plain_size_);
    detail::get(is, p.depth_);
    detail::get(is, p.scale_);
    detail::get(is, b);
    p.in_ntt_domain_ = b;
    detail::get(is, enc);
    detail::get(is, b);
    detail::get(is, st);
    detail::get(is, again);
    if (again != p.plain_size_ || p.plain_size_ < 0)
        throw std::runtime_error("Invalid plaintext size!");
    detail::get_words(is, p, (size_t) p.plain_size_);
    p.plaintext_generated_ = true;
}
inline void save(const Plaintext<Scheme::BFV>& p, std::ostream& os)
{
    if (!p.plaintext_generated_)
        throw std::runtime_error("Plaintext is not generated so can not be serialized!");
    detail::put(os, detail::scheme_tag(Scheme::BFV));
    detail::put(os, (int) p.plain_size_);
    detail::put(os, (bool) p.in_ntt_domain_);
    detail::put(os, (bool) p.plaintext_generated_);
    detail::put(os, (uint8_t) p.storage_type_);
    detail::put(os, (int) p.plain_size_);
    detail::put_words(os, p, (size_t) p.plain_size_);
}
inline void load(Plaintext<Scheme::BFV>& p, std::istream& is)
{
    if (p.plaintext_generated_)
        throw std::runtime_error("Plaintext has been already exist!");
    uint8_t tag, st;
    bool b;
    int again;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(Scheme::BFV))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, p.plain_size_);
    detail::get(is, b);
    p.in_ntt_domain_ = b;
    detail::get(is, b);
    detail::get(is, st);
    detail::get(is, again);
    if (again != p.plain_size_ || p.plain_size_ < 0)
        throw std::runtime_error("Invalid plaintext size!");
    detail::get_words(is, p, (size_t) p.plain_size_);
    p.plaintext_generated_ = true;
}

// ---- Secretkey / Publickey (both schemes) ---------------------------------------------------------------
template <Scheme S> void save(const Secretkey<S>& k, std::ostream& os)
{
    if (!k.secret_key_generated_)
        throw std::runtime_error("Secretkey is not generated so can not be serialized!");
    int n_power = 0;
    while ((1 << n_power) < k.ring_size_)
        ++n_power;
    detail::put(os, detail::scheme_tag(S));
    detail::put(os, (int) k.ring_size_);
    detail::put(os, (int) k.coeff_modulus_count_);
    detail::put(os, n_power);
    detail::put(os, (int) k.hamming_weight_);
    detail::put(os, (bool) k.in_ntt_domain_);
    detail::put(os, (bool) k.secret_key_generated_);
    detail::put(os, (uint8_t) k.storage_type_);
    const uint32_t words = (uint32_t) ((size_t) k.coeff_modulus_count_ * k.ring_size_);
    detail::put(os, words);
    detail::put_words(os, k, words);
}
template <Scheme S> void load(Secretkey<S>& k, std::istream& is)
{
    if (k.secret_key_generated_)
        throw std::runtime_error("Secretkey has been already exist!");
    uint8_t tag, st;
    bool b;
    int n_power;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(S))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, k.ring_size_);
    detail::get(is, k.coeff_modulus_count_);
    detail::get(is, n_power);
    detail::get(is, k.hamming_weight_);
    detail::get(is, b);
    k.in_ntt_domain_ = b;
    detail::get(is, b);
    detail::get(is, st);
    uint32_t words;
    detail::get(is, words);
    if (words != (uint32_t) ((size_t) k.coeff_modulus_count_ * k.ring_size_))
        throw std::runtime_error("Invalid secretkey size!");
    detail::get_words(is, k, words);
    k.secret_key_generated_ = true;
}
template <Scheme S> void save(const Publickey<S>& k, std::ostream& os)
{
    if (!k.public_key_generated_)
        throw std::runtime_error("Publickey is not generated so can not be serialized!");
    detail::put(os, detail::scheme_tag(S));
    detail::put(os, (int) k.ring_size_);
    detail::put(os, (int) k.coeff_modulus_count_);
    detail::put(os, (bool) k.in_ntt_domain_);
    detail::put(os, (bool) k.public_key_generated_);
    detail::put(os, (uint8_t) k.storage_type_);
    const uint32_t words = (uint32_t) ((size_t) 2 * k.coeff_modulus_count_ * k.ring_size_);
    detail::put(os, words);
    detail::put_words(os, k, words);
}
template <Scheme S> void load(Publickey<S>& k, std::istream& is)
{
    if (k.public_key_generated_)
        throw std::runtime_error("Publickey has been already exist!");
    uint8_t tag, st;
    bool b;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(S))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, k.ring_size_);
    detail::get(is, k.coeff_modulus_count_);
    detail::get(is, b);
    k.in_ntt_domain_ = b;
    detail::get(is, b);
    detail::get(is, st);
    uint32_t words;
    detail::get(is, words);
    if (words != (uint32_t) ((size_t) 2 * k.coeff_modulus_count_ * k.ring_size_))
        throw std::runtime_error("Invalid publickey size!");
    detail::get_words(is, k, words);
    k.public_key_generated_ = true;
}

// ---- Relinkey (ckks/evaluationkey.cu:102-140; bfv twin) ------------------------------------------------
template <Scheme S> void save(const Relinkey<S>& k, std::ostream& os)
{
    if (!k.relin_key_generated_)
        throw std::runtime_error("Relinkey is not generated so can not be serialized!");
    const int d = k.d_;
    const Data64 words = (Data64) d * 2 * k.Q_prime_size_ * k.ring_size;
    detail::put(os, detail::scheme_tag(S));
    detail::put(os, (uint8_t) k.key_type);
    detail::put(os, (int) k.ring_size);
    detail::put(os, (int) k.Q_prime_size_);
    detail::put(os, (int) k.Q_size_);
    detail::put(os, d);
    detail::put(os, (int) 0); // d_tilda_ (Method III, unused)
    detail::put(os, (int) 0); // r_prime_
    detail::put(os, (uint8_t) storage_type::DEVICE);
    detail::put(os, (bool) true);
    detail::put(os, words);
    const detail::KeyView kv_ = k.view(cudaStreamDefault); // stages a host-stored key
    detail::put_dev(os, kv_.ptr, (size_t) words);
}
template <Scheme S> void load(Relinkey<S>& k, std::istream& is)
{
    if (k.relin_key_generated_)
        throw std::runtime_error("Relinkey has been already exist!");
    uint8_t tag, kt, st;
    bool b;
    int n, qp, q, d, dt, rp;
    Data64 words;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(S))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, kt);
    detail::get(is, n);
    detail::get(is, qp);
    detail::get(is, q);
    detail::get(is, d);
    detail::get(is, dt);
    detail::get(is, rp);
    detail::get(is, st);
    detail::get(is, b);
    detail::get(is, words);
    if (words != (Data64) d * 2 * qp * n)
        throw std::runtime_error("Invalid relinkey binary!");
    if (k.context_) // an object bound to a context only accepts that context's keys
    {
        const auto& c = *k.context_;
        if (n != c.n || qp != c.Q_prime_size || q != c.Q_size || d != detail::digits0(c))
            throw std::runtime_error("Invalid relinkey binary for this context!");
    }
    k.key_type = (keyswitching_type) kt;
    k.ring_size = n, k.Q_prime_size_ = qp, k.Q_size_ = q, k.d_ = d;
    k.device_location_ = detail::get_dev(is, (size_t) words);
    k.relin_key_generated_ = true;
}

// ---- Galoiskey (ckks/evaluationkey.cu: Galoiskey::save / load) ------------------------------------------
template <Scheme S> void save(const Galoiskey<S>& k, std::ostream& os)
{
    if (!k.galois_key_generated_)
        throw std::runtime_error("Galoiskey is not generated so can not be serialized!");
    const int d = k.d_;
    const Data64 words = (Data64) d * 2 * k.Q_prime_size_ * k.ring_size;
    detail::put(os, detail::scheme_tag(S));
    detail::put(os, (uint8_t) k.key_type);
    detail::put(os, (int) k.ring_size);
    detail::put(os, (int) k.Q_prime_size_);
    detail::put(os, (int) k.Q_size_);
    detail::put(os, d);
    detail::put(os, (bool) k.customized);
    detail::put(os, (int) k.group_order_);
    detail::put(os, (uint8_t) storage_type::DEVICE);
    detail::put(os, (bool) true);
    if (k.customized)
    {
        std::vector<uint32_t> elts(k.custom_galois_elt);
        for (const auto& kv : k.galois_elt)
            elts.push_back((uint32_t) kv.second);
        detail::put(os, (uint32_t) elts.size());
        for (uint32_t e : elts)
            detail::put(os, e);
    }
    else
    {
        detail::put(os, (uint32_t) k.galois_elt.size());
        for (const auto& kv : k.galois_elt)
        {
            detail::put(os, (int) kv.first);
            detail::put(os, (int) kv.second);
        }
    }
    detail::put(os, (int) k.galois_elt_zero);
    detail::put(os, words);
    std::vector<int> present; // keys on the device or (store_in_host) in host memory
    for (const auto& kv : k.device_location_)
        present.push_back(kv.first);
    for (const auto& kv : k.host_location_)
        present.push_back(kv.first);
    uint32_t count = 0;
    for (int e : present)
        if (e != k.galois_elt_zero || S == Scheme::CKKS)
            ++count;
    detail::put(os, count);
    for (int e : present)
    {
        if (!(e != k.galois_elt_zero || S == Scheme::CKKS))
            continue;
        detail::put(os, e);
        const detail::KeyView v = k.view(e, cudaStreamDefault);
        detail::put_dev(os, v.ptr, (size_t) words);
    }
    {
        const detail::KeyView v = k.zero_view(cudaStreamDefault); // the conjugation / column-rotation key
        if (!v)
            throw std::logic_error("Galois key not present!");
        detail::put_dev(os, v.ptr, (size_t) words);
    }
}
template <Scheme S> void load(Galoiskey<S>& k, std::istream& is)
{
    if (k.galois_key_generated_)
        throw std::runtime_error("Galoiskey has been already exist!");
    uint8_t tag, kt, st;
    bool b, customized;
    int n, qp, q, d, order, zero;
    Data64 words;
    detail::get(is, tag);
    if (tag != detail::scheme_tag(S))
        throw std::runtime_error("Invalid scheme binary!");
    detail::get(is, kt);
    detail::get(is, n);
    detail::get(is, qp);
    detail::get(is, q);
    detail::get(is, d);
    detail::get(is, customized);
    detail::get(is, order);
    detail::get(is, st);
    detail::get(is, b);
    if (k.context_)
    {
        const auto& c = *k.context_;
        if (n != c.n || qp != c.Q_prime_size || q != c.Q_size || d != detail::digits0(c))
            throw std::runtime_error("Invalid galoiskey binary for this context!");
    }
    k.key_type = (keyswitching_type) kt;
    k.ring_size = n, k.Q_prime_size_ = qp, k.Q_size_ = q, k.d_ = d;
    k.customized = customized;
    k.group_order_ = order;
    uint32_t cnt;
    detail::get(is, cnt);
    k.galois_elt.clear();
    k.custom_galois_elt.clear();
    for (uint32_t i = 0; i < cnt; ++i)
    {
        if (customized)
        {
            uint32_t e;
            detail::get(is, e);
            k.custom_galois_elt.push_back(e);
        }
        else
        {
            int a, e;
            detail::get(is, a);
            detail::get(is, e);
            k.galois_elt[a] = e;
        }
    }
    detail::get(is, zero);
    detail::get(is, words);
    if (words != (Data64) d * 2 * qp * n)
        throw std::runtime_error("Invalid galoiskey size!");
    uint32_t count;
    detail::get(is, count);
    for (uint32_t i = 0; i < count; ++i)
    {
        int elt;
        detail::get(is, elt);
        k.device_location_[elt] = detail::get_dev(is, (size_t) words);
    }
    k.set_zero_key(zero, detail::get_dev(is, (size_t) words));
    k.galois_key_generated_ = true;
}

// ---- member forms: object.save(os) / object.load(is) -----------------------------------------------------
#define HEON_SERIAL_MEMBERS(TYPE)                                                                  \
    inline void TYPE::save(std::ostream& os) const { heongpu::save(*this, os); }                    \
    inline void TYPE::load(std::istream& is) { heongpu::load(*this, is); }
HEON_SERIAL_MEMBERS(Ciphertext<Scheme::CKKS>)
HEON_SERIAL_MEMBERS(Ciphertext<Scheme::BFV>)
HEON_SERIAL_MEMBERS(Plaintext<Scheme::CKKS>)
HEON_SERIAL_MEMBERS(Plaintext<Scheme::BFV>)
HEON_SERIAL_MEMBERS(Relinkey<Scheme::CKKS>)
HEON_SERIAL_MEMBERS(Relinkey<Scheme::BFV>)
HEON_SERIAL_MEMBERS(Galoiskey<Scheme::CKKS>)
HEON_SERIAL_MEMBERS(Galoiskey<Scheme::BFV>)
#undef HEON_SERIAL_MEMBERS
template <Scheme S> void Secretkey<S>::save(std::ostream& os) const { heongpu::save(*this, os); }
template <Scheme S> void Secretkey<S>::load(std::istream& is) { heongpu::load(*this, is); }
template <Scheme S> void Publickey<S>::save(std::ostream& os) const { heongpu::save(*this, os); }
template <Scheme S> void Publickey<S>::load(std::istream& is) { heongpu::load(*this, is); }

// ---- heongpu::serializer (serializer.h:72-131) -----------------------------------------------------------
namespace serializer {
inline std::vector<uint8_t> compress(const std::vector<uint8_t>& data)
{
    size_t cap = heon_compress_bound(data.size());
    std::vector<uint8_t> out(cap);
    detail::check(heon_compress(data.data(), data.size(), out.data(), &cap));
    out.resize(cap);
    return out;
}
inline std::vector<uint8_t> decompress(const std::vector<uint8_t>& data)
{
    size_t cap = data.size() * 4 + 1024;
    for (;;)
    {
        std::vector<uint8_t> out(cap);
        size_t len = cap;
        const int rc = heon_decompress(data.data(), data.size(), out.data(), &len);
        if (rc == HEON_OK)
        {
            out.resize(len);
            return out;
        }
        if (rc != HEON_ERR_LOGIC)
            throw std::runtime_error("Zlib decompression failed");
        cap *= 2; // the reference guesses 4x once (serializer.cpp:35-50); high-entropy data can need more
    }
}
template <class T> std::vector<uint8_t> serialize(const T& obj)
{
    std::stringstream ss;
    obj.save(ss);
    const std::string str = ss.str();
    return compress(std::vector<uint8_t>(str.begin(), str.end()));
}
// objects of this class layer are bound to a context at construction, so deserialize takes the empty object
// to fill (the reference default-constructs and re-binds, serializer.h:84-95)
template <class T> T& deserialize(const std::vector<uint8_t>& buffer, T& obj)
{
    const std::vector<uint8_t> raw = decompress(buffer);
    std::stringstream ss;
    ss.str(std::string(raw.begin(), raw.end()));
    obj.load(ss);
    return obj;
}
template <class T> T deserialize(const std::vector<uint8_t>& buffer)
{
    T obj;
    deserialize(buffer, obj);
    return obj;
}
template <class T> void save_to_file(const T& obj, const std::string& filename)
{
    const std::vector<uint8_t> data = serialize(obj);
    const uint64_t size = data.size();
    std::ofstream ofs(filename, std::ios::binary);
    if (!ofs)
        throw std::runtime_error("Cannot open file for writing: " + filename);
    ofs.write(reinterpret_cast<const char*>(&size), sizeof(size));
    ofs.write(reinterpret_cast<const char*>(data.data()), (std::streamsize) size);
}
template <class T> T& load_from_file(const std::string& filename, T& obj)
{
    std::ifstream ifs(filename, std::ios::binary);
    if (!ifs)
        throw std::runtime_error("Cannot open file for reading: " + filename);
    uint64_t size = 0;
    ifs.read(reinterpret_cast<char*>(&size), sizeof(size));
    std::vector<uint8_t> buffer(size);
    ifs.read(reinterpret_cast<char*>(buffer.data()), (std::streamsize) size);
    return deserialize(buffer, obj);
}
template <class T> T load_from_file(const std::string& filename)
{
    T obj;
    load_from_file(filename, obj);
    return obj;
}
} // namespace serializer
} // namespace heongpu
