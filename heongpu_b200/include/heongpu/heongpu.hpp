// heongpu.hpp -- source-level mirror of the HEonGPU class layer for the hot path,
// implemented on top of the C ABI in include/heon_b200.h (libheon_b200.so).
//
// Mirrors, with the same names / argument meaning / exception types:
//   heongpu::Scheme, sec_level_type, keyswitching_type, storage_type        src/include/heongpu/util/schemes.h:15-31
//   heongpu::ExecutionOptions                                               src/include/heongpu/util/storagemanager.cuh:34-97
//   heongpu::HEContext<S> / GenHEContext<S>(...)                            src/include/heongpu/host/ckks/context.cuh
//   heongpu::Ciphertext<S>                                                  src/include/heongpu/host/ckks/ciphertext.cuh:77-209
//   heongpu::Relinkey<S>, Galoiskey<S>                                      src/include/heongpu/host/ckks/evaluationkey.cuh
//   heongpu::Plaintext<S>, Switchkey<S>                                    src/include/heongpu/host/ckks/{plaintext,evaluationkey}.cuh
//   heongpu::HEArithmeticOperator<S>::{add,sub,negate,multiply,multiply_inplace,
//       multiply_plain,add_plain[_inplace],sub_plain[_inplace],
//       relinearize_inplace,rescale_inplace,mod_drop_inplace,mod_drop,
//       rotate_rows,rotate_rows_inplace,apply_galois,apply_galois_inplace,
//       keyswitch,conjugate}                                                src/include/heongpu/host/ckks/operator.cuh:95-1600
//   the BFV twins (HEContext<BFV>, Ciphertext<BFV>, Relinkey<BFV>, Galoiskey<BFV>,
//       HEArithmeticOperator<BFV>::{add,sub,negate,multiply,relinearize_inplace,
//       rotate_rows,rotate_columns,apply_galois})                           src/include/heongpu/host/bfv/operator.cuh
//
// Scope: the CKKS and BFV hot path.  Key generation, encoding and
// encryption are client-side "next" rows: Relinkey / Galoiskey here are
// containers in the reference layout that the caller fills (set_data), and
// Ciphertext can be constructed from raw words.  Host-parsable (no CUDA
// language extensions), like the reference's public headers.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iosfwd>
#include <iostream>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../../include/heon_b200.h"

typedef std::uint64_t Data64;
// Modulus<Data64> {value, bit, mu} with mu = floor(2^(2*bit+1) / value) and the host Barrett product of
// OPERATOR<Data64> (thirdparty/GPU-NTT/src/include/gpuntt/common/modular_arith.cuh:28-60, 90-107); both
// live at global namespace in the reference and its tests use them directly
// (test/test_bfv_multiplication.cpp:50-58).
struct Modulus64 {
    Data64 value = 0, bit = 0, mu = 0;
    Modulus64() = default;
    Modulus64(Data64 v, Data64 b, Data64 m) : value(v), bit(b), mu(m) {}
    Modulus64(Data64 v) : value(v)
    {
        for (Data64 t = v; t; t >>= 1)
            ++bit;
        mu = (Data64) ((((unsigned __int128) 1) << (2 * bit + 1)) / v);
    }
};
struct OPERATOR64 {
    static Data64 mult(const Data64& a, const Data64& b, const Modulus64& m)
    {
        unsigned __int128 z = (unsigned __int128) a * b;
        unsigned __int128 r = z >> (m.bit - 2);
        r = r * (unsigned __int128) m.mu;
        r = r >> (m.bit + 3);
        r = r * (unsigned __int128) m.value;
        z = z - r;
        const Data64 res = (Data64) z;
        return res >= m.value ? res - m.value : res;
    }
    static Data64 add(const Data64& a, const Data64& b, const Modulus64& m)
    {
        const Data64 s = a + b;
        return s >= m.value ? s - m.value : s;
    }
    static Data64 sub(const Data64& a, const Data64& b, const Modulus64& m)
    {
        const Data64 d = a + m.value - b;
        return d >= m.value ? d - m.value : d;
    }
};

namespace heongpu {

enum class Scheme { BFV = 1, CKKS = 2, TFHE = 3 };
enum class sec_level_type { none, sec128 = 128, sec192 = 192, sec256 = 256 };
enum class keyswitching_type { NONE = 0, KEYSWITCHING_METHOD_I = 1, KEYSWITCHING_METHOD_II = 2 };
enum class storage_type { HOST = 1, DEVICE = 2 };

struct ExecutionOptions {
    cudaStream_t stream_ = cudaStreamDefault;
    storage_type storage_ = storage_type::DEVICE;
    bool keep_initial_condition_ = true;
    ExecutionOptions& set_stream(cudaStream_t s)
    {
        stream_ = s;
        return *this;
    }
    ExecutionOptions& set_storage_type(storage_type s)
    {
        storage_ = s;
        return *this;
    }
    ExecutionOptions& set_initial_location(bool k)
    {
        keep_initial_condition_ = k;
        return *this;
    }
};

namespace detail {
inline void check(int status)
{
    if (status == HEON_OK)
        return;
    const std::string msg = heon_last_error();
    if (status == HEON_ERR_INVALID)
        throw std::invalid_argument(msg);
    if (status == HEON_ERR_LOGIC)
        throw std::logic_error(msg);
    throw std::runtime_error(msg);
}
inline void cuda(cudaError_t e)
{
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e));
}
} // namespace detail

// Stream-ordered device buffer (the reference's DeviceVector is an
// rmm::device_uvector over a process-wide pool; here cudaMallocAsync's pool).
template <typename T> class DeviceVector {
  public:
    DeviceVector() = default;
    explicit DeviceVector(size_t n, cudaStream_t st = cudaStreamDefault) { resize(n, st); }
    DeviceVector(const std::vector<T>& h, cudaStream_t st = cudaStreamDefault)
    {
        resize(h.size(), st);
        detail::cuda(cudaMemcpyAsync(ptr_, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    DeviceVector(DeviceVector&& o) noexcept { swap(o); }
    DeviceVector& operator=(DeviceVector&& o) noexcept
    {
        swap(o);
        return *this;
    }
    DeviceVector(const DeviceVector&) = delete;
    DeviceVector& operator=(const DeviceVector&) = delete;
    ~DeviceVector()
    {
        if (ptr_)
            cudaFreeAsync(ptr_, stream_);
    }
    void resize(size_t n, cudaStream_t st = cudaStreamDefault)
    {
        if (n == size_)
            return;
        T* p = nullptr;
        if (n)
            detail::cuda(cudaMallocAsync((void**) &p, n * sizeof(T), st));
        if (ptr_ && p)
            detail::cuda(cudaMemcpyAsync(p, ptr_, (n < size_ ? n : size_) * sizeof(T), cudaMemcpyDeviceToDevice, st));
        if (ptr_)
            cudaFreeAsync(ptr_, st);
        ptr_ = p;
        size_ = n;
        stream_ = st;
    }
    T* data() const { return ptr_; }
    size_t size() const { return size_; }
    void swap(DeviceVector& o)
    {
        std::swap(ptr_, o.ptr_);
        std::swap(size_, o.size_);
        std::swap(stream_, o.stream_);
    }

  private:
    T* ptr_ = nullptr;
    size_t size_ = 0;
    cudaStream_t stream_ = cudaStreamDefault;
};
template <typename T> using HostVector = std::vector<T>;

// Pinned host buffer (the reference's HostVector is an rmm pinned-pool vector, hostvector.cuh).
template <typename T> class PinnedVector {
  public:
    PinnedVector() = default;
    PinnedVector(PinnedVector&& o) noexcept { swap(o); }
    PinnedVector& operator=(PinnedVector&& o) noexcept
    {
        swap(o);
        return *this;
    }
    PinnedVector(const PinnedVector&) = delete;
    PinnedVector& operator=(const PinnedVector&) = delete;
    ~PinnedVector() { clear(); }
    void resize(size_t n)
    {
        if (n == size_)
            return;
        clear();
        if (n)
            detail::cuda(cudaMallocHost((void**) &ptr_, n * sizeof(T)));
        size_ = n;
    }
    void clear()
    {
        if (ptr_)
            cudaFreeHost(ptr_);
        ptr_ = nullptr;
        size_ = 0;
    }
    T* data() const { return ptr_; }
    size_t size() const { return size_; }
    void swap(PinnedVector& o)
    {
        std::swap(ptr_, o.ptr_);
        std::swap(size_, o.size_);
    }

  private:
    T* ptr_ = nullptr;
    size_t size_ = 0;
};

namespace detail {
// Where an object's words live and how they move: store_in_host / store_in_device / copy_to_device /
// remove_from_host / remove_from_device / is_on_device of the reference's Ciphertext, Plaintext and key
// classes (ckks/ciphertext.cuh:120-209).
class Storable {
  public:
    Data64* data() const { return device_locations_.data(); }
    Data64* host_data() const { return host_locations_.data(); }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        if (storage_type_ == storage_type::DEVICE)
            return;
        copy_to_device(st);
        detail::cuda(cudaStreamSynchronize(st)); // the pinned source is released next
        host_locations_.clear();
    }
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        if (storage_type_ == storage_type::HOST)
            return;
        host_locations_.resize(device_locations_.size());
        detail::cuda(cudaMemcpyAsync(host_locations_.data(), device_locations_.data(), device_locations_.size() * sizeof(Data64),
                                     cudaMemcpyDeviceToHost, st));
        detail::cuda(cudaStreamSynchronize(st));
        device_locations_ = DeviceVector<Data64>();
        storage_type_ = storage_type::HOST;
    }
    // device copy next to the host copy (the host words stay valid)
    void copy_to_device(cudaStream_t st = cudaStreamDefault)
    {
        if (storage_type_ == storage_type::DEVICE)
            return;
        DeviceVector<Data64> d(host_locations_.size(), st);
        detail::cuda(cudaMemcpyAsync(d.data(), host_locations_.data(), host_locations_.size() * sizeof(Data64),
                                     cudaMemcpyHostToDevice, st));
        device_locations_ = std::move(d);
        storage_type_ = storage_type::DEVICE;
    }
    void remove_from_device(cudaStream_t st = cudaStreamDefault)
    {
        // back to the host state kept by copy_to_device (the words did not change)
        if (host_locations_.size() == 0)
        {
            store_in_host(st);
            return;
        }
        device_locations_ = DeviceVector<Data64>();
        storage_type_ = storage_type::HOST;
    }
    void remove_from_host() { host_locations_.clear(); }
    size_t memory_size() const { return is_on_device() ? device_locations_.size() : host_locations_.size(); }
    void memory_set(DeviceVector<Data64>&& v)
    {
        device_locations_ = std::move(v);
        host_locations_.clear();
        storage_type_ = storage_type::DEVICE;
    }

    DeviceVector<Data64> device_locations_;
    PinnedVector<Data64> host_locations_;
    storage_type storage_type_ = storage_type::DEVICE;
};

// input_storage_manager (storagemanager.cuh:113-167): brings an operand to the device for the call and,
// when the call ends, leaves it where ExecutionOptions says (truth table: README.md:349-366).
template <class T> struct InputGuard {
    T& o;
    ExecutionOptions opt;
    storage_type initial;
    bool same;
    InputGuard(T& obj, const ExecutionOptions& op, bool is_input_output_same = false)
        : o(obj), opt(op), initial(obj.storage_type_), same(is_input_output_same)
    {
        if (!o.is_on_device())
        {
            if (opt.keep_initial_condition_ && !same)
                o.copy_to_device(opt.stream_);
            else
                o.store_in_device(opt.stream_);
        }
    }
    ~InputGuard() noexcept(false)
    {
        if (same)
            return; // the result's location is set by output_storage
        if (opt.keep_initial_condition_)
        {
            if (initial == storage_type::HOST)
                o.remove_from_device(opt.stream_);
        }
        else if (opt.storage_ == storage_type::HOST)
            o.store_in_host(opt.stream_);
        else
            o.store_in_device(opt.stream_);
    }
};
// output_storage_manager: the result goes where set_storage_type says
template <class T> void output_storage(T& out, const ExecutionOptions& opt)
{
    if (opt.storage_ == storage_type::HOST)
        out.store_in_host(opt.stream_);
}
} // namespace detail


namespace detail {
// heongpu_{128,192,256}bit_std_parms (src/include/heongpu/util/secstdparams.h:24-76): the largest total
// coefficient-modulus bit count a ring degree supports at a security level
inline int max_total_bits(sec_level_type sec, size_t n)
{
    static const int tab[3][5] = {{109, 218, 438, 881, 1761}, {74, 149, 300, 605, 1212}, {57, 115, 232, 465, 930}};
    const int row = sec == sec_level_type::sec128 ? 0 : sec == sec_level_type::sec192 ? 1 : 2;
    const int col = n == 4096 ? 0 : n == 8192 ? 1 : n == 16384 ? 2 : n == 32768 ? 3 : n == 65536 ? 4 : -1;
    return col < 0 ? 0 : tab[row][col];
}
inline int bit_size(Data64 v)
{
    int b = 0;
    while (v)
    {
        ++b;
        v >>= 1;
    }
    return b;
}
// set_coeff_modulus_*: "Parameters do not align with the security recommendations" (ckks/context.cu:113,
// bfv/context.cu:113,223) unless the level is sec_level_type::none
inline void check_security(sec_level_type sec, size_t n, int total_bits)
{
    if (sec == sec_level_type::none)
        return;
    if (total_bits > max_total_bits(sec, n))
        throw std::runtime_error("Parameters do not align with the security recommendations.");
}
// A rotation whose own key is absent is carried out as a chain of the +-2^i rotations the key set holds
// (rotate_ckks_method_I/II, operator.cu:1338-1420; rotation_index_generator :2411-2480): the shift is taken
// to its balanced representative modulo the row size, written in non-adjacent form, and digits above
// 2^max_shift are folded into repeated +-2^max_shift steps.
inline std::vector<int> rotation_plan(int shift, int log_slots, int max_shift)
{
    const long long mod = 1LL << log_slots;
    long long x = ((shift % mod) + mod) % mod;
    if (x > mod / 2)
        x -= mod;
    std::vector<int> plan;
    for (int i = 0; x != 0; ++i, x >>= 1)
    {
        if (!(x & 1))
            continue;
        const int d = 2 - (int) (x & 3); // +1 or -1
        x -= d;
        if (i <= max_shift)
            plan.push_back(d * (1 << i));
        else
            for (long long r = 0; r < (1LL << (i - max_shift)); ++r)
                plan.push_back(d * (1 << max_shift));
    }
    return plan;
}
// default 128-bit-security coefficient moduli (src/lib/util/defaultmodulus.cpp:12-90; the SEAL defaults)
inline std::vector<Data64> default_modulus_128(size_t n)
{
    switch (n)
    {
        case 4096: return {0x800004001, 0x800008001, 0x1000002001};
        case 8192: return {0x40000084001, 0x400000b0001, 0x8000002c001, 0x80000050001, 0x80000064001};
        case 16384:
            return {0x800000020001,  0x8000001a8001,  0x8000001e8001,  0x10000000d8001, 0x1000000168001,
                    0x10000001a0001, 0x10000001e0001, 0x10000002b8001, 0x10000002e8001};
        case 32768:
            return {0x2000000002b0001, 0x2000000003a0001, 0x2000000005b0001, 0x200000000640001, 0x400000000270001,
                    0x400000000350001, 0x400000000360001, 0x4000000004d0001, 0x400000000570001, 0x400000000660001,
                    0x4000000008a0001, 0x400000000920001, 0x400000000980001, 0x400000000990001, 0x400000000a40001};
        case 65536:
            return {0x2000000003a0001, 0x200000000640001, 0x200000000f80001, 0x200000001460001, 0x2000000015a0001,
                    0x2000000015e0001, 0x200000001b20001, 0x200000001c00001, 0x200000001ee0001, 0x400000000360001,
                    0x400000000660001, 0x4000000008a0001, 0x400000000920001, 0x400000000980001, 0x400000000a40001,
                    0x400000000c00001, 0x400000000ea0001, 0x400000001460001, 0x400000001700001, 0x400000001740001,
                    0x4000000017a0001, 0x400000001920001, 0x400000001b00001, 0x400000001b60001, 0x400000001c40001,
                    0x400000001ee0001, 0x400000001f20001, 0x4000000020c0001, 0x400000002360001, 0x400000002480001};
    }
    throw std::logic_error("no default modulus for this poly_modulus_degree");
}
} // namespace detail

namespace detail {
// Evaluation keys may live in host memory (ExecutionOptions::set_storage_type(HOST), store_in_host()): the
// operators then stage the one key they need on their stream for the duration of the call, as the reference does
// (e.g. ckks/operator.cu:3117-3131: DeviceVector<Data64> key_location(galois_key.host_location_[elt], stream)).
struct KeyView {
    DeviceVector<Data64> staged;
    const Data64* ptr = nullptr;
    KeyView() = default;
    KeyView(const DeviceVector<Data64>& dev, const PinnedVector<Data64>& host, cudaStream_t st)
    {
        if (dev.data())
            ptr = dev.data();
        else if (host.data())
        {
            staged.resize(host.size(), st);
            cuda(cudaMemcpyAsync(staged.data(), host.data(), host.size() * sizeof(Data64), cudaMemcpyHostToDevice, st));
            ptr = staged.data();
        }
    }
    KeyView(KeyView&&) = default;
    KeyView& operator=(KeyView&&) = default;
    explicit operator bool() const { return ptr != nullptr; }
};
inline void key_to_host(DeviceVector<Data64>& dev, PinnedVector<Data64>& host, cudaStream_t st)
{
    if (!dev.data())
        return;
    host.resize(dev.size());
    cuda(cudaMemcpyAsync(host.data(), dev.data(), dev.size() * sizeof(Data64), cudaMemcpyDeviceToHost, st));
    cuda(cudaStreamSynchronize(st));
    dev = DeviceVector<Data64>();
}
inline void key_to_device(DeviceVector<Data64>& dev, PinnedVector<Data64>& host, cudaStream_t st)
{
    if (!host.data())
        return;
    dev.resize(host.size(), st);
    cuda(cudaMemcpyAsync(dev.data(), host.data(), host.size() * sizeof(Data64), cudaMemcpyHostToDevice, st));
    cuda(cudaStreamSynchronize(st));
    host.clear();
}
} // namespace detail

// ---- MemoryPoolConfig / MemoryPool (src/include/heongpu/util/memorypool.cuh:38-140) ---------------------------
// The reference keeps an RMM pool behind a process-wide singleton.  This engine allocates stream-ordered from the
// CUDA driver's per-device memory pool (cudaMallocAsync); the same configuration surface is mapped onto it: the
// `max_*` limits become the pool's release threshold (memory above it goes back to the driver at the next
// synchronisation), `initial_*` pre-reserves that much by one allocate / free, and `use_memory_pool(false)` sets the
// threshold to zero (every free returns memory to the driver, i.e. plain cudaMalloc / cudaFree behaviour).
// Fractions accept 0.0-1.0 (ratio) or 0-100 (percentage), as in the reference.
struct MemoryPoolConfig {
    std::optional<float> initial_device_fraction, max_device_fraction;
    std::optional<size_t> initial_device_bytes, max_device_bytes;
    std::optional<float> initial_host_fraction, max_host_fraction;
    std::optional<size_t> initial_host_bytes, max_host_bytes;
    bool use_memory_pool = true;
    static MemoryPoolConfig Defaults()
    {
        MemoryPoolConfig c;
        c.initial_device_fraction = 0.5f; // memorypool.cu: half of the free device memory up front, 80 % at most
        c.max_device_fraction = 0.8f;
        return c;
    }
};
class MemoryPool {
  public:
    static MemoryPool& instance()
    {
        static MemoryPool pool;
        return pool;
    }
    void initialize() { initialize(MemoryPoolConfig::Defaults()); }
    void initialize(const MemoryPoolConfig& config)
    {
        config_ = config;
        int dev = 0;
        detail::cuda(cudaGetDevice(&dev));
        cudaMemPool_t pool;
        detail::cuda(cudaDeviceGetDefaultMemPool(&pool, dev));
        size_t free_b = 0, total_b = 0;
        detail::cuda(cudaMemGetInfo(&free_b, &total_b));
        auto frac = [](float f) { return f > 1.0f ? f / 100.0f : f; };
        unsigned long long threshold = ~0ull; // keep everything: allocation never goes back to the driver mid-run
        if (config.max_device_bytes)
            threshold = *config.max_device_bytes;
        else if (config.max_device_fraction)
            threshold = (unsigned long long) ((double) total_b * frac(*config.max_device_fraction));
        if (!config.use_memory_pool)
            threshold = 1; // effectively nothing stays cached (0 would read as "not configured" to the library)
        detail::cuda(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        size_t initial = 0;
        if (config.initial_device_bytes)
            initial = *config.initial_device_bytes;
        else if (config.initial_device_fraction)
            initial = (size_t) ((double) free_b * frac(*config.initial_device_fraction));
        if (config.use_memory_pool && initial > 0)
        {
            initial = std::min<size_t>(initial, (size_t) std::min<unsigned long long>(threshold, free_b / 10 * 9));
            void* p = nullptr;
            if (cudaMallocAsync(&p, initial, cudaStreamDefault) == cudaSuccess)
                cudaFreeAsync(p, cudaStreamDefault); // stays reserved in the pool (below the release threshold)
            else
                cudaGetLastError();
        }
        initialized_ = true;
    }
    void use_memory_pool(bool use)
    {
        config_.use_memory_pool = use;
        initialize(config_);
    }
    bool is_initialized() const { return initialized_; }
    void print_memory_pool_status() const
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess)
            return;
        unsigned long long reserved = 0, used = 0, reserved_high = 0, used_high = 0, threshold = 0;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &reserved_high);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &used_high);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        const double mb = 1.0 / (1024.0 * 1024.0);
        std::cout << "==== Memory pool status (device " << dev << ", CUDA stream-ordered pool) ====" << std::endl;
        std::cout << "-->   reserved: " << reserved * mb << " MB (peak " << reserved_high * mb << " MB)" << std::endl;
        std::cout << "-->   in use:   " << used * mb << " MB (peak " << used_high * mb << " MB)" << std::endl;
        if (threshold == ~0ull)
            std::cout << "-->   release threshold: unlimited" << std::endl;
        else
            std::cout << "-->   release threshold: " << threshold * mb << " MB" << std::endl;
    }

  private:
    MemoryPool() = default;
    MemoryPoolConfig config_ = MemoryPoolConfig::Defaults();
    bool initialized_ = false;
};

namespace detail {
// Like the reference, the class layer is single-device per host thread: ciphertexts, plaintexts and keys allocate on
// the CURRENT CUDA device.  A context created for another device would hand the operators buffers from the wrong
// GPU, so generate() refuses it (cudaSetDevice(device) first; the C ABI itself takes any device per call).
inline void require_current_device(int device)
{
    int cur = -1;
    cuda(cudaGetDevice(&cur));
    if (cur != device)
        throw std::invalid_argument("HEContext: the context's device (" + std::to_string(device) +
                                    ") is not the current CUDA device (" + std::to_string(cur) +
                                    "); call cudaSetDevice first");
}
} // namespace detail

template <Scheme S> class HEContextImpl;
template <Scheme S> using HEContext = std::shared_ptr<HEContextImpl<S>>;

template <> class HEContextImpl<Scheme::CKKS> {
  public:
    explicit HEContextImpl(sec_level_type sec = sec_level_type::sec128, int device = 0) : sec_level_(sec), device_(device) {}
    ~HEContextImpl()
    {
        if (h_)
            heon_context_destroy(h_);
    }
    void set_poly_modulus_degree(size_t n)
    {
        if (coeff_modulus_specified_ || poly_modulus_degree_specified_)
            throw std::logic_error("Poly modulus degree cannot be changed after the coeff_modulus is specified!");
        if (n == 0 || (n & (n - 1)))
            throw std::logic_error("Poly modulus degree have to be power of two");
        if (n > 65536 || n < 4096)
            throw std::logic_error("Poly modulus degree is not supported");
        this->n = (int) n;
        n_power = 0;
        while ((size_t(1) << n_power) < n)
            ++n_power;
        poly_modulus_degree_specified_ = true;
    }
    void set_coeff_modulus_bit_sizes(const std::vector<int>& q_bits, const std::vector<int>& p_bits)
    {
        if (coeff_modulus_specified_ || context_generated_ || !poly_modulus_degree_specified_)
            throw std::logic_error("Coeff_modulus cannot be changed after the context is generated!");
        if (p_bits.empty())
            throw std::logic_error("log_P_bases_bit_sizes cannot be empty!");
        int total = 0;
        for (int b : q_bits)
            total += b;
        for (int b : p_bits)
            total += b;
        detail::check_security(sec_level_, (size_t) n, total);
        q_bits_ = q_bits;
        p_bits_ = p_bits;
        by_value_ = false;
        coeff_modulus_specified_ = true;
    }
    void set_coeff_modulus_values(const std::vector<Data64>& q, const std::vector<Data64>& p)
    {
        if (coeff_modulus_specified_ || context_generated_ || !poly_modulus_degree_specified_)
            throw std::logic_error("Coeff_modulus cannot be changed after the context is generated!");
        if (p.empty())
            throw std::logic_error("log_P_bases_bit_sizes cannot be empty!");
        int total = 0;
        for (Data64 v : q)
            total += detail::bit_size(v);
        for (Data64 v : p)
            total += detail::bit_size(v);
        detail::check_security(sec_level_, (size_t) n, total);
        q_vals_ = q;
        p_vals_ = p;
        by_value_ = true;
        coeff_modulus_specified_ = true;
    }
    // generate(const MemoryPoolConfig&) (context.cu: the pool is configured before the tables are built)
    void generate(const MemoryPoolConfig& pool_config)
    {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(device_);
        MemoryPool::instance().initialize(pool_config);
        cudaSetDevice(prev);
        generate();
    }
    void generate()
    {
        if (context_generated_ || !poly_modulus_degree_specified_ || !coeff_modulus_specified_)
            throw std::runtime_error("Context is already generated!");
        detail::require_current_device(device_);
        if (by_value_)
            detail::check(heon_ckks_context_create_values(device_, n_power, q_vals_.data(), (int) q_vals_.size(),
                                                          p_vals_.data(), (int) p_vals_.size(), &h_));
        else
            detail::check(heon_ckks_context_create(device_, n_power, q_bits_.data(), (int) q_bits_.size(),
                                                   p_bits_.data(), (int) p_bits_.size(), &h_));
        heon_info info;
        detail::check(heon_context_info(h_, &info));
        Q_size = info.q_size;
        P_size = info.p_size;
        Q_prime_size = Q_size + P_size;
        keyswitching_type_ = info.keyswitch_method == 1 ? keyswitching_type::KEYSWITCHING_METHOD_I
                                                        : keyswitching_type::KEYSWITCHING_METHOD_II;
        size_t cnt = 0;
        detail::check(heon_context_table(h_, HEON_TBL_MODULUS, 0, nullptr, 0, &cnt));
        std::vector<Data64> raw(cnt);
        detail::check(heon_context_table(h_, HEON_TBL_MODULUS, 0, raw.data(), cnt, &cnt));
        prime_vector_.clear();
        for (size_t i = 0; i < cnt; i += 3)
            prime_vector_.push_back(Modulus64{raw[i], raw[i + 1], raw[i + 2]});
        context_generated_ = true;
    }
    // ckks/context.cu:576-700: the reference serialises the parameters chosen at set_coeff_modulus time (the primes
    // included) and load() regenerates the context from them.  This class picks its primes inside generate(), so an
    // ungenerated context asks a scratch context for them first.
    void save(std::ostream& os) const
    {
        if (!poly_modulus_degree_specified_ || !coeff_modulus_specified_)
            throw std::runtime_error("Context has no enough parameters to serialize!");
        std::vector<Modulus64> primes = prime_vector_;
        int q_size = Q_size, p_size = P_size;
        uint8_t ks = (uint8_t) keyswitching_type_;
        if (!context_generated_)
        {
            HEContextImpl scratch(sec_level_type::none, device_);
            scratch.set_poly_modulus_degree((size_t) n);
            if (by_value_)
                scratch.set_coeff_modulus_values(q_vals_, p_vals_);
            else
                scratch.set_coeff_modulus_bit_sizes(q_bits_, p_bits_);
            scratch.generate();
            primes = scratch.prime_vector_;
            q_size = scratch.Q_size;
            p_size = scratch.P_size;
            ks = (uint8_t) scratch.keyswitching_type_;
        }
        auto put = [&](const auto& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(v)); };
        put((uint8_t) 0x2); // scheme_type::ckks
        put((uint8_t) sec_level_);
        put(ks);
        put((int) n);
        put((int) n_power);
        put((int) (q_size + p_size)); // coeff_modulus
        int total_bits = 0;
        for (const auto& m : primes)
            total_bits += (int) m.bit;
        put(total_bits);
        put((int) (q_size + p_size));
        put(q_size);
        put(p_size);
        put((uint32_t) primes.size());
        os.write(reinterpret_cast<const char*>(primes.data()), (std::streamsize) (sizeof(Modulus64) * primes.size()));
        put((uint32_t) q_size); // base_q
        for (int i = 0; i < q_size; ++i)
            put((Data64) primes[i].value);
        auto put_bits = [&](int from, int to) {
            put((uint32_t) (to - from));
            for (int i = from; i < to; ++i)
                put((int) primes[i].bit);
        };
        put_bits(0, q_size + p_size);
        put_bits(0, q_size);
        put_bits(q_size, q_size + p_size);
    }
    void load(std::istream& is)
    {
        if (context_generated_)
            throw std::runtime_error("Context has been already exist!");
        auto get = [&](auto& v) {
            is.read(reinterpret_cast<char*>(&v), sizeof(v));
            if (!is)
                throw std::runtime_error("Invalid context binary!");
        };
        uint8_t scheme, sec, ks;
        int coeff_modulus, total_bits, qp, q, pz;
        get(scheme);
        if (scheme != 0x2)
            throw std::runtime_error("Invalid scheme binary!");
        get(sec);
        get(ks);
        get(n);
        get(n_power);
        get(coeff_modulus);
        get(total_bits);
        get(qp);
        get(q);
        get(pz);
        uint32_t cnt;
        get(cnt);
        std::vector<Modulus64> primes(cnt);
        is.read(reinterpret_cast<char*>(primes.data()), (std::streamsize) (sizeof(Modulus64) * cnt));
        for (int v = 0; v < 4; ++v) // base_q and the three bit-size vectors
        {
            get(cnt);
            is.ignore((std::streamsize) cnt * (v == 0 ? sizeof(Data64) : sizeof(int)));
        }
        if (!is || q + pz != (int) primes.size() || q < 1 || pz < 1)
            throw std::runtime_error("Invalid context binary!");
        sec_level_ = (sec_level_type) sec;
        q_vals_.clear();
        p_vals_.clear();
        for (int i = 0; i < q; ++i)
            q_vals_.push_back(primes[i].value);
        for (int i = q; i < q + pz; ++i)
            p_vals_.push_back(primes[i].value);
        by_value_ = true;
        poly_modulus_degree_specified_ = true;
        coeff_modulus_specified_ = true;
        generate();
    }
    int digit_count(int depth) const
    {
        const int L = Q_size - depth;
        return P_size == 1 ? L : (L + P_size - 1) / P_size;
    }
    heon_context_t handle() const { return h_; }
    size_t get_poly_modulus_degree() const { return (size_t) n; }
    int get_log_poly_modulus_degree() const { return n_power; }
    int get_ciphertext_modulus_count() const { return Q_size; }
    int get_key_modulus_count() const { return Q_prime_size; }
    std::vector<Modulus64> get_key_modulus() const { return prime_vector_; }
    // ckks/context.cu: print_parameters
    void print_parameters() const
    {
        if (!context_generated_)
        {
            std::cout << "Parameters is not generated yet!" << std::endl;
            return;
        }
        std::cout << "==== HEonGPU a GPU Based Homomorphic Encryption Library ====\n" << std::endl;
        std::cout << "Encryption parameters:" << std::endl;
        std::cout << "-->   scheme: " << "CKKS" << std::endl;
        std::cout << "-->   poly_modulus_degree: " << n << std::endl;
        std::cout << "-->   Q_tilta size: Q( ";
        for (int i = 0; i < Q_size; ++i)
            std::cout << prime_vector_[i].bit << (i + 1 < Q_size ? " + " : "");
        std::cout << " ) + P( ";
        for (int i = Q_size; i < Q_prime_size; ++i)
            std::cout << prime_vector_[i].bit << (i + 1 < Q_prime_size ? " + " : "");
        std::cout << " ) bits" << std::endl;
        std::cout << std::endl;
    }

    int n = 0, n_power = 0;
    int Q_size = 0, P_size = 0, Q_prime_size = 0;
    keyswitching_type keyswitching_type_ = keyswitching_type::NONE;
    std::vector<Modulus64> prime_vector_;
    bool context_generated_ = false;

  private:
    sec_level_type sec_level_;
    int device_;
    heon_context_t h_ = nullptr;
    std::vector<int> q_bits_, p_bits_;
    std::vector<Data64> q_vals_, p_vals_;
    bool by_value_ = false, poly_modulus_degree_specified_ = false, coeff_modulus_specified_ = false;
};

template <Scheme S> HEContext<S> GenHEContext(sec_level_type sec = sec_level_type::sec128, int device = 0)
{
    return std::make_shared<HEContextImpl<S>>(sec, device);
}

template <Scheme S> class Ciphertext;
template <> class Ciphertext<Scheme::CKKS> : public detail::Storable {
  public:
    Ciphertext() = default;
    void save(std::ostream& os) const;
    void load(std::istream& is);
    // empty ciphertext bound to a context (ckks/ciphertext.cu:8-31); filled by the encryptor / operators
    explicit Ciphertext(HEContext<Scheme::CKKS> ctx, const ExecutionOptions& = ExecutionOptions()) : context_(ctx)
    {
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_size;
        cipher_size_ = 2;
        scale_ = 0;
    }
    // [cipher_size][L][N] words, NTT domain (ciphertext.cu:20-31)
    Ciphertext(HEContext<Scheme::CKKS> ctx, const std::vector<Data64>& words, int cipher_size = 2, int depth = 0,
               double scale = 1.0, const ExecutionOptions& opt = ExecutionOptions())
        : context_(ctx), cipher_size_(cipher_size), depth_(depth), scale_(scale)
    {
        device_locations_ = DeviceVector<Data64>(words, opt.stream_);
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_size;
        if (words.size() < (size_t) cipher_size * (ctx->Q_size - depth) * ctx->n)
            throw std::invalid_argument("Invalid Ciphertexts size!");
        ciphertext_generated_ = true;
    }
    void get_data(std::vector<Data64>& out, cudaStream_t st = cudaStreamDefault) const
    {
        out.resize((size_t) cipher_size_ * (coeff_modulus_count_ - depth_) * ring_size_);
        if (!is_on_device())
        {
            std::copy(host_data(), host_data() + out.size(), out.begin());
            return;
        }
        detail::cuda(cudaMemcpyAsync(out.data(), data(), out.size() * sizeof(Data64), cudaMemcpyDeviceToHost, st));
        detail::cuda(cudaStreamSynchronize(st));
    }
    int level() const { return coeff_modulus_count_ - depth_; }
    int size() const { return cipher_size_; }
    int depth() const { return depth_; }
    double scale() const { return scale_; }
    bool in_ntt_domain() const { return in_ntt_domain_; }
    bool rescale_required() const { return rescale_required_; }
    bool relinearization_required() const { return relinearization_required_; }

    HEContext<Scheme::CKKS> context_;
    int ring_size_ = 0, coeff_modulus_count_ = 0, cipher_size_ = 0, depth_ = 0;
    double scale_ = 0;
    bool in_ntt_domain_ = true, rescale_required_ = false, relinearization_required_ = false,
         ciphertext_generated_ = false;
};

template <Scheme S> class Relinkey;
template <> class Relinkey<Scheme::CKKS> {
  public:
    Relinkey() = default; // filled by load() (serializer::deserialize / load_from_file)
    explicit Relinkey(HEContext<Scheme::CKKS> ctx) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        ring_size = ctx->n, Q_prime_size_ = ctx->Q_prime_size, Q_size_ = ctx->Q_size, d_ = ctx->digit_count(0);
    }
    int ring_size = 0, Q_prime_size_ = 0, Q_size_ = 0, d_ = 0;
    void save(std::ostream& os) const;
    void load(std::istream& is);
    // [digit][2][Q'_0][N] NTT-domain words (keygeneration.cu:180-183)
    void set_data(const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count(0) * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid relinearization key size!");
        device_location_ = DeviceVector<Data64>(words, st);
        relin_key_generated_ = true;
    }
    Data64* data() const { return device_location_.data(); }
    HEContext<Scheme::CKKS> context_;
    keyswitching_type key_type = keyswitching_type::NONE;
    storage_type storage_type_ = storage_type::DEVICE;
    DeviceVector<Data64> device_location_;
    PinnedVector<Data64> host_location_;
    // evaluationkey.cuh: store_in_host / store_in_device / is_on_device
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_host(device_location_, host_location_, st);
        storage_type_ = storage_type::HOST;
    }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_device(device_location_, host_location_, st);
        storage_type_ = storage_type::DEVICE;
    }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    detail::KeyView view(cudaStream_t st) const { return detail::KeyView(device_location_, host_location_, st); }
    bool relin_key_generated_ = false;
};

template <Scheme S> class Galoiskey;
template <> class Galoiskey<Scheme::CKKS> {
  public:
    Galoiskey() = default; // filled by load()
    int ring_size = 0, Q_prime_size_ = 0, Q_size_ = 0, d_ = 0;
    void bind_meta(const HEContextImpl<Scheme::CKKS>& c_) { ring_size = c_.n, Q_prime_size_ = c_.Q_prime_size, Q_size_ = c_.Q_size, d_ = c_.digit_count(0); }
    explicit Galoiskey(HEContext<Scheme::CKKS> ctx) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        bind_meta(*ctx);
        for (int i = 0; i < 8; ++i) // default keys for +-2^i, i < MAX_SHIFT (evaluationkey.cu:306-345)
        {
            galois_elt[1 << i] = heon_steps_to_galois_elt(1 << i, ctx->n, group_order_);
            galois_elt[-(1 << i)] = heon_steps_to_galois_elt(-(1 << i), ctx->n, group_order_);
        }
    }
    Galoiskey(HEContext<Scheme::CKKS> ctx, const std::vector<int>& shifts) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        bind_meta(*ctx);
        customized = true;
        for (int s : shifts)
            galois_elt[s] = heon_steps_to_galois_elt(s, ctx->n, group_order_);
    }
    void set_key(int galois_element, const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count(0) * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid galois key size!");
        device_location_[galois_element] = DeviceVector<Data64>(words, st);
    }
    // conjugation key (galois_elt_zero = 2N-1), Galoiskey::c_data() in the reference
    void set_conjugate_key(const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count(0) * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid galois key size!");
        zero_device_location_ = DeviceVector<Data64>(words, st);
        galois_elt_zero = 2 * context_->n - 1;
    }
    Data64* c_data() const { return zero_device_location_.data(); }
    // keys by explicit Galois element (evaluationkey.cu: Galoiskey(context, std::vector<uint32_t>))
    Galoiskey(HEContext<Scheme::CKKS> ctx, const std::vector<uint32_t>& elts)
        : context_(ctx), key_type(ctx->keyswitching_type_), custom_galois_elt(elts)
    {
        bind_meta(*ctx);
        customized = true;
    }
    void set_zero_key(int elt, DeviceVector<Data64>&& key)
    {
        zero_device_location_ = std::move(key);
        galois_elt_zero = elt;
    }
    const Data64* zero_key_data() const { return zero_device_location_.data(); }
    void save(std::ostream& os) const;
    void load(std::istream& is);
    std::vector<uint32_t> custom_galois_elt;
    bool galois_key_generated_ = false;
    int max_shift_ = 7; // MAX_SHIFT - 1 (evaluationkey.cu:428)
    HEContext<Scheme::CKKS> context_;
    keyswitching_type key_type = keyswitching_type::NONE;
    storage_type storage_type_ = storage_type::DEVICE;
    int group_order_ = 5;
    bool customized = false;
    int galois_elt_zero = 0;
    std::unordered_map<int, int> galois_elt; // shift -> galois element
    std::unordered_map<int, DeviceVector<Data64>> device_location_; // galois element -> key
    DeviceVector<Data64> zero_device_location_;
    std::unordered_map<int, PinnedVector<Data64>> host_location_;
    PinnedVector<Data64> zero_host_location_;
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        for (auto& kv : device_location_)
            detail::key_to_host(kv.second, host_location_[kv.first], st);
        device_location_.clear();
        detail::key_to_host(zero_device_location_, zero_host_location_, st);
        storage_type_ = storage_type::HOST;
    }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        for (auto& kv : host_location_)
            detail::key_to_device(device_location_[kv.first], kv.second, st);
        host_location_.clear();
        detail::key_to_device(zero_device_location_, zero_host_location_, st);
        storage_type_ = storage_type::DEVICE;
    }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    bool has_key(int elt) const { return device_location_.count(elt) || host_location_.count(elt); }
    detail::KeyView view(int elt, cudaStream_t st) const
    {
        static const DeviceVector<Data64> no_dev;
        static const PinnedVector<Data64> no_host;
        auto d = device_location_.find(elt);
        auto h = host_location_.find(elt);
        return detail::KeyView(d == device_location_.end() ? no_dev : d->second, h == host_location_.end() ? no_host : h->second, st);
    }
    detail::KeyView zero_view(cudaStream_t st) const { return detail::KeyView(zero_device_location_, zero_host_location_, st); }
};

template <Scheme S> class Plaintext;
// Plaintext<CKKS>: [L][N] words in the NTT domain (ckks/plaintext.cu); encoding is a client-side
// "next" row, so the caller supplies the encoded words.
template <> class Plaintext<Scheme::CKKS> : public detail::Storable {
  public:
    Plaintext() = default;
    void save(std::ostream& os) const;
    void load(std::istream& is);
    explicit Plaintext(HEContext<Scheme::CKKS> ctx, const ExecutionOptions& = ExecutionOptions()) : context_(ctx) {}
    Plaintext(HEContext<Scheme::CKKS> ctx, const std::vector<Data64>& words, int depth = 0, double scale = 1.0,
              const ExecutionOptions& opt = ExecutionOptions())
        : context_(ctx), depth_(depth), scale_(scale)
    {
        device_locations_ = DeviceVector<Data64>(words, opt.stream_);
        if (words.size() < (size_t) (ctx->Q_size - depth) * ctx->n)
            throw std::invalid_argument("Invalid Plaintext size!");
        plain_size_ = (int) words.size();
        plaintext_generated_ = true;
    }
    size_t size() const { return memory_size(); }
    int depth() const { return depth_; }
    double scale() const { return scale_; }
    HEContext<Scheme::CKKS> context_;
    int plain_size_ = 0, depth_ = 0;
    double scale_ = 0;
    bool in_ntt_domain_ = true, plaintext_generated_ = false;
};

template <Scheme S> class Switchkey;
template <> class Switchkey<Scheme::CKKS> {
  public:
    Switchkey() = default;
    explicit Switchkey(HEContext<Scheme::CKKS> ctx) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        ring_size = ctx->n, Q_prime_size_ = ctx->Q_prime_size, Q_size_ = ctx->Q_size, d_ = ctx->digit_count(0);
    }
    int ring_size = 0, Q_prime_size_ = 0, Q_size_ = 0, d_ = 0;
    void set_data(const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count(0) * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid switch key size!");
        device_location_ = DeviceVector<Data64>(words, st);
        switch_key_generated_ = true;
    }
    Data64* data() const { return device_location_.data(); }
    HEContext<Scheme::CKKS> context_;
    keyswitching_type key_type = keyswitching_type::NONE;
    storage_type storage_type_ = storage_type::DEVICE;
    DeviceVector<Data64> device_location_;
    PinnedVector<Data64> host_location_;
    // evaluationkey.cuh: store_in_host / store_in_device / is_on_device
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_host(device_location_, host_location_, st);
        storage_type_ = storage_type::HOST;
    }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_device(device_location_, host_location_, st);
        storage_type_ = storage_type::DEVICE;
    }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    detail::KeyView view(cudaStream_t st) const { return detail::KeyView(device_location_, host_location_, st); }
    bool switch_key_generated_ = false;
};

template <Scheme S> class HEEncoder; // client-side ("next" row); only named in the operator constructor
template <Scheme S> class HEOperator;

template <> class HEOperator<Scheme::CKKS> {
  protected:
    explicit HEOperator(HEContext<Scheme::CKKS> context)
    {
        if (!context || !context->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        context_ = std::move(context);
    }
    HEContext<Scheme::CKKS> context_;
    heon_context_t h() const { return context_->handle(); }
    size_t words(int comps, int depth) const { return (size_t) comps * (context_->Q_size - depth) * context_->n; }
    static void copy_meta(const Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& o)
    {
        o.context_ = a.context_;
        o.ring_size_ = a.ring_size_;
        o.coeff_modulus_count_ = a.coeff_modulus_count_;
        o.cipher_size_ = a.cipher_size_;
        o.depth_ = a.depth_;
        o.scale_ = a.scale_;
        o.in_ntt_domain_ = a.in_ntt_domain_;
        o.rescale_required_ = a.rescale_required_;
        o.relinearization_required_ = a.relinearization_required_;
        o.ciphertext_generated_ = true;
    }

  public:
    void add(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b, Ciphertext<Scheme::CKKS>& out,
             const ExecutionOptions& opt = ExecutionOptions())
    {
        binary(a, b, out, opt, 0);
    }
    void sub(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b, Ciphertext<Scheme::CKKS>& out,
             const ExecutionOptions& opt = ExecutionOptions())
    {
        binary(a, b, out, opt, 1);
    }
    void negate(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& out, const ExecutionOptions& opt = ExecutionOptions())
    {
        DeviceVector<Data64> mem(words(a.cipher_size_, a.depth_), opt.stream_);
        detail::check(heon_negate(h(), a.data(), 0, mem.data(), 0, a.cipher_size_, a.depth_, 1, opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        detail::output_storage(out, opt);
    }

    // in-place forms (operator.cuh:130-190, 260-300): the result replaces the first operand
    void add_inplace(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b, const ExecutionOptions& opt = ExecutionOptions())
    {
        add(a, b, a, opt);
    }
    void sub_inplace(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b, const ExecutionOptions& opt = ExecutionOptions())
    {
        sub(a, b, a, opt);
    }
    void negate_inplace(Ciphertext<Scheme::CKKS>& a, const ExecutionOptions& opt = ExecutionOptions()) { negate(a, a, opt); }

    // operator.cuh:631-691 + multiply_ckks (operator.cu:796-837)
    void multiply(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b, Ciphertext<Scheme::CKKS>& out,
                  const ExecutionOptions& opt = ExecutionOptions())
    {
        if (a.relinearization_required_ || b.relinearization_required_)
            throw std::invalid_argument("Ciphertexts can not be multiplied because of the non-linear part! Please use relinearization operation!");
        if (a.rescale_required_ || b.rescale_required_)
            throw std::invalid_argument("Ciphertexts can not be multiplied because of the noise! Please use rescale operation to get rid of additional noise!");
        if (a.depth_ != b.depth_)
            throw std::logic_error("Ciphertexts leveled are not equal");
        if (a.memory_size() < words(2, a.depth_) || b.memory_size() < words(2, a.depth_))
            throw std::invalid_argument("Invalid Ciphertexts size!");
        detail::InputGuard<Ciphertext<Scheme::CKKS>> ga(a, opt, &a == &out), gb(b, opt, &b == &out);
        DeviceVector<Data64> mem(words(3, a.depth_), opt.stream_);
        detail::check(heon_ckks_multiply(h(), a.data(), 0, b.data(), 0, mem.data(), 0, a.depth_, 1, opt.stream_));
        const double scale = a.scale_ * b.scale_;
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        out.scale_ = scale;
        out.cipher_size_ = 3;
        detail::output_storage(out, opt);
        out.relinearization_required_ = true;
        out.rescale_required_ = true;
    }
    void multiply_inplace(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b,
                          const ExecutionOptions& opt = ExecutionOptions())
    {
        multiply(a, b, a, opt);
    }

    // operator.cuh:1053-1094
    void relinearize_inplace(Ciphertext<Scheme::CKKS>& ct, Relinkey<Scheme::CKKS>& rk,
                             const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!ct.relinearization_required_)
            throw std::invalid_argument("Ciphertexts can not use relinearization, since no non-linear part!");
        if (!rk.relin_key_generated_)
            throw std::invalid_argument("Relinkey is not generated!");
        if (ct.memory_size() < words(3, ct.depth_))
            throw std::invalid_argument("Invalid Ciphertexts size!");
        detail::InputGuard<Ciphertext<Scheme::CKKS>> g(ct, opt, true);
        const detail::KeyView key = rk.view(opt.stream_);
        detail::check(heon_ckks_relinearize(h(), ct.data(), 0, key.ptr, ct.depth_, 1, opt.stream_));
        ct.relinearization_required_ = false;
        ct.cipher_size_ = 2;
        detail::output_storage(ct, opt);
    }

    // operator.cuh:1423-1445
    void rescale_inplace(Ciphertext<Scheme::CKKS>& ct, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (ct.depth_ >= context_->Q_size - 1)
            throw std::invalid_argument("Ciphertexts can not be rescaled, level is too low!");
        detail::InputGuard<Ciphertext<Scheme::CKKS>> g(ct, opt, true);
        detail::check(heon_ckks_rescale(h(), ct.data(), 0, ct.depth_, 1, opt.stream_));
        ct.scale_ /= (double) context_->prime_vector_[context_->Q_size - ct.depth_ - 1].value;
        ct.depth_++;
        ct.rescale_required_ = false;
        detail::output_storage(ct, opt);
    }

    // operator.cuh:1457-1600
    void mod_drop_inplace(Ciphertext<Scheme::CKKS>& ct, const ExecutionOptions& opt = ExecutionOptions())
    {
        detail::InputGuard<Ciphertext<Scheme::CKKS>> g(ct, opt, true);
        detail::check(heon_ckks_mod_drop_inplace(h(), ct.data(), 0, ct.cipher_size_, ct.depth_, 1, opt.stream_));
        ct.depth_++;
        detail::output_storage(ct, opt);
    }
    // mod_drop of a plaintext (operator.cuh:1512-1578): its first L-1 limbs are the plaintext one level down
    void mod_drop_inplace(Plaintext<Scheme::CKKS>& pt, const ExecutionOptions& = ExecutionOptions())
    {
        if (pt.depth_ >= context_->Q_size - 1)
            throw std::logic_error("Plaintext modulus can not be dropped!");
        pt.depth_++;
    }
    void mod_drop(Plaintext<Scheme::CKKS>& in, Plaintext<Scheme::CKKS>& out, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.depth_ >= context_->Q_size - 1)
            throw std::logic_error("Plaintext modulus can not be dropped!");
        detail::InputGuard<Plaintext<Scheme::CKKS>> g(in, opt, false);
        const size_t w = (size_t) (context_->Q_size - in.depth_ - 1) * context_->n;
        DeviceVector<Data64> mem(w, opt.stream_);
        detail::cuda(cudaMemcpyAsync(mem.data(), in.data(), w * sizeof(Data64), cudaMemcpyDeviceToDevice, opt.stream_));
        out.context_ = in.context_;
        out.memory_set(std::move(mem));
        out.plain_size_ = (int) w;
        out.depth_ = in.depth_ + 1;
        out.scale_ = in.scale_;
        out.in_ntt_domain_ = in.in_ntt_domain_;
        out.plaintext_generated_ = true;
        detail::output_storage(out, opt);
    }
    void mod_drop(Ciphertext<Scheme::CKKS>& in, Ciphertext<Scheme::CKKS>& out, const ExecutionOptions& opt = ExecutionOptions())
    {
        DeviceVector<Data64> mem(words(2, in.depth_ + 1), opt.stream_);
        detail::check(heon_ckks_mod_drop(h(), in.data(), 0, mem.data(), 0, in.depth_, 1, opt.stream_));
        copy_meta(in, out);
        out.memory_set(std::move(mem));
        out.depth_ = in.depth_ + 1;
    }

    // operator.cuh:1105-1270
    void apply_galois(Ciphertext<Scheme::CKKS>& in, Ciphertext<Scheme::CKKS>& out, Galoiskey<Scheme::CKKS>& gk,
                      int galois_elt, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.rescale_required_ || in.relinearization_required_)
            throw std::invalid_argument("Ciphertext can not be rotated because of the non-linear part or noise!");
        const detail::KeyView key = gk.view(galois_elt, opt.stream_);
        if (!key)
            throw std::logic_error("Galois key not present!");
        detail::InputGuard<Ciphertext<Scheme::CKKS>> g(in, opt, &in == &out);
        DeviceVector<Data64> mem(words(2, in.depth_), opt.stream_);
        detail::check(heon_ckks_apply_galois(h(), in.data(), 0, mem.data(), 0, key.ptr,
                                             (uint32_t) galois_elt, in.depth_, 1, opt.stream_));
        copy_meta(in, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = 2;
        detail::output_storage(out, opt);
    }
    void apply_galois_inplace(Ciphertext<Scheme::CKKS>& ct, Galoiskey<Scheme::CKKS>& gk, int galois_elt,
                              const ExecutionOptions& opt = ExecutionOptions())
    {
        apply_galois(ct, ct, gk, galois_elt, opt);
    }
    void rotate_rows(Ciphertext<Scheme::CKKS>& in, Ciphertext<Scheme::CKKS>& out, Galoiskey<Scheme::CKKS>& gk, int shift,
                     const ExecutionOptions& opt = ExecutionOptions())
    {
        if (shift == 0)
        {
            if (&in != &out)
            {
                DeviceVector<Data64> mem(words(2, in.depth_), opt.stream_);
                detail::cuda(cudaMemcpyAsync(mem.data(), in.data(), mem.size() * sizeof(Data64), cudaMemcpyDeviceToDevice, opt.stream_));
                copy_meta(in, out);
                out.memory_set(std::move(mem));
            }
            return;
        }
        const int elt = heon_steps_to_galois_elt(shift, context_->n, gk.group_order_);
        if (elt != 0 && gk.has_key(elt))
        {
            apply_galois(in, out, gk, elt, opt);
            return;
        }
        int log_slots = 0;
        while ((2 << log_slots) < context_->n)
            ++log_slots;
        Ciphertext<Scheme::CKKS>* cur = &in;
        for (int step : detail::rotation_plan(shift, log_slots, gk.max_shift_))
        {
            auto it = gk.galois_elt.find(step);
            if (it == gk.galois_elt.end() || !gk.has_key(it->second))
                throw std::logic_error("Galois key not present!");
            apply_galois(*cur, out, gk, it->second, opt);
            cur = &out;
        }
    }
    void rotate_rows_inplace(Ciphertext<Scheme::CKKS>& ct, Galoiskey<Scheme::CKKS>& gk, int shift,
                             const ExecutionOptions& opt = ExecutionOptions())
    {
        rotate_rows(ct, ct, gk, shift, opt);
    }

    // Baby-step loop of the BSGS matrix-vector product: every shift of `shifts` applied to the same
    // ciphertext (fast_single_hoisting_rotation_ckks_method_I/II, operator.cu:4674-5446; protected in the
    // reference, public here).  Mod-up and the forward NTTs are shared by all rotations; result r is
    // bit-identical to rotate_rows(in, out[r], gk, shifts[r]).
    std::vector<Ciphertext<Scheme::CKKS>> rotate_rows_hoisted(Ciphertext<Scheme::CKKS>& in, Galoiskey<Scheme::CKKS>& gk,
                                                              const std::vector<int>& shifts,
                                                              const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.rescale_required_ || in.relinearization_required_)
            throw std::invalid_argument("Ciphertext can not be rotated because of the non-linear part or noise!");
        std::vector<const uint64_t*> keys;
        std::vector<detail::KeyView> views;
        std::vector<uint32_t> elts;
        for (int s : shifts)
        {
            if (s == 0)
                throw std::invalid_argument("rotate_rows_hoisted: shift 0 is the identity, use the input itself");
            const int elt = heon_steps_to_galois_elt(s, context_->n, gk.group_order_);
            if (elt == 0 || !gk.has_key(elt))
                throw std::logic_error("Galois key not present!");
            views.emplace_back(gk.view(elt, opt.stream_));
            keys.push_back(views.back().ptr);
            elts.push_back((uint32_t) elt);
        }
        const size_t w = words(2, in.depth_);
        DeviceVector<Data64> mem(w * shifts.size(), opt.stream_);
        detail::check(heon_ckks_rotate_hoisted(h(), in.data(), 0, mem.data(), 0, (long long) w, keys.data(), elts.data(),
                                               (int) shifts.size(), in.depth_, 1, opt.stream_));
        std::vector<Ciphertext<Scheme::CKKS>> out(shifts.size());
        for (size_t r = 0; r < shifts.size(); ++r)
        {
            DeviceVector<Data64> one(w, opt.stream_);
            detail::cuda(cudaMemcpyAsync(one.data(), mem.data() + r * w, w * sizeof(Data64), cudaMemcpyDeviceToDevice,
                                         opt.stream_));
            copy_meta(in, out[r]);
            out[r].memory_set(std::move(one));
            out[r].cipher_size_ = 2;
        }
        return out;
    }

    // BSGS diagonal matrix-vector products with SINGLE hoisting: HEOperator<CKKS>::multiply_matrix
    // (operator.cu:2803-2895) and multiply_matrix_less_memory (:3398-3496), protected in the reference, public
    // here.  diags_matrices_bsgs_[m][0] = the baby-step shifts (hoisted: one decomposition for all of them),
    // diags_matrices_bsgs_[m][j][0] = the giant-step shift of group j (or the chain real_shift[m][j] of shifts
    // the Galois key holds); matrix[m]: [terms][L][N] diagonals over Q_l in the NTT domain, group order.
    // Each matrix is followed by scale *= scale_boot_ and rescale_inplace, as in the reference.
    double scale_boot_ = 1.0;
    void set_matrix_scale(double s) { scale_boot_ = s; }
    Ciphertext<Scheme::CKKS> multiply_matrix(Ciphertext<Scheme::CKKS>& cipher, std::vector<DeviceVector<Data64>>& matrix,
                                             std::vector<std::vector<std::vector<int>>>& diags_matrices_bsgs_,
                                             Galoiskey<Scheme::CKKS>& galois_key,
                                             const ExecutionOptions& opt = ExecutionOptions())
    {
        return matrix_single_hoisting(cipher, matrix, diags_matrices_bsgs_, nullptr, galois_key, opt);
    }
    Ciphertext<Scheme::CKKS> multiply_matrix_less_memory(Ciphertext<Scheme::CKKS>& cipher,
                                                         std::vector<DeviceVector<Data64>>& matrix,
                                                         std::vector<std::vector<std::vector<int>>>& diags_matrices_bsgs_,
                                                         std::vector<std::vector<std::vector<int>>>& real_shift,
                                                         Galoiskey<Scheme::CKKS>& galois_key,
                                                         const ExecutionOptions& opt = ExecutionOptions())
    {
        return matrix_single_hoisting(cipher, matrix, diags_matrices_bsgs_, &real_shift, galois_key, opt);
    }

  private:
    Ciphertext<Scheme::CKKS> matrix_single_hoisting(Ciphertext<Scheme::CKKS>& cipher,
                                                    std::vector<DeviceVector<Data64>>& matrix,
                                                    std::vector<std::vector<std::vector<int>>>& diags,
                                                    std::vector<std::vector<std::vector<int>>>* real_shift,
                                                    Galoiskey<Scheme::CKKS>& gk, const ExecutionOptions& opt)
    {
        detail::InputGuard<Ciphertext<Scheme::CKKS>> g(cipher, opt, false);
        Ciphertext<Scheme::CKKS> result;
        copy_meta(cipher, result);
        {
            const size_t w = words(2, cipher.depth_);
            DeviceVector<Data64> mem(w, opt.stream_);
            detail::cuda(cudaMemcpyAsync(mem.data(), cipher.data(), w * sizeof(Data64), cudaMemcpyDeviceToDevice, opt.stream_));
            result.memory_set(std::move(mem));
            result.cipher_size_ = 2;
        }
        for (int m = (int) diags.size() - 1; m > -1; m--)
        {
            const std::vector<int>& baby = diags[m][0];
            const int n1 = (int) baby.size();
            const int L = context_->Q_size - result.depth_;
            const size_t w = words(2, result.depth_);
            // fast_single_hoisting_rotation_ckks: every baby-step rotation from one decomposition
            DeviceVector<Data64> rotated(w * n1, opt.stream_);
            for (int i = 0; i < n1;)
            {
                if (baby[i] == 0)
                {
                    detail::cuda(cudaMemcpyAsync(rotated.data() + i * w, result.data(), w * sizeof(Data64),
                                                 cudaMemcpyDeviceToDevice, opt.stream_));
                    ++i;
                    continue;
                }
                std::vector<const uint64_t*> keys;
                std::vector<detail::KeyView> views;
                std::vector<uint32_t> elts;
                int e = i;
                for (; e < n1 && baby[e] != 0; ++e)
                {
                    const int elt = heon_steps_to_galois_elt(baby[e], context_->n, gk.group_order_);
                    if (!gk.has_key(elt))
                        throw std::logic_error("Galois key not present!");
                    views.emplace_back(gk.view(elt, opt.stream_));
                    keys.push_back(views.back().ptr);
                    elts.push_back((uint32_t) elt);
                }
                detail::check(heon_ckks_rotate_hoisted(h(), result.data(), 0, rotated.data() + i * w, 0, (long long) w,
                                                       keys.data(), elts.data(), e - i, result.depth_, 1, opt.stream_));
                i = e;
            }
            Ciphertext<Scheme::CKKS> acc;
            size_t counter = 0;
            for (size_t j = 0; j < diags[m].size(); ++j)
            {
                const int inner_n1 = (int) diags[m][j].size();
                if (inner_n1 > n1 || matrix[m].size() < (counter + inner_n1) * (size_t) L * context_->n)
                    throw std::invalid_argument("matrix diagonals: wrong size");
                Ciphertext<Scheme::CKKS> inner_sum;
                copy_meta(result, inner_sum);
                DeviceVector<Data64> mem(w, opt.stream_);
                detail::check(heon_ckks_multiply_plain_accumulate(h(), rotated.data(),
                                                                  matrix[m].data() + counter * (size_t) L * context_->n,
                                                                  mem.data(), inner_n1, result.depth_, opt.stream_));
                inner_sum.memory_set(std::move(mem));
                inner_sum.cipher_size_ = 2;
                counter += inner_n1;
                if (real_shift)
                    for (int sft : (*real_shift)[m][j])
                        rotate_rows_inplace(inner_sum, gk, sft, opt);
                else
                    rotate_rows_inplace(inner_sum, gk, diags[m][j][0], opt);
                if (j == 0)
                    acc = std::move(inner_sum);
                else
                    add(acc, inner_sum, acc, opt);
            }
            result = std::move(acc);
            result.scale_ = result.scale_ * scale_boot_;
            result.rescale_required_ = true;
            rescale_inplace(result, opt);
        }
        return result;
    }

  public:
    // BSGS diagonal matrix-vector products with double hoisting, the linear transforms of CKKS bootstrapping:
    // HEOperator<CKKS>::multiply_matrix_v2 (operator.cu:2898-3390; protected in the reference, public here) with
    // the reference's arguments.  matrix[m] holds the diagonals of matrix m encoded over PQ_l in the NTT domain,
    // [terms][L+K][N] in group order; diags_matrices_bsgs_[m][j][k] = rot_n1_[m][j] + (a shift of rot_n2_[m]).
    // The matrices are applied last to first, each followed by rescale_inplace (:3382-3387).
    Ciphertext<Scheme::CKKS> multiply_matrix_v2(Ciphertext<Scheme::CKKS>& cipher,
                                                std::vector<DeviceVector<Data64>>& matrix,
                                                std::vector<std::vector<std::vector<int>>>& diags_matrices_bsgs_,
                                                std::vector<std::vector<int>>& diags_matrices_bsgs_rot_n1_,
                                                std::vector<std::vector<int>>& diags_matrices_bsgs_rot_n2_,
                                                Galoiskey<Scheme::CKKS>& galois_key,
                                                const ExecutionOptions& opt = ExecutionOptions())
    {
        detail::InputGuard<Ciphertext<Scheme::CKKS>> g(cipher, opt, false);
        Ciphertext<Scheme::CKKS> result;
        copy_meta(cipher, result);
        {
            const size_t w = words(2, cipher.depth_);
            DeviceVector<Data64> mem(w, opt.stream_);
            detail::cuda(cudaMemcpyAsync(mem.data(), cipher.data(), w * sizeof(Data64), cudaMemcpyDeviceToDevice, opt.stream_));
            result.memory_set(std::move(mem));
            result.cipher_size_ = 2;
        }
        const int matrix_count = (int) diags_matrices_bsgs_.size();
        for (int m = matrix_count - 1; m > -1; m--)
        {
            std::sort(diags_matrices_bsgs_rot_n2_[m].begin(), diags_matrices_bsgs_rot_n2_[m].end());
            const std::vector<int>& baby = diags_matrices_bsgs_rot_n2_[m];
            std::vector<detail::KeyView> views; // keys staged from host memory live until the call is enqueued
            views.reserve(baby.size() + diags_matrices_bsgs_[m].size());
            auto resolve = [&](int shift, uint32_t& elt, const uint64_t*& key) {
                elt = 0;
                key = nullptr;
                if (shift == 0)
                    return;
                elt = (uint32_t) heon_steps_to_galois_elt(shift, context_->n, galois_key.group_order_);
                if (!galois_key.has_key((int) elt))
                    throw std::logic_error("Galois key not present!");
                views.emplace_back(galois_key.view((int) elt, opt.stream_));
                key = views.back().ptr;
            };
            std::vector<uint32_t> baby_elts(baby.size()), giant_elts(diags_matrices_bsgs_[m].size());
            std::vector<const uint64_t*> baby_keys(baby.size()), giant_keys(giant_elts.size());
            std::vector<int> sizes, terms;
            for (size_t i = 0; i < baby.size(); ++i)
                resolve(baby[i], baby_elts[i], baby_keys[i]);
            for (size_t j = 0; j < giant_elts.size(); ++j)
            {
                const int real_shift = diags_matrices_bsgs_rot_n1_[m][j];
                resolve(real_shift, giant_elts[j], giant_keys[j]);
                sizes.push_back((int) diags_matrices_bsgs_[m][j].size());
                for (int dg : diags_matrices_bsgs_[m][j])
                {
                    auto it = std::find(baby.begin(), baby.end(), dg - real_shift);
                    if (it == baby.end())
                        throw std::invalid_argument("BSGS diagonal without a baby step");
                    terms.push_back((int) std::distance(baby.begin(), it));
                }
            }
            const int L = context_->Q_size - result.depth_;
            if (matrix[m].size() < terms.size() * (size_t) (L + context_->P_size) * context_->n)
                throw std::invalid_argument("matrix diagonals: wrong size");
            DeviceVector<Data64> mem(words(2, result.depth_), opt.stream_);
            detail::check(heon_ckks_multiply_matrix(h(), result.data(), mem.data(), matrix[m].data(), baby_elts.data(),
                                                    baby_keys.data(), (int) baby.size(), giant_elts.data(), giant_keys.data(),
                                                    sizes.data(), terms.data(), (int) giant_elts.size(), result.depth_,
                                                    opt.stream_));
            result.memory_set(std::move(mem));
            result.in_ntt_domain_ = true;
            result.relinearization_required_ = false;
            result.scale_ = result.scale_ * (double) context_->prime_vector_[L].value;
            result.rescale_required_ = true;
            rescale_inplace(result, opt);
        }
        return result;
    }

    // operator.cuh:718-884 + multiply_plain_ckks (operator.cu:839-871)
    void multiply_plain(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p, Ciphertext<Scheme::CKKS>& out,
                        const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, out, opt, 0);
        out.scale_ = a.scale_ * p.scale_;
        out.rescale_required_ = true;
    }
    void multiply_plain_inplace(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p,
                                const ExecutionOptions& opt = ExecutionOptions())
    {
        multiply_plain(a, p, a, opt);
    }
    // operator.cuh:197-620 + add_plain_ckks / sub_plain_ckks (operator.cu:302-345, 434-477)
    void add_plain(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p, Ciphertext<Scheme::CKKS>& out,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, out, opt, 1);
    }
    void add_plain_inplace(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p, const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, a, opt, 1);
    }
    void sub_plain(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p, Ciphertext<Scheme::CKKS>& out,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, out, opt, 2);
    }
    void sub_plain_inplace(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p, const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, a, opt, 2);
    }

    // operator.cuh:1282-1340 + switchkey_ckks_method_I/II (operator.cu:1722-2025)
    void keyswitch(Ciphertext<Scheme::CKKS>& in, Ciphertext<Scheme::CKKS>& out, Switchkey<Scheme::CKKS>& sk,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.rescale_required_ || in.relinearization_required_)
            throw std::invalid_argument("Ciphertext can not be key-switched because of the non-linear part or noise!");
        if (!sk.switch_key_generated_)
            throw std::invalid_argument("Switchkey is not generated!");
        DeviceVector<Data64> mem(words(2, in.depth_), opt.stream_);
        const detail::KeyView key = sk.view(opt.stream_);
        detail::check(heon_ckks_keyswitch(h(), in.data(), 0, mem.data(), 0, key.ptr, in.depth_, 1, opt.stream_));
        copy_meta(in, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = 2;
    }
    void keyswitch_inplace(Ciphertext<Scheme::CKKS>& ct, Switchkey<Scheme::CKKS>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        keyswitch(ct, ct, sk, opt);
    }
    // operator.cuh:1354-1420 + conjugate_ckks_method_I/II (operator.cu:2027-2311)
    void conjugate(Ciphertext<Scheme::CKKS>& in, Ciphertext<Scheme::CKKS>& out, Galoiskey<Scheme::CKKS>& gk,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.rescale_required_ || in.relinearization_required_)
            throw std::invalid_argument("Ciphertext can not be conjugated because of the non-linear part or noise!");
        const detail::KeyView key = gk.zero_view(opt.stream_);
        if (!key)
            throw std::logic_error("Conjugation key not present!");
        DeviceVector<Data64> mem(words(2, in.depth_), opt.stream_);
        detail::check(heon_ckks_conjugate(h(), in.data(), 0, mem.data(), 0, key.ptr, in.depth_, 1, opt.stream_));
        copy_meta(in, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = 2;
    }

    void conjugate_inplace(Ciphertext<Scheme::CKKS>& ct, Galoiskey<Scheme::CKKS>& gk, const ExecutionOptions& opt = ExecutionOptions())
    {
        conjugate(ct, ct, gk, opt);
    }

  private:
    void plain(Ciphertext<Scheme::CKKS>& a, Plaintext<Scheme::CKKS>& p, Ciphertext<Scheme::CKKS>& out,
               const ExecutionOptions& opt, int op)
    {
        detail::InputGuard<Ciphertext<Scheme::CKKS>> ga(a, opt, &a == &out);
        detail::InputGuard<Plaintext<Scheme::CKKS>> gp(p, opt, false);
        if (a.depth_ != p.depth_)
            throw std::logic_error("Ciphertexts leveled are not equal");
        if (a.memory_size() < words(a.cipher_size_, a.depth_))
            throw std::invalid_argument("Invalid Ciphertexts size!");
        if (p.size() < words(1, a.depth_))
            throw std::invalid_argument("Invalid Plaintext size!");
        DeviceVector<Data64> mem(words(a.cipher_size_, a.depth_), opt.stream_);
        auto fn = op == 0 ? heon_ckks_multiply_plain : op == 1 ? heon_ckks_add_plain : heon_ckks_sub_plain;
        detail::check(fn(h(), a.data(), 0, p.data(), 0, mem.data(), 0, a.cipher_size_, a.depth_, 1, opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        detail::output_storage(out, opt);
    }
    void binary(Ciphertext<Scheme::CKKS>& a, Ciphertext<Scheme::CKKS>& b, Ciphertext<Scheme::CKKS>& out,
                const ExecutionOptions& opt, int op)
    {
        detail::InputGuard<Ciphertext<Scheme::CKKS>> ga(a, opt, &a == &out), gb(b, opt, &b == &out);
        if (a.depth_ != b.depth_)
            throw std::logic_error("Ciphertexts leveled are not equal");
        if (a.cipher_size_ != b.cipher_size_)
            throw std::invalid_argument("Ciphertexts should have the same size!");
        DeviceVector<Data64> mem(words(a.cipher_size_, a.depth_), opt.stream_);
        detail::check((op == 0 ? heon_add : heon_sub)(h(), a.data(), 0, b.data(), 0, mem.data(), 0, a.cipher_size_,
                                                      a.depth_, 1, opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        detail::output_storage(out, opt);
    }
};

template <Scheme S> class HEArithmeticOperator;
template <> class HEArithmeticOperator<Scheme::CKKS> : public HEOperator<Scheme::CKKS> {
  public:
    explicit HEArithmeticOperator(HEContext<Scheme::CKKS> context) : HEOperator<Scheme::CKKS>(context) {}
    // reference signature (ckks/operator.cu:6680); the encoder is not used by the hot path
    HEArithmeticOperator(HEContext<Scheme::CKKS> context, HEEncoder<Scheme::CKKS>&) : HEOperator<Scheme::CKKS>(context) {}
};

// ---------------------------------------------------------------------------
// BFV twins (src/include/heongpu/host/bfv/*.cuh).  Ciphertexts live in the
// COEFFICIENT domain, there are no levels (depth 0 everywhere on this path).
// ---------------------------------------------------------------------------
template <> class HEContextImpl<Scheme::BFV> {
  public:
    explicit HEContextImpl(sec_level_type sec = sec_level_type::sec128, int device = 0) : sec_level_(sec), device_(device) {}
    ~HEContextImpl()
    {
        if (h_)
            heon_context_destroy(h_);
    }
    void set_poly_modulus_degree(size_t n)
    {
        if (coeff_modulus_specified_ || poly_modulus_degree_specified_)
            throw std::logic_error("Poly modulus degree cannot be changed after the coeff_modulus is specified!");
        if (n == 0 || (n & (n - 1)))
            throw std::logic_error("Poly modulus degree have to be power of two");
        if (n > 65536 || n < 4096)
            throw std::logic_error("Poly modulus degree is not supported");
        this->n = (int) n;
        n_power = 0;
        while ((size_t(1) << n_power) < n)
            ++n_power;
        poly_modulus_degree_specified_ = true;
    }
    void set_coeff_modulus_bit_sizes(const std::vector<int>& q_bits, const std::vector<int>& p_bits)
    {
        if (coeff_modulus_specified_ || context_generated_ || !poly_modulus_degree_specified_)
            throw std::logic_error("Coeff_modulus cannot be changed after the context is generated!");
        if (p_bits.empty())
            throw std::logic_error("log_P_bases_bit_sizes cannot be empty!");
        int total = 0;
        for (int b : q_bits)
            total += b;
        for (int b : p_bits)
            total += b;
        detail::check_security(sec_level_, (size_t) n, total);
        q_bits_ = q_bits;
        p_bits_ = p_bits;
        coeff_modulus_specified_ = true;
    }
    // bfv/context.cu:135-220
    void set_coeff_modulus_values(const std::vector<Data64>& q, const std::vector<Data64>& p)
    {
        if (coeff_modulus_specified_ || context_generated_ || !poly_modulus_degree_specified_)
            throw std::logic_error("Coeff_modulus cannot be changed after the context is generated!");
        if (p.empty())
            throw std::logic_error("log_P_bases_bit_sizes cannot be empty!");
        int total = 0;
        for (Data64 v : q)
            total += detail::bit_size(v);
        for (Data64 v : p)
            total += detail::bit_size(v);
        detail::check_security(sec_level_, (size_t) n, total);
        q_vals_ = q;
        p_vals_ = p;
        by_value_ = true;
        coeff_modulus_specified_ = true;
    }
    // bfv/context.cu:223-300: the last P_modulus_size primes of the default chain are the special primes
    void set_coeff_modulus_default_values(int P_modulus_size)
    {
        if (coeff_modulus_specified_ || context_generated_ || !poly_modulus_degree_specified_)
            throw std::logic_error("Coeff_modulus cannot be changed after the context is generated!");
        if (sec_level_ != sec_level_type::sec128)
            throw std::runtime_error("default moduli are tabulated for 128-bit security only");
        const std::vector<Data64> all = detail::default_modulus_128((size_t) n);
        if (P_modulus_size < 1 || (size_t) P_modulus_size >= all.size())
            throw std::logic_error("P_modulus_size is not valid!");
        q_vals_.assign(all.begin(), all.end() - P_modulus_size);
        p_vals_.assign(all.end() - P_modulus_size, all.end());
        by_value_ = true;
        coeff_modulus_specified_ = true;
    }
    void set_plain_modulus(int t)
    {
        if (context_generated_)
            throw std::logic_error("Plain modulus cannot be changed after the context is generated!");
        plain_modulus_ = (Data64) t;
        plain_modulus_specified_ = true;
    }
    // generate(const MemoryPoolConfig&) (context.cu: the pool is configured before the tables are built)
    void generate(const MemoryPoolConfig& pool_config)
    {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(device_);
        MemoryPool::instance().initialize(pool_config);
        cudaSetDevice(prev);
        generate();
    }
    void generate()
    {
        if (context_generated_ || !poly_modulus_degree_specified_ || !coeff_modulus_specified_ || !plain_modulus_specified_)
            throw std::runtime_error("Context is already generated or not fully specified!");
        detail::require_current_device(device_);
        if (by_value_)
            detail::check(heon_bfv_context_create_values(device_, n_power, q_vals_.data(), (int) q_vals_.size(), p_vals_.data(),
                                                         (int) p_vals_.size(), plain_modulus_, &h_));
        else
            detail::check(heon_bfv_context_create(device_, n_power, q_bits_.data(), (int) q_bits_.size(), p_bits_.data(),
                                                  (int) p_bits_.size(), plain_modulus_, &h_));
        heon_info info;
        detail::check(heon_context_info(h_, &info));
        Q_size = info.q_size;
        P_size = info.p_size;
        Q_prime_size = Q_size + P_size;
        keyswitching_type_ = info.keyswitch_method == 1 ? keyswitching_type::KEYSWITCHING_METHOD_I
                                                        : keyswitching_type::KEYSWITCHING_METHOD_II;
        size_t cnt = 0;
        detail::check(heon_context_table(h_, HEON_TBL_MODULUS, 0, nullptr, 0, &cnt));
        std::vector<Data64> raw(cnt);
        detail::check(heon_context_table(h_, HEON_TBL_MODULUS, 0, raw.data(), cnt, &cnt));
        prime_vector_.clear();
        for (size_t i = 0; i < cnt && prime_vector_.size() < (size_t) Q_prime_size; i += 3)
            prime_vector_.push_back(Modulus64{raw[i], raw[i + 1], raw[i + 2]});
        context_generated_ = true;
    }
    // bfv/context.cu:803-930: the reference serialises the parameters chosen at set_coeff_modulus time (the primes
    // included) and load() regenerates the context from them.  This class picks its primes inside generate(), so an
    // ungenerated context asks a scratch context for them first.
    void save(std::ostream& os) const
    {
        if (!poly_modulus_degree_specified_ || !coeff_modulus_specified_ || !plain_modulus_specified_)
            throw std::runtime_error("Context has no enough parameters to serialize!");
        std::vector<Modulus64> primes = prime_vector_;
        int q_size = Q_size, p_size = P_size;
        uint8_t ks = (uint8_t) keyswitching_type_;
        if (!context_generated_)
        {
            HEContextImpl scratch(sec_level_type::none, device_);
            scratch.set_poly_modulus_degree((size_t) n);
            if (by_value_)
                scratch.set_coeff_modulus_values(q_vals_, p_vals_);
            else
                scratch.set_coeff_modulus_bit_sizes(q_bits_, p_bits_);
            scratch.set_plain_modulus((int) plain_modulus_);
            scratch.generate();
            primes = scratch.prime_vector_;
            q_size = scratch.Q_size;
            p_size = scratch.P_size;
            ks = (uint8_t) scratch.keyswitching_type_;
        }
        auto put = [&](const auto& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(v)); };
        put((uint8_t) 0x1); // scheme_type::bfv
        put((uint8_t) sec_level_);
        put(ks);
        put((int) n);
        put((int) n_power);
        put((int) (q_size + p_size)); // coeff_modulus
        int total_bits = 0;
        for (const auto& m : primes)
            total_bits += (int) m.bit;
        put(total_bits);
        put((int) (q_size + p_size));
        put(q_size);
        put(p_size);
        put((uint32_t) primes.size());
        os.write(reinterpret_cast<const char*>(primes.data()), (std::streamsize) (sizeof(Modulus64) * primes.size()));
        put((uint32_t) q_size); // base_q
        for (int i = 0; i < q_size; ++i)
            put((Data64) primes[i].value);
        auto put_bits = [&](int from, int to) {
            put((uint32_t) (to - from));
            for (int i = from; i < to; ++i)
                put((int) primes[i].bit);
        };
        put_bits(0, q_size + p_size);
        put_bits(0, q_size);
        put_bits(q_size, q_size + p_size);
        const Modulus64 t(plain_modulus_);
        put(t);
    }
    void load(std::istream& is)
    {
        if (context_generated_)
            throw std::runtime_error("Context has been already exist!");
        auto get = [&](auto& v) {
            is.read(reinterpret_cast<char*>(&v), sizeof(v));
            if (!is)
                throw std::runtime_error("Invalid context binary!");
        };
        uint8_t scheme, sec, ks;
        int coeff_modulus, total_bits, qp, q, pz;
        get(scheme);
        if (scheme != 0x1)
            throw std::runtime_error("Invalid scheme binary!");
        get(sec);
        get(ks);
        get(n);
        get(n_power);
        get(coeff_modulus);
        get(total_bits);
        get(qp);
        get(q);
        get(pz);
        uint32_t cnt;
        get(cnt);
        std::vector<Modulus64> primes(cnt);
        is.read(reinterpret_cast<char*>(primes.data()), (std::streamsize) (sizeof(Modulus64) * cnt));
        for (int v = 0; v < 4; ++v) // base_q and the three bit-size vectors
        {
            get(cnt);
            is.ignore((std::streamsize) cnt * (v == 0 ? sizeof(Data64) : sizeof(int)));
        }
        Modulus64 t;
        get(t);
        if (!is || q + pz != (int) primes.size() || q < 1 || pz < 1)
            throw std::runtime_error("Invalid context binary!");
        plain_modulus_ = t.value;
        plain_modulus_specified_ = true;
        sec_level_ = (sec_level_type) sec;
        q_vals_.clear();
        p_vals_.clear();
        for (int i = 0; i < q; ++i)
            q_vals_.push_back(primes[i].value);
        for (int i = q; i < q + pz; ++i)
            p_vals_.push_back(primes[i].value);
        by_value_ = true;
        poly_modulus_degree_specified_ = true;
        coeff_modulus_specified_ = true;
        generate();
    }
    // Method II digits have size 2 in the reference's BFV context whatever |P| is (contextpool.hpp:29)
    int digit_count() const { return P_size == 1 ? Q_size : (Q_size + 1) / 2; }
    heon_context_t handle() const { return h_; }
    size_t get_poly_modulus_degree() const { return (size_t) n; }
    int get_log_poly_modulus_degree() const { return n_power; }
    int get_ciphertext_modulus_count() const { return Q_size; }
    int get_key_modulus_count() const { return Q_prime_size; }
    std::vector<Modulus64> get_key_modulus() const { return prime_vector_; }
    // ckks/context.cu: print_parameters
    void print_parameters() const
    {
        if (!context_generated_)
        {
            std::cout << "Parameters is not generated yet!" << std::endl;
            return;
        }
        std::cout << "==== HEonGPU a GPU Based Homomorphic Encryption Library ====\n" << std::endl;
        std::cout << "Encryption parameters:" << std::endl;
        std::cout << "-->   scheme: " << "BFV" << std::endl;
        std::cout << "-->   poly_modulus_degree: " << n << std::endl;
        std::cout << "-->   Q_tilta size: Q( ";
        for (int i = 0; i < Q_size; ++i)
            std::cout << prime_vector_[i].bit << (i + 1 < Q_size ? " + " : "");
        std::cout << " ) + P( ";
        for (int i = Q_size; i < Q_prime_size; ++i)
            std::cout << prime_vector_[i].bit << (i + 1 < Q_prime_size ? " + " : "");
        std::cout << " ) bits" << std::endl;
        std::cout << "-->   plain_modulus: " << plain_modulus_ << std::endl;
        std::cout << std::endl;
    }
    Modulus64 get_plain_modulus() const { return Modulus64(plain_modulus_); }

    int n = 0, n_power = 0;
    int Q_size = 0, P_size = 0, Q_prime_size = 0;
    Data64 plain_modulus_ = 0;
    keyswitching_type keyswitching_type_ = keyswitching_type::NONE;
    std::vector<Modulus64> prime_vector_;
    bool context_generated_ = false;

  private:
    sec_level_type sec_level_;
    int device_;
    heon_context_t h_ = nullptr;
    std::vector<int> q_bits_, p_bits_;
    std::vector<Data64> q_vals_, p_vals_;
    bool by_value_ = false;
    bool poly_modulus_degree_specified_ = false, coeff_modulus_specified_ = false, plain_modulus_specified_ = false;
};

template <> class Ciphertext<Scheme::BFV> : public detail::Storable {
  public:
    Ciphertext() = default;
    void save(std::ostream& os) const;
    void load(std::istream& is);
    explicit Ciphertext(HEContext<Scheme::BFV> ctx, const ExecutionOptions& = ExecutionOptions()) : context_(ctx)
    {
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_size;
        cipher_size_ = 2;
    }
    // [cipher_size][Q][N] words, coefficient domain (bfv/ciphertext.cu)
    Ciphertext(HEContext<Scheme::BFV> ctx, const std::vector<Data64>& words, int cipher_size = 2,
               const ExecutionOptions& opt = ExecutionOptions())
        : context_(ctx), cipher_size_(cipher_size)
    {
        device_locations_ = DeviceVector<Data64>(words, opt.stream_);
        ring_size_ = ctx->n;
        coeff_modulus_count_ = ctx->Q_size;
        if (words.size() < (size_t) cipher_size * ctx->Q_size * ctx->n)
            throw std::invalid_argument("Invalid Ciphertexts size!");
        ciphertext_generated_ = true;
    }
    void get_data(std::vector<Data64>& out, cudaStream_t st = cudaStreamDefault) const
    {
        out.resize((size_t) cipher_size_ * coeff_modulus_count_ * ring_size_);
        if (!is_on_device())
        {
            std::copy(host_data(), host_data() + out.size(), out.begin());
            return;
        }
        detail::cuda(cudaMemcpyAsync(out.data(), data(), out.size() * sizeof(Data64), cudaMemcpyDeviceToHost, st));
        detail::cuda(cudaStreamSynchronize(st));
    }
    int size() const { return cipher_size_; }
    bool in_ntt_domain() const { return in_ntt_domain_; }
    bool relinearization_required() const { return relinearization_required_; }

    HEContext<Scheme::BFV> context_;
    int ring_size_ = 0, coeff_modulus_count_ = 0, cipher_size_ = 0;
    bool in_ntt_domain_ = false, relinearization_required_ = false, ciphertext_generated_ = false;
};

// Plaintext<BFV>: [N] values below the plain modulus (bfv/plaintext.cu); batching encode is a client-side row.
template <> class Plaintext<Scheme::BFV> : public detail::Storable {
  public:
    Plaintext() = default;
    void save(std::ostream& os) const;
    void load(std::istream& is);
    explicit Plaintext(HEContext<Scheme::BFV> ctx, const ExecutionOptions& = ExecutionOptions()) : context_(ctx) {}
    Plaintext(HEContext<Scheme::BFV> ctx, const std::vector<Data64>& words, const ExecutionOptions& opt = ExecutionOptions())
        : context_(ctx)
    {
        device_locations_ = DeviceVector<Data64>(words, opt.stream_);
        if (words.size() < (size_t) ctx->n)
            throw std::invalid_argument("Invalid Plaintext size!");
        plain_size_ = (int) words.size();
        plaintext_generated_ = true;
    }
    size_t size() const { return memory_size(); }
    HEContext<Scheme::BFV> context_;
    int plain_size_ = 0;
    bool in_ntt_domain_ = false, plaintext_generated_ = false;
};

template <> class Relinkey<Scheme::BFV> {
  public:
    Relinkey() = default; // filled by load() (serializer::deserialize / load_from_file)
    explicit Relinkey(HEContext<Scheme::BFV> ctx) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        ring_size = ctx->n, Q_prime_size_ = ctx->Q_prime_size, Q_size_ = ctx->Q_size, d_ = ctx->digit_count();
    }
    int ring_size = 0, Q_prime_size_ = 0, Q_size_ = 0, d_ = 0;
    void save(std::ostream& os) const;
    void load(std::istream& is);
    void set_data(const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count() * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid relinearization key size!");
        device_location_ = DeviceVector<Data64>(words, st);
        relin_key_generated_ = true;
    }
    Data64* data() const { return device_location_.data(); }
    HEContext<Scheme::BFV> context_;
    keyswitching_type key_type = keyswitching_type::NONE;
    storage_type storage_type_ = storage_type::DEVICE;
    DeviceVector<Data64> device_location_;
    PinnedVector<Data64> host_location_;
    // evaluationkey.cuh: store_in_host / store_in_device / is_on_device
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_host(device_location_, host_location_, st);
        storage_type_ = storage_type::HOST;
    }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_device(device_location_, host_location_, st);
        storage_type_ = storage_type::DEVICE;
    }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    detail::KeyView view(cudaStream_t st) const { return detail::KeyView(device_location_, host_location_, st); }
    bool relin_key_generated_ = false;
};

template <> class Switchkey<Scheme::BFV> {
  public:
    Switchkey() = default;
    explicit Switchkey(HEContext<Scheme::BFV> ctx) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        ring_size = ctx->n, Q_prime_size_ = ctx->Q_prime_size, Q_size_ = ctx->Q_size, d_ = ctx->digit_count();
    }
    int ring_size = 0, Q_prime_size_ = 0, Q_size_ = 0, d_ = 0;
    void set_data(const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count() * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid switch key size!");
        device_location_ = DeviceVector<Data64>(words, st);
        switch_key_generated_ = true;
    }
    Data64* data() const { return device_location_.data(); }
    HEContext<Scheme::BFV> context_;
    keyswitching_type key_type = keyswitching_type::NONE;
    storage_type storage_type_ = storage_type::DEVICE;
    DeviceVector<Data64> device_location_;
    PinnedVector<Data64> host_location_;
    // evaluationkey.cuh: store_in_host / store_in_device / is_on_device
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_host(device_location_, host_location_, st);
        storage_type_ = storage_type::HOST;
    }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        detail::key_to_device(device_location_, host_location_, st);
        storage_type_ = storage_type::DEVICE;
    }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    detail::KeyView view(cudaStream_t st) const { return detail::KeyView(device_location_, host_location_, st); }
    bool switch_key_generated_ = false;
};

template <> class Galoiskey<Scheme::BFV> {
  public:
    Galoiskey() = default; // filled by load()
    int ring_size = 0, Q_prime_size_ = 0, Q_size_ = 0, d_ = 0;
    void bind_meta(const HEContextImpl<Scheme::BFV>& c_) { ring_size = c_.n, Q_prime_size_ = c_.Q_prime_size, Q_size_ = c_.Q_size, d_ = c_.digit_count(); }
    explicit Galoiskey(HEContext<Scheme::BFV> ctx) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        bind_meta(*ctx);
        for (int i = 0; i < 8; ++i) // default keys for +-2^i, i < MAX_SHIFT (bfv/evaluationkey.cu:306-345)
        {
            galois_elt[1 << i] = heon_steps_to_galois_elt(1 << i, ctx->n, group_order_);
            galois_elt[-(1 << i)] = heon_steps_to_galois_elt(-(1 << i), ctx->n, group_order_);
        }
        galois_elt_zero = 2 * ctx->n - 1;
    }
    Galoiskey(HEContext<Scheme::BFV> ctx, const std::vector<int>& shifts) : context_(ctx), key_type(ctx->keyswitching_type_)
    {
        bind_meta(*ctx);
        customized = true;
        for (int s : shifts)
            galois_elt[s] = heon_steps_to_galois_elt(s, ctx->n, group_order_);
        galois_elt_zero = 2 * ctx->n - 1;
    }
    void set_key(int galois_element, const std::vector<Data64>& words, cudaStream_t st = cudaStreamDefault)
    {
        const size_t need = (size_t) context_->digit_count() * 2 * context_->Q_prime_size * context_->n;
        if (words.size() != need)
            throw std::invalid_argument("Invalid galois key size!");
        device_location_[galois_element] = DeviceVector<Data64>(words, st);
    }
    Galoiskey(HEContext<Scheme::BFV> ctx, const std::vector<uint32_t>& elts)
        : context_(ctx), key_type(ctx->keyswitching_type_), custom_galois_elt(elts)
    {
        bind_meta(*ctx);
        customized = true;
        galois_elt_zero = 2 * ctx->n - 1;
    }
    // the column-rotation key (element 2N-1) lives in the same map (bfv/evaluationkey.cu)
    void set_zero_key(int elt, DeviceVector<Data64>&& key)
    {
        device_location_[elt] = std::move(key);
        galois_elt_zero = elt;
    }
    const Data64* zero_key_data() const
    {
        auto it = device_location_.find(galois_elt_zero);
        if (it == device_location_.end())
            throw std::logic_error("Galois key not present!");
        return it->second.data();
    }
    void save(std::ostream& os) const;
    void load(std::istream& is);
    std::vector<uint32_t> custom_galois_elt;
    bool galois_key_generated_ = false;
    int max_shift_ = 7; // MAX_SHIFT - 1
    HEContext<Scheme::BFV> context_;
    keyswitching_type key_type = keyswitching_type::NONE;
    storage_type storage_type_ = storage_type::DEVICE;
    int group_order_ = 3;
    bool customized = false;
    int galois_elt_zero = 0; // column rotation
    std::unordered_map<int, int> galois_elt;
    std::unordered_map<int, DeviceVector<Data64>> device_location_;
    std::unordered_map<int, PinnedVector<Data64>> host_location_;
    void store_in_host(cudaStream_t st = cudaStreamDefault)
    {
        for (auto& kv : device_location_)
            detail::key_to_host(kv.second, host_location_[kv.first], st);
        device_location_.clear();
        storage_type_ = storage_type::HOST;
    }
    void store_in_device(cudaStream_t st = cudaStreamDefault)
    {
        for (auto& kv : host_location_)
            detail::key_to_device(device_location_[kv.first], kv.second, st);
        host_location_.clear();
        storage_type_ = storage_type::DEVICE;
    }
    bool is_on_device() const { return storage_type_ == storage_type::DEVICE; }
    bool has_key(int elt) const { return device_location_.count(elt) || host_location_.count(elt); }
    detail::KeyView view(int elt, cudaStream_t st) const
    {
        static const DeviceVector<Data64> no_dev;
        static const PinnedVector<Data64> no_host;
        auto d = device_location_.find(elt);
        auto h = host_location_.find(elt);
        return detail::KeyView(d == device_location_.end() ? no_dev : d->second, h == host_location_.end() ? no_host : h->second, st);
    }
    detail::KeyView zero_view(cudaStream_t st) const { return view(galois_elt_zero, st); }
};

template <> class HEOperator<Scheme::BFV> {
  protected:
    explicit HEOperator(HEContext<Scheme::BFV> context)
    {
        if (!context || !context->context_generated_)
            throw std::invalid_argument("HEContext is not generated!");
        context_ = std::move(context);
    }
    HEContext<Scheme::BFV> context_;
    heon_context_t h() const { return context_->handle(); }
    size_t words(int comps) const { return (size_t) comps * context_->Q_size * context_->n; }
    static void copy_meta(const Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& o)
    {
        o.context_ = a.context_;
        o.ring_size_ = a.ring_size_;
        o.coeff_modulus_count_ = a.coeff_modulus_count_;
        o.cipher_size_ = a.cipher_size_;
        o.in_ntt_domain_ = a.in_ntt_domain_;
        o.relinearization_required_ = a.relinearization_required_;
        o.ciphertext_generated_ = true;
    }

  public:
    void add(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, Ciphertext<Scheme::BFV>& out,
             const ExecutionOptions& opt = ExecutionOptions())
    {
        binary(a, b, out, opt, 0);
    }
    void sub(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, Ciphertext<Scheme::BFV>& out,
             const ExecutionOptions& opt = ExecutionOptions())
    {
        binary(a, b, out, opt, 1);
    }
    void negate(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& out, const ExecutionOptions& opt = ExecutionOptions())
    {
        DeviceVector<Data64> mem(words(a.cipher_size_), opt.stream_);
        detail::check(heon_negate(h(), a.data(), 0, mem.data(), 0, a.cipher_size_, 0, 1, opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        detail::output_storage(out, opt);
    }
    void add_inplace(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, const ExecutionOptions& opt = ExecutionOptions())
    {
        add(a, b, a, opt);
    }
    void sub_inplace(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, const ExecutionOptions& opt = ExecutionOptions())
    {
        sub(a, b, a, opt);
    }
    void negate_inplace(Ciphertext<Scheme::BFV>& a, const ExecutionOptions& opt = ExecutionOptions()) { negate(a, a, opt); }
    // bfv/operator.cuh multiply + multiply_bfv (bfv/operator.cu:336-430)
    void multiply(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, Ciphertext<Scheme::BFV>& out,
                  const ExecutionOptions& opt = ExecutionOptions())
    {
        if (a.relinearization_required_ || b.relinearization_required_)
            throw std::invalid_argument("Ciphertexts can not be multiplied because of the non-linear part! Please use relinearization operation!");
        if (a.in_ntt_domain_ || b.in_ntt_domain_)
            throw std::invalid_argument("Ciphertexts should be in the coefficient domain!");
        if (a.memory_size() < words(2) || b.memory_size() < words(2))
            throw std::invalid_argument("Invalid Ciphertexts size!");
        detail::InputGuard<Ciphertext<Scheme::BFV>> ga(a, opt, &a == &out), gb(b, opt, &b == &out);
        DeviceVector<Data64> mem(words(3), opt.stream_);
        detail::check(heon_bfv_multiply(h(), a.data(), 0, b.data(), 0, mem.data(), 0, 1, opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = 3;
        out.relinearization_required_ = true;
        detail::output_storage(out, opt);
    }
    void multiply_inplace(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, const ExecutionOptions& opt = ExecutionOptions())
    {
        multiply(a, b, a, opt);
    }
    // relinearize_seal_method_inplace / relinearize_external_product_method2_inplace (bfv/operator.cu:505-671)
    void relinearize_inplace(Ciphertext<Scheme::BFV>& ct, Relinkey<Scheme::BFV>& rk, const ExecutionOptions& opt = ExecutionOptions())
    {
        if (!ct.relinearization_required_)
            throw std::invalid_argument("Ciphertexts can not use relinearization, since no non-linear part!");
        if (!rk.relin_key_generated_)
            throw std::invalid_argument("Relinkey is not generated!");
        if (ct.memory_size() < words(3))
            throw std::invalid_argument("Invalid Ciphertexts size!");
        detail::InputGuard<Ciphertext<Scheme::BFV>> g(ct, opt, true);
        const detail::KeyView key = rk.view(opt.stream_);
        detail::check(heon_bfv_relinearize(h(), ct.data(), 0, key.ptr, 1, opt.stream_));
        ct.relinearization_required_ = false;
        ct.cipher_size_ = 2;
        detail::output_storage(ct, opt);
    }
    // apply_galois_method_I/II, rotate_rows, rotate_columns (bfv/operator.cu:771-973)
    void apply_galois(Ciphertext<Scheme::BFV>& in, Ciphertext<Scheme::BFV>& out, Galoiskey<Scheme::BFV>& gk, int galois_elt,
                      const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.relinearization_required_)
            throw std::invalid_argument("Ciphertext can not be rotated because of the non-linear part!");
        const detail::KeyView key = gk.view(galois_elt, opt.stream_);
        if (!key)
            throw std::logic_error("Galois key not present!");
        detail::InputGuard<Ciphertext<Scheme::BFV>> g(in, opt, &in == &out);
        DeviceVector<Data64> mem(words(2), opt.stream_);
        detail::check(heon_bfv_apply_galois(h(), in.data(), 0, mem.data(), 0, key.ptr, (uint32_t) galois_elt, 1,
                                            opt.stream_));
        copy_meta(in, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = 2;
        detail::output_storage(out, opt);
    }
    void rotate_rows(Ciphertext<Scheme::BFV>& in, Ciphertext<Scheme::BFV>& out, Galoiskey<Scheme::BFV>& gk, int shift,
                     const ExecutionOptions& opt = ExecutionOptions())
    {
        if (shift == 0)
        {
            // the reference returns the input unchanged (bfv/operator.cuh:591-595); steps_to_galois_elt(0) is the
            // column-rotation element 2N-1 and must not be applied here
            if (&in != &out)
            {
                DeviceVector<Data64> mem(words(2), opt.stream_);
                detail::cuda(cudaMemcpyAsync(mem.data(), in.data(), mem.size() * sizeof(Data64), cudaMemcpyDeviceToDevice, opt.stream_));
                copy_meta(in, out);
                out.memory_set(std::move(mem));
            }
            return;
        }
        const int elt = heon_steps_to_galois_elt(shift, context_->n, gk.group_order_);
        if (elt != 0 && gk.has_key(elt))
        {
            apply_galois(in, out, gk, elt, opt);
            return;
        }
        int log_slots = 0;
        while ((2 << log_slots) < context_->n)
            ++log_slots;
        Ciphertext<Scheme::BFV>* cur = &in;
        for (int step : detail::rotation_plan(shift, log_slots, gk.max_shift_))
        {
            auto it = gk.galois_elt.find(step);
            if (it == gk.galois_elt.end() || !gk.has_key(it->second))
                throw std::logic_error("Galois key not present!");
            apply_galois(*cur, out, gk, it->second, opt);
            cur = &out;
        }
    }
    void rotate_rows_inplace(Ciphertext<Scheme::BFV>& ct, Galoiskey<Scheme::BFV>& gk, int shift,
                             const ExecutionOptions& opt = ExecutionOptions())
    {
        rotate_rows(ct, ct, gk, shift, opt);
    }
    // add_plain_bfv / sub_plain_bfv / multiply_plain_bfv (bfv/operator.cu:216-340, 432-503)
    void add_plain(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, Ciphertext<Scheme::BFV>& out,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, out, opt, 1);
    }
    void add_plain_inplace(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, a, opt, 1);
    }
    void sub_plain(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, Ciphertext<Scheme::BFV>& out,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, out, opt, 2);
    }
    void sub_plain_inplace(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, const ExecutionOptions& opt = ExecutionOptions())
    {
        plain(a, p, a, opt, 2);
    }
    void multiply_plain(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, Ciphertext<Scheme::BFV>& out,
                        const ExecutionOptions& opt = ExecutionOptions())
    {
        if (a.relinearization_required_)
            throw std::invalid_argument("Ciphertexts can not be multiplied because of the non-linear part! Please use relinearization operation!");
        plain(a, p, out, opt, 0);
    }
    // switchkey_method_I/II (bfv/operator.cu:975-1372)
    void keyswitch(Ciphertext<Scheme::BFV>& in, Ciphertext<Scheme::BFV>& out, Switchkey<Scheme::BFV>& sk,
                   const ExecutionOptions& opt = ExecutionOptions())
    {
        if (in.relinearization_required_)
            throw std::invalid_argument("Ciphertext can not be key-switched because of the non-linear part!");
        if (!sk.switch_key_generated_)
            throw std::invalid_argument("Switchkey is not generated!");
        DeviceVector<Data64> mem(words(2), opt.stream_);
        const detail::KeyView key = sk.view(opt.stream_);
        detail::check(heon_bfv_keyswitch(h(), in.data(), 0, mem.data(), 0, key.ptr, 1, opt.stream_));
        copy_meta(in, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = 2;
    }
    void rotate_columns(Ciphertext<Scheme::BFV>& in, Ciphertext<Scheme::BFV>& out, Galoiskey<Scheme::BFV>& gk,
                        const ExecutionOptions& opt = ExecutionOptions())
    {
        apply_galois(in, out, gk, gk.galois_elt_zero, opt);
    }
    void rotate_columns_inplace(Ciphertext<Scheme::BFV>& ct, Galoiskey<Scheme::BFV>& gk, const ExecutionOptions& opt = ExecutionOptions())
    {
        rotate_columns(ct, ct, gk, opt);
    }
    void apply_galois_inplace(Ciphertext<Scheme::BFV>& ct, Galoiskey<Scheme::BFV>& gk, int galois_elt,
                              const ExecutionOptions& opt = ExecutionOptions())
    {
        apply_galois(ct, ct, gk, galois_elt, opt);
    }
    void keyswitch_inplace(Ciphertext<Scheme::BFV>& ct, Switchkey<Scheme::BFV>& sk, const ExecutionOptions& opt = ExecutionOptions())
    {
        keyswitch(ct, ct, sk, opt);
    }
    void multiply_plain_inplace(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, const ExecutionOptions& opt = ExecutionOptions())
    {
        multiply_plain(a, p, a, opt);
    }

  private:
    void plain(Ciphertext<Scheme::BFV>& a, Plaintext<Scheme::BFV>& p, Ciphertext<Scheme::BFV>& out,
               const ExecutionOptions& opt, int op)
    {
        detail::InputGuard<Ciphertext<Scheme::BFV>> ga(a, opt, &a == &out);
        detail::InputGuard<Plaintext<Scheme::BFV>> gp(p, opt, false);
        const int comps = op == 0 ? 2 : (a.relinearization_required_ ? 3 : 2);
        if (a.memory_size() < words(comps))
            throw std::invalid_argument("Invalid Ciphertexts size!");
        if (p.size() < (size_t) context_->n)
            throw std::invalid_argument("Invalid Plaintext size!");
        DeviceVector<Data64> mem(words(comps), opt.stream_);
        if (op == 0)
            detail::check(heon_bfv_multiply_plain(h(), a.data(), 0, p.data(), 0, mem.data(), 0, 1, opt.stream_));
        else
            detail::check((op == 1 ? heon_bfv_add_plain : heon_bfv_sub_plain)(h(), a.data(), 0, p.data(), 0, mem.data(), 0,
                                                                             comps, 1, opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        out.cipher_size_ = comps;
    }
    void binary(Ciphertext<Scheme::BFV>& a, Ciphertext<Scheme::BFV>& b, Ciphertext<Scheme::BFV>& out,
                const ExecutionOptions& opt, int op)
    {
        detail::InputGuard<Ciphertext<Scheme::BFV>> ga(a, opt, &a == &out), gb(b, opt, &b == &out);
        if (a.cipher_size_ != b.cipher_size_)
            throw std::invalid_argument("Ciphertexts should have the same size!");
        DeviceVector<Data64> mem(words(a.cipher_size_), opt.stream_);
        detail::check((op == 0 ? heon_add : heon_sub)(h(), a.data(), 0, b.data(), 0, mem.data(), 0, a.cipher_size_, 0, 1,
                                                      opt.stream_));
        copy_meta(a, out);
        out.memory_set(std::move(mem));
        detail::output_storage(out, opt);
    }
};

template <> class HEArithmeticOperator<Scheme::BFV> : public HEOperator<Scheme::BFV> {
  public:
    explicit HEArithmeticOperator(HEContext<Scheme::BFV> context) : HEOperator<Scheme::BFV>(context) {}
    HEArithmeticOperator(HEContext<Scheme::BFV> context, HEEncoder<Scheme::BFV>&) : HEOperator<Scheme::BFV>(context) {}
};

} // namespace heongpu

#include "heongpu_client.hpp"
#include "heongpu_serial.hpp"
#include "heongpu_tfhe.hpp"
#include "heongpu_logic.hpp"
