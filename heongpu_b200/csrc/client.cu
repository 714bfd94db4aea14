// Client side of the engine: key generation, public-key encryption, decryption and encoding
// (SURVEY.md section 8(f) rank 2), so that the reference's own tests and benchmarks run end to end
// against this library.  Everything is built from the same NTT and element-wise kernels as the hot path.
//
// reference: src/lib/kernel/keygeneration.cu:13-130 (secret / public key), :145-185, :584-629
// (relinearisation key, Method I / II), :757-860 (Galois key), :896-1030 (switch key);
// src/lib/host/ckks/keygenerator.cu:28-1200, src/lib/host/bfv/keygenerator.cu;
// src/lib/kernel/encryption.cu + src/lib/host/{ckks,bfv}/encryptor.cu (pk*u + e over Q', divide-round by
// P, + plaintext); src/lib/kernel/decryption.cu + host/{ckks,bfv}/decryptor.cu (c0 + c1*s);
// src/lib/host/ckks/encoder.cu + src/lib/kernel/encoding.cu (canonical embedding with the 5^j slot
// order, scale, round, RNS, NTT), src/lib/host/bfv/encoder.cu (batching with generator 3).
//
// Randomness: a counter-based generator (splitmix64 mixing of seed, stream and index), deterministic
// from the caller's seed so that keys are reproducible in tests.  The reference draws from RNGonGPU's
// AES-CTR DRBG seeded by OpenSSL; words therefore differ from the reference's by construction and parity
// for this file is decrypt-level (tests/test_client_side.py, tests/cpp/run_reference_tests.py).
// Distributions follow the reference: a uniform mod each prime, e a rounded Gaussian (sigma 3.2, clipped
// at 6 sigma), u uniform ternary, s ternary with a fixed Hamming weight.
#include <algorithm>
#include <cmath>
#include <vector>
#include <complex>
#include <cstring>
#include "modarith.cuh"
#include "ops.hpp"

namespace heon {

// ---------------------------------------------------------------------------
// randomness
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 mix64(u64 z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 rnd64(u64 seed, u64 stream, u64 ctr)
{
    return mix64(mix64(seed ^ (stream * 0xD1342543DE82EF95ull)) + ctr * 0x9E3779B97F4A7C15ull);
}

// out[(poly*limbs + y)*N + idx] uniform in [0, prime(y)); limb slot y uses prime first_prime + y
__global__ void __launch_bounds__(256)
    k_rng_uniform(u64* __restrict__ out, const PrimeConst* __restrict__ pcs, int logn, int limbs, int first_prime, u64 seed,
                  u64 stream)
{
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    const int y = blockIdx.y;
    const long long poly = blockIdx.z;
    const u64 ctr = ((poly * limbs + y) << logn) + idx;
    const u64 lo = rnd64(seed, stream, 2 * ctr), hi = rnd64(seed, stream, 2 * ctr + 1);
    out[ctr] = reduce_u128(lo, hi, pcs[first_prime + y]); // 128 random bits mod p: bias below 2^-66
}

// one small signed integer per (poly, idx), written as residues into every limb.
// kind 0: rounded Gaussian sigma 3.2 clipped at 6 sigma; kind 1: uniform ternary {-1, 0, 1}
__global__ void __launch_bounds__(256)
    k_rng_small(u64* __restrict__ out, const Mod64* __restrict__ mods, int logn, int limbs, int first_prime, u64 seed,
                u64 stream, int kind)
{
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    const long long poly = blockIdx.y;
    const u64 ctr = (poly << logn) + idx;
    int v;
    if (kind == 0)
    {
        const u64 r = rnd64(seed, stream, ctr);
        const float u1 = ((float) (unsigned) (r >> 40) + 1.0f) * (1.0f / 16777217.0f); // (0,1]
        const float u2 = (float) (unsigned) ((r >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
        const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2) * 3.2f;
        v = (int) rintf(g);
        v = v > 19 ? 19 : v < -19 ? -19 : v;
    }
    else
    {
        // unbiased ternary from 64 bits: floor(3 * r / 2^64) - 1
        v = (int) __umul64hi(rnd64(seed, stream, ctr), 3ull) - 1;
    }
    for (int y = 0; y < limbs; ++y)
    {
        const u64 p = mods[first_prime + y].value;
        out[((poly * limbs + y) << logn) + idx] = v < 0 ? p - (u64) (-v) : (u64) v;
    }
}

// secretkey_rns_kernel (keygeneration.cu:63-91)
__global__ void __launch_bounds__(256)
    k_small_to_rns(const int* __restrict__ in, u64* __restrict__ out, const Mod64* __restrict__ mods, int logn, int limbs)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int v = in[idx];
    for (int y = 0; y < limbs; ++y)
        out[((long long) y << logn) + idx] = v < 0 ? mods[y].value - (u64) (-v) : (u64) v;
}

// ---------------------------------------------------------------------------
// key generation kernels
// ---------------------------------------------------------------------------
// publickey_gen_kernel (keygeneration.cu:93-116): pk0 = -(a*s + e), pk1 = a   (all NTT domain, Q' limbs)
__global__ void __launch_bounds__(256)
    k_pk_gen(u64* __restrict__ pk, const u64* __restrict__ sk, const u64* __restrict__ e, const u64* __restrict__ a,
             const Mod64* __restrict__ mods, int logn, int Qp)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const long long loc = idx + ((long long) y << logn);
    const Mod64 m = mods[y];
    u64 t = barrett_mul(sk[loc], a[loc], m);
    t = mod_add(t, e[loc], m.value);
    pk[loc] = mod_sub(0, t, m.value);
    pk[loc + ((long long) Qp << logn)] = a[loc];
}

// Generic evaluation-key generator (relinkey_gen_kernel :145-185, relinkey_gen_II_kernel :584-629,
// galoiskey_gen_kernel :757-805, switchkey_gen_kernel :896-1030): for digit i and limb y of Q'
//   key[i][0][y] = -(a_i * under + e_i) + [y belongs to digit i] * (P mod m_y) * target,   key[i][1][y] = a_i
// `under`: the secret the switched ciphertext ends up under; `target`: the polynomial the key encrypts
// (s^2 for relinearisation, s for a Galois key whose `under` is the permuted secret, the old secret for a
// switch key).  digit_of[y] = digit that owns limb y (Sk_pair in the reference), < 0 for the P limbs.
__global__ void __launch_bounds__(256)
    k_evk_gen(u64* __restrict__ key, const u64* __restrict__ under, const u64* __restrict__ target,
              const u64* __restrict__ e, const u64* __restrict__ a, const Mod64* __restrict__ mods,
              const u64* __restrict__ pfac, const int* __restrict__ digit_of, int logn, int Qp, int d)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const Mod64 m = mods[y];
    const long long loc = idx + ((long long) y << logn);
    const u64 s = under[loc];
    const int mine = digit_of[y];
    u64 boost = 0;
    if (mine >= 0)
        boost = barrett_mul(target[loc], pfac[y], m);
    for (int i = 0; i < d; ++i)
    {
        const long long src = loc + (((long long) Qp * i) << logn);
        const u64 av = a[src];
        u64 k0 = barrett_mul(s, av, m);
        k0 = mod_add(k0, e[src], m.value);
        k0 = mod_sub(0, k0, m.value);
        if (i == mine)
            k0 = mod_add(k0, boost, m.value);
        const long long dst = loc + (((long long) Qp * i) << (logn + 1));
        key[dst] = k0;
        key[dst + ((long long) Qp << logn)] = av;
    }
}

// out[y] = a[y] * b[y] over `limbs` limbs (s^2, pk*u ...); b may have a different component stride
__global__ void __launch_bounds__(256)
    k_mul_limbs(const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out, const Mod64* __restrict__ mods,
                int logn, long long a_cs, long long o_cs)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const int c = blockIdx.z;
    const long long loc = idx + ((long long) y << logn);
    out[c * o_cs + loc] = barrett_mul(a[c * a_cs + loc], b[loc], mods[y]);
}
// x[c][y] += e[c][y]  (coefficient domain, all limbs)
__global__ void __launch_bounds__(256)
    k_add_limbs(u64* __restrict__ x, const u64* __restrict__ e, const Mod64* __restrict__ mods, int logn, long long cs)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const int c = blockIdx.z;
    const long long loc = c * cs + idx + ((long long) y << logn);
    x[loc] = mod_add(x[loc], e[loc], mods[y].value);
}
// NTT-domain automorphism of one polynomial set [limbs][N]: out[i] = in[src(i)] (see k_galois_permute_ntt;
// `permutation` of keygeneration.cu:742-755 is the same index map)
__global__ void __launch_bounds__(256)
    k_permute_ntt(const u64* __restrict__ in, u64* __restrict__ out, int logn, unsigned galois_elt)
{
    const unsigned idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const unsigned k = __brev(idx) >> (32 - logn);
    const unsigned f = (((2u * k + 1u) * galois_elt) & ((2u << logn) - 1u)) >> 1;
    const unsigned src = __brev(f) >> (32 - logn);
    out[((long long) y << logn) + idx] = in[((long long) y << logn) + src];
}
// decryption: out[y] = c0 + c1*s (+ c2*s^2)   (sk_multiplication_ckks, decryption.cu:349-370)
__global__ void __launch_bounds__(256)
    k_decrypt_dot(const u64* __restrict__ ct, const u64* __restrict__ sk, u64* __restrict__ out,
                  const Mod64* __restrict__ mods, int logn, int L, int comps, int add_c0)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const Mod64 m = mods[y];
    const long long loc = idx + ((long long) y << logn);
    const long long cs = (long long) L << logn;
    const u64 s = sk[loc];
    u64 acc = add_c0 ? ct[loc] : 0;
    u64 sp = s;
    for (int c = 1; c < comps; ++c)
    {
        acc = mod_add(acc, barrett_mul(ct[c * cs + loc], sp, m), m.value);
        sp = barrett_mul(sp, s, m);
    }
    out[loc] = acc;
}

static void chk(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
struct Buf {
    void* p = nullptr;
    cudaStream_t st;
    Buf(size_t bytes, cudaStream_t s) : st(s)
    {
        if (cudaMallocAsync(&p, bytes ? bytes : 8, s) != cudaSuccess)
            throw std::runtime_error("cudaMallocAsync failed (client)");
    }
    ~Buf() { cudaFreeAsync(p, st); }
    u64* w() const { return (u64*) p; }
};

static void sample_uniform(const Context& c, u64* out, int polys, int limbs, u64 seed, u64 stream, cudaStream_t st)
{
    LaunchScope scope(KC_ELEMENTWISE, st);
    k_rng_uniform<<<dim3(c.n >> 8, limbs, polys), 256, 0, st>>>(out, c.d_pc, c.logn, limbs, 0, seed, stream);
}
static void sample_small(const Context& c, u64* out, int polys, int limbs, u64 seed, u64 stream, int kind, cudaStream_t st)
{
    LaunchScope scope(KC_ELEMENTWISE, st);
    k_rng_small<<<dim3(c.n >> 8, polys), 256, 0, st>>>(out, c.d_mod, c.logn, limbs, 0, seed, stream, kind);
}

// generate_secret_key (ckks/keygenerator.cu:28-82): sk [Q'][N], NTT domain
void client_keygen_secret(const Context& c, u64 seed, int hamming_weight, u64* sk, cudaStream_t st)
{
    if (hamming_weight <= 0 || hamming_weight > c.n)
        throw std::invalid_argument("hamming weight has to be in range 0 to ring size.");
    // `hamming_weight` non-zero coefficients (+-1) at distinct positions: the positions with the smallest
    // random keys (the reference's collision-free variant, secretkey_gen_kernel_v2 keygeneration.cu:39-61,
    // takes the positions from the host as well).  Deterministic in the seed.
    std::vector<std::pair<u64, int>> order(c.n);
    for (int i = 0; i < c.n; ++i)
        order[i] = {rnd64(seed, 0x5EC, (u64) i), i};
    std::sort(order.begin(), order.end());
    std::vector<int> h_sk(c.n, 0);
    for (int i = 0; i < hamming_weight; ++i)
        h_sk[order[i].second] = (rnd64(seed, 0x5ED, (u64) i) & 1) ? 1 : -1;
    Buf raw((size_t) c.n * sizeof(int), st);
    cudaMemcpyAsync(raw.p, h_sk.data(), (size_t) c.n * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_small_to_rns<<<c.n >> 8, 256, 0, st>>>((const int*) raw.p, sk, c.d_mod, c.logn, c.Qp);
    }
    chk("secret key");
    launch_ntt(c, sk, sk, c.Qp, range_primes(0, c.Qp), false, st);
}

// generate_public_key (ckks/keygenerator.cu:167-243): pk [2][Q'][N]
void client_keygen_public(const Context& c, const u64* sk, u64 seed, u64* pk, cudaStream_t st)
{
    const size_t w = (size_t) c.Qp * c.n;
    Buf ea(2 * w * 8, st);
    u64 *e = ea.w(), *a = ea.w() + w;
    sample_uniform(c, a, 1, c.Qp, seed, 1, st);
    sample_small(c, e, 1, c.Qp, seed, 2, 0, st);
    launch_ntt(c, e, e, c.Qp, range_primes(0, c.Qp), false, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_pk_gen<<<dim3(c.n >> 8, c.Qp), 256, 0, st>>>(pk, sk, e, a, c.d_mod, c.logn, c.Qp);
    }
    chk("public key");
}

static int digits0(const Context& c) { return c.method == 1 ? c.Q_size : c.lvl2[0].d; }

// key [d][2][Q'][N]
void client_keygen_evk(const Context& c, const u64* under, const u64* target, u64 seed, u64* key, cudaStream_t st)
{
    const int d = digits0(c), Qp = c.Qp, Q = c.Q_size, K = c.P_size;
    const size_t w = (size_t) d * Qp * c.n;
    Buf ea(2 * w * 8, st);
    u64 *e = ea.w(), *a = ea.w() + w;
    sample_uniform(c, a, d, Qp, seed, 3, st);
    sample_small(c, e, d, Qp, seed, 4, 0, st);
    launch_ntt(c, e, e, (long long) d * Qp, range_primes(0, Qp), false, st);
    // P mod m_y (product of the factor_ rows, util.cu:751-767) and the digit that owns each limb
    std::vector<u64> pfac(Qp, 0);
    std::vector<int> owner(Qp, -1);
    for (int y = 0; y < Q; ++y)
    {
        u64 f = 1;
        for (int j = 0; j < K; ++j)
            f = mulmod(f, c.factor[(size_t) j * Q + y], c.mod[y].value);
        pfac[y] = f;
        if (c.method == 1)
            owner[y] = y;
        else
            for (int i = 0; i < d; ++i)
                if (y >= c.lvl2[0].I_loc[i] && y < c.lvl2[0].I_loc[i] + c.lvl2[0].I_j[i])
                    owner[y] = i;
    }
    Buf tab((size_t) Qp * 16, st);
    u64* d_pfac = tab.w();
    int* d_owner = (int*) (tab.w() + Qp);
    cudaMemcpyAsync(d_pfac, pfac.data(), Qp * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_owner, owner.data(), Qp * 4, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st); // the staging vectors live on this stack frame
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_evk_gen<<<dim3(c.n >> 8, Qp), 256, 0, st>>>(key, under, target, e, a, c.d_mod, d_pfac, d_owner, c.logn, Qp, d);
    }
    chk("evaluation key");
}

void client_keygen_relin(const Context& c, const u64* sk, u64 seed, u64* key, cudaStream_t st)
{
    Buf s2((size_t) c.Qp * c.n * 8, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_mul_limbs<<<dim3(c.n >> 8, c.Qp, 1), 256, 0, st>>>(sk, sk, s2.w(), c.d_mod, c.logn, 0, 0);
    }
    client_keygen_evk(c, sk, s2.w(), seed, key, st);
}

static unsigned inv_mod_pow2(unsigned a, unsigned m) // a odd, m a power of two
{
    unsigned x = a; // Newton: x <- x * (2 - a*x), correct bits double each step
    for (int i = 0; i < 6; ++i)
        x *= 2u - a * x;
    return x & (m - 1);
}

// generate_galois_key (ckks/keygenerator.cu:416-700): the key switch runs BEFORE the automorphism
// (apply_galois = permute(keyswitch(ct))), so the key is under sigma_{g^-1}(s) and encrypts P*s
void client_keygen_galois(const Context& c, const u64* sk, unsigned galois_elt, u64 seed, u64* key, cudaStream_t st)
{
    if (!(galois_elt & 1) || galois_elt >= 2u * c.n)
        throw std::invalid_argument("invalid Galois element");
    const unsigned inv = inv_mod_pow2(galois_elt, 2u * c.n);
    Buf sp((size_t) c.Qp * c.n * 8, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_permute_ntt<<<dim3(c.n >> 8, c.Qp), 256, 0, st>>>(sk, sp.w(), c.logn, inv);
    }
    client_keygen_evk(c, sp.w(), sk, seed ^ ((u64) galois_elt << 32), key, st);
}

void op_moddown_coeff(const Context& c, const u64* in, u64* out, long long out_bs, int batch, cudaStream_t st);

// encrypt_ckks / encrypt_bfv (host/{ckks,bfv}/encryptor.cu): ct = round((pk*u + e) / P) + plaintext.
// pt: CKKS [Q][N] NTT domain (added to c0); BFV [N] below the plain modulus (scaled by floor(Q/t)).
// ct: [2][Q][N] (CKKS: NTT domain, BFV: coefficient domain).  pt == nullptr encrypts zero.
void client_encrypt(const Context& c, const u64* pk, const u64* pt, u64 seed, u64* ct, cudaStream_t st)
{
    const int Qp = c.Qp, Q = c.Q_size;
    const size_t w = (size_t) Qp * c.n;
    Buf work(5 * w * 8, st);
    u64 *u = work.w(), *e = u + w, *pku = e + 2 * w;
    sample_small(c, u, 1, Qp, seed, 5, 1, st);
    sample_small(c, e, 2, Qp, seed, 6, 0, st);
    launch_ntt(c, u, u, Qp, range_primes(0, Qp), false, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_mul_limbs<<<dim3(c.n >> 8, Qp, 2), 256, 0, st>>>(pk, u, pku, c.d_mod, c.logn, (long long) w, (long long) w);
    }
    launch_ntt(c, pku, pku, 2 * Qp, range_primes(0, Qp), true, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_add_limbs<<<dim3(c.n >> 8, Qp, 2), 256, 0, st>>>(pku, e, c.d_mod, c.logn, (long long) w);
    }
    chk("encrypt");
    op_moddown_coeff(c, pku, ct, 2ll * Q * c.n, 1, st); // enc_div_lastq_*_kernel: divide-round by every P prime
    if (c.scheme == SCHEME_CKKS)
    {
        launch_ntt(c, ct, ct, 2 * Q, range_primes(0, Q), false, st);
        if (pt)
            op_plain(c, ct, 0, pt, 0, ct, 0, 2, 0, 1, 1, st); // cipher_message_add_kernel
    }
    else if (pt)
        op_bfv_addsub_plain(c, ct, 0, pt, 0, ct, 0, 2, 1, 1, st);
}

// decrypt (ckks/decryptor.cu): pt [L][N] NTT domain = c0 + c1*s (+ c2*s^2)
void client_decrypt_ckks(const Context& c, const u64* sk, const u64* ct, int comps, int depth, u64* pt, cudaStream_t st)
{
    const int L = c.Q_size - depth;
    if (depth < 0 || L < 1 || comps < 2 || comps > 3)
        throw std::invalid_argument("invalid ciphertext");
    LaunchScope scope(KC_ELEMENTWISE, st);
    k_decrypt_dot<<<dim3(c.n >> 8, L), 256, 0, st>>>(ct, sk, pt, c.d_mod, c.logn, L, comps, 1);
    chk("decrypt");
}

// ---------------------------------------------------------------------------
// host-side helpers: exact CRT scaling for BFV decryption, canonical embedding
// ---------------------------------------------------------------------------
// BFV: x = c0 + c1*s (+ c2 s^2) in coefficient domain over Q; message = round(t * [x]_Q / Q) mod t.
// [x]_Q / Q = frac(sum_i y_i / q_i), y_i = x_i * (Q/q_i)^-1 mod q_i  (80-bit long double: the sum carries
// more than 60 correct fractional bits, the decision needs log2(t) + noise margin of them)
void client_decrypt_bfv(const Context& c, const u64* sk, const u64* ct, int comps, u64* pt, cudaStream_t st, int* budget = nullptr);
void client_noise_budget_bfv(const Context& c, const u64* sk, const u64* ct, int comps, int* bits, cudaStream_t st)
{
    client_decrypt_bfv(c, sk, ct, comps, nullptr, st, bits);
}
void client_decrypt_bfv(const Context& c, const u64* sk, const u64* ct, int comps, u64* pt, cudaStream_t st, int* budget)
{
    const int Q = c.Q_size;
    const size_t N = c.n;
    if (comps < 2 || comps > 3)
        throw std::invalid_argument("invalid ciphertext");
    Buf work((size_t) (comps + 1) * Q * N * 8, st);
    u64 *tmp = work.w(), *dot = tmp + (size_t) comps * Q * N;
    cudaMemcpyAsync(tmp, ct, (size_t) comps * Q * N * 8, cudaMemcpyDeviceToDevice, st);
    launch_ntt(c, tmp + Q * N, tmp + Q * N, (long long) (comps - 1) * Q, range_primes(0, Q), false, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_decrypt_dot<<<dim3(c.n >> 8, Q), 256, 0, st>>>(tmp, sk, dot, c.d_mod, c.logn, Q, comps, 0);
    }
    launch_ntt(c, dot, dot, Q, range_primes(0, Q), true, st);
    {
        LaunchScope scope(KC_ELEMENTWISE, st);
        k_add_limbs<<<dim3(c.n >> 8, Q, 1), 256, 0, st>>>(dot, ct, c.d_mod, c.logn, 0);
    }
    chk("bfv decrypt");
    std::vector<u64> h((size_t) Q * N);
    cudaMemcpyAsync(h.data(), dot, h.size() * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    std::vector<u64> inv(Q);
    for (int i = 0; i < Q; ++i)
    {
        const u64 qi = c.mod[i].value;
        u64 prod = 1;
        for (int j = 0; j < Q; ++j)
            if (j != i)
                prod = mulmod(prod, c.mod[j].value % qi, qi);
        inv[i] = invmod(prod, qi);
    }
    const u64 t = c.plain_modulus;
    std::vector<u64> msg(N);
    long double worst = 0.0L;
    for (size_t j = 0; j < N; ++j)
    {
        long double frac = 0.0L;
        for (int i = 0; i < Q; ++i)
        {
            const u64 qi = c.mod[i].value;
            const u64 yi = mulmod(h[(size_t) i * N + j], inv[i], qi);
            frac += (long double) yi / (long double) qi;
        }
        frac -= floorl(frac);
        long double v = roundl(frac * (long double) t);
        worst = fmaxl(worst, fabsl(frac * (long double) t - v));
        u64 m = (u64) v;
        msg[j] = m >= t ? m - t : m;
    }
    if (budget)
    {
        const long double b = worst > 0 ? -log2l(2.0L * worst) : 62.0L;
        *budget = b < 0 ? 0 : b > 62 ? 62 : (int) floorl(b);
    }
    if (pt)
    {
        cudaMemcpyAsync(pt, msg.data(), N * 8, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
    }
}

typedef std::complex<double> cplx;
static inline int brev_host(int x, int bits)
{
    int r = 0;
    for (int i = 0; i < bits; ++i)
    {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}
// psi^bitrev(i), psi = exp(i*pi/N): the complex twin of the NTT table (util.cu:398-451)
static std::vector<cplx> embed_roots(int logn, bool inverse)
{
    const int N = 1 << logn;
    std::vector<cplx> r(N);
    const long double step = 3.141592653589793238462643383279502884L / (long double) N;
    for (int i = 0; i < N; ++i)
    {
        const long double ang = step * (long double) brev_host(i, logn) * (inverse ? -1.0L : 1.0L);
        r[i] = cplx((double) cosl(ang), (double) sinl(ang));
    }
    return r;
}
// forward: natural in, bit-reversed out, out[i] = a(psi^(2*brev(i)+1))  (the loop of ntt_cpu.cu:81-130)
static void embed_forward(std::vector<cplx>& a, int logn, const std::vector<cplx>& tw)
{
    const int N = 1 << logn;
    int t = N, m = 1;
    while (m < N)
    {
        t >>= 1;
        for (int i = 0; i < m; ++i)
        {
            const cplx S = tw[m + i];
            const int j1 = 2 * i * t;
            for (int j = j1; j < j1 + t; ++j)
            {
                const cplx U = a[j], V = a[j + t] * S;
                a[j] = U + V;
                a[j + t] = U - V;
            }
        }
        m <<= 1;
    }
}
// inverse: bit-reversed in, natural out, scaled by 1/N (ntt_cpu.cu:132-188)
static void embed_inverse(std::vector<cplx>& a, int logn, const std::vector<cplx>& tw)
{
    const int N = 1 << logn;
    int t = 1, m = N;
    while (m > 1)
    {
        int j1 = 0;
        const int h = m >> 1;
        for (int i = 0; i < h; ++i)
        {
            const cplx S = tw[h + i];
            for (int j = j1; j < j1 + t; ++j)
            {
                const cplx U = a[j], V = a[j + t];
                a[j] = U + V;
                a[j + t] = (U - V) * S;
            }
            j1 += 2 * t;
        }
        t <<= 1;
        m >>= 1;
    }
    const double s = 1.0 / (double) N;
    for (auto& v : a)
        v *= s;
}
// slot i <-> transform index: matrix_reps_index_map of the canonical embedding (generator `gen`)
static std::vector<int> slot_map(int logn, int gen)
{
    const int N = 1 << logn, slots = N >> 1, m = 2 * N;
    std::vector<int> map(N);
    long long pos = 1;
    for (int i = 0; i < slots; ++i)
    {
        const int i1 = (int) ((pos - 1) >> 1), i2 = (int) ((m - pos - 1) >> 1);
        map[i] = brev_host(i1, logn);
        map[slots + i] = brev_host(i2, logn);
        pos = (pos * gen) & (m - 1);
    }
    return map;
}

// HEEncoder<CKKS>::encode (ckks/encoder.cu, encoding.cu:43-141): values = complex slots (re, im pairs),
// count <= N/2; pt [L][N] NTT domain at `depth`
void client_ckks_encode(const Context& c, const double* values, int count, double scale, int depth, u64* pt, cudaStream_t st)
{
    const int N = c.n, slots = N >> 1, L = c.Q_size - depth;
    if (count < 0 || count > slots)
        throw std::invalid_argument("Vector size can not be higher than slot count!");
    if (L < 1 || depth < 0)
        throw std::invalid_argument("invalid depth");
    if (!(scale > 0.0))
        throw std::invalid_argument("Scale can not be negative or zero");
    const std::vector<int> map = slot_map(c.logn, 5);
    std::vector<cplx> v(N, cplx(0, 0));
    for (int i = 0; i < count; ++i)
    {
        const cplx z(values[2 * i], values[2 * i + 1]);
        v[map[i]] = z;
        v[map[slots + i]] = std::conj(z);
    }
    embed_inverse(v, c.logn, embed_roots(c.logn, true));
    std::vector<u64> h((size_t) L * N);
    for (int j = 0; j < N; ++j)
    {
        const double x = nearbyint(v[j].real() * scale);
        if (!(fabs(x) < 9.0e18))
            throw std::invalid_argument("encoded value out of range: scale too large for this message");
        const long long xi = (long long) x;
        for (int y = 0; y < L; ++y)
        {
            const u64 p = c.mod[y].value;
            const long long r = xi % (long long) p;
            h[(size_t) y * N + j] = r < 0 ? (u64) (r + (long long) p) : (u64) r;
        }
    }
    cudaMemcpyAsync(pt, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    launch_ntt(c, pt, pt, L, range_primes(0, L), false, st);
}

// little-endian multi-word helpers for the CRT composition
static void big_muladd(std::vector<u64>& acc, const std::vector<u64>& a, u64 m)
{
    u128 carry = 0;
    for (size_t i = 0; i < acc.size(); ++i)
    {
        const u128 t = (u128) (i < a.size() ? a[i] : 0) * m + acc[i] + carry;
        acc[i] = (u64) t;
        carry = t >> 64;
    }
}
static int big_cmp(const std::vector<u64>& a, const std::vector<u64>& b)
{
    for (size_t i = a.size(); i-- > 0;)
    {
        const u64 x = a[i], y = i < b.size() ? b[i] : 0;
        if (x != y)
            return x < y ? -1 : 1;
    }
    return 0;
}
static void big_sub(std::vector<u64>& a, const std::vector<u64>& b)
{
    u64 borrow = 0;
    for (size_t i = 0; i < a.size(); ++i)
    {
        const u64 y = i < b.size() ? b[i] : 0;
        const u64 t = a[i] - y - borrow;
        borrow = (a[i] < y + borrow) || (y + borrow < y) ? 1 : 0;
        a[i] = t;
    }
}
static double big_to_double(const std::vector<u64>& a)
{
    double r = 0.0;
    for (size_t i = a.size(); i-- > 0;)
        r = r * 18446744073709551616.0 + (double) a[i];
    return r;
}

// HEEncoder<CKKS>::decode (ckks/encoder.cu, encoding.cu:234-400): INTT, CRT composition to the centred
// integer, division by the scale, canonical embedding.  out: `count` complex slots (re, im pairs)
void client_ckks_decode(const Context& c, const u64* pt, int depth, double scale, double* out, int count, cudaStream_t st)
{
    const int N = c.n, slots = N >> 1, L = c.Q_size - depth;
    if (count < 0 || count > slots || L < 1 || depth < 0)
        throw std::invalid_argument("invalid decode request");
    Buf work((size_t) L * N * 8, st);
    cudaMemcpyAsync(work.p, pt, (size_t) L * N * 8, cudaMemcpyDeviceToDevice, st);
    launch_ntt(c, work.w(), work.w(), L, range_primes(0, L), true, st);
    std::vector<u64> h((size_t) L * N);
    cudaMemcpyAsync(h.data(), work.p, h.size() * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    // Q, Q/q_i, (Q/q_i)^-1 mod q_i
    const size_t W = (size_t) L + 1;
    std::vector<u64> bigQ(W, 0);
    bigQ[0] = 1;
    for (int i = 0; i < L; ++i)
    {
        std::vector<u64> t(W, 0);
        big_muladd(t, bigQ, c.mod[i].value);
        bigQ = t;
    }
    std::vector<std::vector<u64>> qhat(L, std::vector<u64>(W, 0));
    std::vector<u64> inv(L);
    for (int i = 0; i < L; ++i)
    {
        qhat[i][0] = 1;
        const u64 qi = c.mod[i].value;
        u64 prod = 1;
        for (int j = 0; j < L; ++j)
            if (j != i)
            {
                std::vector<u64> t(W, 0);
                big_muladd(t, qhat[i], c.mod[j].value);
                qhat[i] = t;
                prod = mulmod(prod, c.mod[j].value % qi, qi);
            }
        inv[i] = invmod(prod, qi);
    }
    std::vector<u64> halfQ(bigQ);
    for (size_t i = 0; i < W; ++i)
        halfQ[i] = (bigQ[i] >> 1) | (i + 1 < W ? bigQ[i + 1] << 63 : 0);
    std::vector<cplx> v(N);
    std::vector<u64> acc(W);
    for (int j = 0; j < N; ++j)
    {
        std::fill(acc.begin(), acc.end(), 0);
        for (int i = 0; i < L; ++i)
            big_muladd(acc, qhat[i], mulmod(h[(size_t) i * N + j], inv[i], c.mod[i].value));
        while (big_cmp(acc, bigQ) >= 0)
            big_sub(acc, bigQ);
        double x;
        if (big_cmp(acc, halfQ) > 0)
        {
            std::vector<u64> neg(bigQ);
            big_sub(neg, acc);
            x = -big_to_double(neg);
        }
        else
            x = big_to_double(acc);
        v[j] = cplx(x / scale, 0.0);
    }
    embed_forward(v, c.logn, embed_roots(c.logn, false));
    const std::vector<int> map = slot_map(c.logn, 5);
    for (int i = 0; i < count; ++i)
    {
        out[2 * i] = v[map[i]].real();
        out[2 * i + 1] = v[map[i]].imag();
    }
}

// ---- BFV batching (bfv/encoder.cu, encode_kernel_bfv / decode_kernel_bfv, encoding.cu:11-41):
// slots through the index map of generator 3, negacyclic transform modulo the plain modulus t
static void plain_ntt(std::vector<u64>& a, int logn, u64 t, bool inverse)
{
    const int N = 1 << logn;
    const u64 psi = minimal_primitive_root(2 * (u64) N, t);
    const u64 root = inverse ? invmod(psi, t) : psi;
    std::vector<u64> pw(N), tw(N);
    pw[0] = 1;
    for (int i = 1; i < N; ++i)
        pw[i] = mulmod(pw[i - 1], root, t);
    for (int i = 0; i < N; ++i)
        tw[i] = pw[brev_host(i, logn)];
    if (!inverse)
    {
        int tt = N, m = 1;
        while (m < N)
        {
            tt >>= 1;
            for (int i = 0; i < m; ++i)
            {
                const u64 S = tw[m + i];
                const int j1 = 2 * i * tt;
                for (int j = j1; j < j1 + tt; ++j)
                {
                    const u64 U = a[j], V = mulmod(a[j + tt], S, t);
                    a[j] = addmod(U, V, t);
                    a[j + tt] = submod(U, V, t);
                }
            }
            m <<= 1;
        }
        return;
    }
    int tt = 1, m = N;
    while (m > 1)
    {
        int j1 = 0;
        const int h = m >> 1;
        for (int i = 0; i < h; ++i)
        {
            const u64 S = tw[h + i];
            for (int j = j1; j < j1 + tt; ++j)
            {
                const u64 U = a[j], V = a[j + tt];
                a[j] = addmod(U, V, t);
                a[j + tt] = mulmod(submod(U, V, t), S, t);
            }
            j1 += 2 * tt;
        }
        tt <<= 1;
        m >>= 1;
    }
    const u64 ninv = invmod((u64) N % t, t);
    for (auto& v : a)
        v = mulmod(v, ninv, t);
}

void client_bfv_encode(const Context& c, const u64* msg, int count, u64* pt, cudaStream_t st)
{
    const int N = c.n;
    const u64 t = c.plain_modulus;
    if (c.scheme != SCHEME_BFV)
        throw std::invalid_argument("not a BFV context");
    if (count < 0 || count > N)
        throw std::invalid_argument("Vector size can not be higher than slot count!");
    if ((t - 1) % (2ull * N))
        throw std::invalid_argument("plain modulus does not support batching (t != 1 mod 2N)");
    const std::vector<int> map = slot_map(c.logn, 3);
    std::vector<u64> a(N, 0);
    for (int i = 0; i < count; ++i)
        a[map[i]] = msg[i] % t;
    plain_ntt(a, c.logn, t, true);
    cudaMemcpyAsync(pt, a.data(), (size_t) N * 8, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
}

void client_bfv_decode(const Context& c, const u64* pt, u64* msg, int count, cudaStream_t st)
{
    const int N = c.n;
    const u64 t = c.plain_modulus;
    if (c.scheme != SCHEME_BFV)
        throw std::invalid_argument("not a BFV context");
    if (count < 0 || count > N)
        throw std::invalid_argument("Vector size can not be higher than slot count!");
    std::vector<u64> a(N);
    cudaMemcpyAsync(a.data(), pt, (size_t) N * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    plain_ntt(a, c.logn, t, false);
    const std::vector<int> map = slot_map(c.logn, 3);
    for (int i = 0; i < count; ++i)
        msg[i] = a[map[i]];
}

} // namespace heon
