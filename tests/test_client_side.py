"""GPU, decrypt-level: the client side (key generation, encoding, public-key encryption, decryption)
against the plaintext computation, in the shape of the reference's own tests
(test/test_ckks_relinearization.cpp:36-830, test_ckks_rotation_method_*.cpp, test_ckks_encoding.cpp,
test_ckks_encryption.cpp, test_bfv_multiplication.cpp, test_bfv_rotation_method_*.cpp)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _api():
    from heongpu_b200 import api
    return api


def _ckks(log_n, qb, pb):
    api = _api()
    ctx = api.HEContext(log_n, qb, pb, device=0)
    kg = api.HEKeyGenerator(ctx, seed=1234)
    sk = kg.generate_secret_key(api.Secretkey(ctx))
    pk = kg.generate_public_key(api.Publickey(ctx), sk)
    return api, ctx, kg, sk, pk


@pytest.mark.parametrize("log_n,qb,pb", [(12, [40, 30, 30], [40]), (13, [40, 30, 30, 30, 30], [40]),
                                         (12, [40, 30, 30, 30], [40, 40]), (13, [50, 40, 40, 40, 40], [50, 50, 50])])
def test_ckks_encode_encrypt_decrypt_roundtrip(log_n, qb, pb):
    api, ctx, kg, sk, pk = _ckks(log_n, qb, pb)
    enc, cry, dec = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk), api.HEDecryptor(ctx, sk)
    rng = np.random.default_rng(1)
    slots = ctx.n // 2
    m = rng.uniform(-1, 1, slots) + 1j * rng.uniform(-1, 1, slots)
    scale = 2.0 ** qb[-1]
    pt = enc.encode(m, scale)
    assert np.allclose(enc.decode(pt), m, atol=1e-6)  # test_ckks_encoding.cpp
    ct = cry.encrypt(pt)
    got = enc.decode(dec.decrypt(ct))
    assert np.allclose(got, m, atol=1e-4)  # test_ckks_encryption.cpp


@pytest.mark.parametrize("log_n,qb,pb", [(12, [40, 30, 30], [40]), (12, [40, 30, 30, 30], [40, 40]),
                                         (14, [50, 40, 40, 40], [48]), (13, [50, 40, 40, 40, 40], [50, 50, 50])])
def test_ckks_multiply_relinearize_rescale_rotate_decrypts_correctly(log_n, qb, pb):
    api, ctx, kg, sk, pk = _ckks(log_n, qb, pb)
    enc, cry, dec = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk), api.HEDecryptor(ctx, sk)
    op = api.HEArithmeticOperator(ctx)
    rk = kg.generate_relin_key(sk)
    gk = kg.generate_galois_key(sk, shifts=[1, -3, 5])
    rng = np.random.default_rng(2)
    slots = ctx.n // 2
    m1, m2 = rng.uniform(0, 1, slots), rng.uniform(0, 1, slots)
    scale = 2.0 ** qb[-1]
    c1, c2 = cry.encrypt(enc.encode(m1, scale)), cry.encrypt(enc.encode(m2, scale))
    L, n = ctx.Q_size, ctx.n
    prod = api.Ciphertext(ctx, torch.zeros(1, 3, L, n, dtype=torch.int64, device="cuda"))
    op.multiply(c1, c2, prod)
    op.relinearize_inplace(prod, rk)
    op.rescale_inplace(prod)
    got = enc.decode(dec.decrypt(prod)).real
    assert np.allclose(got, m1 * m2, atol=1e-4)  # test_ckks_relinearization.cpp
    for shift in (1, -3, 5):
        out = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"))
        op.rotate_rows(c1, out, gk, shift)
        got = enc.decode(dec.decrypt(out)).real
        # Method I with a 40-bit q_0 next to a 40-bit P: the uncentred digit [c1]_{q_0} has mean q_0/2, which puts a
        # noise peak of ~1e-4 at scale 2^30 on the slot whose root is nearest to 1 (the reference's kernels, to
        # which the key switch is bit-identical, carry the same peak); every other slot is two orders below
        err = np.abs(got - np.roll(m1, -shift))
        assert err.max() < 5e-4 and np.median(err) < 1e-5, (shift, err.max())  # test_ckks_rotation_method_*.cpp
    conj = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"))
    z = m1 + 1j * m2
    cz = cry.encrypt(enc.encode(z, scale))
    op.conjugate(cz, conj, gk)
    err = np.abs(enc.decode(dec.decrypt(conj)) - np.conj(z))
    assert err.max() < 5e-4 and np.median(err) < 1e-5, err.max()


def test_ckks_switch_key_re_encrypts_under_the_new_secret():
    api, ctx, kg, sk, pk = _ckks(12, [40, 30, 30, 30], [40, 40])
    sk2 = kg.generate_secret_key(api.Secretkey(ctx))
    swk = kg.generate_switch_key(sk2, sk)
    enc, cry = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk)
    op = api.HEArithmeticOperator(ctx)
    m = np.linspace(-1, 1, ctx.n // 2)
    ct = cry.encrypt(enc.encode(m, 2.0 ** 30))
    out = api.Ciphertext(ctx, torch.zeros(1, 2, ctx.Q_size, ctx.n, dtype=torch.int64, device="cuda"))
    op.keyswitch(ct, out, swk)
    out.scale_ = ct.scale_
    assert np.allclose(api.HEEncoder(ctx).decode(api.HEDecryptor(ctx, sk2).decrypt(out)).real, m, atol=1e-4)


def test_keys_are_reproducible_from_the_seed():
    api = _api()
    ctx = api.HEContext(12, [40, 30, 30], [40], device=0)
    a = api.HEKeyGenerator(ctx, seed=7).generate_secret_key(api.Secretkey(ctx))
    b = api.HEKeyGenerator(ctx, seed=7).generate_secret_key(api.Secretkey(ctx))
    d = api.HEKeyGenerator(ctx, seed=8).generate_secret_key(api.Secretkey(ctx))
    assert torch.equal(a.data, b.data) and not torch.equal(a.data, d.data)
    # Hamming weight n/2, coefficients in {-1, 0, 1}
    coef = a.data[:1].clone()
    ctx.ntt(coef, [0], inverse=True)
    p = ctx.primes[0]
    v = coef.cpu().numpy().view(np.uint64)[0]
    assert set(np.unique(v)) <= {0, 1, p - 1} and int((v != 0).sum()) == ctx.n // 2


@pytest.mark.parametrize("log_n,qb,pb,t", [(12, [36, 36], [37], 1032193), (13, [54, 54, 54], [55], 786433),
                                           (12, [40, 40], [40, 40], 1032193)])
def test_bfv_batching_multiply_relinearize_rotate_decrypts_correctly(log_n, qb, pb, t):
    api = _api()
    ctx = api.HEContext(log_n, qb, pb, device=0, plain_modulus=t)
    kg = api.HEKeyGenerator(ctx, seed=99)
    sk = kg.generate_secret_key(api.Secretkey(ctx))
    pk = kg.generate_public_key(api.Publickey(ctx), sk)
    rk = kg.generate_relin_key(sk)
    gk = kg.generate_galois_key(sk, shifts=[1, 2])
    enc, cry, dec = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk), api.HEDecryptor(ctx, sk)
    op = api.HEArithmeticOperator(ctx)
    rng = np.random.default_rng(3)
    n, Q = ctx.n, ctx.Q_size
    m1, m2 = rng.integers(0, t, n), rng.integers(0, t, n)
    p1 = enc.encode(m1)
    assert np.array_equal(enc.decode(p1), m1.astype(np.uint64))  # test_bfv_encoding.cpp
    c1, c2 = cry.encrypt(p1), cry.encrypt(enc.encode(m2))
    assert np.array_equal(enc.decode(dec.decrypt(c1)), m1.astype(np.uint64))  # test_bfv_encryption.cpp
    prod = api.Ciphertext(ctx, torch.zeros(1, 3, Q, n, dtype=torch.int64, device="cuda"))
    op.multiply_bfv(c1, c2, prod)
    op.relinearize_inplace_bfv(prod, rk)
    got = enc.decode(dec.decrypt(prod))
    assert np.array_equal(got, (m1 * m2 % t).astype(np.uint64))  # test_bfv_multiplication / relinearization
    half = n // 2
    for shift in (1, 2):
        out = api.Ciphertext(ctx, torch.zeros(1, 2, Q, n, dtype=torch.int64, device="cuda"))
        op.rotate_rows_bfv(c1, out, gk, shift)
        got = enc.decode(dec.decrypt(out))
        want = np.concatenate([np.roll(m1[:half], -shift), np.roll(m1[half:], -shift)]).astype(np.uint64)
        assert np.array_equal(got, want), shift  # test_bfv_rotation_method_*.cpp
    out = api.Ciphertext(ctx, torch.zeros(1, 2, Q, n, dtype=torch.int64, device="cuda"))
    op.rotate_columns_bfv(c1, out, gk)
    assert np.array_equal(enc.decode(dec.decrypt(out)), np.concatenate([m1[half:], m1[:half]]).astype(np.uint64))


def test_ckks_bsgs_matrix_product_decrypts_correctly():
    """multiply_matrix_v2 (ckks/operator.cu:2898-3390) with real keys: sum_j rot_{G_j}( sum_k diag_jk * rot_{b_k}(x) )
    against numpy.  The diagonals are encoded over PQ_0 by a second context whose Q chain is this one's Q' chain."""
    api, ctx, kg, sk, pk = _ckks(13, [50, 40, 40, 40, 40], [50, 50, 50])
    spare = api.HEContext(13, [45], [46], device=0).primes[1]
    wide = api.HEContext(13, q_values=list(ctx.primes), p_values=[spare], device=0)
    enc, cry, dec = api.HEEncoder(ctx), api.HEEncryptor(ctx, pk), api.HEDecryptor(ctx, sk)
    wenc = api.HEEncoder(wide)
    op = api.HEArithmeticOperator(ctx)
    rot_n2, rot_n1 = [0, 1, 2, 3], [0, 4, 8]
    diags_bsgs = [[0, 1, 2, 3], [4, 5, 7], [8, 10, 11]]
    gk = kg.generate_galois_key(sk, shifts=[1, 2, 3, 4, 8])
    rng = np.random.default_rng(5)
    slots = ctx.n // 2
    x = rng.uniform(-1, 1, slots)
    scale = 2.0 ** 40
    ct = cry.encrypt(enc.encode(x, scale))
    want = np.zeros(slots)
    planes = []
    for j, group in enumerate(diags_bsgs):
        inner = np.zeros(slots)
        for dg in group:
            p = rng.uniform(-1, 1, slots)
            planes.append(wenc.encode(p, scale).data.reshape(-1, ctx.n))
            inner += p * np.roll(x, -(dg - rot_n1[j]))
        want += np.roll(inner, -rot_n1[j])
    matrix = torch.stack(planes).contiguous()
    L, n = ctx.Q_size, ctx.n
    assert matrix.shape == (len(planes), L + ctx.P_size, n)
    out = api.Ciphertext(ctx, torch.zeros(1, 2, L, n, dtype=torch.int64, device="cuda"))
    op.multiply_matrix(ct, out, matrix, diags_bsgs, rot_n1, rot_n2, gk)
    assert out.depth_ == 1
    out.scale_ = scale * scale / float(ctx.primes[L - 1])
    got = enc.decode(dec.decrypt(out)).real
    err = np.abs(got - want)
    assert err.max() < 1e-4 and np.median(err) < 1e-6, (err.max(), np.median(err))
