"""GPU parity for TFHE gate bootstrapping (BASELINE config 5).
 (a) bit-exact against the reference's OWN kernels (oracle/_ref/libref_tfhe.so: small_ntt.cu + bootstrapping.cu
     compiled unmodified, launch replay of src/lib/host/tfhe/operator.cu): the 1024-point transform, every gate's
     linear part, the whole blind rotation + sample extraction (1 launch here, 1024 there), the key switch;
 (b) decrypt-level: the truth tables of all gates with real keys, in the shape of test/test_tfhe_gate_boot.cpp."""
import numpy as np
import pytest
import torch

from oracle import ref as R

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.have_tfhe(), reason="oracle/_ref/libref_tfhe.so not built")
P = 1152921504606877697


def _t():
    from heongpu_b200 import tfhe
    return tfhe


_STATE = {}


def _keys():
    if "k" not in _STATE:
        t = _t()
        ctx = t.HEContext(0)
        kg = t.HEKeyGenerator(ctx, seed=77)
        sk = kg.generate_secret_key(t.Secretkey(ctx))
        bk = kg.generate_bootstrapping_key(t.Bootstrappingkey(ctx), sk)
        torch.cuda.synchronize()
        _STATE["k"] = (ctx, sk, bk)
    return _STATE["k"]


def _rand_i32(gen, *shape):
    return torch.randint(-2 ** 31, 2 ** 31, shape, generator=gen, device="cuda", dtype=torch.int64).to(torch.int32)


@needs_ref
def test_ntt_1024_matches_reference_small_ntt():
    t = _t()
    ctx, _, _ = _keys()
    op, rt = t.HELogicOperator(ctx), R.RefTfhe()
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, P, (5, 1024), generator=gen, device="cuda", dtype=torch.int64)
    x[0] = 0
    x[1] = P - 1
    x[2] = 0
    x[2, 0] = 1
    ours, theirs = x.clone(), x.clone()
    op.ntt(ours)
    rt.ntt(theirs)
    torch.cuda.synchronize()
    assert torch.equal(ours, theirs), "forward"
    op.ntt(ours, inverse=True)
    rt.ntt(theirs, inverse=True)
    torch.cuda.synchronize()
    assert torch.equal(ours, theirs) and torch.equal(ours, x), "inverse"


@needs_ref
@pytest.mark.parametrize("gate", ["NAND", "AND", "NOR", "OR", "XNOR", "XOR", "ANDNY", "NOT"])
def test_gate_linear_part_matches_reference(gate):
    t = _t()
    ctx, _, _ = _keys()
    op, rt = t.HELogicOperator(ctx), R.RefTfhe()
    gen = torch.Generator(device="cuda").manual_seed(5)
    shape, n = 7, ctx.n_
    c1 = t.Ciphertext(ctx, _rand_i32(gen, shape, n), _rand_i32(gen, shape))
    c2 = t.Ciphertext(ctx, _rand_i32(gen, shape, n), _rand_i32(gen, shape))
    out = op.gate_linear(gate, c1, None if gate == "NOT" else c2)
    ra, rb = torch.zeros(shape, n, dtype=torch.int32, device="cuda"), torch.zeros(shape, dtype=torch.int32, device="cuda")
    rt.gate_linear(t.GATES[gate], c1.a_device_location_, c1.b_device_location_, c2.a_device_location_, c2.b_device_location_,
                   ra, rb, n, shape)
    torch.cuda.synchronize()
    assert torch.equal(out.a_device_location_, ra) and torch.equal(out.b_device_location_, rb)


@needs_ref
@pytest.mark.parametrize("kind", ["real_key", "uniform_key_words"])
def test_blind_rotation_bit_exact_vs_reference_kernels(kind):
    """HELogicOperator<TFHE>::bootstrapping: ours is ONE launch; the replay is the reference's 1 + 2*511 + 1."""
    t = _t()
    ctx, sk, bk = _keys()
    op, rt = t.HELogicOperator(ctx), R.RefTfhe()
    gen = torch.Generator(device="cuda").manual_seed(11)
    shape = 5
    a, b = _rand_i32(gen, shape, ctx.n_), _rand_i32(gen, shape)
    a[0, :7] = 0          # rotation amount 0 (skipped steps)
    a[1, 3] = -2 ** 31    # rotation by N
    b[2] = 0
    key = bk
    if kind == "uniform_key_words":
        key = t.Bootstrappingkey(ctx)
        key.boot_key_device_location_ = torch.randint(0, P, tuple(bk.boot_key_device_location_.shape), generator=gen,
                                                      device="cuda", dtype=torch.int64)
    out = op.bootstrapping(t.Ciphertext(ctx, a, b), key)
    ra = torch.zeros(shape, ctx.N_, dtype=torch.int32, device="cuda")
    rb = torch.zeros(shape, dtype=torch.int32, device="cuda")
    rt.bootstrap(a, b, ra, rb, key.boot_key_device_location_, shape)
    torch.cuda.synchronize()
    assert torch.equal(out.a_device_location_, ra), "extracted a differs"
    assert torch.equal(out.b_device_location_, rb), "extracted b differs"


@needs_ref
def test_key_switch_bit_exact_vs_reference_kernel():
    t = _t()
    ctx, sk, bk = _keys()
    op, rt = t.HELogicOperator(ctx), R.RefTfhe()
    gen = torch.Generator(device="cuda").manual_seed(13)
    shape = 6
    a, b = _rand_i32(gen, shape, ctx.N_), _rand_i32(gen, shape)
    a[0] = 0
    out = op.key_switching(t.Ciphertext(ctx, a, b), bk)
    ra, rb = torch.zeros(shape, ctx.n_, dtype=torch.int32, device="cuda"), torch.zeros(shape, dtype=torch.int32, device="cuda")
    rt.keyswitch(a, b, ra, rb, bk.switch_key_device_location_a_, bk.switch_key_device_location_b_, shape)
    torch.cuda.synchronize()
    assert torch.equal(out.a_device_location_, ra) and torch.equal(out.b_device_location_, rb)


def test_all_gates_decrypt_to_their_truth_tables():
    """test/test_tfhe_gate_boot.cpp:13-94."""
    t = _t()
    ctx, sk, bk = _keys()
    enc, dec, logic = t.HEEncryptor(ctx, sk), t.HEDecryptor(ctx, sk), t.HELogicOperator(ctx)
    rng = np.random.default_rng(9)
    size = 64
    i1, i2, cc = (rng.integers(0, 2, size).astype(bool) for _ in range(3))
    c1, c2, c3 = enc.encrypt(i1), enc.encrypt(i2), enc.encrypt(cc)
    assert dec.decrypt(c1) == list(i1)
    want = {"NAND": ~(i1 & i2), "AND": i1 & i2, "NOR": ~(i1 | i2), "OR": i1 | i2, "XNOR": ~(i1 ^ i2), "XOR": i1 ^ i2}
    for g, w in want.items():
        got = dec.decrypt(getattr(logic, g)(c1, c2, bk))
        assert got == list(w), g
    assert dec.decrypt(logic.NOT(c1)) == list(~i1)
    assert dec.decrypt(logic.MUX(c1, c2, c3, bk)) == list(np.where(cc, i1, i2))
    # gates compose: (a NAND b) XOR c, twice bootstrapped
    assert dec.decrypt(logic.XOR(logic.NAND(c1, c2, bk), c3, bk)) == list((~(i1 & i2)) ^ cc)


def test_gate_errors():
    t = _t()
    ctx, sk, bk = _keys()
    enc, logic = t.HEEncryptor(ctx, sk), t.HELogicOperator(ctx)
    with pytest.raises(t.HeonError):
        logic.AND(enc.encrypt([True, False]), enc.encrypt([True]), bk)
    with pytest.raises(t.HeonError):
        t.HEEncryptor(ctx, t.Secretkey(ctx))


def test_blind_rotation_and_key_switch_match_cpu_oracle():
    """The CUDA path against the CPU restatement (oracle/heon_oracle.c) on the seeded golden inputs."""
    from oracle import oracle as O
    from tests.tfhe_common import golden_inputs
    t = _t()
    ctx, _, _ = _keys()
    op, orc = t.HELogicOperator(ctx), O.TfheOracle()
    g = golden_inputs()
    dev = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()
    key = t.Bootstrappingkey(ctx)
    key.boot_key_device_location_ = dev(g["bk"])
    key.switch_key_device_location_a_, key.switch_key_device_location_b_ = dev(g["ks_a"]), dev(g["ks_b"])
    out = op.bootstrapping(t.Ciphertext(ctx, dev(g["boot_a"]), dev(g["boot_b"])), key)
    oa, ob = orc.bootstrap(g["boot_a"], g["boot_b"], g["bk"])
    assert np.array_equal(out.a_device_location_.cpu().numpy(), oa) and np.array_equal(out.b_device_location_.cpu().numpy(), ob)
    ks = op.key_switching(t.Ciphertext(ctx, dev(g["ks_in_a"]), dev(g["ks_in_b"])), key)
    ka, kb = orc.keyswitch(g["ks_in_a"], g["ks_in_b"], g["ks_a"], g["ks_b"])
    assert np.array_equal(ks.a_device_location_.cpu().numpy(), ka) and np.array_equal(ks.b_device_location_.cpu().numpy(), kb)
    for gate in range(8):
        c1, c2 = t.Ciphertext(ctx, dev(g["a1"]), dev(g["b1"])), t.Ciphertext(ctx, dev(g["a2"]), dev(g["b2"]))
        name = [k for k, v in t.GATES.items() if v == gate][0]
        got = op.gate_linear(name, c1, None if gate == 7 else c2)
        wa, wb = orc.gate_linear(gate, g["a1"], g["b1"], g["a2"], g["b2"])
        assert np.array_equal(got.a_device_location_.cpu().numpy(), wa) and np.array_equal(got.b_device_location_.cpu().numpy(), wb)
