"""GPU tests of the tcgen05 NT-GEMM engine (kfb_gemm_nt) against fp64 torch on the SAME bf16 hi/lo
operands, so the only admissible differences are the dropped lo*lo term (~2^-32 relative) and fp32
accumulation order."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine():
    from kronfluence_b200 import engine

    engine.require_device()
    return engine


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def _operands(engine, batch_a, batch_b, m, n, k, precision, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(batch_a, m, k, device="cuda", generator=gen)
    b = torch.randn(batch_b, n, k, device="cuda", generator=gen)
    sa = engine.split_from_tensor(a, precision)
    sb = engine.split_from_tensor(b, precision)
    ref = torch.matmul(sa.to_float().double(), sb.to_float().double().transpose(1, 2))
    return sa, sb, ref


SHAPES = [
    (1, 1, 128, 256, 64),
    (1, 1, 128, 256, 2048),
    (1, 1, 200, 300, 100),      # ragged everything
    (1, 1, 7, 10, 5),           # tiny
    (3, 3, 130, 70, 96),        # batched, N<128
    (1, 4, 256, 512, 257),      # A broadcast, odd K
    (4, 1, 64, 129, 40),        # B broadcast
    (1, 1, 1000, 1000, 1),      # rank-1 (S=1 outer products)
]


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_store_f32(shape, precision):
    engine = _engine()
    ba, bb, m, n, k = shape
    sa, sb, ref = _operands(engine, ba, bb, m, n, k, precision)
    batch = max(ba, bb)
    out = torch.full((batch, m, n), float("nan"), device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=n, out_batch_stride=m * n, alpha=1.0)
    engine.gemm_nt(sa, sb, epi, precision)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert _rel(out, ref.expand(batch, m, n)) < 1e-5


@pytest.mark.parametrize("simt", [False, True])
def test_store_variants(simt):
    engine = _engine()
    lib = engine.load_library()
    ba, bb, m, n, k = 2, 2, 150, 200, 72
    sa, sb, ref = _operands(engine, ba, bb, m, n, k, 0, seed=3)
    mul = torch.rand(m, n, device="cuda") + 0.5
    out_t = torch.ones(2, n, m, device="cuda")
    dst = engine.Split(n, m, 2, device="cuda")
    dst.storage.fill_(float("nan"))
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out_t.data_ptr(), ldo=m, out_batch_stride=m * n,
                             out_split=dst.struct(), mul=mul.data_ptr(), ldmul=n, transpose_out=1, square=1,
                             accumulate=1, alpha=0.5)
    try:
        if simt:
            lib.kfb_set_gemm_backend(1)
        engine.gemm_nt(sa, sb, epi, 0)
        torch.cuda.synchronize()
    finally:
        lib.kfb_set_gemm_backend(0)
    want = ((0.5 * ref * mul.double()) ** 2).transpose(1, 2)
    assert _rel(out_t - 1.0, want) < 1e-5
    assert _rel(dst.to_float(), want) < 1e-4  # bf16 hi+lo holds ~16 mantissa bits

    # non-transposed split output, padding must be zero-filled
    dst2 = engine.Split(m, n, 2, device="cuda")
    dst2.storage.fill_(float("nan"))
    epi2 = engine.KfbEpilogue(kind=engine.EPI_STORE, out_split=dst2.struct(), alpha=1.0)
    engine.gemm_nt(sa, sb, epi2, 0)
    torch.cuda.synchronize()
    assert _rel(dst2.to_float(), ref) < 1e-4
    assert (dst2.storage[:, :, :, n:].float() == 0).all()


def test_split_k_accumulate():
    engine = _engine()
    sa, sb, ref = _operands(engine, 1, 1, 96, 80, 8192, 0, seed=5)
    out = torch.ones(96, 80, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=80, out_batch_stride=0,
                             accumulate=1, alpha=2.0)
    engine.gemm_nt(sa, sb, epi, 0)
    torch.cuda.synchronize()
    assert _rel(out - 1.0, 2.0 * ref[0]) < 1e-5


@pytest.mark.parametrize("shape", [(5, 300, 520, 130), (2, 128, 256, 64), (3, 17, 10, 33), (4, 1000, 1024, 1025)])
def test_rowdot(shape):
    engine = _engine()
    q, t, n, k = shape
    sa, sb, ref = _operands(engine, 1, q, t, n, k, 0, seed=7)  # A = train rows (shared), B = P[q]
    g = torch.randn(t, n, device="cuda")
    out = torch.ones(q, t + 3, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=out[:, 1:].data_ptr(), out_batch_stride=t + 3,
                             g=g.data_ptr(), ldg=n, alpha=1.5, accumulate=1)
    engine.gemm_nt(sa, sb, epi, 0)
    torch.cuda.synchronize()
    want = 1.5 * (ref * g.double().unsqueeze(0)).sum(-1)
    assert _rel(out[:, 1 : t + 1] - 1.0, want) < 2e-5
    assert (out[:, 0] == 1).all() and (out[:, t + 1 :] == 1).all()


@pytest.mark.parametrize("shape", [(5, 330, 3, 70, 11), (2, 256, 64, 128, 32), (7, 100, 40, 33, 1)])
def test_rowdot_batched_factor_and_row_groups(shape):
    """The low-rank scoring epilogue: every batch entry reads its own column block of the factor matrix
    (g_batch_stride) and the rows of a group (the tokens of an example) are summed into one output."""
    engine = _engine()
    q, t, n, k, group = shape
    sa, sb, ref = _operands(engine, 1, q, t, n, k, 0, seed=5)  # A = token rows (shared), B = left factors of query q
    ldy = (q * n + 3) // 4 * 4
    y = torch.randn(t, ldy, device="cuda")
    groups = (t + group - 1) // group
    out = torch.ones(q, groups + 2, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=out[:, 1:].data_ptr(), out_batch_stride=groups + 2,
                             g=y.data_ptr(), ldg=ldy, g_batch_stride=n, row_group=group, alpha=0.5, accumulate=1)
    engine.gemm_nt(sa, sb, epi, 0)
    torch.cuda.synchronize()
    yq = y[:, : q * n].double().reshape(t, q, n).permute(1, 0, 2)  # [q, t, n]
    per_row = 0.5 * (ref * yq).sum(-1)  # [q, t]
    pad = groups * group - t
    want = torch.nn.functional.pad(per_row, (0, pad)).reshape(q, groups, group).sum(-1)
    assert _rel(out[:, 1 : groups + 1] - 1.0, want) < 2e-5
    assert (out[:, 0] == 1).all() and (out[:, groups + 1 :] == 1).all()


@pytest.mark.parametrize("shape", [(1, 2048, 4096, 512), (3, 200, 1300, 96)])
def test_rowdot_few_row_blocks_overwrite(shape):
    """Few row blocks: the n-tiles of a block are spread over several units (atomics into a zeroed output)."""
    engine = _engine()
    q, t, n, k = shape
    sa, sb, ref = _operands(engine, 1, q, t, n, k, 0, seed=11)
    g = torch.randn(t, n, device="cuda")
    out = torch.full((q, t + 5), 7.0, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=out[:, 2:].data_ptr(), out_batch_stride=t + 5,
                             g=g.data_ptr(), ldg=n, alpha=0.5, accumulate=0)
    engine.gemm_nt(sa, sb, epi, 0)
    torch.cuda.synchronize()
    want = 0.5 * (ref * g.double().unsqueeze(0)).sum(-1)
    assert _rel(out[:, 2 : t + 2], want) < 2e-5
    assert (out[:, :2] == 7).all() and (out[:, t + 2 :] == 7).all()


@pytest.mark.parametrize("shape", [(9, 200, 150, 40), (300, 64, 128, 16), (2, 769, 768, 128)])
def test_sqacc(shape):
    engine = _engine()
    b, m, n, k = shape
    sa, sb, ref = _operands(engine, b, b, m, n, k, 0, seed=11)
    out = torch.ones(m, n, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_SQACC, out_f32=out.data_ptr(), ldo=n, alpha=0.25)
    engine.gemm_nt(sa, sb, epi, 0)
    torch.cuda.synchronize()
    want = 0.25 * (ref ** 2).sum(0)
    assert _rel(out - 1.0, want) < 2e-5


def test_large_persistent():
    """More tiles than SMs, several k-blocks, both accumulator stages and many ring wraps."""
    engine = _engine()
    sa, sb, ref = _operands(engine, 1, 1, 4096, 2048, 1024, 0, seed=13)
    out = torch.empty(4096, 2048, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=2048, alpha=1.0)
    engine.gemm_nt(sa, sb, epi, 0)
    torch.cuda.synchronize()
    assert _rel(out, ref[0]) < 1e-5


@pytest.mark.parametrize("shape", [(1, 200, 150, 9000), (2, 64, 300, 5000)])
def test_long_contraction_passes(shape):
    """K > 2048: the contraction is cut into TMEM passes summed in fp32 registers (and split across
    CTAs when accumulating), keeping the tensor core's truncating accumulation short."""
    engine = _engine()
    b, m, n, k = shape
    sa, sb, ref = _operands(engine, b, b, m, n, k, 0, seed=17)
    out = torch.full((b, m, n), float("nan"), device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=n, out_batch_stride=m * n, alpha=1.0)
    engine.gemm_nt(sa, sb, epi, 0)
    acc = torch.ones(b, m, n, device="cuda")
    epi2 = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=acc.data_ptr(), ldo=n, out_batch_stride=m * n,
                              alpha=1.0, accumulate=1)
    engine.gemm_nt(sa, sb, epi2, 0)
    dst = engine.Split(m, n, b, device="cuda")
    epi3 = engine.KfbEpilogue(kind=engine.EPI_STORE, out_split=dst.struct(), alpha=1.0)
    engine.gemm_nt(sa, sb, epi3, 0)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-5
    assert _rel(acc - 1.0, ref) < 1e-5
    assert _rel(dst.to_float(), ref) < 1e-4


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("shape", [(4097, 1000), (768, 9000), (300, 64), (100, 5000), (1025, 2048), (577, 70)])
def test_symmetric_accumulate(shape, precision):
    """SYRK mode (covariance, tracker/factor.py:85-93): C += alpha X^T X with only the upper-triangular tiles
    computed and mirrored by the epilogue; long contractions are cut into atomically combined passes."""
    engine = _engine()
    d, k = shape
    gen = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(1, d, k, device="cuda", generator=gen)
    sx = engine.split_from_tensor(x, precision)
    xf = sx.to_float().double()[0]
    ref = xf @ xf.t()
    out = torch.ones(d, d, device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=d, out_batch_stride=0, accumulate=1,
                             alpha=0.5, symmetric=1)
    engine.gemm_nt(sx, sx, epi, precision)
    engine.gemm_nt(sx, sx, epi, precision)  # accumulates a second time
    torch.cuda.synchronize()
    got = out - 1.0
    # bf16 mode runs 4x longer TMEM passes (its operands are only good to 2^-9): allow their accumulation bias
    tol = 1e-5 if precision == 0 else 3e-5
    assert _rel(got, ref) < tol
    assert _rel(got.t(), ref) < tol
    # mirrored tiles: same values up to the order of the passes' atomic adds
    assert (got - got.t()).abs().max().item() <= 2e-6 * got.abs().max().item()


@pytest.mark.parametrize("scale", [1.0, 3e-7, 5e4])
@pytest.mark.parametrize("shape", [(1, 1, 200, 150, 700), (3, 1, 64, 300, 4097), (1, 1, 130, 40, 30), (1, 1, 700, 520, 769)])
def test_strict_precision(shape, scale):
    """KFB_PREC_STRICT: scaled FP16 hi/lo planes, 3 MMAs, TMEM drained every 64 contraction elements.  Checked against
    float64 on the ORIGINAL float32 inputs: the whole error (split + truncating accumulation) is below 1e-6."""
    engine = _engine()
    ba, bb, m, n, k = shape
    gen = torch.Generator(device="cuda").manual_seed(23)
    a = torch.randn(ba, m, k, device="cuda", generator=gen) * scale  # the operand scale must not matter
    b = torch.randn(bb, n, k, device="cuda", generator=gen)
    sa = engine.split_from_tensor(a, engine.PREC_STRICT)
    sb = engine.split_from_tensor(b, engine.PREC_STRICT)
    assert (sa.to_float() - a.double()).abs().max() <= 3e-7 * a.abs().max()
    rel_elem = ((sa.to_float() - a.double()).abs() / a.double().abs().clamp_min(1e-30))
    assert rel_elem[a.abs() > 1e-3 * a.abs().max()].max() <= 2.5e-7  # 22 mantissa bits where it matters
    ref = torch.matmul(a.double(), b.double().transpose(1, 2))
    batch = max(ba, bb)
    out = torch.full((batch, m, n), float("nan"), device="cuda")
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=n, out_batch_stride=m * n, alpha=1.0)
    engine.gemm_nt(sa, sb, epi, engine.PREC_STRICT)
    dst = engine.Split(n, m, batch, device="cuda")
    epi2 = engine.KfbEpilogue(kind=engine.EPI_STORE, out_split=dst.struct(), alpha=1.0, transpose_out=1)
    engine.gemm_nt(sa, sb, epi2, engine.PREC_STRICT)
    torch.cuda.synchronize()
    err = _rel(out, ref.expand(batch, m, n))
    assert err < 2e-6, err
    assert _rel(dst.to_float(), ref.expand(batch, m, n).transpose(1, 2)) < 1e-4
