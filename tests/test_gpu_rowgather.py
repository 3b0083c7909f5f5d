"""Row-gather forward splat (csrc/render_rowgather.cuh): the default kernel for the rectified pose class (ldi.py:71-182 with
R = I, translation in the image plane: the KITTI stereo configurations).  Checked against the CPU oracle at small sizes and
against the plain one-thread-per-pixel global-atomic kernel (`_variant=1`, the closest restatement of the reference's
scatter_nd semantics) at sizes the oracle would take minutes for: both layouts, masks, both compose modes, target
down-sampling, widths that are not multiples of 32, rows wider than 1024 pixels (512-thread instantiation), checkerboard-noise
disparities (long per-cell lists), mixed batches (some images outside the class fall through to the reduction kernels) and the
`variant 5` hint the Python mirror selects once it has seen the class flags of a camera set."""
import numpy as np
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    assert torch.cuda.is_available()
    from lsi import _b200
    from lsi.geometry import ldi
    from lsi.nnutils import helpers
    _b200.lib()
    return ldi, helpers


def _scene(L, B, H, W, seed, noisy=False, tx=-0.5327, ty=0.0, rotate=(), max_disp=0.4):
    g = torch.Generator().manual_seed(seed)
    tex = torch.rand(L, B, H, W, 3, generator=g)
    if noisy:
        disp = torch.rand(L, B, H, W, 1, generator=g) * max_disp
    else:
        yy = torch.linspace(0, 1, H).view(1, 1, H, 1, 1)
        xx = torch.linspace(0, 6.28, W).view(1, 1, 1, W, 1)
        base = max_disp * torch.clamp((yy - 0.45) / 0.55, 0.02, 1) + 0.02 * torch.sin(3 * xx + yy * 5)
        disp = torch.cat([base * f for f in np.linspace(1.0, 0.4, L)], dim=0).expand(L, B, H, W, 1).contiguous()
        disp = disp.clamp(1e-3, max_disp)
    mask = torch.rand(L, B, H, W, 1, generator=g)
    fx, fy = 721.54 * W / 1242.0, 721.54 * H / 375.0
    k = torch.tensor([[fx, 0, 609.56 * W / 1242.0], [0, fy, 172.85 * H / 375.0], [0, 0, 1]], dtype=torch.float32)
    k = k.expand(B, 3, 3).contiguous()
    rot = torch.eye(3).expand(B, 3, 3).contiguous().clone()
    for b in rotate:              # a small rotation about the optical axis + y: outside the rectified class
        a = 0.03
        rot[b] = torch.tensor([[np.cos(a), -np.sin(a), 0.01], [np.sin(a), np.cos(a), 0], [-0.01, 0, 1]], dtype=torch.float32)
    t = torch.tensor([tx, ty, 0.0]).expand(B, 3).contiguous().view(B, 3, 1)
    return tex, mask, disp, k, k.clone(), rot, t


def _run(ldi, helpers, sc, use_mask, packed, variant, **kw):
    tex, mask, disp, k_s, k_t, rot, t = [x.cuda() for x in sc]
    L, B, H, W, _ = tex.shape
    if packed:
        pk = torch.cat([tex, disp], dim=-1).contiguous()
        tex, disp = pk[..., :3], pk[..., 3:]
    pc = helpers.pixel_coords(B, H, W)
    return ldi.forward_splat((tex, mask if use_mask else None, disp), pc, k_s, k_t, rot, t, _variant=variant, **kw)


CASES = [
    # L, B, H, W, ds, compose, mask, packed, noisy
    (4, 2, 16, 832, 1.0, True, False, True, False),      # the headline shape of a row, packed head-output layout
    (4, 2, 16, 832, 1.0, True, False, False, False),     # planar tensors
    (2, 3, 12, 416, 0.5, True, True, False, False),      # training setting: two target rows per source row, masks
    (3, 2, 12, 100, 1.0, False, True, True, True),       # per-layer outputs, width not a multiple of 32, noise
    (4, 1, 8, 1664, 1.0, True, False, True, True),       # 512-thread instantiation, noise
    (5, 2, 8, 1664, 0.5, False, True, False, False),     # config-5 shape of a row, per layer, ds = 0.5
    (1, 2, 6, 36, 1.0, True, True, False, True),         # tiny
    (4, 2, 16, 256, 0.25, True, False, True, False),     # four-fold down-sampling (several source rows per target row)
]


@pytest.mark.parametrize('case', CASES)
def test_rowgather_matches_atomic_kernel(mods, case):
    ldi, helpers = mods
    L, B, H, W, ds, compose, use_mask, packed, noisy = case
    sc = _scene(L, B, H, W, seed=hash(case) % 1000, noisy=noisy)
    kw = dict(compose_layers=compose, trg_downsampling=ds, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    got = _run(ldi, helpers, sc, use_mask, packed, 0, **kw)
    ref = _run(ldi, helpers, sc, use_mask, packed, 1, **kw)
    for a, b in zip(got, ref):
        assert torch.isfinite(a).all()
        assert rel_err(a.cpu(), b.cpu()) < 2e-5


def test_rowgather_vertical_shift_and_oracle(mods):
    """A fractional vertical translation (two target rows per source row at ds = 1) against the CPU oracle."""
    ldi, helpers = mods
    from oracle import lsi_oracle as O
    L, B, H, W = 2, 2, 24, 64
    sc = _scene(L, B, H, W, seed=3, ty=0.07, max_disp=1.0)
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=0.2, max_disp=1.0, zbuf_scale=50)
    tex, mask, disp, k_s, k_t, rot, t = sc
    ref_img, ref_wts = O.forward_splat((tex, mask, disp), O.pixel_coords(B, H, W), k_s, k_t, rot, t, **kw)
    img, wts = _run(ldi, helpers, sc, True, False, 0, **kw)
    assert rel_err(img.cpu(), ref_img) < 1e-4
    assert rel_err(wts.cpu(), ref_wts) < 1e-4


def test_rowgather_mixed_batch_and_hint(mods):
    """Images outside the class are rendered by the reduction kernels in the same call; the all-rectified hint (variant 5) is
    chosen by the Python mirror only from the device's own flags of the same camera tensors and gives the same result."""
    ldi, helpers = mods
    L, B, H, W = 3, 4, 16, 128
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    sc = _scene(L, B, H, W, seed=5, rotate=(1, 3))
    got = _run(ldi, helpers, sc, True, True, 0, **kw)
    ref = _run(ldi, helpers, sc, True, True, 1, **kw)
    for a, b in zip(got, ref):
        assert rel_err(a.cpu(), b.cpu()) < 2e-5
    # same camera tensor objects over several calls: flags are read back asynchronously, later calls take variant 5
    tex, mask, disp, k_s, k_t, rot, t = [x.cuda() for x in _scene(L, B, H, W, seed=6)]
    pc = helpers.pixel_coords(B, H, W)
    outs = []
    for it in range(4):
        outs.append(ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, **kw))
        torch.cuda.synchronize()
    e, cls = ldi._POSE_CACHE.lookup((k_s, k_t, rot, t))
    assert cls == 'all' and e['state'] == 'all'
    ref = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, _variant=1, **kw)
    for o in outs:
        for a, b in zip(o, ref):
            assert rel_err(a.cpu(), b.cpu()) < 2e-5
    # an in-place change of the cameras (now outside the class) must drop the hint
    rot.copy_(_scene(L, B, H, W, seed=6, rotate=(0,))[5].cuda())
    e, cls = ldi._POSE_CACHE.lookup((k_s, k_t, rot, t))
    assert cls is None
    got = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, **kw)
    ref = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, _variant=1, **kw)
    for a, b in zip(got, ref):
        assert torch.isfinite(a).all()
        assert rel_err(a.cpu(), b.cpu()) < 2e-5


def test_rowgather_empty_rows_and_loud_misuse(mods):
    """A vertical offset leaves target rows that no source row reaches (background only: the producer's data-less items); and the
    all-rectified hint forced onto a batch that is NOT in the class must fail loudly (NaNs), never silently."""
    ldi, helpers = mods
    L, B, H, W = 2, 2, 32, 96
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    # pure vertical offset through the principal point (still rectified: y' = y + const): rows shifted out of the image
    tex, mask, disp, k_s, k_t, rot, t = [x.cuda() for x in _scene(L, B, H, W, seed=22)]
    k_t = k_t.clone(); k_t[:, 1, 2] += 9.5
    pc = helpers.pixel_coords(B, H, W)
    got = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, **kw)
    ref = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, _variant=1, **kw)
    for a, b in zip(got, ref):
        assert rel_err(a.cpu(), b.cpu()) < 2e-5
    assert torch.allclose(got[0][0, :, :8], torch.ones_like(got[0][0, :, :8]))      # the first rows see only the white canvas
    # misuse: variant 5 on rotated cameras
    sc = _scene(L, B, H, W, seed=23, rotate=(1,))
    img, wts = _run(ldi, helpers, sc, True, True, 5, **kw)
    assert torch.isfinite(img[0, 0]).all() and torch.isnan(img[0, 1]).all() and torch.isnan(wts[0, 1]).all()


def test_hint_on_a_shape_the_kernel_does_not_take(mods):
    """Rows wider than 2048 pixels are outside the row-gather kernel; the all-rectified hint the mirror passes after the first sight of
    the cameras must then fall back to the default path (streaming kernel), not fail."""
    ldi, helpers = mods
    L, B, H, W = 2, 1, 4, 2304
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    tex, mask, disp, k_s, k_t, rot, t = [x.cuda() for x in _scene(L, B, H, W, seed=31)]
    pc = helpers.pixel_coords(B, H, W)
    ref = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, _variant=1, **kw)
    for it in range(3):
        got = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, **kw)
        torch.cuda.synchronize()
        for a, b in zip(got, ref):
            assert torch.isfinite(a).all() and rel_err(a.cpu(), b.cpu()) < 2e-5
    assert ldi._POSE_CACHE.lookup((k_s, k_t, rot, t))[1] == 'all'


def test_no_image_in_class_hint(mods):
    """General poses for the whole batch: after the first sight the mirror passes variant 6 (no row-gather launch); same result."""
    ldi, helpers = mods
    L, B, H, W = 2, 3, 16, 64
    kw = dict(compose_layers=True, trg_downsampling=1, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    tex, mask, disp, k_s, k_t, rot, t = [x.cuda() for x in _scene(L, B, H, W, seed=11, rotate=(0, 1, 2))]
    pc = helpers.pixel_coords(B, H, W)
    outs = []
    for it in range(3):
        outs.append(ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, **kw))
        torch.cuda.synchronize()
    assert ldi._POSE_CACHE.lookup((k_s, k_t, rot, t))[1] == 'none'
    ref = ldi.forward_splat((tex, mask, disp), pc, k_s, k_t, rot, t, _variant=1, **kw)
    for o in outs:
        for a, b in zip(o, ref):
            assert rel_err(a.cpu(), b.cpu()) < 2e-5


def test_rowgather_gradients_unchanged(mods):
    """The backward kernels read the saved forward outputs: gradients through the new forward equal those through the atomic one."""
    ldi, helpers = mods
    L, B, H, W = 2, 2, 16, 96
    sc = _scene(L, B, H, W, seed=9)
    kw = dict(compose_layers=True, trg_downsampling=0.5, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50)
    grads = []
    for variant in (0, 1):
        tex, mask, disp, k_s, k_t, rot, t = [x.cuda() for x in sc]
        leaves = [x.requires_grad_(True) for x in (tex, mask, disp)]
        pc = helpers.pixel_coords(B, H, W)
        img, wts = ldi.forward_splat(tuple(leaves), pc, k_s, k_t, rot, t, _variant=variant, **kw)
        g = torch.autograd.grad((img * torch.linspace(0, 1, img.numel(), device='cuda').view_as(img)).sum() + wts.sum(), leaves)
        grads.append(g)
    for a, b in zip(*grads):
        assert rel_err(a.cpu(), b.cpu()) < 5e-5
