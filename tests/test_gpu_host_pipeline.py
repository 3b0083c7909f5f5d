"""GPU: train_utils.HostViewPipeline (host batches in, rendered views out, copies overlapped with the kernels on three
streams) returns, for every batch, exactly what predict_ldi + forward_splat return for that batch alone."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_direct_calls():
    from lsi.geometry import ldi as ldi_utils
    from lsi.nnutils import helpers, nets, train_utils
    dev = torch.device('cuda')
    B, H, W, L = 2, 128, 128, 2
    opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W)
    store = nets.ParamStore(device=dev, seed=0)
    kw = dict(compose_layers=True, trg_downsampling=1.0, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50.0)
    rs = np.random.RandomState(0)
    K = np.array([[W * 0.58, 0, W * 0.49], [0, H * 1.92, H * 0.46], [0, 0, 1]], dtype=np.float32)
    batches = []
    for k in range(5):
        t = np.zeros((B, 3), dtype=np.float32)
        t[:, 0] = -0.5 + 0.1 * k
        batches.append({'img': torch.tensor(rs.uniform(0, 1, (B, H, W, 3)).astype(np.float32)).pin_memory(),
                        'k_s': torch.tensor(np.stack([K] * B)).pin_memory(), 'k_t': torch.tensor(np.stack([K] * B)).pin_memory(),
                        'rot': torch.tensor(np.stack([np.eye(3, dtype=np.float32)] * B)).pin_memory(),
                        't': torch.tensor(t).pin_memory()})
    with torch.no_grad():
        train_utils.predict_ldi(batches[0]['img'].to(dev), opts, store, reuse=False)
    pc = helpers.pixel_coords(B, H, W, _device=dev)
    direct = []
    for b in batches:
        with torch.no_grad():
            ldi = train_utils.predict_ldi(b['img'].to(dev), opts, store, reuse=True)
            img, wts = ldi_utils.forward_splat(tuple(ldi), pc, b['k_s'].to(dev), b['k_t'].to(dev), b['rot'].to(dev),
                                               b['t'].to(dev), **kw)
        direct.append((img.cpu().clone(), wts.cpu().clone()))
    got = {}
    pipe = train_utils.HostViewPipeline(opts, store, kw, B, H, W, dev, depth=2)
    pipe.run(batches, on_result=lambda k, im, wt: got.__setitem__(k, (im.clone(), wt.clone())))
    assert sorted(got) == list(range(len(batches)))
    assert pipe.d2h_bytes() == 4 * B * H * W * 4 and pipe.h2d_bytes(batches[0]) == 4 * (B * H * W * 3 + B * 30)
    for k, (img, wts) in enumerate(direct):
        # same kernels on the same inputs; the splat's floating-point reductions are order-dependent, hence a tolerance
        assert float((got[k][0] - img).abs().max()) < 1e-4, k
        assert float((got[k][1] - wts).abs().max()) < 1e-4 * float(wts.abs().max()), k
    assert float((direct[0][0] - direct[4][0]).abs().max()) > 1e-3      # the batches really differ


def test_pipeline_uint8_upload_matches_device_scaling():
    """u8_input=True: the images cross the bus as uint8 (a quarter of the bytes) and are scaled to [0,1] on the device; the result
    equals the float pipeline run on the same 8-bit data / 255, at depth 3."""
    from lsi.nnutils import nets, train_utils
    dev = torch.device('cuda')
    B, H, W, L = 2, 128, 128, 2
    opts = train_utils.default_opts(dataset='kitti', n_layers=L, batch_size=B, img_height=H, img_width=W)
    store = nets.ParamStore(device=dev, seed=0)
    kw = dict(compose_layers=True, trg_downsampling=1.0, bg_layer_disp=1e-3, max_disp=0.4, zbuf_scale=50.0)
    rs = np.random.RandomState(1)
    K = np.array([[W * 0.58, 0, W * 0.49], [0, H * 1.92, H * 0.46], [0, 0, 1]], dtype=np.float32)
    cams = {'k_s': torch.tensor(np.stack([K] * B)).pin_memory(), 'k_t': torch.tensor(np.stack([K] * B)).pin_memory(),
            'rot': torch.tensor(np.stack([np.eye(3, dtype=np.float32)] * B)).pin_memory(),
            't': torch.tensor(np.tile(np.array([[-0.5, 0, 0]], dtype=np.float32), (B, 1))).pin_memory()}
    u8 = [rs.randint(0, 256, (B, H, W, 3)).astype(np.uint8) for _ in range(4)]
    b8 = [dict(cams, img=torch.tensor(x).pin_memory()) for x in u8]
    bf = [dict(cams, img=torch.tensor(x.astype(np.float32) * np.float32(1.0 / 255.0)).pin_memory()) for x in u8]
    with torch.no_grad():
        train_utils.predict_ldi(bf[0]['img'].to(dev), opts, store, reuse=False)
    out8, outf = {}, {}
    p8 = train_utils.HostViewPipeline(opts, store, kw, B, H, W, dev, depth=3, u8_input=True)
    p8.run(b8, on_result=lambda k, im, wt: out8.__setitem__(k, im.clone()))
    pf = train_utils.HostViewPipeline(opts, store, kw, B, H, W, dev, depth=2)
    pf.run(bf, on_result=lambda k, im, wt: outf.__setitem__(k, im.clone()))
    assert p8.h2d_bytes(b8[0]) == B * H * W * 3 + 4 * B * 30 and pf.h2d_bytes(bf[0]) == 4 * (B * H * W * 3 + B * 30)
    for k in range(4):
        assert float((out8[k] - outf[k]).abs().max()) < 2e-3, k      # same 8-bit data; x * (1/255) on either side, then the whole network
    assert train_utils.bind_to_gpu_numa_node(0) in (None,) + tuple(range(16))
