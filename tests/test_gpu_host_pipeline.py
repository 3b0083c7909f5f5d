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
