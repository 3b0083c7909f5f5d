"""Mirror of lsi/data/syntheticPlanes/data.py (reference tree): the synthetic planar-room data generator, GPU-native.

Same classes and methods -- sample_views, WorldGenerator, Renderer, DataLoader -- and the same world statistics (box extent,
billboard placement, view sampling, camera intrinsics).  What differs, and why:

  * textures: the reference pastes PASCAL object crops and SUN backgrounds (utils.QueuedRandomTextureLoader, TF file queues);
    neither dataset is available, so textures are procedural (band-limited colour noise; objects carry a soft super-ellipse alpha
    mask whose aspect plays the role of the crop's aspect) and are generated on the GPU (lsi_b200_procedural_texture);
  * rendering: the reference builds a TF graph of per-plane warps + composition and runs it in a second session, one sample at a
    time (data.py:293-450, 598-619).  Here DataLoader.forward renders the source and target views of the WHOLE batch (and, with
    synth_dl_eval_data, their foreground / background disparities and background-only images) with one fused kernel launch per
    kind of output (lsi_b200_render_planes): per target pixel, loop over the planes -- homography, bilinear texture + mask sample,
    plane disparity, layer log-probability, running arg-max; no warped layer is ever materialised;
  * randomness: a numpy RandomState owned by the generator (seed 0, train_utils.py:158-160) instead of the global numpy state.

forward(bs) returns the reference's tuple as float32 CUDA tensors.
"""
import numpy as np
import torch

from lsi import _b200
from lsi.data.syntheticPlanes import utils

MIN_DISP, DEPTH_SOFTMAX_TEMP = 2e-1, 0.4        # data.py:376-380: the renderer's fixed composition constants
EXTENT = (-0.7, -0.5, 2.0, 0.7, 0.5, 3.5)       # data.py:236-243: the room box
N_WAVES = 6


def sample_views(nviews, _rs=None):
    """data.py:29-52 -- camera on the z = 0 plane within +-0.5, looking at a point with z in [3, 3.5]: list of (rot, trans)."""
    rs = _rs or np.random
    out = []
    for _ in range(nviews):
        cam = np.array([rs.uniform(-0.5, 0.5), rs.uniform(-0.5, 0.5), 0.0]).reshape(3, 1)
        lookat = np.array([rs.uniform(-0.5, 0.5), rs.uniform(-0.5, 0.5), rs.uniform(3.0, 3.5)]).reshape(3, 1)
        rot = utils.lookat_rotation(lookat - cam)
        out.append((rot, -rot @ cam))
    return out


class WorldGenerator(object):
    """data.py:55-291 -- random box world: n_box_planes textured faces of the room + n_obj_min..n_obj_max almost fronto-parallel
    billboards standing on the floor (transparent dummies fill up to n_obj_max)."""

    def __init__(self, bg_tex_dir=None, obj_tex_dir=None, h=400, w=400, n_obj_max=4, n_obj_min=1, n_box_planes=5, ext_obj='.png',
                 ext_bg='.jpg', split='all', _seed=0, _device='cuda'):
        self.h, self.w = h, w
        self.n_obj_min, self.n_obj_max, self.n_box_planes = n_obj_min, n_obj_max, n_box_planes
        self.bs = n_box_planes + n_obj_max
        assert self.bs > 0
        self.rs = np.random.RandomState({'all': 0, 'train': 1, 'val': 2, 'test': 3}.get(split, 0) * 7919 + _seed)
        self.device = torch.device(_device)

    def dummy_obj_plane(self, z_max):
        """data.py:120-143 -- a unit plane at the back of the box; its texture is fully transparent."""
        return {'pt': np.array([0.0, 0.0, z_max]), 'x_dir': np.array([1.0, 0, 0]), 'y_dir': np.array([0, 1.0, 0]), 'w': 1, 'h': 1,
                'off_x': 0.5, 'off_y': 0.5}

    def random_obj_plane(self, extent, aspect, fixed_plane=None):
        """data.py:145-205 -- a billboard of (perturbed) aspect h/w standing on the floor of the box."""
        rs = self.rs
        aspect = float(np.exp(np.log(aspect) + rs.uniform(-0.2, 0.2)))
        w_box, h_box, d_box = extent[3] - extent[0], extent[4] - extent[1], extent[5] - extent[2]
        if aspect < h_box / w_box:                       # width is the bottleneck
            w_obj = rs.uniform(0.4, 0.6) * w_box
            h_obj = w_obj * aspect
        else:
            h_obj = rs.uniform(0.4, 0.6) * h_box
            w_obj = h_obj / aspect
        if fixed_plane is not None:
            cx = extent[0] + 0.25 * w_box + 0.25 * fixed_plane * w_box
            cz = extent[2] + 0.2 * fixed_plane * d_box
        else:
            cx = extent[0] + w_box * rs.uniform(0.1, 0.9 - w_obj / w_box) + 0.5 * w_obj
            cz = extent[2] + 0.5 * rs.uniform(0, d_box)
        return {'pt': np.array([cx, extent[4], cz]), 'x_dir': np.array([1.0, 0, 0]), 'y_dir': np.array([0, 1.0, 0]), 'w': w_obj,
                'h': h_obj, 'off_x': 0.5, 'off_y': 1}

    def layout(self):
        """The geometric half of forward(): (rot_w2s, t_w2s, k_w, n_hat_w, a_w) as numpy arrays and the texture recipe
        (wave parameters [bs,3,K,4], kind [bs]: 0 wall, 1 object, 2 transparent dummy)."""
        rs, bs = self.rs, self.bs
        planes = utils.box_planes(EXTENT)[0:self.n_box_planes]
        n_obj = rs.randint(self.n_obj_min, self.n_obj_max + 1)
        kind = np.zeros(bs, dtype=np.int32)
        for ix in range(self.n_obj_max):
            if ix < n_obj:
                aspect_tex = float(np.exp(rs.uniform(-0.5, 0.5)))       # stands for the crop's height / width
                planes.append(self.random_obj_plane(EXTENT, aspect_tex, fixed_plane=ix))
                kind[self.n_box_planes + ix] = 1
            else:
                planes.append(self.dummy_obj_plane(EXTENT[5]))
                kind[self.n_box_planes + ix] = 2
        rot_w2s, t_w2s, k_w = np.zeros((bs, 3, 3)), np.zeros((bs, 3, 1)), np.zeros((bs, 3, 3))
        for ix, pl in enumerate(planes):
            centre = utils.get_centre(pl['pt'], pl['x_dir'], pl['y_dir'], pl['w'], pl['h'], off_x=pl['off_x'], off_y=pl['off_y'])
            rot_w2s[ix], t_w2s[ix] = utils.canonical_transform(centre, pl['x_dir'], pl['y_dir'])
            k_w[ix] = utils.dims2kmat(pl['w'], pl['h'], self.h, self.w)           # (sic) data.py:289 passes (h, w) for (w_tex, h_tex); the loader's textures are square
        n_hat_w = np.tile(np.array([[[0.0, 0.0, 1.0]]]), (bs, 1, 1))
        a_w = -np.ones((bs, 1, 1))
        waves = np.zeros((bs, 3, N_WAVES, 4), dtype=np.float32)
        waves[..., 0] = rs.uniform(0.05, 0.22, (bs, 3, N_WAVES))
        waves[..., 1] = rs.uniform(-0.35, 0.35, (bs, 3, N_WAVES)) * (400.0 / self.w)
        waves[..., 2] = rs.uniform(-0.35, 0.35, (bs, 3, N_WAVES)) * (400.0 / self.h)
        waves[..., 3] = rs.uniform(0, 2 * np.pi, (bs, 3, N_WAVES))
        return rot_w2s, t_w2s, k_w, n_hat_w, a_w, waves, kind

    def textures(self, waves, kind):
        """Procedural textures of any number of planes in one launch: imgs [n,h,w,3], masks [n,h,w,1] (CUDA)."""
        n = waves.shape[0]
        dev = self.device
        imgs = torch.empty(n, self.h, self.w, 3, dtype=torch.float32, device=dev)
        masks = torch.empty(n, self.h, self.w, 1, dtype=torch.float32, device=dev)
        wv = torch.tensor(np.ascontiguousarray(waves, dtype=np.float32), device=dev)
        kd = torch.tensor(np.ascontiguousarray(np.minimum(kind, 1), dtype=np.int32), device=dev)
        _b200.call('lsi_b200_procedural_texture', _b200.ptr(wv), _b200.ptr(kd), n, waves.shape[2], self.h, self.w, _b200.ptr(imgs),
                   _b200.ptr(masks), _b200.stream())
        dummy = torch.tensor(np.asarray(kind) == 2, device=dev)
        if bool(dummy.any()):
            masks[dummy] = 0.0                             # data.py:263-264: planes beyond n_obj are fully transparent
            imgs[dummy] = 1.0
        return imgs, masks

    def forward(self):
        """data.py:207-291 -- (rot_w2s, t_w2s, k_w, n_hat_w, a_w, imgs_w, masks_w); the matrices are numpy arrays, the textures CUDA
        tensors [bs,h,w,3] / [bs,h,w,1]."""
        rot_w2s, t_w2s, k_w, n_hat_w, a_w, waves, kind = self.layout()
        imgs_w, masks_w = self.textures(waves, kind)
        return rot_w2s, t_w2s, k_w, n_hat_w, a_w, imgs_w, masks_w


def _t2w_matrices(k_w, k_t, rot_w2s, t_w2s, n_hat_w, a_w, rot_s2t, t_s2t):
    """Per plane: the 3x3 matrix taking target pixels to texture pixels (homography.inv_homography with the composed world->target
    motion, data.py:383-389) and the row vector giving the plane's disparity at a target pixel (inv_homography_dmat), in fp64."""
    rot = rot_s2t[None] @ rot_w2s
    t = t_s2t[None] + rot_s2t[None] @ t_w2s
    rot_t = np.transpose(rot, (0, 2, 1))
    denom = a_w - n_hat_w @ rot_t @ t
    denom = np.where(denom == 0, 1e-8, denom)
    k_t_inv = np.linalg.inv(k_t)
    hom = k_w @ (rot_t + (rot_t @ t @ n_hat_w @ rot_t) / denom) @ k_t_inv[None]
    dmat = (-1.0 * (n_hat_w @ rot_t @ k_t_inv[None])) / denom
    return hom.reshape(-1, 9), dmat.reshape(-1, 3), n_hat_w @ rot_t, a_w - n_hat_w @ (rot_t @ t)


def render_views(imgs_w, masks_w, hom, dmat, h_out, w_out, want_disps=True):
    """One fused launch for V views: imgs_w [V,n,h,w,3], masks_w [V,n,h,w,1] (CUDA), hom [V,n,9], dmat [V,n,3] (numpy or CUDA) ->
    (render [V,H,W,3], disp_fg [V,H,W,1], disp_bg [V,H,W,1])."""
    dev = imgs_w.device
    V, n, h, w, _ = imgs_w.shape
    hom = torch.as_tensor(np.ascontiguousarray(hom, dtype=np.float32)).to(dev) if not torch.is_tensor(hom) else hom
    dmat = torch.as_tensor(np.ascontiguousarray(dmat, dtype=np.float32)).to(dev) if not torch.is_tensor(dmat) else dmat
    out = torch.empty(V, h_out, w_out, 3, dtype=torch.float32, device=dev)
    fg = torch.empty(V, h_out, w_out, 1, dtype=torch.float32, device=dev) if want_disps else None
    bg = torch.empty(V, h_out, w_out, 1, dtype=torch.float32, device=dev) if want_disps else None
    scratch = torch.empty(V, dtype=torch.float32, device=dev)
    _b200.call('lsi_b200_render_planes', _b200.ptr(_b200.dev_f32(imgs_w, 'imgs_w')), _b200.ptr(_b200.dev_f32(masks_w, 'masks_w')),
               _b200.ptr(hom), _b200.ptr(dmat), V, n, h, w, h_out, w_out, MIN_DISP, DEPTH_SOFTMAX_TEMP, _b200.ptr(out), _b200.ptr(fg),
               _b200.ptr(bg), _b200.ptr(scratch), _b200.stream())
    return out, fg, bg


def _downsample(x, factor):
    if factor == 1:
        return x
    B, H, W, C = x.shape
    out = torch.empty(B, H // factor, W // factor, C, dtype=torch.float32, device=x.device)
    _b200.call('lsi_b200_box_downsample', _b200.ptr(x.contiguous()), _b200.ptr(out), B, H, W, C, factor, _b200.stream())
    return out


class Renderer(object):
    """data.py:293-516 -- renders one world from arbitrary viewpoints.  set_feed_dict takes the reference's keys."""

    def __init__(self, n_imgs, h=400, w=400, ds_factor=1):
        self.n_imgs, self.h, self.w, self.ds_factor = n_imgs, h, w, ds_factor
        self._feed = {}

    def set_feed_dict(self, **kwargs):
        """data.py:446-472: k_w, k_s, k_t, rot_w2s, t_w2s, n_hat_w, a_w, imgs_w, masks_w, pixel_coords (ignored: the standard grid is
        implied), rot_s2t, t_s2t."""
        self._feed.update(kwargs)

    def _render(self, rot, t, want_disps):
        f = self._feed
        hom, dmat, n_hat_t, a_t = _t2w_matrices(np.asarray(f['k_w'], np.float64), np.asarray(f['k_t'], np.float64),
                                                np.asarray(f['rot_w2s'], np.float64), np.asarray(f['t_w2s'], np.float64),
                                                np.asarray(f['n_hat_w'], np.float64), np.asarray(f['a_w'], np.float64),
                                                np.asarray(rot, np.float64), np.asarray(t, np.float64).reshape(3, 1))
        img, fg, bg = render_views(f['imgs_w'][None], f['masks_w'][None], hom[None], dmat[None], self.h, self.w, want_disps)
        return img, fg, bg, n_hat_t, a_t

    def render_planes(self, rot, t):
        """data.py:474-486 -- [h/ds, w/ds, 3] rendering of the world from the view (rot, t) relative to the source frame."""
        return _downsample(self._render(rot, t, False)[0], self.ds_factor)[0]

    def render_disps(self, rot, t):
        """data.py:488-502 -- [fg, bg] disparity maps [h/ds, w/ds, 1]."""
        _, fg, bg, _, _ = self._render(rot, t, True)
        return [_downsample(fg, self.ds_factor)[0], _downsample(bg, self.ds_factor)[0]]

    def plane_geometry(self, rot, t):
        """data.py:504-516 -- plane normals [L,1,3] and displacements [L,1,1] in the frame of the view."""
        _, _, _, n_hat_t, a_t = self._render(rot, t, False)
        return [n_hat_t, a_t]


class DataLoader(object):
    """data.py:519-673 -- generator + renderer.  opts needs img_height, img_width, synth_ds_factor, n_obj_max, n_obj_min,
    n_box_planes, data_split, synth_dl_eval_data (sun_imgs_dir / pascal_objects_dir are accepted and ignored)."""

    def __init__(self, opts, _seed=0, _device='cuda'):
        self.opts = opts
        self.output_gt = bool(getattr(opts, 'synth_dl_eval_data', False))
        ds = int(getattr(opts, 'synth_ds_factor', 1))
        self.ds_factor = ds
        w, h = opts.img_width * ds, opts.img_width * ds            # data.py:531-532: (sic) the reference sizes both from img_width
        self.h, self.w = h, w
        self.generator = WorldGenerator(getattr(opts, 'sun_imgs_dir', None), getattr(opts, 'pascal_objects_dir', None), h=h, w=w,
                                        n_obj_max=opts.n_obj_max, n_obj_min=opts.n_obj_min, n_box_planes=opts.n_box_planes,
                                        split=getattr(opts, 'data_split', 'all'), _seed=_seed, _device=_device)
        self.renderer = Renderer(opts.n_box_planes + opts.n_obj_max, h=opts.img_height * ds, w=opts.img_width * ds, ds_factor=ds)
        self.k_s = np.array([[w, 0, w / 2.0], [0, h, h / 2.0], [0, 0, 1.0]])   # data.py:548-557
        self.k_t = np.copy(self.k_s)

    def forward_instance(self):
        """data.py:559-640 -- one pair; returned WITHOUT the batch dimension."""
        return [v[0] for v in self.forward(1)]

    def forward(self, bs):
        """data.py:642-673 -- (img_s, img_t, k_s, k_t, rot, trans[, n_hat, a, disp_s_fg, disp_s_bg, disp_t_fg, disp_t_bg, img_s_bg,
        img_t_bg]) for bs random worlds, rendered in one fused launch per output kind."""
        gen, ds = self.generator, self.ds_factor
        n, nb = gen.bs, gen.n_box_planes
        Hr, Wr = self.renderer.h, self.renderer.w
        layouts = [gen.layout() for _ in range(bs)]
        views = [sample_views(1, _rs=gen.rs)[0] for _ in range(bs)]
        imgs, masks = gen.textures(np.concatenate([l[5] for l in layouts]), np.concatenate([l[6] for l in layouts]))
        imgs = imgs.view(bs, n, gen.h, gen.w, 3)
        masks = masks.view(bs, n, gen.h, gen.w, 1)
        eye, zero = np.eye(3), np.zeros((3, 1))
        hom, dmat, geo = [], [], []
        for (rot_w2s, t_w2s, k_w, n_hat_w, a_w, _, _), (rot_t, t_t) in zip(layouts, views):
            for rot, t in ((eye, zero), (rot_t, t_t)):               # source view = the world frame, then the sampled target view
                h_, d_, nh, a = _t2w_matrices(k_w, self.k_t, rot_w2s, t_w2s, n_hat_w, a_w, rot, t)
                hom.append(h_); dmat.append(d_); geo.append((nh, a))
        hom, dmat = np.stack(hom), np.stack(dmat)                    # [2 bs, n, .]: (src, trg) interleaved per scene
        rep = lambda x: x.unsqueeze(1).expand(bs, 2, *x.shape[1:]).reshape(2 * bs, *x.shape[1:])
        img, fg, bg = render_views(rep(imgs).contiguous(), rep(masks).contiguous(), hom, dmat, Hr, Wr, want_disps=self.output_gt)
        img = _downsample(img, ds).view(bs, 2, Hr // ds, Wr // ds, 3)
        dev = img.device
        f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)
        k = utils.resize_instrinsic(self.k_s, 1.0 / ds, 1.0 / ds)
        out = [img[:, 0].contiguous(), img[:, 1].contiguous(), f32(np.stack([k] * bs)), f32(np.stack([k] * bs)),
               f32(np.stack([v[0] for v in views])), f32(np.stack([v[1] for v in views]))]      # source pose is the identity: rot, trans = target pose
        if self.output_gt:
            fg = _downsample(fg, ds).view(bs, 2, Hr // ds, Wr // ds, 1)
            masks_bg = masks.clone()
            masks_bg[:, nb:] = 0.0                                   # data.py:610-618: background-only world
            # (sic) data.py:620-623 takes the FOREGROUND selection of the background-only world as disp_*_bg
            img_b, fg_b, _ = render_views(rep(imgs).contiguous(), rep(masks_bg).contiguous(), hom, dmat, Hr, Wr, want_disps=True)
            img_b = _downsample(img_b, ds).view(bs, 2, Hr // ds, Wr // ds, 3)
            fg_b = _downsample(fg_b, ds).view(bs, 2, Hr // ds, Wr // ds, 1)
            out += [f32(np.stack([geo[2 * b][0] for b in range(bs)])), f32(np.stack([geo[2 * b][1] for b in range(bs)])),
                    fg[:, 0].contiguous(), fg_b[:, 0].contiguous(), fg[:, 1].contiguous(), fg_b[:, 1].contiguous(),
                    img_b[:, 0].contiguous(), img_b[:, 1].contiguous()]
        return out
