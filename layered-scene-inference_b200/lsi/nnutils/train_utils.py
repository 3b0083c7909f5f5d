"""Mirror of the training wiring: lsi/nnutils/train_utils.py (Trainer: Adam, seeds, step loop) and the model/loss
wiring of ldi_enc_dec.py (Trainer.define_pred_graph :175-228, define_loss_graph :265-410), re-hosted on torch +
the B200 kernels.  The TF session / Supervisor / summary plumbing of the reference is out of scope; what is kept is the
arithmetic, the hyper-parameters and the checkpoint save / resume / pretrain-restore protocol (lsi.nnutils.checkpoint):

  * two U-Net towers with shared variables (src and trg image), heads, disp *= max_disp      ldi_enc_dec.py:196-221
  * view-synthesis loss                                                                      ldi_enc_dec.py:265-410
  * Adam(learning_rate=1e-4, beta1=0.9, beta2=0.999, eps=1e-8) on every variable with a gradient   train_utils.py:56-57,107-117
  * seed 0                                                                                   train_utils.py:158-160

Data parallelism (not in the reference): one process per GPU, each rank runs the step on its shard of the
batch, then ONE all-reduce (sum) of the flat gradient buffer over NCCL and a fused Adam step on the flat parameter buffer
with the gradients scaled by 1/world_size.  Batch-norm statistics are per replica by default; Trainer(sync_bn=True)
all-reduces the per-channel sums of every BN layer (forward and backward) so that the N-rank step equals the reference's
single-device step on the global batch (tests/test_gpu_dp.py).
"""
import os
import types

import numpy as np
import torch

from lsi import _b200
from lsi.loss import loss as loss_mod
from lsi.nnutils import checkpoint as ckpt
from lsi.nnutils import helpers as nn_helpers
from lsi.nnutils import nets


def default_opts(**kw):
    """Flag defaults of ldi_enc_dec.py:39-123 / train_utils.py:31-57 with the synthetic-dataset constants of main()
    (ldi_enc_dec.py:415-420); pass dataset='kitti' for the KITTI ones (:421-425)."""
    o = types.SimpleNamespace(
        n_layers=2, batch_size=2, img_height=256, img_width=256, learning_rate=1e-4, beta1=0.9,
        self_cons_wt=1.0, indep_splat_wt=1.0, compose_splat_wt=1.0, splat_bdry_ignore=0.1, zbuf_scale=50.0,
        trg_splat_downsampling=0.5, disp_smoothness_wt=0.1, incr_depth_wt=10.0, l0_self_cons=False,
        use_unet=True, n_layerwise_steps=3, pred_ldi_masks=False, bg_layer_disp=0.2, max_disp=1.0, dataset='synthetic',
        # logging / snapshotting and loop control (train_utils.py:38-57)
        checkpoint_dir='/code/lsi/cachedir/snapshots/', pretrain_name='', pretrain_iter=100000, num_iter=100000, log_freq=5,
        checkpoint_freq=50000, save_latest_freq=2000,
        # data (ldi_enc_dec.py:48-82)
        data_split='train', synth_ds_factor=1, n_obj_min=1, n_obj_max=4, n_box_planes=5, synth_dl_eval_data=False,
        sun_imgs_dir=None, pascal_objects_dir=None, kitti_data_root='/datasets/kitti', kitti_dataset_variant='mview',
        kitti_dl_disparities=False, debug_synth_texture=False)
    if kw.get('dataset') == 'kitti':
        o.bg_layer_disp, o.max_disp = 1e-3, 0.4
    o.__dict__.update(kw)
    return o


def shard_batch(batch, rank, world_size):
    """Contiguous, disjoint, complete split of the leading (batch) dimension across ranks."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        if n % world_size:
            raise ValueError('batch size %d is not divisible by world size %d' % (n, world_size))
        per = n // world_size
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def allreduce_sum_(flat, group=None):
    """The one collective of the training step: sum the flat gradient buffer across ranks (NCCL on GPUs; gloo in the
    CPU tests).  No-op when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return dist.get_world_size(group)
    return 1


def predict_ldi(img, opts, store, reuse):
    """ldi_enc_dec.py:196-213 for one tower: U-Net trunk -> heads -> disp *= max_disp.  Images whose size the reference
    U-Net cannot take (not a multiple of 128) are zero-padded and the prediction cropped (nets.pad_to_legal)."""
    padded, (h, w) = nets.pad_to_legal(img)
    if getattr(opts, 'use_unet', True):                               # ldi_enc_dec.py:197-205
        _, feat_dec, skip_feat, _ = nets.encoder_decoder_unet(padded, nl_diff_enc_dec=opts.n_layerwise_steps, reuse=reuse,
                                                              _store=store)
    else:
        _, feat_dec, skip_feat, _ = nets.encoder_decoder_simple(padded, nl_diff_enc_dec=opts.n_layerwise_steps, reuse=reuse,
                                                                _store=store)
    # the crop back to (h, w) is fused into the prediction conv (it only evaluates the top-left window)
    if not torch.is_grad_enabled() and not opts.pred_ldi_masks:
        # inference: the disparity scale is a per-channel output factor of the prediction conv, so textures and
        # disparities stay views of ONE packed [L,B,H,W,4] tensor and the renderer reads 16 bytes per pixel-layer
        return nets.ldi_predictor(feat_dec, n_layers=opts.n_layers, reuse=reuse, n_layerwise_steps=opts.n_layerwise_steps,
                                  skip_feat=skip_feat, pred_masks=False, _store=store, _out_hw=(h, w),
                                  _disp_scale=(None if opts.max_disp == 1 else opts.max_disp))
    tex, masks, disps = nets.ldi_predictor(feat_dec, n_layers=opts.n_layers, reuse=reuse,
                                           n_layerwise_steps=opts.n_layerwise_steps, skip_feat=skip_feat,
                                           pred_masks=opts.pred_ldi_masks, _store=store, _out_hw=(h, w))
    if opts.max_disp == 1:
        return [tex, masks, disps]
    if not torch.is_grad_enabled():
        disps.mul_(opts.max_disp)
        return [tex, masks, disps]
    return [tex, masks, disps * opts.max_disp]


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process (its pipeline thread and the pinned host buffers it allocates afterwards: first touch) to the CPUs of
    the NUMA node the GPU hangs off, read from sysfs.  With one process per GPU this keeps every rank's host<->device copies on
    its own socket's memory controllers and PCIe root instead of all ranks sharing node 0.  Returns the node (or None when
    sysfs does not expose it, e.g. in a container: nothing is changed then)."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError):
        return None


class HostViewPipeline(object):
    """Image -> LDI -> rendered target view (the predict + render path of ldi_pred_eval.py:117-224) over a sequence of
    HOST batches.  Three CUDA streams: while batch k runs its kernels, batch k+1's images and cameras are copied
    host->device and batch k-1's rendered views device->host (PCIe is full duplex), through `depth` device input
    buffers and `depth` pinned host output buffers.  Every byte of every batch still crosses the bus; only the
    waiting is removed.  The arithmetic of a batch is exactly predict_ldi + forward_splat (batch-norm statistics stay
    per batch).

    batches: sequence of dicts of pinned host tensors {'img' [B,H,W,3], 'k_s','k_t','rot' [B,3,3], 't' [B,3] or [B,3,1]}.
    'img' is float32 in [0,1], or -- u8_input=True -- the 8-bit image data as it comes out of a decoder (uint8): a quarter of the
    host->device bytes; the scaling to [0,1] then runs on the device (lsi_b200_area_resize_u8 at unit scale, the kernel the KITTI
    loader uses).  on_result(k, img_host, wts_host) is called once batch k's views are in host memory (the buffers are
    recycled after the callback returns)."""

    def __init__(self, opts, store, render_kw, batch, height, width, device, depth=2, u8_input=False):
        from lsi.geometry import ldi as ldi_utils
        self._render = ldi_utils.forward_splat
        if depth < 2:      # upload(k+1) is issued before batch k's kernels: with one slot it would overwrite batch k's inputs
            raise ValueError('HostViewPipeline needs depth >= 2, got %d' % depth)
        self.opts, self.store, self.kw, self.device, self.depth = opts, store, dict(render_kw), torch.device(device), depth
        ds = float(render_kw.get('trg_downsampling', 1))
        ht, wt = int(height * ds), int(width * ds)
        dev = self.device
        self.pc = nn_helpers.pixel_coords(batch, height, width, _device=dev)
        self.u8 = bool(u8_input)
        if self.u8 and batch * height > 65535:
            raise ValueError('u8_input: batch * height must be <= 65535 (the batch is converted as one tall image)')
        self.x = [torch.empty(batch, height, width, 3, device=dev, dtype=torch.uint8 if self.u8 else torch.float32) for _ in range(depth)]
        self.xf = torch.empty(batch, height, width, 3, device=dev) if self.u8 else None      # converted images of the batch being computed
        self.cams = [dict(k_s=torch.empty(batch, 3, 3, device=dev), k_t=torch.empty(batch, 3, 3, device=dev),
                          rot=torch.empty(batch, 3, 3, device=dev), t=None) for _ in range(depth)]
        self.out_img = [torch.empty(1, batch, ht, wt, 3).pin_memory() for _ in range(depth)]
        self.out_wts = [torch.empty(1, batch, ht, wt, 1).pin_memory() for _ in range(depth)]
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d_bytes(self, batch):
        return sum(batch[k].numel() * batch[k].element_size() for k in ('img', 'k_s', 'k_t', 'rot', 't'))

    def d2h_bytes(self):
        return (self.out_img[0].numel() + self.out_wts[0].numel()) * 4

    def run(self, batches, on_result=None):
        n, depth = len(batches), self.depth
        cur = torch.cuda.current_stream(self.device)
        ready, free_in, out_done = [None] * depth, [None] * depth, [None] * depth

        def upload(k):
            s, b = k % depth, batches[k]
            with torch.cuda.stream(self.s_in):
                if free_in[s] is not None:
                    self.s_in.wait_event(free_in[s])            # batch k - depth has finished reading this buffer
                self.x[s].copy_(b['img'], non_blocking=True)
                c = self.cams[s]
                for key in ('k_s', 'k_t', 'rot'):
                    c[key].copy_(b[key], non_blocking=True)
                c['t'] = b['t'].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                ready[s] = ev

        def deliver(k):
            s = k % depth
            out_done[s].synchronize()
            if on_result is not None:
                on_result(k, self.out_img[s], self.out_wts[s])
            out_done[s] = None

        upload(0)
        for k in range(n):
            if k + 1 < n:
                upload(k + 1)
            s = k % depth
            cur.wait_event(ready[s])
            with torch.no_grad():
                x = self.x[s]
                if self.u8:          # 8-bit -> float32 / 255 on the device (the whole batch as one [B*H, W, 3] image, unit scale)
                    bh, w = x.shape[0] * x.shape[1], x.shape[2]
                    _b200.call('lsi_b200_area_resize_u8', _b200.ptr(x), bh, w, 3, _b200.ptr(self.xf), bh, w, 3, _b200.stream())
                    x = self.xf
                ldi = predict_ldi(x, self.opts, self.store, reuse=True)
                c = self.cams[s]
                c['t'].record_stream(cur)
                # the cameras are re-uploaded every step: identify the set by the content of the host copies, so that the renderer's
                # pose-class memo (lsi/geometry/ldi.py) can recognise a rectified batch it has classified before
                hb = batches[k]
                pose_key = hash(tuple(hb[q].numpy().tobytes() for q in ('k_s', 'k_t', 'rot', 't')))
                img, wts = self._render(tuple(ldi), self.pc, c['k_s'], c['k_t'], c['rot'], c['t'], _pose_key=pose_key, **self.kw)[:2]
            ev = torch.cuda.Event()
            ev.record(cur)
            free_in[s] = ev
            if out_done[s] is not None:
                deliver(k - depth)                              # the host buffer is about to be overwritten
            self.s_out.wait_event(ev)
            with torch.cuda.stream(self.s_out):
                self.out_img[s].copy_(img, non_blocking=True)
                self.out_wts[s].copy_(wts, non_blocking=True)
                img.record_stream(self.s_out)
                wts.record_stream(self.s_out)
                evo = torch.cuda.Event()
                evo.record(self.s_out)
                out_done[s] = evo
        for k in range(max(0, n - depth), n):
            if out_done[k % depth] is not None:
                deliver(k)
        cur.synchronize()


class Trainer(object):
    """One training step of ldi_enc_dec.py on the B200 path."""

    def __init__(self, opts, store=None, group=None, sync_bn=False):
        """sync_bn: batch-norm statistics over the GLOBAL batch (all ranks of `group`), i.e. the reference's single-device
        semantics (nets.py:263-272) under data parallelism; default False = per-replica statistics (documented choice,
        DESIGN.md section 7)."""
        self.opts = opts
        self.sync_bn = bool(sync_bn)
        self.store = store if store is not None else nets.ParamStore(seed=0)
        self.group = group
        self.step_count = 0          # global_step (train_utils.py:107-116)
        self.adam_t = 0              # Adam's own step count (bias correction); restarts whenever the slots are not restored
        self.m = self.v = None
        self._built = False
        self._deferred = None        # optimistic restore waiting for the variables to exist
        self._pending_slots = None   # Adam slots read from a checkpoint before m / v were allocated
        self._probe = None

    def define_pred_graph(self, imgs_src, imgs_trg):
        """ldi_enc_dec.py:175-228: both towers share their variables (reuse=True for the second)."""
        ldi_src = predict_ldi(imgs_src, self.opts, self.store, reuse=self._built)
        self._built = True
        ldi_trg = predict_ldi(imgs_trg, self.opts, self.store, reuse=True)
        return ldi_src, ldi_trg

    def define_loss_graph(self, ldi_src, ldi_trg, batch):
        """ldi_enc_dec.py:265-410."""
        b, h, w, _ = batch['imgs_src'].shape
        pc = nn_helpers.pixel_coords(b, h, w, _device=batch['imgs_src'].device)
        return loss_mod.view_synthesis_loss(ldi_src, ldi_trg, batch['imgs_src'], batch['imgs_trg'], pc, batch['k_s'],
                                            batch['k_t'], batch['rot_mat'], batch['trans_mat'], self.opts)

    # ---- data + loop (train_utils.py:63-66, 149-222; ldi_enc_dec.py:130-137, 230-263) ------------------------------------
    def define_data_loader(self):
        """ldi_enc_dec.py:130-137: the synthetic planar-room generator or the KITTI stereo-pair loader."""
        opts = self.opts
        if opts.dataset == 'synthetic':
            from lsi.data.syntheticPlanes import data as synthetic_planes
            self.data_loader = synthetic_planes.DataLoader(opts)
        elif opts.dataset == 'kitti':
            from lsi.data.kitti import data as kitti_data
            self.data_loader = kitti_data.DataLoader(opts)
            self.data_loader.define_queues()
            self.data_loader.preload_calib_files()
        else:
            raise ValueError('unknown dataset %r' % (opts.dataset,))

    def feed(self):
        """ldi_enc_dec.py:230-263: one batch from the loader, keyed like the reference's placeholders."""
        if getattr(self.opts, 'debug_synth_texture', False):
            raise NotImplementedError('debug_synth_texture (ground-truth disparities in place of the prediction) is not provided')
        data = self.data_loader.forward(self.opts.batch_size)
        img_src, img_trg, k_s, k_t, rot_mat, trans_mat = data[:6]
        dev = self.store.device
        f = lambda x: torch.as_tensor(x, dtype=torch.float32).to(dev)
        return dict(imgs_src=f(img_src), imgs_trg=f(img_trg), k_s=f(k_s), k_t=f(k_t), rot_mat=f(rot_mat), trans_mat=f(trans_mat))

    def train(self, on_log=None):
        """train_utils.py:149-222 -- the training routine: seed 0, data loader, resume from the latest checkpoint of
        opts.checkpoint_dir or else start from the pretrained net, then opts.num_iter steps with the reference's cadence: every
        log_freq steps the losses are reported (on_log(global_step, total, parts); they are also appended to self.log -- the role
        of the TF summaries), every save_latest_freq steps `model.latest` is written, every checkpoint_freq steps
        `model-<global_step>`.  Returns self.log."""
        import numpy as np_
        opts = self.opts
        torch.manual_seed(0)
        np_.random.seed(0)
        self.define_data_loader()
        what, path = self.init_from_checkpoints(opts.checkpoint_dir, getattr(opts, 'pretrain_name', '') or None,
                                                getattr(opts, 'pretrain_iter', 0))
        self.log = [('init', what, path)]
        for step in range(1, opts.num_iter + 1):
            total, parts = self.train_step(self.feed())
            gs = self.step_count
            if step % opts.log_freq == 0:
                rec = (gs, float(total), {k: float(v) for k, v in parts.items() if torch.is_tensor(v) or isinstance(v, float)})
                self.log.append(rec)
                if on_log is not None:
                    on_log(*rec)
            if step % opts.save_latest_freq == 0:
                self.save(opts.checkpoint_dir, 'latest')
            if step % opts.checkpoint_freq == 0:
                self.save(opts.checkpoint_dir, gs)
        return self.log

    # ---- checkpoints (train_utils.py:172-200, 224-232) -----------------------------------------------------------
    def _adam_slots(self):
        if self.m is None or self.store.flat is None:
            return None, None
        m, v, off = {}, {}, 0
        for k in sorted(self.store.vars):
            n = self.store.vars[k].numel()
            m[k] = self.m[off:off + n].view(self.store.vars[k].shape)
            v[k] = self.v[off:off + n].view(self.store.vars[k].shape)
            off += n
        return m, v

    def save(self, checkpoint_dir, step):
        """train_utils.py:224-232: step == 'latest' -> model.latest, else model-<global_step>.  Variables under their TF
        names, global_step, and (beyond the reference, for exact resume) the Adam slots with Adam's step count."""
        m, v = self._adam_slots()
        return ckpt.save_checkpoint(ckpt.checkpoint_path(checkpoint_dir, step), self.store.vars, self.step_count, m, v,
                                    adam_t=self.adam_t)

    def _take_slots(self, saved):
        """Adam slots of a checkpoint: applied now if m / v exist, else stashed until train_step allocates them.  Without
        slots in the file Adam restarts (adam_t = 0), as the reference's does after any restore (its Saver keeps
        neither the slots nor the beta powers)."""
        names = [k for k in self.store.vars if k + '/Adam' in saved and k + '/Adam_1' in saved]
        if not names or len(names) != len(self.store.vars):
            self._pending_slots, self.adam_t = None, 0
            if self.m is not None:
                self.m.zero_()
                self.v.zero_()
            return
        self.adam_t = int(saved['adam_t']) if 'adam_t' in saved else int(saved.get('global_step', 0))
        self._pending_slots = {k: (saved[k + '/Adam'], saved[k + '/Adam_1']) for k in names}
        self._apply_pending_slots()

    def _apply_pending_slots(self):
        m, v = self._adam_slots()
        if m is None or self._pending_slots is None:
            return
        with torch.no_grad():
            for k, (sm, sv) in self._pending_slots.items():
                m[k].copy_(torch.from_numpy(sm).to(m[k].device))
                v[k].copy_(torch.from_numpy(sv).to(v[k].device))
        self._pending_slots = None

    def restore(self, path):
        """saver.restore (train_utils.py:195): every model variable must be present in the file with its shape.  On a Trainer
        whose variables do not exist yet (they are created by the first forward) the variables are created FROM the
        checkpoint, so that a resume can never silently keep a random initialisation.  Resumes global_step and, when the file
        has them, the Adam slots."""
        saved = ckpt.read_checkpoint(path)
        model_keys = [k for k in saved if k not in ckpt.NON_VARIABLE_KEYS and not k.endswith('/Adam') and not k.endswith('/Adam_1')]
        missing_in_store = [k for k in model_keys if k not in self.store.vars]
        if missing_in_store:
            if self.store.flat is not None:
                raise RuntimeError('checkpoint %s has variables the (already flattened) model lacks: %s' % (path, missing_in_store[:5]))
            self.store.load_state_dict({k: torch.from_numpy(np.asarray(saved[k])) for k in missing_in_store})
        r = ckpt.optimistic_restorer(path, self.store)
        if r.new_vars or r.shape_mismatch:
            raise RuntimeError('checkpoint %s does not match the model: missing %s, shape mismatch %s'
                               % (path, r.new_vars[:5], r.shape_mismatch[:5]))
        self.step_count = r.restore(self.store)
        self._deferred = None
        self._take_slots(saved)
        return self.step_count

    def init_from_checkpoints(self, checkpoint_dir, pretrain_name=None, pretrain_iter=0):
        """train_utils.py:176-200: resume from the latest checkpoint of checkpoint_dir if there is one; otherwise, with
        pretrain_name, optimistically restore <checkpoint_dir>/../<pretrain_name>/model-<pretrain_iter> (variables that
        exist there with the same shape; the rest keep their initialisation).  Returns what happened.  The optimistic
        restore of a Trainer whose variables do not exist yet is applied as soon as the first forward has created them."""
        latest = ckpt.latest_checkpoint(checkpoint_dir)
        if latest is not None:
            self.restore(latest)
            return 'resumed', latest
        if pretrain_name:
            path = ckpt.checkpoint_path(os.path.normpath(os.path.join(checkpoint_dir, '..', pretrain_name)), pretrain_iter)
            saved = ckpt.read_checkpoint(path)
            self.step_count = int(saved[ckpt.global_step_key(saved)]) if ckpt.global_step_key(saved) else 0
            self.adam_t, self._pending_slots = 0, None        # a pretrained net starts a new optimisation (reference: slots are never saved)
            if not self.store.vars:
                self._deferred = path
            else:
                ckpt.optimistic_restorer(path, self.store).restore(self.store)
            return 'pretrained', path
        return 'fresh', None

    def _build_variables(self, batch):
        """Create every variable once (one no-grad forward), apply a deferred optimistic restore, flatten, allocate Adam slots."""
        with torch.no_grad():
            self.define_pred_graph(batch['imgs_src'][:1], batch['imgs_trg'][:1])
        if self._deferred is not None:
            ckpt.optimistic_restorer(self._deferred, self.store).restore(self.store)
            self._deferred = None
        flat, _ = self.store.flatten()
        self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
        self._apply_pending_slots()

    def train_step(self, batch, dp_check=False):
        """forward -> loss -> backward -> all-reduce(sum) of the flat gradients -> Adam.  Returns (total_loss, parts); with
        dp_check=True also a dict proving the collective: <sum over ranks of the shard gradients, r> against <all-reduced
        gradient, r> for a fixed random vector r (identical on every rank)."""
        nets.set_sync_bn(self.sync_bn, self.group)
        if self.store.flat is None:
            self._build_variables(batch)
        self.store.zero_grad()
        with nets.grad_sink():      # weight gradients go straight into the flat gradient buffer (no temporaries, no accumulation launches)
            ldi_src, ldi_trg = self.define_pred_graph(batch['imgs_src'], batch['imgs_trg'])
            total, parts = self.define_loss_graph(ldi_src, ldi_trg, batch)
            total.backward()
        chk = None
        g = self.store.flat_grad
        if dp_check:
            if self._probe is None or self._probe.numel() != g.numel():
                gen = torch.Generator(device=g.device)
                gen.manual_seed(1234)
                self._probe = torch.randn(g.numel(), device=g.device, generator=gen)
            shard = torch.dot(g.double(), self._probe.double()).reshape(1)
            allreduce_sum_(shard, self.group)
        world = allreduce_sum_(g, self.group)
        if dp_check:
            chk = {'proj_sum_of_shards': float(shard), 'proj_allreduced': float(torch.dot(g.double(), self._probe.double())),
                   'world': world}
        self.step_count += 1
        self.adam_t += 1
        o = self.opts
        _b200.call('lsi_b200_adam_step', _b200.ptr(self.store.flat), _b200.ptr(self.store.flat_grad), _b200.ptr(self.m),
                   _b200.ptr(self.v), self.store.flat.numel(), o.learning_rate, o.beta1, 0.999, 1e-8, self.adam_t,
                   1.0 / world, _b200.stream())
        self.store.touch()        # parameter values changed through raw pointers: invalidate the inference path's filter memo
        out = (total.detach(), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in parts.items()})
        return out + (chk,) if dp_check else out
