"""The LGD / IEF model with the reference's class surface, executed by libempose_b200.

Mirrors ``empose/nn/models.py``: ``create_model`` (``:23-33``), ``BaseModel`` bookkeeping (``:41-96``)
and ``IterativeErrorFeedback`` (``:369-688``).  Same constructor arguments, same ``forward`` /
``backward`` signatures and outputs, same ``*_history`` attributes, same state-dict keys -- but
``forward`` marshals the batch through the C ABI (``empose_b200.lib``) into hand-written sm_100a CUDA
instead of running PyTorch ops.

Training (``net.train()``): the parameters are re-homed as views of ONE flat CUDA vector (their ``.grad`` as views
of a second one), ``forward`` runs the train-mode pass (BatchNorm on batch statistics) and ``backward`` adds to
``.grad`` exactly what the reference leaves there after ``net(batch)`` + ``net.backward(batch, out)`` -- including
the gradients its forward pass accumulates as a side effect (``models.py:576``).  Any ``torch.optim`` optimiser
built on ``net.parameters()`` then steps in place, and data-parallel training all-reduces the single flat
gradient vector (``allreduce_gradients``).
"""
import re

import numpy as np
import torch
import torch.nn as nn

from empose_b200 import lib as _lib
from empose_b200.helpers.configuration import CONSTANTS as C
from empose_b200.nn.layers import MLP
from empose_b200.nn.layers import RNNLayer


def create_model(config, *args):
    """``models.py:23-33``.  Only the LGD family is implemented by this package."""
    m_type = config.m_type
    if m_type in ('ief', 'lgd'):
        return IterativeErrorFeedback(config, *args)
    if m_type == 'rnn':
        return SimpleRNN(config, *args)
    if m_type == 'resnet':
        raise NotImplementedError("model type 'resnet' is not built by empose_b200")
    raise ValueError("Model type '{}' unknown.".format(m_type))


def _frame_mask(seq_lengths, n_frames):
    t = torch.arange(n_frames, device=seq_lengths.device).unsqueeze(0)
    return t < seq_lengths.reshape(-1, 1)


class IterativeErrorFeedback(nn.Module):
    """The LGD(-RNN) model (``models.py:369-688``)."""

    def __init__(self, config, smpl_model, precision=_lib.PRECISION_FP16):
        super(IterativeErrorFeedback, self).__init__()
        self.config = config
        self.n_markers = config.n_markers if getattr(config, 'n_markers', -1) > -1 else C.N_TRACKERS_WO_ROOT
        assert self.n_markers in [6, 12]                                    # models.py:385
        self.n_frames = config.window_size
        self.smpl = smpl_model
        self.N = config.m_num_iterations
        self.step_size = config.m_step_size
        self.shape_avg = config.m_average_shape
        self.r_weight = config.m_reprojection_loss_weight
        self.use_gradient = config.m_use_gradient
        self.skip_connections = config.m_skip_connections
        self.rnn_init = config.m_rnn_init
        self.estimate_shape = config.m_estimate_shape
        self.fk_loss_weight = config.m_fk_loss
        self.do_fk = self.fk_loss_weight > 0.0
        if self.do_fk:
            assert self.smpl is not None
        self.shape_weight = getattr(config, 'm_shape_loss_weight', 1.0)
        self.pose_weight = getattr(config, 'm_pose_loss_weight', 1.0)
        self.precision = precision
        if getattr(config, 'use_marker_nor', False):
            raise ValueError('Normals currently not supported.')             # models.py:122-123
        if getattr(config, 'm_rnn_bidirectional', False) and self.rnn_init:
            raise NotImplementedError('bidirectional init RNN is not part of the released LGD models')

        # sizes (models.py:397-421); the constructor mutates the config like the reference does
        self.pos_d_start, self.pos_d_end, self.ori_d_start, self.ori_d_end = 0, 0, 0, 0
        input_size = 0
        if config.use_marker_pos:
            input_size += self.n_markers * 3
            self.pos_d_end = self.n_markers * 3
            self.ori_d_start = self.pos_d_end
        if config.use_marker_ori:
            input_size += self.n_markers * 9
            self.ori_d_end = self.ori_d_start + self.n_markers * 9
        self.input_size = input_size
        self.pose_size = (C.N_JOINTS + 1) * 3
        self.shape_size = C.N_SHAPE_PARAMS
        self.input_iter_size = input_size + self.pose_size + self.shape_size
        if self.use_gradient:
            self.input_iter_size += self.pose_size + self.shape_size
        for k in ('input_size', 'pose_size', 'shape_size', 'input_iter_size'):
            setattr(config, k, getattr(self, k))

        # parameter containers with the reference's names (models.py:424-454)
        bn = not config.m_no_batch_norm
        mlp = lambda n_in, n_out: MLP(n_in, n_out, config.m_hidden_size, config.m_num_layers, config.m_dropout_hidden,
                                      self.skip_connections, bn)
        if self.rnn_init:
            self.rnn = RNNLayer(self.input_size, config.m_rnn_hidden_size, config.m_rnn_num_layers,
                                dropout=config.m_dropout, bidirectional=False)
            self.pose_net_init = nn.Linear(config.m_rnn_hidden_size, self.pose_size)
            self.shape_net_init = nn.Linear(config.m_rnn_hidden_size, self.shape_size)
        else:
            self.pose_net_init = mlp(self.input_size, self.pose_size)
            self.shape_net_init = mlp(self.input_size, self.shape_size)
        self.pose_net_iter = mlp(self.input_iter_size, self.pose_size)
        self.shape_net_iter = mlp(self.input_iter_size, self.shape_size)
        self.smpl_loss = nn.L1Loss(reduction='none')

        self.vertex_ids = C.VERTEX_IDS
        self.marker_idxs = list(range(12)) if self.n_markers == 12 else C.S_CONFIG_6
        self.markers_hat_history = None
        self.markers_ori_hat_history = None
        self.pose_hat_history = None
        self.shape_hat_history = None
        self.joints_hat_history = None
        self._ctx = None
        self._ctx_key = None
        self._trainer = None
        self._trainer_device = None
        self._flat = None                    # dict(params=, grads=, bn=, entries=) once training has started
        self._train_batch_shape = None
        self._hist_buf = None                # history tensors of the last single-span forward, reused when the shape repeats
        #: fp16 operands (the inference default) cannot represent everything fp32 can.  With the guard on, (i) a model
        #: whose BatchNorm-folded weights leave the range where fp16 keeps its 11 significant bits is run on tf32 tensor
        #: cores from the start, and (ii) a forward pass whose result is not finite (an activation beyond 65504) is
        #: repeated on tf32.  Costs one scalar device->host read per forward; set False for asynchronous pipelines.
        self.precision_guard = True
        self.precision_fallbacks = 0         # forward passes the guard repeated on tf32

    # ------------------------------------------------------------------------------------------------
    def model_name(self):
        """``models.py:459-469``."""
        c = self.config
        name = "IEF-{}x{}-N{}".format(c.m_num_layers, c.m_hidden_size, c.m_num_iterations)
        if self.rnn_init:
            name += '-{}RNN-{}x{}'.format('', c.m_rnn_num_layers, c.m_rnn_hidden_size)
        name += '-r{}-ws{}-lr{}'.format(self.r_weight, c.window_size, c.lr)
        name += '-grad' if self.use_gradient else ''
        name += '-skip' if self.skip_connections else ''
        name += '-n{}'.format(self.n_markers)
        return name

    # ------------------------------------------------------------------------------------------------
    def _native_config(self, device_index):
        c = self.config
        return dict(n_markers=self.n_markers, num_iterations=self.N, step_size=float(self.step_size),
                    rnn_init=int(self.rnn_init), average_shape=int(self.shape_avg), use_gradient=int(self.use_gradient),
                    use_marker_pos=int(c.use_marker_pos), use_marker_ori=int(c.use_marker_ori),
                    hidden_size=int(c.m_hidden_size), num_layers=int(c.m_num_layers),
                    rnn_hidden_size=int(c.m_rnn_hidden_size), rnn_num_layers=int(c.m_rnn_num_layers),
                    skip_connections=int(self.skip_connections), batch_norm=int(not c.m_no_batch_norm),
                    precision=int(self.precision), device=int(device_index))

    def _weights_key(self, device_index, precision):
        tensors = [t for k, t in self.state_dict(keep_vars=True).items()]
        return (device_index, precision, self._weights_epoch) + tuple((t.data_ptr(), t._version) for t in tensors)

    _weights_epoch = 0

    def invalidate(self):
        """
        Drop the packed copies of the weights (inference context and trainer) so that the next ``forward`` re-reads the
        module's tensors.  The cache notices ordinary in-place updates (``tensor._version``) and re-assignments
        (``data_ptr``), but NOT writes through ``.data`` (``p.data.mul_()``, EMA / weight-surgery code): call this after such
        writes.  ``train()`` / ``eval()`` transitions call it themselves.
        """
        self._weights_epoch += 1
        self._hist_buf = None
        if self._trainer is not None:
            self._trainer.close()
            self._trainer = None
            self._trainer_device = None

    def train(self, mode=True):
        if mode != self.training:
            self._weights_epoch += 1              # the other mode's packed weights are stale once this one has stepped
        return super(IterativeErrorFeedback, self).train(mode)

    def folded_weight_range(self):
        """(largest, smallest non-zero) magnitude of the BatchNorm-folded Linear / LSTM weights and biases -- what the
        fp16 operand mode has to represent (the context folds BN exactly like this, csrc/model_internal.h pack_linear)."""
        sd = self.state_dict()
        hi, lo = 0.0, float('inf')
        for name, w in sd.items():
            if name.startswith('smpl.') or not w.is_floating_point() or w.dim() != 2:
                continue
            w = w.detach().double().cpu()
            bn = None
            if name.endswith('input_to_hidden.weight'):
                bn = name[:-len('input_to_hidden.weight')] + 'batch_norm'
            else:
                m = re.match(r'(.*\.layers\.)(\d+)\.weight$', name)
                if m:
                    bn = m.group(1) + str(int(m.group(2)) + 1)
            if bn is not None and (bn + '.running_var') in sd:
                scale = sd[bn + '.weight'].double().cpu() / torch.sqrt(sd[bn + '.running_var'].double().cpu() + 1e-5)
                w = w * scale.reshape(-1, 1)
            mag = w.abs()
            hi = max(hi, float(mag.max()))
            nz = mag[mag > 0]
            if nz.numel():
                lo = min(lo, float(nz.min()))
        return hi, lo

    def _effective_precision(self):
        """fp16 unless the guard finds folded weights that fp16 would overflow or flush (see ``precision_guard``)."""
        if self.precision != _lib.PRECISION_FP16 or not self.precision_guard:
            return self.precision
        key = (self._weights_epoch,) + tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())
        if getattr(self, '_range_key', None) != key:
            hi, _ = self.folded_weight_range()
            # 65504 is fp16's largest value; activations are sums of ~hundreds of weight x O(1)-input products, so keep
            # two orders of magnitude of head room.  (Small weights are harmless: below 6e-5 fp16 degrades gracefully.)
            self._range_ok = hi < 256.0
            self._range_key = key
        return _lib.PRECISION_FP16 if self._range_ok else _lib.PRECISION_TF32

    def native_context(self, device, precision=None):
        """Build (or reuse) the native context for the current weights on ``device``."""
        if device.type != 'cuda':
            raise _lib.EmposeError('empose_b200 runs on CUDA devices only (no CPU fallback); got %s' % device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        precision = self._effective_precision() if precision is None else precision
        key = self._weights_key(index, precision)
        cache = self.__dict__.setdefault('_ctx_cache', {})
        for prec in list(cache):                         # contexts of other precisions built from older weights
            if cache[prec][0][2:] != key[2:] or cache[prec][0][0] != index:
                cache.pop(prec)[1].close()
        if precision not in cache:
            arrays = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()
                      if not k.startswith('smpl.') and v.is_floating_point()}
            arrays = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in arrays.items()}
            arrays.update(self.smpl.submodel_arrays())
            cfg = self._native_config(index)
            cfg['precision'] = int(precision)
            cache[precision] = (key, _lib.IefContext(cfg, arrays))
        self._ctx, self._ctx_key = cache[precision][1], key
        return self._ctx

    # ------------------------------------------------------------------------------------------------
    # training
    # ------------------------------------------------------------------------------------------------
    def _trainable_items(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return sd

    def trainer(self, device):
        """Build (or reuse) the native training context on ``device``; re-homes parameters into flat vectors."""
        if device.type != 'cuda':
            raise _lib.EmposeError('empose_b200 trains on CUDA devices only (no CPU fallback); got %s' % device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._trainer is not None and self._trainer_device == index:
            return self._trainer
        if getattr(self.config, 'm_dropout_hidden', 0.0) > 0.0 or getattr(self.config, 'm_dropout', 0.0) > 0.0:
            raise NotImplementedError('training with dropout > 0 is not implemented (the released models use 0)')
        if self._trainer is not None:
            self._trainer.close()
            self._trainer = None
        cfg = self._native_config(index)
        if cfg['precision'] == _lib.PRECISION_FP16:          # fp16 operands are an inference mode: train on tf32 tensor cores
            cfg['precision'] = _lib.PRECISION_TF32
        entries, n_params, n_buffers = _lib.train_layout(cfg)
        dev = torch.device('cuda', index)
        params = torch.zeros(max(n_params, 4), dtype=torch.float32, device=dev)
        grads = torch.zeros_like(params)
        bn = torch.zeros(max(n_buffers, 4), dtype=torch.float32, device=dev)
        items = self._trainable_items()
        for name, kind, off, numel in entries:
            t = items[name]
            flat = params if kind == 0 else bn
            view = flat[off:off + numel].view(t.shape)
            view.copy_(t.detach().to(device=dev, dtype=torch.float32))
            t.data = view                                   # the Parameter / buffer object stays, its storage moves
            if kind == 0:
                t.grad = grads[off:off + numel].view(t.shape)
        self._flat = dict(params=params, grads=grads, bn=bn, entries=entries)
        arrays = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()
                  if not k.startswith('smpl.') and v.is_floating_point()}
        arrays = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in arrays.items()}
        arrays.update(self.smpl.submodel_arrays())
        self._trainer = _lib.TrainContext(cfg, arrays, params, grads, bn if n_buffers > 0 else None)
        self._trainer_device = index
        return self._trainer

    def flat_gradients(self):
        """The single flat gradient vector (None before the first training step)."""
        return None if self._flat is None else self._flat['grads']

    def flat_parameters(self):
        return None if self._flat is None else self._flat['params']

    def sync_batchnorm(self, enabled=True, group=None, device=None):
        """Data-parallel training with the statistics of the GLOBAL batch in every BatchNorm1d (``layers.py:26,57``), i.e.
        what ``torch.nn.SyncBatchNorm.convert_sync_batchnorm`` + DDP would give the reference: N ranks on B / N windows
        each then take the step one device takes on B windows.  Every rank must run the same number of rows per step."""
        dev = device if device is not None else next(self.parameters()).device
        self.trainer(dev).set_sync_batchnorm(enabled, group)

    def allreduce_gradients(self, average=True):
        """Data-parallel training: ONE all-reduce over the flat gradient vector.  A no-op when
        ``overlap_gradient_allreduce`` already reduced this step's gradients inside ``backward``."""
        import torch.distributed as dist
        from empose_b200 import sharding
        if self._grads_reduced:
            self._grads_reduced = False
            return
        if average:
            sharding.allreduce_mean_(self._flat['grads'], dist)
        else:
            dist.all_reduce(self._flat['grads'], op=dist.ReduceOp.SUM)

    _grads_reduced = False
    _overlap = None

    def overlap_gradient_allreduce(self, enable=True, average=True, group=None):
        """
        Reduce the gradients INSIDE ``backward`` in two buckets: the iter-MLPs and heads (everything but the
        LSTM) as soon as they are final -- the native pass records an event before the LSTM's backward-through-time sweep
        -- on a side stream, so that the NCCL transfer runs under the sweep; the LSTM bucket when the sweep has finished.
        ``allreduce_gradients`` then becomes a no-op for that step.  Needs an initialised ``torch.distributed``.
        """
        self._overlap = dict(average=average, group=group, stream=None, event=None) if enable else None

    def lstm_bucket_end(self):
        """Number of floats at the start of the flat vectors that belong to the LSTM (``empose_train_layout`` puts the
        ``rnn.lstm.*`` tensors first); 0 for models without the RNN initialisation."""
        end = 0
        for name, kind, off, numel in self._flat['entries']:
            if kind == 0 and name.startswith('rnn.lstm.'):
                end = max(end, off + numel)
        for name, kind, off, numel in self._flat['entries']:
            assert kind != 0 or name.startswith('rnn.lstm.') or off >= end, 'the LSTM must be the first bucket of the flat vector'
        return end

    def _attach_gradients(self):
        """``optimizer.zero_grad()`` sets ``.grad`` to None by default: re-attach (zeroed) views of the flat vector."""
        items = dict(self.named_parameters())
        grads = self._flat['grads']
        todo = []
        n_params = 0
        for name, kind, off, numel in self._flat['entries']:
            if kind != 0:
                continue
            n_params += 1
            p = items[name]
            view = grads[off:off + numel].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                todo.append((p, view))
        if todo and len(todo) == n_params:
            grads.zero_()                    # the usual case (zero_grad() dropped every .grad): ONE fill instead of one per tensor
            for p, view in todo:
                p.grad = view
        else:
            for p, view in todo:
                view.zero_()
                p.grad = view

    def _forward_train(self, batch, window_size):
        if window_size is not None and window_size < batch.seq_length:
            raise NotImplementedError('windowed evaluation is an inference feature; train on whole windows')
        inputs = batch.get_inputs()
        marker_pos = inputs['marker_pos']
        tr = self.trainer(marker_pos.device)
        self._attach_gradients()
        res = tr.forward(marker_pos, inputs['marker_oris'], inputs['offset_r'], inputs['offset_t'], batch.seq_lengths,
                         marker_masks=inputs.get('marker_masks'), want_history=True)
        for mod, times in ((self.pose_net_iter, self.N), (self.shape_net_iter, self.N),
                           (None if self.rnn_init else self.pose_net_init, 1),
                           (None if self.rnn_init else self.shape_net_init, 1)):
            if mod is None:
                continue
            for m in mod.modules():
                if isinstance(m, nn.BatchNorm1d):
                    m.num_batches_tracked += times
        n1 = self.N + 1
        h = res['history']
        self.pose_hat_history = [h['pose'][i] for i in range(n1)]
        self.shape_hat_history = [h['shape'][i] for i in range(n1)]
        self.joints_hat_history = [h['joints'][i] for i in range(n1)]
        self.markers_hat_history = [h['markers'][i] for i in range(n1)]
        self.markers_ori_hat_history = [h['markers_ori'][i] for i in range(n1)]
        pose = res['pose']
        return {'pose_hat': pose[:, :, 3:], 'root_ori_hat': pose[:, :, :3], 'shape_hat': res['shape'],
                'joints_hat': res['joints']}

    def _backward_train(self, batch, writer, global_step):
        pose_gt = torch.cat([batch.poses_root, batch.poses_body], dim=-1)
        joints_gt = batch.joints_gt if self.do_fk else None
        fk_w = self.fk_loss_weight if self.do_fk else 0.0
        if self._overlap is None:
            vals = self._trainer.backward(pose_gt, batch.shapes, joints_gt, self.pose_weight, self.shape_weight, self.r_weight, fk_w)
        else:
            import torch.distributed as dist
            ov = self._overlap
            main = torch.cuda.current_stream(pose_gt.device)
            if ov['stream'] is None:
                ov['stream'] = torch.cuda.Stream(device=pose_gt.device)
                ov['event'] = torch.cuda.Event()
                ov['event'].record(main)                       # creates the handle the native pass re-records
            self._trainer.backward(pose_gt, batch.shapes, joints_gt, self.pose_weight, self.shape_weight, self.r_weight, fk_w,
                                   defer=True, dense_ready_event=ov['event'])
            grads, cut = self._flat['grads'], self.lstm_bucket_end()
            op = dist.ReduceOp.AVG if ov['average'] else dist.ReduceOp.SUM
            ov['stream'].wait_event(ov['event'])
            with torch.cuda.stream(ov['stream']):              # dense bucket: under the LSTM sweep still running on `main`
                dist.all_reduce(grads[cut:], op=op, group=ov['group'])
            if cut > 0:
                dist.all_reduce(grads[:cut], op=op, group=ov['group'])
            main.wait_stream(ov['stream'])
            self._grads_reduced = True
            vals = self._trainer.loss_values()
        if writer is not None:
            for k in vals:
                writer.add_scalar('{}/{}'.format(k, 'train'), vals[k], global_step)
        total = torch.tensor([vals['total_loss']], dtype=torch.float32, device=pose_gt.device)
        return total, vals

    # ------------------------------------------------------------------------------------------------
    def forward(self, batch, window_size=None, is_new_sequence=True):
        """``models.py:485-632``.  ``batch`` is an ``ABatch`` (anything with ``get_inputs``, ``seq_lengths``,
        ``batch_size`` and ``seq_length``).  Works inside ``torch.no_grad()`` like the reference."""
        if self.training:
            return self._forward_train(batch, window_size)
        if self.rnn_init:
            if is_new_sequence:
                self.rnn.final_state = None
            self.rnn.init_state = self.rnn.final_state

        seq_len = batch.seq_length
        if window_size is None:
            spans = [(None, None)]
        else:                                                                # models.py:146-159
            n_windows = seq_len // window_size + int(seq_len % window_size > 0)
            spans = [(i * window_size, min((i + 1) * window_size, seq_len)) for i in range(n_windows)]

        outs, hists = [], []
        for sf, ef in spans:
            inputs = batch.get_inputs(sf=sf, ef=ef) if sf is not None else batch.get_inputs()
            marker_pos = inputs['marker_pos']
            if sf is None:
                lengths = batch.seq_lengths
            else:
                lengths = torch.full((marker_pos.shape[0],), ef - sf, dtype=torch.int32, device=marker_pos.device)
            state = None
            if self.rnn_init and self.rnn.final_state is not None:
                state = torch.stack([self.rnn.final_state[0], self.rnn.final_state[1]])
            run = lambda ctx: ctx.forward(marker_pos, inputs['marker_oris'], inputs['offset_r'], inputs['offset_t'], lengths,
                                          marker_masks=inputs.get('marker_masks'), lstm_state=state,
                                          is_new_sequence=state is None, want_history=True,
                                          history_out=self._hist_buf if len(spans) == 1 else None)
            ctx = self.native_context(marker_pos.device)
            res = run(ctx)
            if (self.precision_guard and ctx.config['precision'] == _lib.PRECISION_FP16 and
                    not bool(torch.isfinite(res['pose']).all() & torch.isfinite(res['joints']).all())):
                # an activation left fp16's range: the documented fallback is the same pass on tf32 tensor cores
                self.precision_fallbacks += 1
                res = run(self.native_context(marker_pos.device, precision=_lib.PRECISION_TF32))
            if self.rnn_init:
                self.rnn.init_state = self.rnn.final_state
                self.rnn.final_state = (res['lstm_state'][0], res['lstm_state'][1])
            outs.append(res)
            hists.append(res['history'])

        n1 = self.N + 1
        if len(outs) == 1:          # the reference's callers: one span, nothing to concatenate (and nothing to copy)
            cat = lambda key: outs[0][key]
            hist_cat = {k: [hists[0][k][i] for i in range(n1)] for k in hists[0]}
            self._hist_buf = hists[0]
        else:
            cat = lambda key: torch.cat([o[key] for o in outs], dim=1)
            hist_cat = {k: [torch.cat([h[k][i] for h in hists], dim=1) for i in range(n1)] for k in hists[0]}
        self.pose_hat_history = hist_cat['pose']
        self.shape_hat_history = hist_cat['shape']
        self.joints_hat_history = hist_cat['joints']
        self.markers_hat_history = hist_cat['markers']
        self.markers_ori_hat_history = hist_cat['markers_ori']
        pose = cat('pose')
        return {'pose_hat': pose[:, :, 3:], 'root_ori_hat': pose[:, :, :3], 'shape_hat': cat('shape'),
                'joints_hat': cat('joints')}

    # ------------------------------------------------------------------------------------------------
    def prepare_inputs(self, batch_inputs):
        """``models.py:106-125`` (host-side view used by ``backward``; the kernels do their own gather)."""
        n, f = batch_inputs['marker_pos'].shape[0], batch_inputs['marker_pos'].shape[1]
        m_pos = batch_inputs['marker_pos'].reshape((n, f, -1, 3))
        m_ori = batch_inputs['marker_oris'].reshape((n, f, -1, 3, 3))
        if self.n_markers == 6:
            m_pos, m_ori = m_pos[:, :, C.S_CONFIG_6], m_ori[:, :, C.S_CONFIG_6]
        parts = []
        if self.config.use_marker_pos:
            parts.append(m_pos.reshape((n, f, -1)))
        if self.config.use_marker_ori:
            parts.append(m_ori.reshape((n, f, -1)))
        return torch.cat(parts, dim=-1)

    @staticmethod
    def _masked_mean(per_frame, seq_lengths):
        mask = _frame_mask(seq_lengths, per_frame.shape[1]).to(per_frame.dtype)
        return ((per_frame * mask).sum(-1) / seq_lengths.to(per_frame.dtype)).mean()

    def _recon(self, gt, hat, seq_lengths, marker_masks):
        """``loss.py:23-41``."""
        diff = hat - gt
        per_frame = torch.sqrt((diff * diff).sum(dim=-1)).sum(dim=-1)
        if marker_masks is not None:
            per_frame = per_frame * marker_masks.logical_not().any(dim=-1).logical_not()
        return self._masked_mean(per_frame, seq_lengths)

    def backward(self, batch, model_out, writer=None, global_step=None):
        """
        ``models.py:634-688``: the training loss over all N+1 iterates, evaluated from the histories of the
        last ``forward``.  Returns ``(total_loss, loss_vals)``.  The loss value is what
        ``empose/eval/helpers.py:86`` logs during validation.  In training mode the native backward pass runs
        instead and the gradients land in ``.grad`` (the reference calls ``total_loss.backward()`` here).
        """
        if self.training:
            return self._backward_train(batch, writer, global_step)
        bsz, n_frames = batch.batch_size, batch.seq_length
        inputs_ = self.prepare_inputs(batch.get_inputs())
        markers_in = inputs_[:, :, self.pos_d_start:self.pos_d_end].reshape((bsz, n_frames, -1, 3))
        markers_ori_in = inputs_[:, :, self.ori_d_start:self.ori_d_end].reshape((bsz, n_frames, -1, 9))
        lengths = batch.seq_lengths
        dev = inputs_.device
        zero = lambda: torch.zeros(1, device=dev)
        recon_t, shape_t, pose_t, fk_t = zero(), zero(), zero(), zero()
        n_hist = len(self.pose_hat_history)
        pose_gt = torch.cat([batch.poses_root, batch.poses_body], dim=-1)
        shape_gt = batch.shapes.unsqueeze(1).repeat((1, n_frames, 1))
        for i in range(n_hist):
            pose_t += self._masked_mean((pose_gt - self.pose_hat_history[i]).abs().mean(-1), lengths)
            shape_t += self._masked_mean((shape_gt - self.shape_hat_history[i]).abs().mean(-1), lengths)
            if self.do_fk:
                fk_t += self._recon(batch.joints_gt.reshape(bsz, n_frames, -1, 3),
                                    model_out['joints_hat'].reshape(bsz, n_frames, -1, 3), lengths, batch.marker_masks)
            if self.config.use_marker_pos:
                hat = self.markers_hat_history[i].reshape((bsz, n_frames, -1, 3))[:, :, self.marker_idxs]
                recon_t += self._recon(markers_in, hat, lengths, batch.marker_masks)
            if self.config.use_marker_ori:
                hat = self.markers_ori_hat_history[i].reshape((bsz, n_frames, -1, 9))[:, :, self.marker_idxs]
                recon_t += self._recon(markers_ori_in, hat, lengths, batch.marker_masks)
        total = (self.pose_weight * pose_t + self.fk_loss_weight * fk_t + self.shape_weight * shape_t +
                 self.r_weight * recon_t) / n_hist
        loss_vals = {'pose': pose_t.cpu().item() / n_hist, 'shape': shape_t.cpu().item() / n_hist,
                     'reconstruction': recon_t.cpu().item() / n_hist, 'fk': fk_t.cpu().item() / n_hist,
                     'total_loss': total.cpu().item()}
        if writer is not None:
            prefix = 'train' if self.training else 'valid'
            for k in loss_vals:
                writer.add_scalar('{}/{}'.format(k, prefix), loss_vals[k], global_step)
        return total, loss_vals


class SimpleRNN(nn.Module):
    """The uni- / bidirectional RNN baseline (``models.py:265-366``; "BiRNN" in the paper), inference on the B200.

    Same constructor arguments, ``forward`` signature and outputs, ``rnn.final_state`` carry and state-dict keys as the
    reference; ``backward`` returns the loss values in eval mode (``eval/helpers.py:86``).  Training this baseline is
    not built."""

    def __init__(self, config, smpl_layer=None, precision=_lib.PRECISION_FP16):
        super(SimpleRNN, self).__init__()
        self.config = config
        self.n_markers = config.n_markers if getattr(config, 'n_markers', -1) > -1 else C.N_TRACKERS_WO_ROOT
        assert self.n_markers in [6, 12]
        self.n_frames = config.window_size
        self.smpl = smpl_layer
        self.estimate_shape = config.m_estimate_shape
        self.shape_avg = config.m_average_shape
        self.fk_loss_weight = config.m_fk_loss
        self.do_fk = self.fk_loss_weight > 0.0
        if self.do_fk:
            assert self.smpl is not None
            assert self.estimate_shape                                       # models.py:55
        self.shape_weight = getattr(config, 'm_shape_loss_weight', 1.0)
        self.pose_weight = getattr(config, 'm_pose_loss_weight', 1.0)
        self.precision = precision
        if getattr(config, 'use_marker_nor', False):
            raise ValueError('Normals currently not supported.')
        if getattr(config, 'm_learn_init_state', False):
            raise NotImplementedError('m_learn_init_state is not supported by empose_b200')
        input_size = 0
        if config.use_marker_pos:
            input_size += self.n_markers * 3
        if config.use_marker_ori:
            input_size += self.n_markers * 9
        self.input_size, self.output_size = input_size, (C.N_JOINTS + 1) * 3
        setattr(config, 'input_size', input_size)
        setattr(config, 'output_size', self.output_size)
        hidden = config.m_hidden_size
        dirs = 2 if config.m_bidirectional else 1
        self.rnn = RNNLayer(input_size, hidden, config.m_num_layers, bidirectional=config.m_bidirectional,
                            dropout=config.m_dropout, learn_init_state=False)
        self.to_pose = nn.Linear(hidden * dirs, self.output_size)
        if self.estimate_shape:
            self.to_shape = MLP(input_size=hidden * dirs, output_size=C.N_SHAPE_PARAMS, hidden_size=config.m_shape_hidden_size,
                                num_layers=2, dropout_p=config.m_dropout_hidden, skip_connection=config.m_skip_connections,
                                use_batch_norm=False)
            if config.m_skip_connections:
                raise NotImplementedError('skip connections in to_shape are not supported')
        else:
            self.to_shape = None
        self.shape_loss = nn.L1Loss(reduction='none')
        self._ctx = None
        self._ctx_key = None

    def model_name(self):
        """``models.py:283-289`` + ``BaseModel.model_name`` (``:84-94``)."""
        c = self.config
        name = "RNN-{}".format('-'.join([str(c.m_hidden_size)] * c.m_num_layers))
        if c.m_bidirectional:
            name = "Bi" + name
        if self.estimate_shape is not None:
            name += '-shape{}{}'.format(c.m_shape_hidden_size, '-avg' if self.shape_avg else '')
        if self.do_fk:
            name += '-fk{}'.format(self.fk_loss_weight)
        name += '-n{}'.format(self.n_markers)
        name += '-lr{}'.format(c.lr)
        return name

    def _native_config(self, device_index):
        c = self.config
        return dict(n_markers=self.n_markers, hidden_size=int(c.m_hidden_size), num_layers=int(c.m_num_layers),
                    bidirectional=int(bool(c.m_bidirectional)), learn_init_state=0, estimate_shape=int(bool(self.estimate_shape)),
                    shape_hidden_size=int(c.m_shape_hidden_size), average_shape=int(bool(self.shape_avg)), do_fk=int(self.do_fk),
                    use_marker_pos=int(c.use_marker_pos), use_marker_ori=int(c.use_marker_ori), precision=int(self.precision),
                    device=int(device_index))

    def invalidate(self):
        """Drop the packed weights: the next ``forward`` re-reads the module's tensors (needed after writes through
        ``.data``, which neither change ``data_ptr`` nor bump ``_version``; see ``IterativeErrorFeedback.invalidate``)."""
        if self._ctx is not None:
            self._ctx.close()
        self._ctx, self._ctx_key = None, None

    def native_context(self, device):
        if device.type != 'cuda':
            raise _lib.EmposeError('empose_b200 runs on CUDA devices only (no CPU fallback); got %s' % device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        tensors = [t for k, t in self.state_dict(keep_vars=True).items()]
        key = (index, self.precision) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self._ctx is None or key != self._ctx_key:
            if self._ctx is not None:
                self._ctx.close()
            arrays = {k: np.ascontiguousarray(v.detach().cpu().numpy(), dtype=np.float32) for k, v in self.state_dict().items()
                      if not k.startswith('smpl.') and v.is_floating_point()}
            if self.do_fk:
                arrays.update(self.smpl.submodel_arrays())
            self._ctx = _lib.RnnContext(self._native_config(index), arrays)
            self._ctx_key = key
        return self._ctx

    def forward(self, batch, window_size=None, is_new_sequence=True):
        """``models.py:291-317``.  ``window_size`` is ignored, as in the reference."""
        if self.training:
            raise NotImplementedError('training the (Bi)RNN baseline is not built in empose_b200; call net.eval()')
        if is_new_sequence:
            self.rnn.final_state = None
        self.rnn.init_state = self.rnn.final_state
        inputs = batch.get_inputs()
        marker_pos = inputs['marker_pos']
        ctx = self.native_context(marker_pos.device)
        state = None
        if self.rnn.final_state is not None:
            state = torch.stack([self.rnn.final_state[0], self.rnn.final_state[1]])
        res = ctx.forward(marker_pos, inputs['marker_oris'], batch.seq_lengths, lstm_state=state, is_new_sequence=state is None)
        self.rnn.final_state = (res['lstm_state'][0], res['lstm_state'][1])
        pose = res['pose']
        return {'pose_hat': pose[:, :, 3:], 'root_ori_hat': pose[:, :, :3], 'shape_hat': res['shape'], 'joints_hat': res['joints']}

    def backward(self, batch, model_out, writer=None, global_step=None):
        """``models.py:319-366`` loss values (eval mode)."""
        if self.training:
            raise NotImplementedError('training the (Bi)RNN baseline is not built in empose_b200')
        pose_hat, root_ori_hat, shape_hat = model_out['pose_hat'], model_out['root_ori_hat'], model_out['shape_hat']
        n, f = batch.batch_size, batch.seq_length
        lengths = batch.seq_lengths

        def normal_mse(gt, hat):                                             # loss.py:44-62
            diff = hat - gt
            per_frame = (diff * diff).sum(dim=-1).sum(dim=-1)
            if batch.marker_masks is not None:
                per_frame = per_frame * batch.marker_masks.logical_not().any(dim=-1).logical_not()
            return IterativeErrorFeedback._masked_mean(per_frame, lengths)

        pose_loss = normal_mse(batch.poses_body.reshape(n, f, -1, 3), pose_hat.reshape(n, f, -1, 3))
        root_loss = normal_mse(batch.poses_root.reshape(n, f, -1, 3), root_ori_hat.reshape(n, f, -1, 3))
        dev = pose_hat.device
        shape_loss = torch.zeros(1, device=dev)
        if self.estimate_shape:
            gt = batch.shapes.unsqueeze(1).repeat((1, f, 1))
            shape_loss = IterativeErrorFeedback._masked_mean((gt - shape_hat).abs().mean(-1), lengths)
        fk_loss = torch.zeros(1, device=dev)
        if self.do_fk:
            diff = model_out['joints_hat'].reshape(n, f, -1, 3) - batch.joints_gt.reshape(n, f, -1, 3)
            per_frame = torch.sqrt((diff * diff).sum(dim=-1)).sum(dim=-1)
            if batch.marker_masks is not None:
                per_frame = per_frame * batch.marker_masks.logical_not().any(dim=-1).logical_not()
            fk_loss = IterativeErrorFeedback._masked_mean(per_frame, lengths)
        total = pose_loss + root_loss + shape_loss + self.fk_loss_weight * fk_loss
        loss_vals = {'pose': float(pose_loss), 'root_pose': float(root_loss), 'shape': float(shape_loss), 'fk': float(fk_loss),
                     'total_loss': float(total)}
        if writer is not None:
            for k in loss_vals:
                writer.add_scalar('{}/{}'.format(k, 'valid'), loss_vals[k], global_step)
        return total, loss_vals
