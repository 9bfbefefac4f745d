"""CPU-side checks of the boundary: the library loads, exports every declared symbol, the Python module
mirrors the reference's state-dict layout, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from empose_b200 import lib as native
from empose_b200 import synthetic

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built_lib():
    from empose_b200 import build
    build.build_library()
    return native.load()


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, 'include', 'empose_b200.h')).read()
    declared = set(re.findall(r'\b(empose_[a-z_0-9]+)\s*\(', header))
    declared = {d for d in declared if not d.endswith('_t')}
    assert declared == set(native.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.empose_abi_version() == 1


def test_struct_layouts_match_header(built_lib):
    assert ctypes.sizeof(native.IefConfig) == 16 * 4
    assert ctypes.sizeof(native.Tensor) == 8 + 8 + 4 + 4 + 32
    assert ctypes.sizeof(native.History) == 5 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_create_fails_loudly_without_gpu(built_lib, smpl_npz):
    net = util.build_module(smpl_npz)
    with pytest.raises(native.EmposeError):
        net.native_context(torch.device('cpu'))
    cfg = net._native_config(0)
    arrays = {k: v.numpy() for k, v in net.state_dict().items() if not k.startswith('smpl.') and v.is_floating_point()}
    arrays.update(net.smpl.submodel_arrays())
    with pytest.raises(native.EmposeError, match='no CUDA device|CUDA'):
        native.IefContext(cfg, arrays)


def test_bad_config_is_rejected_before_touching_the_gpu(built_lib):
    cfg = native.IefConfig()
    cfg.n_markers = 7
    handle = ctypes.c_void_p()
    table, _ = native.make_tensor_table({'x': np.zeros(1, dtype=np.float32)})
    rc = built_lib.empose_ief_create(ctypes.byref(cfg), table, 1, ctypes.byref(handle))
    assert rc == -1 and b'n_markers' in built_lib.empose_last_error()


@pytest.mark.parametrize('n_markers,rnn_init', [(12, True), (6, True), (12, False)])
def test_state_dict_keys_match_reference_layout(smpl_npz, n_markers, rnn_init):
    net = util.build_module(smpl_npz, n_markers=n_markers, rnn_init=rnn_init)
    keys = set(net.state_dict().keys())
    learned = {k for k, _, _ in synthetic.lgd_state_dict_spec(n_markers=n_markers, rnn_init=rnn_init)}
    smpl = {'smpl.bm.' + k for k in ('trans', 'root_orient', 'pose_body', 'pose_hand', 'betas', 'v_template', 'f',
                                      'shapedirs', 'J_regressor', 'posedirs', 'kintree_table', 'weights')}
    assert keys == learned | smpl            # SURVEY.md section 8b key list
    n_params = sum(p.numel() for p in net.parameters() if p.requires_grad)
    if n_markers == 6 and rnn_init:
        pass  # N does not change the count; README.md:228 quotes 5721419 for this layout
    gold = util.load_golden({(12, True): 'lgd_rnn12_n4', (6, True): 'lgd_rnn6_n2_real', (12, False): 'lgd_mlp12_n4'}[(n_markers, rnn_init)])
    assert n_params == int(gold['n_trainable_params'])


def test_training_refuses_cpu_tensors(smpl_npz):
    """No CPU fallback in training either: the train-mode forward needs CUDA tensors."""
    net = util.build_module(smpl_npz).train()
    p = synthetic.synth_window_params(2, 4, seed=3)
    z = torch.zeros
    batch = util.DuckBatch(z(2, 4, 36), z(2, 4, 108), torch.from_numpy(p['offset_r']), torch.from_numpy(p['offset_t']),
                           torch.from_numpy(p['seq_lengths']))
    with pytest.raises(native.EmposeError):
        net(batch)


def test_training_layout_covers_every_trainable_tensor(smpl_npz):
    """empose_train_layout: one flat vector holds exactly the reference's trainable tensors (README.md:228 count)."""
    for n_markers, rnn_init in ((6, True), (12, True), (12, False)):
        net = util.build_module(smpl_npz, n_markers=n_markers, num_iterations=2, rnn_init=rnn_init)
        entries, n_params, n_buffers = native.train_layout(net._native_config(0))
        params = {k: v for k, v in net.named_parameters() if not k.startswith('smpl.')}
        got = {name: numel for name, kind, off, numel in entries if kind == 0}
        assert got == {k: v.numel() for k, v in params.items()}
        bufs = {k: v.numel() for k, v in net.named_buffers() if 'running_' in k}
        assert {name: numel for name, kind, off, numel in entries if kind == 1} == bufs
        spans = sorted((off, off + numel) for name, kind, off, numel in entries if kind == 0)
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= n_params
        assert all(off % 4 == 0 for off, _ in spans)
        if n_markers == 6 and rnn_init:
            assert sum(got.values()) + 169 == 5721419


def test_rnn_state_dict_keys_match_reference_layout(smpl_npz):
    """SimpleRNN mirror: the reference's keys for the bidirectional LSTM, to_pose and the BatchNorm-free to_shape MLP."""
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import Configuration
    from empose_b200.nn.models import SimpleRNN, create_model
    cfg = Configuration(dict(m_type='rnn', m_hidden_size=1024, m_num_layers=2, m_bidirectional=True, m_estimate_shape=True,
                             m_average_shape=True, m_fk_loss=0.1, use_marker_pos=True, use_marker_ori=True, n_markers=12))
    net = create_model(cfg, SMPLLayer(smpl_npz).to(dtype=torch.float32))
    assert isinstance(net, SimpleRNN)
    keys = {k for k in net.state_dict() if not k.startswith('smpl.')}
    spec = {k for k, _, _ in synthetic.rnn_state_dict_spec(n_markers=12, hidden_size=1024, num_layers=2, bidirectional=True,
                                                           estimate_shape=True)}
    assert keys == spec
    gold = util.load_golden('rnn_bi12_shape_fk')
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == int(gold['n_trainable_params'])
    assert net.model_name() == 'BiRNN-1024-1024-shape256-avg-fk0.1-n12-lr0.001'       # what the reference printed for this config
    with pytest.raises(native.EmposeError):
        net.eval()(util.DuckBatch(torch.zeros(1, 2, 36), torch.zeros(1, 2, 108), None, None, torch.tensor([2])))


def test_submodel_arrays_are_complete(smpl_npz):
    net = util.build_module(smpl_npz)
    sub = net.smpl.submodel_arrays()
    assert sub['sub.dims'].tolist()[-1] == 12 and sub['sub.posedirs'].shape[0] == 189
    assert set(sub) >= {'sub.v_template', 'sub.shapedirs', 'sub.posedirs', 'sub.j0', 'sub.jdirs', 'sub.skin_weight',
                        'sub.skin_joint', 'sub.jt_ptr', 'sub.jt_vert', 'sub.jt_weight', 'sub.vj_ptr', 'sub.jvj_ptr', 'sub.vinc_ptr', 'sub.vinc_item', 'sub.vinc_code', 'sub.parents', 'sub.faces',
                        'sub.sensor_vert', 'sub.helper_vert', 'sub.sensor_faces', 'sub.sensor_degree', 'sub.dims'}


def test_dropin_rebinds_the_reference_factories(smpl_npz, asset_dir):
    """With the reference tree present, dropin.install() makes the reference's own factory build our class."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip('reference tree not present on this machine')
    ref_shims.install(asset_dir, seed=0)
    import empose_b200.dropin
    from empose_b200.nn.models import IterativeErrorFeedback
    ref_models, ref_smpl = empose_b200.dropin.install()
    flags = ['--m_type', 'lgd', '--m_num_iterations', '2', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
             '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '6', '--window_size', '32']
    cfg = ref_shims.make_config(flags)                      # the reference's own Configuration object
    smpl = ref_smpl.SMPLLayer(smpl_npz)
    net = ref_models.create_model(cfg, smpl)
    assert isinstance(net, IterativeErrorFeedback) and isinstance(net, ref_models.IterativeErrorFeedback)
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 5721419     # README.md:228
    assert net.model_name().startswith('IEF-2x512-N2-RNN-2x512')


def test_dropin_rebinds_names_imported_before_install(smpl_npz, asset_dir):
    """The reference binds its factories with ``from empose.nn.models import create_model`` (eval/helpers.py:20-25): a module
    imported BEFORE install() keeps the old objects unless install() also rebinds them there."""
    import sys
    import types
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip('reference tree not present on this machine')
    ref_shims.install(asset_dir, seed=0)
    import empose.bodymodels.smpl as ref_smpl
    import empose.nn.models as ref_models
    import empose_b200.dropin
    from empose_b200.nn.models import IterativeErrorFeedback
    early = types.ModuleType('empose._imported_before_install')          # stands for empose.eval.helpers
    early.create_model, early.IterativeErrorFeedback = ref_models.create_model, ref_models.IterativeErrorFeedback
    early.create_default_smpl_model, early.SMPLLayer = ref_smpl.create_default_smpl_model, ref_smpl.SMPLLayer
    sys.modules[early.__name__] = early
    try:
        empose_b200.dropin.install()
        assert early.IterativeErrorFeedback is IterativeErrorFeedback
        assert early.create_model is ref_models.create_model and early.SMPLLayer is ref_smpl.SMPLLayer
        flags = ['--m_type', 'lgd', '--m_num_iterations', '2', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
                 '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '6', '--window_size', '32']
        net = early.create_model(ref_shims.make_config(flags), early.SMPLLayer(smpl_npz))
        assert isinstance(net, IterativeErrorFeedback)
    finally:
        sys.modules.pop(early.__name__, None)


def test_fp16_guard_static_range_check_and_invalidate(smpl_npz):
    """The precision guard of the mirror class: BatchNorm-folded weights outside fp16's comfortable range
    select tf32 before anything runs; wide running statistics alone do not (the fold happens in double); invalidate()
    changes the cache key that .data writes leave untouched."""
    net = util.build_module(smpl_npz, precision=native.PRECISION_FP16)
    hi, lo = net.folded_weight_range()
    assert 0.0 < lo < hi < 16.0
    assert net._effective_precision() == native.PRECISION_FP16
    sd = net.state_dict()
    # (i) BN statistics over six orders of magnitude, the Linear in front scaled to match: same folded weights
    g = torch.Generator().manual_seed(3)
    for name in [k for k in sd if k.endswith('running_var')]:
        var = torch.exp(torch.empty_like(sd[name]).uniform_(np.log(1e-4), np.log(1e2), generator=g))
        ratio = torch.sqrt((var + 1e-5) / (sd[name] + 1e-5))
        lin = name.replace('batch_norm.running_var', 'input_to_hidden') if 'batch_norm' in name else \
            '.'.join(name.split('.')[:-2] + [str(int(name.split('.')[-2]) - 1)])
        sd[lin + '.weight'] = sd[lin + '.weight'] * ratio.reshape(-1, 1)
        sd[lin + '.bias'] = sd[lin + '.bias'] * ratio
        sd[name.replace('running_var', 'running_mean')] = sd[name.replace('running_var', 'running_mean')] * ratio
        sd[name] = var
    net.load_state_dict(sd)
    hi2, _ = net.folded_weight_range()
    assert abs(hi2 - hi) < 1e-3 * hi and net._effective_precision() == native.PRECISION_FP16
    # (ii) a genuinely huge layer: tf32 from the start
    with torch.no_grad():
        net.pose_net_iter.hidden_to_output.weight.mul_(1e4)
    assert net._effective_precision() == native.PRECISION_TF32
    net.precision_guard = False
    assert net._effective_precision() == native.PRECISION_FP16
    # (iii) .data writes change neither data_ptr nor _version: only invalidate() moves the key
    k0 = net._weights_key(0, net.precision)
    net.pose_net_iter.hidden_to_output.weight.data.mul_(0.5)
    assert net._weights_key(0, net.precision) == k0
    net.invalidate()
    assert net._weights_key(0, net.precision) != k0
