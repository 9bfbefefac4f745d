"""Synthetic assets: an SMPL-H ``model.npz`` stand-in, network weights and sensor windows.

Used by the tests, ``bench.py`` and ``__graft_entry__.smoke()`` (there is no network for the real
assets); also handy for users who want to try the path without the licensed model.

The licensed SMPL-H model the reference loads at
``empose/bodymodels/smpl.py:26`` (``$SMPL_MODELS/smplh_amass/neutral/model.npz``)
cannot be shipped, so tests and the benchmark run on a synthetic stand-in with the
same keys, shapes and dtypes:

    v_template (6890,3)  f (13776,3)  shapedirs (6890,3,16)  posedirs (6890,3,459)
    J_regressor (52,6890)  kintree_table (2,52)  weights (6890,52)

The surface is a closed 2-manifold (ellipsoid, 84 latitude rings x 82 segments
plus two poles = 6890 vertices / 13776 triangles) so that every sensor vertex id
of ``empose/helpers/configuration.py:32-34`` has a complete 1-ring.  Skinning
weights are 4-sparse like the real model and deliberately put hand-joint mass on
vertices near the wrists so the hand-folding identity used by the CUDA path is
exercised.  Everything is drawn from ``numpy.random.RandomState`` (bit-stable
across numpy versions), so the same seed gives the same model on every machine.
"""
import os

import numpy as np

N_VERTS = 6890
N_FACES = 13776
N_JOINTS_SMPLH = 52
N_RINGS, N_SEGS = 84, 82

#: SMPL-H kinematic tree: 22 body joints (reference configuration.py:118) followed by
#: 15 joints per hand hanging off the wrists 20 / 21 (three-joint finger chains).
BODY_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]


def smplh_parents():
    parents = list(BODY_PARENTS)
    for wrist in (20, 21):
        for _finger in range(5):
            base = len(parents)
            parents += [wrist, base, base + 1]
    assert len(parents) == N_JOINTS_SMPLH
    return parents


def _ellipsoid_mesh():
    """Lat-long ellipsoid; returns (vertices (6890,3), faces (13776,3)) with outward winding."""
    polar = np.linspace(0.0, np.pi, N_RINGS + 2)[1:-1]
    azim = np.arange(N_SEGS) * (2.0 * np.pi / N_SEGS)
    pp, aa = np.meshgrid(polar, azim, indexing='ij')
    ax, ay, az = 0.28, 0.85, 0.18
    body = np.stack([ax * np.sin(pp) * np.cos(aa), ay * np.cos(pp), az * np.sin(pp) * np.sin(aa)], axis=-1)
    verts = np.concatenate([[[0.0, ay, 0.0]], body.reshape(-1, 3), [[0.0, -ay, 0.0]]], axis=0)

    def vid(ring, seg):
        return 1 + ring * N_SEGS + (seg % N_SEGS)

    faces = []
    for s in range(N_SEGS):  # north cap
        faces.append((0, vid(0, s + 1), vid(0, s)))
    for r in range(N_RINGS - 1):
        for s in range(N_SEGS):
            p00, p01 = vid(r, s), vid(r, s + 1)
            p10, p11 = vid(r + 1, s), vid(r + 1, s + 1)
            faces.append((p00, p01, p10))
            faces.append((p01, p11, p10))
    south = N_VERTS - 1
    for s in range(N_SEGS):  # south cap
        faces.append((south, vid(N_RINGS - 1, s), vid(N_RINGS - 1, s + 1)))
    faces = np.asarray(faces, dtype=np.int64)
    assert verts.shape == (N_VERTS, 3) and faces.shape == (N_FACES, 3)
    return verts, faces


def _rest_joint_targets(rng):
    """Rough humanoid layout inside the ellipsoid (metres); hands fan out from the wrists."""
    body = np.array([
        [0.00, -0.05, 0.0],   # 0 root
        [0.08, -0.15, 0.0], [-0.08, -0.15, 0.0], [0.00, 0.05, 0.0],      # hips, spine1
        [0.09, -0.42, 0.0], [-0.09, -0.42, 0.0], [0.00, 0.18, 0.0],      # knees, spine2
        [0.08, -0.68, 0.0], [-0.08, -0.68, 0.0], [0.00, 0.30, 0.0],      # ankles, spine3
        [0.08, -0.76, 0.05], [-0.08, -0.76, 0.05], [0.00, 0.48, 0.0],    # feet, neck
        [0.07, 0.42, 0.0], [-0.07, 0.42, 0.0], [0.00, 0.62, 0.0],        # collars, head
        [0.15, 0.40, 0.0], [-0.15, 0.40, 0.0],                            # shoulders
        [0.19, 0.22, 0.0], [-0.19, 0.22, 0.0],                            # elbows
        [0.21, 0.04, 0.0], [-0.21, 0.04, 0.0],                            # wrists
    ])
    hands = []
    for wrist, sign in ((20, 1.0), (21, -1.0)):
        for finger in range(5):
            for knuckle in range(3):
                off = np.array([sign * 0.012 * (knuckle + 1), -0.02 * (knuckle + 1), 0.012 * (finger - 2)])
                hands.append(body[wrist] + off + rng.randn(3) * 0.002)
    return np.concatenate([body, np.asarray(hands)], axis=0)


def _flip_edge(faces, a, b):
    """Replace the edge (a, b) shared by two triangles by the other diagonal of their quad; winding is preserved."""
    def find(u, v):          # the face that contains the directed edge u -> v, rotated to (u, v, w)
        for i, f in enumerate(faces):
            for k in range(3):
                if f[k] == u and f[(k + 1) % 3] == v:
                    return i, int(f[(k + 2) % 3])
        raise ValueError('edge %d -> %d not found' % (u, v))
    i1, c = find(a, b)
    i2, d = find(b, a)
    faces[i1] = (a, d, c)
    faces[i2] = (d, b, c)


#: valence of the first sensor vertices in the irregular variants (all others keep the lat-long mesh's 6)
IRREGULAR_VALENCES = {'mild': (5, 7, 6, 7, 5), 'wild': (5, 7, 8, 9, 10, 4, 11)}


def _make_irregular(faces, kind):
    """
    Edge flips around the sensor vertices (``submodel.VERTEX_IDS``) so that their valences differ from 6 -- a real
    SMPL-H mesh is not regular.  The surface stays a closed, consistently wound 2-manifold with the same vertex and
    face counts.  'mild' keeps every ring within 8 vertices, 'wild' needs 12-vertex sensor blocks.
    """
    from empose_b200.submodel import VERTEX_IDS, sensor_ring
    faces = faces.copy()
    for v, want in zip(VERTEX_IDS, IRREGULAR_VALENCES[kind]):
        while True:
            inc = faces[(faces == v).any(axis=1)]
            ring, is_fan = sensor_ring(inc, v)
            assert is_fan
            deg = len(ring) - 1
            if deg == want:
                break
            if deg > want:
                _flip_edge(faces, v, ring[1])                 # the spoke to a neighbour goes: valence - 1
            else:
                _flip_edge(faces, ring[1], ring[2])           # the rim edge of the first face becomes a spoke: valence + 1
    return faces


def make_synthetic_smplh(seed=0, irregular=None):
    """Return a dict with the SMPL-H ``model.npz`` keys (float64 / int64).  ``irregular``: None, 'mild' or 'wild'."""
    rng = np.random.RandomState(seed)
    verts, faces = _ellipsoid_mesh()
    if irregular:
        faces = _make_irregular(faces, irregular)
    targets = _rest_joint_targets(rng)
    n_j = N_JOINTS_SMPLH

    # Joint regressor: sparse convex rows.  Each joint regresses from the 24 vertices closest to its
    # target point and from their mirror images through the target's (x, z) line, so the convex
    # combination can land inside the body.
    j_reg = np.zeros((n_j, N_VERTS))
    for j in range(n_j):
        d2 = ((verts - targets[j]) ** 2).sum(-1)
        near = np.argsort(d2)[:24]
        mirror_pt = targets[j] * np.array([1.0, 1.0, 1.0]) - (verts[near] - targets[j]) * np.array([1.0, 0.0, 1.0])
        far = np.array([np.argmin(((verts - p) ** 2).sum(-1)) for p in mirror_pt])
        ids = np.concatenate([near, far])
        w = rng.rand(ids.shape[0]) + 0.25
        np.add.at(j_reg[j], ids, w / w.sum())
    rest_joints = j_reg @ verts

    # 4-sparse skinning weights from distance to the rest joints.
    d2 = ((verts[:, None, :] - rest_joints[None, :, :]) ** 2).sum(-1)
    order = np.argsort(d2, axis=1)[:, :4]
    weights = np.zeros((N_VERTS, n_j))
    rows = np.arange(N_VERTS)
    for k in range(4):
        weights[rows, order[:, k]] = np.exp(-40.0 * d2[rows, order[:, k]]) + 1e-3
    weights /= weights.sum(axis=1, keepdims=True)

    # Blend shapes: a smooth low-frequency part plus dense noise (the real posedirs are dense).
    freq = rng.randn(16, 3, 3)
    smooth = np.stack([np.sin(verts @ freq[k].T * 3.0) for k in range(16)], axis=-1)  # (V,3,16)
    shapedirs = 0.012 * smooth + 0.004 * rng.randn(N_VERTS, 3, 16)
    posedirs = 0.002 * rng.randn(N_VERTS, 3, 459)

    parents = np.asarray(smplh_parents(), dtype=np.int64)
    kintree = np.stack([parents, np.arange(n_j, dtype=np.int64)], axis=0)
    # Every float array is made exactly float32-representable (stored as float64 like the real file):
    # the reference casts the model to float32 on load (smpl.py:27), so this way the reference, the
    # oracle and the CUDA path all start from bit-identical constants.
    r32 = lambda a: a.astype(np.float32).astype(np.float64)
    return {'v_template': r32(verts), 'f': faces, 'shapedirs': r32(shapedirs), 'posedirs': r32(posedirs),
            'J_regressor': r32(j_reg), 'kintree_table': kintree, 'weights': r32(weights)}


def write_synthetic_smplh(root_dir, seed=0, irregular=None):
    """Write ``<root_dir>/smplh_amass/neutral/model.npz`` (the path of reference smpl.py:26); returns it.
    The irregular-valence variants go to ``<root_dir>/<irregular>/smplh_amass/neutral/model.npz``."""
    if irregular:
        root_dir = os.path.join(root_dir, irregular)
    path = os.path.join(root_dir, 'smplh_amass', 'neutral', 'model.npz')
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + '.tmp%d.npz' % os.getpid()
        np.savez(tmp, **make_synthetic_smplh(seed, irregular))
        os.replace(tmp, path)
    return path


# ----------------------------------------------------------------------------------------------------------------------
# Network weights
# ----------------------------------------------------------------------------------------------------------------------

def lgd_state_dict_spec(n_markers=12, rnn_init=True, hidden_size=512, num_layers=2, rnn_hidden_size=512,
                        rnn_num_layers=2, use_gradient=True, batch_norm=True, use_marker_pos=True,
                        use_marker_ori=True):
    """
    Ordered list of (key, shape, kind) for the learned tensors of an LGD model, with the reference's
    state-dict key layout (``empose/nn/models.py:424-454``, ``empose/nn/layers.py:13-77, 114``).
    ``kind`` is one of 'w' (fan-in scaled weight), 'b' (bias), 'bn_w', 'bn_b', 'bn_mean', 'bn_var',
    'bn_count', 'prelu'.
    """
    in_size = n_markers * ((3 if use_marker_pos else 0) + (9 if use_marker_ori else 0))
    iter_in = in_size + 66 + 10 + ((66 + 10) if use_gradient else 0)
    spec = []

    def linear(prefix, n_in, n_out):
        spec.append((prefix + '.weight', (n_out, n_in), 'w'))
        spec.append((prefix + '.bias', (n_out,), 'b'))

    def bn(prefix, width):
        spec.append((prefix + '.weight', (width,), 'bn_w'))
        spec.append((prefix + '.bias', (width,), 'bn_b'))
        spec.append((prefix + '.running_mean', (width,), 'bn_mean'))
        spec.append((prefix + '.running_var', (width,), 'bn_var'))
        spec.append((prefix + '.num_batches_tracked', (), 'bn_count'))

    def mlp(prefix, n_in, n_out):
        linear(prefix + '.input_to_hidden', n_in, hidden_size)
        if batch_norm:
            bn(prefix + '.batch_norm', hidden_size)
        spec.append((prefix + '.activation_fn.weight', (1,), 'prelu'))
        linear(prefix + '.hidden_to_output', hidden_size, n_out)
        stride = 4 if batch_norm else 3
        for b in range(num_layers):
            for l in range(2):
                base = '%s.hidden_layers.%d.layers' % (prefix, b)
                linear('%s.%d' % (base, l * stride), hidden_size, hidden_size)
                if batch_norm:
                    bn('%s.%d' % (base, l * stride + 1), hidden_size)
                spec.append(('%s.%d.weight' % (base, l * stride + (2 if batch_norm else 1)), (1,), 'prelu'))

    if rnn_init:
        h = rnn_hidden_size
        for layer in range(rnn_num_layers):
            n_in = in_size if layer == 0 else h
            spec.append(('rnn.lstm.weight_ih_l%d' % layer, (4 * h, n_in), 'w_lstm'))
            spec.append(('rnn.lstm.weight_hh_l%d' % layer, (4 * h, h), 'w_lstm'))
            spec.append(('rnn.lstm.bias_ih_l%d' % layer, (4 * h,), 'b_lstm'))
            spec.append(('rnn.lstm.bias_hh_l%d' % layer, (4 * h,), 'b_lstm'))
        linear('pose_net_init', h, 66)
        linear('shape_net_init', h, 10)
    else:
        mlp('pose_net_init', in_size, 66)
        mlp('shape_net_init', in_size, 10)
    mlp('pose_net_iter', iter_in, 66)
    mlp('shape_net_iter', iter_in, 10)
    return spec


def rnn_state_dict_spec(n_markers=12, hidden_size=1024, num_layers=2, bidirectional=True, estimate_shape=False,
                        shape_hidden_size=256, use_marker_pos=True, use_marker_ori=True):
    """
    Ordered (key, shape, kind) list of the reference's ``SimpleRNN`` (``empose/nn/models.py:265-289``): a uni- or
    bidirectional LSTM (``layers.py:114``), ``to_pose`` and, with ``m_estimate_shape``, the BatchNorm-free ``to_shape`` MLP.
    """
    in_size = n_markers * ((3 if use_marker_pos else 0) + (9 if use_marker_ori else 0))
    dirs = 2 if bidirectional else 1
    h = hidden_size
    spec = []
    for layer in range(num_layers):
        n_in = in_size if layer == 0 else h * dirs
        for sfx in ([''] + (['_reverse'] if bidirectional else [])):
            spec.append(('rnn.lstm.weight_ih_l%d%s' % (layer, sfx), (4 * h, n_in), 'w_lstm'))
            spec.append(('rnn.lstm.weight_hh_l%d%s' % (layer, sfx), (4 * h, h), 'w_lstm'))
            spec.append(('rnn.lstm.bias_ih_l%d%s' % (layer, sfx), (4 * h,), 'b_lstm'))
            spec.append(('rnn.lstm.bias_hh_l%d%s' % (layer, sfx), (4 * h,), 'b_lstm'))
    spec.append(('to_pose.weight', (66, h * dirs), 'w'))
    spec.append(('to_pose.bias', (66,), 'b'))
    if estimate_shape:
        sh = shape_hidden_size
        spec.append(('to_shape.input_to_hidden.weight', (sh, h * dirs), 'w'))
        spec.append(('to_shape.input_to_hidden.bias', (sh,), 'b'))
        spec.append(('to_shape.activation_fn.weight', (1,), 'prelu'))
        spec.append(('to_shape.hidden_to_output.weight', (10, sh), 'w'))
        spec.append(('to_shape.hidden_to_output.bias', (10,), 'b'))
        for b in range(2):
            for l in range(2):
                base = 'to_shape.hidden_layers.%d.layers' % b
                spec.append(('%s.%d.weight' % (base, l * 3), (sh, sh), 'w'))
                spec.append(('%s.%d.bias' % (base, l * 3), (sh,), 'b'))
                spec.append(('%s.%d.weight' % (base, l * 3 + 1), (1,), 'prelu'))
    return spec


def synth_rnn_state_dict(seed=0, **model_kwargs):
    """Deterministic random weights for ``rnn_state_dict_spec`` (same streams / scales as ``synth_state_dict``)."""
    return _synth_from_spec(rnn_state_dict_spec(**model_kwargs), seed, model_kwargs.get('hidden_size', 1024))


def synth_state_dict(seed=0, **model_kwargs):
    """
    Deterministic random weights (numpy RandomState, one stream per key) shaped like torch's default
    initialisation, with BatchNorm running statistics perturbed so that folding them is exercised.
    Returns {key: numpy array}; float32 except ``num_batches_tracked`` (int64).
    """
    return _synth_from_spec(lgd_state_dict_spec(**model_kwargs), seed, model_kwargs.get('rnn_hidden_size', 512))


def _synth_from_spec(spec, seed, rnn_h):
    import zlib
    out = {}
    for key, shape, kind in spec:
        rng = np.random.RandomState((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31 - 1))
        if kind == 'w':
            bound = 1.0 / np.sqrt(shape[1])
            val = rng.uniform(-bound, bound, size=shape)
        elif kind == 'b':
            val = rng.uniform(-0.05, 0.05, size=shape)
        elif kind in ('w_lstm', 'b_lstm'):
            bound = 1.0 / np.sqrt(rnn_h)
            val = rng.uniform(-bound, bound, size=shape)
        elif kind == 'bn_w':
            val = rng.uniform(0.25, 1.0, size=shape)
        elif kind == 'bn_b':
            val = rng.normal(0.0, 0.05, size=shape)
        elif kind == 'bn_mean':
            val = rng.normal(0.0, 0.1, size=shape)
        elif kind == 'bn_var':
            val = rng.uniform(0.5, 1.5, size=shape)
        elif kind == 'prelu':
            val = rng.uniform(0.1, 0.4, size=shape)
        elif kind == 'bn_count':
            out[key] = np.asarray(100, dtype=np.int64)
            continue
        else:
            raise ValueError(kind)
        out[key] = val.astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Sensor windows
# ----------------------------------------------------------------------------------------------------------------------

def _random_rotations(rng, n):
    """n small random rotation matrices (axis-angle with sigma 0.15 rad), float64."""
    rv = rng.normal(0.0, 0.15, size=(n, 3))
    ang = np.linalg.norm(rv, axis=1, keepdims=True) + 1e-12
    k = rv / ang
    kx = np.zeros((n, 3, 3))
    kx[:, 0, 1], kx[:, 0, 2] = -k[:, 2], k[:, 1]
    kx[:, 1, 0], kx[:, 1, 2] = k[:, 2], -k[:, 0]
    kx[:, 2, 0], kx[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(ang)[:, :, None], np.cos(ang)[:, :, None]
    return np.eye(3)[None] + s * kx + (1 - c) * (kx @ kx)


def synth_window_params(n_windows, n_frames, seed=0, ragged=False, offsets=False, drop_rate=0.0):
    """
    The pose-side half of a synthetic batch: ``poses ~ 0.2 N(0,1)`` (B,F,66),
    ``shapes ~ N(0,1)`` (B,10), sensor-to-skin offsets (identity / zero, or small random ones),
    ``seq_lengths`` (all F, or ragged in [1, F]) and an optional sensor-dropout mask.
    Returns a dict of float32 / int64 numpy arrays.
    """
    rng = np.random.RandomState(1000003 * seed + 17)
    poses = 0.2 * rng.standard_normal((n_windows, n_frames, 66))
    shapes = rng.standard_normal((n_windows, 10))
    if offsets:
        offset_t = 0.02 * rng.standard_normal((n_windows, 12, 3))
        offset_r = _random_rotations(rng, n_windows * 12).reshape(n_windows, 12, 3, 3)
    else:
        offset_t = np.zeros((n_windows, 12, 3))
        offset_r = np.tile(np.eye(3)[None, None], (n_windows, 12, 1, 1))
    if ragged:
        seq_lengths = rng.randint(1, n_frames + 1, size=(n_windows,))
        seq_lengths[0] = n_frames
    else:
        seq_lengths = np.full((n_windows,), n_frames)
    masks = None
    if drop_rate > 0.0:
        masks = (rng.uniform(size=(n_windows, n_frames, 12)) >= drop_rate).astype(np.float32)
    return {'poses': poses.astype(np.float32), 'shapes': shapes.astype(np.float32),
            'offset_t': offset_t.astype(np.float32), 'offset_r': offset_r.astype(np.float32),
            'seq_lengths': seq_lengths.astype(np.int64), 'marker_masks': masks}


def write_synthetic_offsets(out_dir, n_files=3, seed=0):
    """Stand-ins for the reference's estimated sensor-to-skin offset files (``transforms.py:145-160``: ``means`` (12,3),
    ``covs`` (12,3,3), ``r`` (12,3,3), ``vertex_ids``).  Returns the list of paths."""
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    vertex_ids = [3027, 3748, 5430, 5178, 5006, 4447, 4559, 1961, 1391, 1535, 959, 1072]      # configuration.py:32-34
    for i in range(n_files):
        rng = np.random.RandomState(7000 + 31 * seed + i)
        means = 0.02 * rng.standard_normal((12, 3))
        a = 0.004 * rng.standard_normal((12, 3, 3))
        covs = a @ np.transpose(a, (0, 2, 1)) + 1e-6 * np.eye(3)[None]
        r = _random_rotations(rng, 12)
        path = os.path.join(out_dir, 'offsets_%d_%d.npz' % (seed, i))
        np.savez(path, means=means, covs=covs, r=r, vertex_ids=np.asarray(vertex_ids))
        paths.append(path)
    return paths


def synth_measurements(gt_pos, gt_ori, seed=0, pos_noise=0.01):
    """Measured sensors = projected ground truth + 1 cm position noise; (B,F,12,3)/(B,F,12,3,3) -> (B,F,36)/(B,F,108)."""
    rng = np.random.RandomState(1000003 * seed + 71)
    b, f = gt_pos.shape[0], gt_pos.shape[1]
    pos = np.asarray(gt_pos, dtype=np.float32).reshape(b, f, 36)
    pos = pos + (pos_noise * rng.standard_normal(pos.shape)).astype(np.float32)
    ori = np.asarray(gt_ori, dtype=np.float32).reshape(b, f, 108)
    return pos, ori
