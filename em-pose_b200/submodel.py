"""Host-side preparation of the SMPL-H constants the CUDA path keeps resident.

The reference evaluates the full 6890-vertex mesh N+1 times per window
(``empose/nn/models.py:474`` -> ``empose/bodymodels/smpl.py:121``) although the loop only
consumes 12 sensor frames and 22 joints.  This module extracts, once per model and in float64,
the minimal equivalent ("sub-model"):

* ``J(beta) = J0 + Jdirs . beta`` -- the joint regressor folded through the shape blend shapes;
* hand joints folded into the wrists (the reference always feeds a zero hand pose, ``smpl.py:99``,
  so every hand joint's skinning transform equals its wrist ancestor's and only the first
  21*9 pose-blend features are ever non-zero);
* only the vertices in the 1-ring of the 12 sensor vertices (``virtual_sensors.py:61-75``), stored ring by ring
  ("fan form", see ``extract_submodel`` and ``csrc/fan_math.h``).

The mesh connectivity is derived exactly the way the reference derives it (``topology_from_faces``),
and is plain data for the C-ABI so that kernel and oracle can never disagree about it.
"""
import numpy as np

#: sensor vertex ids in network order (reference ``empose/helpers/configuration.py:32-34``)
VERTEX_IDS = (3027, 3748, 5430, 5178, 5006, 4447, 4559, 1961, 1391, 1535, 959, 1072)
N_BODY_JOINTS = 22
N_POSE_FEATURES = 21 * 9
N_BETAS = 10


def vertex_faces_table(faces, n_vertices):
    """
    Incident faces of every vertex, ascending, padded with -1 -- the ``trimesh.Trimesh.vertex_faces``
    contract the reference relies on (``smpl.py:58-67``).  If trimesh is installed it is used, so the
    helper-vertex choice (first face listed, ``virtual_sensors.py:55-58``) is whatever the reference
    would get on this machine.
    """
    faces = np.asarray(faces, dtype=np.int64)
    try:
        import trimesh  # noqa: F401  (optional)
        if hasattr(trimesh, 'Trimesh') and getattr(trimesh, '__file__', None):
            mesh = trimesh.Trimesh(np.zeros((n_vertices, 3)), faces, process=False)
            return np.asarray(mesh.vertex_faces).astype(np.int64)
    except ImportError:
        pass
    flat_v = faces.reshape(-1)
    flat_f = np.repeat(np.arange(faces.shape[0]), 3)
    order = np.lexsort((flat_f, flat_v))
    flat_v, flat_f = flat_v[order], flat_f[order]
    degree = np.bincount(flat_v, minlength=n_vertices)
    table = np.full((n_vertices, int(degree.max())), -1, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(degree)[:-1]])
    slot = np.arange(flat_v.shape[0]) - start[flat_v]
    table[flat_v, slot] = flat_f
    return table


def topology_from_faces(faces, vertex_ids=VERTEX_IDS):
    """
    Sub-mesh faces, per-sensor incident faces and helper vertices, as the reference computes them
    (``virtual_sensors.py:47-75``).  All ids are GLOBAL vertex ids / rows into ``sub_faces``.
    """
    faces = np.asarray(faces, dtype=np.int64)
    ids = [int(v) for v in vertex_ids]
    n_vertices = int(faces.max()) + 1
    full_vf = vertex_faces_table(faces, n_vertices)
    touched = full_vf[ids]
    sub_faces = faces[np.unique(touched[touched != -1])]
    sensor_faces = vertex_faces_table(sub_faces, int(sub_faces.max()) + 1)[ids]
    helper_ids = []
    for v in ids:
        first_face = faces[full_vf[v, 0]]
        helper_ids.append(int(first_face[first_face != v][0]))
    return {'sub_faces': sub_faces, 'sensor_faces': sensor_faces,
            'helper_ids': np.asarray(helper_ids, dtype=np.int64), 'vertex_ids': np.asarray(ids, dtype=np.int64)}


def _first_body_ancestor(parents, j):
    while j >= N_BODY_JOINTS:
        j = int(parents[j])
    return j


MAX_FAN_JOINTS = 8           # csrc/fan_math.h kMaxFanJoints


def sensor_ring(faces_of_sensor, v):
    """
    Ring of sensor vertex ``v`` from its incident faces (``(deg, 3)`` global ids, in the reference's order).

    Returns ``(ring, is_fan)``: ``ring[0] == v``; if the faces form a closed, consistently wound fan, ``ring[1:]`` are the
    neighbours in winding order, starting with the first face, so that face ``d`` is -- up to a cyclic rotation of its
    corners, which leaves the normal ``(v1-v0)x(v2-v0)`` of ``utils.py:135`` unchanged -- ``(v, ring[1+d], ring[1+(d+1)%deg])``.
    Otherwise (boundary, non-manifold or inconsistently wound neighbourhood) the other vertices follow in order of appearance.
    """
    faces_of_sensor = [[int(a) for a in f] for f in faces_of_sensor]
    deg = len(faces_of_sensor)
    nxt, first = {}, None
    ok = deg >= 3
    for f in faces_of_sensor:
        k = f.index(v)
        p, q = f[(k + 1) % 3], f[(k + 2) % 3]
        if first is None:
            first = p
        if p in nxt or p == v or q == v:
            ok = False
        nxt[p] = q
    if ok:
        order, cur = [first], nxt[first]
        while cur != first and cur in nxt and len(order) <= deg:
            order.append(cur)
            cur = nxt[cur]
        ok = cur == first and len(order) == deg
        if ok:
            return [v] + order, True
    ring = [v]
    for f in faces_of_sensor:
        for a in f:
            if a not in ring:
                ring.append(a)
    return ring, False


def extract_submodel(v_template, shapedirs, posedirs, j_regressor, weights, kintree_table, topology):
    """
    :param v_template: (V,3) or (1,V,3).  shapedirs: (V,3,>=10).  posedirs: (459, V*3) (the BodyModel buffer
        layout found in released checkpoints under ``smpl.bm.posedirs``) or (V,3,459) (the npz layout).
    :param j_regressor: (52,V).  weights: (V,52).  kintree_table: (2,52).  topology: ``topology_from_faces``.
    :return: dict of float32 / int32 arrays keyed ``sub.*`` (see include/empose_b200.h) plus python ints.

    The sub-mesh is stored RING-MAJOR: sensor ``s`` owns the ``slots`` (8 or 12) consecutive sub-model vertices
    ``s*slots ..``, slot 0 being the sensor vertex, then its neighbours (``sensor_ring``), then zero padding.  A vertex
    shared by two rings is simply stored twice (the blend GEMMs are linear, so its gradient columns add up).
    """
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    v_template = f64(v_template).reshape(-1, 3)
    n_v = v_template.shape[0]
    shapedirs = f64(shapedirs)[:, :, :N_BETAS]
    posedirs = f64(posedirs)
    if posedirs.ndim == 3:                                   # npz layout (V,3,459) -> (459, V*3)
        posedirs = posedirs.reshape(n_v * 3, -1).T
    assert posedirs.shape[1] == n_v * 3
    j_regressor, weights = f64(j_regressor), f64(weights)
    parents = [int(p) for p in np.asarray(kintree_table)[0]]
    parents[0] = -1
    for j in range(1, N_BODY_JOINTS):
        assert 0 <= parents[j] < j, 'body joints must be topologically ordered'

    sub_faces = np.asarray(topology['sub_faces'], dtype=np.int64)
    sensor_ids = [int(v) for v in topology['vertex_ids']]
    helper_ids = [int(v) for v in topology['helper_ids']]
    sensor_faces_g = np.asarray(topology['sensor_faces'], dtype=np.int64)          # rows of sub_faces, -1 padded
    n_sensors = len(sensor_ids)
    degree = (sensor_faces_g > -1).sum(axis=1)
    max_deg = int(sensor_faces_g.shape[1])

    rings, fans = [], []
    for s_idx, v in enumerate(sensor_ids):
        ring, is_fan = sensor_ring(sub_faces[sensor_faces_g[s_idx, :degree[s_idx]]], v)
        assert helper_ids[s_idx] in ring
        rings.append(ring)
        fans.append(is_fan)
    max_ring = max(len(r) for r in rings)
    if max_ring > 12:
        raise ValueError('sensor rings of more than 12 vertices are not supported (found %d)' % max_ring)
    slots = 8 if max_ring <= 8 else 12
    n_sub = n_sensors * slots
    gid = -np.ones(n_sub, dtype=np.int64)                                           # global id of every sub-model vertex
    for s_idx, ring in enumerate(rings):
        gid[s_idx * slots:s_idx * slots + len(ring)] = ring
    real = gid >= 0
    local_of = [{g: s_idx * slots + r for r, g in enumerate(ring)} for s_idx, ring in enumerate(rings)]

    # per-sensor face lists in the reference's order, corners as ring-local sub-model vertices
    faces_local, sensor_faces = [], -np.ones((n_sensors, max_deg), dtype=np.int64)
    for s_idx in range(n_sensors):
        for d in range(int(degree[s_idx])):
            sensor_faces[s_idx, d] = len(faces_local)
            faces_local.append([local_of[s_idx][int(g)] for g in sub_faces[sensor_faces_g[s_idx, d]]])
    faces_local = np.asarray(faces_local, dtype=np.int64)
    sensor_vert = np.asarray([s_idx * slots for s_idx in range(n_sensors)], dtype=np.int64)
    helper_vert = np.asarray([local_of[s_idx][helper_ids[s_idx]] for s_idx in range(n_sensors)], dtype=np.int64)

    # joints folded through the shape blend shapes
    j0 = j_regressor[:N_BODY_JOINTS] @ v_template                                   # (22,3)
    jdirs = np.einsum('jv,vck->kjc', j_regressor[:N_BODY_JOINTS], shapedirs)        # (10,22,3)

    # hands folded into their first body ancestor
    w22 = np.zeros((n_sub, N_BODY_JOINTS))
    for j in range(weights.shape[1]):
        w22[real, _first_body_ancestor(parents, j)] += weights[gid[real], j]
    nnz = (w22 != 0.0)
    n_skin = int(nnz.sum(axis=1).max())
    skin_joint = np.zeros((n_sub, n_skin), dtype=np.int32)
    skin_weight = np.zeros((n_sub, n_skin), dtype=np.float64)
    for v in range(n_sub):
        js = np.nonzero(nnz[v])[0]
        skin_joint[v, :js.shape[0]] = js
        skin_weight[v, :js.shape[0]] = w22[v, js]
    # the same relation grouped by joint (gather lists for the reverse pass of the general kernel)
    jt_ptr = np.zeros(N_BODY_JOINTS + 1, dtype=np.int32)
    jt_vert, jt_weight = [], []
    for j in range(N_BODY_JOINTS):
        vs = np.nonzero(nnz[:, j])[0]
        jt_vert += vs.tolist()
        jt_weight += w22[vs, j].tolist()
        jt_ptr[j + 1] = len(jt_vert)

    # the same lists cut into chunks of bounded length ("virtual joints") so GPU lanes stay balanced; at most
    # MAX_CHUNKS chunks because their partial sums share storage with other per-joint scratch (frame_math.h kMaxVj)
    MAX_CHUNKS = 44
    chunk_len = 8
    while True:
        vj_ptr, jvj_ptr = [0], [0]
        for j in range(N_BODY_JOINTS):
            lo, hi = int(jt_ptr[j]), int(jt_ptr[j + 1])
            for start in range(lo, hi, chunk_len):
                vj_ptr.append(min(start + chunk_len, hi))
            jvj_ptr.append(len(vj_ptr) - 1)
        if len(vj_ptr) - 1 <= MAX_CHUNKS:
            break
        chunk_len += 2

    # per-vertex incidence lists for the atomic-free reverse of the sensor phase (general kernel): entry (item, code) with
    # item = sensor * max_degree + slot and code 0/1/2 = corner of that face, 3 = the sensor vertex itself,
    # 4 = the sensor's helper vertex
    inc = [[] for _ in range(n_sub)]
    for s_idx in range(n_sensors):
        for d in range(max_deg):
            fid = int(sensor_faces[s_idx, d])
            if fid < 0:
                continue
            for corner in range(3):
                inc[int(faces_local[fid, corner])].append((s_idx * max_deg + d, corner))
        inc[int(sensor_vert[s_idx])].append((s_idx * max_deg, 3))
        inc[int(helper_vert[s_idx])].append((s_idx * max_deg, 4))
    vinc_ptr = np.zeros(n_sub + 1, dtype=np.int32)
    vinc_item, vinc_code = [], []
    for v in range(n_sub):
        for item, code in inc[v]:
            vinc_item.append(item)
            vinc_code.append(code)
        vinc_ptr[v + 1] = len(vinc_item)

    # fan tables (csrc/fan_math.h FanModel): per sensor the union of its ring's skinning joints with dense weights,
    # and per joint the list of (sensor, joint) partial sums that make up dE/dA_j
    fan_ok = all(fans)
    fan_nj = np.zeros(n_sensors, dtype=np.int32)
    fan_joint = np.zeros((n_sensors, MAX_FAN_JOINTS), dtype=np.int32)
    fan_weight = np.zeros((n_sensors, MAX_FAN_JOINTS, slots), dtype=np.float64)
    fan_helper = np.zeros(n_sensors, dtype=np.int32)
    for s_idx in range(n_sensors):
        blk = slice(s_idx * slots, (s_idx + 1) * slots)
        js = np.nonzero(nnz[blk].any(axis=0))[0]
        if js.shape[0] > MAX_FAN_JOINTS:
            fan_ok = False
            js = js[:MAX_FAN_JOINTS]
        fan_nj[s_idx] = js.shape[0]
        fan_joint[s_idx, :js.shape[0]] = js
        fan_weight[s_idx, :js.shape[0]] = w22[blk][:, js].T
        fan_helper[s_idx] = rings[s_idx].index(helper_ids[s_idx]) - 1
    fan_part_ptr = np.concatenate([[0], np.cumsum(fan_nj)]).astype(np.int32)
    jp = [[] for _ in range(N_BODY_JOINTS)]
    for s_idx in range(n_sensors):
        for u in range(int(fan_nj[s_idx])):
            jp[int(fan_joint[s_idx, u])].append(int(fan_part_ptr[s_idx]) + u)
    fan_jp_ptr = np.concatenate([[0], np.cumsum([len(x) for x in jp])]).astype(np.int32)
    fan_jp_idx = np.asarray([i for x in jp for i in x] or [0], dtype=np.int32)

    vp_dim = ((n_sub * 3 + 15) // 16) * 16                     # padded width of the per-frame vertex vector
    safe = np.where(real, gid, 0)
    mask3 = np.repeat(real, 3)
    pd = posedirs.reshape(posedirs.shape[0], n_v, 3)[:N_POSE_FEATURES, safe].reshape(N_POSE_FEATURES, n_sub * 3) * mask3
    sd = shapedirs[safe].reshape(n_sub * 3, N_BETAS).T * mask3                       # (10, Vs*3)
    vt = v_template[safe].reshape(-1) * mask3
    pad = lambda a: np.concatenate([a, np.zeros(a.shape[:-1] + (vp_dim - n_sub * 3,))], axis=-1)

    out = {
        'sub.v_template': pad(vt),                                                    # (VP,)
        'sub.shapedirs': pad(sd),                                                     # (10, VP)
        'sub.posedirs': pad(pd),                                                      # (189, VP)
        'sub.j0': j0.reshape(-1),                                                     # (66,)
        'sub.jdirs': jdirs.reshape(N_BETAS, N_BODY_JOINTS * 3),                       # (10, 66)
        'sub.skin_weight': skin_weight,                                               # (Vs, n_skin)
        'sub.jt_weight': np.asarray(jt_weight, dtype=np.float64),
        'sub.fan_weight': fan_weight,                                                 # (12, 8, slots)
    }
    out = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}
    out.update({
        'sub.parents': np.asarray(parents[:N_BODY_JOINTS], dtype=np.int32),
        'sub.skin_joint': skin_joint,
        'sub.jt_ptr': jt_ptr,
        'sub.jt_vert': np.asarray(jt_vert, dtype=np.int32),
        'sub.vinc_ptr': vinc_ptr,
        'sub.vinc_item': np.asarray(vinc_item, dtype=np.int32),
        'sub.vinc_code': np.asarray(vinc_code, dtype=np.int32),
        'sub.vj_ptr': np.asarray(vj_ptr, dtype=np.int32),
        'sub.jvj_ptr': np.asarray(jvj_ptr, dtype=np.int32),
        'sub.faces': faces_local.astype(np.int32),                                    # (sum of degrees, 3) sub-model vertex ids
        'sub.sensor_vert': sensor_vert.astype(np.int32),                              # (12,)
        'sub.helper_vert': helper_vert.astype(np.int32),                              # (12,)
        'sub.sensor_faces': sensor_faces.astype(np.int32),                            # (12,deg) rows into faces, -1 pad
        'sub.sensor_degree': degree.astype(np.int32),
        'sub.global_vertex_ids': gid.astype(np.int32),                                # -1: padding slot
        'sub.fan_dims': np.asarray([int(fan_ok), slots, int(degree.max()), int(fan_part_ptr[-1])], dtype=np.int32),
        'sub.fan_helper': fan_helper,                                                 # (12,) fan index of the helper vertex
        'sub.fan_n_joints': fan_nj,                                                   # (12,)
        'sub.fan_part_ptr': fan_part_ptr,                                             # (13,)
        'sub.fan_joint': fan_joint,                                                   # (12, 8)
        'sub.fan_jp_ptr': fan_jp_ptr,                                                 # (23,)
        'sub.fan_jp_idx': fan_jp_idx,
    })
    out['dims'] = {'n_verts': n_sub, 'vp_dim': int(vp_dim), 'n_faces': int(faces_local.shape[0]),
                   'max_degree': max_deg, 'n_skin': n_skin, 'n_sensors': n_sensors, 'slots': slots, 'fan_ok': bool(fan_ok)}
    return out


def submodel_from_npz(npz_path, vertex_ids=VERTEX_IDS):
    """Convenience: SMPL-H ``model.npz`` (the file of ``smpl.py:26``) -> sub-model arrays."""
    with np.load(npz_path) as z:
        topo = topology_from_faces(z['f'], vertex_ids)
        return extract_submodel(z['v_template'], z['shapedirs'], z['posedirs'], z['J_regressor'], z['weights'],
                                z['kintree_table'], topo), topo


def extract_fullmodel(v_template, shapedirs, posedirs, j_regressor, weights, kintree_table):
    """
    Constants of the FULL 6890-vertex mesh for ``SMPLLayer.forward`` (reference ``smpl.py:81-122``): the same folds as
    the sub-model (joints through the shape blend shapes, zero-pose hands into their body ancestor) but for every
    vertex and all 52 joints.  Returns float32 / int32 arrays keyed ``smpl.*``.
    """
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    v_template = f64(v_template).reshape(-1, 3)
    n_v = v_template.shape[0]
    shapedirs = f64(shapedirs)[:, :, :N_BETAS]
    posedirs = f64(posedirs)
    if posedirs.ndim == 3:
        posedirs = posedirs.reshape(n_v * 3, -1).T
    j_regressor, weights = f64(j_regressor), f64(weights)
    parents = [int(p) for p in np.asarray(kintree_table)[0]]
    parents[0] = -1
    n_j = j_regressor.shape[0]
    j0 = j_regressor @ v_template                                           # (52,3)
    jdirs = np.einsum('jv,vck->kjc', j_regressor, shapedirs)                # (10,52,3)
    ancestor = np.asarray([_first_body_ancestor(parents, j) for j in range(n_j)], dtype=np.int32)
    w22 = np.zeros((n_v, N_BODY_JOINTS))
    for j in range(n_j):
        w22[:, ancestor[j]] += weights[:, j]
    nnz = w22 != 0.0
    n_skin = int(nnz.sum(axis=1).max())
    order = np.argsort(~nnz, axis=1, kind='stable')[:, :n_skin]             # non-zero joints first, ascending
    skin_joint = order.astype(np.int32)
    skin_weight = np.take_along_axis(w22, order, axis=1)
    skin_joint[skin_weight == 0.0] = 0
    return {
        'smpl.v_template': v_template.reshape(-1).astype(np.float32),                                   # (V*3,)
        'smpl.shapedirs': np.ascontiguousarray(shapedirs.reshape(n_v * 3, N_BETAS).T, dtype=np.float32),  # (10, V*3)
        'smpl.posedirs': np.ascontiguousarray(posedirs[:N_POSE_FEATURES], dtype=np.float32),            # (189, V*3)
        'smpl.j0': j0.reshape(-1).astype(np.float32),                                                   # (156,)
        'smpl.jdirs': jdirs.reshape(N_BETAS, n_j * 3).astype(np.float32),                               # (10, 156)
        'smpl.parents': np.asarray(parents[:N_BODY_JOINTS], dtype=np.int32),
        'smpl.ancestor': ancestor,                                                                      # (52,)
        'smpl.skin_joint': skin_joint,
        'smpl.skin_weight': skin_weight.astype(np.float32),
        'smpl.dims': np.asarray([n_v, n_j, n_skin], dtype=np.int32),
    }
