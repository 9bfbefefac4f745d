"""Virtual EM-sensor frames from a posed mesh, restated (test infrastructure).

Follows ``empose/data/virtual_sensors.py`` and ``empose/helpers/utils.py:126-146`` of the
reference; pinned by ``tests/golden`` (outputs of the unmodified reference modules).
"""
import numpy as np
import torch

#: sensor vertex ids in network order (reference ``empose/helpers/configuration.py:32-34``)
VERTEX_IDS = [3027, 3748, 5430, 5178, 5006, 4447, 4559, 1961, 1391, 1535, 959, 1072]
#: 6-sensor subset (reference ``configuration.py:89``)
S_CONFIG_6 = [0, 1, 2, 6, 7, 11]


def vertex_faces_table(faces, n_vertices):
    """
    For every vertex the ids of its incident faces in ascending order, padded with -1 to the
    largest degree -- the contract of ``trimesh.Trimesh(...).vertex_faces`` that the reference
    consumes at ``smpl.py:58-67`` and ``virtual_sensors.py:52-58, 66-74`` (trimesh itself is absent).
    """
    faces = np.asarray(faces)
    incident = [[] for _ in range(n_vertices)]
    for face_id, tri in enumerate(faces):
        for v in tri:
            incident[int(v)].append(face_id)
    width = max(len(row) for row in incident)
    table = np.full((n_vertices, width), -1, dtype=np.int64)
    for v, row in enumerate(incident):
        table[v, :len(row)] = row
    return table


def sensor_topology(faces, vertex_ids=VERTEX_IDS):
    """
    Everything about the mesh connectivity the sensor projection needs.
    :return: dict with
       ``sub_faces`` (Fs,3) global vertex ids of faces touching any sensor vertex (``virtual_sensors.py:61-75``),
       ``sensor_faces`` (M,deg) rows into ``sub_faces`` incident to each sensor vertex, -1 padded,
       ``helper_ids`` (M,) first vertex != sensor vertex of the first full-mesh face of that vertex
       (``virtual_sensors.py:47-59``).
    """
    faces = np.asarray(faces).astype(np.int64)
    n_vertices = int(faces.max()) + 1
    full_vf = vertex_faces_table(faces, n_vertices)
    touched = full_vf[list(vertex_ids)]
    face_ids = np.unique(touched[touched != -1])
    sub_faces = faces[face_ids]
    sub_vf = vertex_faces_table(sub_faces, int(sub_faces.max()) + 1)[list(vertex_ids)]
    helpers = []
    for v in vertex_ids:
        for cand in faces[full_vf[v, 0]]:
            if cand != v:
                helpers.append(int(cand))
                break
    return {'sub_faces': sub_faces, 'sensor_faces': sub_vf, 'helper_ids': np.asarray(helpers, dtype=np.int64)}


def area_weighted_vertex_normals(vertices, faces, vertex_faces):
    """``utils.py:126-146``: un-normalised face normals, summed over incident faces, divided by the degree."""
    faces = torch.as_tensor(faces, dtype=torch.long, device=vertices.device)
    vertex_faces = torch.as_tensor(vertex_faces, dtype=torch.long, device=vertices.device)
    corners = vertices[:, faces]                                              # (N,F,3,3)
    face_n = torch.cross(corners[:, :, 1] - corners[:, :, 0], corners[:, :, 2] - corners[:, :, 0], dim=-1)
    valid = (vertex_faces > -1)
    gathered = face_n[:, vertex_faces.clamp(min=0)] * valid[None, :, :, None].to(vertices.dtype)
    degree = valid.sum(dim=-1).to(vertices.dtype)
    return gathered.sum(dim=-2) / degree[None, :, None]


def sensor_frames(vertices, topology, vertex_ids=VERTEX_IDS):
    """
    ``VirtualMarkerHelper.get_virtual_pos_and_rot`` (``virtual_sensors.py:85-96``).
    :param vertices: (N,V,3) posed mesh.
    :return: positions (N,M,3), orientations (N,M,3,3) with columns [on_surface, third, normal], raw normals (N,M,3).
    """
    ids = list(vertex_ids)
    raw_n = area_weighted_vertex_normals(vertices, topology['sub_faces'], topology['sensor_faces'])
    pos = vertices[:, ids]
    unit = lambda x: x / torch.linalg.vector_norm(x, dim=-1, keepdim=True)
    normal = unit(raw_n)
    tangent0 = unit(vertices[:, list(topology['helper_ids'])] - pos)
    # The reference calls torch.cross without ``dim`` (virtual_sensors.py:27,30): that resolves to the
    # first size-3 dimension, i.e. the last one here unless the batch itself has exactly 3 rows.
    third = unit(torch.cross(normal, tangent0, dim=-1))
    tangent = unit(torch.cross(third, normal, dim=-1))
    ori = torch.stack([tangent, third, normal], dim=-1)
    return pos, ori, raw_n


def apply_offsets(pos, ori, offset_r, offset_t):
    """``models.py:478-479``: R' = R R_off, p' = p + R t_off."""
    ori_c = torch.matmul(ori, offset_r)
    pos_c = pos + torch.matmul(ori, offset_t.unsqueeze(-1)).squeeze(-1)
    return pos_c, ori_c
