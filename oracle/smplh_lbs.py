"""SMPL-H linear blend skinning, restated from the published formulation (test infrastructure).

PARITY UNPINNED for this file: the arithmetic the reference runs at
``empose/bodymodels/smpl.py:121`` (``self.bm(root_orient=..., pose_body=..., betas=...,
pose_hand=..., trans=...)`` -> ``.v``, ``.Jtr``) lives in the third-party package
``human-body-prior`` (fork totomobile43/human_body_prior @ 821a0e7ebf93f33702babe21c53ecce7d1aa9582,
reference ``requirements.txt:9``), which is not in ``/root/reference`` and not installable here.
The reference has no tests or golden vectors at this boundary.  What follows restates the
SMPL / SMPL-H model as published (Loper et al. 2015; Romero et al. 2017) in the form the
smplx-style ``lbs`` routine of that package implements it:

  1. shape blend      v_s = v_template + shapedirs . betas
  2. joint regression J   = J_regressor . v_s                                    (52 x 3)
  3. Rodrigues        R_j = I + sin(a) K + (1 - cos(a)) K^2,  a = ||r_j + 1e-8||, K = hat(r_j / a)
  4. pose blend       v_p = v_s + posedirs^T . vec(R_1..R_51 - I)               (459 features)
  5. kinematic chain  G_j = G_parent(j) . [R_j | J_j - J_parent(j)],  Jtr_j = G_j[:3, 3],
                      A_j = G_j - [0 | G_j . J_j]
  6. skinning         v   = (sum_j w_vj A_j) . [v_p; 1] + trans

Known-answer / self-consistency checks are in ``tests/test_oracle_smpl.py``.

All functions are differentiable torch code and run in whatever dtype the inputs have.
"""
import numpy as np
import torch

N_BODY_JOINTS = 22       # root + 21 (reference configuration.py:104)
N_HAND_JOINTS = 15       # per hand (reference configuration.py:106)
N_ALL_JOINTS = 52


class SmplhModel(object):
    """Plain tensor container for one SMPL-H model (no nn.Module on purpose)."""

    def __init__(self, source, num_betas=10, dtype=torch.float64):
        data = np.load(source) if isinstance(source, str) else source
        as_t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dtype)
        self.dtype = dtype
        self.num_betas = num_betas
        self.v_template = as_t(data['v_template'])                               # (V,3)
        self.faces = torch.as_tensor(np.asarray(data['f']).astype(np.int64))      # (F,3)
        self.shapedirs = as_t(np.asarray(data['shapedirs'])[:, :, :num_betas])    # (V,3,nb)
        pd = np.asarray(data['posedirs'], dtype=np.float64)                       # (V,3,459)
        self.posedirs = as_t(pd.reshape(pd.shape[0] * 3, pd.shape[2]).T.copy())   # (459, V*3)
        self.j_regressor = as_t(data['J_regressor'])                              # (52,V)
        kt = np.asarray(data['kintree_table']).astype(np.int64)
        parents = kt[0].tolist()
        parents[0] = -1
        self.parents = parents
        self.kintree_table = torch.as_tensor(kt)
        self.weights = as_t(data['weights'])                                      # (V,52)

    @property
    def n_verts(self):
        return self.v_template.shape[0]

    def to(self, dtype):
        out = object.__new__(SmplhModel)
        out.__dict__.update(self.__dict__)
        out.dtype = dtype
        for name in ('v_template', 'shapedirs', 'posedirs', 'j_regressor', 'weights'):
            setattr(out, name, getattr(self, name).to(dtype))
        return out


def rodrigues(rotvecs):
    """Axis-angle (N,3) -> rotation matrices (N,3,3); the angle is ``||r + 1e-8||`` (step 3 above)."""
    angle = torch.linalg.vector_norm(rotvecs + 1e-8, dim=1, keepdim=True)        # (N,1)
    axis = rotvecs / angle
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    zero = torch.zeros_like(x)
    skew = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=rotvecs.dtype, device=rotvecs.device).unsqueeze(0)
    s = torch.sin(angle).unsqueeze(-1)
    c = torch.cos(angle).unsqueeze(-1)
    return eye + s * skew + (1.0 - c) * torch.bmm(skew, skew)


def kinematic_chain(rotmats, rest_joints, parents):
    """Step 5: returns posed joints (N,J,3) and the skinning transforms A as (rot (N,J,3,3), trans (N,J,3))."""
    n_joints = rest_joints.shape[1]
    world_rot = [rotmats[:, 0]]
    world_pos = [rest_joints[:, 0]]
    for j in range(1, n_joints):
        p = parents[j]
        world_rot.append(torch.bmm(world_rot[p], rotmats[:, j]))
        bone = (rest_joints[:, j] - rest_joints[:, p]).unsqueeze(-1)
        world_pos.append(torch.bmm(world_rot[p], bone).squeeze(-1) + world_pos[p])
    g_rot = torch.stack(world_rot, dim=1)
    g_pos = torch.stack(world_pos, dim=1)
    a_trans = g_pos - torch.matmul(g_rot, rest_joints.unsqueeze(-1)).squeeze(-1)
    return g_pos, g_rot, a_trans


def lbs(model, full_pose, betas, trans=None):
    """
    Evaluate SMPL-H.
    :param full_pose: (N, 156) axis-angle: root (3) | body (63) | hands (90).
    :param betas: (N, num_betas).
    :param trans: (N, 3) or None.
    :return: vertices (N,V,3), posed joints (N,52,3).
    """
    n = full_pose.shape[0]
    n_j = model.j_regressor.shape[0]
    v_shaped = model.v_template.unsqueeze(0) + torch.einsum('vck,nk->nvc', model.shapedirs, betas)
    rest_joints = torch.einsum('jv,nvc->njc', model.j_regressor, v_shaped)
    rotmats = rodrigues(full_pose.reshape(n * n_j, 3)).reshape(n, n_j, 3, 3)
    eye = torch.eye(3, dtype=full_pose.dtype, device=full_pose.device)
    pose_feature = (rotmats[:, 1:] - eye).reshape(n, (n_j - 1) * 9)
    v_posed = v_shaped + torch.matmul(pose_feature, model.posedirs).reshape(n, -1, 3)
    joints, a_rot, a_trans = kinematic_chain(rotmats, rest_joints, model.parents)
    blend_rot = torch.einsum('vj,njrc->nvrc', model.weights, a_rot)
    blend_trans = torch.einsum('vj,njr->nvr', model.weights, a_trans)
    verts = torch.einsum('nvrc,nvc->nvr', blend_rot, v_posed) + blend_trans
    if trans is not None:
        verts = verts + trans.unsqueeze(1)
        joints = joints + trans.unsqueeze(1)
    return verts, joints


def smpl_layer_forward(model, poses_body, betas, poses_root=None, trans=None):
    """
    The call the reference wrapper makes (``empose/bodymodels/smpl.py:81-122``): hands are zero
    (``:99``), root / trans default to zero (``:101-105``), betas are broadcast and cut to
    ``num_betas`` (``:108-110``).  Returns (vertices (N,6890,3), joints (N,52,3)).
    """
    n = poses_body.shape[0]
    dt, dev = poses_body.dtype, poses_body.device
    if poses_root is None:
        poses_root = torch.zeros(n, 3, dtype=dt, device=dev)
    if betas.dim() == 1 or betas.shape[0] == 1:
        betas = betas.reshape(1, -1).repeat(n, 1)
    betas = betas[:, :model.num_betas]
    hands = torch.zeros(n, 2 * N_HAND_JOINTS * 3, dtype=dt, device=dev)
    full_pose = torch.cat([poses_root, poses_body, hands], dim=1)
    return lbs(model, full_pose, betas, trans)
