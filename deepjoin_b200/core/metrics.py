"""Drop-in for the two nearest-neighbour metrics of the reference's ``core/metrics.py`` that follow
meshing in ``infer.sh``: ``chamfer`` (:16-53) and ``normal_consistency`` (:56-95).
The KD-tree queries run as a brute-force nearest-neighbour kernel on the GPU (dj_nn_query); surface
sampling follows ``trimesh.sample.sample_surface`` (area-weighted faces, sqrt-barycentric points) on the
host with numpy's global RNG -- the random stream of trimesh itself is not reproduced (trimesh is not
available offline), so values agree with the reference statistically, not sample for sample."""
import numpy as np
import torch

from deepjoin_b200 import _lib
from . import errors


def nn_query(queries, targets, device=None):
    """cKDTree(targets).query(queries) -> (distances [N] float64, indices [N] int64)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    q = torch.as_tensor(np.ascontiguousarray(queries, np.float32)).to(dev) if not isinstance(queries, torch.Tensor) else queries.to(dev, torch.float32).contiguous()
    t = torch.as_tensor(np.ascontiguousarray(targets, np.float32)).to(dev) if not isinstance(targets, torch.Tensor) else targets.to(dev, torch.float32).contiguous()
    d2 = torch.empty(q.shape[0], device=dev, dtype=torch.float32)
    idx = torch.empty(q.shape[0], device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dj_nn_query(q.data_ptr(), q.shape[0], t.data_ptr(), t.shape[0], d2.data_ptr(), idx.data_ptr(),
                                          _lib.stream_ptr()))
    return torch.sqrt(d2.double()).cpu().numpy(), idx.cpu().numpy().astype(np.int64)


def face_normals(mesh):
    v, f = np.asarray(mesh.vertices, np.float64), np.asarray(mesh.faces)
    return np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])


def sample_surface(mesh, count):
    """Area-weighted surface samples (points [count,3], face index [count]) as trimesh.sample.sample_surface."""
    v, f = np.asarray(mesh.vertices, np.float64), np.asarray(mesh.faces)
    area = 0.5 * np.linalg.norm(face_normals(mesh), axis=1)
    cum = np.cumsum(area)
    face = np.searchsorted(cum, np.random.random(count) * cum[-1])
    face = np.minimum(face, len(f) - 1)
    o, e1, e2 = v[f[face, 0]], v[f[face, 1]] - v[f[face, 0]], v[f[face, 2]] - v[f[face, 0]]
    r = np.random.random((count, 2))
    flip = r.sum(axis=1) > 1.0
    r[flip] = 1.0 - r[flip]
    return o + e1 * r[:, :1] + e2 * r[:, 1:], face


def chamfer(gt_shape, pred_shape, num_mesh_samples=2000):
    """core/metrics.py:16-53: symmetric mean squared nearest-neighbour distance."""
    if np.asarray(pred_shape.vertices).shape[0] == 0:
        raise errors.MeshEmptyError
    if hasattr(gt_shape, "faces"):
        assert np.asarray(gt_shape.vertices).shape[0] != 0, "gt shape has no vertices"
        gt_pts = sample_surface(gt_shape, num_mesh_samples)[0]
    else:
        gt_pts = np.asarray(gt_shape)
        assert gt_pts.shape[0] == num_mesh_samples, \
            "Wrong number of gt points, expected {} got {}".format(num_mesh_samples, gt_pts.shape[0])
    pred_pts = sample_surface(pred_shape, num_mesh_samples)[0]
    one, _ = nn_query(gt_pts, pred_pts)
    two, _ = nn_query(pred_pts, gt_pts)
    return float(np.mean(np.square(one)) + np.mean(np.square(two)))


def normal_consistency(gt_shape, pred_shape, num_mesh_samples=30000):
    """core/metrics.py:56-95: mean dot product of the face normals at nearest surface samples, both ways."""
    if np.asarray(pred_shape.vertices).shape[0] == 0:
        raise errors.MeshEmptyError
    assert np.asarray(gt_shape.vertices).shape[0] != 0, "gt shape has no vertices"

    def normal_diff(obj_from, obj_to):
        n_from_all, n_to_all = face_normals(obj_from), face_normals(obj_to)
        verts_from, fi_from = sample_surface(obj_from, num_mesh_samples)
        verts_to, fi_to = sample_surface(obj_to, num_mesh_samples)
        _, idx = nn_query(verts_from, verts_to)
        n_from = n_from_all[fi_from]
        n_to = n_to_all[fi_to[idx]]
        n_from = n_from / np.linalg.norm(n_from, axis=-1, keepdims=True)
        n_to = n_to / np.linalg.norm(n_to, axis=-1, keepdims=True)
        return (n_to * n_from).sum(axis=-1)

    return float(np.hstack((normal_diff(gt_shape, pred_shape), normal_diff(pred_shape, gt_shape))).mean())
