"""Skeleton constants the affective encoder is built from (reference utils/ted_db_utils.py:14-19).
The constants are on the hot path; the clip-level host helpers at the end serve the long-form entry points."""
import numpy as np

# (parent joint, child joint, bone length) of the 9 upper-body direction vectors
dir_vec_pairs = [(0, 1, 0.26), (1, 2, 0.18), (2, 3, 0.14), (1, 4, 0.22), (4, 5, 0.36),
                 (5, 6, 0.33), (1, 7, 0.22), (7, 8, 0.36), (8, 9, 0.33)]
# adjacency between direction vectors (graph 1: 9 nodes) and between body parts (graph 2: 3 nodes)
dir_edge_pairs = [(0, 1), (1, 2), (0, 3), (3, 4), (4, 5), (0, 6), (6, 7), (7, 8)]
body_parts_edge_idx = [np.arange(0, 3), np.arange(3, 6), np.arange(6, 9)]
max_body_part_edges = 3
body_parts_edge_pairs = [(0, 1), (0, 2)]


# ---- clip-level host helpers of the long-form entry points (utils/ted_db_utils.py:50-60, :81-124).  They touch one
# ---- clip's ground-truth skeleton once per rendered clip (metadata preparation); the per-frame conversion of
# ---- GENERATED motion runs on the device (csrc/longform.cu, ops.dir_vec_to_pose).
def resample_pose_seq(poses, duration_in_sec, fps):
    """Linear resampling of a pose sequence to `fps` (the reference uses scipy interp1d with extrapolation, :50-60)."""
    poses = np.asarray(poses)
    n = len(poses)
    x_new = np.arange(0, n, n / (duration_in_sec * fps))
    lo = np.clip(np.floor(x_new).astype(np.int64), 0, n - 2)
    frac = (x_new - lo).reshape((-1,) + (1,) * (poses.ndim - 1))
    p = poses.astype(np.float64)
    out = (p[lo + 1] - p[lo]) * frac + p[lo]
    return out.astype(poses.dtype)


def convert_pose_seq_to_dir_vec(pose):
    """joint positions [..., 10, 3] (or flattened last dim) -> unit direction vectors [..., 9, 3] (:105-124)"""
    pose = np.asarray(pose)
    if pose.shape[-1] != 3:
        pose = pose.reshape(pose.shape[:-1] + (-1, 3))
    parents = [a for a, _, _ in dir_vec_pairs]
    children = [b for _, b, _ in dir_vec_pairs]
    d = (pose[..., children, :] - pose[..., parents, :]).astype(np.float64)
    norm = np.sqrt((d * d).sum(axis=-1, keepdims=True))
    norm[norm == 0.0] = 1.0
    return d / norm


def convert_dir_vec_to_pose(vec):
    """host version of csrc/longform.cu dir_vec_to_pose (:81-102), for callers that hold numpy data"""
    vec = np.asarray(vec, dtype=np.float64)
    if vec.shape[-1] != 3:
        vec = vec.reshape(vec.shape[:-1] + (-1, 3))
    joint = np.zeros(vec.shape[:-2] + (10, 3))
    for j, (a, b, length) in enumerate(dir_vec_pairs):
        joint[..., b, :] = joint[..., a, :] + length * vec[..., j, :]
    return joint
