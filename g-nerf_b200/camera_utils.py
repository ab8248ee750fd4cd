"""Host-side camera helpers for the frame loop that feeds the renderer (tiny, numpy, not on the hot path).

The reference builds every pose on the device with ~20 small torch kernels per frame
(camera_utils.py:88-106,155-174).  A 120-frame orbit is 120 x 25 floats:
build it once on the host and upload it in one copy.
"""
import math

import numpy as np

FFHQ_FOCAL = 4.2647                      # gen_videos.py:135


def look_at_origin(theta: float, phi: float, radius: float) -> np.ndarray:
    """cam2world [4,4] of a camera on the sphere of ``radius`` at yaw ``theta`` / pitch ``phi`` looking at the
    origin, y up, no roll -- the pose LookAtPoseSampler.sample returns (camera_utils.py:88-106,155-174)."""
    eye = radius * np.array([math.sin(phi) * math.cos(math.pi - theta), math.cos(phi),
                             math.sin(phi) * math.sin(math.pi - theta)])
    z = -eye / np.linalg.norm(eye)
    x = np.cross(z, (0.0, 1.0, 0.0))           # = -cross(up, forward)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    y /= np.linalg.norm(y)
    pose = np.eye(4)
    pose[:3, :3] = np.stack([x, y, z], axis=1)
    pose[:3, 3] = eye
    return pose.astype(np.float32)


def ffhq_intrinsics(n: int) -> np.ndarray:
    k = np.array([[FFHQ_FOCAL, 0, 0.5], [0, FFHQ_FOCAL, 0.5], [0, 0, 1]], np.float32)
    return np.broadcast_to(k, (n, 3, 3)).copy()


def orbit_cameras(n: int, radius: float = 2.7, frames: int = 120, pitch_offset: float = -0.05):
    """``n`` cameras evenly spaced along the gen_videos orbit (gen_videos.py:155-158, which spells pi as 3.14):
    frame i has yaw 3.14/2 + 0.7 sin(2*3.14*i/frames) and pitch 3.14/2 + pitch_offset + 0.3 cos(2*3.14*i/frames).
    Returns (cam2world [n,4,4], intrinsics [n,3,3]) as float32 numpy arrays."""
    poses = []
    for j in range(n):
        i = (j * frames) // max(n, 1)
        poses.append(look_at_origin(3.14 / 2 + 0.7 * math.sin(2 * 3.14 * i / frames),
                                    3.14 / 2 + pitch_offset + 0.3 * math.cos(2 * 3.14 * i / frames), radius))
    return np.stack(poses), ffhq_intrinsics(n)
