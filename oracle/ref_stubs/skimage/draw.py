"""skimage.draw.disk restated from the published scikit-image 0.18.1 algorithm (not vendored in
/root/reference; pinned by Dockerfile:69; call sites obstacle.py:78, fov.py:50-56).

disk(center, radius, shape) == ellipse(r, c, radius, radius, shape): bounding box
ceil(center-radius)..floor(center+radius) clipped to `shape`, keep cells with
((r-r0)/R)**2 + ((c-c0)/R)**2 < 1 evaluated in float64.  R == 0 gives an empty disk (nan/inf compare).
The reference's tests/learning/test_fov.py pins this indirectly (cost 30 + both golden FOVs).
"""
import numpy as np


def disk(center, radius, *, shape=None):
    center = np.array([center[0], center[1]], dtype=np.float64)
    radii = np.array([radius, radius], dtype=np.float64)
    upper_left = np.ceil(center - radii).astype(int)
    lower_right = np.floor(center + radii).astype(int)
    if shape is not None:
        upper_left = np.maximum(upper_left, np.array([0, 0]))
        lower_right = np.minimum(lower_right, np.array(shape[:2]) - 1)
    shifted_center = center - upper_left
    bounding_shape = lower_right - upper_left + 1
    if bounding_shape[0] <= 0 or bounding_shape[1] <= 0:
        return np.zeros(0, dtype=np.intp), np.zeros(0, dtype=np.intp)
    r_lim, c_lim = np.ogrid[0:float(bounding_shape[0]), 0:float(bounding_shape[1])]
    r, c = (r_lim - shifted_center[0]), (c_lim - shifted_center[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        distances = (r / radii[0]) ** 2 + (c / radii[1]) ** 2
    rr, cc = np.nonzero(distances < 1)
    return rr + upper_left[0], cc + upper_left[1]
