"""Bind the CUDA path over the reference's own callables (the boundary is a set of Python names).

    import sys; sys.path.insert(0, "/path/to/Semantic-SuperPoint")
    import ssp_b200; ssp_b200.dropin.install()
    # from here on `from utils.utils import descriptor_loss` (Train_model_heatmap_all.py:133), labels2Dto3D
    # (:278), inv_warp_image_batch (export.py:40), getPtsFromHeatmap (:698) ... resolve to the sm_100a kernels.

Most reference call sites import late (inside the function body), so patching module attributes before the
first training / export step covers them; `export.py:40` imports at module import, hence `install()` must run
before `import export`.  `install()` returns the list of names it bound; `uninstall()` restores the originals.
"""
import importlib
import sys

from . import utils as _u

# reference attribute -> replacement
UTILS_NAMES = [
    "warp_points", "filter_points", "inv_warp_image_batch", "inv_warp_image", "compute_valid_mask", "labels2Dto3D",
    "flattenDetection", "getPtsFromHeatmap", "nms_fast", "box_nms", "descriptor_loss", "normPts", "denormPts",
    "homography_scaling_torch",
]
_saved = []


def _bind(obj, name, fn, bound):
    if hasattr(obj, name):
        _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, fn)
    bound.append("%s.%s" % (getattr(obj, "__name__", type(obj).__name__), name))


def install(utils_module=None, trainer_class=None, frontend_class=None, export_module=None, tracker_class=None):
    """Patch `utils.utils` (imported from sys.path unless given) and, when passed, the trainer class
    (`Train_model_heatmap_all`: detector_loss, getMasks, sem_loss), the inference front-end class
    (`SuperPointFrontend_torch`: getPtsFromHeatmap, nms_fast, sample_desc_from_points), the tracker class
    (`PointTracker`: nn_match_two_way) and the `export` module (combine_heatmap)."""
    bound = []
    if utils_module is None:
        utils_module = importlib.import_module("utils.utils")
    for n in UTILS_NAMES:
        _bind(utils_module, n, getattr(_u, n), bound)
    if trainer_class is not None:
        _bind(trainer_class, "detector_loss",
              lambda self, input, target, mask=None, loss_type="softmax": _u.detector_loss(input, target, mask, loss_type), bound)
        _bind(trainer_class, "getMasks",
              lambda self, mask_2D, cell_size, device="cpu": _u.getMasks(mask_2D, cell_size, device), bound)
        _bind(trainer_class, "sem_loss", lambda self, pred, label, device="cpu": _u.sem_loss(pred, label, device), bound)
    if frontend_class is not None:
        _bind(frontend_class, "getPtsFromHeatmap",
              lambda self, heatmap: _u.getPtsFromHeatmap(heatmap, self.conf_thresh, self.nms_dist), bound)
        _bind(frontend_class, "nms_fast",
              lambda self, in_corners, H, W, dist_thresh: _u.nms_fast(in_corners, H, W, dist_thresh), bound)
        _bind(frontend_class, "sample_desc_from_points",
              lambda self, coarse_desc, pts: _u.sample_desc_from_points(coarse_desc, pts, self.cell), bound)
    if tracker_class is not None:  # models/model_wrap.py PointTracker

        def _nn(self, desc1, desc2, nn_thresh):
            self.mscores = _u.nn_match_two_way(desc1, desc2, nn_thresh)   # the reference keeps the matches here too
            return self.mscores

        _bind(tracker_class, "nn_match_two_way", _nn, bound)
    if export_module is None:
        export_module = sys.modules.get("export")
    if export_module is not None:
        _bind(export_module, "combine_heatmap", _u.combine_heatmap, bound)
        _bind(export_module, "inv_warp_image_batch", _u.inv_warp_image_batch, bound)
    return bound


def uninstall():
    while _saved:
        obj, name, fn = _saved.pop()
        setattr(obj, name, fn)
