"""Entry of the hot path as the scene model sees it (reference: EnvironmentModel.batchified_composer_call and
merge_dictionaries, model/environment_model.py:474-545).

``install(environment_model)`` swaps the composer of an already-built reference ``EnvironmentModel`` (any of its
autoencoder-coupled subclasses) for the B200 one and rebinds ``batchified_composer_call`` — train.py, train_autoencoder.py and
play.py then run unchanged.  See INTEGRATION.md."""
import types
from typing import Dict, List

import torch

from .object_composer import ObjectComposer


def merge_dictionaries(dictionaries: List[Dict], dimension: int) -> Dict:
    """Reference :523-545 (drops the ``pytorch_hook`` dummy)."""
    merged = {}
    for key in list(dictionaries[0].keys()):
        if key == "pytorch_hook":
            continue
        if torch.is_tensor(dictionaries[0][key]):
            merged[key] = torch.cat([d[key] for d in dictionaries], dim=dimension)
        else:
            merged[key] = merge_dictionaries([d[key] for d in dictionaries], dimension)
    return merged


def batchified_composer_call(object_composer: ObjectComposer, ray_origins, ray_directions, focal_normals, transformation_matrix_w2o,
                             style, deformation, object_in_scene, perturb: bool, samples_per_image_batching: int = 0,
                             video_indexes=None, canonical_pose: bool = False) -> Dict:
    """Reference :474-521.  ``samples_per_image_batching`` exists there only to bound activation memory (12 sequential composer
    calls per 288x512 frame); the fused path keeps O(1) state per ray, so the argument is accepted and the frame is ONE call."""
    results = object_composer(ray_origins, ray_directions, focal_normals, transformation_matrix_w2o, style, deformation,
                              object_in_scene, perturb, video_indexes=video_indexes, canonical_pose=canonical_pose)
    results.pop("pytorch_hook", None)
    return results


RAY_SELECTION_FUNCTIONS = ("sample_rays", "sample_rays_weighted", "sample_rays_patched", "sample_rays_strided_patch",
                           "sample_all_rays_strided_grid", "permutation_indices_to_positions")


def install_ray_selection(reference_ray_helper=None):
    """Rebinds the reference's ``RayHelper.sample_rays*`` (utils/lib_3d/ray_helper.py:55-183, 236-482, 611-795; called from
    EnvironmentModel at model/environment_model.py:951-958, 1109-1116) to the tensorised versions of this package: same
    results for the same torch RNG state, but no ``.item()`` device->host sync per image and object."""
    import sys
    from ..utils.lib_3d.ray_helper import RayHelper as Ours
    if reference_ray_helper is None:
        module = sys.modules.get("utils.lib_3d.ray_helper")
        if module is None:
            raise Exception("the reference's utils.lib_3d.ray_helper is not imported: pass its RayHelper class explicitly")
        reference_ray_helper = module.RayHelper
    for name in RAY_SELECTION_FUNCTIONS:
        setattr(reference_ray_helper, name, staticmethod(getattr(Ours, name)))
    return reference_ray_helper


def install(environment_model, precision: str = None, ray_selection: bool = False):
    """Replaces ``environment_model.object_composer`` (a reference ObjectComposer) by the B200 composer carrying the same
    parameters, and routes ``batchified_composer_call`` to the single-call version.  ``ray_selection=True`` also rebinds
    the reference's ray-selection helpers (``install_ray_selection``).  ``precision=None`` keeps the configured
    ``model.b200_precision`` (default: the composer's own).  Note: the rebound ``batchified_composer_call`` ignores
    ``samples_per_image_batching``; in TRAIN mode the BatchNorm batch of the AdaIn layers is therefore the whole call instead of
    one chunk (eval mode is chunk-invariant, bit for bit)."""
    if ray_selection:
        install_ray_selection()
    reference = environment_model.object_composer
    config = dict(environment_model.config)
    composer = ObjectComposer(config)
    composer.load_state_dict(reference.state_dict())
    if precision is not None:
        composer.precision = precision
    composer.to(next(reference.parameters()).device)
    composer.train(reference.training)
    environment_model.object_composer = composer

    def _call(self, *args, **kwargs):
        return batchified_composer_call(self.object_composer, *args, **kwargs)

    environment_model.batchified_composer_call = types.MethodType(_call, environment_model)
    return environment_model


def build_environment_model(reference_architecture: str, config):
    """Builds the reference's environment model ``reference_architecture`` (a dotted module path of the upstream tree exporting
    ``model(config)``) and installs the B200 composer in it (``install``): the factory behind the
    ``playableenvironments_b200.model.environment_model_*`` architecture strings."""
    import importlib
    try:
        module = importlib.import_module(reference_architecture)
    except ImportError as exc:
        raise ImportError(f"{reference_architecture} is the reference's own module: put the PlayableEnvironments tree on PYTHONPATH "
                          f"(this package replaces its render path, not its encoders / decoder / trainers)") from exc
    return install(module.model(config))
