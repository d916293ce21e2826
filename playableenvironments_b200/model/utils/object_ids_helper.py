"""ObjectIDsHelper (reference: model/utils/object_ids_helper.py:4-153): object instance <-> model index bookkeeping,
static models first."""
from typing import Tuple


class ObjectIDsHelper:

    def __init__(self, config):
        self.config = config
        m = config["model"]
        self.static_object_models_count = m["static_object_models"]
        self.object_models_count = len(m["object_models"])
        self.dynamic_object_models_count = self.object_models_count - self.static_object_models_count
        self.object_parameters_encoders_configs = m["object_parameters_encoder"]
        self.object_models_configs = m["object_models"]
        self.object_encoders_configs = m.get("object_encoders")
        self.model_idx_by_object_idx_map = {}
        self.model_idx_by_dynamic_object_idx_map = {}
        self.first_object_idx_by_model_idx_map = {}
        obj, dyn = 0, 0
        for model_idx in range(self.object_models_count):
            self.first_object_idx_by_model_idx_map[model_idx] = obj
            for _ in range(self.objects_count_by_model_idx(model_idx)):
                self.model_idx_by_object_idx_map[obj] = model_idx
                obj += 1
                if self.is_dynamic(model_idx):
                    self.model_idx_by_dynamic_object_idx_map[dyn] = model_idx
                    dyn += 1
        self.dynamic_objects_count = dyn
        self.objects_count = obj
        self.static_objects_count = obj - dyn

    def is_static(self, model_idx: int) -> bool:
        return model_idx < self.static_object_models_count

    def is_dynamic(self, model_idx: int) -> bool:
        return not self.is_static(model_idx)

    def objects_count_by_model_idx(self, model_idx: int) -> int:
        return self.object_parameters_encoders_configs[model_idx]["objects_count"]

    def objects_count_by_animation_model_idx(self, model_idx: int) -> int:
        return self.object_parameters_encoders_configs[self.static_object_models_count + model_idx]["objects_count"]

    def model_idx_by_object_idx(self, object_idx) -> int:
        return self.model_idx_by_object_idx_map[object_idx]

    def model_idx_by_dynamic_object_idx(self, object_idx) -> int:
        return self.model_idx_by_dynamic_object_idx_map[object_idx]

    def animation_model_idx_by_dynamic_object_idx(self, object_idx) -> int:
        return self.model_idx_by_dynamic_object_idx_map[object_idx] - self.model_idx_by_dynamic_object_idx_map[0]

    def object_idx_by_dynamic_object_idx(self, dynamic_object_idx) -> int:
        object_idx = dynamic_object_idx + self.static_objects_count
        if object_idx >= self.objects_count:
            raise Exception(f"The provided object id {dynamic_object_idx} is out of range")
        return object_idx

    def dynamic_object_idx_by_object_idx(self, object_idx) -> int:
        dynamic_object_id = object_idx - self.static_objects_count
        if dynamic_object_id < 0:
            raise Exception(f"The provided object id {object_idx} does not correspond to a dynamic object")
        return dynamic_object_id

    def dynamic_object_idx_range_by_model_idx(self, model_idx) -> Tuple[int, int]:
        if not self.is_dynamic(model_idx):
            raise Exception(f"Model id {model_idx} does not refer to a dynamic object")
        first = self.dynamic_object_idx_by_object_idx(self.first_object_idx_by_model_idx_map[model_idx])
        return first, first + self.objects_count_by_model_idx(model_idx)
