"""
The drop-in boundary.  When lhotse is importable, the B200 extractors subclass the real
``lhotse.features.base.FeatureExtractor`` (base.py:37) and register themselves in lhotse's own
``FEATURE_EXTRACTORS`` registry (base.py:368-405), so ``FeatureExtractor.from_dict/from_yaml``,
``CutSet.compute_and_store_features[_batch]`` and ``OnTheFlyFeatures`` work unchanged.

When lhotse is absent (e.g. the benchmark box), a local protocol class with the same public
methods is used instead: every reference caller is duck-typed on the extractor instance.
"""
from __future__ import annotations

from abc import ABCMeta, abstractmethod
from dataclasses import is_dataclass
from typing import Any, Dict, Optional, Type

import numpy as np

try:  # pragma: no cover - depends on the environment
    from lhotse.features.base import FEATURE_EXTRACTORS as _REGISTRY
    from lhotse.features.base import FeatureExtractor as FeatureExtractor
    from lhotse.features.base import register_extractor as register_extractor

    HAVE_LHOTSE = True
except Exception:  # lhotse (or one of its hard deps) is not importable
    HAVE_LHOTSE = False
    _REGISTRY: Dict[str, Type] = {}

    class FeatureExtractor(metaclass=ABCMeta):  # mirrors lhotse/features/base.py:37-366
        name = None
        config_type = None

        def __init__(self, config: Optional[Any] = None):
            if config is None:
                config = self.config_type()
            assert is_dataclass(config), "The feature configuration object must be a dataclass."
            self.config = config

        @abstractmethod
        def extract(self, samples: np.ndarray, sampling_rate: int) -> np.ndarray:
            ...

        @property
        @abstractmethod
        def frame_shift(self) -> float:
            ...

        @abstractmethod
        def feature_dim(self, sampling_rate: int) -> int:
            ...

        @property
        def device(self):
            return "cpu"

        @staticmethod
        def mix(features_a, features_b, energy_scaling_factor_b):
            raise ValueError('The feature extractor\'s "mix" operation is undefined.')

        @staticmethod
        def compute_energy(features):
            raise ValueError('The feature extractor\'s "compute_energy" operation is undefined.')

        @staticmethod
        def scale(features, energy_scaling_factor):
            raise ValueError('The feature extractor\'s "scale" operation is undefined.')

        @classmethod
        def from_dict(cls, data: dict) -> "FeatureExtractor":
            data = dict(data)
            extractor_type = _REGISTRY[data.pop("feature_type")]
            return extractor_type(extractor_type.config_type.from_dict(data))

        def to_dict(self) -> Dict[str, Any]:
            d = self.config.to_dict()
            d["feature_type"] = self.name
            return d

        @classmethod
        def from_yaml(cls, path) -> "FeatureExtractor":
            import yaml

            with open(path) as f:
                return cls.from_dict(yaml.safe_load(f))

        def to_yaml(self, path):
            import yaml

            data = self.to_dict()
            if "device" in data and not isinstance(data["device"], str):
                data["device"] = data["device"].type
            with open(path, "w") as f:
                yaml.safe_dump(data, f)

    def register_extractor(cls):
        _REGISTRY[cls.name] = cls
        return cls


def get_extractor_type(name: str) -> Type:
    return _REGISTRY[name]
