"""Name -> class maps for encoder wrappers, mirroring the reference's plugin discovery
(A/my_affectgpt/common/registry.py:143-172, lookups :282-310): ``register_visual_encoder``,
``register_acoustic_encoder``, ``register_model`` and their ``get_*_class`` twins.  Registering a name twice
raises KeyError, exactly as the reference does (registry.py:147-152)."""
from __future__ import annotations


class Registry:
    mapping = {"visual_encoder_mapping": {}, "acoustic_encoder_mapping": {}, "model_name_mapping": {}}

    @classmethod
    def _register(cls, table: str, name: str):
        def wrap(klass):
            if name in cls.mapping[table]:
                raise KeyError("Name '{}' already registered for {}.".format(name, cls.mapping[table][name]))
            cls.mapping[table][name] = klass
            return klass
        return wrap

    @classmethod
    def register_visual_encoder(cls, name):
        return cls._register("visual_encoder_mapping", name)

    @classmethod
    def register_acoustic_encoder(cls, name):
        return cls._register("acoustic_encoder_mapping", name)

    @classmethod
    def register_model(cls, name):
        return cls._register("model_name_mapping", name)

    @classmethod
    def get_visual_encoder_class(cls, name):
        return cls.mapping["visual_encoder_mapping"].get(name, None)

    @classmethod
    def get_acoustic_encoder_class(cls, name):
        return cls.mapping["acoustic_encoder_mapping"].get(name, None)

    @classmethod
    def get_model_class(cls, name):
        return cls.mapping["model_name_mapping"].get(name, None)

    @classmethod
    def list_visual_encoders(cls):
        return sorted(cls.mapping["visual_encoder_mapping"].keys())

    @classmethod
    def list_acoustic_encoders(cls):
        return sorted(cls.mapping["acoustic_encoder_mapping"].keys())


registry = Registry()
