"""Stand-in for ``omegaconf`` (absent here): attribute dicts are enough for the closures."""


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


class OmegaConf:
    @staticmethod
    def to_container(cfg):
        return dict(cfg)
