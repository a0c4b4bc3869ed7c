class SharedDict(dict):
    """accelerate.state.SharedDict: a dict shared between instances, attribute-style defaults."""

    def __getattr__(self, key):  # only reached for missing attributes
        raise AttributeError(key)
