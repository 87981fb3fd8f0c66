"""A small stand-in for ``ml_collections.ConfigDict`` (absent from this image): attribute and
item access, nested dict promotion, ``update``, ``to_dict``, ``update_from_flattened_dict``."""


class ConfigDict:
    def __init__(self, initial=None, **kwargs):
        object.__setattr__(self, "_d", {})
        for src in (initial, kwargs):
            if src:
                for k, v in (src.items() if hasattr(src, "items") else src):
                    self[k] = v

    # mapping protocol ---------------------------------------------------------------------
    def __setitem__(self, key, value):
        self._d[key] = ConfigDict(value) if isinstance(value, dict) else value

    def __getitem__(self, key):
        return self._d[key]

    def __delitem__(self, key):
        del self._d[key]

    def __contains__(self, key):
        return key in self._d

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def keys(self):
        return self._d.keys()

    def values(self):
        return self._d.values()

    def items(self):
        return self._d.items()

    def get(self, key, default=None):
        return self._d.get(key, default)

    def pop(self, key, *default):
        return self._d.pop(key, *default)

    # attribute protocol -------------------------------------------------------------------
    def __getattr__(self, name):
        d = object.__getattribute__(self, "_d")
        if name in d:
            return d[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    # bulk operations ----------------------------------------------------------------------
    def update(self, *others, **kwargs):
        for other in others + (kwargs,):
            for k, v in (other.items() if hasattr(other, "items") else other):
                cur = self._d.get(k)
                if isinstance(cur, ConfigDict) and isinstance(v, (dict, ConfigDict)):
                    cur.update(v)
                else:
                    self[k] = v

    def update_from_flattened_dict(self, flat, strip_prefix=""):
        for dotted, v in flat.items():
            node, parts = self, dotted[len(strip_prefix):].split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def to_dict(self):
        return {k: v.to_dict() if isinstance(v, ConfigDict) else v for k, v in self._d.items()}

    def __repr__(self):
        return f"ConfigDict({self._d})"
