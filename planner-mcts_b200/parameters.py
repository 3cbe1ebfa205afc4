"""Hierarchical parameter tree with the access conventions the reference's Python tooling is
written against (BARK's `bark.runtime.commons.parameters.ParameterServer`, not in the
container; behaviour restated from its call sites in
hypothesis/behavior_space/behavior_space.py:19-343 and tests/py_behavior_space_test.py:29-250):

    p = ParameterServer()
    p["BehaviorSpace"]["Sampling"]["RandomSeed", "seed of the sampler", 1000]  # -> 1000, stored
    p.AddChild("A::B")                       # children along a "::" path
    p.getReal("A::B::x", "", 0.0)            # the C++ Params getters on the same tree
    p.flatten()                              # {"A::B::x": value}: what the planners' C ABI takes

A missing plain key creates an empty child (so chained `p["a"]["b"]` works on a fresh tree); a
`(key, description, default)` triple stores and returns the default when the key is missing.
"""
import copy
import json as _json

DELIMITER = "::"


class ParameterServer:
    def __init__(self, filename=None, json=None, log_if_default=False, **_ignored):
        self.store = {}
        self.param_descriptions = {}
        self.log_if_default = log_if_default
        self.param_filename = filename
        if filename is not None:
            with open(filename) as f:
                self._load(_json.load(f))
        elif json is not None:
            self._load(json)

    # ------------------------------------------------------------------ tree <-> dict
    def _load(self, dct):
        for key, val in dct.items():
            if isinstance(val, dict):
                child = ParameterServer(log_if_default=self.log_if_default)
                child._load(val)
                self.store[key] = child
            else:
                self.store[key] = copy.deepcopy(val)

    def ConvertToDict(self, print_description=False):
        out = {}
        for key, val in self.store.items():
            if isinstance(val, ParameterServer):
                out[key] = val.ConvertToDict(print_description)
            elif print_description and key in self.param_descriptions:
                out[key] = {"value": val, "description": self.param_descriptions[key]}
            else:
                out[key] = val
        return out

    def Save(self, filename, print_description=False):
        with open(filename, "w") as f:
            _json.dump(self.ConvertToDict(print_description), f, indent=4, default=_jsonable)

    def clone(self):
        return ParameterServer(json=self.ConvertToDict(), log_if_default=self.log_if_default)

    def flatten(self, prefix=""):
        """{"A::B::leaf": value} over all leaves (the key form of the C++ Params getters)"""
        out = {}
        for key, val in self.store.items():
            path = prefix + key
            if isinstance(val, ParameterServer):
                out.update(val.flatten(path + DELIMITER))
            else:
                out[path] = val
        return out

    # ------------------------------------------------------------------ item access
    def _descend(self, path, create):
        node = self
        for part in path:
            nxt = node.store.get(part)
            if not isinstance(nxt, ParameterServer):
                if not create:
                    return None
                nxt = ParameterServer(log_if_default=self.log_if_default)
                node.store[part] = nxt
            node = nxt
        return node

    def __getitem__(self, key):
        if isinstance(key, tuple):
            name, description, default = key
            parts = name.split(DELIMITER)
            node = self._descend(parts[:-1], create=True)
            leaf = parts[-1]
            node.param_descriptions[leaf] = description
            if leaf not in node.store:
                node.store[leaf] = default
            return node.store[leaf]
        if key not in self.store:
            self.store[key] = ParameterServer(log_if_default=self.log_if_default)
        return self.store[key]

    def __setitem__(self, key, value):
        if isinstance(value, dict):
            value = ParameterServer(json=value, log_if_default=self.log_if_default)
        self.store[key] = value

    def __delitem__(self, key):
        del self.store[key]

    def __contains__(self, key):
        return key in self.store

    def __len__(self):
        return len(self.store)

    def __iter__(self):
        return iter(self.store)

    def items(self):
        """flat (path, value) pairs, so a tree can stand in wherever a flat dict is taken"""
        return self.flatten().items()

    def get(self, path, default=None):
        """read-only lookup of a "::" path (also finds keys that contain "::" themselves)"""
        if path in self.store and not isinstance(self.store[path], ParameterServer):
            return self.store[path]
        flat = self.flatten()
        return flat.get(path, default)

    def AddChild(self, name, delete=False):
        parts = name.split(DELIMITER)
        parent = self._descend(parts[:-1], create=True)
        if delete or not isinstance(parent.store.get(parts[-1]), ParameterServer):
            parent.store[parts[-1]] = ParameterServer(log_if_default=self.log_if_default)
        return parent.store[parts[-1]]

    # ------------------------------------------------------------------ C++ Params getters
    def _typed(self, name, description, default, conv):
        return conv(self[name, description, default])

    def getReal(self, name, description="", default=0.0):
        return self._typed(name, description, default, float)

    def getInt(self, name, description="", default=0):
        return self._typed(name, description, default, int)

    def getBool(self, name, description="", default=False):
        return self._typed(name, description, default, bool)

    def getString(self, name, description="", default=""):
        return self._typed(name, description, default, str)

    def getListFloat(self, name, description="", default=()):
        return [float(v) for v in self[name, description, list(default)]]

    def getListListFloat(self, name, description="", default=()):
        return [[float(v) for v in row] for row in self[name, description, list(default)]]

    GetReal, GetInt, GetBool, GetString = getReal, getInt, getBool, getString
    GetListFloat, GetListListFloat = getListFloat, getListListFloat

    def __repr__(self):
        return f"ParameterServer({self.ConvertToDict()!r})"

    # pickling: the tree as a plain dict
    def __getstate__(self):
        return {"tree": self.ConvertToDict(), "log_if_default": self.log_if_default}

    def __setstate__(self, st):
        self.__init__(json=st["tree"], log_if_default=st["log_if_default"])


def _jsonable(x):
    try:
        return x.item()  # numpy scalars
    except AttributeError:
        return list(x)
