"""Minimal configuration-property system.

The reference derives every configurable class from ``caput.config.Reader``
and declares ``config.Property`` class attributes that are filled from the
YAML sections (drift/core/telescope.py:211-243, drift/core/beamtransfer.py:
186-195, drift/telescope/cylinder.py:30-46).  ``caput`` is an external
dependency that is not available offline, so the same behaviour is provided
here: unknown keys are ignored, ``key=`` renames the YAML key, ``proptype``
coerces the value.
"""


class Property:
    """A descriptor whose value can be set from a config dictionary."""

    def __init__(self, default=None, proptype=None, key=None):
        self.default = default
        self.proptype = (lambda x: x) if proptype is None else proptype
        self.key = key
        self.propname = None

    def __set_name__(self, owner, name):
        self.propname = name
        if self.key is None:
            self.key = name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return obj.__dict__.get("_prop_" + self.propname, self.default)

    def __set__(self, obj, val):
        obj.__dict__["_prop_" + self.propname] = None if val is None else self.proptype(val)

    def _from_config(self, obj, config):
        if self.key in config:
            self.__set__(obj, config[self.key])


def enum(options, default=None):
    """A property restricted to a fixed set of values."""

    def _prop(val):
        if val not in options:
            raise ValueError(f"Input {val!r} is not one of the options {options!r}")
        return val

    if default is not None and default not in options:
        raise ValueError(f"Default {default!r} is not one of the options {options!r}")
    return Property(proptype=_prop, default=default)


def list_type(type_=None, length=None, maxlength=None, default=None):
    """A property holding a list with optional element type / length checks."""

    def _prop(val):
        if not isinstance(val, (list, tuple)):
            raise ValueError("Expected a list")
        if type_ is not None:
            val = [type_(v) for v in val]
        if length is not None and len(val) != length:
            raise ValueError(f"List must have length {length}")
        if maxlength is not None and len(val) > maxlength:
            raise ValueError(f"List must not be longer than {maxlength}")
        return list(val)

    return Property(proptype=_prop, default=default)


class Reader:
    """Base class for objects configurable from a dictionary."""

    @classmethod
    def from_config(cls, config, *args, **kwargs):
        c = cls(*args, **kwargs)
        c.read_config(config)
        return c

    def read_config(self, config):
        for basecls in type(self).__mro__[::-1]:
            for propval in vars(basecls).values():
                if isinstance(propval, Property):
                    propval._from_config(self, config)
        self._finalise_config()

    def _finalise_config(self):
        pass

    def __getstate__(self):
        return self.__dict__.copy()
