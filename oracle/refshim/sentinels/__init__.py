"""Minimal stand-in for the ``sentinels`` package.  TEST INFRASTRUCTURE ONLY."""


class Sentinel(object):
    _existing = {}

    def __new__(cls, name):
        if name not in cls._existing:
            inst = super(Sentinel, cls).__new__(cls)
            inst._name = name
            cls._existing[name] = inst
        return cls._existing[name]

    def __repr__(self):
        return "<{}>".format(self._name)

    def __reduce__(self):
        return (Sentinel, (self._name,))
