"""Shared machinery of the model parameter objects: a declarative (name, default) table per
model, keyword / positional construction with the reference's names and order, attribute access
and the ``.params`` tuple in the order the reference's operators unpack it."""
import unicodedata


def _norm(name):
    # Python identifiers are NFKC-normalised (ϕ U+03D5 and φ U+03C6 name the same variable)
    return unicodedata.normalize("NFKC", name)


class ParameterSet:
    _TABLE = ()          # ((name, default), ...) in constructor order
    _PARAMS_ORDER = ()   # names, in the order of the .params tuple

    def __init__(self, *args, **kwargs):
        names = [_norm(n) for n, _ in self._TABLE]
        if len(args) > len(names):
            raise TypeError(f"{type(self).__name__}() takes at most {len(names)} positional arguments")
        values = {_norm(n): v for n, v in self._TABLE}
        for n, v in zip(names, args):
            values[n] = v
        for k, v in kwargs.items():
            k = _norm(k)
            if k not in values:
                raise TypeError(f"{type(self).__name__}() got an unexpected keyword argument {k!r}")
            if k in names[:len(args)]:
                raise TypeError(f"{type(self).__name__}() got multiple values for argument {k!r}")
            values[k] = v
        for n, v in values.items():
            setattr(self, n, v)
        self.params = tuple(values[_norm(n)] for n in self._PARAMS_ORDER)

    def __repr__(self):
        inner = ", ".join(f"{_norm(n)}={getattr(self, _norm(n))!r}" for n, _ in self._TABLE)
        return f"{type(self).__name__}({inner})"
