from .._overlay import extend

__path__ = extend(list(__path__), "util")
