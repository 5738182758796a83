"""Importable alias of the product package.

The product lives in ``deeprank-gnn_b200/`` (the directory name the project layout fixes);
a hyphen cannot appear in a Python module name, so this shim makes the same directory
importable as ``deeprank_gnn_b200`` by pointing ``__path__`` at it and executing its
``__init__.py`` in this namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'deeprank-gnn_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
