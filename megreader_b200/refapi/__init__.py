"""The reference's own import paths, backed by megreader_b200.

After `install()`, `import ops`, `import decoders`, `import backbones`, `import assets.ops.dcn`
resolve to the packages in this directory, which mirror the reference's module layout, class
names, constructor signatures and state-dict keys for the hot path, so that
`getattr(decoders, name)` / the yaml class lookup of the reference's host code
(structure/model.py:20-21, concern/config.py:77-90) find them.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    return _HERE
