"""Importable name for the package that lives in ``icp-slam-with-loop-closure_b200/``.

The directory name required by the repo layout contains hyphens and cannot be imported
directly; this alias package points its ``__path__`` there, so
``import icp_slam_b200.icp`` loads ``icp-slam-with-loop-closure_b200/icp.py``.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "icp-slam-with-loop-closure_b200")
__path__.insert(0, _impl)

from ._version import __version__  # noqa: E402,F401
