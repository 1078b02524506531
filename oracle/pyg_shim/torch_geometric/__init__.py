"""ORACLE / TEST INFRASTRUCTURE ONLY.  Minimal ``torch_geometric`` stand-in (restated 2.5.3
semantics, CPU, plain PyTorch) so the unmodified reference imports.  See nn/ and utils/."""
__version__ = "2.5.3+oracle-restatement"
from . import nn, utils  # noqa: F401,E402
