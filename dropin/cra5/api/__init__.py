"""cra5/api/__init__.py:1-3 of the reference: `from .cra5_api import cra5_api`"""
from cra5_b200.api.cra5_api import cra5_api

__all__ = ["cra5_api"]
