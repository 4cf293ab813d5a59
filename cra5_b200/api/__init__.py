from .cra5_api import cra5_api  # noqa: F401
