from cra5_b200.api.cra5_api import *  # noqa: F401,F403
from cra5_b200.api.cra5_api import cra5_api  # noqa: F401
