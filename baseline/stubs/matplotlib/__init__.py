"""Import stub (TEST / BASELINE INFRASTRUCTURE): the reference imports matplotlib at module import
(src/envs/ssd/map_env.py:6-8, src/utils/utility_funcs.py:3, src/controllers/homophily_controller.py:7) but only
uses it for replay rendering, which is out of scope.  matplotlib is not installed in this image."""
from . import pyplot, patches  # noqa: F401
