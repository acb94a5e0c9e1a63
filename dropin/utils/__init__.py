"""`from utils.helpers import ...` (reference: utils/helpers.py).  The reference's utils/__init__.py also pulls in its
trainer (and swanlab); the sampling path only needs the helpers."""
from . import helpers  # noqa: F401
from .helpers import *  # noqa: F401,F403
