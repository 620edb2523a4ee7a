"""Drop-in shim: put this repo's root on PYTHONPATH ahead of the reference tree and the reference's
`from models import *` (main.py:3) picks up the B200 implementation -- main.py / functions.py run unchanged."""
from collision_handling_in_instantngp_b200.models import *  # noqa: F401,F403
from collision_handling_in_instantngp_b200.models import __all__  # noqa: F401
