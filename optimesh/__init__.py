"""Drop-in alias: ``import optimesh`` resolves to the B200-native implementation of the
smoothing path (/root/reference/README.md:119-142)."""
from optimesh_b200 import *  # noqa: F401,F403
from optimesh_b200 import __version__, cpt, cvt, odt  # noqa: F401
from optimesh_b200 import cli  # noqa: F401
