"""B200-native GDM hot path: host-side mirror of the reference's include/gdm API over libgdm_b200.so."""
from .api import *  # noqa: F401,F403
