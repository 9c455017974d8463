from nerf_downstream_b200.me.functional import *  # noqa: F401,F403
