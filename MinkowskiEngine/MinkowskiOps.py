from nerf_downstream_b200.me.modules import cat  # noqa: F401
