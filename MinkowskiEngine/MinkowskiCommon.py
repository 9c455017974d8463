from nerf_downstream_b200.me.modules import MinkowskiModuleBase, MinkowskiNetwork  # noqa: F401
